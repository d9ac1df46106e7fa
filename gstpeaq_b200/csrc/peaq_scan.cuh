// Device code of the recurrent half of basic-mode PEAQ (`scan_step`, `scan_epilogue`), shared by
// K2 `scan_basic_kernel` (peaq_scan.cu: state in registers across the frame loop, inputs from the
// per-frame records) and the fused persistent kernel (peaq_fused.cu: inputs from shared memory,
// state in L2-resident global memory between frames).
//
// One CTA = one (ref,test) pair; C groups of 128 threads, thread (c, b) owns band
// b of channel c.  Per frame, in order,
//   - time-domain smearing of the excitation      (fftearmodel.c:496-504)
//   - level and pattern adaptation                (leveladapter.c:243-340)
//   - modulation processing                       (modpatt.c:223-251)
//   - the loudness-reached latch                  (gstpeaq.c:841-845, earmodel.c:890-907)
// then the band terms of modulation difference (movs.c:205-254), noise loudness
// (movs.c:709-743), NMR (movs.c:1002-1021) and detection probability (movs.c:1234-1275),
// reduced over bands with warp shuffles and fed to the 11 MOV accumulators with the reference's
// INIT / NORMAL / TENTATIVE semantics (movaccum.c:317-481; call order of gstpeaq.c:850-921).
// The epilogue evaluates the accumulators, the 11-3-1 network and the ODG mapping
// (nn.c:187-216, :372-375).
//
// Compiled with -fmad=false: expressions round like the reference's C code.
// Powers x^y are evaluated as exp(y ln x) through peaq_math.cuh (integer powers by
// multiplication, 0.5^y as exp(-y ln 2)): ~1e-15 relative deviation from libm's pow.
#pragma once

#include "peaq_engine.h"
#include "peaq_math.cuh"

#include <climits>

namespace peaq {
namespace {

constexpr int kGroup = 128;     // threads per channel group
constexpr int kWarpsPerGroup = 4;

// accumulator slots = MOV order of gstpeaq.c:95-108
enum {
  kBwRef, kBwTest, kTotalNmr, kWinModDiff, kAdb, kEhs, kAvgModDiff1, kAvgModDiff2,
  kRmsNoiseLoud, kMfpd, kRelDistFrames
};
enum { kStInit = 0, kStNormal = 1, kStTentative = 2 };   // movaccum.c:53-58
enum { kModeAvg, kModeAvgLog, kModeRms, kModeAvgWindow, kModeFilteredMax, kModeAdb };

__device__ __forceinline__ int acc_mode_basic(int k) {
  switch (k) {   // gstpeaq.c:536-557
    case kTotalNmr: return kModeAvgLog;
    case kWinModDiff: return kModeAvgWindow;
    case kAdb: return kModeAdb;
    case kRmsNoiseLoud: return kModeRms;
    case kMfpd: return kModeFilteredMax;
    default: return kModeAvg;
  }
}

struct Acc {
  double num, den, x0, x1, x2, snum, sden, smax;   // x0..2: window history | (max, filt, -)
};

// peaq_movaccum_accumulate (movaccum.c:369-425); INIT is handled by the caller
__device__ __forceinline__ void acc_add(Acc& a, int mode, double val, double weight) {
  switch (mode) {
    case kModeRms:
      weight *= weight;
      a.num += weight * val * val;
      a.den += weight;
      break;
    case kModeAvg:
    case kModeAvgLog:
    case kModeAdb:
      a.num += weight * val;
      a.den += weight;
      break;
    case kModeAvgWindow: {
      const double val_sqrt = sqrt(val);
      if (!isnan(a.x0)) {
        double winsum = val_sqrt;
        winsum += a.x0;
        winsum += a.x1;
        winsum += a.x2;
        winsum /= 4.;
        winsum *= winsum;
        winsum *= winsum;
        a.num += winsum;
        a.den += 1.;
      }
      a.x0 = a.x1;
      a.x1 = a.x2;
      a.x2 = val_sqrt;
      break;
    }
    case kModeFilteredMax:
      a.x1 = 0.9 * a.x1 + 0.1 * val;   // x1 = filter state, x0 = max
      if (a.x1 > a.x0) a.x0 = a.x1;
      break;
  }
}

// one channel's term of peaq_movaccum_get_value (movaccum.c:438-481)
__device__ __forceinline__ double acc_value(const Acc& a, int mode, bool tentative) {
  const double num = tentative ? a.snum : a.num;
  const double den = tentative ? a.sden : a.den;
  switch (mode) {
    case kModeAvg: return num / den;
    case kModeAvgLog: return 10. * log10(num / den);
    case kModeAvgWindow:
    case kModeRms: return sqrt(num / den);
    case kModeFilteredMax: return tentative ? a.smax : a.x0;
    case kModeAdb:
      if (den > 0) return num == 0. ? -0.5 : log10(num / den);
      return 0.;
  }
  return 0.;
}

#ifndef PEAQ_WARP_SUM_DEFINED
#define PEAQ_WARP_SUM_DEFINED
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif
__device__ __forceinline__ double warp_prod(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max0(double v) {
  // max that ignores NaN candidates the way `if (x > m) m = x` does
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double x = __shfl_xor_sync(0xffffffffu, v, o);
    v = x > v ? x : v;
  }
  return v;
}

constexpr int kRed1 = 4;   // num, den, loudness ref, loudness test
constexpr int kRed2 = 8;   // md1, md2, wt, nl, nmr sum, nmr max, prod(1-p), sum q

// shared-memory work arrays of one scan CTA
struct ScanShared {
  double red1[2][kMaxChannels][kWarpsPerGroup][kRed1];
  double red2[2][kMaxChannels][kWarpsPerGroup][kRed2];
  double pa[kMaxChannels][2][kGroup];        // pattern adaptation factors
  double pq[kMaxChannels][2][kGroup];        // detection probability / steps
  double movval[kMaxChannels][kNumAcc];
  double movs_sh[kNumAcc];
  int latch_sh;
};

// per-band constants of thread (c, b)
struct BandConst {
  double a_ear, a_proc, in_noise, in03, ethres, thres, loudfac, maskdiff;
  int m1, m2;                                   // leveladapter.c:318-319
};
__device__ __forceinline__ BandConst load_band_const(const DeviceTables* __restrict__ T, int bb, int B) {
  BandConst k;
  k.a_ear = T->fft.a_ear[bb];
  k.a_proc = T->fft.a_proc[bb];
  k.in_noise = T->fft.internal_noise[bb];
  k.in03 = T->fft.internal_noise_pow03[bb];
  k.ethres = T->fft.ethres[bb];
  k.thres = T->fft.thres[bb];
  k.loudfac = T->fft.loudfac[bb];
  k.maskdiff = T->maskdiff[bb];
  k.m1 = min(bb, B / 36);
  k.m2 = min(B - bb - 1, B / 25);
  return k;
}

// what one frame hands to thread (c, b): its band's unsmeared excitations and noise, and the
// frame's scalars (only the accumulator threads look at bw / ehs)
struct ScanInputs {
  double E2r, E2t, nz;
  int flags;
  const int* bw;        // {bw_ref, bw_test} of the thread's channel
  const double* ehs;    // EHS value of the thread's channel
  const double* snr;    // {signal, noise} energy of the frame
  // test tap (keep_records; K2 only) or null: this frame's slot, [2][C][B] smeared excitations
  // (ref | test) followed by kScanTapValues values per channel, see scan_tap_doubles()
  double* dbg;
};
constexpr int kScanTapValues = 8;   // md1, md2, temporal weight, noise loudness, mean N/M, max N/M, P binaural, Q binaural
__host__ __device__ inline int scan_tap_doubles(int C, int B) { return 2 * C * B + C * kScanTapValues; }

// status word and counters every thread tracks
struct ScanCounters {
  int status;
  unsigned frame_counter, loud_frame;
  double sig_energy, noise_energy;
  // segments of long items (peaq_segments.cu; all zero for whole items): sums restart at frame
  // acc_start (the frames before it only warm the recurrences up), `owned_above`: a frame from
  // acc_start on was above the threshold, first_above: 1 + the first frame above it (0: none)
  unsigned acc_start, first_above;
  int owned_above;
};

// thread identity inside the scan CTA
struct ScanThread {
  int c, b, bb, lane, wig;
  bool active, acc_thread;
  int acc_mode;
};
__device__ __forceinline__ ScanThread make_scan_thread(int B) {
  ScanThread th;
  th.c = threadIdx.x / kGroup;
  th.b = threadIdx.x % kGroup;
  th.lane = threadIdx.x & 31;
  th.wig = (threadIdx.x >> 5) % kWarpsPerGroup;   // warp in group
  th.active = th.b < B;
  th.bb = th.active ? th.b : 0;
  th.acc_thread = th.b < kNumAcc;
  th.acc_mode = acc_mode_basic(th.b < kNumAcc ? th.b : 0);
  return th;
}

// How scan_step reaches the 14 recurrent band values of thread (c, b) -- field order of StateLayout:
// 0 Efr, 1 Eft (time smearing) | 2 Rf, 3 Tf, 4 fnum, 5 fden, 6 pcr, 7 pct (level adapter) |
// 8 prev_r, 9 fl_r, 10 fd_r, 11 prev_t, 12 fl_t, 13 fd_t (modulation).
// RegState: registers across the frame loop (K2).  MemState: global memory (L2), each value loaded
// where it is first needed and stored right after its update, so that the fused kernel's scan half
// fits the 80 registers of its frame half.
struct RegState {
  double (&bs)[kBandStateFields];
  __device__ __forceinline__ double get(int i) const { return bs[i]; }
  __device__ __forceinline__ void set(int i, double v) const { bs[i] = v; }
};
struct MemState {
  double* base;      // &state[off_band + c * B + b]
  int stride;        // C * B
  bool active;
  __device__ __forceinline__ double get(int i) const { return active ? __ldcg(base + i * stride) : 0.; }
  __device__ __forceinline__ void set(int i, double v) const { if (active) __stcg(base + i * stride, v); }
};
// the thread's accumulator: registers across the frame loop (K2) or its slot in the state block (fused)
struct RegAcc {
  __device__ __forceinline__ void load(Acc&) const {}
  __device__ __forceinline__ void store(const Acc&) const {}
};
struct MemAcc {
  double* slot;
  __device__ __forceinline__ void load(Acc& a) const {
    a.num = slot[0]; a.den = slot[1]; a.x0 = slot[2]; a.x1 = slot[3]; a.x2 = slot[4];
    a.snum = slot[5]; a.sden = slot[6]; a.smax = slot[7];
  }
  __device__ __forceinline__ void store(const Acc& a) const {
    slot[0] = a.num; slot[1] = a.den; slot[2] = a.x0; slot[3] = a.x1; slot[4] = a.x2;
    slot[5] = a.snum; slot[6] = a.sden; slot[7] = a.smax;
  }
};
// per-band constants: preloaded (K2) or read from the tables where needed (fused)
struct RegConst {
  const BandConst& k;
  __device__ __forceinline__ double a_ear() const { return k.a_ear; }
  __device__ __forceinline__ double a_proc() const { return k.a_proc; }
  __device__ __forceinline__ double in_noise() const { return k.in_noise; }
  __device__ __forceinline__ double in03() const { return k.in03; }
  __device__ __forceinline__ double ethres() const { return k.ethres; }
  __device__ __forceinline__ double thres() const { return k.thres; }
  __device__ __forceinline__ double loudfac() const { return k.loudfac; }
  __device__ __forceinline__ double maskdiff() const { return k.maskdiff; }
};
struct MemConst {
  const DeviceTables* __restrict__ T;
  int bb;
  __device__ __forceinline__ double a_ear() const { return T->fft.a_ear[bb]; }
  __device__ __forceinline__ double a_proc() const { return T->fft.a_proc[bb]; }
  __device__ __forceinline__ double in_noise() const { return T->fft.internal_noise[bb]; }
  __device__ __forceinline__ double in03() const { return T->fft.internal_noise_pow03[bb]; }
  __device__ __forceinline__ double ethres() const { return T->fft.ethres[bb]; }
  __device__ __forceinline__ double thres() const { return T->fft.thres[bb]; }
  __device__ __forceinline__ double loudfac() const { return T->fft.loudfac[bb]; }
  __device__ __forceinline__ double maskdiff() const { return T->maskdiff[bb]; }
};

// One frame of the scan for thread (c, b).  st: the 14 recurrent band values; acc: the
// accumulator owned by threads b < kNumAcc.  Contains three CTA-wide barriers.
template <typename State, typename Konst, typename AccIo>
__device__ __forceinline__ void scan_step(const ScanInputs& in, const Konst& kc, const State& st, const AccIo& acc_io,
                                          Acc& acc, ScanCounters& cnt, ScanShared& sh, const ScanThread& th,
                                          int C, int B) {
  const int c = th.c, b = th.b, bb = th.bb, lane = th.lane, wig = th.wig;
  const bool active = th.active, acc_thread = th.acc_thread;
  const int acc_mode = th.acc_mode;
  const double deriv_factor = (double)48000 / kFftStep;
  const int m1 = min(bb, B / 36), m2 = min(B - bb - 1, B / 25);   // leveladapter.c:318-319
  int& status = cnt.status;
  unsigned& frame_counter = cnt.frame_counter;
  unsigned& loud_frame = cnt.loud_frame;
  const int par = frame_counter & 1;
  const double E2r = in.E2r, E2t = in.E2t, nz = in.nz;
  const int flags = in.flags;
  const bool above = flags & kRecFlagAbove;
  auto& red1 = sh.red1;
  auto& red2 = sh.red2;
  auto& pa = sh.pa;
  auto& pq = sh.pq;
  int& latch_sh = sh.latch_sh;
  const double a_proc = kc.a_proc();

  // time-domain smearing (fftearmodel.c:496-504)
  const double a_ear = kc.a_ear();
  const double Efr = a_ear * st.get(0) + (1. - a_ear) * E2r;
  st.set(0, Efr);
  const double Er = Efr > E2r ? Efr : E2r;
  const double Eft = a_ear * st.get(1) + (1. - a_ear) * E2t;
  st.set(1, Eft);
  const double Et = Eft > E2t ? Eft : E2t;
  if (in.dbg && active) {
    in.dbg[(0 * C + c) * B + b] = Er;
    in.dbg[(1 * C + c) * B + b] = Et;
  }

  // modulation (modpatt.c:234-250): needs nothing but the unsmeared excitations, so it sits in
  // this phase, where its exp/log chain overlaps the detection-probability chain below
  const double Lr = peaq_exp_clamped(0.3 * peaq_log_pos(E2r)), Lt = peaq_exp_clamped(0.3 * peaq_log_pos(E2t));
  const double fd_r = a_proc * st.get(10) + (1 - a_proc) * (deriv_factor * fabs(Lr - st.get(8)));
  const double fl_r = a_proc * st.get(9) + (1. - a_proc) * Lr;
  const double mod_r = peaq_div(fd_r, 1. + peaq_div(fl_r, 0.3));
  st.set(10, fd_r);
  st.set(9, fl_r);
  st.set(8, Lr);
  const double fd_t = a_proc * st.get(13) + (1 - a_proc) * (deriv_factor * fabs(Lt - st.get(11)));
  const double fl_t = a_proc * st.get(12) + (1. - a_proc) * Lt;
  const double mod_t = peaq_div(fd_t, 1. + peaq_div(fl_t, 0.3));
  st.set(13, fd_t);
  st.set(12, fl_t);
  st.set(11, Lt);

  // everything below needs no other band or channel, so it runs here, alongside the
  // detection-probability chain: modulation difference terms (movs.c:226-242), the
  // modulation-only factor of the noise loudness (movs.c:725-738), the noise-to-mask
  // ratio term (movs.c:1002-1011)
  double r2[kRed2];
  {
    const double diff = fabs(mod_r - mod_t);
    r2[0] = peaq_div(diff, 1. + mod_r);
    const double w = mod_t >= mod_r ? 1. : .1;
    r2[1] = peaq_div(w * diff, 0.01 + mod_r);
    r2[2] = peaq_div(fl_r, fl_r + 100. * kc.in03());
  }
  const double sref = 0.15 * mod_r + 0.5;
  const double stest = 0.15 * mod_t + 0.5;
  const double nl_fac = peaq_exp_clamped(0.23 * peaq_log_pos(peaq_div(kc.in_noise(), stest)));
  const double curr_nmr = peaq_div(nz, peaq_div(Er, kc.maskdiff()));
  r2[4] = curr_nmr;
  double nmr_max = curr_nmr > 0. ? curr_nmr : 0.;

  // level adaptation, first part (leveladapter.c:262-277)
  const double Rf = a_proc * st.get(2) + (1 - a_proc) * Er;
  const double Tf = a_proc * st.get(3) + (1 - a_proc) * Et;
  st.set(2, Rf);
  st.set(3, Tf);
  double r1[kRed1];
  r1[0] = active ? peaq_sqrt(Rf * Tf) : 0.;
  r1[1] = active ? Tf : 0.;
  // loudness until the latch is set (earmodel.c:890-907)
  r1[2] = 0.;
  r1[3] = 0.;
  if (loud_frame == UINT_MAX && active) {
    const double loudfac = kc.loudfac(), thres = kc.thres(), ethres = kc.ethres();
    const double lr = loudfac * (peaq_exp_clamped(0.23 * peaq_log_pos(1. - thres + peaq_div(thres * Er, ethres))) - 1.);
    const double lt = loudfac * (peaq_exp_clamped(0.23 * peaq_log_pos(1. - thres + peaq_div(thres * Et, ethres))) - 1.);
    r1[2] = lr > 0. ? lr : 0.;
    r1[3] = lt > 0. ? lt : 0.;
  }
  // detection probability of this channel (movs.c:1240-1260)
  {
    const double eref_db = 10. * (peaq_log_pos(Er) * 0.4342944819032518);
    const double etest_db = 10. * (peaq_log_pos(Et) * 0.4342944819032518);
    const double l = 0.3 * (eref_db > etest_db ? eref_db : etest_db) + 0.7 * etest_db;
    const double l2 = l * l;
    const double s = l > 0. ? 5.95072 * peaq_exp_clamped(1.71332 * peaq_log_pos(peaq_div(6.39468, l))) + 9.01033e-11 * (l2 * l2) +
                                  5.05622e-6 * (l2 * l) - 0.00102438 * l * l + 0.0550197 * l -
                                  0.198719
                            : 1e30;
    const double e = eref_db - etest_db;
    const double t1 = peaq_div(e, s), t2 = t1 * t1;
    const double tb = eref_db > etest_db ? t2 * t2 : t2 * t2 * t2;   // (e/s)^b, b = 4 or 6
    pq[c][0][b] = 1. - peaq_exp_clamped(-0.6931471805599453 * tb);   // 1 - 0.5^tb
    pq[c][1][b] = peaq_div(fabs(trunc(e)), s);
  }
#pragma unroll
  for (int k = 0; k < kRed1; k++) {
    const double v = warp_sum(r1[k]);
    if (lane == 0) red1[par][c][wig][k] = v;
  }
  __syncthreads();   // A
  double tot1[kRed1];
#pragma unroll
  for (int k = 0; k < kRed1; k++)
    tot1[k] = ((red1[par][c][0][k] + red1[par][c][1][k]) + red1[par][c][2][k]) + red1[par][c][3][k];

  // level adaptation, second part (leveladapter.c:278-308)
  const double lev_corr = peaq_div(tot1[0] * tot1[0], tot1[1] * tot1[1]);
  double lcr, lct;
  if (lev_corr > 1) {
    lct = Et;
    lcr = peaq_div(Er, lev_corr);
  } else {
    lcr = Er;
    lct = Et * lev_corr;
  }
  const double fnum = a_proc * st.get(4) + lct * lcr;
  const double fden = a_proc * st.get(5) + lcr * lcr;
  st.set(4, fnum);
  st.set(5, fden);
  double pa_r, pa_t;
  if (fnum >= fden) {
    pa_r = 1.;
    pa_t = peaq_div(fden, fnum);
  } else {
    pa_r = peaq_div(fnum, fden);
    pa_t = 1.;
  }
  pa[c][0][b] = pa_r;
  pa[c][1][b] = pa_t;
  // loudness-reached latch (gstpeaq.c:841-845): any channel with both > 0.1
  if (threadIdx.x == 0) latch_sh = 0;
  __syncthreads();   // B
  if (loud_frame == UINT_MAX && b == 0) {
    const double loud_r = tot1[2] * (24. / B), loud_t = tot1[3] * (24. / B);
    if (loud_r > 0.1 && loud_t > 0.1) latch_sh = 1;   // benign race: all writers store 1
  }

  // pattern adaptation, third part (leveladapter.c:310-339)
  double ra_r = 0., ra_t = 0.;
  for (int l = bb - m1; l <= bb + m2; l++) {
    ra_r += pa[c][0][l];
    ra_t += pa[c][1][l];
  }
  ra_r = peaq_div(ra_r, (double)(m1 + m2 + 1));
  ra_t = peaq_div(ra_t, (double)(m1 + m2 + 1));
  const double pcr = a_proc * st.get(6) + (1 - a_proc) * ra_r;
  const double pct = a_proc * st.get(7) + (1 - a_proc) * ra_t;
  st.set(6, pcr);
  st.set(7, pct);
  const double adr = lcr * pcr, adt = lct * pct;

  // noise loudness term (movs.c:725-738) with alpha 1.5, ThresFac 0.15, S0 0.5; the factor
  // that only depends on the test modulation was prepared in front of barrier A
  {
    const double beta = peaq_exp_clamped(peaq_div(-1.5 * (adt - adr), adr));
    const double d = stest * adt - sref * adr;
    r2[3] = nl_fac * (peaq_exp_clamped(0.23 * peaq_log_pos(1. + peaq_div(d > 0. ? d : 0., kc.in_noise() + sref * adr * beta))) - 1.);
  }
  // binaural detection (movs.c:1261-1267), evaluated by channel 0's threads
  double one_minus_p = 1., qsteps = 0.;
  if (c == 0) {
    double p = 0., q = 0.;
    for (int cc = 0; cc < C; cc++) {
      const double pc = pq[cc][0][b], qc = pq[cc][1][b];
      if (pc > p) p = pc;
      if (cc == 0 || qc > q) q = qc;
    }
    one_minus_p = 1. - p;
    qsteps = q;
  }
  if (!active) {
#pragma unroll
    for (int k = 0; k < 5; k++) r2[k] = 0.;
    nmr_max = 0.;
    one_minus_p = 1.;
    qsteps = 0.;
  }
#pragma unroll
  for (int k = 0; k < 5; k++) {
    const double v = warp_sum(r2[k]);
    if (lane == 0) red2[par][c][wig][k] = v;
  }
  {
    const double vmax = warp_max0(nmr_max);
    const double vprod = warp_prod(one_minus_p);
    const double vq = warp_sum(qsteps);
    if (lane == 0) {
      red2[par][c][wig][5] = vmax;
      red2[par][c][wig][6] = vprod;
      red2[par][c][wig][7] = vq;
    }
  }
  __syncthreads();   // C
  if (loud_frame == UINT_MAX && latch_sh) loud_frame = frame_counter;

  if (in.dbg && b == 0) {
    // per-frame terms as the reference's MOV functions see them (before gates and accumulators)
    double t[kRed2];
#pragma unroll
    for (int k = 0; k < 5; k++)
      t[k] = ((red2[par][c][0][k] + red2[par][c][1][k]) + red2[par][c][2][k]) + red2[par][c][3][k];
    double mx = red2[par][c][0][5], prod = 1., q = 0.;
    for (int w = 0; w < kWarpsPerGroup; w++) {
      if (red2[par][c][w][5] > mx) mx = red2[par][c][w][5];
      prod *= red2[par][0][w][6];
      q += red2[par][0][w][7];
    }
    double* d = in.dbg + 2 * C * B + c * kScanTapValues;
    const double nl = t[3] * (24. / B);
    d[0] = t[0] * (100. / B);
    d[1] = t[1] * (100. / B);
    d[2] = t[2];
    d[3] = nl < 0. ? 0. : nl;
    d[4] = t[4] / B;
    d[5] = mx;
    d[6] = 1. - prod;
    d[7] = q;
  }
  // ---- accumulators: thread (c, k) owns slot k of channel c -------------------
  const bool seg_start = frame_counter == cnt.acc_start;   // (frame 0 of a whole item: everything is zero anyway)
  if (acc_thread) {
    acc_io.load(acc);
    if (seg_start) {
      // the sums restart, histories (window, filter state) carry over from the warm-up frames
      acc.num = acc.den = acc.snum = acc.sden = acc.smax = 0.;
      if (acc_mode == kModeFilteredMax) acc.x0 = 0.;
    }
    // peaq_movaccum_set_tentative (movaccum.c:317-354) for every slot
    int st_new = status;
    if (!above) {
      if (status == kStNormal) {
        acc.snum = acc.num;
        acc.sden = acc.den;
        acc.smax = acc.x0;
        st_new = kStTentative;
      }
    } else {
      st_new = kStNormal;
    }
    if (st_new != kStInit) {
      double t2[kRed2];
#pragma unroll
      for (int k = 0; k < 5; k++)
        t2[k] = ((red2[par][c][0][k] + red2[par][c][1][k]) + red2[par][c][2][k]) + red2[par][c][3][k];
      const bool md_gate = frame_counter >= 24;                        // gstpeaq.c:871
      const bool nl_gate = md_gate && frame_counter - 3 >= loud_frame;   // gstpeaq.c:880-881
      const int k = b;
      const bool single = (k == kAdb || k == kMfpd);   // one accumulator channel (gstpeaq.c:580-584)
      if (!(single && c != 0)) {
        switch (k) {
          case kBwRef:
          case kBwTest: {
            const int bw_ref = in.bw[0], bw_test = in.bw[1];
            if (bw_ref > 346) acc_add(acc, acc_mode, k == kBwRef ? bw_ref : bw_test, 1.);
            break;
          }
          case kTotalNmr:
            acc_add(acc, acc_mode, t2[4] / B, 1.);
            break;
          case kRelDistFrames: {
            double mx = red2[par][c][0][5];
            for (int w = 1; w < kWarpsPerGroup; w++)
              if (red2[par][c][w][5] > mx) mx = red2[par][c][w][5];
            acc_add(acc, acc_mode, mx > 1.41253754462275 ? 1. : 0., 1.);
            break;
          }
          case kWinModDiff:
            if (md_gate) acc_add(acc, acc_mode, t2[0] * (100. / B), 1.);
            break;
          case kAvgModDiff1:
            if (md_gate) acc_add(acc, acc_mode, t2[0] * (100. / B), t2[2]);
            break;
          case kAvgModDiff2:
            if (md_gate) acc_add(acc, acc_mode, t2[1] * (100. / B), t2[2]);
            break;
          case kRmsNoiseLoud:
            if (nl_gate) {
              double nl = t2[3] * (24. / B);
              if (nl < 0.) nl = 0.;
              acc_add(acc, acc_mode, nl, 1.);
            }
            break;
          case kAdb:
          case kMfpd: {
            double prod = 1., q = 0.;
            for (int w = 0; w < kWarpsPerGroup; w++) {
              prod *= red2[par][0][w][6];
              q += red2[par][0][w][7];
            }
            const double p_bin = 1. - prod;
            if (k == kMfpd) acc_add(acc, acc_mode, p_bin, 1.);
            else if (p_bin > 0.5) acc_add(acc, acc_mode, q, 1.);
            break;
          }
          case kEhs:
            if (flags & kRecFlagEhsValid) acc_add(acc, acc_mode, 1000. * in.ehs[0], 1.);
            break;
        }
      }
    }
  }
  if (acc_thread) acc_io.store(acc);
  // every thread tracks the shared status word and the counters
  if (!above) {
    if (status == kStNormal) status = kStTentative;
  } else {
    status = kStNormal;
  }
  if (seg_start) cnt.sig_energy = cnt.noise_energy = 0.;
  if (above) {
    if (cnt.first_above == 0) cnt.first_above = frame_counter + 1;
    if (frame_counter >= cnt.acc_start) cnt.owned_above = 1;
  }
  cnt.sig_energy += in.snr[0];
  cnt.noise_energy += in.snr[1];
  frame_counter++;
}

// Evaluation after the last frame of a chunk: accumulator values, channel average, neural
// network, ODG.  All threads of the CTA call it (two barriers).
__device__ __forceinline__ void scan_epilogue(const DeviceTables* __restrict__ T, const Acc& acc,
                                              const ScanCounters& cnt, ScanShared& sh, const ScanThread& th,
                                              int C, PairResult* __restrict__ result) {
  if (th.acc_thread) {
    const bool single = (th.b == kAdb || th.b == kMfpd);
    sh.movval[th.c][th.b] = (single && th.c != 0) ? 0. : acc_value(acc, th.acc_mode, cnt.status == kStTentative);
  }
  __syncthreads();
  // ---- results: channel average, neural network, ODG ------------------------------
  if (threadIdx.x < kNumAcc) {
    const int k = threadIdx.x;
    const bool single = (k == kAdb || k == kMfpd);
    double v = 0.;
    const int nch = single ? 1 : C;
    for (int cc = 0; cc < nch; cc++) v += sh.movval[cc][k];
    v /= nch;
    sh.movs_sh[k] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    PairResult& r = *result;
    double x[5];
    for (int j = 0; j < T->nn_hidden; j++) x[j] = T->nn_wxb[j];
    for (int i = 0; i < T->nn_inputs; i++) {
      const double m = (sh.movs_sh[i] - T->nn_amin[i]) / (T->nn_amax[i] - T->nn_amin[i]);
      for (int j = 0; j < T->nn_hidden; j++) x[j] += T->nn_wx[i * 5 + j] * m;
    }
    double di = T->nn_wyb;
    for (int j = 0; j < T->nn_hidden; j++) di += T->nn_wy[j] / (1 + exp(-x[j]));
    r.di = di;
    r.odg = -3.98 + (0.22 - -3.98) / (1 + exp(-di));
    r.totalsnr = 10 * log10(cnt.sig_energy / cnt.noise_energy);
    for (int i = 0; i < kNumAcc; i++) r.movs[i] = sh.movs_sh[i];
    r.n_movs = kNumAcc;
    r.frames_fft = cnt.frame_counter;
    r.frames_fb = 0;
    r.loudness_reached_frame = cnt.loud_frame;
  }
}

// state <-> registers (layout: StateLayout in peaq_engine.h)
__device__ __forceinline__ void load_acc(Acc& acc, const double* a) {
  acc.num = a[0]; acc.den = a[1]; acc.x0 = a[2]; acc.x1 = a[3]; acc.x2 = a[4];
  acc.snum = a[5]; acc.sden = a[6]; acc.smax = a[7];
}
__device__ __forceinline__ void store_acc(const Acc& acc, double* a) {
  a[0] = acc.num; a[1] = acc.den; a[2] = acc.x0; a[3] = acc.x1; a[4] = acc.x2;
  a[5] = acc.snum; a[6] = acc.sden; a[7] = acc.smax;
}
__device__ __forceinline__ void load_counters(ScanCounters& cnt, const double* st, const StateLayout& S) {
  const int* st_ints = reinterpret_cast<const int*>(st + S.off_ints);
  cnt.status = st_ints[0];
  cnt.frame_counter = (unsigned)st_ints[1];
  cnt.loud_frame = (unsigned)st_ints[2];
  cnt.acc_start = (unsigned)st_ints[kSegAccStart];
  cnt.owned_above = st_ints[kSegOwnedAbove];
  cnt.first_above = (unsigned)st_ints[kSegFirstAbove];
  cnt.sig_energy = st[S.off_scalar];
  cnt.noise_energy = st[S.off_scalar + 1];
}
__device__ __forceinline__ void store_counters(const ScanCounters& cnt, double* st, const StateLayout& S) {
  int* st_ints = reinterpret_cast<int*>(st + S.off_ints);
  st_ints[0] = cnt.status;
  st_ints[1] = (int)cnt.frame_counter;
  st_ints[2] = (int)cnt.loud_frame;
  st_ints[kSegOwnedAbove] = cnt.owned_above;
  st_ints[kSegFirstAbove] = (int)cnt.first_above;
  st[S.off_scalar] = cnt.sig_energy;
  st[S.off_scalar + 1] = cnt.noise_energy;
}

}  // namespace
}  // namespace peaq
