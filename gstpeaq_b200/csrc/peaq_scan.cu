// K2 `scan_basic`: the recurrent half of basic-mode PEAQ.
//
// One CTA = one (ref,test) pair; C groups of 128 threads, thread (c, b) owns band
// b of channel c.  The CTA walks the pair's frames in order, reading the K1
// records and carrying in registers, across the frame loop,
//   - time-domain smearing of the excitation      (fftearmodel.c:496-504)
//   - level and pattern adaptation                (leveladapter.c:243-340)
//   - modulation processing                       (modpatt.c:223-251)
//   - the loudness-reached latch                  (gstpeaq.c:841-845, earmodel.c:890-907)
// computing per frame the band terms of modulation difference (movs.c:205-254),
// noise loudness (movs.c:709-743), NMR (movs.c:1002-1021) and detection
// probability (movs.c:1234-1275), reducing them over bands with warp shuffles,
// and feeding the 11 MOV accumulators with the reference's INIT / NORMAL /
// TENTATIVE semantics (movaccum.c:317-481; call order of gstpeaq.c:850-921).
// The state is loaded from / stored to global memory at chunk boundaries; the
// epilogue evaluates the accumulators, the 11-3-1 network and the ODG mapping
// (nn.c:187-216, :372-375).
//
// Compiled with -fmad=false: expressions round like the reference's C code.
// Powers x^y are evaluated as exp(y ln x) (integer powers by multiplication,
// 0.5^y as exp2(-y)): ~1e-15 relative deviation from libm's pow, 3x fewer
// instructions.
#include "peaq_engine.h"

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

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
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

__global__ void __launch_bounds__(kGroup * kMaxChannels, 2)
scan_basic_kernel(const DeviceTables* __restrict__ T, const double* __restrict__ records,
                  RecordLayout L, const unsigned* __restrict__ n_frames, unsigned first_frame,
                  unsigned n_chunk_frames, double* __restrict__ state, StateLayout S,
                  PairResult* __restrict__ results) {
  const int C = L.C, B = L.B;
  const int pair = blockIdx.x;
  const int c = threadIdx.x / kGroup;
  const int b = threadIdx.x % kGroup;
  const int lane = threadIdx.x & 31;
  const int wig = (threadIdx.x >> 5) % kWarpsPerGroup;   // warp in group
  const bool active = b < B;
  const int bb = active ? b : 0;

  __shared__ double red1[2][kMaxChannels][kWarpsPerGroup][kRed1];
  __shared__ double red2[2][kMaxChannels][kWarpsPerGroup][kRed2];
  __shared__ double pa[kMaxChannels][2][kGroup];        // pattern adaptation factors
  __shared__ double pq[kMaxChannels][2][kGroup];        // detection probability / steps
  __shared__ double movval[kMaxChannels][kNumAcc];
  __shared__ double movs_sh[kNumAcc];
  __shared__ int latch_sh;

  double* st = state + (size_t)pair * S.stride;
  int* st_ints = reinterpret_cast<int*>(st + S.off_ints);

  // ---- load state -----------------------------------------------------------
  double bs[kBandStateFields];
#pragma unroll
  for (int f = 0; f < kBandStateFields; f++)
    bs[f] = active ? st[S.off_band + (f * C + c) * B + b] : 0.;
  double& Efr = bs[0]; double& Eft = bs[1]; double& Rf = bs[2]; double& Tf = bs[3];
  double& fnum = bs[4]; double& fden = bs[5]; double& pcr = bs[6]; double& pct = bs[7];
  double& prev_r = bs[8]; double& fl_r = bs[9]; double& fd_r = bs[10];
  double& prev_t = bs[11]; double& fl_t = bs[12]; double& fd_t = bs[13];

  Acc acc = {0, 0, 0, 0, 0, 0, 0, 0};
  const bool acc_thread = b < kNumAcc;
  const int acc_mode = acc_mode_basic(b < kNumAcc ? b : 0);
  if (acc_thread) {
    const double* a = st + S.off_acc + (c * kNumAcc + b) * kAccFields;
    acc.num = a[0]; acc.den = a[1]; acc.x0 = a[2]; acc.x1 = a[3]; acc.x2 = a[4];
    acc.snum = a[5]; acc.sden = a[6]; acc.smax = a[7];
  }
  int status = st_ints[0];
  unsigned frame_counter = (unsigned)st_ints[1];
  unsigned loud_frame = (unsigned)st_ints[2];
  double sig_energy = st[S.off_scalar], noise_energy = st[S.off_scalar + 1];

  // per-band constants
  const double a_ear = T->fft.a_ear[bb], a_proc = T->fft.a_proc[bb];
  const double in_noise = T->fft.internal_noise[bb], in03 = T->fft.internal_noise_pow03[bb];
  const double ethres = T->fft.ethres[bb], thres = T->fft.thres[bb], loudfac = T->fft.loudfac[bb];
  const double maskdiff = T->maskdiff[bb];
  const double deriv_factor = (double)48000 / kFftStep;
  const int m1 = min(bb, B / 36), m2 = min(B - bb - 1, B / 25);   // leveladapter.c:318-319

  const unsigned total = n_frames[pair];
  const unsigned end = min(first_frame + n_chunk_frames, total);

  for (unsigned f = first_frame; f < end; f++) {
    const int par = f & 1;
    const double* rec = records + ((size_t)pair * n_chunk_frames + (f - first_frame)) * L.stride;
    const int* rints = reinterpret_cast<const int*>(rec + L.off_ints);
    const double E2r = active ? rec[(0 * C + c) * B + b] : 1.;
    const double E2t = active ? rec[(1 * C + c) * B + b] : 1.;
    const double nz = active ? rec[L.off_noise + c * B + b] : 0.;
    const int flags = rints[0];
    const bool above = flags & kRecFlagAbove;

    // time-domain smearing (fftearmodel.c:496-504)
    Efr = a_ear * Efr + (1. - a_ear) * E2r;
    const double Er = Efr > E2r ? Efr : E2r;
    Eft = a_ear * Eft + (1. - a_ear) * E2t;
    const double Et = Eft > E2t ? Eft : E2t;

    // modulation (modpatt.c:234-250): needs nothing but the unsmeared excitations, so it sits in
    // this phase, where its exp/log chain overlaps the detection-probability chain below
    const double Lr = exp(0.3 * log(E2r)), Lt = exp(0.3 * log(E2t));
    fd_r = a_proc * fd_r + (1 - a_proc) * (deriv_factor * fabs(Lr - prev_r));
    fl_r = a_proc * fl_r + (1. - a_proc) * Lr;
    const double mod_r = fd_r / (1. + fl_r / 0.3);
    prev_r = Lr;
    fd_t = a_proc * fd_t + (1 - a_proc) * (deriv_factor * fabs(Lt - prev_t));
    fl_t = a_proc * fl_t + (1. - a_proc) * Lt;
    const double mod_t = fd_t / (1. + fl_t / 0.3);
    prev_t = Lt;

    // everything below needs no other band or channel, so it runs here, alongside the
    // detection-probability chain: modulation difference terms (movs.c:226-242), the
    // modulation-only factor of the noise loudness (movs.c:725-738), the noise-to-mask
    // ratio term (movs.c:1002-1011)
    double r2[kRed2];
    {
      const double diff = fabs(mod_r - mod_t);
      r2[0] = diff / (1. + mod_r);
      const double w = mod_t >= mod_r ? 1. : .1;
      r2[1] = w * diff / (0.01 + mod_r);
      r2[2] = fl_r / (fl_r + 100. * in03);
    }
    const double sref = 0.15 * mod_r + 0.5;
    const double stest = 0.15 * mod_t + 0.5;
    const double nl_fac = exp(0.23 * log(in_noise / stest));
    const double curr_nmr = nz / (Er / maskdiff);
    r2[4] = curr_nmr;
    double nmr_max = curr_nmr > 0. ? curr_nmr : 0.;

    // level adaptation, first part (leveladapter.c:262-277)
    Rf = a_proc * Rf + (1 - a_proc) * Er;
    Tf = a_proc * Tf + (1 - a_proc) * Et;
    double r1[kRed1];
    r1[0] = active ? sqrt(Rf * Tf) : 0.;
    r1[1] = active ? Tf : 0.;
    // loudness until the latch is set (earmodel.c:890-907)
    r1[2] = 0.;
    r1[3] = 0.;
    if (loud_frame == UINT_MAX && active) {
      const double lr = loudfac * (exp(0.23 * log(1. - thres + thres * Er / ethres)) - 1.);
      const double lt = loudfac * (exp(0.23 * log(1. - thres + thres * Et / ethres)) - 1.);
      r1[2] = lr > 0. ? lr : 0.;
      r1[3] = lt > 0. ? lt : 0.;
    }
    // detection probability of this channel (movs.c:1240-1260)
    {
      const double eref_db = 10. * log10(Er);
      const double etest_db = 10. * log10(Et);
      const double l = 0.3 * (eref_db > etest_db ? eref_db : etest_db) + 0.7 * etest_db;
      const double l2 = l * l;
      const double s = l > 0. ? 5.95072 * exp(1.71332 * log(6.39468 / l)) + 9.01033e-11 * (l2 * l2) +
                                    5.05622e-6 * (l2 * l) - 0.00102438 * l * l + 0.0550197 * l -
                                    0.198719
                              : 1e30;
      const double e = eref_db - etest_db;
      const double t1 = e / s, t2 = t1 * t1;
      const double tb = eref_db > etest_db ? t2 * t2 : t2 * t2 * t2;   // (e/s)^b, b = 4 or 6
      pq[c][0][b] = 1. - exp2(-tb);
      pq[c][1][b] = fabs(trunc(e)) / s;
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
    const double lev_corr = tot1[0] * tot1[0] / (tot1[1] * tot1[1]);
    double lcr, lct;
    if (lev_corr > 1) {
      lct = Et;
      lcr = Er / lev_corr;
    } else {
      lcr = Er;
      lct = Et * lev_corr;
    }
    fnum = a_proc * fnum + lct * lcr;
    fden = a_proc * fden + lcr * lcr;
    double pa_r, pa_t;
    if (fnum >= fden) {
      pa_r = 1.;
      pa_t = fden / fnum;
    } else {
      pa_r = fnum / fden;
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
    ra_r /= (m1 + m2 + 1);
    ra_t /= (m1 + m2 + 1);
    pcr = a_proc * pcr + (1 - a_proc) * ra_r;
    pct = a_proc * pct + (1 - a_proc) * ra_t;
    const double adr = lcr * pcr, adt = lct * pct;

    // noise loudness term (movs.c:725-738) with alpha 1.5, ThresFac 0.15, S0 0.5; the factor
    // that only depends on the test modulation was prepared in front of barrier A
    {
      const double beta = exp(-1.5 * (adt - adr) / adr);
      const double d = stest * adt - sref * adr;
      r2[3] = nl_fac * (exp(0.23 * log(1. + (d > 0. ? d : 0.) / (in_noise + sref * adr * beta))) - 1.);
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

    // ---- accumulators: thread (c, k) owns slot k of channel c -------------------
    if (acc_thread) {
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
              const int bw_ref = rints[1 + 2 * c], bw_test = rints[2 + 2 * c];
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
              if (flags & kRecFlagEhsValid) acc_add(acc, acc_mode, 1000. * rec[L.off_ehs + c], 1.);
              break;
          }
        }
      }
    }
    // every thread tracks the shared status word and the counters
    if (!above) {
      if (status == kStNormal) status = kStTentative;
    } else {
      status = kStNormal;
    }
    sig_energy += rec[L.off_snr];
    noise_energy += rec[L.off_snr + 1];
    frame_counter++;
  }

  // ---- store state ---------------------------------------------------------
  if (active) {
#pragma unroll
    for (int f = 0; f < kBandStateFields; f++) st[S.off_band + (f * C + c) * B + b] = bs[f];
  }
  if (acc_thread) {
    double* a = st + S.off_acc + (c * kNumAcc + b) * kAccFields;
    a[0] = acc.num; a[1] = acc.den; a[2] = acc.x0; a[3] = acc.x1; a[4] = acc.x2;
    a[5] = acc.snum; a[6] = acc.sden; a[7] = acc.smax;
    const bool single = (b == kAdb || b == kMfpd);
    movval[c][b] = (single && c != 0) ? 0. : acc_value(acc, acc_mode, status == kStTentative);
  }
  if (threadIdx.x == 0) {
    st_ints[0] = status;
    st_ints[1] = (int)frame_counter;
    st_ints[2] = (int)loud_frame;
    st[S.off_scalar] = sig_energy;
    st[S.off_scalar + 1] = noise_energy;
  }
  __syncthreads();
  // ---- results: channel average, neural network, ODG ------------------------------
  if (threadIdx.x < kNumAcc) {
    const int k = threadIdx.x;
    const bool single = (k == kAdb || k == kMfpd);
    double v = 0.;
    const int nch = single ? 1 : C;
    for (int cc = 0; cc < nch; cc++) v += movval[cc][k];
    v /= nch;
    movs_sh[k] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    PairResult& r = results[pair];
    double x[5];
    for (int j = 0; j < T->nn_hidden; j++) x[j] = T->nn_wxb[j];
    for (int i = 0; i < T->nn_inputs; i++) {
      const double m = (movs_sh[i] - T->nn_amin[i]) / (T->nn_amax[i] - T->nn_amin[i]);
      for (int j = 0; j < T->nn_hidden; j++) x[j] += T->nn_wx[i * 5 + j] * m;
    }
    double di = T->nn_wyb;
    for (int j = 0; j < T->nn_hidden; j++) di += T->nn_wy[j] / (1 + exp(-x[j]));
    r.di = di;
    r.odg = -3.98 + (0.22 - -3.98) / (1 + exp(-di));
    r.totalsnr = 10 * log10(sig_energy / noise_energy);
    for (int i = 0; i < kNumAcc; i++) r.movs[i] = movs_sh[i];
    r.n_movs = kNumAcc;
    r.frames_fft = frame_counter;
    r.frames_fb = 0;
    r.loudness_reached_frame = loud_frame;
  }
}

}  // namespace

cudaError_t launch_scan_basic(const DeviceTables* d_tables, const double* records, RecordLayout L,
                              const unsigned* n_frames, unsigned first_frame,
                              unsigned n_chunk_frames, double* state, StateLayout S,
                              PairResult* results, int n_pairs, cudaStream_t stream) {
  if (n_pairs <= 0) return cudaSuccess;
  scan_basic_kernel<<<n_pairs, kGroup * L.C, 0, stream>>>(d_tables, records, L, n_frames, first_frame,
                                                           n_chunk_frames, state, S, results);
  return cudaGetLastError();
}

}  // namespace peaq
