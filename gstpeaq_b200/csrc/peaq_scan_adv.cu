// Advanced-mode recurrent kernels.
//
// FB3a `fb_spread_kernel`: frequency-domain spreading and rectification of the
// filter-bank outputs (fbearmodel.c:320-357), once per 32-sample sub-step.  Only the
// first-order smoothing of the slope `cu` carries state from one sub-step to the next;
// everything else is independent per sub-step, so a CTA (one per stream) works on tiles
// of 32 sub-steps: the level-dependent slope of every (band, sub-step) in parallel, the
// 40 `cu` recurrences serially through the tile, then one thread per (sub-step, re|im)
// runs the upward ladder entirely in registers -- in the reference's own order, source
// band by source band, so the sums round exactly as there -- and the downward pass.
//
// FB3b `fb_scan_kernel`: everything that runs on the 192-sample clock.  One CTA per
// pair, one warp per stream (channel, ref|test), lane = band (40 bands: lane l owns band
// l and, for l < 8, band 32 + l).  Per frame: backward-masking FIR over the last eleven
// sub-step energies (kept in registers), internal noise, forward masking
// (fbearmodel.c:371-395); then per channel (the ref warp): level / pattern
// adaptation and modulation at 40 bands (leveladapter.c, modpatt.c), loudness
// latch, RmsModDiffA, RmsNoiseLoudAsymA, AvgLinDistA with their accumulators
// (process_fb_block, gstpeaq.c:965-1010; movs.c:205-254, 551-577, 679-743).
//
// `adv_fft_scan_kernel`: the 1024-sample clock of advanced mode
// (process_fft_block_advanced, gstpeaq.c:924-962): time smearing of the ref
// excitation at 55 bands, SegmentalNMRB, EHSB, SNR sums; its epilogue combines
// all five MOVs through the 5-5-1 network (nn.c:304-335, 372-375).
//
// Compiled with -fmad=false (reference rounding); recurrent state is loaded
// from / stored to global memory at chunk boundaries.
#include "peaq_engine.h"
#include "peaq_math.cuh"

#include <climits>

namespace peaq {
namespace {

constexpr double kSlopeA = 0.993355506255034;   // fbearmodel.c:49
constexpr double kCl = 0.0802581846102741;      // fbearmodel.c:51
constexpr double kLnDist = -0.08137117849224008;  // ln DIST, DIST = 0.921851456499719 (fbearmodel.c:50)
constexpr double kKappa = 0.07067810761028887;    // -0.2 * 10 / ln(10) * ln DIST
enum { kStInit = 0, kStNormal = 1, kStTentative = 2 };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// calc_noise_loudness band term (movs.c:725-738)
__device__ __forceinline__ double nl_term(double alpha, double thres_fac, double S0, double ethres,
                                          double ref_mod, double test_mod, double ep_ref,
                                          double ep_test) {
  const double sref = thres_fac * ref_mod + S0;
  const double stest = thres_fac * test_mod + S0;
  const double beta = peaq_exp_clamped(peaq_div(-alpha * (ep_test - ep_ref), ep_ref));
  const double d = stest * ep_test - sref * ep_ref;
  return peaq_exp_clamped(0.23 * peaq_log_pos(peaq_div(ethres, stest))) *
         (peaq_exp_clamped(0.23 * peaq_log_pos(1. + peaq_div(d > 0. ? d : 0., ethres + sref * ep_ref * beta))) - 1.);
}

// per-band constants of the filter-bank model, staged in shared memory
enum { kCFc, kCNoise, kCNoise03, kCAEar, kCAProc, kCEthres, kCThres, kCLoudfac, kCCount };

struct FbSmem {
  double cst[kCCount][kFbBands];
  double ex_u[2 * kMaxChannels][kFbBands];
  double ex_e[2 * kMaxChannels][kFbBands];
  double lv[kMaxChannels][6][kFbBands];       // ref_filt, test_filt, num, den, pc_ref, pc_test
  double md[kMaxChannels][2][3][kFbBands];    // [ref|test][prev, filt_loud, filt_deriv]
  double pa[kMaxChannels][2][kFbBands];
  double acc[kMaxChannels][3][kAccFields];
  double mod[2 * kMaxChannels][kFbBands];     // modulation of every stream (written by its own warp)
  double ad[kMaxChannels][2][kFbBands];       // spectrally adapted patterns, ref | test (written by the ref warp)
  double tsum[kMaxChannels][2];               // band sums computed by the test warp: missing components, lin. dist.
  int latch;
};

constexpr int kSpTile = 32;   // sub-steps per tile
constexpr int kSpPad = 33;    // row pitch: column reads (fixed sub-step) and row-major read-out both conflict free

struct SpreadSmem {
  double re[kFbBands][kSpPad];
  double im[kFbBands][kSpPad];
  double cu[kFbBands][kSpPad];   // DIST^s, then the smoothed slope
  double slope0[kFbBands];       // (24 + 230 / fc) ln DIST
};

__global__ void __launch_bounds__(2 * kSpTile, 6)
fb_spread_kernel(const DeviceTables* __restrict__ T, const double2* __restrict__ fbout, unsigned n_sub,
                 const unsigned* __restrict__ n_frames, unsigned first_frame, double* __restrict__ state,
                 AdvStateLayout S, double* __restrict__ energy /* [stream][n_sub][40] */) {
  __shared__ SpreadSmem sm;
  const int stream = blockIdx.x;
  const int per_pair = 2 * S.C;
  const int pair = stream / per_pair;
  double* st_stream = state + (size_t)pair * S.stride + S.off_fb_stream + (stream % per_pair) * (2 + 11) * kFbBands;
  const unsigned total = n_frames[pair];
  const unsigned valid = total > first_frame ? min((total - first_frame) * 6u, n_sub) : 0u;
  const int tid = threadIdx.x, t = tid & 31, part = tid >> 5;
  double cu_state = 0.;
  if (tid < kFbBands) {
    cu_state = st_stream[tid];
    sm.slope0[tid] = (24 + 230 / T->fb.fc[tid]) * kLnDist;   // ln DIST^(24 + 230 / fc)
  }
  const double2* __restrict__ my_out = fbout + (size_t)stream * kFbBands * n_sub;
  double* __restrict__ my_e = energy + (size_t)stream * kFbBands * n_sub;
  __syncthreads();

  for (unsigned s0 = 0; s0 < valid; s0 += kSpTile) {
    const int nt = (int)min((unsigned)kSpTile, valid - s0);
    // ---- level dependent slope of every (band, sub-step) (fbearmodel.c:326-331) ----
    if (t < nt) {
      // DIST^max(4, s0 - 0.2 L) with L = 10 log10 p, written as exp(min(4 ln DIST, c0 + kappa ln p)):
      // one ln and one exp per value instead of log10 + exp (same value to ~1e-16).
      // Four bands at a time in a ROLLED loop, the next four loaded meanwhile: unrolled twenty
      // times the ln / exp code alone was 35 KB of this kernel's 87 KB, and a warp walks through
      // all of it once per tile -- 23 % of the kernel's stall samples were instruction fetches.
      constexpr int kG = 4;
      double2 o[kG], nx[kG];
#pragma unroll
      for (int k = 0; k < kG; k++) nx[k] = my_out[(size_t)(part + 2 * k) * n_sub + s0 + t];
#pragma unroll 1
      for (int k0 = 0; k0 < kFbBands / 2; k0 += kG) {
#pragma unroll
        for (int k = 0; k < kG; k++) o[k] = nx[k];
        if (k0 + kG < kFbBands / 2) {
#pragma unroll
          for (int k = 0; k < kG; k++) nx[k] = my_out[(size_t)(part + 2 * (k0 + kG + k)) * n_sub + s0 + t];
        }
#pragma unroll
        for (int k = 0; k < kG; k++) {
          const int b = part + 2 * (k0 + k);
          const double e = sm.slope0[b] + kKappa * peaq_log(o[k].x * o[k].x + o[k].y * o[k].y);
          sm.re[b][t] = o[k].x;
          sm.im[b][t] = o[k].y;
          sm.cu[b][t] = peaq_exp(e < 4 * kLnDist ? e : 4 * kLnDist);
        }
      }
    }
    __syncthreads();
    // ---- first-order smoothing along time, one thread per band (fbearmodel.c:336) ----
    if (tid < kFbBands) {
      for (int i = 0; i < nt; i++) {
        cu_state = cu_state + kSlopeA * (sm.cu[tid][i] - cu_state);
        sm.cu[tid][i] = cu_state;
      }
    }
    __syncthreads();
    // ---- spreading of one sub-step, real or imaginary part, in registers ------------
    if (t < nt) {
      double* xs = part ? &sm.im[0][t] : &sm.re[0][t];
      const double* cs = &sm.cu[0][t];
      double A[kFbBands];
#pragma unroll
      for (int j = 0; j < kFbBands; j++) A[j] = xs[j * kSpPad];
      // upward (fbearmodel.c:337-347): source b adds out[b] cu[b]^(j-b) to every j > b;
      // eight sources advance together so that their multiply chains overlap
#pragma unroll
      for (int g = 0; g < kFbBands; g += 8) {
        double r[8], c[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
          r[q] = xs[(g + q) * kSpPad];
          c[q] = cs[(g + q) * kSpPad];
        }
#pragma unroll
        for (int j = g + 1; j < kFbBands; j++) {
#pragma unroll
          for (int q = 0; q < 8; q++) {
            if (g + q < j) {
              r[q] *= c[q];
              A[j] += r[q];
            }
          }
        }
      }
      // downward with the constant slope CL (fbearmodel.c:350-353)
#pragma unroll
      for (int j = kFbBands - 1; j > 0; j--) A[j - 1] += kCl * A[j];
#pragma unroll
      for (int j = 0; j < kFbBands; j++) xs[j * kSpPad] = A[j];
    }
    __syncthreads();
    // ---- rectification (fbearmodel.c:356-359); [sub-step][band] rows are contiguous ----
    for (int e = tid; e < nt * kFbBands; e += 2 * kSpTile) {
      const int i = e / kFbBands, b = e - i * kFbBands;
      const double ar = sm.re[b][i], ai = sm.im[b][i];
      my_e[(size_t)s0 * kFbBands + e] = ar * ar + ai * ai;
    }
    __syncthreads();
  }
  if (tid < kFbBands) st_stream[tid] = cu_state;
}

__global__ void __launch_bounds__(64 * kMaxChannels, 4)
fb_scan_kernel(const DeviceTables* __restrict__ T, const double* __restrict__ energy,
               unsigned n_sub /* sub-steps per stream in this chunk */,
               const unsigned char* __restrict__ flags, const unsigned* __restrict__ n_frames,
               unsigned first_frame, unsigned n_chunk_frames, double* __restrict__ state,
               AdvStateLayout S, double* __restrict__ dbg) {
  __shared__ FbSmem sm;
  const int C = S.C;
  const int pair = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chan = warp >> 1, side = warp & 1;
  const int stream = pair * 2 * C + warp;
  double* st = state + (size_t)pair * S.stride;
  int* st_ints = reinterpret_cast<int*>(st + S.off_ints);

  // ---- load state and constants -----------------------------------------------
  // exc: forward-masked excitation; prev: the five sub-step energies before the
  // current frame, oldest first (the state keeps them newest first)
  double exc[2], prev[2][5];
  double* st_stream = st + S.off_fb_stream + warp * (2 + 11) * kFbBands;
#pragma unroll
  for (int sl = 0; sl < 2; sl++) {
    const int b = lane + 32 * sl;
    const bool ok = b < kFbBands;
    exc[sl] = ok ? st_stream[kFbBands + b] : 0.;
#pragma unroll
    for (int k = 0; k < 5; k++) prev[sl][4 - k] = ok ? st_stream[(2 + k) * kFbBands + b] : 0.;
    if (ok) {
      if (side == 0) {
        for (int f = 0; f < 6; f++) sm.lv[chan][f][b] = st[S.off_fb_level + (chan * 6 + f) * kFbBands + b];
        for (int sd = 0; sd < 2; sd++)
          for (int f = 0; f < 3; f++)
            sm.md[chan][sd][f][b] = st[S.off_fb_mod + ((chan * 2 + sd) * 3 + f) * kFbBands + b];
      }
      if (warp == 0) {
        sm.cst[kCFc][b] = T->fb.fc[b];
        sm.cst[kCNoise][b] = T->fb.internal_noise[b];
        sm.cst[kCNoise03][b] = T->fb.internal_noise_pow03[b];
        sm.cst[kCAEar][b] = T->fb.a_ear[b];
        sm.cst[kCAProc][b] = T->fb.a_proc[b];
        sm.cst[kCEthres][b] = T->fb.ethres[b];
        sm.cst[kCThres][b] = T->fb.thres[b];
        sm.cst[kCLoudfac][b] = T->fb.loudfac[b];
      }
    }
  }
  if (side == 0 && lane < 3 * kAccFields)
    sm.acc[chan][lane / kAccFields][lane % kAccFields] = st[S.off_fb_acc + chan * 3 * kAccFields + lane];
  int status = st_ints[2];
  unsigned frame_counter = (unsigned)st_ints[3];
  unsigned loud_frame = (unsigned)st_ints[4];
  // segments of long items (peaq_segments.cu; zero for whole items)
  const unsigned acc_start = (unsigned)st_ints[kASegAccStartFb];
  int owned_above = st_ints[kASegOwnedAboveFb];
  unsigned first_above = (unsigned)st_ints[kASegFirstAboveFb];
  __syncthreads();

  const double deriv_factor = (double)48000 / kFbFrame;
  const double* __restrict__ my_e = energy + (size_t)stream * kFbBands * n_sub;
  double bm[6];
#pragma unroll
  for (int i = 0; i < 6; i++) bm[i] = T->fb_back_mask[i];

  const unsigned total = n_frames[pair];
  const unsigned end = min(first_frame + n_chunk_frames, total);
  // the six sub-step energies of a frame, loaded one frame ahead
  double cur[2][6];
  auto load_frame = [&](unsigned fl, double (&dst)[2][6]) {
#pragma unroll
    for (int sl = 0; sl < 2; sl++) {
      const int b = lane + 32 * sl;
#pragma unroll
      for (int sub = 0; sub < 6; sub++)
        dst[sl][sub] = b < kFbBands ? my_e[((size_t)fl * 6 + sub) * kFbBands + b] : 0.;
    }
  };
  if (first_frame < end) load_frame(0, cur);
  for (unsigned f = first_frame; f < end; f++) {
    const unsigned fl = f - first_frame;
    const bool above = flags[(size_t)pair * n_chunk_frames + fl] != 0;
    if (threadIdx.x == 0) sm.latch = 0;   // set after the first barrier, read after the second
    double nxt[2][6];
    if (f + 1 < end) load_frame(fl + 1, nxt);
    double mod_own[2] = {0., 0.}, avl_own[2] = {0., 0.};
    // ---- backward masking, noise, forward masking (fbearmodel.c:371-395) -------
    // E0_buf[i] (i = 0 newest) is cur[5 - i] for i <= 5 and prev[10 - i] beyond
#pragma unroll
    for (int sl = 0; sl < 2; sl++) {
      const int b = lane + 32 * sl;
      if (b < kFbBands) {
        double e1 = 0.;
#pragma unroll
        for (int i = 0; i < 5; i++) e1 += (cur[sl][5 - i] + prev[sl][i]) * bm[i];
        e1 += cur[sl][0] * bm[5];
        const double U = e1 + sm.cst[kCNoise][b];
        const double a_ear = sm.cst[kCAEar][b];
        exc[sl] = a_ear * exc[sl] + (1. - a_ear) * U;
        sm.ex_u[warp][b] = U;
        sm.ex_e[warp][b] = exc[sl];
        {
          // modulation of this stream (modpatt.c:234-250): needs nothing but its own unsmeared
          // excitation, so every warp does its own
          const double a_proc = sm.cst[kCAProc][b];
          const double loud = peaq_exp_clamped(0.3 * peaq_log_pos(U));
          const double fd = a_proc * sm.md[chan][side][2][b] +
                            (1 - a_proc) * (deriv_factor * fabs(loud - sm.md[chan][side][0][b]));
          const double fl_ = a_proc * sm.md[chan][side][1][b] + (1. - a_proc) * loud;
          sm.md[chan][side][2][b] = fd;
          sm.md[chan][side][1][b] = fl_;
          sm.md[chan][side][0][b] = loud;
          mod_own[sl] = peaq_div(fd, 1. + peaq_div(fl_, 0.3));
          avl_own[sl] = fl_;
          sm.mod[warp][b] = mod_own[sl];
        }
        if (dbg) {
          // [pair][frame][stream][U|E][40]
          double* d = dbg + (((size_t)pair * n_chunk_frames + fl) * 2 * C + warp) * 2 * kFbBands;
          d[b] = U;
          d[kFbBands + b] = exc[sl];
        }
      }
#pragma unroll
      for (int k = 0; k < 5; k++) prev[sl][k] = cur[sl][k + 1];
#pragma unroll
      for (int k = 0; k < 6; k++) cur[sl][k] = nxt[sl][k];
    }
    __syncthreads();   // excitations of all streams visible

    // ---- channel processing on the ref warp (apply_ear_model_and_preprocess) ----
    // per-band values the MOV section needs, kept in registers across the barrier
    double adr[2], adt[2];
    if (side == 0) {
      double lcr[2], lct[2];
      double p_num = 0., p_den = 0., l_r = 0., l_t = 0.;
#pragma unroll
      for (int sl = 0; sl < 2; sl++) {
        const int b = lane + 32 * sl;
        if (b < kFbBands) {
          const double a_proc = sm.cst[kCAProc][b];
          const double Er = sm.ex_e[2 * chan][b], Et = sm.ex_e[2 * chan + 1][b];
          // leveladapter.c:262-277
          const double rf = a_proc * sm.lv[chan][0][b] + (1 - a_proc) * Er;
          const double tf = a_proc * sm.lv[chan][1][b] + (1 - a_proc) * Et;
          sm.lv[chan][0][b] = rf;
          sm.lv[chan][1][b] = tf;
          p_num += peaq_sqrt(rf * tf);
          p_den += tf;
          if (loud_frame == UINT_MAX) {   // earmodel.c:890-907
            const double thres = sm.cst[kCThres][b], ethres = sm.cst[kCEthres][b], lfac = sm.cst[kCLoudfac][b];
            const double a = lfac * (peaq_exp_clamped(0.23 * peaq_log_pos(1. - thres + peaq_div(thres * Er, ethres))) - 1.);
            const double c2 = lfac * (peaq_exp_clamped(0.23 * peaq_log_pos(1. - thres + peaq_div(thres * Et, ethres))) - 1.);
            l_r += a > 0. ? a : 0.;
            l_t += c2 > 0. ? c2 : 0.;
          }
        }
      }
      p_num = warp_sum(p_num);
      p_den = warp_sum(p_den);
      if (loud_frame == UINT_MAX) {
        l_r = warp_sum(l_r) * (24. / kFbBands);
        l_t = warp_sum(l_t) * (24. / kFbBands);
        if (lane == 0 && l_r > 0.1 && l_t > 0.1) sm.latch = 1;   // gstpeaq.c:841-845
      }
      const double lev_corr = peaq_div(p_num * p_num, p_den * p_den);
#pragma unroll
      for (int sl = 0; sl < 2; sl++) {
        const int b = lane + 32 * sl;
        lcr[sl] = lct[sl] = 0.;
        if (b < kFbBands) {
          const double a_proc = sm.cst[kCAProc][b];
          const double Er = sm.ex_e[2 * chan][b], Et = sm.ex_e[2 * chan + 1][b];
          if (lev_corr > 1) {
            lct[sl] = Et;
            lcr[sl] = peaq_div(Er, lev_corr);
          } else {
            lcr[sl] = Er;
            lct[sl] = Et * lev_corr;
          }
          const double num = a_proc * sm.lv[chan][2][b] + lct[sl] * lcr[sl];
          const double den = a_proc * sm.lv[chan][3][b] + lcr[sl] * lcr[sl];
          sm.lv[chan][2][b] = num;
          sm.lv[chan][3][b] = den;
          if (num >= den) {
            sm.pa[chan][0][b] = 1.;
            sm.pa[chan][1][b] = peaq_div(den, num);
          } else {
            sm.pa[chan][0][b] = peaq_div(num, den);
            sm.pa[chan][1][b] = 1.;
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int sl = 0; sl < 2; sl++) {
        const int b = lane + 32 * sl;
        adr[sl] = adt[sl] = 0.;
        if (b < kFbBands) {
          const double a_proc = sm.cst[kCAProc][b];
          // leveladapter.c:315-339 with band_count 40: m1 = min(k,1), m2 = min(40-k-1,1)
          const int m1 = min(b, kFbBands / 36), m2 = min(kFbBands - b - 1, kFbBands / 25);
          double ra_r = 0., ra_t = 0.;
          for (int l = b - m1; l <= b + m2; l++) {
            ra_r += sm.pa[chan][0][l];
            ra_t += sm.pa[chan][1][l];
          }
          ra_r /= (m1 + m2 + 1);
          ra_t /= (m1 + m2 + 1);
          const double pcr = a_proc * sm.lv[chan][4][b] + (1 - a_proc) * ra_r;
          const double pct = a_proc * sm.lv[chan][5][b] + (1 - a_proc) * ra_t;
          sm.lv[chan][4][b] = pcr;
          sm.lv[chan][5][b] = pct;
          adr[sl] = lcr[sl] * pcr;
          adt[sl] = lct[sl] * pct;
          sm.ad[chan][0][b] = adr[sl];
          sm.ad[chan][1][b] = adt[sl];
        }
      }
    }
    __syncthreads();   // latch visible
    if (loud_frame == UINT_MAX && sm.latch) loud_frame = frame_counter;

    // ---- MOVs of the filter-bank clock (gstpeaq.c:987-1007) --------------------
    // band sums: the ref warp takes the modulation difference and the noise loudness, the test
    // warp the two other loudness terms; lane 0 of the ref warp then feeds the accumulators
    const bool md_gate = frame_counter >= 125;
    const bool nl_gate = md_gate && frame_counter - 13 >= loud_frame;
    double s_md = 0., s_wt = 0., s_nl = 0.;
    if (side == 0) {
#pragma unroll
      for (int sl = 0; sl < 2; sl++) {
        const int b = lane + 32 * sl;
        if (b < kFbBands) {
          const double mod_r = mod_own[sl], mod_t = sm.mod[warp + 1][b];
          if (md_gate) {   // movs.c:226-242 with levWt = 1 (no second accumulator)
            const double diff = fabs(mod_r - mod_t);
            s_md += peaq_div(diff, 1. + mod_r);
            s_wt += peaq_div(avl_own[sl], avl_own[sl] + 1. * sm.cst[kCNoise03][b]);
          }
          if (nl_gate)   // peaq_mov_noise_loud_asym, first term (movs.c:551-577)
            s_nl += nl_term(2.5, 0.3, 1., sm.cst[kCNoise][b], mod_r, mod_t, adr[sl], adt[sl]);
        }
      }
      s_md = warp_sum(s_md);
      s_wt = warp_sum(s_wt);
      s_nl = warp_sum(s_nl);
    } else {
      double s_mc = 0., s_ld = 0.;
      if (nl_gate) {
#pragma unroll
        for (int sl = 0; sl < 2; sl++) {
          const int b = lane + 32 * sl;
          if (b < kFbBands) {
            const double in_noise = sm.cst[kCNoise][b];
            const double mod_t = mod_own[sl], mod_r = sm.mod[warp - 1][b];
            const double a_r = sm.ad[chan][0][b], a_t = sm.ad[chan][1][b];
            // the missing-components term swaps ref/test patterns AND modulation (settings.h:47)
            s_mc += nl_term(1.5, 0.15, 1., in_noise, mod_t, mod_r, a_t, a_r);
            // peaq_mov_lin_dist (movs.c:679-706): ref modulation on both sides,
            // adapted ref pattern vs ref excitation
            s_ld += nl_term(1.5, 0.15, 1., in_noise, mod_r, mod_r, a_r, sm.ex_e[2 * chan][b]);
          }
        }
        s_mc = warp_sum(s_mc);
        s_ld = warp_sum(s_ld);
      }
      if (lane == 0) {
        sm.tsum[chan][0] = s_mc;
        sm.tsum[chan][1] = s_ld;
      }
    }
    __syncthreads();   // the test warp's sums visible
    if (side == 0 && lane == 0) {
      const double s_mc = sm.tsum[chan][0], s_ld = sm.tsum[chan][1];
      double (*a)[kAccFields] = sm.acc[chan];
      if (frame_counter == acc_start) {   // a segment's sums restart after its warm-up frames
        for (int k = 0; k < 3; k++) a[k][0] = a[k][1] = a[k][2] = a[k][5] = a[k][6] = a[k][7] = 0.;
      }
      // peaq_movaccum_set_tentative on the three fb-clock accumulators (gstpeaq.c:974-979)
      int st_new = status;
      if (!above) {
        if (status == kStNormal) {
          for (int k = 0; k < 3; k++) {
            a[k][5] = a[k][0];
            a[k][6] = a[k][1];
            a[k][7] = a[k][2];
          }
          st_new = kStTentative;
        }
      } else {
        st_new = kStNormal;
      }
      if (st_new != kStInit) {
        if (md_gate) {
          // MODE_RMS: weight squared (movaccum.c:375-379); value scaled by 100/sqrt(B) (movs.c:243-244)
          const double val = s_md * (100. / sqrt((double)kFbBands));
          double w = s_wt;
          w *= w;
          a[0][0] += w * val * val;
          a[0][1] += w;
        }
        if (nl_gate) {
          double nl = s_nl * (24. / kFbBands);
          if (nl < 0.1) nl = 0.;                       // NLmin (movs.c:740-741)
          double mc = s_mc * (24. / kFbBands);
          if (mc < 0.) mc = 0.;
          double ld = s_ld * (24. / kFbBands);
          if (ld < 0.) ld = 0.;
          // MODE_RMS_ASYM (movaccum.c:380-385): field 2 holds the second numerator
          a[1][0] += nl * nl;
          a[1][2] += mc * mc;
          a[1][1] += 1.;
          a[2][0] += 1. * ld;                          // MODE_AVG, weight 1
          a[2][1] += 1.;
        }
      }
      if (dbg) {
        double* d = dbg + (size_t)gridDim.x * n_chunk_frames * 2 * C * 2 * kFbBands +
                    (((size_t)pair * n_chunk_frames + fl) * C + chan) * 8;
        double nl = s_nl * (24. / kFbBands);
        if (nl < 0.1) nl = 0.;
        d[0] = md_gate ? s_md * (100. / sqrt((double)kFbBands)) : 0.;
        d[1] = md_gate ? s_wt : 0.;
        d[2] = nl_gate ? nl : 0.;
        d[3] = nl_gate ? s_mc * (24. / kFbBands) : 0.;
        d[4] = nl_gate ? s_ld * (24. / kFbBands) : 0.;
        d[5] = above;
      }
    }
    if (!above) {
      if (status == kStNormal) status = kStTentative;
    } else {
      status = kStNormal;
      if (first_above == 0) first_above = frame_counter + 1;
      if (frame_counter >= acc_start) owned_above = 1;
    }
    frame_counter++;
    // no barrier here: everything the next frame overwrites before its first barrier was last
    // read before the one above
  }
  __syncthreads();

  // ---- store state, publish the channel-averaged MOVs ----------------------------
#pragma unroll
  for (int sl = 0; sl < 2; sl++) {
    const int b = lane + 32 * sl;
    if (b < kFbBands) {
      st_stream[kFbBands + b] = exc[sl];
#pragma unroll
      for (int k = 0; k < 5; k++) st_stream[(2 + k) * kFbBands + b] = prev[sl][4 - k];
      if (side == 0) {
        for (int f = 0; f < 6; f++) st[S.off_fb_level + (chan * 6 + f) * kFbBands + b] = sm.lv[chan][f][b];
        for (int sd = 0; sd < 2; sd++)
          for (int f = 0; f < 3; f++)
            st[S.off_fb_mod + ((chan * 2 + sd) * 3 + f) * kFbBands + b] = sm.md[chan][sd][f][b];
      }
    }
  }
  __syncthreads();
  if (side == 0 && lane < 3 * kAccFields)
    st[S.off_fb_acc + chan * 3 * kAccFields + lane] = sm.acc[chan][lane / kAccFields][lane % kAccFields];
  if (threadIdx.x == 0) {
    st_ints[2] = status;
    st_ints[3] = (int)frame_counter;
    st_ints[4] = (int)loud_frame;
    st_ints[kASegOwnedAboveFb] = owned_above;
    st_ints[kASegFirstAboveFb] = (int)first_above;
    // peaq_movaccum_get_value (movaccum.c:438-481), averaged over channels
    const bool tent = status == kStTentative;
    double v0 = 0., v1 = 0., v2 = 0.;
    for (int c = 0; c < C; c++) {
      const double (*a)[kAccFields] = sm.acc[c];
      v0 += sqrt((tent ? a[0][5] : a[0][0]) / (tent ? a[0][6] : a[0][1]));
      const double den1 = tent ? a[1][6] : a[1][1];
      v1 += sqrt((tent ? a[1][5] : a[1][0]) / den1);
      v1 += 0.5 * sqrt((tent ? a[1][7] : a[1][2]) / den1);
      v2 += (tent ? a[2][5] : a[2][0]) / (tent ? a[2][6] : a[2][1]);
    }
    st[S.off_fb_movs] = v0 / C;
    st[S.off_fb_movs + 1] = v1 / C;
    st[S.off_fb_movs + 2] = v2 / C;
  }
}

// ---------------------------------------------------------------------------

constexpr int kAdvGroup = 64;   // threads per channel (55 bands)

__global__ void __launch_bounds__(kAdvGroup * kMaxChannels)
adv_fft_scan_kernel(const DeviceTables* __restrict__ T, const double* __restrict__ records,
                    RecordLayout L, const unsigned* __restrict__ n_frames, unsigned first_frame,
                    unsigned n_chunk_frames, double* __restrict__ state, AdvStateLayout S,
                    PairResult* __restrict__ results) {
  const int C = L.C, B = L.B;
  const int pair = blockIdx.x;
  const int c = threadIdx.x / kAdvGroup;
  const int b = threadIdx.x % kAdvGroup;
  const int lane = threadIdx.x & 31;
  const int wig = (threadIdx.x >> 5) & 1;
  const bool active = b < B;
  const int bb = active ? b : 0;
  __shared__ double red[2][kMaxChannels][2];
  __shared__ double val_sh[kMaxChannels][2];

  double* st = state + (size_t)pair * S.stride;
  int* st_ints = reinterpret_cast<int*>(st + S.off_ints);
  double Efr = active ? st[S.off_fft_filtered + c * B + b] : 0.;
  // accumulator slot owned by thread (c, k): k = 0 SegmentalNMR, 1 EHS; both MODE_AVG
  double num = 0., den = 0., snum = 0., sden = 0.;
  const bool acc_thread = b < 2;
  if (acc_thread) {
    const double* a = st + S.off_fft_acc + (c * 2 + b) * kAccFields;
    num = a[0]; den = a[1]; snum = a[5]; sden = a[6];
  }
  int status = st_ints[0];
  unsigned frame_counter = (unsigned)st_ints[1];
  const unsigned acc_start = (unsigned)st_ints[kASegAccStartFft];   // segments of long items (peaq_segments.cu)
  int owned_above = st_ints[kASegOwnedAboveFft];
  unsigned first_above = (unsigned)st_ints[kASegFirstAboveFft];
  double sig_energy = st[S.off_fft_scalar], noise_energy = st[S.off_fft_scalar + 1];
  const double a_ear = T->fft.a_ear[bb], maskdiff = T->maskdiff[bb];

  const unsigned total = n_frames[pair];
  const unsigned end = min(first_frame + n_chunk_frames, total);
  for (unsigned f = first_frame; f < end; f++) {
    const int par = f & 1;
    const double* rec = records + ((size_t)pair * n_chunk_frames + (f - first_frame)) * L.stride;
    const int flags = reinterpret_cast<const int*>(rec + L.off_ints)[0];
    const bool above = flags & kRecFlagAbove;
    const double E2r = active ? rec[(0 * C + c) * B + b] : 1.;
    const double nz = active ? rec[L.off_noise + c * B + b] : 0.;
    Efr = a_ear * Efr + (1. - a_ear) * E2r;                    // fftearmodel.c:496-504
    const double Er = Efr > E2r ? Efr : E2r;
    double curr = active ? nz / (Er / maskdiff) : 0.;          // movs.c:1002-1011
    curr = warp_sum(curr);
    if (lane == 0) red[par][c][wig] = curr;
    __syncthreads();
    const bool seg_start = frame_counter == acc_start;
    if (seg_start) {   // a segment's sums restart after its warm-up frames
      num = den = snum = sden = 0.;
      sig_energy = noise_energy = 0.;
    }
    if (acc_thread) {
      if (!above) {
        if (status == kStNormal) {
          snum = num;
          sden = den;
        }
      }
      const int st_new = above ? kStNormal : (status == kStNormal ? kStTentative : status);
      if (st_new != kStInit) {
        if (b == 0) {
          const double nmr = (red[par][c][0] + red[par][c][1]) / B;
          num += 1. * (10. * log10(nmr));                      // segmental: dB per frame (movs.c:1017-1018)
          den += 1.;
        } else if (flags & kRecFlagEhsValid) {
          num += 1. * (1000. * rec[L.off_ehs + c]);            // movs.c:1441
          den += 1.;
        }
      }
    }
    if (!above) {
      if (status == kStNormal) status = kStTentative;
    } else {
      status = kStNormal;
      if (first_above == 0) first_above = frame_counter + 1;
      if (frame_counter >= acc_start) owned_above = 1;
    }
    sig_energy += rec[L.off_snr];
    noise_energy += rec[L.off_snr + 1];
    frame_counter++;
  }

  if (active) st[S.off_fft_filtered + c * B + b] = Efr;
  if (acc_thread) {
    double* a = st + S.off_fft_acc + (c * 2 + b) * kAccFields;
    a[0] = num; a[1] = den; a[5] = snum; a[6] = sden;
    const bool tent = status == kStTentative;
    val_sh[c][b] = (tent ? snum : num) / (tent ? sden : den);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    st_ints[0] = status;
    st_ints[1] = (int)frame_counter;
    st_ints[kASegOwnedAboveFft] = owned_above;
    st_ints[kASegFirstAboveFft] = (int)first_above;
    st[S.off_fft_scalar] = sig_energy;
    st[S.off_fft_scalar + 1] = noise_energy;
    double movs[5];
    movs[0] = st[S.off_fb_movs];        // RmsModDiffA
    movs[1] = st[S.off_fb_movs + 1];    // RmsNoiseLoudAsymA
    movs[4] = st[S.off_fb_movs + 2];    // AvgLinDistA
    double v2 = 0., v3 = 0.;
    for (int cc = 0; cc < C; cc++) {
      v2 += val_sh[cc][0];
      v3 += val_sh[cc][1];
    }
    movs[2] = v2 / C;                   // SegmentalNMRB
    movs[3] = v3 / C;                   // EHSB
    // peaq_calculate_di_advanced (nn.c:304-335)
    double x[5];
    for (int j = 0; j < 5; j++) x[j] = T->nn_wxb[j];
    for (int i = 0; i < 5; i++) {
      const double m = (movs[i] - T->nn_amin[i]) / (T->nn_amax[i] - T->nn_amin[i]);
      for (int j = 0; j < 5; j++) x[j] += T->nn_wx[i * 5 + j] * m;
    }
    double di = T->nn_wyb;
    for (int j = 0; j < 5; j++) di += T->nn_wy[j] / (1 + exp(-x[j]));
    PairResult& r = results[pair];
    r.di = di;
    r.odg = -3.98 + (0.22 - -3.98) / (1 + exp(-di));
    r.totalsnr = 10 * log10(sig_energy / noise_energy);
    for (int i = 0; i < 11; i++) r.movs[i] = i < 5 ? movs[i] : 0.;
    r.n_movs = 5;
    r.frames_fft = frame_counter;
    r.frames_fb = (unsigned)st_ints[3];
    r.loudness_reached_frame = (unsigned)st_ints[4];
  }
}

// fresh advanced state: zeros, loudness latch unset (gstpeaq.c:359), history slot 10
__global__ void init_adv_state_kernel(double* state, AdvStateLayout S, int n_pairs) {
  const size_t total = (size_t)n_pairs * S.stride;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % S.stride);
    if (o == S.off_ints + 2) {
      int* ints = reinterpret_cast<int*>(state + i);
      ints[0] = (int)UINT_MAX;   // slot 4: loudness_reached_frame
      ints[1] = 0;
    } else {
      state[i] = 0.;
    }
  }
}

}  // namespace

cudaError_t launch_init_adv_state(double* state, AdvStateLayout S, int n_pairs, cudaStream_t stream) {
  if (n_pairs <= 0) return cudaSuccess;
  const size_t total = (size_t)n_pairs * S.stride;
  const unsigned blocks = (unsigned)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  init_adv_state_kernel<<<blocks, 256, 0, stream>>>(state, S, n_pairs);
  return cudaGetLastError();
}

cudaError_t launch_fb_spread(const DeviceTables* d_tables, const double* fbout, unsigned n_sub,
                             const unsigned* n_frames, unsigned first_frame, double* state,
                             AdvStateLayout S, double* energy, int n_pairs, cudaStream_t stream) {
  if (n_pairs <= 0 || n_sub == 0) return cudaSuccess;
  fb_spread_kernel<<<n_pairs * 2 * S.C, 2 * kSpTile, 0, stream>>>(
      d_tables, reinterpret_cast<const double2*>(fbout), n_sub, n_frames, first_frame, state, S, energy);
  return cudaGetLastError();
}

cudaError_t launch_fb_scan(const DeviceTables* d_tables, const double* energy, unsigned n_sub,
                           const unsigned char* flags, const unsigned* n_frames, unsigned first_frame,
                           unsigned n_chunk_frames, double* state, AdvStateLayout S, double* dbg,
                           int n_pairs, cudaStream_t stream) {
  if (n_pairs <= 0) return cudaSuccess;
  fb_scan_kernel<<<n_pairs, 64 * S.C, 0, stream>>>(d_tables, energy, n_sub, flags, n_frames, first_frame,
                                                    n_chunk_frames, state, S, dbg);
  return cudaGetLastError();
}

cudaError_t launch_adv_fft_scan(const DeviceTables* d_tables, const double* records, RecordLayout L,
                                const unsigned* n_frames, unsigned first_frame, unsigned n_chunk_frames,
                                double* state, AdvStateLayout S, PairResult* results, int n_pairs,
                                cudaStream_t stream) {
  if (n_pairs <= 0) return cudaSuccess;
  adv_fft_scan_kernel<<<n_pairs, kAdvGroup * L.C, 0, stream>>>(d_tables, records, L, n_frames, first_frame,
                                                               n_chunk_frames, state, S, results);
  return cudaGetLastError();
}

}  // namespace peaq
