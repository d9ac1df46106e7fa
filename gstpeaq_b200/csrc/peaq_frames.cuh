// Device code of the stateless, frame-parallel half of the PEAQ hot path (`frame_body`), shared by
// K1 `fft_frames_kernel` (peaq_frames.cu: one CTA per frame, results to per-frame records in HBM)
// and the fused persistent kernel (peaq_fused.cu: one CTA per pair, results stay in shared memory
// for the recurrent half of the same frame).
//
// One CTA = one FFT-clock frame of one (ref,test) pair; 4*C warps, two warps per
// stream (channel c, side: 0 ref, 1 test).  Per stream:
//   PCM (interleaved F32, HBM, TMA bulk copy) -> Hann window -> 2048-pt real FFT
//   (1024-pt complex radix-4 in shared memory + split) -> power spectrum ->
//   outer/middle-ear weighting -> critical-band grouping -> + internal noise ->
//   level dependent frequency spreading      (fftearmodel.c:432-515, :603-676)
// and per channel: noise spectrum grouped into bands (movs.c:988-1000), bandwidth
// bins (movs.c:776-809), error-harmonic-structure value (movs.c:1346-1443), energy
// and above-threshold flags (fftearmodel.c:508-514, gstpeaq.c:1081-1099) and the
// SNR partial sums (gstpeaq.c:913-918).  Everything recurrent is left to the scan.
//
// Shared memory: 16 KB per stream (the FFT buffer, reused for the weighted power
// spectrum and the spreading / EHS scratch) + 8 KB of twiddles => 3 CTAs per SM.
// The power spectrum is staged in registers between the FFT read-out and the
// write-back, and the bandwidth decisions are taken on those registers, so the
// unweighted spectrum never needs its own buffer.
//
// Arithmetic is IEEE double; the including files are compiled with -fmad=false so that
// expressions restated from the reference round like the reference's (gcc,
// x86-64, no contraction); fused multiply-adds appear only where written
// explicitly (FFT butterflies, the transcendental kernels of peaq_math.cuh).
#pragma once

#include "peaq_engine.h"
#include "peaq_fft.cuh"
#include "peaq_math.cuh"

#include <cstdlib>
#include <type_traits>

namespace peaq {
namespace {

constexpr int kWorkDoubles = 2048;   // per-warp buffer: 1024 complex points
constexpr int kTwDoubles = 2 * 512;
constexpr int kScratchDlog = 1536;   // dlog[512] (test stream's buffer)

#ifndef PEAQ_WARP_SUM_DEFINED
#define PEAQ_WARP_SUM_DEFINED
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif

__device__ __forceinline__ double warp_max_nonan(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double x = __shfl_xor_sync(0xffffffffu, v, o);
    v = x > v ? x : v;
  }
  return v;
}

__device__ __forceinline__ int warp_max_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int x = __shfl_xor_sync(0xffffffffu, v, o);
    v = x > v ? x : v;
  }
  return v;
}

// ---- TMA (cp.async.bulk) staging of one frame of interleaved PCM into shared memory ----
__device__ __forceinline__ unsigned smem_addr(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_global, unsigned bytes,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_addr(dst_smem)),
      "l"(src_global), "r"(bytes), "r"(smem_addr(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

// sample `i` of channel `c` of an interleaved signal, zero past the end
__device__ __forceinline__ float pcm_at(const float* __restrict__ sig, unsigned long long s0,
                                        unsigned long long n_samples, int i, int c, int C) {
  const unsigned long long s = s0 + (unsigned long long)i;
  return s < n_samples ? __ldg(sig + s * (unsigned long long)C + c) : 0.0f;
}

// literal replay of is_frame_above_threshold for one channel (gstpeaq.c:1088-1096):
// FLOAT running sum, double increments; run by one thread, only for borderline frames
__device__ bool replay_threshold_serial(const float* __restrict__ sig, unsigned long long s0,
                                        unsigned long long n_samples, int c, int C) {
  const double thr = 200. / 32768;
  float sum = 0;
  int i;
  for (i = 0; i < 5; i++)
    sum = (float)((double)sum + fabs((double)pcm_at(sig, s0, n_samples, i, c, C)));
  while (i < kFftFrame) {
    sum = (float)((double)sum + (fabs((double)pcm_at(sig, s0, n_samples, i, c, C)) -
                                 fabs((double)pcm_at(sig, s0, n_samples, i - 5, c, C))));
    if ((double)sum >= thr) return true;
    i++;
  }
  return false;
}

// peaq_fftearmodel_group_into_bands for band i (fftearmodel.c:603-620); `spec`
// already holds the weighted power spectrum
__device__ __forceinline__ double group_band(const DeviceTables* __restrict__ T, const double* spec, int i) {
  const int lo = T->band_lo[i], hi = T->band_hi[i];
  double p = T->band_wl[i] * spec[lo] + T->band_wu[i] * spec[hi];
  for (int k = lo + 1; k < hi; k++) p += spec[k];
  return p < 1e-12 ? 1e-12 : p;
}

// noise spectrum bin (movs.c:993-998)
__device__ __forceinline__ double noise_bin(const double* spec_ref, const double* spec_test, int k) {
  const double r = spec_ref[k], t = spec_test[k];
  return r - 2 * peaq_sqrt(r * t) + t;
}

// noise_in_bands[i] (movs.c:999-1000): the grouping of fftearmodel.c:603-620 applied to the noise
// spectrum.  `nz[k]` = noise_bin(k), computed one bin per thread beforehand (the square roots are
// the cost and band widths range from 1 to ~50 bins, so a lane per band would idle most lanes);
// the sum runs in the reference's order: wl N[lo] + wu N[hi] + N[lo+1] + ... + N[hi-1].
__device__ __forceinline__ double group_band_noise(const DeviceTables* __restrict__ T, const double* nz, int i) {
  const int lo = T->band_lo[i], hi = T->band_hi[i];
  double p = T->band_wl[i] * nz[lo];
  p = p + T->band_wu[i] * nz[hi];   // both edge terms exist even when lo == hi
  for (int k = lo + 1; k < hi; k++) p = p + nz[k];
  return p < 1e-12 ? 1e-12 : p;
}

// Level dependent frequency spreading (do_spreading, fftearmodel.c:636-676), in two parts.
//
// spread_prepare (one warp per stream): slopes, normalisation and the downward pass.
// `se` holds the pitch pattern on entry; on return the stream's scratch holds
//   sa[i] = aUCE[i]^0.4, se[i] = En[i]^0.4, se2[i] = downward-spread pattern (all in the 0.4 domain).
__device__ void spread_prepare(const DeviceTables* __restrict__ T, int B, double* sa, double* se,
                               double* se2, int lane) {
  const double dz02 = 0.2 * T->dz;
  // Four bands per lane, one after the other: unrolling this loop (24 inlined ln / exp) and the
  // ratio loop below bought nothing -- the frame kernel is not bound by a warp's own instruction
  // latency but by the instruction caches and the shared-memory pipe of the SM --
  // and cost 1500 instructions of code (149.7 -> 146.3 ms per 4096 x 10 s rolled).  For the same
  // reason the branch-free division / square root / ln / exp of peaq_math.cuh, which speed the
  // latency-bound scan kernels up by 18 %, are NOT used here: measured 149.7 -> 164.1 ms.
#pragma unroll 1
  for (int m = 0; m < 4; m++) {
    const int i = lane + 32 * m;
    const int ii = i < B ? i : B - 1;
    // aUCE = aUC * Pp^(0.2 dz); gIU = (1 - aUCE^(B-i)) / (1 - aUCE);
    // En = Pp / (gIL + gIU - 1); store aUCE^0.4 and En^0.4   (:647-656)
    const double pp = i < B ? se[i] : 1.;
    const double lp = peaq_log(pp);
    const double la = T->log_aUC[ii] + dz02 * lp;     // ln aUCE
    const double a_uce = peaq_exp(la);
    const double g_iu = (1. - peaq_exp((double)(B - ii) * la)) / (1. - a_uce);
    const double den = T->gIL[ii] + g_iu - 1.;
    const double va = peaq_exp(0.4 * la);
    const double ve = peaq_exp(0.4 * (lp - peaq_log(den)));
    if (i < B) {   // each lane rewrites only the entry it read
      sa[i] = va;
      se[i] = ve;
    }
  }
  __syncwarp();
  // downward spreading, constant slope: E2[i-1] = aLe * E2[i] + Ene[i-1]   (:658-661)
  // = sum_{j >= i} aLe^(j-i) Ene[j]: lane l scans bands 4l..4l+3 locally, the lane
  // carries are combined by a 5-step shuffle scan with the slope raised to 4 * 2^s
  {
    const double a_le = T->aLe;
    double v[4];
#pragma unroll
    for (int m = 0; m < 4; m++) v[m] = 4 * lane + m < B ? se[4 * lane + m] : 0.;
    v[2] = v[2] + a_le * v[3];
    v[1] = v[1] + a_le * v[2];
    v[0] = v[0] + a_le * v[1];
    const double a2 = a_le * a_le, a4 = a2 * a2;
    double carry = v[0];      // sum over this lane's bands, referred to band 4l
    double q = a4;            // slope over one lane (4 bands), squared every step
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_down_sync(0xffffffffu, carry, o);
      if (lane + o < 32) carry = carry + q * up;
      q = q * q;
    }
    // contribution of all higher lanes, referred to band 4(l+1)
    double hi = __shfl_down_sync(0xffffffffu, carry, 1);
    if (lane == 31) hi = 0.;
    const double a3 = a2 * a_le;
    if (4 * lane + 3 < B) se2[4 * lane + 3] = v[3] + a_le * hi;
    if (4 * lane + 2 < B) se2[4 * lane + 2] = v[2] + a2 * hi;
    if (4 * lane + 1 < B) se2[4 * lane + 1] = v[1] + a3 * hi;
    if (4 * lane + 0 < B) se2[4 * lane + 0] = v[0] + a4 * hi;
  }
  __syncwarp();
}

// spread_ladder (one warp per stream): upward spreading (:664-671), source band i adds
// Ene[i] * aUCEe[i]^(j-i) to every band j > i -- O(B^2 / 2) multiply-adds whose cost is the data
// movement, not the arithmetic.  A lane owns the four consecutive bands p = 4 l + m as SOURCES
// (r[m] = running power Ene * a^t, the only multiplications) and four TRAVELLING accumulators:
// at step t accumulator slot p holds the partial sum of target band p + t, takes its source's
// term, and moves one slot down.  Inside a lane that move is a register renaming (the loop is
// unrolled by 4); only slot 0 of each lane crosses to the lane below: ONE shuffle per lane and
// step instead of one per band, and no per-band select.  The accumulator leaving slot 0 of lane 0
// at step t is the finished band t.  Every target still receives its terms nearest source first,
// starting from the downward pattern, each term the same chain of products as before: results are
// bit for bit those of the one-shuffle-per-band ladder this replaces (13 instead of 26
// instructions per step).  On return the complete spread pattern (0.4 domain) is band 0 in se2[0]
// and band i >= 1 in scr[128 + i - 1] (spread_result).
__device__ __forceinline__ double spread_result(const double* scr, int i) { return i ? scr[128 + i - 1] : scr[256]; }

__device__ void spread_ladder(int B, double* scr, int lane) {
  const double* sa = scr;
  const double* se = scr + 128;
  double* se2 = scr + 256;
  double r[4], a[4], acc[4];
#pragma unroll
  for (int m = 0; m < 4; m++) {
    const int p = 4 * lane + m;
    r[m] = p < B ? se[p] : 0.;
    a[m] = p < B ? sa[p] : 0.;
    acc[m] = p + 1 < B ? se2[p + 1] : 0.;
  }
  __syncwarp();   // every lane holds its inputs: se (dead from here on) receives the finished bands
  // Four steps per trip: lane 0 keeps the four bands that finish and writes them with two 128-bit
  // stores, band t to fin[t - 1] (the shift by one makes the groups of four 16-byte aligned; one
  // shared-memory wavefront per two bands instead of one per band).  The last trip may run up to
  // three steps past band B - 1: they only produce entries nobody reads (fin has 128 entries).
  double* fin = scr + 128;
  for (int t = 1; t < B; t += 4) {
    double done[4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
#pragma unroll
      for (int m = 0; m < 4; m++) {
        r[m] *= a[m];
        acc[m] += r[m];
      }
      done[s] = acc[0];
      double in = __shfl_down_sync(0xffffffffu, acc[0], 1);
      if (lane == 31) in = 0.;
#pragma unroll
      for (int m = 0; m < 3; m++) acc[m] = acc[m + 1];
      acc[3] = in;
    }
    if (lane == 0) {
      *reinterpret_cast<double2*>(fin + t - 1) = make_double2(done[0], done[1]);
      *reinterpret_cast<double2*>(fin + t + 1) = make_double2(done[2], done[3]);
    }
  }
  __syncwarp();
}


// barrier of the two warps of one stream (ids 1..4), of the four warps of one channel
// (ids 5, 6) and the arrive/wait pair that guards the ref buffer (ids 7, 8)
struct StreamSync {
  int id;
  __device__ __forceinline__ void operator()() const { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
};

// Error harmonic structure of one channel (peaq_mov_ehs, movs.c:1383-1441),
// run by the two warps of the channel's test stream (t = thread within the stream, 0..63): the
// 512- and 256-point transforms and the loops around them are shared; the last, 128-point
// transform and the peak search are too small for that and stay on the first warp (which returns
// the value).  `xch`: 6 doubles of shared memory for the cross-warp sums.  d[0..511] =
// ln(Pw_test/Pw_ref) is in `dlog` (shared); work: the first 1536 doubles of the ref stream's buffer.
//
// Transform sizes are halved wherever a sequence is real:
//   - both forward transforms of do_xcorr (movs.c:1300-1303) are packed into ONE
//     512-point complex FFT (real part d[0..511], imaginary part d[0..255] | 0);
//   - the inverse of the Hermitian product (movs.c:1304-1313) is a 256-point complex
//     FFT of the even/odd recombined spectrum;
//   - the final 256-point real transform (movs.c:1428) is a 128-point complex FFT.
__device__ double ehs_channel_pair(const DeviceTables* __restrict__ T, const double* dlog, double* work,
                                   const double2* __restrict__ tw, int t, StreamSync sync, double* xch) {
  const int lane = t & 31, half = t >> 5;
  double2* za = reinterpret_cast<double2*>(work);         // 512 complex
  double2* zb = reinterpret_cast<double2*>(work) + 512;   // 256 complex (+ 129 doubles of |C|^2)
  const int sl9 = fft_slot_rt<9>(t);
  const int sl8 = fft_slot_rt<8>(t);
  // The thread's points n = t + 64 u: for fixed parity j of u the four u = j + 2 m are the inputs
  // of one butterfly of the first radix-4 level (top digit of n >> 1, no twiddles), which therefore
  // runs here in registers and its outputs go where the scatter would have put the inputs.
#pragma unroll
  for (int j = 0; j < 2; j++) {
    double2 a0, a1, a2, a3;
    {
      const double v0 = dlog[t + 64 * j], v1 = dlog[t + 64 * (j + 2)];
      a0 = make_double2(v0, v0);
      a1 = make_double2(v1, v1);
      a2 = make_double2(dlog[t + 64 * (j + 4)], 0.);
      a3 = make_double2(dlog[t + 64 * (j + 6)], 0.);
    }
    fft_bfly4_nt(a0, a1, a2, a3);
    za[sl9 ^ fft_slot<9>(64 * j)] = a0;
    za[sl9 ^ fft_slot<9>(64 * (j + 2))] = a1;
    za[sl9 ^ fft_slot<9>(64 * (j + 4))] = a2;
    za[sl9 ^ fft_slot<9>(64 * (j + 6))] = a3;
  }
  sync();
  group_fft<9, 64, StreamSync, 4>(za, tw, t, sync);
  auto g_of = [&](int q) {
    const double2 p = za[fft_swz(q & 511)];
    const double2 m = za[fft_swz((512 - q) & 511)];
    const double f1r = 0.5 * (p.x + m.x), f1i = 0.5 * (p.y - m.y);
    const double f2r = 0.5 * (p.y + m.y), f2i = -0.5 * (p.x - m.x);
    return make_double2((f1r * f2r + f1i * f2i) / (2 * kMaxLag), (f2r * f1i - f1r * f2i) / (2 * kMaxLag));
  };
  // (the thread's four points k = t + 64 u are one butterfly of the first radix-4 level: registers)
  double2 y[4];
#pragma unroll
  for (int u = 0; u < 4; u++) {
    const int k = t + 64 * u;
    const double2 a = g_of(k), b = g_of(256 - k);
    const double er = a.x + b.x, ei = a.y - b.y;
    const double dr = a.x - b.x, di = a.y + b.y;
    const double2 w = tw[fft_twi(2 * t) ^ fft_twi(128 * u)];   // e^{-2 pi i k / 512}; its conjugate is needed
    const double orr = dr * w.x + di * w.y, oi = di * w.x - dr * w.y;
    y[u] = make_double2(er - oi, -(ei + orr));   // conj(E + i O)
  }
  fft_bfly4_nt(y[0], y[1], y[2], y[3]);
#pragma unroll
  for (int u = 0; u < 4; u++) zb[sl8 ^ fft_slot<8>(64 * u)] = y[u];
  sync();
  group_fft<8, 64, StreamSync, 4>(zb, tw, t, sync);
  // thread owns lags i = 4 t .. 4 t + 3: c[2m] = Re, c[2m+1] = -Im of element m = 2 t + p
  double c[4];
#pragma unroll
  for (int p = 0; p < 2; p++) {
    const double2 v = zb[fft_swz(2 * t + p)];
    c[2 * p] = v.x;
    c[2 * p + 1] = -v.y;
  }
  // c[i] /= sqrt(d0 * dk_i), dk_i = d0 + sum_{j<i} (d[j+256]^2 - d[j]^2)   (movs.c:1405-1418)
  double term[4];
  double local = 0.;
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const int i = 4 * t + e;
    const double hi = dlog[i + kMaxLag], lo = dlog[i];
    term[e] = hi * hi - lo * lo;
    local += term[e];
  }
  double incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double x = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += x;
  }
  if (t == 0) xch[0] = c[0];
  if (t == 31) xch[1] = incl;   // total of the first warp
  sync();
  const double d0 = xch[0];
  double dk = d0 + ((incl - local) + (half ? xch[1] : 0.));
  double csum = 0.;
#pragma unroll
  for (int e = 0; e < 4; e++) {
    c[e] = c[e] * rsqrt(d0 * dk);   // c / sqrt(d0 dk) (movs.c:1415); same zero / NaN classes
    csum += c[e];
    dk += term[e];
  }
  csum = warp_sum(csum);
  if (lane == 0) xch[2 + half] = csum;
  sync();
  const double cavg = (xch[2] + xch[3]) / kMaxLag;
  // subtract mean, window (movs.c:1419-1421); 256 real points as 128 complex ones
#pragma unroll
  for (int p = 0; p < 2; p++) {
    const int i = 4 * t + 2 * p;
    za[fft_slot_rt<7>(2 * t + p)] =
        make_double2((c[2 * p] - cavg) * T->ehs_window[i], (c[2 * p + 1] - cavg) * T->ehs_window[i + 1]);
  }
  sync();
  if (half) return 0.;
  warp_fft<7>(za, tw, lane);
  double* s2 = reinterpret_cast<double*>(zb);
  for (int k = lane; k <= kMaxLag / 2; k += 32) {
    const double2 p = za[fft_swz(k & 127)];
    const double2 m = za[fft_swz((128 - k) & 127)];
    const double er = 0.5 * (p.x + m.x), ei = 0.5 * (p.y - m.y);
    const double orr = 0.5 * (p.y + m.y), oi = -0.5 * (p.x - m.x);
    const double2 w = fft_tw(tw, 4 * k);
    const double xr = er + (orr * w.x - oi * w.y);
    const double xi = ei + (orr * w.y + oi * w.x);
    s2[k] = xr * xr + xi * xi;
  }
  __syncwarp();
  double best = 0.;
  for (int k = 1 + lane; k <= kMaxLag / 2; k += 32) {
    const double sv = s2[k], sp = s2[k - 1];
    if (sv > sp && sv > best) best = sv;
  }
  return warp_max_nonan(best);
}

// ---- kernel -------------------------------------------------------------------------------
// Per-stream buffer (2048 doubles), after the FFT:
//   [0, 769)      weighted power spectrum, bins 0..768 (nothing reads higher bins: grouping stops
//                 at 18 kHz = bin 768, the EHS at bin 511; the bandwidth decisions use registers)
//   test buffer:  [776, 1160) spreading scratch of the test stream, [1536, 2048) ln spectrum ratio
//   ref buffer:   [1536, 1920) spreading scratch of the ref stream; [0, 1536) is recycled by the
//                 EHS transforms once every reader of the ref spectrum has arrived
constexpr int kSpecBins = 769;
constexpr int kScratchTest = 776;
constexpr int kScratchRef = 1536;

struct FrameMail {
  double thr_part[kMaxChannels][2];
  double snr_s[kMaxChannels][2], snr_n[kMaxChannels][2];
  double energy[2 * kMaxChannels][2];
  float maxabs[kMaxChannels][2];
  int bw_ref_part[kMaxChannels][2];
  int bw_test_part[kMaxChannels][2];
  double ehs_x[kMaxChannels][6];
  unsigned long long mbar;
  // results of the frame that are not band arrays (fused kernel: read by the scan step)
  double o_ehs[kMaxChannels];
  double o_snr[2];
  int o_flags;
  int o_bw[kMaxChannels][2];
};

// Shared-memory map of a CTA (doubles): [0, kTwDoubles) twiddles, then one kWorkDoubles buffer per
// stream (stream = 2 * channel + side), then the FrameMail.
__device__ __forceinline__ double* frame_stream_buf(double* smem, int stream) {
  return smem + kTwDoubles + stream * kWorkDoubles;
}
__device__ __forceinline__ FrameMail* frame_mail(double* smem, int C) {
  return reinterpret_cast<FrameMail*>(smem + kTwDoubles + 2 * C * kWorkDoubles);
}
// where frame_body<true> leaves the band arrays of channel `chan` (valid until the buffers are reused)
constexpr int kNoiseOut = 1160;   // test buffer: noise in bands
__device__ __forceinline__ const double* frame_out_e2(double* smem, int chan, int side) {
  return side == 0 ? frame_stream_buf(smem, 2 * chan) + kScratchRef + 256
                   : frame_stream_buf(smem, 2 * chan + 1) + kScratchTest + 256;
}
__device__ __forceinline__ const double* frame_out_noise(double* smem, int chan) {
  return frame_stream_buf(smem, 2 * chan + 1) + kNoiseOut;
}

__device__ __forceinline__ void chan_sync(int chan) {
  asm volatile("bar.sync %0, 128;" ::"r"(5 + chan) : "memory");
}
__device__ __forceinline__ void refbuf_arrive(int chan) {
  asm volatile("bar.arrive %0, 128;" ::"r"(7 + chan) : "memory");
}
__device__ __forceinline__ void refbuf_wait(int chan) {
  asm volatile("bar.sync %0, 128;" ::"r"(7 + chan) : "memory");
}
__device__ __forceinline__ void frame_load_twiddles(const DeviceTables* __restrict__ T, double* smem) {
  double2* tw = reinterpret_cast<double2*>(smem);   // 512 complex
  for (int i = threadIdx.x; i < 512; i += blockDim.x)
    tw[fft_twi(i)] = make_double2(T->tw1024[i].x, T->tw1024[i].y);   // swizzled table (peaq_fft.cuh)
}

// Whole frame inside both signals and 16-byte aligned: it is staged with two TMA bulk copies
// (ref -> stream 0's buffer, test -> stream 1's buffer, both unused at that point); otherwise
// (last, zero-padded frame; odd strides) frame_body fills the same buffers with a guarded copy.
__device__ __forceinline__ bool frame_tma_ok(const PcmView& pcm, int pair, unsigned frame) {
  const unsigned long long s0 = (unsigned long long)frame * kFftStep;
  const float* ref_sig = pcm.ref + pcm_pair_offset(pcm, pair);
  const float* test_sig = pcm.test + pcm_pair_offset(pcm, pair);
  return (reinterpret_cast<uintptr_t>(ref_sig) & 15) == 0 && (reinterpret_cast<uintptr_t>(test_sig) & 15) == 0 &&
         s0 + kFftFrame <= pcm.n_samples[pair] && s0 + kFftFrame <= pcm.n_samples_test[pair];
}
// one thread: arm the barrier and start both copies (the mbarrier must have been initialised and
// the two buffers must be free)
__device__ __forceinline__ void frame_tma_issue(const PcmView& pcm, int pair, unsigned frame, double* smem) {
  const int C = pcm.channels;
  FrameMail* mail = frame_mail(smem, C);
  const unsigned long long s0 = (unsigned long long)frame * kFftStep;
  const float* ref_sig = pcm.ref + pcm_pair_offset(pcm, pair);
  const float* test_sig = pcm.test + pcm_pair_offset(pcm, pair);
  const unsigned bytes = kFftFrame * C * sizeof(float);
  mbar_expect_tx(&mail->mbar, 2 * bytes);
  tma_load_1d(frame_stream_buf(smem, 0), ref_sig + s0 * C, bytes, &mail->mbar);
  tma_load_1d(frame_stream_buf(smem, 1), test_sig + s0 * C, bytes, &mail->mbar);
}

// One CTA = one FFT-clock frame of one pair; 4 C warps.  Two warps share a stream (channel c,
// side: 0 ref, 1 test) up to the weighted power spectrum -- staging, window, FFT, power
// spectrum, bandwidth -- so every thread holds 16 bins instead of 32 and a CTA needs half the
// registers per thread: 24 warps per SM instead of 12 for the same shared memory.  After
// that the four warps of a channel split into tasks:
//   ref-h0 : grouping + slopes of the ref stream, then the upward ladder of BOTH streams
//   ref-h1 : grouping + slopes of the test stream
//   test-h0/h1: ln spectrum ratio and noise spectrum (one bin per thread), noise in bands,
//               then both run the EHS together
// kFused = false: results go to the per-frame record `rec` (global memory);
// kFused = true : band arrays stay in shared memory (frame_out_e2 / frame_out_noise), scalars in
//                 the FrameMail; the caller synchronises the CTA before reading them.
// tma_ok / tma_parity: the caller has issued frame_tma_issue for this frame (phase parity of the
// mbarrier) -- or not, then the frame is copied here.  The twiddles must be (being) loaded.
template <bool kFused>
__device__ __forceinline__ void frame_body(const DeviceTables* __restrict__ T, const PcmView& pcm, int pair,
                                           unsigned frame, int B, int advanced, double* smem, bool tma_ok,
                                           unsigned tma_parity, double* __restrict__ rec, const RecordLayout& L) {
  const int C = pcm.channels;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int stream = warp >> 1;       // 2 * channel + side
  const int half = warp & 1;
  const int chan = stream >> 1;
  const int side = stream & 1;
  const int t = half * 32 + lane;     // thread within the stream

  double2* tw = reinterpret_cast<double2*>(smem);                       // 512 complex
  double* work = frame_stream_buf(smem, stream);
  FrameMail* mail = frame_mail(smem, C);

  const unsigned long long n_ref = pcm.n_samples[pair], n_test = pcm.n_samples_test[pair];
  const unsigned long long s0 = (unsigned long long)frame * kFftStep;
  const float* __restrict__ ref_sig = pcm.ref + pcm_pair_offset(pcm, pair);
  const float* __restrict__ test_sig = pcm.test + pcm_pair_offset(pcm, pair);
  float* raw_ref = reinterpret_cast<float*>(frame_stream_buf(smem, 0));
  float* raw_test = reinterpret_cast<float*>(frame_stream_buf(smem, 1));
  if (!tma_ok) {
    // guarded copy of the frame (zeros past the end of either signal) into the same buffers
    const unsigned long long base = s0 * (unsigned long long)C;
    for (int i = threadIdx.x; i < kFftFrame * C; i += blockDim.x) {
      raw_ref[i] = base + i < n_ref * C ? __ldg(ref_sig + base + i) : 0.f;
      raw_test[i] = base + i < n_test * C ? __ldg(test_sig + base + i) : 0.f;
    }
    __syncthreads();
  }

  // ---- phase 1: this thread's 32 samples (complex points n = t + 64 u) into registers ----
  float xs0[16], xs1[16];
  double energy = 0., es = 0., en = 0.;
  if (tma_ok) mbar_wait(&mail->mbar, tma_parity);
  {
    const float* raw = side ? raw_test : raw_ref;
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const int n = t + 64 * u;   // samples 2n, 2n+1
      if (C == 2) {
        const float4 v = *reinterpret_cast<const float4*>(raw + 4 * n);
        xs0[u] = chan ? v.y : v.x;
        xs1[u] = chan ? v.w : v.z;
      } else {
        const float2 v = *reinterpret_cast<const float2*>(raw + 2 * n);
        xs0[u] = v.x;
        xs1[u] = v.y;
      }
      if (side == 1 && u < 8) {
        // SNR partial sums over the first half of the frame (gstpeaq.c:913-918)
        float r0, r1;
        if (C == 2) {
          const float4 v = *reinterpret_cast<const float4*>(raw_ref + 4 * n);
          r0 = chan ? v.y : v.x;
          r1 = chan ? v.w : v.z;
        } else {
          const float2 v = *reinterpret_cast<const float2*>(raw_ref + 2 * n);
          r0 = v.x;
          r1 = v.y;
        }
        es += (double)(r0 * r0);
        es += (double)(r1 * r1);
        en += (double)((r0 - xs0[u]) * (r0 - xs0[u]));
        en += (double)((r1 - xs1[u]) * (r1 - xs1[u]));
      }
    }
  }
  __syncthreads();   // staged PCM consumed: the buffers become FFT work space

  // ---- phase 2: window, first FFT pass, scatter into FFT order; energy; largest |x| ----
  double2* z = reinterpret_cast<double2*>(work);
  {
    const int slot_t = fft_slot_rt<10>(t);
    float amax = 0.f;
    // The thread's 16 points are n = t + 64 j + 256 m: for fixed j the four m are exactly the
    // inputs of one butterfly of the FIRST radix-4 pass (top digit of n = lowest digit of the
    // digit-reversed position, no twiddles), so that pass happens here in registers and the
    // results go straight to the positions the scatter would have filled.
#pragma unroll
    for (int j = 0; j < 4; j++) {
      double2 a[4];
#pragma unroll
      for (int m = 0; m < 4; m++) {
        const int u = j + 4 * m;
        const int n = t + 64 * u;
        const float x0 = xs0[u], x1 = xs1[u];
        const double2 h = *reinterpret_cast<const double2*>(&T->hann[2 * n]);
        a[m] = make_double2(h.x * x0, h.y * x1);
        if (u >= 8) {   // samples 1024..2047: float products, double accumulation (fftearmodel.c:508-511)
          energy += (double)(x0 * x0);
          energy += (double)(x1 * x1);
        }
        // sample 0 never enters a tested window (gstpeaq.c:1088-1096: windows end at i >= 5)
        if (n > 0) amax = fmaxf(amax, fabsf(x0));
        amax = fmaxf(amax, fabsf(x1));
      }
      const double2 t0 = make_double2(a[0].x + a[2].x, a[0].y + a[2].y);
      const double2 t1 = make_double2(a[0].x - a[2].x, a[0].y - a[2].y);
      const double2 t2 = make_double2(a[1].x + a[3].x, a[1].y + a[3].y);
      const double2 t3 = make_double2(a[1].y - a[3].y, a[3].x - a[1].x);   // -i (a1 - a3)
      z[slot_t ^ fft_slot<10>(64 * j)] = make_double2(t0.x + t2.x, t0.y + t2.y);
      z[slot_t ^ fft_slot<10>(64 * (j + 4))] = make_double2(t1.x + t3.x, t1.y + t3.y);
      z[slot_t ^ fft_slot<10>(64 * (j + 8))] = make_double2(t0.x - t2.x, t0.y - t2.y);
      z[slot_t ^ fft_slot<10>(64 * (j + 12))] = make_double2(t1.x - t3.x, t1.y - t3.y);
    }
    energy = warp_sum(energy);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (side == 1) {
      es = warp_sum(es);
      en = warp_sum(en);
    }
    if (lane == 0) {
      mail->energy[stream][half] = energy;
      if (side == 0) mail->maxabs[chan][half] = amax;
      else {
        mail->snr_s[chan][half] = es;
        mail->snr_n[chan][half] = en;
      }
    }
  }
  __syncthreads();   // twiddles loaded, mail published, scatter of both halves complete

  // ---- 2048-point real FFT by the stream's 64 threads; power spectrum into registers ----
  // Levels 4 + 16 in one trip through shared memory, levels 64 + 256 in a second one that ends in
  // registers (peaq_fft.cuh): thread t then holds Z[t + 64 u], u < 16.
  // Z[k] and Z[1024-k] give X[k] = E + P and X[1024-k] = conj(E - P) (E, O the even/odd parts,
  // P = O w^k), so one pair of values and one twiddle product serve two bins: register u < 8
  // holds bin t + 64 u, register 8 + u its mirror 1024 - (t + 64 u); bin 512 (its own mirror)
  // is an extra of thread 0.  Z[k], k = t + 64 u < 512, is the thread's own e[u]; the mirror
  // Z[1024 - k] = Z[(64 - t) + 64 (15 - u)] is e[15 - u] of thread 64 - t, so only the upper
  // halves e[8..15] go through shared memory once more: row u - 8, column (t - 1) & 63 of a plain
  // 8 x 64 exchange array (consecutive lanes, consecutive 16-byte slots, for the store and for the
  // mirrored load alike).  Thread 0 is its own partner, one register further up: Z[1024 - 64 u] =
  // its e[16 - u] (row 8 - u), and Z[0] mirrors itself.
  const StreamSync ssync{1 + stream};
  fft1024_levels_4_16(z, tw, t);
  ssync();
  double pv[16], p_mid = 0.;
  auto bin_of = [&](int u) { return u < 8 ? t + 64 * u : 1024 - (t + 64 * (u - 8)); };
  {
    const double lf = T->level_factor_fft;
    double2 e[16];
    fft1024_load_64_256(e, z, t);
    ssync();   // every thread holds its sixteen inputs: the buffer becomes the exchange array
    double2* xch = z;
    fft1024_levels_64_256(e, xch, tw, t);
    if (t == 0) {
      const double2 p = xch[fft1024_xch_mid()];   // Z[512]: E = (Re, 0), O = (Im, 0), w^512 = -i
      const double wr = T->tw2048[512].x, wi = T->tw2048[512].y;
      const double xr = p.x + p.y * wr, xi = p.y * wi;
      p_mid = (xr * xr + xi * xi) * lf;
    }
    ssync();
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int k = t + 64 * u;
      const double2 p = e[u];
      const double2 q = fft1024_mirror(xch, t, u, p);
      const double er = 0.5 * (p.x + q.x), ei = 0.5 * (p.y - q.y);
      const double orr = 0.5 * (p.y + q.y), oi = -0.5 * (p.x - q.x);
      const double wr = T->tw2048[k].x, wi = T->tw2048[k].y;
      const double pr = orr * wr - oi * wi, pi = orr * wi + oi * wr;
      const double xr = er + pr, xi = ei + pi;
      const double yr = er - pr, yi = ei - pi;
      pv[u] = (xr * xr + xi * xi) * lf;   // fftearmodel.c:464-466
      pv[8 + u] = (yr * yr + yi * yi) * lf;
    }
  }

  // ---- bandwidth on the register-held spectrum (movs.c:783-803) ----------------------
  if (side == 1) {
    // bins 921..1023: register 8 (bin 1024 - t, t >= 1) and register 9 (bin 960 - t, t <= 39)
    double thr = t >= 1 ? pv[8] : pv[9];
    if (t >= 1 && t <= 39 && pv[9] >= thr) thr = pv[9];
    thr = warp_max_nonan(thr);
    if (lane == 0) mail->thr_part[chan][half] = thr;
  }
  chan_sync(chan);
  double zero_thr;
  {
    const double a = mail->thr_part[chan][0], b = mail->thr_part[chan][1];
    zero_thr = b > a ? b : a;
  }
  if (side == 0) {
    int bw_ref = 0;
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const int k = bin_of(u);
      if (k < 921 && pv[u] > 10. * zero_thr) bw_ref = max(bw_ref, k + 1);
    }
    if (t == 0 && p_mid > 10. * zero_thr) bw_ref = max(bw_ref, 513);
    bw_ref = warp_max_int(bw_ref);
    if (lane == 0) mail->bw_ref_part[chan][half] = bw_ref;
  }
  chan_sync(chan);
  const int bw_ref = max(mail->bw_ref_part[chan][0], mail->bw_ref_part[chan][1]);
  if (side == 1) {
    int bw_test = 0;
    if (bw_ref > 346) {
#pragma unroll
      for (int u = 0; u < 16; u++) {
        const int k = bin_of(u);
        if (k < bw_ref && pv[u] >= 3.16227766016838 * zero_thr) bw_test = max(bw_test, k + 1);
      }
      if (t == 0 && 512 < bw_ref && p_mid >= 3.16227766016838 * zero_thr) bw_test = max(bw_test, 513);
      bw_test = warp_max_int(bw_test);
    }
    if (lane == 0) mail->bw_test_part[chan][half] = bw_test;
  }

  // ---- weighted power spectrum back into the (now free) FFT buffer ----------------
  double* spec = work;
#pragma unroll
  for (int u = 0; u < 16; u++) {
    const int k = bin_of(u);
    if (k < kSpecBins) spec[k] = pv[u] * T->earw2[k];   // fftearmodel.c:470-472
  }
  if (t == 0) spec[512] = p_mid * T->earw2[512];
  chan_sync(chan);   // both spectra of the channel (and bw_test_part) visible

  double* buf_ref = frame_stream_buf(smem, 2 * chan);
  double* buf_test = frame_stream_buf(smem, 2 * chan + 1);
  const double* spec_ref = buf_ref;
  const double* spec_test = buf_test;
  double* dlog = buf_test + kScratchDlog;

  bool ehs_valid = false;
  for (int w = 0; w < 2 * C; w++)
    ehs_valid |= mail->energy[w][0] + mail->energy[w][1] >= 8000. / (32768. * 32768.);

  // Tasks of the channel's four warps: 0, 1: spreading of ref, test; 2, 3: ratio / noise / EHS.
  // A task is tied to the warp index and with it to one scheduler (SM sub-partition) of the SM:
  // the six warps a scheduler holds (two per resident CTA) then run the SAME code, which is what
  // keeps this 140 KB kernel inside the instruction caches (L0 ~6 KB per scheduler, L1.5 32 KB per
  // SM).  Rotating the tasks with the frame index would spread the FP64 load over the four pipes
  // but measured 7 % slower: instruction-fetch stalls 8.6 % -> 24.9 % of all warp stall samples.
#if defined(PEAQ_DEV_ROTATION)
  const int role = ((warp & 3) + (int)(frame & 3u)) & 3;
#else
  const int role = warp & 3;
#endif
  const int tk = (role & 1) * 32 + lane;                    // thread within the two-warp task 2 + 3
  const StreamSync pair_sync{2 + 2 * chan};                 // the test stream's barrier, free since the FFT
  if (role <= 1) {
    // ---- grouping + internal noise + frequency spreading of one stream ----------------
    const bool skip = role == 1 && advanced;   // the advanced FFT model only needs the ref excitation
    double* scratch = role == 0 ? buf_ref + kScratchRef : buf_test + kScratchTest;
    double* se = scratch + 128;
    if (!skip) {
      const double* sp = role == 0 ? spec_ref : spec_test;
      for (int i = lane; i < B; i += 32) se[i] = group_band(T, sp, i) + T->fft.internal_noise[i];
    }
    __threadfence_block();
    refbuf_arrive(chan);   // this warp no longer reads the ref spectrum
    if (!skip) {
      __syncwarp();
      spread_prepare(T, B, scratch, se, scratch + 256, lane);
      spread_ladder(B, scratch, lane);
      // E2 = E2s^(1/0.4) / norm  (:673-675); x^2.5 = x^2 sqrt(x)
      double* out = kFused ? const_cast<double*>(frame_out_e2(smem, chan, role)) : rec + (role * C + chan) * B;
      for (int i = lane; i < B; i += 32) {
        const double v = spread_result(scratch, i);
        out[i] = v * v * sqrt(v) / T->spread_norm[i];
      }
    }
    if (role == 0 && lane == 0) {
      const int bw_test = max(mail->bw_test_part[chan][0], mail->bw_test_part[chan][1]);
      if (kFused) {
        mail->o_bw[chan][0] = bw_ref;
        mail->o_bw[chan][1] = bw_test;
      } else {
        int* ints = reinterpret_cast<int*>(rec + L.off_ints);
        ints[1 + 2 * chan] = bw_ref;
        ints[2 + 2 * chan] = bw_test;
      }
    }
  } else {
    // ---- ln spectrum ratio (movs.c:1396-1403) and noise spectrum (movs.c:993-998), one bin per
    // thread; then the noise in bands (movs.c:999-1000) ---------------------------------------
    double* nzs = buf_ref + (kSpecBins - 2);   // nzs[k] = noise bin k, k = 2..768: ref buffer [769, 1536)
#pragma unroll 1
    for (int u = 0; u < 8; u++) {
      const int i = tk + 64 * u;
      const double fref = spec_ref[i], ftest = spec_test[i];
      dlog[i] = (fref == 0. && ftest == 0.) ? 0. : peaq_log(ftest / fref);
    }
    // Noise in bands, two ways (same arithmetic, same bits).  Basic mode: the spreading warps are
    // the critical path of the frame and share the FP64 pipes and the shared-memory pipe with
    // these two, so the noise spectrum is computed where it is consumed, a lane per band, spread
    // out in time (a burst of 767 square roots at the start of the tail slowed the spreading
    // warps: 151.8 -> 169.7 ms per 4096 x 10 s).  Advanced mode (55 bands, only the ref stream is
    // spread): these two warps ARE the critical path, and one bin per thread up front is faster
    // (157.8 -> 128.7 ms).
    if (!advanced) {
      double* out = kFused ? const_cast<double*>(frame_out_noise(smem, chan)) : rec + L.off_noise + chan * B;
      for (int i = tk; i < B; i += 64) {
        const int lo = T->band_lo[i], hi = T->band_hi[i];
        const double wl = T->band_wl[i], wu = T->band_wu[i];
        double p = 0.;
        const int n_terms = hi > lo ? hi - lo + 1 : 2;
#pragma unroll 1
        for (int j = 0; j < n_terms; j++) {
          const int k = j == 0 ? lo : (j == 1 ? hi : lo + j - 1);
          const double w = j == 0 ? wl : (j == 1 ? wu : 1.);
          const double v = w * noise_bin(spec_ref, spec_test, k);
          p = j == 0 ? v : p + v;
        }
        out[i] = p < 1e-12 ? 1e-12 : p;
      }
    } else {
#pragma unroll 4
      for (int k = 2 + tk; k < kSpecBins; k += 64) nzs[k] = noise_bin(spec_ref, spec_test, k);
      __threadfence_block();
      pair_sync();   // both warps of the task: noise spectrum and ln ratio complete
      double* out = kFused ? const_cast<double*>(frame_out_noise(smem, chan)) : rec + L.off_noise + chan * B;
      for (int i = tk; i < B; i += 64) out[i] = group_band_noise(T, nzs, i);
    }
    __threadfence_block();
    // both warps: ref spectrum and noise spectrum dead -- the ref buffer carries the EHS transforms
    refbuf_wait(chan);
    double ehs = 0.;
    if (ehs_valid) ehs = ehs_channel_pair(T, dlog, buf_ref, tw, tk, pair_sync, mail->ehs_x[chan]);
    if (role == 2 && lane == 0) {
      if (kFused) mail->o_ehs[chan] = ehs;
      else rec[L.off_ehs + chan] = ehs;
    }
  }

  if (threadIdx.x == 0) {
    // is_frame_above_threshold (gstpeaq.c:1081-1099) from the largest |x| per channel: a sample
    // of 1.01 thr alone carries its window over the threshold, five of less than 0.99 thr / 5
    // cannot reach it (the float running sum drifts by < 4e-6 while it stays below); only in
    // between is the recurrence replayed literally
    bool above = false;
    double sum_s = 0., sum_n = 0.;
    const float thr_f = 200.f / 32768.f;
    for (int c = 0; c < C; c++) {
      const float m = fmaxf(mail->maxabs[c][0], mail->maxabs[c][1]);
      if (m >= thr_f * 1.01f) above = true;
      sum_s += mail->snr_s[c][0] + mail->snr_s[c][1];
      sum_n += mail->snr_n[c][0] + mail->snr_n[c][1];
    }
    if (!above) {
      for (int c = 0; c < C && !above; c++) {
        const float m = fmaxf(mail->maxabs[c][0], mail->maxabs[c][1]);
        if (5.f * m >= thr_f * 0.99f) above = replay_threshold_serial(ref_sig, s0, n_ref, c, C);
      }
    }
    const int flags = (above ? kRecFlagAbove : 0) | (ehs_valid ? kRecFlagEhsValid : 0);
    if (kFused) {
      mail->o_snr[0] = sum_s;
      mail->o_snr[1] = sum_n;
      mail->o_flags = flags;
    } else {
      rec[L.off_snr] = sum_s;
      rec[L.off_snr + 1] = sum_n;
      reinterpret_cast<int*>(rec + L.off_ints)[0] = flags;
    }
  }
}

constexpr size_t frame_smem_bytes(int channels) {
  return sizeof(double) * (kTwDoubles + 2 * channels * kWorkDoubles) + sizeof(FrameMail);
}

}  // namespace
}  // namespace peaq
