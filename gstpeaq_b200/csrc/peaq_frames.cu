// K1 `fft_frames`: the stateless, frame-parallel half of the PEAQ hot path.
//
// One CTA = one FFT-clock frame of one (ref,test) pair; 2*C warps, warp w handles
// stream (channel c = w/2, side = w%2: 0 ref, 1 test).  Per stream:
//   PCM (interleaved F32, HBM) -> Hann window -> 2048-pt real FFT (1024-pt complex
//   radix-4 in shared memory + split) -> power spectrum -> outer/middle-ear
//   weighting -> critical-band grouping -> + internal noise -> level dependent
//   frequency spreading           (fftearmodel.c:432-515, :603-676)
// and per channel: noise spectrum grouped into bands (movs.c:988-1000), bandwidth
// bins (movs.c:776-809), error-harmonic-structure value (movs.c:1346-1443), energy
// and above-threshold flags (fftearmodel.c:508-514, gstpeaq.c:1081-1099) and the
// SNR partial sums (gstpeaq.c:913-918).  Everything recurrent is left to K2.
//
// Arithmetic is IEEE double; the file is compiled with -fmad=false so that
// expressions restated from the reference round like the reference's (gcc,
// x86-64, no contraction); fused multiply-adds appear only where written
// explicitly (FFT butterflies).
#include "peaq_engine.h"
#include "peaq_fft.cuh"

namespace peaq {
namespace {

constexpr int kWorkDoubles = 2048;   // per-warp scratch: 1024 complex points
constexpr int kSpecDoubles = 1032;   // per-warp power spectrum (1025, padded)
constexpr int kTwDoubles = 2 * 768;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_max_nonan(double v) {
  // v must not be NaN
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

// sample `i` of channel `c` of an interleaved signal, zero past the end
__device__ __forceinline__ float pcm_at(const float* __restrict__ sig, unsigned long long s0,
                                        unsigned long long n_samples, int i, int c, int C) {
  const unsigned long long s = s0 + (unsigned long long)i;
  return s < n_samples ? __ldg(sig + s * (unsigned long long)C + c) : 0.0f;
}

// is_frame_above_threshold for one channel (gstpeaq.c:1081-1099), exact.
//
// The reference keeps the 5-sample running sum of |x| in a FLOAT that is
// updated sequentially with double increments; the test is made from sample 5
// on.  The float recurrence drifts from the exact window sum W_i by less than
// 2048 * 2^-24 relative, so W_i decides the outcome unless some W_i lies within
// 1 % of the threshold; only then is the sequential recurrence replayed.
__device__ bool channel_above_threshold(const float* __restrict__ sig, unsigned long long s0,
                                        unsigned long long n_samples, int c, int C, int lane) {
  const double thr = 200. / 32768;
  double prev = 0.;
  bool sure = false, maybe = false;
  for (int j = 0; j < kFftFrame / 32; j++) {
    const int i = lane + 32 * j;
    const double v = fabs((double)pcm_at(sig, s0, n_samples, i, c, C));
    double w = v;
#pragma unroll
    for (int k = 1; k <= 4; k++) {
      const double a = __shfl_sync(0xffffffffu, v, (lane - k) & 31);
      const double b = __shfl_sync(0xffffffffu, prev, (lane - k) & 31);
      w += (lane >= k) ? a : b;
    }
    if (i >= 5) {
      if (w >= thr * 1.01) sure = true;
      else if (w >= thr * 0.99) maybe = true;
    }
    prev = v;
  }
  if (__any_sync(0xffffffffu, sure)) return true;
  if (!__any_sync(0xffffffffu, maybe)) return false;
  // borderline: replay the reference's recurrence literally on one lane
  int result = 0;
  if (lane == 0) {
    float sum = 0;
    int i;
    for (i = 0; i < 5; i++)
      sum = (float)((double)sum + fabs((double)pcm_at(sig, s0, n_samples, i, c, C)));
    while (i < kFftFrame) {
      sum = (float)((double)sum + (fabs((double)pcm_at(sig, s0, n_samples, i, c, C)) -
                                   fabs((double)pcm_at(sig, s0, n_samples, i - 5, c, C))));
      if ((double)sum >= thr) {
        result = 1;
        break;
      }
      i++;
    }
  }
  return __shfl_sync(0xffffffffu, result, 0) != 0;
}

// weighted power spectrum value (fftearmodel.c:470-472)
__device__ __forceinline__ double weighted(const double* spec, const double* __restrict__ earw2,
                                           int k) {
  return spec[k] * earw2[k];
}

// peaq_fftearmodel_group_into_bands for band i (fftearmodel.c:603-620) over the
// weighted power spectrum of one stream
__device__ __forceinline__ double group_band_weighted(const DeviceTables* __restrict__ T,
                                                      const double* spec, int i) {
  const int lo = T->band_lo[i], hi = T->band_hi[i];
  double p = T->band_wl[i] * weighted(spec, T->earw2, lo) + T->band_wu[i] * weighted(spec, T->earw2, hi);
  for (int k = lo + 1; k < hi; k++) p += weighted(spec, T->earw2, k);
  return p < 1e-12 ? 1e-12 : p;
}

// noise spectrum bin (movs.c:993-998)
__device__ __forceinline__ double noise_bin(const double* spec_ref, const double* spec_test,
                                            const double* __restrict__ earw2, int k) {
  const double r = weighted(spec_ref, earw2, k), t = weighted(spec_test, earw2, k);
  return r - 2 * sqrt(r * t) + t;
}

__device__ __forceinline__ double group_band_noise(const DeviceTables* __restrict__ T,
                                                   const double* spec_ref, const double* spec_test,
                                                   int i) {
  const int lo = T->band_lo[i], hi = T->band_hi[i];
  double p = T->band_wl[i] * noise_bin(spec_ref, spec_test, T->earw2, lo) +
             T->band_wu[i] * noise_bin(spec_ref, spec_test, T->earw2, hi);
  for (int k = lo + 1; k < hi; k++) p += noise_bin(spec_ref, spec_test, T->earw2, k);
  return p < 1e-12 ? 1e-12 : p;
}

// Level dependent frequency spreading of one stream (do_spreading,
// fftearmodel.c:636-676).  `pp` holds the pitch pattern on entry (scratch),
// result goes to `out` (global record).  sa/se/se2 are warp-private shared
// arrays of >= B doubles.
__device__ void spread_bands(const DeviceTables* __restrict__ T, int B, double* sa, double* se,
                             double* se2, double* __restrict__ out, int lane) {
  const double dz02 = 0.2 * T->dz;
  for (int i = lane; i < B; i += 32) {
    const double pp = se[i];
    const double a_uce = T->aUC[i] * pow(pp, dz02);
    const double g_iu = (1. - pow(a_uce, (double)(B - i))) / (1. - a_uce);
    const double en = pp / (T->gIL[i] + g_iu - 1.);
    sa[i] = pow(a_uce, 0.4);
    se[i] = pow(en, 0.4);
  }
  __syncwarp();
  // downward spreading, constant slope: E2[i-1] = aLe * E2[i] + Ene[i-1]
  if (lane == 0) {
    const double a_le = T->aLe;
    double acc = se[B - 1];
    se2[B - 1] = acc;
    for (int i = B - 1; i > 0; i--) {
      acc = a_le * acc + se[i - 1];
      se2[i - 1] = acc;
    }
  }
  __syncwarp();
  // upward spreading: source band i adds Ene[i] * aUCEe[i]^(j-i) to every j > i.
  // Each lane walks its (up to 4) source bands upwards in lock step; at step t
  // all written targets i+t are distinct, so plain read-modify-write is safe.
  double r[4], a[4];
#pragma unroll
  for (int m = 0; m < 4; m++) {
    const int i = lane + 32 * m;
    r[m] = i < B ? se[i] : 0.;
    a[m] = i < B ? sa[i] : 0.;
  }
  for (int t = 1; t < B; t++) {
#pragma unroll
    for (int m = 0; m < 4; m++) {
      const int j = lane + 32 * m + t;
      if (j < B) {
        r[m] *= a[m];
        se2[j] += r[m];
      }
    }
    __syncwarp();
  }
  for (int i = lane; i < B; i += 32) out[i] = pow(se2[i], 2.5) / T->spread_norm[i];
}

// Error harmonic structure of one channel (peaq_mov_ehs, movs.c:1383-1441),
// run by one warp.  d[0..511] = ln(Pw_test/Pw_ref) is already in `dlog`
// (shared).  work: 2048 doubles of warp-private shared memory.
__device__ double ehs_channel(const DeviceTables* __restrict__ T, const double* dlog, double* work,
                              const double2* __restrict__ tw, int lane) {
  double2* za = reinterpret_cast<double2*>(work);         // 512 complex
  double2* zb = reinterpret_cast<double2*>(work) + 512;   // 512 complex
  // both forward transforms of do_xcorr (movs.c:1300-1303) in one complex FFT:
  // real part = d[0..511], imaginary part = d[0..255] followed by zeros
  for (int n = lane; n < 512; n += 32) {
    const double v = dlog[n];
    za[fft_slot<9>(n)] = make_double2(v, n < kMaxLag ? v : 0.);
  }
  __syncwarp();
  warp_fft<9>(za, tw, lane);
  // F1 = (Z[k] + conj Z[N-k]) / 2, F2 = (Z[k] - conj Z[N-k]) / 2i,
  // G = F1 * conj(F2) / 512 (movs.c:1304-1312); store conj(G) for the inverse
  for (int k = lane; k < 512; k += 32) {
    const double2 p = za[fft_swz(k)];
    const double2 q = za[fft_swz((512 - k) & 511)];
    const double f1r = 0.5 * (p.x + q.x), f1i = 0.5 * (p.y - q.y);
    const double f2r = 0.5 * (p.y + q.y), f2i = -0.5 * (p.x - q.x);
    const double gr = (f1r * f2r + f1i * f2i) / (2 * kMaxLag);
    const double gi = (f2r * f1i - f1r * f2i) / (2 * kMaxLag);
    zb[fft_slot<9>(k)] = make_double2(gr, -gi);
  }
  __syncwarp();
  warp_fft<9>(zb, tw, lane);   // c[l] = Re zb[l]  (unnormalised inverse, like GstFFT)
  // normalisation by the running window energy (movs.c:1405-1418):
  //   c[i] /= sqrt(d0 * dk_i),  dk_i = d0 + sum_{j<i} (d[j+256]^2 - d[j]^2)
  const double d0 = zb[fft_swz(0)].x;
  // lane owns i = 8*lane .. 8*lane+7 for the prefix sum
  double term[8], c[8];
  double local = 0.;
#pragma unroll
  for (int e = 0; e < 8; e++) {
    const int i = 8 * lane + e;
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
  double excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 0.;
  double dk = d0 + excl;
  double csum = 0.;
#pragma unroll
  for (int e = 0; e < 8; e++) {
    const int i = 8 * lane + e;
    c[e] = zb[fft_swz(i)].x / sqrt(d0 * dk);
    csum += c[e];
    dk += term[e];
  }
  const double cavg = warp_sum(csum) / kMaxLag;
  __syncwarp();
  // subtract mean, window (movs.c:1419-1421), 256-pt real FFT as a complex one
  for (int e = 0; e < 8; e++) {
    const int i = 8 * lane + e;
    za[fft_slot<8>(i)] = make_double2((c[e] - cavg) * T->ehs_window[i], 0.);
  }
  __syncwarp();
  warp_fft<8>(za, tw, lane);
  // largest |C[k]|^2 among 1 <= k <= 128 that rises above its left neighbour
  // (movs.c:1433-1440; NaNs never win a comparison, exactly as there)
  double best = 0.;
  for (int k = 1 + lane; k <= kMaxLag / 2; k += 32) {
    const double2 x = za[fft_swz(k)], y = za[fft_swz(k - 1)];
    const double s = x.x * x.x + x.y * x.y;
    const double sp = y.x * y.x + y.y * y.y;
    if (s > sp && s > best) best = s;
  }
  return warp_max_nonan(best);
}

__global__ void __launch_bounds__(128)
fft_frames_kernel(const DeviceTables* __restrict__ T, PcmView pcm, unsigned first_frame,
                  unsigned n_chunk_frames, double* __restrict__ records, RecordLayout L, int B,
                  int advanced) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = pcm.channels;
  const int pair = blockIdx.x / n_chunk_frames;
  const unsigned chunk_frame = blockIdx.x - pair * n_chunk_frames;
  const unsigned frame = first_frame + chunk_frame;
  if (frame >= pcm.n_frames[pair]) return;

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int chan = warp >> 1;
  const int side = warp & 1;

  double* smem = reinterpret_cast<double*>(smem_raw);
  double2* tw = reinterpret_cast<double2*>(smem);                       // 768 complex
  double* work = smem + kTwDoubles + warp * (kWorkDoubles + kSpecDoubles);
  double* spec = work + kWorkDoubles;                                   // power spectrum
  // small cross-warp mailbox behind the per-warp areas
  double* mail = smem + kTwDoubles + 2 * C * (kWorkDoubles + kSpecDoubles);
  int* mail_flags = reinterpret_cast<int*>(mail + 8);                   // [2C] energy, [C] above

  for (int i = threadIdx.x; i < 768; i += blockDim.x)
    tw[i] = make_double2(T->tw1024[i].x, T->tw1024[i].y);

  const unsigned long long n_ref = pcm.n_samples[pair], n_test = pcm.n_samples_test[pair];
  const unsigned long long n_samples = side ? n_test : n_ref;
  const unsigned long long s0 = (unsigned long long)frame * kFftStep;
  const float* __restrict__ sig = (side ? pcm.test : pcm.ref) + (size_t)pair * pcm.pair_stride;

  // ---- load, window, scatter into FFT order; energy of the second half --------
  double2* z = reinterpret_cast<double2*>(work);
  double energy = 0.;
  for (int u = 0; u < 32; u++) {
    const int n = lane + 32 * u;   // complex index: samples 2n, 2n+1
    const float x0 = pcm_at(sig, s0, n_samples, 2 * n, chan, C);
    const float x1 = pcm_at(sig, s0, n_samples, 2 * n + 1, chan, C);
    z[fft_slot<10>(n)] = make_double2(T->hann[2 * n] * x0, T->hann[2 * n + 1] * x1);
    if (n >= 512) {   // samples 1024..2047: float products, double accumulation
      energy += (double)(x0 * x0);
      energy += (double)(x1 * x1);
    }
  }
  energy = warp_sum(energy);
  if (lane == 0) mail_flags[warp] = energy >= 8000. / (32768. * 32768.);

  if (side == 0) {
    const bool above = channel_above_threshold(sig, s0, n_samples, chan, C, lane);
    if (lane == 0) mail_flags[2 * C + chan] = above;
  } else {
    // SNR partial sums over the first half of the frame (gstpeaq.c:913-918):
    // float products accumulated in double
    const float* __restrict__ ref_sig = pcm.ref + (size_t)pair * pcm.pair_stride;
    double es = 0., en = 0.;
    for (int i = lane; i < kFftFrame / 2; i += 32) {
      const float r = pcm_at(ref_sig, s0, n_ref, i, chan, C);
      const float t = pcm_at(sig, s0, n_samples, i, chan, C);
      es += (double)(r * r);
      en += (double)((r - t) * (r - t));
    }
    es = warp_sum(es);
    en = warp_sum(en);
    if (lane == 0) {
      mail[2 * chan] = es;
      mail[2 * chan + 1] = en;
    }
  }
  __syncthreads();   // twiddles loaded (and this warp's scatter complete)

  // ---- 2048-point real FFT --------------------------------------------------
  warp_fft<10>(z, tw, lane);
  {
    const double lf = T->level_factor_fft;
    for (int k = lane; k <= 1024; k += 32) {
      const double2 p = z[fft_swz(k & 1023)];
      const double2 q = z[fft_swz((1024 - k) & 1023)];
      const double er = 0.5 * (p.x + q.x), ei = 0.5 * (p.y - q.y);
      const double orr = 0.5 * (p.y + q.y), oi = -0.5 * (p.x - q.x);
      const double wr = T->tw2048[k].x, wi = T->tw2048[k].y;
      const double xr = er + (orr * wr - oi * wi);
      const double xi = ei + (orr * wi + oi * wr);
      spec[k] = (xr * xr + xi * xi) * lf;
    }
  }
  __syncthreads();   // all spectra visible; FFT scratch free for reuse

  double* rec = records + ((size_t)pair * n_chunk_frames + chunk_frame) * L.stride;
  const double* spec_ref = smem + kTwDoubles + (2 * chan) * (kWorkDoubles + kSpecDoubles) + kWorkDoubles;
  const double* spec_test = smem + kTwDoubles + (2 * chan + 1) * (kWorkDoubles + kSpecDoubles) + kWorkDoubles;
  double* work_test = smem + kTwDoubles + (2 * chan + 1) * (kWorkDoubles + kSpecDoubles);

  // ---- own stream: grouping + internal noise + frequency spreading -----------
  double* sa = work;              // scratch arrays inside the warp's work area
  double* se = work + 128;
  double* se2 = work + 256;
  double* dlog = work_test + 512;   // 512 doubles, written by the test warp
  if (!(advanced && side == 1)) {
    for (int i = lane; i < B; i += 32)
      se[i] = group_band_weighted(T, spec, i) + T->fft.internal_noise[i];
    __syncwarp();
    spread_bands(T, B, sa, se, se2, rec + (side * C + chan) * B, lane);
  }

  if (side == 1) {
    // ---- noise in bands (movs.c:988-1000) ----------------------------------
    for (int i = lane; i < B; i += 32)
      rec[L.off_noise + chan * B + i] = group_band_noise(T, spec_ref, spec_test, i);
    // ---- bandwidth (movs.c:783-803) -----------------------------------------
    double thr = spec_test[921];
    for (int k = 922 + lane; k < 1024; k += 32) {
      const double v = spec_test[k];
      if (v >= thr) thr = v;
    }
    thr = warp_max_nonan(thr);
    int bw_ref = 0;
    for (int k = lane; k < 921; k += 32)
      if (spec_ref[k] > 10. * thr) bw_ref = k + 1;
    bw_ref = warp_max_int(bw_ref);
    int bw_test = 0;
    if (bw_ref > 346) {
      for (int k = lane; k < bw_ref; k += 32)
        if (spec_test[k] >= 3.16227766016838 * thr) bw_test = k + 1;
      bw_test = warp_max_int(bw_test);
    }
    if (lane == 0) {
      int* ints = reinterpret_cast<int*>(rec + L.off_ints);
      ints[1 + 2 * chan] = bw_ref;
      ints[2 + 2 * chan] = bw_test;
    }
    // ---- log spectrum ratio for the EHS (movs.c:1396-1403) -------------------
    for (int i = lane; i < 2 * kMaxLag; i += 32) {
      const double fref = weighted(spec_ref, T->earw2, i);
      const double ftest = weighted(spec_test, T->earw2, i);
      dlog[i] = (fref == 0. && ftest == 0.) ? 0. : log(ftest / fref);
    }
  }
  __syncthreads();   // dlog ready; flags in the mailbox

  bool ehs_valid = false;
  for (int w = 0; w < 2 * C; w++) ehs_valid |= mail_flags[w] != 0;
  if (side == 0) {
    double ehs = 0.;
    if (ehs_valid) ehs = ehs_channel(T, dlog, work, tw, lane);
    if (lane == 0) rec[L.off_ehs + chan] = ehs;
  }
  if (threadIdx.x == 0) {
    bool above = false;
    double es = 0., en = 0.;
    for (int c = 0; c < C; c++) {
      above |= mail_flags[2 * C + c] != 0;
      es += mail[2 * c];
      en += mail[2 * c + 1];
    }
    rec[L.off_snr] = es;
    rec[L.off_snr + 1] = en;
    int* ints = reinterpret_cast<int*>(rec + L.off_ints);
    ints[0] = (above ? kRecFlagAbove : 0) | (ehs_valid ? kRecFlagEhsValid : 0);
  }
}

}  // namespace

size_t fft_frames_smem_bytes(int channels) {
  return sizeof(double) * (kTwDoubles + 2 * channels * (kWorkDoubles + kSpecDoubles) + 16);
}

cudaError_t launch_fft_frames(const DeviceTables* d_tables, PcmView pcm, int n_pairs,
                              unsigned first_frame, unsigned n_chunk_frames, double* records,
                              RecordLayout L, int fft_bands, bool advanced, cudaStream_t stream) {
  if (n_pairs <= 0 || n_chunk_frames == 0) return cudaSuccess;
  const size_t smem = fft_frames_smem_bytes(pcm.channels);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(fft_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)fft_frames_smem_bytes(kMaxChannels));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((unsigned)n_chunk_frames * (unsigned)n_pairs);
  dim3 block(64 * pcm.channels);
  fft_frames_kernel<<<grid, block, smem, stream>>>(d_tables, pcm, first_frame, n_chunk_frames, records,
                                                   L, fft_bands, advanced ? 1 : 0);
  return cudaGetLastError();
}

}  // namespace peaq
