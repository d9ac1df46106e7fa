// Warp-level double-precision FFT in shared memory (sm_100a).
//
// One warp transforms N = 2^LOG2N complex points held in shared memory:
// in-place radix-4 decimation in time (plus one radix-2 pass when LOG2N is
// odd).  The caller stores input sample n at slot fft_slot<LOG2N>(n) (base-4
// digit reversal, XOR-swizzled); the result is in natural order behind the
// same swizzle: X[k] = z[fft_swz(k)].
//
// The swizzle folds index bits 3,4,6,8,9 into the low three bits so that every
// 128-bit shared-memory access of every pass, the digit-reversed input scatter
// and the natural-order read-out are bank-conflict free for N = 1024.  It is
// linear over GF(2) (swz(a ^ b) = swz(a) ^ swz(b)) and every index used by a
// pass is lane_part ^ compile_time_part with disjoint bits, so a pass computes
// ONE swizzle per lane; all other addresses are XORs with constants.
//
// Twiddles: exp(-2 pi i q / 1024) for q < 512 in shared memory; the second half
// of the circle is the negated first half.
//
// Replaces the reference's calls into GstFFTF64/kissfft
// (/root/reference/src/fftearmodel.c:457, movs.c:1301-1313,1428).
#pragma once

#include <cuda_runtime.h>

namespace peaq {

__host__ __device__ constexpr int fft_swz(int i) {
  return i ^ (((i >> 3) ^ (i >> 6)) & 1) ^ ((((i >> 4) ^ (i >> 8)) & 1) << 1) ^
         ((((i >> 4) ^ (i >> 9)) & 1) << 2);
}

// base-4 digit reversal of a (2*DIGITS)-bit index (compile-time friendly)
__host__ __device__ constexpr int fft_rev4c(int n, int digits) {
  int r = 0;
  for (int d = 0; d < digits; d++) {
    r = (r << 2) | (n & 3);
    n >>= 2;
  }
  return r;
}

// slot (before swizzle) that input sample n must be written to; a bit
// permutation of n, hence linear: perm(a ^ b) = perm(a) ^ perm(b)
template <int LOG2N>
__host__ __device__ constexpr int fft_perm(int n) {
  return (LOG2N % 2 == 0) ? fft_rev4c(n, LOG2N / 2)
                          : (n & 1) * (1 << (LOG2N - 1)) + fft_rev4c(n >> 1, LOG2N / 2);
}

template <int LOG2N>
__host__ __device__ constexpr int fft_slot(int n) {
  return fft_swz(fft_perm<LOG2N>(n));
}

// run-time versions for lane-dependent indices (bit reversal in one instruction)
template <int LOG2N>
__device__ __forceinline__ int fft_perm_rt(int n) {
  constexpr int E = LOG2N & ~1;   // bits that are digit-reversed
  const int m = (LOG2N % 2 == 0) ? n : (n >> 1);
  unsigned r = __brev((unsigned)m) >> (32 - E);
  r = ((r & 0x55555555u) << 1) | ((r >> 1) & 0x55555555u);
  return (LOG2N % 2 == 0) ? (int)r : (n & 1) * (1 << (LOG2N - 1)) + (int)r;
}

template <int LOG2N>
__device__ __forceinline__ int fft_slot_rt(int n) {
  return fft_swz(fft_perm_rt<LOG2N>(n));
}

__device__ __forceinline__ double2 cmul(double2 a, double2 w) {
  // explicit fused multiply-adds: the FFT's rounding need not (and cannot)
  // match kissfft's, only its accuracy
  return make_double2(fma(a.x, w.x, -a.y * w.y), fma(a.x, w.y, a.y * w.x));
}

// twiddle exp(-2 pi i q / 1024), q < 1024, from the half-circle table
__device__ __forceinline__ double2 fft_tw(const double2* __restrict__ tw, int q) {
  const double2 w = tw[q & 511];
  return (q & 512) ? make_double2(-w.x, -w.y) : w;
}

// index contributed by iteration u of a pass (added to the lane part, disjoint bits)
template <int LQ>
__host__ __device__ constexpr int fft_pass_delta(int u) {
  return LQ <= 32 ? 128 * u : (LQ == 64 ? (32 * (u & 1)) | (256 * (u >> 1)) : 32 * u);
}

// one radix-4 pass; LQ = size of the sub-transforms being combined
template <int LOG2N, int LQ>
__device__ __forceinline__ void fft_pass4(double2* z, const double2* __restrict__ tw, int lane) {
  constexpr int N = 1 << LOG2N;
  constexpr int TS = 1024 / (4 * LQ);  // stride into the 1024-point twiddle circle
  // butterfly b = lane + 32 u combines i0 + {0,1,2,3} LQ with i0 = 4 (b - k) + k, k = b mod LQ
  const int k_lane = LQ <= 32 ? (lane & (LQ - 1)) : lane;
  const int base = LQ <= 32 ? ((lane - k_lane) << 2) + k_lane : lane;
  const int sb = fft_swz(base);
  double2 w1, w2, w3;
  if (LQ > 1 && LQ <= 32) {   // twiddles depend on the lane only
    w1 = tw[k_lane * TS];
    w2 = tw[2 * k_lane * TS];
    w3 = fft_tw(tw, 3 * k_lane * TS);
  }
#pragma unroll
  for (int u = 0; u < N / 128; u++) {
    constexpr int dummy = 0;
    (void)dummy;
    const int d = fft_pass_delta<LQ>(u);
    const int s0 = sb ^ fft_swz(d), s1 = sb ^ fft_swz(d | LQ), s2 = sb ^ fft_swz(d | (2 * LQ)),
              s3 = sb ^ fft_swz(d | (3 * LQ));
    double2 a0 = z[s0], a1 = z[s1], a2 = z[s2], a3 = z[s3];
    if (LQ > 32) {
      const int k = LQ == 64 ? (lane | (32 * (u & 1))) : (lane | (32 * u));
      w1 = tw[k * TS];
      w2 = tw[2 * k * TS];
      w3 = fft_tw(tw, 3 * k * TS);
    }
    if (LQ > 1) {
      a1 = cmul(a1, w1);
      a2 = cmul(a2, w2);
      a3 = cmul(a3, w3);
    }
    const double2 t0 = make_double2(a0.x + a2.x, a0.y + a2.y);
    const double2 t1 = make_double2(a0.x - a2.x, a0.y - a2.y);
    const double2 t2 = make_double2(a1.x + a3.x, a1.y + a3.y);
    const double2 t3 = make_double2(a1.y - a3.y, a3.x - a1.x);   // -i (a1 - a3)
    z[s0] = make_double2(t0.x + t2.x, t0.y + t2.y);
    z[s1] = make_double2(t1.x + t3.x, t1.y + t3.y);
    z[s2] = make_double2(t0.x - t2.x, t0.y - t2.y);
    z[s3] = make_double2(t1.x - t3.x, t1.y - t3.y);
  }
  __syncwarp();
}

template <int LOG2N, int LQ>
struct FftPasses {
  static __device__ __forceinline__ void run(double2* z, const double2* __restrict__ tw, int lane) {
    fft_pass4<LOG2N, LQ>(z, tw, lane);
    FftPasses<LOG2N, LQ * 4>::run(z, tw, lane);
  }
};
// recursion ends once LQ reaches 4^(LOG2N/2)
template <> struct FftPasses<10, 1024> { static __device__ __forceinline__ void run(double2*, const double2*, int) {} };
template <> struct FftPasses<9, 256> { static __device__ __forceinline__ void run(double2*, const double2*, int) {} };
template <> struct FftPasses<8, 256> { static __device__ __forceinline__ void run(double2*, const double2*, int) {} };

// Forward transform (kernel exp(-2 pi i k n / N)).  `tw` = exp(-2 pi i q/1024),
// q < 512, in shared memory.  Caller must __syncwarp() after filling z.
template <int LOG2N>
__device__ __forceinline__ void warp_fft(double2* z, const double2* __restrict__ tw, int lane) {
  constexpr int N = 1 << LOG2N;
  FftPasses<LOG2N, 1>::run(z, tw, lane);
  if (LOG2N % 2 == 1) {
    constexpr int H = N / 2;
    constexpr int TS = 1024 / N;
    const int sl = fft_swz(lane);
#pragma unroll
    for (int u = 0; u < H / 32; u++) {
      const int se = sl ^ fft_swz(32 * u), so = sl ^ fft_swz(32 * u + H);
      const double2 e = z[se];
      const double2 o = cmul(z[so], tw[(lane + 32 * u) * TS]);
      z[se] = make_double2(e.x + o.x, e.y + o.y);
      z[so] = make_double2(e.x - o.x, e.y - o.y);
    }
    __syncwarp();
  }
}

}  // namespace peaq
