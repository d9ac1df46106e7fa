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
// and the natural-order read-out are bank-conflict free for N = 1024
// (checked by enumeration).
//
// Replaces the reference's calls into GstFFTF64/kissfft
// (/root/reference/src/fftearmodel.c:457, movs.c:1301-1313,1428).
#pragma once

#include <cuda_runtime.h>

namespace peaq {

__device__ __forceinline__ int fft_swz(int i) {
  return i ^ (((i >> 3) ^ (i >> 6)) & 1) ^ ((((i >> 4) ^ (i >> 8)) & 1) << 1) ^
         ((((i >> 4) ^ (i >> 9)) & 1) << 2);
}

// base-4 digit reversal of a (2*DIGITS)-bit index
template <int DIGITS>
__device__ __forceinline__ int fft_rev4(int n) {
  unsigned r = __brev((unsigned)n) >> (32 - 2 * DIGITS);  // bit reversal ...
  r = ((r & 0x55555555u) << 1) | ((r >> 1) & 0x55555555u);  // ... un-swap inside each digit
  return (int)r;
}

// slot (before swizzle) that input sample n must be written to
template <int LOG2N>
__device__ __forceinline__ int fft_perm(int n) {
  if (LOG2N % 2 == 0) return fft_rev4<LOG2N / 2>(n);
  return (n & 1) * (1 << (LOG2N - 1)) + fft_rev4<LOG2N / 2>(n >> 1);
}

template <int LOG2N>
__device__ __forceinline__ int fft_slot(int n) {
  return fft_swz(fft_perm<LOG2N>(n));
}

__device__ __forceinline__ double2 cmul(double2 a, double2 w) {
  // explicit fused multiply-adds: the FFT's rounding need not (and cannot)
  // match kissfft's, only its accuracy
  return make_double2(fma(a.x, w.x, -a.y * w.y), fma(a.x, w.y, a.y * w.x));
}

// one radix-4 pass; LQ = size of the sub-transforms being combined
template <int LOG2N, int LQ>
__device__ __forceinline__ void fft_pass4(double2* z, const double2* __restrict__ tw, int lane) {
  constexpr int N = 1 << LOG2N;
  constexpr int L = 4 * LQ;
  constexpr int TS = 1024 / L;  // stride into the 1024-point twiddle table
#pragma unroll 2
  for (int u = 0; u < N / 128; u++) {
    const int b = lane + 32 * u;
    const int k = b & (LQ - 1);
    const int i0 = ((b - k) << 2) + k;
    const int s0 = fft_swz(i0), s1 = fft_swz(i0 + LQ), s2 = fft_swz(i0 + 2 * LQ),
              s3 = fft_swz(i0 + 3 * LQ);
    double2 a0 = z[s0], a1 = z[s1], a2 = z[s2], a3 = z[s3];
    if (LQ > 1) {
      a1 = cmul(a1, tw[k * TS]);
      a2 = cmul(a2, tw[2 * k * TS]);
      a3 = cmul(a3, tw[3 * k * TS]);
    }
    const double2 t0 = make_double2(a0.x + a2.x, a0.y + a2.y);
    const double2 t1 = make_double2(a0.x - a2.x, a0.y - a2.y);
    const double2 t2 = make_double2(a1.x + a3.x, a1.y + a3.y);
    // -i * (a1 - a3)
    const double2 t3 = make_double2(a1.y - a3.y, a3.x - a1.x);
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

// Forward transform (kernel exp(-2 pi i k n / N)).  `tw` = exp(-2 pi i k/1024),
// k < 768, in shared memory.  Caller must __syncwarp() after filling z.
template <int LOG2N>
__device__ __forceinline__ void warp_fft(double2* z, const double2* __restrict__ tw, int lane) {
  constexpr int N = 1 << LOG2N;
  FftPasses<LOG2N, 1>::run(z, tw, lane);
  if (LOG2N % 2 == 1) {
    constexpr int H = N / 2;
    constexpr int TS = 1024 / N;
    for (int k = lane; k < H; k += 32) {
      const int se = fft_swz(k), so = fft_swz(k + H);
      const double2 e = z[se];
      const double2 o = cmul(z[so], tw[k * TS]);
      z[se] = make_double2(e.x + o.x, e.y + o.y);
      z[so] = make_double2(e.x - o.x, e.y - o.y);
    }
    __syncwarp();
  }
}

}  // namespace peaq
