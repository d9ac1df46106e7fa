// Double-precision FFTs in shared memory for the frame kernel (sm_100a).
//
// A group of NT threads (one warp, or the two warps of a stream) transforms N = 2^LOG2N complex
// points held in shared memory: in-place radix-4 decimation in time (plus one radix-2 pass when
// LOG2N is odd).  The caller stores input sample n at slot fft_slot<LOG2N>(n) (base-4 digit
// reversal, XOR-swizzled); the result is in natural order behind the same swizzle:
// X[k] = z[fft_swz(k)].  Two forms:
//   * fft_pass4 / group_fft / warp_fft: one radix-4 level per trip through shared memory (the EHS
//     transforms of 512, 256 and 128 points; their first level runs in the caller's registers);
//   * fft1024_*: the 1024-point transform of the frame's spectrum, TWO levels per trip, ending in
//     registers, the mirror values of the real-FFT split exchanged through a plain 8 x 64 array.
//
// The swizzle folds index bits 3,4,6,8,9 into the low three bits so that every 128-bit
// shared-memory access of every pass, the digit-reversed input scatter and the natural-order
// read-out are bank-conflict free for N = 1024.  It is linear over GF(2) (swz(a ^ b) = swz(a) ^
// swz(b)) and every index used by a pass is lane_part ^ compile_time_part with disjoint bits, so
// a pass computes ONE swizzle per lane; all other addresses are XORs with constants.
//
// Twiddles: exp(-2 pi i q / 1024) for q < 512 in shared memory, behind a swizzle of their own
// (fft_twi); the second half of the circle is the negated first half.
//
// tests/host/fft_check.cu replays the passes thread by thread on the host (bit equality of both
// forms, a long-double DFT, the wavefront count of every access).
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

__host__ __device__ __forceinline__ double2 cmul(double2 a, double2 w) {
  // explicit fused multiply-adds: the FFT's rounding need not (and cannot)
  // match kissfft's, only its accuracy
  return make_double2(fma(a.x, w.x, -a.y * w.y), fma(a.x, w.y, a.y * w.x));
}

// The half-circle twiddle table is stored behind a swizzle of its own: entry q lives at slot
// fft_twi(q), index bits 3..5 and 6..8 folded into the low three bits.  The passes read it with
// strides 64, 16, 4 and 1 (k * 1024 / (4 LQ)) and the EHS with strides 2 and 8; in natural order
// the strided 128-bit reads fell into one or two of the eight 16-byte bank groups (864 instead of
// 160 shared-memory wavefronts per frame, profiles/r2_fft_frames_ncu.txt), behind the swizzle the
// lanes of every quarter-warp hit eight different groups for every one of those strides.
// Linear over GF(2) like fft_swz.
__host__ __device__ constexpr int fft_twi(int q) { return q ^ ((q >> 3) & 7) ^ ((q >> 6) & 7); }

// twiddle exp(-2 pi i q / 1024), q < 1024, from the half-circle table
__host__ __device__ __forceinline__ double2 fft_tw(const double2* __restrict__ tw, int q) {
  const double2 w = tw[fft_twi(q & 511)];
  return (q & 512) ? make_double2(-w.x, -w.y) : w;
}

// Passes are written for NT cooperating threads (32 = one warp, 64 = the two warps
// of a channel); thread t handles butterflies b = t + NT u.  `Sync` is the barrier
// of the group (__syncwarp for a warp, a named barrier for two warps).
struct WarpSync {
  __device__ __forceinline__ void operator()() const { __syncwarp(); }
};

// index contributed by iteration u of a pass (added to the thread part, disjoint bits)
template <int LQ, int NT>
__host__ __device__ constexpr int fft_pass_delta(int u) {
  return LQ <= NT ? 4 * NT * u : (NT * (u % (LQ / NT))) | (4 * LQ * (u / (LQ / NT)));
}

// one radix-4 pass; LQ = size of the sub-transforms being combined
template <int LOG2N, int LQ, int NT, typename Sync>
__host__ __device__ __forceinline__ void fft_pass4(double2* z, const double2* __restrict__ tw, int t, Sync sync) {
  constexpr int N = 1 << LOG2N;
  constexpr int TS = 1024 / (4 * LQ);  // stride into the 1024-point twiddle circle
  // butterfly b = t + NT u combines i0 + {0,1,2,3} LQ with i0 = 4 (b - k) + k, k = b mod LQ
  const int k_t = LQ <= NT ? (t & (LQ - 1)) : t;
  const int base = LQ <= NT ? ((t - k_t) << 2) + k_t : t;
  const int sb = fft_swz(base);
  double2 w1, w2, w3;
  // w^2 and w^3 are computed from w instead of loaded: the strided table reads conflict in the
  // banks (they were half of all excess shared-memory wavefronts of the frame kernel) and the
  // FP64 pipe has room; |w^2 - table| ~ 2e-16, irrelevant for the transform's accuracy
  if (LQ > 1 && LQ <= NT) {   // twiddles depend on the thread only
    w1 = tw[fft_twi(k_t * TS)];
    w2 = cmul(w1, w1);
    w3 = cmul(w2, w1);
  }
  // (fully unrolled on purpose: with the loop rolled the swizzled addresses become run-time
  // arithmetic and the frame kernel measured 146 -> 169 ms, although the code shrinks by 7 %)
#pragma unroll
  for (int u = 0; u < N / (4 * NT); u++) {
    const int d = fft_pass_delta<LQ, NT>(u);
    const int s0 = sb ^ fft_swz(d), s1 = sb ^ fft_swz(d | LQ), s2 = sb ^ fft_swz(d | (2 * LQ)),
              s3 = sb ^ fft_swz(d | (3 * LQ));
    double2 a0 = z[s0], a1 = z[s1], a2 = z[s2], a3 = z[s3];
    if constexpr (LQ > NT) {
      const int k = t | (NT * (u % (LQ / NT)));
      w1 = tw[fft_twi(t * TS) ^ fft_twi(NT * (u % (LQ / NT)) * TS)];
      w2 = cmul(w1, w1);
      w3 = cmul(w2, w1);
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
  sync();
}

template <int LOG2N, int LQ, int NT, typename Sync>
struct FftPasses {
  static __device__ __forceinline__ void run(double2* z, const double2* __restrict__ tw, int t, Sync sync) {
    fft_pass4<LOG2N, LQ, NT, Sync>(z, tw, t, sync);
    FftPasses<LOG2N, LQ * 4, NT, Sync>::run(z, tw, t, sync);
  }
};
// recursion ends once LQ reaches 4^(LOG2N/2)
template <int NT, typename Sync> struct FftPasses<10, 1024, NT, Sync> { static __device__ __forceinline__ void run(double2*, const double2*, int, Sync) {} };
template <int NT, typename Sync> struct FftPasses<9, 256, NT, Sync> { static __device__ __forceinline__ void run(double2*, const double2*, int, Sync) {} };
template <int NT, typename Sync> struct FftPasses<7, 64, NT, Sync> { static __device__ __forceinline__ void run(double2*, const double2*, int, Sync) {} };
template <int NT, typename Sync> struct FftPasses<8, 256, NT, Sync> { static __device__ __forceinline__ void run(double2*, const double2*, int, Sync) {} };

// Forward transform (kernel exp(-2 pi i k n / N)) by NT threads.  `tw` =
// exp(-2 pi i q/1024), q < 512, in shared memory.  The group must be
// synchronised after filling z; it is synchronised again on return.
// LQ0 = 4: the caller has run the first radix-4 level itself (in registers, while filling z).
template <int LOG2N, int NT, typename Sync, int LQ0 = 1>
__device__ __forceinline__ void group_fft(double2* z, const double2* __restrict__ tw, int t, Sync sync) {
  constexpr int N = 1 << LOG2N;
  FftPasses<LOG2N, LQ0, NT, Sync>::run(z, tw, t, sync);
  if (LOG2N % 2 == 1) {
    constexpr int H = N / 2;
    constexpr int TS = 1024 / N;
    const int st = fft_swz(t);
#pragma unroll
    for (int u = 0; u < H / NT; u++) {
      const int se = st ^ fft_swz(NT * u), so = st ^ fft_swz(NT * u + H);
      const double2 e = z[se];
      const double2 o = cmul(z[so], tw[fft_twi(t * TS) ^ fft_twi(NT * u * TS)]);
      z[se] = make_double2(e.x + o.x, e.y + o.y);
      z[so] = make_double2(e.x - o.x, e.y - o.y);
    }
    sync();
  }
}

// ---- 1024 points by 64 threads, two radix-4 levels per trip through shared memory ------------
// A thread of fft_pass4 handles four butterflies per pass -- sixteen points in, sixteen out -- and
// the frame kernel is bound by the shared-memory pipe (81 % of its wavefront rate, the FFT passes
// being 40 % of the wavefronts).  Here the four butterflies a thread runs at level LQ are chosen so
// that their outputs are exactly the inputs of four butterflies at level 4 LQ: the group
//   i = h 16 LQ + b 4 LQ + a LQ + k,  a, b = 0..3    (k < LQ, h < 64 / LQ: 64 groups, one per thread)
// takes level LQ along a (twiddle index k) and level 4 LQ along b (twiddle index a LQ + k) in
// registers, so the 1024-point transform needs two trips (levels 4 + 16, levels 64 + 256) after the
// register-resident first level instead of four.  Butterflies, twiddles and their order are those
// of fft_pass4: the transform is bit for bit the same.
// radix-4 butterfly on twiddled inputs (the additions of fft_pass4)
__host__ __device__ __forceinline__ void fft_bfly4_nt(double2& a0, double2& a1, double2& a2, double2& a3) {
  const double2 t0 = make_double2(a0.x + a2.x, a0.y + a2.y);
  const double2 t1 = make_double2(a0.x - a2.x, a0.y - a2.y);
  const double2 t2 = make_double2(a1.x + a3.x, a1.y + a3.y);
  const double2 t3 = make_double2(a1.y - a3.y, a3.x - a1.x);   // -i (a1 - a3)
  a0 = make_double2(t0.x + t2.x, t0.y + t2.y);
  a1 = make_double2(t1.x + t3.x, t1.y + t3.y);
  a2 = make_double2(t0.x - t2.x, t0.y - t2.y);
  a3 = make_double2(t1.x - t3.x, t1.y - t3.y);
}

// level LQ of group (h, k) on e[4 b + a], b = 0..3: one twiddle for all four butterflies, applied
// power by power so that at most two of w, w^2, w^3 are alive next to the sixteen points
template <int LQ>
__host__ __device__ __forceinline__ void fft_level_inner(double2 (&e)[16], const double2* __restrict__ tw, int k) {
  constexpr int TS1 = 1024 / (4 * LQ);
  const double2 w1 = tw[fft_twi(k * TS1)];
#pragma unroll
  for (int b = 0; b < 4; b++) e[4 * b + 1] = cmul(e[4 * b + 1], w1);
  const double2 w2 = cmul(w1, w1);
#pragma unroll
  for (int b = 0; b < 4; b++) e[4 * b + 2] = cmul(e[4 * b + 2], w2);
  const double2 w3 = cmul(w2, w1);
#pragma unroll
  for (int b = 0; b < 4; b++) e[4 * b + 3] = cmul(e[4 * b + 3], w3);
#pragma unroll
  for (int b = 0; b < 4; b++) fft_bfly4_nt(e[4 * b], e[4 * b + 1], e[4 * b + 2], e[4 * b + 3]);
}

// butterfly a of level 4 LQ of group (h, k): e[a], e[4 + a], e[8 + a], e[12 + a], twiddle index a LQ + k
template <int LQ>
__host__ __device__ __forceinline__ void fft_level_outer(double2 (&e)[16], const double2* __restrict__ tw, int q2, int a) {
  constexpr int TS2 = 1024 / (16 * LQ);
  const double2 w1 = tw[q2 ^ fft_twi(a * LQ * TS2)];   // q2 = fft_twi(k TS2), k < LQ: disjoint bits
  e[4 + a] = cmul(e[4 + a], w1);
  const double2 w2 = cmul(w1, w1);
  e[8 + a] = cmul(e[8 + a], w2);
  const double2 w3 = cmul(w2, w1);
  e[12 + a] = cmul(e[12 + a], w3);
  fft_bfly4_nt(e[a], e[4 + a], e[8 + a], e[12 + a]);
}

// group of thread t < 64 for the trip over levels 4 and 16: k = t & 3, h = 8 t_2 + (t >> 3); the
// lanes of a quarter-warp then differ in index bits 0, 1 and 9, which fft_swz maps to eight
// different bank groups.  Returns the group's first index h 64 + k.
__host__ __device__ constexpr int fft_group_a(int t) { return ((((t >> 2) & 1) * 8 + (t >> 3)) << 6) | (t & 3); }

// levels 4 and 16, in place (z holds the output of level 1 behind fft_swz); the caller
// synchronises the 64 threads before and after
__host__ __device__ __forceinline__ void fft1024_levels_4_16(double2* z, const double2* __restrict__ tw, int t) {
  const int sb = fft_swz(fft_group_a(t));
  double2 e[16];
#pragma unroll
  for (int j = 0; j < 16; j++) e[j] = z[sb ^ fft_swz(4 * j)];        // b 16 + a 4 = 4 (4 b + a)
  fft_level_inner<4>(e, tw, t & 3);
  const int q2 = fft_twi((t & 3) * 16);
#pragma unroll
  for (int a = 0; a < 4; a++) {
    fft_level_outer<4>(e, tw, q2, a);
#pragma unroll
    for (int b = 0; b < 4; b++) z[sb ^ fft_swz(4 * (4 * b + a))] = e[4 * b + a];
  }
}

// inputs of levels 64 and 256 (group k = t, h = 0): e[j] <- position t + 64 j
__host__ __device__ __forceinline__ void fft1024_load_64_256(double2 (&e)[16], const double2* z, int t) {
  const int sb = fft_swz(t);
#pragma unroll
  for (int j = 0; j < 16; j++) e[j] = z[sb ^ fft_swz(64 * j)];
}

// levels 64 and 256 on the loaded inputs; afterwards e[u] = Z[t + 64 u].  The upper halves
// e[8..15] = Z[512 + t + 64 (u - 8)] are left in the exchange array `xch` (the stream's buffer
// itself: every thread must have loaded its inputs, i.e. the 64 threads synchronise between
// fft1024_load_64_256 and this), row u - 8, column (t - 1) & 63, unswizzled: consecutive lanes,
// consecutive 16-byte slots.
__host__ __device__ __forceinline__ void fft1024_levels_64_256(double2 (&e)[16], double2* xch,
                                                               const double2* __restrict__ tw, int t) {
  fft_level_inner<64>(e, tw, t);
  const int col = (t - 1) & 63;
  const int q2 = fft_twi(t);
#pragma unroll
  for (int a = 0; a < 4; a++) {
    fft_level_outer<64>(e, tw, q2, a);
    xch[a * 64 + col] = e[8 + a];
    xch[(4 + a) * 64 + col] = e[12 + a];
  }
}

// Z[512] (thread 0's e[8]) in the exchange array
__host__ __device__ constexpr int fft1024_xch_mid() { return 63; }

// Z[(1024 - k) & 1023] for k = t + 64 u, u < 8, out of the exchange array (after a barrier):
// = e[15 - u] of thread 64 - t for t > 0; thread 0 is its own partner one register further up
// (Z[1024 - 64 u] = its e[16 - u]); Z[0] mirrors itself (`own` = the caller's e[u]).
__host__ __device__ __forceinline__ double2 fft1024_mirror(const double2* xch, int t, int u, double2 own) {
  const double2 q = xch[(7 - u + (t == 0 ? 1 : 0)) * 64 + (63 - t)];   // row 8 (t = 0, u = 0): inside the buffer, unused
  return (u == 0 && t == 0) ? own : q;
}

template <int LOG2N>
__device__ __forceinline__ void warp_fft(double2* z, const double2* __restrict__ tw, int lane) {
  group_fft<LOG2N, 32, WarpSync>(z, tw, lane, WarpSync());
}

}  // namespace peaq
