// Filter-bank ear model of advanced-mode PEAQ (fbearmodel.c), front half:
//
//   FB0 fb_flags_kernel   above-threshold flag of every 192-sample frame
//                         (is_frame_above_threshold, gstpeaq.c:1081-1099, called
//                         from process_fb_block :970-971)
//   FB1 fb_hp_kernel      playback-level scaling + DC-reject filter, two cascaded
//                         biquads (fbearmodel.c:289-303): sequential in time, one
//                         thread per stream, state carried between chunks
//   FB2 fb_bank_rec_kernel  the 40 complex FIR filters of 52..1456 taps, evaluated every
//                         32 samples (apply_filter_bank, fbearmodel.c:398-435), as
//                         sliding windowed DFTs: 384 FMAs per band and sub-step whatever
//                         the filter length (see the comment at the kernel); half of its
//                         coefficients come from constant memory through the uniform datapath
//       fb_bank_kernel    the same filters as polyphase direct FIRs (2 (N - 1) FMAs per
//                         band and sub-step); the engine's cross-check, selected with
//                         PEAQ_B200_FB_DIRECT=1 (tests/test_gpu_parity.py compares the two)
//
// The direct kernel computes 32 polyphase sub-convolutions: with delay d = 32 q + j,
//   out_b[s] = sum_j sum_q G_b[32 q + j] * x[32 (s - q) - j]
// so for a fixed phase j consecutive outputs slide over the same decimated
// input sequence.  Each lane owns R = 7 consecutive outputs and keeps the sliding
// window in registers: one shared-memory load and one coefficient load feed 14
// FMAs.  The input tile is staged in shared memory in polyphase-transposed layout
// (row stride 273 doubles: conflict-free for the transposing store and for the
// stride-7 window loads).
//
// Coefficients G_b[d] are indexed by total delay d (band delay D = 1 + (1456-N)/2
// folded in, fbearmodel.c:408), full length (the reference exploits the
// even/odd symmetry, :418-425; same products, different summation order), and
// reproduce the reference's ring-buffer alias: for band 0 the tap at delay 1456
// reads the NEWEST sample (fb_buf[offset + 1456] == fb_buf[offset]).
#include "peaq_engine.h"

#include <cstdlib>
#include <mutex>
#include <vector>

namespace peaq {
namespace {

constexpr int kR = 7;                    // outputs per lane
constexpr int kTileOut = 32 * kR;        // sub-steps per tile (224)
constexpr int kMaxQ = 46;                // ceil(1457 / 32)
constexpr int kTileM = kTileOut + kMaxQ + 1;   // decimated samples per phase (271)
constexpr int kRowStride = 273;
constexpr int kBankWarps = 8;

// ---------------------------------------------------------------------------

__global__ void fb_flags_kernel(PcmView pcm, unsigned first_frame, unsigned n_chunk_frames,
                                unsigned char* __restrict__ flags) {
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int pair = blockIdx.y;
  if (idx >= n_chunk_frames) return;
  const unsigned frame = first_frame + idx;
  unsigned char above = 0;
  if (frame < pcm.n_frames[pair]) {
    const int C = pcm.channels;
    const unsigned long long n = pcm.n_samples[pair];
    const float* __restrict__ sig = pcm.ref + pcm_pair_offset(pcm, pair);
    const unsigned long long s0 = (unsigned long long)frame * kFbFrame;
    for (int c = 0; c < C && !above; c++) {
      // literal replay: float running sum, double increments (gstpeaq.c:1088-1096)
      float sum = 0;
      double hist[5];
      int i;
      for (i = 0; i < 5; i++) {
        const unsigned long long s = s0 + i;
        hist[i] = fabs((double)(s < n ? __ldg(sig + s * C + c) : 0.f));
        sum = (float)((double)sum + hist[i]);
      }
      for (; i < kFbFrame; i++) {
        const unsigned long long s = s0 + i;
        const double v = fabs((double)(s < n ? __ldg(sig + s * C + c) : 0.f));
        sum = (float)((double)sum + (v - hist[i % 5]));
        hist[i % 5] = v;
        if ((double)sum >= 200. / 32768) {
          above = 1;
          break;
        }
      }
    }
  }
  flags[(size_t)pair * n_chunk_frames + idx] = above;
}

// ---------------------------------------------------------------------------
// FB1: level scaling (fbearmodel.c:289) and the two cascaded DC-reject biquads
// (fbearmodel.c:292-303).  hp buffer per stream: [kFbHist history samples][chunk samples].
//
// The recurrence is sequential in time: one sample costs a loop-carried chain of ~100 cycles,
// 25 ms per 10 s of audio however many SMs idle -- 1.5 s for a batch of 10-minute items.  It is
// therefore evaluated as a BLOCK SCAN over blocks of kHpL = 512 samples at ABSOLUTE positions
// of the item (block j = samples [512 j, 512 j + 512)):
//   e0[j]   = end state of block j run from a zero recursive state (its input history x[n-1],
//             x[n-2] is PCM, not state)
//   s0[j+1] = e0[j] + M s0[j]            M = homogeneous transition over 512 samples, s0[0] = 0
//   e1[j]   = end state of block j run from s0[j]
//   d[j+1]  = (e1[j] - s0[j+1]) + M d[j]                                              d[0] = 0
//   output of block j = the reference's own recurrence started from s[j] = s0[j] + d[j]
// The near-double pole of the sections makes M s cancel ~30:1, so s0 alone is ~1e-10 off; d is
// the (tiny) correction, whose own cancellation no longer matters.  Against the purely
// sequential recurrence the output differs by a rounding sequence: ~1e-11 of the signal level
// (MOVs <= 3e-12, tests/test_gpu_parity.py).
//
// TWO kernels compute exactly this, operation for operation:
//   fb_hp_kernel       one thread per stream walks the chunk sample by sample and carries all
//                      three recurrences (they are independent chains: same latency as one);
//                      any chunk start / length -- sessions, short chunks, huge batches
//   fb_hp_par_*        blocks in parallel (passes 0, 1, 2) with the two sequential per-stream
//                      scans in between; chunks that start on a block boundary -- few long items
// and both leave the same state behind (kHpHdr doubles per stream):
//   [0,1] x[n-1], x[n-2] | [2..5] output recurrence | [6..9] zero-state recurrence |
//   [10..13] recurrence from s0 | [14..17] s0[j] | [18..21] d[j]      (j = current block)
// so results do not depend on how an item is chunked, streamed or which kernel ran: every block
// starts from the same numbers.
constexpr int kHpL = 512;

struct HpTransition {
  double m[4][4];   // [to][from] over (y1a, y2a, y1b, y2b)
};

struct Biquads {
  double x1, x2, y1a, y2a, y1b, y2b;
  // one sample through the two cascaded DC-reject sections (fbearmodel.c:292-303)
  __device__ __forceinline__ double step(double scaled) {
    const double h1 = scaled - 2. * x1 + x2 + 1.99517 * y1a - 0.995174 * y2a;
    const double h2 = h1 - 2. * y1a + y2a + 1.99799 * y1b - 0.997998 * y2b;
    x2 = x1;
    x1 = scaled;
    y2a = y1a;
    y1a = h1;
    y2b = y1b;
    y1b = h2;
    return h2;
  }
  __device__ __forceinline__ void set_y(const double (&s)[4]) { y1a = s[0]; y2a = s[1]; y1b = s[2]; y2b = s[3]; }
  __device__ __forceinline__ void get_y(double (&s)[4]) const { s[0] = y1a; s[1] = y2a; s[2] = y1b; s[3] = y2b; }
};

// o = e + M s
__device__ __forceinline__ void hp_apply(const HpTransition& M, const double (&s)[4], const double (&e)[4],
                                         double (&o)[4]) {
#pragma unroll
  for (int r = 0; r < 4; r++)
    o[r] = fma(M.m[r][0], s[0], fma(M.m[r][1], s[1], fma(M.m[r][2], s[2], fma(M.m[r][3], s[3], e[r]))));
}

// block boundary: from (s0[j], d[j]) and the end states of block j to (s0[j+1], d[j+1])
__device__ __forceinline__ void hp_block_update(const HpTransition& M, double (&s0)[4], double (&d)[4],
                                                const double (&e0)[4], const double (&e1)[4]) {
  double s0n[4], r[4], dn[4];
  hp_apply(M, s0, e0, s0n);
#pragma unroll
  for (int k = 0; k < 4; k++) r[k] = e1[k] - s0n[k];   // nearly equal: exact
  hp_apply(M, d, r, dn);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    s0[k] = s0n[k];
    d[k] = dn[k];
  }
}

constexpr int kHpBlock = 16;   // samples per register block (input prefetch / 128-byte output rows)

// One thread per stream (pair, channel, side), sample by sample.  Input is fetched 16 samples
// ahead with 128-bit loads (the two channel threads of a signal read the same interleaved
// lines; the second read hits L1), output written as 128-byte rows.
template <int C>
__global__ void fb_hp_kernel(const DeviceTables* __restrict__ T, PcmView pcm, int n_streams,
                             unsigned long long t0, unsigned chunk_samples,
                             double* __restrict__ hp, size_t hp_stride,
                             double* __restrict__ hp_state /* [stream][kHpStateDoubles] */, HpTransition M,
                             unsigned long long pos0 /* absolute position of the chunk's first sample in the item */,
                             int first_chunk) {
  const int stream = blockIdx.x * blockDim.x + threadIdx.x;   // pair * 2C + 2c + side, as everywhere
  if (stream >= n_streams) return;
  const int pair = stream / (2 * C), c = (stream >> 1) % C, side = stream & 1;
  const unsigned long long n = side ? pcm.n_samples_test[pair] : pcm.n_samples[pair];
  const float* __restrict__ sig = (side ? pcm.test : pcm.ref) + pcm_pair_offset(pcm, pair);
  const bool aligned = (reinterpret_cast<uintptr_t>(sig) & 15) == 0;
  double* out = hp + (size_t)stream * hp_stride;
  double* st = hp_state + (size_t)stream * kHpStateDoubles;
  Biquads fo = Biquads{0, 0, 0, 0, 0, 0}, f0 = fo, f1 = fo;   // output, zero-state, from-s0 recurrences
  double s0[4] = {0., 0., 0., 0.}, d[4] = {0., 0., 0., 0.};
  if (!first_chunk) {
    fo = Biquads{st[0], st[1], st[2], st[3], st[4], st[5]};
    f0 = Biquads{st[0], st[1], st[6], st[7], st[8], st[9]};
    f1 = Biquads{st[0], st[1], st[10], st[11], st[12], st[13]};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      s0[k] = st[14 + k];
      d[k] = st[18 + k];
    }
  }
  {
    // history for the FIR bank: last kFbHist samples of the previous chunk
    double2* o2 = reinterpret_cast<double2*>(out);
    const double2* src = reinterpret_cast<const double2*>(st + kHpHdr);   // saved by the previous chunk
    for (int i = 0; i < kFbHist / 2; i++) o2[i] = first_chunk ? make_double2(0., 0.) : src[i];
  }
  const double lf = T->level_factor_fb;
  // samples past the end of the item are zero (do_flush pads the last frame,
  // gstpeaq.c:731-736); frames past the padded one are never read.
  // The raw samples of block i+1 are fetched before block i is filtered, so the
  // memory latency hides behind the (sequential) recurrence.
  constexpr int kVec = kHpBlock * C / 4;   // float4 loads per block
  float4 nxt[kVec];
  auto fetch = [&](unsigned i0) {
    const unsigned long long s = t0 + i0;
    if (aligned && s + kHpBlock <= n) {
      const float4* v = reinterpret_cast<const float4*>(sig + s * C);
#pragma unroll
      for (int q = 0; q < kVec; q++) nxt[q] = __ldg(v + q);
    } else {
#pragma unroll
      for (int q = 0; q < kVec; q++) {
        float e[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const unsigned long long idx = s * C + 4 * q + r;   // interleaved index
          e[r] = idx < n * C ? __ldg(sig + idx) : 0.f;
        }
        nxt[q] = make_float4(e[0], e[1], e[2], e[3]);
      }
    }
  };
  if (chunk_samples) fetch(0);
  for (unsigned i0 = 0; i0 < chunk_samples; i0 += kHpBlock) {
    float x[kHpBlock];
#pragma unroll
    for (int q = 0; q < kVec; q++) {
      const float e[4] = {nxt[q].x, nxt[q].y, nxt[q].z, nxt[q].w};
      if (C == 1) {
#pragma unroll
        for (int r = 0; r < 4; r++) x[4 * q + r] = e[r];
      } else {   // interleaved stereo: samples 2q, 2q+1 of channel c
        x[2 * q] = c ? e[1] : e[0];
        x[2 * q + 1] = c ? e[3] : e[2];
      }
    }
    if (i0 + kHpBlock < chunk_samples) fetch(i0 + kHpBlock);
    double y[kHpBlock];
#pragma unroll
    for (int k = 0; k < kHpBlock; k++) {
      const double scaled = x[k] * lf;   // fbearmodel.c:289
      y[k] = fo.step(scaled);
      f0.step(scaled);
      f1.step(scaled);
    }
    double2* o = reinterpret_cast<double2*>(out + kFbHist + i0);
#pragma unroll
    for (int k = 0; k < kHpBlock / 2; k++) o[k] = make_double2(y[2 * k], y[2 * k + 1]);
    // chunks are whole 192-sample frames, so the 16-sample groups tile the 512-sample blocks
    if (((pos0 + i0 + kHpBlock) & (kHpL - 1)) == 0) {
      double e0[4], e1[4], s[4];
      f0.get_y(e0);
      f1.get_y(e1);
      hp_block_update(M, s0, d, e0, e1);
#pragma unroll
      for (int k = 0; k < 4; k++) s[k] = s0[k] + d[k];
      const double z[4] = {0., 0., 0., 0.};
      fo.set_y(s);
      f0.set_y(z);
      f1.set_y(s0);
    }
  }
  st[0] = fo.x1; st[1] = fo.x2;
  st[2] = fo.y1a; st[3] = fo.y2a; st[4] = fo.y1b; st[5] = fo.y2b;
  st[6] = f0.y1a; st[7] = f0.y2a; st[8] = f0.y1b; st[9] = f0.y2b;
  st[10] = f1.y1a; st[11] = f1.y2a; st[12] = f1.y1b; st[13] = f1.y2b;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    st[14 + k] = s0[k];
    st[18 + k] = d[k];
  }
  // the last kFbHist filtered samples ([history | chunk] is contiguous in `out`) become
  // the next chunk's history; works for chunks shorter than the history too
  double2* dst = reinterpret_cast<double2*>(st + kHpHdr);
  const double2* src = reinterpret_cast<const double2*>(out + chunk_samples);
  for (int i = 0; i < kFbHist / 2; i++) dst[i] = src[i];
}

// ---- the same block scan with the blocks in parallel (chunks starting on a block boundary) ----
// starts/ends: [stream][n_blocks + 1][4]; MODE 0: zero-state pass -> ends; MODE 1: pass from
// starts (= s0) -> ends; MODE 2: output pass from starts (= s0 + d).  The last block may be
// partial (end of the chunk): its recurrences stop exactly at the chunk's end and their running
// states go to the stream's state header, like fb_hp_kernel leaves them.
// One CTA = 32 consecutive blocks of one signal (pair, side): lane = block, warp = channel.  The
// blocks' samples lie 2 KB (4 KB stereo) apart, so the PCM goes through shared memory: every
// sub-step the CTA copies the next 32 samples of its 32 blocks with 128-bit loads (16 lanes cover
// one block's 256-byte run), transposed into rows of odd pitch that the recurrences then read
// without bank conflicts; the output pass hands its results back the same way (rows of 32 doubles
// per block and channel, written out as 256-byte runs).
constexpr int kHpTileBlocks = 32;   // blocks per CTA
constexpr int kHpSub = 32;          // samples per sub-step

template <int C, int MODE>
__global__ void __launch_bounds__(32 * C)
fb_hp_par_block_kernel(const DeviceTables* __restrict__ T, PcmView pcm, int n_blocks,
                       unsigned long long t0, unsigned chunk_samples,
                       double* __restrict__ hp, size_t hp_stride,
                       double* __restrict__ hp_state, const double* __restrict__ starts,
                       double* __restrict__ ends, int first_chunk) {
  constexpr int kPitch = kHpSub * C + 1;                 // floats per block row
  constexpr int kVecPerBlock = kHpSub * C / 4;           // float4 per block and sub-step
  constexpr int kVecPerThread = kHpTileBlocks * kVecPerBlock / (32 * C);   // = 8
  __shared__ float in_s[2][kHpTileBlocks * kPitch];
  __shared__ double out_s[MODE == 2 ? C * kHpTileBlocks * (kHpSub + 1) : 1];
  const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;
  const int pair = blockIdx.x >> 1, side = blockIdx.x & 1;
  const int stream = pair * 2 * C + 2 * c + side;
  const int blk0 = blockIdx.y * kHpTileBlocks;
  const int blk = blk0 + lane;
  const bool have = blk < n_blocks;
  const unsigned long long n = side ? pcm.n_samples_test[pair] : pcm.n_samples[pair];
  const float* __restrict__ sig = (side ? pcm.test : pcm.ref) + pcm_pair_offset(pcm, pair);
  const bool aligned = (reinterpret_cast<uintptr_t>(sig + t0 * C) & 15) == 0;
  const double lf = T->level_factor_fb;
  const unsigned b0 = (unsigned)blk * kHpL;
  const unsigned len = have ? min((unsigned)kHpL, chunk_samples - b0) : 0u;   // a multiple of 64 (whole frames)
  // the longest block of the tile decides how many sub-steps the CTA runs
  const unsigned tile_b0 = (unsigned)blk0 * kHpL;
  const unsigned tile_len = min((unsigned)kHpL, chunk_samples - tile_b0);
  double* st = hp_state + (size_t)stream * kHpStateDoubles;
  Biquads f = Biquads{0, 0, 0, 0, 0, 0};
  if (have) {
    if (blk == 0) {
      if (!first_chunk) {
        f.x1 = st[0];
        f.x2 = st[1];
      }
    } else {
      const unsigned long long s1 = t0 + b0 - 1, s2 = t0 + b0 - 2;
      f.x1 = (s1 < n ? __ldg(sig + s1 * C + c) : 0.f) * lf;
      f.x2 = (s2 < n ? __ldg(sig + s2 * C + c) : 0.f) * lf;
    }
    if (MODE != 0) {
      const double* bs = starts + ((size_t)stream * (n_blocks + 1) + blk) * 4;
      f.y1a = bs[0]; f.y2a = bs[1]; f.y1b = bs[2]; f.y2b = bs[3];
    }
  }
  // the CTA's copy of sub-step i0: float4 q of the tile = block q / kVecPerBlock, offset q % kVecPerBlock
  float4 stage[kVecPerThread];
  auto fetch = [&](unsigned i0) {
#pragma unroll
    for (int k = 0; k < kVecPerThread; k++) {
      const int q = threadIdx.x + k * 32 * C;
      const int j = q / kVecPerBlock, w = q - j * kVecPerBlock;
      const unsigned jb0 = (unsigned)(blk0 + j) * kHpL;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (blk0 + j < n_blocks && jb0 + i0 < chunk_samples) {
        const unsigned long long e0 = (t0 + jb0 + i0) * C + 4 * w;   // first float of the vector, from sig
        if (aligned && e0 + 4 <= n * C) {
          v = __ldg(reinterpret_cast<const float4*>(sig + e0));
        } else {
          v.x = e0 < n * C ? __ldg(sig + e0) : 0.f;
          v.y = e0 + 1 < n * C ? __ldg(sig + e0 + 1) : 0.f;
          v.z = e0 + 2 < n * C ? __ldg(sig + e0 + 2) : 0.f;
          v.w = e0 + 3 < n * C ? __ldg(sig + e0 + 3) : 0.f;
        }
      }
      stage[k] = v;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int k = 0; k < kVecPerThread; k++) {
      const int q = threadIdx.x + k * 32 * C;
      const int j = q / kVecPerBlock, w = q - j * kVecPerBlock;
      float* d = &in_s[buf][j * kPitch + 4 * w];
      d[0] = stage[k].x; d[1] = stage[k].y; d[2] = stage[k].z; d[3] = stage[k].w;
    }
  };
  fetch(0);
  stash(0);
  __syncthreads();
  int buf = 0;
  for (unsigned i0 = 0; i0 < tile_len; i0 += kHpSub) {
    const bool more = i0 + kHpSub < tile_len;
    if (more) fetch(i0 + kHpSub);
    if (i0 < len) {
      const float* x = &in_s[buf][lane * kPitch + c];
      double* y = MODE == 2 ? &out_s[(c * kHpTileBlocks + lane) * (kHpSub + 1)] : nullptr;
#pragma unroll 8
      for (int k = 0; k < kHpSub; k++) {
        const double v = f.step(x[k * C] * lf);   // fbearmodel.c:289
        if (MODE == 2) y[k] = v;
      }
    }
    if (MODE == 2) {
      __syncthreads();
      // rows of 32 doubles per (channel, block) -> 256-byte runs of the filtered signal
      for (int e = threadIdx.x; e < C * kHpTileBlocks * (kHpSub / 2); e += 32 * C) {
        const int row = e / (kHpSub / 2), p2 = e - row * (kHpSub / 2);   // row = cc * 32 + j
        const int cc = row / kHpTileBlocks, j = row - cc * kHpTileBlocks;
        const unsigned jb0 = (unsigned)(blk0 + j) * kHpL;
        if (blk0 + j < n_blocks && jb0 + i0 < chunk_samples) {
          const double* r = &out_s[row * (kHpSub + 1) + 2 * p2];
          double* out = hp + (size_t)(pair * 2 * C + 2 * cc + side) * hp_stride + kFbHist + jb0 + i0;
          reinterpret_cast<double2*>(out)[p2] = make_double2(r[0], r[1]);
        }
      }
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  if (!have) return;
  if (MODE != 2 && len == (unsigned)kHpL) {
    double* be = ends + ((size_t)stream * (n_blocks + 1) + blk) * 4;
    be[0] = f.y1a; be[1] = f.y2a; be[2] = f.y1b; be[3] = f.y2b;
  }
  if (blk == n_blocks - 1) {
    // running states at the end of the chunk; when the last block is full the scans have already
    // moved on to the next block and fb_hp_par_scan_kernel<1> writes the header instead
    // (x[n-1], x[n-2] go to spare header slots: the CTA of block 0 may still be reading st[0..1];
    // fb_hp_par_hist_kernel moves them afterwards)
    if (MODE == 2) { st[22] = f.x1; st[23] = f.x2; }
    if (len < (unsigned)kHpL) {
      const int o = MODE == 2 ? 2 : (MODE == 0 ? 6 : 10);
      st[o] = f.y1a; st[o + 1] = f.y2a; st[o + 2] = f.y1b; st[o + 3] = f.y2b;
    }
  }
}

// MODE 0: a = zero-state end states e0[j] -> a = s0[j], j = 0..n_blocks (in place).
// MODE 1: a = s0[j], b = e1[j] -> a = s[j] = s0[j] + d[j]; leaves (s0, d) of the block the chunk
//         ends in -- and, if that block has not begun yet, its three start states -- in the header.
// One WARP per stream: the recurrences over the blocks are sequential (and evaluated in exactly
// the order fb_hp_kernel uses), but 32 blocks' inputs are read -- and their results written --
// together, as one coalesced kilobyte; the chain runs through the warp by shuffles, every lane
// carrying the same state and keeping the one that belongs to its block.
__device__ __forceinline__ void hp_bcast4(const double (&v)[4], int src, double (&o)[4]) {
#pragma unroll
  for (int k = 0; k < 4; k++) o[k] = __shfl_sync(0xffffffffu, v[k], src);
}

template <int MODE>
__global__ void fb_hp_par_scan_kernel(int n_streams, int n_blocks, unsigned chunk_samples,
                                      double* __restrict__ hp_state, double* __restrict__ a,
                                      const double* __restrict__ b, HpTransition M, int first_chunk) {
  const int stream = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (stream >= n_streams) return;
  double* pa = a + (size_t)stream * (n_blocks + 1) * 4;
  double* st = hp_state + (size_t)stream * kHpStateDoubles;
  const int n_full = (int)(chunk_samples / kHpL);   // n_blocks or n_blocks - 1
  if (MODE == 0) {
    double s[4] = {0., 0., 0., 0.};
    if (!first_chunk) { s[0] = st[14]; s[1] = st[15]; s[2] = st[16]; s[3] = st[17]; }
    for (int base = 0; base <= n_full; base += 32) {
      const int j = base + lane;
      double e[4] = {0., 0., 0., 0.}, mine[4] = {0., 0., 0., 0.};
      if (j < n_full) {
#pragma unroll
        for (int k = 0; k < 4; k++) e[k] = pa[4 * j + k];
      }
      const int steps = min(32, n_full + 1 - base);
      for (int t = 0; t < steps; t++) {
        if (lane == t) {
#pragma unroll
          for (int k = 0; k < 4; k++) mine[k] = s[k];
        }
        double et[4], o[4];
        hp_bcast4(e, t, et);
        if (base + t < n_full) {
          hp_apply(M, s, et, o);
#pragma unroll
          for (int k = 0; k < 4; k++) s[k] = o[k];
        }
      }
      if (j <= n_full) {
#pragma unroll
        for (int k = 0; k < 4; k++) pa[4 * j + k] = mine[k];
      }
    }
  } else {
    const double* pb = b + (size_t)stream * (n_blocks + 1) * 4;
    double d[4] = {0., 0., 0., 0.};
    if (!first_chunk) { d[0] = st[18]; d[1] = st[19]; d[2] = st[20]; d[3] = st[21]; }
    for (int base = 0; base <= n_full; base += 32) {
      const int j = base + lane;
      double s0[4] = {0., 0., 0., 0.}, r[4] = {0., 0., 0., 0.}, mine[4] = {0., 0., 0., 0.};
      if (j <= n_full) {
#pragma unroll
        for (int k = 0; k < 4; k++) s0[k] = pa[4 * j + k];
      }
      if (j < n_full) {
#pragma unroll
        for (int k = 0; k < 4; k++) r[k] = pb[4 * j + k] - pa[4 * (j + 1) + k];   // e1[j] - s0[j+1]: nearly equal, exact
      }
      const int steps = min(32, n_full + 1 - base);
      for (int t = 0; t < steps; t++) {
        if (lane == t) {
#pragma unroll
          for (int k = 0; k < 4; k++) mine[k] = d[k];
        }
        double rt[4], o[4];
        hp_bcast4(r, t, rt);
        if (base + t < n_full) {
          hp_apply(M, d, rt, o);
#pragma unroll
          for (int k = 0; k < 4; k++) d[k] = o[k];
        }
      }
      __syncwarp();   // every lane has read s0[j + 1] before lane j + 1 overwrites it
      if (j <= n_full) {
#pragma unroll
        for (int k = 0; k < 4; k++) pa[4 * j + k] = s0[k] + mine[k];
      }
      if (j == n_full) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          st[14 + k] = s0[k];
          st[18 + k] = mine[k];
        }
        if (n_full == n_blocks) {
          // the chunk ends on a block boundary: the next block starts from (s, 0, s0)
#pragma unroll
          for (int k = 0; k < 4; k++) {
            st[2 + k] = s0[k] + mine[k];
            st[6 + k] = 0.;
            st[10 + k] = s0[k];
          }
        }
      }
    }
  }
}

__global__ void fb_hp_par_hist_kernel(double* __restrict__ hp, size_t hp_stride, double* __restrict__ hp_state,
                                      unsigned chunk_samples, int mode, int first_chunk) {
  const int stream = blockIdx.x;
  double2* buf = reinterpret_cast<double2*>(hp + (size_t)stream * hp_stride);
  double2* st = reinterpret_cast<double2*>(hp_state + (size_t)stream * kHpStateDoubles + kHpHdr);
  if (mode == 0) {
    for (int i = threadIdx.x; i < kFbHist / 2; i += blockDim.x) buf[i] = first_chunk ? make_double2(0., 0.) : st[i];
  } else {
    // [history | chunk] is contiguous: the last kFbHist filtered samples, also for short chunks
    const double* src = hp + (size_t)stream * hp_stride + chunk_samples;
    double* dst = hp_state + (size_t)stream * kHpStateDoubles + kHpHdr;
    for (int i = threadIdx.x; i < kFbHist; i += blockDim.x) dst[i] = src[i];
    if (threadIdx.x == 0) {   // x[n-1], x[n-2] left by the output pass
      double* hdr = hp_state + (size_t)stream * kHpStateDoubles;
      hdr[0] = hdr[22];
      hdr[1] = hdr[23];
    }
  }
}

// ---------------------------------------------------------------------------
// FB2

struct BankBands {
  // bands handled by each of the kBankWarps warps (balanced by filter length)
  signed char band[kBankWarps][8];
};

__global__ void __launch_bounds__(32 * kBankWarps)
fb_bank_kernel(const DeviceTables* __restrict__ T, const double* __restrict__ hp, size_t hp_stride,
               unsigned n_sub /* sub-steps in this chunk */, double2* __restrict__ fbout,
               size_t out_stream_stride /* = 40 * n_sub */, BankBands bands, int n_tiles) {
  extern __shared__ __align__(16) double xs[];   // [32][kRowStride]
  const int stream = blockIdx.x / n_tiles;
  const int tile = blockIdx.x - stream * n_tiles;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S0 = tile * kTileOut;          // first sub-step of the tile (chunk-local)
  // decimated index m (chunk-local, x[32 m + p] = hpbuf[kFbHist + 32 m + p]); tile needs
  // m in [S0 - kMaxQ - 1, S0 + kTileOut)
  const int mbase = S0 - kMaxQ - 1;
  const double* __restrict__ src = hp + (size_t)stream * hp_stride + kFbHist;
  for (int i = threadIdx.x; i < kTileM * 32; i += blockDim.x) {
    const int t = mbase * 32 + i;          // chunk-local sample index (>= -kFbHist)
    const int p = i & 31, mm = i >> 5;
    const bool ok = t < (int)(n_sub * 32) && t >= -kFbHist;
    xs[p * kRowStride + mm] = ok ? src[t] : 0.;
  }
  __syncthreads();

  const int s0 = S0 + kR * lane;           // this lane's first output
  for (int slot = 0; slot < 8; slot++) {
    const int b = bands.band[warp][slot];
    if (b < 0) break;
    const int dlo = T->fb_dlo[b], dhi = T->fb_dhi[b];
    const double2* __restrict__ G = reinterpret_cast<const double2*>(T->fb_g) + T->fb_g_offset[b];
    double are[kR], aim[kR];
#pragma unroll
    for (int r = 0; r < kR; r++) are[r] = aim[r] = 0.;
    for (int j = 0; j < 32; j++) {
      // taps d = 32 q + j within [dlo, dhi]
      const int qlo = dlo > j ? (dlo - j + 31) >> 5 : 0;
      const int qhi = dhi >= j ? (dhi - j) >> 5 : -1;
      const int qn = qhi - qlo + 1;
      if (qn <= 0) continue;
      // x[32 (s - q) - j] = xs[p][s - q - moff - mbase]
      const int p = (32 - j) & 31, moff = j > 0 ? 1 : 0;
      const double* __restrict__ row = xs + p * kRowStride + (s0 - qlo - moff - mbase);
      // coefficients of this phase are contiguous: Gp[j][q]
      const double2* __restrict__ g = G + T->fb_phase_offset[b * 32 + j];
      double w[kR];
#pragma unroll
      for (int r = 0; r < kR; r++) w[r] = row[r];
      for (int u0 = 0; u0 < qn; u0 += kR) {
#pragma unroll
        for (int u = 0; u < kR; u++) {
          if (u0 + u < qn) {
            const double2 c = __ldg(g + u0 + u);
#pragma unroll
            for (int r = 0; r < kR; r++) {
              const double x = w[(r - u + kR) % kR];
              are[r] = fma(x, c.x, are[r]);
              aim[r] = fma(x, c.y, aim[r]);
            }
            // slide: the slot of output kR-1 is free, it becomes output 0 of tap q+1
            w[(kR - 1 - u) % kR] = row[-(u0 + u) - 1];
          }
        }
      }
    }
    // band-major output [stream][band][sub-step]: 7 consecutive sub-steps per lane
    double2* __restrict__ o = fbout + (size_t)stream * out_stream_stride + (size_t)b * n_sub;
#pragma unroll
    for (int r = 0; r < kR; r++)
      if (s0 + r < (int)n_sub) o[s0 + r] = make_double2(are[r], aim[r]);
  }
}


// ---------------------------------------------------------------------------
// FB2r: the filter bank as sliding windowed DFTs.
//
// Every tap set is a raised-cosine window times a complex exponential
// (fbearmodel.c:213-220): h[n] = (Wt/N) (2 - e^{j d n} - e^{-j d n}) e^{j w (n - N/2)},
// d = 2 pi / N.  So the output is the sum of three rectangular-window sums
//   S_f[s] = g_f sum_{n<N} e^{j w_f n} x[32 s - D - n],   w_f = w, w + d, w - d,
// and each of those obeys a one-step recursion in the sub-step index s:
//   S_f[s] = r_f S_f[s-1] + W_f[s],   r_f = e^{j 32 w_f},
//   W_f[s] = sum_{k<32} P_f[k] x[32 s - D - k] + Q_f[k] x[32 s - D - k - N]
// (the 32 samples that entered the window and the 32 that left it).  That is
// 384 FMAs per band and sub-step whatever N is, against 2 (N - 1) for the direct
// form (2910 for the longest band, 102 for the shortest): 15.4 k instead of 43.6 k
// FMAs per sub-step for the whole bank.  (Keeping the twelve shortest filters
// direct would save another 3 k FMAs; a fused direct pass for them was measured
// slower because of its constant-operand traffic and is not kept.)
//
// One CTA per stream walks through the chunk in tiles of 192 sub-steps; warp w
// owns bands w, w + 8, ..., w + 32.  Lane j owns one group of six consecutive
// sub-steps (one 192-sample frame): it computes W for them from the staged,
// polyphase-transposed input tile, turns them into the group's zero-state
// response P_j[i] = sum_{l<=i} r^{i-l} W[6 j + l]; lanes 0-2 (one per frequency)
// then chain the group totals sequentially, c_j = r^6 c_{j-1} + P_j[5], and
// every lane finishes with S[6 j + i] = r^{i+1} c_{j-1} + P_j[i].  Groups are
// anchored at absolute frame boundaries and chunks are whole frames, so the
// arithmetic of every output is the same however the item is cut into chunks;
// the chain value (3 complex per band) is the only state carried over.
// Rounding differs from the direct form at the 1e-14 level relative to the
// in-window signal (tests/test_gpu_parity.py compares both with the oracle).
constexpr int kRecTile = 32 * kFbRecGroup;       // 192 sub-steps
constexpr int kRecHistRows = 47;                 // 1488 samples of history / 32, rounded up
constexpr int kRecRows = kRecTile + kRecHistRows;
constexpr int kRecStride = 242;                  // even: the six columns of a lane keep their 16-byte alignment from row to row
constexpr int kRecWarps = 8;                     // 40 bands, 5 per warp (16 warps per SM: 4 per scheduler, 128 registers)
constexpr int kRecSlots = 5;
constexpr int kRecXsDoubles = 32 * kRecStride + 4;   // input tile incl. padding at both ends (even)

__device__ __forceinline__ double2 cfma(double2 a, double2 b, double2 c) {   // a * b + c
  return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}


// six consecutive doubles px[0..5] with 128-bit shared loads.  Lanes are 48 bytes apart, so
// every quarter-warp hits eight distinct 16-byte bank groups (64-bit loads at this stride
// would conflict two-way).  kOdd: px is 8 mod 16 (the same for all lanes of the warp).
template <bool kOdd>
__device__ __forceinline__ void load6(const double* __restrict__ px, double (&x)[6]) {
  if (!kOdd) {
    const double2 v0 = *reinterpret_cast<const double2*>(px);
    const double2 v1 = *reinterpret_cast<const double2*>(px + 2);
    const double2 v2 = *reinterpret_cast<const double2*>(px + 4);
    x[0] = v0.x; x[1] = v0.y; x[2] = v1.x; x[3] = v1.y; x[4] = v2.x; x[5] = v2.y;
  } else {
    const double2 v0 = *reinterpret_cast<const double2*>(px - 1);
    const double2 v1 = *reinterpret_cast<const double2*>(px + 1);
    const double2 v2 = *reinterpret_cast<const double2*>(px + 3);
    const double2 v3 = *reinterpret_cast<const double2*>(px + 5);
    x[0] = v0.y; x[1] = v1.x; x[2] = v1.y; x[3] = v2.x; x[4] = v2.y; x[5] = v3.x;
  }
}

// Coefficients of the ENTERING side in constant memory: P_f[k] of every band, [band][k][f]
// (61 440 bytes).  The inner loop reads them with warp-uniform addresses, i.e. through the uniform
// datapath (SASS: LDCU.64 UR, c[0x3][UR+imm]; DFMA R, R, UR, R), not through the shared-memory
// pipe.  That pipe bounded this kernel (88 % of its wavefront rate against 64 % of the FP64 pipe,
// profiles/r2_fb_bank_rec_ncu.txt): a broadcast 128-bit shared load costs four wavefronts like any
// other, so per tap index k the three coefficients cost as much as the lane's six samples
// (12 + 12..16 wavefronts against 18 cycles of DFMAs).  The constant path has a limit of its own
// -- with BOTH sides on it the loop stalls on the constant loads instead (MIO throttle, measured:
// 362.5 ms per 4096 pairs against 364.0 from shared memory) -- so the load is split: entering side
// from constant memory, leaving side (Q_f[k]) from the warp's shared-memory copy as before:
// 330.3 ms, the same bits.  The table is a model constant (no playback level in it): one copy per
// device.
__constant__ double2 c_fb_rec_p[kFbRecBands * 32 * 3];

// acc[f][i] += coef[3 k + f] * px[i] for k in [k0, k1), px moving up one row (one sample
// back in time) per k.  kConst: coef = c_fb_rec_p + cbase, else the shared-memory copy `coef`.
template <bool kOdd, bool kConst>
__device__ __forceinline__ void rec_accumulate(double2 (&acc)[3][kFbRecGroup], const double* __restrict__ px,
                                               const double2* __restrict__ coef, int cbase, int k0, int k1) {
#pragma unroll 2
  for (int k = k0; k < k1; k++) {
    double x[kFbRecGroup];
    load6<kOdd>(px, x);
    double2 c0, c1, c2;
    if (kConst) {
      c0 = c_fb_rec_p[cbase + 3 * k];
      c1 = c_fb_rec_p[cbase + 3 * k + 1];
      c2 = c_fb_rec_p[cbase + 3 * k + 2];
    } else {
      c0 = coef[3 * k];
      c1 = coef[3 * k + 1];
      c2 = coef[3 * k + 2];
    }
#pragma unroll
    for (int i = 0; i < kFbRecGroup; i++) {
      acc[0][i].x = fma(c0.x, x[i], acc[0][i].x);
      acc[0][i].y = fma(c0.y, x[i], acc[0][i].y);
      acc[1][i].x = fma(c1.x, x[i], acc[1][i].x);
      acc[1][i].y = fma(c1.y, x[i], acc[1][i].y);
      acc[2][i].x = fma(c2.x, x[i], acc[2][i].x);
      acc[2][i].y = fma(c2.y, x[i], acc[2][i].y);
    }
    px -= kRecStride;
  }
}

// the 32 samples x[32 s - a0 - k], k < 32, of the lane's six sub-steps against coef[3 k + f].
// x[32 s - a] sits in row (-a) mod 32, column s - ceil(a / 32) - mbase: the row falls by one
// per k and wraps once, where the column steps back and (the row pitch being even) the
// 16-byte alignment of the lane's six columns flips.
template <bool kConst>
__device__ __forceinline__ void rec_side(double2 (&acc)[3][kFbRecGroup], const double* __restrict__ col0, int a0,
                                         const double2* __restrict__ coef, int cbase) {
  const int row = (-a0) & 31, q = (a0 + 31) >> 5;
  const int k1 = row + 1;   // k < k1: before the wrap
  const double* __restrict__ px = col0 + row * kRecStride - q;
  if (reinterpret_cast<uintptr_t>(px) & 8) {
    rec_accumulate<true, kConst>(acc, px, coef, cbase, 0, k1);
    rec_accumulate<false, kConst>(acc, col0 + 31 * kRecStride - (q + 1), coef, cbase, k1, 32);
  } else {
    rec_accumulate<false, kConst>(acc, px, coef, cbase, 0, k1);
    rec_accumulate<true, kConst>(acc, col0 + 31 * kRecStride - (q + 1), coef, cbase, k1, 32);
  }
}

__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src)
               : "memory");
}

template <bool kConstP>   // entering side's coefficients from constant memory (default) or, like the leaving side's, from shared memory
__global__ void __launch_bounds__(32 * kRecWarps, 2)
fb_bank_rec_kernel(const DeviceTables* __restrict__ T, const double* __restrict__ hp, size_t hp_stride,
                   unsigned n_sub /* sub-steps in this chunk, a multiple of 6 */, double2* __restrict__ fbout,
                   size_t out_stream_stride /* = 40 * n_sub */, double* __restrict__ hp_state,
                   int first_chunk, const unsigned* __restrict__ n_frames, unsigned first_frame, int streams_per_pair) {
  extern __shared__ __align__(16) double xs_raw[];   // 2 doubles of padding, [32][kRecStride], then the warps' buffers
  double* xs = xs_raw + 2;                            // (a misaligned 128-bit load may start one double early)
  const int stream = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double2* xch = reinterpret_cast<double2*>(xs_raw + kRecXsDoubles) + warp * 96;   // [3][32]
  const double* __restrict__ src = hp + (size_t)stream * hp_stride + kFbHist;
  double2* carry_state = reinterpret_cast<double2*>(hp_state + (size_t)stream * kHpStateDoubles + kHpHdr + kFbHist);
  const int n_slots = warp + (kRecSlots - 1) * kRecWarps < kFbRecBands ? kRecSlots : kRecSlots - 1;
  // chain values of this warp's bands: [slot][frequency], kept in shared memory between tiles
  double2* cst = reinterpret_cast<double2*>(xs_raw + kRecXsDoubles) + kRecWarps * 96 + warp * (3 * kRecSlots);
  // this warp's copy of the current band's coefficient table (global loads of it kept
  // missing L1 behind the streaming input tile)
  double2* phs = reinterpret_cast<double2*>(xs_raw + kRecXsDoubles) + kRecWarps * (96 + 3 * kRecSlots) + warp * 192;
  if (lane < 3 * kRecSlots) {
    const int slot = lane / 3, f = lane - 3 * slot;
    cst[lane] = (!first_chunk && slot < n_slots) ? carry_state[(warp + kRecWarps * slot) * 3 + f] : make_double2(0., 0.);
  }
  __syncwarp();
  // tiles that hold frames of this stream's item (nothing downstream reads beyond them; a ragged
  // batch would otherwise filter zeros for as long as its longest item lasts)
  const unsigned total = n_frames[stream / streams_per_pair];
  const unsigned valid = total > first_frame ? min((total - first_frame) * 6u, n_sub) : 0u;
  const int n_tiles = (int)((valid + kRecTile - 1) / kRecTile);
  const int t_end = (int)(n_sub * 32);
  for (int tile = 0; tile < n_tiles; tile++) {
    const int S0 = tile * kRecTile;
    const int mbase = S0 - kRecHistRows;
    __syncthreads();   // the previous tile has been consumed
    // polyphase-transposed tile through asynchronous 8-byte copies: all of a thread's loads
    // are in flight together
    for (int i = threadIdx.x; i < kRecRows * 32; i += blockDim.x) {
      const int t = mbase * 32 + i;          // chunk-local sample index (>= -kFbHist)
      const int p = i & 31, mm = i >> 5;
      if (t < t_end) cp_async8(&xs[p * kRecStride + mm], src + t);
      else xs[p * kRecStride + mm] = 0.;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    // the next tile's new samples towards L2 while this one is being worked on
    {
      const int t_next = (S0 + kRecTile) * 32 + threadIdx.x * 16;   // 128-byte lines
      for (int t = t_next; t < (S0 + 2 * kRecTile) * 32 && t < t_end; t += blockDim.x * 16)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(src + t));
    }
    const int n_groups = min(32, (int)(n_sub - S0) / kFbRecGroup);
    const double* __restrict__ col0 = xs + kFbRecGroup * lane + kRecHistRows;   // column of x[32 s0 - 0]
    for (int slot = 0; slot < n_slots; slot++) {
      // (broadcasts from lane 0: tells the compiler that band and length are the same for the whole
      // warp, which puts the coefficient addresses and the loop counters on the uniform datapath)
      const int b = __shfl_sync(0xffffffffu, warp + kRecWarps * slot, 0);
      const int N = __shfl_sync(0xffffffffu, T->fb_len[b], 0);
      const int D = 1 + (kFbBuf - N) / 2;
      {
        // this band's coefficients [side][k][f] into the warp's shared-memory copy (kConstP: the
        // leaving side only)
        const double2* __restrict__ ph = reinterpret_cast<const double2*>(&T->fb_rec_ph[b][0][0]);
        if (kConstP) {
#pragma unroll
          for (int u = 0; u < 3; u++) {
            const int e = lane + 32 * u;          // = 3 k + f
            const int k = e / 3, f = e - 3 * k;
            phs[96 + e] = __ldg(ph + 6 * k + 3 + f);
          }
        } else {
#pragma unroll
          for (int u = 0; u < 6; u++) {
            const int e = lane + 32 * u;          // = 6 k + side * 3 + f
            const int k = e / 6, r = e - 6 * k;
            phs[(r >= 3 ? 96 : 0) + 3 * k + (r >= 3 ? r - 3 : r)] = __ldg(ph + e);
          }
        }
        __syncwarp();
      }
      double2 acc[3][kFbRecGroup];
#pragma unroll
      for (int f = 0; f < 3; f++)
#pragma unroll
        for (int i = 0; i < kFbRecGroup; i++) acc[f][i] = make_double2(0., 0.);
      rec_side<kConstP>(acc, col0, D, phs, b * 96);       // the 32 samples that entered the window
      rec_side<false>(acc, col0, D + N, phs + 96, 0);     // the 32 that left it
      // zero-state response of the group, P[i] = r P[i-1] + W[i]; group totals to the chain lanes
      const double2* __restrict__ rp = reinterpret_cast<const double2*>(&T->fb_rec_rpow[b][0][0]);
#pragma unroll
      for (int f = 0; f < 3; f++) {
        const double2 r1 = __ldg(rp + f * kFbRecGroup);
#pragma unroll
        for (int i = 1; i < kFbRecGroup; i++) acc[f][i] = cfma(r1, acc[f][i - 1], acc[f][i]);
        xch[f * 32 + lane] = acc[f][kFbRecGroup - 1];
      }
      __syncwarp();
      if (lane < 3) {
        const double2 r6 = __ldg(rp + lane * kFbRecGroup + kFbRecGroup - 1);
        double2 c = cst[3 * slot + lane];
        double2 tot = xch[lane * 32];
        for (int j = 0; j < n_groups; j++) {
          const double2 nxt = xch[lane * 32 + ((j + 1) & 31)];   // fetched ahead of the dependent chain
          xch[lane * 32 + j] = c;                                // chain value in front of group j
          c = cfma(r6, c, tot);
          tot = nxt;
        }
        cst[3 * slot + lane] = c;
      }
      __syncwarp();
      double2 out[kFbRecGroup];
#pragma unroll
      for (int i = 0; i < kFbRecGroup; i++) out[i] = make_double2(0., 0.);
#pragma unroll
      for (int f = 0; f < 3; f++) {
        const double2 cin = xch[f * 32 + lane];
#pragma unroll
        for (int i = 0; i < kFbRecGroup; i++) {
          const double2 sv = cfma(__ldg(rp + f * kFbRecGroup + i), cin, acc[f][i]);
          out[i].x += sv.x;
          out[i].y += sv.y;
        }
      }
      __syncwarp();   // xch and phs are reused by the next band
      if (b == 0) {
        // the reference's ring buffer holds 1456 samples, so band 0's last tap (delay 1456)
        // reads the newest sample instead (fbearmodel.c:311-313,413)
        const double2 g = make_double2(T->fb_rec_alias.x, T->fb_rec_alias.y);
        const double* __restrict__ pn = col0;                                              // x[32 s]
        const double* __restrict__ po = col0 + ((-kFbBuf) & 31) * kRecStride - ((kFbBuf + 31) >> 5);   // x[32 s - 1456]
#pragma unroll
        for (int i = 0; i < kFbRecGroup; i++) {
          const double dx = pn[i] - po[i];
          out[i].x = fma(g.x, dx, out[i].x);
          out[i].y = fma(g.y, dx, out[i].y);
        }
      }
      if (lane < n_groups) {
        double2* __restrict__ o = fbout + (size_t)stream * out_stream_stride + (size_t)b * n_sub + S0 + kFbRecGroup * lane;
#pragma unroll
        for (int i = 0; i < kFbRecGroup; i++) o[i] = out[i];
      }
    }
  }
  __syncwarp();
  if (lane < 3 * n_slots) {
    const int slot = lane / 3, f = lane - 3 * slot;
    carry_state[(warp + kRecWarps * slot) * 3 + f] = cst[lane];
  }
}

}  // namespace

cudaError_t launch_fb_flags(PcmView pcm, int n_pairs, unsigned first_frame, unsigned n_chunk_frames,
                            unsigned char* flags, cudaStream_t stream) {
  if (n_pairs <= 0 || n_chunk_frames == 0) return cudaSuccess;
  for (int p0 = 0; p0 < n_pairs; p0 += 65535) {
    const int np = n_pairs - p0 < 65535 ? n_pairs - p0 : 65535;
    PcmView v = pcm;
    if (pcm.base) {
      v.base += p0;
    } else {
      v.ref += (size_t)p0 * pcm.pair_stride;
      v.test += (size_t)p0 * pcm.pair_stride;
    }
    v.n_samples += p0;
    v.n_samples_test += p0;
    v.n_frames += p0;
    dim3 grid((n_chunk_frames + 127) / 128, np);
    fb_flags_kernel<<<grid, 128, 0, stream>>>(v, first_frame, n_chunk_frames,
                                              flags + (size_t)p0 * n_chunk_frames);
  }
  return cudaGetLastError();
}

cudaError_t launch_fb_hp(const DeviceTables* d_tables, PcmView pcm, int n_pairs,
                         unsigned long long t0, unsigned long long pos0, unsigned chunk_samples,
                         double* hp, size_t hp_stride, double* hp_state, bool first_chunk,
                         cudaStream_t stream) {
  const int n_streams = n_pairs * 2 * pcm.channels;
  if (n_streams <= 0) return cudaSuccess;
  // homogeneous transition of the recursive state over one block of kHpL samples
  static const HpTransition M = [] {
    HpTransition t;
    for (int from = 0; from < 4; from++) {
      double v[4] = {0., 0., 0., 0.};
      v[from] = 1.;
      for (int i = 0; i < kHpL; i++) {
        const double h1 = 1.99517 * v[0] - 0.995174 * v[1];
        const double h2 = h1 - 2. * v[0] + v[1] + 1.99799 * v[2] - 0.997998 * v[3];
        v[1] = v[0]; v[0] = h1; v[3] = v[2]; v[2] = h2;
      }
      for (int to = 0; to < 4; to++) t.m[to][from] = v[to];
    }
    return t;
  }();
  // Blocks in parallel when the sample-by-sample walk would leave the GPU idle: it keeps one
  // thread per stream busy for ~52 ns per sample whatever the batch, the parallel form does three
  // times the arithmetic on every SM.  Same results either way (PEAQ_B200_HP_PARALLEL=0/1 forces one).
  static const int forced = std::getenv("PEAQ_B200_HP_PARALLEL") ? std::atoi(std::getenv("PEAQ_B200_HP_PARALLEL")) : -1;
  // (the block-parallel grid carries the tiles of 32 blocks in its y dimension: 65 535 x 16 384 samples)
  const bool can = chunk_samples >= 4 * kHpL && pos0 % kHpL == 0 && chunk_samples / (kHpL * kHpTileBlocks) < 65535u;
  const bool parallel = can && (forced == 1 || (forced < 0 && n_streams <= 4096));
  if (parallel) {
    const int n_blocks = (int)((chunk_samples + kHpL - 1) / kHpL);
    const size_t n_state = (size_t)n_streams * (n_blocks + 1) * 4;
    double *sa = nullptr, *sb = nullptr;   // scratch from the stream-ordered allocator
    cudaError_t e = cudaMallocAsync(&sa, n_state * sizeof(double), stream);
    if (e != cudaSuccess) return e;
    e = cudaMallocAsync(&sb, n_state * sizeof(double), stream);
    if (e != cudaSuccess) return e;
    const int fc = first_chunk ? 1 : 0;
    const unsigned sgrid = (unsigned)((n_streams + 3) / 4);   // one warp per stream
    const dim3 grid((unsigned)n_pairs * 2, (unsigned)((n_blocks + kHpTileBlocks - 1) / kHpTileBlocks));
    fb_hp_par_hist_kernel<<<n_streams, 128, 0, stream>>>(hp, hp_stride, hp_state, chunk_samples, 0, fc);
    if (pcm.channels == 2) {
      fb_hp_par_block_kernel<2, 0><<<grid, 64, 0, stream>>>(d_tables, pcm, n_blocks, t0, chunk_samples, hp, hp_stride, hp_state, nullptr, sa, fc);
      fb_hp_par_scan_kernel<0><<<sgrid, 128, 0, stream>>>(n_streams, n_blocks, chunk_samples, hp_state, sa, sb, M, fc);
      fb_hp_par_block_kernel<2, 1><<<grid, 64, 0, stream>>>(d_tables, pcm, n_blocks, t0, chunk_samples, hp, hp_stride, hp_state, sa, sb, fc);
      fb_hp_par_scan_kernel<1><<<sgrid, 128, 0, stream>>>(n_streams, n_blocks, chunk_samples, hp_state, sa, sb, M, fc);
      fb_hp_par_block_kernel<2, 2><<<grid, 64, 0, stream>>>(d_tables, pcm, n_blocks, t0, chunk_samples, hp, hp_stride, hp_state, sa, sb, fc);
    } else {
      fb_hp_par_block_kernel<1, 0><<<grid, 32, 0, stream>>>(d_tables, pcm, n_blocks, t0, chunk_samples, hp, hp_stride, hp_state, nullptr, sa, fc);
      fb_hp_par_scan_kernel<0><<<sgrid, 128, 0, stream>>>(n_streams, n_blocks, chunk_samples, hp_state, sa, sb, M, fc);
      fb_hp_par_block_kernel<1, 1><<<grid, 32, 0, stream>>>(d_tables, pcm, n_blocks, t0, chunk_samples, hp, hp_stride, hp_state, sa, sb, fc);
      fb_hp_par_scan_kernel<1><<<sgrid, 128, 0, stream>>>(n_streams, n_blocks, chunk_samples, hp_state, sa, sb, M, fc);
      fb_hp_par_block_kernel<1, 2><<<grid, 32, 0, stream>>>(d_tables, pcm, n_blocks, t0, chunk_samples, hp, hp_stride, hp_state, sa, sb, fc);
    }
    fb_hp_par_hist_kernel<<<n_streams, 128, 0, stream>>>(hp, hp_stride, hp_state, chunk_samples, 1, fc);
    cudaFreeAsync(sa, stream);
    cudaFreeAsync(sb, stream);
    return cudaGetLastError();
  }
  // few threads per block so the streams spread over all SMs (latency-bound scan)
  const int block = 32;
  const int grid = (n_streams + block - 1) / block;
  if (pcm.channels == 2) {
    fb_hp_kernel<2><<<grid, block, 0, stream>>>(d_tables, pcm, n_streams, t0, chunk_samples, hp, hp_stride, hp_state, M,
                                                pos0, first_chunk ? 1 : 0);
  } else {
    fb_hp_kernel<1><<<grid, block, 0, stream>>>(d_tables, pcm, n_streams, t0, chunk_samples, hp, hp_stride, hp_state, M,
                                                pos0, first_chunk ? 1 : 0);
  }
  return cudaGetLastError();
}

// The entering side's recursion coefficients into the constant memory of the current device, once per device and
// process (they are the same for every engine: nothing in them depends on the playback level).
static cudaError_t fb_bank_upload_constants(const DeviceTables* h_tables) {
  static std::mutex mu;
  static bool done[256] = {false};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  if (dev >= 0 && dev < 256 && done[dev]) return cudaSuccess;
  std::vector<double2> p((size_t)kFbRecBands * 96);
  for (int b = 0; b < kFbRecBands; b++)
    for (int k = 0; k < 32; k++)
      for (int f = 0; f < 3; f++) p[(size_t)b * 96 + 3 * k + f] = make_double2(h_tables->fb_rec_ph[b][k][f].x, h_tables->fb_rec_ph[b][k][f].y);
  e = cudaMemcpyToSymbol(c_fb_rec_p, p.data(), p.size() * sizeof(double2));
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 256) done[dev] = true;
  return cudaSuccess;
}

cudaError_t launch_fb_bank(const DeviceTables* d_tables, const DeviceTables* h_tables,
                           const double* hp, size_t hp_stride, int n_streams, unsigned n_sub,
                           double* fbout, double* hp_state, bool first_chunk, bool direct_only,
                           const unsigned* n_frames, unsigned first_frame, int streams_per_pair,
                           cudaStream_t stream) {
  if (n_streams <= 0 || n_sub == 0) return cudaSuccess;
  // default: all filters through the recursion (fb_bank_rec_kernel); direct_only: all 40 as
  // polyphase direct FIRs (cross-check)
  if (!direct_only) {
    cudaError_t e;
    // PEAQ_B200_FB_SMEM_COEF=1: both sides' coefficients from shared memory (the kernel's earlier
    // form; same results bit for bit)
    static const bool smem_coef = std::getenv("PEAQ_B200_FB_SMEM_COEF") && std::atoi(std::getenv("PEAQ_B200_FB_SMEM_COEF"));
    const size_t smem_rec = sizeof(double) * kRecXsDoubles + sizeof(double2) * (96 + 3 * kRecSlots + 192) * kRecWarps;
    auto kernel = smem_coef ? fb_bank_rec_kernel<false> : fb_bank_rec_kernel<true>;
    if (!smem_coef) {
      e = fb_bank_upload_constants(h_tables);
      if (e != cudaSuccess) return e;
    }
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rec);
    if (e != cudaSuccess) return e;
    kernel<<<(unsigned)n_streams, 32 * kRecWarps, smem_rec, stream>>>(
        d_tables, hp, hp_stride, n_sub, reinterpret_cast<double2*>(fbout), (size_t)kFbBands * n_sub, hp_state,
        first_chunk ? 1 : 0, n_frames, first_frame, streams_per_pair);
    return cudaGetLastError();
  }
  // distribute the 40 bands over the warps, longest filters first, always onto
  // the least loaded warp
  BankBands bb;
  int load[kBankWarps] = {0};
  int count[kBankWarps] = {0};
  for (int w = 0; w < kBankWarps; w++)
    for (int s = 0; s < 8; s++) bb.band[w][s] = -1;
  for (int b = 0; b < kFbBands; b++) {   // lengths are sorted descending already
    int best = 0;
    for (int w = 1; w < kBankWarps; w++)
      if (load[w] < load[best]) best = w;
    bb.band[best][count[best]++] = (signed char)b;
    load[best] += h_tables->fb_len[b] + 64;
  }
  const size_t smem = sizeof(double) * 32 * kRowStride;
  // per device and cheap: set on every launch
  cudaError_t e = cudaFuncSetAttribute(fb_bank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int n_tiles = (int)((n_sub + kTileOut - 1) / kTileOut);
  fb_bank_kernel<<<(unsigned)n_streams * n_tiles, 32 * kBankWarps, smem, stream>>>(
      d_tables, hp, hp_stride, n_sub, reinterpret_cast<double2*>(fbout), (size_t)kFbBands * n_sub, bb,
      n_tiles);
  return cudaGetLastError();
}

}  // namespace peaq
