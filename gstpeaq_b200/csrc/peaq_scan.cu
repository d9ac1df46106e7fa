// K2 `scan_basic`: the recurrent half of basic-mode PEAQ as a kernel of its own -- one CTA per
// pair walks the pair's frames in order over the per-frame records K1 wrote, carrying every
// recurrence in registers across the frame loop; the state is loaded from / stored to global
// memory at chunk boundaries, so hour-long items and streaming sessions run as a sequence of
// chunks.  Large batches run the fused persistent kernel (peaq_fused.cu) instead; the device code
// (scan_step, scan_epilogue) is shared, peaq_scan.cuh.
#include "peaq_scan.cuh"

namespace peaq {
namespace {

__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src)
               : "memory");
}

#if defined(PEAQ_DEV_K2_OCC3)
#define PEAQ_K2_MIN_CTAS 3
#else
#define PEAQ_K2_MIN_CTAS 2
#endif
template <bool kTap>
__global__ void __launch_bounds__(kGroup * kMaxChannels, PEAQ_K2_MIN_CTAS)
scan_basic_kernel(const DeviceTables* __restrict__ T, const double* __restrict__ records,
                  RecordLayout L, const unsigned* __restrict__ n_frames, unsigned first_frame,
                  unsigned n_chunk_frames, double* __restrict__ state, StateLayout S,
                  PairResult* __restrict__ results, double* __restrict__ dbg) {
  const int C = L.C, B = L.B;
  const int pair = blockIdx.x;
  const ScanThread th = make_scan_thread(B);
  __shared__ ScanShared sh;

  double* st = state + (size_t)pair * S.stride;

  // ---- load state -----------------------------------------------------------
  double bs[kBandStateFields];
#pragma unroll
  for (int f = 0; f < kBandStateFields; f++)
    bs[f] = th.active ? st[S.off_band + (f * C + th.c) * B + th.b] : 0.;
  Acc acc = {0, 0, 0, 0, 0, 0, 0, 0};
  if (th.acc_thread) load_acc(acc, st + S.off_acc + (th.c * kNumAcc + th.b) * kAccFields);
  ScanCounters cnt;
  load_counters(cnt, st, S);
  const BandConst k = load_band_const(T, th.bb, B);

  const unsigned total = n_frames[pair];
  const unsigned end = min(first_frame + n_chunk_frames, total);

  // The three band values a thread needs from the next frame's record travel through shared
  // memory, started one frame ahead with cp.async: every frame would otherwise begin with a DRAM
  // round trip (the records of a chunk are far larger than L2), and a prefetch into registers
  // costs registers this kernel does not have.  A thread reads back only what it copied itself,
  // so no barrier is involved.  The frame's scalars (flags, bandwidths, EHS, SNR sums) are
  // prefetched towards L1 / L2.
  __shared__ double stage[2][3][kGroup * kMaxChannels];
  auto stage_frame = [&](unsigned fl, int buf) {
    const double* r = records + ((size_t)pair * n_chunk_frames + fl) * L.stride;
    if (th.active) {
      cp_async8(&stage[buf][0][threadIdx.x], r + (0 * C + th.c) * B + th.b);
      cp_async8(&stage[buf][1][threadIdx.x], r + (1 * C + th.c) * B + th.b);
      cp_async8(&stage[buf][2][threadIdx.x], r + L.off_noise + th.c * B + th.b);
      if (th.b == 0) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(r + L.off_ehs));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(r + L.off_ints));
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (first_frame < end) stage_frame(0, 0);
  for (unsigned f = first_frame; f < end; f++) {
    const unsigned fl = f - first_frame;
    const int buf = fl & 1;
    const double* rec = records + ((size_t)pair * n_chunk_frames + fl) * L.stride;
    const int* rints = reinterpret_cast<const int*>(rec + L.off_ints);
    if (f + 1 < end) {
      stage_frame(fl + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");   // this frame's copies have landed
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    ScanInputs in;
    in.E2r = th.active ? stage[buf][0][threadIdx.x] : 1.;
    in.E2t = th.active ? stage[buf][1][threadIdx.x] : 1.;
    in.nz = th.active ? stage[buf][2][threadIdx.x] : 0.;
    in.flags = rints[0];
    in.bw = rints + 1 + 2 * th.c;
    in.ehs = rec + L.off_ehs + th.c;
    in.snr = rec + L.off_snr;
    in.dbg = kTap ? dbg + ((size_t)pair * n_chunk_frames + fl) * scan_tap_doubles(C, B) : nullptr;
#if defined(PEAQ_DEV_K2_OCC3)
    scan_step(in, MemConst{T, th.bb}, RegState{bs}, RegAcc{}, acc, cnt, sh, th, C, B);
#else
    scan_step(in, RegConst{k}, RegState{bs}, RegAcc{}, acc, cnt, sh, th, C, B);
#endif
  }

  // ---- store state ---------------------------------------------------------
  if (th.active) {
#pragma unroll
    for (int f = 0; f < kBandStateFields; f++) st[S.off_band + (f * C + th.c) * B + th.b] = bs[f];
  }
  if (th.acc_thread) store_acc(acc, st + S.off_acc + (th.c * kNumAcc + th.b) * kAccFields);
  if (threadIdx.x == 0) store_counters(cnt, st, S);
  scan_epilogue(T, acc, cnt, sh, th, C, results + pair);
}

}  // namespace

int scan_tap_doubles_per_frame(int C, int B) { return scan_tap_doubles(C, B); }

cudaError_t launch_scan_basic(const DeviceTables* d_tables, const double* records, RecordLayout L,
                              const unsigned* n_frames, unsigned first_frame,
                              unsigned n_chunk_frames, double* state, StateLayout S,
                              PairResult* results, int n_pairs, cudaStream_t stream, double* dbg) {
  if (n_pairs <= 0) return cudaSuccess;
  if (dbg)
    scan_basic_kernel<true><<<n_pairs, kGroup * L.C, 0, stream>>>(d_tables, records, L, n_frames, first_frame,
                                                                  n_chunk_frames, state, S, results, dbg);
  else
    scan_basic_kernel<false><<<n_pairs, kGroup * L.C, 0, stream>>>(d_tables, records, L, n_frames, first_frame,
                                                                   n_chunk_frames, state, S, results, nullptr);
  return cudaGetLastError();
}

}  // namespace peaq
