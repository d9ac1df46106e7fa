// K2 `scan_basic`: the recurrent half of basic-mode PEAQ as a kernel of its own -- one CTA per
// pair walks the pair's frames in order over the per-frame records K1 wrote, carrying every
// recurrence in registers across the frame loop; the state is loaded from / stored to global
// memory at chunk boundaries, so hour-long items and streaming sessions run as a sequence of
// chunks.  Large batches run the fused persistent kernel (peaq_fused.cu) instead; the device code
// (scan_step, scan_epilogue) is shared, peaq_scan.cuh.
#include "peaq_scan.cuh"

namespace peaq {
namespace {

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

  for (unsigned f = first_frame; f < end; f++) {
    const double* rec = records + ((size_t)pair * n_chunk_frames + (f - first_frame)) * L.stride;
    const int* rints = reinterpret_cast<const int*>(rec + L.off_ints);
    ScanInputs in;
    in.E2r = th.active ? rec[(0 * C + th.c) * B + th.b] : 1.;
    in.E2t = th.active ? rec[(1 * C + th.c) * B + th.b] : 1.;
    in.nz = th.active ? rec[L.off_noise + th.c * B + th.b] : 0.;
    in.flags = rints[0];
    in.bw = rints + 1 + 2 * th.c;
    in.ehs = rec + L.off_ehs + th.c;
    in.snr = rec + L.off_snr;
    in.dbg = kTap ? dbg + ((size_t)pair * n_chunk_frames + (f - first_frame)) * scan_tap_doubles(C, B) : nullptr;
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
