// `peaq_fused_basic_kernel`: the whole basic-mode hot path of one (ref,test) pair in ONE persistent
// CTA -- the shape of the reference's own frame loop (gstpeaq.c:850-921: ear model, level adapter,
// modulation and every MOV of a frame back to back on hot data), batched over thousands of pairs.
//
// One CTA = one pair, 4 C warps, looping over the pair's FFT-clock frames.  Per frame
//   1. frame_body<true> (peaq_frames.cuh): PCM (TMA bulk copy, issued one frame ahead) -> window
//      -> FFT -> ear weighting -> band grouping -> spreading, noise in bands, bandwidth, EHS;
//      the results stay in shared memory;
//   2. scan_step (peaq_scan.cuh) on the same threads, thread (c, b) = band b of channel c: time
//      smearing, level / pattern adaptation, modulation, MOV terms, the 11 accumulators.
// Nothing per-frame goes through HBM: the kernel's DRAM traffic is the PCM it reads (16 KB per
// frame; the 50 % frame overlap is re-read from L2).  The recurrent band state (14 doubles per
// band and channel) lives in the pair's state block in global memory between two frames of the
// same CTA -- 24 KB per pair that never leave L2 while the CTA runs -- because the frame half
// needs every register of the 80 that three resident CTAs per SM allow.  The state block is the
// one K2 uses, so a run can be continued by either path.
//
// Arithmetic and its order are those of K1 + K2: results are bit-identical to the two-kernel path.
#include "peaq_frames.cuh"
#include "peaq_scan.cuh"

namespace peaq {
namespace {

__global__ void __launch_bounds__(256, 3)
peaq_fused_basic_kernel(const DeviceTables* __restrict__ T, PcmView pcm, unsigned first_frame,
                        unsigned n_chunk_frames, double* __restrict__ state, StateLayout S,
                        PairResult* __restrict__ results, RecordLayout L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  const int C = pcm.channels, B = S.B;
  const int pair = blockIdx.x;
  const ScanThread th = make_scan_thread(B);
  FrameMail* mail = frame_mail(smem, C);
  // The scan's work arrays alias FFT buffers that are idle during the scan step: stereo -- stream 2's
  // (streams 0 and 1 receive the next frame's PCM meanwhile); mono -- stream 1's (no prefetch then)
  ScanShared& sh = *reinterpret_cast<ScanShared*>(frame_stream_buf(smem, C == 2 ? 2 : 1));
  static_assert(sizeof(ScanShared) <= kWorkDoubles * sizeof(double), "scan arrays must fit one FFT buffer");
  const bool prefetch = C == 2;

  double* st = state + (size_t)pair * S.stride;
  // status word and counters: in shared memory between frames (the frame half needs the registers)
  ScanCounters* cnt_sh = reinterpret_cast<ScanCounters*>(mail + 1);
  if (threadIdx.x == 0) load_counters(*cnt_sh, st, S);
  const unsigned total = pcm.n_frames[pair];
  const unsigned end = min(first_frame + n_chunk_frames, total);

  frame_load_twiddles(T, smem);
  if (threadIdx.x == 0) mbar_init(&mail->mbar, 1);
  __syncthreads();
  unsigned parity = 0;
  bool tma_ok = first_frame < end && frame_tma_ok(pcm, pair, first_frame);
  if (tma_ok && threadIdx.x == 0) frame_tma_issue(pcm, pair, first_frame, smem);

  for (unsigned f = first_frame; f < end; f++) {
    // ---- frame half: results into shared memory ------------------------------------------
    frame_body<true>(T, pcm, pair, f, B, 0, smem, tma_ok, parity, nullptr, L);
    if (tma_ok) parity ^= 1u;
    __syncthreads();
    // ---- scan half ---------------------------------------------------------------------------
    ScanInputs in;
    in.E2r = th.active ? frame_out_e2(smem, th.c, 0)[th.b] : 1.;
    in.E2t = th.active ? frame_out_e2(smem, th.c, 1)[th.b] : 1.;
    in.nz = th.active ? frame_out_noise(smem, th.c)[th.b] : 0.;
    in.flags = mail->o_flags;
    in.bw = mail->o_bw[th.c];
    in.ehs = mail->o_ehs + th.c;
    in.snr = mail->o_snr;
    in.dbg = nullptr;
    ScanCounters cnt = *cnt_sh;
    __syncthreads();   // every input is in registers: the FFT buffers are free
    const bool next_ok = f + 1 < end && frame_tma_ok(pcm, pair, f + 1);
    if (prefetch && next_ok && threadIdx.x == 0) frame_tma_issue(pcm, pair, f + 1, smem);

    Acc acc = {0, 0, 0, 0, 0, 0, 0, 0};
    scan_step(in, MemConst{T, th.bb}, MemState{st + S.off_band + th.c * B + th.b, C * B, th.active},
              MemAcc{st + S.off_acc + (th.c * kNumAcc + th.b) * kAccFields}, acc, cnt, sh, th, C, B);
    if (threadIdx.x == 0) *cnt_sh = cnt;
    __syncthreads();   // scan arrays dead before the next frame reuses the buffers
    if (!prefetch && next_ok && threadIdx.x == 0) frame_tma_issue(pcm, pair, f + 1, smem);
    tma_ok = next_ok;
  }

  const ScanCounters cnt = *cnt_sh;
  if (threadIdx.x == 0) store_counters(cnt, st, S);
  Acc acc = {0, 0, 0, 0, 0, 0, 0, 0};
  if (th.acc_thread) load_acc(acc, st + S.off_acc + (th.c * kNumAcc + th.b) * kAccFields);
  scan_epilogue(T, acc, cnt, sh, th, C, results + pair);
}

}  // namespace

cudaError_t launch_fused_basic(const DeviceTables* d_tables, PcmView pcm, int n_pairs, unsigned first_frame,
                               unsigned n_chunk_frames, double* state, StateLayout S, PairResult* results,
                               cudaStream_t stream) {
  if (n_pairs <= 0) return cudaSuccess;
  const size_t smem = frame_smem_bytes(pcm.channels) + sizeof(ScanCounters);
  cudaError_t e = cudaFuncSetAttribute(peaq_fused_basic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(frame_smem_bytes(kMaxChannels) + sizeof(ScanCounters)));
  if (e != cudaSuccess) return e;
  const RecordLayout L = make_record_layout(S.C, S.B);
  peaq_fused_basic_kernel<<<n_pairs, 128 * pcm.channels, smem, stream>>>(d_tables, pcm, first_frame, n_chunk_frames,
                                                                        state, S, results, L);
  return cudaGetLastError();
}

}  // namespace peaq
