// K1 `fft_frames`: the stateless, frame-parallel half of the PEAQ hot path as a kernel of its own
// -- one CTA per (pair, frame), results to per-frame records in HBM for the scan kernels.  This is
// the default path in both modes (basic: followed by K2 scan_basic_kernel; advanced: the FFT clock
// next to the filter bank); the fused persistent kernel (peaq_fused.cu, PEAQ_B200_FUSED=1) runs the
// same device code, frame_body in peaq_frames.cuh, inside a loop over a pair's frames.
#include "peaq_frames.cuh"

namespace peaq {
namespace {

__global__ void __launch_bounds__(256, 3)
fft_frames_kernel(const DeviceTables* __restrict__ T, PcmView pcm, unsigned first_frame,
                  unsigned n_chunk_frames, double* __restrict__ records, RecordLayout L, int B,
                  int advanced) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int pair = blockIdx.x / n_chunk_frames;
  const unsigned chunk_frame = blockIdx.x - pair * n_chunk_frames;
  const unsigned frame = first_frame + chunk_frame;
  if (frame >= pcm.n_frames[pair]) return;
  double* smem = reinterpret_cast<double*>(smem_raw);

  frame_load_twiddles(T, smem);
  const bool tma_ok = frame_tma_ok(pcm, pair, frame);
  if (tma_ok) {
    if (threadIdx.x == 0) mbar_init(&frame_mail(smem, pcm.channels)->mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) frame_tma_issue(pcm, pair, frame, smem);
  }
  double* rec = records + ((size_t)pair * n_chunk_frames + chunk_frame) * L.stride;
  frame_body<false>(T, pcm, pair, frame, B, advanced, smem, tma_ok, 0u, rec, L);
}

}  // namespace

size_t fft_frames_smem_bytes(int channels) { return frame_smem_bytes(channels); }

cudaError_t launch_fft_frames(const DeviceTables* d_tables, PcmView pcm, int n_pairs,
                              unsigned first_frame, unsigned n_chunk_frames, double* records,
                              RecordLayout L, int fft_bands, bool advanced, cudaStream_t stream) {
  if (n_pairs <= 0 || n_chunk_frames == 0) return cudaSuccess;
  const size_t smem = fft_frames_smem_bytes(pcm.channels);
  // per device (function attributes belong to the context), a few hundred ns: set every time
  cudaError_t e = cudaFuncSetAttribute(fft_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)fft_frames_smem_bytes(kMaxChannels));
  if (e != cudaSuccess) return e;
  dim3 grid((unsigned)n_chunk_frames * (unsigned)n_pairs);
  dim3 block(128 * pcm.channels);
  fft_frames_kernel<<<grid, block, smem, stream>>>(d_tables, pcm, first_frame, n_chunk_frames, records,
                                                   L, fft_bands, advanced ? 1 : 0);
  return cudaGetLastError();
}

}  // namespace peaq
