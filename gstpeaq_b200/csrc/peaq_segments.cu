// Long items as SEGMENTS (BASELINE configs[4]: few, hour-long pairs).
//
// Every recurrence of the model forgets: the slowest one (level / pattern adaptation at the lowest
// band, time constant 92 ms, leveladapter.c:243-340) decays by 5e-20 over 4.1 s, the DC-reject
// filter and the filter bank long before that.  A long item is therefore cut, at ABSOLUTE positions
// that depend on nothing but its length, into segments of kSegSamples samples; segment k > 0 is
// run as a pair of its own whose signals start kSegWarmSamples before the segment -- the warm-up
// frames only bring the recurrent state to what the sequential run has there (to the last bit or
// two) -- and whose sums restart at the segment's first frame (`acc_start` in the state's segment
// words).  All kernels of the engine then see a batch of many ~37 s pairs instead of a few long
// ones and fill the GPU the way they do for BASELINE configs[1..3]; the filter bank's recursion
// restarts every segment, which also bounds its rounding drift whatever the item length.
//
// What a segment cannot know is the item's history of one-way switches.  Segments k > 0 assume
//   A1  a frame above the threshold came before their warm-up (accumulators out of STATUS_INIT,
//       movaccum.c:317-354), and
//   A2  the loudness latch (gstpeaq.c:841-845) was set before their warm-up;
// segment 0 records when both happened, seg_combine_* checks it and raises the item's `redo` flag
// when either did not hold (an item that starts with half a minute of silence): the host then
// runs that item again as a whole.
//
// Combination (seg_combine_*): a segment leaves, per accumulator, its running sums `cur` and --
// when it ended in STATUS_TENTATIVE -- the sums `saved` at its last frame above the threshold.
// With T = sum of `cur` of the segments before s, the item's value is T + M(s) at the LAST
// segment s that owns a frame above the threshold (M = saved or cur): exactly the frames the
// sequential state machine would have committed.  Maxima (MFPD) combine with max instead of +.
// The combined sums go into segment 0's state block; the scan kernels are then launched once more
// with zero frames to evaluate MOVs, DI and ODG from it (their ordinary epilogue).
#include "peaq_engine.h"

#include <climits>

namespace peaq {
namespace {

enum { kStInit = 0, kStNormal = 1, kStTentative = 2 };
enum { kKindSum2, kKindSum2Max };   // fields (0, 1) are sums | ... and field 2 is a maximum

constexpr int kMaxSeg = 512;   // segments per item the combine kernels handle (> 5 hours of audio)

__global__ void seg_init_basic_kernel(double* state, StateLayout S, int n_vp, const int* __restrict__ seg_index,
                                      const unsigned* __restrict__ frame0, const unsigned* __restrict__ acc_start) {
  const int vp = blockIdx.x * blockDim.x + threadIdx.x;
  if (vp >= n_vp || seg_index[vp] == 0) return;
  int* ints = reinterpret_cast<int*>(state + (size_t)vp * S.stride + S.off_ints);
  ints[0] = kStNormal;            // A1
  ints[1] = (int)frame0[vp];      // the frame counter is the item's
  ints[2] = 0;                    // A2: loudness reached at frame 0
  ints[kSegAccStart] = (int)acc_start[vp];
}

__global__ void seg_init_adv_kernel(double* state, AdvStateLayout S, int n_vp, const int* __restrict__ seg_index,
                                    const unsigned* __restrict__ frame0_fft, const unsigned* __restrict__ acc_start_fft,
                                    const unsigned* __restrict__ frame0_fb, const unsigned* __restrict__ acc_start_fb) {
  const int vp = blockIdx.x * blockDim.x + threadIdx.x;
  if (vp >= n_vp || seg_index[vp] == 0) return;
  int* ints = reinterpret_cast<int*>(state + (size_t)vp * S.stride + S.off_ints);
  ints[0] = kStNormal;
  ints[1] = (int)frame0_fft[vp];
  ints[2] = kStNormal;
  ints[3] = (int)frame0_fb[vp];
  ints[4] = 0;
  ints[kASegAccStartFft] = (int)acc_start_fft[vp];
  ints[kASegAccStartFb] = (int)acc_start_fb[vp];
}

// one CTA per item; first_vp[item] .. first_vp[item] + n_seg[item] - 1 are its segments
__global__ void __launch_bounds__(64)
seg_combine_basic_kernel(double* state, StateLayout S, int n_items, const int* __restrict__ first_vp,
                         const int* __restrict__ n_seg_of, const unsigned* __restrict__ frame0,
                         unsigned char* __restrict__ redo) {
  const int item = blockIdx.x;
  const int n_seg = n_seg_of[item] < kMaxSeg ? n_seg_of[item] : kMaxSeg;
  if (n_seg <= 1) {
    if (threadIdx.x == 0) redo[item] = 0;
    return;
  }
  __shared__ int status[kMaxSeg], owned[kMaxSeg];
  const int vp0 = first_vp[item];
  for (int s = threadIdx.x; s < n_seg; s += blockDim.x) {
    const int* ints = reinterpret_cast<const int*>(state + (size_t)(vp0 + s) * S.stride + S.off_ints);
    status[s] = ints[0];
    owned[s] = ints[kSegOwnedAbove];
  }
  __syncthreads();
  const int C = S.C;
  double out[kAccFields];
  const bool acc_thread = (int)threadIdx.x < C * kNumAcc;
  if (acc_thread) {
    const int slot = threadIdx.x % kNumAcc;
    const int kind = slot == 9 /* MFPD: filtered maximum */ ? kKindSum2Max : kKindSum2;
    double T[3] = {0., 0., 0.}, Cm[3] = {0., 0., 0.};
    for (int s = 0; s < n_seg; s++) {
      const double* a = state + (size_t)(vp0 + s) * S.stride + S.off_acc + threadIdx.x * kAccFields;
      const bool tent = status[s] == kStTentative;
      for (int f = 0; f < 3; f++) {
        const bool is_max = kind == kKindSum2Max && f == 2;
        if (f == 2 && !is_max) continue;
        const double cur = a[f], m = tent ? a[5 + f] : cur;
        if (owned[s]) Cm[f] = is_max ? (m > T[f] ? m : T[f]) : T[f] + m;
        T[f] = is_max ? (cur > T[f] ? cur : T[f]) : T[f] + cur;
      }
    }
    out[0] = Cm[0];
    out[1] = Cm[1];
    out[2] = Cm[2];
  }
  double sig = 0., noise = 0.;
  int any_owned = 0;
  if (threadIdx.x == 63) {
    for (int s = 0; s < n_seg; s++) {
      const double* st = state + (size_t)(vp0 + s) * S.stride;
      sig += st[S.off_scalar];
      noise += st[S.off_scalar + 1];
      any_owned |= owned[s];
    }
  }
  __syncthreads();
  double* st0 = state + (size_t)vp0 * S.stride;
  if (acc_thread) {
    double* a = st0 + S.off_acc + threadIdx.x * kAccFields;
    const int slot = threadIdx.x % kNumAcc;
    a[0] = out[0];
    a[1] = out[1];
    if (slot == 9) a[2] = out[2];
    a[5] = a[6] = a[7] = 0.;
  }
  if (threadIdx.x == 63) {
    int* ints0 = reinterpret_cast<int*>(st0 + S.off_ints);
    const int* ints_last = reinterpret_cast<const int*>(state + (size_t)(vp0 + n_seg - 1) * S.stride + S.off_ints);
    const unsigned warm1 = frame0[vp0 + 1];   // first warm-up frame of segment 1
    const unsigned first_above = (unsigned)ints0[kSegFirstAbove], loud = (unsigned)ints0[2];
    redo[item] = (first_above == 0 || first_above - 1 >= warm1 || loud == UINT_MAX || loud >= warm1) ? 1 : 0;
    ints0[0] = any_owned ? kStNormal : kStInit;
    ints0[1] = ints_last[1];   // frames of the whole item
    st0[S.off_scalar] = sig;
    st0[S.off_scalar + 1] = noise;
  }
}

__global__ void __launch_bounds__(64)
seg_combine_adv_kernel(double* state, AdvStateLayout S, int n_items, const int* __restrict__ first_vp,
                       const int* __restrict__ n_seg_of, const unsigned* __restrict__ frame0_fft,
                       const unsigned* __restrict__ frame0_fb, unsigned char* __restrict__ redo) {
  const int item = blockIdx.x;
  const int n_seg = n_seg_of[item] < kMaxSeg ? n_seg_of[item] : kMaxSeg;
  if (n_seg <= 1) {
    if (threadIdx.x == 0) redo[item] = 0;
    return;
  }
  __shared__ int status[2][kMaxSeg], owned[2][kMaxSeg];   // [fft | fb clock]
  const int vp0 = first_vp[item];
  for (int s = threadIdx.x; s < n_seg; s += blockDim.x) {
    const int* ints = reinterpret_cast<const int*>(state + (size_t)(vp0 + s) * S.stride + S.off_ints);
    status[0][s] = ints[0];
    status[1][s] = ints[2];
    owned[0][s] = ints[kASegOwnedAboveFft];
    owned[1][s] = ints[kASegOwnedAboveFb];
  }
  __syncthreads();
  const int C = S.C;
  // threads 0 .. 2C-1: FFT-clock accumulators [c][2]; 2C .. 5C-1: filter-bank clock [c][3]
  const int n_fft = 2 * C, n_fb = 3 * C;
  const bool acc_thread = (int)threadIdx.x < n_fft + n_fb;
  const int clock = (int)threadIdx.x < n_fft ? 0 : 1;
  const int idx = clock ? threadIdx.x - n_fft : threadIdx.x;
  const int off = clock ? S.off_fb_acc + idx * kAccFields : S.off_fft_acc + idx * kAccFields;
  double out[3] = {0., 0., 0.};
  if (acc_thread) {
    double T[3] = {0., 0., 0.};
    for (int s = 0; s < n_seg; s++) {
      const double* a = state + (size_t)(vp0 + s) * S.stride + off;
      const bool tent = status[clock][s] == kStTentative;
      for (int f = 0; f < 3; f++) {
        const double cur = a[f], m = tent ? a[5 + f] : cur;
        if (owned[clock][s]) out[f] = T[f] + m;
        T[f] += cur;
      }
    }
  }
  double sig = 0., noise = 0.;
  int any_owned[2] = {0, 0};
  if (threadIdx.x == 63) {
    for (int s = 0; s < n_seg; s++) {
      const double* st = state + (size_t)(vp0 + s) * S.stride;
      sig += st[S.off_fft_scalar];
      noise += st[S.off_fft_scalar + 1];
      any_owned[0] |= owned[0][s];
      any_owned[1] |= owned[1][s];
    }
  }
  __syncthreads();
  double* st0 = state + (size_t)vp0 * S.stride;
  if (acc_thread) {
    double* a = st0 + off;
    a[0] = out[0];
    a[1] = out[1];
    a[2] = out[2];
    a[5] = a[6] = a[7] = 0.;
  }
  if (threadIdx.x == 63) {
    int* ints0 = reinterpret_cast<int*>(st0 + S.off_ints);
    const int* ints_last = reinterpret_cast<const int*>(state + (size_t)(vp0 + n_seg - 1) * S.stride + S.off_ints);
    const unsigned warm_fft = frame0_fft[vp0 + 1], warm_fb = frame0_fb[vp0 + 1];
    const unsigned fa_fft = (unsigned)ints0[kASegFirstAboveFft], fa_fb = (unsigned)ints0[kASegFirstAboveFb];
    const unsigned loud = (unsigned)ints0[4];
    redo[item] = (fa_fft == 0 || fa_fft - 1 >= warm_fft || fa_fb == 0 || fa_fb - 1 >= warm_fb || loud == UINT_MAX ||
                  loud >= warm_fb)
                     ? 1
                     : 0;
    ints0[0] = any_owned[0] ? kStNormal : kStInit;
    ints0[1] = ints_last[1];
    ints0[2] = any_owned[1] ? kStNormal : kStInit;
    ints0[3] = ints_last[3];
    st0[S.off_fft_scalar] = sig;
    st0[S.off_fft_scalar + 1] = noise;
  }
}

__global__ void seg_gather_results_kernel(const PairResult* __restrict__ res, const int* __restrict__ first_vp,
                                          int n_items, PairResult* __restrict__ out) {
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item < n_items) out[item] = res[first_vp[item]];
}

}  // namespace

cudaError_t launch_seg_init(double* state, const StateLayout* S, const AdvStateLayout* A, int n_vp,
                            const SegTable& t, cudaStream_t stream) {
  if (n_vp <= 0) return cudaSuccess;
  const int grid = (n_vp + 127) / 128;
  if (A) {
    seg_init_adv_kernel<<<grid, 128, 0, stream>>>(state, *A, n_vp, t.seg_index, t.frame0_fft, t.acc_start_fft,
                                                  t.frame0_fb, t.acc_start_fb);
  } else {
    seg_init_basic_kernel<<<grid, 128, 0, stream>>>(state, *S, n_vp, t.seg_index, t.frame0_fft, t.acc_start_fft);
  }
  return cudaGetLastError();
}

cudaError_t launch_seg_combine(double* state, const StateLayout* S, const AdvStateLayout* A, int n_items,
                               const SegTable& t, unsigned char* redo, cudaStream_t stream) {
  if (n_items <= 0) return cudaSuccess;
  if (A) {
    seg_combine_adv_kernel<<<n_items, 64, 0, stream>>>(state, *A, n_items, t.first_vp, t.n_seg, t.frame0_fft,
                                                       t.frame0_fb, redo);
  } else {
    seg_combine_basic_kernel<<<n_items, 64, 0, stream>>>(state, *S, n_items, t.first_vp, t.n_seg, t.frame0_fft, redo);
  }
  return cudaGetLastError();
}

cudaError_t launch_seg_gather_results(const PairResult* res, const int* first_vp, int n_items, PairResult* out,
                                      cudaStream_t stream) {
  if (n_items <= 0) return cudaSuccess;
  seg_gather_results_kernel<<<(n_items + 127) / 128, 128, 0, stream>>>(res, first_vp, n_items, out);
  return cudaGetLastError();
}

}  // namespace peaq
