// C ABI (include/peaq_b200.h) on top of the CUDA kernels: engine (batch path),
// session (streaming path mirroring one `peaq` element instance), synthetic
// input generator and small device-memory helpers for FFI hosts.
#include "../../include/peaq_b200.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "peaq_engine.h"
#include "peaq_synth.h"

namespace peaq {

static thread_local std::string g_last_error;

static int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define PEAQ_CUDA(expr)                                                                 \
  do {                                                                                  \
    cudaError_t err__ = (expr);                                                         \
    if (err__ != cudaSuccess)                                                           \
      return fail(PEAQ_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err__)); \
  } while (0)

static_assert(sizeof(PairResult) == sizeof(peaq_b200_result), "result layout mismatch");

// ---------------------------------------------------------------------------
// kernels that live here: state initialisation and the synthetic generator

// Fresh recurrent state: everything zero (g_new0 in the reference's
// constructors), accumulators in STATUS_INIT, the windowed-average history
// primed with NaN (movaccum.c:293), loudness not yet reached (gstpeaq.c:359).
// Small tables (plans, segment words) are FETCHED from pinned host memory by a kernel on the
// compute stream instead of being copied: a cudaMemcpyAsync would share the copy engine with the
// staging copies of the next sub-batch and wait behind gigabytes of PCM.
__global__ void fetch_words_kernel(unsigned* __restrict__ dst, const unsigned* __restrict__ src, size_t n_words) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

__global__ void init_state_kernel(double* state, StateLayout S, int n_pairs) {
  const size_t total = (size_t)n_pairs * S.stride;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int pair = (int)(i / S.stride);
    const int o = (int)(i - (size_t)pair * S.stride);
    double v = 0.;
    if (o >= S.off_acc && o < S.off_scalar) {
      const int a = (o - S.off_acc) % (kNumAcc * kAccFields);
      const int slot = a / kAccFields, field = a % kAccFields;
      if (slot == 3 /* WinModDiff */ && field >= 2 && field <= 4) v = nan("");
    }
    if (o == S.off_ints + 1) {
      // int32 slots 2,3: loudness_reached_frame = G_MAXUINT (gstpeaq.c:359), spare
      int* ints = reinterpret_cast<int*>(state + i);
      ints[0] = (int)UINT_MAX;
      ints[1] = 0;
    } else {
      state[i] = v;   // all-zero bits: STATUS_INIT, counters 0
    }
  }
}

__global__ void synth_pairs_kernel(float* __restrict__ ref, float* __restrict__ test,
                                   size_t pair_stride, unsigned long long first_pair,
                                   unsigned long long n_samples, int channels) {
  __shared__ int16_t table[PEAQ_SYNTH_TABLE_SIZE];
  __shared__ PeaqSynthPair sp;
  for (int i = threadIdx.x; i < PEAQ_SYNTH_TABLE_SIZE; i += blockDim.x)
    table[i] = (int16_t)peaq_synth_sine_entry((uint32_t)i);
  if (threadIdx.x == 0) peaq_synth_pair_init(&sp, first_pair + blockIdx.y);
  __syncthreads();
  float* r = ref + (size_t)blockIdx.y * pair_stride;
  float* t = test + (size_t)blockIdx.y * pair_stride;
  const unsigned long long total = n_samples * (unsigned long long)channels;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long n = i / channels;
    const int c = (int)(i - n * channels);
    int32_t rv, tv;
    peaq_synth_sample(&sp, table, n, c, &rv, &tv);
    r[i] = (float)rv / 32768.0f;
    t[i] = (float)tv / 32768.0f;
  }
}

cudaError_t launch_synth_pairs(float* ref, float* test, size_t pair_stride, int n_pairs,
                               unsigned long long first_pair_index, unsigned long long n_samples,
                               int channels, cudaStream_t stream) {
  if (n_pairs <= 0 || n_samples == 0) return cudaSuccess;
  const unsigned long long total = n_samples * channels;
  unsigned bx = (unsigned)std::min<unsigned long long>((total + 255) / 256, 1184);
  for (int p0 = 0; p0 < n_pairs; p0 += 65535) {
    const int np = std::min(65535, n_pairs - p0);
    dim3 grid(bx, np);
    synth_pairs_kernel<<<grid, 256, 0, stream>>>(ref + (size_t)p0 * pair_stride,
                                                 test + (size_t)p0 * pair_stride, pair_stride,
                                                 first_pair_index + p0, n_samples, channels);
  }
  return cudaGetLastError();
}

// number of FFT-clock frames for n samples per channel: do_processing
// (gstpeaq.c:596-611) consumes 2048-sample frames every 1024 samples, do_flush
// (:716-745) adds one zero-padded frame if anything is left
static unsigned frames_for_samples(uint64_t n) {
  uint64_t full = n >= kFftFrame ? (n - kFftFrame) / kFftStep + 1 : 0;
  uint64_t left = n - full * kFftStep;
  return (unsigned)(full + (left > 0 ? 1 : 0));
}

// 192-sample frames of the filter-bank clock: whole frames, plus one
// zero-padded frame for the remainder (gstpeaq.c:649-652, :770-772)
static unsigned fb_frames_for_samples(uint64_t n) {
  return (unsigned)(n / kFbFrame + (n % kFbFrame ? 1 : 0));
}

// ---------------------------------------------------------------------------

struct EventPair {
  cudaEvent_t a, b;
  int which;
};

struct Engine {
  int device = 0;
  bool advanced = false;
  double level = 92.0;
  DeviceTables* h_tables = nullptr;
  DeviceTables* d_tables = nullptr;
  cudaStream_t stream = nullptr;
  double* d_records = nullptr;
  size_t records_cap = 0;   // doubles
  double* d_state = nullptr;
  size_t state_cap = 0;
  PairResult* d_results = nullptr;
  unsigned long long* d_nsamples = nullptr;       // [4][pairs_cap]: fft ref, fft test, fb ref, fb test
  unsigned* d_nframes = nullptr;                  // [2][pairs_cap]: fft clock, fb clock
  size_t pairs_cap = 0;
  // advanced mode workspaces
  double* d_hp = nullptr;
  size_t hp_cap = 0;
  double* d_hp_state = nullptr;
  size_t hp_state_cap = 0;
  // whole-item mode of the DC-reject scan: its own high-priority stream, so that the (chain
  // bound, few-thread) scan of the whole batch hides underneath the frame kernel
  cudaStream_t hp_stream = nullptr;
  cudaEvent_t ev_hp_ready = nullptr, ev_hp_done = nullptr;
  size_t hp_whole_budget_bytes = (size_t)72 << 30;
  double* d_fbout = nullptr;
  size_t fbout_cap = 0;
  double* d_fbenergy = nullptr;   // rectified sub-step energies [stream][sub-step][band]
  size_t fbenergy_cap = 0;
  unsigned char* d_fbflags = nullptr;
  size_t fbflags_cap = 0;
  double* d_fbdbg = nullptr;
  size_t fbdbg_cap = 0;
  size_t last_fbdbg_doubles = 0;
  unsigned last_fb_frames = 0;
  size_t fb_budget_bytes = (size_t)32 << 30;
  float* d_stage[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [slot][ref|test]
  size_t stage_cap[2] = {0, 0};   // floats, per buffer of the slot
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_freed[2] = {nullptr, nullptr};
  // Small host <-> device transfers of a batch (plans, segment tables, redo flags) go through
  // PINNED memory: from pageable memory cudaMemcpyAsync first drains the stream (H2D) or waits for
  // the copy (D2H), which would serialise the staging copies of the next sub-batch behind the
  // kernels of this one.  The arena is recycled when the batch has been synchronised.
  struct PinnedArena {
    std::vector<std::pair<char*, size_t>> blocks;
    size_t cur = 0, used = 0;
    void* take(size_t bytes) {
      bytes = (bytes + 63) & ~(size_t)63;
      while (cur < blocks.size() && used + bytes > blocks[cur].second) {
        cur++;
        used = 0;
      }
      if (cur == blocks.size()) {
        const size_t cap = std::max<size_t>(bytes, (size_t)1 << 20);
        char* p = nullptr;
        if (cudaHostAlloc((void**)&p, cap, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {   // kernels read / write it
          cudaGetLastError();
          return nullptr;
        }
        blocks.emplace_back(p, cap);
        used = 0;
      }
      void* r = blocks[cur].first + used;
      used += bytes;
      return r;
    }
    void reset() { cur = 0; used = 0; }
    void destroy() {
      for (auto& b : blocks) cudaFreeHost(b.first);
      blocks.clear();
      reset();
    }
  } arena;
  struct DeferredCopy {   // arena -> caller memory, once the batch has been synchronised
    void* dst;
    const void* src;
    size_t bytes;
  };
  std::vector<DeferredCopy> deferred;
  // device <- pinned host memory through fetch_words_kernel, on the compute stream
  template <typename T>
  int fetch(T* d_dst, const T* pinned_src, size_t n) {
    static_assert(sizeof(T) % 4 == 0, "whole words");
    const size_t words = n * (sizeof(T) / 4);
    if (!words) return 0;
    const unsigned blocks = (unsigned)std::min<size_t>((words + 255) / 256, 64);
    fetch_words_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<unsigned*>(d_dst),
                                                   reinterpret_cast<const unsigned*>(pinned_src), words);
    PEAQ_CUDA(cudaGetLastError());
    return 0;
  }
  // pinned copy of a small host array (alive until finish_batch)
  template <typename T>
  const T* pinned_copy(const T* src, size_t n) {
    T* p = static_cast<T*>(arena.take(std::max<size_t>(n, 1) * sizeof(T)));
    if (p && n) std::memcpy(p, src, n * sizeof(T));
    return p;
  }
  std::vector<EventPair> events;
  size_t events_used = 0;
  double ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // see peaq_b200_engine_last_ms; [7]: flags + DC-reject scan
  uint64_t launches = 0;
  bool keep_records = false;
  RecordLayout last_layout = {};
  size_t last_records_doubles = 0;
  size_t record_budget_bytes = (size_t)16 << 30;
  int sm_count = 148;
  int fused_mode = 0;    // PEAQ_B200_FUSED=1: the fused persistent kernel for basic-mode batches
  unsigned long long fb_pos = 0;   // filter-bank clock: samples consumed so far by the running item (session)
  int pipeline_mode = -1;   // PEAQ_B200_PIPELINE: 1 always, 0 never, default: batches of few pairs
  cudaStream_t scan_stream = nullptr;
  cudaEvent_t ev_rec_ready[2] = {nullptr, nullptr}, ev_rec_free[2] = {nullptr, nullptr};
  // segments of long items (peaq_segments.cu): device copies of the segment table
  int segment_mode = 1;   // PEAQ_B200_SEGMENTS=0: never cut items into segments
  unsigned long long* d_seg_base = nullptr;
  size_t seg_base_cap = 0;
  unsigned* d_seg_words = nullptr;   // [5][n_vp]: segment index, frame0 / acc_start of both clocks; then [2][n_items]
  size_t seg_words_cap = 0;
  PairResult* d_item_results = nullptr;
  size_t item_results_cap = 0;

  int init() {
    PEAQ_CUDA(cudaSetDevice(device));
    PEAQ_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    PEAQ_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    {
      // scratch of the stream-ordered allocator (the DC-reject block scan's block states) stays
      // with the process between batches instead of going back to the driver at every sync
      cudaMemPool_t pool = nullptr;
      if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      cudaGetLastError();
    }
    {
      int prio_lo = 0, prio_hi = 0;
      PEAQ_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
      PEAQ_CUDA(cudaStreamCreateWithPriority(&hp_stream, cudaStreamNonBlocking, prio_hi));
      PEAQ_CUDA(cudaEventCreateWithFlags(&ev_hp_ready, cudaEventDisableTiming));
      PEAQ_CUDA(cudaEventCreateWithFlags(&ev_hp_done, cudaEventDisableTiming));
    }
    for (int i = 0; i < 2; i++) {
      PEAQ_CUDA(cudaEventCreateWithFlags(&ev_copied[i], cudaEventDisableTiming));
      PEAQ_CUDA(cudaEventCreateWithFlags(&ev_freed[i], cudaEventDisableTiming));
    }
    h_tables = new (std::nothrow) DeviceTables;
    if (!h_tables) return fail(PEAQ_B200_ERR_NOMEM, "out of host memory");
    build_tables(h_tables, advanced, level);
    PEAQ_CUDA(cudaMalloc(&d_tables, sizeof(DeviceTables)));
    PEAQ_CUDA(cudaMemcpy(d_tables, h_tables, sizeof(DeviceTables), cudaMemcpyHostToDevice));
    PEAQ_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    if (const char* env = std::getenv("PEAQ_B200_FUSED")) fused_mode = std::atoi(env) ? 1 : 0;
    if (const char* env = std::getenv("PEAQ_B200_PIPELINE")) pipeline_mode = std::atoi(env) ? 1 : 0;
    if (const char* env = std::getenv("PEAQ_B200_SEGMENTS")) segment_mode = std::atoi(env) ? 1 : 0;
    if (const char* env = std::getenv("PEAQ_B200_RECORD_BUDGET_MB")) {
      const long mb = std::atol(env);
      if (mb > 0) record_budget_bytes = (size_t)mb << 20;
    }
    if (const char* env = std::getenv("PEAQ_B200_FB_BUDGET_MB")) {
      const long mb = std::atol(env);
      if (mb > 0) fb_budget_bytes = (size_t)mb << 20;
    }
    if (const char* env = std::getenv("PEAQ_B200_HP_WHOLE_BUDGET_MB")) {
      const long mb = std::atol(env);
      if (mb >= 0) hp_whole_budget_bytes = (size_t)mb << 20;
    }
    return 0;
  }

  // property `playback_level` on a live model (gstpeaq.c:509-514 sets it on both ear models at any
  // time, without touching their state): rebuild the tables, keep everything else
  int set_level(double new_level) {
    PEAQ_CUDA(cudaSetDevice(device));
    PEAQ_CUDA(cudaStreamSynchronize(stream));
    level = new_level;
    build_tables(h_tables, advanced, level);
    PEAQ_CUDA(cudaMemcpy(d_tables, h_tables, sizeof(DeviceTables), cudaMemcpyHostToDevice));
    return 0;
  }

  void destroy() {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    if (hp_stream) cudaStreamSynchronize(hp_stream);
    if (scan_stream) {
      cudaStreamSynchronize(scan_stream);
      for (int i = 0; i < 2; i++) {
        cudaEventDestroy(ev_rec_ready[i]);
        cudaEventDestroy(ev_rec_free[i]);
      }
      cudaStreamDestroy(scan_stream);
    }
    for (auto& e : events) {
      cudaEventDestroy(e.a);
      cudaEventDestroy(e.b);
    }
    cudaFree(d_records);
    cudaFree(d_state);
    cudaFree(d_results);
    arena.destroy();
    cudaFree(d_seg_base);
    cudaFree(d_seg_words);
    cudaFree(d_item_results);
    cudaFree(d_nsamples);
    cudaFree(d_nframes);
    for (int i = 0; i < 2; i++) {
      cudaFree(d_stage[i][0]);
      cudaFree(d_stage[i][1]);
      if (ev_copied[i]) cudaEventDestroy(ev_copied[i]);
      if (ev_freed[i]) cudaEventDestroy(ev_freed[i]);
    }
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (ev_hp_ready) cudaEventDestroy(ev_hp_ready);
    if (ev_hp_done) cudaEventDestroy(ev_hp_done);
    if (hp_stream) cudaStreamDestroy(hp_stream);
    cudaFree(d_hp);
    cudaFree(d_hp_state);
    cudaFree(d_fbout);
    cudaFree(d_fbenergy);
    cudaFree(d_fbflags);
    cudaFree(d_fbdbg);
    cudaFree(d_tables);
    if (stream) cudaStreamDestroy(stream);
    delete h_tables;
  }

  int timer_begin(int which, cudaStream_t on = nullptr) {
    if (events_used == events.size()) {
      EventPair p;
      p.which = which;
      PEAQ_CUDA(cudaEventCreate(&p.a));
      PEAQ_CUDA(cudaEventCreate(&p.b));
      events.push_back(p);
    }
    events[events_used].which = which;
    PEAQ_CUDA(cudaEventRecord(events[events_used].a, on ? on : stream));
    return 0;
  }
  int timer_end(cudaStream_t on = nullptr) {
    PEAQ_CUDA(cudaEventRecord(events[events_used].b, on ? on : stream));
    events_used++;
    return 0;
  }
  int timers_collect() {
    for (size_t i = 0; i < events_used; i++) {
      float t = 0;
      PEAQ_CUDA(cudaEventElapsedTime(&t, events[i].a, events[i].b));
      ms[events[i].which] += t;
    }
    events_used = 0;
    return 0;
  }

  template <typename T>
  int ensure(T** ptr, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    if (*ptr) PEAQ_CUDA(cudaFree(*ptr));
    *ptr = nullptr;
    *cap = 0;
    PEAQ_CUDA(cudaMalloc(ptr, need * sizeof(T)));
    *cap = need;
    return 0;
  }

  int ensure_stage(int slot, size_t floats) {
    if (floats <= stage_cap[slot]) return 0;
    size_t c0 = stage_cap[slot], c1 = stage_cap[slot];
    stage_cap[slot] = 0;   // stays 0 if either allocation fails (one buffer may be gone then)
    int rc;
    if ((rc = ensure(&d_stage[slot][0], &c0, std::max<size_t>(floats, 4)))) return rc;
    if ((rc = ensure(&d_stage[slot][1], &c1, std::max<size_t>(floats, 4)))) return rc;
    stage_cap[slot] = std::min(c0, c1);
    return 0;
  }

  int ensure_pairs(size_t n_pairs, size_t state_stride, bool preserve_state) {
    if (n_pairs > pairs_cap) {
      if (d_results) PEAQ_CUDA(cudaFree(d_results));
      if (d_nsamples) PEAQ_CUDA(cudaFree(d_nsamples));
      if (d_nframes) PEAQ_CUDA(cudaFree(d_nframes));
      d_results = nullptr;
      d_nsamples = nullptr;
      d_nframes = nullptr;
      PEAQ_CUDA(cudaMalloc(&d_results, n_pairs * sizeof(PairResult)));
      PEAQ_CUDA(cudaMalloc(&d_nsamples, 4 * n_pairs * sizeof(unsigned long long)));
      PEAQ_CUDA(cudaMalloc(&d_nframes, 2 * n_pairs * sizeof(unsigned)));
      pairs_cap = n_pairs;
    }
    const size_t need = n_pairs * state_stride;
    if (need > state_cap) {
      if (preserve_state) return fail(PEAQ_B200_ERR_INVALID, "state would be lost on growth");
      int rc = ensure(&d_state, &state_cap, need);
      if (rc) return rc;
    }
    return 0;
  }

  // Per-pair work description of one clock (host arrays, one entry per pair):
  // signal lengths in samples per channel (zero beyond) and frames to run.
  struct ClockPlan {
    const uint64_t* ns_ref;
    const uint64_t* ns_test;
    const unsigned* nf;
  };

  int upload_plan(const ClockPlan& plan, int n_pairs, int slot) {
    // pinned host copies, alive until the batch has been synchronised
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "plan words");
    const unsigned long long* a = pinned_copy(reinterpret_cast<const unsigned long long*>(plan.ns_ref), (size_t)n_pairs);
    const unsigned long long* b = pinned_copy(reinterpret_cast<const unsigned long long*>(plan.ns_test), (size_t)n_pairs);
    const unsigned* f = pinned_copy(plan.nf, (size_t)n_pairs);
    if (!a || !b || !f) return fail(PEAQ_B200_ERR_NOMEM, "out of pinned host memory");
    int rc;
    if ((rc = fetch(d_nsamples + (size_t)(2 * slot) * pairs_cap, a, (size_t)n_pairs))) return rc;
    if ((rc = fetch(d_nsamples + (size_t)(2 * slot + 1) * pairs_cap, b, (size_t)n_pairs))) return rc;
    return fetch(d_nframes + (size_t)slot * pairs_cap, f, (size_t)n_pairs);
  }

  PcmView make_view(const float* d_ref, const float* d_test, size_t pair_stride, int C, int slot) const {
    PcmView pcm;
    pcm.ref = d_ref;
    pcm.test = d_test;
    pcm.pair_stride = pair_stride;
    pcm.n_samples = d_nsamples + (size_t)(2 * slot) * pairs_cap;
    pcm.n_samples_test = d_nsamples + (size_t)(2 * slot + 1) * pairs_cap;
    pcm.n_frames = d_nframes + (size_t)slot * pairs_cap;
    pcm.channels = C;
    return pcm;
  }

  // FFT-clock chunk loop shared by both modes: K1 then the mode's scan kernel.
  template <typename ScanFn>
  int run_fft_clock(const PcmView& pcm, int n_pairs, unsigned max_frames, const RecordLayout& L, ScanFn scan) {
    const int B = h_tables->fft_bands;
    const size_t rec_bytes = (size_t)L.stride * sizeof(double);
    size_t chunk = max_frames;
    if (chunk == 0) chunk = 1;
    if ((size_t)n_pairs * chunk * rec_bytes > record_budget_bytes) {
      chunk = record_budget_bytes / ((size_t)n_pairs * rec_bytes);
      if (chunk < 1) chunk = 1;
    }
    if (keep_records) chunk = std::max<size_t>(max_frames, 1);
    int rc = ensure(&d_records, &records_cap, (size_t)n_pairs * chunk * L.stride);
    if (rc) return rc;
    unsigned first = 0;
    do {
      const unsigned n = (unsigned)std::min<size_t>(chunk, max_frames - first);
      if (n > 0) {
        if ((rc = timer_begin(1))) return rc;
        PEAQ_CUDA(launch_fft_frames(d_tables, pcm, n_pairs, first, n, d_records, L, B, advanced, stream));
        launches++;
        if ((rc = timer_end())) return rc;
      }
      if ((rc = timer_begin(2))) return rc;
      PEAQ_CUDA(scan(first, n));
      launches++;
      if ((rc = timer_end())) return rc;
      first += n;
    } while (first < max_frames);
    last_layout = L;
    last_records_doubles = keep_records ? (size_t)n_pairs * chunk * L.stride : 0;
    return 0;
  }

  // Basic mode, few pairs (long items): K2 runs one CTA per pair and is bound by the latency of its
  // frame step, K1 is frame-parallel.  The FFT clock is cut into chunks of frames and K2 of chunk
  // i (scan stream, high priority: a handful of CTAs) runs underneath K1 of chunk i + 1 (main
  // stream, fills the rest of the GPU); records are double-buffered.  Same kernels, same chunked
  // state hand-over as the serial loop: bit-identical results.
  int run_fft_clock_pipelined(const PcmView& pcm, int n_pairs, unsigned max_frames, const RecordLayout& L,
                              const StateLayout& S, PairResult* d_res) {
    const int B = h_tables->fft_bands;
    const size_t rec_bytes = (size_t)L.stride * sizeof(double);
    // >= 8 chunks so that the first K1 and the last K2 (which nothing hides) stay small, chunks of
    // at least 64 frames, two buffers within the record budget
    size_t chunk = std::max<size_t>((max_frames + 11) / 12, 64);
    const size_t fit = record_budget_bytes / (2 * (size_t)n_pairs * rec_bytes);
    chunk = std::max<size_t>(std::min(chunk, std::max<size_t>(fit, 1)), 1);
    const size_t buf_doubles = (size_t)n_pairs * chunk * L.stride;
    int rc = ensure(&d_records, &records_cap, 2 * buf_doubles);
    if (rc) return rc;
    if (!scan_stream) {
      int prio_lo = 0, prio_hi = 0;
      PEAQ_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
      PEAQ_CUDA(cudaStreamCreateWithPriority(&scan_stream, cudaStreamNonBlocking, prio_hi));
      for (int i = 0; i < 2; i++) {
        PEAQ_CUDA(cudaEventCreateWithFlags(&ev_rec_ready[i], cudaEventDisableTiming));
        PEAQ_CUDA(cudaEventCreateWithFlags(&ev_rec_free[i], cudaEventDisableTiming));
      }
    }
    // everything queued so far on the main stream (plans, state init) precedes the first scan
    PEAQ_CUDA(cudaEventRecord(ev_rec_free[0], stream));
    PEAQ_CUDA(cudaStreamWaitEvent(scan_stream, ev_rec_free[0], 0));
    unsigned first = 0;
    for (int i = 0; first < max_frames; i++) {
      const unsigned n = (unsigned)std::min<size_t>(chunk, max_frames - first);
      const int buf = i & 1;
      double* rec = d_records + (size_t)buf * buf_doubles;
      if (i >= 2) PEAQ_CUDA(cudaStreamWaitEvent(stream, ev_rec_free[buf], 0));   // K2 of chunk i - 2 has read it
      if ((rc = timer_begin(1))) return rc;
      PEAQ_CUDA(launch_fft_frames(d_tables, pcm, n_pairs, first, n, rec, L, B, advanced, stream));
      if ((rc = timer_end())) return rc;
      PEAQ_CUDA(cudaEventRecord(ev_rec_ready[buf], stream));
      PEAQ_CUDA(cudaStreamWaitEvent(scan_stream, ev_rec_ready[buf], 0));
      if ((rc = timer_begin(2, scan_stream))) return rc;
      PEAQ_CUDA(launch_scan_basic(d_tables, rec, L, pcm.n_frames, first, n, d_state, S, d_res, n_pairs, scan_stream));
      if ((rc = timer_end(scan_stream))) return rc;
      PEAQ_CUDA(cudaEventRecord(ev_rec_free[buf], scan_stream));
      launches += 2;
      first += n;
    }
    // join: the main stream continues (result copy, next batch) after the last scan
    PEAQ_CUDA(cudaEventRecord(ev_rec_ready[0], scan_stream));
    PEAQ_CUDA(cudaStreamWaitEvent(stream, ev_rec_ready[0], 0));
    last_layout = L;
    last_records_doubles = 0;
    return 0;
  }

  // Host description of a batch cut into segments (peaq_segments.cu): one entry per virtual pair
  // (an item's segments are consecutive), plus first_vp / n_seg per item.
  struct SegPlan {
    int n_items = 0;
    std::vector<unsigned long long> base;   // floats from the batch's ref / test pointers
    std::vector<unsigned> seg_index, frame0_fft, acc_start_fft, frame0_fb, acc_start_fb;
    std::vector<unsigned> first_vp, n_seg;
  };

  int upload_segments(const SegPlan& sp, int n_vp, SegTable* t) {
    int rc;
    if ((rc = ensure(&d_seg_base, &seg_base_cap, (size_t)n_vp))) return rc;
    if ((rc = ensure(&d_seg_words, &seg_words_cap, (size_t)5 * n_vp + (size_t)2 * sp.n_items))) return rc;
    const unsigned long long* base = pinned_copy(sp.base.data(), (size_t)n_vp);
    if (!base) return fail(PEAQ_B200_ERR_NOMEM, "out of pinned host memory");
    if ((rc = fetch(d_seg_base, base, (size_t)n_vp))) return rc;
    std::vector<unsigned> words;
    words.reserve((size_t)5 * n_vp + (size_t)2 * sp.n_items);
    for (const auto* v : {&sp.seg_index, &sp.frame0_fft, &sp.acc_start_fft, &sp.frame0_fb, &sp.acc_start_fb})
      words.insert(words.end(), v->begin(), v->end());
    words.insert(words.end(), sp.first_vp.begin(), sp.first_vp.end());
    words.insert(words.end(), sp.n_seg.begin(), sp.n_seg.end());
    const unsigned* pw = pinned_copy(words.data(), words.size());
    if (!pw) return fail(PEAQ_B200_ERR_NOMEM, "out of pinned host memory");
    if ((rc = fetch(d_seg_words, pw, words.size()))) return rc;
    const unsigned* w = d_seg_words;
    t->seg_index = reinterpret_cast<const int*>(w);
    t->frame0_fft = w + (size_t)n_vp;
    t->acc_start_fft = w + (size_t)2 * n_vp;
    t->frame0_fb = w + (size_t)3 * n_vp;
    t->acc_start_fb = w + (size_t)4 * n_vp;
    t->first_vp = reinterpret_cast<const int*>(w + (size_t)5 * n_vp);
    t->n_seg = reinterpret_cast<const int*>(w + (size_t)5 * n_vp + sp.n_items);
    return 0;
  }

  // Runs the frames described by the plan(s) for `n_pairs` pairs whose PCM is
  // resident on the device, reading from sample 0 of the given buffers.
  // reset_state: start from fresh state (else continue a session).
  int process_resident(const float* d_ref, const float* d_test, size_t pair_stride, int n_pairs,
                       int C, const ClockPlan& fft, const ClockPlan& fb, bool reset_state,
                       PairResult* h_out, PairResult* d_out = nullptr, bool blocking = true,
                       const float* d_ref_fb = nullptr, const float* d_test_fb = nullptr,
                       size_t pair_stride_fb = 0, const SegPlan* seg = nullptr, unsigned char* h_redo = nullptr) {
    // seg: the n_pairs "pairs" are segments of seg->n_items items (peaq_segments.cu); results
    // (h_out / d_out) are then per ITEM, h_redo[item] tells which items must be run again whole
    // d_ref_fb/d_test_fb: separate buffers for the filter-bank clock (streaming sessions
    // hold different windows of the stream per clock); default: the same buffers
    // d_out: where the kernels write the results (default: the engine's buffer);
    // blocking = false leaves everything queued on `stream` (finish_batch() later)
    const int B = h_tables->fft_bands;
    const RecordLayout L = make_record_layout(C, B);
    const StateLayout S = make_state_layout(C, B);
    const AdvStateLayout A = make_adv_state_layout(C);
    int rc = ensure_pairs((size_t)n_pairs, advanced ? (size_t)A.stride : (size_t)S.stride, !reset_state);
    if (rc) return rc;
    unsigned max_frames = 0, max_fb_frames = 0;
    for (int p = 0; p < n_pairs; p++) {
      max_frames = std::max(max_frames, fft.nf[p]);
      if (advanced) max_fb_frames = std::max(max_fb_frames, fb.nf[p]);
    }
    if ((rc = upload_plan(fft, n_pairs, 0))) return rc;
    PcmView pcm = make_view(d_ref, d_test, pair_stride, C, 0);
    PairResult* d_res = d_out ? d_out : d_results;
    SegTable seg_table = {};
    if (seg) {
      if ((rc = upload_segments(*seg, n_pairs, &seg_table))) return rc;
      pcm.base = d_seg_base;
      d_res = d_results;   // per segment; the items' rows are gathered at the end
    }

    if (!advanced) {
      if (reset_state) {
        const size_t total = (size_t)n_pairs * S.stride;
        const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 8);
        init_state_kernel<<<blocks, 256, 0, stream>>>(d_state, S, n_pairs);
        PEAQ_CUDA(cudaGetLastError());
        launches++;
        if (seg) {
          PEAQ_CUDA(launch_seg_init(d_state, &S, nullptr, n_pairs, seg_table, stream));
          launches++;
        }
      }
      // Two paths, same arithmetic, same bits (tests/test_gpu_parity.py):
      //  - frame-parallel K1 + per-pair K2 over chunks of per-frame records (default), and
      //  - PEAQ_B200_FUSED=1: the fused persistent kernel, one CTA per pair, nothing per-frame
      //    through HBM (DRAM traffic 16.8 KB per frame against 27 KB) -- measured SLOWER on B200
      //    (6.5 against 8.9 M frames/s at 4096 pairs): a CTA serialises the two halves of a frame
      //    and its 240 KB of code thrash the instruction caches, so it is opt-in.
      const bool fused = fused_mode == 1 && !keep_records;
      if (fused && !keep_records) {
        if ((rc = timer_begin(1))) return rc;
        PEAQ_CUDA(launch_fused_basic(d_tables, pcm, n_pairs, 0, std::max(max_frames, 1u), d_state, S, d_res, stream));
        launches++;
        if ((rc = timer_end())) return rc;
        last_layout = L;
        last_records_doubles = 0;
      } else if (!keep_records && max_frames >= 256 &&
                 (pipeline_mode == 1 || (pipeline_mode < 0 && n_pairs < 2 * sm_count))) {
        if ((rc = run_fft_clock_pipelined(pcm, n_pairs, max_frames, L, S, d_res))) return rc;
      } else {
        // keep_records: one chunk; K2's per-frame tap goes to the debug buffer (tests)
        double* tap = nullptr;
        last_fbdbg_doubles = 0;
        if (keep_records) {
          const size_t need = (size_t)n_pairs * std::max(max_frames, 1u) * scan_tap_doubles_per_frame(C, B);
          if ((rc = ensure(&d_fbdbg, &fbdbg_cap, need))) return rc;
          PEAQ_CUDA(cudaMemsetAsync(d_fbdbg, 0, need * sizeof(double), stream));
          tap = d_fbdbg;
          last_fbdbg_doubles = need;
          last_fb_frames = std::max(max_frames, 1u);
        }
        rc = run_fft_clock(pcm, n_pairs, max_frames, L, [&](unsigned first, unsigned n) {
          return launch_scan_basic(d_tables, d_records, L, pcm.n_frames, first, n, d_state, S, d_res,
                                   n_pairs, stream, tap);
        });
        if (rc) return rc;
      }
    } else {
      if ((rc = upload_plan(fb, n_pairs, 1))) return rc;
      PcmView pcm_fb = make_view(d_ref_fb ? d_ref_fb : d_ref, d_test_fb ? d_test_fb : d_test,
                                 d_ref_fb ? pair_stride_fb : pair_stride, C, 1);
      if (seg) pcm_fb.base = d_seg_base;
      if (reset_state) {
        PEAQ_CUDA(launch_init_adv_state(d_state, A, n_pairs, stream));
        launches++;
        if (seg) {
          PEAQ_CUDA(launch_seg_init(d_state, nullptr, &A, n_pairs, seg_table, stream));
          launches++;
        }
      }
      // ---- filter-bank clock: chunks of 192-sample frames -----------------------
      const int n_streams = n_pairs * 2 * C;
      const size_t per_frame = (size_t)n_streams * (kFbFrame * sizeof(double) + 6 * kFbBands * 3 * sizeof(double));
      size_t chunk = std::max<unsigned>(max_fb_frames, 1);
      if (chunk * per_frame > fb_budget_bytes) {
        chunk = std::max<size_t>(fb_budget_bytes / per_frame, 8);
        // 224 frames = 1344 sub-steps = 6 direct-FIR tiles of 224 = 7 recursion tiles of 192:
        // no partially filled tiles
        if (chunk >= 224) chunk -= chunk % 224;
        else if (chunk >= 32) chunk -= chunk % 32;
      }
      if (keep_records) chunk = std::max<unsigned>(max_fb_frames, 1);
      // Whole-item mode (fresh batches whose filtered signal and FFT-clock records fit): the
      // DC-reject scan of ALL frames goes to its own stream first and the frame kernel runs
      // meanwhile -- the scan keeps a few warps busy for 25 ms per 10 s of audio whatever the
      // batch, the frame kernel leaves just enough registers per SM for it.  Results are the
      // same either way (same kernels, same arithmetic).
      const size_t rec_bytes_all = (size_t)n_pairs * std::max<unsigned>(max_frames, 1) * L.stride * sizeof(double);
      // (beyond ~2 scan CTAs per SM the scan starts to displace frame-kernel CTAs and the
      // overlap stops paying: 4096 pairs measured 878 ms against 870 ms without it)
      bool whole = reset_state && !keep_records && max_fb_frames > 0 && d_ref_fb == nullptr &&
                         n_streams <= 8192 &&
                         (size_t)n_streams * (kFbHist + (size_t)max_fb_frames * kFbFrame) * sizeof(double) <=
                             hp_whole_budget_bytes &&
                         rec_bytes_all <= record_budget_bytes;
      if (whole && (size_t)n_streams * (kFbHist + (size_t)max_fb_frames * kFbFrame) > hp_cap) {
        // the whole-item buffer is an optimisation: if the device cannot hold it next to the
        // caller's PCM, fall back to the chunked path instead of failing the batch
        const size_t need = (size_t)n_streams * (kFbHist + (size_t)max_fb_frames * kFbFrame);
        if (d_hp) PEAQ_CUDA(cudaFree(d_hp));
        d_hp = nullptr;
        hp_cap = 0;
        if (cudaMalloc(&d_hp, need * sizeof(double)) == cudaSuccess) {
          hp_cap = need;
        } else {
          cudaGetLastError();
          d_hp = nullptr;
          whole = false;
        }
      }
      const size_t hp_stride = kFbHist + (whole ? (size_t)max_fb_frames : chunk) * kFbFrame;
      if ((rc = ensure(&d_hp, &hp_cap, (size_t)n_streams * hp_stride))) return rc;
      if ((size_t)n_streams * kHpStateDoubles > hp_state_cap && !reset_state)
        return fail(PEAQ_B200_ERR_INVALID, "filter state would be lost on growth");
      if ((rc = ensure(&d_hp_state, &hp_state_cap, (size_t)n_streams * kHpStateDoubles))) return rc;
      if ((rc = ensure(&d_fbout, &fbout_cap, (size_t)n_streams * kFbBands * chunk * 6 * 2))) return rc;
      if ((rc = ensure(&d_fbenergy, &fbenergy_cap, (size_t)n_streams * kFbBands * chunk * 6))) return rc;
      if ((rc = ensure(&d_fbflags, &fbflags_cap, (size_t)n_pairs * chunk))) return rc;
      double* dbg = nullptr;
      last_fbdbg_doubles = 0;
      if (keep_records) {
        const size_t need = (size_t)n_pairs * chunk * (2 * C * 2 * kFbBands + C * 8);
        if ((rc = ensure(&d_fbdbg, &fbdbg_cap, need))) return rc;
        PEAQ_CUDA(cudaMemsetAsync(d_fbdbg, 0, need * sizeof(double), stream));
        dbg = d_fbdbg;
        last_fbdbg_doubles = need;
        last_fb_frames = (unsigned)chunk;
      }
      if (whole) {
        PEAQ_CUDA(cudaEventRecord(ev_hp_ready, stream));          // plans uploaded, PCM resident, state fresh
        PEAQ_CUDA(cudaStreamWaitEvent(hp_stream, ev_hp_ready, 0));
        PEAQ_CUDA(launch_fb_hp(d_tables, pcm_fb, n_pairs, 0ull, 0ull, max_fb_frames * kFbFrame, d_hp, hp_stride,
                               d_hp_state, true, hp_stream));
        PEAQ_CUDA(cudaEventRecord(ev_hp_done, hp_stream));
        launches++;
        // the frame kernel of the whole FFT clock, underneath which the scan runs
        if ((rc = ensure(&d_records, &records_cap, (size_t)n_pairs * std::max<unsigned>(max_frames, 1) * L.stride)))
          return rc;
        if (max_frames > 0) {
          if ((rc = timer_begin(1))) return rc;
          PEAQ_CUDA(launch_fft_frames(d_tables, pcm, n_pairs, 0, max_frames, d_records, L, B, advanced, stream));
          launches++;
          if ((rc = timer_end())) return rc;
        }
        PEAQ_CUDA(cudaStreamWaitEvent(stream, ev_hp_done, 0));
      }
      if (reset_state) fb_pos = 0;   // samples the filter-bank clock has consumed before this call (sessions continue)
      unsigned first = 0;
      while (first < max_fb_frames) {
        const unsigned n = (unsigned)std::min<size_t>(chunk, max_fb_frames - first);
        const unsigned n_sub = n * 6, samples = n * kFbFrame;
        if ((rc = timer_begin(7))) return rc;
        PEAQ_CUDA(launch_fb_flags(pcm_fb, n_pairs, first, n, d_fbflags, stream));
        if (!whole) {
          PEAQ_CUDA(launch_fb_hp(d_tables, pcm_fb, n_pairs, (unsigned long long)first * kFbFrame,
                                 fb_pos + (unsigned long long)first * kFbFrame, samples, d_hp, hp_stride, d_hp_state,
                                 first == 0 && reset_state, stream));
        }
        if ((rc = timer_end())) return rc;
        // PEAQ_B200_FB_DIRECT=1: all 40 filters as direct FIRs (development aid / cross-check)
        static const bool fb_direct = std::getenv("PEAQ_B200_FB_DIRECT") && std::atoi(std::getenv("PEAQ_B200_FB_DIRECT"));
        if ((rc = timer_begin(5))) return rc;
        // whole-item mode: the chunk starts `first` frames into the filtered signal, its FIR
        // history is simply what precedes it there
        PEAQ_CUDA(launch_fb_bank(d_tables, h_tables, d_hp + (whole ? (size_t)first * kFbFrame : 0), hp_stride,
                                 n_streams, n_sub, d_fbout, d_hp_state, first == 0 && reset_state, fb_direct,
                                 pcm_fb.n_frames, first, 2 * C, stream));
        if ((rc = timer_end())) return rc;
        if ((rc = timer_begin(6))) return rc;
        PEAQ_CUDA(launch_fb_spread(d_tables, d_fbout, n_sub, pcm_fb.n_frames, first, d_state, A, d_fbenergy,
                                   n_pairs, stream));
        PEAQ_CUDA(launch_fb_scan(d_tables, d_fbenergy, n_sub, d_fbflags, pcm_fb.n_frames, first, n, d_state, A,
                                 dbg, n_pairs, stream));
        launches += whole ? 4 : 5;
        if ((rc = timer_end())) return rc;
        first += n;
      }
      fb_pos += (unsigned long long)max_fb_frames * kFbFrame;
      if (max_fb_frames == 0 && reset_state) {
        // publish the (empty) fb-clock MOVs so the epilogue sees 0/0 like the reference
        PEAQ_CUDA(launch_fb_scan(d_tables, d_fbenergy, 0, d_fbflags, pcm_fb.n_frames, 0, 0, d_state, A, nullptr,
                                 n_pairs, stream));
        launches++;
      }
      // ---- FFT clock; its epilogue combines all five MOVs ---------------------------
      if (whole) {
        if ((rc = timer_begin(2))) return rc;
        PEAQ_CUDA(launch_adv_fft_scan(d_tables, d_records, L, pcm.n_frames, 0, max_frames, d_state, A, d_res,
                                      n_pairs, stream));
        launches++;
        if ((rc = timer_end())) return rc;
        last_layout = L;
        last_records_doubles = 0;
      } else {
        rc = run_fft_clock(pcm, n_pairs, max_frames, L, [&](unsigned first_f, unsigned n) {
          return launch_adv_fft_scan(d_tables, d_records, L, pcm.n_frames, first_f, n, d_state, A, d_res,
                                     n_pairs, stream);
        });
        if (rc) return rc;
      }
    }

    int n_out = n_pairs;
    if (seg) {
      // sums of the segments -> segment 0 of every item, the scan kernels' epilogues once more
      // (zero frames) on the combined state, then one result row per item
      n_out = seg->n_items;
      if ((rc = ensure(&d_item_results, &item_results_cap, (size_t)n_out))) return rc;
      // the redo flags are written straight into pinned host memory (no copy engine, see fetch())
      unsigned char* redo_pin = static_cast<unsigned char*>(arena.take((size_t)n_out));
      if (!redo_pin) return fail(PEAQ_B200_ERR_NOMEM, "out of pinned host memory");
      if (!advanced) {
        PEAQ_CUDA(launch_seg_combine(d_state, &S, nullptr, n_out, seg_table, redo_pin, stream));
        PEAQ_CUDA(launch_scan_basic(d_tables, d_records, L, pcm.n_frames, 0, 0, d_state, S, d_res, n_pairs, stream));
        launches += 2;
      } else {
        PEAQ_CUDA(launch_seg_combine(d_state, nullptr, &A, n_out, seg_table, redo_pin, stream));
        PEAQ_CUDA(launch_fb_scan(d_tables, d_fbenergy, 0, d_fbflags, pcm.n_frames, 0, 0, d_state, A, nullptr, n_pairs,
                                 stream));
        PEAQ_CUDA(launch_adv_fft_scan(d_tables, d_records, L, pcm.n_frames, 0, 0, d_state, A, d_res, n_pairs, stream));
        launches += 3;
      }
      PairResult* d_items = d_out ? d_out : d_item_results;
      PEAQ_CUDA(launch_seg_gather_results(d_res, seg_table.first_vp, n_out, d_items, stream));
      launches++;
      d_res = d_items;
      if (h_redo) deferred.push_back({h_redo, redo_pin, (size_t)n_out});
    }
    if (!blocking) return 0;
    if (h_out) {
      PEAQ_CUDA(cudaMemcpyAsync(h_out, d_res, n_out * sizeof(PairResult), cudaMemcpyDeviceToHost, stream));
    }
    return finish_batch();
  }

  // how many segments an item of n samples is cut into, and their length (a function of n alone,
  // so that an item's result does not depend on the batch it is part of)
  static unsigned segments_for_samples(uint64_t n, uint64_t* seg_len) {
    // about kSegSamples each, all segments of an item equally long (no ragged tail: the last
    // segment would otherwise keep every per-stream kernel running for the others' sake),
    // boundaries on multiples of both frame steps and of the DC-reject scan's blocks
    const uint64_t grid = 3072;   // lcm(1024, 192), = 6 blocks of 512
    uint64_t k = std::max<uint64_t>(1, (n + kSegSamples / 2) / kSegSamples);
    k = std::min<uint64_t>(k, 512);   // seg_combine_* handle 512 segments per item
    uint64_t len = ((n + k - 1) / k + grid - 1) / grid * grid;
    while (k > 1 && (k - 1) * len >= n) k--;
    *seg_len = len;
    return (unsigned)k;
  }

  // A fresh batch of whole items, PCM resident on the device: items long enough are cut into
  // segments (peaq_segments.cu), the others run as they are.  h_redo (n_items bytes, may be filled
  // asynchronously like h_out): items whose segments' assumptions did not hold; the caller runs
  // them again with segments = false.
  int process_items(const float* d_ref, const float* d_test, size_t pair_stride, int n_items, int C,
                    const uint64_t* ns, const unsigned* nf, const unsigned* nfb, PairResult* h_out,
                    PairResult* d_out, bool blocking, unsigned char* h_redo, bool segments = true) {
    bool any = false;
    if (segments && segment_mode && !keep_records) {
      uint64_t len;
      for (int i = 0; i < n_items && !any; i++) any = segments_for_samples(ns[i], &len) > 1;
    }
    if (!any) {
      if (h_redo) std::memset(h_redo, 0, (size_t)n_items);
      const ClockPlan fft{ns, ns, nf}, fb{ns, ns, nfb};
      return process_resident(d_ref, d_test, pair_stride, n_items, C, fft, fb, true, h_out, d_out, blocking);
    }
    SegPlan sp;
    sp.n_items = n_items;
    std::vector<uint64_t> v_ns_fft, v_ns_fb;
    std::vector<unsigned> v_nf, v_nfb;
    for (int i = 0; i < n_items; i++) {
      uint64_t len;
      const unsigned k_seg = segments_for_samples(ns[i], &len);
      sp.first_vp.push_back((unsigned)sp.base.size());
      sp.n_seg.push_back(k_seg);
      for (unsigned k = 0; k < k_seg; k++) {
        const uint64_t a = (uint64_t)k * len, w = k ? a - kSegWarmSamples : 0;   // owned from a, run from w
        const bool last = k + 1 == k_seg;
        sp.base.push_back((unsigned long long)i * pair_stride + w * C);
        sp.seg_index.push_back(k);
        sp.frame0_fft.push_back((unsigned)(w / kFftStep));
        sp.acc_start_fft.push_back((unsigned)(a / kFftStep));
        sp.frame0_fb.push_back((unsigned)(w / kFbFrame));
        sp.acc_start_fb.push_back((unsigned)(a / kFbFrame));
        if (last) {
          v_ns_fft.push_back(ns[i] - w);
          v_ns_fb.push_back(ns[i] - w);
          v_nf.push_back(nf[i] - (unsigned)(w / kFftStep));
          v_nfb.push_back(nfb[i] - (unsigned)(w / kFbFrame));
        } else {
          const uint64_t span = a + len - w;   // multiples of both frame steps
          v_ns_fft.push_back(span + (kFftFrame - kFftStep));   // the last frame's second half
          v_ns_fb.push_back(span);
          v_nf.push_back((unsigned)(span / kFftStep));
          v_nfb.push_back((unsigned)(span / kFbFrame));
        }
      }
    }
    const int n_vp = (int)sp.base.size();
    const ClockPlan fft{v_ns_fft.data(), v_ns_fft.data(), v_nf.data()}, fb{v_ns_fb.data(), v_ns_fb.data(), v_nfb.data()};
    return process_resident(d_ref, d_test, pair_stride, n_vp, C, fft, fb, true, h_out, d_out, blocking, nullptr, nullptr,
                            0, &sp, h_redo);
  }

  int finish_batch() {
    PEAQ_CUDA(cudaStreamSynchronize(stream));
    for (const auto& d : deferred) std::memcpy(d.dst, d.src, d.bytes);
    deferred.clear();
    arena.reset();
    return timers_collect();
  }
};

static int check_channels(int channels) {
  if (channels < 1 || channels > kMaxChannels)
    return fail(PEAQ_B200_ERR_INVALID, "channels must be 1 or 2");
  return 0;
}

// owners for the few CUDA objects run_batch creates: every early return releases them
struct ScopedEvent {
  cudaEvent_t ev = nullptr;
  ~ScopedEvent() { if (ev) cudaEventDestroy(ev); }
};
struct ScopedDeviceMem {
  void* p = nullptr;
  ~ScopedDeviceMem() { if (p) cudaFree(p); }
};
// a failed batch must not leave half-recorded timers or pinned plan copies behind (the next
// successful batch would read an event pair whose end was never recorded)
struct BatchCleanup {
  Engine* e;
  bool ok = false;
  ~BatchCleanup() {
    if (ok) return;
    cudaStreamSynchronize(e->stream);
    cudaStreamSynchronize(e->copy_stream);
    if (e->scan_stream) cudaStreamSynchronize(e->scan_stream);
    e->events_used = 0;
    e->deferred.clear();
    e->arena.reset();
    cudaGetLastError();
  }
};

static int run_batch(Engine* e, const peaq_b200_batch* b, peaq_b200_result* out) {
  if (!e || !b || !out) return fail(PEAQ_B200_ERR_INVALID, "null argument");
  if (b->n_pairs <= 0) return fail(PEAQ_B200_ERR_INVALID, "n_pairs must be positive");
  int rc = check_channels(b->channels);
  if (rc) return rc;
  if (!b->ref || !b->test) return fail(PEAQ_B200_ERR_INVALID, "null PCM pointer");
  if (b->on_device && ((reinterpret_cast<uintptr_t>(b->ref) | reinterpret_cast<uintptr_t>(b->test)) & 15))
    return fail(PEAQ_B200_ERR_INVALID, "device PCM pointers must be 16-byte aligned");
  const int C = b->channels;
  const int n_pairs = b->n_pairs;
  std::vector<uint64_t> ns(n_pairs);
  std::vector<unsigned> nf(n_pairs), nfb(n_pairs);
  uint64_t max_n = 0;
  for (int p = 0; p < n_pairs; p++) {
    ns[p] = b->n_samples ? b->n_samples[p] : b->n_samples_all;
    if (ns[p] > ((uint64_t)UINT_MAX - 4) * kFftStep)
      return fail(PEAQ_B200_ERR_INVALID, "item too long");
    nf[p] = frames_for_samples(ns[p]);
    nfb[p] = fb_frames_for_samples(ns[p]);
    max_n = std::max(max_n, ns[p]);
    if (n_pairs > 1 && ns[p] * C > b->pair_stride)
      return fail(PEAQ_B200_ERR_INVALID, "pair_stride smaller than an item");
  }
  PEAQ_CUDA(cudaSetDevice(e->device));
  for (int i = 0; i < 8; i++) e->ms[i] = 0;
  e->events_used = 0;
  BatchCleanup cleanup{e};
  ScopedEvent t0, t1;
  PEAQ_CUDA(cudaEventCreate(&t0.ev));
  PEAQ_CUDA(cudaEventCreate(&t1.ev));
  PEAQ_CUDA(cudaEventRecord(t0.ev, e->stream));

  PairResult* res = reinterpret_cast<PairResult*>(out);
  std::vector<unsigned char> redo((size_t)n_pairs, 0);   // items whose segments must be run again whole
  if (b->on_device) {
    rc = e->process_items(b->ref, b->test, b->pair_stride, n_pairs, C, ns.data(), nf.data(), nfb.data(), res, nullptr,
                          true, redo.data());
    if (rc) return rc;
    for (int p = 0; p < n_pairs; p++) {
      if (!redo[p]) continue;
      rc = e->process_items(b->ref + (size_t)p * b->pair_stride, b->test + (size_t)p * b->pair_stride, b->pair_stride,
                            1, C, &ns[p], &nf[p], &nfb[p], res + p, nullptr, true, nullptr, false);
      if (rc) return rc;
    }
  } else {
    // host input: sub-batches of pairs are staged through two device slots; the
    // H2D copy of sub-batch i+1 (copy stream) overlaps the kernels of sub-batch i
    const size_t stride = n_pairs > 1 ? b->pair_stride : (size_t)max_n * C;
    const size_t slot_budget = (size_t)1 << 30;   // floats per staging buffer (4 GiB)
    const int per_max = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_pairs, slot_budget / std::max<size_t>(stride, 1)));
    // Sub-batch sizes.  Basic mode is bound by the copies (16 KB of PCM per frame against
    // ~100 ns of kernels), so the job takes all copies plus the kernels of the LAST sub-batch:
    // small, equal sub-batches (one wave of the scan kernel's CTAs, two per SM).  Advanced mode
    // is bound by the kernels, so it takes the FIRST copy plus all kernels: a small first
    // sub-batch, then growing ones.
    std::vector<int> sizes;
    {
      // the counts below were tuned on 10 s stereo pairs (3.84 MB per signal); longer items weigh
      // as many of those as they are long, so that sub-batches keep their size in bytes
      const double weight = std::max(1.0, (double)stride / 960000.0);
      auto pairs_of = [&](int standard_pairs) { return std::max(1, (int)(standard_pairs / weight)); };
      int left = n_pairs;
      if (e->advanced) {
        // kernel-bound (0.19 ms of kernels against 0.14 ms of copies per standard pair): the job
        // takes the FIRST copy plus all kernels as long as no later copy outlasts the kernels of
        // the sub-batch before it -- sub-batches grow by 1.4 x from 128 standard pairs (one
        // 128 + 4 x 1024 schedule left 140 ms of the second copy exposed: 943 ms per 4096 pairs)
        // ... and no sub-batch so small that it cannot fill the GPU: at least 128 pairs or
        // segments (a 10-minute item is 18 segments: 8 items)
        uint64_t seg_len = 0;
        const int vp_per_pair = e->segment_mode && !e->keep_records ? (int)Engine::segments_for_samples(max_n, &seg_len) : 1;
        const int floor_pairs = std::min(per_max, std::max(pairs_of(128), (128 + vp_per_pair - 1) / vp_per_pair));
        double next = floor_pairs;
        const int cap = std::max(floor_pairs, pairs_of(1200));
        while (left > 0) {
          int n = std::min({left, per_max, std::max(floor_pairs, (int)next)});
          if (left - n < std::max(1, floor_pairs / 2)) n = std::min(left, per_max);   // no crumbs at the end
          sizes.push_back(n);
          left -= n;
          next = std::min(next * 1.4, (double)cap);
        }
      } else {
        const int per = std::min(per_max, pairs_of((e->fused_mode == 1 ? 3 : 2) * e->sm_count));
        while (left > 0) {
          // copy-bound: the job ends with the kernels of the last sub-batches, which are small
          const int n = left <= per ? std::min(left, pairs_of(128)) : std::min(per, left);
          sizes.push_back(n);
          left -= n;
        }
      }
    }
    if (std::getenv("PEAQ_B200_DEBUG")) {
      std::fprintf(stderr, "peaq_b200: %d sub-batches:", (int)sizes.size());
      for (int n : sizes) std::fprintf(stderr, " %d", n);
      std::fprintf(stderr, " pairs\n");
    }
    // size everything once, for the largest sub-batch: growing later would free memory under
    // queued kernels (cudaFree synchronises the device and stalls the copy / compute overlap)
    {
      int np_max = 0;   // in virtual pairs: long items count once per segment
      for (int i = 0, q0 = 0; i < (int)sizes.size(); q0 += sizes[i], i++) {
        int n_vp = 0;
        for (int q = q0; q < q0 + sizes[i]; q++) {
          uint64_t len;
          n_vp += e->segment_mode && !e->keep_records ? (int)Engine::segments_for_samples(ns[q], &len) : 1;
        }
        np_max = std::max(np_max, n_vp);
      }
      const StateLayout S = make_state_layout(C, e->h_tables->fft_bands);
      const AdvStateLayout A = make_adv_state_layout(C);
      if ((rc = e->ensure_pairs((size_t)np_max, e->advanced ? (size_t)A.stride : (size_t)S.stride, false))) return rc;
    }
    ScopedDeviceMem d_all_mem;
    PEAQ_CUDA(cudaMalloc(&d_all_mem.p, (size_t)n_pairs * sizeof(PairResult)));
    PairResult* d_all = static_cast<PairResult*>(d_all_mem.p);
    int p0 = 0;
    for (int i = 0; i < (int)sizes.size(); p0 += sizes[i], i++) {
      const int np = sizes[i];
      const int slot = i & 1;
      // the caller's arrays end with the last pair's samples: never read past them
      const size_t floats = (size_t)(np - 1) * stride + (size_t)ns[p0 + np - 1] * C;
      // (pairs before the last one are read up to the stride; the kernels read ns[p] samples)
      if (i >= 2) PEAQ_CUDA(cudaStreamWaitEvent(e->copy_stream, e->ev_freed[slot], 0));
      const size_t need = (size_t)(np - 1) * stride + (size_t)max_n * C;   // same capacity for every sub-batch
      if (need > e->stage_cap[slot]) {
        // growing a slot frees memory that queued kernels may still read: drain first
        PEAQ_CUDA(cudaStreamSynchronize(e->stream));
        if ((rc = e->ensure_stage(slot, need))) return rc;
      }
      if (floats) {
        PEAQ_CUDA(cudaMemcpyAsync(e->d_stage[slot][0], b->ref + (size_t)p0 * stride, floats * sizeof(float),
                                  cudaMemcpyHostToDevice, e->copy_stream));
        PEAQ_CUDA(cudaMemcpyAsync(e->d_stage[slot][1], b->test + (size_t)p0 * stride, floats * sizeof(float),
                                  cudaMemcpyHostToDevice, e->copy_stream));
      }
      PEAQ_CUDA(cudaEventRecord(e->ev_copied[slot], e->copy_stream));
      PEAQ_CUDA(cudaStreamWaitEvent(e->stream, e->ev_copied[slot], 0));
      rc = e->process_items(e->d_stage[slot][0], e->d_stage[slot][1], stride, np, C, ns.data() + p0, nf.data() + p0,
                            nfb.data() + p0, nullptr, d_all + p0, false, redo.data() + p0);
      if (rc) return rc;
      PEAQ_CUDA(cudaEventRecord(e->ev_freed[slot], e->stream));
    }
    PEAQ_CUDA(cudaMemcpyAsync(res, d_all, (size_t)n_pairs * sizeof(PairResult), cudaMemcpyDeviceToHost, e->stream));
    rc = e->finish_batch();
    if (rc) return rc;
    for (int p = 0; p < n_pairs; p++) {
      if (!redo[p]) continue;
      // (rare: an item whose first half minute is silent) staged again and run as a whole
      const size_t floats = (size_t)ns[p] * C;
      PEAQ_CUDA(cudaMemcpyAsync(e->d_stage[0][0], b->ref + (size_t)p * stride, floats * sizeof(float),
                                cudaMemcpyHostToDevice, e->stream));
      PEAQ_CUDA(cudaMemcpyAsync(e->d_stage[0][1], b->test + (size_t)p * stride, floats * sizeof(float),
                                cudaMemcpyHostToDevice, e->stream));
      rc = e->process_items(e->d_stage[0][0], e->d_stage[0][1], stride, 1, C, &ns[p], &nf[p], &nfb[p], res + p, nullptr,
                            true, nullptr, false);
      if (rc) return rc;
    }
  }
  PEAQ_CUDA(cudaEventRecord(t1.ev, e->stream));
  PEAQ_CUDA(cudaEventSynchronize(t1.ev));
  float t = 0;
  PEAQ_CUDA(cudaEventElapsedTime(&t, t0.ev, t1.ev));
  e->ms[0] = t;
  cleanup.ok = true;
  return 0;
}

// ---------------------------------------------------------------------------
// session = one element instance

// GstAdapter stand-in: append at the back, take from the front.  Consuming advances a read
// offset; the storage is compacted only when more than half of it is dead.
struct Fifo {
  std::vector<float> buf;
  size_t head = 0;
  size_t size() const { return buf.size() - head; }
  bool empty() const { return size() == 0; }
  const float* data() const { return buf.data() + head; }
  void append(const float* p, size_t n) {
    if (head && head >= buf.size() / 2) {
      buf.erase(buf.begin(), buf.begin() + head);
      head = 0;
    }
    buf.insert(buf.end(), p, p + n);
  }
  void consume(size_t n) {
    head += n;
    if (head >= buf.size()) {
      buf.clear();
      head = 0;
    }
  }
  void clear() {
    buf.clear();
    head = 0;
  }
};

struct Session {
  int device = 0;
  bool advanced = false;
  double level = 92.0;
  int channels = 0;
  Engine* engine = nullptr;
  bool started = false;          // recurrent state initialised on the device
  // GstAdapter stand-ins: [0] FFT clock (ref_adapter_fft / test_adapter_fft), [1] filter-bank
  // clock (ref_adapter_fb / test_adapter_fb, advanced mode only; gstpeaq.c:116-119, :626-635)
  Fifo fifo[2][2];   // [clock][ref|test]
  PairResult last = {};
  bool have_result = false;
  // Streaming without stalling the caller (the element's streaming thread, gstpeaq.c:614-661): a
  // push copies the frames that became complete into one of two PINNED staging sets, queues the
  // copy to the device and the kernels on the engine's stream, and returns.  Nothing waits for the
  // GPU until a result is read (properties odg / di / totalsnr, gstpeaq.c:484-497), the session is
  // finished or snapshotted, or kMaxPending pushes are in flight.
  static constexpr int kMaxPending = 64;
  float* pin[2][2][2] = {};          // [set][clock][ref|test]
  size_t pin_cap[2][2] = {};         // floats per buffer of [set][clock]
  cudaEvent_t pin_done[2] = {nullptr, nullptr};
  bool pin_used[2] = {false, false};
  int pin_next = 0;
  int pending = 0;                   // pushes queued since the last synchronisation
  bool result_stale = false;         // the device holds a newer result than `last`

  void reset_stream() {
    sync();
    started = false;
    have_result = false;
    result_stale = false;
    for (auto& c : fifo)
      for (auto& f : c) f.clear();
  }

  void free_pinned() {
    for (int i = 0; i < 2; i++) {
      for (int c = 0; c < 2; c++) {
        for (int sd = 0; sd < 2; sd++) {
          if (pin[i][c][sd]) cudaFreeHost(pin[i][c][sd]);
          pin[i][c][sd] = nullptr;
        }
        pin_cap[i][c] = 0;
      }
      if (pin_done[i]) cudaEventDestroy(pin_done[i]);
      pin_done[i] = nullptr;
      pin_used[i] = false;
    }
  }

  void drop_engine() {
    reset_stream();
    if (engine) {
      engine->destroy();
      delete engine;
      engine = nullptr;
    }
    free_pinned();
  }

  int ensure_engine() {
    if (engine) return 0;
    engine = new (std::nothrow) Engine;
    if (!engine) return fail(PEAQ_B200_ERR_NOMEM, "out of host memory");
    engine->device = device;
    engine->advanced = advanced;
    engine->level = level;
    int rc = engine->init();
    if (rc) {
      engine->destroy();
      delete engine;
      engine = nullptr;
    }
    return rc;
  }

  // waits for everything queued so far; afterwards plan copies and timers are collected
  int sync() {
    if (!engine || pending == 0) return 0;
    pending = 0;
    PEAQ_CUDA(cudaSetDevice(device));
    return engine->finish_batch();
  }

  // latest result on the host
  int fetch_result() {
    if (!result_stale) return 0;
    PEAQ_CUDA(cudaSetDevice(device));
    PEAQ_CUDA(cudaMemcpyAsync(&last, engine->d_results, sizeof(PairResult), cudaMemcpyDeviceToHost, engine->stream));
    pending++;
    int rc = sync();
    if (rc) return rc;
    result_stale = false;
    return 0;
  }

  // copies `floats` of each signal into pinned set `set` and queues the transfer to the device slot
  int stage(int set, int clock, const float* ref, const float* test, size_t floats) {
    int rc;
    if ((rc = engine->ensure_stage(clock, floats))) return rc;
    if (!floats) return 0;
    if (floats > pin_cap[set][clock]) {
      const size_t cap = std::max<size_t>(floats, (size_t)4096 * channels);
      for (int sd = 0; sd < 2; sd++) {
        if (pin[set][clock][sd]) PEAQ_CUDA(cudaFreeHost(pin[set][clock][sd]));
        pin[set][clock][sd] = nullptr;
      }
      pin_cap[set][clock] = 0;
      for (int sd = 0; sd < 2; sd++) PEAQ_CUDA(cudaHostAlloc((void**)&pin[set][clock][sd], cap * sizeof(float), cudaHostAllocDefault));
      pin_cap[set][clock] = cap;
    }
    std::memcpy(pin[set][clock][0], ref, floats * sizeof(float));
    std::memcpy(pin[set][clock][1], test, floats * sizeof(float));
    PEAQ_CUDA(cudaMemcpyAsync(engine->d_stage[clock][0], pin[set][clock][0], floats * sizeof(float),
                              cudaMemcpyHostToDevice, engine->stream));
    PEAQ_CUDA(cudaMemcpyAsync(engine->d_stage[clock][1], pin[set][clock][1], floats * sizeof(float),
                              cudaMemcpyHostToDevice, engine->stream));
    return 0;
  }

  // Queues k_fft frames of the FFT clock over the first n_fft samples of (ref_fft, test_fft)
  // and k_fb frames of the filter-bank clock over the first n_fb samples of (ref_fb, test_fb),
  // continuing from the state on the device.  The host buffers are free again on return.
  int run_frames(unsigned k_fft, const float* ref_fft, const float* test_fft, size_t n_fft, unsigned k_fb,
                 const float* ref_fb, const float* test_fb, size_t n_fb) {
    int rc = ensure_engine();
    if (rc) return rc;
    PEAQ_CUDA(cudaSetDevice(device));
    const size_t fl_fft = n_fft * channels, fl_fb = n_fb * channels;
    // growing a device slot or a pinned set frees memory queued work may still use: drain first
    const int set = pin_next;
    pin_next ^= 1;
    if (fl_fft > engine->stage_cap[0] || (advanced && fl_fb > engine->stage_cap[1]) || fl_fft > pin_cap[set][0] ||
        (advanced && fl_fb > pin_cap[set][1])) {
      if ((rc = sync())) return rc;
      PEAQ_CUDA(cudaStreamSynchronize(engine->stream));
    }
    if (!pin_done[set]) PEAQ_CUDA(cudaEventCreateWithFlags(&pin_done[set], cudaEventDisableTiming));
    if (pin_used[set]) PEAQ_CUDA(cudaEventSynchronize(pin_done[set]));   // the copy issued two pushes ago
    if ((rc = stage(set, 0, ref_fft, test_fft, fl_fft))) return rc;
    if (advanced && (rc = stage(set, 1, ref_fb, test_fb, fl_fb))) return rc;
    PEAQ_CUDA(cudaEventRecord(pin_done[set], engine->stream));
    pin_used[set] = true;
    const uint64_t ns_fft = n_fft, ns_fb = n_fb;
    const Engine::ClockPlan fft{&ns_fft, &ns_fft, &k_fft}, fb{&ns_fb, &ns_fb, &k_fb};
    rc = engine->process_resident(engine->d_stage[0][0], engine->d_stage[0][1], std::max<size_t>(fl_fft, 1), 1,
                                  channels, fft, fb, !started, nullptr, nullptr, false,
                                  advanced ? engine->d_stage[1][0] : nullptr,
                                  advanced ? engine->d_stage[1][1] : nullptr, std::max<size_t>(fl_fb, 1));
    if (rc) return rc;
    started = true;
    have_result = true;
    result_stale = true;
    if (++pending >= kMaxPending) return sync();
    return 0;
  }

  // whole frames available on one clock (do_processing, gstpeaq.c:596-611)
  unsigned frames_ready(int clock, unsigned F, unsigned St, size_t* n_samples) const {
    const size_t fl = std::min(fifo[clock][0].size(), fifo[clock][1].size());
    const size_t frame = (size_t)F * channels, step = (size_t)St * channels;
    if (fl < frame) {
      *n_samples = 0;
      return 0;
    }
    const unsigned k = (unsigned)((fl - frame) / step + 1);
    *n_samples = (size_t)(k - 1) * St + F;
    return k;
  }

  void consume(int clock, unsigned k, unsigned St) {
    const size_t n = (size_t)k * St * channels;
    for (int s = 0; s < 2; s++) fifo[clock][s].consume(n);
  }

  // pad_chain's processing part (gstpeaq.c:637-656): drain both clocks
  int drain() {
    size_t n_fft = 0, n_fb = 0;
    const unsigned k_fft = frames_ready(0, kFftFrame, kFftStep, &n_fft);
    const unsigned k_fb = advanced ? frames_ready(1, kFbFrame, kFbFrame, &n_fb) : 0;
    if (k_fft == 0 && k_fb == 0) return 0;
    int rc = run_frames(k_fft, fifo[0][0].data(), fifo[0][1].data(), n_fft, k_fb, fifo[1][0].data(),
                        fifo[1][1].data(), n_fb);
    if (rc) return rc;
    consume(0, k_fft, kFftStep);
    if (advanced) consume(1, k_fb, kFbFrame);
    return 0;
  }

  // do_flush (gstpeaq.c:716-745) on both clocks: ONE zero-padded frame per adapter
  // pair made of MIN(left, frame) samples of each stream
  int flush() {
    std::vector<float> pad[2][2];
    unsigned k[2] = {0, 0};
    const unsigned F[2] = {kFftFrame, kFbFrame};
    for (int c = 0; c < (advanced ? 2 : 1); c++) {
      if (fifo[c][0].empty() && fifo[c][1].empty()) continue;
      const size_t frame = (size_t)F[c] * channels;
      for (int s = 0; s < 2; s++) {
        pad[c][s].assign(frame, 0.f);
        const size_t n = std::min(fifo[c][s].size(), frame);
        std::copy(fifo[c][s].data(), fifo[c][s].data() + n, pad[c][s].begin());
        fifo[c][s].consume(n);
      }
      k[c] = 1;
    }
    if (!k[0] && !k[1]) return 0;
    return run_frames(k[0], pad[0][0].data(), pad[0][1].data(), k[0] ? kFftFrame : 0, k[1], pad[1][0].data(),
                      pad[1][1].data(), k[1] ? kFbFrame : 0);
  }

  // ---- snapshot / restore ------------------------------------------------------------------
  // Layout: SnapHeader | 4 FIFO contents (float) | device state block (double) | DC-reject and
  // filter-bank chain state (double, advanced) | last result
  struct SnapHeader {
    uint32_t magic, version;
    int32_t advanced, channels, started, have_result;
    double level;
    uint64_t fifo_floats[2][2];
    uint64_t state_doubles, hp_doubles;
    uint64_t fb_pos;   // samples the filter-bank clock has consumed (anchors the DC-reject block scan)
  };
  static constexpr uint32_t kSnapMagic = 0x51414550u;   // "PEAQ"

  size_t state_doubles() const {
    if (!started) return 0;
    const int B = engine->h_tables->fft_bands;
    return advanced ? (size_t)make_adv_state_layout(channels).stride : (size_t)make_state_layout(channels, B).stride;
  }
  size_t hp_doubles() const { return started && advanced ? (size_t)2 * channels * kHpStateDoubles : 0; }

  int snapshot(void* buf, size_t capacity, size_t* size) {
    int rc = fetch_result();
    if (rc) return rc;
    if ((rc = sync())) return rc;
    SnapHeader h = {};
    h.magic = kSnapMagic;
    h.version = 1;
    h.advanced = advanced;
    h.channels = channels;
    h.started = started;
    h.have_result = have_result;
    h.level = level;
    size_t total = sizeof h;
    for (int c = 0; c < 2; c++)
      for (int sd = 0; sd < 2; sd++) {
        h.fifo_floats[c][sd] = fifo[c][sd].size();
        total += fifo[c][sd].size() * sizeof(float);
      }
    h.state_doubles = state_doubles();
    h.hp_doubles = hp_doubles();
    h.fb_pos = engine ? engine->fb_pos : 0;
    total += (h.state_doubles + h.hp_doubles) * sizeof(double) + sizeof(PairResult);
    if (size) *size = total;
    if (!buf) return 0;
    if (capacity < total) return fail(PEAQ_B200_ERR_INVALID, "snapshot buffer too small");
    unsigned char* p = static_cast<unsigned char*>(buf);
    std::memcpy(p, &h, sizeof h);
    p += sizeof h;
    for (int c = 0; c < 2; c++)
      for (int sd = 0; sd < 2; sd++) {
        std::memcpy(p, fifo[c][sd].data(), fifo[c][sd].size() * sizeof(float));
        p += fifo[c][sd].size() * sizeof(float);
      }
    if (h.state_doubles) {
      PEAQ_CUDA(cudaSetDevice(device));
      PEAQ_CUDA(cudaStreamSynchronize(engine->stream));
      PEAQ_CUDA(cudaMemcpy(p, engine->d_state, h.state_doubles * sizeof(double), cudaMemcpyDeviceToHost));
      p += h.state_doubles * sizeof(double);
      if (h.hp_doubles) {
        PEAQ_CUDA(cudaMemcpy(p, engine->d_hp_state, h.hp_doubles * sizeof(double), cudaMemcpyDeviceToHost));
        p += h.hp_doubles * sizeof(double);
      }
    }
    std::memcpy(p, &last, sizeof(PairResult));
    return 0;
  }

  int restore(const void* buf, size_t size) {
    if (size < sizeof(SnapHeader)) return fail(PEAQ_B200_ERR_INVALID, "not a session snapshot");
    SnapHeader h;
    std::memcpy(&h, buf, sizeof h);
    if (h.magic != kSnapMagic || h.version != 1) return fail(PEAQ_B200_ERR_INVALID, "not a session snapshot");
    if (h.channels < 0 || h.channels > kMaxChannels) return fail(PEAQ_B200_ERR_INVALID, "corrupt snapshot");
    size_t total = sizeof h + (h.state_doubles + h.hp_doubles) * sizeof(double) + sizeof(PairResult);
    for (int c = 0; c < 2; c++)
      for (int sd = 0; sd < 2; sd++) total += h.fifo_floats[c][sd] * sizeof(float);
    if (size < total) return fail(PEAQ_B200_ERR_INVALID, "truncated snapshot");
    drop_engine();
    advanced = h.advanced != 0;
    level = h.level;
    channels = h.channels;
    const unsigned char* p = static_cast<const unsigned char*>(buf) + sizeof h;
    for (int c = 0; c < 2; c++)
      for (int sd = 0; sd < 2; sd++) {
        fifo[c][sd].append(reinterpret_cast<const float*>(p), h.fifo_floats[c][sd]);
        p += h.fifo_floats[c][sd] * sizeof(float);
      }
    if (h.started) {
      int rc = ensure_engine();
      if (rc) return rc;
      PEAQ_CUDA(cudaSetDevice(device));
      started = true;
      engine->fb_pos = h.fb_pos;
      if (h.state_doubles != state_doubles() || h.hp_doubles != hp_doubles()) {
        started = false;
        return fail(PEAQ_B200_ERR_INVALID, "snapshot of another engine version");
      }
      if ((rc = engine->ensure_pairs(1, h.state_doubles, false))) return rc;
      PEAQ_CUDA(cudaMemcpy(engine->d_state, p, h.state_doubles * sizeof(double), cudaMemcpyHostToDevice));
      p += h.state_doubles * sizeof(double);
      if (h.hp_doubles) {
        if ((rc = engine->ensure(&engine->d_hp_state, &engine->hp_state_cap, (size_t)h.hp_doubles))) return rc;
        PEAQ_CUDA(cudaMemcpy(engine->d_hp_state, p, h.hp_doubles * sizeof(double), cudaMemcpyHostToDevice));
        p += h.hp_doubles * sizeof(double);
      }
    }
    std::memcpy(&last, p, sizeof(PairResult));
    have_result = h.have_result != 0;
    result_stale = false;
    return 0;
  }
};

}  // namespace peaq

using namespace peaq;

extern "C" {

const char* peaq_b200_last_error(void) { return g_last_error.c_str(); }
const char* peaq_b200_version(void) { return "peaq-b200 1"; }

int peaq_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int peaq_b200_engine_create(peaq_b200_engine** out, int device, int advanced, double playback_level) {
  if (!out) return fail(PEAQ_B200_ERR_INVALID, "null argument");
  *out = nullptr;
  if (peaq_b200_device_count() <= device || device < 0)
    return fail(PEAQ_B200_ERR_CUDA, "no such CUDA device (the engine has no CPU fallback)");
  Engine* e = new (std::nothrow) Engine;
  if (!e) return fail(PEAQ_B200_ERR_NOMEM, "out of host memory");
  e->device = device;
  e->advanced = advanced != 0;
  e->level = playback_level;
  int rc = e->init();
  if (rc) {
    e->destroy();
    delete e;
    return rc;
  }
  *out = reinterpret_cast<peaq_b200_engine*>(e);
  return 0;
}

int peaq_b200_engine_destroy(peaq_b200_engine* h) {
  if (!h) return 0;
  Engine* e = reinterpret_cast<Engine*>(h);
  e->destroy();
  delete e;
  return 0;
}

int peaq_b200_engine_run_batch(peaq_b200_engine* h, const peaq_b200_batch* batch, peaq_b200_result* out) {
  return run_batch(reinterpret_cast<Engine*>(h), batch, out);
}

int peaq_b200_synth_pairs(int device, float* ref, float* test, size_t pair_stride, int32_t n_pairs,
                          uint64_t first_pair, uint64_t n_samples, int32_t channels) {
  if (!ref || !test) return fail(PEAQ_B200_ERR_INVALID, "null argument");
  int rc = check_channels(channels);
  if (rc) return rc;
  if (n_pairs > 1 && n_samples * channels > pair_stride)
    return fail(PEAQ_B200_ERR_INVALID, "pair_stride smaller than an item");
  if (device < 0) {
    static int16_t* table = nullptr;
    if (!table) {
      int16_t* t = new int16_t[PEAQ_SYNTH_TABLE_SIZE];
      for (int i = 0; i < PEAQ_SYNTH_TABLE_SIZE; i++) t[i] = (int16_t)peaq_synth_sine_entry((uint32_t)i);
      table = t;
    }
    for (int32_t p = 0; p < n_pairs; p++) {
      PeaqSynthPair sp;
      peaq_synth_pair_init(&sp, first_pair + p);
      float* r = ref + (size_t)p * pair_stride;
      float* t = test + (size_t)p * pair_stride;
      for (uint64_t n = 0; n < n_samples; n++)
        for (int c = 0; c < channels; c++) {
          int32_t rv, tv;
          peaq_synth_sample(&sp, table, n, c, &rv, &tv);
          r[n * channels + c] = (float)rv / 32768.0f;
          t[n * channels + c] = (float)tv / 32768.0f;
        }
    }
    return 0;
  }
  PEAQ_CUDA(cudaSetDevice(device));
  PEAQ_CUDA(launch_synth_pairs(ref, test, pair_stride, n_pairs, first_pair, n_samples, channels, 0));
  PEAQ_CUDA(cudaDeviceSynchronize());
  return 0;
}

int peaq_b200_device_alloc(int device, size_t bytes, void** out) {
  if (!out) return fail(PEAQ_B200_ERR_INVALID, "null argument");
  PEAQ_CUDA(cudaSetDevice(device));
  PEAQ_CUDA(cudaMalloc(out, bytes));
  return 0;
}
int peaq_b200_device_free(int device, void* ptr) {
  PEAQ_CUDA(cudaSetDevice(device));
  PEAQ_CUDA(cudaFree(ptr));
  return 0;
}
int peaq_b200_memcpy_h2d(int device, void* dst, const void* src, size_t bytes) {
  PEAQ_CUDA(cudaSetDevice(device));
  PEAQ_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
  return 0;
}
int peaq_b200_memcpy_d2h(int device, void* dst, const void* src, size_t bytes) {
  PEAQ_CUDA(cudaSetDevice(device));
  PEAQ_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return 0;
}
int peaq_b200_host_alloc_pinned(size_t bytes, void** out) {
  if (!out) return fail(PEAQ_B200_ERR_INVALID, "null argument");
  PEAQ_CUDA(cudaMallocHost(out, bytes));
  return 0;
}
int peaq_b200_host_free_pinned(void* ptr) {
  PEAQ_CUDA(cudaFreeHost(ptr));
  return 0;
}

double peaq_b200_engine_last_ms(const peaq_b200_engine* h, int which) {
  const Engine* e = reinterpret_cast<const Engine*>(h);
  if (!e || which < 0 || which > 6) return -1.;
  if (which == 4) return e->ms[7] + e->ms[5] + e->ms[6];
  return e->ms[which];
}

namespace {
// 8 independent DFMA chains per thread: with >= 16 warps per scheduler the FP64 pipe is the only limit
__global__ void fp64_peak_kernel(double* out, int iters, double b, double c) {
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = (double)(threadIdx.x + k);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = fma(a[k], b, c);
  }
  double s = 0.;
#pragma unroll
  for (int k = 0; k < 8; k++) s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

double peaq_b200_fp64_peak_tflops(int device) {
  if (cudaSetDevice(device) != cudaSuccess) return -1.;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
  double* d = nullptr;
  if (cudaMalloc(&d, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return -1.;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  double best = -1.;
  for (int rep = 0; rep < 4; rep++) {   // first repetition warms up
    cudaEventRecord(a);
    fp64_peak_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(b);
    if (cudaEventSynchronize(b) != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double tf = 2. * 8. * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d);
  return best;
}

uint64_t peaq_b200_engine_launch_count(const peaq_b200_engine* h) {
  const Engine* e = reinterpret_cast<const Engine*>(h);
  return e ? e->launches : 0;
}

int peaq_b200_engine_keep_records(peaq_b200_engine* h, int enable) {
  if (!h) return fail(PEAQ_B200_ERR_INVALID, "null argument");
  reinterpret_cast<Engine*>(h)->keep_records = enable != 0;
  return 0;
}

int peaq_b200_engine_record_layout(const peaq_b200_engine* h, int32_t* out) {
  if (!h || !out) return fail(PEAQ_B200_ERR_INVALID, "null argument");
  const RecordLayout& L = reinterpret_cast<const Engine*>(h)->last_layout;
  out[0] = L.C; out[1] = L.B; out[2] = L.off_noise; out[3] = L.off_ehs; out[4] = L.off_snr;
  out[5] = L.off_ints; out[6] = L.stride; out[7] = 0; out[8] = 0;
  return 0;
}

int peaq_b200_engine_copy_records(peaq_b200_engine* h, double* dst, size_t max_doubles, size_t* n_doubles) {
  if (!h || !dst) return fail(PEAQ_B200_ERR_INVALID, "null argument");
  Engine* e = reinterpret_cast<Engine*>(h);
  const size_t n = std::min(max_doubles, e->last_records_doubles);
  PEAQ_CUDA(cudaSetDevice(e->device));
  if (n) PEAQ_CUDA(cudaMemcpy(dst, e->d_records, n * sizeof(double), cudaMemcpyDeviceToHost));
  if (n_doubles) *n_doubles = n;
  return 0;
}

int peaq_b200_engine_copy_fb_debug(peaq_b200_engine* h, double* dst, size_t max_doubles, size_t* n_doubles,
                                   uint32_t* frames) {
  if (!h || !dst) return fail(PEAQ_B200_ERR_INVALID, "null argument");
  Engine* e = reinterpret_cast<Engine*>(h);
  const size_t n = std::min(max_doubles, e->last_fbdbg_doubles);
  PEAQ_CUDA(cudaSetDevice(e->device));
  if (n) PEAQ_CUDA(cudaMemcpy(dst, e->d_fbdbg, n * sizeof(double), cudaMemcpyDeviceToHost));
  if (n_doubles) *n_doubles = n;
  if (frames) *frames = e->last_fb_frames;
  return 0;
}

int peaq_b200_segment_plan(uint64_t n_samples, uint64_t* segment_samples, uint64_t* warmup_samples) {
  uint64_t len = 0;
  const unsigned k = Engine::segments_for_samples(n_samples, &len);
  if (segment_samples) *segment_samples = len;
  if (warmup_samples) *warmup_samples = kSegWarmSamples;
  return (int)k;
}

int peaq_b200_table(int advanced, double playback_level, int model, int which, double* out) {
  if (!out) return fail(PEAQ_B200_ERR_INVALID, "null argument");
  DeviceTables* tp = new (std::nothrow) DeviceTables;
  if (!tp) return fail(PEAQ_B200_ERR_NOMEM, "out of host memory");
  build_tables(tp, advanced != 0, playback_level);
  struct Guard { DeviceTables* p; ~Guard() { delete p; } } guard{tp};
  const DeviceTables& t = *tp;
  if (model == 2) {
    // filter-bank tables of band `which`: N, D, the recursion coefficients [32][6] complex and
    // rotations [3][6] complex (fb_bank_rec_kernel), then the reference-form taps re[0..N/2],
    // im[0..N/2] (fbearmodel.c:213-220) -- up to 1880 doubles
    if (which < 0 || which >= kFbBands) return fail(PEAQ_B200_ERR_INVALID, "no such band");
    const int N = t.fb_len[which];
    int n = 0;
    out[n++] = N;
    out[n++] = 1 + (kFbBuf - N) / 2;
    if (which < kFbRecBands) {
      for (int k = 0; k < 32; k++)
        for (int f = 0; f < 6; f++) {
          out[n++] = t.fb_rec_ph[which][k][f].x;
          out[n++] = t.fb_rec_ph[which][k][f].y;
        }
      for (int f = 0; f < 3; f++)
        for (int i = 0; i < kFbRecGroup; i++) {
          out[n++] = t.fb_rec_rpow[which][f][i].x;
          out[n++] = t.fb_rec_rpow[which][f][i].y;
        }
    }
    for (int i = 0; i <= N / 2; i++) out[n++] = t.fb_h_re[t.fb_tap_offset[which] + i];
    for (int i = 0; i <= N / 2; i++) out[n++] = t.fb_h_im[t.fb_tap_offset[which] + i];
    return n;
  }
  const BandTables& b = model ? t.fb : t.fft;
  for (int i = 0; i < b.B; i++) {
    double v;
    switch (which) {
      case 0: v = b.fc[i]; break;
      case 1: v = b.internal_noise[i]; break;
      case 2: v = b.a_ear[i]; break;
      case 3: v = b.ethres[i]; break;
      case 4: v = b.thres[i]; break;
      case 5: v = b.loudfac[i]; break;
      case 14: v = b.a_proc[i]; break;
      default:
        if (model) return fail(PEAQ_B200_ERR_INVALID, "no such table");
        switch (which) {
          case 6: v = t.maskdiff[i]; break;
          case 7: v = t.aUC[i]; break;
          case 8: v = t.gIL[i]; break;
          case 9: v = t.spread_norm[i]; break;
          case 10: v = t.band_wl[i]; break;
          case 11: v = t.band_wu[i]; break;
          case 12: v = t.band_lo[i]; break;
          case 13: v = t.band_hi[i]; break;
          default: return fail(PEAQ_B200_ERR_INVALID, "no such table");
        }
    }
    out[i] = v;
  }
  return b.B;
}

int peaq_b200_session_create(peaq_b200_session** out, int device) {
  if (!out) return fail(PEAQ_B200_ERR_INVALID, "null argument");
  *out = nullptr;
  if (peaq_b200_device_count() <= device || device < 0)
    return fail(PEAQ_B200_ERR_CUDA, "no such CUDA device (the engine has no CPU fallback)");
  Session* s = new (std::nothrow) Session;
  if (!s) return fail(PEAQ_B200_ERR_NOMEM, "out of host memory");
  s->device = device;
  *out = reinterpret_cast<peaq_b200_session*>(s);
  return 0;
}

int peaq_b200_session_destroy(peaq_b200_session* h) {
  if (!h) return 0;
  Session* s = reinterpret_cast<Session*>(h);
  s->drop_engine();
  delete s;
  return 0;
}

int peaq_b200_session_set_advanced(peaq_b200_session* h, int advanced) {
  if (!h) return fail(PEAQ_B200_ERR_INVALID, "null argument");
  Session* s = reinterpret_cast<Session*>(h);
  s->drop_engine();
  s->advanced = advanced != 0;
  return 0;
}

int peaq_b200_session_set_playback_level(peaq_b200_session* h, double level_db) {
  if (!h) return fail(PEAQ_B200_ERR_INVALID, "null argument");
  if (!(level_db >= 0. && level_db <= 130.)) return fail(PEAQ_B200_ERR_INVALID, "playback level out of range");
  Session* s = reinterpret_cast<Session*>(h);
  if (s->engine) {
    // like the reference (gstpeaq.c:509-514): takes effect from the next frame, state is kept
    int rc = s->sync();
    if (rc) return rc;
    if ((rc = s->engine->set_level(level_db))) return rc;
  }
  s->level = level_db;
  return 0;
}

int peaq_b200_session_get_playback_level(const peaq_b200_session* h, double* level_db) {
  if (!h || !level_db) return fail(PEAQ_B200_ERR_INVALID, "null argument");
  *level_db = reinterpret_cast<const Session*>(h)->level;
  return 0;
}

int peaq_b200_session_set_channels(peaq_b200_session* h, int channels) {
  if (!h) return fail(PEAQ_B200_ERR_INVALID, "null argument");
  int rc = check_channels(channels);
  if (rc) return rc;
  Session* s = reinterpret_cast<Session*>(h);
  // set_caps frees and reallocates all per-channel state (gstpeaq.c:575-586)
  s->reset_stream();
  s->channels = channels;
  return 0;
}

int peaq_b200_session_push(peaq_b200_session* h, int pad, const float* data, size_t n) {
  if (!h || (n && !data)) return fail(PEAQ_B200_ERR_INVALID, "null argument");
  if (pad != PEAQ_B200_PAD_REF && pad != PEAQ_B200_PAD_TEST) return fail(PEAQ_B200_ERR_INVALID, "bad pad");
  Session* s = reinterpret_cast<Session*>(h);
  if (s->channels == 0) return fail(PEAQ_B200_ERR_INVALID, "channels not negotiated");
  s->fifo[0][pad].append(data, n * s->channels);
  if (s->advanced) s->fifo[1][pad].append(data, n * s->channels);
  return s->drain();
}

int peaq_b200_session_finish(peaq_b200_session* h) {
  if (!h) return fail(PEAQ_B200_ERR_INVALID, "null argument");
  Session* s = reinterpret_cast<Session*>(h);
  if (s->channels == 0) return fail(PEAQ_B200_ERR_INVALID, "channels not negotiated");
  int rc = s->drain();
  if (rc) return rc;
  if ((rc = s->flush())) return rc;
  return s->fetch_result();
}

int peaq_b200_session_get_result(peaq_b200_session* h, peaq_b200_result* out) {
  if (!h || !out) return fail(PEAQ_B200_ERR_INVALID, "null argument");
  Session* s = reinterpret_cast<Session*>(h);
  if (!s->have_result) {
    // no frame processed yet: the reference would evaluate empty accumulators
    // (0/0); report that state without touching the GPU
    std::memset(out, 0, sizeof *out);
    const double nanv = std::nan("");
    out->odg = out->di = out->totalsnr = nanv;
    out->n_movs = s->advanced ? 5 : 11;
    for (int i = 0; i < out->n_movs; i++) out->movs[i] = nanv;
    out->loudness_reached_frame = UINT_MAX;
    return 0;
  }
  int rc = s->fetch_result();   // waits for the pushes queued so far
  if (rc) return rc;
  std::memcpy(out, &s->last, sizeof *out);
  return 0;
}

int peaq_b200_session_snapshot(peaq_b200_session* h, void* buf, size_t capacity, size_t* size) {
  if (!h) return fail(PEAQ_B200_ERR_INVALID, "null argument");
  return reinterpret_cast<Session*>(h)->snapshot(buf, capacity, size);
}

int peaq_b200_session_restore(peaq_b200_session* h, const void* buf, size_t size) {
  if (!h || !buf) return fail(PEAQ_B200_ERR_INVALID, "null argument");
  return reinterpret_cast<Session*>(h)->restore(buf, size);
}

// ---------------------------------------------------------------------------
// multi-GPU batch: one engine and one host thread per device, contiguous blocks of pairs

struct peaq_b200_multi {
  std::vector<Engine*> engines;
};

int peaq_b200_multi_create(peaq_b200_multi** out, const int* devices, int n_devices, int advanced,
                           double playback_level) {
  if (!out) return fail(PEAQ_B200_ERR_INVALID, "null argument");
  *out = nullptr;
  const int visible = peaq_b200_device_count();
  if (n_devices <= 0) n_devices = visible;
  if (n_devices <= 0) return fail(PEAQ_B200_ERR_CUDA, "no CUDA device (the engine has no CPU fallback)");
  peaq_b200_multi* m = new (std::nothrow) peaq_b200_multi;
  if (!m) return fail(PEAQ_B200_ERR_NOMEM, "out of host memory");
  for (int i = 0; i < n_devices; i++) {
    const int dev = devices ? devices[i] : i;
    peaq_b200_engine* e = nullptr;
    int rc = peaq_b200_engine_create(&e, dev, advanced, playback_level);
    if (rc) {
      peaq_b200_multi_destroy(m);
      return rc;
    }
    m->engines.push_back(reinterpret_cast<Engine*>(e));
  }
  *out = m;
  return 0;
}

int peaq_b200_multi_destroy(peaq_b200_multi* m) {
  if (!m) return 0;
  for (Engine* e : m->engines) peaq_b200_engine_destroy(reinterpret_cast<peaq_b200_engine*>(e));
  delete m;
  return 0;
}

int peaq_b200_multi_device_count(const peaq_b200_multi* m) { return m ? (int)m->engines.size() : 0; }

int peaq_b200_multi_run_batch(peaq_b200_multi* m, const peaq_b200_batch* b, peaq_b200_result* out) {
  if (!m || !b || !out) return fail(PEAQ_B200_ERR_INVALID, "null argument");
  if (b->on_device) return fail(PEAQ_B200_ERR_INVALID, "the multi-GPU batch takes host buffers");
  if (b->n_pairs <= 0) return fail(PEAQ_B200_ERR_INVALID, "n_pairs must be positive");
  const int n_dev = (int)m->engines.size();
  std::vector<int> rcs(n_dev, 0);
  std::vector<std::string> errs(n_dev);
  std::vector<std::thread> workers;
  const int base = b->n_pairs / n_dev, rem = b->n_pairs % n_dev;
  int first = 0;
  for (int d = 0; d < n_dev; d++) {
    const int count = base + (d < rem ? 1 : 0);
    if (count > 0) {
      peaq_b200_batch sub = *b;
      sub.n_pairs = count;
      // a single pair is addressed without a stride (see run_batch): hand over explicit pointers
      const size_t stride = b->n_pairs > 1 ? b->pair_stride : 0;
      sub.ref = b->ref + (size_t)first * stride;
      sub.test = b->test + (size_t)first * stride;
      if (b->n_samples) sub.n_samples = b->n_samples + first;
      if (count == 1 && b->n_pairs > 1) sub.pair_stride = b->pair_stride;
      peaq_b200_result* dst = out + first;
      workers.emplace_back([&, d, sub, dst]() {
        rcs[d] = run_batch(m->engines[d], &sub, dst);   // the D2H of this device's rows IS the gather
        if (rcs[d]) errs[d] = g_last_error;              // thread-local: hand it to the caller's thread
      });
    }
    first += count;
  }
  for (auto& w : workers) w.join();
  for (int d = 0; d < n_dev; d++)
    if (rcs[d]) return fail(rcs[d], "device " + std::to_string(m->engines[d]->device) + ": " + errs[d]);
  return 0;
}

}  // extern "C"
