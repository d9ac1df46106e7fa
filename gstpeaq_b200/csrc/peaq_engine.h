// Internal C++ interface of the PEAQ CUDA engine (not part of the C ABI).
//
// Data flow (DESIGN.md has the long version):
//   PCM in HBM --K1 fft_frames--> per-frame records --K2 scan_basic--> pair state/result
// K1 is frame-parallel and stateless; K2 carries every recurrence of the
// reference (time smearing, level adapter, modulation, accumulators) in
// registers across its frame loop and persists it in `PairState` memory between
// chunks, so hour-long items and streaming sessions run as a sequence of chunks.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "peaq_tables.h"

namespace peaq {

constexpr int kMaxChannels = 2;
constexpr int kNumAcc = 11;        // accumulator slots (11 basic MOVs; 5 used in advanced)
constexpr int kAccFields = 8;      // num, den, x0, x1, x2, saved num, saved den, saved max
constexpr int kBandStateFields = 14;
constexpr int kHpHdr = 24;   // DC-reject scan state per stream, see peaq_fb.cu (FB1)
constexpr int kHpStateDoubles = kHpHdr + kFbHist + 2 * 3 * kFbRecBands;   // + FIR history + filter-bank chain values (even: 16-byte rows)

// Layout of one per-frame record written by K1 and read by K2 (units: doubles).
struct RecordLayout {
  int C;           // channels
  int B;           // FFT bands
  int off_noise;   // noise_in_bands[c][b]
  int off_ehs;     // ehs[c]
  int off_snr;     // signal, noise partial sums of this frame
  int off_ints;    // int32: flags, then (bw_ref, bw_test) per channel
  int stride;      // doubles per record
};

inline RecordLayout make_record_layout(int C, int B) {
  RecordLayout L;
  L.C = C;
  L.B = B;
  L.off_noise = 2 * C * B;
  L.off_ehs = L.off_noise + C * B;
  L.off_snr = L.off_ehs + C;
  L.off_ints = L.off_snr + 2;
  L.stride = L.off_ints + (1 + 2 * C + 1) / 2;
  return L;
}

constexpr int kRecFlagAbove = 1;     // is_frame_above_threshold (gstpeaq.c:1081-1099)
constexpr int kRecFlagEhsValid = 2;  // any energy flag set (movs.c:1375-1381)

// Per-pair recurrent state, basic mode (units: doubles unless noted).
struct StateLayout {
  int C, B;
  int off_band;    // [field][c][b], kBandStateFields fields
  int off_acc;     // [c][kNumAcc][kAccFields]
  int off_scalar;  // signal energy, noise energy
  int off_ints;    // int32: status, frame counter, loudness-reached frame, spare | segment words (kSeg*)
  int stride;
};

// Segment words of the basic state (int32 slots after the three counters; zero for whole items,
// see peaq_segments.cu): first frame whose contributions count, "an owned frame was above the
// threshold", 1 + first frame above the threshold (0: none yet)
constexpr int kSegAccStart = 4, kSegOwnedAbove = 5, kSegFirstAbove = 6;
// the same for the two clocks of the advanced state
constexpr int kASegAccStartFft = 5, kASegAccStartFb = 6, kASegOwnedAboveFft = 7, kASegOwnedAboveFb = 8,
              kASegFirstAboveFft = 9, kASegFirstAboveFb = 10;

inline StateLayout make_state_layout(int C, int B) {
  StateLayout S;
  S.C = C;
  S.B = B;
  S.off_band = 0;
  S.off_acc = kBandStateFields * C * B;
  S.off_scalar = S.off_acc + C * kNumAcc * kAccFields;
  S.off_ints = S.off_scalar + 2;
  S.stride = S.off_ints + 4;
  return S;
}

// Per-pair recurrent state, advanced mode (doubles unless noted).
struct AdvStateLayout {
  int C;
  int off_fft_filtered;   // [c][55]            time smearing of the ref excitation
  int off_fft_acc;        // [c][2][kAccFields] SegmentalNMR, EHS
  int off_fft_scalar;     // signal / noise energy
  int off_fb_stream;      // [2C streams][cu 40 | excitation 40 | last sub-step energies, newest first, 11 x 40 (5 used)]
  int off_fb_level;       // [c][6][40]
  int off_fb_mod;         // [c][ref|test][3][40]
  int off_fb_acc;         // [c][3][kAccFields] RmsModDiff, RmsNoiseLoudAsym, AvgLinDist
  int off_fb_movs;        // 3 channel-averaged MOV values published by the fb scan
  int off_ints;           // int32: fft status, fft frames, fb status, fb frames, loudness frame | segment words (kASeg*)
  int stride;
};

inline AdvStateLayout make_adv_state_layout(int C) {
  AdvStateLayout S;
  S.C = C;
  S.off_fft_filtered = 0;
  S.off_fft_acc = C * 55;
  S.off_fft_scalar = S.off_fft_acc + C * 2 * kAccFields;
  S.off_fb_stream = S.off_fft_scalar + 2;
  S.off_fb_level = S.off_fb_stream + 2 * C * 13 * kFbBands;
  S.off_fb_mod = S.off_fb_level + C * 6 * kFbBands;
  S.off_fb_acc = S.off_fb_mod + C * 2 * 3 * kFbBands;
  S.off_fb_movs = S.off_fb_acc + C * 3 * kAccFields;
  S.off_ints = S.off_fb_movs + 3;
  S.stride = S.off_ints + 6;
  return S;
}

// Mirrors peaq_b200_result of the C ABI (include/peaq_b200.h).
struct PairResult {
  double odg;
  double di;
  double totalsnr;
  double movs[11];
  int32_t n_movs;
  uint32_t frames_fft;
  uint32_t frames_fb;
  uint32_t loudness_reached_frame;
};

// View of resident PCM: pair p's ref signal starts at ref + p*pair_stride
// floats, interleaved [sample][channel]; samples >= n_samples[p] read as zero.
struct PcmView {
  const float* ref;
  const float* test;
  size_t pair_stride;
  const unsigned long long* n_samples;       // device, per pair: length of the ref signal
  const unsigned long long* n_samples_test;  // device, per pair: length of the test signal
  const unsigned* n_frames;                  // device, per pair: frames of this clock to run
  int channels;
  // device, per pair, or null: where the pair's signals start, in floats from ref / test (segments
  // of long items are windows into their item); null: pair * pair_stride
  const unsigned long long* base = nullptr;
};

__host__ __device__ inline size_t pcm_pair_offset(const PcmView& pcm, int pair) {
  return pcm.base ? (size_t)pcm.base[pair] : (size_t)pair * pcm.pair_stride;
}

struct LaunchStats {
  unsigned long long launches;
};

// K0: synthetic (ref,test) pairs generated in place (bench/test input only).
cudaError_t launch_synth_pairs(float* ref, float* test, size_t pair_stride, int n_pairs,
                               unsigned long long first_pair_index, unsigned long long n_samples,
                               int channels, cudaStream_t stream);

// K1: frames [first_frame, first_frame + n_chunk_frames) of every pair.
cudaError_t launch_fft_frames(const DeviceTables* d_tables, PcmView pcm, int n_pairs,
                              unsigned first_frame, unsigned n_chunk_frames, double* records,
                              RecordLayout L, int fft_bands, bool advanced, cudaStream_t stream);

// K2 (basic): consumes the records of one chunk, updates the state, writes results.
cudaError_t launch_scan_basic(const DeviceTables* d_tables, const double* records, RecordLayout L,
                              const unsigned* n_frames, unsigned first_frame,
                              unsigned n_chunk_frames, double* state, StateLayout S,
                              PairResult* results, int n_pairs, cudaStream_t stream,
                              double* dbg = nullptr /* test tap: [pair][chunk frame][scan_tap_doubles_per_frame] */);
int scan_tap_doubles_per_frame(int C, int B);

size_t fft_frames_smem_bytes(int channels);

// Fused persistent kernel (basic mode): K1 + K2 of one pair in one CTA, frames
// [first_frame, first_frame + n_chunk_frames); same state block and results as K1 + K2.
cudaError_t launch_fused_basic(const DeviceTables* d_tables, PcmView pcm, int n_pairs, unsigned first_frame,
                               unsigned n_chunk_frames, double* state, StateLayout S, PairResult* results,
                               cudaStream_t stream);

// advanced mode (peaq_fb.cu, peaq_scan_adv.cu)
cudaError_t launch_fb_flags(PcmView pcm, int n_pairs, unsigned first_frame, unsigned n_chunk_frames,
                            unsigned char* flags, cudaStream_t stream);
// t0: position of the chunk in the PCM buffers; pos0: its absolute position in the item (they
// differ for streaming sessions, whose buffers hold only the newest samples)
cudaError_t launch_fb_hp(const DeviceTables* d_tables, PcmView pcm, int n_pairs,
                         unsigned long long t0, unsigned long long pos0, unsigned chunk_samples,
                         double* hp, size_t hp_stride, double* hp_state, bool first_chunk,
                         cudaStream_t stream);
cudaError_t launch_fb_bank(const DeviceTables* d_tables, const DeviceTables* h_tables,
                           const double* hp, size_t hp_stride, int n_streams, unsigned n_sub,
                           double* fbout, double* hp_state, bool first_chunk, bool direct_only,
                           const unsigned* n_frames /* device, per pair: frames of the filter-bank clock */,
                           unsigned first_frame, int streams_per_pair, cudaStream_t stream);
cudaError_t launch_init_adv_state(double* state, AdvStateLayout S, int n_pairs, cudaStream_t stream);
cudaError_t launch_fb_spread(const DeviceTables* d_tables, const double* fbout, unsigned n_sub,
                             const unsigned* n_frames, unsigned first_frame, double* state,
                             AdvStateLayout S, double* energy, int n_pairs, cudaStream_t stream);
cudaError_t launch_fb_scan(const DeviceTables* d_tables, const double* energy, unsigned n_sub,
                           const unsigned char* flags, const unsigned* n_frames, unsigned first_frame,
                           unsigned n_chunk_frames, double* state, AdvStateLayout S, double* dbg,
                           int n_pairs, cudaStream_t stream);
cudaError_t launch_adv_fft_scan(const DeviceTables* d_tables, const double* records, RecordLayout L,
                                const unsigned* n_frames, unsigned first_frame, unsigned n_chunk_frames,
                                double* state, AdvStateLayout S, PairResult* results, int n_pairs,
                                cudaStream_t stream);

// ---- segments of long items (peaq_segments.cu) -------------------------------------------------
constexpr unsigned long long kSegWarmSamples = 196608;          // 4.096 s = 192 FFT-clock frames = 1024 filter-bank frames
constexpr unsigned long long kSegSamples = 8 * kSegWarmSamples;   // 32.8 s = 1536 / 8192 frames; multiples of the DC-reject scan's blocks

// device arrays describing the virtual pairs (one per segment) of a batch and their items
struct SegTable {
  const int* seg_index;            // [vp] 0 for an item's first segment
  const unsigned* frame0_fft;      // [vp] absolute index of the first frame the vp runs (its warm-up start)
  const unsigned* acc_start_fft;   // [vp] absolute index of the first frame the vp owns
  const unsigned* frame0_fb;
  const unsigned* acc_start_fb;
  const int* first_vp;             // [item]
  const int* n_seg;                // [item]
};
cudaError_t launch_seg_init(double* state, const StateLayout* S, const AdvStateLayout* A, int n_vp,
                            const SegTable& t, cudaStream_t stream);
// sums of the segments -> segment 0's state block; redo[item] = 1 when the item must be run as a whole
cudaError_t launch_seg_combine(double* state, const StateLayout* S, const AdvStateLayout* A, int n_items,
                               const SegTable& t, unsigned char* redo, cudaStream_t stream);
cudaError_t launch_seg_gather_results(const PairResult* res, const int* first_vp, int n_items, PairResult* out,
                                      cudaStream_t stream);

}  // namespace peaq
