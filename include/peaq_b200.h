/* peaq_b200.h -- C ABI of the B200-native PEAQ engine (libpeaq_b200.so).
 *
 * Drop-in boundary for the per-frame hot path of HSU-ANT/gstpeaq.  Each entry
 * point names the reference interface it replaces (paths relative to the
 * reference's src/).  Plain pointers and sizes only; all device memory is
 * owned by the library.  Every function returns 0 on success or a negative
 * peaq_b200_status; peaq_b200_last_error() gives the message of the calling
 * thread's last failure.  There is no CPU fallback: without a CUDA device the
 * calls fail with PEAQ_B200_ERR_CUDA.
 *
 * INTEGRATION.md shows how the GStreamer element (gstpeaq.c) and the CLI
 * (peaq.c) bind to these functions.
 */
#ifndef PEAQ_B200_H
#define PEAQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  PEAQ_B200_OK = 0,
  PEAQ_B200_ERR_INVALID = -1,   /* bad argument / call order */
  PEAQ_B200_ERR_CUDA = -2,      /* CUDA runtime failure or no device */
  PEAQ_B200_ERR_NOMEM = -3
} peaq_b200_status;

/* Result properties of the element: `odg`, `di`, `totalsnr` (gstpeaq.c:484-497)
 * and the model output variables behind them in the order of the enums at
 * gstpeaq.c:86-108 (11 in basic mode, 5 in advanced mode). */
typedef struct {
  double odg;
  double di;
  double totalsnr;
  double movs[11];
  int32_t n_movs;
  uint32_t frames_fft;              /* frame_counter     (gstpeaq.c:124) */
  uint32_t frames_fb;               /* frame_counter_fb  (gstpeaq.c:125) */
  uint32_t loudness_reached_frame;  /* G_MAXUINT = not yet (gstpeaq.c:126) */
} peaq_b200_result;

const char *peaq_b200_last_error(void);
/* library / interface version, "peaq-b200 <n>" */
const char *peaq_b200_version(void);
/* number of visible CUDA devices (0 if none; never fails) */
int peaq_b200_device_count(void);

/* ------------------------------------------------------------------------
 * Engine: one per GPU and mode.  Owns tables, workspaces and a stream.
 * Replaces the model objects created in init() (gstpeaq.c:357-376) and
 * configured by set_property (gstpeaq.c:506-560).                          */
typedef struct peaq_b200_engine peaq_b200_engine;

int peaq_b200_engine_create(peaq_b200_engine **out, int device, int advanced,
                            double playback_level);
int peaq_b200_engine_destroy(peaq_b200_engine *e);

/* Batch of independent (ref,test) pairs -- the data-parallel entry that the
 * reference lacks (it runs one pair per element instance).  Results equal
 * running each pair through a session.
 *
 * Pair p's signals start at ref + p*pair_stride and test + p*pair_stride
 * (units: floats), interleaved F32 [sample][channel] as on the element's pads
 * (gstpeaq.c:146-152), n_samples[p] samples per channel (or n_samples_all for
 * every pair when n_samples is NULL).  Framing follows do_processing /
 * do_flush (gstpeaq.c:596-611, :716-745): 2048-sample frames every 1024
 * samples and one zero-padded frame for the remainder.
 * on_device != 0: ref/test are device pointers on the engine's GPU (must be
 * 16-byte aligned); otherwise host pointers that are copied in. */
typedef struct {
  int32_t n_pairs;
  int32_t channels;                 /* 1 or 2 */
  const float *ref;
  const float *test;
  size_t pair_stride;
  const uint64_t *n_samples;        /* host array [n_pairs] or NULL */
  uint64_t n_samples_all;
  int32_t on_device;
} peaq_b200_batch;

/* `out`: host array of n_pairs results. */
int peaq_b200_engine_run_batch(peaq_b200_engine *e, const peaq_b200_batch *batch,
                               peaq_b200_result *out);

/* Bench/test input: generates synthetic pairs [first_pair, first_pair+n_pairs)
 * (integer-only generator) into device buffers ref/test with
 * the layout described above.  With device < 0 the same pairs are produced on
 * the host into host buffers (bit-identical; lets the CPU oracle see the same
 * input). */
int peaq_b200_synth_pairs(int device, float *ref, float *test, size_t pair_stride,
                          int32_t n_pairs, uint64_t first_pair, uint64_t n_samples,
                          int32_t channels);

/* Device memory helpers so non-CUDA hosts (ctypes, cgo, JNI) can keep inputs
 * resident in HBM between calls. */
int peaq_b200_device_alloc(int device, size_t bytes, void **out);
int peaq_b200_device_free(int device, void *ptr);
int peaq_b200_memcpy_h2d(int device, void *dst, const void *src, size_t bytes);
int peaq_b200_memcpy_d2h(int device, void *dst, const void *src, size_t bytes);
int peaq_b200_host_alloc_pinned(size_t bytes, void **out);
int peaq_b200_host_free_pinned(void *ptr);

/* Timing and accounting of the engine's last run_batch (CUDA events on the
 * engine's stream).  which: 0 whole batch, 1 frame kernel(s), 2 scan
 * kernel(s), 3 host->device copies, 4 all filter-bank-clock kernels, 5 of
 * those the filter bank proper (fb_bank_rec_kernel), 6 the spreading and
 * 192-sample-clock scan kernels.  Milliseconds. */
double peaq_b200_engine_last_ms(const peaq_b200_engine *e, int which);
/* kernels launched by this engine since creation */
uint64_t peaq_b200_engine_launch_count(const peaq_b200_engine *e);
/* Measured FP64 FMA throughput of the device in TFLOP/s (a register-resident DFMA loop on every
 * SM, CUDA-event timed; < 0 on error).  The path is FP64 bound: this is the denominator the
 * kernels' achieved FMA rates are quoted against, next to the HBM roofline. */
double peaq_b200_fp64_peak_tflops(int device);

/* Debug taps for the parity tests: per-frame records of the last run_batch
 * (only kept when enabled before the run).  Layout: see
 * gstpeaq_b200/csrc/peaq_engine.h RecordLayout. */
int peaq_b200_engine_keep_records(peaq_b200_engine *e, int enable);
int peaq_b200_engine_record_layout(const peaq_b200_engine *e, int32_t *layout9);
int peaq_b200_engine_copy_records(peaq_b200_engine *e, double *dst, size_t max_doubles,
                                  size_t *n_doubles);
/* advanced mode, same switch: per 192-sample frame [pair][frame][stream][U|E][40]
 * (unsmeared / smeared filter-bank excitation; stream = 2*channel + (test?1:0)),
 * followed by [pair][frame][channel][8] = mod diff, temp weight, noise loudness,
 * missing components, lin dist, above-threshold flag, 0, 0.  *frames = frames kept.
 * Basic mode: the scan kernel's taps instead, [pair][frame]{[ref|test][channel][109] smeared
 * excitation, [channel][8] = mod diff 1, mod diff 2, temporal weight, noise loudness, mean N/M,
 * max N/M, binaural detection probability, binaural detection steps}. */
int peaq_b200_engine_copy_fb_debug(peaq_b200_engine *e, double *dst, size_t max_doubles,
                                   size_t *n_doubles, uint32_t *frames);
/* constant tables of the engine for a mode / playback level (host code, no
 * GPU needed; same `model`/`which` numbering as the oracle's
 * peaq_oracle_table); returns the count, < 0 on error.  model 2: the filter-bank
 * tables of band `which` -- N, D, recursion coefficients [32][6] and rotations [3][6]
 * (complex), then the taps re[0..N/2], im[0..N/2]; `out` must hold 1880 doubles */
int peaq_b200_table(int advanced, double playback_level, int model, int which, double *out);
/* How a batch item of n_samples per channel is run (host code, no GPU needed): the number of
 * segments it is cut into (1: not cut), their length and the warm-up every segment after the
 * first starts with, in samples.  Long items run as segments so that a handful of hour-long
 * pairs fill the GPU (gstpeaq_b200/csrc/peaq_segments.cu); the cut depends on n_samples alone.
 * Environment PEAQ_B200_SEGMENTS=0 switches segments off; sessions never use them. */
int peaq_b200_segment_plan(uint64_t n_samples, uint64_t *segment_samples, uint64_t *warmup_samples);

/* ------------------------------------------------------------------------
 * Session: one element instance (struct _GstPeaq, gstpeaq.c:110-139).
 * Externally serialised like pad_chain under GST_OBJECT_LOCK (gstpeaq.c:619).*/
typedef struct peaq_b200_session peaq_b200_session;

enum { PEAQ_B200_PAD_REF = 0, PEAQ_B200_PAD_TEST = 1 };

/* init(): basic mode, 92 dB SPL (gstpeaq.c:273-288, :357-376) */
int peaq_b200_session_create(peaq_b200_session **out, int device);
int peaq_b200_session_destroy(peaq_b200_session *s);
/* property `advanced` (gstpeaq.c:516-560); resets the per-channel state */
int peaq_b200_session_set_advanced(peaq_b200_session *s, int advanced);
/* property `playback_level` (gstpeaq.c:509-514): may change at any time; like in the reference the
 * new level applies from the next frame on and the models' state is kept */
int peaq_b200_session_set_playback_level(peaq_b200_session *s, double level_db);
int peaq_b200_session_get_playback_level(const peaq_b200_session *s, double *level_db);
/* CAPS event -> set_caps (gstpeaq.c:569-593); resets the per-channel state */
int peaq_b200_session_set_channels(peaq_b200_session *s, int channels);
/* pad_chain (gstpeaq.c:614-661): n = samples per channel; the buffer is
 * borrowed for the duration of the call */
int peaq_b200_session_push(peaq_b200_session *s, int pad, const float *interleaved, size_t n);
/* PAUSED->READY: do_flush + calculate_odg (gstpeaq.c:764-778) */
int peaq_b200_session_finish(peaq_b200_session *s);
/* properties odg / di / totalsnr, readable at any time (gstpeaq.c:484-497).  A push only
 * QUEUES its work on the GPU (the streaming thread is never stalled, like pad_chain returning
 * GST_FLOW_OK at once, gstpeaq.c:660); reading a result waits for the pushes queued so far. */
int peaq_b200_session_get_result(peaq_b200_session *s, peaq_b200_result *out);
/* Snapshot / restore of a running session: mode, playback level, channels, the samples waiting in
 * the adapters (gstpeaq.c:116-119), the recurrent state on the device and the last result.  A
 * restored session continues bit-identically (also on another device of the same kind).
 * snapshot: *size = bytes needed / written; with buf == NULL only the size is reported. */
int peaq_b200_session_snapshot(peaq_b200_session *s, void *buf, size_t capacity, size_t *size);
int peaq_b200_session_restore(peaq_b200_session *s, const void *buf, size_t size);

/* ------------------------------------------------------------------------
 * Multi-GPU batch: pairs are closed computations (all state is per element
 * instance, gstpeaq.c:110-139), so a batch shards into contiguous blocks of pairs, one engine and
 * one host thread per device; every device writes its result rows straight into the caller's
 * array -- that copy is the gather, no device-to-device collective is involved.
 * devices == NULL / n_devices <= 0: all visible devices.  Host buffers only. */
typedef struct peaq_b200_multi peaq_b200_multi;
int peaq_b200_multi_create(peaq_b200_multi **out, const int *devices, int n_devices, int advanced,
                           double playback_level);
int peaq_b200_multi_destroy(peaq_b200_multi *m);
int peaq_b200_multi_device_count(const peaq_b200_multi *m);
int peaq_b200_multi_run_batch(peaq_b200_multi *m, const peaq_b200_batch *batch, peaq_b200_result *out);

#ifdef __cplusplus
}
#endif
#endif /* PEAQ_B200_H */
