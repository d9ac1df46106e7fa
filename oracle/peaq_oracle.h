/* oracle/peaq_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, double precision, single thread) of the per-frame
 * PEAQ hot path of HSU-ANT/gstpeaq: interleaved F32 PCM of a (ref, test) pair
 * -> 11 (basic) / 5 (advanced) model output variables -> DI -> ODG.  It is the
 * checker for the CUDA engine in gstpeaq_b200/: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product never calls into oracle/.
 *
 * Parity pin: tests/test_oracle.py checks this restatement against
 *   - the reference's own code compiled from /root/reference/src
 *     (oracle/_ref/libpeaq_ref.so) on seeded inputs, per frame and end to end,
 *   - the reference's golden vectors (testpeaq.c:37-599) and known-answer ODGs
 *     (runtest-1.0.sh:18,28,38,48), committed as fixtures in tests/golden/.
 */
#ifndef PEAQ_ORACLE_H
#define PEAQ_ORACLE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PEAQ_ORACLE_MAX_BANDS 109
#define PEAQ_ORACLE_MAX_CHANNELS 8
#define PEAQ_ORACLE_TRACE_CHANNELS 2

typedef struct _PeaqOracle PeaqOracle;

/* Everything the engine is compared against per FFT-clock frame (first two
 * channels).  side 0 = ref, 1 = test. */
typedef struct
{
  int frame;
  int above_threshold;          /* gstpeaq.c:1081-1099 (ref signal) */
  int energy_flag[2][PEAQ_ORACLE_TRACE_CHANNELS]; /* fftearmodel.c:508-514 */
  int bw_ref[PEAQ_ORACLE_TRACE_CHANNELS];  /* movs.c:791-796; valid if > 346 */
  int bw_test[PEAQ_ORACLE_TRACE_CHANNELS]; /* movs.c:798-803 */
  int ehs_valid;                /* movs.c:1375-1381 */
  int pad_;
  double unsmeared[2][PEAQ_ORACLE_TRACE_CHANNELS][PEAQ_ORACLE_MAX_BANDS];
  double excitation[2][PEAQ_ORACLE_TRACE_CHANNELS][PEAQ_ORACLE_MAX_BANDS];
  double noise_in_bands[PEAQ_ORACLE_TRACE_CHANNELS][PEAQ_ORACLE_MAX_BANDS];
  double nmr[PEAQ_ORACLE_TRACE_CHANNELS];      /* linear mean N/M, movs.c:1013 */
  double nmr_max[PEAQ_ORACLE_TRACE_CHANNELS];
  double ehs[PEAQ_ORACLE_TRACE_CHANNELS];      /* before the factor 1000 */
  double mod_diff1[PEAQ_ORACLE_TRACE_CHANNELS];
  double mod_diff2[PEAQ_ORACLE_TRACE_CHANNELS];
  double temp_wt[PEAQ_ORACLE_TRACE_CHANNELS];
  double noise_loud[PEAQ_ORACLE_TRACE_CHANNELS];
  double adb_steps;             /* binaural Q, movs.c:1267 */
  double det_prob;              /* binaural P, movs.c:1269 */
  double signal_energy;         /* running totals, gstpeaq.c:913-918 */
  double noise_energy;
} PeaqOracleFftTrace;

/* Per filter-bank-clock frame (advanced mode). */
typedef struct
{
  int frame;
  int above_threshold;
  double unsmeared[2][PEAQ_ORACLE_TRACE_CHANNELS][40];
  double excitation[2][PEAQ_ORACLE_TRACE_CHANNELS][40];
  double mod_diff[PEAQ_ORACLE_TRACE_CHANNELS];
  double temp_wt[PEAQ_ORACLE_TRACE_CHANNELS];
  double noise_loud[PEAQ_ORACLE_TRACE_CHANNELS];
  double missing_comp[PEAQ_ORACLE_TRACE_CHANNELS];
  double lin_dist[PEAQ_ORACLE_TRACE_CHANNELS];
} PeaqOracleFbTrace;

typedef struct
{
  double odg;
  double di;
  double totalsnr;
  double movs[11];              /* order of gstpeaq.c:86-108 */
  int n_movs;
  unsigned frames_fft;
  unsigned frames_fb;
  unsigned loudness_reached_frame; /* UINT_MAX = never */
} PeaqOracleResult;

PeaqOracle *peaq_oracle_new (int advanced, double playback_level, int channels);
void peaq_oracle_free (PeaqOracle *o);
/* interleaved F32, n_* = samples per channel; either side may be empty */
void peaq_oracle_push (PeaqOracle *o, const float *ref, size_t n_ref,
                       const float *test, size_t n_test);
void peaq_oracle_finish (PeaqOracle *o);
void peaq_oracle_result (const PeaqOracle *o, PeaqOracleResult *out);
void peaq_oracle_set_fft_trace (PeaqOracle *o, PeaqOracleFftTrace *buf,
                                size_t capacity);
void peaq_oracle_set_fb_trace (PeaqOracle *o, PeaqOracleFbTrace *buf,
                               size_t capacity);

/* whole (ref,test) pair in one call; n = samples per channel of each signal */
void peaq_oracle_run_pair (int advanced, double playback_level, int channels,
                           const float *ref, size_t n_ref, const float *test,
                           size_t n_test, PeaqOracleResult *out);

/* Constant tables, for table-parity tests against the engine.
 * model 0 = FFT ear model (109/55 bands per `advanced`), 1 = filter bank.
 * which: 0 fc, 1 internal noise, 2 ear-model time constant, 3 excitation
 * threshold, 4 threshold index, 5 loudness factor, 6 masking difference,
 * 7 aUC, 8 gIL, 9 spreading normalisation, 10 band lower weight, 11 band
 * upper weight, 12 band lower bin, 13 band upper bin, 14 level-adapter /
 * modulation time constant.  Returns the number of values written. */
int peaq_oracle_table (const PeaqOracle *o, int model, int which, double *out);

size_t peaq_oracle_sizeof (int which);

/* single stages, for the reference's golden vectors (testpeaq.c) */
typedef struct _PeaqOracleStage PeaqOracleStage;
PeaqOracleStage *peaq_oracle_stage_new (int bands /* 109, 55 or 40 */ );
void peaq_oracle_stage_free (PeaqOracleStage *s);
/* FFT ear model on one 2048-sample mono frame; outputs may be NULL */
void peaq_oracle_stage_fft_ear (PeaqOracleStage *s, const float *frame,
                                double *power_spectrum /*1025*/ ,
                                double *weighted /*1025*/ ,
                                double *unsmeared, double *excitation);
/* filter-bank ear model on one 192-sample mono frame */
void peaq_oracle_stage_fb_ear (PeaqOracleStage *s, const float *frame,
                               double *unsmeared /*40*/ , double *excitation);
double peaq_oracle_stage_loudness (PeaqOracleStage *s, int filterbank);
void peaq_oracle_stage_level_adapt (PeaqOracleStage *s, const double *ref_exc,
                                    const double *test_exc, double *ref_out,
                                    double *test_out);
void peaq_oracle_stage_modulation (PeaqOracleStage *s, const double *unsmeared,
                                   double *modulation, double *avg_loudness);

#ifdef __cplusplus
}
#endif
#endif
