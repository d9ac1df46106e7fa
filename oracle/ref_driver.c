/* oracle/ref_driver.c -- TEST INFRASTRUCTURE ONLY (never shipped, never on the
 * product path).
 *
 * Drives the REFERENCE's own per-frame code so it can act as parity oracle and
 * as the "reference" CPU baseline.  The numeric classes are the reference's
 * .c files compiled unmodified from /root/reference/src against oracle/shim/.
 * The per-frame functions of gstpeaq.c (which as a whole needs GStreamer) are
 * pulled in by LINE RANGE at build time into oracle/_ref/gstpeaq_extract.inc
 * (see oracle/Makefile; nothing of it is copied into tracked files):
 *   75-139    MOV enums + struct _GstPeaq
 *   171-172   prototypes free/alloc_per_channel_data
 *   183-192   prototypes of the process/calculate functions
 *   399-473   free_per_channel_data, alloc_per_channel_data
 *   793-1099  apply_ear_model ... is_frame_above_threshold
 * Only the GStreamer-bound glue is restated below, each piece citing what it
 * follows:
 *   peaq_ref_new      <- init() gstpeaq.c:357-376, set_property(advanced)
 *                        :516-560, set_caps() :575-586
 *   fifo_* / process  <- GstAdapter use in pad_chain :626-652 and
 *                        do_processing :596-611
 *   peaq_ref_finish   <- change_state PAUSED->READY :764-778, do_flush :716-745
 */
#include <glib-object.h>
#include <glib/gprintf.h>
#include <math.h>
#include <string.h>

#include "fbearmodel.h"
#include "fftearmodel.h"
#include "leveladapter.h"
#include "modpatt.h"
#include "movaccum.h"
#include "movs.h"
#include "nn.h"

typedef struct { GObject parent; } GstElement;
typedef struct _PeaqRefDummyPad GstPad;
typedef struct _PeaqRefDummyAdapter GstAdapter;
typedef struct _GstPeaq GstPeaq;

#include "gstpeaq_extract.inc"

typedef struct
{
  float *data;
  size_t len;                   /* floats available */
  size_t cap;
} Fifo;

static void
fifo_push (Fifo *f, const float *x, size_t n)
{
  if (f->len + n > f->cap) {
    f->cap = (f->len + n) * 2 + 4096;
    f->data = (float *) realloc (f->data, f->cap * sizeof (float));
  }
  if (n)
    memcpy (f->data + f->len, x, n * sizeof (float));
  f->len += n;
}

static void
fifo_flush (Fifo *f, size_t n)
{
  memmove (f->data, f->data + n, (f->len - n) * sizeof (float));
  f->len -= n;
}

typedef struct _PeaqRef
{
  GstPeaq peaq;
  Fifo ref_fft, test_fft, ref_fb, test_fb;
} PeaqRef;

PeaqRef *
peaq_ref_new (int advanced, double playback_level, int channels)
{
  guint i;
  PeaqRef *r = (PeaqRef *) calloc (1, sizeof (PeaqRef));
  GstPeaq *peaq = &r->peaq;

  /* init(), gstpeaq.c:357-376 */
  peaq->frame_counter = 0;
  peaq->frame_counter_fb = 0;
  peaq->loudness_reached_frame = G_MAXUINT;
  peaq->total_signal_energy = 0.;
  peaq->total_noise_energy = 0.;
  peaq->channels = 0;
  peaq->fft_ear_model = g_object_new (PEAQ_TYPE_FFTEARMODEL, NULL);
  peaq->fb_ear_model = g_object_new (PEAQ_TYPE_FILTERBANKEARMODEL, NULL);
  for (i = 0; i < COUNT_MOV_BASIC; i++)
    peaq->mov_accum[i] = peaq_movaccum_new ();
  peaq->console_output = FALSE;

  /* set_property(PROP_PLAYBACK_LEVEL), gstpeaq.c:509-514 */
  g_object_set (peaq->fft_ear_model, "playback-level", playback_level, NULL);
  g_object_set (peaq->fb_ear_model, "playback-level", playback_level, NULL);

  /* set_property(PROP_MODE_ADVANCED), gstpeaq.c:516-560 */
  peaq->advanced = advanced ? TRUE : FALSE;
  g_object_set (peaq->fft_ear_model, "number-of-bands",
                (guint) (advanced ? 55 : 109), NULL);
  if (peaq->advanced) {
    peaq_movaccum_set_mode (peaq->mov_accum[MOVADV_RMS_MOD_DIFF], MODE_RMS);
    peaq_movaccum_set_mode (peaq->mov_accum[MOVADV_SEGMENTAL_NMR], MODE_AVG);
    peaq_movaccum_set_mode (peaq->mov_accum[MOVADV_EHS], MODE_AVG);
    peaq_movaccum_set_mode (peaq->mov_accum[MOVADV_AVG_LIN_DIST], MODE_AVG);
    peaq_movaccum_set_mode (peaq->mov_accum[MOVADV_RMS_NOISE_LOUD_ASYM],
                            MODE_RMS_ASYM);
  } else {
    peaq_movaccum_set_mode (peaq->mov_accum[MOVBASIC_BANDWIDTH_REF], MODE_AVG);
    peaq_movaccum_set_mode (peaq->mov_accum[MOVBASIC_BANDWIDTH_TEST], MODE_AVG);
    peaq_movaccum_set_mode (peaq->mov_accum[MOVBASIC_TOTAL_NMR], MODE_AVG_LOG);
    peaq_movaccum_set_mode (peaq->mov_accum[MOVBASIC_WIN_MOD_DIFF],
                            MODE_AVG_WINDOW);
    peaq_movaccum_set_mode (peaq->mov_accum[MOVBASIC_ADB], MODE_ADB);
    peaq_movaccum_set_mode (peaq->mov_accum[MOVBASIC_EHS], MODE_AVG);
    peaq_movaccum_set_mode (peaq->mov_accum[MOVBASIC_AVG_MOD_DIFF_1], MODE_AVG);
    peaq_movaccum_set_mode (peaq->mov_accum[MOVBASIC_AVG_MOD_DIFF_2], MODE_AVG);
    peaq_movaccum_set_mode (peaq->mov_accum[MOVBASIC_RMS_NOISE_LOUD], MODE_RMS);
    peaq_movaccum_set_mode (peaq->mov_accum[MOVBASIC_MFPD], MODE_FILTERED_MAX);
    peaq_movaccum_set_mode (peaq->mov_accum[MOVBASIC_REL_DIST_FRAMES],
                            MODE_AVG);
  }

  /* set_caps(), gstpeaq.c:575-586 */
  peaq->channels = channels;
  for (i = 0; i < COUNT_MOV_BASIC; i++)
    if (!peaq->advanced && (i == MOVBASIC_ADB || i == MOVBASIC_MFPD))
      peaq_movaccum_set_channels (peaq->mov_accum[i], 1);
    else
      peaq_movaccum_set_channels (peaq->mov_accum[i], peaq->channels);
  alloc_per_channel_data (peaq);
  return r;
}

void
peaq_ref_free (PeaqRef *r)
{
  guint i;
  if (!r)
    return;
  free_per_channel_data (&r->peaq);
  g_object_unref (r->peaq.fft_ear_model);
  g_object_unref (r->peaq.fb_ear_model);
  for (i = 0; i < COUNT_MOV_BASIC; i++)
    g_object_unref (r->peaq.mov_accum[i]);
  free (r->ref_fft.data);
  free (r->test_fft.data);
  free (r->ref_fb.data);
  free (r->test_fb.data);
  free (r);
}

/* do_processing(), gstpeaq.c:596-611, sizes in floats instead of bytes */
static void
process (GstPeaq *peaq, Fifo *ref, Fifo *test,
         void (*process_block) (GstPeaq *, gfloat *, gfloat *),
         size_t frame_floats, size_t step_floats)
{
  size_t pos = 0;
  while (ref->len - pos >= frame_floats && test->len - pos >= frame_floats) {
    process_block (peaq, ref->data + pos, test->data + pos);
    pos += step_floats;
  }
  /* both FIFOs always advance together, so one flush at the end is the same
   * as gst_adapter_flush after every frame */
  fifo_flush (ref, pos);
  fifo_flush (test, pos);
}

/* pad_chain(), gstpeaq.c:626-652; n_* = samples per channel; either side may
 * be empty.  ref/test are interleaved F32 as on the element's pads. */
void
peaq_ref_push (PeaqRef *r, const float *ref, size_t n_ref, const float *test,
               size_t n_test)
{
  GstPeaq *peaq = &r->peaq;
  size_t ch = (size_t) peaq->channels;
  size_t fft_frame = ch * peaq_earmodel_get_frame_size (peaq->fft_ear_model);
  size_t fft_step = ch * peaq_earmodel_get_step_size (peaq->fft_ear_model);
  fifo_push (&r->ref_fft, ref, n_ref * ch);
  fifo_push (&r->test_fft, test, n_test * ch);
  if (peaq->advanced) {
    size_t fb_frame = ch * peaq_earmodel_get_frame_size (peaq->fb_ear_model);
    fifo_push (&r->ref_fb, ref, n_ref * ch);
    fifo_push (&r->test_fb, test, n_test * ch);
    process (peaq, &r->ref_fft, &r->test_fft, process_fft_block_advanced,
             fft_frame, fft_step);
    process (peaq, &r->ref_fb, &r->test_fb, process_fb_block, fb_frame,
             fb_frame);
  } else {
    process (peaq, &r->ref_fft, &r->test_fft, process_fft_block_basic,
             fft_frame, fft_step);
  }
}

/* do_flush(), gstpeaq.c:716-745 */
static void
flush (GstPeaq *peaq, Fifo *ref, Fifo *test,
       void (*process_block) (GstPeaq *, gfloat *, gfloat *), guint frame_size)
{
  if (ref->len || test->len) {
    size_t frame_floats = (size_t) peaq->channels * frame_size;
    gfloat *pr = (gfloat *) calloc (frame_floats, sizeof (gfloat));
    gfloat *pt = (gfloat *) calloc (frame_floats, sizeof (gfloat));
    size_t nr = MIN (ref->len, frame_floats);
    size_t nt = MIN (test->len, frame_floats);
    memcpy (pr, ref->data, nr * sizeof (gfloat));
    memcpy (pt, test->data, nt * sizeof (gfloat));
    process_block (peaq, pr, pt);
    fifo_flush (ref, nr);
    fifo_flush (test, nt);
    free (pr);
    free (pt);
  }
}

/* change_state(PAUSED->READY), gstpeaq.c:764-778 */
void
peaq_ref_finish (PeaqRef *r)
{
  GstPeaq *peaq = &r->peaq;
  if (peaq->advanced) {
    flush (peaq, &r->ref_fft, &r->test_fft, process_fft_block_advanced,
           peaq_earmodel_get_frame_size (peaq->fft_ear_model));
    flush (peaq, &r->ref_fb, &r->test_fb, process_fb_block,
           peaq_earmodel_get_frame_size (peaq->fb_ear_model));
  } else {
    flush (peaq, &r->ref_fft, &r->test_fft, process_fft_block_basic,
           peaq_earmodel_get_frame_size (peaq->fft_ear_model));
  }
}

/* properties odg / di / totalsnr (gstpeaq.c:484-497) + the MOVs behind them */
void
peaq_ref_result (PeaqRef *r, double *odg, double *di, double *movs,
                 double *totalsnr, unsigned *frames_fft, unsigned *frames_fb,
                 unsigned *loudness_reached_frame)
{
  GstPeaq *peaq = &r->peaq;
  guint i, n = peaq->advanced ? COUNT_MOV_ADVANCED : COUNT_MOV_BASIC;
  if (movs)
    for (i = 0; i < n; i++)
      movs[i] = peaq_movaccum_get_value (peaq->mov_accum[i]);
  if (di)
    *di = peaq->advanced ? calculate_di_advanced (peaq)
      : calculate_di_basic (peaq);
  if (odg)
    *odg = calculate_odg (peaq);
  if (totalsnr)
    *totalsnr =
      10 * log10 (peaq->total_signal_energy / peaq->total_noise_energy);
  if (frames_fft)
    *frames_fft = peaq->frame_counter;
  if (frames_fb)
    *frames_fb = peaq->frame_counter_fb;
  if (loudness_reached_frame)
    *loudness_reached_frame = peaq->loudness_reached_frame;
}

/* Per-frame state taps through the reference's own accessors
 * (fftearmodel.h:43-49, earmodel.h:163-166, leveladapter.h:54-55,
 * modpatt.h:55-56).  which: 0 power spectrum (1025), 1 weighted power
 * spectrum (1025), 2 unsmeared excitation (B), 3 excitation (B) of the FFT
 * model; 4/5 unsmeared/excitation of the filter-bank model (40);
 * 6/7 level-adapted ref/test patterns; 8 modulation, 9 average loudness of
 * the modulation processor (side selects ref/test).  Returns the length. */
int
peaq_ref_tap (PeaqRef *r, int which, int side_test, int channel, double *out)
{
  GstPeaq *peaq = &r->peaq;
  const double *src = NULL;
  int n = 0;
  gpointer fft_state = side_test ? peaq->test_fft_ear_state[channel]
    : peaq->ref_fft_ear_state[channel];
  PeaqModulationProcessor *mp = side_test
    ? peaq->test_modulation_processor[channel]
    : peaq->ref_modulation_processor[channel];
  int fft_bands = (int) peaq_earmodel_get_band_count (peaq->fft_ear_model);
  int proc_bands = peaq->advanced ? 40 : fft_bands;
  switch (which) {
    case 0:
      src = peaq_fftearmodel_get_power_spectrum (fft_state);
      n = 1025;
      break;
    case 1:
      src = peaq_fftearmodel_get_weighted_power_spectrum (fft_state);
      n = 1025;
      break;
    case 2:
      src = peaq_earmodel_get_unsmeared_excitation (peaq->fft_ear_model,
                                                    fft_state);
      n = fft_bands;
      break;
    case 3:
      src = peaq_earmodel_get_excitation (peaq->fft_ear_model, fft_state);
      n = fft_bands;
      break;
    case 4:
    case 5:
      if (!peaq->advanced)
        return 0;
      {
        gpointer fb_state = side_test ? peaq->test_fb_ear_state[channel]
          : peaq->ref_fb_ear_state[channel];
        src = which == 4
          ? peaq_earmodel_get_unsmeared_excitation (peaq->fb_ear_model,
                                                    fb_state)
          : peaq_earmodel_get_excitation (peaq->fb_ear_model, fb_state);
        n = 40;
      }
      break;
    case 6:
      src = peaq_leveladapter_get_adapted_ref (peaq->level_adapter[channel]);
      n = proc_bands;
      break;
    case 7:
      src = peaq_leveladapter_get_adapted_test (peaq->level_adapter[channel]);
      n = proc_bands;
      break;
    case 8:
      src = peaq_modulationprocessor_get_modulation (mp);
      n = proc_bands;
      break;
    case 9:
      src = peaq_modulationprocessor_get_average_loudness (mp);
      n = proc_bands;
      break;
    default:
      return 0;
  }
  memcpy (out, src, (size_t) n * sizeof (double));
  return n;
}

/* Constant tables of the reference's ear models, for table parity tests.
 * which: 0 fc, 1 internal noise, 2 ear time constant, 3 excitation
 * threshold, 4 threshold index, 5 loudness factor (earmodel.h fields);
 * 6 masking difference (FFT model only).  model: 0 FFT, 1 filter bank. */
int
peaq_ref_table (PeaqRef *r, int model, int which, double *out)
{
  PeaqEarModel *m = model ? r->peaq.fb_ear_model : r->peaq.fft_ear_model;
  int n = (int) m->band_count, i;
  const double *src = NULL;
  switch (which) {
    case 0: src = m->fc; break;
    case 1: src = m->internal_noise; break;
    case 2: src = m->ear_time_constants; break;
    case 3: src = m->excitation_threshold; break;
    case 4: src = m->threshold; break;
    case 5: src = m->loudness_factor; break;
    case 6:
      if (model)
        return 0;
      src = peaq_fftearmodel_get_masking_difference (PEAQ_FFTEARMODEL (m));
      break;
    default:
      return 0;
  }
  for (i = 0; i < n; i++)
    out[i] = src[i];
  return n;
}
