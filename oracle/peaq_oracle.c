/* oracle/peaq_oracle.c -- TEST INFRASTRUCTURE ONLY; see peaq_oracle.h.
 *
 * Plain-C restatement of the reference's per-frame path.  Each function cites
 * the reference file:line (relative to /root/reference/src) it follows.  The
 * shipped compile-time switches of settings.h are fixed:
 *   SWAP_MOD_PATTS_FOR_NOISE_LOUDNESS_MOVS=1, CENTER_EHS_CORRELATION_WINDOW=0,
 *   EHS_SUBTRACT_DC_BEFORE_WINDOW=1, USE_FLOOR_FOR_STEPS_ABOVE_THRESHOLD=0,
 *   CLAMP_MOVS=0, SWAP_SLOPE_FILTER_COEFFICIENTS=0   (settings.h:47-97).
 * Expression order follows the reference so that, built with the same
 * compiler flags and the same FFT (oracle/fft_f64.c), results agree with
 * oracle/_ref/libpeaq_ref.so to the last bit or two.
 */
#include "peaq_oracle.h"
#include "fft_f64.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXB PEAQ_ORACLE_MAX_BANDS
#define MAXCH PEAQ_ORACLE_MAX_CHANNELS
#define FS 48000.
#define NFFT 2048
#define NBINS (NFFT / 2 + 1)
#define FB_FRAME 192
#define FB_BANDS 40
#define FB_BUF 1456
#define MAXLAG 256

#define DMAX(a, b) (((a) > (b)) ? (a) : (b))
#define DMIN(a, b) (((a) < (b)) ? (a) : (b))

/* ------------------------------------------------------------------------ */
/* per-band constants shared by both ear models (earmodel.c:278-323)         */

typedef struct
{
  int B;
  int step;                     /* samples per frame step */
  double loudness_scale;
  double tau_min, tau_100;
  double fc[MAXB];
  double internal_noise[MAXB];
  double ethres[MAXB];          /* excitation threshold */
  double thres[MAXB];           /* threshold index */
  double loudfac[MAXB];
  double a_ear[MAXB];           /* ear model's own smearing constant */
  double a_proc[MAXB];          /* level adapter / modulation constant */
} Bands;

/* earmodel.c:626-635 */
static double
time_constant (const Bands *b, int band, double tau_min, double tau_100)
{
  double tau = tau_min + 100. / b->fc[band] * (tau_100 - tau_min);
  return exp (b->step / (-48000. * tau));
}

/* earmodel.c:278-323 + update_ear_time_constants; leveladapter.c:203-206,
 * modpatt.c:182-186 for a_proc */
static void
bands_init (Bands *b, const double *fc, int B)
{
  int i;
  b->B = B;
  for (i = 0; i < B; i++) {
    double f = fc[i];
    b->fc[i] = f;
    b->internal_noise[i] = pow (10., 0.4 * 0.364 * pow (f / 1000., -0.8));
    b->ethres[i] = pow (10., 0.364 * pow (f / 1000., -0.8));
    b->thres[i] =
      pow (10.,
           0.1 * (-2. - 2.05 * atan (f / 4000.) -
                  0.75 * atan (f / 1600. * f / 1600.)));
    b->loudfac[i] =
      b->loudness_scale * pow (b->ethres[i] / (1e4 * b->thres[i]), 0.23);
  }
  for (i = 0; i < B; i++) {
    b->a_ear[i] = time_constant (b, i, b->tau_min, b->tau_100);
    b->a_proc[i] = time_constant (b, i, 0.008, 0.05);
  }
}

/* earmodel.c:701-709 */
static double
ear_weight (double frequency)
{
  double f_kHz = frequency / 1000.;
  double W_dB =
    -0.6 * 3.64 * pow (f_kHz, -0.8) + 6.5 * exp (-0.6 * pow (f_kHz - 3.3, 2)) -
    1e-3 * pow (f_kHz, 3.6);
  return pow (10, W_dB / 20);
}

/* earmodel.c:890-907 */
static double
calc_loudness (const Bands *b, const double *excitation)
{
  int i;
  double overall = 0.;
  for (i = 0; i < b->B; i++) {
    double l = b->loudfac[i]
      * (pow (1. - b->thres[i] + b->thres[i] * excitation[i] / b->ethres[i],
              0.23) - 1.);
    overall += DMAX (l, 0.);
  }
  overall *= 24. / b->B;
  return overall;
}

/* ------------------------------------------------------------------------ */
/* FFT ear model (fftearmodel.c)                                             */

typedef struct
{
  Bands bands;
  PeaqOracleFFT *fft;
  double hann[NFFT];
  double earw2[NBINS];          /* squared outer/middle ear weight */
  double level_factor;
  double dz;
  int lo[MAXB], hi[MAXB];
  double wl[MAXB], wu[MAXB];
  double aL;                    /* lower_spreading */
  double aLe;                   /* lower_spreading ^ 0.4 */
  double aUC[MAXB], gIL[MAXB], norm[MAXB], maskdiff[MAXB];
} FftModel;

typedef struct
{
  double filtered[MAXB];
  double unsmeared[MAXB];
  double excitation[MAXB];
  double power[NBINS];
  double weighted[NBINS];
  int energy_flag;
} FftState;

/* fftearmodel.c:636-676 */
static void
fft_spread (const FftModel *m, const double *Pp, double *E2)
{
  int i, j, B = m->bands.B;
  double aUCEe[MAXB], Ene[MAXB];
  const double aLe = m->aLe;
  for (i = 0; i < B; i++) {
    double aUCE = m->aUC[i] * pow (Pp[i], 0.2 * m->dz);
    double gIU = (1. - pow (aUCE, B - i)) / (1. - aUCE);
    double En = Pp[i] / (m->gIL[i] + gIU - 1.);
    aUCEe[i] = pow (aUCE, 0.4);
    Ene[i] = pow (En, 0.4);
  }
  E2[B - 1] = Ene[B - 1];
  for (i = B - 1; i > 0; i--)
    E2[i - 1] = aLe * E2[i] + Ene[i - 1];
  for (i = 0; i < B - 1; i++) {
    double r = Ene[i];
    for (j = i + 1; j < B; j++) {
      r *= aUCEe[i];
      E2[j] += r;
    }
  }
  for (i = 0; i < B; i++)
    E2[i] = pow (E2[i], 1. / 0.4) / m->norm[i];
}

/* fftearmodel.c:603-620 */
static void
fft_group (const FftModel *m, const double *spectrum, double *band_power)
{
  int i, k;
  for (i = 0; i < m->bands.B; i++) {
    band_power[i] = m->wl[i] * spectrum[m->lo[i]] + m->wu[i] * spectrum[m->hi[i]];
    for (k = m->lo[i] + 1; k < m->hi[i]; k++)
      band_power[i] += spectrum[k];
    if (band_power[i] < 1e-12)
      band_power[i] = 1e-12;
  }
}

/* base_init :159-173, init :239-257, set_playback_level :304-314,
 * set_property(number-of-bands) :692-788 */
static void
fft_model_init (FftModel *m, int B, double playback_level)
{
  const double GAMMA = 0.84971762641205;
  int k, band;
  double fc[MAXB], spread[MAXB];
  memset (m, 0, sizeof *m);
  m->fft = peaq_oracle_fft_new (NFFT);
  for (k = 0; k < NFFT; k++)
    m->hann[k] = sqrt (8. / 3.) * 0.5 * (1. - cos (2 * M_PI * k / (NFFT - 1)));
  for (k = 0; k <= NFFT / 2; k++)
    m->earw2[k] = pow (ear_weight ((double) k * FS / NFFT), 2);
  m->level_factor = pow (10, playback_level / 10) /
    (8. / 3. * (GAMMA / 4 * (NFFT - 1)) * (GAMMA / 4 * (NFFT - 1)));

  m->bands.step = NFFT / 2;
  m->bands.loudness_scale = 1.07664;
  m->bands.tau_min = 0.008;
  m->bands.tau_100 = 0.030;

  m->dz = 27. / (B - 1);
  {
    double zL = 7. * asinh (80. / 650.);
    double zU = 7. * asinh (18000. / 650.);
    m->aL = pow (10., -2.7 * m->dz);
    m->aLe = pow (m->aL, 0.4);
    for (band = 0; band < B; band++) {
      double zl = zL + band * m->dz;
      double zu = DMIN (zU, zL + (band + 1) * m->dz);
      double zc = (zu + zl) / 2.;
      double curr_fc = 650. * sinh (zc / 7.);
      double fl = 650. * sinh (zl / 7.);
      double fu = 650. * sinh (zu / 7.);
      double upper_freq, U;
      fc[band] = curr_fc;
      m->lo[band] = (int) (unsigned) round (fl / FS * NFFT);
      m->hi[band] = (int) (unsigned) round (fu / FS * NFFT);
      upper_freq = (2 * m->lo[band] + 1) / 2. * FS / NFFT;
      if (upper_freq > fu)
        upper_freq = fu;
      U = upper_freq - fl;
      m->wl[band] = U * NFFT / FS;
      if (m->lo[band] == m->hi[band]) {
        m->wu[band] = 0;
      } else {
        double lower_freq = (2 * m->hi[band] - 1) / 2. * FS / NFFT;
        U = fu - lower_freq;
        m->wu[band] = U * NFFT / FS;
      }
      m->aUC[band] = pow (10., (-2.4 - 23. / curr_fc) * m->dz);
      m->gIL[band] = (1. - pow (m->aL, band + 1)) / (1. - m->aL);
      m->norm[band] = 1.;
      m->maskdiff[band] =
        pow (10., (band * m->dz <= 12. ? 3. : 0.25 * band * m->dz) / 10.);
    }
  }
  bands_init (&m->bands, fc, B);
  /* normalisation = the spreading routine itself applied to all-ones */
  fft_spread (m, m->norm, spread);
  for (band = 0; band < B; band++)
    m->norm[band] = spread[band];
}

static void
fft_model_clear (FftModel *m)
{
  peaq_oracle_fft_free (m->fft);
  m->fft = NULL;
}

/* fftearmodel.c:432-515 */
static void
fft_process (const FftModel *m, FftState *s, const float *x)
{
  int k, i, B = m->bands.B;
  double windowed[NFFT];
  PeaqOracleComplex out[NBINS];
  double band_power[MAXB], noisy[MAXB];
  double energy;

  for (k = 0; k < NFFT; k++)
    windowed[k] = m->hann[k] * x[k];
  peaq_oracle_fft_forward (m->fft, windowed, out);
  for (k = 0; k < NBINS; k++) {
    s->power[k] = (out[k].r * out[k].r + out[k].i * out[k].i) * m->level_factor;
    s->weighted[k] = s->power[k] * m->earw2[k];
  }
  fft_group (m, s->weighted, band_power);
  for (i = 0; i < B; i++)
    noisy[i] = band_power[i] + m->bands.internal_noise[i];
  fft_spread (m, noisy, s->unsmeared);
  for (i = 0; i < B; i++) {
    double a = m->bands.a_ear[i];
    s->filtered[i] = a * s->filtered[i] + (1. - a) * s->unsmeared[i];
    s->excitation[i] =
      s->filtered[i] > s->unsmeared[i] ? s->filtered[i] : s->unsmeared[i];
  }
  energy = 0.;
  for (k = NFFT / 2; k < NFFT; k++)
    energy += x[k] * x[k];      /* float product, double accumulation */
  s->energy_flag = energy >= 8000. / (32768. * 32768.);
}

/* ------------------------------------------------------------------------ */
/* filter-bank ear model (fbearmodel.c)                                      */

static const int fb_len[FB_BANDS] = {
  1456, 1438, 1406, 1362, 1308, 1244, 1176, 1104, 1030, 956, 884, 814, 748,
  686, 626, 570, 520, 472, 430, 390, 354, 320, 290, 262, 238, 214, 194, 176,
  158, 144, 130, 118, 106, 96, 86, 78, 70, 64, 58, 52
};                              /* Table 8 of BS.1387, fbearmodel.c:57-61 */

#define FB_SLOPE_A 0.993355506255034
#define FB_DIST 0.921851456499719
#define FB_CL 0.0802581846102741

typedef struct
{
  Bands bands;
  double level_factor;
  double *h_re[FB_BANDS];
  double *h_im[FB_BANDS];
  double back_mask[6];
} FbModel;

typedef struct
{
  double x1, x2, y1a, y2a, y1b, y2b;    /* two cascaded DC-reject biquads */
  double buf[2 * FB_BUF];
  unsigned off;
  double cu[FB_BANDS];
  double E0[FB_BANDS][11];
  double excitation[FB_BANDS];
  double unsmeared[FB_BANDS];
} FbState;

/* class_init :150-186, init :188-225, set_playback_level :248-254 */
static void
fb_model_init (FbModel *m, double playback_level)
{
  int band, n, i;
  double fc[FB_BANDS];
  memset (m, 0, sizeof *m);
  m->bands.step = FB_FRAME;
  m->bands.loudness_scale = 1.26539;
  m->bands.tau_min = 0.004;
  m->bands.tau_100 = 0.020;
  for (i = 0; i < 6; i++)
    m->back_mask[i] =
      cos (M_PI * (i - 5.) / 12.) * cos (M_PI * (i - 5.) / 12.) * 0.9761 / 6.;
  for (band = 0; band < FB_BANDS; band++) {
    int N = fb_len[band];
    double f =
      sinh ((asinh (50. / 650.) +
             band * (asinh (18000. / 650.) - asinh (50. / 650.)) / 39.)) * 650.;
    double Wt = ear_weight (f);
    fc[band] = f;
    m->h_re[band] = (double *) malloc (sizeof (double) * (N / 2 + 1));
    m->h_im[band] = (double *) malloc (sizeof (double) * (N / 2 + 1));
    for (n = 0; n < N / 2 + 1; n++) {
      double win = 4. / N * sin (M_PI * n / N) * sin (M_PI * n / N) * Wt;
      m->h_re[band][n] = win * cos (2 * M_PI * f * (n - N / 2.) / 48000.);
      m->h_im[band][n] = win * sin (2 * M_PI * f * (n - N / 2.) / 48000.);
    }
  }
  bands_init (&m->bands, fc, FB_BANDS);
  m->level_factor = pow (10., playback_level / 20.);
}

static void
fb_model_clear (FbModel *m)
{
  int band;
  for (band = 0; band < FB_BANDS; band++) {
    free (m->h_re[band]);
    free (m->h_im[band]);
    m->h_re[band] = m->h_im[band] = NULL;
  }
}

/* fbearmodel.c:398-435 */
static void
fb_apply_bank (const FbModel *m, const FbState *s, double *out_re,
               double *out_im)
{
  int band, n;
  for (band = 0; band < FB_BANDS; band++) {
    int N = fb_len[band];
    int D = 1 + (fb_len[0] - N) / 2;
    int N_2 = N / 2;
    double re = 0, im = 0;
    const double *in1 = s->buf + D + s->off;
    const double *in2 = s->buf + D + N + s->off;
    const double *h_re = m->h_re[band];
    const double *h_im = m->h_im[band];
    for (n = 1; n < N_2; n++) {
      in1++;
      h_re++;
      h_im++;
      in2--;
      re += (*in1 + *in2) * *h_re;
      im += (*in1 - *in2) * *h_im;
    }
    in1++;
    h_re++;
    h_im++;
    re += *in1 * *h_re;
    im += *in1 * *h_im;
    out_re[band] = re;
    out_im[band] = im;
  }
}

/* fbearmodel.c:275-396 */
static void
fb_process (const FbModel *m, FbState *s, const float *x)
{
  int k, band, j, i;
  double level_factor = m->level_factor;
  for (k = 0; k < FB_FRAME; k++) {
    double scaled = x[k] * level_factor;
    double hp1 = scaled - 2. * s->x1 + s->x2 + 1.99517 * s->y1a - 0.995174 * s->y2a;
    double hp2 = hp1 - 2. * s->y1a + s->y2a + 1.99799 * s->y1b - 0.997998 * s->y2b;
    s->x2 = s->x1;
    s->x1 = scaled;
    s->y2a = s->y1a;
    s->y1a = hp1;
    s->y2b = s->y1b;
    s->y1b = hp2;
    if (s->off == 0)
      s->off = FB_BUF;
    s->off--;
    s->buf[s->off] = hp2;
    s->buf[s->off + FB_BUF] = hp2;
    if (k % 32 == 0) {
      double o_re[FB_BANDS], o_im[FB_BANDS], A_re[FB_BANDS], A_im[FB_BANDS];
      fb_apply_bank (m, s, o_re, o_im);
      for (band = 0; band < FB_BANDS; band++) {
        A_re[band] = o_re[band];
        A_im[band] = o_im[band];
      }
      for (band = 0; band < FB_BANDS; band++) {
        double fc = m->bands.fc[band];
        double L = 10 * log10 (o_re[band] * o_re[band] + o_im[band] * o_im[band]);
        double sl = DMAX (4, 24 + 230 / fc - 0.2 * L);
        double dist_s = pow (FB_DIST, sl);
        double d1, d2;
        s->cu[band] = s->cu[band] + FB_SLOPE_A * (dist_s - s->cu[band]);
        d1 = o_re[band];
        d2 = o_im[band];
        for (j = band + 1; j < FB_BANDS; j++) {
          d1 *= s->cu[band];
          d2 *= s->cu[band];
          A_re[j] += d1;
          A_im[j] += d2;
        }
      }
      for (band = FB_BANDS - 1; band > 0; band--) {
        A_re[band - 1] += FB_CL * A_re[band];
        A_im[band - 1] += FB_CL * A_im[band];
      }
      for (band = 0; band < FB_BANDS; band++) {
        memmove (s->E0[band] + 1, s->E0[band], 10 * sizeof (double));
        s->E0[band][0] = A_re[band] * A_re[band] + A_im[band] * A_im[band];
      }
    }
  }
  for (band = 0; band < FB_BANDS; band++) {
    double E1 = 0., a;
    for (i = 0; i < 5; i++)
      E1 += (s->E0[band][i] + s->E0[band][10 - i]) * m->back_mask[i];
    E1 += s->E0[band][5] * m->back_mask[5];
    s->unsmeared[band] = E1 + m->bands.internal_noise[band];
    a = m->bands.a_ear[band];
    s->excitation[band] = a * s->excitation[band] + (1. - a) * s->unsmeared[band];
  }
}

/* ------------------------------------------------------------------------ */
/* level adapter (leveladapter.c:243-340)                                    */

typedef struct
{
  double ref_filt[MAXB], test_filt[MAXB];
  double num[MAXB], den[MAXB];
  double pc_ref[MAXB], pc_test[MAXB];
  double ad_ref[MAXB], ad_test[MAXB];   /* spectrally adapted patterns */
} LevelState;

static void
level_process (const Bands *b, LevelState *s, const double *ref_exc,
               const double *test_exc)
{
  int B = b->B, k, l;
  double num = 0., den = 0., lev_corr;
  double lc[MAXB], pa_ref[MAXB], pa_test[MAXB];
  const double *lc_ref, *lc_test;
  for (k = 0; k < B; k++) {
    double a = b->a_proc[k];
    s->ref_filt[k] = a * s->ref_filt[k] + (1 - a) * ref_exc[k];
    s->test_filt[k] = a * s->test_filt[k] + (1 - a) * test_exc[k];
    num += sqrt (s->ref_filt[k] * s->test_filt[k]);
    den += s->test_filt[k];
  }
  lev_corr = num * num / (den * den);
  if (lev_corr > 1) {
    lc_test = test_exc;
    for (k = 0; k < B; k++)
      lc[k] = ref_exc[k] / lev_corr;
    lc_ref = lc;
  } else {
    lc_ref = ref_exc;
    for (k = 0; k < B; k++)
      lc[k] = test_exc[k] * lev_corr;
    lc_test = lc;
  }
  for (k = 0; k < B; k++) {
    double a = b->a_proc[k];
    s->num[k] = a * s->num[k] + lc_test[k] * lc_ref[k];
    s->den[k] = a * s->den[k] + lc_ref[k] * lc_ref[k];
    if (s->num[k] >= s->den[k]) {
      pa_ref[k] = 1.;
      pa_test[k] = s->den[k] / s->num[k];
    } else {
      pa_ref[k] = s->num[k] / s->den[k];
      pa_test[k] = 1.;
    }
  }
  for (k = 0; k < B; k++) {
    double a = b->a_proc[k];
    int m1 = k < B / 36 ? k : B / 36;
    int m2 = (B - k - 1) < B / 25 ? (B - k - 1) : B / 25;
    double ra_ref = 0., ra_test = 0.;
    for (l = k - m1; l <= k + m2; l++) {
      ra_ref += pa_ref[l];
      ra_test += pa_test[l];
    }
    ra_ref /= (m1 + m2 + 1);
    ra_test /= (m1 + m2 + 1);
    s->pc_ref[k] = a * s->pc_ref[k] + (1 - a) * ra_ref;
    s->pc_test[k] = a * s->pc_test[k] + (1 - a) * ra_test;
    s->ad_ref[k] = lc_ref[k] * s->pc_ref[k];
    s->ad_test[k] = lc_test[k] * s->pc_test[k];
  }
}

/* ------------------------------------------------------------------------ */
/* modulation processor (modpatt.c:223-251)                                  */

typedef struct
{
  double prev[MAXB], filt_loud[MAXB], filt_deriv[MAXB], mod[MAXB];
} ModState;

static void
mod_process (const Bands *b, ModState *s, const double *unsmeared)
{
  int k;
  double derivative_factor = (double) 48000 / b->step;
  for (k = 0; k < b->B; k++) {
    double a = b->a_proc[k];
    double loud = pow (unsmeared[k], 0.3);
    double d = loud - s->prev[k];
    double deriv = derivative_factor * (d < 0 ? -d : d);
    s->filt_deriv[k] = a * s->filt_deriv[k] + (1 - a) * deriv;
    s->filt_loud[k] = a * s->filt_loud[k] + (1. - a) * loud;
    s->mod[k] = s->filt_deriv[k] / (1. + s->filt_loud[k] / 0.3);
    s->prev[k] = loud;
  }
}

/* ------------------------------------------------------------------------ */
/* MOV accumulators (movaccum.c)                                             */

enum
{ ACC_AVG, ACC_AVG_LOG, ACC_RMS, ACC_RMS_ASYM, ACC_AVG_WINDOW,
  ACC_FILTERED_MAX, ACC_ADB
};                              /* movaccum.h:323-332 */
enum
{ ST_INIT, ST_NORMAL, ST_TENTATIVE };   /* movaccum.c:53-58 */

typedef struct
{
  double num, num2, den;        /* Fraction / TwinFraction */
  double past[3];               /* AVG_WINDOW history, NaN primed (:293) */
  double max, filt;             /* FILTERED_MAX */
} AccData;

typedef struct
{
  int mode, status, channels;
  AccData d[MAXCH], saved[MAXCH];
} Accum;

static void
acc_init (Accum *a, int mode, int channels)
{
  int c, i;
  memset (a, 0, sizeof *a);
  a->mode = mode;
  a->status = ST_INIT;
  a->channels = channels;
  for (c = 0; c < channels; c++)
    for (i = 0; i < 3; i++)
      a->d[c].past[i] = NAN;
}

/* movaccum.c:317-354 */
static void
acc_set_tentative (Accum *a, int tentative)
{
  if (tentative) {
    if (a->status == ST_NORMAL) {
      int c;
      for (c = 0; c < a->channels; c++) {
        /* only num/den (and max) are snapshotted; window history and the
         * filter state keep evolving */
        a->saved[c].num = a->d[c].num;
        a->saved[c].num2 = a->d[c].num2;
        a->saved[c].den = a->d[c].den;
        a->saved[c].max = a->d[c].max;
      }
      a->status = ST_TENTATIVE;
    }
  } else {
    a->status = ST_NORMAL;
  }
}

/* movaccum.c:369-425 */
static void
acc_add (Accum *a, int c, double val, double weight)
{
  AccData *d = &a->d[c];
  if (a->status == ST_INIT)
    return;
  switch (a->mode) {
    case ACC_RMS:
      weight *= weight;
      d->num += weight * val * val;
      d->den += weight;
      break;
    case ACC_RMS_ASYM:
      d->num += val * val;
      d->num2 += weight * weight;
      d->den += 1.;
      break;
    case ACC_AVG:
    case ACC_AVG_LOG:
    case ACC_ADB:
      d->num += weight * val;
      d->den += weight;
      break;
    case ACC_AVG_WINDOW:
      {
        int i;
        double val_sqrt = sqrt (val);
        if (!isnan (d->past[0])) {
          double winsum = val_sqrt;
          for (i = 0; i < 3; i++)
            winsum += d->past[i];
          winsum /= 4.;
          winsum *= winsum;
          winsum *= winsum;
          d->num += winsum;
          d->den += 1.;
        }
        for (i = 0; i < 2; i++)
          d->past[i] = d->past[i + 1];
        d->past[2] = val_sqrt;
      }
      break;
    case ACC_FILTERED_MAX:
      d->filt = 0.9 * d->filt + 0.1 * val;
      if (d->filt > d->max)
        d->max = d->filt;
      break;
  }
}

/* movaccum.c:438-481 */
static double
acc_value (const Accum *a)
{
  const AccData *data = a->status == ST_TENTATIVE ? a->saved : a->d;
  double value = 0.;
  int c;
  for (c = 0; c < a->channels; c++) {
    const AccData *d = &data[c];
    switch (a->mode) {
      case ACC_AVG:
        value += d->num / d->den;
        break;
      case ACC_AVG_LOG:
        value += 10. * log10 (d->num / d->den);
        break;
      case ACC_AVG_WINDOW:
      case ACC_RMS:
        value += sqrt (d->num / d->den);
        break;
      case ACC_RMS_ASYM:
        value += sqrt (d->num / d->den);
        value += 0.5 * sqrt (d->num2 / d->den);
        break;
      case ACC_FILTERED_MAX:
        value += d->max;
        break;
      case ACC_ADB:
        if (d->den > 0)
          value += d->num == 0. ? -0.5 : log10 (d->num / d->den);
        break;
    }
  }
  value /= a->channels;
  return value;
}

/* ------------------------------------------------------------------------ */
/* neural network (nn.c)                                                     */

static const double amin_basic[11] = {
  393.916656, 361.965332, -24.045116, 1.110661, -0.206623, 0.074318, 1.113683,
  0.950345, 0.029985, 0.000101, 0.
};
static const double amax_basic[11] = {
  921, 881.131226, 16.212030, 107.137772, 2.886017, 13.933351, 63.257874,
  1145.018555, 14.819740, 1., 1.
};
static const double wx_basic[11][3] = {
  {-0.502657, 0.436333, 1.219602}, {4.307481, 3.246017, 1.123743},
  {4.984241, -2.211189, -0.192096}, {0.051056, -1.762424, 4.331315},
  {2.321580, 1.789971, -0.754560}, {-5.303901, -3.452257, -10.814982},
  {2.730991, -6.111805, 1.519223}, {0.624950, -1.331523, -5.955151},
  {3.102889, 0.871260, -5.922878}, {-1.051468, -0.939882, -0.142913},
  {-1.804679, -0.503610, -0.620456}
};
static const double wxb_basic[3] = { -2.518254, 0.654841, -2.207228 };
static const double wy_basic[3] = { -3.817048, 4.107138, 4.629582 };
static const double wyb_basic = -0.307594;

static const double amin_adv[5] = {
  13.298751, 0.041073, -25.018791, 0.061560, 0.02452
};
static const double amax_adv[5] = {
  2166.5, 13.24326, 13.46708, 10.226771, 14.224874
};
static const double wx_adv[5][5] = {
  {21.211773, -39.013052, -1.382553, -14.545348, -0.320899},
  {-8.981803, 19.956049, 0.935389, -1.686586, -3.238586},
  {1.633830, -2.877505, -7.442935, 5.606502, -1.783120},
  {6.103821, 19.587435, -0.240284, 1.088213, -0.511314},
  {11.556344, 3.892028, 9.720441, -3.287205, -11.031250},
};
static const double wxb_adv[5] = {
  1.330890, 2.686103, 2.096598, -1.327851, 3.087055
};
static const double wy_adv[5] = {
  -4.696996, -3.289959, 7.004782, 6.651897, 4.009144
};
static const double wyb_adv = -1.360308;

/* nn.c:187-216 */
static double
di_basic (const double *movs)
{
  int i, j;
  double x[3], di;
  for (i = 0; i < 3; i++)
    x[i] = wxb_basic[i];
  for (i = 0; i <= 10; i++) {
    double m = (movs[i] - amin_basic[i]) / (amax_basic[i] - amin_basic[i]);
    for (j = 0; j < 3; j++)
      x[j] += wx_basic[i][j] * m;
  }
  di = wyb_basic;
  for (i = 0; i < 3; i++)
    di += wy_basic[i] / (1 + exp (-x[i]));
  return di;
}

/* nn.c:304-335 */
static double
di_advanced (const double *movs)
{
  int i, j;
  double x[5], di;
  for (i = 0; i < 5; i++)
    x[i] = wxb_adv[i];
  for (i = 0; i <= 4; i++) {
    double m = (movs[i] - amin_adv[i]) / (amax_adv[i] - amin_adv[i]);
    for (j = 0; j < 5; j++)
      x[j] += wx_adv[i][j] * m;
  }
  di = wyb_adv;
  for (i = 0; i < 5; i++)
    di += wy_adv[i] / (1 + exp (-x[i]));
  return di;
}

/* nn.c:372-375 */
static double
odg_from_di (double di)
{
  return -3.98 + (0.22 - -3.98) / (1 + exp (-di));
}

/* ------------------------------------------------------------------------ */
/* the per-pair object (struct _GstPeaq, gstpeaq.c:110-139)                  */

/* MOV indices, gstpeaq.c:86-108 */
enum
{ ADV_RMS_MOD_DIFF, ADV_RMS_NOISE_LOUD_ASYM, ADV_SEGMENTAL_NMR, ADV_EHS,
  ADV_AVG_LIN_DIST, N_ADV
};
enum
{ B_BW_REF, B_BW_TEST, B_TOTAL_NMR, B_WIN_MOD_DIFF, B_ADB, B_EHS,
  B_AVG_MOD_DIFF_1, B_AVG_MOD_DIFF_2, B_RMS_NOISE_LOUD, B_MFPD,
  B_REL_DIST_FRAMES, N_BASIC
};

typedef struct
{
  float *data;
  size_t len, cap;
} Fifo;

struct _PeaqOracle
{
  int advanced, channels;
  unsigned frame_counter, frame_counter_fb, loudness_reached_frame;
  FftModel fft;
  FbModel fb;
  FftState *ref_fft, *test_fft; /* [channels] */
  FbState *ref_fb, *test_fb;
  LevelState *level;
  ModState *ref_mod, *test_mod;
  Accum acc[N_BASIC];
  double total_signal_energy, total_noise_energy;
  double ehs_window[MAXLAG];
  PeaqOracleFFT *fft512, *fft256;
  Fifo q_ref_fft, q_test_fft, q_ref_fb, q_test_fb;
  PeaqOracleFftTrace *fft_trace;
  size_t fft_trace_cap;
  PeaqOracleFbTrace *fb_trace;
  size_t fb_trace_cap;
  PeaqOracleFftTrace scratch_fft;
  PeaqOracleFbTrace scratch_fb;
};

static const Bands *
proc_bands (const PeaqOracle *o)
{
  /* level adapter / modulation processors follow the filter bank in advanced
   * mode (alloc_per_channel_data, gstpeaq.c:455-470), else the FFT model */
  return o->advanced ? &o->fb.bands : &o->fft.bands;
}

/* gstpeaq.c:1081-1099; running sum is a FLOAT, the increments are double */
static int
frame_above_threshold (const float *frame, unsigned framesize, unsigned channels)
{
  float sum;
  unsigned i, c;
  for (c = 0; c < channels; c++) {
    sum = 0;
    for (i = 0; i < 5; i++)
      sum += fabs (frame[channels * i + c]);
    while (i < framesize) {
      sum += fabs (frame[channels * i + c]) - fabs (frame[channels * (i - 5) + c]);
      if (sum >= 200. / 32768)
        return 1;
      i++;
    }
  }
  return 0;
}

/* movs.c:205-254 */
static void
mov_modulation_difference (PeaqOracle *o, Accum *acc1, Accum *acc2,
                           Accum *acc_win, double *t_md1, double *t_md2,
                           double *t_wt)
{
  const Bands *b = proc_bands (o);
  int c, i, B = b->B;
  double levWt = acc2 ? 100. : 1.;
  for (c = 0; c < acc1->channels; c++) {
    const double *mr = o->ref_mod[c].mod, *mt = o->test_mod[c].mod;
    const double *lr = o->ref_mod[c].filt_loud;
    double md1 = 0., md2 = 0., wt = 0.;
    for (i = 0; i < B; i++) {
      double w;
      double diff = mr[i] - mt[i];
      if (diff < 0)
        diff = -diff;
      md1 += diff / (1. + mr[i]);
      w = mt[i] >= mr[i] ? 1. : .1;
      md2 += w * diff / (0.01 + mr[i]);
      wt += lr[i] / (lr[i] + levWt * pow (b->internal_noise[i], 0.3));
    }
    if (acc1->mode == ACC_RMS)
      md1 *= 100. / sqrt (B);
    else
      md1 *= 100. / B;
    md2 *= 100. / B;
    acc_add (acc1, c, md1, wt);
    if (acc2)
      acc_add (acc2, c, md2, wt);
    if (acc_win)
      acc_add (acc_win, c, md1, 1.);
    if (c < PEAQ_ORACLE_TRACE_CHANNELS) {
      t_md1[c] = md1;
      if (t_md2)
        t_md2[c] = md2;
      t_wt[c] = wt;
    }
  }
}

/* movs.c:709-743 */
static double
noise_loudness (const Bands *b, double alpha, double thres_fac, double S0,
                double NLmin, const double *ref_mod, const double *test_mod,
                const double *ref_exc, const double *test_exc)
{
  int i;
  double nl = 0.;
  for (i = 0; i < b->B; i++) {
    double sref = thres_fac * ref_mod[i] + S0;
    double stest = thres_fac * test_mod[i] + S0;
    double ethres = b->internal_noise[i];
    double ep_ref = ref_exc[i];
    double ep_test = test_exc[i];
    double beta = exp (-alpha * (ep_test - ep_ref) / ep_ref);
    nl += pow (ethres / stest, 0.23) *
      (pow (1. + DMAX (stest * ep_test - sref * ep_ref, 0.) /
            (ethres + sref * ep_ref * beta), 0.23) - 1.);
  }
  nl *= 24. / b->B;
  if (nl < NLmin)
    nl = 0.;
  return nl;
}

/* movs.c:776-809 */
static void
mov_bandwidth (PeaqOracle *o, PeaqOracleFftTrace *t)
{
  int c, i;
  for (c = 0; c < o->acc[B_BW_REF].channels; c++) {
    const double *rp = o->ref_fft[c].power, *tp = o->test_fft[c].power;
    double zero_threshold = tp[921];
    unsigned bw_ref = 0, bw_test = 0;
    for (i = 922; i < 1024; i++)
      if (tp[i] >= zero_threshold)
        zero_threshold = tp[i];
    for (i = 921; i > 0; i--)
      if (rp[i - 1] > 10. * zero_threshold) {
        bw_ref = i;
        break;
      }
    if (bw_ref > 346) {
      for (i = bw_ref; i > 0; i--)
        if (tp[i - 1] >= 3.16227766016838 * zero_threshold) {
          bw_test = i;
          break;
        }
      acc_add (&o->acc[B_BW_REF], c, bw_ref, 1.);
      acc_add (&o->acc[B_BW_TEST], c, bw_test, 1.);
    }
    if (c < PEAQ_ORACLE_TRACE_CHANNELS) {
      t->bw_ref[c] = (int) bw_ref;
      t->bw_test[c] = (int) bw_test;
    }
  }
}

/* movs.c:971-1023 */
static void
mov_nmr (PeaqOracle *o, Accum *acc_nmr, Accum *acc_rdf, PeaqOracleFftTrace *t)
{
  const FftModel *m = &o->fft;
  int c, i, B = m->bands.B;
  for (c = 0; c < acc_nmr->channels; c++) {
    const double *ref_exc = o->ref_fft[c].excitation;
    const double *rw = o->ref_fft[c].weighted, *tw = o->test_fft[c].weighted;
    double nmr = 0., nmr_max = 0.;
    double noise_in_bands[MAXB];
    double noise_spectrum[NBINS];
    for (i = 0; i < NBINS; i++)
      noise_spectrum[i] = rw[i] - 2 * sqrt (rw[i] * tw[i]) + tw[i];
    fft_group (m, noise_spectrum, noise_in_bands);
    for (i = 0; i < B; i++) {
      double mask = ref_exc[i] / m->maskdiff[i];
      double curr = noise_in_bands[i] / mask;
      nmr += curr;
      if (curr > nmr_max)
        nmr_max = curr;
    }
    nmr /= B;
    if (acc_nmr->mode == ACC_AVG_LOG)
      acc_add (acc_nmr, c, nmr, 1.);
    else
      acc_add (acc_nmr, c, 10. * log10 (nmr), 1.);
    if (acc_rdf)
      acc_add (acc_rdf, c, nmr_max > 1.41253754462275 ? 1. : 0., 1.);
    if (c < PEAQ_ORACLE_TRACE_CHANNELS) {
      memcpy (t->noise_in_bands[c], noise_in_bands, sizeof (double) * B);
      t->nmr[c] = nmr;
      t->nmr_max[c] = nmr_max;
    }
  }
}

/* movs.c:1224-1276 */
static void
mov_prob_detect (PeaqOracle *o, PeaqOracleFftTrace *t)
{
  int c, i, B = o->fft.bands.B;
  double bin_p = 1., bin_q = 0.;
  for (i = 0; i < B; i++) {
    double p = 0., q = 0.;
    for (c = 0; c < o->channels; c++) {
      double eref_db = 10. * log10 (o->ref_fft[c].excitation[i]);
      double etest_db = 10. * log10 (o->test_fft[c].excitation[i]);
      double l = 0.3 * DMAX (eref_db, etest_db) + 0.7 * etest_db;
      double s = l > 0. ? 5.95072 * pow (6.39468 / l, 1.71332) +
        9.01033e-11 * pow (l, 4.) + 5.05622e-6 * pow (l, 3.) -
        0.00102438 * l * l + 0.0550197 * l - 0.198719 : 1e30;
      double e = eref_db - etest_db;
      double b = eref_db > etest_db ? 4. : 6.;
      double pc = 1. - pow (0.5, pow (e / s, b));
      double qc = fabs (trunc (e)) / s;
      if (pc > p)
        p = pc;
      if (c == 0 || qc > q)
        q = qc;
    }
    bin_p *= 1. - p;
    bin_q += q;
  }
  bin_p = 1. - bin_p;
  if (bin_p > 0.5)
    acc_add (&o->acc[B_ADB], 0, bin_q, 1.);
  acc_add (&o->acc[B_MFPD], 0, bin_p, 1.);
  t->adb_steps = bin_q;
  t->det_prob = bin_p;
}

/* movs.c:1279-1315 */
static void
ehs_xcorr (PeaqOracle *o, const double *d, double *c)
{
  int k;
  double timedata[2 * MAXLAG];
  PeaqOracleComplex f1[MAXLAG + 1], f2[MAXLAG + 1];
  memcpy (timedata, d, 2 * MAXLAG * sizeof (double));
  peaq_oracle_fft_forward (o->fft512, timedata, f1);
  memset (timedata + MAXLAG, 0, MAXLAG * sizeof (double));
  peaq_oracle_fft_forward (o->fft512, timedata, f2);
  for (k = 0; k < MAXLAG + 1; k++) {
    double r = (f1[k].r * f2[k].r + f1[k].i * f2[k].i) / (2 * MAXLAG);
    double i = (f2[k].r * f1[k].i - f1[k].r * f2[k].i) / (2 * MAXLAG);
    f1[k].r = r;
    f1[k].i = i;
  }
  peaq_oracle_fft_inverse (o->fft512, f1, timedata);
  memcpy (c, timedata, MAXLAG * sizeof (double));
}

/* movs.c:1346-1443 */
static void
mov_ehs (PeaqOracle *o, Accum *acc, PeaqOracleFftTrace *t)
{
  int i, chan, valid = 0;
  for (chan = 0; chan < acc->channels; chan++)
    if (o->ref_fft[chan].energy_flag || o->test_fft[chan].energy_flag)
      valid = 1;
  t->ehs_valid = valid;
  if (!valid)
    return;
  for (chan = 0; chan < acc->channels; chan++) {
    const double *rp = o->ref_fft[chan].weighted;
    const double *tp = o->test_fft[chan].weighted;
    double d[NBINS], c[MAXLAG];
    double d0, dk, ehs = 0., cavg = 0., s;
    PeaqOracleComplex c_fft[MAXLAG / 2 + 1];
    for (i = 0; i < 2 * MAXLAG; i++) {
      double fref = rp[i], ftest = tp[i];
      if (fref == 0. && ftest == 0.)
        d[i] = 0.;
      else
        d[i] = log (ftest / fref);
    }
    ehs_xcorr (o, d, c);
    d0 = c[0];
    dk = d0;
    for (i = 0; i < MAXLAG; i++) {
      c[i] /= sqrt (d0 * dk);
      cavg += c[i];
      dk += d[i + MAXLAG] * d[i + MAXLAG] - d[i] * d[i];
    }
    cavg /= MAXLAG;
    for (i = 0; i < MAXLAG; i++)
      c[i] = (c[i] - cavg) * o->ehs_window[i];
    peaq_oracle_fft_forward (o->fft256, c, c_fft);
    s = c_fft[0].r * c_fft[0].r + c_fft[0].i * c_fft[0].i;
    for (i = 1; i < MAXLAG / 2 + 1; i++) {
      double new_s = c_fft[i].r * c_fft[i].r + c_fft[i].i * c_fft[i].i;
      if (new_s > s && new_s > ehs)
        ehs = new_s;
      s = new_s;
    }
    acc_add (acc, chan, 1000. * ehs, 1.);
    if (chan < PEAQ_ORACLE_TRACE_CHANNELS)
      t->ehs[chan] = ehs;
  }
}

/* apply_ear_model, gstpeaq.c:794-812 */
static void
deinterleave (const float *data, unsigned channels, unsigned c, unsigned n,
              float *out)
{
  unsigned i;
  for (i = 0; i < n; i++)
    out[i] = data[channels * i + c];
}

static void
apply_fft_model (PeaqOracle *o, const float *data, FftState *st)
{
  float mono[NFFT];
  int c;
  for (c = 0; c < o->channels; c++) {
    deinterleave (data, o->channels, c, NFFT, mono);
    fft_process (&o->fft, &st[c], mono);
  }
}

static void
apply_fb_model (PeaqOracle *o, const float *data, FbState *st)
{
  float mono[FB_FRAME];
  int c;
  for (c = 0; c < o->channels; c++) {
    deinterleave (data, o->channels, c, FB_FRAME, mono);
    fb_process (&o->fb, &st[c], mono);
  }
}

/* apply_ear_model_and_preprocess, gstpeaq.c:815-847 (after the ear models) */
static void
preprocess (PeaqOracle *o, unsigned frame_counter)
{
  const Bands *b = proc_bands (o);
  int c;
  for (c = 0; c < o->channels; c++) {
    const double *re, *te, *ru, *tu;
    if (o->advanced) {
      re = o->ref_fb[c].excitation;
      te = o->test_fb[c].excitation;
      ru = o->ref_fb[c].unsmeared;
      tu = o->test_fb[c].unsmeared;
    } else {
      re = o->ref_fft[c].excitation;
      te = o->test_fft[c].excitation;
      ru = o->ref_fft[c].unsmeared;
      tu = o->test_fft[c].unsmeared;
    }
    level_process (b, &o->level[c], re, te);
    mod_process (b, &o->ref_mod[c], ru);
    mod_process (b, &o->test_mod[c], tu);
    if (o->loudness_reached_frame == UINT_MAX) {
      if (calc_loudness (b, re) > 0.1 && calc_loudness (b, te) > 0.1)
        o->loudness_reached_frame = frame_counter;
    }
  }
}

/* gstpeaq.c:913-918 */
static void
snr_sums (PeaqOracle *o, const float *refdata, const float *testdata)
{
  unsigned i, n = (unsigned) o->channels * NFFT / 2;
  for (i = 0; i < n; i++) {
    o->total_signal_energy += refdata[i] * refdata[i];
    o->total_noise_energy +=
      (refdata[i] - testdata[i]) * (refdata[i] - testdata[i]);
  }
}

static void
trace_fft_states (PeaqOracle *o, PeaqOracleFftTrace *t)
{
  int c, B = o->fft.bands.B;
  for (c = 0; c < o->channels && c < PEAQ_ORACLE_TRACE_CHANNELS; c++) {
    memcpy (t->unsmeared[0][c], o->ref_fft[c].unsmeared, sizeof (double) * B);
    memcpy (t->unsmeared[1][c], o->test_fft[c].unsmeared, sizeof (double) * B);
    memcpy (t->excitation[0][c], o->ref_fft[c].excitation, sizeof (double) * B);
    memcpy (t->excitation[1][c], o->test_fft[c].excitation, sizeof (double) * B);
    t->energy_flag[0][c] = o->ref_fft[c].energy_flag;
    t->energy_flag[1][c] = o->test_fft[c].energy_flag;
  }
}

static PeaqOracleFftTrace *
fft_trace_slot (PeaqOracle *o)
{
  PeaqOracleFftTrace *t = &o->scratch_fft;
  if (o->fft_trace && o->frame_counter < o->fft_trace_cap)
    t = &o->fft_trace[o->frame_counter];
  memset (t, 0, sizeof *t);
  t->frame = (int) o->frame_counter;
  return t;
}

/* process_fft_block_basic, gstpeaq.c:850-921 */
static void
process_fft_block_basic (PeaqOracle *o, const float *refdata,
                         const float *testdata)
{
  int i, c;
  PeaqOracleFftTrace *t = fft_trace_slot (o);
  int above = frame_above_threshold (refdata, NFFT, (unsigned) o->channels);
  t->above_threshold = above;
  for (i = 0; i < N_BASIC; i++)
    acc_set_tentative (&o->acc[i], !above);

  apply_fft_model (o, refdata, o->ref_fft);
  apply_fft_model (o, testdata, o->test_fft);
  preprocess (o, o->frame_counter);
  trace_fft_states (o, t);

  if (o->frame_counter >= 24)
    mov_modulation_difference (o, &o->acc[B_AVG_MOD_DIFF_1],
                               &o->acc[B_AVG_MOD_DIFF_2],
                               &o->acc[B_WIN_MOD_DIFF], t->mod_diff1,
                               t->mod_diff2, t->temp_wt);

  if (o->frame_counter >= 24 &&
      o->frame_counter - 3 >= o->loudness_reached_frame) {
    /* peaq_mov_noise_loudness, movs.c:354-371 */
    const Bands *b = proc_bands (o);
    for (c = 0; c < o->acc[B_RMS_NOISE_LOUD].channels; c++) {
      double nl = noise_loudness (b, 1.5, 0.15, 0.5, 0., o->ref_mod[c].mod,
                                  o->test_mod[c].mod, o->level[c].ad_ref,
                                  o->level[c].ad_test);
      acc_add (&o->acc[B_RMS_NOISE_LOUD], c, nl, 1.);
      if (c < PEAQ_ORACLE_TRACE_CHANNELS)
        t->noise_loud[c] = nl;
    }
  }

  mov_bandwidth (o, t);
  mov_nmr (o, &o->acc[B_TOTAL_NMR], &o->acc[B_REL_DIST_FRAMES], t);
  mov_prob_detect (o, t);
  mov_ehs (o, &o->acc[B_EHS], t);
  snr_sums (o, refdata, testdata);
  t->signal_energy = o->total_signal_energy;
  t->noise_energy = o->total_noise_energy;
  o->frame_counter++;
}

/* process_fft_block_advanced, gstpeaq.c:924-962 */
static void
process_fft_block_advanced (PeaqOracle *o, const float *refdata,
                            const float *testdata)
{
  PeaqOracleFftTrace *t = fft_trace_slot (o);
  int above = frame_above_threshold (refdata, NFFT, (unsigned) o->channels);
  t->above_threshold = above;
  acc_set_tentative (&o->acc[ADV_SEGMENTAL_NMR], !above);
  acc_set_tentative (&o->acc[ADV_EHS], !above);
  apply_fft_model (o, refdata, o->ref_fft);
  apply_fft_model (o, testdata, o->test_fft);
  trace_fft_states (o, t);
  mov_nmr (o, &o->acc[ADV_SEGMENTAL_NMR], NULL, t);
  mov_ehs (o, &o->acc[ADV_EHS], t);
  snr_sums (o, refdata, testdata);
  t->signal_energy = o->total_signal_energy;
  t->noise_energy = o->total_noise_energy;
  o->frame_counter++;
}

/* process_fb_block, gstpeaq.c:965-1010 */
static void
process_fb_block (PeaqOracle *o, const float *refdata, const float *testdata)
{
  const Bands *b = &o->fb.bands;
  int c;
  PeaqOracleFbTrace *t = &o->scratch_fb;
  int above = frame_above_threshold (refdata, FB_FRAME, (unsigned) o->channels);
  if (o->fb_trace && o->frame_counter_fb < o->fb_trace_cap)
    t = &o->fb_trace[o->frame_counter_fb];
  memset (t, 0, sizeof *t);
  t->frame = (int) o->frame_counter_fb;
  t->above_threshold = above;
  acc_set_tentative (&o->acc[ADV_RMS_MOD_DIFF], !above);
  acc_set_tentative (&o->acc[ADV_RMS_NOISE_LOUD_ASYM], !above);
  acc_set_tentative (&o->acc[ADV_AVG_LIN_DIST], !above);

  apply_fb_model (o, refdata, o->ref_fb);
  apply_fb_model (o, testdata, o->test_fb);
  preprocess (o, o->frame_counter_fb);
  for (c = 0; c < o->channels && c < PEAQ_ORACLE_TRACE_CHANNELS; c++) {
    memcpy (t->unsmeared[0][c], o->ref_fb[c].unsmeared, sizeof (double) * 40);
    memcpy (t->unsmeared[1][c], o->test_fb[c].unsmeared, sizeof (double) * 40);
    memcpy (t->excitation[0][c], o->ref_fb[c].excitation, sizeof (double) * 40);
    memcpy (t->excitation[1][c], o->test_fb[c].excitation, sizeof (double) * 40);
  }

  if (o->frame_counter_fb >= 125)
    mov_modulation_difference (o, &o->acc[ADV_RMS_MOD_DIFF], NULL, NULL,
                               t->mod_diff, NULL, t->temp_wt);

  if (o->frame_counter_fb >= 125 &&
      o->frame_counter_fb - 13 >= o->loudness_reached_frame) {
    for (c = 0; c < o->channels; c++) {
      /* peaq_mov_noise_loud_asym, movs.c:551-577 (mod patterns swapped for
       * the missing-components term, settings.h:47) */
      const double *ar = o->level[c].ad_ref, *at = o->level[c].ad_test;
      const double *mr = o->ref_mod[c].mod, *mt = o->test_mod[c].mod;
      double nl = noise_loudness (b, 2.5, 0.3, 1., 0.1, mr, mt, ar, at);
      double mc = noise_loudness (b, 1.5, 0.15, 1., 0., mt, mr, at, ar);
      /* peaq_mov_lin_dist, movs.c:679-706 (ref modulation on both sides) */
      double ld = noise_loudness (b, 1.5, 0.15, 1., 0., mr, mr, ar,
                                  o->ref_fb[c].excitation);
      acc_add (&o->acc[ADV_RMS_NOISE_LOUD_ASYM], c, nl, mc);
      acc_add (&o->acc[ADV_AVG_LIN_DIST], c, ld, 1.);
      if (c < PEAQ_ORACLE_TRACE_CHANNELS) {
        t->noise_loud[c] = nl;
        t->missing_comp[c] = mc;
        t->lin_dist[c] = ld;
      }
    }
  }
  o->frame_counter_fb++;
}

/* ------------------------------------------------------------------------ */
/* construction / framing                                                    */

PeaqOracle *
peaq_oracle_new (int advanced, double playback_level, int channels)
{
  int i;
  PeaqOracle *o;
  if (channels < 1 || channels > MAXCH)
    return NULL;
  o = (PeaqOracle *) calloc (1, sizeof *o);
  o->advanced = advanced ? 1 : 0;
  o->channels = channels;
  o->loudness_reached_frame = UINT_MAX;
  fft_model_init (&o->fft, advanced ? 55 : 109, playback_level);
  fb_model_init (&o->fb, playback_level);
  o->ref_fft = (FftState *) calloc ((size_t) channels, sizeof (FftState));
  o->test_fft = (FftState *) calloc ((size_t) channels, sizeof (FftState));
  o->level = (LevelState *) calloc ((size_t) channels, sizeof (LevelState));
  o->ref_mod = (ModState *) calloc ((size_t) channels, sizeof (ModState));
  o->test_mod = (ModState *) calloc ((size_t) channels, sizeof (ModState));
  if (advanced) {
    o->ref_fb = (FbState *) calloc ((size_t) channels, sizeof (FbState));
    o->test_fb = (FbState *) calloc ((size_t) channels, sizeof (FbState));
  }
  /* accumulator modes, gstpeaq.c:527-557; channel counts :580-584 */
  if (advanced) {
    acc_init (&o->acc[ADV_RMS_MOD_DIFF], ACC_RMS, channels);
    acc_init (&o->acc[ADV_RMS_NOISE_LOUD_ASYM], ACC_RMS_ASYM, channels);
    acc_init (&o->acc[ADV_SEGMENTAL_NMR], ACC_AVG, channels);
    acc_init (&o->acc[ADV_EHS], ACC_AVG, channels);
    acc_init (&o->acc[ADV_AVG_LIN_DIST], ACC_AVG, channels);
  } else {
    acc_init (&o->acc[B_BW_REF], ACC_AVG, channels);
    acc_init (&o->acc[B_BW_TEST], ACC_AVG, channels);
    acc_init (&o->acc[B_TOTAL_NMR], ACC_AVG_LOG, channels);
    acc_init (&o->acc[B_WIN_MOD_DIFF], ACC_AVG_WINDOW, channels);
    acc_init (&o->acc[B_ADB], ACC_ADB, 1);
    acc_init (&o->acc[B_EHS], ACC_AVG, channels);
    acc_init (&o->acc[B_AVG_MOD_DIFF_1], ACC_AVG, channels);
    acc_init (&o->acc[B_AVG_MOD_DIFF_2], ACC_AVG, channels);
    acc_init (&o->acc[B_RMS_NOISE_LOUD], ACC_RMS, channels);
    acc_init (&o->acc[B_MFPD], ACC_FILTERED_MAX, 1);
    acc_init (&o->acc[B_REL_DIST_FRAMES], ACC_AVG, channels);
  }
  /* EHS correlation window, movs.c:1366-1367 */
  for (i = 0; i < MAXLAG; i++)
    o->ehs_window[i] = 0.81649658092773 *
      (1 - cos (2 * M_PI * i / (MAXLAG - 1))) / MAXLAG;
  o->fft512 = peaq_oracle_fft_new (2 * MAXLAG);
  o->fft256 = peaq_oracle_fft_new (MAXLAG);
  return o;
}

void
peaq_oracle_free (PeaqOracle *o)
{
  if (!o)
    return;
  fft_model_clear (&o->fft);
  fb_model_clear (&o->fb);
  peaq_oracle_fft_free (o->fft512);
  peaq_oracle_fft_free (o->fft256);
  free (o->ref_fft);
  free (o->test_fft);
  free (o->ref_fb);
  free (o->test_fb);
  free (o->level);
  free (o->ref_mod);
  free (o->test_mod);
  free (o->q_ref_fft.data);
  free (o->q_test_fft.data);
  free (o->q_ref_fb.data);
  free (o->q_test_fb.data);
  free (o);
}

void
peaq_oracle_set_fft_trace (PeaqOracle *o, PeaqOracleFftTrace *buf,
                           size_t capacity)
{
  o->fft_trace = buf;
  o->fft_trace_cap = capacity;
}

void
peaq_oracle_set_fb_trace (PeaqOracle *o, PeaqOracleFbTrace *buf,
                          size_t capacity)
{
  o->fb_trace = buf;
  o->fb_trace_cap = capacity;
}

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
fifo_drop (Fifo *f, size_t n)
{
  memmove (f->data, f->data + n, (f->len - n) * sizeof (float));
  f->len -= n;
}

typedef void (*BlockFn) (PeaqOracle *, const float *, const float *);

/* do_processing, gstpeaq.c:596-611 */
static void
run_frames (PeaqOracle *o, Fifo *ref, Fifo *test, BlockFn fn, size_t frame,
            size_t step)
{
  size_t pos = 0;
  while (ref->len - pos >= frame && test->len - pos >= frame) {
    fn (o, ref->data + pos, test->data + pos);
    pos += step;
  }
  fifo_drop (ref, pos);
  fifo_drop (test, pos);
}

/* pad_chain, gstpeaq.c:626-652 */
void
peaq_oracle_push (PeaqOracle *o, const float *ref, size_t n_ref,
                  const float *test, size_t n_test)
{
  size_t ch = (size_t) o->channels;
  fifo_push (&o->q_ref_fft, ref, n_ref * ch);
  fifo_push (&o->q_test_fft, test, n_test * ch);
  if (o->advanced) {
    fifo_push (&o->q_ref_fb, ref, n_ref * ch);
    fifo_push (&o->q_test_fb, test, n_test * ch);
    run_frames (o, &o->q_ref_fft, &o->q_test_fft, process_fft_block_advanced,
                ch * NFFT, ch * NFFT / 2);
    run_frames (o, &o->q_ref_fb, &o->q_test_fb, process_fb_block,
                ch * FB_FRAME, ch * FB_FRAME);
  } else {
    run_frames (o, &o->q_ref_fft, &o->q_test_fft, process_fft_block_basic,
                ch * NFFT, ch * NFFT / 2);
  }
}

/* do_flush, gstpeaq.c:716-745: ONE zero-padded frame per FIFO pair */
static void
flush_frames (PeaqOracle *o, Fifo *ref, Fifo *test, BlockFn fn, size_t frame)
{
  if (ref->len || test->len) {
    float *pr = (float *) calloc (frame, sizeof (float));
    float *pt = (float *) calloc (frame, sizeof (float));
    size_t nr = ref->len < frame ? ref->len : frame;
    size_t nt = test->len < frame ? test->len : frame;
    if (nr)
      memcpy (pr, ref->data, nr * sizeof (float));
    if (nt)
      memcpy (pt, test->data, nt * sizeof (float));
    fn (o, pr, pt);
    fifo_drop (ref, nr);
    fifo_drop (test, nt);
    free (pr);
    free (pt);
  }
}

/* change_state PAUSED->READY, gstpeaq.c:764-778 */
void
peaq_oracle_finish (PeaqOracle *o)
{
  size_t ch = (size_t) o->channels;
  if (o->advanced) {
    flush_frames (o, &o->q_ref_fft, &o->q_test_fft, process_fft_block_advanced,
                  ch * NFFT);
    flush_frames (o, &o->q_ref_fb, &o->q_test_fb, process_fb_block,
                  ch * FB_FRAME);
  } else {
    flush_frames (o, &o->q_ref_fft, &o->q_test_fft, process_fft_block_basic,
                  ch * NFFT);
  }
}

/* calculate_di_basic/advanced + calculate_odg, gstpeaq.c:1012-1078;
 * totalsnr :493-497 */
void
peaq_oracle_result (const PeaqOracle *o, PeaqOracleResult *out)
{
  int i, n = o->advanced ? N_ADV : N_BASIC;
  memset (out, 0, sizeof *out);
  for (i = 0; i < n; i++)
    out->movs[i] = acc_value (&o->acc[i]);
  out->n_movs = n;
  out->di = o->advanced ? di_advanced (out->movs) : di_basic (out->movs);
  out->odg = odg_from_di (out->di);
  out->totalsnr = 10 * log10 (o->total_signal_energy / o->total_noise_energy);
  out->frames_fft = o->frame_counter;
  out->frames_fb = o->frame_counter_fb;
  out->loudness_reached_frame = o->loudness_reached_frame;
}

void
peaq_oracle_run_pair (int advanced, double playback_level, int channels,
                      const float *ref, size_t n_ref, const float *test,
                      size_t n_test, PeaqOracleResult *out)
{
  PeaqOracle *o = peaq_oracle_new (advanced, playback_level, channels);
  /* feed in bounded chunks so the FIFOs stay small for hour-long items */
  size_t pos = 0, chunk = 1 << 16;
  size_t n_max = n_ref > n_test ? n_ref : n_test;
  while (pos < n_max) {
    size_t r = pos < n_ref ? (n_ref - pos < chunk ? n_ref - pos : chunk) : 0;
    size_t t = pos < n_test ? (n_test - pos < chunk ? n_test - pos : chunk) : 0;
    peaq_oracle_push (o, ref + pos * (size_t) channels, r,
                      test + pos * (size_t) channels, t);
    pos += chunk;
  }
  peaq_oracle_finish (o);
  peaq_oracle_result (o, out);
  peaq_oracle_free (o);
}

int
peaq_oracle_table (const PeaqOracle *o, int model, int which, double *out)
{
  const Bands *b = model ? &o->fb.bands : &o->fft.bands;
  const FftModel *m = &o->fft;
  int i, n = b->B;
  for (i = 0; i < n; i++) {
    double v;
    switch (which) {
      case 0: v = b->fc[i]; break;
      case 1: v = b->internal_noise[i]; break;
      case 2: v = b->a_ear[i]; break;
      case 3: v = b->ethres[i]; break;
      case 4: v = b->thres[i]; break;
      case 5: v = b->loudfac[i]; break;
      case 14: v = b->a_proc[i]; break;
      default:
        if (model)
          return 0;
        switch (which) {
          case 6: v = m->maskdiff[i]; break;
          case 7: v = m->aUC[i]; break;
          case 8: v = m->gIL[i]; break;
          case 9: v = m->norm[i]; break;
          case 10: v = m->wl[i]; break;
          case 11: v = m->wu[i]; break;
          case 12: v = m->lo[i]; break;
          case 13: v = m->hi[i]; break;
          default: return 0;
        }
    }
    out[i] = v;
  }
  return n;
}

/* ------------------------------------------------------------------------ */
/* single stages for the golden vectors of testpeaq.c                        */

struct _PeaqOracleStage
{
  int bands;
  FftModel fft;
  FbModel fb;
  FftState fft_state;
  FbState fb_state;
  LevelState level;
  ModState mod;
};

PeaqOracleStage *
peaq_oracle_stage_new (int bands)
{
  PeaqOracleStage *s = (PeaqOracleStage *) calloc (1, sizeof *s);
  s->bands = bands;
  fft_model_init (&s->fft, bands == 40 ? 109 : bands, 92.);
  fb_model_init (&s->fb, 92.);
  return s;
}

void
peaq_oracle_stage_free (PeaqOracleStage *s)
{
  if (!s)
    return;
  fft_model_clear (&s->fft);
  fb_model_clear (&s->fb);
  free (s);
}

void
peaq_oracle_stage_fft_ear (PeaqOracleStage *s, const float *frame,
                           double *power_spectrum, double *weighted,
                           double *unsmeared, double *excitation)
{
  int B = s->fft.bands.B;
  fft_process (&s->fft, &s->fft_state, frame);
  if (power_spectrum)
    memcpy (power_spectrum, s->fft_state.power, sizeof (double) * NBINS);
  if (weighted)
    memcpy (weighted, s->fft_state.weighted, sizeof (double) * NBINS);
  if (unsmeared)
    memcpy (unsmeared, s->fft_state.unsmeared, sizeof (double) * B);
  if (excitation)
    memcpy (excitation, s->fft_state.excitation, sizeof (double) * B);
}

void
peaq_oracle_stage_fb_ear (PeaqOracleStage *s, const float *frame,
                          double *unsmeared, double *excitation)
{
  fb_process (&s->fb, &s->fb_state, frame);
  if (unsmeared)
    memcpy (unsmeared, s->fb_state.unsmeared, sizeof (double) * FB_BANDS);
  if (excitation)
    memcpy (excitation, s->fb_state.excitation, sizeof (double) * FB_BANDS);
}

double
peaq_oracle_stage_loudness (PeaqOracleStage *s, int filterbank)
{
  return filterbank ? calc_loudness (&s->fb.bands, s->fb_state.excitation)
    : calc_loudness (&s->fft.bands, s->fft_state.excitation);
}

void
peaq_oracle_stage_level_adapt (PeaqOracleStage *s, const double *ref_exc,
                               const double *test_exc, double *ref_out,
                               double *test_out)
{
  const Bands *b = s->bands == 40 ? &s->fb.bands : &s->fft.bands;
  level_process (b, &s->level, ref_exc, test_exc);
  memcpy (ref_out, s->level.ad_ref, sizeof (double) * b->B);
  memcpy (test_out, s->level.ad_test, sizeof (double) * b->B);
}

void
peaq_oracle_stage_modulation (PeaqOracleStage *s, const double *unsmeared,
                              double *modulation, double *avg_loudness)
{
  const Bands *b = s->bands == 40 ? &s->fb.bands : &s->fft.bands;
  mod_process (b, &s->mod, unsmeared);
  memcpy (modulation, s->mod.mod, sizeof (double) * b->B);
  memcpy (avg_loudness, s->mod.filt_loud, sizeof (double) * b->B);
}

/* layout check for the ctypes/numpy mirrors in tests/refharness.py */
size_t
peaq_oracle_sizeof (int which)
{
  switch (which) {
    case 0: return sizeof (PeaqOracleFftTrace);
    case 1: return sizeof (PeaqOracleFbTrace);
    case 2: return sizeof (PeaqOracleResult);
    default: return 0;
  }
}
