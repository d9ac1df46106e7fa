/* oracle/fft_f64.c -- TEST INFRASTRUCTURE ONLY; see fft_f64.h. */
#include "fft_f64.h"

#include <math.h>
#include <stdlib.h>

struct _PeaqOracleFFT
{
  int n;                        /* real length */
  int h;                        /* n / 2, complex length */
  int *rev;                     /* bit reversal of 0..h-1 */
  PeaqOracleComplex *w;         /* exp(-2 pi i k / h), k < h/2 */
  PeaqOracleComplex *ws;        /* exp(-2 pi i k / n), k <= h */
};

PeaqOracleFFT *
peaq_oracle_fft_new (int len)
{
  PeaqOracleFFT *f = (PeaqOracleFFT *) calloc (1, sizeof *f);
  int k, bits = 0;
  f->n = len;
  f->h = len / 2;
  while ((1 << bits) < f->h)
    bits++;
  f->rev = (int *) malloc (sizeof (int) * f->h);
  for (k = 0; k < f->h; k++) {
    int r = 0, b;
    for (b = 0; b < bits; b++)
      if (k & (1 << b))
        r |= 1 << (bits - 1 - b);
    f->rev[k] = r;
  }
  f->w = (PeaqOracleComplex *) malloc (sizeof (PeaqOracleComplex) * (f->h / 2 + 1));
  for (k = 0; k < f->h / 2; k++) {
    f->w[k].r = cos (-2. * M_PI * k / f->h);
    f->w[k].i = sin (-2. * M_PI * k / f->h);
  }
  f->ws = (PeaqOracleComplex *) malloc (sizeof (PeaqOracleComplex) * (f->h + 1));
  for (k = 0; k <= f->h; k++) {
    f->ws[k].r = cos (-2. * M_PI * k / f->n);
    f->ws[k].i = sin (-2. * M_PI * k / f->n);
  }
  return f;
}

void
peaq_oracle_fft_free (PeaqOracleFFT *f)
{
  if (!f)
    return;
  free (f->rev);
  free (f->w);
  free (f->ws);
  free (f);
}

/* in-place forward complex FFT of length f->h on bit-reversed input */
static void
cfft (const PeaqOracleFFT *f, PeaqOracleComplex *z)
{
  int h = f->h, len, i, j;
  for (len = 2; len <= h; len <<= 1) {
    int half = len >> 1, stride = h / len;
    for (i = 0; i < h; i += len)
      for (j = 0; j < half; j++) {
        PeaqOracleComplex w = f->w[j * stride];
        PeaqOracleComplex a = z[i + j], b = z[i + j + half], t;
        t.r = b.r * w.r - b.i * w.i;
        t.i = b.r * w.i + b.i * w.r;
        z[i + j].r = a.r + t.r;
        z[i + j].i = a.i + t.i;
        z[i + j + half].r = a.r - t.r;
        z[i + j + half].i = a.i - t.i;
      }
  }
}

void
peaq_oracle_fft_forward (const PeaqOracleFFT *f, const double *x,
                         PeaqOracleComplex *X)
{
  int h = f->h, k;
  PeaqOracleComplex *z = (PeaqOracleComplex *) malloc (sizeof *z * h);
  for (k = 0; k < h; k++) {
    z[f->rev[k]].r = x[2 * k];
    z[f->rev[k]].i = x[2 * k + 1];
  }
  cfft (f, z);
  /* X[k] = (Z[k] + conj Z[h-k])/2 - i/2 w^k (Z[k] - conj Z[h-k]) */
  for (k = 0; k <= h; k++) {
    PeaqOracleComplex a = z[k % h], b = z[(h - k) % h], w = f->ws[k];
    double er = 0.5 * (a.r + b.r), ei = 0.5 * (a.i - b.i);
    double or_ = 0.5 * (a.i + b.i), oi = -0.5 * (a.r - b.r);
    X[k].r = er + or_ * w.r - oi * w.i;
    X[k].i = ei + or_ * w.i + oi * w.r;
  }
  free (z);
}

void
peaq_oracle_fft_inverse (const PeaqOracleFFT *f, const PeaqOracleComplex *X,
                         double *x)
{
  /* x[n] = sum over the full Hermitian spectrum of X[k] exp(+2 pi i k n/N).
   * Build Z[k] = E[k] + i O[k] with E, O the even/odd-sample spectra, run the
   * forward machinery on the conjugate, conjugate back. */
  int h = f->h, k;
  PeaqOracleComplex *z = (PeaqOracleComplex *) malloc (sizeof *z * h);
  for (k = 0; k < h; k++) {
    PeaqOracleComplex a = X[k], b = X[h - k], w = f->ws[k];
    /* E = (a + conj b), O = (a - conj b) * conj(w)   (factor 2 kept: result
     * is N times the true inverse) */
    double er = a.r + b.r, ei = a.i - b.i;
    double dr = a.r - b.r, di = a.i + b.i;
    double or_ = dr * w.r + di * w.i, oi = di * w.r - dr * w.i;
    /* Z = E + i O ; store conj(Z) bit-reversed */
    double zr = er - oi, zi = ei + or_;
    z[f->rev[k]].r = zr;
    z[f->rev[k]].i = -zi;
  }
  cfft (f, z);
  for (k = 0; k < h; k++) {
    x[2 * k] = z[k].r;
    x[2 * k + 1] = -z[k].i;
  }
  free (z);
}
