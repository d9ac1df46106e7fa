/* oracle/shim/gst/fft/gstfftf64.h -- TEST INFRASTRUCTURE ONLY.
 * GstFFTF64 facade over oracle/fft_f64.c (same conventions as libgstfft). */
#ifndef PEAQ_ORACLE_SHIM_GSTFFTF64_H
#define PEAQ_ORACLE_SHIM_GSTFFTF64_H

#include <glib.h>
#include "fft_f64.h"

typedef PeaqOracleFFT GstFFTF64;
typedef struct { gdouble r, i; } GstFFTF64Complex;

static inline GstFFTF64 *
gst_fft_f64_new (gint len, gboolean inverse)
{
  (void) inverse;
  return peaq_oracle_fft_new (len);
}

static inline void
gst_fft_f64_free (GstFFTF64 *self)
{
  peaq_oracle_fft_free (self);
}

static inline void
gst_fft_f64_fft (GstFFTF64 *self, const gdouble *timedata,
                 GstFFTF64Complex *freqdata)
{
  peaq_oracle_fft_forward (self, timedata, (PeaqOracleComplex *) freqdata);
}

static inline void
gst_fft_f64_inverse_fft (GstFFTF64 *self, const GstFFTF64Complex *freqdata,
                         gdouble *timedata)
{
  peaq_oracle_fft_inverse (self, (const PeaqOracleComplex *) freqdata,
                           timedata);
}

#endif
