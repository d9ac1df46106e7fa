/* oracle/fft_f64.h -- TEST INFRASTRUCTURE ONLY (CPU oracle, never shipped).
 *
 * Double-precision real FFT with the conventions of GStreamer's GstFFTF64
 * (gst-plugins-base libgstfft, kissfft based; un-vendored third-party
 * dependency of the reference, version unpinned -- configure.ac:26-27):
 *   forward : N real samples -> N/2+1 complex bins, unnormalised
 *             X[k] = sum_n x[n] exp(-2 pi i k n / N)
 *   inverse : N/2+1 complex bins -> N real samples, unnormalised
 *             (N times the true inverse; the reference divides by N itself,
 *              /root/reference/src/movs.c:1306-1309)
 * Call sites in the reference: fftearmodel.c:244,457; movs.c:1284-1313,1355,1428.
 * Any correct double-precision DFT satisfies the reference's own golden
 * vectors (testpeaq.c:680-686, rel 5e-5); this one is a plain radix-2
 * half-length complex FFT plus the real-signal split.
 */
#ifndef PEAQ_ORACLE_FFT_F64_H
#define PEAQ_ORACLE_FFT_F64_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double r, i; } PeaqOracleComplex;
typedef struct _PeaqOracleFFT PeaqOracleFFT;

/* len must be a power of two >= 4 */
PeaqOracleFFT *peaq_oracle_fft_new (int len);
void peaq_oracle_fft_free (PeaqOracleFFT *f);
void peaq_oracle_fft_forward (const PeaqOracleFFT *f, const double *timedata,
                              PeaqOracleComplex *freqdata);
void peaq_oracle_fft_inverse (const PeaqOracleFFT *f,
                              const PeaqOracleComplex *freqdata,
                              double *timedata);

#ifdef __cplusplus
}
#endif
#endif
