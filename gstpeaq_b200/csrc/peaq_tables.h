// Constant tables of the PEAQ engine, built once on the host in double precision
// and uploaded to device global memory.  Every table cites the reference code
// (paths relative to /root/reference/src) whose values it must reproduce; the
// CPU test tests/test_tables.py compares them with the oracle's tables.
#pragma once

#include <cstdint>

namespace peaq {

constexpr int kMaxBands = 109;     // basic FFT ear model; advanced uses 55 (fftearmodel.c:692-705)
constexpr int kFbBands = 40;       // filter-bank ear model (fbearmodel.c:57-61)
constexpr int kFftFrame = 2048;    // fftearmodel.c:49
constexpr int kFftStep = 1024;
constexpr int kFftBins = 1025;
constexpr int kFbFrame = 192;      // fbearmodel.c:48
constexpr int kFbSub = 32;         // filter bank evaluated every 32 samples (fbearmodel.c:314)
constexpr int kFbBuf = 1456;       // fbearmodel.c:52
constexpr int kMaxLag = 256;       // movs.c:42
constexpr int kFbTapsTotal = 10954; // sum over bands of (N/2+1), N from Table 8 of BS.1387
constexpr int kFbHist = 1504;      // filtered samples kept in front of a chunk (>= 1456 + 32, multiple of 32)
constexpr int kFbGTotal = 21868;   // sum over bands of the delay support length (N-1, band 0: N)
constexpr int kFbRecBands = 40;    // all bands run as sliding windowed DFTs (the direct FIR kernel is kept as a cross-check)
constexpr int kFbRecGroup = 6;     // sub-steps per recursion group (= one 192-sample frame)

struct double2_t { double x, y; };

// Per-band constants shared by both ear models (earmodel.c:278-323).
struct BandTables {
  int B;
  int step;                         // samples per frame step (1024 / 192)
  double fc[kMaxBands];
  double internal_noise[kMaxBands];
  double internal_noise_pow03[kMaxBands];  // pow(internal_noise, 0.3), movs.c:241
  double ethres[kMaxBands];         // excitation threshold
  double thres[kMaxBands];          // threshold index
  double loudfac[kMaxBands];
  double a_ear[kMaxBands];          // ear-model smearing constant (earmodel.c:626-635)
  double a_proc[kMaxBands];         // level adapter / modulation constant (tau 8/50 ms)
};

// Everything the frame kernels and the scan kernels read.  One instance per
// mode (basic / advanced) and playback level.
struct DeviceTables {
  int advanced;
  int fft_bands;                    // 109 / 55
  double level_factor_fft;          // fftearmodel.c:304-314
  double level_factor_fb;           // fbearmodel.c:248-254
  double dz;                        // band width in Bark
  double aLe;                       // lower spreading slope ^0.4 (fftearmodel.c:726-728)
  BandTables fft;                   // 109- or 55-band layout
  BandTables fb;                    // 40-band layout
  alignas(16) double hann[kFftFrame];  // fftearmodel.c:159-173
  double earw2[kFftBins];           // squared outer/middle ear weight (fftearmodel.c:251-256)
  int band_lo[kMaxBands];           // fftearmodel.c:742-745
  int band_hi[kMaxBands];
  double band_wl[kMaxBands];
  double band_wu[kMaxBands];
  double aUC[kMaxBands];            // fftearmodel.c:765
  double log_aUC[kMaxBands];        // ln(aUC), for the exp/log form of the spreading slopes
  double gIL[kMaxBands];            // fftearmodel.c:766
  double spread_norm[kMaxBands];    // fftearmodel.c:778-781
  double maskdiff[kMaxBands];       // fftearmodel.c:770-772
  double ehs_window[kMaxLag];       // movs.c:1366-1367
  alignas(16) double2_t tw1024[768];   // exp(-2 pi i k / 1024), k < 768
  double2_t tw2048[kFftBins];       // exp(-2 pi i k / 2048), k <= 1024
  // filter bank (fbearmodel.c:188-225); taps of band b start at fb_tap_offset[b]
  int fb_len[kFbBands];
  int fb_tap_offset[kFbBands + 1];
  double fb_h_re[kFbTapsTotal];
  double fb_h_im[kFbTapsTotal];
  double fb_back_mask[6];           // fbearmodel.c:179-185
  // the same filters indexed by total delay d = D + n (D = 1 + (1456-N)/2,
  // fbearmodel.c:408), full length, phase-major (d = 32 q + j) for the polyphase
  // kernel: band b covers delays fb_dlo[b]..fb_dhi[b]; its coefficients start at
  // fb_g_offset[b] (complex entries), phase j at + fb_phase_offset[b*32+j]
  int fb_dlo[kFbBands];
  int fb_dhi[kFbBands];
  int fb_g_offset[kFbBands];
  int fb_phase_offset[kFbBands * 32];
  alignas(16) double fb_g[2 * kFbGTotal];
  // Long filters as recursions (see fb_bank_rec_kernel).  The taps are a raised-cosine
  // window times a complex exponential (fbearmodel.c:213-220), so with w = 2 pi fc / 48000,
  // d = 2 pi / N, w_f = {w, w + d, w - d}, g_f = {2, -1, -1} (Wt / N) e^{-j w N / 2}
  //   out[s] = sum_f S_f[s],   S_f[s] = g_f sum_{n < N} e^{j w_f n} x[32 s - D - n]
  //   S_f[s] = e^{j 32 w_f} S_f[s-1] + sum_{k < 32} (P_f[k] x[32 s - D - k] + Q_f[k] x[32 s - D - k - N])
  // fb_rec_ph[b][k] = {P_0, P_+, P_-, Q_0, Q_+, Q_-}[k], P_f[k] = g_f e^{j w_f k}, Q_f[k] = -e^{j w N} P_f[k];
  // fb_rec_rpow[b][f][i] = e^{j 32 w_f (i + 1)}, i < 6
  alignas(16) double2_t fb_rec_ph[kFbRecBands][32][6];
  alignas(16) double2_t fb_rec_rpow[kFbRecBands][3][kFbRecGroup];
  double2_t fb_rec_alias;           // band 0: tap at delay 1456, which the reference reads from the newest sample
  // neural network (nn.c:40-93)
  double nn_amin[11], nn_amax[11];
  double nn_wx[11 * 5];             // [input][hidden], row stride 5
  double nn_wxb[5];
  double nn_wy[5];
  double nn_wyb;
  int nn_inputs, nn_hidden;
};

// Fills `t` for the given mode; pure host code, double precision, libm.
void build_tables(DeviceTables* t, bool advanced, double playback_level);

}  // namespace peaq
