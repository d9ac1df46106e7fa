// Host-side construction of the engine's constant tables (double precision).
// Formulas follow the reference so the values agree bit for bit with what its
// GObject constructors compute; citations are relative to /root/reference/src.
#include "peaq_tables.h"

#include <cmath>
#include <cstring>

namespace peaq {
namespace {

constexpr double kFs = 48000.0;

// BS.1387 Table 8 filter lengths (fbearmodel.c:57-61)
const int kFbLen[kFbBands] = {
    1456, 1438, 1406, 1362, 1308, 1244, 1176, 1104, 1030, 956, 884, 814, 748, 686,
    626,  570,  520,  472,  430,  390,  354,  320,  290,  262, 238, 214, 194, 176,
    158,  144,  130,  118,  106,  96,   86,   78,   70,   64,  58,  52};

// outer and middle ear transfer function (earmodel.c:701-709)
double ear_weight(double frequency) {
  const double f_khz = frequency / 1000.;
  const double w_db = -0.6 * 3.64 * std::pow(f_khz, -0.8) +
                      6.5 * std::exp(-0.6 * std::pow(f_khz - 3.3, 2)) -
                      1e-3 * std::pow(f_khz, 3.6);
  return std::pow(10, w_db / 20);
}

// first-order smoothing coefficient for one band (earmodel.c:626-635)
double time_constant(const BandTables& b, int band, double tau_min, double tau_100) {
  const double tau = tau_min + 100. / b.fc[band] * (tau_100 - tau_min);
  return std::exp(b.step / (-48000. * tau));
}

// per-band psychoacoustic constants (earmodel.c:278-323)
void fill_bands(BandTables* b, const double* fc, int count, int step,
                double loudness_scale, double tau_min, double tau_100) {
  b->B = count;
  b->step = step;
  for (int i = 0; i < count; i++) {
    const double f = fc[i];
    b->fc[i] = f;
    b->internal_noise[i] = std::pow(10., 0.4 * 0.364 * std::pow(f / 1000., -0.8));
    b->internal_noise_pow03[i] = std::pow(b->internal_noise[i], 0.3);
    b->ethres[i] = std::pow(10., 0.364 * std::pow(f / 1000., -0.8));
    b->thres[i] = std::pow(10., 0.1 * (-2. - 2.05 * std::atan(f / 4000.) -
                                      0.75 * std::atan(f / 1600. * f / 1600.)));
    b->loudfac[i] = loudness_scale * std::pow(b->ethres[i] / (1e4 * b->thres[i]), 0.23);
  }
  for (int i = 0; i < count; i++) {
    b->a_ear[i] = time_constant(*b, i, tau_min, tau_100);
    b->a_proc[i] = time_constant(*b, i, 0.008, 0.05);  // leveladapter.c:205, modpatt.c:185
  }
}

// Level dependent spreading on the host; used once on an all-ones pattern to
// obtain the normalisation exactly as the reference does (fftearmodel.c:636-676,
// :778-781).
void host_spread(const DeviceTables& t, const double* pp, double* e2) {
  const int B = t.fft_bands;
  double a_e[kMaxBands], en_e[kMaxBands];
  for (int i = 0; i < B; i++) {
    const double a = t.aUC[i] * std::pow(pp[i], 0.2 * t.dz);
    const double g_up = (1. - std::pow(a, B - i)) / (1. - a);
    const double en = pp[i] / (t.gIL[i] + g_up - 1.);
    a_e[i] = std::pow(a, 0.4);
    en_e[i] = std::pow(en, 0.4);
  }
  e2[B - 1] = en_e[B - 1];
  for (int i = B - 1; i > 0; i--) e2[i - 1] = t.aLe * e2[i] + en_e[i - 1];
  for (int i = 0; i < B - 1; i++) {
    double r = en_e[i];
    for (int j = i + 1; j < B; j++) {
      r *= a_e[i];
      e2[j] += r;
    }
  }
  for (int i = 0; i < B; i++) e2[i] = std::pow(e2[i], 1. / 0.4) / t.spread_norm[i];
}

void fill_fft_model(DeviceTables* t, double playback_level) {
  const int N = kFftFrame;
  const int B = t->fft_bands;
  const double gamma = 0.84971762641205;  // fftearmodel.c:50
  for (int k = 0; k < N; k++)
    t->hann[k] = std::sqrt(8. / 3.) * 0.5 * (1. - std::cos(2 * M_PI * k / (N - 1)));
  for (int k = 0; k <= N / 2; k++)
    t->earw2[k] = std::pow(ear_weight((double)k * kFs / N), 2);
  t->level_factor_fft = std::pow(10, playback_level / 10) /
                        (8. / 3. * (gamma / 4 * (N - 1)) * (gamma / 4 * (N - 1)));

  t->dz = 27. / (B - 1);
  const double z_lo = 7. * std::asinh(80. / 650.);
  const double z_hi = 7. * std::asinh(18000. / 650.);
  const double a_low = std::pow(10., -2.7 * t->dz);
  t->aLe = std::pow(a_low, 0.4);
  double fc[kMaxBands];
  for (int band = 0; band < B; band++) {
    const double zl = z_lo + band * t->dz;
    const double zu = std::fmin(z_hi, z_lo + (band + 1) * t->dz);
    const double zc = (zu + zl) / 2.;
    const double f_c = 650. * std::sinh(zc / 7.);
    const double fl = 650. * std::sinh(zl / 7.);
    const double fu = 650. * std::sinh(zu / 7.);
    fc[band] = f_c;
    // C round(): half away from zero (fftearmodel.c:742-745)
    t->band_lo[band] = (int)(unsigned)std::round(fl / kFs * N);
    t->band_hi[band] = (int)(unsigned)std::round(fu / kFs * N);
    double upper_freq = (2 * t->band_lo[band] + 1) / 2. * kFs / N;
    if (upper_freq > fu) upper_freq = fu;
    double u = upper_freq - fl;
    t->band_wl[band] = u * N / kFs;
    if (t->band_lo[band] == t->band_hi[band]) {
      t->band_wu[band] = 0;
    } else {
      const double lower_freq = (2 * t->band_hi[band] - 1) / 2. * kFs / N;
      u = fu - lower_freq;
      t->band_wu[band] = u * N / kFs;
    }
    t->aUC[band] = std::pow(10., (-2.4 - 23. / f_c) * t->dz);
    t->log_aUC[band] = std::log(t->aUC[band]);
    t->gIL[band] = (1. - std::pow(a_low, band + 1)) / (1. - a_low);
    t->spread_norm[band] = 1.;
    t->maskdiff[band] =
        std::pow(10., (band * t->dz <= 12. ? 3. : 0.25 * band * t->dz) / 10.);
  }
  fill_bands(&t->fft, fc, B, kFftStep, 1.07664, 0.008, 0.030);  // fftearmodel.c:51,226-228
  double ones[kMaxBands], spread[kMaxBands];
  for (int i = 0; i < B; i++) ones[i] = 1.;
  host_spread(*t, ones, spread);
  for (int i = 0; i < B; i++) t->spread_norm[i] = spread[i];

  for (int i = 0; i < kMaxLag; i++)
    t->ehs_window[i] =
        0.81649658092773 * (1 - std::cos(2 * M_PI * i / (kMaxLag - 1))) / kMaxLag;
}

void fill_fb_model(DeviceTables* t, double playback_level) {
  double fc[kFbBands];
  int offset = 0;
  for (int band = 0; band < kFbBands; band++) {
    const int N = kFbLen[band];
    const double f = std::sinh((std::asinh(50. / 650.) +
                                band * (std::asinh(18000. / 650.) - std::asinh(50. / 650.)) / 39.)) *
                     650.;
    const double wt = ear_weight(f);
    fc[band] = f;
    t->fb_len[band] = N;
    t->fb_tap_offset[band] = offset;
    for (int n = 0; n < N / 2 + 1; n++) {
      const double win = 4. / N * std::sin(M_PI * n / N) * std::sin(M_PI * n / N) * wt;
      t->fb_h_re[offset + n] = win * std::cos(2 * M_PI * f * (n - N / 2.) / 48000.);
      t->fb_h_im[offset + n] = win * std::sin(2 * M_PI * f * (n - N / 2.) / 48000.);
    }
    offset += N / 2 + 1;
  }
  t->fb_tap_offset[kFbBands] = offset;
  // delay-indexed, full-length, phase-major copy for the polyphase kernel
  int g = 0;
  for (int band = 0; band < kFbBands; band++) {
    const int N = kFbLen[band];
    const int D = 1 + (kFbLen[0] - N) / 2;          // fbearmodel.c:408
    const double* hre = t->fb_h_re + t->fb_tap_offset[band];
    const double* him = t->fb_h_im + t->fb_tap_offset[band];
    // taps n = 1 .. N-1 (h[0] = h[N] = 0) sit at delays D+1 .. D+N-1.  Band 0
    // (N = 1456, D = 1): the reference reads delay 1456 from fb_buf[offset+1456],
    // which aliases the newest sample, i.e. delay 0 (fbearmodel.c:311-313,413).
    int dlo = D + 1, dhi = D + N - 1;
    if (dhi >= kFbBuf) {
      dlo = 0;
      dhi = kFbBuf - 1;
    }
    t->fb_dlo[band] = dlo;
    t->fb_dhi[band] = dhi;
    t->fb_g_offset[band] = g;
    int local = 0;
    for (int j = 0; j < 32; j++) {
      t->fb_phase_offset[band * 32 + j] = local;
      for (int d = j; d <= dhi; d += 32) {
        if (d < dlo) continue;
        int dd = d;
        if (dd == 0 && D + N - 1 >= kFbBuf) dd = kFbBuf;   // the aliased tap
        const int n = dd - D;
        double re = 0., im = 0.;
        if (n >= 1 && n <= N - 1) {
          if (n <= N / 2) {
            re = hre[n];
            im = him[n];
          } else {
            re = hre[N - n];     // even symmetry (fbearmodel.c:424)
            im = -him[N - n];    // odd symmetry (:425)
          }
        }
        t->fb_g[2 * (g + local)] = re;
        t->fb_g[2 * (g + local) + 1] = im;
        local++;
      }
    }
    g += local;
  }
  // recursive form of the long filters; phases in long double so that what a sample adds
  // when it enters the window is what gets removed N samples later (to ~1e-19)
  for (int band = 0; band < kFbRecBands; band++) {
    const int N = kFbLen[band];
    const long double pi = 3.141592653589793238462643383279502884L;
    const long double w = 2 * pi * (long double)fc[band] / 48000.L;
    const long double d = 2 * pi / N;
    const long double wf[3] = {w, w + d, w - d};
    const long double gain[3] = {2.L, -1.L, -1.L};
    const long double amp = (long double)ear_weight(fc[band]) / N;
    for (int f = 0; f < 3; f++) {
      for (int k = 0; k < 32; k++) {
        // g_f e^{j w_f k} with g_f = gain_f amp e^{-j w N / 2}
        const long double a = wf[f] * k - w * N / 2;
        t->fb_rec_ph[band][k][f].x = (double)(gain[f] * amp * cosl(a));
        t->fb_rec_ph[band][k][f].y = (double)(gain[f] * amp * sinl(a));
        const long double a2 = a + w * N;   // times -e^{j w N}
        t->fb_rec_ph[band][k][3 + f].x = (double)(-gain[f] * amp * cosl(a2));
        t->fb_rec_ph[band][k][3 + f].y = (double)(-gain[f] * amp * sinl(a2));
      }
      for (int i = 0; i < kFbRecGroup; i++) {
        t->fb_rec_rpow[band][f][i].x = (double)cosl(32 * wf[f] * (i + 1));
        t->fb_rec_rpow[band][f][i].y = (double)sinl(32 * wf[f] * (i + 1));
      }
    }
  }
  {
    // band 0, n = N - 1 (delay 1456): same expression as the tap table above
    const double* hre = t->fb_h_re + t->fb_tap_offset[0];
    const double* him = t->fb_h_im + t->fb_tap_offset[0];
    t->fb_rec_alias.x = hre[1];
    t->fb_rec_alias.y = -him[1];
  }
  for (int i = 0; i < 6; i++)
    t->fb_back_mask[i] =
        std::cos(M_PI * (i - 5.) / 12.) * std::cos(M_PI * (i - 5.) / 12.) * 0.9761 / 6.;
  fill_bands(&t->fb, fc, kFbBands, kFbFrame, 1.26539, 0.004, 0.020);  // fbearmodel.c:172-176
  t->level_factor_fb = std::pow(10., playback_level / 20.);
}

void fill_twiddles(DeviceTables* t) {
  for (int k = 0; k < 768; k++) {
    t->tw1024[k].x = std::cos(-2. * M_PI * k / 1024.);
    t->tw1024[k].y = std::sin(-2. * M_PI * k / 1024.);
  }
  for (int k = 0; k < kFftBins; k++) {
    t->tw2048[k].x = std::cos(-2. * M_PI * k / 2048.);
    t->tw2048[k].y = std::sin(-2. * M_PI * k / 2048.);
  }
}

void fill_nn(DeviceTables* t) {
  // nn.c:40-93
  static const double amin_b[11] = {393.916656, 361.965332, -24.045116, 1.110661, -0.206623, 0.074318,
                                    1.113683,   0.950345,   0.029985,   0.000101, 0.};
  static const double amax_b[11] = {921,       881.131226,  16.212030, 107.137772, 2.886017, 13.933351,
                                    63.257874, 1145.018555, 14.819740, 1.,         1.};
  static const double wx_b[11][3] = {
      {-0.502657, 0.436333, 1.219602},  {4.307481, 3.246017, 1.123743},  {4.984241, -2.211189, -0.192096},
      {0.051056, -1.762424, 4.331315},  {2.321580, 1.789971, -0.754560}, {-5.303901, -3.452257, -10.814982},
      {2.730991, -6.111805, 1.519223},  {0.624950, -1.331523, -5.955151}, {3.102889, 0.871260, -5.922878},
      {-1.051468, -0.939882, -0.142913}, {-1.804679, -0.503610, -0.620456}};
  static const double wxb_b[3] = {-2.518254, 0.654841, -2.207228};
  static const double wy_b[3] = {-3.817048, 4.107138, 4.629582};
  static const double amin_a[5] = {13.298751, 0.041073, -25.018791, 0.061560, 0.02452};
  static const double amax_a[5] = {2166.5, 13.24326, 13.46708, 10.226771, 14.224874};
  static const double wx_a[5][5] = {{21.211773, -39.013052, -1.382553, -14.545348, -0.320899},
                                    {-8.981803, 19.956049, 0.935389, -1.686586, -3.238586},
                                    {1.633830, -2.877505, -7.442935, 5.606502, -1.783120},
                                    {6.103821, 19.587435, -0.240284, 1.088213, -0.511314},
                                    {11.556344, 3.892028, 9.720441, -3.287205, -11.031250}};
  static const double wxb_a[5] = {1.330890, 2.686103, 2.096598, -1.327851, 3.087055};
  static const double wy_a[5] = {-4.696996, -3.289959, 7.004782, 6.651897, 4.009144};
  std::memset(t->nn_wx, 0, sizeof t->nn_wx);
  if (t->advanced) {
    t->nn_inputs = 5;
    t->nn_hidden = 5;
    for (int i = 0; i < 5; i++) {
      t->nn_amin[i] = amin_a[i];
      t->nn_amax[i] = amax_a[i];
      t->nn_wxb[i] = wxb_a[i];
      t->nn_wy[i] = wy_a[i];
      for (int j = 0; j < 5; j++) t->nn_wx[i * 5 + j] = wx_a[i][j];
    }
    t->nn_wyb = -1.360308;
  } else {
    t->nn_inputs = 11;
    t->nn_hidden = 3;
    for (int i = 0; i < 11; i++) {
      t->nn_amin[i] = amin_b[i];
      t->nn_amax[i] = amax_b[i];
      for (int j = 0; j < 3; j++) t->nn_wx[i * 5 + j] = wx_b[i][j];
    }
    for (int j = 0; j < 3; j++) {
      t->nn_wxb[j] = wxb_b[j];
      t->nn_wy[j] = wy_b[j];
    }
    t->nn_wyb = -0.307594;
  }
}

}  // namespace

void build_tables(DeviceTables* t, bool advanced, double playback_level) {
  std::memset(t, 0, sizeof *t);
  t->advanced = advanced ? 1 : 0;
  t->fft_bands = advanced ? 55 : 109;
  fill_fft_model(t, playback_level);
  fill_fb_model(t, playback_level);
  fill_twiddles(t);
  fill_nn(t);
}

}  // namespace peaq
