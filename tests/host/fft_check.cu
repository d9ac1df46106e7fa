// Host replay of the frame kernel's 1024-point transform (gstpeaq_b200/csrc/peaq_fft.cuh): the
// functions are __host__ __device__, a pass touches only the calling thread's own sixteen points,
// so running "thread" t = 0..63 one after the other between the barriers IS the kernel's schedule.
// Checks, for random input behind the window phase's level-1 scatter:
//   1. two trips (levels 4+16, 64+256) == four single-level passes (fft_pass4), bit for bit;
//   2. both == a long-double DFT to 1e-12 of the largest bin;
//   3. the mirror exchange hands thread t exactly Z[(1024 - k) & 1023];
//   4. shared-memory wavefronts of every 128-bit access of the new path under the quarter-warp
//      rule (8 lanes, eight 16-byte bank groups): all at the minimum.
// Prints "ok" and exits 0, or says what failed.  Built and run by tests/test_host.py (nvcc, no GPU).
#include "../../gstpeaq_b200/csrc/peaq_fft.cuh"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <set>
#include <vector>

using namespace peaq;

struct NoSync {
  __host__ __device__ void operator()() const {}
};

static int wavefronts_of(const std::vector<int>& idx) {   // idx[lane] = 16-byte element index
  int total = 0;
  for (int q = 0; q < 4; q++) {
    std::set<int> addr;
    for (int l = 8 * q; l < 8 * q + 8; l++) addr.insert(idx[l]);
    int per_group[8] = {0};
    for (int a : addr) per_group[a & 7]++;
    int w = 0;
    for (int g = 0; g < 8; g++) w = per_group[g] > w ? per_group[g] : w;
    total += w;
  }
  return total;
}

int main() {
  static double2 tw[512], z_new[1024], z_old[1024], in[1024];
  for (int q = 0; q < 512; q++) tw[fft_twi(q)] = make_double2(std::cos(-2. * M_PI * q / 1024.), std::sin(-2. * M_PI * q / 1024.));
  {
    std::set<int> seen;
    for (int q = 0; q < 512; q++) seen.insert(fft_twi(q));
    if (seen.size() != 512 || *seen.rbegin() != 511) { std::printf("fft_twi is not a permutation of 0..511\n"); return 1; }
  }
  unsigned long long s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992. - 0.5; };
  for (int n = 0; n < 1024; n++) in[n] = make_double2(rnd(), rnd());

  // window phase: level 1 in registers, scatter (peaq_frames.cuh, phase 2)
  for (int t = 0; t < 64; t++) {
    const int slot_t = fft_slot<10>(t);
    for (int j = 0; j < 4; j++) {
      double2 a[4];
      for (int m = 0; m < 4; m++) a[m] = in[t + 64 * (j + 4 * m)];
      const double2 t0 = make_double2(a[0].x + a[2].x, a[0].y + a[2].y);
      const double2 t1 = make_double2(a[0].x - a[2].x, a[0].y - a[2].y);
      const double2 t2 = make_double2(a[1].x + a[3].x, a[1].y + a[3].y);
      const double2 t3 = make_double2(a[1].y - a[3].y, a[3].x - a[1].x);
      z_new[slot_t ^ fft_slot<10>(64 * j)] = make_double2(t0.x + t2.x, t0.y + t2.y);
      z_new[slot_t ^ fft_slot<10>(64 * (j + 4))] = make_double2(t1.x + t3.x, t1.y + t3.y);
      z_new[slot_t ^ fft_slot<10>(64 * (j + 8))] = make_double2(t0.x - t2.x, t0.y - t2.y);
      z_new[slot_t ^ fft_slot<10>(64 * (j + 12))] = make_double2(t1.x - t3.x, t1.y - t3.y);
    }
  }
  std::memcpy(z_old, z_new, sizeof(z_new));

  // four single-level passes
  for (int t = 0; t < 64; t++) fft_pass4<10, 4, 64, NoSync>(z_old, tw, t, NoSync());
  for (int t = 0; t < 64; t++) fft_pass4<10, 16, 64, NoSync>(z_old, tw, t, NoSync());
  for (int t = 0; t < 64; t++) fft_pass4<10, 64, 64, NoSync>(z_old, tw, t, NoSync());
  for (int t = 0; t < 64; t++) fft_pass4<10, 256, 64, NoSync>(z_old, tw, t, NoSync());

  // two trips
  static double2 e[64][16];
  for (int t = 0; t < 64; t++) fft1024_levels_4_16(z_new, tw, t);
  for (int t = 0; t < 64; t++) fft1024_load_64_256(e[t], z_new, t);
  for (int t = 0; t < 64; t++) fft1024_levels_64_256(e[t], z_new, tw, t);

  int bad = 0;
  for (int t = 0; t < 64; t++)
    for (int u = 0; u < 16; u++) {
      const double2 want = z_old[fft_swz(t + 64 * u)];
      if (std::memcmp(&want, &e[t][u], sizeof(double2)) != 0 && bad++ < 5)
        std::printf("Z[%d]: two trips %.17g %.17g, four passes %.17g %.17g\n", t + 64 * u, e[t][u].x, e[t][u].y, want.x, want.y);
    }
  for (int t = 0; t < 64; t++)
    for (int u = 0; u < 8; u++) {
      const int k = t + 64 * u;
      const double2 want = z_old[fft_swz((1024 - k) & 1023)];
      const double2 got = fft1024_mirror(z_new, t, u, e[t][u]);
      if (std::memcmp(&want, &got, sizeof(double2)) != 0 && bad++ < 10) std::printf("mirror of bin %d wrong\n", k);
    }
  {
    const double2 want = z_old[fft_swz(512)], got = z_new[fft1024_xch_mid()];
    if (std::memcmp(&want, &got, sizeof(double2)) != 0) { std::printf("Z[512] not where thread 0 looks for it\n"); bad++; }
  }

  // against a DFT in long double
  long double worst = 0, scale = 0;
  for (int k = 0; k < 1024; k += 37) {
    long double re = 0, im = 0;
    for (int n = 0; n < 1024; n++) {
      const long double ph = -2.0L * 3.141592653589793238462643383279502884L * (long double)((k * n) & 1023) / 1024.0L;
      re += in[n].x * cosl(ph) - in[n].y * sinl(ph);
      im += in[n].x * sinl(ph) + in[n].y * cosl(ph);
    }
    const double2 got = e[k & 63][k >> 6];
    worst = fmaxl(worst, fmaxl(fabsl(re - got.x), fabsl(im - got.y)));
    scale = fmaxl(scale, fmaxl(fabsl(re), fabsl(im)));
  }
  if (worst > 1e-12L * scale) { std::printf("DFT mismatch %Lg of %Lg\n", worst, scale); bad++; }

  // EHS transforms (peaq_frames.cuh, ehs_channel_pair): first radix-4 level in registers while
  // filling == fill, then the LQ = 1 pass, for the 512- and the 256-point transform
  {
    static double2 a_fill[512], a_reg[512], b_fill[256], b_reg[256];
    static double d[512];
    for (int i = 0; i < 512; i++) d[i] = rnd();
    for (int t = 0; t < 64; t++) {
      const int sl9 = fft_slot<9>(t), sl8 = fft_slot<8>(t);
      for (int u = 0; u < 8; u++) a_fill[sl9 ^ fft_slot<9>(64 * u)] = make_double2(d[t + 64 * u], u < 4 ? d[t + 64 * u] : 0.);
      for (int u = 0; u < 4; u++) b_fill[sl8 ^ fft_slot<8>(64 * u)] = in[t + 64 * u];
      for (int j = 0; j < 2; j++) {
        const double v0 = d[t + 64 * j], v1 = d[t + 64 * (j + 2)];
        double2 a0 = make_double2(v0, v0), a1 = make_double2(v1, v1);
        double2 a2 = make_double2(d[t + 64 * (j + 4)], 0.), a3 = make_double2(d[t + 64 * (j + 6)], 0.);
        fft_bfly4_nt(a0, a1, a2, a3);
        a_reg[sl9 ^ fft_slot<9>(64 * j)] = a0;
        a_reg[sl9 ^ fft_slot<9>(64 * (j + 2))] = a1;
        a_reg[sl9 ^ fft_slot<9>(64 * (j + 4))] = a2;
        a_reg[sl9 ^ fft_slot<9>(64 * (j + 6))] = a3;
      }
      double2 y[4];
      for (int u = 0; u < 4; u++) y[u] = in[t + 64 * u];
      fft_bfly4_nt(y[0], y[1], y[2], y[3]);
      for (int u = 0; u < 4; u++) b_reg[sl8 ^ fft_slot<8>(64 * u)] = y[u];
    }
    for (int t = 0; t < 64; t++) fft_pass4<9, 1, 64, NoSync>(a_fill, tw, t, NoSync());
    for (int t = 0; t < 64; t++) fft_pass4<8, 1, 64, NoSync>(b_fill, tw, t, NoSync());
    if (std::memcmp(a_fill, a_reg, sizeof(a_fill)) != 0) { std::printf("512-point first level differs\n"); bad++; }
    if (std::memcmp(b_fill, b_reg, sizeof(b_fill)) != 0) { std::printf("256-point first level differs\n"); bad++; }
  }

  // shared-memory wavefronts of the new path's 128-bit accesses (per warp instruction)
  int excess = 0;
  auto check = [&](const char* what, const std::vector<int>& idx, int ideal) {
    const int w = wavefronts_of(idx);
    if (w > ideal) { if (excess++ < 10) std::printf("%s: %d wavefronts, %d possible\n", what, w, ideal); }
  };
  for (int warp = 0; warp < 2; warp++) {
    std::vector<int> idx(32);
    for (int j = 0; j < 16; j++) {
      for (int l = 0; l < 32; l++) idx[l] = fft_swz(fft_group_a(32 * warp + l)) ^ fft_swz(4 * j);
      check("levels 4+16 data", idx, 4);
      for (int l = 0; l < 32; l++) idx[l] = fft_swz(32 * warp + l) ^ fft_swz(64 * j);
      check("levels 64+256 load", idx, 4);
    }
    for (int r = 0; r < 8; r++) {
      for (int l = 0; l < 32; l++) idx[l] = r * 64 + ((32 * warp + l - 1) & 63);
      check("exchange store", idx, 4);
      for (int l = 0; l < 32; l++) { const int t = 32 * warp + l; idx[l] = (r + (t == 0 ? 1 : 0)) * 64 + (63 - t); }
      check("mirror load", idx, 4);
    }
    // twiddles: level 4 (stride 64, k = t & 3), level 16 (16 k + 64 a), level 64 (4 t), level 256 (t + 64 a)
    for (int l = 0; l < 32; l++) idx[l] = fft_twi(((32 * warp + l) & 3) * 64);
    check("twiddle level 4", idx, 4);
    for (int a = 0; a < 4; a++) {
      for (int l = 0; l < 32; l++) idx[l] = fft_twi(((32 * warp + l) & 3) * 16) ^ fft_twi(a * 64);
      check("twiddle level 16", idx, 4);
      for (int l = 0; l < 32; l++) idx[l] = fft_twi(32 * warp + l) ^ fft_twi(a * 64);
      check("twiddle level 256", idx, 4);
    }
    for (int l = 0; l < 32; l++) idx[l] = fft_twi((32 * warp + l) * 4);
    check("twiddle level 64", idx, 4);
    // the EHS's strides: 2 (tw[2 k], radix-2 pass of the 512-point transform) and 8 (128-point)
    for (int l = 0; l < 32; l++) idx[l] = fft_twi((32 * warp + l) * 2);
    check("twiddle stride 2", idx, 4);
    for (int l = 0; l < 32; l++) idx[l] = fft_twi(l * 8);
    check("twiddle stride 8", idx, 4);
    for (int l = 0; l < 32; l++) idx[l] = fft_twi((l & 15) * 16);
    check("twiddle stride 16, 16 values", idx, 4);
  }
  if (bad || excess) { std::printf("FAILED: %d value mismatches, %d accesses above the minimum\n", bad, excess); return 1; }
  std::printf("ok\n");
  return 0;
}
