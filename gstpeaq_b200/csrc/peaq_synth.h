// Integer-only synthetic (ref, test) signal generator.
//
// Bit-identical on host and device: no libm, no floating point until the final
// exact division by 32768.  A pair is addressed by its global index, a sample
// by (n, channel), so any part of any pair can be produced independently (the
// GPU generates the benchmark batch in place; tests generate the same pairs on
// the host for the CPU oracle).
//
//   ref  : sum of <= 24 partials of a fundamental f0 in [110, 880) Hz, spread
//          up to 16 kHz, gentle amplitude roll-off, slow triangular AM,
//          quantised to 16 bit;
//   test : the same partials below a per-pair cut-off (10..15 kHz) -- a crude
//          codec low-pass -- re-quantised to 8 + (pair mod 7) bits with
//          hash-noise dither, so pairs cover ODGs from ~0 down to ~-3.9.
// Both stay well above the bandwidth MOV's noise-floor gate so every pair has
// a finite ODG (a reference without content above 8.1 kHz makes BandwidthRef
// 0/0, movs.c:797 / movaccum.c:451).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define PEAQ_SYNTH_FN __host__ __device__ __forceinline__
#else
#define PEAQ_SYNTH_FN static inline
#endif

#define PEAQ_SYNTH_MAX_PARTIALS 24
#define PEAQ_SYNTH_TABLE_BITS 12
#define PEAQ_SYNTH_TABLE_SIZE (1 << PEAQ_SYNTH_TABLE_BITS)

PEAQ_SYNTH_FN uint64_t peaq_synth_mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// One period of an integer "sine" (Bhaskara I rational approximation, exact
// integer arithmetic): index in [0, 4096), value in [-32767, 32767].
PEAQ_SYNTH_FN int32_t peaq_synth_sine_entry(uint32_t idx) {
  const int64_t half = PEAQ_SYNTH_TABLE_SIZE / 2;
  int64_t i = idx & (PEAQ_SYNTH_TABLE_SIZE / 2 - 1);
  int64_t p = i * (half - i);
  int64_t v = (16 * p * 32767) / (5 * half * half - 4 * p);
  return (idx & (PEAQ_SYNTH_TABLE_SIZE / 2)) ? (int32_t)-v : (int32_t)v;
}

typedef struct {
  int n_partials;
  int n_test_partials;            // partials kept in the test signal
  int test_shift;                 // 16 - bits of the test quantiser
  uint32_t lfo_mask;              // AM period - 1 (power of two)
  uint32_t lfo_phase;
  uint32_t inc[PEAQ_SYNTH_MAX_PARTIALS];   // 32-bit phase increment per sample
  int32_t amp[PEAQ_SYNTH_MAX_PARTIALS];
  uint32_t phase[2][PEAQ_SYNTH_MAX_PARTIALS];  // start phase per channel
  uint64_t noise_key;
} PeaqSynthPair;

PEAQ_SYNTH_FN void peaq_synth_pair_init(PeaqSynthPair* s, uint64_t pair_index) {
  uint64_t r = peaq_synth_mix64(pair_index * 0xD1342543DE82EF95ull + 1);
  const uint32_t f0_mhz = 110000u + (uint32_t)(r % 770000u);     // milli-Hz
  r = peaq_synth_mix64(r);
  const uint32_t max_h = 16000000u / f0_mhz;                      // harmonics <= 16 kHz
  const uint32_t stride = (max_h + PEAQ_SYNTH_MAX_PARTIALS - 1) / PEAQ_SYNTH_MAX_PARTIALS;
  const uint32_t cutoff_mhz = 10000000u + (uint32_t)(r % 5000000u);
  r = peaq_synth_mix64(r);
  int n = 0, nt = 0;
  for (uint32_t j = 0; j < PEAQ_SYNTH_MAX_PARTIALS; j++) {
    const uint32_t h = 1 + j * stride;
    if (h > max_h) break;
    const uint64_t f = (uint64_t)f0_mhz * h;                      // milli-Hz
    s->inc[n] = (uint32_t)((f << 32) / 48000000ull);
    s->amp[n] = (int32_t)(9600 / (2 + j));                        // 4800 / (1 + j/2)
    r = peaq_synth_mix64(r);
    s->phase[0][n] = (uint32_t)r;
    s->phase[1][n] = (uint32_t)(r >> 32);
    if (f < cutoff_mhz) nt = n + 1;
    n++;
  }
  s->n_partials = n;
  s->n_test_partials = nt;
  s->test_shift = 16 - (8 + (int)(pair_index % 7));
  r = peaq_synth_mix64(r);
  s->lfo_mask = (r & 1) ? 0xFFFFu : 0x7FFFu;
  s->lfo_phase = (uint32_t)(r >> 8);
  s->noise_key = peaq_synth_mix64(r);
}

// sine_table: PEAQ_SYNTH_TABLE_SIZE int16 entries = peaq_synth_sine_entry(i).
// Writes the 16-bit sample values of ref and test for (n, channel).
PEAQ_SYNTH_FN void peaq_synth_sample(const PeaqSynthPair* s, const int16_t* sine_table,
                                     uint64_t n, int channel, int32_t* ref_out,
                                     int32_t* test_out) {
  const uint32_t n32 = (uint32_t)n;
  int64_t acc_ref = 0, acc_test = 0;
  for (int j = 0; j < s->n_partials; j++) {
    const uint32_t ph = s->phase[channel][j] + n32 * s->inc[j];
    const int32_t v = (int32_t)sine_table[ph >> (32 - PEAQ_SYNTH_TABLE_BITS)] * s->amp[j];
    acc_ref += v;
    if (j < s->n_test_partials) acc_test += v;
  }
  // slow triangular amplitude modulation, gain in [160, 256] / 256
  const uint32_t t = (n32 + s->lfo_phase) & s->lfo_mask;
  const uint32_t half = (s->lfo_mask + 1) >> 1;
  const uint32_t tri = t < half ? t : (s->lfo_mask + 1 - t);      // 0 .. half
  const int64_t gain = 160 + (int64_t)((96ull * tri) / half);
  // >> 15 (table scale) and >> 8 (gain scale), rounding toward -inf
  int64_t r16 = (acc_ref * gain) >> 23;
  int64_t t16 = (acc_test * gain) >> 23;
  // test: re-quantise with uniform hash dither
  const int sh = s->test_shift;
  const uint64_t hsh = peaq_synth_mix64(s->noise_key ^ (n * 2 + (uint64_t)channel));
  const int64_t dither = (int64_t)(hsh & ((1u << sh) - 1));
  t16 = ((t16 + dither + 32768) >> sh << sh) - 32768;
  if (r16 > 32767) r16 = 32767;
  if (r16 < -32768) r16 = -32768;
  if (t16 > 32767) t16 = 32767;
  if (t16 < -32768) t16 = -32768;
  *ref_out = (int32_t)r16;
  *test_out = (int32_t)t16;
}
