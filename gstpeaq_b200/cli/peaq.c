/* peaq -- command line front end of the B200-native PEAQ engine.
 *
 * Drop-in for the reference CLI (/root/reference/src/peaq.c): same options
 * (--basic, --advanced, --version, REFFILE TESTFILE; peaq.c:35-45), same two
 * output lines (peaq.c:217-220), same exit codes (0 ok, 1 usage / input
 * error, 2 engine unavailable -- the reference's "element missing",
 * peaq.c:146-150).  The reference lets GStreamer decode, convert and resample
 * (filesrc ! wavparse ! audioconvert ! audioresample, peaq.c:154-209); here a
 * small WAV reader feeds the session API of libpeaq_b200 directly:
 *   - RIFF/WAVE PCM 8/16/24/32 bit, IEEE float 32/64, WAVE_FORMAT_EXTENSIBLE;
 *   - integer samples are scaled by 1/2^(bits-1) like audioconvert (the
 *     thresholds of the algorithm assume 16-bit full scale = 1.0,
 *     gstpeaq.c:1093, fftearmodel.c:511);
 *   - the sample rate must be 48 kHz (no resampler; the element's caps are
 *     rate=48000, gstpeaq.c:146-152);
 *   - a mono file paired with a stereo file is up-mixed by duplication, as the
 *     element's caps negotiation + audioconvert do (runtest-1.0.sh:31-48).
 */
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/peaq_b200.h"

#define PEAQ_CLI_VERSION "peaq-b200 1"

typedef struct {
  float *data;          /* interleaved */
  size_t frames;        /* samples per channel */
  int channels;
  int rate;
} Wav;

static uint32_t rd32(const unsigned char *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint16_t rd16(const unsigned char *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

static int wav_read(const char *path, Wav *w, char *err, size_t errlen) {
  FILE *f = fopen(path, "rb");
  unsigned char hdr[12], ck[8];
  int have_fmt = 0, fmt_tag = 0, bits = 0, block = 0;
  memset(w, 0, sizeof *w);
  if (!f) {
    snprintf(err, errlen, "%s: %s", path, strerror(errno));
    return -1;
  }
  if (fread(hdr, 1, 12, f) != 12 || memcmp(hdr, "RIFF", 4) || memcmp(hdr + 8, "WAVE", 4)) {
    snprintf(err, errlen, "%s: not a RIFF/WAVE file", path);
    fclose(f);
    return -1;
  }
  while (fread(ck, 1, 8, f) == 8) {
    uint32_t size = rd32(ck + 4);
    if (!memcmp(ck, "fmt ", 4)) {
      unsigned char fmt[40];
      uint32_t n = size < sizeof fmt ? size : (uint32_t)sizeof fmt;
      if (size < 16 || fread(fmt, 1, n, f) != n) break;
      fmt_tag = rd16(fmt);
      w->channels = rd16(fmt + 2);
      w->rate = (int)rd32(fmt + 4);
      block = rd16(fmt + 12);
      bits = rd16(fmt + 14);
      if (fmt_tag == 0xFFFE && n >= 26) fmt_tag = rd16(fmt + 24);   /* EXTENSIBLE: sub-format GUID */
      have_fmt = 1;
      if (size > n) fseek(f, (long)(size - n), SEEK_CUR);
      if (size & 1) fseek(f, 1, SEEK_CUR);
    } else if (!memcmp(ck, "data", 4)) {
      size_t bytes, i, n;
      unsigned char *raw;
      int bps;
      if (!have_fmt || w->channels < 1 || bits < 8 || block < 1) break;
      bps = bits / 8;
      if (!((fmt_tag == 1 && (bits == 8 || bits == 16 || bits == 24 || bits == 32)) ||
            (fmt_tag == 3 && (bits == 32 || bits == 64)))) {
        snprintf(err, errlen, "%s: unsupported WAV encoding (format %d, %d bit)", path, fmt_tag, bits);
        fclose(f);
        return -1;
      }
      raw = (unsigned char *)malloc(size ? size : 1);
      bytes = raw ? fread(raw, 1, size, f) : 0;   /* tolerate truncated files / streaming sizes */
      n = bytes / (size_t)bps;
      w->frames = n / (size_t)w->channels;
      n = w->frames * (size_t)w->channels;
      w->data = (float *)malloc((n ? n : 1) * sizeof(float));
      if (!raw || !w->data) {
        snprintf(err, errlen, "%s: out of memory", path);
        free(raw);
        fclose(f);
        return -1;
      }
      for (i = 0; i < n; i++) {
        const unsigned char *p = raw + i * (size_t)bps;
        if (fmt_tag == 3) {
          if (bits == 32) {
            float v;
            memcpy(&v, p, 4);
            w->data[i] = v;
          } else {
            double v;
            memcpy(&v, p, 8);
            w->data[i] = (float)v;
          }
        } else if (bits == 8) {
          w->data[i] = (float)((int)p[0] - 128) / 128.0f;
        } else if (bits == 16) {
          w->data[i] = (float)(int16_t)rd16(p) / 32768.0f;
        } else if (bits == 24) {
          int32_t v = (int32_t)((uint32_t)p[0] << 8 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 24) >> 8;
          w->data[i] = (float)v / 8388608.0f;
        } else {
          w->data[i] = (float)((double)(int32_t)rd32(p) / 2147483648.0);
        }
      }
      free(raw);
      fclose(f);
      return 0;
    } else {
      fseek(f, (long)(size + (size & 1)), SEEK_CUR);
    }
  }
  snprintf(err, errlen, "%s: malformed WAV file", path);
  fclose(f);
  return -1;
}

/* mono -> n channels by duplication (what audioconvert does for caps channels=2) */
static int wav_upmix(Wav *w, int channels) {
  size_t i;
  int c;
  float *d;
  if (w->channels == channels) return 0;
  if (w->channels != 1) return -1;
  d = (float *)malloc((w->frames ? w->frames : 1) * (size_t)channels * sizeof(float));
  if (!d) return -1;
  for (i = 0; i < w->frames; i++)
    for (c = 0; c < channels; c++) d[i * (size_t)channels + (size_t)c] = w->data[i];
  free(w->data);
  w->data = d;
  w->channels = channels;
  return 0;
}

static void usage(const char *argv0) {
  printf("Usage:\n  %s [OPTION...] REFFILE TESTFILE\n\n"
         "peaq computes the Objective Difference Grade based on ITU-R BS.1387-1 (but it\n"
         "does not meet its conformance requirements).\n\n"
         "Options:\n"
         "  --version     print version information\n"
         "  --advanced    use advanced version\n"
         "  --basic       use basic version (default)\n"
         "  --device=N    CUDA device to run on (default 0)\n",
         argv0);
}

int main(int argc, char *argv[]) {
  int advanced = 0, device = 0, i, nfiles = 0, rc, channels;
  const char *files[2] = {NULL, NULL};
  char err[512];
  Wav ref, test;
  peaq_b200_session *s = NULL;
  peaq_b200_result res;
  size_t pos;

  for (i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--advanced")) advanced = 1;
    else if (!strcmp(argv[i], "--basic")) advanced = 0;
    else if (!strcmp(argv[i], "--version")) {
      puts(PEAQ_CLI_VERSION "\n"
           "B200-native PEAQ engine with the command line of GstPEAQ's `peaq`.\n"
           "There is NO WARRANTY, to the extent permitted by law.");
      return 0;
    } else if (!strncmp(argv[i], "--device=", 9)) device = atoi(argv[i] + 9);
    else if (!strcmp(argv[i], "--help") || !strcmp(argv[i], "-h")) {
      usage(argv[0]);
      return 0;
    } else if (argv[i][0] == '-' && argv[i][1] == '-') {
      printf("Failed to initialize: Unknown option %s\n", argv[i]);
      return 1;
    } else if (nfiles < 2) files[nfiles++] = argv[i];
    else nfiles++;
  }
  if (nfiles != 2) {
    usage(argv[0]);
    return 1;
  }
  if (wav_read(files[0], &ref, err, sizeof err) || wav_read(files[1], &test, err, sizeof err)) {
    printf("Error: %s\n", err);
    return 1;
  }
  if (ref.rate != 48000 || test.rate != 48000) {
    printf("Error: both files must be sampled at 48000 Hz (got %d and %d); resample first\n", ref.rate, test.rate);
    return 1;
  }
  channels = ref.channels > test.channels ? ref.channels : test.channels;
  if (channels > 2 || wav_upmix(&ref, channels) || wav_upmix(&test, channels)) {
    printf("Error: unsupported channel layout (%d and %d channels)\n", ref.channels, test.channels);
    return 1;
  }

  if (peaq_b200_session_create(&s, device) != 0) {
    puts("Failed to instantiate peaq element: " );
    puts(peaq_b200_last_error());
    return 2;
  }
  rc = peaq_b200_session_set_advanced(s, advanced);
  if (!rc) rc = peaq_b200_session_set_channels(s, channels);
  /* feed both pads in buffers of 1 s, alternating, like two live sources would */
  for (pos = 0; !rc && (pos < ref.frames || pos < test.frames); pos += 48000) {
    if (pos < ref.frames) {
      size_t n = ref.frames - pos < 48000 ? ref.frames - pos : 48000;
      rc = peaq_b200_session_push(s, PEAQ_B200_PAD_REF, ref.data + pos * (size_t)channels, n);
    }
    if (!rc && pos < test.frames) {
      size_t n = test.frames - pos < 48000 ? test.frames - pos : 48000;
      rc = peaq_b200_session_push(s, PEAQ_B200_PAD_TEST, test.data + pos * (size_t)channels, n);
    }
  }
  if (!rc) rc = peaq_b200_session_finish(s);
  if (!rc) rc = peaq_b200_session_get_result(s, &res);
  if (rc) {
    printf("Error: %s\n", peaq_b200_last_error());
    peaq_b200_session_destroy(s);
    return 2;
  }
  printf("Objective Difference Grade: %.3f\n", res.odg);
  printf("Distortion Index: %.3f\n", res.di);
  peaq_b200_session_destroy(s);
  free(ref.data);
  free(test.data);
  return 0;
}
