/* harness.c -- plays "pipeline" for the element shell against the GStreamer stand-in
 * (TEST INFRASTRUCTURE ONLY).  usage: harness REF.f32 TEST.f32 CHANNELS ADVANCED
 * Mirrors what the reference's CLI pipeline does to the element (peaq.c:154-220): link two
 * sources, negotiate caps, stream buffers on both pads, EOS, PAUSED->READY, read odg / di.
 * Prints "CHECK <name> ok|FAIL" lines and the properties; exit status 0 if every check passed. */
#include "gst_stub.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static int failures = 0;
#define CHECK(name, cond) do { int ok__ = (cond); printf ("CHECK %s %s\n", name, ok__ ? "ok" : "FAIL"); if (!ok__) failures++; } while (0)

static float *
read_f32 (const char *path, size_t *n)
{
  FILE *f = fopen (path, "rb");
  if (!f) { perror (path); exit (2); }
  fseek (f, 0, SEEK_END);
  long sz = ftell (f);
  fseek (f, 0, SEEK_SET);
  float *p = malloc (sz ? sz : 4);
  if (fread (p, 1, sz, f) != (size_t) sz) exit (2);
  fclose (f);
  *n = sz / sizeof (float);
  return p;
}

int
main (int argc, char **argv)
{
  if (argc < 5) { fprintf (stderr, "usage: %s REF.f32 TEST.f32 CHANNELS ADVANCED\n", argv[0]); return 2; }
  size_t n_ref, n_test;
  float *ref = read_f32 (argv[1], &n_ref), *test = read_f32 (argv[2], &n_test);
  const int channels = atoi (argv[3]), advanced = atoi (argv[4]);

  CHECK ("plugin_init", gst_stub_plugin_init_peaqb200 (NULL));
  GstElement *peaq = gst_stub_factory_make ("peaq");          /* same element name as the reference */
  CHECK ("factory_make", peaq != NULL);
  const GstElementClass *klass = gst_stub_element_class (peaq);
  CHECK ("klass_sink_audio", klass->klass && !strcmp (klass->klass, "Sink/Audio"));
  CHECK ("sink_flag", (GST_OBJECT (peaq)->flags & GST_ELEMENT_FLAG_SINK) != 0);
  GstPad *refpad = gst_stub_get_pad (peaq, "ref"), *testpad = gst_stub_get_pad (peaq, "test");
  CHECK ("pads", refpad && testpad && klass->n_templates == 2);

  /* construct-time defaults (gstpeaq.c:273-317) */
  double level = 0; int adv = -1, console = -1;
  g_object_get (peaq, "playback_level", &level, "advanced", &adv, "console-output", &console, NULL);
  CHECK ("defaults", level == 92. && adv == 0 && console == 1);
  g_object_set (peaq, "advanced", advanced, NULL);

  /* caps negotiation: a pad offers what the peer of the OTHER pad can do (gstpeaq.c:215-244) */
  gst_stub_pad_set_peer_caps (refpad, channels);               /* ref source: fixed channel count */
  gst_stub_pad_set_peer_caps (testpad, 0);                     /* test source: any */
  CHECK ("query_test_pad_follows_ref_peer", gst_stub_pad_query_caps (testpad, 0) == channels);
  CHECK ("query_ref_pad_unrestricted", gst_stub_pad_query_caps (refpad, 0) == 0);
  CHECK ("query_filter_conflict_is_empty", gst_stub_pad_query_caps (testpad, channels + 1) == -1);
  CHECK ("caps_event_rejected", !gst_stub_pad_send_caps (testpad, channels + 1));
  CHECK ("caps_event_ref", gst_stub_pad_send_caps (refpad, channels));
  CHECK ("caps_event_test", gst_stub_pad_send_caps (testpad, channels));

  /* stream: unequal buffer sizes on the two pads, like two independent decoders */
  size_t pr = 0, pt = 0;
  const size_t br = 4096 * (size_t) channels, bt = 1000 * (size_t) channels;
  int flow_ok = 1;
  while (pr < n_ref || pt < n_test) {
    if (pr < n_ref) { size_t k = n_ref - pr < br ? n_ref - pr : br; flow_ok &= gst_stub_pad_push (refpad, ref + pr, k) == GST_FLOW_OK; pr += k; }
    for (int i = 0; i < 4 && pt < n_test; i++) { size_t k = n_test - pt < bt ? n_test - pt : bt; flow_ok &= gst_stub_pad_push (testpad, test + pt, k) == GST_FLOW_OK; pt += k; }
  }
  CHECK ("chain_flow_ok", flow_ok);

  /* mid-stream property read does not disturb the stream */
  double odg_mid = 0;
  g_object_get (peaq, "odg", &odg_mid, NULL);

  /* EOS aggregation (gstpeaq.c:668-688): the message goes out once BOTH pads are at EOS */
  gst_stub_pad_send_eos (refpad, 7);
  CHECK ("eos_waits_for_both", peaq->messages_eos == 0);
  gst_stub_pad_send_eos (testpad, 8);
  CHECK ("eos_posted_once", peaq->messages_eos == 1 && peaq->last_eos_seqnum == 8);

  /* PAUSED->READY: flush + evaluate + console output (gstpeaq.c:764-778) */
  printf ("--- console output ---\n");
  fflush (stdout);
  CHECK ("change_state", gst_stub_change_state (peaq, GST_STATE_CHANGE_PAUSED_TO_READY) == GST_STATE_CHANGE_SUCCESS);
  printf ("--- end ---\n");
  double odg = 0, di = 0, snr = 0;
  g_object_get (peaq, "odg", &odg, "di", &di, "totalsnr", &snr, NULL);
  printf ("RESULT odg %.17g di %.17g totalsnr %.17g odg_mid %.17g\n", odg, di, snr, odg_mid);
  CHECK ("no_element_errors", gst_stub_error_count () == 0);
  CHECK ("lock_balanced", GST_OBJECT (peaq)->lock_depth == 0);
  g_object_unref (peaq);
  printf ("%s\n", failures ? "HARNESS FAILED" : "HARNESS OK");
  return failures ? 1 : 0;
}
