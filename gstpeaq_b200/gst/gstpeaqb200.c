/* gstpeaqb200.c -- GStreamer-1.0 element `peaq` on top of libpeaq_b200.
 *
 * The element shell of the reference (/root/reference/src/gstpeaq.c) with its
 * per-frame body replaced by the session API of include/peaq_b200.h.  Surface
 * kept: element name `peaq`, klass Sink/Audio, ALWAYS sink pads `ref` and
 * `test` with caps audio/x-raw,format=F32LE,layout=interleaved,rate=48000
 * (gstpeaq.c:146-165), caps negotiation by intersecting with the peer of the
 * OTHER pad (:215-244), EOS aggregation over both pads (:668-688), properties
 * playback_level / advanced / di / odg / totalsnr / console-output (:273-317),
 * evaluation on PAUSED->READY with the console output formats of
 * calculate_di_basic / _advanced / calculate_odg (:1012-1078).
 *
 * GStreamer is not available in the build image, so the real build is gated:
 * `make -C gstpeaq_b200/csrc gst` builds libgstpeaqb200.so when
 * `pkg-config gstreamer-1.0` succeeds.  What IS verified here: the file compiles with
 * -Wall -Wextra -Werror against a stand-in for the GStreamer / GObject API with the real
 * signatures (tests/gst_stub/), and tests/test_gst_element.py drives the element through that
 * stand-in like a pipeline would -- caps queries and CAPS events on both pads, buffers of unequal
 * sizes, EOS aggregation, PAUSED->READY, the properties -- and compares odg / di and the console
 * output with the reference's known answers.
 *
 * Divergences from gstpeaq.c, on purpose:
 *  - channels are restricted to 1..2 in the pad templates (the engine's kernels are built for
 *    mono and stereo; the reference takes any count, gstpeaq.c:575-586);
 *  - a CAPS event that does not change the channel count keeps the model state (the reference
 *    reallocates it on every CAPS event, gstpeaq.c:575-586); renegotiation mid-stream to the same
 *    format therefore does not restart the measurement.
 */
#include <gst/gst.h>
#include <string.h>

#include "../../include/peaq_b200.h"

#define GST_TYPE_PEAQ_B200 (gst_peaq_b200_get_type ())
G_DECLARE_FINAL_TYPE (GstPeaqB200, gst_peaq_b200, GST, PEAQ_B200, GstElement)

struct _GstPeaqB200
{
  GstElement element;
  GstPad *refpad, *testpad;
  gboolean ref_eos, test_eos;
  gboolean console_output;
  gboolean advanced;
  gdouble playback_level;
  gint channels;
  gint device;
  peaq_b200_session *session;
};

G_DEFINE_TYPE (GstPeaqB200, gst_peaq_b200, GST_TYPE_ELEMENT)

enum { PROP_0, PROP_PLAYBACK_LEVEL, PROP_ADVANCED, PROP_DI, PROP_ODG, PROP_TOTALSNR,
  PROP_CONSOLE_OUTPUT, PROP_DEVICE };

#define PEAQ_CAPS "audio/x-raw, format = F32LE, layout = interleaved, rate = (int) 48000, channels = (int) [ 1, 2 ]"
static GstStaticPadTemplate ref_template =
GST_STATIC_PAD_TEMPLATE ("ref", GST_PAD_SINK, GST_PAD_ALWAYS, GST_STATIC_CAPS (PEAQ_CAPS));
static GstStaticPadTemplate test_template =
GST_STATIC_PAD_TEMPLATE ("test", GST_PAD_SINK, GST_PAD_ALWAYS, GST_STATIC_CAPS (PEAQ_CAPS));

static gboolean
ensure_session (GstPeaqB200 * self)
{
  if (self->session)
    return TRUE;
  if (peaq_b200_session_create (&self->session, self->device) != 0) {
    GST_ELEMENT_ERROR (self, LIBRARY, INIT, ("%s", peaq_b200_last_error ()), (NULL));
    return FALSE;
  }
  if (peaq_b200_session_set_playback_level (self->session, self->playback_level) != 0 ||
      peaq_b200_session_set_advanced (self->session, self->advanced) != 0 ||
      (self->channels > 0 && peaq_b200_session_set_channels (self->session, self->channels) != 0)) {
    GST_ELEMENT_ERROR (self, LIBRARY, SETTINGS, ("%s", peaq_b200_last_error ()), (NULL));
    peaq_b200_session_destroy (self->session);
    self->session = NULL;
    return FALSE;
  }
  return TRUE;
}

/* console output of calculate_di_basic / _advanced / calculate_odg */
static void
print_result (GstPeaqB200 * self, const peaq_b200_result * r)
{
  const double *m = r->movs;
  if (self->advanced)
    g_print ("RmsModDiffA = %f\nRmsNoiseLoudAsymA = %f\nSegmentalNMRB = %f\nEHSB = %f\n"
        "AvgLinDistA = %f\n", m[0], m[1], m[2], m[3], m[4]);
  else
    g_print ("   BandwidthRefB: %f\n  BandwidthTestB: %f\n      Total NMRB: %f\n"
        "    WinModDiff1B: %f\n            ADBB: %f\n            EHSB: %f\n"
        "    AvgModDiff1B: %f\n    AvgModDiff2B: %f\n   RmsNoiseLoudB: %f\n"
        "           MFPDB: %f\n  RelDistFramesB: %f\n", m[0], m[1], m[2], m[3], m[4], m[5], m[6],
        m[7], m[8], m[9], m[10]);
  g_print ("Objective Difference Grade: %.3f\n", r->odg);
}

static void
gst_peaq_b200_get_property (GObject * obj, guint id, GValue * value, GParamSpec * pspec)
{
  GstPeaqB200 *self = GST_PEAQ_B200 (obj);
  peaq_b200_result r;
  memset (&r, 0, sizeof r);
  if ((id == PROP_DI || id == PROP_ODG || id == PROP_TOTALSNR) && ensure_session (self))
    peaq_b200_session_get_result (self->session, &r);
  switch (id) {
    case PROP_PLAYBACK_LEVEL: g_value_set_double (value, self->playback_level); break;
    case PROP_ADVANCED: g_value_set_boolean (value, self->advanced); break;
    case PROP_DI: g_value_set_double (value, r.di); break;
    case PROP_ODG: g_value_set_double (value, r.odg); break;
    case PROP_TOTALSNR: g_value_set_double (value, r.totalsnr); break;
    case PROP_CONSOLE_OUTPUT: g_value_set_boolean (value, self->console_output); break;
    case PROP_DEVICE: g_value_set_int (value, self->device); break;
    default: G_OBJECT_WARN_INVALID_PROPERTY_ID (obj, id, pspec);
  }
}

static void
gst_peaq_b200_set_property (GObject * obj, guint id, const GValue * value, GParamSpec * pspec)
{
  GstPeaqB200 *self = GST_PEAQ_B200 (obj);
  switch (id) {
    case PROP_PLAYBACK_LEVEL:{
      /* takes effect at once on a running measurement, state kept (gstpeaq.c:509-514); the
       * property only reports a level the engine has accepted */
      const gdouble level = g_value_get_double (value);
      if (self->session && peaq_b200_session_set_playback_level (self->session, level) != 0)
        GST_WARNING_OBJECT (self, "playback_level not changed: %s", peaq_b200_last_error ());
      else
        self->playback_level = level;
      break;
    }
    case PROP_ADVANCED:{
      /* rebuilds the models, like the reference (gstpeaq.c:516-560) */
      const gboolean advanced = g_value_get_boolean (value);
      if (self->session && peaq_b200_session_set_advanced (self->session, advanced) != 0)
        GST_WARNING_OBJECT (self, "advanced not changed: %s", peaq_b200_last_error ());
      else
        self->advanced = advanced;
      break;
    }
    case PROP_CONSOLE_OUTPUT: self->console_output = g_value_get_boolean (value); break;
    case PROP_DEVICE: self->device = g_value_get_int (value); break;
    default: G_OBJECT_WARN_INVALID_PROPERTY_ID (obj, id, pspec);
  }
}

/* caps of one pad = template caps restricted to what the peer of the OTHER pad offers */
static gboolean
gst_peaq_b200_pad_query (GstPad * pad, GstObject * parent, GstQuery * query)
{
  GstPeaqB200 *self = GST_PEAQ_B200 (parent);
  if (GST_QUERY_TYPE (query) == GST_QUERY_CAPS) {
    GstCaps *filter, *mine, *other, *result;
    gst_query_parse_caps (query, &filter);
    mine = gst_pad_get_pad_template_caps (pad);
    other = gst_pad_peer_query_caps (pad == self->refpad ? self->testpad : self->refpad, filter);
    result = gst_caps_intersect (mine, other);
    gst_caps_unref (mine);
    gst_caps_unref (other);
    gst_query_set_caps_result (query, result);
    gst_caps_unref (result);
    return TRUE;
  }
  return gst_pad_query_default (pad, parent, query);
}

static gboolean
gst_peaq_b200_pad_event (GstPad * pad, GstObject * parent, GstEvent * event)
{
  GstPeaqB200 *self = GST_PEAQ_B200 (parent);
  gboolean ret = FALSE;
  switch (GST_EVENT_TYPE (event)) {
    case GST_EVENT_EOS:
      GST_OBJECT_LOCK (self);
      if (pad == self->refpad)
        self->ref_eos = TRUE;
      else
        self->test_eos = TRUE;
      ret = self->ref_eos && self->test_eos;
      GST_OBJECT_UNLOCK (self);
      if (ret) {
        GstMessage *msg = gst_message_new_eos (parent);
        gst_message_set_seqnum (msg, gst_event_get_seqnum (event));
        ret = gst_element_post_message (GST_ELEMENT (self), msg);
      }
      gst_event_unref (event);
      break;
    case GST_EVENT_CAPS:{
      GstCaps *caps;
      gst_event_parse_caps (event, &caps);
      if (gst_pad_peer_query_accept_caps (pad == self->refpad ? self->testpad : self->refpad, caps)) {
        gint channels = 0;
        gst_structure_get_int (gst_caps_get_structure (caps, 0), "channels", &channels);
        GST_OBJECT_LOCK (self);
        ret = TRUE;
        if (channels != self->channels) {
          if (self->session && peaq_b200_session_set_channels (self->session, channels) != 0) {
            GST_WARNING_OBJECT (self, "caps refused: %s", peaq_b200_last_error ());
            ret = FALSE;
          } else
            self->channels = channels;
        }
        GST_OBJECT_UNLOCK (self);
      }
      gst_event_unref (event);
      break;
    }
    default:
      ret = gst_pad_event_default (pad, parent, event);
  }
  return ret;
}

static GstFlowReturn
gst_peaq_b200_pad_chain (GstPad * pad, GstObject * parent, GstBuffer * buffer)
{
  GstPeaqB200 *self = GST_PEAQ_B200 (parent);
  GstMapInfo map;
  GST_OBJECT_LOCK (self);
  if (pad == self->refpad)
    self->ref_eos = FALSE;
  else
    self->test_eos = FALSE;
  if (self->channels > 0 && ensure_session (self) && gst_buffer_map (buffer, &map, GST_MAP_READ)) {
    if (peaq_b200_session_push (self->session,
            pad == self->refpad ? PEAQ_B200_PAD_REF : PEAQ_B200_PAD_TEST,
            (const float *) map.data, map.size / (sizeof (gfloat) * self->channels)) != 0)
      GST_WARNING_OBJECT (self, "%s", peaq_b200_last_error ());
    gst_buffer_unmap (buffer, &map);
  }
  GST_OBJECT_UNLOCK (self);
  gst_buffer_unref (buffer);
  return GST_FLOW_OK;           /* the reference never fails the chain (gstpeaq.c:660) */
}

static GstStateChangeReturn
gst_peaq_b200_change_state (GstElement * element, GstStateChange transition)
{
  GstPeaqB200 *self = GST_PEAQ_B200 (element);
  if (transition == GST_STATE_CHANGE_PAUSED_TO_READY && self->session) {
    peaq_b200_result r;
    peaq_b200_session_finish (self->session);   /* do_flush on both clocks */
    if (peaq_b200_session_get_result (self->session, &r) == 0 && self->console_output)
      print_result (self, &r);
  }
  return GST_ELEMENT_CLASS (gst_peaq_b200_parent_class)->change_state (element, transition);
}

static void
gst_peaq_b200_finalize (GObject * obj)
{
  GstPeaqB200 *self = GST_PEAQ_B200 (obj);
  peaq_b200_session_destroy (self->session);
  G_OBJECT_CLASS (gst_peaq_b200_parent_class)->finalize (obj);
}

static void
gst_peaq_b200_class_init (GstPeaqB200Class * klass)
{
  GObjectClass *oc = G_OBJECT_CLASS (klass);
  GstElementClass *ec = GST_ELEMENT_CLASS (klass);
  oc->get_property = gst_peaq_b200_get_property;
  oc->set_property = gst_peaq_b200_set_property;
  oc->finalize = gst_peaq_b200_finalize;
  ec->change_state = gst_peaq_b200_change_state;
  g_object_class_install_property (oc, PROP_PLAYBACK_LEVEL,
      g_param_spec_double ("playback_level", "playback level", "Playback level in dB", 0, 130, 92,
          G_PARAM_READWRITE | G_PARAM_CONSTRUCT));
  g_object_class_install_property (oc, PROP_ADVANCED,
      g_param_spec_boolean ("advanced", "Advanced mode enabled", "True if advanced mode is used", FALSE,
          G_PARAM_READWRITE | G_PARAM_CONSTRUCT));
  g_object_class_install_property (oc, PROP_DI,
      g_param_spec_double ("di", "distortion index", "Distortion Index", -G_MAXDOUBLE, G_MAXDOUBLE, 0,
          G_PARAM_READABLE));
  g_object_class_install_property (oc, PROP_ODG,
      g_param_spec_double ("odg", "objective difference grade", "Objective Difference Grade",
          -G_MAXDOUBLE, G_MAXDOUBLE, 0, G_PARAM_READABLE));
  g_object_class_install_property (oc, PROP_TOTALSNR,
      g_param_spec_double ("totalsnr", "the overall SNR in dB", "the overall signal to noise ratio in dB",
          -G_MAXDOUBLE, G_MAXDOUBLE, 0, G_PARAM_READABLE));
  g_object_class_install_property (oc, PROP_CONSOLE_OUTPUT,
      g_param_spec_boolean ("console-output", "console output", "Enable or disable console output", TRUE,
          G_PARAM_READWRITE | G_PARAM_CONSTRUCT));
  g_object_class_install_property (oc, PROP_DEVICE,
      g_param_spec_int ("device", "CUDA device", "CUDA device the engine runs on", 0, 1023, 0,
          G_PARAM_READWRITE | G_PARAM_CONSTRUCT));
  gst_element_class_add_static_pad_template (ec, &ref_template);
  gst_element_class_add_static_pad_template (ec, &test_template);
  gst_element_class_set_static_metadata (ec, "Perceptual evaluation of audio quality (B200)",
      "Sink/Audio", "Compute objective audio quality measures (ITU-R BS.1387) on a CUDA device",
      "peaq-b200");
}

static void
gst_peaq_b200_init (GstPeaqB200 * self)
{
  self->refpad = gst_pad_new_from_static_template (&ref_template, "ref");
  self->testpad = gst_pad_new_from_static_template (&test_template, "test");
  GstPad *pads[2] = { self->refpad, self->testpad };
  for (int i = 0; i < 2; i++) {
    gst_pad_set_chain_function (pads[i], gst_peaq_b200_pad_chain);
    gst_pad_set_event_function (pads[i], gst_peaq_b200_pad_event);
    gst_pad_set_query_function (pads[i], gst_peaq_b200_pad_query);
    gst_element_add_pad (GST_ELEMENT (self), pads[i]);
  }
  GST_OBJECT_FLAG_SET (self, GST_ELEMENT_FLAG_SINK);
  self->playback_level = 92.;
  self->console_output = TRUE;
}

static gboolean
plugin_init (GstPlugin * plugin)
{
  /* same element name and rank as the reference (gstpeaqplugin.c:30-34) */
  return gst_element_register (plugin, "peaq", GST_RANK_NONE, GST_TYPE_PEAQ_B200);
}

#ifndef PACKAGE
#define PACKAGE "peaq-b200"
#endif
GST_PLUGIN_DEFINE (GST_VERSION_MAJOR, GST_VERSION_MINOR, peaqb200,
    "Perceptual evaluation of audio quality on CUDA (B200)", plugin_init, "1", "LGPL", "peaq-b200",
    "https://example.invalid/peaq-b200")
