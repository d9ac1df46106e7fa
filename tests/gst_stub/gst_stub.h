/* harness-side functions of the GStreamer stand-in (TEST INFRASTRUCTURE ONLY, see gst/gst.h) */
#ifndef PEAQ_GST_STUB_HARNESS_H
#define PEAQ_GST_STUB_HARNESS_H
#include <gst/gst.h>
GstCaps *gst_stub_caps_new (gint channels);
gboolean gst_stub_caps_is_empty (GstCaps * c);
int gst_stub_error_count (void);
GstElement *gst_stub_factory_make (const char *name);
GstPad *gst_stub_get_pad (GstElement * e, const char *name);
void gst_stub_pad_set_peer_caps (GstPad * pad, gint channels);
gboolean gst_stub_pad_send_caps (GstPad * pad, gint channels);
gboolean gst_stub_pad_send_eos (GstPad * pad, guint32 seqnum);
GstFlowReturn gst_stub_pad_push (GstPad * pad, const float *samples, size_t n_floats);
gint gst_stub_pad_query_caps (GstPad * pad, gint filter_channels);
GstStateChangeReturn gst_stub_change_state (GstElement * e, GstStateChange transition);
const GstElementClass *gst_stub_element_class (GstElement * e);
gboolean gst_stub_plugin_init_peaqb200 (GstPlugin * plugin);
#endif
