/* gst/gst.h -- a small stand-in for the GStreamer-1.0 / GObject API (TEST INFRASTRUCTURE ONLY).
 *
 * GStreamer and GLib are not installed in the build image.  This header declares, with the
 * signatures of the real GStreamer 1.x / GLib 2.x headers, exactly the symbols the element shell
 * gstpeaq_b200/gst/gstpeaqb200.c uses; tests/gst_stub/gst_stub.c implements them just far enough
 * to drive the element the way a pipeline would (two peers with caps, buffers, CAPS and EOS
 * events, a bus, properties, state changes).  tests/test_gst_element.py compiles the element
 * against it with -Wall -Werror (catches type / signature rot without a GPU) and runs the harness
 * on the GPU box.  It is NOT GStreamer: the real build is `make -C gstpeaq_b200/csrc gst`.
 */
#ifndef PEAQ_GST_STUB_H
#define PEAQ_GST_STUB_H

#include <float.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

typedef int gboolean;
typedef int gint;
typedef unsigned int guint;
typedef char gchar;
typedef float gfloat;
typedef double gdouble;
typedef void *gpointer;
typedef const void *gconstpointer;
typedef size_t gsize;
typedef unsigned char guint8;
typedef uint32_t guint32;
typedef uint64_t guint64;
typedef size_t GType;

#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
#define G_MAXDOUBLE DBL_MAX

/* ---- GObject ---------------------------------------------------------------------- */
typedef struct _GObject GObject;
typedef struct _GObjectClass GObjectClass;
typedef struct _GParamSpec GParamSpec;
typedef struct _GValue GValue;

typedef enum {
  G_PARAM_READABLE = 1 << 0,
  G_PARAM_WRITABLE = 1 << 1,
  G_PARAM_READWRITE = (1 << 0) | (1 << 1),
  G_PARAM_CONSTRUCT = 1 << 2
} GParamFlags;

struct _GValue {
  int kind;                     /* 0 double, 1 boolean, 2 int */
  union { gdouble d; gboolean b; gint i; } v;
};

struct _GParamSpec {
  const gchar *name;
  int kind;
  GValue def;
  GParamFlags flags;
  guint id;
  GParamSpec *next;
};

struct _GObject {
  GObjectClass *klass;
  int ref_count;
};

struct _GObjectClass {
  GType type;
  void (*set_property) (GObject * object, guint property_id, const GValue * value, GParamSpec * pspec);
  void (*get_property) (GObject * object, guint property_id, GValue * value, GParamSpec * pspec);
  void (*finalize) (GObject * object);
  GParamSpec *pspecs;           /* stub: installed properties */
};

#define G_OBJECT(obj) ((GObject *) (obj))
#define G_OBJECT_CLASS(klass) ((GObjectClass *) (klass))
#define G_OBJECT_WARN_INVALID_PROPERTY_ID(obj, id, pspec) \
  fprintf (stderr, "%p: invalid property id %u (%s)\n", (void *) (obj), (guint) (id), (pspec)->name)

void g_object_class_install_property (GObjectClass * oclass, guint property_id, GParamSpec * pspec);
GParamSpec *g_param_spec_double (const gchar * name, const gchar * nick, const gchar * blurb,
    gdouble minimum, gdouble maximum, gdouble default_value, GParamFlags flags);
GParamSpec *g_param_spec_boolean (const gchar * name, const gchar * nick, const gchar * blurb,
    gboolean default_value, GParamFlags flags);
GParamSpec *g_param_spec_int (const gchar * name, const gchar * nick, const gchar * blurb,
    gint minimum, gint maximum, gint default_value, GParamFlags flags);
void g_value_set_double (GValue * value, gdouble v);
void g_value_set_boolean (GValue * value, gboolean v);
void g_value_set_int (GValue * value, gint v);
gdouble g_value_get_double (const GValue * value);
gboolean g_value_get_boolean (const GValue * value);
gint g_value_get_int (const GValue * value);
void g_print (const gchar * format, ...) __attribute__ ((format (printf, 1, 2)));
gpointer g_object_new (GType object_type, const gchar * first_property_name, ...);
void g_object_set (gpointer object, const gchar * first_property_name, ...);
void g_object_get (gpointer object, const gchar * first_property_name, ...);
void g_object_unref (gpointer object);

/* type registration behind G_DEFINE_TYPE */
GType g_stub_register_type (const gchar * name, GType parent, size_t instance_size, size_t class_size,
    void (*class_init) (gpointer klass), void (*instance_init) (gpointer instance), gpointer * parent_class);

#define G_DECLARE_FINAL_TYPE(ModuleObjName, module_obj_name, MODULE, OBJ_NAME, ParentName) \
  GType module_obj_name##_get_type (void); \
  typedef struct _##ModuleObjName ModuleObjName; \
  typedef struct { ParentName##Class parent_class; } ModuleObjName##Class; \
  static inline ModuleObjName *MODULE##_##OBJ_NAME (gpointer ptr) { return (ModuleObjName *) ptr; }

#define G_DEFINE_TYPE(TN, t_n, T_P) \
  static void t_n##_init (TN * self); \
  static void t_n##_class_init (TN##Class * klass); \
  static gpointer t_n##_parent_class = NULL; \
  GType t_n##_get_type (void) { \
    static GType type = 0; \
    if (!type) \
      type = g_stub_register_type (#TN, T_P, sizeof (TN), sizeof (TN##Class), \
          (void (*)(gpointer)) t_n##_class_init, (void (*)(gpointer)) t_n##_init, &t_n##_parent_class); \
    return type; \
  }

/* ---- GStreamer ---------------------------------------------------------------------- */
typedef struct _GstObject GstObject;
typedef struct _GstElement GstElement;
typedef struct _GstElementClass GstElementClass;
typedef struct _GstPad GstPad;
typedef struct _GstCaps GstCaps;
typedef struct _GstStructure GstStructure;
typedef struct _GstEvent GstEvent;
typedef struct _GstQuery GstQuery;
typedef struct _GstMessage GstMessage;
typedef struct _GstBuffer GstBuffer;
typedef struct _GstPlugin GstPlugin;

struct _GstObject {
  GObject object;
  guint32 flags;
  int lock_depth;               /* stub: GST_OBJECT_LOCK bookkeeping */
  GstObject *parent;
};

typedef enum { GST_PAD_UNKNOWN, GST_PAD_SRC, GST_PAD_SINK } GstPadDirection;
typedef enum { GST_PAD_ALWAYS, GST_PAD_SOMETIMES, GST_PAD_REQUEST } GstPadPresence;
typedef enum { GST_FLOW_OK = 0, GST_FLOW_ERROR = -5 } GstFlowReturn;
typedef enum { GST_MAP_READ = 1, GST_MAP_WRITE = 2 } GstMapFlags;
typedef enum { GST_RANK_NONE = 0 } GstRank;
typedef enum { GST_STATE_CHANGE_FAILURE = 0, GST_STATE_CHANGE_SUCCESS = 1 } GstStateChangeReturn;
typedef enum {
  GST_STATE_CHANGE_NULL_TO_READY = (1 << 3) | 2,
  GST_STATE_CHANGE_READY_TO_PAUSED = (2 << 3) | 3,
  GST_STATE_CHANGE_PAUSED_TO_PLAYING = (3 << 3) | 4,
  GST_STATE_CHANGE_PLAYING_TO_PAUSED = (4 << 3) | 3,
  GST_STATE_CHANGE_PAUSED_TO_READY = (3 << 3) | 2,
  GST_STATE_CHANGE_READY_TO_NULL = (2 << 3) | 1
} GstStateChange;
typedef enum { GST_EVENT_UNKNOWN = 0, GST_EVENT_EOS = 1, GST_EVENT_CAPS = 2, GST_EVENT_SEGMENT = 3 } GstEventType;
typedef enum { GST_QUERY_UNKNOWN = 0, GST_QUERY_CAPS = 1, GST_QUERY_ACCEPT_CAPS = 2 } GstQueryType;
enum { GST_ELEMENT_FLAG_SINK = 1 << 5 };

typedef struct { const gchar *string; } GstStaticCaps;
#define GST_STATIC_CAPS(str) { str }
typedef struct {
  const gchar *name_template;
  GstPadDirection direction;
  GstPadPresence presence;
  GstStaticCaps static_caps;
} GstStaticPadTemplate;
#define GST_STATIC_PAD_TEMPLATE(padname, dir, pres, caps) { padname, dir, pres, caps }

typedef struct {
  GstBuffer *memory;
  GstMapFlags flags;
  guint8 *data;
  gsize size;
  gsize maxsize;
} GstMapInfo;

typedef GstFlowReturn (*GstPadChainFunction) (GstPad * pad, GstObject * parent, GstBuffer * buffer);
typedef gboolean (*GstPadEventFunction) (GstPad * pad, GstObject * parent, GstEvent * event);
typedef gboolean (*GstPadQueryFunction) (GstPad * pad, GstObject * parent, GstQuery * query);

struct _GstElement {
  GstObject object;
  GstPad *pads[4];              /* stub */
  int n_pads;
  int messages_eos;             /* stub bus: EOS messages posted */
  guint32 last_eos_seqnum;
};

struct _GstElementClass {
  GObjectClass parent_class;
  GstStateChangeReturn (*change_state) (GstElement * element, GstStateChange transition);
  GstStaticPadTemplate *templates[4];   /* stub */
  int n_templates;
  const gchar *longname, *klass, *description, *author;
};

GType gst_element_get_type (void);
#define GST_TYPE_ELEMENT (gst_element_get_type ())
#define GST_ELEMENT(obj) ((GstElement *) (obj))
#define GST_ELEMENT_CLASS(klass) ((GstElementClass *) (klass))
#define GST_OBJECT(obj) ((GstObject *) (obj))
#define GST_OBJECT_LOCK(obj) (((GstObject *) (obj))->lock_depth++)
#define GST_OBJECT_UNLOCK(obj) (((GstObject *) (obj))->lock_depth--)
#define GST_OBJECT_FLAG_SET(obj, flag) (((GstObject *) (obj))->flags |= (flag))
#define GST_EVENT_TYPE(event) (gst_stub_event_type (event))
#define GST_QUERY_TYPE(query) (gst_stub_query_type (query))
GstEventType gst_stub_event_type (GstEvent * event);
GstQueryType gst_stub_query_type (GstQuery * query);

#define GST_WARNING_OBJECT(obj, ...) do { fprintf (stderr, "WARNING: "); fprintf (stderr, __VA_ARGS__); fprintf (stderr, "\n"); } while (0)
/* like the real macro: posts an error message; the stub prints it and counts it */
void gst_stub_element_error (GstElement * element, const gchar * text);
#define GST_ELEMENT_ERROR(el, domain, code, text, debug) do { \
    char gst_stub_buf__[512]; snprintf gst_stub_args__ text; gst_stub_element_error (GST_ELEMENT (el), gst_stub_buf__); } while (0)
#define gst_stub_args__(...) (gst_stub_buf__, sizeof gst_stub_buf__, __VA_ARGS__)

void gst_element_class_add_static_pad_template (GstElementClass * klass, GstStaticPadTemplate * static_templ);
void gst_element_class_set_static_metadata (GstElementClass * klass, const gchar * longname,
    const gchar * classification, const gchar * description, const gchar * author);
gboolean gst_element_add_pad (GstElement * element, GstPad * pad);
gboolean gst_element_post_message (GstElement * element, GstMessage * message);
gboolean gst_element_register (GstPlugin * plugin, const gchar * name, guint rank, GType type);

GstPad *gst_pad_new_from_static_template (GstStaticPadTemplate * templ, const gchar * name);
void gst_pad_set_chain_function (GstPad * pad, GstPadChainFunction chain);
void gst_pad_set_event_function (GstPad * pad, GstPadEventFunction event);
void gst_pad_set_query_function (GstPad * pad, GstPadQueryFunction query);
GstCaps *gst_pad_get_pad_template_caps (GstPad * pad);
GstCaps *gst_pad_peer_query_caps (GstPad * pad, GstCaps * filter);
gboolean gst_pad_peer_query_accept_caps (GstPad * pad, GstCaps * caps);
gboolean gst_pad_query_default (GstPad * pad, GstObject * parent, GstQuery * query);
gboolean gst_pad_event_default (GstPad * pad, GstObject * parent, GstEvent * event);

GstCaps *gst_caps_intersect (GstCaps * caps1, GstCaps * caps2);
void gst_caps_unref (GstCaps * caps);
GstStructure *gst_caps_get_structure (const GstCaps * caps, guint index);
gboolean gst_structure_get_int (const GstStructure * structure, const gchar * fieldname, gint * value);

void gst_query_parse_caps (GstQuery * query, GstCaps ** filter);
void gst_query_set_caps_result (GstQuery * query, GstCaps * caps);

void gst_event_parse_caps (GstEvent * event, GstCaps ** caps);
guint32 gst_event_get_seqnum (GstEvent * event);
void gst_event_unref (GstEvent * event);

GstMessage *gst_message_new_eos (GstObject * src);
void gst_message_set_seqnum (GstMessage * message, guint32 seqnum);

gboolean gst_buffer_map (GstBuffer * buffer, GstMapInfo * info, GstMapFlags flags);
void gst_buffer_unmap (GstBuffer * buffer, GstMapInfo * info);
void gst_buffer_unref (GstBuffer * buffer);

#define GST_VERSION_MAJOR 1
#define GST_VERSION_MINOR 18
#define GST_PLUGIN_DEFINE(major, minor, name, description, init, version, license, package, origin) \
  gboolean gst_stub_plugin_init_##name (GstPlugin * plugin) { return init (plugin); }

#endif /* PEAQ_GST_STUB_H */
