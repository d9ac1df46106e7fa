/* gst_stub.c -- implementation of the stand-in declared in gst/gst.h plus the harness functions
 * the test driver uses to play "pipeline" (TEST INFRASTRUCTURE ONLY, see gst/gst.h). */
#include "gst_stub.h"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

/* ---- types ---------------------------------------------------------------------------- */
typedef struct {
  const char *name;
  GType parent;
  size_t instance_size, class_size;
  void (*class_init) (gpointer);
  void (*instance_init) (gpointer);
  void *klass;
} TypeInfo;
static TypeInfo g_types[16];
static int g_n_types = 1;        /* GType 0 = invalid */

static GstStateChangeReturn
element_change_state (GstElement * element, GstStateChange transition)
{
  (void) element; (void) transition;
  return GST_STATE_CHANGE_SUCCESS;
}
static void object_finalize (GObject * obj) { (void) obj; }

GType
gst_element_get_type (void)
{
  static GType t = 0;
  if (!t) {
    t = g_n_types++;
    g_types[t].name = "GstElement";
    g_types[t].instance_size = sizeof (GstElement);
    g_types[t].class_size = sizeof (GstElementClass);
    GstElementClass *k = calloc (1, sizeof *k);
    k->change_state = element_change_state;
    k->parent_class.finalize = object_finalize;
    k->parent_class.type = t;
    g_types[t].klass = k;
  }
  return t;
}

GType
g_stub_register_type (const gchar * name, GType parent, size_t instance_size, size_t class_size,
    void (*class_init) (gpointer), void (*instance_init) (gpointer), gpointer * parent_class)
{
  GType t = g_n_types++;
  TypeInfo *ti = &g_types[t];
  ti->name = name; ti->parent = parent; ti->instance_size = instance_size; ti->class_size = class_size;
  ti->class_init = class_init; ti->instance_init = instance_init;
  ti->klass = calloc (1, class_size);
  memcpy (ti->klass, g_types[parent].klass, g_types[parent].class_size);   /* inherit the vtable */
  ((GObjectClass *) ti->klass)->type = t;
  ((GObjectClass *) ti->klass)->pspecs = NULL;
  *parent_class = g_types[parent].klass;
  class_init (ti->klass);
  return t;
}

/* ---- properties ------------------------------------------------------------------------- */
static GParamSpec *
new_pspec (const gchar * name, int kind, GParamFlags flags)
{
  GParamSpec *p = calloc (1, sizeof *p);
  p->name = name; p->kind = kind; p->flags = flags; p->def.kind = kind;
  return p;
}
GParamSpec *g_param_spec_double (const gchar * name, const gchar * nick, const gchar * blurb, gdouble minimum,
    gdouble maximum, gdouble default_value, GParamFlags flags)
{ (void) nick; (void) blurb; (void) minimum; (void) maximum;
  GParamSpec *p = new_pspec (name, 0, flags); p->def.v.d = default_value; return p; }
GParamSpec *g_param_spec_boolean (const gchar * name, const gchar * nick, const gchar * blurb,
    gboolean default_value, GParamFlags flags)
{ (void) nick; (void) blurb; GParamSpec *p = new_pspec (name, 1, flags); p->def.v.b = default_value; return p; }
GParamSpec *g_param_spec_int (const gchar * name, const gchar * nick, const gchar * blurb, gint minimum,
    gint maximum, gint default_value, GParamFlags flags)
{ (void) nick; (void) blurb; (void) minimum; (void) maximum;
  GParamSpec *p = new_pspec (name, 2, flags); p->def.v.i = default_value; return p; }
void g_object_class_install_property (GObjectClass * oclass, guint property_id, GParamSpec * pspec)
{ pspec->id = property_id; pspec->next = oclass->pspecs; oclass->pspecs = pspec; }
void g_value_set_double (GValue * value, gdouble v) { value->kind = 0; value->v.d = v; }
void g_value_set_boolean (GValue * value, gboolean v) { value->kind = 1; value->v.b = v; }
void g_value_set_int (GValue * value, gint v) { value->kind = 2; value->v.i = v; }
gdouble g_value_get_double (const GValue * value) { return value->v.d; }
gboolean g_value_get_boolean (const GValue * value) { return value->v.b; }
gint g_value_get_int (const GValue * value) { return value->v.i; }

void g_print (const gchar * format, ...)
{ va_list ap; va_start (ap, format); vprintf (format, ap); va_end (ap); }

static GParamSpec *
find_pspec (GObject * o, const char *name)
{
  for (GParamSpec * p = o->klass->pspecs; p; p = p->next) {
    /* GObject treats '-' and '_' in property names alike */
    const char *a = p->name, *b = name;
    while (*a && *b && (*a == *b || ((*a == '-' || *a == '_') && (*b == '-' || *b == '_')))) { a++; b++; }
    if (!*a && !*b) return p;
  }
  fprintf (stderr, "no property %s\n", name);
  abort ();
}

static void
set_valist (GObject * o, const gchar * name, va_list ap)
{
  while (name) {
    GParamSpec *p = find_pspec (o, name);
    GValue v; v.kind = p->kind;
    if (p->kind == 0) v.v.d = va_arg (ap, double);
    else if (p->kind == 1) v.v.b = va_arg (ap, int);
    else v.v.i = va_arg (ap, int);
    o->klass->set_property (o, p->id, &v, p);
    name = va_arg (ap, const gchar *);
  }
}

gpointer
g_object_new (GType type, const gchar * first_property_name, ...)
{
  TypeInfo *ti = &g_types[type];
  GObject *o = calloc (1, ti->instance_size);
  o->klass = ti->klass; o->ref_count = 1;
  ti->instance_init (o);
  for (GParamSpec * p = o->klass->pspecs; p; p = p->next)   /* G_PARAM_CONSTRUCT defaults */
    if (p->flags & G_PARAM_CONSTRUCT) o->klass->set_property (o, p->id, &p->def, p);
  va_list ap; va_start (ap, first_property_name); set_valist (o, first_property_name, ap); va_end (ap);
  return o;
}
void g_object_set (gpointer object, const gchar * first_property_name, ...)
{ va_list ap; va_start (ap, first_property_name); set_valist (object, first_property_name, ap); va_end (ap); }
void g_object_get (gpointer object, const gchar * name, ...)
{
  GObject *o = object;
  va_list ap; va_start (ap, name);
  while (name) {
    GParamSpec *p = find_pspec (o, name);
    GValue v; memset (&v, 0, sizeof v);
    o->klass->get_property (o, p->id, &v, p);
    void *dst = va_arg (ap, void *);
    if (p->kind == 0) *(double *) dst = v.v.d; else if (p->kind == 1) *(int *) dst = v.v.b; else *(int *) dst = v.v.i;
    name = va_arg (ap, const gchar *);
  }
  va_end (ap);
}
void g_object_unref (gpointer object)
{ GObject *o = object; if (--o->ref_count == 0) { o->klass->finalize (o); free (o); } }

/* ---- caps: audio/x-raw F32LE interleaved 48 kHz with a channel count (0 = any) ----------------- */
struct _GstStructure { gint channels; };
struct _GstCaps { int ref_count; gboolean empty; struct _GstStructure s; };
GstCaps *gst_stub_caps_new (gint channels)
{ GstCaps *c = calloc (1, sizeof *c); c->ref_count = 1; c->s.channels = channels; return c; }
gboolean gst_stub_caps_is_empty (GstCaps * c) { return c->empty; }
GstCaps *gst_caps_intersect (GstCaps * a, GstCaps * b)
{
  GstCaps *r = gst_stub_caps_new (0);
  if (a->empty || b->empty || (a->s.channels && b->s.channels && a->s.channels != b->s.channels)) r->empty = TRUE;
  else r->s.channels = a->s.channels ? a->s.channels : b->s.channels;
  return r;
}
void gst_caps_unref (GstCaps * caps) { if (--caps->ref_count == 0) free (caps); }
GstStructure *gst_caps_get_structure (const GstCaps * caps, guint index) { (void) index; return (GstStructure *) &caps->s; }
gboolean gst_structure_get_int (const GstStructure * s, const gchar * fieldname, gint * value)
{ if (strcmp (fieldname, "channels") || !s->channels) return FALSE; *value = s->channels; return TRUE; }

/* ---- pads ------------------------------------------------------------------------------- */
struct _GstPad {
  GstObject object;
  const char *name;
  GstStaticPadTemplate *templ;
  GstPadChainFunction chain;
  GstPadEventFunction event;
  GstPadQueryFunction query;
  GstCaps *peer_caps;           /* what the upstream peer can produce (harness) */
};
GstPad *gst_pad_new_from_static_template (GstStaticPadTemplate * templ, const gchar * name)
{ GstPad *p = calloc (1, sizeof *p); p->name = name; p->templ = templ; return p; }
void gst_pad_set_chain_function (GstPad * pad, GstPadChainFunction f) { pad->chain = f; }
void gst_pad_set_event_function (GstPad * pad, GstPadEventFunction f) { pad->event = f; }
void gst_pad_set_query_function (GstPad * pad, GstPadQueryFunction f) { pad->query = f; }
GstCaps *gst_pad_get_pad_template_caps (GstPad * pad)
{
  /* the template string fixes everything but the channel count, or restricts it to a range */
  (void) pad;
  return gst_stub_caps_new (0);
}
GstCaps *gst_pad_peer_query_caps (GstPad * pad, GstCaps * filter)
{
  GstCaps *peer = pad->peer_caps ? pad->peer_caps : NULL;
  GstCaps *any = gst_stub_caps_new (0);
  GstCaps *r = gst_caps_intersect (peer ? peer : any, filter ? filter : any);
  gst_caps_unref (any);
  return r;
}
gboolean gst_pad_peer_query_accept_caps (GstPad * pad, GstCaps * caps)
{ GstCaps *r = gst_pad_peer_query_caps (pad, caps); gboolean ok = !r->empty; gst_caps_unref (r); return ok; }
gboolean gst_pad_query_default (GstPad * pad, GstObject * parent, GstQuery * query)
{ (void) pad; (void) parent; (void) query; return FALSE; }
gboolean gst_pad_event_default (GstPad * pad, GstObject * parent, GstEvent * event)
{ (void) pad; (void) parent; gst_event_unref (event); return TRUE; }

/* ---- events, queries, messages, buffers ------------------------------------------------------- */
struct _GstEvent { GstEventType type; GstCaps *caps; guint32 seqnum; };
struct _GstQuery { GstQueryType type; GstCaps *filter, *result; };
struct _GstMessage { GstObject *src; guint32 seqnum; int is_eos; };
struct _GstBuffer { void *data; size_t size; int mapped; };
GstEventType gst_stub_event_type (GstEvent * e) { return e->type; }
GstQueryType gst_stub_query_type (GstQuery * q) { return q->type; }
void gst_event_parse_caps (GstEvent * e, GstCaps ** caps) { *caps = e->caps; }
guint32 gst_event_get_seqnum (GstEvent * e) { return e->seqnum; }
void gst_event_unref (GstEvent * e) { if (e->caps) gst_caps_unref (e->caps); free (e); }
void gst_query_parse_caps (GstQuery * q, GstCaps ** filter) { *filter = q->filter; }
void gst_query_set_caps_result (GstQuery * q, GstCaps * caps) { caps->ref_count++; q->result = caps; }
GstMessage *gst_message_new_eos (GstObject * src) { GstMessage *m = calloc (1, sizeof *m); m->src = src; m->is_eos = 1; return m; }
void gst_message_set_seqnum (GstMessage * m, guint32 seqnum) { m->seqnum = seqnum; }
gboolean gst_buffer_map (GstBuffer * b, GstMapInfo * info, GstMapFlags flags)
{ memset (info, 0, sizeof *info); info->memory = b; info->flags = flags; info->data = b->data; info->size = info->maxsize = b->size; b->mapped++; return TRUE; }
void gst_buffer_unmap (GstBuffer * b, GstMapInfo * info) { (void) info; b->mapped--; }
void gst_buffer_unref (GstBuffer * b) { if (b->mapped) { fprintf (stderr, "buffer freed while mapped\n"); abort (); } free (b->data); free (b); }

/* ---- element ----------------------------------------------------------------------------- */
static int g_errors = 0;
void gst_stub_element_error (GstElement * element, const gchar * text)
{ (void) element; g_errors++; fprintf (stderr, "ELEMENT ERROR: %s\n", text); }
int gst_stub_error_count (void) { return g_errors; }
void gst_element_class_add_static_pad_template (GstElementClass * klass, GstStaticPadTemplate * t)
{ klass->templates[klass->n_templates++] = t; }
void gst_element_class_set_static_metadata (GstElementClass * klass, const gchar * longname,
    const gchar * classification, const gchar * description, const gchar * author)
{ klass->longname = longname; klass->klass = classification; klass->description = description; klass->author = author; }
gboolean gst_element_add_pad (GstElement * element, GstPad * pad)
{ pad->object.parent = GST_OBJECT (element); element->pads[element->n_pads++] = pad; return TRUE; }
gboolean gst_element_post_message (GstElement * element, GstMessage * message)
{
  if (GST_OBJECT (element)->lock_depth) { fprintf (stderr, "message posted with the object lock held\n"); abort (); }
  if (message->is_eos) { element->messages_eos++; element->last_eos_seqnum = message->seqnum; }
  free (message);
  return TRUE;
}
static struct { const char *name; GType type; guint rank; } g_factories[4];
static int g_n_factories = 0;
gboolean gst_element_register (GstPlugin * plugin, const gchar * name, guint rank, GType type)
{ (void) plugin; g_factories[g_n_factories].name = name; g_factories[g_n_factories].rank = rank; g_factories[g_n_factories++].type = type; return TRUE; }

/* ---- harness side ------------------------------------------------------------------------- */
GstElement *gst_stub_factory_make (const char *name)
{
  for (int i = 0; i < g_n_factories; i++)
    if (!strcmp (g_factories[i].name, name)) return g_object_new (g_factories[i].type, NULL);
  return NULL;
}
GstPad *gst_stub_get_pad (GstElement * e, const char *name)
{ for (int i = 0; i < e->n_pads; i++) if (!strcmp (e->pads[i]->name, name)) return e->pads[i]; return NULL; }
void gst_stub_pad_set_peer_caps (GstPad * pad, gint channels)
{ if (pad->peer_caps) gst_caps_unref (pad->peer_caps); pad->peer_caps = gst_stub_caps_new (channels); }
gboolean gst_stub_pad_send_caps (GstPad * pad, gint channels)
{
  static guint32 seq = 100;
  GstEvent *e = calloc (1, sizeof *e); e->type = GST_EVENT_CAPS; e->caps = gst_stub_caps_new (channels); e->seqnum = seq++;
  return pad->event (pad, pad->object.parent, e);
}
gboolean gst_stub_pad_send_eos (GstPad * pad, guint32 seqnum)
{ GstEvent *e = calloc (1, sizeof *e); e->type = GST_EVENT_EOS; e->seqnum = seqnum; return pad->event (pad, pad->object.parent, e); }
GstFlowReturn gst_stub_pad_push (GstPad * pad, const float *samples, size_t n_floats)
{
  GstBuffer *b = calloc (1, sizeof *b);
  b->size = n_floats * sizeof (float); b->data = malloc (b->size ? b->size : 1); memcpy (b->data, samples, b->size);
  return pad->chain (pad, pad->object.parent, b);
}
/* caps query on `pad` with an optional channel filter: the channel count of the answer, -1 if empty */
gint gst_stub_pad_query_caps (GstPad * pad, gint filter_channels)
{
  GstQuery q; memset (&q, 0, sizeof q); q.type = GST_QUERY_CAPS;
  q.filter = filter_channels ? gst_stub_caps_new (filter_channels) : NULL;
  gint r = -2;
  if (pad->query (pad, pad->object.parent, &q) && q.result) { r = q.result->empty ? -1 : q.result->s.channels; gst_caps_unref (q.result); }
  if (q.filter) gst_caps_unref (q.filter);
  return r;
}
GstStateChangeReturn gst_stub_change_state (GstElement * e, GstStateChange transition)
{ return ((GstElementClass *) G_OBJECT (e)->klass)->change_state (e, transition); }
const GstElementClass *gst_stub_element_class (GstElement * e) { return (GstElementClass *) G_OBJECT (e)->klass; }
