/* oracle/shim/shim_gobject.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Implementation of the GObject subset declared in glib-object.h.  Semantics
 * that the reference relies on and that are therefore reproduced:
 *  - a new class starts as a byte copy of its parent's class struct, then ALL
 *    ancestors' base_init run root-first, then the type's own class_init
 *    (the Hann window of PeaqFFTEarModel is built in base_init,
 *    /root/reference/src/fftearmodel.c:159-173);
 *  - instance_init functions run root-first, then every G_PARAM_CONSTRUCT
 *    property is set to its default, parent class properties first
 *    (playback-level 92 dB then number-of-bands 109,
 *    /root/reference/src/earmodel.c:100-108, fftearmodel.c:207-214);
 *  - a property write is dispatched to the set_property of the class that
 *    INSTALLED the property, so "band-centers" reaches earmodel.c:539-564
 *    even when set on a subclass instance that overrides set_property.
 */
#include <glib-object.h>

typedef struct _TypeNode
{
  struct _TypeNode *parent;
  const gchar *name;
  GTypeInfo info;
  GTypeClass *klass;
} TypeNode;

static TypeNode *
node_of (GType t)
{
  return (TypeNode *) t;
}

static void
object_base_finalize (GObject *obj)
{
  (void) obj;
}

GType
peaq_shim_object_get_type (void)
{
  static TypeNode *node = NULL;
  if (!node) {
    node = (TypeNode *) calloc (1, sizeof (TypeNode));
    node->name = "GObject";
    node->info.class_size = sizeof (GObjectClass);
    node->info.instance_size = sizeof (GObject);
    GObjectClass *k = (GObjectClass *) calloc (1, sizeof (GObjectClass));
    k->g_type_class.g_type = (GType) node;
    k->finalize = object_base_finalize;
    node->klass = (GTypeClass *) k;
  }
  return (GType) node;
}

GType
g_type_register_static (GType parent, const gchar *name,
                        const GTypeInfo *info, guint flags)
{
  (void) flags;
  TypeNode *node = (TypeNode *) calloc (1, sizeof (TypeNode));
  node->parent = node_of (parent);
  node->name = name;
  node->info = *info;
  return (GType) node;
}

static void
run_base_inits (TypeNode *node, gpointer klass)
{
  if (!node)
    return;
  run_base_inits (node->parent, klass);
  if (node->info.base_init)
    node->info.base_init (klass);
}

gpointer
g_type_class_peek (GType type)
{
  TypeNode *node = node_of (type);
  if (!node->klass) {
    GTypeClass *pk = (GTypeClass *) g_type_class_peek ((GType) node->parent);
    GTypeClass *k = (GTypeClass *) calloc (1, node->info.class_size);
    memcpy (k, pk, node->parent->info.class_size);
    k->g_type = type;
    node->klass = k;
    run_base_inits (node, k);
    if (node->info.class_init)
      node->info.class_init (k, (gpointer) node->info.class_data);
  }
  return node->klass;
}

gpointer
g_type_class_peek_parent (gpointer g_class)
{
  TypeNode *node = node_of (((GTypeClass *) g_class)->g_type);
  return g_type_class_peek ((GType) node->parent);
}

gboolean
peaq_shim_is_a (gconstpointer instance, GType type)
{
  const GTypeInstance *inst = (const GTypeInstance *) instance;
  TypeNode *n = node_of (inst->g_class->g_type);
  while (n) {
    if ((GType) n == type)
      return TRUE;
    n = n->parent;
  }
  return FALSE;
}

static GParamSpec *
new_pspec (const gchar *name, PeaqShimValueKind kind, guint flags)
{
  GParamSpec *p = (GParamSpec *) calloc (1, sizeof (GParamSpec));
  p->name = name;
  p->kind = kind;
  p->flags = flags;
  p->default_value.kind = kind;
  return p;
}

GParamSpec *
g_param_spec_double (const gchar *name, const gchar *nick, const gchar *blurb,
                     gdouble minimum, gdouble maximum, gdouble default_value,
                     guint flags)
{
  (void) nick; (void) blurb; (void) minimum; (void) maximum;
  GParamSpec *p = new_pspec (name, PEAQ_SHIM_VALUE_DOUBLE, flags);
  p->default_value.data.v_double = default_value;
  return p;
}

GParamSpec *
g_param_spec_uint (const gchar *name, const gchar *nick, const gchar *blurb,
                   guint minimum, guint maximum, guint default_value,
                   guint flags)
{
  (void) nick; (void) blurb; (void) minimum; (void) maximum;
  GParamSpec *p = new_pspec (name, PEAQ_SHIM_VALUE_UINT, flags);
  p->default_value.data.v_uint = default_value;
  return p;
}

GParamSpec *
g_param_spec_pointer (const gchar *name, const gchar *nick,
                      const gchar *blurb, guint flags)
{
  (void) nick; (void) blurb;
  return new_pspec (name, PEAQ_SHIM_VALUE_POINTER, flags);
}

GParamSpec *
g_param_spec_boolean (const gchar *name, const gchar *nick,
                      const gchar *blurb, gboolean default_value, guint flags)
{
  (void) nick; (void) blurb;
  GParamSpec *p = new_pspec (name, PEAQ_SHIM_VALUE_BOOLEAN, flags);
  p->default_value.data.v_boolean = default_value;
  return p;
}

void
g_object_class_install_property (GObjectClass *oclass, guint property_id,
                                 GParamSpec *pspec)
{
  pspec->param_id = property_id;
  pspec->owner_set = oclass->set_property;
  pspec->owner_get = oclass->get_property;
  pspec->next = oclass->pspecs;
  oclass->pspecs = pspec;
}

static int
same_name (const gchar *a, const gchar *b)
{
  /* GObject treats '-' and '_' in property names as equivalent */
  for (; *a && *b; a++, b++) {
    gchar ca = *a == '_' ? '-' : *a;
    gchar cb = *b == '_' ? '-' : *b;
    if (ca != cb)
      return 0;
  }
  return *a == *b;
}

static GParamSpec *
find_pspec (GObject *obj, const gchar *name)
{
  GParamSpec *p = G_OBJECT_GET_CLASS (obj)->pspecs;
  for (; p; p = p->next)
    if (same_name (p->name, name))
      return p;
  fprintf (stderr, "peaq shim: no property '%s'\n", name);
  abort ();
  return NULL;
}

void
g_object_set_property (GObject *object, const gchar *name,
                       const GValue *value)
{
  GParamSpec *p = find_pspec (object, name);
  p->owner_set (object, p->param_id, value, p);
}

void
g_object_get_property (GObject *object, const gchar *name, GValue *value)
{
  GParamSpec *p = find_pspec (object, name);
  value->kind = p->kind;
  p->owner_get (object, p->param_id, value, p);
}

static void
set_valist (GObject *obj, const gchar *name, va_list ap)
{
  while (name) {
    GParamSpec *p = find_pspec (obj, name);
    GValue v;
    v.kind = p->kind;
    switch (p->kind) {
      case PEAQ_SHIM_VALUE_DOUBLE:
        v.data.v_double = va_arg (ap, gdouble);
        break;
      case PEAQ_SHIM_VALUE_UINT:
        v.data.v_uint = va_arg (ap, guint);
        break;
      case PEAQ_SHIM_VALUE_POINTER:
        v.data.v_pointer = va_arg (ap, gpointer);
        break;
      case PEAQ_SHIM_VALUE_BOOLEAN:
        v.data.v_boolean = va_arg (ap, gboolean);
        break;
      default:
        abort ();
    }
    p->owner_set (obj, p->param_id, &v, p);
    name = va_arg (ap, const gchar *);
  }
}

void
g_object_set (gpointer object, const gchar *first_property_name, ...)
{
  va_list ap;
  va_start (ap, first_property_name);
  set_valist ((GObject *) object, first_property_name, ap);
  va_end (ap);
}

void
g_object_get (gpointer object, const gchar *first_property_name, ...)
{
  va_list ap;
  const gchar *name = first_property_name;
  va_start (ap, first_property_name);
  while (name) {
    GParamSpec *p = find_pspec ((GObject *) object, name);
    GValue v;
    memset (&v, 0, sizeof v);
    v.kind = p->kind;
    p->owner_get ((GObject *) object, p->param_id, &v, p);
    void *dst = va_arg (ap, void *);
    switch (p->kind) {
      case PEAQ_SHIM_VALUE_DOUBLE:
        *(gdouble *) dst = v.data.v_double;
        break;
      case PEAQ_SHIM_VALUE_UINT:
        *(guint *) dst = v.data.v_uint;
        break;
      case PEAQ_SHIM_VALUE_POINTER:
        *(gpointer *) dst = v.data.v_pointer;
        break;
      case PEAQ_SHIM_VALUE_BOOLEAN:
        *(gboolean *) dst = v.data.v_boolean;
        break;
      default:
        abort ();
    }
    name = va_arg (ap, const gchar *);
  }
  va_end (ap);
}

static void
run_instance_inits (TypeNode *node, GTypeInstance *inst, gpointer klass)
{
  if (!node)
    return;
  run_instance_inits (node->parent, inst, klass);
  if (node->info.instance_init)
    node->info.instance_init (inst, klass);
}

static void
apply_construct_defaults (GObject *obj, GParamSpec *p)
{
  if (!p)
    return;
  /* list is newest-first; recurse so that parent-class properties go first */
  apply_construct_defaults (obj, p->next);
  if (p->flags & G_PARAM_CONSTRUCT)
    p->owner_set (obj, p->param_id, &p->default_value, p);
}

gpointer
g_object_new (GType type, const gchar *first_property_name, ...)
{
  TypeNode *node = node_of (type);
  GTypeClass *klass = (GTypeClass *) g_type_class_peek (type);
  GObject *obj = (GObject *) calloc (1, node->info.instance_size);
  obj->g_type_instance.g_class = klass;
  obj->ref_count = 1;
  run_instance_inits (node, &obj->g_type_instance, klass);
  apply_construct_defaults (obj, ((GObjectClass *) klass)->pspecs);
  if (first_property_name) {
    va_list ap;
    va_start (ap, first_property_name);
    set_valist (obj, first_property_name, ap);
    va_end (ap);
  }
  return obj;
}

gpointer
g_object_ref (gpointer object)
{
  ((GObject *) object)->ref_count++;
  return object;
}

void
g_object_unref (gpointer object)
{
  GObject *obj = (GObject *) object;
  if (--obj->ref_count == 0) {
    GObjectClass *k = G_OBJECT_GET_CLASS (obj);
    if (k->finalize)
      k->finalize (obj);
    free (obj);
  }
}
