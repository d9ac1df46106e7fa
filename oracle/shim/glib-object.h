/* oracle/shim/glib-object.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Minimal stand-in for <glib-object.h>: just enough of the GObject type
 * system (static type registration, class/instance init chains, base_init,
 * properties with G_PARAM_CONSTRUCT defaults, ref counting) for the
 * reference's numeric classes to be instantiated the way its own code does
 * (e.g. /root/reference/src/testpeaq.c:663-664, gstpeaq.c:362-363).
 */
#ifndef PEAQ_ORACLE_SHIM_GLIB_OBJECT_H
#define PEAQ_ORACLE_SHIM_GLIB_OBJECT_H

#include <glib.h>
#include <stdarg.h>

G_BEGIN_DECLS

typedef gsize GType;

typedef struct _GTypeClass
{
  GType g_type;
} GTypeClass;

typedef struct _GTypeInstance
{
  GTypeClass *g_class;
} GTypeInstance;

typedef void (*GBaseInitFunc) (gpointer g_class);
typedef void (*GBaseFinalizeFunc) (gpointer g_class);
typedef void (*GClassInitFunc) (gpointer g_class, gpointer class_data);
typedef void (*GClassFinalizeFunc) (gpointer g_class, gpointer class_data);
typedef void (*GInstanceInitFunc) (GTypeInstance *instance, gpointer g_class);

typedef struct _GTypeInfo
{
  guint class_size;
  GBaseInitFunc base_init;
  GBaseFinalizeFunc base_finalize;
  GClassInitFunc class_init;
  GClassFinalizeFunc class_finalize;
  gconstpointer class_data;
  guint instance_size;
  guint n_preallocs;
  GInstanceInitFunc instance_init;
  gconstpointer value_table;
} GTypeInfo;

typedef enum
{
  PEAQ_SHIM_VALUE_NONE,
  PEAQ_SHIM_VALUE_DOUBLE,
  PEAQ_SHIM_VALUE_UINT,
  PEAQ_SHIM_VALUE_POINTER,
  PEAQ_SHIM_VALUE_BOOLEAN
} PeaqShimValueKind;

typedef struct _GValue
{
  PeaqShimValueKind kind;
  union
  {
    gdouble v_double;
    guint v_uint;
    gpointer v_pointer;
    gboolean v_boolean;
  } data;
} GValue;

typedef enum
{
  G_PARAM_READABLE = 1 << 0,
  G_PARAM_WRITABLE = 1 << 1,
  G_PARAM_READWRITE = (1 << 0) | (1 << 1),
  G_PARAM_CONSTRUCT = 1 << 2
} GParamFlags;

typedef struct _GObject GObject;
typedef struct _GObjectClass GObjectClass;
typedef struct _GParamSpec GParamSpec;

struct _GParamSpec
{
  const gchar *name;
  PeaqShimValueKind kind;
  GValue default_value;
  guint flags;
  guint param_id;
  /* accessor functions of the class that installed the property */
  void (*owner_set) (GObject *, guint, const GValue *, GParamSpec *);
  void (*owner_get) (GObject *, guint, GValue *, GParamSpec *);
  GParamSpec *next;
};

struct _GObject
{
  GTypeInstance g_type_instance;
  guint ref_count;
};

struct _GObjectClass
{
  GTypeClass g_type_class;
  void (*set_property) (GObject *object, guint property_id,
                        const GValue *value, GParamSpec *pspec);
  void (*get_property) (GObject *object, guint property_id, GValue *value,
                        GParamSpec *pspec);
  void (*finalize) (GObject *object);
  GParamSpec *pspecs;           /* newest first; tail shared with the parent */
};

GType peaq_shim_object_get_type (void);
#define G_TYPE_OBJECT (peaq_shim_object_get_type ())

#define G_TYPE_CHECK_INSTANCE_CAST(obj, type, T) ((T *) (obj))
#define G_TYPE_CHECK_CLASS_CAST(klass, type, T) ((T *) (klass))
#define G_TYPE_CHECK_INSTANCE_TYPE(obj, type) (peaq_shim_is_a ((obj), (type)))
#define G_TYPE_CHECK_CLASS_TYPE(klass, type) (TRUE)
#define G_TYPE_INSTANCE_GET_CLASS(obj, type, T) \
  ((T *) (((GTypeInstance *) (obj))->g_class))
#define G_OBJECT(obj) ((GObject *) (obj))
#define G_OBJECT_CLASS(klass) ((GObjectClass *) (klass))
#define G_OBJECT_GET_CLASS(obj) \
  ((GObjectClass *) (((GTypeInstance *) (obj))->g_class))
#define G_OBJECT_WARN_INVALID_PROPERTY_ID(obj, id, pspec) \
  do { (void) (obj); (void) (id); (void) (pspec); } while (0)

GType g_type_register_static (GType parent, const gchar *name,
                              const GTypeInfo *info, guint flags);
gpointer g_type_class_peek (GType type);
gpointer g_type_class_peek_parent (gpointer g_class);
gboolean peaq_shim_is_a (gconstpointer instance, GType type);
#define g_type_init() do { } while (0)

gpointer g_object_new (GType type, const gchar *first_property_name, ...);
void g_object_set (gpointer object, const gchar *first_property_name, ...);
void g_object_get (gpointer object, const gchar *first_property_name, ...);
void g_object_set_property (GObject *object, const gchar *name,
                            const GValue *value);
void g_object_get_property (GObject *object, const gchar *name,
                            GValue *value);
gpointer g_object_ref (gpointer object);
void g_object_unref (gpointer object);

void g_object_class_install_property (GObjectClass *oclass, guint property_id,
                                      GParamSpec *pspec);
GParamSpec *g_param_spec_double (const gchar *name, const gchar *nick,
                                 const gchar *blurb, gdouble minimum,
                                 gdouble maximum, gdouble default_value,
                                 guint flags);
GParamSpec *g_param_spec_uint (const gchar *name, const gchar *nick,
                               const gchar *blurb, guint minimum,
                               guint maximum, guint default_value,
                               guint flags);
GParamSpec *g_param_spec_pointer (const gchar *name, const gchar *nick,
                                  const gchar *blurb, guint flags);
GParamSpec *g_param_spec_boolean (const gchar *name, const gchar *nick,
                                  const gchar *blurb, gboolean default_value,
                                  guint flags);

static inline gdouble g_value_get_double (const GValue *v) { return v->data.v_double; }
static inline guint g_value_get_uint (const GValue *v) { return v->data.v_uint; }
static inline gpointer g_value_get_pointer (const GValue *v) { return v->data.v_pointer; }
static inline gboolean g_value_get_boolean (const GValue *v) { return v->data.v_boolean; }
static inline void g_value_set_double (GValue *v, gdouble d) { v->kind = PEAQ_SHIM_VALUE_DOUBLE; v->data.v_double = d; }
static inline void g_value_set_uint (GValue *v, guint u) { v->kind = PEAQ_SHIM_VALUE_UINT; v->data.v_uint = u; }
static inline void g_value_set_pointer (GValue *v, gpointer p) { v->kind = PEAQ_SHIM_VALUE_POINTER; v->data.v_pointer = p; }
static inline void g_value_set_boolean (GValue *v, gboolean b) { v->kind = PEAQ_SHIM_VALUE_BOOLEAN; v->data.v_boolean = b; }

G_END_DECLS

#endif
