/* oracle/shim/glib.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Minimal stand-in for <glib.h> so that the reference's numeric C files
 * (/root/reference/src/{earmodel,fftearmodel,fbearmodel,leveladapter,modpatt,
 * movaccum,movs,nn}.c) compile UNMODIFIED in a container without GLib.
 * Only the symbols those files use are provided.  Nothing here is part of the
 * product (gstpeaq_b200/); it exists so the reference itself can serve as the
 * parity oracle (oracle/_ref/libpeaq_ref.so).
 */
#ifndef PEAQ_ORACLE_SHIM_GLIB_H
#define PEAQ_ORACLE_SHIM_GLIB_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <alloca.h>
#include <assert.h>

#ifdef __cplusplus
#define G_BEGIN_DECLS extern "C" {
#define G_END_DECLS }
#else
#define G_BEGIN_DECLS
#define G_END_DECLS
#endif

typedef char gchar;
typedef int gint;
typedef unsigned int guint;
typedef int gboolean;
typedef float gfloat;
typedef double gdouble;
typedef void *gpointer;
typedef const void *gconstpointer;
typedef size_t gsize;
typedef uint32_t guint32;
typedef unsigned long gulong;

#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif

#define G_MAXUINT UINT_MAX

#undef MAX
#define MAX(a, b) (((a) > (b)) ? (a) : (b))
#undef MIN
#define MIN(a, b) (((a) < (b)) ? (a) : (b))
#undef ABS
#define ABS(a) (((a) < 0) ? -(a) : (a))
#undef CLAMP
#define CLAMP(x, lo, hi) (((x) > (hi)) ? (hi) : (((x) < (lo)) ? (lo) : (x)))

#define GLIB_CHECK_VERSION(a, b, c) 1

#define g_new(type, n) ((type *) malloc (sizeof (type) * ((n) > 0 ? (n) : 1)))
#define g_new0(type, n) ((type *) calloc (((n) > 0 ? (n) : 1), sizeof (type)))
#define g_renew(type, p, n) \
  ((type *) realloc ((p), sizeof (type) * ((n) > 0 ? (n) : 1)))
#define g_newa(type, n) ((type *) alloca (sizeof (type) * (n)))
#define g_free(p) free ((void *) (p))
#define g_assert(x) assert (x)
#define g_printf printf

/* GArray: only sized_new / append_val / unref and the ->data / ->len fields */
typedef struct _GArray
{
  gchar *data;
  guint len;
  guint elt_size;
  guint cap;
  gint ref;
} GArray;

static inline GArray *
g_array_sized_new (gboolean zero_terminated, gboolean clear_, guint elt_size,
                   guint reserved)
{
  GArray *a = (GArray *) calloc (1, sizeof (GArray));
  (void) zero_terminated;
  (void) clear_;
  a->elt_size = elt_size;
  a->cap = reserved > 0 ? reserved : 1;
  a->data = (gchar *) malloc ((size_t) a->cap * elt_size);
  a->ref = 1;
  return a;
}

static inline void
peaq_shim_array_append (GArray *a, const void *v)
{
  if (a->len == a->cap) {
    a->cap *= 2;
    a->data = (gchar *) realloc (a->data, (size_t) a->cap * a->elt_size);
  }
  memcpy (a->data + (size_t) a->len * a->elt_size, v, a->elt_size);
  a->len++;
}

#define g_array_append_val(a, v) peaq_shim_array_append ((a), &(v))

static inline void
g_array_unref (GArray *a)
{
  if (--a->ref == 0) {
    free (a->data);
    free (a);
  }
}

#endif
