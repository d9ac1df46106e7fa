/* oracle/shim/glib/gprintf.h -- TEST INFRASTRUCTURE ONLY. */
#include <glib.h>
