/*
 * oracle/shim/libgimp/gimp.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A minimal stand-in for <libgimp/gimp.h> (GIMP 2.10 / glib / gegl / babl) that
 * declares just enough for the reference plug-in source to compile unmodified,
 * where it lies under /root/reference, into oracle/_ref/libfixca_ref.so.
 * Nothing here is shipped in the product library.  The behaviour behind these
 * declarations is the in-memory fake backend in oracle/ref_harness.c.
 */
#ifndef FIXCA_ORACLE_SHIM_GIMP_H
#define FIXCA_ORACLE_SHIM_GIMP_H

#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <limits.h>

/* ---- glib basics ---- */
typedef int            gint;
typedef unsigned int   guint;
typedef char           gchar;
typedef unsigned char  guchar;
typedef double         gdouble;
typedef float          gfloat;
typedef int            gboolean;
typedef int32_t        gint32;
typedef uint32_t       guint32;
typedef int16_t        gint16;
typedef int8_t         gint8;
typedef uint8_t        guint8;
typedef void          *gpointer;
typedef const void    *gconstpointer;
typedef size_t         gsize;

#ifndef TRUE
#define TRUE  1
#endif
#ifndef FALSE
#define FALSE 0
#endif
#ifndef MAX
#define MAX(a, b) (((a) > (b)) ? (a) : (b))
#endif
#ifndef MIN
#define MIN(a, b) (((a) < (b)) ? (a) : (b))
#endif
#define G_N_ELEMENTS(a) (sizeof (a) / sizeof ((a)[0]))
typedef void (*GCallback) (void);
#define G_CALLBACK(f) ((GCallback) (f))

#define g_new(type, n) ((type *) shim_g_malloc (sizeof (type) * (size_t) (n)))
void *shim_g_malloc (size_t n);
void  g_free (void *p);
void  g_message (const char *fmt, ...);
void  g_object_unref (void *obj);
unsigned long g_signal_connect (void *inst, const char *sig, GCallback cb, void *data);
unsigned long g_signal_connect_swapped (void *inst, const char *sig, GCallback cb, void *data);

/* ---- enums the plug-in names ---- */
typedef enum {
	GIMP_INTERPOLATION_NONE = 0,
	GIMP_INTERPOLATION_LINEAR = 1,
	GIMP_INTERPOLATION_CUBIC = 2
} GimpInterpolationType;

typedef enum {
	GIMP_RUN_INTERACTIVE = 0,
	GIMP_RUN_NONINTERACTIVE = 1,
	GIMP_RUN_WITH_LAST_VALS = 2
} GimpRunMode;

typedef enum {
	GIMP_PDB_EXECUTION_ERROR = 0,
	GIMP_PDB_CALLING_ERROR = 1,
	GIMP_PDB_PASS_THROUGH = 2,
	GIMP_PDB_SUCCESS = 3,
	GIMP_PDB_CANCEL = 4
} GimpPDBStatusType;

typedef enum {
	GIMP_PDB_INT32 = 0,
	GIMP_PDB_INT16 = 1,
	GIMP_PDB_INT8 = 2,
	GIMP_PDB_FLOAT = 3,
	GIMP_PDB_STRING = 4,
	GIMP_PDB_IMAGE = 13,
	GIMP_PDB_DRAWABLE = 16,
	GIMP_PDB_STATUS = 21
} GimpPDBArgType;

typedef enum { GIMP_PLUGIN = 1 } GimpPDBProcType;

/* ---- PDB parameter plumbing ---- */
typedef struct {
	GimpPDBArgType type;
	gchar *name;
	gchar *description;
} GimpParamDef;

/* One storage cell read through several members: this overlap is what makes
 * the lens arguments arrive the way they do in the reference's run(). */
typedef union {
	gint32   d_int32;
	gint16   d_int16;
	guint8   d_int8;
	gdouble  d_float;
	gchar   *d_string;
	gint32   d_image;
	gint32   d_drawable;
	GimpPDBStatusType d_status;
} GimpParamData;

typedef struct {
	GimpPDBArgType type;
	GimpParamData  data;
} GimpParam;

typedef struct {
	gint32 drawable_id;
	guint  width;
	guint  height;
	guint  bpp;
} GimpDrawable;

typedef void (*GimpInitProc)  (void);
typedef void (*GimpQuitProc)  (void);
typedef void (*GimpQueryProc) (void);
typedef void (*GimpRunProc)   (const gchar *name, gint nparams, const GimpParam *param,
			       gint *nreturn_vals, GimpParam **return_vals);
typedef struct {
	GimpInitProc  init_proc;
	GimpQuitProc  quit_proc;
	GimpQueryProc query_proc;
	GimpRunProc   run_proc;
} GimpPlugInInfo;

#define MAIN()
#define GIMP_CHECK_VERSION(a, b, c) (1)

typedef struct { gdouble r, g, b, a; } GimpRGB;
typedef struct { gdouble h, s, v, a; } GimpHSV;
void gimp_rgb_to_hsv (const GimpRGB *rgb, GimpHSV *hsv);
void gimp_hsv_to_rgb (const GimpHSV *hsv, GimpRGB *rgb);

/* ---- babl / gegl ---- */
typedef struct _Babl Babl;
typedef struct _GeglBuffer GeglBuffer;
typedef struct { gint x, y, width, height; } GeglRectangle;
const GeglRectangle *shim_gegl_rect (gint x, gint y, gint w, gint h);
#define GEGL_RECTANGLE(x, y, w, h) shim_gegl_rect ((x), (y), (w), (h))
#define GEGL_AUTO_ROWSTRIDE 0
#define GEGL_ABYSS_NONE 0

void gegl_init (gint *argc, gchar ***argv);
void gegl_exit (void);
void gegl_buffer_get (GeglBuffer *buf, const GeglRectangle *rect, gdouble scale,
		      const Babl *format, gpointer dest, gint rowstride, gint abyss);
void gegl_buffer_set (GeglBuffer *buf, const GeglRectangle *rect, gint level,
		      const Babl *format, const void *src, gint rowstride);
int          babl_format_get_bytes_per_pixel (const Babl *format);
const char  *babl_get_name (const Babl *format);

/* ---- libgimp calls on the plug-in's path ---- */
const gchar  *gimp_locale_directory (void);
gboolean      gimp_install_procedure (const gchar *name, const gchar *blurb, const gchar *help,
				      const gchar *author, const gchar *copyright, const gchar *date,
				      const gchar *menu_label, const gchar *image_types,
				      GimpPDBProcType type, gint n_params, gint n_return_vals,
				      const GimpParamDef *params, const GimpParamDef *return_vals);
gboolean      gimp_plugin_menu_register (const gchar *name, const gchar *menu);
GimpDrawable *gimp_drawable_get (gint32 id);
void          gimp_drawable_detach (GimpDrawable *d);
void          gimp_tile_cache_ntiles (gint n);
guint         gimp_tile_width (void);
guint         gimp_tile_height (void);
gboolean      gimp_drawable_mask_intersect (gint32 id, gint *x, gint *y, gint *w, gint *h);
const Babl   *gimp_drawable_get_format (gint32 id);
GeglBuffer   *gimp_drawable_get_buffer (gint32 id);
GeglBuffer   *gimp_drawable_get_shadow_buffer (gint32 id);
gint          gimp_drawable_width (gint32 id);
gint          gimp_drawable_height (gint32 id);
gboolean      gimp_drawable_merge_shadow (gint32 id, gboolean undo);
gboolean      gimp_drawable_update (gint32 id, gint x, gint y, gint w, gint h);
void          gimp_displays_flush (void);
gboolean      gimp_get_data (const gchar *key, gpointer data);
gboolean      gimp_set_data (const gchar *key, gconstpointer data, guint32 bytes);
gboolean      gimp_progress_init (const gchar *msg);
gboolean      gimp_progress_update (gdouble frac);
void          gimp_message (const gchar *msg);

#endif
