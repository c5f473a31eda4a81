/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Compiles the reference plug-in's own source, unmodified and where it lies
 * (-DFIXCA_REF_SOURCE='"/root/reference/fix-ca.c"'), the same way the
 * reference's own test does (tests/test-fix-ca.c:1-2: define TEST_FIX_CA and
 * #include the .c), and exports
 *   - ref_fix_ca_region(): the static fix_ca_region() (fix-ca.c:998-1348),
 *   - ref_run():           the real run() (fix-ca.c:197-330) on top of an
 *                          in-memory fake of the ~25 libgimp/gegl/babl calls
 *                          it makes (fix-ca.c:215-217,245,327,340,352-385),
 *   - ref_color_size():    color_size() (fix-ca.c:681-711) on a fake format.
 * The GTK dialog runs headlessly: every widget constructor returns one inert
 * object and gimp_dialog_run() answers OK, which exercises the lens reset at
 * fix-ca.c:427-428.
 *
 * Output goes to oracle/_ref/libfixca_ref.so (git-ignored; travels to the GPU
 * box).  Build: see oracle/Makefile.  No reference source is copied here.
 */
#define TEST_FIX_CA 1
#include FIXCA_REF_SOURCE

#include <stdarg.h>
#include <stdio.h>

/* ------------------------------------------------------------------ */
/* fake GIMP state                                                     */
/* ------------------------------------------------------------------ */
struct _Babl { char name[64]; int bpp; };
struct _GeglBuffer { int is_shadow; };

static struct {
	int width, height, bpp;
	struct _Babl format;
	guchar *pixels;		/* owned by the caller of ref_fake_set_drawable */
	guchar *shadow;		/* owned here */
	int sel_x, sel_y, sel_w, sel_h, has_sel;
	struct _GeglBuffer buf, shadow_buf;
	GimpDrawable drawable;
	char last_message[256];
	int n_progress_init, n_progress_update, n_messages, n_merge, n_flush;
	double last_progress;
	unsigned char saved[256];
	unsigned saved_bytes;
	int preview_x, preview_y, preview_w, preview_h;
	guchar *preview_out;	/* caller-owned capture buffer */
	int preview_rowstride;
	int dialog_response;
} G;

void *shim_g_malloc (size_t n) { void *p = malloc (n ? n : 1); if (!p) abort (); return p; }
void g_free (void *p) { free (p); }
void g_object_unref (void *obj) { (void) obj; }
unsigned long g_signal_connect (void *i, const char *s, GCallback cb, void *d)
{ (void) i; (void) s; (void) cb; (void) d; return 1; }
unsigned long g_signal_connect_swapped (void *i, const char *s, GCallback cb, void *d)
{ (void) i; (void) s; (void) cb; (void) d; return 1; }

void g_message (const char *fmt, ...)
{
	va_list ap;
	va_start (ap, fmt);
	vsnprintf (G.last_message, sizeof G.last_message, fmt, ap);
	va_end (ap);
	G.n_messages++;
}

const GeglRectangle *shim_gegl_rect (gint x, gint y, gint w, gint h)
{
	static GeglRectangle r[4];
	static int k;
	GeglRectangle *p = &r[k++ & 3];
	p->x = x; p->y = y; p->width = w; p->height = h;
	return p;
}

void gegl_init (gint *argc, gchar ***argv) { (void) argc; (void) argv; }
void gegl_exit (void) { }

/* GEGL semantics with GEGL_AUTO_ROWSTRIDE: the client array is packed with a
 * row stride of rect->width * bpp. */
void gegl_buffer_get (GeglBuffer *buf, const GeglRectangle *r, gdouble scale,
		      const Babl *format, gpointer dest, gint rowstride, gint abyss)
{
	int y;
	const guchar *from = buf->is_shadow ? G.shadow : G.pixels;
	(void) scale; (void) format; (void) rowstride; (void) abyss;
	for (y = 0; y < r->height; ++y)
		memcpy ((guchar *) dest + (size_t) y * r->width * G.bpp,
			from + ((size_t) (r->y + y) * G.width + r->x) * G.bpp,
			(size_t) r->width * G.bpp);
}

void gegl_buffer_set (GeglBuffer *buf, const GeglRectangle *r, gint level,
		      const Babl *format, const void *src, gint rowstride)
{
	int y;
	guchar *to = buf->is_shadow ? G.shadow : G.pixels;
	(void) level; (void) format; (void) rowstride;
	for (y = 0; y < r->height; ++y)
		memcpy (to + ((size_t) (r->y + y) * G.width + r->x) * G.bpp,
			(const guchar *) src + (size_t) y * r->width * G.bpp,
			(size_t) r->width * G.bpp);
}

int babl_format_get_bytes_per_pixel (const Babl *f) { return f->bpp; }
const char *babl_get_name (const Babl *f) { return f->name; }

const gchar *gimp_locale_directory (void) { return "/nonexistent"; }
gboolean gimp_install_procedure (const gchar *name, const gchar *blurb, const gchar *help,
				 const gchar *author, const gchar *copyright, const gchar *date,
				 const gchar *menu_label, const gchar *image_types,
				 GimpPDBProcType type, gint n_params, gint n_return_vals,
				 const GimpParamDef *params, const GimpParamDef *return_vals)
{
	(void) name; (void) blurb; (void) help; (void) author; (void) copyright; (void) date;
	(void) menu_label; (void) image_types; (void) type; (void) n_params;
	(void) n_return_vals; (void) params; (void) return_vals;
	return TRUE;
}
gboolean gimp_plugin_menu_register (const gchar *n, const gchar *m) { (void) n; (void) m; return TRUE; }

GimpDrawable *gimp_drawable_get (gint32 id)
{
	G.drawable.drawable_id = id;
	G.drawable.width = G.width;
	G.drawable.height = G.height;
	G.drawable.bpp = G.bpp;
	return &G.drawable;
}
void gimp_drawable_detach (GimpDrawable *d) { (void) d; }
void gimp_tile_cache_ntiles (gint n) { (void) n; }
guint gimp_tile_width (void) { return 64; }
guint gimp_tile_height (void) { return 64; }

gboolean gimp_drawable_mask_intersect (gint32 id, gint *x, gint *y, gint *w, gint *h)
{
	(void) id;
	if (G.has_sel) { *x = G.sel_x; *y = G.sel_y; *w = G.sel_w; *h = G.sel_h; }
	else { *x = 0; *y = 0; *w = G.width; *h = G.height; }
	return (*w > 0 && *h > 0);
}
const Babl *gimp_drawable_get_format (gint32 id) { (void) id; return &G.format; }
GeglBuffer *gimp_drawable_get_buffer (gint32 id) { (void) id; G.buf.is_shadow = 0; return &G.buf; }
GeglBuffer *gimp_drawable_get_shadow_buffer (gint32 id)
{ (void) id; G.shadow_buf.is_shadow = 1; return &G.shadow_buf; }
gint gimp_drawable_width (gint32 id) { (void) id; return G.width; }
gint gimp_drawable_height (gint32 id) { (void) id; return G.height; }

gboolean gimp_drawable_merge_shadow (gint32 id, gboolean undo)
{
	int x, y, w, h, r;
	(void) undo;
	gimp_drawable_mask_intersect (id, &x, &y, &w, &h);
	for (r = 0; r < h; ++r)
		memcpy (G.pixels + ((size_t) (y + r) * G.width + x) * G.bpp,
			G.shadow + ((size_t) (y + r) * G.width + x) * G.bpp, (size_t) w * G.bpp);
	G.n_merge++;
	return TRUE;
}
gboolean gimp_drawable_update (gint32 id, gint x, gint y, gint w, gint h)
{ (void) id; (void) x; (void) y; (void) w; (void) h; return TRUE; }
void gimp_displays_flush (void) { G.n_flush++; }

gboolean gimp_get_data (const gchar *key, gpointer data)
{
	(void) key;
	if (!G.saved_bytes) return FALSE;
	memcpy (data, G.saved, G.saved_bytes);
	return TRUE;
}
gboolean gimp_set_data (const gchar *key, gconstpointer data, guint32 bytes)
{
	(void) key;
	if (bytes > sizeof G.saved) abort ();
	memcpy (G.saved, data, bytes);
	G.saved_bytes = bytes;
	return TRUE;
}
gboolean gimp_progress_init (const gchar *msg) { (void) msg; G.n_progress_init++; return TRUE; }
gboolean gimp_progress_update (gdouble f) { G.n_progress_update++; G.last_progress = f; return TRUE; }
void gimp_message (const gchar *msg) { g_message ("%s", msg); }

/* libgimpcolor (GIMP 2.10, libgimpcolor/gimpcolorspace.c) is a dependency that is NOT part of the
 * reference tree; only the preview-only saturate() (fix-ca.c:922-943) reaches these two.  They are
 * RESTATED here from the published GIMP 2.10 algorithm -- PARITY UNPINNED for this pair: nothing in
 * the reference's tests or fixtures exercises them.  Everything around them (get_pixel, the s *= scale
 * clamp, set_pixel, centerline) is the reference's own compiled code. */
void gimp_rgb_to_hsv (const GimpRGB *rgb, GimpHSV *hsv)
{
	gdouble max, min, delta;
	max = rgb->r > rgb->g ? (rgb->r > rgb->b ? rgb->r : rgb->b) : (rgb->g > rgb->b ? rgb->g : rgb->b);
	min = rgb->r < rgb->g ? (rgb->r < rgb->b ? rgb->r : rgb->b) : (rgb->g < rgb->b ? rgb->g : rgb->b);
	hsv->v = max;
	delta = max - min;
	if (delta > 0.0001) {
		hsv->s = delta / max;
		if (rgb->r == max) {
			hsv->h = (rgb->g - rgb->b) / delta;
			if (hsv->h < 0.0)
				hsv->h += 6.0;
		} else if (rgb->g == max) {
			hsv->h = 2.0 + (rgb->b - rgb->r) / delta;
		} else {
			hsv->h = 4.0 + (rgb->r - rgb->g) / delta;
		}
		hsv->h /= 6.0;
	} else {
		hsv->s = 0.0;
		hsv->h = 0.0;
	}
	hsv->a = rgb->a;
}
void gimp_hsv_to_rgb (const GimpHSV *hsv, GimpRGB *rgb)
{
	gint    i;
	gdouble f, w, q, t, hue;
	if (hsv->s == 0.0) {
		rgb->r = hsv->v;
		rgb->g = hsv->v;
		rgb->b = hsv->v;
	} else {
		hue = hsv->h;
		if (hue == 1.0)
			hue = 0.0;
		hue *= 6.0;
		i = (gint) hue;
		f = hue - i;
		w = hsv->v * (1.0 - hsv->s);
		q = hsv->v * (1.0 - (hsv->s * f));
		t = hsv->v * (1.0 - (hsv->s * (1.0 - f)));
		switch (i) {
		case 0: rgb->r = hsv->v; rgb->g = t;      rgb->b = w;      break;
		case 1: rgb->r = q;      rgb->g = hsv->v; rgb->b = w;      break;
		case 2: rgb->r = w;      rgb->g = hsv->v; rgb->b = t;      break;
		case 3: rgb->r = w;      rgb->g = q;      rgb->b = hsv->v; break;
		case 4: rgb->r = t;      rgb->g = w;      rgb->b = hsv->v; break;
		case 5: rgb->r = hsv->v; rgb->g = w;      rgb->b = q;      break;
		}
	}
	rgb->a = hsv->a;
}

/* ---- inert GTK ---- */
static GtkWidget the_widget = { &the_widget };
void gimp_ui_init (const gchar *p, gboolean v) { (void) p; (void) v; }
GtkWidget *gimp_dialog_new (const gchar *t, const gchar *r, GtkWidget *p, gint f,
			    GimpHelpFunc h, const gchar *id, ...)
{ (void) t; (void) r; (void) p; (void) f; (void) h; (void) id; return &the_widget; }
gint gimp_dialog_run (GtkWidget *d) { (void) d; return G.dialog_response; }
GtkWidget *gtk_vbox_new (gboolean h, gint s) { (void) h; (void) s; return &the_widget; }
void gtk_container_set_border_width (GtkWidget *c, guint w) { (void) c; (void) w; }
void gtk_container_add (GtkWidget *c, GtkWidget *w) { (void) c; (void) w; }
void gtk_widget_show (GtkWidget *w) { (void) w; }
void gtk_widget_destroy (GtkWidget *w) { (void) w; }
void gtk_box_pack_start (GtkWidget *b, GtkWidget *c, gboolean e, gboolean f, guint p)
{ (void) b; (void) c; (void) e; (void) f; (void) p; }
GtkWidget *gtk_table_new (guint r, guint c, gboolean h) { (void) r; (void) c; (void) h; return &the_widget; }
void gtk_table_set_col_spacings (GtkWidget *t, guint s) { (void) t; (void) s; }
void gtk_table_set_row_spacings (GtkWidget *t, guint s) { (void) t; (void) s; }
GtkWidget *gimp_drawable_preview_new_from_drawable_id (gint32 id) { (void) id; return &the_widget; }
gint32 gimp_drawable_preview_get_drawable_id (GtkWidget *p) { (void) p; return 1; }
void gimp_preview_invalidate (GtkWidget *p) { (void) p; }
void gimp_preview_get_position (GtkWidget *p, gint *x, gint *y)
{ (void) p; *x = G.preview_x; *y = G.preview_y; }
void gimp_preview_get_size (GtkWidget *p, gint *w, gint *h)
{ (void) p; *w = G.preview_w; *h = G.preview_h; }
void gimp_preview_draw_buffer (GtkWidget *p, const guchar *buf, gint rowstride)
{
	(void) p;
	G.preview_rowstride = rowstride;
	if (G.preview_out)
		memcpy (G.preview_out, buf, (size_t) rowstride * G.preview_h);
}
GtkObject *gimp_scale_entry_new (GtkWidget *table, gint col, gint row, const gchar *text,
				 gint sw, gint spw, gdouble value, gdouble lower, gdouble upper,
				 gdouble step, gdouble page, guint digits, gboolean constrain,
				 gdouble ul, gdouble uu, const gchar *tip, const gchar *hid)
{
	(void) table; (void) col; (void) row; (void) text; (void) sw; (void) spw; (void) value;
	(void) lower; (void) upper; (void) step; (void) page; (void) digits; (void) constrain;
	(void) ul; (void) uu; (void) tip; (void) hid;
	return &the_widget;
}
void gimp_double_adjustment_update (GtkObject *a, gpointer d) { (void) a; (void) d; }
GtkWidget *gimp_int_combo_box_new (const gchar *l, gint v, ...) { (void) l; (void) v; return &the_widget; }
unsigned long gimp_int_combo_box_connect (GtkWidget *c, gint v, GCallback cb, gpointer d)
{ (void) c; (void) v; (void) cb; (void) d; return 1; }
gboolean gimp_int_combo_box_get_active (GtkWidget *c, gint *v) { (void) c; (void) v; return TRUE; }
GtkWidget *gimp_table_attach_aligned (GtkWidget *t, gint c, gint r, const gchar *l, gfloat xa,
				      gfloat ya, GtkWidget *w, gint cs, gboolean la)
{ (void) t; (void) c; (void) r; (void) l; (void) xa; (void) ya; (void) w; (void) cs; (void) la; return &the_widget; }
GtkWidget *gimp_frame_new (const gchar *l) { (void) l; return &the_widget; }

/* ------------------------------------------------------------------ */
/* exported entry points (ctypes / C callers)                          */
/* ------------------------------------------------------------------ */
#define EXPORT __attribute__ ((visibility ("default")))

/* p[10] = blue, red, lens_x, lens_y, interpolation, saturation, x_blue, x_red, y_blue, y_red */
static void fill_params (FixCaParams *fp, const double *p)
{
	memset (fp, 0, sizeof *fp);
	fp->blue = p[0]; fp->red = p[1]; fp->lens_x = p[2]; fp->lens_y = p[3];
	fp->update_preview = TRUE;
	fp->interpolation = (GimpInterpolationType) (int) p[4];
	fp->saturation = p[5];
	fp->x_blue = p[6]; fp->x_red = p[7]; fp->y_blue = p[8]; fp->y_red = p[9];
}

EXPORT void ref_fix_ca_region (const unsigned char *src, unsigned char *dst, int width, int height,
			       int bytes, int bpc, const double *p, int x1, int x2, int y1, int y2,
			       int show_progress)
{
	FixCaParams fp;
	fill_params (&fp, p);
	fix_ca_region ((guchar *) src, dst, width, height, bytes, bpc, &fp, x1, x2, y1, y2, show_progress);
}

EXPORT int ref_sizeof_params (void) { return (int) sizeof (FixCaParams); }

EXPORT int ref_color_size (const char *format_name, int bpp)
{
	struct _Babl f;
	snprintf (f.name, sizeof f.name, "%s", format_name);
	f.bpp = bpp;
	return color_size (&f);
}

EXPORT void ref_fake_set_drawable (int width, int height, int bpp, const char *format_name,
				   unsigned char *pixels)
{
	free (G.shadow);
	memset (&G, 0, sizeof G);
	G.width = width; G.height = height; G.bpp = bpp;
	snprintf (G.format.name, sizeof G.format.name, "%s", format_name);
	G.format.bpp = bpp;
	G.pixels = pixels;
	G.shadow = calloc ((size_t) width * height * bpp + 1, 1);
	G.dialog_response = GTK_RESPONSE_OK;
}

EXPORT void ref_fake_set_selection (int x, int y, int w, int h)
{ G.has_sel = 1; G.sel_x = x; G.sel_y = y; G.sel_w = w; G.sel_h = h; }

EXPORT void ref_fake_set_dialog_response (int ok) { G.dialog_response = ok ? GTK_RESPONSE_OK : GTK_RESPONSE_CANCEL; }

EXPORT void ref_fake_set_saved_params (const double *p)
{
	FixCaParams fp;
	fill_params (&fp, p);
	gimp_set_data (DATA_KEY_VALS, &fp, sizeof fp);
}

/* out[10] in the same order as fill_params(); returns 0 if nothing was saved. */
EXPORT int ref_fake_get_saved_params (double *out)
{
	FixCaParams fp;
	if (!gimp_get_data (DATA_KEY_VALS, &fp)) return 0;
	out[0] = fp.blue; out[1] = fp.red; out[2] = fp.lens_x; out[3] = fp.lens_y;
	out[4] = fp.interpolation; out[5] = fp.saturation;
	out[6] = fp.x_blue; out[7] = fp.x_red; out[8] = fp.y_blue; out[9] = fp.y_red;
	return 1;
}

EXPORT const char *ref_fake_last_message (void) { return G.last_message; }
EXPORT int ref_fake_counter (int which)
{
	switch (which) {
	case 0: return G.n_progress_init;
	case 1: return G.n_progress_update;
	case 2: return G.n_messages;
	case 3: return G.n_merge;
	case 4: return G.n_flush;
	default: return -1;
	}
}

/*
 * Drive the reference's real run() the way GIMP's PDB would for
 *   (Test-Fix-CA run_mode image drawable blue red lens_x lens_y interp x_blue x_red y_blue y_red)
 * nparams counts from run_mode (5..12 accepted by the reference).  FLOAT args
 * are stored through d_float exactly as libgimp marshals GIMP_PDB_FLOAT.
 * proc_name lets a test pass a wrong name.  Returns values[0].data.d_status.
 */
EXPORT int ref_run (const char *proc_name, int run_mode, int nparams, const double *f /* [8]: blue red lens_x lens_y x_blue x_red y_blue y_red */,
		    int interpolation)
{
	GimpParam param[12];
	GimpParam *ret = NULL;
	gint nret = 0;
	memset (param, 0, sizeof param);
	param[0].type = GIMP_PDB_INT32;    param[0].data.d_int32 = run_mode;
	param[1].type = GIMP_PDB_IMAGE;    param[1].data.d_image = 1;
	param[2].type = GIMP_PDB_DRAWABLE; param[2].data.d_drawable = 1;
	param[3].type = GIMP_PDB_FLOAT;    param[3].data.d_float = f[0];
	param[4].type = GIMP_PDB_FLOAT;    param[4].data.d_float = f[1];
	param[5].type = GIMP_PDB_FLOAT;    param[5].data.d_float = f[2];
	param[6].type = GIMP_PDB_FLOAT;    param[6].data.d_float = f[3];
	param[7].type = GIMP_PDB_INT8;     param[7].data.d_int8 = (guint8) interpolation;
	param[8].type = GIMP_PDB_FLOAT;    param[8].data.d_float = f[4];
	param[9].type = GIMP_PDB_FLOAT;    param[9].data.d_float = f[5];
	param[10].type = GIMP_PDB_FLOAT;   param[10].data.d_float = f[6];
	param[11].type = GIMP_PDB_FLOAT;   param[11].data.d_float = f[7];
	run (proc_name, nparams, param, &nret, &ret);
	if (nret != 1 || !ret) return -1;
	return (int) ret[0].data.d_status;
}

/* The dialog's lens reset (fix-ca.c:427-428), reached through the real dialog code. */
EXPORT void ref_dialog_lens (int width, int height, double *lens_x, double *lens_y)
{
	static unsigned char px[4];
	FixCaParams fp = fix_ca_params_default;
	ref_fake_set_drawable (width, height, 3, "R'G'B' u8", px);
	fp.lens_x = *lens_x; fp.lens_y = *lens_y;
	fix_ca_dialog (1, &fp);
	*lens_x = fp.lens_x; *lens_y = fp.lens_y;
}

EXPORT void ref_query (void) { query (); }

/* The dialog's preview refresh, preview_update() (fix-ca.c:617-679), on the drawable installed with
 * ref_fake_set_drawable(): the visible window is (x, y, w, h); `out` receives the 8-bit buffer the
 * reference hands to gimp_preview_draw_buffer() (w * h * bpp / |bpc| bytes); returns its row stride. */
EXPORT int ref_preview_update (int x, int y, int w, int h, const double *p, unsigned char *out)
{
	FixCaParams fp;
	fill_params (&fp, p);
	G.preview_x = x; G.preview_y = y; G.preview_w = w; G.preview_h = h;
	G.preview_out = out;
	G.preview_rowstride = 0;
	preview_update (&the_widget, &fp);
	G.preview_out = NULL;
	return G.preview_rowstride;
}

/* The reference's function on `nthreads` disjoint full-width row bands, one
 * thread per band: legal because fix_ca_region() only uses locals and g_new
 * (fix-ca.c:1003-1031) and a band call writes exactly the rows the full call
 * would.  Used by bench.py --impl reference as the "all host cores" arm. */
#include <pthread.h>
typedef struct {
	const unsigned char *src; unsigned char *dst;
	int width, height, bytes, bpc, y1, y2;
	FixCaParams fp;
} ref_band;

static void *ref_band_main (void *arg)
{
	ref_band *b = arg;
	fix_ca_region ((guchar *) b->src, b->dst, b->width, b->height, b->bytes, b->bpc, &b->fp,
		       0, b->width, b->y1, b->y2, TRUE);
	return NULL;
}

EXPORT void ref_fix_ca_region_mt (const unsigned char *src, unsigned char *dst, int width, int height,
				  int bytes, int bpc, const double *p, int y1, int y2, int nthreads)
{
	int i, n = nthreads < 1 ? 1 : nthreads;
	pthread_t *th;
	ref_band *bd;
	if (n > y2 - y1) n = y2 - y1 > 0 ? y2 - y1 : 1;
	th = malloc (sizeof *th * n);
	bd = malloc (sizeof *bd * n);
	for (i = 0; i < n; ++i) {
		bd[i].src = src; bd[i].dst = dst; bd[i].width = width; bd[i].height = height;
		bd[i].bytes = bytes; bd[i].bpc = bpc;
		bd[i].y1 = y1 + (int) ((long long) (y2 - y1) * i / n);
		bd[i].y2 = y1 + (int) ((long long) (y2 - y1) * (i + 1) / n);
		fill_params (&bd[i].fp, p);
		pthread_create (&th[i], NULL, ref_band_main, &bd[i]);
	}
	for (i = 0; i < n; ++i)
		pthread_join (th[i], NULL);
	free (th); free (bd);
}
