/*
 * oracle/shim/libgimp/gimpui.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Stand-in for <libgimp/gimpui.h> + GTK2: every widget is the same inert
 * object, so the reference's dialog code runs headlessly (all sliders keep
 * their initial values, the dialog "returns OK").  See oracle/ref_harness.c.
 */
#ifndef FIXCA_ORACLE_SHIM_GIMPUI_H
#define FIXCA_ORACLE_SHIM_GIMPUI_H

#include "gimp.h"

typedef struct _GtkWidget GtkWidget;
struct _GtkWidget { GtkWidget *vbox; };
typedef GtkWidget GtkObject;
typedef GtkWidget GimpPreview;
typedef GtkWidget GimpDrawablePreview;
typedef void (*GimpHelpFunc) (const gchar *help_id, gpointer help_data);

#define GTK_CONTAINER(w)        (w)
#define GTK_DIALOG(w)           (w)
#define GTK_BOX(w)              (w)
#define GTK_TABLE(w)            (w)
#define GIMP_DIALOG(w)          (w)
#define GIMP_INT_COMBO_BOX(w)   (w)
#define GIMP_DRAWABLE_PREVIEW(w) (w)
#define GIMP_PREVIEW(w)         (w)

#define GTK_STOCK_CANCEL   "gtk-cancel"
#define GTK_STOCK_OK       "gtk-ok"
#define GTK_RESPONSE_OK     (-5)
#define GTK_RESPONSE_CANCEL (-6)

void       gimp_ui_init (const gchar *prog, gboolean preview);
GtkWidget *gimp_dialog_new (const gchar *title, const gchar *role, GtkWidget *parent, gint flags,
			    GimpHelpFunc help, const gchar *help_id, ...);
gint       gimp_dialog_run (GtkWidget *dialog);
GtkWidget *gtk_vbox_new (gboolean homogeneous, gint spacing);
void       gtk_container_set_border_width (GtkWidget *c, guint w);
void       gtk_container_add (GtkWidget *c, GtkWidget *w);
void       gtk_widget_show (GtkWidget *w);
void       gtk_widget_destroy (GtkWidget *w);
void       gtk_box_pack_start (GtkWidget *box, GtkWidget *child, gboolean e, gboolean f, guint pad);
GtkWidget *gtk_table_new (guint rows, guint cols, gboolean homogeneous);
void       gtk_table_set_col_spacings (GtkWidget *t, guint s);
void       gtk_table_set_row_spacings (GtkWidget *t, guint s);
GtkWidget *gimp_drawable_preview_new_from_drawable_id (gint32 id);
gint32     gimp_drawable_preview_get_drawable_id (GtkWidget *p);
void       gimp_preview_invalidate (GtkWidget *p);
void       gimp_preview_get_position (GtkWidget *p, gint *x, gint *y);
void       gimp_preview_get_size (GtkWidget *p, gint *w, gint *h);
void       gimp_preview_draw_buffer (GtkWidget *p, const guchar *buf, gint rowstride);
GtkObject *gimp_scale_entry_new (GtkWidget *table, gint col, gint row, const gchar *text,
				 gint scale_width, gint spin_width, gdouble value,
				 gdouble lower, gdouble upper, gdouble step, gdouble page,
				 guint digits, gboolean constrain, gdouble ulower, gdouble uupper,
				 const gchar *tooltip, const gchar *help_id);
void       gimp_double_adjustment_update (GtkObject *adj, gpointer data);
GtkWidget *gimp_int_combo_box_new (const gchar *first_label, gint first_value, ...);
unsigned long gimp_int_combo_box_connect (GtkWidget *combo, gint value, GCallback cb, gpointer data);
gboolean   gimp_int_combo_box_get_active (GtkWidget *combo, gint *value);
GtkWidget *gimp_table_attach_aligned (GtkWidget *table, gint col, gint row, const gchar *label,
				      gfloat xalign, gfloat yalign, GtkWidget *widget,
				      gint colspan, gboolean left_align);
GtkWidget *gimp_frame_new (const gchar *label);

#endif
