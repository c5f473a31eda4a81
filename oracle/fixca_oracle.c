/*
 * oracle/fixca_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A CPU restatement, in plain C, of the reference's per-pixel correction pass
 * (fix_ca_region, /root/reference/fix-ca.c:998-1348, and its helpers :713-920).
 * It is the checker for the CUDA path: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The
 * product library never calls it and has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file (a) against
 * the reference's own golden vector tests/test1.md5 (reproduced through the
 * JPEG/BMP chain of SURVEY.md App. C) and (b) byte-for-byte against the
 * reference's own source compiled into oracle/_ref/libfixca_ref.so, over the
 * format x interpolation x parameter matrix of SURVEY.md section 7; digests of
 * those outputs are committed under tests/golden/.
 *
 * Structure differs from the reference on purpose: the source coordinate of a
 * channel is an affine map applied independently per axis (fix-ca.c:799-820),
 * so the pass is described by two 1-D tables per channel (W + H entries) and
 * evaluated from them; the reference's 120-row cache (fix-ca.c:822-862) is not
 * reproduced, its meaning is "row y of src".
 *
 * Arithmetic contract: IEEE binary64, every operation
 * rounded separately.  Build with -O2 -ffp-contract=off, no -march=native.
 */
#include <math.h>
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define EXPORT __attribute__ ((visibility ("default")))

/* p[10] = blue, red, lens_x, lens_y, interpolation, saturation, x_blue, x_red, y_blue, y_red
 * (the fields of FixCaParams, fix-ca.c:70-82, minus update_preview). */
enum { P_BLUE, P_RED, P_LENS_X, P_LENS_Y, P_INTERP, P_SAT, P_XB, P_XR, P_YB, P_YR };

typedef struct {
	int    *idx;	/* None: source index.  Linear/Cubic: floor of the clamped coordinate */
	double *frac;	/* Linear/Cubic: coordinate - floor */
} axis_table;

/* fix-ca.c:801 / :813 -- (i - center) * scale + center - shift, left to right */
static double src_coord (int i, int center, double scale, double shift)
{
	double d = (double) (i - center) * scale;
	d = d + (double) center;
	d = d - shift;
	return d;
}

/* fix-ca.c:776-789 */
static int nearest_int (double d)
{
	if (d >= 0) {
		if (d > INT_MAX)
			return INT_MAX;
		return (int) (d + 0.5);
	}
	if (d < INT_MIN)
		return INT_MIN;
	return -((int) (0.5 - d));
}

static void build_axis (axis_table *t, int n, int center, double scale, double shift, int interp)
{
	int i;
	t->idx = malloc (sizeof (int) * (size_t) (n > 0 ? n : 1));
	t->frac = malloc (sizeof (double) * (size_t) (n > 0 ? n : 1));
	for (i = 0; i < n; ++i) {
		double d = src_coord (i, center, scale, shift);
		if (interp == 0) {
			/* fix-ca.c:802-808 */
			int j = nearest_int (d);
			if (j <= 0) j = 0;
			else if (j >= n) j = n - 1;
			t->idx[i] = j;
			t->frac[i] = 0.0;
		} else {
			/* fix-ca.c:814-819, then floor / fraction as at :1139-1142, :1207-1210 */
			double f;
			if (d <= 0.0) d = 0.0;
			else if (d >= n - 1) d = n - 1;
			f = floor (d);
			t->idx[i] = (int) f;
			t->frac[i] = d - (int) f;
		}
	}
}

static void free_axis (axis_table *t) { free (t->idx); free (t->frac); }

/* fix-ca.c:713-746 */
static double decode (const unsigned char *p, int bpc)
{
	double r = 0.0;
	switch (bpc) {
	case 1: r += *p; r /= 255; break;
	case 2: { uint16_t v; memcpy (&v, p, 2); r += v; r /= 65535; break; }
	case 4: { uint32_t v; memcpy (&v, p, 4); r += v; r /= 4294967295; break; }
	case 8: { uint64_t v; long double l = 0.0; memcpy (&v, p, 8); l += v;
		  l /= 18446744073709551615UL; r = l; break; }
	case -8: { double v; memcpy (&v, p, 8); r += v; break; }
	case -4: { float v; memcpy (&v, p, 4); r += v; break; }
	/* half precision: the branch the reference keeps commented out (fix-ca.c:740-742), see patch_half.py */
	case -2: { _Float16 v; memcpy (&v, p, 2); r += v; break; }
	/* u15 (0 .. 32768 <-> [0,1]): the reference answers "TODO for another day" (fix-ca.c:694-695); the pattern
	 * of its other unsigned types with max = 32768, see patch_u15.py */
	case 15: { uint16_t v; memcpy (&v, p, 2); r += v; r /= 32768; break; }
	default: break;
	}
	return r;
}

/* fix-ca.c:873-880 then :748-774 */
static void encode (unsigned char *p, double d, int bpc)
{
	if (d <= 0.0) d = 0.0;
	else if (d >= 1.0) d = 1.0;
	switch (bpc) {
	case 1: *p = round (d * 255); break;
	case 2: { uint16_t v = round (d * 65535); memcpy (p, &v, 2); break; }
	case 4: { uint32_t v = round (d * 4294967295); memcpy (p, &v, 4); break; }
	case 8: { uint64_t v = roundl (d * 18446744073709551615UL); memcpy (p, &v, 8); break; }
	case -8: memcpy (p, &d, 8); break;
	case -4: { float v = (float) d; memcpy (p, &v, 4); break; }
	case -2: { _Float16 v = (_Float16) d; memcpy (p, &v, 2); break; }	/* fix-ca.c:768-770, commented out there */
	case 15: { uint16_t v = round (d * 32768); memcpy (p, &v, 2); break; }	/* extension, see decode() */
	default: break;
	}
}

/* fix-ca.c:905-907 / :916-918 */
static double catmull_rom (double m1, double x, double p1, double p2, double t)
{
	return (((( - m1 + 3 * x - 3 * p1 + p2 ) * t +
		  ( 2 * m1 - 5 * x + 4 * p1 - p2 ) ) * t +
				 ( - m1 + p1 ) ) * t + (x + x) ) / 2.0;
}

static int clampi (int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

typedef struct {
	const unsigned char *src;
	unsigned char *dst;
	int W, H, bytes, bpc, b, interp;
	axis_table xt[2], yt[2];	/* [0] = red (offset 0), [1] = blue (offset 2b) */
} job;

static void run_rows (const job *j, int y1, int y2)
{
	const int W = j->W, H = j->H, bytes = j->bytes, bpc = j->bpc, b = j->b;
	const size_t stride = (size_t) W * bytes;
	int x, y, c;

	for (y = y1; y < y2; ++y) {
		unsigned char *out = j->dst + stride * y;
		/* green and alpha: whole-row byte copy, fix-ca.c:1094-1098 */
		memcpy (out, j->src + stride * y, stride);

		for (c = 0; c < 2; ++c) {
			const int off = c ? 2 * b : 0;
			const axis_table *xt = &j->xt[c], *yt = &j->yt[c];
			if (j->interp == 0) {
				/* fix-ca.c:1105-1120: raw b-byte copies */
				const unsigned char *row = j->src + stride * yt->idx[y] + off;
				for (x = 0; x < W; ++x)
					memcpy (out + (size_t) x * bytes + off,
						row + (size_t) xt->idx[x] * bytes, b);
			} else if (j->interp == 1) {
				/* fix-ca.c:1135-1185, bilinear() :882-894 */
				const int r0 = yt->idx[y], r1 = (r0 == H - 1) ? r0 : r0 + 1;
				const double dy = yt->frac[y];
				const unsigned char *row0 = j->src + stride * r0 + off;
				const unsigned char *row1 = j->src + stride * r1 + off;
				for (x = 0; x < W; ++x) {
					const int c0 = xt->idx[x], c1 = (c0 == W - 1) ? c0 : c0 + 1;
					const double dx = xt->frac[x];
					const double p00 = decode (row0 + (size_t) c0 * bytes, bpc);
					const double p10 = decode (row0 + (size_t) c1 * bytes, bpc);
					const double p01 = decode (row1 + (size_t) c0 * bytes, bpc);
					const double p11 = decode (row1 + (size_t) c1 * bytes, bpc);
					const double d = (1 - dy) * (p00 + dx * (p10 - p00))
						       + dy * (p01 + dx * (p11 - p01));
					encode (out + (size_t) x * bytes + off, d, bpc);
				}
			} else {
				/* fix-ca.c:1204-1318: taps clamp to the edge (:1219-1256, :1271-1298);
				 * horizontal Catmull-Rom on 4 rows, then vertical */
				const int r = yt->idx[y];
				const double dy = yt->frac[y];
				const unsigned char *rows[4];
				int k;
				for (k = 0; k < 4; ++k)
					rows[k] = j->src + stride * clampi (r - 1 + k, 0, H - 1) + off;
				for (x = 0; x < W; ++x) {
					const int cx = xt->idx[x];
					const double dx = xt->frac[x];
					const size_t o0 = (size_t) clampi (cx - 1, 0, W - 1) * bytes;
					const size_t o1 = (size_t) cx * bytes;
					const size_t o2 = (size_t) clampi (cx + 1, 0, W - 1) * bytes;
					const size_t o3 = (size_t) clampi (cx + 2, 0, W - 1) * bytes;
					double h[4];
					for (k = 0; k < 4; ++k)
						h[k] = catmull_rom (decode (rows[k] + o0, bpc),
								    decode (rows[k] + o1, bpc),
								    decode (rows[k] + o2, bpc),
								    decode (rows[k] + o3, bpc), dx);
					encode (out + (size_t) x * bytes + off,
						catmull_rom (h[0], h[1], h[2], h[3], dy), bpc);
				}
			}
		}
	}
}

static int setup (job *j, const unsigned char *src, unsigned char *dst, int W, int H,
		  int bytes, int bpc, const double *p)
{
	int xc, yc, m, b, interp;
	double s_blue, s_red;

	b = bpc == 15 ? 2 : bpc < 0 ? -bpc : bpc;
	interp = (int) p[P_INTERP];
	if (!(bpc == 1 || bpc == 2 || bpc == 4 || bpc == 8 || bpc == -4 || bpc == -8 || bpc == -2 || bpc == 15))
		return -2;
	if (bytes != 3 * b && bytes != 4 * b)
		return -2;
	if (interp < 0 || interp > 2)
		return -3;
	if (W <= 0 || H <= 0)
		return -4;

	/* fix-ca.c:1033-1045 */
	xc = (int) p[P_LENS_X];
	yc = (int) p[P_LENS_Y];
	m = xc >= yc ? xc : yc;
	if (W - xc > m) m = W - xc;
	if (H - yc > m) m = H - yc;
	s_blue = m / (m + p[P_BLUE]);
	s_red = m / (m + p[P_RED]);

	j->src = src; j->dst = dst; j->W = W; j->H = H;
	j->bytes = bytes; j->bpc = bpc; j->b = b; j->interp = interp;
	build_axis (&j->xt[0], W, xc, s_red, p[P_XR], interp);
	build_axis (&j->xt[1], W, xc, s_blue, p[P_XB], interp);
	build_axis (&j->yt[0], H, yc, s_red, p[P_YR], interp);
	build_axis (&j->yt[1], H, yc, s_blue, p[P_YB], interp);
	return 0;
}

static void teardown (job *j)
{
	free_axis (&j->xt[0]); free_axis (&j->xt[1]);
	free_axis (&j->yt[0]); free_axis (&j->yt[1]);
}

/*
 * Same meaning as the reference's fix_ca_region() with show_progress=TRUE
 * (fix-ca.c:998-1001) on its defined domain x1 == 0, x2 == width (SURVEY.md
 * App. D #3): writes rows [y1, y2) of dst and nothing else.  Returns 0, or a
 * negative code for arguments the reference cannot take.
 */
EXPORT int fixca_oracle_region (const unsigned char *src, unsigned char *dst, int width, int height,
				int bytes, int bpc, const double *p, int x1, int x2, int y1, int y2)
{
	job j;
	int rc;
	if (x1 != 0 || x2 != width || y1 < 0 || y2 > height || y1 > y2)
		return -5;
	rc = setup (&j, src, dst, width, height, bytes, bpc, p);
	if (rc) return rc;
	run_rows (&j, y1, y2);
	teardown (&j);
	return 0;
}

typedef struct { const job *j; int y1, y2; } band;
static void *band_main (void *arg) { band *b = arg; run_rows (b->j, b->y1, b->y2); return NULL; }

/* The same pass on `nthreads` disjoint full-width row bands (bench.py's CPU
 * baseline uses it; bands are independent, fix-ca.c:1091-1329). */
EXPORT int fixca_oracle_region_mt (const unsigned char *src, unsigned char *dst, int width, int height,
				   int bytes, int bpc, const double *p, int y1, int y2, int nthreads)
{
	job j;
	int rc, i, n = nthreads < 1 ? 1 : nthreads;
	pthread_t *th;
	band *bd;
	if (y1 < 0 || y2 > height || y1 > y2)
		return -5;
	rc = setup (&j, src, dst, width, height, bytes, bpc, p);
	if (rc) return rc;
	if (n > y2 - y1) n = y2 - y1 > 0 ? y2 - y1 : 1;
	th = malloc (sizeof *th * n);
	bd = malloc (sizeof *bd * n);
	for (i = 0; i < n; ++i) {
		bd[i].j = &j;
		bd[i].y1 = y1 + (int) ((long long) (y2 - y1) * i / n);
		bd[i].y2 = y1 + (int) ((long long) (y2 - y1) * (i + 1) / n);
		pthread_create (&th[i], NULL, band_main, &bd[i]);
	}
	for (i = 0; i < n; ++i)
		pthread_join (th[i], NULL);
	free (th); free (bd);
	teardown (&j);
	return 0;
}

/* Coordinate tables alone, for tests of the host-side band/halo logic.
 * channel: 0 = red, 1 = blue; axis: 0 = x, 1 = y.  idx/frac hold n entries. */
EXPORT int fixca_oracle_axis (int width, int height, const double *p, int channel, int axis,
			      int *idx, double *frac)
{
	job j;
	static unsigned char dummy;
	const axis_table *t;
	int n, rc = setup (&j, &dummy, &dummy, width, height, 3, 1, p);
	if (rc) return rc;
	t = axis ? &j.yt[channel] : &j.xt[channel];
	n = axis ? height : width;
	memcpy (idx, t->idx, sizeof (int) * n);
	memcpy (frac, t->frac, sizeof (double) * n);
	teardown (&j);
	return 0;
}

/* The dialog's lens reset, fix-ca.c:427-428: a lens coordinate <= 0 or >= size
 * becomes round (size / 2) with integer division. */
EXPORT void fixca_oracle_resolve_lens (int width, int height, double *lens_x, double *lens_y)
{
	if (*lens_x <= 0 || *lens_x >= width) *lens_x = round (width / 2);
	if (*lens_y <= 0 || *lens_y >= height) *lens_y = round (height / 2);
}

/* ------------------------------------------------------------------------- */
/* Preview epilogue: what fix_ca_region() does to a finished row when it is   */
/* called with show_progress == FALSE (fix-ca.c:1322-1327)                    */
/* ------------------------------------------------------------------------- */

/* set_pixel() WITHOUT clip_d (fix-ca.c:748-774): saturate() and centerline() call it directly */
static void put_sample (unsigned char *p, double d, int bpc)
{
	switch (bpc) {
	case 1: *p = round (d * 255); break;
	case 2: { uint16_t v = round (d * 65535); memcpy (p, &v, 2); break; }
	case 4: { uint32_t v = round (d * 4294967295); memcpy (p, &v, 4); break; }
	case 8: { uint64_t v = roundl (d * 18446744073709551615UL); memcpy (p, &v, 8); break; }
	case -8: memcpy (p, &d, 8); break;
	case -4: { float v = (float) d; memcpy (p, &v, 4); break; }
	case -2: { _Float16 v = (_Float16) d; memcpy (p, &v, 2); break; }	/* fix-ca.c:768-770, commented out there */
	case 15: { uint16_t v = round (d * 32768); memcpy (p, &v, 2); break; }	/* extension, see decode() */
	default: break;
	}
}

/* gimp_rgb_to_hsv / gimp_hsv_to_rgb of libgimpcolor 2.10 (gimpcolorspace.c) -- a dependency that is
 * NOT in the reference tree: restated from the published algorithm, PARITY UNPINNED for this pair. */
static void rgb_to_hsv (double r, double g, double b, double *h, double *s, double *v)
{
	double max = r > g ? (r > b ? r : b) : (g > b ? g : b);
	double min = r < g ? (r < b ? r : b) : (g < b ? g : b);
	double delta = max - min;
	*v = max;
	if (delta > 0.0001) {
		*s = delta / max;
		if (r == max) {
			*h = (g - b) / delta;
			if (*h < 0.0)
				*h += 6.0;
		} else if (g == max) {
			*h = 2.0 + (b - r) / delta;
		} else {
			*h = 4.0 + (r - g) / delta;
		}
		*h /= 6.0;
	} else {
		*s = 0.0;
		*h = 0.0;
	}
}

static void hsv_to_rgb (double h, double s, double v, double *r, double *g, double *b)
{
	if (s == 0.0) {
		*r = *g = *b = v;
	} else {
		double hue = h, f, w, q, t;
		int i;
		if (hue == 1.0)
			hue = 0.0;
		hue *= 6.0;
		i = (int) hue;
		f = hue - i;
		w = v * (1.0 - s);
		q = v * (1.0 - (s * f));
		t = v * (1.0 - (s * (1.0 - f)));
		switch (i) {
		case 0: *r = v; *g = t; *b = w; break;
		case 1: *r = q; *g = v; *b = w; break;
		case 2: *r = w; *g = v; *b = t; break;
		case 3: *r = w; *g = q; *b = v; break;
		case 4: *r = t; *g = w; *b = v; break;
		case 5: *r = v; *g = w; *b = q; break;
		default: break;	/* out-of-range hue: the reference leaves rgb untouched */
		}
	}
}

/* saturate(), fix-ca.c:922-943, on one row */
static void saturate_row (unsigned char *row, int width, int bytes, int bpc, double s_scale)
{
	int b = bpc == 15 ? 2 : bpc < 0 ? -bpc : bpc, x;
	for (x = 0; x < width; ++x) {
		unsigned char *px = row + (size_t) x * bytes;
		double r = decode (px, bpc), g = decode (px + b, bpc), bl = decode (px + 2 * b, bpc);
		double h, s, v;
		rgb_to_hsv (r, g, bl, &h, &s, &v);
		s *= s_scale;
		if (s > 1.0)
			s = 1.0;
		hsv_to_rgb (h, s, v, &r, &g, &bl);
		put_sample (px, r, bpc);
		put_sample (px + b, g, bpc);
		put_sample (px + 2 * b, bl, bpc);
	}
}

static void put_rgb (unsigned char *px, double c, int b, int bpc)
{
	put_sample (px, c, bpc);
	put_sample (px + b, c, bpc);
	put_sample (px + 2 * b, c, bpc);
}

/* centerline(), fix-ca.c:945-996, on row y (x1 == 0): dashed horizontal line through the lens row,
 * dashed vertical line and the two diagonals elsewhere */
static void centerline_row (unsigned char *row, int width, int bytes, int bpc, int y, int xc, int yc)
{
	int b = bpc == 15 ? 2 : bpc < 0 ? -bpc : bpc, i, x;
	double c = 1.0;
	if (y == yc) {
		i = (xc < 0 ? -xc : xc) % 16;
		if (i < 8) c = 0.0;
		for (x = 0; x < width; ++x) {
			put_rgb (row + (size_t) x * bytes, c, b, bpc);
			if (i-- < 0) {
				i = 7;
				c = c > 0 ? 0.0 : 1.0;
			}
		}
		return;
	}
	y = y <= yc ? yc - y : y - yc;
	i = (y < 0 ? -y : y) % 16;
	if (i < 8) c = 0.0;
	if (xc >= 0 && xc < width)
		put_rgb (row + (size_t) xc * bytes, c, b, bpc);
	x = xc - y;
	if (x >= 0 && x < width)
		put_rgb (row + (size_t) x * bytes, c, b, bpc);
	x = xc + y;
	if (x >= 0 && x < width)
		put_rgb (row + (size_t) x * bytes, c, b, bpc);
}

/* fix_ca_region (..., show_progress = FALSE): the pass, then per row saturate (iff saturation != 0)
 * and centerline (fix-ca.c:1322-1327). */
EXPORT int fixca_oracle_region_preview (const unsigned char *src, unsigned char *dst, int width, int height,
					int bytes, int bpc, const double *p, int y1, int y2)
{
	int y, rc = fixca_oracle_region (src, dst, width, height, bytes, bpc, p, 0, width, y1, y2);
	if (rc) return rc;
	for (y = y1; y < y2; ++y) {
		unsigned char *row = dst + (size_t) y * width * bytes;
		if (p[P_SAT] != 0.0)
			saturate_row (row, width, bytes, bpc, 1 + p[P_SAT] / 100);
		centerline_row (row, width, bytes, bpc, y, (int) p[P_LENS_X], (int) p[P_LENS_Y]);
	}
	return 0;
}
