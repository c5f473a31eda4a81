// kernels_preview.cu -- the preview epilogue of fix_ca_region(): what the reference does to every
// finished row when it is called with show_progress == FALSE (fix-ca.c:1322-1327, the dialog's
// preview call at :656-657):
//
//   saturate()    fix-ca.c:922-943   RGB -> HSV, s *= 1 + saturation/100 (clamped to 1), HSV -> RGB, iff saturation != 0
//   centerline()  fix-ca.c:945-996   dashed black/white lines through the lens centre: the lens row, the lens
//                                    column and the two diagonals
//
// Both work on normalised doubles through get_pixel()/set_pixel() (fix-ca.c:713-774; set_pixel without
// clip_d).  gimp_rgb_to_hsv / gimp_hsv_to_rgb belong to libgimpcolor (GIMP 2.10, gimpcolorspace.c), which
// is not part of the reference tree: they are restated from the published algorithm (parity unpinned for
// that pair; oracle/ref_harness.c carries the same restatement under the reference's own saturate()).
//
// This translation unit is compiled with --fmad=false: every FP64 operation is rounded separately, as in
// the reference's x86-64 build.  One thread per pixel; the preview band is small and interactive, so this
// kernel is about exactness, not speed.
#include <cstdint>

#include "fixca_internal.h"

namespace fixca {

namespace {

template <class S> struct Norm;
template <> struct Norm<uint8_t> {
	__device__ static double get(uint8_t v) { return (double)v / 255.0; }
	__device__ static uint8_t put(double d) { return (uint8_t)(uint32_t)__double2int_rz(round(d * 255.0)); }
};
template <> struct Norm<uint16_t> {
	__device__ static double get(uint16_t v) { return (double)v / 65535.0; }
	__device__ static uint16_t put(double d) { return (uint16_t)(uint32_t)__double2int_rz(round(d * 65535.0)); }
};
template <> struct Norm<u15_t> {	// extension, bpc = 15: the pattern of the unsigned types with max = 32768
	__device__ static double get(u15_t v) { return (double)v.v / 32768.0; }
	__device__ static u15_t put(double d) { u15_t o; o.v = (uint16_t)(uint32_t)__double2int_rz(round(d * 32768.0)); return o; }
};
template <> struct Norm<uint32_t> {
	__device__ static double get(uint32_t v) { return (double)v / 4294967295.0; }
	__device__ static uint32_t put(double d) { return (uint32_t)__double2ll_rz(round(d * 4294967295.0)); }
};
template <> struct Norm<uint64_t> {	// the x87 steps of get_pixel / set_pixel restated (fixca_kernels.cuh)
	__device__ static double get(uint64_t v) { return u64_get_pixel(v); }
	__device__ static uint64_t put(double d) { return u64_set_pixel(d); }
};
template <> struct Norm<float> {
	__device__ static double get(float v) { return (double)v; }
	__device__ static float put(double d) { return (float)d; }
};
template <> struct Norm<__half> {	// the commented-out half branches of get_pixel / set_pixel (fix-ca.c:740-742, :768-770)
	__device__ static double get(__half v) { return (double)__half2float(v); }
	__device__ static __half put(double d) { return __double2half(d); }
};
template <> struct Norm<double> {
	__device__ static double get(double v) { return v; }
	__device__ static double put(double d) { return d; }
};

// libgimpcolor 2.10 gimp_rgb_to_hsv (restated)
__device__ void rgb_to_hsv(double r, double g, double b, double &h, double &s, double &v)
{
	const double mx = r > g ? (r > b ? r : b) : (g > b ? g : b);
	const double mn = r < g ? (r < b ? r : b) : (g < b ? g : b);
	const double delta = mx - mn;
	v = mx;
	if (delta > 0.0001) {
		s = delta / mx;
		if (r == mx) {
			h = (g - b) / delta;
			if (h < 0.0)
				h += 6.0;
		} else if (g == mx) {
			h = 2.0 + (b - r) / delta;
		} else {
			h = 4.0 + (r - g) / delta;
		}
		h /= 6.0;
	} else {
		s = 0.0;
		h = 0.0;
	}
}

// libgimpcolor 2.10 gimp_hsv_to_rgb (restated)
__device__ void hsv_to_rgb(double h, double s, double v, double &r, double &g, double &b)
{
	if (s == 0.0) {
		r = g = b = v;
		return;
	}
	double hue = h;
	if (hue == 1.0)
		hue = 0.0;
	hue *= 6.0;
	const int i = (int)hue;
	const double f = hue - i;
	const double w = v * (1.0 - s);
	const double q = v * (1.0 - (s * f));
	const double t = v * (1.0 - (s * (1.0 - f)));
	switch (i) {
	case 0: r = v; g = t; b = w; break;
	case 1: r = q; g = v; b = w; break;
	case 2: r = w; g = v; b = t; break;
	case 3: r = w; g = q; b = v; break;
	case 4: r = t; g = w; b = v; break;
	case 5: r = v; g = w; b = q; break;
	default: break;
	}
}

struct PreviewArgs {
	unsigned char *dst;	// row dst_row0
	long long pitch;
	int dst_row0, y1, y2, width;
	int xc, yc;		// (int) lens_x, (int) lens_y  (fix-ca.c:1033-1034)
	int sat_on;		// saturation != 0
	double s_scale;		// 1 + saturation / 100
};

template <class S, int NCH>
__global__ void __launch_bounds__(256) preview_kernel(const PreviewArgs a)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	const int y = a.y1 + blockIdx.y;
	if (x >= a.width || y >= a.y2)
		return;
	S *px = reinterpret_cast<S *>(a.dst + (long long)(y - a.dst_row0) * a.pitch) + (size_t)x * NCH;

	// centerline(): is this pixel on one of the dashed lines, and in which colour?
	bool line = false;
	double c = 1.0;
	if (y == a.yc) {
		// the loop of fix-ca.c:951-966 in closed form: the first i0 + 2 pixels keep the initial colour,
		// then it alternates every 9 pixels
		const int i0 = (a.xc < 0 ? -a.xc : a.xc) % 16;
		const int toggles = x <= i0 + 1 ? 0 : (x - (i0 + 2)) / 9 + 1;
		const bool dark0 = i0 < 8;
		c = (dark0 != ((toggles & 1) != 0)) ? 0.0 : 1.0;
		line = true;
	} else {
		const int dy = y <= a.yc ? a.yc - y : y - a.yc;
		if (x == a.xc || x == a.xc - dy || x == a.xc + dy) {
			line = true;
			c = (dy % 16) < 8 ? 0.0 : 1.0;
		}
	}
	if (line) {
		px[0] = Norm<S>::put(c);
		px[1] = Norm<S>::put(c);
		px[2] = Norm<S>::put(c);
		return;
	}
	if (a.sat_on) {
		double r = Norm<S>::get(px[0]), g = Norm<S>::get(px[1]), b = Norm<S>::get(px[2]);
		double h, s, v;
		rgb_to_hsv(r, g, b, h, s, v);
		s *= a.s_scale;
		if (s > 1.0)
			s = 1.0;
		hsv_to_rgb(h, s, v, r, g, b);
		px[0] = Norm<S>::put(r);
		px[1] = Norm<S>::put(g);
		px[2] = Norm<S>::put(b);
	}
}

template <class S>
cudaError_t launch_s(int nch, const PreviewArgs &a, cudaStream_t st)
{
	for (int y = a.y1; y < a.y2; y += 32768) {
		PreviewArgs b = a;
		b.y1 = y;
		b.y2 = y + 32768 < a.y2 ? y + 32768 : a.y2;
		const dim3 grid((a.width + 255) / 256, b.y2 - b.y1), block(256);
		if (nch == 3)
			preview_kernel<S, 3><<<grid, block, 0, st>>>(b);
		else
			preview_kernel<S, 4><<<grid, block, 0, st>>>(b);
	}
	return cudaGetLastError();
}

// preview_update()'s down-conversion (fix-ca.c:659-671): every sample of the preview window goes through
// get_pixel() and set_pixel(..., 1), i.e. round(v / max * 255) stored into a guchar (no clip_d: pass-through
// samples of float images may lie outside [0,1], and None copies anything).  The store is C's double -> unsigned
// char conversion as x86-64 compiles it: truncate to a 32-bit int (out of range / NaN -> INT_MIN), keep the low byte.
struct To8Args {
	const unsigned char *src;	// row src_row0 of the corrected rows
	long long pitch;
	int src_row0, y1, y2;		// rows [y1, y2)
	int x, samples;			// first column of the window, samples (pixels * channels) per window row
	unsigned char *out;		// row y1 of the window, tight rows of `samples` bytes
	int bpp;
};

template <class S>
__global__ void __launch_bounds__(256) to8_kernel(const To8Args a)
{
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	const int y = a.y1 + blockIdx.y;
	if (j >= a.samples || y >= a.y2)
		return;
	const S *row = reinterpret_cast<const S *>(a.src + (long long)(y - a.src_row0) * a.pitch + (long long)a.x * a.bpp);
	const double v = round(Norm<S>::get(row[j]) * 255.0);
	const int iv = (v >= -2147483648.0 && v < 2147483648.0) ? (int)v : (int)0x80000000;
	a.out[(long long)(y - a.y1) * a.samples + j] = (unsigned char)iv;
}
template <>
__global__ void __launch_bounds__(256) to8_kernel<uint8_t>(const To8Args a)	// b == 1: a memcpy in the reference (:661-664)
{
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	const int y = a.y1 + blockIdx.y;
	if (j >= a.samples || y >= a.y2)
		return;
	a.out[(long long)(y - a.y1) * a.samples + j] = a.src[(long long)(y - a.src_row0) * a.pitch + (long long)a.x * a.bpp + j];
}

template <class S>
cudaError_t launch_to8_s(To8Args a, cudaStream_t st)
{
	for (int y = a.y1; y < a.y2; y += 32768) {
		To8Args b = a;
		b.y1 = y;
		b.y2 = y + 32768 < a.y2 ? y + 32768 : a.y2;
		b.out = a.out + (long long)(y - a.y1) * a.samples;
		to8_kernel<S><<<dim3((a.samples + 255) / 256, b.y2 - b.y1), 256, 0, st>>>(b);
	}
	return cudaGetLastError();
}

} // namespace

// The 8-bit preview buffer of window columns [x, x + pw) of corrected rows [y1, y2) (fix-ca.c:659-671).
cudaError_t launch_to8(int kind, int nch, const unsigned char *src, long long pitch, int src_row0, int y1, int y2, int x, int pw,
		       int bpp, unsigned char *out, cudaStream_t st)
{
	To8Args a;
	a.src = src; a.pitch = pitch; a.src_row0 = src_row0; a.y1 = y1; a.y2 = y2; a.x = x; a.samples = pw * nch; a.out = out; a.bpp = bpp;
	switch (kind) {
	case SK_U8:  return launch_to8_s<uint8_t>(a, st);
	case SK_U16: return launch_to8_s<uint16_t>(a, st);
	case SK_U32: return launch_to8_s<uint32_t>(a, st);
	case SK_F32: return launch_to8_s<float>(a, st);
	case SK_F64: return launch_to8_s<double>(a, st);
	case SK_F16: return launch_to8_s<__half>(a, st);
	case SK_U15: return launch_to8_s<u15_t>(a, st);
	default: return cudaErrorInvalidValue;	// u64: 80-bit long double in the reference
	}
}

// Applies the preview epilogue to destination rows [y1, y2) in place.  kind: SampleKind.
cudaError_t launch_preview(int kind, int nch, unsigned char *dst, long long pitch, int dst_row0, int y1, int y2, int width,
			   int xc, int yc, double saturation, cudaStream_t st)
{
	PreviewArgs a;
	a.dst = dst; a.pitch = pitch; a.dst_row0 = dst_row0; a.y1 = y1; a.y2 = y2; a.width = width;
	a.xc = xc; a.yc = yc;
	a.sat_on = saturation != 0.0;
	a.s_scale = 1 + saturation / 100;
	switch (kind) {
	case SK_U8:  return launch_s<uint8_t>(nch, a, st);
	case SK_U16: return launch_s<uint16_t>(nch, a, st);
	case SK_U32: return launch_s<uint32_t>(nch, a, st);
	case SK_U64: return launch_s<uint64_t>(nch, a, st);
	case SK_F32: return launch_s<float>(nch, a, st);
	case SK_F64: return launch_s<double>(nch, a, st);
	case SK_F16: return launch_s<__half>(nch, a, st);
	case SK_U15: return launch_s<u15_t>(nch, a, st);
	default: return cudaErrorInvalidValue;
	}
}

} // namespace fixca
