// fixca_geometry.h -- the coordinate half of the pass, shared by host and device.
//
// The reference maps an output index i of channel c to a source coordinate with
//     d = (i - center) * scale_c + center - shift_c            (fix-ca.c:801, :813)
// independently per axis, so one axis of one channel is fully described by
// (center, size, scale, shift).  Everything below is IEEE binary64 with every
// operation rounded separately (no FMA): device code uses the __d*_rn
// intrinsics, host code is compiled with -ffp-contract=off.
#pragma once

#include <climits>
#include <cmath>

#if defined(__CUDACC__)
#define FIXCA_HD __host__ __device__ __forceinline__
#else
#define FIXCA_HD inline
#endif

namespace fixca {

enum { CH_RED = 0, CH_BLUE = 1 };

struct Axis {
	int    center;	// (int) lens_x or lens_y           (fix-ca.c:1033-1034)
	int    size;	// width or height
	double scale;	// max_dim / (max_dim + amount)     (fix-ca.c:1044-1045)
	double shift;	// x_/y_ blue/red                   (fix-ca.c:1105-1106 etc.)
};

struct Geometry {
	int  width, height;
	int  interp;		// 0 None, 1 Linear, 2 Cubic
	Axis x[2], y[2];	// [CH_RED], [CH_BLUE]
	bool monotone;		// both scales finite and > 0: source rows/cols never decrease
};

FIXCA_HD double mul_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
	return __dmul_rn(a, b);
#else
	return a * b;
#endif
}
FIXCA_HD double add_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
	return __dadd_rn(a, b);
#else
	return a + b;
#endif
}
FIXCA_HD double sub_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
	return __dsub_rn(a, b);
#else
	return a - b;
#endif
}

// fix-ca.c:801 / :813, evaluated left to right.
FIXCA_HD double src_coord(const Axis &a, int i)
{
	return sub_rn(add_rn(mul_rn((double)(i - a.center), a.scale), (double)a.center), a.shift);
}

// round_nearest() + the clamp of scale(): fix-ca.c:776-789, :802-808.
FIXCA_HD int nearest_index(const Axis &a, int i)
{
	const double d = src_coord(a, i);
	int j;
	if (d >= 0) {
		j = (d > (double)INT_MAX) ? INT_MAX : (int)add_rn(d, 0.5);
	} else {
		// NaN lands here too, exactly as in the reference (d >= 0 is false).
		j = (d < (double)INT_MIN) ? INT_MIN : -((int)sub_rn(0.5, d));
	}
	if (j <= 0)
		return 0;
	if (j >= a.size)
		return a.size - 1;
	return j;
}

// scale_d() then floor / fraction: fix-ca.c:811-820, :1139-1142, :1207-1210.
// Returns the base index i0 in [0, size-1]; frac = clamped coordinate - i0.
FIXCA_HD int base_index(const Axis &a, int i, double &frac)
{
	double d = src_coord(a, i);
	if (d <= 0.0)
		d = 0.0;
	else if (d >= (double)(a.size - 1))
		d = (double)(a.size - 1);
	// NaN (reference UB, rejected by the host driver) would pass through here.
	const int i0 = (int)floor(d);
	frac = sub_rn(d, (double)i0);
	return i0;
}

FIXCA_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Inclusive range of source indices output index i touches on this axis.
// None: the nearest index.  Linear: {i0, min(i0+1, n-1)} (fix-ca.c:1149-1158,
// :1170-1177).  Cubic: {max(i0-1,0) .. min(i0+2, n-1)} (fix-ca.c:1219-1256,
// :1271-1298).
FIXCA_HD void tap_range(const Axis &a, int interp, int i, int &lo, int &hi)
{
	if (interp == 0) {
		lo = hi = nearest_index(a, i);
	} else {
		double f;
		const int i0 = base_index(a, i, f);
		if (interp == 1) {
			lo = i0;
			hi = i0 + 1 < a.size ? i0 + 1 : a.size - 1;
		} else {
			lo = i0 > 0 ? i0 - 1 : 0;
			hi = i0 + 2 < a.size ? i0 + 2 : a.size - 1;
		}
	}
}

// Inclusive source range needed by outputs [i1, i2] (i1 <= i2) of one axis:
// the outputs' own positions (green/alpha are copied in place, fix-ca.c:1098)
// plus both channels' taps.  Requires a monotone geometry so that the extremes
// sit at the interval ends.
FIXCA_HD void span_needed(const Axis &red, const Axis &blue, int interp, int i1, int i2, int &lo, int &hi)
{
	int l, h;
	lo = i1;
	hi = i2;
	tap_range(red, interp, i1, l, h);
	lo = l < lo ? l : lo;
	tap_range(red, interp, i2, l, h);
	hi = h > hi ? h : hi;
	tap_range(blue, interp, i1, l, h);
	lo = l < lo ? l : lo;
	tap_range(blue, interp, i2, l, h);
	hi = h > hi ? h : hi;
}

} // namespace fixca
