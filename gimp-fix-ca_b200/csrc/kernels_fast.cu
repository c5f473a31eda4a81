// kernels_fast.cu -- Linear / Cubic in FP32 on raw sample values: within +-1 LSB
// of the reference for u8 / u16, ~2 ulp(1.0) for float.
//
//   stream_kernel (fixca_stream.cuh) the measured path: persistent column strips, TMA ring
//   strip_kernel  (fixca_strip.cuh)  per-tile fallback when the ring does not fit (same bytes)
//   direct_kernel (FastF32)          per-pixel gather, any geometry
#include <cstdlib>
#include <cstring>
#include "fixca_internal.h"

namespace fixca {

#define DIRECT_ENTRIES(S, TAG)                                                                                \
	{ (kernel_fn)direct_kernel<S, 3, 1, FastF32>, "direct/linear/f32/" TAG "x3", 0, 0, (int)sizeof(S), 0 },  \
	{ (kernel_fn)direct_kernel<S, 4, 1, FastF32>, "direct/linear/f32/" TAG "x4", 0, 0, (int)sizeof(S), 0 },  \
	{ (kernel_fn)direct_kernel<S, 3, 2, FastF32>, "direct/cubic/f32/" TAG "x3", 0, 0, (int)sizeof(S), 0 },   \
	{ (kernel_fn)direct_kernel<S, 4, 2, FastF32>, "direct/cubic/f32/" TAG "x4", 0, 0, (int)sizeof(S), 0 }

static const KernelEntry direct_table[] = {
	DIRECT_ENTRIES(uint8_t, "u8"),
	DIRECT_ENTRIES(uint16_t, "u16"),
	DIRECT_ENTRIES(float, "f32"),
	DIRECT_ENTRIES(__half, "f16"),
	DIRECT_ENTRIES(u15_t, "u15"),
};

// Columns per thread are picked so that the lane stride in shared memory, P * bytes-per-pixel,
// is conflict-free (an odd number of 32-bit words) or as close as the format allows.
#define STRIP_ENTRIES(S, TAG, P3, TW3, P4, TW4)                                                               \
	{ (kernel_fn)strip_kernel<S, 3, 1, P3, TW3>, "strip/linear/f32/" TAG "x3", TW3, 16, (int)sizeof(S), P3 }, \
	{ (kernel_fn)strip_kernel<S, 4, 1, P4, TW4>, "strip/linear/f32/" TAG "x4", TW4, 16, (int)sizeof(S), P4 }, \
	{ (kernel_fn)strip_kernel<S, 3, 2, P3, TW3>, "strip/cubic/f32/" TAG "x3", TW3, 16, (int)sizeof(S), P3 },  \
	{ (kernel_fn)strip_kernel<S, 4, 2, P4, TW4>, "strip/cubic/f32/" TAG "x4", TW4, 16, (int)sizeof(S), P4 }

static const KernelEntry strip_table[] = {
	STRIP_ENTRIES(uint8_t, "u8", 4, 256, 1, 128),
	STRIP_ENTRIES(uint16_t, "u16", 2, 256, 1, 128),
	STRIP_ENTRIES(float, "f32", 1, 128, 1, 128),
	STRIP_ENTRIES(__half, "f16", 2, 256, 1, 128),
	STRIP_ENTRIES(u15_t, "u15", 2, 256, 1, 128),
};

// ALT4: 4-channel strips put red and blue of one pixel on neighbouring lanes (half the row span
// per warp load: no 2-way / 4-way bank conflicts for 8- and 16-byte pixels).
#define STREAM_ENTRIES(S, TAG, P3, TW3, P4, TW4, ALT4)                                                        \
	FIXCA_STREAM_ENTRY("stream/linear/f32/" TAG "x3", S, 3, 1, P3, TW3, false, false, false),             \
	FIXCA_STREAM_ENTRY("stream/linear/f32/" TAG "x4", S, 4, 1, P4, TW4, ALT4, false, false),              \
	FIXCA_STREAM_ENTRY("stream/cubic/f32/" TAG "x3", S, 3, 2, P3, TW3, false, false, false),              \
	FIXCA_STREAM_ENTRY("stream/cubic/f32/" TAG "x4", S, 4, 2, P4, TW4, ALT4, false, false)

static const KernelEntry stream_table[] = {
	STREAM_ENTRIES(uint8_t, "u8", 4, 256, 3, 192, false),	// 4-byte pixels: 3 columns = 12 bytes per lane, 7 loads per 3 outputs
	STREAM_ENTRIES(uint16_t, "u16", 2, 256, 3, 192, true),	// 8-byte pixels, lanes alternate channels: 3 columns = conflict-free, 7 loads per 3 outputs
	STREAM_ENTRIES(float, "f32", 1, 128, 1, 64, true),	// 16-byte pixels: 64 + halo columns fit one 2 KB TMA box
	STREAM_ENTRIES(__half, "f16", 2, 256, 3, 192, true),	// 16-bit floats: the layouts of u16
	STREAM_ENTRIES(u15_t, "u15", 2, 256, 3, 192, true),	// 15-bit unsigned in 16-bit storage: the layouts of u16
};

// Narrower 16-bit RGBA strips for the calls whose window (strip + shift + slack columns) would not fit one
// 2 KB TMA box at 192 columns (variant 4).
static const KernelEntry stream_u16x4_tw128[] = {
	FIXCA_STREAM_ENTRY("stream/linear/f32/u16x4/tw128", uint16_t, 4, 1, 1, 128, true, false, false),
	FIXCA_STREAM_ENTRY("stream/cubic/f32/u16x4/tw128", uint16_t, 4, 2, 1, 128, true, false, false),
	FIXCA_STREAM_ENTRY("stream/linear/f32/f16x4/tw128", __half, 4, 1, 1, 128, true, false, false),
	FIXCA_STREAM_ENTRY("stream/cubic/f32/f16x4/tw128", __half, 4, 2, 1, 128, true, false, false),
};

// channel-per-warp variants of the 4-channel strips, for A/B runs (FIXCA_STREAM_NOALT=1)
static const KernelEntry stream_x4_noalt[] = {
	FIXCA_STREAM_ENTRY("stream/linear/f32/u16x4/noalt", uint16_t, 4, 1, 1, 128, false, false, false),
	FIXCA_STREAM_ENTRY("stream/cubic/f32/u16x4/noalt", uint16_t, 4, 2, 1, 128, false, false, false),
	FIXCA_STREAM_ENTRY("stream/linear/f32/f32x4/noalt", float, 4, 1, 1, 128, false, false, false),
	FIXCA_STREAM_ENTRY("stream/cubic/f32/f32x4/noalt", float, 4, 2, 1, 128, false, false, false),
};

static const KernelEntry stream_u16x3_tw128[] = {
	FIXCA_STREAM_ENTRY("stream/linear/f32/u16x3/tw128", uint16_t, 3, 1, 2, 128, false, false, false),
	FIXCA_STREAM_ENTRY("stream/cubic/f32/u16x3/tw128", uint16_t, 3, 2, 2, 128, false, false, false),
};

// narrower-tile variants of the headline format, for tuning (FIXCA_STRIP_TW=128)
static const KernelEntry strip_u16x3_tw128[] = {
	{ (kernel_fn)strip_kernel<uint16_t, 3, 1, 2, 128>, "strip/linear/f32/u16x3/tw128", 128, 16, 2, 2 },
	{ (kernel_fn)strip_kernel<uint16_t, 3, 2, 2, 128>, "strip/cubic/f32/u16x3/tw128", 128, 16, 2, 2 },
};

// variant: 0 direct, 2 strip, 3 stream, 4 narrower stream (nullptr when there is none)
const KernelEntry *lookup_fast_variant(SampleKind kind, int nch, int interp, int variant)
{
	if (variant == 4)
		return ((kind == SK_U16 || kind == SK_F16) && nch == 4 && (interp == 1 || interp == 2))
			       ? &stream_u16x4_tw128[(kind == SK_F16 ? 2 : 0) + interp - 1] : nullptr;
	int s;
	switch (kind) {
	case SK_U8:  s = 0; break;
	case SK_U16: s = 1; break;
	case SK_F32: s = 2; break;
	case SK_F16: s = 3; break;
	case SK_U15: s = 4; break;
	default: return nullptr;
	}
	if ((nch != 3 && nch != 4) || (interp != 1 && interp != 2))
		return nullptr;
	const bool tw128 = kind == SK_U16 && nch == 3 && tuning().strip_tw128;	// tuning: narrower tiles for the headline format
	if (variant == 3 && nch == 4 && (kind == SK_U16 || kind == SK_F32) && tuning().stream_noalt)
		return &stream_x4_noalt[(kind == SK_F32 ? 2 : 0) + interp - 1];
	switch (variant) {
	case 3: return tw128 ? &stream_u16x3_tw128[interp - 1] : &stream_table[s * 4 + (interp - 1) * 2 + (nch - 3)];
	case 2: return tw128 ? &strip_u16x3_tw128[interp - 1] : &strip_table[s * 4 + (interp - 1) * 2 + (nch - 3)];
	default: return &direct_table[s * 4 + (interp - 1) * 2 + (nch - 3)];
	}
}

const KernelEntry *lookup_fast(SampleKind kind, int nch, int interp, bool tiled)
{
	if (!tiled)
		return lookup_fast_variant(kind, nch, interp, 0);
	return lookup_fast_variant(kind, nch, interp, tuning().fast_kernel);	// FIXCA_FAST_KERNEL = strip | stream (default), for A/B runs
}

} // namespace fixca
