// kernels_fast.cu -- Linear / Cubic in FP32 on raw sample values (FastF32):
// within +-1 LSB of the reference for u8 / u16, ~2 ulp(1.0) for float.
#include "fixca_internal.h"

namespace fixca {

#define FAST_ENTRIES(S, TAG)                                                                                  \
	{ (kernel_fn)tiled_kernel<S, 3, 1, FastF32, TILE_W>, "tiled/linear/f32/" TAG "x3", TILE_W, (int)sizeof(FastF32::YCoef), (int)sizeof(S) }, \
	{ (kernel_fn)tiled_kernel<S, 4, 1, FastF32, TILE_W>, "tiled/linear/f32/" TAG "x4", TILE_W, (int)sizeof(FastF32::YCoef), (int)sizeof(S) }, \
	{ (kernel_fn)tiled_kernel<S, 3, 2, FastF32, TILE_W>, "tiled/cubic/f32/" TAG "x3", TILE_W, (int)sizeof(FastF32::YCoef), (int)sizeof(S) },  \
	{ (kernel_fn)tiled_kernel<S, 4, 2, FastF32, TILE_W>, "tiled/cubic/f32/" TAG "x4", TILE_W, (int)sizeof(FastF32::YCoef), (int)sizeof(S) },  \
	{ (kernel_fn)direct_kernel<S, 3, 1, FastF32>, "direct/linear/f32/" TAG "x3", 0, 0, (int)sizeof(S) },  \
	{ (kernel_fn)direct_kernel<S, 4, 1, FastF32>, "direct/linear/f32/" TAG "x4", 0, 0, (int)sizeof(S) },  \
	{ (kernel_fn)direct_kernel<S, 3, 2, FastF32>, "direct/cubic/f32/" TAG "x3", 0, 0, (int)sizeof(S) },   \
	{ (kernel_fn)direct_kernel<S, 4, 2, FastF32>, "direct/cubic/f32/" TAG "x4", 0, 0, (int)sizeof(S) }

static const KernelEntry fast_table[] = {
	FAST_ENTRIES(uint8_t, "u8"),
	FAST_ENTRIES(uint16_t, "u16"),
	FAST_ENTRIES(float, "f32"),
};

const KernelEntry *lookup_fast(SampleKind kind, int nch, int interp, bool tiled)
{
	int s;
	switch (kind) {
	case SK_U8:  s = 0; break;
	case SK_U16: s = 1; break;
	case SK_F32: s = 2; break;
	default: return nullptr;
	}
	if ((nch != 3 && nch != 4) || (interp != 1 && interp != 2))
		return nullptr;
	return &fast_table[s * 8 + (tiled ? 0 : 4) + (interp - 1) * 2 + (nch - 3)];
}

} // namespace fixca
