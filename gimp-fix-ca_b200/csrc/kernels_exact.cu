// kernels_exact.cu -- Linear / Cubic in FP64 with the reference's operation
// order (ExactF64): bit-identical to fix-ca.c:1122-1320 for every format.
#include "fixca_internal.h"

namespace fixca {

#define EXACT_ENTRIES(S, TAG)                                                                                   \
	{ (kernel_fn)tiled_kernel<S, 3, 1, ExactF64, TILE_W>, "tiled/linear/f64/" TAG "x3", TILE_W, (int)sizeof(ExactF64::YCoef), (int)sizeof(S) }, \
	{ (kernel_fn)tiled_kernel<S, 4, 1, ExactF64, TILE_W>, "tiled/linear/f64/" TAG "x4", TILE_W, (int)sizeof(ExactF64::YCoef), (int)sizeof(S) }, \
	{ (kernel_fn)tiled_kernel<S, 3, 2, ExactF64, TILE_W>, "tiled/cubic/f64/" TAG "x3", TILE_W, (int)sizeof(ExactF64::YCoef), (int)sizeof(S) },  \
	{ (kernel_fn)tiled_kernel<S, 4, 2, ExactF64, TILE_W>, "tiled/cubic/f64/" TAG "x4", TILE_W, (int)sizeof(ExactF64::YCoef), (int)sizeof(S) },  \
	{ (kernel_fn)direct_kernel<S, 3, 1, ExactF64>, "direct/linear/f64/" TAG "x3", 0, 0, (int)sizeof(S) },   \
	{ (kernel_fn)direct_kernel<S, 4, 1, ExactF64>, "direct/linear/f64/" TAG "x4", 0, 0, (int)sizeof(S) },   \
	{ (kernel_fn)direct_kernel<S, 3, 2, ExactF64>, "direct/cubic/f64/" TAG "x3", 0, 0, (int)sizeof(S) },    \
	{ (kernel_fn)direct_kernel<S, 4, 2, ExactF64>, "direct/cubic/f64/" TAG "x4", 0, 0, (int)sizeof(S) }

static const KernelEntry exact_table[] = {
	EXACT_ENTRIES(uint8_t, "u8"),
	EXACT_ENTRIES(uint16_t, "u16"),
	EXACT_ENTRIES(uint32_t, "u32"),
	EXACT_ENTRIES(float, "f32"),
	EXACT_ENTRIES(double, "f64"),
	EXACT_ENTRIES(__half, "f16"),
	EXACT_ENTRIES(u15_t, "u15"),
	EXACT_ENTRIES(uint64_t, "u64"),	// the x87 long double steps of get_pixel / set_pixel restated (fixca_kernels.cuh)
};

const KernelEntry *lookup_exact(SampleKind kind, int nch, int interp, bool tiled)
{
	int s;
	switch (kind) {
	case SK_U8:  s = 0; break;
	case SK_U16: s = 1; break;
	case SK_U32: s = 2; break;
	case SK_F32: s = 3; break;
	case SK_F64: s = 4; break;
	case SK_F16: s = 5; break;
	case SK_U15: s = 6; break;
	case SK_U64: s = 7; break;
	default: return nullptr;
	}
	if ((nch != 3 && nch != 4) || (interp != 1 && interp != 2))
		return nullptr;
	return &exact_table[s * 8 + (tiled ? 0 : 4) + (interp - 1) * 2 + (nch - 3)];
}

} // namespace fixca
