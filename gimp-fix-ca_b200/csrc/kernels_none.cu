// kernels_none.cu -- interpolation = None: a byte-exact gather (fix-ca.c:1100-1121).
// Instantiated by sample size only; the data is never interpreted.
#include "fixca_internal.h"

namespace fixca {

#define NONE_ENTRIES(U, B)                                                                          \
	{ (kernel_fn)tiled_kernel<U, 3, 0, NoArith, TILE_W>, "tiled/none/copy/b" #B "x3", TILE_W, (int)sizeof(NoArith::YCoef), B }, \
	{ (kernel_fn)tiled_kernel<U, 4, 0, NoArith, TILE_W>, "tiled/none/copy/b" #B "x4", TILE_W, (int)sizeof(NoArith::YCoef), B }, \
	{ (kernel_fn)direct_none_kernel<U, 3>, "direct/none/copy/b" #B "x3", 0, 0, B },                \
	{ (kernel_fn)direct_none_kernel<U, 4>, "direct/none/copy/b" #B "x4", 0, 0, B }

static const KernelEntry none_table[] = {
	NONE_ENTRIES(uint8_t, 1),
	NONE_ENTRIES(uint16_t, 2),
	NONE_ENTRIES(uint32_t, 4),
	NONE_ENTRIES(uint64_t, 8),
};

const KernelEntry *lookup_none(int sample_bytes, int nch, bool tiled)
{
	int s;
	switch (sample_bytes) {
	case 1: s = 0; break;
	case 2: s = 1; break;
	case 4: s = 2; break;
	case 8: s = 3; break;
	default: return nullptr;
	}
	if (nch != 3 && nch != 4)
		return nullptr;
	return &none_table[s * 4 + (tiled ? 0 : 2) + (nch - 3)];
}

} // namespace fixca
