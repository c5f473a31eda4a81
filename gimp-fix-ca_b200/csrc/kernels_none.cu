// kernels_none.cu -- interpolation = None: a byte-exact gather (fix-ca.c:1100-1121).
// Instantiated by sample size only; the data is never interpreted.
//
//   stream_kernel<.., INTERP = 0, ..>  persistent column strips, TMA ring (fixca_stream.cuh): 1-, 2- and 4-byte samples
//   tiled_kernel<.., NoArith, ..>      one CTA per tile: 8-byte samples, windows wider than a TMA box, A/B runs
//   direct_none_kernel                 per-pixel gather, any geometry
#include <cstdlib>
#include <cstring>
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

// Columns per thread / strip width / lane layout as in kernels_fast.cu (conflict-free lane strides).
static const KernelEntry none_stream_table[] = {
	{ (kernel_fn)stream_kernel<uint8_t, 3, 0, 4, 256>, "stream/none/copy/b1x3", 256, 0, 1, 4, 1 },
	{ (kernel_fn)stream_kernel<uint8_t, 4, 0, 3, 192>, "stream/none/copy/b1x4", 192, 0, 1, 3, 1 },
	{ (kernel_fn)stream_kernel<uint16_t, 3, 0, 2, 256>, "stream/none/copy/b2x3", 256, 0, 2, 2, 1 },
	{ (kernel_fn)stream_kernel<uint16_t, 4, 0, 3, 192, true>, "stream/none/copy/b2x4", 192, 0, 2, 3, 1 },
	{ (kernel_fn)stream_kernel<uint32_t, 3, 0, 1, 128>, "stream/none/copy/b4x3", 128, 0, 4, 1, 1 },
	{ (kernel_fn)stream_kernel<uint32_t, 4, 0, 1, 64, true>, "stream/none/copy/b4x4", 64, 0, 4, 1, 1 },
};

const KernelEntry *lookup_none_stream(int sample_bytes, int nch)
{
	if (tuning().none_tiled)	// FIXCA_NONE_KERNEL = tiled | stream (default), for A/B runs
		return nullptr;
	if (nch != 3 && nch != 4)
		return nullptr;
	switch (sample_bytes) {
	case 1: return &none_stream_table[nch - 3];
	case 2: return &none_stream_table[2 + nch - 3];
	case 4: return &none_stream_table[4 + nch - 3];
	default: return nullptr;
	}
}

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
