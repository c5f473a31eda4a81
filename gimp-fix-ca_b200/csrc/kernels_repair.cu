// kernels_repair.cu -- Linear / Cubic, bit-identical to the reference (fix-ca.c:1122-1320), for 8-bit samples:
// stream_kernel<..., REPAIR> runs the FP32 streaming pipeline and recomputes, in the reference's own FP64
// arithmetic, exactly the samples whose FP32 value lies within a proven error bound of a rounding boundary
//  -- behind the streaming grid, from a queue in global memory (REPAIR = 2, repair_patch_kernel: the
// default), or inside it, from per-warp queues in shared memory (REPAIR = 1, FIXCA_EXACT_KERNEL=inline).  Every other format, and geometries the streaming kernel cannot take, stay on
// tiled_kernel<ExactF64> (kernels_exact.cu).
#include "fixca_internal.h"

namespace fixca {

// 8-bit samples: the FP32 pipeline's bound (|FP32 - reference| <= 2.7e-4 LSB) sends 0.05 % of the samples to the FP64
// repair and EXACT runs 2.1x (Cubic) / 1.7x (Linear) faster than the FP64 tile kernel.  For 16-bit samples the same
// bound is 0.07 LSB, 14 % of the samples need the repair, and that kernel measured 2.34 ms against 1.19 ms (100 MP
// RGB16 Cubic): they take the WIDE form instead -- the separable sums in FP64 (12 FP64 operations per sample against
// the reference order's 64), bound 3.5e-8 LSB, two samples in a million recomputed.
// layouts as in kernels_fast.cu (columns per thread, strip width)
#define REPAIR_ENTRIES(S, TAG, P3, TW3, P4, TW4)                                                              \
	FIXCA_STREAM_ENTRY("stream/linear/f32+f64inline/" TAG "x3", S, 3, 1, P3, TW3, false, true, false),    \
	FIXCA_STREAM_ENTRY("stream/linear/f32+f64inline/" TAG "x4", S, 4, 1, P4, TW4, false, true, false),    \
	FIXCA_STREAM_ENTRY("stream/cubic/f32+f64inline/" TAG "x3", S, 3, 2, P3, TW3, false, true, false),     \
	FIXCA_STREAM_ENTRY("stream/cubic/f32+f64inline/" TAG "x4", S, 4, 2, P4, TW4, false, true, false)
#define WIDE_ENTRIES(S, TAG, P3, TW3, P4, TW4, ALT4)                                                          \
	FIXCA_STREAM_ENTRY("stream/linear/f64+exact/" TAG "x3", S, 3, 1, P3, TW3, false, false, true),        \
	FIXCA_STREAM_ENTRY("stream/linear/f64+exact/" TAG "x4", S, 4, 1, P4, TW4, ALT4, false, true),         \
	FIXCA_STREAM_ENTRY("stream/cubic/f64+exact/" TAG "x3", S, 3, 2, P3, TW3, false, false, true),         \
	FIXCA_STREAM_ENTRY("stream/cubic/f64+exact/" TAG "x4", S, 4, 2, P4, TW4, ALT4, false, true)

// the deferred form (REPAIR = 2): near-tie samples queued in global memory, repair_patch_kernel behind the stream kernel
#define DEFER_ENTRY(NAME, S, NCH, INTERP, P, TW)                                                                    \
	{ (kernel_fn)stream_kernel<S, NCH, INTERP, P, TW, false, 2, false>, NAME, TW, 0, (int)sizeof(S), P, 1, 1,   \
	  stream_setup_kernel<S, NCH, INTERP, P, TW, false, 2, false>,                                              \
	  (int)sizeof(StreamColumnState<float, P, (INTERP) == 1 ? 3 : 5>), repair_patch_kernel<S, NCH, INTERP> }
#define DEFER_ENTRIES(S, TAG, P3, TW3, P4, TW4)                                                               \
	DEFER_ENTRY("stream/linear/f32+f64/" TAG "x3", S, 3, 1, P3, TW3),                                    \
	DEFER_ENTRY("stream/linear/f32+f64/" TAG "x4", S, 4, 1, P4, TW4),                                    \
	DEFER_ENTRY("stream/cubic/f32+f64/" TAG "x3", S, 3, 2, P3, TW3),                                     \
	DEFER_ENTRY("stream/cubic/f32+f64/" TAG "x4", S, 4, 2, P4, TW4)

static const KernelEntry defer_table[] = {
	DEFER_ENTRIES(uint8_t, "u8", 4, 256, 3, 192),
};

static const KernelEntry repair_table[] = {
	REPAIR_ENTRIES(uint8_t, "u8", 4, 256, 3, 192),
	WIDE_ENTRIES(uint16_t, "u16", 2, 256, 3, 192, true),
	WIDE_ENTRIES(u15_t, "u15", 2, 256, 3, 192, true),
};

const KernelEntry *lookup_exact_stream(SampleKind kind, int nch, int interp)
{
	if ((nch != 3 && nch != 4) || (interp != 1 && interp != 2) || tuning().exact_tiled)
		return nullptr;
	int s;
	switch (kind) {
	case SK_U8:  s = 0; break;
	case SK_U16: s = 1; break;
	case SK_U15: s = 2; break;
	default: return nullptr;
	}
	if (s == 0 && !tuning().exact_inline)
		return &defer_table[(interp - 1) * 2 + (nch - 3)];
	return &repair_table[s * 4 + (interp - 1) * 2 + (nch - 3)];
}

} // namespace fixca
