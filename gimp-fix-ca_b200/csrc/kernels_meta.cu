// kernels_meta.cu -- the per-plan tables of the streaming kernels (stream_meta_kernel, fixca_stream.cuh): the
// vertical weights / tap rows of every 8-row chunk of a band and the source rows it touches, filled once when a
// launch plan is made instead of once per CTA (fix-ca.c:813-820 coordinates, :1141-1148 / :1219-1256 taps).
#include "fixca_internal.h"

namespace fixca {

size_t stream_meta_record_bytes(int mode) { return mode == 2 ? sizeof(StreamMetaWide) : sizeof(StreamMeta); }

// the scale position_weights() folds into the vertical weights of this sample type (1 / max and the power of two of
// the subnormal-operand codec, fixca_strip.cuh)
static float inv_max_of(SampleKind kind)
{
	switch (kind) {
	case SK_U8:  return StripCodec<uint8_t>::kInvMax;
	case SK_U16: return StripCodec<uint16_t>::kInvMax;
	case SK_U15: return StripCodec<u15_t>::kInvMax;
	default:     return 1.0f;	// float, half (None never reads it)
	}
}

cudaError_t launch_stream_meta(int interp, int mode, SampleKind kind, const KernelArgs &a, void *meta, void *span, int nchunks,
			       cudaStream_t st)
{
	const dim3 block(128), grid((unsigned)((nchunks + 3) / 4));
	const float im = inv_max_of(kind);
	StreamSpan *sp = reinterpret_cast<StreamSpan *>(span);
	if (interp == 0)
		stream_meta_kernel<0, 0><<<grid, block, 0, st>>>(a, im, meta, sp, nchunks);
	else if (interp == 1 && mode == 0)
		stream_meta_kernel<1, 0><<<grid, block, 0, st>>>(a, im, meta, sp, nchunks);
	else if (interp == 1 && mode == 1)
		stream_meta_kernel<1, 1><<<grid, block, 0, st>>>(a, im, meta, sp, nchunks);
	else if (interp == 1 && mode == 2)
		stream_meta_kernel<1, 2><<<grid, block, 0, st>>>(a, im, meta, sp, nchunks);
	else if (interp == 2 && mode == 0)
		stream_meta_kernel<2, 0><<<grid, block, 0, st>>>(a, im, meta, sp, nchunks);
	else if (interp == 2 && mode == 1)
		stream_meta_kernel<2, 1><<<grid, block, 0, st>>>(a, im, meta, sp, nchunks);
	else if (interp == 2 && mode == 2)
		stream_meta_kernel<2, 2><<<grid, block, 0, st>>>(a, im, meta, sp, nchunks);
	else
		return cudaErrorInvalidValue;
	return cudaGetLastError();
}

cudaError_t launch_stream_cols(const KernelArgs &a, int ncols, void *i0, cudaStream_t st)
{
	const dim3 block(256), grid((unsigned)((ncols + 255) / 256), 2);
	stream_cols_kernel<0><<<grid, block, 0, st>>>(a, ncols, reinterpret_cast<int *>(i0));
	return cudaGetLastError();
}

} // namespace fixca
