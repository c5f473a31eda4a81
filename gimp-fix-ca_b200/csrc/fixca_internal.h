// fixca_internal.h -- glue between the host driver (fixca_api.cu) and the
// kernel translation units (kernels_*.cu).
#pragma once

#include <cstddef>
#include "fixca_kernels.cuh"
#include "fixca_strip.cuh"
#include "fixca_stream.cuh"

namespace fixca {

// SK_F16 (bpc = -2) is the extension of SURVEY.md 8(f) #4: the reference carries it as commented-out code
// SK_U15 (bpc = 15) is the other half of that row: 15-bit unsigned samples, which the reference only rejects
enum SampleKind { SK_U8 = 0, SK_U16, SK_U32, SK_U64, SK_F32, SK_F64, SK_F16, SK_U15, SK_COUNT };
enum ArithKind  { AR_COPY = 0, AR_EXACT = 1, AR_FAST = 2 };

// Every kernel has the signature  __global__ void k(const KernelArgs).
typedef void (*kernel_fn)(const KernelArgs);

struct KernelEntry {
	kernel_fn   fn;
	const char *name;	// "tiled/cubic/f32/u16x3"
	int         tw;		// tile width (0 for direct kernels)
	int         ycoef_bytes;// per-row table entry size (tiled)
	int         sample_bytes;
	int         strip_p;	// > 0: strip_kernel with this many columns per thread (blockDim = 2 * tw / strip_p)
	int         stream;	// != 0: stream_kernel (blockDim = 2 * tw / strip_p + 32, grid = strips x segments)
	int         repair;	// 1: the exact-repair form of stream_kernel (FP32 pipeline, per-warp queues in shared memory); 2: WIDE (FP64 pipeline)
	// Linear / Cubic stream kernels: the kernel that fills the plan's column set-up table (grid = strips, block = the
	// compute threads) and the size of one thread's record
	void      (*setup)(const KernelArgs, void *);
	int         setup_rec_bytes;
	// deferred exact-repair form (stream_kernel<..., REPAIR = 2>): the kernel launched behind it, which recomputes the
	// queued near-tie samples (repair_patch_kernel)
	void      (*patch)(const KernelArgs, const PatchArgs);
};

// table entry of stream_kernel<S, NCH, INTERP, P, TW, ALT, REPAIR, WIDE> (INTERP != 0) with its set-up kernel
#define FIXCA_STREAM_ENTRY(NAME, S, NCH, INTERP, P, TW, ALT, REPAIR, WIDE)                                        \
	{ (kernel_fn)stream_kernel<S, NCH, INTERP, P, TW, ALT, REPAIR, WIDE>, NAME, TW, 0, (int)sizeof(S), P, 1,  \
	  (WIDE) ? 2 : (REPAIR) ? 1 : 0, stream_setup_kernel<S, NCH, INTERP, P, TW, ALT, REPAIR, WIDE>,            \
	  (int)sizeof(StreamColumnState<typename std::conditional<WIDE, double, float>::type, P, (P) == 1 ? ((INTERP) == 1 ? 2 : 4) : ((INTERP) == 1 ? 3 : 5)>) }

constexpr int TILE_W = 128;

// Tuning switches, read from the environment ONCE per process (and again only when a
// caller asks through fixca_cuda_reload_tuning(): tests and A/B probes that change the environment
// mid-process).  `generation` is part of the plan-cache key.
struct Tuning {
	int generation = 0;
	int fast_kernel = 3;	// lookup_fast_variant(): 2 strip, 3 stream (FIXCA_FAST_KERNEL=strip|stream)
	int none_tiled = 0;	// FIXCA_NONE_KERNEL=tiled
	int exact_tiled = 0;	// FIXCA_EXACT_KERNEL=tiled: EXACT Linear / Cubic on tiled_kernel<ExactF64> for every format
	int exact_inline = 0;	// FIXCA_EXACT_KERNEL=inline: 8-bit EXACT on the in-kernel repair form (A/B against the deferred form)
	int strip_tw128 = 0;	// FIXCA_STRIP_TW=128
	int stream_noalt = 0;	// FIXCA_STREAM_NOALT=1
	int tile_h = 0, tile_ctas = 3;
	int stream_ctas = 0, stream_depth = 0, stream_segs = 0, stream_waves = 0;
	int stream_tlead = 0;	// FIXCA_STREAM_TLEAD=1|2: chunks the pass-through tile is requested ahead (default 2 at depth >= 2)
	int stream_debug = 0;	// honoured by -DFIXCA_TUNING builds only
	int no_pdl = 0, verbose = 0;
	int chunk_mb = 0, copy_threads = 0;
	int no_chunk_ramp = 0;	// FIXCA_CHUNK_RAMP=0: uniform staging chunks for pinned callers (A/B)
	int precision_fast = 0;	// FIXCA_PRECISION=fast: what the flag-less fixca_cuda_region() computes in
};
const Tuning &tuning();

// sample_bytes in {1,2,4,8}; nch in {3,4}
const KernelEntry *lookup_none(int sample_bytes, int nch, bool tiled);
// the streaming None kernel for this format, or nullptr (8-byte samples, FIXCA_NONE_KERNEL=tiled)
const KernelEntry *lookup_none_stream(int sample_bytes, int nch);
// kind in {SK_U8,SK_U16,SK_U32,SK_F32,SK_F64,SK_F16,SK_U15}; interp in {1,2}
const KernelEntry *lookup_exact(SampleKind kind, int nch, int interp, bool tiled);
// the exact-repair streaming kernel for this format, or nullptr (floats, u32, FIXCA_EXACT_KERNEL=tiled)
const KernelEntry *lookup_exact_stream(SampleKind kind, int nch, int interp);
// kind in {SK_U8,SK_U16,SK_F32,SK_F16,SK_U15}; interp in {1,2}
const KernelEntry *lookup_fast(SampleKind kind, int nch, int interp, bool tiled);
// variant: 0 direct, 2 strip, 3 stream, 4 narrower stream (may be nullptr)
const KernelEntry *lookup_fast_variant(SampleKind kind, int nch, int interp, int variant);

// kernels_meta.cu: fills the per-plan tables of a streaming launch (KernelArgs::meta_tab / span_tab), one record per
// 8-row chunk of [a.y1, a.y2); mode = KernelEntry::repair (0 FP32 weights, 1 exact-repair, 2 WIDE)
size_t stream_meta_record_bytes(int mode);
cudaError_t launch_stream_meta(int interp, int mode, SampleKind kind, const KernelArgs &a, void *meta, void *span, int nchunks,
			       cudaStream_t st);

// ... and None's per-plan column table (KernelArgs::col_i0): ncols nearest source columns per channel
cudaError_t launch_stream_cols(const KernelArgs &a, int ncols, void *i0, cudaStream_t st);

// kernels_preview.cu: saturate() + centerline() on destination rows [y1, y2) (fix-ca.c:1322-1327)
cudaError_t launch_preview(int kind, int nch, unsigned char *dst, long long pitch, int dst_row0, int y1, int y2, int width,
			   int xc, int yc, double saturation, cudaStream_t st);

// kernels_preview.cu: the 8-bit preview buffer (fix-ca.c:659-671) of window columns [x, x + pw) of corrected rows [y1, y2)
cudaError_t launch_to8(int kind, int nch, const unsigned char *src, long long pitch, int src_row0, int y1, int y2, int x, int pw,
		       int bpp, unsigned char *out, cudaStream_t st);

} // namespace fixca
