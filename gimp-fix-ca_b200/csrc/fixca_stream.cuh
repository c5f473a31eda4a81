// fixca_stream.cuh -- the streaming kernel: None, FAST (FP32) Linear and FAST Cubic (fix_ca_region's row
// loop, fix-ca.c:1091-1320), the path bench.py measures.
//
// strip_kernel (fixca_strip.cuh) pays per 16-row tile for: the FP64 column / row
// setup, re-priming its 4-row ring (3 halo rows of horizontal work per channel),
// a halo of window rows re-fetched from L2, and CTA-wide barriers around one
// exposed TMA round trip.  Its ncu capture (profiles/r01_ncu_strip_*_d.md) shows
// ~40 % of the issued instructions and most stall cycles ("barrier", "wait")
// coming from that per-tile overhead rather than from the row loop.
//
// stream_kernel keeps the row loop and amortises everything else:
//
//   * a CTA owns a TW-column strip and a long run of rows (a "segment") of one frame (blockIdx.z); the
//     per-column weights and the ring of horizontal rows live in registers for
//     the whole segment, so every source row is filtered horizontally once per
//     channel and nothing is re-primed;
//   * rows move through shared memory as a TMA + mbarrier pipeline driven by one
//     producer warp: source rows land in a circular window (each row fetched
//     from global memory once per strip), the pass-through pixels of the next
//     output chunk land in a staging buffer, finished chunks leave with TMA bulk
//     stores; D chunks are in flight ahead of the compute warps;
//     Copies are TMA *tensor* tiles (cp.async.bulk.tensor.3d over CUtensorMaps of the source and
//     destination frames): one instruction moves a 4-row group of the window, one the 8 x TW
//     pass-through tile, one stores a finished chunk; rows and columns outside the image are
//     zero-filled / clipped by the TMA unit.  (A first version issued one bulk copy per row --
//     24 per chunk -- and the producer's own instruction stream, ~1.9 us per chunk, was the
//     critical path of the whole kernel: profiles/r01_stream_pipe_sweep_i.md.)
//   * the per-row vertical weights (FP64 coordinates, fix-ca.c:813-820, weights ordered by tap position) and
//     the source rows every chunk touches are the same for every strip and every frame: stream_meta_kernel
//     fills a table of them once per launch plan (one StreamMeta + StreamSpan per 8-row chunk) and the TMA
//     lane copies a chunk's record into shared memory with the chunk's pass-through tile.  (Until r02 a
//     second helper warp recomputed them in every CTA: 12 % of the RGB8 kernel's instructions.)
//   * the compute warps never meet at a CTA barrier: they wait on "full" mbarriers and arrive on
//     "done" mbarriers.
//
// Row loop: per warp one of three horizontal forms -- narrow (T weights over P + T - 1 shared
// samples), regular (T + 1 weights: the extra one absorbs the drift of the tap window), bent (per-column
// samples, tap windows squeezed against an image edge) -- then a vertical pass over a register ring whose
// weights are ordered by tap position, so the arithmetic does not depend on ring phase, tile, segment, band or
// batch layout; strip_kernel computes the same bytes.  4-channel 16/32-bit strips put red and blue of a pixel
// on neighbouring lanes (ALT).  INTERP = 0 copies raw sample bytes instead (bit-exact None).  The kernel is
// launched with programmatic stream serialization: only the TMA lane touches global memory, after
// griddep_wait().
#pragma once

#include <type_traits>
#ifdef FIXCA_EXP_TIMING
#include <cstdio>
#endif

#include <cuda.h>	// CUtensorMap (the maps are encoded on the host, fixca_api.cu)

#include "fixca_strip.cuh"

namespace fixca {

// Timing experiments (FIXCA_STREAM_DEBUG: bit 0 skips the arithmetic -- wrong pixels --, bit 1 disables the narrow
// form, bit 2 skips the window / tile loads, bit 3 the stores) exist in -DFIXCA_TUNING builds only (make TUNING=1); the
// release library ignores KernelArgs::debug.
#ifdef FIXCA_TUNING
#define STREAM_DEBUG_BIT(a, bit) ((a).debug & (bit))
#else
#define STREAM_DEBUG_BIT(a, bit) 0
#endif

constexpr int STREAM_CH = 8;	// output rows per chunk
// Pipeline depth D (KernelArgs::depth, chosen by the host from the shared memory left): chunks whose
// window rows are requested ahead of the compute warps -- these come from HBM and carry the
// latency.  D + 1 chunks have live "full"/"done" barriers and metadata.  The pass-through tile of a
// chunk is an L2 hit (its rows went through the window a moment ago) and is requested only one
// chunk ahead, so three staging buffers suffice: one loading, one computing, one draining.
constexpr int STREAM_MAX_D = 8;
constexpr int STREAM_MAX_NF = STREAM_MAX_D + 1;
constexpr int STREAM_NSTG = 3;
#define STREAM_COL_SLACK(P) ((P) + 4)	// >= NS = P + taps: pixels of window slack per side (host and device)

// Fan-out (the all-gather form of the reassembly): besides tm_out, every finished chunk is stored
// into up to STREAM_MAX_FAN further destination frames -- other GPUs' copies of the frame, mapped over NVLink -- by
// the same TMA lane, from the same staging buffer.  n = 0 for every ordinary launch.
constexpr int STREAM_MAX_FAN = 7;
struct alignas(64) StreamFanout {
	CUtensorMap tm[STREAM_MAX_FAN];
	int n;
};

template <class V4> struct StreamMetaT {
	V4     wy[STREAM_CH][2];	// vertical weights per output row and channel, by tap position (position_weights)
	int    last[2][STREAM_CH + 1];	// highest tap row of each output row (INT_MAX after the chunk's last row)
	int    s_end[2];		// = last[c][nrows - 1]
	int    simple[2];		// full chunk whose rows finish on CH consecutive source rows (the usual case)
	int    first[2];		// lowest tap row of the chunk's first output row (where a segment that starts here begins to filter)
};
typedef StreamMetaT<float4> StreamMeta;		// FP32 pipelines
typedef StreamMetaT<dvec4> StreamMetaWide;	// WIDE: FP64 weights
static_assert(sizeof(StreamMeta) % 16 == 0 && sizeof(StreamMetaWide) % 16 == 0, "records move as bulk copies");
// first / last source row the chunk's output rows touch, over both channels (what the TMA lane has to have in the ring)
struct alignas(8) StreamSpan { int lo, hi; };

struct StreamHeader {
	unsigned long long full[STREAM_MAX_NF];
	unsigned long long done[STREAM_MAX_NF];
	int col_lo[2], col_hi[2];
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Wait of a helper warp, with a suspend-time hint.  The hint does not park the warp: ncu's per-instruction counts
// (r01 Z / b / c, 32 x 4K RGB8 frames) show these two loops coming back every ~20 cycles and executing a quarter of
// all warp instructions of the kernel (60-150 polls per wait) -- whatever the hint, with or without a nanosleep.u32
// between the polls, and whether the compute warps arrive per thread or per warp.  The polls only fill issue slots
// the compute warps leave empty (so "issue slots busy" in the ncu summaries overstates the compute warps' own
// pressure by that quarter), and the prompt wake-up is worth having: with the hand-over on named barriers
// (bar.arrive / bar.sync: the waiting warps parked by the hardware, no polling at all) the Cubic kernels were
// 3-9 % slower (headline 0.198 -> 0.204 ms, 128 x 4K RGB8 68 -> 65 %), only None gained (100 MP RGB16 94.5 -> 97.5 %).
// A/B (FIXCA_TUNING builds, FIXCA_STREAM_DEBUG >> 8 = nanoseconds): a helper warp that tests the barrier and then
// really sleeps (nanosleep.u32, no wake-up on barrier traffic) instead of the try_wait loop below
__device__ __forceinline__ void mbar_wait_naps(uint64_t *bar, uint32_t parity, uint32_t nap_ns)
{
	for (;;) {
		uint32_t done;
		asm volatile(
			"{\n\t.reg .pred p;\n\t"
			"mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			"selp.u32 %0, 1, 0, p;\n\t}"
			: "=r"(done)
			: "r"(smem_u32(bar)), "r"(parity)
			: "memory");
		if (done)
			return;
		asm volatile("nanosleep.u32 %0;" ::"r"(nap_ns) : "memory");
	}
}
__device__ __forceinline__ void mbar_wait_sleepy(uint64_t *bar, uint32_t parity, uint32_t hint_ns)
{
	uint32_t done;
	do {
		asm volatile(
			"{\n\t.reg .pred p;\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
			"selp.u32 %0, 1, 0, p;\n\t}"
			: "=r"(done)
			: "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
			: "memory");
	} while (!done);
}
// A compute warp hands its part of a chunk over: every lane's staging writes are made visible to the async proxy
// (the TMA store the producer issues next), the warp converges, one lane arrives: 4-10 barrier arrivals per chunk
// instead of 128-320.
__device__ __forceinline__ void warp_arrive(uint64_t *bar)
{
	fence_proxy_async_smem();
	__syncwarp();
	if ((threadIdx.x & 31) == 0)
		mbar_arrive(bar);
}
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
	asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// Tensor tile of one frame, global -> shared; coordinates in elements of the map (c0: 8-byte units, c1: rows,
// c2: frame of the batch)
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *tm, int c0, int c1, int c2, uint64_t *bar)
{
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
		     ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
		     : "memory");
}
// Tensor tile, shared -> global (clipped to the tensor's extent)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *tm, int c0, int c1, int c2, const void *smem_src)
{
	asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
		     ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(smem_src))
		     : "memory");
}
// Programmatic dependent launch: stream kernels are launched with programmatic stream serialization, so the
// CTAs of the next launch in the stream may become resident (and run their set-up) while the last CTAs of this
// one drain; nothing touches global memory before griddep_wait(), which returns once the previous grid has
// completed and its writes are visible.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap *tm)
{
	asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

// lowest / highest source row output row y touches, over both channels (None: the nearest row, fix-ca.c:1105-1106;
// Linear :1141-1148; Cubic :1219-1256); the map is monotone, so a chunk's extremes sit at its first and last row
template <int INTERP>
__device__ __forceinline__ void stream_tap_rows(const Geometry &g, int y, int &lo, int &hi)
{
	constexpr int T = INTERP == 0 ? 1 : INTERP == 1 ? 2 : 4;
	constexpr int OFF = INTERP == 2 ? 1 : 0;
	lo = INT_MAX;
	hi = 0;
#pragma unroll
	for (int c = 0; c < 2; ++c) {
		int first, lastr;
		if constexpr (INTERP == 0) {
			first = lastr = nearest_index(g.y[c], y);
		} else {
			double td;
			const int i0 = base_index(g.y[c], y, td);
			first = max(i0 - OFF, 0);
			lastr = min(i0 + T - 1 - OFF, g.height - 1);
		}
		lo = min(lo, first);
		hi = max(hi, lastr);
	}
}

// ---------------------------------------------------------------------------------------------------------
// stream_meta_kernel: the per-plan tables of a streaming launch, one warp per 8-row chunk of [y1, y2).
//   meta[j]  vertical weights by tap position (position_weights: FP64 coordinates fix-ca.c:813-820, Linear
//            :1141-1148, Cubic :1219-1256), newest tap row of every output row, the "simple" flags
//   span[j]  first / last source row the chunk touches over both channels
// The records depend on the y axes, the band [y1, y2) and, for None, the ring geometry -- not on the strip, the
// frame or the buffers -- so every CTA of every launch of the plan reads the same table (L2-resident after the
// first strip).  MODE 0: FP32 weights; 1: the exact-repair form (weights in FP64, rounded once); 2: WIDE (FP64).
// ---------------------------------------------------------------------------------------------------------
template <int INTERP, int MODE>
__global__ void __launch_bounds__(128) stream_meta_kernel(const KernelArgs a, const float inv_max, void *const meta_out,
							   StreamSpan *const span_out, const int nchunks)
{
	constexpr int CH = STREAM_CH;
	[[maybe_unused]] constexpr int OFF = INTERP == 2 ? 1 : 0;
	static_assert(!(INTERP == 0 && MODE != 0), "None has one form");
	typedef typename std::conditional<MODE == 2, StreamMetaWide, StreamMeta>::type Meta;
	const int lane = threadIdx.x & 31;
	const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if (i >= nchunks)
		return;
	const int H = a.g.height;
	const int y_first = a.y1 + i * CH;
	const int nr = min(CH, a.y2 - y_first);
	Meta &m = reinterpret_cast<Meta *>(meta_out)[i];
	const int ch = (lane / CH) & 1, r = lane % CH;
	const bool mine = lane < 2 * CH && r < nr;
	int last = 0;
	if constexpr (INTERP == 0) {
		// None: the ring offset of the one source row each output row copies from (rows past
		// the band's end repeat its last row: the copy loop is not predicated, the store clips)
		if (lane < 2 * CH) {
			const int src = nearest_index(a.g.y[ch], min(y_first + r, a.y2 - 1));
			m.wy[r][ch] = make_float4(__int_as_float((src % a.ring_rows) * a.win_pitch), 0.f, 0.f, 0.f);
		}
	}
	if (mine) {
		if constexpr (INTERP == 0)
			last = nearest_index(a.g.y[ch], y_first + r);
		else if constexpr (MODE == 2)
			m.wy[r][ch] = position_weights_wide<INTERP>(a.g.y[ch], y_first + r, H, kWideScale, last);
		else
			m.wy[r][ch] = position_weights<INTERP, MODE == 1>(a.g.y[ch], y_first + r, H, inv_max, last);
		m.last[ch][r] = last;
		if (r == nr - 1) {
			m.last[ch][nr] = INT_MAX;
			m.s_end[ch] = last;
		}
		if (r == 0) {
			if constexpr (INTERP == 0) {
				m.first[ch] = last;
			} else {
				double td;
				m.first[ch] = max(base_index(a.g.y[ch], y_first, td) - OFF, 0);
			}
		}
	}
	// "simple": a full chunk in which every row completes exactly one source row after
	// the previous one -- one horizontal row in, one output row out, CH times
	const int prev = __shfl_up_sync(0xffffffffu, last, 1);
	const bool ok = mine && (r == 0 || last == prev + 1);
	const unsigned okmask = __ballot_sync(0xffffffffu, ok);
	if (lane < 2) {
		const unsigned want = ((1u << CH) - 1u) << (lane * CH);
		m.simple[lane] = (nr == CH) && ((okmask & want) == want);
	}
	// first source row of the chunk's first output row, last source row of its last one
	if (lane < 2) {
		int lo, hi;
		stream_tap_rows<INTERP>(a.g, lane == 0 ? y_first : y_first + nr - 1, lo, hi);
		if (lane == 0)
			span_out[i].lo = lo;
		else
			span_out[i].hi = hi;
	}
}

// ---------------------------------------------------------------------------------------------------------
// stream_cols_kernel (None): the nearest source column of every column of the strips' column range [0, ncols),
// per channel (fix-ca.c:801-808, :1105-1106) -- what every copy thread used to evaluate in FP64 at CTA start.
// ---------------------------------------------------------------------------------------------------------
template <int UNUSED = 0>	// (a template only so that the header can be included by several translation units)
__global__ void __launch_bounds__(256) stream_cols_kernel(const KernelArgs a, const int ncols, int *const i0_out)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
	if (x < ncols)
		i0_out[c * ncols + x] = nearest_index(a.g.x[c], x);
}

// ---------------------------------------------------------------------------------------------------------
// Column set-up of the Linear / Cubic kernels.  What a compute thread needs for its P columns -- the folded
// horizontal weights, the sample offsets, the form its warp runs -- depends on the strip and the thread, not on the
// segment, the frame or the buffers: stream_setup_kernel evaluates it once per plan with the thread layout of the
// main kernel (one CTA per strip, blockDim = the compute threads, so the warp votes are the same), and every CTA of
// every launch loads its threads' records (7 x 16 bytes for RGB8) instead of re-deriving them: the set-up was
// 4-6 k cycles of every CTA, a fifth of a one-wave launch (scripts/timing_probe.py).
//
// regular: the P columns share one window of NS consecutive samples; wt[k][jj] weighs sample k + jj.
// otherwise ("bent": tap windows squeezed against an image edge, fix-ca.c:1271-1298): column k reads its
// own T consecutive samples from byte offset cofs[k]; wt[k][jj < T] weighs sample jj, clamped taps merged.
// Both forms run the same row loop without index arithmetic, so a bent warp costs about as much as a
// regular one (a per-tap path with clamps in the loop made every CTA of an edge strip a 1.4x straggler).
// ---------------------------------------------------------------------------------------------------------
template <class A, int P, int NW> struct alignas(16) StreamColumnState {
	A   wt[P][NW];
	int cofs[P];		// ... relative to colbase
	int colbase;		// byte offset of shared sample 0 from the window row start
	int cmax;		// bent form: offset of the image's last column (samples past it carry no weight)
	int form;		// of the thread's warp: 0 bent, 1 regular, 2 regular and narrow (the extra weight is 0 everywhere)
};

// the strip's column extent in the window: [col_lo, col_hi] with the slack of the P-column groups, and the byte
// offset wb0 of the window's first column (16-byte aligned, may be negative); lo / hi = the tap ranges of the
// strip's first / last column of both channels (tap_range)
__device__ __forceinline__ void stream_strip_extent(int x0, int xl, int lo0, int lo1, int hi0, int hi1, int slack, int bpp,
						    int &col_lo, int &col_hi, int &wb0)
{
	col_lo = min(x0, min(lo0, lo1));
	col_hi = max(xl, max(hi0, hi1));
	// Slack for the shared-sample windows of the P-column groups.  NOT clamped to the image: the TMA
	// unit zero-fills columns outside it, and the clamp-to-edge rule is folded into the weights (a
	// tap that would fall outside lands on the edge column; the zero-filled samples get weight 0).
	col_lo -= slack;
	col_hi += slack;
	wb0 = (col_lo * bpp) & ~15;
}

template <class S, int NCH, int INTERP, int P, int TW, bool ALT, int REPAIR, bool WIDE>
__global__ void __launch_bounds__(2 * TW / P) stream_setup_kernel(const KernelArgs a, void *const out)
{
	constexpr int BPP = NCH * (int)sizeof(S);
	constexpr int T = INTERP == 1 ? 2 : 4;
	constexpr int OFF = INTERP == 2 ? 1 : 0;
	constexpr int NW = P == 1 ? T : T + 1;
	constexpr int NS = P + NW - 1;
	constexpr int NTC = 2 * TW / P;
	constexpr int HALF = TW / P;
	constexpr bool HAS_NARROW = P == 2 || P == 3 || (P == 4 && REPAIR != 1);
	static_assert(INTERP != 0, "None has no weights");
	typedef typename std::conditional<WIDE, double, float>::type A;
	typedef StripCodec<S> Codec;
	__shared__ int s_lo[2], s_hi[2];
	const int tid = threadIdx.x;
	const int W = a.g.width;
	const int x0 = blockIdx.x * TW;
	const int xl = min(x0 + TW, W) - 1;
	if (tid < 4) {
		const int ch = tid & 1, last = tid >> 1;
		int lo, hi;
		tap_range(a.g.x[ch], INTERP, last ? xl : x0, lo, hi);
		if (last) s_hi[ch] = hi; else s_lo[ch] = lo;
	}
	__syncthreads();
	int col_lo, col_hi, wb0;
	stream_strip_extent(x0, xl, s_lo[0], s_lo[1], s_hi[0], s_hi[1], STREAM_COL_SLACK(P), BPP, col_lo, col_hi, wb0);
	const int c = ALT ? (tid & 1) : tid / HALF;	// 0 red, 1 blue (the main kernel's layout)
	const int lt = ALT ? (tid >> 1) : tid - c * HALF;

	StreamColumnState<A, P, NW> st;
	A w[P][4];
	int tap[P][T];	// clamp-to-edge tap columns minus k
	int bmin = INT_MAX;
#pragma unroll
	for (int k = 0; k < P; ++k) {
		// columns past the tile's last one are computed like any other (their results are
		// clipped by the TMA store); past the image the coordinate clamps to W - 1 ...
		double td;
		const int i0 = base_index(a.g.x[c], x0 + lt * P + k, td);
		if constexpr (WIDE) {
			tap_weights_d<INTERP>(td, w[k]);
#pragma unroll
			for (int j = 0; j < 4; ++j)
				w[k][j] = w[k][j] * kWideScale + 0.0;
		} else {
			if (REPAIR) {	// FP64 weights, rounded once (the error bound counts one rounding per weight)
				double wd[4];
				tap_weights_d<INTERP>(td, wd);
#pragma unroll
				for (int j = 0; j < 4; ++j)
					w[k][j] = (float)wd[j];
			} else {
				tap_weights<INTERP>((float)td, w[k]);
			}
#pragma unroll
			for (int j = 0; j < 4; ++j)
				w[k][j] = w[k][j] * Codec::kHScale + 0.f;	// power of two (integer samples are read as subnormals); -0 -> +0: the sign of an all-zero sum must not depend on the fold
		}
		// ... with zero weights, so that they do not bend a warp of the last strip
		if (x0 + lt * P + k > xl)
			w[k][0] = w[k][1] = w[k][2] = w[k][3] = 0;
#pragma unroll
		for (int j = 0; j < T; ++j) {
			tap[k][j] = clampi(i0 - OFF + j, 0, W - 1) - k;
			if (w[k][j] != 0)
				bmin = min(bmin, tap[k][j]);
		}
	}
	if (bmin == INT_MAX)	// no column of this thread is inside the image
		bmin = col_lo;
	// regular: every tap that carries weight sits at shared sample k + j', 0 <= j' < NW, and the
	// NS samples lie inside the window
	bool regular = bmin >= col_lo && bmin + NS - 1 <= col_hi;
	// Float samples: a zero weight does not silence a NaN.  Columns left of the image and right of
	// its 16-byte-aligned row end are zero-filled by the TMA unit, but the bytes between width * BPP
	// and that row end are whatever the caller's pitch padding holds: such windows go the bent way,
	// which clamps its sample offsets to the last column.
	if (is_float_sample<S>::value)
		regular = regular && bmin + NS - 1 <= W - 1;
#pragma unroll
	for (int k = 0; k < P; ++k)
#pragma unroll
		for (int j = 0; j < T; ++j)
			regular = regular && (w[k][j] == 0 || tap[k][j] - bmin <= NW - 1);
	regular = __all_sync(0xffffffffu, regular);
	// The usual thread: no tap clamped, every column's window starts 0 or 1 samples after the group's
	// first sample -- its weights are the tap weights, shifted by that drift.
	bool plain = regular;
#pragma unroll
	for (int k = 0; k < P; ++k) {
		const int d = tap[k][0] - bmin;
		plain = plain && (d == 0 || (NW > T && d == 1));
#pragma unroll
		for (int j = 1; j < T; ++j)
			plain = plain && tap[k][j] == tap[k][0] + j;
	}
	if (plain) {
#pragma unroll
		for (int k = 0; k < P; ++k) {
			const bool drift = tap[k][0] != bmin;
#pragma unroll
			for (int jj = 0; jj < NW; ++jj) {
				const A w0 = jj < T ? w[k][jj] : 0, w1 = jj >= 1 ? w[k][jj - 1] : 0;
				st.wt[k][jj] = drift ? w1 : w0;
			}
		}
	} else {
#pragma unroll
		for (int k = 0; k < P; ++k)
#pragma unroll
			for (int jj = 0; jj < NW; ++jj) {
				A v = 0;
#pragma unroll
				for (int j = 0; j < T; ++j) {
					const int at = regular ? tap[k][j] - bmin : tap[k][j] - tap[k][0];
					v += (w[k][j] != 0 && at == jj) ? w[k][j] : 0;
				}
				st.wt[k][jj] = v;
			}
	}
#pragma unroll
	for (int k = 0; k < P; ++k)
		st.cofs[k] = (tap[k][0] + k - bmin) * BPP;
	st.colbase = bmin * BPP + 2 * c * (int)sizeof(S) - wb0;
	st.cmax = (W - 1 - bmin) * BPP;
	// narrow: regular, and no column group of the warp straddles a drift of the tap window (the usual case: the map
	// drifts one sample every 1 / |scale - 1| columns), so the extra weight is 0 everywhere
	bool narrow = regular && HAS_NARROW;
#pragma unroll
	for (int k = 0; k < P; ++k)
		narrow = narrow && st.wt[k][NW - 1] == 0;
	narrow = __all_sync(0xffffffffu, narrow);
	st.form = narrow ? 2 : regular ? 1 : 0;
	reinterpret_cast<StreamColumnState<A, P, NW> *>(out)[blockIdx.x * NTC + tid] = st;
}

// ---------------------------------------------------------------------------------------------------------
// The exact-repair kernels' slow path, OUT OF LINE: n <= 32 queued samples of a warp, one per lane, through the
// reference's own arithmetic (fix-ca.c:1135-1186, :1204-1320): coordinates, clamp-to-edge taps read from the window
// ring, FP64 in the reference's operation order.  One in ~2000 samples comes here; inlined at the four emit sites
// of the unrolled rows (r02 (B)) its 180 FP64 instructions made the row loop 2300 instructions long -- four times
// the instruction cache -- and the bit-identical 8-bit kernels ran at a third of the FAST ones' speed.
// entry = chunk row << 8 | column in group << 5 | lane.
// ---------------------------------------------------------------------------------------------------------
template <class S, int NCH, int INTERP, int P, int TW>
__device__ __noinline__ void stream_repair_samples(const Geometry *const g, const uint16_t *const entries, const int n, const int c,
							const int x0, const int y_first, const unsigned char *const win, const int ring_rows,
							const int wpitch, const int wb0, unsigned char *const stg)
{
	constexpr int BPP = NCH * (int)sizeof(S);
	constexpr int OUT_PITCH = TW * BPP;
	constexpr int HALF = TW / P;
	const int tid = threadIdx.x, lane = tid & 31;
	__syncwarp();
	if (lane < n) {
		const unsigned e = entries[lane];
		const int elt = (tid & ~31) + (int)(e & 31u) - c * HALF, ek = (int)(e >> 5) & 7, er = (int)(e >> 8);
		const S v = interp_sample<S, INTERP, ExactF64>(*g, c, x0 + elt * P + ek, y_first + er, [&](int row, int col) {
			return *reinterpret_cast<const S *>(win + (row % ring_rows) * wpitch + col * BPP + 2 * c * (int)sizeof(S) - wb0);
		});
		*reinterpret_cast<S *>(stg + er * OUT_PITCH + (elt * P + ek) * BPP + 2 * c * (int)sizeof(S)) = v;
	}
	__syncwarp();
}

// ---------------------------------------------------------------------------------------------------------
// repair_patch_kernel: the second half of the deferred exact-repair form.  Launched behind stream_kernel<..., REPAIR = 2>
// (programmatic dependent launch: it waits for that grid's completion), it recomputes the queued near-tie samples with
// the reference's own arithmetic (interp_sample<ExactF64>: fix-ca.c:1135-1186, :1204-1320; taps gathered from the source
// image through L2) and overwrites them in the destination frame(s).  A group of threads per region (= compute warp of
// the streaming grid), a thread per queued sample.  A region whose warp flagged more samples than its slots hold (an image
// made of exact ties) is recomputed as a whole: slow, and still the reference's bytes.
// ---------------------------------------------------------------------------------------------------------
struct PatchArgs {
	unsigned long long src_frame_stride, dst_frame_stride;	// bytes between the frames of a batch
	unsigned nregions;			// CTAs of the streaming grid x compute warps
	int grid_x, grid_y;			// the streaming grid (strips, segments; z = frames)
	int warps, tw, p;			// compute warps per CTA, strip width, columns per thread
	unsigned lanes;				// threads per region: a power of two, 8 .. 256
	unsigned nframes;			// frames of the batch (grid z of the streaming launch)
	int nfan;				// further destination frames (fan-out / all-gather form)
	unsigned char *fan[STREAM_MAX_FAN];
};
template <class S, int NCH, int INTERP>
__global__ void __launch_bounds__(256) repair_patch_kernel(const __grid_constant__ KernelArgs a, const __grid_constant__ PatchArgs pa)
{
	griddep_launch_dependents();
	griddep_wait();		// the streaming grid has completed: its queue, counts and destination bytes are visible
	const unsigned W = (unsigned)a.g.width, rows = (unsigned)(a.y2 - a.y1);
	auto fix = [&](const unsigned long long frame, const int x, const int y, const int c) {
		const unsigned char *const src = a.src + frame * pa.src_frame_stride;
		const S v = interp_sample<S, INTERP, ExactF64>(a.g, c, x, y, [&](int row, int col) {
			return reinterpret_cast<const S *>(src + (long long)(row - a.src_row0) * a.src_pitch)[(size_t)col * NCH + 2 * c];
		});
		const size_t at = frame * pa.dst_frame_stride + (size_t)((long long)(y - a.dst_row0) * a.dst_pitch) + ((size_t)x * NCH + 2 * c) * sizeof(S);
		*reinterpret_cast<S *>(a.dst + at) = v;
		for (int f = 0; f < pa.nfan; ++f)
			*reinterpret_cast<S *>(pa.fan[f] + at) = v;
	};
	// pa.lanes threads per region, a queued sample each: a sample is ~450 dependent FP64 instructions behind 16 gathered taps
	// (~5 us of latency, hardly any throughput), so the host sizes the groups for ONE turn of this loop (with 8 lanes for
	// the ~13 samples of a 24 MP image's regions the kernel took 40 us instead of 8)
	const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned sl = t % pa.lanes;
	for (unsigned r = t / pa.lanes; r < pa.nregions; r += gridDim.x * (blockDim.x / pa.lanes)) {
		const unsigned n = __ldcg(a.rq_ctl + r);
		if (n <= a.rq_cap) {
			const unsigned long long *const q = a.rq_entries + (size_t)r * a.rq_cap;
			for (unsigned i = sl; i < n; i += pa.lanes) {
				const unsigned long long e = __ldcg(q + i);
				const unsigned long long px = e >> 1, fr = px / W;
				if (fr / rows < pa.nframes)	// (x < width and y1 <= y < y2 by construction)
					fix(fr / rows, (int)(px % W), a.y1 + (int)(fr % rows), (int)(e & 1ull));
			}
		} else {
			// the warp's whole region: rows of its CTA's segment x the 32 * P columns of its lanes, its channel
			const unsigned cta = r / (unsigned)pa.warps, w = r % (unsigned)pa.warps;
			const int bx = (int)(cta % (unsigned)pa.grid_x), by = (int)((cta / (unsigned)pa.grid_x) % (unsigned)pa.grid_y);
			const unsigned long long frame = cta / ((unsigned)pa.grid_x * (unsigned)pa.grid_y);
			const int half = pa.tw / pa.p, tid0 = (int)w * 32;
			const int c = tid0 / half;
			const int xa = bx * pa.tw + (tid0 - c * half) * pa.p, ncol = 32 * pa.p;
			const int ya = a.y1 + by * a.seg_rows, yb = min(ya + a.seg_rows, a.y2);
			for (int i = (int)sl; i < (yb - ya) * ncol; i += (int)pa.lanes) {
				const int x = xa + i % ncol;
				if (x < (int)W)
					fix(frame, x, ya + i / ncol, c);
			}
		}
	}
}

// Dynamic shared memory: [StreamHeader | StreamMeta[D + 1] | window ring (ring_rows x win_pitch) |
//                         staging (3 x CH x TW x BPP)]
// blockDim.x == 2 * TW / P compute threads + 32 (the TMA warp).
//
// tm_win   source image, box = win_pitch bytes x 4 rows     (window ring groups)
// tm_tile  source image, box = TW * BPP bytes x CH rows      (pass-through pixels of a chunk)
// tm_out   destination rows [dst_row0, y2), same box         (finished chunks; clipped at y2 and at the row end)
// fan      further destinations of the same geometry (fan-out / all-gather form), usually none
// All three are 3-D maps (8-byte elements x rows x frames of a batch, blockIdx.z = frame) over rows of
// align16(width * BPP) bytes; row coordinates are relative to src_row0 / dst_row0.
// REPAIR (8-bit samples, Linear / Cubic): the bit-exact form.  The FP32 pipeline runs as in FAST mode (weights
// rounded once from FP64); every output whose FP32 value lies within Codec::kEps -- a proven bound on |FP32 value -
// reference value| -- of a rounding boundary is queued per warp and recomputed with the reference's own FP64
// arithmetic (interp_sample<ExactF64>, taps read from the window ring), 32 queued samples at a time so that the
// FP64 work runs on full warps.  Everything else rounds to the same integer in both arithmetics.
// REPAIR = 2, the deferred form (the default): the same test, but the near-tie samples are only appended to a queue in
// global memory (KernelArgs::rq_*) and repair_patch_kernel (below), launched behind this grid, recomputes them from the
// source image -- the FP64 function, its registers and the queue drains leave the streaming kernel (4 CTAs per SM again).
// WIDE (16-bit samples, Linear / Cubic): the bit-exact form on an FP64 pipeline.  The same separable sums in FP64 on
// raw sample values (fixca_strip.cuh, WideCodec); a sample within WideCodec::kEps of a rounding boundary -- one in
// half a million -- is recomputed in the reference's operation order by its own thread at the end of the chunk.
template <class S, int NCH, int INTERP, int P, int TW, bool ALT = false, int REPAIR = 0, bool WIDE = false>
#ifdef FIXCA_EXP_LB5
__global__ void __launch_bounds__(2 * TW / P + 32, (2 * TW / P) <= 128 && !WIDE ? 5 : 2)	// experiment: 5 resident CTAs (80 registers)
#else
// register budget: 4 resident CTAs for the narrow-pixel FP32 kernels, 3 for their exact-repair forms (at 96 registers they
// spilled 100-150 bytes per thread into the row loop: local-memory traffic on the shared-memory data pipe that bounds these
// kernels, 24 MP RGB8 Cubic EXACT 0.117 ms), 2 for the wide strips and the FP64 pipelines
__global__ void __launch_bounds__(2 * TW / P + 32, (2 * TW / P) <= 128 && !WIDE ? (REPAIR == 1 ? 3 : 4) : 2)
#endif
stream_kernel(const __grid_constant__ KernelArgs a, const __grid_constant__ CUtensorMap tm_win,
	      const __grid_constant__ CUtensorMap tm_tile, const __grid_constant__ CUtensorMap tm_out,
	      const __grid_constant__ StreamFanout fan)
{
	extern __shared__ __align__(128) unsigned char smem[];
#ifdef FIXCA_EXP_TIMING
	const long long t_entry = clock64();
#endif
	constexpr int BPP = NCH * (int)sizeof(S);
	constexpr int OUT_PITCH = TW * BPP;
	constexpr int STAGE_BYTES = STREAM_CH * OUT_PITCH;
	constexpr int T = INTERP == 0 ? 1 : INTERP == 1 ? 2 : 4;	// taps per axis (None: the nearest sample)
	constexpr int NW = P == 1 ? T : T + 1;
	constexpr int NTC = 2 * TW / P;		// compute threads
	constexpr int HALF = TW / P;
	constexpr int CH = STREAM_CH;
	constexpr int NSTG = STREAM_NSTG;
	const int D = a.depth, NF = D + 1;
	static_assert(ALT || HALF % 32 == 0, "a warp must not straddle the two channels");
	static_assert(NTC % 32 == 0, "whole compute warps");
	static_assert(2 * CH <= 32, "one producer lane per (channel, row) of a chunk");
	static_assert(!(WIDE && (REPAIR || INTERP == 0)), "WIDE is its own exact form");
	typedef typename std::conditional<WIDE, double, float>::type A;		// the pipeline's arithmetic type
	typedef typename std::conditional<WIDE, StreamMetaWide, StreamMeta>::type Meta;
	StreamHeader *hdr = reinterpret_cast<StreamHeader *>(smem);
	Meta *meta = reinterpret_cast<Meta *>(smem + a.off_ytab);
	unsigned char *win = smem + a.off_win;
	unsigned char *stage = smem + a.off_out;

	const int tid = threadIdx.x;
	const int W = a.g.width;
	const int x0 = blockIdx.x * TW;
	const int xl = min(x0 + TW, W) - 1;
	const int ya = a.y1 + blockIdx.y * a.seg_rows;
	const int yb = min(ya + a.seg_rows, a.y2);
	const int nchunks = (yb - ya + CH - 1) / CH;
	const int NR = a.ring_rows;
	const int wpitch = a.win_pitch;

	// the source rows the segment's first D + 1 chunks touch, one lane of the TMA warp each (later chunks: the plan's span table)
	const int cj0 = (ya - a.y1) / CH;		// the segment's first chunk in the plan's tables
	const int2 *const span = reinterpret_cast<const int2 *>(a.span_tab) + cj0;	// StreamSpan {lo, hi} records
	int2 sp_lane = make_int2(0, 0);
	if (tid >= NTC && tid - NTC <= STREAM_MAX_D) {
		// (evaluated, not loaded: at CTA start the table is a cold read -- 1-2 k cycles in front of the first TMA request
		// of a one-wave launch; a lane per chunk takes a few hundred)
		const int jc = min(tid - NTC, nchunks - 1);
		int lo, hi, dummy;
		stream_tap_rows<INTERP>(a.g, ya, lo, dummy);
		stream_tap_rows<INTERP>(a.g, min(ya + jc * STREAM_CH + STREAM_CH, yb) - 1, dummy, hi);
		sp_lane = make_int2(lo, hi);
	}

	// ---- one-time: barriers and the strip's column extent ----
	if (tid == 0) {
		for (int i = 0; i < NF; ++i) {
			mbar_init(reinterpret_cast<uint64_t *>(&hdr->full[i]), 2);	// window rows; pass-through tile + the chunk's record
			mbar_init(reinterpret_cast<uint64_t *>(&hdr->done[i]), NTC / 32);	// one arrival per compute warp (warp_arrive)
		}
		fence_mbar_init();
	}
	if (tid < 4) {
		const int ch = tid & 1, last = tid >> 1;
		int lo, hi;
		tap_range(a.g.x[ch], INTERP, last ? xl : x0, lo, hi);
		if (last) hdr->col_hi[ch] = hi; else hdr->col_lo[ch] = lo;
	}
	__syncthreads();
	int col_lo, col_hi, wb0;	// (wb0 may be negative)
	stream_strip_extent(x0, xl, hdr->col_lo[0], hdr->col_lo[1], hdr->col_hi[0], hdr->col_hi[1], STREAM_COL_SLACK(P), BPP, col_lo, col_hi, wb0);
	(void)col_lo; (void)col_hi;
	uint64_t *full = reinterpret_cast<uint64_t *>(hdr->full);
	uint64_t *done = reinterpret_cast<uint64_t *>(hdr->done);

	if (tid >= NTC) {
		// =====================================================================
		// TMA warp (one elected lane): window groups and pass-through tiles in, finished chunks out
		// =====================================================================
		// Which source rows a chunk needs comes from the plan's span table (written long before this launch: no
		// dependency on the previous grid).  The first D + 1 records are fetched by the warp's lanes at once;
		// after that the lane reads one record per chunk, a whole chunk period before it is needed.
		int hi_next = __shfl_sync(0xffffffffu, sp_lane.y, 0);	// last source row of the next chunk to request
		if (tid == NTC) {
			griddep_launch_dependents();
			griddep_wait();		// the only thread of the CTA that reads or writes image memory
			prefetch_tensormap(&tm_win);
			prefetch_tensormap(&tm_tile);
			prefetch_tensormap(&tm_out);
			for (int e = 0; e < fan.n; ++e)
				prefetch_tensormap(&fan.tm[e]);
		}
		const int NRG = NR >> 2;		// ring capacity in 4-row groups
		const int group_bytes = 4 * wpitch;
		const int c0_win = wb0 >> 3, c0_tile = (x0 * BPP) >> 3;
		const int frame = blockIdx.z;		// a batch of equal frames: one grid layer per frame
		int loaded_g = (sp_lane.x >> 2) - 1;	// highest 4-row group already requested (lane 0: before the segment's first row)
		int gslot = (loaded_g + 1) % NRG;	// ring slot of group loaded_g + 1
		int wnf = 0;				// window requests: i % NF
		auto request_window = [&](const int row_hi) {	// the source rows up to row_hi join the ring (the next chunk's)
			const int hi_g = max(row_hi >> 2, loaded_g);
			uint64_t *bar = &full[wnf];
			if (STREAM_DEBUG_BIT(a, 4)) {	// timing experiment: no window loads (the compute warps filter stale shared memory)
				mbar_arrive(bar);
			} else {
				mbar_arrive_expect_tx(bar, (uint32_t)((hi_g - loaded_g) * group_bytes));
				for (int g = loaded_g + 1; g <= hi_g; ++g) {
					tma_load_3d(win + gslot * group_bytes, &tm_win, c0_win, 4 * g - a.src_row0, frame, bar);
					gslot = gslot + 1 == NRG ? 0 : gslot + 1;
				}
			}
			loaded_g = hi_g;
			wnf = wnf + 1 == NF ? 0 : wnf + 1;
		};
		const unsigned char *const meta_src = reinterpret_cast<const unsigned char *>(a.meta_tab) + (size_t)cj0 * sizeof(Meta);
		int tnf = 0, tstg = 0;			// tile requests: i % NF, i % NSTG
		// Chunk i's own pixels -> its staging buffer, its record -> its metadata slot; requested TLEAD chunks ahead.
		// TLEAD = 2 (depth >= 2): the buffer's previous tenant is the chunk whose store was issued a moment ago, so
		// the lane first waits until that store has read the buffer -- it would only be polling the next `done`
		// barrier otherwise -- and the tile has two chunk periods to arrive instead of one (the compute warps spent 9 %
		// of their time waiting for it: an L2 round trip under load is about one chunk period of a narrow strip).
		// The metadata slot's previous tenant (chunk i - NF) was handed over before that store was issued.
		const int TLEAD = a.tile_lead;
		auto request_tile = [&](int i) {
			if (TLEAD == 2)
				bulk_wait_read<0>();
			else
				bulk_wait_read<1>();	// (tenant stored two iterations ago)
			uint64_t *bar = &full[tnf];
			if (STREAM_DEBUG_BIT(a, 4)) {	// timing experiment: the record only
				mbar_arrive_expect_tx(bar, (uint32_t)sizeof(Meta));
			} else {
				mbar_arrive_expect_tx(bar, (uint32_t)(STAGE_BYTES + sizeof(Meta)));
				tma_load_3d(stage + tstg * STAGE_BYTES, &tm_tile, c0_tile, ya + i * CH - a.src_row0, frame, bar);
			}
			bulk_load(&meta[tnf], meta_src + (size_t)i * sizeof(Meta), (uint32_t)sizeof(Meta), bar);
			tnf = tnf + 1 == NF ? 0 : tnf + 1;
			tstg = tstg + 1 == NSTG ? 0 : tstg + 1;
		};
		// (the whole warp walks the first D + 1 records so that lane 0 gets them by shuffle)
		for (int i = 0; i < D && i < nchunks; ++i) {
			if (tid == NTC)
				request_window(hi_next);
			hi_next = __shfl_sync(0xffffffffu, sp_lane.y, min(i + 1, STREAM_MAX_D));
		}
		if (tid != NTC)
			return;
		for (int i = 0; i < TLEAD && i < nchunks; ++i)
			request_tile(i);
		int jnf = 0, jstg = 0, jpar = 0;	// j % NF, j % NSTG, (j / NF) & 1
		for (int j = 0; j < nchunks; ++j) {
			if (j + D < nchunks) {
				request_window(hi_next);
				hi_next = __ldg(&span[min(j + D + 1, nchunks - 1)]).y;
			}
			if (j + TLEAD < nchunks)
				request_tile(j + TLEAD);
#ifdef FIXCA_TUNING
			if (a.debug >> 8)
				mbar_wait_naps(&done[jnf], (uint32_t)jpar, (uint32_t)(a.debug >> 8));
			else
#endif
			mbar_wait_sleepy(&done[jnf], (uint32_t)jpar, 500u);
			if (!STREAM_DEBUG_BIT(a, 8))	// (timing experiment: no stores either)
			tma_store_3d(&tm_out, c0_tile, ya + j * CH - a.dst_row0, frame, stage + jstg * STAGE_BYTES);
			for (int e = 0; e < fan.n; ++e)		// the same chunk into the other frames (peer GPUs, over NVLink)
				tma_store_3d(&fan.tm[e], c0_tile, ya + j * CH - a.dst_row0, frame, stage + jstg * STAGE_BYTES);
			bulk_commit();
			if (++jnf == NF) { jnf = 0; jpar ^= 1; }
			jstg = jstg + 1 == NSTG ? 0 : jstg + 1;
		}
		bulk_wait_all();
		return;
	}

	// =========================================================================
	// compute warps
	// =========================================================================
	// ALT (4-channel formats): even lanes red, odd lanes blue of the same pixel, so that a warp's
	// samples span 32 * BPP / 2 bytes of a row instead of 32 * BPP (half the shared-memory
	// wavefronts for 8- and 16-byte pixels).  Otherwise the channel is uniform per warp.
	const int c = ALT ? (tid & 1) : tid / HALF;	// 0 red, 1 blue
	const int lt = ALT ? (tid >> 1) : tid - c * HALF;
	const int qoff = lt * P * BPP + 2 * c * (int)sizeof(S);

	if constexpr (INTERP == 0) {
		// ---- None (fix-ca.c:1100-1121): dst.c = src[nearest row][nearest column].c as raw sample bytes,
		// exact for every payload; green / alpha are already in the staging buffer (pass-through tile)
		int cofs[P];
#pragma unroll
		for (int k = 0; k < P; ++k)
			cofs[k] = __ldg(a.col_i0 + c * a.col_n + x0 + lt * P + k) * BPP + 2 * c * (int)sizeof(S) - wb0;
		int jnf = 0, jstg = 0, jpar = 0;	// j % NF, j % NSTG, (j / NF) & 1
		for (int j = 0; j < nchunks; ++j) {
			mbar_wait(&full[jnf], (uint32_t)jpar);
			const Meta &m = meta[jnf];
			uint64_t *const done_bar = &done[jnf];
			unsigned char *q = stage + jstg * STAGE_BYTES + qoff;
			if (++jnf == NF) { jnf = 0; jpar ^= 1; }
			jstg = jstg + 1 == NSTG ? 0 : jstg + 1;
			if (!STREAM_DEBUG_BIT(a, 1)) {
				// rows past the band's end (last chunk) repeat valid offsets and are clipped by the store
				S v[CH][P];
#pragma unroll
				for (int r = 0; r < CH; ++r) {
					const unsigned char *prow = win + __float_as_int(m.wy[r][c].x);
#pragma unroll
					for (int k = 0; k < P; ++k)
						v[r][k] = *reinterpret_cast<const S *>(prow + cofs[k]);
				}
#pragma unroll
				for (int r = 0; r < CH; ++r)
#pragma unroll
					for (int k = 0; k < P; ++k)
						*reinterpret_cast<S *>(q + r * OUT_PITCH + k * BPP) = v[r][k];
			}
			warp_arrive(done_bar);
		}
	} else {
	typedef StripCodec<S> Codec;
	typedef typename std::conditional<WIDE, WideCodec<S>, StripCodec<S>>::type LoadCodec;

	// The thread's column set-up (folded horizontal weights, sample offsets, the form its warp runs): the plan's
	// table, one record per strip and compute thread (stream_setup_kernel).
	typedef StreamColumnState<A, P, NW> ColState;
	static_assert(sizeof(ColState) % 16 == 0, "records are read as 16-byte vectors");
	ColState cst;
	{
		const uint4 *const src = reinterpret_cast<const uint4 *>(reinterpret_cast<const ColState *>(a.setup_tab) + (size_t)blockIdx.x * NTC + tid);
		uint4 *const dstv = reinterpret_cast<uint4 *>(&cst);
#pragma unroll
		for (int i = 0; i < (int)(sizeof(ColState) / 16); ++i)
			dstv[i] = __ldg(src + i);
	}
	A (&wt)[P][NW] = cst.wt;
	int (&cofs)[P] = cst.cofs;	// ... relative to colbase
	const int colbase = cst.colbase;	// byte offset of shared sample 0 from the window row start
	const int cmax = cst.cmax;		// bent form: offset of the image's last column (samples past it carry no weight)
	const bool regular = cst.form != 0;

	// last source row this thread has filtered horizontally: the one below the first tap row of the segment's first output row
	int s_done = __ldg(&reinterpret_cast<const Meta *>(a.meta_tab)[cj0].first[c]) - 1;
	// Ring of the last four horizontal rows.  Between chunks the newest row sits in slot 3 and the row p below it in
	// slot 3 - p: the unrolled steady-state rows write slots 0, 1, 2, 3, 0 ... (whole ring turns, so they leave that
	// order behind), the general walk shifts the ring down a slot per source row (12 register moves for four
	// columns: rows off the steady path are few, and ONE copy of their code instead of one per ring phase is a
	// quarter of the instruction-cache footprint -- they ran at ~800 cycles per row, mostly instruction fetch).  The
	// vertical weights come ordered by the distance p (position_weights), so the arithmetic is independent of the slot.
	A hr[4][P];
#pragma unroll
	for (int u = 0; u < 4; ++u)
#pragma unroll
		for (int k = 0; k < P; ++k)
			hr[u][k] = 0;
	// the thread's view of the window ring: every row pointer already carries its column offset
	// (32-bit shared-window addresses: one add / compare / select per row; with generic pointers the
	// compiler kept a second, converted copy of the chain for the ld.shared operands)
	const uint32_t win_c = smem_u32(win) + (uint32_t)colbase;
	const uint32_t win_end = win_c + (uint32_t)(NR * wpitch);
	uint32_t prow = win_c + (uint32_t)(((s_done + 1) % NR) * wpitch);	// shared sample 0 of row s_done + 1

	// ---- REPAIR: this warp's queue of near-tie samples (entry = chunk row << 8 | column in group << 5 | lane) ----
	// (channel-per-warp layouts only: every lane of a warp runs the same code, so the queue state is warp-uniform)
	static_assert(!(REPAIR && ALT), "the exact-repair form needs warp-uniform control flow");
	constexpr int RQ_CAP = 32 + 32 * P;		// < 32 pending + what one output row can add
	const int lane = tid & 31;
	uint16_t *const rq = reinterpret_cast<uint16_t *>(smem + a.off_rq) + (REPAIR == 1 ? (tid >> 5) * RQ_CAP : 0);
	int rq_n = 0;				// warp-uniform
	// ---- deferred form: this warp's region of the queue in global memory ----
	[[maybe_unused]] const unsigned rq_region_id = ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * (unsigned)(NTC / 32) + (unsigned)(tid >> 5);
	[[maybe_unused]] unsigned long long *const rq_region = REPAIR == 2 ? a.rq_entries + (size_t)rq_region_id * a.rq_cap : nullptr;
	[[maybe_unused]] unsigned rq_cnt = 0;	// warp-uniform
	// n <= 32 queued samples, one per lane, recomputed out of line (stream_repair_samples)
	auto repair = [&](const int n, const int y_first, unsigned char *const stg) {
		if constexpr (REPAIR == 1)
			stream_repair_samples<S, NCH, INTERP, P, TW>(&a.g, rq + (rq_n - n), n, c, x0, y_first, win, NR, wpitch, wb0, stg);
		rq_n -= n;
	};
	// the columns `flags` marks in chunk row r join the queue; a full warp's worth is repaired at once
	auto enqueue = [&](const unsigned flags, const int r, const int y_first, unsigned char *const stg) {
#pragma unroll
		for (int k = 0; k < P; ++k) {
			const bool f = (flags >> k) & 1u;
			const unsigned m = __ballot_sync(0xffffffffu, f);
			if (m) {
				if (f)
					rq[rq_n + __popc(m & ((1u << lane) - 1u))] = (uint16_t)((r << 8) | (k << 5) | lane);
				rq_n += __popc(m);
			}
		}
		while (rq_n >= 32)	// (a row of an image full of exact ties queues up to 32 * P samples)
			repair(32, y_first, stg);
	};

	// form 0: bent; 1: regular, NW weights per column; 2: regular and no column group of the warp
	// straddles a drift of the tap window (the usual case: the map drifts one sample every
	// 1 / |scale - 1| columns), so the extra weight is 0 everywhere and T weights / P + T - 1 samples do
	auto run = [&](auto form) {
		constexpr int FORM = decltype(form)::value;
		constexpr bool REG = FORM != 0;
		constexpr int NWV = FORM == 0 ? T : FORM == 1 ? NW : NW - 1;	// weights per column
		constexpr int NSL = REG ? P + NWV - 1 : P * T;			// samples a thread loads per source row
		// ---- state of the chunk in hand (set by begin_chunk) ----
		int j = 0;				// chunk of the segment
		int jnf = 0, jstg = 0, jpar = 0;	// j % NF, j % NSTG, (j / NF) & 1
		const Meta *m = nullptr;
		uint64_t *done_bar = nullptr;
		unsigned char *stg = nullptr;		// this chunk's staging buffer
		unsigned char *q = nullptr;		// ... at the thread's first column, next row to emit
		int y_first = 0;
		int er = 0;				// REPAIR / WIDE: chunk row the next emit writes
		[[maybe_unused]] unsigned wflags = 0;	// WIDE / deferred repair: this thread's near-tie samples of the chunk, bit = chunk row * P + column
		int s_end = 0;
		typedef typename std::remove_reference<decltype(meta[0].wy[0][0])>::type WY;
		const WY *wy = nullptr;			// [row][channel]: stride 2
		const int *lastp = nullptr;
		int next_last = 0;

		// a source row's samples / the P horizontal results from them, in either form
		auto load_row = [&](const uint32_t p, A (&smp)[NSL]) {
			if (REG) {
#pragma unroll
				for (int mm = 0; mm < NSL; ++mm)
					smp[mm] = LoadCodec::load_at(p + mm * BPP);
			} else {
#pragma unroll
				for (int k = 0; k < P; ++k)
#pragma unroll
					for (int jj = 0; jj < T; ++jj)
						smp[k * T + jj] = LoadCodec::load_at(p + (uint32_t)min(cofs[k] + jj * BPP, cmax));
			}
		};
		auto hfilter = [&](const A (&smp)[NSL], A (&out)[P]) {
#pragma unroll
			for (int k = 0; k < P; ++k) {
				A v = wt[k][0] * smp[REG ? k : k * T];
#pragma unroll
				for (int jj = 1; jj < NWV; ++jj)
					v = fma(wt[k][jj], smp[REG ? k + jj : k * T + jj], v);
				out[k] = v;
			}
		};
		// one output row of the P columns from the ring (newest row in slot U) into chunk row `row` at q
		auto vemit = [&](auto slot, const auto &wrow, unsigned char *const qrow, const int row) {
			constexpr int U = decltype(slot)::value;
			if constexpr (WIDE) {
				wflags |= vertical_emit_wide<INTERP, U, P, BPP, WideCodec<S>>(hr, wrow, qrow) << (row * P);
			} else {
				const unsigned fl = vertical_emit<INTERP, U, P, BPP, Codec, REPAIR>(hr, wrow, qrow);
				if constexpr (REPAIR == 1)
					enqueue(fl, row, y_first, stg);
				else if constexpr (REPAIR == 2)
					wflags |= fl << (row * P);	// (CH * P <= 32 bits)
			}
		};
		// general walk -- one source row through the horizontal filter: the ring shifts down a slot, the new row goes to slot 3
		auto hrow = [&]() {
			A smp[NSL];
			load_row(prow, smp);
#pragma unroll
			for (int u = 0; u < 3; ++u)
#pragma unroll
				for (int k = 0; k < P; ++k)
					hr[u][k] = hr[u + 1][k];
			hfilter(smp, hr[3]);
			++s_done;
			prow += (uint32_t)wpitch;
			if (prow == win_end)
				prow = win_c;
		};
		// ... and the output rows the newest row (slot 3) completes
		auto emit = [&]() {
#pragma unroll 1
			while (next_last <= s_done) {
				vemit(std::integral_constant<int, 3>(), *wy, q, er++);
				wy += 2;
				q += OUT_PITCH;
				next_last = *++lastp;
			}
		};
		// source rows up to `upto`, emitting whatever they complete
		auto walk = [&](const int upto) {
#pragma unroll 1
			while (s_done < upto) {
				hrow();
				emit();
			}
		};

		// wait for chunk j's rows, pass-through tile and record; returns whether it can take the unrolled loop: a full
		// chunk whose rows complete on the next CH source rows
#ifdef FIXCA_EXP_TIMING
		long long t_waitacc = 0;	// cycles inside the `full` barrier waits of begin_chunk
#endif
		auto begin_chunk = [&]() -> bool {
#ifdef FIXCA_EXP_TIMING
			const long long tw0 = clock64();
			mbar_wait(&full[jnf], (uint32_t)jpar);
			t_waitacc += clock64() - tw0;
#else
			mbar_wait(&full[jnf], (uint32_t)jpar);
#endif
			m = &meta[jnf];
			done_bar = &done[jnf];
			stg = stage + jstg * STAGE_BYTES;
			q = stg + qoff;
			y_first = ya + j * CH;
			er = 0;
			wflags = 0;
			s_end = m->s_end[c];
			wy = &m->wy[0][c];
			lastp = m->last[c];
			next_last = lastp[0];
			return m->simple[c] && next_last == s_done + 1;
		};
		// hand the chunk over: repairs first (their rows and taps belong to this chunk), then the barrier
		auto end_chunk = [&]() {
			if (REPAIR == 1)	// what is left of the queue
				while (rq_n > 0)
					repair(min(rq_n, 32), y_first, stg);
			if constexpr (REPAIR == 2) {
				// deferred form: the chunk's near-tie samples (6e-4 of all) join this warp's region of the launch's queue in
				// global memory -- plain stores, slots counted per warp (an atomic counter cost every chunk an L2 round trip:
				// with eight warps per CTA some warp flags a sample in nearly every chunk); repair_patch_kernel, launched
				// behind this grid, recomputes them from the source image.  Samples beyond the region's capacity are only
				// counted: the patch kernel then recomputes the warp's whole region.
				static_assert(STREAM_CH * P <= 32, "one flag bit per sample of a thread's chunk");
				unsigned pending = __ballot_sync(0xffffffffu, wflags != 0);
				while (pending) {	// (one turn per flagged sample of the busiest lane: one, as a rule)
					if (wflags) {
						const int b = __ffs((int)wflags) - 1;
						wflags &= wflags - 1;
						const int wr = b / P, wk = b - wr * P;
						const unsigned slot = rq_cnt + __popc(pending & ((1u << lane) - 1u));
						if (slot < a.rq_cap) {
							const unsigned long long row = (unsigned long long)blockIdx.z * (unsigned)(a.y2 - a.y1) + (unsigned)(y_first + wr - a.y1);
							rq_region[slot] = ((row * (unsigned)W + (unsigned)(x0 + lt * P + wk)) << 1) | (unsigned)c;
						}
					}
					rq_cnt += __popc(pending);
					pending = __ballot_sync(0xffffffffu, wflags != 0);
				}
			}
			if constexpr (WIDE) {
				// near-tie samples (2e-6 of all): the reference's own arithmetic (fix-ca.c:1135-1186, :1204-1320),
				// taps read from the window ring, whose rows stay valid until the chunk is handed over
				while (wflags) {
					const int b = __ffs((int)wflags) - 1;
					wflags &= wflags - 1;
					const int wr = b / P, wk = b - wr * P;
					const S v = interp_sample<S, INTERP, ExactF64>(a.g, c, x0 + lt * P + wk, y_first + wr, [&](int row, int col) {
						return *reinterpret_cast<const S *>(win + (row % NR) * wpitch + col * BPP + 2 * c * (int)sizeof(S) - wb0);
					});
					*reinterpret_cast<S *>(stg + wr * OUT_PITCH + (lt * P + wk) * BPP + 2 * c * (int)sizeof(S)) = v;
				}
			}
			// staging writes -> visible to the TMA store the producer issues after this barrier
			warp_arrive(done_bar);
			if (++jnf == NF) { jnf = 0; jpar ^= 1; }
			jstg = jstg + 1 == NSTG ? 0 : jstg + 1;
			++j;
		};
		// Steady state: CH consecutive source rows in, CH output rows out.  UNR rows unrolled (whole ring turns, so
		// slots stay static); the next row's samples are loaded before the current row's arithmetic (the loads never
		// wait on the stores).  Wide column groups unroll one ring turn only: 8 rows of P = 4 are 8 KB of code and
		// "no instruction" became the top stall (profiles/r01_ncu_stream_narrow_A.md).
		auto steady_chunk = [&]() {
			constexpr int UNR = P >= 4 ? 4 : CH;
			static_assert(CH % UNR == 0 && UNR % 4 == 0, "unroll must divide the chunk and cover whole ring turns");
			// (bent warps load and filter row by row: their P * T samples are not double-buffered)
			A smp[REG ? 2 : 1][NSL];
			if (REG)
				load_row(prow, smp[0]);
#pragma unroll 1
			for (int it = 0; it < CH / UNR; ++it) {
#pragma unroll
				for (int u = 0; u < UNR; ++u) {
					uint32_t pnext = prow + (uint32_t)wpitch;
					if (pnext == win_end)
						pnext = win_c;
					// (after the chunk's last row the samples are simply dropped)
					if (REG)
						load_row(pnext, smp[REG ? (u + 1) & 1 : 0]);
					else
						load_row(prow, smp[0]);
					hfilter(smp[REG ? u & 1 : 0], hr[u & 3]);
					switch (u & 3) {	// (static ring slots: u is a constant after unrolling)
					case 0: vemit(std::integral_constant<int, 0>(), wy[2 * u], q + u * OUT_PITCH, er++); break;
					case 1: vemit(std::integral_constant<int, 1>(), wy[2 * u], q + u * OUT_PITCH, er++); break;
					case 2: vemit(std::integral_constant<int, 2>(), wy[2 * u], q + u * OUT_PITCH, er++); break;
					default: vemit(std::integral_constant<int, 3>(), wy[2 * u], q + u * OUT_PITCH, er++); break;
					}
					prow = pnext;
				}
				wy += 2 * UNR;
				q += UNR * OUT_PITCH;
			}
			s_done += CH;
		};

		// The chunk loop is a nest: the inner loop runs consecutive steady-state chunks and touches the ring of
		// horizontal rows through static register slots only; everything else (a segment's first chunk, whose first
		// T - 1 rows only prime the ring; chunks at a discontinuity of the row map; the band's last, partial chunk) takes
		// the general walk outside it.  (As one loop with both paths in its body the compiler moved the 4 x P ring
		// registers between the two paths' assignments around every chunk: 24 of ~90 hand-over instructions per thread.)
		if (nchunks <= 0)
			return;
#ifdef FIXCA_EXP_TIMING
		const long long t_setup = clock64();
#endif
		bool steady = begin_chunk();
#ifdef FIXCA_EXP_TIMING
		const long long t_data = clock64();
		long long t_steady0 = 0, t_wait0 = 0;
		int n_general = 0, simple0 = m->simple[c], prime0 = next_last - s_done;
#endif
		for (;;) {
			if (STREAM_DEBUG_BIT(a, 1)) {	// timing experiment: memory pipeline only (results are wrong)
				s_done = s_end;
				prow = win_c + (uint32_t)(((s_done + 1) % NR) * wpitch);
				end_chunk();
				if (j == nchunks)
					return;
				begin_chunk();
				continue;
			}
			// A full chunk that only lacks its priming rows (a segment's first: T - 1 of them) is primed by the general walk
			// and then enters the one steady-state loop below (a second, inlined copy of the unrolled rows for this case cost
			// the first chunk of every CTA a cold instruction cache: 4-6 k cycles against 2.4 k for a steady chunk).
			if (!steady && m->simple[c] && next_last > s_done) {
				walk(next_last - 1);
				steady = next_last == s_done + 1;
			}
			if (steady) {
#ifdef FIXCA_EXP_TIMING
				// experiment: where a compute warp's time goes in the steady state (one CTA reports)
				long long t_rows = 0, t_end = 0, t_begin = 0;
				int n_st = 0;
				if (!t_steady0) { t_steady0 = clock64(); t_wait0 = t_waitacc; }
#pragma unroll 1
				do {
					const long long t0 = clock64();
					steady_chunk();
					const long long t1 = clock64();
					end_chunk();
					const long long t2 = clock64();
					t_rows += t1 - t0; t_end += t2 - t1; ++n_st;
					if (j == nchunks) {
						if (blockIdx.x == 3 && (blockIdx.y == 1 || blockIdx.y == 20) && (gridDim.z == 1 || blockIdx.z % 16 == 5) && (tid & 31) == 0)
							printf("cta z%d warp %d form %d: %d steady chunks, cycles per chunk: rows %lld, hand-over %lld, wait+entry %lld (of it inside the barrier wait %lld); CTA: set-up %lld, first data +%lld, first chunk +%lld (chunk 0 simple %d, rows to prime %d, general chunks first %d, barrier wait in it %lld), total %lld\n", (int)blockIdx.z, tid >> 5, FORM, n_st,
							       t_rows / n_st, t_end / n_st, t_begin / max(n_st - 1, 1), t_waitacc / max(n_st, 1), t_setup - t_entry, t_data - t_setup, t_steady0 - t_data, simple0, prime0, n_general, t_wait0, clock64() - t_entry);
						return;
					}
					steady = begin_chunk();
					t_begin += clock64() - t2;
				} while (steady);
#else
#pragma unroll 1
				do {
					steady_chunk();
					end_chunk();
					if (j == nchunks)
						return;
					steady = begin_chunk();
				} while (steady);
#endif
				continue;	// (the chunk in hand is not steady: it may only lack priming rows)
			}
			// general chunk: rows of this chunk whose taps were all produced while walking the previous chunk, then the walk
#ifdef FIXCA_EXP_TIMING
			if (!t_steady0) ++n_general;
#endif
			emit();
			walk(s_end);
			end_chunk();
			if (j == nchunks)
				return;
			steady = begin_chunk();
		}
	};
	// form 2 (measured: 100 MP RGB16 Cubic 0.212 -> 0.208 ms, RGBA8 0.068 -> 0.066 ms.  Four-column groups (RGB8): one
	// shared-memory wavefront and four FMAs less per row, 128 x 4K RGB8 Cubic 0.717 -> 0.732 of the HBM peak, Linear
	// 0.670 -> 0.687; the exact-repair form keeps the one regular form -- with three row loops it spills.)
	constexpr bool HAS_NARROW = P == 2 || P == 3 || (P == 4 && REPAIR != 1);
	const bool narrow = cst.form == 2 && !STREAM_DEBUG_BIT(a, 2);	// debug bit 1: A/B runs without the narrow form
	if (HAS_NARROW && narrow)
		run(std::integral_constant<int, HAS_NARROW ? 2 : 1>());
	else if (regular)
		run(std::integral_constant<int, 1>());
	else
		run(std::integral_constant<int, 0>());
	if constexpr (REPAIR == 2)
		if (lane == 0)
			a.rq_ctl[rq_region_id] = rq_cnt;	// (every launch writes every region's count: nothing to reset)
	}	// INTERP != 0
}

} // namespace fixca
