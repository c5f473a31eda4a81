// fixca_stream.cuh -- the streaming form of the FAST (FP32) Linear / Cubic kernel.
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
//   * a CTA owns a TW-column strip and a long run of rows (a "segment"); the
//     per-column weights and the ring of horizontal rows live in registers for
//     the whole segment, so every source row is filtered horizontally once per
//     channel and nothing is re-primed;
//   * rows move through shared memory as a TMA + mbarrier pipeline driven by one
//     producer warp: source rows land in a circular window (each row fetched
//     from global memory once per strip), the pass-through pixels of the next
//     output chunk land in a staging buffer, finished chunks leave with TMA bulk
//     stores; D chunks are in flight ahead of the compute warps;
//   * the producer warp's idle lanes compute the per-row vertical weights of the
//     chunks it prefetches (FP64 coordinates, fix-ca.c:813-820, weights folded
//     into ring-slot order), so the compute warps never touch FP64;
//   * the 8 compute warps (4 red, 4 blue) never meet at a CTA barrier: they wait
//     on "full" mbarriers and arrive on "done" mbarriers.
//
// Arithmetic is identical to strip_kernel (same weights, same FMA order, ring
// phase tied to absolute source rows), so both produce the same bytes.
#pragma once

#include <type_traits>

#include "fixca_strip.cuh"

namespace fixca {

constexpr int STREAM_CH = 8;	// output rows per chunk
constexpr int STREAM_D = 2;	// chunks prefetched ahead of the compute warps
constexpr int STREAM_NF = STREAM_D + 1;		// chunks with live "full"/"done" barriers and metadata
constexpr int STREAM_NSTG = STREAM_D + 2;	// staging buffers: D loading, 1 computing, 1 draining

struct StreamMeta {
	float4 wy[2][STREAM_CH];	// vertical weights per output row, ring-slot order, pre-scaled by 1/max
	int    last[2][STREAM_CH + 1];	// highest tap row of each output row (INT_MAX after the chunk's last row)
	int    s_end[2];		// = last[c][nrows - 1]
	int    nrows;
	int    pad;
};

struct StreamHeader {
	unsigned long long full[STREAM_NF];
	unsigned long long done[STREAM_NF];
	int col_lo[2], col_hi[2];
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
	asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Dynamic shared memory: [StreamHeader | StreamMeta[NF] | window ring (ring_rows x win_pitch) |
//                         staging (NSTG x CH x TW x BPP)]
// blockDim.x == 2 * TW / P compute threads + 32 (the producer warp).
template <class S, int NCH, int INTERP, int P, int TW>
__global__ void __launch_bounds__(2 * TW / P + 32) stream_kernel(const __grid_constant__ KernelArgs a)
{
	extern __shared__ __align__(128) unsigned char smem[];
	constexpr int BPP = NCH * (int)sizeof(S);
	constexpr int OUT_PITCH = TW * BPP;
	constexpr int STAGE_BYTES = STREAM_CH * OUT_PITCH;
	constexpr int T = INTERP == 1 ? 2 : 4;
	constexpr int OFF = INTERP == 1 ? 0 : 1;
	constexpr int NW = P == 1 ? T : T + 1;
	constexpr int NS = P + NW - 1;
	constexpr int NTC = 2 * TW / P;		// compute threads
	constexpr int HALF = TW / P;
	constexpr int CH = STREAM_CH, NF = STREAM_NF, NSTG = STREAM_NSTG, D = STREAM_D;
	static_assert(HALF % 32 == 0, "a warp must not straddle the two channels");
	static_assert(2 * CH <= 32, "one producer lane per (channel, row) of a chunk");
	typedef StripCodec<S> Codec;

	StreamHeader *hdr = reinterpret_cast<StreamHeader *>(smem);
	StreamMeta *meta = reinterpret_cast<StreamMeta *>(smem + a.off_ytab);
	unsigned char *win = smem + a.off_win;
	unsigned char *stage = smem + a.off_out;

	const int tid = threadIdx.x;
	const int W = a.g.width, H = a.g.height;
	const int x0 = blockIdx.x * TW;
	const int xl = min(x0 + TW, W) - 1;
	const int ya = a.y1 + blockIdx.y * a.seg_rows;
	const int yb = min(ya + a.seg_rows, a.y2);
	const int nchunks = (yb - ya + CH - 1) / CH;
	const int NR = a.ring_rows;
	const int wpitch = a.win_pitch;

	// ---- one-time: barriers and the strip's column extent ----
	if (tid == 0) {
#pragma unroll
		for (int i = 0; i < NF; ++i) {
			mbar_init(reinterpret_cast<uint64_t *>(&hdr->full[i]), 1);
			mbar_init(reinterpret_cast<uint64_t *>(&hdr->done[i]), NTC);
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
	int col_lo = min(x0, min(hdr->col_lo[0], hdr->col_lo[1]));
	int col_hi = max(xl, max(hdr->col_hi[0], hdr->col_hi[1]));
	col_lo = max(col_lo - P, 0);
	col_hi = min(col_hi + P, W - 1);
	const int wb0 = (col_lo * BPP) & ~15;
	const int wbytes = (((col_hi + 1) * BPP + 15) & ~15) - wb0;
	const int tile_bytes = ((xl - x0 + 1) * BPP + 15) & ~15;
	uint64_t *full = reinterpret_cast<uint64_t *>(hdr->full);
	uint64_t *done = reinterpret_cast<uint64_t *>(hdr->done);

	if (tid >= NTC) {
		// =====================================================================
		// producer warp
		// =====================================================================
		const int lane = tid - NTC;
		int loaded_hi;		// highest source row already requested
		{
			double td;
			const int fr = max(base_index(a.g.y[0], ya, td) - OFF, 0);
			const int fb = max(base_index(a.g.y[1], ya, td) - OFF, 0);
			loaded_hi = min(fr, fb) - 1;
		}
		auto produce = [&](int i) {
			const int y_first = ya + i * CH;
			const int nr = min(CH, yb - y_first);
			StreamMeta &m = meta[i % NF];
			if (lane < 2 * CH) {
				const int ch = lane / CH, r = lane - ch * CH;
				if (r < nr) {
					double td;
					const int i0 = base_index(a.g.y[ch], y_first + r, td);
					float w[4];
					tap_weights<INTERP>((float)td, w);
					float slot[4] = {0.f, 0.f, 0.f, 0.f};
					int last = 0;
#pragma unroll
					for (int j = 0; j < T; ++j) {
						const int q = clampi(i0 - OFF + j, 0, H - 1);
						last = q;
#pragma unroll
						for (int s = 0; s < 4; ++s)
							slot[s] += ((q & 3) == s) ? w[j] * Codec::kInvMax : 0.f;
					}
					m.wy[ch][r] = make_float4(slot[0], slot[1], slot[2], slot[3]);
					m.last[ch][r] = last;
					if (r == nr - 1) {
						m.last[ch][nr] = INT_MAX;
						m.s_end[ch] = last;
					}
				}
			}
			if (lane == 0)
				m.nrows = nr;
			int hi;
			{
				double td;
				const int hr_ = min(base_index(a.g.y[0], y_first + nr - 1, td) + T - 1 - OFF, H - 1);
				const int hb_ = min(base_index(a.g.y[1], y_first + nr - 1, td) + T - 1 - OFF, H - 1);
				hi = max(max(hr_, hb_), loaded_hi);
			}
			const int n_new = hi - loaded_hi;
			// the staging buffer's previous tenant (chunk i - NSTG) must have been read out
			bulk_wait_read<1>();
			__syncwarp();
			uint64_t *bar = &full[i % NF];
			if (lane == 0)
				mbar_arrive_expect_tx(bar, (uint32_t)(n_new * wbytes + nr * tile_bytes));
			__syncwarp();
			for (int r = lane; r < n_new; r += 32) {
				const int q = loaded_hi + 1 + r;
				bulk_load(win + (q % NR) * wpitch,
					  a.src + (long long)(q - a.src_row0) * a.src_pitch + wb0, (uint32_t)wbytes, bar);
			}
			unsigned char *st = stage + (i % NSTG) * STAGE_BYTES;
			const unsigned char *t = a.src + (long long)(y_first - a.src_row0) * a.src_pitch + (long long)x0 * BPP;
			for (int r = lane; r < nr; r += 32)
				bulk_load(st + r * OUT_PITCH, t + (long long)r * a.src_pitch, (uint32_t)tile_bytes, bar);
			loaded_hi = hi;
		};
		for (int i = 0; i < D && i < nchunks; ++i)
			produce(i);
		for (int j = 0; j < nchunks; ++j) {
			if (j + D < nchunks)
				produce(j + D);
			mbar_wait(&done[j % NF], (uint32_t)((j / NF) & 1));
			const int y_first = ya + j * CH;
			const int nr = min(CH, yb - y_first);
			const unsigned char *st = stage + (j % NSTG) * STAGE_BYTES;
			unsigned char *g = a.dst + (long long)(y_first - a.dst_row0) * a.dst_pitch + (long long)x0 * BPP;
			for (int r = lane; r < nr; r += 32)
				bulk_store(g + (long long)r * a.dst_pitch, st + r * OUT_PITCH, (uint32_t)tile_bytes);
			bulk_commit();
		}
		bulk_wait_all();
		return;
	}

	// =========================================================================
	// compute warps
	// =========================================================================
	const int c = tid / HALF;		// 0 red, 1 blue: uniform per warp
	const int lt = tid - c * HALF;

	float wt[P][NW];
	int cidx[P];
	int colbase;
	bool regular;
	{
		int idx0[P];
		float w[P][4];
		int bmin = INT_MAX;
#pragma unroll
		for (int k = 0; k < P; ++k) {
			const int x = min(x0 + lt * P + k, xl);
			double td;
			cidx[k] = base_index(a.g.x[c], x, td);
			tap_weights<INTERP>((float)td, w[k]);
			idx0[k] = cidx[k] - OFF - k;
			bmin = min(bmin, idx0[k]);
		}
		regular = bmin >= col_lo && bmin + NS - 1 <= col_hi;
#pragma unroll
		for (int k = 0; k < P; ++k)
			regular = regular && (idx0[k] - bmin + T - 1 <= NW - 1);
		regular = __all_sync(0xffffffffu, regular);
#pragma unroll
		for (int k = 0; k < P; ++k)
#pragma unroll
			for (int j = 0; j < NW; ++j) {
				if (regular) {
					const int d = j - (idx0[k] - bmin);
					float v = 0.f;
#pragma unroll
					for (int m = 0; m < T; ++m)
						v = (d == m) ? w[k][m] : v;
					wt[k][j] = v;
				} else {
					wt[k][j] = j < T ? w[k][j] : 0.f;
				}
			}
		colbase = bmin * BPP + 2 * c * (int)sizeof(S) - wb0;
	}
	const int choff = 2 * c * (int)sizeof(S) - wb0;

	int s_done;	// last source row this thread has filtered horizontally
	{
		double td;
		s_done = max(base_index(a.g.y[c], ya, td) - OFF, 0) - 1;
	}
	float hr[4][P];
#pragma unroll
	for (int u = 0; u < 4; ++u)
#pragma unroll
		for (int k = 0; k < P; ++k)
			hr[u][k] = 0.f;
	const unsigned char *prow = win + ((s_done + 1) % NR) * wpitch;	// row s_done + 1
	const unsigned char *const win_end = win + NR * wpitch;
	const int qoff = lt * P * BPP + 2 * c * (int)sizeof(S);

	auto run = [&](auto fast_path) {
		constexpr bool FAST = decltype(fast_path)::value;
		for (int j = 0; j < nchunks; ++j) {
			mbar_wait(&full[j % NF], (uint32_t)((j / NF) & 1));
			const StreamMeta &m = meta[j % NF];
			const int s_end = m.s_end[c];
			const float4 *wy = m.wy[c];
			const int *lastp = m.last[c];
			int next_last = lastp[0];
			unsigned char *q = stage + (j % NSTG) * STAGE_BYTES + qoff;

#define FIXCA_EMIT()                                                                                          \
	do {                                                                                                  \
		const float4 w_ = *wy++;                                                                      \
		_Pragma("unroll") for (int k = 0; k < P; ++k)                                                 \
		{                                                                                             \
			const float v_ = __saturatef(fmaf(w_.w, hr[3][k], fmaf(w_.z, hr[2][k],                \
							 fmaf(w_.y, hr[1][k], w_.x * hr[0][k]))));            \
			Codec::store(q + k * BPP, v_);                                                        \
		}                                                                                             \
		q += OUT_PITCH;                                                                               \
		next_last = *++lastp;                                                                         \
	} while (0)

#define FIXCA_STEP(U)                                                                                         \
	do {                                                                                                  \
		if (FAST) {                                                                                   \
			float smp[NS];                                                                        \
			_Pragma("unroll") for (int mm = 0; mm < NS; ++mm)                                     \
				smp[mm] = Codec::load(prow + colbase + mm * BPP);                             \
			_Pragma("unroll") for (int k = 0; k < P; ++k)                                         \
			{                                                                                     \
				float v = wt[k][0] * smp[k];                                                  \
				_Pragma("unroll") for (int jj = 1; jj < NW; ++jj)                             \
					v = fmaf(wt[k][jj], smp[k + jj], v);                                  \
				hr[U][k] = v;                                                                 \
			}                                                                                     \
		} else {                                                                                      \
			_Pragma("unroll") for (int k = 0; k < P; ++k)                                         \
			{                                                                                     \
				float v = 0.f;                                                                \
				_Pragma("unroll") for (int jj = 0; jj < T; ++jj)                              \
				{                                                                             \
					const int ix = clampi(cidx[k] - OFF + jj, 0, W - 1);                  \
					v = fmaf(wt[k][jj], Codec::load(prow + ix * BPP + choff), v);         \
				}                                                                             \
				hr[U][k] = v;                                                                 \
			}                                                                                     \
		}                                                                                             \
		++s_done;                                                                                     \
		prow += wpitch;                                                                               \
		_Pragma("unroll 1") while (next_last <= s_done) FIXCA_EMIT();                                 \
	} while (0)

			// rows of this chunk whose taps were all produced while walking the previous chunk
#pragma unroll 1
			while (next_last <= s_done)
				FIXCA_EMIT();
#pragma unroll 1
			while (s_done < s_end) {
				switch ((s_done + 1) & 3) {	// ring slot of the next source row: static per case
				case 0:
					FIXCA_STEP(0);
					if (s_done >= s_end) break;
				case 1:
					FIXCA_STEP(1);
					if (s_done >= s_end) break;
				case 2:
					FIXCA_STEP(2);
					if (s_done >= s_end) break;
				default:
					FIXCA_STEP(3);
					if (prow == win_end)	// ring_rows % 4 == 0: the ring only wraps after slot 3
						prow = win;
				}
			}
#undef FIXCA_STEP
#undef FIXCA_EMIT
			// staging writes -> visible to the TMA store the producer issues after this barrier
			fence_proxy_async_smem();
			mbar_arrive(&done[j % NF]);
		}
	};
	if (regular)
		run(std::true_type());
	else
		run(std::false_type());
}

} // namespace fixca
