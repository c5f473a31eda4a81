// fixca_api.cu -- the C ABI of include/fixca_cuda.h: argument checking, the
// geometry prologue of fix_ca_region (fix-ca.c:1033-1045), kernel selection and
// tile planning, and the host region driver that replaces the middle of
// fix_ca() (fix-ca.c:366-377): pinned staging, H2D / kernel / D2H pipelined by
// row chunks on CUDA streams, one worker per GPU for row-banded multi-GPU runs.
//
// No CPU compute path exists here: without a usable GPU every compute entry
// point returns FIXCA_ERR_NO_DEVICE / FIXCA_ERR_CUDA.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/fixca_cuda.h"
#include "fixca_internal.h"

using namespace fixca;

// ---------------------------------------------------------------------------
// errors, bookkeeping
// ---------------------------------------------------------------------------
static thread_local char tl_error[512] = "";
static thread_local char tl_kernel[96] = "";
static std::atomic<long> g_launches{0};

static int fail(int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(tl_error, sizeof tl_error, fmt, ap);
	va_end(ap);
	return code;
}

#define CUDA_TRY(expr)                                                                          \
	do {                                                                                    \
		cudaError_t e_ = (expr);                                                        \
		if (e_ != cudaSuccess)                                                          \
			return fail(e_ == cudaErrorMemoryAllocation ? FIXCA_ERR_NOMEM : FIXCA_ERR_CUDA, \
				    "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

// Progress callback: per calling thread (the plug-in installs it on, and calls from, its only thread,
// fix-ca.c:1022-1023; concurrent callers on other threads neither see nor race on it).
static thread_local fixca_progress_fn g_progress = nullptr;
static thread_local void *g_progress_user = nullptr;

// ---------------------------------------------------------------------------
// tuning switches: read from the environment once
// ---------------------------------------------------------------------------
static std::mutex g_tuning_mu;
static Tuning g_tuning;
static std::atomic<int> g_tuning_gen{0};	// 0 = not read yet

static int env_int(const char *name, int dflt)
{
	const char *s = getenv(name);
	return (s && *s) ? atoi(s) : dflt;
}

static void read_tuning_locked()
{
	Tuning t;
	const char *e = getenv("FIXCA_FAST_KERNEL");
	t.fast_kernel = (e && !strcmp(e, "strip")) ? 2 : 3;
	e = getenv("FIXCA_NONE_KERNEL");
	t.none_tiled = e && !strcmp(e, "tiled");
	e = getenv("FIXCA_EXACT_KERNEL");
	t.exact_tiled = e && !strcmp(e, "tiled");
	t.exact_inline = e && !strcmp(e, "inline");
	t.strip_tw128 = env_int("FIXCA_STRIP_TW", 0) == 128;
	t.stream_noalt = env_int("FIXCA_STREAM_NOALT", 0) != 0;
	t.tile_h = env_int("FIXCA_TILE_H", 0);
	t.tile_ctas = env_int("FIXCA_TILE_CTAS", 3);
	t.stream_ctas = env_int("FIXCA_STREAM_CTAS", 0);
	t.stream_depth = env_int("FIXCA_STREAM_DEPTH", 0);
	t.stream_segs = env_int("FIXCA_STREAM_SEGS", 0);
	t.stream_waves = env_int("FIXCA_STREAM_WAVES", 0);
	t.stream_tlead = env_int("FIXCA_STREAM_TLEAD", 0);
	t.stream_debug = env_int("FIXCA_STREAM_DEBUG", 0);
	t.no_pdl = env_int("FIXCA_NO_PDL", 0);
	t.verbose = env_int("FIXCA_VERBOSE", 0);
	t.chunk_mb = env_int("FIXCA_CHUNK_MB", 0);
	t.no_chunk_ramp = env_int("FIXCA_CHUNK_RAMP", 1) == 0;
	t.copy_threads = env_int("FIXCA_COPY_THREADS", 0);
	e = getenv("FIXCA_PRECISION");
	t.precision_fast = e && (e[0] == 'f' || e[0] == 'F');
	t.generation = g_tuning_gen.load() + 1;
	g_tuning = t;
	g_tuning_gen.store(t.generation);
}

const Tuning &fixca::tuning()
{
	if (!g_tuning_gen.load()) {
		std::lock_guard<std::mutex> lock(g_tuning_mu);
		if (!g_tuning_gen.load())
			read_tuning_locked();
	}
	return g_tuning;
}

extern "C" void fixca_cuda_reload_tuning(void)
{
	std::lock_guard<std::mutex> lock(g_tuning_mu);
	read_tuning_locked();
}

static void pinned_pool_release();
static thread_local double tl_last_call_ms;

// ---------------------------------------------------------------------------
// format and geometry
// ---------------------------------------------------------------------------
struct Format {
	SampleKind kind;
	int sample_bytes, nch, bpp;
};

static int parse_format(int bytes, int bpc, Format &f)
{
	switch (bpc) {
	case 1:  f.kind = SK_U8;  f.sample_bytes = 1; break;
	case 2:  f.kind = SK_U16; f.sample_bytes = 2; break;
	case 4:  f.kind = SK_U32; f.sample_bytes = 4; break;
	case 8:  f.kind = SK_U64; f.sample_bytes = 8; break;
	case -4: f.kind = SK_F32; f.sample_bytes = 4; break;
	case -8: f.kind = SK_F64; f.sample_bytes = 8; break;
	case -2: f.kind = SK_F16; f.sample_bytes = 2; break;	// extension: the reference's commented-out half branch
	case FIXCA_BPC_U15: f.kind = SK_U15; f.sample_bytes = 2; break;	// extension: 15-bit unsigned in 16-bit storage
	default:
		return fail(FIXCA_ERR_FORMAT, "unsupported bpc %d (the reference handles 1,2,4,8,-4,-8, and -2 in commented-out code; 15 = u15 is this library's extension; fix-ca.c:688-707)", bpc);
	}
	if (bytes == 3 * f.sample_bytes)
		f.nch = 3;
	else if (bytes == 4 * f.sample_bytes)
		f.nch = 4;
	else
		return fail(FIXCA_ERR_FORMAT, "bytes per pixel %d is neither RGB nor RGBA of %d-byte samples", bytes, f.sample_bytes);
	f.bpp = bytes;
	return FIXCA_OK;
}

// fix-ca.c:1033-1045
static int make_geometry(int width, int height, const fixca_params *p, Geometry &g)
{
	if (p->interpolation < 0 || p->interpolation > 2)
		return fail(FIXCA_ERR_INTERP, "interpolation %d outside 0..2", p->interpolation);
	// NaN / infinite amounts, shifts or lens, and lens coordinates whose (int) cast or width - xc would
	// overflow, are undefined behaviour in the reference (fix-ca.c:1033-1045, :801): rejected here
	const double all[8] = {p->blue, p->red, p->lens_x, p->lens_y, p->x_blue, p->x_red, p->y_blue, p->y_red};
	for (double v : all)
		if (!(v == v) || v > 1.0e300 || v < -1.0e300)
			return fail(FIXCA_ERR_ARG, "non-finite parameter (blue %g red %g lens %g,%g shifts %g %g %g %g)",
				    p->blue, p->red, p->lens_x, p->lens_y, p->x_blue, p->x_red, p->y_blue, p->y_red);
	if (p->lens_x > 1073741824.0 || p->lens_x < -1073741824.0 || p->lens_y > 1073741824.0 || p->lens_y < -1073741824.0)
		return fail(FIXCA_ERR_ARG, "lens centre %g,%g outside +-2^30", p->lens_x, p->lens_y);
	const int xc = (int)p->lens_x, yc = (int)p->lens_y;
	int m = xc >= yc ? xc : yc;
	if (width - xc > m) m = width - xc;
	if (height - yc > m) m = height - yc;
	const double den_blue = (double)m + p->blue, den_red = (double)m + p->red;
	if (den_blue == 0.0 || den_red == 0.0 || den_blue != den_blue || den_red != den_red)
		return fail(FIXCA_ERR_DEGENERATE,
			    "max_dim + amount == 0 (max_dim %d, blue %g, red %g): the scale is infinite and the reference indexes out of bounds",
			    m, p->blue, p->red);
	const double s_blue = (double)m / den_blue, s_red = (double)m / den_red;
	g.width = width;
	g.height = height;
	g.interp = p->interpolation;
	g.x[CH_RED]  = Axis{xc, width, s_red, p->x_red};
	g.x[CH_BLUE] = Axis{xc, width, s_blue, p->x_blue};
	g.y[CH_RED]  = Axis{yc, height, s_red, p->y_red};
	g.y[CH_BLUE] = Axis{yc, height, s_blue, p->y_blue};
	auto ok = [](double s) { return s > 0.0 && s <= 1.0e300; };
	auto fin = [](double v) { return v == v && v > -1.0e300 && v < 1.0e300; };
	g.monotone = ok(s_blue) && ok(s_red) && fin(p->x_red) && fin(p->x_blue) && fin(p->y_red) && fin(p->y_blue);
	return FIXCA_OK;
}

// Inclusive source-row range read by output rows [y1, y2).
static void source_rows(const Geometry &g, int y1, int y2, int &lo, int &hi)
{
	if (g.monotone) {
		span_needed(g.y[CH_RED], g.y[CH_BLUE], g.interp, y1, y2 - 1, lo, hi);
		return;
	}
	lo = y1;
	hi = y2 - 1;
	for (int y = y1; y < y2; ++y)
		for (int c = 0; c < 2; ++c) {
			int l, h;
			tap_range(g.y[c], g.interp, y, l, h);
			lo = std::min(lo, l);
			hi = std::max(hi, h);
		}
}

// ---------------------------------------------------------------------------
// planning
// ---------------------------------------------------------------------------
struct Plan {
	const KernelEntry *k = nullptr;
	KernelArgs args;
	dim3 grid, block;
	size_t smem = 0;
	// stream kernel: TMA tensor maps of the source (window groups, pass-through tiles) and destination
	CUtensorMap tm_win, tm_tile, tm_out;
	StreamFanout fan;	// further destinations of a fan-out launch (fan.n = 0 otherwise)
	void *fan_dst[STREAM_MAX_FAN];	// their base pointers (row dst_row0), as requested
	int src_rows_avail = 0;	// rows of the source band present at args.src
	// stream kernel: the per-chunk tables args.meta_tab / args.span_tab point into (device memory, shared by every
	// plan of the same band, y axes and kernel family; freed with the last plan that holds them)
	std::shared_ptr<void> tables, col_tables;
	// deferred exact-repair form: size of the launch's queue of near-tie samples (counts, then args.rq_cap entries per
	// region); the memory belongs to the launching thread's (device, stream) pair and is bound at launch (repair_queue())
	size_t rq_counts_bytes = 0, rq_bytes = 0;
	// a batch of equal frames in one launch (stream kernels: grid.z = frame, 3-D tensor maps)
	int nframes = 1;
	size_t src_frame_stride = 0, dst_frame_stride = 0;
};

// frames of a batch: frame i starts src_stride / dst_stride bytes after frame i - 1; fan-out: further destination
// frames (same pitch and first row as d_dst) every finished chunk is stored into as well
struct Batch {
	int nframes = 1;
	size_t src_stride = 0, dst_stride = 0;
	int nfan = 0;
	void *fan[STREAM_MAX_FAN] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// FAST arithmetic is taken for the formats that have it; every other Linear / Cubic call computes EXACT
static bool wants_fast(const Format &f, unsigned flags)
{
	return (flags & FIXCA_PRECISION_MASK) == FIXCA_PRECISION_FAST &&
	       (f.kind == SK_U8 || f.kind == SK_U16 || f.kind == SK_F32 || f.kind == SK_F16 || f.kind == SK_U15);
}

static const KernelEntry *pick_kernel(const Format &f, int interp, unsigned flags, bool tiled)
{
	if (interp == 0)
		return lookup_none(f.sample_bytes, f.nch, tiled);
	return wants_fast(f, flags) ? lookup_fast(f.kind, f.nch, interp, tiled) : lookup_exact(f.kind, f.nch, interp, tiled);
}

static int g_smem_optin[64];	// per device, 0 = not queried

static int smem_limit(int dev)
{
	if (dev < 0 || dev >= 64)
		return 48 * 1024;
	if (!g_smem_optin[dev]) {
		int v = 0;
		if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || v <= 0)
			v = 48 * 1024;
		g_smem_optin[dev] = v;
	}
	return g_smem_optin[dev];
}

// Largest source window (bytes per row, rows) any tile of a th-row grid needs.
// `slack` widens every tile's column range by that many pixels per side (strip kernel:
// the shared-sample windows of its P-column groups), clamped to the image.
static void window_extent(const Geometry &g, int bpp, int y1, int y2, int tw, int th, int slack, int &max_wbytes, int &max_rows)
{
	max_wbytes = 0;
	max_rows = 0;
	for (int x0 = 0; x0 < g.width; x0 += tw) {
		const int xl = std::min(x0 + tw, g.width) - 1;
		int lo, hi;
		span_needed(g.x[CH_RED], g.x[CH_BLUE], g.interp, x0, xl, lo, hi);
		lo = std::max(lo - slack, 0);
		hi = std::min(hi + slack, g.width - 1);
		const int b0 = (lo * bpp) & ~15, b1 = ((hi + 1) * bpp + 15) & ~15;
		max_wbytes = std::max(max_wbytes, b1 - b0);
	}
	for (int y0 = y1; y0 < y2; y0 += th) {
		const int yl = std::min(y0 + th, y2) - 1;
		int lo, hi;
		span_needed(g.y[CH_RED], g.y[CH_BLUE], g.interp, y0, yl, lo, hi);
		max_rows = std::max(max_rows, hi - lo + 1);
	}
}

// Widest window (bytes) any strip of the streaming kernel needs; the slack is NOT clamped to the image
// (the kernel's TMA tiles are zero-filled outside it).
static void window_extent_stream(const Geometry &g, int bpp, int tw, int slack, int &max_wbytes)
{
	max_wbytes = 0;
	for (int x0 = 0; x0 < g.width; x0 += tw) {
		const int xl = std::min(x0 + tw, g.width) - 1;
		int lo, hi;
		span_needed(g.x[CH_RED], g.x[CH_BLUE], g.interp, x0, xl, lo, hi);
		lo -= slack;
		hi += slack;
		const int b0 = (lo * bpp) & ~15, b1 = ((hi + 1) * bpp + 15) & ~15;
		max_wbytes = std::max(max_wbytes, b1 - b0);
	}
}

static int g_sm_count[64];	// per device, 0 = not queried

static int sm_count(int dev)
{
	if (dev < 0 || dev >= 64)
		return 148;
	if (!g_sm_count[dev]) {
		int v = 0;
		if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
			v = 148;
		g_sm_count[dev] = v;
	}
	return g_sm_count[dev];
}

struct Plan;
// Streaming kernel: every CTA takes one tw-column strip and one segment of rows; the grid is sized
// so that all CTAs are resident at once (strips x segments ~ SMs x CTAs per SM).  Fills the
// kernel-specific parts of `pl` and returns true, or returns false when the window ring for D + 1
// chunks does not fit in shared memory.
static bool plan_stream(const KernelEntry *k, const Format &f, const Geometry &g, int y1, int y2, int dev, int limit, Plan &pl);

static int make_plan_uncached(const Format &f, const Geometry &g, const void *d_src, size_t src_pitch, int src_row0, int src_rows,
			      void *d_dst, size_t dst_pitch, int dst_row0, int y1, int y2, unsigned flags, int dev, Plan &pl,
			      const Batch &batch)
{
	pl = Plan();
	memset(&pl.fan, 0, sizeof pl.fan);
	pl.fan.n = batch.nfan;	// taken by the streaming kernels only; the caller loops over destinations otherwise
	for (int i = 0; i < STREAM_MAX_FAN; ++i)
		pl.fan_dst[i] = i < batch.nfan ? batch.fan[i] : nullptr;
	pl.nframes = batch.nframes;
	pl.src_frame_stride = batch.src_stride;
	pl.dst_frame_stride = batch.dst_stride;
	KernelArgs &a = pl.args;
	memset(&a, 0, sizeof a);
	a.src = (const unsigned char *)d_src;
	a.dst = (unsigned char *)d_dst;
	a.src_pitch = (long long)src_pitch;
	a.dst_pitch = (long long)dst_pitch;
	a.src_row0 = src_row0;
	a.dst_row0 = dst_row0;
	a.y1 = y1;
	a.y2 = y2;
	a.g = g;
	pl.src_rows_avail = src_rows;

	// The TMA kernels store whole 16-byte units, i.e. up to 15 bytes past width * bpp of every destination row.
	// Harmless in row padding, fatal when the rows are views into a wider buffer: taken only when the rows are whole
	// 16-byte units or the caller declares the padding scratch (FIXCA_PADDING_SCRATCH; the host drivers' own pitched
	// staging buffers always are)
	const bool tight_ok = (flags & FIXCA_PADDING_SCRATCH) || ((size_t)g.width * f.bpp) % 16 == 0;
	const bool tma_ok = tight_ok && g.monotone && (src_pitch % 16 == 0) && (dst_pitch % 16 == 0) &&
			    ((uintptr_t)d_src % 16 == 0) && ((uintptr_t)d_dst % 16 == 0) &&
			    (batch.nframes <= 1 || (batch.src_stride % 16 == 0 && batch.dst_stride % 16 == 0));
	bool want_tiled = tma_ok && !(flags & FIXCA_FORCE_DIRECT);
	// only the streaming kernels take a batch in one launch; every other plan is for one frame and the
	// caller loops (pl.nframes tells which)
	struct OneFrame { Plan &p; ~OneFrame() { if (!p.k || !p.k->stream) { p.nframes = 1; p.src_frame_stride = p.dst_frame_stride = 0; p.fan.n = 0; } } } one_frame{pl};

	if (want_tiled) {
		const KernelEntry *k = pick_kernel(f, g.interp, flags, true);
		if (!k)
			return fail(FIXCA_ERR_FORMAT, "no tiled kernel for this format");
		const int limit = smem_limit(dev);
		if (g.interp == 0) {
			const KernelEntry *ks = lookup_none_stream(f.sample_bytes, f.nch);
			if (ks && plan_stream(ks, f, g, y1, y2, dev, limit, pl))
				return FIXCA_OK;
		}
		if (g.interp != 0 && !wants_fast(f, flags)) {
			// EXACT on integer samples: the streaming kernel with exact repair of near-tie samples, when it fits
			const KernelEntry *kr = lookup_exact_stream(f.kind, f.nch, g.interp);
			if (kr && plan_stream(kr, f, g, y1, y2, dev, limit, pl))
				return FIXCA_OK;
		}
		if (k->stream) {
			if (plan_stream(k, f, g, y1, y2, dev, limit, pl))
				return FIXCA_OK;
			// the window is wider than a TMA box: narrower strips if the format has them
			const KernelEntry *k2 = lookup_fast_variant(f.kind, f.nch, g.interp, 4);
			if (k2 && plan_stream(k2, f, g, y1, y2, dev, limit, pl))
				return FIXCA_OK;
			// the ring does not fit (huge shifts): fall back to per-tile windows
			k = lookup_fast_variant(f.kind, f.nch, g.interp, 2);
		}
		// Prefer the tallest tile that still leaves room for `want` CTAs per SM.
		const int target_ctas = std::max(1, tuning().tile_ctas);
		const int forced_th = tuning().tile_h;
		static const int th_choices[] = {64, 48, 32, 24, 16, 12, 8, 4};
		const size_t sm_total = 227 * 1024;
		int best_th = 0;
		size_t best_smem = 0;
		int best_wb = 0, best_rows = 0, best_off[5] = {0, 0, 0, 0, 0};
		for (int pass = 0; pass < 2 && !best_th; ++pass) {
			for (int th : th_choices) {
				if (forced_th && th != forced_th)
					continue;
				if (th > 8 && th >= 2 * (y2 - y1) && !forced_th)
					continue;	// tile much taller than the band
				int wb, rows;
				window_extent(g, f.bpp, y1, y2, k->tw, th, k->strip_p, wb, rows);
				const size_t off_ytab = align_up(sizeof(TileHeader), 16);
				size_t off_nemit = 0, ne_pitch = 0, off_win, win_rows = (size_t)rows;
				if (k->strip_p) {
					// [header | float4 ytab[2][th] | u8 nemit[2][ne_pitch] | window (+3 rows the
					//  4-row-unrolled walker may touch on either side) | staging tile]
					ne_pitch = align_up((size_t)rows + 3 + 4, 4);
					off_nemit = off_ytab + (size_t)2 * th * 16;
					off_win = align_up(off_nemit + 2 * ne_pitch, 128);
					win_rows = (size_t)rows + 6;	// 3 phase-alignment rows in front, 3 unroll-slack rows behind
				} else {
					off_win = align_up(off_ytab + (size_t)2 * th * k->ycoef_bytes, 128);
				}
				const size_t off_out = align_up(off_win + (size_t)wb * win_rows, 128);
				const size_t total = off_out + (size_t)th * k->tw * f.bpp;
				const size_t budget = pass == 0 ? (sm_total / target_ctas - 1024) : (size_t)limit;
				if (total <= budget && total <= (size_t)limit) {
					best_th = th; best_smem = total; best_wb = wb; best_rows = rows;
					best_off[0] = (int)off_ytab; best_off[1] = (int)off_win; best_off[2] = (int)off_out;
					best_off[3] = (int)off_nemit; best_off[4] = (int)ne_pitch;
					break;
				}
			}
		}
		if (best_th) {
			pl.k = k;
			a.th = best_th;
			a.win_pitch = best_wb;
			a.win_rows = best_rows;
			a.off_ytab = best_off[0];
			a.off_win = best_off[1];
			a.off_out = best_off[2];
			a.off_nemit = best_off[3];
			a.ne_pitch = best_off[4];
			pl.smem = best_smem;
			pl.block = dim3(k->strip_p ? 2 * k->tw / k->strip_p : 2 * k->tw);
			pl.grid = dim3((g.width + k->tw - 1) / k->tw, (y2 - y1 + best_th - 1) / best_th);
			if (pl.grid.y > 65535)
				want_tiled = false;
			else
				return FIXCA_OK;
		} else {
			want_tiled = false;
		}
	}
	if (flags & FIXCA_FORCE_TILED)
		return fail(FIXCA_ERR_ARG, "FIXCA_FORCE_TILED: the tiled kernel cannot take this call (%s)",
			    tma_ok ? "source window exceeds shared memory" : "non-monotone map, pitch/pointer not 16-byte aligned, or rows that are not whole 16-byte units without FIXCA_PADDING_SCRATCH");

	pl.k = pick_kernel(f, g.interp, flags, false);
	if (!pl.k)
		return fail(FIXCA_ERR_FORMAT, "no kernel for this format");
	pl.block = dim3(64, 4);
	pl.grid = dim3((g.width + 63) / 64, (y2 - y1 + 3) / 4);
	pl.smem = 0;
	if (pl.grid.y > 65535)
		return fail(FIXCA_ERR_ARG, "band of %d rows is too tall for one launch", y2 - y1);
	return FIXCA_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda).
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
				    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
				    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn encode_tiled()
{
	static encode_tiled_fn fn = nullptr;
	static std::once_flag once;
	std::call_once(once, []() {
		void *p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
		    q == cudaDriverEntryPointSuccess)
			fn = (encode_tiled_fn)p;
		else
			cudaGetLastError();
	});
	return fn;
}

// A 3-D map of 8-byte elements over `frames` frames (frame_stride bytes apart) of `rows` rows of `row_bytes`
// (multiple of 16) bytes at `pitch`; boxes of box_bytes x box_rows x 1.  Out-of-range parts of a box are
// zero-filled on load, dropped on store.
static bool make_tensor_map(CUtensorMap &tm, const void *base, size_t pitch, size_t row_bytes, size_t rows,
			    size_t frames, size_t frame_stride, unsigned box_bytes, unsigned box_rows)
{
	encode_tiled_fn enc = encode_tiled();
	if (!enc || box_bytes % 16 || box_bytes / 8 > 256 || box_rows > 256 || rows == 0 || row_bytes % 16 || frames == 0 ||
	    frame_stride % 16 || pitch % 16)
		return false;
	const cuuint64_t dims[3] = {row_bytes / 8, rows, frames};
	const cuuint64_t strides[2] = {pitch, frame_stride};
	const cuuint32_t box[3] = {box_bytes / 8, box_rows, 1};
	const cuuint32_t estr[3] = {1, 1, 1};
	return enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void *>(base), dims, strides, box, estr,
		   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
		   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// CTAs of kernel k the hardware keeps resident per SM at this block size and dynamic shared memory (the grid of a
// streaming launch is sized to be resident all at once); asked once per (kernel, threads, shared memory, device)
static int resident_ctas(const KernelEntry *k, int threads, size_t smem, int dev)	// (smem: the caller's own limit; part of the key only)
{
	struct Key { const void *fn; int threads; size_t smem; int dev; int n; };
	static thread_local Key cache[16];
	static thread_local int next = 0;
	for (const Key &c : cache)
		if (c.fn == (const void *)k->fn && c.threads == threads && c.smem == smem && c.dev == dev)
			return c.n;
	int n = 64;		// unknown: no extra limit
	cudaFuncAttributes fa;
	if (cudaFuncGetAttributes(&fa, (const void *)k->fn) == cudaSuccess && fa.numRegs > 0) {
		// registers are allocated per warp in units of 256; 64 K registers per SM
		const int per_warp = (fa.numRegs * 32 + 255) / 256 * 256;
		n = std::max(1, 65536 / (per_warp * ((threads + 31) / 32)));
	} else {
		cudaGetLastError();
	}
	cache[next] = Key{(const void *)k->fn, threads, smem, dev, n};
	next = (next + 1) % 16;
	return n;
}

// Per-plan tables of the streaming kernels live in device memory and are shared by every plan with the same inputs:
// a small per-thread cache keyed by exactly what the fill kernel reads.  A table is filled on a private stream and
// waited for here -- a plan is made once per distinct call (the plan cache), a table once per distinct band / image
// width.  `fill(mem, stream)` launches the fill kernel(s).
template <class Key, class Fill>
static std::shared_ptr<void> cached_table(int tag, const Key &key, int dev, size_t bytes, Fill fill)
{
	static_assert(sizeof(Key) <= 192, "table key");
	struct Slot { int tag; unsigned char key[192]; std::shared_ptr<void> mem; bool used; };
	constexpr int SLOTS = 96;
	static thread_local Slot slots[SLOTS];
	static thread_local int next = 0;
	static thread_local cudaStream_t fill_stream[64];	// per device, created on first use (never destroyed: process lifetime)
	for (Slot &sl : slots)
		if (sl.used && sl.tag == tag && !memcmp(sl.key, &key, sizeof key))
			return sl.mem;
	if (dev < 0 || dev >= 64)
		return nullptr;
	int cur = -1;
	if (cudaGetDevice(&cur) != cudaSuccess)
		return nullptr;
	if (cur != dev && cudaSetDevice(dev) != cudaSuccess)
		return nullptr;
	bool ok = false;
	void *mem = nullptr;
	do {
		if (!fill_stream[dev] && cudaStreamCreateWithFlags(&fill_stream[dev], cudaStreamNonBlocking) != cudaSuccess)
			break;
		if (cudaMalloc(&mem, bytes) != cudaSuccess)
			break;
		if (cudaMemsetAsync(mem, 0, bytes, fill_stream[dev]) != cudaSuccess)
			break;
		if (fill(mem, fill_stream[dev]) != cudaSuccess)
			break;
		if (cudaStreamSynchronize(fill_stream[dev]) != cudaSuccess)
			break;
		ok = true;
	} while (0);
	if (!ok) {
		cudaGetLastError();
		if (mem)
			cudaFree(mem);
	}
	if (cur != dev)
		cudaSetDevice(cur);
	if (!ok)
		return nullptr;
	// (cudaFree waits for the device: no launch of a plan that held the table can still be reading it)
	std::shared_ptr<void> sp(mem, [](void *p) { if (cudaFree(p) != cudaSuccess) cudaGetLastError(); });
	Slot &sl = slots[next];
	next = (next + 1) % SLOTS;
	sl.tag = tag;
	memset(sl.key, 0, sizeof sl.key);
	memcpy(sl.key, &key, sizeof key);
	sl.mem = sp;
	sl.used = true;
	return sp;
}

// Queue memory of the deferred exact-repair form (stream kernel -> repair_patch_kernel): one allocation per calling
// thread, device and stream, grown to the largest launch seen.  Launches in one stream are ordered, so they can share it
// (every launch writes every region's count before its patch kernel reads it: nothing carries over); launches in
// different streams, or from different threads, never share one.  Growing frees the old block (cudaFree waits for the
// device: no launch can still be using it).
struct RepairQueueSlot { int dev; cudaStream_t stream; size_t bytes; void *mem; };
static thread_local RepairQueueSlot tl_repair_queues[16];
static void repair_queue_release()	// (the calling thread's; fixca_cuda_release())
{
	int cur = -1;
	cudaGetDevice(&cur);
	for (RepairQueueSlot &c : tl_repair_queues) {
		if (c.mem && cudaSetDevice(c.dev) == cudaSuccess && cudaFree(c.mem) != cudaSuccess)
			cudaGetLastError();
		c.mem = nullptr;
	}
	if (cur >= 0)
		cudaSetDevice(cur);
}
static void *repair_queue(int dev, cudaStream_t stream, size_t bytes)
{
	typedef RepairQueueSlot Slot;
	constexpr int SLOTS = 16;
	Slot (&slots)[SLOTS] = tl_repair_queues;
	static thread_local int next = 0;
	Slot *sl = nullptr;
	for (Slot &c : slots)
		if (c.mem && c.dev == dev && c.stream == stream)
			sl = &c;
	if (sl && sl->bytes >= bytes)
		return sl->mem;
	size_t had = sl ? sl->bytes : 0;	// (this stream's block is too small: grow geometrically)
	if (!sl) {
		sl = &slots[next];
		next = (next + 1) % SLOTS;
	}
	int cur = -1;
	if (cudaGetDevice(&cur) != cudaSuccess)
		return nullptr;
	if (sl->mem) {	// an older stream's block, or one that is too small
		if (sl->dev != cur)
			cudaSetDevice(sl->dev);
		if (cudaFree(sl->mem) != cudaSuccess)
			cudaGetLastError();
		sl->mem = nullptr;
	}
	if (cudaSetDevice(dev) != cudaSuccess)
		return nullptr;
	const size_t want = std::max(bytes + bytes / 4, 2 * had);	// (the host driver's chunks ramp up: few regrowths)
	void *mem = nullptr;
	if (cudaMalloc(&mem, want) != cudaSuccess) {
		cudaGetLastError();
		mem = nullptr;
	}
	if (cur != dev)
		cudaSetDevice(cur);
	if (!mem)
		return nullptr;
	*sl = Slot{dev, stream, want, mem};
	return mem;
}

// The tables of a streaming plan:
//   chunk table (stream_meta_kernel): vertical weights / tap rows and source-row spans of every 8-row chunk of
//     [y1, y2): depends on the y axes, the band, the kernel family and (None) the ring geometry;
//   column tables: None -- the nearest source column of every column of the strips (stream_cols_kernel); Linear /
//     Cubic -- every compute thread's column set-up, one record per strip and thread (stream_setup_kernel): depend on
//     the x axes, the width and the kernel family.
static bool stream_tables(const KernelEntry *k, const Format &f, int dev, int strips, Plan &pl)
{
	KernelArgs &a = pl.args;
	struct ChunkKey {
		int interp, mode, kind, y1, y2, ring_rows, win_pitch, dev, height;
		int center[2], size[2];
		double scale[2], shift[2];
	} ck;
	memset(&ck, 0, sizeof ck);
	ck.interp = a.g.interp; ck.mode = k->repair; ck.kind = a.g.interp ? (int)f.kind : 0;
	ck.y1 = a.y1; ck.y2 = a.y2; ck.dev = dev; ck.height = a.g.height;
	if (a.g.interp == 0) { ck.ring_rows = a.ring_rows; ck.win_pitch = a.win_pitch; }
	for (int c = 0; c < 2; ++c) {
		ck.center[c] = a.g.y[c].center; ck.size[c] = a.g.y[c].size;
		ck.scale[c] = a.g.y[c].scale; ck.shift[c] = a.g.y[c].shift;
	}
	const int nchunks = (a.y2 - a.y1 + STREAM_CH - 1) / STREAM_CH;
	const size_t span_off = align_up((size_t)nchunks * stream_meta_record_bytes(k->repair), 256);
	const KernelArgs args = a;
	const int interp = a.g.interp, mode = k->repair;
	const SampleKind kind = f.kind;
	pl.tables = cached_table(1, ck, dev, span_off + (size_t)nchunks * sizeof(StreamSpan), [&](void *mem, cudaStream_t st) {
		return launch_stream_meta(interp, mode, kind, args, mem, (unsigned char *)mem + span_off, nchunks, st);
	});
	if (!pl.tables)
		return false;
	a.meta_tab = pl.tables.get();
	a.span_tab = (const unsigned char *)pl.tables.get() + span_off;

	struct ColKey {
		const void *entry;	// the kernel family (sample type, layout, arithmetic)
		int interp, ncols, dev, width;
		int center[2], size[2];
		double scale[2], shift[2];
	} xk;
	memset(&xk, 0, sizeof xk);
	const int ncols = strips * k->tw;
	xk.entry = k; xk.interp = interp; xk.ncols = ncols; xk.dev = dev; xk.width = a.g.width;
	for (int c = 0; c < 2; ++c) {
		xk.center[c] = a.g.x[c].center; xk.size[c] = a.g.x[c].size;
		xk.scale[c] = a.g.x[c].scale; xk.shift[c] = a.g.x[c].shift;
	}
	if (interp == 0) {
		// None: the nearest source column of every column (the entry does not matter: one table per width)
		xk.entry = nullptr;
		pl.col_tables = cached_table(2, xk, dev, (size_t)2 * ncols * sizeof(int) + 256, [&](void *mem, cudaStream_t st) {
			return launch_stream_cols(args, ncols, mem, st);
		});
		if (!pl.col_tables)
			return false;
		a.col_i0 = (const int *)pl.col_tables.get();
		a.col_n = ncols;
		return true;
	}
	// Linear / Cubic: every compute thread's column set-up, one record per strip and thread
	if (!k->setup || k->setup_rec_bytes <= 0)
		return false;
	const int nthreads = 2 * k->tw / k->strip_p;
	void (*const setup)(const KernelArgs, void *) = k->setup;
	pl.col_tables = cached_table(3, xk, dev, (size_t)strips * nthreads * k->setup_rec_bytes + 256, [&](void *mem, cudaStream_t st) {
		KernelArgs aa = args;
		void *out = mem;
		void *params[] = {&aa, &out};
		return cudaLaunchKernel((const void *)setup, dim3((unsigned)strips), dim3((unsigned)nthreads), params, 0, st);
	});
	if (!pl.col_tables)
		return false;
	a.setup_tab = pl.col_tables.get();
	return true;
}

static bool plan_stream(const KernelEntry *k, const Format &f, const Geometry &g, int y1, int y2, int dev, int limit, Plan &pl)
{
	const int CH = STREAM_CH;
	int wb;
	window_extent_stream(g, f.bpp, k->tw, STREAM_COL_SLACK(k->strip_p), wb);
	wb = (int)align_up((size_t)wb, 32);	// the window box: 4 * wb must keep ring groups 128-byte aligned
	if (wb > 2048 || k->tw * f.bpp > 2048)
		return false;			// TMA boxes are at most 256 elements (of 8 bytes) wide
	const int threads = 2 * k->tw / k->strip_p + 32;	// compute warps + the TMA warp
	// As many CTAs per SM as shared memory allows up to ~32 compute warps: the narrow pixel formats are
	// instruction-bound and have small CTAs (measured: RGB8 0.086 -> 0.062 ms, RGBA16 0.155 -> 0.131 ms
	// going from 2 to 4 CTAs per SM); the 3-channel 16-bit / float strips fit 2 per SM at depth 2.
	const int compute_warps = 2 * k->tw / k->strip_p / 32;
	const int want_ctas = tuning().stream_ctas > 0 ? tuning().stream_ctas : std::max(2, 32 / compute_warps);
	const int forced_d = tuning().stream_depth;
	size_t total = 0, off_meta = 0, off_win = 0, off_out = 0, off_rq = 0, ring_rows = 0;
	int depth = 0;
	// first / last source row each chunk touches (one scan; the map is monotone, so a chunk's extremes
	// sit at its first and last output row)
	const int nchunks_all = (y2 - y1 + CH - 1) / CH;
	std::vector<int> c_lo(nchunks_all), c_hi(nchunks_all);
	for (int j = 0; j < nchunks_all; ++j) {
		const int y0 = y1 + j * CH, yl = std::min(y0 + CH, y2) - 1;
		span_needed(g.y[CH_RED], g.y[CH_BLUE], g.interp, y0, yl, c_lo[j], c_hi[j]);
	}
	// shared memory of a CTA at pipeline depth d
	auto layout = [&](int d) {
		// ring capacity: the source rows d + 1 consecutive chunks can have live at once, for any chunk start
		int max_rows = 0;
		for (int j = 0; j < nchunks_all; ++j)
			max_rows = std::max(max_rows, c_hi[std::min(j + d, nchunks_all - 1)] - c_lo[j] + 1);
		ring_rows = ((size_t)max_rows + 6) / 4 * 4;	// whole 4-row groups: a run of R rows touches at most floor((R + 6) / 4) of them
		off_meta = align_up(sizeof(StreamHeader), 16);
		off_win = align_up(off_meta + (size_t)(d + 1) * (k->repair == 2 ? sizeof(StreamMetaWide) : sizeof(StreamMeta)), 128);
		off_out = align_up(off_win + ring_rows * (size_t)wb, 128);
		total = off_out + (size_t)STREAM_NSTG * CH * k->tw * f.bpp;
		if (k->repair == 1) {	// per-warp queues of near-tie samples, u16 entries
			off_rq = align_up(total, 16);
			total = off_rq + (size_t)(2 * k->tw / k->strip_p / 32) * (32 + 32 * k->strip_p) * 2;
		}
	};
	if (forced_d > 0) {
		layout(std::min(forced_d, STREAM_MAX_D));
		if (total <= (size_t)limit)
			depth = std::min(forced_d, STREAM_MAX_D);
	} else {
		// Most CTAs per SM that still leave a pipeline depth of 2 (the HBM latency needs ~2 chunks in
		// flight per CTA).  Deeper is not better: with 2+ CTAs per SM a third chunk in flight per CTA costs
		// bandwidth for the 1536-byte strips (100 MP RGB16 None: depth 2 0.199 ms, depth 3-6 0.211 ms, depth 8
		// 0.219 ms; 8K RGBA16 None 0.093 -> 0.089 ms, Linear 0.090 -> 0.087 ms: the strips of a row drift further
		// apart and their requests lose the DRAM pages they share), a lone CTA per SM wants 4 (0.212 against
		// 0.247 / 0.258 ms at depth 3 / 2).  The narrower strips (RGBA f32: 64 pixels = 1024 bytes, 8 compute
		// warps per SM) keep the deepest pipeline that fits for None: capped at 2 the copies lose 8-10 % (r02, with the
		// pass-through tile two chunks ahead: 50 MP RGBA f32 None 0.93 at depth 2, 1.00 at depth 3+; 33 MP RGBA8 None 0.80 /
		// 0.87 at depth 4); Linear / Cubic are level or lose beyond depth 2 on every layout (RGBA8 Cubic 0.75 / 0.68 at 3).
		const bool wide_strip = k->tw * f.bpp >= 1536;
		for (int want = want_ctas; want >= 1 && !depth; --want) {
			const size_t budget = std::min<size_t>(limit, (227 * 1024) / want - 1024);
			for (int d = (!wide_strip && g.interp == 0) ? STREAM_MAX_D : want > 1 ? 2 : 4; d >= (want > 1 ? 2 : 1); --d) {
				layout(d);
				if (total <= budget) {
					depth = d;
					break;
				}
			}
		}
	}
	if (!depth)
		return false;
	int per_sm = (int)((227 * 1024) / (total + 1024));
	per_sm = std::max(1, std::min(per_sm, 2048 / threads));
	per_sm = std::min(per_sm, want_ctas);
	per_sm = std::min(per_sm, resident_ctas(k, threads, total, dev));	// registers (the FP64 pipelines: 2 per SM)
	const int strips = (g.width + k->tw - 1) / k->tw;
	const int rows = y2 - y1;
	int segs = std::max(1, sm_count(dev) * per_sm / strips);
	// One wave of CTAs, all resident at once, except None on 2-byte samples with tall segments: two waves of
	// half-height segments copy 5-7 % faster (100 MP RGB16 0.199 -> 0.187 ms = the measured copy peak, 8K RGBA16
	// 0.089 -> 0.085 ms); Linear / Cubic and the other sample sizes are level or lose (RGB8 None 81 -> 75 %).
	int waves = (g.interp == 0 && f.bpp / f.nch == 2 && rows / segs >= 512) ? 2 : 1;
	if (tuning().stream_waves > 0)
		waves = tuning().stream_waves;
	segs *= waves;
	if (pl.nframes > 1) {
		// a batch fills the GPU with frames x strips x segments CTAs: long segments (>= 512 rows) so that
		// a CTA's set-up, ring priming and first chunk -- a fifth of the lifetime of a 240-row CTA -- are spread over
		// many chunks, as long as there are CTAs for ~8 waves (128 x 4K RGB8 Cubic: 9 segments per frame 0.733, 2-4
		// segments 0.748 of the HBM peak; 64 frames: 4-6 segments 0.727, 9 segments 0.711)
		const int slots = sm_count(dev) * per_sm;
		const int fill = (slots + strips * pl.nframes - 1) / (strips * pl.nframes);
		const int waves8 = (8 * slots + strips * pl.nframes - 1) / (strips * pl.nframes);
		segs = std::max(fill, std::min((rows + 255) / 256, std::max((rows + 511) / 512, waves8)));
	}
	const int forced = tuning().stream_segs;
	if (forced > 0)
		segs = forced;
	int seg_rows = (int)align_up((size_t)(rows + segs - 1) / segs, CH);
	seg_rows = std::max(seg_rows, CH);
	segs = (rows + seg_rows - 1) / seg_rows;
	if (segs > 65535)
		return false;
	KernelArgs &a = pl.args;
	pl.k = k;
	a.th = CH;
	a.win_pitch = wb;
	a.win_rows = (int)ring_rows;
	a.ring_rows = (int)ring_rows;
	a.seg_rows = seg_rows;
	a.depth = depth;
	a.tile_lead = (depth >= 2 && tuning().stream_tlead != 1) ? 2 : 1;
	a.debug = tuning().stream_debug;	// honoured by -DFIXCA_TUNING builds only
	a.off_ytab = (int)off_meta;
	a.off_win = (int)off_win;
	a.off_out = (int)off_out;
	a.off_rq = (int)off_rq;
	pl.smem = total;
	pl.block = dim3(threads);
	pl.grid = dim3(strips, segs, pl.nframes);
	if (pl.nframes > 65535)
		return false;
	if (tuning().verbose)
		fprintf(stderr, "fixca: %s grid %d x %d, %d threads, smem %zu B (ring %zu rows x %d B, depth %d), seg %d rows, %d CTA/SM\n",
			k->name, strips, segs, threads, total, ring_rows, wb, depth, seg_rows, per_sm);
	// rows as the kernels see them: align16(width * bpp) bytes (they may touch the padding bytes)
	const size_t row_bytes = align_up((size_t)g.width * f.bpp, 16);
	const size_t src_rows = (size_t)pl.src_rows_avail, dst_rows = (size_t)(y2 - a.dst_row0);
	const size_t nf = (size_t)pl.nframes;
	const size_t sfs = nf > 1 ? pl.src_frame_stride : (size_t)a.src_pitch * src_rows;
	const size_t dfs = nf > 1 ? pl.dst_frame_stride : (size_t)a.dst_pitch * dst_rows;
	if (!make_tensor_map(pl.tm_win, a.src, (size_t)a.src_pitch, row_bytes, src_rows, nf, sfs, (unsigned)wb, 4) ||
	    !make_tensor_map(pl.tm_tile, a.src, (size_t)a.src_pitch, row_bytes, src_rows, nf, sfs, (unsigned)(k->tw * f.bpp), CH) ||
	    !make_tensor_map(pl.tm_out, a.dst, (size_t)a.dst_pitch, row_bytes, dst_rows, nf, dfs, (unsigned)(k->tw * f.bpp), CH))
		return false;
	for (int i = 0; i < pl.fan.n; ++i) {
		if (!make_tensor_map(pl.fan.tm[i], pl.fan_dst[i], (size_t)a.dst_pitch, row_bytes, dst_rows, nf, dfs, (unsigned)(k->tw * f.bpp), CH))
			return false;
	}
	if (!stream_tables(k, f, dev, strips, pl))
		return false;
	if (k->patch) {
		// Queue of the deferred exact-repair form: 6.2e-4 of the samples are within the FP32 bound of a rounding boundary
		// (DESIGN.md 4.6); every compute warp of the grid has its own region with room for 2.5 times what it expects plus
		// 48 (the patch kernel recomputes a region that overflows), and a count
		const int warps = 2 * k->tw / k->strip_p / 32;
		const unsigned long long nregions = (unsigned long long)strips * segs * pl.nframes * warps;
		const unsigned cap = (unsigned)align_up((size_t)((double)seg_rows * 32 * k->strip_p * 6.2e-4 * 2.5) + 48, 16);
		const size_t counts = align_up((size_t)nregions * sizeof(unsigned), 256);
		if (nregions > 0xffffffffull)
			return false;
		pl.rq_counts_bytes = counts;
		pl.rq_bytes = counts + (size_t)nregions * cap * sizeof(unsigned long long);
		a.rq_cap = cap;
	}
	return true;
}

// Planning costs tens of microseconds (window scans in FP64, three tensor-map encodes), a third of
// the kernel itself; callers that repeat a call (frame streams, benchmarks, the chunks of a band) hit
// a per-thread cache instead, sized for the chunk plans of a few host calls in flight (one 100 MP call
// makes ~19 distinct chunk plans).  The key holds every input of make_plan_uncached, including the
// generation of the tuning switches.
struct PlanKey {
	int kind, nch, bpp, interp, width, height, monotone;
	int axis_center[4], axis_size[4];
	double axis_scale[4], axis_shift[4];
	const void *src, *dst;
	size_t src_pitch, dst_pitch;
	int src_row0, src_rows, dst_row0, y1, y2, dev;
	unsigned flags;
	unsigned env;
	int nframes;
	size_t src_frame_stride, dst_frame_stride;
	int nfan;
	const void *fan[STREAM_MAX_FAN];
};

static int make_plan(const Format &f, const Geometry &g, const void *d_src, size_t src_pitch, int src_row0, int src_rows,
		     void *d_dst, size_t dst_pitch, int dst_row0, int y1, int y2, unsigned flags, int dev, Plan &pl,
		     const Batch &batch = Batch())
{
	PlanKey k;
	memset(&k, 0, sizeof k);
	k.kind = f.kind; k.nch = f.nch; k.bpp = f.bpp;
	k.interp = g.interp; k.width = g.width; k.height = g.height; k.monotone = g.monotone;
	const Axis *ax[4] = {&g.x[0], &g.x[1], &g.y[0], &g.y[1]};
	for (int i = 0; i < 4; ++i) {
		k.axis_center[i] = ax[i]->center; k.axis_size[i] = ax[i]->size;
		k.axis_scale[i] = ax[i]->scale; k.axis_shift[i] = ax[i]->shift;
	}
	k.src = d_src; k.dst = d_dst; k.src_pitch = src_pitch; k.dst_pitch = dst_pitch;
	k.src_row0 = src_row0; k.src_rows = src_rows; k.dst_row0 = dst_row0; k.y1 = y1; k.y2 = y2; k.dev = dev;
	k.flags = flags;
	k.env = (unsigned)tuning().generation;
	k.nframes = batch.nframes; k.src_frame_stride = batch.src_stride; k.dst_frame_stride = batch.dst_stride;
	k.nfan = batch.nfan;
	for (int i = 0; i < batch.nfan; ++i)
		k.fan[i] = batch.fan[i];
	constexpr int SLOTS = 96;
	static thread_local PlanKey keys[SLOTS];
	static thread_local Plan plans[SLOTS];
	static thread_local unsigned long long hashes[SLOTS];	// 0 = empty slot
	static thread_local int next = 0;
	unsigned long long h = 1469598103934665603ull;
	for (size_t i = 0; i < sizeof k; ++i)
		h = (h ^ reinterpret_cast<const unsigned char *>(&k)[i]) * 1099511628211ull;
	h |= 1ull;
	for (int i = 0; i < SLOTS; ++i)
		if (hashes[i] == h && !memcmp(&keys[i], &k, sizeof k)) {
			pl = plans[i];
			return FIXCA_OK;
		}
	const int rc = make_plan_uncached(f, g, d_src, src_pitch, src_row0, src_rows, d_dst, dst_pitch, dst_row0, y1, y2, flags, dev, pl, batch);
	if (rc == FIXCA_OK) {
		keys[next] = k;
		plans[next] = pl;
		hashes[next] = h;
		next = (next + 1) % SLOTS;
	}
	return rc;
}

static int launch_plan(const Plan &pl, cudaStream_t stream)
{
	if (pl.smem > 48 * 1024) {
		// Opt in to large dynamic shared memory; the attribute sticks per function and device, so
		// only raise it when a plan needs more than what was already granted.
		static std::mutex mu;
		static std::vector<std::pair<std::pair<const void *, int>, size_t>> granted;
		int dev = 0;
		cudaGetDevice(&dev);
		std::lock_guard<std::mutex> lock(mu);
		size_t *have = nullptr;
		for (auto &e : granted)
			if (e.first.first == (const void *)pl.k->fn && e.first.second == dev)
				have = &e.second;
		if (!have || *have < pl.smem) {
			CUDA_TRY(cudaFuncSetAttribute((const void *)pl.k->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
			if (have)
				*have = pl.smem;
			else
				granted.push_back({{(const void *)pl.k->fn, dev}, pl.smem});
		}
	}
	KernelArgs a = pl.args;
	if (pl.k->patch) {
		int dev = 0;
		CUDA_TRY(cudaGetDevice(&dev));
		unsigned char *q = (unsigned char *)repair_queue(dev, stream, pl.rq_bytes);
		if (!q)
			return fail(FIXCA_ERR_NOMEM, "no device memory for the exact-repair queue (%zu bytes)", pl.rq_bytes);
		a.rq_ctl = (unsigned *)q;
		a.rq_entries = (unsigned long long *)(q + pl.rq_counts_bytes);
	}
	CUtensorMap tm[3] = {pl.tm_win, pl.tm_tile, pl.tm_out};
	StreamFanout fan = pl.fan;
	void *params[] = {&a, &tm[0], &tm[1], &tm[2], &fan};	// the tensor maps are only declared by stream kernels
	if (pl.k->stream && !tuning().no_pdl) {
		// programmatic dependent launch (see griddep_wait() in fixca_stream.cuh): back-to-back launches in one
		// stream overlap the next grid's set-up with this grid's tail; memory ordering is unchanged
		cudaLaunchConfig_t cfg = {};
		cfg.gridDim = pl.grid;
		cfg.blockDim = pl.block;
		cfg.dynamicSmemBytes = pl.smem;
		cfg.stream = stream;
		cudaLaunchAttribute attr[1];
		attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		attr[0].val.programmaticStreamSerializationAllowed = 1;
		cfg.attrs = attr;
		cfg.numAttrs = 1;
		CUDA_TRY(cudaLaunchKernelExC(&cfg, (const void *)pl.k->fn, params));
	} else {
		CUDA_TRY(cudaLaunchKernel((const void *)pl.k->fn, pl.grid, pl.block, params, pl.smem, stream));
	}
	g_launches.fetch_add(1);
	snprintf(tl_kernel, sizeof tl_kernel, "%s", pl.k->name);
	if (pl.k->patch) {
		// deferred exact repair: the queued near-tie samples are recomputed behind the streaming grid (the patch kernel
		// waits for its completion: griddepcontrol.wait, or plain stream order)
		PatchArgs pa;
		memset(&pa, 0, sizeof pa);
		pa.src_frame_stride = pl.nframes > 1 ? pl.src_frame_stride : 0;
		pa.dst_frame_stride = pl.nframes > 1 ? pl.dst_frame_stride : 0;
		pa.grid_x = (int)pl.grid.x;
		pa.grid_y = (int)pl.grid.y;
		pa.warps = (int)pl.block.x / 32 - 1;
		pa.nregions = pl.grid.x * pl.grid.y * pl.grid.z * (unsigned)pa.warps;
		pa.nframes = (unsigned)pl.nframes;
		pa.tw = pl.k->tw;
		pa.p = pl.k->strip_p;
		pa.nfan = pl.fan.n;
		for (int i = 0; i < pl.fan.n; ++i)
			pa.fan[i] = (unsigned char *)pl.fan_dst[i];
		void *pparams[] = {&a, &pa};
		cudaLaunchConfig_t cfg = {};
		// threads per region: what a region expects (6.2e-4 of its samples) plus four standard deviations, as a power of two
		const double expect = (double)pl.args.seg_rows * 32 * pl.k->strip_p * 6.2e-4;
		pa.lanes = 8;
		while (pa.lanes < 256 && pa.lanes < expect + 4 * sqrt(expect) + 4)
			pa.lanes *= 2;
		cfg.gridDim = dim3((unsigned)std::min<unsigned long long>(((unsigned long long)pa.nregions * pa.lanes + 255) / 256, 1u << 30));
		cfg.blockDim = dim3(256);
		cfg.stream = stream;
		cudaLaunchAttribute attr[1];
		attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		attr[0].val.programmaticStreamSerializationAllowed = 1;
		cfg.attrs = attr;
		cfg.numAttrs = tuning().no_pdl ? 0 : 1;
		CUDA_TRY(cudaLaunchKernelExC(&cfg, (const void *)pl.k->patch, pparams));
		g_launches.fetch_add(1);
	}
	return FIXCA_OK;
}

// What each entry point accepts beyond the precision / kernel-selection bits.
enum { ALLOW_PREVIEW = 1, ALLOW_COLUMNS = 2 };
static int check_flags(unsigned flags, int allow, const Format &f, const fixca_params *p, const char *entry)
{
	const unsigned known = FIXCA_PRECISION_MASK | FIXCA_FORCE_DIRECT | FIXCA_FORCE_TILED | FIXCA_PREVIEW_OVERLAY |
			       FIXCA_COLUMN_SELECTION | FIXCA_PADDING_SCRATCH;
	if (flags & ~known)
		return fail(FIXCA_ERR_ARG, "%s: unknown flag bits %#x", entry, flags & ~known);
	if ((flags & FIXCA_PRECISION_MASK) > FIXCA_PRECISION_FAST)
		return fail(FIXCA_ERR_ARG, "%s: unknown precision %#x", entry, flags & FIXCA_PRECISION_MASK);
	if ((flags & FIXCA_PREVIEW_OVERLAY) && !(allow & ALLOW_PREVIEW))
		return fail(FIXCA_ERR_UNSUPPORTED, "%s: the preview overlay is a single-image call (fixca_cuda_region*, fixca_cuda_region_dev)", entry);
	if ((flags & FIXCA_COLUMN_SELECTION) && !(allow & ALLOW_COLUMNS))
		return fail(FIXCA_ERR_UNSUPPORTED, "%s: FIXCA_COLUMN_SELECTION is an option of fixca_cuda_region_ex only", entry);
	(void)f; (void)p;	// (u64 samples were refused here until the x87 steps of get_pixel / set_pixel were restated: fixca_kernels.cuh)
	return FIXCA_OK;
}

static int check_common(const void *src, const void *dst, int width, int height, const fixca_params *p, int y1, int y2)
{
	if (!src || !dst || !p)
		return fail(FIXCA_ERR_ARG, "NULL pointer argument");
	if (width <= 0 || height <= 0)
		return fail(FIXCA_ERR_ARG, "non-positive image size %dx%d", width, height);
	if (y1 < 0 || y2 > height || y1 > y2)
		return fail(FIXCA_ERR_ARG, "rows [%d,%d) outside 0..%d", y1, y2, height);
	return FIXCA_OK;
}

static int current_device_or(int device, int &out)
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0) {
		cudaGetLastError();
		return fail(FIXCA_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU path", e != cudaSuccess ? cudaGetErrorString(e) : "count 0");
	}
	if (device < 0) {
		CUDA_TRY(cudaGetDevice(&out));
	} else {
		if (device >= n)
			return fail(FIXCA_ERR_NO_DEVICE, "device %d requested, %d present", device, n);
		out = device;
	}
	return FIXCA_OK;
}

// ---------------------------------------------------------------------------
// device-resident entry
// ---------------------------------------------------------------------------
extern "C" int fixca_cuda_region_dev(const void *d_src, size_t src_pitch, int src_row0, int src_rows,
				     void *d_dst, size_t dst_pitch, int dst_row0,
				     int width, int height, int bytes, int bpc,
				     const fixca_params *params, int y1, int y2, unsigned flags, void *stream)
{
	int rc = check_common(d_src, d_dst, width, height, params, y1, y2);
	if (rc) return rc;
	Format f;
	if ((rc = parse_format(bytes, bpc, f))) return rc;
	if ((rc = check_flags(flags, ALLOW_PREVIEW, f, params, "fixca_cuda_region_dev"))) return rc;
	Geometry g;
	if ((rc = make_geometry(width, height, params, g))) return rc;
	if (src_pitch < (size_t)width * bytes || dst_pitch < (size_t)width * bytes)
		return fail(FIXCA_ERR_ARG, "pitch smaller than a row (%zu / %zu < %zu)", src_pitch, dst_pitch, (size_t)width * bytes);
	if (y1 == y2)
		return FIXCA_OK;
	int lo, hi;
	source_rows(g, y1, y2, lo, hi);
	if (src_row0 < 0 || lo < src_row0 || hi >= src_row0 + src_rows || src_row0 + src_rows > height)
		return fail(FIXCA_ERR_ARG, "source rows [%d,%d) do not cover the rows [%d,%d] that output rows [%d,%d) read",
			    src_row0, src_row0 + src_rows, lo, hi, y1, y2);
	if (dst_row0 < 0 || dst_row0 > y1)
		return fail(FIXCA_ERR_ARG, "dst_row0 %d is past the first output row %d", dst_row0, y1);
	int dev;
	if ((rc = current_device_or(-1, dev))) return rc;
	Plan pl;
	if ((rc = make_plan(f, g, d_src, src_pitch, src_row0, src_rows, d_dst, dst_pitch, dst_row0, y1, y2, flags, dev, pl))) return rc;
	if ((rc = launch_plan(pl, (cudaStream_t)stream))) return rc;
	if (flags & FIXCA_PREVIEW_OVERLAY) {
		// fix-ca.c:1322-1327: saturate (iff saturation != 0) and centerline on every finished row
		CUDA_TRY(launch_preview(f.kind, f.nch, (unsigned char *)d_dst, (long long)dst_pitch, dst_row0, y1, y2, width,
					(int)params->lens_x, (int)params->lens_y, params->saturation, (cudaStream_t)stream));
		g_launches.fetch_add(1);
	}
	return FIXCA_OK;
}

// The band stored into several destination frames at once (the all-gather form of the reassembly): the streaming
// kernels fan every finished chunk out from shared memory to all frames in one launch; any other kernel is
// launched once per destination.
extern "C" int fixca_cuda_region_dev_fanout(const void *d_src, size_t src_pitch, int src_row0, int src_rows,
					    void *const *d_dsts, int ndst, size_t dst_pitch, int dst_row0,
					    int width, int height, int bytes, int bpc,
					    const fixca_params *params, int y1, int y2, unsigned flags, void *stream)
{
	if (!d_dsts || ndst < 1 || ndst > STREAM_MAX_FAN + 1)
		return fail(FIXCA_ERR_ARG, "fan-out over %d destinations (1..%d)", ndst, STREAM_MAX_FAN + 1);
	for (int i = 0; i < ndst; ++i)
		if (!d_dsts[i])
			return fail(FIXCA_ERR_ARG, "fan-out destination %d is NULL", i);
	int rc = check_common(d_src, d_dsts[0], width, height, params, y1, y2);
	if (rc) return rc;
	Format f;
	if ((rc = parse_format(bytes, bpc, f))) return rc;
	if ((rc = check_flags(flags, 0, f, params, "fixca_cuda_region_dev_fanout"))) return rc;
	Geometry g;
	if ((rc = make_geometry(width, height, params, g))) return rc;
	if (src_pitch < (size_t)width * bytes || dst_pitch < (size_t)width * bytes)
		return fail(FIXCA_ERR_ARG, "pitch smaller than a row (%zu / %zu < %zu)", src_pitch, dst_pitch, (size_t)width * bytes);
	if (y1 == y2)
		return FIXCA_OK;
	int lo, hi;
	source_rows(g, y1, y2, lo, hi);
	if (src_row0 < 0 || lo < src_row0 || hi >= src_row0 + src_rows || src_row0 + src_rows > height)
		return fail(FIXCA_ERR_ARG, "source rows [%d,%d) do not cover the rows [%d,%d] that output rows [%d,%d) read",
			    src_row0, src_row0 + src_rows, lo, hi, y1, y2);
	if (dst_row0 < 0 || dst_row0 > y1)
		return fail(FIXCA_ERR_ARG, "dst_row0 %d is past the first output row %d", dst_row0, y1);
	int dev;
	if ((rc = current_device_or(-1, dev))) return rc;
	Batch b;
	b.nfan = ndst - 1;
	for (int i = 1; i < ndst; ++i)
		b.fan[i - 1] = d_dsts[i];
	Plan pl;
	if ((rc = make_plan(f, g, d_src, src_pitch, src_row0, src_rows, d_dsts[0], dst_pitch, dst_row0, y1, y2, flags, dev, pl, b))) return rc;
	if ((rc = launch_plan(pl, (cudaStream_t)stream))) return rc;
	if (pl.fan.n != ndst - 1) {	// not a streaming kernel: one launch per remaining destination
		for (int i = 1; i < ndst; ++i) {
			if ((rc = make_plan(f, g, d_src, src_pitch, src_row0, src_rows, d_dsts[i], dst_pitch, dst_row0, y1, y2, flags, dev, pl))) return rc;
			if ((rc = launch_plan(pl, (cudaStream_t)stream))) return rc;
		}
	}
	return FIXCA_OK;
}

// ---------------------------------------------------------------------------
// peer frames: a destination frame on one GPU that the kernels of every rank store into
// ---------------------------------------------------------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == FIXCA_IPC_HANDLE_BYTES, "CUDA IPC handle size");

extern "C" int fixca_cuda_frame_alloc(size_t bytes, void **d_frame, unsigned char *handle)
{
	if (!d_frame || !handle || !bytes)
		return fail(FIXCA_ERR_ARG, "fixca_cuda_frame_alloc: NULL / empty argument");
	int dev, rc;
	if ((rc = current_device_or(-1, dev))) return rc;
	void *p = nullptr;
	CUDA_TRY(cudaMalloc(&p, bytes));
	cudaIpcMemHandle_t h;
	cudaError_t e = cudaIpcGetMemHandle(&h, p);
	if (e != cudaSuccess) {
		cudaFree(p);
		cudaGetLastError();
		return fail(FIXCA_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
	}
	memcpy(handle, &h, sizeof h);
	*d_frame = p;
	return FIXCA_OK;
}

extern "C" int fixca_cuda_frame_open(const unsigned char *handle, void **d_frame)
{
	if (!d_frame || !handle)
		return fail(FIXCA_ERR_ARG, "fixca_cuda_frame_open: NULL argument");
	int dev, rc;
	if ((rc = current_device_or(-1, dev))) return rc;
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, sizeof h);
	void *p = nullptr;
	// maps the owner's allocation into this process and enables peer access to its device
	CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
	*d_frame = p;
	return FIXCA_OK;
}

extern "C" int fixca_cuda_frame_close(void *d_frame)
{
	if (!d_frame) return FIXCA_OK;
	CUDA_TRY(cudaIpcCloseMemHandle(d_frame));
	return FIXCA_OK;
}

extern "C" int fixca_cuda_frame_free(void *d_frame)
{
	if (!d_frame) return FIXCA_OK;
	CUDA_TRY(cudaFree(d_frame));
	return FIXCA_OK;
}

// ---------------------------------------------------------------------------
// device-resident batch of frames
// ---------------------------------------------------------------------------
extern "C" int fixca_cuda_frames_dev(const void *d_src, size_t src_pitch, size_t src_frame_stride,
				     void *d_dst, size_t dst_pitch, size_t dst_frame_stride, int nframes,
				     int width, int height, int bytes, int bpc,
				     const fixca_params *params, unsigned flags, void *stream)
{
	if (nframes < 0)
		return fail(FIXCA_ERR_ARG, "negative frame count");
	if (nframes == 0)
		return FIXCA_OK;
	int rc = check_common(d_src, d_dst, width, height, params, 0, height);
	if (rc) return rc;
	Format f;
	if ((rc = parse_format(bytes, bpc, f))) return rc;
	if ((rc = check_flags(flags, 0, f, params, "fixca_cuda_frames_dev"))) return rc;
	Geometry g;
	if ((rc = make_geometry(width, height, params, g))) return rc;
	const size_t row = (size_t)width * bytes;
	if (src_pitch < row || dst_pitch < row)
		return fail(FIXCA_ERR_ARG, "pitch smaller than a row (%zu / %zu < %zu)", src_pitch, dst_pitch, row);
	if (nframes > 1 && (src_frame_stride < src_pitch * (size_t)(height - 1) + row || dst_frame_stride < dst_pitch * (size_t)(height - 1) + row))
		return fail(FIXCA_ERR_ARG, "frame stride smaller than a frame");
	int dev;
	if ((rc = current_device_or(-1, dev))) return rc;
	const unsigned char *sp = (const unsigned char *)d_src;
	unsigned char *dp = (unsigned char *)d_dst;
	for (int first = 0; first < nframes;) {
		Batch b;
		b.nframes = std::min(nframes - first, 65535);
		b.src_stride = src_frame_stride;
		b.dst_stride = dst_frame_stride;
		Plan pl;
		if ((rc = make_plan(f, g, sp + (size_t)first * src_frame_stride, src_pitch, 0, height,
				    dp + (size_t)first * dst_frame_stride, dst_pitch, 0, 0, height, flags, dev, pl, b))) return rc;
		if ((rc = launch_plan(pl, (cudaStream_t)stream))) return rc;
		first += pl.nframes;	// 1 when the plan's kernel has no batch form
	}
	return FIXCA_OK;
}

// ---------------------------------------------------------------------------
// host region driver
// ---------------------------------------------------------------------------
namespace {

struct DeviceCtx {
	std::mutex mu;
	int dev = -1;
	bool ready = false;
	cudaStream_t s_up = nullptr, s_run = nullptr, s_down = nullptr;
	unsigned char *d_src = nullptr, *d_dst = nullptr, *d_aux = nullptr;	// d_aux: the preview's 8-bit window
	size_t d_src_cap = 0, d_dst_cap = 0, d_aux_cap = 0;
	unsigned char *h_in = nullptr, *h_out = nullptr;	// pinned staging rings
	size_t h_in_cap = 0, h_out_cap = 0;
	std::vector<cudaEvent_t> events;

	int init(int device)
	{
		if (ready) return FIXCA_OK;
		dev = device;
		CUDA_TRY(cudaSetDevice(dev));
		CUDA_TRY(cudaStreamCreateWithFlags(&s_up, cudaStreamNonBlocking));
		CUDA_TRY(cudaStreamCreateWithFlags(&s_run, cudaStreamNonBlocking));
		CUDA_TRY(cudaStreamCreateWithFlags(&s_down, cudaStreamNonBlocking));
		ready = true;
		return FIXCA_OK;
	}
	int reserve_dev(unsigned char *&p, size_t &cap, size_t need)
	{
		if (need <= cap) return FIXCA_OK;
		if (p) { cudaFree(p); p = nullptr; cap = 0; }
		CUDA_TRY(cudaMalloc(&p, need));
		cap = need;
		return FIXCA_OK;
	}
	int reserve_pinned(unsigned char *&p, size_t &cap, size_t need)
	{
		if (need <= cap) return FIXCA_OK;
		if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
		CUDA_TRY(cudaHostAlloc(&p, need, cudaHostAllocDefault));
		cap = need;
		return FIXCA_OK;
	}
	int event(size_t i, cudaEvent_t &e)
	{
		while (events.size() <= i) {
			cudaEvent_t ne;
			CUDA_TRY(cudaEventCreateWithFlags(&ne, cudaEventDisableTiming));
			events.push_back(ne);
		}
		e = events[i];
		return FIXCA_OK;
	}
	void release()
	{
		if (!ready) return;
		cudaSetDevice(dev);
		for (cudaEvent_t e : events) cudaEventDestroy(e);
		events.clear();
		if (d_src) cudaFree(d_src);
		if (d_dst) cudaFree(d_dst);
		if (d_aux) cudaFree(d_aux);
		if (h_in) cudaFreeHost(h_in);
		if (h_out) cudaFreeHost(h_out);
		d_src = d_dst = d_aux = h_in = h_out = nullptr;
		d_src_cap = d_dst_cap = d_aux_cap = h_in_cap = h_out_cap = 0;
		cudaStreamDestroy(s_up); cudaStreamDestroy(s_run); cudaStreamDestroy(s_down);
		ready = false;
	}
};

DeviceCtx g_ctx[16];

// fixca_cuda_frames: per-device ring of frame slots, kept between calls
struct FrameRing {
	static constexpr int SLOTS = 3;
	struct Slot {
		cudaStream_t s = nullptr;
		unsigned char *d_src = nullptr, *d_dst = nullptr, *h_in = nullptr, *h_out = nullptr;
		size_t dev_cap = 0, h_in_cap = 0, h_out_cap = 0;
		int frame = -1;
		bool staged_out = false;
	};
	std::mutex mu;
	int dev = -1;
	Slot slot[SLOTS];
	void release()
	{
		for (Slot &s : slot) {
			if (s.s) { cudaStreamSynchronize(s.s); cudaStreamDestroy(s.s); }
			if (s.d_src) cudaFree(s.d_src);
			if (s.d_dst) cudaFree(s.d_dst);
			if (s.h_in) cudaFreeHost(s.h_in);
			if (s.h_out) cudaFreeHost(s.h_out);
			s = Slot();
		}
	}
};
FrameRing g_frame_ring[16];

bool is_pinned(const void *p)
{
	cudaPointerAttributes at;
	if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
		cudaGetLastError();
		return false;
	}
	return at.type == cudaMemoryTypeHost;
}

} // namespace

// Pageable caller memory (the plug-in's g_new buffers, fix-ca.c:366-367) goes through pinned rings; one thread
// copies at ~12 GB/s, a quarter of what the PCIe link moves in both directions, so the staging copies are split
// over a few worker threads that live for the duration of one call (FIXCA_COPY_THREADS, default min(8, cores/2)).
class CopyPool {
public:
	explicit CopyPool(int nthreads)
	{
		for (int i = 1; i < nthreads; ++i)
			workers_.emplace_back([this]() { run(); });
	}
	~CopyPool()
	{
		{
			std::lock_guard<std::mutex> lock(mu_);
			quit_ = true;
		}
		cv_.notify_all();
		for (std::thread &t : workers_)
			t.join();
	}
	// blocking; slices of >= 1 MB, 64-byte aligned cuts
	void copy(void *dst, const void *src, size_t n)
	{
		const size_t parts = std::min<size_t>(workers_.size() + 1, std::max<size_t>(1, n >> 20));
		if (parts <= 1) {
			memcpy(dst, src, n);
			return;
		}
		const size_t slice = ((n + parts - 1) / parts + 63) & ~(size_t)63;
		{
			std::lock_guard<std::mutex> lock(mu_);
			for (size_t off = slice; off < n; off += slice) {
				tasks_.push_back({(unsigned char *)dst + off, (const unsigned char *)src + off, std::min(slice, n - off)});
				++pending_;
			}
		}
		cv_.notify_all();
		memcpy(dst, src, std::min(slice, n));
		std::unique_lock<std::mutex> lock(mu_);
		done_.wait(lock, [this]() { return pending_ == 0; });
	}

private:
	struct Task { unsigned char *d; const unsigned char *s; size_t n; };
	void run()
	{
		std::unique_lock<std::mutex> lock(mu_);
		for (;;) {
			cv_.wait(lock, [this]() { return quit_ || !tasks_.empty(); });
			if (tasks_.empty())
				return;
			const Task t = tasks_.back();
			tasks_.pop_back();
			lock.unlock();
			memcpy(t.d, t.s, t.n);
			lock.lock();
			if (--pending_ == 0)
				done_.notify_all();
		}
	}
	std::vector<std::thread> workers_;
	std::vector<Task> tasks_;
	std::mutex mu_;
	std::condition_variable cv_, done_;
	size_t pending_ = 0;
	bool quit_ = false;
};

static int copy_threads()
{
	const int hw = (int)std::thread::hardware_concurrency();
	return tuning().copy_threads > 0 ? tuning().copy_threads : std::min(8, std::max(1, hw / 2));
}

// One band [y1,y2) of a host image on one device: upload the source rows the
// band reads, run, download.  Rows move in chunks so that H2D of chunk i+1,
// the kernel of chunk i and D2H of chunk i-1 overlap (PCIe is full duplex).
// preview_update()'s 8-bit window (fix-ca.c:659-671): when given, `dst` of the band call is the tight 8-bit buffer of
// window columns [x, x + pw) (row y1 first) instead of the full-precision image
struct Window8 { int x, pw; };

static int region_host_band_locked(DeviceCtx &cx, int dev, const unsigned char *src, unsigned char *dst, int width, int height,
				   const Format &f, const fixca_params *params, const Geometry &g,
				   int x1, int x2, int y1, int y2, unsigned flags, bool progress, const Window8 *w8);

static int region_host_band(int dev, const unsigned char *src, unsigned char *dst, int width, int height,
			    const Format &f, const fixca_params *params, const Geometry &g,
			    int x1, int x2, int y1, int y2, unsigned flags, bool progress, const Window8 *w8 = nullptr)
{
	if (dev < 0 || dev >= 16)
		return fail(FIXCA_ERR_NO_DEVICE, "device ordinal %d out of range", dev);
	DeviceCtx &cx = g_ctx[dev];
	std::lock_guard<std::mutex> lock(cx.mu);
	int rc;
	if ((rc = cx.init(dev))) return rc;
	CUDA_TRY(cudaSetDevice(dev));
	rc = region_host_band_locked(cx, dev, src, dst, width, height, f, params, g, x1, x2, y1, y2, flags, progress, w8);
	if (rc) {
		// leave no copy in flight behind an error: H2D may still read the pinned ring, D2H may still write the
		// caller's dst, and both outlive this call's lock (the error text of the failure is kept)
		char keep[sizeof tl_error];
		memcpy(keep, tl_error, sizeof keep);
		cudaStreamSynchronize(cx.s_up);
		cudaStreamSynchronize(cx.s_run);
		cudaStreamSynchronize(cx.s_down);
		cudaGetLastError();
		memcpy(tl_error, keep, sizeof keep);
	}
	return rc;
}

static int region_host_band_locked(DeviceCtx &cx, int dev, const unsigned char *src, unsigned char *dst, int width, int height,
				   const Format &f, const fixca_params *params, const Geometry &g,
				   int x1, int x2, int y1, int y2, unsigned flags, bool progress, const Window8 *w8)
{
	int rc;

	const size_t row_bytes = (size_t)width * f.bpp;
	const size_t pitch = align_up(row_bytes, 128);
	// FIXCA_COLUMN_SELECTION: the rows are computed full width on the device (a pixel's arithmetic does not
	// depend on x1 / x2, fix-ca.c:1105-1320) and only columns [x1,x2) come back
	const size_t sel_off = (size_t)x1 * f.bpp, sel_bytes = (size_t)(x2 - x1) * f.bpp;
	// what comes back, and where it goes in the caller's buffer: the full-precision rows at their place in the image,
	// or (w8) the tight 8-bit window whose first row is image row y1
	const size_t o_row = w8 ? (size_t)w8->pw * f.nch : row_bytes;	// host row stride of dst
	const size_t o_off = w8 ? 0 : sel_off;
	const size_t o_bytes = w8 ? o_row : sel_bytes;			// bytes per row that come back
	const int o_y0 = w8 ? y1 : 0;					// dst row 0 is image row o_y0
	int band_lo, band_hi;
	source_rows(g, y1, y2, band_lo, band_hi);
	const int src_rows = band_hi - band_lo + 1;
	if ((rc = cx.reserve_dev(cx.d_src, cx.d_src_cap, pitch * src_rows))) return rc;
	if ((rc = cx.reserve_dev(cx.d_dst, cx.d_dst_cap, pitch * (size_t)(y2 - y1)))) return rc;
	if (w8 && (rc = cx.reserve_dev(cx.d_aux, cx.d_aux_cap, o_row * (size_t)(y2 - y1)))) return rc;

	// Chunking: ~32 MB (pinned caller) / ~16 MB (pageable caller) of rows per chunk (FIXCA_CHUNK_MB), at least 64 rows, a multiple of 8 rows (the
	// streaming kernel's chunk height, so every launch of the band shares one chunk grid), at most 256 chunks.
	// look at the first rows actually touched: callers may pass a whole-image base pointer of
	// which only this band's rows are backed by memory
	const bool src_pinned = is_pinned(src + (size_t)band_lo * row_bytes), dst_pinned = is_pinned(dst + (size_t)(y1 - o_y0) * o_row + o_off);
	// (pageable callers: smaller chunks, the staging copies of a chunk are not overlapped with its own transfers;
	// measured on 100 MP RGB16: 16 MB 27.3 ms, 32 MB 28.7 ms, 64 MB 41.9 ms)
	const size_t chunk_bytes = (size_t)(tuning().chunk_mb > 0 ? tuning().chunk_mb : (src_pinned && dst_pinned) ? 32 : 16) << 20;
	int chunk_rows = (int)std::max<size_t>(64, chunk_bytes / std::max<size_t>(row_bytes, 1));
	chunk_rows = std::max(chunk_rows, (y2 - y1 + 255) / 256);
	chunk_rows = (chunk_rows + 7) & ~7;
	chunk_rows = std::min(chunk_rows, y2 - y1);
	// Chunk boundaries.  Pinned callers: the first chunk's upload and the last chunk's download are the part of the
	// call nothing overlaps (one 32 MB chunk each way = 0.7 of 13.4 ms on 100 MP RGB16), so the chunks ramp up from
	// a sixteenth of a chunk at the start and down again at the end (2, 4, 8, 16, 32 ... 32, 16, 8, 4, 2 MB), all multiples of 8 rows.
	std::vector<int> bounds;
	{
		const bool ramp = src_pinned && dst_pinned && !tuning().no_chunk_ramp && !w8;	// (FIXCA_CHUNK_RAMP=0: uniform chunks, for A/B runs)
		const int small = (int)std::min<size_t>((size_t)chunk_rows, std::max<size_t>(64, ((chunk_bytes / 16) / std::max<size_t>(row_bytes, 1) + 7) & ~(size_t)7));
		std::vector<int> head, tail;
		int lo = y1, hi = y1 + (y2 - y1) / 8 * 8;	// (boundaries stay on the band's 8-row grid; the last chunk takes the odd rows)
		if (ramp && small < chunk_rows) {
			for (int r = small; r < chunk_rows && hi - lo > 4 * chunk_rows; r *= 2) {
				head.push_back(lo);
				lo += r;
				hi -= r;
				tail.push_back(hi);
			}
		}
		bounds = head;
		const int mid_end = tail.empty() ? y2 : hi;
		for (int y = lo; y < mid_end; y += chunk_rows)
			bounds.push_back(y);
		for (size_t k = tail.size(); k-- > 0;)
			bounds.push_back(tail[k]);
		bounds.push_back(y2);
	}
	const int nchunks = (int)bounds.size() - 1;

	// Pageable callers go through pinned rings (2 slots per direction).
	const int ring = 2;
	size_t in_slot = 0, out_slot = 0;
	if (!src_pinned) {
		// a chunk uploads at most its own rows plus the whole halo on the first chunk
		int worst = 0, prev = band_lo - 1;
		for (int i = 0; i < nchunks; ++i) {
			const int c1 = bounds[i], c2 = bounds[i + 1];
			int lo, hi;
			source_rows(g, c1, c2, lo, hi);
			if (!g.monotone) { lo = band_lo; hi = band_hi; }
			worst = std::max(worst, hi - prev);
			prev = std::max(prev, hi);
		}
		in_slot = (size_t)std::max(worst, 1) * row_bytes;
		if ((rc = cx.reserve_pinned(cx.h_in, cx.h_in_cap, in_slot * ring))) return rc;
	}
	if (!dst_pinned) {
		out_slot = (size_t)chunk_rows * o_bytes;
		if ((rc = cx.reserve_pinned(cx.h_out, cx.h_out_cap, out_slot * ring))) return rc;
	}

	// worker threads for the staging copies of pageable callers (none for pinned callers or small bands)
	const bool staged = (!src_pinned || !dst_pinned) && (size_t)(y2 - y1) * row_bytes >= ((size_t)8 << 20);
	CopyPool pool(staged ? copy_threads() : 1);

	if (progress && g_progress)
		g_progress(0, 0.0, g_progress_user);

	int uploaded_hi = band_lo - 1;	// highest source row already sent
	std::vector<int> chunk_y1(nchunks), chunk_y2(nchunks);
	// events: [3*i] upload done, [3*i+1] kernel done, [3*i+2] download done
	auto retire = [&](int i) -> int {
		cudaEvent_t e_down = nullptr;
		int r = cx.event(3 * i + 2, e_down);
		if (r) return r;
		CUDA_TRY(cudaEventSynchronize(e_down));
		if (!dst_pinned) {
			const unsigned char *slot = cx.h_out + (size_t)(i % ring) * out_slot;
			if (o_bytes == o_row)
				pool.copy(dst + (size_t)(chunk_y1[i] - o_y0) * o_row, slot, (size_t)(chunk_y2[i] - chunk_y1[i]) * o_row);
			else
				for (int y = chunk_y1[i]; y < chunk_y2[i]; ++y)
					memcpy(dst + (size_t)(y - o_y0) * o_row + o_off, slot + (size_t)(y - chunk_y1[i]) * o_bytes, o_bytes);
		}
		if (progress && g_progress)
			for (int y = chunk_y1[i]; y < chunk_y2[i]; ++y)
				if ((y - y1) % 8 == 0)
					g_progress(1, (double)(y - y1) / (double)(y2 - y1), g_progress_user);
		return FIXCA_OK;
	};

	for (int i = 0; i < nchunks; ++i) {
		const int c1 = bounds[i], c2 = bounds[i + 1];
		chunk_y1[i] = c1;
		chunk_y2[i] = c2;
		int lo, hi;
		source_rows(g, c1, c2, lo, hi);
		if (!g.monotone) { lo = band_lo; hi = band_hi; }
		cudaEvent_t e_up = nullptr, e_run = nullptr, e_down = nullptr;
		if ((rc = cx.event(3 * i, e_up)) || (rc = cx.event(3 * i + 1, e_run)) || (rc = cx.event(3 * i + 2, e_down))) return rc;

		// the ring slot this chunk reuses must have been drained
		if (i >= ring && (rc = retire(i - ring))) return rc;

		if (hi > uploaded_hi) {
			const int r0 = uploaded_hi + 1, nr = hi - uploaded_hi;
			const unsigned char *from = src + (size_t)r0 * row_bytes;
			if (!src_pinned) {
				unsigned char *slot = cx.h_in + (size_t)(i % ring) * in_slot;
				pool.copy(slot, from, (size_t)nr * row_bytes);
				from = slot;
			}
			CUDA_TRY(cudaMemcpy2DAsync(cx.d_src + (size_t)(r0 - band_lo) * pitch, pitch, from, row_bytes,
						   row_bytes, nr, cudaMemcpyHostToDevice, cx.s_up));
			uploaded_hi = hi;
		}
		CUDA_TRY(cudaEventRecord(e_up, cx.s_up));
		CUDA_TRY(cudaStreamWaitEvent(cx.s_run, e_up, 0));
		Plan pl;
		if ((rc = make_plan(f, g, cx.d_src, pitch, band_lo, src_rows, cx.d_dst, pitch, y1, c1, c2, flags | FIXCA_PADDING_SCRATCH, dev, pl))) return rc;
		if ((rc = launch_plan(pl, cx.s_run))) return rc;
		if (flags & FIXCA_PREVIEW_OVERLAY) {
			CUDA_TRY(launch_preview(f.kind, f.nch, cx.d_dst, (long long)pitch, y1, c1, c2, width,
						(int)params->lens_x, (int)params->lens_y, params->saturation, cx.s_run));
			g_launches.fetch_add(1);
		}
		if (w8) {	// fix-ca.c:659-671: the window's samples as 8-bit
			CUDA_TRY(launch_to8(f.kind, f.nch, cx.d_dst, (long long)pitch, y1, c1, c2, w8->x, w8->pw, f.bpp,
					    cx.d_aux + (size_t)(c1 - y1) * o_row, cx.s_run));
			g_launches.fetch_add(1);
		}
		CUDA_TRY(cudaEventRecord(e_run, cx.s_run));
		CUDA_TRY(cudaStreamWaitEvent(cx.s_down, e_run, 0));
		unsigned char *to = dst_pinned ? dst + (size_t)(c1 - o_y0) * o_row + o_off : cx.h_out + (size_t)(i % ring) * out_slot;
		const unsigned char *from_dev = w8 ? cx.d_aux + (size_t)(c1 - y1) * o_row : cx.d_dst + (size_t)(c1 - y1) * pitch + sel_off;
		CUDA_TRY(cudaMemcpy2DAsync(to, dst_pinned ? o_row : o_bytes, from_dev, w8 ? o_row : pitch,
					   o_bytes, c2 - c1, cudaMemcpyDeviceToHost, cx.s_down));
		CUDA_TRY(cudaEventRecord(e_down, cx.s_down));
	}
	for (int i = std::max(0, nchunks - ring); i < nchunks; ++i)
		if ((rc = retire(i))) return rc;
	CUDA_TRY(cudaStreamSynchronize(cx.s_down));
	if (progress && g_progress)
		g_progress(1, 0.0, g_progress_user);
	(void)height;
	return FIXCA_OK;
}

static int host_prologue(const unsigned char *src, unsigned char *dst, int width, int height, int bytes, int bpc,
			 const fixca_params *params, int x1, int x2, int y1, int y2, Format &f, Geometry &g, unsigned flags,
			 int allow, const char *entry)
{
	int rc = check_common(src, dst, width, height, params, y1, y2);
	if (rc) return rc;
	if ((rc = parse_format(bytes, bpc, f))) return rc;
	if ((rc = check_flags(flags, allow, f, params, entry))) return rc;
	if ((x1 != 0 || x2 != width) && !(flags & FIXCA_COLUMN_SELECTION))
		return fail(FIXCA_ERR_REGION, "columns [%d,%d) of %d: only full-width row bands are defined (the reference's own x1 != 0 path is broken); FIXCA_COLUMN_SELECTION opts in to the repaired form", x1, x2, width);
	if (x1 < 0 || x2 > width || x1 >= x2)
		return fail(FIXCA_ERR_REGION, "columns [%d,%d) outside 0..%d", x1, x2, width);
	return make_geometry(width, height, params, g);
}

extern "C" int fixca_cuda_region_ex(const unsigned char *src, unsigned char *dst, int width, int height,
				    int bytes, int bpc, const fixca_params *params,
				    int x1, int x2, int y1, int y2, int show_progress, unsigned flags, int device)
{
	Format f;
	Geometry g;
	if (!show_progress)	// the preview call (fix-ca.c:656-657): saturation boost + centre lines on top
		flags |= FIXCA_PREVIEW_OVERLAY;
	int rc = host_prologue(src, dst, width, height, bytes, bpc, params, x1, x2, y1, y2, f, g, flags,
			       ALLOW_PREVIEW | ALLOW_COLUMNS, "fixca_cuda_region_ex");
	if (rc) return rc;
	int dev;
	if ((rc = current_device_or(device, dev))) return rc;
	if (y1 == y2)
		return FIXCA_OK;
	int prev = -1;
	cudaGetDevice(&prev);
	const auto t0 = std::chrono::steady_clock::now();
	rc = region_host_band(dev, src, dst, width, height, f, params, g, x1, x2, y1, y2, flags, show_progress != 0);
	tl_last_call_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
	if (prev >= 0)
		cudaSetDevice(prev);
	return rc;
}

// preview_update()'s middle in one call (fix-ca.c:656-671): the pass over rows [y, y + ph) with the preview overlay,
// then the 8-bit down-conversion of window columns [x, x + pw); only the window's bytes come back.
extern "C" int fixca_cuda_preview(const unsigned char *src, unsigned char *prev, int width, int height,
				  int bytes, int bpc, const fixca_params *params, int x, int y, int pw, int ph)
{
	Format f;
	Geometry g;
	if (x < 0 || pw <= 0 || x + pw > width || ph < 0)
		return fail(FIXCA_ERR_ARG, "preview window %d+%d x %d+%d outside the %dx%d image", x, pw, y, ph, width, height);
	const unsigned flags = (tuning().precision_fast ? FIXCA_PRECISION_FAST : FIXCA_PRECISION_EXACT) | FIXCA_PREVIEW_OVERLAY;
	int rc = host_prologue(src, prev, width, height, bytes, bpc, params, 0, width, y, y + ph, f, g, flags,
			       ALLOW_PREVIEW, "fixca_cuda_preview");
	if (rc) return rc;
	int dev;
	if ((rc = current_device_or(-1, dev))) return rc;
	if (ph == 0)
		return FIXCA_OK;
	int prevdev = -1;
	cudaGetDevice(&prevdev);
	const Window8 w8 = {x, pw};
	const auto t0 = std::chrono::steady_clock::now();
	rc = region_host_band(dev, src, prev, width, height, f, params, g, 0, width, y, y + ph, flags, false, &w8);
	tl_last_call_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
	if (prevdev >= 0)
		cudaSetDevice(prevdev);
	return rc;
}

extern "C" int fixca_cuda_region(const unsigned char *src, unsigned char *dst, int width, int height,
				 int bytes, int bpc, const fixca_params *params,
				 int x1, int x2, int y1, int y2, int show_progress)
{
	const unsigned flags = tuning().precision_fast ? FIXCA_PRECISION_FAST : FIXCA_PRECISION_EXACT;
	return fixca_cuda_region_ex(src, dst, width, height, bytes, bpc, params, x1, x2, y1, y2, show_progress, flags, -1);
}

extern "C" int fixca_split_bands(int y1, int y2, int nbands, int *band_y1, int *band_y2)
{
	if (nbands <= 0 || y1 > y2 || !band_y1 || !band_y2)
		return fail(FIXCA_ERR_ARG, "bad band split request");
	for (int i = 0; i < nbands; ++i) {
		band_y1[i] = y1 + (int)((long long)(y2 - y1) * i / nbands);
		band_y2[i] = y1 + (int)((long long)(y2 - y1) * (i + 1) / nbands);
	}
	return FIXCA_OK;
}

extern "C" int fixca_cuda_region_multi(const unsigned char *src, unsigned char *dst, int width, int height,
				       int bytes, int bpc, const fixca_params *params, int y1, int y2,
				       unsigned flags, const int *devices, int ndev)
{
	Format f;
	Geometry g;
	int rc = host_prologue(src, dst, width, height, bytes, bpc, params, 0, width, y1, y2, f, g, flags,
			       ALLOW_PREVIEW, "fixca_cuda_region_multi");
	if (rc) return rc;
	int have = 0;
	if (cudaGetDeviceCount(&have) != cudaSuccess || have <= 0) {
		cudaGetLastError();
		return fail(FIXCA_ERR_NO_DEVICE, "no CUDA device; this library has no CPU path");
	}
	if (ndev <= 0 || ndev > 16)
		return fail(FIXCA_ERR_ARG, "ndev %d outside 1..16", ndev);
	std::vector<int> dv(ndev), b1(ndev), b2(ndev), rcs(ndev, 0);
	std::vector<std::string> errs(ndev);
	for (int i = 0; i < ndev; ++i) {
		dv[i] = devices ? devices[i] : i;
		if (dv[i] < 0 || dv[i] >= have)
			return fail(FIXCA_ERR_NO_DEVICE, "device %d requested, %d present", dv[i], have);
	}
	fixca_split_bands(y1, y2, ndev, b1.data(), b2.data());
	std::vector<std::thread> workers;
	for (int i = 0; i < ndev; ++i) {
		if (b1[i] == b2[i])
			continue;
		workers.emplace_back([&, i]() {
			rcs[i] = region_host_band(dv[i], src, dst, width, height, f, params, g, 0, width, b1[i], b2[i], flags, false);
			if (rcs[i])
				errs[i] = tl_error;
		});
	}
	for (std::thread &t : workers)
		t.join();
	for (int i = 0; i < ndev; ++i)
		if (rcs[i])
			return fail(rcs[i], "band %d on device %d: %s", i, dv[i], errs[i].c_str());
	snprintf(tl_kernel, sizeof tl_kernel, "multi x%d", ndev);
	return FIXCA_OK;
}

extern "C" int fixca_cuda_frames(const unsigned char *const *src_frames, unsigned char *const *dst_frames, int nframes,
				 int width, int height, int bytes, int bpc, const fixca_params *params,
				 unsigned flags, int device)
{
	if (nframes < 0 || (nframes > 0 && (!src_frames || !dst_frames)))
		return fail(FIXCA_ERR_ARG, "bad frame list");
	if (nframes == 0)
		return FIXCA_OK;
	Format f;
	Geometry g;
	int rc = host_prologue(src_frames[0], dst_frames[0], width, height, bytes, bpc, params, 0, width, 0, height, f, g, flags,
			       0, "fixca_cuda_frames");
	if (rc) return rc;
	int dev;
	if ((rc = current_device_or(device, dev))) return rc;
	int prev = -1;
	cudaGetDevice(&prev);
	CUDA_TRY(cudaSetDevice(dev));

	// A ring of 3 slots, each with its own stream, device frame pair and (only for pageable caller memory)
	// pinned pair.  The ring lives as long as the library (fixca_cuda_release frees it): allocating streams,
	// device frames and pinned frames per call cost ~200 ms, 20x the PCIe time of a 16-frame 4K batch.
	const size_t row_bytes = (size_t)width * bytes, pitch = align_up(row_bytes, 128);
	const size_t frame_bytes = row_bytes * height, dev_bytes = pitch * height;
	if (dev >= 16) {
		if (prev >= 0) cudaSetDevice(prev);
		return fail(FIXCA_ERR_NO_DEVICE, "device %d: the frame ring serves devices 0..15", dev);
	}
	FrameRing &fr = g_frame_ring[dev];
	std::lock_guard<std::mutex> lock(fr.mu);
	const int ring = std::min(FrameRing::SLOTS, nframes);
	auto body = [&]() -> int {
		for (int k = 0; k < ring; ++k) {
			FrameRing::Slot &s = fr.slot[k];
			if (!s.s)
				CUDA_TRY(cudaStreamCreateWithFlags(&s.s, cudaStreamNonBlocking));
			if (s.dev_cap < dev_bytes) {
				if (s.d_src) cudaFree(s.d_src);
				if (s.d_dst) cudaFree(s.d_dst);
				s.d_src = s.d_dst = nullptr;
				s.dev_cap = 0;
				CUDA_TRY(cudaMalloc(&s.d_src, dev_bytes));
				CUDA_TRY(cudaMalloc(&s.d_dst, dev_bytes));
				s.dev_cap = dev_bytes;
			}
			s.frame = -1;
		}
		auto pinned_buf = [&](unsigned char *&p, size_t &cap) -> int {
			if (cap >= frame_bytes) return FIXCA_OK;
			if (p) cudaFreeHost(p);
			p = nullptr;
			cap = 0;
			CUDA_TRY(cudaHostAlloc(&p, frame_bytes, cudaHostAllocDefault));
			cap = frame_bytes;
			return FIXCA_OK;
		};
		auto retire = [&](FrameRing::Slot &s) -> int {
			if (s.frame < 0) return FIXCA_OK;
			CUDA_TRY(cudaStreamSynchronize(s.s));
			if (s.staged_out)
				memcpy(dst_frames[s.frame], s.h_out, frame_bytes);
			s.frame = -1;
			return FIXCA_OK;
		};
		for (int i = 0; i < nframes; ++i) {
			FrameRing::Slot &s = fr.slot[i % ring];
			int r = retire(s);
			if (r) return r;
			if (!src_frames[i] || !dst_frames[i])
				return fail(FIXCA_ERR_ARG, "frame %d is NULL", i);
			const unsigned char *from = src_frames[i];
			if (!is_pinned(from)) {
				if ((r = pinned_buf(s.h_in, s.h_in_cap))) return r;
				memcpy(s.h_in, from, frame_bytes);
				from = s.h_in;
			}
			s.staged_out = !is_pinned(dst_frames[i]);
			if (s.staged_out && (r = pinned_buf(s.h_out, s.h_out_cap))) return r;
			unsigned char *to = s.staged_out ? s.h_out : dst_frames[i];
			CUDA_TRY(cudaMemcpy2DAsync(s.d_src, pitch, from, row_bytes, row_bytes, height, cudaMemcpyHostToDevice, s.s));
			Plan pl;
			if ((r = make_plan(f, g, s.d_src, pitch, 0, height, s.d_dst, pitch, 0, 0, height, flags | FIXCA_PADDING_SCRATCH, dev, pl))) return r;
			if ((r = launch_plan(pl, s.s))) return r;
			CUDA_TRY(cudaMemcpy2DAsync(to, row_bytes, s.d_dst, pitch, row_bytes, height, cudaMemcpyDeviceToHost, s.s));
			s.frame = i;
		}
		for (int k = 0; k < ring; ++k) {
			int r = retire(fr.slot[k]);
			if (r) return r;
		}
		return FIXCA_OK;
	};
	rc = body();
	if (rc)		// leave no copy in flight behind an error
		for (int k = 0; k < FrameRing::SLOTS; ++k)
			if (fr.slot[k].s) cudaStreamSynchronize(fr.slot[k].s);
	if (prev >= 0) cudaSetDevice(prev);
	return rc;
}

// Frames sharded by index over several GPUs of the box from one process (BASELINE configs[4]: "batch stream of
// frames sharded across 8xB200"): device d of ndev takes frames d, d + ndev, ... through its own ring, stream
// set and PCIe link, on its own worker thread.  Frames are independent (fix-ca.c:373-374 per frame): no
// exchange between devices.
extern "C" int fixca_cuda_frames_multi(const unsigned char *const *src_frames, unsigned char *const *dst_frames, int nframes,
				       int width, int height, int bytes, int bpc, const fixca_params *params,
				       unsigned flags, const int *devices, int ndev)
{
	if (nframes < 0 || (nframes > 0 && (!src_frames || !dst_frames)))
		return fail(FIXCA_ERR_ARG, "bad frame list");
	if (ndev <= 0 || ndev > 16)
		return fail(FIXCA_ERR_ARG, "ndev %d outside 1..16", ndev);
	if (nframes == 0)
		return FIXCA_OK;
	int have = 0;
	if (cudaGetDeviceCount(&have) != cudaSuccess || have <= 0) {
		cudaGetLastError();
		return fail(FIXCA_ERR_NO_DEVICE, "no CUDA device; this library has no CPU path");
	}
	std::vector<int> dv(ndev), rcs(ndev, 0);
	std::vector<std::string> errs(ndev), kern(ndev);
	for (int i = 0; i < ndev; ++i) {
		dv[i] = devices ? devices[i] : i;
		if (dv[i] < 0 || dv[i] >= have)
			return fail(FIXCA_ERR_NO_DEVICE, "device %d requested, %d present", dv[i], have);
	}
	std::vector<std::vector<const unsigned char *>> srcs(ndev);
	std::vector<std::vector<unsigned char *>> dsts(ndev);
	for (int i = 0; i < nframes; ++i) {
		srcs[i % ndev].push_back(src_frames[i]);
		dsts[i % ndev].push_back(dst_frames[i]);
	}
	std::vector<std::thread> workers;
	for (int d = 0; d < ndev; ++d) {
		if (srcs[d].empty())
			continue;
		workers.emplace_back([&, d]() {
			rcs[d] = fixca_cuda_frames(srcs[d].data(), dsts[d].data(), (int)srcs[d].size(), width, height, bytes, bpc,
						   params, flags, dv[d]);
			if (rcs[d])
				errs[d] = tl_error;
			kern[d] = tl_kernel;
		});
	}
	for (std::thread &t : workers)
		t.join();
	for (int d = 0; d < ndev; ++d)
		if (rcs[d])
			return fail(rcs[d], "frames of device %d: %s", dv[d], errs[d].c_str());
	snprintf(tl_kernel, sizeof tl_kernel, "%s", kern[0].c_str());
	return FIXCA_OK;
}

// ---------------------------------------------------------------------------
// host-side logic
// ---------------------------------------------------------------------------
extern "C" int fixca_band_source_rows(int width, int height, const fixca_params *params, int y1, int y2,
				      int *src_lo, int *src_hi)
{
	if (!params || !src_lo || !src_hi || width <= 0 || height <= 0 || y1 < 0 || y2 > height || y1 >= y2)
		return fail(FIXCA_ERR_ARG, "bad arguments to fixca_band_source_rows");
	Geometry g;
	int rc = make_geometry(width, height, params, g);
	if (rc) return rc;
	source_rows(g, y1, y2, *src_lo, *src_hi);
	return FIXCA_OK;
}

extern "C" void fixca_resolve_lens(int width, int height, double *lens_x, double *lens_y)
{
	// fix-ca.c:427-428: round (xImg/2) with integer division
	if (lens_x && (*lens_x <= 0 || *lens_x >= width))
		*lens_x = (double)(width / 2);
	if (lens_y && (*lens_y <= 0 || *lens_y >= height))
		*lens_y = (double)(height / 2);
}

extern "C" int fixca_check_params(const fixca_params *p)
{
	if (!p)
		return fail(FIXCA_ERR_ARG, "NULL params");
	// fix-ca.c:279-292
	const double v[6] = {p->blue, p->red, p->x_blue, p->x_red, p->y_blue, p->y_red};
	for (double d : v)
		if (d < -FIXCA_INPUT_MAX || d > FIXCA_INPUT_MAX)
			return fail(FIXCA_ERR_RANGE, "Parameter out of range!");
	if (p->interpolation < 0 || p->interpolation > 2)
		return fail(FIXCA_ERR_INTERP, "Parameter out of range!");
	return FIXCA_OK;
}

extern "C" int fixca_color_size(const char *name, int bpp)
{
	// fix-ca.c:681-711
	if (!name)
		return FIXCA_BPC_UNSUPPORTED;
	if (strstr(name, "double")) return -8;
	if (strstr(name, "float")) return -4;
	if (strstr(name, "u15")) return FIXCA_BPC_UNSUPPORTED;
	if (!strstr(name, " u")) return FIXCA_BPC_UNSUPPORTED;
	if (bpp > 32) return FIXCA_BPC_UNSUPPORTED;
	if (bpp >= 24) return 8;
	if (bpp >= 12) return 4;
	if (bpp >= 6) return 2;
	if (bpp >= 3) return 1;
	return FIXCA_BPC_UNSUPPORTED;
}

extern "C" int fixca_color_size_half(const char *name, int bpp)
{
	// color_size() with the reference's commented-out half line (fix-ca.c:692-693) enabled, in its place:
	// after "double" and "float", before the unsigned-integer names
	if (name && !strstr(name, "double") && !strstr(name, "float") && strstr(name, "half"))
		return -2;
	return fixca_color_size(name, bpp);
}

extern "C" int fixca_color_size_ext(const char *name, int bpp)
{
	// color_size() with both "TODO for another day" rows answered: half as in fixca_color_size_half(), and
	// "u15" (fix-ca.c:694-695) -> FIXCA_BPC_U15 for RGB / RGBA of 16-bit storage
	if (name && !strstr(name, "double") && !strstr(name, "float") && !strstr(name, "half") && strstr(name, "u15"))
		return (bpp == 6 || bpp == 8) ? FIXCA_BPC_U15 : FIXCA_BPC_UNSUPPORTED;
	return fixca_color_size_half(name, bpp);
}

extern "C" void fixca_params_default(fixca_params *p)
{
	if (!p) return;
	// fix-ca.c:85-97
	memset(p, 0, sizeof *p);
	p->lens_x = -1.0;
	p->lens_y = -1.0;
	p->update_preview = 1;
	p->interpolation = FIXCA_INTERP_LINEAR;
}

extern "C" void fixca_cuda_set_progress(fixca_progress_fn fn, void *user)
{
	g_progress = fn;
	g_progress_user = user;
}

// Pinned (page-locked, portable) host memory for the caller's image buffers: what fix_ca() allocates with
// g_new at fix-ca.c:366-367 / :648-649.  A buffer from here is recognised by fixca_cuda_region*() and moved
// by DMA directly, without the staging copies pageable memory needs.  NULL when no GPU is usable or the
// allocation fails: the caller then falls back to its own allocator (and to its CPU path).
//
// Page-locking costs ~0.2 ms per MB, far more than the pass itself, and the plug-in allocates two whole-image
// buffers per run() and per preview refresh (:648-649, every slider move): freed buffers are kept in a small
// pool (at most 4 buffers / 4 GiB) and handed out again to requests they fit (within 2x); fixca_cuda_release()
// returns the pool to the system.
namespace {
struct PinnedPool {
	struct Buf { void *p; size_t cap; bool used; };
	std::mutex mu;
	std::vector<Buf> bufs;
};
PinnedPool g_pinned;
}

extern "C" void *fixca_cuda_host_alloc(size_t bytes)
{
	int dev;
	if (!bytes || current_device_or(-1, dev))
		return nullptr;
	{
		std::lock_guard<std::mutex> lock(g_pinned.mu);
		PinnedPool::Buf *best = nullptr;
		for (PinnedPool::Buf &b : g_pinned.bufs)
			if (!b.used && b.cap >= bytes && b.cap / 2 <= bytes && (!best || b.cap < best->cap))
				best = &b;
		if (best) {
			best->used = true;
			return best->p;
		}
	}
	void *p = nullptr;
	if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
		cudaGetLastError();
		fail(FIXCA_ERR_NOMEM, "cudaHostAlloc of %zu bytes failed", bytes);
		return nullptr;
	}
	std::lock_guard<std::mutex> lock(g_pinned.mu);
	g_pinned.bufs.push_back({p, bytes, true});
	return p;
}

extern "C" void fixca_cuda_host_free(void *p)
{
	if (!p)
		return;
	void *drop = nullptr;
	{
		std::lock_guard<std::mutex> lock(g_pinned.mu);
		size_t idle_bytes = 0, idle = 0;
		for (PinnedPool::Buf &b : g_pinned.bufs)
			if (!b.used) { idle_bytes += b.cap; ++idle; }
		for (size_t i = 0; i < g_pinned.bufs.size(); ++i) {
			PinnedPool::Buf &b = g_pinned.bufs[i];
			if (b.p != p)
				continue;
			if (idle < 4 && idle_bytes + b.cap <= ((size_t)4 << 30)) {
				b.used = false;		// kept for the next request
				return;
			}
			drop = b.p;
			g_pinned.bufs.erase(g_pinned.bufs.begin() + (long)i);
			break;
		}
		if (!drop)
			drop = p;	// not ours to pool (allocated before a release): still pinned memory
	}
	if (cudaFreeHost(drop) != cudaSuccess)
		cudaGetLastError();
}

static void pinned_pool_release()
{
	std::lock_guard<std::mutex> lock(g_pinned.mu);
	for (size_t i = 0; i < g_pinned.bufs.size();) {
		if (!g_pinned.bufs[i].used) {
			if (cudaFreeHost(g_pinned.bufs[i].p) != cudaSuccess)
				cudaGetLastError();
			g_pinned.bufs.erase(g_pinned.bufs.begin() + (long)i);
		} else {
			++i;
		}
	}
}

// wall-clock duration of the last fixca_cuda_region*() host call on this thread (what a plug-in spends inside
// the call that replaces its row loop)
extern "C" double fixca_cuda_last_call_ms(void) { return tl_last_call_ms; }

extern "C" const char *fixca_cuda_last_error(void) { return tl_error; }
extern "C" const char *fixca_cuda_last_kernel(void) { return tl_kernel; }
extern "C" long fixca_cuda_launch_count(void) { return g_launches.load(); }

extern "C" const char *fixca_strerror(int code)
{
	switch (code) {
	case FIXCA_OK: return "success";
	case FIXCA_ERR_ARG: return "invalid argument";
	case FIXCA_ERR_FORMAT: return "unsupported sample format";
	case FIXCA_ERR_INTERP: return "interpolation outside 0..2";
	case FIXCA_ERR_REGION: return "only full-width row bands are defined";
	case FIXCA_ERR_DEGENERATE: return "max_dim + amount == 0";
	case FIXCA_ERR_RANGE: return "parameter out of range";
	case FIXCA_ERR_NO_DEVICE: return "no CUDA device";
	case FIXCA_ERR_CUDA: return "CUDA error";
	case FIXCA_ERR_UNSUPPORTED: return "this entry point does not take that flag";
	case FIXCA_ERR_NOMEM: return "out of device or pinned memory";
	default: return "unknown error";
	}
}

extern "C" int fixca_cuda_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

extern "C" void fixca_cuda_release(void)
{
	pinned_pool_release();
	repair_queue_release();
	for (DeviceCtx &c : g_ctx) {
		std::lock_guard<std::mutex> lock(c.mu);
		c.release();
	}
	int prev = -1;
	cudaGetDevice(&prev);
	for (int d = 0; d < 16; ++d) {
		std::lock_guard<std::mutex> lock(g_frame_ring[d].mu);
		bool any = false;
		for (const FrameRing::Slot &s : g_frame_ring[d].slot)
			any = any || s.s || s.d_src;
		if (any && cudaSetDevice(d) == cudaSuccess)
			g_frame_ring[d].release();
	}
	if (prev >= 0) cudaSetDevice(prev);
}

extern "C" const char *fixca_version(void) { return "fixca-b200 0.1 (sm_100a)"; }
