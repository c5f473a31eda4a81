// fixca_kernels.cuh -- sm_100a kernels for Fix-CA's per-pixel correction pass
// (the reference's fix_ca_region row loop, fix-ca.c:1091-1333, and its helpers
// :713-920).  Included by the kernels_*.cu translation units, which instantiate
// the templates per sample format / interpolation / arithmetic.
//
// Two kernel families:
//
//   direct_kernel   one thread per output pixel, taps gathered from global
//                   memory through L1/L2.  Handles every geometry, including
//                   non-monotone maps and buffers TMA cannot address.  Fallback.
//
//   tiled_kernel    the fast path.  One CTA per TW x TH output tile:
//                     1. 8 threads evaluate the affine map at the tile corners ->
//                        source window (tile + halo); 2*TH threads fill the
//                        per-row coefficient table in shared memory;
//                     2. one warp issues a TMA bulk copy (cp.async.bulk ->
//                        UBLKCP) per window row, completing on an mbarrier,
//                        while every thread derives its own column taps and
//                        weights from the lens centre in registers;
//                     3. the tile's own pixels are copied window -> staging
//                        tile (green / alpha pass through, fix-ca.c:1094-1098);
//                     4. each thread owns one (column, channel) and walks down
//                        the tile: the horizontal pass of a source row is done
//                        once and kept in a 2- or 4-deep register window that
//                        slides with the (warp-uniform) source row, the
//                        vertical pass combines the window; results overwrite
//                        the R / B samples of the staging tile;
//                     5. the staging tile leaves with one TMA bulk store per row.
//                   Source coordinates are separable (x taps depend on x only,
//                   y taps on y only), which is what makes step 4 exact: the
//                   reference computes the same horizontal value for every
//                   output row that shares a source row (fix-ca.c:1300-1307).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include "fixca_geometry.h"

namespace fixca {

// ---------------------------------------------------------------------------
// Launch arguments
// ---------------------------------------------------------------------------
struct KernelArgs {
	const unsigned char *src;	// row src_row0 of the image
	unsigned char       *dst;	// row dst_row0 of the image
	long long src_pitch, dst_pitch;
	int  src_row0, dst_row0;
	int  y1, y2;			// output rows [y1, y2)
	Geometry g;
	// tiled kernel only
	int  th;			// tile height (rows)
	int  win_pitch;			// shared-memory window row pitch, bytes (multiple of 16)
	int  win_rows;			// window rows allocated
	int  off_ytab, off_win, off_out;// byte offsets into dynamic shared memory
	// strip kernel only
	int  off_nemit;			// per-source-row emit counts, u8[2][ne_pitch]
	int  ne_pitch;			// multiple of 4, >= win_rows + 4
	// stream kernel only (off_ytab = metadata ring, off_win = window ring, off_out = staging ring)
	int  seg_rows;			// output rows per CTA (multiple of the chunk height)
	int  ring_rows;			// window ring capacity in rows (multiple of 4)
	int  depth;			// chunks prefetched ahead (pipeline depth D)
	int  tile_lead;			// chunks the pass-through tile + record are requested ahead: 2 (needs depth >= 2) or 1
	int  debug;			// bit 0: skip the arithmetic (timing experiments only)
	int  off_rq;			// exact-repair stream kernels: per-warp queues of near-tie samples (u16 entries)
	// per-plan tables in global memory, one record per 8-row chunk of [y1, y2) (stream_meta_kernel fills them)
	const void *meta_tab;		// StreamMeta / StreamMetaWide records: vertical weights and tap rows of the chunk's rows
	const void *span_tab;		// StreamSpan records: first / last source row the chunk touches
	// None: per-plan column table (stream_cols_kernel): [2][col_n] nearest source column, col_n = the strips' column range
	const int  *col_i0;
	int  col_n;
	// Linear / Cubic: per-plan column set-up (stream_setup_kernel): one StreamColumnState per strip and compute thread
	const void *setup_tab;
	// exact-repair stream kernels, deferred form: every compute warp appends its near-tie samples to its own
	// region of a queue in global memory (no atomics); repair_patch_kernel, launched behind the stream kernel, recomputes them
	unsigned long long *rq_entries;	// [regions][rq_cap] linear sample indices ((frame * rows + y - y1) * width + x) * 2 + channel
	unsigned *rq_ctl;		// [regions] samples the warp flagged (> rq_cap: the patch kernel recomputes the warp's whole region)
	unsigned  rq_cap;		// region = (CTA of the grid, x fastest) * compute warps + warp
};

// 15-bit unsigned samples in 16-bit storage (babl's "u15": 0 .. 32768 <-> [0.0, 1.0]).  The reference rejects
// them ("TODO for another day", fix-ca.c:694-695); the extension (bpc code 15) gives them the
// arithmetic of the reference's other unsigned types with max = 32768: get_pixel v / 32768 (exact: a power
// of two), set_pixel round(d * 32768) after clip_d.  Stored values above 32768 decode to > 1.0 and are
// clipped like float samples.  A distinct type so that templates tell it from uint16_t.
struct u15_t { uint16_t v; };

// ---------------------------------------------------------------------------
// u64 samples: the reference decodes them through the x87's 80-bit long double (fix-ca.c:728-733) and encodes them
// through `roundl(d * 18446744073709551615UL)` (:759-761).  Both are restated in integer / FP64 arithmetic:
//   get_pixel  ret = RN53(RN64(v / (2^64 - 1))).  v / (2^64 - 1) = 0.VVVV... in base 2^64, so the 64 bits that
//              follow the leading one are rotl(v, clz(v)) and what follows repeats them: the round bit is the
//              top bit of that rotation -- always set -- and the sticky bits are never all zero, so the 64-bit
//              mantissa is rotl(v, clz(v)) + 1 (v = 2^64 - 1 carries to exactly 1.0); then one rounding to 53
//              bits, nearest-even.
//   set_pixel  the constant 2^64 - 1 converts to the DOUBLE 2^64, the product is exact, roundl rounds half away,
//              and the conversion of 2^64 itself (d = 1.0) to uint64_t is out of range: the compiled reference's
//              x87 sequence (subtract 2^63, fistp, flip the top bit) yields 0 -- white wraps to black, and so it
//              does here (checked against oracle/_ref, tests/golden/golden_u64.json).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double u64_get_pixel(uint64_t v)
{
	if (v == 0)
		return 0.0;
	const int lz = __clzll((long long)v);
	const uint64_t r = lz ? (v << lz) | (v >> (64 - lz)) : v;
	const double scale = __hiloint2double((1023 - 53 - lz) << 20, 0);	// 2^(11 - 64 - lz)
	if (r == ~0ull)							// mantissa 2^64: v / (2^64 - 1) == 1 (lz == 0)
		return 1.0;
	const uint64_t n = r + 1;					// the x87 quotient's mantissa, top bit set
	uint64_t m = n >> 11;
	const uint64_t rem = n & 0x7FF;
	if (rem > 0x400 || (rem == 0x400 && (m & 1)))
		++m;							// <= 2^53: exact as a double
	return __dmul_rn((double)m, scale);
}
__device__ __forceinline__ uint64_t u64_set_pixel(double d)	// d already through clip_d
{
	const double r = round(__dmul_rn(d, 18446744073709551616.0));
	return r >= 18446744073709551616.0 ? 0ull : __double2ull_rz(r);
}

// ---------------------------------------------------------------------------
// Arithmetic policies
// ---------------------------------------------------------------------------
template <class S> struct SampleMax;
template <> struct SampleMax<uint8_t>  { static constexpr double value = 255.0; };
template <> struct SampleMax<uint16_t> { static constexpr double value = 65535.0; };
template <> struct SampleMax<uint32_t> { static constexpr double value = 4294967295.0; };

// Exact: FP64, the reference's operation order, one rounding per operation.
// get_pixel (fix-ca.c:713-746), clip_d + set_pixel (fix-ca.c:873-880, :748-774),
// bilinear (fix-ca.c:882-894), cubicY / cubicX (fix-ca.c:896-920).
struct ExactF64 {
	typedef double acc_t;
	static constexpr const char *name = "f64";

	struct XCoef { double t; };
	struct YCoef { int i0; int pad; double t; };

	__device__ __forceinline__ static XCoef make_x(double t, int /*interp*/) { return XCoef{t}; }
	__device__ __forceinline__ static YCoef make_y(int i0, double t, int /*interp*/) { return YCoef{i0, 0, t}; }

	// get_pixel's true division v / max (fix-ca.c:717-722) without the division subroutine: q0 = v * RN(1/max),
	// one exact residual, one correction (Markstein).  Correctly rounded for EVERY 8- and 16-bit sample:
	// tests/test_host_logic.py checks all 256 + 65536 values against IEEE division in exact rational arithmetic.
	// The sample becomes a double through the 2^52 trick (no I2F.F64 on the XU pipe).
	template <int MAX>
	__device__ __forceinline__ static double div_by_max(unsigned v)
	{
		const double x = __dsub_rn(__hiloint2double(0x43300000, (int)v), 4503599627370496.0);	// exact: 2^52 + v - 2^52
		constexpr double r = 1.0 / (double)MAX;		// correctly rounded reciprocal
		const double q0 = __dmul_rn(x, r);
		const double e = __fma_rn(-(double)MAX, q0, x);	// exact residual
		return __fma_rn(e, r, q0);
	}
	__device__ __forceinline__ static double decode(uint8_t v)  { return div_by_max<255>(v); }
	__device__ __forceinline__ static double decode(uint16_t v) { return div_by_max<65535>(v); }
	__device__ __forceinline__ static double decode(uint32_t v) { return __ddiv_rn((double)v, 4294967295.0); }
	__device__ __forceinline__ static double decode(uint64_t v) { return u64_get_pixel(v); }
	__device__ __forceinline__ static double decode(u15_t v)    { return __dmul_rn((double)v.v, 0x1p-15); }	// v / 32768, exact
	__device__ __forceinline__ static double decode(float v)    { return (double)v; }
	__device__ __forceinline__ static double decode(double v)   { return v; }
	// half: the reference's commented-out branch `ret += *p` (fix-ca.c:740-742): an exact widening
	__device__ __forceinline__ static double decode(__half v)   { return (double)__half2float(v); }

	__device__ __forceinline__ static double clip(double d)
	{
		if (d <= 0.0) return 0.0;
		if (d >= 1.0) return 1.0;
		return d;	// NaN passes through, as in the reference
	}
	__device__ __forceinline__ static void encode(uint8_t &o, double d)  { o = (uint8_t)__double2uint_rz(round(__dmul_rn(clip(d), 255.0))); }
	__device__ __forceinline__ static void encode(uint16_t &o, double d) { o = (uint16_t)__double2uint_rz(round(__dmul_rn(clip(d), 65535.0))); }
	__device__ __forceinline__ static void encode(uint32_t &o, double d) { o = __double2uint_rz(round(__dmul_rn(clip(d), 4294967295.0))); }
	__device__ __forceinline__ static void encode(uint64_t &o, double d) { o = u64_set_pixel(clip(d)); }
	__device__ __forceinline__ static void encode(u15_t &o, double d)    { o.v = (uint16_t)__double2uint_rz(round(__dmul_rn(clip(d), 32768.0))); }
	__device__ __forceinline__ static void encode(float &o, double d)    { o = __double2float_rn(clip(d)); }
	__device__ __forceinline__ static void encode(double &o, double d)   { o = clip(d); }
	// `*p = d` (fix-ca.c:768-770, commented out there): one rounding, double -> half, nearest-even
	__device__ __forceinline__ static void encode(__half &o, double d)   { o = __double2half(clip(d)); }

	// p0 + t * (p1 - p0)
	__device__ __forceinline__ static double hlin(double p0, double p1, const XCoef &c)
	{
		return __dadd_rn(p0, __dmul_rn(c.t, __dsub_rn(p1, p0)));
	}
	// (1 - t) * h0 + t * h1
	__device__ __forceinline__ static double vlin(double h0, double h1, const YCoef &c)
	{
		return __dadd_rn(__dmul_rn(__dsub_rn(1.0, c.t), h0), __dmul_rn(c.t, h1));
	}
	// Catmull-Rom, Horner form, exactly as written at fix-ca.c:905-907.
	__device__ __forceinline__ static double catmull(double m1, double x, double p1, double p2, double t)
	{
		const double a = __dadd_rn(__dsub_rn(__dadd_rn(-m1, __dmul_rn(3.0, x)), __dmul_rn(3.0, p1)), p2);
		const double b = __dsub_rn(__dadd_rn(__dsub_rn(__dmul_rn(2.0, m1), __dmul_rn(5.0, x)), __dmul_rn(4.0, p1)), p2);
		const double c = __dadd_rn(-m1, p1);
		double r = __dadd_rn(__dmul_rn(a, t), b);
		r = __dadd_rn(__dmul_rn(r, t), c);
		r = __dadd_rn(__dmul_rn(r, t), __dadd_rn(x, x));
		return __dmul_rn(r, 0.5);	// "/ 2.0": exact either way
	}
	__device__ __forceinline__ static double hcub(double a, double b, double c, double d, const XCoef &k) { return catmull(a, b, c, d, k.t); }
	__device__ __forceinline__ static double vcub(double a, double b, double c, double d, const YCoef &k) { return catmull(a, b, c, d, k.t); }
};

// Fast: FP32 on raw sample values with separable weights (FMA allowed).  The
// division by / multiplication with the sample maximum cancels, so integer
// samples stay in LSB units end to end.  Within +-1 LSB of ExactF64 for u8/u16
// and ~2 ulp(1.0) for float.
struct FastF32 {
	typedef float acc_t;
	static constexpr const char *name = "f32";

	struct XCoef { float w0, w1, w2, w3; };
	struct __align__(16) YCoef { float w0, w1, w2, w3; int i0; int pad[3]; };

	__device__ __forceinline__ static void weights(double t, int interp, float &w0, float &w1, float &w2, float &w3)
	{
		if (interp == 1) {
			w0 = (float)(1.0 - t); w1 = (float)t; w2 = 0.f; w3 = 0.f;
		} else {
			// Catmull-Rom basis of fix-ca.c:905-907, collected per tap.
			w0 = (float)(((-t + 2.0) * t - 1.0) * t * 0.5);
			w1 = (float)(((3.0 * t - 5.0) * t * t + 2.0) * 0.5);
			w2 = (float)(((-3.0 * t + 4.0) * t + 1.0) * t * 0.5);
			w3 = (float)((t - 1.0) * t * t * 0.5);
		}
	}
	__device__ __forceinline__ static XCoef make_x(double t, int interp)
	{
		XCoef c; weights(t, interp, c.w0, c.w1, c.w2, c.w3); return c;
	}
	__device__ __forceinline__ static YCoef make_y(int i0, double t, int interp)
	{
		YCoef c; weights(t, interp, c.w0, c.w1, c.w2, c.w3); c.i0 = i0; c.pad[0] = c.pad[1] = c.pad[2] = 0; return c;
	}

	__device__ __forceinline__ static float decode(uint8_t v)  { return (float)v; }
	__device__ __forceinline__ static float decode(uint16_t v) { return (float)v; }
	__device__ __forceinline__ static float decode(u15_t v)    { return (float)v.v; }
	__device__ __forceinline__ static float decode(float v)    { return v; }
	__device__ __forceinline__ static float decode(__half v)   { return __half2float(v); }

	// clip_d's order: <= 0 first, then >= max; NaN passes (float images only).
	__device__ __forceinline__ static void encode(uint8_t &o, float d)  { o = (uint8_t)__float2uint_rn(fminf(fmaxf(d, 0.f), 255.f)); }
	__device__ __forceinline__ static void encode(uint16_t &o, float d) { o = (uint16_t)__float2uint_rn(fminf(fmaxf(d, 0.f), 65535.f)); }
	__device__ __forceinline__ static void encode(u15_t &o, float d)    { o.v = (uint16_t)__float2uint_rn(fminf(fmaxf(d, 0.f), 32768.f)); }
	__device__ __forceinline__ static void encode(float &o, float d)    { o = (d <= 0.f) ? 0.f : ((d >= 1.f) ? 1.f : d); }
	__device__ __forceinline__ static void encode(__half &o, float d)   { o = __float2half_rn((d <= 0.f) ? 0.f : ((d >= 1.f) ? 1.f : d)); }

	__device__ __forceinline__ static float hlin(float p0, float p1, const XCoef &c) { return fmaf(c.w1, p1, c.w0 * p0); }
	__device__ __forceinline__ static float vlin(float h0, float h1, const YCoef &c) { return fmaf(c.w1, h1, c.w0 * h0); }
	__device__ __forceinline__ static float hcub(float a, float b, float c, float d, const XCoef &k)
	{
		return fmaf(k.w3, d, fmaf(k.w2, c, fmaf(k.w1, b, k.w0 * a)));
	}
	__device__ __forceinline__ static float vcub(float a, float b, float c, float d, const YCoef &k)
	{
		return fmaf(k.w3, d, fmaf(k.w2, c, fmaf(k.w1, b, k.w0 * a)));
	}
};

// ---------------------------------------------------------------------------
// direct kernel
// ---------------------------------------------------------------------------
template <class S>
__device__ __forceinline__ const S *src_row_ptr(const KernelArgs &a, int r)
{
	return reinterpret_cast<const S *>(a.src + (long long)(r - a.src_row0) * a.src_pitch);
}

// None: raw sample copies, no arithmetic on the data (fix-ca.c:1105-1120).
// U is an unsigned integer of the sample's byte size.
template <class U, int NCH>
__global__ void __launch_bounds__(256) direct_none_kernel(const __grid_constant__ KernelArgs a)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	const int y = a.y1 + blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= a.g.width || y >= a.y2)
		return;
	const U *own = src_row_ptr<U>(a, y) + (size_t)x * NCH;
	U *out = reinterpret_cast<U *>(a.dst + (long long)(y - a.dst_row0) * a.dst_pitch) + (size_t)x * NCH;
	out[1] = own[1];
	if (NCH == 4)
		out[3] = own[3];
#pragma unroll
	for (int c = 0; c < 2; ++c) {
		const int sx = nearest_index(a.g.x[c], x);
		const int sy = nearest_index(a.g.y[c], y);
		out[2 * c] = src_row_ptr<U>(a, sy)[(size_t)sx * NCH + 2 * c];
	}
}

// One output sample of channel c at (x, y) in the reference's own (non-shared) form, fix-ca.c:1135-1186 (Linear) /
// :1204-1320 (Cubic): coordinates, clamp-to-edge taps, horizontal pass per tap row, vertical pass, clip + encode.
// fetch(row, col) returns sample c of the source pixel at absolute (row, col).  Used by direct_kernel (global
// memory) and by the streaming kernel's exact repair of near-tie samples (window ring in shared memory).
template <class S, int INTERP, class A, class Fetch>
__device__ __forceinline__ S interp_sample(const Geometry &g, int c, int x, int y, Fetch fetch)
{
	const int W = g.width, H = g.height;
	double tx, ty;
	const int cx = base_index(g.x[c], x, tx);
	const int cy = base_index(g.y[c], y, ty);
	const typename A::XCoef kx = A::make_x(tx, INTERP);
	const typename A::YCoef ky = A::make_y(cy, ty, INTERP);
	typename A::acc_t r;
	if (INTERP == 1) {
		const int c1 = cx + 1 < W ? cx + 1 : W - 1;
		const int r1 = cy + 1 < H ? cy + 1 : H - 1;
		const typename A::acc_t h0 = A::hlin(A::decode(fetch(cy, cx)), A::decode(fetch(cy, c1)), kx);
		const typename A::acc_t h1 = A::hlin(A::decode(fetch(r1, cx)), A::decode(fetch(r1, c1)), kx);
		r = A::vlin(h0, h1, ky);
	} else {
		const int q0 = cx > 0 ? cx - 1 : 0, q2 = cx + 1 < W ? cx + 1 : W - 1, q3 = cx + 2 < W ? cx + 2 : W - 1;
		typename A::acc_t h[4];
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			const int row = clampi(cy - 1 + k, 0, H - 1);
			h[k] = A::hcub(A::decode(fetch(row, q0)), A::decode(fetch(row, cx)), A::decode(fetch(row, q2)),
				       A::decode(fetch(row, q3)), kx);
		}
		r = A::vcub(h[0], h[1], h[2], h[3], ky);
	}
	S out;
	A::encode(out, r);
	return out;
}

// Linear / Cubic, per-pixel evaluation in the reference's own (non-shared) form.
template <class S, int NCH, int INTERP, class A>
__global__ void __launch_bounds__(256) direct_kernel(const __grid_constant__ KernelArgs a)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	const int y = a.y1 + blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= a.g.width || y >= a.y2)
		return;
	const S *own = src_row_ptr<S>(a, y) + (size_t)x * NCH;
	S *out = reinterpret_cast<S *>(a.dst + (long long)(y - a.dst_row0) * a.dst_pitch) + (size_t)x * NCH;
	out[1] = own[1];
	if (NCH == 4)
		out[3] = own[3];
#pragma unroll
	for (int c = 0; c < 2; ++c)
		out[2 * c] = interp_sample<S, INTERP, A>(a.g, c, x, y, [&](int row, int col) {
			return src_row_ptr<S>(a, row)[(size_t)col * NCH + 2 * c];
		});
}

// ---------------------------------------------------------------------------
// TMA bulk copy + mbarrier wrappers (PTX; SASS: UBLKCP, SYNCS)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
	return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem()
{
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	uint32_t done;
	do {
		asm volatile(
			"{\n\t.reg .pred p;\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			"selp.u32 %0, 1, 0, p;\n\t}"
			: "=r"(done)
			: "r"(smem_u32(bar)), "r"(parity)
			: "memory");
	} while (!done);
}
// global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		     ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
		     : "memory");
}
// shared -> global
__device__ __forceinline__ void bulk_store(void *gmem_dst, const void *smem_src, uint32_t bytes)
{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
		     ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
		     : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---------------------------------------------------------------------------
// tiled kernel
// ---------------------------------------------------------------------------
struct TileHeader {
	unsigned long long bar;
	int col_lo[4], col_hi[4];	// tap ranges at the tile's first / last column, per channel
	int row_lo[4], row_hi[4];
};

// Per-row entry of the None table: just the two source rows.
struct NoneRow { int row; };

// Dynamic shared memory: [TileHeader | ytab[2][th] | window | staging tile].
//
// S      sample type (for None: an unsigned integer of the sample's size)
// NCH    3 or 4 samples per pixel
// INTERP 0 / 1 / 2
// A      arithmetic policy (ignored for None)
// TW     tile width in pixels; blockDim.x == 2 * TW (red half, blue half)
template <class S, int NCH, int INTERP, class A, int TW>
__global__ void __launch_bounds__(2 * TW) tiled_kernel(const __grid_constant__ KernelArgs a)
{
	extern __shared__ __align__(128) unsigned char smem[];
	constexpr int BPP = NCH * (int)sizeof(S);
	constexpr int OUT_PITCH = TW * BPP;
	typedef typename A::YCoef YCoef;
	typedef typename A::XCoef XCoef;
	typedef typename A::acc_t acc_t;

	TileHeader *hdr = reinterpret_cast<TileHeader *>(smem);
	YCoef *ytab = reinterpret_cast<YCoef *>(smem + a.off_ytab);	// [2][th]
	unsigned char *win = smem + a.off_win;
	unsigned char *outt = smem + a.off_out;

	const int tid = threadIdx.x;
	const int c = tid / TW;		// 0 red, 1 blue: uniform per warp (TW % 32 == 0)
	const int lx = tid - c * TW;
	const int W = a.g.width, H = a.g.height;
	const int x0 = blockIdx.x * TW;
	const int y0 = a.y1 + blockIdx.y * a.th;
	const int xl = min(x0 + TW, W) - 1;		// last column of the tile
	const int yl = min(y0 + a.th, a.y2) - 1;	// last row of the tile
	const int nrows_out = yl - y0 + 1;

	// ---- 1. window extent (8 lanes) and per-row coefficients ----
	if (tid == 0) {
		mbar_init(reinterpret_cast<uint64_t *>(&hdr->bar), 1);
		fence_mbar_init();
	}
	if (tid < 8) {
		const int ch = tid & 1, last = (tid >> 1) & 1, isrow = tid >> 2;
		const Axis &ax = isrow ? a.g.y[ch] : a.g.x[ch];
		const int i = isrow ? (last ? yl : y0) : (last ? xl : x0);
		int lo, hi;
		tap_range(ax, INTERP, i, lo, hi);
		int *plo = isrow ? hdr->row_lo : hdr->col_lo;
		int *phi = isrow ? hdr->row_hi : hdr->col_hi;
		plo[tid & 3] = lo;
		phi[tid & 3] = hi;
	}
	for (int k = tid; k < 2 * nrows_out; k += 2 * TW) {
		const int ch = k >= nrows_out;
		const int r = k - ch * nrows_out;
		if (INTERP == 0) {
			reinterpret_cast<NoneRow *>(ytab)[ch * a.th + r].row = nearest_index(a.g.y[ch], y0 + r);
		} else {
			double t;
			const int i0 = base_index(a.g.y[ch], y0 + r, t);
			ytab[ch * a.th + r] = A::make_y(i0, t, INTERP);
		}
	}
	__syncthreads();

	int col_lo = x0, col_hi = xl, row_lo = y0, row_hi = yl;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		col_lo = min(col_lo, hdr->col_lo[k]);
		col_hi = max(col_hi, hdr->col_hi[k]);
		row_lo = min(row_lo, hdr->row_lo[k]);
		row_hi = max(row_hi, hdr->row_hi[k]);
	}
	const int wb0 = (col_lo * BPP) & ~15;			// window start, bytes into the row
	const int wbytes = (((col_hi + 1) * BPP + 15) & ~15) - wb0;	// <= src_pitch - wb0 (pitch % 16 == 0)
	const int wrows = row_hi - row_lo + 1;
	const int wpitch = a.win_pitch;

	// ---- 2. TMA: one bulk copy per window row ----
	uint64_t *bar = reinterpret_cast<uint64_t *>(&hdr->bar);
	if (tid < 32) {
		if (tid == 0)
			mbar_arrive_expect_tx(bar, (uint32_t)(wrows * wbytes));
		__syncwarp();
		const unsigned char *g = a.src + (long long)(row_lo - a.src_row0) * a.src_pitch + wb0;
		for (int r = tid; r < wrows; r += 32)
			bulk_load(win + r * wpitch, g + (long long)r * a.src_pitch, (uint32_t)wbytes, bar);
	}

	// ---- per-thread column state, overlapped with the copies ----
	const int x = x0 + lx;
	const bool active = x <= xl;
	const int choff = 2 * c * (int)sizeof(S);
	int o0 = 0, o1 = 0, o2 = 0, o3 = 0;
	XCoef kx;
	if (INTERP == 0) {
		o0 = nearest_index(a.g.x[c], active ? x : xl) * BPP + choff - wb0;
	} else {
		double t;
		const int cx = base_index(a.g.x[c], active ? x : xl, t);
		kx = A::make_x(t, INTERP);
		if (INTERP == 1) {
			o0 = cx * BPP + choff - wb0;
			o1 = (cx + 1 < W ? cx + 1 : W - 1) * BPP + choff - wb0;
		} else {
			o0 = (cx > 0 ? cx - 1 : 0) * BPP + choff - wb0;
			o1 = cx * BPP + choff - wb0;
			o2 = (cx + 1 < W ? cx + 1 : W - 1) * BPP + choff - wb0;
			o3 = (cx + 2 < W ? cx + 2 : W - 1) * BPP + choff - wb0;
		}
	}

	mbar_wait(bar, 0);

	// ---- 3. green / alpha: the tile's own pixels, window -> staging tile ----
	{
		const int row_bytes = ((xl - x0 + 1) * BPP + 15) & ~15;
		const int vec_per_row = row_bytes >> 4;
		const int src_off = x0 * BPP - wb0;	// multiple of 16: TW * BPP % 16 == 0
		for (int r = tid >> 5; r < nrows_out; r += 2 * TW / 32)		// a warp per row (no division per vector)
			for (int v = tid & 31; v < vec_per_row; v += 32) {
				const int4 q = *reinterpret_cast<const int4 *>(win + (y0 + r - row_lo) * wpitch + src_off + v * 16);
				*reinterpret_cast<int4 *>(outt + r * OUT_PITCH + v * 16) = q;
			}
	}
	__syncthreads();

	// ---- 4. walk down the tile ----
	// (Measured and dropped, r02: the channel's 128 threads decoding each source row once into a double buffer in
	// shared memory -- get_pixel's division is 4 FP64 operations and every sample is a tap of four columns -- with a
	// named barrier per row: 100 MP RGB16 Cubic 1.19 -> 1.42 ms, Linear 0.58 -> 1.08 ms.  The barrier costs more
	// than the 12 FP64 operations per row it saves.)
	if (active) {
		unsigned char *op = outt + lx * BPP + choff;
		if (INTERP == 0) {
			const NoneRow *yt = reinterpret_cast<const NoneRow *>(ytab) + c * a.th;
			for (int r = 0; r < nrows_out; ++r) {
				const S v = *reinterpret_cast<const S *>(win + (yt[r].row - row_lo) * wpitch + o0);
				*reinterpret_cast<S *>(op + r * OUT_PITCH) = v;
			}
		} else if (INTERP == 1) {
			const YCoef *yt = ytab + c * a.th;
			acc_t h0 = 0, h1 = 0;
			int base = 0;
			for (int r = 0; r < nrows_out; ++r) {
				const YCoef ky = yt[r];
				const int i0 = ky.i0;
				if (r == 0 || i0 > base + 1) {
					// (re)load both rows
					const unsigned char *p0 = win + (i0 - row_lo) * wpitch;
					const unsigned char *p1 = win + (min(i0 + 1, H - 1) - row_lo) * wpitch;
					h0 = A::hlin(A::decode(*reinterpret_cast<const S *>(p0 + o0)), A::decode(*reinterpret_cast<const S *>(p0 + o1)), kx);
					h1 = A::hlin(A::decode(*reinterpret_cast<const S *>(p1 + o0)), A::decode(*reinterpret_cast<const S *>(p1 + o1)), kx);
					base = i0;
				} else if (i0 == base + 1) {
					const unsigned char *p1 = win + (min(i0 + 1, H - 1) - row_lo) * wpitch;
					h0 = h1;
					h1 = A::hlin(A::decode(*reinterpret_cast<const S *>(p1 + o0)), A::decode(*reinterpret_cast<const S *>(p1 + o1)), kx);
					base = i0;
				}
				S v;
				A::encode(v, A::vlin(h0, h1, ky));
				*reinterpret_cast<S *>(op + r * OUT_PITCH) = v;
			}
		} else {
			const YCoef *yt = ytab + c * a.th;
			acc_t h0 = 0, h1 = 0, h2 = 0, h3 = 0;
			int base = 0;	// window holds source rows base-1 .. base+2 (clamped to the image)
			auto hrow = [&](int srow) -> acc_t {
				const unsigned char *p = win + (clampi(srow, 0, H - 1) - row_lo) * wpitch;
				return A::hcub(A::decode(*reinterpret_cast<const S *>(p + o0)),
					       A::decode(*reinterpret_cast<const S *>(p + o1)),
					       A::decode(*reinterpret_cast<const S *>(p + o2)),
					       A::decode(*reinterpret_cast<const S *>(p + o3)), kx);
			};
			for (int r = 0; r < nrows_out; ++r) {
				const YCoef ky = yt[r];
				const int i0 = ky.i0;
				if (r == 0 || i0 > base + 3) {
					h0 = hrow(i0 - 1);
					h1 = hrow(i0);
					h2 = hrow(i0 + 1);
					h3 = hrow(i0 + 2);
					base = i0;
				} else {
					while (base < i0) {	// warp-uniform: depends on (row, channel) only
						h0 = h1; h1 = h2; h2 = h3;
						h3 = hrow(base + 3);
						++base;
					}
				}
				S v;
				A::encode(v, A::vcub(h0, h1, h2, h3, ky));
				*reinterpret_cast<S *>(op + r * OUT_PITCH) = v;
			}
		}
	}

	// ---- 5. staging tile -> global, one bulk store per row ----
	fence_proxy_async_smem();
	__syncthreads();
	if (tid < 32) {
		const int row_bytes = ((xl - x0 + 1) * BPP + 15) & ~15;
		unsigned char *g = a.dst + (long long)(y0 - a.dst_row0) * a.dst_pitch + (long long)x0 * BPP;
		for (int r = tid; r < nrows_out; r += 32)
			bulk_store(g + (long long)r * a.dst_pitch, outt + r * OUT_PITCH, (uint32_t)row_bytes);
		bulk_commit();
		bulk_wait_read_all();
	}
}

// Policy placeholder for None instantiations of tiled_kernel (never evaluated).
struct NoArith {
	typedef int acc_t;
	static constexpr const char *name = "copy";
	struct XCoef { int unused; };
	struct YCoef { int i0; };
	__device__ __forceinline__ static XCoef make_x(double, int) { return XCoef{0}; }
	__device__ __forceinline__ static YCoef make_y(int i0, double, int) { return YCoef{i0}; }
	template <class S> __device__ __forceinline__ static int decode(S) { return 0; }
	template <class S> __device__ __forceinline__ static void encode(S &, int) {}
	__device__ __forceinline__ static int hlin(int, int, const XCoef &) { return 0; }
	__device__ __forceinline__ static int vlin(int, int, const YCoef &) { return 0; }
	__device__ __forceinline__ static int hcub(int, int, int, int, const XCoef &) { return 0; }
	__device__ __forceinline__ static int vcub(int, int, int, int, const YCoef &) { return 0; }
};

} // namespace fixca
