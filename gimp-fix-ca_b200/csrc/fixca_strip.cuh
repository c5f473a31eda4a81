// fixca_strip.cuh -- the FAST (FP32) Linear / Cubic kernel for sm_100a.
//
// Same job as tiled_kernel (fix_ca_region's row loop, fix-ca.c:1122-1320, in
// the separable form of SURVEY.md App. A), rebuilt around what the first ncu
// capture showed (profiles/r01_ncu_tiled_cubic_u16x3_a.md): the pass is bound
// by instruction issue and by shared-memory wavefronts, not by HBM.
//
//   * One CTA per TW x TH output tile.  TMA bulk copies bring in (a) the source
//     window (tile + halo reachable through the affine map, fix-ca.c:801/:813)
//     and (b) the tile's own pixels straight into the staging tile, so green /
//     alpha pass through (fix-ca.c:1094-1098) without a single instruction.
//   * A thread owns P adjacent output columns of ONE channel (red or blue) and
//     walks down the SOURCE rows of the window.  Per source row it loads
//     NS = P + NW - 1 consecutive samples once and forms the P horizontal
//     results with NW = T + 1 weights each: the extra weight absorbs the
//     one-sample drift between neighbouring columns' tap windows (scale != 1),
//     so register indexing stays static and 5 loads replace 8 (Cubic, P = 2).
//     P is chosen per pixel format so that the lane stride in shared memory
//     (P * bytes-per-pixel) is conflict-free.
//   * The vertical pass keeps the last four horizontal rows in a ring of
//     registers whose slots are static in the unrolled row loop.  The
//     per-output-row weights are stored in shared memory by tap POSITION
//     (distance below the output row's newest tap row, clamped edge taps
//     folded in) and the FMA chain runs oldest -> newest, so a sample's
//     arithmetic does not depend on ring phase, tile, segment or band:
//     4 FMAs per Cubic output, 2 per Linear output, no register moves.
//   * Output rows are emitted when their last tap row has been produced; the
//     per-source-row emit counts come from a small table built with the row
//     coefficients.  float -> integer conversion saturates in hardware
//     (cvt.rni.sat), which is clip_d + round in one instruction.
//   * Columns whose tap windows are bent by the clamp-to-edge rules
//     (fix-ca.c:1271-1298) cannot use consecutive samples; their warps take a
//     per-tap path (same ring, same emission), exact in the same sense.
//
// Results are within +-1 LSB of the reference for u8 / u16 and ~2 ulp(1.0) for
// float; FAST mode assumes finite float samples (a
// zero weight times an Inf/NaN neighbour is NaN).
#pragma once

#include <type_traits>

#include "fixca_kernels.cuh"

namespace fixca {


// fp32 tap weights from the fraction t (fix-ca.c:891-892 Linear, :905-907 Cubic, per tap).
template <int INTERP>
__device__ __forceinline__ void tap_weights(float t, float (&w)[4])
{
	if (INTERP == 1) {
		w[0] = 1.0f - t; w[1] = t; w[2] = 0.f; w[3] = 0.f;
	} else {
		w[0] = ((2.0f - t) * t - 1.0f) * t * 0.5f;
		w[1] = ((3.0f * t - 5.0f) * t * t + 2.0f) * 0.5f;
		w[2] = ((4.0f - 3.0f * t) * t + 1.0f) * t * 0.5f;
		w[3] = (t - 1.0f) * t * t * 0.5f;
	}
}

// The same weights evaluated in FP64 (the caller rounds them to FP32 once): the exact-repair kernels' error bound
//  counts one rounding per weight.
template <int INTERP>
__device__ __forceinline__ void tap_weights_d(double t, double (&w)[4])
{
	if (INTERP == 1) {
		w[0] = 1.0 - t; w[1] = t; w[2] = 0.0; w[3] = 0.0;
	} else {
		w[0] = ((2.0 - t) * t - 1.0) * t * 0.5;
		w[1] = ((3.0 * t - 5.0) * t * t + 2.0) * 0.5;
		w[2] = ((4.0 - 3.0 * t) * t + 1.0) * t * 0.5;
		w[3] = (t - 1.0) * t * t * 0.5;
	}
}

// Sample codecs.  The integer <-> float conversions cost no instruction at all:
//   load   ld.shared.u8/u16 zero-extends into a 32-bit register, and that register IS the operand: the bit
//          pattern of the integer v read as a float is the subnormal v * 2^-149, exact, and FMUL / FFMA take
//          subnormal operands at full rate.  The horizontal weights carry 2^100 (kHScale) and the vertical
//          weights 2^49 (inside kInvMax), so every intermediate is the value an I2FP-converted sample would
//          give times a power of two: the same mantissa after every rounding (no intermediate that matters
//          comes near the subnormal range: 2^-49 * 1e-30 is still normal).  (First version: cvt.rn.f32.u32 ->
//          I2FP on the ALU pipe, 8 of the 60 instructions per row of an RGB8 Cubic thread; before that I2F
//          on the XU pipe, profiles/r01_*_d.md.)
//   store  the vertical weights are pre-scaled by 1/max so the last FMA saturates to [0,1] (FFMA.SAT
//          = clip_d, fix-ca.c:873-880); one more FMA with 1.5 * 2^23 rounds max * r to nearest-even in
//          the low mantissa bits, which st.shared.u8/u16 then stores.
constexpr float kIntHScale = 0x1p100f, kIntVScale = 0x1p49f;	// 100 + 49 = 149
template <class S> struct StripCodec;
template <> struct StripCodec<uint8_t> {
	static constexpr float kHScale = kIntHScale;
	static constexpr float kInvMax = (float)(1.0 / 255.0) * kIntVScale;
	__device__ __forceinline__ static float load(const unsigned char *p)
	{
		unsigned v;
		asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)));
		return __uint_as_float(v);
	}
	__device__ __forceinline__ static float load_at(uint32_t saddr)	// 32-bit shared-window address
	{
		unsigned v;
		asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
		return __uint_as_float(v);
	}
	__device__ __forceinline__ static void store(unsigned char *p, float sat01)
	{
		*p = (unsigned char)__float_as_uint(fmaf(sat01, 255.0f, 12582912.0f));
	}
	// exact-repair kernels: store, and report whether the FP32 value lies within kEps of a rounding boundary
	// (kEps bounds |FP32 value - reference value| in LSB for raw samples <= kRawMax)
	static constexpr float kMax = 255.0f, kEps = 0.08f * 255.0f / 65535.0f;	// bound 2.6e-4, DESIGN.md 4.6
	__device__ __forceinline__ static bool store_flag(unsigned char *p, float sat01)
	{
		const float r = fmaf(sat01, kMax, 12582912.0f);
		*p = (unsigned char)__float_as_uint(r);
		return fabsf(fmaf(sat01, kMax, 12582912.0f - r)) >= 0.5f - kEps;
	}
	// ... store, and return the distance of the FP32 value from the integer it rounds to (|.| >= 0.5 - kEps: near a boundary)
	__device__ __forceinline__ static float store_resid(unsigned char *p, float sat01)
	{
		const float r = fmaf(sat01, kMax, 12582912.0f);
		*p = (unsigned char)__float_as_uint(r);
		return fmaf(sat01, kMax, 12582912.0f - r);
	}
};
template <> struct StripCodec<uint16_t> {
	static constexpr float kHScale = kIntHScale;
	static constexpr float kInvMax = (float)(1.0 / 65535.0) * kIntVScale;
	__device__ __forceinline__ static float load(const unsigned char *p)
	{
		unsigned v;
		asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)));
		return __uint_as_float(v);
	}
	__device__ __forceinline__ static float load_at(uint32_t saddr)
	{
		unsigned v;
		asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(saddr));
		return __uint_as_float(v);
	}
	__device__ __forceinline__ static void store(unsigned char *p, float sat01)
	{
		*reinterpret_cast<uint16_t *>(p) = (uint16_t)__float_as_uint(fmaf(sat01, 65535.0f, 12582912.0f));
	}
	static constexpr float kMax = 65535.0f, kEps = 0.08f;
	__device__ __forceinline__ static bool store_flag(unsigned char *p, float sat01)
	{
		const float r = fmaf(sat01, kMax, 12582912.0f);
		*reinterpret_cast<uint16_t *>(p) = (uint16_t)__float_as_uint(r);
		return fabsf(fmaf(sat01, kMax, 12582912.0f - r)) >= 0.5f - kEps;
	}
	__device__ __forceinline__ static float store_resid(unsigned char *p, float sat01)
	{
		const float r = fmaf(sat01, kMax, 12582912.0f);
		*reinterpret_cast<uint16_t *>(p) = (uint16_t)__float_as_uint(r);
		return fmaf(sat01, kMax, 12582912.0f - r);
	}
};
template <> struct StripCodec<u15_t> {	// bpc = 15: the loads of u16, max = 32768
	static constexpr float kHScale = kIntHScale;
	static constexpr float kInvMax = (float)(1.0 / 32768.0) * kIntVScale;
	__device__ __forceinline__ static float load(const unsigned char *p) { return StripCodec<uint16_t>::load(p); }
	__device__ __forceinline__ static float load_at(uint32_t saddr) { return StripCodec<uint16_t>::load_at(saddr); }
	__device__ __forceinline__ static void store(unsigned char *p, float sat01)
	{
		*reinterpret_cast<uint16_t *>(p) = (uint16_t)__float_as_uint(fmaf(sat01, 32768.0f, 12582912.0f));
	}
	static constexpr float kMax = 32768.0f, kEps = 0.08f;	// raw codes run to 65535 (out-of-range inputs): the bound of u16
	__device__ __forceinline__ static bool store_flag(unsigned char *p, float sat01)
	{
		const float r = fmaf(sat01, kMax, 12582912.0f);
		*reinterpret_cast<uint16_t *>(p) = (uint16_t)__float_as_uint(r);
		return fabsf(fmaf(sat01, kMax, 12582912.0f - r)) >= 0.5f - kEps;
	}
	__device__ __forceinline__ static float store_resid(unsigned char *p, float sat01)
	{
		const float r = fmaf(sat01, kMax, 12582912.0f);
		*reinterpret_cast<uint16_t *>(p) = (uint16_t)__float_as_uint(r);
		return fmaf(sat01, kMax, 12582912.0f - r);
	}
};
template <> struct StripCodec<float> {
	static constexpr float kHScale = 1.0f;
	static constexpr float kInvMax = 1.0f;
	__device__ __forceinline__ static float load(const unsigned char *p) { return *reinterpret_cast<const float *>(p); }
	__device__ __forceinline__ static float load_at(uint32_t saddr)
	{
		float f;
		asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f) : "r"(saddr));
		return f;
	}
	__device__ __forceinline__ static void store(unsigned char *p, float sat01) { *reinterpret_cast<float *>(p) = sat01; }
	__device__ __forceinline__ static bool store_flag(unsigned char *p, float sat01) { store(p, sat01); return false; }
	static constexpr float kEps = 0.f;
	__device__ __forceinline__ static float store_resid(unsigned char *p, float sat01) { store(p, sat01); return 0.f; }
};

template <> struct StripCodec<__half> {	// bpc = -2: computed like float images, stored with one rounding to half
	static constexpr float kHScale = 1.0f;
	static constexpr float kInvMax = 1.0f;
	__device__ __forceinline__ static float load(const unsigned char *p) { return __half2float(*reinterpret_cast<const __half *>(p)); }
	__device__ __forceinline__ static float load_at(uint32_t saddr)
	{
		unsigned short v;
		asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(saddr));
		return __half2float(__ushort_as_half(v));
	}
	__device__ __forceinline__ static void store(unsigned char *p, float sat01) { *reinterpret_cast<__half *>(p) = __float2half_rn(sat01); }
	__device__ __forceinline__ static bool store_flag(unsigned char *p, float sat01) { store(p, sat01); return false; }
	static constexpr float kEps = 0.f;
	__device__ __forceinline__ static float store_resid(unsigned char *p, float sat01) { store(p, sat01); return 0.f; }
};

// Samples whose bit patterns include NaN / Inf (a zero weight does not silence them)
template <class S> struct is_float_sample { static constexpr bool value = false; };
template <> struct is_float_sample<float> { static constexpr bool value = true; };
template <> struct is_float_sample<__half> { static constexpr bool value = true; };

// Vertical weights of output row y of one channel, ordered by tap position: .w weighs the newest
// tap row `last` (the row whose arrival completes the output), .z row last - 1, .y last - 2, .x
// last - 3.  Clamp-to-edge taps (fix-ca.c:1219-1256, :1149-1158) land on the same row and their
// weights add up; positions that hold no tap get weight 0.  Pre-scaled by inv_max so that the
// vertical pass lands in [0,1] units.
// PRECISE (exact-repair kernels): tap weights and the merge of clamped taps in FP64, one rounding to FP32 at the end.
template <int INTERP, bool PRECISE = false>
__device__ __forceinline__ float4 position_weights(const Axis &ay, int y, int H, float inv_max, int &last)
{
	constexpr int T = INTERP == 1 ? 2 : 4;
	constexpr int OFF = INTERP == 1 ? 0 : 1;
	double td;
	const int i0 = base_index(ay, y, td);
	last = clampi(i0 - OFF + T - 1, 0, H - 1);
	if (PRECISE) {
		double wd[4], posd[4] = {0.0, 0.0, 0.0, 0.0};
		tap_weights_d<INTERP>(td, wd);
#pragma unroll
		for (int j = 0; j < T; ++j) {
			const int p = last - clampi(i0 - OFF + j, 0, H - 1);
#pragma unroll
			for (int m = 0; m < 4; ++m)
				posd[m] += (p == m) ? wd[j] * (double)inv_max : 0.0;
		}
		return make_float4((float)posd[3], (float)posd[2], (float)posd[1], (float)posd[0]);
	}
	float w[4];
	tap_weights<INTERP>((float)td, w);
	float pos[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
	for (int j = 0; j < T; ++j) {
		const int p = last - clampi(i0 - OFF + j, 0, H - 1);
#pragma unroll
		for (int m = 0; m < 4; ++m)
			pos[m] += (p == m) ? w[j] * inv_max : 0.f;
	}
	return make_float4(pos[3], pos[2], pos[1], pos[0]);
}

// One output row of P columns from the ring of horizontal rows; U = ring slot of the newest row.
// Chain order oldest -> newest; the last FMA saturates (clip_d, fix-ca.c:873-880).  Linear has taps
// at the two newest positions only.
// REPAIR: returns a mask of the columns whose FP32 value lies within Codec::kEps of a rounding boundary.
// REPAIR = 2: the same mask, through one test per row: the largest residual of the P columns against the bound, and the
// per-column tests only behind it (one thread-row in 400 has a near-tie sample: 5 instead of 3 * P test instructions).
template <int INTERP, int U, int P, int BPP, class Codec, int REPAIR = 0>
__device__ __forceinline__ unsigned vertical_emit(const float (&hr)[4][P], const float4 w, unsigned char *q)
{
	unsigned flags = 0;
	[[maybe_unused]] float resid[P];
#pragma unroll
	for (int k = 0; k < P; ++k) {
		float v;
		if (INTERP == 1) {
			v = w.z * hr[(U + 3) & 3][k];
		} else {
			v = w.x * hr[(U + 1) & 3][k];
			v = fmaf(w.y, hr[(U + 2) & 3][k], v);
			v = fmaf(w.z, hr[(U + 3) & 3][k], v);
		}
		v = __saturatef(fmaf(w.w, hr[U][k], v));
		if (REPAIR == 2)
			resid[k] = fabsf(Codec::store_resid(q + k * BPP, v));
		else if (REPAIR)
			flags |= Codec::store_flag(q + k * BPP, v) ? 1u << k : 0u;
		else
			Codec::store(q + k * BPP, v);
	}
	if constexpr (REPAIR == 2) {
		float worst = resid[0];
#pragma unroll
		for (int k = 1; k < P; ++k)
			worst = fmaxf(worst, resid[k]);
		if (__builtin_expect(worst >= 0.5f - Codec::kEps, 0)) {
#pragma unroll
			for (int k = 0; k < P; ++k)
				flags |= resid[k] >= 0.5f - Codec::kEps ? 1u << k : 0u;
		}
	}
	return flags;
}

// ---------------------------------------------------------------------------------------------------------
// WIDE: the FP64 primary arithmetic of the exact-repair stream kernels for 16-bit samples.
// Bit-identical output needs the reference's operation order (64 FP64 operations per Cubic sample) only where
// the result is in doubt: the same separable sum evaluated in FP64 on raw sample values -- 12 FP64 operations
// per sample -- is within kEps LSB of the reference's own FP64 value, so every sample further than kEps from a
// rounding boundary rounds to the reference's integer; the others (2 kEps of all samples) are recomputed with
// interp_sample<ExactF64>.  Raw integer samples enter the DFMAs as subnormal doubles (low word = the sample,
// high word 0: v * 2^-1074, exact; no conversion instruction); the horizontal and the vertical weights carry
// 2^537 each, so the vertical sum lands in LSB units.
// ---------------------------------------------------------------------------------------------------------
struct alignas(16) dvec4 { double x, y, z, w; };
constexpr double kWideScale = 0x1p537;		// 537 + 537 = 1074
template <class S> struct WideCodec;
template <> struct WideCodec<uint16_t> {
	// |separable FP64 value - reference FP64 value| <= 3.5e-8 LSB for raw samples <= 65535
	static constexpr double kEps = 1e-6;
	static constexpr int kMaxInt = 65535;
	__device__ __forceinline__ static double load_at(uint32_t saddr)
	{
		unsigned v;
		asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(saddr));
		return __hiloint2double(0, (int)v);
	}
};
template <> struct WideCodec<u15_t> {	// raw codes run to 65535 (out-of-range inputs): the bound of u16; clip at 1.0 = 32768
	static constexpr double kEps = 1e-6;
	static constexpr int kMaxInt = 32768;
	__device__ __forceinline__ static double load_at(uint32_t saddr) { return WideCodec<uint16_t>::load_at(saddr); }
};

// position_weights in FP64 (no rounding to FP32), scaled by `scale`
template <int INTERP>
__device__ __forceinline__ dvec4 position_weights_wide(const Axis &ay, int y, int H, double scale, int &last)
{
	constexpr int T = INTERP == 1 ? 2 : 4;
	constexpr int OFF = INTERP == 1 ? 0 : 1;
	double td;
	const int i0 = base_index(ay, y, td);
	last = clampi(i0 - OFF + T - 1, 0, H - 1);
	double wd[4], posd[4] = {0.0, 0.0, 0.0, 0.0};
	tap_weights_d<INTERP>(td, wd);
#pragma unroll
	for (int j = 0; j < T; ++j) {
		const int p = last - clampi(i0 - OFF + j, 0, H - 1);
#pragma unroll
		for (int m = 0; m < 4; ++m)
			posd[m] += (p == m) ? wd[j] * scale : 0.0;
	}
	dvec4 r;
	r.x = posd[3]; r.y = posd[2]; r.z = posd[1]; r.w = posd[0];
	return r;
}

// vertical_emit in FP64: the sum in LSB units -> clip to [0, max] (clip_d) and round (set_pixel) through the
// 1.5 * 2^52 addition; returns a mask of the columns whose value lies within kEps of a rounding boundary (those are
// recomputed in the reference's own operation order by the caller; an exact tie always is)
template <int INTERP, int U, int P, int BPP, class WC>
__device__ __forceinline__ unsigned vertical_emit_wide(const double (&hr)[4][P], const dvec4 &w, unsigned char *q)
{
	// Rounding and the boundary test in ONE FP64 addition: t = v + 0.5 + 1.5 * 2^22 lies in [2^22, 2^23) for every
	// |v| < 2^21, so its mantissa is the fixed-point number (v + 0.5 + 2^21) * 2^30, rounded once (<= half a unit = 2^-31 LSB):
	// bits 30+ are floor(v + 0.5) + 2^21 -- taken with one funnel shift, the exponent's bits that come along are a constant
	// that goes into the clamp's bounds, and neither it nor 2^21 reaches the 16 bits stored -- and the low 30 bits are the
	// fraction of v + 0.5, which is near 0 or 1 exactly when v is near a rounding boundary.  (r02 (C): v + 1.5 * 2^52, back
	// again, the difference, a DSETP -- three FP64 additions and a compare per sample on the pipe that bounds the kernel.)
	constexpr double MAGIC = 6291456.5;			// 1.5 * 2^22 + 0.5
	constexpr int EXPO = (int)((0x415u << 22) & 0xffffffffu);	// (biased exponent of 2^22) << 22, what is left of it in 32 bits
	constexpr int BASE = EXPO + (1 << 21);			// the funnel-shifted word of v + 0.5 = 0
	constexpr unsigned EPS = (unsigned)(WC::kEps * 1073741824.0) + 26;	// kEps in units of 2^-30, + the addition's own half unit, + margin
	static_assert(WC::kMaxInt <= 65536 && (BASE & 0xffff) == 0, "the stored 16 bits are the integer's own");
	unsigned flags = 0;
#pragma unroll
	for (int k = 0; k < P; ++k) {
		double v;
		if (INTERP == 1) {
			v = w.z * hr[(U + 3) & 3][k];
		} else {
			v = w.x * hr[(U + 1) & 3][k];
			v = fma(w.y, hr[(U + 2) & 3][k], v);
			v = fma(w.z, hr[(U + 3) & 3][k], v);
		}
		v = fma(w.w, hr[U][k], v);
		const double t = v + MAGIC;			// |v| <= 1.6 * 65535
		const unsigned lo = (unsigned)__double2loint(t);
		const int word = (int)__funnelshift_l(lo, (unsigned)__double2hiint(t), 2);	// EXPO + 2^21 + floor(v + 0.5)
		const int n = min(max(word, BASE), BASE + WC::kMaxInt);
		*reinterpret_cast<uint16_t *>(q + k * BPP) = (uint16_t)n;
		// fraction of v + 0.5 within EPS of 0 or 1 (the shift drops the two integer bits of the low word)
		flags |= ((lo + EPS) << 2) <= ((2 * EPS) << 2) ? 1u << k : 0u;
	}
	return flags;
}

// S      sample type (uint8_t, uint16_t, float)
// NCH    3 or 4 samples per pixel
// INTERP 1 Linear, 2 Cubic
// P      adjacent output columns per thread
// TW     tile width in pixels; blockDim.x == 2 * TW / P (red half, blue half)
//
// Dynamic shared memory: [TileHeader | ytab float4[2][th] | nemit u8[2][ne_pitch] | window | staging tile]
template <class S, int NCH, int INTERP, int P, int TW>
__global__ void __launch_bounds__(2 * TW / P) strip_kernel(const __grid_constant__ KernelArgs a)
{
	extern __shared__ __align__(128) unsigned char smem[];
	constexpr int BPP = NCH * (int)sizeof(S);
	constexpr int OUT_PITCH = TW * BPP;
	constexpr int T = INTERP == 1 ? 2 : 4;		// taps per axis
	constexpr int OFF = INTERP == 1 ? 0 : 1;	// first tap = base index - OFF
	constexpr int NW = P == 1 ? T : T + 1;		// weights per output
	constexpr int NS = P + NW - 1;			// consecutive samples a thread loads per source row
	constexpr int NT = 2 * TW / P;			// threads per CTA
	constexpr int HALF = TW / P;			// threads per channel
	static_assert(HALF % 32 == 0, "a warp must not straddle the two channels");
	typedef StripCodec<S> Codec;

	TileHeader *hdr = reinterpret_cast<TileHeader *>(smem);
	float4 *ytab = reinterpret_cast<float4 *>(smem + a.off_ytab);		// [2][th]
	unsigned char *nemit = smem + a.off_nemit;				// [2][ne_pitch]
	unsigned char *win = smem + a.off_win;
	unsigned char *stage = smem + a.off_out;

	const int tid = threadIdx.x;
	const int c = tid / HALF;		// 0 red, 1 blue: uniform per warp
	const int lt = tid - c * HALF;
	const int W = a.g.width, H = a.g.height;
	const int x0 = blockIdx.x * TW;
	const int y0 = a.y1 + blockIdx.y * a.th;
	const int xl = min(x0 + TW, W) - 1;
	const int yl = min(y0 + a.th, a.y2) - 1;
	const int nrows_out = yl - y0 + 1;

	// ---- 1. window extent (tap ranges at the tile corners) ----
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
		(isrow ? hdr->row_lo : hdr->col_lo)[tid & 3] = lo;
		(isrow ? hdr->row_hi : hdr->col_hi)[tid & 3] = hi;
	}
	for (int k = tid; k < 2 * a.ne_pitch / 4; k += NT)
		reinterpret_cast<int *>(nemit)[k] = 0;
	__syncthreads();

	int col_lo = x0, col_hi = xl, row_lo = y0, row_hi = yl;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		col_lo = min(col_lo, hdr->col_lo[k]);
		col_hi = max(col_hi, hdr->col_hi[k]);
		row_lo = min(row_lo, hdr->row_lo[k]);
		row_hi = max(row_hi, hdr->row_hi[k]);
	}
	// slack columns for the shared-sample windows of the P-column groups
	col_lo = max(col_lo - P, 0);
	col_hi = min(col_hi + P, W - 1);
	const int wb0 = (col_lo * BPP) & ~15;
	const int wbytes = (((col_hi + 1) * BPP + 15) & ~15) - wb0;
	const int wrows = row_hi - row_lo + 1;
	const int wpitch = a.win_pitch;
	// The walker's ring slot and window row index are tied to absolute source rows (slot = row & 3,
	// window row index = row - row_base); the arithmetic is phase-free (tap-position weights).
	const int row_base = row_lo & ~3;

	// ---- 2. TMA: window rows, and the tile's own pixels into the staging tile ----
	uint64_t *bar = reinterpret_cast<uint64_t *>(&hdr->bar);
	const int tile_bytes = ((xl - x0 + 1) * BPP + 15) & ~15;
	if (tid < 32) {
		if (tid == 0)
			mbar_arrive_expect_tx(bar, (uint32_t)(wrows * wbytes + nrows_out * tile_bytes));
		__syncwarp();
		const unsigned char *g = a.src + (long long)(row_lo - a.src_row0) * a.src_pitch + wb0;
		for (int r = tid; r < wrows; r += 32)
			bulk_load(win + (r + row_lo - row_base) * wpitch, g + (long long)r * a.src_pitch, (uint32_t)wbytes, bar);
		const unsigned char *t = a.src + (long long)(y0 - a.src_row0) * a.src_pitch + (long long)x0 * BPP;
		for (int r = tid; r < nrows_out; r += 32)
			bulk_load(stage + r * OUT_PITCH, t + (long long)r * a.src_pitch, (uint32_t)tile_bytes, bar);
	}

	// rows [row_base, row_lo) are walked but not loaded: zero them (finite values for the zero weights)
	for (int k = tid; k < (row_lo - row_base) * (wbytes >> 4); k += NT) {
		const int r = k / (wbytes >> 4), v = k - r * (wbytes >> 4);
		*reinterpret_cast<int4 *>(win + r * wpitch + v * 16) = make_int4(0, 0, 0, 0);
	}

	// ---- 3. per-row coefficients by tap position + emit counts (overlaps the copies) ----
	for (int k = tid; k < 2 * nrows_out; k += NT) {
		const int ch = k >= nrows_out;
		const int r = k - ch * nrows_out;
		int last;
		ytab[ch * a.th + r] = position_weights<INTERP>(a.g.y[ch], y0 + r, H, Codec::kInvMax, last);
		// emitted once source row `last` (its highest tap row) has been produced
		atomicAdd(reinterpret_cast<unsigned int *>(nemit + ch * a.ne_pitch + ((last - row_base) & ~3)),
			  1u << (8 * (last & 3)));
	}

	// ---- 4. per-thread column state ----
	float wt[P][NW];	// fast path: weights over the NS shared samples.  slow path: [k][j<T] tap weights
	int cidx[P];		// base index of each column (slow path)
	int colbase;		// byte offset of shared sample 0 from the window row start
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
#pragma unroll
			for (int j = 0; j < 4; ++j)
				w[k][j] *= Codec::kHScale;	// exact: a power of two (integer samples are read as subnormals)
			idx0[k] = cidx[k] - OFF - k;		// first tap minus k (unclamped)
			bmin = min(bmin, idx0[k]);
		}
		// regular: every tap of column k sits at shared sample k + j', 0 <= j' < NW, without clamping
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
					const int d = j - (idx0[k] - bmin);	// tap number landing on shared sample k + j
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

	// walk bounds for this channel: first and last source row any output row of the tile taps
	int s_first, s_last;
	{
		double td;
		const int i0a = base_index(a.g.y[c], y0, td);
		const int i0b = base_index(a.g.y[c], yl, td);
		s_first = max(i0a - OFF, 0);
		s_last = min(i0b + T - 1 - OFF, H - 1);
		s_first &= ~3;				// ring phase: slot = row & 3
	}

	__syncthreads();	// ytab / nemit complete
	mbar_wait(bar, 0);	// window + pass-through tile have landed

	// ---- 5. walk down the source rows ----
	{
		float hr[4][P];
#pragma unroll
		for (int u = 0; u < 4; ++u)
#pragma unroll
			for (int k = 0; k < P; ++k)
				hr[u][k] = 0.f;
		const unsigned char *prow = win + (s_first - row_base) * wpitch;
		unsigned char *q = stage + lt * P * BPP + 2 * c * (int)sizeof(S);
		const float4 *wy = ytab + c * a.th;
		const unsigned int *ne = reinterpret_cast<const unsigned int *>(nemit + c * a.ne_pitch + (s_first - row_base));
		const int choff = 2 * c * (int)sizeof(S) - wb0;

		// one instantiation per path so that the choice is made once, not per source row
		auto walk = [&](auto fast_path) {
			constexpr bool FAST = decltype(fast_path)::value;
			// one source row into ring slot u (= row & 3), then the output rows it completes
			auto row = [&](auto slot, const unsigned int ne4) {
				constexpr int u = decltype(slot)::value;
				if (FAST) {
					float smp[NS];
#pragma unroll
					for (int m = 0; m < NS; ++m)
						smp[m] = Codec::load(prow + colbase + m * BPP);
#pragma unroll
					for (int k = 0; k < P; ++k) {
						float v = wt[k][0] * smp[k];
#pragma unroll
						for (int j = 1; j < NW; ++j)
							v = fmaf(wt[k][j], smp[k + j], v);
						hr[u][k] = v;
					}
				} else {
#pragma unroll
					for (int k = 0; k < P; ++k) {
						float v = 0.f;
#pragma unroll
						for (int j = 0; j < T; ++j) {
							const int ix = clampi(cidx[k] - OFF + j, 0, W - 1);
							v = fmaf(wt[k][j], Codec::load(prow + ix * BPP + choff), v);
						}
						hr[u][k] = v;
					}
				}
				prow += wpitch;
				// usually exactly one output row completes per source row (scale ~ 1)
#pragma unroll 1
				for (int n = (ne4 >> (8 * u)) & 0xff; n > 0; --n) {
					vertical_emit<INTERP, u, P, BPP, Codec>(hr, *wy++, q);
					q += OUT_PITCH;
				}
			};
			for (int s = s_first; s <= s_last; s += 4) {
				const unsigned int ne4 = *ne++;
				row(std::integral_constant<int, 0>(), ne4);
				row(std::integral_constant<int, 1>(), ne4);
				row(std::integral_constant<int, 2>(), ne4);
				row(std::integral_constant<int, 3>(), ne4);
			}
		};
		if (regular)
			walk(std::true_type());
		else
			walk(std::false_type());
	}

	// ---- 6. staging tile -> global, one bulk store per row ----
	fence_proxy_async_smem();
	__syncthreads();
	if (tid < 32) {
		unsigned char *g = a.dst + (long long)(y0 - a.dst_row0) * a.dst_pitch + (long long)x0 * BPP;
		for (int r = tid; r < nrows_out; r += 32)
			bulk_store(g + (long long)r * a.dst_pitch, stage + r * OUT_PITCH, (uint32_t)tile_bytes);
		bulk_commit();
		bulk_wait_read_all();
	}
}

} // namespace fixca
