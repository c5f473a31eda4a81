// ubench_issue.cu -- developer probe (not part of the library): issue cost of FFMA vs FFMA2 (fma.rn.f32x2)
// alone and mixed with the LDS.U8 / I2FP / STS.U8 traffic of the narrow-pixel row loop, per SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_issue ubench_issue.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b)
{
	u64 r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
	return r;
}
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c)
{
	u64 r;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
	return r;
}
__device__ __forceinline__ float lo(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }

constexpr int ITERS = 2048;

// MODE 0: 16 scalar FFMA per iteration (8 chains x 2)
// MODE 1: 8 FFMA2 per iteration (same flops)
// MODE 2: 16 FFMA + 8 LDS.U8 + 8 I2FP + 4 STS.U8
// MODE 3: 8 FFMA2 + 8 LDS.U8 + 8 I2FP + 4 STS.U8
// MODE 4: 8 LDS.U8 + 8 I2FP + 4 STS.U8 only
template <int MODE>
__global__ void k(float *out, long long *cyc, float w0, float w1)
{
	__shared__ unsigned char sm[8192];
	for (int i = threadIdx.x; i < 8192; i += blockDim.x)
		sm[i] = (unsigned char)i;
	__syncthreads();
	float a[16];
	u64 p[8];
#pragma unroll
	for (int i = 0; i < 16; ++i)
		a[i] = threadIdx.x * 0.001f + i;
#pragma unroll
	for (int i = 0; i < 8; ++i)
		p[i] = pack(a[2 * i], a[2 * i + 1]);
	const u64 pw0 = pack(w0, w0 * 1.01f), pw1 = pack(w1, w1 * 0.99f);
	const unsigned char *src = sm + threadIdx.x * 12 % 4096;
	unsigned char *dst = sm + 4096 + threadIdx.x * 12 % 4000;
	long long t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < ITERS; ++it) {
		float s[8];
		if (MODE >= 2) {
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				unsigned v;
				asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(src + 3 * i + (it & 7))));
				asm("cvt.rn.f32.u32 %0, %1;" : "=f"(s[i]) : "r"(v));
			}
		} else {
#pragma unroll
			for (int i = 0; i < 8; ++i)
				s[i] = w1;
		}
		if (MODE == 0 || MODE == 2) {
#pragma unroll
			for (int i = 0; i < 16; ++i)
				a[i] = fmaf(a[i], w0, s[i & 7]);
		}
		if (MODE == 1 || MODE == 3) {
#pragma unroll
			for (int i = 0; i < 8; ++i)
				p[i] = ffma2(p[i], pw0, MODE == 3 ? pack(s[i], s[(i + 1) & 7]) : pw1);
		}
		if (MODE == 4) {
#pragma unroll
			for (int i = 0; i < 8; ++i)
				a[i] += s[i];
		}
		if (MODE >= 2) {
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				const float v = (MODE == 3) ? lo(p[i]) : a[i];
				dst[3 * i] = (unsigned char)__float_as_uint(v);
			}
		}
	}
	long long t1 = clock64();
	float acc = 0.f;
#pragma unroll
	for (int i = 0; i < 16; ++i)
		acc += a[i];
#pragma unroll
	for (int i = 0; i < 8; ++i)
		acc += lo(p[i]) + hi(p[i]);
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
	if (threadIdx.x == 0)
		cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char *name, int threads)
{
	float *out;
	long long *cyc, h[148];
	CK(cudaMalloc(&out, 148 * 1024 * sizeof(float)));
	CK(cudaMalloc(&cyc, 148 * sizeof(long long)));
	k<MODE><<<148, threads>>>(out, cyc, 0.5f, 0.25f);
	CK(cudaDeviceSynchronize());
	k<MODE><<<148, threads>>>(out, cyc, 0.5f, 0.25f);
	CK(cudaDeviceSynchronize());
	CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
	double avg = 0;
	for (int i = 0; i < 148; ++i)
		avg += h[i];
	avg /= 148;
	const int warps_per_smsp = threads / 32 / 4;
	printf("%-44s threads %4d  cycles/iter/warp-on-SMSP %.2f  (cycles per iteration of one SMSP's %d warps: %.2f)\n", name, threads,
	       avg / ITERS / warps_per_smsp, warps_per_smsp, avg / ITERS);
	cudaFree(out);
	cudaFree(cyc);
}

int main()
{
	for (int threads : {128, 256, 512, 1024}) {
		run<0>("16 FFMA", threads);
		run<1>("8 FFMA2", threads);
		run<2>("16 FFMA + 8 LDS.U8 + 8 I2FP + 4 STS.U8", threads);
		run<3>("8 FFMA2 + 8 LDS.U8 + 8 I2FP + 4 STS.U8", threads);
		run<4>("8 FADD + 8 LDS.U8 + 8 I2FP + 4 STS.U8", threads);
	}
	return 0;
}
