// memprobe.cu -- developer probe (not part of the library): how much HBM bandwidth a 2-D strip
// traversal can reach on B200 compared with a linear copy.  Build: nvcc -arch=sm_100a -O3 -o memprobe memprobe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void copy_linear(const int4 *__restrict__ s, int4 *__restrict__ d, size_t n)
{
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
	for (; i + 3 * st < n; i += 4 * st) {
		int4 a = s[i], b = s[i + st], c = s[i + 2 * st], e = s[i + 3 * st];
		d[i] = a; d[i + st] = b; d[i + 2 * st] = c; d[i + 3 * st] = e;
	}
	for (; i < n; i += st) d[i] = s[i];
}

// each CTA copies a strip of `sw` bytes x `seg` rows; UNR rows in flight per thread group
template <int UNR>
__global__ void copy_strips(const char *__restrict__ s, char *__restrict__ d, size_t pitch, int sw, int seg, int rows)
{
	const int v_per_row = sw / 16;                       // int4 per strip row
	const int x0 = blockIdx.x * sw;
	const int ya = blockIdx.y * seg, yb = min(ya + seg, rows);
	const int rows_per_it = blockDim.x / v_per_row;      // rows handled per pass by the CTA
	const int tr = threadIdx.x / v_per_row, tv = threadIdx.x % v_per_row;
	if (tr >= rows_per_it) return;
	for (int y = ya + tr; y < yb; y += rows_per_it * UNR) {
		int4 r[UNR];
#pragma unroll
		for (int u = 0; u < UNR; ++u) {
			const int yy = y + u * rows_per_it;
			if (yy < yb) r[u] = *reinterpret_cast<const int4 *>(s + (size_t)yy * pitch + x0 + tv * 16);
		}
#pragma unroll
		for (int u = 0; u < UNR; ++u) {
			const int yy = y + u * rows_per_it;
			if (yy < yb) *reinterpret_cast<int4 *>(d + (size_t)yy * pitch + x0 + tv * 16) = r[u];
		}
	}
}

static float time_it(void (*f)(void *), void *ctx)
{
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	for (int i = 0; i < 3; ++i) f(ctx);
	float best = 1e9;
	for (int i = 0; i < 10; ++i) {
		cudaEventRecord(a); f(ctx); cudaEventRecord(b); cudaEventSynchronize(b);
		float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
	}
	return best;
}

struct Ctx { const char *s; char *d; size_t pitch; int rows; int sw, seg, threads, unr; };

static void run_linear(void *p) { Ctx *c = (Ctx *)p; copy_linear<<<148 * 8, 512>>>((const int4 *)c->s, (int4 *)c->d, c->pitch * c->rows / 16); }
static void run_strips(void *p)
{
	Ctx *c = (Ctx *)p;
	dim3 grid((unsigned)(c->pitch / c->sw), (c->rows + c->seg - 1) / c->seg);
	if (c->unr == 4) copy_strips<4><<<grid, c->threads>>>(c->s, c->d, c->pitch, c->sw, c->seg, c->rows);
	else copy_strips<8><<<grid, c->threads>>>(c->s, c->d, c->pitch, c->sw, c->seg, c->rows);
}

int main()
{
	const size_t pitch = 12288 * 6;   // 73728
	const int rows = 8192;
	char *s, *d;
	CK(cudaMalloc(&s, pitch * rows)); CK(cudaMalloc(&d, pitch * rows));
	CK(cudaMemset(s, 1, pitch * rows)); CK(cudaMemset(d, 0, pitch * rows));
	Ctx c{s, d, pitch, rows, 0, 0, 0, 4};
	const double gb = 2.0 * pitch * rows / 1e9;
	float ms = time_it(run_linear, &c);
	printf("linear copy                      %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
	const int sws[] = {768, 1536, 3072, 6144, 12288, 73728};
	for (int sw : sws)
		for (int ctas_per_sm : {2, 4, 8})
			for (int unr : {4, 8}) {
				const int strips = (int)(pitch / sw);
				int segs = 148 * ctas_per_sm / strips; if (segs < 1) segs = 1;
				int seg = (rows + segs - 1) / segs;
				c.sw = sw; c.seg = seg; c.unr = unr;
				c.threads = 512; 
				if (sw / 16 > 512) { c.threads = 1024; if (sw / 16 > 1024) continue; }
				ms = time_it(run_strips, &c);
				printf("strip %5d B  %2d CTA/SM (grid %3d x %3d, seg %4d rows) unr %d  %.3f ms  %.0f GB/s\n",
				       sw, ctas_per_sm, strips, (rows + seg - 1) / seg, seg, unr, ms, gb / ms * 1e3);
			}
	CK(cudaDeviceSynchronize());
	return 0;
}
