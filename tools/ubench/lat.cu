// micro-benchmark: dependent-chain latencies of the instructions the blur recurrence is made of (one warp, one SM)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_viaddmnmx(int *out, int d, int hi, int n, long long *cyc)
{
	int acc = out[threadIdx.x];
	long long t0 = clock64();
	#pragma unroll 16
	for (int i = 0; i < n; ++i) acc = __viaddmin_s32_relu(acc, d, hi);
	long long t1 = clock64();
	out[threadIdx.x] = acc; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_addminmax(int *out, int d, int hi, int n, long long *cyc)
{
	int acc = out[threadIdx.x];
	long long t0 = clock64();
	#pragma unroll 16
	for (int i = 0; i < n; ++i) { asm volatile("add.s32 %0, %0, %1;" : "+r"(acc) : "r"(d)); asm volatile("min.s32 %0, %0, %1;" : "+r"(acc) : "r"(hi)); asm volatile("max.s32 %0, %0, 0;" : "+r"(acc)); }
	long long t1 = clock64();
	out[threadIdx.x] = acc; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_iadd(int *out, int d, int n, long long *cyc)
{
	int acc = out[threadIdx.x];
	long long t0 = clock64();
	#pragma unroll 16
	for (int i = 0; i < n; ++i) asm volatile("add.s32 %0, %0, %1;" : "+r"(acc) : "r"(d));
	long long t1 = clock64();
	out[threadIdx.x] = acc; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_lds_chain(int *out, int n, long long *cyc)
{
	__shared__ uint8_t s[4096];
	for (int i = threadIdx.x; i < 4096; i += 32) s[i] = (uint8_t)((i*7+1) & 0xff);
	__syncwarp();
	int idx = threadIdx.x;
	long long t0 = clock64();
	#pragma unroll 8
	for (int i = 0; i < n; ++i) idx = s[(idx*4 + threadIdx.x) & 4095];
	long long t1 = clock64();
	out[threadIdx.x] = idx; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_sts_lds(int *out, int n, long long *cyc)
{
	__shared__ uint8_t s[4096];
	int v = threadIdx.x;
	long long t0 = clock64();
	#pragma unroll 8
	for (int i = 0; i < n; ++i) { s[threadIdx.x*4 + (i & 3)] = (uint8_t)v; v = s[threadIdx.x*4 + (i & 3)] + 1; }
	long long t1 = clock64();
	out[threadIdx.x] = v; if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
	int *d_out; long long *d_cyc, cyc; const int n = 4096;
	cudaMalloc(&d_out, 128); cudaMemset(d_out, 0, 128); cudaMalloc(&d_cyc, 8);
#define RUN(name, ...) for (int r = 0; r < 2; ++r) { name<<<1, 32>>>(__VA_ARGS__); cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost); } printf("%-14s %.2f cycles/iter\n", #name, double(cyc)/n);
	RUN(k_viaddmnmx, d_out, 3, 60000, n, d_cyc)
	RUN(k_addminmax, d_out, 3, 60000, n, d_cyc)
	RUN(k_iadd, d_out, 3, n, d_cyc)
	RUN(k_lds_chain, d_out, n, d_cyc)
	RUN(k_sts_lds, d_out, n, d_cyc)
	printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
	return 0;
}
