// micro-benchmark: the steady-state batch of the old blur recurrence (8 steps: 8 LDS.U8 in, 8 LDS.U8 trailing edge, add-min-relu chain,
// umulhi + min, 8 STS.U8) in isolation, one warp per SM sub-partition, with parts switched off by template flags
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr unsigned kRing = 512;
template <int FLAGS, int KSTEP>
__global__ void __launch_bounds__(128) k_loop(int *out, int nBatches, unsigned kM, unsigned divHi, long long *cyc)
{
	extern __shared__ __align__(16) uint8_t s_all[];
	const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint8_t *s_in = s_all + warp*(2*kRing*32 + 64), *s_out = s_in + kRing*32;
	for (unsigned i = lane; i < kRing*32; i += 32) { s_in[i] = uint8_t(i*37 + 11); s_out[i] = uint8_t(i*11 + 3); }
	__syncwarp();
	const unsigned r = lane >> 2, chan = lane & 3;
	const unsigned laneBase = (KSTEP == 32) ? lane : (r*(kRing*4 + 16) + chan);
	int acc = out[threadIdx.x];
	const unsigned span = 2*kM - 1, edge = kM - 1;
	long long t0 = clock64();
	for (int b = 0; b < nBatches; ++b)
	{
		const unsigned tb = 256 + unsigned(b)*8;
		const uint8_t *inp = s_in + laneBase + (tb & (kRing-1))*KSTEP;
		const unsigned subPos = (tb - span) & (kRing-1) & ~7u, outPos = (tb - edge) & (kRing-1) & ~7u;
		int px[8], spx[8]; unsigned pack = 0;
		#pragma unroll
		for (int j = 0; j < 8; ++j) px[j] = (FLAGS & 32) ? int(*reinterpret_cast<const unsigned *>(s_in + ((laneBase & ~3u) + ((tb + j) & (kRing-1))*KSTEP)) >> (8*chan)) & 0xff : int(inp[j*KSTEP]);
		#pragma unroll
		for (int j = 0; j < 8; ++j) spx[j] = (FLAGS & 2) ? (j + b) & 0xff : (FLAGS & 32) ? int(*reinterpret_cast<const unsigned *>(s_out + ((laneBase & ~3u) + (subPos + j)*KSTEP)) >> (8*chan)) & 0xff : int(s_out[laneBase + (subPos + j)*KSTEP]);
		uint8_t *outp = s_out + laneBase + outPos*KSTEP;
		#pragma unroll
		for (int j = 0; j < 8; ++j)
		{
			acc = __viaddmin_s32_relu(acc, px[j] - spx[j], 65535 - spx[j]);
			const unsigned o = (FLAGS & 4) ? unsigned(acc) : min(__umulhi(unsigned(acc), divHi), 255u);
			if (FLAGS & 8) *reinterpret_cast<unsigned *>(s_out + ((laneBase & ~3u) + (outPos + j)*KSTEP)) = o;
			else if (FLAGS & 16) { pack = (pack >> 8) | (o << 24); if ((j & 3) == 3) *reinterpret_cast<unsigned *>(s_out + ((laneBase & ~3u) + (outPos + j)*KSTEP)) = pack; }
			else if (!(FLAGS & 1)) outp[j*KSTEP] = uint8_t(o);
			else if (o == 0x12345) out[0] = 1;
		}
	}
	long long t1 = clock64();
	out[threadIdx.x] = acc; if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
	int *d_out; long long *d_cyc, cyc; const int n = 2048;
	cudaMalloc(&d_out, 4096); cudaMemset(d_out, 0, 4096); cudaMalloc(&d_cyc, 8);
#define RUN(F, K, W) cudaFuncSetAttribute(k_loop<F, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4*2*kRing*32 + 1024); for (int r = 0; r < 2; ++r) { k_loop<F, K><<<1, W*32, 4*2*kRing*32 + 1024>>>(d_out, n, 14, 0x0924u << 16, d_cyc); cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost); } \
	printf("flags=%d kstep=%2d warps=%d  %.1f cycles/step  (%s)\n", F, K, W, double(cyc)/n/8, cudaGetErrorString(cudaGetLastError()));
	RUN(0, 4, 1) RUN(1, 4, 1) RUN(2, 4, 1) RUN(3, 4, 1) RUN(8, 4, 1) RUN(16, 4, 1) RUN(32, 4, 1) RUN(40, 4, 1) RUN(48, 4, 1) RUN(18, 4, 1)
	RUN(0, 32, 1) RUN(8, 32, 1) RUN(16, 32, 1) RUN(32, 32, 1) RUN(40, 32, 1) RUN(48, 32, 1)
	printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
	return 0;
}
