// Micro-benchmark: cycles per step of XXH64's accumulator recurrence b' = rotl(b, 31) * P1 + x in four formulations,
// one warp, four active lanes (what a big input gets).  nvcc -arch=sm_100a -O3 xxchain.cu -o xxchain && ./xxchain
#include <cstdio>
#include <stdint.h>
typedef uint32_t u32;
typedef uint64_t u64;
#define XXP1 0x9E3779B185EBCA87ULL
#define STEPS 32
template <int V>
__device__ __forceinline__ u64 chain32(u64 b, const u64* x) {
	if (V == 0) {  // plain 64-bit C
#pragma unroll
		for (int i = 0; i < STEPS; i++) {
			u64 r = (b << 31) | (b >> 33);
			b = r * XXP1 + x[i];
		}
		return b;
	}
	if (V == 1) {  // mad.lo.u64
#pragma unroll
		for (int i = 0; i < STEPS; i++) {
			u64 r = (b << 31) | (b >> 33);
			asm("mad.lo.u64 %0, %1, %2, %3;" : "=l"(b) : "l"(r), "l"(XXP1), "l"(x[i]));
		}
		return b;
	}
	u32 lo = (u32)b, hi = (u32)(b >> 32);
	const u32 pl = (u32)XXP1, ph = (u32)(XXP1 >> 32);
#pragma unroll
	for (int i = 0; i < STEPS; i++) {
		u32 rl = __funnelshift_r(hi, lo, 1), rh = __funnelshift_r(lo, hi, 1);
		u32 cross = rl * ph + rh * pl;
		u64 t;
		if (V == 2) {  // the library's form: wide multiply-add, cross terms added to the high word afterwards
			asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"(rl), "r"(pl), "l"(x[i]));
			lo = (u32)t;
			hi = (u32)(t >> 32) + cross;
		} else {  // cross terms folded into the addend first
			u64 add = x[i] + ((u64)cross << 32);
			asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"(rl), "r"(pl), "l"(add));
			lo = (u32)t;
			hi = (u32)(t >> 32);
		}
	}
	return ((u64)hi << 32) | lo;
}
template <int V>
__global__ void k(u64* io, const u64* xs, int iters, long long* cyc) {
	__shared__ u64 sx[STEPS * 4];
	for (int i = threadIdx.x; i < STEPS * 4; i += 32) sx[i] = xs[i];
	__syncwarp();
	u64 b = io[threadIdx.x];
	long long t0 = clock64();
	if (threadIdx.x < 4) {
		for (int it = 0; it < iters; it++) {
			u64 x[STEPS];
#pragma unroll
			for (int i = 0; i < STEPS; i++) x[i] = sx[4 * i + threadIdx.x];
			__syncwarp(0xf);
			b = chain32<V>(b, x);
		}
	}
	long long t1 = clock64();
	io[threadIdx.x] = b;
	if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
	u64 *io, *xs;
	long long* cyc;
	cudaMalloc(&io, 32 * 8);
	cudaMalloc(&xs, STEPS * 4 * 8);
	cudaMallocManaged(&cyc, 8);
	u64 h[STEPS * 4];
	for (int i = 0; i < STEPS * 4; i++) h[i] = 0x9E3779B97F4A7C15ULL * (i + 1);
	cudaMemcpy(xs, h, sizeof(h), cudaMemcpyHostToDevice);
	cudaMemcpy(io, h, 32 * 8, cudaMemcpyHostToDevice);
	const int iters = 20000;
	u64 ref = 0;
	for (int v = 0; v < 4; v++) {
		cudaMemcpy(io, h, 32 * 8, cudaMemcpyHostToDevice);
		for (int rep = 0; rep < 2; rep++) {
			if (v == 0) k<0><<<1, 32>>>(io, xs, iters, cyc);
			if (v == 1) k<1><<<1, 32>>>(io, xs, iters, cyc);
			if (v == 2) k<2><<<1, 32>>>(io, xs, iters, cyc);
			if (v == 3) k<3><<<1, 32>>>(io, xs, iters, cyc);
			cudaDeviceSynchronize();
		}
		u64 out;
		cudaMemcpy(&out, io, 8, cudaMemcpyDeviceToHost);
		if (v == 0) ref = out;
		printf("variant %d: %.2f cycles per step (incl. the 32 shared-memory loads per 32 steps)  result %s\n", v,
		       (double)*cyc / ((double)iters * STEPS), out == ref ? "same" : "DIFFERENT");
	}
	return 0;
}
