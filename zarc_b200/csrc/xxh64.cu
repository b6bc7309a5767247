// K4: batched XXH64 (one thread per input; the per-input recurrence is serial by construction).
#include "common.h"
#include "xxh64.cuh"

__global__ void __launch_bounds__(128)
k_xxh64(const u8* __restrict__ blob, const u64* __restrict__ off, const u64* __restrict__ len, u64 n, u64* __restrict__ out) {
	u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	if (len[i] >= XX_WARP_MIN) return;  // k_xxh64_warp's
	out[i] = xx_hash(blob + off[i], len[i], 0);
}
// one warp per input of XX_WARP_MIN bytes and more (the others' warps leave at once)
__global__ void __launch_bounds__(128)
k_xxh64_warp(const u8* __restrict__ blob, const u64* __restrict__ off, const u64* __restrict__ len, u64 n, u64* __restrict__ out) {
	__shared__ u64 sb[4][XX_SB_WORDS];
	u64 i = (u64)blockIdx.x * 4 + (threadIdx.x >> 5);
	if (i >= n || len[i] < XX_WARP_MIN) return;
	u64 h = xx_hash_warp(blob + off[i], len[i], sb[threadIdx.x >> 5]);
	if ((threadIdx.x & 31) == 0) out[i] = h;
}

size_t zg_xxh64_run(cudaStream_t s, const u8* blob, const u64* off, const u64* len, u64 n, u64* hashes) {
	if (n == 0) return 0;
	ZG_LAUNCH(k_xxh64, (u32)((n + 127) / 128), 128, 0, s, blob, off, len, n, hashes);
	ZG_LAUNCH(k_xxh64_warp, (u32)((n + 3) / 4), 128, 0, s, blob, off, len, n, hashes);
	ZG_COUNT_LAUNCH();
	ZG_COUNT_LAUNCH();
	return cudaGetLastError() == cudaSuccess ? 0 : ZG_ERR(ZG_error_device);
}
