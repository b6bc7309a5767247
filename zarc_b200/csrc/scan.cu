// K7: exclusive prefix sum of u64 (frame lengths -> archive offsets, replacing the running
// `self.offset += bytes` of crates/zarc/src/encode/content_frame.rs:22,45), plus small
// batch-bookkeeping kernels shared by pack and unpack.
#include "common.h"

#define SCAN_T 256
#define SCAN_PER 4
#define SCAN_TILE (SCAN_T * SCAN_PER)

ZG_DEV u64 zg_warp_incl_scan64(u64 v) {
	u32 lane = zg_lane();
	ZG_UNROLL
	for (int d = 1; d < 32; d <<= 1) {
		u64 t = __shfl_up_sync(ZG_FULL, v, d);
		if (lane >= (u32)d) v += t;
	}
	return v;
}
// block-wide exclusive scan of one value per thread (SCAN_T threads); returns the exclusive prefix,
// *total = block sum
ZG_DEV u64 zg_block_excl_scan64(u64 v, u64* total) {
	__shared__ u64 wsum[SCAN_T / 32];
	u32 lane = zg_lane(), warp = threadIdx.x >> 5;
	u64 inc = zg_warp_incl_scan64(v);
	if (lane == 31) wsum[warp] = inc;
	__syncthreads();
	if (warp == 0) {
		u64 w = lane < SCAN_T / 32 ? wsum[lane] : 0;
		u64 winc = zg_warp_incl_scan64(w);
		if (lane < SCAN_T / 32) wsum[lane] = winc - w;
		if (lane == SCAN_T / 32 - 1) *total = winc;
	}
	__syncthreads();
	u64 r = inc - v + wsum[warp];
	__syncthreads();
	return r;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_tiles(const u64* __restrict__ in, u64 n, u64* __restrict__ tile_sum) {
	__shared__ u64 total;
	u64 base = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * SCAN_PER;
	u64 s = 0;
	for (u32 i = 0; i < SCAN_PER; i++)
		if (base + i < n) s += in[base + i];
	zg_block_excl_scan64(s, &total);
	if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}
__global__ void __launch_bounds__(SCAN_T) k_scan_partials(u64* tile_sum, u64 ntiles, u64 base, u64* total_out) {
	__shared__ u64 total;
	u64 carry = base;
	for (u64 t0 = 0; t0 < ntiles; t0 += SCAN_T) {
		u64 i = t0 + threadIdx.x;
		u64 v = i < ntiles ? tile_sum[i] : 0;
		u64 ex = zg_block_excl_scan64(v, &total);
		if (i < ntiles) tile_sum[i] = carry + ex;
		carry += total;
		__syncthreads();
	}
	if (threadIdx.x == 0 && total_out) *total_out = carry;
}
__global__ void __launch_bounds__(SCAN_T) k_scan_apply(const u64* __restrict__ in, u64 n, const u64* __restrict__ tile_sum, u64* __restrict__ out) {
	__shared__ u64 total;
	u64 base = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * SCAN_PER;
	u64 v[SCAN_PER];
	u64 s = 0;
	for (u32 i = 0; i < SCAN_PER; i++) {
		v[i] = base + i < n ? in[base + i] : 0;
		s += v[i];
	}
	u64 ex = zg_block_excl_scan64(s, &total) + tile_sum[blockIdx.x];
	for (u32 i = 0; i < SCAN_PER; i++) {
		if (base + i < n) out[base + i] = ex;
		ex += v[i];
	}
}

// out[i] = base + sum(in[0..i)); *total_out (device, may be null) = base + sum(in).  in may alias out.
size_t zg_scan_run(cudaStream_t s, ZgBuf& tiles, const u64* in, u64 n, u64 base, u64* out, u64* total_out) {
	u64 ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
	if (ntiles == 0) ntiles = 1;
	if (tiles.reserve(ntiles * 8)) return ZG_ERR(ZG_error_memory_allocation);
	ZG_LAUNCH(k_scan_tiles, (u32)ntiles, SCAN_T, 0, s, in, n, tiles.as<u64>());
	ZG_LAUNCH(k_scan_partials, 1, SCAN_T, 0, s, tiles.as<u64>(), ntiles, base, total_out);
	ZG_LAUNCH(k_scan_apply, (u32)ntiles, SCAN_T, 0, s, in, n, tiles.as<u64>(), out);
	g_zg_launches += 3;
	return cudaGetLastError() == cudaSuccess ? 0 : ZG_ERR(ZG_error_device);
}

// A few bytes of results for the host, without the copy engine: a bulk D2H transfer of another stream (the
// host-buffer API downloads one slice while the next one computes) would keep a small cudaMemcpyAsync
// waiting behind it for milliseconds.  The kernel stores straight into pinned host memory (device-visible
// under unified addressing); the caller synchronises the stream before reading.
__global__ void __launch_bounds__(256) k_publish(const u8* __restrict__ src, volatile u8* host_dst, u32 nbytes) {
	for (u32 i = threadIdx.x; i < nbytes; i += blockDim.x) host_dst[i] = src[i];
#ifndef ZG_EMU
	__threadfence_system();
#endif
}
cudaError_t zg_publish(cudaStream_t s, const void* dev_src, void* pinned_dst, u32 nbytes) {
	ZG_LAUNCH(k_publish, 1, 256, 0, s, (const u8*)dev_src, (volatile u8*)pinned_dst, nbytes);
	ZG_COUNT_LAUNCH();
	return cudaGetLastError();
}
