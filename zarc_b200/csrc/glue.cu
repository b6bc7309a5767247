// Small per-frame bookkeeping kernels around the codec kernels (unpack side).
#include "common.h"
#include "xxh64.cuh"
#include "zstd_common.cuh"

// after decode: size check against Frame.uncompressed, optional Content_Checksum verification
// (libzstd verifies it by default; the reference never disables it, decode/zstd_iterator.rs:29)
__global__ void __launch_bounds__(128)
k_unpack_finalize(const u8* __restrict__ out, const u64* __restrict__ out_off, const u64* __restrict__ ulen,
                  const u64* __restrict__ produced, const u32* __restrict__ cksums, u32* __restrict__ status, u64 n, int verify) {
	u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	u32 st = status[k];
	if (st == ZS_OK && produced[k] != ulen[k]) st = ZS_E_CORRUPT;
	if (st == ZS_OK && verify && cksums[2 * k + 1] && ulen[k] < XX_WARP_MIN) {
		u64 h = xx_hash(out + out_off[k], ulen[k], 0);
		if ((u32)h != cksums[2 * k]) st = ZS_E_CHECKSUM;
	}
	status[k] = st;
}
// the Content_Checksum of the frames of XX_WARP_MIN bytes and more, one warp per frame (xxh64.cuh); runs after
// k_unpack_finalize, which has already settled the size check
__global__ void __launch_bounds__(128)
k_unpack_checksum_warp(const u8* __restrict__ out, const u64* __restrict__ out_off, const u64* __restrict__ ulen,
                       const u32* __restrict__ cksums, u32* __restrict__ status, u64 n, const u64* __restrict__ xx_done,
                       const u64* __restrict__ xx_acc) {
	__shared__ u64 sb[4][XX_SB_WORDS];
	u64 k = (u64)blockIdx.x * 4 + (threadIdx.x >> 5);
	if (k >= n) return;
	if (ulen[k] < XX_WARP_MIN || status[k] != ZS_OK || !cksums[2 * k + 1]) return;
	// a staged frame may have been hashed in part already, while its later blocks were being decoded
	u64 c0 = xx_done ? xx_done[k] : 0;
	u64 acc = xx_warp_acc0();
	if (c0 && (threadIdx.x & 31) < 4) acc = xx_acc[4 * k + (threadIdx.x & 31)];
	u64 h = xx_hash_warp_from(out + out_off[k], ulen[k], c0, acc, sb[threadIdx.x >> 5]);
	if ((threadIdx.x & 31) == 0 && (u32)h != cksums[2 * k]) status[k] = ZS_E_CHECKSUM;
}

// what BLAKE3 verification may read: only frames that decoded (a frame rejected with dstSize_tooSmall / srcSize_wrong
// has an output span that may lie outside `out`); the others are hashed as empty inputs at offset 0
__global__ void __launch_bounds__(128)
k_verify_spans(const u32* __restrict__ status, const u64* __restrict__ out_off, const u64* __restrict__ ulen, u64 out_cap, u64 n,
               u64* __restrict__ voff, u64* __restrict__ vlen) {
	u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	u64 o = out_off[k], l = ulen[k];
	bool good = status[k] == ZS_OK && o <= out_cap && l <= out_cap - o;
	voff[k] = good ? o : 0;
	vlen[k] = good ? l : 0;
}

// ok[k] = frame decoded and BLAKE3(out_k) == expected  (FrameIterator::verify, frame_iterator.rs:86-88)
__global__ void __launch_bounds__(128)
k_digest_compare(const u8* __restrict__ got, const u8* __restrict__ want, const u32* __restrict__ status, u8* __restrict__ ok, u64 n) {
	u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	u32 diff = 0;
	for (u32 i = 0; i < 32; i++) diff |= (u32)(got[32 * k + i] ^ want[32 * k + i]);  // constant time, integrity.rs:17-22
	ok[k] = (status[k] == ZS_OK && diff == 0) ? 1 : 0;
}

// first[0] = min over failing k of (k << 8 | code); ~0 when all succeeded
__global__ void __launch_bounds__(128) k_first_error(const u32* __restrict__ status, u64 n, unsigned long long* first) {
	u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	u32 st = status[k];
	if (st) atomicMin(first, (unsigned long long)((k << 8) | (st & 0xff)));
}

size_t zg_unpack_finalize_run(cudaStream_t s, const u8* out, const u64* out_off, const u64* ulen, const u64* produced,
                              const u32* cksums, u32* status, u64 n, int verify, const ZgZdWork* zw) {
	if (!n) return 0;
	ZG_LAUNCH(k_unpack_finalize, (u32)((n + 127) / 128), 128, 0, s, out, out_off, ulen, produced, cksums, status, n, verify);
	ZG_COUNT_LAUNCH();
	if (verify) {
		const bool part = zw && zw->st.xx_n == n;  // the decode run just before left partial hashes for its staged frames
		ZG_LAUNCH(k_unpack_checksum_warp, (u32)((n + 3) / 4), 128, 0, s, out, out_off, ulen, cksums, status, n,
		          part ? zw->st.xx_done.as<u64>() : (const u64*)nullptr, part ? zw->st.xx_acc.as<u64>() : (const u64*)nullptr);
		ZG_COUNT_LAUNCH();
	}
	return 0;
}
size_t zg_verify_spans_run(cudaStream_t s, const u32* status, const u64* out_off, const u64* ulen, u64 out_cap, u64 n, u64* voff, u64* vlen) {
	if (!n) return 0;
	ZG_LAUNCH(k_verify_spans, (u32)((n + 127) / 128), 128, 0, s, status, out_off, ulen, out_cap, n, voff, vlen);
	ZG_COUNT_LAUNCH();
	return 0;
}
size_t zg_digest_compare_run(cudaStream_t s, const u8* got, const u8* want, const u32* status, u8* ok, u64 n) {
	if (!n) return 0;
	ZG_LAUNCH(k_digest_compare, (u32)((n + 127) / 128), 128, 0, s, got, want, status, ok, n);
	ZG_COUNT_LAUNCH();
	return 0;
}
size_t zg_first_error_run(cudaStream_t s, const u32* status, u64 n, u64* first) {
	cudaMemsetAsync(first, 0xff, 8, s);
	if (!n) return 0;
	ZG_LAUNCH(k_first_error, (u32)((n + 127) / 128), 128, 0, s, status, n, (unsigned long long*)first);
	ZG_COUNT_LAUNCH();
	return 0;
}
