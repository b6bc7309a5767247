// K1: batched BLAKE3 (one digest per file).  Replaces blake3::hash at
// crates/zarc/src/encode/content_frame.rs:26 and the Hasher at decode/frame_iterator.rs:99,77.
//
// Small/medium files (<= 1024 chunks): one warp per file, one chunk per lane, shuffle tree merge.
// Big files: (1) one warp per group of 32 chunks -> level-5 node, all groups of all big files in
// one launch; (2) one warp per big file folds its level-5 nodes.  HBM-wise: N bytes read once,
// 32 B written per file (+ N/1024 bytes of nodes for big files).
#include "common.h"
#include "blake3.cuh"

#define B3_WARPS 8
#define B3_BIG_CHUNKS 1024ull

ZG_DEV void b3_store_digest(u8* out, const u32 cv[8]) {
	// digests are 32-byte records in a u8 array: 4-byte aligned by construction of the ABI buffers
	if (((uintptr_t)out & 3) == 0) {
		u32* o = (u32*)out;
		ZG_UNROLL
		for (int i = 0; i < 8; i++) o[i] = cv[i];
	} else {
		ZG_UNROLL
		for (int i = 0; i < 8; i++) {
			out[4 * i] = (u8)cv[i];
			out[4 * i + 1] = (u8)(cv[i] >> 8);
			out[4 * i + 2] = (u8)(cv[i] >> 16);
			out[4 * i + 3] = (u8)(cv[i] >> 24);
		}
	}
}

__global__ void __launch_bounds__(B3_WARPS * 32)
k_blake3_files(const u8* __restrict__ blob, const u64* __restrict__ off, const u64* __restrict__ len, const u32* __restrict__ list,
               u64 n, u8* __restrict__ digests, u64* __restrict__ big, u32* __restrict__ big_count) {
	__shared__ B3Stack stacks[B3_WARPS];
	u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	u64 gw = (u64)blockIdx.x * B3_WARPS + warp, nw = (u64)gridDim.x * B3_WARPS;
	for (u64 i = gw; i < n; i += nw) {
		u64 f = list ? (u64)list[i] : i;
		u64 l = len[f];
		if (((l + 1023) >> 10) > B3_BIG_CHUNKS) {
			if (lane == 0) {
				u32 k = atomicAdd(big_count, 1u);
				big[2 * k] = f;
				big[2 * k + 1] = l;
			}
			continue;
		}
		u32 cv[8];
		b3_warp_hash(blob + off[f], l, &stacks[warp], cv);
		if (lane == 0) b3_store_digest(digests + 32 * f, cv);
	}
}

// Small files (<= B3_SMALL_CHUNKS chunks, i.e. almost every file of a source tree): ONE LANE per
// file.  A warp per file leaves most lanes idle when files average ten chunks; here every lane
// streams its own file and all lanes meet in one converged b3_compress per step.  A step is either
// the next 64-byte block of the lane's current chunk or a parent merge of its subtree stack -- the
// same compression function with different inputs, so divergence is confined to the short
// preparation around it.  Lanes take the next file from a queue as they finish.
#define B3_SMALL_CHUNKS 64u
#define B3_SMALL_THREADS 128
#define B3_SMALL_DEPTH 7   // subtree stack: <= log2(64) + 1 entries

// message words of a block of `len` (1..64) bytes at an arbitrarily aligned address, zero padded;
// only aligned words holding at least one message byte are read
ZG_DEV void b3_load_block(const u8* p, u32 len, u32 m[16]) {
	uintptr_t a = (uintptr_t)p;
	const u32* q = (const u32*)(a & ~(uintptr_t)3);
	u32 mis = (u32)(a & 3), sh = mis * 8;
	u32 prev = q[0];
	ZG_UNROLL
	for (int i = 0; i < 16; i++) {
		u32 nx = (u32)(4 * (i + 1)) < len + mis ? q[i + 1] : 0u;
		u32 w = __funnelshift_r(prev, nx, sh);
		u32 pos = 4u * i;
		m[i] = pos + 4 <= len ? w : (pos < len ? (w & ((1u << (8 * (len - pos))) - 1u)) : 0u);
		prev = nx;
	}
}

__global__ void __launch_bounds__(B3_SMALL_THREADS)
k_blake3_small(const u8* __restrict__ blob, const u64* __restrict__ off, const u64* __restrict__ len, u64 n,
               u8* __restrict__ digests, u32* __restrict__ med, u32* __restrict__ counters) {
	// subtree stack of every thread: [depth][word][thread] keeps the accesses conflict-free
	__shared__ u32 stack[B3_SMALL_DEPTH][8][B3_SMALL_THREADS];
	u32 tid = threadIdx.x;
	const u8* p = blob;   // next block of the current file
	u64 f = 0;
	u32 left = 0;         // bytes of the file not yet compressed
	u32 nchunks = 0, chunk = 0, blk = 0, depth = 0, merges = 0;
	bool active = false, final_merge = false;
	u32 cv[8];
	b3_set_iv(cv);
	for (;;) {
		// ---- take files until this lane has one to hash (or the queue is dry) ----
		while (!active) {
			u64 g = atomicAdd(&counters[1], 1u);
			if (g >= n) break;
			u64 l = len[g];
			if (((l + 1023) >> 10) > B3_SMALL_CHUNKS) {
				med[atomicAdd(&counters[2], 1u)] = (u32)g;  // left to the warp-per-file kernel
				continue;
			}
			f = g;
			p = blob + off[g];
			left = (u32)l;
			nchunks = l == 0 ? 1u : (u32)((l + 1023) >> 10);
			chunk = blk = depth = merges = 0;
			final_merge = false;
			active = true;
			b3_set_iv(cv);
		}
		if (!__any_sync(ZG_FULL, active)) break;
		// ---- prepare this step's compression ----
		u32 m[16];
		u32 ctr = 0, blen = 64, flags = 0;
		if (active && merges) {
			depth--;
			ZG_UNROLL
			for (int i = 0; i < 8; i++) {
				m[i] = stack[depth][i][tid];
				m[8 + i] = cv[i];
			}
			b3_set_iv(cv);
			flags = B3_PARENT | ((final_merge && merges == 1) ? B3_ROOT : 0u);
		} else if (active) {
			u32 in_chunk = zg_min<u32>(left, 1024u - 64u * blk);  // bytes of this chunk still to go
			blen = zg_min<u32>(in_chunk, 64u);
			bool last_blk = in_chunk <= 64;
			if (blen == 64) b3_load_block(p, 64, m);
			else b3_load_block(p, blen, m);  // blen == 0 only for the empty file: all zero words, nothing read
			ctr = chunk;
			flags = (blk == 0 ? B3_CHUNK_START : 0u) | (last_blk ? B3_CHUNK_END : 0u) | ((last_blk && nchunks == 1) ? B3_ROOT : 0u);
		} else {
			ZG_UNROLL
			for (int i = 0; i < 16; i++) m[i] = 0;
		}
		b3_compress(cv, m, ctr, 0, blen, flags);
		// ---- advance ----
		if (!active) continue;
		bool done = false;
		if (merges) {
			merges--;
			if (merges == 0) {
				if (final_merge) done = true;
				else {
					ZG_UNROLL
					for (int i = 0; i < 8; i++) stack[depth][i][tid] = cv[i];
					depth++;
					b3_set_iv(cv);
				}
			}
		} else {
			u32 in_chunk = zg_min<u32>(left, 1024u - 64u * blk);
			bool last_blk = in_chunk <= 64;
			p += blen;
			left -= blen;
			blk++;
			if (last_blk) {
				chunk++;
				blk = 0;
				if (chunk == nchunks) {
					// last chunk: fold the whole stack onto it, the last merge is the root
					final_merge = true;
					merges = depth;
					if (merges == 0) done = true;
				} else {
					// completed `chunk` chunks: merge while that count is even, then push
					merges = (u32)__ffs((int)chunk) - 1u;
					if (merges == 0) {
						ZG_UNROLL
						for (int i = 0; i < 8; i++) stack[depth][i][tid] = cv[i];
						depth++;
						b3_set_iv(cv);
					}
				}
			}
		}
		if (done) {
			b3_store_digest(digests + 32 * f, cv);
			active = false;
		}
	}
}

// big files, pass 1: group g of 32 chunks -> nodes[g]
__global__ void __launch_bounds__(B3_WARPS * 32)
k_blake3_big_groups(const u8* __restrict__ blob, const u64* __restrict__ off, const u64* __restrict__ big,
                    const u64* __restrict__ base, u32 nbig, u64 ngroups, u32* __restrict__ nodes) {
	u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	u64 gw = (u64)blockIdx.x * B3_WARPS + warp, nw = (u64)gridDim.x * B3_WARPS;
	for (u64 g = gw; g < ngroups; g += nw) {
		// binary search: last i with base[i] <= g
		u32 lo = 0, hi = nbig - 1;
		while (lo < hi) {
			u32 mid = (lo + hi + 1) >> 1;
			if (base[mid] <= g) lo = mid;
			else hi = mid - 1;
		}
		u64 f = big[2 * lo], n = big[2 * lo + 1];
		u64 nchunks = (n + 1023) >> 10;
		u64 c0 = (g - base[lo]) * 32;
		u64 c = c0 + lane;
		u32 cnt = (u32)zg_min<u64>((u64)32, nchunks - c0);
		u32 cv[8];
		if (c < nchunks) {
			u64 o = c << 10;
			b3_chunk_cv(blob + off[f] + o, (u32)zg_min<u64>((u64)1024, n - o), c, false, cv);
		}
		b3_warp_reduce(cv, cnt, false);
		if (lane == 0) {
			ZG_UNROLL
			for (int i = 0; i < 8; i++) nodes[8 * g + i] = cv[i];
		}
	}
}

// big files, pass 2: fold the level-5 nodes of one file per warp
__global__ void __launch_bounds__(B3_WARPS * 32)
k_blake3_big_finish(const u64* __restrict__ big, const u64* __restrict__ base, u32 nbig, const u32* __restrict__ nodes,
                    u8* __restrict__ digests) {
	__shared__ B3Stack stacks[B3_WARPS];
	u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	u32 gw = blockIdx.x * B3_WARPS + warp, nw = gridDim.x * B3_WARPS;
	for (u32 i = gw; i < nbig; i += nw) {
		B3Stack* st = &stacks[warp];
		u64 f = big[2 * i];
		u64 nn = base[i + 1] - base[i];
		const u32* nd = nodes + 8 * base[i];
		u64 nbatches = (nn + 31) >> 5;
		if (lane == 0) st->depth = 0;
		__syncwarp();
		u32 cv[8];
		for (u64 b = 0; b < nbatches; b++) {
			u64 k = b * 32 + lane;
			u32 cnt = (u32)zg_min<u64>((u64)32, nn - b * 32);
			if (k < nn) {
				ZG_UNROLL
				for (int j = 0; j < 8; j++) cv[j] = nd[8 * k + j];
			}
			b3_warp_reduce(cv, cnt, nbatches == 1);
			if (lane == 0) {
				if (b + 1 < nbatches) b3_stack_push(st, cv, b + 1);
				else b3_stack_fold(st, cv);
			}
			__syncwarp();
		}
		if (lane == 0) b3_store_digest(digests + 32 * f, cv);
	}
}

void zg_b3work_free(ZgB3Work& w) {
	w.big.release();
	w.med.release();
	w.ctr.release();
	w.base.release();
	w.nodes.release();
	w.h.release();
}

size_t zg_blake3_run(cudaStream_t s, ZgB3Work& w, const u8* blob, const u64* off, const u64* len, u64 n, u8* digests) {
	if (n == 0) return 0;
	if (n >= 0xffffffffull) return ZG_ERR(ZG_error_GENERIC);
	if (w.big.reserve(n * 16) || w.med.reserve(n * 4) || w.ctr.reserve(16) || w.h.reserve(16)) return ZG_ERR(ZG_error_memory_allocation);
	cudaMemsetAsync(w.ctr.p, 0, 16, s);  // [0] big files, [1] small-kernel queue, [2] medium files
	u32* hcount = w.h.as<u32>();
	zg_prof_begin(ZG_K_BLAKE3, s);
	u32 grid = (u32)zg_min<u64>((n + B3_SMALL_THREADS - 1) / B3_SMALL_THREADS, (u64)zg_sm_count() * 7);
	ZG_LAUNCH(k_blake3_small, grid, B3_SMALL_THREADS, 0, s, blob, off, len, n, digests, w.med.as<u32>(), w.ctr.as<u32>());
	zg_prof_end(ZG_K_BLAKE3, s);
	ZG_COUNT_LAUNCH();
	cudaMemcpyAsync(hcount, w.ctr.p, 12, cudaMemcpyDeviceToHost, s);
	if (cudaStreamSynchronize(s) != cudaSuccess) return ZG_ERR(ZG_error_device);
	u32 nmed = hcount[2];
	if (nmed == 0) return cudaGetLastError() == cudaSuccess ? 0 : ZG_ERR(ZG_error_device);
	// files of more than 64 chunks: one warp per file (those above 1024 chunks are listed for the big path)
	grid = (u32)zg_min<u64>(((u64)nmed + B3_WARPS - 1) / B3_WARPS, (u64)zg_sm_count() * 8);
	ZG_LAUNCH(k_blake3_files, grid, B3_WARPS * 32, 0, s, blob, off, len, w.med.as<u32>(), (u64)nmed, digests, w.big.as<u64>(),
	          w.ctr.as<u32>());
	ZG_COUNT_LAUNCH();
	cudaMemcpyAsync(hcount, w.ctr.p, 4, cudaMemcpyDeviceToHost, s);
	if (cudaStreamSynchronize(s) != cudaSuccess) return ZG_ERR(ZG_error_device);
	u32 nbig = *hcount;
	if (nbig == 0) return cudaGetLastError() == cudaSuccess ? 0 : ZG_ERR(ZG_error_device);
	// host bookkeeping on sizes only: group prefix per big file
	if (w.h.reserve((size_t)nbig * 16 + (size_t)(nbig + 1) * 8)) return ZG_ERR(ZG_error_memory_allocation);
	u64* hbig = w.h.as<u64>();
	u64* hbase = hbig + 2 * (size_t)nbig;
	cudaMemcpyAsync(hbig, w.big.p, (size_t)nbig * 16, cudaMemcpyDeviceToHost, s);
	if (cudaStreamSynchronize(s) != cudaSuccess) return ZG_ERR(ZG_error_device);
	u64 ngroups = 0;
	for (u32 i = 0; i < nbig; i++) {
		hbase[i] = ngroups;
		ngroups += (((hbig[2 * i + 1] + 1023) >> 10) + 31) >> 5;
	}
	hbase[nbig] = ngroups;
	if (w.base.reserve((size_t)(nbig + 1) * 8) || w.nodes.reserve((size_t)ngroups * 32)) return ZG_ERR(ZG_error_memory_allocation);
	cudaMemcpyAsync(w.base.p, hbase, (size_t)(nbig + 1) * 8, cudaMemcpyHostToDevice, s);
	u32 g1 = (u32)zg_min<u64>((ngroups + B3_WARPS - 1) / B3_WARPS, (u64)zg_sm_count() * 8);
	ZG_LAUNCH(k_blake3_big_groups, g1, B3_WARPS * 32, 0, s, blob, off, w.big.as<u64>(), w.base.as<u64>(), nbig, ngroups, w.nodes.as<u32>());
	ZG_COUNT_LAUNCH();
	u32 g2 = (u32)zg_min<u64>(((u64)nbig + B3_WARPS - 1) / B3_WARPS, (u64)zg_sm_count() * 8);
	ZG_LAUNCH(k_blake3_big_finish, g2, B3_WARPS * 32, 0, s, w.big.as<u64>(), w.base.as<u64>(), nbig, w.nodes.as<u32>(), digests);
	ZG_COUNT_LAUNCH();
	// hbase must outlive the async upload
	if (cudaStreamSynchronize(s) != cudaSuccess) return ZG_ERR(ZG_error_device);
	return cudaGetLastError() == cudaSuccess ? 0 : ZG_ERR(ZG_error_device);
}
