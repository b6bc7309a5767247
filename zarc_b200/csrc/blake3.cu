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
// 8-chunk unit.  A warp per file leaves most lanes idle when files average ten chunks; here every
// lane streams its own 8 KiB unit (an aligned 8-chunk group is a complete subtree of BLAKE3's
// left-full tree, the last partial group is the right spine) and all lanes meet in one converged
// b3_compress per step.  A step is either the next 64-byte block of the lane's current chunk or a
// parent merge of its subtree stack -- the same compression function with different inputs, so
// divergence is confined to the short preparation around it.  Lanes take the next unit from a
// queue as they finish; a file's <= 8 unit nodes are folded by k_blake3_unit_merge.
#define B3_SMALL_CHUNKS 64u
#define B3_UNIT_CHUNKS 8u
#define B3_SMALL_THREADS 128
#define B3_SMALL_DEPTH 4   // subtree stack inside a unit: <= log2(8) + 1 entries

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

// units per file (0 for files left to the warp-per-file kernel, which are listed in `med`)
__global__ void __launch_bounds__(256) k_blake3_unit_count(const u64* __restrict__ len, u64 n, u64* __restrict__ ucount, u32* __restrict__ med,
                                                            u32* __restrict__ counters) {
	u64 f = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= n) return;
	u64 chunks = (len[f] + 1023) >> 10;
	if (chunks > B3_SMALL_CHUNKS) {
		med[atomicAdd(&counters[2], 1u)] = (u32)f;
		ucount[f] = 0;
	} else {
		ucount[f] = chunks == 0 ? 1 : (chunks + B3_UNIT_CHUNKS - 1) / B3_UNIT_CHUNKS;
	}
}

__global__ void __launch_bounds__(B3_SMALL_THREADS)
k_blake3_small(const u8* __restrict__ blob, const u64* __restrict__ off, const u64* __restrict__ len, u64 n,
               const u64* __restrict__ ubase, const u64* __restrict__ utotal, u8* __restrict__ digests, u32* __restrict__ nodes,
               u32* __restrict__ counters) {
	// subtree stack of every thread: [depth][word][thread] keeps the accesses conflict-free
	__shared__ u32 stack[B3_SMALL_DEPTH][8][B3_SMALL_THREADS];
	u32 tid = threadIdx.x;
	u64 total = *utotal;
	const u8* p = blob;   // next block of the current unit
	u64 f = 0, g = 0;
	u32 left = 0;         // bytes of the unit not yet compressed
	u32 chunk = 0, chunk_end = 0, done_in_unit = 0, blk = 0, depth = 0, merges = 0;
	bool active = false, final_merge = false, whole = false, dry = false;
	u32 cv[8];
	b3_set_iv(cv);
	for (;;) {
		// ---- take the next unit (or find the queue dry) ----
		if (!active && !dry) {
			g = atomicAdd(&counters[1], 1u);
			dry = g >= total;
			if (!dry) {
				u64 lo = 0, hi = n - 1;  // the unit's file: last f with ubase[f] <= g
				while (lo < hi) {
					u64 mid = (lo + hi + 1) >> 1;
					if (ubase[mid] <= g) lo = mid;
					else hi = mid - 1;
				}
				f = lo;
				u64 l = len[f];
				u32 u = (u32)(g - ubase[f]);
				u32 nchunks = l == 0 ? 1u : (u32)((l + 1023) >> 10);
				chunk = u * B3_UNIT_CHUNKS;
				chunk_end = zg_min<u32>(chunk + B3_UNIT_CHUNKS, nchunks);
				whole = nchunks <= B3_UNIT_CHUNKS;
				p = blob + off[f] + (u64)chunk * 1024;
				left = (u32)zg_min<u64>(l - (u64)chunk * 1024, (u64)(chunk_end - chunk) * 1024);
				done_in_unit = blk = depth = merges = 0;
				final_merge = false;
				active = true;
				b3_set_iv(cv);
			}
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
			flags = B3_PARENT | ((final_merge && merges == 1 && whole) ? B3_ROOT : 0u);
		} else if (active) {
			u32 in_chunk = zg_min<u32>(left, 1024u - 64u * blk);  // bytes of this chunk still to go
			blen = zg_min<u32>(in_chunk, 64u);
			bool last_blk = in_chunk <= 64;
			b3_load_block(p, blen, m);  // blen == 0 only for the empty file: all zero words, nothing read
			ctr = chunk;
			flags = (blk == 0 ? B3_CHUNK_START : 0u) | (last_blk ? B3_CHUNK_END : 0u) |
			        ((last_blk && whole && chunk_end == 1) ? B3_ROOT : 0u);
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
				done_in_unit++;
				blk = 0;
				if (chunk == chunk_end) {
					// last chunk of the unit: fold the whole stack onto it
					final_merge = true;
					merges = depth;
					if (merges == 0) done = true;
				} else {
					// completed `done_in_unit` chunks: merge while that count is even, then push
					merges = (u32)__ffs((int)done_in_unit) - 1u;
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
			if (whole) b3_store_digest(digests + 32 * f, cv);
			else {
				ZG_UNROLL
				for (int i = 0; i < 8; i++) nodes[8 * g + i] = cv[i];
			}
			active = false;
		}
	}
}

// files of 2..8 units: fold the unit nodes (left-full over units, like chunks), one lane per file
__global__ void __launch_bounds__(256) k_blake3_unit_merge(const u64* __restrict__ ucount, const u64* __restrict__ ubase, u64 n,
                                                            const u32* __restrict__ nodes, u8* __restrict__ digests) {
	u64 f = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= n) return;
	u32 U = (u32)ucount[f];
	if (U < 2) return;
	const u32* nd = nodes + 8 * ubase[f];
	u32 st[3][8];  // U <= 8: at most 3 pending subtrees
	u32 depth = 0;
	u32 cv[8];
	for (u32 i = 0; i < U; i++) {
		ZG_UNROLL
		for (int j = 0; j < 8; j++) cv[j] = nd[8 * i + j];
		bool last = i + 1 == U;
		u32 merges = last ? depth : (u32)__ffs((int)(i + 1)) - 1u;
		for (u32 k = 0; k < merges; k++) {
			u32 o[8];
			depth--;
			b3_parent_cv(st[depth], cv, last && k + 1 == merges, o);
			ZG_UNROLL
			for (int j = 0; j < 8; j++) cv[j] = o[j];
		}
		if (!last) {
			ZG_UNROLL
			for (int j = 0; j < 8; j++) st[depth][j] = cv[j];
			depth++;
		}
	}
	b3_store_digest(digests + 32 * f, cv);
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
	w.ucount.release();
	w.ubase.release();
	w.unodes.release();
	w.tiles.release();
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
	if (w.ucount.reserve(n * 8) || w.ubase.reserve(n * 8 + 8)) return ZG_ERR(ZG_error_memory_allocation);
	// every file contributes <= 8 units; the node array is sized for the worst case
	if (w.unodes.reserve(n * 8 * 32)) return ZG_ERR(ZG_error_memory_allocation);
	zg_prof_begin(ZG_K_BLAKE3, s);
	ZG_LAUNCH(k_blake3_unit_count, (u32)((n + 255) / 256), 256, 0, s, len, n, w.ucount.as<u64>(), w.med.as<u32>(), w.ctr.as<u32>());
	size_t sr = zg_scan_run(s, w.tiles, w.ucount.as<u64>(), n, 0, w.ubase.as<u64>(), w.ubase.as<u64>() + n);
	if (zg_is_error(sr)) return sr;
	u32 grid = (u32)zg_min<u64>((n + B3_SMALL_THREADS - 1) / B3_SMALL_THREADS, (u64)zg_sm_count() * 7);
	ZG_LAUNCH(k_blake3_small, grid, B3_SMALL_THREADS, 0, s, blob, off, len, n, w.ubase.as<u64>(), w.ubase.as<u64>() + n, digests,
	          w.unodes.as<u32>(), w.ctr.as<u32>());
	ZG_LAUNCH(k_blake3_unit_merge, (u32)((n + 255) / 256), 256, 0, s, w.ucount.as<u64>(), w.ubase.as<u64>(), n, w.unodes.as<u32>(), digests);
	zg_prof_end(ZG_K_BLAKE3, s);
	g_zg_launches += 3;
	zg_publish(s, w.ctr.p, hcount, 12);
	if (cudaStreamSynchronize(s) != cudaSuccess) return ZG_ERR(ZG_error_device);
	u32 nmed = hcount[2];
	if (nmed == 0) return cudaGetLastError() == cudaSuccess ? 0 : ZG_ERR(ZG_error_device);
	// files of more than 64 chunks: one warp per file (those above 1024 chunks are listed for the big path)
	grid = (u32)zg_min<u64>(((u64)nmed + B3_WARPS - 1) / B3_WARPS, (u64)zg_sm_count() * 8);
	ZG_LAUNCH(k_blake3_files, grid, B3_WARPS * 32, 0, s, blob, off, len, w.med.as<u32>(), (u64)nmed, digests, w.big.as<u64>(),
	          w.ctr.as<u32>());
	ZG_COUNT_LAUNCH();
	zg_publish(s, w.ctr.p, hcount, 4);
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
