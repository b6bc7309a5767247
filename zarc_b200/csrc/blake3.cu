// K1: batched BLAKE3 (one digest per file).  Replaces blake3::hash at
// crates/zarc/src/encode/content_frame.rs:26 and the Hasher at decode/frame_iterator.rs:99,77.
//
// One scheme for files of every size.  All 1 KiB chunks of all files form one sequence (file f owns the
// chunks cbase[f] .. cbase[f+1]); a warp takes 32 consecutive chunks of it, ONE LANE PER CHUNK, so every
// lane runs the same sixteen compressions in lockstep whatever the file sizes are.
//   k_blake3_chunks  persistent warps, four per CTA.  The bytes of a chunk reach the lane through shared
//                    memory: each lane has the copy unit fetch its chunk 256 bytes at a time
//                    (cp.async.bulk, 16-byte aligned superset of the piece, completion counted on an
//                    mbarrier per warp and buffer), one piece ahead of the one being compressed, and
//                    reads its message words from its own slot (a byte offset of the file inside the
//                    slot is absorbed by funnel shifts).  No lane ever waits for a global load.
//                    Single-chunk files are finished here (ROOT flag); other chunks leave their
//                    chaining value in cv0[chunk].
//   k_blake3_level   BLAKE3's tree is left-full: at every level nodes pair up from the left and an odd
//                    last node moves up unchanged.  Level k of file f lives at slot (cbase[f] >> k) + f
//                    of a ping-pong array (disjoint for different files without a scan per level); one
//                    thread per slot merges two nodes of level k-1.  log2(largest file in chunks)
//                    launches, halving in size; the merge that leaves one node is the root.
// HBM-wise: N bytes read once, 32 B per chunk written and read again by the first level (6 % of N),
// 32 B per file written.  The chunk kernel is bound by the integer pipes (7 rounds x 8 G per 64 bytes).
#include "common.h"
#include "blake3.cuh"
#include "tma.cuh"

// Staging geometry of k_blake3_chunks<WARPS, PIECE, COPY, VAR>: per warp two buffers of 32 slots of PIECE + 16 bytes
// (the 16-byte alignment slack; 272 and 528 are 16 mod 128, which spreads the lanes' slots over the banks), then the
// two mbarriers of each warp.
ZG_HD size_t b3c_smem_bytes(int warps, int piece) { return (size_t)warps * 2 * 32 * (piece + 16) + (size_t)warps * 2 * 8; }

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

// meta: [0] chunks of all files, [1] end of the data (largest off + len), [2] chunks of the largest file
__global__ void __launch_bounds__(256) k_blake3_count(const u64* __restrict__ off, const u64* __restrict__ len, u64 n, u64* __restrict__ cnt,
                                                       unsigned long long* __restrict__ meta) {
	u64 f = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	u64 end = 0, chunks = 0;
	if (f < n) {
		u64 l = len[f];
		chunks = l == 0 ? 1 : (l + 1023) >> 10;  // the empty input is one empty chunk
		cnt[f] = chunks;
		end = off[f] + l;
	}
	ZG_UNROLL
	for (int d = 16; d > 0; d >>= 1) {
		end = zg_max<u64>(end, __shfl_xor_sync(ZG_FULL, end, d));
		chunks = zg_max<u64>(chunks, __shfl_xor_sync(ZG_FULL, chunks, d));
	}
	if (zg_lane() == 0) {
		atomicMax(&meta[1], (unsigned long long)end);
		atomicMax(&meta[2], (unsigned long long)chunks);
	}
}

// the file holding the first chunk of every group of 32 chunks
__global__ void __launch_bounds__(256) k_blake3_taskfile(const u64* __restrict__ cbase, u64 nfiles, u64 ntasks, u32* __restrict__ taskfile) {
	u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= ntasks) return;
	u64 g = t << 5;
	u64 lo = 0, hi = nfiles - 1;  // last f with cbase[f] <= g
	while (lo < hi) {
		u64 mid = (lo + hi + 1) >> 1;
		if (cbase[mid] <= g) lo = mid;
		else hi = mid - 1;
	}
	taskfile[t] = (u32)lo;
}

struct B3Lane {
	const u8* src;  // first byte of the chunk
	u64 ctr;        // chunk index inside its file
	u64 file;
	u32 len;        // bytes in the chunk (0 only for the empty file)
	u32 kind;       // 0: no chunk (past the end), 1: chunk of a multi-chunk file, 2: whole file
};

ZG_DEV void b3c_lookup(u64 t, u64 total_chunks, const u8* blob, const u64* __restrict__ off, const u64* __restrict__ len, const u64* __restrict__ cbase,
                       const u32* __restrict__ taskfile, u64 nfiles, B3Lane& L) {
	u32 lane = zg_lane();
	u64 g0 = t << 5, g = g0 + lane;
	u64 f0 = taskfile[t];
	// files f0+1 .. f0+32 start at most 32 chunks further on: one bit per file start inside this group
	u64 fj = f0 + 1 + lane;
	u32 bit = 0;
	if (fj < nfiles) {
		u64 rel = cbase[fj] - g0;
		if (rel < 32) bit = 1u << (u32)rel;
	}
	u32 starts = __reduce_or_sync(ZG_FULL, bit);
	u64 f = f0 + (u32)__popc(starts & ((2u << lane) - 1u));
	L.kind = 0;
	L.len = 0;
	L.src = nullptr;
	L.ctr = 0;
	L.file = f;
	if (g < total_chunks) {
		u64 b = cbase[f], nf = cbase[f + 1] - b, fl = len[f];
		u64 c = g - b;
		L.ctr = c;
		L.len = (u32)zg_min<u64>(1024, fl - (c << 10));
		L.src = blob + off[f] + (c << 10);
		L.kind = nf == 1 ? 2u : 1u;
	}
}

// Have piece p of the lane's chunk copied into `slot`.  Whole 16-byte units move from 16-byte aligned addresses: the
// superset [a & ~15, (a+n+15) & ~15) is fetched and the lane reads at slot + (a & 15).  The superset never leaves the
// data: its end is clipped to `end16` (the last 16-byte boundary at or below the end of the data) and the few bytes
// beyond are copied by hand.
// COPY 0: one bulk copy (cp.async.bulk) per lane and piece, completion counted in bytes on the warp's mbarrier; every
//         lane arrives once whether it copies or not.
// COPY 1: sixteen-byte cp.async copies, one group per lane and piece (an empty group when there is nothing to copy).
// COPY 2: the same through L1 (the two halves of a 32-byte sector are fetched by consecutive copies).
template <int PIECE, int COPY>
ZG_DEV void b3c_issue(const B3Lane& L, u32 p, u8* slot, u64* bar, const u8* end16) {
	u32 o = p * (u32)PIECE;
	u32 bytes = 0;
	const u8* a16 = nullptr;
	if (L.kind != 0 && o < L.len) {
		const u8* a = L.src + o;
		u32 n = zg_min<u32>(L.len - o, (u32)PIECE);
		a16 = (const u8*)((uintptr_t)a & ~(uintptr_t)15);
		const u8* e16 = (const u8*)(((uintptr_t)a + n + 15) & ~(uintptr_t)15);
		if (e16 > end16) {
			const u8* from = a > end16 ? a : end16;
			for (const u8* q = from; q < a + n; q++) slot[q - a16] = *q;
			if (COPY == 0) zg_fence_proxy_async();
			e16 = end16 > a16 ? end16 : a16;
		}
		bytes = (u32)(e16 - a16);
	}
	if (COPY == 0) {
		if (bytes) {
			zg_mbar_arrive_tx(bar, bytes);
			zg_bulk_g2s(slot, a16, bytes, bar);
		} else {
			zg_mbar_arrive(bar);
		}
	} else {
		ZG_UNROLL
		for (int k = 0; k < PIECE / 16 + 1; k++)
			if ((u32)(16 * k) < bytes) {
				if (COPY == 2) zg_cp_async16_l1(slot + 16 * k, a16 + 16 * k);
				else zg_cp_async16(slot + 16 * k, a16 + 16 * k);
			}
		zg_cp_async_commit();
	}
}

template <int WARPS, int PIECE, int COPY, int VAR>
__global__ void __launch_bounds__(WARPS * 32)
k_blake3_chunks(const u8* __restrict__ blob, const u64* __restrict__ off, const u64* __restrict__ len, const u64* __restrict__ cbase,
                const u32* __restrict__ taskfile, u64 nfiles, const u64* __restrict__ meta, u32* __restrict__ cv0, u8* __restrict__ digests, u32 one) {
	ZG_DYN_SMEM(u8, smem);
	constexpr u32 SLOT = PIECE + 16, BLOCKS = PIECE / 64;
	u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	u8* buf0 = smem + ((size_t)(warp * 2 + 0) * 32 + lane) * SLOT;  // this lane's slot in either buffer
	u8* buf1 = smem + ((size_t)(warp * 2 + 1) * 32 + lane) * SLOT;
	u64* bar = (u64*)(smem + (size_t)WARPS * 2 * 32 * SLOT) + 2 * warp;
	u64 total = meta[0];
	u64 ntasks = (total + 31) >> 5;
	const u8* end16 = (const u8*)((uintptr_t)(blob + meta[1]) & ~(uintptr_t)15);
	u64 nw = (u64)gridDim.x * WARPS;
	u64 t = (u64)blockIdx.x * WARPS + warp;
	if (COPY == 0) {
		if (lane == 0) {
			zg_mbar_init(&bar[0], 32);
			zg_mbar_init(&bar[1], 32);
			zg_mbar_fence_init();
		}
		__syncwarp();
	}
	if (t >= ntasks) return;
	B3Lane cur, nxt;
	b3c_lookup(t, total, blob, off, len, cbase, taskfile, nfiles, cur);
	u32 stage = 0;
	b3c_issue<PIECE, COPY>(cur, 0, buf0, &bar[0], end16);
	for (;;) {
		u32 maxlen = __reduce_max_sync(ZG_FULL, cur.len);
		u32 npieces = zg_max<u32>(1u, (maxlen + PIECE - 1) / PIECE);
		const u32 mis = (u32)((uintptr_t)cur.src & 15), sh = (mis & 3u) * 8u;
		const bool words = __all_sync(ZG_FULL, sh == 0);  // every lane's chunk starts on a word boundary
		u32 cv[8], fin[8];
		b3_set_iv(cv);
		ZG_UNROLL
		for (int i = 0; i < 8; i++) fin[i] = 0;
		bool have_next = false;
		u64 tn = t + nw;
		for (u32 p = 0; p < npieces; p++, stage++) {
			// the copies run one stage ahead: the next piece of this group, or the first piece of the next group
			u32 nb = (stage + 1) & 1u;
			if (p + 1 < npieces) {
				b3c_issue<PIECE, COPY>(cur, p + 1, nb ? buf1 : buf0, &bar[nb], end16);
			} else if (tn < ntasks) {
				have_next = true;
				b3c_lookup(tn, total, blob, off, len, cbase, taskfile, nfiles, nxt);
				b3c_issue<PIECE, COPY>(nxt, 0, nb ? buf1 : buf0, &bar[nb], end16);
			} else if (COPY != 0) {
				zg_cp_async_commit();  // (keeps "all but the newest group" meaning this stage's copies)
			}
			if (COPY == 0) zg_mbar_wait(&bar[stage & 1u], (stage >> 1) & 1u);
			else zg_cp_async_wait<1>();
			const u8* slot = (stage & 1u) ? buf1 : buf0;
			ZG_UNROLL1
			for (u32 bi = 0; bi < BLOCKS; bi++) {
				u32 o = p * PIECE + 64u * bi;  // offset of this block inside the chunk
				bool active = cur.kind != 0 && (o < cur.len || o == 0);
				if (!__any_sync(ZG_FULL, active)) break;
				u32 rem = active ? cur.len - o : 65u;
				const u32* w = (const u32*)(slot + ((mis + 64u * bi) & ~3u));
				u32 m[16];
				if (words) {
					ZG_UNROLL
					for (int i = 0; i < 16; i++) m[i] = w[i];
				} else {
					u32 prev = w[0];
					ZG_UNROLL
					for (int i = 0; i < 16; i++) {
						u32 nx = w[i + 1];
						m[i] = __funnelshift_r(prev, nx, sh);
						prev = nx;
					}
				}
				bool last = rem <= 64u;
				u32 blen = zg_min<u32>(rem, 64u);
				u32 flags = (o == 0 ? B3_CHUNK_START : 0u) | (last ? B3_CHUNK_END : 0u) | ((last && cur.kind == 2u) ? B3_ROOT : 0u);
				bool ends = __any_sync(ZG_FULL, last);
				if (ends && __any_sync(ZG_FULL, rem < 64u)) {
					// a short last block is padded with zeros: word i keeps its first clamp(blen - 4i, 0, 4) bytes
					ZG_UNROLL
					for (int i = 0; i < 16; i++) {
						i32 keep = zg_min<i32>(zg_max<i32>((i32)blen - 4 * i, 0), 4);
						m[i] &= ~zg_shl_clamp(0xffffffffu, 8u * (u32)keep);
					}
				}
				b3_compress_v<VAR>(cv, m, (u32)cur.ctr, (u32)(cur.ctr >> 32), blen, flags, one);
				if (ends && last) {
					ZG_UNROLL
					for (int i = 0; i < 8; i++) fin[i] = cv[i];
				}
			}
		}
		if (cur.kind == 2u) {
			b3_store_digest(digests + 32 * cur.file, fin);
		} else if (cur.kind == 1u) {
			uint4* o = (uint4*)(cv0 + 8 * ((t << 5) + lane));
			o[0] = make_uint4(fin[0], fin[1], fin[2], fin[3]);
			o[1] = make_uint4(fin[4], fin[5], fin[6], fin[7]);
		}
		if (!have_next) break;
		cur = nxt;
		t = tn;
	}
}

// One level of the tree for all files: slot s of level k (k >= 1) merges nodes 2j and 2j+1 of level k-1 of its file.
// Level 0 is cv0[chunk]; level k >= 1 of file f starts at slot (cbase[f] >> k) + f.
// A block of B3L_THREADS consecutive slots touches at most B3L_THREADS + 1 consecutive files, the first of which
// k_blake3_level_index found beforehand: the slot bases of those files go to shared memory and every thread finds its
// file there.
#define B3L_THREADS 256
#define B3L_MAXLEVELS 56
struct B3Levels {
	u64 tab_off[B3L_MAXLEVELS];  // where level k's block table starts (index k - 1)
	u32 first, count;            // levels first .. first + count - 1
};
ZG_DEV u64 b3l_base(const u64* __restrict__ cbase, u64 f, u32 k) { return (cbase[f] >> k) + f; }

// for every block of every level: the file of its first slot
__global__ void __launch_bounds__(256) k_blake3_level_index(const u64* __restrict__ cbase, u64 nfiles, u64 total_chunks, B3Levels lv,
                                                             u32* __restrict__ tab) {
	u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	u32 k = lv.first;
	for (u32 q = 0; q < lv.count; q++, k++) {
		u64 nblocks = (((total_chunks >> k) + nfiles) + B3L_THREADS - 1) / B3L_THREADS;
		if (i < nblocks) break;
		i -= nblocks;
	}
	if (k >= lv.first + lv.count) return;
	u64 s = i * B3L_THREADS;
	u64 lo = 0, hi = nfiles - 1;  // last f with base_k(f) <= s
	while (lo < hi) {
		u64 mid = (lo + hi + 1) >> 1;
		if (b3l_base(cbase, mid, k) <= s) lo = mid;
		else hi = mid - 1;
	}
	tab[lv.tab_off[k - 1] + i] = (u32)lo;
}

__global__ void __launch_bounds__(B3L_THREADS) k_blake3_level(const u64* __restrict__ cbase, u64 nfiles, u32 k, u64 slots,
                                                               const u32* __restrict__ tab, const u32* __restrict__ in, u32* __restrict__ out,
                                                               u8* __restrict__ digests, u32 one) {
	__shared__ u64 sb[B3L_THREADS + 1];
	u64 f0 = tab[blockIdx.x];
	for (u32 i = threadIdx.x; i <= B3L_THREADS; i += B3L_THREADS) sb[i] = f0 + i < nfiles ? b3l_base(cbase, f0 + i, k) : ~0ull;
	__syncthreads();
	u64 s = (u64)blockIdx.x * B3L_THREADS + threadIdx.x;
	if (s >= slots) return;
	u32 lo = 0, hi = B3L_THREADS;  // last i with sb[i] <= s
	while (lo < hi) {
		u32 mid = (lo + hi + 1) >> 1;
		if (sb[mid] <= s) lo = mid;
		else hi = mid - 1;
	}
	u64 f = f0 + lo, b = cbase[f], n = cbase[f + 1] - b;
	u64 cnt_in = (n + ((1ull << (k - 1)) - 1)) >> (k - 1);
	if (cnt_in <= 1) return;  // the file was finished at a lower level
	u64 cnt_out = (cnt_in + 1) >> 1;
	u64 j = s - sb[lo];
	if (j >= cnt_out) return;
	u64 in_base = k == 1 ? b : (b >> (k - 1)) + f;
	const uint4* src = (const uint4*)(in + 8 * (in_base + 2 * j));
	uint4 a0 = src[0], a1 = src[1];
	u32 o[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
	if (2 * j + 1 < cnt_in) {
		uint4 b0 = src[2], b1 = src[3];
		u32 m[16] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
		b3_set_iv(o);
		b3_compress_v<1>(o, m, 0, 0, 64, B3_PARENT | (cnt_out == 1 ? B3_ROOT : 0u), one);
	}
	if (cnt_out == 1) {
		b3_store_digest(digests + 32 * f, o);
	} else {
		uint4* dst = (uint4*)(out + 8 * (sb[lo] + j));
		dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
		dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
	}
}

// The variants of the chunk kernel that can be selected at run time (a tuning aid; the default is the fastest measured:
// profiles/README.md, "BLAKE3 staging variants").
static int g_zg_b3_variant = 9;
extern "C" void zg_internal_set_b3_variant(int v) { g_zg_b3_variant = v; }

template <int WARPS, int PIECE, int COPY, int VAR>
static size_t b3c_launch_one(cudaStream_t s, int ctas, const u8* blob, const u64* off, const u64* len, const u64* cbase, const u32* taskfile, u64 n,
                             u64 ntasks, const u64* meta, u32* cv0, u8* digests) {
	size_t smem = b3c_smem_bytes(WARPS, PIECE);
	static ZgPerDevice attr_dev;
	bool& attr_set = *attr_dev.slot();
	if (!attr_set) {
		if (cudaFuncSetAttribute(k_blake3_chunks<WARPS, PIECE, COPY, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
			return ZG_ERR(ZG_error_device);
		attr_set = true;
	}
	u32 grid = (u32)zg_min<u64>((ntasks + WARPS - 1) / WARPS, (u64)zg_sm_count() * ctas);
	ZG_LAUNCH((k_blake3_chunks<WARPS, PIECE, COPY, VAR>), grid, WARPS * 32, smem, s, blob, off, len, cbase, taskfile, n, meta, cv0, digests, 1u);
	ZG_COUNT_LAUNCH();
	return 0;
}
static size_t b3c_launch(cudaStream_t s, int variant, const u8* blob, const u64* off, const u64* len, const u64* cbase, const u32* taskfile, u64 n,
                         u64 ntasks, const u64* meta, u32* cv0, u8* digests) {
#define B3C_CASE(id, W, P, C, V, ctas) \
	case id: return b3c_launch_one<W, P, C, V>(s, ctas, blob, off, len, cbase, taskfile, n, ntasks, meta, cv0, digests)
	switch (variant) {
		B3C_CASE(1, 4, 256, 0, 0, 3);  // bulk copies, 256-byte pieces, plain G
		B3C_CASE(2, 2, 512, 0, 1, 3);  // bulk copies, 512-byte pieces, 6 warps per SM
		B3C_CASE(3, 4, 256, 1, 1, 3);  // cp.async, 256-byte pieces
		B3C_CASE(4, 4, 256, 1, 3, 3);  // cp.async, rotations by 12 on the FMA pipe
		B3C_CASE(5, 4, 256, 0, 3, 3);  // bulk copies, rotations by 12 on the FMA pipe
		B3C_CASE(6, 4, 512, 1, 1, 1);  // cp.async, 512-byte pieces, 4 warps per SM
		B3C_CASE(7, 2, 512, 1, 1, 3);  // cp.async, 512-byte pieces, 6 warps per SM
		B3C_CASE(8, 4, 128, 1, 1, 5);  // cp.async, 128-byte pieces, 20 warps per SM
		B3C_CASE(9, 4, 128, 1, 1, 6);  // cp.async, 128-byte pieces, 24 warps per SM
		B3C_CASE(10, 4, 128, 0, 1, 5);  // bulk copies, 128-byte pieces, 20 warps per SM
		B3C_CASE(11, 4, 256, 2, 1, 3);  // cp.async through L1 (.ca), 256-byte pieces
		B3C_CASE(12, 4, 128, 2, 1, 5);  // cp.async through L1 (.ca), 128-byte pieces, 20 warps per SM
	default: B3C_CASE(0, 4, 256, 0, 1, 3);  // bulk copies, 256-byte pieces, additions split over the pipes, 12 warps per SM
	}
#undef B3C_CASE
}

void zg_b3work_free(ZgB3Work& w) {
	w.cnt.release();
	w.cbase.release();
	w.tiles.release();
	w.meta.release();
	w.taskfile.release();
	w.cv0.release();
	w.cv1.release();
	w.ltab.release();
	w.h.release();
}

size_t zg_blake3_run(cudaStream_t s, ZgB3Work& w, const u8* blob, const u64* off, const u64* len, u64 n, u8* digests) {
	if (n == 0) return 0;
	if (n >= 0xffffffffull) return ZG_ERR(ZG_error_GENERIC);
	if (w.cnt.reserve(n * 8) || w.cbase.reserve(n * 8 + 8) || w.meta.reserve(32) || w.h.reserve(32)) return ZG_ERR(ZG_error_memory_allocation);
	u64* meta = w.meta.as<u64>();
	u64* hmeta = w.h.as<u64>();
	cudaMemsetAsync(meta, 0, 32, s);
	zg_prof_begin(ZG_K_BLAKE3, s);
	ZG_LAUNCH(k_blake3_count, (u32)((n + 255) / 256), 256, 0, s, off, len, n, w.cnt.as<u64>(), (unsigned long long*)meta);
	ZG_COUNT_LAUNCH();
	size_t sr = zg_scan_run(s, w.tiles, w.cnt.as<u64>(), n, 0, w.cbase.as<u64>(), meta);
	if (zg_is_error(sr)) return sr;
	// cbase[n] = the total: the scan leaves it in meta[0]; the kernels read cbase[f + 1]
	cudaMemcpyAsync(w.cbase.as<u64>() + n, meta, 8, cudaMemcpyDeviceToDevice, s);
	zg_publish(s, meta, hmeta, 24);
	if (cudaStreamSynchronize(s) != cudaSuccess) return ZG_ERR(ZG_error_device);
	u64 total = hmeta[0], maxchunks = hmeta[2];
	u64 ntasks = (total + 31) >> 5;
	u64 nodes = zg_max<u64>(total, (total >> 1) + n);
	if (w.taskfile.reserve(ntasks * 4) || w.cv0.reserve(nodes * 32) || (maxchunks > 2 && w.cv1.reserve(nodes * 32)))
		return ZG_ERR(ZG_error_memory_allocation);
	ZG_LAUNCH(k_blake3_taskfile, (u32)((ntasks + 255) / 256), 256, 0, s, w.cbase.as<u64>(), n, ntasks, w.taskfile.as<u32>());
	ZG_COUNT_LAUNCH();
	size_t lr = b3c_launch(s, g_zg_b3_variant, blob, off, len, w.cbase.as<u64>(), w.taskfile.as<u32>(), n, ntasks, meta, w.cv0.as<u32>(), digests);
	if (zg_is_error(lr)) return lr;
	u32 nlevels = 0;
	while ((1ull << nlevels) < maxchunks) nlevels++;  // levels 1 .. nlevels
	if (nlevels > B3L_MAXLEVELS) return ZG_ERR(ZG_error_GENERIC);
	if (nlevels) {
		B3Levels lv;
		lv.first = 1;
		lv.count = nlevels;
		u64 nblk = 0;
		for (u32 k = 1; k <= nlevels; k++) {
			lv.tab_off[k - 1] = nblk;
			nblk += (((total >> k) + n) + B3L_THREADS - 1) / B3L_THREADS;
		}
		if (w.ltab.reserve(nblk * 4)) return ZG_ERR(ZG_error_memory_allocation);
		ZG_LAUNCH(k_blake3_level_index, (u32)((nblk + 255) / 256), 256, 0, s, w.cbase.as<u64>(), n, total, lv, w.ltab.as<u32>());
		ZG_COUNT_LAUNCH();
		u32* in = w.cv0.as<u32>();
		u32* out = w.cv1.as<u32>();
		for (u32 k = 1; k <= nlevels; k++) {
			u64 slots = (total >> k) + n;
			ZG_LAUNCH(k_blake3_level, (u32)((slots + B3L_THREADS - 1) / B3L_THREADS), B3L_THREADS, 0, s, w.cbase.as<u64>(), n, k, slots,
			          w.ltab.as<u32>() + lv.tab_off[k - 1], in, out, digests, 1u);
			ZG_COUNT_LAUNCH();
			u32* x = in;
			in = out;
			out = x;
		}
	}
	zg_prof_end(ZG_K_BLAKE3, s);
	return cudaGetLastError() == cudaSuccess ? 0 : ZG_ERR(ZG_error_device);
}
