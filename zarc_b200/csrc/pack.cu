// Pack-side bookkeeping kernels around K1/K2/K3/K4: content-addressed dedup, unique/block lists,
// frame sizing and frame assembly.  Together with zg_pack_core (abi.cu) this is
// Encoder::add_data_frame (crates/zarc/src/encode/content_frame.rs:20-60) for a whole batch:
//   digest (:26) -> "already have this digest?" (:30) -> compress (:41) -> offset += bytes (:45)
//   -> Frame{offset, digest, length, uncompressed} (:48-57)
// with the same observable results as calling it file by file in order.
#include "common.h"
#include "zstd_common.cuh"
#include "xxh64.cuh"

#define ZE_RAW 0x80000000u

// ---- dedup: open-addressing table of (global file id + 1), keyed by digest --------------------
ZG_DEV bool pk_digest_eq(const u8* a, const u8* b) {
	const u32* x = (const u32*)a;
	const u32* y = (const u32*)b;
	u32 d = 0;
	ZG_UNROLL
	for (int i = 0; i < 8; i++) d |= x[i] ^ y[i];
	return d == 0;
}
// inserts ids [lo, hi); afterwards each digest's slot holds its smallest id (= first occurrence, :30)
__global__ void __launch_bounds__(128) k_dedup_insert(const u8* __restrict__ g_digest, u64 lo, u64 hi, u32* table, u32 mask) {
	u64 id = lo + (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (id >= hi) return;
	const u8* mine = g_digest + 32 * id;
	u32 slot = *(const u32*)mine & mask;
	for (;;) {
		u32 cur = table[slot];
		if (cur == 0) {
			cur = atomicCAS(&table[slot], 0u, (u32)id + 1);
			if (cur == 0) return;
		}
		if (pk_digest_eq(g_digest + 32 * (u64)(cur - 1), mine)) {
			atomicMin(&table[slot], (u32)id + 1);
			return;
		}
		slot = (slot + 1) & mask;
	}
}
// rep[i] = global id of the first occurrence of file i's content; first[i] = (rep == own id);
// nblk[i] = number of Zstandard blocks to encode for i (0 for duplicates)
__global__ void __launch_bounds__(128)
k_dedup_resolve(const u8* __restrict__ g_digest, u64 base, u64 n, const u32* __restrict__ table, u32 mask,
                const u64* __restrict__ len, const u8* __restrict__ select, u64* __restrict__ rep, u8* __restrict__ first,
                u64* __restrict__ isfirst64, u64* __restrict__ nblk, u64* __restrict__ clen) {
	u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	u64 id = base + i;
	const u8* mine = g_digest + 32 * id;
	u32 slot = *(const u32*)mine & mask;
	u32 cur;
	for (;;) {
		cur = table[slot];
		if (cur != 0 && pk_digest_eq(g_digest + 32 * (u64)(cur - 1), mine)) break;
		slot = (slot + 1) & mask;
	}
	u64 r = cur - 1;
	bool f = r == id && (!select || select[i]);  // (select: the decision was taken over more files than this context has seen)
	rep[i] = r;
	first[i] = f ? 1 : 0;
	isfirst64[i] = f ? 1 : 0;
	u64 l = len[i];
	nblk[i] = f ? (l == 0 ? 1 : (l + ZS_BLOCK_MAX - 1) / ZS_BLOCK_MAX) : 0;
	clen[i] = f ? ((l + 15) & ~(u64)15) : 0;  // room for the encoded blocks of i in the staging blob
}
// unique list (insertion order) and each unique file's first block index
__global__ void __launch_bounds__(128)
k_build_ulist(const u8* __restrict__ first, const u64* __restrict__ uidx, const u64* __restrict__ blk_first, u64 n,
              u32* __restrict__ ulist, u64* __restrict__ blk_base, u64 nuniq, u64 nblocks) {
	u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (i == 0) blk_base[nuniq] = nblocks;
	if (i >= n || !first[i]) return;
	ulist[uidx[i]] = (u32)i;
	blk_base[uidx[i]] = blk_first[i];
}

// ---- frame header shape (what libzstd writes for a known content size) -----
struct PkHdr {
	u32 size;     // bytes before the first block
	u32 fcs_len;
	u32 desc;
	u32 single;
};
// flags: bit 0 = Content_Checksum (ZSTD_c_checksumFlag), bit 1 = write Frame_Content_Size
ZG_DEV PkHdr pk_frame_header(u64 len, u32 flags) {
	PkHdr h;
	u32 checksum = flags & 1;
	if (!(flags & 2)) {  // contentSizeFlag = 0: windowed frame without FCS
		h.single = 0;
		h.fcs_len = 0;
		h.desc = checksum ? 4u : 0u;
		h.size = 4 + 1 + 1;
		return h;
	}
	h.single = len <= ((u64)1 << 27) ? 1 : 0;  // larger frames must carry a window <= 2^27 (App. C)
	u32 flag;
	if (h.single) {
		if (len < 256) flag = 0;
		else if (len < 65536 + 256) flag = 1;
		else flag = 2;
	} else {
		flag = len < ((u64)1 << 32) ? 2 : 3;
	}
	h.fcs_len = flag == 0 ? 1 : flag == 1 ? 2 : flag == 2 ? 4 : 8;
	h.desc = (flag << 6) | (h.single << 5) | (checksum ? 4u : 0u);
	h.size = 4 + 1 + (h.single ? 0 : 1) + h.fcs_len;
	return h;
}

// per block: bytes it occupies in the frame (3-byte header + body)
__global__ void __launch_bounds__(128) k_block_out_sizes(const u32* __restrict__ blk_csize, u64 nblocks, u64* __restrict__ blk_out) {
	u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nblocks) return;
	u32 c = blk_csize[b];
	blk_out[b] = 3 + (u64)(c & ~ZE_RAW);
}
// per unique file: total frame length (Frame.length, content_frame.rs:54)
__global__ void __launch_bounds__(128)
k_frame_sizes(const u32* __restrict__ ulist, const u64* __restrict__ blk_base, const u64* __restrict__ blk_pos,
              const u64* __restrict__ blk_total, const u64* __restrict__ len, u64 nuniq, u32 flags, u64* __restrict__ frame_len_u) {
	u64 u = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (u >= nuniq) return;
	u64 b0 = blk_base[u], b1 = blk_base[u + 1];
	u64 p0 = blk_pos[b0], p1 = b1 < blk_base[nuniq] ? blk_pos[b1] : *blk_total;
	PkHdr h = pk_frame_header(len[ulist[u]], flags);
	frame_len_u[u] = h.size + (p1 - p0) + ((flags & 1) ? 4 : 0);
}

ZG_DEV void pk_warp_copy(u8* dst, const u8* src, u32 n) {
	u32 lane = zg_lane();
	if (n >= 64) {
		// destination-aligned 4-byte stores; a source that is not aligned the same way is read as aligned words too and
		// shifted into place (a block of a few KiB copied byte by byte was 3/4 of this kernel's time)
		u32 head = (u32)((4 - ((uintptr_t)dst & 3)) & 3);
		if (lane < head) dst[lane] = src[lane];
		u32 nw = (n - head) >> 2;
		u32 sa = (u32)((uintptr_t)(src + head) & 3);
		const u32* s = (const u32*)(src + head - sa);
		u32* d = (u32*)(dst + head);
		if (sa == 0) {
			for (u32 i = lane; i < nw; i += 32) d[i] = s[i];
		} else {
			// word i needs bytes sa.. of s[i] and bytes ..sa of s[i + 1]: aligned words that each hold at least one byte of
			// [src, src + n), so they cannot leave the allocation
			for (u32 i = lane; i < nw; i += 32) d[i] = __funnelshift_r(s[i], s[i + 1], 8 * sa);
		}
		for (u32 i = head + (nw << 2) + lane; i < n; i += 32) dst[i] = src[i];
	} else {
		for (u32 i = lane; i < n; i += 32) dst[i] = src[i];
	}
}

// one warp per block: block header + body into the frame's final place; the first block's warp
// also writes the frame header, the last block's warp the XXH64 checksum.
__global__ void __launch_bounds__(128)
k_frame_assemble(const u8* __restrict__ blob, const u8* __restrict__ comp, const u64* __restrict__ file_off,
                 const u64* __restrict__ comp_off, const u64* __restrict__ file_len, const u32* __restrict__ ulist, const u64* __restrict__ blk_base,
                 const u64* __restrict__ blk_pos, const u32* __restrict__ blk_csize, const u64* __restrict__ frame_off_u,
                 const u64* __restrict__ xxh, u64 nuniq, u64 nblocks, u32 flags, u64 archive_base, u8* frames_out) {
	u32 checksum = flags & 1;
	u32 lane = threadIdx.x & 31;
	u64 b = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if (b >= nblocks) return;
	u64 lo = 0, hi = nuniq - 1;
	while (lo < hi) {
		u64 mid = (lo + hi + 1) >> 1;
		if (blk_base[mid] <= b) lo = mid;
		else hi = mid - 1;
	}
	u64 u = lo;
	u32 f = ulist[u];
	u64 j = b - blk_base[u];
	u64 flen = file_len[f];
	PkHdr h = pk_frame_header(flen, flags);
	u8* frame = frames_out + (frame_off_u[u] - archive_base);
	u8* dst = frame + h.size + (blk_pos[b] - blk_pos[blk_base[u]]);
	u32 c = blk_csize[b];
	u32 body = c & ~ZE_RAW;
	bool raw = (c & ZE_RAW) != 0;
	bool last = b + 1 == blk_base[u + 1];
	u64 boff = j * ZS_BLOCK_MAX;
	const u8* src = raw ? blob + file_off[f] + boff : comp + comp_off[f] + boff;
	if (lane == 0) {
		u32 bh = (last ? 1u : 0u) | ((raw ? 0u : 2u) << 1) | (body << 3);
		dst[0] = (u8)bh;
		dst[1] = (u8)(bh >> 8);
		dst[2] = (u8)(bh >> 16);
		if (j == 0) {
			frame[0] = 0x28;
			frame[1] = 0xB5;
			frame[2] = 0x2F;
			frame[3] = 0xFD;
			frame[4] = (u8)h.desc;
			u32 p = 5;
			if (!h.single) frame[p++] = 0x38;  // Window_Descriptor: windowLog 17 (matches never leave a block)
			u64 fcs = h.fcs_len == 2 ? flen - 256 : flen;
			for (u32 i = 0; i < h.fcs_len; i++) frame[p++] = (u8)(fcs >> (8 * i));
		}
		if (last && checksum) {
			u32 ck = (u32)xxh[u];
			u8* t = dst + 3 + body;
			t[0] = (u8)ck;
			t[1] = (u8)(ck >> 8);
			t[2] = (u8)(ck >> 16);
			t[3] = (u8)(ck >> 24);
		}
	}
	pk_warp_copy(dst + 3, src, body);
}

// XXH64 of the unique files only
__global__ void __launch_bounds__(128)
k_xxh64_list(const u8* __restrict__ blob, const u64* __restrict__ off, const u64* __restrict__ len, const u32* __restrict__ ulist,
             u64 nuniq, u64* __restrict__ out) {
	u64 u = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (u >= nuniq) return;
	u32 f = ulist[u];
	if (len[f] >= XX_WARP_MIN) return;  // k_xxh64_list_warp's
	out[u] = xx_hash(blob + off[f], len[f], 0);
}
// ... and one warp per unique file of XX_WARP_MIN bytes and more (xxh64.cuh)
__global__ void __launch_bounds__(128)
k_xxh64_list_warp(const u8* __restrict__ blob, const u64* __restrict__ off, const u64* __restrict__ len, const u32* __restrict__ ulist,
                  u64 nuniq, u64* __restrict__ out) {
	__shared__ u64 sb[4][XX_SB_WORDS];
	u64 u = (u64)blockIdx.x * 4 + (threadIdx.x >> 5);
	if (u >= nuniq) return;
	u32 f = ulist[u];
	if (len[f] < XX_WARP_MIN) return;
	u64 h = xx_hash_warp(blob + off[f], len[f], sb[threadIdx.x >> 5]);
	if ((threadIdx.x & 31) == 0) out[u] = h;
}

// record the new frames in the encoder's map, then answer every file from the map
__global__ void __launch_bounds__(128)
k_record_frames(const u32* __restrict__ ulist, const u64* __restrict__ frame_off_u, const u64* __restrict__ frame_len_u, u64 nuniq,
                u64 base, u64* __restrict__ g_off, u64* __restrict__ g_len) {
	u64 u = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (u >= nuniq) return;
	u64 id = base + ulist[u];
	g_off[id] = frame_off_u[u];
	g_len[id] = frame_len_u[u];
}
__global__ void __launch_bounds__(128)
k_answer_files(const u64* __restrict__ rep, const u64* __restrict__ g_off, const u64* __restrict__ g_len, u64 n,
               u64* __restrict__ frame_off, u64* __restrict__ frame_len) {
	u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	u64 r = rep[i];
	if (frame_off) frame_off[i] = g_off[r];
	if (frame_len) frame_len[i] = g_len[r];
}

#define PK_GRID(n) (u32)(((n) + 127) / 128)

size_t zg_pk_dedup_insert(cudaStream_t s, const u8* g_digest, u64 lo, u64 hi, u32* table, u32 mask) {
	if (hi <= lo) return 0;
	ZG_LAUNCH(k_dedup_insert, PK_GRID(hi - lo), 128, 0, s, g_digest, lo, hi, table, mask);
	ZG_COUNT_LAUNCH();
	return 0;
}
size_t zg_pk_dedup_resolve(cudaStream_t s, const u8* g_digest, u64 base, u64 n, const u32* table, u32 mask, const u64* len, const u8* select,
                           u64* rep, u8* first, u64* isfirst64, u64* nblk, u64* clen) {
	ZG_LAUNCH(k_dedup_resolve, PK_GRID(n), 128, 0, s, g_digest, base, n, table, mask, len, select, rep, first, isfirst64, nblk, clen);
	ZG_COUNT_LAUNCH();
	return 0;
}
size_t zg_pk_build_ulist(cudaStream_t s, const u8* first, const u64* uidx, const u64* blk_first, u64 n, u32* ulist, u64* blk_base,
                         u64 nuniq, u64 nblocks) {
	ZG_LAUNCH(k_build_ulist, PK_GRID(n), 128, 0, s, first, uidx, blk_first, n, ulist, blk_base, nuniq, nblocks);
	ZG_COUNT_LAUNCH();
	return 0;
}
size_t zg_pk_block_out_sizes(cudaStream_t s, const u32* blk_csize, u64 nblocks, u64* blk_out) {
	ZG_LAUNCH(k_block_out_sizes, PK_GRID(nblocks), 128, 0, s, blk_csize, nblocks, blk_out);
	ZG_COUNT_LAUNCH();
	return 0;
}
size_t zg_pk_frame_sizes(cudaStream_t s, const u32* ulist, const u64* blk_base, const u64* blk_pos, const u64* blk_total, const u64* len,
                         u64 nuniq, u32 flags, u64* frame_len_u) {
	ZG_LAUNCH(k_frame_sizes, PK_GRID(nuniq), 128, 0, s, ulist, blk_base, blk_pos, blk_total, len, nuniq, flags, frame_len_u);
	ZG_COUNT_LAUNCH();
	return 0;
}
size_t zg_pk_xxh64_list(cudaStream_t s, const u8* blob, const u64* off, const u64* len, const u32* ulist, u64 nuniq, u64* out) {
	ZG_LAUNCH(k_xxh64_list, PK_GRID(nuniq), 128, 0, s, blob, off, len, ulist, nuniq, out);
	ZG_LAUNCH(k_xxh64_list_warp, (u32)((nuniq + 3) / 4), 128, 0, s, blob, off, len, ulist, nuniq, out);
	ZG_COUNT_LAUNCH();
	ZG_COUNT_LAUNCH();
	return 0;
}
size_t zg_pk_frame_assemble(cudaStream_t s, const u8* blob, const u8* comp, const u64* file_off, const u64* comp_off, const u64* file_len,
                            const u32* ulist, const u64* blk_base, const u64* blk_pos, const u32* blk_csize, const u64* frame_off_u,
                            const u64* xxh, u64 nuniq, u64 nblocks, u32 flags, u64 archive_base, u8* frames_out) {
	ZG_LAUNCH(k_frame_assemble, (u32)((nblocks + 3) / 4), 128, 0, s, blob, comp, file_off, comp_off, file_len, ulist, blk_base, blk_pos,
	          blk_csize, frame_off_u, xxh, nuniq, nblocks, flags, archive_base, frames_out);
	ZG_COUNT_LAUNCH();
	return 0;
}
size_t zg_pk_record_answer(cudaStream_t s, const u32* ulist, const u64* frame_off_u, const u64* frame_len_u, u64 nuniq, u64 base,
                           u64* g_off, u64* g_len, const u64* rep, u64 n, u64* frame_off, u64* frame_len) {
	if (nuniq) {
		ZG_LAUNCH(k_record_frames, PK_GRID(nuniq), 128, 0, s, ulist, frame_off_u, frame_len_u, nuniq, base, g_off, g_len);
		ZG_COUNT_LAUNCH();
	}
	ZG_LAUNCH(k_answer_files, PK_GRID(n), 128, 0, s, rep, g_off, g_len, n, frame_off, frame_len);
	ZG_COUNT_LAUNCH();
	return 0;
}
