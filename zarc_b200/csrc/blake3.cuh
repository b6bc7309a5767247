// BLAKE3 device functions (hash mode, 32-byte output).
// Replaces blake3::hash (crates/zarc/src/encode/content_frame.rs:26, integrity.rs:110) and
// blake3::Hasher::update/finalize (decode/frame_iterator.rs:99,77).  Constants: BLAKE3 spec
// (SURVEY.md App. B).  One 1 KiB chunk per lane, 32 chunks per warp, shuffle-based tree merge.
#pragma once
#include "simt.h"

#define B3_CHUNK_START 1u
#define B3_CHUNK_END 2u
#define B3_PARENT 4u
#define B3_ROOT 8u

#define B3_IV0 0x6A09E667u
#define B3_IV1 0xBB67AE85u
#define B3_IV2 0x3C6EF372u
#define B3_IV3 0xA54FF53Au
#define B3_IV4 0x510E527Fu
#define B3_IV5 0x9B05688Cu
#define B3_IV6 0x1F83D9ABu
#define B3_IV7 0x5BE0CD19u

ZG_DEV u32 b3_ror16(u32 x) { return __byte_perm(x, x, 0x1032); }
ZG_DEV u32 b3_ror8(u32 x) { return __byte_perm(x, x, 0x0321); }
ZG_DEV u32 b3_ror12(u32 x) { return __funnelshift_r(x, x, 12); }
ZG_DEV u32 b3_ror7(u32 x) { return __funnelshift_r(x, x, 7); }

#define B3_G(a, b, c, d, mx, my) \
	do {                         \
		a = a + b + (mx);        \
		d = b3_ror16(d ^ a);     \
		c = c + d;               \
		b = b3_ror12(b ^ c);     \
		a = a + b + (my);        \
		d = b3_ror8(d ^ a);      \
		c = c + d;               \
		b = b3_ror7(b ^ c);      \
	} while (0)

// One round with the message words addressed through a compile-time schedule row.
#define B3_ROUND(m, i0, i1, i2, i3, i4, i5, i6, i7, i8, i9, i10, i11, i12, i13, i14, i15) \
	do {                                                                                  \
		B3_G(s0, s4, s8, s12, m[i0], m[i1]);                                              \
		B3_G(s1, s5, s9, s13, m[i2], m[i3]);                                              \
		B3_G(s2, s6, s10, s14, m[i4], m[i5]);                                             \
		B3_G(s3, s7, s11, s15, m[i6], m[i7]);                                             \
		B3_G(s0, s5, s10, s15, m[i8], m[i9]);                                             \
		B3_G(s1, s6, s11, s12, m[i10], m[i11]);                                           \
		B3_G(s2, s7, s8, s13, m[i12], m[i13]);                                            \
		B3_G(s3, s4, s9, s14, m[i14], m[i15]);                                            \
	} while (0)

// cv <- compress(cv, m, counter, block_len, flags), truncated to the 8-word chaining value.
// The 7 schedule rows are the message permutation [2,6,3,10,7,0,4,13,1,11,12,5,9,14,15,8] iterated.
ZG_DEV void b3_compress(u32 cv[8], const u32 m[16], u32 ctr_lo, u32 ctr_hi, u32 block_len, u32 flags) {
	u32 s0 = cv[0], s1 = cv[1], s2 = cv[2], s3 = cv[3], s4 = cv[4], s5 = cv[5], s6 = cv[6], s7 = cv[7];
	u32 s8 = B3_IV0, s9 = B3_IV1, s10 = B3_IV2, s11 = B3_IV3, s12 = ctr_lo, s13 = ctr_hi, s14 = block_len, s15 = flags;
	B3_ROUND(m, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
	B3_ROUND(m, 2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8);
	B3_ROUND(m, 3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1);
	B3_ROUND(m, 10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6);
	B3_ROUND(m, 12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4);
	B3_ROUND(m, 9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7);
	B3_ROUND(m, 11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13);
	cv[0] = s0 ^ s8;
	cv[1] = s1 ^ s9;
	cv[2] = s2 ^ s10;
	cv[3] = s3 ^ s11;
	cv[4] = s4 ^ s12;
	cv[5] = s5 ^ s13;
	cv[6] = s6 ^ s14;
	cv[7] = s7 ^ s15;
}

ZG_DEV void b3_set_iv(u32 cv[8]) {
	cv[0] = B3_IV0;
	cv[1] = B3_IV1;
	cv[2] = B3_IV2;
	cv[3] = B3_IV3;
	cv[4] = B3_IV4;
	cv[5] = B3_IV5;
	cv[6] = B3_IV6;
	cv[7] = B3_IV7;
}

// 64 message bytes at an arbitrarily aligned address
ZG_DEV void b3_load_full(const u8* p, u32 m[16]) {
	uintptr_t a = (uintptr_t)p;
	if ((a & 15) == 0) {
		const uint4* q = (const uint4*)p;
		ZG_UNROLL
		for (int i = 0; i < 4; i++) {
			uint4 v = q[i];
			m[4 * i] = v.x;
			m[4 * i + 1] = v.y;
			m[4 * i + 2] = v.z;
			m[4 * i + 3] = v.w;
		}
	} else if ((a & 3) == 0) {
		const u32* q = (const u32*)p;
		ZG_UNROLL
		for (int i = 0; i < 16; i++) m[i] = q[i];
	} else {
		const u32* q = (const u32*)(a & ~(uintptr_t)3);
		u32 sh = (u32)(a & 3) * 8;
		u32 prev = q[0];
		ZG_UNROLL
		for (int i = 0; i < 16; i++) {
			u32 nx = q[i + 1];
			m[i] = __funnelshift_r(prev, nx, sh);
			prev = nx;
		}
	}
}
// the zero-padded final block of a chunk (len in 0..63, or 64 via the full path)
ZG_DEV void b3_load_partial(const u8* p, u32 len, u32 m[16]) {
	ZG_UNROLL
	for (int w = 0; w < 16; w++) {
		u32 v = 0;
		ZG_UNROLL
		for (int b = 0; b < 4; b++)
			if ((u32)(4 * w + b) < len) v |= (u32)p[4 * w + b] << (8 * b);
		m[w] = v;
	}
}

// chaining value of one chunk (n in 0..1024 bytes)
ZG_DEV void b3_chunk_cv(const u8* p, u32 n, u64 chunk_idx, bool is_root, u32 cv[8]) {
	b3_set_iv(cv);
	u32 nblocks = n == 0 ? 1 : (n + 63) >> 6;
	u32 clo = (u32)chunk_idx, chi = (u32)(chunk_idx >> 32);
	u32 m[16];
	for (u32 b = 0; b + 1 < nblocks; b++) {
		b3_load_full(p + 64 * b, m);
		b3_compress(cv, m, clo, chi, 64, b == 0 ? B3_CHUNK_START : 0);
	}
	u32 lastlen = n - 64 * (nblocks - 1);
	if (lastlen == 64) b3_load_full(p + 64 * (nblocks - 1), m);
	else b3_load_partial(p + 64 * (nblocks - 1), lastlen, m);
	u32 flags = (nblocks == 1 ? B3_CHUNK_START : 0) | B3_CHUNK_END | (is_root ? B3_ROOT : 0);
	b3_compress(cv, m, clo, chi, lastlen, flags);
}

ZG_DEV void b3_parent_cv(const u32 l[8], const u32 r[8], bool is_root, u32 out[8]) {
	u32 m[16];
	ZG_UNROLL
	for (int i = 0; i < 8; i++) {
		m[i] = l[i];
		m[8 + i] = r[i];
	}
	b3_set_iv(out);
	b3_compress(out, m, 0, 0, 64, B3_PARENT | (is_root ? B3_ROOT : 0));
}

// Reduce `cnt` (1..32) equal-level nodes held by lanes 0..cnt-1 to one node in lane 0.  Pairs from
// the left at every level, an odd node is promoted: for left-full trees this is exactly BLAKE3's
// tree shape.  `root_at_top`: apply ROOT to the last merge (only when this is the whole input).
ZG_DEV void b3_warp_reduce(u32 cv[8], u32 cnt, bool root_at_top) {
	u32 lane = zg_lane();
	for (u32 w = cnt; w > 1; w = (w + 1) >> 1) {
		u32 l[8], r[8];
		ZG_UNROLL
		for (int i = 0; i < 8; i++) {
			l[i] = __shfl_sync(ZG_FULL, cv[i], (int)((2 * lane) & 31));
			r[i] = __shfl_sync(ZG_FULL, cv[i], (int)((2 * lane + 1) & 31));
		}
		if (2 * lane + 1 < w) b3_parent_cv(l, r, root_at_top && w == 2, cv);
		else if (2 * lane < w) {
			ZG_UNROLL
			for (int i = 0; i < 8; i++) cv[i] = l[i];
		}
	}
}

// Stack of completed subtrees (units of 32 leaves), owned by one warp; lane 0 is the only writer.
struct B3Stack {
	u32 cv[52][8];
	u32 depth;
};

// lane 0: push a full 32-leaf node; `total` = number of full nodes pushed including this one.
ZG_DEV void b3_stack_push(B3Stack* st, u32 cv[8], u64 total) {
	while ((total & 1) == 0) {
		u32 l[8];
		st->depth--;
		ZG_UNROLL
		for (int i = 0; i < 8; i++) l[i] = st->cv[st->depth][i];
		u32 o[8];
		b3_parent_cv(l, cv, false, o);
		ZG_UNROLL
		for (int i = 0; i < 8; i++) cv[i] = o[i];
		total >>= 1;
	}
	ZG_UNROLL
	for (int i = 0; i < 8; i++) st->cv[st->depth][i] = cv[i];
	st->depth++;
}
// lane 0: fold the stack onto the right-most node; the last merge is the root.
ZG_DEV void b3_stack_fold(B3Stack* st, u32 cv[8]) {
	while (st->depth > 0) {
		u32 l[8], o[8];
		st->depth--;
		ZG_UNROLL
		for (int i = 0; i < 8; i++) l[i] = st->cv[st->depth][i];
		b3_parent_cv(l, cv, st->depth == 0, o);
		ZG_UNROLL
		for (int i = 0; i < 8; i++) cv[i] = o[i];
	}
}

// Whole-warp BLAKE3 of [data, data+n): result (8 words) valid in lane 0.
ZG_DEV void b3_warp_hash(const u8* data, u64 n, B3Stack* st, u32 cv[8]) {
	u32 lane = zg_lane();
	u64 nchunks = n == 0 ? 1 : (n + 1023) >> 10;
	u64 nbatches = (nchunks + 31) >> 5;
	if (lane == 0) st->depth = 0;
	__syncwarp();
	for (u64 b = 0; b < nbatches; b++) {
		u64 c = b * 32 + lane;
		u32 cnt = (u32)zg_min<u64>((u64)32, nchunks - b * 32);
		if (c < nchunks) {
			u64 o = c << 10;
			u32 len = (u32)zg_min<u64>((u64)1024, n - o);
			b3_chunk_cv(data + o, len, c, nchunks == 1, cv);
		}
		b3_warp_reduce(cv, cnt, nbatches == 1);
		if (lane == 0) {
			if (b + 1 < nbatches) b3_stack_push(st, cv, b + 1);
			else b3_stack_fold(st, cv);
		}
		__syncwarp();
	}
}
