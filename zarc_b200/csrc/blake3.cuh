// BLAKE3 device functions (hash mode, 32-byte output).
// Replaces blake3::hash (crates/zarc/src/encode/content_frame.rs:26, integrity.rs:110) and
// blake3::Hasher::update/finalize (decode/frame_iterator.rs:99,77).  Constants: BLAKE3 spec
// (SURVEY.md App. B).  The kernels (one 1 KiB chunk per lane, tree levels merged pair by pair) are in blake3.cu.
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

// Two forms of G that trade ALU-pipe instructions (the xors and rotates can only issue there, and it is the busier
// pipe) for FMA-pipe ones.  VAR & 1: the second addition of a + b + m as a multiply-add by `one`, a register holding 1
// that the compiler cannot see through.  VAR & 2: the rotation by 12 as a 32 x 32 -> 64 multiplication by 2^20 whose
// halves are added (`k20` = one << 20).
template <int VAR>
ZG_DEV u32 b3_add3(u32 a, u32 b, u32 m, u32 one) { return (VAR & 1) ? ((a + m) + b * one) : (a + b + m); }
template <int VAR>
ZG_DEV u32 b3_rot12(u32 x, u32 one, u32 k20) {
	if (VAR & 2) {
		u64 p = (u64)x * k20;
		return (u32)(p >> 32) * one + (u32)p;
	}
	return b3_ror12(x);
}
#define B3_G(a, b, c, d, mx, my)                      \
	do {                                              \
		a = b3_add3<VAR>(a, b, mx, one);              \
		d = b3_ror16(d ^ a);                          \
		c = c + d;                                    \
		b = b3_rot12<VAR>(b ^ c, one, k20);           \
		a = b3_add3<VAR>(a, b, my, one);              \
		d = b3_ror8(d ^ a);                           \
		c = c + d;                                    \
		b = b3_ror7(b ^ c);                           \
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
template <int VAR>
ZG_DEV void b3_compress_v(u32 cv[8], const u32 m[16], u32 ctr_lo, u32 ctr_hi, u32 block_len, u32 flags, u32 one) {
	const u32 k20 = one << 20;
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
ZG_DEV void b3_compress(u32 cv[8], const u32 m[16], u32 ctr_lo, u32 ctr_hi, u32 block_len, u32 flags) {
	b3_compress_v<0>(cv, m, ctr_lo, ctr_hi, block_len, flags, 1u);
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
