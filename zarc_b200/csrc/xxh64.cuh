// XXH64 device functions (seed 0 for the Zstandard Content_Checksum, RFC 8878 §3.1.1; the low
// 32 bits are what libzstd appends when ZSTD_c_checksumFlag is set, as crates/zarc-cli/src/pack.rs:227
// does).  Constants: SURVEY.md App. D.  The four lanes are a serial multiply-rotate chain per input.
#pragma once
#include "simt.h"

#define XXP1 0x9E3779B185EBCA87ULL
#define XXP2 0xC2B2AE3D27D4EB4FULL
#define XXP3 0x165667B19E3779F9ULL
#define XXP4 0x85EBCA77C2B2AE63ULL
#define XXP5 0x27D4EB2F165667C5ULL

ZG_DEV u64 xx_rotl(u64 x, int n) { return (x << n) | (x >> (64 - n)); }
ZG_DEV u64 xx_round(u64 acc, u64 v) { return xx_rotl(acc + v * XXP2, 31) * XXP1; }
ZG_DEV u64 xx_merge(u64 h, u64 v) { return (h ^ xx_round(0, v)) * XXP1 + XXP4; }

// The accumulator recurrence acc' = rotl(acc + x, 31) * P1 (x = data * P2, off the chain) is the serial bottleneck of a
// big input: written naively it is six dependent operations per step (64-bit add with carry, two funnel shifts, a
// wide multiply and two cross-term multiplies).  Carrying b = acc + x instead gives b' = rotl(b, 31) * P1 + x', where
// the addition rides on the multiply-add: funnel shift -> mad.wide (64-bit addend) -> the two cross terms = 4 levels.
struct XxChain {
	u32 lo, hi;  // b = acc + x of the step about to be taken
};
ZG_DEV void xx_chain_start(XxChain& c, u64 acc, u64 x) {
	u64 b = acc + x;
	c.lo = (u32)b;
	c.hi = (u32)(b >> 32);
}
// one step: consumes the pending b, takes the NEXT stripe's x (or 0 for the last step) -> new b
ZG_DEV void xx_chain_step(XxChain& c, u64 xnext) {
	const u32 pl = (u32)XXP1, ph = (u32)(XXP1 >> 32);
	u32 rl = __funnelshift_r(c.hi, c.lo, 1);   // rotl64(b, 31), low and high words
	u32 rh = __funnelshift_r(c.lo, c.hi, 1);
	u32 cross = rl * ph + rh * pl;
#ifdef ZG_EMU
	u64 t = (u64)rl * pl + xnext;
#else
	u64 t;
	asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"(rl), "r"(pl), "l"(xnext));
#endif
	c.lo = (u32)t;
	c.hi = (u32)(t >> 32) + cross;
}
ZG_DEV u64 xx_chain_value(const XxChain& c) { return ((u64)c.hi << 32) | c.lo; }

struct XxState {
	u64 v1, v2, v3, v4;
};
ZG_DEV void xx_init(XxState& s, u64 seed) {
	s.v1 = seed + XXP1 + XXP2;
	s.v2 = seed + XXP2;
	s.v3 = seed;
	s.v4 = seed - XXP1;
}
ZG_DEV void xx_stripe(XxState& s, u64 a, u64 b, u64 c, u64 d) {
	s.v1 = xx_round(s.v1, a);
	s.v2 = xx_round(s.v2, b);
	s.v3 = xx_round(s.v3, c);
	s.v4 = xx_round(s.v4, d);
}
// finish: `tail` points at the < 32 remaining bytes, n = total length
ZG_DEV u64 xx_finish(const XxState& s, const u8* tail, u64 n, u64 seed) {
	u64 h;
	if (n >= 32) {
		h = xx_rotl(s.v1, 1) + xx_rotl(s.v2, 7) + xx_rotl(s.v3, 12) + xx_rotl(s.v4, 18);
		h = xx_merge(h, s.v1);
		h = xx_merge(h, s.v2);
		h = xx_merge(h, s.v3);
		h = xx_merge(h, s.v4);
	} else {
		h = seed + XXP5;
	}
	h += n;
	u32 rem = (u32)(n & 31);
	while (rem >= 8) {
		h = xx_rotl(h ^ xx_round(0, zg_ld64(tail)), 27) * XXP1 + XXP4;
		tail += 8;
		rem -= 8;
	}
	if (rem >= 4) {
		h = xx_rotl(h ^ ((u64)zg_ld32(tail) * XXP1), 23) * XXP2 + XXP3;
		tail += 4;
		rem -= 4;
	}
	while (rem) {
		h = xx_rotl(h ^ ((u64)*tail * XXP5), 11) * XXP1;
		tail++;
		rem--;
	}
	h ^= h >> 33;
	h *= XXP2;
	h ^= h >> 29;
	h *= XXP3;
	h ^= h >> 32;
	return h;
}
ZG_DEV u64 xx_hash(const u8* p, u64 n, u64 seed) {
	XxState s;
	xx_init(s, seed);
	u64 stripes = n >> 5;
	if (((uintptr_t)p & 7) == 0) {
		const ulonglong2* q = (const ulonglong2*)p;
		if (((uintptr_t)p & 15) == 0) {
			for (u64 i = 0; i < stripes; i++) {
				ulonglong2 a = q[2 * i], b = q[2 * i + 1];
				xx_stripe(s, a.x, a.y, b.x, b.y);
			}
		} else {
			const u64* w = (const u64*)p;
			for (u64 i = 0; i < stripes; i++) xx_stripe(s, w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
		}
	} else {
		for (u64 i = 0; i < stripes; i++) {
			const u8* b = p + 32 * i;
			xx_stripe(s, zg_ld64(b), zg_ld64(b + 8), zg_ld64(b + 16), zg_ld64(b + 24));
		}
	}
	return xx_finish(s, p + 32 * stripes, n, seed);
}

// Inputs of XX_WARP_MIN bytes and more are hashed by a whole warp.  The recurrence itself cannot be
// split (each accumulator is a serial multiply-rotate chain over the stripes), but a single thread
// also pays a memory round trip per 32-byte stripe; here the 32 lanes fetch a chunk (2 KiB: 64 stripes) at a
// time, coalesced and one chunk ahead, into shared memory, and lanes 0..3 run one accumulator each
// from there: the chain's arithmetic latency is all that is left (2 GB/s per input; many inputs run
// side by side).  `sb`: XX_SB_WORDS u64 of shared memory per warp.  All lanes call; all lanes get the hash.
// (threshold: below it one THREAD per input -- four accumulators interleaved, every lane of a warp busy on its own input.
// On a source tree (1-64 KiB files) thresholds of 8 / 32 / 128 KiB hash at 1 723 / 1 873 / 1 985 GB/s: the warp path pays
// for its hand-offs and its serial tail on inputs this small; it is for inputs whose own chain is the critical path)
#ifndef XX_WARP_MIN
#define XX_WARP_MIN 131072u
#endif
#define XX_PREFETCH_CHUNKS 16u
// a warp's unit of work: 2 KiB (64 stripes per accumulator chain between two warp-wide hand-offs; with 1 KiB the
// hand-off -- products, stores, barrier, the first loads of the chain -- was 22 % of a chunk's time)
// (inputs below XX_BIG_MIN keep 1 KiB: a 2 KiB unit leaves up to 63 stripes to the serial tail of lane 0, and the
// 8-64 KiB files of a source tree hash 10 % slower with it)
#define XX_BIG_MIN (1u << 20)
#define XX_CHUNK_KIB 2                     // unit of inputs of XX_BIG_MIN bytes and more, and of resumed (staged) hashes
#define XX_CHUNK_SHIFT 11
#define XX_SB_WORDS (128 * XX_CHUNK_KIB)   // u64 of shared memory per warp
// The accumulators over the whole chunks (of 2^XX_CHUNK_SHIFT bytes) [c0, c1) of the input at p, continuing from `acc` (lane k < 4 carries
// accumulator k in and out; the other lanes' value is ignored).  Resumable: a big input can be hashed piece by piece as
// it becomes available (the staged decoder does).
// (out of line: inlined, its schedule depended on the kernel around it -- in one of the three kernels that use it the
// chain's shared-memory loads were issued one at a time, right before their use, and the input ran 25 % slower)
template <int KIB>
ZG_DEV_NOINLINE u64 xx_warp_chunks_t(const u8* p, u64 c0, u64 c1, u64 acc, u64* sb) {
	constexpr int SHIFT = 9 + KIB;  // 1 KiB: 10, 2 KiB: 11
	u32 lane = zg_lane();
	const u8* q = p + 32 * lane;
	// The lane's stripes of a chunk (stripe `lane` and, 1 KiB on, stripe `lane + 32`) as aligned words; they are combined
	// into u64 (a funnel shift when the input is not word aligned) only AFTER the chains of the chunk before have run:
	// combined at once, the shifts would wait for the loads in front of the chains, and every chunk would cost a memory
	// round trip on top of its serial steps.
	const u32* wq = (const u32*)((uintptr_t)q & ~(uintptr_t)3);
	const u32 sh = (u32)((uintptr_t)q & 3) * 8;
	u32 raw[KIB][9];
	ZG_UNROLL
	for (int h = 0; h < KIB; h++) {
		ZG_UNROLL
		for (int k = 0; k < 9; k++) raw[h][k] = 0;
	}
	if (c0 < c1) {
		ZG_UNROLL
		for (int h = 0; h < KIB; h++) {
			const u32* w0 = wq + (c0 << (SHIFT - 2)) + 256 * h;
			ZG_UNROLL
			for (int k = 0; k < 8; k++) raw[h][k] = w0[k];
			if (sh) raw[h][8] = w0[8];
		}
	}
	for (u64 c = c0; c < c1; c++) {
		// a warp has one chunk in flight: without help it would stream at one DRAM latency per chunk (measured 1.3 GB/s on
		// a 4 GiB input).  The lines 16 chunks ahead are asked into L2 now, so that the loads below find them there.
		if (c + XX_PREFETCH_CHUNKS < c1) {
			ZG_UNROLL
			for (int h = 0; h < KIB; h++) zg_prefetch_l2(q + ((c + XX_PREFETCH_CHUNKS) << SHIFT) + 1024 * h);
		}
		ZG_UNROLL
		for (int h = 0; h < KIB; h++) {
			ZG_UNROLL
			for (int j = 0; j < 4; j++) {  // the products are off the chain: all 32 lanes make them
				u64 r = ((u64)__funnelshift_r(raw[h][2 * j + 1], raw[h][2 * j + 2], sh) << 32) | __funnelshift_r(raw[h][2 * j], raw[h][2 * j + 1], sh);
				sb[128 * h + 4 * lane + j] = r * XXP2;
			}
		}
		__syncwarp();
		if (c + 1 < c1) {  // in flight while the chains run
			ZG_UNROLL
			for (int h = 0; h < KIB; h++) {
				const u32* wn = wq + ((c + 1) << (SHIFT - 2)) + 256 * h;
				ZG_UNROLL
				for (int k = 0; k < 8; k++) raw[h][k] = wn[k];
				if (sh) raw[h][8] = wn[8];
			}
		}
		if (lane < 4) {
			// b = acc + x0, then one step per stripe, each absorbing the next stripe's x (the last absorbs 0: b becomes acc)
			XxChain ch;
			xx_chain_start(ch, acc, sb[lane]);
			ZG_UNROLL
			for (u32 i = 1; i < 32 * KIB; i++) xx_chain_step(ch, sb[4 * i + lane]);
			xx_chain_step(ch, 0);
			acc = xx_chain_value(ch);
		}
		__syncwarp();
	}
	return acc;
}
ZG_DEV u64 xx_warp_chunks(const u8* p, u64 c0, u64 c1, u64 acc, u64* sb) { return xx_warp_chunks_t<XX_CHUNK_KIB>(p, c0, c1, acc, sb); }
ZG_DEV u64 xx_warp_acc0() {  // seed 0
	u32 lane = zg_lane();
	return lane == 0 ? XXP1 + XXP2 : lane == 1 ? XXP2 : lane == 2 ? 0ull : 0ull - XXP1;
}
// the hash of p[0..n) given the accumulators after its first c0 chunks of KIB KiB (xx_warp_acc0() and 0 for the whole input)
template <int KIB>
ZG_DEV u64 xx_hash_warp_from_t(const u8* p, u64 n, u64 c0, u64 acc, u64* sb) {
	constexpr int SHIFT = 9 + KIB;
	u32 lane = zg_lane();
	u64 chunks = n >> SHIFT;
	acc = xx_warp_chunks_t<KIB>(p, c0, chunks, acc, sb);
	XxState s;
	s.v1 = __shfl_sync(ZG_FULL, acc, 0);
	s.v2 = __shfl_sync(ZG_FULL, acc, 1);
	s.v3 = __shfl_sync(ZG_FULL, acc, 2);
	s.v4 = __shfl_sync(ZG_FULL, acc, 3);
	u64 h = 0;
	if (lane == 0) {
		const u8* t = p + (chunks << SHIFT);
		u64 stripes = (n >> 5) - (chunks << (SHIFT - 5));  // fewer than a chunk's worth left
		for (u64 i = 0; i < stripes; i++) {
			const u8* b = t + 32 * i;
			xx_stripe(s, zg_ld64(b), zg_ld64(b + 8), zg_ld64(b + 16), zg_ld64(b + 24));
		}
		h = xx_finish(s, t + 32 * stripes, n, 0);
	}
	return __shfl_sync(ZG_FULL, h, 0);
}
// resumed from a staged frame's partial hash (units of XX_CHUNK_KIB), or from the start
ZG_DEV u64 xx_hash_warp_from(const u8* p, u64 n, u64 c0, u64 acc, u64* sb) {
	if (c0 || n >= XX_BIG_MIN) return xx_hash_warp_from_t<XX_CHUNK_KIB>(p, n, c0, acc, sb);
	return xx_hash_warp_from_t<1>(p, n, 0, acc, sb);
}
ZG_DEV u64 xx_hash_warp(const u8* p, u64 n, u64* sb) { return xx_hash_warp_from(p, n, 0, xx_warp_acc0(), sb); }
