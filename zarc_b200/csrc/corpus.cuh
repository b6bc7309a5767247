// Synthetic corpus generator (SURVEY.md §8d shapes): text / random / src / log.
// Counter-based and integer-only, so the host (oracle / CPU baseline) and the device (bench)
// materialise bit-identical bytes from the same (key, length).  A file is a concatenation of
// independent <=64 KiB segments; `key` = f(seed, content_id, segment_idx) is computed by the caller.
// This is workload generation, not part of the content path.
#pragma once
#include "simt.h"

#define ZG_SEG_BYTES 65536u

enum { ZG_KIND_TEXT = 0, ZG_KIND_RANDOM = 1, ZG_KIND_SRC = 2, ZG_KIND_LOG = 3 };

struct ZgRng {
	u64 s;
};
ZG_HD u64 zg_mix64(u64 z) {
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}
ZG_HD u64 zg_rng_next(ZgRng& r) {
	r.s += 0x9E3779B97F4A7C15ULL;
	return zg_mix64(r.s);
}
// uniform in [0, n)
ZG_HD u32 zg_rng_below(ZgRng& r, u32 n) { return (u32)(((zg_rng_next(r) >> 32) * (u64)n) >> 32); }
// log-uniform rank in [0, 2^bits - 1): integer stand-in for Zipf(s=1)
ZG_HD u32 zg_rng_zipf(ZgRng& r, u32 bits) {
	u32 u = (u32)(zg_rng_next(r) >> 32);
	u64 t = (u64)u * bits;
	u32 e = (u32)(t >> 32), frac = (u32)t;
	return ((1u << e) + (u32)(((u64)(1u << e) * frac) >> 32)) - 1u;
}

struct ZgSink {
	u8* p;
	u32 n, cap;
};
ZG_HD bool zg_put(ZgSink& o, u8 c) {
	if (o.n >= o.cap) return false;
	o.p[o.n++] = c;
	return true;
}
ZG_HD void zg_put_str(ZgSink& o, const char* s) {
	while (*s) zg_put(o, (u8)*s++);
}
// vocabulary word w: 2..10 lowercase letters, a pure function of w
ZG_HD void zg_put_word(ZgSink& o, u32 w) {
	u64 h = zg_mix64(0x20240120ULL + w);
	u32 len = 2 + (u32)((h >> 56) % 9);
	for (u32 i = 0; i < len; i++) zg_put(o, (u8)('a' + ((h >> (5 * i)) & 31) % 26));
}
ZG_HD void zg_put_uint(ZgSink& o, u32 v, u32 min_digits) {
	char buf[10];
	u32 n = 0;
	do {
		buf[n++] = (char)('0' + v % 10);
		v /= 10;
	} while (v);
	while (n < min_digits) buf[n++] = '0';
	while (n) zg_put(o, (u8)buf[--n]);
}

ZG_HD void zg_gen_segment(u8* out, u32 len, u32 kind, u64 key) {
	ZgRng r{key};
	ZgSink o{out, 0, len};
	if (kind == ZG_KIND_RANDOM) {
		while (o.n < o.cap) {
			u64 v = zg_rng_next(r);
			for (int i = 0; i < 8 && o.n < o.cap; i++) o.p[o.n++] = (u8)(v >> (8 * i));
		}
	} else if (kind == ZG_KIND_TEXT) {
		u32 k = 0;
		while (o.n < o.cap) {
			zg_put_word(o, zg_rng_zipf(r, 12));
			zg_put(o, (++k % 12 == 0) ? '\n' : ' ');
		}
	} else if (kind == ZG_KIND_SRC) {
		const char* kw[15] = {"if", "for", "while", "return", "fn", "let", "const", "struct", "impl", "match", "pub", "use", "mod", "static", "else"};
		while (o.n < o.cap) {
			u32 tabs = zg_rng_below(r, 5);
			for (u32 i = 0; i < tabs; i++) zg_put(o, '\t');
			zg_put_str(o, kw[zg_rng_below(r, 15)]);
			zg_put(o, ' ');
			zg_put_word(o, zg_rng_below(r, 300));
			zg_put(o, '_');
			zg_put_word(o, zg_rng_below(r, 300));
			zg_put(o, '(');
			zg_put_word(o, zg_rng_zipf(r, 8));
			zg_put_str(o, ") { ");
			zg_put_str(o, kw[zg_rng_below(r, 15)]);
			zg_put(o, ' ');
			zg_put_word(o, zg_rng_zipf(r, 8));
			zg_put_str(o, "; }\n");
		}
	} else {
		const char* verb[3] = {"GET", "PUT", "POST"};
		const u32 status[5] = {200, 200, 200, 404, 500};
		u32 t = 1700000000u + (u32)(key & 0xffffff);
		while (o.n < o.cap) {
			t += zg_rng_below(r, 4);
			zg_put_uint(o, t, 1);
			zg_put_str(o, " host");
			zg_put_uint(o, 1 + zg_rng_below(r, 20), 2);
			zg_put_str(o, " svc[");
			zg_put_uint(o, 100 + zg_rng_below(r, 900), 1);
			zg_put_str(o, "]: ");
			zg_put_str(o, verb[zg_rng_below(r, 3)]);
			zg_put_str(o, " /api/v1/");
			zg_put_word(o, zg_rng_zipf(r, 12));
			zg_put(o, '/');
			zg_put_uint(o, 1 + zg_rng_below(r, 99999), 1);
			zg_put_str(o, " status=");
			zg_put_uint(o, status[zg_rng_below(r, 5)], 1);
			zg_put_str(o, " dur=");
			zg_put_uint(o, 1 + zg_rng_below(r, 3000), 1);
			zg_put_str(o, "ms\n");
		}
	}
}
