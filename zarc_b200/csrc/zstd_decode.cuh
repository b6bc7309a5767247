// Device helpers shared by the two Zstandard decode paths (RFC 8878; replaces DCtx::decompress_stream as driven by
// crates/zarc/src/decode/zstd_iterator.rs:88-153): the fused lane-per-frame kernel (zstd_decode.cu) and the staged
// pipeline for multi-block frames (zstd_decode_staged.cu).  Header and section parsers, Huffman / FSE table builds,
// the lane-private FSE sequence decoder and the row-wise sequence executor.
#pragma once
#include "common.h"
#include "zstd_common.cuh"

#define ZD_WARPS 4
#ifndef ZD_MIN_CTAS
#define ZD_MIN_CTAS 5                    // occupancy target: 20 warps / SM
#endif
#define ZD_SEQ_ARENA (1u << 17)          // u64 entries of sequence staging per warp (1 MiB)
#define ZD_LITBUF (ZS_BLOCK_MAX + 64)    // bytes of literal staging per warp
#define ZD_TAB_SLOT 1280u                // u32 entries per lane: LL 512 | ML 512 | OF 256
#define ZD_HUFSAVE 272u                  // bytes per lane: 256 weights + count
#define ZD_OFF_MAX ((1u << 28) - 1u)
#define ZD_BATCH_BYTES (128u << 10)        // least compressed bytes per warp batch (see the hand-out in the kernel)

struct ZdWarp {
	u32 tab[512];      // table under construction (FSE sequence table or Huffman-weight table)
	u16 huf[2048];     // sym | nbBits << 8
	u8 weights[256];
	i16 norm[256];
	u16 next[256];
	u32 misc[8];
};

// flags
#define ZD_F_ACTIVE 1u      // frame still has blocks to decode
#define ZD_F_LAST 2u        // the block in flight is the frame's last
#define ZD_F_CKSUM 4u       // frame has a Content_Checksum
#define ZD_F_HUF_OK 8u      // a Huffman table has been defined (saved weights are valid)
#define ZD_F_LL_OK 16u
#define ZD_F_ML_OK 32u
#define ZD_F_OF_OK 64u
#define ZD_F_LL_DEF 256u    // table = the predefined one (not in the slot)
#define ZD_F_ML_DEF 512u
#define ZD_F_OF_DEF 1024u
#define ZD_F_HAS_CK 2048u   // checksum field was read

// one frame's decoding state; lives in the registers of the frame's lane
struct ZdLane {
	const u8* src;
	u8* out;
	u64 n, ip, cap, opos, fcs;
	u64 base;           // lowest output position a match may reach (0; the block's start for a split frame's block)
	u32 status, flags, fcs_len, cksum;
	u32 rep0, rep1, rep2;
	u32 ll_log, ml_log, of_log;
	// block in flight
	const u8* blk;      // block body
	u32 blk_n;
	u32 nseq, nseq_left, seq_base;
	ZsBack b;
	u32 sl, so, sm;
};

template <typename T>
ZG_DEV T zd_bc(T v, int f) { return __shfl_sync(ZG_FULL, v, f); }
ZG_DEV const u8* zd_bc(const u8* v, int f) { return (const u8*)(uintptr_t)__shfl_sync(ZG_FULL, (u64)(uintptr_t)v, f); }
ZG_DEV u8* zd_bc(u8* v, int f) { return (u8*)(uintptr_t)__shfl_sync(ZG_FULL, (u64)(uintptr_t)v, f); }

// every lane gets a copy of lane f's frame state
ZG_DEV ZdLane zd_bcast(const ZdLane& L, int f) {
	ZdLane U;
	U.src = zd_bc(L.src, f);
	U.out = zd_bc(L.out, f);
	U.n = zd_bc(L.n, f);
	U.ip = zd_bc(L.ip, f);
	U.cap = zd_bc(L.cap, f);
	U.opos = zd_bc(L.opos, f);
	U.fcs = zd_bc(L.fcs, f);
	U.status = zd_bc(L.status, f);
	U.flags = zd_bc(L.flags, f);
	U.fcs_len = zd_bc(L.fcs_len, f);
	U.cksum = zd_bc(L.cksum, f);
	U.rep0 = zd_bc(L.rep0, f);
	U.rep1 = zd_bc(L.rep1, f);
	U.rep2 = zd_bc(L.rep2, f);
	U.ll_log = zd_bc(L.ll_log, f);
	U.ml_log = zd_bc(L.ml_log, f);
	U.of_log = zd_bc(L.of_log, f);
	U.base = zd_bc(L.base, f);
	U.blk = zd_bc(L.blk, f);
	U.blk_n = zd_bc(L.blk_n, f);
	U.nseq = zd_bc(L.nseq, f);
	U.nseq_left = zd_bc(L.nseq_left, f);
	U.seq_base = zd_bc(L.seq_base, f);
	U.b.start = zd_bc(L.b.start, f);
	U.b.ptr = zd_bc(L.b.ptr, f);
	U.b.lo = zd_bc(L.b.lo, f);
	U.b.hi = zd_bc(L.b.hi, f);
	U.b.consumed = zd_bc(L.b.consumed, f);
	U.sl = zd_bc(L.sl, f);
	U.so = zd_bc(L.so, f);
	U.sm = zd_bc(L.sm, f);
	return U;
}

// ---------------------------------------------------------------------------------------------
// Loads of bytes that ANOTHER SM may have written during this kernel (the staged pipeline's match sources: output of
// earlier blocks, executed by other warps) must not be served from this SM's L1: CG = ld.global.cg (L2 only).
template <bool CG>
ZG_DEV u32 zd_ldw(const u32* p) {
#ifndef ZG_EMU
	if (CG) return __ldcg(p);
#endif
	return *p;
}
template <bool CG>
ZG_DEV u32 zd_ldb(const u8* p) {
#ifndef ZG_EMU
	if (CG) return __ldcg(p);
#endif
	return *p;
}
// warp-cooperative copies
template <bool CG>
ZG_DEV_NOINLINE void zg_warp_copy_t(u8* dst, const u8* src, u32 n) {
	u32 lane = zg_lane();
	if (n < 128) {
		for (u32 i = lane; i < n; i += 32) dst[i] = (u8)zd_ldb<CG>(src + i);
		return;
	}
	u32 head = (u32)((16 - ((uintptr_t)dst & 15)) & 15);
	if (lane < head) dst[lane] = (u8)zd_ldb<CG>(src + lane);
	dst += head;
	src += head;
	n -= head;
	u32 nvec = n >> 4;
	uintptr_t sa = (uintptr_t)src;
	const u32* sw = (const u32*)(sa & ~(uintptr_t)3);
	u32 sh = (u32)(sa & 3) * 8;
	for (u32 v = lane; v < nvec; v += 32) {
		const u32* q = sw + 4 * v;
		u32 w0 = zd_ldw<CG>(q), w1 = zd_ldw<CG>(q + 1), w2 = zd_ldw<CG>(q + 2), w3 = zd_ldw<CG>(q + 3);
		uint4 o;
		if (sh == 0) {
			o = make_uint4(w0, w1, w2, w3);
		} else {
			u32 w4 = zd_ldw<CG>(q + 4);
			o = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
			               __funnelshift_r(w3, w4, sh));
		}
		((uint4*)dst)[v] = o;
	}
	for (u32 i = (nvec << 4) + lane; i < n; i += 32) dst[i] = (u8)zd_ldb<CG>(src + i);
}
ZG_DEV void zg_warp_copy(u8* dst, const u8* src, u32 n) { zg_warp_copy_t<false>(dst, src, n); }
ZG_DEV_NOINLINE void zg_warp_fill(u8* dst, u32 byte, u32 n) {
	for (u32 i = zg_lane(); i < n; i += 32) dst[i] = (u8)byte;
}
// match copy inside the frame output: d[i] = d[i - off], forward semantics (overlap allowed)
template <bool CG>
ZG_DEV void zd_warp_match(u8* d, u32 off, u32 ml) {
	u32 lane = zg_lane();
	const u8* s = d - off;
	if (off >= 32) {
		bool overlap = off < ml;
		for (u32 i0 = 0; i0 < ml; i0 += 32) {
			u32 i = i0 + lane;
			if (i < ml) d[i] = (u8)zd_ldb<CG>(s + i);
			if (overlap) __syncwarp();
		}
	} else {
		for (u32 i = lane; i < ml; i += 32) d[i] = (u8)zd_ldb<CG>(s + i % off);
	}
}

// ---------------------------------------------------------------------------------------------
// Huffman tree description -> W->weights[0..nw) (the last weight is implied).  Returns bytes
// consumed (0 on error).  All lanes call.
ZG_DEV_NOINLINE u32 zd_read_huf_weights(ZdWarp* W, const u8* src, u32 n, u32* nw_out) {
	u32 lane = zg_lane();
	if (n < 1) return 0;
	u32 h = src[0];
	u32 used, nw;
	if (h >= 128) {
		nw = h - 127;
		used = 1 + ((nw + 1) >> 1);
		if (used > n) return 0;
		for (u32 i = lane; i < nw; i += 32) {
			u32 byte = src[1 + (i >> 1)];
			W->weights[i] = (u8)((i & 1) ? (byte & 15) : (byte >> 4));
		}
		__syncwarp();
	} else {
		u32 csz = h;
		used = 1 + csz;
		if (used > n || csz < 2) return 0;
		if (lane == 0) {
			u32 nsym = 0, log = 0;
			u32 nc = zs_read_ncount(src + 1, csz, 6, 255, W->norm, &nsym, &log);
			W->misc[0] = nc;
			W->misc[1] = nsym;
			W->misc[2] = log;
		}
		__syncwarp();
		u32 nc = W->misc[0], nsym = W->misc[1], log = W->misc[2];
		__syncwarp();
		if (nc == 0 || nc >= csz) return 0;
		zs_fse_build_dtable(W->tab, W->norm, nsym, log, W->next);
		if (lane == 0) {
			ZsBack b;
			u32 cnt = 0;
			bool ok = zs_back_init(b, src + 1 + nc, csz - nc);
			if (ok) {
				zs_back_reload(b);
				u32 s1 = zs_back_read(b, log), s2 = zs_back_read(b, log);
				// two interleaved states; when the stream runs dry after an update, the other
				// state's symbol is the last one (RFC 8878 §4.2.1.2)
				for (;;) {
					if (cnt > 253) {
						ok = false;
						break;
					}
					u32 e1 = W->tab[s1];
					W->weights[cnt++] = (u8)e1;
					zs_back_reload(b);
					s1 = (e1 >> 16) + zs_back_read(b, (e1 >> 8) & 0xff);
					if (zs_back_overflow(b)) {
						W->weights[cnt++] = (u8)W->tab[s2];
						break;
					}
					u32 e2 = W->tab[s2];
					W->weights[cnt++] = (u8)e2;
					s2 = (e2 >> 16) + zs_back_read(b, (e2 >> 8) & 0xff);
					if (zs_back_overflow(b)) {
						W->weights[cnt++] = (u8)W->tab[s1];
						break;
					}
				}
			}
			W->misc[0] = ok ? cnt : 0;
		}
		__syncwarp();
		nw = W->misc[0];
		__syncwarp();
		if (nw == 0) return 0;
	}
	*nw_out = nw;
	return used;
}

// W->weights[0..nw) -> W->huf.  Returns maxbits (0 on error).  All lanes call.
ZG_DEV_NOINLINE u32 zd_build_huf(ZdWarp* W, u32 nw) {
	u32 lane = zg_lane();
	// weights -> last weight, ranks, per-symbol start index (serial, <= 256 symbols)
	if (lane == 0) {
		u32 sum = 0;
		bool ok = true;
		u32 rank[13];
		for (u32 i = 0; i < 13; i++) rank[i] = 0;
		for (u32 i = 0; i < nw; i++) {
			u32 w = W->weights[i];
			if (w > 11) ok = false;
			else {
				if (w) sum += 1u << (w - 1);
				rank[w]++;
			}
		}
		u32 maxbits = 0;
		if (ok && sum != 0) {
			maxbits = zs_highbit(sum) + 1;
			u32 left = (1u << maxbits) - sum;
			if (maxbits > ZS_HUF_MAXLOG || (left & (left - 1))) ok = false;
			else {
				u32 last = zs_highbit(left) + 1;
				W->weights[nw] = (u8)last;
				rank[last]++;
			}
		} else ok = false;
		if (ok && (rank[1] < 2 || (rank[1] & 1))) ok = false;  // as libzstd HUF_readStats
		if (ok) {
			u32 start = 0;
			u32 rs[13];
			for (u32 w = 1; w <= 11; w++) {
				rs[w] = start;
				start += rank[w] << (w - 1);
			}
			for (u32 i = 0; i <= nw; i++) {
				u32 w = W->weights[i];
				if (w) {
					W->next[i] = (u16)rs[w];
					rs[w] += 1u << (w - 1);
				}
			}
		}
		W->misc[0] = ok ? 1 : 0;
		W->misc[1] = maxbits;
	}
	__syncwarp();
	bool ok = W->misc[0] != 0;
	u32 maxbits = W->misc[1];
	__syncwarp();
	if (!ok) return 0;
	nw += 1;
	// fill: long ranges cooperatively, short ranges by the owning lane
	for (u32 s0 = 0; s0 < nw; s0 += 32) {
		u32 s = s0 + lane;
		u32 w = s < nw ? W->weights[s] : 0;
		u32 len = w ? 1u << (w - 1) : 0;
		u32 start = w ? W->next[s] : 0;
		u32 entry = s | ((maxbits + 1 - w) << 8);
		if (len > 0 && len < 32)
			for (u32 k = 0; k < len; k++) W->huf[start + k] = (u16)entry;
		u32 big = __ballot_sync(ZG_FULL, len >= 32);
		while (big) {
			int l = __ffs((int)big) - 1;
			big &= big - 1;
			u32 bs = __shfl_sync(ZG_FULL, start, l), bl = __shfl_sync(ZG_FULL, len, l), be = __shfl_sync(ZG_FULL, entry, l);
			for (u32 k = lane; k < bl; k += 32) W->huf[bs + k] = (u16)be;
		}
	}
	__syncwarp();
	return maxbits;
}

// one Huffman stream, single thread
ZG_DEV_NOINLINE bool zd_huf_stream(const u16* huf, u32 maxbits, const u8* src, u32 n, u8* dst, u32 count) {
	ZsBack b;
	if (!zs_back_init(b, src, n)) return false;
	u32 i = 0;
	const u32 sh = 32u - maxbits;
	while (i < count) {
		zs_back_reload(b);
		u32 m = zg_min<u32>(count - i, 5u);  // 5 x 11 bits <= 57
		// the unread bits left-aligned in ah:al; every symbol then costs one shift for the index and a two-word
		// shift for the bits it used (bits before the start of the stream shift in as zeros)
		u32 c = b.consumed;
		u32 ah = c < 32 ? __funnelshift_l(b.lo, b.hi, c) : (c < 64 ? b.lo << (c & 31u) : 0u);
		u32 al = c < 32 ? b.lo << c : 0u;
		u32 used = 0;
		for (u32 k = 0; k < m; k++) {
			u32 e = huf[ah >> sh];
			dst[i + k] = (u8)e;
			u32 nb = e >> 8;
			ah = __funnelshift_l(al, ah, nb);
			al <<= nb;
			used += nb;
		}
		b.consumed = c + used;
		i += m;
	}
	zs_back_reload(b);
	return zs_back_finished(b);
}

// sequence table for one of LL/OF/ML into the lane's global slot `slot`.  All lanes call (uniform).
// Returns false on error; advances p.  `flags`: def_flag is set when the predefined table is in use.
ZG_DEV_NOINLINE bool zd_seq_table(ZdWarp* W, u32* slot, u32& log, u32& flags, u32 ok_flag, u32 def_flag, u32 mode, const u8*& p, const u8* end,
                         u32 maxlog, u32 maxsym, u32 def_log) {
	u32 lane = zg_lane();
	if (mode == 0) {
		log = def_log;
		flags |= ok_flag | def_flag;
		return true;
	}
	if (mode == 1) {
		if (p >= end) return false;
		u32 sym = *p++;
		if (sym > maxsym) return false;
		if (lane == 0) slot[0] = sym;
		log = 0;
		flags = (flags | ok_flag) & ~def_flag;
		return true;
	}
	if (mode == 2) {
		if (lane == 0) {
			u32 nsym = 0, lg = 0;
			u32 nc = zs_read_ncount(p, (u32)(end - p), maxlog, maxsym, W->norm, &nsym, &lg);
			W->misc[0] = nc;
			W->misc[1] = nsym;
			W->misc[2] = lg;
		}
		__syncwarp();
		u32 nc = W->misc[0], nsym = W->misc[1], lg = W->misc[2];
		__syncwarp();
		if (nc == 0) return false;
		zs_fse_build_dtable(W->tab, W->norm, nsym, lg, W->next);
		for (u32 i = lane; i < (1u << lg); i += 32) slot[i] = W->tab[i];
		__syncwarp();
		p += nc;
		log = lg;
		flags = (flags | ok_flag) & ~def_flag;
		return true;
	}
	return (flags & ok_flag) != 0;  // Repeat_Mode: the slot (or the predefined flag) still holds the table
}

// ---------------------------------------------------------------------------------------------
ZG_DEV void zd_fail(ZdLane& U, u32 code) {
	U.status = code;
	U.flags &= ~ZD_F_ACTIVE;
	U.nseq = U.nseq_left = 0;
}
// end of frame: optional checksum field, content-size check
ZG_DEV void zd_frame_finish(ZdLane& U) {
	U.flags &= ~ZD_F_ACTIVE;
	U.nseq = U.nseq_left = 0;
	if (U.flags & ZD_F_CKSUM) {
		if (U.ip + 4 > U.n) {
			U.status = ZS_E_SRC_SIZE;
			return;
		}
		U.cksum = zg_ld32(U.src + U.ip);
		U.flags |= ZD_F_HAS_CK;
		U.ip += 4;
	}
	if (U.fcs_len && U.fcs != U.opos) U.status = ZS_E_CORRUPT;
}

// frame header (lane-private)
ZG_DEV void zd_frame_header(ZdLane& L) {
	const u8* src = L.src;
	u64 n = L.n;
	if (n < 6) return zd_fail(L, ZS_E_SRC_SIZE);
	if (zg_ld32(src) != ZS_MAGIC) return zd_fail(L, ZS_E_PREFIX);
	u32 desc = src[4];
	u32 fcs_flag = desc >> 6, single = (desc >> 5) & 1, checksum = (desc >> 2) & 1, did_flag = desc & 3;
	if (desc & 8) return zd_fail(L, ZS_E_UNSUPPORTED);
	u64 ip = 5;
	u64 window = 0;
	if (!single) {
		u32 wd = src[ip++];
		u32 wl = 10 + (wd >> 3);
		if (wl > 31) return zd_fail(L, ZS_E_WINDOW);
		window = ((u64)1 << wl) + ((((u64)1 << wl) >> 3) * (wd & 7));
	}
	if (did_flag) {
		u32 dl = did_flag == 3 ? 4 : did_flag;
		if (ip + dl > n) return zd_fail(L, ZS_E_SRC_SIZE);
		u32 did = 0;
		for (u32 i = 0; i < dl; i++) did |= (u32)src[ip + i] << (8 * i);
		ip += dl;
		if (did) return zd_fail(L, ZS_E_DICT);
	}
	u32 fcs_len = fcs_flag == 0 ? single : fcs_flag == 1 ? 2 : fcs_flag == 2 ? 4 : 8;
	if (ip + fcs_len > n) return zd_fail(L, ZS_E_SRC_SIZE);
	u64 fcs = 0;
	for (u32 i = 0; i < fcs_len; i++) fcs |= (u64)src[ip + i] << (8 * i);
	if (fcs_len == 2) fcs += 256;
	ip += fcs_len;
	// libzstd's streaming decoder (what zstd_iterator.rs:29 creates) refuses windows > 2^27
	// ("Frame requires too much memory for decoding")
	if ((single ? fcs : window) > ((u64)1 << 27)) return zd_fail(L, ZS_E_WINDOW);
	L.ip = ip;
	L.fcs = fcs;
	L.fcs_len = fcs_len;
	if (checksum) L.flags |= ZD_F_CKSUM;
}

// ---------------------------------------------------------------------------------------------
// Literals_Section_Header (RFC 8878 §3.1.1.3.1.1).  Returns false on a malformed header.
struct ZdLitHdr {
	u32 ltype, regen, comp, hdr, streams;
};
ZG_DEV bool zd_lit_header(const u8* src, u32 n, ZdLitHdr& h) {
	if (n < 2) return false;
	u32 b0 = src[0];
	u32 sf = (b0 >> 2) & 3;
	h.ltype = b0 & 3;
	h.comp = 0;
	h.streams = 1;
	if (h.ltype < 2) {
		if (sf == 0 || sf == 2) {
			h.regen = b0 >> 3;
			h.hdr = 1;
		} else if (sf == 1) {
			h.regen = (b0 >> 4) | ((u32)src[1] << 4);
			h.hdr = 2;
		} else {
			if (n < 3) return false;
			h.regen = (b0 >> 4) | ((u32)src[1] << 4) | ((u32)src[2] << 12);
			h.hdr = 3;
		}
		if (h.regen > ZS_BLOCK_MAX) return false;
		h.comp = h.ltype == 0 ? h.regen : 1;
		return h.hdr + h.comp <= n;
	}
	if (n < 3) return false;
	u64 v = (u64)b0 | ((u64)src[1] << 8) | ((u64)src[2] << 16);
	if (sf == 0 || sf == 1) {
		h.regen = (u32)(v >> 4) & 1023;
		h.comp = (u32)(v >> 14) & 1023;
		h.hdr = 3;
		h.streams = sf == 0 ? 1 : 4;
	} else if (sf == 2) {
		if (n < 4) return false;
		v |= (u64)src[3] << 24;
		h.regen = (u32)(v >> 4) & 16383;
		h.comp = (u32)(v >> 18) & 16383;
		h.hdr = 4;
		h.streams = 4;
	} else {
		if (n < 5) return false;
		v |= ((u64)src[3] << 24) | ((u64)src[4] << 32);
		h.regen = (u32)(v >> 4) & 262143;
		h.comp = (u32)(v >> 22) & 262143;
		h.hdr = 5;
		h.streams = 4;
	}
	return h.regen <= ZS_BLOCK_MAX && h.hdr + h.comp <= n && h.regen != 0;
}

// The block's literals (uniform in the warp): Raw -> pointer into the block, RLE -> the byte,
// Huffman -> decoded into litbuf.  Returns a ZS_E_* code.
ZG_DEV_NOINLINE u32 zd_literals(ZdWarp* W, u32& flags, const u8* src, const ZdLitHdr& h, bool last, u8* litbuf, u8* hufsave, const u8*& lit,
                       bool& lit_rle, u32& rle_byte) {
	u32 lane = zg_lane();
	lit_rle = false;
	rle_byte = 0;
	if (h.ltype == 0) {
		lit = src + h.hdr;
		return ZS_OK;
	}
	if (h.ltype == 1) {
		lit_rle = true;
		rle_byte = src[h.hdr];
		lit = src;
		return ZS_OK;
	}
	const u8* lp = src + h.hdr;
	const u8* lend = lp + h.comp;
	u32 regen = h.regen;
	u32 huf_bits;
	if (h.ltype == 2) {
		u32 nw = 0;
		u32 used = zd_read_huf_weights(W, lp, h.comp, &nw);
		if (used == 0) return ZS_E_CORRUPT;
		if (!last) {  // a later Treeless block may need this tree again
			for (u32 i = lane; i < nw; i += 32) hufsave[i] = W->weights[i];
			if (lane == 0) {
				hufsave[256] = (u8)nw;
				hufsave[257] = (u8)(nw >> 8);
			}
		}
		huf_bits = zd_build_huf(W, nw);
		if (huf_bits == 0) return ZS_E_CORRUPT;
		flags |= ZD_F_HUF_OK;
		lp += used;
	} else {
		if (!(flags & ZD_F_HUF_OK)) return ZS_E_CORRUPT;
		u32 nw = (u32)hufsave[256] | ((u32)hufsave[257] << 8);
		__syncwarp();
		for (u32 i = lane; i < nw; i += 32) W->weights[i] = hufsave[i];
		__syncwarp();
		huf_bits = zd_build_huf(W, nw);
		if (huf_bits == 0) return ZS_E_CORRUPT;
	}
	bool ok = true;
	if (h.streams == 1) {
		if (lane == 0) ok = zd_huf_stream(W->huf, huf_bits, lp, (u32)(lend - lp), litbuf, regen);
	} else {
		if (lend - lp < 10) return ZS_E_CORRUPT;
		u32 s1 = zg_ld16(lp), s2 = zg_ld16(lp + 2), s3 = zg_ld16(lp + 4);
		lp += 6;
		u32 avail = (u32)(lend - lp);
		u32 seg = (regen + 3) >> 2;
		if (s1 + s2 + s3 >= avail || seg * 3 > regen) return ZS_E_CORRUPT;
		if (lane < 4) {
			u32 so = lane == 0 ? 0 : lane == 1 ? s1 : lane == 2 ? s1 + s2 : s1 + s2 + s3;
			u32 sn = lane == 0 ? s1 : lane == 1 ? s2 : lane == 2 ? s3 : avail - s1 - s2 - s3;
			u32 cnt = lane < 3 ? seg : regen - 3 * seg;
			ok = zd_huf_stream(W->huf, huf_bits, lp + so, sn, litbuf + lane * seg, cnt);
		}
	}
	if (!__all_sync(ZG_FULL, ok)) return ZS_E_CORRUPT;
	__syncwarp();
	lit = litbuf;
	return ZS_OK;
}

// ---------------------------------------------------------------------------------------------
// Phase A for one Compressed block of the frame viewed by U (uniform in the warp).  Returns 0 = ok
// (U.nseq sequences wait for phase B; 0 means the block was literals only and is complete),
// ZD_DEFER = the sequence arena is full (nothing consumed), else a ZS_E_* code + 1.
#define ZD_DEFER 1u
#define ZD_ERR(c) ((c) + 1u)
ZG_DEV u32 zd_block_setup(ZdWarp* W, ZdLane& U, const u8* src, u32 n, bool last, u32* slot, u32& arena_used, u8* litbuf, u8* hufsave) {
	ZdLitHdr h;
	if (!zd_lit_header(src, n, h)) return ZD_ERR(ZS_E_CORRUPT);
	const u8* end = src + n;
	const u8* p = src + h.hdr + h.comp;
	// ---- sequences section header ----
	if (p >= end) return ZD_ERR(ZS_E_CORRUPT);
	u32 nseq;
	{
		u32 c0 = p[0];
		if (c0 < 128) {
			nseq = c0;
			p += 1;
		} else if (c0 < 255) {
			if (end - p < 2) return ZD_ERR(ZS_E_CORRUPT);
			nseq = ((c0 - 128) << 8) + p[1];
			p += 2;
		} else {
			if (end - p < 3) return ZD_ERR(ZS_E_CORRUPT);
			nseq = (u32)p[1] + ((u32)p[2] << 8) + 0x7F00;
			p += 3;
		}
	}
	if (nseq == 0) {
		if (p != end) return ZD_ERR(ZS_E_CORRUPT);
		const u8* lit;
		bool lit_rle;
		u32 rle_byte;
		u32 r = zd_literals(W, U.flags, src, h, last, litbuf, hufsave, lit, lit_rle, rle_byte);
		if (r) return ZD_ERR(r);
		if (U.opos + h.regen > U.cap) return ZD_ERR(ZS_E_DST_SMALL);
		if (h.regen) {
			if (lit_rle) zg_warp_fill(U.out + U.opos, rle_byte, h.regen);
			else zg_warp_copy(U.out + U.opos, lit, h.regen);
		}
		U.opos += h.regen;
		U.nseq = U.nseq_left = 0;
		__syncwarp();
		return 0;
	}
	if (nseq > ZD_SEQ_ARENA) return ZD_ERR(ZS_E_CORRUPT);  // > 131072 sequences cannot come from a 128 KiB block
	if (arena_used + nseq > ZD_SEQ_ARENA) return ZD_DEFER;    // (an empty arena always fits a block)
	if (p >= end) return ZD_ERR(ZS_E_CORRUPT);
	u32 modes = *p++;
	if (modes & 3) return ZD_ERR(ZS_E_CORRUPT);
	if (!zd_seq_table(W, slot, U.ll_log, U.flags, ZD_F_LL_OK, ZD_F_LL_DEF, (modes >> 6) & 3, p, end, ZS_LL_MAXLOG, 35, 6)) return ZD_ERR(ZS_E_CORRUPT);
	if (!zd_seq_table(W, slot + 1024, U.of_log, U.flags, ZD_F_OF_OK, ZD_F_OF_DEF, (modes >> 4) & 3, p, end, ZS_OF_MAXLOG, 31, 5)) return ZD_ERR(ZS_E_CORRUPT);
	if (!zd_seq_table(W, slot + 512, U.ml_log, U.flags, ZD_F_ML_OK, ZD_F_ML_DEF, (modes >> 2) & 3, p, end, ZS_ML_MAXLOG, 52, 6)) return ZD_ERR(ZS_E_CORRUPT);
	// bitstream + the three initial states (uniform here; the frame's lane keeps them)
	if (!zs_back_init(U.b, p, (u32)(end - p))) return ZD_ERR(ZS_E_CORRUPT);
	zs_back_reload(U.b);
	U.sl = zs_back_read(U.b, U.ll_log);
	U.so = zs_back_read(U.b, U.of_log);
	U.sm = zs_back_read(U.b, U.ml_log);
	if (zs_back_overflow(U.b)) return ZD_ERR(ZS_E_CORRUPT);
	U.nseq = U.nseq_left = nseq;
	U.seq_base = arena_used;
	U.blk = src;
	U.blk_n = n;
	arena_used += nseq;
	__syncwarp();
	return 0;
}

// Advance the frame viewed by U (uniform) through its blocks until one with sequences is ready for
// phase B, the frame ends, fails, or has to wait for arena space.
ZG_DEV void zd_setup_frame(ZdWarp* W, ZdLane& U, u32* slot, u32& arena_used, u8* litbuf, u8* hufsave) {
	for (;;) {
		if (U.ip + 3 > U.n) return zd_fail(U, ZS_E_SRC_SIZE);
		u32 bh = zg_ld24(U.src + U.ip);
		u32 last = bh & 1, type = (bh >> 1) & 3, bsize = bh >> 3;
		if (type == 3) return zd_fail(U, ZS_E_CORRUPT);
		u64 body = U.ip + 3;
		if (type == 0) {
			if (body + bsize > U.n) return zd_fail(U, ZS_E_SRC_SIZE);
			if (bsize > ZS_BLOCK_MAX) return zd_fail(U, ZS_E_CORRUPT);
			if (U.opos + bsize > U.cap) return zd_fail(U, ZS_E_DST_SMALL);
			zg_warp_copy(U.out + U.opos, U.src + body, bsize);
			U.opos += bsize;
			U.ip = body + bsize;
		} else if (type == 1) {
			if (body + 1 > U.n) return zd_fail(U, ZS_E_SRC_SIZE);
			if (bsize > ZS_BLOCK_MAX) return zd_fail(U, ZS_E_CORRUPT);
			if (U.opos + bsize > U.cap) return zd_fail(U, ZS_E_DST_SMALL);
			zg_warp_fill(U.out + U.opos, U.src[body], bsize);
			U.opos += bsize;
			U.ip = body + 1;
		} else {
			if (bsize > ZS_BLOCK_MAX) return zd_fail(U, ZS_E_CORRUPT);
			if (body + bsize > U.n) return zd_fail(U, ZS_E_SRC_SIZE);
			u32 r = zd_block_setup(W, U, U.src + body, bsize, last != 0, slot, arena_used, litbuf, hufsave);
			if (r == ZD_DEFER) return;
			if (r) return zd_fail(U, r - 1u);
			U.ip = body + bsize;
			if (U.nseq) {
				U.flags = last ? (U.flags | ZD_F_LAST) : (U.flags & ~ZD_F_LAST);
				return;
			}
		}
		__syncwarp();
		if (last) return zd_frame_finish(U);
	}
}

// ---------------------------------------------------------------------------------------------
// Phase B, lane-private: decode up to `cnt` sequences of this lane's block and append them, packed,
// to dst.  The caller keeps all lanes in step (cnt is bounded) so that the warp stays converged.
ZG_DEV bool zd_lane_decode(ZdLane& L, u32 cnt, const u32* llt, const u32* mlt, const u32* oft, u64* dst) {
	ZsBack b = L.b;
	u32 sl = L.sl, so = L.so, sm = L.sm;
	u32 rep0 = L.rep0, rep1 = L.rep1, rep2 = L.rep2;
	u32 left = L.nseq_left;
	ZsBelow ahead = {0, 0};
	zs_below_fetch(b, ahead);
	u32 bad = 0;  // checked once per call: a bad sequence only poisons values, never an address (states stay inside their tables)
	for (u32 k = 0; k < cnt; k++) {
		u32 oe = oft[so], me = mlt[sm], le = llt[sl];
		zs_back_reload_ahead(b, ahead);  // >= 57 bits available from here; no memory wait (the bytes came in a step ago)
		// the codes come out of tables whose symbols were range-checked when they were read (zd_seq_table)
		u32 oc = oe & 0xff;
		u32 mp = ZS_ML_PACK[me & 0xff], lp = ZS_LL_PACK[le & 0xff];
		u32 ofv = (1u << oc) + zs_back_read(b, oc);
		u32 used = oc;
		if (oc > 24) {
			zs_back_reload_ahead(b, ahead);
			used = 0;
		}
		u32 mb = mp >> 24, lb = lp >> 24;
		u32 ml = (mp & 0xffffffu) + zs_back_read(b, mb);
		u32 ll = (lp & 0xffffffu) + zs_back_read(b, lb);
		used += mb + lb;
		if (left - k > 1) {
			if (used > 30) zs_back_reload_ahead(b, ahead);  // the three state updates need <= 26 bits
			sl = (le >> 16) + zs_back_read(b, (le >> 8) & 0xff);
			sm = (me >> 16) + zs_back_read(b, (me >> 8) & 0xff);
			so = (oe >> 16) + zs_back_read(b, (oe >> 8) & 0xff);
		}
		// repeat-offset resolution (RFC 8878 §3.1.1.5)
		u32 off;
		if (ofv > 3) {
			off = ofv - 3;
			rep2 = rep1;
			rep1 = rep0;
			rep0 = off;
		} else {
			u32 idx = ofv - 1 + (ll == 0 ? 1 : 0);
			off = idx == 0 ? rep0 : idx == 1 ? rep1 : idx == 2 ? rep2 : rep0 - 1u;
			if (idx > 1) rep2 = rep1;
			if (idx > 0) {
				rep1 = rep0;
				rep0 = off;
			}
		}
		// offsets beyond 2^28 exceed every window this decoder accepts; an offset of 0 only arises from a corrupt stream
		// or from the unknown repeat-offset history of a split frame's block (then the block is not independent);
		// lengths are < 2^18 by construction
		bad |= (off - 1u >= ZD_OFF_MAX) ? 1u : 0u;
		dst[k] = (u64)off | ((u64)ll << 28) | ((u64)ml << 46);
	}
	if (bad || zs_back_overflow(b)) return false;
	left -= cnt;
	if (left == 0) {
		zs_back_reload(b);
		if (!zs_back_finished(b)) return false;
	}
	L.b = b;
	L.sl = sl;
	L.so = so;
	L.sm = sm;
	L.rep0 = rep0;
	L.rep1 = rep1;
	L.rep2 = rep2;
	L.nseq_left = left;
	return true;
}

// Per-lane copy of n bytes between regions that do not overlap: 16 bytes per memory round trip (the
// loads of a group are all in flight before the first store).  Only aligned words that hold at least
// one source byte are touched.
template <bool CG>
ZG_DEV void zd_lane_copy(u8* dst, const u8* src, u32 n) {
	const u8* lim = src + n;
	for (u32 k0 = 0; k0 < n; k0 += 16) {
		uintptr_t a = (uintptr_t)(src + k0);
		const u32* w = (const u32*)(a & ~(uintptr_t)3);
		u32 sh = (u32)(a & 3) * 8;
		u32 x[5];
		ZG_UNROLL
		for (int k = 0; k < 5; k++) x[k] = (const u8*)(w + k) < lim ? zd_ldw<CG>(w + k) : 0u;
		u32 m = n - k0;
		u8* q = dst + k0;
		ZG_UNROLL
		for (int j = 0; j < 4; j++) {
			if (4u * j < m) {
				u32 v = __funnelshift_r(x[j], x[j + 1], sh);
				q[4 * j] = (u8)v;
				if (4u * j + 1 < m) q[4 * j + 1] = (u8)(v >> 8);
				if (4u * j + 2 < m) q[4 * j + 2] = (u8)(v >> 16);
				if (4u * j + 3 < m) q[4 * j + 3] = (u8)(v >> 24);
			}
		}
	}
}
// Per-lane match copy d[i] = d[i - off] for off < ml (the match overlaps its own output): the valid
// span [d - off, d + done) is periodic in `off`, so it is extended by non-overlapping copies whose
// length doubles.
template <bool CG>
ZG_DEV void zd_lane_overlap(u8* d, u32 off, u32 ml) {
	u32 done = 0;
	while (done < ml) {
		u32 c = zg_min<u32>(done + off, ml - done);
		zd_lane_copy<CG>(d + done, d - off, c);
		done += c;
	}
}

#ifndef ZD_LANE_COPY_MAX
#define ZD_LANE_COPY_MAX 64u   // longer literal runs / matches are copied by the whole warp
#endif

// Phase C: execute `cnt` (<= 32) sequences, one per lane.  Output and literal positions come from
// warp scans.  All literal runs are independent and copied first, lane-parallel.  Matches then go in
// waves: a match is ready once every earlier match of the row whose output its source overlaps has been
// written (match starts and ends grow with the lane, so that set is a range of lanes, found by two binary
// searches over shuffles); all ready matches of a wave copy lane-parallel.  A row of n sequences takes (dependency depth) waves, not n
// steps, and every wave is a few 16-byte round trips.
// Where a row's bytes live.  ZdDirect: the frame's output in global memory (CG: match sources are read through L2, for
// bytes another SM may have written during the kernel).  ZdWindow (zstd_decode_staged.cu, the chain executor): the recent
// output in a shared-memory window, older bytes in global memory.
template <bool CG>
struct ZdDirect {
	u8* out;
	ZG_DEV u8* dst(u64 pos) const { return out + pos; }
	ZG_DEV void lane_copy(u64 dpos, u64 spos, u32 n) const { zd_lane_copy<CG>(out + dpos, out + spos, n); }
	ZG_DEV void lane_overlap(u64 dpos, u32 off, u32 ml) const { zd_lane_overlap<CG>(out + dpos, off, ml); }
	ZG_DEV void warp_copy(u64 dpos, u64 spos, u32 n) const { zg_warp_copy_t<CG>(out + dpos, out + spos, n); }
	ZG_DEV void warp_match(u64 dpos, u32 off, u32 ml) const { zd_warp_match<CG>(out + dpos, off, ml); }
};

template <class POL>
ZG_DEV u32 zd_exec_row(u64 sq_lane, u32 cnt, const POL& P, u64& o_io, u64 base, u64 cap, const u8* lit, bool lit_rle, u32 rle_byte,
                       u32 regen, u32& lpos_io) {
	u32 lane = zg_lane();
	u64 o = o_io;
	u32 lpos = lpos_io;
	bool act = lane < cnt;
	u64 sq = act ? sq_lane : 0;
	u32 of = (u32)sq & ZD_OFF_MAX, ll = (u32)(sq >> 28) & 0x3ffffu, ml = (u32)(sq >> 46);
	u32 incl = zg_warp_incl_scan(ll + ml), lincl = zg_warp_incl_scan(ll);
	u32 btot = __shfl_sync(ZG_FULL, incl, 31), ltot = __shfl_sync(ZG_FULL, lincl, 31);
	if (lpos + ltot > regen) return ZS_E_CORRUPT;
	if (o + btot > cap) return ZS_E_DST_SMALL;
	u32 rstart = incl - ml;          // my match's start relative to the row's output start
	u64 mstart = o + rstart;         // ... and in the frame output
	if (__any_sync(ZG_FULL, act && (u64)of > mstart - base)) return ZS_E_CORRUPT;
	u8* d = P.dst(mstart - ll);
	u32 lsrc = lpos + (lincl - ll);
	// (a) literals
	u32 longl = __ballot_sync(ZG_FULL, ll > ZD_LANE_COPY_MAX);
	if (ll <= ZD_LANE_COPY_MAX) {
		if (lit_rle) for (u32 k = 0; k < ll; k++) d[k] = (u8)rle_byte;
		else zd_lane_copy<false>(d, lit + lsrc, ll);
	}
	while (longl) {
		int l = __ffs((int)longl) - 1;
		longl &= longl - 1;
		u32 n2 = __shfl_sync(ZG_FULL, ll, l), s2 = __shfl_sync(ZG_FULL, lsrc, l);
		u64 d2 = __shfl_sync(ZG_FULL, (u64)(uintptr_t)d, l);
		if (lit_rle) zg_warp_fill((u8*)(uintptr_t)d2, rle_byte, n2);
		else zg_warp_copy((u8*)(uintptr_t)d2, lit + s2, n2);
	}
	// (b) matches.  need = the lanes below me whose match starts before the end of my source
	//     (rstart grows with the lane: binary search by shuffles; idle lanes sit at the far end)
	i32 rs = act ? (i32)rstart : 0x7fffffff;
	i32 send = (i32)rstart - (i32)of + (i32)zg_min<u32>(ml, of);  // <= 0: the source lies before this row
	u32 nlow = 0;
	ZG_UNROLL
	for (int s = 16; s >= 1; s >>= 1) {
		i32 v = __shfl_sync(ZG_FULL, rs, (int)(nlow + (u32)s - 1u));
		if (v < send) nlow += (u32)s;
	}
	nlow = zg_min<u32>(nlow, lane);
	// ... and of those, not the ones whose match ends at or before my source begins (match ends grow with the
	// lane too): what is left is exactly the lanes whose output my source overlaps.  Waiting for the whole prefix
	// instead would push every later match behind the deepest chain seen so far.
	i32 re = act ? (i32)(rstart + ml) : 0x7fffffff;
	i32 sbeg = (i32)rstart - (i32)of;
	u32 nlo = 0;
	ZG_UNROLL
	for (int s = 16; s >= 1; s >>= 1) {
		i32 v = __shfl_sync(ZG_FULL, re, (int)(nlo + (u32)s - 1u));
		if (v <= sbeg) nlo += (u32)s;
	}
	nlo = zg_min<u32>(nlo, nlow);
	u32 need = ((1u << nlow) - 1u) & ~((1u << nlo) - 1u);
	u32 pending = __ballot_sync(ZG_FULL, act && ml > 0);
	__syncwarp();  // the literals are written
	while (pending) {
		bool ready = ((pending >> lane) & 1u) && !(pending & need);
		bool shortm = ml <= ZD_LANE_COPY_MAX;
		u32 rdy = __ballot_sync(ZG_FULL, ready);
		u32 longm = __ballot_sync(ZG_FULL, ready && !shortm);
		if (ready && shortm) {
			if (of >= ml) P.lane_copy(mstart, mstart - of, ml);
			else P.lane_overlap(mstart, of, ml);
		}
		while (longm) {
			int l = __ffs((int)longm) - 1;
			longm &= longm - 1;
			u32 n2 = __shfl_sync(ZG_FULL, ml, l), o2 = __shfl_sync(ZG_FULL, of, l);
			u64 m2 = __shfl_sync(ZG_FULL, mstart, l);
			if (o2 >= n2) P.warp_copy(m2, m2 - o2, n2);
			else P.warp_match(m2, o2, n2);
		}
		__syncwarp();
		pending &= ~rdy;
	}
	o_io = o + btot;
	lpos_io = lpos + ltot;
	return ZS_OK;
}

// Phase C for the block in flight of the frame viewed by U (uniform): literals, then all sequences.
ZG_DEV void zd_exec_block(ZdWarp* W, ZdLane& U, const u64* seqs, u8* litbuf, u8* hufsave) {
	ZdLitHdr h;
	zd_lit_header(U.blk, U.blk_n, h);  // validated in phase A
	const u8* lit;
	bool lit_rle;
	u32 rle_byte;
	u32 r = zd_literals(W, U.flags, U.blk, h, (U.flags & ZD_F_LAST) != 0, litbuf, hufsave, lit, lit_rle, rle_byte);
	if (r) return zd_fail(U, r);
	u64 o = U.opos;
	u32 lpos = 0;
	u32 lane = zg_lane();
	u64 sq = lane < U.nseq ? seqs[lane] : 0;
	for (u32 s0 = 0; s0 < U.nseq; s0 += 32) {
		u64 sq_next = s0 + 32 + lane < U.nseq ? seqs[s0 + 32 + lane] : 0;  // the next row's records come in while this row executes
		r = zd_exec_row(sq, zg_min<u32>(32u, U.nseq - s0), ZdDirect<false>{U.out}, o, U.base, U.cap, lit, lit_rle, rle_byte, h.regen, lpos);
		if (r) return zd_fail(U, r);
		sq = sq_next;
	}
	// the literals after the last sequence
	u32 rest = h.regen - lpos;
	if (o + rest > U.cap) return zd_fail(U, ZS_E_DST_SMALL);
	if (rest) {
		if (lit_rle) zg_warp_fill(U.out + o, rle_byte, rest);
		else zg_warp_copy(U.out + o, lit + lpos, rest);
		o += rest;
	}
	if (o - U.opos > ZS_BLOCK_MAX) return zd_fail(U, ZS_E_CORRUPT);
	__syncwarp();
	U.opos = o;
	U.nseq = 0;
	if (U.flags & ZD_F_LAST) zd_frame_finish(U);
}

