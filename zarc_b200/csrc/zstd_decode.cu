// K5 + K6: Zstandard frame decode (RFC 8878).  Replaces DCtx::decompress_stream as driven by
// crates/zarc/src/decode/zstd_iterator.rs:88-153 (one frame per content entry, decode/frame_iterator.rs).
//
// One warp per frame; blocks of a frame are decoded in order by that warp, so Treeless literals,
// Repeat_Mode tables, repeat offsets and cross-block matches (all produced by libzstd at levels
// 1/3/9) come for free.  Within a block: Huffman tree + FSE tables are built
// warp-cooperatively in shared memory, the 4 literal streams are decoded by 4 lanes, the sequence
// bitstream by lane 0 in batches of 32, and each batch is executed by the whole warp (byte-parallel
// copies straight into the frame's output in HBM).  Frames are pulled from an atomic queue.
// HBM traffic: C read + N written (+ literals staged through an L2-resident per-warp buffer).
#include "common.h"
#include "zstd_common.cuh"

#define ZD_WARPS 4
#define ZD_LITBUF (ZS_BLOCK_MAX + 64)

struct ZdWarp {
	u32 ll_tab[512];
	u32 ml_tab[512];
	u32 of_tab[256];
	u32 wtab[64];      // FSE table of the Huffman weights
	u16 huf[2048];     // sym | nbBits << 8
	u32 seq_ll[32], seq_ml[32], seq_of[32];
	u8 weights[256];
	i16 norm[256];
	u16 next[256];
	u32 misc[8];
};

struct ZdState {
	u32 rep0, rep1, rep2;
	u32 ll_log, ml_log, of_log, huf_bits;
	bool huf_ok, ll_ok, ml_ok, of_ok;
};

// ---------------------------------------------------------------------------------------------
// warp-cooperative copies
ZG_DEV void zg_warp_copy(u8* dst, const u8* src, u32 n) {
	u32 lane = zg_lane();
	if (n < 128) {
		for (u32 i = lane; i < n; i += 32) dst[i] = src[i];
		return;
	}
	u32 head = (u32)((16 - ((uintptr_t)dst & 15)) & 15);
	if (lane < head) dst[lane] = src[lane];
	dst += head;
	src += head;
	n -= head;
	u32 nvec = n >> 4;
	uintptr_t sa = (uintptr_t)src;
	const u32* sw = (const u32*)(sa & ~(uintptr_t)3);
	u32 sh = (u32)(sa & 3) * 8;
	for (u32 v = lane; v < nvec; v += 32) {
		const u32* q = sw + 4 * v;
		u32 w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3];
		uint4 o;
		if (sh == 0) {
			o = make_uint4(w0, w1, w2, w3);
		} else {
			u32 w4 = q[4];
			o = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
			               __funnelshift_r(w3, w4, sh));
		}
		((uint4*)dst)[v] = o;
	}
	for (u32 i = (nvec << 4) + lane; i < n; i += 32) dst[i] = src[i];
}
ZG_DEV void zg_warp_fill(u8* dst, u32 byte, u32 n) {
	for (u32 i = zg_lane(); i < n; i += 32) dst[i] = (u8)byte;
}
// match copy inside the frame output: d[i] = d[i - off], forward semantics (overlap allowed)
ZG_DEV void zd_warp_match(u8* d, u32 off, u32 ml) {
	u32 lane = zg_lane();
	const u8* s = d - off;
	if (off >= 32) {
		bool overlap = off < ml;
		for (u32 i0 = 0; i0 < ml; i0 += 32) {
			u32 i = i0 + lane;
			if (i < ml) d[i] = s[i];
			if (overlap) __syncwarp();
		}
	} else {
		for (u32 i = lane; i < ml; i += 32) d[i] = s[i % off];
	}
}

// ---------------------------------------------------------------------------------------------
// Huffman tree description -> W->huf.  Returns bytes consumed (0 on error).  All lanes call.
ZG_DEV u32 zd_read_huf_tree(ZdWarp* W, ZdState& st, const u8* src, u32 n) {
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
		zs_fse_build_dtable(W->wtab, W->norm, nsym, log, W->next);
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
					u32 e1 = W->wtab[s1];
					W->weights[cnt++] = (u8)e1;
					zs_back_reload(b);
					s1 = (e1 >> 16) + zs_back_read(b, (e1 >> 8) & 0xff);
					if (zs_back_overflow(b)) {
						W->weights[cnt++] = (u8)W->wtab[s2];
						break;
					}
					u32 e2 = W->wtab[s2];
					W->weights[cnt++] = (u8)e2;
					s2 = (e2 >> 16) + zs_back_read(b, (e2 >> 8) & 0xff);
					if (zs_back_overflow(b)) {
						W->weights[cnt++] = (u8)W->wtab[s1];
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
	st.huf_bits = maxbits;
	st.huf_ok = true;
	return used;
}

// one Huffman stream, single thread
ZG_DEV bool zd_huf_stream(const u16* huf, u32 maxbits, const u8* src, u32 n, u8* dst, u32 count) {
	ZsBack b;
	if (!zs_back_init(b, src, n)) return false;
	u32 i = 0;
	while (i < count) {
		zs_back_reload(b);
		u32 m = zg_min<u32>(count - i, 5u);  // 5 x 11 bits <= 57
		for (u32 k = 0; k < m; k++) {
			u32 e = huf[zs_back_look(b, maxbits)];
			dst[i + k] = (u8)e;
			zs_back_skip(b, e >> 8);
		}
		i += m;
	}
	zs_back_reload(b);
	return zs_back_finished(b);
}

// sequence table for one of LL/OF/ML.  All lanes call.  Returns false on error; advances *pp.
ZG_DEV bool zd_seq_table(ZdWarp* W, u32* tab, u32& log, bool& have, u32 mode, const u8*& p, const u8* end, u32 maxlog,
                         u32 maxsym, const u32* def_tab, u32 def_log) {
	u32 lane = zg_lane();
	if (mode == 0) {
		for (u32 i = lane; i < (1u << def_log); i += 32) tab[i] = def_tab[i];
		__syncwarp();
		log = def_log;
		have = true;
		return true;
	}
	if (mode == 1) {
		if (p >= end) return false;
		u32 sym = *p++;
		if (sym > maxsym) return false;
		if (lane == 0) tab[0] = sym;
		__syncwarp();
		log = 0;
		have = true;
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
		zs_fse_build_dtable(tab, W->norm, nsym, lg, W->next);
		p += nc;
		log = lg;
		have = true;
		return true;
	}
	return have;  // Repeat_Mode
}

// ---------------------------------------------------------------------------------------------
// one Compressed block.  `out` = start of the frame's output, opos = bytes produced so far.
ZG_DEV u32 zd_compressed_block(ZdWarp* W, ZdState& st, const u8* src, u32 n, u8* out, u64& opos, u64 cap, u8* litbuf) {
	u32 lane = zg_lane();
	if (n < 2) return ZS_E_CORRUPT;
	const u8* end = src + n;
	u32 b0 = src[0];
	u32 ltype = b0 & 3, sf = (b0 >> 2) & 3;
	u32 regen = 0, comp = 0, hdr = 0, streams = 1;
	const u8* lit = litbuf;
	bool lit_rle = false;
	u32 rle_byte = 0;
	const u8* p;
	if (ltype < 2) {
		if (sf == 0 || sf == 2) {
			regen = b0 >> 3;
			hdr = 1;
		} else if (sf == 1) {
			regen = (b0 >> 4) | ((u32)src[1] << 4);
			hdr = 2;
		} else {
			if (n < 3) return ZS_E_CORRUPT;
			regen = (b0 >> 4) | ((u32)src[1] << 4) | ((u32)src[2] << 12);
			hdr = 3;
		}
		if (regen > ZS_BLOCK_MAX) return ZS_E_CORRUPT;
		p = src + hdr;
		if (ltype == 0) {
			if (hdr + regen > n) return ZS_E_CORRUPT;
			lit = p;
			p += regen;
		} else {
			if (hdr + 1 > n) return ZS_E_CORRUPT;
			lit_rle = true;
			rle_byte = *p++;
		}
	} else {
		if (n < 3) return ZS_E_CORRUPT;
		u64 v = (u64)b0 | ((u64)src[1] << 8) | ((u64)src[2] << 16);
		if (sf == 0 || sf == 1) {
			regen = (u32)(v >> 4) & 1023;
			comp = (u32)(v >> 14) & 1023;
			hdr = 3;
			streams = sf == 0 ? 1 : 4;
		} else if (sf == 2) {
			if (n < 4) return ZS_E_CORRUPT;
			v |= (u64)src[3] << 24;
			regen = (u32)(v >> 4) & 16383;
			comp = (u32)(v >> 18) & 16383;
			hdr = 4;
			streams = 4;
		} else {
			if (n < 5) return ZS_E_CORRUPT;
			v |= ((u64)src[3] << 24) | ((u64)src[4] << 32);
			regen = (u32)(v >> 4) & 262143;
			comp = (u32)(v >> 22) & 262143;
			hdr = 5;
			streams = 4;
		}
		if (regen > ZS_BLOCK_MAX || hdr + comp > n || regen == 0) return ZS_E_CORRUPT;
		const u8* lp = src + hdr;
		const u8* lend = lp + comp;
		if (ltype == 2) {
			u32 used = zd_read_huf_tree(W, st, lp, comp);
			if (used == 0) return ZS_E_CORRUPT;
			lp += used;
		} else if (!st.huf_ok) {
			return ZS_E_CORRUPT;
		}
		bool ok = true;
		if (streams == 1) {
			if (lane == 0) ok = zd_huf_stream(W->huf, st.huf_bits, lp, (u32)(lend - lp), litbuf, regen);
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
				ok = zd_huf_stream(W->huf, st.huf_bits, lp + so, sn, litbuf + lane * seg, cnt);
			}
		}
		if (!__all_sync(ZG_FULL, ok)) return ZS_E_CORRUPT;
		p = lend;
	}
	__syncwarp();
	// ---- sequences section ----
	if (p >= end) return ZS_E_CORRUPT;
	u32 nseq;
	{
		u32 c0 = p[0];
		if (c0 < 128) {
			nseq = c0;
			p += 1;
		} else if (c0 < 255) {
			if (end - p < 2) return ZS_E_CORRUPT;
			nseq = ((c0 - 128) << 8) + p[1];
			p += 2;
		} else {
			if (end - p < 3) return ZS_E_CORRUPT;
			nseq = (u32)p[1] + ((u32)p[2] << 8) + 0x7F00;
			p += 3;
		}
	}
	u32 lpos = 0;
	u64 o = opos;
	if (nseq) {
		if (p >= end) return ZS_E_CORRUPT;
		u32 modes = *p++;
		if (modes & 3) return ZS_E_CORRUPT;
		if (!zd_seq_table(W, W->ll_tab, st.ll_log, st.ll_ok, (modes >> 6) & 3, p, end, ZS_LL_MAXLOG, 35, ZS_LL_DEFAULT_DTABLE, 6)) return ZS_E_CORRUPT;
		if (!zd_seq_table(W, W->of_tab, st.of_log, st.of_ok, (modes >> 4) & 3, p, end, ZS_OF_MAXLOG, 31, ZS_OF_DEFAULT_DTABLE, 5)) return ZS_E_CORRUPT;
		if (!zd_seq_table(W, W->ml_tab, st.ml_log, st.ml_ok, (modes >> 2) & 3, p, end, ZS_ML_MAXLOG, 52, ZS_ML_DEFAULT_DTABLE, 6)) return ZS_E_CORRUPT;
		// lane 0 owns the bitstream, the three states and the repeat offsets
		ZsBack b;
		u32 sl = 0, so = 0, sm = 0;
		u32 rep0 = st.rep0, rep1 = st.rep1, rep2 = st.rep2;
		bool ok = true;
		if (lane == 0) {
			ok = zs_back_init(b, p, (u32)(end - p));
			if (ok) {
				zs_back_reload(b);
				sl = zs_back_read(b, st.ll_log);
				so = zs_back_read(b, st.of_log);
				sm = zs_back_read(b, st.ml_log);
				ok = !zs_back_overflow(b);
			}
		}
		if (!__all_sync(ZG_FULL, ok)) return ZS_E_CORRUPT;
		for (u32 s0 = 0; s0 < nseq; s0 += 32) {
			u32 cnt = zg_min<u32>(32u, nseq - s0);
			if (lane == 0) {
				for (u32 k = 0; k < cnt; k++) {
					u32 oe = W->of_tab[so], me = W->ml_tab[sm], le = W->ll_tab[sl];
					u32 oc = oe & 0xff, mc = me & 0xff, lc = le & 0xff;
					if (oc > 31 || mc > 52 || lc > 35) {
						ok = false;
						break;
					}
					zs_back_reload(b);  // >= 57 bits available from here
					u32 ofv = (1u << oc) + zs_back_read(b, oc);
					u32 used = oc;
					if (oc > 24) {
						zs_back_reload(b);
						used = 0;
					}
					u32 mb = ZS_ML_BITS[mc], lb = ZS_LL_BITS[lc];
					u32 ml = ZS_ML_BASE[mc] + zs_back_read(b, mb);
					u32 ll = ZS_LL_BASE[lc] + zs_back_read(b, lb);
					used += mb + lb;
					if (s0 + k + 1 < nseq) {
						if (used > 30) zs_back_reload(b);  // the three state updates need <= 26 bits
						sl = (le >> 16) + zs_back_read(b, (le >> 8) & 0xff);
						sm = (me >> 16) + zs_back_read(b, (me >> 8) & 0xff);
						so = (oe >> 16) + zs_back_read(b, (oe >> 8) & 0xff);
					}
					if (zs_back_overflow(b)) {
						ok = false;
						break;
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
						if (idx == 0) {
							off = rep0;
						} else {
							off = idx == 1 ? rep1 : idx == 2 ? rep2 : rep0 - 1;
							if (off == 0) {
								ok = false;
								break;
							}
							if (idx > 1) rep2 = rep1;
							rep1 = rep0;
							rep0 = off;
						}
					}
					W->seq_ll[k] = ll;
					W->seq_ml[k] = ml;
					W->seq_of[k] = off;
				}
				if (ok && s0 + cnt == nseq) {
					zs_back_reload(b);
					ok = zs_back_finished(b);
				}
			}
			if (!__all_sync(ZG_FULL, ok)) return ZS_E_CORRUPT;
			__syncwarp();  // lane 0's batch in shared memory is visible to the warp
			// execute the batch: one sequence per lane.  Output and literal positions come from warp
			// scans; (a) all literal runs and (b) all matches whose source lies wholly before this
			// batch's output are independent and copied lane-parallel; (c) the remaining matches
			// (sources inside the batch, incl. overlapping ones) go in order, warp-cooperatively.
			{
				bool act = lane < cnt;
				u32 ll = act ? W->seq_ll[lane] : 0, ml = act ? W->seq_ml[lane] : 0, of = act ? W->seq_of[lane] : 0;
				u32 incl = zg_warp_incl_scan(ll + ml), lincl = zg_warp_incl_scan(ll);
				u32 btot = __shfl_sync(ZG_FULL, incl, 31), ltot = __shfl_sync(ZG_FULL, lincl, 31);
				if (lpos + ltot > regen) return ZS_E_CORRUPT;
				if (o + btot > cap) return ZS_E_DST_SMALL;
				u64 mstart = o + (incl - ll - ml) + ll;  // where my match begins in the frame output
				if (__any_sync(ZG_FULL, act && (u64)of > mstart)) return ZS_E_CORRUPT;
				u8* d = out + mstart - ll;
				u32 lsrc = lpos + (lincl - ll);
				// (a) literals
				u32 longl = __ballot_sync(ZG_FULL, ll >= 48);
				if (ll < 48) {
					if (lit_rle) for (u32 k = 0; k < ll; k++) d[k] = (u8)rle_byte;
					else for (u32 k = 0; k < ll; k++) d[k] = lit[lsrc + k];
				}
				while (longl) {
					int l = __ffs((int)longl) - 1;
					longl &= longl - 1;
					u32 n2 = __shfl_sync(ZG_FULL, ll, l), s2 = __shfl_sync(ZG_FULL, lsrc, l);
					u64 d2 = __shfl_sync(ZG_FULL, (u64)(uintptr_t)d, l);
					if (lit_rle) zg_warp_fill((u8*)(uintptr_t)d2, rle_byte, n2);
					else zg_warp_copy((u8*)(uintptr_t)d2, lit + s2, n2);
				}
				// (b) independent matches
				u8* md = out + mstart;
				bool indep = act && ml > 0 && mstart - of + ml <= o;
				u32 longm = __ballot_sync(ZG_FULL, indep && ml >= 48);
				if (indep && ml < 48) {
					const u8* ms = md - of;
					for (u32 k = 0; k < ml; k++) md[k] = ms[k];
				}
				while (longm) {
					int l = __ffs((int)longm) - 1;
					longm &= longm - 1;
					u32 n2 = __shfl_sync(ZG_FULL, ml, l), o2 = __shfl_sync(ZG_FULL, of, l);
					u64 d2 = __shfl_sync(ZG_FULL, (u64)(uintptr_t)md, l);
					zg_warp_copy((u8*)(uintptr_t)d2, (const u8*)(uintptr_t)d2 - o2, n2);
				}
				__syncwarp();
				// (c) dependent matches, in sequence order
				u32 dep = __ballot_sync(ZG_FULL, act && ml > 0 && !indep);
				while (dep) {
					int l = __ffs((int)dep) - 1;
					dep &= dep - 1;
					u32 n2 = __shfl_sync(ZG_FULL, ml, l), o2 = __shfl_sync(ZG_FULL, of, l);
					u64 d2 = __shfl_sync(ZG_FULL, (u64)(uintptr_t)md, l);
					zd_warp_match((u8*)(uintptr_t)d2, o2, n2);
					__syncwarp();
				}
				o += btot;
				lpos += ltot;
			}
			__syncwarp();
		}
		st.rep0 = __shfl_sync(ZG_FULL, rep0, 0);
		st.rep1 = __shfl_sync(ZG_FULL, rep1, 0);
		st.rep2 = __shfl_sync(ZG_FULL, rep2, 0);
	} else if (p != end) {
		return ZS_E_CORRUPT;
	}
	u32 rest = regen - lpos;
	if (o + rest > cap) return ZS_E_DST_SMALL;
	if (rest) {
		if (lit_rle) zg_warp_fill(out + o, rle_byte, rest);
		else zg_warp_copy(out + o, lit + lpos, rest);
		o += rest;
	}
	if (o - opos > ZS_BLOCK_MAX) return ZS_E_CORRUPT;
	__syncwarp();
	opos = o;
	return ZS_OK;
}

// one frame.  Returns status; *produced = bytes written; *cksum = stored checksum (valid if *has_ck).
ZG_DEV u32 zd_frame(ZdWarp* W, const u8* src, u64 n, u8* out, u64 cap, u8* litbuf, u64* produced, u32* cksum, bool* has_ck) {
	*produced = 0;
	*has_ck = false;
	if (n < 6) return ZS_E_SRC_SIZE;
	if (zg_ld32(src) != ZS_MAGIC) return ZS_E_PREFIX;
	u32 desc = src[4];
	u32 fcs_flag = desc >> 6, single = (desc >> 5) & 1, checksum = (desc >> 2) & 1, did_flag = desc & 3;
	if (desc & 8) return ZS_E_UNSUPPORTED;
	u64 ip = 5;
	u64 window = 0;
	if (!single) {
		u32 wd = src[ip++];
		u32 wl = 10 + (wd >> 3);
		if (wl > 31) return ZS_E_WINDOW;
		window = ((u64)1 << wl) + ((((u64)1 << wl) >> 3) * (wd & 7));
	}
	if (did_flag) {
		u32 dl = did_flag == 3 ? 4 : did_flag;
		if (ip + dl > n) return ZS_E_SRC_SIZE;
		u32 did = 0;
		for (u32 i = 0; i < dl; i++) did |= (u32)src[ip + i] << (8 * i);
		ip += dl;
		if (did) return ZS_E_DICT;
	}
	u32 fcs_len = fcs_flag == 0 ? single : fcs_flag == 1 ? 2 : fcs_flag == 2 ? 4 : 8;
	if (ip + fcs_len > n) return ZS_E_SRC_SIZE;
	u64 fcs = 0;
	for (u32 i = 0; i < fcs_len; i++) fcs |= (u64)src[ip + i] << (8 * i);
	if (fcs_len == 2) fcs += 256;
	ip += fcs_len;
	// libzstd's streaming decoder (what zstd_iterator.rs:29 creates) refuses windows > 2^27
	// ("Frame requires too much memory for decoding")
	if ((single ? fcs : window) > ((u64)1 << 27)) return ZS_E_WINDOW;
	ZdState st;
	st.rep0 = 1;
	st.rep1 = 4;
	st.rep2 = 8;
	st.huf_ok = st.ll_ok = st.ml_ok = st.of_ok = false;
	st.ll_log = st.ml_log = st.of_log = st.huf_bits = 0;
	u64 opos = 0;
	for (;;) {
		if (ip + 3 > n) return ZS_E_SRC_SIZE;
		u32 bh = zg_ld24(src + ip);
		ip += 3;
		u32 last = bh & 1, type = (bh >> 1) & 3, bsize = bh >> 3;
		if (type == 3) return ZS_E_CORRUPT;
		if (type == 0) {
			if (ip + bsize > n) return ZS_E_SRC_SIZE;
			if (bsize > ZS_BLOCK_MAX) return ZS_E_CORRUPT;
			if (opos + bsize > cap) return ZS_E_DST_SMALL;
			zg_warp_copy(out + opos, src + ip, bsize);
			opos += bsize;
			ip += bsize;
		} else if (type == 1) {
			if (ip + 1 > n) return ZS_E_SRC_SIZE;
			if (bsize > ZS_BLOCK_MAX) return ZS_E_CORRUPT;
			if (opos + bsize > cap) return ZS_E_DST_SMALL;
			zg_warp_fill(out + opos, src[ip], bsize);
			opos += bsize;
			ip += 1;
		} else {
			if (bsize > ZS_BLOCK_MAX) return ZS_E_CORRUPT;
			if (ip + bsize > n) return ZS_E_SRC_SIZE;
			u32 r = zd_compressed_block(W, st, src + ip, bsize, out, opos, cap, litbuf);
			if (r) return r;
			ip += bsize;
		}
		__syncwarp();
		if (last) break;
	}
	if (checksum) {
		if (ip + 4 > n) return ZS_E_SRC_SIZE;
		*cksum = zg_ld32(src + ip);
		*has_ck = true;
		ip += 4;
	}
	*produced = opos;
	if (fcs_len && fcs != opos) return ZS_E_CORRUPT;
	return ZS_OK;
}

__global__ void __launch_bounds__(ZD_WARPS * 32)
k_zstd_decode_frames(const u8* __restrict__ archive, u64 archive_len, const u64* __restrict__ off, const u64* __restrict__ len,
                     const u64* __restrict__ ulen, const u64* __restrict__ out_off, u64 nframes, u8* out, u64 out_cap,
                     u8* litbufs, u32* queue, u32* status, u64* produced, u32* cksums) {
	__shared__ ZdWarp sm[ZD_WARPS];
	u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	ZdWarp* W = &sm[warp];
	u8* litbuf = litbufs + (size_t)(blockIdx.x * ZD_WARPS + warp) * ZD_LITBUF;
	for (;;) {
		u32 k = 0;
		if (lane == 0) k = atomicAdd(queue, 1u);
		k = __shfl_sync(ZG_FULL, k, 0);
		if (k >= nframes) break;
		u64 fo = off[k], fl = len[k], ul = ulen[k], oo = out_off[k];
		u32 r;
		u64 prod = 0;
		u32 ck = 0;
		bool has = false;
		if (fo > archive_len || fl > archive_len - fo) r = ZS_E_SRC_SIZE;
		else if (oo > out_cap || ul > out_cap - oo) r = ZS_E_DST_SMALL;
		else r = zd_frame(W, archive + fo, fl, out + oo, ul, litbuf, &prod, &ck, &has);
		__syncwarp();
		if (lane == 0) {
			status[k] = r;
			produced[k] = prod;
			cksums[2 * k] = ck;
			cksums[2 * k + 1] = has ? 1u : 0u;
		}
	}
}

size_t zg_zstd_decode_run(cudaStream_t s, ZgZdWork& w, const u8* archive, u64 archive_len, const u64* off, const u64* len,
                          const u64* ulen, const u64* out_off, u64 n, u8* out, u64 out_cap, u32* status, u64* produced,
                          u32* cksums) {
	if (n == 0) return 0;
	u32 grid = (u32)zg_min<u64>((n + ZD_WARPS - 1) / ZD_WARPS, (u64)zg_sm_count() * 5);
	if (w.lit.reserve((size_t)grid * ZD_WARPS * ZD_LITBUF) || w.queue.reserve(16)) return ZG_ERR(ZG_error_memory_allocation);
	cudaMemsetAsync(w.queue.p, 0, 16, s);
	zg_prof_begin(ZG_K_DECODE, s);
	ZG_LAUNCH(k_zstd_decode_frames, grid, ZD_WARPS * 32, 0, s, archive, archive_len, off, len, ulen, out_off, n, out, out_cap,
	          w.lit.as<u8>(), w.queue.as<u32>(), status, produced, cksums);
	zg_prof_end(ZG_K_DECODE, s);
	ZG_COUNT_LAUNCH();
	return cudaGetLastError() == cudaSuccess ? 0 : ZG_ERR(ZG_error_device);
}
