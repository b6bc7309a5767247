// K5 + K6: Zstandard frame decode (RFC 8878).  Replaces DCtx::decompress_stream as driven by
// crates/zarc/src/decode/zstd_iterator.rs:88-153 (one frame per content entry, decode/frame_iterator.rs).
//
// One warp owns a batch of 32 frames, ONE FRAME PER LANE.  Everything that is inherently serial
// inside a frame -- frame/block header parsing, the FSE sequence bitstream with its three states,
// repeat-offset resolution -- runs lane-per-frame, so 32 serial chains advance per warp instruction.
// Everything that is parallel inside a frame is done by the whole warp for one frame at a time:
// Huffman/FSE table construction, the 4 literal streams, and sequence execution (literal and match
// copies straight into the frame's output in HBM).
//
// A batch proceeds in rounds, one block per live frame and round:
//   A. per frame (warp-cooperative): block header (Raw/RLE blocks are copied on the spot), literals
//      and sequences section headers, FSE decode tables into the lane's global slot, bitstream init;
//   B. lock-step (lane-per-frame): every lane decodes ALL sequences of its block, resolves repeat
//      offsets and appends them, packed to 8 bytes (offset:28 | litLength:18 | matchLength:18), to
//      the frame's span of the warp's sequence arena;
//   C. per frame (warp-cooperative): Huffman table + the literal streams into the warp's literal
//      buffer, then the sequences are executed 32 at a time -- so while a frame is being written its
//      literals, its sequences and its recent output (the match sources) are cache-hot.
// Blocks of a frame stay in order, so Treeless literals, Repeat_Mode tables, repeat offsets and
// cross-block matches (all produced by libzstd at levels 1/3/9) work.  State that
// must outlive a round lives in global scratch: the FSE tables (fixed slot per lane) and the Huffman
// weights (for Treeless blocks).  A block whose sequences do not fit the arena waits for the next round.
// HBM traffic: C read + N written (+ 8 B per sequence and the literals staged through L2).
//
// This kernel takes the single-block frames (every file up to 128 KiB: almost all of a source tree) and whatever the
// staged pipeline's header walk refused (the serial decode names the error).  Frames of two blocks and more go through
// the staged pipeline (zstd_decode_staged.cu): all their blocks entropy-decoded in parallel, then executed in block
// order.  Items are handed out largest first, one atomicAdd per batch.
// The uniform per-block helpers (table builds, literals, warp copies) are kept OUT OF LINE on purpose: inlined,
// the kernel was 207 KB of code with 700 B of spills and 12 % slower.
#include "zstd_decode.cuh"
#ifndef ZD_B_STEPS
#define ZD_B_STEPS 64   // sequences a lane decodes before the warp reconverges (phase B)
#endif

// ---------------------------------------------------------------------------------------------
// Frames are handed to the warps in descending order of compressed size (counting sort), so that the
// 32 frames of a batch carry similar numbers of sequences (phase B runs in lock-step) and the big
// frames start first.
#define ZD_BINS 4096u
ZG_DEV u32 zd_bin(u64 len) { return ZD_BINS - 1u - (u32)zg_min<u64>(len >> 4, ZD_BINS - 1u); }
// (`list`: the frames to decode, or null for all of them; an item t is frame list[t])
__global__ void __launch_bounds__(256) k_zd_bin_count(const u64* __restrict__ len, const u32* __restrict__ list, u64 n, u32* bins) {
	u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (t < n) atomicAdd(&bins[zd_bin(len[list ? list[t] : t])], 1u);
}
__global__ void __launch_bounds__(1024) k_zd_bin_scan(u32* bins) {
	// exclusive scan of ZD_BINS counters by one CTA (4 per thread)
	__shared__ u32 part[32];
	u32 t = threadIdx.x, lane = t & 31, w = t >> 5;
	u32 v[4], sum = 0;
	for (u32 i = 0; i < 4; i++) {
		v[i] = bins[4 * t + i];
		sum += v[i];
	}
	u32 incl = zg_warp_incl_scan(sum);
	if (lane == 31) part[w] = incl;
	__syncthreads();
	if (w == 0) {
		u32 p = part[lane];
		u32 pi = zg_warp_incl_scan(p);
		part[lane] = pi - p;
	}
	__syncthreads();
	u32 run = part[w] + incl - sum;
	for (u32 i = 0; i < 4; i++) {
		bins[4 * t + i] = run;
		run += v[i];
	}
}
__global__ void __launch_bounds__(256) k_zd_bin_scatter(const u64* __restrict__ len, const u32* __restrict__ list, u64 n, u32* bins, u32* perm) {
	u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (t < n) perm[atomicAdd(&bins[zd_bin(len[list ? list[t] : t])], 1u)] = (u32)t;
}

__global__ void __launch_bounds__(ZD_WARPS * 32, ZD_MIN_CTAS)
k_zstd_decode_frames(const u8* __restrict__ archive, u64 archive_len, const u64* __restrict__ off, const u64* __restrict__ len,
                     const u64* __restrict__ ulen, const u64* __restrict__ out_off, u64 nframes, u8* out, u64 out_cap,
                     const u32* __restrict__ perm, const u32* __restrict__ list, u64* seq_arenas, u8* litbufs, u32* tabs, u8* hufsaves, u32* queue,
                     u32* status, u64* produced, u32* cksums, u32 cap_div, u32 min_batch, u32 floor_bytes, u32 share_q) {
	ZG_DYN_SMEM(ZdWarp, sm);
	u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	ZdWarp* W = &sm[warp];
	size_t gw = (size_t)blockIdx.x * ZD_WARPS + warp;
	u64* arena = seq_arenas + gw * ZD_SEQ_ARENA;
	u8* litbuf = litbufs + gw * ZD_LITBUF;
	u32* my_slot = tabs + (gw * 32 + lane) * ZD_TAB_SLOT;
	u8* hufsave0 = hufsaves + gw * 32 * ZD_HUFSAVE;
	u64 hint_len = ~0ull;  // compressed size of the last item of this warp's previous batch
	u64 prev_base = 0;
	for (;;) {
		// Hand-out: frames come largest first.  A batch is up to 32 frames, but (a) no more than about
		// ZD_BATCH_BYTES of compressed input -- the phases around the lane-parallel sequence decode run
		// frame after frame, so 32 of the largest frames in one warp would be the kernel's tail --
		// and (b) smaller towards the end, so that the last round is spread over all warps.
		u32 base = 0, want = 32;
		if (lane == 0) {
			// The batch size comes from the size of the last item this warp saw (items come largest first, so it
			// bounds every later one) and from this warp's previous position in the queue: nothing has to be read
			// before the claim, and the claim is one atomicAdd -- no compare-and-swap loop for ~3000 warps to fight over.
			u64 flen = hint_len;
			if (flen == ~0ull) {
				u32 t0 = perm ? perm[0] : 0u;
				flen = len[list ? list[t0] : t0];
			}
			u64 left = nframes > prev_base ? nframes - prev_base : 1;
			u64 share = 4 * left / ((u64)share_q * gridDim.x * ZD_WARPS);  // share_q quarters of ... (8: half a fair share)
			// a quarter of a warp's fair share of the input, but at least min_batch bytes
			u64 cap = zg_max<u64>(min_batch, archive_len / ((u64)cap_div * gridDim.x * ZD_WARPS));
			u64 fit = cap / (flen ? flen : 1);
			// ... and never so small that claiming a batch costs as much as decoding it (the smallest
			// frames come last: they always go out as full rows)
			u64 floor_b = floor_bytes / (flen ? flen : 1);
			want = (u32)zg_min<u64>(zg_min<u64>(32, zg_max<u64>(zg_max<u64>(4, share), floor_b)), zg_max<u64>(2, fit));
			base = atomicAdd(queue, want);
		}
		base = __shfl_sync(ZG_FULL, base, 0);
		want = __shfl_sync(ZG_FULL, want, 0);
		if (base >= nframes) break;
		prev_base = base;
		bool mine = lane < want && (u64)base + lane < nframes;
		u64 t = mine ? (perm ? (u64)perm[base + lane] : (u64)base + lane) : 0;  // work item
		u64 k = (mine && list) ? (u64)list[t] : t;                              // its frame
		// ---- lane-private frame header ----
		ZdLane L;
		L.src = archive;
		L.out = out;
		L.n = L.ip = L.cap = L.opos = L.fcs = L.base = 0;
		L.status = ZS_OK;
		L.flags = 0;
		L.fcs_len = L.cksum = 0;
		L.rep0 = 1;
		L.rep1 = 4;
		L.rep2 = 8;
		L.ll_log = L.ml_log = L.of_log = 0;
		L.blk = archive;
		L.blk_n = L.nseq = L.nseq_left = L.seq_base = 0;
		L.b.start = L.b.ptr = archive;
		L.b.lo = L.b.hi = L.b.consumed = 0;
		L.sl = L.so = L.sm = 0;
		u64 mylen = 0;
		if (mine) {
			u64 fo = off[k], fl = len[k], ul = ulen[k], oo = out_off[k];
			mylen = fl;
			if (fo > archive_len || fl > archive_len - fo) L.status = ZS_E_SRC_SIZE;
			else if (oo > out_cap || ul > out_cap - oo) L.status = ZS_E_DST_SMALL;
			else {
				L.src = archive + fo;
				L.n = fl;
				L.out = out + oo;
				L.cap = ul;
				L.flags = ZD_F_ACTIVE;
				zd_frame_header(L);
			}
		}
		{
			u32 have = __ballot_sync(ZG_FULL, mine);
			hint_len = __shfl_sync(ZG_FULL, mylen, have ? 31 - __clz((int)have) : 0);
		}
		__syncwarp();
		// ---- rounds: one block per live frame ----
		for (;;) {
			u32 live = __ballot_sync(ZG_FULL, (L.flags & ZD_F_ACTIVE) != 0);
			if (!live) break;
			// A: set up the next block of every live frame
			u32 arena_used = 0;
			u32 todo = live;
			while (todo) {
				int f = __ffs((int)todo) - 1;
				todo &= todo - 1;
				ZdLane U = zd_bcast(L, f);
				zd_setup_frame(W, U, tabs + (gw * 32 + (u32)f) * ZD_TAB_SLOT, arena_used, litbuf, hufsave0 + (u32)f * ZD_HUFSAVE);
				if (lane == (u32)f) L = U;
				__syncwarp();
			}
			// B: every lane decodes its block's sequences (in bounded steps, so the warp reconverges)
			{
				const u32* llt = (L.flags & ZD_F_LL_DEF) ? ZS_LL_DEFAULT_DTABLE : my_slot;
				const u32* mlt = (L.flags & ZD_F_ML_DEF) ? ZS_ML_DEFAULT_DTABLE : my_slot + 512;
				const u32* oft = (L.flags & ZD_F_OF_DEF) ? ZS_OF_DEFAULT_DTABLE : my_slot + 1024;
				for (;;) {
					u32 cnt = zg_min<u32>((u32)ZD_B_STEPS, L.nseq_left);
					if (!__any_sync(ZG_FULL, cnt > 0)) break;
					if (cnt && !zd_lane_decode(L, cnt, llt, mlt, oft, arena + L.seq_base + (L.nseq - L.nseq_left))) zd_fail(L, ZS_E_CORRUPT);
				}
			}
			__syncwarp();
			// C: execute frame after frame
			todo = __ballot_sync(ZG_FULL, (L.flags & ZD_F_ACTIVE) && L.nseq > 0);
			while (todo) {
				int f = __ffs((int)todo) - 1;
				todo &= todo - 1;
				ZdLane U = zd_bcast(L, f);
				zd_exec_block(W, U, arena + U.seq_base, litbuf, hufsave0 + (u32)f * ZD_HUFSAVE);
				if (lane == (u32)f) L = U;
				__syncwarp();
			}
		}
		if (mine) {
			bool ok = L.status == ZS_OK;
			status[k] = L.status;
			produced[k] = ok ? L.opos : 0;
			cksums[2 * k] = ok ? L.cksum : 0;
			cksums[2 * k + 1] = (ok && (L.flags & ZD_F_HAS_CK)) ? 1u : 0u;
		}
		__syncwarp();
	}
}

// hand-out tuning (see the kernel): a batch holds at most 1/cap_div of a warp's fair share of the input but at least
// min_batch bytes; frames so small that 32 of them stay below floor_bytes always go out as full rows
static u32 g_zd_tune[4] = {4, ZD_BATCH_BYTES, 32u << 10, 8};
extern "C" void zg_internal_set_decode_share(u32 q) { g_zd_tune[3] = q ? q : 8; }
extern "C" void zg_internal_set_decode_batching(u32 cap_div, u32 min_batch, u32 floor_bytes) {
	g_zd_tune[0] = cap_div ? cap_div : 4;
	g_zd_tune[1] = min_batch ? min_batch : ZD_BATCH_BYTES;
	g_zd_tune[2] = floor_bytes ? floor_bytes : (32u << 10);
}

// one launch of the fused kernel over `count` frames (those in `list`, or all when it is null), handed out largest first.
// Multi-block frames go through the staged pipeline instead (zstd_decode_staged.cu, which also holds zg_zstd_decode_run).
size_t zd_fused_launch(cudaStream_t s, ZgZdWork& w, const u8* archive, u64 archive_len, const u64* off, const u64* len, const u64* ulen,
                       const u64* out_off, const u32* list, u64 count, u8* out, u64 out_cap, u32* status, u64* produced, u32* cksums) {
	if (count == 0) return 0;
	if (count >= 0xff000000ull) return ZG_ERR(ZG_error_GENERIC);  // (the queue counter runs past count by up to one batch per warp)
	u64 batches = (count + 31) / 32;
	u32 grid = (u32)zg_min<u64>((batches + ZD_WARPS - 1) / ZD_WARPS, (u64)zg_sm_count() * ZD_MIN_CTAS);
	size_t warps = (size_t)grid * ZD_WARPS;
	if (w.seqs.reserve(warps * ZD_SEQ_ARENA * 8) || w.lit.reserve(warps * ZD_LITBUF) || w.tabs.reserve(warps * 32 * ZD_TAB_SLOT * 4) ||
	    w.hufsave.reserve(warps * 32 * ZD_HUFSAVE) || w.queue.reserve(16) || w.bins.reserve(ZD_BINS * 4) || w.perm.reserve(count * 4))
		return ZG_ERR(ZG_error_memory_allocation);
	cudaMemsetAsync(w.queue.p, 0, 16, s);
	const u32* perm = nullptr;
	if (count > 64) {
		cudaMemsetAsync(w.bins.p, 0, ZD_BINS * 4, s);
		u32 g = (u32)((count + 255) / 256);
		ZG_LAUNCH(k_zd_bin_count, g, 256, 0, s, len, list, count, w.bins.as<u32>());
		ZG_LAUNCH(k_zd_bin_scan, 1, 1024, 0, s, w.bins.as<u32>());
		ZG_LAUNCH(k_zd_bin_scatter, g, 256, 0, s, len, list, count, w.bins.as<u32>(), w.perm.as<u32>());
		g_zg_launches += 3;
		perm = w.perm.as<u32>();
	}
	size_t smem = sizeof(ZdWarp) * ZD_WARPS;
	static ZgPerDevice attr_dev;
	bool& attr_set = *attr_dev.slot();
	if (!attr_set) {
		if (cudaFuncSetAttribute(k_zstd_decode_frames, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
			return ZG_ERR(ZG_error_device);
		attr_set = true;
	}
	zg_prof_begin(ZG_K_DECODE, s);
	ZG_LAUNCH(k_zstd_decode_frames, grid, ZD_WARPS * 32, smem, s, archive, archive_len, off, len, ulen, out_off, count, out, out_cap,
	          perm, list, w.seqs.as<u64>(), w.lit.as<u8>(), w.tabs.as<u32>(), w.hufsave.as<u8>(), w.queue.as<u32>(), status, produced, cksums,
	          g_zd_tune[0], g_zd_tune[1], g_zd_tune[2], g_zd_tune[3]);
	zg_prof_end(ZG_K_DECODE, s);
	ZG_COUNT_LAUNCH();
	return cudaGetLastError() == cudaSuccess ? 0 : ZG_ERR(ZG_error_device);
}
