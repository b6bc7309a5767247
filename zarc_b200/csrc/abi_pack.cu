// C ABI, pack side: zg_cctx (CCtx analogue + the Encoder's content state), zg_pack_batch[_dev],
// zg_compress2.  Host code here only moves buffers and does bookkeeping on counts/sizes.
#include "common.h"
#include <mutex>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <vector>
#include <string.h>

int zg_abi_dev_count();  // abi.cu

#define ZG_NEED_DEVICE() \
	do {                 \
		if (zg_abi_dev_count() <= 0) return ZG_ERR(ZG_error_no_device); \
	} while (0)
#define ZG_TRY(expr)              \
	do {                          \
		size_t r_ = (expr);       \
		if (zg_is_error(r_)) return r_; \
	} while (0)
#define ZG_CUDA(expr) \
	do {              \
		if ((expr) != cudaSuccess) return ZG_ERR(ZG_error_device); \
	} while (0)
#define ZG_ALLOC(expr) \
	do {               \
		if ((expr) != cudaSuccess) return ZG_ERR(ZG_error_memory_allocation); \
	} while (0)
#define ZG_GUARD(expr)                                    \
	try {                                                 \
		return (expr);                                    \
	} catch (const std::bad_alloc&) {                     \
		return ZG_ERR(ZG_error_memory_allocation);        \
	} catch (...) {                                       \
		return ZG_ERR(ZG_error_GENERIC);                  \
	}

// pack.cu
size_t zg_pk_dedup_insert(cudaStream_t s, const u8* g_digest, u64 lo, u64 hi, u32* table, u32 mask);
size_t zg_pk_dedup_resolve(cudaStream_t s, const u8* g_digest, u64 base, u64 n, const u32* table, u32 mask, const u64* len, const u8* select,
                           u64* rep, u8* first, u64* isfirst64, u64* nblk, u64* clen);
size_t zg_pk_build_ulist(cudaStream_t s, const u8* first, const u64* uidx, const u64* blk_first, u64 n, u32* ulist, u64* blk_base,
                         u64 nuniq, u64 nblocks);
size_t zg_pk_block_out_sizes(cudaStream_t s, const u32* blk_csize, u64 nblocks, u64* blk_out);
size_t zg_pk_frame_sizes(cudaStream_t s, const u32* ulist, const u64* blk_base, const u64* blk_pos, const u64* blk_total, const u64* len,
                         u64 nuniq, u32 flags, u64* frame_len_u);
size_t zg_pk_xxh64_list(cudaStream_t s, const u8* blob, const u64* off, const u64* len, const u32* ulist, u64 nuniq, u64* out);
size_t zg_pk_frame_assemble(cudaStream_t s, const u8* blob, const u8* comp, const u64* file_off, const u64* comp_off, const u64* file_len,
                            const u32* ulist, const u64* blk_base, const u64* blk_pos, const u32* blk_csize, const u64* frame_off_u,
                            const u64* xxh, u64 nuniq, u64 nblocks, u32 flags, u64 archive_base, u8* frames_out);
size_t zg_pk_record_answer(cudaStream_t s, const u32* ulist, const u64* frame_off_u, const u64* frame_len_u, u64 nuniq, u64 base,
                           u64* g_off, u64* g_len, const u64* rep, u64 n, u64* frame_off, u64* frame_len);

// The Encoder's content state: `frames: HashMap<Digest, Frame>` + `offset` (encode.rs:32,36)
struct ZgArchive {
	u64 offset = 0;
	u64 nfiles = 0;  // global file ids issued so far
	ZgBuf g_digest, g_off, g_len;
	ZgBuf table;
	u32 table_size = 0;
	void release() {
		g_digest.release();
		g_off.release();
		g_len.release();
		table.release();
		table_size = 0;
		nfiles = 0;
	}
};

struct zg_cctx {
	cudaStream_t stream = 0;
	bool own_stream = false;
	int level = 3, checksum = 0, content_size = 1;
	int other_params[16] = {0};
	ZgArchive archive, oneshot;
	ZgB3Work b3;
	ZgZeWork ze;
	ZgBuf tiles, rep, isfirst64, nblk, clen, uidx, blkfirst, comp_off, ulist, blk_base, blk_csize, blk_out, blk_pos, frame_len_u,
	    frame_off_u, xxh, comp, totals, first_tmp;
	ZgHostBuf h;
	// host-API staging: two stages, so that the copies of one slice overlap the kernels of another
	struct Stage {
		ZgBuf d_blob, d_meta, d_out_meta, d_frames;
		ZgHostBuf h_meta;
		cudaEvent_t in_done = nullptr, out_done = nullptr;
	} stage[2];
	cudaStream_t s_in = nullptr, s_out = nullptr;
	// the Content_Checksum (XXH64: one serial multiply-rotate chain per file) runs beside the encoder, not after it
	cudaStream_t s_xx = nullptr;
	cudaEvent_t xx_fork = nullptr, xx_join = nullptr;
};

static cudaError_t grow_keep(ZgBuf& b, size_t need, size_t keep, cudaStream_t s) {
	if (need <= b.cap) return cudaSuccess;
	ZgBuf nb;
	cudaError_t e = nb.reserve(need * 2);
	if (e != cudaSuccess) return e;
	if (keep && b.p) {
		e = cudaMemcpyAsync(nb.p, b.p, keep, cudaMemcpyDeviceToDevice, s);
		if (e != cudaSuccess) return e;
		cudaStreamSynchronize(s);
	}
	b.release();
	b = nb;
	return cudaSuccess;
}

// Undo everything a failed batch did to the Encoder state: the map goes back to the ids below `nfiles0` (the table is
// rebuilt from the kept digests -- open addressing has no cheap delete) and the running offset to `offset0`, so a
// retry sees exactly the state the failed call started from.
static void archive_rollback(zg_cctx* c, ZgArchive& A, u64 nfiles0, u64 offset0) {
	cudaStream_t s = c->stream;
	A.nfiles = nfiles0;
	A.offset = offset0;
	if (A.table_size) {
		cudaMemsetAsync(A.table.p, 0, (size_t)A.table_size * 4, s);
		if (nfiles0) zg_pk_dedup_insert(s, A.g_digest.as<u8>(), 0, nfiles0, A.table.as<u32>(), A.table_size - 1);
		cudaStreamSynchronize(s);
	}
}

static size_t pack_core_body(zg_cctx* c, ZgArchive& A, const u8* blob, const u64* off, const u64* len, u64 F, u8* digests_out, u8* first_out,
                             u64* frame_off_out, u64* frame_len_out, u8* frames_out, u64 frames_cap, u64* frames_bytes, const u8* digests_in,
                             const u8* select);
// one batch, all or nothing: on any error the archive state is what it was on entry
static size_t pack_core(zg_cctx* c, ZgArchive& A, const u8* blob, const u64* off, const u64* len, u64 F, u8* digests_out, u8* first_out,
                        u64* frame_off_out, u64* frame_len_out, u8* frames_out, u64 frames_cap, u64* frames_bytes,
                        const u8* digests_in = nullptr, const u8* select = nullptr) {
	u64 nfiles0 = A.nfiles, offset0 = A.offset;
	size_t r = pack_core_body(c, A, blob, off, len, F, digests_out, first_out, frame_off_out, frame_len_out, frames_out, frames_cap, frames_bytes,
	                          digests_in, select);
	if (zg_is_error(r)) {
		archive_rollback(c, A, nfiles0, offset0);
		if (frames_bytes) *frames_bytes = 0;
	}
	return r;
}

static size_t pack_core_body(zg_cctx* c, ZgArchive& A, const u8* blob, const u64* off, const u64* len, u64 F, u8* digests_out, u8* first_out,
                             u64* frame_off_out, u64* frame_len_out, u8* frames_out, u64 frames_cap, u64* frames_bytes, const u8* digests_in,
                             const u8* select) {
	cudaStream_t s = c->stream;
	if (frames_bytes) *frames_bytes = 0;
	if (F == 0) return 0;
	const bool trace = getenv("ZG_TRACE") != nullptr;
	auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	double tc0 = now_ms();
	u64 base = A.nfiles;
	if (base + F >= 0xfffffff0ull) return ZG_ERR(ZG_error_GENERIC);
	// the map's storage, indexed by global file id
	ZG_ALLOC(grow_keep(A.g_digest, (base + F) * 32, base * 32, s));
	ZG_ALLOC(grow_keep(A.g_off, (base + F) * 8, base * 8, s));
	ZG_ALLOC(grow_keep(A.g_len, (base + F) * 8, base * 8, s));
	u8* g_digest = A.g_digest.as<u8>();
	// (1) digests (content_frame.rs:26) -- or the ones the caller already computed (multi-GPU: the digests are needed
	//     before this call for the cross-rank first-occurrence decision)
	if (digests_in) ZG_CUDA(cudaMemcpyAsync(g_digest + 32 * base, digests_in, F * 32, cudaMemcpyDeviceToDevice, s));
	else ZG_TRY(zg_blake3_run(s, c->b3, blob, off, len, F, g_digest + 32 * base));
	if (digests_out && digests_out != digests_in) ZG_CUDA(cudaMemcpyAsync(digests_out, g_digest + 32 * base, F * 32, cudaMemcpyDeviceToDevice, s));
	// files that end up without a frame of their own in this archive part answer with offset 0 / length 0
	ZG_CUDA(cudaMemsetAsync(A.g_off.as<u64>() + base, 0, F * 8, s));
	ZG_CUDA(cudaMemsetAsync(A.g_len.as<u64>() + base, 0, F * 8, s));
	// (2) dedup (content_frame.rs:30): the table keeps, per digest, the smallest global id
	u64 want = 1024;
	while (want < 2 * (base + F)) want <<= 1;
	if (want > A.table_size) {
		A.table.release();
		ZG_ALLOC(A.table.reserve(want * 4));
		A.table_size = (u32)want;
		ZG_CUDA(cudaMemsetAsync(A.table.p, 0, want * 4, s));
		ZG_TRY(zg_pk_dedup_insert(s, g_digest, 0, base, A.table.as<u32>(), A.table_size - 1));
	}
	ZG_TRY(zg_pk_dedup_insert(s, g_digest, base, base + F, A.table.as<u32>(), A.table_size - 1));
	ZG_ALLOC(c->rep.reserve(F * 8));
	ZG_ALLOC(c->isfirst64.reserve(F * 8));
	ZG_ALLOC(c->nblk.reserve(F * 8));
	ZG_ALLOC(c->clen.reserve(F * 8));
	ZG_ALLOC(c->uidx.reserve(F * 8));
	ZG_ALLOC(c->blkfirst.reserve(F * 8));
	ZG_ALLOC(c->comp_off.reserve(F * 8));
	ZG_ALLOC(c->totals.reserve(64));
	ZG_ALLOC(c->h.reserve(64));
	u8* first = first_out;
	if (!first) {
		ZG_ALLOC(c->first_tmp.reserve(F));
		first = c->first_tmp.as<u8>();
	}
	u64* totals = c->totals.as<u64>();
	ZG_TRY(zg_pk_dedup_resolve(s, g_digest, base, F, A.table.as<u32>(), A.table_size - 1, len, select, c->rep.as<u64>(), first,
	                           c->isfirst64.as<u64>(), c->nblk.as<u64>(), c->clen.as<u64>()));
	ZG_TRY(zg_scan_run(s, c->tiles, c->isfirst64.as<u64>(), F, 0, c->uidx.as<u64>(), totals + 0));
	ZG_TRY(zg_scan_run(s, c->tiles, c->nblk.as<u64>(), F, 0, c->blkfirst.as<u64>(), totals + 1));
	ZG_TRY(zg_scan_run(s, c->tiles, c->clen.as<u64>(), F, 0, c->comp_off.as<u64>(), totals + 2));
	u64* ht = c->h.as<u64>();
	ZG_CUDA(zg_publish(s, totals, ht, 24));
	ZG_CUDA(cudaStreamSynchronize(s));
	u64 nuniq = ht[0], nblocks = ht[1], comp_bytes = ht[2];
	double tc1 = now_ms();
	u32 flags = (c->checksum ? 1u : 0u) | (c->content_size ? 2u : 0u);
	u64 new_offset = A.offset;
	if (nuniq) {
		ZG_ALLOC(c->ulist.reserve(nuniq * 4));
		ZG_ALLOC(c->blk_base.reserve((nuniq + 1) * 8));
		ZG_ALLOC(c->blk_csize.reserve(nblocks * 4));
		ZG_ALLOC(c->blk_out.reserve(nblocks * 8));
		ZG_ALLOC(c->blk_pos.reserve(nblocks * 8));
		ZG_ALLOC(c->frame_len_u.reserve(nuniq * 8));
		ZG_ALLOC(c->frame_off_u.reserve(nuniq * 8));
		ZG_ALLOC(c->xxh.reserve(nuniq * 8));
		ZG_ALLOC(c->comp.reserve(comp_bytes + 64));
		ZG_TRY(zg_pk_build_ulist(s, first, c->uidx.as<u64>(), c->blkfirst.as<u64>(), F, c->ulist.as<u32>(), c->blk_base.as<u64>(), nuniq,
		                         nblocks));
		// (3) compress every new content (content_frame.rs:41 -> lowlevel_frames.rs:30)
		const ZgCParams cparams = {c->level, c->other_params[ZG_c_windowLog - 100], c->other_params[ZG_c_hashLog - 100],
		                           c->other_params[ZG_c_searchLog - 100], c->other_params[ZG_c_minMatch - 100],
		                           c->other_params[ZG_c_strategy - 100]};
		bool side = false;
		if (c->checksum) {
			// the checksums only need the input and the unique list: fork them onto a side stream (launched first, higher
			// priority, so that its few long-running warps are resident while the encoder's kernels fill the rest)
			if (!c->s_xx) {
				int lo = 0, hi = 0;
				cudaDeviceGetStreamPriorityRange(&lo, &hi);
				if (cudaStreamCreateWithPriority(&c->s_xx, cudaStreamNonBlocking, hi) != cudaSuccess ||
				    cudaEventCreateWithFlags(&c->xx_fork, cudaEventDisableTiming) != cudaSuccess ||
				    cudaEventCreateWithFlags(&c->xx_join, cudaEventDisableTiming) != cudaSuccess)
					c->s_xx = nullptr;
			}
			side = c->s_xx != nullptr;
			if (side) {
				ZG_CUDA(cudaEventRecord(c->xx_fork, s));
				ZG_CUDA(cudaStreamWaitEvent(c->s_xx, c->xx_fork, 0));
				ZG_TRY(zg_pk_xxh64_list(c->s_xx, blob, off, len, c->ulist.as<u32>(), nuniq, c->xxh.as<u64>()));
				ZG_CUDA(cudaEventRecord(c->xx_join, c->s_xx));
			}
		}
		ZG_TRY(zg_zstd_encode_run(s, c->ze, blob, off, c->comp_off.as<u64>(), len, c->ulist.as<u32>(), c->blk_base.as<u64>(), (u32)nuniq,
		                          nblocks, comp_bytes, c->comp.as<u8>(), c->blk_csize.as<u32>(), cparams));
		if (c->checksum && !side) ZG_TRY(zg_pk_xxh64_list(s, blob, off, len, c->ulist.as<u32>(), nuniq, c->xxh.as<u64>()));
		if (side) ZG_CUDA(cudaStreamWaitEvent(s, c->xx_join, 0));
		// (4) frame lengths -> archive offsets (content_frame.rs:22,45)
		ZG_TRY(zg_pk_block_out_sizes(s, c->blk_csize.as<u32>(), nblocks, c->blk_out.as<u64>()));
		ZG_TRY(zg_scan_run(s, c->tiles, c->blk_out.as<u64>(), nblocks, 0, c->blk_pos.as<u64>(), totals + 3));
		ZG_TRY(zg_pk_frame_sizes(s, c->ulist.as<u32>(), c->blk_base.as<u64>(), c->blk_pos.as<u64>(), totals + 3, len, nuniq, flags,
		                         c->frame_len_u.as<u64>()));
		ZG_TRY(zg_scan_run(s, c->tiles, c->frame_len_u.as<u64>(), nuniq, A.offset, c->frame_off_u.as<u64>(), totals + 4));
		ZG_CUDA(zg_publish(s, totals + 4, ht + 4, 8));
		ZG_CUDA(cudaStreamSynchronize(s));
		new_offset = ht[4];
		if (trace) fprintf(stderr, "[zg pack]   digests+dedup+scans %.2f ms, encode+sizes %.2f ms\n", tc1 - tc0, now_ms() - tc1);
		u64 bytes = new_offset - A.offset;
		if (bytes > frames_cap) return ZG_ERR(ZG_error_dstSize_tooSmall);
		if (!frames_out) return ZG_ERR(ZG_error_dstBuffer_null);
		// (5) write the frames in insertion order
		ZG_TRY(zg_pk_frame_assemble(s, blob, c->comp.as<u8>(), off, c->comp_off.as<u64>(), len, c->ulist.as<u32>(), c->blk_base.as<u64>(),
		                            c->blk_pos.as<u64>(), c->blk_csize.as<u32>(), c->frame_off_u.as<u64>(), c->xxh.as<u64>(), nuniq,
		                            nblocks, flags, A.offset, frames_out));
		if (frames_bytes) *frames_bytes = bytes;
	}
	// (6) Frame records (content_frame.rs:48-57) and per-file answers
	ZG_TRY(zg_pk_record_answer(s, c->ulist.as<u32>(), c->frame_off_u.as<u64>(), c->frame_len_u.as<u64>(), nuniq, base, A.g_off.as<u64>(),
	                           A.g_len.as<u64>(), c->rep.as<u64>(), F, frame_off_out, frame_len_out));
	ZG_CUDA(cudaStreamSynchronize(s));
	ZG_CUDA(cudaGetLastError());
	A.nfiles = base + F;
	A.offset = new_offset;
	return 0;
}

extern "C" {

zg_cctx* zg_cctx_create(void) {
	if (zg_abi_dev_count() <= 0) return nullptr;
	zg_cctx* c = new (std::nothrow) zg_cctx();
	if (!c) return nullptr;
	if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
		delete c;
		return nullptr;
	}
	c->own_stream = true;
	return c;
}
void zg_cctx_free(zg_cctx* c) {
	if (!c) return;
	cudaStreamSynchronize(c->stream);
	c->archive.release();
	c->oneshot.release();
	zg_b3work_free(c->b3);
	c->ze.release();
	for (ZgBuf* b : {&c->tiles, &c->rep, &c->isfirst64, &c->nblk, &c->clen, &c->uidx, &c->blkfirst, &c->comp_off, &c->ulist, &c->blk_base,
	                 &c->blk_csize, &c->blk_out, &c->blk_pos, &c->frame_len_u, &c->frame_off_u, &c->xxh, &c->comp, &c->totals, &c->first_tmp})
		b->release();
	for (auto& st : c->stage) {
		for (ZgBuf* b : {&st.d_blob, &st.d_meta, &st.d_out_meta, &st.d_frames}) b->release();
		st.h_meta.release();
		if (st.in_done) cudaEventDestroy(st.in_done);
		if (st.out_done) cudaEventDestroy(st.out_done);
	}
	if (c->s_in) cudaStreamDestroy(c->s_in);
	if (c->s_out) cudaStreamDestroy(c->s_out);
	if (c->s_xx) {
		cudaStreamSynchronize(c->s_xx);
		cudaStreamDestroy(c->s_xx);
		cudaEventDestroy(c->xx_fork);
		cudaEventDestroy(c->xx_join);
	}
	c->h.release();
	if (c->own_stream) cudaStreamDestroy(c->stream);
	delete c;
}
size_t zg_cctx_set_stream(zg_cctx* c, void* stream) {
	if (!c) return ZG_ERR(ZG_error_GENERIC);
	if (c->own_stream) cudaStreamDestroy(c->stream);
	c->own_stream = false;
	c->stream = (cudaStream_t)stream;
	return 0;
}
// CCtx::init(level) == ZSTD_initCStream(cctx, level): session reset + compressionLevel (encode.rs:62)
size_t zg_cctx_init(zg_cctx* c, int level) {
	if (!c) return ZG_ERR(ZG_error_GENERIC);
	if (level < -131072 || level > 22) return ZG_ERR(ZG_error_parameter_outOfBound);
	c->level = level == 0 ? 3 : level;
	return 0;
}
size_t zg_cctx_set_parameter(zg_cctx* c, int param, int value) {
	if (!c) return ZG_ERR(ZG_error_GENERIC);
	switch (param) {
	case ZG_c_compressionLevel:
		if (value < -131072 || value > 22) return ZG_ERR(ZG_error_parameter_outOfBound);
		c->level = value == 0 ? 3 : value;
		return (size_t)(value < 0 ? 0 : value);
	case ZG_c_checksumFlag:
		c->checksum = value != 0;
		return (size_t)c->checksum;
	case ZG_c_contentSizeFlag:
		c->content_size = value != 0;
		return (size_t)c->content_size;
	case ZG_c_dictIDFlag:
		return (size_t)(value != 0);
	case ZG_c_windowLog:
		if (value != 0 && (value < 10 || value > 31)) return ZG_ERR(ZG_error_parameter_outOfBound);
		c->other_params[param - 100] = value;
		return (size_t)value;
	// the --zstd parameters (pack.rs:140-195), with libzstd's bounds; how each one steers the GPU match finder is in
	// ze_resolve_params (zstd_encode.cu).  A parameter that cannot be honoured is refused, never silently ignored.
	case ZG_c_hashLog:
		if (value != 0 && (value < 6 || value > 30)) return ZG_ERR(ZG_error_parameter_outOfBound);
		c->other_params[param - 100] = value;  // tables above 2^12 positions do not exist here: adjusted down, as libzstd adjusts cparams to the input
		return (size_t)value;
	case ZG_c_searchLog:
		if (value != 0 && (value < 1 || value > 30)) return ZG_ERR(ZG_error_parameter_outOfBound);
		c->other_params[param - 100] = value;
		return (size_t)value;
	case ZG_c_minMatch:
		if (value != 0 && (value < 3 || value > 7)) return ZG_ERR(ZG_error_parameter_outOfBound);
		if (value == 3) return ZG_ERR(ZG_error_parameter_unsupported);  // the 4-byte hash cannot find 3-byte matches
		c->other_params[param - 100] = value;
		return (size_t)value;
	case ZG_c_strategy:
		if (value != 0 && (value < 1 || value > 9)) return ZG_ERR(ZG_error_parameter_outOfBound);
		c->other_params[param - 100] = value;
		return (size_t)value;
	case ZG_c_chainLog:
	case ZG_c_targetLength:
		// no chain table and no optimal parser here: only the "not set" value is accepted
		if (value != 0) return ZG_ERR(ZG_error_parameter_unsupported);
		return 0;
	default:
		return ZG_ERR(ZG_error_parameter_unsupported);
	}
}
size_t zg_cctx_reset(zg_cctx* c, int directive) {
	if (!c) return ZG_ERR(ZG_error_GENERIC);
	if (directive < 1 || directive > 3) return ZG_ERR(ZG_error_parameter_outOfBound);
	// session_only: nothing is in flight between calls (content_frame.rs:37-39 does this per frame)
	if (directive & ZG_reset_parameters) {
		c->level = 3;
		c->checksum = 0;
		c->content_size = 1;
		memset(c->other_params, 0, sizeof c->other_params);
	}
	return 0;
}
size_t zg_cctx_reset_archive(zg_cctx* c, uint64_t first_frame_offset) {
	if (!c) return ZG_ERR(ZG_error_GENERIC);
	cudaStreamSynchronize(c->stream);
	c->archive.nfiles = 0;
	c->archive.offset = first_frame_offset;
	if (c->archive.table_size) ZG_CUDA(cudaMemsetAsync(c->archive.table.p, 0, (size_t)c->archive.table_size * 4, c->stream));
	return 0;
}
uint64_t zg_cctx_archive_offset(const zg_cctx* c) { return c ? c->archive.offset : 0; }
size_t zg_compress_bound(size_t n) { return n + (n >> 8) + (n < (128 << 10) ? (((128 << 10) - n) >> 11) : 0); }  // ZSTD_compressBound

size_t zg_pack_batch_dev(zg_cctx* c, const uint8_t* blob, const uint64_t* off, const uint64_t* len, uint64_t n, uint8_t* digests,
                         uint8_t* first, uint64_t* frame_off, uint64_t* frame_len, uint8_t* frames_out, uint64_t frames_cap,
                         uint64_t* frames_bytes) {
	ZG_NEED_DEVICE();
	if (!c) return ZG_ERR(ZG_error_GENERIC);
	return pack_core(c, c->archive, blob, off, len, n, digests, first, frame_off, frame_len, frames_out, frames_cap, frames_bytes);
}

// add_data_frame with its two decisions made by the caller (multi-GPU): `digests_in` (may be NULL)
// are the files' BLAKE3 digests computed earlier (zg_blake3_batch_dev) -- the digest of content_frame.rs:26 is needed
// before the cross-rank dedup -- and `select` (may be NULL) is the global "first occurrence" answer of
// content_frame.rs:30 for each file: a file with select[i] == 0 gets no frame here (first[i] = 0, frame_len[i] = 0).
size_t zg_pack_batch_dev_ex(zg_cctx* c, const uint8_t* blob, const uint64_t* off, const uint64_t* len, uint64_t n, const uint8_t* digests_in,
                            const uint8_t* select, uint8_t* digests, uint8_t* first, uint64_t* frame_off, uint64_t* frame_len,
                            uint8_t* frames_out, uint64_t frames_cap, uint64_t* frames_bytes) {
	ZG_NEED_DEVICE();
	if (!c) return ZG_ERR(ZG_error_GENERIC);
	return pack_core(c, c->archive, blob, off, len, n, digests, first, frame_off, frame_len, frames_out, frames_cap, frames_bytes, digests_in,
	                 select);
}

// Host-buffer pack.  The batch is cut into slices of about g_zg_pack_slice_bytes of input (in file order) that
// go through two staging sets: while the kernels of slice k run on the context's stream, slice k+1
// is already on its way up (s_in) and the frames of slice k-1 are on their way down (s_out).  The
// archive state (dedup map, running offset) carries from slice to slice exactly as it does from call
// to call, so the results equal one pack_core over the whole batch.
static size_t pack_host(zg_cctx* c, ZgArchive& A, const uint8_t* blob, const uint64_t* off, const uint64_t* len, uint64_t n,
                        uint8_t* digests, uint8_t* first, uint64_t* frame_off, uint64_t* frame_len, uint8_t* frames_out,
                        uint64_t frames_cap, uint64_t* frames_bytes) {
	cudaStream_t s = c->stream;
	if (frames_bytes) *frames_bytes = 0;
	if (n == 0) return 0;
	if (!c->s_in) {
		ZG_CUDA(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
		ZG_CUDA(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
		for (auto& st : c->stage) {
			ZG_CUDA(cudaEventCreateWithFlags(&st.in_done, cudaEventDisableTiming));
			ZG_CUDA(cudaEventCreateWithFlags(&st.out_done, cudaEventDisableTiming));
		}
	}
	// slices: [first file, one past last), and the span of the host blob each one touches
	struct Slice {
		u64 i0, i1, lo, hi, bytes;
	};
	std::vector<Slice> sl;
	{
		Slice cur{0, 0, ~0ull, 0, 0};
		u64 spans = 0, total = 0;
		for (u64 i = 0; i < n; i++) {
			cur.lo = off[i] < cur.lo ? off[i] : cur.lo;
			cur.hi = off[i] + len[i] > cur.hi ? off[i] + len[i] : cur.hi;
			cur.bytes += len[i];
			// the first slice is a quarter of the others: its upload is the one transfer nothing overlaps
			if (cur.bytes >= (sl.empty() ? g_zg_pack_slice_bytes / 4 : g_zg_pack_slice_bytes) || i + 1 == n) {
				cur.i1 = i + 1;
				sl.push_back(cur);
				spans += cur.hi - cur.lo;
				total += cur.bytes;
				cur = Slice{i + 1, 0, ~0ull, 0, 0};
			}
		}
		if (sl.size() > 1 && spans > 2 * total + (64ull << 20)) {  // files scattered over the blob: slicing would re-copy it
			Slice all{0, n, ~0ull, 0, total};
			for (u64 i = 0; i < n; i++) {
				all.lo = off[i] < all.lo ? off[i] : all.lo;
				all.hi = off[i] + len[i] > all.hi ? off[i] + len[i] : all.hi;
			}
			sl.assign(1, all);
		}
	}
	auto upload = [&](size_t k) -> size_t {
		const Slice& q = sl[k];
		auto& st = c->stage[k & 1];
		u64 m = q.i1 - q.i0, span = q.hi - q.lo;
		u64 cap = q.bytes + q.bytes / 128 + 64 * m + 4096;
		ZG_ALLOC(st.d_blob.reserve(span + 64));
		ZG_ALLOC(st.d_meta.reserve(m * 16));
		ZG_ALLOC(st.d_out_meta.reserve(m * 49 + 64));
		ZG_ALLOC(st.d_frames.reserve(cap + 64));
		ZG_ALLOC(st.h_meta.reserve(m * 8));
		u64* h = st.h_meta.as<u64>();
		for (u64 i = 0; i < m; i++) h[i] = off[q.i0 + i] - q.lo;
		// (the stage's input buffers are free: the slice that used them has been packed.  Its RESULTS may still be on
		// their way down -- that is the packing stream's wait below, not this upload's: both directions stay busy)
		u64* dm = st.d_meta.as<u64>();
		if (span) ZG_CUDA(cudaMemcpyAsync(st.d_blob.p, blob + q.lo, span, cudaMemcpyHostToDevice, c->s_in));
		ZG_CUDA(cudaMemcpyAsync(dm, h, m * 8, cudaMemcpyHostToDevice, c->s_in));
		ZG_CUDA(cudaMemcpyAsync(dm + m, len + q.i0, m * 8, cudaMemcpyHostToDevice, c->s_in));
		ZG_CUDA(cudaEventRecord(st.in_done, c->s_in));
		return 0;
	};
	const bool trace = getenv("ZG_TRACE") != nullptr;
	auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	double t_begin = now_ms();
	const u64 nfiles0 = A.nfiles, offset0 = A.offset;  // the call is all or nothing, however many slices went through
	size_t r = upload(0);
	u64 written = 0;
	for (size_t k = 0; k < sl.size() && !zg_is_error(r); k++) {
		const Slice& q = sl[k];
		auto& st = c->stage[k & 1];
		u64 m = q.i1 - q.i0;
		double t0 = now_ms();
		if (k + 1 < sl.size()) {
			r = upload(k + 1);
			if (zg_is_error(r)) break;
		}
		double t1 = now_ms();
		if (trace) cudaStreamSynchronize(c->s_in), fprintf(stderr, "[zg pack] slice %zu: files %llu bytes %llu  upload(next) enqueue %.2f ms, in-stream drained at +%.2f ms\n", k, (unsigned long long)m, (unsigned long long)q.bytes, t1 - t0, now_ms() - t_begin);
		u8* o = st.d_out_meta.as<u8>();
		u64* d_foff = (u64*)o;
		u64* d_flen = d_foff + m;
		u8* d_dig = (u8*)(d_flen + m);
		u8* d_first = d_dig + m * 32;
		u64* dm = st.d_meta.as<u64>();
		u64 cap = zg_min<u64>(st.d_frames.cap, frames_cap - written);
		u64 bytes = 0;
		// the slice's input has arrived, and the stage's previous results have left
		if (cudaStreamWaitEvent(s, st.in_done, 0) != cudaSuccess || cudaStreamWaitEvent(s, st.out_done, 0) != cudaSuccess) {
			r = ZG_ERR(ZG_error_device);
			break;
		}
		double t2 = now_ms();
		r = pack_core(c, A, st.d_blob.as<u8>(), dm, dm + m, m, d_dig, d_first, d_foff, d_flen, st.d_frames.as<u8>(), cap, &bytes);
		if (trace) fprintf(stderr, "[zg pack] slice %zu: core %.2f ms (ends at +%.2f ms)\n", k, now_ms() - t2, now_ms() - t_begin);
		if (zg_is_error(r)) break;
		// pack_core returns with the stream drained: the results can leave while the next slice computes
		cudaError_t e = cudaSuccess;
		if (digests) e = cudaMemcpyAsync(digests + 32 * q.i0, d_dig, m * 32, cudaMemcpyDeviceToHost, c->s_out);
		if (first && e == cudaSuccess) e = cudaMemcpyAsync(first + q.i0, d_first, m, cudaMemcpyDeviceToHost, c->s_out);
		if (frame_off && e == cudaSuccess) e = cudaMemcpyAsync(frame_off + q.i0, d_foff, m * 8, cudaMemcpyDeviceToHost, c->s_out);
		if (frame_len && e == cudaSuccess) e = cudaMemcpyAsync(frame_len + q.i0, d_flen, m * 8, cudaMemcpyDeviceToHost, c->s_out);
		if (bytes && e == cudaSuccess) e = cudaMemcpyAsync(frames_out + written, st.d_frames.p, bytes, cudaMemcpyDeviceToHost, c->s_out);
		if (e == cudaSuccess) e = cudaEventRecord(st.out_done, c->s_out);
		if (e != cudaSuccess) {
			r = ZG_ERR(ZG_error_device);
			break;
		}
		written += bytes;
	}
	cudaError_t e1 = cudaStreamSynchronize(c->s_in), e2 = cudaStreamSynchronize(c->s_out);
	if (trace) fprintf(stderr, "[zg pack] all results on the host at +%.2f ms\n", now_ms() - t_begin);
	if (!zg_is_error(r) && (e1 != cudaSuccess || e2 != cudaSuccess)) r = ZG_ERR(ZG_error_device);
	if (zg_is_error(r)) {
		archive_rollback(c, A, nfiles0, offset0);
		return r;
	}
	if (frames_bytes) *frames_bytes = written;
	return 0;
}

size_t zg_pack_batch(zg_cctx* c, const uint8_t* blob, const uint64_t* off, const uint64_t* len, uint64_t n, uint8_t* digests,
                     uint8_t* first, uint64_t* frame_off, uint64_t* frame_len, uint8_t* frames_out, uint64_t frames_cap,
                     uint64_t* frames_bytes) {
	ZG_NEED_DEVICE();
	if (!c) return ZG_ERR(ZG_error_GENERIC);
	ZG_GUARD(pack_host(c, c->archive, blob, off, len, n, digests, first, frame_off, frame_len, frames_out, frames_cap, frames_bytes));
}

// CCtx::compress2 (lowlevel_frames.rs:30): one complete frame, no dedup, no archive state.
size_t zg_compress2(zg_cctx* c, void* dst, size_t cap, const void* src, size_t n) {
	ZG_NEED_DEVICE();
	if (!c) return ZG_ERR(ZG_error_GENERIC);
	if (!dst) return ZG_ERR(ZG_error_dstBuffer_null);
	ZgArchive& A = c->oneshot;
	cudaStreamSynchronize(c->stream);
	A.nfiles = 0;
	A.offset = 0;
	if (A.table_size) ZG_CUDA(cudaMemsetAsync(A.table.p, 0, (size_t)A.table_size * 4, c->stream));
	uint64_t off = 0, len = n, bytes = 0;
	static const uint8_t empty = 0;
	size_t r;
	try {
		r = pack_host(c, A, src ? (const uint8_t*)src : &empty, &off, &len, 1, nullptr, nullptr, nullptr, nullptr, (uint8_t*)dst, cap, &bytes);
	} catch (...) {
		return ZG_ERR(ZG_error_memory_allocation);
	}
	if (zg_is_error(r)) return r;
	return bytes;
}

}  // extern "C"

// First-occurrence decisions over an ordered list of digests (content_frame.rs:30), on the device.
// Used when digests from several GPUs are gathered into global file order.
extern "C" size_t zg_dedup_dev(void* stream, const uint8_t* digests, uint64_t n, uint8_t* first, uint64_t* rep) {
	ZG_NEED_DEVICE();
	if (n == 0) return 0;
	cudaStream_t s = (cudaStream_t)stream;
	u64 want = 1024;
	while (want < 2 * n) want <<= 1;
	// the table and the scratch are kept per device between calls: cudaMalloc / cudaFree synchronise the whole device,
	// and beside another process's collectives a cudaFree per call was seen to stall for 0.5-0.8 s now and then
	struct Scratch {
		ZgBuf table, tmp;
	};
	static std::mutex mu;
	static Scratch cache[64];
	int dev = 0;
	cudaGetDevice(&dev);
	Scratch local;
	const bool cached = dev >= 0 && dev < 64;
	std::unique_lock<std::mutex> g(mu, std::defer_lock);
	if (cached) g.lock();
	Scratch& S = cached ? cache[dev] : local;
	size_t r = 0;
	if (S.table.reserve(want * 4) || S.tmp.reserve(n * 8 * 4)) r = ZG_ERR(ZG_error_memory_allocation);
	if (!r) {
		cudaMemsetAsync(S.table.p, 0, want * 4, s);
		cudaMemsetAsync(S.tmp.p, 0, n * 8 * 4, s);
		u64* t = S.tmp.as<u64>();
		r = zg_pk_dedup_insert(s, digests, 0, n, S.table.as<u32>(), (u32)want - 1);
		if (!r) r = zg_pk_dedup_resolve(s, digests, 0, n, S.table.as<u32>(), (u32)want - 1, t, nullptr, rep, first, t + n, t + 2 * n, t + 3 * n);
		if (cudaStreamSynchronize(s) != cudaSuccess) r = ZG_ERR(ZG_error_device);
	}
	if (!cached) {
		local.table.release();
		local.tmp.release();
	}
	return r;
}
