// C ABI of libzarcgpu (include/zarcgpu.h): contexts, host<->device staging, error names.
// Everything that touches content bytes is a CUDA kernel; the host code here only moves buffers
// and does bookkeeping on sizes/offsets (frame/block *header* walks to find frame boundaries).
#include "common.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <mutex>
#include <thread>
#include <new>
#include <vector>
#include <string.h>

std::atomic<uint64_t> g_zg_launches{0};
// Slices of the host-buffer API.  The decoder wants many frames in flight (a slice is one wave of its lane-per-frame
// kernel), the results' way down wants to start early and end with a short last slice: with two slices decoding at once
// 1024 / 768 / 640 MiB of output per slice unpack 10.2 GB in 260 / 251 / 247 ms.  The encoder's kernels keep their
// efficiency on less, and smaller pack slices expose less of the first upload.
#define ZG_UNPACK_SLICE (640ull << 20)
#define ZG_PACK_SLICE (768ull << 20)
uint64_t g_zg_slice_bytes = ZG_UNPACK_SLICE;
uint64_t g_zg_pack_slice_bytes = ZG_PACK_SLICE;
extern "C" void zg_internal_set_slice_bytes(uint64_t v) {
	g_zg_slice_bytes = v ? v : ZG_UNPACK_SLICE;
	g_zg_pack_slice_bytes = v ? v : ZG_PACK_SLICE;
}
extern "C" void zg_internal_set_pack_slice_bytes(uint64_t v) { g_zg_pack_slice_bytes = v ? v : ZG_PACK_SLICE; }

static int g_dev_count = -2;
static int g_sm_count = 0;

static int dev_count() {
	if (g_dev_count == -2) {
		int n = 0;
		if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0;
		g_dev_count = n;
	}
	return g_dev_count;
}
int zg_abi_dev_count() { return dev_count(); }
int zg_sm_count() {
	if (!g_sm_count) {
		int d = 0;
		cudaGetDevice(&d);
		cudaDeviceProp p;
		if (cudaGetDeviceProperties(&p, d) == cudaSuccess) g_sm_count = p.multiProcessorCount;
		if (g_sm_count <= 0) g_sm_count = 148;
	}
	return g_sm_count;
}
#define ZG_NEED_DEVICE() \
	do {                 \
		if (dev_count() <= 0) return ZG_ERR(ZG_error_no_device); \
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
// no exception crosses the C boundary: host allocations sized by caller- or archive-controlled numbers fail as error codes
#define ZG_GUARD(expr)                                    \
	try {                                                 \
		return (expr);                                    \
	} catch (const std::bad_alloc&) {                     \
		return ZG_ERR(ZG_error_memory_allocation);        \
	} catch (...) {                                       \
		return ZG_ERR(ZG_error_GENERIC);                  \
	}

// ---------------------------------------------------------------------------------------------
// host-side frame boundary walk: frame header + 3-byte block headers only (no content decoding).
// Returns 0 if more input is needed, an error, or the frame's total length.  *bound = upper bound
// on the decoded size (FCS when present, else 128 KiB per block).
static size_t host_frame_size(const uint8_t* p, size_t n, uint64_t* bound, bool* has_fcs = nullptr) {
	if (n < 5) return 0;
	uint32_t magic = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
	if (magic != 0xFD2FB528u) return ZG_ERR(ZG_error_prefix_unknown);
	uint32_t desc = p[4];
	uint32_t fcs_flag = desc >> 6, single = (desc >> 5) & 1, checksum = (desc >> 2) & 1, did_flag = desc & 3;
	if (desc & 8) return ZG_ERR(ZG_error_frameParameter_unsupported);
	size_t ip = 5 + (single ? 0 : 1) + (did_flag == 3 ? 4 : did_flag);
	uint32_t fcs_len = fcs_flag == 0 ? single : fcs_flag == 1 ? 2 : fcs_flag == 2 ? 4 : 8;
	if (n < ip + fcs_len) return 0;
	uint64_t fcs = 0;
	for (uint32_t i = 0; i < fcs_len; i++) fcs |= (uint64_t)p[ip + i] << (8 * i);
	if (fcs_len == 2) fcs += 256;
	ip += fcs_len;
	uint64_t blocks = 0;
	for (;;) {
		if (n < ip + 3) return 0;
		uint32_t bh = (uint32_t)p[ip] | ((uint32_t)p[ip + 1] << 8) | ((uint32_t)p[ip + 2] << 16);
		uint32_t last = bh & 1, type = (bh >> 1) & 3, bsize = bh >> 3;
		if (type == 3) return ZG_ERR(ZG_error_corruption_detected);
		ip += 3 + (type == 1 ? 1 : bsize);
		blocks++;
		if (last) break;
	}
	if (checksum) ip += 4;
	if (n < ip) return 0;
	// a frame cannot regenerate more than 128 KiB per block: a Frame_Content_Size beyond that is a lie (and would
	// otherwise size host and device buffers from an untrusted 64-bit field)
	if (fcs_len && fcs > blocks * 131072ull) return ZG_ERR(ZG_error_corruption_detected);
	if (bound) *bound = fcs_len ? fcs : blocks * 131072ull;
	if (has_fcs) *has_fcs = fcs_len != 0;
	return ip;
}

// ---------------------------------------------------------------------------------------------
#define ZG_UNPACK_WORKERS_MAX 4
int g_zg_unpack_workers = 2;
extern "C" void zg_internal_set_unpack_workers(int v) { g_zg_unpack_workers = v < 1 ? 2 : v > ZG_UNPACK_WORKERS_MAX ? ZG_UNPACK_WORKERS_MAX : v; }
struct zg_dctx {
	cudaStream_t stream = 0;
	bool own_stream = false;
	int verify_checksum = 1;
	ZgZdWork zd;
	ZgB3Work b3;
	ZgBuf status, produced, cksums, got_digests, first, tiles, packed_off, vspan;
	// host-API staging: two stages per worker, so that the copies of one slice overlap the kernels of another.  Slices
	// are decoded by ZG_UNPACK_WORKERS_MAX contexts at once (this one and its helpers, a host thread each): the
	// persistent decode kernel of one slice drains while the next slice's fills the SMs it leaves.
	struct Stage {
		ZgBuf d_archive, d_meta, d_out, d_digests, d_ok;
		ZgHostBuf h_small;
		cudaEvent_t in_done = nullptr, out_done = nullptr;
	} hstage[2 * ZG_UNPACK_WORKERS_MAX];
	zg_dctx* helper[ZG_UNPACK_WORKERS_MAX - 1] = {};
	cudaStream_t s_in = nullptr, s_out = nullptr;
	ZgBuf d_archive, d_out, d_meta;  // one-shot / streaming paths
	ZgHostBuf h_first, h_small;
	// streaming state (zg_decompress_stream): the frame is collected and decoded ON THE DEVICE; the host keeps only
	// the header walk's few bytes of state, so a multi-GiB frame costs no host memory (the reference streams too)
	struct Stream {
		ZgBuf in, out;             // device: the frame so far / the decoded frame
		uint64_t in_len = 0;       // frame bytes received
		uint8_t carry[24];         // the last bytes received (a 3-byte block header or the frame header may straddle two calls)
		uint32_t ncarry = 0;
		bool have_header = false;
		uint32_t checksum = 0, fcs_len = 0;
		uint64_t fcs = 0, blocks = 0;
		uint64_t next_hdr = 0;     // frame offset of the next block header to read
		uint64_t frame_end = 0;    // frame length, known once the last block's header has been seen (else 0)
		uint64_t out_len = 0, out_pos = 0;
		int stage = 0;             // 0 collecting input, 1 handing out output
		void reset() {
			in_len = ncarry = checksum = fcs_len = 0;
			have_header = false;
			fcs = blocks = next_hdr = frame_end = out_len = out_pos = 0;
			stage = 0;
		}
	} sm;
};

// device-pointer core of unpack: decode, verify sizes/checksums, optional BLAKE3 verify, first error
static size_t unpack_core(zg_dctx* d, const u8* archive, u64 archive_len, u64 n, const u64* off, const u64* len, const u64* ulen,
                          const u8* digests, u8* out, u64 out_cap, const u64* out_off, u8* ok, u32* status) {
	cudaStream_t s = d->stream;
	if (n == 0) return 0;
	ZG_ALLOC(d->produced.reserve(n * 8));
	ZG_ALLOC(d->cksums.reserve(n * 8));
	ZG_ALLOC(d->first.reserve(8));
	ZG_ALLOC(d->h_first.reserve(8));
	u32* st = status;
	if (!st) {
		ZG_ALLOC(d->status.reserve(n * 4));
		st = d->status.as<u32>();
	}
	if (!out_off) {
		ZG_ALLOC(d->packed_off.reserve(n * 8));
		ZG_TRY(zg_scan_run(s, d->tiles, ulen, n, 0, d->packed_off.as<u64>(), nullptr));
		out_off = d->packed_off.as<u64>();
	}
	ZG_TRY(zg_zstd_decode_run(s, d->zd, archive, archive_len, off, len, ulen, out_off, n, out, out_cap, st, d->produced.as<u64>(),
	                          d->cksums.as<u32>(), d->verify_checksum));
	ZG_TRY(zg_unpack_finalize_run(s, out, out_off, ulen, d->produced.as<u64>(), d->cksums.as<u32>(), st, n, d->verify_checksum, &d->zd));
	if (digests && ok) {
		ZG_ALLOC(d->got_digests.reserve(n * 32));
		ZG_ALLOC(d->vspan.reserve(n * 16));
		// only frames that decoded are hashed: a rejected frame's output span may lie outside `out`
		u64* voff = d->vspan.as<u64>();
		ZG_TRY(zg_verify_spans_run(s, st, out_off, ulen, out_cap, n, voff, voff + n));
		ZG_TRY(zg_blake3_run(s, d->b3, out, voff, voff + n, n, d->got_digests.as<u8>()));
		ZG_TRY(zg_digest_compare_run(s, d->got_digests.as<u8>(), digests, st, ok, n));
	}
	ZG_TRY(zg_first_error_run(s, st, n, d->first.as<u64>()));
	u64* hf = d->h_first.as<u64>();
	ZG_CUDA(zg_publish(s, d->first.p, hf, 8));
	ZG_CUDA(cudaStreamSynchronize(s));
	ZG_CUDA(cudaGetLastError());
	if (*hf != ~0ull) return ZG_ERR((size_t)(*hf & 0xff));
	return 0;
}

extern "C" {

int zg_is_error(size_t code) { return code > ZG_ERR(ZG_error_maxCode); }
zg_error_code zg_get_error_code(size_t code) { return zg_is_error(code) ? (zg_error_code)(0 - code) : ZG_error_no_error; }
const char* zg_error_name(size_t code) {
	// libzstd 1.5.5's strings for the shared codes (what zstd_safe::get_error_name returns)
	switch (zg_get_error_code(code)) {
	case ZG_error_no_error: return "No error detected";
	case ZG_error_GENERIC: return "Error (generic)";
	case ZG_error_prefix_unknown: return "Unknown frame descriptor";
	case ZG_error_frameParameter_unsupported: return "Unsupported frame parameter";
	case ZG_error_frameParameter_windowTooLarge: return "Frame requires too much memory for decoding";
	case ZG_error_corruption_detected: return "Data corruption detected";
	case ZG_error_checksum_wrong: return "Restored data doesn't match checksum";
	case ZG_error_dictionary_wrong: return "Dictionary mismatch";
	case ZG_error_parameter_unsupported: return "Unsupported parameter";
	case ZG_error_parameter_outOfBound: return "Parameter is out of bound";
	case ZG_error_memory_allocation: return "Allocation error : not enough memory";
	case ZG_error_dstSize_tooSmall: return "Destination buffer is too small";
	case ZG_error_srcSize_wrong: return "Src size is incorrect";
	case ZG_error_dstBuffer_null: return "Operation on NULL destination buffer";
	case ZG_error_device: return "CUDA device error";
	case ZG_error_no_device: return "No CUDA device (libzarcgpu has no CPU fallback)";
	default: return "Unspecified error code";
	}
}

int zg_device_count(void) { return dev_count(); }
size_t zg_set_device(int d) {
	ZG_NEED_DEVICE();
	g_sm_count = 0;
	return cudaSetDevice(d) == cudaSuccess ? 0 : ZG_ERR(ZG_error_device);
}
const char* zg_build_info(void) {
#ifdef ZG_EMU
	return "simt-emu (test build, not the product)";
#else
	return "sm_100a";
#endif
}
uint64_t zg_kernel_launch_count(void) { return g_zg_launches.load(); }
void* zg_alloc_pinned(size_t n) {
	void* p = nullptr;
	if (dev_count() <= 0 || cudaMallocHost(&p, n ? n : 1) != cudaSuccess) return nullptr;
	return p;
}
void zg_free_pinned(void* p) {
	if (p) cudaFreeHost(p);
}

// ---------------------------------------------------------------------------------------------
// building blocks, device pointers
size_t zg_blake3_batch_dev(void* stream, const uint8_t* blob, const uint64_t* off, const uint64_t* len, uint64_t n, uint8_t* digests) {
	ZG_NEED_DEVICE();
	// scratch (unit lists, scan tiles) is kept per device between calls: allocating it costs more than the kernels
	static std::mutex mu;
	static ZgB3Work cache[16];
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 16) {
		ZgB3Work w;
		size_t r = zg_blake3_run((cudaStream_t)stream, w, blob, off, len, n, digests);
		cudaStreamSynchronize((cudaStream_t)stream);
		zg_b3work_free(w);
		return r;
	}
	std::lock_guard<std::mutex> g(mu);
	size_t r = zg_blake3_run((cudaStream_t)stream, cache[dev], blob, off, len, n, digests);
	cudaStreamSynchronize((cudaStream_t)stream);
	return r;
}
size_t zg_xxh64_batch_dev(void* stream, const uint8_t* blob, const uint64_t* off, const uint64_t* len, uint64_t n, uint64_t* hashes) {
	ZG_NEED_DEVICE();
	return zg_xxh64_run((cudaStream_t)stream, blob, off, len, n, hashes);
}
size_t zg_corpus_generate_dev(void* stream, uint8_t* out, const uint64_t* seg_off, const uint32_t* seg_len,
                              const uint8_t* seg_kind, const uint64_t* seg_key, uint64_t n) {
	ZG_NEED_DEVICE();
	return zg_corpus_run((cudaStream_t)stream, out, seg_off, seg_len, seg_kind, seg_key, n);
}
size_t zg_assign_offsets_dev(void* stream, const uint64_t* frame_len, uint64_t n, uint64_t base, uint64_t* frame_off) {
	ZG_NEED_DEVICE();
	// (scan scratch kept per device: no cudaMalloc / cudaFree, which synchronise the device, per call)
	static std::mutex mu;
	static ZgBuf cache[64];
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64) {
		ZgBuf tiles;
		size_t r = zg_scan_run((cudaStream_t)stream, tiles, frame_len, n, base, frame_off, nullptr);
		cudaStreamSynchronize((cudaStream_t)stream);
		tiles.release();
		return r;
	}
	std::lock_guard<std::mutex> g(mu);
	size_t r = zg_scan_run((cudaStream_t)stream, cache[dev], frame_len, n, base, frame_off, nullptr);
	cudaStreamSynchronize((cudaStream_t)stream);
	return r;
}

// ---------------------------------------------------------------------------------------------
// host-buffer convenience wrappers: stage to the device, run the kernels, copy results back
struct Staged {
	ZgBuf blob, off, len;
	size_t put(cudaStream_t s, const uint8_t* b, const uint64_t* o, const uint64_t* l, uint64_t n) {
		uint64_t end = 0;
		for (uint64_t i = 0; i < n; i++) end = o[i] + l[i] > end ? o[i] + l[i] : end;
		if (blob.reserve(end + 16) || off.reserve(n * 8 + 8) || len.reserve(n * 8 + 8)) return ZG_ERR(ZG_error_memory_allocation);
		if (end) cudaMemcpyAsync(blob.p, b, end, cudaMemcpyHostToDevice, s);
		if (n) {
			cudaMemcpyAsync(off.p, o, n * 8, cudaMemcpyHostToDevice, s);
			cudaMemcpyAsync(len.p, l, n * 8, cudaMemcpyHostToDevice, s);
		}
		return 0;
	}
	void release() {
		blob.release();
		off.release();
		len.release();
	}
};

size_t zg_blake3_batch(const uint8_t* blob, const uint64_t* off, const uint64_t* len, uint64_t n, uint8_t* digests) {
	ZG_NEED_DEVICE();
	Staged st;
	ZgBuf out;
	ZgB3Work w;
	size_t r = st.put(0, blob, off, len, n);
	if (!r && out.reserve(n * 32 + 32)) r = ZG_ERR(ZG_error_memory_allocation);
	if (!r) r = zg_blake3_run(0, w, st.blob.as<u8>(), st.off.as<u64>(), st.len.as<u64>(), n, out.as<u8>());
	if (!r && n) {
		if (cudaMemcpy(digests, out.p, n * 32, cudaMemcpyDeviceToHost) != cudaSuccess) r = ZG_ERR(ZG_error_device);
	}
	st.release();
	out.release();
	zg_b3work_free(w);
	return r;
}
size_t zg_blake3(const void* data, size_t len, uint8_t out[32]) {
	uint64_t o = 0, l = len;
	return zg_blake3_batch((const uint8_t*)data, &o, &l, 1, out);
}
size_t zg_xxh64_batch(const uint8_t* blob, const uint64_t* off, const uint64_t* len, uint64_t n, uint64_t* hashes) {
	ZG_NEED_DEVICE();
	Staged st;
	ZgBuf out;
	size_t r = st.put(0, blob, off, len, n);
	if (!r && out.reserve(n * 8 + 8)) r = ZG_ERR(ZG_error_memory_allocation);
	if (!r) r = zg_xxh64_run(0, st.blob.as<u8>(), st.off.as<u64>(), st.len.as<u64>(), n, out.as<u64>());
	if (!r && n) {
		if (cudaMemcpy(hashes, out.p, n * 8, cudaMemcpyDeviceToHost) != cudaSuccess) r = ZG_ERR(ZG_error_device);
	}
	st.release();
	out.release();
	return r;
}

// ---------------------------------------------------------------------------------------------
// streaming Hasher: blake3::Hasher::{new,update,finalize} (decode/frame_iterator.rs:54,99,77).
// update() buffers; finalize() hashes the buffered bytes on the GPU (the digest of a stream is
// only observable at finalize, so this is semantically identical).
struct zg_hasher {
	ZgBuf dev, meta, out;   // the content so far lives ON THE DEVICE (the host keeps nothing)
	ZgB3Work b3;
	uint64_t len = 0;
};
zg_hasher* zg_hasher_new(void) {
	if (dev_count() <= 0) return nullptr;
	return new (std::nothrow) zg_hasher();
}
size_t zg_hasher_update(zg_hasher* h, const void* data, size_t len) {
	ZG_NEED_DEVICE();
	if (!h || (!data && len)) return ZG_ERR(ZG_error_GENERIC);
	if (!len) return 0;
	if (h->len + len > h->dev.cap) {
		ZgBuf nb;
		size_t want = (h->len + len) + (h->len + len) / 2 + 4096;
		ZG_ALLOC(nb.reserve(want));
		if (h->len) ZG_CUDA(cudaMemcpy(nb.p, h->dev.p, h->len, cudaMemcpyDeviceToDevice));
		h->dev.release();
		h->dev = nb;
	}
	ZG_CUDA(cudaMemcpy(h->dev.as<u8>() + h->len, data, len, cudaMemcpyHostToDevice));
	h->len += len;
	return 0;
}
size_t zg_hasher_finalize(zg_hasher* h, uint8_t out[32]) {
	ZG_NEED_DEVICE();
	if (!h) return ZG_ERR(ZG_error_GENERIC);
	ZG_ALLOC(h->dev.reserve(16));
	ZG_ALLOC(h->meta.reserve(16));
	ZG_ALLOC(h->out.reserve(32));
	uint64_t m[2] = {0, h->len};
	ZG_CUDA(cudaMemcpy(h->meta.p, m, 16, cudaMemcpyHostToDevice));
	ZG_TRY(zg_blake3_run(0, h->b3, h->dev.as<u8>(), h->meta.as<u64>(), h->meta.as<u64>() + 1, 1, h->out.as<u8>()));
	ZG_CUDA(cudaMemcpy(out, h->out.p, 32, cudaMemcpyDeviceToHost));
	return 0;
}
void zg_hasher_free(zg_hasher* h) {
	if (!h) return;
	h->dev.release();
	h->meta.release();
	h->out.release();
	zg_b3work_free(h->b3);
	delete h;
}

// ---------------------------------------------------------------------------------------------
// decompression context
zg_dctx* zg_dctx_create(void) {
	if (dev_count() <= 0) return nullptr;
	zg_dctx* d = new (std::nothrow) zg_dctx();
	if (!d) return nullptr;
	if (cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess) {
		delete d;
		return nullptr;
	}
	d->own_stream = true;
	return d;
}
void zg_dctx_free(zg_dctx* d) {
	if (!d) return;
	for (zg_dctx*& h : d->helper) {
		zg_dctx_free(h);
		h = nullptr;
	}
	cudaStreamSynchronize(d->stream);
	d->zd.release();
	zg_b3work_free(d->b3);
	for (ZgBuf* b : {&d->status, &d->produced, &d->cksums, &d->got_digests, &d->first, &d->tiles, &d->packed_off, &d->vspan, &d->sm.in, &d->sm.out, &d->d_archive,
	                 &d->d_meta, &d->d_out})
		b->release();
	for (auto& st : d->hstage) {
		for (ZgBuf* b : {&st.d_archive, &st.d_meta, &st.d_out, &st.d_digests, &st.d_ok}) b->release();
		st.h_small.release();
		if (st.in_done) cudaEventDestroy(st.in_done);
		if (st.out_done) cudaEventDestroy(st.out_done);
	}
	if (d->s_in) cudaStreamDestroy(d->s_in);
	if (d->s_out) cudaStreamDestroy(d->s_out);
	d->h_first.release();
	d->h_small.release();
	if (d->own_stream) cudaStreamDestroy(d->stream);
	delete d;
}
size_t zg_dctx_set_stream(zg_dctx* d, void* stream) {
	if (!d) return ZG_ERR(ZG_error_GENERIC);
	if (d->own_stream) cudaStreamDestroy(d->stream);
	d->own_stream = false;
	d->stream = (cudaStream_t)stream;
	return 0;
}
size_t zg_dctx_set_verify_checksum(zg_dctx* d, int on) {
	if (!d) return ZG_ERR(ZG_error_GENERIC);
	d->verify_checksum = on ? 1 : 0;
	return 0;
}
size_t zg_dstream_in_size(void) { return 131075; }   // ZSTD_DStreamInSize: block max + block header
size_t zg_dstream_out_size(void) { return 131072; }  // ZSTD_DStreamOutSize
size_t zg_find_frame_compressed_size(const void* src, size_t n) {
	size_t r = host_frame_size((const uint8_t*)src, n, nullptr);
	return r == 0 ? ZG_ERR(ZG_error_srcSize_wrong) : r;
}

size_t zg_unpack_batch_dev(zg_dctx* d, const uint8_t* archive, uint64_t archive_len, uint64_t n, const uint64_t* off,
                           const uint64_t* len, const uint64_t* ulen, const uint8_t* digests, uint8_t* out, uint64_t out_cap,
                           const uint64_t* out_off, uint8_t* ok, uint32_t* status) {
	ZG_NEED_DEVICE();
	if (!d) return ZG_ERR(ZG_error_GENERIC);
	if (!out && n) return ZG_ERR(ZG_error_dstBuffer_null);
	return unpack_core(d, archive, archive_len, n, off, len, ulen, digests, out, out_cap, out_off, ok, status);
}

// Host-buffer unpack, sliced and double-buffered like pack_host (abi_pack.cu): the archive bytes of
// slice k+1 go up and the restored files of slice k-1 come down while slice k decodes.
static size_t unpack_batch_impl(zg_dctx* d, const uint8_t* archive, uint64_t archive_len, uint64_t n, const uint64_t* off,
                                const uint64_t* len, const uint64_t* ulen, const uint8_t* digests, uint8_t* out, uint64_t out_cap,
                                const uint64_t* out_off, uint8_t* ok, uint32_t* status) {
	ZG_NEED_DEVICE();
	if (!d) return ZG_ERR(ZG_error_GENERIC);
	if (n == 0) return 0;
	if (!out) return ZG_ERR(ZG_error_dstBuffer_null);
	cudaStream_t s = d->stream;
	if (!d->s_in) {
		ZG_CUDA(cudaStreamCreateWithFlags(&d->s_in, cudaStreamNonBlocking));
		ZG_CUDA(cudaStreamCreateWithFlags(&d->s_out, cudaStreamNonBlocking));
		for (auto& st : d->hstage) {
			ZG_CUDA(cudaEventCreateWithFlags(&st.in_done, cudaEventDisableTiming));
			ZG_CUDA(cudaEventCreateWithFlags(&st.out_done, cudaEventDisableTiming));
		}
	}
	// bookkeeping on offsets: is the output the dense concatenation (then it can be sliced)?
	u64 total = 0;
	bool dense = true;
	for (u64 k = 0; k < n; k++) {
		if (off[k] > archive_len || len[k] > archive_len - off[k]) return ZG_ERR(ZG_error_srcSize_wrong);
		if (out_off && out_off[k] != total) dense = false;
		if (ulen[k] > UINT64_MAX - total) return ZG_ERR(ZG_error_dstSize_tooSmall);  // (the sum must not wrap)
		total += ulen[k];
	}
	if (dense && total > out_cap) return ZG_ERR(ZG_error_dstSize_tooSmall);
	struct Slice {
		u64 i0, i1, lo, hi, obase, obytes, olo, ohi;
	};
	std::vector<Slice> sl;
	{
		Slice cur{0, 0, ~0ull, 0, 0, 0, ~0ull, 0};
		u64 spans = 0, cbytes = 0, acc = 0;
		for (u64 k = 0; k < n; k++) {
			cur.lo = off[k] < cur.lo ? off[k] : cur.lo;
			cur.hi = off[k] + len[k] > cur.hi ? off[k] + len[k] : cur.hi;
			cur.obytes += ulen[k];
			cbytes += len[k];
			if (!dense) {
				if (out_off[k] > out_cap || ulen[k] > out_cap - out_off[k]) return ZG_ERR(ZG_error_dstSize_tooSmall);
				cur.olo = out_off[k] < cur.olo ? out_off[k] : cur.olo;
				cur.ohi = out_off[k] + ulen[k] > cur.ohi ? out_off[k] + ulen[k] : cur.ohi;
			}
			acc += ulen[k];
			if ((dense && cur.obytes >= g_zg_slice_bytes) || k + 1 == n) {
				cur.i1 = k + 1;
				if (dense) {
					cur.olo = cur.obase;
					cur.ohi = cur.obase + cur.obytes;
				}
				sl.push_back(cur);
				spans += cur.hi - cur.lo;
				cur = Slice{k + 1, 0, ~0ull, 0, acc, 0, ~0ull, 0};
			}
		}
		if (sl.size() > 1 && spans > 2 * cbytes + (64ull << 20)) {  // frames scattered over the archive: one slice
			Slice all{0, n, ~0ull, 0, 0, total, 0, total};
			for (u64 k = 0; k < n; k++) {
				all.lo = off[k] < all.lo ? off[k] : all.lo;
				all.hi = off[k] + len[k] > all.hi ? off[k] + len[k] : all.hi;
			}
			sl.assign(1, all);
		}
	}
	// workers: contexts decoding slices at the same time (slice k belongs to worker k % nw and to stage k % nst)
#ifdef ZG_EMU
	size_t nw = 1;  // (the CPU test build runs kernels on the calling thread)
#else
	size_t nw = (size_t)g_zg_unpack_workers < sl.size() ? (size_t)g_zg_unpack_workers : sl.size();
#endif
	int dev = 0;
	ZG_CUDA(cudaGetDevice(&dev));
	for (size_t w = 1; w < nw; w++) {
		if (!d->helper[w - 1] && !(d->helper[w - 1] = zg_dctx_create())) return ZG_ERR(ZG_error_memory_allocation);
		d->helper[w - 1]->verify_checksum = d->verify_checksum;
	}
	const size_t nst = 2 * nw;
	std::mutex up_mu;
	auto upload = [&](size_t k) -> size_t {
		std::lock_guard<std::mutex> lk(up_mu);
		const Slice& q = sl[k];
		auto& st = d->hstage[k % nst];
		u64 m = q.i1 - q.i0, span = q.hi - q.lo, ospan = q.ohi - q.olo;
		ZG_ALLOC(st.d_archive.reserve(span + 16));
		ZG_ALLOC(st.d_meta.reserve(m * 32));
		ZG_ALLOC(st.d_out.reserve(ospan + 16));
		ZG_ALLOC(st.d_ok.reserve(m * 5 + 8));
		ZG_ALLOC(st.h_small.reserve(m * 16));
		u64* h_off = st.h_small.as<u64>();
		u64* h_oo = h_off + m;
		u64 acc = 0;
		for (u64 i = 0; i < m; i++) {
			h_off[i] = off[q.i0 + i] - q.lo;
			h_oo[i] = dense ? acc : out_off[q.i0 + i] - q.olo;
			acc += ulen[q.i0 + i];
		}
		// (the stage's input buffers are free: the slice that used them has been decoded.  Its RESULTS may still be on
		// their way down -- that is the decoding stream's wait, not this upload's: both directions stay busy)
		u64* dm = st.d_meta.as<u64>();
		ZG_CUDA(cudaMemcpyAsync(st.d_archive.p, archive + q.lo, span, cudaMemcpyHostToDevice, d->s_in));
		ZG_CUDA(cudaMemcpyAsync(dm, h_off, m * 8, cudaMemcpyHostToDevice, d->s_in));
		ZG_CUDA(cudaMemcpyAsync(dm + m, len + q.i0, m * 8, cudaMemcpyHostToDevice, d->s_in));
		ZG_CUDA(cudaMemcpyAsync(dm + 2 * m, ulen + q.i0, m * 8, cudaMemcpyHostToDevice, d->s_in));
		ZG_CUDA(cudaMemcpyAsync(dm + 3 * m, h_oo, m * 8, cudaMemcpyHostToDevice, d->s_in));
		if (digests && ok) {
			ZG_ALLOC(st.d_digests.reserve(m * 32));
			ZG_CUDA(cudaMemcpyAsync(st.d_digests.p, digests + 32 * q.i0, m * 32, cudaMemcpyHostToDevice, d->s_in));
		}
		ZG_CUDA(cudaEventRecord(st.in_done, d->s_in));
		return 0;
	};
	const bool trace = getenv("ZG_TRACE") != nullptr;
	auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	double t_begin = now_ms();
	std::atomic<bool> stop{false};
	// one worker: its slices in order; a fatal error (device, memory) stops every worker, a frame's error is kept and
	// the other frames are still delivered
	struct Result {
		size_t fatal = 0, first_err = 0;
		u64 first_err_slice = ~0ull;
	};
	std::vector<Result> res(nw);
	auto work = [&](size_t w) {
		Result& R = res[w];
		zg_dctx* c = w ? d->helper[w - 1] : d;
		if (w && cudaSetDevice(dev) != cudaSuccess) {
			R.fatal = ZG_ERR(ZG_error_device);
			stop = true;
			return;
		}
		for (size_t k = w; k < sl.size() && !stop; k += nw) {
			const Slice& q = sl[k];
			auto& st = d->hstage[k % nst];
			u64 m = q.i1 - q.i0, span = q.hi - q.lo, ospan = q.ohi - q.olo;
			u64* dm = st.d_meta.as<u64>();
			u8* d_dig = (digests && ok) ? st.d_digests.as<u8>() : nullptr;
			u8* d_okp = st.d_ok.as<u8>();
			u32* d_status = (u32*)(d_okp + ((m + 3) & ~(u64)3));
			// the slice's input has arrived, and the stage's previous results have left
			if (cudaStreamWaitEvent(c->stream, st.in_done, 0) != cudaSuccess || cudaStreamWaitEvent(c->stream, st.out_done, 0) != cudaSuccess) {
				R.fatal = ZG_ERR(ZG_error_device);
				break;
			}
			double t2 = now_ms();
			size_t rk = unpack_core(c, st.d_archive.as<u8>(), span, m, dm, dm + m, dm + 2 * m, d_dig, st.d_out.as<u8>(), ospan, dm + 3 * m,
			                        d_dig ? d_okp : nullptr, d_status);
			if (trace) fprintf(stderr, "[zg unpack] slice %zu (worker %zu): frames %llu out %llu  core starts +%.2f ms, takes %.2f ms\n", k, w,
			                   (unsigned long long)m, (unsigned long long)q.obytes, t2 - t_begin, now_ms() - t2);
			if (zg_is_error(rk)) {
				if (zg_get_error_code(rk) == ZG_error_device || zg_get_error_code(rk) == ZG_error_memory_allocation) {
					R.fatal = rk;
					break;
				}
				if (!R.first_err) R.first_err = rk, R.first_err_slice = k;
			}
			cudaError_t e = cudaSuccess;
			{
				std::lock_guard<std::mutex> lk(up_mu);  // (s_out is shared: one slice's results are enqueued as a unit)
				if (!dense) {
					std::vector<u8> tmp(ospan);
					e = cudaMemcpyAsync(tmp.data(), st.d_out.p, ospan, cudaMemcpyDeviceToHost, c->stream);
					if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
					if (e == cudaSuccess)
						for (u64 i = q.i0; i < q.i1; i++) memcpy(out + out_off[i], tmp.data() + (out_off[i] - q.olo), ulen[i]);
				} else if (q.obytes) {
					e = cudaMemcpyAsync(out + q.obase, st.d_out.p, q.obytes, cudaMemcpyDeviceToHost, d->s_out);
				}
				if (d_dig && e == cudaSuccess) e = cudaMemcpyAsync(ok + q.i0, d_okp, m, cudaMemcpyDeviceToHost, d->s_out);
				if (status && e == cudaSuccess) e = cudaMemcpyAsync(status + q.i0, d_status, m * 4, cudaMemcpyDeviceToHost, d->s_out);
				if (e == cudaSuccess) e = cudaEventRecord(st.out_done, d->s_out);
			}
			if (e != cudaSuccess) {
				R.fatal = ZG_ERR(ZG_error_device);
				break;
			}
			if (k + nst < sl.size()) {  // the stage is free again once these results have left: its next slice may come up
				size_t ru = upload(k + nst);
				if (zg_is_error(ru)) {
					R.fatal = ru;
					break;
				}
			}
		}
		if (R.fatal) stop = true;
	};
	size_t r = 0;
	for (size_t k = 0; k < nst && k < sl.size() && !zg_is_error(r); k++) r = upload(k);
	if (!zg_is_error(r)) {
		struct Joiner {
			std::vector<std::thread> th;
			~Joiner() {
				for (auto& t : th)
					if (t.joinable()) t.join();
			}
		} joiner;
		for (size_t w = 1; w < nw; w++) joiner.th.emplace_back([&work, &res, &stop, w] {
			try {
				work(w);
			} catch (...) {  // (nothing may leave a thread; an allocation failure is all that can be thrown here)
				res[w].fatal = ZG_ERR(ZG_error_memory_allocation);
				stop = true;
			}
		});
		work(0);
	}
	size_t first_err = 0;
	{
		u64 best = ~0ull;
		for (const Result& R : res) {
			if (R.fatal && !zg_is_error(r)) r = R.fatal;
			if (R.first_err && R.first_err_slice < best) best = R.first_err_slice, first_err = R.first_err;
		}
	}
	cudaError_t e1 = cudaStreamSynchronize(d->s_in), e2 = cudaStreamSynchronize(d->s_out);
	if (trace) fprintf(stderr, "[zg unpack] all results on the host at +%.2f ms\n", now_ms() - t_begin);
	if (zg_is_error(r)) return r;
	if (e1 != cudaSuccess || e2 != cudaSuccess) return ZG_ERR(ZG_error_device);
	return first_err;
}
size_t zg_unpack_batch(zg_dctx* d, const uint8_t* archive, uint64_t archive_len, uint64_t n, const uint64_t* off,
                       const uint64_t* len, const uint64_t* ulen, const uint8_t* digests, uint8_t* out, uint64_t out_cap,
                       const uint64_t* out_off, uint8_t* ok, uint32_t* status) {
	ZG_GUARD(unpack_batch_impl(d, archive, archive_len, n, off, len, ulen, digests, out, out_cap, out_off, ok, status));
}

// one frame already on the device (d_frame, fsz bytes) -> d_out (ocap bytes); returns the bytes produced
static size_t decompress_dev(zg_dctx* d, const u8* d_frame, u64 fsz, u8* d_out, u64 ocap) {
	cudaStream_t s = d->stream;
	ZG_ALLOC(d->d_meta.reserve(64));
	ZG_ALLOC(d->produced.reserve(8));
	ZG_ALLOC(d->cksums.reserve(8));
	ZG_ALLOC(d->status.reserve(4));
	ZG_ALLOC(d->h_small.reserve(64));
	u64* hm = d->h_small.as<u64>();
	hm[0] = 0;
	hm[1] = fsz;
	hm[2] = ocap;
	hm[3] = 0;
	u64* m = d->d_meta.as<u64>();
	ZG_CUDA(cudaMemcpyAsync(m, hm, 32, cudaMemcpyHostToDevice, s));
	ZG_TRY(zg_zstd_decode_run(s, d->zd, d_frame, fsz, m, m + 1, m + 2, m + 3, 1, d_out, ocap, d->status.as<u32>(), d->produced.as<u64>(),
	                          d->cksums.as<u32>()));
	// size is whatever the frame produced (FCS is checked inside the kernel when present)
	ZG_CUDA(cudaMemcpyAsync(hm + 4, d->produced.p, 8, cudaMemcpyDeviceToHost, s));
	ZG_CUDA(cudaStreamSynchronize(s));
	u64 produced = hm[4];
	hm[2] = produced;
	ZG_CUDA(cudaMemcpyAsync(m + 2, hm + 2, 8, cudaMemcpyHostToDevice, s));
	ZG_TRY(zg_unpack_finalize_run(s, d_out, m + 3, m + 2, d->produced.as<u64>(), d->cksums.as<u32>(), d->status.as<u32>(), 1, d->verify_checksum));
	u32* hst = (u32*)(hm + 5);
	ZG_CUDA(cudaMemcpyAsync(hst, d->status.p, 4, cudaMemcpyDeviceToHost, s));
	ZG_CUDA(cudaStreamSynchronize(s));
	if (*hst) return ZG_ERR((size_t)*hst);
	return produced;
}

size_t zg_decompress(zg_dctx* d, void* dst, size_t cap, const void* src, size_t n) {
	ZG_NEED_DEVICE();
	if (!d) return ZG_ERR(ZG_error_GENERIC);
	uint64_t bound = 0;
	bool has_fcs = false;
	size_t fsz = host_frame_size((const uint8_t*)src, n, &bound, &has_fcs);
	if (fsz == 0) return ZG_ERR(ZG_error_srcSize_wrong);
	if (zg_is_error(fsz)) return fsz;
	if (has_fcs && bound > cap) return ZG_ERR(ZG_error_dstSize_tooSmall);
	u64 ocap = bound;
	ZG_ALLOC(d->d_archive.reserve(fsz + 16));
	ZG_ALLOC(d->d_out.reserve(ocap + 16));
	ZG_CUDA(cudaMemcpyAsync(d->d_archive.p, src, fsz, cudaMemcpyHostToDevice, d->stream));
	size_t produced = decompress_dev(d, d->d_archive.as<u8>(), fsz, d->d_out.as<u8>(), ocap);
	if (zg_is_error(produced)) return produced;
	if (produced > cap) return ZG_ERR(ZG_error_dstSize_tooSmall);
	if (produced) ZG_CUDA(cudaMemcpy(dst, d->d_out.p, produced, cudaMemcpyDeviceToHost));
	return produced;
}

// DCtx::decompress_stream contract (decode/zstd_iterator.rs:104-107,126-129): consume input until the
// frame is complete, decode it on the GPU, then hand the output out as space allows.  Returns 0
// when the frame is fully decoded and flushed, otherwise a non-zero hint.
// grow a device buffer keeping its first `keep` bytes
static cudaError_t dev_grow_keep(ZgBuf& b, size_t need, size_t keep, cudaStream_t s) {
	if (need <= b.cap) return cudaSuccess;
	ZgBuf nb;
	cudaError_t e = nb.reserve(need + need / 2);
	if (e != cudaSuccess) return e;
	if (keep && b.p) {
		e = cudaMemcpyAsync(nb.p, b.p, keep, cudaMemcpyDeviceToDevice, s);
		if (e == cudaSuccess) e = cudaStreamSynchronize(s);
		if (e != cudaSuccess) {
			nb.release();
			return e;
		}
	}
	b.release();
	b = nb;
	return cudaSuccess;
}
// byte at frame offset `pos` out of (carry, the bytes offered now); the caller only asks for offsets that are in reach
static inline uint8_t stream_byte(const zg_dctx::Stream& S, const uint8_t* src, uint64_t pos) {
	return pos >= S.in_len ? src[pos - S.in_len] : S.carry[S.ncarry - (S.in_len - pos)];
}
static size_t decompress_stream_impl(zg_dctx* d, zg_out_buffer* output, zg_in_buffer* input) {
	ZG_NEED_DEVICE();
	if (!d || !output || !input) return ZG_ERR(ZG_error_GENERIC);
	if (input->pos > input->size || output->pos > output->size) return ZG_ERR(ZG_error_srcSize_wrong);
	zg_dctx::Stream& S = d->sm;
	cudaStream_t s = d->stream;
	if (S.stage == 0) {
		const uint8_t* src = (const uint8_t*)input->src + input->pos;
		size_t avail = input->size - input->pos;
		uint64_t have = S.in_len + avail;  // frame bytes in reach: [0, have)
		// ---- header walk (frame header, then 3-byte block headers): no content byte is looked at on the host ----
		if (!S.have_header && have >= 5) {
			uint8_t hb[18];
			uint32_t nh = (uint32_t)(have < 18 ? have : 18);
			for (uint32_t i = 0; i < nh; i++) hb[i] = stream_byte(S, src, i);
			uint32_t magic = (uint32_t)hb[0] | ((uint32_t)hb[1] << 8) | ((uint32_t)hb[2] << 16) | ((uint32_t)hb[3] << 24);
			if (magic != 0xFD2FB528u) {
				S.reset();
				return ZG_ERR(ZG_error_prefix_unknown);
			}
			uint32_t desc = hb[4];
			if (desc & 8) {
				S.reset();
				return ZG_ERR(ZG_error_frameParameter_unsupported);
			}
			uint32_t fcs_flag = desc >> 6, single = (desc >> 5) & 1, did_flag = desc & 3;
			uint32_t ip = 5 + (single ? 0 : 1) + (did_flag == 3 ? 4 : did_flag);
			uint32_t fcs_len = fcs_flag == 0 ? single : fcs_flag == 1 ? 2 : fcs_flag == 2 ? 4 : 8;
			if (have >= ip + fcs_len) {
				uint64_t fcs = 0;
				for (uint32_t i = 0; i < fcs_len; i++) fcs |= (uint64_t)hb[ip + i] << (8 * i);
				if (fcs_len == 2) fcs += 256;
				S.have_header = true;
				S.checksum = (desc >> 2) & 1;
				S.fcs_len = fcs_len;
				S.fcs = fcs;
				S.next_hdr = ip + fcs_len;
			}
		}
		while (S.have_header && !S.frame_end && S.next_hdr + 3 <= have) {
			uint32_t bh = (uint32_t)stream_byte(S, src, S.next_hdr) | ((uint32_t)stream_byte(S, src, S.next_hdr + 1) << 8) |
			              ((uint32_t)stream_byte(S, src, S.next_hdr + 2) << 16);
			uint32_t last = bh & 1, type = (bh >> 1) & 3, bsize = bh >> 3;
			if (type == 3) {
				S.reset();
				return ZG_ERR(ZG_error_corruption_detected);
			}
			S.next_hdr += 3 + (type == 1 ? 1 : bsize);
			S.blocks++;
			if (last) S.frame_end = S.next_hdr + (S.checksum ? 4 : 0);
		}
		// ---- the bytes of this frame go to the device; never consume past the end of the frame ----
		size_t take = avail;
		if (S.frame_end && S.in_len + take > S.frame_end) take = (size_t)(S.frame_end - S.in_len);
		if (take) {
			if (dev_grow_keep(S.in, S.in_len + take + 16, S.in_len, s) != cudaSuccess) {
				S.reset();
				return ZG_ERR(ZG_error_memory_allocation);
			}
			ZG_CUDA(cudaMemcpyAsync(S.in.as<u8>() + S.in_len, src, take, cudaMemcpyHostToDevice, s));
			ZG_CUDA(cudaStreamSynchronize(s));  // the caller may reuse its input buffer as soon as we return
			// keep the last bytes for headers that straddle this call and the next
			uint8_t nc[24];
			uint32_t keep = (uint32_t)((S.ncarry + take) < 24 ? (S.ncarry + take) : 24);
			for (uint32_t i = 0; i < keep; i++) {
				uint64_t pos = S.in_len + take - keep + i;
				nc[i] = stream_byte(S, src, pos);
			}
			memcpy(S.carry, nc, keep);
			S.ncarry = keep;
			S.in_len += take;
			input->pos += take;
		}
		if (!S.frame_end || S.in_len < S.frame_end) {
			// frame incomplete: hint = what is known to be missing, at least one more block header
			if (S.frame_end) return (size_t)(S.frame_end - S.in_len);
			return S.have_header && S.next_hdr + 3 > S.in_len ? (size_t)(S.next_hdr + 3 - S.in_len) : 3;
		}
		// ---- frame complete: decode it on the device ----
		uint64_t bound = S.fcs_len ? S.fcs : S.blocks * 131072ull;
		if (S.fcs_len && S.fcs > S.blocks * 131072ull) {  // (a frame cannot regenerate more than 128 KiB per block)
			S.reset();
			return ZG_ERR(ZG_error_corruption_detected);
		}
		if (S.out.reserve(bound + 16) != cudaSuccess) {
			S.reset();
			return ZG_ERR(ZG_error_memory_allocation);
		}
		size_t r = decompress_dev(d, S.in.as<u8>(), S.frame_end, S.out.as<u8>(), bound);
		if (zg_is_error(r)) {
			S.reset();
			return r;
		}
		S.out_len = r;
		S.out_pos = 0;
		S.stage = 1;
	}
	size_t room = output->size - output->pos;
	size_t left = (size_t)(S.out_len - S.out_pos);
	size_t take = room < left ? room : left;
	if (take) ZG_CUDA(cudaMemcpy((uint8_t*)output->dst + output->pos, S.out.as<u8>() + S.out_pos, take, cudaMemcpyDeviceToHost));
	output->pos += take;
	S.out_pos += take;
	if (S.out_pos == S.out_len) {
		S.reset();
		return 0;
	}
	return (size_t)(S.out_len - S.out_pos);
}
size_t zg_decompress_stream(zg_dctx* d, zg_out_buffer* output, zg_in_buffer* input) {
	try {
		return decompress_stream_impl(d, output, input);
	} catch (...) {  // drop the half-collected frame, report it libzstd's way
		if (d) d->sm.reset();
		return ZG_ERR(ZG_error_memory_allocation);
	}
}

}  // extern "C"

// the compression context and zg_pack_batch live in abi_pack.cu

// ---------------------------------------------------------------------------------------------
// per-kernel device timing for bench.py's roofline: cudaEvent pairs on the launching stream
#include <utility>
static bool g_prof_on = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof[ZG_K_COUNT];
static std::mutex g_prof_mu;                             // (zg_unpack_batch decodes on two host threads)
static thread_local cudaEvent_t g_prof_open[ZG_K_COUNT];  // begin / end pairs belong to one thread
void zg_prof_begin(int k, cudaStream_t s) {
	if (!g_prof_on) return;
	cudaEvent_t a;
	cudaEventCreate(&a);
	cudaEventRecord(a, s);
	g_prof_open[k] = a;
}
void zg_prof_end(int k, cudaStream_t s) {
	if (!g_prof_on) return;
	cudaEvent_t b;
	cudaEventCreate(&b);
	cudaEventRecord(b, s);
	std::lock_guard<std::mutex> lk(g_prof_mu);
	g_prof[k].push_back({g_prof_open[k], b});
}
extern "C" {
void zg_profile_enable(int on) { g_prof_on = on != 0; }
// total device milliseconds and launch count of kernel class k since the last read
size_t zg_profile_read(int k, double* ms, uint64_t* launches) {
	if (k < 0 || k >= ZG_K_COUNT) return ZG_ERR(ZG_error_parameter_outOfBound);
	std::lock_guard<std::mutex> lk(g_prof_mu);
	double t = 0;
	for (auto& pr : g_prof[k]) {
		cudaEventSynchronize(pr.second);
		float f = 0;
		cudaEventElapsedTime(&f, pr.first, pr.second);
		t += f;
		cudaEventDestroy(pr.first);
		cudaEventDestroy(pr.second);
	}
	if (ms) *ms = t;
	if (launches) *launches = g_prof[k].size();
	g_prof[k].clear();
	return 0;
}
}
