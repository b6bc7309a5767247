// C ABI of libzarcgpu (include/zarcgpu.h): contexts, host<->device staging, error names.
// Everything that touches content bytes is a CUDA kernel; the host code here only moves buffers
// and does bookkeeping on sizes/offsets.
#include "common.h"
#include <new>
#include <vector>
#include <string.h>

uint64_t g_zg_launches = 0;

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
uint64_t zg_kernel_launch_count(void) { return g_zg_launches; }

// ---------------------------------------------------------------------------------------------
// building blocks, device pointers
size_t zg_blake3_batch_dev(void* stream, const uint8_t* blob, const uint64_t* off, const uint64_t* len, uint64_t n, uint8_t* digests) {
	ZG_NEED_DEVICE();
	ZgB3Work w;
	size_t r = zg_blake3_run((cudaStream_t)stream, w, blob, off, len, n, digests);
	cudaStreamSynchronize((cudaStream_t)stream);
	zg_b3work_free(w);
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

}  // extern "C"
