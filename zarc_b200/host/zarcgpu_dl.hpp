// Run-time binding of libzarcgpu.so (include/zarcgpu.h) for the C++ host.
//
// The host side never links a codec: every content byte goes through the C ABI of the CUDA library.
// The library is found through $ZARCGPU_LIB, else next to the executable, else by the loader's search
// path.
#pragma once
#include <dlfcn.h>
#include <stdexcept>
#include <string>
#include <unistd.h>
#include <limits.h>

#include "../../include/zarcgpu.h"

namespace zarc {

struct GpuLib {
	void* handle = nullptr;
#define ZG_FN(name) decltype(&::name) name = nullptr
	ZG_FN(zg_is_error);
	ZG_FN(zg_error_name);
	ZG_FN(zg_get_error_code);
	ZG_FN(zg_device_count);
	ZG_FN(zg_set_device);
	ZG_FN(zg_alloc_pinned);
	ZG_FN(zg_free_pinned);
	ZG_FN(zg_blake3);
	ZG_FN(zg_hasher_new);
	ZG_FN(zg_hasher_update);
	ZG_FN(zg_hasher_finalize);
	ZG_FN(zg_hasher_free);
	ZG_FN(zg_cctx_create);
	ZG_FN(zg_cctx_free);
	ZG_FN(zg_cctx_init);
	ZG_FN(zg_cctx_set_parameter);
	ZG_FN(zg_cctx_reset);
	ZG_FN(zg_compress2);
	ZG_FN(zg_compress_bound);
	ZG_FN(zg_dctx_create);
	ZG_FN(zg_dctx_free);
	ZG_FN(zg_decompress);
	ZG_FN(zg_find_frame_compressed_size);
	ZG_FN(zg_cctx_reset_archive);
	ZG_FN(zg_cctx_archive_offset);
	ZG_FN(zg_pack_batch);
	ZG_FN(zg_unpack_batch);
#undef ZG_FN

	static std::string exe_dir() {
		char buf[PATH_MAX];
		ssize_t n = readlink("/proc/self/exe", buf, sizeof buf - 1);
		if (n <= 0) return ".";
		buf[n] = 0;
		std::string s(buf);
		size_t k = s.rfind('/');
		return k == std::string::npos ? "." : s.substr(0, k);
	}

	GpuLib() {
		const char* env = getenv("ZARCGPU_LIB");
		std::string tried;
		for (std::string p : {std::string(env ? env : ""), exe_dir() + "/libzarcgpu.so", std::string("libzarcgpu.so")}) {
			if (p.empty()) continue;
			handle = dlopen(p.c_str(), RTLD_NOW | RTLD_LOCAL);
			if (handle) break;
			tried += "\n  " + p + ": " + dlerror();
		}
		if (!handle) throw std::runtime_error("cannot load libzarcgpu.so (there is no CPU code path for content bytes):" + tried);
#define ZG_FN(name)                                                                  \
	name = reinterpret_cast<decltype(name)>(dlsym(handle, #name));                   \
	if (!name) throw std::runtime_error(std::string("libzarcgpu.so lacks ") + #name)
		ZG_FN(zg_is_error);
		ZG_FN(zg_error_name);
		ZG_FN(zg_get_error_code);
		ZG_FN(zg_device_count);
		ZG_FN(zg_set_device);
		ZG_FN(zg_alloc_pinned);
		ZG_FN(zg_free_pinned);
		ZG_FN(zg_blake3);
		ZG_FN(zg_hasher_new);
		ZG_FN(zg_hasher_update);
		ZG_FN(zg_hasher_finalize);
		ZG_FN(zg_hasher_free);
		ZG_FN(zg_cctx_create);
		ZG_FN(zg_cctx_free);
		ZG_FN(zg_cctx_init);
		ZG_FN(zg_cctx_set_parameter);
		ZG_FN(zg_cctx_reset);
		ZG_FN(zg_compress2);
		ZG_FN(zg_compress_bound);
		ZG_FN(zg_dctx_create);
		ZG_FN(zg_dctx_free);
		ZG_FN(zg_decompress);
		ZG_FN(zg_find_frame_compressed_size);
		ZG_FN(zg_cctx_reset_archive);
		ZG_FN(zg_cctx_archive_offset);
		ZG_FN(zg_pack_batch);
		ZG_FN(zg_unpack_batch);
#undef ZG_FN
	}
	GpuLib(const GpuLib&) = delete;
	GpuLib& operator=(const GpuLib&) = delete;

	// zstd-style error code -> exception carrying libzstd's error string (crates/zarc/src/lib.rs:27-30
	// maps these to io::Error::other(name); decode/error.rs:35-38 to Error::Zstd(name))
	size_t check(size_t code, const char* what) const {
		if (zg_is_error(code)) throw std::runtime_error(std::string(what) + ": " + zg_error_name(code));
		return code;
	}
};

inline GpuLib& gpu() {
	static GpuLib lib;
	return lib;
}

}  // namespace zarc
