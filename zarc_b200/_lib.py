"""ctypes binding of include/zarcgpu.h.

`lib()` loads the product library zarc_b200/libzarcgpu.so (nvcc, sm_100a) and raises if it is
missing: there is no CPU fallback anywhere in this package.  (`Lib(path)` can be pointed at another
build of the same ABI; the CPU test-suite uses that to load the SIMT-emulator build of the kernel
sources, which is test infrastructure and lives under tests/.)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_SO = os.environ.get("ZARC_B200_LIB") or os.path.join(_HERE, "libzarcgpu.so")  # (the variable: another BUILD of the product library, for tuning runs)

_sz = C.c_size_t
_vp = C.c_void_p
_u64 = C.c_uint64

# every symbol include/zarcgpu.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "zg_is_error": (C.c_int, [_sz]),
    "zg_error_name": (C.c_char_p, [_sz]),
    "zg_get_error_code": (C.c_int, [_sz]),
    "zg_device_count": (C.c_int, []),
    "zg_set_device": (_sz, [C.c_int]),
    "zg_build_info": (C.c_char_p, []),
    "zg_blake3": (_sz, [_vp, _sz, _vp]),
    "zg_hasher_new": (_vp, []),
    "zg_hasher_update": (_sz, [_vp, _vp, _sz]),
    "zg_hasher_finalize": (_sz, [_vp, _vp]),
    "zg_hasher_free": (None, [_vp]),
    "zg_cctx_create": (_vp, []),
    "zg_cctx_free": (None, [_vp]),
    "zg_cctx_init": (_sz, [_vp, C.c_int]),
    "zg_cctx_set_parameter": (_sz, [_vp, C.c_int, C.c_int]),
    "zg_cctx_reset": (_sz, [_vp, C.c_int]),
    "zg_cctx_set_stream": (_sz, [_vp, _vp]),
    "zg_compress2": (_sz, [_vp, _vp, _sz, _vp, _sz]),
    "zg_compress_bound": (_sz, [_sz]),
    "zg_dctx_create": (_vp, []),
    "zg_dctx_free": (None, [_vp]),
    "zg_dctx_set_stream": (_sz, [_vp, _vp]),
    "zg_dctx_set_verify_checksum": (_sz, [_vp, C.c_int]),
    "zg_decompress_stream": (_sz, [_vp, _vp, _vp]),
    "zg_dstream_in_size": (_sz, []),
    "zg_dstream_out_size": (_sz, []),
    "zg_decompress": (_sz, [_vp, _vp, _sz, _vp, _sz]),
    "zg_find_frame_compressed_size": (_sz, [_vp, _sz]),
    "zg_cctx_reset_archive": (_sz, [_vp, _u64]),
    "zg_cctx_archive_offset": (_u64, [_vp]),
    "zg_pack_batch": (_sz, [_vp, _vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _u64, _vp]),
    "zg_pack_batch_dev": (_sz, [_vp, _vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _u64, _vp]),
    "zg_pack_batch_dev_ex": (_sz, [_vp, _vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _vp]),
    "zg_unpack_batch": (_sz, [_vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _u64, _vp, _vp, _vp]),
    "zg_unpack_batch_dev": (_sz, [_vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _u64, _vp, _vp, _vp]),
    "zg_blake3_batch_dev": (_sz, [_vp, _vp, _vp, _vp, _u64, _vp]),
    "zg_blake3_batch": (_sz, [_vp, _vp, _vp, _u64, _vp]),
    "zg_xxh64_batch_dev": (_sz, [_vp, _vp, _vp, _vp, _u64, _vp]),
    "zg_xxh64_batch": (_sz, [_vp, _vp, _vp, _u64, _vp]),
    "zg_assign_offsets_dev": (_sz, [_vp, _vp, _u64, _u64, _vp]),
    "zg_corpus_generate_dev": (_sz, [_vp, _vp, _vp, _vp, _vp, _vp, _u64]),
    "zg_corpus_generate_host": (_sz, [_vp, _vp, _vp, _vp, _vp, _u64]),
    "zg_kernel_launch_count": (_u64, []),
    "zg_dedup_dev": (_sz, [_vp, _vp, _u64, _vp, _vp]),
    "zg_profile_enable": (None, [C.c_int]),
    "zg_profile_read": (_sz, [C.c_int, C.POINTER(C.c_double), C.POINTER(_u64)]),
    "zg_alloc_pinned": (_vp, [_sz]),
    "zg_free_pinned": (None, [_vp]),
}

# libzstd's ZSTD_cParameter numbers (crates/zarc-cli/src/pack.rs:140-195 maps --zstd names to these)
ZG_c_compressionLevel = 100
ZG_c_windowLog = 101
ZG_c_contentSizeFlag = 200
ZG_c_checksumFlag = 201
ZG_reset_session_only = 1
ZG_reset_parameters = 2
ZG_reset_session_and_parameters = 3


class ZgError(RuntimeError):
    """A zstd-style error code crossed the boundary (crates/zarc/src/lib.rs:27-30 turns these into io::Error)."""

    def __init__(self, code: int, name: str):
        super().__init__(name)
        self.code = code
        self.name = name


class OutBuffer(C.Structure):
    _fields_ = [("dst", _vp), ("size", _sz), ("pos", _sz)]


class InBuffer(C.Structure):
    _fields_ = [("src", _vp), ("size", _sz), ("pos", _sz)]


def _ptr(a):
    """Address of a numpy array / bytes-like / int (device pointer) / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    if isinstance(a, (bytes, bytearray)):
        return C.cast(C.c_char_p(bytes(a)) if isinstance(a, bytes) else (C.c_char * len(a)).from_buffer(a), _vp).value
    raise TypeError(type(a))


class Lib:
    def __init__(self, path: str, strict: bool = True):
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build it with `python -m zarc_b200.build` "
                "(nvcc, sm_100a). zarc_b200 has no CPU fallback."
            )
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            try:
                f = getattr(self.dll, name)
            except AttributeError:
                if strict:
                    raise
                continue
            f.restype = res
            f.argtypes = args

    def check(self, code: int) -> int:
        if self.dll.zg_is_error(code):
            raise ZgError(self.dll.zg_get_error_code(code), self.dll.zg_error_name(code).decode())
        return code

    def __getattr__(self, name):
        return getattr(self.dll, name)


_product = None


def lib() -> Lib:
    """The product library. Raises if it has not been built."""
    global _product
    if _product is None:
        _product = Lib(PRODUCT_SO)
    return _product
