"""zarc_b200: B200-native (sm_100a) implementation of zarc's content path -- BLAKE3 digests, dedup,
Zstandard frame encode on pack; Zstandard frame decode + BLAKE3 verify on unpack -- behind the
C ABI in include/zarcgpu.h.  No CPU fallback: the CUDA library must be built and a GPU present."""
from ._lib import Lib, ZgError, lib  # noqa: F401
