"""Shared test helpers: call the C ABI with numpy host buffers."""
import ctypes as C

import numpy as np


def layout(files, align=1):
    """Concatenate byte strings into (blob u8[], off u64[], len u64[])."""
    lens = np.array([len(f) for f in files], dtype=np.uint64)
    padded = (lens + np.uint64(align - 1)) // np.uint64(align) * np.uint64(align)
    off = (np.cumsum(padded) - padded).astype(np.uint64)
    total = int(padded.sum())
    blob = np.zeros(max(total, 1), dtype=np.uint8)
    for f, o in zip(files, off):
        blob[int(o) : int(o) + len(f)] = np.frombuffer(f, dtype=np.uint8)
    return blob, off, lens


def blake3_batch(lib, files, align=1, shift=0):
    blob, off, lens = layout(files, align)
    if shift:  # exercise unaligned base addresses
        blob = np.concatenate([np.zeros(shift, np.uint8), blob])
        off = off + np.uint64(shift)
    out = np.zeros((len(files), 32), dtype=np.uint8)
    lib.check(lib.zg_blake3_batch(blob.ctypes.data, off.ctypes.data, lens.ctypes.data, len(files), out.ctypes.data))
    return [bytes(r) for r in out]


def xxh64_batch(lib, files, align=1, shift=0):
    blob, off, lens = layout(files, align)
    if shift:
        blob = np.concatenate([np.zeros(shift, np.uint8), blob])
        off = off + np.uint64(shift)
    out = np.zeros(len(files), dtype=np.uint64)
    lib.check(lib.zg_xxh64_batch(blob.ctypes.data, off.ctypes.data, lens.ctypes.data, len(files), out.ctypes.data))
    return [int(v) for v in out]
