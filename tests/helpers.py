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


def unpack_batch(lib, frames, ulens, digests=None, dctx=None, verify_checksum=True, junk_prefix=7):
    """Concatenate frames (with a junk prefix so offsets are non-trivial) and run zg_unpack_batch.
    Returns (outputs list[bytes], ok list[int] | None, status list[int], rc)."""
    archive = bytearray(b"\xAA" * junk_prefix)
    offs, lens = [], []
    for f in frames:
        offs.append(len(archive))
        lens.append(len(f))
        archive += f
    n = len(frames)
    arch = np.frombuffer(bytes(archive) or b"\0", dtype=np.uint8).copy()
    off = np.array(offs, dtype=np.uint64)
    ln = np.array(lens, dtype=np.uint64)
    ul = np.array(ulens, dtype=np.uint64)
    total = int(ul.sum())
    out = np.zeros(max(total, 1), dtype=np.uint8)
    ok = np.zeros(max(n, 1), dtype=np.uint8)
    status = np.zeros(max(n, 1), dtype=np.uint32)
    dig = None
    if digests is not None:
        dig = np.frombuffer(b"".join(digests) or b"\0" * 32, dtype=np.uint8).copy()
    own = dctx is None
    if own:
        dctx = lib.zg_dctx_create()
        assert dctx, "zg_dctx_create failed"
    try:
        lib.zg_dctx_set_verify_checksum(dctx, 1 if verify_checksum else 0)
        rc = lib.zg_unpack_batch(dctx, arch.ctypes.data, len(archive), n, off.ctypes.data, ln.ctypes.data, ul.ctypes.data,
                                 dig.ctypes.data if dig is not None else None, out.ctypes.data, total, None,
                                 ok.ctypes.data if dig is not None else None, status.ctypes.data)
    finally:
        if own:
            lib.zg_dctx_free(dctx)
    outs, pos = [], 0
    for u in ulens:
        outs.append(bytes(out[pos : pos + u]))
        pos += u
    return outs, (list(map(int, ok[:n])) if dig is not None else None), list(map(int, status[:n])), rc


def compress2(lib, data, level=3, checksum=True, cctx=None, content_size=True):
    """zg_compress2 with the reference's capacity rule (lowlevel_frames.rs:21)."""
    import ctypes as C

    own = cctx is None
    if own:
        cctx = lib.zg_cctx_create()
        assert cctx
        lib.check(lib.zg_cctx_init(cctx, 0))
        lib.check(lib.zg_cctx_set_parameter(cctx, 201, 1 if checksum else 0))
        lib.check(lib.zg_cctx_set_parameter(cctx, 100, level))
        if not content_size:
            lib.check(lib.zg_cctx_set_parameter(cctx, 200, 0))
    try:
        n = len(data)
        cap = n + max(1024, n // 10)
        dst = C.create_string_buffer(cap)
        r = lib.check(lib.zg_compress2(cctx, dst, cap, bytes(data), n))
        return dst.raw[:r]
    finally:
        if own:
            lib.zg_cctx_free(cctx)


def pack_batch(lib, cctx, files, align=1, cap=None):
    """zg_pack_batch over a list of byte strings. Returns dict of numpy outputs + the new frames bytes."""
    blob, off, lens = layout(files, align)
    n = len(files)
    total = int(lens.sum())
    if cap is None:
        cap = total + max(1024, total // 10) + 32 * n
    digests = np.zeros((max(n, 1), 32), dtype=np.uint8)
    first = np.zeros(max(n, 1), dtype=np.uint8)
    foff = np.zeros(max(n, 1), dtype=np.uint64)
    flen = np.zeros(max(n, 1), dtype=np.uint64)
    frames = np.zeros(max(cap, 1), dtype=np.uint8)
    nbytes = np.zeros(1, dtype=np.uint64)
    rc = lib.zg_pack_batch(cctx, blob.ctypes.data, off.ctypes.data, lens.ctypes.data, n, digests.ctypes.data, first.ctypes.data,
                           foff.ctypes.data, flen.ctypes.data, frames.ctypes.data, cap, nbytes.ctypes.data)
    return dict(rc=rc, digests=[bytes(d) for d in digests[:n]], first=list(map(int, first[:n])), off=list(map(int, foff[:n])),
                len=list(map(int, flen[:n])), frames=bytes(frames[: int(nbytes[0])]))
