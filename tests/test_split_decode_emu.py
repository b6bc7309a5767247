"""The staged pipeline for multi-block frames (zstd_decode_staged.cu): every block of every frame is entropy-decoded in
parallel (tables inherited from earlier blocks, repeat offsets symbolic), then executed -- at once when the block reads
nothing before itself (this library's encoder), after the blocks it reads otherwise (libzstd).  No frame is decoded twice,
and the bytes equal the reference decoder's."""
import ctypes as C
import struct

import numpy as np

from oracle import ref_path
from tests.golden.recipes import rand, text
from tests.helpers import pack_batch, unpack_batch

BLOCK = 128 * 1024


def _b3(d):
    import blake3

    return blake3.blake3(d).digest()


def _stats(lib):
    out = (C.c_uint64 * 3)()
    lib.dll.zg_internal_decode_stats(out)
    return list(out)


def _stats_ex(lib):
    """{frames, work items, frames decoded twice, staged frames, staged blocks, blocks that waited for earlier output, chunks}"""
    out = (C.c_uint64 * 8)()
    lib.dll.zg_internal_decode_stats_ex(out)
    return list(out)


def _own_frames(lib, files, level=3):
    cctx = lib.zg_cctx_create()
    lib.check(lib.zg_cctx_init(cctx, level))
    lib.check(lib.zg_cctx_set_parameter(cctx, 201, 1))
    lib.check(lib.zg_cctx_reset_archive(cctx, 12))
    r = pack_batch(lib, cctx, files)
    lib.zg_cctx_free(cctx)
    assert r["rc"] == 0
    return [bytes(r["frames"][o - 12 : o - 12 + l]) for o, l in zip(r["off"], r["len"])]


def _files():
    return [text(300_000, 31), text(BLOCK + 1, 32), text(2 * BLOCK, 33), text(5000, 34), rand(3 * BLOCK + 77, 35),
            text(200_000, 36) + bytes(150_000) + text(90_000, 37), b"", text(BLOCK, 38), text(1_000_000, 39)]


def nblocks(n):
    return max(1, -(-n // BLOCK))


def test_own_multi_block_frames_decode_block_parallel(emu):
    files = _files()
    frames = _own_frames(emu, files)
    outs, ok, status, rc = unpack_batch(emu, frames, [len(f) for f in files], [_b3(f) for f in files])
    assert rc == 0 and status == [0] * len(files)
    assert outs == files and ok == [1] * len(files)
    n_frames, n_items, n_redo = _stats(emu)
    assert n_frames == len(files)
    assert n_items == sum(nblocks(len(f)) for f in files)  # every block of every multi-block frame was its own item
    assert n_redo == 0                                       # ... and none of them needed its neighbours
    for f, fr in zip(files, frames):                         # the reference decoder agrees
        assert ref_path.ref_decompress(fr, len(f)) == f


def test_reference_multi_block_frames_decode_staged_in_one_pass(emu):
    files = _files()
    for level in (1, 3, 9, 19):
        frames = [ref_path.ref_compress(f, level=level) for f in files]
        outs, ok, status, rc = unpack_batch(emu, frames, [len(f) for f in files], [_b3(f) for f in files])
        assert rc == 0 and status == [0] * len(files)
        assert outs == files and ok == [1] * len(files)
        n_frames, n_items, n_redo = _stats(emu)
        ex = _stats_ex(emu)
        assert n_redo == 0                 # nothing is decoded twice
        assert n_items > n_frames and ex[3] >= 5   # the multi-block frames went through the staged pipeline ...
        assert ex[5] >= 3                  # ... and libzstd's text frames have blocks that read earlier blocks


def test_chunked_staging_equals_one_pass(emu):
    """Tiny chunks (a few blocks each): the per-frame state (output offset, repeat history, completion) carries over."""
    files = _files()
    try:
        for chunk in (1, 3, 7):
            emu.dll.zg_internal_set_decode_chunk_blocks(C.c_uint32(chunk))
            for make in (lambda: _own_frames(emu, files), lambda: [ref_path.ref_compress(f, level=3) for f in files]):
                frames = make()
                outs, ok, status, rc = unpack_batch(emu, frames, [len(f) for f in files], [_b3(f) for f in files])
                assert rc == 0 and status == [0] * len(files) and outs == files and ok == [1] * len(files)
            assert _stats_ex(emu)[6] >= 3
    finally:
        emu.dll.zg_internal_set_decode_chunk_blocks(C.c_uint32(0))


def _window_edge_files():
    """What stresses the chain executor's shared-memory window: matches longer than the window (periods of 70 000 and
    200 000 bytes), matches that overlap themselves across a slide (period 37), sources far behind the window, stored
    blocks in the middle of a chain, and a long tail of literals."""
    rng = np.random.default_rng(99)
    a = rand(70_000, 1)
    b = rand(200_000, 2)
    return [a * 5 + text(30_000, 3),
            b + b + text(150_000, 4) + b[:150_000],
            text(20_000, 5) + rand(37, 6) * 9000 + text(100_000, 5),
            text(300_000, 7) + rand(140_000, 8) + text(300_000, 7),
            text(180_000, 9) + bytes(rng.integers(0, 256, 131_000, dtype=np.uint8)),
            (text(5000, 10) + bytes(3000)) * 60]


def test_chain_executor_window_edges_and_both_modes(emu):
    files = _window_edge_files()
    try:
        for mode in (2, 0):  # 2: always the chain executor, 0: per-block flags
            emu.dll.zg_internal_set_decode_chain_mode(C.c_uint32(mode))
            for level in (1, 3):
                frames = [ref_path.ref_compress(f, level=level) for f in files]
                outs, ok, status, rc = unpack_batch(emu, frames, [len(f) for f in files], [_b3(f) for f in files])
                assert rc == 0 and status == [0] * len(files)
                assert outs == files and ok == [1] * len(files)
                ex = _stats_ex(emu)
                assert ex[5] > 0 and (ex[7] > 0) == (mode == 2)
        # small chunks: the window starts over at every chunk boundary
        emu.dll.zg_internal_set_decode_chain_mode(C.c_uint32(2))
        emu.dll.zg_internal_set_decode_chunk_blocks(C.c_uint32(5))
        frames = [ref_path.ref_compress(f, level=3) for f in files]
        outs, ok, status, rc = unpack_batch(emu, frames, [len(f) for f in files], [_b3(f) for f in files])
        assert rc == 0 and outs == files and ok == [1] * len(files)
    finally:
        emu.dll.zg_internal_set_decode_chain_mode(C.c_uint32(1))
        emu.dll.zg_internal_set_decode_chunk_blocks(C.c_uint32(0))


def test_split_threshold_and_mixed_batch(emu):
    files = _files()
    own = _own_frames(emu, files)
    ref = [ref_path.ref_compress(f) for f in files]
    mixed = [own[i] if i % 2 else ref[i] for i in range(len(files))]
    try:
        for split_min in (0, 4 * BLOCK, 1 << 40):
            emu.dll.zg_internal_set_decode_split_min(C.c_uint64(split_min))
            outs, ok, status, rc = unpack_batch(emu, mixed, [len(f) for f in files], [_b3(f) for f in files])
            assert rc == 0 and outs == files and ok == [1] * len(files)
            if split_min == 1 << 40:
                assert _stats(emu)[1] == len(files)  # nothing split
    finally:
        emu.dll.zg_internal_set_decode_split_min(C.c_uint64(0))


def _raw_frame(parts, checksum=True):
    """A frame of Raw blocks of the given sizes (single segment, 4-byte content size)."""
    import xxhash

    data = b"".join(parts)
    out = bytearray(struct.pack("<IB", 0xFD2FB528, 0x80 | 0x20 | (4 if checksum else 0)) + struct.pack("<I", len(data)))
    for i, p in enumerate(parts):
        h = (len(p) << 3) | (1 if i + 1 == len(parts) else 0)
        out += struct.pack("<I", h)[:3] + p
    if checksum:
        out += struct.pack("<I", xxhash.xxh64(data, seed=0).intdigest() & 0xFFFFFFFF)
    return bytes(out), data


def test_short_blocks_are_not_mistaken_for_full_ones(emu):
    # blocks of other sizes than 128 KiB: every block's place comes from the prefix sum of the sizes before it
    f1, d1 = _raw_frame([rand(100_000, 1), rand(100_000, 2)])
    f2, d2 = _raw_frame([rand(1000, 3), rand(BLOCK, 4), rand(70_000, 5)])
    # full blocks
    f3, d3 = _raw_frame([rand(BLOCK, 6), rand(BLOCK, 7), rand(5, 8)])
    for f, d in ((f1, d1), (f2, d2), (f3, d3)):
        assert ref_path.ref_decompress(f, len(d)) == d
    outs, ok, status, rc = unpack_batch(emu, [f1, f2, f3], [len(d1), len(d2), len(d3)], [_b3(d1), _b3(d2), _b3(d3)])
    assert rc == 0 and status == [0, 0, 0] and outs == [d1, d2, d3] and ok == [1, 1, 1]
    n_frames, n_items, n_redo = _stats(emu)
    assert (n_frames, n_items, n_redo) == (3, 2 + 3 + 3, 0)  # blocks of any size are staged; nothing is decoded twice


def test_checksums_hashed_chunk_by_chunk(emu):
    """With several chunks per launch the Content_Checksum of a staged frame is absorbed piece by piece behind the chunks
    (k_zds_xxh64_partial) and finished by the check: good frames pass, a wrong checksum field and a flipped content byte
    that still decodes are reported as checksum_wrong, whatever the chunk size."""
    files = [text(400_000, 51), text(300_000, 52) + rand(140_000, 53), text(9_000, 54), text(262_144, 55)]
    frames = _own_frames(emu, files)
    bad_ck = bytearray(frames[0])
    bad_ck[-1] ^= 0x80
    ref = [ref_path.ref_compress(f, level=1) for f in files]
    # a stored (Raw) block's payload can be altered without breaking the frame: find one in the random part
    batch = [frames[0], bytes(bad_ck), frames[1], frames[2], frames[3]] + ref
    sizes = [len(files[0]), len(files[0]), len(files[1]), len(files[2]), len(files[3])] + [len(f) for f in files]
    digs = [_b3(files[0]), _b3(files[0]), _b3(files[1]), _b3(files[2]), _b3(files[3])] + [_b3(f) for f in files]
    try:
        for chunk in (0, 1, 2, 5):
            emu.dll.zg_internal_set_decode_chunk_blocks(C.c_uint32(chunk))
            outs, ok, status, rc = unpack_batch(emu, batch, sizes, digs)
            assert status == [0, 22, 0, 0, 0, 0, 0, 0, 0], (chunk, status)
            assert ok == [1, 0, 1, 1, 1, 1, 1, 1, 1]
            assert outs[0] == files[0] and outs[2] == files[1] and outs[5:] == files
    finally:
        emu.dll.zg_internal_set_decode_chunk_blocks(C.c_uint32(0))


def test_corruption_inside_split_frames(emu):
    files = [text(400_000, 41), text(300_000, 42), text(290_000, 43), text(10_000, 44)]
    frames = _own_frames(emu, files)
    bad_mid = bytearray(frames[0])
    bad_mid[len(bad_mid) // 2] ^= 0x5A
    bad_ck = bytearray(frames[1])
    bad_ck[-2] ^= 1
    truncated = frames[2][:-20]
    batch = [bytes(bad_mid), bytes(bad_ck), truncated, frames[3], frames[0]]
    sizes = [len(files[0]), len(files[1]), len(files[2]), len(files[3]), len(files[0])]
    digs = [_b3(files[0]), _b3(files[1]), _b3(files[2]), _b3(files[3]), _b3(files[0])]
    outs, ok, status, rc = unpack_batch(emu, batch, sizes, digs)
    assert status[3] == 0 and status[4] == 0 and outs[3] == files[3] and outs[4] == files[0]
    assert status[0] != 0 and status[2] != 0
    assert status[1] == 22  # checksum_wrong
    assert ok == [0, 0, 0, 1, 1]
    for f, n, st in zip(batch, sizes, status):  # libzstd agrees on which frames are bad
        try:
            ref_path.ref_decompress(f, n)
            assert st == 0
        except ref_path.ZstdError:
            assert st != 0
