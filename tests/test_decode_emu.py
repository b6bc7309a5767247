"""K5/K6 decoder logic on the SIMT emulator: reference-made (libzstd 1.5.5) frames must decode
byte-identically, with digests verified, and corrupt frames must be rejected like libzstd does."""
import base64
import json
import os

import numpy as np
import pytest

from oracle import ref_path
from tests.golden.recipes import RECIPES, make_input, text, rand
from tests.helpers import unpack_batch

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat.json")))


def _b3(d):
    import blake3

    return blake3.blake3(d).digest()


def test_golden_frames(emu):
    frames, datas = [], []
    for case in KAT["cases"]:
        data = make_input(case["recipe"])
        for level, b64 in case["frames"].items():
            frames.append(base64.b64decode(b64))
            datas.append(data)
    outs, ok, status, rc = unpack_batch(emu, frames, [len(d) for d in datas], [_b3(d) for d in datas])
    assert rc == 0, emu.zg_error_name(rc)
    assert status == [0] * len(frames)
    for o, d in zip(outs, datas):
        assert o == d
    assert ok == [1] * len(frames)


@pytest.mark.parametrize("level", [1, 3, 9, 19])
def test_levels_and_features(emu, level):
    rng = np.random.default_rng(5)
    datas = [text(200_000, 21), text(50_000, 22) + rand(140_000, 23) + text(90_000, 22),
             rng.integers(0, 4, 30_000, dtype=np.uint8).tobytes(), b"".join(b"RECORD_" + bytes([v]) for v in range(255)),
             make_input("binary_8000"), make_input("rle_then_text"), b"", b"z"]
    frames = [ref_path.ref_compress(d, level=level) for d in datas]
    outs, ok, status, rc = unpack_batch(emu, frames, [len(d) for d in datas], [_b3(d) for d in datas])
    assert rc == 0 and status == [0] * len(frames)
    assert outs == datas and ok == [1] * len(frames)


def periodic_cases():
    """Matches that overlap their own output (offset < length) for every short period, matches whose source was
    written by earlier sequences of the same 32-sequence row, and literal runs / matches around the per-lane /
    whole-warp copy threshold."""
    rng = np.random.default_rng(77)
    datas = []
    for period in list(range(1, 41)) + [63, 64, 65, 127, 200]:
        pat = rng.integers(0, 256, period, dtype=np.uint8).tobytes()
        reps = int(rng.integers(2, 40))
        tail = int(rng.integers(0, period + 1))
        datas.append(rand(int(rng.integers(0, 70)), period) + pat * reps + pat[:tail] + rand(int(rng.integers(1, 90)), period + 1))
    # chains: each short phrase is re-used right after it was produced (dependent matches inside one row)
    chain = bytearray(rand(24, 5))
    for k in range(600):
        n = int(rng.integers(4, 70))
        back = int(rng.integers(n, min(len(chain), 300) + 1)) if len(chain) >= n else len(chain)
        chain += chain[len(chain) - back : len(chain) - back + n]
        if k % 3 == 0:
            chain += rand(int(rng.integers(1, 80)), k)
    datas.append(bytes(chain))
    datas.append(b"".join(rand(int(rng.integers(40, 80)), k) + b"ab" * int(rng.integers(20, 45)) for k in range(200)))
    return datas


@pytest.mark.parametrize("level", [1, 3, 19])
def test_periodic_and_chained_matches(emu, level):
    datas = periodic_cases()
    frames = [ref_path.ref_compress(d, level=level) for d in datas]
    outs, ok, status, rc = unpack_batch(emu, frames, [len(d) for d in datas], [_b3(d) for d in datas])
    assert rc == 0 and status == [0] * len(frames)
    assert outs == datas and ok == [1] * len(frames)


def test_no_checksum_frames_and_digest_mismatch(emu):
    datas = [text(5000, 1), rand(300, 2)]
    frames = [ref_path.ref_compress(d, checksum=False) for d in datas]
    wrong = [_b3(datas[0]), _b3(b"something else")]
    outs, ok, status, rc = unpack_batch(emu, frames, [len(d) for d in datas], wrong)
    assert rc == 0 and outs == datas
    assert ok == [1, 0]  # a digest mismatch is not an error at this boundary (frame_iterator.rs:86-88)


def test_corruption_is_reported_per_frame(emu):
    data = text(5000, 3)
    good = ref_path.ref_compress(data)
    bad_ck = bytearray(good)
    bad_ck[-1] ^= 1
    bad_magic = b"\x00\x00\x00\x00" + good[4:]
    truncated = good[:-9]
    flipped = bytearray(good)
    flipped[len(good) // 2] ^= 0x55
    frames = [good, bytes(bad_ck), bad_magic, truncated, bytes(flipped), good]
    outs, ok, status, rc = unpack_batch(emu, frames, [len(data)] * 6, [_b3(data)] * 6)
    assert status[0] == 0 and status[5] == 0 and outs[0] == data and outs[5] == data
    assert status[1] == 22  # checksum_wrong
    assert status[2] == 10  # prefix_unknown
    assert status[3] != 0 and status[4] != 0
    assert ok == [1, 0, 0, 0, 0, 1]
    assert emu.zg_get_error_code(rc) == 22  # lowest failing frame
    # libzstd agrees on which frames are bad
    for f, st in zip(frames, status):
        try:
            ref_path.ref_decompress(f, len(data))
            assert st == 0
        except ref_path.ZstdError:
            assert st != 0


def test_one_shot_and_streaming_api(emu):
    import ctypes as C
    from zarc_b200._lib import InBuffer, OutBuffer

    data = text(300_000, 9)
    frame = ref_path.ref_compress(data)
    d = emu.zg_dctx_create()
    dst = C.create_string_buffer(len(data))
    r = emu.check(emu.zg_decompress(d, dst, len(data), frame, len(frame)))
    assert dst.raw[:r] == data
    assert emu.zg_get_error_code(emu.zg_decompress(d, dst, 100, frame, len(frame))) == 70
    # streaming with the reference's gulp sizes (zstd_iterator.rs:88-153), trailing bytes of a next frame present
    archive = frame + ref_path.ref_compress(b"next frame")
    in_size, out_size = emu.zg_dstream_in_size(), emu.zg_dstream_out_size()
    assert (in_size, out_size) == (131075, 131072)
    pos, got, done = 0, b"", False
    while not done:
        gulp = archive[pos : pos + in_size]
        ib = C.create_string_buffer(gulp, len(gulp))
        inb = InBuffer(C.cast(ib, C.c_void_p), len(gulp), 0)
        while True:
            ob = C.create_string_buffer(out_size)
            outb = OutBuffer(C.cast(ob, C.c_void_p), out_size, 0)
            hint = emu.check(emu.zg_decompress_stream(d, C.byref(outb), C.byref(inb)))
            got += ob.raw[: outb.pos]
            if hint == 0:
                done = True
                break
            if outb.pos < out_size and inb.pos == inb.size:
                break
        pos += inb.pos
    assert got == data and pos == len(frame)
    emu.zg_dctx_free(d)


def test_streaming_with_tiny_and_odd_gulps_and_the_streaming_hasher(emu):
    """The frame header and the 3-byte block headers may straddle any two calls; the streaming Hasher sees the same
    chunks the FrameIterator would feed it (decode/frame_iterator.rs:94-103)."""
    import ctypes as C
    from zarc_b200._lib import InBuffer, OutBuffer

    rng = np.random.default_rng(3)
    for data, level in ((text(300_000, 9) + rand(140_000, 4) + bytes(200_000), 3), (b"", 3), (b"x", 1), (text(70_000, 2), 9)):
        frame = ref_path.ref_compress(data, level=level)
        archive = frame + b"\x28\xb5\x2f\xfdjunk of a next frame"
        for gulps in ("bytes", "random"):
            d = emu.zg_dctx_create()
            h = emu.zg_hasher_new()
            pos, got, done = 0, b"", False
            while not done:
                n = 1 if (gulps == "bytes" and pos < 40) else int(rng.integers(1, 70_000))
                gulp = archive[pos : pos + n]
                ib = C.create_string_buffer(gulp, len(gulp))
                inb = InBuffer(C.cast(ib, C.c_void_p), len(gulp), 0)
                while True:
                    cap = int(rng.integers(1, 200_000))
                    ob = C.create_string_buffer(cap)
                    outb = OutBuffer(C.cast(ob, C.c_void_p), cap, 0)
                    hint = emu.check(emu.zg_decompress_stream(d, C.byref(outb), C.byref(inb)))
                    chunk = ob.raw[: outb.pos]
                    got += chunk
                    emu.check(emu.zg_hasher_update(h, chunk, len(chunk)))
                    if hint == 0:
                        done = True
                        break
                    if outb.pos < cap and inb.pos == inb.size:
                        break
                pos += inb.pos
            assert got == data and pos == len(frame)
            dig = C.create_string_buffer(32)
            emu.check(emu.zg_hasher_finalize(h, dig))
            assert dig.raw == _b3(data)
            emu.check(emu.zg_hasher_finalize(h, dig))  # does not consume
            assert dig.raw == _b3(data)
            emu.zg_hasher_free(h)
            emu.zg_dctx_free(d)
