"""Streaming-API fuzz on the SIMT-emulator build (CPU): intact and mutated frames (made by libzstd and by this library)
fed to zg_decompress_stream in random gulps with random output capacities, the way decode/zstd_iterator.rs:88-153 drives
DCtx::decompress_stream, ONE context reused from frame to frame (a frame that failed must leave it usable; only a frame abandoned for lack of input gets a new one).  Whatever the stream
delivers with a final hint of 0 must be exactly what the reference's streaming decoder restores from the same bytes;
nothing may crash, hang or deliver more than the reference does.
Usage: python tools/fuzz_stream.py FIRST_SEED SEEDS"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from zarc_b200 import build, _lib
from zarc_b200._lib import InBuffer, OutBuffer
from oracle import ref_path
from tests.helpers import compress2
from tests.golden.recipes import rand, text
from tests.test_decode_fuzz_emu import _mutations

lib = _lib.Lib(build.build_emu(), strict=False)


def stream(d, archive, rng, limit):
    """-> (bytes delivered, input consumed, error code or 0)"""
    pos, got = 0, b""
    calls = 0
    while True:
        n = int(rng.choice([1, 3, 17, 1000, 70_000, 131075]))
        gulp = archive[pos : pos + n]
        if not gulp:
            return got, pos, -1  # ran out of input before the frame ended
        ib = C.create_string_buffer(gulp, len(gulp))
        inb = InBuffer(C.cast(ib, C.c_void_p), len(gulp), 0)
        while True:
            cap = int(rng.choice([1, 100, 5000, 131072, 300_000]))
            ob = C.create_string_buffer(cap)
            outb = OutBuffer(C.cast(ob, C.c_void_p), cap, 0)
            r = lib.zg_decompress_stream(d, C.byref(outb), C.byref(inb))
            calls += 1
            assert calls < 200_000, "no progress"
            if lib.zg_is_error(r):
                return got, pos + inb.pos, lib.zg_get_error_code(r)
            got += ob.raw[: outb.pos]
            assert len(got) <= limit, "delivered more than any valid frame of this size could hold"
            if r == 0:
                return got, pos + inb.pos, 0
            if outb.pos < cap and inb.pos == inb.size:
                break
        pos += inb.pos


def main():
    first, count = int(sys.argv[1]), int(sys.argv[2])
    t0 = time.time()
    d = lib.zg_dctx_create()
    agree = rejected = 0
    for seed in range(first, first + count):
        rng = np.random.default_rng(seed)
        datas = [text(int(rng.integers(0, 9000)), seed), rand(int(rng.integers(0, 3000)), seed) + bytes(int(rng.integers(0, 5000))),
                 text(int(rng.integers(100_000, 300_000)), seed + 1)]
        for data in datas:
            base = ref_path.ref_compress(data, level=int(rng.choice([1, 3, 9])), checksum=bool(rng.integers(0, 2))) if rng.integers(0, 2) \
                else bytes(compress2(lib, data, level=int(rng.choice([1, 3])), checksum=bool(rng.integers(0, 2))))
            for fr in [base] + _mutations(base, rng, 4):
                archive = fr + b"\x28\xb5\x2f\xfdnext"
                try:
                    ref = ref_path.ref_decompress_stream(archive, 0)
                except ref_path.ZstdError:
                    ref = None
                got, used, err = stream(d, archive, rng, 128 * 1024 * (len(fr) // 3 + 2))
                if err == 0:
                    assert ref is not None and got == ref, (seed, "accepted what the reference rejects or restores differently")
                    agree += 1
                else:
                    rejected += 1
                    if fr is base:
                        raise AssertionError((seed, "intact frame refused", err))
                    if err == -1:  # abandoned in the middle of a frame: the reference would drop this DCtx (one per frame iterator)
                        lib.zg_dctx_free(d)
                        d = lib.zg_dctx_create()
        if seed % 10 == 0:
            print(seed, "agree", agree, "rejected", rejected, round(time.time() - t0, 1), flush=True)
    lib.zg_dctx_free(d)
    print("done: agree", agree, "rejected", rejected)


main()
