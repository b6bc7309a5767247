"""Streaming-API fuzz on the SIMT-emulator build (CPU): intact and mutated frames (made by libzstd and by this library)
fed to zg_decompress_stream in random gulps with random output capacities, the way decode/zstd_iterator.rs:88-153 drives
DCtx::decompress_stream, ONE context reused from frame to frame (a frame that failed must leave it usable; only a frame abandoned for lack of input gets a new one).  Whatever the stream
delivers with a final hint of 0 must be exactly what the reference's streaming decoder restores from the same bytes;
nothing may crash, hang or deliver more than the reference does.
Usage: python tools/fuzz_stream.py FIRST_SEED SEEDS"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import fuzz_cases as fz  # noqa: E402
from zarc_b200 import _lib, build  # noqa: E402

t0 = time.time()


def log(seed, n):
    print(seed, "ok", n, round(time.time() - t0, 1), flush=True)


if __name__ == "__main__":
    lib = _lib.Lib(build.build_emu(), strict=False)
    print("done, all ok: (accepted and equal, rejected)", fz.streaming(lib, int(sys.argv[1]), int(sys.argv[2]), log=log))
