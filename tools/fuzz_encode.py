"""Encoder round-trip fuzz on the SIMT-emulator build (CPU): inputs glued from random / text / runs / repeats / copies of
earlier parts (0-600 KB, so single- and multi-block frames), compressed at levels 1 / 3 / 6 / 9 with and without checksum;
every frame must be restored byte-identically by libzstd 1.5.5 (oracle/ref_path.py) and by this library's decoder, digests
verified.  Usage: python tools/fuzz_encode.py FIRST_SEED SEEDS   (400 seeds = 9 600 frames in ~16 min on one core)"""
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
    print("done, all ok: frames", fz.encode_roundtrips(lib, int(sys.argv[1]), int(sys.argv[2]), log=log))
