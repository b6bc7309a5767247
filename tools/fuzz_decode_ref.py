"""Reference-made frames through the decoder on the SIMT-emulator build (CPU): inputs glued from random / text / runs /
repeats / copies of earlier parts, up to a few MB (multi-block frames whose blocks read earlier blocks, Repeat_Mode
tables, Treeless literals, RLE blocks), compressed by libzstd 1.5.5 at levels 1 / 3 / 9 / 19 the way the reference does;
the decoder must restore every one byte-identically, digests verified, no frame decoded twice.
Usage: python tools/fuzz_decode_ref.py FIRST_SEED SEEDS"""
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
    print("done, all ok: frames", fz.decode_reference_frames(lib, int(sys.argv[1]), int(sys.argv[2]), log=log))
