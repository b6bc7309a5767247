"""Longer runs of the decoder fuzz of tests/test_decode_fuzz_emu.py on the SIMT-emulator build (CPU): mutated frames must
never crash or hang the decoder, and whatever it accepts must be what the reference's streaming decoder restores.
Usage: python tools/fuzz_decode.py SEEDS PER_FRAME   (150 x 30 = 74 400 frames in ~13 min on one core)"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zarc_b200 import build, _lib
from tests.test_decode_fuzz_emu import run_fuzz
lib = _lib.Lib(build.build_emu(), strict=False)
t0 = time.time()
for seed in range(1000, 1000 + int(sys.argv[1])):
    r = run_fuzz(lib, seed, int(sys.argv[2]))
    print(seed, r, round(time.time() - t0, 1), flush=True)
