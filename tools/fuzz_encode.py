"""Encoder round-trip fuzz on the SIMT-emulator build (CPU): inputs glued from random / text / runs / repeats / copies of
earlier parts (0-600 KB, so single- and multi-block frames), compressed at levels 1 / 3 / 6 / 9 with and without checksum;
every frame must be restored byte-identically by libzstd 1.5.5 (oracle/ref_path.py) and by this library's decoder, digests
verified.  Usage: python tools/fuzz_encode.py FIRST_SEED SEEDS   (400 seeds = 9 600 frames in ~16 min on one core)"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zarc_b200 import build, _lib
from oracle import ref_path
from tests.helpers import compress2, unpack_batch
from tests.golden.recipes import rand, text
lib = _lib.Lib(build.build_emu(), strict=False)
t0 = time.time()
def gen(rng):
    parts = []
    n = int(rng.integers(0, 6))
    for _ in range(n):
        k = int(rng.integers(0, 6))
        m = int(rng.choice([0, 1, 3, 7, 31, 100, 1000, 5000, 40000, 131072, 200000])) + int(rng.integers(0, 64))
        if k == 0: parts.append(rand(m, int(rng.integers(1, 1 << 30))))
        elif k == 1: parts.append(text(m, int(rng.integers(1, 1 << 30))))
        elif k == 2: parts.append(bytes([int(rng.integers(0, 256))]) * m)
        elif k == 3 and parts: parts.append(parts[int(rng.integers(0, len(parts)))][:m])
        elif k == 4:
            unit = rand(int(rng.integers(1, 40)), int(rng.integers(1, 1 << 30)))
            parts.append((unit * (m // max(1, len(unit)) + 1))[:m])
        else:
            a = bytearray(text(m, 77))
            for _ in range(m // 50):
                if m: a[int(rng.integers(0, m))] = int(rng.integers(0, 256))
            parts.append(bytes(a))
    return b"".join(parts)
bad = 0
for seed in range(int(sys.argv[1]), int(sys.argv[1]) + int(sys.argv[2])):
    rng = np.random.default_rng(seed)
    datas = [gen(rng) for _ in range(6)]
    for level in (1, 3, 6, 9):
        frames = []
        for d in datas:
            fr = bytes(compress2(lib, d, level=level, checksum=bool(rng.integers(0, 2))))
            if ref_path.ref_decompress(fr, len(d)) != d:
                bad += 1
                print("MISMATCH ref", seed, level, len(d), flush=True)
            frames.append(fr)
        outs, ok, status, rc = unpack_batch(lib, frames, [len(d) for d in datas], [ref_path.c_blake3(d) for d in datas])
        if rc != 0 or outs != datas or not all(ok):
            bad += 1
            print("MISMATCH own", seed, level, rc, status, flush=True)
    if seed % 5 == 0: print(seed, "ok", bad, round(time.time() - t0, 1), flush=True)
print("done bad =", bad)
