"""Reference-made frames through the decoder on the SIMT-emulator build (CPU): inputs glued from random / text / runs /
repeats / copies of earlier parts, up to a few MB (multi-block frames whose blocks read earlier blocks, Repeat_Mode
tables, Treeless literals, RLE blocks), compressed by libzstd 1.5.5 at levels 1 / 3 / 9 / 19 the way the reference does;
the decoder must restore every one byte-identically, digests verified, no frame decoded twice.
Usage: python tools/fuzz_decode_ref.py FIRST_SEED SEEDS"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from zarc_b200 import build, _lib
from oracle import ref_path
from tests.helpers import unpack_batch
from tests.golden.recipes import rand, text

lib = _lib.Lib(build.build_emu(), strict=False)
t0 = time.time()


def gen(rng, scale):
    parts = []
    for _ in range(int(rng.integers(1, 8))):
        k = int(rng.integers(0, 6))
        m = int(rng.choice([0, 1, 7, 100, 1000, 5000, 40000, 131072, 200000, 500000]) * scale) + int(rng.integers(0, 64))
        if k == 0:
            parts.append(rand(m, int(rng.integers(1, 1 << 30))))
        elif k == 1:
            parts.append(text(m, int(rng.integers(1, 1 << 30))))
        elif k == 2:
            parts.append(bytes([int(rng.integers(0, 256))]) * m)
        elif k == 3 and parts:
            parts.append(parts[int(rng.integers(0, len(parts)))][:m])
        elif k == 4:
            unit = rand(int(rng.integers(1, 40)), int(rng.integers(1, 1 << 30)))
            parts.append((unit * (m // max(1, len(unit)) + 1))[:m])
        else:
            a = bytearray(text(m, 77))
            for _ in range(m // 50):
                if m:
                    a[int(rng.integers(0, m))] = int(rng.integers(0, 256))
            parts.append(bytes(a))
    return b"".join(parts)


bad = 0
first, count = int(sys.argv[1]), int(sys.argv[2])
for seed in range(first, first + count):
    rng = np.random.default_rng(seed)
    datas = [gen(rng, float(rng.choice([0.1, 1.0, 1.0, 2.0]))) for _ in range(5)]
    for level in (1, 3, 9, 19):
        frames = [ref_path.ref_compress(d, level=level, checksum=bool(rng.integers(0, 2))) for d in datas]
        outs, ok, status, rc = unpack_batch(lib, frames, [len(d) for d in datas], [ref_path.c_blake3(d) for d in datas])
        if rc != 0 or outs != datas or not all(ok):
            bad += 1
            print("MISMATCH", seed, level, rc, status, [len(d) for d in datas], flush=True)
    if seed % 5 == 0:
        print(seed, "ok", bad, round(time.time() - t0, 1), flush=True)
print("done bad =", bad)
