"""Directory fuzz of the C++ host (zarc_b200/zarc-b200) with the kernels on the SIMT emulator (CPU): archives whose
directory element stream has been mutated AFTER it was built -- bytes flipped, runs overwritten, truncated, elements
duplicated -- and then compressed, digested and given a matching trailer, so that the reader gets past the integrity
checks and has to parse what is in there.  `list-files` and `unpack` must end with exit code 0 or 1 (an error message),
never a signal, never a hang, and nothing may be written outside the extraction directory.
Usage: python tools/fuzz_container.py FIRST_SEED SEEDS"""
import os
import shutil
import struct
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import ref_container, ref_path
from tests.golden.recipes import rand, text
from zarc_b200 import build

EMU = build.build_emu()
HOST = build.build_host()


class Writer(ref_container.RefArchiveWriter):
    def directory_bytes(self) -> bytes:
        d = bytearray(ref_container.edition_element())
        for f in self.files:
            if f["digest"] is not None:
                d += ref_container.frame_element(self.enc.frames[f["digest"]])
            d += ref_container.file_element(**f)
        return bytes(d)

    def finalise_with(self, directory: bytes) -> bytes:
        out = bytearray(self.out)
        digest = ref_path._blake3(directory)
        comp = self.enc.compress_frame(directory)
        out += comp
        tb = ref_container.trailer_bytes(digest, -(len(comp) + 8 + 32 + ref_container.EPILOGUE_LENGTH), len(directory))
        out += bytes([0x5F, 0x2A, 0x4D, 0x18]) + struct.pack("<I", len(tb)) + tb
        return bytes(out)


def mutate(d: bytes, rng) -> bytes:
    f = bytearray(d)
    n = len(f)
    for _ in range(int(rng.integers(1, 4))):
        kind = int(rng.integers(0, 6))
        n = len(f)
        if n == 0:
            break
        if kind <= 1:
            p = int(rng.integers(0, n))
            f[p] ^= 1 << int(rng.integers(0, 8))
        elif kind == 2:
            p = int(rng.integers(0, n))
            m = int(rng.integers(1, 9))
            f[p : p + m] = bytes(rng.integers(0, 256, m, dtype=np.uint8))
        elif kind == 3:
            f = f[: int(rng.integers(0, n))]
        elif kind == 4:  # a big length / count in front of something
            p = int(rng.integers(0, n))
            f[p : p + 1] = bytes([int(rng.choice([0x5B, 0x9B, 0xBB, 0x7B, 0x1B]))]) + bytes(rng.integers(0, 256, 8, dtype=np.uint8))
        else:  # duplicate a run somewhere else
            p, q = int(rng.integers(0, n)), int(rng.integers(0, n))
            m = int(rng.integers(1, 40))
            f[q:q] = f[p : p + m]
    return bytes(f)


def main():
    first, count = int(sys.argv[1]), int(sys.argv[2])
    w = Writer(level=3)
    w.add_file(["a.txt"], text(3000, 1))
    w.add_file(["sub", "b.bin"], rand(2000, 2))
    w.add_file(["sub", "dup.txt"], text(3000, 1))
    w.add_file(["sub", "empty"], b"")
    w.add_file(["big.txt"], text(140_000, 4))
    base = w.directory_bytes()
    # the unmutated archive must unpack
    env = dict(os.environ, ZARCGPU_LIB=EMU)
    t0 = time.time()
    signals = hangs = escapes = accepted = 0
    for seed in range(first, first + count):
        rng = np.random.default_rng(seed)
        d = base if seed == first else mutate(base, rng)
        tmp = tempfile.mkdtemp(prefix="zfuzz")
        try:
            work = os.path.join(tmp, "w")
            os.makedirs(work)
            with open(os.path.join(work, "x.zarc"), "wb") as fh:
                fh.write(w.finalise_with(d))
            for cmd in (["list-files", "x.zarc"], ["unpack", "x.zarc"]):
                try:
                    p = subprocess.run([HOST, *cmd], cwd=work, env=env, capture_output=True, timeout=60)
                except subprocess.TimeoutExpired:
                    hangs += 1
                    print("HANG", seed, cmd, flush=True)
                    continue
                if p.returncode not in (0, 1):
                    signals += 1
                    print("SIGNAL/RC", seed, cmd, p.returncode, p.stderr[-200:], flush=True)
                elif p.returncode == 0 and cmd[0] == "unpack":
                    accepted += 1
            if seed == first:
                assert accepted == 1, "the unmutated archive must unpack"
            left = set(os.listdir(tmp)) - {"w"}
            if left:
                escapes += 1
                print("ESCAPE", seed, left, flush=True)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        if seed % 25 == 0:
            print(seed, "signals", signals, "hangs", hangs, "escapes", escapes, "accepted", accepted, round(time.time() - t0, 1), flush=True)
    print("done: signals", signals, "hangs", hangs, "escapes", escapes, "accepted", accepted)


main()
