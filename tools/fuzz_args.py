"""Hostile metadata fuzz on the SIMT-emulator build (CPU): zg_unpack_batch_dev / zg_unpack_batch over batches in which some
entries carry wild offsets, lengths, sizes or output offsets; no crash, every untouched entry still decodes, every wild
one is reported in its own status entry (tests/fuzz_cases.py: hostile_metadata); then zg_compress2 into exact-size
destinations of every awkward capacity (compress_capacity).  Meant to be run with the emulator built
with AddressSanitizer as well (see profiles/README.md).
Usage: python tools/fuzz_args.py FIRST_SEED SEEDS"""
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
    print("done, all ok: batches", fz.hostile_metadata(lib, int(sys.argv[1]), int(sys.argv[2]), log=log))
    print("done, all ok: zg_compress2 calls (frames made, refused for capacity)", fz.compress_capacity(lib, int(sys.argv[1]), max(1, int(sys.argv[2]) // 10), log=log))
