"""Directory fuzz of the C++ host (zarc_b200/zarc-b200) with the kernels on the SIMT emulator (CPU): archives whose
directory element stream has been mutated AFTER it was built -- bytes flipped, runs overwritten, truncated, elements
duplicated -- and then compressed, digested and given a matching trailer, so that the reader gets past the integrity
checks and has to parse what is in there.  `list-files` and `unpack` must end with exit code 0 or 1 (an error message),
never a signal, never a hang, and nothing may be written outside the extraction directory.
Usage: python tools/fuzz_container.py FIRST_SEED SEEDS"""
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
    print("done, no signal / hang / escape: archives that unpacked", fz.container_directories(build.build_emu(), int(sys.argv[1]), int(sys.argv[2]), log=log))
