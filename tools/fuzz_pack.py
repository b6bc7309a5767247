"""Pack bookkeeping fuzz on the SIMT-emulator build (CPU): random file lists with duplicates (within a batch and across
batches), empty files and multi-block files, packed in a random number of zg_pack_batch calls with random host-API slice
and encoder chunk sizes.  Digests, first-occurrence flags, offsets and lengths must equal the reference Encoder's
bookkeeping (content_frame.rs:20-60 as restated in oracle/ref_path.py: offsets 12 + running sum in insertion order of
unique contents, duplicates answer with the first occurrence's frame), the archive bytes must not depend on how the work
was cut, and every frame must be restored by libzstd and by our decoder.
Usage: python tools/fuzz_pack.py FIRST_SEED SEEDS"""
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
    print("done, all ok: file lists", fz.pack_bookkeeping(lib, int(sys.argv[1]), int(sys.argv[2]), log=log))
