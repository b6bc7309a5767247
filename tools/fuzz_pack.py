"""Pack bookkeeping fuzz on the SIMT-emulator build (CPU): random file lists with duplicates (within a batch and across
batches), empty files and multi-block files, packed in a random number of zg_pack_batch calls with random host-API slice
and encoder chunk sizes.  Digests, first-occurrence flags, offsets and lengths must equal the reference Encoder's
bookkeeping (content_frame.rs:20-60 as restated in oracle/ref_path.py: offsets 12 + running sum in insertion order of
unique contents, duplicates answer with the first occurrence's frame), the archive bytes must not depend on how the work
was cut, and every frame must be restored by libzstd and by our decoder.
Usage: python tools/fuzz_pack.py FIRST_SEED SEEDS"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from zarc_b200 import build, _lib
from oracle import ref_path
from tests.helpers import pack_batch, unpack_batch
from tests.golden.recipes import rand, text

lib = _lib.Lib(build.build_emu(), strict=False)
t0 = time.time()


def one_file(rng):
    k = int(rng.integers(0, 5))
    m = int(rng.choice([0, 0, 1, 50, 700, 3000, 20000, 131072, 140000, 300000])) + int(rng.integers(0, 16))
    if m <= 16 and rng.integers(0, 2):
        m = 0
    s = int(rng.integers(1, 1 << 30))
    if k == 0:
        return rand(m, s)
    if k == 1:
        return bytes([s & 255]) * m
    return text(m, s)


def run(seed):
    rng = np.random.default_rng(seed)
    pool = [one_file(rng) for _ in range(int(rng.integers(1, 12)))]
    files = [pool[int(rng.integers(0, len(pool)))] for _ in range(int(rng.integers(1, 30)))]
    level = int(rng.choice([1, 3, 9]))
    out = bytearray()
    enc = ref_path.RefEncoder(out, level=level)
    ref_digests = [enc.add_data_frame(f) for f in files]
    seen, first = {}, []
    for i, d in enumerate(ref_digests):
        first.append(0 if d in seen else 1)
        seen.setdefault(d, i)
    archives = []
    for trial in range(2):
        slice_bytes = int(rng.choice([0, 10_000, 70_000, 400_000]))
        chunk_bytes = int(rng.choice([0, 4096, 65536, 262144]))
        lib.dll.zg_internal_set_slice_bytes(C.c_uint64(slice_bytes))
        lib.dll.zg_internal_set_encode_chunk_bytes(C.c_uint64(chunk_bytes))
        cuts = sorted(set(int(x) for x in rng.integers(0, len(files) + 1, int(rng.integers(0, 3))))) + [len(files)]
        c = lib.zg_cctx_create()
        lib.check(lib.zg_cctx_init(c, 0))
        lib.check(lib.zg_cctx_set_parameter(c, 201, 1))
        lib.check(lib.zg_cctx_set_parameter(c, 100, level))
        lib.check(lib.zg_cctx_reset_archive(c, 12))
        got = dict(digests=[], first=[], off=[], len=[], frames=b"")
        a = 0
        for b in cuts:
            r = pack_batch(lib, c, files[a:b])
            assert r["rc"] == 0, (seed, r["rc"])
            for k in ("digests", "first", "off", "len"):
                got[k] += r[k]
            got["frames"] += r["frames"]
            a = b
        assert lib.zg_cctx_archive_offset(c) == 12 + len(got["frames"])
        lib.zg_cctx_free(c)
        assert got["digests"] == ref_digests, seed
        assert got["first"] == first, seed
        pos = 12
        for i, f in enumerate(files):
            if first[i]:
                assert got["off"][i] == pos, (seed, i)
                pos += got["len"][i]
            else:
                j = seen[ref_digests[i]]
                assert (got["off"][i], got["len"][i]) == (got["off"][j], got["len"][j]), (seed, i)
        assert pos - 12 == len(got["frames"])
        archives.append(got["frames"])
    assert archives[0] == archives[1], (seed, "archive bytes depend on slices / chunks / batch cuts")
    fr = [got["frames"][o - 12 : o - 12 + l] for o, l, f1 in zip(got["off"], got["len"], first) if f1]
    uniq = [f for f, f1 in zip(files, first) if f1]
    for f, x in zip(uniq, fr):
        assert ref_path.ref_decompress(x, len(f)) == f, seed
    outs, ok, status, rc = unpack_batch(lib, fr, [len(f) for f in uniq], [d for d, f1 in zip(ref_digests, first) if f1])
    assert rc == 0 and outs == uniq and all(ok), seed


first_seed, count = int(sys.argv[1]), int(sys.argv[2])
try:
    for seed in range(first_seed, first_seed + count):
        run(seed)
        if seed % 10 == 0:
            print(seed, "ok", round(time.time() - t0, 1), flush=True)
    print("done, all ok")
finally:
    lib.dll.zg_internal_set_slice_bytes(C.c_uint64(0))
    lib.dll.zg_internal_set_encode_chunk_bytes(C.c_uint64(0))
