"""CPU-side ratio lab: run the encoder kernels on the SIMT emulator over a C2-shaped sample and report
the compression ratio next to libzstd's (test infrastructure; never part of the product path)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zarc_b200 import _lib, build
from tests.helpers import pack_batch
from zarc_b200 import corpus
from oracle import ref_path

mb = float(sys.argv[1]) if len(sys.argv) > 1 else 4
level = int(sys.argv[2]) if len(sys.argv) > 2 else 3
emu = _lib.Lib(build.build_emu(), strict=False)
c = corpus.c2_source_tree(total_bytes=int(mb * 1e6), seed=3)
blob = corpus.materialise_host(emu, c)
files = [bytes(blob[int(o):int(o) + int(l)]) for o, l in zip(c.off, c.len)]
cctx = emu.zg_cctx_create()
emu.check(emu.zg_cctx_init(cctx, level))
emu.check(emu.zg_cctx_set_parameter(cctx, 201, 1))
emu.check(emu.zg_cctx_reset_archive(cctx, 12))
t0 = time.time()
r = pack_batch(emu, cctx, files)
t1 = time.time()
assert r["rc"] == 0
ours = sum(r["len"])
ref = sum(len(ref_path.ref_compress(f, level=level)) for f in files)
bad = 0
if "--check" in sys.argv:
    for f, o, l in zip(files, r["off"], r["len"]):
        if ref_path.ref_decompress(r["frames"][o - 12:o - 12 + l], len(f)) != f:
            bad += 1
n = sum(len(f) for f in files)
print(f"files {len(files)} bytes {n} ours {ours} ratio {n/ours:.4f} ref {n/ref:.4f} ours/ref {ours/ref:.4f} emu {t1-t0:.1f}s bad {bad}")
