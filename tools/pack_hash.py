"""Digest of the archive bytes the encoder kernels produce (on the SIMT emulator) for a fixed C2-shaped sample and a few
levels: an optimisation that is meant to leave the output bit-identical is checked by comparing this before and after."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import pack_batch  # noqa: E402
from zarc_b200 import _lib, build, corpus  # noqa: E402

mb = float(sys.argv[1]) if len(sys.argv) > 1 else 3
emu = _lib.Lib(build.build_emu(), strict=False)
c = corpus.c2_source_tree(total_bytes=int(mb * 1e6), seed=3)
blob = corpus.materialise_host(emu, c)
files = [bytes(blob[int(o):int(o) + int(l)]) for o, l in zip(c.off, c.len)]
big = corpus.c3_huge(n_files=1, file_bytes=400_000, seed=4)
bb = corpus.materialise_host(emu, big)
files.append(bytes(bb[: big.total_bytes]))
for level in (1, 3, 9):
    cctx = emu.zg_cctx_create()
    emu.check(emu.zg_cctx_init(cctx, level))
    emu.check(emu.zg_cctx_set_parameter(cctx, 201, 1))
    emu.check(emu.zg_cctx_reset_archive(cctx, 12))
    r = pack_batch(emu, cctx, files)
    assert r["rc"] == 0
    print(level, sum(r["len"]), hashlib.sha256(bytes(r["frames"])).hexdigest()[:32])
    emu.zg_cctx_free(cctx)
