"""Device-resident pack / unpack timing of the other BASELINE.json corpus shapes (C1, C3, C4) through the
device-pointer C ABI, with the round trip verified on the device.  bench.py measures C2 (the headline)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zarc_b200 import corpus, lib as product_lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="c3", choices=["c1", "c2", "c3", "c4"])
    ap.add_argument("--files", type=int, default=8)
    ap.add_argument("--file-mb", type=float, default=512)
    ap.add_argument("--gb", type=float, default=2.0)
    ap.add_argument("--level", type=int, default=3)
    ap.add_argument("--iters", type=int, default=2)
    a = ap.parse_args()
    lib = product_lib()
    if a.shape == "c3":
        c = corpus.c3_huge(n_files=a.files, file_bytes=int(a.file_mb * (1 << 20)))
    elif a.shape == "c1":
        c = corpus.c1_tree(total_bytes=int(a.gb * 1e9), n_files=max(1, int(a.gb * 1e9 / 131072)))
    else:
        c = corpus.c2_source_tree(total_bytes=int(a.gb * 1e9), dup=a.shape == "c4")
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    s = torch.cuda.current_stream().cuda_stream
    so, sl, sk, key = c.segments()
    blob = torch.empty(c.blob_bytes + 64, dtype=torch.uint8, device="cuda")
    segs = [dev(x) for x in (so, sl, sk, key)]
    lib.check(lib.zg_corpus_generate_dev(s, blob.data_ptr(), *[t.data_ptr() for t in segs], len(so)))
    torch.cuda.synchronize()
    n, B = c.n_files, c.total_bytes
    off, ln = dev(c.off), dev(c.len)
    cctx, dctx = lib.zg_cctx_create(), lib.zg_dctx_create()
    lib.check(lib.zg_cctx_set_stream(cctx, s))
    lib.check(lib.zg_dctx_set_stream(dctx, s))
    lib.check(lib.zg_cctx_init(cctx, 0))
    lib.check(lib.zg_cctx_set_parameter(cctx, 201, 1))
    lib.check(lib.zg_cctx_set_parameter(cctx, 100, a.level))
    cap = B + max(1024, B // 10) + 64 * n
    d_dig = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    d_first = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_foff = torch.empty(n, dtype=torch.int64, device="cuda")
    d_flen = torch.empty(n, dtype=torch.int64, device="cuda")
    d_frames = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(c.blob_bytes + 64, dtype=torch.uint8, device="cuda")
    d_ok = torch.zeros(n, dtype=torch.uint8, device="cuda")
    d_status = torch.zeros(n, dtype=torch.int32, device="cuda")
    nbytes = np.zeros(1, dtype=np.uint64)

    def pack():
        lib.check(lib.zg_cctx_reset_archive(cctx, 12))
        lib.check(lib.zg_pack_batch_dev(cctx, blob.data_ptr(), off.data_ptr(), ln.data_ptr(), n, d_dig.data_ptr(), d_first.data_ptr(),
                                        d_foff.data_ptr(), d_flen.data_ptr(), d_frames.data_ptr(), cap, nbytes.ctypes.data))

    def unpack():
        rel = d_foff - 12
        lib.check(lib.zg_unpack_batch_dev(dctx, d_frames.data_ptr(), int(nbytes[0]), n, rel.data_ptr(), d_flen.data_ptr(), ln.data_ptr(),
                                          d_dig.data_ptr(), d_out.data_ptr(), c.blob_bytes, off.data_ptr(), d_ok.data_ptr(), d_status.data_ptr()))

    def timed(fn):
        best = 1e30
        for _ in range(a.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    pack()
    tp = timed(pack)
    unpack()
    tu = timed(unpack)
    ok = int(d_ok.sum().item()) == n and int(d_status.abs().sum().item()) == 0 and torch.equal(d_out[: c.blob_bytes], blob[: c.blob_bytes])
    uniq = float((ln.view(torch.int64) * d_first.to(torch.int64)).sum().item())
    print(json.dumps({"shape": a.shape, "files": n, "bytes": B, "unique_bytes": uniq, "pack_ms": tp, "unpack_ms": tu, "pack_gbs": B / tp / 1e6,
                      "unpack_gbs": B / tu / 1e6, "ratio_unique": uniq / float(nbytes[0]), "roundtrip_ok": bool(ok)}))
    lib.zg_cctx_free(cctx)
    lib.zg_dctx_free(dctx)


if __name__ == "__main__":
    main()
