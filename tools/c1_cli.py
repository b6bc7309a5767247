"""Config C1 end to end through the CLI, file I/O included: a synthetic 256 MB tree
(2,000 files, 4 KB-1 MB, text/random) in tmpfs -> `zarc-b200 pack` -> .zarc -> `zarc-b200 unpack` -> tree,
wall clock, next to the reference path restated by the oracle (oracle/ref_container.py: the reference's
libzstd-1.5.5 + BLAKE3 call sequence, single-threaded like the reference CLI) on the same files.

usage: python tools/c1_cli.py [--mb 256] [--files 2000] [--level 3] [--no-ref]
Prints one JSON line.  The restored trees are compared byte for byte; the GPU-made archive is also read
back by the oracle (libzstd decodes every frame) and the oracle-made archive by zarc-b200 unpack.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from zarc_b200 import build, corpus, lib as product_lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=256)
    ap.add_argument("--files", type=int, default=2000)
    ap.add_argument("--level", type=int, default=3)
    ap.add_argument("--no-ref", action="store_true")
    args = ap.parse_args()
    lib = product_lib()
    cli = build.build_host()
    base = tempfile.mkdtemp(prefix="zarc_c1_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        c = corpus.c1_tree(total_bytes=int(args.mb * (1 << 20)), n_files=args.files)
        blob = corpus.materialise_host(lib, c)
        tree = os.path.join(base, "tree")
        names = []
        for i, (o, l) in enumerate(zip(c.off, c.len)):
            rel = os.path.join(f"d{i % 40:02d}", f"f{i:05d}.{'txt' if i % 2 == 0 else 'bin'}")
            names.append(rel)
            p = os.path.join(tree, rel)
            os.makedirs(os.path.dirname(p), exist_ok=True)
            blob[int(o) : int(o) + int(l)].tofile(p)
        total = int(c.len.sum())
        env = dict(os.environ, ZARCGPU_LIB=lib.path, ZARC_TIMING="1")

        def cuda_init_s(stderr: bytes) -> float:
            """the CLI's own report of how long creating the CUDA context took (a per-process cost of the box: no persistence daemon)"""
            for ln in stderr.decode().splitlines():
                if "(context) create" in ln:
                    return float(ln.split()[-2]) / 1e3
            return 0.0

        res = {"config": f"C1: {args.files} files, {total / 1e6:.1f} MB, level {args.level}, tmpfs", "bytes": total}
        # warm the device context / page cache with a tiny run so the timed runs measure the steady state
        subprocess.run([cli, "pack", "--output", os.path.join(base, "warm.zarc"), os.path.join("tree", "d00")], cwd=base, env=env, check=True, capture_output=True)
        t0 = time.perf_counter()
        pp = subprocess.run([cli, "pack", "--level", str(args.level), "--output", os.path.join(base, "gpu.zarc"), "tree"], cwd=base, env=env, check=True, capture_output=True)
        t1 = time.perf_counter()
        out = os.path.join(base, "out_gpu")
        os.makedirs(out)
        pu = subprocess.run([cli, "unpack", os.path.join(base, "gpu.zarc")], cwd=out, env=env, check=True, capture_output=True)
        t2 = time.perf_counter()
        res["gpu_cli_pack_s"], res["gpu_cli_unpack_s"] = t1 - t0, t2 - t1
        res["gpu_cli_pack_cuda_init_s"], res["gpu_cli_unpack_cuda_init_s"] = cuda_init_s(pp.stderr), cuda_init_s(pu.stderr)
        res["gpu_cli_pack_gbs_excl_cuda_init"] = total / max(t1 - t0 - res["gpu_cli_pack_cuda_init_s"], 1e-9) / 1e9
        res["gpu_cli_unpack_gbs_excl_cuda_init"] = total / max(t2 - t1 - res["gpu_cli_unpack_cuda_init_s"], 1e-9) / 1e9
        res["gpu_cli_pack_gbs"], res["gpu_cli_unpack_gbs"] = total / (t1 - t0) / 1e9, total / (t2 - t1) / 1e9
        res["gpu_archive_bytes"] = os.path.getsize(os.path.join(base, "gpu.zarc"))
        ok = all(open(os.path.join(out, "tree", r), "rb").read() == open(os.path.join(tree, r), "rb").read() for r in names)
        res["gpu_roundtrip_identical"] = ok
        if not args.no_ref:
            from oracle import ref_container

            t0 = time.perf_counter()
            w = ref_container.RefArchiveWriter(level=args.level)
            for r in names:
                w.add_file(["tree"] + r.split("/"), open(os.path.join(tree, r), "rb").read(), mode=0o100644)
            arc = w.finalise()
            open(os.path.join(base, "ref.zarc"), "wb").write(arc)
            t1 = time.perf_counter()
            ar = ref_container.read_archive(arc)
            out_ref = os.path.join(base, "out_ref")
            good = True
            for f in ar["files"]:
                data, okf = ar["content"](f[2])
                good = good and okf
                p = os.path.join(out_ref, *f[1])
                os.makedirs(os.path.dirname(p), exist_ok=True)
                open(p, "wb").write(data)
            t2 = time.perf_counter()
            res["ref_pack_s"], res["ref_unpack_s"] = t1 - t0, t2 - t1
            res["ref_pack_gbs"], res["ref_unpack_gbs"] = total / (t1 - t0) / 1e9, total / (t2 - t1) / 1e9
            res["ref_archive_bytes"] = len(arc)
            res["ref_cores"] = 1
            res["ref_kind"] = "port (oracle/ref_container.py: libzstd 1.5.5 + BLAKE3, the reference's call sequence, one thread)"
            res["ratio_gpu_over_ref"] = res["gpu_archive_bytes"] / len(arc)
            # cross checks: libzstd restores every GPU-made frame; zarc-b200 unpacks the reference-made archive
            ag = ref_container.read_archive(open(os.path.join(base, "gpu.zarc"), "rb").read())
            res["gpu_archive_read_by_oracle"] = all(ag["content"](f[2])[1] for f in ag["files"] if 2 in f) and len(ag["frames"]) == len(ar["frames"])
            out2 = os.path.join(base, "out_gpu_of_ref")
            os.makedirs(out2)
            subprocess.run([cli, "unpack", os.path.join(base, "ref.zarc")], cwd=out2, env=env, check=True, capture_output=True)
            res["ref_archive_unpacked_by_gpu"] = good and all(
                open(os.path.join(out2, "tree", r), "rb").read() == open(os.path.join(tree, r), "rb").read() for r in names)
        print(json.dumps(res))
    finally:
        shutil.rmtree(base, ignore_errors=True)


if __name__ == "__main__":
    main()
