#!/usr/bin/env python
"""bench.py -- pack & unpack GB/s (uncompressed) of the zarc content path on B200.

A "step" = one pass of the hot path over one batch: pack the rank's shard of the synthetic corpus
(BLAKE3 digests -> dedup -> Zstandard frame encode -> offsets -> frames) and then unpack it again
(Zstandard frame decode -> checksum -> BLAKE3 verify), device-resident.  `value` = uncompressed
corpus bytes through that pack+unpack round trip per second, summed over ranks (so a corpus of B
bytes per rank that packs in tp and unpacks in tu gives N*B/(tp+tu)); `pack_gbs` / `unpack_gbs`
are the two legs on their own.  `e2e` is the same round trip through the host-buffer C ABI
(zg_pack_batch / zg_unpack_batch) with pinned host buffers and every H2D/D2H copy inside the timed
region.  `--impl reference` times the reference's own CPU path (oracle/ref_path.py: libzstd 1.5.5 +
BLAKE3 with the reference's call sequence) on the host cores.

Launch: `python bench.py --gpus 1` or under torchrun for N > 1 (one rank per GPU, NCCL).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pack+unpack round-trip GB/s (uncompressed)"


# ------------------------------------------------------------------------------------------------
# CPU reference arm (also the cpu_baseline of the GPU arm).  Runs BEFORE any CUDA initialisation.
def _cpu_worker(args):
    """One host core: generate its share of the sample, then time the reference path on it."""
    seed, total_bytes, level, wid, nworkers, lib_path = args
    from oracle import ref_path
    from zarc_b200 import corpus

    c = corpus.c2_source_tree(total_bytes=total_bytes, seed=seed)
    idx = np.arange(wid, c.n_files, nworkers)
    sub = corpus.take(c, idx)
    dll = C.CDLL(lib_path)
    blob = np.zeros(max(sub.blob_bytes, 1), dtype=np.uint8)
    so, sl, sk, key = sub.segments()
    dll.zg_corpus_generate_host.argtypes = [C.c_void_p] * 5 + [C.c_uint64]
    dll.zg_corpus_generate_host(blob.ctypes.data, so.ctypes.data, sl.ctypes.data, sk.ctypes.data, key.ctypes.data, len(so))
    files = [bytes(blob[int(o) : int(o) + int(l)]) for o, l in zip(sub.off, sub.len)]
    out = bytearray()
    t0 = time.perf_counter()
    enc = ref_path.RefEncoder(out, checksum=True, level=level)
    digests = [enc.add_data_frame(f) for f in files]
    t1 = time.perf_counter()
    dec = ref_path.RefDecoder(bytes(out), enc.frames)
    ok = True
    for d, f in zip(digests, files):
        data, good = dec.read_content_frame(d)
        ok = ok and good and len(data) == len(f)
    t2 = time.perf_counter()
    return sub.total_bytes, len(out) - 12, t1 - t0, t2 - t1, ok


def cpu_reference_run(seed: int, sample_bytes: int, level: int, cores: int):
    """All host cores, one process each over disjoint files of the sample (the reference itself is
    single-threaded; this is the generous aggregate).  Returns dict(value GB/s round trip, ...)."""
    import multiprocessing as mp

    lib_path = os.path.join(ROOT, "zarc_b200", "libzarcgpu.so")
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(seed, sample_bytes, level, w, cores, lib_path) for w in range(cores)])
    nbytes = sum(r[0] for r in res)
    cbytes = sum(r[1] for r in res)
    tp = max(r[2] for r in res)
    tu = max(r[3] for r in res)
    assert all(r[4] for r in res), "reference round trip failed"
    return dict(bytes=nbytes, ratio=nbytes / max(cbytes, 1), pack_gbs=nbytes / tp / 1e9, unpack_gbs=nbytes / tu / 1e9,
                value=nbytes / (tp + tu) / 1e9, pack_s=tp, unpack_s=tu)


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(float(s[1])) for s in self.samples if len(s) > 8 and s[1].replace(".", "").isdigit())
        mx = [int(float(s[2])) for s in self.samples if len(s) > 8 and s[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            if len(s) > 8:
                for name, v in zip(names, s[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                    samples=len(self.samples))


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel: str):
    """DRAM bytes per launch from the committed `ncu --set full` capture of this workload, or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel, {}).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--corpus-gb", type=float, default=10.2, help="uncompressed GB per GPU (C2 shape: 10.2 GB = ~1M files)")
    ap.add_argument("--e2e-gb", type=float, default=4.0, help="GB per GPU pushed through the host-buffer ABI per e2e step")
    ap.add_argument("--level", type=int, default=3)
    ap.add_argument("--cpu-sample-mb-per-core", type=float, default=96.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    corpus_bytes = int(args.corpus_gb * 1e9)
    workload = (f"C2 synthetic source-tree corpus: {args.corpus_gb:g} GB uncompressed per GPU (~{int(corpus_bytes / 10240):,} files of "
                f"1-64 KiB, 80% src / 15% text / 5% random), zstd level {args.level}, checksumFlag=1, files dealt to ranks by "
                "size-balanced greedy partition")
    cores = host_cores()

    # -------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        sample = int(min(args.cpu_sample_mb_per_core, 48.0) * 1e6 * cores)
        for _ in range(args.warmup):
            cpu_reference_run(2, min(sample, int(8e6) * cores), args.level, cores)
        vals = []
        t0 = time.perf_counter()
        for _ in range(args.steps):
            vals.append(cpu_reference_run(2, sample, args.level, cores))
        dt = time.perf_counter() - t0
        v = sum(x["bytes"] for x in vals) / sum(x["pack_s"] + x["unpack_s"] for x in vals) / 1e9
        sample_desc = (f"{sample / 1e6:.0f} MB of the same C2 corpus per step, split over {cores} processes, libzstd 1.5.5 + BLAKE3 with the "
                       "reference's call sequence (oracle/ref_path.py), in memory")
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / max(args.steps, 1) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": {"workload": workload},
            "pack_gbs": float(np.mean([x["pack_gbs"] for x in vals])), "unpack_gbs": float(np.mean([x["unpack_gbs"] for x in vals])),
            "ratio": float(np.mean([x["ratio"] for x in vals])),
            "cpu_baseline": {"value": v, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample_desc},
            "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # -------------------------------------------------------------------------------------------
    # CPU baseline first (fork before CUDA is initialised), rank 0 at N=1 only
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = int(args.cpu_sample_mb_per_core * 1e6 * cores)
        cpu = cpu_reference_run(2, sample, args.level, cores)
        cpu["sample"] = (f"{sample / 1e6:.0f} MB of the same C2 corpus, split over {cores} processes; libzstd 1.5.5 + BLAKE3 with the "
                         "reference's call sequence (oracle/ref_path.py), in memory, no file I/O")

    import torch
    import torch.distributed as dist

    from zarc_b200 import corpus, lib as product_lib, parallel

    torch.cuda.set_device(local_rank)
    lib = product_lib()
    assert lib.zg_device_count() > 0, "no CUDA device"
    lib.check(lib.zg_set_device(local_rank))
    # rank 0's stdout carries exactly one JSON line: NCCL (version banner) and anything else that writes to fd 1 from
    # native code goes to stderr; the line itself is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/zarc_bench_nccl.%h.%p.log")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream().cuda_stream

    def to_dev(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    # the rank's shard of the global corpus (weak scaling: N x corpus_gb in total)
    glob = corpus.c2_source_tree(total_bytes=corpus_bytes * world, seed=2)
    plan = parallel.ShardPlan(glob.len, world)
    mine = corpus.take(glob, plan.mine(rank), name=f"C2[rank {rank}]")
    n = mine.n_files
    so, sl, sk, key = mine.segments()
    blob = torch.empty(mine.blob_bytes + 64, dtype=torch.uint8, device=dev)
    segs = [to_dev(x) for x in (so, sl, sk, key)]
    lib.check(lib.zg_corpus_generate_dev(stream, blob.data_ptr(), segs[0].data_ptr(), segs[1].data_ptr(), segs[2].data_ptr(),
                                         segs[3].data_ptr(), len(so)))
    torch.cuda.synchronize()
    del segs
    off, ln = to_dev(mine.off), to_dev(mine.len)
    B = mine.total_bytes

    cctx = lib.zg_cctx_create()
    dctx = lib.zg_dctx_create()
    assert cctx and dctx
    lib.check(lib.zg_cctx_set_stream(cctx, stream))
    lib.check(lib.zg_dctx_set_stream(dctx, stream))
    lib.check(lib.zg_cctx_init(cctx, 0))  # encode.rs:62
    lib.check(lib.zg_cctx_set_parameter(cctx, 201, 1))  # ChecksumFlag(true), pack.rs:227
    lib.check(lib.zg_cctx_set_parameter(cctx, 100, args.level))  # pack.rs:229-232

    cap = B + max(1024, B // 10) + 32 * n
    d_dig = torch.empty(n * 32, dtype=torch.uint8, device=dev)
    d_first = torch.empty(n, dtype=torch.uint8, device=dev)
    d_foff = torch.empty(n, dtype=torch.int64, device=dev)
    d_flen = torch.empty(n, dtype=torch.int64, device=dev)
    d_frames = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_out = torch.empty(mine.blob_bytes + 64, dtype=torch.uint8, device=dev)
    d_ok = torch.zeros(n, dtype=torch.uint8, device=dev)
    d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    nbytes = np.zeros(1, dtype=np.uint64)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def pack_step():
        lib.check(lib.zg_cctx_reset_archive(cctx, 12))
        lib.check(lib.zg_pack_batch_dev(cctx, blob.data_ptr(), off.data_ptr(), ln.data_ptr(), n, d_dig.data_ptr(), d_first.data_ptr(),
                                        d_foff.data_ptr(), d_flen.data_ptr(), d_frames.data_ptr(), cap, nbytes.ctypes.data))
        if world > 1:
            # cross-GPU exchange: frame sizes -> exclusive prefix sum in global insertion order -> archive offsets
            flen_local = d_flen * d_first.to(torch.int64)
            first_g = torch.ones(plan.n, dtype=torch.uint8, device=dev)
            rep_g = torch.arange(plan.n, dtype=torch.int64, device=dev)
            parallel.global_offsets(lib, plan, flen_local, first_g, rep_g, base=12, stream=stream)

    def unpack_step():
        rel = d_foff - 12
        lib.check(lib.zg_unpack_batch_dev(dctx, d_frames.data_ptr(), int(nbytes[0]), n, rel.data_ptr(), d_flen.data_ptr(), ln.data_ptr(),
                                          d_dig.data_ptr(), d_out.data_ptr(), mine.blob_bytes, off.data_ptr(), d_ok.data_ptr(),
                                          d_status.data_ptr()))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        pack_step()
        unpack_step()
    sync_all()
    # correctness of what is being timed: digests verified on device, bytes identical
    assert int(d_ok.sum().item()) == n and int(d_status.abs().sum().item()) == 0, "unpack verification failed"
    assert torch.equal(d_out[: mine.blob_bytes], blob[: mine.blob_bytes]), "round trip is not byte-identical"
    ratio = B / float(nbytes[0])

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    if os.environ.get("ZG_DECODE_BATCHING"):  # tuning aid: "cap_div,min_batch_bytes,floor_bytes"
        a, b, c3 = os.environ["ZG_DECODE_BATCHING"].split(",")
        lib.dll.zg_internal_set_decode_batching(C.c_uint32(int(a)), C.c_uint32(int(b)), C.c_uint32(int(c3)))
    if os.environ.get("ZG_SLICE_MB"):  # tuning aid: host-API slice size
        lib.dll.zg_internal_set_slice_bytes(C.c_uint64(int(os.environ["ZG_SLICE_MB"]) << 20))
    if os.environ.get("ZG_PACK_SLICE_MB"):  # tuning aid: host-API slice size, pack only
        lib.dll.zg_internal_set_pack_slice_bytes(C.c_uint64(int(os.environ["ZG_PACK_SLICE_MB"]) << 20))
    lib.zg_profile_enable(1)
    launches0 = lib.zg_kernel_launch_count()
    sync_all()
    e0, e1 = ev(), ev()
    pack_ms, unpack_ms = [], []
    e0.record()
    for _ in range(args.steps):
        a, b, c = ev(), ev(), ev()
        a.record()
        pack_step()
        b.record()
        unpack_step()
        c.record()
        pack_ms.append((a, b))
        unpack_ms.append((b, c))
    e1.record()
    sync_all()
    total_ms = e0.elapsed_time(e1)
    launches = lib.zg_kernel_launch_count() - launches0
    lib.zg_profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    tp = sum(x.elapsed_time(y) for x, y in pack_ms) / args.steps
    tu = sum(x.elapsed_time(y) for x, y in unpack_ms) / args.steps

    def rank_max(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def rank_sum(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    total_ms = rank_max(total_ms)
    tp, tu = rank_max(tp), rank_max(tu)
    Btot = rank_sum(float(B))
    ms_per_step = total_ms / args.steps
    value = Btot / (ms_per_step * 1e-3) / 1e9

    # per-kernel roofline of the dominant kernel (device time from CUDA events on the launching stream)
    prof = {}
    names = ((0, "k_blake3_small"), (6, "k_zstd_match_blocks"), (7, "k_zstd_literals"), (8, "k_zstd_sequences"), (2, "k_zstd_decode_frames"),
             (1, "encode pass (match+literals+sequences, all chunks)"))
    for k, name in names:
        ms, cnt = C.c_double(0), C.c_uint64(0)
        lib.zg_profile_read(k, C.byref(ms), C.byref(cnt))
        prof[name] = (ms.value, cnt.value)
    peak, peak_src = measured_peak_hbm()
    C_bytes = float(nbytes[0])
    # algorithmic bytes per STEP of each kernel class: the encoder reads the unique input once (match) and
    # writes the compressed bytes once (literals + sequences sections); the decoder reads C and writes N; BLAKE3 reads N
    # (pack digests and unpack verification: two launches per step)
    Bu = float(B)  # the bench corpus has no duplicate files
    alg_step = {"k_blake3_small": 2.0 * B, "k_zstd_match_blocks": Bu, "k_zstd_literals": C_bytes, "k_zstd_sequences": C_bytes,
                "k_zstd_decode_frames": C_bytes + B, "encode pass (match+literals+sequences, all chunks)": Bu + C_bytes}
    single = [nm for _, nm in names[:5]]
    dom = max(single, key=lambda k: prof[k][0])
    kernels = {}
    for k, v in prof.items():
        ms_step = v[0] / args.steps
        kernels[k] = {"ms_per_step": ms_step, "launches_per_step": v[1] / args.steps,
                      "ms_per_launch": (v[0] / v[1] if v[1] else None),
                      "algorithmic_gbs": (alg_step[k] / (ms_step * 1e-3) / 1e9 if ms_step > 0 else None),
                      "share_of_step": (ms_step / ms_per_step if ms_per_step else None)}
    lps = max(prof[dom][1] / args.steps, 1)
    dom_ms = prof[dom][0] / max(prof[dom][1], 1)
    alg_launch = alg_step[dom] / lps
    achieved = alg_launch / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(dom), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_launch, "kernels": kernels,
                "note": "integer/latency-bound kernels (bitstream decode, match finding, entropy coding); the HBM roofline is the "
                        "ceiling the north star names; BLAKE3 is bound by the INT32 ALU pipe (see profiles/README.md)"}

    # -------------------------------------------------------------------------------------------
    # end to end through the host-buffer C ABI (pinned host memory, copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        sub = mine.subset(max(1, int(n * min(1.0, args.e2e_gb * 1e9 / max(B, 1)))))
        ne, Be, span = sub.n_files, sub.total_bytes, sub.blob_bytes
        h_blob = torch.empty(span + 64, dtype=torch.uint8, pin_memory=True)
        h_blob[:span].copy_(blob[:span])
        torch.cuda.synchronize()
        cap_e = Be + max(1024, Be // 10) + 32 * ne
        h_frames = torch.empty(cap_e, dtype=torch.uint8, pin_memory=True)
        h_out = torch.empty(Be + 64, dtype=torch.uint8, pin_memory=True)
        h_dig = torch.empty(ne * 32, dtype=torch.uint8, pin_memory=True)
        h_first = torch.empty(ne, dtype=torch.uint8, pin_memory=True)
        h_foff = torch.empty(ne, dtype=torch.int64, pin_memory=True)
        h_flen = torch.empty(ne, dtype=torch.int64, pin_memory=True)
        h_ok = torch.empty(ne, dtype=torch.uint8, pin_memory=True)
        h_status = torch.empty(ne, dtype=torch.int32, pin_memory=True)
        h_off = torch.from_numpy(sub.off.astype(np.int64)).pin_memory()
        h_len = torch.from_numpy(sub.len.astype(np.int64)).pin_memory()
        nb_e = np.zeros(1, dtype=np.uint64)

        e2e_t = [0.0, 0.0]

        def e2e_step():
            t0 = time.perf_counter()
            lib.check(lib.zg_cctx_reset_archive(cctx, 12))
            lib.check(lib.zg_pack_batch(cctx, h_blob.data_ptr(), h_off.data_ptr(), h_len.data_ptr(), ne, h_dig.data_ptr(), h_first.data_ptr(),
                                        h_foff.data_ptr(), h_flen.data_ptr(), h_frames.data_ptr(), cap_e, nb_e.ctypes.data))
            t1 = time.perf_counter()
            rel = h_foff - 12
            lib.check(lib.zg_unpack_batch(dctx, h_frames.data_ptr(), int(nb_e[0]), ne, rel.data_ptr(), h_flen.data_ptr(), h_len.data_ptr(),
                                          h_dig.data_ptr(), h_out.data_ptr(), Be, None, h_ok.data_ptr(), h_status.data_ptr()))
            e2e_t[0] += t1 - t0  # both calls return with their results on the host (blocking API)
            e2e_t[1] += time.perf_counter() - t1

        for _ in range(2):
            e2e_step()
        sync_all()
        assert int(h_ok.sum().item()) == ne, "e2e verification failed"
        s0, s1 = ev(), ev()
        ksteps = max(2, min(args.steps, 5))
        e2e_t[0] = e2e_t[1] = 0.0
        s0.record()
        for _ in range(ksteps):
            e2e_step()
        s1.record()
        sync_all()
        e_ms = rank_max(s0.elapsed_time(s1)) / ksteps
        Ce = float(nb_e[0])
        e2e = {"value": rank_sum(float(Be)) / (e_ms * 1e-3) / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": int(span + 16 * ne + Ce + 56 * ne), "d2h_bytes_per_step": int(Ce + 49 * ne + Be + 5 * ne),
               "ms_per_step": e_ms, "pack_ms": e2e_t[0] / ksteps * 1e3, "unpack_ms": e2e_t[1] / ksteps * 1e3,
               "workload_gb_per_gpu": Be / 1e9,
               "api": "zg_pack_batch + zg_unpack_batch (host buffers, pinned), digests verified"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": workload, "files_per_gpu": n, "bytes_per_gpu": B, "l2": "inputs (>= 10x L2) larger than L2; no flush needed",
                       "parallelism": f"frames sharded over {world} GPU(s); NCCL all-gather of frame sizes only"},
            "pack_gbs": Btot / (tp * 1e-3) / 1e9, "unpack_gbs": Btot / (tu * 1e-3) / 1e9,
            "ratio": ratio, "ratio_reference_level3": cpu["ratio"] if cpu else None,
            "roofline": roofline,
            "cpu_baseline": ({"value": cpu["value"], "unit": "GB/s", "cores": cores, "kind": "port", "sample": cpu["sample"],
                              "pack_gbs": cpu["pack_gbs"], "unpack_gbs": cpu["unpack_gbs"]} if cpu else None),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    lib.zg_cctx_free(cctx)
    lib.zg_dctx_free(dctx)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
