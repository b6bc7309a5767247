"""Host link ceiling with N GPUs copying at once (what bounds bench.py's e2e at N GPUs): pinned H2D, D2H and both
directions together, all ranks started behind a barrier, device time per rank from CUDA events, aggregate = total
bytes / slowest rank.  Run: torchrun --nproc-per-node N tools/linkbench.py (or plain python for N = 1)."""
import json
import os
import sys

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nb = 1 << 30
    hb, hb2 = torch.empty(nb, dtype=torch.uint8, pin_memory=True), torch.empty(nb, dtype=torch.uint8, pin_memory=True)
    db, db2 = torch.empty(nb, dtype=torch.uint8, device="cuda"), torch.empty(nb, dtype=torch.uint8, device="cuda")
    s2 = torch.cuda.Stream()

    def h2d():
        db.copy_(hb, non_blocking=True)

    def d2h():
        hb.copy_(db, non_blocking=True)

    def both():
        s2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s2):
            hb2.copy_(db2, non_blocking=True)
        db.copy_(hb, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s2)

    res = {}
    for name, fn in (("h2d", h2d), ("d2h", d2h), ("duplex_each_way", both)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(4):
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                fn()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t.item()))
        res[name + "_aggregate_gbs"] = world * 3 * nb / best / 1e6
    if rank == 0:
        out = {"n_gpus": world, **res}
        print(json.dumps(out))
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(f"gpurun_out/host_link_{world}.json", "w"))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
