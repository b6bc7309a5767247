"""Multi-GPU sharding of the content path: one process per GPU, files dealt to ranks
by a size-balanced greedy partition, no content bytes ever cross GPUs.

The only exchange steps are metadata-sized all-gathers over the process group (NCCL over NVLink on
GPUs, gloo in the CPU tests):
  * 32-byte digests -> global first-occurrence (dedup) decisions, identical to the reference's
    in-order `frames.contains_key` (crates/zarc/src/encode/content_frame.rs:30);
  * 8-byte frame lengths -> exclusive prefix sum in global insertion order -> archive offsets,
    identical to the running `self.offset += bytes` (content_frame.rs:22,45; first frame at 12,
    encode.rs:65).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def _gather_ragged(local: torch.Tensor, counts: list[int], group=None) -> list[torch.Tensor]:
    """all_gather of per-rank tensors with different first dimensions (padded to the max)."""
    world = dist.get_world_size(group)
    mx = max(counts) if counts else 0
    shape = (mx,) + tuple(local.shape[1:])
    padded = torch.zeros(shape, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    return [o[:c] for o, c in zip(out, counts)]


class ShardPlan:
    """Which global file index lives on which rank (every rank computes the same plan)."""

    def __init__(self, lens: np.ndarray, world: int):
        from .corpus import partition_balanced

        self.world = world
        self.n = int(lens.shape[0])
        self.parts = partition_balanced(lens, world)
        self.counts = [int(p.shape[0]) for p in self.parts]

    def mine(self, rank: int) -> np.ndarray:
        return self.parts[rank]


def global_dedup(lib, plan: ShardPlan, local_digests: torch.Tensor, stream: int = 0, group=None):
    """local_digests: u8[n_local, 32] on this rank's device, in the order of plan.mine(rank).
    Returns (first_local u8[n_local], rep_global i64[n_local]): the reference's dedup decision for each
    local file, taken over the GLOBAL input order."""
    dev = local_digests.device
    parts = _gather_ragged(local_digests.contiguous(), plan.counts, group)
    glob = torch.empty((plan.n, 32), dtype=torch.uint8, device=dev)
    for r, p in enumerate(parts):
        glob[torch.from_numpy(plan.parts[r]).to(dev)] = p
    first = torch.empty(plan.n, dtype=torch.uint8, device=dev)
    rep = torch.empty(plan.n, dtype=torch.int64, device=dev)
    if dev.type == "cuda":
        torch.cuda.current_stream().synchronize()
    lib.check(lib.zg_dedup_dev(stream, glob.data_ptr(), plan.n, first.data_ptr(), rep.data_ptr()))
    idx = torch.from_numpy(plan.parts[dist.get_rank(group)]).to(dev)
    return first[idx], rep[idx], first, rep


def global_offsets(lib, plan: ShardPlan, local_frame_len: torch.Tensor, first_global: torch.Tensor, rep_global: torch.Tensor,
                   base: int = 12, stream: int = 0, group=None):
    """local_frame_len: i64[n_local] (0 for files that are not global first occurrences).
    Returns (off_local, len_local, total): Frame.offset / Frame.length for every local file, offsets
    assigned in global insertion order starting at `base`; total = archive offset after the last frame."""
    dev = local_frame_len.device
    parts = _gather_ragged(local_frame_len.contiguous(), plan.counts, group)
    glen = torch.zeros(plan.n, dtype=torch.int64, device=dev)
    for r, p in enumerate(parts):
        glen[torch.from_numpy(plan.parts[r]).to(dev)] = p
    glen = glen * first_global.to(torch.int64)
    goff = torch.empty(plan.n, dtype=torch.int64, device=dev)
    if dev.type == "cuda":
        torch.cuda.current_stream().synchronize()
    lib.check(lib.zg_assign_offsets_dev(stream, glen.data_ptr(), plan.n, base, goff.data_ptr()))
    # duplicates answer with their first occurrence's frame
    goff = goff[rep_global]
    glen_full = glen[rep_global]
    idx = torch.from_numpy(plan.parts[dist.get_rank(group)]).to(dev)
    total = int(base + glen.sum().item())
    return goff[idx], glen_full[idx], total
