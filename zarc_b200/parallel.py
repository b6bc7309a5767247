"""Multi-GPU sharding of the content path: one process per GPU, files dealt to ranks,
no content bytes ever cross GPUs.

Partition: contiguous byte-balanced ranges of the input order when that is even to within 1 % (then
every rank's frames form ONE contiguous span of the archive and the ordered write is one big write per
rank), else the size-balanced greedy deal (`corpus.partition_balanced`).

The only exchange steps are metadata-sized all-gathers over the process group (NCCL over NVLink on
GPUs, gloo in the CPU tests), ONE collective each, into tensors laid out by a permutation that is
computed once per plan:
  * 32-byte digests -> global first-occurrence (dedup) decisions, identical to the reference's
    in-order `frames.contains_key` (crates/zarc/src/encode/content_frame.rs:30);
  * 8-byte frame lengths -> exclusive prefix sum in global insertion order -> archive offsets,
    identical to the running `self.offset += bytes` (content_frame.rs:22,45; first frame at 12,
    encode.rs:65).  When no content is duplicated across the corpus and the partition is contiguous,
    the per-rank TOTALS are enough (8 bytes per rank).
Nothing here reads a device value on the host (no `.item()`): totals are returned as device tensors.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def partition_contiguous(lens: np.ndarray, nranks: int) -> list[np.ndarray]:
    """Byte-balanced contiguous ranges of the input order: rank r gets the files whose first byte (in the
    concatenation of all files) falls into the r-th N-th of the total."""
    n = int(lens.shape[0])
    csum = np.cumsum(lens.astype(np.uint64))
    total = int(csum[-1]) if n else 0
    start = csum - lens.astype(np.uint64)
    if total == 0:
        bounds = [n * r // nranks for r in range(nranks + 1)]  # all files empty: split by count
    else:
        cuts = [np.uint64((total * r + nranks - 1) // nranks) for r in range(nranks + 1)]
        bounds = [int(np.searchsorted(start, c, side="left")) for c in cuts]
    bounds[0], bounds[-1] = 0, n
    return [np.arange(bounds[r], bounds[r + 1], dtype=np.int64) for r in range(nranks)]


class ShardPlan:
    """Which global file index lives on which rank (every rank computes the same plan)."""

    def __init__(self, lens: np.ndarray, world: int, mode: str = "auto"):
        from .corpus import partition_balanced

        self.world = world
        self.n = int(lens.shape[0])
        parts = None
        if mode in ("auto", "contiguous"):
            parts = partition_contiguous(lens, world)
            sizes = [int(lens[p].sum()) for p in parts]
            even = max(sizes) <= 1.01 * (sum(sizes) / max(world, 1)) + 1
            if mode == "auto" and not even:
                parts = None
        self.contiguous = parts is not None
        self.parts = parts if parts is not None else partition_balanced(lens, world)
        self.counts = [int(p.shape[0]) for p in self.parts]
        self.maxc = max(self.counts) if self.counts else 0
        # slot r * maxc + j of a padded all-gather holds global file parts[r][j]
        self._slots = np.concatenate([r * self.maxc + np.arange(c, dtype=np.int64) for r, c in enumerate(self.counts)]) if self.n else np.zeros(0, np.int64)
        self._globals = np.concatenate(self.parts) if self.n else np.zeros(0, np.int64)
        self._dev = {}

    def mine(self, rank: int) -> np.ndarray:
        return self.parts[rank]

    def tensors(self, device, rank: int):
        """(slots, globals, mine) as int64 tensors on `device`, cached."""
        key = (str(device), rank)
        if key not in self._dev:
            self._dev[key] = (torch.from_numpy(self._slots).to(device), torch.from_numpy(self._globals).to(device),
                              torch.from_numpy(np.ascontiguousarray(self.parts[rank])).to(device))
        return self._dev[key]


def _gather_global(plan: ShardPlan, local: torch.Tensor, group=None) -> torch.Tensor:
    """One all-gather of every rank's per-file rows (padded to the largest shard), returned in GLOBAL file order."""
    rank = dist.get_rank(group)
    slots, globs, _ = plan.tensors(local.device, rank)
    row = tuple(local.shape[1:])
    padded = local
    if local.shape[0] != plan.maxc:
        padded = torch.zeros((plan.maxc,) + row, dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    out = torch.empty((plan.world * plan.maxc,) + row, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    if plan.contiguous and all(c == plan.maxc for c in plan.counts):
        return out  # rank-major order IS global order
    glob = torch.empty((plan.n,) + row, dtype=local.dtype, device=local.device)
    glob[globs] = out[slots]
    return glob


def _stream_sync(t: torch.Tensor):
    if t.device.type == "cuda":
        torch.cuda.current_stream().synchronize()


def global_dedup(lib, plan: ShardPlan, local_digests: torch.Tensor, stream: int = 0, group=None):
    """local_digests: u8[n_local, 32] on this rank's device, in the order of plan.mine(rank).
    Returns (first_local u8[n_local], rep_local i64[n_local], first_global, rep_global): the reference's dedup
    decision for each file, taken over the GLOBAL input order."""
    dev = local_digests.device
    glob = _gather_global(plan, local_digests.reshape(-1, 32).contiguous(), group)
    first = torch.empty(plan.n, dtype=torch.uint8, device=dev)
    rep = torch.empty(plan.n, dtype=torch.int64, device=dev)
    _stream_sync(glob)
    lib.check(lib.zg_dedup_dev(stream, glob.data_ptr(), plan.n, first.data_ptr(), rep.data_ptr()))
    idx = plan.tensors(dev, dist.get_rank(group))[2]
    return first[idx], rep[idx], first, rep


def global_offsets(lib, plan: ShardPlan, local_frame_len: torch.Tensor, first_global: torch.Tensor | None, rep_global: torch.Tensor | None,
                   base: int = 12, stream: int = 0, group=None, no_duplicates: bool = False):
    """local_frame_len: i64[n_local] (0 for files that are not global first occurrences).
    Returns (off_local, len_local, total): Frame.offset / Frame.length for every local file, offsets assigned in
    global insertion order starting at `base`; total (0-dim device tensor) = archive offset after the last frame.

    no_duplicates=True (the caller knows every file is a first occurrence) with a contiguous partition needs only
    the per-rank totals: 8 bytes per rank cross the link."""
    dev = local_frame_len.device
    rank = dist.get_rank(group)
    n_local = int(local_frame_len.shape[0])
    if no_duplicates and plan.contiguous:
        tot = local_frame_len.sum().reshape(1)
        alltot = torch.empty(plan.world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(alltot, tot, group=group)
        before = torch.cumsum(alltot, 0) - alltot  # exclusive prefix over ranks
        off = torch.empty(max(n_local, 1), dtype=torch.int64, device=dev)
        _stream_sync(off)
        if n_local:
            lib.check(lib.zg_assign_offsets_dev(stream, local_frame_len.data_ptr(), n_local, base, off.data_ptr()))
        return off[:n_local] + before[rank], local_frame_len, alltot.sum() + base
    glen = _gather_global(plan, local_frame_len.contiguous(), group)
    if first_global is not None:
        glen = glen * first_global.to(torch.int64)
    goff = torch.empty(max(plan.n, 1), dtype=torch.int64, device=dev)
    _stream_sync(glen)
    if plan.n:
        lib.check(lib.zg_assign_offsets_dev(stream, glen.data_ptr(), plan.n, base, goff.data_ptr()))
    idx = plan.tensors(dev, rank)[2]
    if rep_global is not None:  # duplicates answer with their first occurrence's frame
        ridx = rep_global[idx]
        return goff[ridx], glen[ridx], glen.sum() + base
    return goff[idx], glen[idx], glen.sum() + base
