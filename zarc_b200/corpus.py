"""Synthetic corpora of SURVEY.md §8(d): shapes C1..C4 as (offsets, lengths, kinds, content ids), and
the segment plan the generator kernel / host generator (csrc/corpus.cuh) materialises.

Workload generation only -- not part of the content path.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

KIND_TEXT, KIND_RANDOM, KIND_SRC, KIND_LOG = 0, 1, 2, 3
SEG = 65536
_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z: np.ndarray) -> np.ndarray:
    z = z.astype(np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


@dataclass
class Corpus:
    name: str
    off: np.ndarray  # u64[F] byte offset of each file in the blob
    len: np.ndarray  # u64[F]
    kind: np.ndarray  # u8[F]
    content: np.ndarray  # u64[F] content id (equal ids + equal len => identical bytes)
    seed: int
    align: int

    @property
    def n_files(self) -> int:
        return int(self.len.shape[0])

    @property
    def total_bytes(self) -> int:
        return int(self.len.sum())

    @property
    def blob_bytes(self) -> int:
        return int(self.off[-1] + self.len[-1]) if self.n_files else 0

    def segments(self):
        """(seg_off u64, seg_len u32, seg_kind u8, seg_key u64) for zg_corpus_generate_*."""
        nseg = np.maximum((self.len + np.uint64(SEG - 1)) // np.uint64(SEG), np.uint64(0)).astype(np.int64)
        total = int(nseg.sum())
        file_of = np.repeat(np.arange(self.n_files, dtype=np.int64), nseg)
        first = np.cumsum(nseg) - nseg
        seg_idx = (np.arange(total, dtype=np.int64) - first[file_of]).astype(np.uint64)
        seg_off = self.off[file_of] + seg_idx * np.uint64(SEG)
        seg_len = np.minimum(self.len[file_of] - seg_idx * np.uint64(SEG), np.uint64(SEG)).astype(np.uint32)
        with np.errstate(over="ignore"):
            key = _mix64(
                np.uint64(self.seed) * np.uint64(0x9E3779B97F4A7C15)
                + self.content[file_of] * np.uint64(0xD1B54A32D192ED03)
                + seg_idx * np.uint64(0x8CB92BA72F3D8DD7)
            )
        return (
            np.ascontiguousarray(seg_off, dtype=np.uint64),
            np.ascontiguousarray(seg_len, dtype=np.uint32),
            np.ascontiguousarray(self.kind[file_of], dtype=np.uint8),
            np.ascontiguousarray(key, dtype=np.uint64),
        )

    def subset(self, n: int) -> "Corpus":
        """The first n files, re-based at offset 0 (bounded CPU-baseline samples)."""
        n = min(n, self.n_files)
        return _layout(self.name + f"[:{n}]", self.len[:n].copy(), self.kind[:n].copy(), self.content[:n].copy(), self.seed, self.align)


def _layout(name, lens, kinds, content, seed, align) -> Corpus:
    lens = lens.astype(np.uint64)
    padded = (lens + np.uint64(align - 1)) // np.uint64(align) * np.uint64(align)
    off = np.cumsum(padded) - padded
    return Corpus(name, off.astype(np.uint64), lens, kinds.astype(np.uint8), content.astype(np.uint64), seed, align)


def c1_tree(total_bytes: int = 256 << 20, n_files: int = 2000, seed: int = 1, align: int = 16) -> Corpus:
    """C1: 2 000 files, log-uniform 4 KiB..1 MiB scaled to ~256 MiB, even idx text / odd idx random."""
    rng = np.random.default_rng(seed)
    s = np.exp(rng.uniform(np.log(4096), np.log(1 << 20), n_files))
    s = np.clip(s * (total_bytes / s.sum()), 4096, 1 << 20).astype(np.uint64)
    idx = np.arange(n_files)
    kinds = np.where(idx % 2 == 0, KIND_TEXT, KIND_RANDOM)
    return _layout("C1", s, kinds, idx, seed, align)


def c2_source_tree(total_bytes: int = 10_200_000_000, seed: int = 2, align: int = 16, dup: bool = False) -> Corpus:
    """C2: ~1M files of 1..64 KiB (size = floor(1024 * 64**(u**1.8))), 80% src / 15% text / 5% random.
    dup=True gives C4: every odd file is a byte copy of the preceding even file."""
    mean = 10_240.0
    n_files = max(1, int(total_bytes / mean))
    rng = np.random.default_rng(seed)
    u = rng.uniform(0, 1, n_files)
    s = np.floor(1024.0 * np.power(64.0, np.power(u, 1.8))).astype(np.uint64)
    idx = np.arange(n_files)
    m = idx % 20
    kinds = np.where(m < 16, KIND_SRC, np.where(m < 19, KIND_TEXT, KIND_RANDOM))
    content = idx.copy()
    if dup:
        even = idx - (idx % 2)
        s = s[even]
        kinds = kinds[even]
        content = even
    return _layout("C4" if dup else "C2", s, kinds, content, seed, align)


def c3_huge(n_files: int = 8, file_bytes: int = 1 << 32, seed: int = 3, align: int = 16) -> Corpus:
    """C3: a few huge semi-compressible log files."""
    idx = np.arange(n_files)
    return _layout("C3", np.full(n_files, file_bytes, dtype=np.uint64), np.full(n_files, KIND_LOG), idx, seed, align)


def mixed_small(n_files: int = 64, seed: int = 7, max_len: int = 40000, align: int = 1) -> Corpus:
    """A small ragged mix of every kind incl. empty and 1-byte files (parity tests)."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, max_len, n_files).astype(np.uint64)
    lens[: min(6, n_files)] = np.array([0, 1, 2, 63, 64, 65], dtype=np.uint64)[: min(6, n_files)]
    idx = np.arange(n_files)
    return _layout("mixed", lens, idx % 4, idx, seed, align)


def materialise_host(lib, c: Corpus) -> np.ndarray:
    """Generate the blob on the host with the library's host generator (same code as the kernel)."""
    blob = np.zeros(max(c.blob_bytes, 1), dtype=np.uint8)
    so, sl, sk, key = c.segments()
    if len(so):
        lib.check(lib.zg_corpus_generate_host(blob.ctypes.data, so.ctypes.data, sl.ctypes.data, sk.ctypes.data, key.ctypes.data, len(so)))
    return blob
