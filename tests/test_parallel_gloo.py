"""N>1 path on CPU: world_size-2 gloo processes run the sharded pack (partition -> local digests ->
all-gather -> global dedup -> local encode -> all-gather of frame sizes -> global offsets) with the
SIMT-emulator build standing in for the GPU, and the result must equal the reference path run
serially over the whole ordered file list (dedup decisions, offsets, digests, round trip)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _files(dups=True):
    from tests.golden.recipes import text, rand

    fs = [text(3000 + 137 * i, i) if i % 3 else rand(500 + 91 * i, i) for i in range(14)]
    if dups:
        fs[5] = fs[1]  # duplicates that land on different ranks
        fs[9] = fs[2]
        fs[12] = fs[1]
    fs.append(b"")
    return fs


def _worker(rank, world, port, q, mode="greedy", dups=True):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.helpers import layout
    from zarc_b200 import _lib, build, parallel

    lib = _lib.Lib(build.build_emu(), strict=False)
    files = _files(dups)
    lens = np.array([len(f) for f in files], dtype=np.uint64)
    plan = parallel.ShardPlan(lens, world, mode=mode)
    assert plan.contiguous == (mode == "contiguous")
    mine = plan.mine(rank)
    local = [files[i] for i in mine]
    blob, off, ln = layout(local, 16)
    n = len(local)
    t = lambda a: torch.from_numpy(a)
    d_blob, d_off, d_len = t(blob), t(off.astype(np.int64)), t(ln.astype(np.int64))
    # (1) local digests
    dig = torch.zeros((n, 32), dtype=torch.uint8)
    lib.check(lib.zg_blake3_batch_dev(0, d_blob.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), n, dig.data_ptr()))
    # (2) global first-occurrence decisions
    first_l, rep_l, first_g, rep_g = parallel.global_dedup(lib, plan, dig)
    # (3) encode the local files that are global first occurrences: add_data_frame with the digest (content_frame.rs:26)
    #     and the first-occurrence decision (:30) handed in
    cctx = lib.zg_cctx_create()
    lib.check(lib.zg_cctx_init(cctx, 0))
    lib.check(lib.zg_cctx_set_parameter(cctx, 201, 1))
    lib.check(lib.zg_cctx_reset_archive(cctx, 0))
    cap = int(d_len.sum()) + 4096 + 64 * n
    frames = torch.zeros(cap, dtype=torch.uint8)
    foff = torch.zeros(max(n, 1), dtype=torch.int64)
    flen = torch.zeros(max(n, 1), dtype=torch.int64)
    first_out = torch.zeros(max(n, 1), dtype=torch.uint8)
    nbytes = np.zeros(1, dtype=np.uint64)
    sel = first_l.contiguous()
    lib.check(lib.zg_pack_batch_dev_ex(cctx, d_blob.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), n, dig.data_ptr(), sel.data_ptr(), None,
                                       first_out.data_ptr(), foff.data_ptr(), flen.data_ptr(), frames.data_ptr(), cap, nbytes.ctypes.data))
    lib.zg_cctx_free(cctx)
    assert torch.equal(first_out[:n], sel)
    local_flen = flen[:n] * sel.to(torch.int64)  # (an unselected file answers with an earlier local copy's frame, if any, else 0)
    keep = torch.nonzero(first_l).flatten()
    # (4) global archive offsets
    g_off, g_len, total = parallel.global_offsets(lib, plan, local_flen, first_g, rep_g, base=12, no_duplicates=not dups)
    total = int(total)
    # ship everything to the parent for checking
    local_frames = {}
    for j, i in enumerate(keep.tolist()):
        local_frames[int(mine[i])] = bytes(frames[int(foff[i]) : int(foff[i]) + int(flen[i])].numpy())
    q.put((rank, mine.tolist(), [bytes(d.numpy()) for d in dig], first_l.tolist(), g_off.tolist(), g_len.tolist(), total, local_frames))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,dups", [("greedy", True), ("contiguous", True), ("contiguous", False)])
def test_two_rank_sharded_pack_matches_serial_reference(mode, dups):
    from oracle import ref_path
    from zarc_b200 import build

    build.build_emu()
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, mode, dups)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    files = _files(dups)
    n = len(files)
    digest, first, off, ln, frames = [None] * n, [None] * n, [None] * n, [None] * n, {}
    total = None
    for rank, mine, digs, fl, go, gl, tot, lf in results:
        for j, i in enumerate(mine):
            digest[i], first[i], off[i], ln[i] = digs[j], fl[j], go[j], gl[j]
        frames.update(lf)
        total = tot if total is None else total
        assert tot == total
    # serial reference over the same ordered list
    out = bytearray()
    enc = ref_path.RefEncoder(out, level=3)
    ref_d = [enc.add_data_frame(f) for f in files]
    assert digest == ref_d
    seen, ref_first = set(), []
    for d in ref_d:
        ref_first.append(0 if d in seen else 1)
        seen.add(d)
    assert first == ref_first
    pos = 12
    for i in range(n):
        if ref_first[i]:
            assert off[i] == pos and ln[i] == len(frames[i])
            assert ref_path.ref_decompress(frames[i], len(files[i])) == files[i]
            pos += ln[i]
        else:
            j = ref_d.index(ref_d[i])
            assert (off[i], ln[i]) == (off[j], ln[j])
    assert total == pos


def test_partition_is_balanced_and_complete():
    from zarc_b200 import corpus

    c = corpus.c2_source_tree(total_bytes=50_000_000, seed=4)
    for world in (2, 4, 8):
        parts = corpus.partition_balanced(c.len, world)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(c.n_files))
        sizes = [int(c.len[p].sum()) for p in parts]
        assert max(sizes) - min(sizes) <= 65536 + 0.001 * max(sizes)
        for p in parts:
            assert np.all(np.diff(p) > 0)  # input order preserved inside a rank
    c3 = corpus.c3_huge(n_files=8, file_bytes=1 << 20)
    assert sorted(len(p) for p in corpus.partition_balanced(c3.len, 8)) == [1] * 8


def test_contiguous_partition_and_plan_choice():
    from zarc_b200 import corpus, parallel

    c = corpus.c2_source_tree(total_bytes=50_000_000, seed=4)
    for world in (1, 2, 4, 8):
        parts = parallel.partition_contiguous(c.len, world)
        assert np.array_equal(np.concatenate(parts), np.arange(c.n_files))
        sizes = [int(c.len[p].sum()) for p in parts]
        assert max(sizes) - min(sizes) <= 2 * 65536
    assert parallel.ShardPlan(c.len, 8).contiguous  # many small files: contiguous ranges are even
    c3 = corpus.c3_huge(n_files=8, file_bytes=1 << 20)
    assert [len(p) for p in parallel.partition_contiguous(c3.len, 8)] == [1] * 8
    # sizes sorted descending: contiguous ranges cannot be even, the plan falls back to the greedy deal
    skew = np.sort(c.len[:64])[::-1].copy()
    plan = parallel.ShardPlan(skew, 8)
    assert parallel.ShardPlan(np.array([1 << 20, 10, 10, 10], dtype=np.uint64), 2).contiguous is False
    assert np.array_equal(np.sort(np.concatenate(plan.parts)), np.arange(64))
    empty = parallel.partition_contiguous(np.zeros(5, dtype=np.uint64), 2)
    assert np.array_equal(np.concatenate(empty), np.arange(5))
