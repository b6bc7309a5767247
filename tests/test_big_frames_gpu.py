"""Frames above the single-segment limit (2^27 B) and above 2^32 B on the real GPU (VERDICT r1 "untested configs"):
the windowed frames this encoder writes must be accepted by libzstd's STREAMING decoder (what zstd_iterator.rs:29
creates: it refuses windows above 2^27), BLAKE3's 64-bit chunk counter and the 8-byte Frame_Content_Size are exercised at
2^32 bytes, and reference-made frames of thousands of dependent blocks decode through the staged pipeline."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dev_corpus(lib, torch, c):
    so, sl, sk, key = c.segments()
    blob = torch.empty(c.blob_bytes + 64, dtype=torch.uint8, device="cuda")
    segs = [torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (so, sl, sk, key)]
    lib.check(lib.zg_corpus_generate_dev(0, blob.data_ptr(), *[t.data_ptr() for t in segs], len(so)))
    torch.cuda.synchronize()
    return blob


def _pack_dev(lib, torch, blob, c, level=3):
    n = c.n_files
    off = torch.from_numpy(c.off.astype(np.int64)).cuda()
    ln = torch.from_numpy(c.len.astype(np.int64)).cuda()
    cctx = lib.zg_cctx_create()
    lib.check(lib.zg_cctx_init(cctx, 0))
    lib.check(lib.zg_cctx_set_parameter(cctx, 201, 1))
    lib.check(lib.zg_cctx_set_parameter(cctx, 100, level))
    lib.check(lib.zg_cctx_reset_archive(cctx, 0))
    cap = c.total_bytes + c.total_bytes // 10 + 4096
    dig = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    first = torch.empty(n, dtype=torch.uint8, device="cuda")
    foff = torch.empty(n, dtype=torch.int64, device="cuda")
    flen = torch.empty(n, dtype=torch.int64, device="cuda")
    frames = torch.empty(cap, dtype=torch.uint8, device="cuda")
    nbytes = np.zeros(1, dtype=np.uint64)
    lib.check(lib.zg_pack_batch_dev(cctx, blob.data_ptr(), off.data_ptr(), ln.data_ptr(), n, dig.data_ptr(), first.data_ptr(), foff.data_ptr(),
                                    flen.data_ptr(), frames.data_ptr(), cap, nbytes.ctypes.data))
    lib.zg_cctx_free(cctx)
    return dict(off=off, len=ln, dig=dig, foff=foff, flen=flen, frames=frames[: int(nbytes[0])], nbytes=int(nbytes[0]))


def _unpack_dev(lib, torch, arch, foff, flen, ulen, dig, out_off, out_bytes):
    n = int(foff.shape[0])
    out = torch.empty(out_bytes + 64, dtype=torch.uint8, device="cuda")
    ok = torch.zeros(n, dtype=torch.uint8, device="cuda")
    st = torch.zeros(n, dtype=torch.int32, device="cuda")
    dctx = lib.zg_dctx_create()
    rc = lib.zg_unpack_batch_dev(dctx, arch.data_ptr(), int(arch.shape[0]), n, foff.data_ptr(), flen.data_ptr(), ulen.data_ptr(), dig.data_ptr(),
                                 out.data_ptr(), out_bytes, out_off.data_ptr(), ok.data_ptr(), st.data_ptr())
    stats = (C.c_uint64 * 8)()
    lib.dll.zg_internal_decode_stats_ex(stats)
    lib.zg_dctx_free(dctx)
    return rc, out, ok.cpu().tolist(), st.cpu().tolist(), list(stats)


def _same(torch, a, b, n):
    return all(bool(torch.equal(a[o : min(n, o + (1 << 30))], b[o : min(n, o + (1 << 30))])) for o in range(0, n, 1 << 30))


@pytest.mark.parametrize("mib", [130, 200])
def test_windowed_frames_are_restored_by_the_reference_streaming_decoder(gpu, mib):
    import blake3
    import torch

    from oracle import ref_path
    from zarc_b200 import corpus

    c = corpus.c3_huge(n_files=1, file_bytes=mib << 20, seed=40 + mib)
    blob = _dev_corpus(gpu, torch, c)
    p = _pack_dev(gpu, torch, blob, c)
    frame = p["frames"].cpu().numpy()
    assert frame[4] & 0x20 == 0, "a frame above 2^27 bytes must not be single-segment (App. C: window limit)"
    host = blob[: c.total_bytes].cpu().numpy()
    want = blake3.blake3(host.tobytes(), max_threads=blake3.blake3.AUTO).digest()
    assert bytes(p["dig"].cpu().numpy()) == want
    # the reference's own call sequence (zstd_iterator.rs:88-153 + frame_iterator.rs:94-103) on the GPU-made frame
    dig, produced, consumed = ref_path.ref_stream_digest(frame, 0)
    assert (dig, produced, consumed) == (want, c.total_bytes, len(frame))
    # and this library's decoder on it
    rc, out, ok, st, stats = _unpack_dev(gpu, torch, p["frames"], p["foff"], p["flen"], p["len"], p["dig"], p["off"], c.blob_bytes)
    assert rc == 0 and ok == [1] and st == [0] and _same(torch, out, blob, c.total_bytes)
    assert stats[2] == 0 and stats[3] == 1 and stats[5] == 0  # staged, no block waited for another


def test_reference_frame_of_thousands_of_dependent_blocks(gpu):
    """200 MiB of log lines compressed by libzstd the reference's way: 1 600 blocks chained by cross-block matches,
    repeat offsets, Treeless literals and Repeat_Mode tables -- decoded once, by the staged pipeline."""
    import torch

    from oracle import ref_path
    from zarc_b200 import corpus

    c = corpus.c3_huge(n_files=1, file_bytes=200 << 20, seed=77)
    blob = _dev_corpus(gpu, torch, c)
    host = blob[: c.total_bytes].cpu().numpy().tobytes()
    for level in (1, 3):
        frame = ref_path.ref_compress(host, level=level)
        arch = torch.from_numpy(np.frombuffer(frame, dtype=np.uint8).copy()).cuda()
        z = torch.zeros(1, dtype=torch.int64, device="cuda")
        fl = torch.tensor([len(frame)], dtype=torch.int64, device="cuda")
        ul = torch.tensor([len(host)], dtype=torch.int64, device="cuda")
        dig = torch.from_numpy(np.frombuffer(ref_path._blake3(host), dtype=np.uint8).copy()).cuda()
        rc, out, ok, st, stats = _unpack_dev(gpu, torch, arch, z, fl, ul, dig, z, len(host))
        assert rc == 0 and ok == [1] and st == [0] and _same(torch, out, blob, len(host))
        assert stats[2] == 0 and stats[3] == 1 and stats[4] == 1600 and stats[5] > 1000


def test_4gib_file_digest_frame_and_round_trip(gpu):
    """One file of exactly 2^32 bytes: BLAKE3 with a chunk counter past 2^22, an 8-byte Frame_Content_Size, 32 768 blocks."""
    import blake3
    import torch

    from oracle import ref_path
    from zarc_b200 import corpus

    c = corpus.c3_huge(n_files=1, file_bytes=1 << 32, seed=9)
    blob = _dev_corpus(gpu, torch, c)
    p = _pack_dev(gpu, torch, blob, c)
    host = blob[: c.total_bytes].cpu().numpy()
    h = blake3.blake3(max_threads=blake3.blake3.AUTO)
    for o in range(0, c.total_bytes, 1 << 28):
        h.update(host[o : o + (1 << 28)].tobytes())
    want = h.digest()
    del host
    assert bytes(p["dig"].cpu().numpy()) == want
    frame = p["frames"].cpu().numpy()
    assert frame[4] >> 6 == 3, "Frame_Content_Size of 2^32 needs the 8-byte field"
    dig, produced, consumed = ref_path.ref_stream_digest(frame, 0)  # libzstd streaming, hashing every chunk like FrameIterator
    assert (dig, produced, consumed) == (want, 1 << 32, len(frame))
    rc, out, ok, st, stats = _unpack_dev(gpu, torch, p["frames"], p["foff"], p["flen"], p["len"], p["dig"], p["off"], c.blob_bytes)
    assert rc == 0 and ok == [1] and st == [0] and _same(torch, out, blob, c.total_bytes)
    assert stats[4] == 32768
