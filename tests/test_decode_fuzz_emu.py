"""Mutated frames through the decoder (SIMT emulator): no crash, no hang, and whenever the GPU decoder accepts a frame
the bytes it produced are the bytes the reference's streaming decoder produces for the same frame.  Covers the serial path, the block items of
split frames and the serial redo (zstd_decode.cu), on frames made by libzstd and by this library's encoder."""
import numpy as np

from oracle import ref_path
from tests.golden.recipes import rand, text
from tests.helpers import unpack_batch
from tests.test_split_decode_emu import _b3, _own_frames, _raw_frame

BLOCK = 128 * 1024


def _mutations(frame, rng, k):
    out = []
    n = len(frame)
    for _ in range(k):
        f = bytearray(frame)
        kind = int(rng.integers(0, 6))
        if kind <= 2:  # flip 1-3 bytes anywhere (headers, tables, bitstreams, checksum)
            for _ in range(kind + 1):
                p = int(rng.integers(0, n))
                f[p] ^= int(rng.integers(1, 256))
        elif kind == 3:  # truncate
            f = f[: int(rng.integers(1, n))]
        elif kind == 4:  # flip inside the first 24 bytes (frame / block / section headers)
            p = int(rng.integers(0, min(24, n)))
            f[p] ^= 1 << int(rng.integers(0, 8))
        else:  # overwrite a short run
            p = int(rng.integers(0, n))
            m = int(rng.integers(1, 9))
            f[p : p + m] = bytes(rng.integers(0, 256, min(m, n - p), dtype=np.uint8))
        out.append(bytes(f))
    return out


def _reference_verdict(frame, ulen):
    """What the reference restores from this frame: its own call sequence (a streaming DCtx started at the frame's
    offset, zstd_iterator.rs:88-153), which stops at the end of the frame and never looks at what follows -- plus the
    size check against Frame.uncompressed that the directory implies."""
    try:
        d = ref_path.ref_decompress_stream(bytes(frame), 0)
        return d if len(d) == ulen else None
    except ref_path.ZstdError:
        return None


def run_fuzz(lib, seed, per_frame):
    rng = np.random.default_rng(seed)
    datas = [text(6000, 51), text(300, 52) + bytes(500) + rand(700, 53), text(40_000, 54), text(2 * BLOCK + 5000, 55),
             rand(BLOCK + 100, 56) + text(BLOCK, 57)]
    bases = []
    for level in (1, 3):
        bases += [(ref_path.ref_compress(d, level=level), d) for d in datas]
    bases += list(zip(_own_frames(lib, datas), datas))
    bases.append(_raw_frame([rand(BLOCK, 6), rand(BLOCK, 7), rand(5, 8)]))
    frames, ulens, origs = [], [], []
    for fr, d in bases:
        for m in _mutations(fr, rng, per_frame):
            frames.append(m)
            ulens.append(len(d))
            origs.append(d)
        frames.append(fr)  # and the intact frame in the same batch
        ulens.append(len(d))
        origs.append(d)
    outs, ok, status, rc = unpack_batch(lib, frames, ulens, [_b3(d) for d in origs])
    accepted = rejected_both = 0
    for f, u, d, o, st, okk in zip(frames, ulens, origs, outs, status, ok):
        ref = _reference_verdict(f, u)
        if st == 0:
            # accepted: must be exactly what the reference decoder restores (a mutation can leave a frame valid,
            # e.g. in a stored block, when the frame carries no checksum -- these all do, so it is rare)
            assert ref is not None and o == ref, "accepted a frame the reference decoder rejects or decodes differently"
            assert okk == (1 if o == d else 0)
            accepted += 1
        else:
            assert okk == 0
            if ref is None:
                rejected_both += 1
    n_intact = len(bases)
    assert accepted >= n_intact           # every intact frame decoded
    assert rejected_both >= (len(frames) - accepted) * 0.9  # and what we reject, the reference nearly always rejects too
    return accepted, rejected_both, len(frames)


def test_mutated_frames(emu):
    run_fuzz(emu, seed=1234, per_frame=12)
