"""The decoder parity cases of test_decode_emu.py, on the real GPU through the C ABI."""
import pytest

from tests import test_decode_emu as cases

pytestmark = pytest.mark.gpu


def test_golden_frames(gpu):
    cases.test_golden_frames(gpu)


@pytest.mark.parametrize("level", [1, 3, 9, 19])
def test_levels_and_features(gpu, level):
    cases.test_levels_and_features(gpu, level)


@pytest.mark.parametrize("level", [1, 3, 19])
def test_periodic_and_chained_matches(gpu, level):
    cases.test_periodic_and_chained_matches(gpu, level)


def test_no_checksum_frames_and_digest_mismatch(gpu):
    cases.test_no_checksum_frames_and_digest_mismatch(gpu)


def test_corruption_is_reported_per_frame(gpu):
    cases.test_corruption_is_reported_per_frame(gpu)


def test_one_shot_and_streaming_api(gpu):
    cases.test_one_shot_and_streaming_api(gpu)


def test_many_reference_frames(gpu):
    """A C2-shaped batch (2 000 files, 1..64 KiB, src/text/random) packed by the reference path at
    levels 1/3/9, unpacked + verified on the GPU, byte-identical."""
    import blake3
    import numpy as np

    from oracle import ref_path
    from tests.helpers import unpack_batch
    from zarc_b200 import corpus

    c = corpus.c2_source_tree(total_bytes=20_000_000, seed=11)
    blob = corpus.materialise_host(gpu, c)
    datas = [bytes(blob[int(o) : int(o) + int(l)]) for o, l in zip(c.off, c.len)]
    for level in (1, 3, 9):
        frames = [ref_path.ref_compress(d, level=level) for d in datas]
        outs, ok, status, rc = unpack_batch(gpu, frames, [len(d) for d in datas], [blake3.blake3(d).digest() for d in datas])
        assert rc == 0 and not any(status)
        assert outs == datas and all(ok)
