"""Encoder + pack parity on the real GPU through the C ABI (same cases as the emulator suite, plus
corpus-shaped batches): frames must be restored byte-identically by libzstd 1.5.5 and by our decoder;
digests / dedup / offsets must equal the reference path's."""
import numpy as np
import pytest

from oracle import ref_path
from tests import test_encode_emu as cases
from tests.golden.recipes import RECIPES
from tests.helpers import pack_batch, unpack_batch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(RECIPES))
def test_compress2_recipes(gpu, name):
    cases.test_compress2_recipes(gpu, name)


def test_compress2_known_answer_shapes(gpu):
    cases.test_compress2_known_answer_shapes(gpu)


def test_compress2_flags_and_errors(gpu):
    cases.test_compress2_flags_and_errors(gpu)


@pytest.mark.parametrize("level", [1, 3])
def test_few_sequences_and_chain_ranges(gpu, level):
    cases.test_few_sequences_and_chain_ranges(gpu, level)


def test_ratio_vs_reference_small_corpus(gpu):
    cases.test_ratio_vs_reference_small_corpus(gpu)


def test_pack_batch_matches_reference_bookkeeping(gpu):
    cases.test_pack_batch_matches_reference_bookkeeping(gpu)


def test_pack_batch_capacity_error(gpu):
    cases.test_pack_batch_capacity_error(gpu)


def test_sliced_host_api_and_chunked_encoder_equal_one_shot(gpu):
    cases.test_sliced_host_api_and_chunked_encoder_equal_one_shot(gpu)


@pytest.mark.parametrize("workers", [1, 2, 4])
def test_sliced_unpack_workers_and_frame_errors(gpu, workers):
    cases.test_sliced_unpack_workers_and_frame_errors(gpu, workers)


def _roundtrip_corpus(gpu, c, level, sample_every):
    import blake3

    from zarc_b200 import corpus

    blob = corpus.materialise_host(gpu, c)
    datas = [bytes(blob[int(o) : int(o) + int(l)]) for o, l in zip(c.off, c.len)]
    cctx = gpu.zg_cctx_create()
    gpu.check(gpu.zg_cctx_init(cctx, 0))
    gpu.check(gpu.zg_cctx_set_parameter(cctx, 201, 1))
    gpu.check(gpu.zg_cctx_set_parameter(cctx, 100, level))
    gpu.check(gpu.zg_cctx_reset_archive(cctx, 12))
    r = pack_batch(gpu, cctx, datas, align=16)
    gpu.zg_cctx_free(cctx)
    assert r["rc"] == 0
    digests = [blake3.blake3(d).digest() for d in datas]
    assert r["digests"] == digests
    seen, first = {}, []
    for i, d in enumerate(digests):
        first.append(0 if d in seen else 1)
        seen.setdefault(d, i)
    assert r["first"] == first
    frames = [r["frames"][o - 12 : o - 12 + l] for o, l in zip(r["off"], r["len"])]
    # the reference decoder on a sample, ours on everything (duplicates decode their shared frame)
    for i in range(0, len(datas), sample_every):
        assert ref_path.ref_decompress(frames[i], len(datas[i])) == datas[i]
    outs, ok, status, rc = unpack_batch(gpu, frames, [len(d) for d in datas], digests)
    assert rc == 0 and not any(status)
    assert outs == datas and all(ok)
    uniq_bytes = sum(len(d) for d, f in zip(datas, first) if f)
    return uniq_bytes / max(len(r["frames"]), 1)


@pytest.mark.parametrize("level", [1, 3])
def test_c2_shaped_roundtrip(gpu, level):
    from zarc_b200 import corpus

    ratio = _roundtrip_corpus(gpu, corpus.c2_source_tree(total_bytes=30_000_000, seed=21), level, 7)
    assert ratio > 2.0


def test_c4_dedup_heavy_roundtrip(gpu):
    from zarc_b200 import corpus

    _roundtrip_corpus(gpu, corpus.c2_source_tree(total_bytes=20_000_000, seed=22, dup=True), 3, 11)


def test_c1_shaped_roundtrip(gpu):
    from zarc_b200 import corpus

    ratio = _roundtrip_corpus(gpu, corpus.c1_tree(total_bytes=64 << 20, n_files=500, seed=23), 3, 5)
    assert ratio > 1.2


def test_c3_shaped_roundtrip(gpu):
    from zarc_b200 import corpus

    ratio = _roundtrip_corpus(gpu, corpus.c3_huge(n_files=3, file_bytes=(48 << 20) + 12345, seed=24), 3, 1)
    assert ratio > 2.5
