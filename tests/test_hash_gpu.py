"""K1 BLAKE3 / K4 XXH64 on the real GPU through the C ABI, vs the oracle."""
import numpy as np
import pytest

from oracle import ref_path
from tests.helpers import blake3_batch, xxh64_batch
from tests.test_hash_emu import SIZES, _data

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shift", [0, 1, 4, 8])
def test_blake3_sizes(gpu, shift):
    import blake3

    files = [_data(n) for n in SIZES]
    got = blake3_batch(gpu, files, shift=shift)
    for f, g in zip(files, got):
        assert g == blake3.blake3(f).digest() == ref_path.c_blake3(f), len(f)


def test_blake3_big_files(gpu):
    import blake3

    files = [_data(1024 * 1024 + 1), _data(10), _data(40 * 1024 * 1024 + 12345, 1), _data(2 * 1024 * 1024, 2), b""]
    got = blake3_batch(gpu, files, align=16)
    for f, g in zip(files, got):
        assert g == blake3.blake3(f).digest(), len(f)


def test_blake3_many_ragged(gpu):
    import blake3

    rng = np.random.default_rng(3)
    files = [_data(int(n), 5) for n in rng.integers(0, 70000, 3000)]
    got = blake3_batch(gpu, files)
    for f, g in zip(files, got):
        assert g == blake3.blake3(f).digest(), len(f)


@pytest.mark.parametrize("variant", range(13))
def test_blake3_chunk_kernel_variants(gpu, variant):
    from tests import test_hash_emu as cases

    cases.test_blake3_chunk_kernel_variants(gpu, variant)
    # and a ragged crowd of small files, where groups of 32 chunks straddle many files
    import blake3

    rng = np.random.default_rng(11 + variant)
    files = [_data(int(n), 6) for n in rng.integers(0, 40000, 600)]
    gpu.dll.zg_internal_set_b3_variant(variant)
    try:
        got = blake3_batch(gpu, files, shift=variant)
    finally:
        gpu.dll.zg_internal_set_b3_variant(9)
    for f, g in zip(files, got):
        assert g == blake3.blake3(f).digest(), len(f)


@pytest.mark.parametrize("shift", [0, 1, 8])
def test_xxh64_sizes(gpu, shift):
    import xxhash

    files = [_data(n) for n in SIZES] + [_data(3_000_000, 9)]
    got = xxh64_batch(gpu, files, shift=shift)
    for f, g in zip(files, got):
        assert g == xxhash.xxh64(f, seed=0).intdigest() == ref_path.c_xxh64(f), len(f)


def test_xxh64_large_inputs_whole_warp_path(gpu):
    from tests import test_hash_emu as cases

    cases.test_xxh64_large_inputs_whole_warp_path(gpu)
