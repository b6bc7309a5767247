"""Kernel logic of K1 (BLAKE3) and K4 (XXH64) on the SIMT emulator vs the oracle (CPU, no GPU)."""
import numpy as np
import pytest

from oracle import ref_path
from tests.helpers import blake3_batch, xxh64_batch

SIZES = [0, 1, 3, 63, 64, 65, 127, 128, 1023, 1024, 1025, 2047, 2048, 2049, 3 * 1024, 5000, 8191, 8192, 8193, 9247, 31 * 1024, 32 * 1024,
         32 * 1024 + 1, 33 * 1024, 64 * 1024, 65 * 1024 + 7, 100_000, 128 * 1024 - 1, 128 * 1024, 128 * 1024 + 5, 200_000, 300 * 1024 + 33]


def _data(n, seed=0):
    return np.random.default_rng(seed + n).integers(0, 256, n, dtype=np.uint8).tobytes()


@pytest.mark.parametrize("shift", [0, 1, 4, 8])
def test_blake3_sizes(emu, shift):
    import blake3

    files = [_data(n) for n in SIZES]
    got = blake3_batch(emu, files, shift=shift)
    for f, g in zip(files, got):
        assert g == blake3.blake3(f).digest(), len(f)


@pytest.mark.parametrize("variant", range(13))
def test_blake3_chunk_kernel_variants(emu, variant):
    """Every staging / arithmetic variant of the chunk kernel (bulk copies or cp.async, 256- or 512-byte pieces, the
    additions and rotations of G split over the pipes) gives the same digests."""
    import blake3

    files = [_data(n, 7) for n in (0, 1, 64, 1000, 1024, 1025, 4096, 5000, 33 * 1024 + 3, 70_001)]
    emu.dll.zg_internal_set_b3_variant(variant)
    try:
        for shift, align in ((0, 1), (3, 1), (0, 16)):
            got = blake3_batch(emu, files, align=align, shift=shift)
            for f, g in zip(files, got):
                assert g == blake3.blake3(f).digest(), (variant, shift, len(f))
    finally:
        emu.dll.zg_internal_set_b3_variant(9)


def test_blake3_empty_kat(emu):
    assert blake3_batch(emu, [b""])[0].hex() == "af1349b9f5f9a1a6a0404dea36dcc9499bcb25c9adc112b7cc9a93cae41f3262"


def test_blake3_big_file_path(emu):
    import blake3

    # a few thousand chunks per file: a dozen tree levels; mixed with small files
    files = [_data(1024 * 1024 + 1), _data(10), _data(1024 * 1024 + 1024 * 33 + 5, 1), _data(2 * 1024 * 1024, 2), b""]
    got = blake3_batch(emu, files, align=16)
    for f, g in zip(files, got):
        assert g == blake3.blake3(f).digest(), len(f)


def test_xxh64_large_inputs_whole_warp_path(emu):
    """Large inputs through the one-warp-per-input path (1 KiB chunks through shared memory)."""
    import xxhash

    files = [_data((1 << 20) + 13), _data(1 << 20, 1), _data((1 << 20) - 1, 2), _data(3 * (1 << 20) + 1024 + 31, 3), _data(100, 4)]
    for shift in (0, 3):
        got = xxh64_batch(emu, files, shift=shift)
        for f, g in zip(files, got):
            assert g == xxhash.xxh64(f, seed=0).intdigest(), len(f)


@pytest.mark.parametrize("shift", [0, 1, 8])
def test_xxh64_sizes(emu, shift):
    import xxhash

    files = [_data(n) for n in SIZES]
    got = xxh64_batch(emu, files, shift=shift)
    for f, g in zip(files, got):
        assert g == xxhash.xxh64(f, seed=0).intdigest(), len(f)
