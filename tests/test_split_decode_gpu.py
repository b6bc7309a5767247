"""The block-parallel decoding cases of test_split_decode_emu.py on the real GPU, plus a C3-shaped one."""
import pytest

from tests import test_split_decode_emu as cases

pytestmark = pytest.mark.gpu


def test_own_multi_block_frames_decode_block_parallel(gpu):
    cases.test_own_multi_block_frames_decode_block_parallel(gpu)


def test_reference_multi_block_frames_decode_staged_in_one_pass(gpu):
    cases.test_reference_multi_block_frames_decode_staged_in_one_pass(gpu)


def test_chunked_staging_equals_one_pass(gpu):
    cases.test_chunked_staging_equals_one_pass(gpu)


def test_split_threshold_and_mixed_batch(gpu):
    cases.test_split_threshold_and_mixed_batch(gpu)


def test_short_blocks_are_not_mistaken_for_full_ones(gpu):
    cases.test_short_blocks_are_not_mistaken_for_full_ones(gpu)


def test_corruption_inside_split_frames(gpu):
    cases.test_corruption_inside_split_frames(gpu)


def test_chain_executor_window_edges_and_both_modes(gpu):
    cases.test_chain_executor_window_edges_and_both_modes(gpu)


def test_checksums_hashed_chunk_by_chunk(gpu):
    cases.test_checksums_hashed_chunk_by_chunk(gpu)
