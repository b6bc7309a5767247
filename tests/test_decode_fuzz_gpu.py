"""The mutated-frame cases of test_decode_fuzz_emu.py on the real GPU (more mutations per frame)."""
import pytest

from tests.test_decode_fuzz_emu import run_fuzz

pytestmark = pytest.mark.gpu


def test_mutated_frames(gpu):
    run_fuzz(gpu, seed=99, per_frame=60)
