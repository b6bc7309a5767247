"""The randomised checks of tests/fuzz_cases.py on the real GPU through the C ABI (a few seeds each; the same code runs
for thousands of seeds on the CPU emulator build: tools/fuzz_*.py)."""
import pytest

from tests import fuzz_cases as fz

pytestmark = pytest.mark.gpu


def test_encoder_roundtrips(gpu):
    assert fz.encode_roundtrips(gpu, 11, 2, levels=(1, 3, 9)) == 2 * 6 * 3


def test_reference_made_frames(gpu):
    assert fz.decode_reference_frames(gpu, 5, 2, levels=(1, 3, 19)) == 2 * 5 * 3


def test_pack_bookkeeping(gpu):
    fz.pack_bookkeeping(gpu, 3, 4)


def test_streaming_api(gpu):
    agree, rejected = fz.streaming(gpu, 21, 4)
    assert agree >= 12 and rejected > 0
