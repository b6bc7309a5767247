"""A few seeds of every randomised check of tests/fuzz_cases.py on the SIMT-emulator build (CPU).  The long runs are
tools/fuzz_*.py (thousands of seeds each; results in profiles/README.md)."""
import pytest

from tests import fuzz_cases as fz
from zarc_b200 import build


@pytest.mark.parametrize("first", [11, 4242])
def test_encoder_roundtrips(emu, first):
    assert fz.encode_roundtrips(emu, first, 2, levels=(1, 3, 9)) == 2 * 6 * 3


@pytest.mark.parametrize("first", [5, 777])
def test_reference_made_frames(emu, first):
    assert fz.decode_reference_frames(emu, first, 2, levels=(1, 3, 19)) == 2 * 5 * 3


@pytest.mark.parametrize("first", [3, 900])
def test_pack_bookkeeping(emu, first):
    fz.pack_bookkeeping(emu, first, 4)


def test_streaming_api(emu):
    agree, rejected = fz.streaming(emu, 21, 6)
    assert agree >= 18 and rejected > 0


def test_container_directories():
    assert fz.container_directories(build.build_emu(), 300, 60) >= 1


def test_hostile_metadata(emu):
    fz.hostile_metadata(emu, 40, 60)


def test_compress_capacity(emu):
    made, refused = fz.compress_capacity(emu, 7, 12)
    assert made > 20 and refused > 20
