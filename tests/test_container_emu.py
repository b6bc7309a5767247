"""f1-f3: container writer/reader, CLI drivers and the directory frame through the kernels -- the C++
host driven end to end on the CPU, with the kernels running on the SIMT emulator (tests/simt_emu)."""
import pytest

from tests import container_cases as cc
from zarc_b200 import build


@pytest.fixture(scope="module")
def emu_path():
    return build.build_emu()


def test_pack_matches_reference_container(emu_path, tmp_path):
    cc.check_pack_against_oracle(emu_path, tmp_path)


def test_unpack_reference_archives(emu_path, tmp_path):
    cc.check_unpack_of_reference_archive(emu_path, tmp_path, levels=(3,))


def test_roundtrip_multi_batch_and_errors(emu_path, tmp_path):
    # ZARC_BATCH_MB=0 forces one GPU pass per file (every batch boundary case), dedup across passes included
    cc.check_roundtrip_and_errors(emu_path, tmp_path, env_extra={"ZARC_BATCH_MB": "0"})
