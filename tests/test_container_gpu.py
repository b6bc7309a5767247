"""f1-f3 on the GPU: the C++ host + product library against the oracle's reference container."""
import pytest

from tests import container_cases as cc
from zarc_b200 import _lib

pytestmark = pytest.mark.gpu


def test_pack_matches_reference_container_gpu(gpu, tmp_path):
    cc.check_pack_against_oracle(_lib.PRODUCT_SO, tmp_path)


def test_unpack_reference_archives_gpu(gpu, tmp_path):
    cc.check_unpack_of_reference_archive(_lib.PRODUCT_SO, tmp_path, levels=(1, 3, 9))


def test_roundtrip_multi_batch_and_errors_gpu(gpu, tmp_path):
    cc.check_roundtrip_and_errors(_lib.PRODUCT_SO, tmp_path, env_extra={"ZARC_BATCH_MB": "1"})
