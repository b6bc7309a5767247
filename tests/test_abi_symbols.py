"""The C-ABI library loads and exports every symbol include/zarcgpu.h declares (no compute calls:
this runs without a GPU), and the product path fails loudly -- never falls back -- without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "zarcgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zg_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def product():
    from zarc_b200 import build, _lib

    build.build_product()
    return _lib.Lib(_lib.PRODUCT_SO)  # strict: raises on any missing symbol of the binding table


def test_every_declared_symbol_is_exported(product):
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(product.dll, n), f"{n} is declared in include/zarcgpu.h but not exported"


def test_binding_table_covers_the_header():
    from zarc_b200 import _lib

    assert set(_declared()) <= set(_lib.SIGNATURES), sorted(set(_declared()) - set(_lib.SIGNATURES))


def test_emulator_build_exports_the_same_abi(emu):
    for n in _declared():
        assert hasattr(emu.dll, n), n


def test_product_is_sm100a_and_has_no_cpu_fallback(product):
    assert product.zg_build_info() == b"sm_100a"
    assert product.zg_dstream_in_size() == 131075 and product.zg_dstream_out_size() == 131072
    assert product.zg_error_name((1 << 64) - 20) == b"Data corruption detected"
    assert product.zg_is_error((1 << 64) - 70) and not product.zg_is_error(12345)
    if product.zg_device_count() == 0:
        out = np.zeros(32, dtype=np.uint8)
        rc = product.zg_blake3(b"abc", 3, out.ctypes.data)
        assert product.zg_get_error_code(rc) == 111  # ZG_error_no_device: fails loudly
        assert not product.zg_cctx_create() and not product.zg_dctx_create()


def test_missing_library_raises():
    from zarc_b200 import _lib

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.Lib("/nonexistent/libzarcgpu.so")


def test_product_package_never_references_the_oracle_or_emulator():
    """The shipped path must not import/call oracle/ or the SIMT emulator."""
    pkg = os.path.join(ROOT, "zarc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) and f != "build.py":
                txt = open(os.path.join(dirpath, f)).read()
                for needle in ("import oracle", "from oracle", "oracle/", "oracle.", "libzarc_oracle", "ref_path", "libzstd.so", "CDLL(\"libzstd"):
                    assert needle not in txt, (f, needle)
                if f != "simt.h":
                    assert "simt_emu" not in txt, f
