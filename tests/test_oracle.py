"""Pins the oracle (oracle/zarc_oracle.c + oracle/ref_path.py).

The reference has no tests or golden vectors for this path, so the C restatement
is pinned against the same-version third-party libraries the reference links -- libzstd 1.5.5
(dlopen), Python blake3, xxhash -- and against the committed fixtures generated from them.
"""
import base64
import json
import os

import numpy as np
import pytest

from oracle import ref_path
from tests.golden.recipes import RECIPES, make_input, text, rand

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat.json")))


def test_libzstd_is_the_reference_version():
    assert ref_path.ZSTD_VERSION == 10505  # zstd-sys 2.0.9+zstd.1.5.5 (Cargo.lock:2480-2481)


def test_survey_known_answers():
    # SURVEY.md App. B / App. E
    assert ref_path.c_blake3(b"").hex() == "af1349b9f5f9a1a6a0404dea36dcc9499bcb25c9adc112b7cc9a93cae41f3262"
    assert ref_path.ref_compress(b"").hex() == "28b52ffd2400010000" + "99e9d851"
    assert ref_path.ref_compress(b"a").hex() == "28b52ffd2401090000" + "61" + "5b6e8ca9"
    assert ref_path.ref_compress(b"hello world\n").hex() == "28b52ffd240c610000" + b"hello world\n".hex() + "8c6d7d20"
    assert ref_path.ref_compress(bytes(1000)).hex() == "28b52ffd64e8024d000010000001" + "00e32b80055a074479"
    f = ref_path.ref_compress(b"hello zarc " * 1000)
    assert len(f) == 33 and f.hex().startswith("28b52ffd64f8299d000058") and f.hex().endswith("d269f6f1")


@pytest.mark.parametrize("case", KAT["cases"], ids=[c["recipe"] for c in KAT["cases"]])
def test_golden_vectors(case):
    import blake3
    import xxhash

    data = make_input(case["recipe"])
    assert len(data) == case["len"]
    # the fixture agrees with the live libraries ...
    assert blake3.blake3(data).hexdigest() == case["blake3"]
    assert f"{xxhash.xxh64(data, seed=0).intdigest():016x}" == case["xxh64"]
    # ... and the C restatement agrees with the fixture
    assert ref_path.c_blake3(data).hex() == case["blake3"]
    assert f"{ref_path.c_xxh64(data):016x}" == case["xxh64"]
    for level, b64 in case["frames"].items():
        frame = base64.b64decode(b64)
        assert ref_path.ref_compress(data, level=int(level)) == frame  # libzstd output is deterministic
        got, err, consumed, _ = ref_path.c_zstd_decompress_frame(frame, len(data))
        assert err == 0 and got == data and consumed == len(frame)


def test_c_decoder_vs_libzstd_feature_coverage():
    """The C decoder restores libzstd frames of levels 1/3/9 and the inputs exercise every section type
    the survey saw in reference-made frames (App. E)."""
    rng = np.random.default_rng(5)
    inputs = [make_input(n) for n in RECIPES]
    inputs += [text(400_000, 21), text(120_000, 22) + rand(140_000, 23) + text(200_000, 22)]
    inputs += [rng.integers(0, 4, 50_000, dtype=np.uint8).tobytes()]
    # fixed-shape records: every sequence has the same LL/ML/OF code -> RLE sequence tables
    inputs += [b"".join(b"RECORD_" + bytes([v]) for v in range(255))]
    seen = {}
    for data in inputs:
        for level in (1, 3, 9, 19):
            frame = ref_path.ref_compress(data, level=level)
            got, err, consumed, st = ref_path.c_zstd_decompress_frame(frame, len(data))
            assert err == 0, (len(data), level, err)
            assert got == data and consumed == len(frame)
            for k, v in st.items():
                seen[k] = seen.get(k, 0) + v
    for k in ("blocks_raw", "blocks_compressed", "lit_raw", "lit_huf1", "lit_huf4", "lit_treeless", "seq_predefined",
              "seq_fse", "seq_repeat", "seq_rle", "rep_offsets"):
        assert seen[k] > 0, k


def test_c_decoder_rejects_corruption():
    data = text(5000, 3)
    frame = bytearray(ref_path.ref_compress(data))
    frame[-1] ^= 1  # checksum
    _, err, _, _ = ref_path.c_zstd_decompress_frame(bytes(frame), len(data))
    assert err == 22
    _, err, _, _ = ref_path.c_zstd_decompress_frame(b"\x00" * 16, 16)
    assert err == 10
    good = ref_path.ref_compress(data)
    _, err, _, _ = ref_path.c_zstd_decompress_frame(good, len(data) - 1)
    assert err == 70
    _, err, _, _ = ref_path.c_zstd_decompress_frame(good[:-5], len(data))
    assert err != 0


def test_reference_call_sequence_dedup_and_offsets():
    """RefEncoder restates add_data_frame's bookkeeping (content_frame.rs:20-60)."""
    out = bytearray()
    enc = ref_path.RefEncoder(out, level=3)
    a, b = text(3000, 1), rand(2000, 2)
    da, db, da2 = enc.add_data_frame(a), enc.add_data_frame(b), enc.add_data_frame(a)
    assert da == da2 and da != db and len(enc.frames) == 2
    fa, fb = enc.frames[da], enc.frames[db]
    assert fa.offset == 12 and fb.offset == 12 + fa.length and len(out) == fb.offset + fb.length
    dec = ref_path.RefDecoder(bytes(out), enc.frames)
    assert dec.read_content_frame(da) == (a, True)
    assert dec.read_content_frame(db) == (b, True)
    assert dec.read_content_frame(b"\x00" * 32) is None
