"""Hostile archives and arguments at the boundary (round-1 advisor findings): nothing may crash, throw across the
C ABI, read out of bounds or write outside the extraction directory.  Kernels run on the SIMT emulator."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from oracle import ref_container, ref_path
from tests import container_cases as cc
from tests.golden.recipes import rand, text
from tests.helpers import pack_batch
from zarc_b200 import build
from zarc_b200._lib import InBuffer, OutBuffer


@pytest.fixture(scope="module")
def emu_path():
    return build.build_emu()


class _Writer(ref_container.RefArchiveWriter):
    """RefArchiveWriter whose directory elements can be tampered with before finalise()."""

    def __init__(self, frame_digest=None, file_digest=None):
        super().__init__(level=3)
        self.frame_digest, self.file_digest = frame_digest, file_digest

    def finalise(self) -> bytes:
        directory = bytearray(ref_container.edition_element())
        for f in self.files:
            if f["digest"] is not None:
                fr = self.enc.frames[f["digest"]]
                fd = self.frame_digest(fr.digest) if self.frame_digest else fr.digest
                directory += ref_container.element(3, {0: fr.edition, 1: fr.offset, 2: fd, 3: fr.length, 4: fr.uncompressed})
                f = dict(f, digest=self.file_digest(f["digest"]) if self.file_digest else f["digest"])
            directory += ref_container.file_element(**f)
        digest = ref_path._blake3(bytes(directory))
        comp = self.enc.compress_frame(bytes(directory))
        self.out += comp
        tb = ref_container.trailer_bytes(digest, -(len(comp) + 8 + 32 + ref_container.EPILOGUE_LENGTH), len(directory))
        self.out += bytes([0x5F, 0x2A, 0x4D, 0x18]) + struct.pack("<I", len(tb)) + tb
        return bytes(self.out)


@pytest.mark.parametrize("bad", [b"", b"\x01\x02\x03", b"\x00" * 31, b"\x00" * 33])
def test_directory_digest_of_wrong_length_is_a_parse_error(emu_path, tmp_path, bad):
    w = _Writer(frame_digest=lambda d: bad, file_digest=lambda d: bad)
    w.add_file(["a.txt"], text(3000, 1))
    (tmp_path / "bad.zarc").write_bytes(w.finalise())
    for cmd in (["unpack", "bad.zarc"], ["list-files", "bad.zarc"]):
        p = cc.run_cli(emu_path, cmd, cwd=str(tmp_path), check=False)
        assert p.returncode == 1, (p.returncode, p.stderr)  # an error exit, not a signal
        assert "digest has the wrong length" in p.stderr


def test_unpack_refuses_paths_that_leave_the_extraction_directory(emu_path, tmp_path):
    outside = tmp_path / "outside"
    outside.mkdir()
    work = tmp_path / "work"
    work.mkdir()
    w = ref_container.RefArchiveWriter(level=3)
    w.add_file(["..", "escaped.txt"], b"escaped")
    w.add_file([str(outside / "abs_escaped.txt")], b"absolute")
    w.add_file(["ok", ".", "fine.txt"], b"fine", mode=0o106755)  # setuid + setgid requested
    w.add_file(["ok", "sub/dir.txt"], b"slash inside a component")
    (work / "evil.zarc").write_bytes(w.finalise())
    p = cc.run_cli(emu_path, ["unpack", "evil.zarc"], cwd=str(work))
    assert "unpacked 1 files" in p.stderr and "skipped 3 entries" in p.stderr
    assert (work / "ok" / "fine.txt").read_bytes() == b"fine"
    assert not (tmp_path / "escaped.txt").exists() and not (outside / "abs_escaped.txt").exists()
    assert not (work / "ok" / "sub").exists()
    mode = (work / "ok" / "fine.txt").stat().st_mode
    assert mode & 0o6000 == 0 and mode & 0o777 == 0o755  # setuid / setgid of an untrusted archive are not restored
    # list-files still shows what the archive claims
    ls = cc.run_cli(emu_path, ["list-files", "evil.zarc"], cwd=str(work)).stdout
    assert "../escaped.txt" in ls


def test_stream_decoder_rejects_a_lying_content_size_without_throwing(emu):
    # magic, descriptor 0xE0 (single segment, 8-byte FCS), FCS = 2^62, one empty last raw block
    frame = bytes([0x28, 0xB5, 0x2F, 0xFD, 0xE0]) + struct.pack("<Q", 1 << 62) + bytes([0x01, 0x00, 0x00])
    d = emu.zg_dctx_create()
    out = C.create_string_buffer(1 << 17)
    ob = OutBuffer(C.cast(out, C.c_void_p), len(out), 0)
    src = C.create_string_buffer(frame, len(frame))
    ib = InBuffer(C.cast(src, C.c_void_p), len(frame), 0)
    r = emu.zg_decompress_stream(d, C.byref(ob), C.byref(ib))
    assert emu.zg_is_error(r) and emu.zg_get_error_code(r) == 20  # corruption_detected, like any frame whose FCS lies
    # the same lie, smaller: 1 GiB claimed by a frame of one 10-byte raw block
    frame = bytes([0x28, 0xB5, 0x2F, 0xFD, 0xA0]) + struct.pack("<I", 1 << 30) + bytes([0x51, 0x00, 0x00]) + b"0123456789"
    src = C.create_string_buffer(frame, len(frame))
    ib = InBuffer(C.cast(src, C.c_void_p), len(frame), 0)
    r = emu.zg_decompress_stream(d, C.byref(ob), C.byref(ib))
    assert emu.zg_is_error(r)
    # the context is still usable
    good = ref_path.ref_compress(b"hello " * 100)
    src = C.create_string_buffer(good, len(good))
    ib = InBuffer(C.cast(src, C.c_void_p), len(good), 0)
    ob.pos = 0
    r = emu.zg_decompress_stream(d, C.byref(ob), C.byref(ib))
    assert r == 0 and out.raw[: ob.pos] == b"hello " * 100
    emu.zg_dctx_free(d)


def test_rejected_frames_are_not_hashed_out_of_bounds(emu):
    data = text(3000, 5)
    frame = ref_path.ref_compress(data)
    arch = np.frombuffer(frame + frame, dtype=np.uint8).copy()
    off = np.array([0, len(frame)], dtype=np.uint64)
    ln = np.array([len(frame)] * 2, dtype=np.uint64)
    ul = np.array([len(data), 1 << 40], dtype=np.uint64)  # the second frame claims a terabyte
    oo = np.array([0, 4096 - 8], dtype=np.uint64)
    out = np.zeros(4096, dtype=np.uint8)
    dig = np.frombuffer(ref_path._blake3(data) * 2, dtype=np.uint8).copy()
    ok = np.zeros(2, dtype=np.uint8)
    st = np.zeros(2, dtype=np.uint32)
    d = emu.zg_dctx_create()
    r = emu.zg_unpack_batch_dev(d, arch.ctypes.data, len(arch), 2, off.ctypes.data, ln.ctypes.data, ul.ctypes.data, dig.ctypes.data,
                                out.ctypes.data, 4096, oo.ctypes.data, ok.ctypes.data, st.ctypes.data)
    assert emu.zg_is_error(r) and emu.zg_get_error_code(r) == 70  # dstSize_tooSmall, of the lowest failing frame
    assert list(st) == [0, 70] and list(ok) == [1, 0]
    assert bytes(out[: len(data)]) == data  # the good frame is still delivered
    # host variant: sizes whose sum wraps 2^64 are refused, not wrapped
    ul2 = np.array([(1 << 64) - 16, 32], dtype=np.uint64)
    r = emu.zg_unpack_batch(d, arch.ctypes.data, len(arch), 2, off.ctypes.data, ln.ctypes.data, ul2.ctypes.data, None, out.ctypes.data, 4096,
                            None, None, st.ctypes.data)
    assert emu.zg_is_error(r) and emu.zg_get_error_code(r) == 70
    emu.zg_dctx_free(d)


def test_failed_pack_batch_leaves_the_archive_state_untouched(emu):
    files = [text(4000, 1), rand(3000, 2), text(4000, 1), text(9000, 3)]
    c = emu.zg_cctx_create()
    emu.check(emu.zg_cctx_init(c, 0))
    emu.check(emu.zg_cctx_set_parameter(c, 201, 1))
    emu.check(emu.zg_cctx_reset_archive(c, 12))
    first = pack_batch(emu, c, files[:2])
    assert first["rc"] == 0 and first["first"] == [1, 1]
    off0 = emu.zg_cctx_archive_offset(c)
    # a batch that cannot fit: error, nothing committed
    bad = pack_batch(emu, c, files[2:] + [rand(5000, 9)], cap=64)
    assert emu.zg_is_error(bad["rc"]) and emu.zg_get_error_code(bad["rc"]) == 70 and bad["frames"] == b""
    assert emu.zg_cctx_archive_offset(c) == off0
    # the retry sees the state of before the failed call: same decisions and offsets as a context that never failed
    again = pack_batch(emu, c, files[2:] + [rand(5000, 9)])
    c2 = emu.zg_cctx_create()
    emu.check(emu.zg_cctx_init(c2, 0))
    emu.check(emu.zg_cctx_set_parameter(c2, 201, 1))
    emu.check(emu.zg_cctx_reset_archive(c2, 12))
    pack_batch(emu, c2, files[:2])
    clean = pack_batch(emu, c2, files[2:] + [rand(5000, 9)])
    assert again["rc"] == 0 and again["first"] == clean["first"] == [0, 1, 1]
    assert again["off"] == clean["off"] and again["len"] == clean["len"] and again["frames"] == clean["frames"]
    emu.zg_cctx_free(c)
    emu.zg_cctx_free(c2)
