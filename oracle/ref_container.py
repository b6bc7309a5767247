"""ORACLE (test infrastructure -- never imported by the product path): the zarc container around the
content frames, restated from the reference.  The reference itself (Rust) cannot be built here, so
this follows its source line by line:

  header      crates/zarc/src/header.rs:35-40
  elements    crates/zarc/src/directory/elements.rs:10-25,58-74  (kind u8 | len u16 LE | 00 | CBOR)
  Edition     directory/edition.rs:13-34      File  directory/file.rs:18-62      Frame  directory/frame.rs:12-32
  finalise    encode/directory.rs:40-122      trailer  trailer.rs:66-108,139-173
  open/read   decode/open.rs:71-158, decode/directory.rs:55-119

`RefArchiveWriter` produces archives the way the reference's `Encoder` would (libzstd 1.5.5 frames via
oracle/ref_path.py, python `blake3`); `read_archive` parses any archive and restores content with
libzstd.  Parity is unpinned by reference-owned vectors (the reference has none).
"""
from __future__ import annotations

import datetime
import struct

from . import ref_path

ZARC_MAGIC = bytes([0x65, 0xAA, 0xDC])
FILE_MAGIC = ref_path.RefEncoder.FILE_MAGIC
EPILOGUE_LENGTH = 22


# ---- CBOR (RFC 8949), the subset the directory uses ------------------------------------------------
class Tag:
    def __init__(self, tag, value):
        self.tag, self.value = tag, value

    def __eq__(self, o):
        return isinstance(o, Tag) and (self.tag, self.value) == (o.tag, o.value)

    def __repr__(self):
        return f"Tag({self.tag}, {self.value!r})"


def _head(major, v):
    if v < 24:
        return bytes([major << 5 | v])
    for ai, fmt in ((24, ">B"), (25, ">H"), (26, ">I"), (27, ">Q")):
        if v < 1 << (8 * struct.calcsize(fmt)):
            return bytes([major << 5 | ai]) + struct.pack(fmt, v)
    raise ValueError(v)


def cbor_encode(x) -> bytes:
    if x is None:
        return b"\xf6"
    if x is True:
        return b"\xf5"
    if x is False:
        return b"\xf4"
    if isinstance(x, int):
        return _head(0, x) if x >= 0 else _head(1, -1 - x)
    if isinstance(x, (bytes, bytearray)):
        return _head(2, len(x)) + bytes(x)
    if isinstance(x, str):
        b = x.encode()
        return _head(3, len(b)) + b
    if isinstance(x, (list, tuple)):
        return _head(4, len(x)) + b"".join(cbor_encode(i) for i in x)
    if isinstance(x, dict):  # insertion order (minicbor writes fields in index order)
        return _head(5, len(x)) + b"".join(cbor_encode(k) + cbor_encode(v) for k, v in x.items())
    if isinstance(x, Tag):
        return _head(6, x.tag) + cbor_encode(x.value)
    raise TypeError(type(x))


def cbor_decode(b: bytes, pos: int = 0):
    """Returns (value, next position)."""
    ib = b[pos]
    major, ai = ib >> 5, ib & 31
    pos += 1
    if major == 7:
        if ai in (20, 21):
            return ai == 21, pos
        if ai in (22, 23):
            return None, pos
        if ai == 26:
            return struct.unpack(">f", b[pos : pos + 4])[0], pos + 4
        if ai == 27:
            return struct.unpack(">d", b[pos : pos + 8])[0], pos + 8
        raise ValueError("unsupported simple value")
    indef = ai == 31
    if ai < 24:
        v = ai
    elif not indef:
        n = 1 << (ai - 24)
        v = int.from_bytes(b[pos : pos + n], "big")
        pos += n
    if major == 0:
        return v, pos
    if major == 1:
        return -1 - v, pos
    if major in (2, 3):
        if indef:
            parts = []
            while b[pos] != 0xFF:
                s, pos = cbor_decode(b, pos)
                parts.append(s if major == 2 else s.encode())
            data, pos = b"".join(parts), pos + 1
        else:
            data, pos = bytes(b[pos : pos + v]), pos + v
        return (data if major == 2 else data.decode()), pos
    if major == 4:
        out = []
        while (b[pos] != 0xFF) if indef else (len(out) < v):
            i, pos = cbor_decode(b, pos)
            out.append(i)
        return out, pos + (1 if indef else 0)
    if major == 5:
        out = {}
        while (b[pos] != 0xFF) if indef else (len(out) < v):
            k, pos = cbor_decode(b, pos)
            out[k], pos = cbor_decode(b, pos)
        return out, pos + (1 if indef else 0)
    if major == 6:
        val, pos = cbor_decode(b, pos)
        return Tag(v, val), pos
    raise ValueError(major)


# ---- directory elements -----------------------------------------------------------------------------
def rfc3339(ts: float | datetime.datetime) -> str:
    """chrono::DateTime<Utc>::to_rfc3339(): AutoSi fraction, '+00:00' (timestamps.rs:70-78)."""
    if not isinstance(ts, datetime.datetime):
        ts = datetime.datetime.fromtimestamp(ts, datetime.timezone.utc)
    s = ts.strftime("%Y-%m-%dT%H:%M:%S")
    us = ts.microsecond
    if us:
        s += f".{us // 1000:03d}" if us % 1000 == 0 else f".{us:06d}"
    return s + "+00:00"


def element(kind: int, payload_obj) -> bytes:
    payload = cbor_encode(payload_obj)
    assert len(payload) <= 0xFFFF
    return bytes([kind]) + struct.pack("<H", len(payload)) + b"\x00" + payload


def edition_element(number=1, written_at=None) -> bytes:
    return element(1, {0: number, 1: Tag(0, rfc3339(written_at or datetime.datetime.now(datetime.timezone.utc))), 2: 1})


def file_element(name: list, digest: bytes | None = None, mode=None, user=None, group=None, timestamps=None, special=None, edition=1) -> bytes:
    m = {0: edition, 1: list(name)}
    if digest is not None:
        m[2] = digest
    if mode is not None:
        m[3] = mode
    if user is not None:
        m[4] = list(user)
    if group is not None:
        m[5] = list(group)
    if timestamps is not None:
        m[6] = {k: Tag(0, rfc3339(v)) for k, v in timestamps.items()}
    if special is not None:
        m[7] = list(special)
    return element(2, m)


def frame_element(fr: ref_path.RefFrame) -> bytes:
    return element(3, {0: fr.edition, 1: fr.offset, 2: fr.digest, 3: fr.length, 4: fr.uncompressed})


def trailer_bytes(digest: bytes, directory_offset: int, directory_uncompressed: int) -> bytes:
    """Trailer::to_bytes (no prologue) with the check byte that XORs the prologue in (trailer.rs:66-108)."""
    def epilogue(check):
        return bytes([1]) + struct.pack("<q", directory_offset) + struct.pack("<Q", directory_uncompressed) + bytes([check, 1]) + ZARC_MAGIC

    check = 0
    for x in bytes([0, 1]) + digest + epilogue(0):
        check ^= x
    return digest + epilogue(check)


class RefArchiveWriter:
    """`Encoder` incl. add_file_entry + finalise, restated.  files: list of dicts given to file_element."""

    def __init__(self, level: int | None = None):
        self.out = bytearray()
        self.enc = ref_path.RefEncoder(self.out, checksum=True, level=level)
        self.files: list[dict] = []

    def add_file(self, name: list, content: bytes | None, **meta):
        digest = self.enc.add_data_frame(content) if content is not None else None
        self.files.append(dict(name=name, digest=digest, **meta))
        return digest

    def finalise(self) -> bytes:
        directory = bytearray(edition_element())
        frames = dict(self.enc.frames)
        # BTreeMap<Pathname, _> order: derive(Ord) on Vec<CborString> with Text < Binary (strings.rs:73)
        key = lambda f: [(0, c) if isinstance(c, str) else (1, c) for c in f["name"]]
        for f in sorted(self.files, key=key):
            if f["digest"] is not None and f["digest"] in frames:
                directory += frame_element(frames.pop(f["digest"]))
            directory += file_element(**f)
        for fr in frames.values():
            directory += frame_element(fr)
        digest = ref_path._blake3(bytes(directory))
        comp = self.enc.compress_frame(bytes(directory))  # same CCtx => same level + checksum (encode/directory.rs:95)
        self.out += comp
        tb = trailer_bytes(digest, -(len(comp) + 8 + 32 + EPILOGUE_LENGTH), len(directory))
        self.out += bytes([0x5F, 0x2A, 0x4D, 0x18]) + struct.pack("<I", len(tb)) + tb
        return bytes(self.out)


# ---- reader ----------------------------------------------------------------------------------------
def read_archive(data: bytes) -> dict:
    """Decoder::open + read_directory restated; returns header/trailer fields, decoded elements and a
    `content(digest)` callable that restores a frame with libzstd and checks its BLAKE3."""
    assert data[:12] == FILE_MAGIC, "header magic"
    ep = data[-EPILOGUE_LENGTH:]
    digest_type, off, usz, check, version = ep[0], struct.unpack("<q", ep[1:9])[0], struct.unpack("<Q", ep[9:17])[0], ep[17], ep[18]
    assert ep[19:22] == ZARC_MAGIC and digest_type == 1 and version == 1
    digest = data[-EPILOGUE_LENGTH - 32 : -EPILOGUE_LENGTH]
    x = 0
    for b in bytes([0, digest_type]) + digest + ep[:17] + b"\x00" + ep[18:]:
        x ^= b
    assert x == check, "trailer check byte"
    tr_frame = data[-EPILOGUE_LENGTH - 32 - 8 :]
    assert tr_frame[:4] == bytes([0x5F, 0x2A, 0x4D, 0x18]) and struct.unpack("<I", tr_frame[4:8])[0] == 54
    if off < 0:
        off += len(data)
    dir_len = ref_path.find_frame_compressed_size(data[off : len(data) - 62])
    assert off + dir_len == len(data) - 62, "directory frame must end where the trailer frame starts"
    directory = ref_path.ref_decompress_stream(data, off)
    assert len(directory) == usz, "directory uncompressed size"
    assert ref_path._blake3(directory) == digest, "directory digest"
    editions, files, frames = [], [], {}
    pos = 0
    while pos < len(directory):
        kind, ln, pad = directory[pos], struct.unpack("<H", directory[pos + 1 : pos + 3])[0], directory[pos + 3]
        assert pad == 0
        obj, end = cbor_decode(directory, pos + 4)
        assert end == pos + 4 + ln, "element length"
        pos = end
        if kind == 1:
            editions.append(obj)
        elif kind == 2:
            files.append(obj)
        elif kind == 3:
            frames[obj[2]] = ref_path.RefFrame(obj[0], obj[1], obj[2], obj[3], obj[4])
    dec = ref_path.RefDecoder(data, frames)
    return dict(digest=digest, directory_offset=off, directory_uncompressed=usz, directory=directory, editions=editions, files=files,
                frames=frames, content=dec.read_content_frame)
