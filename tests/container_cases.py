"""Shared bodies of the container / CLI tests: the C++ host (zarc_b200/zarc-b200) against the oracle's
restatement of the reference container (oracle/ref_container.py).  The CPU suite runs them with
$ZARCGPU_LIB pointing at the SIMT-emulator build of the kernels, the GPU suite with the product library."""
import os
import stat
import subprocess

from oracle import ref_container, ref_path
from tests.golden.recipes import rand, text
from zarc_b200 import build


def run_cli(lib_path, args, cwd=None, env_extra=None, check=True):
    env = dict(os.environ, ZARCGPU_LIB=lib_path)
    env.update(env_extra or {})
    p = subprocess.run([build.build_host(), *args], cwd=cwd, env=env, capture_output=True, text=True, timeout=900)
    if check:
        assert p.returncode == 0, p.stderr
    return p


def make_tree(root, scale=1):
    """A small source-tree-like directory: nested dirs, duplicates, an empty file, a symlink, odd sizes."""
    files = {
        "a.txt": text(3000 * scale, 1),
        "b.bin": rand(5000 * scale, 2),
        "dup1.txt": text(7000 * scale, 3),
        "sub/dup2.txt": text(7000 * scale, 3),
        "sub/deep/er/c.txt": text(140_000 * scale, 4),  # two blocks
        "sub/empty": b"",
        "sub/one": b"x",
        "z/été.txt": text(900, 5),
    }
    for rel, data in files.items():
        p = os.path.join(root, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "wb") as f:
            f.write(data)
    os.makedirs(os.path.join(root, "emptydir"), exist_ok=True)
    os.symlink("a.txt", os.path.join(root, "link"))
    os.chmod(os.path.join(root, "a.txt"), 0o640)
    os.utime(os.path.join(root, "b.bin"), ns=(1_600_000_000_123_456_789, 1_500_000_000_000_000_000))
    return files


def check_pack_against_oracle(lib_path, tmp_path, env_extra=None):
    root = tmp_path / "tree"
    root.mkdir()
    files = make_tree(str(root))
    p = run_cli(lib_path, ["pack", "--output", str(tmp_path / "out.zarc"), "tree"], cwd=str(tmp_path), env_extra=env_extra)
    data = (tmp_path / "out.zarc").read_bytes()
    ar = ref_container.read_archive(data)  # header, trailer check byte, directory digest/size are asserted inside
    import base64

    assert p.stdout.strip() == "digest: " + base64.b64encode(ar["digest"]).decode()
    assert len(ar["editions"]) == 1 and ar["editions"][0][0] == 1 and ar["editions"][0][2] == 1
    by_name = {"/".join(f[1]): f for f in ar["files"]}
    # every walked entry is there, in BTreeMap<Pathname> order
    names = ["/".join(f[1]) for f in ar["files"]]
    assert names == sorted(names, key=lambda n: n.split("/"))
    assert set(names) == {"tree", "tree/emptydir", "tree/link", "tree/sub", "tree/sub/deep", "tree/sub/deep/er", "tree/z"} | {"tree/" + k for k in files}
    # content: digests bit-exact, frames valid Zstandard that libzstd restores, verified like FrameIterator
    for rel, content in files.items():
        f = by_name["tree/" + rel]
        assert f[2] == ref_path._blake3(content)
        got, ok = ar["content"](f[2])
        assert got == content and ok
        fr = ar["frames"][f[2]]
        assert fr.uncompressed == len(content) and fr.edition == 1
    # dedup: one frame for the two identical files; frames contiguous from offset 12
    assert by_name["tree/dup1.txt"][2] == by_name["tree/sub/dup2.txt"][2]
    assert len(ar["frames"]) == len({ref_path._blake3(c) for c in files.values()})
    frs = sorted(ar["frames"].values(), key=lambda fr: fr.offset)
    assert frs[0].offset == 12
    for a, b in zip(frs, frs[1:]):
        assert a.offset + a.length == b.offset
    assert frs[-1].offset + frs[-1].length == ar["directory_offset"]
    for fr in frs:
        assert ref_path.find_frame_compressed_size(data[fr.offset : fr.offset + fr.length]) == fr.length
    # metadata bookkeeping: specials, mode, owner, timestamps
    assert by_name["tree/emptydir"][7] == [1] and by_name["tree/sub"][7] == [1]
    assert by_name["tree/link"][7] == [10, ["a.txt"]] and 2 not in by_name["tree/link"]
    st = os.lstat(root / "a.txt")
    assert by_name["tree/a.txt"][3] == st.st_mode and stat.S_IMODE(by_name["tree/a.txt"][3]) == 0o640
    assert by_name["tree/a.txt"][4][0] == st.st_uid and by_name["tree/a.txt"][5][0] == st.st_gid
    ts = by_name["tree/b.bin"][6]
    assert ts[2] == ref_container.Tag(0, "2017-07-14T02:40:00+00:00") and ts[3] == ref_container.Tag(0, "2020-09-13T12:26:40.123456789+00:00")
    # a Frame element comes right before the first File element that links it
    seen, pos, d = set(), 0, ar["directory"]
    order = []
    while pos < len(d):
        kind, ln = d[pos], int.from_bytes(d[pos + 1 : pos + 3], "little")
        obj, _ = ref_container.cbor_decode(d, pos + 4)
        order.append((kind, obj))
        pos += 4 + ln
    assert order[0][0] == 1
    for i, (kind, obj) in enumerate(order):
        if kind == 2 and 2 in obj and obj[2] not in seen:
            assert order[i - 1][0] == 3 and order[i - 1][1][2] == obj[2]
            seen.add(obj[2])
    return data, files


def check_unpack_of_reference_archive(lib_path, tmp_path, levels=(1, 3, 9)):
    """Archives written the reference's way (libzstd frames, reference directory layout) restore byte-identically."""
    import base64

    for level in levels:
        w = ref_container.RefArchiveWriter(level=level)
        contents = {
            ("r", "one.txt"): text(50_000, 10 + level),
            ("r", "two.bin"): rand(20_000, 11),
            ("r", "d", "three.txt"): text(300_000, 12),
            ("r", "d", "dup.txt"): text(50_000, 10 + level),
            ("r", "d", "empty"): b"",
        }
        w.add_file(["r"], None, special=[1], mode=0o40755)
        w.add_file(["r", "d"], None, special=[1], mode=0o40700)
        for name, c in contents.items():
            w.add_file(list(name), c, mode=0o100604, timestamps={2: 1_700_000_000.5})
        arc = w.finalise()
        work = tmp_path / f"ref{level}"
        work.mkdir()
        (work / "ref.zarc").write_bytes(arc)
        p = run_cli(lib_path, ["unpack", "ref.zarc"], cwd=str(work))
        assert "unpacked 5 files" in p.stderr
        assert "digest: " + base64.b64encode(ref_container.read_archive(arc)["digest"]).decode() in p.stderr
        for name, c in contents.items():
            path = work.joinpath(*name)
            assert path.read_bytes() == c
            assert stat.S_IMODE(path.stat().st_mode) == 0o604
            assert abs(path.stat().st_mtime - 1_700_000_000.5) < 1e-6
        assert stat.S_IMODE((work / "r" / "d").stat().st_mode) == 0o700
        ls = run_cli(lib_path, ["list-files", "ref.zarc"], cwd=str(work)).stdout.split("\n")
        assert ls[:3] == ["r/", "r/d/", "r/d/dup.txt"] and "r/one.txt" in ls
        only = run_cli(lib_path, ["list-files", "--only-files", "--filter", "t.*txt", "ref.zarc"], cwd=str(work)).stdout.split()
        assert only == ["r/d/three.txt"]
        # --verify with the right and a wrong digest
        good = base64.b64encode(ref_container.read_archive(arc)["digest"]).decode()
        assert run_cli(lib_path, ["unpack", "--verify", good, "ref.zarc"], cwd=str(work)).returncode == 0
        bad = run_cli(lib_path, ["unpack", "--verify", base64.b64encode(b"\0" * 32).decode(), "ref.zarc"], cwd=str(work), check=False)
        assert bad.returncode != 0 and "integrity failure" in bad.stderr


def check_roundtrip_and_errors(lib_path, tmp_path, env_extra=None):
    data, files = check_pack_against_oracle(lib_path, tmp_path, env_extra=env_extra)
    out = tmp_path / "restore"
    out.mkdir()
    p = run_cli(lib_path, ["unpack", str(tmp_path / "out.zarc")], cwd=str(out), env_extra=env_extra)
    assert f"unpacked {len(files)} files" in p.stderr
    for rel, content in files.items():
        assert (out / "tree" / rel).read_bytes() == content
    assert (out / "tree" / "emptydir").is_dir()
    assert stat.S_IMODE((out / "tree" / "a.txt").stat().st_mode) == 0o640
    assert (out / "tree" / "b.bin").stat().st_mtime_ns == 1_500_000_000_000_000_000
    # corrupt one content frame: unpack reports libzstd's error name and fails
    ar = ref_container.read_archive(data)
    fr = max(ar["frames"].values(), key=lambda f: f.length)
    broken = bytearray(data)
    broken[fr.offset + fr.length // 2] ^= 0x55
    (tmp_path / "broken.zarc").write_bytes(broken)
    bad = run_cli(lib_path, ["unpack", str(tmp_path / "broken.zarc")], cwd=str(out), check=False)
    assert bad.returncode != 0 and ("zstd:" in bad.stderr or "verification failed" in bad.stderr)
    # corrupt the trailer: the check byte catches it (decode/open.rs:108-118)
    broken = bytearray(data)
    broken[-30] ^= 1
    (tmp_path / "broken2.zarc").write_bytes(broken)
    bad = run_cli(lib_path, ["list-files", str(tmp_path / "broken2.zarc")], check=False)
    assert bad.returncode != 0 and "check byte" in bad.stderr
    # corrupt the directory frame: digest or checksum failure
    broken = bytearray(data)
    broken[ar["directory_offset"] + 12] ^= 0x10
    (tmp_path / "broken3.zarc").write_bytes(broken)
    bad = run_cli(lib_path, ["list-files", str(tmp_path / "broken3.zarc")], check=False)
    assert bad.returncode != 0
