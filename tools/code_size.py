"""Static code size of a kernel per enclosing source function (nvdisasm line info): where the bytes of a kernel are.

usage: python tools/code_size.py <object.o> <kernel-substring> <source file>"""
import re
import subprocess
import sys
from collections import Counter


def main():
    obj, kernel, srcf = sys.argv[1:4]
    subprocess.run(["cuobjdump", "-xelf", "all", obj], capture_output=True)
    import glob, os
    cubin = sorted(glob.glob("*.cubin"))[0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    for f in glob.glob("*.cubin"):
        os.remove(f)
    src = open(srcf).read().splitlines()
    base = srcf.split("/")[-1]
    fn_at, cur = {}, "?"
    for i, l in enumerate(src, 1):
        m = re.match(r"^(?:template\s*<[^>]*>\s*)?(?:ZG_DEV_NOINLINE|ZG_DEV|__global__|static|ZG_HD)[^;{]*?\b([A-Za-z_]\w*)\s*\(", l)
        if m and not l.rstrip().endswith(";"):
            cur = m.group(1)
        fn_at[i] = cur
    cnt, infn, where = Counter(), False, ("?", 0)
    for ln in txt.splitlines():
        if ln.startswith("\t.text.") or ln.startswith(".text."):
            infn = kernel in ln
        if not infn:
            continue
        mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if mm:
            where = (mm.group(1).split("/")[-1], int(mm.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            f, line = where
            cnt[fn_at.get(line, "?") if f == base else f] += 1
    tot = sum(cnt.values())
    print(f"{kernel}: {tot} instructions, {tot * 16 / 1024:.1f} KB")
    for k, v in cnt.most_common(25):
        print(f"  {k:34s} {v:6d}  {100 * v / tot:5.1f}%")


if __name__ == "__main__":
    main()
