"""Aggregate an ncu source-page (SASS) CSV per CUDA source line, using nvdisasm line info.

usage: python tools/ncu_lines.py <report.ncu-rep> <cubin> <kernel-substring> [top]
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def line_map(cubin, kernel):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    m = {}
    cur = None
    infn = False
    for ln in txt.splitlines():
        if ln.startswith("\t.text.") or ln.startswith(".text."):
            infn = kernel in ln
        if not infn:
            continue
        mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if mm:
            cur = (mm.group(1).split("/")[-1], int(mm.group(2)))
            continue
        mm = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if mm and cur:
            m[int(mm.group(1), 16)] = cur
    return m


def main():
    rep, cubin, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[hi]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ith = hdr.index("Thread Instructions Executed")
    lm = line_map(cubin, kernel)
    base = None
    agg = defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r in rows[hi + 1:]:
        if len(r) <= ii or not r[ia].startswith("0x"):
            continue
        a = int(r[ia], 16)
        if base is None:
            base = a
        key = lm.get(a - base, ("?", 0))
        v = (int(r[ii] or 0), int(r[isamp] or 0), int(r[ith] or 0))
        for k in range(3):
            agg[key][k] += v[k]
            tot[k] += v[k]
    print(f"total warp-inst {tot[0]:,}  samples {tot[1]:,}  avg threads/inst {tot[2] / max(tot[0], 1):.1f}")
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{key[0]}:{key[1]:<5} inst {100 * v[0] / tot[0]:5.1f}%  samples {100 * v[1] / max(tot[1], 1):5.1f}%  thr/inst {v[2] / max(v[0], 1):4.1f}")


if __name__ == "__main__":
    main()
