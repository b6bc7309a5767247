"""Aggregate an ncu source-page (SASS) CSV per enclosing C function of the .cu source (via nvdisasm line
info): share of executed warp-instructions, of stall samples, and the dominant stall reasons.

usage: python tools/ncu_funcs.py <report.ncu-rep> <cubin> <kernel-substring> <source.cu> [--lines N]
"""
import csv, io, re, subprocess, sys
from collections import defaultdict
sys.path.insert(0, __import__("os").path.dirname(__file__))
import ncu_lines

STALLS = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_no_inst", "stall_branch_resolving", "stall_selected", "stall_math", "stall_not_selected", "stall_barrier", "stall_lg", "stall_mio", "stall_dispatch"]


def main():
    rep, cubin, kernel, srcf = sys.argv[1:5]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 0
    src = open(srcf).read().splitlines()
    base_name = srcf.split("/")[-1]
    fn_at, cur = {}, "?"
    for i, l in enumerate(src, 1):
        m = re.match(r"^(ZG_DEV_NOINLINE|ZG_DEV|__global__|template).*?(\w+)\(", l)
        if m:
            cur = m.group(2)
        if l.startswith("k_"):
            cur = l.split("(")[0]
        fn_at[i] = cur
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[hi]
    ia, ii, isamp, ith = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
    ist = [hdr.index(s) for s in STALLS]
    lm = ncu_lines.line_map(cubin, kernel)
    base = None
    agg = defaultdict(lambda: [0] * (3 + len(STALLS)))
    lagg = defaultdict(lambda: [0] * (3 + len(STALLS)))
    tot = [0] * (3 + len(STALLS))
    for r in rows[hi + 1:]:
        if len(r) <= ii or not r[ia].startswith("0x"):
            continue
        a = int(r[ia], 16)
        if base is None:
            base = a
        f, l = lm.get(a - base, ("?", 0))
        key = fn_at.get(l, "?") if f == base_name else f
        v = [int(r[ii] or 0), int(r[isamp] or 0), int(r[ith] or 0)] + [int(r[k] or 0) for k in ist]
        for k in range(len(v)):
            agg[key][k] += v[k]
            lagg[(f, l)][k] += v[k]
            tot[k] += v[k]
    print(f"total warp-inst {tot[0]:,}  samples {tot[1]:,}  avg threads/inst {tot[2] / max(tot[0], 1):.1f}")
    print("stalls: " + "  ".join(f"{s[6:]} {100 * tot[3 + k] / max(tot[1], 1):.1f}%" for k, s in enumerate(STALLS)))

    def show(key, v):
        top = sorted(range(len(STALLS)), key=lambda k: -v[3 + k])[:3]
        st = " ".join(f"{STALLS[k][6:]}={100 * v[3 + k] / max(v[1], 1):.0f}%" for k in top)
        print(f"{str(key):34s} inst {100 * v[0] / tot[0]:5.1f}%  samples {100 * v[1] / max(tot[1], 1):5.1f}%  thr/inst {v[2] / max(v[0], 1):4.1f}  {st}")

    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        show(key, v)
    if nlines:
        print("--- hottest lines")
        for key, v in sorted(lagg.items(), key=lambda kv: -kv[1][1])[:nlines]:
            show(f"{key[0]}:{key[1]}", v)


if __name__ == "__main__":
    main()
