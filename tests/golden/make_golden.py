"""Generates tests/golden/kat.json from the oracle (oracle/ref_path.py: libzstd 1.5.5 with the
reference's call sequence, Python blake3, xxhash).  Run in the build container:

    python tests/golden/make_golden.py

Inputs are tiny deterministic recipes (tests/golden/recipes.py) so the fixture stays small; every
frame is stored as base64 together with the input's BLAKE3 digest and XXH64.
"""
import base64
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import blake3  # noqa: E402
import xxhash  # noqa: E402

from oracle import ref_path  # noqa: E402
from tests.golden.recipes import RECIPES, make_input  # noqa: E402


def main():
    out = {"libzstd_version": ref_path.ZSTD_VERSION, "blake3_pkg": blake3.__version__, "cases": []}
    for name in RECIPES:
        data = make_input(name)
        case = {
            "recipe": name,
            "len": len(data),
            "blake3": blake3.blake3(data).hexdigest(),
            "xxh64": f"{xxhash.xxh64(data, seed=0).intdigest():016x}",
            "frames": {},
        }
        for level in (1, 3, 9):
            if len(data) > 40000 and level != 3:
                continue
            case["frames"][str(level)] = base64.b64encode(ref_path.ref_compress(data, level=level, checksum=True)).decode()
        out["cases"].append(case)
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("cases:", len(out["cases"]), "bytes:", os.path.getsize(os.path.join(HERE, "kat.json")))


if __name__ == "__main__":
    main()
