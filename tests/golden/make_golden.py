"""Generates tests/golden/golden_search.json from the CPU oracle (oracle/syzgy_oracle.c).

The reference is Go and cannot run in this image (no Go toolchain), so these vectors are
produced by the restated oracle, not by the reference binary: they pin the oracle against
regressions and give the GPU tests committed expected outputs.  The only values here that
come from the reference's own tests are marked "reference_kat".

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as o  # noqa: E402

CASES = [
    # name, n, dims, bits, metric, k, radius, filter_mod
    ("q8_cos_k10", 3000, 96, 8, o.COSINE, 10, 0.0, 0),
    ("q4_euc_k10", 3000, 128, 4, o.EUCLIDEAN, 10, 0.0, 0),
    ("q4_euc_odd_dims", 500, 7, 4, o.EUCLIDEAN, 5, 0.0, 0),
    ("q16_euc_k20", 2000, 40, 16, o.EUCLIDEAN, 20, 0.0, 0),
    ("q16_cos_k3", 1500, 33, 16, o.COSINE, 3, 0.0, 0),
    ("f32_cos_k10", 2000, 50, 32, o.COSINE, 10, 0.0, 0),
    ("f32_euc_k7", 2000, 19, 32, o.EUCLIDEAN, 7, 0.0, 0),
    ("f64_cos_k10_filter", 2500, 48, 64, o.COSINE, 10, 0.0, 3),
    ("f64_euc_radius", 2500, 24, 64, o.EUCLIDEAN, 0, 2.45, 0),
    ("q8_cos_radius_filter", 3000, 64, 8, o.COSINE, 0, 0.43, 2),
    ("q8_euc_k40", 3000, 100, 8, o.EUCLIDEAN, 40, 0.0, 0),
]


def make_case(name, n, dims, bits, metric, k, radius, filter_mod, seed):
    codes = o.synth_rows(seed, 0, n, dims, bits)
    # ids deliberately not in lexicographic order of insertion: 3*i+1 mixes digit counts
    ids = (np.arange(n, dtype=np.uint64) * 3 + 1)
    queries = o.synth_queries(seed + 1, 0, 3, dims)
    passmask = None if not filter_mod else (ids % filter_mod == 0).astype(np.uint8)
    out = []
    for q in queries:
        rid, rd, pct = o.search_exact(codes, ids, dims, bits, metric, q, k=k, radius=radius, passmask=passmask)
        out.append({"ids": [int(x) for x in rid], "dist": [float.hex(float(x)) for x in rd], "percent": pct})
    return {"name": name, "n": n, "dims": dims, "bits": bits, "metric": metric, "k": k, "radius": radius,
            "filter_mod": filter_mod, "seed": seed, "results": out}


def main():
    doc = {
        "reference_kat": {"euclidean": {"a": [1.0, 2.0, 3.0], "b": [4.0, 5.0, 6.0],
                                        "expected": 5.196152422706632,
                                        "source": "collection_test.go:12-21"}},
        "restated_codec": {  # SURVEY.md 8 a7 (restated from quantization.go:5-36, not run in Go)
            "inputs": [-1.5, -1, -.5, 0, .1, .5, 1, 2],
            "4": [0, 0, 4, 8, 8, 11, 15, 15],
            "8": [0, 0, 64, 128, 140, 191, 255, 255],
            "16": [0, 0, 16384, 32768, 36044, 49151, 65535, 65535],
            "dequantize_128_8": float.hex(0.0039215686274509665),
        },
        "cases": [make_case(*c, seed=1000 + i) for i, c in enumerate(CASES)],
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_search.json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
