"""Markdown table of the committed bench lines (profiles/r02_bench_n{1,2,4,8}.json): the numbers README.md quotes."""
import json
import os
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prefix = sys.argv[1] if len(sys.argv) > 1 else "r02_bench_n"
rows = []
for n in (1, 2, 4, 8):
    p = os.path.join(root, "profiles", f"{prefix}{n}.json")
    if not os.path.exists(p):
        continue
    d = json.loads(open(p).read().strip().splitlines()[-1])
    g = lambda *ks: (lambda v: v)(__import__("functools").reduce(lambda a, k: (a or {}).get(k) if isinstance(a, dict) else None, ks, d))
    rows.append((n, d))
hdr = ["GPUs", "value QPS", "ms/step", "e2e QPS (callers)", "kernel ms", "GB/s", "frac", "1-query QPS", "1-query HBM frac", "1024-batch QPS",
       "cfg5 QPS", "sustained QPS", "SM MHz"]
print("| " + " | ".join(hdr) + " |")
print("|" + "---|" * len(hdr))
for n, d in rows:
    r, e = d["roofline"], d.get("e2e") or {}
    sq, b, c5, su = d.get("single_query") or {}, d.get("batched") or {}, d.get("cfg5") or {}, d.get("sustained") or {}
    f = lambda v, fmt="{:,.0f}": fmt.format(v) if isinstance(v, (int, float)) else "–"
    print("| " + " | ".join([str(n), f(d["value"]), f(d["ms_per_step"], "{:.3f}"), f"{f(e.get('value'))} ({e.get('callers', '–')})",
                             f(r.get("mean_launch_ms"), "{:.3f}"), f(r.get("achieved")), f(r.get("frac"), "{:.2f}"), f(sq.get("value")),
                             f(sq.get("aggregate_hbm_frac"), "{:.2f}"), f(b.get("value")), f(c5.get("value")), f(su.get("value")),
                             f((d.get("clocks") or {}).get("sm_mhz"))]) + " |")
