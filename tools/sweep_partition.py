#!/usr/bin/env python
"""One-GPU A/B of partition_kernel's compile-time variants and launch knobs (lsb_tune "pt_*") on the default
two-step path: every configuration sorts 2^k fresh elements a few times in ONE process, verifies order,
element count and multiset hash every time, and prints the time per sort and per scatter launch.
    python tools/sweep_partition.py --log2n 30 --set "" --set pt_variant=1 --set pt_variant=1,pt_pf_tiles=148 ...
    --out FILE   also writes {"baseline_ms", "rows": [...], "winner": {key: value} or null}: the verified configuration with the lowest median
                 time if it beats the first one (the baseline) by more than --margin, else null."""
import argparse
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import distributed_lsb_b200 as lsb  # noqa: E402
from distributed_lsb_b200 import lsbsort as L  # noqa: E402

DEFAULTS = {"pt_direct": 1, "pt_chunks": 2, "pt_variant": 1, "pt_pf_tiles": 0}  # the library's built-in values
ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=30)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--radix", type=int, default=16)
ap.add_argument("--set", action="append", default=[], help="comma-separated key=value list; '' = the library defaults as built")
ap.add_argument("--margin", type=float, default=0.01)
ap.add_argument("--out", default="")
a = ap.parse_args()
n = 1 << a.log2n
rows = []
for idx, cfg in enumerate(a.set or [""]):
    kv = dict(DEFAULTS) if idx else {}  # the first configuration runs the library exactly as built
    for item in filter(None, cfg.split(",")):
        k, v = item.split("=")
        kv[k] = int(v)
    row = {"set": cfg, "tune": kv, "ok": False}
    try:
        for k, v in kv.items():
            lsb.tune(k, v)
        with lsb.DistributedSorter(n, ranks=1, radix_bits=a.radix, flags=L.FLAG_PHASE_EVENTS | L.FLAG_NO_SKIP) as s:
            ms, launch = [], []
            for i in range(a.iters):
                s.generate()
                before = s.checksum()
                st = s.my_sort()
                v = s.verify()
                assert list(v.checksum) == before and v.elements == n and not v.order_violations
                ms.append(st.device_ms)
                launch.append(st.partition_ms / max(st.partition_launches, 1))
        # the median, not the best: a faster kernel runs into the power cap after the first sort or two
        med, lmed = statistics.median(ms), statistics.median(launch)
        row.update(ok=True, sort_ms=med, best_ms=min(ms), all_ms=[round(x, 3) for x in ms], launch_ms=lmed)
        print(f"[{cfg or 'as built'}] n=2^{a.log2n}: sort ms {row['all_ms']} median {med:.3f} = {n / med / 1e3:.0f} M/s; "
              f"scatter launch {lmed:.3f} ms = {n * 32 / lmed / 1e6:.0f} GB/s; verified", flush=True)
    except (lsb.LsbError, AssertionError) as e:
        row["error"] = repr(e)
        print(f"[{cfg}] FAILED: {e!r}", flush=True)
    rows.append(row)
for k, v in DEFAULTS.items():
    lsb.tune(k, v)
winner = None
if rows and rows[0]["ok"]:
    base = rows[0]["sort_ms"]
    ok = [r for r in rows[1:] if r["ok"]]
    if ok:
        best = min(ok, key=lambda r: r["sort_ms"])
        if best["sort_ms"] < base * (1.0 - a.margin):
            winner = {k: v for k, v in best["tune"].items() if DEFAULTS.get(k) != v}
    print(f"baseline {base:.3f} ms; winner: {winner}", flush=True)
if a.out:
    with open(a.out, "w") as f:
        json.dump({"log2n": a.log2n, "baseline_ms": rows[0].get("sort_ms") if rows else None, "rows": rows, "winner": winner}, f, indent=1)
