#!/usr/bin/env python
"""Small driver for ncu: generate + sort once per iteration.  python tools/prof_sort.py --log2n 26"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import distributed_lsb_b200 as lsb  # noqa: E402
from distributed_lsb_b200 import lsbsort as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=26)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--radix", type=int, default=16)
ap.add_argument("--two-level", action="store_true")
ap.add_argument("--one-pass", action="store_true", help="LSB_FLAG_ONE_PASS: the one-pass kernel instead of two 8-bit steps")
ap.add_argument("--no-skip", action="store_true")
ap.add_argument("--mask", type=lambda x: int(x, 0), default=0xFFFFFFFFFFFFFFFF)
ap.add_argument("--and-draws", type=int, default=1)
ap.add_argument("--tune", action="append", default=[], help="key=value for lsb_tune, repeatable")
a = ap.parse_args()
for kv in a.tune:
    k, v = kv.split("=")
    lsb.tune(k, int(v))
flags = (L.FLAG_PHASE_EVENTS | (L.FLAG_TWO_LEVEL if a.two_level else 0) | (L.FLAG_ONE_PASS if a.one_pass else 0)
         | (L.FLAG_NO_SKIP if a.no_skip else 0))
with lsb.DistributedSorter(1 << a.log2n, ranks=1, radix_bits=a.radix, key_mask=a.mask, and_draws=a.and_draws, flags=flags) as s:
    for i in range(a.iters):
        s.generate()
        st = s.my_sort()
        s.verify()
        n = 1 << a.log2n
        print(f"iter {i}: sort {st.device_ms:.3f} ms = {n / st.device_ms / 1e3:.1f} M elem/s; hist {st.hist_ms:.3f} ms "
              f"({n * 16 / st.hist_ms / 1e6:.0f} GB/s); partition {st.partition_ms / st.partition_launches:.3f} ms/launch "
              f"({n * 32 / (st.partition_ms / st.partition_launches) / 1e6:.0f} GB/s) x{st.partition_launches}; "
              f"scan {st.scan_ms:.3f} ms; subpasses {[round(st.subpass_ms[k], 2) for k in range(min(st.subpasses, 32))]}")
