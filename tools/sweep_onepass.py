#!/usr/bin/env python
"""One-GPU sweep of the one-pass kernel's tunables (lsb_tune): prints sort time per configuration.
    python tools/sweep_onepass.py --log2n 30 --set op_lead=2,op_nx=4 --set op_t1=128 --set two_step ...
Every configuration runs with LSB_FLAG_ONE_PASS except the one named `two_step` (the default shape)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import distributed_lsb_b200 as lsb  # noqa: E402
from distributed_lsb_b200 import lsbsort as L  # noqa: E402

DEFAULTS = {"op_cfg": 1, "op_t1": 232, "op_nx": 6, "op_lead": 3, "op_hints": 15, "pt_direct": 1, "pt_chunks": 2}
ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=30)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--radix", type=int, default=16)
ap.add_argument("--mask", type=lambda x: int(x, 0), default=0xFFFFFFFFFFFFFFFF)
ap.add_argument("--and-draws", type=int, default=1)
ap.add_argument("--set", action="append", default=[], help="comma-separated key=value list; 'two_step' = the default two-step shape")
a = ap.parse_args()
n = 1 << a.log2n
for cfg in a.set or [""]:
    flags = L.FLAG_PHASE_EVENTS | L.FLAG_NO_SKIP | L.FLAG_ONE_PASS
    kv = dict(DEFAULTS)
    for item in filter(None, cfg.split(",")):
        if item == "two_step":
            flags &= ~L.FLAG_ONE_PASS
        elif item == "two_level":
            flags |= L.FLAG_TWO_LEVEL
        else:
            k, v = item.split("=")
            kv[k] = int(v)
    for k, v in kv.items():
        lsb.tune(k, v)
    try:
        with lsb.DistributedSorter(n, ranks=1, radix_bits=a.radix, key_mask=a.mask, and_draws=a.and_draws, flags=flags) as s:
            best, rows = None, []
            for i in range(a.iters):
                s.generate()
                before = s.checksum()
                st = s.my_sort()
                v = s.verify()
                assert list(v.checksum) == before and v.elements == n
                rows.append(st.device_ms)
                best = st if best is None or st.device_ms < best.device_ms else best
            sub = [round(best.subpass_ms[k], 3) for k in range(min(best.subpasses, 32))]
            print(f"[{cfg or 'default'}] n=2^{a.log2n} radix {a.radix}: sort ms {['%.3f' % r for r in rows]} best {best.device_ms:.3f} "
                  f"= {n / best.device_ms / 1e3:.0f} M/s; count {best.hist_ms:.3f} scan {best.scan_ms:.3f} scatter {sub} verified", flush=True)
    except lsb.LsbError as e:
        print(f"[{cfg}] FAILED: {e}", flush=True)
