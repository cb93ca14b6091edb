#!/usr/bin/env python
"""A few verified sorts with whatever library LSB_LIBRARY names, as fast as a process can start:
    LSB_LIBRARY=tools/bin/liblsbsort_exp.so python tools/ab_quick.py 28 3"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import distributed_lsb_b200 as lsb  # noqa: E402
from distributed_lsb_b200 import lsbsort as L  # noqa: E402

log2n, iters = int(sys.argv[1]), int(sys.argv[2])
n = 1 << log2n
with lsb.DistributedSorter(n, ranks=1, flags=L.FLAG_PHASE_EVENTS | L.FLAG_NO_SKIP) as s:
    for i in range(iters):
        s.generate()
        before = s.checksum()
        st = s.my_sort()
        v = s.verify()
        ok = list(v.checksum) == before and v.elements == n and not v.order_violations
        print(f"{os.path.basename(lsb.library_path())} 2^{log2n} iter {i}: sort {st.device_ms:.3f} ms, scatter launch "
              f"{st.partition_ms / max(st.partition_launches, 1):.4f} ms, verified={ok}", flush=True)
