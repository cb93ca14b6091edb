#!/usr/bin/env python
"""Per-phase timing of the multi-GPU sort (run under torchrun, one rank per GPU)."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import distributed_lsb_b200 as lsb  # noqa: E402
from distributed_lsb_b200 import lsbsort as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=28, help="log2 elements per GPU")
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--radix", type=int, default=16)
ap.add_argument("--one-pass", action="store_true")
ap.add_argument("--mask", type=lambda x: int(x, 0), default=0xFFFFFFFFFFFFFFFF)
ap.add_argument("--tune", action="append", default=[], help="key=value for lsb_tune, repeatable")
a = ap.parse_args()
for kv in a.tune:
    k, v = kv.split("=")
    lsb.tune(k, int(v))
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = world << a.log2n
s = lsb.DistributedSorter(n, ranks=world, world_size=world, world_rank=rank, device=lr, radix_bits=a.radix,
                          key_mask=a.mask, flags=L.FLAG_PHASE_EVENTS | L.FLAG_NO_SKIP | (L.FLAG_ONE_PASS if a.one_pass else 0))
ids = [lsb.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
s.comm_init(ids[0])
for i in range(a.iters):
    s.generate()
    s.barrier()
    st = s.my_sort()
    v = s.verify()
    if rank == 0:
        sub = [round(st.subpass_ms[k], 3) for k in range(min(st.subpasses, 32))]
        m = s.here
        tail = (f"exchange kernels {st.exchange_ms:.2f} ms busy = "
                f"{m * 16 * (world - 1) / world * st.passes / max(st.exchange_ms, 1e-9) / 1e6:.0f} GB/s out per GPU while running")
        print(f"[{' '.join(a.tune) or 'default'}{' one-pass' if a.one_pass else ''}] iter {i}: sort {st.device_ms:.2f} ms = {n / st.device_ms / 1e3:.0f} M elem/s; count {st.hist_ms:.2f} ms, "
              f"scan+coll {st.scan_ms:.2f} ms, partitions {sub} ms; sent {list(st.sent[:world])}; {tail}", flush=True)
s.close()
dist.destroy_process_group()
