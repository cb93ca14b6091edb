#!/usr/bin/env python
"""Per-phase timing of the multi-GPU sort (run under torchrun, one rank per GPU).  Several tunable sets
can be measured in one launch: --sets "vparts=8;vparts=16,ex_ctas=2;one_pass"  (ALWAYS wrap in `timeout`
on a GPU box: a rank that dies leaves its peers in a collective)."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import distributed_lsb_b200 as lsb  # noqa: E402
from distributed_lsb_b200 import lsbsort as L  # noqa: E402

DEFAULTS = {"vparts": 8, "ex_ctas": 1, "ex_threads": 0, "ex_u": 4, "op_ctas_mgpu": 3}
ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=28, help="log2 elements per GPU")
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--radix", type=int, default=16)
ap.add_argument("--mask", type=lambda x: int(x, 0), default=0xFFFFFFFFFFFFFFFF)
ap.add_argument("--sets", default="", help="';'-separated sets of ','-separated key=value tunables; 'one_pass' = LSB_FLAG_ONE_PASS")
a = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = world << a.log2n
for cfg in a.sets.split(";") if a.sets else [""]:
    flags = L.FLAG_PHASE_EVENTS | L.FLAG_NO_SKIP
    kv = dict(DEFAULTS)
    for item in filter(None, cfg.split(",")):
        if item == "one_pass":
            flags |= L.FLAG_ONE_PASS
        else:
            k, v = item.split("=")
            kv[k] = int(v)
    for k, v in kv.items():
        lsb.tune(k, v)
    s = lsb.DistributedSorter(n, ranks=world, world_size=world, world_rank=rank, device=lr, radix_bits=a.radix,
                              key_mask=a.mask, flags=flags)
    ids = [lsb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    s.comm_init(ids[0])
    for i in range(a.iters):
        s.generate()
        s.barrier()
        st = s.my_sort()
        v = s.verify()
        t = torch.tensor([st.device_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0 and i == a.iters - 1:
            sub = [round(st.subpass_ms[k], 2) for k in range(min(st.subpasses, 8))]
            m = s.here
            print(f"[{cfg or 'default'}] {world} GPUs x 2^{a.log2n}: sort {t.item():.2f} ms (max over ranks) = {n / t.item() / 1e3:.0f} M elem/s; "
                  f"count {st.hist_ms:.2f} ms, scan+coll {st.scan_ms:.2f} ms, first scatter launches {sub} ms; exchange kernels busy "
                  f"{st.exchange_ms:.1f} ms = {m * 16 * (world - 1) / world * st.passes / max(st.exchange_ms, 1e-9) / 1e6:.0f} GB/s out per GPU "
                  f"while running; {v.elements} elements verified", flush=True)
    s.close()
dist.destroy_process_group()
