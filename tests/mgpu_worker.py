"""Worker for tests/test_multigpu.py: one process per GPU (torchrun), checks this rank's shard
of the multi-GPU sort bit-for-bit against the oracle.  Exit code 0 == all checks passed."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import distributed_lsb_b200 as lsb  # noqa: E402
from oracle import oracle as O  # noqa: E402


def make(n, ranks, rank, world, **kw):
    s = lsb.DistributedSorter(n, ranks=ranks, world_size=world, world_rank=rank,
                              device=int(os.environ.get("LOCAL_RANK", "0")), **kw)
    ids = [lsb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    s.comm_init(ids[0])
    return s


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    ALL = 0xFFFFFFFFFFFFFFFF
    cases = [  # (n, pcg streams R, radix bits, key mask, and_draws)
        (1 << 20, 4, 16, ALL, 1), ((1 << 20) + 12345, world, 16, ALL, 1), (100, 4, 16, ALL, 1), (3, 4, 16, ALL, 1),
        (1, 1, 16, ALL, 1), (65537, 3, 16, ALL, 1), (1 << 20, 4, 8, ALL, 1), (1 << 20, 4, 11, ALL, 1),
        (700001, 4, 16, 0xFFFFFF, 1), (500000, 2, 16, ALL, 3), (300000, 4, 16, 0, 1), (1 << 22, world, 16, ALL, 1),
    ]
    from distributed_lsb_b200 import lsbsort as L
    for ci, (n, R, bits, mask, k) in enumerate(cases):
        g = O.generate(n, R, key_mask=mask, and_draws=k)
        want = O.sort(g, n, R, bits)
        # the part sort as two 8-bit steps (default) or as one launch of the one-pass kernel
        flags = [0, L.FLAG_ONE_PASS, L.FLAG_NO_SKIP][ci % 3]
        s = make(n, R, rank, world, radix_bits=bits, key_mask=mask, and_draws=k, flags=flags)
        lo, hi = s.first_global, s.first_global + s.here
        s.generate()
        assert (s.download() == g[lo:hi]).all(), ("generate", n, R)
        st = s.my_sort()
        got = s.download()
        assert (got == want[lo:hi]).all(), ("sort", n, R, bits, hex(mask), k)
        v = s.verify()
        assert v.elements == n and list(v.checksum) == O.checksum(g[:n]), ("verify", n)
        assert sum(st.sent[:world]) == s.here
        s.close()
    # pass by pass against the reference's per-pass tables, ranks == GPUs: the default (pipelined, virtual ranks)
    # pass at a size where a part is a single tile and at one where it is dozens of tiles, then the one-pass part sort
    for n, bits, flags in ((1 << 20, 16, 0), (1 << 24, 16, 0), (1 << 22, 16, L.FLAG_ONE_PASS)):
        a = O.generate(n, world)[:n]
        s = make(n, world, rank, world, radix_bits=bits, flags=flags)
        s.generate()
        lo, hi = s.first_global, s.first_global + s.here
        for p in range(s.num_passes()):
            want, counts, starts, sc = O.one_pass(a, n, world, bits, p)
            assert (s.histogram(p) == counts[rank]).all(), ("counts", n, p)
            assert (s.starts(p) == starts[:, rank]).all(), ("starts", n, p)
            st = s.global_shuffle(p)
            assert list(st.sent[:world]) == sc[rank].tolist(), ("sendcounts", n, p, list(st.sent[:world]), sc[rank].tolist())
            a = want
            assert (s.download() == a[lo:hi]).all(), ("array after pass", n, p)
        s.close()
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}/{world}: multi-GPU parity ok")


if __name__ == "__main__":
    main()
