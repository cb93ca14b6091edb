"""Host-side logic on CPU: index arithmetic against the oracle, and the N>1 path's host
plumbing with two gloo processes (counts all-gather -> starts, communicator-id hand-off)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import distributed_lsb_b200 as lsb
from distributed_lsb_b200 import hostlogic as H
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,R", [(100, 3), (1, 4), (10, 8), (7, 3), (1 << 20, 4), (0, 2)])
def test_distribution_matches_oracle(n, R):
    assert H.per_rank(n, R) == max(O.per_rank(n, R), 1)
    assert [H.elements_here(n, R, r) for r in range(R)] == [O.here(n, R, r) for r in range(R)]
    per = H.per_rank(n, R)
    for g in (0, n // 2, max(n - 1, 0)):
        r, loc = H.global_to_local(per, g)
        assert H.local_to_global(per, r, loc) == g


def test_subpass_plan():
    assert H.num_passes(16) == 4 and H.num_passes(11) == 6 and H.num_passes(8) == 8
    assert H.subpasses(16) == [(0, 8), (8, 8), (16, 8), (24, 8), (32, 8), (40, 8), (48, 8), (56, 8)]
    assert H.plan_pass(11, 5) == (55, 9, 4, 5)  # the last 11-bit digit is 9 bits wide
    assert H.plan_pass(11, 5, one_pass=True) == (55, 9, 8, 1)  # one-pass kernel: low byte + 1 bit
    assert H.plan_pass(11, 0) == (0, 11, 5, 6) and H.plan_pass(16, 1) == (16, 16, 8, 8)
    assert H.subpasses(8) == [(8 * i, 8) for i in range(8)]
    for bits in range(1, 17):  # sub-digits tile the key exactly once
        covered = sorted(b for s, w in H.subpasses(bits) for b in range(s, s + w))
        assert covered == list(range(64))


@pytest.mark.parametrize("bits,R", [(16, 4), (8, 3), (11, 2)])
def test_scan_and_sendcounts_match_oracle(bits, R):
    n = 200001
    a = O.generate(n, R)[:n]
    for p in range(2):
        a2, counts, starts, sc = O.one_pass(a, n, R, bits, p)
        assert (H.global_starts(counts) == starts).all()
        assert (H.send_counts(counts, H.per_rank(n, R)) == sc).all()
        a = a2
    assert abs(H.remote_fraction(sc[0], 0) - (1 - sc[0, 0] / sc[0].sum())) < 1e-12


def test_roofline_arithmetic():
    m = 1 << 31  # 10.66 ms HBM at 6446.9 GB/s, 33.4 ms NVLink at G=8 (900 GB/s)
    assert abs(H.pass_roofline_ms(m, 0.0, 6446.9, 900) - 10.66) < 0.01
    assert abs(H.pass_roofline_ms(m, 7 / 8, 6446.9, 900) - 33.4) < 0.1


def _gloo_worker(rank, world, port, n, bits, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # what bench.py / mgpu_worker.py do around lsb_comm_init: rank 0 makes a token, all receive it
        token = [os.urandom(128) if rank == 0 else None]
        dist.broadcast_object_list(token, src=0)
        # each rank counts ITS shard, counts are all-gathered, every rank scans the same table
        a = O.generate(n, world)[:n]
        per = H.per_rank(n, world)
        mine = a[rank * per: rank * per + H.elements_here(n, world, rank)]
        shift, width, _, _ = H.plan_pass(bits, 0)
        local = np.bincount(((mine["key"] >> np.uint64(shift)) & np.uint64((1 << width) - 1)).astype(np.int64),
                            minlength=1 << width)
        gathered = [torch.zeros(1 << width, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(local))
        counts = torch.stack(gathered).numpy()
        starts = H.global_starts(counts)
        sent = H.send_counts(counts, per)
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # bench.py's max-over-ranks timing reduction
        q.put((rank, token[0], starts[:, rank].copy(), sent[rank].copy(), float(t.item())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bits", [16, 8])
def test_two_rank_gloo_counts_to_starts(bits):
    world, n = 2, 150001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + bits
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, n, bits, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a = O.generate(n, world)[:n]
    _, counts, starts, sc = O.one_pass(a, n, world, bits, 0)
    assert res[0][1] == res[1][1] and len(res[0][1]) == 128
    for r in range(world):
        assert (res[r][2] == starts[:, r]).all()
        assert (res[r][3] == sc[r]).all()
        assert res[r][4] == float(world)


@pytest.mark.parametrize("R,V", [(2, 4), (4, 4), (8, 2), (3, 5)])
def test_virtual_ranks_do_not_change_the_pass(R, V):
    """the pipelined multi-GPU pass treats part q of GPU g as rank g*V+q of the reference's algorithm:
    refining the block distribution leaves every pass's output unchanged (it is a stable partition of
    the global array), only the count tables get finer"""
    n = R * V * 5000  # divisible, so the refined blocks are exactly the parts
    a = O.generate(n, R)[:n]
    for p in range(2):
        coarse, counts_c, _, sc_c = O.one_pass(a, n, R, 16, p)
        fine, counts_f, _, sc_f = O.one_pass(a, n, R * V, 16, p)
        assert (coarse == fine).all()
        assert (counts_f.reshape(R, V, -1).sum(axis=1) == counts_c).all()
        assert (sc_f.reshape(R, V, R, V).sum(axis=(1, 3)) == sc_c).all()
        a = coarse


def test_part_boundaries_cut_the_shard_exactly():
    # the parts of a shard (virtual ranks) are contiguous, in order, cover [0, per) and ramp up then down
    for per, V in ((1, 16), (100, 8), (65535, 16), (65536, 16), (1 << 21, 16), ((1 << 31) + 12345, 16), (1 << 31, 8)):
        b = H.part_boundaries(per, V)
        assert len(b) == V + 1 and b[0] == 0 and b[-1] == per
        assert all(x <= y for x, y in zip(b, b[1:]))
        sizes = [y - x for x, y in zip(b, b[1:])]
        assert sum(sizes) == per
        if per >= V * 4096:
            assert sizes[0] < sizes[V // 2 - 1] and sizes[-1] < sizes[V // 2] and min(sizes) > 0
            assert max(sizes) < 0.2 * per


@pytest.mark.parametrize("nsuper,t1,lead,nx,tiles_last", [(1, 5, 1, 2, 3), (7, 3, 1, 2, 3), (9, 4, 3, 6, 1), (12, 2, 2, 3, 2), (5, 232, 3, 6, 17)])
def test_onepass_ticket_order_has_no_forward_dependency(nsuper, t1, lead, nx, tiles_last):
    """deadlock freedom of the one-pass kernel's spin-waits: every item is named by exactly one ticket, and every item
    an item waits for has a SMALLER ticket (so it is held by a CTA that is already running, or done)"""
    total = (nsuper + lead) * (t1 + 256)
    ticket_of = {}
    for t in range(total):
        item = H.onepass_ticket(t, nsuper, t1, lead, tiles_last)
        if item is not None:
            assert item not in ticket_of
            ticket_of[item] = t
    assert H.onepass_ticket(total, nsuper, t1, lead, tiles_last) is None
    k1 = sum((tiles_last if s == nsuper - 1 else t1) for s in range(nsuper))
    assert len(ticket_of) == k1 + 256 * nsuper
    for item, t in ticket_of.items():
        for dep in H.onepass_dependencies(item, nx, t1, tiles_last, nsuper):
            assert ticket_of[dep] < t, (item, dep)
