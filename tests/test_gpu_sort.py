"""Parity of the CUDA path (through the C ABI) against the oracle -- needs a B200."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import distributed_lsb_b200 as lsb  # noqa: E402
from distributed_lsb_b200 import lsbsort as L  # noqa: E402
from oracle import oracle as O  # noqa: E402

ALL = 0xFFFFFFFFFFFFFFFF


def _golden(golden_dir):
    with open(os.path.join(golden_dir, "survey_vectors.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("n,R,mask,k", [
    (1, 1, ALL, 1), (100, 4, ALL, 1), (100, 3, ALL, 1), (10, 8, ALL, 1), (2049, 2, ALL, 1),
    (1 << 20, 4, ALL, 1), ((1 << 20) + 77, 7, ALL, 1), (300000, 4, 0xFFFFFF, 1), (200000, 3, ALL, 3),
])
def test_generate_matches_oracle(n, R, mask, k):
    # bit-exact: the pcg64(rank) fill loop, mpi/mpi_lsbsort.cpp:650-656
    with lsb.DistributedSorter(n, ranks=R, key_mask=mask, and_draws=k) as s:
        s.generate()
        got = s.download()
    want = O.generate(n, R, key_mask=mask, and_draws=k)[:n]
    assert (got == want).all()


@pytest.mark.parametrize("n,R", [(0, 1), (1, 1), (1, 4), (3, 4), (7, 3), (10, 8), (100, 4), (4095, 2), (4096, 1),
                                 (4097, 2), (65536, 4), (65537, 4), (1 << 20, 4), ((1 << 21) + 12345, 8)])
@pytest.mark.parametrize("flags", [0, L.FLAG_TWO_LEVEL, L.FLAG_ONE_PASS, L.FLAG_TWO_LEVEL | L.FLAG_ONE_PASS])
def test_sort_matches_oracle(n, R, flags):
    with lsb.DistributedSorter(n, ranks=R, flags=flags) as s:
        s.generate()
        st = s.my_sort()
        got = s.download()
    want = O.sort(O.generate(n, R), n, R)
    assert st.passes == 4
    assert (got == want).all()  # bit-exact, keys and stable values


def test_config0_2pow24_r4_golden(golden_dir):
    """BASELINE configs[0]: mpirun -n 4, n = 2^24, --verify"""
    v = next(r for r in _golden(golden_dir)["sort"] if r["n"] == 1 << 24 and r["ranks"] == 4)
    n, R = 1 << 24, 4
    with lsb.DistributedSorter(n, ranks=R) as s:
        s.generate()
        g = s.download()
        assert f"{O.fnv1a64(g):016x}" == v["fnv_in"]
        s.my_sort()
        out = s.download()
        s.verify()
    assert f"{O.fnv1a64(out):016x}" == v["fnv_out"]
    assert [f"{int(out['key'][0]):016x}", int(out["val"][0])] == v["first"]
    assert [f"{int(out['key'][n - 1]):016x}", int(out["val"][n - 1])] == v["last"]
    assert (out == O.sort(g, n, R)).all()


@pytest.mark.parametrize("bits", [4, 8, 11, 13, 16])
@pytest.mark.parametrize("flags", [0, L.FLAG_TWO_LEVEL, L.FLAG_ONE_PASS])
def test_radix_widths(golden_dir, bits, flags):
    # config 5: digit width changes the pass count, never the answer
    final = _golden(golden_dir)["radix_n1048576_r4"]["final"]
    n, R = 1 << 20, 4
    with lsb.DistributedSorter(n, ranks=R, radix_bits=bits, flags=flags) as s:
        assert s.num_passes() == -(-64 // bits)
        s.generate()
        st = s.my_sort()
        out = s.download()
    assert st.passes == -(-64 // bits)
    assert f"{O.fnv1a64(out):016x}" == final


@pytest.mark.parametrize("bits", [8, 11, 16])
@pytest.mark.parametrize("flags", [0, L.FLAG_TWO_LEVEL, L.FLAG_ONE_PASS, L.FLAG_TWO_LEVEL | L.FLAG_ONE_PASS])
def test_each_pass_matches_reference_pass(golden_dir, bits, flags):
    """globalShuffle pass by pass: counts (:226-229), starts (:350,:407-413), array after (:546-575)"""
    gold = _golden(golden_dir)
    n, R = 1 << 20, 4
    a = O.generate(n, R)[:n]
    with lsb.DistributedSorter(n, ranks=R, radix_bits=bits, flags=flags) as s:
        s.generate()
        for p in range(s.num_passes()):
            want, counts, starts, _ = O.one_pass(a, n, 1, bits, p)  # one shard == one rank
            nb = 1 << s.digit_bits(p)  # the last digit of radix 11 is 9 bits wide
            assert counts[0, nb:].sum() == 0
            assert (s.histogram(p) == counts[0, :nb]).all()
            assert (s.starts(p) == starts[:nb, 0]).all()
            st = s.global_shuffle(p)
            assert st.passes == 1
            a = s.download()
            assert (a == want).all(), f"pass {p}"
            if bits == 16:
                assert f"{O.fnv1a64(a):016x}" == gold["passes_n1048576_r4_radix16"][p]["fnv_after"]


@pytest.mark.parametrize("mask,k", [(0xFFFFFF, 1), (ALL, 3), (0xFF, 1), (0, 1), (0xFFFF0000FFFF, 2)])
@pytest.mark.parametrize("flags", [0, L.FLAG_NO_SKIP, L.FLAG_TWO_LEVEL, L.FLAG_ONE_PASS, L.FLAG_ONE_PASS | L.FLAG_NO_SKIP])
def test_skewed_keys_stable(mask, k, flags):
    # config 4: single-bin digits, massive ties; stability decides the answer
    n, R = 700001, 4
    with lsb.DistributedSorter(n, ranks=R, key_mask=mask, and_draws=k, flags=flags) as s:
        s.generate()
        before = s.checksum()
        s.my_sort()
        out = s.download()
        v = s.verify()
    g = O.generate(n, R, key_mask=mask, and_draws=k)
    assert before == O.checksum(g[:n])
    assert (out == O.sort(g, n, R)).all()
    assert list(v.checksum) == before and v.order_violations == 0 and v.elements == n


def test_constant_digits_are_skipped_not_missorted():
    # 24-bit keys: sub-digits 3..7 (digits 2 and 3) are constant -> their steps are skipped, same bytes out
    n, R = 300000, 2
    g = O.generate(n, R, key_mask=0xFFFFFF)
    want = O.sort(g, n, R)
    for flags, skipped, total in ((0, 5, 8), (L.FLAG_NO_SKIP, 0, 8), (L.FLAG_ONE_PASS, 2, 4), (L.FLAG_ONE_PASS | L.FLAG_NO_SKIP, 0, 4)):
        with lsb.DistributedSorter(n, ranks=R, key_mask=0xFFFFFF, flags=flags) as s:
            s.generate()
            st = s.my_sort()
            assert st.skipped == skipped and st.subpasses == total - skipped and st.passes == 4
            assert (s.download() == want).all()
    for flags, total in ((0, 8), (L.FLAG_ONE_PASS, 4)):
        with lsb.DistributedSorter(1000, key_mask=0, flags=flags) as s:  # every digit constant: nothing to do at all
            s.generate()
            st = s.my_sort()
            assert st.subpasses == 0 and st.skipped == total
            assert (s.download() == O.generate(1000, 1, key_mask=0)[:1000]).all()


@pytest.mark.parametrize("t1,nx,lead,cfg", [(1, 2, 1, 1), (3, 3, 1, 0), (7, 4, 2, 1), (5, 6, 3, 1)])
@pytest.mark.parametrize("mask,k,bits", [(ALL, 1, 16), (0xFFFFFF, 1, 16), (ALL, 4, 16), (ALL, 1, 11), (0xF0F0F0F0F0F0F0F0, 2, 13)])
def test_onepass_many_small_supertiles(t1, nx, lead, cfg, mask, k, bits):
    """the one-pass kernel with supertiles of 1-7 tiles: scratch-ring reuse, frontier versions across
    dozens of supertiles, segments cut into sub-tiles with look-back (skew), short last supertile"""
    n, R = 250007, 3
    g = O.generate(n, R, key_mask=mask, and_draws=k)
    want = O.sort(g, n, R, bits)
    lsb.tune("op_t1", t1)
    lsb.tune("op_nx", nx)
    lsb.tune("op_lead", lead)
    lsb.tune("op_cfg", cfg)
    try:
        with lsb.DistributedSorter(n, ranks=R, radix_bits=bits, key_mask=mask, and_draws=k, flags=L.FLAG_NO_SKIP | L.FLAG_ONE_PASS) as s:
            s.generate()
            s.my_sort()
            assert (s.download() == want).all()
    finally:
        for key, val in (("op_t1", 232), ("op_nx", 6), ("op_lead", 3), ("op_cfg", 1)):
            lsb.tune(key, val)


def test_grouped_upper_bits_large_ragged_n():
    """keys whose upper 32 bits are constant over long stretches (whole warps agree on a digit), n > 1M
    and n % 32 != 0: the warp-aggregated count path and the tail chunk of the count kernel"""
    n = (1 << 20) + 300000 + 13
    rng = np.random.default_rng(11)
    a = np.zeros(n, dtype=lsb.ELT)
    hi = np.repeat(rng.integers(0, 5, n // 4096 + 1, dtype=np.uint64), 4096)[:n]
    a["key"] = (hi << np.uint64(32)) | rng.integers(0, 1 << 32, n, dtype=np.uint64)
    a["val"] = np.arange(n, dtype=np.uint64)
    want = O.stable_sort(a, n)
    for flags in (0, L.FLAG_NO_SKIP, L.FLAG_TWO_LEVEL, L.FLAG_ONE_PASS):
        with lsb.DistributedSorter(n, flags=flags) as s:
            s.upload(a)
            for p in range(4):
                c = s.histogram(p)
                assert c.sum() == n and (c == np.bincount((a["key"] >> np.uint64(16 * p)).astype(np.int64) & 0xFFFF, minlength=65536)).all()
            s.my_sort()
            assert (s.download() == want).all()


def test_uploaded_data_and_host_path():
    # caller-supplied records (the reference only sorts generated data): duplicates + arbitrary vals
    rng = np.random.default_rng(5)
    n = 250001
    a = np.zeros(n, dtype=lsb.ELT)
    a["key"] = rng.integers(0, 1 << 20, n, dtype=np.uint64) << np.uint64(13)
    a["val"] = np.arange(n, dtype=np.uint64)
    want = O.stable_sort(a, n)
    with lsb.DistributedSorter(n) as s:
        s.upload(a)
        s.my_sort()
        assert (s.download() == want).all()
        out = np.empty_like(a)
        st = s.sort_host(a, out)
        assert (out == want).all() and st.elements == n


def test_two_contexts_sort_from_two_threads():
    # bench.py's e2e.overlapped: two contexts on one GPU, each driven by its own host thread through
    # lsb_sort_host, must not share any state; different inputs, different pass shapes, several rounds
    import threading
    rng = np.random.default_rng(11)
    n = 700001
    ins, wants = [], []
    for t in range(2):
        a = np.zeros(n, dtype=lsb.ELT)
        a["key"] = rng.integers(0, 1 << 63, n, dtype=np.uint64) >> np.uint64(7 * t)
        a["val"] = np.arange(n, dtype=np.uint64) + np.uint64(t << 40)
        ins.append(a)
        wants.append(O.stable_sort(a, n))
    sorters = [lsb.DistributedSorter(n), lsb.DistributedSorter(n, flags=L.FLAG_ONE_PASS)]
    bad = []

    def run(t):
        out = np.empty_like(ins[t])
        for _ in range(6):
            out[:] = 0
            sorters[t].sort_host(ins[t], out)
            if not (out == wants[t]).all():
                bad.append(t)

    th = [threading.Thread(target=run, args=(t,)) for t in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for s in sorters:
        s.close()
    assert not bad


def test_verifier_rejects_unsorted_and_unstable():
    n = 100000
    with lsb.DistributedSorter(n) as s:
        s.generate()
        v = s.verify(raise_on_failure=False)
        assert v.order_violations > 0
        s.my_sort()
        good = s.download()
        s.verify()
        bad = good.copy()
        bad[[10, 11]] = bad[[11, 10]]
        s.upload(bad)
        assert s.verify(raise_on_failure=False).order_violations == 1
        tie = good.copy()
        tie["key"][500] = tie["key"][499]
        tie["val"][500] = tie["val"][499]  # equal (key,val): not strictly increasing
        s.upload(tie)
        with pytest.raises(lsb.LsbError):
            s.verify()


def test_many_partition_launches_recycle_tile_counters():
    # radix 2 with the virtual-rank shape: 32 passes x 8 parts = 256 launches per sort, 3 sorts on one context
    # (the per-launch tile counters are recycled in stream order once all 1024 have been used)
    n = 40000
    want = O.sort(O.generate(n, 2), n, 2, 2)
    with lsb.DistributedSorter(n, ranks=2, radix_bits=2, flags=L.FLAG_TWO_LEVEL) as s:
        for _ in range(3):
            s.generate()
            s.my_sort()
            assert (s.download() == want).all()
    want1 = O.sort(O.generate(3000, 1), 3000, 1, 1)
    with lsb.DistributedSorter(3000, ranks=1, radix_bits=1, flags=L.FLAG_TWO_LEVEL) as s:  # 64 passes x 8 parts
        for _ in range(3):
            s.generate()
            s.my_sort()
        assert (s.download() == want1).all()


@pytest.mark.parametrize("direct,chunks", [(0, 0), (0, 2), (1, 0), (1, 1), (1, 3), (1, 4)])
def test_scatter_kernel_launch_shapes(direct, chunks, tune):
    # the non-default ways partition_kernel gets its tiles: ticket counter instead of blockIdx, and the tile
    # arriving in 1, 2, 8 or 16 bulk copies (full tiles, a ragged last tile, tiles smaller than one piece)
    tune("pt_direct", direct)
    tune("pt_chunks", chunks)
    for n, ranks, radix in [(5632 * 9 + 1, 2, 16), (5632 * 3 + 353, 1, 8), (300, 1, 16), (1409, 3, 11)]:
        want = O.sort(O.generate(n, ranks), n, ranks, radix)
        with lsb.DistributedSorter(n, ranks=ranks, radix_bits=radix) as s:
            s.generate()
            s.my_sort()
            assert (s.download() == want).all(), (n, ranks, radix)


_want_cache = {}


def _want(n, ranks, radix=16, mask=ALL):
    key = (n, ranks, radix, mask)
    if key not in _want_cache:
        _want_cache[key] = O.sort(O.generate(n, ranks, key_mask=mask), n, ranks, radix)
    return _want_cache[key]


@pytest.mark.parametrize("variant,pf_tiles,direct", [(0, 0, 1), (0, 0, 0), (1, 0, 1), (1, 1, 1), (1, 3, 1), (1, 5, 1), (1, 4, 0)])
def test_scatter_kernel_variants(variant, pf_tiles, direct, tune):
    # partition_kernel with and without the L2 prefetch of a later tile, at several prefetch distances (with tickets
    # instead of blockIdx order the prefetch is off): same bytes as the oracle for one tile, a few tiles with a ragged
    # tail (the prefetch is clipped at the end of the input), hundreds of tiles (a look-back deeper than one window),
    # virtual ranks (several part-sized launches per pass) and ties in the keys
    tune("pt_variant", variant)
    tune("pt_pf_tiles", pf_tiles)
    tune("pt_direct", direct)
    for n, ranks, radix, flags in [(300, 1, 16, 0), (5632 * 5 + 17, 2, 16, 0), (5632 * 300 + 4001, 3, 16, 0),
                                   (5632 * 40 + 1, 2, 11, L.FLAG_TWO_LEVEL), (5632 * 7, 1, 8, 0)]:
        with lsb.DistributedSorter(n, ranks=ranks, radix_bits=radix, flags=flags) as s:
            s.generate()
            s.my_sort()
            assert (s.download() == _want(n, ranks, radix)).all(), (n, ranks, radix, flags)
    n = 5632 * 64 + 99
    with lsb.DistributedSorter(n, ranks=2, key_mask=0xFFF) as s:  # 4096 distinct keys: long runs of ties, stability matters
        s.generate()
        s.my_sort()
        assert (s.download() == _want(n, 2, 16, 0xFFF)).all()


def test_repeated_sorts_reuse_context():
    # look-back words are tagged by generation instead of being cleared: exercise the wrap
    n = 50000
    want = O.sort(O.generate(n, 2), n, 2)
    with lsb.DistributedSorter(n, ranks=2, radix_bits=8) as s:
        for _ in range(20):  # 160 partition launches > 127 tags
            s.generate()
            s.my_sort()
        assert (s.download() == want).all()


@pytest.mark.parametrize("n,flags", [(1 << 26, 0), (1 << 28, 0), (1 << 26, L.FLAG_ONE_PASS), (1 << 28, L.FLAG_ONE_PASS)])
def test_large_properties(n, flags):
    # sizes with no CPU oracle: strictly increasing (key,val) + same multiset == stable sort
    with lsb.DistributedSorter(n, ranks=1, flags=flags) as s:
        s.generate()
        before = s.checksum()
        assert before[3] == (n * (n - 1) // 2) % (1 << 64)
        st = s.my_sort()
        v = s.verify()
        assert list(v.checksum) == before and v.elements == n
        # idempotence: sorting sorted data changes nothing
        s.my_sort()
        assert list(s.verify().checksum) == before
    assert st.subpasses == (4 if flags & L.FLAG_ONE_PASS else 8)


def test_errors_are_reported_not_swallowed():
    with pytest.raises(lsb.LsbError):
        lsb.DistributedSorter(10, radix_bits=17)
    with pytest.raises(lsb.LsbError):
        lsb.DistributedSorter(10, world_size=2, world_rank=2)
    with lsb.DistributedSorter(10, world_size=2, world_rank=0) as s:
        with pytest.raises(lsb.LsbError):
            s.my_sort()  # communicator not initialised


def _run_driver(*args):
    import subprocess
    exe = os.path.join(os.path.dirname(lsb.library_path()), "driver", "lsbsort")
    if not os.path.exists(exe):
        pytest.skip("driver not built")
    r = subprocess.run([exe, *args], capture_output=True, text=True, timeout=600)
    return r.returncode, r.stdout


@pytest.mark.parametrize("R,n", [(1, 10), (1, 100), (4, 100), (3, 100), (4, 1 << 20)])
def test_cpp_driver_prints_what_the_reference_prints(golden_dir, R, n):
    """same knobs, same report lines (mpi/mpi_lsbsort.cpp:619-620,646,662,684,697-699,712) and the
    same A[...] lines as the unmodified reference's --print for the slots both programs show"""
    import re
    with open(os.path.join(golden_dir, "ref_print.json")) as f:
        case = next(c for c in json.load(f)["cases"] if c["ranks"] == R and c["n"] == n)
    rc, out = _run_driver("--n", str(n), "--ranks", str(R), "--print", "--verify")
    assert rc == 0, out
    lines = out.splitlines()
    assert lines[0] == f"Total number of MPI ranks: {R}" and lines[1] == f"Problem size: {n}"
    for must in ("Generating random values", "Sorting", "Verifying"):
        assert must in lines
    assert any(re.match(r"Generated random values in \S+ s$", l) for l in lines)
    assert any(re.match(rf"Sorted {n} values in \S+$", l) for l in lines)
    assert any(re.match(r"That's \S+ M elements sorted / s$", l) for l in lines)
    blocks, cur = [], None
    for l in lines:
        if l.startswith("A: displaying"):
            cur = {}
            blocks.append(cur)
        m = re.match(r"A\[(\d+)\] = \(([0-9a-f]{16}),(\d+)\)$", l)
        if m and cur is not None:
            cur[int(m.group(1))] = (m.group(2), int(m.group(3)))
    assert len(blocks) == 2 and len(blocks[0]) == min(10, n)
    for blk, ref in ((blocks[0], case["before"]), (blocks[1], case["after"])):
        refd = {i: (k, v) for i, k, v in ref}
        for i, kv in blk.items():
            assert refd[i] == kv, (i, kv, refd[i])


def test_cpp_driver_exit_code_and_knobs():
    rc, out = _run_driver("--n", "300000", "--radix", "11", "--key-mask", "0xFFFFFF", "--verify")
    assert rc == 0 and "Verifying" in out and "Verification FAILED" not in out
    rc, out = _run_driver("--n", "1000", "--no-verify")
    assert rc == 0 and "Verifying" not in out


def test_randomised_parity_sweep():
    """seeded sweep over n / streams / digit width / skew / pass shape; every case bit-exact vs the oracle"""
    rng = np.random.default_rng(20261018)
    masks = [ALL, 0xFFFFFF, 0xFFFF0000FFFF0000, 0xF0F0F0F0F0F0F0F0, 0x3FF, 1 << 63]
    for case in range(40):
        n = int(rng.integers(0, 60000)) if case % 4 else int(rng.integers(5000, 400000))
        R = int(rng.integers(1, 9))
        bits = int(rng.choice([1, 3, 5, 7, 8, 9, 10, 11, 12, 14, 15, 16])) if case % 3 == 0 else 16
        mask = masks[int(rng.integers(0, len(masks)))]
        k = int(rng.integers(1, 4))
        flags = [0, L.FLAG_TWO_LEVEL, L.FLAG_NO_SKIP, L.FLAG_TWO_LEVEL | L.FLAG_ONE_PASS, L.FLAG_ONE_PASS][case % 5]
        if bits < 4 and n > 20000:
            n = 20000  # 64 passes of a 1-bit digit: keep the oracle quick
        g = O.generate(n, R, key_mask=mask, and_draws=k)
        want = O.sort(g, n, R, bits)
        with lsb.DistributedSorter(n, ranks=R, radix_bits=bits, key_mask=mask, and_draws=k, flags=flags) as s:
            s.generate()
            s.my_sort()
            got = s.download()
        assert (got == want).all(), (case, n, R, bits, hex(mask), k, flags)
