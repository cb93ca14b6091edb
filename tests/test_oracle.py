"""Pins the oracle (oracle/lsb_oracle.c) before anything trusts it.

Three independent anchors:
  1. what the UNMODIFIED reference printed (tests/golden/ref_print.json, produced by
     tests/golden/make_ref_fixtures.py from /root/reference/mpi/mpi_lsbsort.cpp --print --verify),
  2. the known-answer vectors of SURVEY.md section 8(c) (tests/golden/survey_vectors.json),
  3. the definition of the result (std::stable_sort by key, mpi/mpi_lsbsort.cpp:722-726),
     restated as a merge sort that shares no code with the radix passes.
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O


def _load(golden_dir, name):
    with open(os.path.join(golden_dir, name)) as f:
        return json.load(f)


def test_pcg64_known_answers(golden_dir):
    kat = _load(golden_dir, "survey_vectors.json")["pcg64"]
    for seed, vals in kat.items():
        got = O.pcg64_stream(int(seed), len(vals))
        assert [f"{int(x):016x}" for x in got] == vals


def test_pcg64_jump_ahead_matches_sequential():
    seq = O.pcg64_stream(7, 5000)
    for skip in (0, 1, 2, 3, 255, 256, 1023, 4097):
        got = O.pcg64_stream(7, 64, skip=skip)
        assert (got == seq[skip:skip + 64]).all()


def test_reference_printed_lines(golden_dir):
    """generator and sorted output equal what the reference itself printed"""
    cases = _load(golden_dir, "ref_print.json")["cases"]
    assert len(cases) >= 10
    for c in cases:
        n, R = c["n"], c["ranks"]
        g = O.generate(n, R)
        per = O.per_rank(n, R)
        for idx, key, val in c["before"]:
            assert idx < R * per
            assert f"{int(g['key'][idx]):016x}" == key and int(g["val"][idx]) == val, (n, R, idx)
        s = O.sort(g, n, R)
        for idx, key, val in c["after"]:
            if idx < n:  # padding slots of the last rank are printed by the reference but not sorted
                assert f"{int(s['key'][idx]):016x}" == key and int(s["val"][idx]) == val, (n, R, idx)
        if 10 * R >= n:  # whole array was printed
            shown = [i for i, _, _ in c["after"] if i < n]
            assert shown == list(range(n))


def test_distribution_edge_cases():
    # mpi/mpi_lsbsort.cpp:144-149: per = ceil(n/R); here clamps to >= 0
    assert O.per_rank(100, 3) == 34 and [O.here(100, 3, r) for r in range(3)] == [34, 34, 32]
    assert O.per_rank(1, 4) == 1 and [O.here(1, 4, r) for r in range(4)] == [1, 0, 0, 0]
    assert O.per_rank(10, 8) == 2 and [O.here(10, 8, r) for r in range(8)] == [2, 2, 2, 2, 2, 0, 0, 0]
    assert O.per_rank(7, 3) == 3 and [O.here(7, 3, r) for r in range(3)] == [3, 3, 1]


@pytest.mark.parametrize("row", range(9))
def test_survey_sort_vectors(golden_dir, row):
    v = _load(golden_dir, "survey_vectors.json")["sort"][row]
    n, R = v["n"], v["ranks"]
    assert O.per_rank(n, R) == v["per"]
    g = O.generate(n, R)
    assert f"{O.fnv1a64(g[:n]):016x}" == v["fnv_in"]
    assert f"{int(np.bitwise_xor.reduce(g['key'][:n])):016x}" == v["xor_keys"]
    s = O.sort(g, n, R)
    assert f"{O.fnv1a64(s):016x}" == v["fnv_out"]
    for name, i in (("first", 0), ("mid", n // 2), ("last", n - 1)):
        assert [f"{int(s['key'][i]):016x}", int(s["val"][i])] == v[name]
    assert int(s["val"].astype(np.uint64).sum()) == n * (n - 1) // 2
    assert O.order_violations(s) == 0


def test_per_pass_goldens(golden_dir):
    sv = _load(golden_dir, "survey_vectors.json")
    n, R = 1 << 20, 4
    a = O.generate(n, R)[:n]
    for row in sv["passes_n1048576_r4_radix16"]:
        a, counts, starts, sc = O.one_pass(a, n, R, 16, row["pass"])
        assert f"{O.fnv1a64(counts):016x}" == row["fnv_counts"]
        assert f"{O.fnv1a64(starts):016x}" == row["fnv_starts"]
        assert f"{O.fnv1a64(a):016x}" == row["fnv_after"]
        assert sc.tolist() == row["sendcounts"]
        if row["pass"] == 0:
            spot = sv["spot_n1048576_r4_radix16_pass0"]
            assert counts[0, :4].tolist() == spot["counts_r0_d0_3"]
            assert starts[1, :].tolist() == spot["starts_d1_r0_3"]


def test_tiny_per_pass(golden_dir):
    t = _load(golden_dir, "survey_vectors.json")["tiny_n100_r4_radix16"]
    n, R = 100, 4
    a = O.generate(n, R)[:n]
    for p in range(4):
        a, _, _, sc = O.one_pass(a, n, R, 16, p)
        if p == 0:
            assert sc.tolist() == t["pass0_sendcounts"]
        assert f"{O.fnv1a64(a):016x}" == t["fnv_after"][p]


def test_radix_width_independence(golden_dir):
    r = _load(golden_dir, "survey_vectors.json")["radix_n1048576_r4"]
    n, R = 1 << 20, 4
    g = O.generate(n, R)[:n]
    a, counts, starts, _ = O.one_pass(g, n, R, 8, 0)
    assert f"{O.fnv1a64(counts):016x}" == r["radix8_pass0"]["fnv_counts"]
    assert f"{O.fnv1a64(starts):016x}" == r["radix8_pass0"]["fnv_starts"]
    assert counts[0, :4].tolist() == r["radix8_pass0"]["counts_r0_d0_3"]
    a, _, _, _ = O.one_pass(a, n, R, 8, 1)
    assert f"{O.fnv1a64(a):016x}" == r["radix8_after_pass1_equals_radix16_after_pass0"]
    _, counts, starts, _ = O.one_pass(g, n, R, 11, 0)
    assert f"{O.fnv1a64(counts):016x}" == r["radix11_pass0"]["fnv_counts"]
    assert f"{O.fnv1a64(starts):016x}" == r["radix11_pass0"]["fnv_starts"]
    assert O.num_passes(11) == r["radix11_passes"] and O.num_passes(8) == 8 and O.num_passes(16) == 4
    for bits in (8, 11, 16):
        assert f"{O.fnv1a64(O.sort(g, n, R, bits)):016x}" == r["final"]


@pytest.mark.parametrize("n,R", [(0, 1), (1, 1), (1, 4), (3, 4), (7, 3), (10, 8), (1000, 7), (65537, 4), (200001, 5)])
def test_radix_equals_definitional_stable_sort(n, R):
    g = O.generate(n, R)
    assert (O.sort(g, n, R) == O.stable_sort(g, n)).all()


@pytest.mark.parametrize("mask,k", [(0xFFFFFF, 1), (0xFFFFFFFFFFFFFFFF, 3), (0xFF, 1), (0, 1)])
def test_skewed_keys_stay_stable(mask, k):
    # config 4: heavy ties; stability is what makes the answer unique
    n, R = 50000, 4
    g = O.generate(n, R, key_mask=mask, and_draws=k)
    s = O.sort(g, n, R)
    assert (s == O.stable_sort(g, n)).all()
    assert O.order_violations(s) == 0
    assert O.checksum(s) == O.checksum(g[:n])
    if mask == 0xFFFFFFFFFFFFFFFF and k == 3:
        draws = O.pcg64_stream(0, 9)
        assert int(g["key"][1]) == int(draws[3] & draws[4] & draws[5])


def test_reference_binary_runs_config1_sample():
    """the unmodified reference passes its own --verify here (skipped if not built)"""
    import subprocess
    if not os.path.exists(O.REF_BIN):
        pytest.skip("oracle/_ref not built")
    env = dict(os.environ, SHIM_RANKS="4")
    out = subprocess.run([O.REF_BIN, "--n", "100000", "--verify"], env=env, check=True,
                         capture_output=True, text=True).stdout
    assert "Verifying" in out and "M elements sorted / s" in out
