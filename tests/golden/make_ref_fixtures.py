#!/usr/bin/env python
"""Regenerates tests/golden/ref_print.json from the UNMODIFIED reference.

Runs oracle/_ref/mpi_lsbsort_shim (= /root/reference/mpi/mpi_lsbsort.cpp compiled over the
ranks-as-threads mpi.h shim, see oracle/build_ref.sh) with `--print --verify` and records what
the reference itself printed before and after sorting (mpi/mpi_lsbsort.cpp:171-200,669-671,
706-708).  `--verify` makes the reference compare its own output with std::stable_sort of its
input (mpi/mpi_lsbsort.cpp:710-739, asserts live), so a recorded case is one the reference
certified.  Needs /root/reference, so it only runs in the build container; the JSON it writes
is committed and is what tests/test_oracle.py reads (the GPU box has no /root/reference).

    python tests/golden/make_ref_fixtures.py
"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
BIN = os.path.join(ROOT, "oracle", "_ref", "mpi_lsbsort_shim")
LINE = re.compile(r"^A\[(\d+)\] = \(([0-9a-f]{16}),(\d+)\)$")

# (ranks, n): tiny cases print the whole array (10*ranks >= n); larger print 10 per rank
CASES = [(1, 10), (4, 40), (3, 29), (8, 10), (4, 3), (4, 1), (2, 5), (4, 100), (3, 100), (1, 100),
         (4, 65536), (4, 65537), (8, 1 << 20), (4, 1 << 20), (4, 1 << 24)]


def run(ranks, n):
    env = dict(os.environ, SHIM_RANKS=str(ranks))
    out = subprocess.run([BIN, "--n", str(n), "--print", "--verify"], env=env, check=True,
                         capture_output=True, text=True).stdout.splitlines()
    blocks, cur = [], None
    for ln in out:
        if ln.startswith("A: displaying"):
            cur = []
            blocks.append(cur)
        else:
            m = LINE.match(ln)
            if m and cur is not None:
                cur.append([int(m.group(1)), m.group(2), int(m.group(3))])
    assert len(blocks) == 2, out
    assert any(l == "Verifying" for l in out)
    return {"ranks": ranks, "n": n, "before": blocks[0], "after": blocks[1]}


def main():
    subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref.sh")])
    cases = [run(r, n) for r, n in CASES]
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_print.json")
    with open(dst, "w") as f:
        json.dump({"source": "mpi/mpi_lsbsort.cpp --print --verify via oracle/_ref (shim transport)",
                   "cases": cases}, f, separators=(",", ":"))
        f.write("\n")
    print("wrote", dst, sum(len(c["before"]) + len(c["after"]) for c in cases), "lines")


if __name__ == "__main__":
    sys.exit(main())
