"""ctypes binding of oracle/lsb_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module (see the header of lsb_oracle.c).  The product package
never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblsb_oracle.so")
REF_BIN = os.path.join(_HERE, "_ref", "mpi_lsbsort_shim")

ELT = np.dtype([("key", "<u8"), ("val", "<u8")])


def build(force=False):
    """Compile the C restatement (and oracle/_ref when /root/reference is present)."""
    src = os.path.join(_HERE, "lsb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v2", "-fPIC", "-shared", "-o", _SO, src])
    subprocess.check_call([os.path.join(_HERE, "build_ref.sh")])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        i64, u64, vp, ci = ctypes.c_int64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int
        L.lsbo_pcg64_stream.argtypes = [u64, u64, vp, i64]
        L.lsbo_pcg64_stream.restype = None
        L.lsbo_per_rank.argtypes = [i64, ci]
        L.lsbo_per_rank.restype = i64
        L.lsbo_here.argtypes = [i64, ci, ci]
        L.lsbo_here.restype = i64
        L.lsbo_generate.argtypes = [vp, i64, ci, u64, u64, ci]
        L.lsbo_generate.restype = None
        L.lsbo_num_passes.argtypes = [ci]
        L.lsbo_num_passes.restype = ci
        L.lsbo_pass.argtypes = [vp, vp, i64, ci, ci, ci, vp, vp, vp]
        L.lsbo_pass.restype = ci
        L.lsbo_sort.argtypes = [vp, vp, i64, ci, ci]
        L.lsbo_sort.restype = ci
        L.lsbo_stable_sort.argtypes = [vp, i64]
        L.lsbo_stable_sort.restype = ci
        L.lsbo_fnv1a64.argtypes = [vp, i64]
        L.lsbo_fnv1a64.restype = u64
        L.lsbo_checksum.argtypes = [vp, i64, vp]
        L.lsbo_checksum.restype = None
        L.lsbo_order_violations.argtypes = [vp, i64]
        L.lsbo_order_violations.restype = i64
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def pcg64_stream(seed, count, skip=0):
    out = np.empty(count, dtype=np.uint64)
    lib().lsbo_pcg64_stream(seed, skip, _p(out), count)
    return out


def per_rank(n, ranks):
    return lib().lsbo_per_rank(n, ranks)


def here(n, ranks, r):
    return lib().lsbo_here(n, ranks, r)


def num_passes(radix_bits):
    return lib().lsbo_num_passes(radix_bits)


def generate(n, ranks, seed_base=0, key_mask=0xFFFFFFFFFFFFFFFF, and_draws=1):
    """All ranks*per slots, as the reference fills them; [:n] is what gets sorted."""
    per = per_rank(n, ranks)
    out = np.zeros(max(per * ranks, 1), dtype=ELT)[: per * ranks]
    lib().lsbo_generate(_p(out), n, ranks, seed_base, key_mask, and_draws)
    return out


def one_pass(src, n, ranks, radix_bits, p):
    """-> (dst[:n], counts[ranks][nb], starts[nb][ranks], sendcounts[ranks][ranks])"""
    nb = 1 << radix_bits
    src = np.ascontiguousarray(src)
    dst = np.zeros(max(n, 1), dtype=ELT)[:n]
    counts = np.zeros((ranks, nb), dtype=np.int64)
    starts = np.zeros((nb, ranks), dtype=np.int64)
    sc = np.zeros((ranks, ranks), dtype=np.int64)
    rc = lib().lsbo_pass(_p(src), _p(dst), n, ranks, radix_bits, p, _p(counts), _p(starts), _p(sc))
    assert rc == 0
    return dst, counts, starts, sc


def sort(elts, n, ranks, radix_bits=16):
    a = np.array(elts[:n], dtype=ELT, copy=True)
    b = np.zeros_like(a)
    rc = lib().lsbo_sort(_p(a), _p(b), n, ranks, radix_bits)
    assert rc == 0
    return a


def stable_sort(elts, n):
    a = np.array(elts[:n], dtype=ELT, copy=True)
    rc = lib().lsbo_stable_sort(_p(a), n)
    assert rc == 0
    return a


def fnv1a64(arr):
    arr = np.ascontiguousarray(arr)
    return lib().lsbo_fnv1a64(_p(arr), arr.nbytes)


def checksum(elts):
    elts = np.ascontiguousarray(elts)
    out = np.zeros(4, dtype=np.uint64)
    lib().lsbo_checksum(_p(elts), len(elts), _p(out))
    return [int(x) for x in out]


def order_violations(elts):
    elts = np.ascontiguousarray(elts)
    return lib().lsbo_order_violations(_p(elts), len(elts))
