"""Host-side mirror of the reference driver's interface for the sort path, over the C ABI.

The reference (mpi/mpi_lsbsort.cpp) is C++ with no Python surface, so this module is a thin
ctypes binding whose names follow the reference: `DistributedSorter.create` ~
DistributedArray<SortElement>::create for A and B (:138-161,:638-639), `generate` ~ the pcg64
fill loop (:650-656), `my_sort` ~ mySort (:580-585), `global_shuffle(digit)` ~ globalShuffle
(:481-577), `verify` ~ the verify block (:710-739).  All compute happens in
liblsbsort.so (CUDA, sm_100a).  There is no CPU fallback: if the library is missing or there
is no GPU, calls raise.
"""
import ctypes
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_LIB_PATH = os.environ.get("LSB_LIBRARY", os.path.join(_HERE, "liblsbsort.so"))  # override: kernel experiments
_HEADER = os.path.join(_ROOT, "include", "lsbsort.h")

ELT = np.dtype([("key", "<u8"), ("val", "<u8")])  # SortElement, mpi/mpi_lsbsort.cpp:29-32

LSB_MAX_GPUS = 8
LSB_MAX_SUBPASSES = 32
LSB_COMM_ID_BYTES = 128
FLAG_PHASE_EVENTS = 1
FLAG_TWO_LEVEL = 2
FLAG_ONE_PASS = 4
FLAG_NO_SKIP = 8


class LsbError(RuntimeError):
    def __init__(self, code, what, detail=""):
        super().__init__(f"{what} failed: status {code}" + (f" ({detail})" if detail else ""))
        self.code = code


class _Config(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int64), ("ranks", ctypes.c_int32), ("world_size", ctypes.c_int32),
                ("world_rank", ctypes.c_int32), ("device", ctypes.c_int32), ("radix_bits", ctypes.c_int32),
                ("and_draws", ctypes.c_int32), ("seed_base", ctypes.c_uint64), ("key_mask", ctypes.c_uint64),
                ("flags", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


class Stats(ctypes.Structure):
    _fields_ = [("device_ms", ctypes.c_double), ("passes", ctypes.c_int32), ("subpasses", ctypes.c_int32), ("skipped", ctypes.c_int32), ("reserved0", ctypes.c_int32),
                ("elements", ctypes.c_int64), ("hist_ms", ctypes.c_double), ("scan_ms", ctypes.c_double),
                ("partition_ms", ctypes.c_double), ("exchange_ms", ctypes.c_double), ("subpass_ms", ctypes.c_double * LSB_MAX_SUBPASSES),
                ("sent", ctypes.c_int64 * LSB_MAX_GPUS), ("partition_launches", ctypes.c_int64), ("partition_elements", ctypes.c_int64),
                ("kernel_launches", ctypes.c_int64)]


class Verify(ctypes.Structure):
    _fields_ = [("order_violations", ctypes.c_int64), ("elements", ctypes.c_int64),
                ("checksum", ctypes.c_uint64 * 4)]


def library_path():
    return _LIB_PATH


def header_symbols():
    """names of every function include/lsbsort.h declares"""
    with open(_HEADER) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(lsb_[a-z0-9_]+)\s*\(", text)))


_lib = None


def load_library():
    """dlopen liblsbsort.so; raises (never falls back) if it has not been built"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise LsbError(-4, "load_library", f"{_LIB_PATH} not built; run python -c 'import __graft_entry__ as g; g.build()'")
    L = ctypes.CDLL(_LIB_PATH)
    vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    pvp = ctypes.POINTER(ctypes.c_void_p)
    L.lsb_abi_version.restype = ci
    L.lsb_create.argtypes = [pvp, ctypes.POINTER(_Config)]
    L.lsb_destroy.argtypes = [vp]
    L.lsb_destroy.restype = None
    L.lsb_last_error.argtypes = [vp]
    L.lsb_last_error.restype = ctypes.c_char_p
    L.lsb_status_string.argtypes = [ci]
    L.lsb_status_string.restype = ctypes.c_char_p
    L.lsb_tune.argtypes = [ctypes.c_char_p, ci]
    L.lsb_comm_unique_id.argtypes = [vp]
    L.lsb_comm_init.argtypes = [vp, vp]
    L.lsb_barrier.argtypes = [vp]
    L.lsb_shard_info.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(i64)]
    L.lsb_generate.argtypes = [vp]
    L.lsb_upload.argtypes = [vp, vp, i64, i64]
    L.lsb_download.argtypes = [vp, vp, i64, i64]
    L.lsb_device_ptr.argtypes = [vp, pvp]
    L.lsb_host_alloc.argtypes = [pvp, i64]
    L.lsb_host_free.argtypes = [vp]
    L.lsb_sort.argtypes = [vp, ctypes.POINTER(Stats)]
    L.lsb_pass.argtypes = [vp, ci, ctypes.POINTER(Stats)]
    L.lsb_sort_host.argtypes = [vp, vp, vp, i64, ctypes.POINTER(Stats)]
    L.lsb_histogram.argtypes = [vp, ci, vp]
    L.lsb_starts.argtypes = [vp, ci, vp]
    L.lsb_num_passes.argtypes = [vp]
    L.lsb_digit_bits.argtypes = [vp, ci]
    L.lsb_checksum.argtypes = [vp, vp]
    L.lsb_verify_device.argtypes = [vp, ctypes.POINTER(Verify)]
    for name in header_symbols():
        getattr(L, name)  # every declared entry point must be exported
    _lib = L
    return L


def abi_symbols():
    L = load_library()
    return [s for s in header_symbols() if hasattr(L, s)]


def tune(key, value):
    """experiment knob for profiling sweeps (lsb_tune): applies to sorters created afterwards"""
    rc = load_library().lsb_tune(key.encode(), int(value))
    if rc:
        raise LsbError(rc, "lsb_tune", load_library().lsb_last_error(None).decode())


def comm_unique_id():
    """MPI_Init's rendezvous token: rank 0 makes it, every rank passes it to comm_init"""
    buf = ctypes.create_string_buffer(LSB_COMM_ID_BYTES)
    rc = load_library().lsb_comm_unique_id(buf)
    if rc:
        raise LsbError(rc, "lsb_comm_unique_id", load_library().lsb_last_error(None).decode())
    return buf.raw


class PinnedBuffer:
    """page-locked host array of ELT (cudaHostAlloc through the C ABI)"""

    def __init__(self, count):
        self._ptr = ctypes.c_void_p()
        rc = load_library().lsb_host_alloc(ctypes.byref(self._ptr), max(count, 1) * ELT.itemsize)
        if rc:
            raise LsbError(rc, "lsb_host_alloc")
        raw = (ctypes.c_char * (max(count, 1) * ELT.itemsize)).from_address(self._ptr.value)
        self.array = np.frombuffer(raw, dtype=ELT)[:count]

    def free(self):
        if self._ptr:
            self.array = None
            load_library().lsb_host_free(self._ptr)
            self._ptr = None


class DistributedSorter:
    """One process's view of the distributed arrays A and B (one shard of each on one GPU)."""

    def __init__(self, n, ranks=0, world_size=1, world_rank=0, device=0, radix_bits=16, seed_base=0,
                 key_mask=0xFFFFFFFFFFFFFFFF, and_draws=1, flags=0):
        self._L = load_library()
        self._ctx = ctypes.c_void_p()
        cfg = _Config(n=n, ranks=ranks, world_size=world_size, world_rank=world_rank, device=device,
                      radix_bits=radix_bits, and_draws=and_draws, seed_base=seed_base, key_mask=key_mask,
                      flags=flags, reserved=0)
        rc = self._L.lsb_create(ctypes.byref(self._ctx), ctypes.byref(cfg))
        if rc:
            raise LsbError(rc, "lsb_create", self._L.lsb_last_error(None).decode())
        per, here, first = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        self._L.lsb_shard_info(self._ctx, ctypes.byref(per), ctypes.byref(here), ctypes.byref(first))
        self.n, self.world_size, self.world_rank = n, world_size, world_rank
        self.per, self.here, self.first_global = per.value, here.value, first.value
        self.radix_bits = radix_bits

    create = classmethod(lambda cls, *a, **k: cls(*a, **k))

    # -- plumbing -----------------------------------------------------------------------
    def _check(self, rc, what):
        if rc:
            raise LsbError(rc, what, self._L.lsb_last_error(self._ctx).decode())

    def close(self):
        if self._ctx:
            self._L.lsb_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def comm_init(self, unique_id):
            self._check(self._L.lsb_comm_init(self._ctx, unique_id), "lsb_comm_init")

    def barrier(self):
        """MPI_Barrier(MPI_COMM_WORLD): all GPUs drained their sort streams"""
        self._check(self._L.lsb_barrier(self._ctx), "lsb_barrier")

    # -- data ---------------------------------------------------------------------------
    def generate(self):
        self._check(self._L.lsb_generate(self._ctx), "lsb_generate")

    def upload(self, elts, local_off=0):
        elts = np.ascontiguousarray(elts, dtype=ELT)
        self._check(self._L.lsb_upload(self._ctx, elts.ctypes.data, local_off, len(elts)), "lsb_upload")

    def download(self, local_off=0, count=None, out=None):
        count = self.here - local_off if count is None else count
        if out is None:
            out = np.empty(count, dtype=ELT)
        self._check(self._L.lsb_download(self._ctx, out.ctypes.data, local_off, count), "lsb_download")
        return out

    def device_ptr(self):
        p = ctypes.c_void_p()
        self._check(self._L.lsb_device_ptr(self._ctx, ctypes.byref(p)), "lsb_device_ptr")
        return p.value

    # -- the hot path ---------------------------------------------------------------------
    def my_sort(self):
        """mySort(A, B): all passes; returns Stats"""
        st = Stats()
        self._check(self._L.lsb_sort(self._ctx, ctypes.byref(st)), "lsb_sort")
        return st

    def global_shuffle(self, digit):
        """globalShuffle(A, B, digit): one stable pass"""
        st = Stats()
        self._check(self._L.lsb_pass(self._ctx, digit, ctypes.byref(st)), "lsb_pass")
        return st

    def sort_host(self, host_in, host_out):
        """host buffers in, host buffers out (copies inside the call)"""
        st = Stats()
        assert host_in.dtype == ELT and host_out.dtype == ELT and len(host_in) == len(host_out)
        self._check(self._L.lsb_sort_host(self._ctx, host_in.ctypes.data, host_out.ctypes.data, len(host_in),
                                          ctypes.byref(st)), "lsb_sort_host")
        return st

    # -- test hooks -----------------------------------------------------------------------
    def num_passes(self):
        return self._L.lsb_num_passes(self._ctx)

    def digit_bits(self, digit):
        return self._L.lsb_digit_bits(self._ctx, digit)

    def histogram(self, digit):
        out = np.zeros(1 << self.digit_bits(digit), dtype=np.int64)
        self._check(self._L.lsb_histogram(self._ctx, digit, out.ctypes.data), "lsb_histogram")
        return out

    def starts(self, digit):
        out = np.zeros(1 << self.digit_bits(digit), dtype=np.int64)
        self._check(self._L.lsb_starts(self._ctx, digit, out.ctypes.data), "lsb_starts")
        return out

    # -- verification ---------------------------------------------------------------------
    def checksum(self):
        out = (ctypes.c_uint64 * 4)()
        self._check(self._L.lsb_checksum(self._ctx, out), "lsb_checksum")
        return [int(x) for x in out]

    def verify(self, raise_on_failure=True):
        v = Verify()
        rc = self._L.lsb_verify_device(self._ctx, ctypes.byref(v))
        if rc and (raise_on_failure or rc != -6):
            self._check(rc, "lsb_verify_device")
        return v
