"""A/B on the same box: round 1's library (tools/bin/liblsbsort_r1.so, built from commit 161055e) vs the current
one, through the entry points and struct layouts both ABIs share (raw ctypes, no binding)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_lsb_b200 import lsbsort as L  # noqa: E402  (struct layouts only)

n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 30
for name in ("tools/bin/liblsbsort_r1.so", "distributed-lsb_b200/liblsbsort.so"):
    lib = ctypes.CDLL(os.path.join(ROOT, name))
    lib.lsb_create.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(L._Config)]
    lib.lsb_sort.argtypes = [ctypes.c_void_p, ctypes.POINTER(L.Stats)]
    lib.lsb_generate.argtypes = [ctypes.c_void_p]
    lib.lsb_destroy.argtypes = [ctypes.c_void_p]
    lib.lsb_destroy.restype = None
    ctx = ctypes.c_void_p()
    cfg = L._Config(n=n, ranks=1, world_size=1, world_rank=0, device=0, radix_bits=16, and_draws=1, seed_base=0,
                    key_mask=0xFFFFFFFFFFFFFFFF, flags=1 | 8, reserved=0)
    assert lib.lsb_create(ctypes.byref(ctx), ctypes.byref(cfg)) == 0
    st = L.Stats()
    for i in range(3):
        lib.lsb_generate(ctx)
        assert lib.lsb_sort(ctx, ctypes.byref(st)) == 0
    print(f"{name}: sort {st.device_ms:.3f} ms, count {st.hist_ms:.3f} ms, scatter launch {st.partition_ms / st.partition_launches:.3f} ms "
          f"x {st.partition_launches}")
    lib.lsb_destroy(ctx)
