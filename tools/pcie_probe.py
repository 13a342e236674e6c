"""Pinned-memory PCIe bandwidth of the box (context for the e2e number): H2D, D2H, both at once."""
import ctypes as C
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "min-tfs-client_b200"), os.path.join(REPO, "tests")]
from devutil import Dev  # noqa: E402
from min_tfs_client import _native as N  # noqa: E402

a, b = Dev(0), Dev(0)
lib = a.lib
for size in (4 << 20, 64 << 20):
    ha, hb = N.PinnedBuffer(size), N.PinnedBuffer(size)
    da, db = a.malloc(size), b.malloc(size)
    reps = 50

    def run(fn):
        fn(); a.sync(); b.sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        a.sync(); b.sync()
        return (time.perf_counter() - t0) / reps

    t_h2d = run(lambda: N.check(lib.b200tfs_memcpy_h2d(a.ctx, da, ha.ptr, size)))
    t_d2h = run(lambda: N.check(lib.b200tfs_memcpy_d2h(b.ctx, hb.ptr, db, size)))
    t_both = run(lambda: (N.check(lib.b200tfs_memcpy_h2d(a.ctx, da, ha.ptr, size)), N.check(lib.b200tfs_memcpy_d2h(b.ctx, hb.ptr, db, size))))
    print(f"size {size >> 20} MiB: H2D {size / t_h2d / 1e9:.1f} GB/s, D2H {size / t_d2h / 1e9:.1f} GB/s, both at once {2 * size / t_both / 1e9:.1f} GB/s total "
          f"({t_h2d * 1e6:.0f} / {t_d2h * 1e6:.0f} / {t_both * 1e6:.0f} us)")

# pageable memory (what a numpy array or a bytes object is): the driver stages it; plus plain host memcpy for scale
import numpy as np  # noqa: E402

for size in (4 << 20, 64 << 20):
    src = np.ones(size, dtype=np.uint8)
    dst = np.empty(size, dtype=np.uint8)
    pin = N.PinnedBuffer(size)
    da = a.malloc(size)
    reps = 20

    def run1(fn):
        fn(); a.sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        a.sync()
        return (time.perf_counter() - t0) / reps

    t_h2d = run1(lambda: N.check(lib.b200tfs_memcpy_h2d(a.ctx, da, src.ctypes.data, size)))
    t_d2h = run1(lambda: N.check(lib.b200tfs_memcpy_d2h(a.ctx, dst.ctypes.data, da, size)))
    t_cp = run1(lambda: C.memmove(pin.ptr, src.ctypes.data, size))
    t_cp2 = run1(lambda: C.memmove(dst.ctypes.data, pin.ptr, size))
    t_new = run1(lambda: bytes(size))
    print(f"size {size >> 20} MiB pageable: H2D {size / t_h2d / 1e9:.1f} GB/s, D2H {size / t_d2h / 1e9:.1f} GB/s; host memcpy pageable->pinned "
          f"{size / t_cp / 1e9:.1f} GB/s, pinned->pageable {size / t_cp2 / 1e9:.1f} GB/s; fresh bytes({size >> 20} MiB) {t_new * 1e6:.0f} us")
