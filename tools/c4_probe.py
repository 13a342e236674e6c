import ctypes as C, sys, os, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "min-tfs-client_b200"), os.path.join(REPO, "tests"), REPO]
import numpy as np
from devutil import Dev
from min_tfs_client import _native as N
from oracle import wire_oracle
dev = Dev(0); lib = dev.lib
n = 32
x = np.random.default_rng(0).standard_normal((8, 512, 1024)).astype(np.float16).astype(np.float32)
w = wire_oracle.build_predict_response([("y", x)])
stride = (len(w) + 255) & ~255
buf = np.zeros(stride * n + 256, np.uint8)
for i in range(n): buf[i * stride: i * stride + len(w)] = np.frombuffer(w, np.uint8)
wd = dev.upload(buf)
off = (C.c_uint64 * n)(*[i * stride for i in range(n)]); ln = (C.c_uint64 * n)(*[len(w)] * n)
for cast, dstride in ((0, x.nbytes), (19, x.nbytes // 2)):
    dst = dev.malloc(dstride * n + 256)
    N.check(lib.b200tfs_set_decode_cast(dev.ctx, cast))
    e0, e1 = C.c_void_p(), C.c_void_p(); lib.b200tfs_event_create(C.byref(e0)); lib.b200tfs_event_create(C.byref(e1))
    def stats():
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64(); N.check(lib.b200tfs_decode_stats(dev.ctx, C.byref(a), C.byref(b), C.byref(c))); return a.value, b.value, c.value
    for rep in range(4):
        s0 = stats()
        lib.b200tfs_event_record(dev.ctx, e0)
        N.check(lib.b200tfs_decode_responses(dev.ctx, wd, n, off, ln, dst, dstride))
        lib.b200tfs_event_record(dev.ctx, e1); dev.sync()
        ms = C.c_float(); lib.b200tfs_event_elapsed_ms(e0, e1, C.byref(ms))
        s1 = stats()
        print("cast", cast, "rep", rep, "us", round(ms.value * 1e3, 1), "served (param, device tpl, walk)", tuple(b - a for a, b in zip(s0, s1)))
    got = dev.download(dst, dstride)
    want = x.astype(np.float16) if cast else x
    print("  ok" if got[: want.nbytes].tobytes() == want.tobytes() else "  MISMATCH")

# the two-phase route on the same batch: parse once (table), then time the unpack kernel alone
N.check(lib.b200tfs_set_decode_cast(dev.ctx, 0))
outs = (N.Output * n)(); n_outs = (C.c_int32 * n)(); specs = (N.ModelSpec * n)(); status = (C.c_int32 * n)()
N.check(lib.b200tfs_parse_responses(dev.ctx, wd, n, off, ln, 1, outs, n_outs, specs, status))
dstride = x.nbytes // 2
dst = dev.malloc(dstride * n + 256)
dptr = (C.c_void_p * n)(*[dst + j * dstride for j in range(n)]); dcode = (C.c_int32 * n)(*[19] * n)
for rep in range(4):
    lib.b200tfs_event_record(dev.ctx, e0)
    N.check(lib.b200tfs_unpack_outputs(dev.ctx, wd, n, outs, off, dptr, dcode, None))
    lib.b200tfs_event_record(dev.ctx, e1); dev.sync()
    ms = C.c_float(); lib.b200tfs_event_elapsed_ms(e0, e1, C.byref(ms))
    print("two-phase unpack (move_kernel OP_F2H) rep", rep, "us", round(ms.value * 1e3, 1))
got = dev.download(dst, dstride)
print("  ok" if got[: x.nbytes // 2].tobytes() == x.astype(np.float16).tobytes() else "  MISMATCH")
for tb in (0,):
    pass
