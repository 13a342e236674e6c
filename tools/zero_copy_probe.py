#!/usr/bin/env python
"""Experiment: run the encode / fused-decode kernels DIRECTLY on pinned host memory (UVA), i.e. let the SMs do the PCIe
traffic, instead of H2D copy -> kernel -> D2H copy.  Prints payload GB/s for one C2 encode+decode step pipeline."""
import ctypes as C
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "min-tfs-client_b200"), os.path.join(REPO, "tests"), REPO]
from bench import response_wire_parts  # noqa: E402
from devutil import Dev  # noqa: E402
from min_tfs_client import _native as N  # noqa: E402

P = 4 << 20
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 4
devs = [Dev(0) for _ in range(2 * depth)]
lib = devs[0].lib
x = np.random.default_rng(0).standard_normal((1024, 1024), dtype=np.float32)
pre, suf = response_wire_parts(b"y", (1024, 1024), P)
resp = np.frombuffer(pre + x.tobytes() + suf, dtype=np.uint8)
slots = []
for j in range(depth):
    s = {"x": N.PinnedBuffer(P), "wire": N.PinnedBuffer(P + 4096), "resp": N.PinnedBuffer(resp.size + 256), "out": N.PinnedBuffer(P + 4096),
         "enc": devs[2 * j], "dec": devs[2 * j + 1]}
    s["x"].array[:] = x.view(np.uint8).reshape(-1)
    s["resp"].array[: resp.size] = resp
    dims = (C.c_int64 * 2)(1024, 1024)
    s["dims"] = dims
    s["t"] = (N.Tensor * 1)(N.Tensor(data=s["x"].ptr, src_dtype=1, wire_dtype=1, rank=2, flags=0, dims=dims, key=b"x", key_len=1, packed_len=0))
    s["rq"] = (N.Request * 1)(N.Request(model_name=b"default", model_name_len=7, has_version=1, order=N.ORDER_UPB, version=1, n_inputs=1, flags=0,
                                        inputs=s["t"]))
    s["off"], s["ln"] = (C.c_uint64 * 1)(), (C.c_uint64 * 1)()
    s["roff"], s["rlen"] = (C.c_uint64 * 1)(0), (C.c_uint64 * 1)(resp.size)
    slots.append(s)


def step(s):
    N.check(lib.b200tfs_encode_requests(s["enc"].ctx, 1, s["rq"], s["wire"].ptr, P + 4096, s["off"], s["ln"]))
    N.check(lib.b200tfs_decode_responses(s["dec"].ctx, s["resp"].ptr, 1, s["roff"], s["rlen"], s["out"].ptr, P + 4096 - (P + 4096) % 256))


def wait(s):
    s["enc"].sync(); s["dec"].sync()


for s in slots:
    step(s); wait(s)
    w = s["wire"].array[int(s["off"][0]): int(s["off"][0]) + int(s["ln"][0])].tobytes()
    from oracle import wire_oracle
    assert w == wire_oracle.encode_predict_request("default", 1, [("x", x)])
    assert s["out"].array[:P].tobytes() == x.tobytes()
n = 300
for k in range(2 * depth):
    step(slots[k % depth]) if k < depth else (wait(slots[k % depth]), step(slots[k % depth]))
for s in slots:
    wait(s)
t0 = time.perf_counter()
for k in range(n):
    s = slots[k % depth]
    wait(s)
    step(s)
for s in slots:
    wait(s)
t = time.perf_counter() - t0
print(f"zero-copy (kernels on pinned host memory), depth {depth}: {2 * P * n / t / 1e9:.1f} GB/s payload, {t / n * 1e6:.0f} us per step")
