#!/usr/bin/env python
"""One host-buffer encode / decode call at a time (pinned buffers): wall time per call and the part of it spent inside the call
(enqueue), for the monolithic path and 2..8 slices.  `python tools/pipeline_probe.py [--mib 4]`; one process per setting, since
the slice count is read when a context is created."""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "min-tfs-client_b200"), os.path.join(REPO, "tests"), REPO]


def child(mib, reps):
    import numpy as np
    from devutil import Dev
    from min_tfs_client import _native as N
    from oracle import wire_oracle

    dev = Dev(0)
    lib = dev.lib
    elems = mib << 18
    x = np.random.default_rng(0).standard_normal(elems, dtype=np.float32)
    xin = N.PinnedBuffer(x.nbytes)
    xin.array[:] = x.view(np.uint8)
    dims = (C.c_int64 * 1)(elems)
    t = (N.Tensor * 1)(N.Tensor(data=xin.ptr, src_dtype=1, wire_dtype=1, rank=1, flags=0, dims=dims, key=b"x", key_len=1, packed_len=0))
    rq = (N.Request * 1)(N.Request(model_name=b"default", model_name_len=7, has_version=1, order=N.ORDER_UPB, version=1, n_inputs=1, flags=0, inputs=t))
    cap = x.nbytes + 4096
    wire = N.PinnedBuffer(cap)
    off, ln = (C.c_uint64 * 1)(), (C.c_uint64 * 1)()
    resp = wire_oracle.build_predict_response([("y", x)])
    rbuf = N.PinnedBuffer(len(resp) + 256)
    rbuf.array[: len(resp)] = np.frombuffer(resp, dtype=np.uint8)
    stride = (len(resp) + 2303 + 255) & ~255
    obuf = N.PinnedBuffer(stride)
    roff, rln = (C.c_uint64 * 1)(0), (C.c_uint64 * 1)(len(resp))
    st = (C.c_int32 * 1)()
    out = {}

    def enc():
        N.check(lib.b200tfs_encode_requests_host_async(dev.ctx, 1, rq, wire.ptr, cap, off, ln))

    def dec():
        N.check(lib.b200tfs_decode_responses_host_async(dev.ctx, rbuf.ptr, 1, roff, rln, obuf.ptr, stride))

    def enc_zero_copy():     # the kernels work on the pinned buffers themselves (unified addressing): SM loads / stores cross PCIe
        N.check(lib.b200tfs_encode_requests(dev.ctx, 1, rq, wire.ptr, cap, off, ln))

    def dec_zero_copy():
        N.check(lib.b200tfs_decode_responses(dev.ctx, rbuf.ptr, 1, roff, rln, obuf.ptr, stride))

    # device-resident input, output written by the kernel straight into pinned host memory (posted PCIe writes, no D2H copy)
    xd = dev.upload(x)
    td = (N.Tensor * 1)(N.Tensor(data=xd, src_dtype=1, wire_dtype=1, rank=1, flags=0, dims=dims, key=b"x", key_len=1, packed_len=0))
    rqd = (N.Request * 1)(N.Request(model_name=b"default", model_name_len=7, has_version=1, order=N.ORDER_UPB, version=1, n_inputs=1, flags=0, inputs=td))
    rd = dev.upload(np.frombuffer(resp, dtype=np.uint8))

    def enc_host_out():
        N.check(lib.b200tfs_encode_requests(dev.ctx, 1, rqd, wire.ptr, cap, off, ln))

    def dec_host_out():
        N.check(lib.b200tfs_decode_responses(dev.ctx, rd, 1, roff, rln, obuf.ptr, stride))

    for name, fn in (("encode", enc), ("decode", dec), ("encode_zero_copy", enc_zero_copy), ("decode_zero_copy", dec_zero_copy),
                     ("encode_device_in_host_out", enc_host_out), ("decode_device_in_host_out", dec_host_out)):
        for _ in range(5):
            fn()
            dev.sync()
        t_call = t_all = 0.0
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            t1 = time.perf_counter()
            dev.sync()
            t2 = time.perf_counter()
            t_call += t1 - t0
            t_all += t2 - t0
        out[name] = {"us_per_call": t_all / reps * 1e6, "us_inside_the_call": t_call / reps * 1e6}
    assert wire.array[off[0]: off[0] + ln[0]].tobytes() == wire_oracle.encode_predict_request("default", 1, [("x", x)])
    n = C.c_uint64()
    N.check(lib.b200tfs_pipelined_calls(dev.ctx, C.byref(n)))
    out["sliced_calls"] = n.value
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=4)
    ap.add_argument("--reps", type=int, default=100)
    ap.add_argument("--child", action="store_true")
    a = ap.parse_args()
    if a.child:
        return child(a.mib, a.reps)
    rows = {}
    for label, env in (("monolithic", {"B200TFS_PIPELINE_MIN": "0"}), ("2 slices", {"B200TFS_PIPELINE_SLICES": "2"}),
                       ("3 slices", {"B200TFS_PIPELINE_SLICES": "3"}), ("4 slices", {"B200TFS_PIPELINE_SLICES": "4"}),
                       ("8 slices", {"B200TFS_PIPELINE_SLICES": "8"})):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--mib", str(a.mib), "--reps", str(a.reps)],
                           env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        rows[label] = json.loads(line[-1]) if line else {"error": (r.stdout + r.stderr)[-300:]}
        print(label, rows[label], file=sys.stderr, flush=True)
    print(json.dumps({"mib": a.mib, "rows": rows}))


if __name__ == "__main__":
    main()
