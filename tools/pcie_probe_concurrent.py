#!/usr/bin/env python
"""Pinned-memory PCIe bandwidth with ALL GPUs of the box copying at the same time: the ceiling of the e2e leg at N GPUs.

    python tools/pcie_probe_concurrent.py [--gpus N] [--size-mib 64] [--seconds 2]

One process per GPU (each bound to its GPU's NUMA node, like bench.py's ranks), a barrier on a shared file, then every process
runs duplex copies (H2D on one stream, D2H on another) for a fixed time; the parent adds up the rates and prints one JSON object:
per-GPU and aggregate GB/s for 1, 2, 4, ... GPUs at once, per direction.  The e2e leg moves (src + response wire) H2D and
(request wire + decoded tensor) D2H per unit - about the same in both directions - so its payload ceiling is the duplex total / 2.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "min-tfs-client_b200"), os.path.join(REPO, "tests"), REPO]


def child(gpu, size, seconds, start_at):
    import bench
    from devutil import Dev
    from min_tfs_client import _native as N

    numa = bench.bind_to_gpu_numa_node(gpu)
    a, b = Dev(gpu), Dev(gpu)
    lib = a.lib
    ha, hb = N.PinnedBuffer(size), N.PinnedBuffer(size)
    da, db = a.malloc(size), b.malloc(size)
    out = {"gpu": gpu, "numa": numa}
    for mode in ("h2d", "d2h", "duplex"):
        while time.time() < start_at[mode]:
            time.sleep(0.0005)
        n = 0
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            for _ in range(8):
                if mode in ("h2d", "duplex"):
                    N.check(lib.b200tfs_memcpy_h2d(a.ctx, da, ha.ptr, size))
                if mode in ("d2h", "duplex"):
                    N.check(lib.b200tfs_memcpy_d2h(b.ctx, hb.ptr, db, size))
            a.sync()
            b.sync()
            n += 8
        dt = time.perf_counter() - t0
        out[mode] = (2 if mode == "duplex" else 1) * n * size / dt / 1e9
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--size-mib", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=1.5)
    ap.add_argument("--child", type=int, default=-1)
    ap.add_argument("--start", type=str, default="")
    args = ap.parse_args()
    if args.child >= 0:
        child(args.child, args.size_mib << 20, args.seconds, json.loads(args.start))
        return
    from min_tfs_client import _native as N

    total = args.gpus or N.device_count()
    report = {"size_mib": args.size_mib, "seconds_per_mode": args.seconds, "runs": []}
    counts = [c for c in (1, 2, 4, 8) if c <= total]
    for count in counts:
        t = time.time() + 14.0      # process start (imports, pinned allocations) takes a while: a generous common start time
        start = {"h2d": t, "d2h": t + args.seconds + 1.0, "duplex": t + 2 * (args.seconds + 1.0)}
        procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--child", str(g), "--size-mib", str(args.size_mib),
                                   "--seconds", str(args.seconds), "--start", json.dumps(start)], stdout=subprocess.PIPE, text=True)
                 for g in range(count)]
        rows = []
        for p in procs:
            so, _ = p.communicate(timeout=300)
            line = [ln for ln in so.strip().splitlines() if ln.startswith("{")]
            rows.append(json.loads(line[-1]) if line else {"error": so[-200:]})
        agg = {m: sum(r.get(m, 0.0) for r in rows) for m in ("h2d", "d2h", "duplex")}
        report["runs"].append({"gpus_at_once": count, "aggregate_GBs": agg, "e2e_payload_ceiling_GBs": agg["duplex"] / 2, "per_gpu": rows})
        print(f"{count} GPU(s) at once: H2D {agg['h2d']:.1f}  D2H {agg['d2h']:.1f}  duplex total {agg['duplex']:.1f} GB/s", file=sys.stderr, flush=True)
    print(json.dumps(report))


if __name__ == "__main__":
    main()
