#!/usr/bin/env python
"""bench.py - TensorProto encode+decode throughput on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c3|c5]

Workload (N=1 default, the config the metric is quoted on): **C2** - one fp32 [1024,1024] tensor per
step: encode it into a PredictRequest{model_spec{"default",1}, inputs{"x"}} wire buffer AND decode one
PredictResponse{outputs{"y": fp32[1024,1024]}, model_spec} back to a tensor.  A *step* = that one
encode + one decode.  ``value`` = tensor payload bytes processed per second (2 x 4 MiB per step),
inputs already resident in HBM; ``e2e`` = the same through the host-buffer C-ABI entry points, with the
H2D / D2H copies inside the timed region.  Between steps the buffers rotate through a ring whose
footprint exceeds the 126 MB L2 (stated in ``config``), so every step reads HBM.

Roofline: the dominant kernel is ``move_kernel`` (one launch per encode, one per decode-unpack);
algorithmic bytes per launch = 2P + H (read P, write P+H on encode; read P+H, write P on decode;
P = 4 194 304, H = 47 / 58) - DESIGN.md "Roofline".  Its average launch duration is measured live with
CUDA events on the launching stream over back-to-back launches of that kernel alone.

Multi-GPU (torchrun, one rank per GPU): independent requests shard across ranks with no data-path
collective; each rank runs the same per-GPU workload (weak scaling); the only communication is the
barrier + MAX-reduce of the device-timed region.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(REPO, "min-tfs-client_b200"), REPO]

L2_BYTES = 126 * 1024 * 1024


# ------------------------------------------------------------------------------------------------
# small wire helpers (bench-local; used to fabricate the response the decode leg consumes and to
# check results - the oracle is only touched by the cpu_baseline / --impl reference legs)
# ------------------------------------------------------------------------------------------------
def _uv(x):
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        out.append(b | (0x80 if x else 0))
        if not x:
            return bytes(out)


def _ld(tag, body):
    return bytes([tag]) + _uv(len(body)) + body


def _f32_tensor_header(shape, nbytes):
    dims = b"".join(_ld(0x12, b"\x08" + _uv(d)) for d in shape)
    return b"\x08\x01" + _ld(0x12, dims) + b"\x2a" + _uv(nbytes)


def response_wire_parts(key, shape, nbytes, model=b"default", version=1, sig=b"serving_default"):
    """(prefix, suffix) such that prefix + payload + suffix is a canonical PredictResponse."""
    th = _f32_tensor_header(shape, nbytes)
    tp_len = len(th) + nbytes
    entry_len = 1 + len(_uv(len(key))) + len(key) + 1 + len(_uv(tp_len)) + tp_len
    prefix = b"\x0a" + _uv(entry_len) + _ld(0x0A, key) + b"\x12" + _uv(tp_len) + th
    spec = _ld(0x0A, model) + _ld(0x12, b"\x08" + _uv(version)) + _ld(0x1A, sig)
    return prefix, _ld(0x12, spec)


def request_wire_parts(key, shape, nbytes, model=b"default", version=1):
    th = _f32_tensor_header(shape, nbytes)
    tp_len = len(th) + nbytes
    entry_len = 1 + len(_uv(len(key))) + len(key) + 1 + len(_uv(tp_len)) + tp_len
    spec = _ld(0x0A, model) + _ld(0x12, b"\x08" + _uv(version))
    return _ld(0x0A, spec) + b"\x12" + _uv(entry_len) + _ld(0x0A, key) + b"\x12" + _uv(tp_len) + th


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = str(gpu_index)
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 9 and parts[0] == self.gpu:
                self.rows.append((time.time(), parts))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [p for (t, p) in self.rows if t0 <= t <= t1 + 0.1] or [p for (_, p) in self.rows[-3:]]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no sample"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows),
                "samples": len(rows), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# distributed plumbing (control plane only)
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(gpu_index):
    """Run this process on the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers of the e2e leg are
    first-touched there and the PCIe copies do not cross the socket interconnect (what `numactl --cpunodebind` would do
    per rank).  Returns a description for the bench line; does nothing when the topology cannot be read
    (B200TFS_BENCH_NUMA=0 turns it off)."""
    if os.environ.get("B200TFS_BENCH_NUMA", "1") == "0":
        return "off"
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(gpu_index)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        bdf = out[-12:] if len(out) >= 12 else out            # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return "unknown node"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "no cpus allowed on node"
        os.sched_setaffinity(0, cpus)
        return f"node {node} ({len(cpus)} cpus)"
    except Exception as e:  # noqa: BLE001 - best effort
        return f"unavailable ({type(e).__name__})"


class World:
    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.size = int(os.environ.get("WORLD_SIZE", "1"))
        self.numa = bind_to_gpu_numa_node(self.local_rank)
        self.dist = None
        self.torch = None
        if self.size > 1:
            import torch
            import torch.distributed as dist

            self.torch = torch
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max(self, x):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# the GPU arm
# ------------------------------------------------------------------------------------------------
class Lane:
    """One native context (= one CUDA stream).  All lanes share one ring of device-resident C2 buffers."""

    def __init__(self, device, shared):
        from min_tfs_client import _native as N

        self.N, self.lib, self.S = N, N.load(), shared
        ctx = C.c_void_p()
        N.check(self.lib.b200tfs_create(device, C.byref(ctx)))
        self.ctx = ctx
        S = shared
        self.dims = (C.c_int64 * 2)(*S.SHAPE)
        self.tensors = (N.Tensor * 1)(N.Tensor(data=256, src_dtype=1, wire_dtype=1, rank=2, flags=0, dims=self.dims, key=b"x", key_len=1,
                                               packed_len=0))
        self.requests = (N.Request * 1)(N.Request(model_name=b"default", model_name_len=7, has_version=1, order=N.ORDER_UPB, version=1,
                                                  n_inputs=1, flags=0, inputs=self.tensors))
        need = C.c_uint64(0)
        N.check(self.lib.b200tfs_request_arena_size(1, self.requests, C.byref(need)))
        self.arena_cap = int(need.value)
        self.rec_off, self.rec_len = (C.c_uint64 * 1)(), (C.c_uint64 * 1)()
        self.p_off, self.p_len = (C.c_uint64 * 1)(0), (C.c_uint64 * 1)(S.resp_len)
        self.graphs = {}
        self.ring = None

    def malloc(self, nbytes):
        p = C.c_void_p()
        self.N.check(self.lib.b200tfs_malloc(self.ctx, nbytes, C.byref(p)))
        return p.value

    def sync(self):
        self.N.check(self.lib.b200tfs_sync(self.ctx))

    def encode(self, i):
        R = self.ring
        self.tensors[0].data = R.src[i]
        self.N.check(self.lib.b200tfs_encode_requests(self.ctx, 1, self.requests, R.arena[i], self.arena_cap, self.rec_off, self.rec_len))

    def decode(self, i):
        R = self.ring
        self.N.check(self.lib.b200tfs_decode_responses(self.ctx, R.resp[i], 1, self.p_off, self.p_len, R.dst[i], self.S.P))

    def capture(self, name, body, slots):
        """Record body(slot) for every slot of `slots`, in order, into one CUDA graph."""
        lib, N = self.lib, self.N
        for i in slots[:4]:
            body(i)          # warm: sizes every scratch buffer outside the capture
        self.sync()
        N.check(lib.b200tfs_capture_begin(self.ctx))
        for i in slots:
            body(i)
        g = C.c_void_p()
        N.check(lib.b200tfs_capture_end(self.ctx, C.byref(g)))
        old = self.graphs.get(name)
        if old:
            lib.b200tfs_graph_destroy(old[0])
        self.graphs[name] = (g, len(slots))
        return g

    def launch(self, name):
        self.N.check(self.lib.b200tfs_graph_launch(self.ctx, self.graphs[name][0]))

    def launches(self):
        n = C.c_uint64(0)
        self.N.check(self.lib.b200tfs_kernel_launches(self.ctx, C.byref(n)))
        return int(n.value)


class Ring:
    """`slots` sets of {tensor, request arena, response wire, decoded tensor} resident in HBM."""

    def __init__(self, lane, slots, shared):
        S, N, lib = shared, lane.N, lane.lib
        self.slots = slots
        self.src, self.arena, self.resp, self.dst = [], [], [], []
        for i in range(slots):
            self.src.append(lane.malloc(S.P))
            self.arena.append(lane.malloc(lane.arena_cap))
            self.resp.append(lane.malloc(S.resp_len + 256))
            self.dst.append(lane.malloc(S.P))
            k = i % len(S.host_x)
            N.check(lib.b200tfs_memcpy_h2d(lane.ctx, self.src[i], S.host_x[k].ctypes.data, S.P))
            N.check(lib.b200tfs_memcpy_h2d(lane.ctx, self.resp[i], S.resp_host[k].ctypes.data, S.resp_len))
            N.check(lib.b200tfs_memset(lane.ctx, self.arena[i], 0, lane.arena_cap))
            N.check(lib.b200tfs_memset(lane.ctx, self.dst[i], 0, S.P))
        lane.sync()
        self.bytes = slots * (S.P * 2 + lane.arena_cap + S.resp_len)


class Shared:
    """Host-side constants of the C2 workload."""

    SHAPE = (1024, 1024)

    def __init__(self):
        self.P = int(np.prod(self.SHAPE)) * 4
        self.host_x = [np.random.default_rng(i).standard_normal(self.SHAPE, dtype=np.float32) for i in range(4)]
        self.resp_prefix, self.resp_suffix = response_wire_parts(b"y", self.SHAPE, self.P)
        self.req_header = request_wire_parts(b"x", self.SHAPE, self.P)
        self.resp_len = len(self.resp_prefix) + self.P + len(self.resp_suffix)
        self.H_req, self.H_resp = len(self.req_header), len(self.resp_prefix) + len(self.resp_suffix)
        self.resp_host = [np.frombuffer(self.resp_prefix + x.tobytes() + self.resp_suffix, dtype=np.uint8) for x in self.host_x]


class C2Bench:
    """fp32 [1024,1024]: `streams` lanes working through one ring of `ring` buffer sets (lane s takes slots s, s+streams, ...)."""

    def __init__(self, device, ring, streams):
        from min_tfs_client import _native as N

        self.N, self.lib = N, N.load()
        self.S = Shared()
        self.P = self.S.P
        self.lanes = [Lane(device, self.S) for _ in range(streams)]
        self.main = self.lanes[0]
        self.ring = Ring(self.main, max(ring, streams), self.S)
        for l in self.lanes:
            l.ring = self.ring
        # pinned host buffers for the e2e leg: `depth` requests in flight, each on its own pair of lanes
        self.e2e_depth = max(1, min(int(os.environ.get("B200TFS_E2E_DEPTH", "4")), streams // 2))
        self.e2e = []
        for j in range(self.e2e_depth):
            slot = {"x": N.PinnedBuffer(self.P), "wire": N.PinnedBuffer(self.main.arena_cap), "resp": N.PinnedBuffer(self.S.resp_len),
                    "out": N.PinnedBuffer(self.P), "enc": self.lanes[(2 * j) % streams], "dec": self.lanes[(2 * j + 1) % streams],
                    "outs": (N.Output * N.FUSED_MAX_OUTPUTS)(), "n_outs": (C.c_int32 * 1)(), "status": (C.c_int32 * 1)(), "busy": False}
            slot["x"].array[:] = self.S.host_x[j % len(self.S.host_x)].view(np.uint8).reshape(-1)
            slot["resp"].array[:] = self.S.resp_host[j % len(self.S.resp_host)]
            self.e2e.append(slot)
        self.e2e_k = 0
        self.outs = (N.Output * 4)()
        self.e2e_outs = (N.Output * N.FUSED_MAX_OUTPUTS)()
        self.n_outs, self.specs, self.status = (C.c_int32 * 1)(), (N.ModelSpec * 1)(), (C.c_int32 * 1)()
        self.dst_ptr = (C.c_void_p * 1)()
        self.events = {}

    def footprint(self):
        return self.ring.bytes

    def sync(self):
        for l in self.lanes:
            l.sync()

    def launches(self):
        return sum(l.launches() for l in self.lanes)

    def _event(self, key):
        if key not in self.events:
            e = C.c_void_p()
            self.N.check(self.lib.b200tfs_event_create(C.byref(e)))
            self.events[key] = e
        return self.events[key]

    def timed_region(self, enqueue):
        """Device time of whatever `enqueue(lane)` submits on every lane: fork from lane 0, join back."""
        lib, N = self.lib, self.N
        self.sync()
        e0, e1 = self._event("t0"), self._event("t1")
        N.check(lib.b200tfs_event_record(self.main.ctx, e0))
        for l in self.lanes[1:]:
            N.check(lib.b200tfs_wait_event(l.ctx, e0))
        for l in self.lanes:
            enqueue(l)
        for k, l in enumerate(self.lanes[1:]):
            f = self._event(("join", k))
            N.check(lib.b200tfs_event_record(l.ctx, f))
            N.check(lib.b200tfs_wait_event(self.main.ctx, f))
        N.check(lib.b200tfs_event_record(self.main.ctx, e1))
        N.check(lib.b200tfs_event_sync(e1))
        self.sync()
        ms = C.c_float(0)
        N.check(lib.b200tfs_event_elapsed_ms(e0, e1, C.byref(ms)))
        return float(ms.value)

    # ---- the timed workload: K steps split over the lanes, replayed from graphs --------------------
    def prepare(self, steps, graph_steps):
        self.plan = []
        n, R = len(self.lanes), self.ring.slots
        for s, l in enumerate(self.lanes):
            mine = steps // n + (1 if s < steps % n else 0)
            g = min(graph_steps, max(mine, 1))
            full, rem = divmod(mine, g)
            slots = [(s + k * n) % R for k in range(g)]   # lane s walks the ring with stride n
            body = lambda i, l=l: (l.encode(i), l.decode(i))  # noqa: E731
            if full:
                l.capture("step", body, slots)
            if rem:
                l.capture("step_rem", body, slots[:rem])
            self.plan.append((full, rem))

    def run_steps(self):
        def enqueue(l):
            full, rem = self.plan[self.lanes.index(l)]
            for _ in range(full):
                l.launch("step")
            if rem:
                l.launch("step_rem")
        return self.timed_region(enqueue)

    # ---- one step through the host-buffer entry points (e2e) ---------------------------------------
    def _e2e_wait(self, slot):
        if slot["busy"]:
            lib, N = self.lib, self.N
            N.check(lib.b200tfs_decode_results(slot["dec"].ctx, 1, slot["outs"], slot["n_outs"], None, slot["status"]))  # synchronises
            N.check(lib.b200tfs_sync(slot["enc"].ctx))
            slot["busy"] = False

    def step_e2e(self):
        """Host tensor -> request wire bytes in host memory AND response wire bytes in host memory -> host tensor,
        through the host-buffer C-ABI entry points.  The two halves are independent, so they run on two contexts;
        up to `e2e_depth` steps are in flight (each with its own pinned buffers), so H2D and D2H copies overlap."""
        lib, N = self.lib, self.N
        slot = self.e2e[self.e2e_k % self.e2e_depth]
        self.e2e_k += 1
        self._e2e_wait(slot)
        a, b = slot["enc"], slot["dec"]
        a.tensors[0].data = slot["x"].ptr
        N.check(lib.b200tfs_encode_requests_host_async(a.ctx, 1, a.requests, slot["wire"].ptr, a.arena_cap, a.rec_off, a.rec_len))
        N.check(lib.b200tfs_decode_responses_host_async(b.ctx, slot["resp"].ptr, 1, b.p_off, b.p_len, slot["out"].ptr, self.P))
        slot["busy"] = True

    def e2e_drain(self):
        for slot in self.e2e:
            self._e2e_wait(slot)

    def timed_main(self, fn, steps):
        def enqueue(l):
            if l is self.main:
                for _ in range(steps):
                    fn()
        return self.timed_region(enqueue)

    def verify(self):
        """Bit-exact check, on every lane, of one ring slot against bytes built here from the inputs."""
        lib, N, S, R = self.lib, self.N, self.S, self.ring
        for li, l in enumerate(self.lanes):
            i = li % R.slots
            N.check(lib.b200tfs_memset(l.ctx, R.arena[i], 0, l.arena_cap))
            N.check(lib.b200tfs_memset(l.ctx, R.dst[i], 0, S.P))
            for _ in range(2):   # second pass takes the decode kernel's template fast path
                l.encode(i)
                l.decode(i)
            l.sync()
            wire = np.empty(int(l.rec_len[0]), dtype=np.uint8)
            N.check(lib.b200tfs_memcpy_d2h(l.ctx, wire.ctypes.data, R.arena[i] + int(l.rec_off[0]), wire.size))
            out = np.empty(S.SHAPE, dtype=np.float32)
            N.check(lib.b200tfs_memcpy_d2h(l.ctx, out.ctypes.data, R.dst[i], S.P))
            outs = (N.Output * N.FUSED_MAX_OUTPUTS)()
            n_outs, status = (C.c_int32 * 1)(), (C.c_int32 * 1)()
            N.check(lib.b200tfs_decode_results(l.ctx, 1, outs, n_outs, None, status))
            x = S.host_x[i % len(S.host_x)]
            assert wire.tobytes() == S.req_header + x.tobytes(), "encoded request differs from the expected wire bytes"
            assert out.tobytes() == x.tobytes(), "decoded tensor differs from the payload"
            assert status[0] == 0 and n_outs[0] == 1 and outs[0].dst_off == 0 and outs[0].dst_bytes == S.P
        return True

    def verify_e2e(self):
        for slot in self.e2e:
            slot["wire"].array[:] = 0
            slot["out"].array[:] = 0
        for _ in range(2 * self.e2e_depth):
            self.step_e2e()
        self.e2e_drain()
        for j, slot in enumerate(self.e2e):
            x = self.S.host_x[j % len(self.S.host_x)]
            assert slot["status"][0] == 0 and slot["n_outs"][0] == 1 and slot["outs"][0].dst_bytes == self.P and slot["outs"][0].dims[0] == 1024
            n, o = int(slot["enc"].rec_len[0]), int(slot["enc"].rec_off[0])
            assert slot["wire"].array[o:o + n].tobytes() == self.S.req_header + x.tobytes()
            assert slot["out"].array.tobytes() == x.tobytes()
        return True


def ncu_traffic():
    """dram read+write bytes per launch of the two hot kernels, from the newest committed ncu --set full capture."""
    import glob

    paths = sorted(glob.glob(os.path.join(REPO, "profiles", "r*_ncu_summary.json")))
    if not paths:
        return None, None
    with open(paths[-1]) as fh:
        caps = json.load(fh).get("full_capture", [])
    unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = {}
    for c in caps:
        r, w = c.get("dram__bytes_read.sum"), c.get("dram__bytes_write.sum")
        if r and w:
            per.setdefault(c["kernel"], []).append(float(r["value"]) * unit.get(r["unit"], 1) + float(w["value"]) * unit.get(w["unit"], 1))
    if not per:
        return None, None
    avg = {k: sum(v) / len(v) for k, v in per.items()}
    return sum(avg.values()) / len(avg), {"source": os.path.relpath(paths[-1], REPO), "per_kernel_bytes": avg,
                                          "note": "dram__bytes_write is ~0 inside the capture window: the 4 MiB of stores sit in the 126 MB L2 "
                                                  "when the kernel ends and are written back later; reads equal the algorithmic read bytes (1.02x)"}


def saturation_pass(bench, peak, n=1024):
    N, lib, lane = bench.N, bench.lib, bench.main
    P = 3 * 224 * 224 * 4
    try:
        src = lane.malloc(n * P)
    except Exception:
        return None
    N.check(lib.b200tfs_memset(lane.ctx, src, 0x3C, n * P))
    dims = (C.c_int64 * 3)(3, 224, 224)
    ts, rq = (N.Tensor * n)(), (N.Request * n)()
    for i in range(n):
        ts[i] = N.Tensor(data=src + i * P, src_dtype=1, wire_dtype=1, rank=3, flags=0, dims=dims, key=b"image", key_len=5, packed_len=0)
        rq[i] = N.Request(model_name=b"default", model_name_len=7, has_version=1, order=N.ORDER_UPB, version=1, n_inputs=1, flags=0,
                          inputs=C.cast(C.byref(ts, i * C.sizeof(N.Tensor)), C.POINTER(N.Tensor)))
    need = C.c_uint64()
    N.check(lib.b200tfs_request_arena_size(n, rq, C.byref(need)))
    arena = lane.malloc(need.value)
    off, ln = (C.c_uint64 * n)(), (C.c_uint64 * n)()

    def encode(_):
        N.check(lib.b200tfs_encode_requests(lane.ctx, n, rq, arena, need.value, off, ln))
    lane.capture("sat", encode, [0, 0])
    bench.timed_main(lambda: lane.launch("sat"), 2)
    reps = 10
    ms = bench.timed_main(lambda: lane.launch("sat"), reps)
    us = ms / (reps * 2) * 1e3
    alg = n * (2 * P + int(ln[0]) - P)
    out = {"workload": f"{n} PredictRequests of fp32[3,224,224] ({n * P >> 20} MiB) encoded by one move_kernel launch (BASELINE configs[4], per-GPU share)",
           "launch_us": us, "algorithmic_bytes_per_launch": alg, "achieved": alg / (us * 1e-6) / 1e9, "frac": alg / (us * 1e-6) / 1e9 / peak,
           "how": "CUDA graph of 2 launches replayed 10x on one stream, CUDA events; 1.2 GB working set"}
    # the decode side of the same share: 1024 PredictResponses of that size through ONE decode_fused_kernel launch
    prefix, suffix = response_wire_parts(b"image", (3, 224, 224), P)
    rec = np.frombuffer(prefix + bytes(P) + suffix, dtype=np.uint8)
    stride = (rec.size + 255) & ~255
    lib.b200tfs_free(lane.ctx, arena)
    wire = lane.malloc(stride * n)
    N.check(lib.b200tfs_memcpy_h2d(lane.ctx, wire, rec.ctypes.data, rec.size))
    lane.sync()
    for i in range(1, n):
        N.check(lib.b200tfs_memcpy_d2d(lane.ctx, wire + i * stride, wire, rec.size))
    roff = (C.c_uint64 * n)(*[i * stride for i in range(n)])
    rlen = (C.c_uint64 * n)(*[rec.size] * n)
    dst_stride = (P + 255) & ~255
    dst = src   # the request tensors are no longer needed: decode into their buffer (n * P >= n * dst_stride? P is 256-aligned: yes)
    assert dst_stride == P

    def decode(_):
        N.check(lib.b200tfs_decode_responses(lane.ctx, wire, n, roff, rlen, dst, dst_stride))
    lane.capture("sat_dec", decode, [0, 0])
    bench.timed_main(lambda: lane.launch("sat_dec"), 2)
    ms = bench.timed_main(lambda: lane.launch("sat_dec"), reps)
    st = (C.c_int32 * n)()
    N.check(lib.b200tfs_decode_results(lane.ctx, n, None, None, None, st))
    assert all(v == 0 for v in st)
    us_d = ms / (reps * 2) * 1e3
    alg_d = n * (2 * P + rec.size - P)
    out["decode"] = {"workload": f"{n} PredictResponses of fp32[3,224,224] decoded by one decode_fused_kernel launch", "launch_us": us_d,
                     "algorithmic_bytes_per_launch": alg_d, "achieved": alg_d / (us_d * 1e-6) / 1e9, "frac": alg_d / (us_d * 1e-6) / 1e9 / peak}
    lib.b200tfs_free(lane.ctx, src)
    lib.b200tfs_free(lane.ctx, wire)
    return out


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU legs (the only place oracle/ is touched)
# ------------------------------------------------------------------------------------------------
def _cpu_roundtrip(seed):
    """One C2 unit on one core through the reference port: encode request + decode response."""
    from oracle import ref_port

    x = np.random.default_rng(seed).standard_normal((1024, 1024), dtype=np.float32)
    prefix, suffix = response_wire_parts(b"y", (1024, 1024), 4194304)
    resp = prefix + x.tobytes() + suffix  # fabricating the response is not part of the measured path (and is cheap)
    t0 = time.perf_counter()
    wire = ref_port.encode_predict_request("default", 1, [("x", x)])
    out = ref_port.decode_predict_response(resp)["y"]
    t1 = time.perf_counter()
    assert out.tobytes() == x.tobytes() and len(wire) == 4194351
    return t1 - t0


def cpu_baseline_port(budget_s=12.0, max_units=64):
    """Scalar (1 core) timing of the reference port on C2 tensors: as many whole units as fit ~budget_s of CPU work."""
    t, units = 0.0, 0
    while units < 2 or (t + t / units <= budget_s and units < max_units):
        t += _cpu_roundtrip(units)
        units += 1
    payload = units * 2 * 4194304
    return {"value": payload / t / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
            "sample": f"{units} x fp32[1024,1024] encode+decode through oracle/ref_port.py (per-element Python, protobuf upb), {t:.2f} s"}


def cpu_c_oracle(units=8):
    from oracle import wire_oracle

    x = np.random.default_rng(0).standard_normal((1024, 1024), dtype=np.float32)
    resp = wire_oracle.build_predict_response([("y", x)])
    t0 = time.perf_counter()
    for _ in range(units):
        wire_oracle.encode_predict_request("default", 1, [("x", x)])
        wire_oracle.decode_predict_response(resp)
    t = time.perf_counter() - t0
    return {"value": units * 2 * 4194304 / t / 1e9, "unit": "GB/s", "cores": 1, "kind": "port (plain C, oracle/wire_oracle.c)"}


def run_reference(args, world):
    """--impl reference: the reference's CPU implementation (Python port over protobuf) on all host cores."""
    if world.rank != 0:
        return
    import multiprocessing as mp

    cores = len(os.sched_getaffinity(0))
    per_step = cores  # one C2 unit per core per step
    steps, warmup = max(1, min(args.steps, 5)), max(0, min(args.warmup, 1))
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(warmup):
            pool.map(_cpu_roundtrip, range(per_step))
        t0 = time.perf_counter()
        for s in range(steps):
            pool.map(_cpu_roundtrip, range(per_step))
        wall = time.perf_counter() - t0
    value = steps * per_step * 2 * 4194304 / wall / 1e9
    line = {
        "impl": "reference", "metric": "TensorProto encode+decode GB/s", "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": wall / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "C2 fp32[1024,1024] single-tensor PredictRequest encode + PredictResponse decode",
                   "step": f"{per_step} tensors per step, one per host core (bounded sample of the same workload)"},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} steps x {per_step} tensors, multiprocessing pool over {cores} cores, oracle/ref_port.py"},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100, help="timed steps; one step = one batch of --batch request/response pairs")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=480, help="C2 PredictRequest/PredictResponse pairs per step (spread over the lanes)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ring", type=int, default=48, help="ring slots in total (split over the streams)")
    ap.add_argument("--streams", type=int, default=16, help="independent lanes (native contexts = CUDA streams) per GPU")
    ap.add_argument("--graph-steps", type=int, default=48, help="steps recorded per CUDA graph")
    ap.add_argument("--e2e-steps", type=int, default=200, help="request/response PAIRS timed by the e2e leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":   # CPU only: rank 0 works alone, nobody needs a process group
        class _Solo:
            rank = int(os.environ.get("RANK", "0"))
        run_reference(args, _Solo())
        return
    # stdout carries exactly ONE JSON line: park fd 1 on stderr while libraries (NCCL prints its version) are chatty
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    world = World()
    warmup = max(args.warmup, 3)
    batch = max(args.batch, 1)
    pairs = args.steps * batch                     # a step is a batch of `batch` independent pairs: K steps = K * batch pairs
    bench = C2Bench(world.local_rank, args.ring, args.streams)
    assert bench.footprint() > 2 * L2_BYTES, "ring must exceed L2"
    bench.verify()
    bench.prepare(warmup * batch, args.graph_steps)
    bench.run_steps()                              # W untimed warm-up steps
    bench.prepare(pairs, args.graph_steps)         # graphs sized so that exactly K steps run
    bench.run_steps()                              # one untimed pass so the new graphs are uploaded
    launches0 = bench.launches()
    sampler = ClockSampler(world.local_rank)
    sampler.start()
    time.sleep(0.3)
    world.barrier()
    t0 = time.time()
    ms = bench.run_steps()
    t1 = time.time()
    world.barrier()
    clocks = sampler.stop(t0, t1)
    launches_eager = bench.launches() - launches0  # graph replays do not pass through the counter ...
    launches = 2 * pairs                           # ... each pair replays one move_kernel + one decode_fused_kernel
    ms_max = world.max(ms)
    S = bench.S
    payload_per_pair = 2 * S.P
    payload_per_step = payload_per_pair * batch
    total_payload = world.sum(float(payload_per_pair * pairs))
    value = total_payload / (ms_max * 1e-3) / 1e9
    enc_bytes, dec_bytes = 2 * S.P + S.H_req, 2 * S.P + S.H_resp
    peak, peak_src = peaks()

    # roofline pass: the dominant kernel alone on ONE stream, back to back inside a graph (no CPU in
    # the loop), ring > L2.  avg launch duration = region / launches (includes the inter-kernel gap).
    m = bench.main
    reps = 20
    every = list(range(bench.ring.slots))            # the whole ring, so this pass is out of L2 as well
    m.capture("enc", m.encode, every)
    m.capture("dec", m.decode, every)
    per = {}
    for name in ("enc", "dec"):
        bench.timed_main(lambda: m.launch(name), 3)
        t = bench.timed_main(lambda: m.launch(name), reps)
        per[name] = t / (reps * m.graphs[name][1]) * 1e3  # us per launch
    avg_us = (per["enc"] + per["dec"]) / 2
    achieved = (enc_bytes + dec_bytes) / 2 / (avg_us * 1e-6) / 1e9
    agg = (enc_bytes + dec_bytes) * pairs / (ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic()
    roofline = {"bound": "hbm", "kernel": "move_kernel (encode) / decode_fused_kernel (decode)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": (enc_bytes + dec_bytes) / 2, "avg_launch_us": avg_us,
                "encode_launch_us": per["enc"], "decode_launch_us": per["dec"],
                "encode_frac": enc_bytes / (per["enc"] * 1e-6) / 1e9 / peak, "decode_frac": dec_bytes / (per["dec"] * 1e-6) / 1e9 / peak,
                "how": "one stream, graph of back-to-back launches of that kernel alone, CUDA events on the launching stream, ring > L2",
                "timed_region_aggregate": {"achieved": agg, "frac": agg / peak, "streams": len(bench.lanes),
                                           "how": "algorithmic bytes of every launch in the timed region / region time (launches of "
                                                  "independent requests overlap across the streams)"}}

    # saturation pass: the same kernel on a batch far above the bandwidth-delay product - the per-GPU share of
    # BASELINE configs[4] (1024 requests of fp32 [3,224,224] = 617 MB) encoded by ONE launch, replayed from a graph
    sat = saturation_pass(bench, peak)
    if sat:
        roofline["saturated"] = sat

    # e2e: host buffers in, host buffers out, copies inside the timed region
    bench.verify_e2e()
    for _ in range(8):
        bench.step_e2e()
    bench.e2e_drain()
    e2e_steps = max(10, min(args.e2e_steps, pairs))    # pairs, not steps: the leg is PCIe-bound, 200 pairs are 50 ms
    world.barrier()
    def e2e_region(l):
        if l is bench.main:
            for _ in range(e2e_steps):
                bench.step_e2e()
            bench.e2e_drain()
    e2e_ms = world.max(bench.timed_region(e2e_region))
    e2e_value = world.sum(float(payload_per_pair * e2e_steps)) / (e2e_ms * 1e-3) / 1e9
    e2e = {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": (S.P + S.resp_len) * batch, "d2h_bytes_per_step": (S.P + S.H_req + S.P) * batch,
           "ms_per_step": e2e_ms / e2e_steps * batch, "pairs_timed": e2e_steps, "pairs_per_step": batch,
           "in_flight": bench.e2e_depth,
           "how": "b200tfs_encode_requests_host_async + b200tfs_decode_responses_host_async / b200tfs_decode_results on pinned host "
                  "buffers; every pair copies its tensor and its response wire H2D and its request wire and decoded tensor D2H "
                  f"(4 x ~4 MiB over PCIe); up to {bench.e2e_depth} pairs in flight so the two copy directions overlap"}

    # the drop-in Python API on ordinary numpy arrays / bytes objects (pageable memory, bytes copies): reported, not the headline
    py_api = None
    if world.rank == 0:
        try:
            from min_tfs_client.codec import Codec

            codec = Codec(world.local_rank)
            x = S.host_x[0]
            resp_bytes = S.resp_host[0].tobytes()
            for _ in range(3):
                w = codec.encode_predict_request("default", {"x": x}, 1)
                y = codec.decode_predict_response(resp_bytes)[0]["y"]
            assert w == S.req_header + x.tobytes() and y.tobytes() == x.tobytes()
            reps = 20
            t0 = time.perf_counter()
            for _ in range(reps):
                codec.encode_predict_request("default", {"x": x}, 1)
                codec.decode_predict_response(resp_bytes)
            dt = (time.perf_counter() - t0) / reps
            py_api = {"value": payload_per_pair / dt / 1e9, "unit": "GB/s", "ms_per_pair": dt * 1e3,
                      "how": "min_tfs_client.codec.Codec.encode_predict_request + decode_predict_response on numpy arrays and bytes (wall clock)"}
            codec.close()
        except Exception as exc:  # pragma: no cover
            py_api = {"error": repr(exc)}
    e2e["python_api"] = py_api

    if world.rank == 0:
        line = {
            "metric": "TensorProto encode+decode GB/s", "value": value, "unit": "GB/s", "n_gpus": world.size, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "C2 fp32[1024,1024] single-tensor PredictRequest encode + PredictResponse decode (BASELINE.json configs[1])",
                       "step": f"one batch of {batch} independent request/response pairs, spread over the lanes",
                       "pairs_per_step": batch, "payload_bytes_per_step": payload_per_step, "ring_slots": bench.ring.slots, "ring_bytes": bench.footprint(),
                       "l2": f"inputs rotate through a ring of {bench.ring.slots} slots = {bench.footprint() >> 20} MiB > 126 MiB L2 "
                             "(timed region and roofline pass alike)",
                       "streams": len(bench.lanes), "cuda_graph_steps": args.graph_steps,
                       "wire_mode": "typed (float_val, sNaN-quieting on: bit-exact vs reference)", "sharding": "independent requests per GPU, no collective", "cpu_binding": world.numa},
            "roofline": roofline, "e2e": e2e, "gpu_launches": launches, "gpu_launches_outside_graphs": launches_eager, "clocks": clocks,
        }
        if world.size == 1 and not args.no_cpu:
            cb = cpu_baseline_port()
            cb["c_oracle_1core_gbs"] = cpu_c_oracle(8)["value"]
            line["cpu_baseline"] = cb
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    world.close()


if __name__ == "__main__":
    main()
