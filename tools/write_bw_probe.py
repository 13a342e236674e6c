#!/usr/bin/env python
"""HBM write bandwidth on its own: cudaMemset / a torch fill over 1 GiB against the 1:1 copy the measured peak comes from, and a
read-only reduction.  Context for the store-heavy C4 encode (fp16 -> fp32: one byte read per two written).  `python tools/write_bw_probe.py`"""
import json
import torch

n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
f = torch.empty(n // 4, dtype=torch.float32, device="cuda")
h = torch.empty(n // 4, dtype=torch.float16, device="cuda").normal_()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


out = {}
t = timed(lambda: a.fill_(7)); out["fill_1GiB_write_only_GBs"] = n / t / 1e9
t = timed(lambda: a.zero_()); out["memset_1GiB_write_only_GBs"] = n / t / 1e9
t = timed(lambda: b.copy_(a)); out["copy_1GiB_read_plus_write_GBs"] = 2 * n / t / 1e9
t = timed(lambda: a.sum(dtype=torch.int64)); out["sum_1GiB_read_only_GBs"] = n / t / 1e9
t = timed(lambda: f.copy_(h)); out["torch_fp16_to_fp32_copy_read_plus_write_GBs"] = (h.numel() * 2 + f.numel() * 4) / t / 1e9
t = timed(lambda: h.copy_(f)); out["torch_fp32_to_fp16_copy_read_plus_write_GBs"] = (h.numel() * 2 + f.numel() * 4) / t / 1e9
print(json.dumps(out, indent=1))
