"""bench.py pieces that need no GPU: the response bytes the timed decode consumes are what the oracle builds, the CPU legs
return well-formed records, the reference arm prints one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from oracle import wire_oracle  # noqa: E402


def test_response_wire_parts_are_the_oracles_bytes():
    x = np.random.default_rng(0).standard_normal((64, 48)).astype(np.float32)
    pre, suf = bench.response_wire_parts(b"y", (64, 48), x.nbytes)
    assert pre + x.tobytes() + suf == wire_oracle.build_predict_response([("y", x)])


def test_cpu_baseline_leg_shape():
    r = bench.cpu_baseline_port(budget_s=0.5, max_units=2)
    assert r["kind"] == "port" and r["cores"] == 1 and r["unit"] == "GB/s" and 0 < r["value"] < 1.0
    assert r["sample"].startswith("2 x fp32[1024,1024]")
    assert bench.cpu_c_oracle(1)["value"] > r["value"]          # plain C beats per-element Python


def test_numa_binding_is_best_effort():
    assert isinstance(bench.bind_to_gpu_numa_node(0), str)       # no GPU here: says why it did nothing


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, RANK="0")
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=600).stdout.strip().splitlines()
    assert len(out) == 1
    line = json.loads(out[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config", "cpu_baseline", "e2e"):
        assert key in line
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    # a rank other than 0 does no work and prints nothing
    out1 = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1"], capture_output=True, text=True,
                          env=dict(os.environ, RANK="1"), timeout=60)
    assert out1.returncode == 0 and out1.stdout.strip() == ""
