"""Helpers shared by the tests: rebuild golden inputs from their recipes (tests/golden/*.json).

``make_array`` must stay in step with ``tests/golden/make_golden.py`` (the script that produced the
fixtures by running the unmodified reference).
"""
import hashlib
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLDEN_DIR, name)) as fh:
        return json.load(fh)["cases"]


def make_array(recipe):
    kind = recipe["gen"]
    if kind == "hex":
        a = np.frombuffer(bytes.fromhex(recipe["data"]), dtype=np.dtype(recipe["dtype"])).copy()
        return a.reshape(recipe["shape"])
    if kind == "strings":
        return np.array(recipe["data"], dtype=np.str_).reshape(recipe["shape"])
    if kind == "arange":
        return np.arange(int(np.prod(recipe["shape"])), dtype=np.dtype(recipe["dtype"])).reshape(recipe["shape"])
    rng = np.random.default_rng(recipe["seed"])
    if recipe["dtype"] == "bfloat16":
        import ml_dtypes

        return rng.standard_normal(recipe["shape"]).astype(ml_dtypes.bfloat16)
    dt = np.dtype(recipe["dtype"])
    if kind == "standard_normal":
        if dt in (np.dtype(np.float32), np.dtype(np.float64)):
            return rng.standard_normal(recipe["shape"], dtype=dt)
        return rng.standard_normal(recipe["shape"]).astype(dt)
    if kind == "random_bits":
        n = int(np.prod(recipe["shape"]))
        raw = rng.integers(0, 256, size=n * dt.itemsize, dtype=np.uint8)
        return raw.view(dt).reshape(recipe["shape"])
    if kind == "varint_mix":
        n = int(np.prod(recipe["shape"]))
        bits = rng.integers(0, dt.itemsize * 8 + 1, size=n)
        raw = rng.integers(0, 2 ** 63, size=n, dtype=np.uint64) | (rng.integers(0, 2, size=n, dtype=np.uint64) << np.uint64(63))
        mask = np.where(bits >= 64, np.uint64(0xFFFFFFFFFFFFFFFF), (np.uint64(1) << bits.astype(np.uint64)) - np.uint64(1))
        vals = raw & mask
        if dt.kind == "i":
            neg = rng.integers(0, 2, size=n).astype(bool)
            v = vals.astype(np.uint64)
            v = np.where(neg, ~v, v)
            return v.astype(np.uint64).view(np.int64).astype(dt).reshape(recipe["shape"])
        return vals.astype(dt).reshape(recipe["shape"])
    raise ValueError(kind)


def apply_transform(a, transform):
    if transform is None:
        return a
    if transform == "T":
        return a.T
    if transform == "[:, ::2]":
        return a[:, ::2]
    raise ValueError(transform)


def check_wire(wire: bytes, summary: dict, label=""):
    """Assert `wire` equals the golden summary (full hex for small cases; len+sha256+edges for big)."""
    assert len(wire) == summary["len"], f"{label}: length {len(wire)} != {summary['len']}"
    if "hex" in summary:
        assert wire.hex() == summary["hex"], f"{label}: bytes differ"
    else:
        assert wire[:96].hex() == summary["head"], f"{label}: head differs"
        assert wire[-96:].hex() == summary["tail"], f"{label}: tail differs"
    assert hashlib.sha256(wire).hexdigest() == summary["sha256"], f"{label}: sha256 differs"


# ---- hand-built response wires (same builders as make_golden.py, for the two large decode cases) ---
def vi(x):
    x &= (1 << 64) - 1
    b = bytearray()
    while True:
        if x < 0x80:
            b.append(x)
            return bytes(b)
        b.append((x & 0x7F) | 0x80)
        x >>= 7


def ld(tag, payload):
    return bytes([tag]) + vi(len(payload)) + payload


def shape(*dims):
    return b"".join(ld(0x12, (b"\x08" + vi(d)) if d else b"") for d in dims)


def tproto(dtype, dims, values_field):
    return b"\x08" + vi(dtype) + ld(0x12, shape(*dims)) + values_field


def entry(key, tp):
    return ld(0x0A, ld(0x0A, key.encode()) + ld(0x12, tp))


def mspec(name=b"default", version=1, sig=b"serving_default"):
    return ld(0x12, ld(0x0A, name) + ld(0x12, b"\x08" + vi(version)) + ld(0x1A, sig))


def decode_case_wire(name, rec):
    if rec.get("wire") is not None:
        return bytes.fromhex(rec["wire"])
    if name == "c2_response":
        big = make_array({"gen": "standard_normal", "seed": 0, "dtype": "<f4", "shape": [1024, 1024]})
        wire = entry("y", tproto(1, [1024, 1024], ld(0x2A, big.tobytes()))) + mspec()
    else:
        raise KeyError(name)
    assert len(wire) == rec["wire_len"] and hashlib.sha256(wire).hexdigest() == rec["wire_sha256"]
    return wire
