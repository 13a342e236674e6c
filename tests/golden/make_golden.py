#!/usr/bin/env python
"""Generate tests/golden/*.json by running the UNMODIFIED reference in this container.

Runs only where ``/root/reference`` exists (the build container).  It puts the reference's
``tensor_serving_client/`` first on ``sys.path`` (so ``min_tfs_client`` is the reference's own
package: tensors.py / types.py / constants.py) and this repo's schema modules second (the
``*_pb2`` modules the reference would otherwise generate with protoc at install time), then records

* encode vectors: ``ndarray_to_tensor_proto(x).SerializeToString()`` and the PredictRequest the
  reference's ``_make_inference_request`` builds (reference requests.py:41-48), serialised the way
  the gRPC stub does (prediction_service_pb2_grpc.py:52), with ``deterministic=True`` when there is
  more than one input (map order is otherwise per-process random - SURVEY 8a Q1);
* decode vectors: ``PredictResponse.FromString(wire)`` (…pb2_grpc.py:53) followed by
  ``tensor_proto_to_ndarray`` (tensors.py:42-46) per output, or the exception type it raises.

Small cases carry full bytes; large ones (C2-C5 sizes) carry a generator recipe, the length, the
SHA-256 of the expected wire and its first/last bytes.

    python tests/golden/make_golden.py        # rewrites encode.json / decode.json / requests.json
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/tensor_serving_client"
if not os.path.isdir(REF):
    sys.exit("reference not present; goldens can only be regenerated in the build container")
sys.path[:0] = [REF, os.path.join(REPO, "min-tfs-client_b200")]

# The reference's package has no __init__.py (a namespace package); this repo's drop-in of the same name is a
# regular package and would win the import.  Pin the name to the reference's directory explicitly.
import types as _types  # noqa: E402

_ref_pkg = _types.ModuleType("min_tfs_client")
_ref_pkg.__path__ = [os.path.join(REF, "min_tfs_client")]
sys.modules["min_tfs_client"] = _ref_pkg

import numpy as np  # noqa: E402
import google.protobuf  # noqa: E402
import min_tfs_client.tensors as ref_tensors  # noqa: E402
from tensorflow.core.framework.tensor_pb2 import TensorProto  # noqa: E402
from tensorflow_serving.apis.predict_pb2 import PredictRequest, PredictResponse  # noqa: E402

assert ref_tensors.__file__.startswith("/root/reference/"), "not the reference package: " + ref_tensors.__file__

enc = ref_tensors.ndarray_to_tensor_proto
dec = ref_tensors.tensor_proto_to_ndarray


# ------------------------------------------------------------------------------------------
# input recipes (shared with tests/golden_util.py - keep in sync)
# ------------------------------------------------------------------------------------------
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
        # integers whose varint lengths cover every byte count for the dtype
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


def hexarr(a):
    a = np.asarray(a)
    shape = list(a.shape)          # np.ascontiguousarray would promote a 0-d array to shape (1,)
    return {"gen": "hex", "dtype": a.dtype.str, "shape": shape, "data": np.ascontiguousarray(a).tobytes().hex()}


def summarize(wire, full_limit=4096):
    out = {"len": len(wire), "sha256": hashlib.sha256(wire).hexdigest()}
    if len(wire) <= full_limit:
        out["hex"] = wire.hex()
    else:
        out["head"] = wire[:96].hex()
        out["tail"] = wire[-96:].hex()
    return out


# ------------------------------------------------------------------------------------------
# encode: TensorProto
# ------------------------------------------------------------------------------------------
def tensor_cases():
    c = []

    def add(name, recipe):
        c.append((name, recipe))

    add("kat1_f32_4x4", {"gen": "arange", "dtype": "<f4", "shape": [4, 4]})
    add("f64_ref_unit_golden", hexarr(np.array([0.314, 0.159, 0.268, 0.358], dtype=np.float64)))  # tensors_test.py:66-83
    add("f32_rank0", hexarr(np.float32(3.0).reshape(())))     # KAT-7: 08 01 12 00 2a 04 00 00 40 40
    add("i64_rank0", hexarr(np.int64(-3)))
    add("f64_rank0", hexarr(np.float64(2.5)))
    add("f32_0x3", hexarr(np.zeros((0, 3), dtype=np.float32)))
    add("f32_3x0x2", hexarr(np.zeros((3, 0, 2), dtype=np.float32)))
    add("f32_rank8", {"gen": "standard_normal", "seed": 5, "dtype": "<f4", "shape": [1, 2, 1, 3, 1, 2, 1, 5]})
    add("f32_dim_300", {"gen": "standard_normal", "seed": 6, "dtype": "<f4", "shape": [300]})
    add("f32_dim_16384x2", {"gen": "standard_normal", "seed": 7, "dtype": "<f4", "shape": [16384, 2]})
    add("f32_snan", hexarr(np.array([0x7F800001, 0xFF800001, 0x7FC00001, 0x7FBFFFFF, 0x80000000, 1, 0x7F800000, 0xFF800000, 0x7FFFFFFF],
                                    dtype=np.uint32).view(np.float32)))
    add("f32_random_bits", {"gen": "random_bits", "seed": 11, "dtype": "<f4", "shape": [257, 33]})
    add("f64_random_bits", {"gen": "random_bits", "seed": 12, "dtype": "<f8", "shape": [129, 17]})
    add("f64_snan", hexarr(np.array([0x7FF0000000000001, 0xFFF0000000000001, 0x8000000000000000, 1], dtype=np.uint64).view(np.float64)))
    add("i8_edges", hexarr(np.array([-1, 127, -128, 0, 1], dtype=np.int8)))
    add("i16_edges", hexarr(np.array([-1, 32767, -32768, 0, 128, 16384], dtype=np.int16)))
    add("i32_edges", hexarr(np.array([-1, 0, 1, 127, 128, 16383, 16384, 2 ** 31 - 1, -2 ** 31], dtype=np.int32)))
    add("i64_edges", hexarr(np.array([-1, 0, 1, 127, 128, 2 ** 35, 2 ** 63 - 1, -2 ** 63, 2 ** 56 - 1, 2 ** 56], dtype=np.int64)))
    add("u8_edges", hexarr(np.array([0, 1, 127, 128, 255], dtype=np.uint8)))
    add("u16_edges", hexarr(np.array([0, 127, 128, 16383, 16384, 65535], dtype=np.uint16)))
    add("u32_edges", hexarr(np.array([0, 127, 128, 2 ** 21 - 1, 2 ** 21, 2 ** 28, 2 ** 32 - 1], dtype=np.uint32)))
    add("u64_edges", hexarr(np.array([0, 127, 128, 2 ** 63 - 1, 2 ** 63, 2 ** 64 - 1, 2 ** 49, 2 ** 49 - 1], dtype=np.uint64)))
    add("bool_mix", hexarr(np.array([True, False, True, True, False], dtype=np.bool_)))
    add("bool_2x3", hexarr(np.array([[1, 0, 1], [0, 0, 1]], dtype=np.bool_)))
    add("i64_label", hexarr(np.array([7], dtype=np.int64)))  # KAT-4
    for dt in ("<i1", "<i2", "<i4", "<i8", "<u1", "<u2", "<u4", "<u8"):
        add("varint_mix_" + dt[1:], {"gen": "varint_mix", "seed": 21, "dtype": dt, "shape": [37, 29]})
    add("str_two", {"gen": "strings", "shape": [2], "data": ["hello world", "a"]})
    add("str_unicode_2x2", {"gen": "strings", "shape": [2, 2], "data": ["Ceci", "n'est", "pas", "une pipe é中"]})
    add("str_empty_elem", {"gen": "strings", "shape": [3], "data": ["", "x", ""]})
    # C3 shapes (small enough to store by recipe)
    add("c3_image", {"gen": "standard_normal", "seed": 0, "dtype": "<f4", "shape": [3, 224, 224]})
    add("c3_scores", {"gen": "standard_normal", "seed": 10000, "dtype": "<f4", "shape": [1000]})
    # C2
    add("c2_f32_1024x1024", {"gen": "standard_normal", "seed": 0, "dtype": "<f4", "shape": [1024, 1024]})
    add("c2_f32_random_bits", {"gen": "random_bits", "seed": 1, "dtype": "<f4", "shape": [1024, 1024]})
    add("f64_512x300", {"gen": "standard_normal", "seed": 3, "dtype": "<f8", "shape": [512, 300]})
    out = {}
    for name, recipe in c:
        x = make_array(recipe)
        wire = enc(x).SerializeToString()
        out[name] = {"input": recipe, "wire": summarize(wire)}
    # non-contiguous / non-native inputs: the reference ravel()s in C order and takes .item()
    a = np.arange(12, dtype=np.float32).reshape(3, 4)
    out["f32_transposed_view"] = {"input": hexarr(a), "transform": "T", "wire": summarize(enc(a.T).SerializeToString())}
    out["f32_strided_view"] = {"input": hexarr(a), "transform": "[:, ::2]", "wire": summarize(enc(a[:, ::2]).SerializeToString())}
    be = np.arange(5, dtype=">f4")
    out["f32_big_endian"] = {"input": {"gen": "hex", "dtype": ">f4", "shape": [5], "data": be.tobytes().hex()},
                             "wire": summarize(enc(be).SerializeToString())}
    # what the reference refuses to encode
    errs = {}
    for label, arr in (("float16", np.array([1, 2], dtype=np.float16)), ("complex64", np.array([1 + 2j], dtype=np.complex64)),
                       ("complex128", np.array([1 + 2j], dtype=np.complex128)), ("bytes_", np.array([b"ab"])),
                       ("object_", np.array([None], dtype=object)), ("datetime64", np.array(["2020-01-01"], dtype="datetime64[D]"))):
        try:
            enc(arr)
            errs[label] = None
        except Exception as e:  # noqa: BLE001
            errs[label] = type(e).__name__
    out["_refused"] = errs
    return out


# ------------------------------------------------------------------------------------------
# encode: PredictRequest (reference requests.py:41-48 restated only as the call sequence)
# ------------------------------------------------------------------------------------------
def build_request(model_name, model_version, inputs):
    request = PredictRequest()
    request.model_spec.name = model_name
    if model_version is not None:
        request.model_spec.version.value = model_version
    for k, v in inputs:
        request.inputs[k].CopyFrom(enc(v))
    return request.SerializeToString(deterministic=len(inputs) > 1)


def request_cases():
    f1 = {"gen": "hex", "dtype": "<f4", "shape": [1], "data": np.array([1.0], dtype=np.float32).tobytes().hex()}
    cases = [
        ("kat2_c1", "default", 1, [("x", {"gen": "arange", "dtype": "<f4", "shape": [4, 4]})]),
        ("kat3_c2", "default", 1, [("x", {"gen": "standard_normal", "seed": 0, "dtype": "<f4", "shape": [1024, 1024]})]),
        ("no_version", "default", None, [("x", f1)]),
        ("version_0", "m", 0, [("x", f1)]),
        ("version_neg", "m", -1, [("x", f1)]),
        ("version_big", "m", 2 ** 40 + 5, [("x", f1)]),
        ("empty_name", "", None, [("x", f1)]),
        ("empty_key", "m", 1, [("", f1)]),
        ("no_inputs", "m", 1, []),
        ("long_names", "resnet50_v2_" + "z" * 200, 123456789, [("input_tensor_" + "k" * 140, f1)]),
        ("unicode_names", "modèle", 2, [("clé", f1)]),
        ("order_quirk", "m", None, [(k, f1) for k in ("b", "a", "aa", "ab", "B", "", "abc", "aé")]),
        ("kat4_c3_req0", "default", 1, [("image", {"gen": "standard_normal", "seed": 0, "dtype": "<f4", "shape": [3, 224, 224]}),
                                          ("label", hexarr(np.array([0], dtype=np.int64)))]),
        ("c3_req7", "default", 1, [("image", {"gen": "standard_normal", "seed": 7, "dtype": "<f4", "shape": [3, 224, 224]}),
                                    ("label", hexarr(np.array([7], dtype=np.int64)))]),
        ("c5_req3", "default", 1, [("image", {"gen": "standard_normal", "seed": 3, "dtype": "<f4", "shape": [3, 224, 224]})]),
        ("mixed_dtypes", "identity", 4, [("float_input", {"gen": "standard_normal", "seed": 1, "dtype": "<f4", "shape": [2, 3]}),
                                          ("int_input", hexarr(np.array([[1, -2], [300, 2 ** 40]], dtype=np.int64))),
                                          ("string_input", {"gen": "strings", "shape": [2], "data": ["hello", "world"]}),
                                          ("bool_input", hexarr(np.array([True, False]))),
                                          ("double_input", {"gen": "standard_normal", "seed": 2, "dtype": "<f8", "shape": [5]})]),
        # scalar placeholders (keep_prob and friends): 0-d arrays keep an EMPTY tensor_shape (`12 00`), SURVEY Q4
        ("scalar_inputs", "m", 1, [("keep_prob", hexarr(np.float32(0.5))), ("step", hexarr(np.int64(-7))), ("flag", hexarr(np.bool_(True)))]),
    ]
    out = {}
    for name, model, ver, ins in cases:
        wire = build_request(model, ver, [(k, make_array(r)) for k, r in ins])
        out[name] = {"model_name": model, "model_version": ver, "inputs": [[k, r] for k, r in ins], "wire": summarize(wire)}
    # C4: f16 / bf16 sources, DT_FLOAT on the wire: expected bytes are the reference's encoding of x.astype(f32)
    for label, np_dt in (("f16", np.float16),):
        x = make_array({"gen": "standard_normal", "seed": 4, "dtype": "<f2", "shape": [8, 512, 1024]})
        wire = build_request("default", 1, [("x", x.astype(np.float32))])
        out["c4_%s_as_float" % label] = {"model_name": "default", "model_version": 1,
                                         "inputs": [["x", {"gen": "standard_normal", "seed": 4, "dtype": "<f2", "shape": [8, 512, 1024]}]],
                                         "wire_dtype": "DT_FLOAT", "wire": summarize(wire)}
    try:
        import ml_dtypes
        x = np.random.default_rng(4).standard_normal((8, 512, 1024)).astype(ml_dtypes.bfloat16)
        wire = build_request("default", 1, [("x", x.astype(np.float32))])
        out["c4_bf16_as_float"] = {"model_name": "default", "model_version": 1,
                                   "inputs": [["x", {"gen": "standard_normal", "seed": 4, "dtype": "bfloat16", "shape": [8, 512, 1024]}]],
                                   "wire_dtype": "DT_FLOAT", "wire": summarize(wire)}
    except ImportError:
        pass
    return out


# ------------------------------------------------------------------------------------------
# decode: hand-built wire -> reference FromString + tensor_proto_to_ndarray
# ------------------------------------------------------------------------------------------
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


def decode_cases():
    f = lambda a: np.asarray(a, dtype=np.float32).tobytes()  # noqa: E731
    scores = make_array({"gen": "standard_normal", "seed": 10000, "dtype": "<f4", "shape": [1000]})
    big = make_array({"gen": "standard_normal", "seed": 0, "dtype": "<f4", "shape": [1024, 1024]})
    cases = {
        "kat5_canonical": entry("scores", tproto(1, [1000], ld(0x2A, scores.tobytes()))) + mspec(),
        "c2_response": entry("y", tproto(1, [1024, 1024], ld(0x2A, big.tobytes()))) + mspec(),
        "empty_response": b"",
        "model_spec_first": mspec() + entry("a", tproto(1, [2], ld(0x2A, f([1, 2])))),
        "values_before_dtype_shape": entry("a", ld(0x2A, f([1, 2, 3])) + ld(0x12, shape(3)) + b"\x08\x01"),
        "split_packed": entry("a", tproto(1, [4], ld(0x2A, f([1, 2])) + ld(0x2A, f([3, 4])))),
        "mixed_packed_unpacked": entry("a", tproto(1, [3], ld(0x2A, f([1, 2])) + b"\x2D" + f([3]))),
        "all_unpacked": entry("a", tproto(1, [2], b"\x2D" + f([5]) + b"\x2D" + f([6]))),
        "unknown_fields": ld(0x0A, ld(0x0A, b"a") + ld(0x12, b"\xA0\x06\x05" + tproto(1, [2], b"\xAA\x06\x04junk" + ld(0x2A, f([1, 2]))) + b"\xAA\x06\x00")) + b"\xB8\x06\x07",
        "nonminimal_len": entry("a", b"\x08\x01" + ld(0x12, shape(2)) + b"\x2A\x88\x00" + f([1, 2])),
        "dim_name_present": entry("a", b"\x08\x01" + ld(0x12, ld(0x12, b"\x08\x02" + ld(0x12, b"batch"))) + ld(0x2A, f([1, 2]))),
        "dtype_twice_last_wins": entry("a", b"\x08\x02\x08\x01" + ld(0x12, shape(2)) + ld(0x2A, f([1, 2]))),
        "shape_twice_merges": entry("a", b"\x08\x01" + ld(0x12, shape(1)) + ld(0x12, shape(2)) + ld(0x2A, f([1, 2]))),
        "dup_key_last_wins": entry("a", tproto(1, [1], ld(0x2A, f([1])))) + entry("a", tproto(1, [2], ld(0x2A, f([8, 9])))),
        "value_before_key": ld(0x0A, ld(0x12, tproto(1, [1], ld(0x2A, f([4])))) + ld(0x0A, b"k")),
        "entry_without_key": ld(0x0A, ld(0x12, tproto(1, [1], ld(0x2A, f([4]))))),
        "two_outputs": entry("scores", tproto(1, [2, 2], ld(0x2A, f([1, 2, 3, 4])))) + entry("classes", tproto(9, [2], ld(0x52, vi(3) + vi(-1)))) + mspec(),
        "int_val_truncates": entry("a", tproto(3, [2], ld(0x3A, vi(-1) + vi(2 ** 32 + 5)))),
        "bool_byte_2": entry("a", tproto(10, [3], ld(0x5A, b"\x02\x00\x01"))),
        "int8_overflow": entry("a", tproto(6, [1], ld(0x3A, vi(300)))),
        "uint8_ok": entry("a", tproto(4, [3], ld(0x3A, vi(0) + vi(200) + vi(255)))),
        "int16_neg": entry("a", tproto(5, [2], ld(0x3A, vi(-2) + vi(1234)))),
        "uint16": entry("a", tproto(17, [2], ld(0x3A, vi(65535) + vi(1)))),
        "uint32": entry("a", tproto(22, [2], b"\x82\x01" + vi(6) + vi(2 ** 32 - 1) + vi(7))),
        "uint64": entry("a", tproto(23, [2], b"\x8A\x01" + vi(11) + vi(2 ** 64 - 1) + vi(7))),
        "int64_edges": entry("a", tproto(9, [3], ld(0x52, vi(-2 ** 63) + vi(2 ** 63 - 1) + vi(0)))),
        "double": entry("a", tproto(2, [2], ld(0x32, np.array([1.5, -2.25]).tobytes()))),
        "f32_snan": entry("a", tproto(1, [3], ld(0x2A, np.array([0x7F800001, 0xFF800001, 0x7FC00000], dtype=np.uint32).tobytes()))),
        "f64_snan": entry("a", tproto(2, [1], ld(0x32, np.array([0x7FF0000000000001], dtype=np.uint64).tobytes()))),
        "strings_ascii": entry("a", tproto(7, [2, 1], ld(0x42, b"hello") + ld(0x42, b""))),
        "strings": entry("a", tproto(7, [2], ld(0x42, b"hello") + ld(0x42, b"\xc3\xa9"))),
        "dim_minus_one": entry("a", tproto(1, [-1, 2], ld(0x2A, f([1, 2, 3, 4])))),
        "count_mismatch": entry("a", tproto(1, [3], ld(0x2A, f([1, 2])))),
        "tensor_content_only": entry("a", b"\x08\x01" + ld(0x12, shape(2)) + ld(0x22, f([1, 2]))),
        "scalar_broadcast": entry("a", tproto(1, [4], ld(0x2A, f([7])))),
        "rank0": entry("a", tproto(1, [], ld(0x2A, f([7])))),
        "zero_dim": entry("a", tproto(1, [0, 3], b"")),
        "dtype_absent": entry("a", ld(0x12, shape(1)) + ld(0x2A, f([1]))),
        "dtype_bfloat16_unmapped": entry("a", tproto(14, [1], ld(0x6A, vi(16256)))),
        "dtype_half_ref_quirk": entry("a", tproto(19, [2], ld(0x6A, vi(18688) + vi(19712)))),
        "truncated": (entry("scores", tproto(1, [1000], ld(0x2A, scores.tobytes()))))[:-5],
        "packed_len_not_mult4": entry("a", tproto(1, [1], ld(0x2A, b"\x00\x00\x00"))),
        "bad_utf8_key": ld(0x0A, ld(0x0A, b"\xff\xfe") + ld(0x12, tproto(1, [1], ld(0x2A, f([1]))))),
        "tag_zero": b"\x00\x00",
        "wiretype_fixed64_unknown": entry("a", tproto(1, [1], ld(0x2A, f([1])))) + b"\xC1\x06" + b"\x00" * 8,
        "wiretype_fixed32_unknown": b"\xC5\x06" + b"\x00" * 4 + entry("a", tproto(1, [1], ld(0x2A, f([1])))),
        "group_unknown": b"\xC3\x06\xC4\x06" + entry("a", tproto(1, [1], ld(0x2A, f([1])))),
        "version_label_spec": entry("a", tproto(1, [1], ld(0x2A, f([1])))) + ld(0x12, ld(0x0A, b"m") + ld(0x22, b"stable")),
        "empty_model_spec": entry("a", tproto(1, [1], ld(0x2A, f([1])))) + b"\x12\x00",
        "entry_with_foreign_field": ld(0x0A, ld(0x0A, b"k") + b"\x08\x01" + ld(0x12, tproto(1, [1], ld(0x2A, f([4]))))) + entry("b", tproto(1, [1], ld(0x2A, f([5])))),
        "entry_with_unknown_number": ld(0x0A, ld(0x0A, b"k") + ld(0x12, tproto(1, [1], ld(0x2A, f([4])))) + b"\xB8\x06\x07"),
        "entry_key_wrong_wiretype": ld(0x0A, b"\x08\x05" + ld(0x12, tproto(1, [1], ld(0x2A, f([4]))))),
        "resource_handle_ok": entry("a", tproto(1, [1], ld(0x2A, f([4])) + ld(0x72, ld(0x0A, b"/dev") + ld(0x32, b"\x08\x01" + ld(0x12, shape(2)))))),
        "resource_handle_malformed": entry("a", tproto(1, [1], ld(0x2A, f([4])) + ld(0x72, b"\x00"))),
        "resource_handle_bad_utf8": entry("a", tproto(1, [1], ld(0x2A, f([4])) + ld(0x72, ld(0x0A, b"\xff")))),
        "variant_nested_ok": entry("a", tproto(1, [1], ld(0x2A, f([4])) + ld(0x7A, ld(0x0A, b"T") + ld(0x12, b"\xff\x00") + ld(0x1A, tproto(1, [1], ld(0x2A, f([9]))))))),
        "variant_nested_malformed": entry("a", tproto(1, [1], ld(0x2A, f([4])) + ld(0x7A, ld(0x1A, tproto(1, [1], b"\x2A\x03\x00\x00\x00"))))),
        "complex64": entry("a", tproto(8, [1], ld(0x4A, f([1, 2])))),
        # layouts the reference reads like any other (it iterates the merged repeated field, tensors.py:42-46) and that go past
        # the decode table's inline arrays: long rows of unpacked elements, many packed occurrences, many outputs, deep shapes
        "all_unpacked_1000": entry("a", tproto(1, [10, 100], b"".join(b"\x2D" + f([i * 0.25 - 3]) for i in range(1000)))) + mspec(),
        "all_unpacked_ints_300": entry("a", tproto(9, [300], b"".join(b"\x50" + vi((i * 7919) % 100000 - 500) for i in range(300)))),
        "split_packed_x20": entry("a", tproto(1, [sum(1 + 3 * (k % 5) for k in range(20))],
                                              b"".join(ld(0x2A, f(np.arange(1 + 3 * (k % 5)) + 100 * k)) for k in range(20)))),
        "split_packed_ints_x20": entry("a", tproto(3, [sum(2 + k % 4 for k in range(20))],
                                                   b"".join(ld(0x3A, b"".join(vi(((-1) ** j) * (k * 1000 + j) ** 2) for j in range(2 + k % 4))) for k in range(20)))),
        "unpacked_between_foreign_runs": entry("d", tproto(2, [30], b"".join(
            ld(0x2A, f([k] * (k % 3 + 1))) + b"\x31" + np.float64(k + 0.5).tobytes() + ld(0x32, np.array([k, -k], dtype=np.float64).tobytes())
            for k in range(10)))),
        "outputs_x40": b"".join(entry("out_%02d" % k, tproto(1, [2], ld(0x2A, f([k, -k])))) for k in range(40)) + mspec(),
        "rank_20": entry("a", tproto(1, [1, 2, 1, 1, 3, 1, 1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1, 1, 5, 1], ld(0x2A, f(np.arange(60))))),
        "rank_20_ints_dim_minus_one": entry("a", tproto(1, [1, 2, 1, 1, 3, 1, 1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1, 1, -1, 1], ld(0x2A, f(np.arange(60))))),
    }
    out = {}
    for name, wire in cases.items():
        rec = {"wire": wire.hex() if len(wire) <= 8192 else None}
        if rec["wire"] is None:
            rec["wire_recipe"] = name  # rebuilt by tests/golden_util.py from the same seeds
            rec["wire_sha256"] = hashlib.sha256(wire).hexdigest()
            rec["wire_len"] = len(wire)
        try:
            resp = PredictResponse.FromString(wire)
        except Exception as e:  # noqa: BLE001
            rec["parse_raises"] = type(e).__name__
            out[name] = rec
            continue
        rec["model_spec"] = {"name": resp.model_spec.name, "version": resp.model_spec.version.value,
                             "has_version": resp.model_spec.HasField("version"),
                             "version_label": resp.model_spec.version_label,
                             "signature_name": resp.model_spec.signature_name}
        outs = {}
        for k in sorted(resp.outputs.keys()):
            try:
                arr = dec(resp.outputs[k])
                if arr.dtype.kind == "U":
                    outs[k] = {"strings": [s for s in arr.ravel().tolist()], "shape": list(arr.shape), "dtype": "str"}
                elif arr.nbytes <= 8192:
                    outs[k] = {"dtype": arr.dtype.str, "shape": list(arr.shape), "data": arr.tobytes().hex()}
                else:
                    outs[k] = {"dtype": arr.dtype.str, "shape": list(arr.shape), "sha256": hashlib.sha256(arr.tobytes()).hexdigest()}
            except Exception as e:  # noqa: BLE001
                outs[k] = {"raises": type(e).__name__, "message": str(e)[:120]}
        rec["outputs"] = outs
        out[name] = rec
    return out


def main():
    meta = {"generator": "tests/golden/make_golden.py", "reference": "zendesk/min-tfs-client v1.0.2 (/root/reference, unmodified)",
            "protobuf": google.protobuf.__version__, "numpy": np.__version__}
    for fname, payload in (("encode.json", tensor_cases()), ("requests.json", request_cases()), ("decode.json", decode_cases())):
        with open(os.path.join(HERE, fname), "w") as fh:
            json.dump({"_meta": meta, "cases": payload}, fh, indent=1, sort_keys=True)
            fh.write("\n")
        print(fname, len(payload), "cases")


if __name__ == "__main__":
    main()
