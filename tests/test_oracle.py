"""Pin the oracle (C restatement + Python port) against the vectors the unmodified reference produced.

CPU only.  If these fail the oracle is wrong and no GPU parity claim that leans on it means anything.
"""
import hashlib

import numpy as np
import pytest

import golden_util as G
from oracle import ref_port, wire_oracle

ENC = G.load("encode.json")
REQ = G.load("requests.json")
DEC = G.load("decode.json")

_BIG = {"c2_f32_1024x1024", "c2_f32_random_bits", "f64_512x300", "c3_image"}  # per-element Python: keep the port to small cases


@pytest.mark.parametrize("name", [k for k in ENC if not k.startswith("_")])
def test_c_oracle_tensor_proto(name):
    case = ENC[name]
    x = G.apply_transform(G.make_array(case["input"]), case.get("transform"))
    G.check_wire(wire_oracle.encode_tensor_proto(x), case["wire"], name)


@pytest.mark.parametrize("name", [k for k in ENC if not k.startswith("_") and k not in _BIG])
def test_port_tensor_proto(name):
    case = ENC[name]
    x = G.apply_transform(G.make_array(case["input"]), case.get("transform"))
    G.check_wire(ref_port.encode_tensor_proto(x), case["wire"], name)


def _has_strings(case):
    return any(r["gen"] == "strings" for _, r in case["inputs"])


@pytest.mark.parametrize("name", [k for k in REQ if not _has_strings(REQ[k])])
def test_c_oracle_request(name):
    case = REQ[name]
    inputs = [(k, G.make_array(r)) for k, r in case["inputs"]]
    wd = np.float32 if case.get("wire_dtype") == "DT_FLOAT" else None
    G.check_wire(wire_oracle.encode_predict_request(case["model_name"], case["model_version"], inputs, wire_dtype=wd), case["wire"], name)


@pytest.mark.parametrize("name", [k for k in REQ if "c4_" not in k and k not in ("kat3_c2", "kat4_c3_req0", "c3_req7", "c5_req3")])
def test_port_request(name):
    case = REQ[name]
    inputs = [(k, G.make_array(r)) for k, r in case["inputs"]]
    G.check_wire(ref_port.encode_predict_request(case["model_name"], case["model_version"], inputs), case["wire"], name)


def test_port_one_c3_request():
    case = REQ["c3_req7"]
    inputs = [(k, G.make_array(r)) for k, r in case["inputs"]]
    G.check_wire(ref_port.encode_predict_request(case["model_name"], case["model_version"], inputs), case["wire"], "c3_req7")


_EXC = {"ValueError": ValueError, "KeyError": KeyError, "TypeError": TypeError, "OverflowError": OverflowError}


@pytest.mark.parametrize("name", list(DEC))
def test_c_oracle_decode(name):
    rec = DEC[name]
    wire = G.decode_case_wire(name, rec)
    if "parse_raises" in rec:
        with pytest.raises(wire_oracle.ParseError):
            wire_oracle.decode_predict_response(wire)
        return
    expected = rec["outputs"]
    if any(v.get("dtype") == "str" or v.get("raises") == "UnicodeDecodeError" for v in expected.values()):
        pytest.skip("string outputs are host objects; covered by the port")
    raising = [v for v in expected.values() if "raises" in v]
    if raising:
        with pytest.raises(_EXC[raising[0]["raises"]]):
            wire_oracle.decode_predict_response(wire)
        return
    outs, spec = wire_oracle.decode_predict_response(wire, with_spec=True)
    assert set(outs) == set(expected)
    for k, v in expected.items():
        got = outs[k]
        assert got.dtype.str == v["dtype"] and list(got.shape) == v["shape"], (k, got.dtype, got.shape)
        if "data" in v:
            assert got.tobytes().hex() == v["data"], k
        else:
            assert hashlib.sha256(got.tobytes()).hexdigest() == v["sha256"], k
    assert spec == rec["model_spec"]


@pytest.mark.parametrize("name", [k for k in DEC if k != "c2_response"])
def test_port_decode(name):
    from google.protobuf.message import DecodeError

    rec = DEC[name]
    wire = G.decode_case_wire(name, rec)
    if "parse_raises" in rec:
        with pytest.raises(DecodeError):
            ref_port.decode_predict_response(wire)
        return
    expected = rec["outputs"]
    raising = [v for v in expected.values() if "raises" in v]
    if raising:
        with pytest.raises(Exception) as ei:
            ref_port.decode_predict_response(wire)
        assert type(ei.value).__name__ == raising[0]["raises"]
        return
    outs = ref_port.decode_predict_response(wire)
    for k, v in expected.items():
        got = outs[k]
        if v["dtype"] == "str":
            assert got.ravel().tolist() == v["strings"]
        else:
            assert got.dtype.str == v["dtype"] and list(got.shape) == v["shape"] and got.tobytes().hex() == v["data"], k


def test_oracle_round_trip_property():
    """oracle encode -> oracle decode is the identity on every numeric dtype (incl. varint edge values)."""
    rng = np.random.default_rng(7)
    for dt in (np.float32, np.float64, np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64, np.bool_):
        if dt is np.bool_:
            x = rng.integers(0, 2, size=(5, 9)).astype(np.bool_)
        elif np.dtype(dt).kind == "f":
            x = rng.standard_normal((5, 9)).astype(dt)
        else:
            info = np.iinfo(dt)
            x = rng.integers(info.min, info.max, size=(5, 9), dtype=dt, endpoint=True)
        back = wire_oracle.decode_tensor_proto(wire_oracle.encode_tensor_proto(x))
        assert back.dtype == x.dtype and back.tobytes() == x.tobytes()
    resp = wire_oracle.build_predict_response([("a", np.arange(6, dtype=np.float32).reshape(2, 3)), ("b", np.array([1, -1], dtype=np.int64))])
    outs = wire_oracle.decode_predict_response(resp)
    assert outs["a"].shape == (2, 3) and outs["b"].tolist() == [1, -1]
    assert ref_port.decode_predict_response(resp)["a"].tobytes() == outs["a"].tobytes()


def _padding_cases():
    """TensorProtos with fewer typed values than the shape holds (TF writes constants this way), built with the runtime."""
    from tensorflow.core.framework import tensor_pb2

    def tp(dtype, shape, **fields):
        m = tensor_pb2.TensorProto(dtype=dtype)
        for d in shape:
            m.tensor_shape.dim.add().size = d
        for k, v in fields.items():
            getattr(m, k).extend(v)
        return m.SerializeToString()

    return {
        "f32_broadcast": tp(1, [4, 3], float_val=[1.5]),
        "f32_three_of_twelve": tp(1, [4, 3], float_val=[1.0, -2.0, 3.25]),
        "f32_none": tp(1, [2, 5]),
        "f64_edge": tp(2, [7], double_val=[0.5, 0.25]),
        "i64_broadcast": tp(9, [3, 3], int64_val=[-7]),
        "i8_two": tp(6, [5], int_val=[-3, 5]),
        "u64_big": tp(23, [4], uint64_val=[2 ** 64 - 1]),
        "bool_one": tp(10, [6], bool_val=[True]),
        "half_bits": tp(19, [4], half_val=[0x3C00, 0x4000]),
        "c64_pair": tp(8, [3], scomplex_val=[1.0, 2.0]),
        "i32_none": tp(3, [3]),
        "i64_large": tp(9, [20000], int64_val=list(range(-50, 50))),
        "f32_large": tp(1, [300, 100], float_val=[float(i) for i in range(7)]),
    }


def test_tolerant_padding_follows_tensorflow():
    """strict=False on fewer values than elements: TensorFlow's MakeNdarray rule (zeros / repeat the last value), restated in
    oracle/ref_port.make_ndarray_tf from the vendored tensor_util.py; strict=True keeps the reference's ValueError."""
    for name, w in _padding_cases().items():
        want = ref_port.make_ndarray_tf(w)
        got = wire_oracle.decode_tensor_proto(w, strict=False)
        assert got.dtype == want.dtype and got.shape == want.shape and got.tobytes() == want.tobytes(), name
        with pytest.raises(ValueError):
            wire_oracle.decode_tensor_proto(w, strict=True)
    # more values than the shape holds is an error in both modes
    from tensorflow.core.framework import tensor_pb2

    m = tensor_pb2.TensorProto(dtype=1, float_val=[1.0, 2.0, 3.0])
    m.tensor_shape.dim.add().size = 2
    for strict in (True, False):
        with pytest.raises(ValueError):
            wire_oracle.decode_tensor_proto(m.SerializeToString(), strict=strict)
