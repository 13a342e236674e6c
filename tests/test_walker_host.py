"""The decode kernels' tag walker, compiled for the host, against the reference's golden decode cases.

CPU only: checks the table the walker builds (status, dtype, dims, where the values lie) and, by
slicing the wire at the tabulated offsets with numpy, the values themselves.  The GPU tests check
the same cases end to end through the kernels.
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

import golden_util as G
from min_tfs_client import _native as N
from min_tfs_client.constants import numpy_for_enum

HERE = os.path.dirname(os.path.abspath(__file__))
DEC = G.load("decode.json")


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(HERE, "native", "_build", "libwalker_host.so")
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "native")], check=True)
    L = C.CDLL(so)
    L.wh_parse_response.restype = C.c_int
    L.wh_parse_response.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.POINTER(N.Output), C.POINTER(C.c_int), C.POINTER(N.ModelSpec)]
    assert L.wh_sizeof_output() == C.sizeof(N.Output), "ctypes mirror of b200tfs_output is out of date"
    return L


def _varints(b):
    out, v, sh = [], 0, 0
    for x in b:
        v |= (x & 0x7F) << sh
        sh += 7
        if not x & 0x80:
            out.append(v & (2 ** 64 - 1))
            v, sh = 0, 0
    return out


def _values(wire, o):
    """numpy array for one tabulated output, applying the reference's element semantics."""
    raw = b"".join(wire[o.chunk_off[k]: o.chunk_off[k] + o.chunk_len[k]] for k in range(o.n_chunks))
    np_type = numpy_for_enum(o.dtype)
    shape = tuple(o.dims[k] for k in range(o.rank))
    if o.flags & N.OF_VARINT:
        vals = _varints(raw)
        if o.value_field in (7, 13):      # int32 fields truncate
            vals = [((v & 0xFFFFFFFF) ^ 0x80000000) - 0x80000000 for v in vals]
        elif o.value_field == 16:
            vals = [v & 0xFFFFFFFF for v in vals]
        elif o.value_field == 10:
            vals = [(v ^ (1 << 63)) - (1 << 63) for v in vals]
        elif o.value_field == 11:
            vals = [v != 0 for v in vals]
        if len(vals) != o.n_elems:
            raise ValueError("count")
        return np.array(vals, dtype=np_type).reshape(shape)
    a = np.frombuffer(raw, dtype=np_type).copy()
    if np_type is np.float32:             # float32 passes through a double: signalling NaNs are quieted
        u = a.view(np.uint32)
        u[((u & 0x7FFFFFFF) > 0x7F800000)] |= 0x00400000
    return a.reshape(shape)


@pytest.mark.parametrize("name", list(DEC))
def test_walker_against_golden(lib, name):
    rec = DEC[name]
    wire = G.decode_case_wire(name, rec)
    outs = (N.Output * 17)()
    n = C.c_int(0)
    spec = N.ModelSpec()
    st = lib.wh_parse_response(wire, len(wire), 16, outs, C.byref(n), C.byref(spec))
    if "parse_raises" in rec:
        assert st == N.E_PARSE
        return
    assert st == N.OK
    table = {wire[outs[i].key_off: outs[i].key_off + outs[i].key_len].decode(): outs[i] for i in range(n.value)}
    assert set(table) == set(rec["outputs"])
    ms = rec["model_spec"]
    assert wire[spec.name_off: spec.name_off + spec.name_len].decode() == ms["name"]
    assert (spec.version, bool(spec.has_version)) == (ms["version"], ms["has_version"])
    assert wire[spec.label_off: spec.label_off + spec.label_len].decode() == ms["version_label"]
    assert wire[spec.signature_off: spec.signature_off + spec.signature_len].decode() == ms["signature_name"]
    for key, exp in rec["outputs"].items():
        o = table[key]
        if "raises" in exp:
            kind = exp["raises"]
            if kind == "KeyError":
                assert o.status == N.E_KEY or o.dtype == 14      # bfloat16 is mapped here, the reference has no row for it
            elif kind == "TypeError":
                assert o.flags & N.OF_RANK0 and o.status == N.OK
            elif kind == "OverflowError":
                with pytest.raises(OverflowError):
                    _values(wire, o)
            elif kind == "UnicodeDecodeError":
                assert o.dtype == 7 and o.status == N.OK
            elif o.dtype in (8, 18):
                assert o.status == N.OK                           # complex: TF pairs here, the reference cannot reshape
            else:
                assert o.status == N.E_SHAPE, (name, o.status)
            continue
        assert o.status == N.OK, (name, key, o.status)
        if exp["dtype"] == "str":
            assert o.dtype == 7 and o.n_strings == int(np.prod(exp["shape"])) and [o.dims[k] for k in range(o.rank)] == exp["shape"]
            continue
        if name == "dtype_half_ref_quirk":
            assert _varints(wire[o.chunk_off[0]: o.chunk_off[0] + o.chunk_len[0]]) == [18688, 19712]
            continue
        got = _values(wire, o)
        assert got.dtype.str == exp["dtype"] and list(got.shape) == exp["shape"]
        if "data" in exp:
            assert got.tobytes().hex() == exp["data"]
        else:
            assert hashlib.sha256(got.tobytes()).hexdigest() == exp["sha256"]


def test_walker_table_limits(lib):
    f1 = G.ld(0x2A, np.float32([1]).tobytes())
    many = b"".join(G.entry("k%d" % i, G.tproto(1, [1], f1)) for i in range(20))
    outs = (N.Output * 17)()
    n = C.c_int(0)
    spec = N.ModelSpec()
    assert lib.wh_parse_response(many, len(many), 16, outs, C.byref(n), C.byref(spec)) == N.E_SIZE
    # nine unpacked elements exceed the eight tabulated chunks: flagged, not mis-decoded
    tp = G.tproto(1, [9], b"".join(b"\x2D" + np.float32([i]).tobytes() for i in range(9)))
    wire = G.entry("a", tp)
    assert lib.wh_parse_response(wire, len(wire), 16, outs, C.byref(n), C.byref(spec)) == N.OK
    assert outs[0].status == N.E_NONCANONICAL
    # rank 17 exceeds the table
    tp = G.tproto(1, [1] * 17, f1)
    wire = G.entry("a", tp)
    assert lib.wh_parse_response(wire, len(wire), 16, outs, C.byref(n), C.byref(spec)) == N.OK
    assert outs[0].status == N.E_NONCANONICAL
    # deep group nesting is refused, shallow nesting skipped
    deep = b"\xC3\x06" * 17 + b"\xC4\x06" * 17
    assert lib.wh_parse_response(deep, len(deep), 16, outs, C.byref(n), C.byref(spec)) == N.E_PARSE
    ok = b"\xC3\x06" * 3 + b"\xC4\x06" * 3 + wire
    assert lib.wh_parse_response(ok, len(ok), 16, outs, C.byref(n), C.byref(spec)) == N.OK
