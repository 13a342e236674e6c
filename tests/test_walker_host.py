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
    L.wh_parse_response.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.POINTER(N.Output), C.POINTER(C.c_int), C.POINTER(N.ModelSpec),
                                    C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    assert L.wh_sizeof_output() == C.sizeof(N.Output), "ctypes mirror of b200tfs_output is out of date"
    assert L.wh_sizeof_spill_entry() == C.sizeof(SpillEntry) == 32
    return L


class SpillEntry(C.Structure):      # walker.h
    _fields_ = [("kind", C.c_uint32), ("seq", C.c_uint32), ("run", N.Run)]


def parse(lib, wire, max_outputs=16, spill_cap=0):
    """(status, outs, n, spec, spill entries the record wanted, the area)"""
    outs = (N.Output * (max_outputs + 1))()
    n, spec, used = C.c_int(0), N.ModelSpec(), C.c_uint32(0)
    area = (SpillEntry * max(spill_cap, 1))()
    st = lib.wh_parse_response(wire, len(wire), max_outputs, outs, C.byref(n), C.byref(spec), area if spill_cap else None, spill_cap, C.byref(used))
    return st, outs, n.value, spec, used.value, area


def all_runs(o, area, used):
    """inline runs + the spilled ones of this output's map entry and value field, in wire order"""
    runs = [o.runs[k] for k in range(o.n_inline)]
    runs += [area[i].run for i in range(used) if area[i].kind == 2 and area[i].seq == o.spill_seq and area[i].run.field == o.value_field]
    assert len(runs) == o.n_runs
    return runs


def all_dims(o, area, used):
    dims = [o.dims[k] for k in range(min(o.rank, N.MAX_RANK))]
    dims += [area[i].run.off for i in range(used) if area[i].kind == 1 and area[i].seq == o.spill_seq]
    assert len(dims) == o.rank
    return dims


def _varints(b):
    out, v, sh = [], 0, 0
    for x in b:
        v |= (x & 0x7F) << sh
        sh += 7
        if not x & 0x80:
            out.append(v & (2 ** 64 - 1))
            v, sh = 0, 0
    return out


def _values(wire, o, runs=None, dims=None):
    """numpy array for one tabulated output, applying the reference's element semantics."""
    runs = [o.runs[k] for k in range(o.n_runs)] if runs is None else runs
    raw = b"".join(wire[r.off + q * r.stride: r.off + q * r.stride + r.len] for r in runs for q in range(r.count))
    np_type = numpy_for_enum(o.dtype)
    shape = tuple(o.dims[k] for k in range(o.rank)) if dims is None else tuple(dims)
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
    st, outs, n, spec, used, area = parse(lib, wire, 64, 4096)
    if "parse_raises" in rec:
        assert st == N.E_PARSE
        return
    assert st == N.OK
    table = {wire[outs[i].key_off: outs[i].key_off + outs[i].key_len].decode(): outs[i] for i in range(n)}
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
            assert o.dtype == 7 and o.n_strings == int(np.prod(exp["shape"])) and all_dims(o, area, used) == exp["shape"]
            continue
        if name == "dtype_half_ref_quirk":
            assert _varints(wire[o.runs[0].off: o.runs[0].off + o.runs[0].len]) == [18688, 19712]
            continue
        got = _values(wire, o, all_runs(o, area, used), all_dims(o, area, used))
        assert got.dtype.str == exp["dtype"] and list(got.shape) == exp["shape"]
        if "data" in exp:
            assert got.tobytes().hex() == exp["data"]
        else:
            assert hashlib.sha256(got.tobytes()).hexdigest() == exp["sha256"]


def test_walker_outputs_limit_and_groups(lib):
    f1 = G.ld(0x2A, np.float32([1]).tobytes())
    many = b"".join(G.entry("k%d" % i, G.tproto(1, [1], f1)) for i in range(20))
    assert parse(lib, many, 16)[0] == N.E_SIZE            # the caller's table is too small: the Python layer doubles it
    st, outs, n, *_ = parse(lib, many, 32)
    assert st == N.OK and n == 20
    wire = G.entry("a", G.tproto(1, [1], f1))
    # deep group nesting is refused, shallow nesting skipped
    deep = b"\xC3\x06" * 17 + b"\xC4\x06" * 17
    assert parse(lib, deep)[0] == N.E_PARSE
    ok = b"\xC3\x06" * 3 + b"\xC4\x06" * 3 + wire
    assert parse(lib, ok)[0] == N.OK


def test_unpacked_elements_coalesce_into_one_strided_run(lib):
    """A field written element by element (tag + value, tag + value, ...) is ONE run however long: count n, stride = tag + value."""
    vals = np.arange(1000, dtype=np.float32) * 0.5
    tp = G.tproto(1, [1000], b"".join(b"\x2D" + v.tobytes() for v in vals))
    wire = G.entry("a", tp)
    st, outs, n, spec, used, area = parse(lib, wire)
    o = outs[0]
    assert st == N.OK and o.status == N.OK and used == 0
    assert (o.n_runs, o.n_inline) == (1, 1) and (o.runs[0].len, o.runs[0].count, o.runs[0].stride) == (4, 1000, 5)
    assert o.flags & N.OF_UNPACKED and o.flags & N.OF_MULTI_CHUNK and not o.flags & N.OF_SPILLED
    assert _values(wire, o).tobytes() == vals.tobytes()
    # unpacked varints of equal length coalesce too; a change of length starts a new run
    ints = [1, 2, 3, 300, 301, 5]
    tp = G.tproto(9, [6], b"".join(b"\x50" + G.vi(v) for v in ints))
    wire = G.entry("a", tp)
    st, outs, n, spec, used, area = parse(lib, wire)
    o = outs[0]
    assert st == N.OK and o.status == N.OK and o.n_runs == 3
    assert [(o.runs[k].len, o.runs[k].count, o.runs[k].stride) for k in range(3)] == [(1, 3, 2), (2, 2, 3), (1, 1, 0)]
    assert _values(wire, o).tolist() == ints
    # a packed occurrence in between breaks the row; long packed occurrences never coalesce (they keep the tiled copy path)
    tp = G.tproto(1, [5], b"\x2D" + np.float32([1]).tobytes() + b"\x2D" + np.float32([2]).tobytes() + G.ld(0x2A, np.float32([3, 4]).tobytes())
                  + b"\x2D" + np.float32([5]).tobytes())
    wire = G.entry("a", tp)
    o = parse(lib, wire)[1][0]
    assert o.status == N.OK and [(o.runs[k].len, o.runs[k].count) for k in range(o.n_runs)] == [(4, 2), (8, 1), (4, 1)]
    big = np.arange(64, dtype=np.float32)
    tp = G.tproto(1, [128], G.ld(0x2A, big.tobytes()) + G.ld(0x2A, big.tobytes()))
    o = parse(lib, G.entry("a", tp))[1][0]
    assert o.status == N.OK and o.n_runs == 2 and o.runs[0].count == 1 and o.runs[1].count == 1


def test_spill_more_runs_than_the_table_holds(lib):
    """20 packed occurrences of different lengths: 8 runs inline, 12 in the spill area; without room the record says how much
    it needs (E_SPILL + used), with room it decodes like the canonical layout."""
    rng = np.random.default_rng(5)
    parts = [rng.standard_normal(1 + 3 * (k % 5)).astype(np.float32) for k in range(20)]
    vals = np.concatenate(parts)
    tp = G.tproto(1, [vals.size], b"".join(G.ld(0x2A, p.tobytes()) for p in parts))
    wire = G.entry("a", tp) + G.mspec()
    st, outs, n, spec, used, area = parse(lib, wire, 16, 0)
    assert st == N.E_SPILL and used == 12
    st, outs, n, spec, used, area = parse(lib, wire, 16, 4)
    assert st == N.E_SPILL and used == 12                      # the count is exact even when the area overflows
    st, outs, n, spec, used, area = parse(lib, wire, 16, used)
    o = outs[0]
    assert st == N.OK and o.status == N.OK and (o.n_runs, o.n_inline) == (20, 8) and o.flags & N.OF_SPILLED and used == 12
    assert _values(wire, o, all_runs(o, area, used)).tobytes() == vals.tobytes()
    # runs of ANOTHER field in between (float_val while the dtype says double) are dropped on both sides of the table's edge
    dbl = [rng.standard_normal(2 + k % 3) for k in range(12)]
    body = b""
    for k in range(12):
        body += G.ld(0x2A, np.float32([k] * (k % 3 + 1)).tobytes()) + G.ld(0x32, dbl[k].tobytes())
    want = np.concatenate(dbl)
    wire = G.entry("d", G.tproto(2, [want.size], body))
    st, outs, n, spec, used, area = parse(lib, wire, 16, 64)
    o = outs[0]
    assert st == N.OK and o.status == N.OK and o.n_runs == 12 and o.n_inline == 4 and used == 16
    assert _values(wire, o, all_runs(o, area, used)).tobytes() == want.tobytes()
    # a later entry with the same key replaces the earlier one: only ITS spill entries count
    wire2 = wire + G.entry("d", G.tproto(2, [want.size], body))
    st, outs, n, spec, used, area = parse(lib, wire2, 16, 64)
    o = outs[0]
    assert st == N.OK and n == 1 and o.spill_seq == 1 and used == 32
    assert _values(wire2, o, all_runs(o, area, used)).tobytes() == want.tobytes()


def test_spill_rank_beyond_the_table(lib):
    dims = [1, 2, 1, 1, 3, 1, 1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1, 1, 5, 1]      # rank 20
    vals = np.arange(int(np.prod(dims)), dtype=np.float32)
    wire = G.entry("a", G.tproto(1, dims, G.ld(0x2A, vals.tobytes())))
    st, outs, n, spec, used, area = parse(lib, wire, 16, 0)
    assert st == N.E_SPILL and used == 4
    st, outs, n, spec, used, area = parse(lib, wire, 16, 8)
    o = outs[0]
    assert st == N.OK and o.status == N.OK and o.rank == 20 and o.flags & N.OF_SPILLED and o.n_elems == vals.size
    assert all_dims(o, area, used) == dims
    # the inferred dim may be one of the spilled ones
    dims2 = list(dims)
    dims2[18] = -1
    wire = G.entry("a", G.tproto(1, dims2, G.ld(0x2A, vals.tobytes())))
    st, outs, n, spec, used, area = parse(lib, wire, 16, 8)
    o = outs[0]
    assert st == N.OK and o.status == N.OK and o.flags & N.OF_DIM_INFERRED and all_dims(o, area, used) == dims
