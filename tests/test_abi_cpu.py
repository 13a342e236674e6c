"""CPU-only checks of the C ABI: the library loads, exports every symbol include/b200tfs.h declares,
fails loudly without a GPU, and its host-side planner (sizes, framing bytes, key order) reproduces
the reference's wire bytes for every golden vector - no kernel runs here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import golden_util as G
from min_tfs_client import _native as N
from min_tfs_client.codec import _Prepared

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENC = G.load("encode.json")
REQ = G.load("requests.json")


def test_header_symbols_all_exported_and_bound():
    hdr = open(os.path.join(REPO, "include", "b200tfs.h")).read()
    declared = set(re.findall(r"\b(b200tfs_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"b200tfs_ctx"}
    lib = N.load()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"declared in b200tfs.h but not exported: {missing}"
    unbound = sorted(declared - set(N.SIGNATURES))
    assert not unbound, f"declared in b200tfs.h but not bound in _native.SIGNATURES: {unbound}"
    extra = sorted(set(N.SIGNATURES) - declared)
    assert not extra, f"bound but not declared in b200tfs.h: {extra}"
    assert lib.b200tfs_abi_version() == 2


def test_struct_mirrors_match_header_layout():
    # sizes follow from the header's field lists (natural alignment)
    assert C.sizeof(N.Tensor) == 56 and C.sizeof(N.Request) == 48
    assert C.sizeof(N.Output) == 32 + 8 * N.MAX_RANK + 24 * N.MAX_RUNS + 32 + 32 + 16 and C.sizeof(N.Run) == 24
    assert C.sizeof(N.ModelSpec) == 48


@pytest.mark.skipif(N.device_count() > 0, reason="this box has a GPU")
def test_no_cpu_fallback_without_gpu():
    """The product must fail loudly when there is no CUDA device - there is no CPU codec behind it."""
    from min_tfs_client import tensors
    from min_tfs_client.codec import Codec

    with pytest.raises(RuntimeError, match="no CPU path"):
        Codec(0)
    with pytest.raises(RuntimeError):
        tensors.ndarray_to_tensor_proto(np.zeros(3, np.float32))
    ctx = C.c_void_p()
    assert N.load().b200tfs_create(0, C.byref(ctx)) == N.E_CUDA


def test_dtype_table():
    lib = N.load()
    rows = {1: (5, 4), 2: (6, 8), 3: (7, 4), 4: (7, 1), 5: (7, 2), 6: (7, 1), 7: (8, 0), 8: (9, 8), 9: (10, 8), 10: (11, 1), 14: (13, 2),
            17: (7, 2), 18: (12, 16), 19: (13, 2), 22: (16, 4), 23: (17, 8)}
    for dt, (field, size) in rows.items():
        assert lib.b200tfs_dtype_field(dt) == field and lib.b200tfs_dtype_size(dt) == size, dt
    for dt in (0, 11, 12, 13, 15, 16, 20, 21, 101):
        assert lib.b200tfs_dtype_field(dt) == 0
    assert lib.b200tfs_cast_supported(19, 1) and lib.b200tfs_cast_supported(14, 1) and not lib.b200tfs_cast_supported(1, 2)


def _payload_bytes(arr, wire_dtype=None):
    """Expected payload of the typed field for fixed-width dtypes, built with numpy (test-side)."""
    a = np.ascontiguousarray(arr)
    if wire_dtype is not None:
        a = a.astype(np.float32)
    if a.dtype == np.float32:
        u = a.view(np.uint32).copy()
        u[(u & 0x7FFFFFFF) > 0x7F800000] |= 0x00400000
        return u.tobytes()
    if a.dtype == np.bool_:
        return (a.view(np.uint8) != 0).astype(np.uint8).tobytes()
    return a.tobytes()


_FIXED = (np.float32, np.float64, np.bool_, np.complex64, np.complex128)


@pytest.mark.parametrize("name", [k for k in ENC if not k.startswith("_")])
def test_tensor_proto_header_matches_golden(name):
    case = ENC[name]
    x = G.apply_transform(G.make_array(case["input"]), case.get("transform"))
    if x.dtype.kind == "U":
        pytest.skip("DT_STRING protos are assembled on the host by protobuf")
    lib = N.load()
    p = _Prepared(x, b"", None, False, False)
    t = p.struct
    if p.array.dtype.type not in _FIXED and p.array.size:
        # varint dtypes: the packed length comes from the measure kernel; take it from the golden total instead
        hl, tl = C.c_uint64(), C.c_uint64()
        t.packed_len = 1
        N.check(lib.b200tfs_tensor_proto_size(C.byref(t), C.byref(hl), C.byref(tl)))
        guess = case["wire"]["len"] - (hl.value - 1)
        for delta in (0, -1, -2, -3, 1):   # the length prefix itself may need another byte
            t.packed_len = guess + delta
            N.check(lib.b200tfs_tensor_proto_size(C.byref(t), C.byref(hl), C.byref(tl)))
            if tl.value == case["wire"]["len"]:
                break
        assert tl.value == case["wire"]["len"]
    buf = C.create_string_buffer(4096)
    n = C.c_uint64()
    N.check(lib.b200tfs_tensor_proto_header(C.byref(t), buf, 4096, C.byref(n)))
    hl, tl = C.c_uint64(), C.c_uint64()
    N.check(lib.b200tfs_tensor_proto_size(C.byref(t), C.byref(hl), C.byref(tl)))
    assert hl.value == n.value and tl.value == case["wire"]["len"]
    header = buf.raw[: n.value]
    expect_head = bytes.fromhex(case["wire"].get("hex", case["wire"].get("head", "")))
    assert expect_head[: min(len(header), len(expect_head))] == header[: min(len(header), len(expect_head))]
    if p.array.dtype.type in _FIXED and "hex" in case["wire"]:
        assert header + _payload_bytes(p.array) == bytes.fromhex(case["wire"]["hex"])


@pytest.mark.parametrize("name", list(REQ))
def test_request_frame_matches_golden(name):
    """Host planner framing + numpy payloads == the reference's PredictRequest bytes (fixed-width inputs)."""
    case = REQ[name]
    inputs = [(k, G.make_array(r)) for k, r in case["inputs"]]
    if any(a.dtype.type not in _FIXED and a.dtype.type is not np.float16 and a.dtype.name != "bfloat16" for _, a in inputs):
        pytest.skip("varint / string inputs: framing depends on the measure kernel (covered on the GPU)")
    wd = "DT_FLOAT" if case.get("wire_dtype") == "DT_FLOAT" else None
    preps = [_Prepared(a, k.encode(), wd, False, False) for k, a in inputs]
    arr = (N.Tensor * max(len(preps), 1))(*[p.struct for p in preps])
    name_b = case["model_name"].encode()
    req = N.Request(model_name=name_b, model_name_len=len(name_b), has_version=int(case["model_version"] is not None), order=N.ORDER_UPB,
                    version=case["model_version"] or 0, n_inputs=len(preps), flags=0, inputs=arr)
    lib = N.load()
    total = C.c_uint64()
    N.check(lib.b200tfs_request_size(C.byref(req), C.byref(total)))
    assert total.value == case["wire"]["len"]
    cap = 1 << 16
    buf = C.create_string_buffer(cap)
    flen = C.c_uint64()
    n = max(len(preps), 1)
    poff, plen, perm = (C.c_uint64 * n)(), (C.c_uint64 * n)(), (C.c_int32 * n)()
    N.check(lib.b200tfs_request_frame(C.byref(req), buf, cap, C.byref(flen), poff, plen, perm))
    frame = buf.raw[: flen.value]
    wire = bytearray()
    fpos = 0
    for j in range(len(preps)):
        take = poff[j] - len(wire)
        wire += frame[fpos: fpos + take]
        fpos += take
        payload = _payload_bytes(preps[perm[j]].array, wd)
        assert len(payload) == plen[j]
        wire += payload
    wire += frame[fpos:]
    # the same request behind gRPC's length-prefixed-message header: five more bytes in front, every payload 5 further on
    req.flags = N.RF_GRPC_FRAME
    total5, flen5 = C.c_uint64(), C.c_uint64()
    N.check(lib.b200tfs_request_size(C.byref(req), C.byref(total5)))
    buf5 = C.create_string_buffer(cap)
    poff5, plen5, perm5 = (C.c_uint64 * n)(), (C.c_uint64 * n)(), (C.c_int32 * n)()
    N.check(lib.b200tfs_request_frame(C.byref(req), buf5, cap, C.byref(flen5), poff5, plen5, perm5))
    assert total5.value == total.value + 5 and flen5.value == flen.value + 5
    assert buf5.raw[: flen5.value] == b"\x00" + int(total.value).to_bytes(4, "big") + frame
    assert [poff5[j] for j in range(len(preps))] == [poff[j] + 5 for j in range(len(preps))]
    req.flags = 0x40
    assert lib.b200tfs_request_size(C.byref(req), C.byref(total5)) == N.E_ARG
    G.check_wire(bytes(wire), case["wire"], name)


def test_key_order_modes():
    lib = N.load()
    keys = [b"b", b"a", b"aa", b"ab", b"B", b"", b"abc"]
    arr = (C.c_char_p * len(keys))(*keys)
    lens = (C.c_int64 * len(keys))(*[len(k) for k in keys])
    perm = (C.c_int32 * len(keys))()
    N.check(lib.b200tfs_order_keys(len(keys), arr, lens, N.ORDER_UPB, perm))
    assert [keys[i] for i in perm] == [b"B", b"aa", b"abc", b"ab", b"a", b"b", b""]      # prefix sorts AFTER the longer key (upb)
    N.check(lib.b200tfs_order_keys(len(keys), arr, lens, N.ORDER_BYTES, perm))
    assert [keys[i] for i in perm] == sorted(keys)
    N.check(lib.b200tfs_order_keys(len(keys), arr, lens, N.ORDER_GIVEN, perm))
    assert list(perm) == list(range(len(keys)))


def test_size_errors():
    lib = N.load()
    dims = (C.c_int64 * 2)(1 << 20, 1 << 12)
    t = N.Tensor(data=256, src_dtype=1, wire_dtype=1, rank=2, flags=0, dims=dims, key=b"", key_len=0, packed_len=0)
    assert lib.b200tfs_tensor_proto_size(C.byref(t), None, None) == N.E_TOOBIG       # 16 GiB > protobuf's 2 GiB limit
    t.wire_dtype = 7
    assert lib.b200tfs_tensor_proto_size(C.byref(t), None, None) == N.E_DTYPE
    t.wire_dtype, t.rank = 1, 300
    assert lib.b200tfs_tensor_proto_size(C.byref(t), None, None) == N.E_SHAPE
    dims[0] = -1
    t.rank = 2
    assert lib.b200tfs_tensor_proto_size(C.byref(t), None, None) == N.E_SHAPE
    assert b"negative" in lib.b200tfs_last_error()


def test_header_is_plain_c99_and_the_library_links_from_c():
    """include/b200tfs.h through gcc -std=c99 -pedantic -Werror, linked against libb200tfs.so and run: struct sizes as the ctypes
    mirrors have them, and - no GPU here - b200tfs_create refusing loudly instead of falling back to a CPU path."""
    import subprocess

    native = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native")
    N.load()   # builds nothing, but fails early if the library is missing
    subprocess.run(["make", "-s", "-C", native, "_build/abi_c99"], check=True)
    out = subprocess.run([os.path.join(native, "_build", "abi_c99")], capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[0] == f"abi {N.load().b200tfs_abi_version()} sizes {C.sizeof(N.Tensor)} {C.sizeof(N.Request)} {C.sizeof(N.Output)} {C.sizeof(N.ModelSpec)}"
    if N.device_count() == 0:
        assert out[1].startswith(f"create {N.E_CUDA} ") and "no CPU path" in out[1]
