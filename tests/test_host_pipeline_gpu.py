"""The pipelined host-buffer entry points (b200tfs_encode_requests_host_async / b200tfs_decode_responses_host_async): a large
request or response is cut into slices that travel H2D, through the kernel and D2H on three streams.  Slicing must not change a
single byte: every case is checked against the CPU oracle, with the slice threshold lowered (B200TFS_PIPELINE_MIN) so that
small, quickly-checked tensors are sliced as well and the cuts fall at awkward places.
"""
import ctypes as C
import os

import numpy as np
import pytest

from devutil import Dev
from min_tfs_client import _native as N
from min_tfs_client.constants import enum_for_numpy
from oracle import wire_oracle

pytestmark = pytest.mark.gpu


def _dev_with_threshold(nbytes):
    old = os.environ.get("B200TFS_PIPELINE_MIN")
    os.environ["B200TFS_PIPELINE_MIN"] = str(nbytes)
    try:
        return Dev(0)          # the threshold is read when the context is created
    finally:
        if old is None:
            del os.environ["B200TFS_PIPELINE_MIN"]
        else:
            os.environ["B200TFS_PIPELINE_MIN"] = old


def _pipelined(dev):
    n = C.c_uint64()
    N.check(dev.lib.b200tfs_pipelined_calls(dev.ctx, C.byref(n)))
    return n.value


def _with_snan(x):
    flat = x.reshape(-1)
    if x.dtype == np.float32 and flat.size >= 8:
        bits = flat.view(np.uint32)
        for pos in (0, flat.size // 3, flat.size - 1):
            bits[pos] = 0x7F800001 if pos % 2 == 0 else 0xFFA00000
    return x


def _direct(dev):
    n = C.c_uint64()
    N.check(dev.lib.b200tfs_direct_calls(dev.ctx, C.byref(n)))
    return n.value


def _encode_host(dev, batch, pinned_inputs=True, wire_dtypes=None, grpc=False, pinned_wire=True):
    """batch: [(model, version, [(key, ndarray)])] with HOST arrays -> list of wires via b200tfs_encode_requests_host_async."""
    keep, reqs = [], []
    for bi, (model, version, inputs) in enumerate(batch):
        ts = []
        for ki, (k, a) in enumerate(inputs):
            a = np.ascontiguousarray(a)
            if pinned_inputs:
                pb = N.PinnedBuffer(max(a.nbytes, 1))
                pb.array[: a.nbytes] = a.reshape(-1).view(np.uint8)
                ptr = pb.ptr
                keep.append(pb)
            else:
                ptr = a.ctypes.data
                keep.append(a)
            dims = (C.c_int64 * max(a.ndim, 1))(*a.shape)
            e = enum_for_numpy(a.dtype)
            w = (wire_dtypes or {}).get((bi, ki), e)
            key = k.encode()
            ts.append(N.Tensor(data=ptr, src_dtype=e, wire_dtype=w, rank=a.ndim, flags=0, dims=dims, key=key, key_len=len(key), packed_len=0))
            keep += [dims, key]
        arr = (N.Tensor * max(len(ts), 1))(*ts)
        name = model.encode()
        keep += [arr, name]
        reqs.append(N.Request(model_name=name, model_name_len=len(name), has_version=int(version is not None), order=N.ORDER_UPB,
                              version=version or 0, n_inputs=len(ts), flags=N.RF_GRPC_FRAME if grpc else 0, inputs=arr))
    n = len(reqs)
    rq = (N.Request * n)(*reqs)
    cap = sum(sum(a.nbytes * 2 for _, a in inputs) + 4096 for _, _, inputs in batch) + 4096
    if pinned_wire:      # page-locked: the kernels write it themselves (no device-to-host copy); record 0 need not start at 0
        wire = N.PinnedBuffer(cap)
        arr, ptr = wire.array, wire.ptr
    else:                # pageable: staged on the device and copied back, record 0 at offset 0
        arr = np.empty(cap, dtype=np.uint8)
        ptr = arr.ctypes.data
    arr[:] = 0xEE
    off, ln = (C.c_uint64 * n)(), (C.c_uint64 * n)()
    N.check(dev.lib.b200tfs_encode_requests_host_async(dev.ctx, n, rq, ptr, cap, off, ln))
    dev.sync()
    if not pinned_wire:
        assert off[0] == 0
    return [arr[off[i]: off[i] + ln[i]].tobytes() for i in range(n)]


@pytest.mark.parametrize("elems", [65536 + 1, 131072 + 13, 200001, 262144, 1 << 20])
def test_sliced_encode_matches_the_oracle(elems):
    dev = _dev_with_threshold(4096)
    try:
        x = _with_snan(np.random.default_rng(elems).standard_normal(elems, dtype=np.float32))
        before, d0 = _pipelined(dev), _direct(dev)
        ref = wire_oracle.encode_predict_request("default", 3, [("x", x)])
        for pinned in (True, False):
            for pinned_wire in (True, False):
                assert _encode_host(dev, [("default", 3, [("x", x)])], pinned_inputs=pinned, pinned_wire=pinned_wire)[0] == ref
        assert _pipelined(dev) == before + 4 and _direct(dev) == d0 + 2
    finally:
        dev.close()


def test_sliced_encode_of_a_mixed_batch():
    """Several requests, several inputs each, payload sizes from a few bytes (small items: travel with slice 0) to megabytes, a
    widening cast (fp16 source -> DT_FLOAT on the wire: source bytes per output byte = 1/2), bools and int8 tensor_content-free
    fixed-width bytes; cuts fall inside and between tensors."""
    dev = _dev_with_threshold(4096)
    try:
        rng = np.random.default_rng(7)
        a = _with_snan(rng.standard_normal((300, 301), dtype=np.float32))
        h = rng.standard_normal(100003).astype(np.float16)
        b = rng.integers(0, 2, size=70001).astype(np.bool_)
        d = rng.standard_normal(40000)                      # float64
        tiny = np.arange(5, dtype=np.float32)
        batch = [("m1", 1, [("a", a), ("tiny", tiny)]), ("m2", None, [("h", h), ("b", b)]), ("m3", 7, [("d", d), ("a2", a[:17])])]
        wires = _encode_host(dev, batch, wire_dtypes={(1, 0): 1}, grpc=True)
        assert _pipelined(dev) == 1 and _direct(dev) == 1
        assert wires == _encode_host(dev, batch, wire_dtypes={(1, 0): 1}, grpc=True, pinned_wire=False)
        want = [wire_oracle.encode_predict_request("m1", 1, [("a", a), ("tiny", tiny)]),
                wire_oracle.encode_predict_request("m2", None, [("h", h.astype(np.float32)), ("b", b)]),
                wire_oracle.encode_predict_request("m3", 7, [("d", d), ("a2", a[:17])])]
        for got, ref in zip(wires, want):
            assert got == b"\x00" + len(ref).to_bytes(4, "big") + ref      # gRPC's length-prefixed message
    finally:
        dev.close()


def test_default_threshold_slices_a_c2_request_and_leaves_small_ones_alone():
    dev = Dev(0)
    try:
        x = _with_snan(np.random.default_rng(1).standard_normal((1024, 1024), dtype=np.float32))
        small = np.random.default_rng(2).standard_normal((64, 64), dtype=np.float32)
        assert _encode_host(dev, [("default", 1, [("x", small)])])[0] == wire_oracle.encode_predict_request("default", 1, [("x", small)])
        assert _pipelined(dev) == 0
        assert _encode_host(dev, [("default", 1, [("x", x)])])[0] == wire_oracle.encode_predict_request("default", 1, [("x", x)])
        assert _pipelined(dev) == 1
    finally:
        dev.close()


def _decode_host(dev, wire, dst_stride, shift=0, pinned_dst=True):
    wire_buf = N.PinnedBuffer(len(wire) + 512)
    wire_buf.array[shift: shift + len(wire)] = np.frombuffer(wire, dtype=np.uint8)
    if pinned_dst:       # the kernel writes the tensors into it itself
        out_buf = N.PinnedBuffer(dst_stride)
        out, optr = out_buf.array, out_buf.ptr
    else:
        out_buf = out = np.empty(dst_stride, dtype=np.uint8)
        optr = out.ctypes.data
    out[:] = 0xEE
    off, ln = (C.c_uint64 * 1)(0), (C.c_uint64 * 1)(len(wire))
    N.check(dev.lib.b200tfs_decode_responses_host_async(dev.ctx, wire_buf.ptr + shift, 1, off, ln, optr, dst_stride))
    outs = (N.Output * N.FUSED_MAX_OUTPUTS)()
    n_outs, specs, status = (C.c_int32 * 1)(), (N.ModelSpec * 1)(), (C.c_int32 * 1)()
    N.check(dev.lib.b200tfs_decode_results(dev.ctx, 1, outs, n_outs, specs, status))
    return out, outs, n_outs[0], status[0], (wire_buf, out_buf)


@pytest.mark.parametrize("elems,key", [(20000, "y"), (65536 + 3, "scores"), (100001, "a_rather_long_output_name"), (1 << 20, "y")])
def test_sliced_decode_matches_the_oracle(elems, key):
    dev = _dev_with_threshold(4096)
    try:
        x = _with_snan(np.random.default_rng(elems).standard_normal(elems, dtype=np.float32))
        wire = wire_oracle.build_predict_response([(key, x)], keep_snan=True)
        ref = wire_oracle.decode_predict_response(wire)[key]
        stride = (len(wire) + 2303 + 255) & ~255
        for rep, shift in enumerate((0, 0, 5)):          # twice the same buffer geometry, then a misaligned host pointer
            out, outs, n_outs, status, keep = _decode_host(dev, wire, stride, shift, pinned_dst=(rep != 1))
            assert status == 0 and n_outs == 1 and outs[0].status == 0 and outs[0].n_elems == elems
            got = out[outs[0].dst_off: outs[0].dst_off + outs[0].dst_bytes]
            assert got.tobytes() == ref.tobytes(), (rep, shift)
        assert _pipelined(dev) == 3
    finally:
        dev.close()


def test_sliced_decode_of_float64_and_int8_content():
    dev = _dev_with_threshold(4096)
    try:
        d = np.random.default_rng(3).standard_normal(50001)
        wire = wire_oracle.build_predict_response([("d", d)])
        out, outs, n_outs, status, keep = _decode_host(dev, wire, (len(wire) + 2560) & ~255)
        assert status == 0 and out[outs[0].dst_off: outs[0].dst_off + outs[0].dst_bytes].tobytes() == d.tobytes()
        assert _pipelined(dev) == 1
        # two outputs: values in two chunks -> not sliced, still right
        a = np.random.default_rng(4).standard_normal(30000, dtype=np.float32)
        wire = wire_oracle.build_predict_response([("a", a), ("d", d)])
        out, outs, n_outs, status, keep = _decode_host(dev, wire, (len(wire) + 4096) & ~255)
        assert status == 0 and n_outs == 2 and _pipelined(dev) == 1
        ref = wire_oracle.decode_predict_response(wire)
        for k in range(2):
            name = bytes(np.frombuffer(wire, dtype=np.uint8)[outs[k].key_off: outs[k].key_off + outs[k].key_len]).decode()
            assert out[outs[k].dst_off: outs[k].dst_off + outs[k].dst_bytes].tobytes() == ref[name].tobytes()
    finally:
        dev.close()


def test_python_api_round_trip_takes_the_pipeline():
    """The drop-in API (Codec.encode_predict_request / decode_predict_response) on one C2-sized tensor."""
    from min_tfs_client.codec import Codec

    codec = Codec(0)
    x = _with_snan(np.random.default_rng(9).standard_normal((1024, 1024), dtype=np.float32))
    wire = codec.encode_predict_request("default", {"x": x}, 1)
    assert bytes(wire) == wire_oracle.encode_predict_request("default", 1, [("x", x)])
    resp = wire_oracle.build_predict_response([("y", x)])
    got, spec = codec.decode_predict_response(resp)
    assert got["y"].tobytes() == wire_oracle.decode_predict_response(resp)["y"].tobytes()


def test_set_pipeline_switches_slicing_per_context():
    dev = Dev(0)
    try:
        x = np.random.default_rng(11).standard_normal(1 << 19, dtype=np.float32)     # 2 MiB
        ref = wire_oracle.encode_predict_request("default", 1, [("x", x)])
        N.check(dev.lib.b200tfs_set_pipeline(dev.ctx, 0, 0))
        assert _encode_host(dev, [("default", 1, [("x", x)])])[0] == ref and _pipelined(dev) == 0
        N.check(dev.lib.b200tfs_set_pipeline(dev.ctx, 1 << 16, 8))
        assert _encode_host(dev, [("default", 1, [("x", x)])])[0] == ref and _pipelined(dev) == 1
        assert dev.lib.b200tfs_set_pipeline(dev.ctx, 1 << 16, 1) == N.E_ARG
    finally:
        dev.close()
