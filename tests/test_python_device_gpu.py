"""The drop-in Python API on device-resident and page-locked arrays: the north_star's input is a tensor that already lives in
HBM - ``__cuda_array_interface__`` / DLPack objects are encoded in place (no host-to-device copy) - and pinned arrays take the
copies that remain at the PCIe rate.  Every result is compared with the oracle, bit for bit."""
import numpy as np
import pytest

from min_tfs_client import device as D
from min_tfs_client.tensors import ndarray_to_tensor_proto_bytes
from oracle import wire_oracle

pytestmark = pytest.mark.gpu


def test_device_array_inputs_encode_like_host_arrays(codec):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((33, 65)).astype(np.float32)
    x.reshape(-1)[:2] = np.array([0x7F800001, 0xFFC00000], dtype=np.uint32).view(np.float32)
    ids = rng.integers(-1000, 1 << 40, size=(5, 7), dtype=np.int64)
    h = rng.standard_normal((8, 16)).astype(np.float16)
    xd, idd, hd = codec.device_array(x), codec.device_array(ids), codec.device_array(h)
    assert D.is_device_object(xd) and not D.is_device_object(x)
    assert xd.copy_to_host().tobytes() == x.tobytes()
    want = wire_oracle.encode_predict_request("m", 3, [("x", x), ("ids", ids)])
    assert codec.encode_predict_request("m", {"x": xd, "ids": idd}, 3) == want
    assert codec.encode_predict_request("m", {"x": xd, "ids": ids}, 3) == want                  # device and host inputs in one request
    assert codec.encode_tensor_protos([xd])[0] == wire_oracle.encode_tensor_proto(x)
    assert ndarray_to_tensor_proto_bytes(idd) == wire_oracle.encode_tensor_proto(ids)
    assert codec.encode_predict_request("m", {"h": hd}, None, wire_dtype="DT_FLOAT") == \
        wire_oracle.encode_predict_request("m", None, [("h", h.astype(np.float32))])
    scalar = codec.device_array(np.float32(2.5))
    assert codec.encode_tensor_protos([scalar])[0] == wire_oracle.encode_tensor_proto(np.float32(2.5).reshape(()))
    for a in (xd, idd, hd, scalar):
        a.free()


def test_torch_cuda_tensors_through_cuda_array_interface_and_dlpack(codec):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("torch sees no CUDA device")
    rng = np.random.default_rng(4)
    x = rng.standard_normal((17, 9)).astype(np.float32)
    t = torch.from_numpy(x).cuda()
    assert hasattr(t, "__cuda_array_interface__")
    want = wire_oracle.encode_predict_request("default", 1, [("x", x)])
    assert codec.encode_predict_request("default", {"x": t}, 1) == want

    class OnlyDLPack:                      # an object that offers nothing but the DLPack protocol
        def __init__(self, t):
            self.t = t

        def __dlpack__(self, stream=None):
            return self.t.__dlpack__()

        def __dlpack_device__(self):
            return self.t.__dlpack_device__()
    assert codec.encode_predict_request("default", {"x": OnlyDLPack(t)}, 1) == want
    ids = torch.arange(-5, 300, dtype=torch.int32, device="cuda")
    assert codec.encode_tensor_protos([OnlyDLPack(ids)])[0] == wire_oracle.encode_tensor_proto(np.arange(-5, 300, dtype=np.int32))
    bf = torch.from_numpy(x).cuda().to(torch.bfloat16)
    import ml_dtypes

    host_bf = x.astype(ml_dtypes.bfloat16)
    assert codec.encode_predict_request("default", {"x": OnlyDLPack(bf)}, 1, wire_dtype="DT_FLOAT") == \
        wire_oracle.encode_predict_request("default", 1, [("x", host_bf.astype(np.float32))])
    with pytest.raises(ValueError):        # the reference ravel()s in C order: a transposed device view is refused, not silently re-ordered
        codec.encode_predict_request("default", {"x": t.t()}, 1)


def test_pinned_arrays_and_out(codec):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((256, 300)).astype(np.float32)
    xp = codec.pinned_empty(x.shape, x.dtype)
    xp[...] = x
    want = wire_oracle.encode_predict_request("default", 1, [("x", x)])
    assert codec.encode_predict_request("default", {"x": xp}, 1) == want
    view = codec.encode_predict_request("default", {"x": xp}, 1, out="pinned")
    assert isinstance(view, np.ndarray) and view.tobytes() == want
    resp = wire_oracle.build_predict_response([("y", x)])
    yp = codec.pinned_empty(x.shape, x.dtype)
    yp[...] = 0
    outs, spec = codec.decode_predict_response(resp, out={"y": yp})
    assert outs["y"] is yp and yp.tobytes() == x.tobytes() and spec.name == "default"
    rp = codec.pinned_empty((len(resp),), np.uint8)            # the wire itself in page-locked memory
    rp[...] = np.frombuffer(resp, np.uint8)
    yp[...] = 0
    assert codec.decode_predict_response(rp, out={"y": yp})[0]["y"].tobytes() == x.tobytes()
    # general path: an ordinary array as destination, two outputs, a varint output
    ids = np.arange(-7, 50, dtype=np.int64)
    resp2 = wire_oracle.build_predict_response([("y", x), ("ids", ids)])
    y2, i2 = np.zeros_like(x), np.zeros_like(ids)
    outs, _ = codec.decode_predict_response(resp2, out={"y": y2, "ids": i2})
    assert outs["y"] is y2 and y2.tobytes() == x.tobytes() and i2.tolist() == ids.tolist()
    with pytest.raises(ValueError):
        codec.decode_predict_response(resp2, out={"y": np.zeros((3,), np.float32)})
    with pytest.raises(KeyError):
        codec.decode_predict_response(resp2, out={"nope": y2})
    # a pinned destination whose response turns out to have two outputs: the general path fills it all the same
    yp[...] = 0
    assert codec.decode_predict_response(resp2, out={"y": yp})[0]["y"].tobytes() == x.tobytes()
