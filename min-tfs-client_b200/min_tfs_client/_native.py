"""ctypes binding of ``libb200tfs.so`` (C ABI declared in ``include/b200tfs.h``).

No PyTorch and no CUDA Python package: the shared library is the only native dependency, loaded from
``min-tfs-client_b200/lib/``.  If the library is missing or no CUDA device is present the codec fails
loudly (``RuntimeError``) - there is no CPU fallback behind this module.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200TFS_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libb200tfs.so")   # B200TFS_LIB: experiment builds

# ---- status codes (b200tfs.h) -----------------------------------------------------------------
OK = 0
E_DTYPE, E_SHAPE, E_SIZE, E_PARSE, E_CUDA, E_TOOBIG, E_ARG, E_NONCANONICAL, E_RANGE, E_KEY, E_SPILL = range(-1, -12, -1)

F_TENSOR_CONTENT = 0x1
F_KEEP_SNAN = 0x2
F_PRESERIALIZED = 0x4
F_DEVICE_DATA = 0x8
RF_GRPC_FRAME = 0x1
OF_TENSOR_CONTENT, OF_MULTI_CHUNK, OF_DIM_INFERRED, OF_HAS_UNKNOWN, OF_RANK0, OF_VARINT, OF_PAD_EDGE = 0x1, 0x2, 0x4, 0x8, 0x10, 0x20, 0x40
OF_UNPACKED, OF_SPILLED = 0x80, 0x100
ORDER_GIVEN, ORDER_UPB, ORDER_BYTES = 0, 1, 2
MAX_RANK, MAX_RUNS, FUSED_MAX_OUTPUTS = 16, 8, 8
DT_HALF_REFQUIRK = -19


class NativeError(RuntimeError):
    """A non-OK status from libb200tfs that has no closer Python exception type."""

    def __init__(self, code, message):
        super().__init__(f"libb200tfs error {code}: {message}")
        self.code = code


class Tensor(C.Structure):
    _fields_ = [
        ("data", C.c_void_p), ("src_dtype", C.c_int32), ("wire_dtype", C.c_int32), ("rank", C.c_int32),
        ("flags", C.c_uint32), ("dims", C.POINTER(C.c_int64)), ("key", C.c_char_p), ("key_len", C.c_int64),
        ("packed_len", C.c_uint64),
    ]


class Request(C.Structure):
    _fields_ = [
        ("model_name", C.c_char_p), ("model_name_len", C.c_int64), ("has_version", C.c_int32), ("order", C.c_int32),
        ("version", C.c_int64), ("n_inputs", C.c_int32), ("flags", C.c_int32), ("inputs", C.POINTER(Tensor)),
    ]


class Run(C.Structure):
    """b200tfs_run: `count` pieces of `len` value bytes, `stride` bytes apart, from TensorProto field `field`."""
    _fields_ = [("off", C.c_uint64), ("len", C.c_uint32), ("count", C.c_uint32), ("stride", C.c_uint32), ("field", C.c_uint32)]


class Output(C.Structure):
    _fields_ = [
        ("key_off", C.c_uint64), ("key_len", C.c_uint32), ("dtype", C.c_int32), ("rank", C.c_int32), ("flags", C.c_uint32),
        ("value_field", C.c_int32), ("n_runs", C.c_int32), ("dims", C.c_int64 * MAX_RANK),
        ("runs", Run * MAX_RUNS),
        ("content_off", C.c_uint64), ("content_len", C.c_uint64), ("msg_off", C.c_uint64), ("msg_len", C.c_uint64),
        ("n_elems", C.c_uint64), ("dst_bytes", C.c_uint64), ("n_strings", C.c_uint64), ("dst_off", C.c_uint64), ("status", C.c_int32),
        ("n_inline", C.c_uint32), ("spill_rec", C.c_uint32), ("spill_seq", C.c_uint32),
    ]


class ModelSpec(C.Structure):
    _fields_ = [
        ("name_off", C.c_uint64), ("name_len", C.c_uint32), ("signature_len", C.c_uint32), ("signature_off", C.c_uint64),
        ("label_off", C.c_uint64), ("label_len", C.c_uint32), ("has_version", C.c_int32), ("version", C.c_int64),
    ]


_u64p = C.POINTER(C.c_uint64)
_i32p = C.POINTER(C.c_int32)
_vp = C.c_void_p
_vpp = C.POINTER(C.c_void_p)

# every symbol include/b200tfs.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "b200tfs_abi_version": (C.c_int, []),
    "b200tfs_last_error": (C.c_char_p, []),
    "b200tfs_device_count": (C.c_int, [_i32p]),
    "b200tfs_create": (C.c_int, [C.c_int, _vpp]),
    "b200tfs_destroy": (C.c_int, [_vp]),
    "b200tfs_sync": (C.c_int, [_vp]),
    "b200tfs_stream": (C.c_void_p, [_vp]),
    "b200tfs_set_stream": (C.c_int, [_vp, _vp]),
    "b200tfs_kernel_launches": (C.c_int, [_vp, _u64p]),
    "b200tfs_malloc": (C.c_int, [_vp, C.c_uint64, _vpp]),
    "b200tfs_free": (C.c_int, [_vp, _vp]),
    "b200tfs_host_alloc": (C.c_int, [C.c_uint64, _vpp]),
    "b200tfs_host_free": (C.c_int, [_vp]),
    "b200tfs_memcpy_h2d": (C.c_int, [_vp, _vp, _vp, C.c_uint64]),
    "b200tfs_memcpy_d2h": (C.c_int, [_vp, _vp, _vp, C.c_uint64]),
    "b200tfs_memcpy_d2d": (C.c_int, [_vp, _vp, _vp, C.c_uint64]),
    "b200tfs_memset": (C.c_int, [_vp, _vp, C.c_int, C.c_uint64]),
    "b200tfs_event_create": (C.c_int, [_vpp]),
    "b200tfs_event_destroy": (C.c_int, [_vp]),
    "b200tfs_event_record": (C.c_int, [_vp, _vp]),
    "b200tfs_event_sync": (C.c_int, [_vp]),
    "b200tfs_event_elapsed_ms": (C.c_int, [_vp, _vp, C.POINTER(C.c_float)]),
    "b200tfs_dtype_size": (C.c_int, [C.c_int32]),
    "b200tfs_dtype_field": (C.c_int, [C.c_int32]),
    "b200tfs_cast_supported": (C.c_int, [C.c_int32, C.c_int32]),
    "b200tfs_tensor_proto_size": (C.c_int, [C.POINTER(Tensor), _u64p, _u64p]),
    "b200tfs_request_size": (C.c_int, [C.POINTER(Request), _u64p]),
    "b200tfs_tensor_proto_header": (C.c_int, [C.POINTER(Tensor), _vp, C.c_uint64, _u64p]),
    "b200tfs_request_frame": (C.c_int, [C.POINTER(Request), _vp, C.c_uint64, _u64p, _u64p, _u64p, _i32p]),
    "b200tfs_order_keys": (C.c_int, [C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.c_int32, _i32p]),
    "b200tfs_tensor_arena_size": (C.c_int, [C.c_int32, C.POINTER(Tensor), _u64p]),
    "b200tfs_request_arena_size": (C.c_int, [C.c_int32, C.POINTER(Request), _u64p]),
    "b200tfs_measure": (C.c_int, [_vp, C.c_int32, C.POINTER(Tensor)]),
    "b200tfs_encode_tensor_protos": (C.c_int, [_vp, C.c_int32, C.POINTER(Tensor), _vp, C.c_uint64, _u64p, _u64p]),
    "b200tfs_encode_requests": (C.c_int, [_vp, C.c_int32, C.POINTER(Request), _vp, C.c_uint64, _u64p, _u64p]),
    "b200tfs_encode_requests_async": (C.c_int, [_vp, C.c_int32, C.POINTER(Request), _vp, C.c_uint64]),
    "b200tfs_encode_results": (C.c_int, [_vp, C.c_int32, _u64p, _u64p]),
    "b200tfs_request_frame_deferred": (C.c_int, [C.POINTER(Request), _u64p, _vp, C.c_uint64, _u64p, _u64p, _u64p, _u64p]),
    "b200tfs_parse_responses": (C.c_int, [_vp, _vp, C.c_int32, _u64p, _u64p, C.c_int32, C.POINTER(Output), _i32p,
                                          C.POINTER(ModelSpec), _i32p]),
    "b200tfs_parse_tensor_protos": (C.c_int, [_vp, _vp, C.c_int32, _u64p, _u64p, C.POINTER(Output), _i32p]),
    "b200tfs_output_dims": (C.c_int, [_vp, C.POINTER(Output), C.POINTER(C.c_int64), C.c_int32]),
    "b200tfs_output_runs": (C.c_int, [_vp, C.POINTER(Output), C.POINTER(Run), C.c_int32]),
    "b200tfs_unpack_outputs": (C.c_int, [_vp, _vp, C.c_int32, C.POINTER(Output), _u64p, _vpp, _i32p, _i32p]),
    "b200tfs_decode_responses": (C.c_int, [_vp, _vp, C.c_int32, _u64p, _u64p, _vp, C.c_uint64]),
    "b200tfs_decode_results": (C.c_int, [_vp, C.c_int32, C.POINTER(Output), _i32p, C.POINTER(ModelSpec), _i32p]),
    "b200tfs_decode_stats": (C.c_int, [_vp, _u64p, _u64p, _u64p]),
    "b200tfs_capture_begin": (C.c_int, [_vp]),
    "b200tfs_capture_end": (C.c_int, [_vp, _vpp]),
    "b200tfs_graph_launch": (C.c_int, [_vp, _vp]),
    "b200tfs_graph_destroy": (C.c_int, [_vp]),
    "b200tfs_wait_event": (C.c_int, [_vp, _vp]),
    "b200tfs_encode_requests_host": (C.c_int, [_vp, C.c_int32, C.POINTER(Request), _vp, C.c_uint64, _u64p, _u64p]),
    "b200tfs_encode_requests_host_async": (C.c_int, [_vp, C.c_int32, C.POINTER(Request), _vp, C.c_uint64, _u64p, _u64p]),
    "b200tfs_pipelined_calls": (C.c_int, [_vp, _u64p]),
    "b200tfs_direct_calls": (C.c_int, [_vp, _u64p]),
    "b200tfs_set_pipeline": (C.c_int, [_vp, C.c_uint64, C.c_int32]),
    "b200tfs_set_decode_cast": (C.c_int, [_vp, C.c_int32]),
    "b200tfs_decode_responses_host_async": (C.c_int, [_vp, _vp, C.c_int32, _u64p, _u64p, _vp, C.c_uint64]),
    "b200tfs_encode_tensor_protos_host": (C.c_int, [_vp, C.c_int32, C.POINTER(Tensor), _vp, C.c_uint64, _u64p, _u64p]),
    "b200tfs_parse_responses_host": (C.c_int, [_vp, _vp, C.c_int32, _u64p, _u64p, C.c_int32, C.POINTER(Output), _i32p,
                                               C.POINTER(ModelSpec), _i32p]),
    "b200tfs_parse_tensor_protos_host": (C.c_int, [_vp, _vp, C.c_int32, _u64p, _u64p, C.POINTER(Output), _i32p]),
    "b200tfs_unpack_outputs_host": (C.c_int, [_vp, C.c_int32, C.POINTER(Output), _u64p, _vpp, _i32p, _i32p]),
}

_lib = None
_lib_lock = threading.Lock()


def load():
    """Load libb200tfs.so once and bind every declared symbol; RuntimeError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is not built (run `python -c 'import __graft_entry__ as g; g.build()'` at the repo root); "
                "min_tfs_client has no CPU codec to fall back to"
            )
        lib = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError = header/library out of sync: fail loudly
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def last_error() -> str:
    return (load().b200tfs_last_error() or b"").decode("utf-8", "replace")


def check(code: int) -> None:
    """Map a status code to the exception the reference raises in the same situation."""
    if code == OK:
        return
    msg = last_error()
    if code in (E_DTYPE, E_SHAPE):
        raise ValueError(msg)
    if code == E_KEY:
        raise KeyError(msg)
    if code == E_RANGE:
        raise OverflowError(msg)
    if code == E_PARSE:
        from google.protobuf.message import DecodeError

        raise DecodeError(msg)
    if code == E_TOOBIG:
        raise ValueError(msg)
    if code == E_NONCANONICAL:
        raise NotImplementedError(msg)
    raise NativeError(code, msg)


def device_count() -> int:
    n = C.c_int32(0)
    rc = load().b200tfs_device_count(C.byref(n))
    return int(n.value) if rc == OK else 0


class PinnedBuffer:
    """Page-locked host memory exposed as a numpy uint8 array (``.array``)."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        check(load().b200tfs_host_alloc(max(self.nbytes, 1), C.byref(p)))
        self.ptr = p.value
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr))[: self.nbytes]

    def free(self):
        if self.ptr:
            self.array = None
            load().b200tfs_host_free(self.ptr)
            self.ptr = None

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.free()
        except Exception:
            pass
