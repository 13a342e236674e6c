"""Device-resident and page-locked arrays for the drop-in API (ctypes only: no PyTorch, no CuPy needed).

``north_star``: the kernels "read the device-resident tensor".  The reference only knows host ``ndarray``s
(tensors.py:28-35); here ``ndarray_to_tensor_proto`` / ``predict_request`` also accept anything that exposes
``__cuda_array_interface__`` (CuPy, Numba, PyTorch CUDA tensors, ``DeviceArray`` below) or ``__dlpack__`` on a CUDA
device: such an input is encoded straight from HBM - no host-to-device copy at all.

``pinned_empty`` hands out page-locked numpy arrays: host inputs that live in them (and ``out=`` destinations for
decode) are copied by the DMA engines at the full PCIe rate instead of through the driver's pageable staging.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _native as N


class DeviceArray:
    """A C-contiguous array in device memory owned through the C ABI (``b200tfs_malloc``), with the CUDA array interface."""

    def __init__(self, codec, shape, dtype):
        self._codec = codec
        self.shape = tuple(int(d) for d in shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        p = C.c_void_p()
        N.check(codec._lib.b200tfs_malloc(codec._ctx, max(self.nbytes, 1), C.byref(p)))
        self.ptr = p.value

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self.ptr, False), "version": 3, "strides": None}

    def copy_from_host(self, arr):
        arr = np.require(arr, dtype=self.dtype, requirements="C")
        if arr.shape != self.shape:
            raise ValueError(f"shape {arr.shape} != {self.shape}")
        if self.nbytes:
            N.check(self._codec._lib.b200tfs_memcpy_h2d(self._codec._ctx, self.ptr, arr.ctypes.data, self.nbytes))
            self._codec.sync()
        return self

    def copy_to_host(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=self.dtype)
        if self.nbytes:
            N.check(self._codec._lib.b200tfs_memcpy_d2h(self._codec._ctx, out.ctypes.data, self.ptr, self.nbytes))
            self._codec.sync()
        return out

    def free(self):
        if self.ptr and getattr(self._codec, "_ctx", None):
            self._codec._lib.b200tfs_free(self._codec._ctx, self.ptr)
        self.ptr = None

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.free()
        except Exception:
            pass


# ---- DLPack (dlpack.h, stable ABI of DLManagedTensor) ------------------------------------------------
class _DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int32), ("device_id", C.c_int32)]


class _DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class _DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", _DLDevice), ("ndim", C.c_int32), ("dtype", _DLDataType),
                ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


class _DLManagedTensor(C.Structure):
    pass


_DLManagedTensor._fields_ = [("dl_tensor", _DLTensor), ("manager_ctx", C.c_void_p),
                             ("deleter", C.CFUNCTYPE(None, C.POINTER(_DLManagedTensor)))]

_KDL_CUDA, _KDL_CUDA_HOST, _KDL_CUDA_MANAGED = 2, 3, 13
_DL_CODES = {0: "i", 1: "u", 2: "f", 4: "bfloat", 5: "c", 6: "b"}


class _DLPackHold:
    """Keeps an imported DLPack tensor alive; calls its deleter when dropped (consumer side of the protocol)."""

    def __init__(self, capsule):
        api = C.pythonapi
        api.PyCapsule_IsValid.restype, api.PyCapsule_IsValid.argtypes = C.c_int, [C.py_object, C.c_char_p]
        api.PyCapsule_GetPointer.restype, api.PyCapsule_GetPointer.argtypes = C.c_void_p, [C.py_object, C.c_char_p]
        api.PyCapsule_SetName.restype, api.PyCapsule_SetName.argtypes = C.c_int, [C.py_object, C.c_char_p]
        if not api.PyCapsule_IsValid(capsule, b"dltensor"):
            raise TypeError("not a DLPack capsule (or already consumed)")
        self._capsule = capsule
        self.managed = C.cast(api.PyCapsule_GetPointer(capsule, b"dltensor"), C.POINTER(_DLManagedTensor))
        api.PyCapsule_SetName(capsule, b"used_dltensor")      # ownership is ours now

    def __del__(self):  # pragma: no cover
        try:
            m = self.managed
            if m and m.contents.deleter:
                m.contents.deleter(m)
            self.managed = None
        except Exception:
            pass


def _c_contiguous(shape, strides_in_elems) -> bool:
    expect = 1
    for d, s in zip(reversed(shape), reversed(strides_in_elems)):
        if d != 1 and s != expect:
            return False
        expect *= d
    return True


def is_device_object(obj) -> bool:
    """True for objects this module can read in place from device memory."""
    if isinstance(obj, np.ndarray) or isinstance(obj, (bytes, str, int, float, list, tuple)):
        return False
    if hasattr(obj, "__cuda_array_interface__"):
        return True
    if hasattr(obj, "__dlpack_device__"):
        try:
            return obj.__dlpack_device__()[0] in (_KDL_CUDA, _KDL_CUDA_MANAGED)
        except Exception:  # noqa: BLE001
            return False
    return False


def device_view(obj) -> Tuple[int, Tuple[int, ...], np.dtype, object]:
    """(device pointer, shape, numpy dtype, keep-alive) of a C-contiguous device array; ValueError otherwise."""
    if hasattr(obj, "__cuda_array_interface__"):
        cai = obj.__cuda_array_interface__
        shape = tuple(int(d) for d in cai["shape"])
        dtype = np.dtype(cai["typestr"])
        strides = cai.get("strides")
        if strides is not None and not _c_contiguous(shape, [s // dtype.itemsize for s in strides]):
            raise ValueError("device input must be C-contiguous (the reference ravel()s in C order, tensors.py:34)")
        if cai.get("mask") is not None:
            raise ValueError("masked device arrays are not supported")
        ptr = cai["data"][0] or 0
        return int(ptr), shape, dtype, obj
    hold = _DLPackHold(obj.__dlpack__())
    t = hold.managed.contents.dl_tensor
    if t.device.device_type not in (_KDL_CUDA, _KDL_CUDA_MANAGED):
        raise ValueError("DLPack tensor is not on a CUDA device")
    if t.dtype.lanes != 1:
        raise ValueError("vector dtypes are not supported")
    shape = tuple(int(t.shape[i]) for i in range(t.ndim))
    kind = _DL_CODES.get(t.dtype.code)
    if kind == "bfloat":
        from .constants import BFLOAT16

        if BFLOAT16 is None or t.dtype.bits != 16:
            raise ValueError("bfloat16 needs ml_dtypes")
        dtype = np.dtype(BFLOAT16)
    elif kind == "b":
        dtype = np.dtype(np.bool_)
    elif kind is None:
        raise ValueError(f"DLPack dtype code {t.dtype.code} is not supported")
    else:
        dtype = np.dtype(f"{kind}{t.dtype.bits // 8}")
    if t.strides and not _c_contiguous(shape, [int(t.strides[i]) for i in range(t.ndim)]):
        raise ValueError("device input must be C-contiguous (the reference ravel()s in C order, tensors.py:34)")
    return int(t.data or 0) + int(t.byte_offset), shape, dtype, hold


class PinnedArrays:
    """Registry of page-locked arrays handed out by ``pinned_empty``: address -> (buffer, capacity)."""

    def __init__(self):
        self._by_addr = {}

    def empty(self, shape, dtype) -> np.ndarray:
        dtype = np.dtype(dtype)
        shape = tuple(int(d) for d in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        cap = ((nbytes + 255) & ~255) + 256                 # room for the 256-byte granularity of the fused decode's slots
        buf = N.PinnedBuffer(cap)
        arr = buf.array[:nbytes].view(dtype).reshape(shape)
        self._by_addr[buf.ptr] = (buf, cap)
        return arr

    def capacity(self, arr: np.ndarray) -> Optional[int]:
        """Bytes available from the start of `arr` if it is (the start of) one of our page-locked buffers."""
        if not isinstance(arr, np.ndarray) or not arr.flags.c_contiguous:
            return None
        hit = self._by_addr.get(arr.ctypes.data)
        return hit[1] if hit else None

    def release(self):
        for buf, _ in self._by_addr.values():
            buf.free()
        self._by_addr.clear()
