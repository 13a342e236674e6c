"""Thin test helpers over the C ABI for device-resident buffers."""
import ctypes as C

import numpy as np

from min_tfs_client import _native as N


class Dev:
    def __init__(self, device=0):
        self.lib = N.load()
        self.ctx = C.c_void_p()
        N.check(self.lib.b200tfs_create(device, C.byref(self.ctx)))
        self.allocs = []

    def close(self):
        for p in self.allocs:
            self.lib.b200tfs_free(self.ctx, p)
        self.lib.b200tfs_destroy(self.ctx)

    def malloc(self, nbytes):
        p = C.c_void_p()
        N.check(self.lib.b200tfs_malloc(self.ctx, max(int(nbytes), 1), C.byref(p)))
        self.allocs.append(p.value)
        return p.value

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        p = self.malloc(arr.nbytes + 256)
        if arr.nbytes:
            N.check(self.lib.b200tfs_memcpy_h2d(self.ctx, p, arr.ctypes.data, arr.nbytes))
        self.sync()
        return p

    def download(self, ptr, nbytes, dtype=np.uint8):
        out = np.empty(int(nbytes), dtype=np.uint8)
        if nbytes:
            N.check(self.lib.b200tfs_memcpy_d2h(self.ctx, out.ctypes.data, ptr, int(nbytes)))
        self.sync()
        return out.view(dtype)

    def sync(self):
        N.check(self.lib.b200tfs_sync(self.ctx))


def tensor_struct(ptr, arr, key=b"", wire_dtype=None, flags=0):
    from min_tfs_client.constants import enum_for_numpy

    dims = (C.c_int64 * max(arr.ndim, 1))(*arr.shape)
    e = enum_for_numpy(arr.dtype)
    t = N.Tensor(data=ptr, src_dtype=e, wire_dtype=e if wire_dtype is None else wire_dtype, rank=arr.ndim, flags=flags, dims=dims,
                 key=key, key_len=len(key), packed_len=0)
    return t, dims
