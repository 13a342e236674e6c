"""GPU wire codec behind the drop-in API: numpy arrays <-> TensorProto / PredictRequest /
PredictResponse wire bytes, through ``libb200tfs.so`` (ctypes, no PyTorch).

What this replaces in the reference (paths relative to its checkout):
  encode  ``ndarray_to_tensor_proto`` + ``write_values_to_tensor_proto`` (tensors.py:17-35), the request
          assembly in ``_make_inference_request`` (requests.py:41-48) and ``SerializeToString``
          (prediction_service_pb2_grpc.py:52);
  decode  ``PredictResponse.FromString`` (…pb2_grpc.py:53), ``extract_shape`` and
          ``tensor_proto_to_ndarray`` (tensors.py:38-46).

A ``Codec`` owns one native context (one CUDA stream + scratch) and is not thread-safe; use
``get_codec()`` for a per-thread instance.  ``DT_STRING`` tensors are variable-length host objects:
their TensorProto is assembled on the host and spliced into the request by the kernel verbatim.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Dict, Iterable, List, Mapping, Optional, Sequence, Tuple, Union

import numpy as np

from . import _native as N
from . import device as D
from .constants import BFLOAT16, DT_BFLOAT16, NP_TO_ENUM_MAPPING, enum_for_numpy, numpy_for_enum
from tensorflow.core.framework import types_pb2

DT_FLOAT, DT_HALF, DT_STRING = types_pb2.DT_FLOAT, types_pb2.DT_HALF, types_pb2.DT_STRING
DT_COMPLEX64, DT_COMPLEX128 = types_pb2.DT_COMPLEX64, types_pb2.DT_COMPLEX128

_ORDER = {"given": N.ORDER_GIVEN, "insertion": N.ORDER_GIVEN, "deterministic": N.ORDER_UPB, "upb": N.ORDER_UPB,
          "bytes": N.ORDER_BYTES}


def _as_enum(dtype) -> int:
    """DT_* enum from an enum int, a "DT_*" name or a numpy dtype."""
    if isinstance(dtype, (int, np.integer)):
        return int(dtype)
    if isinstance(dtype, str) and dtype.startswith("DT_"):
        return getattr(types_pb2, dtype)
    return enum_for_numpy(dtype)


def _validated_enum(arr) -> int:
    """Same gate as ``DataType(ndarray.dtype.type)`` (types.py:27-32): ValueError for foreign dtypes.  `arr`: anything with a numpy dtype."""
    t = arr.dtype.type
    if t in NP_TO_ENUM_MAPPING:
        return NP_TO_ENUM_MAPPING[t]
    if BFLOAT16 is not None and t is BFLOAT16:
        return DT_BFLOAT16
    allowed = ", ".join(k.__name__ for k in NP_TO_ENUM_MAPPING)
    raise ValueError(f"Dtype {t.__name__} is not valid. Allowable values: {allowed}")


def _string_tensor_proto_bytes(arr: np.ndarray) -> bytes:
    """Host assembly of a DT_STRING TensorProto (tensors.py:24, :10-14: str -> utf-8, bytes as is)."""
    from tensorflow.core.framework.tensor_pb2 import TensorProto
    from tensorflow.core.framework.tensor_shape_pb2 import TensorShapeProto

    proto = TensorProto(dtype=DT_STRING,
                        tensor_shape=TensorShapeProto(dim=[TensorShapeProto.Dim(size=d) for d in arr.shape]))
    proto.string_val.extend(v.encode("utf-8") if isinstance(v, str) else v for v in arr.ravel().tolist())
    return proto.SerializeToString()


class _Prepared:
    """One input readied for the C ABI; keeps every buffer the Tensor struct points at alive."""

    __slots__ = ("array", "dims", "key", "struct", "on_device", "nbytes", "ndim", "size", "itemsize", "kind")

    def __init__(self, value, key: bytes, wire_dtype, tensor_content: bool, keep_snan: bool):
        self.on_device = D.is_device_object(value)
        if self.on_device:
            # a tensor that already lives in HBM (CUDA array interface / DLPack): encoded in place, no host-to-device copy
            ptr, shape, dtype, keep = D.device_view(value)
            probe = np.empty(0, dtype=dtype)
            src_enum = _validated_enum(probe)
            if src_enum == DT_STRING or not dtype.isnative:
                raise ValueError("device inputs must be native-endian numeric arrays")
            wire_enum = src_enum if wire_dtype is None else _as_enum(wire_dtype)
            flags = N.F_DEVICE_DATA | (N.F_TENSOR_CONTENT if tensor_content else 0) | (N.F_KEEP_SNAN if keep_snan else 0)
            self.array = keep
            self.dims = (C.c_int64 * max(len(shape), 1))(*shape)
            n = int(np.prod(shape, dtype=np.int64))
            self.nbytes, self.ndim, self.size, self.itemsize, self.kind = n * dtype.itemsize, len(shape), n, dtype.itemsize, dtype.kind
            self.key = key
            self.struct = N.Tensor(data=ptr if n else None, src_dtype=src_enum, wire_dtype=wire_enum, rank=len(shape), flags=flags, dims=self.dims,
                                   key=key, key_len=len(key), packed_len=0)
            return
        arr = np.asarray(value)
        src_enum = _validated_enum(arr)
        flags = 0
        if src_enum == DT_STRING:
            blob = np.frombuffer(_string_tensor_proto_bytes(arr), dtype=np.uint8)
            self.array, self.dims = blob, (C.c_int64 * 1)(0)
            t = N.Tensor(data=blob.ctypes.data if blob.size else None, src_dtype=DT_STRING, wire_dtype=DT_STRING, rank=0,
                         flags=N.F_PRESERIALIZED, dims=self.dims, key=key, key_len=len(key), packed_len=blob.size)
        else:
            if not arr.dtype.isnative:
                arr = arr.astype(arr.dtype.newbyteorder("="))
            # C order, like ndarray.ravel() in tensors.py:34.  (np.ascontiguousarray promotes a 0-d array to shape (1,):
            # the reference keeps the empty shape - `tensor_shape {}` = 12 00 - so ask for the layout only.)
            arr = np.require(arr, requirements="C")
            wire_enum = src_enum if wire_dtype is None else _as_enum(wire_dtype)
            if tensor_content:
                flags |= N.F_TENSOR_CONTENT
            if keep_snan:
                flags |= N.F_KEEP_SNAN
            self.array = arr
            self.dims = (C.c_int64 * max(arr.ndim, 1))(*arr.shape)
            t = N.Tensor(data=arr.ctypes.data if arr.size else None, src_dtype=src_enum, wire_dtype=wire_enum, rank=arr.ndim,
                         flags=flags, dims=self.dims, key=key, key_len=len(key), packed_len=0)
        self.key = key
        self.struct = t
        a = self.array
        self.nbytes, self.ndim, self.size, self.itemsize, self.kind = int(a.nbytes), a.ndim, int(a.size), a.dtype.itemsize, a.dtype.kind


def _wire_bound(p: "_Prepared", tensor_content: bool) -> int:
    """Upper bound of the bytes one prepared tensor occupies on the wire (payload + its own framing)."""
    frame = 64 + 16 * max(p.ndim, 1) + len(p.key)
    if p.struct.flags & N.F_PRESERIALIZED:
        return p.size + frame
    n = p.size
    if p.struct.src_dtype != p.struct.wire_dtype:
        return 4 * n + frame                      # f16 / bf16 -> DT_FLOAT
    if tensor_content or p.kind in "fc" and p.itemsize >= 4 or p.kind == "b":
        return p.nbytes + frame
    per = {1: 10, 2: 10, 4: 10, 8: 10} if p.kind == "i" else {1: 2, 2: 3, 4: 5, 8: 10}   # varint bytes per element
    return n * per[p.itemsize] + frame


class DecodedSpec:
    """model_spec of a parsed response (model.proto:9-33)."""

    __slots__ = ("name", "version", "has_version", "version_label", "signature_name")

    def __init__(self, name="", version=0, has_version=False, version_label="", signature_name=""):
        self.name, self.version, self.has_version = name, version, has_version
        self.version_label, self.signature_name = version_label, signature_name

    def __repr__(self):
        return (f"DecodedSpec(name={self.name!r}, version={self.version}, has_version={self.has_version}, "
                f"version_label={self.version_label!r}, signature_name={self.signature_name!r})")


_FUSED_MOVES = frozenset((1, 2, 8, 18))   # DT_FLOAT, DT_DOUBLE, DT_COMPLEX64, DT_COMPLEX128: what decode_fused_kernel moves itself


class OpenResponse:
    """One PredictResponse after the fused launch: the table, and the host buffer the fixed-width outputs landed in."""

    def __init__(self, codec, buf, base, dst, table):
        self._codec, self._buf, self._base, self._dst, self.table = codec, buf, base, dst, table
        self._handed = set()

    def wire_of(self, key) -> bytes:
        o = self.table[key]
        return self._buf[self._base + o.msg_off: self._base + o.msg_off + o.msg_len].tobytes()

    def array(self, key, strict: bool) -> Optional[np.ndarray]:
        """The decoded output, or None when the launch did not move it (varint-packed, string, tensor_content only): the
        caller then decodes ``wire_of(key)`` on its own.  Raises what the reference raises for this output."""
        o = self.table[key]
        if int(o.dtype) not in _FUSED_MOVES or o.status != N.OK or not o.n_runs or not o.n_elems:
            return None
        np_type, dst_code, shape = self._codec._resolve_output(o, strict, None)
        at = int(o.dst_off)
        arr = self._dst[at: at + int(o.dst_bytes)].view(np_type).reshape(shape)
        if key in self._handed:      # tensor_proto_to_ndarray returns a fresh array per call (tensors.py:46): never alias two results
            return arr.copy()
        self._handed.add(key)
        return arr


class ParsedResponse:
    """Table the parse kernel produced for one PredictResponse: where every output's values lie."""

    def __init__(self, wire, offset: int, length: int, status: int, outputs: Dict[str, N.Output], spec: DecodedSpec):
        self.wire, self.offset, self.length = wire, offset, length
        self.status, self.outputs, self.model_spec = status, outputs, spec

    def keys(self):
        return self.outputs.keys()


class Codec:
    def __init__(self, device: int = 0):
        self._lib = N.load()
        ctx = C.c_void_p()
        rc = self._lib.b200tfs_create(device, C.byref(ctx))
        if rc != N.OK:
            raise RuntimeError(f"cannot create a B200 codec context on device {device}: {N.last_error()}")
        self._ctx = ctx
        self.device = device
        self._pinned = D.PinnedArrays()
        self._wire_pinned = None      # page-locked landing buffer of encode(..., out="pinned")

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.b200tfs_destroy(self._ctx)
            self._ctx = None
            self._pinned.release()
            if self._wire_pinned is not None:
                self._wire_pinned.free()
                self._wire_pinned = None

    # ---- page-locked and device-resident arrays ---------------------------------------------------
    def pinned_empty(self, shape, dtype=np.float32) -> np.ndarray:
        """An uninitialised page-locked numpy array (lives until the codec is closed).  Inputs that sit in one are copied by the
        DMA engines at the PCIe rate; handed to ``decode_predict_response(..., out={key: arr})`` the decoded tensor lands in it
        directly."""
        return self._pinned.empty(shape, dtype)

    def device_array(self, array) -> "D.DeviceArray":
        """Upload `array` once; the returned device array can be encoded any number of times without a host-to-device copy."""
        a = np.require(np.asarray(array), requirements="C")
        return D.DeviceArray(self, a.shape, a.dtype).copy_from_host(a)

    def _pinned_wire(self, cap: int) -> "N.PinnedBuffer":
        if self._wire_pinned is None or self._wire_pinned.nbytes < cap:
            if self._wire_pinned is not None:
                self._wire_pinned.free()
            self._wire_pinned = N.PinnedBuffer(max(cap, 1 << 20))
        return self._wire_pinned

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing -----------------------------------------------------------------------------
    @property
    def ctx(self):
        return self._ctx

    def sync(self):
        N.check(self._lib.b200tfs_sync(self._ctx))

    def kernel_launches(self) -> int:
        n = C.c_uint64(0)
        N.check(self._lib.b200tfs_kernel_launches(self._ctx, C.byref(n)))
        return int(n.value)

    # ---- encode --------------------------------------------------------------------------------
    def encode_tensor_protos(self, arrays: Sequence, *, wire_dtype=None, tensor_content: bool = False,
                             keep_snan: bool = False) -> List[bytes]:
        """Wire bytes of ``ndarray_to_tensor_proto(a).SerializeToString()`` for each array."""
        preps = [_Prepared(a, b"", wire_dtype, tensor_content, keep_snan) for a in arrays]
        n = len(preps)
        if n == 0:
            return []
        out: List[Optional[bytes]] = [None] * n
        dev_idx = []
        for i, p in enumerate(preps):  # a bare DT_STRING proto is host bytes already
            if p.struct.flags & N.F_PRESERIALIZED:
                out[i] = p.array.tobytes()
            else:
                dev_idx.append(i)
        if dev_idx:
            ts = (N.Tensor * len(dev_idx))(*[preps[i].struct for i in dev_idx])
            cap = sum(_wire_bound(preps[i], tensor_content) + 512 for i in dev_idx)
            wire = np.empty(cap, dtype=np.uint8)
            off = (C.c_uint64 * len(dev_idx))()
            ln = (C.c_uint64 * len(dev_idx))()
            N.check(self._lib.b200tfs_encode_tensor_protos_host(self._ctx, len(dev_idx), ts, wire.ctypes.data, cap, off, ln))
            for j, i in enumerate(dev_idx):
                out[i] = wire[off[j]: off[j] + ln[j]].tobytes()
        return out  # type: ignore[return-value]

    def _build_requests(self, requests, order, wire_dtype, tensor_content, keep_snan, grpc_frame=False):
        order_code = _ORDER[order] if isinstance(order, str) else int(order)
        keep = []
        structs = []
        for model_name, model_version, inputs in requests:
            items = list(inputs.items()) if isinstance(inputs, Mapping) else list(inputs)
            preps = []
            for k, v in items:
                kb = k.encode("utf-8") if isinstance(k, str) else bytes(k)
                wd = wire_dtype.get(k) if isinstance(wire_dtype, Mapping) else wire_dtype
                preps.append(_Prepared(v, kb, wd, tensor_content, keep_snan))
            arr = (N.Tensor * max(len(preps), 1))(*[p.struct for p in preps])
            name = model_name.encode("utf-8") if isinstance(model_name, str) else bytes(model_name)
            req = N.Request(model_name=name, model_name_len=len(name), has_version=int(model_version is not None), order=order_code,
                            version=int(model_version) if model_version is not None else 0, n_inputs=len(preps), flags=N.RF_GRPC_FRAME if grpc_frame else 0,
                            inputs=arr)
            keep.append((preps, arr, name))
            structs.append(req)
        return keep, structs

    def encode_predict_requests(self, requests: Iterable[Tuple[str, Optional[int], Union[Mapping, Sequence]]], *,
                                order="deterministic", wire_dtype=None, tensor_content: bool = False,
                                keep_snan: bool = False, grpc_frame: bool = False, out=None) -> List[bytes]:
        """Each item is ``(model_name, model_version, inputs)``; returns one PredictRequest wire per item.

        The bytes equal ``PredictRequest.SerializeToString(deterministic=True)`` of the message the
        reference builds in requests.py:41-48 (``order="deterministic"``), or list the map entries in
        the order given (``order="given"``).  ``grpc_frame=True`` puts gRPC's five-byte length-prefixed-message
        header in front (for a transport that writes HTTP/2 DATA frames itself; grpc-python adds it on its own).
        Inputs may be numpy arrays (pageable or ``pinned_empty``) or device arrays (``__cuda_array_interface__`` / DLPack: no
        host-to-device copy).  ``out="pinned"`` returns uint8 views of the codec's page-locked landing buffer instead of ``bytes``
        objects (no copy on the host; valid until the next encode on this codec).
        """
        keep, structs = self._build_requests(list(requests), order, wire_dtype, tensor_content, keep_snan, grpc_frame)
        n = len(structs)
        if n == 0:
            return []
        reqs = (N.Request * n)(*structs)
        if out is not None and out != "pinned":
            raise ValueError('out must be None or "pinned"')
        if n == 1 and out is None:
            # one request whose size is closed-form (no varint-packed input): copy straight into the bytes object
            total = C.c_uint64()
            if _HAVE_NEW_BYTES and self._lib.b200tfs_request_size(reqs, C.byref(total)) == N.OK:
                obj, addr = _new_bytes(int(total.value))
                off = (C.c_uint64 * 1)()
                ln = (C.c_uint64 * 1)()
                N.check(self._lib.b200tfs_encode_requests_host(self._ctx, 1, reqs, addr, total.value, off, ln))
                if off[0] != 0 or ln[0] != total.value:
                    raise N.NativeError(N.E_ARG, f"encoded length {ln[0]} at {off[0]} does not match the planned {total.value}")
                return [obj]
        cap = 0
        for preps, _, name in keep:
            cap += 1024 + len(name)
            for p in preps:
                cap += _wire_bound(p, tensor_content) + 512
        off = (C.c_uint64 * n)()
        ln = (C.c_uint64 * n)()
        if out == "pinned":
            pw = self._pinned_wire(cap)
            N.check(self._lib.b200tfs_encode_requests_host(self._ctx, n, reqs, pw.ptr, cap, off, ln))
            return [pw.array[off[i]: off[i] + ln[i]] for i in range(n)]
        wire = np.empty(cap, dtype=np.uint8)
        N.check(self._lib.b200tfs_encode_requests_host(self._ctx, n, reqs, wire.ctypes.data, cap, off, ln))
        return [wire[off[i]: off[i] + ln[i]].tobytes() for i in range(n)]

    def encode_predict_request(self, model_name: str, input_dict: Mapping, model_version: Optional[int] = None, **kw) -> bytes:
        return self.encode_predict_requests([(model_name, model_version, input_dict)], **kw)[0]

    # ---- decode --------------------------------------------------------------------------------
    def _pack_wires(self, wires: Sequence[bytes]):
        n = len(wires)
        off = (C.c_uint64 * max(n, 1))()
        ln = (C.c_uint64 * max(n, 1))()
        if n == 1 and len(wires[0]) and (isinstance(wires[0], (bytes, bytearray, memoryview)) or
                                         (isinstance(wires[0], np.ndarray) and wires[0].dtype == np.uint8 and wires[0].flags.c_contiguous)):
            # a single message: hand its own buffer to the library (read-only view, no copy)
            view = wires[0] if isinstance(wires[0], np.ndarray) else np.frombuffer(wires[0], dtype=np.uint8)
            ln[0] = view.size
            return view, off, ln
        cur = 0
        for i, w in enumerate(wires):
            off[i] = cur
            ln[i] = len(w)
            cur += (len(w) + 255) & ~255  # records start 256-byte aligned, like a received buffer would
        buf = np.empty(max(cur, 1), dtype=np.uint8)
        for i, w in enumerate(wires):
            if len(w):
                buf[off[i]: off[i] + len(w)] = w if isinstance(w, np.ndarray) else np.frombuffer(w, dtype=np.uint8)
        return buf, off, ln

    @staticmethod
    def _text(buf: np.ndarray, off: int, length: int) -> str:
        return buf[off: off + length].tobytes().decode("utf-8")

    def parse_predict_responses(self, wires: Sequence[bytes], max_outputs: int = 16) -> List[ParsedResponse]:
        """Run the parse kernel over each PredictResponse; raises DecodeError like ``FromString``.  ``max_outputs`` sizes
        the first attempt only: a response with more outputs is parsed again with a table twice as wide, and so on
        (``PredictResponse.FromString`` has no limit on the map size)."""
        n = len(wires)
        if n == 0:
            return []
        buf, off, ln = self._pack_wires(wires)
        while True:
            outs = (N.Output * (n * max_outputs))()
            n_outs = (C.c_int32 * n)()
            specs = (N.ModelSpec * n)()
            status = (C.c_int32 * n)()
            N.check(self._lib.b200tfs_parse_responses_host(self._ctx, buf.ctypes.data, n, off, ln, max_outputs, outs, n_outs, specs, status))
            if max_outputs < (1 << 20) and any(status[i] == N.E_SIZE for i in range(n)):
                max_outputs *= 2
                continue
            break
        parsed = []
        for i in range(n):
            if status[i] == N.E_SIZE:
                raise ValueError(f"response {i} has more than max_outputs={max_outputs} outputs")
            if status[i] != N.OK:
                from google.protobuf.message import DecodeError

                raise DecodeError(f"Error parsing message (response {i}, status {status[i]})")
            table = {}
            base = int(off[i])  # table offsets are relative to the record
            for j in range(n_outs[i]):
                o = outs[i * max_outputs + j]
                table[self._text(buf, base + o.key_off, o.key_len)] = o
            s = specs[i]
            spec = DecodedSpec(self._text(buf, base + s.name_off, s.name_len), int(s.version), bool(s.has_version),
                               self._text(buf, base + s.label_off, s.label_len), self._text(buf, base + s.signature_off, s.signature_len))
            parsed.append(ParsedResponse(buf, int(off[i]), int(ln[i]), int(status[i]), table, spec))
        return parsed

    def _resolve_output(self, o: N.Output, strict: bool, out_dtype):
        """(numpy dtype, native dst dtype code, shape) for one tabulated output, or raise what the
        reference raises for it (tensors.py:42-46, types.py:39-40)."""
        enum = int(o.dtype)
        if strict and enum == DT_BFLOAT16:
            raise KeyError(enum)  # not in the reference's ENUM_TO_TF_MAPPING
        if o.status == N.E_KEY:
            raise KeyError(enum)
        np_type = numpy_for_enum(enum)
        shape = self._shape(o)
        if strict and enum in (DT_COMPLEX64, DT_COMPLEX128) and o.n_elems:
            raise ValueError("cannot reshape array: the reference reads complex values as separate floats")
        content_only = o.n_runs == 0 and o.content_len and o.n_strings == 0
        if content_only and not strict and o.content_len == int(np.prod(shape, dtype=np.int64)) * np.dtype(np_type).itemsize:
            pass  # tolerant: raw little-endian tensor_content, as TF writes it
        elif o.status == N.E_SHAPE and not strict and out_dtype is None and self._may_pad(o, np_type):
            # tolerant: TensorFlow's MakeNdarray padding - no values: zeros; fewer than the shape holds: the last one repeats
            o.flags |= N.OF_PAD_EDGE
        elif o.status == N.OK and not strict and out_dtype is None and (o.flags & N.OF_VARINT) and o.n_elems > 0:
            o.flags |= N.OF_PAD_EDGE   # packed varints: whether there are fewer values than elements is only known to the decode kernels
        elif o.status == N.E_SHAPE:
            raise ValueError(f"cannot reshape array into shape {shape}")
        elif o.status == N.E_NONCANONICAL:
            raise NotImplementedError("wire layout not tabulated by the device parser")
        elif o.status != N.OK:
            N.check(o.status)
        if (o.flags & N.OF_RANK0) and strict:
            raise TypeError("reshape() takes exactly 1 argument (0 given)")
        dst_code = enum
        if out_dtype is not None:
            dst_code = _as_enum(out_dtype)
            np_type = numpy_for_enum(dst_code)
        elif strict and enum == DT_HALF:
            dst_code = N.DT_HALF_REFQUIRK
        return np_type, dst_code, shape

    def _shape(self, o: N.Output) -> Tuple[int, ...]:
        """Every dim of an output: the table holds MAX_RANK inline, deeper shapes come from the context's spill area."""
        if o.rank <= N.MAX_RANK:
            return tuple(int(o.dims[k]) for k in range(o.rank))
        dims = (C.c_int64 * o.rank)()
        N.check(self._lib.b200tfs_output_dims(self._ctx, C.byref(o), dims, o.rank))
        return tuple(int(d) for d in dims)

    def _runs(self, o: N.Output) -> List[N.Run]:
        """Every value run of an output (inline + spilled), in wire order."""
        if o.n_runs <= o.n_inline:
            return [o.runs[k] for k in range(o.n_runs)]
        runs = (N.Run * o.n_runs)()
        N.check(self._lib.b200tfs_output_runs(self._ctx, C.byref(o), runs, o.n_runs))
        return list(runs)

    def _may_pad(self, o: N.Output, np_type) -> bool:
        """An E_SHAPE output the padding rule applies to: the shape is fully known and holds MORE elements than there are
        values (packed varints: more elements than value bytes would be needed; the kernel then counts exactly)."""
        if o.n_elems <= 0 or o.content_len or any(d < 0 for d in self._shape(o)):
            return False
        value_bytes = sum(int(r.len) * int(r.count) for r in self._runs(o))
        if o.flags & N.OF_VARINT:
            return True     # fewer value bytes than elements was what raised E_SHAPE here; a surplus is caught by the decode kernels
        size = np.dtype(np_type).itemsize
        return value_bytes % size == 0 and value_bytes // size < o.n_elems

    def decode_predict_responses(self, wires: Sequence[bytes], *, strict: bool = False, out_dtypes: Optional[Mapping] = None,
                                 max_outputs: int = 16) -> List[Tuple[Dict[str, np.ndarray], DecodedSpec]]:
        """``PredictResponse.FromString`` + ``tensor_proto_to_ndarray`` on every output, on the GPU.

        ``strict=True`` reproduces the reference's behaviour case for case (including the inputs it
        rejects); the default additionally accepts what TF itself emits: ``tensor_content``, rank-0
        tensors, complex pairs, bfloat16, and reads ``half_val`` as bit patterns.
        """
        if out_dtypes is None and len(wires):
            fused = self._decode_fused(wires, strict)
            if fused is not None:
                return fused
        elif out_dtypes and len(wires) and not strict:
            # every requested cast is the same narrowing of float32 (BASELINE config C4): one launch does it (b200tfs_set_decode_cast)
            try:
                wanted = {int(enum_for_numpy(np.dtype(v).type)) for v in out_dtypes.values()}
            except (KeyError, ValueError, TypeError):
                wanted = set()
            if len(wanted) == 1 and next(iter(wanted)) in (int(DT_HALF), int(DT_BFLOAT16)):
                fused = self._decode_fused(wires, strict, cast=(next(iter(wanted)), dict(out_dtypes)))
                if fused is not None:
                    return fused
        parsed = self.parse_predict_responses(wires, max_outputs=max_outputs)
        if not parsed:
            return []
        jobs = []  # (response idx, key, Output, np_type, dst_code, shape)
        results: List[Tuple[Dict[str, np.ndarray], DecodedSpec]] = []
        for i, pr in enumerate(parsed):
            results.append(({}, pr.model_spec))
            for key, o in pr.outputs.items():
                od = out_dtypes.get(key) if out_dtypes else None
                if int(o.dtype) == DT_STRING and o.status == N.OK:
                    results[i][0][key] = self._decode_strings(pr.wire, pr.offset, o)
                    continue
                np_type, dst_code, shape = self._resolve_output(o, strict, od)
                jobs.append((i, key, o, np_type, dst_code, shape))
        if jobs:
            m = len(jobs)
            outs = (N.Output * m)(*[j[2] for j in jobs])
            arrays = [np.empty(j[5], dtype=j[3]) for j in jobs]
            dst = (C.c_void_p * m)(*[a.ctypes.data if a.size else None for a in arrays])
            codes = (C.c_int32 * m)(*[j[4] for j in jobs])
            status = (C.c_int32 * m)()
            rec = (C.c_uint64 * m)(*[parsed[j[0]].offset for j in jobs])
            N.check(self._lib.b200tfs_unpack_outputs_host(self._ctx, m, outs, rec, dst, codes, status))
            for k, (i, key, o, np_type, dst_code, shape) in enumerate(jobs):
                if status[k] == N.E_SHAPE:
                    raise ValueError(f"cannot reshape array into shape {shape}")
                N.check(status[k])
                results[i][0][key] = arrays[k]
        return results

    def _fused_launch(self, wires: Sequence[bytes], cast_code: int = 0):
        """H2D of the wire, decode_fused_kernel, D2H of the decoded fixed-width outputs, one synchronise.  None if any
        record was not tabulated (malformed, or more than FUSED_MAX_OUTPUTS outputs)."""
        n = len(wires)
        buf, off, ln = self._pack_wires(wires)
        K = N.FUSED_MAX_OUTPUTS
        stride = (max(int(ln[i]) for i in range(n)) + 256 * (K + 1) + 255) & ~255   # every fixed output fits, each 256-aligned
        dst = np.empty(n * stride, dtype=np.uint8)
        if cast_code:
            N.check(self._lib.b200tfs_set_decode_cast(self._ctx, cast_code))
        try:
            N.check(self._lib.b200tfs_decode_responses_host_async(self._ctx, buf.ctypes.data, n, off, ln, dst.ctypes.data, stride))
        finally:
            if cast_code:
                N.check(self._lib.b200tfs_set_decode_cast(self._ctx, 0))
        outs = (N.Output * (n * K))()
        n_outs = (C.c_int32 * n)()
        specs = (N.ModelSpec * n)()
        status = (C.c_int32 * n)()
        N.check(self._lib.b200tfs_decode_results(self._ctx, n, outs, n_outs, specs, status))   # synchronises
        if any(status[i] != N.OK for i in range(n)):
            return None
        return n, buf, off, dst, stride, outs, n_outs, specs

    def open_predict_response(self, wire: bytes) -> Optional["OpenResponse"]:
        """Decode one PredictResponse now, hand its outputs out later (``PredictResponseView.outputs``): the launch moves
        every fixed-width output; ``OpenResponse.array(key, strict)`` applies the reference's per-output rules on demand."""
        if not len(wire):
            return None
        launched = self._fused_launch([wire])
        if launched is None:
            return None
        _, buf, off, dst, _, outs, n_outs, _ = launched
        table = {}
        for j in range(n_outs[0]):
            o = outs[j]
            table[self._text(buf, int(off[0]) + o.key_off, o.key_len)] = o
        return OpenResponse(self, buf, int(off[0]), dst, table)

    def _decode_fused(self, wires: Sequence[bytes], strict: bool, cast=None):
        """One launch, one synchronise: tag walk (or framing-template check), destination layout and the move of every
        fixed-width output in ``decode_fused_kernel``; the outputs come back as views of one host buffer.  Varint-packed
        and tensor_content-only outputs are tabulated by the same launch and unpacked by a second one.  Returns None when
        a record needs the two-phase path (more than eight outputs, a malformed record: that path raises what the
        reference raises)."""
        cast_code, cast_keys = cast if cast else (0, {})
        launched = self._fused_launch(wires, cast_code)
        if launched is None:
            return None
        n, buf, off, dst, stride, outs, n_outs, specs = launched
        K = N.FUSED_MAX_OUTPUTS
        results: List[Tuple[Dict[str, np.ndarray], DecodedSpec]] = []
        jobs = []
        for i in range(n):
            base = int(off[i])
            s = specs[i]
            spec = DecodedSpec(self._text(buf, base + s.name_off, s.name_len), int(s.version), bool(s.has_version),
                               self._text(buf, base + s.label_off, s.label_len), self._text(buf, base + s.signature_off, s.signature_len))
            arrays: Dict[str, np.ndarray] = {}
            results.append((arrays, spec))
            for j in range(n_outs[i]):
                o = outs[i * K + j]
                key = self._text(buf, base + o.key_off, o.key_len)
                if int(o.dtype) == DT_STRING and o.status == N.OK:
                    arrays[key] = self._decode_strings(buf, base, o)
                    continue
                if cast_code and ((int(o.dtype) == 1) != (key in cast_keys)):
                    return None      # the launch narrowed every float32 output: only right when exactly those were asked for
                np_type, dst_code, shape = self._resolve_output(o, strict, cast_keys.get(key))
                if o.status == N.OK and int(o.dtype) in _FUSED_MOVES and o.n_runs and o.n_elems and \
                        (dst_code == int(o.dtype) or (cast_code and int(o.dtype) == 1 and dst_code == cast_code)):
                    at = i * stride + int(o.dst_off)
                    arrays[key] = dst[at: at + int(o.dst_bytes)].view(np_type).reshape(shape)
                else:
                    jobs.append((i, key, o, np_type, dst_code, shape))
        if jobs:
            m = len(jobs)
            o_arr = (N.Output * m)(*[j[2] for j in jobs])
            made = [np.empty(j[5], dtype=j[3]) for j in jobs]
            ptrs = (C.c_void_p * m)(*[a.ctypes.data if a.size else None for a in made])
            codes = (C.c_int32 * m)(*[j[4] for j in jobs])
            st = (C.c_int32 * m)()
            rec = (C.c_uint64 * m)(*[int(off[j[0]]) for j in jobs])
            N.check(self._lib.b200tfs_unpack_outputs_host(self._ctx, m, o_arr, rec, ptrs, codes, st))
            for k, (i, key, o, np_type, dst_code, shape) in enumerate(jobs):
                if st[k] == N.E_SHAPE:
                    raise ValueError(f"cannot reshape array into shape {shape}")
                N.check(st[k])
                results[i][0][key] = made[k]
        return results

    @staticmethod
    def _decode_strings(buf: np.ndarray, base: int, o: N.Output) -> np.ndarray:
        from tensorflow.core.framework.tensor_pb2 import TensorProto

        proto = TensorProto.FromString(buf[base + o.msg_off: base + o.msg_off + o.msg_len].tobytes())
        shape = tuple(int(d.size) for d in proto.tensor_shape.dim)     # the host message is at hand: any rank
        return np.array([e for e in proto.string_val], dtype=np.str_).reshape(*shape)

    def decode_predict_response(self, wire: bytes, *, out: Optional[Mapping[str, np.ndarray]] = None, **kw) -> Tuple[Dict[str, np.ndarray], DecodedSpec]:
        """One response.  ``out={key: array}``: the named outputs are written into the caller's arrays (dtype and shape must
        match); when the response has that one fixed-width output and the array came from ``pinned_empty``, the device-to-host
        copy lands in it directly (no staging buffer, no extra copy on the host)."""
        if out:
            direct = self._decode_into_pinned(wire, out, kw.get("strict", False)) if len(out) == 1 and not kw.get("out_dtypes") else None
            if direct is not None:
                return direct
            arrays, spec = self.decode_predict_responses([wire], **kw)[0]
            for k, dst in out.items():
                if k not in arrays:
                    raise KeyError(k)
                if dst.dtype != arrays[k].dtype or dst.shape != arrays[k].shape:
                    raise ValueError(f"out[{k!r}]: {dst.dtype}{dst.shape} does not match the decoded {arrays[k].dtype}{arrays[k].shape}")
                np.copyto(dst, arrays[k])
                arrays[k] = dst
            return arrays, spec
        return self.decode_predict_responses([wire], **kw)[0]

    def _decode_into_pinned(self, wire, out: Mapping[str, np.ndarray], strict: bool):
        (key, arr), = out.items()
        cap = self._pinned.capacity(arr)
        if cap is None or not len(wire):
            return None
        buf, off, ln = self._pack_wires([wire])
        stride = cap & ~255
        N.check(self._lib.b200tfs_decode_responses_host_async(self._ctx, buf.ctypes.data, 1, off, ln, arr.ctypes.data, stride))
        K = N.FUSED_MAX_OUTPUTS
        outs, n_outs, specs, status = (N.Output * K)(), (C.c_int32 * 1)(), (N.ModelSpec * 1)(), (C.c_int32 * 1)()
        N.check(self._lib.b200tfs_decode_results(self._ctx, 1, outs, n_outs, specs, status))
        if status[0] != N.OK or n_outs[0] != 1:
            return None                                  # not the single-output case after all: the general path redoes it
        o = outs[0]
        if self._text(buf, o.key_off, o.key_len) != key or int(o.dtype) not in _FUSED_MOVES or o.status != N.OK or not o.n_runs or o.dst_off != 0:
            return None
        np_type, dst_code, shape = self._resolve_output(o, strict, None)
        if dst_code != int(o.dtype) or np.dtype(np_type) != arr.dtype or tuple(shape) != arr.shape:
            return None
        s = specs[0]
        spec = DecodedSpec(self._text(buf, s.name_off, s.name_len), int(s.version), bool(s.has_version),
                           self._text(buf, s.label_off, s.label_len), self._text(buf, s.signature_off, s.signature_len))
        return {key: arr}, spec

    def decode_tensor_protos(self, wires: Sequence[bytes], *, strict: bool = False, out_dtype=None) -> List[np.ndarray]:
        """``tensor_proto_to_ndarray`` for serialised TensorProto messages."""
        n = len(wires)
        if n == 0:
            return []
        buf, off, ln = self._pack_wires(wires)
        outs = (N.Output * n)()
        status = (C.c_int32 * n)()
        N.check(self._lib.b200tfs_parse_tensor_protos_host(self._ctx, buf.ctypes.data, n, off, ln, outs, status))
        results: List[Optional[np.ndarray]] = [None] * n
        jobs = []
        for i in range(n):
            N.check(status[i])
            o = outs[i]
            if int(o.dtype) == DT_STRING and o.status == N.OK:
                results[i] = self._decode_strings(buf, int(off[i]), o)
                continue
            np_type, dst_code, shape = self._resolve_output(o, strict, out_dtype)
            jobs.append((i, o, np_type, dst_code, shape))
        if jobs:
            m = len(jobs)
            o_arr = (N.Output * m)(*[j[1] for j in jobs])
            arrays = [np.empty(j[4], dtype=j[2]) for j in jobs]
            dst = (C.c_void_p * m)(*[a.ctypes.data if a.size else None for a in arrays])
            codes = (C.c_int32 * m)(*[j[3] for j in jobs])
            st = (C.c_int32 * m)()
            rec = (C.c_uint64 * m)(*[int(off[j[0]]) for j in jobs])
            N.check(self._lib.b200tfs_unpack_outputs_host(self._ctx, m, o_arr, rec, dst, codes, st))
            for k, (i, o, np_type, dst_code, shape) in enumerate(jobs):
                if st[k] == N.E_SHAPE:
                    raise ValueError(f"cannot reshape array into shape {shape}")
                N.check(st[k])
                results[i] = arrays[k]
        return results  # type: ignore[return-value]


# ---- zero-copy helpers -----------------------------------------------------------------------------
# CPython only: PyBytes_FromStringAndSize(NULL, n) is the documented way to make a bytes object whose buffer the caller
# fills before anyone else sees the object (https://docs.python.org/3/c-api/bytes.html).  The object is not hashed,
# interned or shared until the library has written every byte of it; other interpreters take the copying path.
import platform as _platform

_HAVE_NEW_BYTES = _platform.python_implementation() == "CPython" and hasattr(C, "pythonapi")
if _HAVE_NEW_BYTES:
    _PyBytes_FromStringAndSize = C.pythonapi.PyBytes_FromStringAndSize
    _PyBytes_FromStringAndSize.restype = C.py_object
    _PyBytes_FromStringAndSize.argtypes = [C.c_void_p, C.c_ssize_t]
    _PyBytes_AsString = C.pythonapi.PyBytes_AsString
    _PyBytes_AsString.restype = C.c_void_p
    _PyBytes_AsString.argtypes = [C.py_object]


def _new_bytes(n: int):
    """An uninitialised bytes object of n bytes and the address of its buffer: the device-to-host copy lands
    directly in the object grpc will send (no intermediate buffer, no .tobytes())."""
    obj = _PyBytes_FromStringAndSize(None, n)
    return obj, _PyBytes_AsString(obj)


_tls = threading.local()


def get_codec(device: int = 0) -> Codec:
    """Per-thread codec on `device` (contexts are not re-entrant)."""
    table = getattr(_tls, "codecs", None)
    if table is None:
        table = _tls.codecs = {}
    c = table.get(device)
    if c is None:
        c = table[device] = Codec(device)
    return c
