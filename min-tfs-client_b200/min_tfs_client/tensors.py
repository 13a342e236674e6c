"""Drop-in for the reference's ``min_tfs_client/tensors.py`` with the work done on the GPU.

Same public names and signatures (reference tensors.py:10-46):
``coerce_to_bytes, write_values_to_tensor_proto, ndarray_to_tensor_proto, extract_shape,
tensor_proto_to_ndarray`` - plus the TensorFlow-style aliases ``make_tensor_proto`` / ``make_ndarray``
and byte-level entry points that skip the protobuf message object altogether.
"""
from typing import AnyStr, Iterable, Optional, Tuple, Union

import numpy as np
from tensorflow.core.framework.tensor_pb2 import TensorProto

from .codec import get_codec
from .types import DataType


def coerce_to_bytes(text: AnyStr) -> bytes:
    """str -> UTF-8 bytes, bytes unchanged (reference tensors.py:10-14)."""
    return text.encode("utf-8") if isinstance(text, str) else text


def ndarray_to_tensor_proto_bytes(ndarray: np.ndarray, **options) -> bytes:
    """Serialised TensorProto for `ndarray`: what ``ndarray_to_tensor_proto(x).SerializeToString()``
    returns in the reference, produced by the encode kernels."""
    return get_codec().encode_tensor_protos([ndarray], **options)[0]


def ndarray_to_tensor_proto(ndarray: np.ndarray, **options) -> TensorProto:
    """ndarray -> TensorProto (reference tensors.py:28-35)."""
    return TensorProto.FromString(ndarray_to_tensor_proto_bytes(ndarray, **options))


def write_values_to_tensor_proto(tensor_proto: TensorProto, values: Iterable, dtype: DataType) -> TensorProto:
    """Fill the typed repeated field of `tensor_proto` from `values` (reference tensors.py:17-25).

    The packed field is produced by the encode kernel for a 1-D tensor of `values` and merged into the
    message, so the caller's dtype/shape fields are left as they are.
    """
    arr = np.asarray(list(values) if not isinstance(values, np.ndarray) else values)
    if not dtype.is_numeric:
        getattr(tensor_proto, dtype.proto_field_name).extend(coerce_to_bytes(v) for v in arr.ravel().tolist())
        return tensor_proto
    arr = arr.astype(dtype.numpy_dtype, copy=False).ravel()
    filled = TensorProto.FromString(ndarray_to_tensor_proto_bytes(arr))
    getattr(tensor_proto, dtype.proto_field_name).extend(getattr(filled, dtype.proto_field_name))
    return tensor_proto


def extract_shape(tensor_proto: TensorProto) -> Tuple[int, ...]:
    """Shape tuple of a TensorProto (reference tensors.py:38-39)."""
    return tuple(int(d.size) for d in tensor_proto.tensor_shape.dim)


def tensor_proto_to_ndarray(tensor_proto: Union[TensorProto, bytes, "WireTensor"], *, strict: bool = True, **options) -> np.ndarray:
    """TensorProto -> ndarray (reference tensors.py:42-46), decoded by the parse + unpack kernels.

    Accepts a TensorProto message, its serialised bytes, or a ``WireTensor`` handed out by
    ``TensorServingClient.predict_request`` (the zero-reparse path).  ``strict`` (default) keeps the
    reference's behaviour on the inputs it rejects; ``strict=False`` also accepts ``tensor_content``,
    rank-0 tensors, complex, bfloat16 and reads ``half_val`` as bit patterns (TF's conventions).
    """
    if isinstance(tensor_proto, WireTensor):
        return tensor_proto.to_ndarray(strict=strict, **options)
    wire = tensor_proto if isinstance(tensor_proto, (bytes, bytearray, memoryview)) else tensor_proto.SerializeToString()
    return get_codec().decode_tensor_protos([bytes(wire)], strict=strict, **options)[0]


class WireTensor:
    """One output of a response: wire bytes, decoded on demand by the GPU - or, when the response was opened by the fused
    decode launch (``PredictResponseView.outputs``), already decoded and handed out here."""

    def __init__(self, wire: Optional[bytes] = None, opened=None, key=None):
        self._bytes, self._opened, self._key = wire, opened, key

    @property
    def _wire(self) -> bytes:
        if self._bytes is None:
            self._bytes = self._opened.wire_of(self._key)
        return self._bytes

    def to_ndarray(self, strict: bool = False, **options) -> np.ndarray:
        if self._opened is not None and not options:
            arr = self._opened.array(self._key, strict)
            if arr is not None:
                return arr
        return get_codec().decode_tensor_protos([self._wire], strict=strict, **options)[0]

    def to_proto(self) -> TensorProto:
        return TensorProto.FromString(self._wire)

    def SerializeToString(self) -> bytes:  # noqa: N802 - protobuf spelling
        return self._wire

    def __getattr__(self, name):  # dtype, tensor_shape, float_val, ... through a real message
        return getattr(self.to_proto(), name)


# TensorFlow's names for the same two operations (tensor_util.py:356, :565 in the vendored tree)
make_tensor_proto = ndarray_to_tensor_proto
make_ndarray = tensor_proto_to_ndarray
