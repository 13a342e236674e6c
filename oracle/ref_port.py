"""Python port of the reference's Predict codec over the protobuf runtime.  TEST INFRASTRUCTURE ONLY.

Only ``tests/`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs import this; the
product never does.  It exists because the reference is Python and cannot travel to the GPU box
(``/root/reference`` is absent there): this file performs the same work per element - one Python
object per tensor element appended to a protobuf repeated field, one more per element when reading
back - on the same third-party runtimes (protobuf, numpy), so timing it stands in for timing the
reference (``cpu_baseline.kind = "port"``).  Parity status: PINNED - ``tests/test_oracle.py`` checks
its bytes against every golden vector the unmodified reference produced.

Restates (paths relative to the reference checkout):
  tensor_serving_client/min_tfs_client/tensors.py:17-35   ndarray -> TensorProto, element by element
  tensor_serving_client/min_tfs_client/tensors.py:38-46   TensorProto -> ndarray, element by element
  tensor_serving_client/min_tfs_client/requests.py:41-48  PredictRequest assembly (CopyFrom per input)
  protobuf_srcs/tensorflow_serving/apis/prediction_service_pb2_grpc.py:52-53  SerializeToString / FromString
  tensor_serving_client/min_tfs_client/constants.py:13-29 dtype table
"""
import os
import sys

import numpy as np

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "min-tfs-client_b200")
if _PKG not in sys.path:  # schema modules (the reference generates these with protoc at install time)
    sys.path.insert(0, _PKG)

from tensorflow.core.framework import tensor_pb2, tensor_shape_pb2, types_pb2  # noqa: E402
from tensorflow_serving.apis import predict_pb2  # noqa: E402

# numpy scalar type -> (DataType enum, TensorProto field, values are numbers)
TABLE = {
    np.float16: (types_pb2.DT_HALF, "half_val", True), np.float32: (types_pb2.DT_FLOAT, "float_val", True),
    np.float64: (types_pb2.DT_DOUBLE, "double_val", True), np.int8: (types_pb2.DT_INT8, "int_val", True),
    np.int16: (types_pb2.DT_INT16, "int_val", True), np.int32: (types_pb2.DT_INT32, "int_val", True),
    np.int64: (types_pb2.DT_INT64, "int64_val", True), np.uint8: (types_pb2.DT_UINT8, "int_val", True),
    np.uint16: (types_pb2.DT_UINT16, "int_val", True), np.uint32: (types_pb2.DT_UINT32, "uint32_val", True),
    np.uint64: (types_pb2.DT_UINT64, "uint64_val", True), np.complex64: (types_pb2.DT_COMPLEX64, "scomplex_val", True),
    np.complex128: (types_pb2.DT_COMPLEX128, "dcomplex_val", True), np.str_: (types_pb2.DT_STRING, "string_val", False),
    np.bool_: (types_pb2.DT_BOOL, "bool_val", True),
}
BY_ENUM = {row[0]: (np_type, row[1]) for np_type, row in TABLE.items()}


def to_tensor_proto(x: np.ndarray) -> tensor_pb2.TensorProto:
    if x.dtype.type not in TABLE:
        raise ValueError(f"Dtype {x.dtype.type.__name__} is not valid")
    enum, field, numeric = TABLE[x.dtype.type]
    dims = [tensor_shape_pb2.TensorShapeProto.Dim(size=n) for n in x.shape]
    message = tensor_pb2.TensorProto(dtype=enum, tensor_shape=tensor_shape_pb2.TensorShapeProto(dim=dims))
    flat = x.ravel()
    target = getattr(message, field)
    if numeric:
        target.extend([element.item() for element in flat])  # the reference's hot loop #1: one Python object per element
    else:
        target.extend([s.encode("utf-8") if isinstance(s, str) else s for s in flat])
    return message


def from_tensor_proto(message: tensor_pb2.TensorProto) -> np.ndarray:
    np_type, field = BY_ENUM[message.dtype]  # KeyError for unmapped enums, as in the reference
    shape = tuple(int(d.size) for d in message.tensor_shape.dim)
    elements = [e for e in getattr(message, field)]  # hot loop #2
    return np.array(elements, dtype=np_type).reshape(*shape)


def request_message(model_name, model_version, inputs):
    """inputs: iterable of (key, ndarray)."""
    request = predict_pb2.PredictRequest()
    request.model_spec.name = model_name
    if model_version is not None:
        request.model_spec.version.value = model_version
    for key, value in inputs:
        request.inputs[key].CopyFrom(to_tensor_proto(value))
    return request


def encode_predict_request(model_name, model_version, inputs, deterministic=None) -> bytes:
    inputs = list(inputs)
    if deterministic is None:
        deterministic = len(inputs) > 1  # the only order that is stable across processes
    return request_message(model_name, model_version, inputs).SerializeToString(deterministic=deterministic)


def encode_tensor_proto(x) -> bytes:
    return to_tensor_proto(np.asarray(x)).SerializeToString()


def decode_predict_response(wire: bytes):
    response = predict_pb2.PredictResponse.FromString(wire)
    return {key: from_tensor_proto(response.outputs[key]) for key in response.outputs}


def decode_tensor_proto(wire: bytes):
    return from_tensor_proto(tensor_pb2.TensorProto.FromString(wire))


def make_ndarray_tf(wire: bytes) -> np.ndarray:
    """TensorFlow's MakeNdarray for the numeric typed fields, restated over the protobuf runtime from the reference's vendored
    protobuf_srcs/tensorflow/python/framework/tensor_util.py:565-642: tensor_content wins; else the typed field; no values ->
    zeros; fewer values than the shape holds -> np.pad(..., "edge").  Pins the tolerant decoder's padding rule."""
    tp = tensor_pb2.TensorProto.FromString(wire)
    shape = [d.size for d in tp.tensor_shape.dim]
    n = int(np.prod(shape, dtype=np.int64))
    dt = np.dtype({1: "float32", 2: "float64", 3: "int32", 4: "uint8", 5: "int16", 6: "int8", 9: "int64", 10: "bool", 17: "uint16",
                   22: "uint32", 23: "uint64", 19: "float16", 8: "complex64", 18: "complex128"}[tp.dtype])
    if tp.tensor_content:
        return np.frombuffer(tp.tensor_content, dtype=dt).copy().reshape(shape)
    if tp.dtype == 19:
        values = np.fromiter(tp.half_val, dtype=np.uint16).view(np.float16)
    elif tp.dtype == 1:
        values = np.fromiter(tp.float_val, dtype=dt)
    elif tp.dtype == 2:
        values = np.fromiter(tp.double_val, dtype=dt)
    elif tp.dtype in (3, 4, 5, 6, 17):
        values = np.fromiter(tp.int_val, dtype=dt)
    elif tp.dtype == 9:
        values = np.fromiter(tp.int64_val, dtype=dt)
    elif tp.dtype == 22:
        values = np.fromiter(tp.uint32_val, dtype=dt)
    elif tp.dtype == 23:
        values = np.fromiter(tp.uint64_val, dtype=dt)
    elif tp.dtype == 10:
        values = np.fromiter(tp.bool_val, dtype=dt)
    else:
        it = iter(tp.scomplex_val if tp.dtype == 8 else tp.dcomplex_val)
        values = np.array([complex(a, b) for a, b in zip(it, it)], dtype=dt)
    if values.size == 0:
        return np.zeros(shape, dt)
    if values.size != n:
        values = np.pad(values, (0, n - values.size), "edge")
    return values.reshape(shape)


def response_message(outputs, model_name="default", version=1, signature="serving_default") -> bytes:
    """A PredictResponse as a server would send it (for round-trip timing)."""
    response = predict_pb2.PredictResponse()
    for key, value in outputs:
        response.outputs[key].CopyFrom(to_tensor_proto(value))
    response.model_spec.name = model_name
    response.model_spec.version.value = version
    response.model_spec.signature_name = signature
    return response.SerializeToString(deterministic=True)
