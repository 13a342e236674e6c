"""Drop-in for the Predict part of the reference's ``min_tfs_client/requests.py``.

``TensorServingClient(host, port, credentials=None).predict_request(model_name, input_dict,
timeout=60, model_version=None)`` keeps the reference's signature (requests.py:22-65).  The request
is packed by the encode kernels and sent as raw bytes through the same gRPC method the generated stub
binds (``/tensorflow.serving.PredictionService/Predict``, prediction_service_pb2_grpc.py:50-54); the
response bytes are handed to the parse/unpack kernels.

The reference's other three calls are kept as well (requests.py:67-110).  They are outside the Predict hot path
(SURVEY.md 8(f) rank 4): small pointer-chasing messages assembled on the host by the protobuf runtime, like
``DT_STRING`` tensors.  ``model_status_request`` is the reference's; ``classification_request`` /
``regression_request`` cannot work in the reference (it fills ``request.inputs[k]``, a field neither
``ClassificationRequest`` nor ``RegressionRequest`` has, and sends them to ``Predict``): here they build the
``Input{example_list}`` those RPCs define - one ``tf.Example`` per row of ``input_dict`` - and call
``PredictionService/Classify`` and ``/Regress``.
"""
from typing import Dict, Optional

import numpy as np

from .codec import get_codec
from .tensors import WireTensor

PREDICT_METHOD = "/tensorflow.serving.PredictionService/Predict"
CLASSIFY_METHOD = "/tensorflow.serving.PredictionService/Classify"
REGRESS_METHOD = "/tensorflow.serving.PredictionService/Regress"
MODEL_STATUS_METHOD = "/tensorflow.serving.ModelService/GetModelStatus"


def examples_from_input_dict(input_dict: Dict[str, np.ndarray]):
    """``Input{example_list{examples}}`` for Classify / Regress (input.proto:13-79, example.proto, feature.proto).

    Row i of every array is example i: ``feature[k]`` holds the row's values flattened - ``float_list`` for floating
    dtypes, ``int64_list`` for integers and bools, ``bytes_list`` for str / bytes (``coerce_to_bytes``).  0-d arrays are
    repeated in every example; all other arrays must agree on their first dimension.
    """
    from tensorflow_serving.apis.input_pb2 import Input

    from .tensors import coerce_to_bytes

    arrays = {k: np.asarray(v) for k, v in input_dict.items()}
    rows = {a.shape[0] for a in arrays.values() if a.ndim}
    if len(rows) > 1:
        raise ValueError(f"inputs disagree on the number of examples: {sorted(rows)}")
    n = rows.pop() if rows else (1 if arrays else 0)
    inp = Input()
    inp.example_list.SetInParent()
    for i in range(n):
        ex = inp.example_list.examples.add()
        for k, a in arrays.items():
            row = a if a.ndim == 0 else a[i]
            feat = ex.features.feature[k]
            if row.dtype.kind == "f":
                feat.float_list.value.extend(np.asarray(row, dtype=np.float32).ravel().tolist())
            elif row.dtype.kind in "iub":
                feat.int64_list.value.extend(np.asarray(row, dtype=np.int64).ravel().tolist())
            elif row.dtype.kind in "US":
                feat.bytes_list.value.extend(coerce_to_bytes(s) for s in np.asarray(row).ravel().tolist())
            else:
                raise ValueError(f"input {k!r}: dtype {row.dtype} has no tf.Example feature kind")
    return inp


class PredictResponseView:
    """What ``predict_request`` returns: ``.outputs[key]`` (decode with ``tensor_proto_to_ndarray``),
    ``.model_spec``; anything else is answered by a real ``PredictResponse`` parsed on demand."""

    def __init__(self, wire: bytes):
        self._wire = wire
        self._proto = None
        self._views = None

    def to_proto(self):
        if self._proto is None:
            from tensorflow_serving.apis.predict_pb2 import PredictResponse

            self._proto = PredictResponse.FromString(self._wire)
        return self._proto

    @property
    def outputs(self) -> Dict[str, WireTensor]:
        if self._views is None:
            opened = get_codec().open_predict_response(self._wire)   # one launch decodes every fixed-width output
            if opened is not None:
                self._views = {k: WireTensor(opened=opened, key=k) for k in opened.table}
            else:   # empty / malformed / more outputs than the fused launch tabulates: the two-phase path (raises like FromString)
                parsed = get_codec().parse_predict_responses([self._wire])[0]
                buf, base = parsed.wire, parsed.offset
                self._views = {k: WireTensor(buf[base + o.msg_off: base + o.msg_off + o.msg_len].tobytes()) for k, o in parsed.outputs.items()}
        return self._views

    def to_ndarrays(self, **options) -> Dict[str, np.ndarray]:
        """Every output decoded in one parse + one unpack launch."""
        return get_codec().decode_predict_response(self._wire, **options)[0]

    def SerializeToString(self) -> bytes:  # noqa: N802
        return self._wire

    def __getattr__(self, name):
        return getattr(self.to_proto(), name)


def gpu_request_serializer(request) -> bytes:
    """``request_serializer`` for ``channel.unary_unary``: (model_name, model_version, input_dict) -> bytes."""
    model_name, model_version, input_dict = request
    return get_codec().encode_predict_request(model_name, input_dict, model_version)


def gpu_response_deserializer(wire: bytes) -> PredictResponseView:
    """``response_deserializer`` for ``channel.unary_unary``: bytes -> lazy response view."""
    return PredictResponseView(wire)


# grpc's default receive limit is 4 MiB and the reference leaves it alone (requests.py:27-30): a response that carries a
# fp32[1024,1024] tensor (4 194 3xx bytes) is refused with RESOURCE_EXHAUSTED.  Pass these to lift both limits.
LARGE_MESSAGE_CHANNEL_OPTIONS = (("grpc.max_send_message_length", -1), ("grpc.max_receive_message_length", -1))


class TensorServingClient:
    def __init__(self, host: str, port: int, credentials=None, channel_options=None) -> None:
        """Same arguments as the reference (requests.py:22-30); ``channel_options`` (default None: grpc's defaults, like the
        reference) is handed to ``grpc.insecure_channel`` / ``grpc.secure_channel``, e.g. ``LARGE_MESSAGE_CHANNEL_OPTIONS``."""
        import grpc

        self._host_address = f"{host}:{port}"
        options = list(channel_options) if channel_options else None
        if credentials:
            self._channel = grpc.secure_channel(self._host_address, credentials, options=options)
        else:
            self._channel = grpc.insecure_channel(self._host_address, options=options)
        self._predict = self._channel.unary_unary(PREDICT_METHOD, request_serializer=gpu_request_serializer,
                                                  response_deserializer=gpu_response_deserializer)

    def predict_request(self, model_name: str, input_dict: Dict[str, np.ndarray], timeout: int = 60,
                        model_version: Optional[int] = None) -> PredictResponseView:
        return self._predict((model_name, model_version, input_dict), timeout)

    def _make_example_request(self, request_pb, model_name, input_dict, model_version):
        request = request_pb()
        request.model_spec.name = model_name
        if model_version is not None:
            request.model_spec.version.value = model_version
        request.input.CopyFrom(examples_from_input_dict(input_dict))
        return request

    def classification_request(self, model_name: str, input_dict: Dict[str, np.ndarray], timeout: int = 60,
                               model_version: Optional[int] = None):
        """Same signature as the reference (requests.py:67-81); returns a ``ClassificationResponse``."""
        from tensorflow_serving.apis.classification_pb2 import ClassificationRequest, ClassificationResponse

        call = self._channel.unary_unary(CLASSIFY_METHOD, request_serializer=ClassificationRequest.SerializeToString,
                                         response_deserializer=ClassificationResponse.FromString)
        return call(self._make_example_request(ClassificationRequest, model_name, input_dict, model_version), timeout)

    def regression_request(self, model_name: str, input_dict: Dict[str, np.ndarray], timeout: int = 60,
                           model_version: Optional[int] = None):
        """Same signature as the reference (requests.py:83-97); returns a ``RegressionResponse``."""
        from tensorflow_serving.apis.regression_pb2 import RegressionRequest, RegressionResponse

        call = self._channel.unary_unary(REGRESS_METHOD, request_serializer=RegressionRequest.SerializeToString,
                                         response_deserializer=RegressionResponse.FromString)
        return call(self._make_example_request(RegressionRequest, model_name, input_dict, model_version), timeout)

    def model_status_request(self, model_name: str, model_version: Optional[int] = None, timeout: Optional[int] = 10):
        """``ModelService/GetModelStatus`` as the reference issues it (requests.py:99-110: the version is set only when truthy)."""
        from tensorflow_serving.apis.get_model_status_pb2 import GetModelStatusRequest, GetModelStatusResponse

        request = GetModelStatusRequest()
        request.model_spec.name = model_name
        if model_version:
            request.model_spec.version.value = model_version
        call = self._channel.unary_unary(MODEL_STATUS_METHOD, request_serializer=GetModelStatusRequest.SerializeToString,
                                         response_deserializer=GetModelStatusResponse.FromString)
        return call(request, timeout)
