"""Drop-in for the Predict part of the reference's ``min_tfs_client/requests.py``.

``TensorServingClient(host, port, credentials=None).predict_request(model_name, input_dict,
timeout=60, model_version=None)`` keeps the reference's signature (requests.py:22-65).  The request
is packed by the encode kernels and sent as raw bytes through the same gRPC method the generated stub
binds (``/tensorflow.serving.PredictionService/Predict``, prediction_service_pb2_grpc.py:50-54); the
response bytes are handed to the parse/unpack kernels.  The Classify / Regress / GetModelStatus
helpers of the reference are outside the Predict hot path and raise.
"""
from typing import Dict, Optional

import numpy as np

from .codec import get_codec
from .tensors import WireTensor

PREDICT_METHOD = "/tensorflow.serving.PredictionService/Predict"


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

    def _out_of_scope(self, what):
        raise NotImplementedError(f"{what} is outside the Predict hot path this package rebuilds (see DESIGN.md)")

    def classification_request(self, *a, **kw):
        self._out_of_scope("classification_request")

    def regression_request(self, *a, **kw):
        self._out_of_scope("regression_request")

    def model_status_request(self, *a, **kw):
        self._out_of_scope("model_status_request")
