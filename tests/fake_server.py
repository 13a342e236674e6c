"""In-process stand-in for tensorflow_model_server, for the drop-in's integration test.

Serves the identity model of the reference's fixture (tests/integration/fixtures/generate_tensorflow_model.py:
`x_input` -> `x_output` through tf.identity) on /tensorflow.serving.PredictionService/Predict, using the
protobuf runtime on the server side (the server is not the product): parse PredictRequest, copy each
`<name>_input` TensorProto to `<name>_output`, answer with model_spec{name, version, signature_name}.
It also records the request bytes it received, so the test can check what went over the wire.
"""
from concurrent import futures

import grpc

from tensorflow_serving.apis import predict_pb2

PREDICT = "/tensorflow.serving.PredictionService/Predict"


class IdentityServer:
    def __init__(self):
        self.received = []
        self.server = grpc.server(futures.ThreadPoolExecutor(max_workers=2),
                                  options=[("grpc.max_receive_message_length", 64 << 20), ("grpc.max_send_message_length", 64 << 20)])
        handler = grpc.method_handlers_generic_handler("tensorflow.serving.PredictionService", {
            "Predict": grpc.unary_unary_rpc_method_handler(self._predict, request_deserializer=lambda b: b, response_serializer=lambda b: b)})
        self.server.add_generic_rpc_handlers((handler,))
        self.port = self.server.add_insecure_port("127.0.0.1:0")
        self.server.start()

    def _predict(self, request_bytes, context):
        self.received.append(request_bytes)
        req = predict_pb2.PredictRequest.FromString(request_bytes)
        resp = predict_pb2.PredictResponse()
        for key, proto in req.inputs.items():
            out = key[: -len("_input")] + "_output" if key.endswith("_input") else key
            resp.outputs[out].CopyFrom(proto)
        resp.model_spec.name = req.model_spec.name
        resp.model_spec.version.value = req.model_spec.version.value if req.model_spec.HasField("version") else 1
        resp.model_spec.signature_name = "serving_default"
        return resp.SerializeToString()

    def stop(self):
        self.server.stop(0)
