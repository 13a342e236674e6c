"""In-process stand-in for tensorflow_model_server, for the drop-in's integration test.

Serves the identity model of the reference's fixture (tests/integration/fixtures/generate_tensorflow_model.py:
`x_input` -> `x_output` through tf.identity) on /tensorflow.serving.PredictionService/Predict, using the
protobuf runtime on the server side (the server is not the product): parse PredictRequest, copy each
`<name>_input` TensorProto to `<name>_output`, answer with model_spec{name, version, signature_name}.
It also records the request bytes it received, so the test can check what went over the wire.
"""
from concurrent import futures

import grpc

from tensorflow_serving.apis import classification_pb2, get_model_status_pb2, predict_pb2, regression_pb2

PREDICT = "/tensorflow.serving.PredictionService/Predict"


class IdentityServer:
    def __init__(self):
        self.received = []
        self.server = grpc.server(futures.ThreadPoolExecutor(max_workers=2),
                                  options=[("grpc.max_receive_message_length", 64 << 20), ("grpc.max_send_message_length", 64 << 20)])
        raw = dict(request_deserializer=lambda b: b, response_serializer=lambda b: b)
        handler = grpc.method_handlers_generic_handler("tensorflow.serving.PredictionService", {
            "Predict": grpc.unary_unary_rpc_method_handler(self._predict, **raw),
            "Classify": grpc.unary_unary_rpc_method_handler(self._classify, **raw),
            "Regress": grpc.unary_unary_rpc_method_handler(self._regress, **raw)})
        status = grpc.method_handlers_generic_handler("tensorflow.serving.ModelService", {
            "GetModelStatus": grpc.unary_unary_rpc_method_handler(self._status, **raw)})
        self.server.add_generic_rpc_handlers((handler, status))
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

    # the other RPCs of the reference's client: a toy model - an example's score is the sum of its float features
    @staticmethod
    def _score(example):
        return float(sum(sum(f.float_list.value) for f in example.features.feature.values()))

    def _classify(self, request_bytes, context):
        self.received.append(request_bytes)
        req = classification_pb2.ClassificationRequest.FromString(request_bytes)
        resp = classification_pb2.ClassificationResponse()
        resp.model_spec.CopyFrom(req.model_spec)
        resp.result.SetInParent()
        for ex in req.input.example_list.examples:
            cls = resp.result.classifications.add()
            s = self._score(ex)
            cls.classes.add(label="positive", score=s)
            cls.classes.add(label="negative", score=-s)
        return resp.SerializeToString()

    def _regress(self, request_bytes, context):
        self.received.append(request_bytes)
        req = regression_pb2.RegressionRequest.FromString(request_bytes)
        resp = regression_pb2.RegressionResponse()
        resp.model_spec.CopyFrom(req.model_spec)
        resp.result.SetInParent()
        for ex in req.input.example_list.examples:
            resp.result.regressions.add(value=self._score(ex))
        return resp.SerializeToString()

    def _status(self, request_bytes, context):
        self.received.append(request_bytes)
        req = get_model_status_pb2.GetModelStatusRequest.FromString(request_bytes)
        resp = get_model_status_pb2.GetModelStatusResponse()
        v = resp.model_version_status.add()
        v.version = req.model_spec.version.value if req.model_spec.HasField("version") else 1
        v.state = get_model_status_pb2.ModelVersionStatus.AVAILABLE
        v.status.SetInParent()
        return resp.SerializeToString()

    def stop(self):
        self.server.stop(0)
