# -*- coding: utf-8 -*-
# Schema module written by tools/gen_pb2.py (no protoc in this image).  DO NOT EDIT BY HAND.
# source: tensorflow_serving/apis/predict.proto
"""Message classes for ``tensorflow_serving/apis/predict.proto`` built from a serialised FileDescriptorProto."""
from google.protobuf import descriptor_pool as _descriptor_pool
from google.protobuf import symbol_database as _symbol_database
from google.protobuf.internal import builder as _builder
from tensorflow.core.framework import tensor_pb2 as tensorflow_dot_core_dot_framework_dot_tensor_pb2  # noqa: F401
from tensorflow_serving.apis import model_pb2 as tensorflow_serving_dot_apis_dot_model_pb2  # noqa: F401
_sym_db = _symbol_database.Default()

DESCRIPTOR = _descriptor_pool.Default().AddSerializedFile(b'\n%tensorflow_serving/apis/predict.proto\x12\x12tensorflow.serving\x1a&tensorflow/core/framework/tensor.proto\x1a#tensorflow_serving/apis/model.proto"\x8f\x02\n\x0ePredictRequest\x12<\n\nmodel_spec\x18\x01 \x01(\x0b2\x1d.tensorflow.serving.ModelSpecR\tmodelSpec\x12F\n\x06inputs\x18\x02 \x03(\x0b2..tensorflow.serving.PredictRequest.InputsEntryR\x06inputs\x12#\n\routput_filter\x18\x03 \x03(\tR\x0coutputFilter\x1aR\n\x0bInputsEntry\x12\x10\n\x03key\x18\x01 \x01(\tR\x03key\x12-\n\x05value\x18\x02 \x01(\x0b2\x17.tensorflow.TensorProtoR\x05value:\x028\x01"\xf0\x01\n\x0fPredictResponse\x12<\n\nmodel_spec\x18\x02 \x01(\x0b2\x1d.tensorflow.serving.ModelSpecR\tmodelSpec\x12J\n\x07outputs\x18\x01 \x03(\x0b20.tensorflow.serving.PredictResponse.OutputsEntryR\x07outputs\x1aS\n\x0cOutputsEntry\x12\x10\n\x03key\x18\x01 \x01(\tR\x03key\x12-\n\x05value\x18\x02 \x01(\x0b2\x17.tensorflow.TensorProtoR\x05value:\x028\x01b\x06proto3')

_globals = globals()
_builder.BuildMessageAndEnumDescriptors(DESCRIPTOR, _globals)
_builder.BuildTopDescriptorsAndMessages(DESCRIPTOR, 'tensorflow_serving.apis.predict_pb2', _globals)
