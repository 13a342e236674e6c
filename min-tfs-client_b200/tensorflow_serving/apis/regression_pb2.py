# -*- coding: utf-8 -*-
# Schema module written by tools/gen_pb2.py (no protoc in this image).  DO NOT EDIT BY HAND.
# source: tensorflow_serving/apis/regression.proto
"""Message classes for ``tensorflow_serving/apis/regression.proto`` built from a serialised FileDescriptorProto."""
from google.protobuf import descriptor_pool as _descriptor_pool
from google.protobuf import symbol_database as _symbol_database
from google.protobuf.internal import builder as _builder
from tensorflow_serving.apis import input_pb2 as tensorflow_serving_dot_apis_dot_input_pb2  # noqa: F401
from tensorflow_serving.apis import model_pb2 as tensorflow_serving_dot_apis_dot_model_pb2  # noqa: F401
_sym_db = _symbol_database.Default()

DESCRIPTOR = _descriptor_pool.Default().AddSerializedFile(b'\n(tensorflow_serving/apis/regression.proto\x12\x12tensorflow.serving\x1a#tensorflow_serving/apis/input.proto\x1a#tensorflow_serving/apis/model.proto""\n\nRegression\x12\x14\n\x05value\x18\x01 \x01(\x02R\x05value"T\n\x10RegressionResult\x12@\n\x0bregressions\x18\x01 \x03(\x0b2\x1e.tensorflow.serving.RegressionR\x0bregressions"\x82\x01\n\x11RegressionRequest\x12<\n\nmodel_spec\x18\x01 \x01(\x0b2\x1d.tensorflow.serving.ModelSpecR\tmodelSpec\x12/\n\x05input\x18\x02 \x01(\x0b2\x19.tensorflow.serving.InputR\x05input"\x90\x01\n\x12RegressionResponse\x12<\n\nmodel_spec\x18\x02 \x01(\x0b2\x1d.tensorflow.serving.ModelSpecR\tmodelSpec\x12<\n\x06result\x18\x01 \x01(\x0b2$.tensorflow.serving.RegressionResultR\x06resultb\x06proto3')

_globals = globals()
_builder.BuildMessageAndEnumDescriptors(DESCRIPTOR, _globals)
_builder.BuildTopDescriptorsAndMessages(DESCRIPTOR, 'tensorflow_serving.apis.regression_pb2', _globals)
