# -*- coding: utf-8 -*-
# Schema module written by tools/gen_pb2.py (no protoc in this image).  DO NOT EDIT BY HAND.
# source: tensorflow_serving/apis/get_model_status.proto
"""Message classes for ``tensorflow_serving/apis/get_model_status.proto`` built from a serialised FileDescriptorProto."""
from google.protobuf import descriptor_pool as _descriptor_pool
from google.protobuf import symbol_database as _symbol_database
from google.protobuf.internal import builder as _builder
from tensorflow_serving.apis import model_pb2 as tensorflow_serving_dot_apis_dot_model_pb2  # noqa: F401
from tensorflow_serving.util import status_pb2 as tensorflow_serving_dot_util_dot_status_pb2  # noqa: F401
_sym_db = _symbol_database.Default()

DESCRIPTOR = _descriptor_pool.Default().AddSerializedFile(b'\n.tensorflow_serving/apis/get_model_status.proto\x12\x12tensorflow.serving\x1a#tensorflow_serving/apis/model.proto\x1a$tensorflow_serving/util/status.proto"U\n\x15GetModelStatusRequest\x12<\n\nmodel_spec\x18\x01 \x01(\x0b2\x1d.tensorflow.serving.ModelSpecR\tmodelSpec"\x80\x02\n\x12ModelVersionStatus\x12\x18\n\x07version\x18\x01 \x01(\x03R\x07version\x12B\n\x05state\x18\x02 \x01(\x0e2,.tensorflow.serving.ModelVersionStatus.StateR\x05state\x127\n\x06status\x18\x03 \x01(\x0b2\x1f.tensorflow.serving.StatusProtoR\x06status"S\n\x05State\x12\x0b\n\x07UNKNOWN\x10\x00\x12\t\n\x05START\x10\n\x12\x0b\n\x07LOADING\x10\x14\x12\r\n\tAVAILABLE\x10\x1e\x12\r\n\tUNLOADING\x10(\x12\x07\n\x03END\x102"t\n\x16GetModelStatusResponse\x12Z\n\x14model_version_status\x18\x01 \x03(\x0b2&.tensorflow.serving.ModelVersionStatusR\x14model_version_statusb\x06proto3')

_globals = globals()
_builder.BuildMessageAndEnumDescriptors(DESCRIPTOR, _globals)
_builder.BuildTopDescriptorsAndMessages(DESCRIPTOR, 'tensorflow_serving.apis.get_model_status_pb2', _globals)
