# -*- coding: utf-8 -*-
# Schema module written by tools/gen_pb2.py (no protoc in this image).  DO NOT EDIT BY HAND.
# source: tensorflow_serving/apis/model.proto
"""Message classes for ``tensorflow_serving/apis/model.proto`` built from a serialised FileDescriptorProto."""
from google.protobuf import descriptor_pool as _descriptor_pool
from google.protobuf import symbol_database as _symbol_database
from google.protobuf.internal import builder as _builder
from google.protobuf import wrappers_pb2 as google_dot_protobuf_dot_wrappers_pb2  # noqa: F401
_sym_db = _symbol_database.Default()

DESCRIPTOR = _descriptor_pool.Default().AddSerializedFile(b'\n#tensorflow_serving/apis/model.proto\x12\x12tensorflow.serving\x1a\x1egoogle/protobuf/wrappers.proto"\xb8\x01\n\tModelSpec\x12\x12\n\x04name\x18\x01 \x01(\tR\x04name\x127\n\x07version\x18\x02 \x01(\x0b2\x1b.google.protobuf.Int64ValueH\x00R\x07version\x12%\n\rversion_label\x18\x04 \x01(\tH\x00R\x0cversionLabel\x12%\n\x0esignature_name\x18\x03 \x01(\tR\rsignatureNameB\x10\n\x0eversion_choiceb\x06proto3')

_globals = globals()
_builder.BuildMessageAndEnumDescriptors(DESCRIPTOR, _globals)
_builder.BuildTopDescriptorsAndMessages(DESCRIPTOR, 'tensorflow_serving.apis.model_pb2', _globals)
