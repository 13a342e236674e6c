# -*- coding: utf-8 -*-
# Schema module written by tools/gen_pb2.py (no protoc in this image).  DO NOT EDIT BY HAND.
# source: tensorflow_serving/util/status.proto
"""Message classes for ``tensorflow_serving/util/status.proto`` built from a serialised FileDescriptorProto."""
from google.protobuf import descriptor_pool as _descriptor_pool
from google.protobuf import symbol_database as _symbol_database
from google.protobuf.internal import builder as _builder
from tensorflow.core.lib.core import error_codes_pb2 as tensorflow_dot_core_dot_lib_dot_core_dot_error_codes_pb2  # noqa: F401
_sym_db = _symbol_database.Default()

DESCRIPTOR = _descriptor_pool.Default().AddSerializedFile(b'\n$tensorflow_serving/util/status.proto\x12\x12tensorflow.serving\x1a*tensorflow/core/lib/core/error_codes.proto"k\n\x0bStatusProto\x126\n\nerror_code\x18\x01 \x01(\x0e2\x16.tensorflow.error.CodeR\nerror_code\x12$\n\rerror_message\x18\x02 \x01(\tR\rerror_messageb\x06proto3')

_globals = globals()
_builder.BuildMessageAndEnumDescriptors(DESCRIPTOR, _globals)
_builder.BuildTopDescriptorsAndMessages(DESCRIPTOR, 'tensorflow_serving.util.status_pb2', _globals)
