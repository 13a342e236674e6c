# -*- coding: utf-8 -*-
# Schema module written by tools/gen_pb2.py (no protoc in this image).  DO NOT EDIT BY HAND.
# source: tensorflow/core/framework/resource_handle.proto
"""Message classes for ``tensorflow/core/framework/resource_handle.proto`` built from a serialised FileDescriptorProto."""
from google.protobuf import descriptor_pool as _descriptor_pool
from google.protobuf import symbol_database as _symbol_database
from google.protobuf.internal import builder as _builder
from tensorflow.core.framework import tensor_shape_pb2 as tensorflow_dot_core_dot_framework_dot_tensor_shape_pb2  # noqa: F401
from tensorflow.core.framework import types_pb2 as tensorflow_dot_core_dot_framework_dot_types_pb2  # noqa: F401
_sym_db = _symbol_database.Default()

DESCRIPTOR = _descriptor_pool.Default().AddSerializedFile(b'\n/tensorflow/core/framework/resource_handle.proto\x12\ntensorflow\x1a,tensorflow/core/framework/tensor_shape.proto\x1a%tensorflow/core/framework/types.proto"\xf0\x02\n\x13ResourceHandleProto\x12\x16\n\x06device\x18\x01 \x01(\tR\x06device\x12\x1c\n\tcontainer\x18\x02 \x01(\tR\tcontainer\x12\x12\n\x04name\x18\x03 \x01(\tR\x04name\x12\x1b\n\thash_code\x18\x04 \x01(\x04R\x08hashCode\x12&\n\x0fmaybe_type_name\x18\x05 \x01(\tR\rmaybeTypeName\x12Y\n\x11dtypes_and_shapes\x18\x06 \x03(\x0b2-.tensorflow.ResourceHandleProto.DtypeAndShapeR\x0fdtypesAndShapes\x1ao\n\rDtypeAndShape\x12*\n\x05dtype\x18\x01 \x01(\x0e2\x14.tensorflow.DataTypeR\x05dtype\x122\n\x05shape\x18\x02 \x01(\x0b2\x1c.tensorflow.TensorShapeProtoR\x05shapeb\x06proto3')

_globals = globals()
_builder.BuildMessageAndEnumDescriptors(DESCRIPTOR, _globals)
_builder.BuildTopDescriptorsAndMessages(DESCRIPTOR, 'tensorflow.core.framework.resource_handle_pb2', _globals)
