# -*- coding: utf-8 -*-
# Schema module written by tools/gen_pb2.py (no protoc in this image).  DO NOT EDIT BY HAND.
# source: tensorflow/core/framework/tensor_shape.proto
"""Message classes for ``tensorflow/core/framework/tensor_shape.proto`` built from a serialised FileDescriptorProto."""
from google.protobuf import descriptor_pool as _descriptor_pool
from google.protobuf import symbol_database as _symbol_database
from google.protobuf.internal import builder as _builder

_sym_db = _symbol_database.Default()

DESCRIPTOR = _descriptor_pool.Default().AddSerializedFile(b'\n,tensorflow/core/framework/tensor_shape.proto\x12\ntensorflow"\x98\x01\n\x10TensorShapeProto\x122\n\x03dim\x18\x02 \x03(\x0b2 .tensorflow.TensorShapeProto.DimR\x03dim\x12!\n\x0cunknown_rank\x18\x03 \x01(\x08R\x0bunknownRank\x1a-\n\x03Dim\x12\x12\n\x04size\x18\x01 \x01(\x03R\x04size\x12\x12\n\x04name\x18\x02 \x01(\tR\x04nameb\x06proto3')

_globals = globals()
_builder.BuildMessageAndEnumDescriptors(DESCRIPTOR, _globals)
_builder.BuildTopDescriptorsAndMessages(DESCRIPTOR, 'tensorflow.core.framework.tensor_shape_pb2', _globals)
