# -*- coding: utf-8 -*-
# Schema module written by tools/gen_pb2.py (no protoc in this image).  DO NOT EDIT BY HAND.
# source: tensorflow/core/example/example.proto
"""Message classes for ``tensorflow/core/example/example.proto`` built from a serialised FileDescriptorProto."""
from google.protobuf import descriptor_pool as _descriptor_pool
from google.protobuf import symbol_database as _symbol_database
from google.protobuf.internal import builder as _builder
from tensorflow.core.example import feature_pb2 as tensorflow_dot_core_dot_example_dot_feature_pb2  # noqa: F401
_sym_db = _symbol_database.Default()

DESCRIPTOR = _descriptor_pool.Default().AddSerializedFile(b'\n%tensorflow/core/example/example.proto\x12\ntensorflow\x1a%tensorflow/core/example/feature.proto";\n\x07Example\x120\n\x08features\x18\x01 \x01(\x0b2\x14.tensorflow.FeaturesR\x08features"\x80\x01\n\x0fSequenceExample\x12.\n\x07context\x18\x01 \x01(\x0b2\x14.tensorflow.FeaturesR\x07context\x12=\n\rfeature_lists\x18\x02 \x01(\x0b2\x18.tensorflow.FeatureListsR\x0cfeatureListsb\x06proto3')

_globals = globals()
_builder.BuildMessageAndEnumDescriptors(DESCRIPTOR, _globals)
_builder.BuildTopDescriptorsAndMessages(DESCRIPTOR, 'tensorflow.core.example.example_pb2', _globals)
