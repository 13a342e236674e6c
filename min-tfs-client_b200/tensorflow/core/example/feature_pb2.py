# -*- coding: utf-8 -*-
# Schema module written by tools/gen_pb2.py (no protoc in this image).  DO NOT EDIT BY HAND.
# source: tensorflow/core/example/feature.proto
"""Message classes for ``tensorflow/core/example/feature.proto`` built from a serialised FileDescriptorProto."""
from google.protobuf import descriptor_pool as _descriptor_pool
from google.protobuf import symbol_database as _symbol_database
from google.protobuf.internal import builder as _builder

_sym_db = _symbol_database.Default()

DESCRIPTOR = _descriptor_pool.Default().AddSerializedFile(b'\n%tensorflow/core/example/feature.proto\x12\ntensorflow"!\n\tBytesList\x12\x14\n\x05value\x18\x01 \x03(\x0cR\x05value"%\n\tFloatList\x12\x18\n\x05value\x18\x01 \x03(\x02B\x02\x10\x01R\x05value"%\n\tInt64List\x12\x18\n\x05value\x18\x01 \x03(\x03B\x02\x10\x01R\x05value"\xb9\x01\n\x07Feature\x126\n\nbytes_list\x18\x01 \x01(\x0b2\x15.tensorflow.BytesListH\x00R\tbytesList\x126\n\nfloat_list\x18\x02 \x01(\x0b2\x15.tensorflow.FloatListH\x00R\tfloatList\x126\n\nint64_list\x18\x03 \x01(\x0b2\x15.tensorflow.Int64ListH\x00R\tint64ListB\x06\n\x04kind"\x98\x01\n\x08Features\x12;\n\x07feature\x18\x01 \x03(\x0b2!.tensorflow.Features.FeatureEntryR\x07feature\x1aO\n\x0cFeatureEntry\x12\x10\n\x03key\x18\x01 \x01(\tR\x03key\x12)\n\x05value\x18\x02 \x01(\x0b2\x13.tensorflow.FeatureR\x05value:\x028\x01"<\n\x0bFeatureList\x12-\n\x07feature\x18\x01 \x03(\x0b2\x13.tensorflow.FeatureR\x07feature"\xb5\x01\n\x0cFeatureLists\x12L\n\x0cfeature_list\x18\x01 \x03(\x0b2).tensorflow.FeatureLists.FeatureListEntryR\x0bfeatureList\x1aW\n\x10FeatureListEntry\x12\x10\n\x03key\x18\x01 \x01(\tR\x03key\x12-\n\x05value\x18\x02 \x01(\x0b2\x17.tensorflow.FeatureListR\x05value:\x028\x01b\x06proto3')

_globals = globals()
_builder.BuildMessageAndEnumDescriptors(DESCRIPTOR, _globals)
_builder.BuildTopDescriptorsAndMessages(DESCRIPTOR, 'tensorflow.core.example.feature_pb2', _globals)
