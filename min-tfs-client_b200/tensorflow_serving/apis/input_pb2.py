# -*- coding: utf-8 -*-
# Schema module written by tools/gen_pb2.py (no protoc in this image).  DO NOT EDIT BY HAND.
# source: tensorflow_serving/apis/input.proto
"""Message classes for ``tensorflow_serving/apis/input.proto`` built from a serialised FileDescriptorProto."""
from google.protobuf import descriptor_pool as _descriptor_pool
from google.protobuf import symbol_database as _symbol_database
from google.protobuf.internal import builder as _builder
from tensorflow.core.example import example_pb2 as tensorflow_dot_core_dot_example_dot_example_pb2  # noqa: F401
_sym_db = _symbol_database.Default()

DESCRIPTOR = _descriptor_pool.Default().AddSerializedFile(b'\n#tensorflow_serving/apis/input.proto\x12\x12tensorflow.serving\x1a%tensorflow/core/example/example.proto">\n\x0bExampleList\x12/\n\x08examples\x18\x01 \x03(\x0b2\x13.tensorflow.ExampleR\x08examples"x\n\x16ExampleListWithContext\x12/\n\x08examples\x18\x01 \x03(\x0b2\x13.tensorflow.ExampleR\x08examples\x12-\n\x07context\x18\x02 \x01(\x0b2\x13.tensorflow.ExampleR\x07context"\xc6\x01\n\x05Input\x12H\n\x0cexample_list\x18\x01 \x01(\x0b2\x1f.tensorflow.serving.ExampleListB\x02(\x01H\x00R\x0bexampleList\x12k\n\x19example_list_with_context\x18\x02 \x01(\x0b2*.tensorflow.serving.ExampleListWithContextB\x02(\x01H\x00R\x16exampleListWithContextB\x06\n\x04kindb\x06proto3')

_globals = globals()
_builder.BuildMessageAndEnumDescriptors(DESCRIPTOR, _globals)
_builder.BuildTopDescriptorsAndMessages(DESCRIPTOR, 'tensorflow_serving.apis.input_pb2', _globals)
