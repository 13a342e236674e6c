# -*- coding: utf-8 -*-
# Schema module written by tools/gen_pb2.py (no protoc in this image).  DO NOT EDIT BY HAND.
# source: tensorflow/core/framework/tensor.proto
"""Message classes for ``tensorflow/core/framework/tensor.proto`` built from a serialised FileDescriptorProto."""
from google.protobuf import descriptor_pool as _descriptor_pool
from google.protobuf import symbol_database as _symbol_database
from google.protobuf.internal import builder as _builder
from tensorflow.core.framework import resource_handle_pb2 as tensorflow_dot_core_dot_framework_dot_resource_handle_pb2  # noqa: F401
from tensorflow.core.framework import tensor_shape_pb2 as tensorflow_dot_core_dot_framework_dot_tensor_shape_pb2  # noqa: F401
from tensorflow.core.framework import types_pb2 as tensorflow_dot_core_dot_framework_dot_types_pb2  # noqa: F401
_sym_db = _symbol_database.Default()

DESCRIPTOR = _descriptor_pool.Default().AddSerializedFile(b'\n&tensorflow/core/framework/tensor.proto\x12\ntensorflow\x1a/tensorflow/core/framework/resource_handle.proto\x1a,tensorflow/core/framework/tensor_shape.proto\x1a%tensorflow/core/framework/types.proto"\xd1\x05\n\x0bTensorProto\x12*\n\x05dtype\x18\x01 \x01(\x0e2\x14.tensorflow.DataTypeR\x05dtype\x12?\n\x0ctensor_shape\x18\x02 \x01(\x0b2\x1c.tensorflow.TensorShapeProtoR\x0btensorShape\x12%\n\x0eversion_number\x18\x03 \x01(\x05R\rversionNumber\x12%\n\x0etensor_content\x18\x04 \x01(\x0cR\rtensorContent\x12\x1d\n\x08half_val\x18\r \x03(\x05B\x02\x10\x01R\x07halfVal\x12\x1f\n\tfloat_val\x18\x05 \x03(\x02B\x02\x10\x01R\x08floatVal\x12!\n\ndouble_val\x18\x06 \x03(\x01B\x02\x10\x01R\tdoubleVal\x12\x1b\n\x07int_val\x18\x07 \x03(\x05B\x02\x10\x01R\x06intVal\x12\x1d\n\nstring_val\x18\x08 \x03(\x0cR\tstringVal\x12%\n\x0cscomplex_val\x18\t \x03(\x02B\x02\x10\x01R\x0bscomplexVal\x12\x1f\n\tint64_val\x18\n \x03(\x03B\x02\x10\x01R\x08int64Val\x12\x1d\n\x08bool_val\x18\x0b \x03(\x08B\x02\x10\x01R\x07boolVal\x12%\n\x0cdcomplex_val\x18\x0c \x03(\x01B\x02\x10\x01R\x0bdcomplexVal\x12O\n\x13resource_handle_val\x18\x0e \x03(\x0b2\x1f.tensorflow.ResourceHandleProtoR\x11resourceHandleVal\x12C\n\x0bvariant_val\x18\x0f \x03(\x0b2".tensorflow.VariantTensorDataProtoR\nvariantVal\x12!\n\nuint32_val\x18\x10 \x03(\rB\x02\x10\x01R\tuint32Val\x12!\n\nuint64_val\x18\x11 \x03(\x04B\x02\x10\x01R\tuint64Val"\x84\x01\n\x16VariantTensorDataProto\x12\x1b\n\ttype_name\x18\x01 \x01(\tR\x08typeName\x12\x1a\n\x08metadata\x18\x02 \x01(\x0cR\x08metadata\x121\n\x07tensors\x18\x03 \x03(\x0b2\x17.tensorflow.TensorProtoR\x07tensorsb\x06proto3')

_globals = globals()
_builder.BuildMessageAndEnumDescriptors(DESCRIPTOR, _globals)
_builder.BuildTopDescriptorsAndMessages(DESCRIPTOR, 'tensorflow.core.framework.tensor_pb2', _globals)
