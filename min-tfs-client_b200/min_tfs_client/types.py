"""``DataType``: one row of the dtype table, looked up by numpy type, ``"DT_*"`` name or enum value.

Public behaviour of the reference's ``min_tfs_client/types.py:13-42``: the constructor accepts a numpy
scalar type, a TensorFlow dtype name or a ``DataType`` enum integer and exposes ``numpy_dtype``,
``is_numeric``, ``tf_dtype``, ``enum`` and ``proto_field_name``.  Failure modes are the reference's too:
``ValueError`` for a type outside the table or an argument of another kind, ``KeyError`` for a name or
enum value without a row.
"""
import numpy as np

from . import constants as _c


def _row_for(key):
    """numpy scalar type for any of the three accepted spellings."""
    if isinstance(key, type):
        return key
    if isinstance(key, str):
        return np.dtype(_c.TF_TO_NP_MAPPING[key]).type
    if isinstance(key, int):
        return np.dtype(_c.TF_TO_NP_MAPPING[_c.ENUM_TO_TF_MAPPING[key]]).type
    raise ValueError(f"Expected dtype of types: type, str, or int, got {type(key)}")


class DataType:
    VALID_TYPES = frozenset(_c.NUMERICAL_TYPES) | {np.str_, np.bool_}

    __slots__ = ("numpy_dtype", "is_numeric", "tf_dtype", "enum", "proto_field_name")

    def __init__(self, dtype):
        np_type = _row_for(dtype)
        if np_type not in self.VALID_TYPES:
            names = ", ".join(t.__name__ for t in self.VALID_TYPES)
            raise ValueError(f"Dtype {np_type.__name__} is not valid. Allowable values: {names}")
        tf_name, field = _c.NP_TO_TF_MAPPING[np_type]
        self.numpy_dtype = np_type
        self.is_numeric = np_type in _c.NUMERICAL_TYPES
        self.tf_dtype = tf_name
        self.enum = _c.NP_TO_ENUM_MAPPING[np_type]
        self.proto_field_name = field

    def __repr__(self):
        return f"DataType({self.tf_dtype})"
