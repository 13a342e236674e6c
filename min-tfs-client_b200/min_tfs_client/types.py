"""``DataType``: resolve a numpy type, a ``"DT_*"`` string or a DataType enum int.

Interface and error behaviour of the reference's ``min_tfs_client/types.py:13-42``: attributes
``numpy_dtype, is_numeric, tf_dtype, enum, proto_field_name``; ``ValueError`` for a type outside the
table or an argument that is not a type/str/int; ``KeyError`` for an unmapped string or enum.
"""
from typing import Union

import numpy as np

from .constants import (
    ENUM_TO_TF_MAPPING,
    NP_TO_ENUM_MAPPING,
    NP_TO_TF_MAPPING,
    NUMERICAL_TYPES,
    TF_TO_NP_MAPPING,
)


class DataType:
    VALID_TYPES = NUMERICAL_TYPES.union({np.str_, np.bool_})

    def __init__(self, dtype: Union[type, str, int]):
        resolved = self._get_numpy_dtype(dtype)
        self._validate_dtype(resolved)
        row = NP_TO_TF_MAPPING[resolved]
        self.numpy_dtype = resolved
        self.is_numeric = resolved in NUMERICAL_TYPES
        self.tf_dtype = row.TFDType
        self.enum = NP_TO_ENUM_MAPPING[resolved]
        self.proto_field_name = row.TensorProtoField

    def _validate_dtype(self, numpy_dtype: type) -> None:
        if numpy_dtype in self.VALID_TYPES:
            return
        allowed = ", ".join(t.__name__ for t in self.VALID_TYPES)
        raise ValueError(f"Dtype {numpy_dtype.__name__} is not valid. Allowable values: {allowed}")

    def _get_numpy_dtype(self, dtype: Union[type, str, int]) -> type:
        if isinstance(dtype, type):
            return dtype
        if isinstance(dtype, str):
            return np.dtype(TF_TO_NP_MAPPING[dtype]).type
        if isinstance(dtype, int):
            return np.dtype(TF_TO_NP_MAPPING[ENUM_TO_TF_MAPPING[dtype]]).type
        raise ValueError(f"Expected dtype of types: type, str, or int, got {type(dtype)}")
