// wire.h - protobuf wire-format primitives and the dtype table, usable from host C++ and from
// device code (everything is constexpr / HD-inline; no allocation, no libc beyond <stdint.h>).
//
// Restates, for this path only:
//   * the proto3 wire grammar the reference gets from the protobuf runtime (third-party; reference
//     call sites tensors.py:30-33, requests.py:48, prediction_service_pb2_grpc.py:52-53);
//   * the numpy <-> DT_* <-> TensorProto-field table of reference constants.py:13-29.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#else
#define B2_HD inline
#endif

namespace b200tfs {

// --- wire types ------------------------------------------------------------------------------
enum : uint32_t { WT_VARINT = 0, WT_I64 = 1, WT_LEN = 2, WT_SGROUP = 3, WT_EGROUP = 4, WT_I32 = 5 };

// --- TensorProto field numbers (tensor.proto:14-84) -----------------------------------------
enum : uint32_t {
  F_DTYPE = 1, F_SHAPE = 2, F_VERSION = 3, F_CONTENT = 4, F_FLOAT = 5, F_DOUBLE = 6, F_INT = 7,
  F_STRING = 8, F_SCOMPLEX = 9, F_INT64 = 10, F_BOOL = 11, F_DCOMPLEX = 12, F_HALF = 13,
  F_RESOURCE = 14, F_VARIANT = 15, F_UINT32 = 16, F_UINT64 = 17
};

// --- DataType enum values on this path (types.proto:12-68) -----------------------------------
enum : int32_t {
  DT_INVALID = 0, DT_FLOAT = 1, DT_DOUBLE = 2, DT_INT32 = 3, DT_UINT8 = 4, DT_INT16 = 5, DT_INT8 = 6,
  DT_STRING = 7, DT_COMPLEX64 = 8, DT_INT64 = 9, DT_BOOL = 10, DT_BFLOAT16 = 14, DT_UINT16 = 17,
  DT_COMPLEX128 = 18, DT_HALF = 19, DT_UINT32 = 22, DT_UINT64 = 23
};

// How a dtype's values travel inside their TensorProto field.
enum ValueKind : uint32_t {
  VK_NONE = 0,
  VK_FIXED = 1,   // packed fixed32/fixed64 == raw little-endian memory (float_val, double_val, s/dcomplex_val)
  VK_VARINT = 2,  // packed varints (int_val, int64_val, uint32_val, uint64_val, half_val)
  VK_BOOL = 3,    // packed varints that are always one byte 0/1 on encode (bool_val)
  VK_STRING = 4   // unpacked length-delimited elements (string_val); host-side path
};

struct DtypeInfo {
  uint32_t field;     // TensorProto field number, 0 = unmapped
  uint32_t elem_size; // bytes per element in memory
  uint32_t kind;      // ValueKind
  uint32_t is_signed; // sign-extend to 64 bits before varint (int_val carries int32, Q5)
};

// constants.py:13-29 (15 rows) + DT_BFLOAT16 by TF's convention (tensor_util.py:60-68).
B2_HD DtypeInfo dtype_info(int32_t dt) {
  switch (dt) {
    case DT_FLOAT:      return {F_FLOAT, 4, VK_FIXED, 0};
    case DT_DOUBLE:     return {F_DOUBLE, 8, VK_FIXED, 0};
    case DT_INT32:      return {F_INT, 4, VK_VARINT, 1};
    case DT_UINT8:      return {F_INT, 1, VK_VARINT, 0};
    case DT_INT16:      return {F_INT, 2, VK_VARINT, 1};
    case DT_INT8:       return {F_INT, 1, VK_VARINT, 1};
    case DT_STRING:     return {F_STRING, 0, VK_STRING, 0};
    case DT_COMPLEX64:  return {F_SCOMPLEX, 8, VK_FIXED, 0};
    case DT_INT64:      return {F_INT64, 8, VK_VARINT, 1};
    case DT_BOOL:       return {F_BOOL, 1, VK_BOOL, 0};
    case DT_BFLOAT16:   return {F_HALF, 2, VK_VARINT, 0};
    case DT_UINT16:     return {F_INT, 2, VK_VARINT, 0};
    case DT_COMPLEX128: return {F_DCOMPLEX, 16, VK_FIXED, 0};
    case DT_HALF:       return {F_HALF, 2, VK_VARINT, 0};
    case DT_UINT32:     return {F_UINT32, 4, VK_VARINT, 0};
    case DT_UINT64:     return {F_UINT64, 8, VK_VARINT, 0};
    default:            return {0, 0, VK_NONE, 0};
  }
}

// wire element width of a packed fixed field (float_val / scomplex_val carry fixed32, double_val /
// dcomplex_val fixed64): packed length must be a multiple of this or the parser rejects (D4).
B2_HD uint32_t fixed_wire_width(uint32_t field) {
  return (field == F_FLOAT || field == F_SCOMPLEX) ? 4u : (field == F_DOUBLE || field == F_DCOMPLEX) ? 8u : 0u;
}
// scalar wire type an UNPACKED element of a repeated numeric field must have
B2_HD uint32_t scalar_wire_type(uint32_t field) {
  switch (field) {
    case F_FLOAT: case F_SCOMPLEX: return WT_I32;
    case F_DOUBLE: case F_DCOMPLEX: return WT_I64;
    case F_INT: case F_INT64: case F_BOOL: case F_HALF: case F_UINT32: case F_UINT64: return WT_VARINT;
    default: return 0xFFu;
  }
}

// --- varints -----------------------------------------------------------------------------------
B2_HD uint32_t varint_len(uint64_t v) {
  // 1 + floor(log128(v)); v==0 -> 1
  uint32_t n = 1;
  while (v >= 0x80) { v >>= 7; ++n; }
  return n;
}
// branch-free form for device loops: bits = 64 - clz(v|1); len = (bits + 6) / 7
#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t varint_len_fast(uint64_t v) {
  return (uint32_t)(64 - __clzll((long long)(v | 1)) + 6) / 7u;
}
#endif
B2_HD uint32_t put_varint(uint8_t* p, uint64_t v) {
  uint32_t n = 0;
  while (v >= 0x80) { p[n++] = (uint8_t)(v | 0x80); v >>= 7; }
  p[n++] = (uint8_t)v;
  return n;
}
B2_HD uint32_t tag_of(uint32_t field, uint32_t wt) { return (field << 3) | wt; }

// float32 signalling-NaN quieting (the reference's float32 -> Python double -> float32
// trip at tensors.py:22 / :46 sets the quiet bit and keeps sign + payload).
B2_HD uint32_t quiet_f32(uint32_t w) { return ((w & 0x7FFFFFFFu) > 0x7F800000u) ? (w | 0x00400000u) : w; }

}  // namespace b200tfs
