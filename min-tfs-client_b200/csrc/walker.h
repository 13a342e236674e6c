// walker.h - the PredictResponse / TensorProto tag walk that the parse kernel runs (one lane per
// record).  It tabulates where every output's values lie in the wire so the unpack kernels can move
// them; it moves no payload bytes itself.
//
// Behaviour follows what the reference observes through PredictResponse.FromString
// (prediction_service_pb2_grpc.py:53 -> protobuf runtime) + extract_shape / tensor_proto_to_ndarray
// (tensors.py:38-46), case by case as pinned in tests/golden/decode.json:
//   fields in any order; scalar dtype: last wins; tensor_shape repeated: dims concatenate; packed
//   fields split over several occurrences and unpacked elements: concatenate in wire order; unknown
//   fields (varint / fixed / length-delimited / groups) skipped; duplicate map key: last entry wins;
//   truncated input, tag 0, bad UTF-8 in a string field, packed fixed32 length % 4 != 0: parse error.
//
// Written as host+device inline code so the same source is unit-tested on the CPU against the golden
// vectors (tests/native/) and runs unchanged inside parse_kernel.
#pragma once
#include "../../include/b200tfs.h"
#include "wire.h"

namespace b200tfs {

// Two 128-byte lines of the wire cached next to the walking lane (shared memory on the device).  A
// single lane chasing bytes through HBM pays a full DRAM round trip per miss, so it fetches whole
// lines with eight independent 16-byte loads; the canonical response (header ... payload ... model_spec)
// then costs two misses, both issued up front by win_prefetch().  On the host the window is unused.
struct Win {
  uint64_t line[2];  // absolute address of each cached line (aligned to 128), ~0 = empty
  uint8_t* buf;      // 256 bytes
  uint32_t victim;
};

struct Cursor {
  const uint8_t* w;  // arena base
  uint64_t p;        // current offset
  uint64_t end;      // limit of the enclosing message
  int err;           // sticky B200TFS_E_PARSE
  Win* win;          // device: line cache; host: nullptr
};

#if defined(__CUDACC__)
__device__ __forceinline__ void win_fill(Win* W, uint32_t k, uint64_t line) {
  const uint4* g = reinterpret_cast<const uint4*>(line);
  uint4 t0 = g[0], t1 = g[1], t2 = g[2], t3 = g[3], t4 = g[4], t5 = g[5], t6 = g[6], t7 = g[7];
  uint4* s = reinterpret_cast<uint4*>(W->buf + 128 * k);
  s[0] = t0; s[1] = t1; s[2] = t2; s[3] = t3; s[4] = t4; s[5] = t5; s[6] = t6; s[7] = t7;
  W->line[k] = line;
}
// fetch the first and the last line of a record together (both loads in flight at once)
__device__ __forceinline__ void win_prefetch(Win* W, const uint8_t* w, uint64_t off, uint64_t len) {
  W->line[0] = W->line[1] = ~0ull; W->victim = 0;
  if (!len) return;
  const uint64_t a = (uint64_t)(uintptr_t)(w + off) & ~127ull, b = (uint64_t)(uintptr_t)(w + off + len - 1) & ~127ull;
  const uint4* ga = reinterpret_cast<const uint4*>(a);
  const uint4* gb = reinterpret_cast<const uint4*>(b);
  uint4 x0 = ga[0], x1 = ga[1], x2 = ga[2], x3 = ga[3], x4 = ga[4], x5 = ga[5], x6 = ga[6], x7 = ga[7];
  uint4 y0 = gb[0], y1 = gb[1], y2 = gb[2], y3 = gb[3], y4 = gb[4], y5 = gb[5], y6 = gb[6], y7 = gb[7];
  uint4* s = reinterpret_cast<uint4*>(W->buf);
  s[0] = x0; s[1] = x1; s[2] = x2; s[3] = x3; s[4] = x4; s[5] = x5; s[6] = x6; s[7] = x7;
  s[8] = y0; s[9] = y1; s[10] = y2; s[11] = y3; s[12] = y4; s[13] = y5; s[14] = y6; s[15] = y7;
  W->line[0] = a; W->line[1] = b;
}
#endif

// byte at arena offset p
B2_HD uint8_t rd8(const Cursor& c, uint64_t p) {
#if defined(__CUDA_ARCH__)
  Win* W = c.win;
  const uint64_t addr = (uint64_t)(uintptr_t)(c.w + p), line = addr & ~127ull;
  if (line == W->line[0]) return W->buf[addr & 127];
  if (line == W->line[1]) return W->buf[128 + (addr & 127)];
  const uint32_t k = W->victim;
  W->victim = k ^ 1;
  win_fill(W, k, line);
  return W->buf[128 * k + (addr & 127)];
#else
  return c.w[p];
#endif
}

B2_HD uint64_t rd_varint(Cursor& c) {
  uint64_t v = 0;
  for (int i = 0; i < 10; ++i) {
    if (c.p >= c.end) { c.err = B200TFS_E_PARSE; return 0; }
    uint8_t b = rd8(c, c.p++);
    v |= (uint64_t)(b & 0x7F) << (7 * i);  // bits past 64 fall off, as in the runtime
    if (!(b & 0x80)) return v;
  }
  c.err = B200TFS_E_PARSE;  // 10 continuation bytes
  return 0;
}

// tag: must fit 32 bits, field number != 0
B2_HD uint32_t rd_tag(Cursor& c) {
  uint64_t t = rd_varint(c);
  if (c.err) return 0;
  if (t > 0xFFFFFFFFull || (t >> 3) == 0) { c.err = B200TFS_E_PARSE; return 0; }
  return (uint32_t)t;
}

// length prefix: bounded by the enclosing message (and by int32, like the runtime)
B2_HD uint64_t rd_len(Cursor& c) {
  uint64_t n = rd_varint(c);
  if (c.err) return 0;
  if (n > 0x7FFFFFFFull || n > c.end - c.p) { c.err = B200TFS_E_PARSE; return 0; }
  return n;
}

// Skip one non-group field body (tag already consumed).
B2_HD void skip_scalar(Cursor& c, uint32_t wt) {
  if (wt == WT_VARINT) { (void)rd_varint(c); return; }
  if (wt == WT_I64) { if (c.end - c.p < 8) c.err = B200TFS_E_PARSE; else c.p += 8; return; }
  if (wt == WT_I32) { if (c.end - c.p < 4) c.err = B200TFS_E_PARSE; else c.p += 4; return; }
  if (wt == WT_LEN) { uint64_t n = rd_len(c); if (!c.err) c.p += n; return; }
  c.err = B200TFS_E_PARSE;
}

// Skip one field body of any wire type.  Groups nest (explicit stack, no recursion: device stack
// frames stay static); an END_GROUP that does not close a group we opened is malformed.
B2_HD void skip_field(Cursor& c, uint32_t tag) {
  const uint32_t wt = tag & 7;
  if (wt != WT_SGROUP) { skip_scalar(c, wt); return; }  // stray END_GROUP and wire types 6, 7 fail in there
  uint32_t open[32];
  int depth = 0;
  open[depth++] = tag >> 3;
  while (depth > 0 && !c.err) {
    if (c.p >= c.end) { c.err = B200TFS_E_PARSE; return; }
    const uint32_t t = rd_tag(c);
    if (c.err) return;
    const uint32_t w2 = t & 7;
    if (w2 == WT_SGROUP) {
      if (depth >= 32) { c.err = B200TFS_E_PARSE; return; }
      open[depth++] = t >> 3;
    } else if (w2 == WT_EGROUP) {
      if (open[depth - 1] != (t >> 3)) { c.err = B200TFS_E_PARSE; return; }
      --depth;
    } else {
      skip_scalar(c, w2);
    }
  }
}

// Structural UTF-8 check the runtime applies to proto3 `string` fields (shortest form, no
// surrogates, <= U+10FFFF).
B2_HD bool utf8_ok(const Cursor& c, uint64_t off, uint64_t n) {
  uint64_t i = 0;
  while (i < n) {
    uint8_t b = rd8(c, off + i);
    if (b < 0x80) { ++i; continue; }
    uint32_t need; uint32_t cp;
    if (b >= 0xC2 && b <= 0xDF) { need = 1; cp = b & 0x1F; }
    else if (b >= 0xE0 && b <= 0xEF) { need = 2; cp = b & 0x0F; }
    else if (b >= 0xF0 && b <= 0xF4) { need = 3; cp = b & 0x07; }
    else return false;
    if (n - i - 1 < need) return false;
    for (uint32_t k = 1; k <= need; ++k) {
      uint8_t x = rd8(c, off + i + k);
      if ((x & 0xC0) != 0x80) return false;
      cp = (cp << 6) | (x & 0x3F);
    }
    if (need == 2 && (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF))) return false;
    if (need == 3 && (cp < 0x10000 || cp > 0x10FFFF)) return false;
    i += need + 1;
  }
  return true;
}

// Raw value-field occurrences seen before the dtype is known.
struct RawChunks {
  uint32_t field[B200TFS_MAX_CHUNKS];
  uint64_t off[B200TFS_MAX_CHUNKS];
  uint64_t len[B200TFS_MAX_CHUNKS];
  int n;
  bool overflow;
};

B2_HD void out_reset(b200tfs_output& o) {
  o.key_off = 0; o.key_len = 0; o.dtype = 0; o.rank = 0; o.flags = 0; o.value_field = 0; o.n_chunks = 0;
  for (int i = 0; i < B200TFS_MAX_RANK; ++i) o.dims[i] = 0;
  for (int i = 0; i < B200TFS_MAX_CHUNKS; ++i) { o.chunk_off[i] = 0; o.chunk_len[i] = 0; }
  o.content_off = 0; o.content_len = 0; o.msg_off = 0; o.msg_len = 0;
  o.n_elems = 0; o.dst_bytes = 0; o.n_strings = 0; o.dst_off = 0; o.status = B200TFS_OK; o.reserved = 0;
}

// TensorShapeProto (tensor_shape.proto:13-46): dims append (merge); Dim.size last wins inside a Dim.
B2_HD void walk_shape(Cursor& c, b200tfs_output& o, bool& rank_overflow) {
  while (c.p < c.end && !c.err) {
    uint32_t tag = rd_tag(c);
    if (c.err) return;
    if (tag == tag_of(2, WT_LEN)) {  // dim
      uint64_t n = rd_len(c);
      if (c.err) return;
      Cursor d{c.w, c.p, c.p + n, 0, c.win};
      int64_t size = 0;
      while (d.p < d.end && !d.err) {
        uint32_t t = rd_tag(d);
        if (d.err) break;
        if (t == tag_of(1, WT_VARINT)) size = (int64_t)rd_varint(d);
        else if (t == tag_of(2, WT_LEN)) {  // name: a proto3 string, validated then ignored (tensors.py:38-39)
          uint64_t m = rd_len(d);
          if (d.err) break;
          if (!utf8_ok(d, d.p, m)) d.err = B200TFS_E_PARSE;
          d.p += m;
        } else skip_field(d, t);
      }
      if (d.err) { c.err = d.err; return; }
      if (o.rank < B200TFS_MAX_RANK) o.dims[o.rank++] = size; else rank_overflow = true;
      c.p += n;
    } else {
      skip_field(c, tag);  // unknown_rank (3) and anything else
    }
  }
}

// One TensorProto (tensor.proto:14-84); accumulates into o / raw so a repeated `value` merges.
B2_HD void walk_tensor(Cursor& c, b200tfs_output& o, RawChunks& raw, bool& rank_overflow) {
  while (c.p < c.end && !c.err) {
    uint32_t tag = rd_tag(c);
    if (c.err) return;
    uint32_t field = tag >> 3, wt = tag & 7;
    if (field == F_DTYPE && wt == WT_VARINT) {
      o.dtype = (int32_t)(uint32_t)rd_varint(c);
    } else if (field == F_SHAPE && wt == WT_LEN) {
      uint64_t n = rd_len(c);
      if (c.err) return;
      Cursor s{c.w, c.p, c.p + n, 0, c.win};
      walk_shape(s, o, rank_overflow);
      if (s.err) { c.err = s.err; return; }
      c.p += n;
    } else if (field == F_CONTENT && wt == WT_LEN) {
      uint64_t n = rd_len(c);
      if (c.err) return;
      o.content_off = c.p; o.content_len = n;  // `bytes`: last occurrence replaces
      c.p += n;
    } else if (field == F_STRING && wt == WT_LEN) {
      uint64_t n = rd_len(c);
      if (c.err) return;
      o.n_strings += 1;
      c.p += n;
    } else if (scalar_wire_type(field) != 0xFFu && (wt == WT_LEN || wt == scalar_wire_type(field))) {
      uint64_t off, len;
      if (wt == WT_LEN) {
        len = rd_len(c);
        if (c.err) return;
        off = c.p;
        uint32_t fw = fixed_wire_width(field);
        if (fw) {
          if (len % fw) { c.err = B200TFS_E_PARSE; return; }  // packed fixed32/64 must be whole elements
        } else if (len && (rd8(c, off + len - 1) & 0x80)) {
          c.err = B200TFS_E_PARSE; return;                    // packed varints must end on a terminator
        }
        c.p += len;
      } else {
        off = c.p;
        skip_field(c, tag);
        if (c.err) return;
        len = c.p - off;
      }
      if (len) {
        if (raw.n < B200TFS_MAX_CHUNKS) {
          raw.field[raw.n] = field; raw.off[raw.n] = off; raw.len[raw.n] = len; ++raw.n;
        } else raw.overflow = true;
      }
    } else {
      if (field > 17 || field == 0) o.flags |= B200TFS_OF_HAS_UNKNOWN;
      skip_field(c, tag);  // version_number, resource_handle_val, variant_val, mismatched wire types, unknown
    }
  }
}

// Settle dtype -> field, pick that field's chunks, element counts.  Mirrors what
// tensor_proto_to_ndarray (tensors.py:42-46) would conclude from the parsed message.
B2_HD void finalize_output(b200tfs_output& o, const RawChunks& raw, bool rank_overflow) {
  if (rank_overflow || raw.overflow) { o.status = B200TFS_E_NONCANONICAL; return; }
  DtypeInfo di = dtype_info(o.dtype);
  if (di.field == 0) { o.status = B200TFS_E_KEY; return; }  // types.py:40 KeyError
  o.value_field = (int32_t)di.field;
  uint64_t total = 0;
  for (int i = 0; i < raw.n; ++i) {
    if (raw.field[i] == di.field) {
      o.chunk_off[o.n_chunks] = raw.off[i]; o.chunk_len[o.n_chunks] = raw.len[i];
      total += raw.len[i];
      ++o.n_chunks;
    }
  }
  if (o.n_chunks > 1) o.flags |= B200TFS_OF_MULTI_CHUNK;
  if (o.content_len) o.flags |= B200TFS_OF_TENSOR_CONTENT;
  if (o.rank == 0) o.flags |= B200TFS_OF_RANK0;
  // prod(dims) with at most one -1
  uint64_t prod = 1; int infer = -1; bool bad = false;
  for (int i = 0; i < o.rank; ++i) {
    int64_t d = o.dims[i];
    if (d == -1 && infer < 0) { infer = i; continue; }
    if (d < 0) { bad = true; break; }
    if (d != 0 && prod > 0xFFFFFFFFFFFFFFFFull / (uint64_t)d) { bad = true; break; }
    prod *= (uint64_t)d;
  }
  if (bad) { o.status = B200TFS_E_SHAPE; return; }
  if (di.kind == VK_FIXED) {
    uint64_t count = total / di.elem_size;  // complex: interleaved (re, im) pairs, TF convention
    if (total % di.elem_size) { o.status = B200TFS_E_SHAPE; return; }
    if (infer >= 0) {
      if (prod == 0 || count % prod) { o.status = B200TFS_E_SHAPE; return; }
      o.dims[infer] = (int64_t)(count / prod); prod = count; o.flags |= B200TFS_OF_DIM_INFERRED;
    }
    if (count != prod) { o.status = B200TFS_E_SHAPE; return; }  // reshape() ValueError: no broadcast, no padding
  } else if (di.kind == VK_VARINT || di.kind == VK_BOOL) {
    o.flags |= B200TFS_OF_VARINT;
    if (infer >= 0) { o.status = B200TFS_E_NONCANONICAL; return; }
    if ((total == 0) != (prod == 0)) { o.status = B200TFS_E_SHAPE; return; }
    if (total < prod) { o.status = B200TFS_E_SHAPE; return; }   // every element needs at least one byte
  } else {  // strings: unpacked on the host from msg_off/msg_len
    if (infer >= 0) {
      if (prod == 0 || o.n_strings % prod) { o.status = B200TFS_E_SHAPE; return; }
      o.dims[infer] = (int64_t)(o.n_strings / prod); prod = o.n_strings; o.flags |= B200TFS_OF_DIM_INFERRED;
    }
    if (o.n_strings != prod) { o.status = B200TFS_E_SHAPE; return; }
  }
  o.n_elems = prod;
  o.dst_bytes = prod * di.elem_size;
}

B2_HD bool bytes_equal(const Cursor& c, uint64_t a, uint64_t b, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) if (rd8(c, a + i) != rd8(c, b + i)) return false;
  return true;
}

B2_HD void spec_reset(b200tfs_model_spec& s) {
  s.name_off = 0; s.name_len = 0; s.signature_len = 0; s.signature_off = 0; s.label_off = 0; s.label_len = 0;
  s.has_version = 0; s.version = 0;
}

// ModelSpec (model.proto:9-33); repeated occurrences merge.
B2_HD void walk_model_spec(Cursor& c, b200tfs_model_spec& s) {
  while (c.p < c.end && !c.err) {
    uint32_t tag = rd_tag(c);
    if (c.err) return;
    if (tag == tag_of(1, WT_LEN) || tag == tag_of(3, WT_LEN) || tag == tag_of(4, WT_LEN)) {
      uint64_t n = rd_len(c);
      if (c.err) return;
      if (!utf8_ok(c, c.p, n)) { c.err = B200TFS_E_PARSE; return; }
      if ((tag >> 3) == 1) { s.name_off = c.p; s.name_len = (uint32_t)n; }
      else if ((tag >> 3) == 3) { s.signature_off = c.p; s.signature_len = (uint32_t)n; }
      else { s.label_off = c.p; s.label_len = (uint32_t)n; s.has_version = 0; s.version = 0; }  // oneof: label displaces version
      c.p += n;
    } else if (tag == tag_of(2, WT_LEN)) {  // google.protobuf.Int64Value version
      uint64_t n = rd_len(c);
      if (c.err) return;
      Cursor v{c.w, c.p, c.p + n, 0, c.win};
      if (!s.has_version) s.version = 0;
      while (v.p < v.end && !v.err) {
        uint32_t t = rd_tag(v);
        if (v.err) break;
        if (t == tag_of(1, WT_VARINT)) s.version = (int64_t)rd_varint(v); else skip_field(v, t);
      }
      if (v.err) { c.err = v.err; return; }
      s.has_version = 1; s.label_off = 0; s.label_len = 0;  // oneof: version displaces label
      c.p += n;
    } else skip_field(c, tag);
  }
}

// One PredictResponse (predict.proto:30-40).  Returns the record status; *n_outs distinct keys.
B2_HD int walk_response(const uint8_t* w, uint64_t off, uint64_t len, int max_outputs, b200tfs_output* outs,
                        int* n_outs, b200tfs_model_spec* spec, Win* win = nullptr) {
  Cursor c{w, off, off + len, 0, win};
  int n = 0;
  spec_reset(*spec);
  *n_outs = 0;
  while (c.p < c.end && !c.err) {
    uint32_t tag = rd_tag(c);
    if (c.err) break;
    if (tag == tag_of(1, WT_LEN)) {  // outputs map entry
      uint64_t elen = rd_len(c);
      if (c.err) break;
      Cursor e{w, c.p, c.p + elen, 0, c.win};
      b200tfs_output tmp; out_reset(tmp);
      RawChunks raw; raw.n = 0; raw.overflow = false;
      bool rank_overflow = false;
      while (e.p < e.end && !e.err) {
        uint32_t t = rd_tag(e);
        if (e.err) break;
        if (t == tag_of(1, WT_LEN)) {
          uint64_t k = rd_len(e);
          if (e.err) break;
          if (!utf8_ok(e, e.p, k)) { e.err = B200TFS_E_PARSE; break; }
          tmp.key_off = e.p; tmp.key_len = (uint32_t)k;
          e.p += k;
        } else if (t == tag_of(2, WT_LEN)) {
          uint64_t m = rd_len(e);
          if (e.err) break;
          Cursor tc{w, e.p, e.p + m, 0, c.win};
          tmp.msg_off = e.p; tmp.msg_len = m;
          walk_tensor(tc, tmp, raw, rank_overflow);
          if (tc.err) { e.err = tc.err; break; }
          e.p += m;
        } else skip_field(e, t);
      }
      if (e.err) { c.err = e.err; break; }
      c.p += elen;
      finalize_output(tmp, raw, rank_overflow);
      // duplicate key: the later entry replaces the earlier one
      int slot = -1;
      for (int i = 0; i < n; ++i)
        if (outs[i].key_len == tmp.key_len && bytes_equal(c, outs[i].key_off, tmp.key_off, tmp.key_len)) { slot = i; break; }
      if (slot < 0) {
        if (n >= max_outputs) return B200TFS_E_SIZE;
        slot = n++;
      }
      outs[slot] = tmp;
    } else if (tag == tag_of(2, WT_LEN)) {
      uint64_t m = rd_len(c);
      if (c.err) break;
      Cursor sc{w, c.p, c.p + m, 0, c.win};
      walk_model_spec(sc, *spec);
      if (sc.err) { c.err = sc.err; break; }
      c.p += m;
    } else skip_field(c, tag);
  }
  if (c.err) return c.err;
  *n_outs = n;
  return B200TFS_OK;
}

// A bare TensorProto message (what tensor_proto_to_ndarray receives).
B2_HD int walk_tensor_proto(const uint8_t* w, uint64_t off, uint64_t len, b200tfs_output* out, Win* win = nullptr) {
  Cursor c{w, off, off + len, 0, win};
  out_reset(*out);
  RawChunks raw; raw.n = 0; raw.overflow = false;
  bool rank_overflow = false;
  out->msg_off = off; out->msg_len = len;
  walk_tensor(c, *out, raw, rank_overflow);
  if (c.err) return c.err;
  finalize_output(*out, raw, rank_overflow);
  return B200TFS_OK;
}

}  // namespace b200tfs
