// plan.h - the launch plan shared by the host planner (codec_host.cpp) and the kernels (kernels.cu).
//
// A plan is a flat byte image: PlanHeader, then MoveItem[n_items], TileRef[n_tiles] (absent when
// uniform_tpi != 0), SmallItem[n_small], then the header blob (the varint tags / lengths / dims /
// keys the planner computed - the non-payload bytes of the wire).  It reaches the kernel either
// by value in the kernel parameter space (<= kInlinePlanBytes, no copy at all: the C2 single-tensor
// case) or through one pinned-host -> device copy.
#pragma once
#include <stdint.h>

#include "../../include/b200tfs.h"

namespace b200tfs {

// what a MoveItem / SmallItem does to the bytes it moves
enum MoveOp : uint32_t {
  OP_COPY = 0,        // raw little-endian bytes (float64, complex, tensor_content, KEEP_SNAN float32)
  OP_QUIET_SRC = 1,   // float32 sNaN -> qNaN, 4-byte elements aligned with the SOURCE start (encode)
  OP_QUIET_DST = 2,   // same, elements aligned with the DESTINATION start (decode)
  OP_BOOL = 3,        // byte != 0 -> 1 (bool_val written from numpy bool memory)
  OP_H2F = 4,         // float16  -> float32 (exact), encode-side cast
  OP_B2F = 5,         // bfloat16 -> float32 (exact), encode-side cast
  OP_F2H = 6,         // float32 -> float16, round-to-nearest-even, decode-side cast
  OP_F2B = 7,         // float32 -> bfloat16, round-to-nearest-even, decode-side cast
  OP_FLAG_BLOB = 0x80000000u  // SmallItem: src is an offset into the plan image, not a pointer
};

struct MoveItem {     // one large payload, tiled across CTAs
  const uint8_t* src;
  uint8_t* dst;
  uint64_t n_out;     // bytes written
  uint32_t op;
  uint32_t n_tiles;   // max(1, ceil(ceil(n_out/16) / vec_per_tile))
  uint32_t glen, gstride;  // gstride != 0: the source is a row of pieces, `glen` value bytes every `gstride` bytes (a run of
                           // unpacked elements, b200tfs_run): logical source byte j lies at src[(j / glen) * gstride + j % glen]
};

struct TileRef { uint32_t item; uint32_t tile; };

struct SmallItem {    // one header fragment or small payload: a warp moves it
  uint64_t src;       // pointer, or offset into the plan image with OP_FLAG_BLOB
  uint8_t* dst;
  uint32_t n_out;
  uint32_t op;
  uint32_t glen, gstride;  // as in MoveItem
};

struct PlanHeader {
  uint32_t n_items, n_tiles, n_small, uniform_tpi;  // uniform_tpi: every item has this many tiles (no TileRef table)
  uint32_t vec_per_tile;                            // 16-byte vectors of destination per tile
  uint32_t off_items, off_tiles, off_small;         // byte offsets inside the plan image
  uint32_t guard_div;                               // move_guarded_kernel: item i is stored only if guard[i / guard_div] != 0
  uint32_t independent;                             // != 0: nothing in this plan is written by the kernel launched just before (the framing
                                                    // kernel of a deferred encode whose movers all have host-fixed destinations): no wait for it
  uint32_t pad[2];
  const uint32_t* guard;
};

constexpr uint32_t kInlinePlanBytes = 3840;  // fits the classic 4 KB kernel-parameter window
struct InlinePlan { uint8_t bytes[kInlinePlanBytes]; };

constexpr uint32_t kMoveThreads = 256;      // threads per CTA of move_kernel
constexpr uint32_t kSmallMax = 2048;        // payloads up to this many bytes take the warp path

// ---- fused single-launch decode ---------------------------------------------------------------
constexpr int kFusedMaxOutputs = 8;    // outputs tabulated per record by decode_fused_kernel
constexpr int kFusedInlineRecs = 16;   // up to this many records travel in the kernel parameters
constexpr uint32_t kFusedSlackTiles = 8;  // tiles budgeted per record beyond ceil(rec_len / tile)

struct FusedInline { uint64_t off[kFusedInlineRecs], len[kFusedInlineRecs]; uint32_t tile_start[kFusedInlineRecs + 1]; };

// framing template left by one decode launch for the next (see decode_fused_kernel)
constexpr uint32_t kTplChunks = 4;
constexpr uint32_t kTplFraming = 256;   // == kMoveThreads: one framing byte per thread
struct TplChunk { uint32_t wire_off, len, dst_off, op, n_tiles, is_varint, fpos, pad; };  // record-relative
struct TplHead {
  uint32_t valid, n_chunks, n_outs, framing_len;
  uint64_t rec_len, dst_need;
  uint32_t vpt, total_tiles;
  uint32_t serial, cast;       // serial: which learning produced it (the table entries below belong to exactly this serial);
                               // cast: the float narrowing it was laid out for (0 / DT_HALF / DT_BFLOAT16: FusedParams::cast)
};
// the part of a template every CTA needs before it can start on its tile: small enough to ride in the kernel parameters
// (no load at all ahead of the tile's loads) when the host knows it - because it walked record 0 itself (host-buffer entry
// points) or because an earlier launch left it in pinned memory and the stream has been idle since
struct TplInline {
  TplHead head;
  TplChunk chunk[kTplChunks];
  uint8_t framing[kTplFraming];
};
struct Template {
  TplInline in;
  b200tfs_model_spec spec;
  b200tfs_output outs[kFusedMaxOutputs];
};

struct FusedParams {
  const uint8_t* w;          // wire arena
  uint8_t* dst;              // destination base; record r owns [r*dst_stride, (r+1)*dst_stride)
  uint64_t dst_stride;
  int32_t n;
  uint32_t vpt;
  b200tfs_output* outs;      // device table, n * kFusedMaxOutputs
  int32_t* n_outs;
  b200tfs_model_spec* specs;
  int32_t* status;
  const uint32_t* cta_rec;   // n > kFusedInlineRecs: record of every CTA, tile_start[n+1], rec_off, rec_len
  const uint32_t* tile_start;
  const uint64_t* rec_off;
  const uint64_t* rec_len;
  const Template* tpl_read;  // template written by the previous launch on this context (may be invalid)
  Template* tpl_write;       // where record 0 of this launch leaves its template
  TplInline* tpl_pinned;     // pinned host copy of the inline part, written whenever a template is learnt (valid flag last)
  unsigned long long* stats; // device counters: records served by [0] the template in the parameters, [1] the device template, [2] the walk
  uint32_t serial;           // stamp for a template learnt by THIS launch
  uint32_t cast;             // DT_FLOAT outputs are narrowed on the way out: 0 = no, DT_HALF (19) / DT_BFLOAT16 (14) (b200tfs_set_decode_cast)
  uint32_t mode;             // 0: the whole decode.  The narrowing decode of a batch whose template the host knows runs as three launches:
                             // 1 = verify only (two CTAs per record: framing verdict -> guard[r], table), then move_guarded_kernel over a
                             // host-built plan, then 2 = the whole decode for the records whose guard is still 0 (the others leave at once)
  uint32_t* guard;
  uint32_t tile_bias;        // a SLICE of a one-record launch (the pipelined host path): CTA b works as CTA b + tile_bias of the full grid
  uint32_t trusted;          // != 0: the host built the inline template from THIS record's own bytes: no verdict (a slice's launch runs
                             // before the record's tail - and with it part of the framing - has arrived on the device)
  FusedInline inl;
  TplInline tpli;            // head.valid != 0: the template as the host knows it (tier 1; tpl_read is tier 2, the walk tier 3)
};

// ---- packed-varint jobs ------------------------------------------------------------------------
constexpr uint32_t kVarThreads = 256;
constexpr uint32_t kVarPerThread = 8;
constexpr uint32_t kVarTileElems = kVarThreads * kVarPerThread;  // elements per encode tile
constexpr uint32_t kVarTileBytes = kVarThreads * 32;             // wire bytes per decode tile; tiles are cut at 16-byte-aligned ADDRESSES
constexpr int32_t kVarFlagHalfAsValue = 1;  // decode half_val ints as VALUES (the reference's DT_HALF quirk)
constexpr int32_t kVarFlagPadEdge = 2;      // fewer values than elements is not an error: the caller pads (B200TFS_OF_PAD_EDGE)

struct VarSeg {       // a contiguous run of one job: the whole tensor (encode) or one wire chunk (decode)
  const uint8_t* src;
  uint64_t n;         // elements (encode) or bytes (decode)
  uint32_t job;
  uint32_t first_tile;
};

constexpr uint32_t kVarGroupTiles = kVarThreads;  // tiles per counter group (one counter per thread when a CTA sums its prefix)

struct VarJobDev {
  uint8_t* dst;        // encode: first payload byte on the wire; decode: first element of the tensor
  uint64_t n_elems;
  uint64_t cap;        // encode: payload bytes the header announced (never written past)
  uint32_t* tile_val;  // [n_tiles] bytes (encode) / terminators (decode) of every tile of the job
  uint32_t* group_sum; // [ceil(n_tiles / kVarGroupTiles)] sums of tile_val; zeroed before the counting kernel
  unsigned long long* total;  // sum over the job; zeroed likewise
  int32_t* status;     // decode: B200TFS_OK or the first error
  int32_t dtype;       // DT_* of the tensor in memory
  uint32_t elem_size;
  uint32_t is_signed;
  int32_t flags;
  uint32_t first_tile;
  uint32_t n_tiles;
  uint32_t fuse_group0;  // single-pass kernels: the job's first group descriptor (VarFuse)
  uint32_t pad;
};

// what every varint kernel receives (by value, in the parameter space).  A single job with a single segment - one
// big tensor, the case where the bandwidth matters - travels inline, so that no CTA starts with three dependent loads.
struct VarTables {
  const VarSeg* segs;
  const uint32_t* tile_seg;
  const VarJobDev* jobs;
  uint32_t n_tiles;
  uint32_t single;      // 1: seg0 / job0 below describe every tile
  VarSeg seg0;
  VarJobDev job0;
};

// ---- deferred framing: the length prefixes of packed-varint inputs computed ON THE DEVICE -----------------------------
// Every length on the wire precedes its content, and a packed-varint payload's length is only known once the counting
// kernel has run.  Instead of bringing it to the host (b200tfs_measure: a stream synchronise in the middle of an encode),
// the host describes each request as a little program - byte runs it knows, varints of VALUES the device evaluates
// (sums of constants, job totals, other values and the varint lengths of other values), and the payloads whose
// destinations depend on those values - and frame_requests_kernel (one thread per request) lays the record out, writes the
// framing and patches the destinations into the move plan / the emit jobs that run right behind it.  No host round trip:
// the whole encode is asynchronous and CUDA-graph capturable.
enum FrameSegKind : uint32_t {
  FS_BYTES = 0,   // a = offset into the frame blob, b = byte count
  FS_VARINT = 1,  // a = value id (request-local): varint(value)
  FS_BE32 = 2,    // a = value id: four bytes, big endian (gRPC's message length)
  FS_ITEM = 3,    // a = MoveItem index, b = its byte count: the payload lands here (dst patched)
  FS_SMALL = 4,   // a = SmallItem index, b = its byte count: likewise
  FS_VARJOB = 5,  // a = varint job index, b = value id of its packed length: the varints land here (dst and cap patched)
  FS_ANCHOR = 7,  // a = fused varint job index, b = value id of its packed length: the payload is ALREADY in place at the record's
                  // anchor (venc_fused_kernel wrote it before the framing kernel ran); everything before it is laid out backwards
  FS_TINYVAR = 6  // a = index into FrameTables::tiny, b = value id of its packed length: a packed-varint input of at most
                  // kTinyVarElems elements (a label, an id, a few flags) is counted AND written by the framing kernel itself -
                  // no counting kernel, no emit kernel, no counters to zero for it
};
constexpr uint32_t kTinyVarElems = 32;
struct TinyVar { const uint8_t* src; uint32_t n, elem_size, is_signed, pad; };
struct FrameSeg { uint32_t kind, a, b, pad; };
enum FrameTermKind : uint32_t { FT_TOTAL = 0, FT_VAL = 1, FT_VLEN = 2, FT_TINY = 3, FT_TOTALF = 4 };   // FT_TOTALF: + totals_fused[job]   // + total[job], + value[i], + varint_len(value[i]), + packed length of tiny[idx]
struct FrameTerm { uint32_t kind, idx; };
struct FrameVal { int64_t c; uint32_t first_term, n_terms; };               // evaluated in order: terms refer to earlier values only
struct FrameReq {
  uint32_t first_seg, n_seg, first_val, n_val;
  uint32_t first_term, n_term, first_blob, n_blob;
  uint32_t align_seg;       // the payload segment that should start 128-byte aligned (index relative to first_seg), ~0u: none
  uint32_t anchor_seg;      // ~0u, or the FS_ANCHOR segment: its first byte lies at anchor_off and the record is laid out around it
  uint32_t pad0;
  uint64_t anchor_off;
  uint32_t total_val;       // value id of the record's byte length (incl. a gRPC prefix)
  uint64_t slot_off, slot_cap;   // where the record may lie inside the arena (worst-case sized by the host)
};
struct FrameTables {
  const FrameReq* reqs; const FrameSeg* segs; const FrameVal* vals; const FrameTerm* terms; const uint8_t* blob;
  const unsigned long long* totals;   // packed length of every varint job (the counting kernel's result), by job index
  const TinyVar* tiny;                // FS_TINYVAR / FT_TINY
  const unsigned long long* totals_fused;   // packed length of every single-pass job (FT_TOTALF)
  uint64_t* scratch_vals;   // one evaluated value per FrameVal
  uint64_t* scratch_terms;  // one fetched total per FrameTerm (used by the table-walking path)
  uint8_t* arena;
  MoveItem* items; SmallItem* smalls; VarJobDev* jobs;    // patched
  uint64_t* rec_off; uint64_t* rec_len; int32_t* status;  // pinned host memory: read by b200tfs_encode_results
  uint32_t n;
};

// single-pass varint encode (venc_fused_kernel): the look-back state behind VarTables.  VarJobDev::tile_val holds the per-tile
// state (flag | count), VarJobDev::fuse_group0 the job's first group descriptor; everything is zeroed before the launch.
struct VarFuse {
  uint32_t* ticket;                    // tiles take their number from here
  unsigned long long* group_state;     // per group of 32 tiles: 2-bit flag (1 = sum known, 2 = inclusive prefix known) | bytes
  uint32_t* group_arrivals;            // tiles of the group that have published their count
};

// decode tiles of a chunk [src, src + n): aligned windows of kVarTileBytes starting at src rounded down to 16
#if defined(__CUDACC__)
#define B2_PLAN_HD __host__ __device__ __forceinline__
#else
#define B2_PLAN_HD inline
#endif
B2_PLAN_HD uint64_t var_decode_tiles(const void* src, uint64_t n) {
  return n ? (((uint64_t)((uintptr_t)src & 15) + n + kVarTileBytes - 1) / kVarTileBytes) : 0;
}

}  // namespace b200tfs
