// codec_host.cpp - host half of libb200tfs.so: the C ABI of include/b200tfs.h, the wire-size
// arithmetic, the header planner (tags, varint lengths, dims, keys - everything on the wire that is
// not tensor payload) and the launch plans handed to the kernels in kernels.cu.
//
// The planner restates the framing the reference gets from protobuf for
//   TensorProto{dtype, tensor_shape{dim{size}}, <typed packed field>}      tensors.py:28-35
//   PredictRequest{model_spec{name, version{value}}, inputs{key -> proto}}  requests.py:41-48
// (proto3 rules: fields in ascending number, zero scalars elided, packed repeated scalars, map entries
// always carry key and value) - SURVEY.md 8(a) a1-a3 and quirks Q1-Q5.  There is deliberately no CPU
// implementation of the payload path here: payload bytes only ever move inside the CUDA kernels.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/b200tfs.h"
#include "frame.h"
#include "kernels.h"
#include "plan.h"
#include "tpl.h"
#include "walker.h"
#include "wire.h"

using namespace b200tfs;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail(B200TFS_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_));   \
  } while (0)

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
namespace {

struct Growable {  // device or pinned buffer that only grows
  void* p = nullptr;
  uint64_t cap = 0;
};

constexpr int kSlots = 4;

struct Slot {  // one in-flight plan upload: pinned image + device image + completion event
  Growable host, dev;
  cudaEvent_t done = nullptr;
  bool pending = false;
};

}  // namespace

struct MeasuredTensor {     // what b200tfs_measure learnt about one varint tensor (device addresses of its counters)
  uint64_t n_elems = 0, packed_len = 0;
  int32_t dtype = 0;
  void* tile_val = nullptr;
  void* group_sum = nullptr;
  void* total = nullptr;
};

struct b200tfs_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;       // the stream every call is ordered on: the context's own, or the caller's (b200tfs_set_stream)
  cudaStream_t own_stream = nullptr;
  cudaStream_t aux_stream = nullptr;   // uploads done at capture time, outside the graph being recorded; the pipelined host path's H2D copies
  cudaStream_t d2h_stream = nullptr;   // the pipelined host path's D2H copies
  static constexpr int kPipeMax = 8;   // slices of one pipelined host call
  cudaEvent_t pipe_ev[2 * kPipeMax + 2] = {};
  uint64_t pipe_min = 1ull << 20;      // *_host calls moving at least this many payload bytes are sliced (B200TFS_PIPELINE_MIN; 0 = never)
  int pipe_max = 4;                    // at most this many slices (B200TFS_PIPELINE_SLICES, 2..kPipeMax): every slice costs ~7 driver calls
  uint64_t pipelined_calls = 0;        // how many host calls took the sliced path (tests)
  uint64_t direct_calls = 0;           // ... and how many wrote their output straight into the caller's pinned buffer
  bool opt_direct_out = true;          // B200TFS_DIRECT_OUT=0: always stage the output on the device and copy it back
  Growable guard_dev;                  // the narrowing batch decode's per-record verdicts (FusedParams::guard)
  uint32_t decode_cast = 0;            // b200tfs_set_decode_cast: DT_FLOAT outputs of the single-launch decode leave as DT_HALF / DT_BFLOAT16
  Slot slots[kSlots];
  int next_slot = 0;
  Growable scratch_dev;   // parse tables / varint tile tables
  Growable scratch_host;  // pinned mirror of the parse tables
  Growable stage_dev;     // *_host entry points: tensors (encode) or wire (decode) staged on the device
  Growable arena_dev;     // *_host entry points: wire arena (encode) / unpacked tensors (decode)
  uint64_t launches = 0;
  uint32_t tile_bytes_override = 0;
  bool capturing = false;   // between b200tfs_capture_begin / _end: no syncs; uploads get buffers that live as long as the context
  std::vector<Slot*> graph_slots;  // plan images referenced by captured graphs
  Growable fused_host;      // decode_fused tables: pinned host memory the kernel writes directly
  void* tpl_dev = nullptr;  // two framing templates (device), used alternately by successive decode launches
  uint32_t tpl_flip = 0;
  TplInline* tpl_pinned = nullptr;   // where a kernel that learns a template leaves its inline part (pinned, mapped)
  TplInline tpl_known{};             // the newest template the host knows: its own walk of a host-resident record 0, or tpl_pinned
                                     // as found with the stream idle; rides in the kernel parameters of the next single-response launch
  uint32_t serial = 0;               // stamp of the next template learnt
  cudaEvent_t tpl_event = nullptr;   // recorded behind every eager decode launch: once it has completed, tpl_pinned is current
  bool tpl_event_pending = false;
  bool opt_no_inline = false;        // B200TFS_NO_INLINE_TEMPLATE=1: never hand the template over in the kernel parameters (experiments)
  bool opt_flat_budget = false;      // B200TFS_FLAT_BUDGET=1: every record gets ceil(len / tile) + 8 CTAs whatever the host knows (experiments)
  bool opt_table_dev = false;        // B200TFS_TABLE_DEV=1: the decode table goes to device memory and is fetched by b200tfs_decode_results
  Growable fused_dev;
  uint64_t stage_shift = 0;          // *_host decode: the wire sits at stage_dev + stage_shift (placed so that the payload is 16-byte aligned)
  int32_t fused_n = 0;      // records of the last b200tfs_decode_responses
  std::vector<int32_t> pending_status;   // varint decode: which output each status word in scratch_host belongs to
  // counters left by b200tfs_measure, keyed by tensor address; consumed by the encode that follows
  Growable measured_dev;
  uint64_t measured_used = 0;
  std::unordered_map<const void*, MeasuredTensor> measured;
  // spill area of the two-phase parse: dims / value runs beyond the table's inline arrays (walker.h SpillEntry)
  Growable spill_dev;
  std::vector<SpillEntry> spill_host;   // host copy of the last parse's area (empty when no record spilled)
  uint32_t spill_per_rec = 16;          // entries per record the next parse starts with
  uint32_t spill_last_per = 0;          // geometry of spill_host
  int32_t spill_last_n = 0;
  Growable gather_dev;                  // unpack: strided / many-run varint outputs are first gathered into one stream here
  Growable enc_host;                    // b200tfs_encode_requests_async: rec_off | rec_len | status, written by frame_requests_kernel (pinned)
  int32_t enc_n = 0;
  bool has_graphs = false;              // a CUDA graph captured on this context refers to the scratch buffers: they may not move any more
};

// `baked`: the buffer's address ends up inside captured graphs (every context-owned scratch buffer except the plan-upload
// slots, which captured launches replace by private ones).  Once a graph exists, such a buffer must not be reallocated:
// replays would write through the stale address.  Size everything with the largest call BEFORE capturing.
static const char* kGraphPinned = "a CUDA graph captured on this context refers to its scratch buffers, which this larger call would have to "
                                  "reallocate: run the largest call once before capturing, or use another context";
static int grow_dev(b200tfs_ctx* c, Growable& g, uint64_t need, bool baked = true) {
  if (need <= g.cap) return B200TFS_OK;
  if (c->capturing) return fail(B200TFS_E_ARG, "scratch buffer would have to grow during graph capture: run the call once before capturing");
  if (baked && c->has_graphs) return fail(B200TFS_E_ARG, "%s", kGraphPinned);
  uint64_t cap = std::max<uint64_t>(need, g.cap * 2);
  cap = (cap + 0xFFFFull) & ~0xFFFFull;
  CU(cudaStreamSynchronize(c->stream));
  if (g.p) CU(cudaFree(g.p));
  g.p = nullptr; g.cap = 0;
  CU(cudaMalloc(&g.p, cap));
  g.cap = cap;
  return B200TFS_OK;
}
static int grow_host(b200tfs_ctx* c, Growable& g, uint64_t need, bool baked = true) {
  if (need <= g.cap) return B200TFS_OK;
  if (c->capturing) return fail(B200TFS_E_ARG, "scratch buffer would have to grow during graph capture: run the call once before capturing");
  if (baked && c->has_graphs) return fail(B200TFS_E_ARG, "%s", kGraphPinned);
  uint64_t cap = std::max<uint64_t>(need, g.cap * 2);
  cap = (cap + 0xFFFull) & ~0xFFFull;
  CU(cudaStreamSynchronize(c->stream));
  if (g.p) CU(cudaFreeHost(g.p));
  g.p = nullptr; g.cap = 0;
  CU(cudaHostAlloc(&g.p, cap, cudaHostAllocPortable | cudaHostAllocMapped));
  g.cap = cap;
  return B200TFS_OK;
}

// claim an upload slot with room for `bytes` in both images
static int claim_slot(b200tfs_ctx* c, uint64_t bytes, Slot** out) {
  if (c->capturing) {
    // a captured launch must keep its plan image for as long as the graph may be replayed: give it private
    // buffers (the upload itself is recorded as a copy node and simply repeats on every replay)
    Slot* g = new Slot();
    cudaError_t e = cudaHostAlloc(&g->host.p, bytes, cudaHostAllocPortable);
    if (e == cudaSuccess) e = cudaMalloc(&g->dev.p, bytes);
    if (e != cudaSuccess) {
      if (g->host.p) cudaFreeHost(g->host.p);
      delete g;
      return fail(B200TFS_E_CUDA, "plan buffers for a captured launch: %s", cudaGetErrorString(e));
    }
    g->host.cap = g->dev.cap = bytes;
    c->graph_slots.push_back(g);
    *out = g;
    return B200TFS_OK;
  }
  Slot& s = c->slots[c->next_slot];
  c->next_slot = (c->next_slot + 1) % kSlots;
  if (s.pending) { CU(cudaEventSynchronize(s.done)); s.pending = false; }
  int rc;
  if ((rc = grow_host(c, s.host, bytes, false))) return rc;
  if ((rc = grow_dev(c, s.dev, bytes, false))) return rc;
  *out = &s;
  return B200TFS_OK;
}

// Bring a slot's pinned image to its device image.  While a graph is being captured the slot is private to that graph and its
// image never changes (plans, tables and framing programs are functions of the call's arguments, which a graph freezes
// anyway): it is copied NOW, on a side stream, instead of being recorded as a copy node that every replay would repeat
// (2-3 us of stream time per node; a captured C3 encode had three of them).
static int upload_slot(b200tfs_ctx* c, Slot* slot, uint64_t bytes) {
  if (c->capturing) {
    CU(cudaMemcpyAsync(slot->dev.p, slot->host.p, bytes, cudaMemcpyHostToDevice, c->aux_stream));
    CU(cudaStreamSynchronize(c->aux_stream));
    return B200TFS_OK;
  }
  CU(cudaMemcpyAsync(slot->dev.p, slot->host.p, bytes, cudaMemcpyHostToDevice, c->stream));
  if (slot->done) { CU(cudaEventRecord(slot->done, c->stream)); slot->pending = true; }
  return B200TFS_OK;
}

extern "C" {

int b200tfs_abi_version(void) { return B200TFS_ABI_VERSION; }
const char* b200tfs_last_error(void) { return g_err; }

int b200tfs_device_count(int* count) {
  if (!count) return fail(B200TFS_E_ARG, "count is NULL");
  *count = 0;
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) { *count = 0; return fail(B200TFS_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
  return B200TFS_OK;
}

int b200tfs_create(int device, b200tfs_ctx** out) {
  if (!out) return fail(B200TFS_E_ARG, "out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(B200TFS_E_CUDA, "no CUDA device (%s): this library has no CPU path", e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(B200TFS_E_ARG, "device %d out of range (have %d)", device, n);
  CU(cudaSetDevice(device));
  b200tfs_ctx* c = new b200tfs_ctx();
  c->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
  e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete c; return fail(B200TFS_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
  c->stream = c->own_stream;
  e = cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete c; return fail(B200TFS_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
  for (auto& s : c->slots) {
    e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
    if (e != cudaSuccess) { delete c; return fail(B200TFS_E_CUDA, "cudaEventCreate: %s", cudaGetErrorString(e)); }
  }
  const char* tb = getenv("B200TFS_TILE_BYTES");
  if (tb) c->tile_bytes_override = (uint32_t)strtoul(tb, nullptr, 10);
  const char* oi = getenv("B200TFS_NO_INLINE_TEMPLATE");
  c->opt_no_inline = oi && oi[0] == '1';
  const char* ofb = getenv("B200TFS_FLAT_BUDGET");
  c->opt_flat_budget = ofb && ofb[0] == '1';
  const char* od = getenv("B200TFS_TABLE_DEV");
  c->opt_table_dev = od && od[0] == '1';
  e = cudaEventCreateWithFlags(&c->tpl_event, cudaEventDisableTiming);
  if (e != cudaSuccess) { delete c; return fail(B200TFS_E_CUDA, "cudaEventCreate: %s", cudaGetErrorString(e)); }
  e = cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking);
  for (auto& ev : c->pipe_ev) if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  if (e != cudaSuccess) { delete c; return fail(B200TFS_E_CUDA, "pipeline stream / events: %s", cudaGetErrorString(e)); }
  const char* pm = getenv("B200TFS_PIPELINE_MIN");
  if (pm) c->pipe_min = strtoull(pm, nullptr, 10);
  const char* pd = getenv("B200TFS_DIRECT_OUT");
  if (pd && pd[0] == '0') c->opt_direct_out = false;
  const char* ps = getenv("B200TFS_PIPELINE_SLICES");
  if (ps) c->pipe_max = std::min<int>(b200tfs_ctx::kPipeMax, std::max(2, atoi(ps)));
  *out = c;
  return B200TFS_OK;
}

int b200tfs_destroy(b200tfs_ctx* c) {
  if (!c) return B200TFS_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->stream != c->own_stream) cudaStreamSynchronize(c->own_stream);
  for (auto& s : c->slots) {
    if (s.host.p) cudaFreeHost(s.host.p);
    if (s.dev.p) cudaFree(s.dev.p);
    if (s.done) cudaEventDestroy(s.done);
  }
  if (c->fused_host.p) cudaFreeHost(c->fused_host.p);
  if (c->tpl_dev) cudaFree(c->tpl_dev);
  if (c->tpl_pinned) cudaFreeHost(c->tpl_pinned);
  if (c->tpl_event) cudaEventDestroy(c->tpl_event);
  if (c->fused_dev.p) cudaFree(c->fused_dev.p);
  for (Slot* g : c->graph_slots) { cudaFreeHost(g->host.p); cudaFree(g->dev.p); delete g; }
  if (c->scratch_dev.p) cudaFree(c->scratch_dev.p);
  if (c->spill_dev.p) cudaFree(c->spill_dev.p);
  if (c->gather_dev.p) cudaFree(c->gather_dev.p);
  if (c->guard_dev.p) cudaFree(c->guard_dev.p);
  if (c->enc_host.p) cudaFreeHost(c->enc_host.p);
  if (c->measured_dev.p) cudaFree(c->measured_dev.p);
  if (c->scratch_host.p) cudaFreeHost(c->scratch_host.p);
  if (c->stage_dev.p) cudaFree(c->stage_dev.p);
  if (c->arena_dev.p) cudaFree(c->arena_dev.p);
  cudaStreamDestroy(c->own_stream);
  if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
  if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
  for (auto& ev : c->pipe_ev) if (ev) cudaEventDestroy(ev);
  delete c;
  return B200TFS_OK;
}

int b200tfs_set_stream(b200tfs_ctx* c, void* stream) {
  if (!c) return fail(B200TFS_E_ARG, "ctx is NULL");
  if (c->capturing) return fail(B200TFS_E_ARG, "cannot change streams during graph capture");
  cudaStream_t next = stream ? (cudaStream_t)stream : c->own_stream;
  if (next == c->stream) return B200TFS_OK;
  CU(cudaSetDevice(c->device));
  // work already queued on the old stream (plan uploads, kernels reading the scratch buffers) must be visible to the new one:
  // an event edge, not a host synchronise
  cudaEvent_t e;
  CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  cudaError_t rc = cudaEventRecord(e, c->stream);
  if (rc == cudaSuccess) rc = cudaStreamWaitEvent(next, e, 0);
  cudaEventDestroy(e);
  if (rc != cudaSuccess) return fail(B200TFS_E_CUDA, "switching streams: %s", cudaGetErrorString(rc));
  c->stream = next;
  return B200TFS_OK;
}

int b200tfs_sync(b200tfs_ctx* c) {
  if (!c) return fail(B200TFS_E_ARG, "ctx is NULL");
  if (c->capturing) return fail(B200TFS_E_ARG, "cannot synchronise during graph capture");
  CU(cudaStreamSynchronize(c->stream));
  return B200TFS_OK;
}
void* b200tfs_stream(b200tfs_ctx* c) { return c ? (void*)c->stream : nullptr; }
int b200tfs_kernel_launches(b200tfs_ctx* c, uint64_t* count) {
  if (!c || !count) return fail(B200TFS_E_ARG, "NULL argument");
  *count = c->launches;
  return B200TFS_OK;
}

// ---- memory / events ------------------------------------------------------------------------
int b200tfs_malloc(b200tfs_ctx* c, uint64_t bytes, void** dptr) {
  if (!c || !dptr) return fail(B200TFS_E_ARG, "NULL argument");
  CU(cudaSetDevice(c->device));
  CU(cudaMalloc(dptr, bytes ? bytes : 1));
  return B200TFS_OK;
}
int b200tfs_free(b200tfs_ctx* c, void* dptr) {
  if (!c) return fail(B200TFS_E_ARG, "ctx is NULL");
  CU(cudaSetDevice(c->device));
  CU(cudaFree(dptr));
  return B200TFS_OK;
}
int b200tfs_host_alloc(uint64_t bytes, void** hptr) {
  if (!hptr) return fail(B200TFS_E_ARG, "hptr is NULL");
  CU(cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocPortable));
  return B200TFS_OK;
}
int b200tfs_host_free(void* hptr) { CU(cudaFreeHost(hptr)); return B200TFS_OK; }
int b200tfs_memcpy_h2d(b200tfs_ctx* c, void* d, const void* h, uint64_t n) {
  if (!c) return fail(B200TFS_E_ARG, "ctx is NULL");
  if (n) CU(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, c->stream));
  return B200TFS_OK;
}
int b200tfs_memcpy_d2h(b200tfs_ctx* c, void* h, const void* d, uint64_t n) {
  if (!c) return fail(B200TFS_E_ARG, "ctx is NULL");
  if (n) CU(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, c->stream));
  return B200TFS_OK;
}
int b200tfs_memcpy_d2d(b200tfs_ctx* c, void* d, const void* s, uint64_t n) {
  if (!c) return fail(B200TFS_E_ARG, "ctx is NULL");
  if (n) CU(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, c->stream));
  return B200TFS_OK;
}
int b200tfs_memset(b200tfs_ctx* c, void* d, int v, uint64_t n) {
  if (!c) return fail(B200TFS_E_ARG, "ctx is NULL");
  if (n) CU(cudaMemsetAsync(d, v, n, c->stream));
  return B200TFS_OK;
}
int b200tfs_event_create(void** ev) {
  if (!ev) return fail(B200TFS_E_ARG, "ev is NULL");
  cudaEvent_t e;
  CU(cudaEventCreate(&e));
  *ev = e;
  return B200TFS_OK;
}
int b200tfs_event_destroy(void* ev) { CU(cudaEventDestroy((cudaEvent_t)ev)); return B200TFS_OK; }
int b200tfs_event_record(b200tfs_ctx* c, void* ev) {
  if (!c) return fail(B200TFS_E_ARG, "ctx is NULL");
  CU(cudaEventRecord((cudaEvent_t)ev, c->stream));
  return B200TFS_OK;
}
int b200tfs_event_sync(void* ev) { CU(cudaEventSynchronize((cudaEvent_t)ev)); return B200TFS_OK; }
int b200tfs_event_elapsed_ms(void* a, void* b, float* ms) {
  if (!ms) return fail(B200TFS_E_ARG, "ms is NULL");
  CU(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
  return B200TFS_OK;
}

// ---- dtype table -------------------------------------------------------------------------------
int b200tfs_dtype_size(int32_t dt) { return (int)dtype_info(dt).elem_size; }
int b200tfs_dtype_field(int32_t dt) { return (int)dtype_info(dt).field; }
int b200tfs_cast_supported(int32_t src, int32_t wire) {
  if (src == wire) return dtype_info(src).kind != VK_NONE && dtype_info(src).kind != VK_STRING;
  return (wire == DT_FLOAT && (src == DT_HALF || src == DT_BFLOAT16)) ? 1 : 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// wire-size arithmetic and header bytes
// ------------------------------------------------------------------------------------------------
static bool request_needs_deferred(const b200tfs_request& r);                  // varint_host.inc
static int deferred_slot_bound(const b200tfs_request& r, uint64_t* bound);     // varint_host.inc

namespace {

constexpr uint64_t kProtoLimit = 0x7FFFFFFFull;  // protobuf's 2 GiB message limit

struct TensorLayout {
  uint64_t n_elems = 0;
  uint64_t payload_len = 0;  // bytes of the values field body on the wire (0: field omitted)
  uint64_t header_len = 0;   // bytes before the payload
  uint32_t op = OP_COPY;     // MoveOp for fixed-width payloads
  bool varint = false;       // payload produced by the varint kernels
  uint32_t field = 0;        // field number the values go into
  uint64_t shape_len = 0;    // bytes of the TensorShapeProto body
  DtypeInfo src_info{}, wire_info{};
};

// the header_len bytes in front of a tensor's payload, written at w: 08 vi(dtype) 12 vi(shape_len)
// {12 vi(dim_len) [08 vi(size)]}* [tag vi(payload_len)]
uint8_t* write_tensor_header(const b200tfs_tensor& t, const struct TensorLayout& L, uint8_t* w) {
  *w++ = 0x08; w += put_varint(w, (uint64_t)(uint32_t)t.wire_dtype);
  *w++ = 0x12; w += put_varint(w, L.shape_len);
  for (int i = 0; i < t.rank; ++i) {
    const uint64_t d = (uint64_t)t.dims[i];
    *w++ = 0x12;
    if (d) { *w++ = (uint8_t)(1 + varint_len(d)); *w++ = 0x08; w += put_varint(w, d); }
    else *w++ = 0x00;  // Dim(size=0) is an empty sub-message (Q2)
  }
  if (L.payload_len) { w += put_varint(w, tag_of(L.field, WT_LEN)); w += put_varint(w, L.payload_len); }
  return w;
}

int tensor_layout(const b200tfs_tensor& t, TensorLayout* L, std::vector<uint8_t>* hdr) {
  *L = TensorLayout{};   // layouts are reused from request to request (request_layout): no field may survive
  if (t.flags & B200TFS_F_PRESERIALIZED) {  // an already serialised TensorProto: all payload, no header
    if (t.packed_len > kProtoLimit) return fail(B200TFS_E_TOOBIG, "serialised TensorProto exceeds 2 GiB");
    L->n_elems = t.packed_len; L->payload_len = t.packed_len; L->header_len = 0; L->op = OP_COPY; L->varint = false;
    return B200TFS_OK;
  }
  if (t.rank < 0 || t.rank > 254) return fail(B200TFS_E_SHAPE, "rank %d outside [0, 254]", t.rank);
  if (t.rank && !t.dims) return fail(B200TFS_E_ARG, "dims is NULL");
  L->src_info = dtype_info(t.src_dtype);
  L->wire_info = dtype_info(t.wire_dtype);
  if (L->src_info.kind == VK_NONE || L->wire_info.kind == VK_NONE)
    return fail(B200TFS_E_DTYPE, "dtype %d -> %d is not in the TensorProto table", t.src_dtype, t.wire_dtype);
  if (L->src_info.kind == VK_STRING || L->wire_info.kind == VK_STRING)
    return fail(B200TFS_E_DTYPE, "DT_STRING tensors are assembled on the host (string_val is not a device payload)");
  if (!b200tfs_cast_supported(t.src_dtype, t.wire_dtype))
    return fail(B200TFS_E_DTYPE, "cast DT %d -> DT %d is not supported", t.src_dtype, t.wire_dtype);
  uint64_t n = 1;
  for (int i = 0; i < t.rank; ++i) {
    int64_t d = t.dims[i];
    if (d < 0) return fail(B200TFS_E_SHAPE, "negative dim %lld", (long long)d);
    if (d && n > kProtoLimit * 16 / (uint64_t)d) return fail(B200TFS_E_TOOBIG, "tensor too large for one protobuf message");
    n *= (uint64_t)d;
  }
  L->n_elems = n;
  const bool content = (t.flags & B200TFS_F_TENSOR_CONTENT) != 0;
  const bool cast = t.src_dtype != t.wire_dtype;
  uint32_t field;
  if (content) {
    field = F_CONTENT;
    L->payload_len = n * L->wire_info.elem_size;
    L->op = cast ? (t.src_dtype == DT_HALF ? OP_H2F : OP_B2F) : OP_COPY;
  } else {
    field = L->wire_info.field;
    switch (L->wire_info.kind) {
      case VK_FIXED:
        L->payload_len = n * L->wire_info.elem_size;
        if (cast) L->op = (t.src_dtype == DT_HALF) ? OP_H2F : OP_B2F;
        else L->op = (t.wire_dtype == DT_FLOAT && !(t.flags & B200TFS_F_KEEP_SNAN)) ? OP_QUIET_SRC : OP_COPY;
        break;
      case VK_BOOL:
        L->payload_len = n;
        L->op = OP_BOOL;
        break;
      default:  // VK_VARINT
        L->varint = true;
        if (n && t.packed_len == 0)
          return fail(B200TFS_E_ARG, "varint dtype %d: packed_len not set, call b200tfs_measure first", t.wire_dtype);
        L->payload_len = n ? t.packed_len : 0;
        break;
    }
  }
  if (L->payload_len > kProtoLimit) return fail(B200TFS_E_TOOBIG, "payload of %llu bytes exceeds protobuf's 2 GiB limit", (unsigned long long)L->payload_len);
  uint64_t shape_len = 0;
  for (int i = 0; i < t.rank; ++i) shape_len += 2 + (t.dims[i] ? 1 + varint_len((uint64_t)t.dims[i]) : 0);
  uint64_t hl = 1 + varint_len((uint64_t)(uint32_t)t.wire_dtype) + 1 + varint_len(shape_len) + shape_len;
  if (L->payload_len) hl += varint_len(tag_of(field, WT_LEN)) + varint_len(L->payload_len);
  L->header_len = hl; L->field = field; L->shape_len = shape_len;
  if (hdr) {   // one resize, then raw writes
    const size_t base = hdr->size();
    hdr->resize(base + hl);
    if ((uint64_t)(write_tensor_header(t, *L, hdr->data() + base) - (hdr->data() + base)) != hl)
      return fail(B200TFS_E_ARG, "internal: header length mismatch");
  }
  return B200TFS_OK;
}

// key order.  UPB: bytewise on the common prefix; on a tie the LONGER key first.
int order_keys(int32_t n, const char* const* keys, const int64_t* lens, int32_t order, int32_t* perm) {
  for (int i = 0; i < n; ++i) perm[i] = i;
  if (order == B200TFS_ORDER_GIVEN) return B200TFS_OK;
  if (order != B200TFS_ORDER_UPB && order != B200TFS_ORDER_BYTES) return fail(B200TFS_E_ARG, "unknown key order %d", order);
  std::stable_sort(perm, perm + n, [&](int a, int b) {
    size_t la = (size_t)lens[a], lb = (size_t)lens[b];
    int c = memcmp(keys[a], keys[b], std::min(la, lb));
    if (c) return c < 0;
    if (la == lb) return false;
    return order == B200TFS_ORDER_UPB ? la > lb : la < lb;
  });
  return B200TFS_OK;
}

struct VarJob {  // one varint-packed payload (handled by the varint kernels after move_kernel)
  const uint8_t* src;
  uint8_t* dst;
  uint64_t n_elems;
  uint64_t packed_len;
  int32_t src_dtype;
};

// Accumulates the pieces of a batch of records into a plan image.
struct PlanBuilder {
  std::vector<MoveItem> items;
  std::vector<SmallItem> smalls;
  std::vector<uint8_t> blob;
  std::vector<VarJob> varjobs;
  uint64_t large_bytes = 0;
  const uint32_t* guard = nullptr;   // move_guarded_kernel: item i is stored only if guard[i / guard_div] != 0
  uint32_t guard_div = 1;
  bool independent = false;          // PlanHeader::independent

  void header(uint8_t* dst, size_t blob_off, size_t n) {
    // split long headers so one warp never walks more than kSmallMax bytes
    while (n) {
      uint32_t k = (uint32_t)std::min<size_t>(n, kSmallMax);
      smalls.push_back(SmallItem{(uint64_t)blob_off, dst, k, OP_COPY | OP_FLAG_BLOB, 0, 0});
      dst += k; blob_off += k; n -= k;
    }
  }
  // glen / gstride != 0: the source is a row of pieces (b200tfs_run with count > 1), gathered byte-exact
  void payload(const uint8_t* src, uint8_t* dst, uint64_t n_out, uint32_t op, uint32_t glen = 0, uint32_t gstride = 0) {
    if (!n_out) return;
    if (n_out <= kSmallMax) smalls.push_back(SmallItem{(uint64_t)(uintptr_t)src, dst, (uint32_t)n_out, op, glen, gstride});
    else { items.push_back(MoveItem{src, dst, n_out, op, 0, glen, gstride}); large_bytes += n_out; }
  }
};

uint32_t pick_vec_per_tile(const b200tfs_ctx* c, uint64_t large_bytes, uint64_t max_tile = 65536) {
  uint64_t tile = c->tile_bytes_override;
  if (!tile) {
    // 32 KB per CTA: both vector paths keep all of it in flight at once (8 x 16 B per thread), which
    // measured best both for one 4 MiB tensor alone and for many overlapping launches
    uint64_t target_tiles = (uint64_t)c->sm_count * 8;
    tile = (large_bytes + target_tiles - 1) / target_tiles;
    tile = (tile + 32767) & ~32767ull;
    tile = std::min<uint64_t>(std::max<uint64_t>(tile, 32768), max_tile);
  }
  tile = std::max<uint64_t>(tile & ~31ull, 32);
  return (uint32_t)(tile / 16);
}

// Serialise the plan image (blob offsets inside SmallItems are rebased onto the image).  Small images travel in the kernel
// parameters unless `force_dev`: then - and for large images - the image is uploaded and *plan_dev points at it.
struct BuiltPlan {
  PlanHeader ph{};
  uint8_t inline_buf[kInlinePlanBytes];
  const uint8_t* host_img = nullptr;
  uint8_t* plan_dev = nullptr;   // nullptr: inline
  uint64_t image = 0;
  Slot* slot = nullptr;
};

int build_plan(b200tfs_ctx* c, PlanBuilder& pb, bool force_dev, BuiltPlan* bp) {
  const uint32_t vpt = pick_vec_per_tile(c, pb.large_bytes);
  // tiles
  uint64_t n_tiles = 0;
  uint32_t uniform = 0;
  bool is_uniform = !pb.items.empty();
  for (auto& it : pb.items) {
    uint64_t t = tiles_for_host(it.n_out, vpt);
    if (n_tiles + t > 0x7FFFFFFFull) return fail(B200TFS_E_TOOBIG, "batch needs more than 2^31 tiles");
    it.n_tiles = (uint32_t)t;
    if (&it == &pb.items[0]) uniform = (uint32_t)t; else if (t != uniform) is_uniform = false;
    n_tiles += t;
  }
  if (!is_uniform) uniform = 0;
  PlanHeader& ph = bp->ph;
  ph = PlanHeader{};
  ph.n_items = (uint32_t)pb.items.size();
  ph.n_tiles = (uint32_t)n_tiles;
  ph.n_small = (uint32_t)pb.smalls.size();
  ph.uniform_tpi = uniform;
  ph.vec_per_tile = vpt;
  ph.guard = pb.guard; ph.guard_div = pb.guard_div ? pb.guard_div : 1u;
  ph.independent = pb.independent ? 1u : 0u;
  uint64_t off = (sizeof(PlanHeader) + 15) & ~15ull;
  ph.off_items = (uint32_t)off; off += pb.items.size() * sizeof(MoveItem);
  ph.off_tiles = (uint32_t)off; if (!uniform) off += n_tiles * sizeof(TileRef);
  off = (off + 15) & ~15ull;
  ph.off_small = (uint32_t)off; off += pb.smalls.size() * sizeof(SmallItem);
  const uint64_t off_blob = off;
  off += pb.blob.size();
  const uint64_t image = (off + 15) & ~15ull;
  if (image > 0xFFFFFFFFull) return fail(B200TFS_E_TOOBIG, "plan image larger than 4 GiB");
  bp->image = image;
  uint8_t* img;
  const bool inl = !force_dev && image <= kInlinePlanBytes;
  if (inl) img = bp->inline_buf;
  else {
    int rc = claim_slot(c, image, &bp->slot);
    if (rc) return rc;
    img = (uint8_t*)bp->slot->host.p;
  }
  bp->host_img = img;
  memcpy(img, &ph, sizeof ph);
  if (!pb.items.empty()) memcpy(img + ph.off_items, pb.items.data(), pb.items.size() * sizeof(MoveItem));
  if (!uniform && n_tiles) {
    TileRef* tr = (TileRef*)(img + ph.off_tiles);
    uint32_t g = 0;
    for (uint32_t i = 0; i < pb.items.size(); ++i)
      for (uint32_t t = 0; t < pb.items[i].n_tiles; ++t) tr[g++] = TileRef{i, t};
  }
  SmallItem* sm = (SmallItem*)(img + ph.off_small);
  for (size_t i = 0; i < pb.smalls.size(); ++i) {
    sm[i] = pb.smalls[i];
    if (sm[i].op & OP_FLAG_BLOB) sm[i].src += off_blob;
  }
  if (!pb.blob.empty()) memcpy(img + off_blob, pb.blob.data(), pb.blob.size());
  if (!inl) {
    int rc = upload_slot(c, bp->slot, image);
    if (rc) return rc;
    bp->plan_dev = (uint8_t*)bp->slot->dev.p;
  }
  return B200TFS_OK;
}

int launch_built_plan(b200tfs_ctx* c, BuiltPlan& bp) {
  CU(launch_move(bp.plan_dev, bp.host_img, (uint32_t)bp.image, bp.ph.n_tiles, bp.ph.n_small, c->stream));
  c->launches += 1;
  if (bp.slot && bp.slot->done && !c->capturing) { CU(cudaEventRecord(bp.slot->done, c->stream)); bp.slot->pending = true; }   // the kernel still reads the image
  return B200TFS_OK;
}

int launch_plan(b200tfs_ctx* c, PlanBuilder& pb) {
  if (pb.items.empty() && pb.smalls.empty()) return B200TFS_OK;
  BuiltPlan bp;
  int rc = build_plan(c, pb, false, &bp);
  if (rc) return rc;
  return launch_built_plan(c, bp);
}

// record placement: slots start 256-byte aligned, then padded so the record's largest payload
// starts 128-byte aligned (the vector path then needs no realignment for it)
inline uint64_t place_record(uint64_t cursor, uint64_t largest_payload_off) {
  uint64_t slot = (cursor + 255) & ~255ull;
  uint64_t pad = (128 - (largest_payload_off & 127)) & 127;
  return slot + pad;
}

// Lay one TensorProto at arena+rec_off; appends its pieces to the plan.  `hdr_prefix` bytes (entry
// framing) have already been appended to pb.blob starting at blob_mark and precede the proto header.
int plan_tensor(const b200tfs_tensor& t, const TensorLayout& L, uint8_t* dst_payload, PlanBuilder& pb) {
  if (!L.payload_len) return B200TFS_OK;
  if (L.varint) {
    pb.varjobs.push_back(VarJob{(const uint8_t*)t.data, dst_payload, L.n_elems, L.payload_len, t.src_dtype});
    return B200TFS_OK;
  }
  pb.payload((const uint8_t*)t.data, dst_payload, L.payload_len, L.op);
  return B200TFS_OK;
}

int run_varjobs(b200tfs_ctx* c, PlanBuilder& pb);  // varint.cpp part below

struct RequestLayout {
  std::vector<TensorLayout> tl;
  std::vector<int32_t> perm;
  std::vector<uint64_t> tp_len, entry_len;
  uint64_t spec_len = 0, version_len = 0, total = 0;
  uint64_t prefix = 0;        // bytes in front of the message: gRPC's 5-byte frame header when asked for
  uint64_t largest_off = 0;
  std::vector<uint64_t> payload_off;
};

int request_layout(const b200tfs_request& r, RequestLayout* R) {
  if (r.n_inputs < 0) return fail(B200TFS_E_ARG, "n_inputs < 0");
  if (r.n_inputs && !r.inputs) return fail(B200TFS_E_ARG, "inputs is NULL");
  if (r.model_name_len < 0 || (r.model_name_len && !r.model_name)) return fail(B200TFS_E_ARG, "bad model_name");
  const int n = r.n_inputs;
  R->tl.resize(n); R->perm.resize(n); R->tp_len.resize(n); R->entry_len.resize(n); R->payload_off.resize(n);
  // (no heap traffic for the usual handful of inputs: this runs once per request of a batch)
  constexpr int kInline = 16;
  const char* keys_in[kInline];
  int64_t lens_in[kInline];
  std::vector<const char*> keys_v;
  std::vector<int64_t> lens_v;
  const char** keys = keys_in;
  int64_t* lens = lens_in;
  if (n > kInline) { keys_v.resize(n); lens_v.resize(n); keys = keys_v.data(); lens = lens_v.data(); }
  for (int i = 0; i < n; ++i) {
    if (r.inputs[i].key_len < 0 || (r.inputs[i].key_len && !r.inputs[i].key)) return fail(B200TFS_E_ARG, "bad key on input %d", i);
    keys[i] = r.inputs[i].key ? r.inputs[i].key : "";
    lens[i] = r.inputs[i].key_len;
  }
  int rc = B200TFS_OK;
  if (n == 1 && (r.order == B200TFS_ORDER_GIVEN || r.order == B200TFS_ORDER_UPB || r.order == B200TFS_ORDER_BYTES)) R->perm[0] = 0;
  else rc = order_keys(n, keys, lens, r.order, R->perm.data());
  if (rc) return rc;
  // model_spec{ 0A vi name  [12 vi {08 vi(version)}] }
  uint64_t spec = 0;
  if (r.model_name_len) spec += 1 + varint_len((uint64_t)r.model_name_len) + (uint64_t)r.model_name_len;
  R->version_len = 0;
  if (r.has_version) {
    R->version_len = r.version ? 1 + varint_len((uint64_t)r.version) : 0;
    spec += 2 + R->version_len;
  }
  R->spec_len = spec;
  if (r.flags & ~B200TFS_RF_GRPC_FRAME) return fail(B200TFS_E_ARG, "unknown request flags 0x%x", (unsigned)r.flags);
  R->prefix = (r.flags & B200TFS_RF_GRPC_FRAME) ? 5 : 0;
  uint64_t total = R->prefix + 1 + varint_len(spec) + spec;
  uint64_t largest = 0;
  R->largest_off = 0;
  for (int j = 0; j < n; ++j) {
    const b200tfs_tensor& t = r.inputs[R->perm[j]];
    TensorLayout& L = R->tl[j];
    if ((rc = tensor_layout(t, &L, nullptr))) return rc;
    uint64_t tp = L.header_len + L.payload_len;
    if (tp > kProtoLimit) return fail(B200TFS_E_TOOBIG, "TensorProto of %llu bytes exceeds 2 GiB", (unsigned long long)tp);
    R->tp_len[j] = tp;
    uint64_t el = 1 + varint_len((uint64_t)t.key_len) + (uint64_t)t.key_len + 1 + varint_len(tp) + tp;
    R->entry_len[j] = el;
    uint64_t entry_hdr = 1 + varint_len(el) + 1 + varint_len((uint64_t)t.key_len) + (uint64_t)t.key_len + 1 + varint_len(tp);
    uint64_t payload_off = total + entry_hdr + L.header_len;
    R->payload_off[j] = payload_off;
    if (L.payload_len > largest) { largest = L.payload_len; R->largest_off = payload_off; }
    total += 1 + varint_len(el) + el;
  }
  if (total - R->prefix > kProtoLimit)
    return fail(B200TFS_E_TOOBIG, "PredictRequest of %llu bytes exceeds protobuf's 2 GiB limit", (unsigned long long)(total - R->prefix));
  R->total = total;
  return B200TFS_OK;
}

// append the wire bytes of request r to the plan, record at arena + rec_off.  Every framing byte of the request is
// written into the blob with one resize and raw stores (this runs once per request of a batch); the tensor layouts are
// the ones request_layout computed.
int plan_request(const b200tfs_request& r, const RequestLayout& R, uint8_t* rec, PlanBuilder& pb) {
  uint64_t frame = R.total;
  for (int j = 0; j < r.n_inputs; ++j) frame -= R.tl[j].payload_len;
  const size_t base = pb.blob.size();
  pb.blob.resize(base + frame);
  uint8_t* const w0 = pb.blob.data() + base;
  uint8_t* w = w0;
  size_t mark = base;     // blob offset where the pending header run starts
  uint8_t* cursor = rec;  // where that run will land
  if (R.prefix) {         // gRPC length-prefixed message: compressed-flag 0, big-endian uint32 length
    const uint64_t m = R.total - R.prefix;
    *w++ = 0; *w++ = (uint8_t)(m >> 24); *w++ = (uint8_t)(m >> 16); *w++ = (uint8_t)(m >> 8); *w++ = (uint8_t)m;
  }
  *w++ = 0x0A; w += put_varint(w, R.spec_len);
  if (r.model_name_len) {
    *w++ = 0x0A; w += put_varint(w, (uint64_t)r.model_name_len);
    memcpy(w, r.model_name, (size_t)r.model_name_len); w += r.model_name_len;
  }
  if (r.has_version) {
    *w++ = 0x12; *w++ = (uint8_t)R.version_len;
    if (r.version) { *w++ = 0x08; w += put_varint(w, (uint64_t)r.version); }
  }
  for (int j = 0; j < r.n_inputs; ++j) {
    const b200tfs_tensor& t = r.inputs[R.perm[j]];
    const TensorLayout& L = R.tl[j];
    *w++ = 0x12; w += put_varint(w, R.entry_len[j]);
    *w++ = 0x0A; w += put_varint(w, (uint64_t)t.key_len);
    if (t.key_len) { memcpy(w, t.key, (size_t)t.key_len); w += t.key_len; }
    *w++ = 0x12; w += put_varint(w, R.tp_len[j]);
    if (!(t.flags & B200TFS_F_PRESERIALIZED)) w = write_tensor_header(t, L, w);
    if (L.payload_len) {
      const size_t at = base + (size_t)(w - w0), run = at - mark;
      pb.header(cursor, mark, run);
      cursor += run;
      int rc = plan_tensor(t, L, cursor, pb);
      if (rc) return rc;
      cursor += L.payload_len;
      mark = at;
    }
  }
  const size_t end = base + (size_t)(w - w0), run = end - mark;
  if (run) { pb.header(cursor, mark, run); cursor += run; }
  if ((uint64_t)(w - w0) != frame || (uint64_t)(cursor - rec) != R.total) return fail(B200TFS_E_ARG, "internal: request length mismatch");
  return B200TFS_OK;
}

}  // namespace

extern "C" {

int b200tfs_tensor_proto_size(const b200tfs_tensor* t, uint64_t* header_len, uint64_t* total_len) {
  if (!t) return fail(B200TFS_E_ARG, "tensor is NULL");
  TensorLayout L;
  int rc = tensor_layout(*t, &L, nullptr);
  if (rc) return rc;
  if (L.header_len + L.payload_len > kProtoLimit) return fail(B200TFS_E_TOOBIG, "TensorProto exceeds 2 GiB");
  if (header_len) *header_len = L.header_len;
  if (total_len) *total_len = L.header_len + L.payload_len;
  return B200TFS_OK;
}

int b200tfs_request_size(const b200tfs_request* r, uint64_t* total_len) {
  if (!r) return fail(B200TFS_E_ARG, "request is NULL");
  RequestLayout R;
  int rc = request_layout(*r, &R);
  if (rc) return rc;
  if (total_len) *total_len = R.total;
  return B200TFS_OK;
}

int b200tfs_tensor_proto_header(const b200tfs_tensor* t, void* buf, uint64_t cap, uint64_t* len) {
  if (!t || !len) return fail(B200TFS_E_ARG, "NULL argument");
  TensorLayout L;
  std::vector<uint8_t> h;
  int rc = tensor_layout(*t, &L, &h);
  if (rc) return rc;
  *len = h.size();
  if (h.size() > cap) return fail(B200TFS_E_SIZE, "header needs %zu bytes", h.size());
  if (!h.empty()) memcpy(buf, h.data(), h.size());
  return B200TFS_OK;
}

int b200tfs_request_frame(const b200tfs_request* r, void* buf, uint64_t cap, uint64_t* frame_len, uint64_t* payload_off,
                          uint64_t* payload_len, int32_t* perm) {
  if (!r || !frame_len) return fail(B200TFS_E_ARG, "NULL argument");
  RequestLayout R;
  int rc = request_layout(*r, &R);
  if (rc) return rc;
  PlanBuilder pb;  // planned against a NULL arena: only the blob (frame bytes in wire order) is used
  if ((rc = plan_request(*r, R, nullptr, pb))) return rc;
  *frame_len = pb.blob.size();
  for (int j = 0; j < r->n_inputs; ++j) {
    if (payload_off) payload_off[j] = R.payload_off[j];
    if (payload_len) payload_len[j] = R.tl[j].payload_len;
    if (perm) perm[j] = R.perm[j];
  }
  if (pb.blob.size() > cap) return fail(B200TFS_E_SIZE, "frame needs %zu bytes", pb.blob.size());
  if (!pb.blob.empty()) memcpy(buf, pb.blob.data(), pb.blob.size());
  return B200TFS_OK;
}

int b200tfs_order_keys(int32_t n, const char* const* keys, const int64_t* key_lens, int32_t order, int32_t* perm) {
  if (n < 0 || (n && (!keys || !key_lens || !perm))) return fail(B200TFS_E_ARG, "bad arguments");
  return order_keys(n, keys, key_lens, order, perm);
}

int b200tfs_tensor_arena_size(int32_t n, const b200tfs_tensor* tensors, uint64_t* bytes) {
  if (n < 0 || (n && !tensors) || !bytes) return fail(B200TFS_E_ARG, "bad arguments");
  uint64_t cursor = 0;
  for (int i = 0; i < n; ++i) {
    TensorLayout L;
    int rc = tensor_layout(tensors[i], &L, nullptr);
    if (rc) return rc;
    cursor = place_record(cursor, L.header_len) + L.header_len + L.payload_len;
  }
  *bytes = (cursor + 255) & ~255ull;
  return B200TFS_OK;
}

int b200tfs_request_arena_size(int32_t n, const b200tfs_request* reqs, uint64_t* bytes) {
  if (n < 0 || (n && !reqs) || !bytes) return fail(B200TFS_E_ARG, "bad arguments");
  // a batch with an unmeasured packed-varint input (packed_len == 0) is sized for b200tfs_encode_requests_async: every record
  // gets a slot for its worst case; measured batches are sized exactly, as b200tfs_encode_requests lays them out
  bool deferred = false;
  for (int i = 0; i < n && !deferred; ++i) deferred = request_needs_deferred(reqs[i]);
  uint64_t cursor = 0;
  if (deferred) {
    for (int i = 0; i < n; ++i) {
      uint64_t bound = 0;
      int rc = deferred_slot_bound(reqs[i], &bound);
      if (rc) return rc;
      cursor = ((cursor + 255) & ~255ull) + bound + 128;
    }
    *bytes = (cursor + 255) & ~255ull;
    return B200TFS_OK;
  }
  RequestLayout R;
  for (int i = 0; i < n; ++i) {
    int rc = request_layout(reqs[i], &R);
    if (rc) return rc;
    cursor = place_record(cursor, R.largest_off) + R.total;
  }
  *bytes = (cursor + 255) & ~255ull;
  return B200TFS_OK;
}

int b200tfs_encode_tensor_protos(b200tfs_ctx* c, int32_t n, const b200tfs_tensor* tensors, void* arena_dev, uint64_t arena_cap,
                                 uint64_t* rec_off, uint64_t* rec_len) {
  if (!c || n < 0 || (n && (!tensors || !arena_dev || !rec_off || !rec_len))) return fail(B200TFS_E_ARG, "bad arguments");
  if ((uintptr_t)arena_dev & 255) return fail(B200TFS_E_ARG, "arena must be 256-byte aligned");
  CU(cudaSetDevice(c->device));
  PlanBuilder pb;
  uint64_t cursor = 0;
  for (int i = 0; i < n; ++i) {
    TensorLayout L;
    size_t mark = pb.blob.size();
    int rc = tensor_layout(tensors[i], &L, &pb.blob);
    if (rc) return rc;
    uint64_t total = L.header_len + L.payload_len;
    if (total > kProtoLimit) return fail(B200TFS_E_TOOBIG, "TensorProto exceeds 2 GiB");
    if (L.payload_len && !tensors[i].data) return fail(B200TFS_E_ARG, "tensor %d: data pointer is NULL", i);
    uint64_t off = place_record(cursor, L.header_len);
    if (off + total > arena_cap) return fail(B200TFS_E_SIZE, "arena too small: need %llu bytes", (unsigned long long)(off + total));
    uint8_t* rec = (uint8_t*)arena_dev + off;
    pb.header(rec, mark, L.header_len);
    if ((rc = plan_tensor(tensors[i], L, rec + L.header_len, pb))) return rc;
    rec_off[i] = off; rec_len[i] = total;
    cursor = off + total;
  }
  int rc = launch_plan(c, pb);
  if (rc) return rc;
  return run_varjobs(c, pb);
}

// lay the batch out in the arena and collect its pieces (b200tfs_encode_requests launches them at once, the pipelined host path in slices)
static int plan_requests(int32_t n, const b200tfs_request* reqs, void* arena_dev, uint64_t arena_cap, uint64_t* rec_off, uint64_t* rec_len,
                         PlanBuilder& pb) {
  if (n > 16) {   // a batch: one allocation per table instead of a doubling series
    size_t inputs = 0;
    for (int i = 0; i < n; ++i) inputs += (size_t)std::max(reqs[i].n_inputs, 0);
    pb.items.reserve(inputs); pb.smalls.reserve(inputs + (size_t)n); pb.blob.reserve(64 * inputs + 48 * (size_t)n);
  }
  RequestLayout R;
  uint64_t cursor = 0;
  for (int i = 0; i < n; ++i) {
    int rc = request_layout(reqs[i], &R);
    if (rc) return rc;
    for (int j = 0; j < reqs[i].n_inputs; ++j)
      if (R.tl[j].payload_len && !reqs[i].inputs[R.perm[j]].data) return fail(B200TFS_E_ARG, "request %d: tensor data pointer is NULL", i);
    uint64_t off = place_record(cursor, R.largest_off);
    if (off + R.total > arena_cap) return fail(B200TFS_E_SIZE, "arena too small: need %llu bytes", (unsigned long long)(off + R.total));
    if ((rc = plan_request(reqs[i], R, (uint8_t*)arena_dev + off, pb))) return rc;
    rec_off[i] = off; rec_len[i] = R.total;
    cursor = off + R.total;
  }
  return B200TFS_OK;
}

int b200tfs_encode_requests(b200tfs_ctx* c, int32_t n, const b200tfs_request* reqs, void* arena_dev, uint64_t arena_cap,
                            uint64_t* rec_off, uint64_t* rec_len) {
  if (!c || n < 0 || (n && (!reqs || !arena_dev || !rec_off || !rec_len))) return fail(B200TFS_E_ARG, "bad arguments");
  if ((uintptr_t)arena_dev & 255) return fail(B200TFS_E_ARG, "arena must be 256-byte aligned");
  CU(cudaSetDevice(c->device));
  PlanBuilder pb;
  int rc = plan_requests(n, reqs, arena_dev, arena_cap, rec_off, rec_len, pb);
  if (rc) return rc;
  if ((rc = launch_plan(c, pb))) return rc;
  return run_varjobs(c, pb);
}

// ---- decode ------------------------------------------------------------------------------------
static int parse_common(b200tfs_ctx* c, const void* arena_dev, int32_t n, const uint64_t* rec_off, const uint64_t* rec_len,
                        int32_t max_outputs, bool bare, b200tfs_output* outs, int32_t* n_outs, b200tfs_model_spec* specs,
                        int32_t* rec_status) {
  if (!c || n < 0 || (n && (!arena_dev || !rec_off || !rec_len || !outs || !rec_status))) return fail(B200TFS_E_ARG, "bad arguments");
  if (!bare && (max_outputs <= 0 || !n_outs || !specs)) return fail(B200TFS_E_ARG, "bad arguments");
  if (n == 0) return B200TFS_OK;
  if (c->capturing) return fail(B200TFS_E_ARG, "the two-phase parse synchronises and cannot be captured: use b200tfs_decode_responses");
  CU(cudaSetDevice(c->device));
  if (bare) max_outputs = 1;
  const uint64_t stride = bare ? 1 : (uint64_t)max_outputs + 1;  // responses: one scratch slot per record
  const uint64_t b_off = 0, b_len = b_off + 8ull * n, b_outs = (b_len + 8ull * n + 15) & ~15ull;
  const uint64_t b_nouts = b_outs + sizeof(b200tfs_output) * (uint64_t)n * stride;
  const uint64_t b_specs = (b_nouts + 4ull * n + 15) & ~15ull;
  const uint64_t b_status = b_specs + sizeof(b200tfs_model_spec) * (uint64_t)n;
  const uint64_t b_spill = (b_status + 4ull * n + 15) & ~15ull;      // uint32 spill_used[n]
  const uint64_t total = (b_spill + 4ull * n + 15) & ~15ull;
  int rc;
  if ((rc = grow_dev(c, c->scratch_dev, total))) return rc;
  if ((rc = grow_host(c, c->scratch_host, total))) return rc;
  uint8_t* h = (uint8_t*)c->scratch_host.p;
  uint8_t* d = (uint8_t*)c->scratch_dev.p;
  memcpy(h + b_off, rec_off, 8ull * n);
  memcpy(h + b_len, rec_len, 8ull * n);
  CU(cudaMemcpyAsync(d, h, b_outs, cudaMemcpyHostToDevice, c->stream));
  static_assert(sizeof(SpillEntry) == kSpillEntryBytes, "SpillEntry layout");
  // The walk runs with a spill area of spill_per_rec entries per record; a record that wants more says how many
  // (B200TFS_E_SPILL + spill_used) and the walk runs once more with that much - the count is exact, so twice at most.
  uint32_t per = c->spill_per_rec;
  c->spill_host.clear(); c->spill_last_n = 0; c->spill_last_per = 0;
  for (int attempt = 0;; ++attempt) {
    if ((rc = grow_dev(c, c->spill_dev, (uint64_t)n * per * sizeof(SpillEntry) + 16))) return rc;
    if (bare)
      CU(launch_parse_tensors((const uint8_t*)arena_dev, (const uint64_t*)(d + b_off), (const uint64_t*)(d + b_len), n,
                              (b200tfs_output*)(d + b_outs), (int32_t*)(d + b_status), c->spill_dev.p, per, (uint32_t*)(d + b_spill), c->stream));
    else
      CU(launch_parse_responses((const uint8_t*)arena_dev, (const uint64_t*)(d + b_off), (const uint64_t*)(d + b_len), n, max_outputs,
                                (b200tfs_output*)(d + b_outs), (int32_t*)(d + b_nouts), (b200tfs_model_spec*)(d + b_specs),
                                (int32_t*)(d + b_status), c->spill_dev.p, per, (uint32_t*)(d + b_spill), c->stream));
    c->launches += 1;
    CU(cudaMemcpyAsync(h + b_outs, d + b_outs, total - b_outs, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const int32_t* st = (const int32_t*)(h + b_status);
    const uint32_t* used = (const uint32_t*)(h + b_spill);
    uint32_t want = 0, any = 0;
    bool again = false;
    for (int i = 0; i < n; ++i) { any |= used[i]; if (st[i] == B200TFS_E_SPILL) { again = true; want = std::max(want, used[i]); } }
    if (again && attempt < 2) {
      if ((uint64_t)n * want * sizeof(SpillEntry) > (1ull << 32)) return fail(B200TFS_E_TOOBIG, "spill area of %u entries x %d records", want, n);
      per = want;
      continue;
    }
    if (any) {   // bring the area over: unpack and the b200tfs_output_* accessors read it on the host
      c->spill_host.resize((size_t)n * per);
      CU(cudaMemcpyAsync(c->spill_host.data(), c->spill_dev.p, (size_t)n * per * sizeof(SpillEntry), cudaMemcpyDeviceToHost, c->stream));
      CU(cudaStreamSynchronize(c->stream));
      c->spill_last_n = n; c->spill_last_per = per;
    }
    break;
  }
  memcpy(rec_status, h + b_status, 4ull * n);
  for (int i = 0; i < n; ++i) if (rec_status[i] == B200TFS_E_SPILL) rec_status[i] = B200TFS_E_NONCANONICAL;   // cannot happen: the second walk had room
  if (bare) {
    memcpy(outs, h + b_outs, sizeof(b200tfs_output) * (uint64_t)n);
  } else {
    memcpy(n_outs, h + b_nouts, 4ull * n);
    memcpy(specs, h + b_specs, sizeof(b200tfs_model_spec) * (uint64_t)n);
    for (int i = 0; i < n; ++i) {
      int k = rec_status[i] == B200TFS_OK ? n_outs[i] : 0;
      memcpy(outs + (size_t)i * max_outputs, h + b_outs + sizeof(b200tfs_output) * (uint64_t)i * stride, sizeof(b200tfs_output) * (size_t)k);
    }
  }
  return B200TFS_OK;
}

// every value run of an output, in wire order: the inline ones, then the spilled ones of its map entry and value field
static int collect_runs(const b200tfs_ctx* c, const b200tfs_output& o, std::vector<b200tfs_run>& runs) {
  runs.clear();
  const uint32_t inl = std::min<uint32_t>(o.n_inline, B200TFS_MAX_RUNS);
  for (uint32_t k = 0; k < inl; ++k) runs.push_back(o.runs[k]);
  if ((uint32_t)o.n_runs > inl) {
    if (!(o.flags & B200TFS_OF_SPILLED) || (int32_t)o.spill_rec >= c->spill_last_n || c->spill_host.empty())
      return fail(B200TFS_E_ARG, "output lists %d value runs, %u inline, but the context holds no spill area for it (another parse ran since?)",
                  o.n_runs, inl);
    const SpillEntry* e = c->spill_host.data() + (size_t)o.spill_rec * c->spill_last_per;
    for (uint32_t k = 0; k < c->spill_last_per && (int)runs.size() < o.n_runs; ++k)
      if (e[k].kind == SPILL_RUN && e[k].seq == o.spill_seq && (int32_t)e[k].run.field == o.value_field) runs.push_back(e[k].run);
    if ((int)runs.size() != o.n_runs) return fail(B200TFS_E_ARG, "spill area does not hold the %d runs this output lists", o.n_runs);
  }
  return B200TFS_OK;
}

int b200tfs_output_runs(b200tfs_ctx* c, const b200tfs_output* o, b200tfs_run* runs, int32_t cap) {
  if (!c || !o || (cap > 0 && !runs) || cap < 0) return fail(B200TFS_E_ARG, "bad arguments");
  std::vector<b200tfs_run> all;
  int rc = collect_runs(c, *o, all);
  if (rc) return rc;
  for (int32_t k = 0; k < cap && k < (int32_t)all.size(); ++k) runs[k] = all[k];
  return B200TFS_OK;
}

int b200tfs_output_dims(b200tfs_ctx* c, const b200tfs_output* o, int64_t* dims, int32_t cap) {
  if (!c || !o || (cap > 0 && !dims) || cap < 0) return fail(B200TFS_E_ARG, "bad arguments");
  int32_t k = 0;
  for (; k < cap && k < o->rank && k < B200TFS_MAX_RANK; ++k) dims[k] = o->dims[k];
  if (o->rank > B200TFS_MAX_RANK && cap > B200TFS_MAX_RANK) {
    if ((int32_t)o->spill_rec >= c->spill_last_n || c->spill_host.empty())
      return fail(B200TFS_E_ARG, "rank %d output, but the context holds no spill area for it (another parse ran since?)", o->rank);
    const SpillEntry* e = c->spill_host.data() + (size_t)o->spill_rec * c->spill_last_per;
    for (uint32_t i = 0; i < c->spill_last_per && k < cap && k < o->rank; ++i)
      if (e[i].kind == SPILL_DIM && e[i].seq == o->spill_seq) dims[k++] = (int64_t)e[i].run.off;
    if (k < std::min(cap, o->rank)) return fail(B200TFS_E_ARG, "spill area does not hold the %d dims this output lists", o->rank);
  }
  return B200TFS_OK;
}

int b200tfs_parse_responses(b200tfs_ctx* c, const void* arena_dev, int32_t n, const uint64_t* rec_off, const uint64_t* rec_len,
                            int32_t max_outputs, b200tfs_output* outs, int32_t* n_outs, b200tfs_model_spec* specs,
                            int32_t* rec_status) {
  return parse_common(c, arena_dev, n, rec_off, rec_len, max_outputs, false, outs, n_outs, specs, rec_status);
}

int b200tfs_parse_tensor_protos(b200tfs_ctx* c, const void* arena_dev, int32_t n, const uint64_t* rec_off, const uint64_t* rec_len,
                                b200tfs_output* outs, int32_t* rec_status) {
  return parse_common(c, arena_dev, n, rec_off, rec_len, 1, true, outs, nullptr, nullptr, rec_status);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// unpack
// ------------------------------------------------------------------------------------------------
namespace {

struct VarDecodeJob {
  const uint8_t* src[B200TFS_MAX_RUNS];
  uint64_t len[B200TFS_MAX_RUNS];
  int n_chunks;
  uint8_t* dst;
  uint64_t n_elems;
  int32_t dtype;
  int32_t out_index;
  bool half_as_value;
  bool pad_edge;   // fewer values than elements: repeat the last one (B200TFS_OF_PAD_EDGE)
};

int run_vardecode(b200tfs_ctx* c, std::vector<VarDecodeJob>& jobs, int32_t* status);  // below

}  // namespace

// after a synchronise: the status words the varint decoder left in pinned memory -> the caller's array
static void collect_varint_status(b200tfs_ctx* c, int32_t* status) {
  if (status)
    for (size_t i = 0; i < c->pending_status.size(); ++i) status[c->pending_status[i]] = ((const int32_t*)c->scratch_host.p)[i];
  c->pending_status.clear();
}

// wait = false: everything is enqueued, nothing synchronised; the caller synchronises and calls collect_varint_status
static int unpack_outputs_impl(b200tfs_ctx* c, const void* arena_dev, int32_t m, const b200tfs_output* outs, const uint64_t* out_rec_off,
                               void* const* dst_dev, const int32_t* dst_dtype, int32_t* status, bool wait) {
  if (!c || m < 0 || (m && (!arena_dev || !outs || !dst_dev))) return fail(B200TFS_E_ARG, "bad arguments");
  CU(cudaSetDevice(c->device));
  PlanBuilder pb;
  std::vector<VarDecodeJob> vjobs;
  struct Fill { uint8_t* dst; uint32_t elem_size; uint64_t have, n_elems; };
  std::vector<Fill> fills;   // B200TFS_OF_PAD_EDGE on fixed-width outputs: pad once the values are in place
  // value runs of every output (inline + spilled).  Packed-varint outputs whose values lie in strided runs (rows of unpacked
  // elements) or in more runs than a decode job lists are first GATHERED into one contiguous stream - the concatenation of
  // the pieces is itself a packed varint stream - and decoded from there.
  std::vector<std::vector<b200tfs_run>> all_runs(m);
  uint64_t gather_bytes = 0;
  for (int j = 0; j < m; ++j) {
    const b200tfs_output& o = outs[j];
    if (!o.n_elems || o.n_runs == 0) continue;
    int rc = collect_runs(c, o, all_runs[j]);
    if (rc) return rc;
    const uint32_t kind = dtype_info(o.dtype).kind;
    if (kind == VK_VARINT || kind == VK_BOOL) {
      bool plain = all_runs[j].size() <= (size_t)B200TFS_MAX_RUNS;
      uint64_t bytes = 0;
      for (const b200tfs_run& r : all_runs[j]) { plain = plain && r.count == 1; bytes += (uint64_t)r.len * r.count; }
      if (!plain) gather_bytes = ((gather_bytes + 15) & ~15ull) + bytes;
    }
  }
  if (gather_bytes) { int rc = grow_dev(c, c->gather_dev, gather_bytes + 64); if (rc) return rc; }
  uint64_t gather_cur = 0;
  for (int j = 0; j < m; ++j) {
    const b200tfs_output& o = outs[j];
    const std::vector<b200tfs_run>& runs = all_runs[j];
    const uint8_t* w = (const uint8_t*)arena_dev + (out_rec_off ? out_rec_off[j] : 0);  // table offsets are record-relative
    if (status) status[j] = B200TFS_OK;
    if (!o.n_elems) continue;
    const bool content_only = o.n_runs == 0 && o.content_len && o.content_len == o.dst_bytes;
    // TF's MakeNdarray convention, asked for by the caller.  For packed varints the table cannot know the element count
    // (status OK unless there are fewer value BYTES than elements): the flag then tells the decode kernels to tolerate it.
    const bool pad = (o.flags & B200TFS_OF_PAD_EDGE) &&
                     (o.status == B200TFS_E_SHAPE || (o.status == B200TFS_OK && (o.flags & B200TFS_OF_VARINT)));
    if (o.status != B200TFS_OK && !content_only && !pad)
      return fail(B200TFS_E_ARG, "output %d was tabulated with status %d: nothing to unpack", j, o.status);
    DtypeInfo di = dtype_info(o.dtype);
    if (di.kind == VK_NONE || di.kind == VK_STRING) return fail(B200TFS_E_DTYPE, "output %d: dtype %d has no device payload", j, o.dtype);
    if (!dst_dev[j]) return fail(B200TFS_E_ARG, "output %d: dst is NULL", j);
    int32_t want = dst_dtype ? dst_dtype[j] : o.dtype;
    const bool half_as_value = (want == B200TFS_DT_HALF_REFQUIRK && o.dtype == DT_HALF);
    if (half_as_value) want = DT_HALF;
    uint8_t* dst = (uint8_t*)dst_dev[j];
    if (pad && want != o.dtype) return fail(B200TFS_E_DTYPE, "output %d: cast together with padding is not supported", j);
    if (o.n_runs == 0 && pad && !content_only) {   // no values at all: zeros
      CU(cudaMemsetAsync(dst, 0, o.dst_bytes, c->stream));
      continue;
    }
    if (o.n_runs == 0) {
      // tolerant path chosen by the caller: raw little-endian bytes from tensor_content
      if (o.content_len != o.dst_bytes) return fail(B200TFS_E_SHAPE, "output %d: no values (tensor_content %llu bytes, need %llu)", j,
                                                     (unsigned long long)o.content_len, (unsigned long long)o.dst_bytes);
      if (want != o.dtype) return fail(B200TFS_E_DTYPE, "output %d: cast from tensor_content is not supported", j);
      pb.payload(w + o.content_off, dst, o.content_len, OP_COPY);
      continue;
    }
    if (di.kind == VK_FIXED) {
      uint32_t op;
      uint32_t num = 1, den = 1;  // dst bytes per src byte
      if (want == o.dtype) op = (o.dtype == DT_FLOAT) ? OP_QUIET_DST : OP_COPY;
      else if (o.dtype == DT_FLOAT && want == DT_HALF) { op = OP_F2H; den = 2; }
      else if (o.dtype == DT_FLOAT && want == DT_BFLOAT16) { op = OP_F2B; den = 2; }
      else return fail(B200TFS_E_DTYPE, "output %d: cast DT %d -> DT %d is not supported", j, o.dtype, want);
      uint64_t run = 0;
      for (const b200tfs_run& r : runs) {
        const uint64_t bytes = (uint64_t)r.len * r.count * num / den;
        if (r.count == 1) pb.payload(w + r.off, dst + run, bytes, op);
        else pb.payload(w + r.off, dst + run, bytes, op, r.len, r.stride);
        run += bytes;
      }
      if (pad) {
        if (run % di.elem_size || run > o.dst_bytes) return fail(B200TFS_E_SHAPE, "output %d: %llu value bytes for a tensor of %llu", j,
                                                                  (unsigned long long)run, (unsigned long long)o.dst_bytes);
        fills.push_back(Fill{dst, di.elem_size, run / di.elem_size, o.n_elems});
      }
    } else {  // packed varints (incl. bool_val)
      if (want != o.dtype) return fail(B200TFS_E_DTYPE, "output %d: cast on varint dtypes is not supported", j);
      VarDecodeJob vj{};
      bool plain = runs.size() <= (size_t)B200TFS_MAX_RUNS;
      for (const b200tfs_run& r : runs) plain = plain && r.count == 1;
      if (plain) {
        vj.n_chunks = (int)runs.size();
        for (size_t k = 0; k < runs.size(); ++k) { vj.src[k] = w + runs[k].off; vj.len[k] = runs[k].len; }
      } else {   // gather the pieces into one stream (same launch as the fixed-width moves, ahead of the decode kernels)
        gather_cur = (gather_cur + 15) & ~15ull;
        uint8_t* g = (uint8_t*)c->gather_dev.p + gather_cur;
        uint64_t at = 0;
        for (const b200tfs_run& r : runs) {
          const uint64_t bytes = (uint64_t)r.len * r.count;
          if (r.count == 1) pb.payload(w + r.off, g + at, bytes, OP_COPY);
          else pb.payload(w + r.off, g + at, bytes, OP_COPY, r.len, r.stride);
          at += bytes;
        }
        gather_cur += at;
        vj.n_chunks = 1; vj.src[0] = g; vj.len[0] = at;
      }
      vj.dst = dst; vj.n_elems = o.n_elems; vj.dtype = o.dtype; vj.out_index = j; vj.half_as_value = half_as_value; vj.pad_edge = pad;
      vjobs.push_back(vj);
    }
  }
  int rc = launch_plan(c, pb);
  if (rc) return rc;
  for (const Fill& f : fills) { CU(launch_fill_edge(f.dst, f.elem_size, f.have, nullptr, f.n_elems, c->stream)); c->launches += 1; }
  c->pending_status.clear();
  if (!vjobs.empty()) {
    if ((rc = run_vardecode(c, vjobs, status))) return rc;
  }
  if (status && wait) {
    CU(cudaStreamSynchronize(c->stream));
    collect_varint_status(c, status);
  }
  return B200TFS_OK;
}

extern "C" int b200tfs_unpack_outputs(b200tfs_ctx* c, const void* arena_dev, int32_t m, const b200tfs_output* outs,
                                      const uint64_t* out_rec_off, void* const* dst_dev, const int32_t* dst_dtype, int32_t* status) {
  return unpack_outputs_impl(c, arena_dev, m, outs, out_rec_off, dst_dev, dst_dtype, status, true);
}

// ------------------------------------------------------------------------------------------------
// fused single-launch decode + CUDA graph capture
// ------------------------------------------------------------------------------------------------
namespace {
struct FusedLayout { uint64_t outs, nouts, specs, status, total; };
FusedLayout fused_layout(int32_t n) {
  FusedLayout L;
  L.outs = 0;
  L.nouts = L.outs + sizeof(b200tfs_output) * (uint64_t)n * kFusedMaxOutputs;
  L.specs = (L.nouts + 4ull * n + 15) & ~15ull;
  L.status = L.specs + sizeof(b200tfs_model_spec) * (uint64_t)n;
  L.total = (L.status + 4ull * n + 15) & ~15ull;
  return L;
}
}  // namespace

extern "C" {

// take over the template a kernel left in pinned memory, if it is newer than what the host knows (stream must be idle)
static void adopt_pinned_template(b200tfs_ctx* c) {
  if (!c->tpl_pinned) return;
  const TplInline& p = *c->tpl_pinned;
  // newer than what the host knows - valid or not: a record 0 that could not be learnt retires the old template too
  if (p.head.serial && (int32_t)(p.head.serial - c->tpl_known.head.serial) > 0) c->tpl_known = p;
}

// tile size of a decode launch over `wire_total` bytes: every CTA of the fused kernel first verifies the record's framing, so
// big batches get fatter tiles than the plain move: measured 0.745 / 0.775 / 0.80 of peak at 64 / 128 / 256 KB
static uint32_t decode_vpt(const b200tfs_ctx* c, int32_t n, const uint64_t* rec_len) {
  uint64_t wire_total = 0;
  for (int i = 0; i < n; ++i) wire_total += rec_len[i];
  // a narrowing launch (b200tfs_set_decode_cast) moves its tiles through the general tile routine, whose rounds of 8 KB run one
  // after the other inside a CTA: small tiles, many CTAs (256 KB tiles: 270 us for the C4 batch; 64 KB: see profiles/r02_c4.md)
  return pick_vec_per_tile(c, c->decode_cast ? wire_total / 2 : wire_total, c->decode_cast ? 65536 : 262144);
}

// Walk record 0 on the host (its bytes are in host memory) and build its template: the launch that follows then takes the
// template path from its first CTA on.  Returns false when the record does not qualify (the kernel will walk it).
static bool host_template(const uint8_t* rec0, uint64_t len, uint32_t vpt, uint64_t dst_stride, uint32_t serial, uint32_t cast, Template* T) {
  T->in.head.valid = 0;
  if (!rec0 || len == 0 || len > 0x7FFFFFFFull) return false;
  b200tfs_output outs[kFusedMaxOutputs + 1];
  b200tfs_model_spec spec;
  int cnt = 0;
  Cursor cur;
  cur_open_host(cur, rec0, (uint32_t)len);
  SpillArea sp{nullptr, 0u, 0u};
  const int st = walk_response(cur, kFusedMaxOutputs, outs, &cnt, &spec, sp);
  if (st != B200TFS_OK) return false;
  const uint64_t used = tpl_layout_outputs(outs, cnt, dst_stride, cast);
  for (int k = 0; k < cnt; ++k) if (outs[k].status == B200TFS_E_SIZE) return false;
  // the kernel's own check: the chunks' tiles must fit the budget the launch gives the record
  tpl_learn(T, cur, (uint32_t)len, outs, cnt, spec, st, vpt, (used + 255) & ~255ull, serial, cast);
  return T->in.head.valid != 0;
}

// the pipelined host decode of ONE record: launch k covers tiles [tile_lo[k], tile_lo[k+1]) of the full grid (the last one also the
// slack CTAs behind them), waits for before[k] and is followed by after[k] on the context's stream
struct DecodeSlices {
  int K = 0;
  uint32_t tile_lo[b200tfs_ctx::kPipeMax + 1] = {};
  cudaEvent_t* before = nullptr;
  cudaEvent_t* after = nullptr;
};

// host_tpl: the template of record 0 as the HOST walked it, when the caller has the record's bytes (the *_host entry points), else nullptr
static int decode_launch(b200tfs_ctx* c, const void* arena_dev, int32_t n, const uint64_t* rec_off, const uint64_t* rec_len, void* dst_dev,
                         uint64_t dst_stride, uint32_t vpt, const Template* host_tpl, const DecodeSlices* sl = nullptr) {
  const uint64_t tile_bytes = 16ull * vpt;
  FusedLayout L = fused_layout(n);
  int rc;
  if ((rc = grow_host(c, c->fused_host, L.total))) return rc;
  if (!c->tpl_dev) {
    if (c->capturing) return fail(B200TFS_E_ARG, "run b200tfs_decode_responses once before capturing it");
    CU(cudaMalloc(&c->tpl_dev, 2 * sizeof(Template) + 64));   // + the three path counters (b200tfs_decode_stats)
    CU(cudaMemsetAsync(c->tpl_dev, 0, 2 * sizeof(Template) + 64, c->stream));
    CU(cudaHostAlloc((void**)&c->tpl_pinned, sizeof(TplInline), cudaHostAllocPortable | cudaHostAllocMapped));
    memset(c->tpl_pinned, 0, sizeof(TplInline));
  }
  FusedParams fp{};
  fp.w = (const uint8_t*)arena_dev; fp.dst = (uint8_t*)dst_dev; fp.dst_stride = dst_stride; fp.n = n; fp.vpt = vpt;
  fp.tpl_read = (const Template*)c->tpl_dev + (c->tpl_flip & 1);
  fp.tpl_write = (Template*)c->tpl_dev + ((c->tpl_flip & 1) ^ 1);
  c->tpl_flip ^= 1;
  fp.tpl_pinned = c->tpl_pinned;
  fp.stats = (unsigned long long*)((uint8_t*)c->tpl_dev + 2 * sizeof(Template));
  if (++c->serial == 0) c->serial = 1;
  fp.serial = c->serial;
  fp.cast = c->decode_cast;
  fp.tpli.head.valid = 0;
  if (host_tpl && host_tpl->in.head.valid) {
    if (vpt <= kStageVecsHost) {   // the single-response / small-batch kernel takes its template from the parameters when the host has one
      // this launch's device template IS the host's walk of record 0: upload it where the kernel reads it
      Slot* slot;
      if ((rc = claim_slot(c, sizeof(Template), &slot))) return rc;
      memcpy(slot->host.p, host_tpl, sizeof(Template));
      CU(cudaMemcpyAsync((void*)fp.tpl_read, slot->host.p, sizeof(Template), cudaMemcpyHostToDevice, c->stream));
      if (slot->done) { CU(cudaEventRecord(slot->done, c->stream)); slot->pending = true; }
      c->tpl_known = host_tpl->in;
      if (!c->opt_no_inline) fp.tpli = host_tpl->in;
    }
  } else {
    // the pinned copy is current once the previous decode launch of this context has completed (the stream itself may
    // well be busy again: the caller's own copy of the next response usually precedes this call)
    if (!c->capturing && (!c->tpl_event_pending || cudaEventQuery(c->tpl_event) == cudaSuccess)) { c->tpl_event_pending = false; adopt_pinned_template(c); }
    const TplHead& h = c->tpl_known.head;
    if (vpt <= kStageVecsHost && !c->opt_no_inline && h.valid && h.rec_len == rec_len[0] && h.vpt == vpt && h.cast == fp.cast && h.dst_need <= dst_stride)
      fp.tpli = c->tpl_known;
  }
  // CTAs per record: a record of the length the host knows a template for gets that template's tiles + the publishing CTA + one
  // spare; any other record ceil(len / tile) + kFusedSlackTiles (its values may lie in several chunks, each rounding up).  With
  // the flat slack, 8 of the 9 CTAs of a 4 KB response and 137 of the 265 of a narrowed 16 MiB one had no tile.  Should a record
  // of that length carry OTHER framing that needs more tiles, its status says so (B200TFS_E_NONCANONICAL, b200tfs.h).
  const TplHead& kh = (host_tpl && host_tpl->in.head.valid) ? host_tpl->in.head : c->tpl_known.head;
  const bool budget_known = kh.valid && kh.vpt == vpt && kh.cast == fp.cast && kh.dst_need <= dst_stride && !c->opt_flat_budget;
  auto ctas_for = [&](uint64_t len) -> uint64_t {
    if (budget_known && len == kh.rec_len) return (uint64_t)kh.total_tiles + 2;
    return (len + tile_bytes - 1) / tile_bytes + kFusedSlackTiles;
  };
  // the table is written by the kernel straight into pinned host memory (unified addressing): ~1 KB
  // of posted PCIe writes per record instead of a device table plus a copy node behind every launch
  uint8_t* d = (uint8_t*)c->fused_host.p;
  if (c->opt_table_dev) {
    if ((rc = grow_dev(c, c->fused_dev, L.total))) return rc;
    d = (uint8_t*)c->fused_dev.p;
  }
  fp.outs = (b200tfs_output*)(d + L.outs); fp.n_outs = (int32_t*)(d + L.nouts);
  fp.specs = (b200tfs_model_spec*)(d + L.specs); fp.status = (int32_t*)(d + L.status);
  // lay the CTAs of a launch out: record r owns `per(len)` consecutive CTAs; small batches in the parameters, else one uploaded table
  auto lay_out = [&](FusedParams& q, auto&& per, uint64_t* grid_out) -> int {
    uint64_t grid = 0;
    if (n <= kFusedInlineRecs) {
      for (int i = 0; i < n; ++i) {
        q.inl.off[i] = rec_off[i]; q.inl.len[i] = rec_len[i]; q.inl.tile_start[i] = (uint32_t)grid;
        grid += per(rec_len[i]);
      }
      q.inl.tile_start[n] = (uint32_t)grid;
    } else {
      // tables: cta_rec[grid] | tile_start[n+1] | rec_off[n] | rec_len[n]
      std::vector<uint32_t> ts(n + 1);
      for (int i = 0; i < n; ++i) { ts[i] = (uint32_t)grid; grid += per(rec_len[i]); }
      ts[n] = (uint32_t)grid;
      if (grid > 0x7FFFFFFFull) return fail(B200TFS_E_TOOBIG, "batch needs more than 2^31 tiles");
      const uint64_t o_ts = (grid * 4 + 15) & ~15ull, o_off = (o_ts + 4ull * (n + 1) + 15) & ~15ull, o_len = o_off + 8ull * n;
      const uint64_t image = o_len + 8ull * n;
      Slot* slot;
      int rc2;
      if ((rc2 = claim_slot(c, image, &slot))) return rc2;
      uint8_t* h = (uint8_t*)slot->host.p;
      uint32_t* cr = (uint32_t*)h;
      for (int i = 0; i < n; ++i) for (uint32_t t = ts[i]; t < ts[i + 1]; ++t) cr[t] = (uint32_t)i;
      memcpy(h + o_ts, ts.data(), 4ull * (n + 1));
      memcpy(h + o_off, rec_off, 8ull * n);
      memcpy(h + o_len, rec_len, 8ull * n);
      if ((rc2 = upload_slot(c, slot, image))) return rc2;
      uint8_t* sd = (uint8_t*)slot->dev.p;
      q.cta_rec = (const uint32_t*)sd; q.tile_start = (const uint32_t*)(sd + o_ts);
      q.rec_off = (const uint64_t*)(sd + o_off); q.rec_len = (const uint64_t*)(sd + o_len);
    }
    if (grid > 0x7FFFFFFFull) return fail(B200TFS_E_TOOBIG, "batch needs more than 2^31 tiles");
    *grid_out = grid;
    return B200TFS_OK;
  };
  // The narrowing decode of a batch whose template the host knows, as three launches: the narrowing tile move runs twice as
  // long inside the fused kernel as in the generic move engine (profiles/r02_c4.md), so (1) two CTAs per record verify the
  // framing against the host's template - handed over in the parameters, the very one the plan below is built from - leave the
  // verdict in guard[r] and publish the table, (2) move_guarded_kernel moves every record's chunks from a host-built plan
  // and stores only where guard[r] says so, (3) the whole decode runs for the records still unguarded (none, normally: its
  // CTAs leave after one load).
  const TplInline* kt = (host_tpl && host_tpl->in.head.valid) ? &host_tpl->in : &c->tpl_known;
  bool split = fp.cast && !sl && budget_known && kh.total_tiles > 0 && (uint64_t)n * kh.rec_len >= (4ull << 20) && !c->opt_no_inline && !c->opt_flat_budget;
  for (int i = 0; i < n && split; ++i) split = rec_len[i] == kh.rec_len;
  if (split) {
    if ((rc = grow_dev(c, c->guard_dev, 4ull * n))) return rc;
    FusedParams fa = fp;
    fa.mode = 1; fa.guard = (uint32_t*)c->guard_dev.p; fa.tpli = *kt;
    uint64_t ga = 0;
    if ((rc = lay_out(fa, [](uint64_t) -> uint64_t { return 2; }, &ga))) return rc;
    CU(launch_decode_fused(fa, (uint32_t)ga, c->stream));
    PlanBuilder pb;
    uint32_t per_rec = 0;
    for (uint32_t q = 0; q < kt->head.n_chunks; ++q) if (kt->chunk[q].n_tiles) ++per_rec;
    pb.items.reserve((size_t)n * per_rec);
    for (int i = 0; i < n; ++i)
      for (uint32_t q = 0; q < kt->head.n_chunks; ++q) {
        const TplChunk& ch = kt->chunk[q];
        if (!ch.n_tiles) continue;
        const bool narrow = ch.op == OP_F2H || ch.op == OP_F2B;
        const uint64_t n_out = narrow ? ch.len / 2 : ch.len;
        pb.items.push_back(MoveItem{(const uint8_t*)arena_dev + rec_off[i] + ch.wire_off, (uint8_t*)dst_dev + (uint64_t)i * dst_stride + ch.dst_off, n_out,
                                    ch.op, 0, 0, 0});
        pb.large_bytes += n_out;
      }
    pb.guard = (const uint32_t*)c->guard_dev.p; pb.guard_div = per_rec;
    BuiltPlan bp;
    if ((rc = build_plan(c, pb, true, &bp))) return rc;
    CU(launch_move_guarded(bp.plan_dev, bp.ph.n_tiles, c->stream));
    if (bp.slot && bp.slot->done && !c->capturing) { CU(cudaEventRecord(bp.slot->done, c->stream)); bp.slot->pending = true; }
    c->launches += 2;
    fp.mode = 2; fp.guard = (uint32_t*)c->guard_dev.p;
  }
  uint64_t grid = 0;
  if ((rc = lay_out(fp, ctas_for, &grid))) return rc;
  if (sl) {
    if (n != 1 || !fp.tpli.head.valid) return fail(B200TFS_E_ARG, "internal: sliced decode without a host template");
    fp.trusted = 1;
    for (int k = 0; k < sl->K; ++k) {
      CU(cudaStreamWaitEvent(c->stream, sl->before[k], 0));
      fp.tile_bias = sl->tile_lo[k];
      const uint32_t end = (k + 1 == sl->K) ? (uint32_t)grid : sl->tile_lo[k + 1];
      CU(launch_decode_fused(fp, end - sl->tile_lo[k], c->stream));
      CU(cudaEventRecord(sl->after[k], c->stream));
      c->launches += 1;
    }
  } else {
    CU(launch_decode_fused(fp, (uint32_t)grid, c->stream));
    c->launches += 1;
  }
  c->fused_n = n;
  if (!c->capturing) { CU(cudaEventRecord(c->tpl_event, c->stream)); c->tpl_event_pending = true; }
  return B200TFS_OK;
}

int b200tfs_decode_responses(b200tfs_ctx* c, const void* arena_dev, int32_t n, const uint64_t* rec_off, const uint64_t* rec_len,
                             void* dst_dev, uint64_t dst_stride) {
  if (!c || n < 0 || (n && (!arena_dev || !rec_off || !rec_len || !dst_dev))) return fail(B200TFS_E_ARG, "bad arguments");
  if (n == 0) return B200TFS_OK;
  if (dst_stride & 255) return fail(B200TFS_E_ARG, "dst_stride must be a multiple of 256");
  if (!c->capturing) CU(cudaSetDevice(c->device));
  return decode_launch(c, arena_dev, n, rec_off, rec_len, dst_dev, dst_stride, decode_vpt(c, n, rec_len), nullptr);
}

int b200tfs_set_decode_cast(b200tfs_ctx* c, int32_t float_as) {
  if (!c) return fail(B200TFS_E_ARG, "ctx is NULL");
  if (c->capturing) return fail(B200TFS_E_ARG, "cannot change the decode cast during graph capture");
  if (float_as != 0 && float_as != DT_FLOAT && float_as != DT_HALF && float_as != DT_BFLOAT16)
    return fail(B200TFS_E_DTYPE, "DT_FLOAT outputs can leave as DT_FLOAT, DT_HALF or DT_BFLOAT16, not as dtype %d", float_as);
  c->decode_cast = (float_as == DT_HALF || float_as == DT_BFLOAT16) ? (uint32_t)float_as : 0u;
  return B200TFS_OK;
}

int b200tfs_decode_results(b200tfs_ctx* c, int32_t n, b200tfs_output* outs, int32_t* n_outs, b200tfs_model_spec* specs,
                           int32_t* rec_status) {
  if (!c || n < 0) return fail(B200TFS_E_ARG, "bad arguments");
  if (c->capturing) return fail(B200TFS_E_ARG, "cannot collect results during graph capture");
  if (n > c->fused_n) return fail(B200TFS_E_ARG, "only %d records were decoded", c->fused_n);
  FusedLayout L = fused_layout(c->fused_n);
  if (c->opt_table_dev && c->fused_dev.p) CU(cudaMemcpyAsync(c->fused_host.p, c->fused_dev.p, L.total, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  const uint8_t* h = (const uint8_t*)c->fused_host.p;
  if (outs) memcpy(outs, h + L.outs, sizeof(b200tfs_output) * (uint64_t)n * kFusedMaxOutputs);
  if (n_outs) memcpy(n_outs, h + L.nouts, 4ull * n);
  if (specs) memcpy(specs, h + L.specs, sizeof(b200tfs_model_spec) * (uint64_t)n);
  if (rec_status) memcpy(rec_status, h + L.status, 4ull * n);
  return B200TFS_OK;
}

int b200tfs_decode_stats(b200tfs_ctx* c, uint64_t* param_template, uint64_t* device_template, uint64_t* walked) {
  if (!c) return fail(B200TFS_E_ARG, "ctx is NULL");
  if (c->capturing) return fail(B200TFS_E_ARG, "cannot read the counters during graph capture");
  unsigned long long v[3] = {0, 0, 0};
  if (c->tpl_dev) {
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(v, (uint8_t*)c->tpl_dev + 2 * sizeof(Template), sizeof v, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  if (param_template) *param_template = v[0];
  if (device_template) *device_template = v[1];
  if (walked) *walked = v[2];
  return B200TFS_OK;
}

int b200tfs_capture_begin(b200tfs_ctx* c) {
  if (!c) return fail(B200TFS_E_ARG, "ctx is NULL");
  if (c->capturing) return fail(B200TFS_E_ARG, "already capturing");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  adopt_pinned_template(c);   // captured single-response decodes carry the template known now in their parameters
  for (auto& s : c->slots) s.pending = false;
  CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
  c->capturing = true;
  return B200TFS_OK;
}

int b200tfs_capture_end(b200tfs_ctx* c, void** graph_exec) {
  if (!c || !graph_exec) return fail(B200TFS_E_ARG, "NULL argument");
  if (!c->capturing) return fail(B200TFS_E_ARG, "not capturing");
  c->capturing = false;
  cudaGraph_t g = nullptr;
  CU(cudaStreamEndCapture(c->stream, &g));
  cudaGraphExec_t ge = nullptr;
  cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return fail(B200TFS_E_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
  c->has_graphs = true;
  *graph_exec = ge;
  return B200TFS_OK;
}

int b200tfs_graph_launch(b200tfs_ctx* c, void* graph_exec) {
  if (!c || !graph_exec) return fail(B200TFS_E_ARG, "NULL argument");
  CU(cudaGraphLaunch((cudaGraphExec_t)graph_exec, c->stream));
  return B200TFS_OK;
}

int b200tfs_graph_destroy(void* graph_exec) {
  if (graph_exec) CU(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
  return B200TFS_OK;
}

int b200tfs_wait_event(b200tfs_ctx* c, void* ev) {
  if (!c || !ev) return fail(B200TFS_E_ARG, "NULL argument");
  CU(cudaStreamWaitEvent(c->stream, (cudaEvent_t)ev, 0));
  return B200TFS_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// host-buffer entry points
// ------------------------------------------------------------------------------------------------
namespace {

// bytes of source memory a tensor occupies
uint64_t tensor_src_bytes(const b200tfs_tensor& t) {
  if (t.flags & B200TFS_F_PRESERIALIZED) return t.packed_len;
  uint64_t n = 1;
  for (int i = 0; i < t.rank; ++i) n *= (uint64_t)t.dims[i];
  return n * dtype_info(t.src_dtype).elem_size;
}

// one host tensor's place in the device staging buffer
struct StagePiece { const uint8_t* dev; const uint8_t* host; uint64_t nb; };

// Host-to-device copies of a batch, merged where consecutive ones continue each other on BOTH sides (a batch whose tensors lie
// back to back in one pinned buffer - bench.py's lanes, a caller's arena - becomes one copy instead of one per tensor: ~2 us of
// driver time each, 15-20 % of a 602 KB tensor's time on the link).
struct CopyMerger {
  cudaStream_t stream;
  uint8_t* d = nullptr; const uint8_t* h = nullptr; uint64_t n = 0;
  explicit CopyMerger(cudaStream_t s) : stream(s) {}
  cudaError_t add(const void* dst, const void* src, uint64_t nb) {
    if (!nb) return cudaSuccess;
    if (n && (const uint8_t*)dst >= d + n && (const uint8_t*)dst - (d + n) == (const uint8_t*)src - (h + n) && (const uint8_t*)dst - (d + n) < 256) {
      n = (uint64_t)((const uint8_t*)dst - d) + nb;      // the gap (alignment padding, equal on both sides) travels along
      return cudaSuccess;
    }
    cudaError_t e = flush();
    d = (uint8_t*)dst; h = (const uint8_t*)src; n = nb;
    return e;
  }
  cudaError_t flush() {
    cudaError_t e = cudaSuccess;
    if (n) e = cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, stream);
    n = 0;
    return e;
  }
};

// Give every tensor of the batch a place in the device staging buffer and return device-pointing clones.  `defer` == nullptr:
// the copies are queued on the context's stream right here; else they are only listed (the pipelined path issues them in slices).
int stage_tensors(b200tfs_ctx* c, std::vector<b200tfs_tensor>& ts, std::vector<StagePiece>* defer = nullptr) {
  uint64_t total = 0;
  for (auto& t : ts) {
    if (!(t.flags & B200TFS_F_PRESERIALIZED)) {
      if (dtype_info(t.src_dtype).kind == VK_NONE || dtype_info(t.src_dtype).kind == VK_STRING)
        return fail(B200TFS_E_DTYPE, "dtype %d has no device payload", t.src_dtype);
      if (t.rank < 0 || t.rank > 254 || (t.rank && !t.dims)) return fail(B200TFS_E_SHAPE, "bad rank/dims");
      for (int i = 0; i < t.rank; ++i) if (t.dims[i] < 0) return fail(B200TFS_E_SHAPE, "negative dim");
    }
    if (t.flags & B200TFS_F_DEVICE_DATA) continue;   // already in HBM
    total += tensor_src_bytes(t) + 255;      // whatever order the places are handed out in
  }
  int rc = grow_dev(c, c->stage_dev, total + 256);
  if (rc) return rc;
  // places are handed out in order of HOST address: tensors that lie back to back in the caller's memory (all images of a batch
  // in one buffer, all labels in another) then lie back to back in the staging buffer too, and their copies merge
  std::vector<uint32_t> order;
  order.reserve(ts.size());
  for (uint32_t i = 0; i < ts.size(); ++i) {
    if (ts[i].flags & B200TFS_F_DEVICE_DATA) { ts[i].flags &= ~B200TFS_F_DEVICE_DATA; continue; }
    order.push_back(i);
  }
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return (uintptr_t)ts[a].data < (uintptr_t)ts[b].data; });
  uint64_t cur = 0;
  CopyMerger cm(c->stream);
  for (uint32_t i : order) {
    b200tfs_tensor& t = ts[i];
    cur = (cur + 255) & ~255ull;
    uint64_t nb = tensor_src_bytes(t);
    if (nb) {
      if (!t.data) return fail(B200TFS_E_ARG, "tensor data pointer is NULL");
      if (defer) defer->push_back(StagePiece{(const uint8_t*)c->stage_dev.p + cur, (const uint8_t*)t.data, nb});
      else CU(cm.add((uint8_t*)c->stage_dev.p + cur, t.data, nb));
    }
    t.data = (uint8_t*)c->stage_dev.p + cur;
    cur += nb;
  }
  CU(cm.flush());
  return B200TFS_OK;
}

bool needs_measure_one(const b200tfs_tensor& t) {
  return !(t.flags & (B200TFS_F_PRESERIALIZED | B200TFS_F_TENSOR_CONTENT)) && dtype_info(t.wire_dtype).kind == VK_VARINT;
}
// a packed-varint input the host can measure itself: its values are in host memory, few, and well-formed
bool host_measurable_varint(const b200tfs_tensor& t) {
  if (!needs_measure_one(t) || t.src_dtype != t.wire_dtype || (t.flags & B200TFS_F_DEVICE_DATA)) return false;
  if (t.rank < 0 || t.rank > 254 || (t.rank && !t.dims)) return false;
  uint64_t ne = 1;
  for (int k = 0; k < t.rank; ++k) { if (t.dims[k] < 0 || t.dims[k] > 4096) return false; ne *= (uint64_t)t.dims[k]; if (ne > 4096) return false; }
  return ne == 0 || t.data != nullptr;
}

// Is this host pointer page-locked memory the device can address (cudaHostAlloc / cudaHostRegister under unified addressing)?
// Then a kernel may write its output there itself - posted PCIe writes at the link rate (4 MiB: 96 us from launch to synchronise
// against 178 for H2D + kernel + D2H, profiles/r02_pipeline.md) - and the device-to-host copy disappears.  (The other direction
// does not pay: SM-issued reads of host memory run at ~34 GB/s, the copy engine's at 55.)
uint8_t* device_view_of_host(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
  return (uint8_t*)a.devicePointer;
}

// source bytes per output byte of a move op, as a fraction num/den; 0/0: the op cannot be cut
void op_ratio(uint32_t op, uint32_t* num, uint32_t* den) {
  switch (op) {
    case OP_COPY: case OP_QUIET_SRC: case OP_QUIET_DST: case OP_BOOL: *num = 1; *den = 1; break;
    case OP_H2F: case OP_B2F: *num = 1; *den = 2; break;
    case OP_F2H: case OP_F2B: *num = 2; *den = 1; break;
    default: *num = 0; *den = 0; break;
  }
}

// The pipelined encode: the batch's large payloads are cut into up to kPipeMax slices of consecutive wire bytes; slice k's source
// bytes travel H2D on one stream while slice k-1 is being encoded on the context's stream and slice k-2's wire bytes travel D2H on
// a third - one big request alone keeps both PCIe directions busy (monolithic H2D -> kernel -> D2H).
// Framing bytes and small payloads go with slice 0.  Returns B200TFS_OK with *done = false when the batch does not qualify.
// `direct`: the plan's destinations already lie in the caller's (pinned) wire buffer: no device-to-host copies, two streams.
int encode_pipelined(b200tfs_ctx* c, PlanBuilder& pb, const std::vector<StagePiece>& pieces, uint8_t* wire_host, uint64_t lo, uint64_t hi,
                     bool direct, bool* done) {
  *done = false;
  if (!c->pipe_min || pb.large_bytes < c->pipe_min || !pb.varjobs.empty() || pb.items.empty()) return B200TFS_OK;
  for (auto& it : pb.items) {
    uint32_t num, den;
    op_ratio(it.op, &num, &den);
    if (!num || it.glen) return B200TFS_OK;
  }
  int K = (int)std::min<uint64_t>(c->pipe_max, std::max<uint64_t>(2, pb.large_bytes / std::max<uint64_t>(c->pipe_min, 1ull << 18)));
  const uint64_t per = ((pb.large_bytes + K - 1) / K + 65535) & ~65535ull;    // output bytes per slice, cut on 64 KB
  const uint8_t* arena = (const uint8_t*)c->arena_dev.p;
  // which piece feeds an item: the one that contains its source (device-resident inputs have none)
  auto piece_of = [&](const uint8_t* src) -> const StagePiece* {      // pieces are in ascending staging-address order (stage_tensors)
    auto it = std::upper_bound(pieces.begin(), pieces.end(), src, [](const uint8_t* s, const StagePiece& p) { return s < p.dev; });
    if (it == pieces.begin()) return nullptr;
    --it;
    return (src >= it->dev && src < it->dev + it->nb) ? &*it : nullptr;
  };
  std::vector<const StagePiece*> feeds(pb.items.size());
  std::vector<char> is_large(pieces.size(), 0);
  for (size_t i = 0; i < pb.items.size(); ++i) {
    feeds[i] = piece_of(pb.items[i].src);
    if (feeds[i]) is_large[feeds[i] - pieces.data()] = 1;
  }
  // everything queued on this context so far (the previous call's kernels read the staging buffer) precedes our copies
  cudaEvent_t* ev = c->pipe_ev;
  CU(cudaEventRecord(ev[2 * b200tfs_ctx::kPipeMax], c->stream));
  CU(cudaStreamWaitEvent(c->aux_stream, ev[2 * b200tfs_ctx::kPipeMax], 0));
  if (!direct) CU(cudaStreamWaitEvent(c->d2h_stream, ev[2 * b200tfs_ctx::kPipeMax], 0));
  size_t item = 0;
  uint64_t item_done = 0;        // output bytes of pb.items[item] already handed to a slice
  uint64_t wire_done = lo;       // arena offset up to which the wire has been copied back
  int rc;
  for (int k = 0; k < K && (item < pb.items.size() || k == 0); ++k) {
    PlanBuilder sub;
    CopyMerger cm(c->aux_stream);
    if (k == 0) {
      sub.smalls.swap(pb.smalls);
      sub.blob.swap(pb.blob);
      for (size_t q = 0; q < pieces.size(); ++q)     // sources of small payloads
        if (!is_large[q]) CU(cm.add(pieces[q].dev, pieces[q].host, pieces[q].nb));
      CU(cm.flush());
    }
    uint64_t room = per;
    uint64_t wire_end = wire_done;
    while (item < pb.items.size() && room) {
      const MoveItem& it = pb.items[item];
      uint32_t num, den;
      op_ratio(it.op, &num, &den);
      const bool last_slice = (k == K - 1);
      uint64_t take = std::min<uint64_t>(it.n_out - item_done, last_slice ? ~0ull : room);
      if (take < it.n_out - item_done) take &= ~1023ull;     // cuts fall on 1 KB of output: whole elements and whole 16-byte vectors of either side
      if (!take) break;
      if (!last_slice) room -= std::min(room, take);
      const uint64_t s_off = item_done * num / den, s_len = take * num / den;
      if (feeds[item]) CU(cm.add(it.src + s_off, feeds[item]->host + (it.src + s_off - feeds[item]->dev), s_len));
      sub.payload(it.src + s_off, it.dst + item_done, take, it.op);
      if (!direct) wire_end = (uint64_t)(it.dst + item_done + take - arena);
      item_done += take;
      if (item_done == it.n_out) { ++item; item_done = 0; }
    }
    const bool final_slice = item >= pb.items.size();
    if (final_slice) wire_end = hi;
    CU(cm.flush());
    CU(cudaEventRecord(ev[2 * k], c->aux_stream));
    CU(cudaStreamWaitEvent(c->stream, ev[2 * k], 0));
    if ((rc = launch_plan(c, sub))) return rc;
    if (!direct) {
      CU(cudaEventRecord(ev[2 * k + 1], c->stream));
      CU(cudaStreamWaitEvent(c->d2h_stream, ev[2 * k + 1], 0));
      if (wire_end > wire_done)
        CU(cudaMemcpyAsync(wire_host + (wire_done - lo), arena + wire_done, wire_end - wire_done, cudaMemcpyDeviceToHost, c->d2h_stream));
      wire_done = wire_end;
    }
    if (final_slice) break;
  }
  if (!direct) {   // the context's stream is where callers wait: it ends behind the last copy
    CU(cudaEventRecord(ev[2 * b200tfs_ctx::kPipeMax + 1], c->d2h_stream));
    CU(cudaStreamWaitEvent(c->stream, ev[2 * b200tfs_ctx::kPipeMax + 1], 0));
  }
  c->pipelined_calls += 1;
  *done = true;
  return B200TFS_OK;
}

}  // namespace

extern "C" {

int b200tfs_encode_tensor_protos_host(b200tfs_ctx* c, int32_t n, const b200tfs_tensor* tensors, void* wire_host, uint64_t wire_cap,
                                      uint64_t* rec_off, uint64_t* rec_len) {
  if (!c || n < 0 || (n && (!tensors || !wire_host || !rec_off || !rec_len))) return fail(B200TFS_E_ARG, "bad arguments");
  if (n == 0) return B200TFS_OK;
  CU(cudaSetDevice(c->device));
  std::vector<b200tfs_tensor> ts(tensors, tensors + n);
  int rc = stage_tensors(c, ts);
  if (rc) return rc;
  if ((rc = b200tfs_measure(c, n, ts.data()))) return rc;
  uint64_t need = 0;
  if ((rc = b200tfs_tensor_arena_size(n, ts.data(), &need))) return rc;
  if ((rc = grow_dev(c, c->arena_dev, need))) return rc;
  if ((rc = b200tfs_encode_tensor_protos(c, n, ts.data(), c->arena_dev.p, c->arena_dev.cap, rec_off, rec_len))) return rc;
  const uint64_t lo = rec_off[0], hi = rec_off[n - 1] + rec_len[n - 1];
  if (hi - lo > wire_cap) return fail(B200TFS_E_SIZE, "wire buffer too small: need %llu bytes", (unsigned long long)(hi - lo));
  CU(cudaMemcpyAsync(wire_host, (uint8_t*)c->arena_dev.p + lo, hi - lo, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < n; ++i) rec_off[i] -= lo;
  return B200TFS_OK;
}

int b200tfs_encode_requests_host_async(b200tfs_ctx* c, int32_t n, const b200tfs_request* reqs, void* wire_host, uint64_t wire_cap,
                                       uint64_t* rec_off, uint64_t* rec_len) {
  if (!c || n < 0 || (n && (!reqs || !wire_host || !rec_off || !rec_len))) return fail(B200TFS_E_ARG, "bad arguments");
  if (n == 0) return B200TFS_OK;
  CU(cudaSetDevice(c->device));
  // flatten every input of every request, stage them, then rebuild request structs on the clones
  std::vector<b200tfs_tensor> ts;
  for (int i = 0; i < n; ++i) {
    if (reqs[i].n_inputs < 0 || (reqs[i].n_inputs && !reqs[i].inputs)) return fail(B200TFS_E_ARG, "bad request %d", i);
    ts.insert(ts.end(), reqs[i].inputs, reqs[i].inputs + reqs[i].n_inputs);
  }
  // Packed-varint inputs of up to 4096 elements that lie in HOST memory (labels, ids, one sequence of token ids) are measured
  // right here - a few microseconds of host arithmetic instead of a counting kernel, a device-to-host copy and a stream
  // synchronise (b200tfs_measure: 30-40 us before anything else of the call can be queued).
  bool device_measure = false;
  for (auto& t : ts) {
    if (!host_measurable_varint(t)) { device_measure = device_measure || needs_measure_one(t); continue; }
    uint64_t ne = 1;
    for (int k = 0; k < t.rank; ++k) ne *= (uint64_t)t.dims[k];
    const DtypeInfo di = dtype_info(t.src_dtype);
    t.packed_len = ne ? tiny_total(TinyVar{(const uint8_t*)t.data, (uint32_t)ne, di.elem_size, di.is_signed, 0}) : 0;
  }
  // a batch that needs no device measuring may take the pipelined path: its H2D copies are then issued slice by slice
  const bool try_pipe = c->pipe_min && !c->capturing && !device_measure;
  std::vector<StagePiece> pieces;
  int rc = stage_tensors(c, ts, try_pipe ? &pieces : nullptr);
  if (rc) return rc;
  if (device_measure && (rc = b200tfs_measure(c, (int32_t)ts.size(), ts.data()))) return rc;  // synchronises
  std::vector<b200tfs_request> rq(reqs, reqs + n);
  size_t k = 0;
  for (int i = 0; i < n; ++i) { rq[i].inputs = ts.data() + k; k += (size_t)rq[i].n_inputs; }
  uint64_t need = 0;
  if ((rc = b200tfs_request_arena_size(n, rq.data(), &need))) return rc;
  if (try_pipe) {
    // A pinned wire buffer with room for the arena layout (records 256-byte aligned, largest payload 128-byte aligned) is written
    // by the kernels themselves; rec_off then counts from wire_host like always, but record 0 does not start at 0.
    uint8_t* out_dev = (c->opt_direct_out && c->pipe_min && need <= wire_cap) ? device_view_of_host(wire_host) : nullptr;
    if (out_dev && ((uintptr_t)out_dev & 255)) out_dev = nullptr;
    const bool direct = out_dev != nullptr;
    if (!direct && (rc = grow_dev(c, c->arena_dev, need))) return rc;
    PlanBuilder pb;
    if ((rc = plan_requests(n, rq.data(), direct ? (void*)out_dev : c->arena_dev.p, direct ? wire_cap : c->arena_dev.cap, rec_off, rec_len, pb))) return rc;
    const uint64_t lo = rec_off[0], hi = rec_off[n - 1] + rec_len[n - 1];
    if (!direct && hi - lo > wire_cap) return fail(B200TFS_E_SIZE, "wire buffer too small: need %llu bytes", (unsigned long long)(hi - lo));
    bool done = false;
    if ((rc = encode_pipelined(c, pb, pieces, (uint8_t*)wire_host, lo, hi, direct, &done))) return rc;
    if (!done) {   // too small to be worth slicing: everything on the context's stream, as one piece
      CopyMerger cm(c->stream);
      for (auto& p : pieces) CU(cm.add(p.dev, p.host, p.nb));
      CU(cm.flush());
      if ((rc = launch_plan(c, pb))) return rc;
      if ((rc = run_varjobs(c, pb))) return rc;
      if (!direct) CU(cudaMemcpyAsync(wire_host, (uint8_t*)c->arena_dev.p + lo, hi - lo, cudaMemcpyDeviceToHost, c->stream));
    }
    if (direct) c->direct_calls += 1;
    else for (int i = 0; i < n; ++i) rec_off[i] -= lo;
    return B200TFS_OK;
  }
  if ((rc = grow_dev(c, c->arena_dev, need))) return rc;
  if ((rc = b200tfs_encode_requests(c, n, rq.data(), c->arena_dev.p, c->arena_dev.cap, rec_off, rec_len))) return rc;
  const uint64_t lo = rec_off[0], hi = rec_off[n - 1] + rec_len[n - 1];
  if (hi - lo > wire_cap) return fail(B200TFS_E_SIZE, "wire buffer too small: need %llu bytes", (unsigned long long)(hi - lo));
  CU(cudaMemcpyAsync(wire_host, (uint8_t*)c->arena_dev.p + lo, hi - lo, cudaMemcpyDeviceToHost, c->stream));
  for (int i = 0; i < n; ++i) rec_off[i] -= lo;
  return B200TFS_OK;
}

int b200tfs_pipelined_calls(b200tfs_ctx* c, uint64_t* count) {
  if (!c || !count) return fail(B200TFS_E_ARG, "bad arguments");
  *count = c->pipelined_calls;
  return B200TFS_OK;
}

int b200tfs_direct_calls(b200tfs_ctx* c, uint64_t* count) {
  if (!c || !count) return fail(B200TFS_E_ARG, "bad arguments");
  *count = c->direct_calls;
  return B200TFS_OK;
}

int b200tfs_set_pipeline(b200tfs_ctx* c, uint64_t min_bytes, int32_t max_slices) {
  if (!c) return fail(B200TFS_E_ARG, "ctx is NULL");
  if (min_bytes && (max_slices < 2 || max_slices > b200tfs_ctx::kPipeMax)) return fail(B200TFS_E_ARG, "max_slices must be in [2, %d]", b200tfs_ctx::kPipeMax);
  c->pipe_min = min_bytes;
  if (min_bytes) c->pipe_max = max_slices;
  return B200TFS_OK;
}

int b200tfs_encode_requests_host(b200tfs_ctx* c, int32_t n, const b200tfs_request* reqs, void* wire_host, uint64_t wire_cap,
                                 uint64_t* rec_off, uint64_t* rec_len) {
  int rc = b200tfs_encode_requests_host_async(c, n, reqs, wire_host, wire_cap, rec_off, rec_len);
  if (rc) return rc;
  CU(cudaStreamSynchronize(c->stream));
  return B200TFS_OK;
}

static int stage_wire(b200tfs_ctx* c, const void* wire_host, int32_t n, const uint64_t* rec_off, const uint64_t* rec_len, uint64_t* span) {
  uint64_t hi = 0;
  for (int i = 0; i < n; ++i) hi = std::max(hi, rec_off[i] + rec_len[i]);
  int rc = grow_dev(c, c->stage_dev, hi + 64);
  if (rc) return rc;
  if (hi) CU(cudaMemcpyAsync(c->stage_dev.p, wire_host, hi, cudaMemcpyHostToDevice, c->stream));
  c->stage_shift = 0;
  *span = hi;
  return B200TFS_OK;
}

int b200tfs_parse_responses_host(b200tfs_ctx* c, const void* wire_host, int32_t n, const uint64_t* rec_off, const uint64_t* rec_len,
                                 int32_t max_outputs, b200tfs_output* outs, int32_t* n_outs, b200tfs_model_spec* specs,
                                 int32_t* rec_status) {
  if (!c || n < 0 || (n && (!wire_host || !rec_off || !rec_len))) return fail(B200TFS_E_ARG, "bad arguments");
  if (n == 0) return B200TFS_OK;
  CU(cudaSetDevice(c->device));
  uint64_t span;
  int rc = stage_wire(c, wire_host, n, rec_off, rec_len, &span);
  if (rc) return rc;
  return parse_common(c, c->stage_dev.p, n, rec_off, rec_len, max_outputs, false, outs, n_outs, specs, rec_status);
}

int b200tfs_parse_tensor_protos_host(b200tfs_ctx* c, const void* wire_host, int32_t n, const uint64_t* rec_off, const uint64_t* rec_len,
                                     b200tfs_output* outs, int32_t* rec_status) {
  if (!c || n < 0 || (n && (!wire_host || !rec_off || !rec_len))) return fail(B200TFS_E_ARG, "bad arguments");
  if (n == 0) return B200TFS_OK;
  CU(cudaSetDevice(c->device));
  uint64_t span;
  int rc = stage_wire(c, wire_host, n, rec_off, rec_len, &span);
  if (rc) return rc;
  return parse_common(c, c->stage_dev.p, n, rec_off, rec_len, 1, true, outs, nullptr, nullptr, rec_status);
}

int b200tfs_decode_responses_host_async(b200tfs_ctx* c, const void* wire_host, int32_t n, const uint64_t* rec_off, const uint64_t* rec_len,
                                        void* dst_host, uint64_t dst_stride) {
  if (!c || n < 0 || (n && (!wire_host || !rec_off || !rec_len || !dst_host))) return fail(B200TFS_E_ARG, "bad arguments");
  if (n == 0) return B200TFS_OK;
  if (dst_stride & 255) return fail(B200TFS_E_ARG, "dst_stride must be a multiple of 256");
  CU(cudaSetDevice(c->device));
  // The wire is in host memory: walk record 0 HERE (sub-microsecond: the walk steps over the payload) so that the launch
  // takes the template path from its first CTA on, and place the copy so that the largest value chunk lands 16-byte
  // aligned on the device (aligned loads and stores, no funnel shifts).
  const uint32_t vpt = decode_vpt(c, n, rec_len);
  Template T;
  uint64_t shift = 0;
  const bool have = vpt <= kStageVecsHost && !c->capturing &&
                    host_template((const uint8_t*)wire_host + rec_off[0], rec_len[0], vpt, dst_stride, c->serial + 1 ? c->serial + 1 : 1, c->decode_cast, &T);
  if (have) {
    uint32_t big = 0;
    for (uint32_t q = 1; q < T.in.head.n_chunks; ++q) if (T.in.chunk[q].len > T.in.chunk[big].len) big = q;
    if (T.in.head.n_chunks) shift = (16 - ((rec_off[0] + T.in.chunk[big].wire_off) & 15)) & 15;
  }
  uint64_t hi = 0;
  for (int i = 0; i < n; ++i) hi = std::max(hi, rec_off[i] + rec_len[i]);
  int rc = grow_dev(c, c->stage_dev, hi + 64);
  if (rc) return rc;
  c->stage_shift = shift;
  // a pinned destination is written by the kernel itself (posted PCIe writes): no staging buffer, no device-to-host copy
  uint8_t* dst_direct = (c->opt_direct_out && c->pipe_min) ? device_view_of_host(dst_host) : nullptr;
  if (dst_direct && ((uintptr_t)dst_direct & 255)) dst_direct = nullptr;
  if (!dst_direct && (rc = grow_dev(c, c->arena_dev, dst_stride * (uint64_t)n + 256))) return rc;
  uint8_t* dst_dev = dst_direct ? dst_direct : (uint8_t*)c->arena_dev.p;
  if (dst_direct) c->direct_calls += 1;
  uint8_t* wire_dev = (uint8_t*)c->stage_dev.p + shift;
  // One large record whose values lie in one fixed-width chunk: slices of tiles, pipelined like the encode (wire bytes of slice
  // k+1 travel H2D while slice k is decoded and the tensor bytes of slice k-1 travel D2H).  The kernel skips its framing verdict:
  // the template was built from these very bytes, and a slice runs before the record's tail has arrived.
  const TplChunk& ch = T.in.chunk[0];
  const uint64_t tile_bytes = 16ull * vpt;
  if (have && n == 1 && c->pipe_min && rec_len[0] >= c->pipe_min && !c->opt_no_inline && T.in.head.n_chunks == 1 && !ch.is_varint &&
      (ch.op == OP_COPY || ch.op == OP_QUIET_DST) && ch.n_tiles >= 2 && ch.n_tiles == T.in.head.total_tiles) {
    DecodeSlices sl;
    sl.K = (int)std::min<uint64_t>(std::min<uint64_t>(c->pipe_max, ch.n_tiles),
                                   std::max<uint64_t>(2, rec_len[0] / std::max<uint64_t>(c->pipe_min, 1ull << 18)));
    for (int k = 0; k <= sl.K; ++k) sl.tile_lo[k] = (uint32_t)((uint64_t)ch.n_tiles * k / sl.K);
    cudaEvent_t* ev = c->pipe_ev;
    sl.before = ev; sl.after = ev + b200tfs_ctx::kPipeMax;
    CU(cudaEventRecord(ev[2 * b200tfs_ctx::kPipeMax], c->stream));
    CU(cudaStreamWaitEvent(c->aux_stream, ev[2 * b200tfs_ctx::kPipeMax], 0));
    if (!dst_direct) CU(cudaStreamWaitEvent(c->d2h_stream, ev[2 * b200tfs_ctx::kPipeMax], 0));
    uint64_t wire_done = 0;
    for (int k = 0; k < sl.K; ++k) {
      // tiles below tile_lo[k+1] read no further than 48 bytes past their last vector (the next 16-byte block of a shifted source)
      const uint64_t w_end = (k + 1 == sl.K) ? hi : std::min<uint64_t>(hi, rec_off[0] + ch.wire_off + sl.tile_lo[k + 1] * tile_bytes + 64);
      if (w_end > wire_done)
        CU(cudaMemcpyAsync(wire_dev + wire_done, (const uint8_t*)wire_host + wire_done, w_end - wire_done, cudaMemcpyHostToDevice, c->aux_stream));
      wire_done = std::max(wire_done, w_end);
      CU(cudaEventRecord(sl.before[k], c->aux_stream));
    }
    if ((rc = decode_launch(c, wire_dev, n, rec_off, rec_len, dst_dev, dst_stride, vpt, &T, &sl))) return rc;
    if (!dst_direct) {
      const uint64_t need = std::min<uint64_t>(dst_stride, T.in.head.dst_need);
      uint64_t dst_done = 0;
      for (int k = 0; k < sl.K; ++k) {
        // tiles below tile_lo[k+1] have written every byte of the slot below the first vector of tile tile_lo[k+1]
        const uint64_t d_end = (k + 1 == sl.K) ? need : std::min<uint64_t>(need, ch.dst_off + sl.tile_lo[k + 1] * tile_bytes);
        CU(cudaStreamWaitEvent(c->d2h_stream, sl.after[k], 0));
        if (d_end > dst_done)
          CU(cudaMemcpyAsync((uint8_t*)dst_host + dst_done, (uint8_t*)c->arena_dev.p + dst_done, d_end - dst_done, cudaMemcpyDeviceToHost, c->d2h_stream));
        dst_done = std::max(dst_done, d_end);
      }
      CU(cudaEventRecord(ev[2 * b200tfs_ctx::kPipeMax + 1], c->d2h_stream));
      CU(cudaStreamWaitEvent(c->stream, ev[2 * b200tfs_ctx::kPipeMax + 1], 0));
    }
    c->pipelined_calls += 1;
    return B200TFS_OK;
  }
  if (hi) CU(cudaMemcpyAsync(wire_dev, wire_host, hi, cudaMemcpyHostToDevice, c->stream));
  if ((rc = decode_launch(c, wire_dev, n, rec_off, rec_len, dst_dev, dst_stride, vpt, have ? &T : nullptr))) return rc;
  if (!dst_direct) CU(cudaMemcpyAsync(dst_host, c->arena_dev.p, dst_stride * (uint64_t)n, cudaMemcpyDeviceToHost, c->stream));
  return B200TFS_OK;
}

int b200tfs_unpack_outputs_host(b200tfs_ctx* c, int32_t m, const b200tfs_output* outs, const uint64_t* out_rec_off, void* const* dst_host,
                                const int32_t* dst_dtype, int32_t* status) {
  if (!c || m < 0 || (m && (!outs || !dst_host))) return fail(B200TFS_E_ARG, "bad arguments");
  if (m == 0) return B200TFS_OK;
  CU(cudaSetDevice(c->device));
  if (!c->stage_dev.p) return fail(B200TFS_E_ARG, "no staged wire: call b200tfs_parse_*_host first");
  std::vector<uint64_t> off(m), nb(m);
  uint64_t total = 0;
  for (int j = 0; j < m; ++j) {
    int32_t want = dst_dtype ? dst_dtype[j] : outs[j].dtype;
    if (want == B200TFS_DT_HALF_REFQUIRK) want = DT_HALF;
    nb[j] = outs[j].n_elems * dtype_info(want).elem_size;
    total = (total + 255) & ~255ull;
    off[j] = total;
    total += nb[j];
  }
  int rc = grow_dev(c, c->arena_dev, total + 256);
  if (rc) return rc;
  std::vector<void*> dd(m);
  for (int j = 0; j < m; ++j) dd[j] = (uint8_t*)c->arena_dev.p + off[j];
  if ((rc = unpack_outputs_impl(c, (uint8_t*)c->stage_dev.p + c->stage_shift, m, outs, out_rec_off, dd.data(), dst_dtype, status, false))) return rc;
  for (int j = 0; j < m; ++j)
    if (nb[j]) {
      if (!dst_host[j]) return fail(B200TFS_E_ARG, "output %d: dst is NULL", j);
      CU(cudaMemcpyAsync(dst_host[j], dd[j], nb[j], cudaMemcpyDeviceToHost, c->stream));
    }
  CU(cudaStreamSynchronize(c->stream));
  collect_varint_status(c, status);
  return B200TFS_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// varint dtypes (phase 2: kernels in kernels.cu; wired here)
// ------------------------------------------------------------------------------------------------
#include "varint_host.inc"
