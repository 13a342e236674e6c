// frame.h - deferred framing (plan.h): what one thread of frame_requests_kernel does for one request.  Host/device inline so
// that the host can run the very same code (b200tfs_request_frame_deferred: tests and bindings, no device needed).
#pragma once
#include "plan.h"
#include "wire.h"

namespace b200tfs {

B2_HD void frame_request(const FrameTables& ft, uint32_t r) {
  const FrameReq rq = ft.reqs[r];
  uint64_t* val = ft.scratch_vals + rq.first_val;
  for (uint32_t v = 0; v < rq.n_val; ++v) {
    const FrameVal fv = ft.vals[rq.first_val + v];
    uint64_t x = (uint64_t)fv.c;
    for (uint32_t k = 0; k < fv.n_terms; ++k) {
      const FrameTerm t = ft.terms[fv.first_term + k];
      if (t.kind == FT_TOTAL) x += *ft.jobs[t.idx].total;
      else if (t.kind == FT_VAL) x += val[t.idx];
      else x += varint_len(val[t.idx]);
    }
    val[v] = x;
  }
  auto seg_len = [&](const FrameSeg& sg) -> uint64_t {
    switch (sg.kind) {
      case FS_BYTES: return sg.b;
      case FS_VARINT: return varint_len(val[sg.a]);
      case FS_BE32: return 4;
      case FS_ITEM: return ft.items[sg.a].n_out;
      case FS_SMALL: return ft.smalls[sg.a].n_out;
      default: return *ft.jobs[sg.a].total;
    }
  };
  uint64_t pad = 0;
  if (rq.align_seg != ~0u) {
    uint64_t before = 0;
    for (uint32_t k = 0; k < rq.align_seg; ++k) before += seg_len(ft.segs[rq.first_seg + k]);
    pad = (128 - ((rq.slot_off + before) & 127)) & 127;
  }
  const uint64_t total = val[rq.total_val];
  const uint64_t start = rq.slot_off + pad;
  ft.rec_off[r] = start; ft.rec_len[r] = total;
  if (pad + total > rq.slot_cap || total > 0x7FFFFFFFull + 5) {   // cannot happen with the host's worst-case slots; never write outside one
    ft.status[r] = total > 0x7FFFFFFFull + 5 ? B200TFS_E_TOOBIG : B200TFS_E_SIZE;
    for (uint32_t k = 0; k < rq.n_seg; ++k) {   // park the movers on an empty range
      const FrameSeg sg = ft.segs[rq.first_seg + k];
      if (sg.kind == FS_ITEM) ft.items[sg.a].n_out = 0;
      else if (sg.kind == FS_SMALL) ft.smalls[sg.a].n_out = 0;
      else if (sg.kind == FS_VARJOB) { ft.jobs[sg.a].dst = ft.arena + rq.slot_off; ft.jobs[sg.a].cap = 0; }
    }
    return;
  }
  ft.status[r] = B200TFS_OK;
  uint8_t* w = ft.arena + start;
  for (uint32_t k = 0; k < rq.n_seg; ++k) {
    const FrameSeg sg = ft.segs[rq.first_seg + k];
    switch (sg.kind) {
      case FS_BYTES: { const uint8_t* b = ft.blob + sg.a; for (uint32_t q = 0; q < sg.b; ++q) w[q] = b[q]; w += sg.b; break; }
      case FS_VARINT: w += put_varint(w, val[sg.a]); break;
      case FS_BE32: { const uint64_t m = val[sg.a]; w[0] = (uint8_t)(m >> 24); w[1] = (uint8_t)(m >> 16); w[2] = (uint8_t)(m >> 8); w[3] = (uint8_t)m; w += 4; break; }
      case FS_ITEM: ft.items[sg.a].dst = w; w += ft.items[sg.a].n_out; break;
      case FS_SMALL: ft.smalls[sg.a].dst = w; w += ft.smalls[sg.a].n_out; break;
      default: { const uint64_t L = *ft.jobs[sg.a].total; ft.jobs[sg.a].dst = w; ft.jobs[sg.a].cap = L; w += L; break; }
    }
  }
}

}  // namespace b200tfs
