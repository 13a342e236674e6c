// kernels.h - launchers exported by kernels.cu to the host side of the library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200tfs.h"
#include "plan.h"

namespace b200tfs {

// Launch the pack/unpack engine.  plan_dev == nullptr: the plan image (plan_bytes <= kInlinePlanBytes)
// travels in the kernel parameters; otherwise it has already been copied to plan_dev on `stream`.
cudaError_t launch_move(const uint8_t* plan_dev, const uint8_t* plan_host, uint32_t plan_bytes, uint32_t n_tiles,
                        uint32_t n_small, cudaStream_t stream);

// the same engine over a device-resident plan whose items are stored only where PlanHeader::guard says so (codec_host.cpp: the
// narrowing batch decode)
cudaError_t launch_move_guarded(const uint8_t* plan_dev, uint32_t n_tiles, cudaStream_t stream);

// spill: n * spill_per_rec entries of kSpillEntryBytes (walker.h SpillEntry) for dims / value runs beyond the table's inline
// arrays; spill_used[r] receives how many entries record r wanted (more than spill_per_rec: its status is B200TFS_E_SPILL)
constexpr uint32_t kSpillEntryBytes = 32;
cudaError_t launch_parse_responses(const uint8_t* w, const uint64_t* rec_off, const uint64_t* rec_len, int n, int max_outputs,
                                   b200tfs_output* outs, int32_t* n_outs, b200tfs_model_spec* specs, int32_t* status,
                                   void* spill, uint32_t spill_per_rec, uint32_t* spill_used, cudaStream_t stream);
cudaError_t launch_parse_tensors(const uint8_t* w, const uint64_t* rec_off, const uint64_t* rec_len, int n, b200tfs_output* outs,
                                 int32_t* status, void* spill, uint32_t spill_per_rec, uint32_t* spill_used, cudaStream_t stream);

cudaError_t launch_decode_fused(const FusedParams& fp, uint32_t grid, cudaStream_t stream);
constexpr uint32_t kStageVecsHost = 2048;   // == kStageVecs (kernels.cu): launches with fatter tiles run the TMA-staged batch kernel
uint32_t tiles_for_host(uint64_t n_out, uint32_t vpt);
// pad elements [have, n_elems) of dst with element have-1 (zeros if have == 0); have from the host or, if have_dev, the device
cudaError_t launch_fill_edge(uint8_t* dst, uint32_t elem_size, uint64_t have, const unsigned long long* have_dev, uint64_t n_elems,
                             cudaStream_t stream);
// packed varints (varint_kernels.cuh): the tables and counters every kernel uses are addressed through VarTables
cudaError_t launch_frame_requests(const FrameTables& ft, cudaStream_t stream);
cudaError_t launch_venc_len(const VarTables& tb, cudaStream_t stream);
cudaError_t launch_venc_emit(const VarTables& tb, cudaStream_t stream);
cudaError_t launch_vdec_fused(const VarTables& tb, const VarFuse& fz, cudaStream_t stream);   // single-pass decode: counters zeroed beforehand
cudaError_t launch_venc_fused(const VarTables& tb, const VarFuse& fz, cudaStream_t stream);   // single-pass: counters zeroed beforehand
cudaError_t launch_vdec_count(const VarTables& tb, cudaStream_t stream);
cudaError_t launch_vdec_emit(const VarTables& tb, cudaStream_t stream);

}  // namespace b200tfs
