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

cudaError_t launch_parse_responses(const uint8_t* w, const uint64_t* rec_off, const uint64_t* rec_len, int n, int max_outputs,
                                   b200tfs_output* outs, int32_t* n_outs, b200tfs_model_spec* specs, int32_t* status,
                                   cudaStream_t stream);
cudaError_t launch_parse_tensors(const uint8_t* w, const uint64_t* rec_off, const uint64_t* rec_len, int n, b200tfs_output* outs,
                                 int32_t* status, cudaStream_t stream);

cudaError_t launch_decode_fused(const FusedParams& fp, uint32_t grid, cudaStream_t stream);
uint32_t tiles_for_host(uint64_t n_out, uint32_t vpt);
cudaError_t launch_venc_len(const VarSeg* segs, const uint32_t* tile_seg, const VarJobDev* jobs, uint32_t* tile_val, uint32_t n_tiles,
                            cudaStream_t stream);
cudaError_t launch_vscan(const VarJobDev* jobs, const uint32_t* tile_val, uint64_t* tile_off, uint64_t* job_total, uint32_t n_jobs,
                         cudaStream_t stream);
cudaError_t launch_venc_emit(const VarSeg* segs, const uint32_t* tile_seg, const VarJobDev* jobs, const uint64_t* tile_off,
                             uint32_t n_tiles, cudaStream_t stream);
cudaError_t launch_vdec_count(const VarSeg* segs, const uint32_t* tile_seg, uint32_t* tile_val, uint32_t n_tiles, cudaStream_t stream);
cudaError_t launch_vdec_emit(const VarSeg* segs, const uint32_t* tile_seg, const VarJobDev* jobs, const uint64_t* tile_off,
                             const uint64_t* job_total, int32_t* job_status, uint32_t n_tiles, cudaStream_t stream);

}  // namespace b200tfs
