// Parameters shared by the cost-volume kernels (cost_volume.cu, cost_volume_tma.cu).
#pragma once
#include "common.cuh"

namespace pwc {

struct CvParams {
    const float* f0; const float* f1; const float* flow;
    float* out; float* f0_copy;
    int f0_cs, f1_cs, flow_cs, out_cs, f0_copy_cs;
    int B, H, W, C;
    float flow_scale, alpha, inv_c;
    int warp_type;   // 0 bilinear, 1 nearest
};

// TMA-pipelined r = 4 kernel (cost_volume_tma.cu).  Returns 0 on success, a cudaError_t on failure and
// CV_TMA_UNSUPPORTED when the arguments need the generic kernels of cost_volume.cu.
constexpr int CV_TMA_UNSUPPORTED = -1000;
int launch_cv_tma(const CvParams& p, cudaStream_t st);
// tcgen05 band-GEMM r = 4 kernel (cost_volume_tc.cu), same contract.
int launch_cv_tc(const CvParams& p, cudaStream_t st);

}  // namespace pwc
