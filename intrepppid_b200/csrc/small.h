// Launchers for the small kernels that the C ABI exposes directly.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ib200.h"

namespace ib200 {
cudaError_t launch_pool_fc_fwd(int N, int H, int mode, const float* hn, const float* fc_w, const float* fc_b, float* z,
                               float* pooled_out, uint8_t* argmax_out, cudaStream_t st);
cudaError_t launch_pool_fc_bwd(int N, int H, int mode, const float* dz, const float* pooled, const uint8_t* argmax,
                               const float* fc_w, float* d_hn, float* d_fc_w, float* d_fc_b, cudaStream_t st);
cudaError_t launch_loss_head_fwd(int B, int H, float beta, const float* z, const long long* y, const ib200_head_params& hp,
                                 const ib200_head_masks& hm, float* losses, float* y_hat, cudaStream_t st);
cudaError_t launch_loss_head_bwd(int B, int H, float beta, const float* z, const long long* y, const ib200_head_params& hp,
                                 const ib200_head_masks& hm, const float* d_loss, const float* d_y_hat, float* dz,
                                 const ib200_head_grads& hg,
                                 cudaStream_t st);
bool head_supports(int H);  // loss/head/pair kernels: H multiple of 32 in [32, 256]
// P pairs: explicit (idx_a, idx_b), or -- both null -- the flat upper-triangle indices [p0, p0 + P)
cudaError_t launch_pair_score(int M, int H, const float* z, const int* idx_a, const int* idx_b, long long P, long long p0,
                              const ib200_head_params& hp, float* prob, cudaStream_t st);
}  // namespace ib200
