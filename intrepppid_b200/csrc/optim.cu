// Multi-tensor AdamW: the optimizer step that follows the hot path (e2e/e2e_triplet.py:231-255, `AdamW(self.parameters(), lr)`).
// The reference's step is torch.optim.AdamW over 29 small tensors (23 with gradients); this is ONE launch over all of them,
// HBM-bound: 16 B read + 12 B written per parameter (p, g, m, v -> p, m, v).
//
// Arithmetic follows torch/optim/adamw.py `_single_tensor_adamw` (decoupled weight decay, no amsgrad), in fp32:
//   p *= 1 - lr*wd;  m += (g - m)(1 - b1);  v = v*b2 + (1 - b2) g*g;  p -= lr/(1 - b1^t) * m / (sqrt(v)/sqrt(1 - b2^t) + eps)
#include "kernels.h"

namespace ib200 {
namespace {

constexpr int kAdamThreads = 256;
constexpr int kAdamPerThread = 8;
constexpr int kAdamChunk = kAdamThreads * kAdamPerThread;

struct AdamTable {
  float* p[kAdamMaxTensors];
  const float* g[kAdamMaxTensors];
  float* m[kAdamMaxTensors];
  float* v[kAdamMaxTensors];
  long long numel[kAdamMaxTensors];
  int chunk_start[kAdamMaxTensors + 1];  // first block of tensor k; [n] = grid size
  int n;
};

__global__ void __launch_bounds__(kAdamThreads) adamw_kernel(const __grid_constant__ AdamTable tb, const AdamScalars s) {
  // block -> (tensor, chunk): binary search over at most 32 prefix entries held in the kernel parameters
  int lo = 0, hi = tb.n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tb.chunk_start[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
  }
  const long long base = (long long)((int)blockIdx.x - tb.chunk_start[lo]) * kAdamChunk;
  const long long n = tb.numel[lo];
  float* __restrict__ P = tb.p[lo];
  const float* __restrict__ G = tb.g[lo];
  float* __restrict__ M = tb.m[lo];
  float* __restrict__ V = tb.v[lo];
#pragma unroll
  for (int j = 0; j < kAdamPerThread; ++j) {
    const long long i = base + j * kAdamThreads + threadIdx.x;  // consecutive threads -> consecutive words: coalesced
    if (i >= n) break;
    float g = G[i] * s.grad_scale;
    float p = P[i] * s.decay;
    float m = M[i], v = V[i];
    m = fmaf(g - m, s.one_minus_b1, m);
    v = fmaf(s.one_minus_b2 * g, g, v * s.b2);
    const float denom = sqrtf(v) / s.bc2_sqrt + s.eps;
    p = fmaf(-s.step_size, m / denom, p);
    P[i] = p; M[i] = m; V[i] = v;
  }
}

}  // namespace

cudaError_t launch_adamw(int n, float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                         const long long* numel, const AdamScalars& s, cudaStream_t st, int* launches) {
  *launches = 0;
  for (int k0 = 0; k0 < n;) {
    AdamTable tb{};
    int cnt = 0, blocks = 0;
    for (; k0 < n && cnt < kAdamMaxTensors; ++k0) {
      if (numel[k0] <= 0 || grads[k0] == nullptr) continue;  // a parameter without a gradient is skipped, as torch skips p.grad is None
      tb.p[cnt] = params[k0]; tb.g[cnt] = grads[k0]; tb.m[cnt] = exp_avg[k0]; tb.v[cnt] = exp_avg_sq[k0];
      tb.numel[cnt] = numel[k0];
      tb.chunk_start[cnt] = blocks;
      blocks += (int)((numel[k0] + kAdamChunk - 1) / kAdamChunk);
      ++cnt;
    }
    if (cnt == 0) continue;
    tb.chunk_start[cnt] = blocks;
    tb.n = cnt;
    adamw_kernel<<<blocks, kAdamThreads, 0, st>>>(tb, s);
    ++*launches;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace ib200
