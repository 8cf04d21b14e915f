// Multi-tensor Ranger21: the optimizer step the reference's factory default selects (e2e/e2e_triplet.py:200-226,
// `Ranger21(self.parameters(), lr, weight_decay=1e-2, use_warmup, warmdown_active, ...)`; intrepppid/__init__.py:37
// optimizer_type="ranger21_xx").  The third-party package walks every parameter twice with ~40 small torch kernels each and one
// host sync per step (math.sqrt of a device scalar): 6 ms per step on the headline network, more than the whole training step on
// these kernels; this is TWO launches for all tensors and no sync.
//
// Ranger21 is pinned third-party code absent from the image (requirements.txt:65): the arithmetic follows the published
// algorithm (arXiv:2106.13731) in the step order oracle/ranger21_restated.py spells out -- PARITY UNPINNED against the package.
//
//   phase 1 (one CTA per tensor: every reduction of the step lives here).  The gradient is staged in SHARED MEMORY (tensors up to
//            48 K elements; larger ones are worked on in place in global memory through the same code), so its six passes cost one
//            global read and one global write:
//              rows:   adaptive gradient clipping (unit-wise norms of p and g), centralization; ||p_row|| is kept for phase 2
//              tensor: std -> normalized gradient g1; variance_ma = b2 variance_ma + (1-b2) g1^2; partial[k] = sum(variance_ma)/(1-b2^t)
//              rows + tensor: the SECOND centralization + normalization the package applies before the momentum update -> g2,
//                      written back as the gradient (what the package leaves in p.grad)
//            the last CTA to finish adds the partial sums in index order: variance_normalized = sqrt(sum_k partial[k] / elements)
//   phase 2 (elementwise over all tensors, 2048 elements per CTA):  stable weight decay, norm loss from the kept row norms
//            (||decay p_row|| = decay ||p_row||), positive-negative momentum, softplus denominator, update, lookahead merge.
// Reductions: fp32 inside a row, double across a tensor; fixed order (deterministic).
#include "kernels.h"

namespace ib200 {
namespace {

constexpr int kR21Threads = 1024;  // phase 1: one CTA per tensor
constexpr int kR21Warps = kR21Threads / 32;
constexpr int kR21StageFloats = 48 * 1024;  // gradient staging buffer in shared memory (192 KB)
constexpr int kR21P2Threads = 256, kR21P2PerThread = 8, kR21P2Chunk = kR21P2Threads * kR21P2PerThread;  // phase 2: elementwise chunks

struct R21Table {
  R21Tensor t[kR21MaxTensors];
};
struct R21Table2 {
  R21Tensor t[kR21MaxTensors];
  int chunk_start[kR21MaxTensors + 1];  // first CTA of tensor k; [n] = grid size
  int n;
};

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
// sum over the CTA, the same value in every thread (fixed order: deterministic)
__device__ __forceinline__ double block_sum(double x, double* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  x = warp_sum(x);
  __syncthreads();  // red[] may still be read from the previous call; also orders the caller's writes before the next pass
  if (lane == 0) red[warp] = x;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < kR21Warps; ++w) s += red[w];
  return s;
}

// whole-tensor unbiased standard deviation of X (two passes), as x.std()
__device__ __forceinline__ float tensor_std(const float* X, long long n, double* red) {
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += kR21Threads) s += (double)X[i];
  const double mean = block_sum(s, red) / (double)n;
  double q = 0.0;
  for (long long i = threadIdx.x; i < n; i += kR21Threads) {
    const double d = (double)X[i] - mean;
    q += d * d;
  }
  return (float)sqrt(block_sum(q, red) / (double)(n - 1));
}

// AGC scale of one row from its squared norms (1 = not clipped); the row norm of p is kept for the norm loss of phase 2
__device__ __forceinline__ float agc_scale(float sp, float sg, const R21Scalars& s, bool& clip) {
  const float pn = fmaxf(sqrtf(sp), s.agc_eps), gn = sqrtf(sg), maxn = pn * s.agc_clip;
  clip = s.use_agc && gn > maxn;
  return clip ? maxn / fmaxf(gn, 1e-6f) : 1.0f;
}

__global__ void __launch_bounds__(kR21Threads, 1) r21_phase1_kernel(const __grid_constant__ R21Table tb, const R21Scalars s,
                                                                   double* scratch, int slot0, int n_total) {
  double* partial = scratch + 3;  // scratch: [0] variance_normalized, [1] its inverse, [2] arrival counter, [3 ..] per-tensor sums
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* red = reinterpret_cast<double*>(smem_raw);                  // [kR21Warps]
  float* stage = reinterpret_cast<float*>(smem_raw + kR21Warps * 8);  // [kR21StageFloats]
  const R21Tensor& t = tb.t[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* G = t.g;
  const float* __restrict__ P = t.p;
  const int cols = t.cols, rows = t.rows;
  const long long n = t.numel;
  const bool centralize = s.use_gc && t.multi_dim;
  const bool staged = n <= kR21StageFloats;
  float* X = staged ? stage : G;  // working copy of the gradient
  // the optimizer state is DRAM-cold every step (a training step streams GBs through the L2 in between): start its lines towards
  // the L2 now, under the row pass -- variance_ma for the loop below, the momentum buffers (and slow weights) for phase 2
  for (long long i = (long long)threadIdx.x * 32; i < n; i += (long long)kR21Threads * 32) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(t.v + i));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(t.grad_ma + i));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(t.neg_grad_ma + i));
    if (s.lookahead_merge) asm volatile("prefetch.global.L2 [%0];" ::"l"(t.slow + i));
  }

  // ---- rows: AGC + centralization (a 0-d / 1-d tensor is one row: whole-tensor norm, never centralized) ----------------------
  if (cols <= 128) {
    // four rows per warp at a time, held in registers: one global round trip per four rows
    for (int r0 = warp * 4; r0 < rows; r0 += kR21Warps * 4) {
      float p[4][4], g[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = r0 + u, c = lane + 32 * e;
          const bool ok = r < rows && c < cols;
          p[u][e] = ok ? P[(size_t)r * cols + c] : 0.f;
          g[u][e] = ok ? G[(size_t)r * cols + c] : 0.f;
        }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u;
        if (r >= rows) break;  // warp-uniform
        float sp = 0.f, sg = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          sp = fmaf(p[u][e], p[u][e], sp);
          sg = fmaf(g[u][e], g[u][e], sg);
        }
        sp = warp_sum(sp);
        sg = warp_sum(sg);
        if (lane == 0) t.pnorm[r] = sqrtf(sp);
        bool clip;
        const float scale = agc_scale(sp, sg, s, clip);
        float sum = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (clip) g[u][e] *= scale;
          sum += g[u][e];
        }
        const float mean = centralize ? warp_sum(sum) / (float)cols : 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = lane + 32 * e;
          if (c < cols) X[(size_t)r * cols + c] = g[u][e] - mean;
        }
      }
    }
  } else {
    for (int r = warp; r < rows; r += kR21Warps) {
      const size_t base = (size_t)r * cols;
      float sp = 0.f, sg = 0.f;
#pragma unroll 4
      for (int c = lane; c < cols; c += 32) {
        const float pv = P[base + c], gv = G[base + c];
        sp = fmaf(pv, pv, sp);
        sg = fmaf(gv, gv, sg);
        if (staged) X[base + c] = gv;
      }
      sp = warp_sum(sp);
      sg = warp_sum(sg);
      if (lane == 0) t.pnorm[r] = sqrtf(sp);
      bool clip;
      const float scale = agc_scale(sp, sg, s, clip);
      if (centralize) {  // (each lane re-reads only what it wrote itself)
        float sum = 0.f;
        for (int c = lane; c < cols; c += 32) sum += clip ? X[base + c] * scale : X[base + c];
        const float mean = warp_sum(sum) / (float)cols;
        for (int c = lane; c < cols; c += 32) X[base + c] = (clip ? X[base + c] * scale : X[base + c]) - mean;
      } else if (clip) {
        for (int c = lane; c < cols; c += 32) X[base + c] *= scale;
      }
    }
  }
  __syncthreads();

  // ---- tensor: normalization by the std -> g1; variance_ma and its debiased sum ------------------------------------------------
  const bool norm = s.use_gcnorm && n > 2;
  // (x * (1/s) instead of x / s, here and below: within one ulp of the division, and the IEEE division sequence takes its slow path
  //  on the exactly-zero gradients this network is full of -- unused vocabulary rows, the dead top-layer chain)
  const float inv1 = norm ? 1.0f / (tensor_std(X, n, red) + 1e-8f) : 1.0f;
  float* __restrict__ V = t.v;
  double acc = 0.0;
#pragma unroll 8
  for (long long i = threadIdx.x; i < n; i += kR21Threads) {
    float gv = X[i];
    if (norm) {
      gv = gv * inv1;
      X[i] = gv;
    }
    const float v = fmaf(s.one_minus_b2 * gv, gv, V[i] * s.b2);
    V[i] = v;
    acc += (double)v;
  }
  const double tot = block_sum(acc, red);  // (its barriers also order the X writes above before the row pass below)
  if (threadIdx.x == 0) {
    partial[slot0 + blockIdx.x] = tot * t.inv_bc2;
    // the LAST CTA of the step (over all phase-1 launches) adds the partial sums in index order and publishes
    // 1 / variance_normalized for phase 2: no thread of phase 2 repeats the sum, nobody waits on anybody
    __threadfence();
    unsigned* ticket = reinterpret_cast<unsigned*>(scratch + 2);
    if (atomicAdd(ticket, 1u) == (unsigned)n_total - 1u) {
      __threadfence();
      double vsum = 0.0;
      for (int k = 0; k < n_total; ++k) vsum += *reinterpret_cast<volatile double*>(partial + k);
      const double vn = sqrt(vsum / s.param_size);
      scratch[0] = vn;                  // NaN here = the package's "hit nan for variance_normalized"
      scratch[1] = 1.0 / vn;
      *ticket = 0u;                     // ready for the next step (stream order)
    }
  }

  // ---- the second centralization + normalization (applied by the package right before the momentum update) -> g2 -------------
  if (centralize) {
    for (int r = warp; r < rows; r += kR21Warps) {
      const size_t base = (size_t)r * cols;
      float sum = 0.f;
      for (int c = lane; c < cols; c += 32) sum += X[base + c];
      const float mean = warp_sum(sum) / (float)cols;
      for (int c = lane; c < cols; c += 32) X[base + c] -= mean;
    }
    __syncthreads();
  }
  const float inv2 = norm ? 1.0f / (tensor_std(X, n, red) + 1e-8f) : 1.0f;
  if (staged || norm) {
#pragma unroll 8
    for (long long i = threadIdx.x; i < n; i += kR21Threads) G[i] = norm ? X[i] * inv2 : X[i];
  }
}

__global__ void __launch_bounds__(kR21P2Threads) r21_phase2_kernel(const __grid_constant__ R21Table2 tb, const R21Scalars s,
                                                                  const double* __restrict__ inv_vn) {
  int lo = 0, hi = tb.n;  // CTA -> (tensor, chunk): binary search over the prefix table in the kernel parameters
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tb.chunk_start[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
  }
  const R21Tensor& t = tb.t[lo];
  const unsigned base = (unsigned)((int)blockIdx.x - tb.chunk_start[lo]) * kR21P2Chunk;
  const float decay = s.use_decay ? (float)(1.0 - t.wd_lr * inv_vn[0]) : 1.0f;

  float* __restrict__ P = t.p;
  const float* __restrict__ G = t.g;
  const float* __restrict__ V = t.v;
  float* __restrict__ M = t.grad_ma;
  const float* __restrict__ Mneg = t.neg_grad_ma;
  float* __restrict__ S = t.slow;
  const float* __restrict__ pnorm = t.pnorm;
  const unsigned n = (unsigned)t.numel, cols = (unsigned)t.cols;  // (numel < 2^31, checked by the C ABI)
  const float inv_sqrt_bc2 = 1.0f / t.sqrt_bc2, inv_beta = 1.0f / s.softplus_beta;
#pragma unroll
  for (int j = 0; j < kR21P2PerThread; ++j) {
    const unsigned i = base + j * kR21P2Threads + threadIdx.x;
    if (i >= n) break;
    float mul = 1.0f;
    if (s.use_normloss) {  // unit norm of the decayed row = |decay| * the norm phase 1 kept
      const float unorm = fabsf(decay) * pnorm[i / cols];
      mul = 1.0f - t.lr * (s.normloss2 * (1.0f - 1.0f / (unorm + s.eps)));
    }
    const float g = G[i];
    float denom = fmaf(sqrtf(V[i]), inv_sqrt_bc2, s.eps);
    const float m = fmaf(s.one_minus_b1sq, g, M[i] * s.b1sq);
    M[i] = m;
    if (s.use_softplus) {
      const float x = denom * s.softplus_beta;
      if (x <= 20.0f) denom = log1pf(expf(x)) * inv_beta;
    }
    const float pn = fmaf(-s.pnm_factor, Mneg[i], m * s.one_plus_pnm) * s.inv_noise_norm;
    float p = fmaf(-t.step_size, pn * __frcp_rn(denom), (P[i] * decay) * mul);
    if (s.lookahead_merge) {
      p = fmaf(S[i], s.one_minus_la_alpha, p * s.la_alpha);
      S[i] = p;
    }
    P[i] = p;
  }
}

}  // namespace

cudaError_t launch_ranger21(int n, const R21Tensor* tensors, const R21Scalars& s, double* scratch, cudaStream_t st, int* launches) {
  *launches = 0;
  constexpr int kSmem = kR21Warps * 8 + kR21StageFloats * 4;
  static cudaError_t attr = cudaFuncSetAttribute(r21_phase1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  if (attr != cudaSuccess) return attr;
  for (int k0 = 0; k0 < n; k0 += kR21MaxTensors) {
    const int cnt = n - k0 < kR21MaxTensors ? n - k0 : kR21MaxTensors;
    R21Table tb{};
    for (int k = 0; k < cnt; ++k) tb.t[k] = tensors[k0 + k];
    r21_phase1_kernel<<<cnt, kR21Threads, kSmem, st>>>(tb, s, scratch, k0, n);
    ++*launches;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  for (int k0 = 0; k0 < n; k0 += kR21MaxTensors) {
    const int cnt = n - k0 < kR21MaxTensors ? n - k0 : kR21MaxTensors;
    R21Table2 tb{};
    int blocks = 0;
    for (int k = 0; k < cnt; ++k) {
      tb.t[k] = tensors[k0 + k];
      tb.chunk_start[k] = blocks;
      blocks += (int)((tensors[k0 + k].numel + kR21P2Chunk - 1) / kR21P2Chunk);
    }
    tb.chunk_start[cnt] = blocks;
    tb.n = cnt;
    r21_phase2_kernel<<<blocks, kR21P2Threads, 0, st>>>(tb, s, scratch + 1);
    ++*launches;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace ib200
