// K5: fused triplet-projection + TripletMarginLoss + MLPHead + BCEWithLogits + beta mix, forward and backward, and the
// eval-mode pair scorer.  Reference: e2e/e2e_triplet.py:113-136 (step), :82-85 (projection), classifier/head/mlp.py:35-68.
// These tensors are tiny ([5,B,H], B=80): one CTA, one warp per sample, deterministic reductions (no atomics).
#include "kernels.h"
#include "small.h"

namespace ib200 {
namespace {

__device__ __forceinline__ float mish_acc(float x) {  // x * tanh(softplus(x)) = x * n/(n+2), n = e^x (e^x + 2)
  if (x > 20.0f) return x;
  const float e = expf(x), n = e * (e + 2.0f);
  return x * n / (n + 2.0f);
}
__device__ __forceinline__ float mish_grad_acc(float x) {
  if (x > 20.0f) return 1.0f;
  const float e = expf(x), n = e * (e + 2.0f), d = n + 2.0f;
  return n / d + x * 4.0f * e * (e + 1.0f) / (d * d);
}

constexpr int kHeadWarps = 8;
constexpr int kChunk = 32;  // samples per backward chunk
constexpr float kEps = 1e-6f;  // nn.TripletMarginLoss eps (pairwise_distance adds it to the difference)

template <int H>
struct HeadSmem {
  static constexpr int HH = H / 2;
  float w1[HH][H + 1];           // fc1 weight * mask
  float wp[H][H + 1];            // projection weight
  float scratch[kHeadWarps][H];  // per-warp broadcast vector
  float red[kHeadWarps][4];
};

// projection a' = Wp mish(z) + bp for one sample held as z[e = lane + 32 i]
template <int H>
__device__ __forceinline__ void project(HeadSmem<H>& sm, int warp, int lane, const float* __restrict__ bp, float (&z)[H / 32],
                                        float (&out)[H / 32]) {
  constexpr int FPL = H / 32;
#pragma unroll
  for (int i = 0; i < FPL; ++i) sm.scratch[warp][lane + 32 * i] = mish_acc(z[i]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < FPL; ++i) {
    const int e = lane + 32 * i;
    float s = bp[e];
    for (int k = 0; k < H; ++k) s = fmaf(sm.wp[e][k], sm.scratch[warp][k], s);
    out[i] = s;
  }
  __syncwarp();
}

template <int H>
__device__ __forceinline__ void load_head_weights(HeadSmem<H>& sm, const ib200_head_params& hp, const ib200_head_masks& hm) {
  constexpr int HH = H / 2;
  for (int i = threadIdx.x; i < HH * H; i += blockDim.x) {
    const float m = hm.fc1_w != nullptr ? hm.fc1_w[i] : 1.0f;
    sm.w1[i / H][i % H] = hp.fc1_w[i] * m;  // WeightDrop on fc1.weight (mlp.py:38-46, weightdrop.py:100-102)
  }
  if (hp.proj_w != nullptr)
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) sm.wp[i / H][i % H] = hp.proj_w[i];
}

// forward pieces of the head for one sample; returns the logit (valid in all lanes) and keeps intermediates for backward
template <int H>
struct HeadFwd {
  float x[H / 32];   // (z1+z2)/2
  float a1, d1;      // lane j < H/2: fc1 pre-activation, dropout-1 output
  float d2;          // dropout-2 output
  float logit;
};

template <int H>
__device__ __forceinline__ void head_forward(HeadSmem<H>& sm, int warp, int lane, int b, const float* __restrict__ z1,
                                             const float* __restrict__ z2, const ib200_head_params& hp,
                                             const ib200_head_masks& hm, HeadFwd<H>& o) {
  constexpr int FPL = H / 32, HH = H / 2;
#pragma unroll
  for (int i = 0; i < FPL; ++i) {
    const int e = lane + 32 * i;
    o.x[i] = (z1[e] + z2[e]) * 0.5f;                 // mlp.py:66
    sm.scratch[warp][e] = mish_acc(o.x[i]);          // nl0
  }
  __syncwarp();
  float contrib = 0.f;
  o.a1 = o.d1 = o.d2 = 0.f;
  if (lane < HH) {
    float s = hp.fc1_b[lane];
    for (int k = 0; k < H; ++k) s = fmaf(sm.w1[lane][k], sm.scratch[warp][k], s);
    o.a1 = s;
    const float m1 = mish_acc(s);                                                    // nl1
    o.d1 = m1 * (hm.do1 != nullptr ? hm.do1[(size_t)b * HH + lane] : 1.0f);           // do1
    const float m2 = mish_acc(o.d1);                                                 // nl2
    o.d2 = m2 * (hm.do2 != nullptr ? hm.do2[(size_t)b * HH + lane] : 1.0f);           // do2
    const float w2 = hp.fc2_w[lane] * (hm.fc2_w != nullptr ? hm.fc2_w[lane] : 1.0f);  // WeightDrop on fc2.weight
    contrib = w2 * o.d2;
  }
  __syncwarp();
  o.logit = warp_sum(contrib) + hp.fc2_b[0];
}

template <int H>
__global__ void __launch_bounds__(kHeadWarps * 32) loss_head_fwd_kernel(int B, float beta, const float* __restrict__ z,
                                                                         const long long* __restrict__ y, ib200_head_params hp,
                                                                         ib200_head_masks hm, float* __restrict__ losses,
                                                                         float* __restrict__ y_hat) {
  constexpr int FPL = H / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HeadSmem<H>& sm = *reinterpret_cast<HeadSmem<H>*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  load_head_weights<H>(sm, hp, hm);
  __syncthreads();
  const bool proj = hp.proj_w != nullptr;
  float trip_sum = 0.f, bce_sum = 0.f;
  for (int b = warp; b < B; b += kHeadWarps) {
    float v[3][FPL];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float zr[FPL];
#pragma unroll
      for (int i = 0; i < FPL; ++i) zr[i] = z[((size_t)r * B + b) * H + lane + 32 * i];
      if (proj) project<H>(sm, warp, lane, hp.proj_b, zr, v[r]);
      else {
#pragma unroll
        for (int i = 0; i < FPL; ++i) v[r][i] = zr[i];
      }
    }
    float sap = 0.f, san = 0.f;
#pragma unroll
    for (int i = 0; i < FPL; ++i) {
      const float dp = v[0][i] - v[1][i] + kEps, dn = v[0][i] - v[2][i] + kEps;
      sap = fmaf(dp, dp, sap);
      san = fmaf(dn, dn, san);
    }
    const float dap = sqrtf(warp_sum(sap)), dan = sqrtf(warp_sum(san));
    trip_sum += fmaxf(dap - dan + 1.0f, 0.0f);  // margin = 1.0, p = 2 (e2e_triplet.py:80)

    HeadFwd<H> hf;
    head_forward<H>(sm, warp, lane, b, z + ((size_t)3 * B + b) * H, z + ((size_t)4 * B + b) * H, hp, hm, hf);
    const float xl = hf.logit, yb = (float)y[b];
    bce_sum += fmaxf(xl, 0.0f) - xl * yb + log1pf(expf(-fabsf(xl)));  // BCEWithLogitsLoss (e2e_triplet.py:76)
    if (lane == 0) y_hat[b] = xl;
  }
  if (lane == 0) {
    sm.red[warp][0] = trip_sum;
    sm.red[warp][1] = bce_sum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f, c = 0.f;
    for (int w = 0; w < kHeadWarps; ++w) {
      t += sm.red[w][0];
      c += sm.red[w][1];
    }
    t /= (float)B;
    c /= (float)B;
    const float nb_ssl = 1.0f / beta, nb_cls = 1.0f - nb_ssl;  // e2e_triplet.py:133-136 (code is authoritative)
    losses[0] = nb_cls * c + nb_ssl * t;
    losses[1] = c;
    losses[2] = t;
  }
}

template <int H>
struct HeadBwdSmem {
  static constexpr int HH = H / 2;
  HeadSmem<H> f;
  float dlt1[kChunk][HH];        // delta at fc1 pre-activation
  float m0s[kChunk][H];          // mish((z1+z2)/2)
  float dltp[3][kChunk][H];      // gradient at the projection output per role
  float mzs[3][kChunk][H];       // mish(z_role)
  float vec[kHeadWarps][4 * H];  // per-warp partial vector grads: db1[HH] | dw2[HH] | dbp[H] | (db2 at [2H])
};

template <int H>
__global__ void __launch_bounds__(kHeadWarps * 32) loss_head_bwd_kernel(int B, float beta, const float* __restrict__ z,
                                                                         const long long* __restrict__ y, ib200_head_params hp,
                                                                         ib200_head_masks hm, const float* __restrict__ d_loss,
                                                                         const float* __restrict__ d_y_hat, float* __restrict__ dz,
                                                                         ib200_head_grads hg) {
  constexpr int FPL = H / 32, HH = H / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HeadBwdSmem<H>& sm = *reinterpret_cast<HeadBwdSmem<H>*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  load_head_weights<H>(sm.f, hp, hm);
  for (int i = tid; i < kHeadWarps * 4 * H; i += blockDim.x) (&sm.vec[0][0])[i] = 0.f;
  __syncthreads();
  const bool proj = hp.proj_w != nullptr;
  const float dL = d_loss[0];
  const float w_t = dL * (1.0f / beta) / (float)B, w_c = dL * (1.0f - 1.0f / beta) / (float)B;

  // matrix-gradient accumulators owned by this thread: fc1 [HH*H/256], proj [H*H/256]
  constexpr int N1 = (HH * H + 255) / 256, NP = (H * H + 255) / 256;
  float acc1[N1], accp[NP];
#pragma unroll
  for (int i = 0; i < N1; ++i) acc1[i] = 0.f;
#pragma unroll
  for (int i = 0; i < NP; ++i) accp[i] = 0.f;
  float db1 = 0.f, dw2 = 0.f, db2 = 0.f, dbp[FPL];
#pragma unroll
  for (int i = 0; i < FPL; ++i) dbp[i] = 0.f;

  for (int c0 = 0; c0 < B; c0 += kChunk) {
    const int cn = min(kChunk, B - c0);
    // ---- phase A: one warp per sample ------------------------------------------------------------------------------------
    for (int c = warp; c < cn; c += kHeadWarps) {
      const int b = c0 + c;
      // triplet branch
      float zr[3][FPL], v[3][FPL];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int i = 0; i < FPL; ++i) zr[r][i] = z[((size_t)r * B + b) * H + lane + 32 * i];
        if (proj) {
          project<H>(sm.f, warp, lane, hp.proj_b, zr[r], v[r]);
#pragma unroll
          for (int i = 0; i < FPL; ++i) sm.mzs[r][c][lane + 32 * i] = mish_acc(zr[r][i]);
        } else {
#pragma unroll
          for (int i = 0; i < FPL; ++i) v[r][i] = zr[r][i];
        }
      }
      float sap = 0.f, san = 0.f;
#pragma unroll
      for (int i = 0; i < FPL; ++i) {
        const float dp = v[0][i] - v[1][i] + kEps, dn = v[0][i] - v[2][i] + kEps;
        sap = fmaf(dp, dp, sap);
        san = fmaf(dn, dn, san);
      }
      const float dap = sqrtf(warp_sum(sap)), dan = sqrtf(warp_sum(san));
      const float act = (dap - dan + 1.0f > 0.0f) ? w_t : 0.0f;
      const float iap = dap > 0.f ? act / dap : 0.f, ian = dan > 0.f ? act / dan : 0.f;
      float gv[3][FPL];
#pragma unroll
      for (int i = 0; i < FPL; ++i) {
        const float dp = v[0][i] - v[1][i] + kEps, dn = v[0][i] - v[2][i] + kEps;
        gv[0][i] = dp * iap - dn * ian;
        gv[1][i] = -dp * iap;
        gv[2][i] = dn * ian;
      }
      if (proj) {
        // back through Linear(H,H) and Mish: dz = (Wp^T g) * mish'(z)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
          for (int i = 0; i < FPL; ++i) {
            sm.dltp[r][c][lane + 32 * i] = gv[r][i];
            dbp[i] += gv[r][i];
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < FPL; ++i) {
            const int k = lane + 32 * i;
            float s = 0.f;
            for (int e = 0; e < H; ++e) s = fmaf(sm.f.wp[e][k], sm.dltp[r][c][e], s);
            dz[((size_t)r * B + b) * H + k] = s * mish_grad_acc(zr[r][i]);
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int i = 0; i < FPL; ++i) dz[((size_t)r * B + b) * H + lane + 32 * i] = gv[r][i];
      }
      // classifier branch
      HeadFwd<H> hf;
      head_forward<H>(sm.f, warp, lane, b, z + ((size_t)3 * B + b) * H, z + ((size_t)4 * B + b) * H, hp, hm, hf);
#pragma unroll
      for (int i = 0; i < FPL; ++i) sm.m0s[c][lane + 32 * i] = sm.f.scratch[warp][lane + 32 * i];
      const float dlogit = w_c * (1.0f / (1.0f + expf(-hf.logit)) - (float)y[b]) + (d_y_hat != nullptr ? d_y_hat[b] : 0.0f);
      float delta1 = 0.f;
      if (lane < HH) {
        const float m2mask = hm.fc2_w != nullptr ? hm.fc2_w[lane] : 1.0f;
        const float dd2 = dlogit * hp.fc2_w[lane] * m2mask;
        dw2 += dlogit * hf.d2 * m2mask;
        const float dm2 = dd2 * (hm.do2 != nullptr ? hm.do2[(size_t)b * HH + lane] : 1.0f);
        const float dd1 = dm2 * mish_grad_acc(hf.d1);
        const float dm1 = dd1 * (hm.do1 != nullptr ? hm.do1[(size_t)b * HH + lane] : 1.0f);
        delta1 = dm1 * mish_grad_acc(hf.a1);
        db1 += delta1;
        sm.dlt1[c][lane] = delta1;
      }
      if (lane == 0) db2 += dlogit;
      __syncwarp();
#pragma unroll
      for (int i = 0; i < FPL; ++i) {
        const int k = lane + 32 * i;
        float s = 0.f;
        for (int jj = 0; jj < HH; ++jj) s = fmaf(sm.f.w1[jj][k], sm.dlt1[c][jj], s);
        const float dx = s * mish_grad_acc(hf.x[i]) * 0.5f;
        dz[((size_t)3 * B + b) * H + k] = dx;
        dz[((size_t)4 * B + b) * H + k] = dx;
      }
      __syncwarp();
    }
    __syncthreads();
    // ---- phase B: every thread reduces its matrix entries over the chunk (fixed order => deterministic) ----------------------
#pragma unroll
    for (int i = 0; i < N1; ++i) {
      const int idx = tid + 256 * i;
      if (idx < HH * H) {
        const int jj = idx / H, k = idx % H;
        float s = acc1[i];
        for (int c = 0; c < cn; ++c) s = fmaf(sm.dlt1[c][jj], sm.m0s[c][k], s);
        acc1[i] = s;
      }
    }
    if (proj) {
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int idx = tid + 256 * i;
        if (idx < H * H) {
          const int e = idx / H, k = idx % H;
          float s = accp[i];
          for (int r = 0; r < 3; ++r)
            for (int c = 0; c < cn; ++c) s = fmaf(sm.dltp[r][c][e], sm.mzs[r][c][k], s);
          accp[i] = s;
        }
      }
    }
    __syncthreads();
  }

  // ---- write parameter gradients -------------------------------------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < N1; ++i) {
    const int idx = tid + 256 * i;
    if (idx < HH * H) hg.fc1_w[idx] = acc1[i] * (hm.fc1_w != nullptr ? hm.fc1_w[idx] : 1.0f);  // grad of weight_raw
  }
  if (proj && hg.proj_w != nullptr) {
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int idx = tid + 256 * i;
      if (idx < H * H) hg.proj_w[idx] = accp[i];
    }
  }
  if (lane < HH) {
    sm.vec[warp][lane] = db1;
    sm.vec[warp][HH + lane] = dw2;
  }
#pragma unroll
  for (int i = 0; i < FPL; ++i) sm.vec[warp][H + lane + 32 * i] = dbp[i];
  if (lane == 0) sm.vec[warp][2 * H] = db2;
  __syncthreads();
  for (int i = tid; i <= 2 * H; i += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < kHeadWarps; ++w) s += sm.vec[w][i];
    if (i < HH) hg.fc1_b[i] = s;
    else if (i < H) hg.fc2_w[i - HH] = s;
    else if (i < 2 * H) { if (proj && hg.proj_b != nullptr) hg.proj_b[i - H] = s; }
    else hg.fc2_b[0] = s;
  }
}

// ---- eval-mode head + sigmoid over pairs (one thread per pair) ------------------------------------------------------------------
__device__ __forceinline__ float mish_fast(float x) {
  if (x > 20.0f) return x;
  const float e = __expf(x), n = e * (e + 2.0f);
  return x * __fdividef(n, n + 2.0f);
}

template <int H>
__global__ void __launch_bounds__(128) pair_score_kernel(int M, const float* __restrict__ z, const int* __restrict__ ia,
                                                          const int* __restrict__ ib, long long P, ib200_head_params hp,
                                                          float* __restrict__ prob) {
  constexpr int HH = H / 2;
  __shared__ __align__(16) float w1[HH][H];
  __shared__ float b1[HH], w2[HH];
  for (int i = threadIdx.x; i < HH * H; i += blockDim.x) w1[i / H][i % H] = hp.fc1_w[i];
  for (int i = threadIdx.x; i < HH; i += blockDim.x) {
    b1[i] = hp.fc1_b[i];
    w2[i] = hp.fc2_w[i];
  }
  __syncthreads();
  const float b2 = hp.fc2_b[0];
  for (long long pidx = (long long)blockIdx.x * blockDim.x + threadIdx.x; pidx < P; pidx += (long long)gridDim.x * blockDim.x) {
    int i, j;
    if (ia != nullptr) {
      i = ia[pidx];
      j = ib[pidx];
    } else {
      // upper triangle, row-major: row i holds M-i pairs (i,i) .. (i,M-1); first index of row i = i*M - i(i-1)/2
      const double Md = (double)M;
      long long r = (long long)floor(((2.0 * Md + 1.0) - sqrt((2.0 * Md + 1.0) * (2.0 * Md + 1.0) - 8.0 * (double)pidx)) * 0.5);
      if (r < 0) r = 0;
      while (r * M - r * (r - 1) / 2 > pidx) --r;
      while ((r + 1) * M - (r + 1) * r / 2 <= pidx) ++r;
      i = (int)r;
      j = (int)(pidx - (r * M - r * (r - 1) / 2)) + i;
    }
    float m0[H];
    const float4* za = reinterpret_cast<const float4*>(z + (size_t)i * H);
    const float4* zb = reinterpret_cast<const float4*>(z + (size_t)j * H);
#pragma unroll
    for (int k = 0; k < H / 4; ++k) {
      const float4 a = __ldg(za + k), b = __ldg(zb + k);
      m0[4 * k + 0] = mish_fast((a.x + b.x) * 0.5f);
      m0[4 * k + 1] = mish_fast((a.y + b.y) * 0.5f);
      m0[4 * k + 2] = mish_fast((a.z + b.z) * 0.5f);
      m0[4 * k + 3] = mish_fast((a.w + b.w) * 0.5f);
    }
    float logit = b2;
#pragma unroll 4
    for (int jj = 0; jj < HH; ++jj) {
      float s = b1[jj];
#pragma unroll
      for (int k = 0; k < H / 4; ++k) {
        const float4 w = *reinterpret_cast<const float4*>(&w1[jj][4 * k]);
        s = fmaf(w.x, m0[4 * k], s);
        s = fmaf(w.y, m0[4 * k + 1], s);
        s = fmaf(w.z, m0[4 * k + 2], s);
        s = fmaf(w.w, m0[4 * k + 3], s);
      }
      logit = fmaf(w2[jj], mish_fast(mish_fast(s)), logit);  // nl1, (do1=id), nl2, (do2=id), fc2
    }
    prob[pidx] = __fdividef(1.0f, 1.0f + __expf(-logit));  // torch.sigmoid (cli/infer.py:224)
  }
}

}  // namespace

cudaError_t launch_loss_head_fwd(int B, int H, float beta, const float* z, const long long* y, const ib200_head_params& hp,
                                 const ib200_head_masks& hm, float* losses, float* y_hat, cudaStream_t st) {
  if (H == 64) {
    const size_t smem = sizeof(HeadSmem<64>);
    cudaError_t e = cudaFuncSetAttribute(loss_head_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    loss_head_fwd_kernel<64><<<1, kHeadWarps * 32, smem, st>>>(B, beta, z, y, hp, hm, losses, y_hat);
  } else if (H == 32) {
    const size_t smem = sizeof(HeadSmem<32>);
    loss_head_fwd_kernel<32><<<1, kHeadWarps * 32, smem, st>>>(B, beta, z, y, hp, hm, losses, y_hat);
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_loss_head_bwd(int B, int H, float beta, const float* z, const long long* y, const ib200_head_params& hp,
                                 const ib200_head_masks& hm, const float* d_loss, const float* d_y_hat, float* dz,
                                 const ib200_head_grads& hg, cudaStream_t st) {
  if (H == 64) {
    const size_t smem = sizeof(HeadBwdSmem<64>);
    cudaError_t e = cudaFuncSetAttribute(loss_head_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    loss_head_bwd_kernel<64><<<1, kHeadWarps * 32, smem, st>>>(B, beta, z, y, hp, hm, d_loss, d_y_hat, dz, hg);
  } else if (H == 32) {
    const size_t smem = sizeof(HeadBwdSmem<32>);
    cudaError_t e = cudaFuncSetAttribute(loss_head_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    loss_head_bwd_kernel<32><<<1, kHeadWarps * 32, smem, st>>>(B, beta, z, y, hp, hm, d_loss, d_y_hat, dz, hg);
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_pair_score(int M, int H, const float* z, const int* idx_a, const int* idx_b, long long P,
                              const ib200_head_params& hp, float* prob, cudaStream_t st) {
  if (P <= 0) return cudaSuccess;
  const unsigned grid = (unsigned)std::min<long long>((P + 127) / 128, 148LL * 16);
  if (H == 64) pair_score_kernel<64><<<grid, 128, 0, st>>>(M, z, idx_a, idx_b, P, hp, prob);
  else if (H == 32) pair_score_kernel<32><<<grid, 128, 0, st>>>(M, z, idx_a, idx_b, P, hp, prob);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace ib200
