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

constexpr int kHeadWarps = 16;  // one warp per sample: 16 samples in flight in the single CTA
constexpr int kHeadThreads = kHeadWarps * 32;
constexpr float kEps = 1e-6f;  // nn.TripletMarginLoss eps (pairwise_distance adds it to the difference)
// samples per backward chunk; the wide variants (H > 64) keep the per-chunk vectors small enough for shared memory
template <int H>
constexpr int chunk_of() { return H <= 64 ? 32 : 8; }

// H <= 64: both weight matrices are staged in shared memory (masked fc1 weight, projection weight).  H > 64 (up to 256): they no
// longer fit, the kernels read them from global memory / L2 (a few hundred KB, read by one CTA) through the same accessors.
template <int H>
struct HeadSmem {
  static constexpr int HH = H / 2;
  static constexpr bool kStaged = H <= 64;
  float w1[kStaged ? HH : 1][H + 1];  // fc1 weight * mask
  float wp[kStaged ? H : 1][H + 1];   // projection weight
  float scratch[kHeadWarps][H];       // per-warp broadcast vector
  float red[kHeadWarps][4];
};
template <int H>
__device__ __forceinline__ float w1_at(const HeadSmem<H>& sm, const ib200_head_params& hp, const ib200_head_masks& hm, int j, int k) {
  if constexpr (HeadSmem<H>::kStaged) return sm.w1[j][k];
  else return hp.fc1_w[j * H + k] * (hm.fc1_w != nullptr ? hm.fc1_w[j * H + k] : 1.0f);
}
template <int H>
__device__ __forceinline__ float wp_at(const HeadSmem<H>& sm, const ib200_head_params& hp, int e, int k) {
  if constexpr (HeadSmem<H>::kStaged) return sm.wp[e][k];
  else return hp.proj_w[e * H + k];
}

// projection a' = Wp mish(z) + bp for one sample held as z[e = lane + 32 i]
template <int H>
__device__ __forceinline__ void project(HeadSmem<H>& sm, const ib200_head_params& hp, int warp, int lane, float (&z)[H / 32],
                                        float (&out)[H / 32]) {
  const float* __restrict__ bp = hp.proj_b;
  constexpr int FPL = H / 32;
#pragma unroll
  for (int i = 0; i < FPL; ++i) sm.scratch[warp][lane + 32 * i] = mish_acc(z[i]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < FPL; ++i) {
    const int e = lane + 32 * i;
    float s = bp[e];
    for (int k = 0; k < H; ++k) s = fmaf(wp_at<H>(sm, hp, e, k), sm.scratch[warp][k], s);
    out[i] = s;
  }
  __syncwarp();
}

template <int H>
__device__ __forceinline__ void load_head_weights(HeadSmem<H>& sm, const ib200_head_params& hp, const ib200_head_masks& hm) {
  constexpr int HH = H / 2;
  if constexpr (!HeadSmem<H>::kStaged) return;
#pragma unroll 4
  for (int i = threadIdx.x; i < HH * H; i += blockDim.x) {
    const float m = hm.fc1_w != nullptr ? hm.fc1_w[i] : 1.0f;
    sm.w1[i / H][i % H] = hp.fc1_w[i] * m;  // WeightDrop on fc1.weight (mlp.py:38-46, weightdrop.py:100-102)
  }
  if (hp.proj_w != nullptr) {
#pragma unroll 8
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) sm.wp[i / H][i % H] = hp.proj_w[i];
  }
}

// forward pieces of the head for one sample; returns the logit (valid in all lanes) and keeps intermediates for backward
template <int H>
struct HeadFwd {
  static constexpr int JPL = (H / 2 + 31) / 32;  // fc1 outputs per lane: j = lane + 32 m
  float x[H / 32];       // (z1+z2)/2
  float a1[JPL], d1[JPL];  // fc1 pre-activation, dropout-1 output
  float d2[JPL];         // dropout-2 output
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
#pragma unroll
  for (int m = 0; m < HeadFwd<H>::JPL; ++m) {
    const int j = lane + 32 * m;
    o.a1[m] = o.d1[m] = o.d2[m] = 0.f;
    if (j < HH) {
      float s = hp.fc1_b[j];
      for (int k = 0; k < H; ++k) s = fmaf(w1_at<H>(sm, hp, hm, j, k), sm.scratch[warp][k], s);
      o.a1[m] = s;
      const float m1 = mish_acc(s);                                                 // nl1
      o.d1[m] = m1 * (hm.do1 != nullptr ? hm.do1[(size_t)b * HH + j] : 1.0f);        // do1
      const float m2 = mish_acc(o.d1[m]);                                           // nl2
      o.d2[m] = m2 * (hm.do2 != nullptr ? hm.do2[(size_t)b * HH + j] : 1.0f);        // do2
      const float w2 = hp.fc2_w[j] * (hm.fc2_w != nullptr ? hm.fc2_w[j] : 1.0f);     // WeightDrop on fc2.weight
      contrib += w2 * o.d2[m];
    }
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
      if (proj) project<H>(sm, hp, warp, lane, zr, v[r]);
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
  static constexpr int kChunk = chunk_of<H>();
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
  constexpr int FPL = H / 32, HH = H / 2, kChunk = chunk_of<H>(), JPL = HeadFwd<H>::JPL;
  constexpr bool kRegAcc = H <= 64;  // matrix-gradient accumulators in registers; wider heads accumulate in the output buffers
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HeadBwdSmem<H>& sm = *reinterpret_cast<HeadBwdSmem<H>*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  load_head_weights<H>(sm.f, hp, hm);
  for (int i = tid; i < kHeadWarps * 4 * H; i += blockDim.x) (&sm.vec[0][0])[i] = 0.f;
  __syncthreads();
  const bool proj = hp.proj_w != nullptr;
  const float dL = d_loss[0];
  const float w_t = dL * (1.0f / beta) / (float)B, w_c = dL * (1.0f - 1.0f / beta) / (float)B;

  // matrix-gradient accumulators owned by this thread: fc1 [HH*H/threads], proj [H*H/threads]
  constexpr int N1 = kRegAcc ? (HH * H + kHeadThreads - 1) / kHeadThreads : 1, NP = kRegAcc ? (H * H + kHeadThreads - 1) / kHeadThreads : 1;
  float acc1[N1], accp[NP];
#pragma unroll
  for (int i = 0; i < N1; ++i) acc1[i] = 0.f;
#pragma unroll
  for (int i = 0; i < NP; ++i) accp[i] = 0.f;
  float db1[JPL], dw2[JPL], db2 = 0.f, dbp[FPL];
#pragma unroll
  for (int m = 0; m < JPL; ++m) db1[m] = dw2[m] = 0.f;
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
          project<H>(sm.f, hp, warp, lane, zr[r], v[r]);
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
            for (int e = 0; e < H; ++e) s = fmaf(wp_at<H>(sm.f, hp, e, k), sm.dltp[r][c][e], s);
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
#pragma unroll
      for (int m = 0; m < JPL; ++m) {
        const int j = lane + 32 * m;
        if (j < HH) {
          const float m2mask = hm.fc2_w != nullptr ? hm.fc2_w[j] : 1.0f;
          const float dd2 = dlogit * hp.fc2_w[j] * m2mask;
          dw2[m] += dlogit * hf.d2[m] * m2mask;
          const float dm2 = dd2 * (hm.do2 != nullptr ? hm.do2[(size_t)b * HH + j] : 1.0f);
          const float dd1 = dm2 * mish_grad_acc(hf.d1[m]);
          const float dm1 = dd1 * (hm.do1 != nullptr ? hm.do1[(size_t)b * HH + j] : 1.0f);
          const float delta1 = dm1 * mish_grad_acc(hf.a1[m]);
          db1[m] += delta1;
          sm.dlt1[c][j] = delta1;
        }
      }
      if (lane == 0) db2 += dlogit;
      __syncwarp();
#pragma unroll
      for (int i = 0; i < FPL; ++i) {
        const int k = lane + 32 * i;
        float s = 0.f;
        for (int jj = 0; jj < HH; ++jj) s = fmaf(w1_at<H>(sm.f, hp, hm, jj, k), sm.dlt1[c][jj], s);
        const float dx = s * mish_grad_acc(hf.x[i]) * 0.5f;
        dz[((size_t)3 * B + b) * H + k] = dx;
        dz[((size_t)4 * B + b) * H + k] = dx;
      }
      __syncwarp();
    }
    __syncthreads();
    // ---- phase B: every thread reduces its matrix entries over the chunk (fixed order => deterministic) ----------------------
    if constexpr (kRegAcc) {
#pragma unroll
      for (int i = 0; i < N1; ++i) {
        const int idx = tid + kHeadThreads * i;
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
          const int idx = tid + kHeadThreads * i;
          if (idx < H * H) {
            const int e = idx / H, k = idx % H;
            float s = accp[i];
            for (int r = 0; r < 3; ++r)
              for (int c = 0; c < cn; ++c) s = fmaf(sm.dltp[r][c][e], sm.mzs[r][c][k], s);
            accp[i] = s;
          }
        }
      }
    } else {
      // wide head: every matrix entry is owned by one thread of this single CTA and accumulated in the output buffer itself
      for (int idx = tid; idx < HH * H; idx += kHeadThreads) {
        const int jj = idx / H, k = idx % H;
        float s = c0 == 0 ? 0.f : hg.fc1_w[idx];
        for (int c = 0; c < cn; ++c) s = fmaf(sm.dlt1[c][jj], sm.m0s[c][k], s);
        hg.fc1_w[idx] = s;
      }
      if (proj && hg.proj_w != nullptr) {
        for (int idx = tid; idx < H * H; idx += kHeadThreads) {
          const int e = idx / H, k = idx % H;
          float s = c0 == 0 ? 0.f : hg.proj_w[idx];
          for (int r = 0; r < 3; ++r)
            for (int c = 0; c < cn; ++c) s = fmaf(sm.dltp[r][c][e], sm.mzs[r][c][k], s);
          hg.proj_w[idx] = s;
        }
      }
    }
    __syncthreads();
  }

  // ---- write parameter gradients -------------------------------------------------------------------------------------------
  if constexpr (kRegAcc) {
#pragma unroll
    for (int i = 0; i < N1; ++i) {
      const int idx = tid + kHeadThreads * i;
      if (idx < HH * H) hg.fc1_w[idx] = acc1[i] * (hm.fc1_w != nullptr ? hm.fc1_w[idx] : 1.0f);  // grad of weight_raw
    }
    if (proj && hg.proj_w != nullptr) {
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int idx = tid + kHeadThreads * i;
        if (idx < H * H) hg.proj_w[idx] = accp[i];
      }
    }
  } else if (hm.fc1_w != nullptr) {
    for (int idx = tid; idx < HH * H; idx += kHeadThreads) hg.fc1_w[idx] *= hm.fc1_w[idx];  // same owner thread as the accumulation
  }
#pragma unroll
  for (int m = 0; m < JPL; ++m) {
    const int j = lane + 32 * m;
    if (j < HH) {
      sm.vec[warp][j] = db1[m];
      sm.vec[warp][HH + j] = dw2[m];
    }
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
                                                          const int* __restrict__ ib, long long P, long long p0,
                                                          ib200_head_params hp, float* __restrict__ prob) {
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
      const long long pg = p0 + pidx;  // flat index inside the whole triangle
      long long r = (long long)floor(((2.0 * Md + 1.0) - sqrt((2.0 * Md + 1.0) * (2.0 * Md + 1.0) - 8.0 * (double)pg)) * 0.5);
      if (r < 0) r = 0;
      while (r * M - r * (r - 1) / 2 > pg) --r;
      while ((r + 1) * M - (r + 1) * r / 2 <= pg) ++r;
      i = (int)r;
      j = (int)(pg - (r * M - r * (r - 1) / 2)) + i;
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

// wide variant (H > 64): one WARP per pair; the fc1 weight sits transposed in shared memory ([k][j], conflict-free across lanes),
// the warp stages mish((z_a+z_b)/2) in its smem slot and every lane owns the fc1 outputs j = lane + 32 m.
template <int H>
__global__ void __launch_bounds__(256) pair_score_wide_kernel(int M, const float* __restrict__ z, const int* __restrict__ ia,
                                                               const int* __restrict__ ib, long long P, long long p0,
                                                               ib200_head_params hp, float* __restrict__ prob) {
  constexpr int HH = H / 2, JPL = (HH + 31) / 32, FPL = H / 32;
  extern __shared__ __align__(16) float smw[];
  float* w1t = smw;                  // [H][HH]
  float* m0 = smw + H * HH;          // [8 warps][H]
  for (int i = threadIdx.x; i < HH * H; i += blockDim.x) w1t[(i % H) * HH + i / H] = hp.fc1_w[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* mine = m0 + warp * H;
  float b1[JPL], w2[JPL];
#pragma unroll
  for (int m = 0; m < JPL; ++m) {
    const int j = lane + 32 * m;
    b1[m] = j < HH ? hp.fc1_b[j] : 0.f;
    w2[m] = j < HH ? hp.fc2_w[j] : 0.f;
  }
  const float b2 = hp.fc2_b[0];
  const long long nw = (long long)gridDim.x * 8;
  for (long long pidx = (long long)blockIdx.x * 8 + warp; pidx < P; pidx += nw) {
    int i, j;
    if (ia != nullptr) {
      i = ia[pidx];
      j = ib[pidx];
    } else {
      const double Md = (double)M;
      const long long pg = p0 + pidx;
      long long r = (long long)floor(((2.0 * Md + 1.0) - sqrt((2.0 * Md + 1.0) * (2.0 * Md + 1.0) - 8.0 * (double)pg)) * 0.5);
      if (r < 0) r = 0;
      while (r * M - r * (r - 1) / 2 > pg) --r;
      while ((r + 1) * M - (r + 1) * r / 2 <= pg) ++r;
      i = (int)r;
      j = (int)(pg - (r * M - r * (r - 1) / 2)) + i;
    }
#pragma unroll
    for (int f = 0; f < FPL; ++f) {
      const int k = lane + 32 * f;
      mine[k] = mish_fast((z[(size_t)i * H + k] + z[(size_t)j * H + k]) * 0.5f);
    }
    __syncwarp();
    float s[JPL];
#pragma unroll
    for (int m = 0; m < JPL; ++m) s[m] = b1[m];
    for (int k = 0; k < H; ++k) {
      const float x = mine[k];
#pragma unroll
      for (int m = 0; m < JPL; ++m)
        if (lane + 32 * m < HH) s[m] = fmaf(w1t[k * HH + lane + 32 * m], x, s[m]);
    }
    float contrib = 0.f;
#pragma unroll
    for (int m = 0; m < JPL; ++m)
      if (lane + 32 * m < HH) contrib = fmaf(w2[m], mish_fast(mish_fast(s[m])), contrib);
    const float logit = warp_sum(contrib) + b2;
    if (lane == 0) prob[pidx] = __fdividef(1.0f, 1.0f + __expf(-logit));
    __syncwarp();
  }
}

template <int H>
cudaError_t head_fwd_h(int B, float beta, const float* z, const long long* y, const ib200_head_params& hp, const ib200_head_masks& hm,
                       float* losses, float* y_hat, cudaStream_t st) {
  const size_t smem = sizeof(HeadSmem<H>);
  cudaError_t e = cudaFuncSetAttribute(loss_head_fwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  loss_head_fwd_kernel<H><<<1, kHeadWarps * 32, smem, st>>>(B, beta, z, y, hp, hm, losses, y_hat);
  return cudaGetLastError();
}
template <int H>
cudaError_t head_bwd_h(int B, float beta, const float* z, const long long* y, const ib200_head_params& hp, const ib200_head_masks& hm,
                       const float* d_loss, const float* d_y_hat, float* dz, const ib200_head_grads& hg, cudaStream_t st) {
  const size_t smem = sizeof(HeadBwdSmem<H>);
  cudaError_t e = cudaFuncSetAttribute(loss_head_bwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  loss_head_bwd_kernel<H><<<1, kHeadWarps * 32, smem, st>>>(B, beta, z, y, hp, hm, d_loss, d_y_hat, dz, hg);
  return cudaGetLastError();
}
template <int H>
cudaError_t pair_wide_h(int M, const float* z, const int* idx_a, const int* idx_b, long long P, long long p0,
                        const ib200_head_params& hp, float* prob, cudaStream_t st) {
  const size_t smem = ((size_t)H * (H / 2) + 8 * H) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(pair_score_wide_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)std::min<long long>((P + 7) / 8, 148LL * 8);
  pair_score_wide_kernel<H><<<grid, 256, smem, st>>>(M, z, idx_a, idx_b, P, p0, hp, prob);
  return cudaGetLastError();
}

}  // namespace

bool head_supports(int H) { return H % 32 == 0 && H >= 32 && H <= 256; }

#define IB200_HEAD_DISPATCH(FN, ...)                 \
  switch (H) {                                       \
    case 32: return FN<32>(__VA_ARGS__);             \
    case 64: return FN<64>(__VA_ARGS__);             \
    case 96: return FN<96>(__VA_ARGS__);             \
    case 128: return FN<128>(__VA_ARGS__);           \
    case 160: return FN<160>(__VA_ARGS__);           \
    case 192: return FN<192>(__VA_ARGS__);           \
    case 224: return FN<224>(__VA_ARGS__);           \
    case 256: return FN<256>(__VA_ARGS__);           \
    default: return cudaErrorInvalidValue;           \
  }

cudaError_t launch_loss_head_fwd(int B, int H, float beta, const float* z, const long long* y, const ib200_head_params& hp,
                                 const ib200_head_masks& hm, float* losses, float* y_hat, cudaStream_t st) {
  IB200_HEAD_DISPATCH(head_fwd_h, B, beta, z, y, hp, hm, losses, y_hat, st)
}

cudaError_t launch_loss_head_bwd(int B, int H, float beta, const float* z, const long long* y, const ib200_head_params& hp,
                                 const ib200_head_masks& hm, const float* d_loss, const float* d_y_hat, float* dz,
                                 const ib200_head_grads& hg, cudaStream_t st) {
  IB200_HEAD_DISPATCH(head_bwd_h, B, beta, z, y, hp, hm, d_loss, d_y_hat, dz, hg, st)
}

cudaError_t launch_pair_score(int M, int H, const float* z, const int* idx_a, const int* idx_b, long long P, long long p0,
                              const ib200_head_params& hp, float* prob, cudaStream_t st) {
  if (P <= 0) return cudaSuccess;
  if (H > 64) {
    switch (H) {
      case 96: return pair_wide_h<96>(M, z, idx_a, idx_b, P, p0, hp, prob, st);
      case 128: return pair_wide_h<128>(M, z, idx_a, idx_b, P, p0, hp, prob, st);
      case 160: return pair_wide_h<160>(M, z, idx_a, idx_b, P, p0, hp, prob, st);
      case 192: return pair_wide_h<192>(M, z, idx_a, idx_b, P, p0, hp, prob, st);
      case 224: return pair_wide_h<224>(M, z, idx_a, idx_b, P, p0, hp, prob, st);
      case 256: return pair_wide_h<256>(M, z, idx_a, idx_b, P, p0, hp, prob, st);
      default: return cudaErrorInvalidValue;
    }
  }
  const unsigned grid = (unsigned)std::min<long long>((P + 127) / 128, 148LL * 16);
  if (H == 64) pair_score_kernel<64><<<grid, 128, 0, st>>>(M, z, idx_a, idx_b, P, p0, hp, prob);
  else if (H == 32) pair_score_kernel<32><<<grid, 128, 0, st>>>(M, z, idx_a, idx_b, P, p0, hp, prob);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace ib200
