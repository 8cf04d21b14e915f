// Per-step classification metrics of TripletE2ENet.step (e2e/e2e_triplet.py:171-184): binary AUROC, average precision, Matthews
// correlation, precision and recall of the batch, the values torchmetrics 0.11.1's `metric(y_hat, y)` forward returns and the
// reference logs.  The reference issues five torchmetrics calls (dozens of small kernels and several host syncs: the [0,1] range
// check, argsort, unique thresholds, cumsum, ...); this is ONE launch of one CTA and no sync.
//
// Algorithm (torchmetrics functional/classification: _binary_*_format, _binary_clf_curve, _binary_roc_compute,
// _binary_precision_recall_curve_compute, _matthews_corrcoef_reduce, _precision_recall_reduce):
//   scores = y_hat if every value lies in [0,1] else sigmoid(y_hat);  sort descending, one curve point per DISTINCT score;
//   AUROC = trapezoid area under (fpr, tpr) starting at (0,0), a missing class gives an all-zero rate (=> 0);
//   AP    = sum_k (recall_k - recall_{k-1}) * precision_k (NaN without positives);
//   hard  = score > 0.5 -> confusion counts -> MCC (0 when its denominator vanishes), precision, recall (0 on 0/0).
#include "kernels.h"

namespace ib200 {
namespace {

constexpr int kMT = 1024;  // threads == maximum batch

__device__ __forceinline__ unsigned desc_key(float s) {  // unsigned ascending order of the result == DESCENDING order of s
  unsigned u = __float_as_uint(s);
  u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;  // ascending-float -> ascending-unsigned
  return ~u;
}

__device__ __forceinline__ int block_sum(int v, int* red) {
  v = __reduce_add_sync(0xffffffffu, v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int w = 0; w < kMT / 32; ++w) t += red[w];
  return t;
}

__device__ __forceinline__ float block_sum_f(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < kMT / 32; ++w) t += red[w];
  return t;
}

__global__ void __launch_bounds__(kMT) batch_metrics_kernel(int B, const float* __restrict__ y_hat, const long long* __restrict__ y,
                                                            float threshold, float* __restrict__ out, int* __restrict__ conf) {
  __shared__ unsigned long long keys[kMT];
  __shared__ int scan[2][kMT];
  __shared__ int tp_at[kMT];
  __shared__ int red_i[kMT / 32];
  __shared__ float red_f[kMT / 32];
  const int i = threadIdx.x;
  const bool have = i < B;
  const float x = have ? y_hat[i] : 0.5f;
  const int label = have ? (y[i] != 0 ? 1 : 0) : 0;

  // _binary_*_format: squash with a sigmoid only when some prediction lies outside [0,1]
  const int outside = block_sum(have && !(x >= 0.f && x <= 1.f) ? 1 : 0, red_i);
  const float s = outside ? 1.0f / (1.0f + expf(-x)) : x;

  // confusion counts at `threshold`
  const bool hard = s > threshold;
  const int tp = block_sum(have && hard && label ? 1 : 0, red_i);
  const int fp = block_sum(have && hard && !label ? 1 : 0, red_i);
  const int P = block_sum(label, red_i);
  const int N = B - P, fn = P - tp, tn = N - fp;

  // descending bitonic sort of (score, label); the padding sorts last
  keys[i] = have ? ((unsigned long long)desc_key(s) << 32) | (unsigned)label : ~0ull;
  __syncthreads();
  for (int k = 2; k <= kMT; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int o = i ^ j;
      if (o > i) {
        const unsigned long long a = keys[i], b = keys[o];
        if (((i & k) == 0) == (a > b)) { keys[i] = b; keys[o] = a; }
      }
      __syncthreads();
    }
  const unsigned long long mine = keys[i];
  const int l_sorted = have ? (int)(mine & 1ull) : 0;
  const bool last_of_score = have && (i == B - 1 || (unsigned)(mine >> 32) != (unsigned)(keys[i + 1] >> 32));

  // inclusive scans: tp_i = positives among the first i+1 scores; prev_i = index of the last curve point before i (-1: none)
  scan[0][i] = l_sorted;
  __syncthreads();
  int cur = 0;
  for (int d = 1; d < kMT; d <<= 1) {
    scan[cur ^ 1][i] = scan[cur][i] + (i >= d ? scan[cur][i - d] : 0);
    cur ^= 1;
    __syncthreads();
  }
  const int tp_i = scan[cur][i];
  tp_at[i] = tp_i;
  __syncthreads();
  scan[0][i] = last_of_score ? i : -1;
  __syncthreads();
  cur = 0;
  for (int d = 1; d < kMT; d <<= 1) {
    scan[cur ^ 1][i] = max(scan[cur][i], i >= d ? scan[cur][i - d] : -1);
    cur ^= 1;
    __syncthreads();
  }
  const int prev = i > 0 ? scan[cur][i - 1] : -1;

  float a_auc = 0.f, a_ap = 0.f;
  if (last_of_score) {
    const int fp_i = i + 1 - tp_i;
    const int tp_p = prev >= 0 ? tp_at[prev] : 0, fp_p = prev >= 0 ? prev + 1 - tp_p : 0;
    const float tpr = P > 0 ? (float)tp_i / (float)P : 0.f, tpr_p = P > 0 ? (float)tp_p / (float)P : 0.f;
    const float fpr = N > 0 ? (float)fp_i / (float)N : 0.f, fpr_p = N > 0 ? (float)fp_p / (float)N : 0.f;
    a_auc = (fpr - fpr_p) * (tpr + tpr_p) * 0.5f;
    const float rec = (float)tp_i / (float)P, rec_p = (float)tp_p / (float)P;  // 0/0 = NaN without positives, as torchmetrics
    a_ap = (rec - rec_p) * ((float)tp_i / (float)(i + 1));
  }
  const float auroc = block_sum_f(a_auc, red_f);
  const float ap = block_sum_f(a_ap, red_f);
  if (i == 0) {
    // _matthews_corrcoef_reduce on [[tn, fp], [fn, tp]] in fp32 (exact: every term is an integer below 2^24 for B <= 1024)
    const float t0 = (float)(tn + fp), t1 = (float)(fn + tp), p0 = (float)(tn + fn), p1 = (float)(fp + tp);
    const float c = (float)(tn + tp), sN = (float)B;
    const float cov_ytyp = c * sN - (t0 * p0 + t1 * p1), cov_ypyp = sN * sN - (p0 * p0 + p1 * p1), cov_ytyt = sN * sN - (t0 * t0 + t1 * t1);
    const float denom = cov_ypyp * cov_ytyt;
    out[0] = auroc;
    out[1] = P > 0 ? ap : __int_as_float(0x7fc00000);
    out[2] = denom == 0.f ? 0.f : cov_ytyp / sqrtf(denom);
    out[3] = tp + fp > 0 ? (float)tp / (float)(tp + fp) : 0.f;
    out[4] = tp + fn > 0 ? (float)tp / (float)(tp + fn) : 0.f;
    if (conf != nullptr) { conf[0] = tp; conf[1] = fp; conf[2] = tn; conf[3] = fn; }
  }
}

}  // namespace

cudaError_t launch_batch_metrics(int B, const float* y_hat, const long long* y, float threshold, float* out, int* conf, cudaStream_t st) {
  if (B < 1 || B > kMT) return cudaErrorInvalidValue;
  batch_metrics_kernel<<<1, kMT, 0, st>>>(B, y_hat, y, threshold, out, conf);
  return cudaGetLastError();
}

}  // namespace ib200
