// Production-mode dropout masks: all Bernoulli(keep)/keep masks of one training step (5 x vocabulary-row mask, 5 x
// weight_hh_l0 DropConnect mask, 4 head masks) in ONE launch instead of twelve torch kernels (bernoulli_ + divide per mask).
// The reference draws them with torch's RNG (utils/embedding_do.py:26-28, utils/weightdrop.py:92-102, classifier/head/mlp.py:38-58);
// masks stay INPUTS of the compute kernels, so parity tests keep feeding explicit masks to both sides.
//
// Generator: Philox4x32-10, counter = offset + (global draw index / 4), key = seed ^ "IB200_MS"; draw j of a counter is output
// word j; u = (word >> 8) * 2^-24; value = (u < keep) ? 1/keep : 0.  Every mask starts on a fresh counter.  Bit-exact restatement:
// oracle/restatement.py::draw_masks (Random123 known-answer vectors in tests/test_next_rows_cpu.py).
#include "kernels.h"

namespace ib200 {
namespace {

struct MaskTable {
  float* out[kMaskMaxSpecs];
  long long numel[kMaskMaxSpecs];
  long long draws[kMaskMaxSpecs];
  long long quad0[kMaskMaxSpecs + 1];  // first counter (relative to `offset`) of spec s; [n] = total
  float keep[kMaskMaxSpecs], inv[kMaskMaxSpecs];
  int row_len[kMaskMaxSpecs];
  int n;
};

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t lo0 = 0xD2511F53u * c[0], hi0 = __umulhi(0xD2511F53u, c[0]);
    const uint32_t lo1 = 0xCD9E8D57u * c[2], hi1 = __umulhi(0xCD9E8D57u, c[2]);
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

__global__ void __launch_bounds__(256) draw_masks_kernel(const __grid_constant__ MaskTable tb, unsigned long long key,
                                                         unsigned long long offset) {
  const long long total = tb.quad0[tb.n];
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    int s = 0;
    while (s + 1 < tb.n && q >= tb.quad0[s + 1]) ++s;
    const unsigned long long ctr = offset + (unsigned long long)q;
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
    philox4x32_10(c, (uint32_t)key, (uint32_t)(key >> 32));
    const long long d0 = (q - tb.quad0[s]) * 4;
    const float keep = tb.keep[s], inv = tb.inv[s];
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = ((float)(c[j] >> 8) * (1.0f / 16777216.0f) < keep) ? inv : 0.0f;
    float* __restrict__ out = tb.out[s];
    const int rl = tb.row_len[s];
    if (rl <= 1) {
      if (d0 + 4 <= tb.draws[s] && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
        *reinterpret_cast<float4*>(out + d0) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (d0 + j < tb.draws[s]) out[d0 + j] = v[j];
      }
    } else {  // one draw per row of rl elements (variational row mask, utils/weightdrop.py:92-95)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (d0 + j < tb.draws[s])
          for (int e = 0; e < rl; ++e) {
            const long long i = (d0 + j) * rl + e;
            if (i < tb.numel[s]) out[i] = v[j];
          }
    }
  }
}

}  // namespace

cudaError_t launch_draw_masks(int n, const ib200_mask_spec* specs, unsigned long long seed, unsigned long long offset,
                              unsigned long long* counters_used, cudaStream_t st, int* launches) {
  const unsigned long long key = seed ^ 0x49423230305F4D53ull;
  unsigned long long used = 0;
  *launches = 0;
  for (int s0 = 0; s0 < n; s0 += kMaskMaxSpecs) {
    MaskTable tb{};
    const int cnt = n - s0 < kMaskMaxSpecs ? n - s0 : kMaskMaxSpecs;
    long long quads = 0;
    for (int i = 0; i < cnt; ++i) {
      const ib200_mask_spec& sp = specs[s0 + i];
      const int rl = sp.row_len > 1 ? sp.row_len : 1;
      tb.out[i] = sp.out; tb.numel[i] = sp.numel; tb.row_len[i] = rl;
      tb.draws[i] = (sp.numel + rl - 1) / rl;
      tb.keep[i] = sp.keep_prob; tb.inv[i] = 1.0f / sp.keep_prob;
      tb.quad0[i] = quads;
      quads += (tb.draws[i] + 3) / 4;
    }
    tb.quad0[cnt] = quads;
    tb.n = cnt;
    if (quads > 0) {
      const long long blocks = (quads + 255) / 256;
      draw_masks_kernel<<<(unsigned)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, st>>>(tb, key, offset + used);
      ++*launches;
      const cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return e;
    }
    used += (unsigned long long)quads;
  }
  if (counters_used) *counters_used = used;
  return cudaSuccess;
}

}  // namespace ib200
