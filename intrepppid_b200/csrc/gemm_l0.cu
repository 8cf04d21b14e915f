// Layer-0 parameter gradients of the H = 64 path in ONE pass over the layer-0 dgates (tcgen05 + TMA, bf16 hi/lo planes).
//
// The layer-0 input is an embedding lookup, x[row] = scale_g[tok] * emb[tok] (utils/embedding_do.py:26-43 folded into the gather), so
// both gradients that involve x factor through the token-indexed sums
//     S_gd[k][v] = sum over rows (n,t) of group g with tok = v, t < T_eff :  dA_d[row][k]            (k: gate column, GI order)
//     dW_ih_l0,d = sum_g S_gd * (scale_g (.) emb)                                   [4H, H]
//     dEmb       = sum_g scale_g (.) sum_d S_gd^T * W_ih_l0,d     (row 0 = padding_idx -> 0)   [V, H]
// S_gd = dA_d^T * onehot(tok) is a tensor-core GEMM whose B operand never touches HBM: the one-hot tile is generated in shared
// memory from the token ids (one bf16 1.0 per row, exact).  The same pass also forms dW_hh_l0,d = dA_d^T * Hprev with
// Hprev = Y_0 shifted by one step (TMA).  This replaces, for layer 0: the gathered-embedding dW GEMM, the dX_0 = dA W_ih GEMM
// over all token rows and the atomic scatter of dX_0 into the embedding gradient.
//
// Grid (ctas_per_group, G, 2 directions x 2 halves of the 256 gate columns): each CTA owns 128 gate columns (UMMA M = 128) of
// one direction; per 64-row item it loads its half of the dA rows (hi, lo) and the Hprev tile with TMA, four warps write the
// one-hot tile, one thread issues per k16:  D_hh[128,64] += A_hi Hp_hi + A_hi Hp_lo + A_lo Hp_hi ;  D_S[128,256] += A_hi OH + A_lo OH.
#include <algorithm>

#include "kernels.h"
#include "tc05.cuh"
#include "tma_host.h"

namespace ib200 {
namespace {

using namespace tc;

// Operand pipeline: the TMA stages (dgates + Hprev, 48 KB in fp32 mode / 24 KB in bf16 mode) and the one-hot tiles (32 KB, written from
// the token ids by four warps: no HBM latency to hide) are SEPARATE rings.  With one ring of two 80 KB stages only one item's loads
// were in flight and every item paid part of the TMA latency: 0.32 ms at 54 % of HBM with the tensor pipe 65 % busy -- neither bound.
// fp32 mode: 3 TMA stages + 2 one-hot tiles (0.319 -> 0.290 ms for the three launches of the family; 4 + 1 measured 0.310: a single
// one-hot tile serialises the writers with the MMAs); bf16 mode: 5 + 2 (0.204 ms, 85 % of HBM).
template <bool SPLIT>
constexpr int ay_stages() { return SPLIT ? 3 : 5; }
template <bool SPLIT>
constexpr int oh_stages() { return 2; }
constexpr int kMaxStagesAY = 5, kMaxStagesOH = 2;
constexpr int kBlk = 64 * 128;        // one [64 k-rows x 64 columns] bf16 block (128-byte swizzled rows)
constexpr int kNS = 256;              // one-hot columns (vocabulary slots; V <= 256)
constexpr int kNOut = 64 + kNS;       // accumulator columns per gate row: [dW_hh (64) | S (256)]
constexpr uint32_t kTmemCols = 512;   // 320 used

struct Bars {
  uint64_t full[kMaxStagesAY], empty[kMaxStagesAY], ohfull[kMaxStagesOH], ohempty[kMaxStagesOH], done;
  uint32_t tmem_base;
};
struct Maps {
  CUtensorMap a[2][2];  // [direction][plane] dgates planes: 3D {4H, Tmax, N}, box {64, 64, 1}
  CUtensorMap y[2];     // [plane] Y_0 planes: 3D {2H, Tmax, N}, box {64, 64, 1}
};

template <bool SPLIT>
__global__ void __launch_bounds__(192, 1) l0_grad_gemm_kernel(const __grid_constant__ Maps maps, const L0GradArgs p) {
  constexpr int NPART = SPLIT ? 2 : 1;
  constexpr int kABytes = 2 * kBlk, kYBytes = kBlk, kOHBytes = (kNS / 64) * kBlk;
  constexpr int kStageBytes = NPART * (kABytes + kYBytes), kStages = ay_stages<SPLIT>(), kStagesOH = oh_stages<SPLIT>();
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* ohbuf = smem + (size_t)kStages * kStageBytes;  // [kStagesOH][kOHBytes]
  Bars* bars = reinterpret_cast<Bars*>(ohbuf + (size_t)kStagesOH * kOHBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = blockIdx.y, cta = blockIdx.x, d = p.dir0 + (int)(blockIdx.z >> 1), mh = blockIdx.z & 1;
  const int T = p.lens[p.G + g];
  const int tiles_per_seq = (T + 63) / 64;
  const int items = p.B * tiles_per_seq;
  const int my_items = cta < items ? (items - cta + p.ctas_per_group - 1) / p.ctas_per_group : 0;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int s = 0; s < kStagesOH; ++s) {
      mbar_init(&bars->ohfull[s], 128);
      mbar_init(&bars->ohempty[s], 1);
    }
    mbar_init(&bars->done, 1);
    mbar_init_fence();
    for (int pl = 0; pl < NPART; ++pl) {
      tma_prefetch_desc(&maps.a[d][pl]);
      tma_prefetch_desc(&maps.y[pl]);
    }
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, kTmemCols);
  // the one-hot regions start as zeros; afterwards only the 64 ones of the previous use of a stage are cleared
  for (int i = tid; i < kStagesOH * kOHBytes / 16; i += 192) reinterpret_cast<uint4*>(ohbuf)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = cta; item < items; item += p.ctas_per_group, ++it) {
        const int stage = it % kStages;
        const int n = g * p.B + item / tiles_per_seq, t0 = (item % tiles_per_seq) * 64;
        mbar_wait(&bars->empty[stage], ((it / kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->full[stage], NPART * (kABytes + kYBytes));
        unsigned char* a_dst = smem + (size_t)stage * kStageBytes;
        unsigned char* y_dst = a_dst + NPART * kABytes;
#pragma unroll
        for (int pl = 0; pl < NPART; ++pl) {
          tma_load_3d(a_dst + pl * kABytes, &maps.a[d][pl], &bars->full[stage], 128 * mh, t0, n);
          tma_load_3d(a_dst + pl * kABytes + kBlk, &maps.a[d][pl], &bars->full[stage], 128 * mh + 64, t0, n);
          // h of the previous scan position: forward chain t-1, reverse chain t+1 (rows outside [0, T) are zero: OOB / zeroed tail)
          tma_load_3d(y_dst + pl * kYBytes, &maps.y[pl], &bars->full[stage], d * 64, t0 + (d == 0 ? -1 : 1), n);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp, uniform control flow; one elected lane issues: tc05.cuh) =====================
    {
      constexpr uint32_t idesc_hh = idesc_bf16(128, 64, true, true), idesc_s = idesc_bf16(128, kNS, true, true);
      const uint32_t d_hh = tmem_base, d_s = tmem_base + 64;
      for (int it = 0; it < my_items; ++it) {
        const int stage = it % kStages, ohs = it % kStagesOH;
        mbar_wait(&bars->full[stage], (it / kStages) & 1);
        mbar_wait(&bars->ohfull[ohs], (it / kStagesOH) & 1);
        fence_after_sync();
        const uint32_t a_hi = smem_u32(smem + (size_t)stage * kStageBytes), a_lo = a_hi + kABytes;
        const uint32_t y_hi = a_hi + NPART * kABytes, y_lo = y_hi + kYBytes;
        const uint32_t oh = smem_u32(ohbuf + (size_t)ohs * kOHBytes);
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) {
          const uint32_t ko = k16 * 16 * 128;  // 16 k-rows of 128 bytes inside every block
          const uint64_t ah = smem_desc_sw128(a_hi + ko, 1024, kBlk), yh = smem_desc_sw128(y_hi + ko, 1024, kBlk);
          const uint64_t ohd = smem_desc_sw128(oh + ko, 1024, kBlk);
          const bool acc = (it | k16) != 0;
          mma_bf16_ss_elect(d_hh, ah, yh, idesc_hh, acc);
          mma_bf16_ss_elect(d_s, ah, ohd, idesc_s, acc);
          if constexpr (SPLIT) {
            const uint64_t al = smem_desc_sw128(a_lo + ko, 1024, kBlk), yl = smem_desc_sw128(y_lo + ko, 1024, kBlk);
            mma_bf16_ss_elect(d_hh, ah, yl, idesc_hh, true);
            mma_bf16_ss_elect(d_hh, al, yh, idesc_hh, true);
            mma_bf16_ss_elect(d_s, al, ohd, idesc_s, true);
          }
        }
        mma_commit_elect(&bars->empty[stage]);
        mma_commit_elect(&bars->ohempty[ohs]);
      }
      mma_commit_elect(&bars->done);
    }
  } else {
    // ===================== one-hot writers (4 warps): thread r < 64 owns k-row r of the tile =====================
    const int r = tid - 64;
    int prev[kStagesOH];
#pragma unroll
    for (int s = 0; s < kStagesOH; ++s) prev[s] = -1;
    uint32_t it = 0;
    for (int item = cta; item < items; item += p.ctas_per_group, ++it) {
      const int stage = it % kStagesOH;
      const int n = g * p.B + item / tiles_per_seq, t0 = (item % tiles_per_seq) * 64;
      int v = -1;
      if (r < 64 && t0 + r < T) v = p.tok[(size_t)n * p.Tmax + t0 + r];
      mbar_wait(&bars->ohempty[stage], ((it / kStagesOH) & 1) ^ 1);
      unsigned char* oh = ohbuf + (size_t)stage * kOHBytes;
      if (r < 64) {
        // element (k-row r, column v) of the MN-major tile: block v/64, 128-byte row r, 16-byte chunk ((v%64)/8) ^ (r%8)
        if (prev[stage] >= 0) *reinterpret_cast<unsigned short*>(oh + prev[stage]) = 0;
        int off = -1;
        if (v >= 0 && v < kNS) {
          off = (v >> 6) * kBlk + (int)sw128_offset((uint32_t)r, (uint32_t)((v & 63) >> 3)) + (v & 7) * 2;
          *reinterpret_cast<unsigned short*>(oh + off) = 0x3F80;  // bf16 1.0
        }
        prev[stage] = off;
      }
      fence_async_smem();
      mbar_arrive(&bars->ohfull[stage]);
    }
  }
  __syncthreads();

  // ===================== epilogue: TMEM -> partial [128 gate columns of this half][320] =====================
  float* out = p.partial + ((((size_t)(blockIdx.z >> 1) * p.G + g) * p.ctas_per_group + cta) * 256 + (size_t)mh * 128) * kNOut;
  if (warp >= 2) {
    const int q = warp & 3;
    if (my_items > 0) {
      mbar_wait(&bars->done, 0);
      fence_after_sync();
    }
    float* orow = out + (size_t)(q * 32 + lane) * kNOut;
#pragma unroll 1
    for (int c = 0; c < kNOut / 32; ++c) {
      uint32_t rr[32];
      if (my_items > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, rr);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) rr[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(orow + c * 32 + 4 * j) = make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]),
                                                                     __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// R[dslot][g][k][c] = sum over the CTAs of group g of partial[dslot][g][cta][k][c]      (fixed order: deterministic)
__global__ void __launch_bounds__(kNOut) l0_reduce_kernel(const float* __restrict__ partial, float* __restrict__ R, int ctas_per_group) {
  const size_t blk = blockIdx.x;  // (dslot * G + g) * 256 + k
  const size_t dg = blk / 256, k = blk % 256;
  const int c = threadIdx.x;
  float s = 0.f;
#pragma unroll 8
  for (int cta = 0; cta < ctas_per_group; ++cta) s += partial[((dg * ctas_per_group + cta) * 256 + k) * kNOut + c];
  R[blk * kNOut + c] = s;
}

// dW_hh (per-group DropConnect mask on the forward direction), dW_ih, both bias gradients; one block per (direction slot, gate row k).
// 256 threads = 64 output columns x 4 slices of the vocabulary; the slices are added in a fixed order.
__device__ __forceinline__ void l0_finish_w(const L0GradArgs& p, const float* __restrict__ R, int ds, int k) {
  constexpr int H = 64;
  __shared__ float sr[kNS];        // S_gd[k][v] * scale_g[v]
  __shared__ float part[4][H];
  const int d = p.dir0 + ds, h = threadIdx.x & 63, q = threadIdx.x >> 6;
  const int row = gi_to_torch_row(k, H);
  float whh = 0.f, wih = 0.f;
  for (int g = 0; g < p.G; ++g) {
    const float* __restrict__ Rk = R + (((size_t)ds * p.G + g) * 256 + k) * kNOut;
    __syncthreads();
    for (int v = threadIdx.x; v < p.V; v += 256)
      sr[v] = Rk[64 + v] * (p.emb_row_scale != nullptr ? p.emb_row_scale[(size_t)g * p.V + v] : 1.0f);
    __syncthreads();
    if (q == 0) {
      float m = 1.0f;
      if (d == 0 && p.whh_mask != nullptr) m = p.whh_mask[((size_t)g * 4 * H + row) * H + h];
      whh = fmaf(m, Rk[h], whh);
    }
    float acc = 0.f;
#pragma unroll 8
    for (int v = q; v < p.V; v += 4) acc = fmaf(sr[v], p.emb[(size_t)v * H + h], acc);
    wih += acc;
  }
  part[q][h] = wih;
  __syncthreads();
  if (q == 0) {
    p.d_whh[d][(size_t)row * H + h] = whh;
    p.d_wih[d][(size_t)row * H + h] = (part[0][h] + part[1][h]) + (part[2][h] + part[3][h]);
  }
  if (threadIdx.x == 64) {  // bias gradient = column sum of the dgates: per-CTA partials left by the BPTT kernel
    float bsum = 0.f;
    const float* cs = p.bias_partial + (size_t)ds * p.bias_count * 4 * H;
    for (int i = 0; i < p.bias_count; ++i) bsum += cs[(size_t)i * 4 * H + k];
    p.d_bih[d][row] = bsum;
    p.d_bhh[d][row] = bsum;
  }
}

// dEmb[v] = sum_g scale_g[v] * sum_d sum_k S_gd[k][v] W_ih,d[row(k)]; one block per vocabulary row, 64 columns x 4 slices of k
__device__ __forceinline__ void l0_finish_emb(const L0GradArgs& p, const float* __restrict__ R, int v) {
  constexpr int H = 64;
  __shared__ float col[256];
  __shared__ float part[4][H];
  const int h = threadIdx.x & 63, q = threadIdx.x >> 6;
  float tot = 0.f;
  if (v != 0) {  // padding_idx = 0 receives no gradient (nn.Embedding(..., padding_idx=0), e2e_triplet.py:345)
    for (int g = 0; g < p.G; ++g) {
      const float sc = p.emb_row_scale != nullptr ? p.emb_row_scale[(size_t)g * p.V + v] : 1.0f;
      if (sc == 0.0f) continue;  // (uniform over the block)
      float acc = 0.f;
      for (int ds = 0; ds < p.ndir; ++ds) {
        __syncthreads();
        col[threadIdx.x] = R[(((size_t)ds * p.G + g) * 256 + threadIdx.x) * kNOut + 64 + v];
        __syncthreads();
        const float* __restrict__ W = p.w_ih[p.dir0 + ds];
#pragma unroll 16
        for (int k = 64 * q; k < 64 * q + 64; ++k) acc = fmaf(col[k], W[(size_t)gi_to_torch_row(k, H) * H + h], acc);
      }
      tot = fmaf(sc, acc, tot);
    }
  }
  part[q][h] = tot;
  __syncthreads();
  if (q == 0) p.d_emb[(size_t)v * H + h] = (part[0][h] + part[1][h]) + (part[2][h] + part[3][h]);
}

// both finishing passes in ONE launch (they read the same R and write disjoint outputs): blocks [0, 256*ndir) finish the weights,
// the remaining V blocks the embedding rows
__global__ void __launch_bounds__(256) l0_finish_kernel(const L0GradArgs p, const float* __restrict__ R) {
  const int nw = 256 * p.ndir;
  if ((int)blockIdx.x < nw) l0_finish_w(p, R, (int)blockIdx.x / 256, (int)blockIdx.x % 256);
  else l0_finish_emb(p, R, (int)blockIdx.x - nw);
}

}  // namespace

size_t l0_grad_scratch_floats(int G, int ndir) { return (size_t)ndir * G * 256 * kNOut; }
int l0_grad_ctas_per_group(int G, int ndir) { return std::max(1, 148 / (G * ndir * 2)); }
size_t l0_grad_partial_floats(int G, int ndir) { return (size_t)ndir * G * l0_grad_ctas_per_group(G, ndir) * 256 * kNOut; }

// returns cudaErrorInvalidConfiguration when the shape is not covered (H != 64, V > 256): the caller keeps the general path
cudaError_t launch_l0_grads(const L0GradArgs& a0, int precision, cudaStream_t st) {
  if (a0.H != 64 || a0.V > kNS || a0.ndir < 1 || a0.ndir > 2) return cudaErrorInvalidConfiguration;
  L0GradArgs a = a0;
  a.ctas_per_group = l0_grad_ctas_per_group(a.G, a.ndir);
  const int npart = precision == 0 ? 2 : 1;
  const size_t smem = 1024 + (size_t)(precision == 0 ? ay_stages<true>() : ay_stages<false>()) * npart * 3 * kBlk +
                      (size_t)(precision == 0 ? oh_stages<true>() : oh_stages<false>()) * (kNS / 64) * kBlk + sizeof(Bars) + 64;
  Maps maps;
  const uint32_t box[3] = {64, 64, 1};
  for (int pl = 0; pl < npart; ++pl) {
    for (int ds = 0; ds < a.ndir; ++ds) {
      const int d = a.dir0 + ds;
      const uint64_t da[3] = {256, (uint64_t)a.Tmax, (uint64_t)a.G * a.B}, sa[2] = {1024, (uint64_t)a.Tmax * 1024};
      if (!make_tmap_bf16_sw128(&maps.a[d][pl], reinterpret_cast<const unsigned char*>(a.dA[d]) + (size_t)pl * 512, 3, da, sa, box))
        return cudaErrorInvalidConfiguration;
    }
    const uint64_t dy[3] = {128, (uint64_t)a.Tmax, (uint64_t)a.G * a.B}, sy[2] = {512, (uint64_t)a.Tmax * 512};
    if (!make_tmap_bf16_sw128(&maps.y[pl], reinterpret_cast<const unsigned char*>(a.Y0) + (size_t)pl * 256, 3, dy, sy, box))
      return cudaErrorInvalidConfiguration;
  }
  dim3 grid(a.ctas_per_group, a.G, 2 * a.ndir);
  cudaError_t e;
  if (precision == 0) {
    e = cudaFuncSetAttribute(l0_grad_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    l0_grad_gemm_kernel<true><<<grid, 192, smem, st>>>(maps, a);
  } else {
    e = cudaFuncSetAttribute(l0_grad_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    l0_grad_gemm_kernel<false><<<grid, 192, smem, st>>>(maps, a);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  l0_reduce_kernel<<<(unsigned)(a.ndir * a.G * 256), kNOut, 0, st>>>(a.partial, a.R, a.ctas_per_group);
  l0_finish_kernel<<<256 * a.ndir + a.V, 256, 0, st>>>(a, a.R);
  return cudaGetLastError();
}

}  // namespace ib200
