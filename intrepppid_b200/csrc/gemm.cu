// Tensor-core GEMMs over token rows (everything in the LSTM that is NOT on the dependent chain):
//   NT : xproj_l = Y_{l-1} W_ih^T + b           (input projection of layers >= 1, reference: nn.LSTM's x W_ih^T)
//        dY_{l-1} = sum_d dA_{l,d} W_ih,d        (input gradient), dX_0 likewise
//   TN : dW_ih = dA^T X , dW_hh = dA^T H_prev , db = colsum(dA)   (split over CTAs, deterministic two-pass reduction)
// Rows are (n,t) pairs of the [N, Tmax] activation tensors; rows with t >= T_eff[group(n)] do not exist in the reference
// (quirk Q2) and are skipped / zero-filled here.
// Operands are fp32 in HBM; fp32 mode splits them into bf16 hi + lo on the fly and issues 3 mma.m16n8k16 per product.
#include <algorithm>

#include "kernels.h"

namespace ib200 {
namespace {

constexpr int kBM = 128;   // rows per CTA tile (NT)
constexpr int kBK = 32;    // k slice (NT) / row slice (TN)
constexpr int kPad = 8;    // NT smem row padding (floats): stride 40 words => conflict-free float2 fragment loads

// ------------------------------------------------------------------------------------------------------------------------
// NT
// ------------------------------------------------------------------------------------------------------------------------
template <int NC, int WM, int WN, bool SPLIT>
__global__ void __launch_bounds__(256, 1) gemm_nt_kernel(const GemmNTArgs p) {
  constexpr int MTW = kBM / WM / 16, NTW = NC / WN / 8, S = kBK + kPad;
  static_assert(WM * WN == 8, "8 warps");
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                   // [2][kBM][S]
  float* Ws = smem + 2 * kBM * S;     // [2][NC][S]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tig = lane & 3;
  const int wm = warp / WN, wn = warp % WN;
  const long long row0 = (long long)blockIdx.x * kBM;
  const long long nrows = (long long)p.G * p.B * p.Tmax;
  // tile-level skip: the whole tile lies in the dead tail of one sequence
  {
    const long long rl = min(row0 + kBM, nrows) - 1;
    const int na = (int)(row0 / p.Tmax), nb = (int)(rl / p.Tmax);
    if (na == nb && (int)(row0 % p.Tmax) >= p.lens[p.G + na / p.B]) return;
  }
  const int kslices = p.K / kBK, total = p.nsrc * kslices;

  auto load_slice = [&](int sl, int buf) {
    const int src = sl / kslices, k0 = (sl % kslices) * kBK;
    const float* __restrict__ A = p.A[src];
    const float* __restrict__ W = p.W[src];
    float* as = As + buf * kBM * S;
    float* ws = Ws + buf * NC * S;
    for (int c = tid; c < kBM * (kBK / 4); c += 256) {
      const int r = c / (kBK / 4), q = c % (kBK / 4);
      const long long row = row0 + r;
      bool valid = row < nrows;
      if (valid) {
        const int n = (int)(row / p.Tmax), t = (int)(row % p.Tmax);
        valid = t < p.lens[p.G + n / p.B];
      }
      cp_async16(as + r * S + q * 4, A + (valid ? row : 0) * p.lda + k0 + q * 4, valid);
    }
    for (int c = tid; c < NC * (kBK / 4); c += 256) {
      const int r = c / (kBK / 4), q = c % (kBK / 4);
      cp_async16(ws + r * S + q * 4, W + (size_t)r * p.K + k0 + q * 4, true);
    }
    cp_async_commit();
  };

  float acc[MTW][NTW][4];
#pragma unroll
  for (int i = 0; i < MTW; ++i)
#pragma unroll
    for (int j = 0; j < NTW; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;

  load_slice(0, 0);
  for (int sl = 0; sl < total; ++sl) {
    if (sl + 1 < total) {
      load_slice(sl + 1, (sl + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* as = As + (sl & 1) * kBM * S + (wm * (kBM / WM)) * S;
    const float* ws = Ws + (sl & 1) * NC * S + (wn * (NC / WN)) * S;
#pragma unroll
    for (int kk = 0; kk < kBK / 16; ++kk) {
      uint32_t ah[MTW][4], al[MTW][4];
#pragma unroll
      for (int i = 0; i < MTW; ++i) {
        const float* a0 = as + (i * 16 + gq) * S + kk * 16 + 2 * tig;
        const float2 x0 = *reinterpret_cast<const float2*>(a0);
        const float2 x1 = *reinterpret_cast<const float2*>(a0 + 8 * S);
        const float2 x2 = *reinterpret_cast<const float2*>(a0 + 8);
        const float2 x3 = *reinterpret_cast<const float2*>(a0 + 8 * S + 8);
        if constexpr (SPLIT) {
          split_bf16(x0.x, x0.y, ah[i][0], al[i][0]);
          split_bf16(x1.x, x1.y, ah[i][1], al[i][1]);
          split_bf16(x2.x, x2.y, ah[i][2], al[i][2]);
          split_bf16(x3.x, x3.y, ah[i][3], al[i][3]);
        } else {
          ah[i][0] = pack_bf16(x0.x, x0.y);
          ah[i][1] = pack_bf16(x1.x, x1.y);
          ah[i][2] = pack_bf16(x2.x, x2.y);
          ah[i][3] = pack_bf16(x3.x, x3.y);
        }
      }
#pragma unroll
      for (int j = 0; j < NTW; ++j) {
        const float* b0 = ws + (j * 8 + gq) * S + kk * 16 + 2 * tig;
        const float2 y0 = *reinterpret_cast<const float2*>(b0);
        const float2 y1 = *reinterpret_cast<const float2*>(b0 + 8);
        uint32_t bh0, bh1, bl0 = 0u, bl1 = 0u;
        if constexpr (SPLIT) {
          split_bf16(y0.x, y0.y, bh0, bl0);
          split_bf16(y1.x, y1.y, bh1, bl1);
        } else {
          bh0 = pack_bf16(y0.x, y0.y);
          bh1 = pack_bf16(y1.x, y1.y);
        }
#pragma unroll
        for (int i = 0; i < MTW; ++i) {
          mma_bf16(acc[i][j], ah[i], bh0, bh1);
          if constexpr (SPLIT) {
            mma_bf16(acc[i][j], ah[i], bl0, bl1);
            mma_bf16(acc[i][j], al[i], bh0, bh1);
          }
        }
      }
    }
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < MTW; ++i) {
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const long long row = row0 + wm * (kBM / WM) + i * 16 + gq + hrow * 8;
      if (row >= nrows) continue;
      const int n = (int)(row / p.Tmax), t = (int)(row % p.Tmax);
      if (t >= p.lens[p.G + n / p.B]) continue;
      float* crow = p.C + row * p.ldc;
#pragma unroll
      for (int j = 0; j < NTW; ++j) {
        const int col = wn * (NC / WN) + j * 8 + 2 * tig;
        float2 v = make_float2(acc[i][j][hrow * 2 + 0], acc[i][j][hrow * 2 + 1]);
        if (p.bias != nullptr) {
          v.x += p.bias[col];
          v.y += p.bias[col + 1];
        }
        float2* dst = reinterpret_cast<float2*>(crow + col);
        if (p.accumulate) {
          const float2 o = *dst;
          v.x += o.x;
          v.y += o.y;
        }
        *dst = v;
      }
    }
  }
}

template <int NC, int WM, int WN>
cudaError_t launch_nt(const GemmNTArgs& a, int precision, cudaStream_t st) {
  const size_t smem = (size_t)2 * (kBM + NC) * (kBK + kPad) * sizeof(float);
  const long long nrows = (long long)a.G * a.B * a.Tmax;
  const unsigned grid = (unsigned)((nrows + kBM - 1) / kBM);
  cudaError_t e;
  if (precision == 0) {
    e = cudaFuncSetAttribute(gemm_nt_kernel<NC, WM, WN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    gemm_nt_kernel<NC, WM, WN, true><<<grid, 256, smem, st>>>(a);
  } else {
    e = cudaFuncSetAttribute(gemm_nt_kernel<NC, WM, WN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    gemm_nt_kernel<NC, WM, WN, false><<<grid, 256, smem, st>>>(a);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------------
// TN
// ------------------------------------------------------------------------------------------------------------------------
template <int KA, int NB, bool SPLIT>
__global__ void __launch_bounds__(256, 1) gemm_tn_kernel(const GemmTNArgs p) {
  constexpr int WM = 4, WN = 2, MTW = KA / WM / 16, NTW = NB / WN / 8;
  constexpr int SA = KA + 4, SB = NB + 4;  // strides == 4 (mod 32): conflict-free transposed fragment reads
  constexpr int BPT = kBK * NB / 4 / 256;  // float4 of the B tile per thread (gathered path)
  static_assert(kBK * NB / 4 % 256 == 0 || kBK * NB / 4 < 256, "B tile split");
  extern __shared__ __align__(16) float smem[];
  float* At = smem;                 // [2][kBK][SA]
  float* Bt = smem + 2 * kBK * SA;  // [2][kBK][SB]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tig = lane & 3;
  const int wm = warp / WN, wn = warp % WN;
  const int g = blockIdx.y, cta = blockIdx.x;
  const int T = p.lens[p.G + g];
  const int tiles_per_seq = (T + kBK - 1) / kBK;
  const int items = p.B * tiles_per_seq;
  const bool gathered = p.tok != nullptr;
  const int lda = p.lda > 0 ? p.lda : KA, a_col0 = p.lda > 0 ? p.a_col0 : 0;
  const int emb_ld = p.emb_ld > 0 ? p.emb_ld : NB, emb_col0 = p.emb_ld > 0 ? p.emb_col0 : 0;

  float acc[MTW][NTW][4];
#pragma unroll
  for (int i = 0; i < MTW; ++i)
#pragma unroll
    for (int j = 0; j < NTW; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;
  float csum = 0.f;

  auto issue_async = [&](int item, int buf) {
    const int n = g * p.B + item / tiles_per_seq, t0 = (item % tiles_per_seq) * kBK;
    float* at = At + buf * kBK * SA;
    for (int c = tid; c < kBK * (KA / 4); c += 256) {
      const int r = c / (KA / 4), q = c % (KA / 4);
      const bool valid = t0 + r < T;
      cp_async16(at + r * SA + q * 4, p.A + ((size_t)n * p.Tmax + (valid ? t0 + r : 0)) * lda + a_col0 + q * 4, valid);
    }
    if (!gathered) {
      float* bt = Bt + buf * kBK * SB;
      for (int c = tid; c < kBK * (NB / 4); c += 256) {
        const int r = c / (NB / 4), q = c % (NB / 4);
        const int ts = t0 + r + p.shift;
        const bool valid = (t0 + r < T) && ts >= 0 && ts < T;
        cp_async16(bt + r * SB + q * 4, p.Bsrc + ((size_t)n * p.Tmax + (valid ? ts : 0)) * p.ldb + p.col0 + q * 4, valid);
      }
    }
    cp_async_commit();
  };
  // gathered B operand (layer-0 input = masked embedding rows): global -> registers now, registers -> smem after the MMAs
  float4 breg[BPT > 0 ? BPT : 1];
  auto gather_load = [&](int item) {
    const int n = g * p.B + item / tiles_per_seq, t0 = (item % tiles_per_seq) * kBK;
#pragma unroll
    for (int i = 0; i < (BPT > 0 ? BPT : 1); ++i) {
      const int c = tid + i * 256;
      breg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < kBK * (NB / 4)) {
        const int r = c / (NB / 4), q = c % (NB / 4);
        if (t0 + r < T) {
          const int tk = p.tok[(size_t)n * p.Tmax + t0 + r];
          const float sc = p.emb_row_scale != nullptr ? p.emb_row_scale[(size_t)g * p.V + tk] : 1.0f;
          float4 e = *reinterpret_cast<const float4*>(p.emb + (size_t)tk * emb_ld + emb_col0 + q * 4);
          breg[i] = make_float4(sc * e.x, sc * e.y, sc * e.z, sc * e.w);
        }
      }
    }
  };
  auto gather_store = [&](int buf) {
    float* bt = Bt + buf * kBK * SB;
#pragma unroll
    for (int i = 0; i < (BPT > 0 ? BPT : 1); ++i) {
      const int c = tid + i * 256;
      if (c < kBK * (NB / 4)) {
        const int r = c / (NB / 4), q = c % (NB / 4);
        *reinterpret_cast<float4*>(bt + r * SB + q * 4) = breg[i];
      }
    }
  };

  int it = cta, buf = 0;
  if (it < items) {
    issue_async(it, 0);
    if (gathered) {
      gather_load(it);
      gather_store(0);
    }
  }
  for (; it < items; it += p.ctas_per_group, buf ^= 1) {
    const int nxt = it + p.ctas_per_group;
    if (nxt < items) {
      issue_async(nxt, buf ^ 1);
      if (gathered) gather_load(nxt);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* at = At + buf * kBK * SA;
    const float* bt = Bt + buf * kBK * SB;
    if (p.colsum && tid < KA) {
#pragma unroll 8
      for (int r = 0; r < kBK; ++r) csum += at[r * SA + tid];
    }
#pragma unroll
    for (int kk = 0; kk < kBK / 16; ++kk) {
      // B fragments for this warp's n tiles: b0b1 = (k=2tig,2tig+1 ; n=gq), b2b3 = (k+8)
      uint32_t bh[NTW][2], bl[NTW][2];
#pragma unroll
      for (int j = 0; j < NTW; ++j) {
        const float* b = bt + (kk * 16 + 2 * tig) * SB + wn * (NB / WN) + j * 8 + gq;
        const float y0 = b[0], y1 = b[SB], y2 = b[8 * SB], y3 = b[9 * SB];
        if constexpr (SPLIT) {
          split_bf16(y0, y1, bh[j][0], bl[j][0]);
          split_bf16(y2, y3, bh[j][1], bl[j][1]);
        } else {
          bh[j][0] = pack_bf16(y0, y1);
          bh[j][1] = pack_bf16(y2, y3);
        }
      }
#pragma unroll
      for (int i = 0; i < MTW; ++i) {
        // A fragment (transposed read): rows = gate columns, k = token rows
        const float* a = at + (kk * 16 + 2 * tig) * SA + wm * (KA / WM) + i * 16 + gq;
        const float x00 = a[0], x01 = a[SA];                   // (m=gq   ; k=2tig,2tig+1)
        const float x10 = a[8], x11 = a[SA + 8];               // (m=gq+8 ; k=2tig,2tig+1)
        const float x20 = a[8 * SA], x21 = a[9 * SA];          // (m=gq   ; k=2tig+8,+9)
        const float x30 = a[8 * SA + 8], x31 = a[9 * SA + 8];  // (m=gq+8 ; k=2tig+8,+9)
        uint32_t ah[4], al[4];
        if constexpr (SPLIT) {
          split_bf16(x00, x01, ah[0], al[0]);
          split_bf16(x10, x11, ah[1], al[1]);
          split_bf16(x20, x21, ah[2], al[2]);
          split_bf16(x30, x31, ah[3], al[3]);
        } else {
          ah[0] = pack_bf16(x00, x01);
          ah[1] = pack_bf16(x10, x11);
          ah[2] = pack_bf16(x20, x21);
          ah[3] = pack_bf16(x30, x31);
        }
#pragma unroll
        for (int j = 0; j < NTW; ++j) {
          mma_bf16(acc[i][j], ah, bh[j][0], bh[j][1]);
          if constexpr (SPLIT) {
            mma_bf16(acc[i][j], ah, bl[j][0], bl[j][1]);
            mma_bf16(acc[i][j], al, bh[j][0], bh[j][1]);
          }
        }
      }
    }
    if (gathered && nxt < items) gather_store(buf ^ 1);
    __syncthreads();
  }

  float* out = p.partial + ((size_t)g * p.ctas_per_group + cta) * ((size_t)KA * NB + (p.colsum ? KA : 0));
#pragma unroll
  for (int i = 0; i < MTW; ++i)
#pragma unroll
    for (int j = 0; j < NTW; ++j) {
      const int r = wm * (KA / WM) + i * 16 + gq, c = wn * (NB / WN) + j * 8 + 2 * tig;
      *reinterpret_cast<float2*>(out + (size_t)r * NB + c) = make_float2(acc[i][j][0], acc[i][j][1]);
      *reinterpret_cast<float2*>(out + (size_t)(r + 8) * NB + c) = make_float2(acc[i][j][2], acc[i][j][3]);
    }
  if (p.colsum && tid < KA) out[(size_t)KA * NB + tid] = csum;
}

template <int KA, int NB>
cudaError_t launch_tn(const GemmTNArgs& a, int precision, cudaStream_t st) {
  const size_t smem = (size_t)2 * kBK * (KA + 4 + NB + 4) * sizeof(float);
  dim3 grid(a.ctas_per_group, a.G);
  cudaError_t e;
  if (precision == 0) {
    e = cudaFuncSetAttribute(gemm_tn_kernel<KA, NB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    gemm_tn_kernel<KA, NB, true><<<grid, 256, smem, st>>>(a);
  } else {
    e = cudaFuncSetAttribute(gemm_tn_kernel<KA, NB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    gemm_tn_kernel<KA, NB, false><<<grid, 256, smem, st>>>(a);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------------------------------
// one thread per 4 consecutive output elements: adjacent threads read adjacent float4 of every (group, cta) partial (coalesced),
// sum them in a fixed order (deterministic); the per-group mask is applied before the cross-group sum
__global__ void __launch_bounds__(128) dw_reduce_kernel(const DwReduceArgs p) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const bool any_cs = p.has_colsum || p.cs_ptr != nullptr;
  const int nmat4 = p.KA * p.NB / 4, ncs4 = any_cs ? p.KA / 4 : 0;
  if (w >= nmat4 + ncs4) return;
  const size_t blk = (size_t)p.KA * p.NB + (p.has_colsum ? p.KA : 0);
  const bool is_mat = w < nmat4;
  const int idx = w * 4;  // element offset inside a partial block (matrix part first, then the column sums)
  const int gi = is_mat ? idx / p.NB : (idx - p.KA * p.NB), c = is_mat ? idx % p.NB : 0;
  const int ldo = p.ldo > 0 ? p.ldo : p.NB;  // chunked GEMMs: this block is rows [gi0, gi0+KA) x columns [c0, c0+NB) of [4H, ldo]
  float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!is_mat && p.cs_ptr != nullptr) {
    // bias partials written by the BPTT kernel: [cs_count][KA]
#pragma unroll 8
    for (int k = 0; k < p.cs_count; ++k) {
      const float4 v = *reinterpret_cast<const float4*>(p.cs_ptr + (size_t)k * p.KA + gi);
      tot.x += v.x; tot.y += v.y; tot.z += v.z; tot.w += v.w;
    }
  } else {
    for (int g = 0; g < p.G; ++g) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8  // (8 partial blocks in flight per thread; the additions keep their fixed order)
      for (int k = 0; k < p.ctas_per_group; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(p.partial + ((size_t)g * p.ctas_per_group + k) * blk + idx);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      if (is_mat && p.mask != nullptr && c >= (p.out2 != nullptr ? p.NB1 : 0)) {
        const int mc = p.out2 != nullptr ? c - p.NB1 : p.c0 + c, mld = p.out2 != nullptr ? p.NB - p.NB1 : ldo;
        const float4 m = *reinterpret_cast<const float4*>(p.mask + ((size_t)g * 4 * p.H + gi_to_torch_row(p.gi0 + gi, p.H)) * mld + mc);
        s.x *= m.x; s.y *= m.y; s.z *= m.z; s.w *= m.w;
      }
      tot.x += s.x; tot.y += s.y; tot.z += s.z; tot.w += s.w;
    }
  }
  if (is_mat) {
    const int row = gi_to_torch_row(p.gi0 + gi, p.H);
    if (p.out2 == nullptr) *reinterpret_cast<float4*>(p.out + (size_t)row * ldo + p.c0 + c) = tot;
    else if (c < p.NB1) *reinterpret_cast<float4*>(p.out + (size_t)row * p.NB1 + c) = tot;
    else *reinterpret_cast<float4*>(p.out2 + (size_t)row * (p.NB - p.NB1) + (c - p.NB1)) = tot;
  } else {
    const float t4[4] = {tot.x, tot.y, tot.z, tot.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // gi+j are the 4 gates of one unit: rows q*H + u
      const int row = gi_to_torch_row(p.gi0 + gi + j, p.H);
      if (p.out_b1 != nullptr) p.out_b1[row] = t4[j];
      if (p.out_b2 != nullptr) p.out_b2[row] = t4[j];
    }
  }
}

// each CTA owns a contiguous slice of one group's token rows, accumulates into a shared-memory copy of the [V,H] table with
// shared atomics (spread addresses), then flushes the non-zero rows once with global atomics.  Falls back to direct global
// atomics when the table does not fit in shared memory.
__global__ void emb_grad_kernel(const EmbGradArgs p, int ctas_per_group, int use_smem) {
  extern __shared__ float tab[];  // [V*H]
  const int g = blockIdx.y, cta = blockIdx.x;
  const int T = p.lens[p.G + g];
  const int VH = p.V * p.H;
  if (use_smem) {
    for (int i = threadIdx.x; i < VH; i += blockDim.x) tab[i] = 0.f;
    __syncthreads();
  }
  const long long rows = (long long)p.B * T;  // live rows of this group
  const long long per = (rows + ctas_per_group - 1) / ctas_per_group;
  const long long lo = cta * per, hi = min(rows, lo + per);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (long long r = lo + warp; r < hi; r += nwarp) {
    const int n = g * p.B + (int)(r / T), t = (int)(r % T);
    const long long row = (long long)n * p.Tmax + t;
    const int tk = p.tok[row];
    if (tk == 0) continue;  // padding_idx=0 receives no gradient (nn.Embedding(..., padding_idx=0), e2e_triplet.py:345)
    const float sc = p.emb_row_scale != nullptr ? p.emb_row_scale[(size_t)g * p.V + tk] : 1.0f;
    if (sc == 0.0f) continue;
    float* dst = (use_smem ? tab : p.demb) + (size_t)tk * p.H;
    for (int c = lane; c < p.H; c += 32) atomicAdd(dst + c, sc * p.dx[row * p.H + c]);
  }
  if (use_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < VH; i += blockDim.x) {
      const float v = tab[i];
      if (v != 0.f) atomicAdd(p.demb + i, v);
    }
  }
}

__global__ void prep_wih_kernel(const float* __restrict__ w, const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                int H, int K, float* __restrict__ out_w, float* __restrict__ out_wT, float* __restrict__ out_b) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 4 * H * K) return;
  const int gi = idx / K, k = idx % K, row = gi_to_torch_row(gi, H);
  const float v = w[(size_t)row * K + k];
  if (out_w != nullptr) out_w[idx] = v;
  if (out_wT != nullptr) out_wT[(size_t)k * 4 * H + gi] = v;
  if (out_b != nullptr && k == 0) out_b[gi] = b_ih[row] + b_hh[row];
}

__global__ void fill_zero_kernel(float* p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 0.f;
}

}  // namespace

cudaError_t launch_gemm_nt(const GemmNTArgs& a, int precision, cudaStream_t st) {
  if (a.K % kBK != 0 || a.lda % 4 != 0 || a.ldc % 2 != 0 || a.NC % 32 != 0) return cudaErrorInvalidValue;
  // wide outputs (NC = 4H, 2H for H > 64) are produced in column blocks of 256 / 128 / 64 / 32: W rows, bias and C shift together
  for (int c0 = 0; c0 < a.NC;) {
    const int rem = a.NC - c0, nc = rem >= 256 ? 256 : (rem >= 128 ? 128 : (rem >= 64 ? 64 : 32));
    GemmNTArgs b = a;
    b.NC = nc;
    b.C = a.C + c0;
    b.bias = a.bias != nullptr ? a.bias + c0 : nullptr;
    for (int s = 0; s < a.nsrc; ++s) b.W[s] = a.W[s] + (size_t)c0 * a.K;
    cudaError_t e;
    switch (nc) {
      case 256: e = launch_nt<256, 2, 4>(b, precision, st); break;
      case 128: e = launch_nt<128, 4, 2>(b, precision, st); break;
      case 64: e = launch_nt<64, 8, 1>(b, precision, st); break;
      default: e = launch_nt<32, 8, 1>(b, precision, st); break;
    }
    if (e != cudaSuccess) return e;
    c0 += nc;
  }
  return cudaSuccess;
}

cudaError_t launch_gemm_tn(const GemmTNArgs& a, int precision, cudaStream_t st) {
  if (a.NB1 != a.NB) return cudaErrorInvalidValue;  // the two-source B operand exists in the tcgen05 kernel only
  if (a.KA == 256 && a.NB == 128) return launch_tn<256, 128>(a, precision, st);
  if (a.KA == 256 && a.NB == 64) return launch_tn<256, 64>(a, precision, st);
  if (a.KA == 256 && a.NB == 32) return launch_tn<256, 32>(a, precision, st);
  if (a.KA == 128 && a.NB == 128) return launch_tn<128, 128>(a, precision, st);
  if (a.KA == 128 && a.NB == 64) return launch_tn<128, 64>(a, precision, st);
  if (a.KA == 128 && a.NB == 32) return launch_tn<128, 32>(a, precision, st);
  return cudaErrorInvalidValue;
}

cudaError_t launch_dw_reduce(const DwReduceArgs& a, cudaStream_t st) {
  const int quads = (a.KA * a.NB + ((a.has_colsum || a.cs_ptr != nullptr) ? a.KA : 0)) / 4;
  dw_reduce_kernel<<<(quads + 127) / 128, 128, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_emb_grad(const EmbGradArgs& a, cudaStream_t st) {
  cudaError_t e = launch_fill_zero(a.demb, (size_t)a.V * a.H, st);
  if (e != cudaSuccess) return e;
  const size_t smem = (size_t)a.V * a.H * sizeof(float);
  const int use_smem = 0;  // measured on B200: direct global atomics (0.12 ms) beat the shared-memory privatised variant (0.14-0.5 ms)
  (void)smem;
  const int cpg = std::max(1, 296 / a.G);
  emb_grad_kernel<<<dim3(cpg, a.G), 512, 0, st>>>(a, cpg, use_smem);
  return cudaGetLastError();
}

cudaError_t launch_prep_wih(const float* w, const float* b_ih, const float* b_hh, int H, int K, float* out_w, float* out_wT,
                            float* out_b, cudaStream_t st) {
  const int total = 4 * H * K;
  prep_wih_kernel<<<(total + 255) / 256, 256, 0, st>>>(w, b_ih, b_hh, H, K, out_w, out_wT, out_b);
  return cudaGetLastError();
}

cudaError_t launch_fill_zero(float* p, size_t n, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  const unsigned grid = (unsigned)min((size_t)1184, (n + 255) / 256);
  fill_zero_kernel<<<grid, 256, 0, st>>>(p, n);
  return cudaGetLastError();
}

}  // namespace ib200
