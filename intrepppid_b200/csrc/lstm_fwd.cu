// K2: persistent recurrent-cell forward kernel (one CTA = 8 sequences of one group, one direction, all T_eff steps).
//
// Replaces the recurrence inside nn.LSTM -> _VF.lstm (reference: encoders/awd_lstm.py:35-41,56 via
// utils/weightdrop.py:109-111).  Per step:  a = xproj_t + W_hh h_{t-1};  (i,f,g,o) = (s,s,tanh,s)(a);
// c = f*c + i*g;  h = o*tanh(c).  Pads are stepped through (no packing); the scan covers t in [0,T_eff[g]).
//
// Mapping (H = 64: 8 warps, 256 threads):
//   * gates^T[4H, 8] = W_hh[4H, H] * h^T[H, 8] with mma.m16n8k16 (bf16 operands, fp32 accumulate).  W_hh is the A operand and
//     lives in REGISTERS for the whole kernel (per warp two 16-row tiles: rows (i_u, f_u) and (g_u, o_u) for its 8 units u),
//     so after the MMAs each thread holds i,f,g,o of one unit for two sequences: no shuffles, no smem for gates.
//   * fp32 mode: operands are split into bf16 hi + lo and three MMAs (hi*hi, hi*lo, lo*hi) are issued per product
//     (operand error ~2^-17; measured end-to-end error at T=1500 ~2e-6).  bf16 mode: one MMA.
//   * h is exchanged through a double-buffered smem tile (bf16 hi/lo, padded rows => conflict-free B-fragment loads);
//     ONE __syncthreads per step.
//   * the input projection is consumed as one float4 per cell (gate-interleaved layout): for layer 0 it is gathered from the
//     per-group table P[g][dir][token] (K1, the V x 4H lookup-table identity); for layers >= 1 it is the dense
//     xproj tensor written by the tensor-core GEMM.  It is prefetched two steps ahead into registers.
#include "kernels.h"

namespace ib200 {

namespace {

constexpr int kTokChunk = 32;  // time steps of token ids staged in smem per refill (layer 0)
constexpr int kPF = 2;         // xproj prefetch distance in steps

template <int H, bool SPLIT, bool FAST_ACT, bool LAYER0, bool TRAIN>
__global__ void __launch_bounds__(H * 4, 1) lstm_fwd_kernel(const LstmFwdArgs p) {
  constexpr int KT = H / 16;       // k tiles over the hidden units
  constexpr int HS = H + 8;        // padded row stride (bf16) of the h tile: conflict-free fragment loads
  constexpr int NPART = SPLIT ? 2 : 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tig = lane & 3;
  const int g = blockIdx.y, dir = p.dir0 + (int)blockIdx.z;
  const int T = p.lens[p.G + g];  // T_eff of this group
  if (T <= 0) return;
  const int b0 = blockIdx.x * kBC;
  const int nvalid = min(kBC, p.B - b0);
  const int nbase = g * p.B + b0;  // first global sequence index of this CTA
  const int Tmax = p.Tmax;

  __shared__ __align__(16) __nv_bfloat16 hs[2][NPART][kBC][HS];
  __shared__ int toks[2][kBC][kTokChunk];

  // ---- A fragments: W_hh rows of this warp's 8 units, masked per group for (layer 0, forward) -----------------------------
  const int u = warp * 8 + gq;  // the unit this thread owns
  uint32_t Ahi[2][KT][4], Alo[2][KT][4];
  {
    const float* __restrict__ W = p.whh[dir];
    const float* __restrict__ M = (dir == 0 && p.whh_mask != nullptr) ? p.whh_mask + (size_t)g * 4 * H * H : nullptr;
#pragma unroll
    for (int tile = 0; tile < 2; ++tile) {
      const int r0 = (2 * tile) * H + u;      // fragment rows gq     : gate i (tile 0) / g (tile 1)
      const int r1 = (2 * tile + 1) * H + u;  // fragment rows gq + 8 : gate f (tile 0) / o (tile 1)
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        const int k0 = kt * 16 + 2 * tig;
        const int rr[4] = {r0, r1, r0, r1};
        const int kk[4] = {k0, k0, k0 + 8, k0 + 8};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float w0 = W[rr[j] * H + kk[j]], w1 = W[rr[j] * H + kk[j] + 1];
          if (M != nullptr) {
            w0 *= M[rr[j] * H + kk[j]];
            w1 *= M[rr[j] * H + kk[j] + 1];
          }
          if constexpr (SPLIT) {
            split_bf16(w0, w1, Ahi[tile][kt][j], Alo[tile][kt][j]);
          } else {
            Ahi[tile][kt][j] = pack_bf16(w0, w1);
            Alo[tile][kt][j] = 0u;
          }
        }
      }
    }
  }

  // ---- init h = 0 (both buffers), token chunks 0 and 1 ---------------------------------------------------------------------
  for (int i = tid; i < 2 * NPART * kBC * HS; i += blockDim.x) (&hs[0][0][0][0])[i] = __float2bfloat16(0.0f);
  auto load_tok_chunk = [&](int c) {
    if constexpr (LAYER0) {
      for (int i = tid; i < kBC * kTokChunk; i += blockDim.x) {
        const int n = i / kTokChunk, ss = i % kTokChunk, s = c * kTokChunk + ss;
        int v = 0;
        if (s < T && n < nvalid) {
          const int t = dir ? (T - 1 - s) : s;
          v = p.tok[(size_t)(nbase + n) * Tmax + t];
        }
        toks[c & 1][n][ss] = v;
      }
    }
  };
  load_tok_chunk(0);
  load_tok_chunk(1);
  __syncthreads();

  const int n0 = 2 * tig, n1 = 2 * tig + 1;  // the two sequences (columns) this thread owns
  const bool v0 = n0 < nvalid, v1 = n1 < nvalid;
  const float4* __restrict__ xsrc =
      LAYER0 ? reinterpret_cast<const float4*>(p.table) + (size_t)(g * 2 + dir) * p.V * H
             : reinterpret_cast<const float4*>(p.xproj[dir]);

  auto fetch_x = [&](int s, float4 (&x)[2]) {
    x[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    x[1] = x[0];
    if (s < T) {
      const int t = dir ? (T - 1 - s) : s;
      if constexpr (LAYER0) {
        const int c = (s / kTokChunk) & 1, ss = s % kTokChunk;
        if (v0) x[0] = __ldg(xsrc + (size_t)toks[c][n0][ss] * H + u);
        if (v1) x[1] = __ldg(xsrc + (size_t)toks[c][n1][ss] * H + u);
      } else {
        if (v0) x[0] = __ldg(xsrc + ((size_t)(nbase + n0) * Tmax + t) * H + u);
        if (v1) x[1] = __ldg(xsrc + ((size_t)(nbase + n1) * Tmax + t) * H + u);
      }
    }
  };

  float4 xq[kPF][2];
  fetch_x(0, xq[0]);
  fetch_x(1, xq[1]);

  float c0 = 0.f, c1 = 0.f, h0 = 0.f, h1 = 0.f;

  auto step = [&](const int s, float4 (&xslot)[2]) {
    const int t = dir ? (T - 1 - s) : s;
    const float4 x0 = xslot[0], x1 = xslot[1];
    if constexpr (LAYER0) {
      // refill the token ring: chunk (s/CH + 1) replaces chunk (s/CH - 1), whose last use was at step s - 1 - kPF
      if (s > 0 && (s % kTokChunk) == 0) load_tok_chunk(s / kTokChunk + 1);
    }
    fetch_x(s + kPF, xslot);

    // B fragments: h_{t-1}^T
    const __nv_bfloat16(*hb)[kBC][HS] = hs[s & 1];
    uint32_t bh[KT][2], bl[KT][2];
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      bh[kt][0] = *reinterpret_cast<const uint32_t*>(&hb[0][gq][kt * 16 + 2 * tig]);
      bh[kt][1] = *reinterpret_cast<const uint32_t*>(&hb[0][gq][kt * 16 + 2 * tig + 8]);
      if constexpr (SPLIT) {
        bl[kt][0] = *reinterpret_cast<const uint32_t*>(&hb[NPART - 1][gq][kt * 16 + 2 * tig]);
        bl[kt][1] = *reinterpret_cast<const uint32_t*>(&hb[NPART - 1][gq][kt * 16 + 2 * tig + 8]);
      }
    }
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    float acs[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};  // small cross terms (fp32 mode)
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
#pragma unroll
      for (int tile = 0; tile < 2; ++tile) {
        mma_bf16(acc[tile], Ahi[tile][kt], bh[kt][0], bh[kt][1]);
        if constexpr (SPLIT) {
          mma_bf16(acs[tile], Ahi[tile][kt], bl[kt][0], bl[kt][1]);
          mma_bf16(acs[tile], Alo[tile][kt], bh[kt][0], bh[kt][1]);
        }
      }
    }
    // accumulator layout: [0]=(row gq, col 2tig) [1]=(row gq, col 2tig+1) [2]=(row gq+8, col 2tig) [3]=(row gq+8, col 2tig+1)
    // tile 0 rows: gq -> i_u, gq+8 -> f_u ; tile 1 rows: gq -> g_u, gq+8 -> o_u
    float ai0 = acc[0][0] + acs[0][0] + x0.x, af0 = acc[0][2] + acs[0][2] + x0.y;
    float ag0 = acc[1][0] + acs[1][0] + x0.z, ao0 = acc[1][2] + acs[1][2] + x0.w;
    float ai1 = acc[0][1] + acs[0][1] + x1.x, af1 = acc[0][3] + acs[0][3] + x1.y;
    float ag1 = acc[1][1] + acs[1][1] + x1.z, ao1 = acc[1][3] + acs[1][3] + x1.w;

    const float i0 = sigmoid_f<FAST_ACT>(ai0), f0 = sigmoid_f<FAST_ACT>(af0), gg0 = tanh_f<FAST_ACT>(ag0),
                o0 = sigmoid_f<FAST_ACT>(ao0);
    const float i1 = sigmoid_f<FAST_ACT>(ai1), f1 = sigmoid_f<FAST_ACT>(af1), gg1 = tanh_f<FAST_ACT>(ag1),
                o1 = sigmoid_f<FAST_ACT>(ao1);
    c0 = fmaf(f0, c0, i0 * gg0);
    c1 = fmaf(f1, c1, i1 * gg1);
    h0 = o0 * tanh_f<FAST_ACT>(c0);
    h1 = o1 * tanh_f<FAST_ACT>(c1);

    // publish h_t for the next step
    __nv_bfloat16(*hn)[kBC][HS] = hs[(s + 1) & 1];
    {
      const __nv_bfloat16 h0h = __float2bfloat16_rn(h0), h1h = __float2bfloat16_rn(h1);
      hn[0][n0][u] = h0h;
      hn[0][n1][u] = h1h;
      if constexpr (SPLIT) {
        hn[NPART - 1][n0][u] = __float2bfloat16_rn(h0 - __bfloat162float(h0h));
        hn[NPART - 1][n1][u] = __float2bfloat16_rn(h1 - __bfloat162float(h1h));
      }
    }
    // stream out what later stages need
    const size_t row0 = (size_t)(nbase + n0) * Tmax + t, row1 = (size_t)(nbase + n1) * Tmax + t;
    if (p.y != nullptr) {
      if (v0) p.y[row0 * p.y_stride + dir * H + u] = h0;
      if (v1) p.y[row1 * p.y_stride + dir * H + u] = h1;
    }
    if constexpr (TRAIN) {
      float4* G4 = reinterpret_cast<float4*>(p.gates[dir]);
      float* C = p.cstate[dir];
      if (v0) {
        G4[row0 * H + u] = make_float4(i0, f0, gg0, o0);
        C[row0 * H + u] = c0;
      }
      if (v1) {
        G4[row1 * H + u] = make_float4(i1, f1, gg1, o1);
        C[row1 * H + u] = c1;
      }
    }
    __syncthreads();
  };

  for (int s = 0; s < T; s += kPF) {
    step(s, xq[0]);
    if (s + 1 < T) step(s + 1, xq[1]);
  }

  if (p.hn != nullptr) {
    const size_t N = (size_t)p.G * p.B;
    if (v0) p.hn[((size_t)dir * N + nbase + n0) * H + u] = h0;
    if (v1) p.hn[((size_t)dir * N + nbase + n1) * H + u] = h1;
  }
}

template <int H, bool SPLIT, bool FAST>
cudaError_t launch_h(const LstmFwdArgs& a, cudaStream_t st) {
  dim3 grid((a.B + kBC - 1) / kBC, a.G, a.ndir), block(H * 4);
  const bool l0 = a.tok != nullptr, tr = a.gates[a.dir0] != nullptr;
  if (l0 && tr) lstm_fwd_kernel<H, SPLIT, FAST, true, true><<<grid, block, 0, st>>>(a);
  else if (l0) lstm_fwd_kernel<H, SPLIT, FAST, true, false><<<grid, block, 0, st>>>(a);
  else if (tr) lstm_fwd_kernel<H, SPLIT, FAST, false, true><<<grid, block, 0, st>>>(a);
  else lstm_fwd_kernel<H, SPLIT, FAST, false, false><<<grid, block, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_lstm_fwd(const LstmFwdArgs& a, int H, int precision, cudaStream_t st) {
  if (H == 64) return precision == 0 ? launch_h<64, true, false>(a, st) : launch_h<64, false, true>(a, st);
  if (H == 32) return precision == 0 ? launch_h<32, true, false>(a, st) : launch_h<32, false, true>(a, st);
  return cudaErrorInvalidValue;
}

}  // namespace ib200
