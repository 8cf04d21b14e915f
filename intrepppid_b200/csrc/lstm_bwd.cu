// K3: persistent BPTT kernel (one CTA = 8 sequences of one group, one direction, all T_eff steps in reverse scan order).
//
// Backward of the recurrence in nn.LSTM (reference: autograd through encoders/awd_lstm.py:56).  Per step, for each cell:
//   dh = dy_t + W_hh^T da_{t'} (from the previously processed step) ; tc = tanh(c_t)
//   do = dh*tc ; dc += dh*o*(1-tc^2) ; di = dc*g ; dg = dc*i ; df = dc*c_{prev} ; dc <- dc*f
//   da = (di*i(1-i), df*f(1-f), dg*(1-g^2), do*o(1-o))
// da overwrites the saved gates IN PLACE (same float4 slot, gate-interleaved order) and is the only output: dW_ih, dW_hh,
// db and dX are tensor-core GEMMs over the da stream afterwards (gemm.cu), off the dependent chain.
//
// Mapping (H=64: 8 warps): dh^T[H,8] = W_hh^T[H,4H] * da[4H,8] with mma.m16n8k16; W_hh^T lives in registers (A operand), each warp
// owns one 16-unit output tile and one half of K (two partial sums, exchanged through smem).  The K order is chosen so that
// the B fragment of lane (col n, unit%4) is exactly the cell's (da_i,da_f | da_g,da_o) pair of packed bf16x2 words.
// Two __syncthreads per step.  fp32 mode: bf16 hi/lo split, 3 MMAs per product.
#include "kernels.h"

namespace ib200 {
namespace {

constexpr int kPF = 2;

struct CellIn {
  float4 g;     // saved (i,f,g,o)
  float cprev;  // c at the previous scan position (0 at the chain start)
  float dy;     // upstream gradient of h_t
};

template <int H, bool SPLIT, bool FAST_ACT>
__global__ void __launch_bounds__(H * 4, 1) lstm_bwd_kernel(const LstmBwdArgs p) {
  constexpr int NW = H / 8, MT = H / 16, KTT = H / 4, KTH = KTT / 2;
  constexpr int NPART = SPLIT ? 2 : 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tig = lane & 3;
  const int mt = warp % MT, kh = warp / MT;
  const int g = blockIdx.y, dir = p.dir0 + (int)blockIdx.z;
  const int T = p.lens[p.G + g];
  if (T <= 0) return;
  const int b0 = blockIdx.x * kBC;
  const int nvalid = min(kBC, p.B - b0);
  const int nbase = g * p.B + b0;
  const int Tmax = p.Tmax;
  const size_t N = (size_t)p.G * p.B;

  __shared__ __align__(16) uint32_t dafrag[NPART][KTT][32][2];
  __shared__ __align__(8) float2 xch[NW][32];

  // ---- A fragments: W_hh^T, K ordered as (unit, gate) with the fragment positions {2tig,2tig+1,2tig+8,2tig+9} = gates i,f,g,o ----
  uint32_t Ahi[KTH][4], Alo[KTH][4];
  {
    const float* __restrict__ W = p.whh[dir];
    const float* __restrict__ M = (dir == 0 && p.whh_mask != nullptr) ? p.whh_mask + (size_t)g * 4 * H * H : nullptr;
    const int j0 = 16 * mt + gq, j1 = j0 + 8;
#pragma unroll
    for (int ktl = 0; ktl < KTH; ++ktl) {
      const int uu = 4 * (kh * KTH + ktl) + tig;
      const int jj[4] = {j0, j1, j0, j1};
      const int qa[4] = {0, 0, 2, 2};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int ia = (qa[r] * H + uu) * H + jj[r], ib = ((qa[r] + 1) * H + uu) * H + jj[r];
        float w0 = W[ia], w1 = W[ib];
        if (M != nullptr) {
          w0 *= M[ia];
          w1 *= M[ib];
        }
        if constexpr (SPLIT) {
          split_bf16(w0, w1, Ahi[ktl][r], Alo[ktl][r]);
        } else {
          Ahi[ktl][r] = pack_bf16(w0, w1);
          Alo[ktl][r] = 0u;
        }
      }
    }
  }

  // ---- the two cells this thread owns -----------------------------------------------------------------------------------------
  const int j = 16 * mt + gq + 8 * kh;
  const int n0 = 2 * tig, n1 = n0 + 1;
  const bool v0 = n0 < nvalid, v1 = n1 < nvalid;
  float4* __restrict__ G4 = reinterpret_cast<float4*>(p.gates[dir]);
  const float* __restrict__ C = p.cstate[dir];

  auto time_of = [&](int s) { return dir ? s : (T - 1 - s); };  // reverse of the forward scan order
  auto fetch = [&](int s, CellIn (&q)[2]) {
    q[0].g = make_float4(0.f, 0.f, 0.f, 0.f);
    q[0].cprev = 0.f;
    q[0].dy = 0.f;
    q[1] = q[0];
    if (s < T) {
      const int t = time_of(s);
      const int tp = dir ? t + 1 : t - 1;            // scan predecessor of t in the forward pass
      const bool has_prev = dir ? (t + 1 < T) : (t > 0);
      const bool vv[2] = {v0, v1};
      const int nn[2] = {n0, n1};
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (vv[c]) {
          const size_t row = (size_t)(nbase + nn[c]) * Tmax + t;
          q[c].g = G4[row * H + j];
          if (has_prev) q[c].cprev = __ldg(C + ((size_t)(nbase + nn[c]) * Tmax + tp) * H + j);
          if (p.dy != nullptr) q[c].dy = __ldg(p.dy + row * p.dy_stride + dir * H + j);
        }
      }
    }
  };

  CellIn q[kPF][2];
  fetch(0, q[0]);
  fetch(1, q[1]);
  float ccur[2] = {0.f, 0.f}, dc[2] = {0.f, 0.f}, dhrec[2] = {0.f, 0.f};
  {
    const int t = time_of(0);
    if (v0) ccur[0] = C[((size_t)(nbase + n0) * Tmax + t) * H + j];
    if (v1) ccur[1] = C[((size_t)(nbase + n1) * Tmax + t) * H + j];
    if (p.dhn != nullptr) {
      if (v0) dhrec[0] = p.dhn[((size_t)dir * N + nbase + n0) * H + j];
      if (v1) dhrec[1] = p.dhn[((size_t)dir * N + nbase + n1) * H + j];
    }
  }

  auto step = [&](const int s, CellIn (&slot)[2]) {
    const int t = time_of(s);
    CellIn in[2] = {slot[0], slot[1]};
    fetch(s + kPF, slot);

    const int nn[2] = {n0, n1};
    const bool vv[2] = {v0, v1};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float gi = in[c].g.x, gf = in[c].g.y, gg = in[c].g.z, go = in[c].g.w;
      const float dh = dhrec[c] + in[c].dy;
      const float tc = tanh_f<FAST_ACT>(ccur[c]);
      const float d_o = dh * tc;
      const float dct = fmaf(dh * go, 1.0f - tc * tc, dc[c]);
      const float d_i = dct * gg, d_g = dct * gi, d_f = dct * in[c].cprev;
      dc[c] = dct * gf;
      ccur[c] = in[c].cprev;
      const float da_i = d_i * gi * (1.0f - gi), da_f = d_f * gf * (1.0f - gf);
      const float da_g = d_g * (1.0f - gg * gg), da_o = d_o * go * (1.0f - go);
      if (vv[c]) G4[((size_t)(nbase + nn[c]) * Tmax + t) * H + j] = make_float4(da_i, da_f, da_g, da_o);
      // B-fragment slot of this cell: k tile j/4, lane (col*4 + j%4)
      uint32_t hi0, hi1, lo0 = 0u, lo1 = 0u;
      if constexpr (SPLIT) {
        split_bf16(da_i, da_f, hi0, lo0);
        split_bf16(da_g, da_o, hi1, lo1);
      } else {
        hi0 = pack_bf16(da_i, da_f);
        hi1 = pack_bf16(da_g, da_o);
      }
      const int fl = nn[c] * 4 + (j & 3);
      *reinterpret_cast<uint2*>(&dafrag[0][j >> 2][fl][0]) = make_uint2(hi0, hi1);
      if constexpr (SPLIT) *reinterpret_cast<uint2*>(&dafrag[NPART - 1][j >> 2][fl][0]) = make_uint2(lo0, lo1);
    }
    __syncthreads();  // (A) all da of this step are in smem

    if (s + 1 < T) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f}, acs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ktl = 0; ktl < KTH; ++ktl) {
        const int kt = kh * KTH + ktl;
        const uint2 bh = *reinterpret_cast<const uint2*>(&dafrag[0][kt][lane][0]);
        mma_bf16(acc, Ahi[ktl], bh.x, bh.y);
        if constexpr (SPLIT) {
          const uint2 bl = *reinterpret_cast<const uint2*>(&dafrag[NPART - 1][kt][lane][0]);
          mma_bf16(acs, Ahi[ktl], bl.x, bl.y);
          mma_bf16(acs, Alo[ktl], bh.x, bh.y);
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] += acs[r];
      // acc: [0]=(unit 16mt+gq, col n0) [1]=(.., n1) [2]=(unit 16mt+gq+8, n0) [3]=(.., n1).  Keep the rows of my unit, hand
      // the other two to the partner warp (same mt, other K half), which owns that unit.
      const float2 mine = kh ? make_float2(acc[2], acc[3]) : make_float2(acc[0], acc[1]);
      const float2 theirs = kh ? make_float2(acc[0], acc[1]) : make_float2(acc[2], acc[3]);
      xch[warp][lane] = theirs;
      __syncthreads();  // (B)
      const float2 other = xch[kh ? warp - MT : warp + MT][lane];
      dhrec[0] = mine.x + other.x;
      dhrec[1] = mine.y + other.y;
    }
  };

  for (int s = 0; s < T; s += kPF) {
    step(s, q[0]);
    if (s + 1 < T) step(s + 1, q[1]);
  }
}

}  // namespace

cudaError_t launch_lstm_bwd(const LstmBwdArgs& a, int H, int precision, cudaStream_t st) {
  dim3 grid((a.B + kBC - 1) / kBC, a.G, a.ndir);
  if (H == 64) {
    if (precision == 0) lstm_bwd_kernel<64, true, false><<<grid, 256, 0, st>>>(a);
    else lstm_bwd_kernel<64, false, true><<<grid, 256, 0, st>>>(a);
  } else if (H == 32) {
    if (precision == 0) lstm_bwd_kernel<32, true, false><<<grid, 128, 0, st>>>(a);
    else lstm_bwd_kernel<32, false, true><<<grid, 128, 0, st>>>(a);
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace ib200
