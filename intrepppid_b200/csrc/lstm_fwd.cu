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
//     The accumulators are INITIALISED with the input projection, so the MMA result is the pre-activation.
//   * fp32 mode: operands are split into bf16 hi + lo and three MMAs (hi*hi, hi*lo, lo*hi) are issued per product, each on its
//     own accumulator chain (depth H/16) (operand error ~2^-17; end-to-end error at T=1500 ~1e-6).  bf16 mode: one MMA.
//   * h is exchanged through a double-buffered smem tile laid out [unit][sequence] (16-byte rows): the producer writes one
//     packed bf16x2 word per part, the consumer gets its B fragments with ldmatrix.x4.trans.  ONE __syncthreads per step.
//   * the input projection is consumed as one float4 per cell (gate-interleaved layout): for layer 0 it is gathered from the
//     per-group table P[g][dir][token] (K1, the V x 4H lookup-table identity) -- the CTA's token ids are staged
//     once in smem as uint16 in scan order; for layers >= 1 it is the dense xproj tensor written by the tensor-core GEMM.
//     It is prefetched kD steps ahead with cp.async into per-thread smem slots (a register prefetch ring serialises on the
//     counting scoreboard: profiles/r1_lstm_ncu_full_summary.txt).  The step loop is branch-free: pointers advance by a
//     constant stride, columns beyond the batch are clamped for loads and masked for stores.
#include "kernels.h"

namespace ib200 {

namespace {

constexpr int kD = 4;  // async prefetch depth in steps (power of two)

template <int H, bool SPLIT>
struct FwdSmem {
  static constexpr int NT = H * 4, NPART = SPLIT ? 2 : 1;
  float4 xring[kD][2][NT];                 // per-thread slots of the input projection
  __nv_bfloat16 hs[2][NPART][H][kBC];      // h_{t-1}: [buffer][hi/lo][unit][sequence]
  // followed by uint16 toks[T + kD][kBC] (layer 0)
};

// HALF: 4 sequences per CTA on the EVEN mma columns (odd columns stay zero): one cell per thread instead of two.  Used when the
// launch would otherwise fill at most half of the SMs (e.g. the single live top-layer chain): the MMA work per CTA is unchanged
// but every other per-step cost (activations, loads, stores) halves and twice as many SMs work.
template <int H, bool SPLIT, bool FAST_ACT, bool LAYER0, bool TRAIN, bool HALF>
__global__ void __launch_bounds__(H * 4, HALF ? 2 : 1) lstm_fwd_kernel(const LstmFwdArgs p) {
  constexpr int NT = H * 4, KT = H / 16, NPART = SPLIT ? 2 : 1;
  using Smem = FwdSmem<H, SPLIT>;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tig = lane & 3;
  const int g = blockIdx.y, dir = p.dir0 + (int)blockIdx.z;
  const int T = p.lens[p.G + g];  // T_eff of this group
  if (T <= 0) return;
  constexpr int SEQ = HALF ? kBC / 2 : kBC;  // sequences per CTA
  const int b0 = blockIdx.x * SEQ;
  const int nvalid = min(SEQ, p.B - b0);
  const int nbase = g * p.B + b0;  // first global sequence index of this CTA
  const int Tmax = p.Tmax;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  uint16_t* toks = reinterpret_cast<uint16_t*>(smem_raw + sizeof(Smem));

  // ---- A fragments: W_hh rows of this warp's 8 units, masked per group for (layer 0, forward) -----------------------------
  const int u = warp * 8 + gq;  // the unit this thread owns
  uint32_t Ahi[2][KT][4], Alo[2][KT][4];
  {
    const float* __restrict__ W = (dir ? p.whh[1] : p.whh[0]);
    const float* __restrict__ M = (dir == 0 && p.whh_mask != nullptr) ? p.whh_mask + (size_t)g * 4 * H * H : nullptr;
#pragma unroll
    for (int tile = 0; tile < 2; ++tile) {
      const int r0 = (2 * tile) * H + u;      // fragment rows gq     : gate i (tile 0) / g (tile 1)
      const int r1 = (2 * tile + 1) * H + u;  // fragment rows gq + 8 : gate f (tile 0) / o (tile 1)
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        const int k0 = kt * 16 + 2 * tig;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int idx = ((j & 1) ? r1 : r0) * H + k0 + ((j & 2) ? 8 : 0);
          float w0 = W[idx], w1 = W[idx + 1];
          if (M != nullptr) {
            w0 *= M[idx];
            w1 *= M[idx + 1];
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

  // ---- init h = 0 (both buffers); stage this CTA's token ids in scan order (layer 0) ------------------------------------------
  for (int i = tid; i < 2 * NPART * H * kBC / 2; i += NT) reinterpret_cast<uint32_t*>(&sm.hs[0][0][0][0])[i] = 0u;
  if constexpr (LAYER0) {
    for (int i = tid; i < (T + kD) * SEQ; i += NT) {
      const int n = i / (T + kD), s = i % (T + kD);  // consecutive threads read consecutive time steps (coalesced)
      int v = 0;
      if (s < T && n < nvalid) v = p.tok[(size_t)(nbase + n) * Tmax + (dir ? (T - 1 - s) : s)];
      toks[s * kBC + (HALF ? 2 * n : n)] = (uint16_t)v;
    }
  }
  __syncthreads();

  const int n0 = 2 * tig, n1 = 2 * tig + 1;  // the two mma columns this thread owns (HALF: only n0 carries a sequence)
  const int q0 = HALF ? tig : n0, q1 = HALF ? tig : n1;  // sequence index within the CTA
  const bool v0 = q0 < nvalid, v1 = !HALF && q1 < nvalid;
  // columns beyond the batch read a valid sequence (clamped) and never store
  const int rb0 = (nbase + min(q0, nvalid - 1)) * Tmax, rb1 = (nbase + min(q1, nvalid - 1)) * Tmax;
  const int t_first = dir ? T - 1 : 0, dt = dir ? -1 : 1;
  const float4* __restrict__ xsrc =
      LAYER0 ? reinterpret_cast<const float4*>(p.table) + (size_t)(g * 2 + dir) * p.V * H + u
             : reinterpret_cast<const float4*>((dir ? p.xproj[1] : p.xproj[0])) + u;
  // layer >= 1: running source pointers of the prefetch (kD steps ahead of the compute), advanced by one row per step
  const float4* xp0 = xsrc + (size_t)(rb0 + t_first) * H;
  const float4* xp1 = xsrc + (size_t)(rb1 + t_first) * H;
  const ptrdiff_t xstride = (ptrdiff_t)dt * H;
  float4* slot = &sm.xring[0][0][tid];
  constexpr int kStage = 2 * NT;  // float4 per ring stage

  // one cp.async group per step; slot (s % kD) of this thread is re-armed right after it has been read
  auto issue = [&](int s) {
    float4* dst = slot + (s & (kD - 1)) * kStage;
    if constexpr (LAYER0) {
      const uint32_t tw = *reinterpret_cast<const uint32_t*>(toks + s * kBC + n0);  // tokens of (n0, n1); rows >= T hold 0
      cp_async16(dst, xsrc + (size_t)(tw & 0xffffu) * H, true);
      if constexpr (!HALF) cp_async16(dst + NT, xsrc + (size_t)(tw >> 16) * H, true);
    } else {
      const bool in = s < T;
      cp_async16(dst, xp0, in);
      if constexpr (!HALF) cp_async16(dst + NT, xp1, in);
      if (s + 1 < T) {
        xp0 += xstride;
        if constexpr (!HALF) xp1 += xstride;
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < kD; ++s) issue(s);

  float c0 = 0.f, c1 = 0.f, h0 = 0.f, h1 = 0.f;
  // running output pointers at time t(s)
  float4* g40 = nullptr;
  float4* g41 = nullptr;
  float* cs0 = nullptr;
  float* cs1 = nullptr;
  if constexpr (TRAIN) {
    float4* G4 = reinterpret_cast<float4*>((dir ? p.gates[1] : p.gates[0]));
    float* Cst = (dir ? p.cstate[1] : p.cstate[0]);
    g40 = G4 + (size_t)(rb0 + t_first) * H + u;
    g41 = G4 + (size_t)(rb1 + t_first) * H + u;
    cs0 = Cst + (size_t)(rb0 + t_first) * H + u;
    cs1 = Cst + (size_t)(rb1 + t_first) * H + u;
  }
  const bool has_y = p.y != nullptr, planes = p.planes != 0;
  // fp32 layout: float at column dir*H+u.  planes layout: the same row bytes hold bf16 [hi | lo]; pointer kept in float units of
  // the ROW START and the element is addressed as bf16 inside the row.
  float* y0 = has_y ? p.y + (size_t)(rb0 + t_first) * p.y_stride + (planes ? 0 : dir * H + u) : nullptr;
  float* y1 = has_y ? p.y + (size_t)(rb1 + t_first) * p.y_stride + (planes ? 0 : dir * H + u) : nullptr;
  const int ycol = dir * H + u;
  const ptrdiff_t gstride = (ptrdiff_t)dt * H, ystride = (ptrdiff_t)dt * p.y_stride;

  const __nv_bfloat16* hrow = &sm.hs[0][0][lane % H][0];  // ldmatrix row address of this lane (k-pair block 0, buffer 0)
  __nv_bfloat16* hput = &sm.hs[0][0][u][n0];
  constexpr int kBufElems = NPART * H * kBC, kPartElems = H * kBC;

  const int dbg = p.dbg;
  for (int s = 0; s < T; ++s) {
    float4 x0 = make_float4(0.1f, 0.2f, 0.3f, 0.4f), x1 = x0;
    if (!(dbg & 8)) {
      cp_async_wait<kD - 1>();  // this thread's copies for step s have landed
      const float4* cur = slot + (s & (kD - 1)) * kStage;
      x0 = cur[0];
      if constexpr (!HALF) x1 = cur[NT];
      issue(s + kD);
    }

    // accumulators start from the input projection: [0]=(row gq, col n0) [1]=(row gq, col n1) [2]=(row gq+8, n0) [3]=(row gq+8, n1)
    // tile 0 rows: gq -> i_u, gq+8 -> f_u ; tile 1 rows: gq -> g_u, gq+8 -> o_u
    float acc[2][4] = {{x0.x, x1.x, x0.y, x1.y}, {x0.z, x1.z, x0.w, x1.w}};
    float ac1[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    float ac2[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    const int buf = s & 1;
    if (!(dbg & 1))
#pragma unroll
    for (int kp = 0; kp < (KT + 1) / 2; ++kp) {
      uint32_t bh[4], bl[4];
      ldmatrix_x4_trans(bh, hrow + buf * kBufElems + kp * 32 * kBC);
      if constexpr (SPLIT) ldmatrix_x4_trans(bl, hrow + buf * kBufElems + kPartElems + kp * 32 * kBC);
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int kt = kp * 2 + kk;
        if (kt < KT) {
#pragma unroll
          for (int tile = 0; tile < 2; ++tile) {
            mma_bf16(acc[tile], Ahi[tile][kt], bh[2 * kk], bh[2 * kk + 1]);
            if constexpr (SPLIT) {
              mma_bf16(ac1[tile], Ahi[tile][kt], bl[2 * kk], bl[2 * kk + 1]);
              mma_bf16(ac2[tile], Alo[tile][kt], bh[2 * kk], bh[2 * kk + 1]);
            }
          }
        }
      }
    }
    float ai0 = acc[0][0], af0 = acc[0][2], ag0 = acc[1][0], ao0 = acc[1][2];
    float ai1 = acc[0][1], af1 = acc[0][3], ag1 = acc[1][1], ao1 = acc[1][3];
    if constexpr (SPLIT) {
      ai0 += ac1[0][0] + ac2[0][0]; af0 += ac1[0][2] + ac2[0][2]; ag0 += ac1[1][0] + ac2[1][0]; ao0 += ac1[1][2] + ac2[1][2];
      ai1 += ac1[0][1] + ac2[0][1]; af1 += ac1[0][3] + ac2[0][3]; ag1 += ac1[1][1] + ac2[1][1]; ao1 += ac1[1][3] + ac2[1][3];
    }

    float i0, f0, gg0, o0, i1, f1, gg1, o1;
    if (!(dbg & 2)) {
      i0 = sigmoid_f<FAST_ACT>(ai0), f0 = sigmoid_f<FAST_ACT>(af0), gg0 = tanh_f<FAST_ACT>(ag0), o0 = sigmoid_f<FAST_ACT>(ao0);
      c0 = fmaf(f0, c0, i0 * gg0);
      h0 = o0 * tanh_f<FAST_ACT>(c0);
      if constexpr (!HALF) {
        i1 = sigmoid_f<FAST_ACT>(ai1), f1 = sigmoid_f<FAST_ACT>(af1), gg1 = tanh_f<FAST_ACT>(ag1), o1 = sigmoid_f<FAST_ACT>(ao1);
        c1 = fmaf(f1, c1, i1 * gg1);
        h1 = o1 * tanh_f<FAST_ACT>(c1);
      } else {
        i1 = f1 = gg1 = o1 = 0.f;
      }
    } else {
      i0 = ai0 * 0.5f, f0 = af0 * 0.5f, gg0 = ag0 * 0.5f, o0 = ao0 * 0.5f, i1 = ai1 * 0.5f, f1 = af1 * 0.5f, gg1 = ag1 * 0.5f, o1 = ao1 * 0.5f;
      c0 = fmaf(f0, c0, i0 * gg0) * 0.1f;
      c1 = fmaf(f1, c1, i1 * gg1) * 0.1f;
      h0 = o0 * c0;
      h1 = o1 * c1;
    }

    // publish h_t for the next step: one packed word (sequences n0,n1 of unit u) per part
    {
      __nv_bfloat16* dst = hput + (buf ^ 1) * kBufElems;
      const __nv_bfloat162 hh = __floats2bfloat162_rn(h0, h1);
      *reinterpret_cast<__nv_bfloat162*>(dst) = hh;
      if constexpr (SPLIT) {
        const float2 hf = __bfloat1622float2(hh);
        *reinterpret_cast<__nv_bfloat162*>(dst + kPartElems) = __floats2bfloat162_rn(h0 - hf.x, h1 - hf.y);
      }
    }
    // stream out what later stages need (placement relative to the barrier makes no measurable difference: ablation in DESIGN.md)
    if (has_y && !(dbg & 4)) {
      if (planes) {
        // bf16 hi/lo of my two values as [hi | lo << 16] words; one shuffle with the neighbouring unit (lane ^ 4) lets every
        // thread store a 2-unit bf16x2 word per plane (even gq: units (u,u+1) of column n0; odd gq: units (u-1,u) of column n1)
        const __nv_bfloat162 a = __floats2bfloat162_rn(h0, h1);
        const float2 af = __bfloat1622float2(a);
        const __nv_bfloat162 l = __floats2bfloat162_rn(h0 - af.x, h1 - af.y);
        const uint32_t ab = *reinterpret_cast<const uint32_t*>(&a), lb = *reinterpret_cast<const uint32_t*>(&l);
        const uint32_t w0 = (ab & 0xffffu) | (lb << 16), w1 = (ab >> 16) | (lb & 0xffff0000u);  // value n0 / value n1
        const bool odd = gq & 1;
        const uint32_t recv = __shfl_xor_sync(0xffffffffu, (odd || HALF) ? w0 : w1, 4);
        const uint32_t mine = (odd && !HALF) ? w1 : w0;
        const uint32_t first = odd ? recv : mine, second = odd ? mine : recv;
        const uint32_t hiw = (first & 0xffffu) | (second << 16), low = (first >> 16) | (second & 0xffff0000u);
        const bool st_ok = HALF ? (!odd && v0) : (odd ? v1 : v0);
        if (st_ok) {
          uint32_t* dst = reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>((odd && !HALF) ? y1 : y0) + (ycol & ~1));
          dst[0] = hiw;
          if constexpr (SPLIT) dst[p.y_stride / 2] = low;
        }
      } else {
        if (v0) *y0 = h0;
        if (v1) *y1 = h1;
      }
      y0 += ystride;
      y1 += ystride;
    }
    if (TRAIN && !(dbg & 4)) {
      if (v0) {
        *g40 = make_float4(i0, f0, gg0, o0);
        *cs0 = c0;
      }
      if (v1) {
        *g41 = make_float4(i1, f1, gg1, o1);
        *cs1 = c1;
      }
      g40 += gstride;
      g41 += gstride;
      cs0 += gstride;
      cs1 += gstride;
    }
    if (!(dbg & 16)) __syncthreads();
  }
  cp_async_wait<0>();

  // planes mode: the weight-gradient GEMM reads 64-row TMA boxes (and the row after the last one for the shifted operand), so the
  // rows [T, tail_end) of this CTA's sequences / this direction's columns must be finite zeros, not uninitialised memory
  if (has_y && planes) {
    const int tail_end = min(Tmax, ((T + 63) / 64) * 64 + 1), ntail = tail_end - T;  // +1: the row a shifted box touches
    const int chunks = H * 2 / 16;  // 16-byte chunks per plane half-row of this direction
    for (int i = tid; i < nvalid * ntail * chunks * 2; i += NT) {
      const int c = i % chunks, pl = (i / chunks) & 1, r = (i / (2 * chunks)) % ntail, q = i / (2 * chunks * ntail);
      unsigned char* row = reinterpret_cast<unsigned char*>(p.y + ((size_t)(nbase + q) * Tmax + T + r) * p.y_stride);
      *reinterpret_cast<uint4*>(row + pl * p.y_stride * 2 + dir * H * 2 + c * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }

  if (p.hn != nullptr) {
    const size_t N = (size_t)p.G * p.B;
    if (v0) p.hn[((size_t)dir * N + nbase + q0) * H + u] = h0;
    if (v1) p.hn[((size_t)dir * N + nbase + q1) * H + u] = h1;
  }
}

template <int H, bool SPLIT, bool FAST, bool L0, bool TR, bool HALF>
cudaError_t launch_kh(const LstmFwdArgs& a, cudaStream_t st) {
  constexpr int SEQ = HALF ? kBC / 2 : kBC;
  dim3 grid((a.B + SEQ - 1) / SEQ, a.G, a.ndir), block(H * 4);
  size_t smem = sizeof(FwdSmem<H, SPLIT>);
  if (L0) smem += (size_t)(a.Tmax + kD) * kBC * sizeof(uint16_t);
  if (smem > 220 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(lstm_fwd_kernel<H, SPLIT, FAST, L0, TR, HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  lstm_fwd_kernel<H, SPLIT, FAST, L0, TR, HALF><<<grid, block, smem, st>>>(a);
  return cudaGetLastError();
}
template <int H, bool SPLIT, bool FAST, bool L0, bool TR>
cudaError_t launch_k(const LstmFwdArgs& a, cudaStream_t st) {
  // one cell per thread (4 sequences per CTA) whenever the doubled CTA count is still co-resident: HALF kernels are capped at
  // 128 registers so two of them share an SM and interleave their MMA / MUFU phases
  const int full_ctas = ((a.B + kBC - 1) / kBC) * a.G * a.ndir;
  const bool half = full_ctas <= ((a.dbg & 128) ? 74 : 148) && !(a.dbg & 64);
  return half ? launch_kh<H, SPLIT, FAST, L0, TR, true>(a, st) : launch_kh<H, SPLIT, FAST, L0, TR, false>(a, st);
}


// ------------------------------------------------------------------------------------------------------------------------------
// 16-warp variant: ONE cell per thread.  Each warp owns one 16-row tile = (i,f,g,o) x 4 units; lanes 0-15 hold the i/g rows,
// lanes 16-31 the f/o rows of the accumulator fragment, and one shfl_xor(16) pair gives every thread the four gates of its cell
// (lower half-warp: sequence 2*tig, upper half-warp: sequence 2*tig+1).  Twice the warps per scheduler of the 8-warp kernel:
// the MMA, MUFU and LSU phases of different warps overlap instead of running back to back.
// ------------------------------------------------------------------------------------------------------------------------------
template <int H, bool SPLIT>
struct Fwd16Smem {
  static constexpr int NT = H * 8, NPART = SPLIT ? 2 : 1;
  float4 xring[kD][NT];
  __nv_bfloat16 hs[2][NPART][H][kBC];
};

template <int H, bool SPLIT, bool FAST_ACT, bool LAYER0, bool TRAIN>
__global__ void __launch_bounds__(H * 8, 1) lstm_fwd16_kernel(const LstmFwdArgs p) {
  constexpr int NT = H * 8, KT = H / 16, NPART = SPLIT ? 2 : 1;
  using Smem = Fwd16Smem<H, SPLIT>;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tig = lane & 3;
  const int g = blockIdx.y, dir = p.dir0 + (int)blockIdx.z;
  const int T = p.lens[p.G + g];
  if (T <= 0) return;
  const int b0 = blockIdx.x * kBC;
  const int nvalid = min(kBC, p.B - b0);
  const int nbase = g * p.B + b0;
  const int Tmax = p.Tmax;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  uint16_t* toks = reinterpret_cast<uint16_t*>(smem_raw + sizeof(Smem));

  const bool upper = gq >= 4;
  const int u = warp * 4 + (gq & 3);  // the unit of this thread's cell
  const int n = 2 * tig + (upper ? 1 : 0);  // the sequence (column) of this thread's cell
  uint32_t Ahi[KT][4], Alo[KT][4];
  {
    const float* __restrict__ W = (dir ? p.whh[1] : p.whh[0]);
    const float* __restrict__ M = (dir == 0 && p.whh_mask != nullptr) ? p.whh_mask + (size_t)g * 4 * H * H : nullptr;
    const int r0 = (gq >> 2) * H + u;        // fragment row gq     : gate i (lanes 0-15) / f (lanes 16-31)
    const int r1 = (2 + (gq >> 2)) * H + u;  // fragment row gq + 8 : gate g / o
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      const int k0 = kt * 16 + 2 * tig;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = ((j & 1) ? r1 : r0) * H + k0 + ((j & 2) ? 8 : 0);
        float w0 = W[idx], w1 = W[idx + 1];
        if (M != nullptr) {
          w0 *= M[idx];
          w1 *= M[idx + 1];
        }
        if constexpr (SPLIT) {
          split_bf16(w0, w1, Ahi[kt][j], Alo[kt][j]);
        } else {
          Ahi[kt][j] = pack_bf16(w0, w1);
          Alo[kt][j] = 0u;
        }
      }
    }
  }

  for (int i = tid; i < 2 * NPART * H * kBC / 2; i += NT) reinterpret_cast<uint32_t*>(&sm.hs[0][0][0][0])[i] = 0u;
  if constexpr (LAYER0) {
    for (int i = tid; i < (T + kD) * kBC; i += NT) {
      const int nn = i / (T + kD), s = i % (T + kD);
      int v = 0;
      if (s < T && nn < nvalid) v = p.tok[(size_t)(nbase + nn) * Tmax + (dir ? (T - 1 - s) : s)];
      toks[s * kBC + nn] = (uint16_t)v;
    }
  }
  __syncthreads();

  const bool valid = n < nvalid;
  const int rb = (nbase + min(n, nvalid - 1)) * Tmax;
  const int t_first = dir ? T - 1 : 0, dt = dir ? -1 : 1;
  const float4* __restrict__ xsrc =
      LAYER0 ? reinterpret_cast<const float4*>(p.table) + (size_t)(g * 2 + dir) * p.V * H + u
             : reinterpret_cast<const float4*>((dir ? p.xproj[1] : p.xproj[0])) + u;
  const float4* xp = xsrc + (size_t)(rb + t_first) * H;
  const ptrdiff_t xstride = (ptrdiff_t)dt * H;
  float4* slot = &sm.xring[0][tid];

  auto issue = [&](int s) {
    float4* dst = slot + (s & (kD - 1)) * NT;
    if constexpr (LAYER0) {
      cp_async16(dst, xsrc + (size_t)toks[s * kBC + n] * H, true);
    } else {
      cp_async16(dst, xp, s < T);
      if (s + 1 < T) xp += xstride;
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < kD; ++s) issue(s);

  float c = 0.f, h = 0.f;
  float4* g4 = nullptr;
  float* cs = nullptr;
  if constexpr (TRAIN) {
    g4 = reinterpret_cast<float4*>((dir ? p.gates[1] : p.gates[0])) + (size_t)(rb + t_first) * H + u;
    cs = (dir ? p.cstate[1] : p.cstate[0]) + (size_t)(rb + t_first) * H + u;
  }
  const bool has_y = p.y != nullptr;
  float* yp = has_y ? p.y + (size_t)(rb + t_first) * p.y_stride + dir * H + u : nullptr;
  const ptrdiff_t gstride = (ptrdiff_t)dt * H, ystride = (ptrdiff_t)dt * p.y_stride;
  const __nv_bfloat16* hrow = &sm.hs[0][0][lane % H][0];
  __nv_bfloat16* hput = &sm.hs[0][0][u][n];
  constexpr int kBufElems = NPART * H * kBC, kPartElems = H * kBC;

  for (int s = 0; s < T; ++s) {
    cp_async_wait<kD - 1>();
    const float4 x = slot[(s & (kD - 1)) * NT];
    issue(s + kD);

    float acc[4] = {0.f, 0.f, 0.f, 0.f}, ac1[4] = {0.f, 0.f, 0.f, 0.f}, ac2[4] = {0.f, 0.f, 0.f, 0.f};
    const int buf = s & 1;
#pragma unroll
    for (int kp = 0; kp < (KT + 1) / 2; ++kp) {
      uint32_t bh[4], bl[4];
      ldmatrix_x4_trans(bh, hrow + buf * kBufElems + kp * 32 * kBC);
      if constexpr (SPLIT) ldmatrix_x4_trans(bl, hrow + buf * kBufElems + kPartElems + kp * 32 * kBC);
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int kt = kp * 2 + kk;
        if (kt < KT) {
          mma_bf16(acc, Ahi[kt], bh[2 * kk], bh[2 * kk + 1]);
          if constexpr (SPLIT) {
            mma_bf16(ac1, Ahi[kt], bl[2 * kk], bl[2 * kk + 1]);
            mma_bf16(ac2, Alo[kt], bh[2 * kk], bh[2 * kk + 1]);
          }
        }
      }
    }
    float v0 = acc[0], v1 = acc[1], v2 = acc[2], v3 = acc[3];
    if constexpr (SPLIT) {
      v0 += ac1[0] + ac2[0];
      v1 += ac1[1] + ac2[1];
      v2 += ac1[2] + ac2[2];
      v3 += ac1[3] + ac2[3];
    }
    // lanes 0-15: v0,v1 = i(2tig),i(2tig+1)  v2,v3 = g(..)   lanes 16-31: v0,v1 = f(..)  v2,v3 = o(..)
    const float r0 = __shfl_xor_sync(0xffffffffu, upper ? v0 : v1, 16);
    const float r1 = __shfl_xor_sync(0xffffffffu, upper ? v2 : v3, 16);
    const float ai = (upper ? r0 : v0) + x.x, af = (upper ? v1 : r0) + x.y;
    const float ag = (upper ? r1 : v2) + x.z, ao = (upper ? v3 : r1) + x.w;
    const float gi = sigmoid_f<FAST_ACT>(ai), gf = sigmoid_f<FAST_ACT>(af), gg = tanh_f<FAST_ACT>(ag), go = sigmoid_f<FAST_ACT>(ao);
    c = fmaf(gf, c, gi * gg);
    h = go * tanh_f<FAST_ACT>(c);
    {
      __nv_bfloat16* dst = hput + (buf ^ 1) * kBufElems;
      const __nv_bfloat16 hh = __float2bfloat16_rn(h);
      *dst = hh;
      if constexpr (SPLIT) dst[kPartElems] = __float2bfloat16_rn(h - __bfloat162float(hh));
    }
    if (has_y) {
      if (valid) *yp = h;
      yp += ystride;
    }
    if constexpr (TRAIN) {
      if (valid) {
        *g4 = make_float4(gi, gf, gg, go);
        *cs = c;
      }
      g4 += gstride;
      cs += gstride;
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  if (p.hn != nullptr && valid) p.hn[((size_t)dir * p.G * p.B + nbase + n) * H + u] = h;
}

template <int H, bool SPLIT, bool FAST, bool L0, bool TR>
cudaError_t launch_k16(const LstmFwdArgs& a, cudaStream_t st) {
  dim3 grid((a.B + kBC - 1) / kBC, a.G, a.ndir), block(H * 8);
  size_t smem = sizeof(Fwd16Smem<H, SPLIT>);
  if (L0) smem += (size_t)(a.Tmax + kD) * kBC * sizeof(uint16_t);
  if (smem > 220 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(lstm_fwd16_kernel<H, SPLIT, FAST, L0, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  lstm_fwd16_kernel<H, SPLIT, FAST, L0, TR><<<grid, block, smem, st>>>(a);
  return cudaGetLastError();
}

template <int H, bool SPLIT, bool FAST>
cudaError_t launch_h(const LstmFwdArgs& a, cudaStream_t st) {
  const bool l0 = a.tok != nullptr, tr = a.gates[a.dir0] != nullptr;
  if (l0 && a.V > 65536) return cudaErrorInvalidValue;  // token ids are staged as uint16
  if (a.dbg & 32) {
    if (l0 && tr) return launch_k16<H, SPLIT, FAST, true, true>(a, st);
    if (l0) return launch_k16<H, SPLIT, FAST, true, false>(a, st);
    if (tr) return launch_k16<H, SPLIT, FAST, false, true>(a, st);
    return launch_k16<H, SPLIT, FAST, false, false>(a, st);
  }
  if (l0 && tr) return launch_k<H, SPLIT, FAST, true, true>(a, st);
  if (l0) return launch_k<H, SPLIT, FAST, true, false>(a, st);
  if (tr) return launch_k<H, SPLIT, FAST, false, true>(a, st);
  return launch_k<H, SPLIT, FAST, false, false>(a, st);
}

}  // namespace

cudaError_t launch_lstm_fwd(const LstmFwdArgs& a, int H, int precision, cudaStream_t st) {
  if (H == 64) return precision == 0 ? launch_h<64, true, false>(a, st) : launch_h<64, false, true>(a, st);
  if (H == 32) return precision == 0 ? launch_h<32, true, false>(a, st) : launch_h<32, false, true>(a, st);
  return cudaErrorInvalidValue;
}

}  // namespace ib200
