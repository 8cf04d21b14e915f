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
// Saved gates, c and dy are prefetched kD steps ahead with cp.async into per-thread smem slots.
// Two __syncthreads per step.  fp32 mode: bf16 hi/lo split, 3 MMAs per product, every product on short accumulator chains.
#include "kernels.h"

namespace ib200 {
static bool bwd_half(const LstmBwdArgs& a, bool split);
namespace {

constexpr int kD = 8;  // async prefetch depth in steps (power of two)

__device__ __forceinline__ void stg_v4_if(void* ptr, const float4& v, bool pred) {
  asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %5, 0;\n @q st.global.v4.f32 [%0], {%1,%2,%3,%4};\n}\n" ::"l"(ptr), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w), "r"((int)pred)
               : "memory");
}
__device__ __forceinline__ void stg_v2_if(void* ptr, uint32_t a, uint32_t b, bool pred) {
  asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %3, 0;\n @q st.global.v2.u32 [%0], {%1,%2};\n}\n" ::"l"(ptr), "r"(a), "r"(b),
               "r"((int)pred)
               : "memory");
}

template <int H>
struct BwdSmem {
  static constexpr int NT = H * 4, NW = H / 8, KTT = H / 4;
  float4 rg[kD][2][NT];            // saved (i,f,g,o) of my two cells
  float rc[kD][2][NT];             // c at the scan predecessor (0 at the chain start)
  float rdy[kD][2][NT];            // upstream gradient of h_t
  uint32_t dafrag[2][KTT][32][2];  // da as mma B fragments: [hi/lo][k tile][lane][2 words]
  float2 xch[NW][32];              // partial dh handed to the partner warp
  int shared_sm;                   // phase 1 (two-phase rebalancing, common.cuh): 1 when a second CTA is resident on this SM
};

// HALF: 4 sequences per CTA, one cell per thread (see lstm_fwd.cu): bf16 mode on the even mma columns; fp32 mode ("HL") with
// da_hi on columns 0-3 and da_lo on columns 4-7, so 2 MMAs per product (A_hi, A_lo) instead of 3 and one shfl_xor(2) to add the
// hi-column and lo-column sums of a sequence.
// PHASED: takes part in the two-phase rebalancing of common.cuh (window of steps, state save / restore); plain variants carry none of it.
template <int H, bool SPLIT, bool FAST_ACT, bool HAS_DY, bool HALF, bool PHASED>
__global__ void __launch_bounds__(H * 4, HALF ? 2 : 1) lstm_bwd_kernel(const LstmBwdArgs p) {
  static_assert(!PHASED || HALF, "PHASED is a HALF-mode variant");
  constexpr int NT = H * 4, MT = H / 16, KTT = H / 4, KTH = KTT / 2;
  constexpr bool HL = HALF && SPLIT;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tig = lane & 3;
  const int mt = warp % MT, kh = warp / MT;
  int bx = blockIdx.x, g = blockIdx.y, dz = blockIdx.z;
  if constexpr (PHASED) {
    if (!phase_cta(p.ph, p.G, bx, g, dz)) return;
  }
  const int dir = p.dir0 + dz;
  const int T = p.lens[p.G + g];
  if (T <= 0) return;
  // scan steps [s_begin, s_end) of the chain run in this launch (two-phase rebalancing, common.cuh)
  const int split = PHASED ? phase_split(p.ph, T) : T;
  const int s_begin = (PHASED && p.ph.phase == 2) ? split : 0;
  int s_end = T;
  constexpr int SEQ = HALF ? kBC / 2 : kBC, NC = HALF ? 1 : 2;  // sequences per CTA, cells per thread
  const int b0 = bx * SEQ;
  const int nvalid = min(SEQ, p.B - b0);
  const int nbase = g * p.B + b0;
  const int Tmax = p.Tmax;
  const size_t N = (size_t)p.G * p.B;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem<H>& sm = *reinterpret_cast<BwdSmem<H>*>(smem_raw);

  // ---- A fragments: W_hh^T, K ordered as (unit, gate) with the fragment positions {2tig,2tig+1,2tig+8,2tig+9} = gates i,f,g,o ----
  uint32_t Ahi[KTH][4], Alo[KTH][4];
  {
    const float* __restrict__ W = (dir ? p.whh[1] : p.whh[0]);
    const float* __restrict__ M = (dir == 0 && p.whh_mask != nullptr) ? p.whh_mask + (size_t)g * 4 * H * H : nullptr;
    const int j0 = 16 * mt + gq;
#pragma unroll
    for (int ktl = 0; ktl < KTH; ++ktl) {
      const int uu = 4 * (kh * KTH + ktl) + tig;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int jj = j0 + ((r & 1) ? 8 : 0), q = (r & 2) ? 2 : 0;
        const int ia = (q * H + uu) * H + jj, ib = ((q + 1) * H + uu) * H + jj;
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
  const int n0 = 2 * tig, n1 = n0 + 1;  // mma columns of this thread's accumulators
  const bool low = tig < 2;
  // sequence index within the CTA.  HALF/bf16: sequence tig on the even column.  HL: low lanes own sequence 2*tig (their even
  // column carries its hi part), high lanes own sequence 2*(tig-2)+1 (their odd column carries its lo part).
  const int q0 = HL ? (low ? 2 * tig : 2 * (tig - 2) + 1) : (HALF ? tig : n0), q1 = HALF ? q0 : n1;
  const bool v0 = q0 < nvalid, v1 = !HALF && q1 < nvalid;
  // columns beyond the batch read a valid sequence (clamped) and never store
  const int rb0 = (nbase + min(q0, nvalid - 1)) * Tmax, rb1 = (nbase + min(q1, nvalid - 1)) * Tmax;
  // backward scan: s = 0..T-1 visits t = T-1..0 (forward chain) or t = 0..T-1 (reverse chain)
  const int t_first = dir ? 0 : T - 1, dt = dir ? 1 : -1;
  float4* __restrict__ G4 = reinterpret_cast<float4*>((dir ? p.gates[1] : p.gates[0]));
  const float* __restrict__ C = (dir ? p.cstate[1] : p.cstate[0]);
  const ptrdiff_t gstride = (ptrdiff_t)dt * H;
  // running pointers of the prefetch (kD steps ahead): gates at t(s), c of the scan predecessor = c at t(s+1), dy at t(s)
  const int tb = t_first + s_begin * dt;  // time index of the first step of this launch
  const float4* gp0 = G4 + (size_t)(rb0 + tb) * H + j;
  const float4* gp1 = G4 + (size_t)(rb1 + tb) * H + j;
  const float* cp0 = C + (size_t)(rb0 + tb) * H + j;  // advanced BEFORE use: first use is the step after
  const float* cp1 = C + (size_t)(rb1 + tb) * H + j;
  const float* dp0 = HAS_DY ? p.dy + (size_t)(rb0 + tb) * (2 * H) + dir * H + j : nullptr;
  const float* dp1 = HAS_DY ? p.dy + (size_t)(rb1 + tb) * (2 * H) + dir * H + j : nullptr;
  const ptrdiff_t dstride = (ptrdiff_t)dt * (2 * H);  // == p.dy_stride (checked by the launcher)

  auto issue = [&](int s) {
    const int st = s & (kD - 1);
    const bool in = s < T, has_prev = s + 1 < T;
    cp_async16(&sm.rg[st][0][tid], gp0, in);
    if constexpr (!HALF) cp_async16(&sm.rg[st][1][tid], gp1, in);
    if (has_prev) {
      cp0 += gstride;
      gp0 += gstride;
      if constexpr (!HALF) {
        cp1 += gstride;
        gp1 += gstride;
      }
    }
    cp_async4(&sm.rc[st][0][tid], cp0, has_prev);  // zero-filled at the chain start (c_{-1} = 0)
    if constexpr (!HALF) cp_async4(&sm.rc[st][1][tid], cp1, has_prev);
    if constexpr (HAS_DY) {
      cp_async4(&sm.rdy[st][0][tid], dp0, in);
      if constexpr (!HALF) cp_async4(&sm.rdy[st][1][tid], dp1, in);
      if (has_prev) {
        dp0 += dstride;
        if constexpr (!HALF) dp1 += dstride;
      }
    }
    cp_async_commit();
  };

  for (int i = tid; i < 2 * KTT * 32 * 2; i += NT) (&sm.dafrag[0][0][0][0])[i] = 0u;  // HALF: odd columns stay zero
  if constexpr (PHASED) {
    if (tid == 0) {
      sm.shared_sm = 0;
      if (p.ph.phase == 1) atomicAdd(p.ph.sm_load + sm_id(), 1);
    }
  }
  __syncthreads();
  float ccur[2], dc[2] = {0.f, 0.f}, dhrec[2] = {0.f, 0.f};
  float bsum[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};  // column sums of the dgates of my cells (bias gradient)
  const bool planes = p.planes != 0;
  ccur[0] = cp0[0];
  ccur[1] = cp1[0];
  // two-phase rebalancing: recurrent state of my cell(s), [dir slot][sequence][unit][dc, dh]
  auto state_slot = [&](int q) {
    return reinterpret_cast<float2*>(p.ph.state) + ((size_t)dz * N + nbase + min(q, nvalid - 1)) * H + j;
  };
  if (PHASED && s_begin > 0) {  // phase 2: resume
    float2* st0 = state_slot(q0);
    float2* st1 = state_slot(q1);
    const float2 a = *st0;
    dc[0] = a.x;
    dhrec[0] = a.y;
    if constexpr (!HALF) {
      const float2 b = *st1;
      dc[1] = b.x;
      dhrec[1] = b.y;
    }
  } else if (p.dhn != nullptr) {
    dhrec[0] = p.dhn[((size_t)dir * N + nbase + min(q0, nvalid - 1)) * H + j];
    dhrec[1] = p.dhn[((size_t)dir * N + nbase + min(q1, nvalid - 1)) * H + j];
  }
#pragma unroll
  for (int s = 0; s < kD; ++s) issue(s_begin + s);

  float4* gs0 = G4 + (size_t)(rb0 + tb) * H + j;  // da store pointers at t(s)
  float4* gs1 = G4 + (size_t)(rb1 + tb) * H + j;
  uint32_t* dst_hi0 = &sm.dafrag[0][j >> 2][(HL ? q0 : n0) * 4 + (j & 3)][0];
  constexpr int kFragPart = KTT * 32 * 2;  // words per hi/lo part
  const uint32_t* bsrc = &sm.dafrag[0][kh * KTH][lane][0];

  // The factors of step s that do not depend on the recurrent gradient are computed one step EARLY, inside the MMA phase of
  // step s-1 (ring read, tanh(c), gate-derivative products), so the dependent chain of a step is only
  //   dh = dhrec + dy ; dct = dh*Ac + dc ; da = (dct*Ai, dct*Af, dct*Ag, dh*Ao) ; dc = dct*gf
  // followed by the pack, STS, barrier and the MMAs.   Ao = tanh(c) o(1-o), Ac = o (1-tanh(c)^2), Ai = g i(1-i), Ag = i (1-g^2),
  // Af = c_prev f(1-f).
  struct Coef {
    float Ao, Ac, Ai, Ag, Af, gf, dy;
  };
  Coef K[2];
  auto load_coef = [&](int s) {
    const int st = s & (kD - 1);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float4 g = sm.rg[st][c][tid];
      const float cprev = sm.rc[st][c][tid];
      const float gi = g.x, gf = g.y, gg = g.z, go = g.w;
      const float tc = tanh_f<FAST_ACT>(ccur[c]);
      K[c].Ao = tc * (go * (1.0f - go));
      K[c].Ac = go * fmaf(-tc, tc, 1.0f);
      K[c].Ai = gg * (gi * (1.0f - gi));
      K[c].Ag = gi * fmaf(-gg, gg, 1.0f);
      K[c].Af = cprev * (gf * (1.0f - gf));
      K[c].gf = gf;
      K[c].dy = 0.f;
      if constexpr (HAS_DY) K[c].dy = sm.rdy[st][c][tid];
      ccur[c] = cprev;
    }
  };
  cp_async_wait<kD - 1>();
  load_coef(s_begin);
  issue(s_begin + kD);

  for (int s = s_begin; s < s_end; ++s) {
    if constexpr (PHASED) {
      if (p.ph.phase == 1) {  // (uniform; two compares per step)
        if (s == kPhaseCheck) {
          if (tid == 0) sm.shared_sm = *reinterpret_cast<volatile int*>(p.ph.sm_load + sm_id()) >= 2;  // visible after this step's barriers
        } else if (s == kPhaseCheck + 1) {
          if (sm.shared_sm) s_end = split;
        }
      }
    }
    float4 sv[2];
    uint4 pk[2];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float dh = dhrec[c] + K[c].dy;
      const float dct = fmaf(dh, K[c].Ac, dc[c]);
      const float da_i = dct * K[c].Ai, da_f = dct * K[c].Af, da_g = dct * K[c].Ag, da_o = dh * K[c].Ao;
      dc[c] = dct * K[c].gf;
      sv[c] = make_float4(da_i, da_f, da_g, da_o);
      if (c == 0 ? v0 : v1) {
        bsum[c][0] += da_i; bsum[c][1] += da_f; bsum[c][2] += da_g; bsum[c][3] += da_o;
      }
      // B-fragment slot of this cell: k tile j/4, lane (col*4 + j%4)
      uint32_t hi0, hi1, lo0 = 0u, lo1 = 0u;
      if constexpr (SPLIT) {
        split_bf16(da_i, da_f, hi0, lo0);
        split_bf16(da_g, da_o, hi1, lo1);
      } else {
        hi0 = pack_bf16(da_i, da_f);
        hi1 = pack_bf16(da_g, da_o);
      }
      *reinterpret_cast<uint2*>(dst_hi0 + c * 8) = make_uint2(hi0, hi1);
      if constexpr (HL) *reinterpret_cast<uint2*>(dst_hi0 + (kBC / 2) * 4 * 2) = make_uint2(lo0, lo1);  // column q0 + 4
      else if constexpr (SPLIT) *reinterpret_cast<uint2*>(dst_hi0 + c * 8 + kFragPart) = make_uint2(lo0, lo1);
      pk[c] = make_uint4(hi0, hi1, lo0, lo1);
    }
    __syncthreads();  // (A) all da of this step are in smem

    // the in-place da store (volatile asm: it keeps its place among the HMMAs below and issues in the gaps between them)
    auto store_da = [&]() {
      if (planes) {  // bf16 planes over the same 4H-float row: [hi: 4H bf16 | lo: 4H bf16], gate-interleaved columns 4j..4j+3
        uint2* r0p = reinterpret_cast<uint2*>(gs0 - j) + j;  // row start, then 8-byte slot j
        stg_v2_if(r0p, pk[0].x, pk[0].y, v0);
        if constexpr (SPLIT) stg_v2_if(r0p + H, pk[0].z, pk[0].w, v0);
        if constexpr (!HALF) {
          uint2* r1p = reinterpret_cast<uint2*>(gs1 - j) + j;
          stg_v2_if(r1p, pk[1].x, pk[1].y, v1);
          if constexpr (SPLIT) stg_v2_if(r1p + H, pk[1].z, pk[1].w, v1);
        }
      } else {
        stg_v4_if(gs0, sv[0], v0);
        if constexpr (!HALF) stg_v4_if(gs1, sv[1], v1);
      }
      gs0 += gstride;
      gs1 += gstride;
    };

    if (s + 1 < T) {
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      float ac1[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      float ac2[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int ktl = 0; ktl < KTH; ++ktl) {
        const uint2 bh = *reinterpret_cast<const uint2*>(bsrc + ktl * 64);
        if constexpr (HL) {  // two chains of depth KTH are enough to keep the tensor pipe busy: fewer adds afterwards
          mma_bf16(acc[0], Ahi[ktl], bh.x, bh.y);
          mma_bf16(ac2[0], Alo[ktl], bh.x, bh.y);
        } else {
          mma_bf16(acc[ktl & 1], Ahi[ktl], bh.x, bh.y);
        }
        if constexpr (SPLIT && !HL) {
          const uint2 bl = *reinterpret_cast<const uint2*>(bsrc + ktl * 64 + kFragPart);
          mma_bf16(ac1[ktl & 1], Ahi[ktl], bl.x, bl.y);
          mma_bf16(ac2[ktl & 1], Alo[ktl], bh.x, bh.y);
        }
        if (ktl == (KTH > 2 ? 1 : 0)) store_da();
        if (ktl == (KTH > 2 ? KTH / 2 : KTH - 1)) {  // next step's factors + the prefetch kD steps ahead
          cp_async_wait<kD - 1>();
          load_coef(s + 1);
          issue(s + 1 + kD);
        }
      }
      float r4[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if constexpr (HL) r4[r] = acc[0][r] + ac2[0][r];
        else if constexpr (SPLIT) r4[r] = (acc[0][r] + acc[1][r]) + ((ac1[0][r] + ac1[1][r]) + (ac2[0][r] + ac2[1][r]));
        else r4[r] = acc[0][r] + acc[1][r];
      }
      // r4: [0]=(unit 16mt+gq, col n0) [1]=(.., n1) [2]=(unit 16mt+gq+8, n0) [3]=(.., n1).  Keep the rows of my unit, hand
      // the other two to the partner warp (same mt, other K half), which owns that unit.
      const float2 mine = kh ? make_float2(r4[2], r4[3]) : make_float2(r4[0], r4[1]);
      sm.xch[warp][lane] = kh ? make_float2(r4[0], r4[1]) : make_float2(r4[2], r4[3]);
      __syncthreads();  // (B)
      const float2 other = sm.xch[kh ? warp - MT : warp + MT][lane];
      if constexpr (HL) {  // my column + the partner lane's column of the same sequence (hi part + lo part)
        const float ev = mine.x + other.x, od = mine.y + other.y;
        dhrec[0] = (low ? ev : od) + __shfl_xor_sync(0xffffffffu, low ? od : ev, 2);
      } else {
        dhrec[0] = mine.x + other.x;
        dhrec[1] = mine.y + other.y;
      }
    } else {
      store_da();
    }
  }
  cp_async_wait<0>();

  const bool stopped = PHASED && s_end < T;  // phase 1 on a shared SM: park the state, ask for a phase-2 CTA
  if (stopped) {
    if (v0) *state_slot(q0) = make_float2(dc[0], dhrec[0]);
    if (v1) *state_slot(q1) = make_float2(dc[1], dhrec[1]);
    if (tid == 0) {
      const int lin = bx + p.ph.grid_x * (g + p.G * dz);
      p.ph.resume_list[1 + atomicAdd(p.ph.resume_list, 1)] = lin;
    }
  }

  if (planes) {
    // (1) bias-gradient partials of this CTA: sum my cells over the 8 columns held by the 4 lanes of a quad, lane tig==0 writes
    if (p.bias_partial != nullptr) {
      float4 b = make_float4(bsum[0][0] + bsum[1][0], bsum[0][1] + bsum[1][1], bsum[0][2] + bsum[1][2], bsum[0][3] + bsum[1][3]);
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        b.x += __shfl_xor_sync(0xffffffffu, b.x, o);
        b.y += __shfl_xor_sync(0xffffffffu, b.y, o);
        b.z += __shfl_xor_sync(0xffffffffu, b.z, o);
        b.w += __shfl_xor_sync(0xffffffffu, b.w, o);
      }
      const int tiles_x = (PHASED && p.ph.phase == 2) ? p.ph.grid_x : (int)gridDim.x;
      const size_t ncta = (size_t)p.G * tiles_x, ci = (size_t)g * tiles_x + bx;
      if (tig == 0) {
        float4* slot = reinterpret_cast<float4*>(p.bias_partial + ((size_t)dz * ncta + ci) * 4 * H + 4 * j);
        if (s_begin > 0) {  // phase 2 adds to what the same logical CTA summed in phase 1
          const float4 o = *slot;
          b.x += o.x; b.y += o.y; b.z += o.z; b.w += o.w;
        }
        *slot = b;
      }
    }
    if (stopped) return;
    // (2) zero tail rows [T, tail_end) of my sequences: the TN GEMM reads whole 64-row TMA boxes
    const int tail_end = min(Tmax, ((T + 63) / 64) * 64 + 1), ntail = tail_end - T;  // +1: the row a shifted box touches
    const int chunks = 4 * H * 4 / 16;  // 16-byte chunks per row (both planes)
    for (int i = tid; i < nvalid * ntail * chunks; i += NT) {
      const int c = i % chunks, r = (i / chunks) % ntail, q = i / (chunks * ntail);
      reinterpret_cast<uint4*>(G4 + ((size_t)(nbase + q) * Tmax + T + r) * H)[c] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

template <int H, bool SPLIT, bool FAST, bool HAS_DY, bool HALF, bool PHASED = false>
cudaError_t launch_kh(const LstmBwdArgs& a, cudaStream_t st) {
  constexpr int SEQ = HALF ? kBC / 2 : kBC;
  dim3 grid((a.B + SEQ - 1) / SEQ, a.G, a.ndir);
  if (PHASED && a.ph.phase == 2) grid = dim3(grid.x * grid.y * grid.z, 1, 1);  // 1-D: CTA i resumes resume_list[i], the rest exit at once
  const size_t smem = sizeof(BwdSmem<H>);
  if (a.dy != nullptr && a.dy_stride != 2 * H) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(lstm_bwd_kernel<H, SPLIT, FAST, HAS_DY, HALF, PHASED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  lstm_bwd_kernel<H, SPLIT, FAST, HAS_DY, HALF, PHASED><<<grid, H * 4, smem, st>>>(a);
  return cudaGetLastError();
}
template <int H, bool SPLIT, bool FAST, bool HAS_DY>
cudaError_t launch_k(const LstmBwdArgs& a, cudaStream_t st) {
  // one cell per thread and two co-resident CTAs per SM whenever the full-width launch would leave SMs idle
  const bool half = bwd_half(a, SPLIT);
  LstmBwdArgs b = a;
  b.ph = PhaseArgs{};
  if (!half) return launch_kh<H, SPLIT, FAST, HAS_DY, false>(b, st);
  const int tiles = (a.B + kBC / 2 - 1) / (kBC / 2), half_ctas = tiles * a.G * a.ndir;
  if (a.ph.state != nullptr && half_ctas > 148 && half_ctas < 2 * 148 && a.Tmax >= 8 * kPhaseCheck && !(a.dbg & 2048)) {
    // between one and two CTAs per SM: two-phase rebalancing (common.cuh)
    b.ph = a.ph;
    b.ph.phase = 1;
    b.ph.grid_x = tiles;
    cudaError_t e = launch_kh<H, SPLIT, FAST, HAS_DY, true, true>(b, st);
    if (e != cudaSuccess) return e;
    b.ph.phase = 2;
    return launch_kh<H, SPLIT, FAST, HAS_DY, true, true>(b, st);
  }
  return launch_kh<H, SPLIT, FAST, HAS_DY, true>(b, st);
}
template <int H, bool SPLIT, bool FAST>
cudaError_t launch_b(const LstmBwdArgs& a, cudaStream_t st) {
  return a.dy != nullptr ? launch_k<H, SPLIT, FAST, true>(a, st) : launch_k<H, SPLIT, FAST, false>(a, st);
}

}  // namespace

static bool bwd_half(const LstmBwdArgs& a, bool split) {
  const int full_ctas = ((a.B + kBC - 1) / kBC) * a.G * a.ndir;
  (void)split;  // fp32 mode too since the HL column layout cut its MMA phase by a third (bench: 1.14 -> 0.98 ms on layer 0)
  return full_ctas <= ((a.dbg & 128) ? 74 : 148) && !(a.dbg & 64);
}
int lstm_bwd_cta_count(const LstmBwdArgs& a, int precision) {
  const int seq = bwd_half(a, precision == 0) ? kBC / 2 : kBC;
  return a.G * ((a.B + seq - 1) / seq);
}

cudaError_t launch_lstm_bwd(const LstmBwdArgs& a, int H, int precision, cudaStream_t st) {
  if (H == 64) return precision == 0 ? launch_b<64, true, false>(a, st) : launch_b<64, false, true>(a, st);
  if (H == 32) return precision == 0 ? launch_b<32, true, false>(a, st) : launch_b<32, false, true>(a, st);
  return cudaErrorInvalidValue;
}

}  // namespace ib200
