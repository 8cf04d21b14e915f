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

// predicated global stores as volatile asm: they keep their place BETWEEN the (volatile) mma instructions, so the scheduler issues
// them in the gaps the tensor pipe leaves between the HMMAs of one warp instead of after the last one
__device__ __forceinline__ void stg_v4_if(void* ptr, const float4& v, bool pred) {
  asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %5, 0;\n @q st.global.v4.f32 [%0], {%1,%2,%3,%4};\n}\n" ::"l"(ptr), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w), "r"((int)pred)
               : "memory");
}
__device__ __forceinline__ void stg_f32_if(void* ptr, float v, bool pred) {
  asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %2, 0;\n @q st.global.f32 [%0], %1;\n}\n" ::"l"(ptr), "f"(v), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void stg_bf16_if(void* ptr, __nv_bfloat16 v, bool pred) {
  asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %2, 0;\n @q st.global.u16 [%0], %1;\n}\n" ::"l"(ptr),
               "h"(*reinterpret_cast<const unsigned short*>(&v)), "r"((int)pred)
               : "memory");
}

// In-kernel ablation switches (tools/ablate_fwd.py) exist only in -DIB200_ABLATE builds: the production loop carries no flag tests.
#ifdef IB200_ABLATE
#define IB200_DBGBITS(p) ((p).dbg)
#else
#define IB200_DBGBITS(p) 0
#endif

template <int H, int NPART>
struct FwdSmem {
  static constexpr int NT = H * 4;
  float4 xring[kD][2][NT];                 // per-thread slots of the input projection
  __nv_bfloat16 hs[2][NPART][H][kBC];      // h_{t-1}: [buffer][part][unit][mma column]
  int shared_sm;                           // phase 1: 1 when a second CTA is resident on this SM
  // followed by uint16 toks[T + kD][kBC] (layer 0)
};

// HALF: 4 sequences per CTA, one cell per thread instead of two, two CTAs per SM.  Used when the launch would otherwise leave SMs
// idle (e.g. the single live top-layer chain): every per-step cost except the MMAs halves and twice as many SMs work.
//   * bf16 mode: the sequences sit on the EVEN mma columns (odd columns stay zero);
//   * fp32 mode ("HL"): columns 0-3 carry h_hi and columns 4-7 carry h_lo of the same 4 sequences, so ONE mma per weight part
//     yields hi*hi + hi*lo (A_hi) and lo*hi + lo*lo (A_lo): 2 MMAs per product instead of 3.  The hi-column and lo-column
//     partial sums of a sequence live in lanes tig and tig^2: one shfl_xor(2) per gate hands every lane one complete cell
//     (lanes tig<2 keep their even column = sequence 2*tig, lanes tig>=2 keep their odd column = sequence 2*(tig-2)+1).
// DEFER (HALF modes, launches with at most one CTA per SM): the global stores of a step and the input-projection read are moved
// into the NEXT step's MMA phase, interleaved with the HMMAs.  With two co-resident CTAs per SM the other CTA already fills
// those issue gaps and the deferral only costs registers (measured: -13 % alone on an SM, +3 % when sharing it).
// PHASED: the kernel takes part in the two-phase rebalancing of common.cuh (window of steps, state save / restore); the plain
// variants carry none of that code.
template <int H, bool SPLIT, bool FAST_ACT, bool LAYER0, bool TRAIN, bool HALF, bool DEFER, bool PHASED>
__global__ void __launch_bounds__(H * 4, HALF ? 2 : 1) lstm_fwd_kernel(const LstmFwdArgs p) {
  static_assert(!DEFER || HALF, "DEFER is a HALF-mode variant");
  static_assert(!PHASED || HALF, "PHASED is a HALF-mode variant");
  constexpr int NT = H * 4, KT = H / 16;
  constexpr bool HL = HALF && SPLIT;
  constexpr int NPART = (SPLIT && !HL) ? 2 : 1;
  using Smem = FwdSmem<H, NPART>;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tig = lane & 3;
  int bx = blockIdx.x, g = blockIdx.y, dz = blockIdx.z;
  if constexpr (PHASED) {
    if (!phase_cta(p.ph, p.G, bx, g, dz)) return;
  }
  const int dir = p.dir0 + dz;
  const int T = p.lens[p.G + g];  // T_eff of this group
  if (T <= 0) return;
  // steps [s_begin, s_end) of the chain run in this launch (two-phase rebalancing, common.cuh)
  const int split = PHASED ? phase_split(p.ph, T) : T;
  const int s_begin = (PHASED && p.ph.phase == 2) ? split : 0;
  int s_end = T;
  constexpr int SEQ = HALF ? kBC / 2 : kBC;  // sequences per CTA
  const int b0 = bx * SEQ;
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

  // the mma columns of this thread's accumulators and the sequence(s) of its cell(s)
  const int n0 = 2 * tig, n1 = 2 * tig + 1;
  const bool low = tig < 2;
  const int q0 = HL ? (low ? 2 * tig : 2 * (tig - 2) + 1) : (HALF ? tig : n0);
  const int q1 = HALF ? q0 : n1;
  const bool v0 = q0 < nvalid, v1 = !HALF && q1 < nvalid;

  // ---- init h = 0 (both buffers); stage this CTA's token ids in scan order (layer 0) ------------------------------------------
  if constexpr (PHASED) {
    if (tid == 0) {
      sm.shared_sm = 0;
      if (p.ph.phase == 1) atomicAdd(p.ph.sm_load + sm_id(), 1);
    }
  }
  for (int i = tid; i < 2 * NPART * H * kBC / 2; i += NT) reinterpret_cast<uint32_t*>(&sm.hs[0][0][0][0])[i] = 0u;
  if constexpr (LAYER0) {
    const int pad_slot = (int)((blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) % kPadRows);
    for (int i = tid; i < (T + kD) * SEQ; i += NT) {
      const int n = i / (T + kD), s = i % (T + kD);  // consecutive threads read consecutive time steps (coalesced)
      int v = 0;
      if (s < T && n < nvalid) v = p.tok[(size_t)(nbase + n) * Tmax + (dir ? (T - 1 - s) : s)];
      if (v == 0) v = p.V + pad_slot;  // this CTA's copy of row 0 (kernels.h: pad replicas)
      toks[s * kBC + ((HALF && !HL) ? 2 * n : n)] = (uint16_t)v;
    }
  }
  __syncthreads();

  // columns beyond the batch read a valid sequence (clamped) and never store
  const int rb0 = (nbase + min(q0, nvalid - 1)) * Tmax, rb1 = (nbase + min(q1, nvalid - 1)) * Tmax;
  const int t_first = dir ? T - 1 : 0, dt = dir ? -1 : 1;
  const float4* __restrict__ xsrc =
      LAYER0 ? reinterpret_cast<const float4*>(p.table) + (size_t)((p.table_shared ? 0 : g) * 2 + dir) * (p.V + kPadRows) * H + u
             : reinterpret_cast<const float4*>((dir ? p.xproj[1] : p.xproj[0])) + u;
  // layer >= 1: running source pointers of the prefetch (kD steps ahead of the compute), advanced by one row per step
  const float4* xp0 = xsrc + (size_t)(rb0 + t_first + s_begin * dt) * H;
  const float4* xp1 = xsrc + (size_t)(rb1 + t_first + s_begin * dt) * H;
  const ptrdiff_t xstride = (ptrdiff_t)dt * H;
  float4* slot = &sm.xring[0][0][tid];
  constexpr int kStage = 2 * NT;  // float4 per ring stage

  // one cp.async group per step; slot (s % kD) of this thread is re-armed right after it has been read
  auto issue = [&](int s) {
    float4* dst = slot + (s & (kD - 1)) * kStage;
    if constexpr (LAYER0) {
      if constexpr (HL) {
        cp_async16(dst, xsrc + (size_t)toks[s * kBC + q0] * H, true);  // rows >= T hold token 0
      } else {
        const uint32_t tw = *reinterpret_cast<const uint32_t*>(toks + s * kBC + n0);  // tokens of (n0, n1); rows >= T hold 0
        cp_async16(dst, xsrc + (size_t)(tw & 0xffffu) * H, true);
        if constexpr (!HALF) cp_async16(dst + NT, xsrc + (size_t)(tw >> 16) * H, true);
      }
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
  for (int s = 0; s < kD; ++s) issue(s_begin + s);

  float c0 = 0.f, c1 = 0.f, h0 = 0.f, h1 = 0.f;
  // output bases; the token row of my cells at the current step is (row0, row1), advanced by dt per step
  float4* const G4 = TRAIN ? reinterpret_cast<float4*>((dir ? p.gates[1] : p.gates[0])) + u : nullptr;
  float* const Cst = TRAIN ? (dir ? p.cstate[1] : p.cstate[0]) + u : nullptr;
  const bool has_y = p.y != nullptr, planes = p.planes != 0;
  const int ycol = dir * H + u;
  constexpr int ystr = 2 * H;  // == p.y_stride (checked by the launcher): a compile-time row pitch keeps the address math short
  int row0 = rb0 + t_first + s_begin * dt, row1 = rb1 + t_first + s_begin * dt;

  const __nv_bfloat16* hrow = &sm.hs[0][0][lane % H][0];  // ldmatrix row address of this lane (k-pair block 0, buffer 0)
  __nv_bfloat16* hput = &sm.hs[0][0][u][HL ? q0 : n0];
  constexpr int kBufElems = NPART * H * kBC, kPartElems = H * kBC;

  const int dbg = IB200_DBGBITS(p);

  // two-phase rebalancing: recurrent state of my cell(s), [dir slot][sequence][unit][c, h]
  auto state_slot = [&](int q) {
    return reinterpret_cast<float2*>(p.ph.state) + ((size_t)dz * p.G * p.B + nbase + min(q, nvalid - 1)) * H + u;
  };
  if (PHASED && s_begin > 0) {
    float2* st0 = state_slot(q0);
    float2* st1 = state_slot(q1);  // phase 2: resume.  h re-enters the MMA exactly as it left: the same bf16 hi (+ lo) split of the fp32 value
    const float2 a = *st0;
    c0 = a.x;
    h0 = a.y;
    if constexpr (!HALF) {
      const float2 b = *st1;
      c1 = b.x;
      h1 = b.y;
    }
    const __nv_bfloat162 hh = __floats2bfloat162_rn(h0, h1);
    __nv_bfloat16* dst = hput + (s_begin & 1) * kBufElems;
    if constexpr (HL) {
      dst[0] = hh.x;
      dst[kBC / 2] = __float2bfloat16_rn(h0 - __bfloat162float(hh.x));
    } else {
      *reinterpret_cast<__nv_bfloat162*>(dst) = hh;
      if constexpr (SPLIT) {
        const float2 hf = __bfloat1622float2(hh);
        *reinterpret_cast<__nv_bfloat162*>(dst + kPartElems) = __floats2bfloat162_rn(h0 - hf.x, h1 - hf.y);
      }
    }
    __syncthreads();
  }

  // HALF modes: the global stores of a step (one cell per thread) are DEFERRED into the MMA shadow of the next step, off the
  // chain  STS h -> barrier -> LDSM -> HMMA  that bounds the step time; the last step is flushed after the loop.
  auto emit_half = [&](int row, const float4& gt, float cc, float hv, __nv_bfloat16 hhx, __nv_bfloat16 hlx) {
    if (!v0 || (dbg & 4)) return;
    if (has_y) {
      float* yr = p.y + (size_t)row * ystr;
      if (planes) {  // the row bytes hold bf16 [hi plane: y_stride values | lo plane: y_stride values]
        __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(yr) + ycol;
        d[0] = hhx;
        if constexpr (SPLIT) d[ystr] = hlx;
      } else {
        yr[ycol] = hv;
      }
    }
    if constexpr (TRAIN) {
      G4[(size_t)row * H] = gt;
      Cst[(size_t)row * H] = cc;
    }
  };
  float4 pgt = make_float4(0.f, 0.f, 0.f, 0.f);
  float pcc = 0.f, phv = 0.f;
  __nv_bfloat16 phh = __float2bfloat16_rn(0.f), phl = phh;
  int prow = row0;

  for (int s = s_begin; s < s_end; ++s) {
    if constexpr (PHASED) {
      if (p.ph.phase == 1) {  // (uniform; two compares per step)
        if (s == kPhaseCheck) {
          if (tid == 0) sm.shared_sm = *reinterpret_cast<volatile int*>(p.ph.sm_load + sm_id()) >= 2;  // visible after this step's barrier
        } else if (s == kPhaseCheck + 1) {
          if (sm.shared_sm) s_end = split;
        }
      }
    }
    float4 x0 = make_float4(0.1f, 0.2f, 0.3f, 0.4f), x1 = x0;
    auto fetch_x = [&]() {
      cp_async_wait<kD - 1>();  // this thread's copies for step s have landed
      const float4* cur = slot + (s & (kD - 1)) * kStage;
      x0 = cur[0];
      if constexpr (!HALF) x1 = cur[NT];
      issue(s + kD);
    };
    if constexpr (!(HL && DEFER)) {  // (HL + DEFER fetches x between the HMMAs; it is only added after them)
      if (!(dbg & 8)) fetch_x();
    }

    const int buf = s & 1;
    float ai0, af0, ag0, ao0, ai1 = 0.f, af1 = 0.f, ag1 = 0.f, ao1 = 0.f;
    if constexpr (HL) {
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      float ac2[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      // work that is off the dependent chain is interleaved with the HMMAs (slot i after the i-th group of four): the previous
      // step's global stores, this step's input projection read, the prefetch kD steps ahead
      const bool st_ok = v0 && s > s_begin && !(dbg & 4);
      auto shadow = [&](int slot_i) {
        if constexpr (!DEFER) return;
        if (slot_i == 0) {
          if constexpr (TRAIN) stg_v4_if(G4 + (size_t)prow * H, pgt, st_ok);
        } else if (slot_i == 1) {
          if constexpr (TRAIN) stg_f32_if(Cst + (size_t)prow * H, pcc, st_ok);
          if (has_y) {
            float* yr = p.y + (size_t)prow * ystr;
            if (planes) {
              __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(yr) + ycol;
              stg_bf16_if(d, phh, st_ok);
              stg_bf16_if(d + ystr, phl, st_ok);
            } else {
              stg_f32_if(yr + ycol, phv, st_ok);
            }
          }
        } else if (slot_i == 2) {
          if (!(dbg & 8)) fetch_x();
        }
      };
#pragma unroll
      for (int kp = 0; kp < (KT + 1) / 2; ++kp) {
        uint32_t b[4];
        ldmatrix_x4_trans(b, hrow + buf * kBufElems + kp * 32 * kBC);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const int kt = kp * 2 + kk;
          if (kt < KT) {
            if (!(dbg & 1)) {
#pragma unroll
              for (int tile = 0; tile < 2; ++tile) {
                mma_bf16(acc[tile], Ahi[tile][kt], b[2 * kk], b[2 * kk + 1]);
                mma_bf16(ac2[tile], Alo[tile][kt], b[2 * kk], b[2 * kk + 1]);
              }
            }
            shadow(kt);
          }
        }
      }
#pragma unroll
      for (int i = KT; i < 3; ++i) shadow(i);
      // [0]=(row gq, col n0) [1]=(row gq, col n1) [2]=(row gq+8, n0) [3]=(row gq+8, n1); tile 0 rows: i,f ; tile 1 rows: g,o.
      // Keep my column (low lanes: even = hi part of my sequence; high lanes: odd = lo part), send the other one to lane^2.
      float keep[4], send[4];
#pragma unroll
      for (int tile = 0; tile < 2; ++tile)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float ev = acc[tile][2 * r] + ac2[tile][2 * r], od = acc[tile][2 * r + 1] + ac2[tile][2 * r + 1];
          keep[2 * tile + r] = low ? ev : od;
          send[2 * tile + r] = low ? od : ev;
        }
      ai0 = x0.x + keep[0] + __shfl_xor_sync(0xffffffffu, send[0], 2);
      af0 = x0.y + keep[1] + __shfl_xor_sync(0xffffffffu, send[1], 2);
      ag0 = x0.z + keep[2] + __shfl_xor_sync(0xffffffffu, send[2], 2);
      ao0 = x0.w + keep[3] + __shfl_xor_sync(0xffffffffu, send[3], 2);
    } else {
      // accumulators start from the input projection: [0]=(row gq, col n0) [1]=(row gq, col n1) [2]=(row gq+8, n0) [3]=(row gq+8, n1)
      // tile 0 rows: gq -> i_u, gq+8 -> f_u ; tile 1 rows: gq -> g_u, gq+8 -> o_u
      float acc[2][4] = {{x0.x, x1.x, x0.y, x1.y}, {x0.z, x1.z, x0.w, x1.w}};
      float ac1[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      float ac2[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
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
      if constexpr (DEFER) {
        if (s > s_begin) emit_half(prow, pgt, pcc, phv, phh, phl);  // after the HMMAs have been issued
      }
      ai0 = acc[0][0], af0 = acc[0][2], ag0 = acc[1][0], ao0 = acc[1][2];
      ai1 = acc[0][1], af1 = acc[0][3], ag1 = acc[1][1], ao1 = acc[1][3];
      if constexpr (SPLIT) {
        ai0 += ac1[0][0] + ac2[0][0]; af0 += ac1[0][2] + ac2[0][2]; ag0 += ac1[1][0] + ac2[1][0]; ao0 += ac1[1][2] + ac2[1][2];
        ai1 += ac1[0][1] + ac2[0][1]; af1 += ac1[0][3] + ac2[0][3]; ag1 += ac1[1][1] + ac2[1][1]; ao1 += ac1[1][3] + ac2[1][3];
      }
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

    // bf16 hi (and lo) of the new hidden state
    const __nv_bfloat162 hh = __floats2bfloat162_rn(h0, h1);
    __nv_bfloat162 hl = hh;
    if constexpr (SPLIT) {
      const float2 hf = __bfloat1622float2(hh);
      hl = __floats2bfloat162_rn(h0 - hf.x, h1 - hf.y);
    }
    // publish h_t for the next step
    {
      __nv_bfloat16* dst = hput + (buf ^ 1) * kBufElems;
      if constexpr (HL) {
        dst[0] = hh.x;        // column q0     : hi part
        dst[kBC / 2] = hl.x;  // column q0 + 4 : lo part
      } else {
        *reinterpret_cast<__nv_bfloat162*>(dst) = hh;  // columns (n0, n1); HALF: h1 == 0
        if constexpr (SPLIT) *reinterpret_cast<__nv_bfloat162*>(dst + kPartElems) = hl;
      }
    }
    if constexpr (DEFER) {  // remember this step's outputs; they are stored in the next step's MMA phase
      pgt = make_float4(i0, f0, gg0, o0);
      pcc = c0;
      phv = h0;
      phh = hh.x;
      phl = hl.x;
      prow = row0;
    } else if constexpr (HALF) {
      emit_half(row0, make_float4(i0, f0, gg0, o0), c0, h0, hh.x, hl.x);
    }
    // FULL mode: stream out what later stages need right away
    if (!HALF && has_y && !(dbg & 4)) {
      float* yr0 = p.y + (size_t)row0 * ystr;
      if (planes) {
        // planes layout: the row bytes hold bf16 [hi plane: y_stride values | lo plane: y_stride values]
        {
          // one shuffle with the neighbouring unit (lane ^ 4) lets every thread store a 2-unit bf16x2 word per plane
          // (even gq: units (u,u+1) of column n0; odd gq: units (u-1,u) of column n1)
          const uint32_t ab = *reinterpret_cast<const uint32_t*>(&hh), lb = *reinterpret_cast<const uint32_t*>(&hl);
          const uint32_t w0 = (ab & 0xffffu) | (lb << 16), w1 = (ab >> 16) | (lb & 0xffff0000u);  // value n0 / value n1
          const bool odd = gq & 1;
          const uint32_t recv = __shfl_xor_sync(0xffffffffu, odd ? w0 : w1, 4);
          const uint32_t mine = odd ? w1 : w0;
          const uint32_t first = odd ? recv : mine, second = odd ? mine : recv;
          const uint32_t hiw = (first & 0xffffu) | (second << 16), low_w = (first >> 16) | (second & 0xffff0000u);
          if (odd ? v1 : v0) {
            float* yr = odd ? p.y + (size_t)row1 * ystr : yr0;
            uint32_t* d = reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(yr) + (ycol & ~1));
            d[0] = hiw;
            if constexpr (SPLIT) d[ystr / 2] = low_w;
          }
        }
      } else {
        if (v0) yr0[ycol] = h0;
        if (v1) p.y[(size_t)row1 * ystr + ycol] = h1;
      }
    }
    if (!HALF && TRAIN && !(dbg & 4)) {
      if (v0) {
        G4[(size_t)row0 * H] = make_float4(i0, f0, gg0, o0);
        Cst[(size_t)row0 * H] = c0;
      }
      if (v1) {
        G4[(size_t)row1 * H] = make_float4(i1, f1, gg1, o1);
        Cst[(size_t)row1 * H] = c1;
      }
    }
    row0 += dt;
    row1 += dt;
    if (!(dbg & 16)) __syncthreads();
  }
  if constexpr (DEFER) emit_half(prow, pgt, pcc, phv, phh, phl);  // flush the last step
  cp_async_wait<0>();

  if (PHASED && s_end < T) {  // phase 1, shared SM: park the state and ask for a phase-2 CTA
    if (v0) *state_slot(q0) = make_float2(c0, h0);
    if (v1) *state_slot(q1) = make_float2(c1, h1);
    if (tid == 0) {
      const int lin = bx + p.ph.grid_x * (g + p.G * dz);
      p.ph.resume_list[1 + atomicAdd(p.ph.resume_list, 1)] = lin;
    }
    return;
  }

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

template <int H, bool SPLIT, bool FAST, bool L0, bool TR, bool HALF, bool DEFER = false, bool PHASED = false>
cudaError_t launch_kh(const LstmFwdArgs& a, cudaStream_t st) {
  constexpr int SEQ = HALF ? kBC / 2 : kBC;
  dim3 grid((a.B + SEQ - 1) / SEQ, a.G, a.ndir), block(H * 4);
  if (PHASED && a.ph.phase == 2) grid = dim3(grid.x * grid.y * grid.z, 1, 1);  // 1-D: CTA i resumes resume_list[i], the rest exit at once
  size_t smem = sizeof(FwdSmem<H, (SPLIT && !HALF) ? 2 : 1>);
  if (L0) smem += (size_t)(a.Tmax + kD) * kBC * sizeof(uint16_t);
  if (smem > 220 * 1024) return cudaErrorInvalidValue;
  if (a.y != nullptr && a.y_stride != 2 * H) return cudaErrorInvalidValue;  // the kernel addresses y rows with a fixed pitch of 2H
  cudaError_t e = cudaFuncSetAttribute(lstm_fwd_kernel<H, SPLIT, FAST, L0, TR, HALF, DEFER, PHASED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  lstm_fwd_kernel<H, SPLIT, FAST, L0, TR, HALF, DEFER, PHASED><<<grid, block, smem, st>>>(a);
  return cudaGetLastError();
}
template <int H, bool SPLIT, bool FAST, bool L0, bool TR>
cudaError_t launch_k(const LstmFwdArgs& a, cudaStream_t st) {
  // one cell per thread (4 sequences per CTA) whenever the doubled CTA count is still co-resident: HALF kernels are capped at
  // 128 registers so two of them share an SM and interleave their MMA / MUFU phases
  const int full_ctas = ((a.B + kBC - 1) / kBC) * a.G * a.ndir;
  const bool half = full_ctas <= ((a.dbg & 128) ? 74 : 148) && !(a.dbg & 64);
  if (!half) {
    LstmFwdArgs b = a;
    b.ph = PhaseArgs{};
    return launch_kh<H, SPLIT, FAST, L0, TR, false>(b, st);
  }
  const int tiles = (a.B + kBC / 2 - 1) / (kBC / 2), half_ctas = tiles * a.G * a.ndir;
  const bool defer = (half_ctas <= 148 || (a.dbg & 512)) && !(a.dbg & 1024);  // at most one CTA per SM
  if (defer) {
    LstmFwdArgs b = a;
    b.ph = PhaseArgs{};
    return launch_kh<H, SPLIT, FAST, L0, TR, true, true>(b, st);
  }
  if (a.ph.state != nullptr && half_ctas < 2 * 148 && a.Tmax >= 8 * kPhaseCheck && !(a.dbg & 2048)) {
    // between one and two CTAs per SM: two-phase rebalancing (common.cuh).  Phase 2 has at most one CTA per SM => DEFER variant.
    LstmFwdArgs b = a;
    b.ph.phase = 1;
    b.ph.grid_x = tiles;
    cudaError_t e = launch_kh<H, SPLIT, FAST, L0, TR, true, false, true>(b, st);
    if (e != cudaSuccess) return e;
    b.ph.phase = 2;
    return launch_kh<H, SPLIT, FAST, L0, TR, true, true, true>(b, st);
  }
  LstmFwdArgs b = a;
  b.ph = PhaseArgs{};
  return launch_kh<H, SPLIT, FAST, L0, TR, true, false>(b, st);
}


template <int H, bool SPLIT, bool FAST>
cudaError_t launch_h(const LstmFwdArgs& a, cudaStream_t st) {
  const bool l0 = a.tok != nullptr, tr = a.gates[a.dir0] != nullptr;
  if (l0 && a.V + kPadRows > 65536) return cudaErrorInvalidValue;  // token ids are staged as uint16
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
