// K2c / K3c: recurrent forward and BPTT kernels for hidden sizes the register-resident kernels do not cover (H = 96 .. 256,
// multiple of 32; e.g. the stress configuration E = H = 256, 3 layers).  Same math and the same HBM layouts as lstm_fwd.cu /
// lstm_bwd.cu (reference: nn.LSTM inside encoders/awd_lstm.py:35-41,56).
//
// W_hh (4H x H) no longer fits one SM, so a thread-block CLUSTER of C = H/32 CTAs shares one tile of sequences:
//   * CTA r owns hidden units [32r, 32r+32): its 128 gate rows of W_hh stay resident in shared memory as bf16 hi (+ lo) for the
//     whole kernel, in the row order (warp w: tile (i_u,f_u) then tile (g_u,o_u) for its 8 units) that leaves i,f,g,o of one
//     cell in one thread after the MMAs -- exactly the fragment mapping of the H=64 kernels;
//   * forward: every CTA needs the full h_{t-1} (H x NS) as its B operand: each CTA sends its 32-unit slice of h_t into the h tile
//     of ALL C CTAs through distributed shared memory with 16-byte st.async stores whose bytes are counted (complete_tx) on an
//     mbarrier of the RECEIVING CTA, double buffered; a CTA waits on its own barrier for "all of h_t has landed" -- there is no
//     cluster barrier in the step loop.  The sequence groups of a cluster (warps with the same nh) have their own barriers and run
//     as independent chains;
//   * backward: dh = W_hh^T da needs all 4H rows of da as K.  Each CTA multiplies its own 128-row K slice (A = the same
//     resident W slice read transposed with ldmatrix.trans) into partial sums for ALL H units and REDUCE-SCATTERS them: the
//     partials of units [32q, 32q+32) go to CTA q's exchange buffer with the same st.async + mbarrier scheme; the owner adds the C
//     partials once its barrier reports them complete.
// Measured on B200 at the config-5 shape (profiles/r2_cluster_*): the kernels are latency bound per warp (ncu: tensor pipe 36 %,
// issue slots 22 %, 8 warps per SM, ~950 warp-instructions per warp and step), not bound by the exchange or by shared memory.
// NS = 8*NTILE sequences per cluster (NTILE n-tiles of the m16n8k16 MMA) so large batches amortise the A-fragment traffic.
#include <algorithm>

#include "kernels.h"
#include "tc05.cuh"

namespace ib200 {
namespace {

using tc::mbar_arrive_expect_tx;
using tc::mbar_init;
using tc::mbar_init_fence;
using tc::mbar_wait;

// timing experiments only (IB200_ABLATE=1 builds, env IB200_DBG): 1 no remote h stores, 2 CTA barrier instead of the cluster barrier,
// 4 no global stores, 8 no MMAs.  Production builds compile the switches out.
#ifdef IB200_ABLATE
#define IB200_CL_DBG(p) ((p).dbg)
#else
#define IB200_CL_DBG(p) 0
#endif

constexpr int kUS = 32;          // hidden units per CTA
constexpr int kRows = 4 * kUS;   // gate rows per CTA
constexpr int kUnitWarps = 4;     // warps along the units: warp group ug owns local units 8ug .. 8ug+7
// NWG warp groups along the sequences (1 or 2): with 32 sequences per cluster, 8 warps (two per scheduler) keep the tensor pipe fed,
// one warp per scheduler is limited by its own HMMA issue interval; warp (ug, nh) owns units 8ug.. and n-tiles [nh*NTW, (nh+1)*NTW)
constexpr int kWPad = 8;         // bf16 padding of a W slice row (16 bytes): conflict-free ldmatrix
constexpr int kNPad = 8;         // bf16 padding of an h / da tile row

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory"); }
__device__ __forceinline__ uint32_t map_to_rank(const void* local_smem, uint32_t rank) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(local_smem)), r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
// 16-byte asynchronous store into the shared memory of a CTA of the cluster; its completion is counted (complete_tx, 16 bytes) on an
// mbarrier of THAT CTA, so the receiver waits on a local barrier for "all bytes of this step have landed" -- no cluster-wide barrier
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];\n" ::"r"(addr), "r"(a),
               "r"(b), "r"(c), "r"(d), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a)
               : "memory");
}

// local gate row lr (0..127) -> (local unit, gate): warp w = lr/32 owns units 8w..8w+7; inside a warp tile 0 holds rows
// (i_u | f_u), tile 1 holds (g_u | o_u), 8 units each -- the accumulator rows gq / gq+8 of one thread are gates of ONE unit
__device__ __forceinline__ void local_row(int lr, int& ul, int& gate) {
  const int w = lr >> 5, t2 = (lr >> 4) & 1, r = lr & 15;
  ul = 8 * w + (r & 7);
  gate = 2 * t2 + (r >> 3);
}

// resident W slice: Wsm[part][lr][H + kWPad] bf16
template <bool SPLIT>
__device__ __forceinline__ void load_w_slice(__nv_bfloat16* Wsm, const float* __restrict__ W, const float* __restrict__ M, int H,
                                             int rank) {
  const int ldw = H + kWPad;
  for (int idx = threadIdx.x; idx < kRows * (H / 2); idx += blockDim.x) {
    const int lr = idx / (H / 2), k = (idx % (H / 2)) * 2;
    int ul, gate;
    local_row(lr, ul, gate);
    const size_t src = ((size_t)gate * H + kUS * rank + ul) * H + k;
    float w0 = W[src], w1 = W[src + 1];
    if (M != nullptr) {
      w0 *= M[src];
      w1 *= M[src + 1];
    }
    uint32_t hi, lo = 0u;
    if constexpr (SPLIT) split_bf16(w0, w1, hi, lo);
    else hi = pack_bf16(w0, w1);
    *reinterpret_cast<uint32_t*>(Wsm + (size_t)lr * ldw + k) = hi;
    if constexpr (SPLIT) *reinterpret_cast<uint32_t*>(Wsm + (size_t)(kRows + lr) * ldw + k) = lo;
  }
}

// =================================================================================================================================
// forward
// =================================================================================================================================
template <int NTILE, bool SPLIT, int NWG>
__global__ void __launch_bounds__(32 * kUnitWarps * NWG, 1) lstm_fwd_cl_kernel(const LstmFwdArgs p, const int H) {
  static_assert(NTILE % NWG == 0, "n-tiles split evenly over the warp groups");
  constexpr int NS = 8 * NTILE, NSP = NS + kNPad, NPART = SPLIT ? 2 : 1, NTW = NTILE / NWG, NCELL = 2 * NTW;
  constexpr int kClThreads = 32 * kUnitWarps * NWG;
  constexpr bool FAST = !SPLIT;
  const int C = H / kUS, KT = H / 16, ldw = H + kWPad;
  const int tid = threadIdx.x, lane = tid & 31, gq = lane >> 2, tig = lane & 3;
  const int warp = (tid >> 5) & (kUnitWarps - 1), nh = (tid >> 5) / kUnitWarps;  // unit group, sequence group
  const int ncol0 = nh * NTW * 8;  // first mma column (sequence) of this warp
  const int rank = (int)cluster_ctarank(), tile = blockIdx.x / C;
  const int g = blockIdx.y, dir = p.dir0 + (int)blockIdx.z;
  const int T = p.lens[p.G + g];
  if (T <= 0) return;  // uniform over the cluster
  const int b0 = tile * NS, nvalid = min(NS, p.B - b0), nbase = g * p.B + b0, Tmax = p.Tmax;
  const bool layer0 = p.tok != nullptr, train = p.gates[dir] != nullptr;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* Wsm = reinterpret_cast<__nv_bfloat16*>(smem_raw);           // [NPART][128][ldw]
  __nv_bfloat16* hs = Wsm + (size_t)NPART * kRows * ldw;                      // [2][NPART][H][NSP]
  const int kPartElems = H * NSP, kBufElems = NPART * kPartElems;
  // hbar[buffer][sequence group]: "h of this group's columns has landed in buffer b" (complete_tx bytes).  The NWG sequence groups
  // (warps with the same nh) never read each other's columns, so each group runs its own chain of steps and only waits for its own
  // columns: the groups drift apart and one group's MMAs fill the exchange / activation latency of the other
  uint64_t* hbar = reinterpret_cast<uint64_t*>(hs + (size_t)2 * kBufElems);
  const uint32_t kStepBytes = (uint32_t)(NPART * H * (NS / NWG) * 2);  // bytes a group receives per step: its columns of the h tile

  {
    const float* __restrict__ W = dir ? p.whh[1] : p.whh[0];
    const float* __restrict__ M = (dir == 0 && p.whh_mask != nullptr) ? p.whh_mask + (size_t)g * 4 * H * H : nullptr;
    load_w_slice<SPLIT>(Wsm, W, M, H, rank);
  }
  for (int i = tid; i < 2 * kBufElems / 2; i += kClThreads) reinterpret_cast<uint32_t*>(hs)[i] = 0u;
  if (tid == 0) {
    for (int i = 0; i < 2 * NWG; ++i) mbar_init(&hbar[i], 1);
    mbar_init_fence();
    if (T > 1)
      for (int i = 0; i < NWG; ++i) mbar_arrive_expect_tx(&hbar[NWG + i], kStepBytes);  // armed for h_0, written at the end of step 0
  }
  __syncthreads();
  cluster_arrive();  // every CTA's tiles and barriers are initialised before any remote h write lands
  cluster_wait();

  const int ul = 8 * warp + gq, u = kUS * rank + ul;  // the unit of this thread's cells
  int rb[NCELL];
  bool valid[NCELL];
#pragma unroll
  for (int c = 0; c < NCELL; ++c) {
    const int q = ncol0 + 8 * (c >> 1) + 2 * tig + (c & 1);  // sequence (= mma column) of cell c
    valid[c] = q < nvalid;
    rb[c] = (nbase + min(q, nvalid - 1)) * Tmax;  // columns beyond the batch read a valid sequence and never store
  }
  const int t_first = dir ? T - 1 : 0, dt = dir ? -1 : 1;
  const int pad_row = p.V + (int)((tile + gridDim.x / C * (blockIdx.y + gridDim.y * blockIdx.z)) % kPadRows);
  const float4* __restrict__ xsrc = layer0 ? reinterpret_cast<const float4*>(p.table) + (size_t)((p.table_shared ? 0 : g) * 2 + dir) * (p.V + kPadRows) * H + u
                                           : reinterpret_cast<const float4*>(dir ? p.xproj[1] : p.xproj[0]) + u;
  auto load_x = [&](int s, float4 (&x)[NCELL]) {
    const int t = t_first + s * dt;
#pragma unroll
    for (int c = 0; c < NCELL; ++c) {
      if (layer0) {
        const int tk = p.tok[(size_t)rb[c] + t];
        x[c] = __ldg(xsrc + (size_t)(tk == 0 ? pad_row : tk) * H);  // pads read this cluster's copy of row 0 (kernels.h)
      } else {
        x[c] = __ldg(xsrc + (size_t)(rb[c] + t) * H);
      }
    }
  };
  float4 xn[NCELL];
  load_x(0, xn);

  float4* const G4 = train ? reinterpret_cast<float4*>(dir ? p.gates[1] : p.gates[0]) + u : nullptr;
  float* const Cst = train ? (dir ? p.cstate[1] : p.cstate[0]) + u : nullptr;
  const bool has_y = p.y != nullptr, planes = p.planes != 0;
  const int ycol = dir * H + u, ystr = p.y_stride;

  // h exchange: the four lanes that share a unit (same gq) hold the 8 columns of an n-tile = one 16-byte piece of the h tile row;
  // after a 4-lane gather, lane tig sends that piece to CTAs tig and tig + 4 of the cluster with ONE 16-byte st.async each
  uint32_t hdst[2], rbar[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = tig + 4 * i;
    hdst[i] = r < C ? map_to_rank(hs + (size_t)u * NSP + ncol0, (uint32_t)r) : 0u;
    rbar[i] = r < C ? map_to_rank(hbar + nh, (uint32_t)r) : 0u;
  }
  const bool group_leader = (tid & (32 * kUnitWarps - 1)) == 0;

  // ldmatrix lane addresses: A (non transposed) row = tile*16 + (lane&7) + (lane&8), k offset (lane&16 ? 8 : 0)
  const __nv_bfloat16* a_lane = Wsm + (size_t)((2 * warp) * 16 + (lane & 7) + (lane & 8)) * ldw + ((lane & 16) ? 8 : 0);
  // B (transposed): k row = lane & 15, n-tile offset (lane & 16 ? 8 : 0) (clamped to the last n-tile for NTILE == 1)
  const __nv_bfloat16* b_lane = hs + (size_t)(lane & 15) * NSP + ncol0 + ((NTW > 1 && (lane & 16)) ? 8 : 0);

  float cst[NCELL], hv[NCELL];
#pragma unroll
  for (int c = 0; c < NCELL; ++c) cst[c] = hv[c] = 0.f;
  const int dbg = IB200_CL_DBG(p);

  for (int s = 0; s < T; ++s) {
    float4 x[NCELL];
#pragma unroll
    for (int c = 0; c < NCELL; ++c) x[c] = xn[c];
    if (s + 1 < T) load_x(s + 1, xn);  // register prefetch, one step ahead
    const int buf = s & 1;
    if (s > 0) mbar_wait(&hbar[buf * NWG + nh], (uint32_t)(((s - 1) >> 1) & 1));  // h_{s-1} of my group's columns has landed in `buf`
    if (group_leader && s + 2 < T) mbar_arrive_expect_tx(&hbar[buf * NWG + nh], kStepBytes);  // next fill: end of step s + 1

    // acc[t2][j]: [0]=(row gq, col n0) [1]=(row gq, col n1) [2]=(row gq+8, n0) [3]=(row gq+8, n1); t2=0: rows i,f ; t2=1: rows g,o
    float acc[2][NTW][4];
#pragma unroll
    for (int j = 0; j < NTW; ++j) {
      acc[0][j][0] = x[2 * j].x; acc[0][j][1] = x[2 * j + 1].x; acc[0][j][2] = x[2 * j].y; acc[0][j][3] = x[2 * j + 1].y;
      acc[1][j][0] = x[2 * j].z; acc[1][j][1] = x[2 * j + 1].z; acc[1][j][2] = x[2 * j].w; acc[1][j][3] = x[2 * j + 1].w;
    }
    const __nv_bfloat16* hb = b_lane + (size_t)buf * kBufElems;
#pragma unroll 2
    for (int kt = 0; kt < ((dbg & 8) ? 0 : KT); ++kt) {
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int t2 = 0; t2 < 2; ++t2) {
        ldmatrix_x4(ah[t2], a_lane + (size_t)t2 * 16 * ldw + kt * 16);
        if constexpr (SPLIT) ldmatrix_x4(al[t2], a_lane + (size_t)(kRows + t2 * 16) * ldw + kt * 16);
      }
#pragma unroll
      for (int jp = 0; jp < (NTW + 1) / 2; ++jp) {
        uint32_t bh[4], bl[4];
        ldmatrix_x4_trans(bh, hb + (size_t)kt * 16 * NSP + jp * 16);
        if constexpr (SPLIT) ldmatrix_x4_trans(bl, hb + kPartElems + (size_t)kt * 16 * NSP + jp * 16);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int j = 2 * jp + jj;
          if (j < NTW) {
#pragma unroll
            for (int t2 = 0; t2 < 2; ++t2) {
              mma_bf16(acc[t2][j], ah[t2], bh[2 * jj], bh[2 * jj + 1]);
              if constexpr (SPLIT) {
                mma_bf16(acc[t2][j], ah[t2], bl[2 * jj], bl[2 * jj + 1]);
                mma_bf16(acc[t2][j], al[t2], bh[2 * jj], bh[2 * jj + 1]);
              }
            }
          }
        }
      }
    }

    const int t = t_first + s * dt;
#pragma unroll
    for (int c = 0; c < NCELL; ++c) {
      const int j = c >> 1, o = c & 1;
      const float gi = sigmoid_f<FAST>(acc[0][j][o]), gf = sigmoid_f<FAST>(acc[0][j][2 + o]);
      const float gg = tanh_f<FAST>(acc[1][j][o]), go = sigmoid_f<FAST>(acc[1][j][2 + o]);
      cst[c] = fmaf(gf, cst[c], gi * gg);
      hv[c] = go * tanh_f<FAST>(cst[c]);
      if (valid[c] && !(dbg & 4)) {
        const size_t row = (size_t)(rb[c] + t);
        if (train) {
          G4[row * H] = make_float4(gi, gf, gg, go);
          Cst[row * H] = cst[c];
        }
        if (has_y) {
          if (planes) {
            // bf16 hi | lo planes over the bytes of the fp32 row (what the TMA-fed tcgen05 GEMMs read, gemm_wide.cu)
            __nv_bfloat16* yrow = reinterpret_cast<__nv_bfloat16*>(p.y + row * ystr);
            const __nv_bfloat16 hb16 = __float2bfloat16_rn(hv[c]);
            yrow[ycol] = hb16;
            if constexpr (SPLIT) yrow[ystr + ycol] = __float2bfloat16_rn(hv[c] - __bfloat162float(hb16));
          } else {
            p.y[row * ystr + ycol] = hv[c];
          }
        }
      }
    }
    // publish my slice of h_t to every CTA of the cluster (buffer buf^1); the receivers count the bytes on their hbar[buf^1].
    // No barrier: a CTA can only be one step ahead of the slowest one (it needs everybody's h), and with two buffers a buffer is
    // rewritten only after every reader has sent the h it computed from it.
    if (s + 1 < T) {
      const uint32_t boff = (uint32_t)((buf ^ 1) * kBufElems) * 2u, bar_off = (uint32_t)((buf ^ 1) * NWG) * 8u;
      const int lbase = lane & ~3;
#pragma unroll
      for (int j = 0; j < NTW; ++j) {
        uint32_t hi, lo = 0u;
        if constexpr (SPLIT) split_bf16(hv[2 * j], hv[2 * j + 1], hi, lo);
        else hi = pack_bf16(hv[2 * j], hv[2 * j + 1]);
        uint32_t wh[4], wl[4];
#pragma unroll
        for (int t4 = 0; t4 < 4; ++t4) {
          wh[t4] = __shfl_sync(0xffffffffu, hi, lbase | t4);
          if constexpr (SPLIT) wl[t4] = __shfl_sync(0xffffffffu, lo, lbase | t4);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
          if (tig + 4 * i < C) {
            st_async_v4(hdst[i] + boff + j * 16, wh[0], wh[1], wh[2], wh[3], rbar[i] + bar_off);
            if constexpr (SPLIT)
              st_async_v4(hdst[i] + boff + (uint32_t)kPartElems * 2u + j * 16, wl[0], wl[1], wl[2], wl[3], rbar[i] + bar_off);
          }
      }
    }
  }
  cluster_arrive();  // nobody exits while a peer could still be sending to it
  cluster_wait();

  // planes mode: the weight-gradient GEMM reads whole 64-row TMA boxes (and the row after the last one for the shifted operand): rows
  // [T, tail_end) of this cluster's sequences must be zeros in this CTA's 32 columns of this direction (both planes)
  if (has_y && planes) {
    const int tail_end = min(Tmax, ((T + 63) / 64) * 64 + 1), ntail = tail_end - T;
    constexpr int kChunks = kUS * 2 / 16;  // 16-byte chunks of 32 bf16
    for (int i = tid; i < nvalid * ntail * kChunks * 2; i += kClThreads) {
      const int cc = i % kChunks, pl = (i / kChunks) & 1, r = (i / (2 * kChunks)) % ntail, q = i / (2 * kChunks * ntail);
      unsigned char* row = reinterpret_cast<unsigned char*>(p.y + ((size_t)(nbase + q) * Tmax + T + r) * ystr);
      *reinterpret_cast<uint4*>(row + pl * ystr * 2 + (dir * H + kUS * rank) * 2 + cc * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }

  if (p.hn != nullptr) {
    const size_t N = (size_t)p.G * p.B;
#pragma unroll
    for (int c = 0; c < NCELL; ++c) {
      const int q = ncol0 + 8 * (c >> 1) + 2 * tig + (c & 1);
      if (valid[c]) p.hn[((size_t)dir * N + nbase + q) * H + u] = hv[c];
    }
  }
}

// =================================================================================================================================
// backward
// =================================================================================================================================
template <int NTILE, bool SPLIT, int NWG>
__global__ void __launch_bounds__(32 * kUnitWarps * NWG, 1) lstm_bwd_cl_kernel(const LstmBwdArgs p, const int H) {
  static_assert(NTILE % NWG == 0, "n-tiles split evenly over the warp groups");
  constexpr int NS = 8 * NTILE, NSP = NS + kNPad, NPART = SPLIT ? 2 : 1, NTW = NTILE / NWG, NCELL = 2 * NTW;
  constexpr int kWarps = kUnitWarps * NWG;
  constexpr bool FAST = !SPLIT;
  const int C = H / kUS, MT = H / 16, ldw = H + kWPad;
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31, gq = lane >> 2, tig = lane & 3;
  const int warp = wid & (kUnitWarps - 1), nh = wid / kUnitWarps;  // unit group (cells), sequence group (cells)
  const int ncol0 = nh * NTW * 8;
  const int rank = (int)cluster_ctarank(), tile = blockIdx.x / C;
  const int g = blockIdx.y, dir = p.dir0 + (int)blockIdx.z;
  const int T = p.lens[p.G + g];
  if (T <= 0) return;
  const int b0 = tile * NS, nvalid = min(NS, p.B - b0), nbase = g * p.B + b0, Tmax = p.Tmax;
  const size_t N = (size_t)p.G * p.B;
  const bool has_dy = p.dy != nullptr, planes = p.planes != 0;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* Wsm = reinterpret_cast<__nv_bfloat16*>(smem_raw);      // [NPART][128][ldw]
  __nv_bfloat16* das = Wsm + (size_t)NPART * kRows * ldw;                // [NPART][128][NSP]: da of my units, k = local gate row
  float* xbuf = reinterpret_cast<float*>(das + (size_t)NPART * kRows * NSP);  // [2][C][32][NS] partial dh for my units, per source CTA
  const int kXBuf = C * kUS * NS;
  // xbar[buffer][sequence group]: "all C partials of this group's columns have landed in exchange buffer b".  As in the forward
  // kernel the NWG sequence groups are independent chains: own columns of `das`, own named barrier, own exchange barrier
  uint64_t* xbar = reinterpret_cast<uint64_t*>(xbuf + (size_t)2 * kXBuf);
  const uint32_t kXBytes = (uint32_t)(kXBuf / NWG) * 4u;

  {
    const float* __restrict__ W = dir ? p.whh[1] : p.whh[0];
    const float* __restrict__ M = (dir == 0 && p.whh_mask != nullptr) ? p.whh_mask + (size_t)g * 4 * H * H : nullptr;
    load_w_slice<SPLIT>(Wsm, W, M, H, rank);
  }
  if (tid == 0) {
    for (int i = 0; i < 2 * NWG; ++i) mbar_init(&xbar[i], 1);
    mbar_init_fence();
    for (int i = 0; i < NWG; ++i) {
      if (T > 1) mbar_arrive_expect_tx(&xbar[NWG + i], kXBytes);  // step s fills buffer (s + 1) & 1: step 0 -> buffer 1, step 1 -> buffer 0
      if (T > 2) mbar_arrive_expect_tx(&xbar[i], kXBytes);
    }
  }
  __syncthreads();
  cluster_arrive();
  cluster_wait();

  const int ul = 8 * warp + gq, u = kUS * rank + ul;
  int rb[NCELL];
  bool valid[NCELL];
#pragma unroll
  for (int c = 0; c < NCELL; ++c) {
    const int q = ncol0 + 8 * (c >> 1) + 2 * tig + (c & 1);
    valid[c] = q < nvalid;
    rb[c] = (nbase + min(q, nvalid - 1)) * Tmax;
  }
  // backward scan: s = 0..T-1 visits t = T-1..0 (forward chain) or t = 0..T-1 (reverse chain)
  const int t_first = dir ? 0 : T - 1, dt = dir ? 1 : -1;
  float4* const G4 = reinterpret_cast<float4*>(dir ? p.gates[1] : p.gates[0]) + u;
  const float* const Cst = (dir ? p.cstate[1] : p.cstate[0]) + u;
  const float* const DY = has_dy ? p.dy + dir * H + u : nullptr;

  struct In {
    float4 g[NCELL];
    float cprev[NCELL], dy[NCELL];
  };
  auto load_in = [&](int s, In& in) {
    const int t = t_first + s * dt;
    const bool has_prev = s + 1 < T;
#pragma unroll
    for (int c = 0; c < NCELL; ++c) {
      const size_t row = (size_t)(rb[c] + t);
      in.g[c] = G4[row * H];
      in.cprev[c] = has_prev ? Cst[(size_t)(rb[c] + t + dt) * H] : 0.f;  // c of the scan predecessor; 0 at the chain start
      in.dy[c] = has_dy ? DY[row * p.dy_stride] : 0.f;
    }
  };
  In nxt;
  load_in(0, nxt);
  float ccur[NCELL], dc[NCELL], dhrec[NCELL];
  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);  // column sums of my cells' dgates over all steps (bias gradient partials)
#pragma unroll
  for (int c = 0; c < NCELL; ++c) {
    const int q = ncol0 + 8 * (c >> 1) + 2 * tig + (c & 1);
    ccur[c] = Cst[(size_t)(rb[c] + t_first) * H];
    dc[c] = 0.f;
    dhrec[c] = p.dhn != nullptr ? p.dhn[((size_t)dir * N + nbase + min(q, nvalid - 1)) * H + u] : 0.f;
  }

  // my da rows in the K order of the resident slice: i -> (2w)*16+gq, f -> +8, g -> (2w+1)*16+gq, o -> +8
  __nv_bfloat16* da_i = das + (size_t)((2 * warp) * 16 + gq) * NSP + ncol0 + 2 * tig;
  const int kDaPart = kRows * NSP;
  // A = W slice read transposed: m = unit j (columns of the slice), k = local gate row.  matrix l/8: k + (l&16 ? 8:0), m + (l&8 ? 8:0)
  const __nv_bfloat16* a_lane = Wsm + (size_t)((lane & 7) + ((lane & 16) ? 8 : 0)) * ldw + ((lane & 8) ? 8 : 0);
  const __nv_bfloat16* b_lane = das + (size_t)(lane & 15) * NSP + ncol0 + ((NTW > 1 && (lane & 16)) ? 8 : 0);
  const bool group_leader = (tid & (32 * kUnitWarps - 1)) == 0;
  // remote exchange slots: partials of units [32q,32q+32) go to CTA q, slot [src = my rank][unit % 32][column]
  const float* xslot = xbuf + (size_t)rank * kUS * NS;

  for (int s = 0; s < T; ++s) {
    In in = nxt;
    if (s + 1 < T) load_in(s + 1, nxt);
    const int t = t_first + s * dt;
#pragma unroll
    for (int c = 0; c < NCELL; ++c) {
      const float gi = in.g[c].x, gf = in.g[c].y, gg = in.g[c].z, go = in.g[c].w;
      const float dh = dhrec[c] + in.dy[c];
      const float tc = tanh_f<FAST>(ccur[c]);
      const float d_o = dh * tc;
      const float dct = fmaf(dh * go, fmaf(-tc, tc, 1.0f), dc[c]);
      const float d_i = dct * gg, d_g = dct * gi, d_f = dct * in.cprev[c];
      dc[c] = dct * gf;
      ccur[c] = in.cprev[c];
      const float da_ii = d_i * gi * (1.0f - gi), da_f = d_f * gf * (1.0f - gf);
      const float da_g = d_g * fmaf(-gg, gg, 1.0f), da_o = d_o * go * (1.0f - go);
      in.g[c] = make_float4(da_ii, da_f, da_g, da_o);
      if (valid[c]) {  // dgates overwrite the saved gates in place
        if (planes) {
          // bf16 hi | lo planes over the 4H-float gate row: [4H bf16 hi | 4H bf16 lo], gate-interleaved column 4u + q
          unsigned char* grow = reinterpret_cast<unsigned char*>(G4 + (size_t)(rb[c] + t) * H) - (size_t)u * 16;  // start of the row
          uint32_t h0, h1, l0 = 0u, l1 = 0u;
          if constexpr (SPLIT) {
            split_bf16(da_ii, da_f, h0, l0);
            split_bf16(da_g, da_o, h1, l1);
          } else {
            h0 = pack_bf16(da_ii, da_f);
            h1 = pack_bf16(da_g, da_o);
          }
          *reinterpret_cast<uint2*>(grow + (size_t)u * 8) = make_uint2(h0, h1);
          if constexpr (SPLIT) *reinterpret_cast<uint2*>(grow + (size_t)H * 8 + (size_t)u * 8) = make_uint2(l0, l1);
          bsum.x += da_ii; bsum.y += da_f; bsum.z += da_g; bsum.w += da_o;
        } else {
          G4[(size_t)(rb[c] + t) * H] = in.g[c];
        }
      }
    }
    if (s + 1 == T) break;
    // da -> smem B tile (bf16 hi / lo), packed pairs of columns (n0, n1)
#pragma unroll
    for (int j = 0; j < NTW; ++j) {
      const float4 a = in.g[2 * j], b = in.g[2 * j + 1];
      const float v0[4] = {a.x, a.y, a.z, a.w}, v1[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t hi, lo = 0u;
        if constexpr (SPLIT) split_bf16(v0[q], v1[q], hi, lo);
        else hi = pack_bf16(v0[q], v1[q]);
        __nv_bfloat16* dst = da_i + (size_t)(((q >> 1) * 16) + ((q & 1) * 8)) * NSP + 8 * j;
        *reinterpret_cast<uint32_t*>(dst) = hi;
        if constexpr (SPLIT) *reinterpret_cast<uint32_t*>(dst + kDaPart) = lo;
      }
    }
    // my group's columns of `das` are complete (the four warps of the group own the 128 gate rows between them)
    asm volatile("bar.sync %0, %1;\n" ::"r"(1 + nh), "r"(32 * kUnitWarps) : "memory");

    // partial dh^T[H, my columns] = Wslice^T[H, 128] * da[128, my columns]; warp (ug, nh) takes the unit tiles mt = ug, ug+4, ...
    const uint32_t xoff = (uint32_t)(((s + 1) & 1) * kXBuf) * 4u;
    for (int mt = warp; mt < MT; mt += kUnitWarps) {
      float acc[NTW][4];
#pragma unroll
      for (int j = 0; j < NTW; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
      for (int kt = 0; kt < kRows / 16; ++kt) {
        uint32_t ah[4], al[4];
        ldmatrix_x4_trans(ah, a_lane + (size_t)kt * 16 * ldw + mt * 16);
        if constexpr (SPLIT) ldmatrix_x4_trans(al, a_lane + (size_t)(kRows + kt * 16) * ldw + mt * 16);
#pragma unroll
        for (int jp = 0; jp < (NTW + 1) / 2; ++jp) {
          uint32_t bh[4], bl[4];
          ldmatrix_x4_trans(bh, b_lane + (size_t)kt * 16 * NSP + jp * 16);
          if constexpr (SPLIT) ldmatrix_x4_trans(bl, b_lane + kDaPart + (size_t)kt * 16 * NSP + jp * 16);
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int j = 2 * jp + jj;
            if (j < NTW) {
              mma_bf16(acc[j], ah, bh[2 * jj], bh[2 * jj + 1]);
              if constexpr (SPLIT) {
                mma_bf16(acc[j], ah, bl[2 * jj], bl[2 * jj + 1]);
                mma_bf16(acc[j], al, bh[2 * jj], bh[2 * jj + 1]);
              }
            }
          }
        }
      }
      // rows: unit mt*16+gq ([0],[1]) and mt*16+gq+8 ([2],[3]); columns 8j+2tig, +1.  Lane pairs (tig, tig^1) swap halves so that the
      // even lane owns 4 consecutive columns of row gq and the odd lane 4 consecutive columns of row gq+8: ONE 16-byte st.async each,
      // counted on the owner CTA's xbar (owner: 32 units per CTA = 2 unit tiles)
      const uint32_t owner = (uint32_t)(mt >> 1);
      const uint32_t xdst = map_to_rank(xslot, owner), xbr = map_to_rank(xbar + nh, owner) + (uint32_t)(((s + 1) & 1) * NWG) * 8u;
      const bool odd = (tig & 1) != 0;
      const int xrow = (mt & 1) * 16 + gq + (odd ? 8 : 0);
#pragma unroll
      for (int j = 0; j < NTW; ++j) {
        const float r0 = __shfl_xor_sync(0xffffffffu, odd ? acc[j][0] : acc[j][2], 1);
        const float r1 = __shfl_xor_sync(0xffffffffu, odd ? acc[j][1] : acc[j][3], 1);
        const uint32_t col = (uint32_t)(ncol0 + 8 * j + 2 * (tig & ~1));
        const uint32_t a0 = __float_as_uint(odd ? r0 : acc[j][0]), a1 = __float_as_uint(odd ? r1 : acc[j][1]);
        const uint32_t a2 = __float_as_uint(odd ? acc[j][2] : r0), a3 = __float_as_uint(odd ? acc[j][3] : r1);
        st_async_v4(xdst + xoff + ((uint32_t)xrow * NS + col) * 4u, a0, a1, a2, a3, xbr);
      }
    }
    // all C partials of my group's columns have landed (every CTA of the cluster, myself included, sent 32 units x my columns).
    // No cluster barrier: passing this wait also means every warp of my group has finished reading `das` (sends follow the MMAs)
    mbar_wait(&xbar[((s + 1) & 1) * NWG + nh], (uint32_t)((s >> 1) & 1));
    if (group_leader && s + 3 < T) mbar_arrive_expect_tx(&xbar[((s + 1) & 1) * NWG + nh], kXBytes);  // refilled at step s + 2
    // recurrent gradient of my cells for the next step: sum of the C partials
    const float* xb = xbuf + (size_t)((s + 1) & 1) * kXBuf + (size_t)ul * NS;
#pragma unroll
    for (int c = 0; c < NCELL; ++c) {
      const int col = ncol0 + 8 * (c >> 1) + 2 * tig + (c & 1);
      float sum = 0.f;
      for (int r = 0; r < C; ++r) sum += xb[(size_t)r * kUS * NS + col];
      dhrec[c] = sum;
    }
  }
  cluster_arrive();  // nobody exits while a peer could still be sending to it
  cluster_wait();

  if (planes) {
    // (1) bias-gradient partials: one [4H] row (GI order) per (direction slot, group, tile); CTA `rank` owns columns [128 rank, +128)
    if (p.bias_partial != nullptr) {
      bsum.x += __shfl_xor_sync(0xffffffffu, bsum.x, 1); bsum.y += __shfl_xor_sync(0xffffffffu, bsum.y, 1);
      bsum.z += __shfl_xor_sync(0xffffffffu, bsum.z, 1); bsum.w += __shfl_xor_sync(0xffffffffu, bsum.w, 1);
      bsum.x += __shfl_xor_sync(0xffffffffu, bsum.x, 2); bsum.y += __shfl_xor_sync(0xffffffffu, bsum.y, 2);
      bsum.z += __shfl_xor_sync(0xffffffffu, bsum.z, 2); bsum.w += __shfl_xor_sync(0xffffffffu, bsum.w, 2);
      float4* red = reinterpret_cast<float4*>(xbuf);  // [NWG][32 units] (the exchange buffers are idle now)
      __syncthreads();
      if (tig == 0) red[nh * kUS + ul] = bsum;
      __syncthreads();
      if (nh == 0 && tig == 0) {
        float4 b = red[ul];
#pragma unroll
        for (int w2 = 1; w2 < NWG; ++w2) {
          const float4 o = red[w2 * kUS + ul];
          b.x += o.x; b.y += o.y; b.z += o.z; b.w += o.w;
        }
        const int ntiles = (int)gridDim.x / C;
        float* dst = p.bias_partial + (((size_t)blockIdx.z * p.G + g) * ntiles + tile) * 4 * H + 4 * u;
        *reinterpret_cast<float4*>(dst) = b;
      }
    }
    // (2) zero tail rows [T, tail_end) of my sequences in my 128 gate columns (both planes): the TN GEMM reads whole 64-row boxes
    const int tail_end = min(Tmax, ((T + 63) / 64) * 64 + 1), ntail = tail_end - T;
    constexpr int kChunks = kRows * 2 / 16;  // 16-byte chunks of 128 bf16
    float* const Gbase = dir ? p.gates[1] : p.gates[0];
    for (int i = tid; i < nvalid * ntail * kChunks * 2; i += 32 * kWarps) {
      const int cc = i % kChunks, pl = (i / kChunks) & 1, r = (i / (2 * kChunks)) % ntail, q = i / (2 * kChunks * ntail);
      unsigned char* row = reinterpret_cast<unsigned char*>(Gbase + ((size_t)(nbase + q) * Tmax + T + r) * 4 * H);
      *reinterpret_cast<uint4*>(row + (size_t)pl * 4 * H * 2 + (size_t)kRows * rank * 2 + cc * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

size_t fwd_smem(int H, int ntile, bool split) {
  const int npart = split ? 2 : 1, nsp = 8 * ntile + kNPad;
  return (size_t)npart * kRows * (H + kWPad) * 2 + (size_t)2 * npart * H * nsp * 2 + 64;  // + hbar[2][NWG]
}
size_t bwd_smem(int H, int ntile, bool split) {
  const int npart = split ? 2 : 1, ns = 8 * ntile, nsp = ns + kNPad;
  return (size_t)npart * kRows * (H + kWPad) * 2 + (size_t)npart * kRows * nsp * 2 + (size_t)2 * (H / kUS) * kUS * ns * 4 + 64;  // + xbar[2][NWG]
}

// sequences per cluster: as many n-tiles as the batch fills and shared memory allows (4 n-tiles run with two warp groups)
int pick_ntile(int B, int H, bool split, bool bwd) {
  int nt = B >= 24 ? 4 : (B >= 12 ? 2 : 1);
  while (nt > 1 && (bwd ? bwd_smem(H, nt, split) : fwd_smem(H, nt, split)) > 220 * 1024) nt >>= 1;
  return nt;
}

template <typename Kern, typename Args>
cudaError_t launch_cluster(Kern kern, const Args& a, int H, int ntile, int nwg, size_t smem, int ndir, cudaStream_t st) {
  const int C = H / kUS, NS = 8 * ntile;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(C * ((a.B + NS - 1) / NS)), (unsigned)a.G, (unsigned)ndir);
  cfg.blockDim = dim3(32 * kUnitWarps * nwg);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, a, H);
}

}  // namespace

bool lstm_cluster_supports(int H) { return H % kUS == 0 && H >= kUS && H <= 8 * kUS; }

// bias partial rows written per direction by launch_lstm_bwd_cluster in planes mode: one per (group, sequence tile)
int lstm_bwd_cluster_cta_count(const LstmBwdArgs& a, int H, int precision) {
  const int nt = pick_ntile(a.B, H, precision == 0, true);
  return a.G * ((a.B + 8 * nt - 1) / (8 * nt));
}

cudaError_t launch_lstm_fwd_cluster(const LstmFwdArgs& a, int H, int precision, cudaStream_t st) {
  if (!lstm_cluster_supports(H)) return cudaErrorInvalidValue;
  const bool split = precision == 0;
  const int nt = pick_ntile(a.B, H, split, false);
  const size_t smem = fwd_smem(H, nt, split);
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
#define IB200_FWD_CL(NT_, SP_, NWG_) return launch_cluster(lstm_fwd_cl_kernel<NT_, SP_, NWG_>, a, H, NT_, NWG_, smem, a.ndir, st)
  if (split) {
    if (nt == 4) IB200_FWD_CL(4, true, 2);
    if (nt == 2) IB200_FWD_CL(2, true, 1);
    IB200_FWD_CL(1, true, 1);
  }
  if (nt == 4) IB200_FWD_CL(4, false, 2);
  if (nt == 2) IB200_FWD_CL(2, false, 1);
  IB200_FWD_CL(1, false, 1);
#undef IB200_FWD_CL
}

cudaError_t launch_lstm_bwd_cluster(const LstmBwdArgs& a, int H, int precision, cudaStream_t st) {
  if (!lstm_cluster_supports(H)) return cudaErrorInvalidValue;
  const bool split = precision == 0;
  const int nt = pick_ntile(a.B, H, split, true);
  const size_t smem = bwd_smem(H, nt, split);
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
#define IB200_BWD_CL(NT_, SP_, NWG_) return launch_cluster(lstm_bwd_cl_kernel<NT_, SP_, NWG_>, a, H, NT_, NWG_, smem, a.ndir, st)
  if (split) {
    if (nt == 4) IB200_BWD_CL(4, true, 2);
    if (nt == 2) IB200_BWD_CL(2, true, 1);
    IB200_BWD_CL(1, true, 1);
  }
  if (nt == 4) IB200_BWD_CL(4, false, 2);
  if (nt == 2) IB200_BWD_CL(2, false, 1);
  IB200_BWD_CL(1, false, 1);
#undef IB200_BWD_CL
}

}  // namespace ib200
