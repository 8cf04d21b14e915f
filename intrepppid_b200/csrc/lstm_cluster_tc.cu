// K2t / K3t: recurrent forward and BPTT kernels for H = 128 / 256 on the 5th-generation tensor cores (tcgen05 + TMEM), e.g. the stress
// configuration E = H = 256, 3 layers (BASELINE config 5).  Same math and the same HBM layouts as lstm_cluster.cu / lstm_fwd.cu
// (reference: nn.LSTM inside encoders/awd_lstm.py:35-41,56); the mma.sync cluster kernels stay for the other hidden sizes.
//
// A thread-block CLUSTER of C = H/32 CTAs shares one tile of NS = 32 or 40 sequences.  CTA r owns hidden units [32r, 32r+32): its 128
// gate rows of W_hh stay resident in TENSOR MEMORY for the whole kernel as bf16 hi (+ lo) -- the A operand of TS-mode MMAs.  Per step
// and CTA the product is one accumulator tile in TMEM:
//   forward : gates^T[128 rows, NS] = Wslice[128, H] h_{t-1}[H, NS]          (B = h tile in shared memory, K = H)
//   backward: dh^T[H units, NS]     = Wslice^T[H, 128] da_t[128, NS]         (B = da tile in shared memory, K = 128; H/128 accumulators)
// issued from warp-uniform code by one elected lane as tcgen05.mma M=128, N=NS (3 MMAs per product in fp32 mode: hi*hi + hi*lo +
// lo*hi).  Such an MMA costs ~50 cycles whatever N <= 64 is (tools/microbench_cl.cu, profiles/r2_microbench_cl.txt): 0.4 us (bf16
// mode) / 1.2 us (fp32 mode) of tensor time per step, against ~3.3 us of HMMA + ldmatrix time in the mma.sync kernels.
//   * forward: eight cell warps read the accumulator with the 16x256b TMEM load shape -- the gate rows are ordered so that every
//     thread receives i, f, g, o of whole cells (the mma.sync fragment), no lane exchange --, add the input projection (float4 per
//     cell, register-prefetched one step ahead), update c / h and write their slice of h_t (bf16 hi | lo) directly in the B-operand
//     layout of the next step.  One warp sends the slice to the other CTAs in ring order, ONE cp.async.bulk shared::cta ->
//     shared::cluster per destination, counted (complete_tx) on a per-source mbarrier of the receiver: no cluster barrier in the
//     loop, the receiving tensor core reads what the async proxy wrote, and its MMAs start on the slices that have landed.
//   * backward: the cell warps form da_t from factors precomputed in the shadow of the previous step's MMAs, store the dgates and
//     write the bf16 da tile; the MMA result -- partial sums for ALL H units -- is reduce-scattered: the TMEM rows of accumulator a,
//     lane quarter q belong to CTA 4a + q; they are staged in the idle receive buffer and leave as one bulk copy per destination
//     (fp32 partials in fp32 mode, bf16 in bf16 mode), counted on the owner's mbarrier; the owner adds the C partials.
// DESIGN.md section 5 (K2t / K3t) lists the measurements behind each of these choices.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <type_traits>
#include <utility>

#include "kernels.h"
#include "tc05.cuh"

namespace ib200 {
namespace {

using namespace tc;

// IB200_PROF builds (timing experiments only): per-phase cycle counters of cluster 0, printed at the end of the kernel
#ifdef IB200_PROF
#define PROF_DECL long long pf_[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, pt_ = clock64()
#define PROF_MARK(i) do { const long long now_ = clock64(); pf_[i] += now_ - pt_; pt_ = now_; } while (0)
#define PROF_PRINT(tag, cond, T) do { if (cond) printf("%s T=%d cycles/step: %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld\n", tag, T, pf_[0] / T, pf_[1] / T, pf_[2] / T, pf_[3] / T, pf_[4] / T, pf_[5] / T, pf_[6] / T, pf_[7] / T, pf_[8] / T, pf_[9] / T); } while (0)
#else
#define PROF_DECL do { } while (0)
#define PROF_MARK(i) do { } while (0)
#define PROF_PRINT(tag, cond, T) do { } while (0)
#endif

constexpr int kUS = 32;                 // hidden units per CTA
constexpr int kRows = 4 * kUS;          // gate rows per CTA = UMMA M
// sequences per cluster = UMMA N = 8 * NB (NB = 4 or 5 blocks of eight): 15 clusters of 8 CTAs are co-resident on a B200
// (tools/probe_clusters.cu), so 256 sequences x 2 directions in tiles of 32 (16 clusters) would run in TWO waves; tiles of 40
// (14 clusters) run in one
constexpr int kCellWarps = 8;           // warps 0..7: TMEM read-out + cell math; warp 8: MMA issuer and TMEM owner
constexpr int kCellThreads = 32 * kCellWarps;
// 12 warps = three warpgroups: two of cell warps, one with the MMA issuer (+ three idle warps that only complete the warpgroup).  A CTA
// of 9..12 warps gets 168 registers per thread at launch (three warps per scheduler partition); the third warpgroup hands its
// registers back (setmaxnreg.dec) and the cell warps grow to 208 (setmaxnreg.inc): the cell code keeps its prefetched inputs and the
// step's state in registers without spilling (a spilled prefetch register turns the prefetch into a blocking load).
constexpr int kThreads = kCellThreads + 128;
constexpr int kCellRegs = 208, kAuxRegs = 64;
template <int NB>
struct TileT {
  static constexpr int NS = 8 * NB;
  static constexpr int kChunk = NS * 16 + 16;   // byte pitch of one 8-row k chunk [NS sequences][16 B] of a B operand tile (+16: bank spread)
  static constexpr int kHSlice = 4 * kChunk;    // forward: h slice of one source CTA and one part: [4 unit chunks][NS sequences][16 B]
  // backward: partial dh of one source CTA for my 32 units: [NS sequences][32 units], fp32 (fp32 mode) or bf16 (bf16 mode: half the
  // exchange bytes; the partials are rounded to bf16 before the C-way sum, inside that mode's 2e-2 gate)
  static constexpr int xslice(int elem_bytes) { return NS * kUS * elem_bytes; }
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(local_saddr), "r"(rank));
  return r;
}
// shared-memory matrix descriptor without swizzle (K-major "interleaved" canonical layout): core matrices of 8 rows x 16 bytes
// (128 contiguous bytes); sbo = bytes between 8-row groups along M/N, lbo = bytes between the two 16-byte chunks along K
__device__ __forceinline__ uint64_t smem_desc_nosw(uint32_t saddr, uint32_t sbo_bytes, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (sm_100); layout type 0 = no swizzle
  return d;
}
// 32 lanes x 4 consecutive 32-bit columns -> 4 registers per thread; no wait
__device__ __forceinline__ void tmem_ld4_nowait(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
// 32 lanes x 8 consecutive 32-bit columns -> 8 registers per thread (thread i of the warp reads lane base+i); no wait
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// 16 lanes x 256 bits (8 columns): thread t of the warp gets rows t/4 and t/4 + 8 of the 16-lane window at columns 2(t%4), +1 -- the
// mma.sync accumulator fragment: r[2*(row half) + column parity]; no wait
__device__ __forceinline__ void tmem_ld_16x256b_x1(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
// (mma_bf16_ss_elect / mma_commit_elect: tc05.cuh)
// the same with the A operand in TMEM (lane = M row, 8 columns of packed bf16 pairs per k16 step)
__device__ __forceinline__ void mma_bf16_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n .reg .pred p, e;\n elect.sync _|e, 0xffffffff;\n setp.ne.b32 p, %4, 0;\n"
      " @e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// 32 lanes x 8 consecutive 32-bit columns <- 8 registers per thread (thread i of the warp writes lane base+i)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx_elect(uint64_t* bar, uint32_t bytes) {
  asm volatile(
      "{\n .reg .pred e;\n elect.sync _|e, 0xffffffff;\n"
      " @e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n}\n" ::"r"(smem_u32(bar)), "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void bulk_s2c_elect(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile(
      "{\n .reg .pred e;\n elect.sync _|e, 0xffffffff;\n"
      " @e cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n}\n" ::"r"(dst_cluster),
      "r"(src_cta), "r"(bytes), "r"(bar_cluster)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_elect(uint64_t* bar) {
  asm volatile("{\n .reg .pred e;\n elect.sync _|e, 0xffffffff;\n @e mbarrier.arrive.shared::cta.b64 _, [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pair_bar_sync(int id) { asm volatile("bar.sync %0, 64;\n" ::"r"(id) : "memory"); }
template <int REGS>
__device__ __forceinline__ void reg_grow() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(REGS)); }
template <int REGS>
__device__ __forceinline__ void reg_shrink() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(REGS)); }
__device__ __forceinline__ void cell_bar_sync() { asm volatile("bar.sync 1, %0;\n" ::"n"(kCellThreads) : "memory"); }

// The resident W slice is the A operand of every MMA and lives in TENSOR MEMORY (lane = M row, one 32-bit column = two consecutive
// k as packed bf16; hi part, then lo part): the 48 MMAs of a step then read only their small B tiles from shared memory.  With the
// slice in shared memory (128 KB, 4 KB per MMA) the operand reads shared the memory pipe with the DSMEM exchange and the cell warps
// and an MMA took 68-85 cycles in the kernel against 50 in isolation.
//   forward : M = gate row L = 32*(unit/8) + 8*gate + unit%8 (rows q, q+8, q+16, q+24 of a 32-lane TMEM quarter are i,f,g,o of one
//             unit), K = all H units: H/2 columns per part
//   backward: M = unit (accumulator a holds units [128a, 128a+128)), K = my 128 gate rows L = 4*local unit + gate (the four dgates of
//             a cell are consecutive k of the da tile): 64 columns per part and accumulator
// Row L of the slice is W_hh[gate*H + 32*rank + local unit][.] (masked per group for layer 0, forward direction).  Called by the eight
// cell warps: warp (quarter, half) fills its TMEM lane quarter, `half` selects the column half.
template <bool SPLIT>
__device__ __forceinline__ void load_w_tmem_fwd(uint32_t tb, const float* __restrict__ W, const float* __restrict__ M, int H, int rank) {
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, quarter = wid & 3, half = wid >> 2;
  const int L = 32 * quarter + lane, gate = (L >> 3) & 3, ul = 8 * (L >> 5) + (L & 7);
  const size_t row = ((size_t)gate * H + kUS * rank + ul) * H;
  const int KC = H / 2, c_begin = half * (KC / 2), c_end = c_begin + KC / 2;  // packed columns of this warp
  const uint32_t lane_addr = tb + ((uint32_t)(32 * quarter) << 16);
  for (int c0 = c_begin; c0 < c_end; c0 += 8) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const size_t src = row + 2 * (c0 + j);
      float w0 = W[src], w1 = W[src + 1];
      if (M != nullptr) {
        w0 *= M[src];
        w1 *= M[src + 1];
      }
      lo[j] = 0u;
      if constexpr (SPLIT) split_bf16(w0, w1, hi[j], lo[j]);
      else hi[j] = pack_bf16(w0, w1);
    }
    tmem_st8(lane_addr + (uint32_t)c0, hi);
    if constexpr (SPLIT) tmem_st8(lane_addr + (uint32_t)(KC + c0), lo);
  }
  tmem_wait_st();
}
template <bool SPLIT>
__device__ __forceinline__ void load_w_tmem_bwd(uint32_t tb, const float* __restrict__ W, const float* __restrict__ M, int H, int rank) {
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, quarter = wid & 3, half = wid >> 2;
  const int NACC = H / 128;
  const uint32_t lane_addr = tb + ((uint32_t)(32 * quarter) << 16);
  for (int a = 0; a < NACC; ++a) {
    const int unit = 128 * a + 32 * quarter + lane;  // M row of this thread = a column of W_hh
    for (int c0 = 32 * half; c0 < 32 * half + 32; c0 += 8) {  // packed columns: k = gate rows 2c, 2c+1
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int L0 = 2 * (c0 + j);  // L0 even: gates (0,1) or (2,3) of local unit L0 / 4
        const size_t s0 = ((size_t)(L0 & 3) * H + kUS * rank + (L0 >> 2)) * H + unit, s1 = s0 + (size_t)H * H;  // gate + 1: H rows further
        float w0 = W[s0], w1 = W[s1];
        if (M != nullptr) {
          w0 *= M[s0];
          w1 *= M[s1];
        }
        lo[j] = 0u;
        if constexpr (SPLIT) split_bf16(w0, w1, hi[j], lo[j]);
        else hi[j] = pack_bf16(w0, w1);
      }
      tmem_st8(lane_addr + (uint32_t)(a * 64 + c0), hi);
      if constexpr (SPLIT) tmem_st8(lane_addr + (uint32_t)(NACC * 64 + a * 64 + c0), lo);
    }
  }
  tmem_wait_st();
}

// =================================================================================================================================
// forward
// =================================================================================================================================
template <int NB, bool SPLIT, bool TRAIN>
__global__ void __launch_bounds__(kThreads, 1) lstm_fwd_cltc_kernel(const LstmFwdArgs p, const int H) {
  using TT = TileT<NB>;
  // cells per thread: warp `half` takes NFB whole 8-sequence blocks and, for odd NB, one column parity of the middle block, so both
  // halves carry the same number of cells (NB = 5: 5 + 5 instead of 6 + 4: the publish waits for the slower half)
  constexpr int NS = TT::NS, NFB = NB / 2, NLD = NFB + (NB & 1), NCELL = 2 * NFB + (NB & 1), NPART = SPLIT ? 2 : 1;
  constexpr int kChunk = TT::kChunk, kHSlice = TT::kHSlice;
  constexpr uint32_t kTmemCols = 512;  // W slice: H/2 columns per part (<= 256), accumulator at column 256
  constexpr uint32_t kDCol = 256;
  constexpr bool FAST = !SPLIT;
  const int C = H / kUS;
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_ctarank(), tile = (int)blockIdx.x / C;
  const int g = blockIdx.y, dir = p.dir0 + (int)blockIdx.z;
  const int T = p.lens[p.G + g];
  if (T <= 0) return;  // uniform over the cluster
  const int b0 = tile * NS, nvalid = min(NS, p.B - b0), nbase = g * p.B + b0, Tmax = p.Tmax;
  const bool layer0 = p.tok != nullptr;

  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* hB = smem;                                    // [2 buffers][C source CTAs][NPART][kHSlice]: B operand of the step
  const uint32_t sliceBytes = NPART * kHSlice, bufBytes = (uint32_t)C * sliceBytes;
  // hbar[buffer][source CTA]: "the slice of h that CTA `source` owns has landed in buffer b" -- the MMAs of a step start on the
  // slices that are there (my own first) while the others are still in flight
  uint64_t* hbar = reinterpret_cast<uint64_t*>(hB + 2 * (size_t)bufBytes);
  uint64_t* mma_bar = hbar + 2 * C;                            // "the step's MMAs have completed"
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);

  for (int i = tid; i < (int)(2 * bufBytes / 4); i += kThreads) reinterpret_cast<uint32_t*>(hB)[i] = 0u;  // h_{-1} = 0
  if (tid == 0) {
    for (int i = 0; i < 2 * C; ++i) mbar_init(&hbar[i], 1);
    mbar_init(mma_bar, 1);
    mbar_init_fence();
    for (int r = 0; r < C; ++r) {
      if (r == rank) continue;  // (my own slice: a plain arrive of the publishing warp)
      if (T > 1) mbar_arrive_expect_tx(&hbar[C + r], sliceBytes);  // h_0 -> buffer 1
      if (T > 2) mbar_arrive_expect_tx(&hbar[r], sliceBytes);      // h_1 -> buffer 0
    }
  }
  if (wid == kCellWarps) tmem_alloc(tmem_slot, kTmemCols);
  fence_async_smem();  // the generic-proxy writes above (zero h tile) are operands of the tensor core (async proxy)
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = *tmem_slot;
  if (wid < kCellWarps) {
    const float* __restrict__ W = dir ? p.whh[1] : p.whh[0];
    const float* __restrict__ M = (dir == 0 && p.whh_mask != nullptr) ? p.whh_mask + (size_t)g * 4 * H * H : nullptr;
    load_w_tmem_fwd<SPLIT>(tb, W, M, H, rank);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  cluster_sync_all();  // every CTA's tiles and barriers exist before any remote copy lands

  if (wid >= kCellWarps) {
    reg_shrink<kAuxRegs>();
  }
  if (wid > kCellWarps) {
    // (idle warps of the third warpgroup)
  } else if (wid == kCellWarps) {
    // ===================== MMA issuer (whole warp, uniform control flow; one elected lane issues) =====================
    constexpr uint32_t idesc = idesc_bf16(kRows, NS, false, false);  // A = W slice in TMEM (M = gate rows), B = h tile K-major
    const uint64_t b_base = smem_desc_nosw(smem_u32(hB), 128, kChunk);
    const uint64_t b_lo = (uint64_t)(kHSlice >> 4);
    const uint32_t a_lo = (uint32_t)(H / 2);
    PROF_DECL;
    for (int s = 0; s < T; ++s) {
      const int buf = s & 1;
      const uint32_t par = (uint32_t)(((s - 1) >> 1) & 1);
      const uint64_t bb = b_base + (uint64_t)(((uint32_t)buf * bufBytes) >> 4);
      for (int i = 0; i < C; ++i) {
        const int r = rank - i >= 0 ? rank - i : rank - i + C;  // source CTA: mine first, then rank-1, rank-2, ... (arrival order)
        PROF_MARK(1);
        if (s > 0) {
          mbar_wait(&hbar[buf * C + r], par);
          if (r != rank && s + 2 < T) mbar_arrive_expect_tx_elect(&hbar[buf * C + r], sliceBytes);  // next fill: h_{s+1}
        }
        PROF_MARK(i == 0 ? 0 : (i == 1 ? 2 : 3));  // wait for: my own slice | the first remote slice | the others
        if (i == 0) fence_after_sync();  // (the cell warps' tcgen05.ld of the previous step precede these MMAs)
        // units [32 r, +32) = two k16 steps: A = 8 TMEM columns each; B = the four unit chunks of source r
        const uint32_t ah = tb + (uint32_t)(16 * r);
        const uint64_t bh = bb + (uint64_t)(((uint32_t)r * sliceBytes) >> 4);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint32_t ahj = ah + (uint32_t)(8 * j);
          const uint64_t bhj = bh + (uint64_t)(j * ((2 * kChunk) >> 4));
          mma_bf16_ts_elect(tb + kDCol, ahj, bhj, idesc, (i | j) != 0);
          if constexpr (SPLIT) {
            mma_bf16_ts_elect(tb + kDCol, ahj, bhj + b_lo, idesc, true);
            mma_bf16_ts_elect(tb + kDCol, ahj + a_lo, bhj, idesc, true);
          }
        }
      }
      mma_commit_elect(mma_bar);
      PROF_MARK(1);
    }
    PROF_PRINT("fwd mma  [own-wait issue first-remote-wait other-remote-waits]", blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0, T);
  } else {
    // ===================== cell warps =====================
    // accumulator read-out: warp (quarter, half) reads TMEM lanes [32 quarter, +32) with the 16x256b shape, one 8-column block at a
    // time: rows gq, gq + 8 (first 16 lanes) and gq + 16, gq + 24 (second 16 lanes) = gates i, f, g, o of unit 8 quarter + gq, for
    // the sequences 8 b + 2 tig + (0, 1): whole cells per thread, no exchange between lanes
    reg_grow<kCellRegs>();
    const int quarter = wid & 3, half = wid >> 2, gq = lane >> 2, tig = lane & 3;
    const int u = kUS * rank + 8 * quarter + gq;
    const int t_first = dir ? T - 1 : 0, dt = dir ? -1 : 1;
    const int pad_row = p.V + (int)((tile + gridDim.x / C * (blockIdx.y + gridDim.y * blockIdx.z)) % kPadRows);
    int ncell[NCELL], rowb[NCELL];  // sequence of cell e inside the tile; its first token row (beyond the batch: clamped, never stored)
    uint32_t xoff[NCELL];
    bool valid[NCELL];
#pragma unroll
    for (int e = 0; e < NCELL; ++e) {
      const int b = e < 2 * NFB ? half * (NB - NFB) + (e >> 1) : NFB;  // whole blocks of this half | the shared middle block
      ncell[e] = 8 * b + 2 * tig + (e < 2 * NFB ? (e & 1) : half);
      valid[e] = ncell[e] < nvalid;
      const int ncl = min(ncell[e], nvalid - 1);
      rowb[e] = (nbase + ncl) * Tmax;
      xoff[e] = (uint32_t)ncl * (uint32_t)Tmax * (uint32_t)H;  // float4 units, relative to the tile's first sequence
    }
    // layer >= 1: dense input projection rows [N, Tmax, 4H] (GI: one float4 per cell); layer 0: table rows by token
    const float4* const xdense = layer0 ? nullptr : reinterpret_cast<const float4*>(dir ? p.xproj[1] : p.xproj[0]) + (size_t)nbase * Tmax * H + u;
    const float4* const xtab = layer0 ? reinterpret_cast<const float4*>(p.table) + (size_t)((p.table_shared ? 0 : g) * 2 + dir) * (p.V + kPadRows) * H + u : nullptr;
    auto load_tok = [&](int s, int (&tk)[NCELL]) {
#pragma unroll
      for (int e = 0; e < NCELL; ++e) tk[e] = __ldg(p.tok + (size_t)rowb[e] + t_first + s * dt);
    };
    auto load_x = [&](int s, const int (&tk)[NCELL], float4 (&x)[NCELL]) {  // (layer 0: tk = the tokens of step s)
      if (layer0) {
#pragma unroll
        for (int e = 0; e < NCELL; ++e) x[e] = __ldg(xtab + (size_t)(tk[e] == 0 ? pad_row : tk[e]) * H);  // pads: this cluster's copy of row 0
      } else {
        const float4* xt = xdense + (ptrdiff_t)(t_first + s * dt) * H;
#pragma unroll
        for (int e = 0; e < NCELL; ++e) x[e] = __ldg(xt + xoff[e]);
      }
    };

    float4* const G4 = TRAIN ? reinterpret_cast<float4*>(dir ? p.gates[1] : p.gates[0]) + u : nullptr;
    float* const Cst = TRAIN ? (dir ? p.cstate[1] : p.cstate[0]) + u : nullptr;
    const bool has_y = p.y != nullptr;
    const int ycol = dir * H + u, ystr = p.y_stride;
    float cst[NCELL], hv[NCELL];
#pragma unroll
    for (int e = 0; e < NCELL; ++e) cst[e] = hv[e] = 0.f;
    const uint32_t my_slot = smem_u32(hB) + (uint32_t)rank * sliceBytes;  // + buffer offset: my slice of the h tile
    unsigned char* const my_slot_ptr = hB + (size_t)rank * sliceBytes;
    const uint32_t taddr = tb + ((uint32_t)(32 * quarter) << 16) + kDCol;

    // input projection of the current step (register prefetch: the next step's loads are issued right after this step's values have
    // been consumed and the h slice is on its way; tokens two steps ahead)
    float4 xc[NCELL];
    int tk[NCELL];
#pragma unroll
    for (int e = 0; e < NCELL; ++e) tk[e] = 0;
    if (layer0) load_tok(0, tk);
    load_x(0, tk, xc);
    if (layer0 && T > 1) load_tok(1, tk);
    PROF_DECL;
    for (int s = 0; s < T; ++s) {
      const int buf = s & 1;
      PROF_MARK(0);
      mbar_wait(mma_bar, (uint32_t)(s & 1));
      PROF_MARK(1);
      fence_after_sync();
      uint32_t ra[NLD][4], rb[NLD][4];
#pragma unroll
      for (int jb = 0; jb < NLD; ++jb) {
        const uint32_t col = (uint32_t)(8 * (jb < NFB ? half * (NB - NFB) + jb : NFB));
        tmem_ld_16x256b_x1(taddr + col, ra[jb]);
        tmem_ld_16x256b_x1(taddr + col + (16u << 16), rb[jb]);
      }
      tmem_wait_ld();
      fence_before_sync();
      PROF_MARK(2);

      const int t = t_first + s * dt;
      float gi[NCELL], gf[NCELL], gg[NCELL], go[NCELL];
      __nv_bfloat16 hb16[NCELL], lb16[NCELL];
#pragma unroll
      for (int e = 0; e < NCELL; ++e) {
        uint32_t ai, af, ag, ao;
        if (e < 2 * NFB) {
          ai = ra[e >> 1][e & 1]; af = ra[e >> 1][2 + (e & 1)]; ag = rb[e >> 1][e & 1]; ao = rb[e >> 1][2 + (e & 1)];
        } else {  // the middle block: this half's column parity
          ai = half ? ra[NLD - 1][1] : ra[NLD - 1][0]; af = half ? ra[NLD - 1][3] : ra[NLD - 1][2];
          ag = half ? rb[NLD - 1][1] : rb[NLD - 1][0]; ao = half ? rb[NLD - 1][3] : rb[NLD - 1][2];
        }
        gi[e] = sigmoid_f<FAST>(__uint_as_float(ai) + xc[e].x);
        gf[e] = sigmoid_f<FAST>(__uint_as_float(af) + xc[e].y);
        gg[e] = tanh_f<FAST>(__uint_as_float(ag) + xc[e].z);
        go[e] = sigmoid_f<FAST>(__uint_as_float(ao) + xc[e].w);
        cst[e] = fmaf(gf[e], cst[e], gi[e] * gg[e]);
        hv[e] = go[e] * tanh_f<FAST>(cst[e]);
        hb16[e] = __float2bfloat16_rn(hv[e]);
        lb16[e] = __float2bfloat16_rn(hv[e] - __bfloat162float(hb16[e]));
      }
      if (s + 1 < T) {  // my elements of the next step's B operand: [part][unit chunk = quarter][sequence][8 units x bf16]
        unsigned char* dst = my_slot_ptr + (size_t)(buf ^ 1) * bufBytes + quarter * kChunk + gq * 2;
#pragma unroll
        for (int e = 0; e < NCELL; ++e) {
          *reinterpret_cast<__nv_bfloat16*>(dst + ncell[e] * 16) = hb16[e];
          if constexpr (SPLIT) *reinterpret_cast<__nv_bfloat16*>(dst + kHSlice + ncell[e] * 16) = lb16[e];
        }
        PROF_MARK(3);
        fence_async_smem();  // my slice (generic-proxy stores) is read by the bulk copies and by my own tensor core (async proxy)
        PROF_MARK(4);
        cell_bar_sync();
        PROF_MARK(5);
        if (wid == 0) {
          // one warp issues the copies in RING order (to rank+1 first, rank+2 next, ...): the copy engine works through them roughly
          // in order, so at every receiver the slices arrive staggered (from rank-1 first) and its MMAs start on the early ones while
          // the late ones are still in flight; eight warps issuing side by side made all slices land together at the end
          const uint32_t boff = (uint32_t)(buf ^ 1) * bufBytes;
          uint64_t* bar = &hbar[(buf ^ 1) * C + rank];  // slot `rank` of the receiver's barriers
          mbar_arrive_elect(bar);
          for (int i = 1; i < C; ++i) {
            const uint32_t d = (uint32_t)(rank + i < C ? rank + i : rank + i - C);
            bulk_s2c_elect(map_to_rank(my_slot + boff, d), my_slot + boff, sliceBytes, map_to_rank(smem_u32(bar), d));
          }
        }
        PROF_MARK(6);
        load_x(s + 1, tk, xc);
        if (layer0 && s + 2 < T) load_tok(s + 2, tk);
      }
#pragma unroll
      for (int e = 0; e < NCELL; ++e) {
        if (valid[e]) {
          const size_t row = (size_t)(rowb[e] + t);
          if constexpr (TRAIN) {
            G4[row * H] = make_float4(gi[e], gf[e], gg[e], go[e]);
            Cst[row * H] = cst[e];
          }
          if (has_y) {  // bf16 hi | lo planes over the bytes of the fp32 row (operands of the TMA-fed GEMMs, gemm_wide.cu)
            __nv_bfloat16* yrow = reinterpret_cast<__nv_bfloat16*>(p.y + row * ystr);
            yrow[ycol] = hb16[e];
            if constexpr (SPLIT) yrow[ystr + ycol] = lb16[e];
          }
        }
      }
      PROF_MARK(7);
    }
    PROF_PRINT("fwd cell [loop wait-mma tmem-ld math+sts fence bar bulk-issue stores]", blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (tid == 0 || tid == 255), T);

    if (p.hn != nullptr) {
      const size_t N = (size_t)p.G * p.B;
#pragma unroll
      for (int e = 0; e < NCELL; ++e)
        if (valid[e]) p.hn[((size_t)dir * N + nbase + ncell[e]) * H + u] = hv[e];
    }
  }
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // nobody exits while a peer's copy could still read from or write to it
  if (wid == kCellWarps) tmem_dealloc(tb, kTmemCols);

  // planes mode: the weight-gradient GEMM reads whole 64-row TMA boxes (and the row after the last one for the shifted operand): rows
  // [T, tail_end) of this cluster's sequences must be zeros in this CTA's 32 columns of this direction (both planes)
  if (p.y != nullptr) {
    const int tail_end = min(Tmax, ((T + 63) / 64) * 64 + 1), ntail = tail_end - T;
    constexpr int kChunks = kUS * 2 / 16;  // 16-byte chunks of 32 bf16
    for (int i = tid; i < nvalid * ntail * kChunks * 2; i += kThreads) {
      const int cc = i % kChunks, pl = (i / kChunks) & 1, rr = (i / (2 * kChunks)) % ntail, qq = i / (2 * kChunks * ntail);
      unsigned char* row = reinterpret_cast<unsigned char*>(p.y + ((size_t)(nbase + qq) * Tmax + T + rr) * p.y_stride);
      *reinterpret_cast<uint4*>(row + pl * p.y_stride * 2 + (dir * H + kUS * rank) * 2 + cc * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

// =================================================================================================================================
// backward
// =================================================================================================================================
template <int NB, bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1) lstm_bwd_cltc_kernel(const LstmBwdArgs p, const int H) {
  using TT = TileT<NB>;
  using XT = typename std::conditional<SPLIT, float, __nv_bfloat16>::type;  // element type of the exchanged partial sums
  constexpr int NS = TT::NS, NPART = SPLIT ? 2 : 1;
  constexpr int kChunk = TT::kChunk, kXSlice = TT::xslice((int)sizeof(XT));
  constexpr int NG8 = (NS / 2) / 8, NG4 = ((NS / 2) % 8) / 4;  // read-out per warp: NS/2 columns = NG8 groups of 8 (+ one of 4)
  constexpr bool FAST = !SPLIT;
  const int C = H / kUS, NACC = H / 128;
  constexpr uint32_t kTmemCols = 512;  // W slice: 64 columns per part and accumulator (<= 256), accumulators from column 256
  constexpr uint32_t kDCol = 256;
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_ctarank(), tile = (int)blockIdx.x / C;
  const int g = blockIdx.y, dir = p.dir0 + (int)blockIdx.z;
  const int T = p.lens[p.G + g];
  if (T <= 0) return;
  const int b0 = tile * NS, nvalid = min(NS, p.B - b0), nbase = g * p.B + b0, Tmax = p.Tmax;
  const size_t N = (size_t)p.G * p.B;
  const bool has_dy = p.dy != nullptr, planes = p.planes != 0;

  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* daB = smem;                                   // [NPART][16 k chunks][kChunk]: B operand (da of my 128 gate rows)
  constexpr uint32_t kDaPart = 16 * kChunk;
  // partial dh for my 32 units: xbuf[2 buffers][C - 1 remote CTAs][NS sequences][32 units] fp32 (slot of CTA r: (r - rank - 1) mod C)
  // + xown[NS][32]: my own partial.  The idle receive buffer doubles as the staging area of the outgoing slices (see below).
  XT* xbuf = reinterpret_cast<XT*>(daB + NPART * kDaPart);
  const uint32_t xBufBytes = (uint32_t)(C - 1) * kXSlice;
  constexpr int kXElems = NS * kUS;  // elements per slice
  XT* xown = xbuf + 2 * (size_t)(C - 1) * kXElems;
  // xbar[buffer]: "all partials of this step are in place": C - 1 bulk copies (complete_tx) + the two warps that wrote xown
  uint64_t* xbar = reinterpret_cast<uint64_t*>(xown + kXElems);
  uint64_t* da_bar = xbar + 2;   // "the da tile of this step is complete" (all cell threads arrive)
  uint64_t* mma_bar = xbar + 3;  // [2]: "the step's MMAs into accumulator a have completed" (accumulator 0 is sent while 1 is computed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xbar + 5);

  for (int i = tid; i < (int)(NPART * kDaPart / 4); i += kThreads) reinterpret_cast<uint32_t*>(daB)[i] = 0u;
  if (tid == 0) {
    mbar_init(&xbar[0], 3);
    mbar_init(&xbar[1], 3);
    mbar_init(da_bar, kCellThreads);
    mbar_init(&mma_bar[0], 1);
    mbar_init(&mma_bar[1], 1);
    mbar_init_fence();
    if (T > 1) mbar_arrive_expect_tx(&xbar[1], xBufBytes);  // step s fills buffer (s + 1) & 1: step 0 -> buffer 1, step 1 -> buffer 0
    if (T > 2) mbar_arrive_expect_tx(&xbar[0], xBufBytes);
  }
  if (wid == kCellWarps) tmem_alloc(tmem_slot, kTmemCols);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = *tmem_slot;
  if (wid < kCellWarps) {
    const float* __restrict__ W = dir ? p.whh[1] : p.whh[0];
    const float* __restrict__ M = (dir == 0 && p.whh_mask != nullptr) ? p.whh_mask + (size_t)g * 4 * H * H : nullptr;
    load_w_tmem_bwd<SPLIT>(tb, W, M, H, rank);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  cluster_sync_all();
  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);  // column sums of my cells' dgates over all steps (bias gradient partials)

  if (wid >= kCellWarps) {
    reg_shrink<kAuxRegs>();
  }
  if (wid > kCellWarps) {
    // (idle warps of the third warpgroup)
  } else if (wid == kCellWarps) {
    // ===================== MMA issuer (whole warp, uniform control flow; one elected lane issues) =====================
    constexpr uint32_t idesc = idesc_bf16(128, NS, false, false);  // A = W slice^T in TMEM (M = units), B = da tile K-major
    const uint64_t b_base = smem_desc_nosw(smem_u32(daB), 128, kChunk);
    const uint64_t b_lo = (uint64_t)(kDaPart >> 4);
    const uint32_t a_lo = (uint32_t)(NACC * 64);
    PROF_DECL;
    for (int s = 0; s + 1 < T; ++s) {
      mbar_wait(da_bar, (uint32_t)(s & 1));
      PROF_MARK(0);
      fence_after_sync();
      for (int a = 0; a < NACC; ++a) {
        const uint32_t aa = tb + (uint32_t)(64 * a), dd = tb + kDCol + (uint32_t)a * NS;
#pragma unroll
        for (int k16 = 0; k16 < kRows / 16; ++k16) {
          // k = gate rows [16 k16, +16): A = 8 TMEM columns of accumulator a's slice; B = k chunks 2 k16, +1
          const uint32_t ah = aa + (uint32_t)(8 * k16);
          const uint64_t bh = b_base + (uint64_t)((2 * k16 * kChunk) >> 4);
          mma_bf16_ts_elect(dd, ah, bh, idesc, k16 != 0);
          if constexpr (SPLIT) {
            mma_bf16_ts_elect(dd, ah, bh + b_lo, idesc, true);
            mma_bf16_ts_elect(dd, ah + a_lo, bh, idesc, true);
          }
        }
        mma_commit_elect(&mma_bar[a]);
      }
      PROF_MARK(1);
    }
    PROF_PRINT("bwd mma  [wait-da issue]", blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0, T);
  } else {
    // ===================== cell warps: lane = local unit, warp wi owns sequences NB wi .. NB wi + NB - 1 =====================
    reg_grow<kCellRegs>();
    const int wi = wid, u = kUS * rank + lane, s0 = NB * wi;
    const int t_first = dir ? 0 : T - 1, dt = dir ? 1 : -1;  // backward scan: t = T-1..0 (forward chain) or 0..T-1 (reverse chain)
    float4* const G4 = reinterpret_cast<float4*>(dir ? p.gates[1] : p.gates[0]) + u;
    const float* const Cst = (dir ? p.cstate[1] : p.cstate[0]) + u;
    const float* const DY = has_dy ? p.dy + dir * H + u : nullptr;
    const int dy_stride = p.dy_stride;
    bool valid[NB];
    int rowb[NB];  // first token row of my cell's sequence (columns beyond the batch: clamped for loads, zeroed, never stored)
#pragma unroll
    for (int e = 0; e < NB; ++e) {
      valid[e] = s0 + e < nvalid;
      rowb[e] = (nbase + min(s0 + e, nvalid - 1)) * Tmax;
    }
    struct In {
      float4 g[NB];
      float cprev[NB], dy[NB];
    };
    auto load_in = [&](int s, In& in) {
      const int t = t_first + s * dt;
      const bool has_prev = s + 1 < T;
#pragma unroll
      for (int e = 0; e < NB; ++e) {
        const size_t row = (size_t)(rowb[e] + t);
        in.g[e] = G4[row * H];
        in.cprev[e] = has_prev ? Cst[(row + dt) * H] : 0.f;  // c of the scan predecessor; 0 at the chain start
        in.dy[e] = has_dy ? DY[row * dy_stride] : 0.f;
      }
    };
    // everything of a cell step that does not depend on the recurrent gradient is formed ahead of it (while the previous step's MMAs
    // and exchange are in flight):  dh = dhrec + dy;  dct = dh * A + dc;  da = (dct Fi, dct Ff, dct Fg, dh Fo);  dc' = dct * gf
    float fA[NB], fI[NB], fF[NB], fG[NB], fO[NB], fDy[NB], fGf[NB];
    auto prep = [&](const In& in, const float (&ccur)[NB]) {
#pragma unroll
      for (int e = 0; e < NB; ++e) {
        const bool ok = valid[e];
        const float gi = ok ? in.g[e].x : 0.f, gf = ok ? in.g[e].y : 0.f, gg = ok ? in.g[e].z : 0.f, go = ok ? in.g[e].w : 0.f;
        const float tc = tanh_f<FAST>(ccur[e]);
        fA[e] = go * fmaf(-tc, tc, 1.0f);
        fI[e] = gg * gi * (1.0f - gi);
        fF[e] = in.cprev[e] * gf * (1.0f - gf);
        fG[e] = gi * fmaf(-gg, gg, 1.0f);
        fO[e] = tc * go * (1.0f - go);
        fDy[e] = ok ? in.dy[e] : 0.f;
        fGf[e] = gf;
      }
    };
    float dc[NB], dhrec[NB];
    // my da values in the B tile: k = 4 lane + gate -> chunk lane/2, bytes (lane & 1) * 8 of the 16-byte row of sequence n
    unsigned char* const da_put = daB + (size_t)(lane >> 1) * kChunk + (size_t)(lane & 1) * 8 + (size_t)s0 * 16;
    // read-out: warp (quarter, half) reads TMEM lanes [32 quarter, +32) of every accumulator, columns [half NS/2, +NS/2); lane = unit
    // 128 a + 32 quarter + lane, which CTA 4a + quarter owns as its local unit `lane`
    const int quarter = wid & 3, half = wid >> 2;
    const int xc0 = half * (NS / 2);
    const uint32_t xb_local = smem_u32(xbuf), xbar_local = smem_u32(xbar);

    // `in` holds the saved gates / c / dy of the NEXT step (register prefetch: loaded one step ahead, consumed by prep() in the shadow
    // of this step's MMAs, reloaded right after); ccur = the cell state of the step whose factors prep() forms next
    In in;
    float ccur[NB];
    load_in(0, in);
#pragma unroll
    for (int e = 0; e < NB; ++e) {
      ccur[e] = Cst[(size_t)(rowb[e] + t_first) * H];
      dc[e] = 0.f;
      dhrec[e] = (valid[e] && p.dhn != nullptr) ? p.dhn[((size_t)dir * N + nbase + s0 + e) * H + u] : 0.f;
    }
    prep(in, ccur);
#pragma unroll
    for (int e = 0; e < NB; ++e) ccur[e] = in.cprev[e];
    if (T > 1) load_in(1, in);
    PROF_DECL;
    for (int s = 0; s < T; ++s) {
      PROF_MARK(0);
      const int t = t_first + s * dt;
      const bool more = s + 1 < T;
      uint32_t h0[NB], h1[NB], l0[NB], l1[NB];
      float da[NB][4];
#pragma unroll
      for (int e = 0; e < NB; ++e) {
        const float dh = dhrec[e] + fDy[e];
        const float dct = fmaf(dh, fA[e], dc[e]);
        dc[e] = dct * fGf[e];
        da[e][0] = dct * fI[e];
        da[e][1] = dct * fF[e];
        da[e][2] = dct * fG[e];
        da[e][3] = dh * fO[e];
        l0[e] = l1[e] = 0u;
        if constexpr (SPLIT) {
          split_bf16(da[e][0], da[e][1], h0[e], l0[e]);
          split_bf16(da[e][2], da[e][3], h1[e], l1[e]);
        } else {
          h0[e] = pack_bf16(da[e][0], da[e][1]);
          h1[e] = pack_bf16(da[e][2], da[e][3]);
        }
        if (more) {
          *reinterpret_cast<uint2*>(da_put + e * 16) = make_uint2(h0[e], h1[e]);
          if constexpr (SPLIT) *reinterpret_cast<uint2*>(da_put + kDaPart + e * 16) = make_uint2(l0[e], l1[e]);
        }
      }
      PROF_MARK(1);
      if (more) {
        fence_async_smem();  // the da tile (generic-proxy stores) is the tensor core's B operand (async proxy)
        mbar_arrive(da_bar);
      }
      PROF_MARK(2);
      // off the chain: dgates overwrite the saved gates in place; bias-gradient column sums
#pragma unroll
      for (int e = 0; e < NB; ++e) {
        if (valid[e]) {
          float4* gp = G4 + (size_t)(rowb[e] + t) * H;
          if (planes) {
            // bf16 hi | lo planes over the 4H-float gate row: [4H bf16 hi | 4H bf16 lo], gate-interleaved column 4u + q
            unsigned char* grow = reinterpret_cast<unsigned char*>(gp) - (size_t)u * 16;  // start of the row
            *reinterpret_cast<uint2*>(grow + (size_t)u * 8) = make_uint2(h0[e], h1[e]);
            if constexpr (SPLIT) *reinterpret_cast<uint2*>(grow + (size_t)H * 8 + (size_t)u * 8) = make_uint2(l0[e], l1[e]);
            bsum.x += da[e][0]; bsum.y += da[e][1]; bsum.z += da[e][2]; bsum.w += da[e][3];
          } else {
            *gp = make_float4(da[e][0], da[e][1], da[e][2], da[e][3]);
          }
        }
      }
      PROF_MARK(3);
      if (!more) break;
      prep(in, ccur);  // the next step's factors
#pragma unroll
      for (int e = 0; e < NB; ++e) ccur[e] = in.cprev[e];
      if (s + 2 < T) load_in(s + 2, in);
      PROF_MARK(4);

      // partial dh^T[H, NS] = Wslice^T da is in TMEM once mma_bar flips; reduce-scatter: my TMEM rows of accumulator a are units of CTA
      // 4a + quarter.  They are staged in the IDLE receive buffer (xbuf[s & 1] took step s-1's partials, summed long ago), in the slot
      // of the destination, and go out as ONE bulk copy per destination (the copy engine moves them: 16-byte remote stores from the
      // cell warps kept the load/store pipe busy for ~2500 cycles per step).  The slot of CTA d in my idle buffer is next written by
      // CTA d's own step-(s+1) copy, which CTA d can only issue after it has received this one.  My own partial goes to xown.
      const int xb = (s + 1) & 1;
      for (int a = 0; a < NACC; ++a) {
        mbar_wait(&mma_bar[a], (uint32_t)(s & 1));
        PROF_MARK(5);
        fence_after_sync();
        uint32_t r[NG8][8], r4[4] = {0u, 0u, 0u, 0u};
        const uint32_t tcol = tb + ((uint32_t)(32 * quarter) << 16) + kDCol + (uint32_t)(a * NS + xc0);
#pragma unroll
        for (int j = 0; j < NG8; ++j) tmem_ld8_nowait(tcol + 8 * j, r[j]);
        if constexpr (NG4 != 0) tmem_ld4_nowait(tcol + 8 * NG8, r4);
        tmem_wait_ld();
        const int owner = 4 * a + quarter;
        const int slot = owner - rank - 1 + (owner > rank ? 0 : C);  // (owner - rank - 1) mod C; == C - 1 for my own units
        XT* stg = (owner == rank ? xown : xbuf + ((size_t)(s & 1) * (C - 1) + slot) * kXElems) + (size_t)xc0 * kUS + lane;
#pragma unroll
        for (int j = 0; j < NG8; ++j) {
#pragma unroll
          for (int i = 0; i < 8; ++i) stg[(8 * j + i) * kUS] = (XT)__uint_as_float(r[j][i]);
        }
        if constexpr (NG4 != 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i) stg[(8 * NG8 + i) * kUS] = (XT)__uint_as_float(r4[i]);
        }
        if (owner == rank) {
          __syncwarp();
          mbar_arrive_elect(&xbar[xb]);  // (release: the cell threads that wait on xbar see my rows of xown)
        } else {
          fence_async_smem();
          pair_bar_sync(2 + quarter);  // the two warps (quarter, half 0 | 1) hold the NS sequences of this slice between them
          if (half == 0) {
            const int back = rank - owner - 1 + (rank > owner ? 0 : C);  // my slot in the owner's buffers
            bulk_s2c_elect(map_to_rank(xb_local, (uint32_t)owner) + (uint32_t)xb * xBufBytes + (uint32_t)back * kXSlice,
                           xb_local + (uint32_t)(s & 1) * xBufBytes + (uint32_t)slot * kXSlice, kXSlice,
                           map_to_rank(xbar_local, (uint32_t)owner) + (uint32_t)xb * 8u);
          }
        }
        PROF_MARK(6);
      }
      fence_before_sync();
      // all C partials of my units are in place
      mbar_wait(&xbar[xb], (uint32_t)((s >> 1) & 1));
      PROF_MARK(7);
      if (tid == 0 && s + 3 < T) mbar_arrive_expect_tx(&xbar[xb], xBufBytes);  // refilled at step s + 2
      // recurrent gradient of my cells for the next step: sum of the C partials
      const XT* xr = xbuf + (size_t)xb * (C - 1) * kXElems + (size_t)s0 * kUS + lane;
#pragma unroll
      for (int e = 0; e < NB; ++e) dhrec[e] = (float)xown[(s0 + e) * kUS + lane];
      for (int r2 = 0; r2 < C - 1; ++r2) {
#pragma unroll
        for (int e = 0; e < NB; ++e) dhrec[e] += (float)xr[(size_t)r2 * kXElems + e * kUS];
      }
      PROF_MARK(8);
    }
    PROF_PRINT("bwd cell [loop chain-math+sts fence+arrive dgate-stores prep+prefetch wait-mma ld+send wait-xchg sum]", blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (tid == 0 || tid == 255), T);
  }
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // nobody exits while a peer could still be sending to it
  if (wid == kCellWarps) tmem_dealloc(tb, kTmemCols);

  if (planes) {
    // (1) bias-gradient partials: one [4H] row (GI order) per (direction slot, group, tile); CTA `rank` owns columns [128 rank, +128)
    if (p.bias_partial != nullptr) {
      float4* red = reinterpret_cast<float4*>(xbuf);  // [8 warps][32 units] (the exchange buffers are idle now)
      if (wid < kCellWarps) red[wid * kUS + lane] = bsum;
      __syncthreads();
      if (wid == 0) {
        float4 b = red[lane];
#pragma unroll
        for (int w2 = 1; w2 < kCellWarps; ++w2) {
          const float4 o = red[w2 * kUS + lane];
          b.x += o.x; b.y += o.y; b.z += o.z; b.w += o.w;
        }
        const int ntiles = (int)gridDim.x / C;
        float* dst = p.bias_partial + (((size_t)blockIdx.z * p.G + g) * ntiles + tile) * 4 * H + 4 * (kUS * rank + lane);
        *reinterpret_cast<float4*>(dst) = b;
      }
    }
    // (2) zero tail rows [T, tail_end) of my sequences in my 128 gate columns (both planes): the TN GEMM reads whole 64-row boxes
    const int tail_end = min(Tmax, ((T + 63) / 64) * 64 + 1), ntail = tail_end - T;
    constexpr int kChunks = kRows * 2 / 16;  // 16-byte chunks of 128 bf16
    float* const Gbase = dir ? p.gates[1] : p.gates[0];
    for (int i = tid; i < nvalid * ntail * kChunks * 2; i += kThreads) {
      const int cc = i % kChunks, pl = (i / kChunks) & 1, rr = (i / (2 * kChunks)) % ntail, qq = i / (2 * kChunks * ntail);
      unsigned char* row = reinterpret_cast<unsigned char*>(Gbase + ((size_t)(nbase + qq) * Tmax + T + rr) * 4 * H);
      *reinterpret_cast<uint4*>(row + (size_t)pl * 4 * H * 2 + (size_t)kRows * rank * 2 + cc * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

template <int NB>
size_t fwd_smem_tc(int H, bool split) {
  const int npart = split ? 2 : 1, C = H / kUS;
  return 1024 + (size_t)2 * C * npart * TileT<NB>::kHSlice + (size_t)(2 * C + 1) * 8 + 64;
}
template <int NB>
size_t bwd_smem_tc(int H, bool split) {
  const int npart = split ? 2 : 1, C = H / kUS;
  return 1024 + (size_t)npart * 16 * TileT<NB>::kChunk + (size_t)(2 * (C - 1) + 1) * TileT<NB>::xslice(split ? 4 : 2) + 64;
}

// clusters of C CTAs of this kernel that are co-resident on the device (cached per kernel and shared-memory size)
template <typename Kern>
int max_active_clusters(Kern kern, int C, size_t smem) {
  static std::mutex mu;
  static std::map<std::pair<const void*, size_t>, int> cache;
  std::lock_guard<std::mutex> lk(mu);
  const auto key = std::make_pair(reinterpret_cast<const void*>(kern), smem * 64 + (size_t)C);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  int n = 0;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(C * 64));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) n = 0;
  }
  (void)cudaGetLastError();
  if (n <= 0) n = C == 8 ? 15 : 148 / C;  // (measured on B200; only used when the query is not available)
  cache[key] = n;
  return n;
}

// tile width: 32 sequences per cluster unless tiles of 40 save a whole wave of clusters (a tile of 40 costs ~15 % more per step)
int pick_nb(int clusters32, int clusters40, int max_active) {
  const int w32 = (clusters32 + max_active - 1) / max_active, w40 = (clusters40 + max_active - 1) / max_active;
  return 1.15 * w40 < (double)w32 ? 5 : 4;
}

template <typename Kern, typename Args>
cudaError_t launch_cluster_tc(Kern kern, const Args& a, int H, int NS, size_t smem, int ndir, cudaStream_t st) {
  const int C = H / kUS;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(C * ((a.B + NS - 1) / NS)), (unsigned)a.G, (unsigned)ndir);
  cfg.blockDim = dim3(kThreads);
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

int forced_nb() {  // IB200_CLUSTER_NB = 4 | 5 pins the tile width (tests)
  static const int v = [] { const char* e = getenv("IB200_CLUSTER_NB"); return e ? atoi(e) : 0; }();
  return v == 4 || v == 5 ? v : 0;
}

int bwd_nb(const LstmBwdArgs& a, int H, bool split) {
  if (forced_nb()) return forced_nb();
  const int C = H / kUS, per = a.G * a.ndir;
  const int ma = split ? max_active_clusters(lstm_bwd_cltc_kernel<4, true>, C, bwd_smem_tc<4>(H, true))
                       : max_active_clusters(lstm_bwd_cltc_kernel<4, false>, C, bwd_smem_tc<4>(H, false));
  if (bwd_smem_tc<5>(H, split) > 227 * 1024) return 4;
  return pick_nb(per * ((a.B + 31) / 32), per * ((a.B + 39) / 40), ma);
}

}  // namespace

bool lstm_cluster_tc_supports(int H) { return H == 128 || H == 256; }

// bias partial rows written per direction by launch_lstm_bwd_cluster_tc in planes mode: one per (group, sequence tile)
int lstm_bwd_cluster_tc_cta_count(const LstmBwdArgs& a, int H, int precision) {
  const int ns = 8 * bwd_nb(a, H, precision == 0);
  return a.G * ((a.B + ns - 1) / ns);
}

cudaError_t launch_lstm_fwd_cluster_tc(const LstmFwdArgs& a, int H, int precision, cudaStream_t st) {
  if (!lstm_cluster_tc_supports(H)) return cudaErrorInvalidValue;
  if (a.y != nullptr && !a.planes) return cudaErrorInvalidValue;  // this path writes y as bf16 planes only
  const bool split = precision == 0, train = a.gates[a.dir0] != nullptr;
  const int C = H / kUS, per = a.G * a.ndir;
  int nb = forced_nb();
  if (!nb) {
    const size_t s4 = fwd_smem_tc<4>(H, split);
    const int ma = split ? max_active_clusters(lstm_fwd_cltc_kernel<4, true, true>, C, s4) : max_active_clusters(lstm_fwd_cltc_kernel<4, false, true>, C, s4);
    nb = fwd_smem_tc<5>(H, split) > 227 * 1024 ? 4 : pick_nb(per * ((a.B + 31) / 32), per * ((a.B + 39) / 40), ma);
  }
#define IB200_FWD_TC(NB_)                                                                                                          \
  {                                                                                                                                \
    const size_t smem = fwd_smem_tc<NB_>(H, split);                                                                                \
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;                                                                   \
    if (split) return train ? launch_cluster_tc(lstm_fwd_cltc_kernel<NB_, true, true>, a, H, 8 * NB_, smem, a.ndir, st)            \
                            : launch_cluster_tc(lstm_fwd_cltc_kernel<NB_, true, false>, a, H, 8 * NB_, smem, a.ndir, st);          \
    return train ? launch_cluster_tc(lstm_fwd_cltc_kernel<NB_, false, true>, a, H, 8 * NB_, smem, a.ndir, st)                      \
                 : launch_cluster_tc(lstm_fwd_cltc_kernel<NB_, false, false>, a, H, 8 * NB_, smem, a.ndir, st);                    \
  }
  if (nb == 5) IB200_FWD_TC(5)
  IB200_FWD_TC(4)
#undef IB200_FWD_TC
}

cudaError_t launch_lstm_bwd_cluster_tc(const LstmBwdArgs& a, int H, int precision, cudaStream_t st) {
  if (!lstm_cluster_tc_supports(H)) return cudaErrorInvalidValue;
  const bool split = precision == 0;
  const int nb = bwd_nb(a, H, split);
#define IB200_BWD_TC(NB_)                                                                                       \
  {                                                                                                             \
    const size_t smem = bwd_smem_tc<NB_>(H, split);                                                             \
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;                                                \
    return split ? launch_cluster_tc(lstm_bwd_cltc_kernel<NB_, true>, a, H, 8 * NB_, smem, a.ndir, st)          \
                 : launch_cluster_tc(lstm_bwd_cltc_kernel<NB_, false>, a, H, 8 * NB_, smem, a.ndir, st);        \
  }
  if (nb == 5) IB200_BWD_TC(5)
  IB200_BWD_TC(4)
#undef IB200_BWD_TC
}

}  // namespace ib200
