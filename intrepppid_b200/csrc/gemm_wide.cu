// tcgen05 + TMA token-row GEMMs for the hidden sizes of the cluster path (H = 128, 192, 256; BASELINE config 5 is H = 256).
//
// Same contracts as gemm_tc.cu, but neither operand is resident: both stream through shared memory in 64-wide k chunks.
//   NT : C[row, NC] (=|+=) sum_s A_s[row, K] W_s[NC, K]^T (+ bias)     xproj (K = 2H, NC = 4H), dY (K = 4H per live direction,
//        NC = 2H), dX0 (NC = H).  Tile 128 rows x BN columns (BN <= 256), persistent CTAs over (row tile, column tile) items with
//        the column tile fastest, so the CTAs that share an A tile run side by side and hit it in L2; accumulators double
//        buffered in TMEM (2 x BN columns).
//   TN : P[split][4H, NB] = sum over the split's token rows of dA[row, 4H]^T [B1 | B2][row, NB]   (dW_ih | dW_hh of one
//        (layer, direction) in ONE pass over the dgates; B2 = Y_l shifted by one step).  Tile 128 gate columns x H B-columns,
//        split-K over 64-row items of one group; partials go through launch_dw_reduce (deterministic).
// Every operand is a bf16 hi|lo PLANE matrix (a row of K values lives in the bytes of K floats: [K bf16 hi | K bf16 lo]) written by
// the cluster recurrent kernels (lstm_cluster.cu, planes = 1) or by the preparation kernels below; one thread issues the TMA loads
// (cp.async.bulk.tensor, 128-byte swizzle), one thread issues tcgen05.mma (3 MMAs per product in fp32 mode: hi*hi + hi*lo + lo*hi),
// four warps drain TMEM.  fp32 mode: 2 stages of 96 KB; bf16 mode: 4 stages of 48 KB.
#include <algorithm>

#include "kernels.h"
#include "tc05.cuh"
#include "tma_host.h"

namespace ib200 {
namespace {

using namespace tc;

constexpr int kBM = 128;          // rows per NT tile / gate columns per TN tile (UMMA M)
constexpr int kBK = 64;           // k elements per stage (one 128-byte swizzle row of bf16)
constexpr int kATile = kBM * 128; // [128 x 64] bf16 plane tile
constexpr int kLensSmem = 1024;   // groups whose T_eff is cached in shared memory

template <bool SPLIT>
constexpr int wide_stages() { return SPLIT ? 2 : 4; }

template <int STAGES>
struct WideBars {
  uint64_t full[STAGES], empty[STAGES], tfull[2], tempty[2];
  uint32_t tmem_base;
};

constexpr uint32_t tmem_cols_pow2(uint32_t n) { return n <= 32 ? 32 : (n <= 64 ? 64 : (n <= 128 ? 128 : (n <= 256 ? 256 : 512))); }

struct NTWMaps {
  CUtensorMap a[2][2];  // [source][plane]: 2D {K, rows}, box {64, 128}
  CUtensorMap w[2][2];  // [source][plane]: 2D {K, NC},   box {64, BN}
};

__device__ __forceinline__ bool tile_live(const GemmNTArgs& p, const int* teff, long long row0, long long nrows) {
  const long long rl = min(row0 + kBM, nrows) - 1;
  const int na = (int)(row0 / p.Tmax), nb = (int)(rl / p.Tmax);
  return !(na == nb && (int)(row0 % p.Tmax) >= teff[na / p.B]);
}

// ------------------------------------------------------------------------------------------------------------------------------
// NT.  6 warps: 0 TMA producer, 1 MMA issuer + TMEM owner, 2-5 epilogue (TMEM -> smem transpose -> coalesced stores).
// ------------------------------------------------------------------------------------------------------------------------------
template <int BN, bool SPLIT>
__global__ void __launch_bounds__(192, 1) gemm_nt_wide_kernel(const __grid_constant__ NTWMaps maps, const GemmNTArgs p) {
  constexpr int NPART = SPLIT ? 2 : 1, STAGES = wide_stages<SPLIT>();
  constexpr uint32_t kWTile = BN * 128;                      // [BN x 64] bf16 plane tile
  constexpr uint32_t kStageBytes = NPART * (kATile + kWTile);
  constexpr uint32_t kTmemCols = tmem_cols_pow2(2 * BN);
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  float* Est = reinterpret_cast<float*>(smem + (size_t)STAGES * kStageBytes);  // [4 warps][32 rows][36] epilogue transpose
  WideBars<STAGES>* bars = reinterpret_cast<WideBars<STAGES>*>(Est + 4 * 32 * 36);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long nrows = (long long)p.G * p.B * p.Tmax;
  const int row_tiles = (int)((nrows + kBM - 1) / kBM), col_tiles = p.NC / BN;
  const long long items = (long long)row_tiles * col_tiles;
  const int kslices = p.K / kBK, KC = p.nsrc * kslices;
  __shared__ int teff_s[kLensSmem];
  const int* teff = p.G <= kLensSmem ? teff_s : p.lens + p.G;
  for (int i = tid; i < min(p.G, kLensSmem); i += 192) teff_s[i] = p.lens[p.G + i];

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tfull[a], 1);
      mbar_init(&bars->tempty[a], 128);
    }
    mbar_init_fence();
    for (int sidx = 0; sidx < p.nsrc; ++sidx)
      for (int pl = 0; pl < NPART; ++pl) {
        tma_prefetch_desc(&maps.a[sidx][pl]);
        tma_prefetch_desc(&maps.w[sidx][pl]);
      }
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const long long row0 = (item / col_tiles) * kBM;
        const int n0 = (int)(item % col_tiles) * BN;
        if (!tile_live(p, teff, row0, nrows)) continue;
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const int stage = it % STAGES, src = kc / kslices, k0 = (kc % kslices) * kBK;
          mbar_wait(&bars->empty[stage], ((it / STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&bars->full[stage], kStageBytes);
          unsigned char* dst = smem + (size_t)stage * kStageBytes;
#pragma unroll
          for (int pl = 0; pl < NPART; ++pl) {
            tma_load_2d(dst + pl * kATile, &maps.a[src][pl], &bars->full[stage], k0, (int)row0);
            tma_load_2d(dst + NPART * kATile + pl * kWTile, &maps.w[src][pl], &bars->full[stage], k0, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    {  // (whole warp, uniform control flow; one elected lane issues: tc05.cuh)
      constexpr uint32_t idesc = idesc_bf16(kBM, BN, false, false);
      uint32_t it = 0, tl = 0;
      for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        if (!tile_live(p, teff, (item / col_tiles) * kBM, nrows)) continue;
        const uint32_t acc = tl & 1;
        mbar_wait(&bars->tempty[acc], ((tl >> 1) & 1) ^ 1);
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const int stage = it % STAGES;
          mbar_wait(&bars->full[stage], (it / STAGES) & 1);
          fence_after_sync();
          const uint32_t a_hi = smem_u32(smem + (size_t)stage * kStageBytes), a_lo = a_hi + kATile;
          const uint32_t b_hi = a_hi + NPART * kATile, b_lo = b_hi + kWTile;
#pragma unroll
          for (int k16 = 0; k16 < kBK / 16; ++k16) {
            const uint32_t ko = k16 * 32;  // 16 bf16 = 32 bytes along K inside the swizzle row
            const uint64_t ah = smem_desc_sw128(a_hi + ko, 1024, 0), bh = smem_desc_sw128(b_hi + ko, 1024, 0);
            mma_bf16_ss_elect(d_tmem, ah, bh, idesc, (kc | k16) != 0);
            if constexpr (SPLIT) {
              const uint64_t al = smem_desc_sw128(a_lo + ko, 1024, 0), bl = smem_desc_sw128(b_lo + ko, 1024, 0);
              mma_bf16_ss_elect(d_tmem, ah, bl, idesc, true);
              mma_bf16_ss_elect(d_tmem, al, bh, idesc, true);
            }
          }
          mma_commit_elect(&bars->empty[stage]);  // stage free once these MMAs have read it
        }
        mma_commit_elect(&bars->tfull[acc]);
        ++tl;
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    float* est = Est + q * 32 * 36;
    const int tr = lane >> 3, tc4 = (lane & 7) * 4;  // transposed phase: 4 rows x 8 float4 per pass
    uint32_t tl = 0;
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
      const long long row0 = (item / col_tiles) * kBM;
      const int n0 = (int)(item % col_tiles) * BN;
      if (!tile_live(p, teff, row0, nrows)) continue;
      const uint32_t acc = tl & 1;
      mbar_wait(&bars->tfull[acc], (tl >> 1) & 1);
      fence_after_sync();
      const long long rbase = row0 + q * 32;
      bool ok = rbase + lane < nrows;
      if (ok) ok = (int)((rbase + lane) % p.Tmax) < teff[(int)((rbase + lane) / p.Tmax) / p.B];
      const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
      uint32_t r[2][32];
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      tmem_ld32_nowait(t_row, r[0]);
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        tmem_wait_ld();
        if (c + 1 < BN / 32) tmem_ld32_nowait(t_row + (c + 1) * 32, r[(c + 1) & 1]);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(est + lane * 36 + 4 * j) =
              make_uint4(r[c & 1][4 * j], r[c & 1][4 * j + 1], r[c & 1][4 * j + 2], r[c & 1][4 * j + 3]);
        __syncwarp();
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c * 32 + tc4));
#pragma unroll
        for (int itr = 0; itr < 8; ++itr) {  // 8 lanes cover one row's 32 columns: every store instruction writes 4 full 128-byte rows
          const int rr = itr * 4 + tr;
          if (okmask & (1u << rr)) {
            float4 o = *reinterpret_cast<const float4*>(est + rr * 36 + tc4);
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            float4* dst = reinterpret_cast<float4*>(p.C + (rbase + rr) * p.ldc + n0 + c * 32 + tc4);
            if (p.accumulate) {
              const float4 e = *dst;
              o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
            }
            *dst = o;
          }
        }
        __syncwarp();
      }
      fence_before_sync();
      mbar_arrive(&bars->tempty[acc]);
      ++tl;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

template <int BN>
cudaError_t launch_nt_wide(const GemmNTArgs& a, int precision, cudaStream_t st) {
  const int npart = precision == 0 ? 2 : 1, stages = precision == 0 ? 2 : 4;
  const size_t smem = 1024 + (size_t)stages * npart * (kATile + BN * 128) + 4 * 32 * 36 * sizeof(float) + sizeof(WideBars<4>) + 64;
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  NTWMaps maps;
  const long long nrows = (long long)a.G * a.B * a.Tmax;
  for (int s = 0; s < a.nsrc; ++s)
    for (int pl = 0; pl < npart; ++pl) {
      const uint64_t da[2] = {(uint64_t)a.K, (uint64_t)nrows}, sa[1] = {(uint64_t)a.lda * 4};
      const uint32_t ba[2] = {64, 128};
      if (!make_tmap_bf16_sw128(&maps.a[s][pl], reinterpret_cast<const unsigned char*>(a.A[s]) + (size_t)pl * a.plane_bytes, 2, da, sa, ba))
        return cudaErrorInvalidConfiguration;
      const uint64_t dw[2] = {(uint64_t)a.K, (uint64_t)a.NC}, sw[1] = {(uint64_t)a.K * 4};
      const uint32_t bw[2] = {64, (uint32_t)BN};
      if (!make_tmap_bf16_sw128(&maps.w[s][pl], reinterpret_cast<const unsigned char*>(a.W[s]) + (size_t)pl * a.K * 2, 2, dw, sw, bw))
        return cudaErrorInvalidConfiguration;
    }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long items = ((nrows + kBM - 1) / kBM) * (a.NC / BN);
  const unsigned grid = (unsigned)std::min<long long>(items, sms);
  cudaError_t e;
  if (precision == 0) {
    e = cudaFuncSetAttribute(gemm_nt_wide_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    gemm_nt_wide_kernel<BN, true><<<grid, 192, smem, st>>>(maps, a);
  } else {
    e = cudaFuncSetAttribute(gemm_nt_wide_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    gemm_nt_wide_kernel<BN, false><<<grid, 192, smem, st>>>(maps, a);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------------------
// TN.  grid (m_tiles * n_tiles * splits, G).  CTA = 128 gate columns (A, MN-major) x BN B-columns (MN-major), K = the 64-row items
// of its split.  6 warps: 0 TMA producer, 1 MMA issuer + TMEM owner, 2-5 drain TMEM at the end.
// ------------------------------------------------------------------------------------------------------------------------------
constexpr int kBlk = 64 * 128;  // one [64 k-rows x 64 columns] bf16 block

template <int STAGES>
struct TNWBars {
  uint64_t full[STAGES], empty[STAGES], done;
  uint32_t tmem_base;
};
struct TNWMaps {
  CUtensorMap a[2];   // dgates planes: 3D {4H, Tmax, N}, box {64, 64, 1}
  CUtensorMap b[2];   // first B source planes
  CUtensorMap b2[2];  // second B source planes (NB2 > 0)
};

template <int BN, bool SPLIT>
__global__ void __launch_bounds__(192, 1) gemm_tn_wide_kernel(const __grid_constant__ TNWMaps maps, const GemmTNArgs p, const int m_tiles,
                                                              const int n_tiles, const int splits) {
  constexpr int NPART = SPLIT ? 2 : 1, STAGES = wide_stages<SPLIT>();
  constexpr int kABytes = 2 * kBlk, kBBytes = (BN / 64) * kBlk;  // per plane
  constexpr int kStageBytes = NPART * (kABytes + kBBytes);
  constexpr uint32_t kTmemCols = tmem_cols_pow2(BN);
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  TNWBars<STAGES>* bars = reinterpret_cast<TNWBars<STAGES>*>(smem + (size_t)STAGES * kStageBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = blockIdx.y;
  const int tile = (int)blockIdx.x % (m_tiles * n_tiles), split = (int)blockIdx.x / (m_tiles * n_tiles);
  const int mt = tile % m_tiles, nt = tile / m_tiles;
  const int T = p.lens[p.G + g];
  const int tiles_per_seq = (T + 63) / 64;
  const int items = p.B * tiles_per_seq;
  const int my_items = split < items ? (items - split + splits - 1) / splits : 0;
  // which B source this column tile comes from
  const int n1_tiles = p.NB1 / BN;
  const bool second = nt >= n1_tiles;
  const int bcol = second ? p.col02 + (nt - n1_tiles) * BN : p.col0 + nt * BN, bshift = second ? p.shift2 : p.shift;
  const CUtensorMap* bmap = second ? maps.b2 : maps.b;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    mbar_init(&bars->done, 1);
    mbar_init_fence();
    for (int pl = 0; pl < NPART; ++pl) {
      tma_prefetch_desc(&maps.a[pl]);
      tma_prefetch_desc(&bmap[pl]);
    }
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = split; item < items; item += splits, ++it) {
        const int stage = it % STAGES;
        const int n = g * p.B + item / tiles_per_seq, t0 = (item % tiles_per_seq) * 64;
        mbar_wait(&bars->empty[stage], ((it / STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->full[stage], kStageBytes);
        unsigned char* a_dst = smem + (size_t)stage * kStageBytes;
        unsigned char* b_dst = a_dst + NPART * kABytes;
#pragma unroll
        for (int pl = 0; pl < NPART; ++pl) {
#pragma unroll
          for (int blk = 0; blk < 2; ++blk)
            tma_load_3d(a_dst + pl * kABytes + blk * kBlk, &maps.a[pl], &bars->full[stage], mt * 128 + blk * 64, t0, n);
#pragma unroll
          for (int blk = 0; blk < BN / 64; ++blk)
            tma_load_3d(b_dst + pl * kBBytes + blk * kBlk, &bmap[pl], &bars->full[stage], bcol + blk * 64, t0 + bshift, n);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    {  // (whole warp, uniform control flow; one elected lane issues: tc05.cuh)
      constexpr uint32_t idesc = idesc_bf16(128, BN, true, true);
      for (int it = 0; it < my_items; ++it) {
        const int stage = it % STAGES;
        mbar_wait(&bars->full[stage], (it / STAGES) & 1);
        fence_after_sync();
        const uint32_t a_hi = smem_u32(smem + (size_t)stage * kStageBytes), a_lo = a_hi + kABytes;
        const uint32_t b_hi = a_hi + NPART * kABytes, b_lo = b_hi + kBBytes;
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) {
          const uint32_t ko = k16 * 16 * 128;  // 16 k-rows of 128 bytes inside every block
          const uint64_t ah = smem_desc_sw128(a_hi + ko, 1024, kBlk), bh = smem_desc_sw128(b_hi + ko, 1024, kBlk);
          mma_bf16_ss_elect(tmem_base, ah, bh, idesc, (it | k16) != 0);
          if constexpr (SPLIT) {
            const uint64_t al = smem_desc_sw128(a_lo + ko, 1024, kBlk), bl = smem_desc_sw128(b_lo + ko, 1024, kBlk);
            mma_bf16_ss_elect(tmem_base, ah, bl, idesc, true);
            mma_bf16_ss_elect(tmem_base, al, bh, idesc, true);
          }
        }
        mma_commit_elect(&bars->empty[stage]);
      }
      mma_commit_elect(&bars->done);
    }
  }
  __syncthreads();

  // ===================== epilogue: TMEM -> partial[g][split][4H][NB], rows mt*128.., columns nt*BN.. =====================
  if (warp >= 2) {
    const int q = warp & 3;
    if (my_items > 0) {
      mbar_wait(&bars->done, 0);
      fence_after_sync();
    }
    float* orow = p.partial + ((size_t)g * splits + split) * ((size_t)p.KA * p.NB) + (size_t)(mt * 128 + q * 32 + lane) * p.NB + (size_t)nt * BN;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      if (my_items > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, r);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(orow + c * 32 + 4 * j) = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                                     __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

template <int BN>
cudaError_t launch_tn_wide(const GemmTNArgs& a, int precision, int splits, cudaStream_t st) {
  const int npart = precision == 0 ? 2 : 1, stages = precision == 0 ? 2 : 4;
  const size_t smem = 1024 + (size_t)stages * npart * (2 + BN / 64) * kBlk + sizeof(TNWBars<4>) + 64;
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  TNWMaps maps;
  const uint32_t box[3] = {64, 64, 1};
  const uint64_t nseq = (uint64_t)a.G * a.B;
  for (int pl = 0; pl < npart; ++pl) {
    const uint64_t da[3] = {(uint64_t)a.KA, (uint64_t)a.Tmax, nseq}, sa[2] = {(uint64_t)a.KA * 4, (uint64_t)a.Tmax * a.KA * 4};
    if (!make_tmap_bf16_sw128(&maps.a[pl], reinterpret_cast<const unsigned char*>(a.A) + (size_t)pl * a.KA * 2, 3, da, sa, box))
      return cudaErrorInvalidConfiguration;
    const uint64_t db[3] = {(uint64_t)a.ldb, (uint64_t)a.Tmax, nseq}, sb[2] = {(uint64_t)a.ldb * 4, (uint64_t)a.Tmax * a.ldb * 4};
    if (!make_tmap_bf16_sw128(&maps.b[pl], reinterpret_cast<const unsigned char*>(a.Bsrc) + (size_t)pl * a.ldb * 2, 3, db, sb, box))
      return cudaErrorInvalidConfiguration;
    if (a.NB > a.NB1) {
      const uint64_t d2[3] = {(uint64_t)a.ldb2, (uint64_t)a.Tmax, nseq}, s2[2] = {(uint64_t)a.ldb2 * 4, (uint64_t)a.Tmax * a.ldb2 * 4};
      if (!make_tmap_bf16_sw128(&maps.b2[pl], reinterpret_cast<const unsigned char*>(a.Bsrc2) + (size_t)pl * a.ldb2 * 2, 3, d2, s2, box))
        return cudaErrorInvalidConfiguration;
    } else {
      maps.b2[pl] = maps.b[pl];
    }
  }
  const int m_tiles = a.KA / 128, n_tiles = a.NB / BN;
  dim3 grid((unsigned)(m_tiles * n_tiles * splits), (unsigned)a.G);
  cudaError_t e;
  if (precision == 0) {
    e = cudaFuncSetAttribute(gemm_tn_wide_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    gemm_tn_wide_kernel<BN, true><<<grid, 192, smem, st>>>(maps, a, m_tiles, n_tiles, splits);
  } else {
    e = cudaFuncSetAttribute(gemm_tn_wide_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    gemm_tn_wide_kernel<BN, false><<<grid, 192, smem, st>>>(maps, a, m_tiles, n_tiles, splits);
  }
  return cudaGetLastError();
}

// ---- operand preparation ----------------------------------------------------------------------------------------------------------
// W_ih of one (layer, direction) as plane matrices: out_w[gi][.] = rows permuted to gate-interleaved order ([4H rows][K] -- the xproj
// W operand), out_wT[k][.] = the same transposed ([K rows][4H] -- the dY / dX0 W operand); out_b[gi] = b_ih + b_hh permuted.
template <bool SPLIT>
__global__ void prep_wih_planes_kernel(const float* __restrict__ w, const float* __restrict__ b_ih, const float* __restrict__ b_hh, int H, int K,
                                       float* __restrict__ out_w, float* __restrict__ out_wT, float* __restrict__ out_b) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 4 * H * K) return;
  const int gi = idx / K, k = idx % K, row = gi_to_torch_row(gi, H);
  const float v = w[(size_t)row * K + k];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v), lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  if (out_w != nullptr) {
    __nv_bfloat16* r = reinterpret_cast<__nv_bfloat16*>(out_w + (size_t)gi * K);
    r[k] = hi;
    if (SPLIT) r[K + k] = lo;
  }
  if (out_wT != nullptr) {
    __nv_bfloat16* r = reinterpret_cast<__nv_bfloat16*>(out_wT + (size_t)k * 4 * H);
    r[gi] = hi;
    if (SPLIT) r[4 * H + gi] = lo;
  }
  if (out_b != nullptr && k == 0) out_b[gi] = b_ih[row] + b_hh[row];
}

// layer-0 input rows x[row] = scale[g][tok] * emb[tok] as planes [rows][H] (the first B source of the layer-0 weight-gradient GEMM);
// rows t >= T_eff up to the end of the last 64-row box are zeros
template <bool SPLIT>
__global__ void gather_x0_planes_kernel(int G, int B, int Tmax, int V, int H, const int* __restrict__ lens, const int* __restrict__ tok,
                                        const float* __restrict__ emb, const float* __restrict__ scale, float* __restrict__ out) {
  const int n = blockIdx.y, g = n / B;
  const int T = lens[G + g], tail_end = min(Tmax, ((T + 63) / 64) * 64);
  for (int t = blockIdx.x; t < tail_end; t += gridDim.x) {
    __nv_bfloat16* r = reinterpret_cast<__nv_bfloat16*>(out + ((size_t)n * Tmax + t) * H);
    const int tk = t < T ? tok[(size_t)n * Tmax + t] : -1;
    const float sc = (tk >= 0 && scale != nullptr) ? scale[(size_t)g * V + tk] : 1.0f;
    for (int e = threadIdx.x; e < H; e += blockDim.x) {
      const float v = tk >= 0 ? sc * emb[(size_t)tk * H + e] : 0.f;
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      r[e] = hi;
      if (SPLIT) r[H + e] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
  }
}

}  // namespace

bool gemm_wide_supports(int H) { return H > 64 && H <= 256 && H % 64 == 0; }

// A_s: plane rows of K values (lda floats pitch, plane_bytes = K * 2), W_s: plane matrices [NC][K] from launch_prep_wih_planes
cudaError_t launch_gemm_nt_wide(const GemmNTArgs& a, int precision, cudaStream_t st) {
  if (a.K % kBK != 0 || a.lda % 4 != 0 || a.ldc % 4 != 0 || a.plane_bytes != a.K * 2 || a.nsrc < 1 || a.nsrc > 2) return cudaErrorInvalidConfiguration;
  if (a.NC % 256 == 0) return launch_nt_wide<256>(a, precision, st);
  if (a.NC % 192 == 0) return launch_nt_wide<192>(a, precision, st);
  if (a.NC % 128 == 0) return launch_nt_wide<128>(a, precision, st);
  return cudaErrorInvalidConfiguration;
}

int gemm_tn_wide_splits(int KA, int NB, int BN, int G) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  return std::max(1, sms / std::max(1, (KA / 128) * (NB / BN) * G));
}

// A: dgates planes [rows][4H]; Bsrc / Bsrc2: plane rows of ldb / ldb2 values; the column tile width is H = KA / 4 (NB1, NB2 multiples
// of it).  partial: [G][ctas_per_group][KA][NB] -- a.ctas_per_group is the split count (gemm_tn_wide_splits)
cudaError_t launch_gemm_tn_wide(const GemmTNArgs& a, int precision, cudaStream_t st) {
  const int H = a.KA / 4;
  if (!gemm_wide_supports(H) || a.colsum || a.tok != nullptr || a.Bsrc == nullptr) return cudaErrorInvalidConfiguration;
  if (a.NB1 % H != 0 || (a.NB - a.NB1) % H != 0 || a.col0 % 8 != 0 || a.ldb % 4 != 0) return cudaErrorInvalidConfiguration;
  if (a.NB > a.NB1 && (a.Bsrc2 == nullptr || a.col02 % 8 != 0 || a.ldb2 % 4 != 0)) return cudaErrorInvalidConfiguration;
  switch (H) {
    case 256: return launch_tn_wide<256>(a, precision, a.ctas_per_group, st);
    case 192: return launch_tn_wide<192>(a, precision, a.ctas_per_group, st);
    case 128: return launch_tn_wide<128>(a, precision, a.ctas_per_group, st);
    default: return cudaErrorInvalidConfiguration;
  }
}

cudaError_t launch_prep_wih_planes(const float* w, const float* b_ih, const float* b_hh, int H, int K, float* out_w, float* out_wT,
                                   float* out_b, int precision, cudaStream_t st) {
  const int total = 4 * H * K;
  if (precision == 0) prep_wih_planes_kernel<true><<<(total + 255) / 256, 256, 0, st>>>(w, b_ih, b_hh, H, K, out_w, out_wT, out_b);
  else prep_wih_planes_kernel<false><<<(total + 255) / 256, 256, 0, st>>>(w, b_ih, b_hh, H, K, out_w, out_wT, out_b);
  return cudaGetLastError();
}

cudaError_t launch_gather_x0_planes(int G, int B, int Tmax, int V, int H, const int* lens, const int* tok, const float* emb,
                                    const float* scale, float* out, int precision, cudaStream_t st) {
  dim3 grid((unsigned)std::min(Tmax, 256), (unsigned)(G * B));
  if (precision == 0) gather_x0_planes_kernel<true><<<grid, 128, 0, st>>>(G, B, Tmax, V, H, lens, tok, emb, scale, out);
  else gather_x0_planes_kernel<false><<<grid, 128, 0, st>>>(G, B, Tmax, V, H, lens, tok, emb, scale, out);
  return cudaGetLastError();
}

}  // namespace ib200
