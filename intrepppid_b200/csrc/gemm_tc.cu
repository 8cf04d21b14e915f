// tcgen05 (5th-gen tensor core) versions of the token-row GEMMs, H = 64 shapes.  Same contracts as gemm.cu:
//   NT : C[row, NC] (=|+=) sum_s A_s[row, K] W_s[NC, K]^T (+ bias)     xproj (NC=256,K=128), dY (NC=128,K=256), dX0 (NC=64,K=512)
//   TN : P[cta][KA, NB] = sum_rows A[row, KA]^T Bop[row, NB]           dW_ih / dW_hh / db partials (KA=256, NB=128|64)
// Operands are fp32 in HBM.  Loader warps read them (coalesced, masked by T_eff, optionally gathered/shifted), split every value
// into bf16 hi + lo and write both planes into 128-byte-swizzled shared-memory tiles; ONE thread issues tcgen05.mma
// (kind::f16, bf16 x bf16 -> fp32 in TMEM), three MMAs per k-step in fp32 mode (hi*hi + hi*lo + lo*hi); accumulators live in TMEM
// and are drained by four epilogue warps with tcgen05.ld.  Stages are handed over with mbarriers; tcgen05.commit releases them.
#include <algorithm>

#include "kernels.h"
#include "tc05.cuh"
#include "tma_host.h"

namespace ib200 {
namespace {

using namespace tc;

constexpr int kBM = 128;      // rows per tile (UMMA M)
constexpr int kBK = 64;       // k elements per stage (= one 128-byte swizzle row of bf16)
constexpr int kStagesNT = 2;
constexpr int kTileBytes = kBM * 128;  // one [128 x 64] bf16 plane

struct NTBarriers {
  uint64_t full[kStagesNT], empty[kStagesNT], tfull[2], tempty[2];
  uint32_t tmem_base;
};

// teff = the T_eff row of lens ([G] ints; the TMA kernel keeps a shared-memory copy: one L2 round trip per tile and role otherwise)
__device__ __forceinline__ bool nt_tile_live(const GemmNTArgs& p, const int* teff, long long row0, long long nrows) {
  const long long rl = min(row0 + kBM, nrows) - 1;
  const int na = (int)(row0 / p.Tmax), nb = (int)(rl / p.Tmax);
  return !(na == nb && (int)(row0 % p.Tmax) >= teff[na / p.B]);
}
__device__ __forceinline__ bool nt_tile_live(const GemmNTArgs& p, long long row0, long long nrows) {
  return nt_tile_live(p, p.lens + p.G, row0, nrows);
}
constexpr int kNTLensSmem = 1024;  // groups whose T_eff is cached in shared memory by the TMA kernel

// write 4 consecutive k (or m) values as bf16 hi / lo into the two planes of a swizzled tile
template <bool SPLIT>
__device__ __forceinline__ void store_split4(unsigned char* hi_plane, unsigned char* lo_plane, uint32_t off, float4 v) {
  uint32_t h0, h1, l0 = 0u, l1 = 0u;
  if constexpr (SPLIT) {
    split_bf16(v.x, v.y, h0, l0);
    split_bf16(v.z, v.w, h1, l1);
  } else {
    h0 = pack_bf16(v.x, v.y);
    h1 = pack_bf16(v.z, v.w);
  }
  *reinterpret_cast<uint2*>(hi_plane + off) = make_uint2(h0, h1);
  if constexpr (SPLIT) *reinterpret_cast<uint2*>(lo_plane + off) = make_uint2(l0, l1);
}

// ------------------------------------------------------------------------------------------------------------------------------
// NT.  13 warps: 0-7 loaders (two groups of four, group g owns stage g so two tiles are in flight), 8-11 epilogue (TMEM ->
// smem transpose -> coalesced HBM stores), 12 MMA issuer + TMEM owner.  W (all of K) is resident in smem.
// ------------------------------------------------------------------------------------------------------------------------------
template <int NC, bool SPLIT>
__global__ void __launch_bounds__(416, 1) gemm_nt_tc_kernel(const GemmNTArgs p) {
  constexpr int NPART = SPLIT ? 2 : 1;
  constexpr uint32_t kWTile = NC * 128;  // one [NC x 64] bf16 plane of W
  constexpr uint32_t kTmemCols = 2 * NC < 32 ? 32 : 2 * NC;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int kslices = p.K / kBK, KC = p.nsrc * kslices;
  unsigned char* Wres = smem;                                    // [KC][NPART][NC x 128 B]
  unsigned char* Ast = Wres + (size_t)KC * NPART * kWTile;        // [stage][NPART][128 x 128 B]
  float* Est = reinterpret_cast<float*>(Ast + (size_t)kStagesNT * NPART * kTileBytes);  // [4 warps][32 rows][36] epilogue transpose
  NTBarriers* bars = reinterpret_cast<NTBarriers*>(Est + 4 * 32 * 36);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long nrows = (long long)p.G * p.B * p.Tmax;
  const int ntiles = (int)((nrows + kBM - 1) / kBM);

  if (tid == 0) {
    for (int s = 0; s < kStagesNT; ++s) {
      mbar_init(&bars->full[s], 128);
      mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tfull[a], 1);
      mbar_init(&bars->tempty[a], 128);
    }
    mbar_init_fence();
  }
  if (warp == 12) tmem_alloc(&bars->tmem_base, kTmemCols);
  // resident W: all 256 loader+epilogue threads convert fp32 -> bf16 hi/lo, swizzled K-major
  if (warp < 8) {
    const int f4_per_row = p.K / 4;
    for (int src = 0; src < p.nsrc; ++src) {
      const float* __restrict__ W = src ? p.W[1] : p.W[0];
      for (int i = tid; i < NC * f4_per_row; i += 256) {
        const int n = i / f4_per_row, k4 = i % f4_per_row, k = k4 * 4;
        const float4 v = *reinterpret_cast<const float4*>(W + (size_t)n * p.K + k);
        const int kc = src * kslices + k / kBK, kk = k % kBK;
        unsigned char* hi = Wres + (size_t)(kc * NPART) * kWTile;
        store_split4<SPLIT>(hi, hi + kWTile, sw128_offset(n, kk >> 3) + ((kk >> 2) & 1) * 8, v);
      }
    }
    fence_async_smem();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp < 8) {
    // ===================== loaders =====================
    const int grp = warp >> 2, tg = tid & 127;  // loader group (= the stage it fills), thread index inside the group
    const int f = tg & 15, rg = tg >> 4;        // float4 index within the 64-wide k slice, row group
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long row0 = (long long)tile * kBM;
      if (!nt_tile_live(p, row0, nrows)) continue;
      // row validity of the 16 rows this thread touches
      uint32_t vmask = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const long long row = row0 + j * 8 + rg;
        bool ok = row < nrows;
        if (ok) ok = (int)(row % p.Tmax) < p.lens[p.G + (int)(row / p.Tmax) / p.B];
        vmask |= (ok ? 1u : 0u) << j;
      }
      for (int kc = 0; kc < KC; ++kc, ++it) {
        const int stage = it % kStagesNT;
        if (stage != grp) continue;
        const float* __restrict__ A = (kc / kslices) ? p.A[1] : p.A[0];
        const int k0 = (kc % kslices) * kBK + f * 4;
        float4 v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (vmask & (1u << j)) v[j] = __ldg(reinterpret_cast<const float4*>(A + (row0 + j * 8 + rg) * p.lda + k0));
        }
        mbar_wait(&bars->empty[stage], ((it / kStagesNT) & 1) ^ 1);
        unsigned char* hi = Ast + (size_t)(stage * NPART) * kTileBytes;
#pragma unroll
        for (int j = 0; j < 16; ++j) store_split4<SPLIT>(hi, hi + kTileBytes, sw128_offset(j * 8 + rg, f >> 1) + (f & 1) * 8, v[j]);
        fence_async_smem();
        mbar_arrive(&bars->full[stage]);
      }
    }
  } else if (warp == 12) {
    // ===================== MMA issuer =====================
    {  // (whole warp, uniform control flow; one elected lane issues: tc05.cuh)
      constexpr uint32_t idesc = idesc_bf16(kBM, NC, false, false);
      uint32_t it = 0, tl = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (!nt_tile_live(p, (long long)tile * kBM, nrows)) continue;
        const uint32_t acc = tl & 1;
        mbar_wait(&bars->tempty[acc], ((tl >> 1) & 1) ^ 1);
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * NC;
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const int stage = it % kStagesNT;
          mbar_wait(&bars->full[stage], (it / kStagesNT) & 1);
          fence_after_sync();
          const uint32_t a_hi = smem_u32(Ast + (size_t)(stage * NPART) * kTileBytes), a_lo = a_hi + kTileBytes;
          const uint32_t b_hi = smem_u32(Wres + (size_t)(kc * NPART) * kWTile), b_lo = b_hi + kWTile;
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
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long row0 = (long long)tile * kBM;
      if (!nt_tile_live(p, row0, nrows)) continue;
      const uint32_t acc = tl & 1;
      mbar_wait(&bars->tfull[acc], (tl >> 1) & 1);
      fence_after_sync();
      const long long rbase = row0 + q * 32;
      bool ok = rbase + lane < nrows;
      if (ok) ok = (int)((rbase + lane) % p.Tmax) < p.lens[p.G + (int)((rbase + lane) / p.Tmax) / p.B];
      const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
#pragma unroll 1
      for (int c = 0; c < NC / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * NC + c * 32, r);
#pragma unroll
        for (int j = 0; j < 8; ++j)  // thread = row: stage the 32 columns (row stride 36 floats: conflict-free both ways)
          *reinterpret_cast<uint4*>(est + lane * 36 + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        __syncwarp();
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(p.bias + c * 32 + tc4));
#pragma unroll
        for (int itr = 0; itr < 8; ++itr) {  // 8 lanes cover one row's 32 columns: every store instruction writes 4 full 128-byte rows
          const int rr = itr * 4 + tr;
          if (okmask & (1u << rr)) {
            float4 o = *reinterpret_cast<const float4*>(est + rr * 36 + tc4);
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            float4* dst = reinterpret_cast<float4*>(p.C + (rbase + rr) * p.ldc + c * 32 + tc4);
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
  if (warp == 12) tmem_dealloc(tmem_base, kTmemCols);
}

template <int NC>
cudaError_t launch_nt_tc(const GemmNTArgs& a, int precision, cudaStream_t st) {
  const int npart = precision == 0 ? 2 : 1;
  const int KC = a.nsrc * (a.K / kBK);
  const size_t smem = 1024 + (size_t)KC * npart * NC * 128 + (size_t)kStagesNT * npart * kTileBytes + 4 * 32 * 36 * sizeof(float) +
                      sizeof(NTBarriers) + 64;
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;  // caller falls back (one source per pass / legacy kernel)
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long nrows = (long long)a.G * a.B * a.Tmax;
  const int ntiles = (int)((nrows + kBM - 1) / kBM);
  const unsigned grid = (unsigned)std::min(ntiles, sms);
  cudaError_t e;
  if (precision == 0) {
    e = cudaFuncSetAttribute(gemm_nt_tc_kernel<NC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    gemm_nt_tc_kernel<NC, true><<<grid, 416, smem, st>>>(a);
  } else {
    e = cudaFuncSetAttribute(gemm_nt_tc_kernel<NC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    gemm_nt_tc_kernel<NC, false><<<grid, 416, smem, st>>>(a);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------------------
// TN.  P[cta][KA=256, NB] = sum over this CTA's token rows of A[row, :]^T Bop[row, :].  Both operands are "MN-major" for the
// tensor core: a stage holds 64 token rows (= K); each 64-column block of A / Bop is a [64 k-rows x 128 B] swizzled tile.
// 9 warps: 0-7 loaders (fp32 -> bf16 hi/lo, masked / shifted / gathered), 8 MMA issuer + TMEM owner; warps 0-3 drain TMEM at the end.
// ------------------------------------------------------------------------------------------------------------------------------
constexpr int kStagesTN = 2;
constexpr int kBlkBytes = 64 * 128;  // one [64 k-rows x 64 columns] bf16 block

struct TNBarriers {
  uint64_t full[kStagesTN], empty[kStagesTN], done;
  uint32_t tmem_base;
};

template <int NB, bool SPLIT>
__global__ void __launch_bounds__(288, 1) gemm_tn_tc_kernel(const GemmTNArgs p) {
  constexpr int KA = 256, NPART = SPLIT ? 2 : 1;
  constexpr int kABytes = (KA / 64) * kBlkBytes, kBBytes = (NB / 64) * kBlkBytes;  // per plane
  constexpr int kStageBytes = NPART * (kABytes + kBBytes);
  constexpr uint32_t kTmemCols = 2 * NB;
  constexpr int FPR = NB / 4, RPP = 256 / FPR, BPASS = 64 / RPP;  // B operand: float4 per row, rows per pass, passes
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  TNBarriers* bars = reinterpret_cast<TNBarriers*>(smem + (size_t)kStagesTN * kStageBytes);
  float* csum_sm = reinterpret_cast<float*>(bars + 1);  // [4][KA] column-sum exchange
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = blockIdx.y, cta = blockIdx.x;
  const int T = p.lens[p.G + g];
  const int tiles_per_seq = (T + 63) / 64;
  const int items = p.B * tiles_per_seq;
  const bool gathered = p.tok != nullptr;

  if (tid == 0) {
    for (int s = 0; s < kStagesTN; ++s) {
      mbar_init(&bars->full[s], 256);
      mbar_init(&bars->empty[s], 1);
    }
    mbar_init(&bars->done, 1);
    mbar_init_fence();
  }
  if (warp == 8) tmem_alloc(&bars->tmem_base, kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  const int my_items = cta < items ? (items - cta + p.ctas_per_group - 1) / p.ctas_per_group : 0;

  if (warp < 8) {
    // ===================== loaders =====================
    const int fa = tid & 63, ra = tid >> 6;      // A: float4 index in the 256-wide row, first row (rows ra + 4j)
    const int fb = tid % FPR, rb = tid / FPR;    // B: float4 index in the NB-wide row, first row (rows rb + RPP*j)
    const uint32_t a_off = (fa >> 4) * kBlkBytes + (fa & 1) * 8, a_chunk = (fa & 15) >> 1;
    const uint32_t b_off = (fb >> 4) * kBlkBytes + (fb & 1) * 8, b_chunk = (fb & 15) >> 1;
    float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t it = 0;
    for (int item = cta; item < items; item += p.ctas_per_group, ++it) {
      const int stage = it % kStagesTN;
      const int n = g * p.B + item / tiles_per_seq, t0 = (item % tiles_per_seq) * 64;
      float4 va[16], vb[BPASS];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int t = t0 + ra + 4 * j;
        va[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < T) va[j] = __ldg(reinterpret_cast<const float4*>(p.A + ((size_t)n * p.Tmax + t) * KA) + fa);
      }
#pragma unroll
      for (int j = 0; j < BPASS; ++j) {
        const int t = t0 + rb + RPP * j;
        vb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < T) {
          const int col = fb * 4;
          if (col >= p.NB1) {  // second dense source (h of the previous scan position)
            const int ts = t + p.shift2;
            if (ts >= 0 && ts < T) vb[j] = __ldg(reinterpret_cast<const float4*>(p.Bsrc2 + ((size_t)n * p.Tmax + ts) * p.ldb2 + p.col02 + (col - p.NB1)));
          } else if (gathered) {
            const int tk = p.tok[(size_t)n * p.Tmax + t];
            const float sc = p.emb_row_scale != nullptr ? p.emb_row_scale[(size_t)g * p.V + tk] : 1.0f;
            const float4 e = __ldg(reinterpret_cast<const float4*>(p.emb + (size_t)tk * p.NB1 + col));
            vb[j] = make_float4(sc * e.x, sc * e.y, sc * e.z, sc * e.w);
          } else {
            const int ts = t + p.shift;
            if (ts >= 0 && ts < T) vb[j] = __ldg(reinterpret_cast<const float4*>(p.Bsrc + ((size_t)n * p.Tmax + ts) * p.ldb + p.col0 + col));
          }
        }
      }
      if (p.colsum) {
#pragma unroll
        for (int j = 0; j < 16; ++j) { csum.x += va[j].x; csum.y += va[j].y; csum.z += va[j].z; csum.w += va[j].w; }
      }
      mbar_wait(&bars->empty[stage], ((it / kStagesTN) & 1) ^ 1);
      unsigned char* st_base = smem + (size_t)stage * kStageBytes;
      unsigned char* a_hi = st_base;                     // [A hi][A lo][B hi][B lo]
      unsigned char* b_hi = st_base + NPART * kABytes;
#pragma unroll
      for (int j = 0; j < 16; ++j) store_split4<SPLIT>(a_hi, a_hi + kABytes, a_off + sw128_offset(ra + 4 * j, a_chunk), va[j]);
#pragma unroll
      for (int j = 0; j < BPASS; ++j) store_split4<SPLIT>(b_hi, b_hi + kBBytes, b_off + sw128_offset(rb + RPP * j, b_chunk), vb[j]);
      fence_async_smem();
      mbar_arrive(&bars->full[stage]);
    }
    if (p.colsum) *reinterpret_cast<float4*>(csum_sm + ra * KA + fa * 4) = csum;
  } else if (lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = idesc_bf16(128, NB, true, true);
    for (int it = 0; it < my_items; ++it) {
      const int stage = it % kStagesTN;
      mbar_wait(&bars->full[stage], (it / kStagesTN) & 1);
      fence_after_sync();
      const uint32_t a_hi = smem_u32(smem + (size_t)stage * kStageBytes), a_lo = a_hi + kABytes;
      const uint32_t b_hi = a_hi + NPART * kABytes, b_lo = b_hi + kBBytes;
#pragma unroll
      for (int k16 = 0; k16 < 4; ++k16) {
        const uint32_t ko = k16 * 16 * 128;  // 16 k-rows of 128 bytes
        const uint64_t bh = smem_desc_sw128(b_hi + ko, 1024, kBlkBytes), bl = smem_desc_sw128(b_lo + ko, 1024, kBlkBytes);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint32_t mo = mt * 2 * kBlkBytes;  // 128 columns of A = two 64-column blocks
          const uint64_t ah = smem_desc_sw128(a_hi + mo + ko, 1024, kBlkBytes);
          const uint32_t d = tmem_base + mt * NB;
          mma_bf16_ss(d, ah, bh, idesc, (it | k16) != 0);
          if constexpr (SPLIT) {
            const uint64_t al = smem_desc_sw128(a_lo + mo + ko, 1024, kBlkBytes);
            mma_bf16_ss(d, ah, bl, idesc, true);
            mma_bf16_ss(d, al, bh, idesc, true);
          }
        }
      }
      mma_commit(&bars->empty[stage]);
    }
    mma_commit(&bars->done);
  }
  __syncthreads();  // column sums visible; every role has finished issuing

  float* out = p.partial + ((size_t)g * p.ctas_per_group + cta) * ((size_t)KA * NB + (p.colsum ? KA : 0));
  if (warp < 4) {
    // ===================== epilogue: TMEM -> partial[KA][NB] =====================
    if (my_items > 0) {
      mbar_wait(&bars->done, 0);
      fence_after_sync();
    }
#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) {
      float* orow = out + (size_t)(mt * 128 + warp * 32 + lane) * NB;
#pragma unroll 1
      for (int c = 0; c < NB / 32; ++c) {
        uint32_t r[32];
        if (my_items > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + mt * NB + c * 32, r);
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
  } else if (warp < 8 && p.colsum) {
    for (int c = tid - 128; c < KA; c += 128)
      out[(size_t)KA * NB + c] = my_items > 0 ? (csum_sm[c] + csum_sm[KA + c]) + (csum_sm[2 * KA + c] + csum_sm[3 * KA + c]) : 0.f;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, kTmemCols);
}

template <int NB>
cudaError_t launch_tn_tc(const GemmTNArgs& a, int precision, cudaStream_t st) {
  const int npart = precision == 0 ? 2 : 1;
  const size_t smem = 1024 + (size_t)kStagesTN * npart * ((256 / 64) + (NB / 64)) * kBlkBytes + sizeof(TNBarriers) + 4 * 256 * sizeof(float) + 64;
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  dim3 grid(a.ctas_per_group, a.G);
  cudaError_t e;
  if (precision == 0) {
    e = cudaFuncSetAttribute(gemm_tn_tc_kernel<NB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    gemm_tn_tc_kernel<NB, true><<<grid, 288, smem, st>>>(a);
  } else {
    e = cudaFuncSetAttribute(gemm_tn_tc_kernel<NB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    gemm_tn_tc_kernel<NB, false><<<grid, 288, smem, st>>>(a);
  }
  return cudaGetLastError();
}

// ==============================================================================================================================
// TMA-fed variants.  The recurrent kernels store Y and the dgates as bf16 hi/lo PLANES (interleaved per row over the bytes of
// the fp32 row), so the GEMM operands need no conversion: one thread issues cp.async.bulk.tensor loads straight into the
// 128-byte-swizzled tiles, completion is counted on the stage's mbarrier (expect_tx), and the loader warps disappear.
// ==============================================================================================================================
struct NTTmaMaps {
  CUtensorMap a[2][2];  // [source][plane]: 2D {K, rows}, box {64, 128}
  CUtensorMap c;        // TSTORE: the fp32 result, 2D {NC, rows}, box {32, 32}, 128-byte swizzle
};
constexpr int kNTStoreTile = 32 * 128;  // TSTORE: one staging tile [32 rows x 32 floats]; 4 epilogue warps x (1 or 2) tiles

// TSTORE (C = ..., not +=; chosen by the launcher only where two staging tiles per epilogue warp fit next to the resident W -- at the
// headline shapes, with 128 KB of W, they do not, and a single tile measured slower than the thread stores): the epilogue writes 32x32 blocks into swizzled staging tiles and hands them to the TMA store unit
// (cp.async.bulk.tensor, double buffered per warp) instead of transposing through shared memory and storing row by row from the
// threads.  Whole 32-row boxes are written: rows t >= T_eff of a live tile receive values nobody reads (the consumers of the input
// projection and of dY only touch rows t < T_eff).
template <int NC, bool SPLIT, bool TSTORE>
__global__ void __launch_bounds__(192, 1) gemm_nt_tma_kernel(const __grid_constant__ NTTmaMaps maps, const GemmNTArgs p, const int nbuf) {
  constexpr int NPART = SPLIT ? 2 : 1;
  constexpr uint32_t kWTile = NC * 128;
  constexpr uint32_t kTmemCols = 2 * NC < 32 ? 32 : 2 * NC;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int kslices = p.K / kBK, KC = p.nsrc * kslices;
  unsigned char* Wres = smem;
  unsigned char* Ast = Wres + (size_t)KC * NPART * kWTile;
  float* Est = reinterpret_cast<float*>(Ast + (size_t)kStagesNT * NPART * kTileBytes);
  NTBarriers* bars = reinterpret_cast<NTBarriers*>(reinterpret_cast<unsigned char*>(Est) + (TSTORE ? 4 * nbuf * kNTStoreTile : 4 * 32 * 36 * (int)sizeof(float)));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long nrows = (long long)p.G * p.B * p.Tmax;
  const int ntiles = (int)((nrows + kBM - 1) / kBM);
  __shared__ int teff_s[kNTLensSmem];
  __shared__ __align__(16) float bias_s[TSTORE ? NC : 4];
  if constexpr (TSTORE) {
    for (int i = tid; i < NC; i += 192) bias_s[i] = p.bias != nullptr ? p.bias[i] : 0.f;
    if (tid == 0) tma_prefetch_desc(&maps.c);
  }
  const int* teff = p.G <= kNTLensSmem ? teff_s : p.lens + p.G;
  for (int i = tid; i < min(p.G, kNTLensSmem); i += 192) teff_s[i] = p.lens[p.G + i];

  if (tid == 0) {
    for (int s = 0; s < kStagesNT; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tfull[a], 1);
      mbar_init(&bars->tempty[a], 128);
    }
    mbar_init_fence();
    for (int sidx = 0; sidx < p.nsrc; ++sidx)
      for (int pl = 0; pl < NPART; ++pl) tma_prefetch_desc(&maps.a[sidx][pl]);
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, kTmemCols);
  {  // resident W: fp32 -> bf16 hi/lo, swizzled K-major (tiny; every thread helps)
    const int f4_per_row = p.K / 4;
    for (int src = 0; src < p.nsrc; ++src) {
      const float* __restrict__ W = src ? p.W[1] : p.W[0];
#pragma unroll 8  // (8 loads in flight per thread: a rolled loop pays one L2 round trip per 16 bytes)
      for (int i = tid; i < NC * f4_per_row; i += 192) {
        const int n = i / f4_per_row, k4 = i % f4_per_row, k = k4 * 4;
        const float4 v = *reinterpret_cast<const float4*>(W + (size_t)n * p.K + k);
        const int kc = src * kslices + k / kBK, kk = k % kBK;
        unsigned char* hi = Wres + (size_t)(kc * NPART) * kWTile;
        store_split4<SPLIT>(hi, hi + kWTile, sw128_offset(n, kk >> 3) + ((kk >> 2) & 1) * 8, v);
      }
    }
    fence_async_smem();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long row0 = (long long)tile * kBM;
        if (!nt_tile_live(p, teff, row0, nrows)) continue;
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const int stage = it % kStagesNT, src = kc / kslices, k0 = (kc % kslices) * kBK;
          mbar_wait(&bars->empty[stage], ((it / kStagesNT) & 1) ^ 1);
          mbar_arrive_expect_tx(&bars->full[stage], NPART * kTileBytes);
          unsigned char* dst = Ast + (size_t)(stage * NPART) * kTileBytes;
          tma_load_2d(dst, &maps.a[src][0], &bars->full[stage], k0, (int)row0);
          if constexpr (SPLIT) tma_load_2d(dst + kTileBytes, &maps.a[src][1], &bars->full[stage], k0, (int)row0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_bf16(kBM, NC, false, false);
      uint32_t it = 0, tl = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (!nt_tile_live(p, teff, (long long)tile * kBM, nrows)) continue;
        const uint32_t acc = tl & 1;
        mbar_wait(&bars->tempty[acc], ((tl >> 1) & 1) ^ 1);
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * NC;
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const int stage = it % kStagesNT;
          mbar_wait(&bars->full[stage], (it / kStagesNT) & 1);
          fence_after_sync();
          const uint32_t a_hi = smem_u32(Ast + (size_t)(stage * NPART) * kTileBytes), a_lo = a_hi + kTileBytes;
          const uint32_t b_hi = smem_u32(Wres + (size_t)(kc * NPART) * kWTile), b_lo = b_hi + kWTile;
#pragma unroll
          for (int k16 = 0; k16 < kBK / 16; ++k16) {
            const uint32_t ko = k16 * 32;
            const uint64_t ah = smem_desc_sw128(a_hi + ko, 1024, 0), bh = smem_desc_sw128(b_hi + ko, 1024, 0);
            mma_bf16_ss(d_tmem, ah, bh, idesc, (kc | k16) != 0);
            if constexpr (SPLIT) {
              const uint64_t al = smem_desc_sw128(a_lo + ko, 1024, 0), bl = smem_desc_sw128(b_lo + ko, 1024, 0);
              mma_bf16_ss(d_tmem, ah, bl, idesc, true);
              mma_bf16_ss(d_tmem, al, bh, idesc, true);
            }
          }
          mma_commit(&bars->empty[stage]);
        }
        mma_commit(&bars->tfull[acc]);
        ++tl;
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;
    if constexpr (TSTORE) {
      // TMEM -> registers (+ bias) -> swizzled 32x32 staging tile -> TMA store; two staging tiles per warp
      unsigned char* stage = reinterpret_cast<unsigned char*>(Est) + q * (nbuf * kNTStoreTile);
      uint32_t tl = 0, nst = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long row0 = (long long)tile * kBM;
        if (!nt_tile_live(p, teff, row0, nrows)) continue;
        const uint32_t acc = tl & 1;
        mbar_wait(&bars->tfull[acc], (tl >> 1) & 1);
        fence_after_sync();
        const int rbase = (int)(row0 + q * 32);
        bool ok = row0 + q * 32 + lane < nrows;
        if (ok) ok = (int)((row0 + q * 32 + lane) % p.Tmax) < teff[(int)((row0 + q * 32 + lane) / p.Tmax) / p.B];
        const bool live_box = __ballot_sync(0xffffffffu, ok) != 0u;  // a 32-row box without a valid row is not stored
        uint32_t r[2][32];
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * NC;
        if (live_box) tmem_ld32_nowait(t_row, r[0]);
#pragma unroll
        for (int c = 0; c < NC / 32; ++c) {
          if (!live_box) break;
          tmem_wait_ld();
          if (c + 1 < NC / 32) tmem_ld32_nowait(t_row + (c + 1) * 32, r[(c + 1) & 1]);
          unsigned char* buf = stage + (nbuf == 2 ? (nst & 1) : 0) * kNTStoreTile;
          if (lane == 0) {  // the store issued from this buffer (two blocks ago with two buffers) has read it
            if (nbuf == 2) bulk_wait_group_read<1>();
            else bulk_wait_group_read<0>();
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = *reinterpret_cast<const float4*>(bias_s + c * 32 + 4 * j);
            float4 o;
            o.x = __uint_as_float(r[c & 1][4 * j]) + b.x;
            o.y = __uint_as_float(r[c & 1][4 * j + 1]) + b.y;
            o.z = __uint_as_float(r[c & 1][4 * j + 2]) + b.z;
            o.w = __uint_as_float(r[c & 1][4 * j + 3]) + b.w;
            *reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) = o;
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&maps.c, buf, c * 32, rbase);
            bulk_commit_group();
          }
          ++nst;
        }
        fence_before_sync();
        mbar_arrive(&bars->tempty[acc]);
        ++tl;
      }
      if (lane == 0) bulk_wait_group_read<0>();
    } else {
    float* est = Est + q * 32 * 36;
    const int tr = lane >> 3, tc4 = (lane & 7) * 4;
    float4 bias4[NC / 32];
#pragma unroll
    for (int c = 0; c < NC / 32; ++c)
      bias4[c] = p.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(p.bias + c * 32 + tc4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long row0 = (long long)tile * kBM;
      if (!nt_tile_live(p, teff, row0, nrows)) continue;
      const uint32_t acc = tl & 1;
      mbar_wait(&bars->tfull[acc], (tl >> 1) & 1);
      fence_after_sync();
      const long long rbase = row0 + q * 32;
      bool ok = rbase + lane < nrows;
      if (ok) ok = (int)((rbase + lane) % p.Tmax) < teff[(int)((rbase + lane) / p.Tmax) / p.B];
      const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
      // TMEM loads are software-pipelined over the 32-column chunks: chunk c+1 is in flight while chunk c goes through the
      // transpose and out to global memory (the bias of this thread's columns was loaded once, before the tile loop)
      uint32_t r[2][32];
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * NC;
      tmem_ld32_nowait(t_row, r[0]);
#pragma unroll
      for (int c = 0; c < NC / 32; ++c) {
        tmem_wait_ld();
        if (c + 1 < NC / 32) tmem_ld32_nowait(t_row + (c + 1) * 32, r[(c + 1) & 1]);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(est + lane * 36 + 4 * j) =
              make_uint4(r[c & 1][4 * j], r[c & 1][4 * j + 1], r[c & 1][4 * j + 2], r[c & 1][4 * j + 3]);
        __syncwarp();
        const float4 b = bias4[c];
#pragma unroll
        for (int itr = 0; itr < 8; ++itr) {
          const int rr = itr * 4 + tr;
          if (okmask & (1u << rr)) {
            float4 o = *reinterpret_cast<const float4*>(est + rr * 36 + tc4);
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            float4* dst = reinterpret_cast<float4*>(p.C + (rbase + rr) * p.ldc + c * 32 + tc4);
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
    }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

struct TNTmaMaps {
  CUtensorMap a[2];   // dgates planes: 3D {4H, Tmax, N}, box {64, 64, 1}
  CUtensorMap b[2];   // first B source planes (unused when gathered)
  CUtensorMap b2[2];  // second B source planes (NB2 > 0): fuses dW_ih and dW_hh into one pass over the dgates
};

// B operand = [first source: NB1 columns (dense planes, or gathered masked embeddings) | second source: NB2 columns (dense planes)]
template <int NB1, int NB2, bool SPLIT, bool GATHER>
__global__ void __launch_bounds__(192, 1) gemm_tn_tma_kernel(const __grid_constant__ TNTmaMaps maps, const GemmTNArgs p) {
  constexpr int KA = 256, NPART = SPLIT ? 2 : 1, NB = NB1 + NB2;
  constexpr int kABytes = (KA / 64) * kBlkBytes, kBBytes = (NB / 64) * kBlkBytes, kB2Bytes = (NB2 / 64) * kBlkBytes;
  constexpr int kStageBytes = NPART * (kABytes + kBBytes);
  constexpr uint32_t kTmemCols = 2 * NB <= 256 ? 2 * NB : 512;  // power of two
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  TNBarriers* bars = reinterpret_cast<TNBarriers*>(smem + (size_t)kStagesTN * kStageBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = blockIdx.y, cta = blockIdx.x;
  const int T = p.lens[p.G + g];
  const int tiles_per_seq = (T + 63) / 64;
  const int items = p.B * tiles_per_seq;

  if (tid == 0) {
    for (int s = 0; s < kStagesTN; ++s) {
      mbar_init(&bars->full[s], GATHER ? 1 + 128 : 1);
      mbar_init(&bars->empty[s], 1);
    }
    mbar_init(&bars->done, 1);
    mbar_init_fence();
    for (int pl = 0; pl < NPART; ++pl) {
      tma_prefetch_desc(&maps.a[pl]);
      if (!GATHER) tma_prefetch_desc(&maps.b[pl]);
      if (NB2 > 0) tma_prefetch_desc(&maps.b2[pl]);
    }
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  const int my_items = cta < items ? (items - cta + p.ctas_per_group - 1) / p.ctas_per_group : 0;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = cta; item < items; item += p.ctas_per_group, ++it) {
        const int stage = it % kStagesTN;
        const int n = g * p.B + item / tiles_per_seq, t0 = (item % tiles_per_seq) * 64;
        mbar_wait(&bars->empty[stage], ((it / kStagesTN) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->full[stage], NPART * (kABytes + (GATHER ? kB2Bytes : kBBytes)));
        unsigned char* a_dst = smem + (size_t)stage * kStageBytes;
        unsigned char* b_dst = a_dst + NPART * kABytes;
#pragma unroll
        for (int pl = 0; pl < NPART; ++pl) {
#pragma unroll
          for (int blk = 0; blk < KA / 64; ++blk)
            tma_load_3d(a_dst + pl * kABytes + blk * kBlkBytes, &maps.a[pl], &bars->full[stage], blk * 64, t0, n);
          if constexpr (!GATHER) {
#pragma unroll
            for (int blk = 0; blk < NB1 / 64; ++blk)
              tma_load_3d(b_dst + pl * kBBytes + blk * kBlkBytes, &maps.b[pl], &bars->full[stage], p.col0 + blk * 64, t0 + p.shift, n);
          }
#pragma unroll
          for (int blk = 0; blk < NB2 / 64; ++blk)
            tma_load_3d(b_dst + pl * kBBytes + (NB1 / 64 + blk) * kBlkBytes, &maps.b2[pl], &bars->full[stage], p.col02 + blk * 64,
                        t0 + p.shift2, n);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    {  // (whole warp, uniform control flow; one elected lane issues: tc05.cuh)
      constexpr uint32_t idesc = idesc_bf16(128, NB, true, true);
      for (int it = 0; it < my_items; ++it) {
        const int stage = it % kStagesTN;
        mbar_wait(&bars->full[stage], (it / kStagesTN) & 1);
        fence_after_sync();
        const uint32_t a_hi = smem_u32(smem + (size_t)stage * kStageBytes), a_lo = a_hi + kABytes;
        const uint32_t b_hi = a_hi + NPART * kABytes, b_lo = b_hi + kBBytes;
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) {
          const uint32_t ko = k16 * 16 * 128;
          const uint64_t bh = smem_desc_sw128(b_hi + ko, 1024, kBlkBytes), bl = smem_desc_sw128(b_lo + ko, 1024, kBlkBytes);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const uint32_t mo = mt * 2 * kBlkBytes;
            const uint64_t ah = smem_desc_sw128(a_hi + mo + ko, 1024, kBlkBytes);
            const uint32_t d = tmem_base + mt * NB;
            mma_bf16_ss_elect(d, ah, bh, idesc, (it | k16) != 0);
            if constexpr (SPLIT) {
              const uint64_t al = smem_desc_sw128(a_lo + mo + ko, 1024, kBlkBytes);
              mma_bf16_ss_elect(d, ah, bl, idesc, true);
              mma_bf16_ss_elect(d, al, bh, idesc, true);
            }
          }
        }
        mma_commit_elect(&bars->empty[stage]);
      }
      mma_commit_elect(&bars->done);
    }
  } else if (GATHER) {
    // ===================== B operand = scale[g][tok] * emb[tok] (layer-0 input), staged by 4 warps =====================
    constexpr int FPR = NB1 / 4, RPP = 128 / FPR, BPASS = 64 / RPP;
    const int tg = tid - 64, fb = tg % FPR, rb = tg / FPR;
    const uint32_t b_off = (fb >> 4) * kBlkBytes + (fb & 1) * 8, b_chunk = (fb & 15) >> 1;
    uint32_t it = 0;
    for (int item = cta; item < items; item += p.ctas_per_group, ++it) {
      const int stage = it % kStagesTN;
      const int n = g * p.B + item / tiles_per_seq, t0 = (item % tiles_per_seq) * 64;
      float4 vb[BPASS];
#pragma unroll
      for (int j = 0; j < BPASS; ++j) {
        const int t = t0 + rb + RPP * j;
        vb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < T) {
          const int tk = p.tok[(size_t)n * p.Tmax + t];
          const float sc = p.emb_row_scale != nullptr ? p.emb_row_scale[(size_t)g * p.V + tk] : 1.0f;
          const float4 e = __ldg(reinterpret_cast<const float4*>(p.emb + (size_t)tk * NB1) + fb);
          vb[j] = make_float4(sc * e.x, sc * e.y, sc * e.z, sc * e.w);
        }
      }
      mbar_wait(&bars->empty[stage], ((it / kStagesTN) & 1) ^ 1);
      unsigned char* b_hi = smem + (size_t)stage * kStageBytes + NPART * kABytes;
#pragma unroll
      for (int j = 0; j < BPASS; ++j) store_split4<SPLIT>(b_hi, b_hi + kBBytes, b_off + sw128_offset(rb + RPP * j, b_chunk), vb[j]);
      fence_async_smem();
      mbar_arrive(&bars->full[stage]);
    }
  }
  __syncthreads();

  float* out = p.partial + ((size_t)g * p.ctas_per_group + cta) * ((size_t)KA * NB);
  if (warp >= 2) {
    const int q = warp & 3;
    if (my_items > 0) {
      mbar_wait(&bars->done, 0);
      fence_after_sync();
    }
#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) {
      float* orow = out + (size_t)(mt * 128 + q * 32 + lane) * NB;
#pragma unroll 1
      for (int c = 0; c < NB / 32; ++c) {
        uint32_t r[32];
        if (my_items > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + mt * NB + c * 32, r);
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
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

bool nt_maps(const GemmNTArgs& a, int npart, int row_floats, NTTmaMaps* m) {
  // A_s points at the fp32-sized rows that hold the planes: hi at byte 0, lo at byte K*2; row pitch = lda floats
  const long long nrows = (long long)a.G * a.B * a.Tmax;
  (void)row_floats;
  for (int s = 0; s < a.nsrc; ++s)
    for (int pl = 0; pl < npart; ++pl) {
      const uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)nrows}, strides[1] = {(uint64_t)a.lda * 4};
      const uint32_t box[2] = {64, 128};
      if (!make_tmap_bf16_sw128(&m->a[s][pl], reinterpret_cast<const unsigned char*>(a.A[s]) + (size_t)pl * a.plane_bytes, 2, dims, strides, box))
        return false;
    }
  return true;
}

template <int NC>
cudaError_t launch_nt_tma(const GemmNTArgs& a, int precision, cudaStream_t st) {
  const int npart = precision == 0 ? 2 : 1;
  const int KC = a.nsrc * (a.K / kBK);
  static const bool no_tstore = getenv("IB200_NO_TMA_STORE") != nullptr;
  const size_t base_smem = 1024 + (size_t)KC * npart * NC * 128 + (size_t)kStagesNT * npart * kTileBytes + sizeof(NTBarriers) + 64;
  const size_t static_smem = kNTLensSmem * sizeof(int) + NC * sizeof(float) + 64, limit = 227 * 1024 - static_smem;
  const int nbuf = base_smem + 8 * kNTStoreTile <= limit ? 2 : 1;  // two staging tiles per epilogue warp when they fit
  // (with a single staging tile per warp the store and the next block's fill serialise: measured slower than the thread stores)
  bool tstore = !no_tstore && !a.accumulate && a.ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(a.C) & 15) == 0 && nbuf == 2;
  const size_t smem = base_smem + (tstore ? (size_t)(4 * nbuf * kNTStoreTile) : 4 * 32 * 36 * sizeof(float));
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  NTTmaMaps maps;
  if (!nt_maps(a, npart, 0, &maps)) return cudaErrorInvalidConfiguration;
  if (tstore) {
    const uint64_t dc[2] = {(uint64_t)NC, (uint64_t)((long long)a.G * a.B * a.Tmax)}, sc[1] = {(uint64_t)a.ldc * 4};
    const uint32_t bc[2] = {32, 32};
    if (!make_tmap_f32_sw128(&maps.c, a.C, 2, dc, sc, bc)) tstore = false;
  }
  if (!tstore && smem != base_smem + 4 * 32 * 36 * sizeof(float)) return cudaErrorInvalidConfiguration;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long nrows = (long long)a.G * a.B * a.Tmax;
  const int ntiles = (int)((nrows + kBM - 1) / kBM);
  const unsigned grid = (unsigned)std::min(ntiles, sms);
  cudaError_t e;
#define IB200_NT_LAUNCH(SP_, TS_)                                                                                          \
  {                                                                                                                       \
    e = cudaFuncSetAttribute(gemm_nt_tma_kernel<NC, SP_, TS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    if (e != cudaSuccess) return e;                                                                                       \
    gemm_nt_tma_kernel<NC, SP_, TS_><<<grid, 192, smem, st>>>(maps, a, nbuf);                                             \
  }
  if (precision == 0) {
    if (tstore) IB200_NT_LAUNCH(true, true) else IB200_NT_LAUNCH(true, false)
  } else {
    if (tstore) IB200_NT_LAUNCH(false, true) else IB200_NT_LAUNCH(false, false)
  }
#undef IB200_NT_LAUNCH
  return cudaGetLastError();
}

template <int NB1, int NB2, bool GATHER>
cudaError_t launch_tn_tma(const GemmTNArgs& a, int precision, cudaStream_t st) {
  constexpr int NB = NB1 + NB2;
  const int npart = precision == 0 ? 2 : 1;
  const size_t smem = 1024 + (size_t)kStagesTN * npart * ((256 / 64) + (NB / 64)) * kBlkBytes + sizeof(TNBarriers) + 64;
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  TNTmaMaps maps;
  const uint32_t box[3] = {64, 64, 1};
  for (int pl = 0; pl < npart; ++pl) {
    const uint64_t da[3] = {256, (uint64_t)a.Tmax, (uint64_t)a.G * a.B}, sa[2] = {1024, (uint64_t)a.Tmax * 1024};
    if (!make_tmap_bf16_sw128(&maps.a[pl], reinterpret_cast<const unsigned char*>(a.A) + (size_t)pl * 512, 3, da, sa, box))
      return cudaErrorInvalidConfiguration;
    if (!GATHER) {
      const uint64_t db[3] = {(uint64_t)a.ldb, (uint64_t)a.Tmax, (uint64_t)a.G * a.B};
      const uint64_t sb[2] = {(uint64_t)a.ldb * 4, (uint64_t)a.Tmax * a.ldb * 4};
      if (!make_tmap_bf16_sw128(&maps.b[pl], reinterpret_cast<const unsigned char*>(a.Bsrc) + (size_t)pl * a.ldb * 2, 3, db, sb, box))
        return cudaErrorInvalidConfiguration;
    }
    if (NB2 > 0) {
      const uint64_t db[3] = {(uint64_t)a.ldb2, (uint64_t)a.Tmax, (uint64_t)a.G * a.B};
      const uint64_t sb[2] = {(uint64_t)a.ldb2 * 4, (uint64_t)a.Tmax * a.ldb2 * 4};
      if (!make_tmap_bf16_sw128(&maps.b2[pl], reinterpret_cast<const unsigned char*>(a.Bsrc2) + (size_t)pl * a.ldb2 * 2, 3, db, sb, box))
        return cudaErrorInvalidConfiguration;
    }
  }
  dim3 grid(a.ctas_per_group, a.G);
  cudaError_t e;
  if (precision == 0) {
    e = cudaFuncSetAttribute(gemm_tn_tma_kernel<NB1, NB2, true, GATHER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    gemm_tn_tma_kernel<NB1, NB2, true, GATHER><<<grid, 192, smem, st>>>(maps, a);
  } else {
    e = cudaFuncSetAttribute(gemm_tn_tma_kernel<NB1, NB2, false, GATHER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    gemm_tn_tma_kernel<NB1, NB2, false, GATHER><<<grid, 192, smem, st>>>(maps, a);
  }
  return cudaGetLastError();
}

}  // namespace

// returns cudaErrorInvalidConfiguration when the shape is not covered (caller uses the legacy mma.sync kernel)
cudaError_t launch_gemm_nt_tc(const GemmNTArgs& a, int precision, cudaStream_t st) {
  if (a.K % kBK != 0 || a.lda % 4 != 0 || a.ldc % 4 != 0) return cudaErrorInvalidConfiguration;
  switch (a.NC) {
    case 256: return launch_nt_tc<256>(a, precision, st);
    case 128: return launch_nt_tc<128>(a, precision, st);
    case 64: return launch_nt_tc<64>(a, precision, st);
    default: return cudaErrorInvalidConfiguration;
  }
}

cudaError_t launch_gemm_tn_tc(const GemmTNArgs& a, int precision, cudaStream_t st) {
  if (a.KA != 256) return cudaErrorInvalidConfiguration;
  if (!a.tok && (a.ldb % 4 != 0 || a.col0 % 4 != 0)) return cudaErrorInvalidConfiguration;
  if (a.NB1 != a.NB && (a.NB1 % 4 != 0 || a.ldb2 % 4 != 0 || a.col02 % 4 != 0 || a.Bsrc2 == nullptr)) return cudaErrorInvalidConfiguration;
  if (a.NB == 128) return launch_tn_tc<128>(a, precision, st);
  if (a.NB == 64) return launch_tn_tc<64>(a, precision, st);
  return cudaErrorInvalidConfiguration;
}

// TMA-fed versions: operands are bf16 hi/lo planes (see kernels.h: GemmNTArgs::plane_bytes, GemmTNArgs planes layout)
cudaError_t launch_gemm_nt_tma(const GemmNTArgs& a, int precision, cudaStream_t st) {
  if (a.K % kBK != 0 || a.lda % 4 != 0 || a.ldc % 4 != 0 || a.plane_bytes <= 0) return cudaErrorInvalidConfiguration;
  switch (a.NC) {
    case 256: return launch_nt_tma<256>(a, precision, st);
    case 128: return launch_nt_tma<128>(a, precision, st);
    case 64: return launch_nt_tma<64>(a, precision, st);
    default: return cudaErrorInvalidConfiguration;
  }
}
cudaError_t launch_gemm_tn_tma(const GemmTNArgs& a, int precision, cudaStream_t st) {
  if (a.KA != 256 || a.colsum) return cudaErrorInvalidConfiguration;
  const bool gather = a.tok != nullptr;
  const int nb2 = a.NB - a.NB1;
  if (!gather && (a.ldb % 4 != 0 || a.col0 % 8 != 0)) return cudaErrorInvalidConfiguration;
  if (nb2 != 0 && (a.Bsrc2 == nullptr || a.ldb2 % 4 != 0 || a.col02 % 8 != 0)) return cudaErrorInvalidConfiguration;
  if (nb2 == 0) {
    if (a.NB == 128 && !gather) return launch_tn_tma<128, 0, false>(a, precision, st);
    if (a.NB == 64 && !gather) return launch_tn_tma<64, 0, false>(a, precision, st);
    if (a.NB == 64 && gather) return launch_tn_tma<64, 0, true>(a, precision, st);
  } else if (nb2 == 64) {  // dW_ih | dW_hh in one pass over the dgates
    if (a.NB1 == 128 && !gather) return launch_tn_tma<128, 64, false>(a, precision, st);
    if (a.NB1 == 64 && gather) return launch_tn_tma<64, 64, true>(a, precision, st);
  }
  return cudaErrorInvalidConfiguration;
}

}  // namespace ib200
