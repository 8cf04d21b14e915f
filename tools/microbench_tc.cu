// Microbenchmark + layout check for the tcgen05 recurrent step (B200 / sm_100a):
//   gates^T[4H=256, 16] = W_hh[256, 64] (A operand, RESIDENT IN TMEM as bf16 hi | lo) * [h_hi | h_lo]^T (B operand, smem, K-major)
// (1) verifies the TMEM A layout, the TS-mode MMA and the .16x256b load fragment against a host computation;
// (2) times the dependent round trip  STS h -> fence -> sync -> 16 x tcgen05.mma -> commit -> mbarrier -> tcgen05.ld.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/microbench_tc tools/microbench_tc.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../intrepppid_b200/csrc/tc05.cuh"

using namespace ib200::tc;

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
// 16 lanes x 256 bit, two repetitions (16 columns): the mma.m16n8 accumulator fragment, twice
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}

constexpr int H = 64, M = 256, NCOL = 16;
constexpr uint32_t kColAhi = 0, kColAlo = 64, kColD = 128, kTmemCols = 256;

struct Smem {
  alignas(1024) unsigned char B[2][2048];  // [buffer][16 rows x 128 B], 128B swizzle, K-major: row = column n of h^T, 64 k
  uint64_t bar;
  uint64_t hbar;
  uint32_t tmem_base;
};

// mode bit 0: issue MMAs; bit 1: only 8 MMAs (hi only); bit 2: fake activations (20 MUFU / thread); bit 3: mbarrier handoff with a
// dedicated MMA warp instead of __syncthreads; bit 4: no tcgen05.ld
__global__ void __launch_bounds__(288, 1) k_step(const float* __restrict__ W, const float* __restrict__ Hin, float* __restrict__ out,
                                                  long long* cycles, int steps, int mode) {
  extern __shared__ unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tig = lane & 3;
  const bool worker = warp < 8;
  if (tid == 0) {
    mbar_init(&sm.bar, 1);
    mbar_init(&sm.hbar, 256);
    mbar_init_fence();
  }
  if (warp == 8) tmem_alloc(&sm.tmem_base, kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = sm.tmem_base;
  const int half = warp >> 2, quad = warp & 3;

  // ---- A (W rows, MMA row order) -> TMEM: thread = row 128*half + 32*quad + lane; 64 k as 32 packed words, hi and lo
  if (worker) {
    const int row = 128 * half + 32 * quad + lane;
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = (c8 * 8 + j) * 2;
        const float w0 = W[row * H + k], w1 = W[row * H + k + 1];
        __nv_bfloat162 h2 = __floats2bfloat162_rn(w0, w1);
        float2 hf = __bfloat1622float2(h2);
        __nv_bfloat162 l2 = __floats2bfloat162_rn(w0 - hf.x, w1 - hf.y);
        hi[j] = *reinterpret_cast<uint32_t*>(&h2);
        lo[j] = *reinterpret_cast<uint32_t*>(&l2);
      }
      const uint32_t lane_addr = (uint32_t)(32 * quad) << 16;
      tmem_st8(tb + lane_addr + kColAhi + half * 32 + c8 * 8, hi);
      tmem_st8(tb + lane_addr + kColAlo + half * 32 + c8 * 8, lo);
    }
    tmem_st_wait();
  }
  // ---- initial B tile (buffer 0): h^T rows n = 0..15 (0-7 hi part, 8-15 lo part), k = unit
  for (int i = tid; i < NCOL * H; i += blockDim.x) {
    const int n = i / H, k = i % H;
    const float v = Hin[k * NCOL + n];
    *reinterpret_cast<__nv_bfloat16*>(sm.B[0] + sw128_offset(n, k >> 3) + (k & 7) * 2) = __float2bfloat16_rn(v);
    *reinterpret_cast<__nv_bfloat16*>(sm.B[1] + sw128_offset(n, k >> 3) + (k & 7) * 2) = __float2bfloat16_rn(v);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  constexpr uint32_t idesc = ib200::tc::idesc_bf16(128, NCOL, false, false);
  auto issue_mmas = [&](int buf) {
    const uint32_t b0 = smem_u32(sm.B[buf]);
    const int nparts = (mode & 2) ? 1 : 2;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const uint32_t d = tb + kColD + hf * NCOL;
      for (int part = 0; part < nparts; ++part) {
#pragma unroll
        for (int k16 = 0; k16 < H / 16; ++k16) {
          const uint64_t bd = smem_desc_sw128(b0 + k16 * 32, 1024, 0);
          mma_bf16_ts(d, tb + (part ? kColAlo : kColAhi) + hf * 32 + k16 * 8, bd, idesc, (part | k16) != 0);
        }
      }
    }
    mma_commit(&sm.bar);
  };

  // the unit / columns of this worker thread (same mapping as the mma.sync kernels)
  const int u = 32 * half + 8 * quad + gq;
  float c0 = 0.1f, c1 = 0.2f;
  uint32_t r0[8] = {}, r1[8] = {};
  long long t0 = 0;
  if (tid == 0 && (mode & 1)) {
    fence_after_sync();
    issue_mmas(0);
  }
  __syncthreads();
  t0 = clock64();
  for (int s = 0; s < steps; ++s) {
    const int nb = (s + 1) & 1;
    if (worker) {
      if (mode & 1) {
        mbar_wait(&sm.bar, s & 1);
        fence_after_sync();
        if (!(mode & 16)) {
          const uint32_t la = tb + ((uint32_t)(32 * quad) << 16) + kColD + half * NCOL;
          tmem_ld_16x256b_x2(la, r0);                    // rows gq (i), gq+8 (f) of this quadrant
          tmem_ld_16x256b_x2(la + (16u << 16), r1);      // rows 16+gq (g), 24+gq (o)
          tmem_ld_wait();
        }
      }
      // pre-activations of my two cells: hi-part columns + lo-part columns
      float ai0 = __uint_as_float(r0[0]) + __uint_as_float(r0[4]), ai1 = __uint_as_float(r0[1]) + __uint_as_float(r0[5]);
      float af0 = __uint_as_float(r0[2]) + __uint_as_float(r0[6]), af1 = __uint_as_float(r0[3]) + __uint_as_float(r0[7]);
      float ag0 = __uint_as_float(r1[0]) + __uint_as_float(r1[4]), ag1 = __uint_as_float(r1[1]) + __uint_as_float(r1[5]);
      float ao0 = __uint_as_float(r1[2]) + __uint_as_float(r1[6]), ao1 = __uint_as_float(r1[3]) + __uint_as_float(r1[7]);
      float h0, h1;
      if (mode & 4) {
        auto sg = [](float x) { float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4427f * x)); float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(1.0f + e)); return y; };
        c0 = sg(af0) * c0 + sg(ai0) * (2.f * sg(2.f * ag0) - 1.f);
        c1 = sg(af1) * c1 + sg(ai1) * (2.f * sg(2.f * ag1) - 1.f);
        h0 = sg(ao0) * (2.f * sg(2.f * c0) - 1.f);
        h1 = sg(ao1) * (2.f * sg(2.f * c1) - 1.f);
      } else {
        h0 = ai0 + af0 + ag0 + ao0 + c0;
        h1 = ai1 + af1 + ag1 + ao1 + c1;
      }
      if (steps > 1) {  // timing runs: publish h (hi rows n0,n1 ; lo rows 8+n0, 8+n1) for the next step
        const __nv_bfloat16 a0 = __float2bfloat16_rn(h0), a1 = __float2bfloat16_rn(h1);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(h0 - __bfloat162float(a0)), l1 = __float2bfloat16_rn(h1 - __bfloat162float(a1));
        unsigned char* Bn = sm.B[nb];
        const int n0 = 2 * tig, n1 = n0 + 1;
        const uint32_t ko = (u & 7) * 2, kc = u >> 3;
        *reinterpret_cast<__nv_bfloat16*>(Bn + sw128_offset(n0, kc) + ko) = a0;
        *reinterpret_cast<__nv_bfloat16*>(Bn + sw128_offset(n1, kc) + ko) = a1;
        *reinterpret_cast<__nv_bfloat16*>(Bn + sw128_offset(8 + n0, kc) + ko) = l0;
        *reinterpret_cast<__nv_bfloat16*>(Bn + sw128_offset(8 + n1, kc) + ko) = l1;
        fence_async_smem();
      }
      fence_before_sync();
      if (mode & 8) mbar_arrive(&sm.hbar);
    }
    if (mode & 8) {
      if (warp == 8 && lane == 0) {
        mbar_wait(&sm.hbar, s & 1);
        fence_after_sync();
        if ((mode & 1) && s + 1 < steps) issue_mmas(nb);
      }
    } else {
      __syncthreads();
      if (tid == 0 && (mode & 1) && s + 1 < steps) {
        fence_after_sync();
        issue_mmas(nb);
      }
    }
  }
  long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  if (worker && out != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      out[(size_t)tid * 16 + j] = __uint_as_float(r0[j]);
      out[(size_t)tid * 16 + 8 + j] = __uint_as_float(r1[j]);
    }
  }
  if (c0 + c1 == 123.456f) cycles[0] = 0;
  fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tb, kTmemCols);
}

static float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

int main() {
  std::vector<float> W(M * H), Hin(H * NCOL), D(M * NCOL);
  float *dW, *dH, *dOut;
  long long* dCyc;
  cudaMalloc(&dW, W.size() * 4);
  cudaMalloc(&dH, Hin.size() * 4);
  cudaMalloc(&dOut, 256 * 16 * 4);
  cudaMalloc(&dCyc, 256 * 8);
  const size_t smem = sizeof(Smem) + 1024;
  cudaFuncSetAttribute(k_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);

  for (int test = 0; test < 2; ++test) {
    for (int r = 0; r < M; ++r)
      for (int k = 0; k < H; ++k) {
        if (test == 0) W[r * H + k] = k == 0 ? (float)(r % 128) : (k == 1 ? 1.f : (k == 2 ? (float)(r / 128) : 0.f));
        else W[r * H + k] = (float)((r * 7 + k * 3) % 9 - 4) + (float)((r + 2 * k) % 5 - 2) / 1024.f;  // hi + lo parts
      }
    for (int k = 0; k < H; ++k)
      for (int n = 0; n < NCOL; ++n) {
        if (test == 0) Hin[k * NCOL + n] = k == 0 ? 16.f : (k == 1 ? (float)n : (k == 2 ? 4096.f : 0.f));
        else Hin[k * NCOL + n] = (float)((k * 5 + n * 11) % 7 - 3);
      }
    for (int r = 0; r < M; ++r)
      for (int n = 0; n < NCOL; ++n) {
        double a = 0;
        for (int k = 0; k < H; ++k) a += (double)W[r * H + k] * bf16r(Hin[k * NCOL + n]);
        D[r * NCOL + n] = (float)a;
      }
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dH, Hin.data(), Hin.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dOut, 0, 256 * 16 * 4);
    k_step<<<1, 288, smem>>>(dW, dH, dOut, dCyc, 1, 1);
    cudaError_t e = cudaDeviceSynchronize();
    printf("test %d: run %s\n", test, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> out(256 * 16);
    cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int tid = 0; tid < 256; ++tid) {
      const int warp = tid / 32, lane = tid % 32, gq = lane / 4, tig = lane % 4, half = warp / 4, quad = warp % 4;
      for (int ld = 0; ld < 2; ++ld)
        for (int j = 0; j < 8; ++j) {
          const int rep = j / 4, jj = j % 4;
          const int row = 128 * half + 32 * quad + 16 * ld + gq + ((jj & 2) ? 8 : 0), col = rep * 8 + 2 * tig + (jj & 1);
          const float got = out[tid * 16 + ld * 8 + j], exp = D[row * NCOL + col];
          if (fabsf(got - exp) > 1e-3f * fmaxf(1.f, fabsf(exp))) {
            if (bad < 12) {
              printf("  mismatch tid %d ld %d reg %d: got %g expected D[%d][%d]=%g", tid, ld, j, got, row, col, exp);
              if (test == 0) {  // decode which element it is: D = 16*(r%128) + n + 4096*(r/128)
                const int v = (int)got;
                printf("  -> looks like row %d col %d", (v % 4096) / 16 + 128 * (v / 4096), v % 16);
              }
              printf("\n");
            }
            ++bad;
          }
        }
    }
    printf("test %d: %d mismatches of 4096\n", test, bad);
  }

  // ---- timing ----
  struct V { int mode; const char* name; };
  const V vs[] = {{0, "skeleton: STS h + fence + bar.sync (no MMA)"},
                  {1, "16 MMA (hi/lo) + commit + mbar wait + 2 tcgen05.ld"},
                  {1 | 2, "8 MMA (hi only) + commit + mbar wait + 2 tcgen05.ld"},
                  {1 | 16, "16 MMA, no tcgen05.ld"},
                  {1 | 4, "16 MMA + fake LSTM activations (20 MUFU/thread)"},
                  {4, "no MMA + fake activations"},
                  {1 | 8, "16 MMA, mbarrier handoff to a dedicated MMA warp"},
                  {1 | 4 | 8, "16 MMA + activations, mbarrier handoff"},
                  {1 | 2 | 4 | 8, "8 MMA + activations, mbarrier handoff"}};
  const int steps = 2000;
  for (int grid : {1, 148}) {
    for (const V& v : vs) {
      k_step<<<grid, 288, smem>>>(dW, dH, nullptr, dCyc, steps, v.mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", v.name, cudaGetErrorString(e)); return 1; }
      std::vector<long long> cyc(grid);
      cudaMemcpy(cyc.data(), dCyc, grid * 8, cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (long long c : cyc) mx = c > mx ? c : mx;
      printf("grid %3d  %-58s %8.1f cycles/step\n", grid, v.name, (double)mx / steps);
    }
  }
  return 0;
}
