// Microbenchmarks behind the tcgen05 cluster recurrent kernels (lstm_cluster_tc.cu), B200 / sm_100a:
//   (1) cost of a batch of small-N tcgen05.mma (M = 128, N = 32 / 64, K = 16 each) issued back to back on ONE accumulator, A from
//       shared memory (SS) or from TMEM (TS), + commit + mbarrier wait (+ tcgen05.ld): the per-step tensor time of a CTA that owns
//       128 gate rows of W_hh (K = 256) for 32 / 64 sequences;
//   (2) cost of the per-step h all-gather inside a cluster of 8 CTAs: every CTA sends `bytes_per_dst` to each of the 8 CTAs and waits
//       for its own 8 slices -- with 16-byte st.async stores, or with one cp.async.bulk shared::cta -> shared::cluster per destination.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/microbench_cl tools/microbench_cl.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../intrepppid_b200/csrc/tc05.cuh"

using namespace ib200::tc;

__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}

// ---- (1) -------------------------------------------------------------------------------------------------------------------------
// smem: A_hi | A_lo: [4 k-blocks][128 rows x 128 B]; B_hi | B_lo: [4 k-blocks][N rows x 128 B]
template <int N, int NPARTS, bool TS, int NACC = 1>
__global__ void __launch_bounds__(160, 1) k_mma(long long* cyc, int steps, int do_ld) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int kABlk = 128 * 128, kBBlk = N * 128, kblocks = 4;
  unsigned char* A = smem;                       // 2 parts x 4 blocks
  unsigned char* B = smem + 2 * 4 * kABlk;       // 2 parts x 4 blocks
  uint64_t* bar = reinterpret_cast<uint64_t*>(B + 2 * 4 * kBBlk);
  uint32_t* tbase = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (2 * 4 * kABlk + 2 * 4 * kBBlk) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init_fence();
  }
  if (warp == 4) tmem_alloc(tbase, 512);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = *tbase;
  constexpr uint32_t idesc = idesc_bf16(128, N, false, false);
  // loop-invariant base descriptors: every MMA's descriptor is base + a compile-time constant (start address field, 16-byte units)
  const uint64_t ah0 = smem_desc_sw128(smem_u32(A), 1024, 0), al0 = smem_desc_sw128(smem_u32(A) + 4 * kABlk, 1024, 0);
  const uint64_t bh0 = smem_desc_sw128(smem_u32(B), 1024, 0), bl0 = smem_desc_sw128(smem_u32(B) + 4 * kBBlk, 1024, 0);
  long long t0 = clock64();
  for (int s = 0; s < steps; ++s) {
    if (tid == 128) {
      fence_after_sync();
#pragma unroll
      for (int kb = 0; kb < kblocks; ++kb)
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) {
          const uint64_t bo = (uint64_t)((kb * kBBlk + k16 * 32) >> 4), ao = (uint64_t)((kb * kABlk + k16 * 32) >> 4);
          const bool first = (kb * 4 + k16) < NACC;
          const uint32_t dcol = 256 + ((kb * 4 + k16) % NACC) * N;  // NACC accumulators used round-robin
          if constexpr (TS) {
            const uint32_t ah = tb + (kb * 4 + k16) * 8, al = tb + 128 + (kb * 4 + k16) * 8;
            mma_bf16_ts(tb + dcol, ah, bh0 + bo, idesc, !first);
            if constexpr (NPARTS == 3) {
              mma_bf16_ts(tb + dcol, ah, bl0 + bo, idesc, true);
              mma_bf16_ts(tb + dcol, al, bh0 + bo, idesc, true);
            }
          } else {
            mma_bf16_ss(tb + dcol, ah0 + ao, bh0 + bo, idesc, !first);
            if constexpr (NPARTS == 3) {
              mma_bf16_ss(tb + dcol, ah0 + ao, bl0 + bo, idesc, true);
              mma_bf16_ss(tb + dcol, al0 + ao, bh0 + bo, idesc, true);
            }
          }
        }
      mma_commit(bar);
    }
    if (warp < 4) {
      mbar_wait(bar, s & 1);
      fence_after_sync();
      if (do_ld) {
        uint32_t r[32];
        tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + 256, r);
        if (r[0] == 0x12345678u && r[31] == 0x9abcdef0u) cyc[1] = 1;
      }
      fence_before_sync();
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (tid == 0) cyc[0] = t1 - t0;
  fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tb, 512);
}

template <int N, int NPARTS, bool TS, int NACC = 1>
void run_mma(long long* dCyc, int steps) {
  const size_t smem = 1024 + 2 * 4 * 128 * 128 + 2 * 4 * N * 128 + 64;
  cudaFuncSetAttribute(k_mma<N, NPARTS, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int do_ld = 0; do_ld < 2; ++do_ld) {
    cudaMemset(dCyc, 0, 16);
    k_mma<N, NPARTS, TS, NACC><<<1, 160, smem>>>(dCyc, steps, do_ld);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("k_mma: %s\n", cudaGetErrorString(e)); exit(1); }
    long long c;
    cudaMemcpy(&c, dCyc, 8, cudaMemcpyDeviceToHost);
    const int nm = 16 * NPARTS;
    printf("mma M=128 N=%3d %s %d acc  %2d MMAs (K=256, %d part%s) %s : %8.1f cycles/step  (%.1f per MMA)\n", N, TS ? "TS" : "SS", NACC, nm, NPARTS,
           NPARTS > 1 ? "s" : " ", do_ld ? "+ld" : "   ", (double)c / steps, (double)c / steps / nm);
  }
}

// ---- (2) -------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t map_to_rank(uint32_t a, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];\n" ::"r"(addr), "r"(a),
               "r"(b), "r"(c), "r"(d), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void bulk_s2c(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(bar_cluster)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// every CTA: slice (bytes_per_dst) -> buffer [2][C][bytes_per_dst] of each CTA, slot = my rank; wait for my own C slices
// mode 0: st.async 16 B from all threads; mode 1: local staging + fence.proxy.async + bar + one bulk copy per destination
__global__ void __launch_bounds__(256, 1) k_xchg(long long* cyc, int steps, int bytes_per_dst, int mode, int C) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* rbuf = smem;                                    // [2][C][bytes_per_dst]
  unsigned char* stage = smem + 2 * C * bytes_per_dst;           // [2][bytes_per_dst]
  uint64_t* bar = reinterpret_cast<uint64_t*>(stage + 2 * bytes_per_dst);  // [2]
  const int tid = threadIdx.x;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
  const uint32_t step_bytes = (uint32_t)(C * bytes_per_dst);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_init_fence();
    mbar_arrive_expect_tx(&bar[1], step_bytes);
  }
  __syncthreads();
  cluster_sync_all();
  const uint32_t rb = smem_u32(rbuf), bb = smem_u32(bar);
  long long t0 = clock64();
  for (int s = 0; s < steps; ++s) {
    const int buf = s & 1;
    if (s > 0) mbar_wait(&bar[buf], (uint32_t)(((s - 1) >> 1) & 1));
    if (tid == 0 && s + 2 < steps) mbar_arrive_expect_tx(&bar[buf], step_bytes);
    if (s + 1 < steps) {
      const uint32_t dst_off = (uint32_t)((buf ^ 1) * C + rank) * bytes_per_dst;
      if (mode == 0) {
        const int pieces = bytes_per_dst / 16;
        for (int i = tid; i < pieces * C; i += blockDim.x) {
          const int r = i / pieces, pc = i % pieces;
          st_async_v4(map_to_rank(rb, r) + dst_off + pc * 16, s, i, 2, 3, map_to_rank(bb, r) + (buf ^ 1) * 8);
        }
      } else {
        unsigned char* stg = stage + (buf ^ 1) * bytes_per_dst;
        for (int i = tid; i < bytes_per_dst / 16; i += blockDim.x) reinterpret_cast<uint4*>(stg)[i] = make_uint4(s, i, 2, 3);
        fence_async_smem();
        __syncthreads();
        if (tid < C) bulk_s2c(map_to_rank(rb, tid) + dst_off, smem_u32(stg), bytes_per_dst, map_to_rank(bb, tid) + (buf ^ 1) * 8);
      }
    }
  }
  long long t1 = clock64();
  cluster_sync_all();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  long long* dCyc;
  cudaMalloc(&dCyc, 1024 * 8);
  const int steps = 2000;
  run_mma<32, 1, false, 1>(dCyc, steps); run_mma<32, 3, false, 1>(dCyc, steps);
  run_mma<32, 1, false, 2>(dCyc, steps); run_mma<32, 3, false, 2>(dCyc, steps);
  run_mma<32, 1, false, 4>(dCyc, steps); run_mma<32, 3, false, 4>(dCyc, steps);
  run_mma<32, 3, true, 2>(dCyc, steps);
  run_mma<40, 1, false, 1>(dCyc, steps); run_mma<40, 3, false, 1>(dCyc, steps);
  run_mma<48, 1, false, 1>(dCyc, steps); run_mma<48, 3, false, 1>(dCyc, steps);
  run_mma<64, 3, false, 1>(dCyc, steps);
  if (getenv("MB_XCHG") == nullptr) return 0;
  for (int nclusters : {1, 16})
  for (int C : {4, 8})
    for (int bytes : {1024, 2048, 4096})
      for (int mode = 0; mode < 2; ++mode) {
        const size_t smem = 1024 + 2 * C * bytes + 2 * bytes + 64;
        cudaFuncSetAttribute(k_xchg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(C * nclusters);
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = C;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_xchg, dCyc, steps, bytes, mode, C);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("k_xchg: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> c(C * nclusters);
        cudaMemcpy(c.data(), dCyc, c.size() * 8, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (long long v : c) mx = v > mx ? v : mx;
        printf("xchg %2d clusters C=%d  %4d B per destination (%5d B out per CTA-step) %s : %8.1f cycles/step\n", nclusters, C, bytes, C * bytes,
               mode ? "bulk s2c (stage+fence+bar+8 copies)" : "st.async 16 B                      ", (double)mx / steps);
      }
  return 0;
}
