// One-shot mean all-reduce of a small gradient bucket over NVLink peer memory (the only exchange step of the path).
//
// The data-parallel gradient exchange is latency bound (753 KB per step at the headline shape, in buckets of 6 .. 600 KB): a ring /
// tree collective pays several launches and hops.  Here every rank owns a staging region that its peers have mapped (CUDA IPC), and
// ONE kernel per rank does the whole reduction:
//   kernel A: copy my bucket into my staging region (parity half `epoch & 1`)
//   kernel B: ONE warp publishes "rank r has staged epoch e" into the flag array of EVERY rank (system-scope release stores over
//             NVLink) and polls (with nanosleep back-off) until all ranks have published epoch e in MY flag array -- a single
//             sleeping warp, because the early buckets are reduced while the layer-0 BPTT kernel owns the SMs and a grid of spinning
//             blocks lengthens its dependent chains (measured on 8 GPUs: 0.28 ms exposed instead of 0.14)
//   kernel C: sums the world staged copies -- 16-byte loads straight from the peers' memory -- and writes the mean back into my bucket.
// Two parity halves make the buffers safe without a trailing barrier: a rank overwrites half p again only after it has passed the wait
// of the NEXT epoch, which every peer publishes only after its reads of half p have completed (stream order on that peer).
// A lost peer makes the poll give up with a trap after ~15 s instead of hanging the GPU.
#include <algorithm>
#include <cstdint>

#include "kernels.h"

namespace ib200 {
namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer data changes every step: never through a non-coherent path
__device__ __forceinline__ float4 ld_peer4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer1(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];\n" : "=f"(v) : "l"(p) : "memory");
  return v;
}

__global__ void p2p_stage_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

__global__ void __launch_bounds__(32) p2p_sync_kernel(const P2PArgs a) {
  if (threadIdx.x < a.world) {
    // publish: my bucket of this epoch is staged (kernel A has completed: stream order)
    __threadfence_system();
    st_release_sys(a.flags[threadIdx.x] + a.rank, a.epoch);
    // wait until every rank has published this epoch in MY flag array
    const uint32_t* f = a.flags[a.rank] + threadIdx.x;
    uint32_t polls = 0;
    while ((int32_t)(ld_acquire_sys(f) - a.epoch) < 0) {
      __nanosleep(200);
      if (++polls > (1u << 26)) __trap();  // ~15 s: a lost peer fails loudly
    }
  }
}

__global__ void __launch_bounds__(256) p2p_reduce_kernel(const P2PArgs a) {
  // sum the staged copies (fixed rank order: the result is bit-identical on every rank) and write the mean
  const float inv = 1.0f / (float)a.world;
  const size_t stride = (size_t)gridDim.x * blockDim.x, n4 = a.n / 4;
  const bool aligned = (reinterpret_cast<uintptr_t>(a.data) & 15) == 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < a.world; ++r) {
      const float4 v = ld_peer4(a.stage[r] + 4 * i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    s.x *= inv; s.y *= inv; s.z *= inv; s.w *= inv;
    if (aligned) {
      reinterpret_cast<float4*>(a.data)[i] = s;
    } else {
      a.data[4 * i] = s.x; a.data[4 * i + 1] = s.y; a.data[4 * i + 2] = s.z; a.data[4 * i + 3] = s.w;
    }
  }
  for (size_t i = 4 * n4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    float s = 0.f;
    for (int r = 0; r < a.world; ++r) s += ld_peer1(a.stage[r] + i);
    a.data[i] = s * inv;
  }
}

}  // namespace

cudaError_t launch_p2p_allreduce_mean(const P2PArgs& a, float* my_stage, cudaStream_t st) {
  if (a.world < 1 || a.world > kP2PMaxWorld || a.rank < 0 || a.rank >= a.world) return cudaErrorInvalidValue;
  if (a.n == 0) return cudaSuccess;
  const unsigned blocks = (unsigned)std::min<size_t>(64, (a.n / 4 + 255) / 256 + 1);
  p2p_stage_kernel<<<blocks, 256, 0, st>>>(a.data, my_stage, a.n);
  p2p_sync_kernel<<<1, 32, 0, st>>>(a);
  p2p_reduce_kernel<<<blocks, 256, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace ib200
