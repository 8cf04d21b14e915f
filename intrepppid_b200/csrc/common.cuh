// Shared device helpers for libib200 (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ib200 {

constexpr int kBC = 8;  // sequences per CTA in the recurrent kernels (= N of mma.m16n8k16)

// ---------------------------------------------------------------------------------------------------------------------
// Gate-interleaved ("GI") column order used for every [.,4H] activation tensor (input projections, saved gates, dgates):
//   gi(u, q) = 4*u + q     u = hidden unit, q = gate in PyTorch order (0=i, 1=f, 2=g, 3=o)
// so that the four gates of one cell are one aligned float4.  PyTorch row r = q*H + u.
// ---------------------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int gi_to_torch_row(int gi, int H) { return (gi & 3) * H + (gi >> 2); }

// D(16x8,f32) += A(16x16,bf16,row) * B(16x8,bf16,col)
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_bf16(float lo_elem, float hi_elem) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo_elem, hi_elem);  // .x (low 16 bits) = lo_elem
  return *reinterpret_cast<uint32_t*>(&v);
}

// (x,y) -> packed bf16 pair `hi` = round(x,y) and `lo` = round((x,y) - hi): x ~= hi + lo to ~2^-17 relative.
__device__ __forceinline__ void split_bf16(float x, float y, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  float2 hf = __bfloat1622float2(h);
  __nv_bfloat162 l = __floats2bfloat162_rn(x - hf.x, y - hf.y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ---------------------------------------------------------------------------------------------------------------------
// activations.  ACCURATE: ex2.approx + rcp.approx (abs error ~3e-7) for the fp32 mode; FAST: one MUFU.TANH each.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool FAST>
__device__ __forceinline__ float sigmoid_f(float x) {
  if constexpr (FAST) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return fmaf(0.5f, t, 0.5f);
  } else {
    // 1/(1+2^(-x log2 e)): FMUL, MUFU.EX2, FADD, MUFU.RCP.  ex2 -> +inf gives 0, -> 0 gives 1: no special cases needed.
    return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x));
  }
}
template <bool FAST>
__device__ __forceinline__ float tanh_f(float x) {
  if constexpr (FAST) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
    return t;
  } else {
    // 1 - 2/(1+e^{2x}); saturates cleanly (e^{2x} -> inf gives 1, -> 0 gives -1)
    return fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(2.8853900817779268f * x)), 1.0f);
  }
}

// four 8x8 b16 matrices from smem, transposed on load: with rows = k (8 b16 = one 16-byte row of 8 sequences) this yields
// the mma B fragments {b0,b1} of two consecutive k16 tiles directly.
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row_ptr) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row_ptr));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a)
               : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src, bool valid) {
  uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gmem_src), "r"(sz) : "memory");
}

__device__ __forceinline__ float mish_f(float x) {
  // x * tanh(softplus(x)); softplus with PyTorch's threshold=20 (nn.Mish -> F.mish)
  float sp = x > 20.0f ? x : log1pf(expf(x));
  return x * tanhf(sp);
}
__device__ __forceinline__ float mish_grad_f(float x) {
  float sp = x > 20.0f ? x : log1pf(expf(x));
  float tsp = tanhf(sp);
  float sig = 1.0f / (1.0f + expf(-x));
  return tsp + x * (1.0f - tsp * tsp) * sig;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
  uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  int sz = valid ? 16 : 0;  // src-size 0 => zero-fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// Two-phase rebalancing of a recurrent launch whose CTAs do not divide over the SMs (e.g. 200 CTAs, two resident per SM at most,
// on 148 SMs: 52 SMs hold two CTAs and set the kernel time while 96 hold one and idle 40 % of it).
//   phase 1: the whole grid starts.  Every CTA registers on its SM (atomic counter per %smid); a CTA that finds a second resident
//            on its SM stops after `split` steps, saves its recurrent state and appends its id to `resume_list`.
//   phase 2: a 1-D grid (at most one CTA per SM) resumes the listed CTAs from the saved state to the end of the chain.
// The split point is chosen so that the shared SMs and the exclusive SMs finish phase 1 together.
// ---------------------------------------------------------------------------------------------------------------------
struct PhaseArgs {
  float* state;       // [ndir][N][H][2] recurrent state carried from phase 1 to phase 2 (forward: c, h; backward: dc, dh)
  int* sm_load;       // [256] resident-CTA counters per SM id, zeroed before phase 1
  int* resume_list;   // [0] = number of stopped CTAs (zeroed before phase 1), [1..] = their linear ids (x + grid_x*(g + G*dz))
  int phase;          // 0: one launch runs the whole chain (default)
  float split_frac;   // fraction of the chain a shared CTA runs in phase 1
  int grid_x;         // blockIdx.x extent of the phase-1 grid (tiles per group)
};
constexpr int kPhaseCheck = 16;  // step at which a phase-1 CTA looks at its SM's counter (all CTAs of the launch have started by then)

__device__ __forceinline__ int sm_id() {
  unsigned id;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
  return (int)(id & 255u);  // index into PhaseArgs::sm_load[256]
}
// which (tile, group, direction slot) this CTA works on; false: nothing to do (phase 2 CTA beyond the list)
__device__ __forceinline__ bool phase_cta(const PhaseArgs& ph, int G, int& bx, int& g, int& dz) {
  bx = blockIdx.x; g = blockIdx.y; dz = blockIdx.z;
  if (ph.phase != 2) return true;
  if ((int)blockIdx.x >= ph.resume_list[0]) return false;
  const int id = ph.resume_list[1 + blockIdx.x];
  bx = id % ph.grid_x; g = (id / ph.grid_x) % G; dz = id / (ph.grid_x * G);
  return true;
}
__device__ __forceinline__ int phase_split(const PhaseArgs& ph, int T) {
  const int s = max((int)(ph.split_frac * (float)T), kPhaseCheck + 2);
  return s < T ? s : T;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace ib200
