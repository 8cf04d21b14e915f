// K0 / K1a / K4: truncation lengths, layer-0 input-projection table, bi_reduce pooling + fc.
#include "kernels.h"
#include "small.h"

namespace ib200 {
namespace {

// ---- A1: T1[g] = max_b #{t : tok != 0} (a COUNT, encoders/awd_lstm.py:149-150) + int32 copy of the ids ---------------------
// TOK = the caller's id type: int64 as the reference's dataloader ships them (data/ppi_oma.py:388-390), or the narrowed
// int32 / int16 / uint8 ids of the input-feeding path (V = 250 fits a byte -> 8x less H2D traffic).
template <typename TOK>
__global__ void len1_kernel(const LengthArgs p) {
  const int n = blockIdx.x, g = n / p.B;
  const TOK* __restrict__ src = (const TOK*)p.tokens + (size_t)n * p.Tin;
  int* __restrict__ dst = p.tok32 + (size_t)n * p.Tin;
  int cnt = 0, bad = 0;
  for (int t = threadIdx.x; t < p.Tin; t += blockDim.x) {
    const long long v64 = (long long)src[t];
    // ids outside [0, V) make F.embedding raise in the reference; here they are clamped so that no kernel can index outside the
    // [V, .] tables, and the group's sticky status flag (lens row 2, bit 0) records it for the caller's lazy check (no host sync)
    const int v = v64 < 0 ? 0 : (v64 >= p.V ? p.V - 1 : (int)v64);
    bad |= (v64 < 0 || v64 >= p.V);
    dst[t] = v;
    cnt += (v64 != 0);
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(p.lens + 2 * p.G + g, kStatusBadToken);
  __shared__ int ws[32];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += ws[i];
    atomicMax(p.lens + g, tot);
  }
}

// ---- A3: T_eff[g] = max_{b,e} #{t < T1 : (scale[g][tok] * emb[tok][e]) != 0} (encoders/awd_lstm.py:53-54, quirk Q2) -----------
// count[b][e] = sum_v hist_b[v] * [scale[g][v] != 0 && emb[v][e] != 0]   (scale >= 1 when non-zero, so no underflow to 0).
// A vocabulary row is "full" when it is kept and has no zero entry: it adds hist_b[v] to EVERY column.  Rows that are kept but
// contain zeros (normally only the zero-initialised padding row) are listed per group and handled column by column.
__global__ void nz_rows_kernel(const LengthArgs p, int* __restrict__ row_kind) {  // row_kind[g][v]: 0 contributes nothing, 1 full, 2 partial
  const int g = blockIdx.y, v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (v >= p.V) return;
  int nz = 0;
  for (int e = lane; e < p.H; e += 32) nz += p.emb[(size_t)v * p.H + e] != 0.0f;
  nz = __reduce_add_sync(0xffffffffu, nz);
  const bool keep = p.emb_row_scale == nullptr || p.emb_row_scale[(size_t)g * p.V + v] != 0.0f;
  if (lane == 0) row_kind[(size_t)g * p.V + v] = (!keep || nz == 0) ? 0 : (nz == p.H ? 1 : 2);
}

__global__ void len2_kernel(const LengthArgs p, const int* __restrict__ row_kind) {
  extern __shared__ int hist[];  // [V] histogram, then [V] list of partial rows
  int* partial = hist + p.V;
  __shared__ int n_partial, full_sum;
  const int n = blockIdx.x, g = n / p.B;
  const int T1 = p.lens[g];
  for (int v = threadIdx.x; v < p.V; v += blockDim.x) hist[v] = 0;
  if (threadIdx.x == 0) n_partial = full_sum = 0;
  __syncthreads();
  const int* __restrict__ tk = p.tok32 + (size_t)n * p.Tin;
  for (int t = threadIdx.x; t < T1; t += blockDim.x) atomicAdd(&hist[tk[t]], 1);  // ids were clamped to [0, V) by len1_kernel
  __syncthreads();
  int mine = 0;
  for (int v = threadIdx.x; v < p.V; v += blockDim.x) {
    const int kind = row_kind[(size_t)g * p.V + v], hv = hist[v];
    if (kind == 1) mine += hv;
    else if (kind == 2 && hv != 0) partial[atomicAdd(&n_partial, 1)] = v;
  }
  mine = __reduce_add_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0 && mine != 0) atomicAdd(&full_sum, mine);
  __syncthreads();
  int best = 0;
  if (n_partial == 0) {
    best = full_sum;
  } else {
    for (int e = threadIdx.x; e < p.H; e += blockDim.x) {
      int cnt = full_sum;
      for (int i = 0; i < n_partial; ++i) {
        const int v = partial[i];
        cnt += (p.emb[(size_t)v * p.H + e] != 0.0f) ? hist[v] : 0;
      }
      best = max(best, cnt);
    }
    best = __reduce_max_sync(0xffffffffu, best);
  }
  if ((threadIdx.x & 31) == 0 && best > 0) atomicMax(p.lens + p.G + g, best);
}

__global__ void zero_int_kernel(int* p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0;
}

// ---- K1a: P[g][d][v][gi] = (scale[g][v] * emb[v]) . w_ih_d[row(gi)] + b_ih_d[row] + b_hh_d[row] -----------------------------------
// (utils/embedding_do.py:26-43 folded with the layer-0 x W_ih^T of nn.LSTM: the lookup-table identity)
constexpr int kTableVpb = 32;  // vocabulary rows per block

// One block = 32 vocabulary rows of one (group, direction).  W_ih is staged once per block, transposed, in shared memory (coalesced
// global reads; the +1 pitch keeps the transposing stores conflict-free), every thread owns one gate column gi and keeps 32 running
// sums; the embedding rows are read as broadcast float4.  The sum over k runs in ascending order with fmaf, one accumulator per
// output, exactly like a plain dot product loop.
template <int H>
__global__ void __launch_bounds__(4 * H) l0_table_kernel(const TableArgs p) {
  constexpr int NG = 4 * H, WP = NG + 1;
  extern __shared__ __align__(16) float tsm[];
  float* Wt = tsm;                 // [H][WP]: Wt[k][gi] = w_ih[row(gi)][k]
  float* xs = tsm + H * WP;        // [kTableVpb][H] (16-byte aligned: H is a multiple of 4)
  static_assert((H * WP) % 4 == 0, "xs alignment");
  const int gi = threadIdx.x, row = gi_to_torch_row(gi, H);
  const int g = blockIdx.y >> 1, d = blockIdx.y & 1;
  const int VT = p.V + kPadRows;  // rows >= V replicate row 0 (pad replicas, kernels.h)
  const int v0 = blockIdx.x * kTableVpb, nv = min(VT - v0, kTableVpb);
  const float* __restrict__ W = p.w_ih[d];
  if ((reinterpret_cast<uintptr_t>(W) & 15) == 0) {
    // the whole 4H x H matrix in flight at once: H/4 independent 16-byte loads per thread, then the transposing stores
    const float4* __restrict__ W4 = reinterpret_cast<const float4*>(W);
    float4 w4[H / 4];
#pragma unroll
    for (int it = 0; it < H / 4; ++it) w4[it] = W4[gi + it * NG];
#pragma unroll
    for (int it = 0; it < H / 4; ++it) {
      const int i4 = gi + it * NG, r = i4 / (H / 4), k = (i4 % (H / 4)) * 4;  // PyTorch row r = q*H + u -> gate-interleaved column 4u + q
      float* dst = Wt + k * WP + 4 * (r % H) + r / H;
      dst[0] = w4[it].x; dst[WP] = w4[it].y; dst[2 * WP] = w4[it].z; dst[3 * WP] = w4[it].w;
    }
  } else {
#pragma unroll 16
    for (int it = 0; it < H; ++it) {
      const int i = gi + it * NG, r = i / H, k = i % H;
      Wt[k * WP + 4 * (r % H) + r / H] = W[i];
    }
  }
  // the reference multiplies mask/(1-p) into the row first (embedding_do.py:26-29)
#pragma unroll
  for (int it = 0; it < kTableVpb / 4; ++it) {
    const int i = gi + it * NG;
    const int j = i / H, k = i % H, vt = v0 + j, v = vt < p.V ? vt : 0;
    float x = 0.f;
    if (j < nv) x = (p.emb_row_scale != nullptr ? p.emb_row_scale[(size_t)g * p.V + v] : 1.0f) * p.emb[(size_t)v * H + k];
    xs[i] = x;
  }
  __syncthreads();
  float acc[kTableVpb];
#pragma unroll
  for (int j = 0; j < kTableVpb; ++j) acc[j] = 0.f;
  const float4* xs4 = reinterpret_cast<const float4*>(xs);
#pragma unroll 2
  for (int k4 = 0; k4 < H / 4; ++k4) {
    const float w0 = Wt[(4 * k4 + 0) * WP + gi], w1 = Wt[(4 * k4 + 1) * WP + gi], w2 = Wt[(4 * k4 + 2) * WP + gi],
                w3 = Wt[(4 * k4 + 3) * WP + gi];
#pragma unroll
    for (int j = 0; j < kTableVpb; ++j) {
      const float4 x = xs4[j * (H / 4) + k4];
      acc[j] = fmaf(x.w, w3, fmaf(x.z, w2, fmaf(x.y, w1, fmaf(x.x, w0, acc[j]))));
    }
  }
  const float bias = p.b_ih[d][row] + p.b_hh[d][row];
#pragma unroll
  for (int j = 0; j < kTableVpb; ++j)  // (compile-time indices keep acc[] in registers)
    if (j < nv) p.table[(((size_t)(g * 2 + d)) * VT + v0 + j) * 4 * H + gi] = acc[j] + bias;
}

// any H (used for H > 64, where the weight row no longer fits the register file): same summation order, weights from L2.  A block
// covers kTableVpbGeneric vocabulary rows at once, so every weight is read once per block instead of once per vocabulary row (the
// per-row version took 1.1 ms per call at H = 256: a thread walks its own weight row, 32 sectors per warp request)
constexpr int kTableVpbGeneric = 8;
__global__ void __launch_bounds__(256) l0_table_generic_kernel(const TableArgs p) {
  extern __shared__ float xs[];  // [kTableVpbGeneric][H]
  const int H = p.H, g = blockIdx.y >> 1, d = blockIdx.y & 1;
  const int VT = p.V + kPadRows;
  const int v0 = blockIdx.x * kTableVpbGeneric;
  for (int i = threadIdx.x; i < kTableVpbGeneric * H; i += blockDim.x) {
    const int j = i / H, e = i % H, vt = v0 + j;
    const int v = vt < p.V ? vt : 0;  // rows >= V: replicas of row 0 (pad rows, kernels.h); beyond V + kPadRows: unused
    const float sc = p.emb_row_scale != nullptr ? p.emb_row_scale[(size_t)g * p.V + v] : 1.0f;
    xs[i] = vt < VT ? sc * p.emb[(size_t)v * H + e] : 0.f;
  }
  __syncthreads();
  for (int gi = threadIdx.x; gi < 4 * H; gi += blockDim.x) {
    const int row = gi_to_torch_row(gi, H);
    const float* __restrict__ wr = p.w_ih[d] + (size_t)row * H;
    float s[kTableVpbGeneric];
#pragma unroll
    for (int j = 0; j < kTableVpbGeneric; ++j) s[j] = 0.f;
#pragma unroll 4
    for (int k = 0; k < H; ++k) {
      const float w = wr[k];
#pragma unroll
      for (int j = 0; j < kTableVpbGeneric; ++j) s[j] = fmaf(xs[j * H + k], w, s[j]);
    }
    const float bias = p.b_ih[d][row] + p.b_hh[d][row];
#pragma unroll
    for (int j = 0; j < kTableVpbGeneric; ++j)
      if (v0 + j < VT) p.table[(((size_t)(g * 2 + d)) * VT + v0 + j) * 4 * H + gi] = s[j] + bias;
  }
}

// ---- K4: bi_reduce on h_n[-2:] + fc (encoders/awd_lstm.py:58-71) -----------------------------------------------------------------
__global__ void pool_fc_fwd_kernel(int N, int H, int mode, const float* __restrict__ hn, const float* __restrict__ fc_w,
                                   const float* __restrict__ fc_b, float* __restrict__ z, float* __restrict__ pooled_out,
                                   uint8_t* __restrict__ argmax_out) {
  extern __shared__ float pooled[];
  const int n = blockIdx.x;
  for (int e = threadIdx.x; e < H; e += blockDim.x) {
    const float f = hn[(size_t)n * H + e], r = hn[((size_t)N + n) * H + e];
    float v;
    if (mode == 0) v = r;                 // "last": h_n[-1] = top-layer REVERSE final state (quirk Q4)
    else if (mode == 1) v = (f + r) * 0.5f;  // "mean"
    else {                                // "max" (ties -> first = forward, as torch.max(dim=0))
      v = f >= r ? f : r;
      if (argmax_out != nullptr) argmax_out[(size_t)n * H + e] = f >= r ? 0 : 1;
    }
    pooled[e] = v;
    if (pooled_out != nullptr) pooled_out[(size_t)n * H + e] = v;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < H; e += blockDim.x) {
    float s = fc_b[e];
    for (int k = 0; k < H; ++k) s = fmaf(pooled[k], fc_w[(size_t)e * H + k], s);
    z[(size_t)n * H + e] = s;
  }
}

__global__ void pool_fc_bwd_dh_kernel(int N, int H, int mode, const float* __restrict__ dz, const uint8_t* __restrict__ argmax,
                                      const float* __restrict__ fc_w, float* __restrict__ d_hn) {
  extern __shared__ float dzs[];
  const int n = blockIdx.x;
  for (int e = threadIdx.x; e < H; e += blockDim.x) dzs[e] = dz[(size_t)n * H + e];
  __syncthreads();
  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    float s = 0.f;
    for (int e = 0; e < H; ++e) s = fmaf(dzs[e], fc_w[(size_t)e * H + k], s);
    float df, dr;
    if (mode == 0) { df = 0.f; dr = s; }
    else if (mode == 1) { df = 0.5f * s; dr = 0.5f * s; }
    else { const bool rev = argmax[(size_t)n * H + k] != 0; df = rev ? 0.f : s; dr = rev ? s : 0.f; }
    d_hn[(size_t)n * H + k] = df;
    d_hn[((size_t)N + n) * H + k] = dr;
  }
}

// d_fc_w[e][k] = sum_n dz[n][e] pooled[n][k], d_fc_b[e] = sum_n dz[n][e].  One block per output row e; the N samples are
// split over the warps and combined in a fixed order (deterministic).
__global__ void pool_fc_bwd_dw_kernel(int N, int H, const float* __restrict__ dz, const float* __restrict__ pooled,
                                      float* __restrict__ d_fc_w, float* __restrict__ d_fc_b) {
  extern __shared__ float part[];  // [nwarp][H+1]
  const int e = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int k0 = 0; k0 < H; k0 += 32) {
    const int k = k0 + lane;
    float s = 0.f, sb = 0.f;
#pragma unroll 8  // (loads of 8 rows in flight; the fmaf chain keeps its order)
    for (int n = warp; n < N; n += nwarp) {
      const float d = dz[(size_t)n * H + e];
      if (k < H) s = fmaf(d, pooled[(size_t)n * H + k], s);
      sb += d;
    }
    if (k < H) part[warp * (H + 1) + k] = s;
    if (k0 == 0 && lane == 0) part[warp * (H + 1) + H] = sb;
  }
  __syncthreads();
  for (int k = threadIdx.x; k <= H; k += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < nwarp; ++w) s += part[w * (H + 1) + k];
    if (k < H) d_fc_w[(size_t)e * H + k] = s;
    else d_fc_b[e] = s;
  }
}


// ---- per-sequence (T1, T_eff) of a batch-of-one encoder call in eval mode (cli/infer.py:196-225 encodes every protein alone, so
// each one is truncated to ITS OWN lengths: awd_lstm.py:149-150 and :53-54 on a batch of one).  Bookkeeping for the embedding
// cache of intrepppid_b200.infer (bucketing by length); the encoder kernels recompute both lengths per group (K0 above).
__global__ void seq_row_kind_kernel(int V, int H, const float* __restrict__ emb, int* __restrict__ row_kind) {
  const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (v >= V) return;
  int nz = 0;
  for (int e = lane; e < H; e += 32) nz += emb[(size_t)v * H + e] != 0.0f;
  nz = __reduce_add_sync(0xffffffffu, nz);
  if (lane == 0) row_kind[v] = nz == 0 ? 0 : (nz == H ? 1 : 2);  // contributes nothing / to every column / to some columns
}

template <typename TOK>
__global__ void __launch_bounds__(256) seq_lengths_kernel(int T, int V, int H, const TOK* __restrict__ tokens,
                                                          const float* __restrict__ emb, const int* __restrict__ row_kind,
                                                          int* __restrict__ t1_out, int* __restrict__ teff_out) {
  extern __shared__ int hist[];  // [V] histogram, then [V] list of partial rows
  int* partial = hist + V;
  __shared__ int n_partial, full_sum, t1_s, best_s;
  const TOK* __restrict__ src = tokens + (size_t)blockIdx.x * T;
  for (int v = threadIdx.x; v < V; v += blockDim.x) hist[v] = 0;
  if (threadIdx.x == 0) n_partial = full_sum = t1_s = best_s = 0;
  __syncthreads();
  int cnt = 0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) cnt += ((long long)src[t] != 0);
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0 && cnt != 0) atomicAdd(&t1_s, cnt);
  __syncthreads();
  const int T1 = t1_s;  // a COUNT of non-zero ids; the reference then keeps the FIRST T1 positions
  for (int t = threadIdx.x; t < T1; t += blockDim.x) {
    const long long v64 = (long long)src[t];
    atomicAdd(&hist[v64 < 0 ? 0 : (v64 >= V ? V - 1 : (int)v64)], 1);  // clamped like K0 (the encoder call reports bad ids)
  }
  __syncthreads();
  int mine = 0;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const int kind = row_kind[v], hv = hist[v];
    if (kind == 1) mine += hv;
    else if (kind == 2 && hv != 0) partial[atomicAdd(&n_partial, 1)] = v;
  }
  mine = __reduce_add_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0 && mine != 0) atomicAdd(&full_sum, mine);
  __syncthreads();
  int best = 0;
  if (n_partial == 0) {
    best = full_sum;
  } else {
    for (int e = threadIdx.x; e < H; e += blockDim.x) {
      int c = full_sum;
      for (int i = 0; i < n_partial; ++i) c += (emb[(size_t)partial[i] * H + e] != 0.0f) ? hist[partial[i]] : 0;
      best = max(best, c);
    }
    best = __reduce_max_sync(0xffffffffu, best);
  }
  if ((threadIdx.x & 31) == 0 && best > 0) atomicMax(&best_s, best);
  __syncthreads();
  if (threadIdx.x == 0) {
    t1_out[blockIdx.x] = T1;
    teff_out[blockIdx.x] = best_s;
  }
}

template <typename TOK>
cudaError_t launch_seq_lengths_t(int M, int T, int V, int H, const void* tokens, const float* emb, int* row_kind, int* t1_out,
                                 int* teff_out, cudaStream_t st) {
  const size_t hist_bytes = 2 * (size_t)V * sizeof(int);
  if (hist_bytes > kLen2MaxSmem) return cudaErrorInvalidConfiguration;
  if (hist_bytes > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(seq_lengths_kernel<TOK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_bytes);
    if (e != cudaSuccess) return e;
  }
  seq_row_kind_kernel<<<(V + 7) / 8, 256, 0, st>>>(V, H, emb, row_kind);
  seq_lengths_kernel<TOK><<<M, 256, hist_bytes, st>>>(T, V, H, (const TOK*)tokens, emb, row_kind, t1_out, teff_out);
  return cudaGetLastError();
}
}  // namespace

cudaError_t launch_lengths(const LengthArgs& a, cudaStream_t st) {
  const size_t hist_bytes = 2 * (size_t)a.V * sizeof(int);
  if (hist_bytes > kLen2MaxSmem) return cudaErrorInvalidConfiguration;  // cfg_ok bounds V (kMaxVocab) so that this cannot happen
  if (hist_bytes > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(len2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_bytes);
    if (e != cudaSuccess) return e;
  }
  zero_int_kernel<<<(3 * a.G + 63) / 64, 64, 0, st>>>(a.lens, 3 * a.G);
  switch (a.token_dtype) {
    case IB200_TOK_I64: len1_kernel<long long><<<a.G * a.B, 256, 0, st>>>(a); break;
    case IB200_TOK_I32: len1_kernel<int><<<a.G * a.B, 256, 0, st>>>(a); break;
    case IB200_TOK_I16: len1_kernel<short><<<a.G * a.B, 256, 0, st>>>(a); break;
    case IB200_TOK_U8: len1_kernel<unsigned char><<<a.G * a.B, 256, 0, st>>>(a); break;
    default: return cudaErrorInvalidValue;
  }
  nz_rows_kernel<<<dim3((a.V + 7) / 8, a.G), 256, 0, st>>>(a, a.row_kind);
  len2_kernel<<<a.G * a.B, 128, hist_bytes, st>>>(a, a.row_kind);
  return cudaGetLastError();
}

cudaError_t launch_seq_lengths(int M, int T, int V, int H, const void* tokens, int token_dtype, const float* emb, int* row_kind,
                               int* t1_out, int* teff_out, cudaStream_t st) {
  switch (token_dtype) {
    case IB200_TOK_I64: return launch_seq_lengths_t<long long>(M, T, V, H, tokens, emb, row_kind, t1_out, teff_out, st);
    case IB200_TOK_I32: return launch_seq_lengths_t<int>(M, T, V, H, tokens, emb, row_kind, t1_out, teff_out, st);
    case IB200_TOK_I16: return launch_seq_lengths_t<short>(M, T, V, H, tokens, emb, row_kind, t1_out, teff_out, st);
    case IB200_TOK_U8: return launch_seq_lengths_t<unsigned char>(M, T, V, H, tokens, emb, row_kind, t1_out, teff_out, st);
    default: return cudaErrorInvalidValue;
  }
}

template <int H>
cudaError_t launch_l0_table_h(const TableArgs& a, cudaStream_t st) {
  const size_t smem = sizeof(float) * ((size_t)H * (4 * H + 1) + (size_t)kTableVpb * H);
  const cudaError_t e = cudaFuncSetAttribute(l0_table_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  l0_table_kernel<H><<<dim3((a.V + kPadRows + kTableVpb - 1) / kTableVpb, a.G * 2), 4 * H, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_l0_table(const TableArgs& a, cudaStream_t st) {
  if (a.H == 64) return launch_l0_table_h<64>(a, st);
  if (a.H == 32) return launch_l0_table_h<32>(a, st);
  const int vpb = kTableVpbGeneric;
  dim3 grid((a.V + kPadRows + vpb - 1) / vpb, a.G * 2);
  l0_table_generic_kernel<<<grid, 256, (size_t)vpb * a.H * sizeof(float), st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_pool_fc_fwd(int N, int H, int mode, const float* hn, const float* fc_w, const float* fc_b, float* z,
                               float* pooled_out, uint8_t* argmax_out, cudaStream_t st) {
  pool_fc_fwd_kernel<<<N, 64, H * sizeof(float), st>>>(N, H, mode, hn, fc_w, fc_b, z, pooled_out, argmax_out);
  return cudaGetLastError();
}

cudaError_t launch_pool_fc_bwd(int N, int H, int mode, const float* dz, const float* pooled, const uint8_t* argmax,
                               const float* fc_w, float* d_hn, float* d_fc_w, float* d_fc_b, cudaStream_t st) {
  pool_fc_bwd_dh_kernel<<<N, 64, H * sizeof(float), st>>>(N, H, mode, dz, argmax, fc_w, d_hn);
  pool_fc_bwd_dw_kernel<<<H, 512, 16 * (H + 1) * sizeof(float), st>>>(N, H, dz, pooled, d_fc_w, d_fc_b);
  return cudaGetLastError();
}

}  // namespace ib200
