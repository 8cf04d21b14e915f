// Internal launcher interface between capi.cu and the kernel translation units.
#pragma once
#include "../../include/ib200.h"
#include "common.cuh"

namespace ib200 {

// ---- K0: truncation lengths + int32 token copy (A1/A3) ------------------------------------------------------------------
struct LengthArgs {
  int G, B, Tin, V, H;
  const void* tokens;          // [G*B, Tin] of token_dtype (IB200_TOK_*)
  int token_dtype;
  const float* emb;            // [V,H]
  const float* emb_row_scale;  // [G,V] or null
  int* tok32;                  // [G*B, Tin]
  int* lens;                   // [3,G]: T1, T_eff, status flags (kStatus*), all zeroed by the launcher
  int* row_kind;               // scratch [G,V]
};
constexpr int kStatusBadToken = 1;            // a token id outside [0, V) was clamped (F.embedding raises on it in the reference)
constexpr size_t kLen2MaxSmem = 224 * 1024;   // len2_kernel keeps a [V] histogram + a [V] list in shared memory
constexpr int kMaxVocab = (int)(kLen2MaxSmem / (2 * sizeof(int)));  // = 28672 vocabulary rows
cudaError_t launch_lengths(const LengthArgs& a, cudaStream_t st);
// per-sequence (T1, T_eff) of batch-of-one eval calls; row_kind: scratch [V]
cudaError_t launch_seq_lengths(int M, int T, int V, int H, const void* tokens, int token_dtype, const float* emb, int* row_kind,
                               int* t1_out, int* teff_out, cudaStream_t st);

// ---- K1a: layer-0 input-projection table P[g][d][v][4H] (GI order) --------------------------------------------------------
// Pad replicas: a padded batch makes every short sequence read vocabulary row 0 at the same step (all of them at once at the start
// of the reverse scan), i.e. hundreds of requests per step for the same few L2 lines.  The table therefore carries kPadRows extra
// copies of row 0 (rows V .. V+kPadRows-1) and every CTA of the recurrent kernel redirects its pad ids to "its" copy.
constexpr int kPadRows = 64;
struct TableArgs {
  int G, V, H;
  const float* emb;            // [V,H]
  const float* emb_row_scale;  // [G,V] or null
  const float* w_ih[2];        // [4H,H]
  const float* b_ih[2];
  const float* b_hh[2];
  float* table;                // [G,2,V+kPadRows,4H]
};
cudaError_t launch_l0_table(const TableArgs& a, cudaStream_t st);

// ---- weight preparation: GI-permuted copies used by the GEMMs ------------------------------------------------------------
// out_w[gi][k] = w[torch_row(gi)][k]  (rows permuted), out_wT[k][gi] = same transposed, out_b[gi] = b_ih + b_hh permuted
cudaError_t launch_prep_wih(const float* w, const float* b_ih, const float* b_hh, int H, int K, float* out_w, float* out_wT,
                            float* out_b, cudaStream_t st);

// ---- K2: recurrent forward ----------------------------------------------------------------------------------------------
struct LstmFwdArgs {
  int G, B, Tmax, V;
  int dir0, ndir;            // directions run: dir0 .. dir0+ndir-1 (grid.z)
  const int* lens;           // [2,G]
  const int* tok;            // [N,Tmax] (layer 0) or null
  const float* table;        // layer 0: [G,2,V+kPadRows,4H] (rows >= V are copies of row 0)
  int table_shared;          // 1: no per-group row scale (eval / p = 0) -- all groups read the table of group 0
  const float* xproj[2];     // layer >= 1: per direction [N,Tmax,4H] (GI)
  const float* whh[2];       // [4H,H] per direction
  const float* whh_mask;     // [G,4H,H] or null; direction 0 only
  float* y;                  // [N,Tmax,y_stride] (+dir*H) or null
  int y_stride;
  int planes;                // 1: y rows hold bf16 [hi plane (y_stride) | lo plane (y_stride)] instead of y_stride floats (TMA GEMMs)
  float* gates[2];           // per direction [N,Tmax,H,4] (training) or null
  float* cstate[2];          // per direction [N,Tmax,H]
  float* hn;                 // [2,N,H] or null
  int dbg;                   // ablation flags for timing experiments (env IB200_DBG; 0 in production)
  PhaseArgs ph;              // two-phase rebalancing (common.cuh); buffers null => always one phase
};
cudaError_t launch_lstm_fwd(const LstmFwdArgs& a, int H, int precision, cudaStream_t st);

// ---- K3: recurrent backward (BPTT); dgates overwrite the saved gates in place ------------------------------------------------
struct LstmBwdArgs {
  int G, B, Tmax;
  int dir0, ndir;
  const int* lens;
  const float* whh[2];
  const float* whh_mask;
  float* gates[2];           // in: (i,f,g,o) ; out: (da_i,da_f,da_g,da_o), GI order
  const float* cstate[2];
  const float* dy;           // [N,Tmax,dy_stride] (+dir*H) gradient w.r.t. this layer's output, or null
  int dy_stride;
  const float* dhn;          // [2,N,H] gradient w.r.t. the final hidden state, or null
  int planes;                // 1: dgates rows are written as bf16 [hi plane (4H) | lo plane (4H)] over the 4H-float gate row
  float* bias_partial;       // planes mode: [ndir][G*gridDim.x][4H] per-CTA column sums of the dgates (GI order)
  int dbg;                   // ablation flags (env IB200_DBG)
  PhaseArgs ph;              // two-phase rebalancing (common.cuh); buffers null => always one phase
};
cudaError_t launch_lstm_bwd(const LstmBwdArgs& a, int H, int precision, cudaStream_t st);
int lstm_bwd_cta_count(const LstmBwdArgs& a, int precision);  // G * gridDim.x of the launch above (bias partials per direction)

// ---- K2c / K3c: cluster versions for H = 32k, 32 <= H <= 256 (lstm_cluster.cu): W_hh sliced over H/32 CTAs, h exchanged over DSMEM.
// planes = 0: fp32 y / dgates, bias gradients from the TN GEMM's column sums; planes = 1: y and dgates as bf16 hi|lo planes (the
// operands of gemm_wide.cu), bias partials [ndir][lstm_bwd_cluster_cta_count][4H] written by the BPTT kernel.
bool lstm_cluster_supports(int H);
int lstm_bwd_cluster_cta_count(const LstmBwdArgs& a, int H, int precision);
cudaError_t launch_lstm_fwd_cluster(const LstmFwdArgs& a, int H, int precision, cudaStream_t st);
cudaError_t launch_lstm_bwd_cluster(const LstmBwdArgs& a, int H, int precision, cudaStream_t st);

// ---- K2t / K3t: tcgen05 cluster versions for H = 128 / 256 (lstm_cluster_tc.cu): W_hh slice resident as a UMMA operand, gates^T /
// dh^T accumulated in TMEM, h all-gather by bulk DSMEM copies, dh reduce-scatter by st.async.  Same arguments / layouts as above;
// 32 sequences per cluster.
bool lstm_cluster_tc_supports(int H);
int lstm_bwd_cluster_tc_cta_count(const LstmBwdArgs& a, int H, int precision);
cudaError_t launch_lstm_fwd_cluster_tc(const LstmFwdArgs& a, int H, int precision, cudaStream_t st);
cudaError_t launch_lstm_bwd_cluster_tc(const LstmBwdArgs& a, int H, int precision, cudaStream_t st);

// ---- tensor-core GEMMs over token rows -------------------------------------------------------------------------------------
// NT:  C[row, NC] (=|+=) sum_s A_s[row, K] * W_s[NC, K]^T (+ bias)   for rows (n,t), t < T_eff[group(n)]
struct GemmNTArgs {
  int G, B, Tmax;
  const int* lens;
  int nsrc;                  // 1 or 2 A/W source pairs (K-concatenation, e.g. both directions' dgates)
  const float* A[2];         // [N*Tmax, lda]
  int lda;
  int K;                     // per source
  const float* W[2];         // [NC, K] row-major
  const float* bias;         // [NC] or null
  float* C;                  // [N*Tmax, ldc]
  int ldc;
  int NC;
  int accumulate;            // C += instead of C =
  int plane_bytes;           // TMA path: A rows hold bf16 planes; byte offset of the lo plane inside a row (= K*2); 0 = fp32 rows
};
cudaError_t launch_gemm_nt(const GemmNTArgs& a, int precision, cudaStream_t st);      // legacy mma.sync (any supported H)
// tcgen05 version (H=64 shapes); returns cudaErrorInvalidConfiguration when the shape / smem budget is not covered
cudaError_t launch_gemm_nt_tc(const GemmNTArgs& a, int precision, cudaStream_t st);
// tcgen05 + TMA version: A_s rows are bf16 hi/lo planes written by the recurrent kernels (plane_bytes > 0)
cudaError_t launch_gemm_nt_tma(const GemmNTArgs& a, int precision, cudaStream_t st);

// tcgen05 + TMA with BOTH operands streamed (gemm_wide.cu; H = 128, 192, 256): A_s rows and W_s ([NC][K], from launch_prep_wih_planes)
// are bf16 hi|lo plane matrices
bool gemm_wide_supports(int H);
cudaError_t launch_gemm_nt_wide(const GemmNTArgs& a, int precision, cudaStream_t st);
cudaError_t launch_prep_wih_planes(const float* w, const float* b_ih, const float* b_hh, int H, int K, float* out_w, float* out_wT,
                                   float* out_b, int precision, cudaStream_t st);
// layer-0 input rows scale[g][tok] * emb[tok] as planes [N*Tmax][H] (zeros in the tail rows of the last 64-row box)
cudaError_t launch_gather_x0_planes(int G, int B, int Tmax, int V, int H, const int* lens, const int* tok, const float* emb,
                                    const float* scale, float* out, int precision, cudaStream_t st);

// TN:  P[cta][KA, NB] = sum_{rows of cta} A[row, KA]^T * Bop[row, NB]   (partials; reduced by launch_dw_reduce)
struct GemmTNArgs {
  int G, B, Tmax;
  const int* lens;
  const float* A;            // dgates [N*Tmax, lda] (GI); the GEMM uses columns [a_col0, a_col0 + KA)
  int KA;                    // 4H, or a column chunk of it (legacy kernel: lda / a_col0 select the chunk)
  int lda, a_col0;           // legacy kernel only; lda == 0 means lda = KA, a_col0 = 0
  // B operand: either dense rows Bsrc[(n, t+shift), col0 .. col0+NB) with zero outside [0,T_eff), or gathered embeddings
  const float* Bsrc;
  int ldb, col0, shift;
  const int* tok;            // if non-null: Bop[row] = scale[g][tok] * emb[tok] (layer-0 input), Bsrc/ldb unused
  const float* emb;
  const float* emb_row_scale;
  int V;
  int emb_ld, emb_col0;      // legacy kernel only: embedding row pitch and first column (emb_ld == 0 means NB, 0)
  int NB;
  // optional second dense source for the columns [NB1, NB) (tcgen05 kernel only): fuses dW_ih and dW_hh into one pass over A
  int NB1;                   // columns taken from the first source (== NB when there is no second source)
  const float* Bsrc2;
  int ldb2, col02, shift2;
  float* partial;            // [G][ctas_per_group][KA*NB (+ KA if colsum)]
  int ctas_per_group;
  int colsum;                // also produce column sums of A (bias gradient) after the KA*NB block
};
cudaError_t launch_gemm_tn(const GemmTNArgs& a, int precision, cudaStream_t st);      // legacy mma.sync
cudaError_t launch_gemm_tn_tc(const GemmTNArgs& a, int precision, cudaStream_t st);   // tcgen05 (KA=256, NB=128|64)
// tcgen05 + TMA: A = dgates planes (row = [hi 4H bf16 | lo 4H bf16]); Bsrc = planes rows of ldb floats ([hi ldb bf16 | lo ldb bf16])
// or gathered embeddings (tok != null); no colsum (the BPTT kernel produces the bias partials)
cudaError_t launch_gemm_tn_tma(const GemmTNArgs& a, int precision, cudaStream_t st);

// wide version (gemm_wide.cu): KA = 4H, column tiles of H, split-K over a.ctas_per_group = gemm_tn_wide_splits(...) CTAs per tile;
// partial [G][splits][KA][NB]; dense plane sources only (no gather, no colsum)
int gemm_tn_wide_splits(int KA, int NB, int BN, int G);
cudaError_t launch_gemm_tn_wide(const GemmTNArgs& a, int precision, cudaStream_t st);

// out[torch_row(gi)][c] = sum_g mask_g[torch_row][c] * sum_cta partial[g][cta][gi][c];  optional bias outputs
struct DwReduceArgs {
  int G, ctas_per_group, KA, NB, H;
  int gi0, c0, ldo;          // chunked GEMMs: first GI row / first column of this block inside the full [4H, ldo] result (ldo == 0: NB)
  const float* partial;
  int has_colsum;            // 1: column sums follow each KA*NB block of `partial` (legacy / thread-loader GEMMs)
  const float* cs_ptr;       // alternative column-sum partials [cs_count][KA] written by the BPTT kernel (TMA path), or null
  int cs_count;
  const float* mask;         // [G, KA, NB] in torch row order, or null
  float* out;                // [KA, NB1] torch row order (NB1 == NB unless split)
  int NB1;                   // columns [0,NB1) -> out (no mask); columns [NB1,NB) -> out2 (mask applies to these)
  float* out2;               // [KA, NB-NB1] or null
  float* out_b1;             // [KA] or null (bias_ih grad)
  float* out_b2;             // [KA] or null (bias_hh grad, identical values)
};
cudaError_t launch_dw_reduce(const DwReduceArgs& a, cudaStream_t st);

// embedding gradient: demb[v] = sum_g scale[g][v] * sum_{(n,t): tok=v, t<T_eff} dx[n,t]   (row 0 = padding_idx gets zero)
struct EmbGradArgs {
  int G, B, Tmax, V, H;
  const int* lens;
  const int* tok;
  const float* dx;           // [N*Tmax, H]
  const float* emb_row_scale;
  float* demb;               // [V,H], zero-filled by the launcher
};
cudaError_t launch_emb_grad(const EmbGradArgs& a, cudaStream_t st);

// ---- layer-0 parameter gradients of the H = 64 planes path in one pass over the layer-0 dgates (gemm_l0.cu) --------------------------
struct L0GradArgs {
  int G, B, Tmax, V, H;
  int dir0, ndir;              // live directions of layer 0: dir0 .. dir0 + ndir - 1
  const int* lens;
  const int* tok;              // [N, Tmax]
  const float* dA[2];          // per direction: dgates planes (rows of 4H floats holding [hi 4H bf16 | lo 4H bf16])
  const float* Y0;             // layer-0 output planes (rows of 2H floats holding [hi 2H bf16 | lo 2H bf16])
  const float* emb;            // [V, H]
  const float* emb_row_scale;  // [G, V] or null
  const float* whh_mask;       // [G, 4H, H] or null (forward direction)
  const float* w_ih[2];        // [4H, H]
  const float* bias_partial;   // [ndir][bias_count][4H] column sums of the dgates left by the BPTT kernel
  int bias_count;
  float* partial;              // l0_grad_partial_floats() floats
  float* R;                    // l0_grad_scratch_floats() floats
  int ctas_per_group;          // (set by the launcher)
  float* d_wih[2];
  float* d_whh[2];
  float* d_bih[2];
  float* d_bhh[2];
  float* d_emb;                // [V, H] (overwritten)
};
size_t l0_grad_scratch_floats(int G, int ndir);
size_t l0_grad_partial_floats(int G, int ndir);
// cudaErrorInvalidConfiguration when the shape is not covered (H != 64 or V > 256): the caller keeps the general path
cudaError_t launch_l0_grads(const L0GradArgs& a, int precision, cudaStream_t st);

cudaError_t launch_fill_zero(float* p, size_t n, cudaStream_t st);

// ---- production-mode mask generation (masks.cu) --------------------------------------------------------------------------------
constexpr int kMaskMaxSpecs = 8;  // masks per launch (table travels in the kernel parameters)
cudaError_t launch_draw_masks(int n, const ib200_mask_spec* specs, unsigned long long seed, unsigned long long offset,
                              unsigned long long* counters_used, cudaStream_t st, int* launches);

// ---- per-step classification metrics (metrics.cu): out[5] = auroc, ap, mcc, precision, recall; conf[4] = tp, fp, tn, fn ---------
cudaError_t launch_batch_metrics(int B, const float* y_hat, const long long* y, float threshold, float* out, int* conf, cudaStream_t st);

// ---- one-shot mean all-reduce over NVLink peer memory (p2p.cu) -------------------------------------------------------------------
constexpr int kP2PMaxWorld = 8;
struct P2PArgs {
  int world, rank;
  const float* stage[kP2PMaxWorld];   // every rank's staged copy of this bucket (peer-mapped device pointers; [rank] is my own)
  uint32_t* flags[kP2PMaxWorld];      // every rank's flag array of this bucket ([world] words each)
  float* data;                        // my bucket: n floats, reduced in place
  size_t n;
  uint32_t epoch;                     // 1, 2, 3, ... per bucket
};
cudaError_t launch_p2p_allreduce_mean(const P2PArgs& a, float* my_stage, cudaStream_t st);

// ---- multi-tensor AdamW (optim.cu) -------------------------------------------------------------------------------------------
constexpr int kAdamMaxTensors = 32;  // tensors per launch (pointer table travels in the kernel parameters)
struct AdamScalars {
  float grad_scale, decay, one_minus_b1, b2, one_minus_b2, bc2_sqrt, eps, step_size;
};
// pointer arrays are HOST arrays of device pointers; *launches = kernels launched
cudaError_t launch_adamw(int n, float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                         const long long* numel, const AdamScalars& s, cudaStream_t st, int* launches);

// ---- multi-tensor Ranger21 (ranger21.cu) -------------------------------------------------------------------------------------
constexpr int kR21MaxTensors = 24;  // tensors per launch (the table travels in the kernel parameters: 24 x 112 B)
struct R21Tensor {
  float* p; float* g; float* grad_ma; const float* neg_grad_ma; float* v; float* slow;
  float* pnorm;               // [rows] scratch: ||p_row|| of this step (rows kernel -> update kernel)
  double* rowsum;             // [rows][2] scratch: sum / sum of squares of the clipped, centralized gradient row
  long long numel;
  double inv_bc2, wd_lr;      // 1 / (1 - b2^step); weight_decay * lr(step)
  int rows, cols, multi_dim;  // rows x cols view (one row for 0-d / 1-d tensors); multi_dim: dim() > 1 (rows are centralized)
  float lr, sqrt_bc2, step_size;  // lr(step) after warm-up / warm-down; sqrt(1 - b2^step); lr / (1 - b1^step)
};
struct R21Scalars {
  double param_size;  // elements of all tensors of this step
  float b2, one_minus_b2, b1sq, one_minus_b1sq, eps, agc_clip, agc_eps, normloss2, softplus_beta, pnm_factor, one_plus_pnm,
      inv_noise_norm, la_alpha, one_minus_la_alpha;
  int use_agc, use_gc, use_gcnorm, use_normloss, use_softplus, use_decay, lookahead_merge;
};
long long ranger21_elem_ctas(long long numel);  // CTAs of the elementwise kernels for one tensor
// tensors: HOST array (pnorm / rowsum already point into the scratch); scratch: DEVICE doubles -- [0] variance_normalized, [1] its
// inverse, [2] the arrival counter (zero between steps), [3 ..] one sum per elementwise CTA; *launches = kernels launched
cudaError_t launch_ranger21(int n, const R21Tensor* tensors, const R21Scalars& s, double* scratch, cudaStream_t st, int* launches);

}  // namespace ib200
