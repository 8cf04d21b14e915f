/*
 * ib200.h -- C ABI of libib200.so: the B200 (sm_100a) implementation of INTREPPPID's sequence-encoder hot path.
 *
 * This is the drop-in boundary.  Every entry point takes plain structs, raw DEVICE pointers, sizes
 * and a cudaStream_t (as void*); no C++ or torch types cross it.  The caller (PyTorch caching allocator) owns every
 * buffer, including the workspace that carries saved activations from *_fwd to *_bwd.  Launchers are re-entrant, never
 * allocate device memory, never synchronise, and all their work is ordered on the given stream: ib200_encoder_fwd / _bwd may run
 * small independent kernels on one library-owned side stream per device, forked from and joined back into the given stream
 * within the call (IB200_NO_SIDE=1 disables that).
 *
 * Reference interfaces replaced (paths relative to /root/reference/intrepppid/):
 *   ib200_encoder_fwd / _bwd    encoders/awd_lstm.py:147-155 (AWDLSTMEncoder.forward: truncation + embedding dropout)
 *                               utils/embedding_do.py:20-44  (embedding_dropout)
 *                               utils/weightdrop.py:65-111   (WeightDrop._setweights + forward on nn.LSTM)
 *                               encoders/awd_lstm.py:51-56   (AWDLSTM.forward: 2nd truncation + nn.LSTM -> _VF.lstm)
 *   ib200_pool_fc_fwd / _bwd    encoders/awd_lstm.py:58-71   (bi_reduce on h_n[-2:], fc Linear)
 *   ib200_loss_head_fwd / _bwd  e2e/e2e_triplet.py:113-136   (TripletE2ENet.step: triplet projection, TripletMarginLoss,
 *                               classifier/head/mlp.py:35-68  MLPHead, BCEWithLogitsLoss, beta mix)
 *   ib200_batch_metrics         e2e/e2e_triplet.py:171-184   (torchmetrics AUROC / AP / MCC / Precision / Recall of the batch)
 *   ib200_adamw_step            e2e/e2e_triplet.py:231-255   (configure_optimizers: torch.optim.AdamW over self.parameters())
 *   ib200_ranger21_step         e2e/e2e_triplet.py:200-226   (configure_optimizers: ranger21.Ranger21, the factory default; parity unpinned)
 *   ib200_pair_score            e2e/e2e_triplet.py:105-111 + cli/infer.py:216-225 (head + sigmoid over pairs of cached embeddings)
 *   ib200_sequence_lengths      encoders/awd_lstm.py:149-150,53-54 on the batch-of-one calls of cli/infer.py:196-225 (lengths only)
 *   ib200_p2p_allreduce_mean    (no reference counterpart: the reference trains on one device, e2e/e2e_triplet.py:392-400)
 *
 * Conventions
 *   - weights are in PyTorch layout: weight_ih [4H,in], weight_hh [4H,H], gate row-blocks in order (i,f,g,o); two biases.
 *   - E == H (nn.LSTM(embedding_size, embedding_size, ...), awd_lstm.py:35-41).  Supported H: multiples of 32 in [32, 256]
 *     (32 / 64: register-resident recurrent kernels + tcgen05/TMA GEMMs; 96..256: thread-block-cluster recurrent kernels).
 *     2 <= V <= 28672 (the lengths kernel keeps a per-sequence vocabulary histogram in shared memory).
 *   - "group" = one encoder call of the reference.  A training step fuses G=5 calls (anchor, positive, negative, p1, p2 --
 *     e2e_triplet.py:116-129) into one launch set; every group has its own masks and its own truncation lengths.
 *   - masks are INPUTS (nullptr = no drop).  emb_row_scale[g][v] = keep/(1-p) per vocabulary row; whh_l0_mask[g] = scaled
 *     DropConnect mask for weight_hh_l0 (forward direction of layer 0 only).
 *   - return value: 0 ok; <0 invalid argument (IB200_E_*); >0 a cudaError_t.  ib200_last_error() gives text (thread-local).
 */
#ifndef IB200_H_
#define IB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IB200_VERSION 100
#define IB200_MAX_LAYERS 4

enum { IB200_REDUCE_LAST = 0, IB200_REDUCE_MEAN = 1, IB200_REDUCE_MAX = 2 };
/* IB200_PREC_FP32: bf16 hi/lo split operands (3 tensor-core MMAs per product, ~2^-16 operand error), fp32 state, exact-ish
 * exp/rcp activations.  IB200_PREC_BF16: single bf16 MMA per product, tanh.approx activations.  In BOTH modes the saved
 * activations (gates, c) are fp32 and the workspace size is the same; on the plane paths (H = 64 / 128 / 192 / 256) the layer
 * outputs and dgates that feed the GEMMs are stored as bf16 planes (hi | lo in fp32 mode, hi only in bf16 mode). */
enum { IB200_PREC_FP32 = 0, IB200_PREC_BF16 = 1 };
/* Token id storage.  The reference ships int64 ids (data/ppi_oma.py:388-390, 8 B/token over PCIe); ids < V <= 256 fit a byte.
 * Narrow types are read as they are by the lengths kernel (which also makes the int32 working copy), so a caller that feeds
 * uint8 / int16 / int32 ids cuts the per-step H2D traffic 8x / 4x / 2x (SURVEY 8f "input feeding").  Values outside [0, V) are
 * clamped exactly as for int64 (negative int16/int32 -> 0). */
enum { IB200_TOK_I64 = 0, IB200_TOK_I32 = 1, IB200_TOK_I16 = 2, IB200_TOK_U8 = 3 };
enum {
  IB200_E_NULL = -1, IB200_E_SHAPE = -2, IB200_E_UNSUPPORTED = -3, IB200_E_WORKSPACE = -4, IB200_E_ALIGN = -5
};

typedef struct ib200_cfg {
  int32_t G;          /* groups (encoder calls) fused in this launch set */
  int32_t B;          /* sequences per group */
  int32_t T;          /* token row width == trunc_len */
  int32_t V;          /* vocabulary rows */
  int32_t H;          /* embedding size == LSTM hidden size */
  int32_t L;          /* rnn_num_layers (1..IB200_MAX_LAYERS) */
  int32_t bi_reduce;  /* IB200_REDUCE_* ("concat" is not functional in the reference either) */
  int32_t precision;  /* IB200_PREC_* */
  int32_t training;   /* 1: keep activations for ib200_encoder_bwd in the workspace */
  int32_t token_dtype; /* IB200_TOK_*: element type of `tokens` (0 = int64, the reference's dataloader format) */
} ib200_cfg;

/* Parameters of the encoder, PyTorch layout (state_dict names in comments; d=0 forward, d=1 "_reverse"). */
typedef struct ib200_encoder_params {
  const float* emb;                         /* encoder.embedder.weight                [V,H] */
  const float* w_ih[IB200_MAX_LAYERS][2];   /* encoder.encoder.rnn.weight_ih_l{l}{d}  [4H,H] (l=0) / [4H,2H] */
  const float* w_hh[IB200_MAX_LAYERS][2];   /* ...weight_hh_l{l}{d} ([0][0] = weight_hh_l0_raw)  [4H,H] */
  const float* b_ih[IB200_MAX_LAYERS][2];   /* ...bias_ih_l{l}{d}  [4H] */
  const float* b_hh[IB200_MAX_LAYERS][2];   /* ...bias_hh_l{l}{d}  [4H] */
} ib200_encoder_params;

/* Gradients, same shapes; every non-null tensor is OVERWRITTEN (not accumulated).  Dead tensors (top-layer forward
 * chain under bi_reduce=last) are written as exact zeros. */
typedef struct ib200_encoder_grads {
  float* emb;
  float* w_ih[IB200_MAX_LAYERS][2];
  float* w_hh[IB200_MAX_LAYERS][2];
  float* b_ih[IB200_MAX_LAYERS][2];
  float* b_hh[IB200_MAX_LAYERS][2];
} ib200_encoder_grads;

int ib200_version(void);
const char* ib200_last_error(void);

/* Instrumentation (used by bench.py).  ib200_launch_count: kernels launched by this library in this process so far.
 * ib200_timing_enable(1) makes every launcher bracket its kernels with CUDA events on the launching stream;
 * ib200_timing_read sums the elapsed ms and the number of timed launcher calls per kernel family since the last read
 * (arrays of ib200_timing_families() entries, names from ib200_timing_family_name) and synchronises on those events. */
unsigned long long ib200_launch_count(void);
int ib200_timing_enable(int on);
int ib200_timing_families(void);
const char* ib200_timing_family_name(int family);
int ib200_timing_read(int n, float* ms_out, int* launches_out);

/* Bytes of workspace ib200_encoder_fwd needs for cfg (includes everything _bwd needs when cfg.training). 0 on bad cfg. */
size_t ib200_workspace_bytes(const ib200_cfg* cfg);

/*
 * Encoder forward for G groups.
 *   tokens          [G*B, T] row-major of cfg.token_dtype (int64 by default), pad id 0, ids in [0,V)  (data/ppi_oma.py:388-390)
 *   emb_row_scale   float [G,V] or NULL
 *   whh_l0_mask     float [G,4H,H] or NULL
 *   lengths_out     int32 [2,G]: row 0 = T1 (awd_lstm.py:149-150), row 1 = T_eff (awd_lstm.py:53-54); exact integers.
 *                   A group with T_eff==0 is NOT run (the reference raises there, Q13); callers check lengths_out.
 *   hn_top          float [2, G*B, H]: final hidden state of the top layer, [0]=forward dir, [1]=reverse dir.  Under
 *                   bi_reduce=last only [1] is produced ([0] is zero-filled: the dead chain is skipped, Q16).
 */
int ib200_encoder_fwd(const ib200_cfg* cfg, const void* tokens, const ib200_encoder_params* params,
                      const float* emb_row_scale, const float* whh_l0_mask, int32_t* lengths_out, float* hn_top,
                      void* workspace, size_t workspace_bytes, void* stream);

/*
 * Device-side status of the ib200_encoder_fwd call that last filled `workspace`: status_out int32 [3,G] (device memory) =
 * row 0 T1, row 1 T_eff, row 2 flags (IB200_STATUS_*), copied device-to-device on `stream`.  This is how a caller keeps the reference's
 * error behaviour WITHOUT a host sync in the step: F.embedding raises on ids outside [0,V) (utils/embedding_do.py:35-43) and nn.LSTM
 * raises on an all-pad batch (T_eff == 0, encoders/awd_lstm.py:53-56); here the kernels clamp / skip and record, and the
 * caller reads the word whenever it next touches the host (intrepppid_b200/ops.py checks it lazily, one step late).
 */
#define IB200_STATUS_BAD_TOKEN 1 /* a token id outside [0, V) was clamped */
int ib200_encoder_status(const ib200_cfg* cfg, const void* workspace, size_t workspace_bytes, int32_t* status_out, void* stream);

/*
 * Encoder backward.  `workspace` is the buffer the matching _fwd call filled (cfg.training must have been 1).
 *   d_hn_top        float [2, G*B, H] gradient w.r.t. hn_top (entries of a dead direction are ignored)
 *   grads           overwritten; w_hh[0][0] receives the gradient of the RAW tensor (mask applied per group).
 */
int ib200_encoder_bwd(const ib200_cfg* cfg, const ib200_encoder_params* params, const float* emb_row_scale,
                      const float* whh_l0_mask, const float* d_hn_top, const ib200_encoder_grads* grads,
                      void* workspace, size_t workspace_bytes, void* stream);

/*
 * The same backward cut at a layer boundary: runs layers layer_hi, layer_hi-1, ..., layer_lo only (ib200_encoder_bwd is
 * (L-1, 0)).  A full backward is any sequence of calls that covers L-1 .. 0 in descending order on the same workspace and stream
 * (the gradient w.r.t. a layer's input stays in the workspace between calls); only the gradient tensors of the layers run are
 * written (grads->emb with layer 0).  A data-parallel caller uses the cut to start the all-reduce of the upper layers' gradients
 * while the layer-0 BPTT is still running (intrepppid_b200/parallel.py).
 */
int ib200_encoder_bwd_layers(const ib200_cfg* cfg, const ib200_encoder_params* params, const float* emb_row_scale,
                             const float* whh_l0_mask, const float* d_hn_top, const ib200_encoder_grads* grads,
                             void* workspace, size_t workspace_bytes, int32_t layer_hi, int32_t layer_lo, void* stream);

/* bi_reduce + fc.  hn_top [2,N,H] -> z [N,H].  pooled_out [N,H] and (max only) argmax_out uint8 [N,H] are saved for bwd. */
int ib200_pool_fc_fwd(int32_t N, int32_t H, int32_t bi_reduce, const float* hn_top, const float* fc_w, const float* fc_b,
                      float* z, float* pooled_out, uint8_t* argmax_out, void* stream);
/* dz [N,H] -> d_hn_top [2,N,H], d_fc_w [H,H], d_fc_b [H] (all overwritten). */
int ib200_pool_fc_bwd(int32_t N, int32_t H, int32_t bi_reduce, const float* dz, const float* pooled, const uint8_t* argmax,
                      const float* fc_w, float* d_hn_top, float* d_fc_w, float* d_fc_b, void* stream);

/* Triplet + head + BCE.  z is [5,B,H] in group order (anchor, positive, negative, p1, p2). */
typedef struct ib200_head_params {
  const float* fc1_w; /* head.classify.fc1.module.weight_raw [H/2,H] */
  const float* fc1_b; /* [H/2] */
  const float* fc2_w; /* head.classify.fc2.module.weight_raw [1,H/2] */
  const float* fc2_b; /* [1] */
  const float* proj_w; /* triplet_projection.1.weight [H,H] or NULL (use_projection=False) */
  const float* proj_b; /* [H] or NULL */
} ib200_head_params;
typedef struct ib200_head_grads {
  float *fc1_w, *fc1_b, *fc2_w, *fc2_b, *proj_w, *proj_b;
} ib200_head_grads;
typedef struct ib200_head_masks { /* all already scaled by 1/(1-p); NULL = identity */
  const float* fc1_w; /* [H/2,H] */
  const float* do1;   /* [B,H/2] */
  const float* do2;   /* [B,H/2] */
  const float* fc2_w; /* [1,H/2] */
} ib200_head_masks;

/* losses_out float[3] = {loss, classifier_loss, triplet_loss}; y_hat_out float [B] logits.  y: int64 [B] labels. */
int ib200_loss_head_fwd(int32_t B, int32_t H, float beta_classifier, const float* z, const int64_t* y,
                        const ib200_head_params* params, const ib200_head_masks* masks, float* losses_out,
                        float* y_hat_out, void* stream);
/* d_loss: device scalar (upstream gradient of `loss`); d_y_hat: float [B] upstream gradient of the logits or NULL.
 * dz_out [5,B,H] and grads are overwritten.  Recomputes the forward intermediates from z (they are tiny), so no
 * workspace is carried.  (A stand-alone MLPHead backward is d_loss = 0 with d_y_hat given.) */
int ib200_loss_head_bwd(int32_t B, int32_t H, float beta_classifier, const float* z, const int64_t* y,
                        const ib200_head_params* params, const ib200_head_masks* masks, const float* d_loss,
                        const float* d_y_hat, float* dz_out, const ib200_head_grads* grads, void* stream);

/* Eval-mode head + sigmoid over pairs of cached embeddings: prob_out[p] = sigmoid(head(z[idx_a[p]], z[idx_b[p]])).
 * idx_a/idx_b int32 [P]; if both NULL, P must equal M*(M+1)/2 and pairs are the upper triangle (i<=j) in row-major order. */
int ib200_pair_score(int32_t M, int32_t H, const float* z, const int32_t* idx_a, const int32_t* idx_b, int64_t P,
                     const ib200_head_params* params, float* prob_out, void* stream);

/* The same over a contiguous range [p_begin, p_begin + p_count) of the row-major upper triangle: prob_out[k] is the score of flat
 * pair index p_begin + k.  A rank of a multi-GPU inference job scores its block of triangle rows with this (shard the
 * proteins, all-gather the [M,H] embeddings, split the pair matrix by rows). */
int ib200_pair_score_range(int32_t M, int32_t H, const float* z, int64_t p_begin, int64_t p_count,
                           const ib200_head_params* params, float* prob_out, void* stream);

/* Per-sequence truncation lengths of BATCH-OF-ONE eval encoder calls -- what `infer from_csv` does for every protein of every row
 * (cli/infer.py:196-225: each protein goes through TripletE2ENet.forward alone, so it is truncated to its own lengths):
 *   t1_out[m]   = #{t : tokens[m][t] != 0}                                                 (encoders/awd_lstm.py:149-150, a COUNT)
 *   teff_out[m] = max_e #{t < t1_out[m] : emb[tokens[m][t]][e] != 0}                        (encoders/awd_lstm.py:53-54, quirk Q2)
 * Integer bookkeeping for the embedding cache (intrepppid_b200.infer buckets proteins of equal lengths into encoder groups); the
 * encoder entry points recompute both lengths per group.  tokens [M,T] of `token_dtype` (IB200_TOK_*), emb float [V,H], outputs int32
 * [M], scratch int32 [V]; 2 <= V <= 28672 as for the encoder.  Ids outside [0, V) are clamped (the encoder call reports them). */
int ib200_sequence_lengths(int32_t M, int32_t T, int32_t V, int32_t H, const void* tokens, int32_t token_dtype, const float* emb,
                           int32_t* t1_out, int32_t* teff_out, int32_t* scratch, void* stream);

/*
 * Production-mode dropout masks: every Bernoulli(keep)/keep mask of a step in one launch (the reference draws its 14 masks per step
 * with torch's RNG: utils/embedding_do.py:26-28, utils/weightdrop.py:92-102, classifier/head/mlp.py:38-58).  Masks remain
 * plain inputs of the compute entry points; this only replaces the twelve torch kernels that used to fill them.
 *   specs          HOST array: out (device, numel floats), keep_prob in (0,1], row_len > 1 = one draw per row of row_len elements
 *   seed, offset   Philox4x32-10 key / first counter; *counters_used (host, may be NULL) = counters consumed: pass offset + that
 *                  as the next call's offset.  Element i of mask s = (u < keep ? 1/keep : 0), u from draw (i % 4) of counter
 *                  offset + first_counter(s) + i / 4 (oracle/restatement.py::draw_masks is the bit-exact restatement).
 */
typedef struct ib200_mask_spec {
  float* out;
  int64_t numel;
  float keep_prob;
  int32_t row_len;
} ib200_mask_spec;
int ib200_draw_masks(int32_t n_specs, const ib200_mask_spec* specs, uint64_t seed, uint64_t offset, uint64_t* counters_used,
                     void* stream);

/*
 * Per-step classification metrics of TripletE2ENet.step (e2e/e2e_triplet.py:86-90,171-184): the batch values
 * of torchmetrics' (0.11.1) binary AUROC, AveragePrecision, MatthewsCorrCoef(threshold), Precision and Recall, in one launch and
 * without a host sync.  y_hat float [B] logits (squashed with a sigmoid unless every value already lies in [0,1], as torchmetrics
 * does), y int64 [B] labels, 1 <= B <= 1024.  metrics_out float [5] = {auroc, ap, mcc, precision, recall} (a missing class: auroc 0,
 * ap NaN without positives; 0/0 -> 0 for mcc / precision / recall); confusion_out int32 [4] = {tp, fp, tn, fn} or NULL.
 */
int ib200_batch_metrics(int32_t B, const float* y_hat, const int64_t* y, float threshold, float* metrics_out, int32_t* confusion_out,
                        void* stream);

/*
 * Multi-tensor AdamW: the optimizer step right after the hot path (e2e/e2e_triplet.py:231-255 -- torch.optim.AdamW(self.parameters(), lr)
 * with torch defaults betas (0.9, 0.999), eps 1e-8, weight_decay 1e-2).  One launch per 32 tensors instead of
 * torch's per-tensor / foreach kernels; arithmetic follows torch's `_single_tensor_adamw` (decoupled decay, no amsgrad) in fp32.
 *   params / grads / exp_avg / exp_avg_sq   HOST arrays of n_tensors DEVICE pointers (fp32, numel[k] elements each); a tensor whose
 *                                           grads[k] is NULL is skipped (p.grad is None)
 *   hyper.step                              1-based step count of this update (host integer: no device sync)
 *   hyper.grad_scale                        multiplied into every gradient first (1/world_size folds the data-parallel mean into
 *                                           the step when the all-reduce was a SUM; 1.0 otherwise)
 */
typedef struct ib200_adamw_hyper {
  double lr, beta1, beta2, eps, weight_decay, grad_scale; /* doubles: torch derives 1-beta, lr*wd and the bias corrections from Python floats */
  int32_t step;
  int32_t maximize;
} ib200_adamw_hyper;
int ib200_adamw_step(int32_t n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                     float* const* exp_avg_sq, const int64_t* numel, const ib200_adamw_hyper* hyper, void* stream);

/*
 * Multi-tensor Ranger21: the optimizer the reference's factory default selects (e2e/e2e_triplet.py:200-226 -- ranger21.Ranger21(
 * self.parameters(), lr, weight_decay=1e-2, use_warmup, warmdown_active, num_batches_per_epoch, num_epochs, warmdown_start_pct=0.72);
 * intrepppid/__init__.py:37 optimizer_type="ranger21_xx").  Ranger21 is a pinned third-party package
 * (requirements.txt:65, lessw2020/Ranger21 @ 1a96777) whose source is absent from the image: the arithmetic is the published
 * algorithm (arXiv:2106.13731) in the step order of oracle/ranger21_restated.py -- PARITY UNPINNED against the package itself.
 * Covers the AdamW core with positive-negative momentum, adaptive gradient clipping, gradient centralization + normalization, norm
 * loss, stable weight decay, softplus denominator and lookahead; the learning-rate schedule (warm-up / warm-down) is host logic:
 * the caller passes lr(step) per tensor.  Three launches per 24 tensors, each parallel over all tensors (row pass: clipping,
 * centralization, row sums; normalization + variance; update), no host sync.  Like the package, the step rewrites the
 * gradients in place (clipped, centralized, normalized).
 *   tensors      HOST array [n]: DEVICE pointers of one parameter's tensors (fp32, contiguous, rows * cols elements each; a 0-d / 1-d
 *                tensor is one row), `multi_dim` = the parameter has more than one dimension (its rows are centralized; tensors with
 *                3 dimensions are not supported: the package takes their unit norm over dimension 1 only), `step` = 1-based count
 *                of this update for this tensor, `lr` = learning rate of this step; grad_ma / neg_grad_ma are the buffer that
 *                receives this step's momentum and the other one (the package swaps their roles every step: odd steps write
 *                state["grad_ma"]); lookahead may be NULL when hyper.lookahead_merge == 0
 *   scratch      DEVICE memory, 8-byte aligned, ib200_ranger21_scratch_bytes(n, tensors) bytes, ZERO before the first call and owned
 *                by the optimizer between calls: scratch[0] = variance_normalized of this step afterwards (NaN = the package's
 *                RuntimeError), [1] its inverse, [2] an arrival counter, then per-CTA sums, per-row sums and per-row norms
 */
typedef struct ib200_ranger21_tensor {
  float *param, *grad, *grad_ma, *neg_grad_ma, *variance_ma, *lookahead;
  int64_t rows, cols;
  int32_t multi_dim, step;
  double lr;
} ib200_ranger21_tensor;
typedef struct ib200_ranger21_hyper {
  double beta1, beta2, eps, weight_decay, agc_clip, agc_eps, normloss_factor, softplus_beta, pnm_factor, lookahead_alpha;
  int32_t use_agc, use_gc, use_gcnorm, use_normloss, use_softplus;
  int32_t lookahead_merge; /* 1 on the calls that merge the slow weights (every lookahead_mergetime-th) */
} ib200_ranger21_hyper;
int ib200_ranger21_step(int32_t n_tensors, const ib200_ranger21_tensor* tensors, const ib200_ranger21_hyper* hyper, double* scratch,
                        void* stream);
size_t ib200_ranger21_scratch_bytes(int32_t n_tensors, const ib200_ranger21_tensor* tensors); /* only rows / cols are read; 0 on bad input */

/* Data-parallel gradient exchange (the reference itself is single-GPU, e2e/e2e_triplet.py:392-400): one-shot MEAN all-reduce
 * of a small bucket over NVLink peer memory, in place.  Every rank owns a staging region of 2 * stage_floats floats (two parity halves)
 * and a flag array of `world` 32-bit words per bucket, both zero-initialised and mapped by every peer (CUDA IPC; the caller exchanges
 * the handles, e.g. intrepppid_b200.parallel.P2PAllReduce).  The call copies `data` into my half (epoch & 1), publishes the epoch to
 * every rank's flags, waits for all ranks, sums the `world` staged copies straight from the peers' memory and writes the mean to `data`.
 *   stage_ptrs / flag_ptrs   HOST arrays [world] of DEVICE pointers: the base of each rank's staging region / flag array of THIS bucket
 *   epoch                    1, 2, 3, ... : the call count of this bucket, identical on every rank
 * Every rank must issue the same sequence of calls per bucket; world <= 8. */
int ib200_p2p_allreduce_mean(int32_t world, int32_t rank, void* const* stage_ptrs, void* const* flag_ptrs, size_t stage_floats,
                             float* data, size_t n, uint32_t epoch, void* stream);
/* Peer-mappable device memory for the exchange above.  ib200_p2p_alloc: `bytes` of zeroed device memory on the current device + its
 * 64-byte CUDA IPC handle (send it to the other ranks of the node by any host channel).  ib200_p2p_open: map a peer's region into
 * this process for kernels of the CURRENT device (enables peer access to the owning GPU).  _close / _free release them. */
#define IB200_P2P_HANDLE_BYTES 64
int ib200_p2p_alloc(size_t bytes, void** ptr_out, unsigned char* handle_out);
int ib200_p2p_open(const unsigned char* handle, void** ptr_out);
int ib200_p2p_close(void* ptr);
int ib200_p2p_free(void* ptr);

/* Test hook (tests/test_gpu_gemm.py): the token-row NT GEMM in isolation.  impl: 0 legacy mma.sync, 1 tcgen05, 2 auto.
 * C[row,NC] (=|+=) sum_s A_s[row,K] W_s[NC,K]^T (+bias) for rows (n,t) with t < lens[G + n/B] of the [G*B, T] row space. */
int ib200_dbg_gemm_nt(int32_t G, int32_t B, int32_t T, const int32_t* lens, int32_t nsrc, const float* A0, const float* A1,
                      int32_t lda, int32_t K, const float* W0, const float* W1, const float* bias, float* C, int32_t ldc,
                      int32_t NC, int32_t accumulate, int32_t precision, int32_t impl, void* stream);

/* Test hook: the weight-gradient TN GEMM in isolation.  partial[g][cta][KA*NB (+KA column sums)] = sum over the token rows
 * assigned to (g, cta) of A[row,KA]^T Bop[row,NB]; Bop = Bsrc rows shifted by `shift` steps (zero outside [0,T_eff)), or
 * scale[g][tok]*emb[tok] when tok != NULL. */
int ib200_dbg_gemm_tn(int32_t G, int32_t B, int32_t T, const int32_t* lens, const float* A, int32_t KA, const float* Bsrc,
                      int32_t ldb, int32_t col0, int32_t shift, const int32_t* tok, const float* emb, const float* emb_row_scale,
                      int32_t V, int32_t NB, float* partial, int32_t ctas_per_group, int32_t colsum, int32_t precision,
                      int32_t impl, void* stream);

/* Test hooks for the PRODUCTION kernels of the H = 64 path (tests/test_gpu_gemm.py): the TMA-fed tcgen05 GEMMs on bf16 hi|lo plane
 * operands -- a row of K values is stored over the bytes of K floats as [K bf16 hi | K bf16 lo], exactly what the recurrent kernels
 * write.  _nt_planes -> gemm_nt_tma_kernel (xproj / dY), _tn_planes -> gemm_tn_tma_kernel (dW_ih | dW_hh in one pass: first B source
 * dense planes or gathered scale*emb rows, optional second dense source shifted by +-1 step), _l0_grads -> l0_grad_gemm_kernel +
 * reduce + finish (layer-0 dW_hh, dW_ih, bias and embedding gradients from one pass over the dgates). */
int ib200_dbg_gemm_nt_planes(int32_t G, int32_t B, int32_t T, const int32_t* lens, int32_t nsrc, const float* A0, const float* A1,
                             int32_t lda, int32_t K, const float* W0, const float* W1, const float* bias, float* C, int32_t ldc,
                             int32_t NC, int32_t accumulate, int32_t precision, void* stream);
int ib200_dbg_gemm_tn_planes(int32_t G, int32_t B, int32_t T, const int32_t* lens, const float* A, const float* Bsrc, int32_t ldb,
                             int32_t col0, int32_t shift, const int32_t* tok, const float* emb, const float* emb_row_scale, int32_t V,
                             int32_t NB1, const float* Bsrc2, int32_t ldb2, int32_t col02, int32_t shift2, int32_t NB2, float* partial,
                             int32_t ctas_per_group, int32_t precision, void* stream);
size_t ib200_dbg_l0_scratch_floats(int32_t G, int32_t ndir, int32_t which); /* 0: `partial` floats, 1: `scratch` floats */
int ib200_dbg_l0_grads(int32_t G, int32_t B, int32_t T, int32_t V, const int32_t* lens, const int32_t* tok, const float* const* dA,
                       const float* Y0, const float* emb, const float* emb_row_scale, const float* whh_mask, const float* const* w_ih,
                       const float* bias_partial, int32_t bias_count, int32_t dir0, int32_t ndir, float* partial, float* scratch,
                       float* const* d_wih, float* const* d_whh, float* const* d_bih, float* const* d_bhh, float* d_emb,
                       int32_t precision, void* stream);

/* The same for the H = 128 / 192 / 256 path (gemm_wide.cu: both operands streamed by TMA): A_s [rows][K] and W_s [NC][K] plane matrices;
 * _tn_wide: partial [G][splits][KA][NB1 + NB2] (splits <= 0: library default, returned through *splits_out; partial NULL: query only). */
int ib200_dbg_gemm_nt_wide(int32_t G, int32_t B, int32_t T, const int32_t* lens, int32_t nsrc, const float* A0, const float* A1,
                           int32_t lda, int32_t K, const float* W0, const float* W1, const float* bias, float* C, int32_t ldc, int32_t NC,
                           int32_t accumulate, int32_t precision, void* stream);
int ib200_dbg_gemm_tn_wide(int32_t G, int32_t B, int32_t T, const int32_t* lens, const float* A, int32_t KA, const float* Bsrc, int32_t ldb,
                           int32_t col0, int32_t shift, int32_t NB1, const float* Bsrc2, int32_t ldb2, int32_t col02, int32_t shift2,
                           int32_t NB2, float* partial, int32_t splits, int32_t* splits_out, int32_t precision, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IB200_H_ */
