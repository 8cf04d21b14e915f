// C ABI of libib200.so (see include/ib200.h): argument validation, workspace carve-up and the launch sequences.
#include <algorithm>
#include <cmath>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ib200.h"
#include "kernels.h"
#include "small.h"

using namespace ib200;

namespace {

thread_local std::string g_err;

int fail(int code, const char* msg) {
  g_err = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* where) {
  g_err = std::string(where) + ": " + cudaGetErrorString(e);
  return (int)e;
}
#define CK(call, where)                                  \
  do {                                                   \
    cudaError_t e__ = (call);                            \
    if (e__ != cudaSuccess) return cuda_fail(e__, where); \
  } while (0)

inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

// ---- instrumentation: launch counter (always on) and optional CUDA-event timing per kernel family (bench.py) ----------------
enum Family { F_LENGTHS, F_L0_TABLE, F_PREP, F_LSTM_FWD_L0, F_LSTM_FWD_UP, F_GEMM_XPROJ, F_LSTM_BWD_UP, F_LSTM_BWD_L0, F_GEMM_DW,
              F_DW_REDUCE, F_GEMM_DGRAD, F_EMB_GRAD, F_POOL_FC, F_LOSS_HEAD, F_PAIR_SCORE, F_FILL, F_ADAMW, F_METRICS, F_MASKS, F_L0_GRADS, F_ALLREDUCE, F_RANGER21, F_COUNT };
const char* kFamilyNames[F_COUNT] = {"lengths", "l0_table", "prep_wih", "lstm_fwd_l0", "lstm_fwd_upper", "gemm_nt_xproj",
                                     "lstm_bwd_upper", "lstm_bwd_l0", "gemm_tn_dw", "dw_reduce", "gemm_nt_dgrad", "emb_grad",
                                     "pool_fc", "loss_head", "pair_score", "fill_zero", "adamw", "batch_metrics", "draw_masks", "l0_grads", "p2p_allreduce", "ranger21"};
std::atomic<unsigned long long> g_launches{0};
struct TimingState {
  std::mutex mu;
  bool on = false;
  std::vector<cudaEvent_t> pool;   // pairs (start, stop)
  std::vector<int> fam;            // family of pair i
  size_t used = 0;                 // pairs in use
} g_timing;

struct TimedScope {
  cudaEvent_t stop = nullptr;
  cudaStream_t st;
  TimedScope(int fam, int nkernels, cudaStream_t s) : st(s) {
    g_launches.fetch_add((unsigned long long)nkernels);
    if (!g_timing.on) return;
    std::lock_guard<std::mutex> lk(g_timing.mu);
    if (g_timing.used * 2 + 2 > g_timing.pool.size()) {
      cudaEvent_t a, b;
      if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
      g_timing.pool.push_back(a);
      g_timing.pool.push_back(b);
      g_timing.fam.push_back(fam);
    }
    g_timing.fam[g_timing.used] = fam;
    cudaEventRecord(g_timing.pool[g_timing.used * 2], st);
    stop = g_timing.pool[g_timing.used * 2 + 1];
    ++g_timing.used;
  }
  ~TimedScope() {
    if (stop) cudaEventRecord(stop, st);
  }
};
#define TIMED(fam, nk, call, where)  \
  do {                               \
    TimedScope ts__(fam, nk, st);    \
    CK(call, where);                 \
  } while (0)

// Workspace layout.  Everything the backward needs is inside, so the caller only keeps one buffer alive.
struct Plan {
  int G, B, T, V, H, L, N;
  size_t R;  // token rows = N*T
  bool train;
  bool live[IB200_MAX_LAYERS][2];
  size_t lens, tok32, table, row_kind;
  size_t wih_gi[IB200_MAX_LAYERS][2], b_gi[IB200_MAX_LAYERS][2], wihT_gi[IB200_MAX_LAYERS][2];
  size_t Y[IB200_MAX_LAYERS];
  size_t X[2];
  size_t gates[IB200_MAX_LAYERS][2], cst[IB200_MAX_LAYERS][2];
  size_t bwd_scratch;  // dY [R,2H] then dX0 [R,H]
  size_t partial;
  size_t bias_partial[2];  // planes mode: [2][G*ceil(B/4)][4H] per-CTA dgate column sums from the BPTT kernel; one buffer per layer
                           // parity: layer l's reduce (side stream) reads its sums while layer l-1's BPTT kernel writes the other
  size_t l0_scratch;  // R of gemm_l0.cu
  size_t x0p;         // wide tcgen05 path: layer-0 input rows scale*emb[tok] as planes [R,H] (B operand of the layer-0 dW_ih GEMM)
  size_t ph_state, ph_sched;  // two-phase rebalancing (common.cuh): [2][N][H][2] floats; sm_load[256] + resume_list[1 + CTAs] ints
  int ctas_per_group;
  size_t total;
};

// SMs of the current device (148 on B200); 148 when no device can be queried (ib200_workspace_bytes on a CPU-only host)
int sm_count() {
  thread_local int cached_dev = -1, cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return 148; }
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
    else { (void)cudaGetLastError(); cached = 148; }
    cached_dev = dev;
  }
  return cached;
}

// ablation flags for timing experiments (tools/ablate_*.py); read once per process
int dbg_flags() {
  static const int v = [] { const char* e = getenv("IB200_DBG"); return e ? atoi(e) : 0; }();
  return v;
}

bool cfg_ok(const ib200_cfg* c) {
  if (!(c && c->G >= 1 && c->B >= 1 && c->T >= 1 && c->V >= 2 && (c->H == 32 || c->H == 64 || lstm_cluster_supports(c->H)) && c->L >= 1 &&
        c->L <= IB200_MAX_LAYERS && c->bi_reduce >= 0 && c->bi_reduce <= 2 && (c->precision == 0 || c->precision == 1) &&
        c->token_dtype >= IB200_TOK_I64 && c->token_dtype <= IB200_TOK_U8))
    return false;
  // token rows are indexed with 32-bit integers inside the kernels; ids are staged as uint16 by the H <= 64 layer-0 kernel, whose
  // per-CTA token stage ((T + 4) x 16 bytes of shared memory) bounds trunc_len
  if ((long long)c->G * c->B * c->T >= (1LL << 31)) return false;
  if (c->V > kMaxVocab) return false;  // the lengths kernel keeps a [V] histogram + a [V] row list in shared memory (small.cu)
  if ((c->H == 32 || c->H == 64) && c->T > 11000) return false;
  return true;
}

bool wide_tc(int H);  // (defined with the GEMM dispatch below)

Plan make_plan(const ib200_cfg* c) {
  Plan p{};
  p.G = c->G; p.B = c->B; p.T = c->T; p.V = c->V; p.H = c->H; p.L = c->L;
  p.N = c->G * c->B;
  p.R = (size_t)p.N * p.T;
  p.train = c->training != 0;
  const int H = p.H;
  for (int l = 0; l < p.L; ++l) {
    p.live[l][0] = !(l == p.L - 1 && c->bi_reduce == IB200_REDUCE_LAST);  // dead chain under "last" 
    p.live[l][1] = true;
  }
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
  p.lens = take(sizeof(int) * 3 * p.G);  // T1 | T_eff | status flags
  p.tok32 = take(sizeof(int) * p.R);
  p.row_kind = take(sizeof(int) * (size_t)p.G * p.V);
  p.table = take(sizeof(float) * (size_t)p.G * 2 * (p.V + kPadRows) * 4 * H);
  for (int l = 0; l < p.L; ++l)
    for (int d = 0; d < 2; ++d) {
      const size_t K = l == 0 ? H : 2 * H;
      if (l > 0) {
        p.wih_gi[l][d] = take(sizeof(float) * 4 * H * K);
        p.b_gi[l][d] = take(sizeof(float) * 4 * H);
      }
      if (p.train) p.wihT_gi[l][d] = take(sizeof(float) * 4 * H * K);
    }
  p.ph_state = take(sizeof(float) * 2 * (size_t)p.N * H * 2);
  p.ph_sched = take(sizeof(int) * (256 + 1 + 2 * (size_t)p.G * ((p.B + 3) / 4)));
  for (int l = 0; l < p.L; ++l)
    if (p.train || l < p.L - 1) p.Y[l] = take(sizeof(float) * p.R * 2 * H);
  if (p.L > 1)
    for (int d = 0; d < 2; ++d) p.X[d] = take(sizeof(float) * p.R * 4 * H);
  if (p.train) {
    for (int l = 0; l < p.L; ++l)
      for (int d = 0; d < 2; ++d)
        if (p.live[l][d]) {
          p.gates[l][d] = take(sizeof(float) * p.R * 4 * H);
          p.cst[l][d] = take(sizeof(float) * p.R * H);
        }
    p.bwd_scratch = p.L > 1 ? p.X[0] : take(sizeof(float) * p.R * 3 * H);  // X is dead once the forward is done
    p.ctas_per_group = std::max(1, sm_count() / p.G);
    p.partial = take(sizeof(float) * std::max((size_t)p.G * p.ctas_per_group * ((size_t)4 * H * 3 * H + 4 * H),  // up to [dW_ih | dW_hh] fused
                                              l0_grad_partial_floats(p.G, 2)));
    p.l0_scratch = take(sizeof(float) * l0_grad_scratch_floats(p.G, 2));
    if (wide_tc(H)) p.x0p = take(sizeof(float) * p.R * H);
    for (int i = 0; i < 2; ++i) p.bias_partial[i] = take(sizeof(float) * 2 * (size_t)p.G * ((p.B + 3) / 4) * 4 * H);
  }
  p.total = off;
  return p;
}

// GEMM dispatch: tcgen05 kernels where the shape is covered, legacy mma.sync otherwise (H=32 test shapes; IB200_GEMM=legacy)
// IB200_GEMM = "tma" (default: bf16 hi/lo planes + TMA-fed tcgen05), "threads" (fp32 storage, thread-staged tcgen05), "legacy" (mma.sync)
int gemm_mode() {
  static const int v = [] {
    const char* e = getenv("IB200_GEMM");
    if (e && std::string(e) == "legacy") return 0;
    if (e && std::string(e) == "threads") return 1;
    return 2;
  }();
  return v;
}
// recurrent kernels: register-resident W_hh for H = 32 / 64, thread-block clusters (lstm_cluster.cu) for every other supported H.
// IB200_FORCE_CLUSTER=1 routes H = 32 / 64 through the cluster kernels too (cross-check of the two implementations in the tests).
bool use_cluster(int H) {
  static const bool force = [] { const char* e = getenv("IB200_FORCE_CLUSTER"); return e && atoi(e) != 0; }();
  return force || (H != 32 && H != 64);
}
bool use_tc() { return gemm_mode() >= 1; }
// cluster path on the tcgen05 recurrent kernels (lstm_cluster_tc.cu): H = 128 / 256 with plane operands; IB200_CLUSTER_TC=0 keeps the
// mma.sync cluster kernels (the tests cross-check the two)
// (2 / 3: only the forward / only the backward kernel -- the HBM layouts are identical, so the two families mix freely)
bool use_cluster_tc(int H, bool planes, bool bwd) {
  static const int mode = [] { const char* e = getenv("IB200_CLUSTER_TC"); return e ? atoi(e) : 1; }();
  const bool on = mode == 1 || (mode == 2 && !bwd) || (mode == 3 && bwd);
  return on && planes && use_cluster(H) && lstm_cluster_tc_supports(H);
}
// hidden sizes of the cluster path whose GEMMs run on the TMA-fed tcgen05 kernels of gemm_wide.cu (H = 128, 192, 256); the other
// cluster sizes (96, 160, 224) and IB200_GEMM=legacy keep the column-blocked mma.sync kernels
bool wide_tc(int H) { return gemm_wide_supports(H) && gemm_mode() == 2; }
bool use_planes(int H) { return (H == 64 && gemm_mode() == 2 && !use_cluster(H)) || wide_tc(H); }
cudaError_t gemm_nt_auto(const GemmNTArgs& a, int prec, cudaStream_t st, bool wide = false) {
  if (wide && a.plane_bytes > 0) return launch_gemm_nt_wide(a, prec, st);  // H = 128 / 192 / 256: streamed-operand tcgen05 kernel
  if (wide) return launch_gemm_nt(a, prec, st);  // other H > 64: column-blocked legacy kernel
  if (a.plane_bytes > 0) {  // operands are bf16 planes: only the TMA kernels can read them
    cudaError_t e = launch_gemm_nt_tma(a, prec, st);
    if (e != cudaErrorInvalidConfiguration || a.nsrc != 2) return e;
    (void)cudaGetLastError();
    GemmNTArgs b = a;  // W of both sources does not fit: one source per pass, the second pass accumulates
    b.nsrc = 1;
    e = launch_gemm_nt_tma(b, prec, st);
    if (e != cudaSuccess) return e;
    b.A[0] = a.A[1]; b.W[0] = a.W[1]; b.accumulate = 1; b.bias = nullptr;
    return launch_gemm_nt_tma(b, prec, st);
  }
  if (use_tc()) {
    cudaError_t e = launch_gemm_nt_tc(a, prec, st);
    if (e != cudaErrorInvalidConfiguration) return e;
    (void)cudaGetLastError();
    if (a.nsrc == 2) {  // W of both sources does not fit in shared memory: one source per pass, second pass accumulates
      GemmNTArgs b = a;
      b.nsrc = 1;
      e = launch_gemm_nt_tc(b, prec, st);
      if (e == cudaSuccess) {
        b.A[0] = a.A[1]; b.W[0] = a.W[1]; b.accumulate = 1; b.bias = nullptr;
        return launch_gemm_nt_tc(b, prec, st);
      }
      if (e != cudaErrorInvalidConfiguration) return e;
      (void)cudaGetLastError();
    }
  }
  return launch_gemm_nt(a, prec, st);
}

cudaError_t gemm_tn_auto(const GemmTNArgs& a, int prec, cudaStream_t st, bool planes = false) {
  if (planes) return launch_gemm_tn_tma(a, prec, st);
  if (use_tc()) {
    cudaError_t e = launch_gemm_tn_tc(a, prec, st);
    if (e != cudaErrorInvalidConfiguration) return e;
    (void)cudaGetLastError();
  }
  return launch_gemm_tn(a, prec, st);
}

// dW block = dA^T * Bop over all token rows: per-CTA partials + deterministic reduction.  `wide` (H > 64): the [KA, NB] product is
// cut into blocks the mma.sync kernel covers (KA 256|128 x NB 128|64|32); every block is reduced into its place of the result.
cudaError_t tn_and_reduce(const GemmTNArgs& ta, const DwReduceArgs& ra, int prec, cudaStream_t st, bool planes, bool wide) {
  if (!wide) {
    cudaError_t e = gemm_tn_auto(ta, prec, st, planes);
    if (e != cudaSuccess) return e;
    return launch_dw_reduce(ra, st);
  }
  const int KA = ta.KA, NB = ta.NB;
  for (int ka0 = 0; ka0 < KA;) {
    const int kac = KA - ka0 >= 256 ? 256 : 128;
    for (int nb0 = 0; nb0 < NB;) {
      const int rem = NB - nb0, nbc = rem >= 128 ? 128 : (rem >= 64 ? 64 : 32);
      GemmTNArgs t = ta;
      t.KA = kac; t.lda = KA; t.a_col0 = ka0; t.NB = t.NB1 = nbc;
      if (t.tok != nullptr) { t.emb_ld = NB; t.emb_col0 = nb0; } else { t.col0 = ta.col0 + nb0; }
      t.colsum = (ta.colsum && nb0 == 0) ? 1 : 0;
      cudaError_t e = launch_gemm_tn(t, prec, st);
      if (e != cudaSuccess) return e;
      DwReduceArgs r = ra;
      r.KA = kac; r.NB = r.NB1 = nbc; r.gi0 = ka0; r.c0 = nb0; r.ldo = NB; r.has_colsum = t.colsum;
      if (!t.colsum) r.out_b1 = r.out_b2 = nullptr;
      e = launch_dw_reduce(r, st);
      if (e != cudaSuccess) return e;
      nb0 += nbc;
    }
    ka0 += kac;
  }
  return cudaSuccess;
}
int tn_launches(int KA, int NB, bool wide) {
  if (!wide) return 2;
  int n = 0;
  for (int ka0 = 0; ka0 < KA; ka0 += (KA - ka0 >= 256 ? 256 : 128))
    for (int nb0 = 0; nb0 < NB; nb0 += (NB - nb0 >= 128 ? 128 : (NB - nb0 >= 64 ? 64 : 32))) n += 2;
  return n;
}
int nt_launches(int NC, bool wide, bool planes = false) {
  if (!wide || planes) return 1;
  int n = 0;
  for (int c0 = 0; c0 < NC; c0 += (NC - c0 >= 256 ? 256 : (NC - c0 >= 128 ? 128 : (NC - c0 >= 64 ? 64 : 32)))) ++n;
  return n;
}

template <typename T>
T* at(void* ws, size_t off) { return reinterpret_cast<T*>(reinterpret_cast<char*>(ws) + off); }

// gemm_l0.cu covers layer 0 of the TMA path when the vocabulary fits its one-hot tile (the decision must be the same in _fwd and _bwd)
bool l0_fused_ok(const Plan& p, bool planes) {
  static const bool disabled = getenv("IB200_NO_L0_FUSED") != nullptr;
  return planes && p.H == 64 && p.V <= 256 && !disabled;
}

// split point of the two-phase rebalancing = (step time of a CTA alone on its SM) / (step time of two co-resident CTAs): the shared
// SMs and the exclusive SMs then finish phase 1 together.  Measured on B200; IB200_SPLIT_FWD / IB200_SPLIT_BWD override.
float split_frac(bool bwd) {
  static const float f[2] = {[] { const char* e = getenv("IB200_SPLIT_FWD"); return e ? (float)atof(e) : 0.66f; }(),
                             [] { const char* e = getenv("IB200_SPLIT_BWD"); return e ? (float)atof(e) : 0.65f; }()};
  return f[bwd ? 1 : 0];
}
// the launchers rebalance in two phases exactly when a HALF launch has between one and two CTAs per SM (lstm_fwd.cu / lstm_bwd.cu)
bool wants_phases(const Plan& p, int ndir) {
  const int full_ctas = ((p.B + 7) / 8) * p.G * ndir, half_ctas = ((p.B + 3) / 4) * p.G * ndir, sms = sm_count();
  return full_ctas <= sms && half_ctas > sms && half_ctas < 2 * sms && p.T >= 128;
}
// One library-owned non-blocking stream + two events per host thread and device.  Used inside ib200_encoder_fwd / _bwd to run SMALL independent
// kernels side by side (layer-0 table + W_ih preparation next to the length kernels; the upper layer's dW reduce next to the dY GEMM).
// All of its work is forked from and joined back into the caller's stream within the call, so the "everything is ordered on the
// stream you pass" contract of the ABI holds.  One per (host thread, device).  IB200_NO_SIDE=1 puts everything back on the caller's stream.
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
// Experiment (OFF by default): the whole upper-layer dW GEMM under the next layer's BPTT kernel.  Measured on B200
// (profiles/r1_overlap_ab.txt): the step gets SLOWER (3.86 -> 3.98 ms) -- the 200 recurrent CTAs already occupy every SM, a co-resident
// GEMM CTA lengthens the dependent chain of its neighbour (BPTT 0.93 -> 1.18 ms) and disturbs the placement the two-phase
// rebalancing relies on.  Small kernels do not have that problem.
bool gemm_overlap_enabled() {
  static const bool on = getenv("IB200_OVERLAP") != nullptr;
  return on;
}
SideStream* side_stream() {
  static const bool disabled = getenv("IB200_NO_SIDE") != nullptr;
  if (disabled) return nullptr;
  // per host thread and device: concurrent callers (e.g. the forward on the main thread and autograd's backward worker) never
  // share a stream or an event, so the fork / join records of one call cannot be overtaken by another
  thread_local SideStream per_dev[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStream& s = per_dev[dev];
  if (s.stream == nullptr) {
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) {
      (void)cudaGetLastError();
      s = SideStream{};
      return nullptr;
    }
  }
  return &s;
}

// Joins the side stream back into the caller's stream on EVERY exit path of a launcher: an early error return between fork and join
// must not leave side-stream kernels writing into a workspace the caller is about to free.
struct SideGuard {
  SideStream* side;
  cudaStream_t st;
  bool armed = false;
  SideGuard(SideStream* s, cudaStream_t main_st) : side(s), st(main_st) {}
  ~SideGuard() {
    if (armed && side != nullptr) {
      (void)cudaEventRecord(side->join, side->stream);
      (void)cudaStreamWaitEvent(st, side->join, 0);
    }
  }
};

PhaseArgs phase_args(void* ws, const Plan& p, bool bwd) {
  PhaseArgs a{};
  a.state = at<float>(ws, p.ph_state);
  a.sm_load = at<int>(ws, p.ph_sched);
  a.resume_list = a.sm_load + 256;
  a.split_frac = split_frac(bwd);
  return a;
}

}  // namespace

extern "C" {

int ib200_version(void) { return IB200_VERSION; }
const char* ib200_last_error(void) { return g_err.c_str(); }

unsigned long long ib200_launch_count(void) { return g_launches.load(); }

int ib200_timing_enable(int on) {
  std::lock_guard<std::mutex> lk(g_timing.mu);
  g_timing.on = on != 0;
  g_timing.used = 0;
  return 0;
}

int ib200_timing_families(void) { return F_COUNT; }
const char* ib200_timing_family_name(int i) { return (i >= 0 && i < F_COUNT) ? kFamilyNames[i] : ""; }

int ib200_timing_read(int n, float* ms_out, int* launches_out) {
  std::lock_guard<std::mutex> lk(g_timing.mu);
  if (!ms_out || !launches_out || n < F_COUNT) return fail(IB200_E_SHAPE, "ib200_timing_read: need arrays of ib200_timing_families() entries");
  for (int i = 0; i < F_COUNT; ++i) { ms_out[i] = 0.f; launches_out[i] = 0; }
  for (size_t i = 0; i < g_timing.used; ++i) {
    cudaError_t e = cudaEventSynchronize(g_timing.pool[2 * i + 1]);
    if (e != cudaSuccess) return cuda_fail(e, "timing sync");
    float ms = 0.f;
    e = cudaEventElapsedTime(&ms, g_timing.pool[2 * i], g_timing.pool[2 * i + 1]);
    if (e != cudaSuccess) return cuda_fail(e, "timing elapsed");
    ms_out[g_timing.fam[i]] += ms;
    launches_out[g_timing.fam[i]] += 1;
  }
  g_timing.used = 0;
  return 0;
}

size_t ib200_workspace_bytes(const ib200_cfg* cfg) {
  if (!cfg_ok(cfg)) return 0;
  return make_plan(cfg).total;
}

int ib200_encoder_fwd(const ib200_cfg* cfg, const void* tokens, const ib200_encoder_params* P, const float* emb_row_scale,
                      const float* whh_l0_mask, int32_t* lengths_out, float* hn_top, void* ws, size_t ws_bytes, void* stream) {
  if (!cfg_ok(cfg)) return fail(IB200_E_UNSUPPORTED, "ib200_encoder_fwd: unsupported cfg (H multiple of 32 in [32, 256], 1<=L<=4, bi_reduce in last/mean/max, token_dtype in IB200_TOK_*, G*B*T < 2^31, 2 <= V <= 28672; for H <= 64 also T <= 11000)");
  if (!tokens || !P || !hn_top || !ws || !P->emb) return fail(IB200_E_NULL, "ib200_encoder_fwd: null pointer");
  const Plan p = make_plan(cfg);
  if (ws_bytes < p.total) return fail(IB200_E_WORKSPACE, "ib200_encoder_fwd: workspace too small");
  if (((uintptr_t)ws & 255) != 0) return fail(IB200_E_ALIGN, "ib200_encoder_fwd: workspace must be 256-byte aligned");
  for (int l = 0; l < p.L; ++l)
    for (int d = 0; d < 2; ++d)
      if (!P->w_ih[l][d] || !P->w_hh[l][d] || !P->b_ih[l][d] || !P->b_hh[l][d]) return fail(IB200_E_NULL, "ib200_encoder_fwd: null LSTM parameter");
  cudaStream_t st = (cudaStream_t)stream;
  const int H = p.H, prec = cfg->precision;
  const bool planes = use_planes(H), cluster = use_cluster(H), wide = H != 32 && H != 64, cluster_tc = use_cluster_tc(H, planes, false);

  // the layer-0 table and the W_ih preparation do not depend on the length kernels: they run next to them on the side stream
  SideStream* side = side_stream();
  SideGuard guard(side, st);
  if (side != nullptr) {
    CK(cudaEventRecord(side->fork, st), "fork record");
    CK(cudaStreamWaitEvent(side->stream, side->fork, 0), "fork wait");
    guard.armed = true;
  }
  // without a row scale every group has the same table: build it once (inference launches fuse up to hundreds of groups)
  const bool table_shared = emb_row_scale == nullptr;
  {
    cudaStream_t main_st = st;
    cudaStream_t st = side != nullptr ? side->stream : main_st;
    TableArgs ta{table_shared ? 1 : p.G, p.V, H, P->emb, emb_row_scale, {P->w_ih[0][0], P->w_ih[0][1]}, {P->b_ih[0][0], P->b_ih[0][1]},
                 {P->b_hh[0][0], P->b_hh[0][1]}, at<float>(ws, p.table)};
    TIMED(F_L0_TABLE, 1, launch_l0_table(ta, st), "l0 table");
    for (int l = 0; l < p.L; ++l)
      for (int d = 0; d < 2; ++d) {
        if (!p.live[l][d]) continue;
        const int K = l == 0 ? H : 2 * H;
        float* w = l > 0 ? at<float>(ws, p.wih_gi[l][d]) : nullptr;
        float* b = l > 0 ? at<float>(ws, p.b_gi[l][d]) : nullptr;
        float* wT = p.train ? at<float>(ws, p.wihT_gi[l][d]) : nullptr;
        if (l == 0 && l0_fused_ok(p, planes)) wT = nullptr;  // W_ih^T of layer 0 only feeds the dX_0 GEMM, which the fused path replaces
        if ((w || wT) && wide && planes)  // W_ih (and its transpose) as bf16 hi|lo plane matrices: TMA operands of gemm_wide.cu
          TIMED(F_PREP, 1, launch_prep_wih_planes(P->w_ih[l][d], P->b_ih[l][d], P->b_hh[l][d], H, K, w, wT, b, prec, st), "prep wih planes");
        else if (w || wT) TIMED(F_PREP, 1, launch_prep_wih(P->w_ih[l][d], P->b_ih[l][d], P->b_hh[l][d], H, K, w, wT, b, st), "prep wih");
      }
    if (side != nullptr) CK(cudaEventRecord(side->join, side->stream), "join record");
  }

  LengthArgs la{p.G, p.B, p.T, p.V, H, tokens, cfg->token_dtype, P->emb, emb_row_scale, at<int>(ws, p.tok32), at<int>(ws, p.lens),
                at<int>(ws, p.row_kind)};
  TIMED(F_LENGTHS, 4, launch_lengths(la, st), "lengths");
  if (lengths_out) CK(cudaMemcpyAsync(lengths_out, at<int>(ws, p.lens), sizeof(int) * 2 * p.G, cudaMemcpyDeviceToDevice, st), "lengths copy");
  if (side != nullptr) {
    CK(cudaStreamWaitEvent(st, side->join, 0), "join wait");
    guard.armed = false;
  }

  if (!p.live[p.L - 1][0]) TIMED(F_FILL, 1, launch_fill_zero(hn_top, (size_t)p.N * H, st), "hn zero");

  for (int l = 0; l < p.L; ++l) {
    const int dir0 = p.live[l][0] ? 0 : 1, ndir = p.live[l][0] ? 2 : 1;
    if (l > 0) {
      for (int d = dir0; d < 2; ++d) {
        GemmNTArgs ga{};
        ga.G = p.G; ga.B = p.B; ga.Tmax = p.T; ga.lens = at<int>(ws, p.lens);
        ga.nsrc = 1; ga.A[0] = at<float>(ws, p.Y[l - 1]); ga.lda = 2 * H; ga.K = 2 * H;
        ga.W[0] = at<float>(ws, p.wih_gi[l][d]); ga.bias = at<float>(ws, p.b_gi[l][d]);
        ga.C = at<float>(ws, p.X[d]); ga.ldc = 4 * H; ga.NC = 4 * H; ga.accumulate = 0;
        ga.plane_bytes = planes ? 2 * H * 2 : 0;
        TIMED(F_GEMM_XPROJ, nt_launches(ga.NC, wide, planes), gemm_nt_auto(ga, prec, st, wide), "input projection gemm");
      }
    }
    LstmFwdArgs fa{};
    fa.G = p.G; fa.B = p.B; fa.Tmax = p.T; fa.V = p.V; fa.dir0 = dir0; fa.ndir = ndir;
    fa.lens = at<int>(ws, p.lens);
    if (l == 0) {
      fa.tok = at<int>(ws, p.tok32);
      fa.table = at<float>(ws, p.table);
      fa.table_shared = table_shared ? 1 : 0;
    } else {
      fa.xproj[0] = at<float>(ws, p.X[0]);
      fa.xproj[1] = at<float>(ws, p.X[1]);
    }
    fa.whh[0] = P->w_hh[l][0]; fa.whh[1] = P->w_hh[l][1];
    fa.whh_mask = l == 0 ? whh_l0_mask : nullptr;
    const bool need_y = p.train || l < p.L - 1;
    fa.y = need_y ? at<float>(ws, p.Y[l]) : nullptr;
    fa.y_stride = 2 * H;
    fa.planes = planes ? 1 : 0;
    for (int d = 0; d < 2; ++d) {
      fa.gates[d] = (p.train && p.live[l][d]) ? at<float>(ws, p.gates[l][d]) : nullptr;
      fa.cstate[d] = (p.train && p.live[l][d]) ? at<float>(ws, p.cst[l][d]) : nullptr;
    }
    fa.hn = l == p.L - 1 ? hn_top : nullptr;
    fa.dbg = dbg_flags();
    if (!cluster && wants_phases(p, ndir)) {  // scheduling state of the two-phase rebalancing
      fa.ph = phase_args(ws, p, false);
      TIMED(F_FILL, 1, launch_fill_zero(reinterpret_cast<float*>(fa.ph.sm_load), 257, st), "phase counters");
    }
    TIMED(l == 0 ? F_LSTM_FWD_L0 : F_LSTM_FWD_UP, 1,
          cluster_tc ? launch_lstm_fwd_cluster_tc(fa, H, prec, st)
                     : (cluster ? launch_lstm_fwd_cluster(fa, H, prec, st) : launch_lstm_fwd(fa, H, prec, st)),
          "lstm fwd");
  }
  return 0;
}

int ib200_encoder_status(const ib200_cfg* cfg, const void* ws, size_t ws_bytes, int32_t* status_out, void* stream) {
  if (!cfg_ok(cfg)) return fail(IB200_E_UNSUPPORTED, "ib200_encoder_status: unsupported cfg");
  if (!ws || !status_out) return fail(IB200_E_NULL, "ib200_encoder_status: null pointer");
  const Plan p = make_plan(cfg);
  if (ws_bytes < p.total) return fail(IB200_E_WORKSPACE, "ib200_encoder_status: workspace too small");
  CK(cudaMemcpyAsync(status_out, reinterpret_cast<const char*>(ws) + p.lens, sizeof(int) * 3 * p.G, cudaMemcpyDeviceToDevice,
                     (cudaStream_t)stream), "status copy");
  return 0;
}

int ib200_encoder_bwd(const ib200_cfg* cfg, const ib200_encoder_params* P, const float* emb_row_scale, const float* whh_l0_mask,
                      const float* d_hn_top, const ib200_encoder_grads* Gr, void* ws, size_t ws_bytes, void* stream) {
  if (!cfg_ok(cfg)) return fail(IB200_E_UNSUPPORTED, "ib200_encoder_bwd: cfg must be the training cfg used for _fwd");
  return ib200_encoder_bwd_layers(cfg, P, emb_row_scale, whh_l0_mask, d_hn_top, Gr, ws, ws_bytes, cfg->L - 1, 0, stream);
}

int ib200_encoder_bwd_layers(const ib200_cfg* cfg, const ib200_encoder_params* P, const float* emb_row_scale, const float* whh_l0_mask,
                             const float* d_hn_top, const ib200_encoder_grads* Gr, void* ws, size_t ws_bytes, int32_t layer_hi,
                             int32_t layer_lo, void* stream) {
  if (!cfg_ok(cfg) || !cfg->training) return fail(IB200_E_UNSUPPORTED, "ib200_encoder_bwd: cfg must be the training cfg used for _fwd");
  if (!P || !d_hn_top || !Gr || !ws) return fail(IB200_E_NULL, "ib200_encoder_bwd: null pointer");
  if (layer_lo < 0 || layer_hi >= cfg->L || layer_lo > layer_hi) return fail(IB200_E_SHAPE, "ib200_encoder_bwd_layers: need 0 <= layer_lo <= layer_hi < L");
  const Plan p = make_plan(cfg);
  if (ws_bytes < p.total) return fail(IB200_E_WORKSPACE, "ib200_encoder_bwd: workspace too small");
  for (int l = layer_lo; l <= layer_hi; ++l)
    for (int d = 0; d < 2; ++d)
      if (!Gr->w_ih[l][d] || !Gr->w_hh[l][d] || !Gr->b_ih[l][d] || !Gr->b_hh[l][d]) return fail(IB200_E_NULL, "ib200_encoder_bwd: null gradient tensor");
  if (layer_lo == 0 && !Gr->emb) return fail(IB200_E_NULL, "ib200_encoder_bwd: null embedding gradient");
  cudaStream_t st = (cudaStream_t)stream;
  const int H = p.H, prec = cfg->precision;
  const bool planes = use_planes(H), cluster = use_cluster(H), wide = H != 32 && H != 64, cluster_tc = use_cluster_tc(H, planes, true);
  const int* lens = at<int>(ws, p.lens);
  float* dY = at<float>(ws, p.bwd_scratch);
  float* dX0 = dY + p.R * 2 * H;
  float* partial = at<float>(ws, p.partial);

  SideStream* side = side_stream();
  SideGuard guard(side, st);  // armed from the first fork on: any exit path joins the side stream (a redundant join is harmless)
  const bool overlap_gemm = side != nullptr && gemm_overlap_enabled();
  bool pending_join = false;
  for (int l = layer_hi; l >= layer_lo; --l) {
    const int dir0 = p.live[l][0] ? 0 : 1, ndir = p.live[l][0] ? 2 : 1;
    LstmBwdArgs ba{};
    ba.G = p.G; ba.B = p.B; ba.Tmax = p.T; ba.dir0 = dir0; ba.ndir = ndir; ba.lens = lens;
    ba.whh[0] = P->w_hh[l][0]; ba.whh[1] = P->w_hh[l][1];
    ba.whh_mask = l == 0 ? whh_l0_mask : nullptr;
    for (int d = 0; d < 2; ++d) {
      ba.gates[d] = p.live[l][d] ? at<float>(ws, p.gates[l][d]) : nullptr;
      ba.cstate[d] = p.live[l][d] ? at<float>(ws, p.cst[l][d]) : nullptr;
    }
    ba.dy = l == p.L - 1 ? nullptr : dY;
    ba.dy_stride = 2 * H;
    ba.dhn = l == p.L - 1 ? d_hn_top : nullptr;
    ba.dbg = dbg_flags();
    ba.planes = planes ? 1 : 0;
    ba.bias_partial = planes ? at<float>(ws, p.bias_partial[l & 1]) : nullptr;
    if (!cluster && wants_phases(p, ndir)) {
      ba.ph = phase_args(ws, p, true);
      TIMED(F_FILL, 1, launch_fill_zero(reinterpret_cast<float*>(ba.ph.sm_load), 257, st), "phase counters");
    }
    TIMED(l == 0 ? F_LSTM_BWD_L0 : F_LSTM_BWD_UP, 1,
          cluster_tc ? launch_lstm_bwd_cluster_tc(ba, H, prec, st)
                     : (cluster ? launch_lstm_bwd_cluster(ba, H, prec, st) : launch_lstm_bwd(ba, H, prec, st)),
          "lstm bwd");
    // bias partial rows per direction left by the BPTT kernel
    const int bwd_ctas = !planes ? 0 : (cluster_tc ? lstm_bwd_cluster_tc_cta_count(ba, H, prec)
                                                   : (cluster ? lstm_bwd_cluster_cta_count(ba, H, prec) : lstm_bwd_cta_count(ba, prec)));

    // layer 0 of the TMA path: dW_hh, dW_ih, the bias gradients and the embedding gradient from ONE pass over the dgates
    // (gemm_l0.cu: the token-indexed sums S = dA^T onehot(tok) replace the gathered dW GEMM, the dX_0 GEMM and the atomic scatter)
    bool l0_done = false;
    if (l == 0 && pending_join) {  // the side stream's GEMMs share the `partial` scratch with everything below
      CK(cudaStreamWaitEvent(st, side->join, 0), "join wait");
      pending_join = false;
    }
    if (l == 0 && l0_fused_ok(p, planes)) {
      L0GradArgs la{};
      la.G = p.G; la.B = p.B; la.Tmax = p.T; la.V = p.V; la.H = H; la.dir0 = dir0; la.ndir = ndir;
      la.lens = lens; la.tok = at<int>(ws, p.tok32);
      for (int d = 0; d < 2; ++d) {
        la.dA[d] = p.live[0][d] ? at<float>(ws, p.gates[0][d]) : nullptr;
        la.w_ih[d] = P->w_ih[0][d];
        la.d_wih[d] = Gr->w_ih[0][d]; la.d_whh[d] = Gr->w_hh[0][d]; la.d_bih[d] = Gr->b_ih[0][d]; la.d_bhh[d] = Gr->b_hh[0][d];
      }
      la.Y0 = at<float>(ws, p.Y[0]);
      la.emb = P->emb; la.emb_row_scale = emb_row_scale; la.whh_mask = whh_l0_mask;
      la.bias_partial = at<float>(ws, p.bias_partial[0]); la.bias_count = bwd_ctas;
      la.partial = partial; la.R = at<float>(ws, p.l0_scratch); la.d_emb = Gr->emb;
      TimedScope ts(F_L0_GRADS, 3, st);
      // no fallback: the forward skipped the W_ih^T preparation of layer 0 because this path was chosen (same l0_fused_ok decision)
      const cudaError_t e = launch_l0_grads(la, prec, st);
      if (e != cudaSuccess) return cuda_fail(e, "layer-0 gradient gemm");
      l0_done = true;
    }

    // weight gradients of this layer (read dA = gates buffers, Y_{l-1} / embeddings, Y_l); `st` = the stream they are issued on
    bool x0_ready = false;
    auto weight_grads = [&](cudaStream_t st) -> int {
    for (int d = 0; d < 2; ++d) {
      const int K = l == 0 ? H : 2 * H;
      if (!p.live[l][d]) {  // dead chain: exact zeros
        const size_t n_ih = (size_t)4 * H * K, n_hh = (size_t)4 * H * H, n_b = (size_t)4 * H;
        if (Gr->w_hh[l][d] == Gr->w_ih[l][d] + n_ih && Gr->b_ih[l][d] == Gr->w_hh[l][d] + n_hh && Gr->b_hh[l][d] == Gr->b_ih[l][d] + n_b) {
          // the four tensors are consecutive slices of one flat gradient buffer (the layout ops.py hands in): one fill
          TIMED(F_FILL, 1, launch_fill_zero(Gr->w_ih[l][d], n_ih + n_hh + 2 * n_b, st), "zero");
        } else {
          TIMED(F_FILL, 1, launch_fill_zero(Gr->w_ih[l][d], n_ih, st), "zero");
          TIMED(F_FILL, 1, launch_fill_zero(Gr->w_hh[l][d], n_hh, st), "zero");
          TIMED(F_FILL, 1, launch_fill_zero(Gr->b_ih[l][d], n_b, st), "zero");
          TIMED(F_FILL, 1, launch_fill_zero(Gr->b_hh[l][d], n_b, st), "zero");
        }
        continue;
      }
      if (l0_done) continue;
      GemmTNArgs ta{};
      ta.G = p.G; ta.B = p.B; ta.Tmax = p.T; ta.lens = lens;
      ta.A = at<float>(ws, p.gates[l][d]); ta.KA = 4 * H;
      ta.partial = partial; ta.ctas_per_group = p.ctas_per_group;
      const float* mask_hh = (l == 0 && d == 0) ? whh_l0_mask : nullptr;
      // first B source: the layer input (dW_ih)
      if (l == 0) {
        ta.tok = at<int>(ws, p.tok32); ta.emb = P->emb; ta.emb_row_scale = emb_row_scale; ta.V = p.V; ta.NB = H;
      } else {
        ta.Bsrc = at<float>(ws, p.Y[l - 1]); ta.ldb = 2 * H; ta.col0 = 0; ta.shift = 0; ta.NB = 2 * H;
      }
      ta.NB1 = ta.NB;
      ta.colsum = 0;
      DwReduceArgs ra{};
      ra.G = p.G; ra.ctas_per_group = p.ctas_per_group; ra.KA = 4 * H; ra.NB = ta.NB; ra.H = H; ra.partial = partial;
      ra.out = Gr->w_ih[l][d]; ra.NB1 = ta.NB;
      if (planes) {
        // TMA path: ONE pass over the dgates produces [dW_ih | dW_hh]: second B source = h of the previous scan position = Y_l
        // shifted by one step; the bias gradients come from the column sums the BPTT kernel left behind
        ta.Bsrc2 = at<float>(ws, p.Y[l]); ta.ldb2 = 2 * H; ta.col02 = d * H; ta.shift2 = d == 0 ? -1 : +1;
        ta.NB = ta.NB1 + H;
        ra.NB = ta.NB; ra.out2 = Gr->w_hh[l][d]; ra.mask = mask_hh;
        ra.out_b1 = Gr->b_ih[l][d]; ra.out_b2 = Gr->b_hh[l][d];
        ra.cs_ptr = at<float>(ws, p.bias_partial[l & 1]) + (size_t)(d - dir0) * bwd_ctas * 4 * H;
        ra.cs_count = bwd_ctas;
        if (pending_join && st != side->stream) {  // a reduce of this call may still be reading `partial` on the side stream
          CK(cudaStreamWaitEvent(st, side->join, 0), "join wait");
          pending_join = false;
        }
        if (wide) {
          if (l == 0) {  // the layer-0 input rows scale*emb[tok] as a dense plane operand (one small gather instead of a gathering GEMM loader)
            float* x0p = at<float>(ws, p.x0p);
            if (!x0_ready) {
              TIMED(F_EMB_GRAD, 1, launch_gather_x0_planes(p.G, p.B, p.T, p.V, H, lens, at<int>(ws, p.tok32), P->emb, emb_row_scale, x0p, prec, st),
                    "layer-0 input planes");
              x0_ready = true;
            }
            ta.tok = nullptr; ta.emb = nullptr; ta.emb_row_scale = nullptr;
            ta.Bsrc = x0p; ta.ldb = H; ta.col0 = 0; ta.shift = 0;
          }
          ta.ctas_per_group = ra.ctas_per_group = gemm_tn_wide_splits(ta.KA, ta.NB, H, p.G);
          TIMED(F_GEMM_DW, 1, launch_gemm_tn_wide(ta, prec, st), "dW_ih|dW_hh gemm (wide)");
        } else {
          TIMED(F_GEMM_DW, 1, gemm_tn_auto(ta, prec, st, true), "dW_ih|dW_hh gemm");
        }
        if (l > 0 && side != nullptr && st != side->stream) {
          // the reduce (a small grid) runs on the side stream next to this layer's dY GEMM; joined before `partial` is written again
          CK(cudaEventRecord(side->fork, st), "fork record");
          CK(cudaStreamWaitEvent(side->stream, side->fork, 0), "fork wait");
          guard.armed = true;
          {
            cudaStream_t st = side->stream;
            TIMED(F_DW_REDUCE, 1, launch_dw_reduce(ra, st), "dW reduce");
          }
          CK(cudaEventRecord(side->join, side->stream), "join record");
          pending_join = true;
        } else {
          TIMED(F_DW_REDUCE, 1, launch_dw_reduce(ra, st), "dW reduce");
        }
        continue;
      }
      TIMED(F_GEMM_DW, tn_launches(ta.KA, ta.NB, wide), tn_and_reduce(ta, ra, prec, st, planes, wide), "dW_ih gemm + reduce");
      // dW_hh (+ bias gradients): B operand = h of the previous scan position = Y_l shifted by one step
      ta.tok = nullptr; ta.emb = nullptr; ta.emb_row_scale = nullptr;
      ta.Bsrc = at<float>(ws, p.Y[l]); ta.ldb = 2 * H; ta.col0 = d * H; ta.shift = d == 0 ? -1 : +1; ta.NB = H; ta.NB1 = H;
      ta.colsum = 1;
      DwReduceArgs rb{};
      rb.G = p.G; rb.ctas_per_group = p.ctas_per_group; rb.KA = 4 * H; rb.NB = H; rb.H = H; rb.partial = partial;
      rb.has_colsum = 1; rb.mask = mask_hh; rb.out = Gr->w_hh[l][d]; rb.NB1 = H;
      rb.out_b1 = Gr->b_ih[l][d]; rb.out_b2 = Gr->b_hh[l][d];
      TIMED(F_GEMM_DW, tn_launches(ta.KA, ta.NB, wide), tn_and_reduce(ta, rb, prec, st, planes, wide), "dW_hh gemm + reduce");
    }
    return 0;
    };

    // input gradient of this layer
    auto input_grads = [&](cudaStream_t st) -> int {
    if (l0_done) return 0;
    GemmNTArgs ga{};
    ga.G = p.G; ga.B = p.B; ga.Tmax = p.T; ga.lens = lens;
    ga.nsrc = 0;
    for (int d = 0; d < 2; ++d)
      if (p.live[l][d]) {
        ga.A[ga.nsrc] = at<float>(ws, p.gates[l][d]);
        ga.W[ga.nsrc] = at<float>(ws, p.wihT_gi[l][d]);
        ++ga.nsrc;
      }
    ga.lda = 4 * H; ga.K = 4 * H; ga.bias = nullptr; ga.accumulate = 0;
    ga.plane_bytes = planes ? 4 * H * 2 : 0;
    if (l > 0) {
      ga.C = dY; ga.ldc = 2 * H; ga.NC = 2 * H;
      TIMED(F_GEMM_DGRAD, nt_launches(ga.NC, wide, planes), gemm_nt_auto(ga, prec, st, wide), "dY gemm");
    } else {
      ga.C = dX0; ga.ldc = H; ga.NC = H;
      TIMED(F_GEMM_DGRAD, nt_launches(ga.NC, wide, planes), gemm_nt_auto(ga, prec, st, wide), "dX0 gemm");
      EmbGradArgs ea{p.G, p.B, p.T, p.V, H, lens, at<int>(ws, p.tok32), dX0, emb_row_scale, Gr->emb};
      TIMED(F_EMB_GRAD, 2, launch_emb_grad(ea, st), "embedding grad");
    }
    return 0;
    };

    if (l > 0 && planes && overlap_gemm) {
      // (experiment) Upper layers of the TMA path: the next BPTT launch only needs dY, so dY goes first and this layer's weight-gradient GEMM
      // + reduce (HBM-bound, needed by nobody until the end) run on the library's side stream UNDERNEATH the next layer's
      // latency-bound recurrent kernel.  Joined before anything else touches the `partial` scratch (top of the l == 0 iteration).
      if (int rc = input_grads(st)) return rc;
      CK(cudaEventRecord(side->fork, st), "fork record");
      CK(cudaStreamWaitEvent(side->stream, side->fork, 0), "fork wait");
      guard.armed = true;
      if (int rc = weight_grads(side->stream)) return rc;
      CK(cudaEventRecord(side->join, side->stream), "join record");
      pending_join = true;
    } else {
      if (int rc = weight_grads(st)) return rc;
      if (int rc = input_grads(st)) return rc;
    }
  }
  if (pending_join) CK(cudaStreamWaitEvent(st, side->join, 0), "join wait");
  guard.armed = false;
  return 0;
}

int ib200_pool_fc_fwd(int32_t N, int32_t H, int32_t bi_reduce, const float* hn_top, const float* fc_w, const float* fc_b, float* z,
                      float* pooled_out, uint8_t* argmax_out, void* stream) {
  if (!hn_top || !fc_w || !fc_b || !z) return fail(IB200_E_NULL, "ib200_pool_fc_fwd: null pointer");
  if (N < 1 || H < 1 || H > 1024 || bi_reduce < 0 || bi_reduce > 2) return fail(IB200_E_SHAPE, "ib200_pool_fc_fwd: bad shape / bi_reduce");
  cudaStream_t st = (cudaStream_t)stream;
  TIMED(F_POOL_FC, 1, launch_pool_fc_fwd(N, H, bi_reduce, hn_top, fc_w, fc_b, z, pooled_out, argmax_out, st), "pool_fc fwd");
  return 0;
}

int ib200_pool_fc_bwd(int32_t N, int32_t H, int32_t bi_reduce, const float* dz, const float* pooled, const uint8_t* argmax,
                      const float* fc_w, float* d_hn_top, float* d_fc_w, float* d_fc_b, void* stream) {
  if (!dz || !pooled || !fc_w || !d_hn_top || !d_fc_w || !d_fc_b) return fail(IB200_E_NULL, "ib200_pool_fc_bwd: null pointer");
  if (bi_reduce == IB200_REDUCE_MAX && !argmax) return fail(IB200_E_NULL, "ib200_pool_fc_bwd: max needs argmax");
  if (N < 1 || H < 1 || H > 1024 || bi_reduce < 0 || bi_reduce > 2) return fail(IB200_E_SHAPE, "ib200_pool_fc_bwd: bad shape / bi_reduce");
  cudaStream_t st = (cudaStream_t)stream;
  TIMED(F_POOL_FC, 2, launch_pool_fc_bwd(N, H, bi_reduce, dz, pooled, argmax, fc_w, d_hn_top, d_fc_w, d_fc_b, st), "pool_fc bwd");
  return 0;
}

int ib200_loss_head_fwd(int32_t B, int32_t H, float beta, const float* z, const int64_t* y, const ib200_head_params* hp,
                        const ib200_head_masks* hm, float* losses_out, float* y_hat_out, void* stream) {
  if (!z || !y || !hp || !losses_out || !y_hat_out || !hp->fc1_w || !hp->fc1_b || !hp->fc2_w || !hp->fc2_b)
    return fail(IB200_E_NULL, "ib200_loss_head_fwd: null pointer");
  if (B < 1 || !head_supports(H) || !(beta > 0.f)) return fail(IB200_E_SHAPE, "ib200_loss_head_fwd: bad shape (H must be a multiple of 32 in [32, 256])");
  ib200_head_masks none{};
  cudaStream_t st = (cudaStream_t)stream;
  TIMED(F_LOSS_HEAD, 1, launch_loss_head_fwd(B, H, beta, z, (const long long*)y, *hp, hm ? *hm : none, losses_out, y_hat_out, st), "loss_head fwd");
  return 0;
}

int ib200_loss_head_bwd(int32_t B, int32_t H, float beta, const float* z, const int64_t* y, const ib200_head_params* hp,
                        const ib200_head_masks* hm, const float* d_loss, const float* d_y_hat, float* dz_out, const ib200_head_grads* hg,
                        void* stream) {
  if (!z || !y || !hp || !d_loss || !dz_out || !hg || !hg->fc1_w || !hg->fc1_b || !hg->fc2_w || !hg->fc2_b)
    return fail(IB200_E_NULL, "ib200_loss_head_bwd: null pointer");
  if (B < 1 || !head_supports(H) || !(beta > 0.f)) return fail(IB200_E_SHAPE, "ib200_loss_head_bwd: bad shape (H must be a multiple of 32 in [32, 256])");
  ib200_head_masks none{};
  cudaStream_t st = (cudaStream_t)stream;
  TIMED(F_LOSS_HEAD, 1, launch_loss_head_bwd(B, H, beta, z, (const long long*)y, *hp, hm ? *hm : none, d_loss, d_y_hat, dz_out, *hg, st), "loss_head bwd");
  return 0;
}

int ib200_pair_score(int32_t M, int32_t H, const float* z, const int32_t* idx_a, const int32_t* idx_b, int64_t P,
                     const ib200_head_params* hp, float* prob_out, void* stream) {
  if (!z || !hp || !prob_out || !hp->fc1_w || !hp->fc1_b || !hp->fc2_w || !hp->fc2_b) return fail(IB200_E_NULL, "ib200_pair_score: null pointer");
  if ((idx_a == nullptr) != (idx_b == nullptr)) return fail(IB200_E_NULL, "ib200_pair_score: idx_a and idx_b must both be given or both be null");
  if (M < 1 || !head_supports(H) || P < 0) return fail(IB200_E_SHAPE, "ib200_pair_score: bad shape (H must be a multiple of 32 in [32, 256])");
  if (!idx_a && P != (int64_t)M * (M + 1) / 2) return fail(IB200_E_SHAPE, "ib200_pair_score: P must be M(M+1)/2 for the implicit upper triangle");
  cudaStream_t st = (cudaStream_t)stream;
  TIMED(F_PAIR_SCORE, 1, launch_pair_score(M, H, z, idx_a, idx_b, (long long)P, 0, *hp, prob_out, st), "pair_score");
  return 0;
}

int ib200_pair_score_range(int32_t M, int32_t H, const float* z, int64_t p_begin, int64_t p_count, const ib200_head_params* hp,
                           float* prob_out, void* stream) {
  if (M < 1 || !head_supports(H)) return fail(IB200_E_SHAPE, "ib200_pair_score_range: bad shape (H must be a multiple of 32 in [32, 256])");
  const int64_t total = (int64_t)M * (M + 1) / 2;
  if (p_begin < 0 || p_count < 0 || p_begin + p_count > total) return fail(IB200_E_SHAPE, "ib200_pair_score_range: range outside the M(M+1)/2 triangle");
  if (p_count == 0) return 0;
  if (!z || !hp || !prob_out || !hp->fc1_w || !hp->fc1_b || !hp->fc2_w || !hp->fc2_b) return fail(IB200_E_NULL, "ib200_pair_score_range: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  TIMED(F_PAIR_SCORE, 1, launch_pair_score(M, H, z, nullptr, nullptr, (long long)p_count, (long long)p_begin, *hp, prob_out, st), "pair_score_range");
  return 0;
}

int ib200_sequence_lengths(int32_t M, int32_t T, int32_t V, int32_t H, const void* tokens, int32_t token_dtype, const float* emb,
                           int32_t* t1_out, int32_t* teff_out, int32_t* scratch, void* stream) {
  if (M < 0 || T < 1 || V < 2 || V > kMaxVocab || H < 1) return fail(IB200_E_SHAPE, "ib200_sequence_lengths: bad shape (2 <= V <= 28672)");
  if (token_dtype != IB200_TOK_I64 && token_dtype != IB200_TOK_I32 && token_dtype != IB200_TOK_I16 && token_dtype != IB200_TOK_U8)
    return fail(IB200_E_SHAPE, "ib200_sequence_lengths: unknown token_dtype");
  if (M == 0) return 0;
  if (!tokens || !emb || !t1_out || !teff_out || !scratch) return fail(IB200_E_NULL, "ib200_sequence_lengths: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  TIMED(F_LENGTHS, 2, launch_seq_lengths(M, T, V, H, tokens, token_dtype, emb, scratch, t1_out, teff_out, st), "sequence lengths");
  return 0;
}

int ib200_draw_masks(int32_t n_specs, const ib200_mask_spec* specs, uint64_t seed, uint64_t offset, uint64_t* counters_used, void* stream) {
  if (n_specs < 0) return fail(IB200_E_SHAPE, "ib200_draw_masks: negative mask count");
  if (counters_used) *counters_used = 0;
  if (n_specs == 0) return 0;
  if (!specs) return fail(IB200_E_NULL, "ib200_draw_masks: null pointer");
  for (int i = 0; i < n_specs; ++i) {
    if (specs[i].numel < 0 || !(specs[i].keep_prob > 0.f && specs[i].keep_prob <= 1.f)) return fail(IB200_E_SHAPE, "ib200_draw_masks: keep_prob must be in (0, 1]");
    if (specs[i].numel > 0 && !specs[i].out) return fail(IB200_E_NULL, "ib200_draw_masks: null output");
  }
  cudaStream_t st = (cudaStream_t)stream;
  int launches = 0;
  unsigned long long used = 0;
  {
    TimedScope ts__(F_MASKS, 0, st);
    const cudaError_t e = launch_draw_masks(n_specs, specs, seed, offset, &used, st, &launches);
    g_launches.fetch_add((unsigned long long)launches);
    if (e != cudaSuccess) return cuda_fail(e, "draw_masks");
  }
  if (counters_used) *counters_used = used;
  return 0;
}

int ib200_batch_metrics(int32_t B, const float* y_hat, const int64_t* y, float threshold, float* metrics_out, int32_t* confusion_out,
                        void* stream) {
  if (B < 1 || B > 1024) return fail(IB200_E_SHAPE, "ib200_batch_metrics: batch must be in [1, 1024]");
  if (!y_hat || !y || !metrics_out) return fail(IB200_E_NULL, "ib200_batch_metrics: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  TIMED(F_METRICS, 1, launch_batch_metrics(B, y_hat, (const long long*)y, threshold, metrics_out, confusion_out, st), "batch_metrics");
  return 0;
}

int ib200_p2p_allreduce_mean(int32_t world, int32_t rank, void* const* stage_ptrs, void* const* flag_ptrs, size_t stage_floats,
                             float* data, size_t n, uint32_t epoch, void* stream) {
  if (world < 1 || world > kP2PMaxWorld || rank < 0 || rank >= world) return fail(IB200_E_SHAPE, "ib200_p2p_allreduce_mean: need 1 <= world <= 8, 0 <= rank < world");
  if (!stage_ptrs || !flag_ptrs || (!data && n > 0)) return fail(IB200_E_NULL, "ib200_p2p_allreduce_mean: null pointer");
  if (n > stage_floats) return fail(IB200_E_SHAPE, "ib200_p2p_allreduce_mean: bucket larger than a staging half");
  if (epoch == 0) return fail(IB200_E_SHAPE, "ib200_p2p_allreduce_mean: epochs count from 1 (the flags start at 0)");
  cudaStream_t st = (cudaStream_t)stream;
  P2PArgs a{};
  a.world = world; a.rank = rank; a.data = data; a.n = n; a.epoch = epoch;
  const size_t half = (size_t)(epoch & 1u) * stage_floats;
  for (int r = 0; r < world; ++r) {
    if (!stage_ptrs[r] || !flag_ptrs[r]) return fail(IB200_E_NULL, "ib200_p2p_allreduce_mean: null peer pointer");
    if ((reinterpret_cast<uintptr_t>(stage_ptrs[r]) & 15) != 0 || (stage_floats & 3) != 0)
      return fail(IB200_E_ALIGN, "ib200_p2p_allreduce_mean: staging regions must be 16-byte aligned, stage_floats a multiple of 4");
    a.stage[r] = reinterpret_cast<const float*>(stage_ptrs[r]) + half;
    a.flags[r] = reinterpret_cast<uint32_t*>(flag_ptrs[r]);
  }
  TIMED(F_ALLREDUCE, 3, launch_p2p_allreduce_mean(a, reinterpret_cast<float*>(stage_ptrs[rank]) + half, st), "p2p allreduce");
  return 0;
}

int ib200_p2p_alloc(size_t bytes, void** ptr_out, unsigned char* handle_out) {
  if (!ptr_out || !handle_out || bytes == 0) return fail(IB200_E_NULL, "ib200_p2p_alloc: null pointer / zero size");
  static_assert(sizeof(cudaIpcMemHandle_t) == IB200_P2P_HANDLE_BYTES, "CUDA IPC handle size");
  void* p = nullptr;
  CK(cudaMalloc(&p, bytes), "p2p alloc");
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    (void)cudaFree(p);
    return cuda_fail(e, "p2p alloc (memset / ipc handle)");
  }
  memcpy(handle_out, &h, sizeof(h));
  *ptr_out = p;
  return 0;
}
int ib200_p2p_open(const unsigned char* handle, void** ptr_out) {
  if (!handle || !ptr_out) return fail(IB200_E_NULL, "ib200_p2p_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  CK(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess), "p2p open");
  return 0;
}
int ib200_p2p_close(void* ptr) {
  if (ptr) CK(cudaIpcCloseMemHandle(ptr), "p2p close");
  return 0;
}
int ib200_p2p_free(void* ptr) {
  if (ptr) CK(cudaFree(ptr), "p2p free");
  return 0;
}

int ib200_adamw_step(int32_t n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                     float* const* exp_avg_sq, const int64_t* numel, const ib200_adamw_hyper* h, void* stream) {
  if (n_tensors < 0) return fail(IB200_E_SHAPE, "ib200_adamw_step: negative tensor count");
  if (n_tensors == 0) return 0;
  if (!params || !grads || !exp_avg || !exp_avg_sq || !numel || !h) return fail(IB200_E_NULL, "ib200_adamw_step: null pointer");
  if (h->step < 1) return fail(IB200_E_SHAPE, "ib200_adamw_step: step counts from 1");
  if (!(h->beta1 >= 0. && h->beta1 < 1. && h->beta2 >= 0. && h->beta2 < 1. && h->eps >= 0. && h->lr >= 0. && h->weight_decay >= 0.))
    return fail(IB200_E_SHAPE, "ib200_adamw_step: hyper-parameters out of range (torch.optim.AdamW raises on these too)");
  for (int k = 0; k < n_tensors; ++k)
    if (numel[k] > 0 && grads[k] && (!params[k] || !exp_avg[k] || !exp_avg_sq[k])) return fail(IB200_E_NULL, "ib200_adamw_step: null tensor pointer");
  // every scalar is derived in double on the host and rounded once, as torch does with Python floats (torch/optim/adamw.py)
  const double bc1 = 1.0 - std::pow(h->beta1, (double)h->step), bc2 = 1.0 - std::pow(h->beta2, (double)h->step);
  AdamScalars s;
  s.grad_scale = (float)(h->maximize ? -h->grad_scale : h->grad_scale);
  s.decay = (float)(1.0 - h->lr * h->weight_decay);
  s.one_minus_b1 = (float)(1.0 - h->beta1);
  s.b2 = (float)h->beta2;
  s.one_minus_b2 = (float)(1.0 - h->beta2);
  s.bc2_sqrt = (float)std::sqrt(bc2);
  s.eps = (float)h->eps;
  s.step_size = (float)(h->lr / bc1);
  cudaStream_t st = (cudaStream_t)stream;
  int launches = 0;
  {
    TimedScope ts__(F_ADAMW, 0, st);
    const cudaError_t e = launch_adamw(n_tensors, params, grads, exp_avg, exp_avg_sq, (const long long*)numel, s, st, &launches);
    g_launches.fetch_add((unsigned long long)launches);
    if (e != cudaSuccess) return cuda_fail(e, "adamw");
  }
  return 0;
}

size_t ib200_ranger21_scratch_bytes(int32_t n_tensors, const ib200_ranger21_tensor* tensors) {
  if (n_tensors <= 0 || !tensors) return 0;
  long long ctas = 0, rows = 0;
  for (int k = 0; k < n_tensors; ++k) {
    if (tensors[k].rows < 1 || tensors[k].cols < 1) return 0;
    ctas += ranger21_elem_ctas((long long)tensors[k].rows * (long long)tensors[k].cols);
    rows += tensors[k].rows;
  }
  return (size_t)(8 * (3 + ctas + 2 * rows) + 4 * rows);
}

int ib200_ranger21_step(int32_t n_tensors, const ib200_ranger21_tensor* tensors, const ib200_ranger21_hyper* h, double* scratch,
                        void* stream) {
  if (n_tensors < 0 || n_tensors > 4096) return fail(IB200_E_SHAPE, "ib200_ranger21_step: tensor count out of range");
  if (n_tensors == 0) return 0;
  if (!tensors || !h || !scratch) return fail(IB200_E_NULL, "ib200_ranger21_step: null pointer");
  if (!(h->beta1 >= 0. && h->beta1 < 1. && h->beta2 >= 0. && h->beta2 < 1. && h->eps >= 0. && h->weight_decay >= 0.))
    return fail(IB200_E_SHAPE, "ib200_ranger21_step: hyper-parameters out of range");
  std::vector<R21Tensor> tb((size_t)n_tensors);
  double param_size = 0.0;
  // scratch: 3 doubles (variance_normalized, its inverse, arrival counter) | one double per elementwise CTA | 2 doubles per row | 1 float per row
  long long total_ctas = 0, total_rows = 0;
  for (int k = 0; k < n_tensors; ++k) {
    const ib200_ranger21_tensor& t = tensors[k];
    if (!t.param || !t.grad || !t.grad_ma || !t.neg_grad_ma || !t.variance_ma || (h->lookahead_merge && !t.lookahead))
      return fail(IB200_E_NULL, "ib200_ranger21_step: null tensor pointer");
    if (t.rows < 1 || t.cols < 1 || t.rows > INT32_MAX || t.cols > INT32_MAX || t.rows * t.cols > INT32_MAX)
      return fail(IB200_E_SHAPE, "ib200_ranger21_step: empty or oversized tensor");
    if (t.step < 1) return fail(IB200_E_SHAPE, "ib200_ranger21_step: step counts from 1");
    if (!(t.lr >= 0.)) return fail(IB200_E_SHAPE, "ib200_ranger21_step: negative learning rate");
    total_ctas += ranger21_elem_ctas(t.rows * t.cols);
    total_rows += t.rows;
  }
  double* rowsum = scratch + 3 + total_ctas;
  float* pnorm = reinterpret_cast<float*>(rowsum + 2 * total_rows);
  for (int k = 0; k < n_tensors; ++k) {
    const ib200_ranger21_tensor& t = tensors[k];
    // every scalar is derived in double on the host and rounded once, as the package derives them from Python floats
    const double bc1 = 1.0 - std::pow(h->beta1, (double)t.step), bc2 = 1.0 - std::pow(h->beta2, (double)t.step);
    R21Tensor& o = tb[(size_t)k];
    o.p = t.param; o.g = t.grad; o.grad_ma = t.grad_ma; o.neg_grad_ma = t.neg_grad_ma; o.v = t.variance_ma; o.slow = t.lookahead;
    o.numel = (long long)t.rows * (long long)t.cols;
    o.pnorm = pnorm;
    o.rowsum = rowsum;
    pnorm += t.rows;
    rowsum += 2 * t.rows;
    o.inv_bc2 = 1.0 / bc2;
    o.wd_lr = h->weight_decay * t.lr;
    o.rows = (int)t.rows; o.cols = (int)t.cols; o.multi_dim = t.multi_dim ? 1 : 0;
    o.lr = (float)t.lr; o.sqrt_bc2 = (float)std::sqrt(bc2); o.step_size = (float)(t.lr / bc1);
    param_size += (double)o.numel;
  }
  R21Scalars s{};
  s.param_size = param_size;
  s.b2 = (float)h->beta2; s.one_minus_b2 = (float)(1.0 - h->beta2);
  s.b1sq = (float)(h->beta1 * h->beta1); s.one_minus_b1sq = (float)(1.0 - h->beta1 * h->beta1);
  s.eps = (float)h->eps; s.agc_clip = (float)h->agc_clip; s.agc_eps = (float)h->agc_eps;
  s.normloss2 = (float)(2.0 * h->normloss_factor); s.softplus_beta = (float)h->softplus_beta;
  s.pnm_factor = (float)h->pnm_factor; s.one_plus_pnm = (float)(1.0 + h->pnm_factor);
  s.inv_noise_norm = (float)(1.0 / std::sqrt((1.0 + h->beta2) * (1.0 + h->beta2) + h->beta2 * h->beta2));
  s.la_alpha = (float)h->lookahead_alpha; s.one_minus_la_alpha = (float)(1.0 - h->lookahead_alpha);
  s.use_agc = h->use_agc != 0; s.use_gc = h->use_gc != 0; s.use_gcnorm = h->use_gcnorm != 0; s.use_normloss = h->use_normloss != 0;
  s.use_softplus = h->use_softplus != 0; s.use_decay = h->weight_decay != 0.; s.lookahead_merge = h->lookahead_merge != 0;
  cudaStream_t st = (cudaStream_t)stream;
  int launches = 0;
  {
    TimedScope ts__(F_RANGER21, 0, st);
    const cudaError_t e = launch_ranger21(n_tensors, tb.data(), s, scratch, st, &launches);
    g_launches.fetch_add((unsigned long long)launches);
    if (e != cudaSuccess) return cuda_fail(e, "ranger21");
  }
  return 0;
}

// ---- test hooks: the token-row GEMMs in isolation (tests/test_gpu_gemm.py) ------------------------------------------------------
int ib200_dbg_gemm_nt(int32_t G, int32_t B, int32_t T, const int32_t* lens, int32_t nsrc, const float* A0, const float* A1,
                      int32_t lda, int32_t K, const float* W0, const float* W1, const float* bias, float* C, int32_t ldc,
                      int32_t NC, int32_t accumulate, int32_t precision, int32_t impl, void* stream) {
  GemmNTArgs a{};
  a.G = G; a.B = B; a.Tmax = T; a.lens = lens; a.nsrc = nsrc; a.A[0] = A0; a.A[1] = A1; a.lda = lda; a.K = K;
  a.W[0] = W0; a.W[1] = W1; a.bias = bias; a.C = C; a.ldc = ldc; a.NC = NC; a.accumulate = accumulate;
  cudaStream_t st = (cudaStream_t)stream;
  CK(impl == 0 ? launch_gemm_nt(a, precision, st) : (impl == 1 ? launch_gemm_nt_tc(a, precision, st) : gemm_nt_auto(a, precision, st)),
     "dbg gemm nt");
  return 0;
}

int ib200_dbg_gemm_tn(int32_t G, int32_t B, int32_t T, const int32_t* lens, const float* A, int32_t KA, const float* Bsrc,
                      int32_t ldb, int32_t col0, int32_t shift, const int32_t* tok, const float* emb, const float* emb_row_scale,
                      int32_t V, int32_t NB, float* partial, int32_t ctas_per_group, int32_t colsum, int32_t precision,
                      int32_t impl, void* stream) {
  GemmTNArgs a{};
  a.G = G; a.B = B; a.Tmax = T; a.lens = lens; a.A = A; a.KA = KA; a.Bsrc = Bsrc; a.ldb = ldb; a.col0 = col0; a.shift = shift;
  a.tok = tok; a.emb = emb; a.emb_row_scale = emb_row_scale; a.V = V; a.NB = NB; a.NB1 = NB; a.partial = partial;
  a.ctas_per_group = ctas_per_group; a.colsum = colsum;
  cudaStream_t st = (cudaStream_t)stream;
  CK(impl == 0 ? launch_gemm_tn(a, precision, st) : (impl == 1 ? launch_gemm_tn_tc(a, precision, st) : gemm_tn_auto(a, precision, st)),
     "dbg gemm tn");
  return 0;
}

// ---- test hooks for the PRODUCTION kernels of the H = 64 path (bf16 hi|lo plane operands, TMA-fed tcgen05) -------------------------
// A row of K values is stored over the bytes of K floats: [K bf16 hi | K bf16 lo] (what the recurrent kernels write in planes mode).
int ib200_dbg_gemm_nt_planes(int32_t G, int32_t B, int32_t T, const int32_t* lens, int32_t nsrc, const float* A0, const float* A1,
                             int32_t lda, int32_t K, const float* W0, const float* W1, const float* bias, float* C, int32_t ldc,
                             int32_t NC, int32_t accumulate, int32_t precision, void* stream) {
  if (!lens || !A0 || !W0 || !C || (nsrc == 2 && (!A1 || !W1))) return fail(IB200_E_NULL, "ib200_dbg_gemm_nt_planes: null pointer");
  GemmNTArgs a{};
  a.G = G; a.B = B; a.Tmax = T; a.lens = lens; a.nsrc = nsrc; a.A[0] = A0; a.A[1] = A1; a.lda = lda; a.K = K;
  a.W[0] = W0; a.W[1] = W1; a.bias = bias; a.C = C; a.ldc = ldc; a.NC = NC; a.accumulate = accumulate; a.plane_bytes = K * 2;
  CK(gemm_nt_auto(a, precision, (cudaStream_t)stream), "dbg gemm nt planes");
  return 0;
}

int ib200_dbg_gemm_tn_planes(int32_t G, int32_t B, int32_t T, const int32_t* lens, const float* A, const float* Bsrc, int32_t ldb,
                             int32_t col0, int32_t shift, const int32_t* tok, const float* emb, const float* emb_row_scale, int32_t V,
                             int32_t NB1, const float* Bsrc2, int32_t ldb2, int32_t col02, int32_t shift2, int32_t NB2, float* partial,
                             int32_t ctas_per_group, int32_t precision, void* stream) {
  if (!lens || !A || !partial || (!tok && !Bsrc) || (NB2 > 0 && !Bsrc2)) return fail(IB200_E_NULL, "ib200_dbg_gemm_tn_planes: null pointer");
  GemmTNArgs a{};
  a.G = G; a.B = B; a.Tmax = T; a.lens = lens; a.A = A; a.KA = 256; a.Bsrc = Bsrc; a.ldb = ldb; a.col0 = col0; a.shift = shift;
  a.tok = tok; a.emb = emb; a.emb_row_scale = emb_row_scale; a.V = V; a.NB1 = NB1; a.NB = NB1 + NB2;
  a.Bsrc2 = Bsrc2; a.ldb2 = ldb2; a.col02 = col02; a.shift2 = shift2; a.partial = partial; a.ctas_per_group = ctas_per_group;
  CK(launch_gemm_tn_tma(a, precision, (cudaStream_t)stream), "dbg gemm tn planes");
  return 0;
}

size_t ib200_dbg_l0_scratch_floats(int32_t G, int32_t ndir, int32_t which) {
  return which == 0 ? l0_grad_partial_floats(G, ndir) : l0_grad_scratch_floats(G, ndir);
}

// d_* arrays: [2] pointers indexed by direction (entries of a direction that is not run may be null)
int ib200_dbg_l0_grads(int32_t G, int32_t B, int32_t T, int32_t V, const int32_t* lens, const int32_t* tok, const float* const* dA,
                       const float* Y0, const float* emb, const float* emb_row_scale, const float* whh_mask, const float* const* w_ih,
                       const float* bias_partial, int32_t bias_count, int32_t dir0, int32_t ndir, float* partial, float* scratch,
                       float* const* d_wih, float* const* d_whh, float* const* d_bih, float* const* d_bhh, float* d_emb,
                       int32_t precision, void* stream) {
  if (!lens || !tok || !dA || !Y0 || !emb || !w_ih || !bias_partial || !partial || !scratch || !d_wih || !d_whh || !d_bih || !d_bhh || !d_emb)
    return fail(IB200_E_NULL, "ib200_dbg_l0_grads: null pointer");
  L0GradArgs la{};
  la.G = G; la.B = B; la.Tmax = T; la.V = V; la.H = 64; la.dir0 = dir0; la.ndir = ndir; la.lens = lens; la.tok = tok;
  for (int d = 0; d < 2; ++d) {
    la.dA[d] = dA[d]; la.w_ih[d] = w_ih[d];
    la.d_wih[d] = d_wih[d]; la.d_whh[d] = d_whh[d]; la.d_bih[d] = d_bih[d]; la.d_bhh[d] = d_bhh[d];
  }
  la.Y0 = Y0; la.emb = emb; la.emb_row_scale = emb_row_scale; la.whh_mask = whh_mask;
  la.bias_partial = bias_partial; la.bias_count = bias_count; la.partial = partial; la.R = scratch; la.d_emb = d_emb;
  CK(launch_l0_grads(la, precision, (cudaStream_t)stream), "dbg l0 grads");
  return 0;
}

// ---- test hooks for the streamed-operand tcgen05 kernels of the H = 128 / 192 / 256 path (gemm_wide.cu) ----------------------------
// A_s and W_s are plane matrices ([rows][K] and [NC][K], a row of K values = [K bf16 hi | K bf16 lo] over the bytes of K floats)
int ib200_dbg_gemm_nt_wide(int32_t G, int32_t B, int32_t T, const int32_t* lens, int32_t nsrc, const float* A0, const float* A1,
                           int32_t lda, int32_t K, const float* W0, const float* W1, const float* bias, float* C, int32_t ldc, int32_t NC,
                           int32_t accumulate, int32_t precision, void* stream) {
  if (!lens || !A0 || !W0 || !C || (nsrc == 2 && (!A1 || !W1))) return fail(IB200_E_NULL, "ib200_dbg_gemm_nt_wide: null pointer");
  GemmNTArgs a{};
  a.G = G; a.B = B; a.Tmax = T; a.lens = lens; a.nsrc = nsrc; a.A[0] = A0; a.A[1] = A1; a.lda = lda; a.K = K;
  a.W[0] = W0; a.W[1] = W1; a.bias = bias; a.C = C; a.ldc = ldc; a.NC = NC; a.accumulate = accumulate; a.plane_bytes = K * 2;
  CK(launch_gemm_nt_wide(a, precision, (cudaStream_t)stream), "dbg gemm nt wide");
  return 0;
}

// partial: [G][splits][KA][NB1 + NB2]; splits <= 0 picks gemm_tn_wide_splits; returns the split count through *splits_out
int ib200_dbg_gemm_tn_wide(int32_t G, int32_t B, int32_t T, const int32_t* lens, const float* A, int32_t KA, const float* Bsrc, int32_t ldb,
                           int32_t col0, int32_t shift, int32_t NB1, const float* Bsrc2, int32_t ldb2, int32_t col02, int32_t shift2,
                           int32_t NB2, float* partial, int32_t splits, int32_t* splits_out, int32_t precision, void* stream) {
  if (!lens || !A || !Bsrc || (NB2 > 0 && !Bsrc2)) return fail(IB200_E_NULL, "ib200_dbg_gemm_tn_wide: null pointer");
  GemmTNArgs a{};
  a.G = G; a.B = B; a.Tmax = T; a.lens = lens; a.A = A; a.KA = KA; a.Bsrc = Bsrc; a.ldb = ldb; a.col0 = col0; a.shift = shift;
  a.NB1 = NB1; a.NB = NB1 + NB2; a.Bsrc2 = Bsrc2; a.ldb2 = ldb2; a.col02 = col02; a.shift2 = shift2; a.partial = partial;
  a.ctas_per_group = splits > 0 ? splits : gemm_tn_wide_splits(KA, a.NB, KA / 4, G);
  if (splits_out) *splits_out = a.ctas_per_group;
  if (!partial) return 0;  // query of the split count only
  CK(launch_gemm_tn_wide(a, precision, (cudaStream_t)stream), "dbg gemm tn wide");
  return 0;
}

}  // extern "C"
