// torch operator library `intrepppid_b200` over the C ABI of libib200.so (include/ib200.h) -- SURVEY 8b "C++ side".
//
//   TORCH_LIBRARY(intrepppid_b200)              schemas
//   TORCH_LIBRARY_IMPL(intrepppid_b200, CUDA)   validation (TORCH_CHECK) + output allocation (caching allocator) + one C-ABI call on
//                                               the current CUDA stream; a non-zero status becomes a Python exception
//   TORCH_LIBRARY_IMPL(intrepppid_b200, Meta)   shape inference only (FakeTensor / torch.compile tracing of the surrounding
//                                               Lightning step, torch.library.opcheck)
// There is no CPU kernel: a CPU tensor fails in the dispatcher.  No arithmetic lives here.
//
// Reference seam replaced: nn.LSTM.forward -> torch._VF.lstm plus F.embedding / F.dropout / nn.Linear / the loss modules
// (encoders/awd_lstm.py:35-74,147-155; e2e/e2e_triplet.py:105-136; classifier/head/mlp.py:35-68).
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <cstdio>
#include <tuple>
#include <vector>

#include "../../include/ib200.h"

namespace {

using at::Tensor;
using c10::optional;

// Every message with numbers in it is formatted with snprintf and handed to TORCH_CHECK as ONE string (c10::str's stream formatting of
// integer arguments crashed in this build; tests/test_host_abi.py walks the error paths).
#define IB200_REQUIRE(cond, ...)                      \
  do {                                                \
    if (!(cond)) {                                    \
      char msg__[640];                                \
      snprintf(msg__, sizeof(msg__), __VA_ARGS__);    \
      TORCH_CHECK(false, "[ib200] ", msg__);          \
    }                                                 \
  } while (0)
#define IB200_CHECK(call, what)                                                                                                  \
  do {                                                                                                                           \
    const int st__ = (call);                                                                                                     \
    IB200_REQUIRE(st__ == 0, "%s: %s %d: %s", what, (st__ < 0 ? "invalid argument" : "CUDA error"), st__, ib200_last_error());   \
  } while (0)

void* stream_of(const Tensor& t) { return (void*)c10::cuda::getCurrentCUDAStream(t.get_device()).stream(); }

const float* fptr(const Tensor& t) { return t.const_data_ptr<float>(); }
const float* fptr(const optional<Tensor>& t) { return t.has_value() && t->defined() ? t->const_data_ptr<float>() : nullptr; }

void need_f32(const Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda(), "[ib200] ", name, " must be a CUDA tensor (there is no CPU path)");
  TORCH_CHECK(t.scalar_type() == at::kFloat && t.is_contiguous(), "[ib200] ", name, " must be contiguous float32");
}
void need_f32(const optional<Tensor>& t, const char* name) {
  if (t.has_value() && t->defined()) need_f32(*t, name);
}

int token_dtype_of(const Tensor& tokens) {
  switch (tokens.scalar_type()) {
    case at::kLong: return IB200_TOK_I64;
    case at::kInt: return IB200_TOK_I32;
    case at::kShort: return IB200_TOK_I16;
    case at::kByte: return IB200_TOK_U8;
    default: TORCH_CHECK(false, "[ib200] token ids must be int64 / int32 / int16 / uint8, got ", c10::toString(tokens.scalar_type()));
  }
}

ib200_cfg make_cfg(int64_t G, int64_t B, int64_t T, int64_t V, int64_t H, int64_t L, int64_t bi_reduce, int64_t precision,
                   bool training, int token_dtype) {
  ib200_cfg c;
  c.G = (int32_t)G; c.B = (int32_t)B; c.T = (int32_t)T; c.V = (int32_t)V; c.H = (int32_t)H; c.L = (int32_t)L;
  c.bi_reduce = (int32_t)bi_reduce; c.precision = (int32_t)precision; c.training = training ? 1 : 0; c.token_dtype = token_dtype;
  return c;
}

size_t workspace_bytes_checked(const ib200_cfg& c) {
  const size_t n = ib200_workspace_bytes(&c);
  IB200_REQUIRE(n != 0,
                "unsupported encoder configuration for the sm_100a kernels: H=%d (multiple of 32 in 32..256), L=%d (1..4), G*B*T=%lld "
                "(< 2^31), T=%d (<= 11000 when H <= 64), V=%d (2..28672), bi_reduce=%d (0 last, 1 mean, 2 max; 'concat' is not functional "
                "in the reference either)", c.H, c.L, (long long)c.G * c.B * c.T, c.T, c.V, c.bi_reduce);
  return n;
}

// the 8L LSTM tensors arrive in C-ABI order: for l, for d: weight_ih, weight_hh, bias_ih, bias_hh
template <typename Struct, typename Ptr>
void fill_lstm(Struct& s, const std::vector<Ptr>& p, int64_t L) {
  size_t i = 0;
  for (int64_t l = 0; l < L; ++l)
    for (int d = 0; d < 2; ++d) {
      s.w_ih[l][d] = p[i]; s.w_hh[l][d] = p[i + 1]; s.b_ih[l][d] = p[i + 2]; s.b_hh[l][d] = p[i + 3];
      i += 4;
    }
}

ib200_encoder_params encoder_params(const Tensor& emb, at::TensorList lstm, int64_t L) {
  IB200_REQUIRE(L >= 1 && L <= IB200_MAX_LAYERS && (int64_t)lstm.size() == 8 * L, "expected %lld LSTM tensors for %lld layers (1..4), got %zu",
                (long long)(8 * L), (long long)L, lstm.size());
  need_f32(emb, "emb");
  TORCH_CHECK(emb.dim() == 2, "[ib200] emb must be [V,H]");
  const int64_t H = emb.size(1);
  std::vector<const float*> ptrs;
  for (size_t i = 0; i < lstm.size(); ++i) {
    need_f32(lstm[i], "LSTM parameter");
    const int64_t l = (int64_t)i / 8, k = (int64_t)i % 4;
    const int64_t want = k == 0 ? 4 * H * (l == 0 ? H : 2 * H) : (k == 1 ? 4 * H * H : 4 * H);
    IB200_REQUIRE(lstm[i].numel() == want, "LSTM tensor %zu has %lld elements, expected %lld", i, (long long)lstm[i].numel(), (long long)want);
    ptrs.push_back(fptr(lstm[i]));
  }
  ib200_encoder_params P{};
  P.emb = fptr(emb);
  fill_lstm(P, ptrs, L);
  return P;
}

int64_t encoder_grad_numel(const Tensor& emb, at::TensorList lstm) {
  int64_t n = emb.numel();
  for (const Tensor& t : lstm) n += t.numel();
  return n;
}

// ---- encoder ----------------------------------------------------------------------------------------------------------------------
void check_encoder_inputs(const Tensor& tokens, const Tensor& emb, const optional<Tensor>& ers, const optional<Tensor>& whm) {
  TORCH_CHECK(tokens.dim() == 3, "[ib200] tokens must be [G,B,T]");
  TORCH_CHECK(emb.dim() == 2, "[ib200] emb must be [V,H]");
  const int64_t G = tokens.size(0), V = emb.size(0), H = emb.size(1);
  if (ers.has_value() && ers->defined())
    IB200_REQUIRE(ers->dim() == 2 && ers->size(0) == G && ers->size(1) == V, "emb_row_scale must be [G=%lld, V=%lld]", (long long)G, (long long)V);
  if (whm.has_value() && whm->defined())
    IB200_REQUIRE(whm->dim() == 3 && whm->size(0) == G && whm->size(1) == 4 * H && whm->size(2) == H, "whh_l0_mask must be [G=%lld, %lld, %lld]",
                  (long long)G, (long long)(4 * H), (long long)H);
}

// -> (hn_top [2,G*B,H], status int32 [3,G] = T1 | T_eff | flags, workspace uint8)
std::tuple<Tensor, Tensor, Tensor> encoder_fwd_cuda(const Tensor& tokens, const Tensor& emb, at::TensorList lstm,
                                                    const optional<Tensor>& ers, const optional<Tensor>& whm, int64_t L,
                                                    int64_t bi_reduce, int64_t precision, bool training) {
  check_encoder_inputs(tokens, emb, ers, whm);
  TORCH_CHECK(tokens.is_cuda() && tokens.is_contiguous(), "[ib200] tokens must be a contiguous CUDA tensor (there is no CPU path)");
  need_f32(ers, "emb_row_scale");
  need_f32(whm, "whh_l0_mask");
  const c10::cuda::CUDAGuard guard(tokens.device());
  const int64_t G = tokens.size(0), B = tokens.size(1), T = tokens.size(2), V = emb.size(0), H = emb.size(1);
  const ib200_cfg cfg = make_cfg(G, B, T, V, H, L, bi_reduce, precision, training, token_dtype_of(tokens));
  const size_t nbytes = workspace_bytes_checked(cfg);
  const ib200_encoder_params P = encoder_params(emb, lstm, L);
  Tensor ws = at::empty({(int64_t)nbytes}, tokens.options().dtype(at::kByte));
  Tensor status = at::empty({3, G}, tokens.options().dtype(at::kInt));
  Tensor hn = at::empty({2, G * B, H}, emb.options());
  void* st = stream_of(tokens);
  IB200_CHECK(ib200_encoder_fwd(&cfg, tokens.const_data_ptr(), &P, fptr(ers), fptr(whm), nullptr, hn.data_ptr<float>(), ws.data_ptr(),
                                nbytes, st), "ib200_encoder_fwd");
  IB200_CHECK(ib200_encoder_status(&cfg, ws.data_ptr(), nbytes, status.data_ptr<int32_t>(), st), "ib200_encoder_status");
  return {hn, status, ws};
}

std::tuple<Tensor, Tensor, Tensor> encoder_fwd_meta(const Tensor& tokens, const Tensor& emb, at::TensorList lstm,
                                                    const optional<Tensor>& ers, const optional<Tensor>& whm, int64_t L,
                                                    int64_t bi_reduce, int64_t precision, bool training) {
  check_encoder_inputs(tokens, emb, ers, whm);
  const int64_t G = tokens.size(0), B = tokens.size(1), T = tokens.size(2), V = emb.size(0), H = emb.size(1);
  const ib200_cfg cfg = make_cfg(G, B, T, V, H, L, bi_reduce, precision, training, token_dtype_of(tokens));
  const size_t nbytes = workspace_bytes_checked(cfg);  // pure host arithmetic
  return {at::empty({2, G * B, H}, emb.options()), at::empty({3, G}, tokens.options().dtype(at::kInt)),
          at::empty({(int64_t)nbytes}, tokens.options().dtype(at::kByte))};
}

// writes the gradients of layers layer_hi .. layer_lo into `flat` = [d_emb | d_lstm tensors in `lstm` order] (d_emb with layer 0)
void encoder_bwd_layers_cuda(Tensor& ws, const Tensor& d_hn, const Tensor& emb, at::TensorList lstm, const optional<Tensor>& ers,
                             const optional<Tensor>& whm, int64_t G, int64_t B, int64_t T, int64_t L, int64_t bi_reduce,
                             int64_t precision, Tensor& flat, int64_t layer_hi, int64_t layer_lo) {
  need_f32(d_hn, "d_hn");
  need_f32(flat, "flat gradient buffer");
  need_f32(ers, "emb_row_scale");
  need_f32(whm, "whh_l0_mask");
  TORCH_CHECK(ws.is_cuda() && ws.scalar_type() == at::kByte && ws.is_contiguous(), "[ib200] workspace must be the uint8 tensor encoder_fwd returned");
  const c10::cuda::CUDAGuard guard(ws.device());
  const int64_t V = emb.size(0), H = emb.size(1);
  TORCH_CHECK(d_hn.numel() == 2 * G * B * H, "[ib200] d_hn must be [2, G*B, H]");
  TORCH_CHECK(flat.numel() == encoder_grad_numel(emb, lstm), "[ib200] flat gradient buffer has the wrong size");
  const ib200_cfg cfg = make_cfg(G, B, T, V, H, L, bi_reduce, precision, true, IB200_TOK_I64);
  const ib200_encoder_params P = encoder_params(emb, lstm, L);
  ib200_encoder_grads Gr{};
  float* base = flat.data_ptr<float>();
  Gr.emb = base;
  std::vector<float*> gp;
  int64_t off = emb.numel();
  for (const Tensor& t : lstm) {
    gp.push_back(base + off);
    off += t.numel();
  }
  fill_lstm(Gr, gp, L);
  IB200_CHECK(ib200_encoder_bwd_layers(&cfg, &P, fptr(ers), fptr(whm), fptr(d_hn), &Gr, ws.data_ptr(), (size_t)ws.numel(),
                                       (int32_t)layer_hi, (int32_t)layer_lo, stream_of(ws)), "ib200_encoder_bwd");
}
void encoder_bwd_layers_meta(Tensor& ws, const Tensor& d_hn, const Tensor& emb, at::TensorList lstm, const optional<Tensor>& ers,
                             const optional<Tensor>& whm, int64_t G, int64_t B, int64_t T, int64_t L, int64_t bi_reduce,
                             int64_t precision, Tensor& flat, int64_t layer_hi, int64_t layer_lo) {
  TORCH_CHECK(flat.numel() == encoder_grad_numel(emb, lstm), "[ib200] flat gradient buffer has the wrong size");
  TORCH_CHECK(0 <= layer_lo && layer_lo <= layer_hi && layer_hi < L, "[ib200] need 0 <= layer_lo <= layer_hi < L");
}

Tensor encoder_bwd_cuda(Tensor& ws, const Tensor& d_hn, const Tensor& emb, at::TensorList lstm, const optional<Tensor>& ers,
                        const optional<Tensor>& whm, int64_t G, int64_t B, int64_t T, int64_t L, int64_t bi_reduce, int64_t precision) {
  Tensor flat = at::empty({encoder_grad_numel(emb, lstm)}, emb.options());
  encoder_bwd_layers_cuda(ws, d_hn, emb, lstm, ers, whm, G, B, T, L, bi_reduce, precision, flat, L - 1, 0);
  return flat;
}
Tensor encoder_bwd_meta(Tensor& ws, const Tensor& d_hn, const Tensor& emb, at::TensorList lstm, const optional<Tensor>& ers,
                        const optional<Tensor>& whm, int64_t G, int64_t B, int64_t T, int64_t L, int64_t bi_reduce, int64_t precision) {
  return at::empty({encoder_grad_numel(emb, lstm)}, emb.options());
}

// ---- pool + fc ----------------------------------------------------------------------------------------------------------------------
std::tuple<Tensor, Tensor, Tensor> pool_fc_fwd_cuda(const Tensor& hn, const Tensor& fc_w, const Tensor& fc_b, int64_t mode) {
  need_f32(hn, "hn"); need_f32(fc_w, "fc.weight"); need_f32(fc_b, "fc.bias");
  TORCH_CHECK(hn.dim() == 3 && hn.size(0) == 2, "[ib200] hn must be [2,N,H]");
  const c10::cuda::CUDAGuard guard(hn.device());
  const int64_t N = hn.size(1), H = hn.size(2);
  TORCH_CHECK(fc_w.numel() == H * H && fc_b.numel() == H, "[ib200] fc must be Linear(H,H)");
  Tensor z = at::empty({N, H}, hn.options()), pooled = at::empty({N, H}, hn.options());
  Tensor argmax = mode == IB200_REDUCE_MAX ? at::empty({N, H}, hn.options().dtype(at::kByte)) : at::empty({0}, hn.options().dtype(at::kByte));
  IB200_CHECK(ib200_pool_fc_fwd((int32_t)N, (int32_t)H, (int32_t)mode, fptr(hn), fptr(fc_w), fptr(fc_b), z.data_ptr<float>(),
                                pooled.data_ptr<float>(), mode == IB200_REDUCE_MAX ? argmax.data_ptr<uint8_t>() : nullptr,
                                stream_of(hn)), "ib200_pool_fc_fwd");
  return {z, pooled, argmax};
}
std::tuple<Tensor, Tensor, Tensor> pool_fc_fwd_meta(const Tensor& hn, const Tensor& fc_w, const Tensor& fc_b, int64_t mode) {
  TORCH_CHECK(hn.dim() == 3 && hn.size(0) == 2, "[ib200] hn must be [2,N,H]");
  const int64_t N = hn.size(1), H = hn.size(2);
  return {at::empty({N, H}, hn.options()), at::empty({N, H}, hn.options()),
          mode == IB200_REDUCE_MAX ? at::empty({N, H}, hn.options().dtype(at::kByte)) : at::empty({0}, hn.options().dtype(at::kByte))};
}

// -> (d_hn [2,N,H], flat [fc.weight grad H*H | fc.bias grad H])
std::tuple<Tensor, Tensor> pool_fc_bwd_cuda(const Tensor& dz, const Tensor& pooled, const optional<Tensor>& argmax, const Tensor& fc_w,
                                            int64_t mode) {
  need_f32(dz, "dz"); need_f32(pooled, "pooled"); need_f32(fc_w, "fc.weight");
  TORCH_CHECK(dz.dim() == 2, "[ib200] dz must be [N,H]");
  const c10::cuda::CUDAGuard guard(dz.device());
  const int64_t N = dz.size(0), H = dz.size(1);
  const bool has_arg = argmax.has_value() && argmax->defined() && argmax->numel() > 0;
  TORCH_CHECK(mode != IB200_REDUCE_MAX || (has_arg && argmax->numel() == N * H), "[ib200] bi_reduce=max needs the argmax tensor of the forward");
  Tensor d_hn = at::empty({2, N, H}, dz.options()), flat = at::empty({H * H + H}, dz.options());
  IB200_CHECK(ib200_pool_fc_bwd((int32_t)N, (int32_t)H, (int32_t)mode, fptr(dz), fptr(pooled),
                                mode == IB200_REDUCE_MAX ? argmax->const_data_ptr<uint8_t>() : nullptr, fptr(fc_w), d_hn.data_ptr<float>(),
                                flat.data_ptr<float>(), flat.data_ptr<float>() + H * H, stream_of(dz)), "ib200_pool_fc_bwd");
  return {d_hn, flat};
}
std::tuple<Tensor, Tensor> pool_fc_bwd_meta(const Tensor& dz, const Tensor& pooled, const optional<Tensor>& argmax, const Tensor& fc_w,
                                            int64_t mode) {
  const int64_t N = dz.size(0), H = dz.size(1);
  return {at::empty({2, N, H}, dz.options()), at::empty({H * H + H}, dz.options())};
}

// ---- triplet + head + BCE -----------------------------------------------------------------------------------------------------------
// params = [fc1_w, fc1_b, fc2_w, fc2_b (, proj_w, proj_b)], masks = [fc1_w, do1, do2, fc2_w] (None = no drop)
ib200_head_params head_params(at::TensorList params) {
  TORCH_CHECK(params.size() == 4 || params.size() == 6, "[ib200] head params = [fc1_w, fc1_b, fc2_w, fc2_b (, proj_w, proj_b)]");
  for (const Tensor& t : params) need_f32(t, "head parameter");
  ib200_head_params hp{};
  hp.fc1_w = fptr(params[0]); hp.fc1_b = fptr(params[1]); hp.fc2_w = fptr(params[2]); hp.fc2_b = fptr(params[3]);
  if (params.size() == 6) { hp.proj_w = fptr(params[4]); hp.proj_b = fptr(params[5]); }
  return hp;
}
ib200_head_masks head_masks(const c10::List<optional<Tensor>>& masks) {
  TORCH_CHECK(masks.size() == 4, "[ib200] head masks = [fc1_w, do1, do2, fc2_w]");
  const float* p[4];
  for (size_t i = 0; i < 4; ++i) {
    const optional<Tensor> m = masks.get(i);
    need_f32(m, "head mask");
    p[i] = fptr(m);
  }
  ib200_head_masks hm{};
  hm.fc1_w = p[0]; hm.do1 = p[1]; hm.do2 = p[2]; hm.fc2_w = p[3];
  return hm;
}
void check_head_shapes(const Tensor& z, const Tensor& y, at::TensorList params) {
  TORCH_CHECK(z.dim() == 3 && z.size(0) == 5, "[ib200] loss_head expects z of shape [5,B,H] in group order (anchor, positive, negative, p1, p2)");
  const int64_t B = z.size(1), H = z.size(2), HH = H / 2;
  TORCH_CHECK(y.numel() == B && y.scalar_type() == at::kLong, "[ib200] y must be int64 [B]");
  TORCH_CHECK(params.size() == 4 || params.size() == 6, "[ib200] head params = [fc1_w, fc1_b, fc2_w, fc2_b (, proj_w, proj_b)]");
  TORCH_CHECK(params[0].numel() == HH * H && params[1].numel() == HH && params[2].numel() == HH && params[3].numel() == 1,
              "[ib200] head parameter shapes must be [H/2,H], [H/2], [1,H/2], [1]");
}

// -> (losses [3] = loss, classifier_loss, triplet_loss ; y_hat [B])
std::tuple<Tensor, Tensor> loss_head_fwd_cuda(const Tensor& z, const Tensor& y, at::TensorList params,
                                              const c10::List<optional<Tensor>>& masks, double beta) {
  need_f32(z, "z");
  check_head_shapes(z, y, params);
  TORCH_CHECK(y.is_cuda() && y.is_contiguous(), "[ib200] y must be a contiguous CUDA tensor");
  const c10::cuda::CUDAGuard guard(z.device());
  const int64_t B = z.size(1), H = z.size(2);
  const ib200_head_params hp = head_params(params);
  const ib200_head_masks hm = head_masks(masks);
  Tensor losses = at::empty({3}, z.options()), y_hat = at::empty({B}, z.options());
  IB200_CHECK(ib200_loss_head_fwd((int32_t)B, (int32_t)H, (float)beta, fptr(z), y.const_data_ptr<int64_t>(), &hp, &hm,
                                  losses.data_ptr<float>(), y_hat.data_ptr<float>(), stream_of(z)), "ib200_loss_head_fwd");
  return {losses, y_hat};
}
std::tuple<Tensor, Tensor> loss_head_fwd_meta(const Tensor& z, const Tensor& y, at::TensorList params,
                                              const c10::List<optional<Tensor>>& masks, double beta) {
  check_head_shapes(z, y, params);
  return {at::empty({3}, z.options()), at::empty({z.size(1)}, z.options())};
}

int64_t head_grad_numel(int64_t H, bool proj) { return (H / 2) * H + H / 2 + H / 2 + 1 + (proj ? H * H + H : 0); }

// -> (dz [5,B,H], flat [fc1_w | fc1_b | fc2_w | fc2_b (| proj_w | proj_b)])
std::tuple<Tensor, Tensor> loss_head_bwd_cuda(const Tensor& z, const Tensor& y, at::TensorList params,
                                              const c10::List<optional<Tensor>>& masks, double beta, const Tensor& d_loss,
                                              const optional<Tensor>& d_y_hat) {
  need_f32(z, "z"); need_f32(d_loss, "d_loss"); need_f32(d_y_hat, "d_y_hat");
  check_head_shapes(z, y, params);
  const c10::cuda::CUDAGuard guard(z.device());
  const int64_t B = z.size(1), H = z.size(2), HH = H / 2;
  const bool proj = params.size() == 6;
  const ib200_head_params hp = head_params(params);
  const ib200_head_masks hm = head_masks(masks);
  Tensor dz = at::empty_like(z), flat = at::empty({head_grad_numel(H, proj)}, z.options());
  float* f = flat.data_ptr<float>();
  ib200_head_grads hg{};
  hg.fc1_w = f; hg.fc1_b = f + HH * H; hg.fc2_w = hg.fc1_b + HH; hg.fc2_b = hg.fc2_w + HH;
  if (proj) { hg.proj_w = hg.fc2_b + 1; hg.proj_b = hg.proj_w + H * H; }
  IB200_CHECK(ib200_loss_head_bwd((int32_t)B, (int32_t)H, (float)beta, fptr(z), y.const_data_ptr<int64_t>(), &hp, &hm, fptr(d_loss),
                                  fptr(d_y_hat), dz.data_ptr<float>(), &hg, stream_of(z)), "ib200_loss_head_bwd");
  return {dz, flat};
}
std::tuple<Tensor, Tensor> loss_head_bwd_meta(const Tensor& z, const Tensor& y, at::TensorList params,
                                              const c10::List<optional<Tensor>>& masks, double beta, const Tensor& d_loss,
                                              const optional<Tensor>& d_y_hat) {
  check_head_shapes(z, y, params);
  return {at::empty_like(z), at::empty({head_grad_numel(z.size(2), params.size() == 6)}, z.options())};
}

// ---- inference scorer + metrics -----------------------------------------------------------------------------------------------------
ib200_head_params scorer_params(const Tensor& fc1_w, const Tensor& fc1_b, const Tensor& fc2_w, const Tensor& fc2_b) {
  need_f32(fc1_w, "fc1.weight"); need_f32(fc1_b, "fc1.bias"); need_f32(fc2_w, "fc2.weight"); need_f32(fc2_b, "fc2.bias");
  ib200_head_params hp{};
  hp.fc1_w = fptr(fc1_w); hp.fc1_b = fptr(fc1_b); hp.fc2_w = fptr(fc2_w); hp.fc2_b = fptr(fc2_b);
  return hp;
}

Tensor pair_score_cuda(const Tensor& z, const Tensor& fc1_w, const Tensor& fc1_b, const Tensor& fc2_w, const Tensor& fc2_b,
                       const optional<Tensor>& idx_a, const optional<Tensor>& idx_b) {
  need_f32(z, "z");
  TORCH_CHECK(z.dim() == 2, "[ib200] z must be [M,H]");
  const c10::cuda::CUDAGuard guard(z.device());
  const int64_t M = z.size(0), H = z.size(1);
  const bool has = idx_a.has_value() && idx_a->defined();
  TORCH_CHECK(has == (idx_b.has_value() && idx_b->defined()), "[ib200] idx_a and idx_b must both be given or both be None");
  if (has)
    TORCH_CHECK(idx_a->is_cuda() && idx_b->is_cuda() && idx_a->scalar_type() == at::kInt && idx_b->scalar_type() == at::kInt &&
                    idx_a->is_contiguous() && idx_b->is_contiguous() && idx_a->numel() == idx_b->numel(),
                "[ib200] idx_a / idx_b must be contiguous int32 CUDA tensors of equal length");
  const int64_t P = has ? idx_a->numel() : M * (M + 1) / 2;
  const ib200_head_params hp = scorer_params(fc1_w, fc1_b, fc2_w, fc2_b);
  Tensor out = at::empty({P}, z.options());
  IB200_CHECK(ib200_pair_score((int32_t)M, (int32_t)H, fptr(z), has ? idx_a->const_data_ptr<int32_t>() : nullptr,
                               has ? idx_b->const_data_ptr<int32_t>() : nullptr, P, &hp, out.data_ptr<float>(), stream_of(z)),
              "ib200_pair_score");
  return out;
}
Tensor pair_score_meta(const Tensor& z, const Tensor& fc1_w, const Tensor& fc1_b, const Tensor& fc2_w, const Tensor& fc2_b,
                       const optional<Tensor>& idx_a, const optional<Tensor>& idx_b) {
  const int64_t M = z.size(0);
  const bool has = idx_a.has_value() && idx_a->defined();
  return at::empty({has ? idx_a->numel() : M * (M + 1) / 2}, z.options());
}

Tensor pair_score_range_cuda(const Tensor& z, const Tensor& fc1_w, const Tensor& fc1_b, const Tensor& fc2_w, const Tensor& fc2_b,
                             int64_t p_begin, int64_t p_count) {
  need_f32(z, "z");
  TORCH_CHECK(z.dim() == 2, "[ib200] z must be [M,H]");
  TORCH_CHECK(p_count >= 0, "[ib200] p_count must be non-negative");
  const c10::cuda::CUDAGuard guard(z.device());
  const ib200_head_params hp = scorer_params(fc1_w, fc1_b, fc2_w, fc2_b);
  Tensor out = at::empty({p_count}, z.options());
  IB200_CHECK(ib200_pair_score_range((int32_t)z.size(0), (int32_t)z.size(1), fptr(z), p_begin, p_count, &hp, out.data_ptr<float>(),
                                     stream_of(z)), "ib200_pair_score_range");
  return out;
}
Tensor pair_score_range_meta(const Tensor& z, const Tensor& fc1_w, const Tensor& fc1_b, const Tensor& fc2_w, const Tensor& fc2_b,
                             int64_t p_begin, int64_t p_count) {
  TORCH_CHECK(p_count >= 0, "[ib200] p_count must be non-negative");
  return at::empty({p_count}, z.options());
}

// -> (float32 [5] = auroc, ap, mcc, precision, recall ; int32 [4] = tp, fp, tn, fn)
std::tuple<Tensor, Tensor> batch_metrics_cuda(const Tensor& y_hat, const Tensor& y, double threshold) {
  need_f32(y_hat, "y_hat");
  TORCH_CHECK(y.is_cuda() && y.scalar_type() == at::kLong && y.is_contiguous() && y.numel() == y_hat.numel(), "[ib200] y must be contiguous int64 CUDA, one label per score");
  const c10::cuda::CUDAGuard guard(y_hat.device());
  Tensor out = at::empty({5}, y_hat.options()), conf = at::empty({4}, y_hat.options().dtype(at::kInt));
  IB200_CHECK(ib200_batch_metrics((int32_t)y_hat.numel(), fptr(y_hat), y.const_data_ptr<int64_t>(), (float)threshold,
                                  out.data_ptr<float>(), conf.data_ptr<int32_t>(), stream_of(y_hat)), "ib200_batch_metrics");
  return {out, conf};
}
std::tuple<Tensor, Tensor> batch_metrics_meta(const Tensor& y_hat, const Tensor& y, double threshold) {
  return {at::empty({5}, y_hat.options()), at::empty({4}, y_hat.options().dtype(at::kInt))};
}

}  // namespace

TORCH_LIBRARY(intrepppid_b200, m) {
  m.def("encoder_fwd(Tensor tokens, Tensor emb, Tensor[] lstm, Tensor? emb_row_scale, Tensor? whh_mask, int num_layers, int bi_reduce, "
        "int precision, bool training) -> (Tensor, Tensor, Tensor)");
  m.def("encoder_bwd(Tensor(a!) ws, Tensor d_hn, Tensor emb, Tensor[] lstm, Tensor? emb_row_scale, Tensor? whh_mask, int G, int B, int T, "
        "int num_layers, int bi_reduce, int precision) -> Tensor");
  m.def("encoder_bwd_layers(Tensor(a!) ws, Tensor d_hn, Tensor emb, Tensor[] lstm, Tensor? emb_row_scale, Tensor? whh_mask, int G, int B, "
        "int T, int num_layers, int bi_reduce, int precision, Tensor(b!) flat, int layer_hi, int layer_lo) -> ()");
  m.def("pool_fc_fwd(Tensor hn, Tensor fc_w, Tensor fc_b, int bi_reduce) -> (Tensor, Tensor, Tensor)");
  m.def("pool_fc_bwd(Tensor dz, Tensor pooled, Tensor? argmax, Tensor fc_w, int bi_reduce) -> (Tensor, Tensor)");
  m.def("loss_head_fwd(Tensor z, Tensor y, Tensor[] params, Tensor?[] masks, float beta) -> (Tensor, Tensor)");
  m.def("loss_head_bwd(Tensor z, Tensor y, Tensor[] params, Tensor?[] masks, float beta, Tensor d_loss, Tensor? d_y_hat) -> (Tensor, Tensor)");
  m.def("pair_score(Tensor z, Tensor fc1_w, Tensor fc1_b, Tensor fc2_w, Tensor fc2_b, Tensor? idx_a, Tensor? idx_b) -> Tensor");
  m.def("pair_score_range(Tensor z, Tensor fc1_w, Tensor fc1_b, Tensor fc2_w, Tensor fc2_b, int p_begin, int p_count) -> Tensor");
  m.def("batch_metrics(Tensor y_hat, Tensor y, float threshold) -> (Tensor, Tensor)");
}

TORCH_LIBRARY_IMPL(intrepppid_b200, CUDA, m) {
  m.impl("encoder_fwd", &encoder_fwd_cuda);
  m.impl("encoder_bwd", &encoder_bwd_cuda);
  m.impl("encoder_bwd_layers", &encoder_bwd_layers_cuda);
  m.impl("pool_fc_fwd", &pool_fc_fwd_cuda);
  m.impl("pool_fc_bwd", &pool_fc_bwd_cuda);
  m.impl("loss_head_fwd", &loss_head_fwd_cuda);
  m.impl("loss_head_bwd", &loss_head_bwd_cuda);
  m.impl("pair_score", &pair_score_cuda);
  m.impl("pair_score_range", &pair_score_range_cuda);
  m.impl("batch_metrics", &batch_metrics_cuda);
}

TORCH_LIBRARY_IMPL(intrepppid_b200, Meta, m) {
  m.impl("encoder_fwd", &encoder_fwd_meta);
  m.impl("encoder_bwd", &encoder_bwd_meta);
  m.impl("encoder_bwd_layers", &encoder_bwd_layers_meta);
  m.impl("pool_fc_fwd", &pool_fc_fwd_meta);
  m.impl("pool_fc_bwd", &pool_fc_bwd_meta);
  m.impl("loss_head_fwd", &loss_head_fwd_meta);
  m.impl("loss_head_bwd", &loss_head_bwd_meta);
  m.impl("pair_score", &pair_score_meta);
  m.impl("pair_score_range", &pair_score_range_meta);
  m.impl("batch_metrics", &batch_metrics_meta);
}
