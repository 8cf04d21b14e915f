"""torch custom ops + autograd wiring over the C ABI (include/ib200.h).  Torch is plumbing here: device memory, streams, the
dispatcher and the autograd graph.

The raw launchers are registered with the PyTorch dispatcher as `torch.ops.intrepppid_b200.*` (schemas below, CUDA kernels only:
there is no CPU / Meta implementation, a CPU tensor fails in the dispatcher or earlier with IB200Error):
  encoder_fwd, encoder_bwd, pool_fc_fwd, pool_fc_bwd, loss_head_fwd, loss_head_bwd, pair_score
On top of them, three differentiable ops:
  encode_hidden   tokens[G,B,T] -> top-layer final hidden states hn[2,G*B,H]   (ib200_encoder_fwd / _bwd)
  pool_fc         hn -> z[G*B,H]                                               (ib200_pool_fc_fwd / _bwd)
  loss_head       z[5,B,H], y -> (loss, classifier_loss, triplet_loss, y_hat)  (ib200_loss_head_fwd / _bwd)
and `pair_score` (inference only).  Every op requires CUDA tensors and raises otherwise -- there is no CPU path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import Cfg, EncoderParams, HeadMasks, HeadParams, check, lib, ptr


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.IB200Error("intrepppid_b200 ops run on CUDA tensors only (no CPU fallback); got a CPU tensor")


def _f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.detach().to(torch.float32).contiguous()
    return t


@dataclass
class EncoderConfig:
    num_layers: int
    bi_reduce: str
    precision: str = "fp32"

    def cfg(self, G, B, T, V, H, training) -> Cfg:
        if self.bi_reduce not in _lib.REDUCE:
            # "concat" yields [B,2H], which the reference's E->E fc cannot consume either (awd_lstm.py:47,58-60,71)
            raise ValueError(f"bi_reduce={self.bi_reduce!r} is not functional (the reference raises a shape error at fc); "
                             "use 'last', 'mean' or 'max'")
        return Cfg(G, B, T, V, H, self.num_layers, _lib.REDUCE[self.bi_reduce], _lib.PRECISION[self.precision],
                   1 if training else 0, 0)


def lstm_param_order(num_layers: int) -> List[str]:
    names = []
    for l in range(num_layers):
        for sfx in ("", "_reverse"):
            for w in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                names.append(f"{w}_l{l}{sfx}")
    return names


def _fill_encoder_struct(emb, lstm: Sequence[torch.Tensor], L: int) -> EncoderParams:
    s = EncoderParams()
    s.emb = ptr(emb)
    i = 0
    for l in range(L):
        for d in range(2):
            s.w_ih[l][d], s.w_hh[l][d], s.b_ih[l][d], s.b_hh[l][d] = (ptr(lstm[i + k]) for k in range(4))
            i += 4
    return s


# ------------------------------------------------------------------------------------------------------------------------------
# dispatcher registration: torch.ops.intrepppid_b200.*  (CUDA only)
# ------------------------------------------------------------------------------------------------------------------------------
_TORCH_LIB = torch.library.Library("intrepppid_b200", "DEF")
_TORCH_LIB.define("encoder_fwd(Tensor tokens, Tensor emb, Tensor[] lstm, Tensor? emb_row_scale, Tensor? whh_mask, int num_layers, "
                  "int bi_reduce, int precision, bool training) -> (Tensor, Tensor, Tensor)")
_TORCH_LIB.define("encoder_bwd(Tensor(a!) ws, Tensor d_hn, Tensor emb, Tensor[] lstm, Tensor? emb_row_scale, Tensor? whh_mask, int G, "
                  "int B, int T, int num_layers, int bi_reduce, int precision) -> Tensor")
_TORCH_LIB.define("pool_fc_fwd(Tensor hn, Tensor fc_w, Tensor fc_b, int bi_reduce) -> (Tensor, Tensor, Tensor)")
_TORCH_LIB.define("pool_fc_bwd(Tensor dz, Tensor pooled, Tensor? argmax, Tensor fc_w, int bi_reduce) -> (Tensor, Tensor)")
_TORCH_LIB.define("loss_head_fwd(Tensor z, Tensor y, Tensor[] params, Tensor?[] masks, float beta) -> (Tensor, Tensor)")
_TORCH_LIB.define("loss_head_bwd(Tensor z, Tensor y, Tensor[] params, Tensor?[] masks, float beta, Tensor d_loss, Tensor? d_y_hat) "
                  "-> (Tensor, Tensor)")
_TORCH_LIB.define("pair_score(Tensor z, Tensor fc1_w, Tensor fc1_b, Tensor fc2_w, Tensor fc2_b, Tensor? idx_a, Tensor? idx_b) -> Tensor")
_TORCH_LIB.define("batch_metrics(Tensor y_hat, Tensor y, float threshold) -> (Tensor, Tensor)")
_TORCH_LIB.define("pair_score_range(Tensor z, Tensor fc1_w, Tensor fc1_b, Tensor fc2_w, Tensor fc2_b, int p_begin, int p_count) -> Tensor")


_TOKEN_DTYPES = {torch.int64: _lib.TOKEN_DTYPE["int64"], torch.int32: _lib.TOKEN_DTYPE["int32"],
                 torch.int16: _lib.TOKEN_DTYPE["int16"], torch.uint8: _lib.TOKEN_DTYPE["uint8"]}


def _cfg(G, B, T, V, H, L, bi_reduce, precision, training, token_dtype=0) -> Cfg:
    return Cfg(G, B, T, V, H, L, bi_reduce, precision, 1 if training else 0, token_dtype)


def _encoder_fwd_cuda(tokens, emb, lstm, emb_row_scale, whh_mask, num_layers, bi_reduce, precision, training):
    """-> (hn_top [2,G*B,H], lengths int32 [2,G], workspace uint8).  ib200_encoder_fwd."""
    G, B, T = tokens.shape
    V, H = emb.shape
    if tokens.dtype not in _TOKEN_DTYPES:
        raise _lib.IB200Error(f"token ids must be int64 / int32 / int16 / uint8, got {tokens.dtype}")
    cfg = _cfg(G, B, T, V, H, num_layers, bi_reduce, precision, training, _TOKEN_DTYPES[tokens.dtype])
    nbytes = lib().ib200_workspace_bytes(cfg)
    if nbytes == 0:
        raise _lib.IB200Error(f"unsupported encoder configuration for the sm_100a kernels: H={H} (multiple of 32 in 32..256), "
                              f"L={num_layers} (1..4), G*B*T={G * B * T} (< 2^31), T={T} (<= 11000 when H <= 64), V={V}")
    dev = tokens.device
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    lens = torch.empty(2, G, dtype=torch.int32, device=dev)
    hn = torch.empty(2, G * B, H, dtype=torch.float32, device=dev)
    P = _fill_encoder_struct(emb, lstm, num_layers)
    check(lib().ib200_encoder_fwd(cfg, ptr(tokens), P, ptr(emb_row_scale), ptr(whh_mask), ptr(lens), ptr(hn), ptr(ws), nbytes,
                                  _stream()), "ib200_encoder_fwd")
    return hn, lens, ws


def _encoder_bwd_cuda(ws, d_hn, emb, lstm, emb_row_scale, whh_mask, G, B, T, num_layers, bi_reduce, precision):
    """-> one flat fp32 gradient buffer: [d_emb | d_lstm tensors in `lstm` order] (a single allreduce bucket).  ib200_encoder_bwd."""
    V, H = emb.shape
    cfg = _cfg(G, B, T, V, H, num_layers, bi_reduce, precision, True)
    sizes = [emb.numel()] + [p.numel() for p in lstm]
    flat = torch.empty(sum(sizes), dtype=torch.float32, device=emb.device)
    views, off = [], 0
    for n, ref in zip(sizes, [emb] + list(lstm)):
        views.append(flat[off:off + n].view(ref.shape))
        off += n
    Gs = _fill_encoder_struct(views[0], views[1:], num_layers)
    P = _fill_encoder_struct(emb, lstm, num_layers)
    check(lib().ib200_encoder_bwd(cfg, P, ptr(emb_row_scale), ptr(whh_mask), ptr(d_hn), Gs, ptr(ws), ws.numel(), _stream()),
          "ib200_encoder_bwd")
    return flat


def _pool_fc_fwd_cuda(hn, fc_w, fc_b, mode):
    _, N, H = hn.shape
    z = torch.empty(N, H, dtype=torch.float32, device=hn.device)
    pooled = torch.empty(N, H, dtype=torch.float32, device=hn.device)
    argmax = torch.empty((N, H) if mode == 2 else (0,), dtype=torch.uint8, device=hn.device)
    check(lib().ib200_pool_fc_fwd(N, H, mode, ptr(hn), ptr(fc_w), ptr(fc_b), ptr(z), ptr(pooled), ptr(argmax) if mode == 2 else None,
                                  _stream()), "ib200_pool_fc_fwd")
    return z, pooled, argmax


def _pool_fc_bwd_cuda(dz, pooled, argmax, fc_w, mode):
    N, H = dz.shape
    d_hn = torch.empty(2, N, H, dtype=torch.float32, device=dz.device)
    flat = torch.empty(H * H + H, dtype=torch.float32, device=dz.device)
    d_w, d_b = flat[:H * H].view(H, H), flat[H * H:]
    check(lib().ib200_pool_fc_bwd(N, H, mode, ptr(dz), ptr(pooled), ptr(argmax) if mode == 2 else None, ptr(fc_w), ptr(d_hn), ptr(d_w),
                                  ptr(d_b), _stream()), "ib200_pool_fc_bwd")
    return d_hn, flat


def _head_structs(params, masks):
    params = list(params) + [None] * (6 - len(params))
    hp = HeadParams(*(ptr(t) for t in params))
    hm = HeadMasks(*(ptr(m) for m in masks))
    return hp, hm


def _loss_head_fwd_cuda(z, y, params, masks, beta):
    """params = [fc1_w, fc1_b, fc2_w, fc2_b (, proj_w, proj_b)], masks = [fc1_w, do1, do2, fc2_w] (None = no drop)."""
    _, B, H = z.shape
    hp, hm = _head_structs(params, masks)
    losses = torch.empty(3, dtype=torch.float32, device=z.device)
    y_hat = torch.empty(B, dtype=torch.float32, device=z.device)
    check(lib().ib200_loss_head_fwd(B, H, float(beta), ptr(z), ptr(y), hp, hm, ptr(losses), ptr(y_hat), _stream()),
          "ib200_loss_head_fwd")
    return losses, y_hat


def _loss_head_bwd_cuda(z, y, params, masks, beta, d_loss, d_y_hat):
    """-> (dz [5,B,H], flat gradient buffer [fc1_w | fc1_b | fc2_w | fc2_b (| proj_w | proj_b)])."""
    _, B, H = z.shape
    has_proj = len(params) == 6
    hp, hm = _head_structs(params, masks)
    dz = torch.empty_like(z)
    HH = H // 2
    n_flat = HH * H + HH + HH + 1 + (H * H + H if has_proj else 0)
    flat = torch.empty(n_flat, dtype=torch.float32, device=z.device)
    o = 0
    g_fc1_w = flat[o:o + HH * H]; o += HH * H
    g_fc1_b = flat[o:o + HH]; o += HH
    g_fc2_w = flat[o:o + HH]; o += HH
    g_fc2_b = flat[o:o + 1]; o += 1
    g_pw = g_pb = None
    if has_proj:
        g_pw = flat[o:o + H * H]; o += H * H
        g_pb = flat[o:o + H]
    hg = HeadParams(ptr(g_fc1_w), ptr(g_fc1_b), ptr(g_fc2_w), ptr(g_fc2_b), ptr(g_pw), ptr(g_pb))
    check(lib().ib200_loss_head_bwd(B, H, float(beta), ptr(z), ptr(y), hp, hm, ptr(d_loss), ptr(d_y_hat), ptr(dz), hg, _stream()),
          "ib200_loss_head_bwd")
    return dz, flat


def _pair_score_cuda(z, fc1_w, fc1_b, fc2_w, fc2_b, idx_a, idx_b):
    M, H = z.shape
    P = idx_a.numel() if idx_a is not None else M * (M + 1) // 2
    out = torch.empty(P, dtype=torch.float32, device=z.device)
    hp = HeadParams(ptr(fc1_w), ptr(fc1_b), ptr(fc2_w), ptr(fc2_b), None, None)
    check(lib().ib200_pair_score(M, H, ptr(z), ptr(idx_a), ptr(idx_b), P, hp, ptr(out), _stream()), "ib200_pair_score")
    return out


def _pair_score_range_cuda(z, fc1_w, fc1_b, fc2_w, fc2_b, p_begin, p_count):
    M, H = z.shape
    out = torch.empty(p_count, dtype=torch.float32, device=z.device)
    hp = HeadParams(ptr(fc1_w), ptr(fc1_b), ptr(fc2_w), ptr(fc2_b), None, None)
    check(lib().ib200_pair_score_range(M, H, ptr(z), int(p_begin), int(p_count), hp, ptr(out), _stream()), "ib200_pair_score_range")
    return out


def _batch_metrics_cuda(y_hat, y, threshold):
    """-> (float32 [5] = auroc, ap, mcc, precision, recall ; int32 [4] = tp, fp, tn, fn).  ib200_batch_metrics."""
    B = y_hat.numel()
    out = torch.empty(5, dtype=torch.float32, device=y_hat.device)
    conf = torch.empty(4, dtype=torch.int32, device=y_hat.device)
    check(lib().ib200_batch_metrics(B, ptr(y_hat), ptr(y), float(threshold), ptr(out), ptr(conf), _stream()), "ib200_batch_metrics")
    return out, conf


for _name, _fn in (("batch_metrics", _batch_metrics_cuda), ("pair_score_range", _pair_score_range_cuda), ("encoder_fwd", _encoder_fwd_cuda), ("encoder_bwd", _encoder_bwd_cuda), ("pool_fc_fwd", _pool_fc_fwd_cuda),
                   ("pool_fc_bwd", _pool_fc_bwd_cuda), ("loss_head_fwd", _loss_head_fwd_cuda), ("loss_head_bwd", _loss_head_bwd_cuda),
                   ("pair_score", _pair_score_cuda)):
    _TORCH_LIB.impl(_name, _fn, "CUDA")
_OPS = torch.ops.intrepppid_b200


class _EncodeHidden(torch.autograd.Function):
    """tokens, masks, emb, 8L LSTM tensors -> hn_top [2,N,H].  Saved activations live in one workspace tensor."""

    @staticmethod
    def forward(ctx, econf: EncoderConfig, tokens, emb_row_scale, whh_mask, check_lengths, lengths_holder, emb, *lstm):
        G, B, T = tokens.shape
        V, H = emb.shape
        L = econf.num_layers
        training = any(ctx.needs_input_grad[6:])  # (grad mode is off inside Function.forward; this is the reliable signal)
        _need_cuda(tokens, emb, emb_row_scale, whh_mask, *lstm)
        if tokens.dtype not in _TOKEN_DTYPES:  # int64 as the reference ships them, or narrowed ids (SURVEY 8f input feeding)
            tokens = tokens.long()
        tokens = tokens.contiguous()
        emb_c = _f32c(emb)
        lstm_c = [_f32c(p) for p in lstm]
        ers, whm = _f32c(emb_row_scale), _f32c(whh_mask)
        if ers is not None and tuple(ers.shape) != (G, V):
            raise ValueError(f"emb_row_scale must be [G={G}, V={V}], got {tuple(ers.shape)}")
        if whm is not None and tuple(whm.shape) != (G, 4 * H, H):
            raise ValueError(f"whh_l0_mask must be [G={G}, {4 * H}, {H}], got {tuple(whm.shape)}")
        cfg = econf.cfg(G, B, T, V, H, training)  # validates bi_reduce (concat raises like the reference)
        hn, lens, ws = _OPS.encoder_fwd(tokens, emb_c, lstm_c, ers, whm, L, cfg.bi_reduce, cfg.precision, training)
        if lengths_holder is not None:
            lengths_holder.append(lens)
        if check_lengths:
            lo, hi = int(tokens.min()), int(tokens.max())
            if lo < 0 or hi >= V:  # F.embedding raises on such ids; the kernels clamp them for memory safety only
                raise IndexError(f"token id out of range: ids must lie in [0, {V}), got [{lo}, {hi}]")
            if int(lens[1].min()) <= 0:  # one host sync; the reference does two per encoder call (awd_lstm.py:53-54,149-150)
                raise RuntimeError("Expected sequence length to be larger than 0 in RNN")
        if training:
            ctx.dims = (G, B, T, L, cfg.bi_reduce, cfg.precision)
            ctx.save_for_backward(ws, emb_c, ers, whm, *lstm_c)
        return hn

    @staticmethod
    def backward(ctx, d_hn):
        ws, emb, ers, whm, *lstm = ctx.saved_tensors
        flat = _OPS.encoder_bwd(ws, _f32c(d_hn), emb, lstm, ers, whm, *ctx.dims)
        views, off = [], 0
        for ref in [emb] + list(lstm):
            views.append(flat[off:off + ref.numel()].view(ref.shape))
            off += ref.numel()
        return (None, None, None, None, None, None, *views)


def encode_hidden(econf: EncoderConfig, tokens, emb, lstm: Sequence[torch.Tensor], emb_row_scale=None, whh_mask=None,
                  check_lengths=True, lengths_holder=None):
    return _EncodeHidden.apply(econf, tokens, emb_row_scale, whh_mask, check_lengths, lengths_holder, emb, *lstm)


class _PoolFc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, bi_reduce: str, hn, fc_w, fc_b):
        _need_cuda(hn, fc_w, fc_b)
        hn, fc_w, fc_b = _f32c(hn), _f32c(fc_w), _f32c(fc_b)
        mode = _lib.REDUCE[bi_reduce]
        z, pooled, argmax = _OPS.pool_fc_fwd(hn, fc_w, fc_b, mode)
        ctx.mode = mode
        ctx.save_for_backward(pooled, argmax, fc_w)
        return z

    @staticmethod
    def backward(ctx, dz):
        pooled, argmax, fc_w = ctx.saved_tensors
        H = fc_w.shape[0]
        d_hn, flat = _OPS.pool_fc_bwd(_f32c(dz), pooled, argmax if ctx.mode == 2 else None, fc_w, ctx.mode)
        return None, d_hn, flat[:H * H].view(H, H), flat[H * H:]


def pool_fc(bi_reduce: str, hn, fc_w, fc_b):
    return _PoolFc.apply(bi_reduce, hn, fc_w, fc_b)


class _LossHead(torch.autograd.Function):
    """z[5,B,H], y -> losses[3], y_hat[B].  Backward recomputes the (tiny) forward intermediates inside the kernel."""

    @staticmethod
    def forward(ctx, beta, z, y, m_fc1, m_do1, m_do2, m_fc2, fc1_w, fc1_b, fc2_w, fc2_b, proj_w, proj_b):
        _need_cuda(z, y, fc1_w, fc1_b, fc2_w, fc2_b, proj_w, proj_b, m_fc1, m_do1, m_do2, m_fc2)
        z = _f32c(z)
        G5, B, H = z.shape
        if G5 != 5:
            raise ValueError("loss_head expects z of shape [5,B,H] in group order (anchor, positive, negative, p1, p2)")
        y = y.long().contiguous()
        params = [_f32c(t) for t in (fc1_w, fc1_b, fc2_w, fc2_b)] + ([_f32c(proj_w), _f32c(proj_b)] if proj_w is not None else [])
        masks = [_f32c(t) for t in (m_fc1, m_do1, m_do2, m_fc2)]
        losses, y_hat = _OPS.loss_head_fwd(z, y, params, masks, float(beta))
        ctx.set_materialize_grads(False)  # an unused output (normally y_hat) arrives as None instead of a zero-filled tensor
        ctx.beta, ctx.npar = float(beta), len(params)
        ctx.save_for_backward(z, y, *params, *[t for t in masks if t is not None])
        ctx.mask_present = [m is not None for m in masks]
        return losses, y_hat

    @staticmethod
    def backward(ctx, d_losses, d_y_hat):
        saved = list(ctx.saved_tensors)
        z, y = saved[0], saved[1]
        params = saved[2:2 + ctx.npar]
        rest = saved[2 + ctx.npar:]
        masks = [rest.pop(0) if present else None for present in ctx.mask_present]
        _, B, H = z.shape
        if d_losses is None:
            d_loss = torch.zeros(1, dtype=torch.float32, device=z.device)
        else:
            # only `loss` (element 0) is an optimisation target; classifier/triplet losses are logged detached by the reference
            d_loss = _f32c(d_losses)[0:1].contiguous()
        dz, flat = _OPS.loss_head_bwd(z, y, params, masks, ctx.beta, d_loss, _f32c(d_y_hat))
        HH, o = H // 2, 0
        g_fc1_w = flat[o:o + HH * H].view(HH, H); o += HH * H
        g_fc1_b = flat[o:o + HH]; o += HH
        g_fc2_w = flat[o:o + HH].view(1, HH); o += HH
        g_fc2_b = flat[o:o + 1]; o += 1
        g_pw = g_pb = None
        if ctx.npar == 6:
            g_pw = flat[o:o + H * H].view(H, H); o += H * H
            g_pb = flat[o:o + H]
        return None, dz, None, None, None, None, None, g_fc1_w, g_fc1_b, g_fc2_w, g_fc2_b, g_pw, g_pb


def loss_head(beta, z, y, fc1_w, fc1_b, fc2_w, fc2_b, proj_w=None, proj_b=None, masks=(None, None, None, None)):
    return _LossHead.apply(beta, z, y, *masks, fc1_w, fc1_b, fc2_w, fc2_b, proj_w, proj_b)


@torch.no_grad()
def pair_score(z, fc1_w, fc1_b, fc2_w, fc2_b, idx_a=None, idx_b=None):
    """sigmoid(head(z[i], z[j])) in eval mode for explicit pairs, or for the whole upper triangle (i<=j) when no indices are given."""
    _need_cuda(z, fc1_w, fc1_b, fc2_w, fc2_b, idx_a, idx_b)
    if idx_a is not None:
        idx_a, idx_b = idx_a.to(torch.int32).contiguous(), idx_b.to(torch.int32).contiguous()
    return _OPS.pair_score(_f32c(z), _f32c(fc1_w), _f32c(fc1_b), _f32c(fc2_w), _f32c(fc2_b), idx_a, idx_b)


@torch.no_grad()
def pair_score_range(z, fc1_w, fc1_b, fc2_w, fc2_b, p_begin: int, p_count: int):
    """Scores of the flat upper-triangle pair indices [p_begin, p_begin + p_count) (row-major, i <= j) of the M embeddings in z."""
    _need_cuda(z, fc1_w, fc1_b, fc2_w, fc2_b)
    return _OPS.pair_score_range(_f32c(z), _f32c(fc1_w), _f32c(fc1_b), _f32c(fc2_w), _f32c(fc2_b), int(p_begin), int(p_count))


def draw_masks(specs, device, seed: int, offset: int):
    """All Bernoulli(keep)/keep masks of a step in ONE launch (`ib200_draw_masks`, Philox4x32-10).
    specs: [(shape, keep_prob, row_len)] with row_len > 1 for one draw per row (variational row masks) or 0.
    Returns ([mask tensors, views of one flat buffer], counters consumed -- add them to `offset` for the next call)."""
    import ctypes as C

    sizes = [int(torch.Size(shape).numel()) for shape, _, _ in specs]
    starts, total = [], 0
    for n in sizes:
        starts.append(total)
        total += (n + 3) // 4 * 4  # every mask starts 16-byte aligned: float4 stores
    flat = torch.empty(total, dtype=torch.float32, device=device)
    _need_cuda(flat)
    arr = (_lib.MaskSpec * len(specs))()
    out = []
    for i, ((shape, keep, row_len), n, st0) in enumerate(zip(specs, sizes, starts)):
        view = flat[st0:st0 + n].view(shape)
        out.append(view)
        arr[i] = _lib.MaskSpec(view.data_ptr(), n, float(keep), int(row_len))
    used = C.c_uint64(0)
    check(lib().ib200_draw_masks(len(specs), arr, int(seed) & 0xFFFFFFFFFFFFFFFF, int(offset) & 0xFFFFFFFFFFFFFFFF, C.byref(used),
                                 _stream()), "ib200_draw_masks")
    return out, int(used.value)


METRIC_NAMES = ("auroc", "ap", "mcc", "precision", "rec")  # suffixes of the reference's log keys (e2e_triplet.py:171-184)


@torch.no_grad()
def batch_metrics(y_hat, y, threshold: float = 0.5):
    """Batch AUROC / AP / MCC / precision / recall as torchmetrics' binary metrics return them from `metric(y_hat, y)`
    (e2e_triplet.py:171-184) -> (float32 [5] in METRIC_NAMES order, int32 [4] = tp, fp, tn, fn), one launch, no host sync."""
    _need_cuda(y_hat, y)
    return _OPS.batch_metrics(_f32c(y_hat.reshape(-1)), y.reshape(-1).long().contiguous(), float(threshold))
