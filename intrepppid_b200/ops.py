"""torch.autograd wiring over the C ABI (include/ib200.h).  Torch is plumbing here: device memory, streams, autograd graph.

Three differentiable ops:
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


class _EncodeHidden(torch.autograd.Function):
    """tokens, masks, emb, 8L LSTM tensors -> hn_top [2,N,H].  Saved activations live in one workspace tensor."""

    @staticmethod
    def forward(ctx, econf: EncoderConfig, tokens, emb_row_scale, whh_mask, check_lengths, lengths_holder, emb, *lstm):
        G, B, T = tokens.shape
        V, H = emb.shape
        L = econf.num_layers
        training = any(ctx.needs_input_grad[6:])  # (grad mode is off inside Function.forward; this is the reliable signal)
        _need_cuda(tokens, emb, emb_row_scale, whh_mask, *lstm)
        if tokens.dtype != torch.int64:
            tokens = tokens.long()
        tokens = tokens.contiguous()
        emb_c = _f32c(emb)
        lstm_c = [_f32c(p) for p in lstm]
        ers, whm = _f32c(emb_row_scale), _f32c(whh_mask)
        if ers is not None and tuple(ers.shape) != (G, V):
            raise ValueError(f"emb_row_scale must be [G={G}, V={V}], got {tuple(ers.shape)}")
        if whm is not None and tuple(whm.shape) != (G, 4 * H, H):
            raise ValueError(f"whh_l0_mask must be [G={G}, {4 * H}, {H}], got {tuple(whm.shape)}")
        cfg = econf.cfg(G, B, T, V, H, training)
        nbytes = lib().ib200_workspace_bytes(cfg)
        if nbytes == 0:
            raise _lib.IB200Error(f"unsupported encoder configuration for the sm_100a kernels: H={H} (multiple of 32 in 32..256), L={L} (1..4)")
        dev = tokens.device
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        lens = torch.empty(2, G, dtype=torch.int32, device=dev)
        hn = torch.empty(2, G * B, H, dtype=torch.float32, device=dev)
        P = _fill_encoder_struct(emb_c, lstm_c, L)
        check(lib().ib200_encoder_fwd(cfg, ptr(tokens), P, ptr(ers), ptr(whm), ptr(lens), ptr(hn), ptr(ws), nbytes, _stream()),
              "ib200_encoder_fwd")
        if lengths_holder is not None:
            lengths_holder.append(lens)
        if check_lengths:
            if int(lens[1].min()) <= 0:  # one host sync; the reference does two per encoder call (awd_lstm.py:53-54,149-150)
                raise RuntimeError("Expected sequence length to be larger than 0 in RNN")
        if training:
            ctx.econf, ctx.cfg, ctx.nbytes, ctx.L = econf, cfg, nbytes, L
            ctx.save_for_backward(ws, emb_c, ers, whm, *lstm_c)
        return hn

    @staticmethod
    def backward(ctx, d_hn):
        ws, emb, ers, whm, *lstm = ctx.saved_tensors
        L = ctx.L
        d_hn = _f32c(d_hn)
        # one flat gradient buffer (a single allreduce bucket for data parallelism)
        sizes = [emb.numel()] + [p.numel() for p in lstm]
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=emb.device)
        views, off = [], 0
        for n, ref in zip(sizes, [emb] + list(lstm)):
            views.append(flat[off:off + n].view(ref.shape))
            off += n
        Gs = _fill_encoder_struct(views[0], views[1:], L)
        P = _fill_encoder_struct(emb, lstm, L)
        check(lib().ib200_encoder_bwd(ctx.cfg, P, ptr(ers), ptr(whm), ptr(d_hn), Gs, ptr(ws), ctx.nbytes, _stream()),
              "ib200_encoder_bwd")
        return (None, None, None, None, None, None, *views)


def encode_hidden(econf: EncoderConfig, tokens, emb, lstm: Sequence[torch.Tensor], emb_row_scale=None, whh_mask=None,
                  check_lengths=True, lengths_holder=None):
    return _EncodeHidden.apply(econf, tokens, emb_row_scale, whh_mask, check_lengths, lengths_holder, emb, *lstm)


class _PoolFc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, bi_reduce: str, hn, fc_w, fc_b):
        _need_cuda(hn, fc_w, fc_b)
        hn, fc_w, fc_b = _f32c(hn), _f32c(fc_w), _f32c(fc_b)
        _, N, H = hn.shape
        mode = _lib.REDUCE[bi_reduce]
        z = torch.empty(N, H, dtype=torch.float32, device=hn.device)
        pooled = torch.empty(N, H, dtype=torch.float32, device=hn.device)
        argmax = torch.empty(N, H, dtype=torch.uint8, device=hn.device) if mode == 2 else None
        check(lib().ib200_pool_fc_fwd(N, H, mode, ptr(hn), ptr(fc_w), ptr(fc_b), ptr(z), ptr(pooled), ptr(argmax), _stream()),
              "ib200_pool_fc_fwd")
        ctx.mode = mode
        ctx.save_for_backward(pooled, argmax, fc_w)
        return z

    @staticmethod
    def backward(ctx, dz):
        pooled, argmax, fc_w = ctx.saved_tensors
        dz = _f32c(dz)
        N, H = dz.shape
        d_hn = torch.empty(2, N, H, dtype=torch.float32, device=dz.device)
        flat = torch.empty(H * H + H, dtype=torch.float32, device=dz.device)
        d_w, d_b = flat[:H * H].view(H, H), flat[H * H:]
        check(lib().ib200_pool_fc_bwd(N, H, ctx.mode, ptr(dz), ptr(pooled), ptr(argmax), ptr(fc_w), ptr(d_hn), ptr(d_w), ptr(d_b),
                                      _stream()), "ib200_pool_fc_bwd")
        return None, d_hn, d_w, d_b


def pool_fc(bi_reduce: str, hn, fc_w, fc_b):
    return _PoolFc.apply(bi_reduce, hn, fc_w, fc_b)


def _head_structs(fc1_w, fc1_b, fc2_w, fc2_b, proj_w, proj_b, masks):
    hp = HeadParams(ptr(fc1_w), ptr(fc1_b), ptr(fc2_w), ptr(fc2_b), ptr(proj_w), ptr(proj_b))
    hm = HeadMasks(*(ptr(m) for m in masks))
    return hp, hm


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
        params = [_f32c(t) for t in (fc1_w, fc1_b, fc2_w, fc2_b, proj_w, proj_b)]
        masks = [_f32c(t) for t in (m_fc1, m_do1, m_do2, m_fc2)]
        hp, hm = _head_structs(*params, masks)
        losses = torch.empty(3, dtype=torch.float32, device=z.device)
        y_hat = torch.empty(B, dtype=torch.float32, device=z.device)
        check(lib().ib200_loss_head_fwd(B, H, float(beta), ptr(z), ptr(y), hp, hm, ptr(losses), ptr(y_hat), _stream()),
              "ib200_loss_head_fwd")
        ctx.beta, ctx.has_proj = float(beta), proj_w is not None
        ctx.save_for_backward(z, y, *[t for t in params if t is not None], *[t for t in masks if t is not None])
        ctx.mask_present = [m is not None for m in masks]
        return losses, y_hat

    @staticmethod
    def backward(ctx, d_losses, d_y_hat):
        saved = list(ctx.saved_tensors)
        z, y = saved[0], saved[1]
        npar = 6 if ctx.has_proj else 4
        params = saved[2:2 + npar] + [None] * (6 - npar)
        rest = saved[2 + npar:]
        masks = []
        for present in ctx.mask_present:
            masks.append(rest.pop(0) if present else None)
        _, B, H = z.shape
        hp, hm = _head_structs(*params, masks)
        dev = z.device
        if d_losses is None:
            d_loss = torch.zeros(1, dtype=torch.float32, device=dev)
        else:
            # only `loss` (element 0) is an optimisation target; classifier/triplet losses are logged detached by the reference
            d_loss = _f32c(d_losses)[0:1].contiguous()
        dz = torch.empty_like(z)
        HH = H // 2
        n_flat = HH * H + HH + HH + 1 + (H * H + H if ctx.has_proj else 0)
        flat = torch.empty(n_flat, dtype=torch.float32, device=dev)
        o = 0
        g_fc1_w = flat[o:o + HH * H].view(HH, H); o += HH * H
        g_fc1_b = flat[o:o + HH]; o += HH
        g_fc2_w = flat[o:o + HH].view(1, HH); o += HH
        g_fc2_b = flat[o:o + 1]; o += 1
        g_pw = g_pb = None
        if ctx.has_proj:
            g_pw = flat[o:o + H * H].view(H, H); o += H * H
            g_pb = flat[o:o + H]
        hg = HeadParams(ptr(g_fc1_w), ptr(g_fc1_b), ptr(g_fc2_w), ptr(g_fc2_b), ptr(g_pw), ptr(g_pb))
        check(lib().ib200_loss_head_bwd(B, H, ctx.beta, ptr(z), ptr(y), hp, hm, ptr(d_loss), ptr(_f32c(d_y_hat)), ptr(dz), hg,
                                        _stream()), "ib200_loss_head_bwd")
        return None, dz, None, None, None, None, None, g_fc1_w, g_fc1_b, g_fc2_w, g_fc2_b, g_pw, g_pb


def loss_head(beta, z, y, fc1_w, fc1_b, fc2_w, fc2_b, proj_w=None, proj_b=None, masks=(None, None, None, None)):
    return _LossHead.apply(beta, z, y, *masks, fc1_w, fc1_b, fc2_w, fc2_b, proj_w, proj_b)


@torch.no_grad()
def pair_score(z, fc1_w, fc1_b, fc2_w, fc2_b, idx_a=None, idx_b=None):
    """sigmoid(head(z[i], z[j])) in eval mode for explicit pairs, or for the whole upper triangle (i<=j) when no indices are given."""
    _need_cuda(z, fc1_w, fc1_b, fc2_w, fc2_b, idx_a, idx_b)
    z = _f32c(z)
    M, H = z.shape
    if idx_a is not None:
        idx_a, idx_b = idx_a.to(torch.int32).contiguous(), idx_b.to(torch.int32).contiguous()
        P = idx_a.numel()
    else:
        P = M * (M + 1) // 2
    out = torch.empty(P, dtype=torch.float32, device=z.device)
    hp = HeadParams(ptr(_f32c(fc1_w)), ptr(_f32c(fc1_b)), ptr(_f32c(fc2_w)), ptr(_f32c(fc2_b)), None, None)
    check(lib().ib200_pair_score(M, H, ptr(z), ptr(idx_a), ptr(idx_b), P, hp, ptr(out), _stream()), "ib200_pair_score")
    return out
