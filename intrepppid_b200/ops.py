"""torch custom ops + autograd wiring over the C ABI (include/ib200.h).  Torch is plumbing here: device memory, streams, the
dispatcher and the autograd graph.

The raw launchers are registered with the PyTorch dispatcher as `torch.ops.intrepppid_b200.*` by the C++ operator library
libib200_torch.so (csrc/torch_ops.cpp: TORCH_LIBRARY schemas, CUDA implementations that call the C ABI, Meta shape functions;
there is no CPU kernel -- a CPU tensor fails in the dispatcher or earlier with IB200Error):
  encoder_fwd, encoder_bwd, encoder_bwd_layers, pool_fc_fwd, pool_fc_bwd, loss_head_fwd, loss_head_bwd, pair_score,
  pair_score_range, batch_metrics
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
from ._lib import Cfg, check, lib


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


# ------------------------------------------------------------------------------------------------------------------------------
# operator library: torch.ops.intrepppid_b200.*  (C++ TORCH_LIBRARY shim csrc/torch_ops.cpp: CUDA kernels + Meta shape functions)
# ------------------------------------------------------------------------------------------------------------------------------
def _load_operator_library():
    import os

    path = os.path.join(_lib.PKG, "libib200_torch.so")
    if not os.path.exists(path):
        raise _lib.IB200Error(f"{path} is missing: build it with `python -m intrepppid_b200.build` "
                              "(intrepppid_b200 has no CPU or PyTorch fallback)")
    lib()  # libib200.so first: the shim resolves the C ABI against it
    torch.ops.load_library(path)
    return torch.ops.intrepppid_b200


class _Ops:
    """torch.ops.intrepppid_b200 with the C ABI's status codes surfaced as IB200Error (the shim raises them as RuntimeError
    tagged "[ib200]"; everything else -- dispatcher errors for CPU tensors, shape errors -- passes through unchanged).
    The libraries are loaded when the package is imported; if they are missing or stale at that point (a fresh checkout before
    `python -m intrepppid_b200.build`), the import still succeeds and the FIRST USE raises IB200Error -- there is no fallback."""

    def __init__(self):
        self._ns = None
        self._error = None
        self.load()

    def load(self):
        try:
            _lib._lib = None
            self._ns = _load_operator_library()
            self._error = None
        except (_lib.IB200Error, OSError, AttributeError) as e:
            self._ns, self._error = None, e
        for name in [k for k in self.__dict__ if not k.startswith("_")]:
            delattr(self, name)
        return self._ns is not None

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        if self._ns is None and not self.load():
            raise _lib.IB200Error(f"the CUDA libraries of intrepppid_b200 are not usable ({self._error}); build them with "
                                  "`python -m intrepppid_b200.build` (there is no CPU or PyTorch fallback)")
        op = getattr(self._ns, name)

        def call(*args):
            try:
                return op(*args)
            except RuntimeError as e:
                if "[ib200]" in str(e):
                    raise _lib.IB200Error(str(e).split("\n")[0]) from None
                raise

        call.__name__ = name
        setattr(self, name, call)
        return call


_OPS = _Ops()

_TOKEN_DTYPES = (torch.int64, torch.int32, torch.int16, torch.uint8)


# ------------------------------------------------------------------------------------------------------------------------------
# error behaviour without host syncs.  The reference raises immediately on a token id outside [0,V) (F.embedding,
# utils/embedding_do.py:35-43) and on an all-pad batch (nn.LSTM: "Expected sequence length to be larger than 0", quirk Q13) -- it
# pays two host syncs per encoder call for its truncations anyway (awd_lstm.py:53-54,149-150).  Here the kernels record both
# conditions in a device-side status word (ib200_encoder_status); it travels to pinned host memory asynchronously and is
# examined the next time the op is entered (or when `check_pending(sync=True)` is called), i.e. at most one call late and
# without ever blocking the host.  `check_lengths="sync"` restores the immediate (synchronising) check.
# ------------------------------------------------------------------------------------------------------------------------------
class _PendingStatus:
    __slots__ = ("host", "event", "V")

    def __init__(self, status: torch.Tensor, V: int):
        self.host = torch.empty(status.shape, dtype=torch.int32, pin_memory=True)
        self.host.copy_(status, non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record()
        self.V = V

    def examine(self):
        st = self.host
        if int(st[2].max()) & 1:
            raise IndexError(f"token id out of range: ids must lie in [0, {self.V}) (detected by the encoder kernels; the ids were "
                             "clamped for memory safety)")
        if int(st[1].min()) <= 0:
            raise RuntimeError("Expected sequence length to be larger than 0 in RNN")


_PENDING: List[_PendingStatus] = []


def check_pending(sync: bool = False) -> None:
    """Raise the errors recorded by earlier encoder calls (out-of-range token id -> IndexError, all-pad batch -> RuntimeError).
    sync=False looks only at status words that have already arrived on the host (never blocks); sync=True waits for all."""
    while _PENDING:
        head = _PENDING[0]
        if sync:
            head.event.synchronize()
        elif not head.event.query():
            return
        _PENDING.pop(0)
        try:
            head.examine()
        except Exception:
            _PENDING.clear()
            raise


# Data-parallel hook (parallel.GradientAllReducer): called from inside the encoder backward as soon as the gradients of the layers
# >= 1 are final, with the flat fp32 slice that holds them -- the all-reduce of that slice then runs under the layer-0 BPTT kernel.
EARLY_GRAD_HOOKS: list = []


class _EncodeHidden(torch.autograd.Function):
    """tokens, masks, emb, 8L LSTM tensors -> hn_top [2,N,H].  Saved activations live in one workspace tensor."""

    @staticmethod
    def forward(ctx, econf: EncoderConfig, tokens, emb_row_scale, whh_mask, check_lengths, lengths_holder, emb, *lstm):
        G, B, T = tokens.shape
        V, H = emb.shape
        L = econf.num_layers
        training = any(ctx.needs_input_grad[6:])  # (grad mode is off inside Function.forward; this is the reliable signal)
        _need_cuda(tokens, emb, emb_row_scale, whh_mask, *lstm)
        if check_lengths:
            check_pending()  # errors recorded by earlier calls whose status word has reached the host by now (non-blocking)
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
        hn, status, ws = _OPS.encoder_fwd(tokens, emb_c, lstm_c, ers, whm, L, cfg.bi_reduce, cfg.precision, training)
        if lengths_holder is not None:
            lengths_holder.append(status[:2])
        if check_lengths:
            _PENDING.append(_PendingStatus(status, V))
            if check_lengths == "sync":  # the reference's timing: raise before returning (one host sync)
                check_pending(sync=True)
        if training:
            ctx.dims = (G, B, T, L, cfg.bi_reduce, cfg.precision)
            ctx.consumed = False
            ctx.save_for_backward(ws, emb_c, ers, whm, *lstm_c)
        return hn

    @staticmethod
    def backward(ctx, d_hn):
        if ctx.consumed:
            # ib200_encoder_bwd overwrites the saved gates in place with the gate gradients: a second pass would read garbage
            raise RuntimeError("the encoder's saved activations were consumed by the first backward (the BPTT kernels work in place); "
                               "retain_graph=True / double backward through encode_hidden is not supported")
        ctx.consumed = True
        ws, emb, ers, whm, *lstm = ctx.saved_tensors
        G, B, T, L, bi, prec = ctx.dims
        refs = [emb] + list(lstm)
        flat = torch.empty(sum(r.numel() for r in refs), dtype=torch.float32, device=emb.device)
        d_hn = _f32c(d_hn)
        if L > 1 and EARLY_GRAD_HOOKS:
            # layers L-1 .. 1 first; their gradients (the tail of `flat`) are final here and can be all-reduced under layer 0's BPTT
            _OPS.encoder_bwd_layers(ws, d_hn, emb, lstm, ers, whm, G, B, T, L, bi, prec, flat, L - 1, 1)
            upper = flat[sum(r.numel() for r in refs[:9]):]
            for hook in list(EARLY_GRAD_HOOKS):
                hook(upper)
            _OPS.encoder_bwd_layers(ws, d_hn, emb, lstm, ers, whm, G, B, T, L, bi, prec, flat, 0, 0)
        else:
            _OPS.encoder_bwd_layers(ws, d_hn, emb, lstm, ers, whm, G, B, T, L, bi, prec, flat, L - 1, 0)
        views, off = [], 0
        for ref in refs:
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
    """z[5,B,H], y -> loss, classifier_loss, triplet_loss, y_hat[B].  Only `loss` and `y_hat` are differentiable: the reference
    logs the two component losses detached (e2e_triplet.py:138-170), and marking them non-differentiable makes a backward through
    them an autograd error instead of a silently dropped gradient.  Backward recomputes the (tiny) forward intermediates."""

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
        loss, classifier_loss, triplet_loss = losses[0], losses[1], losses[2]
        ctx.mark_non_differentiable(classifier_loss, triplet_loss)
        return loss, classifier_loss, triplet_loss, y_hat

    @staticmethod
    def backward(ctx, d_loss, _d_cl, _d_tl, d_y_hat):
        saved = list(ctx.saved_tensors)
        z, y = saved[0], saved[1]
        params = saved[2:2 + ctx.npar]
        rest = saved[2 + ctx.npar:]
        masks = [rest.pop(0) if present else None for present in ctx.mask_present]
        _, B, H = z.shape
        if d_loss is None:
            d_loss = torch.zeros(1, dtype=torch.float32, device=z.device)
        else:
            d_loss = _f32c(d_loss).reshape(1)
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
    """-> ((loss, classifier_loss, triplet_loss), y_hat); the two component losses carry no gradient (see _LossHead)."""
    loss, classifier_loss, triplet_loss, y_hat = _LossHead.apply(beta, z, y, *masks, fc1_w, fc1_b, fc2_w, fc2_b, proj_w, proj_b)
    return (loss, classifier_loss, triplet_loss), y_hat


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
