"""The optimizer step that follows the hot path, on multi-tensor sm_100a kernels: `FusedAdamW`
(`ib200_adamw_step`; reference e2e/e2e_triplet.py:231-255, `AdamW(self.parameters(), lr=self.lr)`) and `FusedRanger21`
(`ib200_ranger21_step`; reference e2e/e2e_triplet.py:200-226, the factory default -- parity unpinned, see its docstring).

FusedAdamW:

Same constructor arguments, defaults, param_groups and state_dict layout (`step`, `exp_avg`, `exp_avg_sq` per parameter) as
torch.optim.AdamW, so learning-rate schedulers (OneCycleLR, CosineAnnealingWarmRestarts -- e2e_triplet.py:239-253) and optimizer
checkpoints interchange with the reference's.  All parameters of a group are updated by ONE kernel launch (per 32 tensors).
CUDA fp32 parameters only; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
from torch.optim import Optimizer

from . import _lib
from ._lib import AdamWHyper, check, lib


class FusedAdamW(Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 amsgrad: bool = False, *, maximize: bool = False, grad_scale: float = 1.0):
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 0: {betas[0]}")
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 1: {betas[1]}")
        if not 0.0 <= weight_decay:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        if amsgrad:
            raise ValueError("FusedAdamW does not implement amsgrad (the reference never enables it)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=maximize))
        self.grad_scale = float(grad_scale)  # multiplied into every gradient (1/world_size after a SUM all-reduce)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            live = [p for p in group["params"] if p.grad is not None]
            if not live:
                continue
            steps = set()
            grads = []
            for p in live:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise _lib.IB200Error("FusedAdamW updates contiguous fp32 CUDA parameters only (no CPU fallback)")
                if p.grad.is_sparse:
                    raise RuntimeError("AdamW does not support sparse gradients")
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.to(torch.float32).contiguous()
                grads.append(g)
                st = self.state[p]
                if len(st) == 0:  # same lazily created state as torch.optim.AdamW (step kept as a CPU float tensor)
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] = st["step"] + 1  # out of place: a state_dict loaded without a copy may share this tensor
                steps.add(int(st["step"]))
            # tensors that joined later (a gradient that was None so far) carry their own step count: one launch set per count
            for t in sorted(steps):
                sel = [(p, g) for p, g in zip(live, grads) if int(self.state[p]["step"]) == t]
                n = len(sel)
                arr = lambda xs: (C.c_void_p * n)(*xs)  # noqa: E731
                lr = group["lr"]
                hyper = AdamWHyper(float(lr), float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
                                   float(group["weight_decay"]), self.grad_scale, t, 1 if group["maximize"] else 0)
                check(lib().ib200_adamw_step(
                    n, arr([p.data_ptr() for p, _ in sel]), arr([g.data_ptr() for _, g in sel]),
                    arr([self.state[p]["exp_avg"].data_ptr() for p, _ in sel]),
                    arr([self.state[p]["exp_avg_sq"].data_ptr() for p, _ in sel]),
                    (C.c_int64 * n)(*[p.numel() for p, _ in sel]), hyper, torch.cuda.current_stream().cuda_stream),
                    "ib200_adamw_step")
        return loss


class FusedRanger21(Optimizer):
    """Ranger21 on the multi-tensor sm_100a kernels (`ib200_ranger21_step`) -- the optimizer `configure_optimizers` builds for
    optimizer_type "ranger21" / "ranger21_xx", the factory default (reference: e2e/e2e_triplet.py:200-226, intrepppid/__init__.py:37).

    Same constructor arguments and defaults as `ranger21.Ranger21` (0.1.0, the reference's pin requirements.txt:65), same
    param_groups and per-parameter state keys (`step`, `grad_ma`, `variance_ma`, `lookahead_params`, `neg_grad_ma`,
    `max_variance_ma`), so optimizer checkpoints interchange.  The package is absent from the build image: the arithmetic follows the
    published algorithm (arXiv:2106.13731) -- PARITY UNPINNED against the package (see oracle/ranger21_restated.py and
    tests/test_ranger21.py).  Implemented: the AdamW core with positive-negative momentum, AGC, gradient centralization and
    normalization, norm loss, stable weight decay, softplus, linear warm-up, warm-down, lookahead.  The package's other switches
    (madgrad / adabelief cores, Chebyshev schedule, non-pnm momentum, non-stable decay, gc_conv_only) raise: the reference never
    sets them.  Three launches per step for all tensors and NO host sync (the package pays one per step for its variance scalar);
    like the package, the step rewrites p.grad in place.  CUDA fp32 parameters only; there is no CPU path.
    """

    def __init__(self, params, lr, lookahead_active=True, lookahead_mergetime=5, lookahead_blending_alpha=0.5,
                 lookahead_load_at_validation=False, use_madgrad=False, use_adabelief=False, softplus=True, beta_softplus=50,
                 use_gc=True, use_gcnorm=True, gc_conv_only=False, normloss_active=True, normloss_factor=1e-4,
                 use_adaptive_gradient_clipping=True, agc_clipping_value=1e-2, agc_eps=1e-3, betas=(0.9, 0.999),
                 momentum_type="pnm", pnm_momentum_factor=1.0, momentum=0.9, eps=1e-8, num_batches_per_epoch=None, num_epochs=None,
                 use_cheb=False, use_warmup=True, num_warmup_iterations=None, warmdown_active=True, warmdown_start_pct=0.72,
                 warmdown_min_lr=3e-5, weight_decay=1e-4, decay_type="stable", warmup_type="linear", warmup_pct_default=0.22,
                 logging_active=True):
        for flag, name in ((use_madgrad, "use_madgrad"), (use_adabelief, "use_adabelief"), (use_cheb, "use_cheb"),
                           (gc_conv_only, "gc_conv_only"), (lookahead_load_at_validation, "lookahead_load_at_validation")):
            if flag:
                raise NotImplementedError(f"FusedRanger21: {name}=True is not implemented (the reference never sets it)")
        if momentum_type != "pnm" or decay_type != "stable" or warmup_type != "linear":
            raise NotImplementedError("FusedRanger21 implements momentum_type='pnm', decay_type='stable', warmup_type='linear' only")
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        super().__init__(params, dict(lr=lr, momentum=momentum, betas=betas, eps=eps, weight_decay=weight_decay))
        self.starting_lr = lr
        self.current_lr = lr
        self.num_batches, self.num_epochs = num_batches_per_epoch, num_epochs
        self.total_iterations = (num_epochs or 0) * (num_batches_per_epoch or 0)
        if not self.total_iterations:
            raise ValueError("missing total iterations, which is calced from num epochs and num iters per epoch param")
        self.use_warmup = use_warmup
        self.warmup_complete = False
        if num_warmup_iterations is None:  # the package's untuned linear warm-up: 2 / (1 - beta2) steps, capped by a share of the run
            beta_warmup_iters = math.ceil(2 / (1 - betas[1]))
            if beta_warmup_iters / self.total_iterations > 0.45:
                self.num_warmup_iters = int(warmup_pct_default * self.total_iterations)
            else:
                self.num_warmup_iters = beta_warmup_iters
        else:
            self.num_warmup_iters = num_warmup_iterations
        self.min_lr = warmdown_min_lr
        self.warmdown_lr_delta = self.starting_lr - self.min_lr
        self.warmdown_active = warmdown_active
        if warmdown_active:
            self.warm_down_start_pct = warmdown_start_pct
            self.start_warm_down = int(warmdown_start_pct * num_epochs * num_batches_per_epoch)
            self.warmdown_total_iterations = self.total_iterations - self.start_warm_down
        self.lookahead_active = lookahead_active
        self.lookahead_mergetime = lookahead_mergetime
        self.lookahead_alpha = lookahead_blending_alpha
        self.lookahead_step = 0
        self.softplus, self.beta_softplus = softplus, beta_softplus
        self.use_gc, self.use_gcnorm = use_gc, use_gcnorm
        self.normloss_active, self.normloss_factor = normloss_active, normloss_factor
        self.agc_active, self.agc_clip_val, self.agc_eps = use_adaptive_gradient_clipping, agc_clipping_value, agc_eps
        self.momentum_pnm, self.pnm_momentum_factor = True, pnm_momentum_factor
        self.eps = eps
        self._scratch = None  # device scratch: per-tensor variance sums + variance_normalized (doubles), row norms (floats)

    # -- learning-rate schedule (host logic; the package's warmup_dampening / get_warm_down) ----------------------------------------
    def warmup_dampening(self, lr, step):
        if step > self.num_warmup_iters:
            self.warmup_complete = True
            return lr
        return lr * min(1.0, step / self.num_warmup_iters)

    def get_warm_down(self, lr, iteration):
        if iteration < self.start_warm_down:
            return lr
        it = max(1, (iteration + 1) - self.start_warm_down)
        pct = min(1.0, it / (self.warmdown_total_iterations + 1))
        return max(self.min_lr, self.starting_lr - self.warmdown_lr_delta * pct)

    def lr_at(self, lr, step):
        if self.use_warmup:
            lr = self.warmup_dampening(lr, step)
        if self.warmdown_active:
            lr = self.get_warm_down(lr, step)
        self.current_lr = lr
        return lr

    def variance_normalized(self) -> float:
        """variance_normalized of the most recent step (one device read; NaN is what makes the package raise)."""
        return float("nan") if self._scratch is None else float(self._scratch[0].item())

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        entries = []
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = p.grad
                if g.is_sparse:
                    raise RuntimeError("sparse matrix not supported atm")
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or g.dtype != torch.float32 or not g.is_contiguous():
                    raise _lib.IB200Error("FusedRanger21 updates contiguous fp32 CUDA parameters with contiguous fp32 gradients only "
                                          "(no CPU fallback)")
                if p.dim() == 3:
                    raise NotImplementedError("FusedRanger21: 3-d parameters are not supported (the package norms them over dim 1 only)")
                if p.numel() == 0:
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["grad_ma"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["variance_ma"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    if self.lookahead_active:
                        st["lookahead_params"] = p.detach().clone()
                    st["neg_grad_ma"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["max_variance_ma"] = torch.zeros_like(p, memory_format=torch.preserve_format)  # never written by the package either
                st["step"] += 1
                entries.append((group, p, g, st))
        if not entries:
            return loss
        merge = False
        if self.lookahead_active:
            self.lookahead_step += 1
            if self.lookahead_step >= self.lookahead_mergetime:
                self.lookahead_step = 0
                merge = True
        n = len(entries)
        tb = (_lib.Ranger21Tensor * n)()
        for k, (group, p, g, st) in enumerate(entries):
            step = st["step"]
            odd = step % 2 == 1  # the two momentum buffers swap roles every step
            t = tb[k]
            t.param, t.grad = p.data_ptr(), g.data_ptr()
            t.grad_ma = (st["grad_ma"] if odd else st["neg_grad_ma"]).data_ptr()
            t.neg_grad_ma = (st["neg_grad_ma"] if odd else st["grad_ma"]).data_ptr()
            t.variance_ma = st["variance_ma"].data_ptr()
            t.lookahead = st["lookahead_params"].data_ptr() if self.lookahead_active else None
            t.rows = p.shape[0] if p.dim() > 1 else 1
            t.cols = p.numel() // t.rows
            t.multi_dim = 1 if p.dim() > 1 else 0
            t.step = step
            t.lr = self.lr_at(group["lr"], step)
        # one group's betas / eps / weight_decay drive a launch set: the reference builds a single group (self.parameters())
        groups = {id(e[0]) for e in entries}
        if len(groups) != 1:
            raise NotImplementedError("FusedRanger21 steps one param_group (the reference passes self.parameters())")
        group = entries[0][0]
        dev = entries[0][1].device
        need = (int(lib().ib200_ranger21_scratch_bytes(n, tb)) + 7) // 8  # doubles: vn, 1/vn, arrival counter, per-CTA / per-row sums, row norms
        if self._scratch is None or self._scratch.numel() < need or self._scratch.device != dev:
            self._scratch = torch.zeros(max(64, need), dtype=torch.float64, device=dev)
        hyper = _lib.Ranger21Hyper(float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]), float(group["weight_decay"]),
                                   float(self.agc_clip_val), float(self.agc_eps), float(self.normloss_factor), float(self.beta_softplus),
                                   float(self.pnm_momentum_factor), float(self.lookahead_alpha), int(bool(self.agc_active)),
                                   int(bool(self.use_gc)), int(bool(self.use_gcnorm)), int(bool(self.normloss_active)),
                                   int(bool(self.softplus)), int(merge))
        check(lib().ib200_ranger21_step(n, tb, C.byref(hyper), self._scratch.data_ptr(), torch.cuda.current_stream().cuda_stream),
              "ib200_ranger21_step")
        return loss
