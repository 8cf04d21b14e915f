"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch) of the Ranger21 optimizer step as INTREPPPID configures it.

PARITY UNPINNED.  Ranger21 is a third-party dependency of the reference, pinned at
`ranger21 @ git+https://github.com/lessw2020/Ranger21.git@1a96777278cdd14bc11afd865112724386d26a44` (requirements.txt:65;
pyproject.toml:27 `ranger21 = "^0.1.0"`), and its source is NOT in /root/reference, not in the wheelhouse and not fetchable
(no network).  This file restates the published algorithm -- Wright & Demeure, "Ranger21: a synergistic deep learning
optimizer", arXiv:2106.13731, and the step order of the authors' public implementation (`ranger21/ranger21.py`, class Ranger21,
0.1.0) -- for the one configuration the reference constructs (e2e/e2e_triplet.py:200-226):

    Ranger21(params, lr=lr, weight_decay=1e-2, use_warmup=xx, warmdown_active=xx, num_batches_per_epoch=steps_per_epoch,
             num_epochs=num_epochs, warmdown_start_pct=0.72)          # xx = (optimizer_type == "ranger21_xx")

i.e. every other switch at the package default: AdamW core with positive-negative momentum (pnm_momentum_factor 1.0), adaptive
gradient clipping (1e-2, eps 1e-3), gradient centralization + gradient normalization, norm loss (1e-4), stable weight decay,
softplus (beta 50) on the denominator, linear warm-up, linear warm-down to 3e-5, lookahead (k = 5, alpha = 0.5).
`tests/test_ranger21.py::test_against_the_ranger21_package_when_it_is_importable` compares this file with the real package
whenever `import ranger21` works (it does not in the build image) -- that test is the pin to run where the package exists.

Only `tests/` may import this module; the product (`intrepppid_b200.optim.FusedRanger21`) never does.

Step order restated (numbers = phases of `Ranger21.step`):
  1  for every parameter with a gradient:  AGC (unit-wise clip of p.grad, in place) -> state init -> gradient centralization
     (rows of tensors with more than one dimension) -> gradient normalization by the whole-tensor std (more than 2 elements)
     -- both in place on p.grad -> step += 1 -> variance_ma = b2 variance_ma + (1-b2) g^2 -> sum of variance_ma / (1 - b2^step)
  .  variance_normalized = sqrt(sum / number of elements of those parameters)
  2  per parameter: lr through warm-up / warm-down -> p *= 1 - wd lr / variance_normalized (stable weight decay) -> norm loss
     p *= 1 - lr 2 nf (1 - 1/(unit_norm(p) + eps)) -> the two momentum buffers swap roles every step (positive-negative momentum) ->
     variance_ma = max(max_variance_ma, variance_ma) with max_variance_ma never written (all zeros: a no-op on a non-negative
     tensor; kept because it is what the public implementation executes) -> denom = sqrt(variance_ma)/sqrt(1 - b2^step) + eps ->
     the (already centralized and normalized) gradient is centralized and normalized a SECOND time in place -> grad_ma = b1^2
     grad_ma + (1 - b1^2) g -> denom = softplus(denom, beta 50) -> p -= lr/(1 - b1^step) ((1+pf) grad_ma - pf neg_grad_ma) /
     sqrt((1+b2)^2 + b2^2) / denom
  3  lookahead: every 5th call p = 0.5 p + 0.5 slow ; slow = p
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


def unit_norm(x: torch.Tensor) -> torch.Tensor:
    """Euclidean norm over everything but the leading dimension (whole tensor for 0-d / 1-d; dimension 1 only for 2-d / 3-d)."""
    n = x.dim()
    if n <= 1:
        return x.norm(p=2.0)
    if n in (2, 3):
        return x.norm(dim=1, keepdim=True, p=2.0)
    return x.norm(dim=tuple(range(1, n)), keepdim=True, p=2.0)


def agc_(p: torch.Tensor, g: torch.Tensor, clip: float, agc_eps: float) -> None:
    p_norm = unit_norm(p).clamp_(agc_eps)
    g_norm = unit_norm(g)
    max_norm = p_norm * clip
    clipped = g * (max_norm / g_norm.clamp(min=1e-6))
    g.copy_(torch.where(g_norm > max_norm, clipped, g))


def centralize_(g: torch.Tensor) -> torch.Tensor:
    if g.dim() > 1:
        g.add_(-g.mean(dim=tuple(range(1, g.dim())), keepdim=True))
    return g


def normalize_(g: torch.Tensor, epsilon: float = 1e-8) -> torch.Tensor:
    if g.numel() > 2:
        g.div_(g.std() + epsilon)
    return g


class Ranger21Schedule:
    """The learning-rate part (warm-up / warm-down) of the constructor and of `warmup_dampening` / `get_warm_down`."""

    def __init__(self, lr: float, betas=(0.9, 0.999), num_batches_per_epoch: Optional[int] = None, num_epochs: Optional[int] = None,
                 use_warmup: bool = True, num_warmup_iterations: Optional[int] = None, warmup_pct_default: float = 0.22,
                 warmdown_active: bool = True, warmdown_start_pct: float = 0.72, warmdown_min_lr: float = 3e-5):
        self.starting_lr = lr
        self.total_iterations = (num_epochs or 0) * (num_batches_per_epoch or 0)
        if not self.total_iterations:
            raise ValueError("missing total iterations, which is calced from num epochs and num iters per epoch param")
        self.use_warmup = use_warmup
        if num_warmup_iterations is None:
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
            self.start_warm_down = int(warmdown_start_pct * num_epochs * num_batches_per_epoch)
            self.warmdown_total_iterations = self.total_iterations - self.start_warm_down

    def lr_at(self, lr: float, step: int) -> float:
        if self.use_warmup and step <= self.num_warmup_iters:
            lr = lr * min(1.0, step / self.num_warmup_iters)
        if self.warmdown_active and step >= self.start_warm_down:
            it = max(1, (step + 1) - self.start_warm_down)
            pct = min(1.0, it / (self.warmdown_total_iterations + 1))
            lr = max(self.min_lr, self.starting_lr - self.warmdown_lr_delta * pct)
        return lr


class Ranger21Restated:
    """Functional restatement over explicit tensor lists (any floating dtype: the tests run it in fp64 and fp32)."""

    def __init__(self, params: List[torch.Tensor], lr: float, *, weight_decay: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 num_batches_per_epoch: Optional[int] = None, num_epochs: Optional[int] = None, use_warmup: bool = True,
                 num_warmup_iterations: Optional[int] = None, warmdown_active: bool = True, warmdown_start_pct: float = 0.72,
                 warmdown_min_lr: float = 3e-5, lookahead_active: bool = True, lookahead_mergetime: int = 5,
                 lookahead_blending_alpha: float = 0.5, softplus: bool = True, beta_softplus: float = 50.0, use_gc: bool = True,
                 use_gcnorm: bool = True, normloss_active: bool = True, normloss_factor: float = 1e-4,
                 use_adaptive_gradient_clipping: bool = True, agc_clipping_value: float = 1e-2, agc_eps: float = 1e-3,
                 pnm_momentum_factor: float = 1.0):
        self.params = params
        self.lr, self.weight_decay, self.betas, self.eps = lr, weight_decay, betas, eps
        self.sched = Ranger21Schedule(lr, betas, num_batches_per_epoch, num_epochs, use_warmup, num_warmup_iterations, 0.22,
                                      warmdown_active, warmdown_start_pct, warmdown_min_lr)
        self.lookahead_active, self.lookahead_mergetime, self.lookahead_alpha = lookahead_active, lookahead_mergetime, lookahead_blending_alpha
        self.lookahead_step = 0
        self.softplus, self.beta_softplus = softplus, beta_softplus
        self.use_gc, self.use_gcnorm = use_gc, use_gcnorm
        self.normloss_active, self.normloss_factor = normloss_active, normloss_factor
        self.agc_active, self.agc_clip_val, self.agc_eps = use_adaptive_gradient_clipping, agc_clipping_value, agc_eps
        self.pnm_factor = pnm_momentum_factor
        self.state: Dict[int, Dict[str, object]] = {}

    @torch.no_grad()
    def step(self, grads: List[Optional[torch.Tensor]]) -> None:
        """`grads[k]` is the gradient of `params[k]` (None = no gradient: skipped) and is modified in place exactly as the
        implementation modifies p.grad (clipped, centralized, normalized -- twice)."""
        b1, b2 = self.betas
        param_size, variance_sum = 0, 0.0
        live = [(k, p, g) for k, (p, g) in enumerate(zip(self.params, grads)) if g is not None]
        for k, p, g in live:
            param_size += p.numel()
            if self.agc_active:
                agc_(p, g, self.agc_clip_val, self.agc_eps)
            st = self.state.setdefault(k, {})
            if not st:
                st["step"] = 0
                st["grad_ma"] = torch.zeros_like(p)
                st["variance_ma"] = torch.zeros_like(p)
                if self.lookahead_active:
                    st["lookahead_params"] = p.clone()
                st["neg_grad_ma"] = torch.zeros_like(p)
                st["max_variance_ma"] = torch.zeros_like(p)
            if self.use_gc:
                centralize_(g)
            if self.use_gcnorm:
                normalize_(g)
            st["step"] += 1
            v = st["variance_ma"]
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            variance_sum = variance_sum + (v / (1 - b2 ** st["step"])).sum()
        if not live:
            return
        variance_normalized = math.sqrt(float(variance_sum) / param_size)
        if math.isnan(variance_normalized):
            raise RuntimeError("hit nan for variance_normalized")

        for k, p, g in live:
            st = self.state[k]
            step = st["step"]
            lr = self.sched.lr_at(self.lr, step)
            if self.weight_decay:
                p.mul_(1 - self.weight_decay * lr / variance_normalized)
            if self.normloss_active:
                unorm = unit_norm(p)
                correction = 2 * self.normloss_factor * (1 - torch.div(1, unorm + self.eps))
                p.mul_(1 - lr * correction)
            if step % 2 == 1:
                grad_ma, neg_grad_ma = st["grad_ma"], st["neg_grad_ma"]
            else:
                grad_ma, neg_grad_ma = st["neg_grad_ma"], st["grad_ma"]
            bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
            v = st["variance_ma"]
            torch.max(st["max_variance_ma"], v, out=v)
            denom = (v.sqrt() / math.sqrt(bc2)).add_(self.eps)
            if self.use_gc:
                centralize_(g)
            if self.use_gcnorm:
                normalize_(g)
            grad_ma.mul_(b1 ** 2).add_(g, alpha=1 - b1 ** 2)
            noise_norm = math.sqrt((1 + b2) ** 2 + b2 ** 2)
            if self.softplus:
                denom = F.softplus(denom, beta=self.beta_softplus)
            pnmomentum = grad_ma.mul(1 + self.pnm_factor).add(neg_grad_ma, alpha=-self.pnm_factor).mul(1 / noise_norm)
            p.addcdiv_(pnmomentum, denom, value=-(lr / bc1))

        if self.lookahead_active:
            self.lookahead_step += 1
            if self.lookahead_step >= self.lookahead_mergetime:
                self.lookahead_step = 0
                for k, p, g in live:
                    slow = self.state[k]["lookahead_params"]
                    p.mul_(self.lookahead_alpha).add_(slow, alpha=1.0 - self.lookahead_alpha)
                    slow.copy_(p)
