"""AdamW on the multi-tensor sm_100a kernel (`ib200_adamw_step`) -- the optimizer step that follows the hot path
(reference: e2e/e2e_triplet.py:231-255, `AdamW(self.parameters(), lr=self.lr)`).

Same constructor arguments, defaults, param_groups and state_dict layout (`step`, `exp_avg`, `exp_avg_sq` per parameter) as
torch.optim.AdamW, so learning-rate schedulers (OneCycleLR, CosineAnnealingWarmRestarts -- e2e_triplet.py:239-253) and optimizer
checkpoints interchange with the reference's.  All parameters of a group are updated by ONE kernel launch (per 32 tensors).
CUDA fp32 parameters only; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

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
