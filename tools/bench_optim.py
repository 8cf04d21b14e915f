#!/usr/bin/env python
"""The optimizer step next to the hot path on the headline network's parameters: FusedAdamW and FusedRanger21
(the reference's factory default, e2e_triplet.py:212-224) timed alone with CUDA events, and the full headline training step with
each.  For scale, `eager_torch_ranger21` times the same Ranger21 arithmetic written with one torch call per tensor operation (the
way the third-party package executes it: ~40 small kernels per tensor and a host sync per step).  One JSON line."""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import intrepppid_b200 as ib  # noqa: E402
from intrepppid_b200 import _lib  # noqa: E402

B, T, V = 80, 1500, 250
KW = dict(lr=1e-2, weight_decay=1e-2, use_warmup=True, warmdown_active=True, num_batches_per_epoch=1000, num_epochs=100,
          warmdown_start_pct=0.72)


def timed(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def eager_ranger21_step(params, state, t, lr=1e-2, wd=1e-2, b1=0.9, b2=0.999, eps=1e-8):
    """Per-tensor torch calls in the package's order (timing yardstick only; the parity checks live in tests/test_ranger21.py)."""
    def unorm(x):
        return x.norm(p=2.0) if x.dim() <= 1 else x.norm(dim=1, keepdim=True, p=2.0)

    vsum, n = 0.0, 0
    for p in params:
        g = p.grad
        n += p.numel()
        pn = unorm(p).clamp_(1e-3)
        gn = unorm(g)
        mx = pn * 1e-2
        g.copy_(torch.where(gn > mx, g * (mx / gn.clamp(min=1e-6)), g))
        st = state.setdefault(p, {})
        if not st:
            st.update(m=torch.zeros_like(p), mn=torch.zeros_like(p), v=torch.zeros_like(p), slow=p.detach().clone())
        if g.dim() > 1:
            g.add_(-g.mean(dim=1, keepdim=True))
        if g.numel() > 2:
            g.div_(g.std() + 1e-8)
        st["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
        vsum = vsum + (st["v"] / (1 - b2 ** t)).sum()
    vn = math.sqrt(vsum / n)  # the package's host sync
    for p in params:
        g, st = p.grad, state[p]
        p.mul_(1 - wd * lr / vn)
        p.mul_(1 - lr * (2e-4 * (1 - torch.div(1, unorm(p) + eps))))
        m, mn = (st["m"], st["mn"]) if t % 2 == 1 else (st["mn"], st["m"])
        torch.max(torch.zeros_like(st["v"]), st["v"], out=st["v"])
        denom = (st["v"].sqrt() / math.sqrt(1 - b2 ** t)).add_(eps)
        if g.dim() > 1:
            g.add_(-g.mean(dim=1, keepdim=True))
        if g.numel() > 2:
            g.div_(g.std() + 1e-8)
        m.mul_(b1 ** 2).add_(g, alpha=1 - b1 ** 2)
        denom = F.softplus(denom, beta=50)
        pn = m.mul(2.0).add(mn, alpha=-1.0).mul(1 / math.sqrt((1 + b2) ** 2 + b2 ** 2))
        p.addcdiv_(pn, denom, value=-lr / (1 - b1 ** t))
    if t % 5 == 0:
        for p in params:
            p.mul_(0.5).add_(state[p]["slow"], alpha=0.5)
            state[p]["slow"].copy_(p)


def main():
    torch.manual_seed(0)
    out = {"what": "optimizer step on the headline network (23 tensors with gradients, 0.87 MB of parameters)"}
    g = torch.Generator().manual_seed(1)
    data = [torch.randint(1, V, (B, T), generator=g).cuda() for _ in range(5)] + [torch.randint(0, 2, (B,), generator=g).cuda()]
    for name in ("adamw", "ranger21_xx"):
        net = ib.intrepppid_network(1000, optimizer_type=name).cuda().train()
        opt = net.configure_optimizers()
        net.step(data, "train").backward()
        live = [p for p in net.parameters() if p.grad is not None]
        grads = [p.grad.clone() for p in live]

        def restore():  # Ranger21 rewrites p.grad in place: restore it (one foreach copy, timed alone and subtracted)
            torch._foreach_copy_([p.grad for p in live], grads)

        def opt_with_restore():
            restore()
            opt.step()

        l0 = _lib.launch_count()
        opt.step()
        launches = _lib.launch_count() - l0
        ms_restore = timed(restore)
        out[name] = {"host_enqueue_us": round((timed(opt_with_restore) - ms_restore) * 1e3, 2), "launches_per_step": launches,
                     "class": type(opt).__name__}
        _lib.timing_enable(True)  # CUDA events around the library's launches: the device time of the step
        for _ in range(10):
            opt_with_restore()
        fam = _lib.timing_read()
        _lib.timing_enable(False)
        key = "adamw" if name == "adamw" else "ranger21"
        out[name]["device_us"] = round(fam[key][0] / fam[key][1] * 1e3, 2)

        def step():
            opt.zero_grad(set_to_none=True)
            net.step(data, "train").backward()
            opt.step()

        out[name]["train_step_ms"] = round(timed(step, n=10, warm=3), 4)
        if name == "ranger21_xx":
            out[name]["variance_normalized"] = opt.variance_normalized()
            state, k = {}, [0]

            def eager():
                restore()
                k[0] += 1
                with torch.no_grad():
                    eager_ranger21_step(live, state, k[0])

            out["eager_torch_ranger21"] = {"step_us": round((timed(eager) - ms_restore) * 1e3, 2),
                                           "note": "same arithmetic as per-tensor torch calls + one host sync per step (how the "
                                                   "third-party package executes); timing yardstick, not a parity check"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
