"""Ranger21 (the reference's factory-default optimizer, e2e/e2e_triplet.py:200-226) -- SURVEY 8f rank 2.

PARITY UNPINNED: the package is a pinned third-party dependency (requirements.txt:65) absent from the image, so the chain here is
    ib200_ranger21_step (CUDA)  ==  oracle/ranger21_restated.py (published algorithm, plain torch, fp64)   [GPU tests]
    oracle/ranger21_restated.py ==  ranger21.Ranger21                                                     [only where the package is importable]
plus properties the published algorithm fixes whatever the implementation (CPU tests): the schedule's end points, the lookahead
period, the alternation of the two momentum buffers, centralized / unit-std gradients after a step, AGC's norm bound.
"""
import math

import pytest
import torch

from conftest import rel_l2
from oracle import ranger21_restated as RR

SHAPES = [(250, 64), (256, 64), (256, 64), (256,), (256,), (256, 128), (64, 128), (64,), (32, 64), (32,), (1, 32), (1,), (2,), (3,),
          (5, 3, 2, 2)]


def _problem(seed, shapes=SHAPES, steps=12, grad_scale=0.05):
    g = torch.Generator().manual_seed(seed)
    params = [torch.randn(s, generator=g) * 0.3 for s in shapes]
    params[0][0].zero_()  # the padding row of the embedding table
    grads = []
    for _ in range(steps):
        gs = [torch.randn(s, generator=g) * grad_scale for s in shapes]
        gs[0][0].zero_()          # ... never receives a gradient
        gs[3] = gs[3] * 50.0      # a tensor AGC clips on every step
        gs[4].zero_()             # an exactly-zero gradient (the dead top-layer chain, quirk Q16)
        grads.append(gs)
    return params, grads


KW = dict(lr=1e-2, weight_decay=1e-2, num_batches_per_epoch=10, num_epochs=3, use_warmup=True, warmdown_active=True,
          warmdown_start_pct=0.72)


# ---- CPU: the restatement itself ---------------------------------------------------------------------------------------------------
def test_schedule_end_points():
    s = RR.Ranger21Schedule(1e-2, num_batches_per_epoch=100, num_epochs=100)  # 10 000 iterations: warm-up = ceil(2 / (1 - 0.999)) = 2000
    assert s.num_warmup_iters == 2000 and s.start_warm_down == 7200
    assert s.lr_at(1e-2, 1) == pytest.approx(1e-2 / 2000) and s.lr_at(1e-2, 2000) == pytest.approx(1e-2)
    assert s.lr_at(1e-2, 5000) == 1e-2                       # plateau
    assert s.lr_at(1e-2, 7200) < 1e-2 and s.lr_at(1e-2, 10_000) == pytest.approx(3e-5, rel=1e-2)
    assert all(s.lr_at(1e-2, t) >= s.lr_at(1e-2, t + 1) for t in range(7200, 10_000, 97))
    short = RR.Ranger21Schedule(1e-2, num_batches_per_epoch=10, num_epochs=3)  # 30 iterations: 2000 > 45 % -> 22 % of the run
    assert short.num_warmup_iters == 6
    with pytest.raises(ValueError):
        RR.Ranger21Schedule(1e-2)
    off = RR.Ranger21Schedule(1e-2, num_batches_per_epoch=10, num_epochs=3, use_warmup=False, warmdown_active=False)  # "ranger21"
    assert off.lr_at(1e-2, 1) == 1e-2 and off.lr_at(1e-2, 30) == 1e-2


def test_restatement_properties():
    params, grads = _problem(3)
    P = [p.clone().double() for p in params]
    opt = RR.Ranger21Restated(P, **KW)
    before = [p.clone() for p in P]
    for t, gs in enumerate(grads, 1):
        G = [g.clone().double() for g in gs]
        G[1] = None  # a parameter without a gradient is never touched (projection.*, quirk Q10)
        raw3 = G[3].clone()
        opt.step(G)
        assert torch.equal(P[1], before[1]) and 1 not in opt.state
        # the gradients are left centralized (rows of matrices) and at unit std (more than two elements), as the package leaves p.grad
        assert float(G[5].mean(dim=1).abs().max()) < 1e-12 and float(G[5].std()) == pytest.approx(1.0, abs=1e-6)
        assert float(G[7].std()) == pytest.approx(1.0, abs=1e-6) and float(G[7].mean().abs()) > 1e-3  # vectors: not centralized
        assert torch.equal(G[4], torch.zeros_like(G[4]))
        assert abs(float(G[12][0] * gs[12][1] - G[12][1] * gs[12][0])) < 1e-12  # two elements: clipped at most, never normalized
        # AGC bound on the clipped tensor before normalization: direction kept
        assert rel_l2(G[3] / G[3].norm(), raw3 / raw3.norm()) < 1e-9
        # positive-negative momentum: odd steps write grad_ma, even steps neg_grad_ma
        st = opt.state[5]
        assert st["step"] == t and float(st["max_variance_ma"].abs().max()) == 0.0
        if t == 1:
            assert float(st["neg_grad_ma"].abs().max()) == 0.0 and float(st["grad_ma"].abs().max()) > 0
        # lookahead: slow weights equal the parameters exactly on every 5th step
        same = torch.equal(opt.state[5]["lookahead_params"], P[5])
        assert same == (t % 5 == 0)
    assert torch.equal(P[0][0], torch.zeros(64, dtype=torch.float64))  # the padding row stays zero through decay / norm loss / update
    assert all(torch.isfinite(p).all() for p in P)


def test_restatement_fp32_tracks_fp64():
    params, grads = _problem(4)
    P64, P32 = [p.clone().double() for p in params], [p.clone() for p in params]
    o64, o32 = RR.Ranger21Restated(P64, **KW), RR.Ranger21Restated(P32, **KW)
    for gs in grads:
        o64.step([g.clone().double() for g in gs])
        o32.step([g.clone() for g in gs])
    for a, b in zip(P32, P64):
        assert rel_l2(a, b) < 5e-6


def test_restatement_descends():
    torch.manual_seed(0)
    target = torch.randn(64, 32, dtype=torch.float64)
    w = [torch.zeros(64, 32, dtype=torch.float64) + 0.01]
    opt = RR.Ranger21Restated(w, lr=3e-2, weight_decay=1e-2, num_batches_per_epoch=100, num_epochs=3, num_warmup_iterations=10)
    first = float(((w[0] - target) ** 2).mean())
    for _ in range(300):
        opt.step([2 * (w[0] - target) / target.numel()])
    assert float(((w[0] - target) ** 2).mean()) < 0.6 * first


def test_against_the_ranger21_package_when_it_is_importable():
    """THE PIN, for environments that have the reference's dependency installed (not this image: skipped here)."""
    ranger21 = pytest.importorskip("ranger21")
    if getattr(ranger21, "__file__", None) is None:  # oracle/ref_shim.py seeds sys.modules with a stub of the missing dependency
        pytest.skip("ranger21 is the reference shim's stub, not the package")
    params, grads = _problem(5)
    mine = [p.clone() for p in params]
    theirs = [torch.nn.Parameter(p.clone()) for p in params]
    o_mine = RR.Ranger21Restated(mine, **KW)
    o_pkg = ranger21.Ranger21(theirs, **KW)
    for gs in grads:
        for p, g in zip(theirs, gs):
            p.grad = g.clone()
        o_pkg.step()
        o_mine.step([g.clone() for g in gs])
    for a, b in zip(mine, theirs):
        assert rel_l2(a, b) < 1e-5


# ---- CPU: the product's host side ----------------------------------------------------------------------------------------------------
def test_fused_ranger21_host_contract():
    import intrepppid_b200 as ib
    from intrepppid_b200 import _lib, build

    build.build()  # (no-op when libib200.so is current)

    w = [torch.nn.Parameter(torch.randn(4, 4))]
    with pytest.raises(ValueError):
        ib.FusedRanger21(w, lr=1e-2)                                    # the package needs the run length too
    for bad in (dict(use_madgrad=True), dict(use_adabelief=True), dict(use_cheb=True), dict(momentum_type="x"), dict(decay_type="x")):
        with pytest.raises(NotImplementedError):
            ib.FusedRanger21(w, lr=1e-2, num_batches_per_epoch=10, num_epochs=3, **bad)
    opt = ib.FusedRanger21(w, **KW)
    ref = RR.Ranger21Schedule(1e-2, num_batches_per_epoch=10, num_epochs=3)
    assert opt.num_warmup_iters == ref.num_warmup_iters and opt.start_warm_down == ref.start_warm_down
    for t in range(1, 40):
        assert opt.lr_at(1e-2, t) == ref.lr_at(1e-2, t)
    assert opt.defaults["weight_decay"] == 1e-2 and opt.defaults["betas"] == (0.9, 0.999) and opt.defaults["eps"] == 1e-8
    opt.step()                                                          # no gradients: nothing to do, nothing launched
    w[0].grad = torch.zeros(4, 4)
    with pytest.raises(_lib.IB200Error):
        opt.step()                                                      # CPU tensors: no fallback
    # the factory default lands on it
    net = ib.intrepppid_network(10, num_epochs=3)
    o = net.configure_optimizers()
    assert isinstance(o, ib.FusedRanger21) and o.use_warmup and o.warmdown_active and o.defaults["weight_decay"] == 1e-2
    net = ib.intrepppid_network(10, num_epochs=3, optimizer_type="ranger21")
    o = net.configure_optimizers()
    assert isinstance(o, ib.FusedRanger21) and not o.use_warmup and not o.warmdown_active


def test_ranger21_abi_rejects_bad_arguments_on_the_host():
    import ctypes as C

    from intrepppid_b200 import _lib, build
    from intrepppid_b200._lib import Ranger21Hyper, Ranger21Tensor

    build.build()
    lib = _lib.lib()
    h = Ranger21Hyper(0.9, 0.999, 1e-8, 1e-2, 1e-2, 1e-3, 1e-4, 50.0, 1.0, 0.5, 1, 1, 1, 1, 1, 0)
    assert lib.ib200_ranger21_step(0, None, C.byref(h), None, None) == 0
    assert lib.ib200_ranger21_step(-1, None, C.byref(h), None, None) == -2
    assert lib.ib200_ranger21_step(1, None, C.byref(h), None, None) == -1
    tb = (Ranger21Tensor * 1)()
    assert lib.ib200_ranger21_step(1, tb, C.byref(h), 8, None) == -1 and b"null tensor" in lib.ib200_last_error()
    t = tb[0]
    t.param = t.grad = t.grad_ma = t.neg_grad_ma = t.variance_ma = 8
    t.rows, t.cols, t.step, t.lr = 1, 4, 0, 1e-2
    assert lib.ib200_ranger21_step(1, tb, C.byref(h), 8, None) == -2 and b"step counts from 1" in lib.ib200_last_error()
    t.step, t.cols = 1, 0
    assert lib.ib200_ranger21_step(1, tb, C.byref(h), 8, None) == -2
    t.cols = 4
    h.lookahead_merge = 1                                               # merging needs the slow weights
    assert lib.ib200_ranger21_step(1, tb, C.byref(h), 8, None) == -1
    h.lookahead_merge, h.beta2 = 0, 1.0
    assert lib.ib200_ranger21_step(1, tb, C.byref(h), 8, None) == -2


# ---- GPU: the kernels against the restatement --------------------------------------------------------------------------------------
def _run_pair(kw, seed, shapes=SHAPES, steps=12, none_at=1, extra=None):
    import intrepppid_b200 as ib

    params, grads = _problem(seed, shapes, steps)
    P64, P32 = [p.clone().double() for p in params], [p.clone() for p in params]
    mine = [torch.nn.Parameter(p.clone().cuda()) for p in params]
    okw = dict(kw)
    okw.update(extra or {})
    o64, o32 = RR.Ranger21Restated(P64, **okw), RR.Ranger21Restated(P32, **okw)
    om = ib.FusedRanger21(mine, **okw)
    G64 = None
    for gs in grads:
        G64 = [None if k == none_at else g.clone().double() for k, g in enumerate(gs)]
        for k, g in enumerate(gs):
            mine[k].grad = None if k == none_at else g.clone().cuda()
        o64.step(G64)
        o32.step([None if k == none_at else g.clone() for k, g in enumerate(gs)])
        om.step()
    torch.cuda.synchronize()
    return params, P64, P32, mine, o64, om, G64


BIG = [(300, 200), (64,), (1000, 64), (64, 512), (7,), (70000,), (3, 40000)]  # past the 48 K-element staging buffer / rows wider than 128


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["ranger21_xx", "ranger21", "no_extras", "two_tables", "big_tensors"])
def test_fused_ranger21_matches_the_restatement(variant):
    from intrepppid_b200 import _lib

    kw, shapes, extra = dict(KW), SHAPES, None
    if variant == "ranger21":
        kw.update(use_warmup=False, warmdown_active=False)
    elif variant == "no_extras":
        extra = dict(use_gc=False, use_gcnorm=False, normloss_active=False, use_adaptive_gradient_clipping=False, softplus=False,
                     lookahead_active=False)
    elif variant == "two_tables":
        shapes = SHAPES * 2   # 30 tensors, 29 with a gradient: crosses the 24-tensor kernel-parameter table
    elif variant == "big_tensors":
        shapes = BIG
    l0 = _lib.launch_count()
    params, P64, P32, mine, o64, om, G64 = _run_pair(kw, 21, shapes, extra=extra)
    n_live = len(shapes) - 1
    assert _lib.launch_count() - l0 == 12 * 3 * math.ceil(n_live / 24)
    for k in range(len(shapes)):
        got = mine[k].detach().cpu()
        if k == 1:
            assert torch.equal(got, params[k]) and mine[k] not in om.state
            continue
        e64, e32 = rel_l2(got, P64[k]), rel_l2(P32[k], P64[k])
        assert e64 < 2e-5 and e64 < 4 * e32 + 2e-6, (variant, k, shapes[k], e64, e32)
        st, sr = om.state[mine[k]], o64.state[k]
        assert st["step"] == sr["step"] == 12
        for name in ("grad_ma", "neg_grad_ma", "variance_ma") + (("lookahead_params",) if "lookahead_params" in sr else ()):
            assert rel_l2(st[name], sr[name]) < 2e-5, (variant, k, name)
        assert float(st["max_variance_ma"].abs().max()) == 0.0
        assert rel_l2(mine[k].grad, G64[k]) < 2e-5 or float(G64[k].abs().max()) == 0.0   # p.grad is rewritten in place like the package's
    assert torch.equal(mine[0].detach()[0].cpu(), torch.zeros(shapes[0][1]))
    assert om.variance_normalized() > 0


@pytest.mark.gpu
def test_default_network_trains_with_fused_ranger21():
    """The factory-default module (optimizer_type='ranger21_xx') takes optimisation steps on the kernels and the loss goes down."""
    import intrepppid_b200 as ib
    from oracle import restatement as R

    torch.manual_seed(0)
    net = ib.intrepppid_network(50, num_epochs=1, embedding_droprate=0.0, rnn_dropout_rate=0.0, do_rate=0.0).cuda().train()
    opt = net.configure_optimizers()
    assert isinstance(opt, ib.FusedRanger21)
    batch = tuple(t.cuda() for t in R.synthetic_batch(B=8, T=48, V=250, seed=3))
    losses = []
    for _ in range(30):
        opt.zero_grad(set_to_none=True)
        loss = net.step(batch, "train")
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert all(math.isfinite(x) for x in losses) and min(losses[-5:]) < losses[0], losses
    assert net.encoder.embedder.weight.detach()[0].abs().max().item() == 0.0  # padding row untouched
    assert all(p not in opt.state for n, p in net.named_parameters() if "projection" in n)


@pytest.mark.gpu
def test_fused_ranger21_state_dict_round_trip_continues_the_run():
    """An optimizer checkpoint (state_dict -> a fresh FusedRanger21 -> load_state_dict) continues bit-for-bit: the per-parameter
    state carries the package's keys; the lookahead counter is an attribute (as in the package) and is restored by the caller."""
    import copy

    import intrepppid_b200 as ib

    params, grads = _problem(31, steps=9)
    a = [torch.nn.Parameter(p.clone().cuda()) for p in params]
    oa = ib.FusedRanger21(a, **KW)
    for gs in grads[:4]:
        for p, g in zip(a, gs):
            p.grad = g.clone().cuda()
        oa.step()
    sd = copy.deepcopy(oa.state_dict())
    assert set(sd["state"][0]) == {"step", "grad_ma", "variance_ma", "lookahead_params", "neg_grad_ma", "max_variance_ma"}
    assert sd["state"][0]["step"] == 4 and sd["param_groups"][0]["weight_decay"] == 1e-2
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    ob = ib.FusedRanger21(b, **KW)
    ob.load_state_dict(sd)
    ob.lookahead_step = oa.lookahead_step
    for gs in grads[4:]:
        for pa, pb, g in zip(a, b, gs):
            pa.grad, pb.grad = g.clone().cuda(), g.clone().cuda()
        oa.step()
        ob.step()
    torch.cuda.synchronize()
    for pa, pb in zip(a, b):
        assert torch.equal(pa.detach(), pb.detach())
    assert ob.state[b[0]]["step"] == 9


def test_second_centralization_and_normalization_are_rounding_level():
    """The kernels apply the package's SECOND centralization + normalization of the gradient analytically (csrc/ranger21.cu, elem1):
    on an already centralized, unit-std tensor the second centralization subtracts rounding residue and the second std follows from
    the first, std(g1) = std(g) / (std(g) + 1e-8).  Checked here in fp32 against the literal two-pass sequence."""
    g = torch.Generator().manual_seed(7)
    for shape in [(256, 64), (250, 64), (256,), (3,), (64, 128), (5, 3, 2, 2)]:
        x = torch.randn(shape, generator=g) * 0.03
        lit = RR.normalize_(RR.centralize_(RR.normalize_(RR.centralize_(x.clone()))))        # what the package leaves in p.grad
        gc = RR.centralize_(x.clone())
        sd = gc.double().std().float()
        inv1 = 1.0 / (sd + 1e-8)
        inv2 = 1.0 / (sd * inv1 + 1e-8)
        mine = (gc * inv1) * inv2
        assert rel_l2(mine, lit) < 5e-7, (shape, rel_l2(mine, lit))
