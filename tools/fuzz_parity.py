"""Randomised parity sweep on the GPU box: random shapes / options against the fp64 CPU oracle (same explicit masks).
    python tools/fuzz_parity.py [n_cases] [seed]"""
import os, random, sys, time, traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import rel_l2
from helpers import run_oracle_step, run_product_step
from intrepppid_b200 import ops
from oracle import restatement as R

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
t0 = time.time()
for case in range(n_cases):
    E = rng.choice([32, 64, 64, 96, 128, 128, 192, 256, 256])
    L = rng.choice([1, 2, 2, 3])
    bi = rng.choice(["last", "last", "mean", "max"])
    B = rng.choice([1, 2, 3, 5, 8, 9, 13, 17, 24, 31, 33, 41, 70])
    T = rng.choice([1, 2, 7, 33, 64, 65, 100, 130, 257])
    V = rng.choice([11, 60, 250, 256, 257, 400])
    proj = rng.random() < 0.3
    p = rng.choice([0.0, 0.3, 0.5])
    prec = rng.choice(["fp32", "fp32", "bf16"])
    beta = rng.choice([2.0, 4.0])
    seed = rng.randrange(10 ** 6)
    desc = f"E={E} L={L} bi={bi} B={B} T={T} V={V} proj={proj} p={p} {prec} beta={beta} seed={seed}"
    try:
        P = R.init_params(vocab=V, E=E, L=L, seed=seed, use_projection=proj)
        batch = list(R.synthetic_batch(B, T, V, seed=seed + 1, padded=rng.random() < 0.7))
        masks = R.draw_step_masks(B, V, E, emb_droprate=p, rnn_droprate=p, do_rate=p, seed=seed + 2) if p > 0 else R.StepMasks()
        kw = dict(L=L, bi=bi, beta=beta, use_projection=proj, p_emb=p)
        try:
            ref = run_oracle_step(P, batch, masks, **kw)
        except RuntimeError as e:  # all-pad batch: the product must raise too
            try:
                run_product_step(P, batch, masks, p_rnn=p, p_do=p, precision=prec, **kw)
                ops.check_pending(sync=True)  # the length checks are a device-side status word: raised lazily, at the latest here
                print(f"[{case}] FAIL (oracle raised {e!r}, product did not): {desc}")
                bad += 1
            except RuntimeError:
                print(f"[{case}] ok (both raise): {desc}")
            continue
        got = run_product_step(P, batch, masks, p_rnn=p, p_do=p, precision=prec, **kw)
        ops.check_pending(sync=True)
        tol = 1e-4 if prec == "fp32" else 2e-2
        worst, where = 0.0, ""
        assert torch.equal(got["lengths"].cpu(), ref["lengths"]), "lengths"
        for g in range(5):
            e = rel_l2(got["z"][g], ref["z"][g])
            if e > worst: worst, where = e, f"z[{g}]"
        e = abs(float(got["loss"]) - float(ref["loss"])) / max(1.0, abs(float(ref["loss"])))
        if e > worst: worst, where = e, "loss"
        if not (bi == "max" and prec == "bf16"):  # argmax flips under bf16 noise: no meaningful gradient gate
            for n, gr in ref["grads"].items():
                gg = got["grads"][n]
                if float(gr.abs().max()) == 0.0:
                    # dead chains are exactly zero; the projection bias cancels analytically (a-p, a-n): only rounding noise is left
                    assert float(gg.abs().max()) <= (1e-6 if n == "proj_b" else 0.0), f"{n} must be exactly zero"
                elif float(torch.linalg.norm(gr.double())) >= 1e-6:
                    e = rel_l2(gg, gr)
                    if e > worst: worst, where = e, n
        ok = worst < tol
        print(f"[{case}] {'ok  ' if ok else 'FAIL'} worst {worst:.2e} at {where}: {desc}", flush=True)
        bad += 0 if ok else 1
    except Exception as e:  # noqa: BLE001
        bad += 1
        print(f"[{case}] ERROR {type(e).__name__}: {e}: {desc}")
        traceback.print_exc(limit=2)
print(f"{n_cases} cases, {bad} failures, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
