"""Stage-by-stage GPU-vs-oracle diagnostic (development aid; run on the GPU box from the repo root: python tools/gpu_diag.py)."""
import sys, time, traceback
import torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from helpers import run_product_step, run_oracle_step
from oracle import restatement as R
from conftest import rel_l2


def case(name, *, E, L, bi, B, T, V=60, proj=False, beta=2.0, p=0.3, precision="fp32", training=True, padded=True, seed=3):
    P = R.init_params(vocab=V, E=E, L=L, use_projection=proj, seed=seed)
    batch = list(R.synthetic_batch(B, T, V, seed=seed))
    if padded:
        for k, s in enumerate(batch[:5]):
            s[1, T // 2:] = 0
            s[2 % B, :] = 0
            s[3 % B, 4 % T] = 0
            s[(4 + k) % B, T - 3:] = 0
    masks = R.draw_step_masks(B, V, E, emb_droprate=p, rnn_droprate=p, do_rate=p, seed=seed + 1) if training and p > 0 else R.StepMasks()
    kw = dict(L=L, bi=bi, beta=beta, use_projection=proj, p_emb=p)
    t0 = time.time()
    ref = run_oracle_step(P, batch, masks, training=training, **kw)
    t1 = time.time()
    try:
        got = run_product_step(P, batch, masks, p_rnn=p, p_do=p, precision=precision, training=training, **kw)
    except Exception:
        print(f"[{name}] PRODUCT RAISED:\n{traceback.format_exc()}")
        return
    print(f"[{name}] E={E} L={L} bi={bi} B={B} T={T} prec={precision} train={training}  (oracle {t1-t0:.1f}s)")
    print(f"   lengths ref {ref['lengths'].tolist()} got {got['lengths'].cpu().tolist()}  exact={torch.equal(ref['lengths'], got['lengths'].cpu())}")
    for g in range(5):
        print(f"   z[{g}] rel {rel_l2(got['z'][g], ref['z'][g]):.3e}", end="")
    print()
    for k in ("loss", "classifier_loss", "triplet_loss"):
        print(f"   {k}: ref {float(ref[k]):.7f} got {float(got[k]):.7f}")
    print(f"   y_hat rel {rel_l2(got['y_hat'], ref['y_hat']):.3e}")
    if training:
        for n, g in ref["grads"].items():
            gg = got["grads"][n]
            if gg is None:
                print(f"   grad {n:24s} MISSING")
                continue
            nrm = float(g.norm())
            print(f"   grad {n:24s} rel {rel_l2(gg, g) if nrm > 0 else float(gg.abs().max()):.3e}  |ref| {nrm:.3e} |got| {float(gg.norm()):.3e}")
    sys.stdout.flush()


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    sel = sys.argv[1:] 
    cases = {
        "eval_L1_last_H32": dict(E=32, L=1, bi="last", B=5, T=20, training=False),
        "eval_L1_mean_H64": dict(E=64, L=1, bi="mean", B=9, T=37, training=False),
        "eval_L2_last_H64": dict(E=64, L=2, bi="last", B=9, T=37, training=False),
        "eval_L2_max_H32": dict(E=32, L=2, bi="max", B=11, T=70, training=False),
        "train_L1_mean_H32_nodrop": dict(E=32, L=1, bi="mean", B=5, T=20, p=0.0),
        "train_L1_mean_H64": dict(E=64, L=1, bi="mean", B=9, T=37),
        "train_L2_last_H64": dict(E=64, L=2, bi="last", B=9, T=37),
        "train_L2_mean_H32_proj": dict(E=32, L=2, bi="mean", B=11, T=70, proj=True, beta=4.0),
        "train_L3_max_H64": dict(E=64, L=3, bi="max", B=17, T=150),
        "train_L2_last_H64_bf16": dict(E=64, L=2, bi="last", B=9, T=37, precision="bf16"),
        "train_L2_last_H64_T600": dict(E=64, L=2, bi="last", B=16, T=600, V=250, padded=False),
    }
    for n, c in cases.items():
        if sel and n not in sel:
            continue
        case(n, **c)
