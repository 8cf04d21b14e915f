"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.pt from the UNMODIFIED reference (build container only).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden

Every file holds the inputs (tokens, labels, explicit masks, parameters under canonical names) and the outputs of the
reference's own `TripletE2ENet.step(batch, "train")` + `.backward()` and eval-mode `forward`, produced by running
/root/reference/intrepppid/{utils/weightdrop,utils/embedding_do,encoders/awd_lstm,classifier/head/mlp,e2e/e2e_triplet}.py
through `oracle/ref_shim.py` with the 14 RNG draws of a step replaced by the stored masks.
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim, restatement as R  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: dict(E, L, bi_reduce, use_projection, beta, B, T, rates, seeds)
    "train_last_E64": dict(E=64, L=2, bi="last", proj=False, beta=2.0, B=8, T=120, V=250, p_emb=0.3, p_rnn=0.3, p_do=0.3),
    "train_mean_proj_E32": dict(E=32, L=2, bi="mean", proj=True, beta=4.0, B=6, T=64, V=250, p_emb=0.3, p_rnn=0.3, p_do=0.3),
    "train_max_L3_E32": dict(E=32, L=3, bi="max", proj=False, beta=2.0, B=5, T=48, V=100, p_emb=0.2, p_rnn=0.5, p_do=0.1),
    "train_nodrop_E32": dict(E=32, L=2, bi="last", proj=False, beta=2.0, B=4, T=33, V=250, p_emb=0.0, p_rnn=0.0, p_do=0.0),
}


def ragged_batch(B, T, V, seed):
    """Mixed lengths incl. one full-length row, a short row, an ALL-PAD row (Q13) and an interior id 0 (Q1)."""
    g = torch.Generator().manual_seed(seed)
    seqs = []
    for k in range(5):
        s = torch.randint(1, V, (B, T), generator=g)
        lens = torch.randint(3, T + 1, (B,), generator=g)
        lens[k % B] = T
        for b in range(B):
            s[b, int(lens[b]):] = 0
        s[(k + 1) % B, :] = 0                       # all-pad row
        s[(k + 2) % B, 1] = 0                       # interior <unk>/pad id
        seqs.append(s)
    y = torch.randint(0, 2, (B,), generator=g)
    return seqs + [y]


def make_case(name, c):
    torch.manual_seed(0)
    net = ref_shim.build_reference_net(vocab=c["V"], E=c["E"], L=c["L"], bi_reduce=c["bi"], use_projection=c["proj"],
                                       beta=c["beta"], emb_droprate=c["p_emb"], rnn_droprate=c["p_rnn"], do_rate=c["p_do"])
    batch = ragged_batch(c["B"], c["T"], c["V"], seed=11)
    m = R.draw_step_masks(c["B"], c["V"], c["E"], emb_droprate=c["p_emb"], rnn_droprate=c["p_rnn"], do_rate=c["p_do"], seed=5)
    rows = [m.emb_row_keep[g].reshape(-1, 1) for g in range(5)]
    drops = [m.whh_mask[g] for g in range(5)] + [m.fc1_w, m.do1, m.do2, m.fc2_w]

    # capture encoder outputs in call order
    zs = []
    hook = net.encoder.register_forward_hook(lambda mod, inp, out: zs.append(out.detach().clone()))
    net.train()
    with ref_shim.injected_masks(rows, drops):
        loss = net.step(batch, "train")
    loss.backward()
    hook.remove()
    train_z = zs[:5]
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    named = dict(net.named_parameters())
    P = R.params_from_state_dict(sd, c["L"])
    # gradients under canonical names
    key_of = {"emb": "encoder.embedder.weight", "fc_w": "encoder.encoder.fc.weight", "fc_b": "encoder.encoder.fc.bias",
              "fc1_w": "head.classify.fc1.module.weight_raw", "fc1_b": "head.classify.fc1.module.bias",
              "fc2_w": "head.classify.fc2.module.weight_raw", "fc2_b": "head.classify.fc2.module.bias",
              "proj_w": "triplet_projection.1.weight", "proj_b": "triplet_projection.1.bias"}
    grads = {}
    for n in P:
        if n in key_of:
            k = key_of[n]
        else:
            k = "encoder.encoder.rnn." + (n + "_raw" if n == "weight_hh_l0" else n)
        g = named[k].grad
        grads[n] = None if g is None else g.detach().clone()
    dead = [k for k, p in named.items() if p.grad is None]

    # recompute the pieces the Lightning step only logs, using the restatement's formulas on the REFERENCE's z
    # (they are cross-checked against the reference's returned `loss` below)
    za, zp, zn = train_z[0], train_z[1], train_z[2]
    with torch.no_grad():
        if c["proj"]:
            za, zp, zn = (net.triplet_projection(t) for t in (za, zp, zn))
        trip = net.triplet_criterion(za, zp, zn)
        with ref_shim.injected_masks([], [m.fc1_w, m.do1, m.do2, m.fc2_w]):
            y_hat = net.head(train_z[3], train_z[4]).squeeze(1)
        cls = net.classifier_criterion(y_hat, batch[5].float())
        recomposed = (1 - 1 / c["beta"]) * cls + (1 / c["beta"]) * trip
    assert abs(float(recomposed) - float(loss)) < 1e-6, (float(recomposed), float(loss))

    # eval-mode forward on (p1, p2)
    net.eval()
    zs.clear()
    hook = net.encoder.register_forward_hook(lambda mod, inp, out: zs.append(out.detach().clone()))
    with torch.no_grad():
        logits = net(batch[0], batch[1])
    hook.remove()

    out = dict(
        config=dict(c), tokens=[b.clone() for b in batch[:5]], y=batch[5].clone(),
        masks=dict(emb_row_keep=m.emb_row_keep, whh_mask=m.whh_mask, fc1_w=m.fc1_w, do1=m.do1, do2=m.do2, fc2_w=m.fc2_w),
        params={k: v.clone() for k, v in P.items()},
        state_dict_keys=sorted(sd.keys()), dead_parameter_keys=sorted(dead),
        train=dict(loss=loss.detach().clone(), classifier_loss=cls, triplet_loss=trip, y_hat=y_hat, z=train_z, grads=grads),
        eval=dict(logits=logits.clone(), z1=zs[0], z2=zs[1]),
    )
    path = os.path.join(OUT, name + ".pt")
    torch.save(out, path)
    print(f"{name}: loss={float(loss):.9f} cls={float(cls):.6f} trip={float(trip):.6f}  -> {path} ({os.path.getsize(path)/1e6:.2f} MB)")


def make_infer_rows():
    """`infer from_csv` (cli/infer.py:196-225): the reference's own row loop -- net(a.unsqueeze(0), b.unsqueeze(0)) + sigmoid at
    batch 1 -- over a small ragged protein set (a 1-token protein, equal lengths, an interior <unk>, an unknown id)."""
    torch.manual_seed(7)
    V, E, L, T = 60, 64, 2, 48
    net = ref_shim.build_reference_net(vocab=V, E=E, L=L, bi_reduce="last", use_projection=True).eval()
    g = torch.Generator().manual_seed(8)
    lens = [T, T, 1, 17, 17, 17, 5, 33, 40, 9, 29, 12]
    toks = {}
    for i, n in enumerate(lens):
        row = torch.zeros(T, dtype=torch.long)
        row[:n] = torch.randint(1, V, (n,), generator=g)
        toks[f"P{i:02d}"] = row
    toks["P07"][3] = 0
    names = sorted(toks)
    rows = [(f"itx{i}", names[int(torch.randint(0, len(names), (1,), generator=g))], names[int(torch.randint(0, len(names), (1,), generator=g))])
            for i in range(40)]
    rows[4] = ("itx4", "P01", "NOT_THERE")
    rows[9] = ("itx9", "P02", "P02")
    scored = []
    with torch.no_grad():
        for itx, a, b in rows:
            if a not in toks or b not in toks:
                continue  # the reference prints and continues (:203-213)
            prob = torch.sigmoid(net(toks[a].unsqueeze(0), toks[b].unsqueeze(0)))
            scored.append((itx, float(prob.detach().cpu().numpy().tolist()[0][0])))
        z = {n: net.encoder(toks[n].unsqueeze(0))[0].clone() for n in names}
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    out = dict(config=dict(V=V, E=E, L=L, T=T, bi="last"), params={k: v.clone() for k, v in R.params_from_state_dict(sd, L).items()},
               tokens=toks, rows=rows, scored=scored, z=z)
    path = os.path.join(OUT, "infer_rows.pt")
    torch.save(out, path)
    print(f"infer_rows: {len(scored)} of {len(rows)} rows scored -> {path} ({os.path.getsize(path)/1e6:.2f} MB)")


if __name__ == "__main__":
    if not ref_shim.available():
        raise SystemExit("the reference is not mounted; golden vectors can only be regenerated in the build container")
    os.makedirs(OUT, exist_ok=True)
    for n, c in CASES.items():
        make_case(n, c)
    make_infer_rows()
