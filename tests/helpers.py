"""Shared test plumbing: canonical-parameter <-> module mapping, running the CUDA product and the CPU oracle side by side."""
from __future__ import annotations

import torch

from oracle import restatement as R

HEAD_KEYS = {"fc1_w": "head.classify.fc1.module.weight_raw", "fc1_b": "head.classify.fc1.module.bias",
             "fc2_w": "head.classify.fc2.module.weight_raw", "fc2_b": "head.classify.fc2.module.bias",
             "proj_w": "triplet_projection.1.weight", "proj_b": "triplet_projection.1.bias",
             "emb": "encoder.embedder.weight", "fc_w": "encoder.encoder.fc.weight", "fc_b": "encoder.encoder.fc.bias"}


def module_key(name: str) -> str:
    if name in HEAD_KEYS:
        return HEAD_KEYS[name]
    return "encoder.encoder.rnn." + (name + "_raw" if name == "weight_hh_l0" else name)


def build_product(P, *, L, bi, beta=2.0, use_projection=False, p_emb=0.3, p_rnn=0.3, p_do=0.3, precision="fp32", device="cuda"):
    """intrepppid_b200 network holding the canonical parameters P (the dead `projection.*` tensors keep their init)."""
    import intrepppid_b200 as ib

    V, E = P["emb"].shape
    net = ib.intrepppid_network(1, vocab_size=V, embedding_size=E, rnn_num_layers=L, rnn_dropout_rate=p_rnn, bi_reduce=bi,
                                embedding_droprate=p_emb, do_rate=p_do, beta_classifier=beta, use_projection=use_projection,
                                optimizer_type="adamw", precision=precision)
    sd = net.state_dict()
    for n, v in P.items():
        k = module_key(n)
        sd[k] = v.clone()
        if k.startswith("encoder.encoder.rnn."):
            sd[k.replace("encoder.encoder.rnn.", "encoder.encoder.rnn_dp.module.")] = v.clone()
    net.load_state_dict(sd, strict=True)
    return net.to(device)


def product_masks(masks: R.StepMasks, p_emb: float, device="cuda"):
    from intrepppid_b200 import StepMasks

    def dev(t):
        return None if t is None else t.to(device)

    ers = None
    if masks.emb_row_keep is not None and p_emb:
        ers = dev(masks.emb_row_keep / (1.0 - p_emb))
    return StepMasks(ers, dev(masks.whh_mask), (dev(masks.fc1_w), dev(masks.do1), dev(masks.do2), dev(masks.fc2_w)))


def product_grads(net, names):
    named = dict(net.named_parameters())
    return {n: (None if named[module_key(n)].grad is None else named[module_key(n)].grad.detach().cpu()) for n in names}


def run_product_step(P, batch, masks, *, L, bi, beta, use_projection, p_emb, p_rnn, p_do, precision="fp32", training=True):
    net = build_product(P, L=L, bi=bi, beta=beta, use_projection=use_projection, p_emb=p_emb, p_rnn=p_rnn, p_do=p_do,
                        precision=precision)
    net.train(training)
    cuda_batch = [t.cuda() for t in batch]
    loss = net.step(cuda_batch, "train", masks=product_masks(masks, p_emb) if training else None)
    if training:
        loss.backward()
    torch.cuda.synchronize()
    out = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in net.last_step.items()}
    out["grads"] = product_grads(net, list(P.keys())) if training else None
    return out


def run_oracle_step(P, batch, masks, *, L, bi, beta, use_projection, p_emb, training=True, dtype=torch.float64):
    """fp64 oracle by default: the tighter gold for the 1e-4 gate (fp32-vs-fp64 oracle noise is ~5e-6)."""
    Pd = {k: v.to(dtype).clone().requires_grad_(training) for k, v in P.items()}
    m = masks
    if masks is not None:
        def cast(t):
            return None if t is None else t.to(dtype)

        m = R.StepMasks(masks.emb_row_keep, cast(masks.whh_mask), cast(masks.fc1_w), cast(masks.do1), cast(masks.do2),
                        cast(masks.fc2_w))
    out = R.step(batch, Pd, num_layers=L, bi_reduce=bi, beta_classifier=beta, training=training, emb_droprate=p_emb,
                 use_projection=use_projection, masks=m if training else None)
    grads = None
    if training:
        out.loss.backward()
        grads = {n: p.grad.float() for n, p in Pd.items()}
    return {"loss": out.loss.detach().float(), "classifier_loss": out.classifier_loss.detach().float(),
            "triplet_loss": out.triplet_loss.detach().float(), "y_hat": out.y_hat.detach().float(),
            "z": torch.stack([z.detach().float() for z in out.z]), "grads": grads,
            "lengths": torch.tensor([[i.T1 for i in out.info], [i.T_eff for i in out.info]], dtype=torch.int32)}
