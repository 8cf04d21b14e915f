#!/usr/bin/env python
"""The headline training step on PADDED batches (clipped log-normal lengths, SURVEY 8d's third synthetic variant) next to the
full-length batch of bench.py: real dataloader batches are padded to trunc_len, and without packing (awd_lstm.py:56) every sequence
steps through its pads up to the longest one of the call.  One JSON line; CUDA events, 3 warm-up + 10 timed steps."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import intrepppid_b200 as ib  # noqa: E402
from intrepppid_b200 import _lib  # noqa: E402

B, T, V = 80, 1500, 250


def batch(padded: bool, seed=1234):
    g = torch.Generator().manual_seed(seed)
    seqs = [torch.randint(1, V, (B, T), generator=g) for _ in range(5)]
    y = torch.randint(0, 2, (B,), generator=g)
    mean_len = float(T)
    if padded:
        tot = 0
        for s in seqs:
            lens = torch.clamp(torch.exp(torch.randn(B, generator=g) * 0.6 + 6.0).long(), 50, T)
            lens[0] = T  # at least one full-length row per call: T1 = trunc_len
            s[torch.arange(T).unsqueeze(0) >= lens.unsqueeze(1)] = 0
            tot += float(lens.float().mean())
        mean_len = tot / 5
    return [t.cuda() for t in seqs + [y]], mean_len


def run(padded: bool):
    torch.manual_seed(0)
    net = ib.intrepppid_network(1, optimizer_type="adamw").cuda().train()
    net.encoder.check_lengths = False
    opt = ib.FusedAdamW([p for p in net.parameters() if p.requires_grad], lr=1e-3)
    data, mean_len = batch(padded)

    def step():
        opt.zero_grad(set_to_none=True)
        net.step(data, "train").backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    _lib.timing_enable(True)
    step()
    fam = {k: round(v[0], 3) for k, v in _lib.timing_read().items() if v[0] > 0.05}
    _lib.timing_enable(False)
    return {"ms_per_step": ms, "seqs_per_s": 5 * B / ms * 1e3, "mean_tokens_per_seq": mean_len,
            "T_eff": net.encoder.last_lengths[1].tolist(), "kernel_ms": fam}


if __name__ == "__main__":
    print(json.dumps({"workload": "e2e_rnn_triplet train step, batch 80 x 5, trunc_len 1500, fp32 mode, dropout 0.3",
                      "full_length": run(False), "padded_lognormal": run(True)}))
