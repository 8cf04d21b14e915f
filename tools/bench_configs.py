#!/usr/bin/env python
"""Timings of the two BASELINE.json configurations that are parity cases rather than the headline bench line:

    python tools/bench_configs.py --config 4   # inference: 20k-protein synthetic proteome (T=1500, batch 512) + all-pairs PPI scores
    python tools/bench_configs.py --config 5   # stress encoder: E=256, 3-layer bi-LSTM, mean pooling, T=4000, batch 256 (one GPU's share)

One JSON line per config on stdout.  CUDA-event timing after a warm-up; inputs follow SURVEY.md 8(d) (seeds 4321 / 777)."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import intrepppid_b200 as ib  # noqa: E402
from intrepppid_b200 import _lib  # noqa: E402


def timed(fn, warm=1, reps=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


def config4(args):
    M, T, bs = args.proteins, 1500, 512
    torch.manual_seed(0)
    net = ib.intrepppid_network(1, precision=args.mode).cuda().eval()
    net.encoder.check_lengths = False
    x = torch.randint(1, 250, (M, T), generator=torch.Generator().manual_seed(4321)).cuda()
    with torch.no_grad():
        ms_enc, z = timed(lambda: net.embed(x, bs), warm=1, reps=1)
        P = M * (M + 1) // 2
        ms_pairs, prob = timed(lambda: net.score_pairs(z), warm=1, reps=1)
    print(json.dumps({"config": 4, "mode": args.mode, "proteins": M, "trunc_len": T, "batch": bs,
                      "encode_ms": ms_enc, "encode_seqs_per_s": M / ms_enc * 1e3,
                      "pairs": P, "pairs_ms": ms_pairs, "pairs_per_s": P / ms_pairs * 1e3,
                      "pairs_out_GBps": P * 4 / ms_pairs / 1e6, "prob_mean": float(prob.mean())}))


def config5(args):
    E, L, B, T = 256, 3, args.batch, args.len
    torch.manual_seed(0)
    net = ib.intrepppid_network(1, embedding_size=E, rnn_num_layers=L, bi_reduce="mean", precision=args.mode).cuda()
    net.encoder.check_lengths = False
    x = torch.randint(1, 250, (B, T), generator=torch.Generator().manual_seed(777)).cuda()
    tokens = float(B * T)
    net.eval()
    with torch.no_grad():
        _lib.timing_enable(True)
        ms_inf, z = timed(lambda: net.encoder(x), warm=1, reps=1)
        fam_inf = {k: round(v[0] / 2, 3) for k, v in _lib.timing_read().items() if v[0] > 0}  # warm-up + 1 rep were both timed
        _lib.timing_enable(False)
    out = {"config": 5, "mode": args.mode, "E": E, "layers": L, "bi_reduce": "mean", "batch": B, "trunc_len": T,
           "infer_ms": ms_inf, "infer_tokens_per_s": tokens / ms_inf * 1e3,
           "infer_tflops_dense": tokens * 8388608 / ms_inf / 1e9, "infer_kernel_ms": fam_inf}
    if not args.no_train:
        net.train()

        def step():
            net.zero_grad(set_to_none=True)
            zz = net.encoder(x)
            zz.square().mean().backward()
            return zz

        _lib.timing_enable(True)
        ms_tr, _ = timed(step, warm=1, reps=1)
        fam_tr = {k: round(v[0] / 2, 3) for k, v in _lib.timing_read().items() if v[0] > 0}  # warm-up + 1 rep were both timed
        _lib.timing_enable(False)
        out.update({"train_ms": ms_tr, "train_tokens_per_s": tokens / ms_tr * 1e3, "train_kernel_ms": fam_tr,
                    "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9})
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[4, 5])
    ap.add_argument("--mode", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--proteins", type=int, default=20000)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--len", type=int, default=4000)
    ap.add_argument("--no-train", action="store_true")
    a = ap.parse_args()
    (config4 if a.config == 4 else config5)(a)
