#!/usr/bin/env python
"""Timings of the two BASELINE.json configurations that are parity cases rather than the headline bench line:

    python tools/bench_configs.py --config 4   # inference: 20k-protein synthetic proteome (T=1500, batch 512) + all-pairs PPI scores
    python tools/bench_configs.py --config 6   # infer from_csv: ragged 20k proteome embedded once at batch-of-one semantics + 1M scored rows
    python tools/bench_configs.py --config 5   # stress encoder: E=256, 3-layer bi-LSTM, mean pooling, T=4000, batch 256 (one GPU's share)

One JSON line per config on stdout; bench.py imports the same functions for the `other_configs` object of its line.  CUDA-event timing after a warm-up; inputs follow SURVEY.md 8(d) (seeds 4321 / 777)."""
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
    return ({"config": 4, "mode": args.mode, "proteins": M, "trunc_len": T, "batch": bs,
                      "encode_ms": ms_enc, "encode_seqs_per_s": M / ms_enc * 1e3,
                      "pairs": P, "pairs_ms": ms_pairs, "pairs_per_s": P / ms_pairs * 1e3,
                      "pairs_out_GBps": P * 4 / ms_pairs / 1e6, "prob_mean": float(prob.mean())})


def config_csv(args):
    """`infer from_csv` workload: a ragged synthetic proteome (clipped log-normal lengths) and a list
    of interaction rows; every distinct protein is embedded once with batch-of-one semantics (intrepppid_b200.infer), then the rows
    are scored from the cache.  The reference encodes 2 proteins per ROW at batch 1 (cli/infer.py:216-222)."""
    from intrepppid_b200 import infer

    M, T, R = args.proteins, 1500, args.rows
    torch.manual_seed(0)
    net = ib.intrepppid_network(1, precision=args.mode).cuda().eval()
    g = torch.Generator().manual_seed(4321)
    x = torch.randint(1, 250, (M, T), generator=g)
    lens = torch.clamp(torch.exp(torch.randn(M, generator=g) * 0.6 + 6.0).long(), 50, T)
    x[torch.arange(T).unsqueeze(0) >= lens.unsqueeze(1)] = 0
    x = x.cuda()
    ia = torch.randint(0, M, (R,), generator=g).int().cuda()
    ib_ = torch.randint(0, M, (R,), generator=g).int().cuda()
    with torch.no_grad():
        l0 = _lib.launch_count()
        infer.embed_batch1(net, x)
        launches = _lib.launch_count() - l0
        _lib.timing_enable(True)
        ms_enc, z = timed(lambda: infer.embed_batch1(net, x), warm=0, reps=1)
        fam = {k: round(v[0], 2) for k, v in _lib.timing_read().items() if v[0] > 0}
        _lib.timing_enable(False)
        import time
        torch.cuda.synchronize(); t0 = time.perf_counter()
        keys_t = torch.stack(infer.batch1_lengths(x, net.encoder.embedder.weight), 1).cpu().tolist()
        t1 = time.perf_counter()
        infer.plan_buckets(keys_t)
        t2 = time.perf_counter()
        ms_pairs, prob = timed(lambda: net.score_pairs(z, ia, ib_), warm=1, reps=1)
        net.encoder.check_lengths = False
        ms_mixed, _ = timed(lambda: net.embed(x, 512), warm=1, reps=1)
    keys = torch.stack(infer.batch1_lengths(x, net.encoder.embedder.weight), 1).cpu().tolist()
    plan = infer.plan_buckets(keys)
    return ({"config": "from_csv", "mode": args.mode, "proteins": M, "rows": R, "mean_len": float(lens.float().mean()),
                      "distinct_lengths": len({tuple(k) for k in keys}), "launch_sets": len(plan),
                      "groups_by_size": {b: sum(len(gr) for bb, gr in plan if bb == b) for b in (8, 4, 2, 1)},
                      "embed_batch1_ms": ms_enc, "embed_batch1_seqs_per_s": M / ms_enc * 1e3, "kernel_launches": launches,
                      "embed_batch1_kernel_ms": fam, "lengths_ms": (t1 - t0) * 1e3, "plan_ms": (t2 - t1) * 1e3,
                      "score_rows_ms": ms_pairs, "rows_per_s_end_to_end": R / (ms_enc + ms_pairs) * 1e3,
                      "mixed_batch512_embed_ms (different semantics: pads are stepped)": ms_mixed,
                      "prob_mean": float(prob.mean())})


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
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[4, 5, 6], help="6 = the infer-from_csv workload")
    ap.add_argument("--mode", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--proteins", type=int, default=20000)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--len", type=int, default=4000)
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--rows", type=int, default=1000000)
    a = ap.parse_args()
    print(json.dumps({4: config4, 5: config5, 6: config_csv}[a.config](a)))
