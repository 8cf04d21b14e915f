#!/usr/bin/env python
"""bench.py -- train sequences/sec of the INTREPPPID e2e_rnn_triplet step (5 encoder fwd + triplet + head + BCE + full backward
+ AdamW) at trunc_len 1500, batch 80 per GPU, vocab 250, embed 64, 2-layer bi-LSTM, bi_reduce last (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode fp32|bf16] [--variant dropout|full_t]
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores, same config/metric
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (one rank per GPU, NCCL; weak scaling: 80 samples / GPU)

One JSON line on stdout (rank 0).  `value` = sequences encoded per second, whole job, inputs resident in HBM, device-timed
(CUDA events, max over ranks).  `e2e` = the same through the public module API with HOST (pinned) token buffers, H2D copies
and the loss D2H inside the timed region.  `roofline` = dominant kernel family, algorithmic bytes / event-timed duration against
the measured HBM peak.  `cpu_baseline` = the oracle port on this box's host cores (N=1, rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

B, T, V, E, L, BI = 80, 1500, 250, 64, 2, "last"
WORKLOAD = "e2e_rnn_triplet train step: batch 80/GPU x 5 sequences, trunc_len 1500, vocab 250, embed 64, 2-layer bi-LSTM, bi_reduce last"


def synthetic_batch(seed: int):
    """five randint(1,250,(80,1500)) int64 tensors (p1,p2,anchor,pos,neg) + labels, full length, seed 1234+rank."""
    g = torch.Generator().manual_seed(seed)
    seqs = [torch.randint(1, V, (B, T), generator=g) for _ in range(5)]
    y = torch.randint(0, 2, (B,), generator=g)
    return seqs + [y]


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as fh:
                d = json.load(fh)
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------------------------
def algorithmic_bytes(family: str, tokens_per_chain: float, H: int) -> float:
    """Minimum HBM bytes one launch of the recurrent kernels moves (fp32 storage), per DESIGN.md section 4:
    gates 16H B, c 4H B, h 4H B, dy 4H B, token id 4 B per (token, chain)."""
    g, c, h = 16 * H, 4 * H, 4 * H
    per_token = {
        "lstm_fwd_l0": 2 * (4 + g + c + h),        # two chains: read id, write gates + c + h
        "lstm_fwd_upper": 1 * (g + g + c + h),     # one live chain under "last": read xproj, write gates + c + h
        "lstm_bwd_upper": 1 * (g + c + g),         # read gates + c, write dgates in place
        "lstm_bwd_l0": 2 * (g + c + h + g),        # read gates + c + dy, write dgates
    }[family]
    return per_token * tokens_per_chain


def ncu_traffic(family: str, tokens_per_chain: float, mode: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `family`, from the committed `ncu --set full` capture
    (profiles/r1_traffic.json: bytes per token measured on the same workload, scaled to this run's token count).  None when no
    capture exists for this precision mode."""
    path = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if mode != "fp32" or not os.path.exists(path):
        return None
    try:
        with open(path) as fh:
            return float(json.load(fh)["kernels"][family]["dram_bytes_per_token"]) * tokens_per_chain
    except Exception:
        return None


def run_b200(args):
    import torch.distributed as dist

    import intrepppid_b200 as ib
    from intrepppid_b200 import _lib
    from intrepppid_b200.optim import FusedAdamW
    from intrepppid_b200.parallel import GradientAllReducer

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout for the single JSON line
        dist.init_process_group("nccl", device_id=dev)

    def make_net(variant, mode):
        torch.manual_seed(0)
        net = ib.intrepppid_network(1, precision=mode, optimizer_type="adamw",
                                    embedding_droprate=0.0 if variant == "full_t" else 0.3).to(dev).train()
        net.encoder.check_lengths = False  # no host sync in the step; lengths are read back after the timed region
        return net

    host_batch = synthetic_batch(1234 + rank)
    dev_batch = [t.to(dev) for t in host_batch]

    def timed_steps(net, K, W, e2e=False, collect=None, host_batch=host_batch):
        params = [p for p in net.parameters() if p.requires_grad]
        opt = FusedAdamW(params, lr=1e-3)  # ib200_adamw_step: one launch over the 23 live tensors
        reducer = GradientAllReducer(net) if world > 1 else None
        pinned = [t.pin_memory() for t in host_batch] if e2e else None
        lens_log = []
        if e2e:
            # end-to-end pipeline of a training loop with a prefetching loader: the inputs of step k+1 stream from pinned host
            # memory into the second device buffer on a copy stream while step k computes, and the loss of step k is read on the
            # host (pinned D2H) while step k+1 runs.  Every step still copies its own inputs and delivers its own loss.
            copy_stream = torch.cuda.Stream()
            bufs = [[torch.empty(t.shape, dtype=t.dtype, device=dev) for t in host_batch] for _ in range(2)]
            copied = [torch.cuda.Event() for _ in range(2)]
            consumed = [torch.cuda.Event() for _ in range(2)]
            loss_ev = [torch.cuda.Event() for _ in range(2)]
            loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
            state = {"k": 0, "losses": []}
            for ev in consumed:
                ev.record()

            def prefetch(k):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[k % 2])  # the step that last read this buffer is done with it
                    for d, h in zip(bufs[k % 2], pinned):
                        d.copy_(h, non_blocking=True)
                    copied[k % 2].record(copy_stream)

            prefetch(0)

        def one_step():
            if e2e:
                k = state["k"]
                torch.cuda.current_stream().wait_event(copied[k % 2])
                prefetch(k + 1)
                batch = bufs[k % 2]
            else:
                batch = dev_batch
            opt.zero_grad(set_to_none=True)
            loss = net.step(batch, "train")
            loss.backward()
            if reducer is not None:
                reducer.finish()
            opt.step()
            lens_log.append(net.encoder.last_lengths)
            if e2e:
                consumed[k % 2].record()
                loss_host[k % 2:k % 2 + 1].copy_(loss.detach().reshape(1), non_blocking=True)
                loss_ev[k % 2].record()
                if k > 0:  # the caller reads every step's loss, one step late
                    loss_ev[(k - 1) % 2].synchronize()
                    state["losses"].append(float(loss_host[(k - 1) % 2]))
                state["k"] = k + 1

        for _ in range(W):
            one_step()
        lens_log.clear()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if collect is not None:
            _lib.timing_enable(True)
        l0 = _lib.launch_count()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(K):
            one_step()
        end.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = start.elapsed_time(end)
        launches = _lib.launch_count() - l0
        if collect is not None:
            collect.update(_lib.timing_read())
            _lib.timing_enable(False)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        if reducer is not None:
            reducer.remove()
        lens = torch.stack(lens_log).float().mean(0).cpu() if lens_log else None  # [2,5] mean over steps
        return ms, launches, lens

    K, W = args.steps, args.warmup
    net = make_net(args.variant, args.mode)
    sampler = ClockSampler(local_rank)
    fam = {}
    if rank == 0:
        sampler.start()
    # pass 1 (the headline): K steps, no per-launch events.  pass 2: the same K steps again with a CUDA-event pair around every
    # launch of the library (per-family times, roofline); the event records cost ~0.3 ms/step, so they stay out of `value`.
    ms, launches, lens = timed_steps(net, K, W)
    clocks = sampler.stop() if rank == 0 else None
    ms_instr, _, lens = timed_steps(net, K, 1, collect=fam)
    ms_e2e, _, _ = timed_steps(net, K, max(1, W // 2), e2e=True)
    # the same end-to-end loop fed with narrowed ids (uint8: V = 250 fits a byte; IB200_TOK_U8) -- SURVEY 8f "input feeding"
    narrow_batch = [t.to(torch.uint8) for t in host_batch[:5]] + [host_batch[5]]
    ms_e2e_u8, _, _ = timed_steps(net, K, max(1, W // 2), e2e=True, host_batch=narrow_batch)

    seqs_per_step = 5 * B * world
    value = seqs_per_step * K / (ms / 1e3)
    e2e_value = seqs_per_step * K / (ms_e2e / 1e3)

    extra = {}
    if rank == 0 and args.extras:
        other_mode = "bf16" if args.mode == "fp32" else "fp32"
        other_variant = "full_t" if args.variant == "dropout" else "dropout"
        for tag, (v, m) in {f"{other_mode}/{args.variant}": (args.variant, other_mode),
                            f"{args.mode}/{other_variant}": (other_variant, args.mode)}.items():
            if world > 1:
                break  # extras are single-GPU context only
            n2 = make_net(v, m)
            ms2, _, lens2 = timed_steps(n2, max(3, K // 2), 3)
            extra[tag] = {"seqs_per_s": 5 * B * max(3, K // 2) / (ms2 / 1e3), "ms_per_step": ms2 / max(3, K // 2),
                          "mean_T_eff": float(lens2[1].mean())}
            del n2

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family (rank 0's launches) -----------------------------------------------------------
    peak, peak_src = measured_peaks()
    teff = lens[1]  # mean T_eff per group over the timed steps
    tokens_per_chain = float(B * teff.sum())
    total_kernel_ms = sum(v[0] for v in fam.values())
    shares = {k: {"ms_per_step": v[0] / K, "launches_per_step": v[1] / K, "share": v[0] / total_kernel_ms} for k, v in fam.items()}
    dom = max((k for k in fam if k.startswith("lstm_")), key=lambda k: fam[k][0])
    dom_ms = fam[dom][0] / fam[dom][1]
    dom_bytes = algorithmic_bytes(dom, tokens_per_chain, E)
    achieved = dom_bytes / (dom_ms / 1e3) / 1e9
    chains_steps = float(teff.max())
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(dom, tokens_per_chain, args.mode), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes,
                "timed_in": f"second pass of the same {K} steps with per-launch CUDA events on the launch stream "
                            f"({ms_instr / K:.3f} ms/step instrumented vs {ms / K:.3f} ms/step in the headline pass)",
                "avg_launch_ms": dom_ms, "tau_us_per_cell_step": dom_ms * 1e3 / chains_steps,
                "note": "recurrent kernels are bound by the dependent chain (tau per cell step), not by HBM; see DESIGN.md"}

    out = {
        "metric": "train seqs/sec (fwd+bwd+triplet) @ trunc_len 1500", "value": value, "unit": "seqs/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 state; bf16 hi/lo-split tensor-core products (3 MMAs), fp32 accumulate" if args.mode == "fp32"
                 else "bf16 tensor-core products, fp32 accumulate/state",
        "data": "synthetic", "impl": "b200",
        "config": {"workload": WORKLOAD, "variant": args.variant, "mode": args.mode, "batch_per_gpu": B, "global_batch": B * world,
                   "seqs_per_sample": 5, "trunc_len": T, "mean_T_eff_per_group": [round(float(x), 1) for x in teff],
                   "dropout_rates": 0.3 if args.variant == "dropout" else "embedding_droprate=0, others 0.3",
                   "optimizer": "AdamW (ib200_adamw_step, one multi-tensor launch) inside the timed step", "parallelism": f"dp{world}",
                   "l2": "no flush needed: each step streams ~4 GB of activations (>> 126 MB L2)"},
        "samples_per_s": value / 5,
        "e2e": {"value": e2e_value, "unit": "seqs/s", "ms_per_step": ms_e2e / K,
                "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host_batch), "d2h_bytes_per_step": 4,
                "pipeline": "public module API (TripletE2ENet.step + backward + AdamW); int64 token ids from pinned host memory, "
                            "double-buffered H2D on a copy stream (step k+1's copy overlaps step k), loss D2H every step, read on "
                            "the host one step late",
                "uint8_tokens": {"value": seqs_per_step * K / (ms_e2e_u8 / 1e3), "ms_per_step": ms_e2e_u8 / K,
                                 "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in narrow_batch)}},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "kernel_families": shares,
    }
    if extra:
        out["other_variants_1gpu"] = extra
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args.variant)
    if world > 1:
        dist.destroy_process_group()
    emit(out)


# ------------------------------------------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------------------------------------
def _cpu_step_fn(variant: str):
    from oracle import restatement as R

    torch.set_num_threads(os.cpu_count())
    P = {k: v.clone().requires_grad_(True) for k, v in R.init_params(vocab=V, E=E, L=L, seed=0).items()}
    opt = torch.optim.AdamW(list(P.values()), lr=1e-3)
    p_emb = 0.0 if variant == "full_t" else 0.3
    counter = [0]

    def step(batch):
        nb = batch[0].shape[0]
        counter[0] += 1
        masks = R.draw_step_masks(nb, V, E, emb_droprate=p_emb, rnn_droprate=0.3, do_rate=0.3, seed=100 + counter[0])
        opt.zero_grad(set_to_none=True)
        out = R.step(batch, P, num_layers=L, bi_reduce=BI, beta_classifier=2.0, training=True, emb_droprate=p_emb, masks=masks,
                     impl="vf")
        out.loss.backward()
        opt.step()
        return float(out.loss)

    return step


def cpu_baseline(variant: str):
    """The oracle port (same ATen LSTM entry point the reference's nn.LSTM uses on CPU) on this box's host cores:
    one full-size step (B=80, T=1500) after a small warm-up."""
    step = _cpu_step_fn(variant)
    batch = synthetic_batch(1234)
    small = [t[:4] for t in batch]
    step(small)
    t0 = time.perf_counter()
    step(batch)
    dt = time.perf_counter() - t0
    return {"value": 5 * B / dt, "unit": "seqs/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"1 full step (B={B}, T={T}, fwd+bwd+AdamW, fp32, torch {torch.__version__} CPU LSTM) after a B=4 warm-up; {dt:.2f} s",
            "threads": torch.get_num_threads()}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    step = _cpu_step_fn(args.variant)
    batch = synthetic_batch(1234)
    K, W = args.steps, args.warmup
    probe = [t[:4] for t in batch]
    step(probe)
    t0 = time.perf_counter()
    step(probe)
    per_sample = (time.perf_counter() - t0) / 4
    budget = 150.0
    bs = int(max(4, min(B, budget / ((K + W) * per_sample))))
    sample = [t[:bs] for t in batch]
    for _ in range(W):
        step(sample)
    t0 = time.perf_counter()
    for _ in range(K):
        step(sample)
    dt = time.perf_counter() - t0
    value = 5 * bs * K / dt
    cb = {"value": value, "unit": "seqs/s", "cores": os.cpu_count(), "kind": "port", "threads": torch.get_num_threads(),
          "sample": f"each step = the first {bs} of the {B} samples of the workload batch (5 sequences each, T={T}), "
                    f"fwd+bwd+AdamW, fp32, all host threads"}
    emit({
        "impl": "reference", "metric": "train seqs/sec (fwd+bwd+triplet) @ trunc_len 1500", "value": value, "unit": "seqs/s",
        "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "variant": args.variant, "sample_batch": bs, "trunc_len": T,
                   "note": "reference = CPU oracle port (oracle/restatement.py: the reference's algorithm on torch CPU, same "
                           "ATen LSTM as nn.LSTM); the reference itself is Python and is not present on the GPU box"},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "seqs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


_JSON_FD = None


def _reserve_stdout():
    """stdout carries exactly ONE JSON line: native libraries (NCCL prints its version banner to fd 1) are pointed at stderr and
    the result line is written to a saved duplicate of the original stdout."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def main():
    _reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--variant", default="dropout", choices=["dropout", "full_t"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extras", action="store_true", help="also time the other precision mode / variant (1 GPU)")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
