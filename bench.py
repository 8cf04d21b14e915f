#!/usr/bin/env python
"""bench.py -- train sequences/sec of the INTREPPPID e2e_rnn_triplet step (5 encoder fwd + triplet + head + BCE + full backward
+ AdamW) at trunc_len 1500, batch 80 per GPU, vocab 250, embed 64, 2-layer bi-LSTM, bi_reduce last (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode fp32|bf16] [--variant dropout|full_t] [--no-extras]
    python bench.py --impl reference ...      # the reference's own CPU path on the host cores, same config / metric
    python bench.py --workload config5 ...    # BASELINE configs[4]: E=256, 3 layers, mean, T=4000, batch 256 per GPU (encoder fwd+bwd)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (one rank per GPU, NCCL; weak scaling: 80 samples / GPU)

One JSON line on stdout (rank 0):
  value          sequences encoded per second, whole job, inputs resident in HBM, device-timed (CUDA events, max over ranks)
  e2e            the same through the public API with HOST batches: intrepppid_b200.feed.DeviceFeeder (pinned staging, H2D on a copy
                 stream) -> TripletE2ENet.step -> backward -> AdamW, loss read on the host every step
  roofline       the dominant kernel family; roofline_all = every family (HBM GB/s and fraction, tensor-pipe fraction for the GEMMs,
                 tau per cell step for the recurrent kernels)
  cpu_baseline   the reference's own step on this box's host cores (N=1): median of 3 full-batch steps + a 1-thread figure
  gpu_reference  the UNMODIFIED reference modules on this same GPU through torch CUDA (cuDNN LSTM) -- the reference as users run it
  other_configs  bf16 mode, the full-T variant, BASELINE configs 4 / 5 and the from_csv workload (N=1; --no-extras skips them)
The only code that touches oracle/ is the CPU / reference legs at the bottom.
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
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

B, T, V, E, L, BI = 80, 1500, 250, 64, 2, "last"
METRIC = "train seqs/sec (fwd+bwd+triplet) @ trunc_len 1500"
WORKLOAD = "e2e_rnn_triplet train step: batch 80/GPU x 5 sequences, trunc_len 1500, vocab 250, embed 64, 2-layer bi-LSTM, bi_reduce last"
C5 = dict(E=256, L=3, B=256, T=4000, bi="mean")
C5_METRIC = "train seqs/sec (encoder fwd+bwd) @ trunc_len 4000, embed 256, 3-layer bi-LSTM, bi_reduce mean"
C5_WORKLOAD = "scaled encoder stress: batch 256/GPU, trunc_len 4000, vocab 250, embed 256, 3-layer bi-LSTM, bi_reduce mean"


def synthetic_batch(seed: int):
    """five randint(1,250,(80,1500)) int64 tensors (p1,p2,anchor,pos,neg) + labels, full length, seed 1234+rank."""
    g = torch.Generator().manual_seed(seed)
    seqs = [torch.randint(1, V, (B, T), generator=g) for _ in range(5)]
    y = torch.randint(0, 2, (B,), generator=g)
    return seqs + [y]


def workload_config(variant: str, world: int) -> dict:
    """The part of `config` that names the workload -- identical in the b200 arm and the reference arm."""
    return {"workload": WORKLOAD, "variant": variant, "batch_per_gpu": B, "global_batch": B * world, "seqs_per_sample": 5,
            "trunc_len": T, "vocab": V, "embed": E, "layers": L, "bi_reduce": BI,
            "dropout_rates": 0.3 if variant == "dropout" else "embedding_droprate=0, others 0.3",
            "optimizer": "AdamW step inside the timed step", "parallelism": f"dp{world}"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as fh:
                d = json.load(fh)
            return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", 1386.5)),
                    "bf16_tflops_burst": float(d.get("bf16_tflops", 1644.4)), "source": "measured (MEASURED_PEAKS.json)"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0, "bf16_tflops_burst": 1590.0, "source": "fallback (B200_PROFILING.md)"}


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
# roofline accounting (DESIGN.md sections 4 / 5: algorithmic bytes and FLOPs per token row, per kernel family)
# ------------------------------------------------------------------------------------------------------------------------------
def family_work(rows: float, H: int, n_seq_rows: float, upper_chains: int):
    """family -> (algorithmic HBM bytes per step, useful FLOPs per step or None, kind).  rows = token rows (n, t < T_eff) of the
    step; fp32-sized storage (the bf16 hi|lo planes of the TMA path take the same 4 bytes per element).  upper_chains = live
    chains of the top layer (1 under bi_reduce=last).  L = 2 layout of the headline workload."""
    g, c, h = 16 * H, 4 * H, 4 * H
    nl = upper_chains
    return {
        "lengths": (n_seq_rows * (8 + 4 + 4), None, "hbm"),                                  # read int64 ids, write + re-read int32 copy
        "lstm_fwd_l0": (rows * 2 * (4 + g + c + h), None, "chain"),                            # read id; write gates + c + h (2 chains)
        "gemm_nt_xproj": (rows * nl * (2 * H * 4 + g), rows * nl * 2 * (2 * H) * (4 * H), "gemm"),   # read Y0 [2H], write X [4H]
        "lstm_fwd_upper": (rows * nl * (g + g + c + h), None, "chain"),                        # read xproj; write gates + c + h
        "lstm_bwd_upper": (rows * nl * (g + c + g), None, "chain"),                            # read gates + c; dgates in place
        "gemm_tn_dw": (rows * nl * (g + 2 * H * 4 + H * 4), rows * nl * 2 * (4 * H) * (3 * H), "gemm"),  # dA^T [Y0 | Y1 shifted]
        "gemm_nt_dgrad": (rows * (nl * g + 2 * H * 4), rows * nl * 2 * (4 * H) * (2 * H), "gemm"),    # read dgates, write dY [2H]
        "lstm_bwd_l0": (rows * 2 * (g + c + h + g), None, "chain"),                            # read gates + c + dy; dgates in place
        "l0_grads": (rows * 2 * (g + H * 4 + 4), rows * 2 * 2 * (4 * H) * H, "gemm"),          # dA^T [Y0 shifted | onehot(tok)]; FLOPs: dW_hh only
    }


def build_roofline(fam, K, lens, mode, peaks):
    teff = lens[1]
    rows = float(B * teff.sum())
    chain_steps = float(teff.max())
    work = family_work(rows, E, float(5 * B * T), 1 if BI == "last" else 2)
    mma_per_product = 3 if mode == "fp32" else 1
    out = {}
    for name, (ms_total, calls) in fam.items():
        ms = ms_total / K
        ent = {"ms_per_step": ms, "launcher_calls_per_step": calls / K}
        if name in work:
            nbytes, flops, kind = work[name]
            gbs = nbytes / (ms / 1e3) / 1e9
            ent.update({"algorithmic_bytes_per_step": nbytes, "GBps": gbs, "hbm_frac": gbs / peaks["hbm_gbs"]})
            if flops is not None:
                tf = flops / (ms / 1e3) / 1e12
                ent.update({"useful_tflops": tf, "tensor_frac_useful": tf / peaks["bf16_tflops_sustained"],
                            "tensor_frac_executed": tf * mma_per_product / peaks["bf16_tflops_sustained"]})
            if kind == "chain":
                ent["tau_us_per_cell_step"] = ms * 1e3 / chain_steps
                ent["bound"] = "latency (dependent chain of T_eff cell steps); HBM fraction reported for the activation streaming"
            elif kind == "gemm":
                ent["bound"] = "hbm (K <= 256: arithmetic intensity far below the tensor ridge)"
            else:
                ent["bound"] = "hbm / launch latency"
        else:
            ent["bound"] = "launch latency (single small grid)"
        out[name] = ent
    return out, rows, chain_steps


def ncu_traffic(family: str, rows: float, mode: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `family`, from the committed `ncu --set full` capture
    (bytes per token row measured on the same workload, scaled to this run's row count).  None when no capture exists."""
    if mode != "fp32":
        return None
    for name in ("r2_traffic.json", "r1_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            try:
                with open(path) as fh:
                    return float(json.load(fh)["kernels"][family]["dram_bytes_per_token"]) * rows
            except Exception:
                continue
    return None


# ------------------------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------------------------
def _dist_setup(args):
    import torch.distributed as dist

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout for the single JSON line
        dist.init_process_group("nccl", device_id=dev)
    return dist, rank, world, local_rank, dev


def _timed_region(one_step, K, W, world, dist, dev, collect=None):
    """W warm-up steps, then exactly K steps between barrier + synchronize on both sides; CUDA events; max over ranks."""
    from intrepppid_b200 import _lib

    import gc

    for _ in range(W):
        one_step()
    torch.cuda.synchronize()
    # a full-heap pass of Python's cyclic collector costs several ms in this process (torch + the reference shim loaded) and would
    # land in one of the K steps at random: collect now and park the survivors in the permanent generation for the timed region
    gc.collect()
    gc.freeze()
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
    gc.unfreeze()
    launches = _lib.launch_count() - l0
    if collect is not None:
        collect.update(_lib.timing_read())
        _lib.timing_enable(False)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    return ms, launches


def run_b200(args):
    import intrepppid_b200 as ib
    from intrepppid_b200 import feed, ops
    from intrepppid_b200.optim import FusedAdamW
    from intrepppid_b200.parallel import GradientAllReducer

    dist, rank, world, local_rank, dev = _dist_setup(args)

    def make_net(variant, mode):
        torch.manual_seed(0)
        # default module settings: check_lengths stays on (device-side status word, checked lazily -- no host sync in the step)
        return ib.intrepppid_network(1, precision=mode, optimizer_type="adamw",
                                     embedding_droprate=0.0 if variant == "full_t" else 0.3).to(dev).train()

    host_batch = synthetic_batch(1234 + rank)
    dev_batch = [t.to(dev) for t in host_batch]

    def timed_steps(net, K, W, e2e=None, collect=None, comm=True, optimizer="adamw"):
        """e2e: None = inputs resident in HBM; "packed" = host batches as the product's loader workers deliver them (narrow_collate:
        uint8 ids in one [5,B,T] tensor); "int64" = the reference loader's default-collated int64 tuples (narrowed by the feeder)."""
        params = [p for p in net.parameters() if p.requires_grad]
        if optimizer == "adamw":
            opt = FusedAdamW(params, lr=1e-3)  # ib200_adamw_step: one launch over the 23 live tensors
        else:  # the reference's factory default (e2e_triplet.py:212-224): ib200_ranger21_step, three launches over the live tensors
            from intrepppid_b200.optim import FusedRanger21

            opt = FusedRanger21(params, lr=1e-2, weight_decay=1e-2, use_warmup=True, warmdown_active=True, num_batches_per_epoch=1000,
                                num_epochs=100, warmdown_start_pct=0.72)
        reducer = GradientAllReducer(net) if (world > 1 and comm) else None
        lens_log, state = [], {"k": 0, "losses": []}
        feeder = batches = None
        if e2e is not None:
            src = host_batch if e2e == "int64" else feed.pack_batch(host_batch[:5], host_batch[5], V)
            feeder = feed.DeviceFeeder((src for _ in range(K + W)), dev, V)
            batches = iter(feeder)
            loss_ev = [torch.cuda.Event() for _ in range(2)]
            loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()

        def one_step():
            batch = next(batches) if e2e is not None else dev_batch
            opt.zero_grad(set_to_none=True)
            loss = net.step(batch, "train")
            loss.backward()
            if reducer is not None:
                reducer.finish()
            opt.step()
            lens_log.append(net.encoder.last_lengths)
            if e2e is not None:
                k = state["k"]
                loss_host[k % 2:k % 2 + 1].copy_(loss.detach().reshape(1), non_blocking=True)
                loss_ev[k % 2].record()
                if k > 0:  # the caller reads every step's loss, one step late (the pinned D2H of step k-1 overlaps step k)
                    loss_ev[(k - 1) % 2].synchronize()
                    state["losses"].append(float(loss_host[(k - 1) % 2]))
                state["k"] = k + 1

        for _ in range(W):
            one_step()
        lens_log.clear()
        ms, launches = _timed_region(one_step, K, 0, world, dist, dev, collect)
        if reducer is not None:
            reducer.remove()
        ops.check_pending(sync=True)
        lens = torch.stack(lens_log).float().mean(0).cpu() if lens_log else None  # [2,5] mean over steps
        h2d = feeder.h2d_bytes / (K + W) if feeder is not None else 0
        return ms, launches, lens, h2d

    K, W = args.steps, args.warmup
    net = make_net(args.variant, args.mode)
    sampler = ClockSampler(local_rank)
    fam = {}
    if rank == 0:
        sampler.start()
    # pass 1 (the headline): K steps, no per-launch events.  pass 2: the same K steps again with a CUDA-event pair around every
    # launch of the library (per-family times, roofline); the event records cost ~0.3 ms/step, so they stay out of `value`.
    ms, launches, lens, _ = timed_steps(net, K, W)
    clocks = sampler.stop() if rank == 0 else None
    ms_instr, _, lens, _ = timed_steps(net, K, 1, collect=fam)
    # end-to-end passes: W warm-up steps each (a fresh feeder allocates its pinned staging buffers and copy stream in the first ones)
    ms_e2e_i64, _, _, _ = timed_steps(net, K, W, e2e="int64")
    # the end-to-end loop reads every step's loss on the host, so the host runs at most one step ahead and any host hiccup (scheduler,
    # allocator) lands in the figure: two passes of K steps, both reported, the better one is `e2e.value`
    e2e_passes = [timed_steps(net, K, W, e2e="packed") for _ in range(2)]
    ms_e2e, _, _, h2d_packed = min(e2e_passes, key=lambda r: r[0])
    ms_nocomm = timed_steps(net, K, 1, comm=False)[0] if world > 1 else None

    seqs_per_step = 5 * B * world
    value = seqs_per_step * K / (ms / 1e3)
    e2e_value = seqs_per_step * K / (ms_e2e / 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    teff = lens[1]  # mean T_eff per group over the timed steps
    roofline_all, rows, chain_steps = build_roofline(fam, K, lens, args.mode, peaks)
    dom = max((k for k in fam if k.startswith("lstm_")), key=lambda k: fam[k][0])
    d = roofline_all[dom]
    dom_ms = fam[dom][0] / fam[dom][1]
    dom_bytes = d["algorithmic_bytes_per_step"] * K / fam[dom][1]
    achieved = dom_bytes / (dom_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "limiter": "latency: a dependent chain of T_eff cell steps (tau_us_per_cell_step); the HBM fraction is the "
                                            "activation streaming that rides on it", "kernel": dom, "achieved": achieved,
                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": ncu_traffic(dom, rows, args.mode), "peak_source": peaks["source"],
                "algorithmic_bytes_per_launch": dom_bytes,
                "timed_in": f"second pass of the same {K} steps with per-launch CUDA events on the launch stream "
                            f"({ms_instr / K:.3f} ms/step instrumented vs {ms / K:.3f} ms/step in the headline pass)",
                "avg_launch_ms": dom_ms, "tau_us_per_cell_step": dom_ms * 1e3 / chain_steps}
    total_kernel_ms = sum(v[0] for v in fam.values())
    for k, v in fam.items():
        roofline_all[k]["share_of_kernel_time"] = v[0] / total_kernel_ms

    cfg = workload_config(args.variant, world)
    out = {
        "metric": METRIC, "value": value, "unit": "seqs/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 state; bf16 hi/lo-split tensor-core products (3 MMAs), fp32 accumulate" if args.mode == "fp32"
                 else "bf16 tensor-core products, fp32 accumulate/state",
        "data": "synthetic", "impl": "b200", "mode": args.mode,
        "config": cfg,
        "run": {"mean_T_eff_per_group": [round(float(x), 1) for x in teff], "token_rows_per_step": rows,
                "l2": "no flush needed: each step streams ~4 GB of activations (>> 126 MB L2)",
                "check_lengths": "default (device-side status word, lazy raise; no host sync in the step)",
                "gc": "gc.collect() + gc.freeze() after the warm-up steps of every timed region (no full-heap cyclic-GC pass inside the K steps)"},
        "samples_per_s": value / 5,
        "e2e": {"value": e2e_value, "unit": "seqs/s", "ms_per_step": ms_e2e / K,
                "passes_ms_per_step": [r[0] / K for r in e2e_passes],
                "h2d_bytes_per_step": int(h2d_packed), "d2h_bytes_per_step": 4,
                "pipeline": "public API: intrepppid_b200.feed.DeviceFeeder (uint8 ids as narrow_collate packs them in the loader workers, "
                            "pinned staging, H2D on a copy stream under the previous step) -> TripletE2ENet.step -> backward -> "
                            "FusedAdamW; loss D2H every step, read on the host one step late",
                "int64_loader": {"value": seqs_per_step * K / (ms_e2e_i64 / 1e3), "ms_per_step": ms_e2e_i64 / K,
                                 "note": "the reference loader's default-collated int64 batches (4.8 MB), narrowed to uint8 by the "
                                         "feeder on the consumer thread before the same H2D"}},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_all": roofline_all,
        "peaks": peaks,
    }
    if world > 1:
        out["comm_exposed_ms"] = (ms - ms_nocomm) / K
        out["ms_per_step_without_gradient_exchange"] = ms_nocomm / K
        dist.destroy_process_group()
    if world == 1:
        del net
        torch.cuda.empty_cache()
        if not args.no_extras:
            out["other_configs"] = other_configs(args, make_net, timed_steps)
        if not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args.variant)
            gr = gpu_reference(args.variant, dev)
            if gr is not None:
                gr["b200_over_gpu_reference"] = value / gr["value"]
                out["gpu_reference"] = gr
    emit(out)


def other_configs(args, make_net, timed_steps):
    """Driver-visible numbers for everything that is not the fp32 headline (N = 1): the other precision mode, the full-T variant
    (SURVEY 8d roofline-accounting run), BASELINE configs 4 / 5 and the from_csv workload (tools/bench_configs.py)."""
    import types

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_configs as bc

    out = {}
    other_mode = "bf16" if args.mode == "fp32" else "fp32"
    other_variant = "full_t" if args.variant == "dropout" else "dropout"
    for tag, (v, m) in {f"{other_mode}/{args.variant}": (args.variant, other_mode),
                        f"{args.mode}/{other_variant}": (other_variant, args.mode)}.items():
        n2 = make_net(v, m)
        k2 = max(5, args.steps // 2)
        ms2, _, lens2, _ = timed_steps(n2, k2, 3)
        out[tag] = {"seqs_per_s": 5 * B * k2 / (ms2 / 1e3), "ms_per_step": ms2 / k2, "mean_T_eff": float(lens2[1].mean()), "steps": k2}
        del n2
    try:  # the headline step with the reference's factory-default optimizer (ranger21_xx) instead of AdamW
        n2 = make_net(args.variant, args.mode)
        k2 = max(5, args.steps // 2)
        fam2 = {}
        ms2, _, _, _ = timed_steps(n2, k2, 3, optimizer="ranger21_xx")
        timed_steps(n2, k2, 1, optimizer="ranger21_xx", collect=fam2)
        out[f"{args.mode}/{args.variant}/ranger21_xx"] = {
            "seqs_per_s": 5 * B * k2 / (ms2 / 1e3), "ms_per_step": ms2 / k2, "steps": k2,
            "optimizer_ms_per_step": fam2["ranger21"][0] / k2 if "ranger21" in fam2 else None,
            "optimizer": "intrepppid_b200.optim.FusedRanger21 (ib200_ranger21_step: 3 launches, no host sync; parity unpinned against "
                         "the third-party package, which is absent from the image)"}
        del n2
    except Exception as e:  # noqa: BLE001
        out[f"{args.mode}/{args.variant}/ranger21_xx"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    torch.cuda.empty_cache()
    ns = types.SimpleNamespace(mode=args.mode, proteins=20000, rows=1000000, batch=C5["B"], len=C5["T"], no_train=False)
    for tag, fn in (("config4_inference_20k_proteome_all_pairs", bc.config4), ("from_csv_20k_ragged_proteome_1M_rows", bc.config_csv),
                    ("config5_stress_E256_L3_T4000_B256", bc.config5)):
        try:
            out[tag] = fn(ns)
        except Exception as e:  # noqa: BLE001  (an extra must never cost the headline line)
            out[tag] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()
    ns.mode = "bf16"
    try:
        r = bc.config5(ns)
        out["config5_stress_bf16"] = {k: r[k] for k in ("infer_ms", "train_ms", "peak_mem_GB") if k in r}
    except Exception as e:  # noqa: BLE001
        out["config5_stress_bf16"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    torch.cuda.empty_cache()
    return out


def run_b200_config5(args):
    """BASELINE configs[4] through the same harness: encoder forward + backward + gradient all-reduce + AdamW, 256 sequences of
    4000 tokens per GPU (seed 777 + rank)."""
    import intrepppid_b200 as ib
    from intrepppid_b200 import ops
    from intrepppid_b200.optim import FusedAdamW
    from intrepppid_b200.parallel import GradientAllReducer

    dist, rank, world, local_rank, dev = _dist_setup(args)
    torch.manual_seed(0)
    net = ib.intrepppid_network(1, embedding_size=C5["E"], rnn_num_layers=C5["L"], bi_reduce=C5["bi"], precision=args.mode,
                                optimizer_type="adamw").to(dev).train()
    x = torch.randint(1, V, (C5["B"], C5["T"]), generator=torch.Generator().manual_seed(777 + rank)).to(dev)
    enc_params = [p for p in net.encoder.parameters() if p.requires_grad]
    opt = FusedAdamW(enc_params, lr=1e-3)

    def make_step(reducer):
        def one_step():
            opt.zero_grad(set_to_none=True)
            z = net.encoder(x)
            z.square().mean().backward()
            if reducer is not None:
                reducer.finish()
            opt.step()
        return one_step

    K, W = args.steps, args.warmup
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    reducer = GradientAllReducer(net.encoder) if world > 1 else None
    ms, launches = _timed_region(make_step(reducer), K, W, world, dist, dev)
    clocks = sampler.stop() if rank == 0 else None
    fam = {}
    ms_instr, _ = _timed_region(make_step(reducer), max(1, K // 2), 1, world, dist, dev, collect=fam)
    if reducer is not None:
        reducer.remove()
    ms_nocomm = _timed_region(make_step(None), max(1, K // 2), 1, world, dist, dev)[0] / max(1, K // 2) if world > 1 else None
    ops.check_pending(sync=True)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    k2 = max(1, K // 2)
    tokens = float(C5["B"] * C5["T"])
    H, Lc = C5["E"], C5["L"]
    flops_train = tokens * 3 * (32 * H * H + 48 * H * H * (Lc - 1))  # SURVEY 8d dense count, fwd + bwd (all chains live under mean)
    fams = {k: {"ms_per_step": v[0] / k2, "launcher_calls_per_step": v[1] / k2} for k, v in fam.items()}
    rec = sum(v["ms_per_step"] for k, v in fams.items() if k.startswith("lstm_"))
    gem = sum(v["ms_per_step"] for k, v in fams.items() if k.startswith("gemm_") or k in ("dw_reduce", "l0_grads", "emb_grad"))
    out = {"metric": C5_METRIC, "value": C5["B"] * world * K / (ms / 1e3), "unit": "seqs/s", "n_gpus": world, "steps": K, "warmup": W,
           "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32 state; bf16 hi/lo-split tensor-core products, fp32 accumulate" if args.mode == "fp32" else "bf16 tensor-core products, fp32 accumulate/state",
           "data": "synthetic", "impl": "b200", "mode": args.mode,
           "config": {"workload": C5_WORKLOAD, "batch_per_gpu": C5["B"], "global_batch": C5["B"] * world, "trunc_len": C5["T"], "vocab": V,
                      "embed": C5["E"], "layers": C5["L"], "bi_reduce": C5["bi"], "optimizer": "AdamW step inside the timed step",
                      "parallelism": f"dp{world}"},
           "tokens_per_s": tokens * world * K / (ms / 1e3),
           "useful_tflops_per_gpu": flops_train / (ms / K / 1e3) / 1e12,
           "tensor_frac_useful": flops_train / (ms / K / 1e3) / 1e12 / peaks["bf16_tflops_sustained"],
           "recurrent_ms_per_step": rec, "gemm_ms_per_step": gem,
           "tau_us_per_cell_step": {k: v["ms_per_step"] * 1e3 / C5["T"] / max(1.0, v["launcher_calls_per_step"]) for k, v in fams.items()
                                    if k.startswith("lstm_")},
           "kernel_families": fams, "gpu_launches": int(launches), "clocks": clocks, "peaks": peaks,
           "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}
    if world > 1:
        out["comm_exposed_ms"] = ms / K - ms_nocomm
        dist.destroy_process_group()
    emit(out)


# ------------------------------------------------------------------------------------------------------------------------------
# CPU / reference legs (the only place bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------------------------------------
def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as fh:
            for ln in fh:
                if ln.lower().startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def _reference_step_fn(variant: str, device="cpu"):
    """-> (step(batch) -> loss tensor, kind).  kind "reference": the reference's own TripletE2ENet (the five unmodified files, loaded
    by oracle/ref_shim.py from /root/reference or from the sha256-verified staged copies under oracle/_ref) with its own RNG draws;
    kind "port": oracle/restatement.py (same ATen LSTM entry point as nn.LSTM) when the reference files are not reachable."""
    p_emb = 0.0 if variant == "full_t" else 0.3
    from oracle import ref_shim

    if ref_shim.available():
        torch.manual_seed(0)
        net = ref_shim.build_reference_net(vocab=V, E=E, L=L, bi_reduce=BI, emb_droprate=p_emb).to(device).train()
        opt = torch.optim.AdamW(net.parameters(), lr=1e-3)

        def step(batch):
            opt.zero_grad(set_to_none=True)
            loss = net.step(batch, "train")   # e2e/e2e_triplet.py:113-187, masks drawn by the reference's own F.dropout / bernoulli_
            loss.backward()
            opt.step()
            return loss.detach()

        return step, "reference"

    from oracle import restatement as R

    P = {k: v.clone().to(device).requires_grad_(True) for k, v in R.init_params(vocab=V, E=E, L=L, seed=0).items()}
    opt = torch.optim.AdamW(list(P.values()), lr=1e-3)
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
        return out.loss.detach()

    return step, "port"


def _time_cpu_steps(step, batch, n):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        step(batch)
        ts.append(time.perf_counter() - t0)
    return ts


def _one_thread_figure(step, batch):
    """The same step on ONE host thread, on a bounded sample (the first 8 samples of the batch), scaled to sequences/s."""
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        small = [t[:8] for t in batch]
        step(small)
        t0 = time.perf_counter()
        step(small)
        dt = time.perf_counter() - t0
    finally:
        torch.set_num_threads(n)
    return {"value": 5 * 8 / dt, "unit": "seqs/s", "threads": 1, "sample": f"1 step on the first 8 of the {B} samples after 1 warm-up step; {dt:.2f} s"}


def cpu_baseline(variant: str):
    """BASELINE.md section 3 protocol on this box's host cores: 1 warm-up step, then the median of 3 full-batch steps
    (zero_grad -> step -> backward -> AdamW), all host threads; plus a 1-thread figure on a bounded sample."""
    warnings.filterwarnings("ignore")
    torch.set_num_threads(os.cpu_count())
    step, kind = _reference_step_fn(variant)
    batch = synthetic_batch(1234)
    step([t[:4] for t in batch])
    ts = _time_cpu_steps(step, batch, 3)
    med = statistics.median(ts)
    return {"value": 5 * B / med, "unit": "seqs/s", "cores": os.cpu_count(), "kind": kind, "threads": torch.get_num_threads(),
            "cpu_model": cpu_model(), "torch": torch.__version__, "median_step_s": med, "step_s": [round(t, 3) for t in ts],
            "sample": f"median of 3 full steps (B={B}, T={T}, fwd+bwd+AdamW, fp32, torch CPU LSTM) after a B=4 warm-up step",
            "one_thread": _one_thread_figure(step, batch)}


def gpu_reference(variant: str, dev):
    """The UNMODIFIED reference modules on this same GPU through torch CUDA (nn.LSTM -> cuDNN), same batch, fp32, step + backward
    + AdamW -- the reference as its users run it (the CPU arm is the contract's baseline; this is the same-hardware one).  None when
    the reference files are not reachable (the restatement is not a stand-in for this figure)."""
    from oracle import ref_shim

    if not ref_shim.available():
        return None
    warnings.filterwarnings("ignore")
    try:
        step, kind = _reference_step_fn(variant, device=dev)
        batch = [t.to(dev) for t in synthetic_batch(1234)]
        for _ in range(3):
            step(batch)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        a.record()
        for _ in range(n):
            step(batch)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        return {"value": 5 * B / (ms / 1e3), "unit": "seqs/s", "ms_per_step": ms, "kind": kind, "steps": n, "warmup": 3,
                "how": "oracle/ref_shim.py: the five unmodified reference files, .cuda(), torch " + torch.__version__ +
                       " nn.LSTM (cuDNN; WeightDrop leaves the weights unflattened as in the reference), fp32, same synthetic batch"}
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the step on this box's host cores, all threads, on the b200
    arm's config / metric / unit.  Each step is a FULL workload batch (same config) unless that cannot fit the driver's limit."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    warnings.filterwarnings("ignore")
    torch.set_num_threads(os.cpu_count())
    step, kind = _reference_step_fn(args.variant)
    batch = synthetic_batch(1234)
    K, W = args.steps, args.warmup
    probe = [t[:4] for t in batch]
    step(probe)
    t0 = time.perf_counter()
    step(probe)
    per_sample = (time.perf_counter() - t0) / 4
    budget = 1500.0  # seconds for the K + W steps (the driver's limit for this arm is 1800 s)
    bs = B if (K + W) * per_sample * B <= budget else int(max(4, min(B, budget / ((K + W) * per_sample))))
    sample = batch if bs == B else [t[:bs] for t in batch]
    for _ in range(W):
        step(sample)
    t0 = time.perf_counter()
    ts = _time_cpu_steps(step, sample, K)
    dt = time.perf_counter() - t0
    value = 5 * bs * K / dt
    cfg = workload_config(args.variant, args.gpus)
    cb = {"value": value, "unit": "seqs/s", "cores": os.cpu_count(), "kind": kind, "threads": torch.get_num_threads(),
          "cpu_model": cpu_model(), "torch": torch.__version__, "median_step_s": statistics.median(ts),
          "sample": (f"each step = the full workload batch ({B} samples x 5 sequences, T={T})" if bs == B else
                     f"each step = the first {bs} of the {B} samples of the workload batch (time budget)") +
                    ", fwd+bwd+AdamW, fp32, all host threads",
          "one_thread": _one_thread_figure(step, batch)}
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "seqs/s",
        "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "run": {"sample_batch": bs, "same_config": bs == B,
                "reference": ("the reference's own TripletE2ENet.step + backward (five unmodified files via oracle/ref_shim.py" +
                              (", staged under oracle/_ref" if getattr(sys.modules.get("oracle.ref_shim"), "STAGED", False) else "") + ")")
                if kind == "reference" else "oracle/restatement.py (CPU port; reference files not reachable)"},
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
    ap.add_argument("--workload", default="headline", choices=["headline", "config5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip other_configs (bf16 / full_t / configs 4, 5, from_csv; N = 1 only)")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "config5":
        run_b200_config5(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
