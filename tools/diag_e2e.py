#!/usr/bin/env python
"""Where does the end-to-end loop of bench.py spend its time?  Runs the e2e passes (packed uint8 / int64 loader batches / resident) in
several orders and prints, per pass, the device time per step and the host-side time of every phase of a step."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import intrepppid_b200 as ib  # noqa: E402
from intrepppid_b200 import feed  # noqa: E402
from intrepppid_b200.optim import FusedAdamW  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    net = ib.intrepppid_network(1, precision="fp32", optimizer_type="adamw").to(dev).train()
    host_batch = bench.synthetic_batch(1234)
    dev_batch = [t.to(dev) for t in host_batch]
    params = [p for p in net.parameters() if p.requires_grad]
    K, W = 10, 2

    def run(mode):
        opt = FusedAdamW(params, lr=1e-3)
        feeder = batches = None
        if mode != "resident":
            src = host_batch if mode == "int64" else feed.pack_batch(host_batch[:5], host_batch[5], bench.V)
            feeder = feed.DeviceFeeder((src for _ in range(K + W)), dev, bench.V)
            batches = iter(feeder)
        loss_ev = [torch.cuda.Event() for _ in range(2)]
        loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        st = {"k": 0}
        phases = {"next": [], "step": [], "bwd": [], "opt": [], "loss": []}

        def one():
            t0 = time.perf_counter()
            batch = next(batches) if batches is not None else dev_batch
            t1 = time.perf_counter()
            opt.zero_grad(set_to_none=True)
            loss = net.step(batch, "train")
            t2 = time.perf_counter()
            loss.backward()
            t3 = time.perf_counter()
            opt.step()
            t4 = time.perf_counter()
            k = st["k"]
            loss_host[k % 2:k % 2 + 1].copy_(loss.detach().reshape(1), non_blocking=True)
            loss_ev[k % 2].record()
            if k > 0:
                loss_ev[(k - 1) % 2].synchronize()
            st["k"] = k + 1
            t5 = time.perf_counter()
            for name, a, b in (("next", t0, t1), ("step", t1, t2), ("bwd", t2, t3), ("opt", t3, t4), ("loss", t4, t5)):
                phases[name].append((b - a) * 1e3)

        for _ in range(W):
            one()
        torch.cuda.synchronize()
        for v in phases.values():
            v.clear()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        a.record()
        for _ in range(K):
            one()
        b.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - w0) * 1e3
        print(f"{mode:9s} device {a.elapsed_time(b) / K:.3f} ms/step  wall {wall / K:.3f} ms/step  host phases (ms, mean): " +
              "  ".join(f"{n} {sum(v) / len(v):.3f}" for n, v in phases.items()) +
              "  | per-step next: " + " ".join(f"{x:.2f}" for x in phases["next"]), flush=True)

    for mode in ("resident", "packed", "int64", "packed", "resident", "int64", "packed"):
        run(mode)


if __name__ == "__main__":
    main()
