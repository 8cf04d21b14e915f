"""How long does the HOST take to enqueue one training step (python + ctypes + torch launches), vs the device time of the step?
    python tools/host_overhead.py [fp32|bf16]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import intrepppid_b200 as ib

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
torch.manual_seed(0)
net = ib.intrepppid_network(1, precision=mode, optimizer_type="adamw").cuda().train()
net.encoder.check_lengths = False
batch = [t.cuda() for t in bench.synthetic_batch(1234)]
opt = torch.optim.AdamW([p for p in net.parameters() if p.requires_grad], lr=1e-3, fused=True)


def step():
    opt.zero_grad(set_to_none=True)
    loss = net.step(batch, "train")
    loss.backward()
    opt.step()


for _ in range(5):
    step()
torch.cuda.synchronize()
K = 20
t0 = time.perf_counter()
for _ in range(K):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"{mode}: host enqueue {1e3 * (t1 - t0) / K:.3f} ms/step, wall incl. device {1e3 * (t2 - t0) / K:.3f} ms/step")
# phases of the host side
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
