"""Real-NCCL data-parallel equivalence check (run on a multi-GPU box, not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dp_check.py

Every rank runs one training step on its shard (same weights, same weight masks, its own slice of the per-sample masks) and the
gradients are mean-all-reduced by GradientAllReducer; rank 0 then runs the same step on the CONCATENATED batch in one process and
compares.  embedding_droprate is 0 here: with row dropout the reference's training-mode truncation (quirk Q2) makes T_eff depend
on which sequences share a batch, so sharding legitimately changes the function."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import intrepppid_b200 as ib
    from intrepppid_b200 import StepMasks
    from intrepppid_b200.parallel import GradientAllReducer

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl")
    Bs, T, V, E = 24, 300, 250, 64
    B = Bs * world
    g = torch.Generator().manual_seed(77)
    toks = [torch.randint(1, V, (B, T), generator=g) for _ in range(5)]
    for t in toks:
        t[3, 200:] = 0
        t[B - 2, 50:] = 0
    y = torch.randint(0, 2, (B,), generator=g)
    whh = (torch.rand(5, 4 * E, E, generator=g) >= 0.3).float() / 0.7
    fc1 = (torch.rand(E // 2, E, generator=g) >= 0.3).float() / 0.7
    fc2 = (torch.rand(1, E // 2, generator=g) >= 0.3).float() / 0.7
    do1 = (torch.rand(B, E // 2, generator=g) >= 0.3).float() / 0.7
    do2 = (torch.rand(B, E // 2, generator=g) >= 0.3).float() / 0.7

    def run(lo, hi, reduce):
        torch.manual_seed(0)
        net = ib.intrepppid_network(1, embedding_droprate=0.0, optimizer_type="adamw").cuda().train()
        red = GradientAllReducer(net) if reduce else None
        batch = [t[lo:hi].cuda() for t in toks] + [y[lo:hi].cuda()]
        m = StepMasks(None, whh.cuda(), (fc1.cuda(), do1[lo:hi].cuda(), do2[lo:hi].cuda(), fc2.cuda()))
        loss = net.step(batch, "train", masks=m)
        loss.backward()
        if red is not None:
            red.finish()
        torch.cuda.synchronize()
        return float(loss), {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}

    loss_r, grads_r = run(rank * Bs, (rank + 1) * Bs, True)
    lt = torch.tensor([loss_r], device="cuda")
    dist.all_reduce(lt)
    ok = True
    if rank == 0:
        loss_f, grads_f = run(0, B, False)
        worst = 0.0
        for n, gf in grads_f.items():
            nf = float(gf.norm())
            if nf < 1e-7:
                continue
            worst = max(worst, float((grads_r[n] - gf).norm()) / nf)
        dl = abs(float(lt) / world - loss_f)
        ok = worst < 1e-4 and dl < 1e-5
        print(f"dp_check world={world}: mean shard loss {float(lt) / world:.7f} vs full-batch {loss_f:.7f}; worst grad rel err {worst:.2e}  -> {'OK' if ok else 'FAIL'}")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
