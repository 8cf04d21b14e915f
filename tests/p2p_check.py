"""Real-NVLink check of the peer-memory all-reduce (run on a multi-GPU box, not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tests/p2p_check.py

Every rank reduces random buckets of several sizes / alignments many times with intrepppid_b200.parallel.P2PAllReduce and with NCCL
(all_reduce AVG) and compares; then times both on the bucket sizes of the training step."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from intrepppid_b200.parallel import P2PAllReduce

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    sizes = [6209, 4160, 148480, 29312 + 1, 1, 7]  # head | fc | upper LSTM layers | layer 0 + embedding (+1: odd) | tiny
    red = P2PAllReduce(sizes, dev)
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    worst = 0.0
    for it in range(40):
        for b, n in enumerate(sizes):
            base = torch.randn(n + 3, generator=g, device=dev)
            x = base[(it + b) % 4:][:n]  # unaligned views too
            ref = x.clone()
            dist.all_reduce(ref, op=dist.ReduceOp.AVG)
            red.all_reduce_mean_(b, x).wait()
            worst = max(worst, float((x - ref).abs().max()))
    torch.cuda.synchronize()
    # identical on every rank (fixed summation order)
    x = torch.randn(sizes[2], generator=g, device=dev)
    red.all_reduce_mean_(2, x).wait()
    gathered = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(gathered, x)
    same = all(torch.equal(gathered[0], t) for t in gathered)

    def timeit(fn, reps=200):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e3

    bufs = [torch.randn(n, device=dev) for n in sizes[:4]]
    t_p2p = timeit(lambda: [red.all_reduce_mean_(b, t).wait() for b, t in enumerate(bufs)])
    t_nccl = timeit(lambda: [dist.all_reduce(t, op=dist.ReduceOp.AVG) for t in bufs])
    if rank == 0:
        print(f"p2p_check world={world}: worst |p2p - nccl| {worst:.3e}; bit-identical across ranks: {same}; "
              f"four training-step buckets back to back: p2p {t_p2p:.1f} us, NCCL {t_nccl:.1f} us -> {'OK' if worst < 1e-5 and same else 'FAIL'}")
    red.close()
    dist.destroy_process_group()
    if not (worst < 1e-5 and same):
        sys.exit(1)


if __name__ == "__main__":
    main()
