"""Real-NCCL check of the sharded inference path (BASELINE config 4); run on a multi-GPU box, not collected by pytest:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/infer_check.py

Every rank embeds its shard of the proteins, the embeddings are all-gathered, every rank scores its block of triangle rows; the
blocks are gathered on rank 0 and compared, bit for bit, with the single-process embed + score_pairs."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import intrepppid_b200 as ib
    from intrepppid_b200.parallel import pair_block, sharded_proteome_scores

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl")
    M, T = int(os.environ.get("IB200_M", 4000)), 1500
    torch.manual_seed(0)
    net = ib.intrepppid_network(1).cuda().eval()
    net.encoder.check_lengths = False
    x = torch.randint(1, 250, (M, T), generator=torch.Generator().manual_seed(4321))
    sharded_proteome_scores(net, x[:64 * world], 32)  # warm-up (NCCL, kernels)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    z_all, p0, probs = sharded_proteome_scores(net, x, 512)
    torch.cuda.synchronize()
    dist.barrier()
    dt = time.perf_counter() - t0
    sizes = [pair_block(M, r, world)[3] for r in range(world)]
    if rank == 0:
        parts = [probs] + [torch.empty(n, device="cuda") for n in sizes[1:]]
        for r in range(1, world):
            dist.recv(parts[r], src=r)
        with torch.no_grad():
            z_ref = net.embed(x.cuda(), 512)
            ref = net.score_pairs(z_ref)
        ok = torch.equal(z_all, z_ref) and torch.equal(torch.cat(parts), ref)
        print(f"infer_check world={world}: M={M} proteins, {ref.numel()} pairs in {dt * 1e3:.1f} ms "
              f"({M / dt:.0f} proteins/s incl. all-gather and scoring) -> {'OK' if ok else 'MISMATCH'}")
        assert ok
    else:
        dist.send(probs, dst=0)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
