"""world_size-2 gloo tests (CPU) of the data-parallel host logic: bucketed gradient mean-allreduce launched from grad hooks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class Toy(nn.Module):
    """Same naming shape as the product net: head, encoder fc and (rnn, embedder) buckets, plus a dead branch."""

    def __init__(self):
        super().__init__()
        self.encoder = nn.Module()
        self.encoder.embedder = nn.Embedding(10, 4)
        self.encoder.rnn = nn.Linear(4, 4)
        self.encoder.fc = nn.Linear(4, 4)
        self.encoder.projection = nn.Linear(4, 4)  # never used in forward (quirk Q10)
        self.head = nn.Linear(4, 1)

    def forward(self, x):
        return self.head(self.encoder.fc(torch.tanh(self.encoder.rnn(self.encoder.embedder(x).mean(1))))).mean()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from intrepppid_b200.parallel import GradientAllReducer, default_buckets, shard_range

    torch.manual_seed(0)
    net = Toy()
    buckets = default_buckets(net)
    names = {id(p): n for n, p in net.named_parameters()}
    assert len(buckets) == 3  # head | encoder fc | rnn + embedder: one flat kernel buffer each in the product
    assert all("projection" not in names[id(p)] for b in buckets for p in b)
    assert all(names[id(p)].startswith("head.") for p in buckets[0]) and all(".fc." in names[id(p)] for p in buckets[1])
    assert all(("rnn" in names[id(p)] or "embedder" in names[id(p)]) for p in buckets[2])
    red = GradientAllReducer(net)
    g = torch.Generator().manual_seed(100)
    x_all = torch.randint(0, 10, (world * 3, 5), generator=g)
    lo, hi = shard_range(x_all.shape[0], rank, world)
    for _ in range(2):  # two steps: the reducer must re-arm itself
        net.zero_grad(set_to_none=True)
        net(x_all[lo:hi]).backward()
        red.finish()
    mine = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
    # single-process reference: mean over ranks of per-shard gradients == gradient of the mean of shard losses
    torch.manual_seed(0)
    ref = Toy()
    loss = sum(ref(x_all[shard_range(x_all.shape[0], r, world)[0]:shard_range(x_all.shape[0], r, world)[1]]) for r in range(world)) / world
    loss.backward()
    ok = all(torch.allclose(mine[n], p.grad, atol=1e-6) for n, p in ref.named_parameters() if p.grad is not None)
    ok = ok and net.encoder.projection.weight.grad is None
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_bucketed_allreduce_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_range_partitions_exactly():
    from intrepppid_b200.parallel import shard_range

    for n in (0, 1, 7, 20000):
        for w in (1, 2, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_pair_block_partitions_the_triangle_and_balances_pairs():
    """Multi-GPU inference: blocks of whole triangle rows, contiguous in the flat pair index, balanced by pair count."""
    from intrepppid_b200.parallel import pair_block, triangle_row_start

    for M in (1, 2, 7, 513, 20000):
        total = M * (M + 1) // 2
        assert triangle_row_start(M, M) == total
        for w in (1, 2, 3, 8):
            blocks = [pair_block(M, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == M and blocks[0][2] == 0
            assert sum(b[3] for b in blocks) == total
            for a, b in zip(blocks, blocks[1:]):
                assert a[1] == b[0] and a[2] + a[3] == b[2]
            if M >= 513:  # no rank is more than one row of pairs away from the ideal share
                assert max(b[3] for b in blocks) - total // w <= M


# ---- the early bucket: upper LSTM layers all-reduced from INSIDE the encoder backward (ops.EARLY_GRAD_HOOKS) ---------------------------
class _FlatGradLinear(torch.autograd.Function):
    """Stand-in for ops._EncodeHidden: its backward writes the gradients of (l0 weight, l1 weight) into ONE flat buffer [l0 | l1],
    calls the early hooks with the final l1 slice before "running layer 0", and returns views of the buffer -- exactly the
    protocol of the CUDA op (intrepppid_b200/ops.py)."""

    @staticmethod
    def forward(ctx, x, w0, w1):
        ctx.save_for_backward(x, w0, w1)
        return torch.tanh(x @ w0.t()) @ w1.t()

    @staticmethod
    def backward(ctx, dy):
        from intrepppid_b200 import ops

        x, w0, w1 = ctx.saved_tensors
        h = torch.tanh(x @ w0.t())
        flat = torch.empty(w0.numel() + w1.numel())
        flat[w0.numel():] = (dy.t() @ h).reshape(-1)          # "layers >= 1" are final first
        for hook in list(ops.EARLY_GRAD_HOOKS):
            hook(flat[w0.numel():])
        dh = (dy @ w1) * (1 - h * h)
        flat[:w0.numel()] = (dh.t() @ x).reshape(-1)          # then "layer 0"
        return None, flat[:w0.numel()].view_as(w0), flat[w0.numel():].view_as(w1)


class ToyLayers(nn.Module):
    def __init__(self):
        super().__init__()
        self.encoder = nn.Module()
        self.encoder.encoder = nn.Module()
        self.encoder.encoder.rnn = nn.Module()
        self.encoder.encoder.rnn.weight_ih_l0 = nn.Parameter(torch.randn(6, 4) * 0.3)
        self.encoder.encoder.rnn.weight_ih_l1 = nn.Parameter(torch.randn(3, 6) * 0.3)
        self.head = nn.Linear(3, 1)

    def forward(self, x):
        r = self.encoder.encoder.rnn
        return self.head(_FlatGradLinear.apply(x, r.weight_ih_l0, r.weight_ih_l1)).mean()


def _worker_early(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from intrepppid_b200 import ops
    from intrepppid_b200.parallel import GradientAllReducer, default_buckets

    torch.manual_seed(0)
    net = ToyLayers()
    names = {id(p): n for n, p in net.named_parameters()}
    buckets = default_buckets(net)
    assert [sorted(names[id(p)].split(".")[-1] for p in b) for b in buckets] == [["bias", "weight"], ["weight_ih_l1"], ["weight_ih_l0"]]
    red = GradientAllReducer(net)
    assert red._early == 1 and ops.EARLY_GRAD_HOOKS == [red._on_early]
    x_all = torch.randn(world * 3, 4, generator=torch.Generator().manual_seed(5))
    seen = []
    for step in range(3):  # the early path must re-arm itself; step 2 accumulates into existing .grad (regular path takes over)
        if step < 2:
            net.zero_grad(set_to_none=True)
        else:
            keep = {n: p.grad.clone() for n, p in net.named_parameters()}
        net(x_all[rank * 3:(rank + 1) * 3]).backward()
        seen.append(red._early_done)
        red.finish()
    mine = {n: p.grad.clone() for n, p in net.named_parameters()}
    torch.manual_seed(0)
    ref = ToyLayers()
    (sum(ref(x_all[r * 3:(r + 1) * 3]) for r in range(world)) / world).backward()
    want = {n: p.grad for n, p in ref.named_parameters()}
    ok = seen == [True, True, False]
    # steps 0 / 1 gave the averaged gradient g; step 2 accumulated a LOCAL gradient onto it and averaged the sum: mean_r(g + g_r) = 2 g
    ok = ok and all(torch.allclose(keep[n], want[n], atol=1e-6) for n in want)
    ok = ok and all(torch.allclose(mine[n], 2 * want[n], atol=1e-6) for n in want)
    red.remove()
    ok = ok and ops.EARLY_GRAD_HOOKS == []
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_early_bucket_is_reduced_from_inside_the_backward():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_early, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
