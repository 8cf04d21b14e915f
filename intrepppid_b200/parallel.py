"""Data-parallel gradient exchange for one-process-per-GPU training.

The path shards by samples: every rank holds a full replica (0.87 MB) and 80 samples; the only exchange is ONE mean-allreduce
of the live gradients (188,161 fp32 = 753 KB) per step, over NCCL/NVLink.  It is latency-bound, so the gradients are packed
into three buckets ordered by backward completion and each bucket is reduced asynchronously as soon as it is complete:
  bucket 0: head (+ triplet projection)  -- the flat buffer ib200_loss_head_bwd wrote; ready before the recurrent BPTT starts
  bucket 1: encoder fc                   -- the flat buffer ib200_pool_fc_bwd wrote; reduced underneath the BPTT kernels too
  bucket 2: LSTM + embedding             -- the flat buffer ib200_encoder_bwd wrote; ready when it returns
Every bucket is exactly one buffer of the kernels, so it is reduced IN PLACE (no torch.cat before, no copy back after), and over
NCCL the mean is taken by the collective itself (ReduceOp.AVG): no extra kernel at all on the path.
The reference has no distributed code at all (devices=1 is hard-wired, e2e_triplet.py:392-400); parameters that never
receive a gradient (encoder.projection.*, quirk Q10) are left out of the buckets.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def default_buckets(module: torch.nn.Module) -> List[List[torch.nn.Parameter]]:
    head, fc, late, seen = [], [], [], set()
    for name, p in module.named_parameters():  # named_parameters de-duplicates the rnn / rnn_dp.module aliases
        if not p.requires_grad or id(p) in seen or ".projection." in name or name.startswith("projection."):
            continue
        seen.add(id(p))
        if ".rnn." in name or ".rnn_dp." in name or "embedder" in name:
            late.append(p)
        elif name.startswith("encoder.") and ".fc." in name:
            fc.append(p)
        else:
            head.append(p)
    return [b for b in (head, fc, late) if b]


class GradientAllReducer:
    def __init__(self, module: torch.nn.Module, buckets: Optional[Sequence[Sequence[torch.nn.Parameter]]] = None,
                 process_group=None):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.buckets = [list(b) for b in (buckets if buckets is not None else default_buckets(module))]
        self._bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b}
        self._pending = [len(b) for b in self.buckets]
        self._flat: List[Optional[torch.Tensor]] = [None] * len(self.buckets)
        self._work = [None] * len(self.buckets)
        self._inplace = [False] * len(self.buckets)
        self._hooks = []
        for b in self.buckets:
            for p in b:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.bytes_per_step = sum(p.numel() * 4 for b in self.buckets for p in b)
        # NCCL averages inside the collective; gloo (CPU tests) has no AVG: sum, then divide
        self._avg_in_collective = dist.is_initialized() and dist.get_backend(process_group) == "nccl"

    def _on_grad(self, p):
        i = self._bucket_of[id(p)]
        self._pending[i] -= 1
        if self._pending[i] == 0:
            self._launch(i)

    @staticmethod
    def _shared_flat(grads):
        """The kernels hand back the gradients of one op as views of ONE flat buffer (ops.py): when a bucket is exactly such a
        buffer it is reduced in place -- no torch.cat before and no copy back after the collective."""
        st = grads[0].untyped_storage()
        if any(g.untyped_storage().data_ptr() != st.data_ptr() or not g.is_contiguous() or g.dtype != torch.float32 for g in grads):
            return None
        if sum(g.numel() for g in grads) * 4 != st.nbytes():
            return None
        return torch.empty(0, dtype=torch.float32, device=grads[0].device).set_(st, 0, (st.nbytes() // 4,))

    def _launch(self, i):
        grads = [p.grad for p in self.buckets[i]]
        base = self._shared_flat(grads)
        self._inplace[i] = base is not None
        flat = base if base is not None else torch.cat([g.reshape(-1) for g in grads])
        self._flat[i] = flat
        if self.world > 1:
            op = dist.ReduceOp.AVG if self._avg_in_collective else dist.ReduceOp.SUM
            self._work[i] = dist.all_reduce(flat, op=op, group=self.group, async_op=True)

    def finish(self):
        """Wait for the outstanding reductions and write the averaged gradients back.  Call after loss.backward()."""
        for i, b in enumerate(self.buckets):
            if self._flat[i] is None:
                if self._pending[i] != len(b):
                    raise RuntimeError(f"gradient bucket {i} is incomplete: {self._pending[i]} of {len(b)} gradients missing")
                continue  # backward did not touch this bucket at all (e.g. frozen part)
            if self._work[i] is not None:
                self._work[i].wait()
            flat = self._flat[i]
            if self.world > 1 and not self._avg_in_collective:
                flat.div_(self.world)
            if not self._inplace[i]:  # (in-place buckets: the gradients ARE views of `flat`)
                off = 0
                for p in b:
                    n = p.numel()
                    p.grad.copy_(flat[off:off + n].view_as(p.grad))
                    off += n
            self._flat[i], self._work[i] = None, None
            self._pending[i] = len(b)

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n_items for `rank` (inference: proteins are independent)."""
    per, rem = divmod(n_items, world)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


def triangle_row_start(M: int, i: int) -> int:
    """Flat index of pair (i, i) in the row-major upper triangle of an M x M pair matrix (row i holds M - i pairs)."""
    return i * M - i * (i - 1) // 2


def pair_block(M: int, rank: int, world: int):
    """(row_lo, row_hi, p_begin, p_count): the block of whole triangle ROWS scored by `rank`, balanced by pair count (row i has
    M - i pairs, so equal row counts would give rank 0 almost twice the average).  Blocks are contiguous in the flat index."""
    total = M * (M + 1) // 2
    bounds = [0]
    for r in range(1, world):
        target = total * r // world
        lo, hi = bounds[-1], M
        while lo < hi:  # first row whose start index reaches the target
            mid = (lo + hi) // 2
            if triangle_row_start(M, mid) < target:
                lo = mid + 1
            else:
                hi = mid
        bounds.append(lo)
    bounds.append(M)
    r0, r1 = bounds[rank], bounds[rank + 1]
    p0 = triangle_row_start(M, r0)
    return r0, r1, p0, triangle_row_start(M, r1) - p0


@torch.no_grad()
def sharded_proteome_scores(net, tokens: torch.Tensor, batch_size: int = 512, process_group=None):
    """Multi-GPU inference of BASELINE config 4: every rank embeds its contiguous shard of the M proteins, the [M,E]
    embedding matrix is all-gathered (M*E*4 bytes, 5 MB at M=20k), and each rank scores its block of triangle rows.
    tokens: the FULL [M, trunc_len] id matrix (every rank indexes its own shard).  Returns (z_all [M,E], p_begin, probs_of_block)."""
    world = dist.get_world_size(process_group) if dist.is_initialized() else 1
    rank = dist.get_rank(process_group) if dist.is_initialized() else 0
    M = tokens.shape[0]
    lo, hi = shard_range(M, rank, world)
    dev = next(net.parameters()).device
    z_local = net.embed(tokens[lo:hi].to(dev), batch_size) if hi > lo else torch.empty(0, net.encoder.encoder.embedding_size, device=dev)
    if world > 1:
        per = (M + world - 1) // world  # equal-size slots for all_gather_into_tensor; the tail of a short shard is padding
        slot = torch.zeros(per, z_local.shape[1], dtype=z_local.dtype, device=dev)
        slot[: hi - lo] = z_local
        gathered = torch.empty(world * per, z_local.shape[1], dtype=z_local.dtype, device=dev)
        dist.all_gather_into_tensor(gathered, slot, group=process_group)
        z_all = torch.cat([gathered[r * per: r * per + (shard_range(M, r, world)[1] - shard_range(M, r, world)[0])] for r in range(world)])
    else:
        z_all = z_local
    _, _, p0, pc = pair_block(M, rank, world)
    return z_all, p0, net.score_pairs_range(z_all, p0, pc)
