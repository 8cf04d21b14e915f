"""Data-parallel gradient exchange for one-process-per-GPU training.

The path shards by samples: every rank holds a full replica (0.87 MB) and 80 samples; the only exchange is ONE mean-allreduce
of the live gradients (188,161 fp32 = 753 KB) per step, over NCCL/NVLink.  It is latency-bound, so the gradients are packed
into four buckets ordered by backward completion and each bucket is reduced asynchronously as soon as it is complete:
  bucket 0: head (+ triplet projection)  -- the flat buffer ib200_loss_head_bwd wrote; ready before the recurrent BPTT starts
  bucket 1: encoder fc                   -- the flat buffer ib200_pool_fc_bwd wrote; reduced underneath the BPTT kernels too
  bucket 2: LSTM layers >= 1             -- the tail of the flat buffer of the encoder backward; final once the upper layers' BPTT and
                                            weight-gradient GEMMs are done.  The encoder backward is cut at the layer boundary
                                            (ib200_encoder_bwd_layers) and calls ops.EARLY_GRAD_HOOKS in between, so this bucket -- 80 %
                                            of all gradient bytes at L = 2 -- is all-reduced UNDER the layer-0 BPTT kernel
  bucket 3: LSTM layer 0 + embedding     -- the head of the same flat buffer; ready when the backward returns (the only exposed part)
Every bucket is one contiguous range of a kernel-written buffer, so it is reduced IN PLACE (no torch.cat before, no copy back after),
and over NCCL the mean is taken by the collective itself (ReduceOp.AVG): no extra kernel at all on the path.
The reference has no distributed code at all (devices=1 is hard-wired, e2e_triplet.py:392-400); parameters that never
receive a gradient (encoder.projection.*, quirk Q10) are left out of the buckets.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import re

import torch
import torch.distributed as dist

from . import ops


_UPPER_LAYER = re.compile(r"_l([1-9][0-9]*)(_reverse)?(_raw)?$")


def default_buckets(module: torch.nn.Module) -> List[List[torch.nn.Parameter]]:
    head, fc, upper, late, seen = [], [], [], [], set()
    for name, p in module.named_parameters():  # named_parameters de-duplicates the rnn / rnn_dp.module aliases
        if not p.requires_grad or id(p) in seen or ".projection." in name or name.startswith("projection."):
            continue
        seen.add(id(p))
        if ".rnn." in name or ".rnn_dp." in name:
            (upper if _UPPER_LAYER.search(name) else late).append(p)
        elif "embedder" in name:
            late.append(p)
        elif name.startswith("encoder.") and ".fc." in name:
            fc.append(p)
        else:
            head.append(p)
    return [b for b in (head, fc, upper, late) if b]


class _EventWork:
    """`.wait()` of a reduction that ran on the reducer's own stream: the caller's stream waits for its event (no host sync)."""

    def __init__(self, event: "torch.cuda.Event"):
        self.event = event

    def wait(self):
        torch.cuda.current_stream().wait_event(self.event)


class P2PAllReduce:
    """One-shot mean all-reduce of small buckets over NVLink peer memory (`ib200_p2p_allreduce_mean`, csrc/p2p.cu).

    Every rank allocates, per bucket, a staging tensor of two parity halves and a flag row; the CUDA IPC handles are exchanged once
    through the process group (`all_gather_object`) and every rank maps the regions of all its peers.  A reduction is then two
    launches on this rank -- stage, then publish / wait / sum the peers' copies straight out of their memory -- instead of a ring or
    tree of sends: the exchange of the training step is latency bound (753 KB in four buckets).  All ranks must sit on one node with
    peer access (NVLink / NVSwitch) and issue the same sequence of reductions per bucket."""

    MAX_WORLD = 8

    def __init__(self, bucket_numels: Sequence[int], device, process_group=None):
        from . import _lib

        self._lib = _lib
        self.group = process_group
        self.world, self.rank = dist.get_world_size(process_group), dist.get_rank(process_group)
        if self.world > self.MAX_WORLD:
            raise ValueError(f"P2PAllReduce supports up to {self.MAX_WORLD} ranks")
        self.device = torch.device(device)
        self.stage_floats = [(int(n) + 3) // 4 * 4 for n in bucket_numels]
        # ONE peer-mappable allocation per rank: [bucket 0: 2 parity halves][bucket 1: ...] ... [flags: buckets x 8 words], zeroed
        import ctypes as C

        offs, off = [], 0
        for n in self.stage_floats:
            offs.append(off)
            off += (2 * n * 4 + 255) // 256 * 256
        flag_off, total = off, off + len(bucket_numels) * self.MAX_WORLD * 4
        L = _lib.lib()
        with torch.cuda.device(self.device):
            base, handle = C.c_void_p(), C.create_string_buffer(64)
            _lib.check(L.ib200_p2p_alloc(total, C.byref(base), handle), "ib200_p2p_alloc")
            everyone = [None] * self.world
            dist.all_gather_object(everyone, handle.raw, group=process_group)
            self._base, self._opened, bases = base, [], []
            for r, h in enumerate(everyone):
                if r == self.rank:
                    bases.append(base.value)
                    continue
                p = C.c_void_p()
                _lib.check(L.ib200_p2p_open(h, C.byref(p)), "ib200_p2p_open")
                self._opened.append(p)
                bases.append(p.value)
        arr = C.c_void_p * self.world
        self._stage_arr = [arr(*[b + o for b in bases]) for o in offs]
        self._flag_arr = [arr(*[b + flag_off + 4 * self.MAX_WORLD * i for b in bases]) for i in range(len(bucket_numels))]
        self._epoch = [0] * len(bucket_numels)
        self.stream = torch.cuda.Stream(self.device)
        dist.barrier(group=process_group)  # nobody publishes before everybody has mapped everything

    def close(self):
        """Unmap the peers' regions and free mine (after a barrier: nobody may still be reading)."""
        if getattr(self, "_base", None) is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        L = self._lib.lib()
        for p in self._opened:
            L.ib200_p2p_close(p)
        L.ib200_p2p_free(self._base)
        self._base, self._opened = None, []

    def all_reduce_mean_(self, bucket: int, flat: torch.Tensor) -> _EventWork:
        """Average `flat` (fp32, contiguous, at most the bucket's size) over all ranks, in place, ordered after the work already on
        the current stream; returns a handle whose .wait() orders the current stream after the reduction."""
        if flat.dtype != torch.float32 or not flat.is_contiguous() or flat.numel() > self.stage_floats[bucket]:
            raise ValueError("P2PAllReduce: fp32 contiguous tensor of at most the bucket's size expected")
        self._epoch[bucket] += 1
        ready = torch.cuda.Event()
        ready.record()
        self.stream.wait_event(ready)
        self._lib.check(self._lib.lib().ib200_p2p_allreduce_mean(self.world, self.rank, self._stage_arr[bucket], self._flag_arr[bucket],
                                                                 self.stage_floats[bucket], flat.data_ptr(), flat.numel(),
                                                                 self._epoch[bucket] & 0xFFFFFFFF or 1, self.stream.cuda_stream),
                        "ib200_p2p_allreduce_mean")
        flat.record_stream(self.stream)
        done = torch.cuda.Event()
        done.record(self.stream)
        return _EventWork(done)


def _p2p_wanted(total_bytes: int, process_group) -> bool:
    """IB200_ALLREDUCE = nccl (default) | p2p: the peer-memory reduction is opt-in.  Measured on 8 B200 (profiles/r2_p2p_ab.txt) the
    NCCL collectives launched from the gradient hooks leave 0.14 ms of the exchange exposed per step; see there for the p2p figure."""
    import os

    mode = os.environ.get("IB200_ALLREDUCE", "nccl").lower()
    if mode != "p2p" or not dist.is_initialized() or dist.get_backend(process_group) != "nccl":
        return False
    world = dist.get_world_size(process_group)
    return 2 <= world <= P2PAllReduce.MAX_WORLD


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
        # the bucket of the LSTM layers >= 1 can start mid-backward (ops.EARLY_GRAD_HOOKS); found by name, only in the default layout
        self._early = None
        if buckets is None:
            names = {id(p): n for n, p in module.named_parameters()}
            for i, b in enumerate(self.buckets):
                if all((".rnn." in names[id(p)] or ".rnn_dp." in names[id(p)]) and _UPPER_LAYER.search(names[id(p)]) for p in b):
                    self._early = i
            if self._early is not None:
                ops.EARLY_GRAD_HOOKS.append(self._on_early)
        self._early_done = False
        # NCCL averages inside the collective; gloo (CPU tests) has no AVG: sum, then divide
        self._avg_in_collective = dist.is_initialized() and dist.get_backend(process_group) == "nccl"
        # small exchanges between the GPUs of one node: one-shot reduction over NVLink peer memory instead of the NCCL collective
        self._p2p = None
        if self.world > 1 and _p2p_wanted(self.bytes_per_step, process_group):
            dev = next(p.device for b in self.buckets for p in b)
            self._p2p = P2PAllReduce([sum(p.numel() for p in b) for b in self.buckets], dev, process_group)
            self._avg_in_collective = True  # (the kernel writes the mean)

    def _on_grad(self, p):
        i = self._bucket_of[id(p)]
        self._pending[i] -= 1
        if self._pending[i] == 0:
            if i == self._early and self._early_done:
                self._check_early_views(i)
            else:
                self._launch(i)

    def _on_early(self, upper: torch.Tensor):
        """Called from inside the encoder backward (ops._EncodeHidden.backward) with the flat slice that holds the final gradients of
        the LSTM layers >= 1, before the layer-0 BPTT is enqueued: the collective runs underneath that kernel."""
        i = self._early
        b = self.buckets[i]
        if self._early_done or self._flat[i] is not None or upper.numel() != sum(p.numel() for p in b):
            return
        if any(p.grad is not None for p in b):
            return  # gradient accumulation into existing .grad tensors: the regular (post-accumulate) path handles it
        self._early_done = True
        self._inplace[i] = True
        self._flat[i] = upper
        if self.world > 1:
            self._work[i] = self._reduce(i, upper)

    def _check_early_views(self, i):
        """The early bucket was reduced in place inside `flat`; autograd normally hands those very views to .grad.  If it copied
        instead (not the case for freshly created gradients), finish() writes the reduced values back."""
        flat = self._flat[i]
        lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
        self._inplace[i] = all(lo <= p.grad.data_ptr() < hi for p in self.buckets[i])

    @staticmethod
    def _shared_flat(grads):
        """The kernels hand back the gradients of one op as views of ONE flat buffer (ops.py): when a bucket is a contiguous range
        of such a buffer (in any order, without gaps) it is reduced in place -- no torch.cat before, no copy back after."""
        st = grads[0].untyped_storage()
        if any(g.untyped_storage().data_ptr() != st.data_ptr() or not g.is_contiguous() or g.dtype != torch.float32 for g in grads):
            return None
        spans = sorted((g.storage_offset(), g.numel()) for g in grads)
        pos = spans[0][0]
        for off, n in spans:
            if off != pos:
                return None
            pos += n
        return torch.empty(0, dtype=torch.float32, device=grads[0].device).set_(st, spans[0][0], (pos - spans[0][0],))

    def _launch(self, i):
        grads = [p.grad for p in self.buckets[i]]
        base = self._shared_flat(grads)
        self._inplace[i] = base is not None
        flat = base if base is not None else torch.cat([g.reshape(-1) for g in grads])
        self._flat[i] = flat
        if self.world > 1:
            self._work[i] = self._reduce(i, flat)

    def _reduce(self, i, flat):
        if self._p2p is not None:
            return self._p2p.all_reduce_mean_(i, flat)
        op = dist.ReduceOp.AVG if self._avg_in_collective else dist.ReduceOp.SUM
        return dist.all_reduce(flat, op=op, group=self.group, async_op=True)

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
                for p in (self._early_order(b) if i == self._early and self._early_done else b):
                    n = p.numel()
                    p.grad.copy_(flat[off:off + n].view_as(p.grad))
                    off += n
            self._flat[i], self._work[i] = None, None
            self._pending[i] = len(b)
        self._early_done = False

    @staticmethod
    def _early_order(b):
        return b  # named_parameters order of the layers >= 1 == the C-ABI order of the flat buffer (ops.lstm_param_order)

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
        if self._on_early in ops.EARLY_GRAD_HOOKS:
            ops.EARLY_GRAD_HOOKS.remove(self._on_early)
        if self._p2p is not None:
            self._p2p.close()
            self._p2p = None


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
