"""Input feeding for the training step -- drop-in around the reference's dataloaders.

Reference: `IntrepppidDataModule.train_dataloader` (data/ppi_oma.py:611-620) is a plain `DataLoader(dataset, batch_size, num_workers=4,
shuffle=True)`; every sample is five int64 id rows of width trunc_len plus a label (data/ppi_oma.py:457-503), default-collated to
5 x int64 [B, trunc_len] -- 4.8 MB per batch of 80 over worker IPC and over PCIe, copied synchronously by Lightning.

This module keeps the reference's batch format at the consumer (a 6-tuple in the order p1, p2, anchor, positive, negative, y) and
changes how it gets to the GPU:
  * `narrow_collate(vocab_size)`   a `collate_fn` that packs the five id rows of every sample into ONE [5, B, T] tensor of the
                                    narrowest id type the vocabulary allows (uint8 for the manuscript's V = 250: 8x fewer bytes
                                    through worker IPC and PCIe); ids are range-checked while packing, like F.embedding would;
  * `DeviceFeeder(loader, device)` iterates any loader of such batches (or of the reference's default-collated 6-tuples) and yields
                                    DEVICE batches: one pinned staging buffer + one device buffer per pipeline slot, the H2D copy of
                                    batch k+1 issued on a copy stream while the caller's stream computes batch k, no host sync;
  * `feeding_dataloader(...)`       the reference's DataLoader call with both plugged in.
The kernels read the narrow ids directly (ib200_cfg.token_dtype), so nothing is widened on the device either.
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Optional, Sequence

import torch


def token_dtype_for(vocab_size: int) -> torch.dtype:
    """Narrowest id type of the C ABI (IB200_TOK_*) that holds ids in [0, vocab_size)."""
    if vocab_size <= 256:
        return torch.uint8
    if vocab_size <= 32768:
        return torch.int16
    return torch.int32


class PackedBatch(tuple):
    """The reference's batch tuple (p1, p2, anchor, positive, negative, y) whose five id tensors are views of ONE [5, B, T] tensor
    (`.tokens`, same order).  It IS a 6-tuple, so `TripletE2ENet.step` and user code unpack it unchanged."""

    tokens: torch.Tensor

    def __new__(cls, tokens: torch.Tensor, y: torch.Tensor):
        self = super().__new__(cls, (tokens[0], tokens[1], tokens[2], tokens[3], tokens[4], y))
        self.tokens = tokens
        return self

    def __reduce__(self):  # worker -> main process transfer of a DataLoader (tuple subclasses pickle by their items otherwise)
        return (PackedBatch, (self.tokens, self[5]))

    @property
    def y(self) -> torch.Tensor:
        return self[5]


def pack_batch(seqs: Sequence[torch.Tensor], y: torch.Tensor, vocab_size: int, out: Optional[torch.Tensor] = None,
               check_range: bool = True) -> PackedBatch:
    """Five [B, T] id tensors (any integer type) -> PackedBatch of the narrow type.  `out`: optional [5, B, T] destination
    (e.g. a pinned staging buffer).  Ids outside [0, vocab_size) raise IndexError here -- the narrowing would otherwise wrap them."""
    if len(seqs) != 5:
        raise ValueError("a training batch holds five id tensors: p1, p2, anchor, positive, negative (data/ppi_oma.py:500-503)")
    B, T = seqs[0].shape
    dt = token_dtype_for(vocab_size)
    if out is None:
        out = torch.empty(5, B, T, dtype=dt)
    elif tuple(out.shape) != (5, B, T) or out.dtype != dt:
        raise ValueError(f"staging buffer must be {dt} [5, {B}, {T}], got {out.dtype} {tuple(out.shape)}")
    for i, s in enumerate(seqs):
        if tuple(s.shape) != (B, T):
            raise ValueError("the five id tensors of a batch must have the same [B, trunc_len] shape")
        if s.dtype.is_floating_point or s.dtype == torch.bool:
            raise TypeError(f"token ids must be integers, got {s.dtype}")
        if check_range and s.dtype != dt and s.numel():
            lo, hi = int(s.min()), int(s.max())
            if lo < 0 or hi >= vocab_size:
                raise IndexError(f"token id out of range: ids must lie in [0, {vocab_size}), got [{lo}, {hi}]")
        out[i].copy_(s)
    return PackedBatch(out, y.to(torch.int64).reshape(-1))


def narrow_collate(vocab_size: int):
    """collate_fn for `DataLoader(IntrepppidDataset(...))`: samples are (p1, p2, anchor, positive, negative, label) of int64 rows
    (data/ppi_oma.py:489-503); returns a PackedBatch of narrow ids.  Runs inside the worker processes."""
    dt = token_dtype_for(vocab_size)

    def collate(samples) -> PackedBatch:
        B = len(samples)
        if B == 0 or len(samples[0]) != 6:
            raise ValueError("expected samples of (p1, p2, anchor, positive, negative, label); build the dataset with negative_omid=True")
        T = int(torch.as_tensor(samples[0][0]).shape[0])
        out = torch.empty(5, B, T, dtype=dt)
        wide = torch.empty(B, T, dtype=torch.int64)
        for i in range(5):
            for b, smp in enumerate(samples):
                row = torch.as_tensor(smp[i])
                if row.shape[0] != T:
                    raise ValueError("every id row must be padded to trunc_len (data/ppi_oma.py:388-390)")
                wide[b].copy_(row)
            lo, hi = int(wide.min()), int(wide.max())
            if lo < 0 or hi >= vocab_size:
                raise IndexError(f"token id out of range: ids must lie in [0, {vocab_size}), got [{lo}, {hi}]")
            out[i].copy_(wide)
        y = torch.as_tensor([int(smp[5]) for smp in samples], dtype=torch.int64)
        return PackedBatch(out, y)

    return collate


class DeviceFeeder:
    """Iterate `loader` and yield device-resident PackedBatches, copies double-buffered on a side stream.

    Per pipeline slot: one pinned [5, B, T] staging tensor + label row, one device twin, a `ready` event (H2D done) and a `free`
    event (the consumer's stream is done with the slot).  `__next__` hands out batch k after enqueueing the H2D of batch k+1, so
    that copy runs under step k; the consumer's stream waits on `ready`, the copy stream on `free` (both on the GPU), and the host
    only ever waits for an H2D copy that was issued `depth` batches earlier -- it never waits for compute.
    A batch handed out stays valid until the NEXT batch has been requested (`depth` = 2) -- the Lightning loop's usage pattern.
    Ragged last batches re-use the buffers' leading rows.  Accepts PackedBatches (narrow_collate) or the reference's
    default-collated 6-tuples of int64 tensors (packed here, on the consumer thread)."""

    def __init__(self, loader: Iterable, device, vocab_size: int, depth: int = 2, check_range: bool = True):
        if depth < 2:
            raise ValueError("depth >= 2: one slot computing, one slot copying")
        self.loader, self.device, self.vocab_size, self.depth = loader, torch.device(device), int(vocab_size), int(depth)
        self.check_range = check_range
        if self.device.type != "cuda":
            raise RuntimeError("DeviceFeeder feeds a CUDA device (intrepppid_b200 has no CPU path)")
        self.dtype = token_dtype_for(vocab_size)
        self._slots: List[dict] = []
        self._stream: Optional[torch.cuda.Stream] = None
        self.h2d_bytes = 0  # bytes copied host -> device so far (bench.py reports them per step)

    def __len__(self):
        return len(self.loader)

    def _slot(self, k: int, B: int, T: int) -> dict:
        if len(self._slots) <= k:
            self._slots.append({})
        s = self._slots[k]
        if not s or s["cap"] < B or s["T"] != T:
            s.clear()
            s.update(cap=B, T=T,
                     host=torch.empty(5, B, T, dtype=self.dtype).pin_memory(), host_y=torch.empty(B, dtype=torch.int64).pin_memory(),
                     dev=torch.empty(5, B, T, dtype=self.dtype, device=self.device),
                     dev_y=torch.empty(B, dtype=torch.int64, device=self.device),
                     ready=torch.cuda.Event(), free=None, used=False)
        return s

    def _stage(self, k: int, batch) -> PackedBatch:
        """Pack `batch` into slot k's pinned buffer and enqueue its H2D copy on the copy stream."""
        if isinstance(batch, PackedBatch):
            tokens, y = batch.tokens, batch.y
            if tokens.dtype != self.dtype:
                raise TypeError(f"PackedBatch holds {tokens.dtype} ids, the feeder was built for {self.dtype} (vocab_size={self.vocab_size})")
        else:
            *seqs, y = batch
            tokens = None
        B, T = (tokens.shape[1], tokens.shape[2]) if tokens is not None else tuple(seqs[0].shape)
        s = self._slot(k, B, T)
        # the HOST only waits for the previous H2D copy out of this pinned buffer (done long ago in steady state); the hazard on the
        # DEVICE buffer -- the step that still reads it -- is ordered on the GPU: the copy stream waits for the slot's `free` event
        if s["used"]:
            s["ready"].synchronize()
        host, host_y = s["host"][:, :B], s["host_y"][:B]
        if tokens is not None:
            host.copy_(tokens)
            host_y.copy_(y)
        else:
            pack_batch(seqs, y, self.vocab_size, out=host, check_range=self.check_range)
            host_y.copy_(torch.as_tensor(y).reshape(-1))
        dev, dev_y = s["dev"][:, :B], s["dev_y"][:B]
        with torch.cuda.stream(self._stream):
            if s["free"] is not None:
                self._stream.wait_event(s["free"])
            dev.copy_(host, non_blocking=True)
            dev_y.copy_(host_y, non_blocking=True)
            s["ready"].record(self._stream)
        s["used"] = True
        self.h2d_bytes += host.numel() * host.element_size() + host_y.numel() * 8
        return PackedBatch(dev, dev_y)

    def __iter__(self) -> Iterator[PackedBatch]:
        with torch.cuda.device(self.device):
            if self._stream is None:
                self._stream = torch.cuda.Stream(self.device)
            it = iter(self.loader)
            k = 0
            try:
                nxt = self._stage(0, next(it))
            except StopIteration:
                return
            while nxt is not None:
                cur, cur_slot = nxt, self._slots[k % self.depth]
                try:
                    nxt = self._stage((k + 1) % self.depth, next(it))  # H2D of batch k+1 under the compute of batch k
                except StopIteration:
                    nxt = None
                torch.cuda.current_stream(self.device).wait_event(cur_slot["ready"])
                yield cur
                # the consumer asked for the next batch: everything it enqueued on its stream so far covers batch k
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))
                cur_slot["free"] = ev
                k += 1


def feeding_dataloader(dataset, batch_size: int, vocab_size: int, device, num_workers: int = 4, shuffle: bool = True, **kw):
    """`DataLoader(dataset, batch_size=..., num_workers=..., shuffle=...)` as in data/ppi_oma.py:611-620, with the narrow collate in
    the workers and the double-buffered device feeder around it.  Iterating it yields device PackedBatches."""
    from torch.utils.data import DataLoader

    loader = DataLoader(dataset, batch_size=batch_size, num_workers=num_workers, shuffle=shuffle,
                        collate_fn=narrow_collate(vocab_size), **kw)
    return DeviceFeeder(loader, device, vocab_size)
