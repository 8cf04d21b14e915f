"""WeightDrop -- parameter-naming mirror of the reference's wrapper (reference: intrepppid/utils/weightdrop.py:22-111).

The reference re-samples a DropConnect mask and `setattr`s the masked tensor onto the wrapped module on every forward, then
runs the module through PyTorch.  Here the masked product is formed INSIDE the CUDA kernels that consume the weight
(lstm_fwd/lstm_bwd load `weight_hh_l0_raw * mask` straight into tensor-core fragments; the head kernel does the same for
`fc1/fc2.weight_raw`), so this class only has to reproduce the reference's parameter surgery -- `W` is re-registered as
`W_raw` -- which is what fixes the checkpoint key names (`...rnn_dp.module.weight_hh_l0_raw`, `...fc1.module.weight_raw`).
"""
from __future__ import annotations

import torch
from torch.nn import Parameter


class WeightDrop(torch.nn.Module):
    def __init__(self, module, weights, dropout=0, variational=True):
        super().__init__()
        self.module = module
        self.weights = weights
        self.dropout = dropout
        self.variational = variational
        self._setup()

    def _setup(self):
        # weightdrop.py:49-63: weight -> weight_raw (moves it to the end of the module's parameter order, like the reference)
        for name_w in self.weights:
            w = getattr(self.module, name_w)
            del self.module._parameters[name_w]
            self.module.register_parameter(name_w + "_raw", Parameter(w.data))

    def sample_mask(self, name_w: str, groups: int | None = None, generator=None):
        """Draw the scaled mask the reference would draw in `_setweights` (weightdrop.py:65-107) for `name_w`, on the weight's
        device.  Returns None when the reference applies no drop (DropConnect in eval mode, or p == 0).
        variational=True draws a row mask [rows,1] and, like the reference (`training=True` at :94), also in eval mode."""
        raw_w = getattr(self.module, name_w + "_raw")
        p = float(self.dropout)
        if self.variational:
            shape = (raw_w.size(0), 1)
        else:
            if not self.training:
                return None
            shape = tuple(raw_w.shape)
        if p <= 0.0:
            return None
        if groups is not None:
            shape = (groups, *shape)
        if p >= 1.0:
            m = torch.zeros(shape, dtype=torch.float32, device=raw_w.device)
        else:
            m = torch.empty(shape, dtype=torch.float32, device=raw_w.device).bernoulli_(1.0 - p, generator=generator) / (1.0 - p)
        if self.variational:
            m = m.expand(*m.shape[:-1], raw_w.size(1)).contiguous()
        return m

    def forward(self, *args):
        raise RuntimeError(
            "intrepppid_b200.WeightDrop does not execute the wrapped module through PyTorch: the weight-dropped LSTM and "
            "Linear layers run inside the fused CUDA kernels (AWDLSTMEncoder / MLPHead / TripletE2ENet). There is no eager fallback.")
