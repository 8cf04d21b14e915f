"""MLPHead on the fused head kernel -- drop-in for intrepppid/classifier/head/mlp.py:22-68 (same constructor, forward
signature, state_dict keys `classify.fc{1,2}.module.{bias,weight_raw}` and init order)."""
from __future__ import annotations

from collections import OrderedDict

import torch
from torch import nn

from ... import ops
from ...utils import WeightDrop


class MLPHead(nn.Module):
    def __init__(self, embedding_size, do_rate):
        super().__init__()
        self.embedding_size = embedding_size
        self.do_rate = do_rate
        # The Sequential only fixes parameter names / init order; it is never called (the fused kernel evaluates
        # (z1+z2)/2 -> Mish -> WD-Linear -> Mish -> Dropout -> Mish -> Dropout -> WD-Linear in one launch).
        self.classify = nn.Sequential(OrderedDict([
            ("nl0", nn.Mish()),
            ("fc1", WeightDrop(nn.Linear(embedding_size, embedding_size // 2), ["weight"], dropout=do_rate, variational=False)),
            ("nl1", nn.Mish()),
            ("do1", nn.Dropout(p=do_rate)),
            ("nl2", nn.Mish()),
            ("do2", nn.Dropout(p=do_rate)),
            ("fc2", WeightDrop(nn.Linear(embedding_size // 2, 1), ["weight"], dropout=do_rate, variational=False)),
        ]))

    def tensors(self):
        fc1, fc2 = self.classify.fc1.module, self.classify.fc2.module
        return fc1.weight_raw, fc1.bias, fc2.weight_raw, fc2.bias

    def draw_masks(self, batch: int, generator=None):
        """(fc1_w, do1, do2, fc2_w) scaled masks in the reference's draw order, or Nones in eval mode."""
        if not self.training or self.do_rate <= 0:
            return (None, None, None, None)
        p, dev, hh = float(self.do_rate), self.classify.fc1.module.bias.device, self.embedding_size // 2

        def drop(shape):
            return torch.empty(shape, dtype=torch.float32, device=dev).bernoulli_(1.0 - p, generator=generator) / (1.0 - p)

        m1 = self.classify.fc1.sample_mask("weight", generator=generator)
        d1, d2 = drop((batch, hh)), drop((batch, hh))
        m2 = self.classify.fc2.sample_mask("weight", generator=generator)
        return (m1, d1, d2, m2)

    def forward(self, x1, x2, masks=None):
        """logits [B,1].  Runs the fused loss/head kernel with the triplet slots unused (zero upstream gradient)."""
        if masks is None:
            masks = self.draw_masks(x1.shape[0])
        z = torch.stack((x1, x1, x1, x1, x2), dim=0)
        y = torch.zeros(x1.shape[0], dtype=torch.int64, device=x1.device)
        _losses, y_hat = ops.loss_head(1.0, z, y, *self.tensors(), masks=masks)
        return y_hat.unsqueeze(1)
