"""Embedding dropout (reference: intrepppid/utils/embedding_do.py:20-44).

The reference multiplies a Bernoulli(1-p)/(1-p) mask over VOCABULARY ROWS into the table and gathers.  In this package the
row scale is an input of the fused kernels (it is folded into the layer-0 projection table, see csrc/small.cu), so the only
thing exposed here is the mask draw."""
from __future__ import annotations

import torch


def embedding_row_scale(training: bool, embed, p: float, groups: int = 1, generator=None):
    """[groups, V] scaled keep mask, or None when the reference applies no mask (eval, or p == 0; embedding_do.py:21-24)."""
    if not training or not p:
        return None
    V = embed.weight.size(0)
    keep = torch.empty(groups, V, dtype=torch.float32, device=embed.weight.device).bernoulli_(1.0 - p, generator=generator)
    return keep / (1.0 - p)


def embedding_dropout(training, embed, words, p=0.2):
    raise RuntimeError("intrepppid_b200 fuses embedding dropout + gather into the encoder kernels; call AWDLSTMEncoder instead "
                       "(use embedding_row_scale() to draw the mask). There is no eager fallback.")
