"""AWD-LSTM encoder on the sm_100a kernels -- drop-in for intrepppid/encoders/awd_lstm.py (same constructors, same forward
signature, same state_dict keys, same initial weights under the same seed).

Reference behaviour reproduced (file:line in /root/reference/intrepppid/):
  * AWDLSTMEncoder.forward  encoders/awd_lstm.py:147-155  first truncation (count of non-zero ids), embedding dropout over
                                                          vocabulary rows, then AWDLSTM
  * AWDLSTM.forward         encoders/awd_lstm.py:51-74    second truncation on the embedded tensor (training-mode quirk Q2),
                                                          weight-dropped 2-layer bi-LSTM without packing, bi_reduce on h_n, fc
  * WeightDrop on ["weight_hh_l0"] only (layer 0, forward direction)  :43-45
  * Projection              encoders/awd_lstm.py:77-105   constructed, never called (its parameters stay in the checkpoint)
The arithmetic itself runs in csrc/ (K0 lengths, K1 table/GEMM, K2 recurrent forward, K3 BPTT, K4 pool+fc).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
from torch import nn

from .. import ops
from ..utils import WeightDrop
from ..utils.embedding_do import embedding_row_scale


class LSTMWeights(nn.Module):
    """Parameter container with nn.LSTM's names, shapes, registration order and init (U(-1/sqrt(H), 1/sqrt(H)) drawn in the
    same order as nn.LSTM.reset_parameters), so seeds and checkpoints are interchangeable with the reference's
    `nn.LSTM(E, E, L, bidirectional=True, batch_first=True)` (awd_lstm.py:35-41).  It has no forward: the recurrence is
    executed by the CUDA kernels, never by cuDNN/ATen."""

    def __init__(self, input_size: int, hidden_size: int, num_layers: int):
        super().__init__()
        self.input_size, self.hidden_size, self.num_layers, self.bidirectional = input_size, hidden_size, num_layers, True
        for l in range(num_layers):
            in_l = input_size if l == 0 else 2 * hidden_size
            for sfx in ("", "_reverse"):
                self.register_parameter(f"weight_ih_l{l}{sfx}", nn.Parameter(torch.empty(4 * hidden_size, in_l)))
                self.register_parameter(f"weight_hh_l{l}{sfx}", nn.Parameter(torch.empty(4 * hidden_size, hidden_size)))
                self.register_parameter(f"bias_ih_l{l}{sfx}", nn.Parameter(torch.empty(4 * hidden_size)))
                self.register_parameter(f"bias_hh_l{l}{sfx}", nn.Parameter(torch.empty(4 * hidden_size)))
        stdv = 1.0 / math.sqrt(hidden_size)
        for w in self.parameters():
            nn.init.uniform_(w, -stdv, stdv)

    def ordered(self) -> List[torch.Tensor]:
        """The 8L tensors in C-ABI order; a weight renamed by WeightDrop is found under `<name>_raw`."""
        out = []
        for n in ops.lstm_param_order(self.num_layers):
            out.append(getattr(self, n) if n in self._parameters else getattr(self, n + "_raw"))
        return out

    def forward(self, *a, **k):
        raise RuntimeError("LSTMWeights only holds parameters; the LSTM runs inside AWDLSTM's CUDA kernels")


class AWDLSTM(nn.Module):
    def __init__(self, embedding_size, rnn_num_layers, lstm_dropout_rate, variational_dropout, bi_reduce):
        super().__init__()
        self.bi_reduce = bi_reduce
        self.rnn = LSTMWeights(embedding_size, embedding_size, rnn_num_layers)
        self.rnn_dp = WeightDrop(self.rnn, ["weight_hh_l0"], lstm_dropout_rate, variational_dropout)
        self.fc = nn.Linear(embedding_size, embedding_size)
        self.nl = nn.Mish()  # constructed but not applied, as in the reference (awd_lstm.py:48,72)
        self.embedding_size = embedding_size

    def forward(self, x):
        raise RuntimeError("AWDLSTM consumes token ids through AWDLSTMEncoder (embedding gather, both truncations and the LSTM "
                           "are fused); it cannot be fed an embedded float tensor. There is no eager fallback.")


class Projection(nn.Module):
    """Unused MLP kept for checkpoint-key compatibility (encoder.projection.model.{0,2,4}.*; awd_lstm.py:77-105,140-142)."""

    def __init__(self, in_dim, out_dim, num_layers):
        super().__init__()
        diff_dim = (out_dim - in_dim) // num_layers
        layers, dim = [], in_dim
        for _ in range(num_layers - 1):
            layers.append(nn.Linear(dim, dim + diff_dim))
            layers.append(nn.ReLU())
            dim += diff_dim
        layers.append(nn.Linear(dim, out_dim))
        self.model = nn.Sequential(*layers)

    def forward(self, x):
        return self.model(x)


class AWDLSTMEncoder(nn.Module):
    def __init__(self, embedder: nn.Module, embedding_size: int, embedding_droprate: float, rnn_num_layers: int,
                 rnn_dropout_rate: float, variational_dropout: bool, bi_reduce: str):
        super().__init__()
        # the kernels implement exactly what the reference passes to F.embedding (utils/embedding_do.py:30-43 with the embedder of
        # e2e_triplet.py:345): padding_idx 0 (row 0 never receives a gradient), no max_norm renormalisation, no frequency scaling
        if isinstance(embedder, nn.Embedding):
            pad = embedder.padding_idx
            if pad is None:
                pad = -1  # embedding_do.py:31-32 turns None into -1 (= no padding row); the kernels always skip row 0
            if pad != 0 or embedder.max_norm is not None or embedder.scale_grad_by_freq or embedder.sparse:
                raise ValueError("AWDLSTMEncoder on the sm_100a kernels needs nn.Embedding(V, E, padding_idx=0) without max_norm, "
                                 "scale_grad_by_freq or sparse gradients (the reference's embedder, e2e_triplet.py:345)")
        self.embedder = embedder
        self.embedding_droprate = embedding_droprate
        self.encoder = AWDLSTM(embedding_size, rnn_num_layers, rnn_dropout_rate, variational_dropout, bi_reduce)
        self.projection = Projection(self.encoder.embedding_size, self.encoder.embedding_size * 2, 3)
        self.precision = "fp32"       # "fp32" (bf16x2-split tensor-core, 1e-4 class) or "bf16" (2e-2 class)
        # True: errors of the reference (id outside [0,V) -> IndexError, all-pad batch -> RuntimeError) are detected on the device
        # and raised at the next call / ops.check_pending(), without a host sync; "sync": raised before returning (one host sync,
        # the reference pays two per call); False: not checked
        self.check_lengths = True
        self.last_lengths: Optional[torch.Tensor] = None  # int32 [2,G] (T1, T_eff) of the most recent call, on device

    # -- masks ---------------------------------------------------------------------------------------------------------------
    def draw_masks(self, groups: int, generator=None):
        """(emb_row_scale [G,V] | None, whh_l0_mask [G,4H,H] | None), drawn the way the reference draws them per encoder call
        (embedding_do.py:26-29 then weightdrop.py:92-102)."""
        ers = embedding_row_scale(self.training, self.embedder, self.embedding_droprate, groups, generator)
        whm = self.encoder.rnn_dp.sample_mask("weight_hh_l0", groups, generator)
        return ers, whm

    # -- forward -------------------------------------------------------------------------------------------------------------
    def forward_groups(self, tokens: torch.Tensor | Sequence[torch.Tensor], emb_row_scale=None, whh_mask=None,
                       draw: bool = True) -> torch.Tensor:
        """Encode G independent encoder calls at once.  tokens: [G,B,T] (or a list of G [B,T] tensors).  Each group has its
        own masks and truncation lengths, exactly as if `forward` had been called G times.  Returns z [G,B,E]."""
        if not torch.is_tensor(tokens):
            tokens = torch.stack(list(tokens), dim=0)
        if tokens.dim() != 3:
            raise ValueError("tokens must be [G,B,T]")
        G, B, _ = tokens.shape
        if draw and emb_row_scale is None and whh_mask is None:
            emb_row_scale, whh_mask = self.draw_masks(G)
        econf = ops.EncoderConfig(self.encoder.rnn.num_layers, self.encoder.bi_reduce, self.precision)
        holder: list = []
        hn = ops.encode_hidden(econf, tokens, self.embedder.weight, self.encoder.rnn.ordered(), emb_row_scale, whh_mask,
                               self.check_lengths, holder)
        self.last_lengths = holder[0] if holder else None
        z = ops.pool_fc(self.encoder.bi_reduce, hn, self.encoder.fc.weight, self.encoder.fc.bias)
        return z.view(G, B, -1)

    def forward(self, x, emb_row_scale=None, whh_mask=None):
        """x: Long[B, trunc_len] -> Float[B, E] (encoders/awd_lstm.py:147-155)."""
        ers = None if emb_row_scale is None else emb_row_scale.reshape(1, -1)
        whm = None if whh_mask is None else whh_mask.reshape(1, *whh_mask.shape[-2:])
        return self.forward_groups(x.unsqueeze(0), ers, whm)[0]

    def embedding_dropout(self, embed, words, p=0.2):
        raise RuntimeError("embedding dropout is fused into the encoder kernels; use AWDLSTMEncoder.forward")
