"""intrepppid_b200 -- B200 (sm_100a) implementation of INTREPPPID's sequence-encoder training/inference hot path.

Drop-in surface (same names, arguments and state_dict keys as the reference package `intrepppid`):
    intrepppid_network(...)                       intrepppid/__init__.py:23-88
    encoders.AWDLSTMEncoder / AWDLSTM             intrepppid/encoders/awd_lstm.py
    classifier.head.MLPHead                       intrepppid/classifier/head/mlp.py
    e2e.e2e_triplet.TripletE2ENet                 intrepppid/e2e/e2e_triplet.py
    utils.WeightDrop                              intrepppid/utils/weightdrop.py
The compute path is libib200.so (hand-written CUDA, C ABI in include/ib200.h); it must be built
(`python -m intrepppid_b200.build`) and needs a CUDA device.  There is no CPU / eager fallback.
"""
from torch import nn

from .classifier.head import MLPHead
from .e2e.e2e_triplet import StepMasks, TripletE2ENet
from .encoders.awd_lstm import AWDLSTMEncoder
from .optim import FusedAdamW, FusedRanger21

__all__ = ["intrepppid_network", "AWDLSTMEncoder", "MLPHead", "TripletE2ENet", "StepMasks", "FusedAdamW", "FusedRanger21"]


def intrepppid_network(steps_per_epoch: int, vocab_size: int = 250, embedding_size: int = 64, rnn_num_layers: int = 2,
                       rnn_dropout_rate: float = 0.3, variational_dropout: bool = False, bi_reduce: str = "last",
                       embedding_droprate: float = 0.3, num_epochs: int = 100, do_rate: float = 0.3, beta_classifier: int = 2,
                       lr: float = 1e-2, use_projection: bool = False, optimizer_type: str = "ranger21_xx",
                       precision: str = "fp32"):
    """The INTREPPPID network with the manuscript defaults (signature and defaults of intrepppid/__init__.py:23-38; the
    extra keyword `precision` selects the kernel arithmetic: "fp32" or "bf16")."""
    embedder = nn.Embedding(vocab_size, embedding_size, padding_idx=0)
    encoder = AWDLSTMEncoder(embedder, embedding_size, embedding_droprate, rnn_num_layers, rnn_dropout_rate,
                             variational_dropout, bi_reduce)
    encoder.precision = precision
    head = MLPHead(embedding_size, do_rate)
    return TripletE2ENet(embedding_size, encoder, head, embedding_droprate, num_epochs, steps_per_epoch, beta_classifier,
                         use_projection, optimizer_type, lr)
