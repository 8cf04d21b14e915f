"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference hot-path files from /root/reference.

This module exists to (a) validate `oracle/restatement.py` against the real reference and
(b) generate the golden vectors under `tests/golden/` (see `oracle/make_golden.py`).  It only
works in the build container (where /root/reference is mounted); it is never imported by the
product package, by `-m gpu` tests or by `smoke()`.  `bench.py`'s CPU legs (`--impl reference`, `cpu_baseline`) use it when
the reference files are reachable: mounted at /root/reference (build container) or staged byte-for-byte under oracle/_ref by
oracle/stage_reference.py (the GPU box).

`import intrepppid` fails here (pytorch_lightning / torchmetrics / ranger21 / tables are not
installed), so the five hot-path files are loaded by path under their real dotted names after
seeding `sys.modules` with inert stand-ins for the missing third-party packages:

    intrepppid/utils/weightdrop.py, intrepppid/utils/embedding_do.py,
    intrepppid/encoders/awd_lstm.py, intrepppid/classifier/head/mlp.py,
    intrepppid/e2e/e2e_triplet.py
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types

import torch
from torch import nn

def _find_root() -> str:
    """The mounted reference (build container), else the byte-for-byte staged copies under oracle/_ref (oracle/stage_reference.py:
    git-ignored, travels to the GPU box with the snapshot, sha256-verified before use)."""
    env = os.environ.get("IB200_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile("/root/reference/intrepppid/encoders/awd_lstm.py"):
        return "/root/reference"
    from . import stage_reference

    if stage_reference.verify():
        return stage_reference.DEST
    return "/root/reference"


REF_ROOT = _find_root()
REF_PKG = os.path.join(REF_ROOT, "intrepppid")
STAGED = REF_ROOT != "/root/reference" and not os.environ.get("IB200_REFERENCE_ROOT")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_PKG, "encoders", "awd_lstm.py"))


class _LightningModule(nn.Module):
    def log(self, *a, **k):
        pass


class _Anything:
    def __init__(self, *a, **k):
        pass


class _Metric(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, y_hat, y):
        return torch.zeros(())


_loaded = None


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__path__ = []
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def load():
    """Returns a namespace with the reference's AWDLSTMEncoder, MLPHead, TripletE2ENet, WeightDrop, embedding_dropout."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not mounted at {REF_ROOT}")
    sys.dont_write_bytecode = True
    saved = {k: v for k, v in sys.modules.items() if k == "intrepppid" or k.startswith("intrepppid.")}
    _stub("pytorch_lightning", LightningModule=_LightningModule, LightningDataModule=object, Trainer=_Anything)
    _stub("pytorch_lightning.callbacks", ModelCheckpoint=_Anything, LearningRateMonitor=_Anything,
          StochasticWeightAveraging=_Anything)
    _stub("pytorch_lightning.loggers", TensorBoardLogger=_Anything, Logger=object)
    _stub("pytorch_lightning.utilities")
    _stub("pytorch_lightning.utilities.seed", seed_everything=torch.manual_seed)
    _stub("torchmetrics", AUROC=_Metric, AveragePrecision=_Metric, MatthewsCorrCoef=_Metric, Precision=_Metric,
          Recall=_Metric)
    _stub("ranger21", Ranger21=_Anything)
    for p in ("intrepppid", "intrepppid.utils", "intrepppid.data", "intrepppid.classifier", "intrepppid.encoders",
              "intrepppid.e2e"):
        _stub(p)
    _stub("intrepppid.data.ppi_oma", IntrepppidDataModule=_Anything)
    with contextlib.redirect_stdout(None):
        wd = _load("intrepppid.utils.weightdrop", f"{REF_PKG}/utils/weightdrop.py")
        ed = _load("intrepppid.utils.embedding_do", f"{REF_PKG}/utils/embedding_do.py")
        u = sys.modules["intrepppid.utils"]
        u.WeightDrop, u.embedding_dropout, u.DictLogger = wd.WeightDrop, ed.embedding_dropout, _Anything
        enc = _load("intrepppid.encoders.awd_lstm", f"{REF_PKG}/encoders/awd_lstm.py")
        mlp = _load("intrepppid.classifier.head.mlp", f"{REF_PKG}/classifier/head/mlp.py")
        _stub("intrepppid.classifier.head", MLPHead=mlp.MLPHead)
        e2e = _load("intrepppid.e2e.e2e_triplet", f"{REF_PKG}/e2e/e2e_triplet.py")
    ns = types.SimpleNamespace(AWDLSTMEncoder=enc.AWDLSTMEncoder, AWDLSTM=enc.AWDLSTM, MLPHead=mlp.MLPHead,
                               TripletE2ENet=e2e.TripletE2ENet, WeightDrop=wd.WeightDrop,
                               embedding_dropout=ed.embedding_dropout)
    # Do not leave the stub "intrepppid" package in sys.modules: the product package must never resolve to it.
    for k in [k for k in sys.modules if k == "intrepppid" or k.startswith("intrepppid.")]:
        del sys.modules[k]
    sys.modules.update(saved)
    _loaded = ns
    return ns


def build_reference_net(vocab=250, E=64, L=2, bi_reduce="last", use_projection=False, beta=2.0, emb_droprate=0.3,
                        rnn_droprate=0.3, do_rate=0.3, variational=False):
    """Mirrors intrepppid/__init__.py:71-88 / e2e_triplet.py:345-373 with the reference's own classes."""
    ref = load()
    with contextlib.redirect_stdout(None):  # WeightDrop._setup prints
        emb = nn.Embedding(vocab, E, padding_idx=0)
        encoder = ref.AWDLSTMEncoder(emb, E, emb_droprate, L, rnn_droprate, variational, bi_reduce)
        head = ref.MLPHead(E, do_rate)
        net = ref.TripletE2ENet(E, encoder, head, emb_droprate, 1, 1, beta, use_projection, "adamw", 1e-2)
    return net


@contextlib.contextmanager
def injected_masks(row_masks, dropout_masks):
    """Replace the RNG draws of one reference `step()` with supplied masks (SURVEY Q6 order).

    row_masks:     iterable of 0/1 keep masks [V,1] consumed by `Tensor.bernoulli_` (embedding_do.py:26-28;
                   the reference divides by (1-p) itself).
    dropout_masks: iterable of ALREADY SCALED masks (keep/(1-p)) consumed by every `F.dropout` call in
                   training mode (weightdrop.py:100-102, nn.Dropout in mlp.py:49-51).
    """
    import torch.nn.functional as F

    rows, drops = iter(row_masks), iter(dropout_masks)
    orig_dropout, orig_bern = F.dropout, torch.Tensor.bernoulli_

    def fake_dropout(inp, p=0.5, training=True, inplace=False):
        if not training:
            return inp
        m = next(drops)
        return inp * m.to(inp.dtype)

    def fake_bernoulli_(self, p=0.5, generator=None):
        return self.copy_(next(rows).to(self.dtype))

    F.dropout, torch.Tensor.bernoulli_ = fake_dropout, fake_bernoulli_
    try:
        yield
    finally:
        F.dropout, torch.Tensor.bernoulli_ = orig_dropout, orig_bern
