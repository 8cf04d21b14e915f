"""TripletE2ENet on the sm_100a kernels -- drop-in for intrepppid/e2e/e2e_triplet.py:43-255 (same constructor, forward /
step / *_step / configure_optimizers, same state_dict keys).

`step()` fuses the reference's five encoder calls (anchor, positive, negative, then p1, p2 -- e2e_triplet.py:116-129) into one
G=5 launch set with per-group masks and truncation lengths, then runs triplet + head + BCE + beta mix in one kernel.
Lightning / torchmetrics are optional: when they are importable the class is a LightningModule and logs exactly what the
reference logs; otherwise it is a plain nn.Module and `log` is a no-op (the arithmetic is identical).
"""
from __future__ import annotations

import os
from typing import Optional

import torch
from torch import nn
from torch.optim.lr_scheduler import CosineAnnealingWarmRestarts, OneCycleLR

from .. import ops
from ..optim import FusedAdamW as AdamW  # torch.optim.AdamW's arguments / state layout on the multi-tensor CUDA kernel
from ..optim import FusedRanger21

try:  # optional orchestration dependencies (absent in the build image)
    import pytorch_lightning as pl

    _Base = pl.LightningModule
except Exception:  # pragma: no cover
    pl = None

    class _Base(nn.Module):
        def log(self, *a, **k):
            pass

try:
    import torchmetrics
except Exception:  # pragma: no cover
    torchmetrics = None


class StepMasks:
    """Explicit masks for one training step (SURVEY Q6 draw order); any field may be None (= no drop).
    emb_row_scale [5,V] (keep/(1-p)), whh_mask [5,4H,H], head = (fc1_w [H/2,H], do1 [B,H/2], do2 [B,H/2], fc2_w [1,H/2])."""

    def __init__(self, emb_row_scale=None, whh_mask=None, head=(None, None, None, None)):
        self.emb_row_scale, self.whh_mask, self.head = emb_row_scale, whh_mask, tuple(head)


class TripletE2ENet(_Base):
    def __init__(self, embedding_size: int, encoder: nn.Module, head: nn.Module, embedding_droprate: float, num_epochs: int,
                 steps_per_epoch: int, beta_classifier: float, use_projection: bool, optimizer_type: str, lr: float):
        super().__init__()
        self.encoder = encoder
        self.embedding_droprate = embedding_droprate
        self.classifier_criterion = nn.BCEWithLogitsLoss()   # kept for attribute compatibility; evaluated in the fused kernel
        self.num_epochs = num_epochs
        self.steps_per_epoch = steps_per_epoch
        self.triplet_criterion = nn.TripletMarginLoss(margin=1.0, p=2)
        if use_projection:
            self.triplet_projection = nn.Sequential(nn.Mish(), nn.Linear(embedding_size, embedding_size))
        if torchmetrics is not None:
            self.auroc = torchmetrics.AUROC(task="binary")
            self.average_precision = torchmetrics.AveragePrecision(task="binary")
            self.mcc = torchmetrics.MatthewsCorrCoef(task="binary", threshold=0.5)
            self.precision_metric = torchmetrics.Precision(task="binary")
            self.recall = torchmetrics.Recall(task="binary")
        self.do_rate = 0.3
        self.head = head
        self.beta_classifier = beta_classifier
        self.optimizer_type = optimizer_type
        self.lr = lr
        self.use_projection = use_projection
        self.last_step = None  # dict of detached tensors from the most recent step()
        self.fused_masks = True   # draw the step's dropout masks in one launch (ib200_draw_masks) instead of torch RNG kernels
        self._mask_offset = 0     # Philox counters consumed so far by this module
        self.compute_metrics = True  # log the reference's five per-step metrics (ib200_batch_metrics)

    # -- inference API (e2e_triplet.py:105-111): no projection here (quirk Q11); the caller applies sigmoid ---------------------
    def forward(self, x1, x2):
        z = self.encoder.forward_groups(torch.stack((x1, x2), dim=0))
        return self.head(z[0], z[1])

    @torch.no_grad()
    def embed(self, x, batch_size: int = 512):
        """Eval-mode embeddings [M,E] of M sequences, for encode-once / score-all-pairs inference."""
        out = []
        for i in range(0, x.shape[0], batch_size):
            out.append(self.encoder(x[i:i + batch_size]))
        return torch.cat(out, dim=0)

    @torch.no_grad()
    def score_pairs(self, z, idx_a=None, idx_b=None):
        """sigmoid(head(z[i], z[j])) for explicit pairs or the whole upper triangle (eval mode)."""
        return ops.pair_score(z, *self.head.tensors(), idx_a, idx_b)

    @torch.no_grad()
    def score_pairs_range(self, z, p_begin: int, p_count: int):
        """The slice [p_begin, p_begin + p_count) of `score_pairs(z)` (flat row-major upper triangle) -- one rank's block in
        multi-GPU inference (intrepppid_b200.parallel.sharded_proteome_scores)."""
        return ops.pair_score_range(z, *self.head.tensors(), p_begin, p_count)

    # -- training step (e2e_triplet.py:113-187) ---------------------------------------------------------------------------------
    def step(self, batch, stage, masks: Optional[StepMasks] = None):
        p1_seq, p2_seq, omid_anchor_seq, omid_positive_seq, omid_negative_seq, y = batch
        tokens = torch.stack((omid_anchor_seq, omid_positive_seq, omid_negative_seq, p1_seq, p2_seq), dim=0)
        if masks is None:
            ers, whm, head_masks = self._draw_step_masks(y.shape[0])
        else:
            ers, whm, head_masks = masks.emb_row_scale, masks.whh_mask, masks.head
        # draw order of the reference: 5 x (row mask, W_hh mask), then the 4 head masks
        z = self.encoder.forward_groups(tokens, ers, whm, draw=False)
        if head_masks is None:
            head_masks = self.head.draw_masks(y.shape[0])
        proj_w = proj_b = None
        if self.use_projection:
            proj_w, proj_b = self.triplet_projection[1].weight, self.triplet_projection[1].bias
        losses, y_hat = ops.loss_head(self.beta_classifier, z, y, *self.head.tensors(), proj_w, proj_b, masks=head_masks)
        loss, classifier_loss, triplet_loss = losses[0], losses[1].detach(), losses[2].detach()
        self.last_step = {"loss": loss.detach(), "classifier_loss": classifier_loss, "triplet_loss": triplet_loss,
                          "y_hat": y_hat.detach(), "z": z.detach(), "lengths": self.encoder.last_lengths}

        self.log(f"{stage}_classifier_loss", classifier_loss, on_epoch=True, on_step=False, prog_bar=True)
        self.log(f"{stage}_triplet_loss", triplet_loss, on_epoch=True, on_step=False, prog_bar=True)
        self.log(f"{stage}_loss", loss, on_epoch=True, on_step=False, prog_bar=True)
        self.log(f"{stage}_classifier_loss_step", classifier_loss, on_epoch=False, on_step=True, prog_bar=False)
        self.log(f"{stage}_triplet_loss_step", triplet_loss, on_epoch=False, on_step=True, prog_bar=False)
        self.log(f"{stage}_loss_step", loss, on_epoch=False, on_step=True, prog_bar=False)
        # the five torchmetrics forwards of the reference (batch values, e2e_triplet.py:171-184) in one launch, no host sync
        if self.compute_metrics and y.shape[0] <= 1024:
            m, conf = ops.batch_metrics(y_hat.detach(), y)
            self.last_step["metrics"], self.last_step["confusion"] = m, conf
            for k, name in enumerate(ops.METRIC_NAMES):
                self.log(f"{stage}_{name}", m[k], on_epoch=True, on_step=False)
        return loss

    def _draw_step_masks(self, batch_size: int):
        """The 14 dropout masks of a training step -> (emb_row_scale [5,V] | None, whh_mask [5,4H,H] | None, head masks
        | None).  Production mode draws them in one launch (`ib200_draw_masks`, seeded from torch's CUDA seed plus a per-module
        counter); anything that kernel does not cover (variational row masks, p >= 1, `fused_masks = False`) goes through the
        torch-RNG draws of the encoder / head modules, which return None for the head so that step() draws them later."""
        enc, rnn_dp = self.encoder, self.encoder.encoder.rnn_dp
        p_emb, p_rnn, p_do = float(enc.embedding_droprate or 0.0), float(rnn_dp.dropout or 0.0), float(self.head.do_rate or 0.0)
        dev = enc.embedder.weight.device
        if (not self.training or not self.fused_masks or rnn_dp.variational or dev.type != "cuda"
                or any(not (0.0 <= p < 1.0) for p in (p_emb, p_rnn, p_do))):
            ers, whm = enc.draw_masks(5)
            return ers, whm, None
        V, H = enc.embedder.weight.shape
        HH = H // 2
        want = []
        if p_emb > 0:
            want.append(("ers", (5, V), 1.0 - p_emb))
        if p_rnn > 0:
            want.append(("whm", (5, 4 * H, H), 1.0 - p_rnn))
        if p_do > 0:
            want += [("fc1", (HH, H), 1.0 - p_do), ("do1", (batch_size, HH), 1.0 - p_do), ("do2", (batch_size, HH), 1.0 - p_do),
                     ("fc2", (1, HH), 1.0 - p_do)]
        got = {}
        if self._mask_offset == 0 and torch.distributed.is_available() and torch.distributed.is_initialized():
            self._mask_offset = torch.distributed.get_rank() << 44  # data-parallel ranks share the seed: disjoint counter ranges
        if want:
            tensors, used = ops.draw_masks([(shape, keep, 0) for _, shape, keep in want], dev, torch.cuda.initial_seed(),
                                           self._mask_offset)
            self._mask_offset += used
            got = {name: t for (name, _, _), t in zip(want, tensors)}
        head = (got.get("fc1"), got.get("do1"), got.get("do2"), got.get("fc2"))
        return got.get("ers"), got.get("whm"), head

    # -- checkpoint hooks (Lightning calls them; the state_dict keys stay exactly the reference's 45) ---------------------------------
    def on_save_checkpoint(self, checkpoint):
        checkpoint["ib200_mask_offset"] = int(self._mask_offset)  # Philox counters consumed: a resumed run continues the mask stream

    def on_load_checkpoint(self, checkpoint):
        self._mask_offset = int(checkpoint.get("ib200_mask_offset", 0))

    def on_train_epoch_end(self):
        ops.check_pending(sync=True)  # surface any error the kernels recorded (out-of-range ids, all-pad batch) at the latest here

    def training_step(self, batch, batch_idx):
        return self.step(batch, "train")

    def validation_step(self, batch, batch_idx):
        return self.step(batch, "val")

    def test_step(self, batch, batch_idx):
        return self.step(batch, "test")

    # -- optimizers (e2e_triplet.py:198-255): AdamW variants -> ib200_adamw_step, Ranger21 variants -> ib200_ranger21_step ----------
    def configure_optimizers(self):
        if self.optimizer_type in ("ranger21", "ranger21_xx"):
            xx = self.optimizer_type == "ranger21_xx"
            kw = dict(use_warmup=xx, warmdown_active=xx, lr=self.lr, weight_decay=1e-2, num_batches_per_epoch=self.steps_per_epoch,
                      num_epochs=self.num_epochs, warmdown_start_pct=0.72)
            if os.environ.get("IB200_RANGER21", "fused") == "package":  # the third-party package itself (the reference's pin), when installed
                from ranger21 import Ranger21

                return Ranger21(self.parameters(), **kw)
            return FusedRanger21(self.parameters(), **kw)  # parity unpinned against the package: see optim.FusedRanger21
        if self.optimizer_type == "adamw":
            return AdamW(self.parameters(), lr=self.lr)
        if self.optimizer_type == "adamw_1cycle":
            optimizer = AdamW(self.parameters(), lr=self.lr)
            return [optimizer], [OneCycleLR(optimizer, self.lr, epochs=self.num_epochs, steps_per_epoch=self.steps_per_epoch)]
        if self.optimizer_type == "adamw_cosine":
            optimizer = AdamW(self.parameters(), lr=self.lr)
            return [optimizer], [CosineAnnealingWarmRestarts(optimizer, T_0=10, T_mult=2, eta_min=1e-6)]
        raise ValueError('Expected one of "ranger21", "adamw", "ranger21_xx", or "adamw_1cycle" as the optimizer type.')
