"""CPU-side checks for the SURVEY 8f rows built around the hot path: the AdamW restatement against torch.optim.AdamW (the
reference's optimizer, e2e_triplet.py:231-255), FusedAdamW's host-side contract, and the batch-of-one bucketing of the
inference path (cli/infer.py:196-225).  No CUDA compute here."""
import copy

import pytest
import torch

from oracle import restatement as R


def _rand_problem(seed, n=7, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    shapes = [(1,), (7,), (33, 5), (256, 64), (2049,), (64,), (3, 3, 3)][:n]
    params = [torch.randn(s, generator=g, dtype=dtype) for s in shapes]
    grads = [[torch.randn(s, generator=g, dtype=dtype) * 0.1 for s in shapes] for _ in range(6)]
    return params, grads


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-7), (torch.float64, 1e-14)])
def test_adamw_restatement_matches_torch_adamw(dtype, tol):
    params, grads = _rand_problem(3, dtype=dtype)
    ref = [torch.nn.Parameter(p.clone()) for p in params]
    opt = torch.optim.AdamW(ref, lr=1e-2, foreach=False)  # the reference's call: AdamW(self.parameters(), lr=self.lr)
    mine = [p.clone() for p in params]
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    for t, gs in enumerate(grads, start=1):
        for p, g in zip(ref, gs):
            p.grad = g.clone()
        opt.step()
        R.adamw_step(mine, gs, m, v, step=t, lr=1e-2)
        for a, b in zip(mine, ref):
            assert float((a - b.detach()).abs().max()) <= tol * max(1.0, float(b.abs().max()))
    st = opt.state[ref[3]]
    assert torch.allclose(st["exp_avg"], m[3], rtol=1e-6, atol=1e-9) and torch.allclose(st["exp_avg_sq"], v[3], rtol=1e-6, atol=1e-12)


def test_fused_adamw_host_contract():
    from intrepppid_b200 import FusedAdamW
    from intrepppid_b200._lib import IB200Error

    params = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5))]
    opt = FusedAdamW(params, lr=1e-2)
    tref = torch.optim.AdamW([torch.nn.Parameter(torch.randn(1))], lr=1e-2)
    for k in ("lr", "betas", "eps", "weight_decay", "amsgrad", "maximize"):  # same defaults as the reference's optimizer
        assert opt.param_groups[0][k] == tref.param_groups[0][k], k
    with pytest.raises(ValueError):
        FusedAdamW(params, lr=-1.0)
    with pytest.raises(ValueError):
        FusedAdamW(params, betas=(1.0, 0.999))
    with pytest.raises(ValueError):
        FusedAdamW(params, amsgrad=True)
    opt.step()  # no gradients yet: nothing to do, no state created (torch skips p.grad is None)
    assert len(opt.state) == 0
    params[0].grad = torch.zeros(4, 3)
    with pytest.raises(IB200Error):  # CPU parameters: there is no fallback
        opt.step()
    # schedulers used by the reference drive it like any torch optimizer (e2e_triplet.py:239-253)
    sched = torch.optim.lr_scheduler.OneCycleLR(FusedAdamW(params, lr=1e-2), 1e-2, epochs=2, steps_per_epoch=3)
    assert sched.get_last_lr()[0] < 1e-2


def test_configure_optimizers_uses_the_fused_kernel_for_adamw_types():
    import intrepppid_b200 as ib

    for kind in ("adamw", "adamw_1cycle", "adamw_cosine"):
        net = ib.intrepppid_network(3, optimizer_type=kind, num_epochs=2)
        got = net.configure_optimizers()
        opt = got if isinstance(got, torch.optim.Optimizer) else got[0][0]
        assert isinstance(opt, ib.FusedAdamW) and opt.param_groups[0]["lr"] in (net.lr, pytest.approx(net.lr / 25))
    with pytest.raises(ValueError):
        ib.intrepppid_network(3, optimizer_type="sgd").configure_optimizers()


# ---- inference bucketing --------------------------------------------------------------------------------------------------------
def _ragged_tokens(M, T, V, seed):
    g = torch.Generator().manual_seed(seed)
    tok = torch.randint(1, V, (M, T), generator=g)
    lens = torch.randint(1, T + 1, (M,), generator=g)
    lens[:4] = torch.tensor([T, T, 1, 5])
    for m in range(M):
        tok[m, lens[m]:] = 0
    tok[6, 2] = 0  # interior <unk>: T1 is a COUNT, so the slice loses the last real token
    return tok


def test_batch1_lengths_have_no_cpu_path():
    from intrepppid_b200 import _lib, build
    from intrepppid_b200.infer import batch1_lengths

    build.build()  # (no-op when libib200.so is current)
    with pytest.raises(_lib.IB200Error):  # (the per-sequence parity against the oracle runs on the GPU: tests/test_gpu_next_rows.py)
        batch1_lengths(_ragged_tokens(8, 10, 30, 2), torch.randn(30, 32))
    lib = _lib.lib()
    assert lib.ib200_sequence_lengths(0, 10, 30, 32, None, 0, None, None, None, None, None) == 0          # nothing to do
    assert lib.ib200_sequence_lengths(4, 10, 1, 32, None, 0, None, None, None, None, None) == -2          # V < 2
    assert lib.ib200_sequence_lengths(4, 10, 40000, 32, None, 0, None, None, None, None, None) == -2      # V past the histogram
    assert lib.ib200_sequence_lengths(4, 10, 30, 32, None, 9, None, None, None, None, None) == -2         # unknown id type
    assert lib.ib200_sequence_lengths(4, 10, 30, 32, None, 0, None, None, None, None, None) == -1         # null pointers


def test_plan_buckets_is_an_exact_partition_into_homogeneous_groups():
    from intrepppid_b200.infer import plan_buckets

    g = torch.Generator().manual_seed(5)
    t1 = torch.randint(1, 12, (300,), generator=g).tolist()
    keys = [(a, a - (i % 3 == 0 and a > 1)) for i, a in enumerate(t1)]
    plan = plan_buckets(keys, max_groups=16)
    seen = []
    for b, groups in plan:
        assert b in (8, 4, 2, 1) and 1 <= len(groups) <= 16
        for grp in groups:
            assert len(grp) == b and len({keys[m] for m in grp}) == 1  # one (T1, T_eff) per group = batch-of-one semantics
            seen += grp
    assert sorted(seen) == list(range(300))
    # at most 3 groups smaller than 8 per distinct key (binary decomposition of the remainder)
    small = sum(len(groups) for b, groups in plan if b < 8)
    assert small <= 3 * len(set(keys))
    with pytest.raises(ValueError):
        plan_buckets(keys, group_sizes=(4, 8, 1))


def test_infer_pairs_skips_unknown_ids_without_touching_the_gpu():
    from intrepppid_b200.infer import infer_pairs

    missing = []
    out = infer_pairs(None, {"A": torch.ones(4, dtype=torch.long)}, [("i0", "A", "B"), ("i1", "C", "A")],
                      on_missing=lambda *r: missing.append(r))
    assert out == [] and missing == [("i0", "A", "B"), ("i1", "C", "A")]


# ---- per-step metrics: the restatement of torchmetrics' binary metrics against scikit-learn -------------------------------------
def _metric_case(trial, g):
    B = int(torch.randint(2, 300, (1,), generator=g))
    y = torch.randint(0, 2, (B,), generator=g)
    if y.sum() == 0 or y.sum() == B:
        y[0] = 1 - y[0]
    x = torch.randn(B, generator=g) * 2
    if trial % 3 == 0:
        x = torch.round(x * 2) / 2            # tied scores: one curve point per distinct score
    if trial % 7 == 0:
        x = torch.rand(B, generator=g)        # all inside [0,1]: torchmetrics does NOT apply the sigmoid
    return x, y


def test_metrics_restatement_matches_scikit_learn():
    from sklearn import metrics as M

    g = torch.Generator().manual_seed(0)
    for trial in range(60):
        x, y = _metric_case(trial, g)
        m = R.batch_metrics(x, y)
        s = x if trial % 7 == 0 else torch.sigmoid(x)
        hard = (s > 0.5).numpy().astype(int)
        ref = dict(auroc=M.roc_auc_score(y.numpy(), s.numpy()), ap=M.average_precision_score(y.numpy(), s.numpy()),
                   mcc=M.matthews_corrcoef(y.numpy(), hard), precision=M.precision_score(y.numpy(), hard, zero_division=0),
                   recall=M.recall_score(y.numpy(), hard, zero_division=0))
        for k, v in ref.items():
            assert abs(m[k] - v) < 2e-6, (trial, k, m[k], v)


def test_metrics_restatement_degenerate_batches():
    import math

    x = torch.tensor([2.0, -1.0, 0.3])
    no_pos = R.batch_metrics(x, torch.zeros(3, dtype=torch.long))
    assert no_pos["auroc"] == 0.0 and math.isnan(no_pos["ap"]) and no_pos["mcc"] == 0.0 and no_pos["recall"] == 0.0
    no_neg = R.batch_metrics(x, torch.ones(3, dtype=torch.long))
    assert no_neg["auroc"] == 0.0 and abs(no_neg["ap"] - 1.0) < 1e-6 and no_neg["precision"] == 1.0
    assert R.batch_metrics(torch.tensor([0.2, 0.7]), torch.tensor([0, 1]))["confusion"] == (1, 0, 1, 0)  # no sigmoid inside [0,1]
    assert R.batch_metrics(torch.tensor([0.2, 1.7]), torch.tensor([0, 1]))["confusion"] == (1, 1, 0, 0)  # sigmoid(0.2) > 0.5


# ---- production-mode mask generator: Philox4x32-10 restatement against the Random123 known-answer vectors -----------------------
def test_philox_known_answer_vectors():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
           ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
           ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]
    for ctr, key, want in kat:
        assert tuple(R.philox4x32_10(ctr, key)) == want


def test_mask_restatement_is_a_scaled_bernoulli_and_respects_offsets():
    a, b = R.draw_masks([4000, 10], [0.7, 0.5], seed=11, offset=0)
    assert set(a.unique().tolist()) <= {0.0, float(torch.tensor(1.0) / torch.tensor(0.7))} and abs(float((a != 0).float().mean()) - 0.7) < 0.03
    assert set(b.unique().tolist()) <= {0.0, 2.0}
    a2, = R.draw_masks([4000], [0.7], seed=11, offset=1)  # one counter later = the same stream shifted by 4 draws
    assert torch.equal(a[4:], a2[:-4])
    rows, = R.draw_masks([12], [0.5], seed=3, offset=0, row_lens=[4])
    assert all(len(set(rows[i:i + 4].tolist())) == 1 for i in (0, 4, 8))


# ---- input feeding (intrepppid_b200/feed.py) -----------------------------------------------------------------
class _ToyPairs(torch.utils.data.Dataset):
    """Samples shaped like IntrepppidDataset.__getitem__ (data/ppi_oma.py:489-503): five int64 id rows + a label."""

    def __init__(self, n, T, V, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.rows = torch.randint(0, V, (n, 5, T), generator=g)
        self.y = torch.randint(0, 2, (n,), generator=g)

    def __len__(self):
        return self.rows.shape[0]

    def __getitem__(self, i):
        r = self.rows[i]
        return r[0].long(), r[1].long(), r[2].long(), r[3].long(), r[4].long(), self.y[i].long()


def test_narrow_collate_matches_the_default_collate():
    from torch.utils.data import DataLoader

    from intrepppid_b200 import feed

    assert feed.token_dtype_for(250) == torch.uint8 and feed.token_dtype_for(256) == torch.uint8
    assert feed.token_dtype_for(257) == torch.int16 and feed.token_dtype_for(32768) == torch.int16
    assert feed.token_dtype_for(32769) == torch.int32
    for V, dt in ((250, torch.uint8), (5000, torch.int16)):
        ds = _ToyPairs(23, 40, V, seed=V)
        ref = list(DataLoader(ds, batch_size=8, shuffle=False))                      # the reference's loader (ppi_oma.py:611-620)
        got = list(DataLoader(ds, batch_size=8, shuffle=False, collate_fn=feed.narrow_collate(V), num_workers=2))  # through worker IPC
        assert len(ref) == len(got) == 3
        for r, g in zip(ref, got):
            assert isinstance(g, feed.PackedBatch) and len(g) == 6 and g.tokens.dtype == dt
            assert g.tokens.shape == (5, r[0].shape[0], 40) and g.tokens.is_contiguous()
            for i in range(5):
                assert torch.equal(g[i].long(), r[i]) and g[i].data_ptr() == g.tokens[i].data_ptr()
            assert torch.equal(g[5], r[5]) and g[5].dtype == torch.int64


def test_packing_refuses_ids_that_would_wrap():
    from intrepppid_b200 import feed

    seqs = [torch.randint(0, 250, (3, 9)) for _ in range(5)]
    y = torch.zeros(3, dtype=torch.long)
    feed.pack_batch(seqs, y, 250)
    seqs[3][1, 4] = 250
    with pytest.raises(IndexError):
        feed.pack_batch(seqs, y, 250)       # F.embedding would raise on it; a silent uint8 wrap would turn it into id 250 % 256
    seqs[3][1, 4] = -1
    with pytest.raises(IndexError):
        feed.pack_batch(seqs, y, 250)
    with pytest.raises(IndexError):
        feed.narrow_collate(250)([tuple([s[1] for s in seqs] + [0])])
    with pytest.raises(ValueError):
        feed.pack_batch(seqs[:4], y, 250)
    with pytest.raises(TypeError):
        feed.pack_batch([s.float() for s in seqs], y, 250)
    with pytest.raises(RuntimeError):
        feed.DeviceFeeder([], "cpu", 250)   # no CPU path
