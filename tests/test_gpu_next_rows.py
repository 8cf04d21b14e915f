"""GPU parity for the SURVEY 8f rows around the hot path, through the C ABI:
  * ib200_adamw_step / FusedAdamW against torch.optim.AdamW (the reference's optimizer, e2e_triplet.py:231-255);
  * narrowed token ids (IB200_TOK_I32 / I16 / U8) against the int64 ids the reference ships -- bit-identical results;
  * batch-of-one inference with the embedding cache (cli/infer.py:196-225) against the oracle's per-row batch-1 loop and against
    rows recorded from the reference itself (tests/golden/infer_rows.pt);
  * ib200_batch_metrics against the restatement of torchmetrics' binary metrics; ib200_draw_masks bit-exact against the Philox
    restatement; the regression test for launch sets with more than 32 groups."""
import copy

import pytest
import torch

from conftest import rel_l2
from helpers import build_product, product_masks
from oracle import restatement as R

pytestmark = pytest.mark.gpu


# ---- AdamW ---------------------------------------------------------------------------------------------------------------------
def _problem(seed, n_tensors):
    g = torch.Generator().manual_seed(seed)
    base = [(1,), (7,), (33, 5), (256, 64), (2049,), (64,), (3, 3, 3), (250, 64), (256, 128), (4097,)]
    shapes = [base[i % len(base)] for i in range(n_tensors)]
    params = [torch.randn(s, generator=g) for s in shapes]
    grads = [[torch.randn(s, generator=g) * 0.1 for s in shapes] for _ in range(8)]
    return params, grads


@pytest.mark.parametrize("n_tensors", [3, 23, 40])  # 40 crosses the 32-tensor pointer table: two launches
def test_fused_adamw_matches_torch_adamw(n_tensors):
    from intrepppid_b200 import FusedAdamW, _lib

    params, grads = _problem(11, n_tensors)
    ref = [torch.nn.Parameter(p.clone().double()) for p in params]           # fp64 gold
    ref32 = [torch.nn.Parameter(p.clone()) for p in params]                  # torch's own fp32 CPU path
    mine = [torch.nn.Parameter(p.clone().cuda()) for p in params]
    o_ref, o32 = torch.optim.AdamW(ref, lr=1e-2), torch.optim.AdamW(ref32, lr=1e-2, foreach=False)
    o_mine = FusedAdamW(mine, lr=1e-2)
    skip = 1 if n_tensors > 1 else None  # a parameter that never receives a gradient (the dead projection.* tensors, Q10)
    l0 = _lib.launch_count()
    for gs in grads:
        for k, g in enumerate(gs):
            if k == skip:
                continue
            ref[k].grad, ref32[k].grad, mine[k].grad = g.double(), g.clone(), g.cuda()
        o_ref.step(); o32.step(); o_mine.step()
    torch.cuda.synchronize()
    assert _lib.launch_count() - l0 == len(grads) * (1 if n_tensors <= 33 else 2)
    for k in range(n_tensors):
        got = mine[k].detach().cpu()
        if k == skip:
            assert torch.equal(got, params[k]) and mine[k] not in o_mine.state
            continue
        e64, e32 = rel_l2(got, ref[k]), rel_l2(ref32[k], ref[k])
        assert e64 < 1e-6 and e64 < 4 * e32 + 1e-7, (k, e64, e32)  # as close to fp64 as torch's fp32 path is
        st = o_mine.state[mine[k]]
        assert int(st["step"]) == len(grads)
        assert rel_l2(st["exp_avg"], o_ref.state[ref[k]]["exp_avg"]) < 1e-6
        assert rel_l2(st["exp_avg_sq"], o_ref.state[ref[k]]["exp_avg_sq"]) < 1e-6


def test_fused_adamw_state_dict_interchanges_with_torch_and_grad_scale_folds_the_dp_mean():
    from intrepppid_b200 import FusedAdamW

    params, grads = _problem(12, 6)
    t_params = [torch.nn.Parameter(p.clone().cuda()) for p in params]
    t_opt = torch.optim.AdamW(t_params, lr=3e-3, weight_decay=0.05, betas=(0.8, 0.99), foreach=False)
    for gs in grads[:3]:
        for p, g in zip(t_params, gs):
            p.grad = g.cuda()
        t_opt.step()
    mine = [torch.nn.Parameter(p.detach().clone()) for p in t_params]
    m_opt = FusedAdamW(mine, lr=1.0)
    m_opt.load_state_dict(copy.deepcopy(t_opt.state_dict()))  # a checkpoint written by the reference's optimizer
    assert m_opt.param_groups[0]["lr"] == 3e-3 and m_opt.param_groups[0]["betas"] == (0.8, 0.99)
    m_opt.grad_scale = 0.25  # gradients arrive as a SUM over 4 ranks
    for gs in grads[3:]:
        for p, q, g in zip(t_params, mine, gs):
            p.grad, q.grad = g.cuda(), 4.0 * g.cuda()
        t_opt.step(); m_opt.step()
    for p, q in zip(t_params, mine):
        assert rel_l2(q, p) < 1e-6
    t_opt.load_state_dict(copy.deepcopy(m_opt.state_dict()))  # and back


def test_training_run_with_fused_adamw_tracks_the_oracle_run():
    """Three full optimisation steps (step + backward + AdamW) of the product against the fp64 oracle driven by torch.optim.AdamW."""
    import intrepppid_b200 as ib

    B, T, V, E, L = 12, 48, 60, 64, 2
    P = R.init_params(vocab=V, E=E, L=L, seed=3)
    net = build_product(P, L=L, bi="last", p_emb=0.3)
    net.train()
    opt = ib.FusedAdamW([p for p in net.parameters() if p.requires_grad], lr=1e-2)
    Pd = {k: v.double().clone().requires_grad_(True) for k, v in P.items()}
    o_ref = torch.optim.AdamW(list(Pd.values()), lr=1e-2)
    for it in range(3):
        batch = list(R.synthetic_batch(B, T, V, seed=100 + it, padded=True))
        masks = R.draw_step_masks(B, V, E, emb_droprate=0.3, rnn_droprate=0.3, do_rate=0.3, seed=200 + it)
        opt.zero_grad(set_to_none=True)
        loss = net.step([t.cuda() for t in batch], "train", masks=product_masks(masks, 0.3))
        loss.backward()
        opt.step()
        md = R.StepMasks(masks.emb_row_keep, *(None if m is None else m.double() for m in
                                               (masks.whh_mask, masks.fc1_w, masks.do1, masks.do2, masks.fc2_w)))
        o_ref.zero_grad(set_to_none=True)
        out = R.step(batch, Pd, num_layers=L, bi_reduce="last", beta_classifier=2.0, training=True, emb_droprate=0.3,
                     use_projection=False, masks=md)
        out.loss.backward()
        o_ref.step()
        assert abs(float(loss) - float(out.loss)) < 1e-4 * max(1.0, abs(float(out.loss))), it
    from helpers import module_key

    named = dict(net.named_parameters())
    for n, v in Pd.items():
        # the dead top-layer forward chain gets zero gradients (Q16): AdamW still decays those weights, identically on both sides
        assert rel_l2(named[module_key(n)].detach().cpu(), v.detach()) < 2e-4, n  # Adam normalises: tiny grad errors stay tiny


# ---- narrowed token ids -----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.int32, torch.int16, torch.uint8])
def test_narrow_token_ids_are_bit_identical_to_int64(dtype):
    B, T, V, E, L = 9, 70, 250, 64, 2
    P = R.init_params(vocab=V, E=E, L=L, seed=4)
    batch = list(R.synthetic_batch(B, T, V, seed=9, padded=True))
    masks = R.draw_step_masks(B, V, E, emb_droprate=0.3, rnn_droprate=0.3, do_rate=0.3, seed=10)
    outs = []
    for dt in (torch.int64, dtype):
        net = build_product(P, L=L, bi="mean", p_emb=0.3).train()
        cb = [t.cuda().to(dt) for t in batch[:5]] + [batch[5].cuda()]
        loss = net.step(cb, "train", masks=product_masks(masks, 0.3))
        loss.backward()
        torch.cuda.synchronize()
        outs.append((net.last_step, {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}))
    (a, ga), (b, gb) = outs
    assert torch.equal(a["lengths"], b["lengths"]) and torch.equal(a["z"], b["z"]) and torch.equal(a["loss"], b["loss"])
    assert ga.keys() == gb.keys() and all(torch.equal(ga[n], gb[n]) for n in ga)


def test_unsupported_token_dtype_raises():
    from intrepppid_b200._lib import Cfg, lib

    assert lib().ib200_workspace_bytes(Cfg(1, 2, 8, 10, 32, 1, 0, 0, 0, 7)) == 0
    net = build_product(R.init_params(vocab=20, E=32, L=1, seed=0), L=1, bi="last").eval()
    z = net.encoder(torch.randint(1, 20, (2, 8), device="cuda").to(torch.int8))  # not an id type of the ABI: widened on the host
    assert z.shape == (2, 32)
    with pytest.raises(RuntimeError, match="token ids must be"):  # (the raw op; intrepppid_b200.ops re-raises "[ib200]" errors as IB200Error)
        torch.ops.intrepppid_b200.encoder_fwd(torch.ones(1, 2, 8, device="cuda"), net.encoder.embedder.weight,
                                              net.encoder.encoder.rnn.ordered(), None, None, 1, 0, 0, False)


# ---- batch-of-one inference with the embedding cache ------------------------------------------------------------------------
def _proteins(M, T, V, seed):
    g = torch.Generator().manual_seed(seed)
    toks = {}
    for m in range(M):
        n = int(torch.randint(1, T + 1, (1,), generator=g)) if m >= 6 else (T, T, 1, 17, 17, 17)[m]
        row = torch.zeros(T, dtype=torch.long)
        row[:n] = torch.randint(1, V, (n,), generator=g)
        toks[f"P{m:03d}"] = row
    toks["P007"][3] = 0  # interior <unk>
    return toks


@pytest.mark.parametrize("dtype", [torch.int64, torch.int32, torch.int16, torch.uint8])
def test_batch1_lengths_match_the_oracle_per_sequence(dtype):
    """ib200_sequence_lengths against the oracle's own truncation on a batch of one, sequence by sequence (interior zero, all-zero
    and partially zero vocabulary rows, one full-length and one single-token protein), for every id type of the ABI."""
    from intrepppid_b200.infer import batch1_lengths

    M, T, V, E = 24, 40, 30, 32
    P = R.init_params(vocab=V, E=E, L=1, seed=1)
    P["emb"][7].zero_()       # an all-zero vocabulary row besides the padding row
    P["emb"][9, :5].zero_()   # a partially zero row
    g = torch.Generator().manual_seed(2)
    tok = torch.randint(1, V, (M, T), generator=g)
    lens = torch.randint(1, T + 1, (M,), generator=g)
    lens[:4] = torch.tensor([T, T, 1, 5])
    for m in range(M):
        tok[m, lens[m]:] = 0
    tok[6, 2] = 0  # interior <unk>: T1 is a COUNT, so the slice loses the last real token
    tok[8, :3] = 7
    tok[10, :] = 9
    tok[10, 30:] = 0
    t1, te = batch1_lengths(tok.to(dtype).cuda(), P["emb"].cuda())
    assert t1.dtype == te.dtype == torch.int64 and t1.is_cuda
    for m in range(M):
        _, info = R.encoder_forward(tok[m:m + 1], P, num_layers=1, bi_reduce="last", training=False)
        assert (int(t1[m]), int(te[m])) == (info.T1, info.T_eff), m
    # a vocabulary past the 48 KB default of dynamic shared memory (the opt-in path of the histogram)
    V2 = 9000
    emb2 = torch.randn(V2, 32)
    emb2[0].zero_()
    tok2 = torch.randint(1, V2, (5, 64), generator=g)
    tok2[1, 50:] = 0
    t1b, teb = batch1_lengths(tok2.cuda(), emb2.cuda())
    assert t1b.tolist() == [64, 50, 64, 64, 64] and teb.tolist() == [64, 50, 64, 64, 64]


@pytest.mark.parametrize("bi,L,E", [("last", 2, 64), ("mean", 1, 32), ("max", 2, 96)])
def test_infer_pairs_equals_the_reference_batch_of_one_loop(bi, L, E, tmp_path):
    from intrepppid_b200 import infer

    M, T, V = 45, 60, 40
    P = R.init_params(vocab=V, E=E, L=L, seed=21)
    toks = _proteins(M, T, V, 22)
    names = sorted(toks)
    g = torch.Generator().manual_seed(23)
    rows = [(f"itx{i}", names[int(torch.randint(0, M, (1,), generator=g))], names[int(torch.randint(0, M, (1,), generator=g))])
            for i in range(120)]
    rows[5] = ("itx5", "P001", "NOT_THERE")
    rows[6] = ("itx6", "P003", "P003")
    ref = R.infer_from_csv_rows({k: v for k, v in toks.items()}, rows, {k: v.double() for k, v in P.items()}, num_layers=L, bi_reduce=bi)
    net = build_product(P, L=L, bi=bi).eval()
    missing = []
    got = infer.infer_pairs(net, toks, rows, on_missing=lambda *r: missing.append(r))
    assert [i for i, _ in got] == [i for i, _ in ref] and missing == [("itx5", "P001", "NOT_THERE")]
    worst = max(abs(a - b) for (_, a), (_, b) in zip(got, ref))
    assert worst < 2e-5, worst  # probabilities; logits agree to ~1e-5 relative in fp32 mode

    # embeddings: each row equals a batch-of-one encoder call, and differs from the mixed-length batched call (no packing)
    tok = torch.stack([toks[n] for n in names]).cuda()
    z = infer.embed_batch1(net, tok, max_groups=4)
    for m in (0, 2, 7, 20, 44):
        z1, _ = R.encoder_forward(tok[m:m + 1].cpu(), {k: v.double() for k, v in P.items()}, num_layers=L, bi_reduce=bi, training=False)
        assert rel_l2(z[m:m + 1], z1) < 1e-4, m
        assert rel_l2(net.encoder(tok[m:m + 1]), z1) < 1e-4, m
    assert rel_l2(net.embed(tok)[2:3], z[2:3]) > 1e-3  # the 1-token protein stepped through 59 pads in the mixed batch

    # file-level mirror of `infer from_csv`
    src, dst = tmp_path / "pairs.csv", tmp_path / "out.csv"
    src.write_text("".join(f"{i},{a},{b}\n" for i, a, b in rows))
    assert infer.from_csv(net, toks, str(src), str(dst)) == len(ref)
    lines = dst.read_text().strip().splitlines()
    assert len(lines) == len(ref) and lines[0].split(",")[0] == ref[0][0]
    assert abs(float(lines[0].split(",")[1]) - ref[0][1]) < 2e-5


def test_embed_batch1_rejects_an_all_pad_protein():
    from intrepppid_b200 import infer

    net = build_product(R.init_params(vocab=20, E=32, L=1, seed=0), L=1, bi="last").eval()
    tok = torch.randint(1, 20, (3, 8), device="cuda")
    tok[1] = 0
    with pytest.raises(RuntimeError, match="sequence length"):
        infer.embed_batch1(net, tok)
    with pytest.raises(RuntimeError, match="eval"):
        infer.embed_batch1(net.train(), tok)


def test_many_groups_keep_their_own_lengths_when_the_workspace_is_recycled():
    """Regression: the length counters of ALL groups are reset by every call (G > 32 used to keep stale maxima from the previous
    call that owned the same workspace memory), and a padded launch equals the unpadded one (pad replicas read the same row 0)."""
    G, B, T, V = 96, 4, 64, 250
    net = build_product(R.init_params(vocab=V, E=64, L=2, seed=2), L=2, bi="last").eval()
    net.encoder.check_lengths = False
    g = torch.Generator().manual_seed(3)
    long = torch.randint(1, V, (G, B, T), generator=g).cuda()
    short = long.clone()
    glen = torch.randint(1, T // 2, (G,), generator=g)
    for i in range(G):
        short[i, :, glen[i]:] = 0
    with torch.no_grad():
        net.encoder.forward_groups(long, draw=False)
        z_short = net.encoder.forward_groups(short, draw=False)   # same shapes: the caching allocator hands back the same workspace
        lens = net.encoder.last_lengths.cpu()
        assert torch.equal(lens[0], glen.int()) and torch.equal(lens[1], glen.int())
        for i in (0, 40, 95):
            z1 = net.encoder(short[i, :, :int(glen[i])].contiguous())
            assert rel_l2(z_short[i], z1) < 2e-5, i  # no pad is stepped: same as a call on the truncated rows (other CTA shape)


# ---- per-step metrics ---------------------------------------------------------------------------------------------------------
def test_batch_metrics_kernel_matches_the_restatement():
    import math

    from intrepppid_b200 import ops

    g = torch.Generator().manual_seed(7)
    cases = []
    for trial in range(40):
        B = (1, 2, 80, 1023, 1024)[trial] if trial < 5 else int(torch.randint(2, 1025, (1,), generator=g))
        y = torch.randint(0, 2, (B,), generator=g)
        x = torch.randn(B, generator=g) * 3
        if trial % 3 == 0:
            x = torch.round(x * 2) / 2
        if trial % 7 == 0:
            x = torch.rand(B, generator=g)
        if trial == 10:
            y.zero_()
        if trial == 11:
            y.fill_(1)
        cases.append((x, y))
    for k, (x, y) in enumerate(cases):
        ref = R.batch_metrics(x, y)
        m, conf = ops.batch_metrics(x.cuda(), y.cuda())
        m, conf = m.cpu().tolist(), tuple(conf.cpu().tolist())
        assert conf == ref["confusion"], (k, conf, ref["confusion"])
        for got, name in zip(m, ("auroc", "ap", "mcc", "precision", "recall")):
            want = ref[name]
            assert (math.isnan(got) and math.isnan(want)) or abs(got - want) < 2e-5, (k, name, got, want)
    with pytest.raises(Exception):
        ops.batch_metrics(torch.zeros(1025, device="cuda"), torch.zeros(1025, dtype=torch.long, device="cuda"))


def test_step_logs_the_five_reference_metrics_without_torchmetrics():
    B, T, V = 16, 40, 60
    net = build_product(R.init_params(vocab=V, E=64, L=2, seed=5), L=2, bi="last").train()
    logged = {}
    net.log = lambda name, value, **kw: logged.__setitem__(name, value)
    batch = [t.cuda() for t in R.synthetic_batch(B, T, V, seed=6, padded=True)]
    net.step(batch, "val")
    ref = R.batch_metrics(net.last_step["y_hat"].cpu(), batch[5].cpu())
    for key, name in (("val_auroc", "auroc"), ("val_ap", "ap"), ("val_mcc", "mcc"), ("val_precision", "precision"), ("val_rec", "recall")):
        assert abs(float(logged[key]) - ref[name]) < 2e-5, key
    assert {"val_loss", "val_classifier_loss", "val_triplet_loss", "val_loss_step"} <= set(logged)


# ---- production-mode masks ------------------------------------------------------------------------------------------------------
def test_draw_masks_kernel_is_bit_exact_with_the_philox_restatement():
    from intrepppid_b200 import ops

    specs = [((5, 250), 0.7, 0), ((3, 128, 32), 0.7, 0), ((32, 64), 0.5, 0), ((9, 32), 0.9, 0), ((1, 32), 0.7, 0), ((7, 16), 0.5, 16),
             ((3,), 1.0, 0), ((2, 5), 0.25, 0), ((6, 6), 0.7, 0), ((1,), 0.5, 0)]  # 10 masks: two launches
    got, used = ops.draw_masks(specs, "cuda", seed=1234567890123, offset=77)
    numels = [int(torch.Size(s).numel()) for s, _, _ in specs]
    # the second launch starts where the first one stopped: restate it as two calls of 8 and 2 masks
    first = R.draw_masks(numels[:8], [k for _, k, _ in specs[:8]], 1234567890123, 77, row_lens=[r for _, _, r in specs[:8]])
    used_first = sum(((n + max(r, 1) - 1) // max(r, 1) + 3) // 4 for n, (_, _, r) in zip(numels[:8], specs[:8]))
    second = R.draw_masks(numels[8:], [k for _, k, _ in specs[8:]], 1234567890123, 77 + used_first, row_lens=[r for _, _, r in specs[8:]])
    for g, w, (shape, _, _) in zip(got, first + second, specs):
        assert g.shape == torch.Size(shape) and torch.equal(g.cpu().reshape(-1), w), shape
    assert used == used_first + sum((n + 3) // 4 for n in numels[8:])
    big, _ = ops.draw_masks([((5, 256, 64), 0.7, 0)], "cuda", seed=5, offset=0)
    assert abs(float((big[0] != 0).float().mean()) - 0.7) < 0.01
    with pytest.raises(Exception):
        ops.draw_masks([((4,), 0.0, 0)], "cuda", seed=1, offset=0)


def test_step_draws_its_masks_in_one_launch_and_is_reproducible():
    from intrepppid_b200 import _lib

    B, T, V = 8, 40, 60
    batch = [t.cuda() for t in R.synthetic_batch(B, T, V, seed=6, padded=True)]

    def run(fused):
        torch.manual_seed(123)
        net = build_product(R.init_params(vocab=V, E=64, L=2, seed=5), L=2, bi="last").train()
        net.fused_masks = fused
        net.compute_metrics = False
        net.step(batch, "train")
        l0 = _lib.launch_count()
        losses = [float(net.step(batch, "train").detach()) for _ in range(2)]
        return losses, (_lib.launch_count() - l0) // 2, net

    a, launches_fused, net = run(True)
    b, _, _ = run(True)
    c, launches_torch, _ = run(False)
    assert a == b and a[0] != a[1]            # same seed -> same masks; successive steps -> fresh masks
    assert all(0.0 < x < 10.0 for x in a + c)
    assert launches_fused == launches_torch + 1  # the library launches one more kernel (torch launches twelve fewer)
    assert net._mask_offset > 0
    net.eval()
    assert net._draw_step_masks(B) == (None, None, None)


def test_infer_pairs_vs_reference_golden():
    """The embedding-cache path against numbers recorded from the reference's own batch-of-one row loop (tests/golden/infer_rows.pt)."""
    from conftest import load_golden
    from intrepppid_b200 import infer

    gold = load_golden("infer_rows")
    c = gold["config"]
    net = build_product(gold["params"], L=c["L"], bi=c["bi"], use_projection=True).eval()
    got = infer.infer_pairs(net, gold["tokens"], gold["rows"])
    assert [i for i, _ in got] == [i for i, _ in gold["scored"]]
    assert max(abs(a - b) for (_, a), (_, b) in zip(got, gold["scored"])) < 2e-5
    names = sorted(gold["tokens"])
    z = infer.embed_batch1(net, torch.stack([gold["tokens"][n] for n in names]).cuda())
    for k, n in enumerate(names):
        assert rel_l2(z[k], gold["z"][n]) < 1e-4, n


def test_device_feeder_step_is_bit_identical_to_int64_batches():
    """a step fed by intrepppid_b200.feed (narrow ids packed in the loader workers, pinned staging, H2D on a copy
    stream, ragged last batch) gives bit-identical losses / gradients to the reference's int64 batches copied synchronously."""
    from torch.utils.data import DataLoader, Dataset

    from intrepppid_b200 import feed

    class Pairs(Dataset):
        def __init__(self):
            g = torch.Generator().manual_seed(77)
            self.rows = torch.randint(1, 250, (21, 5, 64), generator=g)
            self.rows[:, :, 50:] *= (torch.rand(21, 5, 14, generator=g) > 0.5)  # interior zeros and ragged tails
            self.y = torch.randint(0, 2, (21,), generator=g)

        def __len__(self):
            return 21

        def __getitem__(self, i):
            r = self.rows[i]
            return r[0], r[1], r[2], r[3], r[4], self.y[i]

    ds = Pairs()
    P = R.init_params(vocab=250, E=64, L=2, seed=4)
    masks = R.draw_step_masks(8, 250, 64, emb_droprate=0.3, rnn_droprate=0.3, do_rate=0.3, seed=3)

    def run(batches):
        out = []
        for batch in batches:
            net = build_product(P, L=2, bi="last").train()
            B = batch[5].shape[0]
            m = product_masks(masks, 0.3)
            m.head = (m.head[0], m.head[1][:B], m.head[2][:B], m.head[3])
            loss = net.step([t.cuda() for t in batch] if not batch[0].is_cuda else batch, "train", masks=m)
            loss.backward()
            out.append((loss.detach().clone(), net.last_step["lengths"].clone(),
                        {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}))
        return out

    ref = run(DataLoader(ds, batch_size=8, shuffle=False))                                    # int64, 3 batches (8, 8, 5)
    feeder = feed.feeding_dataloader(ds, 8, 250, "cuda", num_workers=2, shuffle=False)
    got = run(feeder)
    assert len(ref) == len(got) == 3 and feeder.h2d_bytes == 21 * 5 * 64 + 21 * 8
    for (l0, n0, g0), (l1, n1, g1) in zip(ref, got):
        assert torch.equal(l0, l1) and torch.equal(n0, n1)
        assert g0.keys() == g1.keys() and all(torch.equal(g0[k], g1[k]) for k in g0)
    # the reference's default-collated int64 tuples are accepted too (packed on the consumer thread)
    got2 = run(feed.DeviceFeeder(DataLoader(ds, batch_size=8, shuffle=False), "cuda", 250))
    assert all(torch.equal(a[0], b[0]) for a, b in zip(ref, got2))
