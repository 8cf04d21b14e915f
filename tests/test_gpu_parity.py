"""GPU parity: the CUDA path (through the C ABI, via the reference-shaped modules) against the CPU oracle on the same seeded
inputs and the same explicit masks, and against the golden vectors produced by the unmodified reference.

Gates (BASELINE.json north_star / SURVEY 8d): fp32 mode -- embeddings, losses and every gradient within 1e-4 relative L2;
bf16 mode -- within 2e-2; truncation lengths (T1, T_eff) bit-exact; dead-chain gradients exactly zero."""
import pytest
import torch

from conftest import GOLDEN_CASES, assert_grad_close, load_golden, rel_l2
from helpers import build_product, product_masks, run_oracle_step, run_product_step
from oracle import restatement as R

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16": 2e-2}


def _golden_inputs(gold):
    c = gold["config"]
    mk = gold["masks"]
    masks = R.StepMasks(mk["emb_row_keep"], mk["whh_mask"], mk["fc1_w"], mk["do1"], mk["do2"], mk["fc2_w"])
    batch = gold["tokens"] + [gold["y"]]
    kw = dict(L=c["L"], bi=c["bi"], beta=c["beta"], use_projection=c["proj"], p_emb=c["p_emb"])
    return c, masks, batch, kw


def _check(got, ref, tol, grads=True):
    assert torch.equal(got["lengths"].cpu(), ref["lengths"]), "T1 / T_eff must be bit-exact"
    for g in range(5):
        assert rel_l2(got["z"][g], ref["z"][g]) < tol, f"z[{g}]"
    for k in ("loss", "classifier_loss", "triplet_loss"):
        assert abs(float(got[k]) - float(ref[k])) <= tol * max(1.0, abs(float(ref[k]))), k
    assert rel_l2(got["y_hat"], ref["y_hat"]) < tol or float((got["y_hat"].cpu().double() - ref["y_hat"].double()).abs().max()) < tol
    if grads:
        for n, g in ref["grads"].items():
            assert got["grads"][n] is not None, f"missing gradient {n}"
            assert_grad_close(got["grads"][n], g, tol, n)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_training_step_vs_reference_golden(name):
    """fp32 mode against numbers produced by the reference itself (tests/golden, oracle/make_golden.py)."""
    gold = load_golden(name)
    c, masks, batch, kw = _golden_inputs(gold)
    got = run_product_step(gold["params"], batch, masks, p_rnn=c["p_rnn"], p_do=c["p_do"], **kw)
    t = gold["train"]
    ref = {"lengths": run_oracle_step(gold["params"], batch, masks, **kw)["lengths"], "z": torch.stack(t["z"]), "loss": t["loss"],
           "classifier_loss": t["classifier_loss"], "triplet_loss": t["triplet_loss"], "y_hat": t["y_hat"], "grads": t["grads"]}
    _check(got, ref, 2e-4)  # the golden numbers are fp32 themselves (oracle noise ~5e-6..2e-5 on small-norm tensors)


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_training_step_vs_fp64_oracle(name, precision):
    gold = load_golden(name)
    c, masks, batch, kw = _golden_inputs(gold)
    got = run_product_step(gold["params"], batch, masks, p_rnn=c["p_rnn"], p_do=c["p_do"], precision=precision, **kw)
    ref = run_oracle_step(gold["params"], batch, masks, **kw)
    _check(got, ref, TOL[precision])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_eval_forward_vs_reference_golden(name):
    gold = load_golden(name)
    c = gold["config"]
    net = build_product(gold["params"], L=c["L"], bi=c["bi"], beta=c["beta"], use_projection=c["proj"], p_emb=c["p_emb"],
                        p_rnn=c["p_rnn"], p_do=c["p_do"]).eval()
    with torch.no_grad():
        logits = net(gold["tokens"][0].cuda(), gold["tokens"][1].cuda()).cpu()
    assert logits.shape == gold["eval"]["logits"].shape
    assert float((logits - gold["eval"]["logits"]).abs().max()) < 1e-4


@pytest.mark.parametrize("E,L,bi,B,T", [(64, 2, "last", 13, 257), (32, 3, "mean", 9, 100), (64, 1, "max", 8, 64)])
def test_seeded_cases_vs_oracle(E, L, bi, B, T):
    """Ragged batches whose sizes are not multiples of the kernel tiles (8 sequences per CTA, 32/128-row GEMM tiles)."""
    V = 97
    P = R.init_params(vocab=V, E=E, L=L, seed=21)
    batch = list(R.synthetic_batch(B, T, V, seed=22, padded=True))
    masks = R.draw_step_masks(B, V, E, emb_droprate=0.3, rnn_droprate=0.3, do_rate=0.3, seed=23)
    kw = dict(L=L, bi=bi, beta=2.0, use_projection=False, p_emb=0.3)
    got = run_product_step(P, batch, masks, p_rnn=0.3, p_do=0.3, **kw)
    ref = run_oracle_step(P, batch, masks, **kw)
    _check(got, ref, 1e-4)


def test_all_pad_batch_raises_like_the_reference():
    """nn.LSTM raises on T = 0 (quirk Q13).  check_lengths="sync" raises before returning; the default records the condition on the
    device and raises it without a host sync in the step: at ops.check_pending(sync=True) or on a later call."""
    from intrepppid_b200 import ops

    P = R.init_params(E=32, L=2)
    net = build_product(P, L=2, bi="last").eval()
    pads = torch.zeros(4, 16, dtype=torch.long, device="cuda")
    net.encoder.check_lengths = "sync"
    with pytest.raises(RuntimeError, match="sequence length"):
        net.encoder(pads)
    net.encoder.check_lengths = True
    with torch.no_grad():
        net.encoder(pads)  # returns: nothing has been read on the host yet
    with pytest.raises(RuntimeError, match="sequence length"):
        ops.check_pending(sync=True)
    ops.check_pending(sync=True)  # the error was consumed
    with torch.no_grad():
        net.encoder(pads)
        torch.cuda.synchronize()
        with pytest.raises(RuntimeError, match="sequence length"):
            net.encoder(torch.ones(4, 16, dtype=torch.long, device="cuda"))  # the next call surfaces it
    ops.check_pending(sync=True)


def test_all_pad_row_is_legal_and_bias_driven():
    """Q13: a row of pads inside a batch yields a non-zero embedding."""
    P = R.init_params(E=32, L=2)
    net = build_product(P, L=2, bi="last").eval()
    x = torch.randint(1, 250, (4, 16))
    x[2] = 0
    with torch.no_grad():
        z = net.encoder(x.cuda()).cpu()
        zr, _ = R.encoder_forward(x, P, num_layers=2, bi_reduce="last", training=False)
    assert float(z[2].norm()) > 0
    assert rel_l2(z, zr) < 1e-4


def test_concat_rejected():
    P = R.init_params(E=32, L=2)
    net = build_product(P, L=2, bi="last")
    net.encoder.encoder.bi_reduce = "concat"
    with pytest.raises(ValueError):
        net.encoder(torch.ones(2, 8, dtype=torch.long, device="cuda"))


def test_cpu_tensors_are_refused():
    """No CPU fallback: the ops raise instead of silently running elsewhere."""
    from intrepppid_b200._lib import IB200Error

    P = R.init_params(E=32, L=2)
    net = build_product(P, L=2, bi="last", device="cpu").eval()
    with pytest.raises(IB200Error):
        net.encoder(torch.ones(2, 8, dtype=torch.long))


def test_pair_score_block_vs_oracle():
    """Config 4 parity: head + sigmoid on all pairs of a block of embeddings (explicit indices and implicit upper triangle)."""
    P = R.init_params(E=64, L=2, seed=5)
    net = build_product(P, L=2, bi="last").eval()
    z = torch.randn(96, 64)
    ia, ib = torch.triu_indices(96, 96)
    with torch.no_grad():
        ref = torch.sigmoid(R.mlp_head(z[ia], z[ib], P).squeeze(1))
    got_tri = net.score_pairs(z.cuda()).cpu()
    got_idx = net.score_pairs(z.cuda(), ia.cuda(), ib.cuda()).cpu()
    assert float((got_tri - ref).abs().max()) < 1e-5
    assert float((got_idx - ref).abs().max()) < 1e-5


def test_headline_shape_determinism_and_group_fusion():
    """Full-size (B=80, T=1500) properties that need no oracle: the step is deterministic run to run, and the per-group
    embeddings of the fused G=5 launch equal five separate encoder calls with the same masks (bit for bit)."""
    import intrepppid_b200 as ib

    torch.manual_seed(0)
    net = ib.intrepppid_network(1).cuda().train()
    batch = [t.cuda() for t in R.synthetic_batch(80, 1500, 250, seed=1234)]
    m = R.draw_step_masks(80, 250, 64, emb_droprate=0.3, rnn_droprate=0.3, do_rate=0.3, seed=5)
    pm = product_masks(m, 0.3)
    emb_w = net.encoder.embedder.weight
    l1 = net.step(batch, "train", masks=pm)
    l1.backward()
    g1 = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
    z_fused, lens = net.last_step["z"].clone(), net.last_step["lengths"].clone()
    net.zero_grad()
    l2 = net.step(batch, "train", masks=pm)
    l2.backward()
    assert float(l1) == float(l2)
    for n, p in net.named_parameters():
        if p.grad is None:
            continue
        if p is emb_w:  # the embedding scatter uses float atomics: equal to rounding, not bitwise
            assert rel_l2(p.grad, g1[n]) < 1e-5
        else:
            assert torch.equal(p.grad, g1[n]), n
    order = [batch[2], batch[3], batch[4], batch[0], batch[1]]
    with torch.no_grad():
        for g in range(5):
            zg = net.encoder(order[g], pm.emb_row_scale[g], pm.whh_mask[g])
            assert torch.equal(net.encoder.last_lengths[:, 0], lens[:, g])
            assert float((zg - z_fused[g]).abs().max()) == 0.0
    assert int(lens[1].max()) <= 1500 and int(lens[1].min()) > 900


_HEADLINE_REF = {}


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_headline_shape_step_vs_oracle(precision):
    """BASELINE.json's full configuration (batch 80 x 5 sequences, trunc_len 1500, E=64, 2 layers, `last`, rates 0.3) against the
    CPU oracle with the same 14 masks: T1/T_eff bit-exact, embeddings, the three losses and all 23 live gradients inside the gate.
    The oracle runs in fp32 here (6 s of host time at this size; its own distance to fp64 is <= 5e-6) and is shared by
    the two precision modes."""
    P = R.init_params(vocab=250, E=64, L=2, seed=0)
    batch = list(R.synthetic_batch(80, 1500, 250, seed=1234))
    masks = R.draw_step_masks(80, 250, 64, emb_droprate=0.3, rnn_droprate=0.3, do_rate=0.3, seed=5)
    kw = dict(L=2, bi="last", beta=2.0, use_projection=False, p_emb=0.3)
    got = run_product_step(P, batch, masks, p_rnn=0.3, p_do=0.3, precision=precision, **kw)
    if "ref" not in _HEADLINE_REF:
        _HEADLINE_REF["ref"] = run_oracle_step(P, batch, masks, dtype=torch.float32, **kw)
    ref = _HEADLINE_REF["ref"]
    assert int(ref["lengths"][1].min()) > 900  # the quirk-Q2 truncation is active (T_eff ~ 1100 of 1500)
    _check(got, ref, TOL[precision])


def test_quirks_interior_zeros_zero_weights_and_grad_sparsity():
    """Q1: an interior id 0 counts as a pad for T1 (a COUNT of non-zero ids, so the slice is shorter than the last non-zero
    position).  Q2 with exact zeros in the embedding matrix: T_eff counts non-zero VALUES per embedding column.  Q14: the embedding
    gradient is exactly zero on row 0, on rows dropped by the row mask and on ids that occur only in the truncated tail."""
    V, E, L, B, T = 61, 32, 2, 6, 48
    P = R.init_params(vocab=V, E=E, L=L, seed=7)
    P["emb"][5] = 0.0          # a whole row of exact zeros (besides the padding row)
    P["emb"][9, ::2] = 0.0     # and a row that is zero in every other column
    g = torch.Generator().manual_seed(8)
    batch = [torch.randint(1, V, (B, T), generator=g) for _ in range(5)] + [torch.randint(0, 2, (B,), generator=g)]
    for t in batch[:5]:
        t[2, :] = 5            # a sequence made only of the all-zero-embedding id
        t[:, 10:14] = 0        # interior zeros in every sequence
        t[1, 30:] = 0          # plus an ordinary padded tail
        t[:, 40:] = torch.where(t[:, 40:] != 0, torch.full_like(t[:, 40:], 17), t[:, 40:])  # id 17 lives in the tail
        t[:, :40][t[:, :40] == 17] = 18
    masks = R.draw_step_masks(B, V, E, emb_droprate=0.3, rnn_droprate=0.3, do_rate=0.3, seed=9)
    kw = dict(L=L, bi="last", beta=2.0, use_projection=False, p_emb=0.3)
    got = run_product_step(P, batch, masks, p_rnn=0.3, p_do=0.3, **kw)
    ref = run_oracle_step(P, batch, masks, **kw)
    assert int(ref["lengths"][0].max()) <= T - 4          # the interior zeros shortened T1
    assert int(ref["lengths"][1].max()) < int(ref["lengths"][0].min())  # and the row mask shortened it again
    _check(got, ref, 1e-4)
    zero_rows = (ref["grads"]["emb"].abs().sum(1) == 0)
    assert bool(zero_rows[0]) and bool(zero_rows[17])  # padding row; the id that occurs only beyond T_eff
    assert torch.equal(got["grads"]["emb"].abs().sum(1) == 0, zero_rows), "sparsity pattern of the embedding gradient"


def test_single_step_and_single_sequence():
    """Smallest shapes: one time step, one sequence per group."""
    for B, T in ((1, 1), (1, 9), (3, 1)):
        P = R.init_params(vocab=40, E=32, L=2, seed=3)
        batch = list(R.synthetic_batch(B, T, 40, seed=4))
        masks = R.draw_step_masks(B, 40, 32, emb_droprate=0.0, rnn_droprate=0.3, do_rate=0.3, seed=5)
        masks.emb_row_keep = None
        kw = dict(L=2, bi="mean", beta=2.0, use_projection=False, p_emb=0.0)
        got = run_product_step(P, batch, masks, p_rnn=0.3, p_do=0.3, **kw)
        ref = run_oracle_step(P, batch, masks, **kw)
        _check(got, ref, 1e-4)


def test_config4_full_size_proteome_and_all_pairs():
    """BASELINE.json config 4 at FULL size: 20 000 synthetic proteins x 1500 tokens embedded in batches of 512 (eval mode), then all
    200 010 000 pairs (i <= j) scored.  Rows of the embedding matrix and a random sample of the pair scores are compared with the
    CPU oracle; the triangle enumeration is checked at its corners and against explicit-index scoring."""
    M, T = 20000, 1500
    P = R.init_params(vocab=250, E=64, L=2, seed=0)
    net = build_product(P, L=2, bi="last").eval()
    x = torch.randint(1, 250, (M, T), generator=torch.Generator().manual_seed(4321))
    with torch.no_grad():
        z = net.embed(x.cuda(), 512)
        prob = net.score_pairs(z)
    assert z.shape == (M, 64) and prob.numel() == M * (M + 1) // 2
    rows = [0, 511, 512, 9999, M - 1]  # batch boundaries included
    with torch.no_grad():
        zr, _ = R.encoder_forward(x[rows], {k: v.double() for k, v in P.items()}, num_layers=2, bi_reduce="last", training=False)
    assert rel_l2(z[rows].cpu(), zr.float()) < 1e-4
    g = torch.Generator().manual_seed(1)
    ia = torch.randint(0, M, (2000,), generator=g)
    ib = torch.randint(0, M, (2000,), generator=g)
    ia, ib = torch.minimum(ia, ib), torch.maximum(ia, ib)
    ia[:3], ib[:3] = torch.tensor([0, 0, M - 1]), torch.tensor([0, M - 1, M - 1])  # corners of the triangle
    flat = ia * M - ia * (ia - 1) // 2 + (ib - ia)  # row-major index of (i, j), i <= j
    zc = z.cpu()
    with torch.no_grad():
        ref = torch.sigmoid(R.mlp_head(zc[ia], zc[ib], P).squeeze(1))
    got = prob[flat.cuda()].cpu()
    assert float((got - ref).abs().max()) < 1e-5
    assert torch.equal(net.score_pairs(z, ia.cuda(), ib.cuda()).cpu(), got)


def test_pair_score_ranges_tile_the_triangle():
    """ib200_pair_score_range: the blocks three ranks would score (parallel.pair_block) concatenate to the full triangle, bit for bit;
    sharded_proteome_scores on one process equals embed + score_pairs."""
    from intrepppid_b200.parallel import pair_block, sharded_proteome_scores

    P = R.init_params(E=64, L=2, seed=5)
    net = build_product(P, L=2, bi="last").eval()
    M = 301
    z = torch.randn(M, 64).cuda()
    full = net.score_pairs(z)
    parts = [net.score_pairs_range(z, *pair_block(M, r, 3)[2:]) for r in range(3)]
    assert torch.equal(torch.cat(parts), full)
    assert net.score_pairs_range(z, 17, 0).numel() == 0
    with pytest.raises(RuntimeError):
        net.score_pairs_range(z, M * (M + 1) // 2 - 3, 4)  # past the end of the triangle
    x = torch.randint(1, 250, (40, 64))
    z_all, p0, probs = sharded_proteome_scores(net, x, batch_size=16)
    with torch.no_grad():
        assert torch.equal(z_all, net.embed(x.cuda(), 16)) and p0 == 0
        assert torch.equal(probs, net.score_pairs(z_all))


def test_out_of_range_token_ids_raise_like_f_embedding():
    P = R.init_params(vocab=40, E=32, L=1)
    net = build_product(P, L=1, bi="last").eval()
    x = torch.randint(1, 40, (3, 12))
    from intrepppid_b200 import ops

    net.encoder.check_lengths = "sync"
    x[1, 5] = 40
    with pytest.raises(IndexError):
        net.encoder(x.cuda())
    x[1, 5] = -1
    with pytest.raises(IndexError):
        net.encoder(x.cuda())
    for dt in (torch.int32, torch.int16):  # the device-side flag sees narrowed ids too
        with pytest.raises(IndexError):
            net.encoder(x.to(dt).cuda())
    net.encoder.check_lengths = True  # default: recorded on the device, raised lazily (no host sync inside the call)
    with torch.no_grad():
        net.encoder(x.cuda())
    with pytest.raises(IndexError):
        ops.check_pending(sync=True)
    x[1, 5] = 39
    with torch.no_grad():
        net.encoder(x.cuda())
    ops.check_pending(sync=True)  # in-range ids: nothing recorded
    x[1, 5] = -1
    net.encoder.check_lengths = False  # no check at all: the ids are clamped, nothing faults
    with torch.no_grad():
        z = net.encoder(x.cuda())
    assert bool(torch.isfinite(z).all())


def test_variational_dropout_row_mask_also_in_eval():
    """Q8: variational_dropout=True draws a ROW mask [4H,1] for weight_hh_l0 and -- as the reference does (`training=True` at
    utils/weightdrop.py:94) -- applies it in eval mode too; DropConnect (the default) is the identity in eval."""
    import intrepppid_b200 as ib

    V, E, L = 50, 32, 2
    P = R.init_params(vocab=V, E=E, L=L, seed=11)
    x = torch.randint(1, V, (5, 20), generator=torch.Generator().manual_seed(12))
    row = (torch.rand(4 * E, 1, generator=torch.Generator().manual_seed(13)) >= 0.3).float() / 0.7
    net = build_product(P, L=L, bi="last").eval()
    with torch.no_grad():
        z = net.encoder(x.cuda(), None, row.expand(4 * E, E).contiguous().cuda()).cpu()
        zr, _ = R.encoder_forward(x, P, num_layers=L, bi_reduce="last", training=False, whh_mask=row.expand(4 * E, E))
        assert rel_l2(z, zr) < 1e-4
        assert torch.equal(net.encoder(x.cuda()), net.encoder(x.cuda()))  # DropConnect: no mask in eval, deterministic
    torch.manual_seed(0)
    vnet = ib.intrepppid_network(1, vocab_size=V, embedding_size=E, rnn_num_layers=L, variational_dropout=True).cuda().eval()
    m = vnet.encoder.encoder.rnn_dp.sample_mask("weight_hh_l0", 3)
    assert m.shape == (3, 4 * E, E) and bool((m == m[:, :, :1]).all())  # one value per row, per group
    with torch.no_grad():
        assert not torch.equal(vnet.encoder(x.cuda()), vnet.encoder(x.cuda()))  # a fresh row mask per call, even in eval


def test_long_sequences_t4000():
    """trunc_len 4000 (the stress configuration's length) on the register-resident kernels: token staging, ring prefetch and the
    error growth over 4000 recurrent steps stay inside the fp32 gate."""
    V, E, L, B, T = 250, 64, 2, 6, 4000
    P = R.init_params(vocab=V, E=E, L=L, seed=2)
    batch = list(R.synthetic_batch(B, T, V, seed=3, padded=True))
    masks = R.draw_step_masks(B, V, E, emb_droprate=0.3, rnn_droprate=0.3, do_rate=0.3, seed=4)
    kw = dict(L=L, bi="mean", beta=2.0, use_projection=False, p_emb=0.3)
    got = run_product_step(P, batch, masks, p_rnn=0.3, p_do=0.3, **kw)
    ref = run_oracle_step(P, batch, masks, **kw)
    assert int(ref["lengths"][1].max()) > 2500
    _check(got, ref, 1e-4)


def test_vocabulary_larger_than_the_one_hot_tile():
    """vocab_size > 256 does not fit the one-hot tile of the fused layer-0 gradient kernel: the general path (gathered dW GEMM,
    dX_0 GEMM, scatter) must take over transparently."""
    V, E, L, B, T = 300, 64, 2, 7, 90
    P = R.init_params(vocab=V, E=E, L=L, seed=17)
    batch = list(R.synthetic_batch(B, T, V, seed=18, padded=True))
    masks = R.draw_step_masks(B, V, E, emb_droprate=0.3, rnn_droprate=0.3, do_rate=0.3, seed=19)
    kw = dict(L=L, bi="last", beta=2.0, use_projection=False, p_emb=0.3)
    got = run_product_step(P, batch, masks, p_rnn=0.3, p_do=0.3, **kw)
    ref = run_oracle_step(P, batch, masks, **kw)
    assert int(torch.stack(batch[:5]).max()) > 256
    _check(got, ref, 1e-4)
