"""Side-by-side: the restatement against the UNMODIFIED reference files, executed here through oracle/ref_shim.py.
Container only (the GPU box has no /root/reference)."""
import pytest
import torch

from conftest import assert_grad_close, rel_l2
from oracle import ref_shim, restatement as R

pytestmark = pytest.mark.reference


def _both(bi, proj, L, E=32, B=5, T=30, V=60, beta=3.0, training=True, seed=3, p=0.3):
    torch.manual_seed(seed)
    net = ref_shim.build_reference_net(vocab=V, E=E, L=L, bi_reduce=bi, use_projection=proj, beta=beta,
                                       emb_droprate=p, rnn_droprate=p, do_rate=p)
    batch = list(R.synthetic_batch(B, T, V, seed=seed))
    for s in batch[:5]:
        s[1, T // 2:] = 0
        s[2, :] = 0
        s[3, 4] = 0
    m = R.draw_step_masks(B, V, E, emb_droprate=p, rnn_droprate=p, do_rate=p, seed=seed + 1)
    net.train(training)
    rows = [m.emb_row_keep[g].reshape(-1, 1) for g in range(5)]
    drops = [m.whh_mask[g] for g in range(5)] + [m.fc1_w, m.do1, m.do2, m.fc2_w]
    with ref_shim.injected_masks(rows, drops):
        loss = net.step(batch, "train" if training else "val")
    P = {k: v.detach().clone().requires_grad_(True) for k, v in R.params_from_state_dict(net.state_dict(), L).items()}
    out = R.step(batch, P, num_layers=L, bi_reduce=bi, beta_classifier=beta, training=training, emb_droprate=p,
                 use_projection=proj, masks=m if training else None)
    return net, loss, P, out


@pytest.mark.parametrize("bi,proj,L", [("last", False, 2), ("mean", True, 2), ("max", False, 3), ("last", True, 1)])
def test_step_and_grads_match_reference(bi, proj, L):
    net, loss, P, out = _both(bi, proj, L)
    assert abs(float(loss) - float(out.loss)) < 1e-6
    loss.backward()
    out.loss.backward()
    named = dict(net.named_parameters())
    for n, p in P.items():
        key = {"emb": "encoder.embedder.weight", "fc_w": "encoder.encoder.fc.weight", "fc_b": "encoder.encoder.fc.bias",
               "fc1_w": "head.classify.fc1.module.weight_raw", "fc1_b": "head.classify.fc1.module.bias",
               "fc2_w": "head.classify.fc2.module.weight_raw", "fc2_b": "head.classify.fc2.module.bias",
               "proj_w": "triplet_projection.1.weight", "proj_b": "triplet_projection.1.bias"}.get(
            n, "encoder.encoder.rnn." + (n + "_raw" if n == "weight_hh_l0" else n))
        assert_grad_close(p.grad, named[key].grad, 1e-4, n)


def test_eval_step_matches_reference():
    net, loss, P, out = _both("last", False, 2, training=False)
    assert abs(float(loss) - float(out.loss)) < 1e-6


def test_q14_embedding_grad_sparsity_pattern():
    """Row 0, rows dropped by the row mask and ids only seen in the truncated tail get exactly-zero gradient."""
    net, loss, P, out = _both("last", False, 2)
    loss.backward()
    out.loss.backward()
    g_ref = dict(net.named_parameters())["encoder.embedder.weight"].grad
    assert torch.equal(g_ref.abs().sum(1) == 0, P["emb"].grad.abs().sum(1) == 0)
    assert float(g_ref[0].abs().max()) == 0


def test_q15_layer0_table_identity_is_exact():
    """F.embedding(x, m*W) @ W_ih.T == ((m*W) @ W_ih.T)[x]: the V x 4H lookup-table form of the layer-0 projection."""
    P = R.init_params(E=64)
    x = torch.randint(0, 250, (4, 50))
    a = torch.nn.functional.embedding(x, P["emb"]) @ P["weight_ih_l0"].T
    b = (P["emb"] @ P["weight_ih_l0"].T)[x]
    assert torch.equal(a, b)


def test_reference_parameter_census():
    torch.manual_seed(0)
    net = ref_shim.build_reference_net()
    assert sum(p.numel() for p in net.parameters()) == 216498
    assert len(net.state_dict()) == 45


def test_infer_from_csv_row_loop_matches_reference_batch_of_one():
    """cli/infer.py:196-225 calls `net(embed_a.unsqueeze(0), embed_b.unsqueeze(0))` per row: the restatement of that loop (the oracle
    of intrepppid_b200.infer) against the unmodified TripletE2ENet.forward + sigmoid, including a 1-token protein, an interior
    <unk> and a row with an unknown id."""
    torch.manual_seed(4)
    V, E, L, T = 40, 32, 2, 24
    net = ref_shim.build_reference_net(vocab=V, E=E, L=L, bi_reduce="last", use_projection=True).eval()
    g = torch.Generator().manual_seed(5)
    toks = {}
    for i, n in enumerate((T, 1, 7, 13, 24, 9)):
        row = torch.zeros(T, dtype=torch.long)
        row[:n] = torch.randint(1, V, (n,), generator=g)
        toks[f"P{i}"] = row
    toks["P3"][2] = 0
    rows = [("a", "P0", "P1"), ("b", "P2", "P3"), ("c", "P4", "NOPE"), ("d", "P5", "P5"), ("e", "P1", "P3")]
    P = R.params_from_state_dict(net.state_dict(), L)
    got = R.infer_from_csv_rows(toks, rows, P, num_layers=L, bi_reduce="last")
    assert [i for i, _ in got] == ["a", "b", "d", "e"]
    with torch.no_grad():
        for (itx, p), (_, a, b) in zip(got, [r for r in rows if r[2] in toks]):
            ref = float(torch.sigmoid(net(toks[a].unsqueeze(0), toks[b].unsqueeze(0)))[0, 0])
            assert abs(p - ref) < 1e-6, itx
        # and the batch-of-one result differs from a mixed-length batch (no packing): the reason the buckets exist
        mixed = torch.sigmoid(net(torch.stack([toks["P1"], toks["P0"]]), torch.stack([toks["P3"], toks["P0"]])))[0, 0]
        assert abs(float(mixed) - dict(got)["e"]) > 1e-6
