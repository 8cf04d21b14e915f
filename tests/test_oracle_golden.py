"""The CPU restatement (oracle/restatement.py) against golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py).  Runs everywhere (no GPU, no /root/reference needed)."""
import pytest
import torch

from conftest import GOLDEN_CASES, assert_grad_close, load_golden, rel_l2
from oracle import restatement as R


def _run(gold, impl, dtype=torch.float32):
    c = gold["config"]
    P = {k: v.to(dtype).clone().requires_grad_(True) for k, v in gold["params"].items()}
    mk = gold["masks"]
    masks = R.StepMasks(mk["emb_row_keep"], mk["whh_mask"].to(dtype), mk["fc1_w"].to(dtype), mk["do1"].to(dtype),
                        mk["do2"].to(dtype), mk["fc2_w"].to(dtype))
    # golden tokens are stored in dataset order (p1, p2, anchor, pos, neg)
    batch = gold["tokens"] + [gold["y"]]
    out = R.step(batch, P, num_layers=c["L"], bi_reduce=c["bi"], beta_classifier=c["beta"], training=True,
                 emb_droprate=c["p_emb"], use_projection=c["proj"], masks=masks, impl=impl)
    out.loss.backward()
    return out, P


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("impl", ["vf", "manual"])
def test_training_step_matches_reference_golden(name, impl):
    gold = load_golden(name)
    out, P = _run(gold, impl)
    t = gold["train"]
    assert abs(float(out.loss) - float(t["loss"])) < 2e-6
    assert abs(float(out.classifier_loss) - float(t["classifier_loss"])) < 2e-6
    assert abs(float(out.triplet_loss) - float(t["triplet_loss"])) < 2e-6
    assert rel_l2(out.y_hat, t["y_hat"]) < 1e-5
    for g in range(5):
        assert rel_l2(out.z[g], t["z"][g]) < 1e-5
    for n, gref in t["grads"].items():
        assert gref is not None, n
        assert_grad_close(P[n].grad, gref, 2e-4, n)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_fp64_restatement_is_the_tighter_gold(name):
    """fp64 run of the restatement agrees with the reference's fp32 numbers to fp32 noise."""
    gold = load_golden(name)
    out, P = _run(gold, "vf", torch.float64)
    assert abs(float(out.loss) - float(gold["train"]["loss"])) < 1e-6
    for n, g in gold["train"]["grads"].items():
        assert_grad_close(P[n].grad.float(), g, 1e-4, n)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_eval_forward_matches_reference_golden(name):
    gold = load_golden(name)
    c = gold["config"]
    with torch.no_grad():
        logits = R.infer_logits(gold["tokens"][0], gold["tokens"][1], gold["params"], num_layers=c["L"], bi_reduce=c["bi"])
    assert logits.shape == gold["eval"]["logits"].shape
    assert rel_l2(logits, gold["eval"]["logits"]) < 1e-5


def test_truncation_integers_are_exact():
    """Q1/Q2: T1 is a COUNT of non-zero ids; T_eff counts kept (row-mask != 0) non-pad tokens among the first T1."""
    gold = load_golden("train_last_E64")
    c = gold["config"]
    for g, tok in enumerate([gold["tokens"][2], gold["tokens"][3], gold["tokens"][4], gold["tokens"][0], gold["tokens"][1]]):
        T1 = R.first_truncation(tok)
        assert T1 == int((tok != 0).sum(1).max())
        keep = gold["masks"]["emb_row_keep"][g]
        table = R.masked_embedding_table(gold["params"]["emb"], keep, c["p_emb"], True)
        x = torch.nn.functional.embedding(tok[:, :T1], table)
        T_eff = R.second_truncation(x)
        manual = int(((tok[:, :T1] != 0) & (keep[tok[:, :T1]] != 0)).sum(1).max())
        assert T_eff == manual <= T1


def test_dead_chain_under_last_has_exactly_zero_grads():
    gold = load_golden("train_last_E64")
    for n in ("weight_ih_l1", "weight_hh_l1", "bias_ih_l1", "bias_hh_l1"):
        assert float(gold["train"]["grads"][n].abs().max()) == 0.0
    assert float(gold["train"]["grads"]["weight_ih_l1_reverse"].abs().max()) > 0


def test_concat_is_rejected_like_the_reference():
    with pytest.raises(ValueError):
        R.bi_reduce_hn(torch.zeros(2, 3, 4), "concat")


def test_all_pad_batch_raises():
    P = R.init_params(E=32)
    with pytest.raises(RuntimeError):
        R.encoder_forward(torch.zeros(3, 10, dtype=torch.long), P, num_layers=2, bi_reduce="last", training=False)


def test_state_dict_contract_of_reference_recorded():
    gold = load_golden("train_last_E64")
    keys = gold["state_dict_keys"]
    assert len(keys) == 45
    assert "encoder.encoder.rnn.weight_hh_l0_raw" in keys and "encoder.encoder.rnn.weight_hh_l0" not in keys
    assert "encoder.encoder.rnn_dp.module.weight_hh_l0_raw" in keys
    assert all(k.startswith("encoder.projection.") for k in gold["dead_parameter_keys"])


def test_infer_row_loop_matches_reference_golden():
    """`infer from_csv` at batch 1 (cli/infer.py:196-225): probabilities and per-protein embeddings recorded from the reference."""
    gold = load_golden("infer_rows")
    c = gold["config"]
    got = R.infer_from_csv_rows(gold["tokens"], gold["rows"], gold["params"], num_layers=c["L"], bi_reduce=c["bi"])
    assert [i for i, _ in got] == [i for i, _ in gold["scored"]] and len(got) == len(gold["rows"]) - 1
    assert max(abs(a - b) for (_, a), (_, b) in zip(got, gold["scored"])) < 1e-6
    for name, z in gold["z"].items():
        z1, _ = R.encoder_forward(gold["tokens"][name].unsqueeze(0), gold["params"], num_layers=c["L"], bi_reduce=c["bi"], training=False)
        assert rel_l2(z1[0], z) < 1e-5, name
