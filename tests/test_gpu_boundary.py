"""GPU tests of the drop-in boundary itself: the C++ operator library under torch.library.opcheck, the backward cut at
the layer boundary, the autograd contracts of the differentiable ops, and vocabulary sizes up to the documented bound."""
import pytest
import torch

from conftest import rel_l2
from helpers import build_product, product_masks
from oracle import restatement as R

pytestmark = pytest.mark.gpu


def _lstm_tensors(net):
    return [p.detach() for p in net.encoder.encoder.rnn.ordered()]


def test_opcheck_schema_and_fake_tensor_of_every_op():
    """torch.library.opcheck: the schemas (mutation / aliasing annotations) are truthful and the Meta kernels return exactly the
    shapes, dtypes and devices the CUDA kernels return."""
    ops = torch.ops.intrepppid_b200
    checks = ("test_schema", "test_faketensor")
    P = R.init_params(vocab=60, E=32, L=2, seed=3)
    net = build_product(P, L=2, bi="max").train()
    lstm = _lstm_tensors(net)
    emb = net.encoder.embedder.weight.detach()
    G, B, T, V, H, L = 2, 5, 24, 60, 32, 2
    tok = torch.randint(1, V, (G, B, T), device="cuda")
    ers = (torch.rand(G, V, device="cuda") > 0.3).float() / 0.7
    whm = (torch.rand(G, 4 * H, H, device="cuda") > 0.3).float() / 0.7
    enc_args = (tok, emb, lstm, ers, whm, L, 2, 0, True)
    torch.library.opcheck(ops.encoder_fwd.default, enc_args, test_utils=checks)
    hn, status, ws = ops.encoder_fwd(*enc_args)
    assert status.dtype == torch.int32 and status.shape == (3, G) and int(status[2].max()) == 0
    d_hn = torch.randn_like(hn)
    # the backward consumes its workspace (dgates overwrite the saved gates): every opcheck run gets a fresh one
    torch.library.opcheck(ops.encoder_bwd.default, (ops.encoder_fwd(*enc_args)[2], d_hn, emb, lstm, ers, whm, G, B, T, L, 2, 0),
                          test_utils=("test_faketensor",))
    fc_w, fc_b = net.encoder.encoder.fc.weight.detach(), net.encoder.encoder.fc.bias.detach()
    for mode in (0, 1, 2):
        torch.library.opcheck(ops.pool_fc_fwd.default, (hn, fc_w, fc_b, mode), test_utils=checks)
    z, pooled, argmax = ops.pool_fc_fwd(hn, fc_w, fc_b, 2)
    torch.library.opcheck(ops.pool_fc_bwd.default, (torch.randn_like(z), pooled, argmax, fc_w, 2), test_utils=checks)
    torch.library.opcheck(ops.pool_fc_bwd.default, (torch.randn_like(z), pooled, None, fc_w, 1), test_utils=checks)
    head = [t.detach() for t in net.head.tensors()]
    z5 = torch.randn(5, B, H, device="cuda")
    y = torch.randint(0, 2, (B,), device="cuda")
    masks = [None, (torch.rand(B, H // 2, device="cuda") > 0.3).float() / 0.7, None, None]
    torch.library.opcheck(ops.loss_head_fwd.default, (z5, y, head, masks, 2.0), test_utils=checks)
    torch.library.opcheck(ops.loss_head_bwd.default, (z5, y, head, masks, 2.0, torch.ones(1, device="cuda"), None), test_utils=checks)
    proj = [torch.randn(H, H, device="cuda"), torch.randn(H, device="cuda")]
    torch.library.opcheck(ops.loss_head_bwd.default, (z5, y, head + proj, masks, 2.0, torch.ones(1, device="cuda"),
                                                      torch.randn(B, device="cuda")), test_utils=checks)
    zz = torch.randn(9, H, device="cuda")
    torch.library.opcheck(ops.pair_score.default, (zz, *head, None, None), test_utils=checks)
    ia = torch.randint(0, 9, (13,), device="cuda", dtype=torch.int32)
    torch.library.opcheck(ops.pair_score.default, (zz, *head, ia, ia.flip(0).contiguous()), test_utils=checks)
    torch.library.opcheck(ops.pair_score_range.default, (zz, *head, 4, 20), test_utils=checks)
    torch.library.opcheck(ops.batch_metrics.default, (torch.randn(B, device="cuda"), y, 0.5), test_utils=checks)


def test_shim_translates_c_abi_status_codes():
    """A non-zero status of the C ABI surfaces as IB200Error with the library's own message."""
    from intrepppid_b200 import ops
    from intrepppid_b200._lib import IB200Error

    z = torch.randn(6, 64, device="cuda")
    P = R.init_params(E=64, L=1)
    net = build_product(P, L=1, bi="last").eval()
    with pytest.raises(IB200Error, match="triangle"):
        ops.pair_score_range(z, *net.head.tensors(), 20, 5)  # 6*7/2 = 21 pairs only


def test_backward_cut_at_the_layer_boundary_is_bit_identical():
    """ib200_encoder_bwd_layers(L-1..1) + (0..0) == ib200_encoder_bwd, and the hook sees the final upper-layer gradients."""
    from intrepppid_b200 import ops

    P = R.init_params(E=64, L=3, seed=11)
    batch = list(R.synthetic_batch(6, 70, 250, seed=4, padded=True))
    masks = R.draw_step_masks(6, 250, 64, emb_droprate=0.3, rnn_droprate=0.3, do_rate=0.3, seed=9)

    def run(hooked):
        net = build_product(P, L=3, bi="mean").train()
        seen = []
        hook = lambda upper: seen.append(upper.clone())  # noqa: E731
        if hooked:
            ops.EARLY_GRAD_HOOKS.append(hook)
        try:
            loss = net.step([t.cuda() for t in batch], "train", masks=product_masks(masks, 0.3))
            loss.backward()
        finally:
            if hooked:
                ops.EARLY_GRAD_HOOKS.remove(hook)
        return {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}, seen

    whole, _ = run(False)
    cut, seen = run(True)
    assert whole.keys() == cut.keys()
    for n in whole:
        assert torch.equal(whole[n], cut[n]), n
    assert len(seen) == 1
    order = [n for n in ops.lstm_param_order(3) if "_l0" not in n]
    flat = torch.cat([cut["encoder.encoder.rnn." + n].reshape(-1) for n in order])
    assert torch.equal(seen[0], flat)  # what the hook saw mid-backward is what .grad holds at the end


def test_autograd_contracts_of_the_differentiable_ops():
    from intrepppid_b200 import ops

    P = R.init_params(E=32, L=2, seed=2)
    net = build_product(P, L=2, bi="last").train()
    batch = [t.cuda() for t in R.synthetic_batch(4, 20, 250, seed=8, padded=False)]
    z = net.encoder.forward_groups(torch.stack(batch[:5]))
    (loss, cl, tl), y_hat = ops.loss_head(2.0, z, batch[5], *net.head.tensors())
    assert loss.requires_grad and y_hat.requires_grad
    assert not cl.requires_grad and not tl.requires_grad  # logged detached by the reference: no silent gradient drop possible
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="consumed"):
        loss.backward()  # the BPTT kernels overwrote the saved gates: a second pass must not return garbage


@pytest.mark.parametrize("V", [8192, 28672])
def test_large_vocabulary_lengths_and_embeddings(V):
    """The lengths kernel keeps a [V] histogram + row list per sequence in shared memory (opt-in above 48 KB): the whole documented
    range 2..28672 runs and matches the oracle (advisor finding: V > 6144 used to fail at launch)."""
    P = R.init_params(vocab=V, E=32, L=1, seed=1)
    net = build_product(P, L=1, bi="last", p_emb=0.3).train()
    g = torch.Generator().manual_seed(V)
    x = torch.randint(1, V, (3, 6, 50), generator=g)
    x[0, 2, 30:] = 0
    x[1, :, 44:] = 0
    keep = (torch.rand(3, V, generator=g) > 0.3).float()
    with torch.no_grad():
        z = net.encoder.forward_groups(x.cuda(), (keep / 0.7).cuda(), None, draw=False).cpu()
    lens = net.encoder.last_lengths.cpu()
    for gi in range(3):
        zr, info = R.encoder_forward(x[gi], {k: v.double() for k, v in P.items()}, num_layers=1, bi_reduce="last", training=True,
                                     emb_droprate=0.3, row_keep=keep[gi].double(), whh_mask=None)
        assert (int(lens[0, gi]), int(lens[1, gi])) == (info.T1, info.T_eff)
        assert rel_l2(z[gi], zr) < 1e-4
    from intrepppid_b200 import ops

    ops.check_pending(sync=True)


def test_p2p_allreduce_world_of_one_is_the_identity():
    """ib200_p2p_alloc / ib200_p2p_allreduce_mean / ib200_p2p_free on one GPU: with world = 1 the staged copy, the flag handshake with
    myself and the mean over one rank must give the bucket back unchanged, for both parity halves, unaligned views and odd sizes
    (the multi-GPU equivalence with NCCL is tests/p2p_check.py under torchrun)."""
    import ctypes as C

    from intrepppid_b200 import _lib

    L = _lib.lib()
    stage_floats, world = 4096, 1
    total = 2 * stage_floats * 4 + 256
    base, handle = C.c_void_p(), C.create_string_buffer(64)
    _lib.check(L.ib200_p2p_alloc(total, C.byref(base), handle), "ib200_p2p_alloc")
    try:
        assert any(handle.raw), "an IPC handle was written"
        arr = C.c_void_p * 1
        stage, flags = arr(base.value), arr(base.value + 2 * stage_floats * 4)
        st = torch.cuda.current_stream().cuda_stream
        for epoch, n, off in ((1, 4096, 0), (2, 1001, 1), (3, 7, 3), (4, 1, 0)):
            x = torch.randn(n + 4, device="cuda")[off:off + n]
            ref = x.clone()
            _lib.check(L.ib200_p2p_allreduce_mean(world, 0, stage, flags, stage_floats, x.data_ptr(), n, epoch, st), "p2p")
            torch.cuda.synchronize()
            assert torch.equal(x, ref), (epoch, n)
        # argument validation
        assert L.ib200_p2p_allreduce_mean(world, 0, stage, flags, stage_floats, 0, 8, 5, st) < 0          # null data
        assert L.ib200_p2p_allreduce_mean(world, 0, stage, flags, 16, base.value, 17, 5, st) < 0          # bucket > staging half
        assert L.ib200_p2p_allreduce_mean(world, 0, stage, flags, stage_floats, base.value, 8, 0, st) < 0  # epochs count from 1
        assert L.ib200_p2p_allreduce_mean(9, 0, stage, flags, stage_floats, base.value, 8, 5, st) < 0      # world > 8
    finally:
        _lib.check(L.ib200_p2p_free(base), "ib200_p2p_free")
