"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads and exports every symbol include/ib200.h declares,
host-only entry points behave, and the Python mirror keeps the reference's names, signatures and checkpoint keys.
No compute entry point is called here (no GPU in this tier)."""
import ctypes
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from intrepppid_b200 import build as b

    b.build()
    from intrepppid_b200 import _lib

    return _lib.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ib200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ib200_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ib200.h but not exported by libib200.so"
    from intrepppid_b200 import _lib

    assert sorted(_lib.EXPORTS) == names


def test_version_and_workspace_are_host_only(lib):
    from intrepppid_b200._lib import Cfg

    assert lib.ib200_version() == 100
    train = Cfg(5, 80, 1500, 250, 64, 2, 0, 0, 1, 0)
    infer = Cfg(1, 512, 1500, 250, 64, 2, 0, 0, 0, 0)
    wt, wi = lib.ib200_workspace_bytes(train), lib.ib200_workspace_bytes(infer)
    # training keeps gates/c/h for 3 live chains: a few GB at the headline shape, far below 180 GB
    assert 3e9 < wt < 8e9 and 1e9 < wi < 3e9
    assert lib.ib200_workspace_bytes(Cfg(5, 80, 1500, 250, 48, 2, 0, 0, 1, 0)) == 0      # unsupported H
    assert lib.ib200_workspace_bytes(Cfg(5, 80, 1500, 250, 64, 9, 0, 0, 1, 0)) == 0      # too many layers
    assert lib.ib200_workspace_bytes(Cfg(1, 8, 64, 28672, 64, 1, 0, 0, 0, 0)) > 0         # largest vocabulary the lengths kernel holds
    assert lib.ib200_workspace_bytes(Cfg(1, 8, 64, 28673, 64, 1, 0, 0, 0, 0)) == 0
    assert lib.ib200_workspace_bytes(Cfg(1, 8, 64, 40000, 256, 1, 0, 0, 0, 0)) == 0       # ... for every hidden size
    assert lib.ib200_workspace_bytes(Cfg(5, 80, 1500, 250, 64, 2, 3, 0, 1, 0)) == 0      # "concat" is not a mode


def test_null_arguments_are_rejected_before_any_launch(lib):
    from intrepppid_b200._lib import Cfg

    cfg = Cfg(1, 2, 8, 10, 32, 1, 0, 0, 0, 0)
    st = lib.ib200_encoder_fwd(cfg, None, None, None, None, None, None, None, 0, None)
    assert st == -1 and b"null" in lib.ib200_last_error()
    assert lib.ib200_pair_score(4, 48, None, None, None, 10, None, None, None) < 0


def test_reference_signatures_are_mirrored():
    import intrepppid_b200 as ib
    from intrepppid_b200.classifier.head import MLPHead
    from intrepppid_b200.encoders import AWDLSTMEncoder
    from intrepppid_b200.e2e.e2e_triplet import TripletE2ENet

    sig = inspect.signature(ib.intrepppid_network)
    want = ["steps_per_epoch", "vocab_size", "embedding_size", "rnn_num_layers", "rnn_dropout_rate", "variational_dropout",
            "bi_reduce", "embedding_droprate", "num_epochs", "do_rate", "beta_classifier", "lr", "use_projection", "optimizer_type"]
    assert list(sig.parameters)[:len(want)] == want
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["vocab_size"], d["embedding_size"], d["rnn_num_layers"], d["bi_reduce"], d["beta_classifier"]) == (250, 64, 2, "last", 2)
    assert list(inspect.signature(AWDLSTMEncoder.__init__).parameters)[1:] == [
        "embedder", "embedding_size", "embedding_droprate", "rnn_num_layers", "rnn_dropout_rate", "variational_dropout", "bi_reduce"]
    assert list(inspect.signature(MLPHead.__init__).parameters)[1:] == ["embedding_size", "do_rate"]
    assert list(inspect.signature(TripletE2ENet.__init__).parameters)[1:] == [
        "embedding_size", "encoder", "head", "embedding_droprate", "num_epochs", "steps_per_epoch", "beta_classifier",
        "use_projection", "optimizer_type", "lr"]


def test_checkpoint_keys_match_the_reference_recorded_in_golden():
    import intrepppid_b200 as ib
    from conftest import load_golden

    torch.manual_seed(0)
    net = ib.intrepppid_network(1)
    assert sorted(net.state_dict().keys()) == load_golden("train_last_E64")["state_dict_keys"]
    assert sum(p.numel() for p in net.parameters()) == 216498
    torch.manual_seed(0)
    net_p = ib.intrepppid_network(0, use_projection=True)   # cli/infer.py:170 builds it this way
    assert "triplet_projection.1.weight" in net_p.state_dict()


@pytest.mark.reference
def test_same_seed_same_initial_weights_and_cross_loading():
    import intrepppid_b200 as ib
    from oracle import ref_shim

    torch.manual_seed(0)
    mine = ib.intrepppid_network(1)
    torch.manual_seed(0)
    ref = ref_shim.build_reference_net()
    a, b = mine.state_dict(), ref.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(torch.equal(a[k], b[k]) for k in b)
    ref.load_state_dict(a, strict=True)
    mine.load_state_dict(b, strict=True)


def test_product_refuses_to_run_without_cuda():
    """The compute path fails loudly instead of falling back (this container has no GPU)."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import intrepppid_b200 as ib
    from intrepppid_b200._lib import IB200Error

    net = ib.intrepppid_network(1).eval()
    with pytest.raises(IB200Error):
        net.encoder(torch.ones(2, 8, dtype=torch.long))


def test_product_never_imports_the_oracle():
    import subprocess
    import sys

    code = "import sys; import intrepppid_b200, intrepppid_b200.ops; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "intrepppid_b200")):
        for f in files:
            if f.endswith(".py"):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().replace("oracle/", ""), f


OP_NAMES = ("encoder_fwd", "encoder_bwd", "encoder_bwd_layers", "pool_fc_fwd", "pool_fc_bwd", "loss_head_fwd", "loss_head_bwd",
            "pair_score", "pair_score_range", "batch_metrics")


def test_operator_library_is_the_cpp_shim():
    """The ops come from the C++ TORCH_LIBRARY shim (csrc/torch_ops.cpp -> libib200_torch.so), not from Python registrations."""
    import intrepppid_b200  # noqa: F401
    from intrepppid_b200 import _lib

    assert os.path.exists(os.path.join(_lib.PKG, "libib200_torch.so"))
    with open("/proc/self/maps") as fh:
        assert "libib200_torch.so" in fh.read()
    src = open(os.path.join(ROOT, "intrepppid_b200", "ops.py")).read()
    assert "torch.library.Library(" not in src and ".impl(" not in src


def test_meta_kernels_infer_shapes_without_a_gpu():
    """Shape functions (Meta dispatch key): what FakeTensor tracing / torch.compile of the surrounding Lightning step needs."""
    import intrepppid_b200  # noqa: F401
    from intrepppid_b200 import _lib
    from intrepppid_b200.ops import lstm_param_order

    ops = torch.ops.intrepppid_b200
    G, B, T, V, H, L = 5, 6, 40, 250, 64, 2
    m = lambda *shape, dtype=torch.float32: torch.empty(*shape, dtype=dtype, device="meta")  # noqa: E731
    lstm = []
    for l in range(L):
        for _ in range(2):
            lstm += [m(4 * H, H if l == 0 else 2 * H), m(4 * H, H), m(4 * H), m(4 * H)]
    assert len(lstm) == len(lstm_param_order(L))
    hn, status, ws = ops.encoder_fwd(m(G, B, T, dtype=torch.int64), m(V, H), lstm, m(G, V), m(G, 4 * H, H), L, 0, 0, True)
    assert hn.shape == (2, G * B, H) and status.shape == (3, G) and status.dtype == torch.int32
    assert ws.dtype == torch.uint8 and ws.numel() == _lib.lib().ib200_workspace_bytes(_lib.Cfg(G, B, T, V, H, L, 0, 0, 1, 0))
    n_flat = V * H + sum(t.numel() for t in lstm)
    assert ops.encoder_bwd(ws, m(2, G * B, H), m(V, H), lstm, None, None, G, B, T, L, 0, 0).shape == (n_flat,)
    flat = m(n_flat)
    assert ops.encoder_bwd_layers(ws, m(2, G * B, H), m(V, H), lstm, None, None, G, B, T, L, 0, 0, flat, 1, 1) is None
    # error paths of the shim (every message that carries numbers): unsupported hidden size / bi_reduce, wrong tensor count / shape
    with pytest.raises(RuntimeError, match="H=48"):
        ops.encoder_fwd(m(G, B, T, dtype=torch.int64), m(V, 48), lstm, None, None, L, 0, 0, False)
    with pytest.raises(RuntimeError, match="bi_reduce=3"):
        ops.encoder_fwd(m(G, B, T, dtype=torch.int64), m(V, H), lstm, None, None, L, 3, 0, False)
    with pytest.raises(RuntimeError, match="emb_row_scale must be"):
        ops.encoder_fwd(m(G, B, T, dtype=torch.int64), m(V, H), lstm, m(G, V + 1), None, L, 0, 0, False)
    with pytest.raises(RuntimeError, match="whh_l0_mask must be"):
        ops.encoder_fwd(m(G, B, T, dtype=torch.int64), m(V, H), lstm, None, m(G, 4 * H, H + 1), L, 0, 0, False)
    with pytest.raises(RuntimeError, match="int64 / int32 / int16 / uint8"):
        ops.encoder_fwd(m(G, B, T, dtype=torch.float32), m(V, H), lstm, None, None, L, 0, 0, False)
    z, pooled, argmax = ops.pool_fc_fwd(hn, m(H, H), m(H), 2)
    assert z.shape == (G * B, H) and pooled.shape == (G * B, H) and argmax.shape == (G * B, H) and argmax.dtype == torch.uint8
    assert ops.pool_fc_fwd(hn, m(H, H), m(H), 0)[2].numel() == 0
    d_hn, fc_flat = ops.pool_fc_bwd(z, pooled, None, m(H, H), 0)
    assert d_hn.shape == (2, G * B, H) and fc_flat.shape == (H * H + H,)
    head = [m(H // 2, H), m(H // 2), m(1, H // 2), m(1)]
    losses, y_hat = ops.loss_head_fwd(m(5, B, H), m(B, dtype=torch.int64), head, [None] * 4, 2.0)
    assert losses.shape == (3,) and y_hat.shape == (B,)
    dz, hflat = ops.loss_head_bwd(m(5, B, H), m(B, dtype=torch.int64), head + [m(H, H), m(H)], [None] * 4, 2.0, m(1), None)
    assert dz.shape == (5, B, H) and hflat.numel() == (H // 2) * H + H // 2 + H // 2 + 1 + H * H + H
    assert ops.pair_score(m(7, H), *head, None, None).shape == (28,)
    assert ops.pair_score(m(7, H), *head, m(11, dtype=torch.int32), m(11, dtype=torch.int32)).shape == (11,)
    assert ops.pair_score_range(m(7, H), *head, 3, 9).shape == (9,)
    mt, conf = ops.batch_metrics(m(B), m(B, dtype=torch.int64), 0.5)
    assert mt.shape == (5,) and conf.shape == (4,) and conf.dtype == torch.int32


def test_torch_custom_ops_are_registered_cuda_only():
    """The launchers are torch custom ops (torch.ops.intrepppid_b200.*) with a CUDA kernel only: the dispatcher has nothing to run
    for CPU tensors, so there is no silent fallback."""
    import torch

    import intrepppid_b200  # noqa: F401  (registers the ops)

    ops = torch.ops.intrepppid_b200
    for name in OP_NAMES:
        schema = str(getattr(ops, name).default._schema)
        assert schema.startswith(f"intrepppid_b200::{name}("), schema
        assert torch._C._dispatch_has_kernel_for_dispatch_key(f"intrepppid_b200::{name}", "CUDA")
        assert torch._C._dispatch_has_kernel_for_dispatch_key(f"intrepppid_b200::{name}", "Meta")
        assert not torch._C._dispatch_has_kernel_for_dispatch_key(f"intrepppid_b200::{name}", "CPU")
    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.pool_fc_fwd(torch.zeros(2, 3, 32), torch.zeros(32, 32), torch.zeros(32), 0)


def test_side_entry_points_validate_arguments_before_any_launch(lib):
    """ib200_adamw_step / ib200_batch_metrics / ib200_draw_masks reject bad arguments on the host (no GPU needed, nothing launched)."""
    import ctypes as C

    from intrepppid_b200._lib import AdamWHyper, MaskSpec

    launches = lib.ib200_launch_count()
    h = AdamWHyper(1e-3, 0.9, 0.999, 1e-8, 1e-2, 1.0, 1, 0)
    assert lib.ib200_adamw_step(0, None, None, None, None, None, h, None) == 0           # nothing to do
    assert lib.ib200_adamw_step(-1, None, None, None, None, None, h, None) == -2
    assert lib.ib200_adamw_step(2, None, None, None, None, None, h, None) == -1 and b"null" in lib.ib200_last_error()
    ptrs, n = (C.c_void_p * 1)(0), (C.c_int64 * 1)(4)
    bad_step = AdamWHyper(1e-3, 0.9, 0.999, 1e-8, 1e-2, 1.0, 0, 0)
    assert lib.ib200_adamw_step(1, ptrs, ptrs, ptrs, ptrs, n, bad_step, None) == -2      # steps count from 1
    bad_beta = AdamWHyper(1e-3, 1.0, 0.999, 1e-8, 1e-2, 1.0, 1, 0)
    assert lib.ib200_adamw_step(1, ptrs, ptrs, ptrs, ptrs, n, bad_beta, None) == -2      # torch.optim.AdamW raises on beta1 = 1 too
    assert lib.ib200_batch_metrics(0, None, None, 0.5, None, None, None) == -2
    assert lib.ib200_batch_metrics(2000, None, None, 0.5, None, None, None) == -2        # one CTA: B <= 1024
    assert lib.ib200_batch_metrics(8, None, None, 0.5, None, None, None) == -1
    used = C.c_uint64(7)
    assert lib.ib200_draw_masks(0, None, 1, 0, C.byref(used), None) == 0 and used.value == 0
    spec = (MaskSpec * 1)(MaskSpec(None, 16, 0.0, 0))
    assert lib.ib200_draw_masks(1, spec, 1, 0, None, None) == -2                          # keep_prob must be in (0, 1]
    spec[0] = MaskSpec(None, 16, 0.7, 0)
    assert lib.ib200_draw_masks(1, spec, 1, 0, None, None) == -1                          # null output
    assert lib.ib200_launch_count() == launches
