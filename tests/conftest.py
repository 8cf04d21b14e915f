import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    import torch

    from oracle import ref_shim

    have_gpu = torch.cuda.is_available()
    have_ref = ref_shim.available()
    for item in items:
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "reference" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="/root/reference not mounted"))


@pytest.fixture(autouse=True)
def _no_leftover_status_words():
    """The encoder records reference-style errors on the device and raises them lazily (ops.check_pending): a test must not
    inherit an error recorded by the previous one."""
    yield
    mod = sys.modules.get("intrepppid_b200.ops")
    if mod is not None:
        mod._PENDING.clear()


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["train_last_E64", "train_mean_proj_E32", "train_max_L3_E32", "train_nodrop_E32"]


def load_golden(name):
    import torch

    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


def rel_l2(a, b):
    """max relative L2 error used for every floating-point gate: ||a-b|| / max(||b||, tiny)."""
    import torch

    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float(torch.linalg.norm(a - b) / max(float(torch.linalg.norm(b)), 1e-30))


def assert_grad_close(got, ref, tol, name=""):
    """Gradient gate: relative L2 error <= tol, except for tensors whose reference gradient is (a) exactly zero
    (dead chain, Q16: must be exactly zero too) or (b) pure rounding noise (e.g. the triplet-projection bias, which
    cancels analytically in a-p and a-n): there only an absolute bound is meaningful."""
    import torch

    ref_norm = float(torch.linalg.norm(ref.detach().double().cpu()))
    if float(ref.abs().max()) == 0.0:
        assert float(got.abs().max()) == 0.0, f"{name}: reference gradient is exactly zero"
    elif ref_norm < 1e-6:
        assert float(torch.linalg.norm(got.detach().double().cpu())) < 1e-5, f"{name}: noise-level gradient"
    else:
        err = rel_l2(got, ref)
        assert err < tol, f"{name}: rel L2 error {err:.3e} >= {tol:g}"
