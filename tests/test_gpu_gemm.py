"""The token-row GEMMs in isolation (test hook ib200_dbg_gemm_nt): legacy mma.sync and tcgen05 kernels against torch fp64."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run_nt(impl, precision, G, B, T, lens_eff, nsrc, K, NC, bias=True, accumulate=False, seed=0):
    from intrepppid_b200 import _lib
    from intrepppid_b200._lib import check, lib, ptr

    g = torch.Generator().manual_seed(seed)
    rows = G * B * T
    A = [torch.randn(rows, K, generator=g).cuda() for _ in range(nsrc)]
    W = [(torch.randn(NC, K, generator=g) * 0.1).cuda() for _ in range(nsrc)]
    b = torch.randn(NC, generator=g).cuda() if bias else None
    C0 = torch.randn(rows, NC, generator=g).cuda()
    C = C0.clone()
    lens = torch.tensor([[T] * G, lens_eff], dtype=torch.int32).cuda()
    st = torch.cuda.current_stream().cuda_stream
    check(lib().ib200_dbg_gemm_nt(G, B, T, ptr(lens), nsrc, ptr(A[0]), ptr(A[1]) if nsrc > 1 else None, K, K, ptr(W[0]),
                                  ptr(W[1]) if nsrc > 1 else None, ptr(b), ptr(C), NC, NC, int(accumulate), precision, impl, st),
          "ib200_dbg_gemm_nt")
    torch.cuda.synchronize()
    ref = sum(a.double() @ w.double().T for a, w in zip(A, W))
    if bias:
        ref = ref + b.double()
    if accumulate:
        ref = ref + C0.double()
    valid = torch.zeros(G, B, T, dtype=torch.bool)
    for gi, te in enumerate(lens_eff):
        valid[gi, :, :te] = True
    valid = valid.reshape(-1).cuda()
    err = (C.double() - ref)[valid].norm() / ref[valid].norm()
    untouched = torch.equal(C[~valid], C0[~valid])
    return float(err), untouched


SHAPES = [  # (nsrc, K, NC): xproj, dY (1 and 2 sources), dX0
    (1, 128, 256), (1, 256, 128), (2, 256, 128), (2, 256, 64)]


@pytest.mark.parametrize("nsrc,K,NC", SHAPES)
@pytest.mark.parametrize("impl", [0, 1, 2])
def test_nt_fp32_mode(nsrc, K, NC, impl):
    if impl == 1 and (nsrc, K, NC) == (2, 256, 128):
        pytest.skip("W of both sources exceeds shared memory: the auto path splits it into two accumulate passes")
    err, untouched = _run_nt(impl, 0, G=2, B=3, T=300, lens_eff=[257, 100], nsrc=nsrc, K=K, NC=NC)
    assert err < 2e-5, err
    assert untouched, "rows with t >= T_eff must not be written"


@pytest.mark.parametrize("impl", [0, 1])
def test_nt_bf16_mode_and_accumulate(impl):
    err, untouched = _run_nt(impl, 1, G=1, B=5, T=200, lens_eff=[131], nsrc=1, K=128, NC=256, accumulate=True)
    assert err < 1e-2 and untouched


def test_nt_many_tiles_persistent_loop():
    """More tiles than SMs so every CTA loops (TMEM double buffering, barrier phase wrap-around)."""
    err, untouched = _run_nt(1, 0, G=4, B=40, T=256, lens_eff=[256, 255, 129, 1], nsrc=1, K=128, NC=256, seed=3)
    assert err < 2e-5 and untouched


def _run_tn(impl, precision, G, B, T, lens_eff, NB, mode, ctas=3, colsum=True, seed=0):
    """mode: 'dense' (shift 0), 'prev' (shift -1), 'next' (shift +1), 'gather' (embedding rows by token)."""
    from intrepppid_b200._lib import check, lib, ptr

    g = torch.Generator().manual_seed(seed)
    KA, rows, V = 256, G * B * T, 50
    A = torch.randn(rows, KA, generator=g).cuda()
    lens = torch.tensor([[T] * G, lens_eff], dtype=torch.int32).cuda()
    ldb, col0 = 2 * NB, NB  # take the second half of a wider matrix
    Bsrc = torch.randn(rows, ldb, generator=g).cuda()
    tok = torch.randint(0, V, (rows,), generator=g, dtype=torch.int32).cuda()
    emb = torch.randn(V, NB, generator=g).cuda()
    scale = (torch.rand(G, V, generator=g) > 0.3).float().cuda() / 0.7
    shift = {"dense": 0, "prev": -1, "next": 1, "gather": 0}[mode]
    blk = KA * NB + (KA if colsum else 0)
    partial = torch.full((G, ctas, blk), float("nan")).cuda()
    st = torch.cuda.current_stream().cuda_stream
    gather = mode == "gather"
    check(lib().ib200_dbg_gemm_tn(G, B, T, ptr(lens), ptr(A), KA, None if gather else ptr(Bsrc), ldb, col0, shift,
                                  ptr(tok) if gather else None, ptr(emb) if gather else None, ptr(scale) if gather else None, V, NB,
                                  ptr(partial), ctas, int(colsum), precision, impl, st), "ib200_dbg_gemm_tn")
    torch.cuda.synchronize()
    got = partial.sum(1).double()  # [G, blk]
    worst = 0.0
    A3, B3, tok3 = A.view(G, B, T, KA).double(), Bsrc.view(G, B, T, ldb)[..., col0:col0 + NB].double(), tok.view(G, B, T).long()
    for gi, te in enumerate(lens_eff):
        a = A3[gi, :, :te]
        if gather:
            b = (scale[gi][tok3[gi, :, :te]].unsqueeze(-1) * emb[tok3[gi, :, :te]]).double()
        else:
            b = torch.zeros(B, te, NB, dtype=torch.float64, device="cuda")
            if shift == 0:
                b = B3[gi, :, :te]
            elif shift == -1:
                b[:, 1:] = B3[gi, :, :te - 1]
            else:
                b[:, :te - 1] = B3[gi, :, 1:te]
        ref = torch.einsum("btk,btn->kn", a, b)
        worst = max(worst, float((got[gi, :KA * NB].view(KA, NB) - ref).norm() / ref.norm()))
        if colsum:
            cs = a.sum((0, 1))
            worst = max(worst, float((got[gi, KA * NB:] - cs).norm() / cs.norm()))
    return worst


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("NB,mode", [(128, "dense"), (64, "prev"), (64, "next"), (64, "gather")])
def test_tn_fp32_mode(impl, NB, mode):
    err = _run_tn(impl, 0, G=2, B=3, T=200, lens_eff=[200, 77], NB=NB, mode=mode)
    assert err < 2e-5, err


@pytest.mark.parametrize("impl", [0, 1])
def test_tn_bf16_mode_more_ctas_than_items(impl):
    err = _run_tn(impl, 1, G=1, B=2, T=100, lens_eff=[65], NB=64, mode="dense", ctas=7)
    assert err < 1e-2, err
