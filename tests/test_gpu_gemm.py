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


# ==================================================================================================================================
# The kernels the headline bench actually runs (H = 64 "planes" path): gemm_nt_tma_kernel, gemm_tn_tma_kernel, l0_grad_gemm_kernel.
# Operands are bf16 hi|lo planes: a row of K values lives in the bytes of K floats as [K bf16 hi | K bf16 lo].
# ==================================================================================================================================
def _planes(x: torch.Tensor) -> torch.Tensor:
    """fp32 [rows, K] -> the plane layout, returned as float32 [rows, K] (a byte container) + the value the planes represent."""
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, lo], dim=-1).contiguous().view(torch.float32), hi, lo


def _plane_value(hi, lo, precision):
    return (hi.double() + lo.double()) if precision == 0 else hi.double()


def _valid_rows(G, B, T, lens_eff):
    valid = torch.zeros(G, B, T, dtype=torch.bool)
    for gi, te in enumerate(lens_eff):
        valid[gi, :, :te] = True
    return valid.reshape(-1).cuda()


def _run_nt_planes(precision, G, B, T, lens_eff, nsrc, K, NC, bias=True, accumulate=False, seed=0):
    from intrepppid_b200._lib import check, lib, ptr

    g = torch.Generator().manual_seed(seed)
    rows = G * B * T
    packs = [_planes(torch.randn(rows, K, generator=g).cuda()) for _ in range(nsrc)]
    W = [(torch.randn(NC, K, generator=g) * 0.1).cuda() for _ in range(nsrc)]
    b = torch.randn(NC, generator=g).cuda() if bias else None
    C0 = torch.randn(rows, NC, generator=g).cuda()
    C = C0.clone()
    lens = torch.tensor([[T] * G, lens_eff], dtype=torch.int32).cuda()
    st = torch.cuda.current_stream().cuda_stream
    check(lib().ib200_dbg_gemm_nt_planes(G, B, T, ptr(lens), nsrc, ptr(packs[0][0]), ptr(packs[1][0]) if nsrc > 1 else None, K, K,
                                         ptr(W[0]), ptr(W[1]) if nsrc > 1 else None, ptr(b), ptr(C), NC, NC, int(accumulate), precision,
                                         st), "ib200_dbg_gemm_nt_planes")
    torch.cuda.synchronize()
    # the kernel splits W into bf16 hi + lo itself (fp32 mode) or rounds it to bf16 (bf16 mode): the products below are what it forms
    ref = 0
    for (_, hi, lo), w in zip(packs, W):
        ref = ref + _plane_value(hi, lo, precision) @ (w.double() if precision == 0 else w.to(torch.bfloat16).double()).T
    if bias:
        ref = ref + b.double()
    if accumulate:
        ref = ref + C0.double()
    valid = _valid_rows(G, B, T, lens_eff)
    err = (C.double() - ref)[valid].norm() / ref[valid].norm()
    # output contract: C = ... goes out through TMA stores of whole 32-row boxes, so a row t >= T_eff may be written (with a value
    # nobody reads) when its 32-row box holds a valid row; boxes without a valid row -- and every row t >= T_eff of the C += path,
    # which stores row by row -- keep their contents
    keep = ~valid
    if not accumulate:
        pad = (-rows) % 32
        live_box = torch.nn.functional.pad(valid, (0, pad)).view(-1, 32).any(dim=1).repeat_interleave(32)[:rows]
        keep = ~live_box
    return float(err), torch.equal(C[keep], C0[keep])


def test_nt_tma_row_by_row_epilogue_still_matches():
    """IB200_NO_TMA_STORE=1 keeps the thread-store epilogue (also the C += path): same results, strict row masking."""
    import subprocess, sys
    code = ("import torch, sys; sys.path.insert(0, 'tests'); import test_gpu_gemm as t; "
            "e, u = t._run_nt_planes(0, G=2, B=3, T=300, lens_eff=[257, 100], nsrc=1, K=128, NC=256, accumulate=True); "
            "assert e < 1e-5 and u; "
            "e, u = t._run_nt_planes(0, G=4, B=40, T=256, lens_eff=[256, 255, 129, 1], nsrc=1, K=128, NC=256, seed=3); assert e < 2e-5 and u")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, IB200_NO_TMA_STORE="1"), capture_output=True, text=True, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("nsrc,K,NC", [(1, 128, 256), (1, 256, 128), (2, 256, 128), (2, 256, 64)])  # xproj, dY (1 / 2 live directions), dX0
def test_nt_tma_planes_fp32_mode(nsrc, K, NC):
    err, untouched = _run_nt_planes(0, G=2, B=3, T=300, lens_eff=[257, 100], nsrc=nsrc, K=K, NC=NC)
    assert err < 2e-5, err      # hi*hi + hi*lo + lo*hi, fp32 accumulation: the dropped lo*lo term is ~2^-16 relative
    assert untouched, "rows outside the live 32-row boxes must not be written"


def test_nt_tma_planes_bf16_mode_accumulate_and_many_tiles():
    err, untouched = _run_nt_planes(1, G=1, B=5, T=200, lens_eff=[131], nsrc=1, K=128, NC=256, accumulate=True)
    assert err < 1e-5 and untouched   # against the bf16-rounded operands the kernel really multiplies: only accumulation order differs
    # more tiles than SMs (persistent loop, TMEM double buffering, phase wrap-around), ragged groups incl. T_eff = 1 and a dead group tail
    err, untouched = _run_nt_planes(0, G=4, B=40, T=256, lens_eff=[256, 255, 129, 1], nsrc=1, K=128, NC=256, seed=3)
    assert err < 2e-5 and untouched


def _run_tn_planes(precision, G, B, T, lens_eff, NB1, first, second, ctas=3, seed=0):
    """first: 'dense' | 'gather' (scale*emb rows by token); second: None | 'prev' (shift -1) | 'next' (shift +1), 64 columns taken
    at column offset col02 of a 128-wide plane matrix -- exactly the [dW_ih | dW_hh] call of the encoder backward."""
    from intrepppid_b200._lib import check, lib, ptr

    g = torch.Generator().manual_seed(seed)
    KA, rows, V = 256, G * B * T, 50
    NB2 = 64 if second else 0
    lens = torch.tensor([[T] * G, lens_eff], dtype=torch.int32).cuda()
    ldb, col0 = 2 * NB1, 0 if NB1 == 128 else 64   # H = 64: Y_{l-1} is [.,128] -> NB1 = 128 takes all of it; a 64-wide take starts at 64
    ldb = max(ldb, 128)
    valid = _valid_rows(G, B, T, lens_eff)
    # operand contract of the TMA kernels (whole 64-row boxes are loaded): rows t >= T_eff hold zeros in A and B -- the BPTT kernel
    # zeroes the dgate rows [T_eff, next multiple of 64] (lstm_bwd.cu), the forward kernel the same rows of Y (lstm_fwd.cu)
    Ap, Ahi, Alo = _planes(torch.randn(rows, KA, generator=g).cuda() * valid.unsqueeze(1))
    B1 = torch.randn(rows, ldb, generator=g).cuda() * valid.unsqueeze(1)     # rows t >= T_eff hold zeros (the forward zeroes the tail row)
    B1p, B1hi, B1lo = _planes(B1)
    B2 = torch.randn(rows, 128, generator=g).cuda() * valid.unsqueeze(1)
    B2p, B2hi, B2lo = _planes(B2)
    col02 = 64 if second == "next" else 0                                       # reverse direction reads the second half of Y_l
    shift2 = {None: 0, "prev": -1, "next": 1}[second]
    tok = torch.randint(0, V, (rows,), generator=g, dtype=torch.int32).cuda()
    emb = torch.randn(V, NB1, generator=g).cuda()
    scale = (torch.rand(G, V, generator=g) > 0.3).float().cuda() / 0.7
    NB = NB1 + NB2
    partial = torch.full((G, ctas, KA * NB), float("nan")).cuda()
    gather = first == "gather"
    st = torch.cuda.current_stream().cuda_stream
    check(lib().ib200_dbg_gemm_tn_planes(G, B, T, ptr(lens), ptr(Ap), None if gather else ptr(B1p), ldb, col0, 0,
                                         ptr(tok) if gather else None, ptr(emb) if gather else None, ptr(scale) if gather else None, V, NB1,
                                         ptr(B2p) if second else None, 128, col02, shift2, NB2, ptr(partial), ctas, precision, st),
          "ib200_dbg_gemm_tn_planes")
    torch.cuda.synchronize()
    got = partial.sum(1).double().view(G, KA, NB)
    A3 = _plane_value(Ahi, Alo, precision).view(G, B, T, KA)
    b1v = _plane_value(B1hi, B1lo, precision).view(G, B, T, ldb)[..., col0:col0 + NB1]
    b2v = _plane_value(B2hi, B2lo, precision).view(G, B, T, 128)[..., col02:col02 + 64]
    tok3 = tok.view(G, B, T).long()
    worst = 0.0
    for gi, te in enumerate(lens_eff):
        a = A3[gi, :, :te]
        if gather:
            e = scale[gi][tok3[gi, :, :te]].unsqueeze(-1) * emb[tok3[gi, :, :te]]   # fp32 product, then split by the kernel
            b = e.double() if precision == 0 else e.to(torch.bfloat16).double()
        else:
            b = b1v[gi, :, :te]
        if second:
            s = torch.zeros(B, te, 64, dtype=torch.float64, device="cuda")
            if shift2 == -1:
                s[:, 1:] = b2v[gi, :, :te - 1]
            else:
                s[:, :te - 1] = b2v[gi, :, 1:te]
            b = torch.cat([b, s], dim=-1)
        ref = torch.einsum("btk,btn->kn", a, b)
        worst = max(worst, float((got[gi] - ref).norm() / ref.norm()))
    return worst


@pytest.mark.parametrize("NB1,first,second", [(128, "dense", "prev"), (128, "dense", "next"), (128, "dense", None), (64, "dense", None),
                                              (64, "gather", None), (64, "gather", "prev"), (64, "gather", "next")])
def test_tn_tma_planes_fp32_mode(NB1, first, second):
    err = _run_tn_planes(0, G=2, B=3, T=200, lens_eff=[200, 77], NB1=NB1, first=first, second=second)
    assert err < 2e-5, err


def test_tn_tma_planes_bf16_mode_more_ctas_than_items_and_short_groups():
    assert _run_tn_planes(1, G=1, B=2, T=100, lens_eff=[65], NB1=128, first="dense", second="prev", ctas=7) < 1e-5
    assert _run_tn_planes(0, G=3, B=5, T=130, lens_eff=[1, 64, 129], NB1=128, first="dense", second="next", ctas=2, seed=5) < 2e-5


def _run_l0_grads(precision, G, B, T, lens_eff, V, dir0, ndir, masked=True, seed=0):
    """ib200_dbg_l0_grads (l0_grad_gemm_kernel + reduce + finish) against fp64: dW_hh, dW_ih, both bias gradients, dEmb."""
    import ctypes as C

    from intrepppid_b200._lib import check, lib, ptr

    H, g = 64, torch.Generator().manual_seed(seed)
    rows = G * B * T
    valid = _valid_rows(G, B, T, lens_eff)
    dA = [_planes(torch.randn(rows, 4 * H, generator=g).cuda() * valid.unsqueeze(1)) for _ in range(2)]  # GI columns k = 4u + q; zero tail rows
    Y0p, Yhi, Ylo = _planes(torch.randn(rows, 2 * H, generator=g).cuda() * valid.unsqueeze(1))
    tok = torch.randint(0, V, (rows,), generator=g, dtype=torch.int32).cuda()
    emb = torch.randn(V, H, generator=g).cuda()
    scale = ((torch.rand(G, V, generator=g) > 0.3).float() / 0.7).cuda() if masked else None
    whm = ((torch.rand(G, 4 * H, H, generator=g) > 0.3).float() / 0.7).cuda() if masked else None
    w_ih = [torch.randn(4 * H, H, generator=g).cuda() * 0.1 for _ in range(2)]
    bias_count = 3
    lens = torch.tensor([[T] * G, lens_eff], dtype=torch.int32).cuda()
    bias_partial = torch.randn(ndir, bias_count, 4 * H, generator=g).cuda()
    partial = torch.empty(lib().ib200_dbg_l0_scratch_floats(G, ndir, 0), device="cuda")
    scratch = torch.empty(lib().ib200_dbg_l0_scratch_floats(G, ndir, 1), device="cuda")
    out = {n: [torch.full((4 * H, H) if n in ("wih", "whh") else (4 * H,), float("nan"), device="cuda") for _ in range(2)]
           for n in ("wih", "whh", "bih", "bhh")}
    d_emb = torch.full((V, H), float("nan"), device="cuda")
    arr = lambda ts: (C.c_void_p * 2)(*[ptr(t) for t in ts])  # noqa: E731
    st = torch.cuda.current_stream().cuda_stream
    check(lib().ib200_dbg_l0_grads(G, B, T, V, ptr(lens), ptr(tok), arr([dA[0][0], dA[1][0]]), ptr(Y0p), ptr(emb), ptr(scale), ptr(whm),
                                   arr(w_ih), ptr(bias_partial), bias_count, dir0, ndir, ptr(partial), ptr(scratch), arr(out["wih"]),
                                   arr(out["whh"]), arr(out["bih"]), arr(out["bhh"]), ptr(d_emb), precision, st), "ib200_dbg_l0_grads")
    torch.cuda.synchronize()
    # torch row r = q*H + u  <->  gate-interleaved column k = 4u + q
    k_of_row = torch.tensor([4 * (r % H) + r // H for r in range(4 * H)], device="cuda")
    Yv = _plane_value(Yhi, Ylo, precision).view(G, B, T, 2 * H)
    tok3 = tok.view(G, B, T).long()
    worst, demb_ref = 0.0, torch.zeros(V, H, dtype=torch.float64, device="cuda")
    for ds in range(ndir):
        d = dir0 + ds
        Av = _plane_value(dA[d][1], dA[d][2], precision).view(G, B, T, 4 * H)[..., k_of_row]   # torch row order
        dwhh = torch.zeros(4 * H, H, dtype=torch.float64, device="cuda")
        dwih = torch.zeros(4 * H, H, dtype=torch.float64, device="cuda")
        for gi, te in enumerate(lens_eff):
            a = Av[gi, :, :te]
            hp = torch.zeros(B, te, H, dtype=torch.float64, device="cuda")   # h of the previous scan position
            if d == 0:
                hp[:, 1:] = Yv[gi, :, :te - 1, :H]
            else:
                hp[:, :te - 1] = Yv[gi, :, 1:te, H:]
            m = whm[gi].double() if (whm is not None and d == 0) else 1.0
            dwhh += m * torch.einsum("btk,bth->kh", a, hp)
            sc = scale[gi].double() if scale is not None else torch.ones(V, dtype=torch.float64, device="cuda")
            x = sc[tok3[gi, :, :te]].unsqueeze(-1) * emb.double()[tok3[gi, :, :te]]
            dwih += torch.einsum("btk,bth->kh", a, x)
            dx = a @ w_ih[d].double()                                        # [B, te, H]
            contrib = torch.zeros(V, H, dtype=torch.float64, device="cuda")
            contrib.index_add_(0, tok3[gi, :, :te].reshape(-1), dx.reshape(-1, H))
            demb_ref += sc.unsqueeze(1) * contrib
        bsum = bias_partial[ds].double().sum(0)[k_of_row]
        for name, ref in (("whh", dwhh), ("wih", dwih), ("bih", bsum), ("bhh", bsum)):
            worst = max(worst, float((out[name][d].double() - ref).norm() / ref.norm()))
    demb_ref[0] = 0  # padding_idx
    assert float(d_emb[0].abs().max()) == 0.0
    worst = max(worst, float((d_emb.double() - demb_ref).norm() / demb_ref.norm()))
    for d in range(2):  # a direction that was not run is not written
        if not (dir0 <= d < dir0 + ndir):
            assert bool(torch.isnan(out["wih"][d]).all())
    return worst


@pytest.mark.parametrize("dir0,ndir", [(0, 2), (1, 1)])
def test_l0_grad_kernels_fp32_mode(dir0, ndir):
    err = _run_l0_grads(0, G=2, B=5, T=150, lens_eff=[150, 67], V=250, dir0=dir0, ndir=ndir)
    assert err < 2e-5, err


def test_l0_grad_kernels_bf16_mode_unmasked_small_vocabulary():
    assert _run_l0_grads(1, G=3, B=2, T=70, lens_eff=[1, 64, 70], V=21, dir0=0, ndir=2, masked=False, seed=2) < 1e-5


# ==================================================================================================================================
# gemm_wide.cu: the streamed-operand tcgen05 + TMA kernels behind the cluster recurrent kernels (H = 128, 192, 256; config 5 = 256)
# ==================================================================================================================================
def _run_nt_wide(precision, G, B, T, lens_eff, nsrc, K, NC, bias=True, accumulate=False, seed=0):
    from intrepppid_b200._lib import check, lib, ptr

    g = torch.Generator().manual_seed(seed)
    rows = G * B * T
    packs = [_planes(torch.randn(rows, K, generator=g).cuda()) for _ in range(nsrc)]
    wpacks = [_planes((torch.randn(NC, K, generator=g) * 0.1).cuda()) for _ in range(nsrc)]
    b = torch.randn(NC, generator=g).cuda() if bias else None
    C0 = torch.randn(rows, NC, generator=g).cuda()
    C = C0.clone()
    lens = torch.tensor([[T] * G, lens_eff], dtype=torch.int32).cuda()
    st = torch.cuda.current_stream().cuda_stream
    check(lib().ib200_dbg_gemm_nt_wide(G, B, T, ptr(lens), nsrc, ptr(packs[0][0]), ptr(packs[1][0]) if nsrc > 1 else None, K, K,
                                       ptr(wpacks[0][0]), ptr(wpacks[1][0]) if nsrc > 1 else None, ptr(b), ptr(C), NC, NC, int(accumulate),
                                       precision, st), "ib200_dbg_gemm_nt_wide")
    torch.cuda.synchronize()
    ref = 0
    for (_, hi, lo), (_, whi, wlo) in zip(packs, wpacks):
        ref = ref + _plane_value(hi, lo, precision) @ _plane_value(whi, wlo, precision).T
    if bias:
        ref = ref + b.double()
    if accumulate:
        ref = ref + C0.double()
    valid = _valid_rows(G, B, T, lens_eff)
    err = (C.double() - ref)[valid].norm() / ref[valid].norm()
    return float(err), torch.equal(C[~valid], C0[~valid])


@pytest.mark.parametrize("H", [128, 192, 256])
def test_nt_wide_fp32_mode_all_three_gemm_shapes(H):
    for nsrc, K, NC in ((1, 2 * H, 4 * H), (2, 4 * H, 2 * H), (2, 4 * H, H)):   # xproj, dY (both directions live), dX0
        err, untouched = _run_nt_wide(0, G=2, B=3, T=150, lens_eff=[131, 64], nsrc=nsrc, K=K, NC=NC, seed=H + NC)
        assert err < 2e-5, (H, nsrc, K, NC, err)
        assert untouched, "rows outside the live 32-row boxes must not be written"


def test_nt_wide_bf16_mode_accumulate_and_persistent_loop():
    err, untouched = _run_nt_wide(1, G=1, B=5, T=200, lens_eff=[131], nsrc=1, K=512, NC=1024, accumulate=True)
    assert err < 1e-5 and untouched
    # more (row tile, column tile) items than SMs: persistent loop, TMEM double buffering, stage phase wrap-around, dead tiles skipped
    err, untouched = _run_nt_wide(0, G=4, B=20, T=256, lens_eff=[256, 255, 129, 1], nsrc=1, K=256, NC=512, seed=3)
    assert err < 2e-5 and untouched


def _run_tn_wide(precision, H, G, B, T, lens_eff, l0, second, splits=0, seed=0):
    """[dW_ih | dW_hh] of one (layer, direction): A = dgates planes [rows, 4H]; B1 = Y_{l-1} planes [rows, 2H] (all columns) or, for
    layer 0, the gathered input planes [rows, H]; B2 = H columns of Y_l planes at column offset 0 / H, shifted by -1 / +1."""
    import ctypes as C

    from intrepppid_b200._lib import check, lib, ptr

    g = torch.Generator().manual_seed(seed)
    KA, rows = 4 * H, G * B * T
    NB1, NB2 = (H if l0 else 2 * H), H
    lens = torch.tensor([[T] * G, lens_eff], dtype=torch.int32).cuda()
    valid = _valid_rows(G, B, T, lens_eff)
    Ap, Ahi, Alo = _planes(torch.randn(rows, KA, generator=g).cuda() * valid.unsqueeze(1))
    B1p, B1hi, B1lo = _planes(torch.randn(rows, NB1, generator=g).cuda() * valid.unsqueeze(1))
    B2p, B2hi, B2lo = _planes(torch.randn(rows, 2 * H, generator=g).cuda() * valid.unsqueeze(1))
    col02, shift2 = (H, 1) if second == "next" else (0, -1)
    n = C.c_int32(0)
    st = torch.cuda.current_stream().cuda_stream
    args = (G, B, T, ptr(lens), ptr(Ap), KA, ptr(B1p), NB1, 0, 0, NB1, ptr(B2p), 2 * H, col02, shift2, NB2)
    check(lib().ib200_dbg_gemm_tn_wide(*args, None, splits, C.byref(n), precision, st), "ib200_dbg_gemm_tn_wide (query)")
    NB = NB1 + NB2
    partial = torch.full((G, n.value, KA * NB), float("nan")).cuda()
    check(lib().ib200_dbg_gemm_tn_wide(*args, ptr(partial), splits, C.byref(n), precision, st), "ib200_dbg_gemm_tn_wide")
    torch.cuda.synchronize()
    got = partial.sum(1).double().view(G, KA, NB)
    A3 = _plane_value(Ahi, Alo, precision).view(G, B, T, KA)
    b1v = _plane_value(B1hi, B1lo, precision).view(G, B, T, NB1)
    b2v = _plane_value(B2hi, B2lo, precision).view(G, B, T, 2 * H)[..., col02:col02 + H]
    worst = 0.0
    for gi, te in enumerate(lens_eff):
        a = A3[gi, :, :te]
        s = torch.zeros(B, te, H, dtype=torch.float64, device="cuda")
        if shift2 == -1:
            s[:, 1:] = b2v[gi, :, :te - 1]
        else:
            s[:, :te - 1] = b2v[gi, :, 1:te]
        ref = torch.einsum("btk,btn->kn", a, torch.cat([b1v[gi, :, :te], s], dim=-1))
        worst = max(worst, float((got[gi] - ref).norm() / ref.norm()))
    return worst


@pytest.mark.parametrize("H", [128, 192, 256])
@pytest.mark.parametrize("l0,second", [(False, "prev"), (False, "next"), (True, "prev")])
def test_tn_wide_fp32_mode(H, l0, second):
    err = _run_tn_wide(0, H, G=2, B=3, T=200, lens_eff=[200, 77], l0=l0, second=second, seed=H)
    assert err < 2e-5, err


def test_tn_wide_bf16_mode_many_splits_and_short_groups():
    assert _run_tn_wide(1, 256, G=1, B=2, T=100, lens_eff=[65], l0=False, second="prev", splits=5) < 1e-5   # more splits than items
    assert _run_tn_wide(0, 128, G=3, B=5, T=130, lens_eff=[1, 64, 129], l0=True, second="next", splits=2, seed=5) < 2e-5
