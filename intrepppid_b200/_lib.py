"""ctypes binding of libib200.so (include/ib200.h).  There is NO fallback: if the CUDA library is missing or cannot be
loaded, importing the compute path raises."""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libib200.so")

MAX_LAYERS = 4
REDUCE = {"last": 0, "mean": 1, "max": 2}
PRECISION = {"fp32": 0, "bf16": 1}
TOKEN_DTYPE = {"int64": 0, "int32": 1, "int16": 2, "uint8": 3}  # IB200_TOK_*

F = C.POINTER(C.c_float)
vp = C.c_void_p


class Cfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("G", "B", "T", "V", "H", "L", "bi_reduce", "precision", "training", "token_dtype")]


class EncoderParams(C.Structure):
    _fields_ = [("emb", vp), ("w_ih", vp * 2 * MAX_LAYERS), ("w_hh", vp * 2 * MAX_LAYERS), ("b_ih", vp * 2 * MAX_LAYERS),
                ("b_hh", vp * 2 * MAX_LAYERS)]


class HeadParams(C.Structure):
    _fields_ = [(n, vp) for n in ("fc1_w", "fc1_b", "fc2_w", "fc2_b", "proj_w", "proj_b")]


class AdamWHyper(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("lr", "beta1", "beta2", "eps", "weight_decay", "grad_scale")] + [("step", C.c_int32),
                                                                                                            ("maximize", C.c_int32)]


class Ranger21Tensor(C.Structure):
    _fields_ = [(n, vp) for n in ("param", "grad", "grad_ma", "neg_grad_ma", "variance_ma", "lookahead")] + [
        ("rows", C.c_int64), ("cols", C.c_int64), ("multi_dim", C.c_int32), ("step", C.c_int32), ("lr", C.c_double)]


class Ranger21Hyper(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("beta1", "beta2", "eps", "weight_decay", "agc_clip", "agc_eps", "normloss_factor",
                                          "softplus_beta", "pnm_factor", "lookahead_alpha")] + [
        (n, C.c_int32) for n in ("use_agc", "use_gc", "use_gcnorm", "use_normloss", "use_softplus", "lookahead_merge")]


class MaskSpec(C.Structure):
    _fields_ = [("out", vp), ("numel", C.c_int64), ("keep_prob", C.c_float), ("row_len", C.c_int32)]


class HeadMasks(C.Structure):
    _fields_ = [(n, vp) for n in ("fc1_w", "do1", "do2", "fc2_w")]


EXPORTS = ("ib200_version", "ib200_last_error", "ib200_workspace_bytes", "ib200_launch_count", "ib200_timing_enable",
           "ib200_timing_families", "ib200_timing_family_name", "ib200_timing_read", "ib200_encoder_fwd", "ib200_encoder_status", "ib200_encoder_bwd", "ib200_encoder_bwd_layers",
           "ib200_pool_fc_fwd", "ib200_pool_fc_bwd", "ib200_loss_head_fwd", "ib200_loss_head_bwd", "ib200_pair_score", "ib200_pair_score_range", "ib200_sequence_lengths", "ib200_adamw_step", "ib200_ranger21_step", "ib200_ranger21_scratch_bytes", "ib200_batch_metrics", "ib200_draw_masks", "ib200_p2p_allreduce_mean", "ib200_p2p_alloc", "ib200_p2p_open", "ib200_p2p_close", "ib200_p2p_free",
           "ib200_dbg_gemm_nt", "ib200_dbg_gemm_tn", "ib200_dbg_gemm_nt_planes", "ib200_dbg_gemm_tn_planes", "ib200_dbg_l0_scratch_floats",
           "ib200_dbg_l0_grads", "ib200_dbg_gemm_nt_wide", "ib200_dbg_gemm_tn_wide")

_lib = None


class IB200Error(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IB200Error(f"{LIB_PATH} is missing: build it with `python -m intrepppid_b200.build` "
                         "(intrepppid_b200 has no CPU or PyTorch fallback)")
    L = C.CDLL(LIB_PATH)
    L.ib200_version.restype = C.c_int
    L.ib200_last_error.restype = C.c_char_p
    L.ib200_workspace_bytes.restype = C.c_size_t
    L.ib200_workspace_bytes.argtypes = [C.POINTER(Cfg)]
    L.ib200_encoder_fwd.argtypes = [C.POINTER(Cfg), vp, C.POINTER(EncoderParams), vp, vp, vp, vp, vp, C.c_size_t, vp]
    L.ib200_encoder_status.argtypes = [C.POINTER(Cfg), vp, C.c_size_t, vp, vp]
    L.ib200_encoder_bwd.argtypes = [C.POINTER(Cfg), C.POINTER(EncoderParams), vp, vp, vp, C.POINTER(EncoderParams), vp,
                                    C.c_size_t, vp]
    L.ib200_encoder_bwd_layers.argtypes = [C.POINTER(Cfg), C.POINTER(EncoderParams), vp, vp, vp, C.POINTER(EncoderParams), vp,
                                           C.c_size_t, C.c_int32, C.c_int32, vp]
    L.ib200_pool_fc_fwd.argtypes = [C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, vp, vp, vp, vp]
    L.ib200_pool_fc_bwd.argtypes = [C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ib200_loss_head_fwd.argtypes = [C.c_int32, C.c_int32, C.c_float, vp, vp, C.POINTER(HeadParams), C.POINTER(HeadMasks), vp,
                                      vp, vp]
    L.ib200_loss_head_bwd.argtypes = [C.c_int32, C.c_int32, C.c_float, vp, vp, C.POINTER(HeadParams), C.POINTER(HeadMasks), vp,
                                      vp, vp, C.POINTER(HeadParams), vp]
    L.ib200_pair_score.argtypes = [C.c_int32, C.c_int32, vp, vp, vp, C.c_int64, C.POINTER(HeadParams), vp, vp]
    L.ib200_pair_score_range.argtypes = [C.c_int32, C.c_int32, vp, C.c_int64, C.c_int64, C.POINTER(HeadParams), vp, vp]
    L.ib200_sequence_lengths.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, C.c_int32, vp, vp, vp, vp, vp]
    L.ib200_adamw_step.argtypes = [C.c_int32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_int64),
                                   C.POINTER(AdamWHyper), vp]
    L.ib200_ranger21_scratch_bytes.restype = C.c_size_t
    L.ib200_ranger21_scratch_bytes.argtypes = [C.c_int32, C.POINTER(Ranger21Tensor)]
    L.ib200_ranger21_step.argtypes = [C.c_int32, C.POINTER(Ranger21Tensor), C.POINTER(Ranger21Hyper), vp, vp]
    L.ib200_p2p_allreduce_mean.argtypes = [C.c_int32, C.c_int32, C.POINTER(vp), C.POINTER(vp), C.c_size_t, vp, C.c_size_t, C.c_uint32, vp]
    L.ib200_p2p_alloc.argtypes = [C.c_size_t, C.POINTER(vp), C.c_char_p]
    L.ib200_p2p_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.ib200_p2p_close.argtypes = [vp]
    L.ib200_p2p_free.argtypes = [vp]
    L.ib200_draw_masks.argtypes = [C.c_int32, C.POINTER(MaskSpec), C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), vp]
    L.ib200_batch_metrics.argtypes = [C.c_int32, vp, vp, C.c_float, vp, vp, vp]
    L.ib200_dbg_gemm_nt.argtypes = [C.c_int32, C.c_int32, C.c_int32, vp, C.c_int32, vp, vp, C.c_int32, C.c_int32, vp, vp, vp, vp,
                                    C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]
    L.ib200_dbg_gemm_tn.argtypes = [C.c_int32, C.c_int32, C.c_int32, vp, vp, C.c_int32, vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp,
                                    C.c_int32, C.c_int32, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]
    i32 = C.c_int32
    L.ib200_dbg_gemm_nt_planes.argtypes = [i32, i32, i32, vp, i32, vp, vp, i32, i32, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    L.ib200_dbg_gemm_tn_planes.argtypes = [i32, i32, i32, vp, vp, vp, i32, i32, i32, vp, vp, vp, i32, i32, vp, i32, i32, i32, i32, vp, i32,
                                           i32, vp]
    L.ib200_dbg_gemm_nt_wide.argtypes = [i32, i32, i32, vp, i32, vp, vp, i32, i32, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    L.ib200_dbg_gemm_tn_wide.argtypes = [i32, i32, i32, vp, vp, i32, vp, i32, i32, i32, i32, vp, i32, i32, i32, i32, vp, i32, C.POINTER(i32),
                                         i32, vp]
    L.ib200_dbg_l0_scratch_floats.restype = C.c_size_t
    L.ib200_dbg_l0_scratch_floats.argtypes = [i32, i32, i32]
    L.ib200_dbg_l0_grads.argtypes = [i32, i32, i32, i32, vp, vp, C.POINTER(vp), vp, vp, vp, vp, C.POINTER(vp), vp, i32, i32, i32, vp, vp,
                                     C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), vp, i32, vp]
    L.ib200_launch_count.restype = C.c_ulonglong
    L.ib200_timing_family_name.restype = C.c_char_p
    L.ib200_timing_family_name.argtypes = [C.c_int]
    L.ib200_timing_enable.argtypes = [C.c_int]
    L.ib200_timing_read.argtypes = [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    for name in EXPORTS:
        if name not in ("ib200_version", "ib200_last_error", "ib200_workspace_bytes", "ib200_launch_count",
                        "ib200_timing_family_name", "ib200_dbg_l0_scratch_floats", "ib200_ranger21_scratch_bytes"):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().ib200_last_error().decode("utf-8", "replace")
        kind = "invalid argument" if status < 0 else "CUDA error"
        raise IB200Error(f"{what}: {kind} {status}: {msg}")


def ptr(t) -> int | None:
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def timing_enable(on: bool) -> None:
    check(lib().ib200_timing_enable(1 if on else 0), "ib200_timing_enable")


def timing_read() -> dict:
    """{family: (total_ms, timed_launcher_calls)} since the last read (synchronises on the recorded events)."""
    L = lib()
    n = L.ib200_timing_families()
    ms, cnt = (C.c_float * n)(), (C.c_int * n)()
    check(L.ib200_timing_read(n, ms, cnt), "ib200_timing_read")
    return {L.ib200_timing_family_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i] > 0}


def launch_count() -> int:
    return int(lib().ib200_launch_count())
