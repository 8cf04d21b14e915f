"""Builds intrepppid_b200/libib200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m intrepppid_b200.build [--verbose] [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with the source tree.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libib200.so")
TORCH_LIB = os.path.join(PKG, "libib200_torch.so")  # TORCH_LIBRARY shim (csrc/torch_ops.cpp) over the C ABI
OBJDIR = os.path.join(ROOT, "build", "obj")
SOURCES = ["capi.cu", "small.cu", "head.cu", "lstm_fwd.cu", "lstm_bwd.cu", "lstm_cluster.cu", "lstm_cluster_tc.cu", "gemm.cu", "gemm_tc.cu", "gemm_l0.cu", "gemm_wide.cu", "optim.cu", "ranger21.cu", "metrics.cu", "masks.cu", "p2p.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
              "-Xptxas", "-v", "--expt-relaxed-constexpr"]
if os.environ.get("IB200_PROF"):  # timing experiments only: per-phase cycle counters printed by the tcgen05 cluster kernels
    NVCC_FLAGS.append("-DIB200_PROF")
if os.environ.get("IB200_ABLATE"):  # timing experiments only (tools/ablate_fwd.py): in-kernel ablation switches of IB200_DBG
    NVCC_FLAGS.append("-DIB200_ABLATE")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libib200.so cannot be built (there is no CPU fallback)")


def _digest() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join(ROOT, "include", "ib200.h")]
    for f in files:
        path = f if os.path.isabs(f) else os.path.join(CSRC, f)
        with open(path, "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(verbose: bool = False, force: bool = False) -> str:
    stamp = os.path.join(OBJDIR, "stamp")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(TORCH_LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    with ThreadPoolExecutor(max_workers=len(SOURCES) + 1) as ex:
        shim = ex.submit(_compile_torch_shim)
        results = list(ex.map(compile_one, SOURCES))
    log = []
    for src, obj, r in results:
        log.append(f"== {src}\n{r.stdout}{r.stderr}")
        if r.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src}")
    with open(os.path.join(OBJDIR, "ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    r = subprocess.run([nvcc, "-shared", "-o", LIB, *[o for _, o, _ in results], "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    _link_torch_shim(shim.result())
    with open(stamp, "w") as fh:
        fh.write(dig)
    return LIB


def _compile_torch_shim() -> str:
    """g++ -c of the operator library (schemas + CUDA dispatch + Meta kernels; no CUDA code in it: it only calls the C ABI).
    Runs beside the nvcc jobs; linked by _link_torch_shim once libib200.so exists."""
    import torch
    from torch.utils import cpp_extension as ce

    cxx = os.environ.get("CXX") or shutil.which("g++")
    if not cxx:
        raise RuntimeError("g++ not found: libib200_torch.so cannot be built")
    obj = os.path.join(OBJDIR, "torch_ops.o")
    inc = [f"-I{p}" for p in ce.include_paths("cuda")]
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-c", "-Wno-deprecated-declarations",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", *inc, os.path.join(CSRC, "torch_ops.cpp"), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("torch shim compile failed:\n" + r.stdout + r.stderr)
    return obj


def _link_torch_shim(obj: str) -> str:
    """Link against libib200.so (NEEDED entry, found through the $ORIGIN rpath) and the torch libraries of the running interpreter."""
    import torch

    cxx = os.environ.get("CXX") or shutil.which("g++")
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = [cxx, "-shared", obj, "-o", TORCH_LIB, f"-L{tlib}", "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_cuda", "-lc10_cuda",
           "-Wl,-rpath,$ORIGIN", "-Wl,--no-as-needed", f"-L{PKG}", "-l:libib200.so"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("torch shim link failed:\n" + r.stdout + r.stderr)
    return TORCH_LIB


if __name__ == "__main__":
    print(build(verbose="--verbose" in sys.argv, force="--force" in sys.argv))
