"""In-tree builds of the native libraries (nvcc / g++).  Outputs travel to the GPU box with the snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
LIB_CUDA = os.path.join(PKG, "libb200rt.so")
LIB_HOST = os.path.join(PKG, "librtiaw_host.so")
APP = os.path.join(PKG, "rtiaw_app")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "--fmad=false",          # the reference (Rust) never contracts a*b+c: bit-parity needs single-rounded ops
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off",
    "-shared",
]
GXX_FLAGS = ["-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-Wall", "-Wextra"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")
    return exe


def _stale(out: str, srcs: list[str]) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in srcs if os.path.exists(s))


def _sources(d: str, exts: tuple[str, ...]) -> list[str]:
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts))


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources(CSRC, (".cu", ".cuh", ".hpp")) + [os.path.join(ROOT, "include", "b200rt.h")]
    if force or _stale(LIB_CUDA, srcs):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_CUDA, os.path.join(CSRC, "b200rt.cu")]
        subprocess.check_call(cmd)
    return LIB_CUDA


def build_host(force: bool = False) -> str:
    srcs = _sources(HOST, (".cpp", ".hpp")) + [os.path.join(ROOT, "include", "b200rt.h"), os.path.join(ROOT, "include", "rtiaw_host.h")]
    if force or _stale(LIB_HOST, srcs):
        build_cuda()
        cmd = ["g++"] + GXX_FLAGS + ["-I", os.path.join(ROOT, "include"), "-o", LIB_HOST, os.path.join(HOST, "rtiaw_capi.cpp"),
               "-L", PKG, "-lb200rt", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
    return LIB_HOST


def build_app(force: bool = False) -> str:
    """The reference CLI (src/app/main.rs) over the host mirror."""
    srcs = _sources(HOST, (".cpp", ".hpp")) + [os.path.join(ROOT, "include", "b200rt.h")]
    if force or _stale(APP, srcs):
        build_cuda()
        flags = [f for f in GXX_FLAGS if f not in ("-shared", "-fPIC")]
        cmd = ["g++"] + flags + ["-I", os.path.join(ROOT, "include"), "-o", APP, os.path.join(HOST, "main.cpp"),
               "-L", PKG, "-lb200rt", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
    return APP


def build_all(force: bool = False) -> None:
    build_cuda(force)
    build_host(force)
    build_app(force)
