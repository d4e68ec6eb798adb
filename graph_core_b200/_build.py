"""Build the sm_100a shared library and the C++ host layer.

The library is built IN-TREE with explicit nvcc calls so that the `.so` travels with a
repository snapshot to the GPU box (a JIT cache under ~/.cache would not).  nvcc cross-compiles
without a GPU, so this also runs in the CPU-only build container.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
HOST = PKG / "host"
BUILD = ROOT / "build"
LIB = PKG / "libgcb200.so"
HOSTLIB = PKG / "libgcb200_host.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-ffp-contract=off",  # host-side fp64 decisions (anytime.cu) follow the reference's rounding
    "-diag-suppress", "128",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def _newer(target: Path, sources) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(s).stat().st_mtime <= t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(map(str, cmd)), r.stdout[-4000:], r.stderr[-4000:]))
    return r


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    """Compile graph_core_b200/csrc/*.cu for sm_100a into graph_core_b200/libgcb200.so."""
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + sorted((ROOT / "include").glob("*.h"))
    if not force and _newer(LIB, sources + headers):
        return LIB
    nvcc = _nvcc()
    BUILD.mkdir(exist_ok=True)

    def compile_one(src: Path) -> Path:
        obj = BUILD / (src.stem + ".o")
        if not force and _newer(obj, [src] + headers):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd), flush=True)
        _run(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs),
           "--cudart=static", "-ldl"]
    if verbose:
        print(" ".join(cmd), flush=True)
    _run(cmd)
    return LIB


def build_host(force: bool = False, verbose: bool = False) -> Path | None:
    """Compile the C++ host layer (plugin classes + planners) that sits above the C ABI."""
    sources = sorted(HOST.glob("*.cpp"))
    if not sources:
        return None
    headers = sorted(HOST.glob("*.h")) + sorted((ROOT / "include").glob("*.h"))
    build_cuda(force=False, verbose=verbose)
    if not force and _newer(HOSTLIB, sources + headers + [LIB]):
        return HOSTLIB
    cxx = shutil.which("g++") or "g++"
    cmd = [cxx, "-std=c++17", "-O2", "-fno-fast-math", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
           "-I", str(ROOT / "include"), "-I", str(HOST), "-o", str(HOSTLIB), *map(str, sources),
           "-L", str(PKG), "-lgcb200", "-Wl,-rpath,$ORIGIN"]
    if verbose:
        print(" ".join(cmd), flush=True)
    _run(cmd)
    return HOSTLIB


def build_probe(force: bool = False, verbose: bool = False) -> Path | None:
    """Compile tests/cpp/plugin_probe.cpp (C++ exerciser of the plugin classes) against the host library."""
    src = ROOT / "tests" / "cpp" / "plugin_probe.cpp"
    if not src.exists():
        return None
    out = ROOT / "tests" / "cpp" / "plugin_probe"
    hostlib = build_host(force=False, verbose=verbose)
    headers = sorted(HOST.rglob("*.h")) + sorted((ROOT / "include").glob("*.h"))
    if not force and _newer(out, [src, hostlib] + headers):
        return out
    cxx = shutil.which("g++") or "g++"
    cmd = [cxx, "-std=c++17", "-O2", "-fno-fast-math", "-ffp-contract=off", "-I", str(ROOT / "include"), "-I", str(HOST),
           "-o", str(out), str(src), "-L", str(PKG), "-lgcb200_host", "-lgcb200", "-pthread",
           "-Wl,-rpath," + str(PKG), "-Wl,-rpath,$ORIGIN/../../graph_core_b200"]
    if verbose:
        print(" ".join(cmd), flush=True)
    _run(cmd)
    return out


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cuda(force, verbose)
    build_host(force, verbose)
    build_probe(force, verbose)


if __name__ == "__main__":
    import sys

    build_all(force="--force" in sys.argv, verbose=True)
    print("built", LIB)
