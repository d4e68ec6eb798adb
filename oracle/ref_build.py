#!/usr/bin/env python
"""Builds oracle/_ref/libgc_ref.so: the graph_core reference's OWN translation units, compiled where
they lie under /root/reference (never copied) against the stand-in headers of oracle/ref_shim/
(Eigen, cnr_logger, cnr_param, yaml-cpp -- none of which exist in this image), plus oracle/ref_driver.cpp.

Test infrastructure only.  oracle/_ref/ is git-ignored but travels to the GPU box.  Where
/root/reference is absent (the GPU box) this script does nothing and the prebuilt library is used.

  libgc_ref.so       -O2 -ffp-contract=off     parity build (pins oracle/gc_oracle.cpp)
  libgc_ref_fast.so  the reference's Release flags (graph_core/CMakeLists.txt:13: -O3 -Ofast -funroll-loops)
                     plus -mfma -mavx2 -fno-finite-math-only: the CPU timing baseline ("kind": "reference")
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path(os.environ.get("GRAPH_CORE_REFERENCE", "/root/reference")) / "graph_core"
OUT = HERE / "_ref"

# the plugin registrations (src/graph_core/plugins/**) need cnr_class_loader and are not on the hot path
SOURCES = [
    "src/graph_core/datastructure/kdtree.cpp", "src/graph_core/datastructure/vector.cpp",
    "src/graph_core/graph/node.cpp", "src/graph_core/graph/connection.cpp", "src/graph_core/graph/tree.cpp",
    "src/graph_core/graph/subtree.cpp", "src/graph_core/graph/net.cpp", "src/graph_core/graph/path.cpp",
    "src/graph_core/metrics/euclidean_metrics.cpp",
    "src/graph_core/samplers/uniform_sampler.cpp", "src/graph_core/samplers/informed_sampler.cpp",
    "src/graph_core/samplers/ball_sampler.cpp", "src/graph_core/samplers/tube_informed_sampler.cpp",
    "src/graph_core/solvers/tree_solver.cpp", "src/graph_core/solvers/rrt.cpp", "src/graph_core/solvers/rrt_star.cpp",
    "src/graph_core/solvers/birrt.cpp", "src/graph_core/solvers/anytime_rrt.cpp",
    "src/graph_core/solvers/path_optimizers/path_optimizer_base.cpp",
    "src/graph_core/solvers/path_optimizers/path_local_optimizer.cpp",
]
VARIANTS = {
    "libgc_ref.so": ["-O2", "-fno-fast-math", "-ffp-contract=off", "-mfma"],
    "libgc_ref_fast.so": ["-O3", "-Ofast", "-fno-finite-math-only", "-funroll-loops", "-mfma", "-mavx2"],
}


def host_compiler():
    """The system g++ first: a $CXX wrapper that only finds libstdc++.a links the C++ runtime statically into the
    library, and a second copy of the iostream/locale machinery inside a Python process (numpy and torch load the
    shared one) crashes on the first formatted double (Tree::toYAML)."""
    import shutil
    return shutil.which("g++") or os.environ.get("CXX") or "g++"


def newest_input():
    files = [HERE / "ref_driver.cpp", Path(__file__)] + list((HERE / "ref_shim").rglob("*")) + [REF / s for s in SOURCES]
    return max(f.stat().st_mtime for f in files if f.is_file())


def build(verbose=False):
    if not REF.exists():
        if verbose:
            print("ref_build: %s not present; using prebuilt oracle/_ref if any" % REF)
        return False
    OUT.mkdir(exist_ok=True)
    stamp = newest_input()
    cxx = host_compiler()
    inc = ["-I", str(HERE / "ref_shim"), "-I", str(REF / "include"), "-include", str(HERE / "ref_shim" / "shim_random.hpp")]
    for lib, flags in VARIANTS.items():
        target = OUT / lib
        if target.exists() and target.stat().st_mtime >= stamp:
            continue
        objdir = OUT / ("obj_" + lib.replace(".so", ""))
        objdir.mkdir(exist_ok=True)
        jobs = [(REF / s, objdir / (Path(s).stem + ".o")) for s in SOURCES] + [(HERE / "ref_driver.cpp", objdir / "ref_driver.o")]

        def cc(job):
            src, obj = job
            cmd = [cxx, "-std=c++17", "-fPIC", "-w", "-DNDEBUG", *flags, *inc, "-c", str(src), "-o", str(obj)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("ref_build: %s failed\n%s" % (src, r.stderr[-4000:]))
            return obj

        with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
            objs = list(ex.map(cc, jobs))
        cmd = [cxx, "-shared", "-o", str(target), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("ref_build: link failed\n" + r.stderr[-4000:])
        if verbose:
            print("ref_build: built", target)
    return True


def patched_tree_sources():
    """INTEGRATION.md section 3 applied to a BUILD-TIME copy of the reference's tree.h / tree.cpp under the git-ignored
    oracle/_ref/patched/: a process-wide NearestNeighbors factory consulted by Tree::Tree before the stock
    KdTree / Vector choice (tree.cpp:43-50).  The hook adds two static members (no change of the object layout), so
    every other reference object stays as compiled from the untouched sources; nothing is copied into the repository."""
    pdir = OUT / "patched"
    hdir = pdir / "include" / "graph_core" / "graph"
    hdir.mkdir(parents=True, exist_ok=True)
    h = (REF / "include/graph_core/graph/tree.h").read_text()
    anchor = "  bool use_kdtree_;"
    assert h.count(anchor) == 1, "tree.h changed upstream: revisit the factory-hook patch"
    h = h.replace(anchor, anchor + """
public:
  using NearestNeighborsFactory = std::function<NearestNeighborsPtr(const cnr_logger::TraceLoggerPtr&)>;
  static void setNearestNeighborsFactory(NearestNeighborsFactory f) { nn_factory_() = std::move(f); }
protected:
  static NearestNeighborsFactory& nn_factory_() { static NearestNeighborsFactory f; return f; }
""", 1)
    h = "#include <functional>\n" + h
    (hdir / "tree.h").write_text(h)
    c = (REF / "src/graph_core/graph/tree.cpp").read_text()
    anchor = "  if (use_kdtree)\n  {\n    nodes_ = std::make_shared<KdTree>(logger_);"
    assert c.count(anchor) == 1, "tree.cpp changed upstream: revisit the factory-hook patch"
    c = c.replace(anchor, "  if (nn_factory_())\n  {\n    nodes_ = nn_factory_()(logger_);\n  }\n  else if (use_kdtree)\n  {\n"
                          "    nodes_ = std::make_shared<KdTree>(logger_);", 1)
    (pdir / "tree.cpp").write_text(c)
    return pdir


def build_dropin(verbose=False):
    """oracle/_ref/libgc_ref_dropin.so: the reference's objects + graph_core_b200/host/cuda_plugins.cpp compiled with
    -DGCB200_WITH_GRAPH_CORE against the GENUINE graph_core headers + oracle/ref_dropin.cpp, linked with libgcb200.so
    (loads anywhere; needs a GPU only when a CUDA world is actually used)."""
    if not REF.exists():
        return False
    root = HERE.parent
    pkg = root / "graph_core_b200"
    lib = pkg / "libgcb200.so"
    if not lib.exists():
        raise RuntimeError("ref_build: build graph_core_b200/libgcb200.so first")
    target = OUT / "libgc_ref_dropin.so"
    srcs = [HERE / "ref_dropin.cpp", HERE / "ref_driver.cpp", pkg / "host" / "cuda_plugins.cpp", pkg / "host" / "cuda_plugins.h",
            pkg / "host" / "compat" / "graph_core_min.h", root / "include" / "gcb200.h", Path(__file__), lib]
    stamp = max(max(f.stat().st_mtime for f in srcs), newest_input())
    if target.exists() and target.stat().st_mtime >= stamp:
        return True
    build(verbose)  # the reference objects
    cxx = host_compiler()
    flags = VARIANTS["libgc_ref.so"]
    inc = ["-I", str(HERE / "ref_shim"), "-I", str(REF / "include"), "-I", str(root / "include"), "-I", str(pkg / "host"),
           "-I", str(HERE), "-include", str(HERE / "ref_shim" / "shim_random.hpp")]
    objdir = OUT / "obj_libgc_ref"
    # ref_dropin.cpp includes ref_driver.cpp; tree.o is rebuilt from the patched copy (NearestNeighbors factory hook)
    objs = [o for o in objdir.glob("*.o") if o.name not in ("ref_driver.o", "tree.o")]
    pdir = patched_tree_sources()
    inc = ["-I", str(pdir / "include")] + inc
    extra = []
    for src in (HERE / "ref_dropin.cpp", pkg / "host" / "cuda_plugins.cpp", pdir / "tree.cpp"):
        obj = OUT / ("dropin_" + src.stem + ".o")
        cmd = [cxx, "-std=c++17", "-fPIC", "-w", "-DNDEBUG", "-DGCB200_WITH_GRAPH_CORE", "-DGCB200_TREE_FACTORY_HOOK",
               *flags, *inc, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("ref_build: %s failed\n%s" % (src, r.stderr[-4000:]))
        extra.append(obj)
    cmd = [cxx, "-shared", "-o", str(target), *map(str, objs), *map(str, extra), "-L", str(pkg), "-lgcb200",
           "-Wl,-rpath,$ORIGIN/../../graph_core_b200"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("ref_build: drop-in link failed\n" + r.stderr[-4000:])
    if verbose:
        print("ref_build: built", target)
    return True


if __name__ == "__main__":
    build(verbose=True)
    build_dropin(verbose=True)
    sys.exit(0)
