"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/gcb200.h declares, and refuses to work without a CUDA device (no CPU fallback)."""
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "gcb200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from graph_core_b200 import capi
    lib = capi.load()
    declared = _declared_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), "libgcb200.so does not export %s" % name
    # the ctypes table covers the whole header, nothing more, nothing less
    assert sorted(capi.PROTOTYPES) == declared
    assert lib.gc_version() >= 100


def test_header_cites_reference_interfaces():
    text = (ROOT / "include" / "gcb200.h").read_text()
    for cite in ("nearest_neighbors.h:41-195", "collision_checker_base.h", "kdtree.cpp", "vector.cpp", "tree.cpp"):
        assert cite in text


def test_no_cpu_fallback_without_device():
    from graph_core_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.GcError):
        capi.NodeStore(3, 16)
    with pytest.raises(capi.GcError):
        capi.CollisionWorld.cube(3, 1.0)
    with pytest.raises(capi.GcError):
        capi.sample_uniform(np.zeros(2), np.ones(2), 1, 0, 4)
    assert capi.load().gc_last_error().decode() != ""  # the reason is reported, e.g. "no CUDA-capable device"


def test_product_never_touches_the_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    for p in list((ROOT / "graph_core_b200").rglob("*")):
        if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".h"):
            body = p.read_text()
            assert "oracle_py" not in body and "gc_oracle" not in body and "libgc_oracle" not in body, p


def test_philox_numpy_matches_known_answer():
    """Philox4x32-10 known-answer test (Random123 kat_vectors: counter/key all zero and all ones)."""
    from graph_core_b200.worlds import philox4x32_10
    z = np.zeros(1, np.uint64)
    out = [int(v[0]) for v in philox4x32_10(z, z, z, z, 0, 0)]
    assert out == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = np.full(1, 0xFFFFFFFF, np.uint64)
    out = [int(v[0]) for v in philox4x32_10(f, f, f, f, 0xFFFFFFFF, 0xFFFFFFFF)]
    assert out == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
