"""pytest configuration: markers, import paths, shared fixtures.

`-m "not gpu"` covers the oracle, the host logic and that the C-ABI library loads and exports
every symbol of include/gcb200.h.  `-m gpu` tests are the parity tests proper: they call the
CUDA path through the C ABI and compare it with the oracle on identical seeded inputs.
"""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def gc():
    """The product package with the CUDA library loaded (fails loudly when it is not built)."""
    import graph_core_b200 as g
    g.capi.load()
    return g


@pytest.fixture(scope="session")
def c3(orc):
    from graph_core_b200 import worlds
    return worlds.scenario_c3(1000, orc.arm_checker())
