"""The drop-in C++ plugin classes (graph_core_b200/host/cuda_plugins.h): CudaNearestNeighbors behind
graph::core::NearestNeighbors, CudaCollisionChecker behind graph::core::CollisionCheckerBase, loaded by
their cnr_class_loader names.  tests/cpp/plugin_probe.cpp drives them the way Tree and the solvers do."""
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PROBE = ROOT / "tests" / "cpp" / "plugin_probe"


def _build():
    from graph_core_b200 import _build
    return _build.build_probe()


def test_probe_builds_and_refuses_to_run_without_a_device():
    exe = _build()
    assert exe is not None and exe.exists()
    from graph_core_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    assert "PROBE OK" not in r.stdout
    assert "no CPU fallback" in r.stdout or "plugin init failed" in r.stdout


def test_plugin_classes_override_the_reference_virtuals():
    """Same names / signatures as nearest_neighbors.h:59-173 and collision_checker_base.h:166-319."""
    h = (ROOT / "graph_core_b200" / "host" / "cuda_plugins.h").read_text()
    for sig in [r"void insert\(const NodePtr& node\) override", r"bool clear\(\) override",
                r"void nearestNeighbor\(const Eigen::VectorXd& configuration, NodePtr& best, double& best_distance\) override",
                r"std::multimap<double, NodePtr> near\(const Eigen::VectorXd& configuration, const double& radius\) override",
                r"std::multimap<double, NodePtr> kNearestNeighbors\(const Eigen::VectorXd& configuration, const size_t& k\) override",
                r"bool findNode\(const NodePtr& node\) override",
                r"bool deleteNode\(const NodePtr& node, const bool& disconnect_node = false\) override",
                r"bool restoreNode\(const NodePtr& node\) override", r"std::vector<NodePtr> getNodes\(\) override",
                r"void disconnectNodes\(const std::vector<NodePtr>& white_list\) override",
                r"bool check\(const Eigen::VectorXd& configuration\) override",
                r"bool checkConnection\(const ConnectionPtr& conn\) override",
                r"bool checkConnections\(const std::vector<ConnectionPtr>& connections\) override",
                r"bool checkConnFromConf\(const ConnectionPtr& conn, const Eigen::VectorXd& this_conf\) override",
                r"CollisionCheckerPtr clone\(\) override"]:
        assert re.search(sig, h), sig
    cpp = (ROOT / "graph_core_b200" / "host" / "cuda_plugins.cpp").read_text()
    assert "CLASS_LOADER_REGISTER_CLASS(graph::core::CudaCollisionCheckerPlugin, graph::core::CollisionCheckerBasePlugin)" in cpp
    assert "CLASS_LOADER_REGISTER_CLASS(graph::core::CudaNearestNeighborsPlugin, graph::core::NearestNeighborsPlugin)" in cpp


@pytest.mark.gpu
def test_plugin_probe_on_gpu():
    exe = _build()
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "PROBE OK" in r.stdout
