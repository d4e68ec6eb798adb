"""The drop-in claim, tested at the reference's own interface: graph_core's UNMODIFIED solver classes (RRT,
RRTStar, BiRRT, AnytimeRRT; oracle/_ref/libgc_ref_dropin.so holds their objects compiled from /root/reference)
run with graph::core::CudaCollisionChecker / CudaNearestNeighbors -- graph_core_b200/host/cuda_plugins.cpp
compiled against the GENUINE graph_core headers -- and must build exactly the trees and paths they build with
their CPU KdTree and a per-state CPU checker."""
import numpy as np
import pytest

from graph_core_b200 import worlds as W

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    import ref_py
    try:
        ref_py.lib(dropin=True)
    except (FileNotFoundError, OSError) as ex:
        pytest.skip("oracle/_ref/libgc_ref_dropin.so unavailable: %s" % ex)
    return ref_py


def _cpu_checker(ref, orc, sc):
    if sc.kind == "cube":
        return ref.RefChecker.cube(sc.dim, sc.params["threshold"], sc.min_distance)
    return ref.RefChecker.over_oracle_world(orc.OracleWorld.from_scenario(sc), sc.min_distance)


def _same(ref, a, b):
    assert a.solved == b.solved
    assert a.cost == b.cost or (np.isinf(a.cost) and np.isinf(b.cost))
    assert a.tree_size == b.tree_size
    ta, tb = ref.tree_as_edges(*a.dump()), ref.tree_as_edges(*b.dump())
    assert ta == tb
    assert np.array_equal(a.solution(), b.solution())


@pytest.mark.parametrize("name,iters", [("C0", 1200), ("C1", 1500), ("C3", 250)])
def test_rrtstar_with_cuda_plugins_builds_the_reference_tree(gc, orc, ref, c3, name, iters):
    sc = {"C0": W.scenario_c0, "C1": W.scenario_c1}.get(name, lambda: c3)()
    samples = W.uniform_samples(sc.lb, sc.ub, 9100 + sc.dim, 0, iters)
    cpu = ref.RefSolver(ref.RefSolver.RRTSTAR, sc, _cpu_checker(ref, orc, sc), informed=True)
    gpu = ref.DropinSolver(ref.RefSolver.RRTSTAR, sc, sc.make_world(), informed=True, cuda_nn=True)
    assert gpu.uses_cuda_nn
    for lo, hi in ((0, iters // 3), (iters // 3, iters)):
        assert cpu.run(samples[lo:hi]) == gpu.run(samples[lo:hi])
        _same(ref, cpu, gpu)
    assert gpu.device_calls > iters // 2


@pytest.mark.parametrize("extend", [False, True])
def test_rrt_with_cuda_plugins(gc, orc, ref, extend):
    sc = W.scenario_c1()
    samples = W.uniform_samples(sc.lb, sc.ub, 9200, 0, 4000)
    cpu = ref.RefSolver(ref.RefSolver.RRT, sc, _cpu_checker(ref, orc, sc), extend=extend, informed=False)
    gpu = ref.DropinSolver(ref.RefSolver.RRT, sc, sc.make_world(), extend=extend, informed=False, cuda_nn=True)
    assert cpu.run(samples) == gpu.run(samples)
    assert cpu.solved
    _same(ref, cpu, gpu)


def test_birrt_with_cuda_checker_c2(gc, orc, ref):
    """BASELINE config 2: 6-DoF BiRRT vs C-space spheres.  BiRRT builds its goal tree itself, so only the checker is
    the CUDA one here (the NearestNeighbors hook needs the upstream factory patch of INTEGRATION.md)."""
    sc = W.scenario_c2()
    samples = W.uniform_samples(sc.lb, sc.ub, 9300, 0, 3000)
    cpu = ref.RefSolver(ref.RefSolver.BIRRT, sc, _cpu_checker(ref, orc, sc), informed=False)
    gpu = ref.DropinSolver(ref.RefSolver.BIRRT, sc, sc.make_world(), informed=False, cuda_nn=False)
    assert cpu.run(samples) == gpu.run(samples)
    assert cpu.solved
    _same(ref, cpu, gpu)
    assert gpu.device_calls > 10


@pytest.mark.parametrize("name", ["C1", "C3"])
def test_anytime_rrt_with_cuda_checker(gc, orc, ref, c3, name):
    """BASELINE config 5's solver: AnytimeRRT::solve with its own (std::rand) informed sampler, same seed on both sides."""
    sc = W.scenario_c1() if name == "C1" else c3
    iters = 1500 if name == "C1" else 400
    L = ref.lib()  # std::rand is process-global: seeding through either library seeds both solvers
    L.ref_srand(12345)
    cpu = ref.RefSolver(ref.RefSolver.ANYTIME, sc, _cpu_checker(ref, orc, sc), informed=True)
    ok_cpu = cpu.solve(iters, 1.0e9)
    # the CUDA side is fully warmed up BEFORE seeding: lazy kernel loading inside the driver may draw from rand()
    gw = sc.make_world()
    q = W.uniform_samples(sc.lb, sc.ub, 5, 0, 4)
    gw.check(q)
    gw.check_connection(q[:2], q[2:], sc.min_distance)
    gw.check_connection(q[:2], q[2:], sc.min_distance, mode=1)
    gw.check_conn_from_conf(q[:2], q[2:], q[:2], sc.min_distance)
    L.ref_srand(12345)
    gpu = ref.DropinSolver(ref.RefSolver.ANYTIME, sc, gw, informed=True, cuda_nn=False)
    ok_gpu = gpu.solve(iters, 1.0e9)
    assert ok_cpu == ok_gpu
    _same(ref, cpu, gpu)
    assert gpu.device_calls > 10
