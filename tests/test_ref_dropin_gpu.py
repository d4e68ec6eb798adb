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
    """BASELINE config 5's solver: AnytimeRRT::solve with its own informed sampler, same seed on both sides."""
    sc = W.scenario_c1() if name == "C1" else c3
    iters = 1500 if name == "C1" else 400
    # every draw of the reference (Eigen Random, the random_device behind each sampler's mt19937) comes from one
    # seedable stream per library (oracle/ref_shim/shim_random.hpp): same seed, same samples on both sides
    ref.lib().ref_srand(12345)
    cpu = ref.RefSolver(ref.RefSolver.ANYTIME, sc, _cpu_checker(ref, orc, sc), informed=True)
    ok_cpu = cpu.solve(iters, 1.0e9)
    ref.lib(dropin=True).ref_srand(12345)
    gpu = ref.DropinSolver(ref.RefSolver.ANYTIME, sc, sc.make_world(), informed=True, cuda_nn=False)
    ok_gpu = gpu.solve(iters, 1.0e9)
    assert ok_cpu == ok_gpu
    _same(ref, cpu, gpu)
    assert gpu.device_calls > 10


@pytest.fixture
def nn_factory(ref):
    """INTEGRATION.md section 3 applied (oracle/ref_build.py: patched_tree_sources): while set, every Tree the
    reference's solvers construct owns a CudaNearestNeighbors."""
    L = ref.lib(dropin=True)
    L.dropin_set_nn_factory(0)
    yield L
    L.dropin_set_nn_factory(-1)


def test_birrt_goal_tree_gets_the_cuda_nn_through_the_factory_hook(gc, orc, ref, nn_factory):
    """BiRRT builds its goal tree itself (birrt.cpp:67): with the Tree factory hook both trees query the CUDA store
    (nearest neighbour of extend / connect, tree.cpp:56-59, 217-233) and the bidirectional search is unchanged."""
    sc = W.scenario_c2()
    samples = W.uniform_samples(sc.lb, sc.ub, 9300, 0, 3000)
    cpu = ref.RefSolver(ref.RefSolver.BIRRT, sc, _cpu_checker(ref, orc, sc), informed=False)
    calls0 = nn_factory.dropin_nn_factory_calls()
    gpu = ref.DropinSolver(ref.RefSolver.BIRRT, sc, sc.make_world(), informed=False, cuda_nn=False)
    assert nn_factory.dropin_nn_factory_calls() - calls0 == 2  # start tree (rrt.cpp:73) + goal tree (birrt.cpp:67)
    assert cpu.run(samples) == gpu.run(samples)
    assert cpu.solved
    _same(ref, cpu, gpu)


@pytest.mark.parametrize("name", ["C1", "C3"])
def test_anytime_rrt_round_trees_get_the_cuda_nn_through_the_factory_hook(gc, orc, ref, c3, nn_factory, name):
    """AnytimeRRT::improve starts a fresh Tree per round (anytime_rrt.cpp:298) and grows it with Tree::informedExtend
    (tree.cpp:235-302), whose k-nearest query asks for k = nearK >= N in 7-D: every round's tree lives in a CUDA store
    and the solver takes the same decisions as with its KdTree."""
    sc = W.scenario_c1() if name == "C1" else c3
    iters = 1500 if name == "C1" else 400
    ref.lib().ref_srand(12345)
    cpu = ref.RefSolver(ref.RefSolver.ANYTIME, sc, _cpu_checker(ref, orc, sc), informed=True)
    ok_cpu = cpu.solve(iters, 1.0e9)
    ref.lib(dropin=True).ref_srand(12345)
    calls0 = nn_factory.dropin_nn_factory_calls()
    gpu = ref.DropinSolver(ref.RefSolver.ANYTIME, sc, sc.make_world(), informed=True, cuda_nn=False)
    ok_gpu = gpu.solve(iters, 1.0e9)
    assert ok_cpu == ok_gpu
    _same(ref, cpu, gpu)
    # the start tree (rrt.cpp:73) and every improve round that got past its utopia tests (anytime_rrt.cpp:298)
    assert nn_factory.dropin_nn_factory_calls() - calls0 >= 1


def _box_scene_with_new_obstacle(sc, path):
    """The C1 scene plus one box dropped onto the middle of the given path."""
    import copy
    mid = path[len(path) // 2]
    sc2 = copy.deepcopy(sc)
    sc2.params = {"lo": np.vstack([sc.params["lo"], mid - 0.04]), "hi": np.vstack([sc.params["hi"], mid + 0.04])}
    return sc2


def test_path_validity_and_local_optimizer_through_the_cuda_checker(gc, orc, ref):
    """SURVEY 8f rows N1-N3 at the reference's own call sites: Path::isValid / isValidFromConf (path.cpp:1230-1407),
    Tree::recheckCollision (tree.cpp:1405-1429) and PathLocalOptimizer::warp / simplify
    (path_local_optimizer.cpp:59-204) issue their checkConnection / checkConnFromConf calls to CudaCollisionChecker
    and must decide exactly as with the CPU checker."""
    sc = W.scenario_c1()
    samples = W.uniform_samples(sc.lb, sc.ub, 9400, 0, 2500)
    # both solvers live in the drop-in library so that checkers can be exchanged between them
    L = ref.lib(dropin=True)

    cpu_chk_in_dropin = ref.RefChecker(L.ref_checker_callback(sc.dim, *_cb(orc, sc), sc.min_distance), sc.dim, L,
                                       sc.min_distance)
    cpu = ref.RefSolver.__new__(ref.RefSolver)
    cpu.L, cpu.checker, cpu.dim = L, cpu_chk_in_dropin, sc.dim
    lb, ub, s0, g0 = (np.ascontiguousarray(x, dtype=np.float64) for x in (sc.lb, sc.ub, sc.start, sc.goal))
    import ctypes as C
    dp = C.POINTER(C.c_double)
    cpu.h = L.ref_solver_new(ref.RefSolver.RRTSTAR, sc.dim, cpu_chk_in_dropin.h, lb.ctypes.data_as(dp),
                             ub.ctypes.data_as(dp), s0.ctypes.data_as(dp), g0.ctypes.data_as(dp), sc.max_distance,
                             sc.rewire_factor, sc.utopia_tolerance, 1, 0, 1)
    gw = sc.make_world()
    gpu = ref.DropinSolver(ref.RefSolver.RRTSTAR, sc, gw, informed=True, cuda_nn=True)
    assert cpu.run(samples) == gpu.run(samples)
    _same(ref, cpu, gpu)
    assert cpu.solved
    path = cpu.solution()
    # N2: the untouched path is valid on both sides, connection costs re-evaluated identically
    v_cpu, c_cpu = cpu.solution_is_valid()
    v_gpu, c_gpu = gpu.solution_is_valid()
    assert v_cpu == v_gpu == 1 and np.array_equal(c_cpu, c_gpu)
    # N1: whole-tree re-validation
    assert cpu.recheck_tree() == gpu.recheck_tree()
    # a new obstacle on the path: same connections become infinite, same verdict from any point of the path on
    sc2 = _box_scene_with_new_obstacle(sc, path)
    ow2 = orc.OracleWorld.from_scenario(sc2)
    cpu_chk2 = ref.RefChecker(L.ref_checker_callback(sc.dim, C.cast(ow2.L.orc_check_states, C.c_void_p), ow2.h,
                                                     sc.min_distance), sc.dim, L, sc.min_distance, keep=(ow2,))
    gpu_chk2 = ref.cuda_ref_checker(sc2.make_world(), sc.min_distance)
    v_cpu, c_cpu = cpu.solution_is_valid(cpu_chk2)
    v_gpu, c_gpu = gpu.solution_is_valid(gpu_chk2)
    assert v_cpu == v_gpu == 0 and np.array_equal(c_cpu, c_gpu) and np.isinf(c_cpu).any()
    for ci in range(len(path) - 1):
        for t in (0.0, 0.37, 1.0):
            assert cpu.solution_is_valid_from(ci, t, cpu_chk2) == gpu.solution_is_valid_from(ci, t, gpu_chk2), (ci, t)
    # N3: local optimisation of the (restored) path takes the same decisions
    cpu.solution_is_valid()
    gpu.solution_is_valid()
    for mode, iters in ((0, 5), (1, 5), (2, 30)):
        wc, cc = cpu.optimize_solution(mode, iters)
        wg, cg = gpu.optimize_solution(mode, iters)
        assert np.array_equal(wc, wg) and cc == cg, mode
    assert gpu.device_calls > 100


def _cb(orc, sc):
    import ctypes as C
    ow = orc.OracleWorld.from_scenario(sc)
    _cb.keep = getattr(_cb, "keep", []) + [ow]
    return C.cast(ow.L.orc_check_states, C.c_void_p), ow.h


def _solved_pair(gc, orc, ref, sc, seed, iters):
    """The reference's RRT* with its CPU checker and with the CUDA plugins, both inside the drop-in library (so that
    checkers can be exchanged), run on the same samples until solved."""
    import ctypes as C
    L = ref.lib(dropin=True)
    samples = W.uniform_samples(sc.lb, sc.ub, seed, 0, iters)
    chk = ref.RefChecker(L.ref_checker_callback(sc.dim, *_cb(orc, sc), sc.min_distance), sc.dim, L, sc.min_distance)
    cpu = ref.RefSolver.__new__(ref.RefSolver)
    cpu.L, cpu.checker, cpu.dim = L, chk, sc.dim
    lb, ub, s0, g0 = (np.ascontiguousarray(x, dtype=np.float64) for x in (sc.lb, sc.ub, sc.start, sc.goal))
    dp = C.POINTER(C.c_double)
    cpu.h = L.ref_solver_new(ref.RefSolver.RRTSTAR, sc.dim, chk.h, lb.ctypes.data_as(dp), ub.ctypes.data_as(dp),
                             s0.ctypes.data_as(dp), g0.ctypes.data_as(dp), sc.max_distance, sc.rewire_factor,
                             sc.utopia_tolerance, 1, 0, 1)
    gpu = ref.DropinSolver(ref.RefSolver.RRTSTAR, sc, sc.make_world(), informed=True, cuda_nn=True)
    assert cpu.run(samples) == gpu.run(samples)
    _same(ref, cpu, gpu)
    assert cpu.solved
    return cpu, gpu, L


def test_batched_path_and_tree_revalidation_counts_device_calls(gc, orc, ref):
    """SURVEY 8f N1 / N2 with the edge checks of a whole path or tree in ONE device call
    (graph_core_b200/host/cuda_batched_helpers.h): Path::isValid, Path::isValidFromConf, Tree::recheckCollision and
    Tree::checkPathToNode keep deciding -- verdicts, connection costs, early returns, recently-checked flags are the
    CPU reference's -- while the number of gc_check_edges calls drops from one per connection to one or two."""
    import ctypes as C
    sc = W.scenario_c1()
    cpu, gpu, L = _solved_pair(gc, orc, ref, sc, 9500, 2500)
    path = cpu.solution()
    n_conn = len(path) - 1
    assert n_conn >= 4
    # N2 Path::isValid: n_conn calls one by one, 1 batched
    c0 = gpu.device_calls
    v_one, c_one = gpu.solution_is_valid()
    c1 = gpu.device_calls
    v_bat, c_bat = gpu.solution_is_valid_batched()
    c2 = gpu.device_calls
    v_cpu, c_cpu = cpu.solution_is_valid()
    assert v_cpu == v_one == v_bat == 1 and np.array_equal(c_cpu, c_one) and np.array_equal(c_cpu, c_bat)
    assert c1 - c0 == n_conn and c2 - c1 == 1
    # N1 Tree::recheckCollision: one call for the whole tree instead of one per connection
    c0 = gpu.device_calls
    assert gpu.recheck_tree() == cpu.recheck_tree() == True  # noqa: E712
    c1 = gpu.device_calls
    assert gpu.recheck_tree_batched()
    c2 = gpu.device_calls
    # (the connections of the path are still in the cache of the batched isValid above)
    assert gpu.tree_size - 1 - n_conn <= c1 - c0 <= gpu.tree_size - 1 and c2 - c1 == 1
    # the same scene with a new obstacle on the path
    sc2 = _box_scene_with_new_obstacle(sc, path)
    ow2 = orc.OracleWorld.from_scenario(sc2)
    cpu_chk2 = ref.RefChecker(L.ref_checker_callback(sc.dim, C.cast(ow2.L.orc_check_states, C.c_void_p), ow2.h,
                                                     sc.min_distance), sc.dim, L, sc.min_distance, keep=(ow2,))
    gpu_chk2 = ref.cuda_ref_checker(sc2.make_world(), sc.min_distance)
    v_cpu, c_cpu = cpu.solution_is_valid(cpu_chk2)
    k0 = gpu.device_calls_of(gpu_chk2)
    v_bat, c_bat = gpu.solution_is_valid_batched(gpu_chk2)
    assert gpu.device_calls_of(gpu_chk2) - k0 == 1
    assert v_cpu == v_bat == 0 and np.array_equal(c_cpu, c_bat) and np.isinf(c_cpu).any()
    for ci in range(n_conn):
        for t in (0.0, 0.41, 1.0):
            if ci == n_conn - 1 and t == 1.0:
                continue  # conf == goal: the reference asserts (path.cpp:1352)
            k0 = gpu.device_calls_of(gpu_chk2)
            got = gpu.solution_is_valid_from_batched(ci, t, gpu_chk2)
            assert got == cpu.solution_is_valid_from(ci, t, cpu_chk2), (ci, t)
            assert gpu.device_calls_of(gpu_chk2) - k0 <= 2  # checkConnFromConf + one batch for the rest of the path
    # N1 / a19 Tree::checkPathToNode towards the goal and towards inner nodes, flags kept between two calls
    cpu.solution_is_valid()
    gpu.solution_is_valid_batched()
    for node in (path[-1], path[len(path) // 2], path[1]):
        r_cpu = cpu.check_path_to_node(node, keep_flags=False)
        r_one = gpu.check_path_to_node(node, keep_flags=False)
        k0 = gpu.device_calls
        r_bat = gpu.check_path_to_node_batched(node, keep_flags=False)
        assert gpu.device_calls - k0 == 1
        for r in (r_one, r_bat):
            assert r[0] == r_cpu[0] == 1 and np.array_equal(r[1], r_cpu[1]) and r[2] == r_cpu[2]
    # with the new obstacle: stops at the first blocked connection, marks it infinite, remembers what it checked
    r_cpu = cpu.check_path_to_node(path[-1], cpu_chk2, keep_flags=True)
    r_bat = gpu.check_path_to_node_batched(path[-1], gpu_chk2, keep_flags=True)
    assert r_cpu[0] == r_bat[0] == 0 and np.array_equal(r_cpu[1], r_bat[1]) and r_cpu[2] == r_bat[2] >= 1
    assert np.isinf(r_cpu[1]).sum() == 1
    # second call: the recently-checked connections are not checked again (tree.cpp:345-352)
    r_cpu2 = cpu.check_path_to_node(path[-1], cpu_chk2, keep_flags=True)
    k0 = gpu.device_calls_of(gpu_chk2)
    r_bat2 = gpu.check_path_to_node_batched(path[-1], gpu_chk2, keep_flags=True)
    assert r_cpu2[0] == r_bat2[0] == 0 and r_cpu2[2] == r_bat2[2] == 0
    assert gpu.device_calls_of(gpu_chk2) == k0


def test_speculative_local_optimizer_counts_device_calls(gc, orc, ref):
    """PathLocalOptimizer::bisection tries up to five dependent mid points (2 edge checks each,
    path_local_optimizer.cpp:80-99); SpeculativePathLocalOptimizer checks the 31 possible mid points of one bisection in a
    single device call and then lets the reference's own bisection decide from the cache."""
    sc = W.scenario_c1()
    cpu, gpu, L = _solved_pair(gc, orc, ref, sc, 9600, 2500)
    for mode, iters in ((0, 4), (1, 3), (2, 30)):
        wc, cc = cpu.optimize_solution(mode, iters)
        k0 = gpu.device_calls
        wg, cg = gpu.optimize_solution(mode, iters)
        k1 = gpu.device_calls
        ws, cs = gpu.optimize_solution_speculative(mode, iters)
        k2 = gpu.device_calls
        assert np.array_equal(wc, wg) and cc == cg, mode
        assert np.array_equal(wc, ws) and cc == cs, mode
        assert k2 - k1 <= k1 - k0, (mode, k1 - k0, k2 - k1)
        if mode == 0:
            # warp: at most one device call per bisection instead of up to ten
            assert k2 - k1 <= iters * (len(cpu.solution()) - 2)
            assert (k1 - k0) >= 2 * (k2 - k1)


def test_rewire_only_with_path_check_batched(gc, orc, ref):
    """Tree::rewireOnlyWithPathCheck (tree.cpp:522-679), the replanning form of rewireOnly that re-validates
    the connections it relies on.  cuda_batched::rewireOnlyWithPathCheck hands the near set's edges (both directions) and
    every unflagged connection on the way to those nodes to the device as ONE batch and then runs the reference's own
    function: the boolean, the connections it flagged and the rewired tree (infinite costs of the blocked connections
    included) must be those of the CPU reference with its per-state checker on the changed scene."""
    import ctypes as C
    sc = W.scenario_c1()
    cpu, gpu, L = _solved_pair(gc, orc, ref, sc, 9500, 2500)
    path = cpu.solution()
    sc2 = _box_scene_with_new_obstacle(sc, path)
    ow2 = orc.OracleWorld.from_scenario(sc2)
    cpu_chk2 = ref.RefChecker(L.ref_checker_callback(sc.dim, C.cast(ow2.L.orc_check_states, C.c_void_p), ow2.h,
                                                     sc.min_distance), sc.dim, L, sc.min_distance, keep=(ow2,))
    gpu_chk2 = ref.cuda_ref_checker(sc2.make_world(), sc.min_distance)
    n_improved = n_flagged = 0
    calls = []
    for node, what, r in ((path[-1], 0, 0.15), (path[len(path) // 2 + 1], 0, 0.15), (path[2], 2, 0.15),
                          (path[len(path) // 2 - 1], 1, 0.2), (path[-2], 0, 0.0)):
        want = cpu.rewire_only_with_path_check(node, r, cpu_chk2, what_rewire=what, keep_flags=True)
        k0 = gpu.device_calls_of(gpu_chk2)
        got = gpu.rewire_only_with_path_check_batched(node, r, gpu_chk2, what_rewire=what, keep_flags=True)
        calls.append(gpu.device_calls_of(gpu_chk2) - k0)
        assert got == want, (what, r, got, want)
        _same(ref, cpu, gpu)
        n_improved += want[0]
        n_flagged += want[1]
    assert n_improved >= 2 and n_flagged >= 50
    # one device call per rewiring (the prefetch); nothing the reference asked for was outside the batch
    assert max(calls) <= 1, calls
    conf, par, pc = gpu.dump()
    assert np.isinf(pc).any()  # the connections under the new obstacle were found and marked
