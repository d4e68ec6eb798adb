"""BASELINE config 2 as a workload: the device-resident lock-step BiRRT / RRT (graph_core_b200/csrc/birrt.cu, one kernel
launch per update with the whole Tree::connect ray inside) against the reference's OWN BiRRT and RRT classes
(oracle/_ref, compiled from /root/reference: birrt.cpp:93-152, rrt.cpp:119-157, tree.cpp:217-233) fed the same
injected samples: same meeting iteration, bit-identical path, cost and tree."""
import numpy as np
import pytest

from graph_core_b200 import worlds as W

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    import ref_py
    try:
        ref_py.lib()
    except (FileNotFoundError, OSError) as ex:
        pytest.skip("oracle/_ref/libgc_ref.so unavailable: %s" % ex)
    return ref_py


def _cpu_checker(ref, orc, sc):
    if sc.kind == "cube":
        return ref.RefChecker.cube(sc.dim, sc.params["threshold"], sc.min_distance)
    return ref.RefChecker.over_oracle_world(orc.OracleWorld.from_scenario(sc), sc.min_distance)


def _pairs(sc, orc, n, seed):
    """n start/goal pairs with collision-free end points (the scenario's own pair first)."""
    ow = orc.OracleWorld.from_scenario(sc)
    starts, goals = [sc.start], [sc.goal]
    first = 0
    while len(starts) < n:
        u = W.philox_u01(seed, first, 64, 2 * sc.dim)
        first += 64
        s = sc.lb + u[:, :sc.dim] * (sc.ub - sc.lb)
        g = sc.lb + u[:, sc.dim:] * (sc.ub - sc.lb)
        ok = ow.check(s).astype(bool) & ow.check(g).astype(bool)
        for i in np.nonzero(ok)[0]:
            if len(starts) < n:
                starts.append(s[i])
                goals.append(g[i])
    return np.array(starts), np.array(goals)


def _device_tree_edges(lk, p, solved):
    """The start tree of problem p as {child conf: (parent conf | None, cost)}; after a solution BiRRT::update has moved
    the goal branch into the start tree (birrt.cpp:128-141): added from the way points."""
    par, pc, cf = lk.dump(p, 0)
    out = {}
    for i in range(len(par)):
        out[cf[i].tobytes()] = (None if par[i] < 0 else cf[par[i]].tobytes(), float(pc[i]) if par[i] >= 0 else 0.0)
    if solved:
        wp, cs = lk.solution(p)
        for k in range(len(wp) - 1):
            key = wp[k + 1].tobytes()
            if key not in out:
                out[key] = (wp[k].tobytes(), float(cs[k]))
    return out


def _ref_tree_edges(ref, solver):
    conf, parent_conf, pcost = solver.dump()
    out = {}
    for i in range(conf.shape[0]):
        par = None if np.isnan(parent_conf[i]).any() else parent_conf[i].tobytes()
        out[conf[i].tobytes()] = (par, float(pcost[i]) if par is not None else 0.0)
    return out


@pytest.mark.parametrize("name,n_prob,iters,extend", [("C2", 6, 400, False), ("C1", 5, 600, False), ("C0", 3, 400, False),
                                                     ("C2", 4, 3000, True)])
def test_lockstep_birrt_matches_the_reference_birrt(gc, orc, ref, name, n_prob, iters, extend):
    from graph_core_b200.planner import LockstepBiRRT
    sc = {"C0": W.scenario_c0, "C1": W.scenario_c1, "C2": W.scenario_c2}[name]()
    starts, goals = _pairs(sc, orc, n_prob, 7200 + sc.dim)
    samples = W.uniform_samples(sc.lb, sc.ub, 7300 + sc.dim, 0, iters * n_prob).reshape(iters, n_prob, sc.dim)
    lk = LockstepBiRRT(sc, sc.make_world(), starts, goals, capacity=8192, bidirectional=True, extend=extend)
    for i in range(iters):
        lk.step(samples[i])
    lk.sync()
    n_solved = 0
    for p in range(n_prob):
        cpu = ref.RefSolver(ref.RefSolver.BIRRT, sc, _cpu_checker(ref, orc, sc), start=starts[p], goal=goals[p],
                            extend=extend, informed=False)
        used = cpu.run(samples[:, p])
        info = lk.info(p)
        assert info["solved"] == cpu.solved, p
        if cpu.solved:
            n_solved += 1
            if info["solved_iteration"] > 0:  # 0: solved by setProblem's direct connection (tree_solver.cpp:84-98)
                assert info["solved_iteration"] == used, p
            assert info["cost"] == cpu.cost, p
            wp, cs = lk.solution(p)
            assert np.array_equal(wp, cpu.solution()), p
        assert _device_tree_edges(lk, p, cpu.solved) == _ref_tree_edges(ref, cpu), p
    assert n_solved >= 1
    st = lk.stats()
    assert st["rescans"] == 0 and st["scans"] > 0 and st["edges"] > 0


@pytest.mark.parametrize("name,extend,iters", [("C1", False, 2500), ("C1", True, 4000), ("C0", False, 3000)])
def test_lockstep_rrt_matches_the_reference_rrt(gc, orc, ref, name, extend, iters):
    from graph_core_b200.planner import LockstepBiRRT
    sc = {"C0": W.scenario_c0, "C1": W.scenario_c1}[name]()
    n_prob = 4
    starts, goals = _pairs(sc, orc, n_prob, 7400 + sc.dim)
    samples = W.uniform_samples(sc.lb, sc.ub, 7500 + sc.dim, 0, iters * n_prob).reshape(iters, n_prob, sc.dim)
    lk = LockstepBiRRT(sc, sc.make_world(), starts, goals, capacity=16384, bidirectional=False, extend=extend)
    for i in range(iters):
        lk.step(samples[i])
    lk.sync()
    n_solved = 0
    for p in range(n_prob):
        cpu = ref.RefSolver(ref.RefSolver.RRT, sc, _cpu_checker(ref, orc, sc), start=starts[p], goal=goals[p],
                            extend=extend, informed=False)
        used = cpu.run(samples[:, p])
        info = lk.info(p)
        assert info["solved"] == cpu.solved, p
        if cpu.solved:
            n_solved += 1
            assert info["solved_iteration"] in (0, used) and info["cost"] == cpu.cost, p
            wp, cs = lk.solution(p)
            assert np.array_equal(wp, cpu.solution()), p
        assert info["size_start"] == cpu.tree_size, p
        par, pc, cf = lk.dump(p, 0)
        mine = {cf[i].tobytes(): (None if par[i] < 0 else cf[par[i]].tobytes(), float(pc[i]) if par[i] >= 0 else 0.0)
                for i in range(len(par))}
        assert mine == _ref_tree_edges(ref, cpu), p
    assert n_solved >= 1


def test_lockstep_birrt_reports_capacity_overflow(gc, orc):
    from graph_core_b200 import capi
    from graph_core_b200.planner import LockstepBiRRT
    sc = W.scenario_c1()
    samples = W.uniform_samples(sc.lb, sc.ub, 7600, 0, 200).reshape(200, 1, sc.dim)
    lk = LockstepBiRRT(sc, sc.make_world(), sc.start[None], sc.goal[None], capacity=8, bidirectional=True)
    for i in range(200):
        lk.step(samples[i])
    with pytest.raises(capi.GcError):
        lk.sync()
