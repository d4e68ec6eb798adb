"""GPU parity at planner level: lock-step CUDA RRT* vs the oracle's restatement of RRTStar::update
(src/graph_core/solvers/rrt_star.cpp:89-163) driven by the same injected samples
(tree_solver.h:338).  Every problem must build the SAME tree: node configurations and
parent indices bit-exact, connection costs and path costs bit-exact (BASELINE asks for 1e-5 rel)."""
import numpy as np
import pytest

from graph_core_b200 import worlds as W

pytestmark = pytest.mark.gpu


def _run_both(gc, orc, sc, starts, goals, iters, seed, use_kdtree=True, threads=4, device_trees=True):
    from graph_core_b200.planner import LockstepRRTStar
    P = len(starts)
    gw = sc.make_world()
    lock = LockstepRRTStar(sc, gw, starts, goals, capacity=max(4096, iters + 64), threads=threads,
                           device_trees=device_trees)
    samples = np.stack([W.uniform_samples(sc.lb, sc.ub, seed + p, 0, iters) for p in range(P)], axis=1)
    lock.run(samples)
    ow = orc.OracleWorld.from_scenario(sc)
    for p in range(P):
        o = orc.OracleSolver(orc.OracleSolver.RRTSTAR, sc, ow, start=starts[p], goal=goals[p], use_kdtree=use_kdtree)
        o.run(samples[:, p, :])
        par, pc, conf = o.dump()
        gpar, gpc, gconf = lock.dump(p)
        info = lock.info(p)
        assert info["num_nodes"] == len(par), (p, info, len(par))
        assert np.array_equal(gconf, conf)
        assert np.array_equal(gpar, par)
        assert np.array_equal(gpc, pc)
        assert info["solved"] == o.solved and info["tree_size"] == o.tree_size
        assert info["cost"] == o.cost or (np.isinf(info["cost"]) and np.isinf(o.cost))
        assert np.array_equal(lock.solution(p), o.solution())
    st = lock.stats()
    assert st["edges_fallback"] == 0          # the speculative edge set always covered the replay
    assert st["edges_used"] <= st["edges_device"] + st["iterations"]
    return lock, st


TREES = pytest.mark.parametrize("device_trees", [True, False], ids=["device_trees", "host_replay"])


@TREES
def test_rrtstar_reference_scene_c0(gc, orc, device_trees):
    sc = W.scenario_c0()
    lock, st = _run_both(gc, orc, sc, sc.start[None, :], sc.goal[None, :], 1500, seed=1, device_trees=device_trees)
    assert lock.info(0)["solved"]


@TREES
def test_rrtstar_c1_boxes_many_problems(gc, orc, device_trees):
    sc = W.scenario_c1()
    ow = orc.OracleWorld.from_scenario(sc)
    cand = W.uniform_samples(sc.lb, sc.ub, 77, 0, 64)
    cand = cand[ow.check(cand).astype(bool)]
    starts, goals = cand[:6], cand[6:12]
    lock, st = _run_both(gc, orc, sc, starts, goals, 1200, seed=300, use_kdtree=False, device_trees=device_trees)
    assert sum(lock.info(p)["solved"] for p in range(6)) >= 4


@TREES
def test_rrtstar_c3_arm(gc, orc, c3, device_trees):
    from graph_core_b200 import worlds
    pairs = worlds.c3_problems(orc.arm_checker(), c3.params, 3)
    starts = np.array([s for s, _ in pairs])
    goals = np.array([g for _, g in pairs])
    lock, st = _run_both(gc, orc, c3, starts, goals, 700, seed=3020, device_trees=device_trees)
    assert st["extended"] > 100


def test_rrtstar_direct_connection_and_thread_counts(gc, orc):
    # straight line start->goal is free: setProblem's connectToNode solves it without sampling
    sc = W.scenario_c0()
    starts = np.array([[1.5, 1.5, 1.5], [-1.5, -1.5, -1.5]])
    goals = np.array([[2.2, 1.2, 2.0], [1.5, 1.5, 1.5]])
    for threads, device_trees in ((1, False), (3, False), (1, True)):
        lock, st = _run_both(gc, orc, sc, starts, goals, 300, seed=9, threads=threads, device_trees=device_trees)
        assert lock.info(0)["solved"] and not lock.info(0)["can_improve"]


def test_rrtstar_informed_sampler_inside_the_planner(gc, orc):
    """RRTStar::update with the solver's own sampler (rrt_star.cpp:147-161: sampler_->setCost after every improvement):
    the device forest draws sample i of problem p from the informed set of p's CURRENT path cost (Philox counter
    i * P + p).  The oracle is fed the same stream -- gc_sample_informed with the oracle's own cost before every
    iteration -- so equal trees prove that the device used the right cost at every iteration."""
    from graph_core_b200 import capi
    from graph_core_b200.planner import LockstepRRTStar
    sc = W.scenario_c0()
    starts = np.array([sc.start, [-1.5, -1.5, -1.5], [1.6, -1.4, 1.5]])
    goals = np.array([sc.goal, [1.5, 1.5, 1.5], [-1.5, 1.6, -1.4]])
    P, iters, seed = len(starts), 500, 4242
    lock = LockstepRRTStar(sc, sc.make_world(), starts, goals, capacity=1024)
    for i in range(iters):
        lock.step_informed(seed, i * P)
    lock.sync()
    ow = orc.OracleWorld.from_scenario(sc)
    solvers = [orc.OracleSolver(orc.OracleSolver.RRTSTAR, sc, ow, start=starts[p], goal=goals[p]) for p in range(P)]
    informed_draws = 0
    for i in range(iters):
        cost = np.array([o.cost if o.solved else np.inf for o in solvers])
        informed_draws += int(np.isfinite(cost).sum())
        smp, _ = capi.sample_informed(sc.lb, sc.ub, starts, goals, cost, seed, i * P)
        for p, o in enumerate(solvers):
            o.run(smp[p][None, :])
    assert informed_draws > 100  # the informed branch was exercised, not only the uniform one
    for p, o in enumerate(solvers):
        par, pc, conf = o.dump()
        gpar, gpc, gconf = lock.dump(p)
        assert np.array_equal(gconf, conf) and np.array_equal(gpar, par) and np.array_equal(gpc, pc), p
        assert lock.info(p)["cost"] == o.cost or (np.isinf(lock.info(p)["cost"]) and np.isinf(o.cost))
        assert np.array_equal(lock.solution(p), o.solution())


def test_rrtstar_deep_trees_c1(gc, orc):
    """BASELINE config 1 depth: 10 000 iterations per problem (trees had only been compared at <= 1 500)."""
    sc = W.scenario_c1()
    ow = orc.OracleWorld.from_scenario(sc)
    cand = W.uniform_samples(sc.lb, sc.ub, 78, 0, 64)
    cand = cand[ow.check(cand).astype(bool)]
    lock, st = _run_both(gc, orc, sc, cand[0:2], cand[8:10], 10000, seed=910, use_kdtree=True)
    assert max(lock.info(p)["tree_size"] for p in range(2)) > 3000


def test_rrtstar_deep_trees_c3(gc, orc, c3):
    """BASELINE config 3 depth: 10 000 iterations of the 7-DoF arm problem."""
    from graph_core_b200 import worlds
    pairs = worlds.c3_problems(orc.arm_checker(), c3.params, 2)
    starts = np.array([s for s, _ in pairs])
    goals = np.array([g for _, g in pairs])
    lock, st = _run_both(gc, orc, c3, starts, goals, 10000, seed=3090)
    assert st["extended"] > 2000
