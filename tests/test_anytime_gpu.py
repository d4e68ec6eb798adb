"""BASELINE config 5 as a workload: the lock-step AnytimeRRT (graph_core_b200/csrc/anytime.cu) against the reference's OWN
AnytimeRRT (oracle/_ref, compiled from /root/reference: anytime_rrt.cpp, rrt.cpp, tree.cpp informedExtend) driven through
its public update / improve / improveUpdate with the same per-problem Philox sample stream, and against the oracle's
restatement: same rounds, iterations per round, tree sizes, bias / cost2beat sequence, path costs and way points, bit for
bit."""
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


def _reference_run(orc, ref, sc, start, goal, cost0, max_iter, seed, p, P, extend, improve_in_solve=True):
    rc = ref.RefChecker.over_oracle_world(orc.OracleWorld.from_scenario(sc), sc.min_distance)
    r = ref.RefSolver(ref.RefSolver.ANYTIME, sc, rc, start=start, goal=goal, extend=extend, informed=True)
    lb, ub = np.asarray(sc.lb, float), np.asarray(sc.ub, float)

    def sample_fn(k, cost):
        c = cost0 if cost is None else cost
        q, _ = orc.sample_informed(lb, ub, start[None, :], goal[None, :], np.array([c]), seed, k * P + p)
        return q[0]

    solved, log = ref.anytime_solve(r, sample_fn, max_iter, improve_in_solve=improve_in_solve,
                                    utopia_tolerance=1.0 + sc.utopia_tolerance,
                                    best_utopia=float(np.sqrt(((goal - start) ** 2).sum())))
    return r, solved, log


def _free_pairs(sc, orc, n, seed):
    ow = orc.OracleWorld.from_scenario(sc)
    starts, goals = [np.asarray(sc.start, float)], [np.asarray(sc.goal, float)]
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


def _compare(lk, p, r, solved, log):
    info = lk.info(p)
    assert info["done"], p
    assert info["solved"] == solved, p
    got = lk.log(p)
    assert got.shape == log.shape, (p, got, log)
    assert np.array_equal(got, log), (p, got, log)
    if solved:
        assert info["cost"] == r.cost, p
        wp, cs = lk.solution(p)
        assert np.array_equal(wp, r.solution()), p
        assert float(np.sum(cs)) == pytest.approx(r.cost, rel=1e-12)


@pytest.mark.parametrize("extend", [False, True])
def test_lockstep_anytime_matches_the_reference_c1(gc, orc, ref, extend):
    """2-D boxes (k-nearest really cuts the candidate set, several successful improve rounds, the utopia exits)."""
    from graph_core_b200.planner import LockstepAnytimeRRT
    sc = W.scenario_c1()
    P, max_iter, seed = 6, 400, 4100
    starts, goals = _free_pairs(sc, orc, P, 4200)
    lk = LockstepAnytimeRRT(sc, sc.make_world(), starts, goals, max_iter=max_iter, seed=seed, extend=extend,
                            capacity=4096, keep_trees=True)
    assert lk.run() == 0
    n_solved = n_improved = 0
    for p in range(P):
        r, solved, log = _reference_run(orc, ref, sc, starts[p], goals[p], float("inf"), max_iter, seed, p, P, extend)
        _compare(lk, p, r, solved, log)
        n_solved += int(solved)
        n_improved += int(log[log[:, 0] == 1][:, 2].sum()) if log.size else 0
    assert n_solved >= 3 and n_improved >= 2
    st = lk.stats()
    assert st["improve_rows"] > 0 and st["rrt_rows"] > 0 and st["improve_edges"] > 0


def test_lockstep_anytime_solve_as_written(gc, orc, ref):
    """improve_in_solve = False: AnytimeRRT::solve literally (rrt.cpp:150 makes its improve loop a no-op)."""
    from graph_core_b200.planner import LockstepAnytimeRRT
    sc = W.scenario_c1()
    P, max_iter, seed = 4, 300, 4300
    starts, goals = _free_pairs(sc, orc, P, 4400)
    lk = LockstepAnytimeRRT(sc, sc.make_world(), starts, goals, max_iter=max_iter, seed=seed, improve_in_solve=False,
                            capacity=4096)
    assert lk.run() == 0
    for p in range(P):
        r, solved, log = _reference_run(orc, ref, sc, starts[p], goals[p], float("inf"), max_iter, seed, p, P, False,
                                        improve_in_solve=False)
        _compare(lk, p, r, solved, log)
        assert (log[:, 0] == 0).all()
    assert lk.stats()["improve_rows"] == 0


def test_lockstep_anytime_matches_the_reference_c5(gc, orc, ref, c3):
    """C5' queries on the 7-DoF sphere-arm world, informed first phase; also against the oracle's own solve()."""
    from graph_core_b200.planner import LockstepAnytimeRRT
    arm = orc.arm_checker()
    P, max_iter, seed = 6, 300, 501
    starts, goals, cost0 = W.c5_problems(arm, c3.params, P)
    lk = LockstepAnytimeRRT(c3, c3.make_world(), starts, goals, max_iter=max_iter, seed=seed, sampler_cost0=cost0,
                            capacity=8192, keep_trees=True)
    assert lk.run() == 0
    n_solved = 0
    for p in range(P):
        r, solved, log = _reference_run(orc, ref, c3, starts[p], goals[p], float(cost0[p]), max_iter, seed, p, P, False)
        _compare(lk, p, r, solved, log)
        o = orc.OracleAnytime(c3, orc.OracleWorld.from_scenario(c3), start=starts[p], goal=goals[p],
                              sampler_cost0=float(cost0[p]))
        assert o.solve(max_iter, seed, p, P) == solved
        assert np.array_equal(o.log(), lk.log(p))
        if solved:
            n_solved += 1
            assert np.array_equal(o.solution(), lk.solution(p)[0])
            # the tree that holds the solution: root first, goal last, every parent link the oracle's
            par, pc, cf = lk.dump(p, 1)
            assert par[0] == -1 and np.array_equal(cf[0], starts[p]) and np.array_equal(cf[-1], goals[p])
            opar, opc, oconf = o.dump()
            # the oracle's pool holds the nodes of every round; the goal configuration occurs once per round
            want = {}
            for k in range(len(opar)):
                if opar[k] >= 0:
                    want.setdefault(oconf[k].tobytes(), []).append((oconf[opar[k]].tobytes(), float(opc[k])))
            for k in range(1, len(par)):
                assert (cf[par[k]].tobytes(), float(pc[k])) in want[cf[k].tobytes()], (p, k)
    assert n_solved >= 2


def test_lockstep_anytime_step_by_step_equals_run(gc, orc):
    """gc_anytime_step in a caller's loop and gc_anytime_run are the same computation."""
    from graph_core_b200.planner import LockstepAnytimeRRT
    sc = W.scenario_c1()
    starts, goals = _free_pairs(sc, orc, 3, 4500)
    full = LockstepAnytimeRRT(sc, sc.make_world(), starts, goals, max_iter=200, seed=77, capacity=4096)
    assert full.run() == 0
    again = LockstepAnytimeRRT(sc, sc.make_world(), starts, goals, max_iter=200, seed=77, capacity=4096)
    live = 3
    while live:
        live = again.step()
    for p in range(3):
        assert np.array_equal(full.log(p), again.log(p))
        assert full.info(p) == again.info(p)


def test_lockstep_anytime_results_do_not_depend_on_sharding(gc, orc):
    """BASELINE config 5 shards the queries over the GPUs (query j on GPU j mod G): with gc_anytime_set_streams a query
    draws the same samples in a shard as in the full set, so every shard reproduces the full run."""
    from graph_core_b200.planner import LockstepAnytimeRRT
    sc = W.scenario_c1()
    P = 6
    starts, goals = _free_pairs(sc, orc, P, 4600)
    full = LockstepAnytimeRRT(sc, sc.make_world(), starts, goals, max_iter=250, seed=91, capacity=4096)
    assert full.run() == 0
    for shard in (np.arange(0, P, 2), np.arange(1, P, 2)):
        part = LockstepAnytimeRRT(sc, sc.make_world(), starts[shard], goals[shard], max_iter=250, seed=91, capacity=4096)
        part.set_streams(P, shard)
        assert part.run() == 0
        for i, j in enumerate(shard):
            assert np.array_equal(part.log(i), full.log(int(j)))
            a, b = part.info(i), full.info(int(j))
            assert a == b
            if a["solved"]:
                assert np.array_equal(part.solution(i)[0], full.solution(int(j))[0])


def test_lockstep_anytime_reports_a_full_tree(gc, orc):
    """A tree that outgrows its segment is an error (GC_E_CAPACITY with the problem named), never a silent truncation."""
    from graph_core_b200.planner import LockstepAnytimeRRT
    sc = W.scenario_c1()
    starts, goals = _free_pairs(sc, orc, 2, 4700)
    # the scenario's own pair needs a few hundred nodes; 32 slots are exhausted during the first round
    lk = LockstepAnytimeRRT(sc, sc.make_world(), starts[:1], goals[:1], max_iter=2000, seed=5, capacity=32)
    with pytest.raises(gc.GcError) as ei:
        lk.run()
    assert "capacity" in str(ei.value)
