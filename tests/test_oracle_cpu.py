"""CPU tests of the oracle itself (oracle/gc_oracle.cpp).

The reference's own tests pin no numerical result on either hot path, so the
restatement is pinned by (a) the reference's single fixture tests/path.yaml, (b) golden vectors
minted from the restatement and committed under tests/golden/, (c) internal consistency
KdTree == Vector (two independent restatements of kdtree.cpp / vector.cpp must agree except on
exact ties) and (d) where available the reference's own sources compiled against stand-in headers
(tests/test_oracle_vs_ref.py).
"""
import json
from pathlib import Path

import numpy as np
import pytest

from graph_core_b200 import worlds as W

GOLD = Path(__file__).parent / "golden"


@pytest.mark.parametrize("dim,n", [(2, 500), (3, 2000), (7, 4000), (12, 3000)])
def test_kdtree_equals_vector(orc, dim, n):
    pts = W.c4_nodes(n, dim, seed=100 + dim)
    qs = W.c4_queries(40, dim, seed=200 + dim)
    v, k = orc.OracleNN("vector", dim), orc.OracleNN("kdtree", dim)
    v.insert(pts)
    k.insert(pts)
    assert v.size() == k.size() == n
    iv, dv = v.nearest(qs)
    ik, dk = k.nearest(qs)
    assert np.array_equal(iv, ik) and np.array_equal(dv, dk)
    brute = np.sqrt(((pts[None, :, :] - qs[:, None, :]) ** 2).sum(-1))
    assert np.array_equal(iv, brute.argmin(1))
    r = W.c4_radius(25, n, dim)
    a, b = v.near(qs, r, 512), k.near(qs, r, 512)
    assert np.array_equal(a[2], b[2])
    for i in range(40):
        c = a[2][i]
        assert np.array_equal(a[0][i, :c], b[0][i, :c]) and np.array_equal(a[1][i, :c], b[1][i, :c])
        assert c == (brute[i] < r).sum()
        assert np.all(np.diff(a[1][i, :c]) >= 0)
    for kk in (1, 7, n, n + 5):
        a, b = v.knn(qs, kk, n), k.knn(qs, kk, n)
        assert np.array_equal(a[2], b[2]) and (a[2] == min(kk, n)).all()
        assert np.array_equal(a[0], b[0])


def test_kdtree_lazy_deletion_semantics(orc):
    """kdtree.cpp:261-278, 395-456: tombstones, size bookkeeping, rebuild threshold, restore quirk."""
    pts = W.c4_nodes(300, 3, seed=7)
    k = orc.OracleNN("kdtree", 3)
    v = orc.OracleNN("vector", 3)
    k.insert(pts)
    v.insert(pts)
    for i in (0, 5, 17):  # includes the root
        assert k.delete(i) and v.delete(i)
    assert k.size() == v.size() == 297 and k.deleted() == 3
    assert not k.find(5) and not v.find(5)
    # KdTree::deleteNode returns false when findNode fails (kdtree.cpp:397-399); a tombstone is "not found"
    assert not k.delete(5)
    qs = W.c4_queries(30, 3, seed=8)
    assert np.array_equal(k.nearest(qs)[0], v.nearest(qs)[0])
    assert not k.restore(5)       # restoreNode goes through findNode, which skips tombstones (:447-456)
    k.reinsert(5)                 # KdNode::insert of an existing pointer un-deletes it (:93-97)
    assert k.find(5) and k.size() == 298 and k.deleted() == 2
    assert sorted(k.get_nodes().tolist()) == sorted(set(range(300)) - {0, 17})
    k.set_deleted_threshold(2)
    assert k.delete(40)           # deleted_nodes_ (3) > threshold (2): rebuilt from scratch
    assert k.size() == 297
    assert np.array_equal(np.sort(k.get_nodes()), np.sort(np.array(sorted(set(range(300)) - {0, 17, 40}))))
    v.delete(40)
    v2 = orc.OracleNN("vector", 3)
    keep = np.array(sorted(set(range(300)) - {0, 17, 40}))
    v2.insert(pts[keep])
    assert np.array_equal(keep[v2.nearest(qs)[0]], k.nearest(qs)[0])


def test_delete_returns(orc):
    k = orc.OracleNN("kdtree", 2)
    k.insert(np.array([[0.0, 0.0], [1.0, 1.0]]))
    assert k.delete(1)
    # second delete: findNode fails on the tombstone -> false (kdtree.cpp:397-399)
    assert not k.delete(1)


def test_bisection_state_counts(orc):
    """F6: interior states = n_last - 1 with n_last the largest 2^p (p>=1) such that dist > 2^p * min_d."""
    w = orc.OracleWorld.cube(3, 0.0)  # max|q| > 0: everything except the origin is free
    md = 0.01
    for length, interior in [(0.005, None), (0.01, 0), (0.02, 0), (0.0200001, 1), (0.25, 15), (1.0, 63),
                             (2.29, 127), (0.32, 15), (0.3200001, 31)]:
        a = np.array([[0.0, 1.0, 1.0]])  # x from 0 so that c2 - c1 is exactly `length`
        b = a + np.array([[length, 0.0, 0.0]])
        w.reset_counters()
        assert w.check_connection(a, b, md)[0] == 1
        states = w.counters()["states"]
        assert states == (2 if interior is None else 2 + interior), (length, states)


def test_linear_variant_first_failure_and_conf(orc):
    """collision_checker_base.h:178-205: npnt = ceil(dist/min_d), conf = last feasible configuration."""
    w = orc.OracleWorld.cube(2, 1.0)  # collision inside the square |q|inf <= 1
    a = np.array([[1.5, 0.0]])
    b = np.array([[0.5, 0.0]])
    ok, fb, conf = w.check_connection(a, b, 0.01, mode=1)
    assert ok[0] == 0
    npnt = int(np.ceil(1.0 / 0.01))
    step = (b[0] - a[0]) / npnt
    # first colliding interior point: 1.5 + step*i <= 1.0
    i = next(i for i in range(1, npnt) if max(abs(a[0] + step * i)) <= 1.0)
    assert fb[0] == i
    assert np.array_equal(conf[0], a[0] + step * (i - 1))
    ok, fb, conf = w.check_connection(b, a, 0.01, mode=1)  # starts in collision: conf untouched
    assert ok[0] == 0 and fb[0] == 0 and np.isnan(conf[0]).all()
    ok, fb, conf = w.check_connection(a, a + [[0.3, 0.0]], 0.01, mode=1)
    assert ok[0] == 1 and fb[0] == -1


def test_path_yaml_fixture(orc):
    """The reference's fixture: a valid RRT* path of its own test scene (tests/path.yaml)."""
    wp = np.array(json.loads((GOLD / "path_yaml.json").read_text())["waypoints"])
    assert wp.shape == (14, 3)
    assert np.array_equal(wp[0], [-1.5] * 3) and np.array_equal(wp[-1], [1.5] * 3)
    w = orc.OracleWorld.cube(3, 1.0)
    assert w.check(wp).all()
    assert w.check_connection(wp[:-1], wp[1:], 0.01).all()
    # the straight line start->goal crosses the cube: must be rejected
    assert w.check_connection(wp[:1], wp[-1:], 0.01)[0] == 0
    seg_len = np.linalg.norm(np.diff(wp, axis=0), axis=1)
    assert seg_len.sum() > np.linalg.norm(wp[-1] - wp[0])


def test_golden_vectors(orc):
    """Vectors minted from the restatement by tests/golden/make_golden.py guard against drift."""
    g = json.loads((GOLD / "oracle_golden.json").read_text())
    from make_golden import compute  # noqa: E402  (tests/golden is put on sys.path below)
    now = compute(orc)
    assert now == g


def test_ref_golden_vectors(orc):
    """tests/golden/ref_golden.json holds outputs of the REFERENCE ITSELF (oracle/_ref, minted by
    tests/golden/make_ref_golden.py): NN ids/distances, edge booleans, last feasible configurations and complete
    RRT* trees.  The restatement must reproduce every one of them bit for bit."""
    import make_ref_golden as m
    want = json.loads((GOLD / "ref_golden.json").read_text())
    assert "reference sources" in want.pop("_about")
    got = m.compute(m.OracleBackend(None, orc))
    assert got == want


def test_rrtstar_oracle_solves_reference_scene(orc):
    """compute_path_test.cpp:51-132 with injected samples: RRT* finds and improves a path in C0."""
    sc = W.scenario_c0()
    w = orc.OracleWorld.from_scenario(sc)
    for use_kd in (True, False):
        s = orc.OracleSolver(orc.OracleSolver.RRTSTAR, sc, w, use_kdtree=use_kd)
        samples = W.uniform_samples(sc.lb, sc.ub, 1, 0, 3000)
        used = s.run(samples)
        assert s.solved and used == 3000
        ids = s.solution()
        par, pc, conf = s.dump()
        assert ids[0] == 0 and ids[-1] == 1  # start node id 0, goal id 1
        # the solution is a chain of parent links whose edges are collision free
        for a, b in zip(ids[:-1], ids[1:]):
            assert par[b] == a
        assert w.check_connection(conf[ids[:-1]], conf[ids[1:]], sc.min_distance).all()
        cost = sum(pc[i] for i in ids[1:])
        assert abs(cost - s.cost) < 1e-9
        assert s.cost >= np.linalg.norm(sc.goal - sc.start)
        if use_kd:
            ref = (s.cost, s.tree_size, par.copy())
        else:
            # KdTree and Vector drive the planner identically (no exact ties in random data)
            assert s.cost == ref[0] and s.tree_size == ref[1] and np.array_equal(par, ref[2])


def test_rrt_connect_mode(orc):
    sc = W.scenario_c1()
    w = orc.OracleWorld.from_scenario(sc)
    s = orc.OracleSolver(orc.OracleSolver.RRT, sc, w)
    samples = W.uniform_samples(sc.lb, sc.ub, 102, 0, 5000)
    used = s.run(samples)
    assert s.solved and used < 5000 and not s.can_improve
    ids = s.solution()
    par, pc, conf = s.dump()
    assert w.check_connection(conf[ids[:-1]], conf[ids[1:]], sc.min_distance).all()


def test_rewire_radius_formula(orc):
    """rrt_star.cpp:278-286 and sampler_base.h:144-149."""
    import math
    sc = W.scenario_c0()
    w = orc.OracleWorld.from_scenario(sc)
    s = orc.OracleSolver(orc.OracleSolver.RRTSTAR, sc, w)
    d = 3
    spec = math.gamma(d / 2 + 1) / math.pi ** (d / 2) * 5.0 ** 3
    assert abs(s.specific_volume - spec) < 1e-9 * spec
    # setProblem's connectToNode(goal) already grew the tree towards the goal until the cube blocked it
    n = s.tree_size + 1.0  # the radius is computed before the extension (rrt_star.cpp:108)
    assert s.tree_size == 4  # start + three 0.25-steps along the diagonal before |q|inf <= 1
    s.update(np.array([0.0, 2.0, 2.0]))
    r = 1.1 * (2 * (1 + 1 / d)) ** (1 / d) * spec ** (1 / d) * (math.log(n) / n) ** (1 / d)
    assert abs(s.radius - r) < 1e-12


import sys  # noqa: E402

sys.path.insert(0, str(GOLD))


def test_oracle_grid_culling_keeps_every_verdict(orc):
    """The uniform-grid culling the CPU arm of the benchmark may switch on (orc_world_enable_grid) must not change a
    single verdict of the sphere-arm check: random states, states on the first contact of random edges (bisected to one
    ulp), clouds of several sizes (partial cells, empty grid regions)."""
    from graph_core_b200 import worlds as W
    rng = np.random.default_rng(77)
    for nobs in (1, 37, 1000, 1500):
        sc = W.scenario_c3(nobs)
        plain = orc.OracleWorld.from_scenario(sc)
        grid = orc.OracleWorld.from_scenario(sc)
        assert grid.enable_grid(True) > 0
        q = rng.uniform(-np.pi, np.pi, size=(6000, 7))
        a = plain.check(q)
        assert np.array_equal(a, grid.check(q)), nobs
        # grazing contacts: bisect free -> colliding pairs down to neighbouring doubles along the segment
        free, hit = q[a == 1][:60], q[a == 0][:60]
        n = min(len(free), len(hit))
        lo, hi = np.zeros(n), np.ones(n)
        for _ in range(60):
            mid = 0.5 * (lo + hi)
            qm = free[:n] + (hit[:n] - free[:n]) * mid[:, None]
            ok = plain.check(qm).astype(bool)
            lo, hi = np.where(ok, mid, lo), np.where(ok, hi, mid)
        for t in (lo, hi, np.nextafter(lo, 2.0), np.nextafter(hi, -1.0)):
            qg = free[:n] + (hit[:n] - free[:n]) * t[:, None]
            assert np.array_equal(plain.check(qg), grid.check(qg)), nobs
        grid.reset_counters()
        plain.reset_counters()
        grid.check(q[:500])
        plain.check(q[:500])
        if nobs >= 1000:
            assert grid.counters()["pair_tests"] * 20 < plain.counters()["pair_tests"]


def test_anytime_golden(orc):
    """tests/golden/anytime_golden.json holds the round logs, costs and way points of the REFERENCE's own AnytimeRRT
    (oracle/_ref, minted by tests/golden/make_anytime_golden.py) on 2-D box and 7-DoF sphere-arm queries: the
    restatement reproduces every one of them bit for bit, also where oracle/_ref is not available."""
    import make_anytime_golden as m
    from graph_core_b200 import worlds as W
    want = json.loads((GOLD / "anytime_golden.json").read_text())
    assert want["columns"] == list(orc.ANYTIME_LOG_COLUMNS)
    for name, scn, n, max_iter, seed, extend, iis in m.CASES:
        case = want["cases"][name]
        sc, starts, goals, cost0 = m.problems(orc, W, scn, n)
        for p in range(n):
            o = orc.OracleAnytime(sc, orc.OracleWorld.from_scenario(sc), start=starts[p], goal=goals[p], extend=extend,
                                  sampler_cost0=float(cost0[p]))
            solved = o.solve(max_iter, seed, p, n, improve_in_solve=iis)
            w = case["results"][p]
            assert solved == w["solved"], (name, p)
            assert m.hexf(o.log()) == w["log"], (name, p)
            if solved:
                assert float(o.state()["cost"]).hex() == w["cost"], (name, p)
                assert m.hexf(o.solution()) == w["solution"], (name, p)
