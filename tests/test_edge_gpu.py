"""GPU parity: CUDA state / edge checks vs the oracle's restatement of CollisionCheckerBase.

Reference semantics under test (include/graph_core/collision_checkers/collision_checker_base.h):
  check(q) :166, checkConnection(c1,c2) :207-237 (bisection), checkConnection(c1,c2,conf&) :178-205
  (linear, first failure), checkConnFromConf :264-313; Cube3dCollisionChecker::check
  (cube_3d_collision_checker.h:61-64).  Collision booleans and first-failure indices are compared
  bit for bit.
"""
import numpy as np
import pytest

from graph_core_b200 import worlds as W

pytestmark = pytest.mark.gpu


def _scenarios(c3):
    return {"C0": W.scenario_c0(), "C1": W.scenario_c1(), "C2": W.scenario_c2(), "C3": c3}


def _segments(sc, n, seed, lo_len, hi_len):
    """n random segments inside the bounds with lengths U[lo_len, hi_len]."""
    a = W.uniform_samples(sc.lb, sc.ub, seed, 0, n)
    u = W.philox_u01(seed + 1, 0, n, sc.dim + 1)
    direction = u[:, :sc.dim] - 0.5
    direction /= np.linalg.norm(direction, axis=1, keepdims=True)
    length = lo_len + u[:, sc.dim] * (hi_len - lo_len)
    b = np.clip(a + direction * length[:, None], sc.lb, sc.ub)
    return a, b


@pytest.mark.parametrize("name", ["C0", "C1", "C2", "C3"])
def test_state_check_matches_oracle(gc, orc, c3, name):
    sc = _scenarios(c3)[name]
    gw = sc.make_world()
    ow = orc.OracleWorld.from_scenario(sc)
    n = 20000 if name != "C3" else 6000
    q = W.uniform_samples(sc.lb, sc.ub, 900, 0, n)
    got = gw.check(q)
    want = ow.check(q)
    assert np.array_equal(got, want)
    assert 0 < want.sum() < n  # both outcomes present
    # tiny batches take the warp-per-state path of the arm world
    assert np.array_equal(gw.check(q[:3]), want[:3])


@pytest.mark.parametrize("name,lo_len,hi_len", [("C0", 0.0, 1.5), ("C1", 0.0, 0.3), ("C2", 0.05, 1.0),
                                                ("C3", 0.0, 0.6)])
def test_bisect_edges_match_oracle(gc, orc, c3, name, lo_len, hi_len):
    sc = _scenarios(c3)[name]
    gw = sc.make_world()
    ow = orc.OracleWorld.from_scenario(sc)
    n = 3000 if name != "C3" else 700
    a, b = _segments(sc, n, 910, lo_len, hi_len)
    got = gw.check_connection(a, b, sc.min_distance)
    want = ow.check_connection(a, b, sc.min_distance)
    assert np.array_equal(got, want)
    assert 0 < want.sum() < n
    # one edge at a time = the plugin's checkConnection(c1,c2) call pattern
    for i in range(5):
        assert gw.check_connection(a[i], b[i], sc.min_distance)[0] == want[i]


@pytest.mark.parametrize("name", ["C0", "C1", "C2", "C3"])
def test_linear_edges_match_oracle(gc, orc, c3, name):
    sc = _scenarios(c3)[name]
    gw = sc.make_world()
    ow = orc.OracleWorld.from_scenario(sc)
    n = 2000 if name != "C3" else 400
    a, b = _segments(sc, n, 920, 0.0, 0.8)
    ok, fb = gw.check_connection(a, b, sc.min_distance, mode=gc.capi.EDGE_LINEAR)
    wok, wfb, wconf = ow.check_connection(a, b, sc.min_distance, mode=1)
    assert np.array_equal(ok, wok)
    assert np.array_equal(fb, wfb)
    assert (fb >= 0).any() and (fb < 0).any()


@pytest.mark.parametrize("name", ["C0", "C2", "C3"])
def test_from_conf_edges_match_oracle(gc, orc, c3, name):
    sc = _scenarios(c3)[name]
    gw = sc.make_world()
    ow = orc.OracleWorld.from_scenario(sc)
    n = 1500 if name != "C3" else 300
    p, c = _segments(sc, n, 930, 0.1, 0.9)
    t = W.philox_u01(931, 0, n, 1)
    conf = p + (c - p) * t
    # off the segment: the reference's guard is (dist - dist_child - dist_parent) > 1e-4 (:273-277), which the
    # triangle inequality keeps <= 0, so it never fires; both sides must then agree on the ordinary result
    conf[::7] += 0.05
    got = gw.check_conn_from_conf(p, c, conf, sc.min_distance)
    want = ow.check_conn_from_conf(p, c, conf, sc.min_distance)
    assert np.array_equal(got, want)
    assert not (want == 255).any() and (want == 1).any() and (want == 0).any()


def test_bisection_state_set_edge_cases(gc, orc):
    """Lengths exactly on the reference's loop boundaries."""
    sc = W.scenario_c0()
    gw = sc.make_world()
    ow = orc.OracleWorld.from_scenario(sc)
    md = 0.01
    base = np.array([1.2, 0.0, 0.0])  # free (|q|max > 1); moving along -x enters the cube at x = 1.0
    lens = [0.0, 1e-12, md * 0.999, md, np.nextafter(md, 1), 2 * md, np.nextafter(2 * md, 1), 4 * md, 0.32,
            np.nextafter(0.32, 1), 0.64, 0.1999, 0.2, 0.2001, 0.25, 0.5, 1.0]
    a = np.repeat(base[None, :], len(lens), 0)
    b = a.copy()
    b[:, 0] -= np.array(lens)
    for mode in (0, 1):
        if mode == 0:
            got = gw.check_connection(a, b, md)
            want = ow.check_connection(a, b, md)
            assert np.array_equal(got, want)
        else:
            ok, fb = gw.check_connection(a, b, md, mode=1)
            wok, wfb, _ = ow.check_connection(a, b, md, mode=1)
            assert np.array_equal(ok, wok) and np.array_equal(fb, wfb)
    # an obstacle thinner than the effective spacing is missed by BOTH (bit-exact parity, not geometry)
    thin_lo = np.array([[0.50, -1.0]])
    thin_hi = np.array([[0.5001, 1.0]])
    gwb = gc.CollisionWorld.boxes(thin_lo, thin_hi)
    owb = orc.OracleWorld.boxes(thin_lo, thin_hi)
    a2 = np.array([[0.0, 0.0]] * 64) + np.linspace(0, 0.003, 64)[:, None]
    b2 = a2 + np.array([1.0, 0.0])
    assert np.array_equal(gwb.check_connection(a2, b2, md), owb.check_connection(a2, b2, md))


def test_path_yaml_fixture_is_collision_free(gc, orc):
    """The reference's only golden artefact (graph_core/tests/path.yaml, copied as data into
    tests/golden/path_yaml.json): every consecutive waypoint pair must pass
    Cube3dCollisionChecker(1.0) at min_distance 0.01, on the device as in the oracle."""
    import json
    from pathlib import Path
    wp = np.array(json.loads((Path(__file__).parent / "golden" / "path_yaml.json").read_text())["waypoints"])
    sc = W.scenario_c0()
    gw = sc.make_world()
    ow = orc.OracleWorld.from_scenario(sc)
    got = gw.check_connection(wp[:-1], wp[1:], 0.01)
    want = ow.check_connection(wp[:-1], wp[1:], 0.01)
    assert np.array_equal(got, want) and got.all()


def test_clone_shares_world_and_counters(gc, orc, c3):
    gw = c3.make_world()
    cl = gw.clone()
    q = W.uniform_samples(c3.lb, c3.ub, 940, 0, 64)
    assert np.array_equal(gw.check(q), cl.check(q))
    gw.close()  # the clone keeps the device world alive
    ow = orc.OracleWorld.from_scenario(c3)
    assert np.array_equal(cl.check(q), ow.check(q))
    cl.reset_counters()
    cl.check(q)
    c = cl.counters()
    assert c["states"] == 64 and c["pair_tests"] > 0 and c["launches"] == 1


def test_large_edge_batch_lane_per_state_path(gc, orc, c3):
    """Enough states to take the lane-per-state arm kernel (the throughput path used by bench.py)."""
    gw = c3.make_world()
    ow = orc.OracleWorld.from_scenario(c3)
    a, b = _segments(c3, 1500, 950, 0.3, 1.2)
    got = gw.check_connection(a, b, c3.min_distance)
    want = ow.check_connection(a, b, c3.min_distance)
    assert np.array_equal(got, want)
    ok, fb = gw.check_connection(a, b, c3.min_distance, mode=1)
    wok, wfb, _ = ow.check_connection(a, b, c3.min_distance, mode=1)
    assert np.array_equal(ok, wok) and np.array_equal(fb, wfb)


def test_clones_are_independent_handles_across_threads(gc, orc, c3):
    """CollisionCheckerBase::clone (:319) exists so that several threads can check at once: gc_world_clone gives
    every thread its own stream / scratch / counters over the one immutable device world."""
    import threading
    root = c3.make_world()
    ow = orc.OracleWorld.from_scenario(c3)
    batches = [_segments(c3, 300, 960 + t, 0.2, 0.9) for t in range(4)]
    want = [ow.check_connection(a, b, c3.min_distance) for a, b in batches]
    clones = [root.clone() for _ in range(4)]
    root.close()  # clones keep the device world alive
    got = [None] * 4

    def work(t):
        for _ in range(5):
            got[t] = clones[t].check_connection(*batches[t], c3.min_distance)

    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for t in range(4):
        assert np.array_equal(got[t], want[t])
        assert clones[t].counters()["launches"] == 5


@pytest.mark.parametrize("n_obs", [48, 64, 100, 333, 1000, 1024, 1100])
def test_block_culling_keeps_the_brute_force_verdict(gc, orc, n_obs):
    """The arm kernel skips blocks of 16 obstacles whose inflated bounding sphere no robot sphere reaches (edge.cu,
    arm_block_bounds): clouds below 64 and above 1024 obstacles take the plain loop, the others the culled one, with a
    partial last block when the size is not a multiple of 16.  Verdicts must be the oracle's brute-force ones for
    states (throughput and warp-per-state paths) and for edges in both modes, touching obstacles included."""
    sc = W.scenario_c3(n_obs, orc.arm_checker())
    gw = sc.make_world()
    ow = orc.OracleWorld.from_scenario(sc)
    q = W.uniform_samples(sc.lb, sc.ub, 4000 + n_obs, 0, 5000)
    want = ow.check(q)
    assert np.array_equal(gw.check(q), want)
    assert 0 < want.sum() < len(q)
    assert np.array_equal(gw.check(q[:5]), want[:5])
    a, b = _segments(sc, 1200, 4100 + n_obs, 0.0, 1.5)
    wok = ow.check_connection(a, b, sc.min_distance)
    assert np.array_equal(gw.check_connection(a, b, sc.min_distance), wok)
    assert 0 < wok.sum() < len(a)
    ok, fb = gw.check_connection(a, b, sc.min_distance, mode=1)
    lok, lfb, _ = ow.check_connection(a, b, sc.min_distance, mode=1)
    assert np.array_equal(ok, lok) and np.array_equal(fb, lfb)


def test_block_culling_grazing_contacts(gc, orc, c3):
    """States whose closest sphere pair is within a few ulps of touching: the culling margin must never hide the pair
    the exact fp32 test would call a collision.  Bisect between a free and a colliding state until they are adjacent
    in fp32 joint space, then check the whole bracket."""
    gw = c3.make_world()
    ow = orc.OracleWorld.from_scenario(c3)
    q = W.uniform_samples(c3.lb, c3.ub, 4321, 0, 4000)
    free = ow.check(q).astype(bool)
    lo, hi = q[free][:150], q[~free][:150]
    n = min(len(lo), len(hi))
    lo, hi = lo[:n].copy(), hi[:n].copy()
    for _ in range(40):
        mid = 0.5 * (lo + hi)
        ok = ow.check(mid).astype(bool)
        lo[ok] = mid[ok]
        hi[~ok] = mid[~ok]
    bracket = np.concatenate([lo, hi, 0.5 * (lo + hi), np.nextafter(lo, hi), np.nextafter(hi, lo)])
    assert np.array_equal(gw.check(bracket), ow.check(bracket))
