"""GPU parity of the fused planner step (gc_extend_batch) and the Philox sampler.

gc_extend_batch = Tree::findClosestNode (tree.cpp:56-59) + selectNextConfiguration (:110-128) +
tryExtendFromNode's edge check (:69-81) + Tree::near (:751-754), for many lock-step trees.
"""
import numpy as np
import pytest

from graph_core_b200 import worlds as W

pytestmark = pytest.mark.gpu


def _steer_oracle(node, q, max_distance):
    d = np.sqrt(np.sum((node - q) * (node - q)))  # same op order only matters to 1 ulp; compare with tolerance
    if d <= max_distance:
        return d, q.copy()
    return d, node + (q - node) / d * max_distance


@pytest.mark.parametrize("name", ["C1", "C3"])
def test_extend_batch_matches_oracle_ops(gc, orc, c3, name):
    sc = {"C1": W.scenario_c1(), "C3": c3}[name]
    dim, nseg = sc.dim, 6
    gw = sc.make_world()
    ow = orc.OracleWorld.from_scenario(sc)
    st = gc.NodeStore(dim, 2048, segments=nseg)
    oracles = []
    for s in range(nseg):
        pts = W.uniform_samples(sc.lb, sc.ub, 1000 + s, 0, 40 + 300 * s)
        st.append(pts, segment=s)
        o = orc.OracleNN("vector", dim)
        o.insert(pts)
        oracles.append((o, pts))
    n = 48
    q = W.uniform_samples(sc.lb, sc.ub, 1100, 0, n)
    seg = (np.arange(n) % nseg).astype(np.uint32)
    q[5] = oracles[seg[5]][1][3]  # a sample that coincides with a tree node: dist 0 < TOLERANCE
    radius = np.full(n, 0.2 if name == "C1" else 1.5)
    closest, dist, nxt, ok, nidx, ndist, ncount = gc.extend_batch(st, gw, q, seg, sc.max_distance, sc.min_distance,
                                                                  1e-6, radius, near_cap=2048)
    for i in range(n):
        o, pts = oracles[seg[i]]
        iv, dv = o.nearest(q[i])
        assert closest[i] == iv[0] and dist[i] == dv[0]
        d, want_next = _steer_oracle(pts[iv[0]], q[i], sc.max_distance)
        if dv[0] <= sc.max_distance:
            assert np.array_equal(nxt[i], q[i])  # same bits as the sample (tree.cpp:117-119)
        else:
            assert np.allclose(nxt[i], want_next, rtol=0, atol=1e-15)
        want_ok = 1 if dv[0] < 1e-6 else int(ow.check_connection(pts[iv[0]], nxt[i], sc.min_distance)[0])
        assert ok[i] == want_ok
        oi, od, oc = o.near(nxt[i], radius[i], cap=2048)
        assert ncount[i] == oc[0]
        assert np.array_equal(nidx[i, :oc[0]].astype(np.int64), oi[0, :oc[0]].astype(np.int64))
        assert np.array_equal(ndist[i, :oc[0]], od[0, :oc[0]])
    assert ok[5] == 1 and dist[5] == 0.0
    assert 0 < ok.sum() <= n


def test_extend_batch_empty_tree_and_no_radius(gc, orc):
    sc = W.scenario_c1()
    gw = sc.make_world()
    st = gc.NodeStore(2, 64, segments=2)
    st.append(np.array([[0.05, 0.05]]), segment=0)
    q = np.array([[0.06, 0.05], [0.5, 0.5]])
    closest, dist, nxt, ok = gc.extend_batch(st, gw, q, np.array([0, 1], np.uint32), 0.1, 0.01)
    assert closest[0] == 0 and ok[0] == 1
    assert closest[1] == 0xFFFFFFFF and ok[1] == 0 and np.isinf(dist[1])


def test_philox_uniform_sampler_bit_exact(gc):
    lb = np.array([-3.0, -1.0, 0.0, 2.0, -np.pi, 0.5, -0.25])
    ub = np.array([3.0, 1.0, 1.0, 2.5, np.pi, 0.75, 0.25])
    got = gc.sample_uniform(lb, ub, seed=0x1234567890ABCDEF, first_index=(1 << 32) - 5, n=4099)
    want = W.uniform_samples(lb, ub, 0x1234567890ABCDEF, (1 << 32) - 5, 4099)
    assert np.array_equal(got, want)
    assert (got >= np.minimum(lb, ub)).all() and (got <= np.maximum(lb, ub)).all()
    # distribution sanity: mean close to the box centre
    assert np.allclose(got.mean(0), 0.5 * (lb + ub), atol=0.1)
