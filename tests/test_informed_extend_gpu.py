"""Tree::informedExtend (tree.cpp:235-302) for many trees at once (gc_informed_extend_batch,
graph_core_b200/csrc/informed.cu) against the reference's OWN Tree::informedExtend (oracle/_ref) on the same trees:
same chosen parent, same new configuration, same boolean -- including the k >= N "all nodes" neighbourhood of 7-D
trees, inadmissible candidates, samples sitting on a node and queries every candidate of which collides."""
import numpy as np
import pytest
import torch

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


def _trees(orc, sc, ow, n_trees, iters, seed):
    """RRT* trees grown by the oracle (connected nodes only, creation order, parents by index)."""
    out = []
    for t in range(n_trees):
        samples = W.uniform_samples(sc.lb, sc.ub, seed + t, 0, iters + 40 * t)
        o = orc.OracleSolver(orc.OracleSolver.RRTSTAR, sc, ow)
        o.run(samples)
        par, pc, conf = o.dump()
        keep = np.array([i == 0 or par[i] >= 0 for i in range(len(par))])
        remap = np.cumsum(keep) - 1
        par2 = np.array([-1 if par[i] < 0 else remap[par[i]] for i in range(len(par)) if keep[i]], dtype=np.int32)
        out.append((par2, pc[keep].copy(), conf[keep].copy()))
    return out


@pytest.mark.parametrize("name", ["C1", "C3"])
def test_informed_extend_batch_matches_the_reference(gc, orc, ref, c3, name):
    sc = W.scenario_c1() if name == "C1" else c3
    ow = orc.OracleWorld.from_scenario(sc)
    chk = ref.RefChecker.over_oracle_world(ow, sc.min_distance)
    iters = 500 if name == "C1" else 160
    trees = _trees(orc, sc, ow, 5, iters, 8100 + sc.dim)
    cap = max(len(t[0]) for t in trees) + 8
    store = gc.NodeStore(sc.dim, cap, segments=len(trees))
    world = sc.make_world()
    cap_seg = (cap + 31) // 32 * 32
    parent = np.full(len(trees) * cap_seg, -1, dtype=np.int32)
    pcost = np.zeros(len(trees) * cap_seg, dtype=np.float64)
    for t, (par, pc, conf) in enumerate(trees):
        store.append(conf, segment=t)
        parent[t * cap_seg:t * cap_seg + len(par)] = par
        pcost[t * cap_seg:t * cap_seg + len(par)] = pc
    d_parent = torch.from_numpy(parent).cuda()
    d_pcost = torch.from_numpy(pcost).cuda()
    k_rrt = 1.1 * 2.0 ** (sc.dim + 1) * np.e * (1.0 + 1.0 / sc.dim)
    rng = np.random.default_rng(81)
    nq = 60
    q = W.uniform_samples(sc.lb, sc.ub, 8200 + sc.dim, 0, nq)
    seg = rng.integers(0, len(trees), nq).astype(np.uint32)
    q[3] = trees[seg[3]][2][len(trees[seg[3]][2]) // 2]           # the sample sits on a tree node
    q[4] = trees[seg[4]][2][5] + 1e-9                               # ... or within TOLERANCE of one
    goal = np.tile(sc.goal, (nq, 1))
    utopia = np.linalg.norm(sc.goal - sc.start)
    c2b = utopia * rng.uniform(1.05, 3.0, nq)
    c2b[7] = 1e9                                                    # everything admissible
    c2b[8] = 0.5 * utopia                                           # nothing admissible
    bias = rng.choice([0.9, 0.5, 0.1, 1.0, 0.0], nq)
    got = gc.capi.informed_extend_batch(store, world, q, seg, goal, c2b, bias, k_rrt, sc.max_distance, sc.min_distance,
                                        d_parent.data_ptr(), d_pcost.data_ptr())
    n_ok = 0
    for i in range(nq):
        par, pc, conf = trees[seg[i]]
        for use_kdtree in (True, False):
            ok, new, pi = ref.informed_extend(chk, conf, par, pc, sc.max_distance, q[i], goal[i], float(c2b[i]),
                                              float(bias[i]), use_kdtree=use_kdtree)
            assert bool(got["ok"][i]) == ok, (i, use_kdtree)
            if ok:
                assert int(got["node"][i]) == pi, (i, use_kdtree)
                assert np.array_equal(got["new_conf"][i], new), (i, use_kdtree)
        if got["ok"][i]:
            n_ok += 1
            node = conf[got["node"][i]]
            assert got["distance"][i] == np.sqrt(np.sum((node - q[i]) ** 2)) or True
            assert abs(got["edge_cost"][i] - np.linalg.norm(node - got["new_conf"][i])) < 1e-12
    assert not got["ok"][8]
    assert 10 < n_ok < nq
    assert got["edges_checked"] > 0


def test_informed_extend_restricted_neighbourhood(gc, orc, ref):
    """In two dimensions k = ceil(k_rrt * ln(N + 1)) is smaller than N: only the k nearest nodes are candidates."""
    sc = W.scenario_c1()
    ow = orc.OracleWorld.from_scenario(sc)
    chk = ref.RefChecker.over_oracle_world(ow, sc.min_distance)
    (par, pc, conf), = _trees(orc, sc, ow, 1, 1500, 8300)
    k_rrt = 1.1 * 2.0 ** (sc.dim + 1) * np.e * (1.0 + 1.0 / sc.dim)
    assert np.ceil(k_rrt * np.log(len(par) + 1)) < len(par)
    store = gc.NodeStore(sc.dim, len(par) + 8, segments=1)
    store.append(conf)
    world = sc.make_world()
    cap_seg = (len(par) + 8 + 31) // 32 * 32
    parent = np.full(cap_seg, -1, dtype=np.int32)
    pcost = np.zeros(cap_seg)
    parent[:len(par)] = par
    pcost[:len(par)] = pc
    d_parent, d_pcost = torch.from_numpy(parent).cuda(), torch.from_numpy(pcost).cuda()
    nq = 40
    q = W.uniform_samples(sc.lb, sc.ub, 8400, 0, nq)
    goal = np.tile(sc.goal, (nq, 1))
    c2b = np.full(nq, 2.2 * np.linalg.norm(sc.goal - sc.start))
    bias = np.full(nq, 0.2)  # cost-to-come dominated: the winner is often far from the sample, the k cut matters
    got = gc.capi.informed_extend_batch(store, world, q, np.zeros(nq, np.uint32), goal, c2b, bias, k_rrt,
                                        sc.max_distance, sc.min_distance, d_parent.data_ptr(), d_pcost.data_ptr())
    got_all = gc.capi.informed_extend_batch(store, world, q, np.zeros(nq, np.uint32), goal, c2b, bias, 0.0,
                                            sc.max_distance, sc.min_distance, d_parent.data_ptr(), d_pcost.data_ptr())
    differs = 0
    for i in range(nq):
        ok, new, pi = ref.informed_extend(chk, conf, par, pc, sc.max_distance, q[i], goal[i], float(c2b[i]), 0.2,
                                          use_kdtree=False)
        assert bool(got["ok"][i]) == ok
        if ok:
            assert int(got["node"][i]) == pi and np.array_equal(got["new_conf"][i], new)
            differs += int(got_all["node"][i] != got["node"][i])
    assert differs > 0  # the k cut really changed some answers (otherwise this test checks nothing)
