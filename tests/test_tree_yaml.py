"""Tree YAML round trip: graph_core_b200/tree_yaml.py against the reference's own
`Tree::toYAML` / `Tree::fromYAML` (tree.cpp:1188-1227, :1272-1389) run from oracle/_ref.

The document is compared as data (nodes bit-exact and in order, connection pairs as a set, max_distance, use_kdtree):
yaml-cpp itself is not in this image, so the emitted text is checked with PyYAML and the semantics with the reference.
"""
import numpy as np
import pytest
import yaml

from graph_core_b200 import worlds as W
from graph_core_b200.tree_yaml import connections_from_parents, tree_from_yaml, tree_to_yaml


@pytest.fixture(scope="module")
def ref():
    import ref_py
    if not ref_py.available():
        try:
            ref_py.build()
        except Exception as ex:  # pragma: no cover
            pytest.skip("oracle/_ref not buildable here: %s" % ex)
    if not ref_py.available():
        pytest.skip("oracle/_ref/libgc_ref.so absent (built only where /root/reference exists)")
    ref_py.lib()
    return ref_py


def _parents_from_parent_conf(conf, parent_conf):
    index = {row.tobytes(): i for i, row in enumerate(conf)}
    assert len(index) == len(conf)
    return np.array([-1 if np.isnan(p[0]) else index[p.tobytes()] for p in parent_conf], np.int64)


def _reference_tree(ref, use_kdtree, iters=1500):
    sc = W.scenario_c0()
    samples = W.uniform_samples(sc.lb, sc.ub, 1, 0, iters)
    r = ref.RefSolver(ref.RefSolver.RRTSTAR, sc, ref.RefChecker.cube(3, 1.0, sc.min_distance), use_kdtree=use_kdtree,
                      informed=True)
    r.run(samples)
    return sc, r


@pytest.mark.parametrize("use_kdtree", [False, True])
def test_document_matches_reference_to_yaml(ref, use_kdtree):
    sc, r = _reference_tree(ref, use_kdtree)
    nodes, conns, max_distance, uk = r.tree_to_yaml()
    conf, parent_conf, _ = r.dump()
    assert np.array_equal(nodes, conf)  # document order = getNodes() order
    assert uk == use_kdtree and max_distance == sc.max_distance
    parents = _parents_from_parent_conf(conf, parent_conf)
    text = tree_to_yaml(conf, parents, sc.max_distance, use_kdtree)
    doc = yaml.safe_load(text)
    assert doc["max_distance"] == max_distance and doc["use_kdtree"] == uk
    assert np.array_equal(np.array(doc["nodes"], np.float64), nodes)  # bit-exact after the text round trip
    mine = {tuple(c) for c in doc["connections"]}
    assert len(mine) == len(doc["connections"]) == len(conns) == len(nodes) - 1
    assert mine == {(int(a), int(b)) for a, b in conns}
    # flow style for nodes and connections like the reference's emitter settings
    assert "  - [" in text and text.count("\n  - [") == len(nodes) + len(conns)


@pytest.mark.parametrize("with_max_distance", [True, False])
def test_loading_matches_reference_from_yaml(ref, with_max_distance):
    sc, r = _reference_tree(ref, False, iters=800)
    conf, parent_conf, pcost = r.dump()
    parents = _parents_from_parent_conf(conf, parent_conf)
    text = tree_to_yaml(conf, parents, sc.max_distance, use_kdtree=False)
    if not with_max_distance:
        text = "\n".join(line for line in text.splitlines() if not line.startswith("max_distance")) + "\n"
    doc = tree_from_yaml(text)
    assert np.array_equal(doc.conf, conf) and np.array_equal(doc.parents, parents)
    assert np.array_equal(doc.pcost, pcost)
    assert doc.use_kdtree is False
    want = ref.ref_tree_from_yaml(conf, connections_from_parents(parents),
                                  sc.max_distance if with_max_distance else None, use_kdtree=False)
    assert want is not None
    wconf, wparent, wcost, wmax = want
    assert np.array_equal(wconf, doc.conf)  # Vector store: insertion order = document order
    assert np.array_equal(_parents_from_parent_conf(wconf, wparent), doc.parents)
    assert np.array_equal(wcost, doc.pcost)
    assert wmax == doc.max_distance
    if not with_max_distance:
        assert doc.max_distance == pcost.max()
    # and the text is a fixed point
    assert tree_to_yaml(doc.conf, doc.parents, doc.max_distance, doc.use_kdtree) == \
        tree_to_yaml(conf, parents, doc.max_distance, False)


def test_reference_loads_what_we_write_with_kdtree(ref):
    """use_kdtree = true: the loaded reference tree enumerates its nodes in KdTree preorder; same node set, same
    parent of every node."""
    sc, r = _reference_tree(ref, True, iters=600)
    conf, parent_conf, pcost = r.dump()
    # put the root first (KdTree preorder starts at the root of the k-d tree, which is the first inserted node)
    parents = _parents_from_parent_conf(conf, parent_conf)
    assert parents[0] == -1
    doc = tree_from_yaml(tree_to_yaml(conf, parents, sc.max_distance, True))
    wconf, wparent, wcost, wmax = ref.ref_tree_from_yaml(doc.conf, connections_from_parents(doc.parents), doc.max_distance,
                                                         use_kdtree=True)
    want = ref.tree_as_edges(wconf, wparent, wcost)
    got = ref.tree_as_edges(conf, parent_conf, pcost)
    assert want == got and wmax == sc.max_distance


def test_malformed_documents_are_refused_like_upstream(ref):
    conf = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0]])
    good = tree_to_yaml(conf, [-1, 0, 1], 1.5)
    assert tree_from_yaml(good).parents.tolist() == [-1, 0, 1]
    for missing in ("nodes", "connections"):
        d = yaml.safe_load(good)
        del d[missing]
        with pytest.raises(ValueError):
            tree_from_yaml(d)
    with pytest.raises(ValueError):
        tree_from_yaml({"nodes": 3, "connections": []})
    with pytest.raises(ValueError):
        tree_from_yaml({"nodes": [[0.0, 0.0]], "connections": [[0, 4]]})
    with pytest.raises(ValueError):
        tree_to_yaml(conf, [-1, 0, 2], 1.0)
    # empty and single-node trees
    assert tree_from_yaml(tree_to_yaml(np.zeros((1, 3)), [-1], 0.3)).conf.shape == (1, 3)
    assert yaml.safe_load(tree_to_yaml(np.zeros((0, 3)), [], 0.3))["nodes"] == []
    # special values survive the text
    odd = np.array([[0.1, -0.0, 1e-310], [np.pi, 1e300, -2.5e-7]])
    back = tree_from_yaml(tree_to_yaml(odd, [-1, 0], 0.1))
    assert back.conf.tobytes() == odd.tobytes()


@pytest.mark.gpu
def test_lockstep_tree_to_yaml_and_back_into_a_store(ref):
    """The planner's device-resident tree -> YAML -> (a) the reference loads the same tree, (b) a fresh device store
    answers nearest queries like the original nodes."""
    from graph_core_b200 import capi
    from graph_core_b200.planner import LockstepRRTStar
    sc = W.scenario_c0()
    iters = 1500
    samples = W.uniform_samples(sc.lb, sc.ub, 1, 0, iters)
    world = capi.CollisionWorld.cube(3, 1.0)
    ls = LockstepRRTStar(sc, world, sc.start[None, :], sc.goal[None, :], capacity=2048)
    ls.run(samples[:, None, :])
    par, pc, conf = ls.dump(0)
    assert ls.info(0)["solved"]  # the goal node (slot 1) is connected: every slot is in the tree
    doc = tree_from_yaml(ls.to_yaml(0, use_kdtree=False))
    assert np.array_equal(doc.conf, conf) and np.array_equal(doc.parents, par)
    assert np.array_equal(doc.pcost[1:], pc[doc.parents >= 0]) and doc.max_distance == sc.max_distance
    # the reference, run on the same samples, writes the same tree (it lists the goal where it was connected, the
    # store keeps it in its creation slot: compare order-free)
    r = ref.RefSolver(ref.RefSolver.RRTSTAR, sc, ref.RefChecker.cube(3, 1.0, sc.min_distance), use_kdtree=False,
                      informed=True)
    r.run(samples)
    nodes, conns, max_distance, _ = r.tree_to_yaml()
    assert max_distance == doc.max_distance and len(nodes) == len(doc.conf)
    theirs = {nodes[b].tobytes(): nodes[a].tobytes() for a, b in conns}
    ours = {doc.conf[b].tobytes(): doc.conf[a].tobytes() for a, b in connections_from_parents(doc.parents)}
    assert theirs == ours and {q.tobytes() for q in nodes} == {q.tobytes() for q in doc.conf}
    # and the reference loads ours into the same tree
    wconf, wparent, wcost, _ = ref.ref_tree_from_yaml(doc.conf, connections_from_parents(doc.parents), doc.max_distance,
                                                      use_kdtree=False)
    assert ref.tree_as_edges(wconf, wparent, wcost) == ref.tree_as_edges(*r.dump())
    # an unsolved problem: the unconnected goal is left out
    ls2 = LockstepRRTStar(sc, world, sc.start[None, :], sc.goal[None, :], capacity=2048)
    ls2.run(samples[:40, None, :])
    assert not ls2.info(0)["solved"]
    d2 = tree_from_yaml(ls2.to_yaml(0))
    assert d2.conf.shape[0] == ls2.info(0)["tree_size"]
    assert not any(np.array_equal(q, sc.goal) for q in d2.conf)
    ls2.close()
    store = doc.to_store()
    assert store.size() == (conf.shape[0], conf.shape[0])  # (slots used, live nodes)
    assert np.array_equal(store.read(0, conf.shape[0]), conf)
    q = W.uniform_samples(sc.lb, sc.ub, 77, 0, 64)
    idx, dist = store.nearest(q)
    d = np.sqrt(((conf[None, :, :] - q[:, None, :]) ** 2).sum(-1))
    assert np.array_equal(idx, d.argmin(1))
    ls.close()
