"""GPU tests of the large-N k-nearest path (sample threshold -> collecting scan -> exact radix select)
and of the node-sharded merge (gc_merge_shard_lists_dev).  On the single GPU the driver's `-m gpu` run
has, G shards are emulated as G stores in one process: the per-shard kernels, the global-slot arithmetic
and the merge kernel are exactly what every rank of a multi-GPU run executes; the all-gather itself is
covered by tests/test_sharded_cpu.py (gloo) and tools/sharded_check.py (NCCL, torchrun)."""
import numpy as np
import pytest
import torch

from graph_core_b200 import worlds as W

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,dim,k", [(20000, 12, 1), (50000, 12, 64), (70001, 7, 1024), (40000, 3, 2500)])
def test_knn_large_tree_matches_vector(gc, orc, n, dim, k):
    pts = W.c4_nodes(n, dim, seed=91)
    qs = W.c4_queries(12, dim, seed=92)
    st = gc.NodeStore(dim, n)
    st.append(pts)
    ov = orc.OracleNN("vector", dim)
    ov.insert(pts)
    gi, gd, gcnt = st.knn(qs, k, cap=k)
    oi, od, oc = ov.knn(qs, k, k)
    assert np.array_equal(gcnt.astype(np.int64), oc.astype(np.int64))
    assert np.array_equal(gi.astype(np.int64), oi.astype(np.int64))
    assert np.array_equal(gd, od)


def test_knn_large_tree_duplicates_and_tombstones(gc, orc):
    # many exactly equal distances across the k-th rank: the slot order decides (vector.cpp:83-96)
    base = W.c4_nodes(6000, 5, seed=93)
    pts = np.vstack([base, base, base, base])  # every configuration four times -> 24000 nodes
    st = gc.NodeStore(5, len(pts))
    st.append(pts)
    ok = orc.OracleNN("kdtree", 5)
    ok.insert(pts)
    ov = orc.OracleNN("vector", 5)
    ov.insert(pts)
    qs = np.vstack([base[:4], W.c4_queries(4, 5, seed=94)])
    for k in (2, 6, 50):
        gi, gd, gcnt = st.knn(qs, k, cap=k)
        oi, od, oc = ov.knn(qs, k, k)
        assert np.array_equal(gi.astype(np.int64), oi.astype(np.int64)) and np.array_equal(gd, od)
    dead = sorted(set(ov.knn(qs, 3, 3)[0].ravel().tolist()))
    for s in dead:
        st.tombstone(s)
        ov.delete(s)
    # after the deletions the Vector oracle has re-numbered its nodes: compare through configurations
    gi, gd, gcnt = st.knn(qs, 20, cap=20)
    oi, od, oc = ov.knn(qs, 20, 20)
    assert np.array_equal(gd, od)
    assert not np.isin(gi, dead).any()


@pytest.mark.parametrize("G", [2, 4, 8])
def test_sharded_merge_emulated_shards(gc, orc, G):
    from graph_core_b200.sharded import CudaShardEngine
    dim, n = 12, 30011
    pts = W.c4_nodes(n, dim, seed=95)
    qs = W.c4_queries(16, dim, seed=96)
    engines = []
    for g in range(G):
        e = CudaShardEngine(dim, n // G + 2, 0)
        e.append(pts[g::G])  # node i -> shard i mod G, local slot i div G
        engines.append(e)
    ov = orc.OracleNN("vector", dim)
    ov.insert(pts)

    def exchange(lists, k_out, out_cap):
        idx = torch.stack([l[0] for l in lists])
        dd = torch.stack([l[1] for l in lists])
        cnt = torch.stack([l[2] for l in lists])
        oi, od, oc = engines[0].merge(idx, dd, cnt, k_out, out_cap)
        torch.cuda.synchronize()
        return oi.cpu().numpy(), od.cpu().numpy(), oc.cpu().numpy()

    oi, od, oc = exchange([e.nearest_lists(qs) for e in engines], 1, 1)
    ri, rd = ov.nearest(qs)
    assert np.array_equal(oi[:, 0], ri) and np.array_equal(od[:, 0], rd)
    k = 33
    oi, od, oc = exchange([e.knn_lists(qs, k) for e in engines], k, k)
    ri, rd, rc = ov.knn(qs, k, k)
    assert np.array_equal(oi, ri) and np.array_equal(od, rd) and np.array_equal(oc, rc)
    radius = W.c4_radius(40, n, dim)
    cap = 256
    oi, od, oc = exchange([e.near_lists(qs, radius, cap) for e in engines], G * cap, G * cap)
    ri, rd, rc = ov.near(qs, radius, 1024)
    assert np.array_equal(oc, rc) and rc.sum() > 0
    for q in range(len(qs)):
        assert np.array_equal(oi[q, :rc[q]], ri[q, :rc[q]]) and np.array_equal(od[q, :rc[q]], rd[q, :rc[q]])


def test_sharded_store_world_size_one(gc, orc):
    from graph_core_b200.sharded import ShardedNodeStore
    dim, n = 7, 5000
    pts = W.c4_nodes(n, dim, seed=97)
    qs = W.c4_queries(8, dim, seed=98)
    st = ShardedNodeStore(dim, n, device=0)
    st.append(pts)
    ov = orc.OracleNN("vector", dim)
    ov.insert(pts)
    gi, gd = st.nearest(qs)
    ri, rd = ov.nearest(qs)
    assert np.array_equal(gi.cpu().numpy(), ri) and np.array_equal(gd.cpu().numpy(), rd)


@pytest.mark.parametrize("G", [2, 8])
def test_sharded_merge_p2p_kernel_emulated_shards(gc, orc, G):
    """gc_merge_shard_lists_p2p_dev reads every shard's lists through its own pointer (peer memory on a multi-GPU
    box; here G stores of one process) and must equal the gathered merge."""
    import ctypes as C
    from graph_core_b200.sharded import CudaShardEngine
    from graph_core_b200 import capi
    dim, n, k = 12, 20011, 21
    pts = W.c4_nodes(n, dim, seed=195)
    qs = W.c4_queries(9, dim, seed=196)
    engines = []
    for g in range(G):
        e = CudaShardEngine(dim, n // G + 2, 0)
        e.append(pts[g::G])
        engines.append(e)
    lists = [e.knn_lists(qs, k) for e in engines]
    torch.cuda.synchronize()
    arr = C.c_void_p * G
    ip = arr(*[C.c_void_p(l[0].data_ptr()) for l in lists])
    dp = arr(*[C.c_void_p(l[1].data_ptr()) for l in lists])
    cp = arr(*[C.c_void_p(l[2].data_ptr()) for l in lists])
    nq = len(qs)
    oi = torch.full((nq, k), -1, dtype=torch.int32, device="cuda")
    od = torch.full((nq, k), float("inf"), dtype=torch.float64, device="cuda")
    oc = torch.zeros(nq, dtype=torch.int32, device="cuda")
    capi._check(capi.load().gc_merge_shard_lists_p2p_dev(0, torch.cuda.current_stream().cuda_stream, G, nq, k, ip, dp, cp,
                                                         k, k, oi.data_ptr(), od.data_ptr(), oc.data_ptr()))
    torch.cuda.synchronize()
    ov = orc.OracleNN("vector", dim)
    ov.insert(pts)
    ri, rd, rc = ov.knn(qs, k, k)
    assert np.array_equal(oi.cpu().numpy(), ri) and np.array_equal(od.cpu().numpy(), rd) and np.array_equal(oc.cpu().numpy(), rc)


# ---- beyond one CTA's sort capacity (16 384 entries): run sort + rank-merge passes -------------------------------
def test_knn_all_nodes_sorted_like_tree_nearK(gc, orc):
    """Tree::nearK asks for k = ceil(k_rrt * ln(N + 1)) neighbours (tree.cpp:52-53, 761-765); in 12 dimensions k_rrt =
    1.1 * 2^13 * e * 13/12 is far above N, so the answer is every live node in (distance, slot) order."""
    dim, n = 12, 1 << 17
    pts = W.c4_nodes(n, dim, seed=291)
    qs = W.c4_queries(3, dim, seed=292)
    k_rrt = 1.1 * 2.0 ** (dim + 1) * np.e * (1.0 + 1.0 / dim)
    k = int(np.ceil(k_rrt * np.log(n + 1)))
    assert k >= n
    st = gc.NodeStore(dim, n)
    st.append(pts)
    dead = [5, 70000, n - 1]
    for s in dead:
        st.tombstone(s)
    gi, gd, gcnt = st.knn(qs, k, cap=n)
    assert np.all(gcnt == n - len(dead))
    d = np.empty((len(qs), n))
    for j, q in enumerate(qs):  # Eigen (a-b).norm(): sequential sum of squares
        s = np.zeros(n)
        for c in range(dim):
            t = pts[:, c] - q[c]
            s = s + t * t
        d[j] = np.sqrt(s)
    d[:, dead] = np.inf
    for j in range(len(qs)):
        order = np.lexsort((np.arange(n), d[j]))[: n - len(dead)]
        assert np.array_equal(gi[j, : n - len(dead)].astype(np.int64), order)
        assert np.array_equal(gd[j, : n - len(dead)], d[j, order])


@pytest.mark.parametrize("n,dim,k", [(50000, 7, 20000), (40000, 12, 39999), (33000, 3, 16385)])
def test_knn_k_above_sort_capacity(gc, orc, n, dim, k):
    pts = W.c4_nodes(n, dim, seed=293)
    pts[1000:1400] = pts[:400]  # exact ties across the result: the lower slot comes first
    qs = np.vstack([W.c4_queries(3, dim, seed=294), pts[:1]])
    st = gc.NodeStore(dim, n)
    st.append(pts)
    ov = orc.OracleNN("vector", dim)
    ov.insert(pts)
    gi, gd, gcnt = st.knn(qs, k, cap=k)
    oi, od, oc = ov.knn(qs, k, k)
    assert np.array_equal(gcnt.astype(np.int64), oc.astype(np.int64))
    assert np.array_equal(gd, od)
    assert np.array_equal(gi.astype(np.int64), oi.astype(np.int64))


def test_radius_list_above_sort_capacity(gc, orc):
    dim, n = 3, 70000
    pts = W.c4_nodes(n, dim, seed=295)
    qs = W.c4_queries(4, dim, seed=296)
    st = gc.NodeStore(dim, n)
    st.append(pts)
    ov = orc.OracleNN("vector", dim)
    ov.insert(pts)
    radius = 3.2
    ri, rd, rc = ov.near(qs, radius, n)
    assert rc.max() > 16384
    gi, gd, gcnt = st.near(qs, radius, cap=512)  # grows on GC_W_TRUNCATED
    assert np.array_equal(gcnt.astype(np.int64), rc.astype(np.int64))
    for q in range(len(qs)):
        assert np.array_equal(gi[q, :rc[q]].astype(np.int64), ri[q, :rc[q]].astype(np.int64))
        assert np.array_equal(gd[q, :rc[q]], rd[q, :rc[q]])


@pytest.mark.parametrize("p2p", [False, True])
def test_sharded_merge_above_sort_capacity(gc, orc, p2p):
    import ctypes as C
    from graph_core_b200.sharded import CudaShardEngine
    from graph_core_b200 import capi
    G, dim, n, k = 4, 5, 60000, 9000  # 4 x 9000 gathered entries > 16 384
    pts = W.c4_nodes(n, dim, seed=297)
    qs = W.c4_queries(5, dim, seed=298)
    engines = []
    for g in range(G):
        e = CudaShardEngine(dim, n // G + 2, 0)
        e.append(pts[g::G])
        engines.append(e)
    lists = [e.knn_lists(qs, k) for e in engines]
    torch.cuda.synchronize()
    ov = orc.OracleNN("vector", dim)
    ov.insert(pts)
    ri, rd, rc = ov.knn(qs, k, k)
    if not p2p:
        idx = torch.stack([l[0] for l in lists])
        dd = torch.stack([l[1] for l in lists])
        cnt = torch.stack([l[2] for l in lists])
        oi, od, oc = engines[0].merge(idx, dd, cnt, k, k)
    else:
        arr = C.c_void_p * G
        ip = arr(*[C.c_void_p(l[0].data_ptr()) for l in lists])
        dp = arr(*[C.c_void_p(l[1].data_ptr()) for l in lists])
        cp = arr(*[C.c_void_p(l[2].data_ptr()) for l in lists])
        nq = len(qs)
        oi = torch.full((nq, k), -1, dtype=torch.int32, device="cuda")
        od = torch.full((nq, k), float("inf"), dtype=torch.float64, device="cuda")
        oc = torch.zeros(nq, dtype=torch.int32, device="cuda")
        capi._check(capi.load().gc_merge_shard_lists_p2p_dev(0, torch.cuda.current_stream().cuda_stream, G, nq, k, ip, dp,
                                                             cp, k, k, oi.data_ptr(), od.data_ptr(), oc.data_ptr()))
    torch.cuda.synchronize()
    assert np.array_equal(oc.cpu().numpy(), rc)
    assert np.array_equal(od.cpu().numpy(), rd) and np.array_equal(oi.cpu().numpy(), ri)


# ---- the two-launch latency path: a few queries, small k, large single tree (knn_small_scan / knn_small_merge) --------
@pytest.mark.parametrize("n,dim,k,nq", [(50000, 12, 16, 1), (70001, 7, 64, 8), (40000, 3, 5, 3), (1 << 20, 12, 16, 2),
                                        (33000, 2, 1, 8)])
def test_knn_few_queries_latency_path(gc, orc, n, dim, k, nq):
    pts = W.c4_nodes(n, dim, seed=391)
    qs = W.c4_queries(nq, dim, seed=392)
    st = gc.NodeStore(dim, n)
    st.append(pts)
    l0 = gc.capi.launch_count()
    gi, gd, gcnt = st.knn(qs, k, cap=k)
    assert gc.capi.launch_count() - l0 == 2  # scan + merge, nothing else
    ov = orc.OracleNN("vector", dim)
    ov.insert(pts)
    oi, od, oc = ov.knn(qs, k, k)
    assert np.array_equal(gcnt.astype(np.int64), oc.astype(np.int64))
    assert np.array_equal(gd, od)
    assert np.array_equal(gi.astype(np.int64), oi.astype(np.int64))


def test_knn_few_queries_ties_tombstones_and_fallback(gc, orc):
    base = W.c4_nodes(9000, 5, seed=393)
    pts = np.vstack([base, base, base, base])  # every configuration four times: exact ties across the k-th rank
    st = gc.NodeStore(5, len(pts))
    st.append(pts)
    ov = orc.OracleNN("vector", 5)
    ov.insert(pts)
    qs = np.vstack([base[:3], W.c4_queries(3, 5, seed=394)])
    for k in (1, 2, 6, 50):
        gi, gd, gcnt = st.knn(qs, k, cap=k)
        oi, od, oc = ov.knn(qs, k, k)
        assert np.array_equal(gi.astype(np.int64), oi.astype(np.int64)) and np.array_equal(gd, od)
    dead = sorted(set(ov.knn(qs, 3, 3)[0].ravel().tolist()))
    for s in dead:
        st.tombstone(s)
        ov.delete(s)
    gi, gd, gcnt = st.knn(qs, 20, cap=20)
    oi, od, oc = ov.knn(qs, 20, 20)
    assert np.array_equal(gd, od)
    assert not np.isin(gi, dead).any()
    # 40 000 copies of one point: every strip list overflows -> the general pipeline answers (lowest slots win)
    same = np.tile(W.c4_nodes(1, 4, seed=395), (40000, 1))
    st2 = gc.NodeStore(4, len(same))
    st2.append(same)
    gi, gd, gcnt = st2.knn(same[:2], 7, cap=7)
    assert np.array_equal(gi, np.tile(np.arange(7, dtype=np.uint32), (2, 1))) and np.all(gd == 0.0)


# ---- the node-sharded tree behind the C ABI (gc_comm_init / gc_store_create_sharded, csrc/sharded.cu) ----------------
@pytest.mark.parametrize("G", [1, 3, 8])
def test_c_abi_sharded_store_matches_vector(gc, orc, G):
    """All shards on device 0 here (the driver's GPU tests have one GPU): the per-shard kernels, the global-slot
    arithmetic, the event ordering between the shards' streams and the peer-pointer merge are what a multi-GPU box runs
    (tools/sharded_capi_check.py repeats this with one shard per GPU)."""
    dim, n = 12, 40013
    pts = W.c4_nodes(n, dim, seed=491)
    qs = W.c4_queries(20, dim, seed=492)
    st = gc.ShardedStore(dim, n + 64, [0] * G)
    first = st.append(pts[:1000])
    assert first == 0
    assert st.append(pts[1000:1007]) == 1000      # a batch that does not divide by G
    assert st.append(pts[1007:]) == 1007
    assert st.size() == (n, G)
    ov = orc.OracleNN("vector", dim)
    ov.insert(pts)
    gi, gd = st.nearest(qs)
    ri, rd = ov.nearest(qs)
    assert np.array_equal(gi.astype(np.int64), ri.astype(np.int64)) and np.array_equal(gd, rd)
    for k in (1, 7, 300):
        gi, gd, gcnt = st.knn(qs, k)
        ri, rd, rc = ov.knn(qs, k, k)
        assert np.array_equal(gcnt.astype(np.int64), rc.astype(np.int64))
        assert np.array_equal(gi.astype(np.int64), ri.astype(np.int64)) and np.array_equal(gd, rd)
    radius = W.c4_radius(60, n, dim)
    gi, gd, gcnt = st.near(qs, radius, cap=16)  # too small on purpose: GC_W_TRUNCATED, the wrapper grows it
    ri, rd, rc = ov.near(qs, radius, 4096)
    assert np.array_equal(gcnt.astype(np.int64), rc.astype(np.int64)) and rc.max() > 16
    for q in range(len(qs)):
        assert np.array_equal(gi[q, :rc[q]].astype(np.int64), ri[q, :rc[q]].astype(np.int64))
        assert np.array_equal(gd[q, :rc[q]], rd[q, :rc[q]])


def test_c_abi_sharded_store_edge_cases(gc, orc):
    from graph_core_b200 import capi
    st = gc.ShardedStore(3, 100, [0, 0, 0, 0])
    gi, gd = st.nearest(np.zeros((2, 3)))
    assert np.all(gi == capi.GC_NO_NODE) and np.all(np.isinf(gd))   # empty tree (kdtree.cpp:356-358)
    pts = W.c4_nodes(2, 3, seed=493)                                 # fewer nodes than shards
    st.append(pts)
    gi, gd, gcnt = st.knn(pts[:1], 5)
    assert gcnt[0] == 2 and gi[0, 0] == 0 and gd[0, 0] == 0.0 and gi[0, 1] == 1
    dup = np.tile(pts[:1], (9, 1))                                   # exact ties across shards: lowest global slot first
    st.append(dup)
    gi, gd, gcnt = st.knn(pts[:1], 6)
    assert list(gi[0]) == [0, 2, 3, 4, 5, 6] and np.all(gd[0] == 0.0)
    with pytest.raises(capi.GcError):
        st.append(np.zeros((200, 3)))                                # beyond the capacity
    with pytest.raises(capi.GcError):
        gc.ShardedStore(3, 100, [0, 99])                             # no such device
