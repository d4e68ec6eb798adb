"""Multi-process (gloo, world_size 2) test of the node-sharded query plumbing on CPU: round-robin
ownership, the single all-gather exchange and the global-slot arithmetic of
graph_core_b200/sharded.py.  The local slice and the merge are test doubles (oracle-backed / numpy);
the CUDA engine and merge kernel are covered by tests/test_sharded_gpu.py."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
NO = -1  # 0xFFFFFFFF as int32


class OracleShardEngine:
    """Local slice answered by the oracle's Vector restatement (test double of CudaShardEngine)."""

    def __init__(self, dim):
        sys.path.insert(0, str(ROOT / "oracle"))
        import oracle_py
        self.nn = oracle_py.OracleNN("vector", dim)
        self.dim = dim

    def append(self, rows):
        if len(rows):
            self.nn.insert(rows)

    @staticmethod
    def _t(ids, d, cnt):
        return (torch.from_numpy(ids.astype(np.int32)), torch.from_numpy(d), torch.from_numpy(cnt.astype(np.int32)))

    def nearest_lists(self, q):
        ids, d = self.nn.nearest(q)
        return self._t(ids[:, None], d[:, None], (ids >= 0).astype(np.int32))

    def knn_lists(self, q, k):
        return self._t(*self.nn.knn(q, k, k))

    def near_lists(self, q, radius, cap):
        return self._t(*self.nn.near(q, radius, cap))

    def merge(self, idx, dd, cnt, k_out, out_cap):
        G, nq, cap = idx.shape
        oi = np.full((nq, out_cap), NO, np.int32)
        od = np.full((nq, out_cap), np.inf)
        oc = np.zeros(nq, np.int32)
        for q in range(nq):
            ent = []
            for g in range(G):
                for i in range(min(int(cnt[g, q]), cap)):
                    ent.append((float(dd[g, q, i]), int(idx[g, q, i]) * G + g))
            ent.sort()
            ent = ent[:min(k_out, out_cap)]
            oc[q] = len(ent)
            for i, (d, s) in enumerate(ent):
                od[q, i], oi[q, i] = d, s
        return torch.from_numpy(oi), torch.from_numpy(od), torch.from_numpy(oc)


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle_py
    from graph_core_b200 import worlds as W
    from graph_core_b200.sharded import ShardedNodeStore
    dim, n = 7, 2001
    pts = W.c4_nodes(n, dim, seed=77)
    qs = W.c4_queries(23, dim, seed=78)
    st = ShardedNodeStore(dim, n, engine=OracleShardEngine(dim))
    assert st.world == world and st.rank == rank
    # incremental appends of uneven sizes keep the round-robin ownership
    pos = 0
    for chunk in (1, 2, 500, 1498):
        assert st.append(pts[pos:pos + chunk]) == pos
        pos += chunk
    assert st.engine.nn.size() == len(range(rank, n, world))
    ref = oracle_py.OracleNN("vector", dim)
    ref.insert(pts)
    gi, gd = st.nearest(qs)
    ri, rd = ref.nearest(qs)
    assert np.array_equal(gi.numpy(), ri) and np.array_equal(gd.numpy(), rd)
    ki, kd, kc = st.knn(qs, 9)
    ri, rd, rc = ref.knn(qs, 9, 9)
    assert np.array_equal(ki.numpy(), ri) and np.array_equal(kd.numpy(), rd) and np.array_equal(kc.numpy(), rc)
    ni, nd, nc = st.near(qs, 2.4, cap=128)
    ri, rd, rc = ref.near(qs, 2.4, 256)
    assert np.array_equal(nc.numpy(), rc)
    for q in range(len(qs)):
        assert np.array_equal(ni.numpy()[q, :rc[q]], ri[q, :rc[q]])
        assert np.array_equal(nd.numpy()[q, :rc[q]], rd[q, :rc[q]])
    # ties across shards: the same configuration stored twice lands on two different shards;
    # the lower GLOBAL slot must win (vector.cpp:55-67)
    st2 = ShardedNodeStore(dim, 16, engine=OracleShardEngine(dim))
    st2.append(np.vstack([pts[5], pts[6], pts[5], pts[6], pts[5]]))
    i2, d2 = st2.nearest(pts[5:7])
    assert i2.tolist() == [0, 1] and d2.tolist() == [0.0, 0.0]
    dist.barrier()
    Path(tmp, "ok%d" % rank).write_text("ok")
    dist.destroy_process_group()


def test_sharded_store_two_ranks_gloo(tmp_path):
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def _c5_worker(rank, world, port, tmp):
    """BASELINE config 5's sharding on CPU: query j on rank j mod world, per-query sample streams (k * total + j), one
    all-reduce gather of (solved, cost).  The planner is the oracle's AnytimeRRT restatement (test double of
    LockstepAnytimeRRT, which takes the same (stride, index) stream arguments)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle_py
    from graph_core_b200 import worlds as W
    from graph_core_b200.planner import gather_query_results, query_shard
    sc = W.scenario_c1()
    total, max_iter, seed = 5, 150, 61
    ow = oracle_py.OracleWorld.from_scenario(sc)
    u = W.philox_u01(62, 0, 64, 4)
    pts = sc.lb + u.reshape(-1, 2) * (sc.ub - sc.lb)
    pts = pts[ow.check(pts).astype(bool)][:2 * total]
    starts, goals = pts[:total], pts[total:2 * total]

    def solve(j):
        o = oracle_py.OracleAnytime(sc, oracle_py.OracleWorld.from_scenario(sc), start=starts[j], goal=goals[j])
        ok = o.solve(max_iter, seed, j, total)
        return ok, o.state()["cost"]

    mine = query_shard(total, rank, world)
    assert mine.tolist() == list(range(rank, total, world))
    res = [solve(int(j)) for j in mine]
    s_all, c_all = gather_query_results(dist, total, mine, np.array([r[0] for r in res]), np.array([r[1] for r in res]))
    want = [solve(j) for j in range(total)]  # the whole set in one process
    assert s_all.tolist() == [w[0] for w in want]
    assert all((c == w[1]) or (not w[0] and np.isinf(c)) for c, w in zip(c_all.tolist(), want))
    dist.barrier()
    Path(tmp, "c5ok%d" % rank).write_text("ok")
    dist.destroy_process_group()


def test_c5_query_sharding_two_ranks_gloo(tmp_path):
    port = 29900 + os.getpid() % 90
    mp.spawn(_c5_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "c5ok0").exists() and (tmp_path / "c5ok1").exists()
