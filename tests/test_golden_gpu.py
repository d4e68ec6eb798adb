"""GPU parity against the REFERENCE's committed outputs: tests/golden/ref_golden.json was minted by
tests/golden/make_ref_golden.py from oracle/_ref (graph_core's own sources compiled in the build container).
The same quantities are recomputed here through the CUDA path (C ABI) on identical seeded inputs and
must match bit for bit: neighbour ids, ordering, fp64 distances, collision booleans, first-failure
configurations and the complete RRT* trees."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = Path(__file__).parent / "golden"
sys.path.insert(0, str(GOLD))


class _CudaNN:
    """Vector semantics on the device store; node id == slot (insertion order)."""

    def __init__(self, gc, dim):
        self.gc, self.dim, self.st = gc, dim, None

    def insert(self, pts):
        self.st = self.gc.NodeStore(self.dim, max(1024, len(pts)))
        self.st.append(pts)

    def nearest(self, q):
        i, d = self.st.nearest(q)
        return i.astype(np.int64), d

    def near(self, q, r, cap):
        i, d, c = self.st.near(q, r, cap=cap)
        return i.astype(np.int64), d, c.astype(np.int64)

    def knn(self, q, k, cap):
        i, d, c = self.st.knn(q, k, cap=cap)
        return i.astype(np.int64), d, c.astype(np.int64)

    def delete(self, slot):
        self.st.tombstone(slot)

    def size(self):
        return self.st.size()[1]


class CudaBackend:
    name = "cuda"

    def __init__(self, gc, orc):
        self.gc, self.orc = gc, orc

    def nn(self, kind, dim):
        # both upstream structures answer queries identically (KdTree == Vector without ties); one device store
        # stands behind both
        return _CudaNN(self.gc, dim)

    def checker(self, sc):
        return sc.make_world()

    def edges(self, w, sc, a, b):
        ok = w.check_connection(a, b, sc.min_distance)
        lok, fb = w.check_connection(a, b, sc.min_distance, mode=self.gc.capi.EDGE_LINEAR)
        # last feasible configuration as CudaCollisionChecker::checkConnection(c1,c2,conf) derives it from first_bad
        # (collision_checker_base.h:178-205)
        conf = np.full_like(a, np.nan)
        for i in range(len(a)):
            diff = b[i] - a[i]
            s = 0.0
            for v in diff:
                s += v * v
            npnt = int(np.ceil(np.sqrt(s) / sc.min_distance))
            if fb[i] == 0 or npnt <= 1:
                continue
            step = diff / npnt
            k = npnt - 1 if (fb[i] < 0 or fb[i] >= npnt) else fb[i] - 1
            conf[i] = a[i] + step * float(k)
        return w.check(a), ok, lok, conf

    def from_conf(self, w, sc, a, b, c):
        return w.check_conn_from_conf(a, b, c, sc.min_distance)

    def rrtstar(self, w, sc, samples):
        from graph_core_b200.planner import LockstepRRTStar
        from make_ref_golden import tree_edges_from_ids
        lock = LockstepRRTStar(sc, w, sc.start[None, :], sc.goal[None, :], capacity=len(samples) + 64, threads=1)
        used = lock.run(samples[:, None, :])
        info = lock.info(0)
        par, pc, conf = lock.dump(0)
        return used, info["solved"], info["cost"], info["tree_size"], tree_edges_from_ids(par, pc, conf)


def test_cuda_path_reproduces_reference_golden(gc, orc):
    from make_ref_golden import compute
    want = json.loads((GOLD / "ref_golden.json").read_text())
    want.pop("_about", None)
    got = compute(CudaBackend(gc, orc))
    for section in want:
        for key in want[section]:
            assert got[section][key] == want[section][key], (section, key)


def test_cuda_anytime_reproduces_reference_golden(gc, orc):
    """The lock-step AnytimeRRT (csrc/anytime.cu) against tests/golden/anytime_golden.json = round logs, costs and way
    points of the reference's own AnytimeRRT on the same per-problem sample streams (2-D boxes with connect / extend /
    solve() as written, 7-DoF sphere arm)."""
    import make_anytime_golden as m
    from graph_core_b200 import worlds as W
    from graph_core_b200.planner import LockstepAnytimeRRT
    want = json.loads((GOLD / "anytime_golden.json").read_text())
    for name, scn, n, max_iter, seed, extend, iis in m.CASES:
        case = want["cases"][name]
        sc, starts, goals, cost0 = m.problems(orc, W, scn, n)
        lk = LockstepAnytimeRRT(sc, sc.make_world(), starts, goals, max_iter=max_iter, seed=seed, extend=extend,
                                improve_in_solve=iis, sampler_cost0=cost0, capacity=8192)
        assert lk.run() == 0
        for p in range(n):
            w = case["results"][p]
            info = lk.info(p)
            assert info["solved"] == w["solved"], (name, p)
            assert m.hexf(lk.log(p)) == w["log"], (name, p)
            if w["solved"]:
                assert float(info["cost"]).hex() == w["cost"], (name, p)
                assert m.hexf(lk.solution(p)[0]) == w["solution"], (name, p)
