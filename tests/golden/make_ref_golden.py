"""Mints tests/golden/ref_golden.json from the REFERENCE ITSELF (oracle/_ref/libgc_ref.so: graph_core's own
sources compiled against the stand-in headers of oracle/ref_shim, see oracle/ref_build.py).

  python tests/golden/make_ref_golden.py        # needs /root/reference (or a prebuilt oracle/_ref)

tests/test_oracle_cpu.py::test_ref_golden_vectors recomputes the same quantities with the oracle
restatement and compares them with the committed file, so the restatement stays pinned to reference
outputs even where oracle/_ref is not available.  `compute(backend)` is shared by both sides.
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))


class RefBackend:
    """The reference's classes (ref_py) behind the few calls compute() makes."""
    name = "reference"

    def __init__(self, ref, orc):
        self.ref, self.orc = ref, orc

    def nn(self, kind, dim):
        return self.ref.RefNN(kind, dim)

    def checker(self, sc):
        if sc.kind == "cube":
            return self.ref.RefChecker.cube(sc.dim, sc.params["threshold"], sc.min_distance)
        # synthetic worlds are not in the reference: its checkConnection loops call the oracle's single-state check
        return self.ref.RefChecker.over_oracle_world(self.orc.OracleWorld.from_scenario(sc), sc.min_distance)

    def edges(self, chk, sc, a, b):
        ok = chk.check_connection(a, b)
        lok, conf = chk.check_connection(a, b, mode=1)
        return chk.check(a), ok, lok, conf

    def from_conf(self, chk, sc, a, b, c):
        return chk.check_conn_from_conf(a, b, c)

    def rrtstar(self, chk, sc, samples):
        s = self.ref.RefSolver(self.ref.RefSolver.RRTSTAR, sc, chk, informed=True)
        used = s.run(samples)
        return used, s.solved, s.cost, s.tree_size, self.ref.tree_as_edges(*s.dump())


class OracleBackend:
    """The restatement (oracle_py) behind the same calls."""
    name = "oracle"

    def __init__(self, ref, orc):
        self.ref, self.orc = ref, orc

    def nn(self, kind, dim):
        return self.orc.OracleNN(kind, dim)

    def checker(self, sc):
        return self.orc.OracleWorld.from_scenario(sc)

    def edges(self, w, sc, a, b):
        ok = w.check_connection(a, b, sc.min_distance)
        lok, _, conf = w.check_connection(a, b, sc.min_distance, mode=1)
        return w.check(a), ok, lok, conf

    def from_conf(self, w, sc, a, b, c):
        return w.check_conn_from_conf(a, b, c, sc.min_distance)

    def rrtstar(self, w, sc, samples):
        s = self.orc.OracleSolver(self.orc.OracleSolver.RRTSTAR, sc, w)
        used = s.run(samples)
        par, pc, conf = s.dump()
        return used, s.solved, s.cost, s.tree_size, tree_edges_from_ids(par, pc, conf)


def tree_edges_from_ids(parents, pcost, conf):
    out = {}
    for i in range(len(parents)):
        if parents[i] < 0 and i != 0:
            continue
        p = None if parents[i] < 0 else conf[parents[i]].tobytes()
        out[conf[i].tobytes()] = (p, float(pcost[i]) if parents[i] >= 0 else 0.0)
    return out


def _tree_digest(edges):
    h = hashlib.sha256()
    for k in sorted(edges):
        p, c = edges[k]
        h.update(k)
        h.update(p if p is not None else b"root")
        h.update(np.float64(c).tobytes())
    return h.hexdigest()


def _bits(a):
    return "".join(map(str, np.asarray(a).astype(int).tolist()))


def _hexf(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).ravel()]


def compute(be):
    from graph_core_b200 import worlds as W
    out = {}
    # ---- NN path: Vector and KdTree, 7-DoF (C3/C5 shape) and 12-DoF (C4 shape)
    out["nn"] = {}
    for kind in ("vector", "kdtree"):
        for dim, n in ((7, 3000), (12, 2000)):
            pts = W.c4_nodes(n, dim, seed=8100 + dim)
            qs = W.c4_queries(16, dim, seed=8200 + dim)
            s = be.nn(kind, dim)
            s.insert(pts)
            idx, dist = s.nearest(qs)
            rad = W.c4_radius(20, n, dim)
            ni, nd, nc = s.near(qs, rad, 256)
            ki, kd, kc = s.knn(qs, 6, 6)
            for v in (3, 10, 11):
                s.delete(v)
            idx2, _ = s.nearest(qs)
            out["nn"]["%s_d%d" % (kind, dim)] = {
                "nearest_idx": idx.tolist(), "nearest_dist": _hexf(dist), "near_count": nc.tolist(),
                "near_idx": [ni[i, :nc[i]].tolist() for i in range(len(qs))],
                "near_dist_head": [_hexf(nd[i, :min(2, nc[i])]) for i in range(len(qs))],
                "knn_idx": ki.tolist(), "knn_dist_last": _hexf(kd[:, -1]), "nearest_idx_after_delete": idx2.tolist(),
                "size_after_delete": int(s.size()),
            }
    # ---- edge path: every world; bisection, linear (+ last feasible conf), from-conf
    scen = {"C0": W.scenario_c0(), "C1": W.scenario_c1(), "C2": W.scenario_c2(),
            "C3": W.scenario_c3(1000, be.orc.arm_checker())}
    out["edge"] = {}
    for name, sc in scen.items():
        chk = be.checker(sc)
        n = 96 if name == "C3" else 384
        a = W.uniform_samples(sc.lb, sc.ub, 8300, 0, n)
        u = W.philox_u01(8301, 0, n, sc.dim + 2)
        dirn = u[:, :sc.dim] - 0.5
        dirn /= np.linalg.norm(dirn, axis=1, keepdims=True)
        b = np.clip(a + dirn * (u[:, sc.dim:sc.dim + 1] * 0.9), sc.lb, sc.ub)
        st, ok, lok, conf = be.edges(chk, sc, a, b)
        fc = be.from_conf(chk, sc, a, b, a + (b - a) * u[:, sc.dim + 1:])
        m = ~np.isnan(conf).any(axis=1)
        out["edge"][name] = {"state_ok": _bits(st), "bisect_ok": _bits(ok), "linear_ok": _bits(lok),
                             "linear_conf_set": _bits(m),
                             "linear_conf_sha256": hashlib.sha256(np.ascontiguousarray(conf[m]).tobytes()).hexdigest(),
                             "from_conf": _bits(fc)}
    # ---- planner level: RRT* with injected samples (rrt_star.cpp:89-163) in every world
    out["rrtstar"] = {}
    for name, iters in (("C0", 2500), ("C1", 2000), ("C2", 1000), ("C3", 400)):
        sc = scen[name]
        chk = be.checker(sc)
        samples = W.uniform_samples(sc.lb, sc.ub, 8400 + sc.dim, 0, iters)
        used, solved, cost, tree_size, edges = be.rrtstar(chk, sc, samples)
        out["rrtstar"][name] = {"used": int(used), "solved": bool(solved), "cost": float(cost).hex(),
                                "tree_size": int(tree_size), "tree_sha256": _tree_digest(edges)}
    return json.loads(json.dumps(out))


if __name__ == "__main__":
    import oracle_py
    import ref_py
    oracle_py.build()
    ref_py.build()
    g = compute(RefBackend(ref_py, oracle_py))
    g["_about"] = ref_py.lib().ref_about().decode()
    (HERE / "ref_golden.json").write_text(json.dumps(g, indent=1))
    print("wrote", HERE / "ref_golden.json")
