"""Mints tests/golden/oracle_golden.json from the oracle restatement (seeded Philox inputs).

These vectors come from the restatement itself and guard it against drift.  The vectors that pin the
restatement to the REFERENCE are tests/golden/ref_golden.json (make_ref_golden.py).  Regenerate with:  python tests/golden/make_golden.py
"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))


def compute(orc):
    from graph_core_b200 import worlds as W
    out = {}
    # ---- NN path: 7-DoF tree of 6000 nodes, 24 queries
    pts = W.c4_nodes(6000, 7, seed=4001)
    qs = W.c4_queries(24, 7, seed=4002)
    v = orc.OracleNN("vector", 7)
    v.insert(pts)
    idx, dist = v.nearest(qs)
    ni, nd, nc = v.near(qs, 2.2, 256)
    ki, kd, kc = v.knn(qs, 5, 5)
    out["nn"] = {
        "nearest_idx": idx.tolist(), "nearest_dist": dist.tolist(), "near_count": nc.tolist(),
        "near_first": [ni[i, :min(4, nc[i])].tolist() for i in range(len(qs))],
        "near_dist_sum": [float(nd[i, :nc[i]].sum()) for i in range(len(qs))],
        "knn_idx": ki.tolist(), "knn_dist": kd.tolist(),
    }
    # ---- edge path: every world, bisection + linear
    scen = {"C0": W.scenario_c0(), "C1": W.scenario_c1(), "C2": W.scenario_c2(),
            "C3": W.scenario_c3(1000, orc.arm_checker())}
    out["edge"] = {}
    for name, sc in scen.items():
        w = orc.OracleWorld.from_scenario(sc)
        n = 96 if name == "C3" else 256
        a = W.uniform_samples(sc.lb, sc.ub, 5000, 0, n)
        u = W.philox_u01(5001, 0, n, sc.dim + 1)
        dirn = u[:, :sc.dim] - 0.5
        dirn /= np.linalg.norm(dirn, axis=1, keepdims=True)
        b = np.clip(a + dirn * (u[:, sc.dim:] * 0.8), sc.lb, sc.ub)
        ok = w.check_connection(a, b, sc.min_distance)
        lok, fb, _ = w.check_connection(a, b, sc.min_distance, mode=1)
        st = w.check(a)
        out["edge"][name] = {"state_ok": "".join(map(str, st.tolist())), "bisect_ok": "".join(map(str, ok.tolist())),
                             "linear_ok": "".join(map(str, lok.tolist())), "first_bad": fb.tolist()}
    # ---- planner level: RRT* on the reference's own test scene with 1500 injected samples
    sc = scen["C0"]
    w = orc.OracleWorld.from_scenario(sc)
    s = orc.OracleSolver(orc.OracleSolver.RRTSTAR, sc, w)
    s.run(W.uniform_samples(sc.lb, sc.ub, 1, 0, 1500))
    par, pc, conf = s.dump()
    out["rrtstar_c0"] = {"solved": s.solved, "cost": s.cost, "tree_size": s.tree_size,
                         "parents_head": par[:64].tolist(), "solution": s.solution().tolist(),
                         "counters": s.counters()}
    return json.loads(json.dumps(out))


if __name__ == "__main__":
    import oracle_py
    oracle_py.build()
    g = compute(oracle_py)
    (HERE / "oracle_golden.json").write_text(json.dumps(g, indent=1))
    print("wrote", HERE / "oracle_golden.json")
