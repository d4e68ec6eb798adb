"""Mints tests/golden/anytime_golden.json from the REFERENCE ITSELF: graph_core's own AnytimeRRT (oracle/_ref, compiled
from /root/reference) driven through its public update / improve(max_iter = 0) / improveUpdate with the per-problem
Philox sample streams of the lock-step planner (sample k of problem p of P: counter k * P + p).

  python tests/golden/make_anytime_golden.py        # needs /root/reference (or a prebuilt oracle/_ref)

tests/test_oracle_cpu.py::test_anytime_golden recomputes the cases with the oracle restatement, tests/test_golden_gpu.py
with the CUDA path (LockstepAnytimeRRT); both compare bit for bit (floats are stored as hex strings).
"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))

CASES = [
    # name, scenario, problems, max_iter, seed, extend, improve_in_solve
    ("c1_connect", "C1", 4, 300, 7001, False, True),
    ("c1_extend", "C1", 3, 400, 7002, True, True),
    ("c1_solve_as_written", "C1", 2, 300, 7003, False, False),
    ("c5_sphere_arm", "C5", 4, 250, 501, False, True),
]


def hexf(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).ravel()]


def problems(orc, W, name, n):
    """(scenario, starts, goals, sampler_cost0) of a case; deterministic."""
    if name == "C5":
        arm = orc.arm_checker()
        sc = W.scenario_c3(1000, arm)
        starts, goals, cost0 = W.c5_problems(arm, sc.params, n)
        return sc, starts, goals, cost0
    sc = W.scenario_c1()
    ow = orc.OracleWorld.from_scenario(sc)
    u = W.philox_u01(7100, 0, 256, 4)
    pts = sc.lb + u.reshape(-1, 2) * (sc.ub - sc.lb)
    pts = pts[ow.check(pts).astype(bool)]
    starts = np.vstack([np.asarray(sc.start, float)[None, :], pts[:n - 1]])
    goals = np.vstack([np.asarray(sc.goal, float)[None, :], pts[n - 1:2 * (n - 1)]])
    return sc, starts, goals, np.full(n, np.inf)


def main():
    import oracle_py as orc
    import ref_py as ref
    from graph_core_b200 import worlds as W
    orc.build()
    if not ref.available():
        ref.build()
    out = {"_about": "reference AnytimeRRT (oracle/_ref) round logs and solutions; see make_anytime_golden.py",
           "columns": list(orc.ANYTIME_LOG_COLUMNS), "cases": {}}
    for name, scn, n, max_iter, seed, extend, iis in CASES:
        sc, starts, goals, cost0 = problems(orc, W, scn, n)
        lb, ub = np.asarray(sc.lb, float), np.asarray(sc.ub, float)
        rows = []
        for p in range(n):
            rc = ref.RefChecker.over_oracle_world(orc.OracleWorld.from_scenario(sc), sc.min_distance)
            r = ref.RefSolver(ref.RefSolver.ANYTIME, sc, rc, start=starts[p], goal=goals[p], extend=extend, informed=True)

            def sample_fn(k, cost, p=p):
                c = float(cost0[p]) if cost is None else cost
                q, _ = orc.sample_informed(lb, ub, starts[p][None, :], goals[p][None, :], np.array([c]), seed, k * n + p)
                return q[0]

            solved, log = ref.anytime_solve(r, sample_fn, max_iter, improve_in_solve=iis,
                                            utopia_tolerance=1.0 + sc.utopia_tolerance,
                                            best_utopia=float(np.sqrt(((goals[p] - starts[p]) ** 2).sum())))
            rows.append({"solved": bool(solved), "cost": float(r.cost).hex() if solved else None,
                         "log_shape": list(log.shape), "log": hexf(log),
                         "solution": hexf(r.solution()) if solved else []})
        out["cases"][name] = {"scenario": scn, "problems": n, "max_iter": max_iter, "seed": seed, "extend": extend,
                              "improve_in_solve": iis, "results": rows}
        print(name, [x["solved"] for x in rows], [x["log_shape"][0] for x in rows])
    (HERE / "anytime_golden.json").write_text(json.dumps(out, indent=0))


if __name__ == "__main__":
    main()
