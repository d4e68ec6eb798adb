#!/usr/bin/env python
"""Single-call latency of the C ABI's host-pointer entry points (what one graph_core planner sees through the
plugin classes): one state, one edge, one nearest / near query.  Prints one JSON line."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
import graph_core_b200 as g  # noqa: E402
from graph_core_b200 import worlds as W  # noqa: E402


def timeit(fn, n=2000, warm=200):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e6


def main():
    import oracle_py as orc
    sc = W.scenario_c3(1000, orc.arm_checker(fast=True))
    w = sc.make_world()
    a = W.uniform_samples(sc.lb, sc.ub, 3, 0, 64)
    b = np.clip(a + 0.25, sc.lb, sc.ub)
    out = {}
    out["check_state_us"] = timeit(lambda: w.check(a[:1]))
    out["check_edge_us"] = timeit(lambda: w.check_connection(a[:1], b[:1], sc.min_distance))
    out["check_edge_linear_us"] = timeit(lambda: w.check_connection(a[:1], b[:1], sc.min_distance, mode=1))
    out["check_64_edges_us"] = timeit(lambda: w.check_connection(a, b, sc.min_distance), n=500)
    st = g.NodeStore(7, 16384)
    st.append(W.c4_nodes(10000, 7, seed=1))
    q = W.c4_queries(1, 7, seed=2)
    out["nearest_10k_us"] = timeit(lambda: st.nearest(q))
    out["near_10k_us"] = timeit(lambda: st.near(q, 2.0, cap=256))
    out["append_1_us"] = timeit(lambda: st.append(q), n=1000)
    ow = orc.OracleWorld.from_scenario(sc, fast=True)
    out["cpu_check_edge_us"] = timeit(lambda: ow.check_connection(a[:1], b[:1], sc.min_distance), n=300, warm=20)
    print(json.dumps({k: round(v, 1) for k, v in out.items()}))


if __name__ == "__main__":
    main()
