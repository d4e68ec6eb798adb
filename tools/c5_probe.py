#!/usr/bin/env python
"""Quick timing probe of the lock-step AnytimeRRT on C5' (development tool)."""
import sys, time, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
import numpy as np
import oracle_py as orc
from graph_core_b200 import worlds as W
from graph_core_b200.planner import LockstepAnytimeRRT
P = int(sys.argv[1]) if len(sys.argv) > 1 else 512
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
arm = orc.arm_checker(fast=True)
sc = W.scenario_c3(1000, arm)
t0 = time.time()
starts, goals, cost0 = W.c5_problems(arm, sc.params, P)
t1 = time.time()
lk = LockstepAnytimeRRT(sc, sc.make_world(), starts, goals, max_iter=max_iter, seed=501, sampler_cost0=cost0, capacity=8192)
t2 = time.time()
max_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 0
live = lk.run(max_steps)
t3 = time.time()
solved, cost = lk.solved()
st = lk.stats()
print(json.dumps({"P": P, "max_iter": max_iter, "gen_s": t1 - t0, "create_s": t2 - t1, "run_s": t3 - t2, "live": live,
                  "queries_per_s": P / (t3 - t2), "solved": int(solved.sum()), "mean_cost": float(cost[solved].mean()) if solved.any() else None,
                  "stats": st, "timing": lk.timing(), "us_per_step": 1e6 * (t3 - t2) / max(st["steps"], 1)}))
