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
inst = int(sys.argv[4]) if len(sys.argv) > 4 else 1
max_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 0
import threading
world = sc.make_world()
parts = [np.arange(i, P, inst) for i in range(inst)]
lks = [None] * inst
t2 = time.time()


lock = threading.Lock()


def work(i):
    idx = parts[i]
    with lock:
            lk = LockstepAnytimeRRT(sc, world, starts[idx], goals[idx], max_iter=max_iter, seed=501, sampler_cost0=cost0[idx],
                                capacity=8192)
    lk.set_streams(P, idx)
    lks[i] = lk
    lk.run(max_steps)


ts = [threading.Thread(target=work, args=(i,)) for i in range(inst)]
for t in ts:
    t.start()
for t in ts:
    t.join()
t3 = time.time()
solved = np.zeros(P, bool)
cost = np.zeros(P)
steps = 0
for i, lk in enumerate(lks):
    s_, c_ = lk.solved()
    solved[parts[i]] = s_
    cost[parts[i]] = c_
    steps = max(steps, lk.stats()["steps"])
print(json.dumps({"P": P, "max_iter": max_iter, "instances": inst, "gen_s": t1 - t0, "run_s": t3 - t2,
                  "queries_per_s": P / (t3 - t2), "solved": int(solved.sum()),
                  "mean_cost": float(cost[solved].mean()) if solved.any() else None, "steps": steps,
                  "stats0": lks[0].stats(), "timing0": lks[0].timing()}))
