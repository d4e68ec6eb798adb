#!/usr/bin/env python
"""BASELINE config 5 shape: N independent AnytimeRRT queries on the 7-DoF sphere-arm world, problem j on GPU j mod G.

The planner is the REFERENCE's own, unmodified AnytimeRRT (oracle/_ref/libgc_ref_dropin.so); only its collision
checker is swapped: graph::core::CudaCollisionChecker (drop-in plugin, one gc_world_clone + stream per host thread)
versus the reference's per-state CPU loop over the same world.  One host thread per planner instance, T threads per
GPU; no collective on the data path, rank 0 gathers (solved, cost).  Test/measurement tool -- it loads oracle/.

  python tools/c5_anytime_dropin.py [--problems 256] [--threads 16] [--max-iter 2000] [--cpu-problems 32]
  torchrun --nproc-per-node G tools/c5_anytime_dropin.py ...
"""
import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))


def run_pool(solvers, max_iter, nthreads):
    done = [None] * len(solvers)

    def work(tid):
        for i in range(tid, len(solvers), nthreads):
            t0 = time.perf_counter()
            ok = solvers[i].solve(max_iter, 1.0e9)
            done[i] = (bool(ok), solvers[i].cost, time.perf_counter() - t0)

    ts = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return time.perf_counter() - t0, done


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problems", type=int, default=256)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--max-iter", type=int, default=2000)
    ap.add_argument("--cpu-problems", type=int, default=32)
    args = ap.parse_args()
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    import oracle_py as orc
    import ref_py as ref
    import graph_core_b200 as g
    from graph_core_b200 import worlds as W
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    torch.cuda.set_device(local_rank)
    ncores = len(os.sched_getaffinity(0))
    nthreads = args.threads or max(1, ncores // world)
    sc = W.scenario_c3(1000, orc.arm_checker(fast=True))
    pairs = W.c3_problems(orc.arm_checker(fast=True), sc.params, args.problems, seed=501)
    mine = list(range(rank, args.problems, world))
    root = sc.make_world(local_rank)
    q = W.uniform_samples(sc.lb, sc.ub, 5, 0, 4)  # warm the CUDA path before anything is timed
    root.check_connection(q[:2], q[2:], sc.min_distance)
    ref.lib(dropin=True).ref_srand(501 + rank)
    solvers = [ref.DropinSolver(ref.RefSolver.ANYTIME, sc, root.clone(), start=pairs[j][0], goal=pairs[j][1],
                                informed=True, cuda_nn=False, device=local_rank) for j in mine]
    dt, done = run_pool(solvers, args.max_iter, nthreads)
    calls = sum(s.device_calls for s in solvers)
    res = {"rank": rank, "n": len(mine), "seconds": dt, "solved": sum(1 for d in done if d[0]),
           "device_calls": calls, "mean_cost": float(np.mean([d[1] for d in done if d[0]] or [np.nan]))}
    if world > 1:
        allres = [None] * world
        dist.all_gather_object(allres, res)
    else:
        allres = [res]
    if rank == 0:
        tmax = max(r["seconds"] for r in allres)
        n = sum(r["n"] for r in allres)
        line = {"config": "C5: %d AnytimeRRT queries, 7-DoF sphere arm vs 1000 spheres, max_iter %d" % (n, args.max_iter),
                "planner": "graph_core AnytimeRRT (unmodified) + CudaCollisionChecker", "n_gpus": world,
                "host_threads_per_gpu": nthreads, "queries_per_s": n / tmax, "seconds": tmax,
                "solved": sum(r["solved"] for r in allres), "device_calls": sum(r["device_calls"] for r in allres)}
        # the same planner with the reference's CPU checking loop on a bounded sample of the same problems
        ncpu = min(args.cpu_problems, args.problems)
        ref.lib().ref_srand(501)
        cpu = []
        for j in range(ncpu):
            ow = orc.OracleWorld.from_scenario(sc, fast=True)
            cpu.append(ref.RefSolver(ref.RefSolver.ANYTIME, sc, ref.RefChecker.over_oracle_world(ow, sc.min_distance, fast=True),
                                     start=pairs[j][0], goal=pairs[j][1], informed=True, fast=True))
        dtc, donec = run_pool(cpu, args.max_iter, min(ncores, ncpu))
        line["cpu_reference"] = {"queries_per_s": ncpu / dtc, "problems": ncpu, "threads": min(ncores, ncpu),
                                 "solved": sum(1 for d in donec if d[0]), "seconds": dtc}
        line["speedup_vs_all_host_cores"] = line["queries_per_s"] / line["cpu_reference"]["queries_per_s"]
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
