#!/usr/bin/env python
"""bench.py -- RRT* samples/s on BASELINE.json config 3 (7-DoF sphere-model arm vs 1k-sphere cloud).

One "step" = one lock-step RRT* iteration over P independent planning problems (P samples), i.e. for
every problem: nearest neighbour + steer + extension edge check + radius-near + the rewiring edge
checks + append (RRTStar::update, src/graph_core/solvers/rrt_star.cpp:89-163).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--problems P] [--impl ours|reference]

  value  : samples/s with the sample blocks resident in HBM (written by the device Philox sampler)
  e2e    : samples/s through the host-facing call with HOST sample buffers (H2D of every step's samples,
           D2H of every step's results inside the timed region)
  roofline: the dominant kernel of the step, the arm edge-check kernel (FP32 pipe bound): executed
           sphere-pair tests x 10 flop + evaluated states x 1 kflop FK / device time of the check
           kernels (CUDA events on the launching stream, inside the timed region) against the FFMA peak measured in
           the same run; the executed-instruction view (9 flop per pair) is reported next to it
  cpu_baseline / --impl reference: the CPU restatement of the reference (oracle, the reference's
           -Ofast flags) running the same problems and sample streams on the host cores.

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "rrtstar_samples_per_sec"
UNIT = "samples/s"
# ALGORITHMIC work of the C3 state check as SURVEY.md 8(d) counts it: 10 flop per sphere pair (3 sub, 1 mul + 2 fma,
# radius add, square, compare) and ~1.0 kflop of forward kinematics per state.  The kernel EXECUTES less per pair
# (expanded form, include/gcb200_spec.h: 4 fused multiply-adds + 1 compare = 9 flop); both views are reported.
FLOP_PER_PAIR = 10.0
FLOP_FK_PER_STATE = 1000.0
EXECUTED_FLOP_PER_PAIR = 9.0


def edge_flops(wc):
    return FLOP_PER_PAIR * wc["pair_tests"] + FLOP_FK_PER_STATE * wc["states"]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--iters-per-step", type=int, default=50,
                    help="lock-step RRT* iterations per step (one step = this many samples per problem)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--problems", type=int, default=32768, help="independent planning problems per GPU")
    ap.add_argument("--threads", type=int, default=0, help="host replay threads (0 = all cores)")
    ap.add_argument("--instances", type=int, default=1,
                    help="lock-step planner instances per GPU, each with problems/instances problems, its own stream and "
                         "enqueueing host thread: the serial kernels of one instance overlap the wide ones of the others")
    ap.add_argument("--host-replay", action="store_true", help="round-1 path: trees on the host, replay on host threads")
    ap.add_argument("--skip-nn", action="store_true", help="skip the kNN measurements")
    ap.add_argument("--skip-extras", action="store_true", help="skip the informed-sampler arm and the depth sweep")
    ap.add_argument("--cpu-problems", type=int, default=0, help="problems of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-optimized", action="store_true", help="reference arm: skip the grid-culled CPU leg")
    ap.add_argument("--kernel-timing", action="store_true", help="with --profile: CUDA events around the check kernels")
    ap.add_argument("--profile", action="store_true",
                    help="profiling run: only the device-sample arm, no CPU baseline / parity / kNN side measurement")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ---- clocks ---------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.lines = []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def window(self, t0, t1):
        sm, smax, reasons = [], [], set()
        for t, line in self.lines:
            if t < t0 - 0.2 or t > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}

    def stop(self):
        if self.proc:
            self.proc.terminate()


# ---- workload -------------------------------------------------------------------------------------
def make_samples(sc, problems, iters, seed0=30200):
    from graph_core_b200 import worlds as W
    # sample of (iteration i, problem p): Philox key seed0, counter i*problems + p  -> [iters, problems, dim]
    return W.uniform_samples(sc.lb, sc.ub, seed0, 0, iters * problems).reshape(iters, problems, sc.dim)


def cuda_checker(device=0, world_cache={}):
    """State checker backed by the CUDA world (product path) for problem selection."""
    import graph_core_b200 as g

    def chk(params, q):
        key = (id(params["obs"]), device)
        if key not in world_cache:
            world_cache[key] = g.CollisionWorld.sphere_arm(params["base_tf"], params["joint_tf"], params["rs_link"],
                                                           params["rs_local"], params["obs"], device)
        return world_cache[key].check(q).astype(bool)

    return chk


def cpu_reference_run(sc, starts, goals, samples, warmup, steps, threads, prefer_ref=True, grid=False):
    """The reference's CPU RRT* on len(starts) problems, `threads` host threads; returns samples/s.

    kind "reference": oracle/_ref -- graph_core's own RRTStar / Tree / KdTree / checkConnection sources compiled with
    its Release flags (oracle/ref_build.py), single-state checks answered by the oracle's sphere-arm world (the
    reference ships no such checker).  kind "port": the oracle restatement, used only when oracle/_ref is absent."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle_py as orc
    import ref_py
    orc.build()
    kind = "port"
    if prefer_ref:
        try:
            ref_py.lib(fast=True)
            kind = "reference"
        except Exception:
            kind = "port"
    n = len(starts)
    solvers = []
    worlds_ = []
    for p in range(n):
        w = orc.OracleWorld.from_scenario(sc, fast=True)  # one world per solver: counters are not shared
        if grid:
            w.enable_grid(True)  # the uniform-grid culling the CUDA path uses (same verdicts, ~100x fewer pair tests)
        worlds_.append(w)
        if kind == "reference":
            chk = ref_py.RefChecker.over_oracle_world(w, sc.min_distance, fast=True)
            solvers.append(ref_py.RefSolver(ref_py.RefSolver.RRTSTAR, sc, chk, start=starts[p], goal=goals[p],
                                            use_kdtree=True, informed=True, fast=True))
        else:
            solvers.append(orc.OracleSolver(orc.OracleSolver.RRTSTAR, sc, w, start=starts[p], goal=goals[p],
                                            use_kdtree=True, fast=True))
    nthreads = max(1, min(threads, n))
    done = [0] * n

    def work(tid, lo, hi):
        for p in range(tid, n, nthreads):
            done[p] += solvers[p].run(samples[lo:hi, p, :])

    def phase(lo, hi):
        ts = [threading.Thread(target=work, args=(t, lo, hi)) for t in range(nthreads)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0

    phase(0, warmup)
    before = sum(done)
    dt = phase(warmup, warmup + steps)
    consumed = sum(done) - before
    return consumed / dt, dt, nthreads, solvers, kind


CPU_DESC = {
    "reference": "graph_core's own RRTStar + KdTree + checkConnection sources (oracle/_ref, -O3 -Ofast, stand-in Eigen), "
                 "single-state sphere-arm check from the oracle",
    "port": "CPU restatement of KdTree + checkConnection (-O3 -Ofast)",
}



def depth_sweep(g, torch, LockstepRRTStar, sc, root_world, starts, goals, dev, targets=(100, 1000, 10000), problems=1024,
                window=100, seed=31500, budget_s=90.0):
    """samples/s of the lock-step step at given tree sizes: `problems` problems are grown with uniform Philox samples
    until their mean tree size reaches each target, then `window` iterations are timed with CUDA events."""
    import ctypes as C
    P = min(problems, len(starts))
    cap = int(max(targets) * 1.25) + 512
    w = root_world.clone()
    st = torch.cuda.Stream()
    w.set_stream(st.cuda_stream)
    lk = LockstepRRTStar(sc, w, starts[:P], goals[:P], capacity=cap, device=dev, near_cap=4096)
    lk.set_stream(st.cuda_stream)
    lib = g.capi.load()
    lb = np.ascontiguousarray(sc.lb, dtype=np.float64)
    ub = np.ascontiguousarray(sc.ub, dtype=np.float64)
    chunk = 256
    buf = torch.empty((chunk, P, sc.dim), dtype=torch.float64, device="cuda")
    state = {"it": 0}

    def run(n):
        done = 0
        while done < n:
            m = min(chunk, n - done)
            rc = lib.gc_sample_uniform_dev(dev, sc.dim, lb.ctypes.data_as(C.POINTER(C.c_double)),
                                           ub.ctypes.data_as(C.POINTER(C.c_double)), seed, state["it"] * P, m * P,
                                           buf.data_ptr(), C.c_void_p(st.cuda_stream))
            assert rc == 0
            with torch.cuda.stream(st):
                for j in range(m):
                    lk.step_device_samples(buf[j].data_ptr())
            lk.sync()  # the block is refilled next: wait for its readers
            state["it"] += m
            done += m

    def mean_nodes():
        return float(np.mean([lk.info(p)["tree_size"] for p in range(0, P, max(1, P // 64))]))

    out = []
    t_start = time.time()
    for target in targets:
        while mean_nodes() < target and time.time() - t_start < budget_s:
            run(chunk)
        if mean_nodes() < target * 0.8:
            out.append({"target_nodes": target, "skipped": "time budget of the sweep (%.0f s) reached at %.0f nodes" %
                        (budget_s, mean_nodes())})
            break
        n0 = mean_nodes()
        s0 = lk.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        run(window)
        e1.record(st)
        torch.cuda.synchronize()
        s1 = lk.stats()
        ms = e0.elapsed_time(e1)
        out.append({"tree_nodes": [n0, mean_nodes()], "problems": P, "iterations": window,
                    "samples_per_s": (s1["iterations"] - s0["iterations"]) / (ms * 1e-3), "ms_per_iteration": ms / window,
                    "near_hits_per_extension": (s1["near_hits"] - s0["near_hits"]) / max(s1["extended"] - s0["extended"], 1)})
    lk.close()
    return out


def informed_arm(g, torch, LockstepRRTStar, sc, root_world, starts, goals, dev, iters_warm, iters_timed, seed=30900):
    """The same problems with the planner's own sampler: the device informed sampler draws sample i of problem p from the
    informed set of p's current path cost (InformedSampler::sample, rrt_star.cpp:147-161); nothing crosses PCIe."""
    P = len(starts)
    w = root_world.clone()
    st = torch.cuda.Stream()
    w.set_stream(st.cuda_stream)
    lk = LockstepRRTStar(sc, w, starts, goals, capacity=iters_warm + iters_timed + 64, device=dev)
    lk.set_stream(st.cuda_stream)
    with torch.cuda.stream(st):
        for i in range(iters_warm):
            lk.step_informed(seed, i * P)
        lk.sync()
        s0 = lk.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(iters_warm, iters_warm + iters_timed):
            lk.step_informed(seed, i * P)
        e1.record(st)
    torch.cuda.synchronize()
    lk.sync()
    s1 = lk.stats()
    ms = e0.elapsed_time(e1)
    solved = sum(lk.info(p)["solved"] for p in range(0, P, max(1, P // 256))) / len(range(0, P, max(1, P // 256)))
    lk.close()
    return {"value": (s1["iterations"] - s0["iterations"]) / (ms * 1e-3), "unit": UNIT, "ms_per_iteration": ms / iters_timed,
            "problems": P, "iterations": iters_timed, "solved_fraction_at_end": solved,
            "sampler": "device informed sampler inside the forest (Philox key %d, counter iteration*P + problem)" % seed,
            "parity": "tests/test_planner_gpu.py::test_rrtstar_informed_sampler_inside_the_planner (oracle fed the same "
                      "stream, trees bit-exact)"}


def knn_c4(g, torch, dist_mod, rank, world, dev, hbm_peak, fp32_peak, n_nodes=1 << 20, dim=12, reps=5):
    """BASELINE config 4: 2^20-node 12-DoF tree, node-sharded round-robin over the ranks (NCCL all-gather of the per-shard
    lists + device merge); queries/s = Q / max-over-ranks device time, L2 not flushed (the shard is larger than L2 only at
    one rank: 50 MB of fp32 SoA per 2^20 nodes)."""
    from graph_core_b200 import worlds as W
    from graph_core_b200.sharded import ShardedNodeStore
    pts = W.c4_nodes(n_nodes, dim)
    qs = W.c4_queries(4096, dim)
    st = ShardedNodeStore(dim, n_nodes, device=dev, exchange="nccl", nq_max=4096, cap_max=1024)
    st.append(pts)
    torch.cuda.synchronize()

    def timed(fn):
        times = []
        for it in range(reps + 2):
            if world > 1:
                dist_mod.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], dtype=torch.float64, device="cuda")
                dist_mod.all_reduce(t, op=dist_mod.ReduceOp.MAX)
                ms = float(t.item())
            if it >= 2:
                times.append(ms)
        return float(np.median(times))

    res = {"nodes": n_nodes, "dim": dim, "n_gpus": world, "sharding": "round-robin over ranks, NCCL all-gather + merge kernel",
           "cases": []}
    for nq in (4096, 1):
        for k in (1, 16, 64):
            q = qs[:nq]
            dq = torch.from_numpy(np.ascontiguousarray(q)).cuda()
            ms = timed((lambda: st.nearest(dq)) if k == 1 else (lambda: st.knn(dq, k)))  # queries resident in HBM
            ms_host = timed((lambda: st.nearest(q)) if k == 1 else (lambda: st.knn(q, k)))  # host queries, H2D inside
            case = {"queries": nq, "k": k, "ms": ms, "queries_per_s": nq / (ms * 1e-3),
                    "e2e_ms_host_queries": ms_host, "e2e_queries_per_s": nq / (ms_host * 1e-3)}
            if nq == 1:
                # one pass over the fp32 SoA of every shard: HBM-bound regime
                case["GBps_aggregate"] = 4.0 * dim * n_nodes / (ms * 1e-3) / 1e9
                case["frac_of_hbm_peak"] = case["GBps_aggregate"] / (hbm_peak * world)
            else:
                case["pairs_per_s"] = nq * float(n_nodes) / (ms * 1e-3)
                # expanded form: D fused multiply-adds per (node, query) pair
                case["frac_of_fp32_peak"] = case["pairs_per_s"] * 2.0 * dim / 1e12 / (fp32_peak * world) if fp32_peak else None
            res["cases"].append(case)
    head = next(c for c in res["cases"] if c["queries"] == 4096 and c["k"] == 16)
    res["queries_per_s"] = head["queries_per_s"]
    res["headline_case"] = "k = 16, 4096 queries"
    lat = next(c for c in res["cases"] if c["queries"] == 1 and c["k"] == 1)
    res["roofline"] = {"kernel": "scan_stream_kernel<12,nearest>", "bound": "hbm", "achieved": lat["GBps_aggregate"],
                       "peak": hbm_peak * world, "unit": "GB/s", "frac": lat["frac_of_hbm_peak"],
                       "what": "nearest neighbour of ONE query over the whole sharded tree (4*D*N bytes / max-over-ranks "
                               "time, L2-resident after the first repetition at this size)"}
    return res


def c5_leg(g, torch, dist_mod, rank, world, dev, sc, root_world, ncores, total=4096, max_iter=2000, seed=501,
           instances=4):
    """BASELINE config 5: `total` independent AnytimeRRT queries in the C3 world, query j on GPU j mod G (no data-path
    collective; a query's sample stream depends on j only, so the results do not depend on G).  The queries are C5'
    (graph_core_b200/worlds.py c5_problems: goal 1.5 rad from the start, informed first phase) because with BASELINE's
    literal C5 the reference finds no first solution and its improve rounds never start.  Timed region: creation
    (TreeSolver::setProblem rays) + AnytimeRRT::solve with the improve loop for every query of this rank; whole-job
    queries/s = total / max over ranks.  Rank 0 replays a few queries through the oracle (bit-identical round logs) and
    times the reference's own AnytimeRRT (oracle/_ref) on the host cores beside it."""
    from graph_core_b200 import worlds as W
    from graph_core_b200.planner import LockstepAnytimeRRT, query_shard
    starts, goals, cost0 = W.c5_problems(cuda_checker(dev), sc.params, total, seed=seed)
    mine = query_shard(total, rank, world)
    import threading
    # `instances` lock-step planners per GPU, each with its own stream and host thread: a step is a chain of short
    # launches and read-backs, so the instances overlap one another's latencies (results do not depend on the split)
    parts = [mine[i::instances] for i in range(instances)]
    lks = [None] * instances
    errs = []
    create_lock = threading.Lock()  # gc_world_clone of one parent handle is not meant to run concurrently

    def work(i):
        try:
            with create_lock:
                lk_i = LockstepAnytimeRRT(sc, root_world, starts[parts[i]], goals[parts[i]], max_iter=max_iter,
                                          seed=seed, sampler_cost0=cost0[parts[i]], capacity=8192)
            lk_i.set_streams(total, parts[i])
            lks[i] = lk_i
            lk_i.run()
        except Exception as ex:  # noqa: BLE001
            errs.append(str(ex))

    if world > 1:
        dist_mod.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ts = [threading.Thread(target=work, args=(i,)) for i in range(instances)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if errs:
        raise RuntimeError(errs[0])
    if world > 1:
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist_mod.all_reduce(tt, op=dist_mod.ReduceOp.MAX)
        dt_all = float(tt.item())
    else:
        dt_all = dt
    solved = np.zeros(len(mine), bool)
    cost = np.zeros(len(mine))
    st = {k: 0 for k in LockstepAnytimeRRT.STATS}
    local_of = {}
    for i, lk_i in enumerate(lks):
        s_i, c_i = lk_i.solved()
        solved[i::instances] = s_i
        cost[i::instances] = c_i
        for k, v in lk_i.stats().items():
            st[k] = max(st[k], v) if k == "steps" else st[k] + v
        for j in range(len(parts[i])):
            local_of[i + j * instances] = (lk_i, j)
    live = 0
    agg = torch.tensor([float(solved.sum()), float(cost[solved].sum()), float(st["samples"]), float(st["steps"]),
                        float(st["improve_edges"] + st["goal_edges"] + st["rrt_rows"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist_mod.all_reduce(agg)
    out = {"queries_per_s": total / dt_all, "unit": "queries/s", "queries": total, "queries_per_gpu": len(mine),
           "max_iter_per_round": max_iter, "seconds": dt_all, "live_after_run": int(live),
           "solved": int(agg[0].item()), "mean_cost": float(agg[1].item() / max(agg[0].item(), 1.0)),
           "planner_updates": int(agg[2].item()), "lockstep_steps_rank0": st["steps"],
           "edges_checked": int(agg[4].item()), "scaling": "strong (BASELINE: the same 4 096 queries over 1/2/4/8 GPUs)",
           "workload": "C5': C3 world, start uniform (seed 501), goal 1.5 rad away, InformedSampler(cost 1.5 |goal - "
                       "start|) in the first phase, delta 0.1, cost_impr 0.1, improve loop run (improve_in_solve)",
           "instances_per_gpu": instances, "host_seconds_rank0_instance0": lks[0].timing()}
    if rank != 0:
        return out
    sys.path.insert(0, str(ROOT / "oracle"))
    import ctypes as C
    import threading
    import oracle_py as orc
    orc.build()
    # parity: a few queries of this rank through the oracle's restatement (pinned to the reference's AnytimeRRT)
    ok_all, checked = True, 0
    for i in range(0, len(mine), max(1, len(mine) // 4))[:4]:
        j = int(mine[i])
        o = orc.OracleAnytime(sc, orc.OracleWorld.from_scenario(sc), start=starts[j], goal=goals[j],
                              sampler_cost0=float(cost0[j]))
        so = o.solve(max_iter, seed, j, total)
        lk, li = local_of[i]
        ok = bool(so == bool(solved[i]) and np.array_equal(o.log(), lk.log(li)))
        if so:
            ok = ok and bool(np.array_equal(o.solution(), lk.solution(li)[0]) and o.state()["cost"] == cost[i])
        ok_all = ok_all and ok
        checked += 1
    out["parity"] = {"round_logs_paths_costs_bit_exact": ok_all, "queries_checked": checked}
    # CPU arm: the reference's own AnytimeRRT, its state check answered by the oracle world with the grid culling
    try:
        import ref_py
        L = ref_py.lib(fast=True)
        fn = C.cast(orc.lib(fast=True).orc_sample_informed, C.c_void_p)
        ncpu = min(total, 2 * ncores)
        res = [None] * ncpu

        def work(tid):
            for j in range(tid, ncpu, ncores):
                ow = orc.OracleWorld.from_scenario(sc, fast=True)
                ow.enable_grid(True)
                rc = ref_py.RefChecker.over_oracle_world(ow, sc.min_distance, fast=True)
                r = ref_py.RefSolver(ref_py.RefSolver.ANYTIME, sc, rc, start=starts[j], goal=goals[j], informed=True,
                                     fast=True)
                res[j] = r.anytime_solve_injected(max_iter, fn, sc.lb, sc.ub, float(cost0[j]), seed, j, total,
                                                  sc.utopia_tolerance)

        ts = [threading.Thread(target=work, args=(t,)) for t in range(min(ncores, ncpu))]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        dtc = time.perf_counter() - t0
        out["cpu_baseline_optimized"] = {
            "value": ncpu / dtc, "unit": "queries/s", "cores": min(ncores, ncpu), "kind": "reference",
            "sample": "the first %d queries, graph_core's own AnytimeRRT (update / improve / improveUpdate, oracle/_ref, "
                      "Release flags) on the same sample streams, its state check answered by the oracle world WITH the "
                      "CUDA path's grid culling (the plain per-state loop is ~20x slower), one query per thread, %.1f s"
                      % (ncpu, dtc),
            "solved": int(sum(1 for r in res if r and r[0])), "updates": int(sum(r[2] for r in res if r)),
            "gpu_over_cpu": out["queries_per_s"] / (ncpu / dtc)}
    except Exception as ex:
        out["cpu_baseline_optimized"] = {"error": str(ex)[:300]}
    return out


def c2_legs(g, torch, dev, ncores, problems=16384, iters=64):
    """BASELINE config 2 (rank 0): (i) the 2^20-segment edge micro-benchmark of SURVEY 8(d) (tools/edge_bench.py) with the
    reference's checkConnection loop timed beside it on one host core; (ii) lock-step BiRRT (graph_core_b200/csrc/birrt.cu):
    `problems` independent 6-DoF problems (the 256 start/goal pairs of seed 202, replicated with their own sample streams),
    `iters` updates each or until solved, samples resident in HBM; a few problems are replayed through the reference's own
    BiRRT (oracle/_ref) on the same samples as the in-bench parity check."""
    sys.path.insert(0, str(ROOT / "tools"))
    import ctypes as C
    import edge_bench
    from graph_core_b200 import worlds as W
    from graph_core_b200.planner import LockstepBiRRT
    out = {}
    eb, h1, h2, free = edge_bench.gpu_legs()
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle_py as orc
    orc.build()
    sc = W.scenario_c2()
    ow = orc.OracleWorld.from_scenario(sc, fast=True)
    m = 20000
    t0 = time.perf_counter()
    cok = ow.check_connection(h1[:m], h2[:m], 0.01, 0)
    dt = time.perf_counter() - t0
    eb["parity"] = {"verdicts_equal_to_oracle": bool(np.array_equal(np.asarray(cok).astype(bool), free[:m])), "edges": m}
    eb["cpu_baseline"] = {"value": m / dt, "unit": "edges/s", "cores": 1, "kind": "port",
                          "sample": "first %d segments of the batch, the restatement of checkConnection "
                                    "(collision_checker_base.h:207-237) with the reference's Release flags, one host core "
                                    "(the reference is single-threaded), %.2f s" % (m, dt)}
    out["edge_microbench"] = eb
    # ---- lock-step BiRRT ---------------------------------------------------------------------------------------
    npairs = 256
    u = W.philox_u01(202, 0, 4 * npairs, 2 * sc.dim)
    s_all = sc.lb + u[:, :sc.dim] * (sc.ub - sc.lb)
    g_all = sc.lb + u[:, sc.dim:] * (sc.ub - sc.lb)
    w = g.CollisionWorld.cspheres(sc.params["centers"], sc.params["radii"], dev)
    okp = w.check(s_all).astype(bool) & w.check(g_all).astype(bool)
    s_all, g_all = s_all[okp][:npairs], g_all[okp][:npairs]
    P = problems
    starts = np.tile(s_all, (P // len(s_all) + 1, 1))[:P]
    goals = np.tile(g_all, (P // len(g_all) + 1, 1))[:P]
    st = torch.cuda.Stream()
    w.set_stream(st.cuda_stream)
    lk = LockstepBiRRT(sc, w, starts, goals, capacity=2048, bidirectional=True, extend=False)
    d_samples = torch.empty((iters, P, sc.dim), dtype=torch.float64, device="cuda")
    lb64 = np.ascontiguousarray(sc.lb, dtype=np.float64)
    ub64 = np.ascontiguousarray(sc.ub, dtype=np.float64)
    dp = C.POINTER(C.c_double)
    rc = g.capi.load().gc_sample_uniform_dev(dev, sc.dim, lb64.ctypes.data_as(dp), ub64.ctypes.data_as(dp), 20300, 0,
                                             iters * P, d_samples.data_ptr(), None)
    assert rc == 0
    torch.cuda.synchronize()
    solved0, _ = lk.solved()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        for i in range(iters):
            lk.step_device_samples(d_samples[i].data_ptr())
        e1.record(st)
    torch.cuda.synchronize()
    lk.sync()
    ms = e0.elapsed_time(e1)
    stt = lk.stats()
    solved, costs = lk.solved()
    bi = {"problems": P, "iterations": iters, "ms": ms, "updates": stt["updates"],
          "updates_per_s": stt["updates"] / (ms * 1e-3), "ms_per_iteration": ms / iters,
          "solved_by_direct_connection": float(solved0.mean()), "solved_fraction": float(solved.mean()),
          "nodes_added": stt["nodes_added"], "edges_checked": stt["edges"], "states_evaluated": stt["states"],
          "launches_per_iteration": 1,
          "what": "one kernel launch per BiRRT::update of every unsolved problem (birrt.cpp:93-152), one warp per tree, "
                  "the whole Tree::connect ray inside the kernel; solved problems idle"}
    # parity + CPU arm on the reference's own BiRRT (same samples)
    try:
        import ref_py
        hs = d_samples.cpu().numpy()
        npar, ncpu = 8, 512
        same = True
        t0 = time.perf_counter()
        upd = 0
        ow_fast = orc.OracleWorld.from_scenario(sc, fast=True)
        for jj, p in enumerate([P * j // ncpu for j in range(ncpu)]):
            chk = ref_py.RefChecker.over_oracle_world(ow_fast, sc.min_distance, fast=True)
            cpu = ref_py.RefSolver(ref_py.RefSolver.BIRRT, sc, chk, start=starts[p], goal=goals[p], informed=False, fast=True)
            solved_directly = cpu.solved
            used = 0 if solved_directly else cpu.run(hs[:, p])
            upd += used if cpu.solved else iters
            if jj % (ncpu // npar) != 0:
                continue
            info = lk.info(p)
            ok = info["solved"] == cpu.solved
            if ok and cpu.solved and info["solved_iteration"] > 0:
                wp, _ = lk.solution(p)
                ok = info["solved_iteration"] == used and abs(info["cost"] - cpu.cost) <= 1e-9 * cpu.cost and \
                    np.allclose(wp, cpu.solution(), rtol=0, atol=1e-12)
            same = same and ok
        dt = time.perf_counter() - t0
        bi["parity"] = {"same_as_reference_birrt": bool(same), "problems_checked": npar,
                        "note": "timing build of the reference (-Ofast): iteration, cost and way points compared with "
                                "1e-9 / 1e-12 tolerances here, bit-exact against the parity build in tests/test_birrt_gpu.py"}
        bi["cpu_baseline"] = {"value": upd / dt, "unit": "updates/s", "cores": 1, "kind": "reference",
                              "sample": "%d of the problems through graph_core's own BiRRT (oracle/_ref, Release flags; set-up "
                                        "with its direct-connection attempt included), one host core, %.2f s" % (ncpu, dt)}
    except Exception as ex:
        bi["parity"] = {"error": str(ex)[:200]}
    lk.close()
    out["birrt"] = bi
    return out


def main():
    args = parse_args()
    rank, local_rank, world_size = dist_env()
    if world_size > 1 and args.gpus != world_size:
        args.gpus = world_size
    W_, K = args.warmup, args.steps
    from graph_core_b200 import worlds as W

    ncores = os.cpu_count() or 1
    try:
        ncores = len(os.sched_getaffinity(0)) or ncores
    except Exception:
        pass
    ncores_rank = max(1, ncores // max(world_size, 1))  # the ranks of one box share its host cores

    # ------------------------------------------------------------------ reference arm (CPU only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        sys.path.insert(0, str(ROOT / "oracle"))
        import oracle_py as orc
        orc.build()
        sc = W.scenario_c3(1000, orc.arm_checker(fast=True))
        I = max(1, args.iters_per_step)
        nprob = args.cpu_problems or min(args.problems, max(ncores, 1) * 2)
        pairs = W.c3_problems(orc.arm_checker(fast=True), sc.params, nprob)
        starts = np.array([s for s, _ in pairs])
        goals = np.array([g for _, g in pairs])
        samples = make_samples(sc, nprob, (W_ + K) * I)
        rate, dt, nthreads, _, kind = cpu_reference_run(sc, starts, goals, samples, W_ * I, K * I, ncores)
        line = {
            "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
            "warmup": W_, "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64 planner / f32 collision geometry", "data": "synthetic",
            "config": {"workload": "C3: 7-DoF sphere-arm RRT* vs 1000-sphere cloud, %d problems; one step = %d RRT* "
                                   "iterations of every problem (%d samples)" % (nprob, I, nprob * I),
                       "problems": nprob, "iterations_per_step": I, "max_distance": sc.max_distance,
                       "min_distance": sc.min_distance, "sampler": "uniform (Philox, injected)"},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": nthreads, "kind": kind,
                             "sample": "%d problems x %d RRT* iterations after %d warm-up iterations, %s, one problem per "
                                       "thread, %.1f s" % (nprob, K * I, W_ * I, CPU_DESC[kind], dt)},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        if not args.no_cpu_optimized:
            nprob2 = args.cpu_problems or min(args.problems, max(ncores, 1) * 8)
            pairs2 = W.c3_problems(orc.arm_checker(fast=True), sc.params, nprob2)
            samples2 = make_samples(sc, nprob2, (W_ + K) * I)
            rate2, dt2, nt2, _, kind2 = cpu_reference_run(sc, np.array([s for s, _ in pairs2]),
                                                          np.array([g for _, g in pairs2]), samples2, W_ * I, K * I,
                                                          ncores, grid=True)
            line["cpu_baseline_optimized"] = {
                "value": rate2, "unit": UNIT, "cores": nt2, "kind": kind2,
                "sample": "%d problems x %d RRT* iterations, the same planner sources with the sphere-arm state check given "
                          "the uniform-grid culling of the CUDA path (verdicts identical), %.1f s" % (nprob2, K * I, dt2)}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (CUDA)
    import torch

    import graph_core_b200 as g
    from graph_core_b200.planner import LockstepRRTStar

    if not torch.cuda.is_available() or g.device_count() < 1:
        raise RuntimeError("bench.py: no CUDA device; graph_core_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = local_rank
    if world_size > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    P = args.problems
    I = max(1, args.iters_per_step)
    sc = W.scenario_c3(1000, cuda_checker(dev))
    # rank r plans problems [r*P, (r+1)*P) of the shared stream: independent queries, no data-path collective
    pairs = W.c3_problems(cuda_checker(dev), sc.params, P * world_size)[rank * P:(rank + 1) * P]
    starts = np.array([s for s, _ in pairs])
    goals = np.array([gl for _, gl in pairs])
    iters = (W_ + K) * I
    stream = torch.cuda.Stream()
    # sample of (iteration i, problem p): Philox key 30200 + rank, counter i*P + p.  The blocks are written by the device
    # Philox sampler outside the timed region; the host copy the e2e run and the CPU baseline consume is the same stream
    # (a prefix is checked against the host implementation of the sampler)
    d_samples = torch.empty((iters, P, sc.dim), dtype=torch.float64, device="cuda")
    lb64 = np.ascontiguousarray(sc.lb, dtype=np.float64)
    ub64 = np.ascontiguousarray(sc.ub, dtype=np.float64)
    import ctypes as C
    rc = g.capi.load().gc_sample_uniform_dev(dev, sc.dim, lb64.ctypes.data_as(C.POINTER(C.c_double)),
                                             ub64.ctypes.data_as(C.POINTER(C.c_double)), 30200 + rank, 0, iters * P,
                                             d_samples.data_ptr(), None)
    assert rc == 0, g.capi.load().gc_last_error()
    torch.cuda.synchronize()
    h_samples = torch.empty((iters, P, sc.dim), dtype=torch.float64).pin_memory()
    h_samples.copy_(d_samples)
    samples = h_samples.numpy()
    npre = min(iters, max(1, 65536 // P))
    assert np.array_equal(samples[:npre], make_samples(sc, P, npre, seed0=30200 + rank)), \
        "device Philox stream differs from the host stream"

    clocks = ClockSampler(dev)
    fp32_peak, fp32_mhz = g.capi.measure_fp32_peak(dev)

    root_world = sc.make_world(dev)
    device_trees = not args.host_replay

    def timed_run(device_samples: bool, instances: int, steps_warm=W_, steps_timed=K, kernel_timing=False):
        """(steps_warm + steps_timed) x I lock-step iterations of all P problems, split over `instances` planner instances."""
        G = max(1, min(instances, P))
        bounds = [P * i // G for i in range(G + 1)]
        thr_each = max(1, (args.threads or ncores_rank) // G)
        inst = []
        for i in range(G):
            lo, hi = bounds[i], bounds[i + 1]
            st_i = torch.cuda.Stream()
            w_i = root_world.clone()
            w_i.set_stream(st_i.cuda_stream)
            lk = LockstepRRTStar(sc, w_i, starts[lo:hi], goals[lo:hi], capacity=iters + 64, device=dev, threads=thr_each,
                                 device_trees=device_trees)
            lk.set_stream(st_i.cuda_stream)
            x = {"lo": lo, "hi": hi, "stream": st_i, "world": w_i, "lock": lk, "d": None, "h": None,
                 "costs": torch.empty(hi - lo, dtype=torch.float64).pin_memory()}
            if device_samples:
                x["d"] = d_samples if G == 1 else d_samples[:, lo:hi].contiguous()
            else:
                hs = h_samples if G == 1 else h_samples[:, lo:hi].contiguous().pin_memory()
                x["h"] = hs
                x["hn"] = hs.numpy()
            inst.append(x)

        def run_steps(x, first, last):
            lk = x["lock"]
            with torch.cuda.stream(x["stream"]):
                for s_ in range(first, last):
                    for i2 in range(s_ * I, (s_ + 1) * I):
                        if device_samples:
                            lk.step_device_samples(x["d"][i2].data_ptr())
                        else:
                            lk.step(x["hn"][i2])
                    if device_trees and not device_samples:
                        # the step's result goes back to the host: current path cost of every problem
                        lk.read_costs_async(x["costs"].data_ptr())
                        lk.sync()

        def run_all(first, last):
            if G == 1:
                run_steps(inst[0], first, last)
                return
            ts = [threading.Thread(target=run_steps, args=(x, first, last)) for x in inst]
            for t in ts:
                t.start()
            for t in ts:
                t.join()

        run_all(0, steps_warm)
        for x in inst:
            x["lock"].sync()
            # per-call CUDA events around the check kernels: only in the roofline run (they keep the step out of its
            # CUDA graph)
            x["world"].enable_timing(kernel_timing)
            x["world"].reset_counters()
        st0 = [x["lock"].stats() for x in inst]
        nodes0 = float(np.mean([x["lock"].info(p)["tree_size"] for x in inst for p in range(0, x["hi"] - x["lo"], 97)]))
        l0 = g.capi.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if args.profile:
            torch.cuda.profiler.start()  # ncu --profile-from-start off: only the timed steps are captured
        t0 = time.time()
        e0.record(stream)
        for x in inst:
            x["stream"].wait_event(e0)
        run_all(steps_warm, steps_warm + steps_timed)
        for x in inst:
            ev = torch.cuda.Event()
            ev.record(x["stream"])
            stream.wait_event(ev)
        e1.record(stream)
        barrier()
        t1 = time.time()
        if args.profile:
            torch.cuda.profiler.stop()
        ms = e0.elapsed_time(e1)
        for x in inst:
            x["lock"].sync()
        st1 = [x["lock"].stats() for x in inst]
        stats = {k: sum(b_[k] - a_[k] for a_, b_ in zip(st0, st1)) for k in st1[0]}
        consumed = stats["iterations"]
        if world_size > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            c = torch.tensor([consumed], dtype=torch.float64, device="cuda")
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
            consumed_all = float(c.item())
        else:
            consumed_all = float(consumed)
        cnts = [x["world"].counters() for x in inst]
        cnt = {k: sum(c[k] for c in cnts) for k in cnts[0]}
        launches = g.capi.launch_count() - l0
        nodes1 = float(np.mean([x["lock"].info(p)["tree_size"] for x in inst for p in range(0, x["hi"] - x["lo"], 97)]))

        def info(p):
            for x in inst:
                if x["lo"] <= p < x["hi"]:
                    return x["lock"].info(p - x["lo"])

        def dump(p):
            for x in inst:
                if x["lo"] <= p < x["hi"]:
                    return x["lock"].dump(p - x["lo"])

        return {"ms": ms, "consumed": consumed_all, "stats": stats, "world": cnt, "launches": launches, "t0": t0,
                "t1": t1, "info": info, "dump": dump, "instances": G, "keep": inst, "nodes": (nodes0, nodes1)}

    if args.profile:
        val_run = timed_run(device_samples=True, instances=args.instances, kernel_timing=args.kernel_timing)
        if rank == 0:
            vs = val_run["stats"]
            wc = val_run["world"]
            print(json.dumps({"profile_run": True, "ms_per_step": val_run["ms"] / K,
                              "value": val_run["consumed"] / (val_run["ms"] * 1e-3),
                              "edge_kernel_us_per_iteration": round(wc["kernel_ns"] / (K * I) / 1e3, 1),
                              "edge_tflops": round(edge_flops(wc) / max(wc["kernel_ns"], 1) / 1e3, 2),
                              "edges_per_iteration": vs["edges_device"] / (K * I),
                              "extended_per_iteration": vs["extended"] / (K * I), "tree_nodes": val_run["nodes"]}))
        return 0
    val_run = timed_run(device_samples=True, instances=args.instances)
    e2e_run = timed_run(device_samples=False, instances=args.instances)
    # the dominant kernel is timed in a run of its own with ONE instance: with several instances their kernels share
    # the SMs, so a per-launch duration taken there would not be that of the kernel alone
    roof_run = timed_run(device_samples=True, instances=1, kernel_timing=True)
    value = val_run["consumed"] / (val_run["ms"] * 1e-3)
    e2e_value = e2e_run["consumed"] / (e2e_run["ms"] * 1e-3)
    clk = clocks.window(val_run["t0"], val_run["t1"])
    clocks.stop()

    # all runs consumed the same samples: they must have built the same trees
    same = all(e2e_run["info"](p) == val_run["info"](p) == roof_run["info"](p) for p in range(0, P, max(1, P // 64)))

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": K, "warmup": W_,
        "ms_per_step": val_run["ms"] / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 planner / f32 collision geometry", "data": "synthetic",
        "config": {
            "workload": "C3: 7-DoF sphere-arm RRT* vs 1000-sphere cloud, %d lock-step problems per GPU; one step = %d "
                        "RRT* iterations of every problem (%d samples)" % (P, I, P * I),
            "problems_per_gpu": P, "iterations_per_step": I, "samples_per_step": P * I,
            "tree_nodes": {"at_start_of_timed_region": val_run["nodes"][0], "at_end": val_run["nodes"][1],
                           "grown_by": "%d warm-up steps (%d iterations) outside the timed region" % (W_, W_ * I)},
            "max_distance": sc.max_distance, "min_distance": sc.min_distance,
            "sampler": "uniform over the joint box, Philox blocks injected through the reference's sample-injection entry "
                       "(tree_solver.h:338) so that the CPU arm and the oracle consume the same samples; the rewire radius "
                       "follows the informed set (InformedSampler::getSpecificVolume).  The device informed sampler arm is "
                       "reported under `informed`",
            "trees": "device-resident (gc_forest)" if device_trees else "host replay (round 1)",
            "host_threads": args.threads or ncores_rank, "instances_per_gpu": val_run["instances"],
            "l2_policy": "every iteration has new samples and larger trees; no artificial L2 flush between planner "
                         "iterations (the kNN measurement flushes L2 between launches)",
            "parallelism": "independent problems per GPU, no data-path collective",
        },
        "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"]},
        "gpu_launches": int(val_run["launches"]),
    }
    st = e2e_run["stats"]
    line["e2e"] = {"value": e2e_value, "unit": UNIT,
                   "h2d_bytes_per_step": int(st.get("h2d_bytes", 0) / K), "d2h_bytes_per_step": int(st.get("d2h_bytes", 0) / K + 8 * P),
                   "ms_per_step": e2e_run["ms"] / K, "same_trees_as_value_run": bool(same),
                   "what": "every iteration's sample block goes H2D from pinned host memory; every step ends with a D2H read "
                           "of the %d path costs and a stream synchronisation" % P}
    wc = roof_run["world"]
    brute_pairs_per_state = float(len(sc.params["rs_link"]) * len(sc.params["obs"]))
    achieved = edge_flops(wc) / max(wc["kernel_ns"], 1) / 1e3  # TFLOP/s
    executed = EXECUTED_FLOP_PER_PAIR * wc["pair_tests"] / max(wc["kernel_ns"], 1) / 1e3
    traffic = {}
    for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        try:
            traffic = json.loads((ROOT / "profiles" / name).read_text())
            break
        except Exception:
            pass
    line["roofline"] = {
        "kernel": "arm_state_kernel (edge checks)", "bound": "fp32", "achieved": achieved, "peak": fp32_peak,
        "unit": "TFLOP/s", "frac": achieved / fp32_peak if fp32_peak > 0 else None,
        "traffic": traffic.get("arm_state_kernel", {}).get("dram_bytes_per_launch"),
        "traffic_source": traffic.get("arm_state_kernel", {}).get("capture"),
        "peak_source": "FFMA micro-benchmark in this run (implies %.0f MHz); MEASURED_PEAKS.json has no FP32 entry" % fp32_mhz,
        "accounting": "SURVEY 8(d): %g flop per executed sphere-pair test + %g flop FK per evaluated state (counted by "
                      "the kernel, so early exits do not inflate it)" % (FLOP_PER_PAIR, FLOP_FK_PER_STATE),
        "executed_pair_flops": {"flop_per_pair": EXECUTED_FLOP_PER_PAIR, "achieved": executed,
                                "frac": executed / fp32_peak if fp32_peak > 0 else None},
        "brute_force_equivalent": {
            "note": "what the same states would cost without block culling (SURVEY 8(d): every robot sphere against "
                    "every obstacle); not a roofline figure, it shows what the culling saves",
            "tflops": (FLOP_PER_PAIR * brute_pairs_per_state + FLOP_FK_PER_STATE) * wc["states"] / max(wc["kernel_ns"], 1) / 1e3,
            "pair_tests_executed_share": wc["pair_tests"] / max(brute_pairs_per_state * wc["states"], 1)},
        "pair_tests_per_launch": wc["pair_tests"] / max(wc["launches"], 1),
        "states_per_launch": wc["states"] / max(wc["launches"], 1),
        "us_per_launch": wc["kernel_ns"] / max(wc["launches"], 1) / 1e3,
        "kernel_share_of_step": wc["kernel_ns"] / 1e6 / roof_run["ms"],
        "measured_in": "timed run of the same %d steps with 1 planner instance, launched kernel by kernel with CUDA events "
                       "around the check kernels (%.3f ms/step); the headline runs replay the step from a CUDA graph with "
                       "%d instance(s)" % (K, roof_run["ms"] / K, val_run["instances"]),
    }
    vs = val_run["stats"]
    nit = K * I
    line["step_breakdown"] = {"extended_per_iteration": vs["extended"] / nit,
                              "device_edges_per_iteration": vs["edges_device"] / nit,
                              "edges_used_per_iteration": vs["edges_used"] / nit, "edge_fallbacks": vs["edges_fallback"],
                              "near_hits_per_extension": vs["near_hits"] / max(vs["extended"], 1),
                              "abi_calls_per_iteration": vs["gpu_calls"] / nit,
                              "launches_per_iteration": val_run["launches"] / nit / val_run["instances"]}

    # ---- kNN half of BASELINE's metric: C4 node-sharded over the ranks (every rank takes part) ------------------
    if not args.skip_nn:
        try:
            peaks = {}
            try:
                peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
            except Exception:
                pass
            line["knn"] = knn_c4(g, torch, dist if world_size > 1 else None, rank, world_size, dev,
                                 float(peaks.get("hbm_gbs", 6650.0)), fp32_peak)
        except Exception as ex:
            line["knn"] = {"error": str(ex)[:300]}
    if not args.skip_extras:
        try:
            line["c5"] = c5_leg(g, torch, dist if world_size > 1 else None, rank, world_size, dev, sc, root_world, ncores)
        except Exception as ex:
            line["c5"] = {"error": str(ex)[:300]}
    if rank == 0 and not args.skip_extras:
        try:
            line["informed"] = informed_arm(g, torch, LockstepRRTStar, sc, root_world, starts, goals, dev, W_ * I, K * I)
        except Exception as ex:
            line["informed"] = {"error": str(ex)[:300]}
        try:
            line["depth_sweep"] = depth_sweep(g, torch, LockstepRRTStar, sc, root_world, starts, goals, dev)
        except Exception as ex:
            line["depth_sweep"] = {"error": str(ex)[:300]}
        try:
            line["c2"] = c2_legs(g, torch, dev, ncores)
        except Exception as ex:
            line["c2"] = {"error": str(ex)[:300]}
    if rank == 0:
        # ---- in-bench parity check: 16 problems, full depth, against the oracle on the same samples -----------
        sys.path.insert(0, str(ROOT / "oracle"))
        import oracle_py as orc
        orc.build()
        ow = orc.OracleWorld.from_scenario(sc)
        npar = min(16, P)
        ok_all, nodes_tot = True, 0
        for p in [P * j // npar for j in range(npar)]:
            o = orc.OracleSolver(orc.OracleSolver.RRTSTAR, sc, ow, start=starts[p], goal=goals[p])
            o.run(samples[:, p, :])
            par, pc, conf = o.dump()
            gpar, gpc, gconf = val_run["dump"](p)
            ok = bool(len(par) == len(gpar) and np.array_equal(par, gpar) and np.array_equal(pc, gpc) and
                      np.array_equal(conf, gconf))
            ok_all = ok_all and ok
            nodes_tot += int(len(par))
        line["parity"] = {"trees_bit_exact": ok_all, "problems_checked": npar, "iterations": iters,
                          "nodes_per_tree": nodes_tot / npar}
        # ---- CPU baseline on a bounded sample of the same workload ---------------------------------
        ncpu = args.cpu_problems or min(P, ncores * 2)
        rate, dt, nthreads, _, kind = cpu_reference_run(sc, starts[:ncpu], goals[:ncpu], samples[:, :ncpu, :], W_ * I,
                                                        K * I, ncores)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": nthreads, "kind": kind,
                                "sample": "%d of the %d problems x %d RRT* iterations (same samples, same %d warm-up "
                                          "iterations), %s, one problem per thread, %.1f s" %
                                          (ncpu, P, K * I, W_ * I, CPU_DESC[kind], dt)}
        # the same reference planner with the state check given the CUDA path's grid culling (bit-identical verdicts):
        # the fair denominator -- the plain baseline above credits the GPU for an algorithm the CPU was never given
        ncpu2 = args.cpu_problems or min(P, ncores * 8)
        rate2, dt2, nthreads2, _, kind2 = cpu_reference_run(sc, starts[:ncpu2], goals[:ncpu2], samples[:, :ncpu2, :],
                                                            W_ * I, K * I, ncores, grid=True)
        line["cpu_baseline_optimized"] = {
            "value": rate2, "unit": UNIT, "cores": nthreads2, "kind": kind2,
            "sample": "%d of the %d problems x %d RRT* iterations, same planner sources, sphere-arm check with the "
                      "uniform-grid culling of the CUDA path (oracle restatement, verdicts identical to the plain loop; "
                      "after culling a state costs one FK chain + ~20 pair tests, so an 8-wide pair loop has nothing left "
                      "to vectorise), %.1f s" % (ncpu2, P, K * I, dt2),
            "gpu_over_cpu": {"value_ratio": value / rate2 if rate2 > 0 else None,
                             "e2e_ratio": e2e_value / rate2 if rate2 > 0 else None}}
        # ---- kNN side measurement: BASELINE config 4 shape, HBM-bound regime --------------------------
        if not args.skip_nn:
            try:
                line["roofline_nn"] = nn_roofline(g, torch, dev)
            except Exception as ex:  # the headline line must still be printed
                line["roofline_nn"] = {"error": str(ex)[:200]}
        print(json.dumps(line))
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def nn_roofline(g, torch, dev, dim=12):
    """Nearest-neighbour scan (scan_stream_kernel) of a 12-DoF node store, L2 flushed between launches, CUDA events on
    the launching stream.  HBM roofline: one query group per pass over the node set, so the algorithmic bytes are
    4*D*N (fp32 SoA read once) + queries + results.  Reported for a store large enough that launch latency is
    amortised (2^23 nodes, 403 MB) and for BASELINE config 4's own size (2^20 nodes, 50 MB, where a 7.7 us transfer
    sits inside a ~30 us launch); the batched regime (4096 queries) is FP32-bound and reported in pairs/s."""
    from graph_core_b200 import worlds as W
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    lib = g.capi.load()
    stream = torch.cuda.Stream()
    out = {}

    def run_case(st, n, nq, reps=5):
        q = torch.from_numpy(W.c4_queries(nq, dim)).cuda()
        idx = torch.empty(nq, dtype=torch.int32, device="cuda")
        dd = torch.empty(nq, dtype=torch.float64, device="cuda")
        times = []
        for it in range(reps + 3):
            g.capi.flush_l2(dev)
            with torch.cuda.stream(stream):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                rc = lib.gc_nn1_dev(st.handle, q.data_ptr(), None, nq, idx.data_ptr(), dd.data_ptr())
                e1.record(stream)
            torch.cuda.synchronize()
            assert rc == 0
            if it >= 3:
                times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        passes = (nq + 63) // 64  # queries per pass over the node set (scan_stream_kernel: 64 per group)
        alg_bytes = passes * 4.0 * dim * n + 8.0 * dim * nq + 12.0 * nq
        return {"ms": ms, "queries_per_s": nq / (ms * 1e-3), "GBps": alg_bytes / (ms * 1e-3) / 1e9,
                "pairs_per_s": nq * n / (ms * 1e-3), "alg_bytes": alg_bytes}

    n4 = 1 << 20
    st = g.NodeStore(dim, n4, device=dev)
    st.append(W.c4_nodes(n4, dim))
    st.set_stream(stream.cuda_stream)
    for nq in (1, 8, 4096):
        out["c4_n2^20_nq%d" % nq] = run_case(st, n4, nq)
    st.close()
    nl = 1 << 23
    rng = np.random.default_rng(4010)
    big = g.NodeStore(dim, nl, device=dev)
    for part in range(8):  # appended in 8 slices to bound the host staging buffers
        big.append(rng.uniform(-np.pi, np.pi, size=(nl // 8, dim)))
    big.set_stream(stream.cuda_stream)
    a = run_case(big, nl, 1)
    out["large_n2^23_nq1"] = a
    out["large_n2^23_nq8"] = run_case(big, nl, 8)
    big.close()
    return {"kernel": "scan_stream_kernel<12,nearest>", "bound": "hbm", "achieved": a["GBps"], "peak": hbm,
            "unit": "GB/s", "frac": a["GBps"] / hbm, "traffic": None, "peak_source": src,
            "alg_bytes_per_launch": a["alg_bytes"], "us_per_launch": a["ms"] * 1e3,
            "workload": "nearest neighbour of 1 query over 2^23 nodes x 12-DoF (403 MB fp32 SoA), L2 flushed between "
                        "launches; BASELINE config 4's 2^20-node store is in detail",
            "detail": out}


if __name__ == "__main__":
    sys.exit(main())
