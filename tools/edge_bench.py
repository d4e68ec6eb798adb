#!/usr/bin/env python
"""Edge-batch micro-benchmark of SURVEY.md 8(d), config C2: 2^20 random 6-DoF segments with length U[0.05, 1.0]
(Philox seed 203) against 64 C-space hyperspheres, bisection order, min_distance 0.01, through gc_check_edges_dev
(device pointers, CUDA events on the launching stream).  Two legs: the random batch as drawn (colliding edges exit
early) and the collision-free subset of it (every state evaluated: the roofline figure).  Algorithmic work per
evaluated state: 64 spheres x (3*6 + 1) = 1 216 flop; the kernel counts the states it evaluates, so early
exits do not inflate the figure.  The reference's base-class checkConnection over the oracle world is timed beside it
on a bounded sample (one host core, the reference is single-threaded).  Prints one JSON line."""
import argparse
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
import graph_core_b200 as g  # noqa: E402
from graph_core_b200 import worlds as W  # noqa: E402

FLOP_PER_STATE = 64 * (3 * 6 + 1)
FLOP_PER_PAIR = 3 * 6 + 1  # one state against one sphere: 6 sub, 6 mul, 5 add, 1 compare


def segments(n, dim, lb, ub, seed=203):
    u = W.philox_u01(seed, 0, n, 2 * dim + 1)
    c1 = lb + u[:, :dim] * (ub - lb)
    d = u[:, dim:2 * dim] - 0.5
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    length = 0.05 + 0.95 * u[:, 2 * dim]
    c2 = np.clip(c1 + d * length[:, None], lb, ub)
    return np.ascontiguousarray(c1), np.ascontiguousarray(c2)


def timed(lib, w, stream, c1, c2, n, min_d, ok, reps):
    times = []
    for it in range(reps + 2):
        g.capi.flush_l2(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            rc = lib.gc_check_edges_dev(w.handle, c1.data_ptr(), c2.data_ptr(), n, min_d, 0, ok.data_ptr(), None)
            e1.record(stream)
        torch.cuda.synchronize()
        assert rc >= 0, lib.gc_last_error()
        if it >= 2:
            times.append(e0.elapsed_time(e1))
    return float(np.median(times))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--segments", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu-edges", type=int, default=20000)
    ap.add_argument("--loop", type=int, default=0, help="profiling: launch the free-edge batch this many times and exit")
    args = ap.parse_args()
    sc = W.scenario_c2()
    lib = g.capi.load()
    w = g.CollisionWorld.cspheres(sc.params["centers"], sc.params["radii"])
    stream = torch.cuda.Stream()
    w.set_stream(stream.cuda_stream)
    n = args.segments
    h1, h2 = segments(n, sc.dim, sc.lb, sc.ub)
    c1, c2 = torch.from_numpy(h1).cuda(), torch.from_numpy(h2).cuda()
    ok = torch.empty(n, dtype=torch.uint8, device="cuda")
    fp32_peak, mhz = g.capi.measure_fp32_peak(0)
    min_d = 0.01
    out = {"workload": "C2 edge micro-benchmark: %d segments, 6-DoF, 64 C-space spheres, bisection, min_distance 0.01" % n,
           "kernel": "edge_warp_kernel<6,BISECT>", "flop_per_state": FLOP_PER_STATE, "fp32_peak_tflops": fp32_peak}
    legs = {}
    w.reset_counters()
    ms = timed(lib, w, stream, c1, c2, n, min_d, ok, args.reps)
    free = ok.cpu().numpy().astype(bool)
    cnt = w.counters()
    states = cnt["states"] / (args.reps + 2)
    pairs = cnt["pair_tests"] / (args.reps + 2)
    legs["as_drawn"] = {"edges": n, "free_fraction": float(free.mean()), "ms": ms, "edges_per_s": n / ms * 1e3,
                        "states_evaluated": states, "pair_tests": pairs, "tflops": pairs * FLOP_PER_PAIR / ms / 1e9,
                        "frac_of_fp32_peak": pairs * FLOP_PER_PAIR / ms / 1e9 / fp32_peak}
    f1, f2 = c1[torch.from_numpy(free).cuda()].contiguous(), c2[torch.from_numpy(free).cuda()].contiguous()
    nf = f1.shape[0]
    okf = torch.empty(nf, dtype=torch.uint8, device="cuda")
    if args.loop:
        with torch.cuda.stream(stream):
            for _ in range(args.loop):
                lib.gc_check_edges_dev(w.handle, f1.data_ptr(), f2.data_ptr(), nf, min_d, 0, okf.data_ptr(), None)
        torch.cuda.synchronize()
        print(json.dumps({"loop": args.loop, "free_edges": nf}))
        return
    w.reset_counters()
    ms = timed(lib, w, stream, f1, f2, nf, min_d, okf, args.reps)
    assert bool(okf.all().item())
    cnt = w.counters()
    states = cnt["states"] / (args.reps + 2)
    pairs = cnt["pair_tests"] / (args.reps + 2)
    legs["collision_free"] = {"edges": nf, "ms": ms, "edges_per_s": nf / ms * 1e3, "states_evaluated": states,
                              "states_per_edge": states / nf, "pair_tests": pairs,
                              "tflops": pairs * FLOP_PER_PAIR / ms / 1e9,
                              "frac_of_fp32_peak": pairs * FLOP_PER_PAIR / ms / 1e9 / fp32_peak,
                              "alg_bytes": nf * (2 * 6 * 8 + 1), "launches_per_call": cnt["launches"] / (args.reps + 2)}
    out["legs"] = legs
    out["roofline"] = {"kernel": "edge_warp_kernel<6,BISECT>", "bound": "fp32", "achieved": legs["collision_free"]["tflops"],
                       "peak": fp32_peak, "unit": "TFLOP/s", "frac": legs["collision_free"]["frac_of_fp32_peak"],
                       "accounting": "19 flop per executed (state, sphere) test = 1 216 per collision-free state (SURVEY 8d), "
                                     "counted by the kernel; time = the "
                                     "whole gc_check_edges_dev call (plan + check kernels)"}
    # ---- the reference's checkConnection on one host core (bounded sample) -------------------------------------
    try:
        import oracle_py as orc
        orc.build()
        ow = orc.OracleWorld.from_scenario(sc, fast=True)
        m = min(args.cpu_edges, n)
        t0 = time.perf_counter()
        cok = ow.check_connection(h1[:m], h2[:m], min_d, 0)
        dt = time.perf_counter() - t0
        assert np.array_equal(np.asarray(cok).astype(bool), free[:m]), "verdicts differ from the oracle"
        out["cpu_baseline"] = {"value": m / dt, "unit": "edges/s", "cores": 1, "kind": "port",
                               "sample": "first %d segments of the batch, oracle restatement of checkConnection "
                                         "(collision_checker_base.h:207-237) with the reference's Release flags, %.2f s; "
                                         "verdicts equal to the CUDA path's" % (m, dt)}
        out["gpu_over_cpu_core"] = legs["as_drawn"]["edges_per_s"] / (m / dt)
    except Exception as ex:  # the oracle is test infrastructure: its absence must not hide the GPU numbers
        out["cpu_baseline"] = {"error": str(ex)[:200]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
