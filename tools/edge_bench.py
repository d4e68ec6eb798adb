#!/usr/bin/env python
"""Edge-batch micro-benchmark of SURVEY.md 8(d), config C2: 2^20 random 6-DoF segments with length U[0.05, 1.0]
(Philox seed 203) against 64 C-space hyperspheres, bisection order, min_distance 0.01, through gc_check_edges_dev
(device pointers, CUDA events on the launching stream).  Two legs: the random batch as drawn (colliding edges exit
early) and the collision-free subset of it (every state evaluated: the roofline figure).  Algorithmic work per
evaluated state: 64 spheres x (3*6 + 1) = 1 216 flop; the kernel counts the states it evaluates, so early
exits do not inflate the figure.  `bench.py` runs the same legs under its `edge_c2` key and times the reference's
checkConnection loop beside them (its cpu_baseline leg).  Prints one JSON line."""
import argparse
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
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


def gpu_legs(n_segments=1 << 20, reps=5, loop=0):
    """The two GPU legs; returns (result dict, host end points c1, c2, per-edge verdicts of the batch as drawn)."""
    sc = W.scenario_c2()
    lib = g.capi.load()
    w = g.CollisionWorld.cspheres(sc.params["centers"], sc.params["radii"])
    stream = torch.cuda.Stream()
    w.set_stream(stream.cuda_stream)
    n = n_segments
    h1, h2 = segments(n, sc.dim, sc.lb, sc.ub)
    c1, c2 = torch.from_numpy(h1).cuda(), torch.from_numpy(h2).cuda()
    ok = torch.empty(n, dtype=torch.uint8, device="cuda")
    fp32_peak, mhz = g.capi.measure_fp32_peak(0)
    min_d = 0.01
    out = {"workload": "C2 edge micro-benchmark: %d segments, 6-DoF, 64 C-space spheres, bisection, min_distance 0.01" % n,
           "kernel": "edge_warp_kernel<6,BISECT>", "flop_per_state": FLOP_PER_STATE, "fp32_peak_tflops": fp32_peak}
    legs = {}
    w.reset_counters()
    ms = timed(lib, w, stream, c1, c2, n, min_d, ok, reps)
    free = ok.cpu().numpy().astype(bool)
    cnt = w.counters()
    states = cnt["states"] / (reps + 2)
    pairs = cnt["pair_tests"] / (reps + 2)
    legs["as_drawn"] = {"edges": n, "free_fraction": float(free.mean()), "ms": ms, "edges_per_s": n / ms * 1e3,
                        "states_evaluated": states, "pair_tests": pairs, "tflops": pairs * FLOP_PER_PAIR / ms / 1e9,
                        "frac_of_fp32_peak": pairs * FLOP_PER_PAIR / ms / 1e9 / fp32_peak}
    f1, f2 = c1[torch.from_numpy(free).cuda()].contiguous(), c2[torch.from_numpy(free).cuda()].contiguous()
    nf = f1.shape[0]
    okf = torch.empty(nf, dtype=torch.uint8, device="cuda")
    if loop:
        with torch.cuda.stream(stream):
            for _ in range(loop):
                lib.gc_check_edges_dev(w.handle, f1.data_ptr(), f2.data_ptr(), nf, min_d, 0, okf.data_ptr(), None)
        torch.cuda.synchronize()
        return {"loop": loop, "free_edges": nf}, h1, h2, free
    w.reset_counters()
    ms = timed(lib, w, stream, f1, f2, nf, min_d, okf, reps)
    assert bool(okf.all().item())
    cnt = w.counters()
    states = cnt["states"] / (reps + 2)
    pairs = cnt["pair_tests"] / (reps + 2)
    legs["collision_free"] = {"edges": nf, "ms": ms, "edges_per_s": nf / ms * 1e3, "states_evaluated": states,
                              "states_per_edge": states / nf, "pair_tests": pairs,
                              "tflops": pairs * FLOP_PER_PAIR / ms / 1e9,
                              "frac_of_fp32_peak": pairs * FLOP_PER_PAIR / ms / 1e9 / fp32_peak,
                              "alg_bytes": nf * (2 * 6 * 8 + 1), "launches_per_call": cnt["launches"] / (reps + 2)}
    out["legs"] = legs
    out["roofline"] = {"kernel": "edge_warp_kernel<6,BISECT>", "bound": "fp32", "achieved": legs["collision_free"]["tflops"],
                       "peak": fp32_peak, "unit": "TFLOP/s", "frac": legs["collision_free"]["frac_of_fp32_peak"],
                       "accounting": "19 flop per executed (state, sphere) test = 1 216 per collision-free state (SURVEY 8d), "
                                     "counted by the kernel; time = the whole gc_check_edges_dev call"}
    return out, h1, h2, free


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--segments", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--loop", type=int, default=0, help="profiling: launch the free-edge batch this many times and exit")
    args = ap.parse_args()
    out, _, _, _ = gpu_legs(args.segments, args.reps, args.loop)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
