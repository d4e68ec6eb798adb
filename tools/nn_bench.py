#!/usr/bin/env python
"""Kernel-level timing of the large-tree NN path (BASELINE config 4 shape: 2^20 nodes x 12 DoF) through the
C ABI's device-pointer entry points: nearest / radius / k-nearest at several query-tile sizes, L2 flushed
between launches, CUDA events on the launching stream.  Prints one JSON line per case."""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import graph_core_b200 as g  # noqa: E402
from graph_core_b200 import worlds as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=1 << 20)
    ap.add_argument("--dim", type=int, default=12)
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--ops", default="nearest,near,knn")
    ap.add_argument("--tiles", default="1,8,64,512,4096")
    ap.add_argument("--no-flush", action="store_true")
    args = ap.parse_args()
    n, dim = args.nodes, args.dim
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    lib = g.capi.load()
    st = g.NodeStore(dim, n)
    st.append(W.c4_nodes(n, dim))
    stream = torch.cuda.Stream()
    st.set_stream(stream.cuda_stream)
    fp32_peak, _ = g.capi.measure_fp32_peak(0)
    for op in args.ops.split(","):
        for nq in map(int, args.tiles.split(",")):
            q = torch.from_numpy(W.c4_queries(nq, dim)).cuda()
            cap = 1 if op == "nearest" else (256 if op == "near" else 16)
            idx = torch.empty(nq * cap, dtype=torch.int32, device="cuda")
            dd = torch.empty(nq * cap, dtype=torch.float64, device="cuda")
            cnt = torch.empty(nq, dtype=torch.int32, device="cuda")
            rad = torch.full((nq,), W.c4_radius(16, n, dim), dtype=torch.float64, device="cuda")
            times = []
            for it in range(args.reps + 2):
                if not args.no_flush:
                    g.capi.flush_l2(0)
                with torch.cuda.stream(stream):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    if op == "nearest":
                        rc = lib.gc_nn1_dev(st.handle, q.data_ptr(), None, nq, idx.data_ptr(), dd.data_ptr())
                    elif op == "near":
                        rc = lib.gc_radius_dev(st.handle, q.data_ptr(), None, rad.data_ptr(), nq, cap, idx.data_ptr(),
                                               dd.data_ptr(), cnt.data_ptr())
                    else:
                        rc = lib.gc_knn_dev(st.handle, q.data_ptr(), None, nq, 16, cap, idx.data_ptr(), dd.data_ptr(),
                                            cnt.data_ptr())
                    e1.record(stream)
                torch.cuda.synchronize()
                assert rc >= 0, lib.gc_last_error()
                if it >= 2:
                    times.append(e0.elapsed_time(e1))
            ms = float(np.median(times))
            groups = (nq + 63) // 64
            alg_bytes = groups * 4.0 * dim * n + 8.0 * dim * nq + 12.0 * nq * cap
            pairs = float(nq) * n
            print(json.dumps({"op": op, "nq": nq, "ms": round(ms, 4), "queries_per_s": round(nq / ms * 1e3),
                              "GBps_alg": round(alg_bytes / ms / 1e6, 1), "hbm_frac": round(alg_bytes / ms / 1e6 / hbm, 3),
                              "Tpairs_per_s": round(pairs / ms / 1e9, 3),
                              "fp32_frac_3D": round(3 * dim * pairs / ms / 1e9 / fp32_peak, 3)}), flush=True)


if __name__ == "__main__":
    main()
