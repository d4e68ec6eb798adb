#!/usr/bin/env python
"""The C-ABI node-sharded tree (gc_comm_init / gc_store_create_sharded) with ONE SHARD PER GPU of the box, driven by
this single process: answers compared with the unsharded store on device 0 (bit-exact) and timed (wall clock around
the synchronous C-ABI call, host queries in, host results out).  Prints one JSON line per case."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import graph_core_b200 as g  # noqa: E402
from graph_core_b200 import worlds as W  # noqa: E402


def main():
    ndev = torch.cuda.device_count()
    n, dim = 1 << 20, 12
    pts = W.c4_nodes(n, dim)
    qs = W.c4_queries(4096, dim)
    one = g.NodeStore(dim, n)
    one.append(pts)
    st = g.ShardedStore(dim, n, list(range(ndev)))
    st.append(pts)
    for nq in (1, 16, 4096):
        for k in (1, 16, 64):
            q = qs[:nq]
            if k == 1:
                ri, rd = one.nearest(q)
                gi, gd = st.nearest(q)
                same = bool(np.array_equal(ri, gi) and np.array_equal(rd, gd))
                f_one, f_sh = (lambda: one.nearest(q)), (lambda: st.nearest(q))
            else:
                ri, rd, rc = one.knn(q, k, cap=k)
                gi, gd, gc_ = st.knn(q, k)
                same = bool(np.array_equal(ri, gi) and np.array_equal(rd, gd) and np.array_equal(rc, gc_))
                f_one, f_sh = (lambda: one.knn(q, k, cap=k)), (lambda: st.knn(q, k))
            t = {}
            for name, f in (("one_gpu", f_one), ("sharded", f_sh)):
                f()
                ts = []
                for _ in range(7):
                    t0 = time.perf_counter()
                    f()
                    ts.append(time.perf_counter() - t0)
                t[name] = float(np.median(ts))
            print(json.dumps({"n_gpus": ndev, "queries": nq, "k": k, "same_as_one_store": same,
                              "one_gpu_ms": t["one_gpu"] * 1e3, "sharded_ms": t["sharded"] * 1e3,
                              "sharded_queries_per_s": nq / t["sharded"], "speedup": t["one_gpu"] / t["sharded"]}), flush=True)


if __name__ == "__main__":
    main()
