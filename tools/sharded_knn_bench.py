#!/usr/bin/env python
"""BASELINE config 4: kNN / radius-near sweep over a 2^20-node 12-DoF tree, node-sharded over the GPUs of
one box with a single NCCL all-gather + device merge per query batch.

  python tools/sharded_knn_bench.py                         # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
         tools/sharded_knn_bench.py [--nodes N] [--queries Q] [--check]

Prints one JSON line per (operation, k): queries/s = Q / max-over-ranks device time (CUDA events on the
launching stream, barrier on both sides).  --check compares rank 0's answers with the CPU oracle on the
first 32 queries (bit-exact indices and distances).
"""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=1 << 20)
    ap.add_argument("--dim", type=int, default=12)
    ap.add_argument("--queries", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "p2p"],
                    help="nccl: all-gather + merge; p2p: the merge kernel loads the shard lists over NVLink itself")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from graph_core_b200 import worlds as W
    from graph_core_b200.sharded import ShardedNodeStore

    pts = W.c4_nodes(args.nodes, args.dim)
    qs = W.c4_queries(args.queries, args.dim)
    st = ShardedNodeStore(args.dim, args.nodes, device=local, exchange=args.exchange, nq_max=args.queries, cap_max=1024)
    st.append(pts)
    torch.cuda.synchronize()

    def timed(fn):
        times = []
        out = None
        for it in range(args.reps + 2):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            if it >= 2:
                times.append(ms)
        return float(np.median(times)), out

    results = []
    ms, (ni, nd) = timed(lambda: st.nearest(qs))
    results.append({"op": "nearest", "k": 1, "ms": ms, "queries_per_s": args.queries / (ms * 1e-3)})
    # latency regime: a handful of queries, where the exchange step is a visible share of the call
    for nq_small in (1, 16):
        ms, _ = timed(lambda: st.knn(qs[:nq_small], 16))
        results.append({"op": "knn_small_batch", "k": 16, "batch": nq_small, "ms": ms,
                        "queries_per_s": nq_small / (ms * 1e-3)})
    knn_out = {}
    for k in (16, 64, 256, 1024):
        ms, out = timed(lambda: st.knn(qs, k))
        knn_out[k] = out
        results.append({"op": "knn", "k": k, "ms": ms, "queries_per_s": args.queries / (ms * 1e-3),
                        "allgather_bytes_per_rank": args.queries * (k * 12 + 4)})
    for hits in (16, 128):
        r = W.c4_radius(hits, args.nodes, args.dim)
        ms, out = timed(lambda: st.near(qs, r, cap=1024 // max(world, 1) if hits > 64 else 256))
        results.append({"op": "near", "expected_hits": hits, "radius": r, "ms": ms,
                        "queries_per_s": args.queries / (ms * 1e-3), "mean_hits": float(out[2].float().mean().item())})
    if rank == 0:
        check = None
        if args.check:
            sys.path.insert(0, str(ROOT / "oracle"))
            import oracle_py as orc
            orc.build()
            ov = orc.OracleNN("vector", args.dim)
            ov.insert(pts)
            m = 32
            ri, rd = ov.nearest(qs[:m])
            ok = np.array_equal(ni[:m].cpu().numpy(), ri) and np.array_equal(nd[:m].cpu().numpy(), rd)
            for k, (oi, od, oc) in knn_out.items():
                ki, kd, kc = ov.knn(qs[:m], k, k)
                ok = ok and np.array_equal(oi[:m].cpu().numpy(), ki) and np.array_equal(od[:m].cpu().numpy(), kd)
            check = bool(ok)
        for r in results:
            r.update({"exchange": args.exchange if world > 1 else "none", "n_gpus": world, "nodes": args.nodes, "dim": args.dim, "queries": args.queries,
                      "oracle_check": check})
            print(json.dumps(r))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
