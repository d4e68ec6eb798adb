#!/usr/bin/env python
"""Debug driver of the two-launch k-nearest path: runs the cases of tests/test_sharded_gpu.py one by one."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import graph_core_b200 as g  # noqa: E402
from graph_core_b200 import worlds as W  # noqa: E402

case = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n, dim, k, nq = [(50000, 12, 16, 1), (70001, 7, 64, 8), (40000, 3, 5, 3), (1 << 20, 12, 16, 2), (33000, 2, 1, 8)][case]
pts = W.c4_nodes(n, dim, seed=391)
qs = W.c4_queries(nq, dim, seed=392)
st = g.NodeStore(dim, n)
st.append(pts)
l0 = g.capi.launch_count()
gi, gd, gcnt = st.knn(qs, k, cap=k)
print("case", case, "launches", g.capi.launch_count() - l0, gcnt, gd[0, :4])
