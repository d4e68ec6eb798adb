"""Developer probe: coarse wall-clock timings of the C-ABI calls on the GPU box (not a benchmark)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
import numpy as np
import graph_core_b200 as g
from graph_core_b200 import worlds as W
import oracle_py as orc

def tm(f, n=5):
    f(); ts = []
    for _ in range(n):
        t = time.perf_counter(); f(); ts.append(time.perf_counter() - t)
    return min(ts)

print(g.device_info(0))
N, D = 1 << 20, 12
pts = W.c4_nodes(N, D); qs = W.c4_queries(4096, D)
st = g.NodeStore(D, N); t = time.perf_counter(); st.append(pts); print("append 1M x12: %.1f ms" % ((time.perf_counter() - t) * 1e3))
for nq in (1, 8, 16, 64, 256, 4096):
    t = tm(lambda: st.nearest(qs[:nq]))
    print("nn1  nq=%5d  %.3f ms  -> %.0f q/s, %.1f GB/s algorithmic (per 8-query group pass)" % (nq, t * 1e3, nq / t, (4.0 * D * N * ((nq + 7) // 8)) / t / 1e9))
r = W.c4_radius(16, N, D)
for nq in (1, 16, 256, 4096):
    t = tm(lambda: st.near(qs[:nq], r, cap=256))
    print("near nq=%5d  %.3f ms  -> %.0f q/s" % (nq, t * 1e3, nq / t))
sc = W.scenario_c3(1000, orc.arm_checker())
gw = sc.make_world(); gw.enable_timing(True)
a = W.uniform_samples(sc.lb, sc.ub, 3, 0, 4096)
u = W.philox_u01(4, 0, 4096, 8); dirn = u[:, :7] - 0.5; dirn /= np.linalg.norm(dirn, axis=1, keepdims=True)
b = np.clip(a + dirn * (0.1 + u[:, 7:] * 0.9), sc.lb, sc.ub)
for ne in (1, 8, 64, 512, 4096):
    gw.reset_counters()
    t = tm(lambda: gw.check_connection(a[:ne], b[:ne], 0.01), 3)
    c = gw.counters()
    print("edges n=%5d  wall %.3f ms  kernel %.3f ms/launch  states/launch %d  pairs/launch %.3g -> %.2f Tpair/s (kernel), %.2f TFLOP/s @10flop" % (
        ne, t * 1e3, c["kernel_ns"] / c["launches"] / 1e6, c["states"] // c["launches"], c["pair_tests"] / c["launches"],
        c["pair_tests"] / c["kernel_ns"] / 1e3, 10 * c["pair_tests"] / c["kernel_ns"] / 1e3))
# free-space edges only (all states evaluated)
ok = gw.check_connection(a, b, 0.01).astype(bool)
af, bf = a[ok], b[ok]
gw.reset_counters(); t = tm(lambda: gw.check_connection(af, bf, 0.01), 3); c = gw.counters()
print("free edges n=%d wall %.3f ms kernel %.3f ms  pairs/launch %.3g -> %.2f TFLOP/s" % (len(af), t * 1e3, c["kernel_ns"] / c["launches"] / 1e6, c["pair_tests"] / c["launches"], 10 * c["pair_tests"] / c["kernel_ns"] / 1e3))
st7 = g.NodeStore(7, 16384, segments=32)
for s in range(32):
    st7.append(W.uniform_samples(sc.lb, sc.ub, 50 + s, 0, 10000), segment=s)
q = W.uniform_samples(sc.lb, sc.ub, 99, 0, 32); seg = np.arange(32, dtype=np.uint32)
t = tm(lambda: g.extend_batch(st7, gw, q, seg, 0.5, 0.01, 1e-6, np.full(32, 2.3), 512), 5)
print("extend_batch 32 trees x 10k nodes: %.3f ms" % (t * 1e3))
t = tm(lambda: st7.nearest(q, seg=seg), 5); print("nn1 32 trees: %.3f ms" % (t * 1e3))
t = tm(lambda: st7.near(q, 2.3, cap=512, seg=seg), 5); print("near 32 trees: %.3f ms" % (t * 1e3))
t = tm(lambda: st7.knn(q, 100000, cap=10000, seg=seg), 3); print("knn-all 32 trees: %.3f ms" % (t * 1e3))
