"""ctypes binding of the C++ host layer (graph_core_b200/host/lockstep.h -> libgcb200_host.so).

LockstepRRTStar advances many independent RRT* problems (RRTStar::update,
src/graph_core/solvers/rrt_star.cpp:89-163) in lock step; the hot paths of all problems are batched
into a few CUDA launches per iteration through the C ABI of libgcb200.so.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import capi

_PKG = Path(__file__).resolve().parent
HOSTLIB_PATH = _PKG / "libgcb200_host.so"
_dp = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_hlib = None


def load_host():
    global _hlib
    if _hlib is not None:
        return _hlib
    capi.load()  # libgcb200.so first (no CPU fallback)
    if not HOSTLIB_PATH.exists():
        raise RuntimeError("%s is missing: build it with `python -m graph_core_b200._build`" % HOSTLIB_PATH)
    L = C.CDLL(str(HOSTLIB_PATH))
    vp = C.c_void_p
    sig = {
        "gcp_rrtstar_create": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _dp, _dp, vp, C.c_int,
                                         C.c_uint32, _dp, _dp, C.c_uint32, C.c_int, C.POINTER(vp)]),
        "gcp_rrtstar_create2": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _dp, _dp, vp, C.c_int,
                                          C.c_uint32, _dp, _dp, C.c_uint32, C.c_int, C.c_int, C.POINTER(vp)]),
        "gcp_rrtstar_step_informed": (C.c_int, [vp, C.c_uint64, C.c_uint64]),
        "gcp_rrtstar_sync": (C.c_int, [vp]),
        "gcp_rrtstar_read_costs_async": (C.c_int, [vp, vp]),
        "gcp_rrtstar_destroy": (C.c_int, [vp]),
        "gcp_rrtstar_step": (C.c_int, [vp, _dp]),
        "gcp_rrtstar_step_dsamples": (C.c_int, [vp, vp]),
        "gcp_rrtstar_set_stream": (C.c_int, [vp, vp]),
        "gcp_rrtstar_run": (C.c_int, [vp, _dp, C.c_uint32]),
        "gcp_rrtstar_info": (C.c_int, [vp, C.c_uint32, _i32p, _i32p, _dp, _dp, _u32p, _u32p]),
        "gcp_rrtstar_dump": (C.c_int, [vp, C.c_uint32, _i32p, _dp, _dp]),
        "gcp_rrtstar_solution": (C.c_int, [vp, C.c_uint32, _i32p, C.c_uint32]),
        "gcp_rrtstar_stats": (C.c_int, [vp, _u64p]),
        "gcp_last_error": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _hlib = L
    return L


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a.reshape(shape) if shape is not None else a


class LockstepRRTStar:
    """P independent RRT* problems sharing one collision world, advanced in lock step.

    device_trees=True (default): parent slots and connection costs live on the GPU (gc_forest); `step*` only enqueue
    work on the stream and the accessors synchronise on demand.  device_trees=False: the host replay of round 1."""

    def __init__(self, scenario, world: capi.CollisionWorld, starts, goals, capacity: int = 16384, device: int = 0,
                 threads: int = 0, device_trees: bool = True, near_cap: int | None = None):
        self.L = load_host()
        self.dim = scenario.dim
        self.max_distance = float(scenario.max_distance)
        self.world = world  # keep alive
        starts = _f64(starts, (-1, self.dim))
        goals = _f64(goals, (-1, self.dim))
        assert starts.shape == goals.shape
        self.n = starts.shape[0]
        lb, ub = _f64(scenario.lb), _f64(scenario.ub)
        h = C.c_void_p()
        import os
        old_cap = os.environ.get("GCB200_NEAR_CAP")
        if near_cap is not None:  # longest radius-near list the device forest can hold (default 512 entries)
            os.environ["GCB200_NEAR_CAP"] = str(int(near_cap))
        try:
            rc = self._create(scenario, world, device, starts, goals, capacity, threads, device_trees, lb, ub, h)
        finally:
            if near_cap is not None:
                if old_cap is None:
                    os.environ.pop("GCB200_NEAR_CAP", None)
                else:
                    os.environ["GCB200_NEAR_CAP"] = old_cap
        if rc != 0:
            raise capi.GcError(rc, self.L.gcp_last_error().decode())
        self.h = h

    def _create(self, scenario, world, device, starts, goals, capacity, threads, device_trees, lb, ub, h):
        return self.L.gcp_rrtstar_create2(self.dim, scenario.max_distance, scenario.min_distance, scenario.rewire_factor,
                                        scenario.utopia_tolerance, lb.ctypes.data_as(_dp), ub.ctypes.data_as(_dp),
                                        world.handle, device, self.n, starts.ctypes.data_as(_dp),
                                        goals.ctypes.data_as(_dp), capacity, threads, 1 if device_trees else 0,
                                        C.byref(h))

    def close(self):
        if getattr(self, "h", None):
            self.L.gcp_rrtstar_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, samples) -> int:
        s = _f64(samples, (self.n, self.dim))
        rc = self.L.gcp_rrtstar_step(self.h, s.ctypes.data_as(_dp))
        if rc < 0:
            raise capi.GcError(rc, self.L.gcp_last_error().decode())
        return rc

    def step_device_samples(self, d_samples_ptr: int) -> int:
        """d_samples_ptr: device address of a [problems, dim] float64 block (e.g. torch tensor .data_ptr())."""
        rc = self.L.gcp_rrtstar_step_dsamples(self.h, C.c_void_p(d_samples_ptr))
        if rc < 0:
            raise capi.GcError(rc, self.L.gcp_last_error().decode())
        return rc

    def step_informed(self, seed: int, first_index: int) -> int:
        """One iteration whose samples come from the device informed sampler (problem p: Philox counter first_index+p)."""
        rc = self.L.gcp_rrtstar_step_informed(self.h, seed, first_index)
        if rc < 0:
            raise capi.GcError(rc, self.L.gcp_last_error().decode())
        return rc

    def sync(self):
        rc = self.L.gcp_rrtstar_sync(self.h)
        if rc < 0:
            raise capi.GcError(rc, self.L.gcp_last_error().decode())

    def read_costs_async(self, pinned_ptr: int):
        """Enqueue a D2H copy of the P current path costs into pinned host memory at `pinned_ptr`."""
        rc = self.L.gcp_rrtstar_read_costs_async(self.h, C.c_void_p(pinned_ptr))
        if rc < 0:
            raise capi.GcError(rc, self.L.gcp_last_error().decode())

    def set_stream(self, cuda_stream: int | None):
        rc = self.L.gcp_rrtstar_set_stream(self.h, C.c_void_p(capi.stream_handle(cuda_stream)))
        if rc < 0:
            raise capi.GcError(rc, self.L.gcp_last_error().decode())

    def run(self, samples) -> int:
        """samples: [iterations, problems, dim]; returns the number of planner updates executed."""
        s = _f64(samples)
        assert s.ndim == 3 and s.shape[1] == self.n and s.shape[2] == self.dim
        rc = self.L.gcp_rrtstar_run(self.h, s.ctypes.data_as(_dp), s.shape[0])
        if rc < 0:
            raise capi.GcError(rc, self.L.gcp_last_error().decode())
        return rc

    def info(self, p: int = 0):
        solved, can = C.c_int32(0), C.c_int32(0)
        cost, radius = C.c_double(0), C.c_double(0)
        nn, ts = C.c_uint32(0), C.c_uint32(0)
        self.L.gcp_rrtstar_info(self.h, p, C.byref(solved), C.byref(can), C.byref(cost), C.byref(radius), C.byref(nn),
                                C.byref(ts))
        return {"solved": bool(solved.value), "can_improve": bool(can.value), "cost": cost.value,
                "radius": radius.value, "num_nodes": nn.value, "tree_size": ts.value}

    def dump(self, p: int = 0):
        n = self.info(p)["num_nodes"]
        par = np.empty(n, np.int32)
        pc = np.empty(n, np.float64)
        conf = np.empty((n, self.dim), np.float64)
        self.L.gcp_rrtstar_dump(self.h, p, par.ctypes.data_as(_i32p), pc.ctypes.data_as(_dp), conf.ctypes.data_as(_dp))
        return par, pc, conf

    def to_yaml(self, p: int = 0, use_kdtree: bool = True) -> str:
        """The tree of problem p in the reference's on-disk format (`Tree::toYAML`, tree.cpp:1188-1227)."""
        from .tree_yaml import tree_to_yaml
        par, _, conf = self.dump(p)
        # nodes are listed in slot (creation) order; a goal that is not connected yet (parent -1, not the root) is not
        # part of the tree and is left out -- the reference inserts it only when it connects
        keep = (par >= 0) | (np.arange(par.shape[0]) == 0)
        new_id = np.cumsum(keep) - 1
        par = np.where(par[keep] >= 0, new_id[np.maximum(par[keep], 0)], -1)
        return tree_to_yaml(conf[keep], par, self.max_distance, use_kdtree)

    def solution(self, p: int = 0):
        ids = np.empty(1 << 16, np.int32)
        n = self.L.gcp_rrtstar_solution(self.h, p, ids.ctypes.data_as(_i32p), ids.shape[0])
        return ids[:n].copy()

    def stats(self):
        out = (C.c_uint64 * 16)()
        self.L.gcp_rrtstar_stats(self.h, out)
        keys = ["iterations", "extended", "edges_device", "edges_used", "edges_fallback", "near_hits", "gpu_calls",
                "h2d_bytes", "d2h_bytes", "_", "ns_prologue", "ns_fused_pass", "ns_candidates", "ns_append_edges",
                "ns_replay"]
        return dict(zip(keys, [int(v) for v in out]))


class BirrtConfig(C.Structure):
    _fields_ = [("max_distance", C.c_double), ("min_distance", C.c_double), ("tolerance", C.c_double),
                ("bidirectional", C.c_int32), ("extend", C.c_int32), ("capacity", C.c_uint32), ("depth_cap", C.c_uint32)]


class LockstepBiRRT:
    """P independent BiRRT (or, bidirectional=False, RRT) problems advanced together, entirely on the device: one kernel
    launch per update, one warp per tree, the whole Tree::connect ray inside the kernel (csrc/birrt.cu; reference:
    src/graph_core/solvers/birrt.cpp:93-152, rrt.cpp:119-157, tree.cpp:217-233).  Light worlds only."""

    STATS = ("updates", "nodes_added", "edges", "states", "object_tests", "scans", "rescans", "solved")

    def __init__(self, scenario, world: capi.CollisionWorld, starts, goals, capacity: int = 4096, bidirectional: bool = True,
                 extend: bool = False, depth_cap: int = 0):
        self.L = capi.load()
        self.dim = scenario.dim
        starts = _f64(np.atleast_2d(starts))
        goals = _f64(np.atleast_2d(goals))
        assert starts.shape == goals.shape and starts.shape[1] == self.dim
        self.n = starts.shape[0]
        self.world = world
        cfg = BirrtConfig(scenario.max_distance, scenario.min_distance, 1e-6, 1 if bidirectional else 0,
                          1 if extend else 0, int(capacity), int(depth_cap))
        h = C.c_void_p()
        capi._check(self.L.gc_birrt_create(world.handle, C.byref(cfg), self.n, starts.ctypes.data_as(_dp),
                                           goals.ctypes.data_as(_dp), C.byref(h)))
        self.h = h
        self.capacity = int(capacity)

    def close(self):
        if getattr(self, "h", None):
            self.L.gc_birrt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, samples):
        """One update of every unsolved problem; samples[p] is problem p's sample (host array, staged through pinned
        memory)."""
        s = _f64(samples)
        assert s.shape == (self.n, self.dim)
        capi._check(self.L.gc_birrt_step(self.h, s.ctypes.data_as(_dp), None))

    def step_device_samples(self, d_samples_ptr: int):
        capi._check(self.L.gc_birrt_step(self.h, None, C.c_void_p(d_samples_ptr)))

    def sync(self):
        capi._check(self.L.gc_birrt_sync(self.h))

    def stats(self):
        out = (C.c_uint64 * 8)()
        capi._check(self.L.gc_birrt_stats(self.h, out))
        return dict(zip(self.STATS, (int(v) for v in out)))

    def info(self, p: int):
        solved, it, cost = C.c_int32(0), C.c_uint32(0), C.c_double(0.0)
        ns, ng = C.c_uint32(0), C.c_uint32(0)
        capi._check(self.L.gc_birrt_info(self.h, p, C.byref(solved), C.byref(it), C.byref(cost), C.byref(ns), C.byref(ng)))
        return {"solved": bool(solved.value), "solved_iteration": it.value, "cost": cost.value, "size_start": ns.value,
                "size_goal": ng.value}

    def solved(self):
        s = np.zeros(self.n, dtype=np.uint8)
        c = np.zeros(self.n, dtype=np.float64)
        capi._check(self.L.gc_birrt_read_solved(self.h, s.ctypes.data_as(C.POINTER(C.c_uint8)), c.ctypes.data_as(_dp)))
        return s.astype(bool), c

    def dump(self, p: int, tree: int):
        """(parents, connection costs, configurations) of the start (0) or goal (1) tree of problem p in slot order."""
        par = np.empty(self.capacity, dtype=np.int32)
        pc = np.empty(self.capacity, dtype=np.float64)
        cf = np.empty((self.capacity, self.dim), dtype=np.float64)
        n = self.L.gc_birrt_dump(self.h, p, tree, par.ctypes.data_as(C.POINTER(C.c_int32)), pc.ctypes.data_as(_dp),
                                 cf.ctypes.data_as(_dp))
        if n < 0:
            raise capi.GcError(n, self.L.gc_last_error().decode())
        return par[:n].copy(), pc[:n].copy(), cf[:n].copy()

    def solution(self, p: int, cap: int = 8192):
        wp = np.empty((cap, self.dim), dtype=np.float64)
        cs = np.empty(cap, dtype=np.float64)
        n = self.L.gc_birrt_solution(self.h, p, wp.ctypes.data_as(_dp), cs.ctypes.data_as(_dp), cap)
        if n < 0:
            raise capi.GcError(n, self.L.gc_last_error().decode())
        return wp[:n].copy(), cs[:max(n - 1, 0)].copy()


def query_shard(total: int, rank: int, world: int) -> np.ndarray:
    """BASELINE config 5's partition of `total` independent queries over `world` GPUs: query j runs on GPU j mod world.
    The returned positions are also the queries' sample-stream indices (LockstepAnytimeRRT.set_streams(total, idx))."""
    return np.arange(rank, total, world, dtype=np.int64)


def gather_query_results(dist_mod, total: int, mine: np.ndarray, solved: np.ndarray, cost: np.ndarray, device=None):
    """Whole-set (solved, cost) arrays from the per-rank shards of query_shard: one all-reduce of two vectors of length
    `total` (every query is owned by exactly one rank).  dist_mod = torch.distributed or None for a single process."""
    import torch
    full_s = torch.zeros(total, dtype=torch.float64, device=device)
    full_c = torch.zeros(total, dtype=torch.float64, device=device)
    idx = torch.as_tensor(np.asarray(mine), dtype=torch.long, device=device)
    full_s[idx] = torch.as_tensor(np.asarray(solved, dtype=np.float64), device=device)
    c = np.where(np.asarray(solved, dtype=bool), np.asarray(cost, dtype=np.float64), 0.0)
    full_c[idx] = torch.as_tensor(c, device=device)
    if dist_mod is not None and dist_mod.is_initialized() and dist_mod.get_world_size() > 1:
        dist_mod.all_reduce(full_s)
        dist_mod.all_reduce(full_c)
    s_np = full_s.cpu().numpy() > 0.5
    c_np = np.where(s_np, full_c.cpu().numpy(), np.inf)
    return s_np, c_np


class AnytimeConfig(C.Structure):
    _fields_ = [("max_distance", C.c_double), ("min_distance", C.c_double), ("tolerance", C.c_double),
                ("utopia_tolerance", C.c_double), ("delta", C.c_double), ("cost_impr", C.c_double), ("seed", C.c_uint64),
                ("max_iter", C.c_uint32), ("capacity", C.c_uint32), ("extend", C.c_int32),
                ("improve_in_solve", C.c_int32), ("keep_trees", C.c_int32), ("reserved", C.c_int32)]


class LockstepAnytimeRRT:
    """P independent AnytimeRRT queries (src/graph_core/solvers/anytime_rrt.cpp, BASELINE config 5) advanced together:
    RRT::update during the first phase, AnytimeRRT::improve rounds (Tree::informedExtend) afterwards; csrc/anytime.cu.
    Problem p draws its k-th sample from the informed Philox sampler with counter k * P + p."""

    STATS = ("steps", "samples", "rrt_rows", "improve_rows", "nodes_added", "goal_edges", "improve_edges", "solved")
    LOG_COLUMNS = ("phase", "iterations", "success", "path_cost", "tree_nodes", "bias", "cost2beat", "set_up")

    def __init__(self, scenario, world: capi.CollisionWorld, starts, goals, max_iter: int, seed: int,
                 sampler_cost0=None, capacity: int = 8192, extend: bool = False, improve_in_solve: bool = True,
                 delta: float = 0.1, cost_impr: float = 0.1, keep_trees: bool = False):
        self.L = capi.load()
        self.dim = scenario.dim
        starts = _f64(np.atleast_2d(starts))
        goals = _f64(np.atleast_2d(goals))
        assert starts.shape == goals.shape and starts.shape[1] == self.dim
        self.n = starts.shape[0]
        self.world = world
        lb, ub = _f64(scenario.lb), _f64(scenario.ub)
        c0 = None
        if sampler_cost0 is not None:
            c0 = np.ascontiguousarray(np.broadcast_to(np.asarray(sampler_cost0, dtype=np.float64), (self.n,)))
        cfg = AnytimeConfig(scenario.max_distance, scenario.min_distance, 1e-6, scenario.utopia_tolerance, delta,
                            cost_impr, int(seed), int(max_iter), int(capacity), 1 if extend else 0,
                            1 if improve_in_solve else 0, 1 if keep_trees else 0, 0)
        h = C.c_void_p()
        capi._check(self.L.gc_anytime_create(world.handle, C.byref(cfg), lb.ctypes.data_as(_dp), ub.ctypes.data_as(_dp),
                                             self.n, starts.ctypes.data_as(_dp), goals.ctypes.data_as(_dp),
                                             c0.ctypes.data_as(_dp) if c0 is not None else None, C.byref(h)))
        self.h = h
        self.capacity = int(capacity)

    def close(self):
        if getattr(self, "h", None):
            self.L.gc_anytime_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _rc(self, r):
        if r < 0:
            raise capi.GcError(r, self.L.gc_last_error().decode())
        return r

    def set_streams(self, stride: int, index):
        """These queries are positions `index` of a set of `stride` queries (sharding over GPUs): sample k of local
        query i uses Philox counter k * stride + index[i]."""
        idx = np.ascontiguousarray(index, dtype=np.uint64)
        assert idx.shape == (self.n,)
        capi._check(self.L.gc_anytime_set_streams(self.h, int(stride), idx.ctypes.data_as(C.POINTER(C.c_uint64))))

    def step(self) -> int:
        """One planner update of every live query; returns how many are still live."""
        return self._rc(self.L.gc_anytime_step(self.h))

    def run(self, max_steps: int = 0) -> int:
        """AnytimeRRT::solve for all queries; returns how many are still live (0 unless max_steps cut the run)."""
        return self._rc(self.L.gc_anytime_run(self.h, int(max_steps)))

    def stats(self):
        out = (C.c_uint64 * 8)()
        capi._check(self.L.gc_anytime_stats(self.h, out))
        return dict(zip(self.STATS, (int(v) for v in out)))

    def timing(self):
        out = (C.c_double * 8)()
        capi._check(self.L.gc_anytime_timing(self.h, out))
        keys = ("control", "rrt_rows", "improve_rows", "wait", "decisions", "appends", "goal_edges")
        return dict(zip(keys, (float(v) for v in out)))

    def info(self, p: int):
        solved, done = C.c_int32(0), C.c_int32(0)
        cost, bias, c2b = C.c_double(0), C.c_double(0), C.c_double(0)
        k = C.c_uint64(0)
        nodes, rows = C.c_uint32(0), C.c_uint32(0)
        capi._check(self.L.gc_anytime_info(self.h, p, C.byref(solved), C.byref(done), C.byref(cost), C.byref(bias),
                                           C.byref(c2b), C.byref(k), C.byref(nodes), C.byref(rows)))
        return {"solved": bool(solved.value), "done": bool(done.value), "cost": cost.value, "bias": bias.value,
                "cost2beat": c2b.value, "samples": k.value, "tree_nodes": nodes.value, "rounds": rows.value}

    def solved(self):
        s = np.zeros(self.n, dtype=np.uint8)
        c = np.zeros(self.n, dtype=np.float64)
        capi._check(self.L.gc_anytime_read_solved(self.h, s.ctypes.data_as(C.POINTER(C.c_uint8)), c.ctypes.data_as(_dp)))
        return s.astype(bool), c

    def log(self, p: int):
        n = self._rc(self.L.gc_anytime_log(self.h, p, None, 0))
        rows = np.zeros((max(n, 1), 8))
        self._rc(self.L.gc_anytime_log(self.h, p, rows.ctypes.data_as(_dp), n))
        return rows[:n]

    def solution(self, p: int, cap: int = 8192):
        wp = np.empty((cap, self.dim), dtype=np.float64)
        cs = np.empty(cap, dtype=np.float64)
        n = self._rc(self.L.gc_anytime_solution(self.h, p, wp.ctypes.data_as(_dp), cs.ctypes.data_as(_dp), cap))
        return wp[:n].copy(), cs[:max(n - 1, 0)].copy()

    def dump(self, p: int, which: int = 0):
        """(parents, connection costs, configurations) of the growing tree (0) or the tree holding the solution (1)."""
        cap = self.capacity + 1
        par = np.empty(cap, dtype=np.int32)
        pc = np.empty(cap, dtype=np.float64)
        cf = np.empty((cap, self.dim), dtype=np.float64)
        n = self._rc(self.L.gc_anytime_dump(self.h, p, which, par.ctypes.data_as(C.POINTER(C.c_int32)),
                                            pc.ctypes.data_as(_dp), cf.ctypes.data_as(_dp), cap))
        return par[:n].copy(), pc[:n].copy(), cf[:n].copy()
