"""ctypes wrapper of the CPU oracle (oracle/gc_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs;
never by anything under graph_core_b200/.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_u8p = C.POINTER(C.c_uint8)
_i64p = C.POINTER(C.c_longlong)
_u64p = C.POINTER(C.c_ulonglong)

_libs = {}


def build(force: bool = False):
    src = HERE / "gc_oracle.cpp"
    out = [HERE / "libgc_oracle.so", HERE / "libgc_oracle_fast.so"]
    if force or any((not o.exists()) or o.stat().st_mtime < src.stat().st_mtime for o in out):
        subprocess.run(["make", "-C", str(HERE)] + (["-B"] if force else []), check=True, capture_output=True)


def lib(fast: bool = False):
    """fast=False: strict-IEEE parity build; fast=True: the reference's -Ofast flags (timing only)."""
    key = "fast" if fast else "parity"
    if key in _libs:
        return _libs[key]
    path = HERE / ("libgc_oracle_fast.so" if fast else "libgc_oracle.so")
    if not path.exists():
        build()
    L = C.CDLL(str(path))
    vp = C.c_void_p
    sig = {
        "orc_nn_new": (vp, [C.c_int, C.c_int]),
        "orc_nn_free": (None, [vp]),
        "orc_nn_insert": (C.c_int, [vp, _dp]),
        "orc_nn_insert_many": (None, [vp, _dp, C.c_int]),
        "orc_nn_reinsert": (None, [vp, C.c_int]),
        "orc_nn_delete": (C.c_int, [vp, C.c_int]),
        "orc_nn_restore": (C.c_int, [vp, C.c_int]),
        "orc_nn_find": (C.c_int, [vp, C.c_int]),
        "orc_nn_clear": (C.c_int, [vp]),
        "orc_nn_size": (C.c_uint, [vp]),
        "orc_nn_deleted": (C.c_uint, [vp]),
        "orc_nn_set_deleted_threshold": (None, [vp, C.c_uint]),
        "orc_nn_get_nodes": (C.c_int, [vp, _ip, C.c_int]),
        "orc_nn_nearest": (None, [vp, _dp, C.c_int, _ip, _dp]),
        "orc_nn_near": (None, [vp, _dp, C.c_int, _dp, C.c_int, _ip, _dp, _ip]),
        "orc_nn_knn": (None, [vp, _dp, C.c_int, C.c_ulonglong, C.c_int, _ip, _dp, _ip]),
        "orc_world_cube": (vp, [C.c_int, C.c_double]),
        "orc_world_boxes": (vp, [C.c_int, C.c_int, _dp, _dp]),
        "orc_world_cspheres": (vp, [C.c_int, C.c_int, _fp, _fp]),
        "orc_world_sphere_arm": (vp, [C.c_int, _fp, _fp, C.c_int, _ip, _fp, C.c_int, _fp]),
        "orc_world_free": (None, [vp]),
        "orc_world_pair_tests": (C.c_ulonglong, [vp]),
        "orc_world_state_checks": (C.c_ulonglong, [vp]),
        "orc_world_reset_counters": (None, [vp]),
        "orc_world_enable_grid": (C.c_int, [vp, C.c_int]),
        "orc_check_states": (None, [vp, _dp, C.c_int, _u8p]),
        "orc_check_edges": (None, [vp, _dp, _dp, C.c_int, C.c_double, C.c_int, _u8p, _i64p, _dp]),
        "orc_check_from_conf": (None, [vp, _dp, _dp, _dp, C.c_int, C.c_double, _u8p]),
        "orc_sample_informed": (None, [C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_ulonglong, C.c_ulonglong, _dp,
                                       C.POINTER(C.c_uint)]),
        "orc_solver_new": (vp, [C.c_int, C.c_int, vp, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double,
                                C.c_double, C.c_int, C.c_int]),
        "orc_solver_free": (None, [vp]),
        "orc_solver_run": (C.c_int, [vp, _dp, C.c_int]),
        "orc_solver_update": (C.c_int, [vp, _dp]),
        "orc_solver_solved": (C.c_int, [vp]),
        "orc_solver_can_improve": (C.c_int, [vp]),
        "orc_solver_cost": (C.c_double, [vp]),
        "orc_solver_radius": (C.c_double, [vp]),
        "orc_solver_specific_volume": (C.c_double, [vp]),
        "orc_solver_num_nodes": (C.c_int, [vp]),
        "orc_solver_tree_size": (C.c_int, [vp]),
        "orc_solver_goal_id": (C.c_int, [vp]),
        "orc_solver_dump": (None, [vp, _ip, _dp, _dp]),
        "orc_solver_solution": (C.c_int, [vp, _ip, C.c_int]),
        "orc_solver_counters": (None, [vp, _u64p]),
        "orc_anytime_new": (vp, [C.c_int, vp, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                 C.c_double, C.c_double, C.c_double]),
        "orc_anytime_free": (None, [vp]),
        "orc_anytime_base": (vp, [vp]),
        "orc_anytime_solve": (C.c_int, [vp, C.c_uint, C.c_int, C.c_ulonglong, C.c_ulonglong, C.c_ulonglong]),
        "orc_anytime_rrt_update": (C.c_int, [vp, _dp]),
        "orc_anytime_improve_begin": (C.c_int, [vp]),
        "orc_anytime_improve_update": (C.c_int, [vp, _dp]),
        "orc_anytime_state": (None, [vp, _dp]),
        "orc_anytime_log": (C.c_int, [vp, C.c_int, _dp]),
        "orc_anytime_solution": (C.c_int, [vp, C.c_int, _dp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _libs[key] = L
    return L


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a.reshape(shape) if shape is not None else a


def _p(a, t):
    return a.ctypes.data_as(t)


class OracleNN:
    """kind 'vector' (src/datastructure/vector.cpp) or 'kdtree' (src/datastructure/kdtree.cpp)."""

    def __init__(self, kind: str, dim: int, fast: bool = False):
        self.L = lib(fast)
        self.dim = dim
        self.h = self.L.orc_nn_new(1 if kind == "kdtree" else 0, dim)

    def __del__(self):
        try:
            self.L.orc_nn_free(self.h)
        except Exception:
            pass

    def insert(self, q):
        q = _f64(q, (-1, self.dim))
        self.L.orc_nn_insert_many(self.h, _p(q, _dp), q.shape[0])

    def reinsert(self, node_id):
        self.L.orc_nn_reinsert(self.h, node_id)

    def delete(self, node_id):
        return bool(self.L.orc_nn_delete(self.h, node_id))

    def restore(self, node_id):
        return bool(self.L.orc_nn_restore(self.h, node_id))

    def find(self, node_id):
        return bool(self.L.orc_nn_find(self.h, node_id))

    def clear(self):
        return bool(self.L.orc_nn_clear(self.h))

    def size(self):
        return self.L.orc_nn_size(self.h)

    def deleted(self):
        return self.L.orc_nn_deleted(self.h)

    def set_deleted_threshold(self, t):
        self.L.orc_nn_set_deleted_threshold(self.h, t)

    def get_nodes(self, cap=1 << 22):
        ids = np.empty(cap, dtype=np.int32)
        n = self.L.orc_nn_get_nodes(self.h, _p(ids, _ip), cap)
        return ids[:n].copy()

    def nearest(self, q):
        q = _f64(q, (-1, self.dim))
        ids = np.empty(q.shape[0], dtype=np.int32)
        d = np.empty(q.shape[0], dtype=np.float64)
        self.L.orc_nn_nearest(self.h, _p(q, _dp), q.shape[0], _p(ids, _ip), _p(d, _dp))
        return ids, d

    def near(self, q, radius, cap=256):
        q = _f64(q, (-1, self.dim))
        nq = q.shape[0]
        r = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, dtype=np.float64), (nq,)))
        ids = np.full((nq, cap), -1, dtype=np.int32)
        d = np.full((nq, cap), np.inf, dtype=np.float64)
        cnt = np.zeros(nq, dtype=np.int32)
        self.L.orc_nn_near(self.h, _p(q, _dp), nq, _p(r, _dp), cap, _p(ids, _ip), _p(d, _dp), _p(cnt, _ip))
        return ids, d, cnt

    def knn(self, q, k, cap):
        q = _f64(q, (-1, self.dim))
        nq = q.shape[0]
        ids = np.full((nq, cap), -1, dtype=np.int32)
        d = np.full((nq, cap), np.inf, dtype=np.float64)
        cnt = np.zeros(nq, dtype=np.int32)
        self.L.orc_nn_knn(self.h, _p(q, _dp), nq, int(k), cap, _p(ids, _ip), _p(d, _dp), _p(cnt, _ip))
        return ids, d, cnt


class OracleWorld:
    def __init__(self, h, dim, L, keep=()):
        self.h, self.dim, self.L, self._keep = h, dim, L, keep

    @classmethod
    def cube(cls, dim, thr=1.0, fast=False):
        L = lib(fast)
        return cls(L.orc_world_cube(dim, thr), dim, L)

    @classmethod
    def boxes(cls, lo, hi, fast=False):
        L = lib(fast)
        lo, hi = _f64(lo), _f64(hi)
        return cls(L.orc_world_boxes(lo.shape[1], lo.shape[0], _p(lo, _dp), _p(hi, _dp)), lo.shape[1], L)

    @classmethod
    def cspheres(cls, centers, radii, fast=False):
        L = lib(fast)
        c = np.ascontiguousarray(centers, np.float32)
        r = np.ascontiguousarray(radii, np.float32)
        return cls(L.orc_world_cspheres(c.shape[1], c.shape[0], _p(c, _fp), _p(r, _fp)), c.shape[1], L)

    @classmethod
    def sphere_arm(cls, base_tf, joint_tf, rs_link, rs_local, obs, fast=False):
        L = lib(fast)
        b = np.ascontiguousarray(base_tf, np.float32).reshape(12)
        j = np.ascontiguousarray(joint_tf, np.float32).reshape(-1, 12)
        l = np.ascontiguousarray(rs_link, np.int32)
        lo = np.ascontiguousarray(rs_local, np.float32).reshape(-1, 4)
        o = np.ascontiguousarray(obs, np.float32).reshape(-1, 4)
        h = L.orc_world_sphere_arm(j.shape[0], _p(b, _fp), _p(j, _fp), lo.shape[0], _p(l, _ip), _p(lo, _fp), o.shape[0],
                                   _p(o, _fp))
        return cls(h, j.shape[0], L)

    @classmethod
    def from_scenario(cls, sc, fast=False):
        p = sc.params
        if sc.kind == "cube":
            return cls.cube(sc.dim, p["threshold"], fast)
        if sc.kind == "boxes":
            return cls.boxes(p["lo"], p["hi"], fast)
        if sc.kind == "cspheres":
            return cls.cspheres(p["centers"], p["radii"], fast)
        return cls.sphere_arm(p["base_tf"], p["joint_tf"], p["rs_link"], p["rs_local"], p["obs"], fast)

    def __del__(self):
        try:
            self.L.orc_world_free(self.h)
        except Exception:
            pass

    def counters(self):
        return {"states": self.L.orc_world_state_checks(self.h), "pair_tests": self.L.orc_world_pair_tests(self.h)}

    def reset_counters(self):
        self.L.orc_world_reset_counters(self.h)

    def enable_grid(self, on=True):
        """Uniform-grid culling of the sphere-arm check (verdicts unchanged); returns the number of cells (0 = off)."""
        return self.L.orc_world_enable_grid(self.h, 1 if on else 0)

    def check(self, q):
        q = _f64(q, (-1, self.dim))
        ok = np.empty(q.shape[0], np.uint8)
        self.L.orc_check_states(self.h, _p(q, _dp), q.shape[0], _p(ok, _u8p))
        return ok

    def check_connection(self, c1, c2, min_distance=0.01, mode=0):
        c1, c2 = _f64(c1, (-1, self.dim)), _f64(c2, (-1, self.dim))
        n = c1.shape[0]
        ok = np.empty(n, np.uint8)
        if mode == 1:
            fb = np.empty(n, np.int64)
            conf = np.full((n, self.dim), np.nan)
            self.L.orc_check_edges(self.h, _p(c1, _dp), _p(c2, _dp), n, min_distance, 1, _p(ok, _u8p),
                                   fb.ctypes.data_as(_i64p), _p(conf, _dp))
            return ok, fb, conf
        self.L.orc_check_edges(self.h, _p(c1, _dp), _p(c2, _dp), n, min_distance, 0, _p(ok, _u8p), None, None)
        return ok

    def check_conn_from_conf(self, parent, child, conf, min_distance=0.01):
        p, c, f = _f64(parent, (-1, self.dim)), _f64(child, (-1, self.dim)), _f64(conf, (-1, self.dim))
        res = np.empty(p.shape[0], np.uint8)
        self.L.orc_check_from_conf(self.h, _p(p, _dp), _p(c, _dp), _p(f, _dp), p.shape[0], min_distance, _p(res, _u8p))
        return res


def sample_informed(lb, ub, f1, f2, cost, seed, first=0, fast=False):
    """gc_sample_informed restated on the CPU: f1/f2 [n][dim], cost [n] -> samples [n][dim], attempts [n]."""
    L = lib(fast)
    lb, ub = _f64(lb), _f64(ub)
    dim = lb.shape[0]
    f1, f2 = _f64(f1, (-1, dim)), _f64(f2, (-1, dim))
    n = f1.shape[0]
    cost = np.ascontiguousarray(np.broadcast_to(np.asarray(cost, dtype=np.float64), (n,)))
    out = np.empty((n, dim))
    att = np.empty(n, np.uint32)
    L.orc_sample_informed(dim, n, _p(lb, _dp), _p(ub, _dp), _p(f1, _dp), _p(f2, _dp), _p(cost, _dp), seed, first,
                          _p(out, _dp), att.ctypes.data_as(C.POINTER(C.c_uint)))
    return out, att


def arm_checker(fast=False):
    """checker(params, q) for worlds.scenario_c3."""

    def chk(params, q):
        w = OracleWorld.sphere_arm(params["base_tf"], params["joint_tf"], params["rs_link"], params["rs_local"],
                                   params["obs"], fast)
        return w.check(q).astype(bool)

    return chk


class OracleSolver:
    RRT, RRTSTAR = 0, 1

    def __init__(self, kind, scenario, world: OracleWorld, start=None, goal=None, use_kdtree=True, extend=False,
                 fast=False):
        self.L = lib(fast)
        self.world = world
        self.dim = scenario.dim
        s = _f64(scenario.start if start is None else start)
        g = _f64(scenario.goal if goal is None else goal)
        lb, ub = _f64(scenario.lb), _f64(scenario.ub)
        self.h = self.L.orc_solver_new(kind, self.dim, world.h, _p(lb, _dp), _p(ub, _dp), _p(s, _dp), _p(g, _dp),
                                       scenario.max_distance, scenario.min_distance, scenario.rewire_factor,
                                       scenario.utopia_tolerance, 1 if use_kdtree else 0, 1 if extend else 0)

    def __del__(self):
        try:
            self.L.orc_solver_free(self.h)
        except Exception:
            pass

    def run(self, samples):
        s = _f64(samples, (-1, self.dim))
        return self.L.orc_solver_run(self.h, _p(s, _dp), s.shape[0])

    def update(self, q):
        q = _f64(q)
        return bool(self.L.orc_solver_update(self.h, _p(q, _dp)))

    @property
    def solved(self):
        return bool(self.L.orc_solver_solved(self.h))

    @property
    def can_improve(self):
        return bool(self.L.orc_solver_can_improve(self.h))

    @property
    def cost(self):
        return self.L.orc_solver_cost(self.h)

    @property
    def radius(self):
        return self.L.orc_solver_radius(self.h)

    @property
    def specific_volume(self):
        return self.L.orc_solver_specific_volume(self.h)

    @property
    def tree_size(self):
        return self.L.orc_solver_tree_size(self.h)

    def dump(self):
        n = self.L.orc_solver_num_nodes(self.h)
        par = np.empty(n, np.int32)
        pc = np.empty(n, np.float64)
        conf = np.empty((n, self.dim), np.float64)
        self.L.orc_solver_dump(self.h, _p(par, _ip), _p(pc, _dp), _p(conf, _dp))
        return par, pc, conf

    def solution(self):
        ids = np.empty(1 << 16, np.int32)
        n = self.L.orc_solver_solution(self.h, _p(ids, _ip), ids.shape[0])
        return ids[:n].copy()

    def counters(self):
        out = (C.c_ulonglong * 7)()
        self.L.orc_solver_counters(self.h, out)
        keys = ["nn1", "near", "knn", "edges", "near_hits", "state_checks", "pair_tests"]
        return dict(zip(keys, [int(v) for v in out]))


ANYTIME_LOG_COLUMNS = ("phase", "iterations", "success", "path_cost", "tree_nodes", "bias", "cost2beat", "set_up")


class OracleAnytime:
    """AnytimeRRT (src/graph_core/solvers/anytime_rrt.cpp) restated; sample k of problem p of P comes from the informed
    Philox sampler with counter k * P + p (foci start/goal, cost = sampler_cost0 in the RRT phase, cost2beat in a round).
    """

    def __init__(self, scenario, world: OracleWorld, start=None, goal=None, use_kdtree=True, extend=False, delta=0.1,
                 cost_impr=0.1, sampler_cost0=float("inf"), fast=False):
        self.L = lib(fast)
        self.world = world
        self.dim = scenario.dim
        s = _f64(scenario.start if start is None else start)
        g = _f64(scenario.goal if goal is None else goal)
        lb, ub = _f64(scenario.lb), _f64(scenario.ub)
        self.h = self.L.orc_anytime_new(self.dim, world.h, _p(lb, _dp), _p(ub, _dp), _p(s, _dp), _p(g, _dp),
                                        scenario.max_distance, scenario.min_distance, scenario.utopia_tolerance,
                                        1 if use_kdtree else 0, 1 if extend else 0, delta, cost_impr, sampler_cost0)

    def __del__(self):
        try:
            self.L.orc_anytime_free(self.h)
        except Exception:
            pass

    def solve(self, max_iter, seed, p=0, P=1, improve_in_solve=True):
        return bool(self.L.orc_anytime_solve(self.h, max_iter, 1 if improve_in_solve else 0, seed, p, P))

    def rrt_update(self, q):
        q = _f64(q)
        return bool(self.L.orc_anytime_rrt_update(self.h, _p(q, _dp)))

    def improve_begin(self):
        return bool(self.L.orc_anytime_improve_begin(self.h))

    def improve_update(self, q):
        q = _f64(q)
        return bool(self.L.orc_anytime_improve_update(self.h, _p(q, _dp)))

    def state(self):
        out = np.zeros(8)
        self.L.orc_anytime_state(self.h, _p(out, _dp))
        keys = ["bias", "cost2beat", "path_cost", "cost", "can_improve", "solved", "round_tree_nodes", "samples_drawn"]
        return dict(zip(keys, out.tolist()))

    def log(self):
        n = self.L.orc_anytime_log(self.h, 0, None)
        rows = np.zeros((max(n, 1), 8))
        self.L.orc_anytime_log(self.h, n, _p(rows, _dp))
        return rows[:n]

    def solution(self):
        wp = np.empty((1 << 14, self.dim))
        n = self.L.orc_anytime_solution(self.h, wp.shape[0], _p(wp, _dp))
        return wp[:n].copy()

    def dump(self):
        """parents, connection costs, configurations of every node the solver created (ids = creation order); the
        tree in use is the one rooted at the last solution's first waypoint."""
        b = self.L.orc_anytime_base(self.h)
        n = self.L.orc_solver_num_nodes(b)
        par = np.empty(n, np.int32)
        pc = np.empty(n, np.float64)
        conf = np.empty((n, self.dim), np.float64)
        self.L.orc_solver_dump(b, _p(par, _ip), _p(pc, _dp), _p(conf, _dp))
        return par, pc, conf
