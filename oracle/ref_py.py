"""ctypes wrapper of oracle/_ref/libgc_ref.so -- the graph_core reference's own sources compiled against
the stand-in headers of oracle/ref_shim (see oracle/ref_build.py, oracle/ref_driver.cpp).

TEST INFRASTRUCTURE ONLY: used by tests/ (to pin the oracle restatement), tests/golden/make_ref_golden.py
and bench.py's cpu_baseline / --impl reference legs.  Never imported by graph_core_b200/.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_u8p = C.POINTER(C.c_uint8)
_libs = {}


def available(fast: bool = False) -> bool:
    return (HERE / "_ref" / ("libgc_ref_fast.so" if fast else "libgc_ref.so")).exists()


def build():
    """(Re)build when the reference tree is present; otherwise rely on the prebuilt library."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_build", HERE / "ref_build.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.build()


def lib(fast: bool = False, dropin: bool = False):
    key = "dropin" if dropin else ("fast" if fast else "parity")
    if key in _libs:
        return _libs[key]
    path = HERE / "_ref" / ("libgc_ref_dropin.so" if dropin else ("libgc_ref_fast.so" if fast else "libgc_ref.so"))
    if not path.exists():
        build()
    if not path.exists():
        raise FileNotFoundError("%s missing: run oracle/ref_build.py where /root/reference exists" % path)
    L = C.CDLL(str(path))
    vp = C.c_void_p
    sig = {
        "ref_about": (C.c_char_p, []),
        "ref_nn_new": (vp, [C.c_int, C.c_int]),
        "ref_nn_free": (None, [vp]),
        "ref_nn_insert_many": (None, [vp, _dp, C.c_int]),
        "ref_nn_reinsert": (None, [vp, C.c_int]),
        "ref_nn_delete": (C.c_int, [vp, C.c_int]),
        "ref_nn_restore": (C.c_int, [vp, C.c_int]),
        "ref_nn_find": (C.c_int, [vp, C.c_int]),
        "ref_nn_clear": (C.c_int, [vp]),
        "ref_nn_size": (C.c_uint, [vp]),
        "ref_nn_set_deleted_threshold": (None, [vp, C.c_uint]),
        "ref_nn_get_nodes": (C.c_int, [vp, _ip, C.c_int]),
        "ref_nn_nearest": (None, [vp, _dp, C.c_int, _ip, _dp]),
        "ref_nn_near": (None, [vp, _dp, C.c_int, _dp, C.c_int, _ip, _dp, _ip]),
        "ref_nn_knn": (None, [vp, _dp, C.c_int, C.c_ulonglong, C.c_int, _ip, _dp, _ip]),
        "ref_checker_cube": (vp, [C.c_int, C.c_double, C.c_double]),
        "ref_checker_callback": (vp, [C.c_int, vp, vp, C.c_double]),
        "ref_checker_free": (None, [vp]),
        "ref_checker_states": (C.c_ulonglong, [vp]),
        "ref_check_states": (None, [vp, _dp, C.c_int, _u8p]),
        "ref_check_edges": (None, [vp, _dp, _dp, C.c_int, C.c_int, _u8p, _dp]),
        "ref_check_from_conf": (None, [vp, _dp, _dp, _dp, C.c_int, _u8p]),
        "ref_solver_new": (vp, [C.c_int, C.c_int, vp, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_int,
                                C.c_int, C.c_int]),
        "ref_solver_free": (None, [vp]),
        "ref_solver_run": (C.c_int, [vp, _dp, C.c_int]),
        "ref_solver_update": (C.c_int, [vp, _dp]),
        "ref_solver_solve": (C.c_int, [vp, C.c_uint, C.c_double]),
        "ref_solver_solved": (C.c_int, [vp]),
        "ref_solver_can_improve": (C.c_int, [vp]),
        "ref_solver_cost": (C.c_double, [vp]),
        "ref_solver_specific_volume": (C.c_double, [vp]),
        "ref_solver_tree_size": (C.c_int, [vp]),
        "ref_solver_dump": (C.c_int, [vp, C.c_int, _dp, _dp, _dp]),
        "ref_tree_to_yaml": (C.c_int, [vp, C.c_int, _dp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                       C.POINTER(C.c_double), C.POINTER(C.c_int)]),
        "ref_tree_from_yaml": (C.c_int, [C.c_int, C.c_int, _dp, C.c_int, C.POINTER(C.c_int), C.c_double, C.c_int,
                                         C.c_int, _dp, _dp, _dp, C.POINTER(C.c_double)]),
        "ref_solver_solution": (C.c_int, [vp, C.c_int, _dp]),
        "ref_solver_path_cost": (C.c_double, [vp]),
        "ref_solution_is_valid": (C.c_int, [vp, vp, C.c_int, _dp]),
        "ref_solution_is_valid_from": (C.c_int, [vp, vp, C.c_int, C.c_double]),
        "ref_solution_optimize": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, _dp, _dp]),
        "ref_tree_recheck": (C.c_int, [vp]),
        "ref_anytime_solve_injected": (C.c_int, [vp, C.c_uint, C.c_int, vp, _dp, _dp, C.c_double, C.c_ulonglong,
                                               C.c_ulonglong, C.c_ulonglong, C.c_double, _dp]),
        "ref_anytime_improve_begin": (C.c_int, [vp]),
        "ref_anytime_improve_update": (C.c_int, [vp, _dp]),
        "ref_anytime_state": (None, [vp, _dp]),
        "ref_anytime_set_parameters": (None, [vp, C.c_double, C.c_double]),
        "ref_informed_extend": (C.c_int, [C.c_int, vp, C.c_int, _dp, C.POINTER(C.c_int), _dp, C.c_double, C.c_int, _dp, _dp,
                                          C.c_double, C.c_double, _dp, C.POINTER(C.c_int)]),
        "ref_check_path_to_node": (C.c_int, [vp, vp, _dp, C.c_int, C.c_int, _dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "ref_rewire_only_with_path_check": (C.c_int, [vp, vp, _dp, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_int)]),
        "ref_srand": (None, [C.c_uint]),
        "ref_sample_informed": (None, [C.c_int, _dp, _dp, _dp, _dp, C.c_double, C.c_int, _dp]),
        "ref_informed_specific_volume": (C.c_double, [C.c_int, _dp, _dp, _dp, _dp, C.c_double]),
        "ref_metrics_cost": (C.c_double, [C.c_int, _dp, _dp]),
    }
    if dropin:
        sig["dropin_solver_new"] = (vp, [C.c_int, C.c_int, vp, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double,
                                         C.c_double, C.c_int, C.c_int, C.c_int, C.c_int])
        sig["dropin_checker_new"] = (vp, [C.c_int, vp, C.c_double])
        sig["dropin_checker_device_calls"] = (C.c_ulonglong, [vp])
        sig["dropin_uses_cuda_nn"] = (C.c_int, [vp])
        sig["dropin_set_nn_factory"] = (None, [C.c_int])
        sig["dropin_nn_factory_calls"] = (C.c_ulonglong, [])
        sig["dropin_device_calls_of"] = (C.c_ulonglong, [vp, vp])
        sig["dropin_solution_is_valid_batched"] = (C.c_int, [vp, vp, C.c_int, _dp])
        sig["dropin_solution_is_valid_from_batched"] = (C.c_int, [vp, vp, C.c_int, C.c_double])
        sig["dropin_tree_recheck_batched"] = (C.c_int, [vp])
        sig["dropin_check_path_to_node_batched"] = (C.c_int, [vp, vp, _dp, C.c_int, C.c_int, _dp, C.POINTER(C.c_int),
                                                            C.POINTER(C.c_int)])
        sig["dropin_rewire_only_with_path_check_batched"] = (C.c_int, [vp, vp, _dp, C.c_double, C.c_int, C.c_int,
                                                                     C.POINTER(C.c_int)])
        sig["dropin_solution_optimize_speculative"] = (C.c_int, [vp, C.c_int, C.c_int, C.c_int, _dp, _dp])
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


class RefNN:
    """graph::core::Vector ('vector') or graph::core::KdTree ('kdtree') -- the reference's own classes.
    Same method surface as oracle_py.OracleNN."""

    def __init__(self, kind: str, dim: int, fast: bool = False):
        self.L = lib(fast)
        self.dim = dim
        self.h = self.L.ref_nn_new(1 if kind == "kdtree" else 0, dim)

    def __del__(self):
        try:
            self.L.ref_nn_free(self.h)
        except Exception:
            pass

    def insert(self, q):
        q = _f64(q, (-1, self.dim))
        self.L.ref_nn_insert_many(self.h, _p(q, _dp), q.shape[0])

    def reinsert(self, node_id):
        self.L.ref_nn_reinsert(self.h, node_id)

    def delete(self, node_id):
        return bool(self.L.ref_nn_delete(self.h, node_id))

    def restore(self, node_id):
        return bool(self.L.ref_nn_restore(self.h, node_id))

    def find(self, node_id):
        return bool(self.L.ref_nn_find(self.h, node_id))

    def clear(self):
        return bool(self.L.ref_nn_clear(self.h))

    def size(self):
        return self.L.ref_nn_size(self.h)

    def set_deleted_threshold(self, t):
        self.L.ref_nn_set_deleted_threshold(self.h, t)

    def get_nodes(self, cap=1 << 22):
        ids = np.empty(cap, dtype=np.int32)
        n = self.L.ref_nn_get_nodes(self.h, _p(ids, _ip), cap)
        return ids[:n].copy()

    def nearest(self, q):
        q = _f64(q, (-1, self.dim))
        ids = np.empty(q.shape[0], dtype=np.int32)
        d = np.empty(q.shape[0], dtype=np.float64)
        self.L.ref_nn_nearest(self.h, _p(q, _dp), q.shape[0], _p(ids, _ip), _p(d, _dp))
        return ids, d

    def near(self, q, radius, cap=256):
        q = _f64(q, (-1, self.dim))
        nq = q.shape[0]
        r = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, dtype=np.float64), (nq,)))
        ids = np.full((nq, cap), -1, dtype=np.int32)
        d = np.full((nq, cap), np.inf, dtype=np.float64)
        cnt = np.zeros(nq, dtype=np.int32)
        self.L.ref_nn_near(self.h, _p(q, _dp), nq, _p(r, _dp), cap, _p(ids, _ip), _p(d, _dp), _p(cnt, _ip))
        return ids, d, cnt

    def knn(self, q, k, cap):
        q = _f64(q, (-1, self.dim))
        nq = q.shape[0]
        ids = np.full((nq, cap), -1, dtype=np.int32)
        d = np.full((nq, cap), np.inf, dtype=np.float64)
        cnt = np.zeros(nq, dtype=np.int32)
        self.L.ref_nn_knn(self.h, _p(q, _dp), nq, int(k), cap, _p(ids, _ip), _p(d, _dp), _p(cnt, _ip))
        return ids, d, cnt


class RefChecker:
    """The reference's CollisionCheckerBase loops over either its own Cube3dCollisionChecker::check or
    a per-state callback into an oracle world (for the synthetic worlds the reference does not ship)."""

    def __init__(self, h, dim, L, min_distance, keep=()):
        self.h, self.dim, self.L, self.min_distance, self._keep = h, dim, L, min_distance, keep

    @classmethod
    def cube(cls, dim, threshold=1.0, min_distance=0.01, fast=False):
        L = lib(fast)
        return cls(L.ref_checker_cube(dim, threshold, min_distance), dim, L, min_distance)

    @classmethod
    def over_oracle_world(cls, world, min_distance=0.01, fast=False):
        """world: oracle_py.OracleWorld; its orc_check_states answers the single-state checks."""
        L = lib(fast)
        fn = C.cast(world.L.orc_check_states, C.c_void_p)
        return cls(L.ref_checker_callback(world.dim, fn, world.h, min_distance), world.dim, L, min_distance, keep=(world,))

    def __del__(self):
        try:
            self.L.ref_checker_free(self.h)
        except Exception:
            pass

    def states(self):
        return int(self.L.ref_checker_states(self.h))

    def check(self, q):
        q = _f64(q, (-1, self.dim))
        ok = np.empty(q.shape[0], np.uint8)
        self.L.ref_check_states(self.h, _p(q, _dp), q.shape[0], _p(ok, _u8p))
        return ok

    def check_connection(self, c1, c2, mode=0):
        c1, c2 = _f64(c1, (-1, self.dim)), _f64(c2, (-1, self.dim))
        n = c1.shape[0]
        ok = np.empty(n, np.uint8)
        if mode == 1:
            conf = np.full((n, self.dim), np.nan)
            self.L.ref_check_edges(self.h, _p(c1, _dp), _p(c2, _dp), n, 1, _p(ok, _u8p), _p(conf, _dp))
            return ok, conf
        self.L.ref_check_edges(self.h, _p(c1, _dp), _p(c2, _dp), n, 0, _p(ok, _u8p), None)
        return ok

    def check_conn_from_conf(self, parent, child, conf):
        p, c, f = _f64(parent, (-1, self.dim)), _f64(child, (-1, self.dim)), _f64(conf, (-1, self.dim))
        res = np.empty(p.shape[0], np.uint8)
        self.L.ref_check_from_conf(self.h, _p(p, _dp), _p(c, _dp), _p(f, _dp), p.shape[0], _p(res, _u8p))
        return res


class RefSolver:
    """graph::core::{RRT, RRTStar, BiRRT, AnytimeRRT} with EuclideanMetrics and a Uniform/Informed sampler."""
    RRT, RRTSTAR, BIRRT, ANYTIME = 0, 1, 2, 3

    def __init__(self, kind, scenario, checker: RefChecker, start=None, goal=None, use_kdtree=True, extend=False,
                 informed=True, fast=False):
        self.L = lib(fast)
        self.checker = checker
        self.dim = scenario.dim
        s = _f64(scenario.start if start is None else start)
        g = _f64(scenario.goal if goal is None else goal)
        lb, ub = _f64(scenario.lb), _f64(scenario.ub)
        self.h = self.L.ref_solver_new(kind, self.dim, checker.h, _p(lb, _dp), _p(ub, _dp), _p(s, _dp), _p(g, _dp),
                                       scenario.max_distance, scenario.rewire_factor, scenario.utopia_tolerance,
                                       1 if use_kdtree else 0, 1 if extend else 0, 1 if informed else 0)

    def __del__(self):
        try:
            self.L.ref_solver_free(self.h)
        except Exception:
            pass

    def run(self, samples):
        s = _f64(samples, (-1, self.dim))
        return self.L.ref_solver_run(self.h, _p(s, _dp), s.shape[0])

    def update(self, q):
        q = _f64(q)
        return bool(self.L.ref_solver_update(self.h, _p(q, _dp)))

    def solve(self, max_iter, max_time=float("inf")):
        return bool(self.L.ref_solver_solve(self.h, max_iter, max_time))

    @property
    def solved(self):
        return bool(self.L.ref_solver_solved(self.h))

    @property
    def can_improve(self):
        return bool(self.L.ref_solver_can_improve(self.h))

    @property
    def cost(self):
        return self.L.ref_solver_cost(self.h)

    @property
    def specific_volume(self):
        return self.L.ref_solver_specific_volume(self.h)

    @property
    def tree_size(self):
        return self.L.ref_solver_tree_size(self.h)

    def dump(self):
        """conf[n][dim], parent_conf[n][dim] (NaN for the root), pcost[n] in getNodes() order."""
        n = self.L.ref_solver_dump(self.h, 0, None, None, None)
        conf = np.empty((n, self.dim))
        par = np.empty((n, self.dim))
        pc = np.empty(n)
        self.L.ref_solver_dump(self.h, n, _p(conf, _dp), _p(par, _dp), _p(pc, _dp))
        return conf, par, pc

    def tree_to_yaml(self):
        """Tree::toYAML() of the start tree as arrays: nodes[n][dim], connections[m][2], max_distance, use_kdtree."""
        n = self.tree_size + 8
        conf = np.empty((n, self.dim))
        conn = np.empty((n, 2), np.int32)
        m, md, uk = C.c_int(0), C.c_double(0), C.c_int(0)
        nn = self.L.ref_tree_to_yaml(self.h, n, _p(conf, _dp), n, conn.ctypes.data_as(C.POINTER(C.c_int)), C.byref(m),
                                     C.byref(md), C.byref(uk))
        assert 0 <= nn <= n and m.value <= n
        return conf[:nn].copy(), conn[:m.value].astype(np.int64), md.value, bool(uk.value)

    def solution(self):
        wp = np.empty((1 << 14, self.dim))
        n = self.L.ref_solver_solution(self.h, wp.shape[0], _p(wp, _dp))
        return wp[:n].copy()

    # ---- AnytimeRRT step by step (kind ANYTIME only) -------------------------------------------------------------
    def improve_begin(self):
        r = self.L.ref_anytime_improve_begin(self.h)
        assert r >= 0, "not an AnytimeRRT"
        return bool(r)

    def improve_update(self, q):
        q = _f64(q)
        return bool(self.L.ref_anytime_improve_update(self.h, _p(q, _dp)))

    def anytime_state(self):
        out = np.zeros(8)
        self.L.ref_anytime_state(self.h, _p(out, _dp))
        keys = ["bias", "cost2beat", "path_cost", "cost", "can_improve", "solved", "round_tree_nodes"]
        return dict(zip(keys, out.tolist()))

    def anytime_solve_injected(self, max_iter, sampler_fn_ptr, lb, ub, cost0, seed, p, P, utopia_tolerance,
                               improve_in_solve=True):
        """AnytimeRRT::solve in one C++ call (no Python in the loop): samples from the C function sampler_fn_ptr
        (orc_sample_informed of the oracle library) with Philox counter k * P + p.  Returns (solved, rounds, updates,
        improved rounds)."""
        lb, ub = _f64(lb), _f64(ub)
        out = np.zeros(4)
        r = self.L.ref_anytime_solve_injected(self.h, max_iter, 1 if improve_in_solve else 0, sampler_fn_ptr, _p(lb, _dp),
                                              _p(ub, _dp), float(cost0), seed, p, P, utopia_tolerance, _p(out, _dp))
        assert r >= 0, "not an AnytimeRRT"
        return bool(r), int(out[0]), int(out[1]), int(out[2])

    def set_anytime_parameters(self, delta=0.1, cost_impr=0.1):
        self.L.ref_anytime_set_parameters(self.h, delta, cost_impr)

    @property
    def path_cost(self):
        return self.L.ref_solver_path_cost(self.h)

    def solution_is_valid(self, other_checker=None):
        """Path::isValid with the path's own checker or another RefChecker; returns (valid, connection costs)."""
        costs = np.empty(1 << 14)
        r = self.L.ref_solution_is_valid(self.h, other_checker.h if other_checker is not None else None, costs.shape[0],
                                         _p(costs, _dp))
        n = max(len(self.solution()) - 1, 0)
        return r, costs[:n].copy()

    def solution_is_valid_from(self, conn_idx, t, other_checker=None):
        return self.L.ref_solution_is_valid_from(self.h, other_checker.h if other_checker is not None else None,
                                                 conn_idx, t)

    def optimize_solution(self, mode, max_iter):
        """PathLocalOptimizer on a clone of the solution: mode 0 warp, 1 simplify, 2 solve; (waypoints, cost)."""
        wp = np.empty((1 << 14, self.dim))
        cost = C.c_double(0.0)
        n = self.L.ref_solution_optimize(self.h, mode, max_iter, wp.shape[0], _p(wp, _dp), C.byref(cost))
        return wp[:n].copy(), cost.value

    def recheck_tree(self):
        return bool(self.L.ref_tree_recheck(self.h))

    def check_path_to_node(self, conf, other_checker=None, keep_flags=False, _fn=None):
        """Tree::checkPathToNode (tree.cpp:335-373) towards the tree node at `conf`: (ok, costs of the connections
        root -> node afterwards, number of connections the call checked)."""
        conf = _f64(conf)
        costs = np.empty(1 << 14)
        nc, npth = C.c_int(0), C.c_int(0)
        fn = _fn or self.L.ref_check_path_to_node
        r = fn(self.h, other_checker.h if other_checker is not None else None, _p(conf, _dp), 1 if keep_flags else 0,
               costs.shape[0], _p(costs, _dp), C.byref(nc), C.byref(npth))
        return r, costs[:npth.value].copy(), nc.value


    def rewire_only_with_path_check(self, conf, r_rewire, other_checker=None, what_rewire=0, keep_flags=False, _fn=None):
        """Tree::rewireOnlyWithPathCheck (tree.cpp:522-679) around the tree node at `conf`: (improved, connections the
        call checked)."""
        conf = _f64(conf)
        nc = C.c_int(0)
        fn = _fn or self.L.ref_rewire_only_with_path_check
        r = fn(self.h, other_checker.h if other_checker is not None else None, _p(conf, _dp), float(r_rewire),
               int(what_rewire), 1 if keep_flags else 0, C.byref(nc))
        return r, nc.value


def cuda_ref_checker(cuda_world, min_distance):
    """A RefChecker whose CollisionCheckerBase is graph::core::CudaCollisionChecker over a CUDA world."""
    L = lib(dropin=True)
    h = L.dropin_checker_new(cuda_world.dim, cuda_world.handle, min_distance)
    if not h:
        raise RuntimeError("dropin_checker_new failed")
    return RefChecker(h, cuda_world.dim, L, min_distance, keep=(cuda_world,))


class DropinSolver(RefSolver):
    """The reference's solver classes with this repository's CUDA plugin classes plugged in (oracle/ref_dropin.cpp):
    CudaCollisionChecker for every solver, CudaNearestNeighbors for the start tree when cuda_nn is set."""

    def __init__(self, kind, scenario, cuda_world, start=None, goal=None, extend=False, informed=True, cuda_nn=True,
                 device=0):
        self.L = lib(dropin=True)
        self.checker = cuda_world  # keep alive
        self.dim = scenario.dim
        s = _f64(scenario.start if start is None else start)
        g = _f64(scenario.goal if goal is None else goal)
        lb, ub = _f64(scenario.lb), _f64(scenario.ub)
        self.h = self.L.dropin_solver_new(kind, self.dim, cuda_world.handle, _p(lb, _dp), _p(ub, _dp), _p(s, _dp),
                                          _p(g, _dp), scenario.max_distance, scenario.min_distance,
                                          scenario.rewire_factor, scenario.utopia_tolerance, 1 if extend else 0,
                                          1 if informed else 0, 1 if cuda_nn else 0, device)
        if not self.h:
            raise RuntimeError("dropin_solver_new failed (no CUDA device?)")

    @property
    def device_calls(self):
        return int(self.L.dropin_checker_device_calls(self.h))

    @property
    def uses_cuda_nn(self):
        return bool(self.L.dropin_uses_cuda_nn(self.h))

    # SURVEY 8(f) N1-N3 through graph_core_b200/host/cuda_batched_helpers.h (edge checks gathered into device batches)
    def device_calls_of(self, other_checker=None):
        return int(self.L.dropin_device_calls_of(self.h, other_checker.h if other_checker is not None else None))

    def solution_is_valid_batched(self, other_checker=None):
        costs = np.empty(1 << 14)
        r = self.L.dropin_solution_is_valid_batched(self.h, other_checker.h if other_checker is not None else None,
                                                    costs.shape[0], _p(costs, _dp))
        n = max(len(self.solution()) - 1, 0)
        return r, costs[:n].copy()

    def solution_is_valid_from_batched(self, conn_idx, t, other_checker=None):
        return self.L.dropin_solution_is_valid_from_batched(self.h, other_checker.h if other_checker is not None else None,
                                                            conn_idx, t)

    def recheck_tree_batched(self):
        return bool(self.L.dropin_tree_recheck_batched(self.h))

    def check_path_to_node_batched(self, conf, other_checker=None, keep_flags=False):
        return self.check_path_to_node(conf, other_checker, keep_flags, _fn=self.L.dropin_check_path_to_node_batched)

    def rewire_only_with_path_check_batched(self, conf, r_rewire, other_checker=None, what_rewire=0, keep_flags=False):
        return self.rewire_only_with_path_check(conf, r_rewire, other_checker, what_rewire, keep_flags,
                                                _fn=self.L.dropin_rewire_only_with_path_check_batched)

    def optimize_solution_speculative(self, mode, max_iter):
        wp = np.empty((1 << 14, self.dim))
        cost = C.c_double(0.0)
        n = self.L.dropin_solution_optimize_speculative(self.h, mode, max_iter, wp.shape[0], _p(wp, _dp), C.byref(cost))
        return wp[:n].copy(), cost.value


def tree_as_edges(conf, parent_conf, pcost):
    """Canonical, order-free form of a tree dump: {child conf bytes: (parent conf bytes | None, cost)}."""
    out = {}
    for i in range(conf.shape[0]):
        p = None if np.isnan(parent_conf[i]).any() else parent_conf[i].tobytes()
        out[conf[i].tobytes()] = (p, float(pcost[i]))
    return out


def oracle_tree_as_edges(parents, pcost, conf, in_tree=None):
    """Same canonical form from oracle_py.OracleSolver.dump() / LockstepRRTStar.dump() (ids = creation order);
    nodes with parent -1 other than the root (id 0) are not in the tree yet (e.g. an unconnected goal)."""
    out = {}
    for i in range(len(parents)):
        if parents[i] < 0 and i != 0:
            continue
        p = None if parents[i] < 0 else conf[parents[i]].tobytes()
        out[conf[i].tobytes()] = (p, float(pcost[i]) if parents[i] >= 0 else 0.0)
    return out


def ref_tree_from_yaml(conf, connections, max_distance=None, use_kdtree=True, fast=False):
    """Tree::fromYAML on a document given as arrays; returns (conf, parent_conf, pcost, max_distance) of the loaded
    tree in getNodes() order, or None when the reference refuses the document."""
    L = lib(fast)
    conf = np.ascontiguousarray(conf, np.float64)
    n, dim = conf.shape
    conn = np.ascontiguousarray(connections, np.int32).reshape(-1, 2)
    oc = np.empty((n + 8, dim))
    op = np.empty((n + 8, dim))
    pc = np.empty(n + 8)
    md = C.c_double(0)
    r = L.ref_tree_from_yaml(dim, n, _p(conf, _dp), conn.shape[0], conn.ctypes.data_as(C.POINTER(C.c_int)),
                             float("nan") if max_distance is None else float(max_distance), int(use_kdtree), n + 8,
                             _p(oc, _dp), _p(op, _dp), _p(pc, _dp), C.byref(md))
    if r < 0:
        return None
    return oc[:r].copy(), op[:r].copy(), pc[:r].copy(), md.value


def informed_extend(checker: RefChecker, conf, parent, pcost, max_distance, q, goal, cost2beat, bias, use_kdtree=True):
    """The reference's Tree::informedExtend (tree.cpp:235-302) on a tree given as arrays; (ok, new_conf, parent index)."""
    L = checker.L
    conf = _f64(conf)
    n, dim = conf.shape
    parent = np.ascontiguousarray(parent, dtype=np.int32)
    pcost = _f64(pcost)
    q, goal = _f64(q), _f64(goal)
    new = np.full(dim, np.nan)
    pi = C.c_int(-1)
    ok = L.ref_informed_extend(dim, checker.h, n, _p(conf, _dp), parent.ctypes.data_as(C.POINTER(C.c_int)), _p(pcost, _dp),
                               max_distance, 1 if use_kdtree else 0, _p(q, _dp), _p(goal, _dp), cost2beat, bias,
                               _p(new, _dp), C.byref(pi))
    return bool(ok), new, pi.value


def anytime_solve(solver, sample_fn, max_iter, improve_in_solve=True, failed_iter=3, utopia_tolerance=None,
                  best_utopia=None):
    """AnytimeRRT::solve (anytime_rrt.cpp:73-236) driven through the reference's own public methods with injected
    samples: RRT::update(configuration) for the first phase (TreeSolver::solve's loop, tree_solver.cpp:117-127),
    improve(max_iter = 0) + improveUpdate(configuration) for the rounds.  sample_fn(k, cost) -> the k-th sample of this
    problem from the informed set of that cost.  improve_in_solve: enter the improve loop although RRT::update left
    can_improve_ false (rrt.cpp:150) -- see oracle/gc_oracle.cpp AnytimeSolver.  Returns (solved, log rows) with the
    columns of oracle_py.ANYTIME_LOG_COLUMNS (tree_nodes / set_up / bias / cost2beat read from the solver)."""
    log = []
    k = 0

    def row(phase, iters, ok, setup, tree_nodes):
        st = solver.anytime_state()
        # cost2beat_ is uninitialised in the reference until the first improve()
        log.append([phase, iters, 1.0 if ok else 0.0, st["path_cost"], tree_nodes, st["bias"],
                    st["cost2beat"] if phase == 1 else 0.0, 1.0 if setup else 0.0])

    tol_utopia = utopia_tolerance * best_utopia
    if solver.solved and solver.cost <= tol_utopia:
        return True, np.array(log).reshape(-1, 8)
    n_failed = 0
    while (not solver.solved) and n_failed < failed_iter:
        ok = False
        it = 0
        while it < max_iter:
            q = sample_fn(k, None)
            k += 1
            it += 1
            if solver.update(q):
                ok = True
                break
        row(0, it, ok, True, solver.tree_size)
        if not ok:
            n_failed += 1
    if not solver.solved:
        return False, np.array(log).reshape(-1, 8)
    if solver.cost <= tol_utopia:
        return True, np.array(log).reshape(-1, 8)
    n_failed = 0
    while (improve_in_solve or solver.can_improve) and n_failed < failed_iter:
        if not solver.improve_begin():
            row(1, 0, False, False, 0.0)
            improved = False
        else:
            c2b = solver.anytime_state()["cost2beat"]
            improved = False
            it = 0
            while it < max_iter:
                q = sample_fn(k, c2b)
                k += 1
                it += 1
                if solver.improve_update(q):
                    improved = True
                    break
            st = solver.anytime_state()
            row(1, it, improved, True, solver.tree_size if improved else st["round_tree_nodes"])
        n_failed = 0 if improved else n_failed + 1
        if solver.cost <= tol_utopia:
            break
    return solver.solved, np.array(log).reshape(-1, 8)
