"""ctypes binding of include/gcb200.h plus thin numpy-friendly wrappers.

This is the Python mirror of the reference-facing operator interface:
  NodeStore        <-> graph::core::NearestNeighbors   (nearest_neighbors.h:41-195)
  CollisionWorld   <-> graph::core::CollisionCheckerBase (collision_checker_base.h:118-319)
There is no CPU fallback: loading fails loudly when the CUDA library is missing, and every
create() fails when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libgcb200.so"

GC_OK = 0
GC_W_TRUNCATED = 1
GC_W_NOOP = 2
GC_E_INVALID = -1
GC_E_CUDA = -2
GC_E_CAPACITY = -3
GC_E_UNSUPPORTED = -4
GC_NO_NODE = 0xFFFFFFFF

EDGE_BISECT = 0
EDGE_LINEAR = 1
EDGE_FROM_CONF = 2

_c_double_p = C.POINTER(C.c_double)
_c_float_p = C.POINTER(C.c_float)
_c_u32_p = C.POINTER(C.c_uint32)
_c_i32_p = C.POINTER(C.c_int32)
_c_u8_p = C.POINTER(C.c_uint8)
_c_i64_p = C.POINTER(C.c_int64)
_c_u64_p = C.POINTER(C.c_uint64)

# name -> (restype, argtypes); every symbol declared in include/gcb200.h
PROTOTYPES = {
    "gc_version": (C.c_int, []),
    "gc_last_error": (C.c_char_p, []),
    "gc_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "gc_device_info": (C.c_int, [C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gc_store_create": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_void_p)]),
    "gc_store_destroy": (C.c_int, [C.c_void_p]),
    "gc_store_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gc_store_append": (C.c_int, [C.c_void_p, C.c_uint32, _c_double_p, C.c_uint32, _c_u32_p]),
    "gc_store_append_multi": (C.c_int, [C.c_void_p, _c_u32_p, _c_double_p, C.c_uint32, _c_u32_p]),
    "gc_store_tombstone": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "gc_store_restore": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "gc_store_clear": (C.c_int, [C.c_void_p, C.c_uint32]),
    "gc_store_size": (C.c_int, [C.c_void_p, C.c_uint32, _c_u32_p, _c_u32_p]),
    "gc_store_read": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, _c_double_p]),
    "gc_store_dim": (C.c_int, [C.c_void_p]),
    "gc_nn1": (C.c_int, [C.c_void_p, _c_double_p, _c_u32_p, C.c_uint32, _c_u32_p, _c_double_p]),
    "gc_radius": (C.c_int, [C.c_void_p, _c_double_p, _c_u32_p, _c_double_p, C.c_uint32, C.c_uint32, _c_u32_p,
                            _c_double_p, _c_u32_p]),
    "gc_knn": (C.c_int, [C.c_void_p, _c_double_p, _c_u32_p, C.c_uint32, C.c_uint64, C.c_uint32, _c_u32_p,
                         _c_double_p, _c_u32_p]),
    "gc_nn1_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]),
    "gc_radius_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p,
                                C.c_void_p, C.c_void_p]),
    "gc_knn_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_void_p,
                             C.c_void_p, C.c_void_p]),
    "gc_world_create_cube": (C.c_int, [C.c_int, C.c_double, C.c_int, C.POINTER(C.c_void_p)]),
    "gc_world_create_boxes": (C.c_int, [C.c_int, C.c_uint32, _c_double_p, _c_double_p, C.c_int,
                                        C.POINTER(C.c_void_p)]),
    "gc_world_create_cspheres": (C.c_int, [C.c_int, C.c_uint32, _c_float_p, _c_float_p, C.c_int,
                                           C.POINTER(C.c_void_p)]),
    "gc_world_create_sphere_arm": (C.c_int, [C.c_int, _c_float_p, _c_float_p, C.c_uint32, _c_i32_p, _c_float_p,
                                             C.c_uint32, _c_float_p, C.c_int, C.POINTER(C.c_void_p)]),
    "gc_world_retain": (C.c_int, [C.c_void_p]),
    "gc_world_clone": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "gc_world_destroy": (C.c_int, [C.c_void_p]),
    "gc_world_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gc_world_dim": (C.c_int, [C.c_void_p]),
    "gc_world_counters": (C.c_int, [C.c_void_p, _c_u64_p]),
    "gc_world_reset_counters": (C.c_int, [C.c_void_p]),
    "gc_world_enable_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "gc_check_states": (C.c_int, [C.c_void_p, _c_double_p, C.c_uint32, _c_u8_p]),
    "gc_check_edges": (C.c_int, [C.c_void_p, _c_double_p, _c_double_p, C.c_uint32, C.c_double, C.c_int, _c_u8_p,
                                 _c_i64_p]),
    "gc_check_edges_from_conf": (C.c_int, [C.c_void_p, _c_double_p, _c_double_p, _c_double_p, C.c_uint32, C.c_double,
                                           _c_u8_p]),
    "gc_check_states_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "gc_check_edges_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_double, C.c_int, C.c_void_p,
                                     C.c_void_p]),
    "gc_extend_batch": (C.c_int, [C.c_void_p, C.c_void_p, _c_double_p, _c_u32_p, C.c_uint32, C.c_double, C.c_double,
                                  C.c_double, _c_double_p, C.c_uint32, _c_u32_p, _c_double_p, _c_double_p, _c_u8_p,
                                  _c_u32_p, _c_double_p, _c_u32_p]),
    "gc_sample_uniform": (C.c_int, [C.c_int, C.c_int, _c_double_p, _c_double_p, C.c_uint64, C.c_uint64, C.c_uint32,
                                    _c_double_p]),
    "gc_sample_uniform_dev": (C.c_int, [C.c_int, C.c_int, _c_double_p, _c_double_p, C.c_uint64, C.c_uint64, C.c_uint32,
                                        C.c_void_p, C.c_void_p]),
    "gc_sample_informed": (C.c_int, [C.c_int, C.c_int, _c_double_p, _c_double_p, _c_double_p, _c_double_p, _c_double_p,
                                     C.c_uint64, C.c_uint64, C.c_uint32, _c_double_p, _c_u32_p]),
    "gc_sample_informed_dev": (C.c_int, [C.c_int, C.c_int, _c_double_p, _c_double_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
    "gc_extend_batch_dsamples": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _c_u32_p, _c_u32_p, C.c_uint32,
                                           C.c_double, C.c_double, C.c_double, _c_double_p, C.c_uint32, _c_u32_p,
                                           _c_double_p, _c_double_p, _c_u8_p, _c_u32_p, _c_double_p, _c_u32_p]),
    "gc_merge_shard_lists_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p,
                                           C.c_void_p]),
    "gc_merge_shard_lists_p2p_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                               C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                               C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    # device-resident lock-step trees (driven from C++ by host/lockstep.cpp; bound here for completeness)
    "gc_forest_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _c_double_p, _c_double_p, C.c_uint32, _c_double_p,
                                   _c_double_p, _c_u32_p, _c_i32_p, _c_double_p, _c_i32_p, _c_u8_p, _c_double_p,
                                   C.POINTER(C.c_void_p)]),
    "gc_forest_destroy": (C.c_int, [C.c_void_p]),
    "gc_forest_step": (C.c_int, [C.c_void_p, _c_double_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "gc_forest_read_costs_async": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gc_forest_sync": (C.c_int, [C.c_void_p]),
    "gc_forest_stats": (C.c_int, [C.c_void_p, _c_u64_p]),
    "gc_forest_info": (C.c_int, [C.c_void_p, C.c_uint32, _c_i32_p, _c_i32_p, _c_double_p, _c_double_p, _c_u32_p,
                                 _c_u32_p]),
    "gc_forest_dump": (C.c_int, [C.c_void_p, C.c_uint32, _c_i32_p, _c_double_p, _c_double_p]),
    "gc_forest_solution": (C.c_int, [C.c_void_p, C.c_uint32, _c_i32_p, C.c_uint32]),
    # device-resident lock-step BiRRT / RRT (csrc/birrt.cu)
    "gc_birrt_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, _c_double_p, _c_double_p, C.POINTER(C.c_void_p)]),
    "gc_birrt_destroy": (C.c_int, [C.c_void_p]),
    "gc_birrt_step": (C.c_int, [C.c_void_p, _c_double_p, C.c_void_p]),
    "gc_birrt_sync": (C.c_int, [C.c_void_p]),
    "gc_birrt_stats": (C.c_int, [C.c_void_p, _c_u64_p]),
    "gc_birrt_info": (C.c_int, [C.c_void_p, C.c_uint32, _c_i32_p, _c_u32_p, _c_double_p, _c_u32_p, _c_u32_p]),
    "gc_birrt_read_solved": (C.c_int, [C.c_void_p, _c_u8_p, _c_double_p]),
    "gc_birrt_dump": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int, _c_i32_p, _c_double_p, _c_double_p]),
    "gc_birrt_solution": (C.c_int, [C.c_void_p, C.c_uint32, _c_double_p, _c_double_p, C.c_uint32]),
    "gc_informed_extend_batch": (C.c_int, [C.c_void_p, C.c_void_p, _c_double_p, _c_u32_p, _c_double_p, _c_double_p,
                                           _c_double_p, C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_double,
                                           C.c_void_p, C.c_void_p, _c_u8_p, _c_u32_p, _c_double_p, _c_double_p,
                                           _c_double_p, _c_u64_p]),
    # lock-step AnytimeRRT (csrc/anytime.cu)
    "gc_anytime_create": (C.c_int, [C.c_void_p, C.c_void_p, _c_double_p, _c_double_p, C.c_uint32, _c_double_p,
                                    _c_double_p, _c_double_p, C.POINTER(C.c_void_p)]),
    "gc_anytime_destroy": (C.c_int, [C.c_void_p]),
    "gc_anytime_set_streams": (C.c_int, [C.c_void_p, C.c_uint64, _c_u64_p]),
    "gc_anytime_step": (C.c_int, [C.c_void_p]),
    "gc_anytime_run": (C.c_int, [C.c_void_p, C.c_uint64]),
    "gc_anytime_stats": (C.c_int, [C.c_void_p, _c_u64_p]),
    "gc_anytime_timing": (C.c_int, [C.c_void_p, _c_double_p]),
    "gc_anytime_info": (C.c_int, [C.c_void_p, C.c_uint32, _c_i32_p, _c_i32_p, _c_double_p, _c_double_p, _c_double_p,
                                  _c_u64_p, _c_u32_p, _c_u32_p]),
    "gc_anytime_read_solved": (C.c_int, [C.c_void_p, _c_u8_p, _c_double_p]),
    "gc_anytime_log": (C.c_int, [C.c_void_p, C.c_uint32, _c_double_p, C.c_uint32]),
    "gc_anytime_solution": (C.c_int, [C.c_void_p, C.c_uint32, _c_double_p, _c_double_p, C.c_uint32]),
    "gc_anytime_dump": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int, _c_i32_p, _c_double_p, _c_double_p, C.c_uint32]),
    # node-sharded tree behind the C ABI (csrc/sharded.cu)
    "gc_comm_init": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]),
    "gc_comm_size": (C.c_int, [C.c_void_p]),
    "gc_comm_destroy": (C.c_int, [C.c_void_p]),
    "gc_store_create_sharded": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.POINTER(C.c_void_p)]),
    "gc_sharded_store_destroy": (C.c_int, [C.c_void_p]),
    "gc_sharded_store_size": (C.c_int, [C.c_void_p, _c_u64_p, _c_u32_p]),
    "gc_sharded_store_append": (C.c_int, [C.c_void_p, _c_double_p, C.c_uint32, _c_u64_p]),
    "gc_sharded_nn1": (C.c_int, [C.c_void_p, _c_double_p, C.c_uint32, _c_u32_p, _c_double_p]),
    "gc_sharded_knn": (C.c_int, [C.c_void_p, _c_double_p, C.c_uint32, C.c_uint64, C.c_uint32, _c_u32_p, _c_double_p,
                                 _c_u32_p]),
    "gc_sharded_radius": (C.c_int, [C.c_void_p, _c_double_p, _c_double_p, C.c_uint32, C.c_uint32, _c_u32_p, _c_double_p,
                                    _c_u32_p]),
    "gc_launch_count": (C.c_int, [_c_u64_p]),
    "gc_measure_fp32_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "gc_flush_l2": (C.c_int, [C.c_int, C.c_size_t]),
}

_lib = None


class GcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("gcb200 error %d: %s" % (code, msg))
        self.code = code


def load(path: Path | None = None):
    """Load libgcb200.so and attach prototypes.  Raises when the library is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise RuntimeError(
            "%s is missing: build it with `python -m graph_core_b200._build` (nvcc, sm_100a). "
            "graph_core_b200 has no CPU fallback." % p)
    lib = C.CDLL(str(p))
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError when the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _check(rc: int) -> int:
    if rc < 0:
        raise GcError(rc, load().gc_last_error().decode("utf-8", "replace"))
    return rc


def _f64(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _ptr(a: np.ndarray, typ):
    return a.ctypes.data_as(typ)


def stream_handle(cuda_stream) -> int:
    """cudaStream_t for the C ABI: None = the handle's private stream (NULL), 0 = CUDA's legacy default
    stream (what torch.cuda.current_stream().cuda_stream is unless a stream context is active; the C ABI
    takes it as cudaStreamLegacy = 0x1 because NULL already means "private"), anything else as is."""
    if cuda_stream is None:
        return 0
    return 1 if int(cuda_stream) == 0 else int(cuda_stream)


def device_count() -> int:
    n = C.c_int(0)
    rc = load().gc_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def device_info(device: int = 0):
    name = C.create_string_buffer(256)
    sms, cc = C.c_int(0), C.c_int(0)
    _check(load().gc_device_info(device, name, 256, C.byref(sms), C.byref(cc)))
    return {"name": name.value.decode(), "sms": sms.value, "cc": cc.value}


class NodeStore:
    """Device-resident SoA node store with the NearestNeighbors query set.

    Slots are insertion order per segment (Vector semantics, vector.cpp:38-43); one segment is one
    tree.  Methods mirror NearestNeighbors: insert -> append, nearestNeighbor -> nearest,
    near -> near, kNearestNeighbors -> knn, deleteNode/restoreNode -> tombstone/restore.
    """

    def __init__(self, dim: int, capacity: int, segments: int = 1, device: int = 0):
        self._lib = load()
        self.dim = int(dim)
        self.segments = int(segments)
        self.device = int(device)
        h = C.c_void_p()
        _check(self._lib.gc_store_create(self.dim, int(capacity), self.segments, self.device, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gc_store_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def set_stream(self, cuda_stream: int | None):
        _check(self._lib.gc_store_set_stream(self._h, C.c_void_p(stream_handle(cuda_stream))))

    def append(self, q, segment: int = 0) -> int:
        q = _f64(q, (-1, self.dim))
        first = C.c_uint32(0)
        _check(self._lib.gc_store_append(self._h, segment, _ptr(q, _c_double_p), q.shape[0], C.byref(first)))
        return first.value

    def append_multi(self, segments, q) -> np.ndarray:
        q = _f64(q, (-1, self.dim))
        seg = np.ascontiguousarray(segments, dtype=np.uint32)
        assert seg.shape[0] == q.shape[0]
        slots = np.empty(q.shape[0], dtype=np.uint32)
        _check(self._lib.gc_store_append_multi(self._h, _ptr(seg, _c_u32_p), _ptr(q, _c_double_p), q.shape[0],
                                               _ptr(slots, _c_u32_p)))
        return slots

    def tombstone(self, slot: int, segment: int = 0) -> int:
        return _check(self._lib.gc_store_tombstone(self._h, segment, slot))

    def restore(self, slot: int, segment: int = 0) -> int:
        return _check(self._lib.gc_store_restore(self._h, segment, slot))

    def clear(self, segment: int | None = None):
        _check(self._lib.gc_store_clear(self._h, GC_NO_NODE if segment is None else segment))

    def size(self, segment: int = 0):
        used, live = C.c_uint32(0), C.c_uint32(0)
        _check(self._lib.gc_store_size(self._h, segment, C.byref(used), C.byref(live)))
        return used.value, live.value

    def read(self, first: int, n: int, segment: int = 0) -> np.ndarray:
        out = np.empty((n, self.dim), dtype=np.float64)
        _check(self._lib.gc_store_read(self._h, segment, first, n, _ptr(out, _c_double_p)))
        return out

    def _seg(self, seg, nq):
        if seg is None:
            return None, None
        seg = np.ascontiguousarray(seg, dtype=np.uint32)
        assert seg.shape[0] == nq
        return seg, _ptr(seg, _c_u32_p)

    def nearest(self, q, seg=None):
        q = _f64(q, (-1, self.dim))
        nq = q.shape[0]
        seg, segp = self._seg(seg, nq)
        idx = np.empty(nq, dtype=np.uint32)
        dist = np.empty(nq, dtype=np.float64)
        _check(self._lib.gc_nn1(self._h, _ptr(q, _c_double_p), segp, nq, _ptr(idx, _c_u32_p), _ptr(dist, _c_double_p)))
        return idx, dist

    def near(self, q, radius, cap: int = 256, seg=None, grow: bool = True):
        """Returns (idx[nq,cap], dist[nq,cap], count[nq]); rows sorted by (dist, slot)."""
        q = _f64(q, (-1, self.dim))
        nq = q.shape[0]
        radius = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, dtype=np.float64), (nq,)))
        seg, segp = self._seg(seg, nq)
        while True:
            idx = np.full((nq, cap), GC_NO_NODE, dtype=np.uint32)
            dist = np.full((nq, cap), np.inf, dtype=np.float64)
            count = np.zeros(nq, dtype=np.uint32)
            rc = _check(self._lib.gc_radius(self._h, _ptr(q, _c_double_p), segp, _ptr(radius, _c_double_p), nq, cap,
                                            _ptr(idx, _c_u32_p), _ptr(dist, _c_double_p), _ptr(count, _c_u32_p)))
            if rc == GC_W_TRUNCATED and grow:
                cap = int(2 ** np.ceil(np.log2(int(count.max()))))
                continue
            return idx, dist, count

    def knn(self, q, k: int, cap: int | None = None, seg=None):
        q = _f64(q, (-1, self.dim))
        nq = q.shape[0]
        seg, segp = self._seg(seg, nq)
        if cap is None:
            cap = max(1, min(int(k), max(self.size(s)[0] for s in range(self.segments))))
        idx = np.full((nq, cap), GC_NO_NODE, dtype=np.uint32)
        dist = np.full((nq, cap), np.inf, dtype=np.float64)
        count = np.zeros(nq, dtype=np.uint32)
        _check(self._lib.gc_knn(self._h, _ptr(q, _c_double_p), segp, nq, int(k), cap, _ptr(idx, _c_u32_p),
                                _ptr(dist, _c_double_p), _ptr(count, _c_u32_p)))
        return idx, dist, count


class ShardedStore:
    """A tree node-sharded over the GPUs of one box, driven by ONE process through the C ABI (gc_comm_init /
    gc_store_create_sharded): node i on shard i mod G, exchange fused into the merge kernel (peer loads)."""

    def __init__(self, dim: int, capacity_total: int, devices):
        self._lib = load()
        self.dim = int(dim)
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        c = C.c_void_p()
        _check(self._lib.gc_comm_init(len(devices), devs, C.byref(c)))
        self._comm = c
        h = C.c_void_p()
        rc = self._lib.gc_store_create_sharded(c, self.dim, int(capacity_total), C.byref(h))
        if rc < 0:
            self._lib.gc_comm_destroy(c)
            _check(rc)
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gc_sharded_store_destroy(self._h)
            self._lib.gc_comm_destroy(self._comm)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def append(self, q) -> int:
        q = _f64(q, (-1, self.dim))
        first = C.c_uint64(0)
        _check(self._lib.gc_sharded_store_append(self._h, _ptr(q, _c_double_p), q.shape[0], C.byref(first)))
        return first.value

    def size(self):
        n, g = C.c_uint64(0), C.c_uint32(0)
        _check(self._lib.gc_sharded_store_size(self._h, C.byref(n), C.byref(g)))
        return n.value, g.value

    def nearest(self, q):
        q = _f64(q, (-1, self.dim))
        nq = q.shape[0]
        idx = np.empty(nq, dtype=np.uint32)
        dist = np.empty(nq, dtype=np.float64)
        _check(self._lib.gc_sharded_nn1(self._h, _ptr(q, _c_double_p), nq, _ptr(idx, _c_u32_p), _ptr(dist, _c_double_p)))
        return idx, dist

    def knn(self, q, k: int, cap: int | None = None):
        q = _f64(q, (-1, self.dim))
        nq = q.shape[0]
        cap = int(k) if cap is None else int(cap)
        idx = np.full((nq, cap), GC_NO_NODE, dtype=np.uint32)
        dist = np.full((nq, cap), np.inf, dtype=np.float64)
        count = np.zeros(nq, dtype=np.uint32)
        _check(self._lib.gc_sharded_knn(self._h, _ptr(q, _c_double_p), nq, int(k), cap, _ptr(idx, _c_u32_p),
                                        _ptr(dist, _c_double_p), _ptr(count, _c_u32_p)))
        return idx, dist, count

    def near(self, q, radius, cap: int = 256, grow: bool = True):
        q = _f64(q, (-1, self.dim))
        nq = q.shape[0]
        radius = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, dtype=np.float64), (nq,)))
        while True:
            idx = np.full((nq, cap), GC_NO_NODE, dtype=np.uint32)
            dist = np.full((nq, cap), np.inf, dtype=np.float64)
            count = np.zeros(nq, dtype=np.uint32)
            rc = _check(self._lib.gc_sharded_radius(self._h, _ptr(q, _c_double_p), _ptr(radius, _c_double_p), nq, cap,
                                                    _ptr(idx, _c_u32_p), _ptr(dist, _c_double_p), _ptr(count, _c_u32_p)))
            if rc == GC_W_TRUNCATED and grow:
                cap = int(2 ** np.ceil(np.log2(int(count.max()))))
                continue
            return idx, dist, count


class CollisionWorld:
    """Device-resident collision world with CollisionCheckerBase's checking entry points."""

    def __init__(self, handle, dim, lib):
        self._h = handle
        self.dim = dim
        self._lib = lib

    # ---- constructors ---------------------------------------------------------------------------
    @classmethod
    def cube(cls, dim: int, threshold: float = 1.0, device: int = 0):
        lib = load()
        h = C.c_void_p()
        _check(lib.gc_world_create_cube(dim, threshold, device, C.byref(h)))
        return cls(h, dim, lib)

    @classmethod
    def boxes(cls, lo, hi, device: int = 0):
        lib = load()
        lo = _f64(lo)
        hi = _f64(hi)
        assert lo.shape == hi.shape and lo.ndim == 2
        h = C.c_void_p()
        _check(lib.gc_world_create_boxes(lo.shape[1], lo.shape[0], _ptr(lo, _c_double_p), _ptr(hi, _c_double_p), device,
                                         C.byref(h)))
        return cls(h, lo.shape[1], lib)

    @classmethod
    def cspheres(cls, centers, radii, device: int = 0):
        lib = load()
        c = np.ascontiguousarray(centers, dtype=np.float32)
        r = np.ascontiguousarray(radii, dtype=np.float32)
        assert c.ndim == 2 and r.shape[0] == c.shape[0]
        h = C.c_void_p()
        _check(lib.gc_world_create_cspheres(c.shape[1], c.shape[0], _ptr(c, _c_float_p), _ptr(r, _c_float_p), device,
                                            C.byref(h)))
        return cls(h, c.shape[1], lib)

    @classmethod
    def sphere_arm(cls, base_tf, joint_tf, rs_link, rs_local, obs, device: int = 0):
        lib = load()
        base = np.ascontiguousarray(base_tf, dtype=np.float32).reshape(12)
        jt = np.ascontiguousarray(joint_tf, dtype=np.float32).reshape(-1, 12)
        link = np.ascontiguousarray(rs_link, dtype=np.int32)
        loc = np.ascontiguousarray(rs_local, dtype=np.float32).reshape(-1, 4)
        ob = np.ascontiguousarray(obs, dtype=np.float32).reshape(-1, 4)
        h = C.c_void_p()
        _check(lib.gc_world_create_sphere_arm(jt.shape[0], _ptr(base, _c_float_p), _ptr(jt, _c_float_p), loc.shape[0],
                                              _ptr(link, _c_i32_p), _ptr(loc, _c_float_p), ob.shape[0],
                                              _ptr(ob, _c_float_p), device, C.byref(h)))
        return cls(h, jt.shape[0], lib)

    def clone(self):
        """CollisionCheckerBase::clone: an independent host object sharing the immutable device world."""
        h = C.c_void_p()
        _check(self._lib.gc_world_clone(self._h, C.byref(h)))
        return CollisionWorld(h, self.dim, self._lib)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gc_world_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def set_stream(self, cuda_stream: int | None):
        _check(self._lib.gc_world_set_stream(self._h, C.c_void_p(stream_handle(cuda_stream))))

    def enable_timing(self, on: bool = True):
        _check(self._lib.gc_world_enable_timing(self._h, 1 if on else 0))

    def reset_counters(self):
        _check(self._lib.gc_world_reset_counters(self._h))

    def counters(self):
        out = (C.c_uint64 * 4)()
        _check(self._lib.gc_world_counters(self._h, out))
        return {"states": out[0], "pair_tests": out[1], "launches": out[2], "kernel_ns": out[3]}

    # ---- checks ---------------------------------------------------------------------------------
    def check(self, q) -> np.ndarray:
        q = _f64(q, (-1, self.dim))
        ok = np.empty(q.shape[0], dtype=np.uint8)
        _check(self._lib.gc_check_states(self._h, _ptr(q, _c_double_p), q.shape[0], _ptr(ok, _c_u8_p)))
        return ok

    def check_connection(self, c1, c2, min_distance: float = 0.01, mode: int = EDGE_BISECT):
        c1 = _f64(c1, (-1, self.dim))
        c2 = _f64(c2, (-1, self.dim))
        assert c1.shape == c2.shape
        n = c1.shape[0]
        ok = np.empty(n, dtype=np.uint8)
        if mode == EDGE_LINEAR:
            fb = np.empty(n, dtype=np.int64)
            _check(self._lib.gc_check_edges(self._h, _ptr(c1, _c_double_p), _ptr(c2, _c_double_p), n, min_distance,
                                            mode, _ptr(ok, _c_u8_p), _ptr(fb, _c_i64_p)))
            return ok, fb
        _check(self._lib.gc_check_edges(self._h, _ptr(c1, _c_double_p), _ptr(c2, _c_double_p), n, min_distance, mode,
                                        _ptr(ok, _c_u8_p), None))
        return ok

    def check_conn_from_conf(self, parent, child, conf, min_distance: float = 0.01) -> np.ndarray:
        p = _f64(parent, (-1, self.dim))
        c = _f64(child, (-1, self.dim))
        f = _f64(conf, (-1, self.dim))
        res = np.empty(p.shape[0], dtype=np.uint8)
        _check(self._lib.gc_check_edges_from_conf(self._h, _ptr(p, _c_double_p), _ptr(c, _c_double_p),
                                                  _ptr(f, _c_double_p), p.shape[0], min_distance, _ptr(res, _c_u8_p)))
        return res


def extend_batch(store: NodeStore, world: CollisionWorld, q, seg, max_distance, min_distance, tolerance=1e-6,
                 radius=None, near_cap: int = 256):
    """Fused Tree::extend device part for n lock-step trees (gc_extend_batch)."""
    lib = load()
    q = _f64(q, (-1, store.dim))
    n = q.shape[0]
    segp = None
    if seg is not None:
        seg = np.ascontiguousarray(seg, dtype=np.uint32)
        segp = _ptr(seg, _c_u32_p)
    closest = np.empty(n, dtype=np.uint32)
    dist = np.empty(n, dtype=np.float64)
    nxt = np.empty((n, store.dim), dtype=np.float64)
    ok = np.empty(n, dtype=np.uint8)
    if radius is None:
        rc = lib.gc_extend_batch(store.handle, world.handle, _ptr(q, _c_double_p), segp, n, max_distance, min_distance,
                                 tolerance, None, 0, _ptr(closest, _c_u32_p), _ptr(dist, _c_double_p),
                                 _ptr(nxt, _c_double_p), _ptr(ok, _c_u8_p), None, None, None)
        _check(rc)
        return closest, dist, nxt, ok
    radius = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, dtype=np.float64), (n,)))
    nidx = np.full((n, near_cap), GC_NO_NODE, dtype=np.uint32)
    ndist = np.full((n, near_cap), np.inf, dtype=np.float64)
    ncount = np.zeros(n, dtype=np.uint32)
    rc = lib.gc_extend_batch(store.handle, world.handle, _ptr(q, _c_double_p), segp, n, max_distance, min_distance,
                             tolerance, _ptr(radius, _c_double_p), near_cap, _ptr(closest, _c_u32_p),
                             _ptr(dist, _c_double_p), _ptr(nxt, _c_double_p), _ptr(ok, _c_u8_p), _ptr(nidx, _c_u32_p),
                             _ptr(ndist, _c_double_p), _ptr(ncount, _c_u32_p))
    _check(rc)
    return closest, dist, nxt, ok, nidx, ndist, ncount


def informed_extend_batch(store: NodeStore, world: CollisionWorld, q, seg, goal, cost2beat, bias, k_rrt, max_distance,
                          min_distance, d_parent_ptr: int, d_pcost_ptr: int, tolerance=1e-6):
    """Tree::informedExtend (tree.cpp:235-302) for n (sample, tree) pairs; d_parent_ptr / d_pcost_ptr are device
    pointers to int32 / float64 arrays indexed by global slot.  Returns a dict of host arrays."""
    lib = load()
    q = _f64(q, (-1, store.dim))
    n = q.shape[0]
    goal = _f64(goal, (-1, store.dim))
    seg = np.ascontiguousarray(seg, dtype=np.uint32)
    c2b = np.ascontiguousarray(np.broadcast_to(np.asarray(cost2beat, dtype=np.float64), (n,)))
    bs = np.ascontiguousarray(np.broadcast_to(np.asarray(bias, dtype=np.float64), (n,)))
    ok = np.zeros(n, dtype=np.uint8)
    node = np.zeros(n, dtype=np.uint32)
    new = np.zeros((n, store.dim), dtype=np.float64)
    dist = np.zeros(n, dtype=np.float64)
    ecost = np.zeros(n, dtype=np.float64)
    edges = C.c_uint64(0)
    _check(lib.gc_informed_extend_batch(store.handle, world.handle, _ptr(q, _c_double_p), _ptr(seg, _c_u32_p),
                                        _ptr(goal, _c_double_p), _ptr(c2b, _c_double_p), _ptr(bs, _c_double_p), n,
                                        float(k_rrt), float(max_distance), float(min_distance), float(tolerance),
                                        C.c_void_p(d_parent_ptr), C.c_void_p(d_pcost_ptr), _ptr(ok, _c_u8_p),
                                        _ptr(node, _c_u32_p), _ptr(new, _c_double_p), _ptr(dist, _c_double_p),
                                        _ptr(ecost, _c_double_p), C.byref(edges)))
    return {"ok": ok.astype(bool), "node": node, "new_conf": new, "distance": dist, "edge_cost": ecost,
            "edges_checked": int(edges.value)}


def launch_count() -> int:
    n = C.c_uint64(0)
    _check(load().gc_launch_count(C.byref(n)))
    return n.value


def measure_fp32_peak(device: int = 0):
    """(TFLOP/s, implied SM MHz) of a dependent-chain FFMA micro-benchmark."""
    tf, mhz = C.c_double(0), C.c_double(0)
    _check(load().gc_measure_fp32_peak(device, C.byref(tf), C.byref(mhz)))
    return tf.value, mhz.value


def flush_l2(device: int = 0, nbytes: int = 512 << 20):
    _check(load().gc_flush_l2(device, nbytes))


def sample_uniform(lb, ub, seed: int, first_index: int, n: int, device: int = 0) -> np.ndarray:
    lb = _f64(lb)
    ub = _f64(ub)
    out = np.empty((n, lb.shape[0]), dtype=np.float64)
    _check(load().gc_sample_uniform(device, lb.shape[0], _ptr(lb, _c_double_p), _ptr(ub, _c_double_p), seed,
                                    first_index, n, _ptr(out, _c_double_p)))
    return out


def sample_informed(lb, ub, f1, f2, cost, seed: int, first_index: int = 0, device: int = 0):
    """InformedSampler::sample for n problems (foci f1/f2 [n][dim], cost [n] or scalar) on the device Philox stream.
    Returns (samples [n][dim], attempts [n])."""
    lb, ub = _f64(lb), _f64(ub)
    dim = lb.shape[0]
    f1, f2 = _f64(f1, (-1, dim)), _f64(f2, (-1, dim))
    n = f1.shape[0]
    cost = np.ascontiguousarray(np.broadcast_to(np.asarray(cost, dtype=np.float64), (n,)))
    out = np.empty((n, dim), dtype=np.float64)
    att = np.empty(n, dtype=np.uint32)
    _check(load().gc_sample_informed(device, dim, _ptr(lb, _c_double_p), _ptr(ub, _c_double_p), _ptr(f1, _c_double_p),
                                     _ptr(f2, _c_double_p), _ptr(cost, _c_double_p), seed, first_index, n,
                                     _ptr(out, _c_double_p), _ptr(att, _c_u32_p)))
    return out, att
