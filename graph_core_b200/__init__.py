"""graph_core_b200 -- B200-native (sm_100a) nearest-neighbour and edge-checking hot paths of
JRL-CARI-CNR-UNIBS/graph_core behind the reference's operator interfaces.

Layout
  csrc/      hand-written CUDA kernels + the C ABI (include/gcb200.h) -> libgcb200.so
  host/      C++ host mirror of the reference interfaces (plugin classes, lock-step planners)
  capi.py    ctypes binding used by tests, bench.py and Python callers
  worlds.py  the synthetic benchmark worlds of BASELINE.json (seeded, Philox)
  tree_yaml.py  the reference's tree YAML format, written from / read into the SoA store
"""
from . import capi  # noqa: F401
from .capi import (CollisionWorld, GcError, NodeStore, ShardedStore, device_count, device_info, extend_batch,  # noqa: F401
                   sample_uniform)

__all__ = ["capi", "NodeStore", "ShardedStore", "CollisionWorld", "GcError", "device_count", "device_info", "extend_batch",
           "sample_uniform"]
