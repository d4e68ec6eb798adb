"""Tree YAML round trip from the SoA node store (the reference's only on-disk tree format).

Mirrors `Tree::toYAML` / `Tree::fromYAML` (reference `src/graph_core/graph/tree.cpp:1188-1227`, `:1272-1389`) and
`Node::toYAML` / `Node::fromYAML` (`src/graph_core/graph/node.cpp:541-568`):

    max_distance: 0.5
    use_kdtree: true
    nodes:
      - [q0, q1, ...]          # one flow sequence per node, in the order of NearestNeighbors::getNodes()
    connections:
      - [i, j]                 # parent index, child index (indices into `nodes`)

The SoA store keeps the nodes in append order and a parent slot per node, so the document is produced straight
from `LockstepRRTStar.dump()` (or any `(conf, parents)` pair) without building Node/Connection objects.

Differences from the reference, on purpose:
  * connection pairs are listed by parent, children in ascending slot order.  The reference lists the children of a
    node in the order its child connections were added, which after rewiring depends on history; the SET of pairs
    (and therefore the loaded tree) is the same.
  * doubles are written with the shortest representation that reads back to the same bits (yaml-cpp writes 17
    significant digits); both parse to identical doubles.
`tree_from_yaml` keeps upstream's rules: the root is the first node without a parent connection, `max_distance`
absent -> the longest connection, `use_kdtree` absent -> true, and a document whose root is not node 0 is refused
(upstream would silently drop node 0 and add the root twice: `tree.cpp:1378-1386` adds `nodes[1:]` only).
"""
from __future__ import annotations

import math

import numpy as np

__all__ = ["tree_to_yaml", "tree_from_yaml", "connections_from_parents", "TreeDocument"]


class TreeDocument:
    """A loaded tree: `conf[n, dim]`, `parents[n]` (-1 for the root), `pcost[n]` (Euclidean cost of the parent
    connection, `EuclideanMetrics::cost`), `max_distance`, `use_kdtree`."""

    def __init__(self, conf, parents, pcost, max_distance, use_kdtree):
        self.conf = conf
        self.parents = parents
        self.pcost = pcost
        self.max_distance = max_distance
        self.use_kdtree = use_kdtree

    def to_store(self, capacity: int | None = None, device: int = 0):
        """Uploads the nodes into a fresh device store (slot i = document node i)."""
        from . import capi
        n, dim = self.conf.shape
        store = capi.NodeStore(dim, max(int(capacity or 0), n, 1), device=device)
        store.append(self.conf)
        return store


def connections_from_parents(parents) -> np.ndarray:
    """(parent, child) pairs grouped by parent ascending, children ascending."""
    parents = np.asarray(parents, dtype=np.int64)
    child = np.nonzero(parents >= 0)[0]
    order = np.lexsort((child, parents[child]))
    return np.stack([parents[child][order], child[order]], axis=1).astype(np.int64) if child.size else \
        np.zeros((0, 2), np.int64)


def _num(v: float) -> str:
    v = float(v)
    if math.isnan(v):
        return ".nan"
    if math.isinf(v):
        return ".inf" if v > 0 else "-.inf"
    return repr(v)


def tree_to_yaml(conf, parents, max_distance: float, use_kdtree: bool = True) -> str:
    """conf: [n, dim] float64 in store order; parents: [n] parent slot (-1 for the root)."""
    conf = np.asarray(conf, dtype=np.float64)
    if conf.ndim != 2:
        raise ValueError("conf must be [n, dim]")
    parents = np.asarray(parents, dtype=np.int64)
    if parents.shape != (conf.shape[0],):
        raise ValueError("parents must have one entry per node")
    if np.any(parents >= conf.shape[0]) or np.any(parents == np.arange(conf.shape[0])):
        raise ValueError("parent slots must name other nodes of the tree")
    out = ["max_distance: " + _num(max_distance), "use_kdtree: " + ("true" if use_kdtree else "false")]
    if conf.shape[0]:
        out.append("nodes:")
        out.extend("  - [" + ", ".join(_num(x) for x in row) + "]" for row in conf)
    else:
        out.append("nodes: []")
    pairs = connections_from_parents(parents)
    if pairs.shape[0]:
        out.append("connections:")
        out.extend("  - [%d, %d]" % (int(a), int(b)) for a, b in pairs)
    else:
        out.append("connections: []")
    return "\n".join(out) + "\n"


def tree_from_yaml(text_or_doc) -> TreeDocument:
    """Parses a tree document (text, or the already-parsed mapping).  Raises ValueError where `Tree::fromYAML` logs an
    error and returns nullptr."""
    if isinstance(text_or_doc, (str, bytes)):
        import yaml
        doc = yaml.safe_load(text_or_doc)
    else:
        doc = text_or_doc
    if not isinstance(doc, dict):
        raise ValueError("a tree document is a mapping")
    if "nodes" not in doc:
        raise ValueError("Cannot load a tree from a YAML without 'nodes' field")
    if "connections" not in doc:
        raise ValueError("Cannot load a tree from a YAML without 'connections' field")
    nodes, conns = doc["nodes"], doc["connections"]
    if not isinstance(nodes, list):
        raise ValueError("'nodes' is not a sequence")
    if not isinstance(conns, list):
        raise ValueError("'connections' is not a sequence")
    use_kdtree = bool(doc.get("use_kdtree", True))
    for q in nodes:
        if not isinstance(q, list):
            raise ValueError("Cannot load a node from YAML::Node which does not contain a sequence")
    n = len(nodes)
    dim = len(nodes[0]) if n else 0
    if any(len(q) != dim for q in nodes):
        raise ValueError("nodes of different dimension")
    conf = np.array(nodes, dtype=np.float64).reshape(n, dim)
    parents = np.full(n, -1, np.int64)
    pcost = np.zeros(n, np.float64)
    longest = -1.0
    for c in conns:
        if not isinstance(c, list) or len(c) < 2:
            raise ValueError("'connections' is not a sequence of sequences")
        a, b = int(c[0]), int(c[1])
        if not (0 <= a < n and 0 <= b < n):
            raise ValueError("connection [%d, %d] names a node outside the document" % (a, b))  # upstream: vector::at
        if parents[b] != -1:
            raise ValueError("node %d has two parents: not a tree" % b)
        parents[b] = a
        # EuclideanMetrics::cost = fp64 2-norm of the difference, sequential sum of squares (euclidean_metrics.h:72-95)
        acc = 0.0
        for k in range(dim):
            d = float(conf[a, k]) - float(conf[b, k])
            acc += d * d
        pcost[b] = math.sqrt(acc)
        longest = max(longest, pcost[b])
    roots = np.nonzero(parents < 0)[0]
    if n and (roots.size == 0 or roots[0] != 0):
        raise ValueError("the root (first node without a parent) must be node 0")
    max_distance = float(doc["max_distance"]) if "max_distance" in doc else longest
    return TreeDocument(conf, parents, pcost, max_distance, use_kdtree)
