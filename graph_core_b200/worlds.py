"""Synthetic benchmark worlds of BASELINE.json `configs` (C0..C4).

Everything is generated from Philox4x32-10 streams (key = seed, counter = element index) so that
the CPU oracle and the CUDA path see identical inputs on any machine.  The Philox restatement
here is pure numpy integer arithmetic; the device sampler (csrc/planner.cu) is checked against it
bit for bit in tests/test_sampler_gpu.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Vectorised Philox4x32-10; c* are uint64 arrays holding 32-bit values."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def philox_u01(seed: int, first: int, n: int, dim: int) -> np.ndarray:
    """n x dim uniforms in [0,1): sample i, dims (2b, 2b+1) come from counter (i_lo, i_hi, b, 0)."""
    idx = np.arange(first, first + n, dtype=np.uint64)
    pairs = (dim + 1) // 2
    out = np.empty((n, 2 * pairs), dtype=np.float64)
    for b in range(pairs):
        r0, r1, r2, r3 = philox4x32_10(idx & _MASK, idx >> np.uint64(32), np.full(n, b, np.uint64),
                                       np.zeros(n, np.uint64), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        out[:, 2 * b] = (((r0 << np.uint64(32)) | r1) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
        out[:, 2 * b + 1] = (((r2 << np.uint64(32)) | r3) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    return out[:, :dim]


def uniform_samples(lb, ub, seed: int, first: int, n: int) -> np.ndarray:
    """UniformSampler::sample (uniform_sampler.cpp:34-39) with Random() = 2u-1; same op order as the device."""
    lb = np.asarray(lb, dtype=np.float64)
    ub = np.asarray(ub, dtype=np.float64)
    u = philox_u01(seed, first, n, lb.shape[0])
    c = 0.5 * (lb + ub)
    hw = 0.5 * (lb - ub)
    return c + (2.0 * u - 1.0) * hw


def _u(seed, n, dim, lo, hi):
    return lo + philox_u01(seed, 0, n, dim) * (hi - lo)


@dataclass
class Scenario:
    name: str
    dim: int
    kind: str                      # cube | boxes | cspheres | sphere_arm
    lb: np.ndarray
    ub: np.ndarray
    start: np.ndarray
    goal: np.ndarray
    max_distance: float
    min_distance: float = 0.01
    rewire_factor: float = 1.1
    utopia_tolerance: float = 0.01
    params: dict = field(default_factory=dict)

    def make_world(self, device: int = 0):
        """CUDA world (product path)."""
        from .capi import CollisionWorld
        p = self.params
        if self.kind == "cube":
            return CollisionWorld.cube(self.dim, p["threshold"], device)
        if self.kind == "boxes":
            return CollisionWorld.boxes(p["lo"], p["hi"], device)
        if self.kind == "cspheres":
            return CollisionWorld.cspheres(p["centers"], p["radii"], device)
        return CollisionWorld.sphere_arm(p["base_tf"], p["joint_tf"], p["rs_link"], p["rs_local"], p["obs"], device)


def scenario_c0() -> Scenario:
    """The reference's own test scene (tests/src/compute_path_test.cpp:51-83, graph_core_tests_params.yaml)."""
    d = 3
    return Scenario("C0", d, "cube", np.full(d, -2.5), np.full(d, 2.5), np.full(d, -1.5), np.full(d, 1.5),
                    max_distance=0.25, params={"threshold": 1.0})


def scenario_c1() -> Scenario:
    """2-D unit square with 10 axis-aligned boxes."""
    start, goal = np.array([0.05, 0.05]), np.array([0.95, 0.95])
    lo, hi = [], []
    i = 0
    while len(lo) < 10:
        u = philox_u01(101, i, 1, 4)[0]
        i += 1
        c = 0.15 + u[:2] * 0.70
        h = 0.03 + u[2:] * 0.07
        l, r = c - h, c + h
        if any(np.all(p >= l) and np.all(p <= r) for p in (start, goal)):
            continue
        lo.append(l)
        hi.append(r)
    return Scenario("C1", 2, "boxes", np.zeros(2), np.ones(2), start, goal, max_distance=0.1,
                    params={"lo": np.array(lo), "hi": np.array(hi)})


def scenario_c2(n_spheres: int = 64) -> Scenario:
    """6-DoF C-space with hypersphere obstacles."""
    d = 6
    start, goal = np.full(d, -2.0), np.full(d, 2.0)
    cs, rs = [], []
    i = 0
    while len(cs) < n_spheres:
        u = philox_u01(201, i, 1, d + 1)[0]
        i += 1
        c = (-math.pi + u[:d] * 2 * math.pi).astype(np.float32)
        r = np.float32(1.2 + u[d] * 1.0)  # large enough that ~1/4 of the 6-D box is blocked
        if min(np.linalg.norm(c - start), np.linalg.norm(c - goal)) < float(r) + 0.1:
            continue
        cs.append(c)
        rs.append(r)
    return Scenario("C2", d, "cspheres", np.full(d, -math.pi), np.full(d, math.pi), start, goal, max_distance=1.0,
                    params={"centers": np.array(cs, np.float32), "radii": np.array(rs, np.float32)})


def arm_model(n_joints: int = 7):
    """Serial revolute arm: joint j rotates about z, then F_j = Trans(0,0,L_j) * Rx(+-90 deg)."""
    lengths = [0.30, 0.0, 0.30, 0.0, 0.25, 0.0, 0.10]
    lengths = (lengths * ((n_joints + 6) // 7))[:n_joints]
    base = np.zeros(12, np.float32)
    base[[0, 4, 8]] = 1.0
    joint = np.zeros((n_joints, 12), np.float32)
    for j in range(n_joints):
        sgn = 1.0 if j % 2 == 0 else -1.0
        # Rx(a), a = +-90deg : [[1,0,0],[0,0,-s],[0,s,0]]
        joint[j, :9] = [1, 0, 0, 0, 0, -sgn, 0, sgn, 0]
        joint[j, 9:] = [0, 0, lengths[j]]
    link, local = [], []
    radii = [0.08, 0.08, 0.07, 0.07, 0.07, 0.07, 0.06, 0.06, 0.06, 0.06, 0.05, 0.05, 0.05, 0.05, 0.05, 0.05]
    k = 0
    for j in range(n_joints):
        L = lengths[j]
        for f in (1.0 / 3.0, 2.0 / 3.0):
            z = L * f if L > 0 else 0.0
            y = 0.0 if L > 0 else (0.03 if f < 0.5 else -0.03)
            link.append(j)
            local.append([0.0, y, z, radii[k % len(radii)]])
            k += 1
    for z in (0.03, 0.08):  # tool spheres in the frame after the last joint transform
        link.append(n_joints)
        local.append([0.0, 0.0, z, radii[k % len(radii)]])
        k += 1
    return base, joint, np.array(link, np.int32), np.array(local, np.float32)


def _c3_obstacles(n_obs: int):
    obs = []
    i = 0
    while len(obs) < n_obs:
        u = philox_u01(301, i, 1, 4)[0]
        i += 1
        c = -1.1 + u[:3] * 2.2
        r = 0.02 + u[3] * 0.03
        if math.hypot(c[0], c[1]) < 0.25 + r and -0.1 < c[2] < 0.45:
            continue  # keep the base column clear: link 0 never moves
        obs.append([c[0], c[1], c[2], r])
    return np.array(obs, np.float32)


def scenario_c3(n_obs: int = 1000, checker=None, problem: int = 0) -> Scenario:
    """7-DoF sphere-model arm vs a 1k-sphere cloud (obstacle centres U[-1.1,1.1]^3, radii U[0.02,0.05]).

    `checker(params, q[n,7]) -> bool[n]` (True = free) selects the start/goal pair: pair number
    `problem` of the Philox stream (seed 302) among those whose endpoints are free and whose straight
    connection is blocked (so the planner has work to do).  Tests pass the oracle's checker, the
    product passes the CUDA world's; both give the same booleans.  Without a checker a fixed pair is
    returned unvalidated.
    """
    d = 7
    base, joint, link, local = arm_model(d)
    obs = _c3_obstacles(n_obs)
    params = {"base_tf": base, "joint_tf": joint, "rs_link": link, "rs_local": local, "obs": obs}
    start = np.array([0.0, -0.6, 0.0, 0.9, 0.0, 0.6, 0.0])
    goal = np.array([1.8, 0.7, -0.8, -0.9, 0.7, -0.5, 1.0])
    if checker is not None:
        pairs = c3_problems(checker, params, problem + 1)
        start, goal = pairs[problem]
    return Scenario("C3", d, "sphere_arm", np.full(d, -math.pi), np.full(d, math.pi), start, goal, max_distance=0.5,
                    params=params)


def c3_problems(checker, params, n: int, seed: int = 302):
    """First n start/goal pairs (Philox seed 302) with free endpoints and a blocked straight connection."""
    out = []
    first = 0
    t = np.linspace(0.0, 1.0, 65)[:, None]
    chunk = 64 if n <= 256 else 4096  # candidates per round (the stream is counter based: chunking does not change it)
    while len(out) < n:
        s, g = c3_problem_pairs(chunk, seed, first)
        first += chunk
        ok = np.nonzero(checker(params, s) & checker(params, g))[0]
        if ok.size == 0:
            continue
        # the 65 states of every candidate's straight line, one checker call for the whole chunk
        lines = s[ok][:, None, :] + (g[ok] - s[ok])[:, None, :] * t[None, :, :]
        free = checker(params, lines.reshape(-1, s.shape[1])).reshape(ok.size, t.shape[0]).all(axis=1)
        for i in ok[~free]:
            out.append((s[i].copy(), g[i].copy()))
            if len(out) == n:
                break
    return out


def c5_problems(checker, params, n: int, seed: int = 501, reach: float = 1.5, cost_factor: float = 1.5):
    """BASELINE config 5 as benchmarked (C5'): n AnytimeRRT queries in the C3 world.  Start uniform in the joint box
    (Philox seed 501), goal = start + reach * (uniform direction of the same draw), both collision free and inside the
    box, straight connection blocked.  The solver's own sampler is an InformedSampler(start, goal, lb, ub, cost =
    cost_factor * |goal - start|) -- with BASELINE's literal C5 (goals anywhere in the (2 pi)^7 box, unbounded sampler)
    the reference does not find a first solution within 3 x 2 000 updates, so its improve rounds never start.
    Returns (starts [n][7], goals [n][7], sampler_cost0 [n])."""
    starts, goals = [], []
    first = 0
    t = np.linspace(0.0, 1.0, 65)[:, None]
    chunk = 64 if n <= 256 else 4096  # candidates per round (counter-based stream: chunking does not change it)
    while len(starts) < n:
        u = philox_u01(seed, first, chunk, 14)
        first += chunk
        s = -math.pi + u[:, :7] * 2 * math.pi
        v = 2.0 * u[:, 7:] - 1.0
        v = v / np.sqrt((v * v).sum(axis=1))[:, None]
        g = s + reach * v
        ok = np.nonzero(checker(params, s) & checker(params, g) & (np.abs(g) <= math.pi).all(axis=1))[0]
        if ok.size == 0:
            continue
        lines = s[ok][:, None, :] + (g[ok] - s[ok])[:, None, :] * t[None, :, :]
        free = checker(params, lines.reshape(-1, 7)).reshape(ok.size, t.shape[0]).all(axis=1)
        for i in ok[~free]:
            starts.append(s[i].copy())
            goals.append(g[i].copy())
            if len(starts) == n:
                break
    starts, goals = np.array(starts), np.array(goals)
    d = np.sqrt(((goals - starts) ** 2).sum(axis=1))
    return starts, goals, cost_factor * d


def c3_problem_pairs(n: int, seed: int = 302, first: int = 0):
    """Raw start/goal pairs for the multi-problem C3/C5 runs (before validity filtering)."""
    u = philox_u01(seed, first, n, 14)
    q = -math.pi + u * 2 * math.pi
    return q[:, :7], q[:, 7:]


def c4_nodes(n: int = 1 << 20, dim: int = 12, seed: int = 401) -> np.ndarray:
    return -math.pi + philox_u01(seed, 0, n, dim) * (2 * math.pi)


def c4_queries(n: int = 4096, dim: int = 12, seed: int = 402) -> np.ndarray:
    return -math.pi + philox_u01(seed, 0, n, dim) * (2 * math.pi)


def c4_radius(expected_hits: float, n: int = 1 << 20, dim: int = 12) -> float:
    """Radius whose ball holds `expected_hits` of n uniform nodes in [-pi,pi]^dim (ignoring the boundary)."""
    unit_ball = math.pi ** (dim / 2) / math.gamma(dim / 2 + 1)
    return (expected_hits / n * (2 * math.pi) ** dim / unit_ball) ** (1.0 / dim)
