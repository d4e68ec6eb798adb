// forest.cu -- device-resident lock-step RRT* trees (include/gcb200.h "gc_forest"): parent slot and connection
// cost of every node live in HBM next to the node store, and one RRTStar::update of P independent problems is a
// fixed sequence of stream-ordered launches with no host decision in between.
//
//   RRTStar::update              src/graph_core/solvers/rrt_star.cpp:89-163
//   RRTStar::updateRewireRadius  src/graph_core/solvers/rrt_star.cpp:278-286
//   Tree::extend / extendOnly    src/graph_core/graph/tree.cpp:130-161
//   Tree::rewireOnly             src/graph_core/graph/tree.cpp:381-513
//   Tree::costToNode             src/graph_core/graph/tree.cpp:767-789
//   Tree::getConnectionToNode    src/graph_core/graph/tree.cpp:791-813 (the solution path keeps its connection costs)
//   InformedSampler::setCost     src/graph_core/samplers/informed_sampler.cpp:181-215 (specific volume)
//
// Launch sequence of one step (all on the store's stream):
//   [samples H2D | informed sampler] -> prologue -> nearest -> steer -> extension edges -> radius-near ->
//   extend (append the node, cost-to-come walks, parent candidates in cost order, first edge batch) -> edges A ->
//   resolve (parent winner or open candidates, children candidates, goal edge: second edge batch) -> edges B ->
//   replay (parent choice, children phase in near order, goal connection, solution cost) -> D2H of the path costs
//   -> host callback (libm pow() of the rewire radius for the problems whose path cost changed).
// The sequential decisions of rewireOnly (a rewired child changes the cost-to-come of the near nodes behind it)
// are replayed by one warp per problem: the lanes evaluate 32 near nodes at a time, the first taker in near order
// is applied and the lanes behind it are re-evaluated with fresh cost-to-come walks.
#include <math_constants.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "gc_common.cuh"

namespace gc
{
enum : uint32_t
{
  kErrNearOverflow = 1u,   // a radius-near list was longer than near_cap
  kErrEdgeOverflow = 2u,   // the second edge batch did not fit edges_cap
  kErrSegmentFull = 4u,    // a tree reached the segment capacity
  kErrDepthOverflow = 8u,  // a solution path is longer than depth_cap
  kErrSpecMiss = 16u,      // the replay asked for an edge outside the speculative set (cannot happen by construction)
  kErrHostTimeout = 32u,   // the host radius thread did not answer
};

struct __align__(16) FNode
{
  double pcost;    // cost of the connection parent -> node
  int32_t parent;  // slot of the parent, -1 for the root
  int32_t pad;
};

struct FProb
{
  double cost, best_utopia, goal_cost, r_rewire;
  double c2n0, c2n1;
  int32_t goal_slot;
  uint32_t tree_size;
  uint32_t new_slot, closest, near_n, nc;
  uint32_t edgeA_first, edgeA_count, par_known, parB_count, edge_first, edge_count;
  int32_t par_winner, goal_edge;
  uint32_t sol_len;
  uint8_t solved, can_improve, active, extended;
};

struct ForestDev
{
  int D;
  uint32_t P, cap_seg, near_cap, edges_cap, depth_cap, first_batch;
  size_t row_stride;
  double max_distance, tolerance, utopia_tol;
  float* x32;
  double* x64;
  uint32_t* used;
  FNode* nodes;  // [P][cap_seg]
  FProb* prob;   // [P]
  double* path_cost;         // [P] (+inf while unsolved): the informed sampler's cost
  double* r_rrt;             // [P] rewire_factor * pow(2(1+1/d),1/d) * pow(specific_volume,1/d), host libm
  const double* card_table;  // [cap_seg+1] pow(log(n+1)/(n+1), 1/d), host libm
  uint8_t* changed;          // [P] path cost changed in this step
  const double* goals;       // [P][D]
  const double* samples;     // [P][D]
  uint32_t* seg_q;           // [P] segment queried by problem p: p, or P (empty) when inactive
  uint32_t* seg_near;        // [P] the same for the radius-near pass: P also when the extension of this step failed
  double* radius;
  uint32_t* closest;
  double* dist;
  double* c1;
  double* c2;
  uint8_t* ok_ext;
  uint32_t* near_idx;
  double* near_dist;
  uint32_t* near_count;
  double* c2near;      // [P][near_cap] cost-to-come of the near nodes at the start of the step
  uint32_t* par_cand;  // [P][near_cap] near positions of the parent candidates by (cost through them, position)
  double* par_total;
  uint32_t* tmp_pos;
  double* tmp_total;
  int32_t* chi_edge;  // [P][near_cap] edge of the second batch (relative to the problem's slice) or -1
  double *eA1, *eA2;
  uint8_t* okA;
  double *eB1, *eB2;
  uint8_t* okB;
  uint32_t* ctr;              // [0] edges in batch A, [1] edges in batch B, [2] error bits, [3] finished replay
                              // blocks, [4] problems whose path cost changed, [5] requests made, [6] requests installed
  uint32_t* chg_p;            // [P] the problems of [4]
  // mapped pinned host memory shared with the radius thread
  uint32_t* hs_ctl;           // [0] request number (device -> host), [1] entries, [2] answered request (host -> device)
  uint32_t* hs_p;             // [P] problem
  double* hs_cost;            // [P] its new path cost
  double* hs_rrt;             // [P] the host's answer
  unsigned long long* stats;  // see gc_forest_stats
  uint32_t* sol_slot;         // [P][depth_cap] slots of the solution, root's child first
  double* sol_cost;
};

// Eigen (a-b).norm(): sequential, individually rounded
__device__ __forceinline__ double f_dist(const double* a, const double* b, int D)
{
  double s = 0.0;
  for (int i = 0; i < D; i++)
  {
    const double t = __dsub_rn(a[i], b[i]);
    s = __dadd_rn(s, __dmul_rn(t, t));
  }
  return __dsqrt_rn(s);
}

// Tree::costToNode (tree.cpp:767-789): connection costs summed from the node up to the root
__device__ __forceinline__ double f_walk(const FNode* nodes, int32_t node)
{
  double cost = 0.0;
  while (node != 0)
  {
    if (node < 0)
      return CUDART_INF;
    const FNode nd = nodes[node];
    if (nd.parent < 0)
      return CUDART_INF;
    cost = __dadd_rn(cost, nd.pcost);
    node = nd.parent;
  }
  return cost;
}

__device__ __forceinline__ void f_store_conf(const ForestDev& f, uint32_t p, uint32_t slot, const double* conf, int d)
{
  const size_t g = (size_t)p * f.cap_seg + slot;
  const double v = conf[d];
  f.x64[g * f.D + d] = v;
  f.x32[(size_t)d * f.row_stride + g] = (float)v;
}

// ---- prologue: RRTStar::update's early returns and updateRewireRadius (rrt_star.cpp:93-108, 278-286) ----------
__global__ void forest_prologue_kernel(ForestDev f)
{
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p == 0)
  {
    f.ctr[0] = 0;
    f.ctr[1] = 0;
    atomicAdd(&f.stats[6], 1ull);
  }
  if (p >= f.P)
    return;
  FProb& pr = f.prob[p];
  bool act = false;
  if (pr.can_improve)
  {
    if (pr.cost <= __dmul_rn(f.utopia_tol, pr.best_utopia))
      pr.can_improve = 0;  // "Solution already optimal"
    else
    {
      pr.r_rewire = __dmul_rn(f.r_rrt[p], f.card_table[pr.tree_size]);
      act = true;
    }
  }
  pr.active = act;
  pr.extended = 0;
  f.changed[p] = 0;
  f.seg_q[p] = act ? p : f.P;
  f.radius[p] = pr.r_rewire;
  const uint32_t am = __activemask();
  const uint32_t m = __ballot_sync(am, act);
  if (act && (threadIdx.x & 31) == (uint32_t)(__ffs(m) - 1))
    atomicAdd(&f.stats[0], (unsigned long long)__popc(m));
  // largest tree (the host sizes the scans of later steps with it, read back asynchronously)
  const uint32_t mx = __reduce_max_sync(am, pr.tree_size);
  if ((threadIdx.x & 31) == (uint32_t)(__ffs(am) - 1) && mx > f.ctr[7])
    atomicMax(&f.ctr[7], mx);
}

// ---- gate: only the problems whose extension succeeded (Tree::tryExtendFromNode, tree.cpp:69-81) go on to the
// radius-near pass -- three quarters of the C3 samples end at a colliding extension edge, and their near scans were
// read for nothing
__global__ void forest_gate_kernel(ForestDev f)
{
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= f.P)
    return;
  const bool ok = f.seg_q[p] == p && f.closest[p] != GC_NO_NODE && ((f.dist[p] < f.tolerance) ? true : f.ok_ext[p] != 0);
  f.seg_near[p] = ok ? p : f.P;
}

// ---- extend: Tree::extend's bookkeeping + the parent phase's candidate set (tree.cpp:130-161, 428-463) -------
__global__ void __launch_bounds__(128) forest_extend_kernel(ForestDev f)
{
  const uint32_t p = blockIdx.x;
  FProb& pr = f.prob[p];
  __shared__ int s_go;
  __shared__ uint32_t s_cnt, s_base;
  __shared__ double s_c2n0;
  const int D = f.D;
  if (threadIdx.x == 0)
  {
    int go = 0;
    if (pr.active)
    {
      const uint32_t closest = f.closest[p];
      // Tree::tryExtendFromNode (tree.cpp:69-81): no edge check when the sample sits on the node
      bool ok = closest != GC_NO_NODE && ((f.dist[p] < f.tolerance) ? true : f.ok_ext[p] != 0);
      if (ok && f.used[p] + 2u > f.cap_seg)
      {
        atomicOr(&f.ctr[2], kErrSegmentFull);
        ok = false;
      }
      if (ok && f.near_count[p] > f.near_cap)
      {
        atomicOr(&f.ctr[2], kErrNearOverflow);
        ok = false;
      }
      go = ok ? 1 : 0;
    }
    s_go = go;
    s_cnt = 0;
  }
  __syncthreads();
  if (!s_go)
    return;
  const uint32_t slot = f.used[p];
  const uint32_t closest = f.closest[p];
  const double* next = f.c2 + (size_t)p * D;
  FNode* nodes = f.nodes + (size_t)p * f.cap_seg;
  const double* seg64 = f.x64 + (size_t)p * f.cap_seg * D;
  if ((int)threadIdx.x < D)
    f_store_conf(f, p, slot, next, threadIdx.x);
  const uint32_t near_n = f.near_count[p];
  const uint32_t* nidx = f.near_idx + (size_t)p * f.near_cap;
  const double* ndist = f.near_dist + (size_t)p * f.near_cap;
  double* c2near = f.c2near + (size_t)p * f.near_cap;
  if (threadIdx.x == 0)
  {
    FNode nd;
    nd.parent = (int32_t)closest;
    nd.pcost = f_dist(seg64 + (size_t)closest * D, next, D);
    nd.pad = 0;
    nodes[slot] = nd;
    // costToNode(new node): its own connection first, then the chain of the closest node
    double c = __dadd_rn(0.0, nd.pcost);
    int32_t node = (int32_t)closest;
    while (node != 0)
    {
      const FNode q = nodes[node];
      c = __dadd_rn(c, q.pcost);
      node = q.parent;
    }
    s_c2n0 = c;
  }
  else
  {
    // the walks of the other threads never visit the new slot: the near lists were taken before the append
    for (uint32_t i = threadIdx.x - 1; i < near_n; i += blockDim.x - 1)
      c2near[i] = f_walk(nodes, (int32_t)nidx[i]);
  }
  __syncthreads();
  const double c2n0 = s_c2n0;
  uint32_t* tpos = f.tmp_pos + (size_t)p * f.near_cap;
  double* ttot = f.tmp_total + (size_t)p * f.near_cap;
  for (uint32_t i = threadIdx.x; i < near_n; i += blockDim.x)
  {
    const uint32_t nn = nidx[i];
    if (nn == closest)
      continue;
    const double c = c2near[i];
    if (c >= c2n0)
      continue;
    const double tot = __dadd_rn(c, ndist[i]);
    if (tot >= c2n0)
      continue;
    const uint32_t k = atomicAdd(&s_cnt, 1u);
    tpos[k] = i;
    ttot[k] = tot;
  }
  __syncthreads();
  const uint32_t nc = s_cnt;
  // rank by (cost through the candidate, near position): the reference's walk keeps the first strictly cheaper
  // collision-free candidate, i.e. the first free one in this order
  uint32_t* pcand = f.par_cand + (size_t)p * f.near_cap;
  double* ptot = f.par_total + (size_t)p * f.near_cap;
  for (uint32_t k = threadIdx.x; k < nc; k += blockDim.x)
  {
    const double tk = ttot[k];
    const uint32_t ik = tpos[k];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < nc; j++)
    {
      const double tj = ttot[j];
      rank += (tj < tk || (tj == tk && tpos[j] < ik)) ? 1u : 0u;
    }
    pcand[rank] = ik;
    ptot[rank] = tk;
  }
  const uint32_t cntA = min(nc, f.first_batch);
  if (threadIdx.x == 0)
  {
    s_base = cntA ? atomicAdd(&f.ctr[0], cntA) : 0u;
    f.used[p] = slot + 1;
    pr.tree_size = slot + 1;
    pr.new_slot = slot;
    pr.closest = closest;
    pr.near_n = near_n;
    pr.nc = nc;
    pr.c2n0 = c2n0;
    pr.extended = 1;
    pr.edgeA_count = cntA;
    pr.par_winner = -1;
    atomicAdd(&f.stats[1], 1ull);
    atomicAdd(&f.stats[5], (unsigned long long)near_n);
  }
  __syncthreads();
  if (threadIdx.x == 0)
    pr.edgeA_first = s_base;
  // first edge batch: near node -> new node for the best candidates
  for (uint32_t t = threadIdx.x; t < cntA * (uint32_t)D; t += blockDim.x)
  {
    const uint32_t k = t / D, d = t % D;
    const uint32_t nn = nidx[pcand[k]];
    f.eA1[(size_t)(s_base + k) * D + d] = seg64[(size_t)nn * D + d];
    f.eA2[(size_t)(s_base + k) * D + d] = next[d];
  }
}

// ---- resolve: what the first batch decided; second batch = open parent candidates + children + goal edge ---
__global__ void __launch_bounds__(128) forest_resolve_kernel(ForestDev f)
{
  const uint32_t p = blockIdx.x;
  FProb& pr = f.prob[p];
  if (!pr.extended)
    return;
  __shared__ uint32_t s_parB, s_chi, s_base, s_known;
  __shared__ int s_goal, s_fit;
  __shared__ double s_c2n1;
  const int D = f.D;
  const uint32_t slot = pr.new_slot;
  const double* seg64 = f.x64 + (size_t)p * f.cap_seg * D;
  const double* me = seg64 + (size_t)slot * D;
  const double* goal = f.goals + (size_t)p * D;
  if (threadIdx.x == 0)
  {
    const uint32_t nc = pr.nc, cntA = pr.edgeA_count;
    int winner = -1;
    for (uint32_t k = 0; k < cntA && winner < 0; k++)
      if (f.okA[pr.edgeA_first + k])
        winner = (int)k;
    const double* ptot = f.par_total + (size_t)p * f.near_cap;
    uint32_t parB = 0;
    double c2n1 = pr.c2n0;
    if (winner >= 0)
      c2n1 = ptot[winner];
    else if (cntA < nc)
    {
      parB = nc - cntA;
      c2n1 = ptot[cntA];  // the best the open candidates can still reach: a lower bound
    }
    pr.par_winner = winner;
    pr.par_known = cntA;
    pr.parB_count = parB;
    pr.c2n1 = c2n1;
    s_parB = parB;
    s_known = cntA;
    s_c2n1 = c2n1;
    s_chi = 0;
    s_goal = (!pr.solved && f_dist(me, goal, D) < f.max_distance) ? 1 : 0;
  }
  __syncthreads();
  const uint32_t near_n = pr.near_n, parB = s_parB;
  const double c2n1 = s_c2n1;
  const uint32_t* nidx = f.near_idx + (size_t)p * f.near_cap;
  const double* ndist = f.near_dist + (size_t)p * f.near_cap;
  const double* c2near = f.c2near + (size_t)p * f.near_cap;
  int32_t* chi = f.chi_edge + (size_t)p * f.near_cap;
  // children phase candidates (tree.cpp:471-510) under the exact cost-to-come when the parent is known, under its
  // lower bound otherwise: a superset either way, the cost of a near node only falls while the phase runs
  for (uint32_t i = threadIdx.x; i < near_n; i += blockDim.x)
  {
    const uint32_t nn = nidx[i];
    int32_t e = -1;
    if (nn != 0)
    {
      const double c = c2near[i];
      if (!(c2n1 >= c) && !(__dadd_rn(c2n1, ndist[i]) >= c))
        e = (int32_t)(parB + atomicAdd(&s_chi, 1u));
    }
    chi[i] = e;
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    const uint32_t count = parB + s_chi + (uint32_t)s_goal;
    uint32_t base = count ? atomicAdd(&f.ctr[1], count) : 0u;
    int fit = 1;
    if (base + count > f.edges_cap)
    {
      atomicOr(&f.ctr[2], kErrEdgeOverflow);
      fit = 0;
    }
    s_base = base;
    s_fit = fit;
    pr.edge_first = base;
    pr.edge_count = fit ? count : 0u;
    pr.goal_edge = s_goal ? (int32_t)(parB + s_chi) : -1;
    if (!fit)
      pr.extended = 2;  // the tree of this problem is no longer the reference's; the replay leaves it alone
  }
  __syncthreads();
  if (!s_fit)
    return;
  const uint32_t base = s_base, known = s_known, nchi = s_chi;
  const uint32_t* pcand = f.par_cand + (size_t)p * f.near_cap;
  for (uint32_t t = threadIdx.x; t < parB * (uint32_t)D; t += blockDim.x)  // near node -> new node
  {
    const uint32_t k = t / D, d = t % D;
    const uint32_t nn = nidx[pcand[known + k]];
    f.eB1[(size_t)(base + k) * D + d] = seg64[(size_t)nn * D + d];
    f.eB2[(size_t)(base + k) * D + d] = me[d];
  }
  for (uint32_t i = threadIdx.x; i < near_n; i += blockDim.x)  // new node -> near node
  {
    const int32_t e = chi[i];
    if (e < 0)
      continue;
    const uint32_t nn = nidx[i];
    for (int d = 0; d < D; d++)
    {
      f.eB1[(size_t)(base + e) * D + d] = me[d];
      f.eB2[(size_t)(base + e) * D + d] = seg64[(size_t)nn * D + d];
    }
  }
  if (s_goal && (int)threadIdx.x < D)  // new node -> goal
  {
    const uint32_t e = base + parB + nchi;
    f.eB1[(size_t)e * D + threadIdx.x] = me[threadIdx.x];
    f.eB2[(size_t)e * D + threadIdx.x] = goal[threadIdx.x];
  }
}

// Path(getConnectionToNode(goal)) + Path::cost(): the solution keeps the connection costs it was built from and sums
// them root -> goal (tree.cpp:791-813, path.cpp:285-291)
__device__ double f_make_solution(const ForestDev& f, uint32_t p, FProb& pr, const FNode* nodes)
{
  uint32_t len = 0;
  for (int32_t t = pr.goal_slot; t != 0; t = nodes[t].parent)
    len++;
  if (len > f.depth_cap)
  {
    atomicOr(&f.ctr[2], kErrDepthOverflow);
    // the cost is still exact: sum root -> goal by repeated walks (rare, slow path)
    double c = 0.0;
    for (uint32_t k = 0; k < len; k++)
    {
      int32_t t = pr.goal_slot;
      for (uint32_t j = 0; j + k + 1 < len; j++)
        t = nodes[t].parent;
      c = __dadd_rn(c, nodes[t].pcost);
    }
    pr.sol_len = 0;
    return c;
  }
  uint32_t* ss = f.sol_slot + (size_t)p * f.depth_cap;
  double* sc = f.sol_cost + (size_t)p * f.depth_cap;
  uint32_t k = len;
  for (int32_t t = pr.goal_slot; t != 0;)
  {
    const FNode nd = nodes[t];
    k--;
    ss[k] = (uint32_t)t;
    sc[k] = nd.pcost;
    t = nd.parent;
  }
  double c = 0.0;
  for (uint32_t j = 0; j < len; j++)
    c = __dadd_rn(c, sc[j]);
  pr.sol_len = len;
  return c;
}

// ---- replay: Tree::rewireOnly's decisions and RRTStar::update's epilogue, one warp per problem --------------
__device__ __forceinline__ void forest_replay_problem(const ForestDev& f, uint32_t p, uint32_t lane)
{
  FProb& pr = f.prob[p];
  const int D = f.D;
  FNode* nodes = f.nodes + (size_t)p * f.cap_seg;
  const uint32_t* nidx = f.near_idx + (size_t)p * f.near_cap;
  const double* ndist = f.near_dist + (size_t)p * f.near_cap;
  const double* c2near = f.c2near + (size_t)p * f.near_cap;
  const int32_t* chi = f.chi_edge + (size_t)p * f.near_cap;
  const uint8_t* okB = f.okB + pr.edge_first;
  const uint32_t node = pr.new_slot, near_n = pr.near_n, parB = pr.parB_count;
  double cost_to_node = pr.c2n0;
  bool improved = false;
  uint32_t used = 0, miss = 0;
  // parent phase (tree.cpp:428-463)
  int winner = pr.par_winner;
  if (winner < 0)
  {
    for (uint32_t base = 0; base < parB && winner < 0; base += 32)
    {
      const uint32_t k = base + lane;
      const uint32_t m = __ballot_sync(0xffffffffu, k < parB && okB[k] != 0);
      if (m)
        winner = (int)(pr.par_known + base + (uint32_t)__ffs(m) - 1u);
    }
    used += (winner >= 0) ? (uint32_t)winner + 1u : pr.par_known + parB;
  }
  else
    used += (uint32_t)winner + 1u;
  uint32_t par = pr.closest;
  if (winner >= 0)
  {
    const uint32_t i = f.par_cand[(size_t)p * f.near_cap + winner];
    par = nidx[i];
    if (lane == 0)
    {
      FNode nd;
      nd.parent = (int32_t)par;
      nd.pcost = ndist[i];  // metrics_->cost(n, node): the exact fp64 near distance
      nd.pad = 0;
      nodes[node] = nd;
    }
    cost_to_node = f.par_total[(size_t)p * f.near_cap + winner];
    improved = true;
  }
  __syncwarp();
  // children phase (tree.cpp:471-510) in near order
  bool rewired = false;
  for (uint32_t base = 0; base < near_n; base += 32)
  {
    const uint32_t i = base + lane;
    uint32_t pending = __ballot_sync(0xffffffffu, i < near_n);
    while (pending)
    {
      bool want = false, consult = false;
      uint32_t nn = 0;
      double c = 0.0;
      if ((pending >> lane) & 1u)
      {
        nn = nidx[i];
        if (nn != par && nn != node && nn != 0)
        {
          const double c2 = rewired ? f_walk(nodes, (int32_t)nn) : c2near[i];
          c = ndist[i];
          if (!(cost_to_node >= c2) && !(__dadd_rn(cost_to_node, c) >= c2))
          {
            const int32_t e = chi[i];
            if (e < 0)
              miss++;
            else
            {
              consult = true;
              want = okB[e] != 0;
            }
          }
        }
      }
      const uint32_t m = __ballot_sync(0xffffffffu, want);
      const uint32_t cm = __ballot_sync(0xffffffffu, consult);
      if (!m)
      {
        used += __popc(cm & pending);
        break;
      }
      const uint32_t L = (uint32_t)__ffs(m) - 1u;
      const uint32_t done = (2u << L) - 1u;  // lanes up to and including the taker
      used += __popc(cm & pending & done);
      if (lane == L)
      {
        FNode nd;
        nd.parent = (int32_t)node;
        nd.pcost = c;
        nd.pad = 0;
        nodes[nn] = nd;
      }
      improved = true;
      rewired = true;  // descendants of nn just got cheaper: walk from here on
      pending &= ~done;
      __syncwarp();
    }
  }
  if (lane != 0)
    return;
  // RRTStar::update epilogue (rrt_star.cpp:110-162); Tree::rewire returns rewireOnly's `improved`
  bool changed = false;
  if (!pr.solved)
  {
    if (improved)
    {
      const double* me = f.x64 + ((size_t)p * f.cap_seg + node) * D;
      const double* goal = f.goals + (size_t)p * D;
      const double dg = f_dist(me, goal, D);
      if (dg < f.max_distance)
      {
        bool free_edge = false;
        if (pr.goal_edge >= 0)
        {
          free_edge = okB[pr.goal_edge] != 0;
          used++;
        }
        else
          miss++;
        if (free_edge)
        {
          const uint32_t gs = f.used[p];
          for (int d = 0; d < D; d++)
            f_store_conf(f, p, gs, goal, d);  // start_tree_->addNode(goal_node_)
          FNode nd;
          nd.parent = (int32_t)node;
          nd.pcost = dg;
          nd.pad = 0;
          nodes[gs] = nd;
          f.used[p] = gs + 1;
          pr.tree_size = gs + 1;
          pr.goal_slot = (int32_t)gs;
          const double pc = f_make_solution(f, p, pr, nodes);
          f.path_cost[p] = pc;
          pr.cost = __dadd_rn(pc, pr.goal_cost);
          pr.solved = 1;
          changed = true;
        }
      }
    }
  }
  else if (improved)
  {
    if (!(f_walk(nodes, pr.goal_slot) >= __dsub_rn(f.path_cost[p], 1e-8)))
    {
      const double pc = f_make_solution(f, p, pr, nodes);
      f.path_cost[p] = pc;
      pr.cost = __dadd_rn(pc, pr.goal_cost);
      changed = true;
    }
  }
  if (changed)
  {
    // the rewire radius of this problem needs libm's pow() before the next step: queue it for the host
    f.changed[p] = 1;
    const uint32_t k = atomicAdd(&f.ctr[4], 1u);
    f.chg_p[k] = p;
    atomicAdd(&f.stats[9], 1ull);
  }
  atomicAdd(&f.stats[3], (unsigned long long)used);
  if (miss)
  {
    atomicAdd(&f.stats[4], (unsigned long long)miss);
    atomicOr(&f.ctr[2], kErrSpecMiss);
  }
}

__global__ void __launch_bounds__(128) forest_replay_kernel(ForestDev f)
{
  const uint32_t p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint32_t lane = threadIdx.x & 31;
  if (p == 0 && lane == 0)
    atomicAdd(&f.stats[2], (unsigned long long)f.ctr[0] + f.ctr[1]);
  if (p < f.P && f.prob[p].extended == 1)
    forest_replay_problem(f, p, lane);
  // the last block to finish hands the problems whose path cost changed to the host (mapped pinned memory): the
  // radius thread answers with rewire_factor * pow(2(1+1/d),1/d) * pow(specific volume,1/d) from the host libm
  __syncthreads();
  __shared__ uint32_t s_last;
  if (threadIdx.x == 0)
  {
    __threadfence();
    s_last = (atomicAdd(&f.ctr[3], 1u) == gridDim.x - 1u) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last)
    return;
  __threadfence();
  const uint32_t n = *reinterpret_cast<volatile uint32_t*>(&f.ctr[4]);
  if (threadIdx.x == 0)
    f.ctr[3] = 0;
  if (n == 0)
    return;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
  {
    const uint32_t q = *reinterpret_cast<volatile uint32_t*>(&f.chg_p[i]);
    f.hs_p[i] = q;
    f.hs_cost[i] = *reinterpret_cast<volatile double*>(&f.path_cost[q]);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0)
  {
    const uint32_t req = f.ctr[5] + 1u;
    f.ctr[5] = req;
    f.hs_ctl[1] = n;
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t*>(&f.hs_ctl[0]) = req;
  }
}

// {largest tree, steps, edges} in one place for the asynchronous probe
__global__ void forest_probe_kernel(ForestDev f, uint32_t* out)
{
  if (threadIdx.x == 0)
  {
    out[0] = f.ctr[7];
    out[1] = (uint32_t)f.stats[6];
    out[2] = (uint32_t)f.stats[2];
    out[3] = (uint32_t)(f.stats[2] >> 32);
  }
}

// waits for the host's answer to the last request (if one is outstanding) and installs the new radius factors
__global__ void forest_radius_wait_kernel(ForestDev f)
{
  __shared__ uint32_t s_ok;
  const uint32_t req = f.ctr[5];
  if (f.ctr[6] == req)
    return;  // nothing was asked since the last wait
  if (threadIdx.x == 0)
  {
    const long long t0 = clock64();
    uint32_t ok = 1;
    while (*reinterpret_cast<volatile uint32_t*>(&f.hs_ctl[2]) != req)
    {
      if (clock64() - t0 > 8000000000ll)  // ~4 s: the host thread is gone
      {
        atomicOr(&f.ctr[2], kErrHostTimeout);
        ok = 0;
        break;
      }
      __nanosleep(200);
    }
    s_ok = ok;
  }
  __syncthreads();
  const uint32_t n = f.ctr[4];
  if (s_ok)
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
      f.r_rrt[f.chg_p[i]] = *reinterpret_cast<volatile double*>(&f.hs_rrt[i]);
  __syncthreads();
  if (threadIdx.x == 0)
  {
    f.ctr[4] = 0;
    f.ctr[6] = req;
  }
}
}  // namespace gc

using namespace gc;

// =================================================================================================
struct gc_forest
{
  gc_store* store = nullptr;
  gc_world* world = nullptr;
  gc_forest_config cfg{};
  int D = 0;
  uint32_t P = 0;
  ForestDev dev{};
  std::vector<void*> allocs;
  double *d_lb = nullptr, *d_ub = nullptr, *d_starts = nullptr, *d_samples = nullptr, *d_rrt = nullptr;
  uint32_t* d_probe = nullptr;
  // host side of the radius callback
  std::vector<double> lb, ub, focii;
  // radius thread: answers the device's requests through mapped pinned memory (see forest_replay_kernel)
  void* h_shared = nullptr;  // ctl[4] | p[P] | cost[P] | rrt[P]
  std::thread radius_thread;
  std::atomic<int> radius_stop{ 0 };
  double box_volume_term = 0.0;   // specific volume at infinite cost
  double r_rrt_unit = 0.0;        // rewire_factor * pow(2(1+1/d), 1/d)
  uint64_t steps = 0, steps_at_sync = 0;
  uint32_t used_at_sync = 0;
  std::atomic<uint64_t> radius_updates{ 0 };
  uint64_t per_edge_states = 2;
  double edgesB_per_step = 0.0;
  // asynchronous probe of {largest tree, edges so far, steps so far}: keeps the grid sizing of the scans and of the
  // second edge batch close to the truth between two gc_forest_sync calls without ever blocking
  uint32_t* h_probe = nullptr;  // pinned [4]: max tree size, steps (low word), edges (64 bit)
  cudaEvent_t probe_ev = nullptr;
  bool probe_pending = false;
  uint64_t probe_step = 0, probe_edges_prev = 0, probe_steps_prev = 0;
  // CUDA graph of one step (see gc_forest_step)
  bool graph_enabled = true;
  cudaGraphExec_t graph_exec = nullptr;
  uint64_t graph_sig = 0;
  unsigned graph_launches = 0;
  double graph_hintB = 0.0;
  bool snapshot_valid = false;
  std::vector<FProb> snap;
  std::vector<double> snap_path_cost;
  unsigned long long stats[16] = {};
  uint32_t err_bits = 0;
  // host sample blocks go H2D on a copy stream into one of two staging buffers while the previous iteration computes;
  // the step's own stream only waits for the block and moves it into place with a device copy
  cudaStream_t copy_stream = nullptr;
  double* d_stage[2] = { nullptr, nullptr };
  cudaEvent_t ev_ready[2] = { nullptr, nullptr }, ev_consumed[2] = { nullptr, nullptr };
  bool stage_used[2] = { false, false };
  uint32_t stage_next = 0;
};

namespace
{
// InformedSampler::setCost (informed_sampler.cpp:181-215), scale 1, then RRTStar::updateRewireRadius' r_rrt
double rrt_radius_unit_volume(const gc_forest* f, uint32_t p, double cost)
{
  const int D = f->D;
  double c = cost;
  const bool inf_cost = c >= std::numeric_limits<double>::infinity();
  double min_radius;
  if (c < f->focii[p])
  {
    c = f->focii[p];
    min_radius = 0.0;
  }
  else
    min_radius = 0.5 * std::sqrt(std::pow(c, 2) - std::pow(f->focii[p], 2));
  const double max_radius = 0.5 * c;
  double specific_volume;
  if (inf_cost)
    specific_volume = f->box_volume_term;
  else
    specific_volume = max_radius * std::pow(min_radius, D - 1);
  const double dimension = (double)D;
  return f->r_rrt_unit * std::pow(specific_volume, 1.0 / dimension);
}

// Host side of the radius handshake: spins on the request word the device writes into mapped pinned memory, evaluates
// the pow() terms with the host libm (bit-identical to the reference's, whatever variant the C library dispatches to)
// and publishes the answers.  Sleeps between polls after a stretch without requests.
void forest_radius_loop(gc_forest* f)
{
  volatile uint32_t* ctl = static_cast<volatile uint32_t*>(f->h_shared);
  const uint32_t* hp = f->dev.hs_p;
  const double* hc = f->dev.hs_cost;
  double* hr = f->dev.hs_rrt;
  uint32_t acked = 0;
  uint64_t idle = 0;
  while (!f->radius_stop.load(std::memory_order_relaxed))
  {
    const uint32_t req = ctl[0];
    if (req != acked)
    {
      std::atomic_thread_fence(std::memory_order_acquire);
      const uint32_t n = ctl[1];
      for (uint32_t i = 0; i < n && i < f->P; i++)
        hr[i] = rrt_radius_unit_volume(f, hp[i], hc[i]);
      f->radius_updates.fetch_add(n, std::memory_order_relaxed);
      std::atomic_thread_fence(std::memory_order_release);
      ctl[2] = req;
      acked = req;
      idle = 0;
    }
    else if (++idle > 200000)
      std::this_thread::sleep_for(std::chrono::microseconds(50));
  }
}

double host_dist(const double* a, const double* b, int d)
{
  double s = 0.0;
  for (int i = 0; i < d; i++)
  {
    const double t = a[i] - b[i];
    s = s + t * t;
  }
  return std::sqrt(s);
}

template <typename T>
int forest_alloc(gc_forest* f, T** out, size_t count, bool zero = false)
{
  void* p = nullptr;
  GC_CUDA(cudaMalloc(&p, sizeof(T) * (count ? count : 1)));
  f->allocs.push_back(p);
  if (zero)
    GC_CUDA(cudaMemset(p, 0, sizeof(T) * (count ? count : 1)));
  *out = static_cast<T*>(p);
  return GC_OK;
}

int forest_free(gc_forest* f)
{
  for (void* p : f->allocs)
    cudaFree(p);
  f->allocs.clear();
  if (f->radius_thread.joinable())
  {
    f->radius_stop.store(1);
    f->radius_thread.join();
  }
  if (f->h_shared)
    cudaFreeHost(f->h_shared);
  if (f->graph_exec)
    cudaGraphExecDestroy(f->graph_exec);
  if (f->probe_ev)
    cudaEventDestroy(f->probe_ev);
  if (f->h_probe)
    cudaFreeHost(f->h_probe);
  for (int b = 0; b < 2; b++)
  {
    if (f->ev_ready[b])
      cudaEventDestroy(f->ev_ready[b]);
    if (f->ev_consumed[b])
      cudaEventDestroy(f->ev_consumed[b]);
  }
  if (f->copy_stream)
    cudaStreamDestroy(f->copy_stream);
  delete f;
  return GC_OK;
}

int forest_snapshot(gc_forest* f)
{
  if (f->snapshot_valid)
    return GC_OK;
  GC_TRY(gc_forest_sync(f));
  return GC_OK;
}
}  // namespace

extern "C" {

int gc_forest_create(gc_store_t* s, gc_world_t* w, const gc_forest_config_t* cfg, const double* lb, const double* ub,
                     uint32_t P, const double* starts, const double* goals, const uint32_t* node_first,
                     const int32_t* node_parent, const double* node_pcost, const int32_t* goal_slot,
                     const uint8_t* solved, const double* path_cost, gc_forest_t** out)
{
  GC_REQUIRE(s && w && cfg && lb && ub && starts && goals && node_first && node_parent && node_pcost && goal_slot &&
                 solved && path_cost && out,
             "gc_forest_create: NULL argument");
  *out = nullptr;
  GC_REQUIRE(P >= 1 && s->nseg == P + 1, "gc_forest_create: the store needs problems + 1 segments (%u for %u problems)",
             s->nseg, P);
  GC_REQUIRE(s->dim == w->dim && s->device == w->device, "gc_forest_create: store and world do not match");
  GC_REQUIRE(cfg->max_distance > 0.0 && cfg->min_distance > 0.0 && cfg->near_cap >= 1 && cfg->near_cap <= 16384 &&
                 cfg->edges_cap >= 1 && cfg->depth_cap >= 2,
             "gc_forest_create: bad configuration");
  GC_REQUIRE(s->used[P] == 0, "gc_forest_create: the last segment must stay empty");
  GC_SET_DEVICE(s->device);
  GC_CUDA(cudaStreamSynchronize(s->stream));
  gc_forest* f = new gc_forest();
  f->store = s;
  f->world = w;
  f->cfg = *cfg;
  if (f->cfg.first_batch == 0)
    f->cfg.first_batch = 3;
  const int D = s->dim;
  f->D = D;
  f->P = P;
  f->lb.assign(lb, lb + D);
  f->ub.assign(ub, ub + D);
  f->focii.resize(P);
  const double dimension = (double)D;
  {
    double v = std::tgamma(((double)D) * 0.5 + 1.0) / std::pow(M_PI, (double)D * 0.5);
    for (int i = 0; i < D; i++)
      v *= (ub[i] - lb[i]);
    f->box_volume_term = v;
    f->r_rrt_unit = cfg->rewire_factor * std::pow(2.0 * (1.0 + 1.0 / dimension), 1.0 / dimension);
  }
  int rc = GC_OK;
  auto bail = [&](int code) {
    forest_free(f);
    return code;
  };
  ForestDev& d = f->dev;
  d.D = D;
  d.P = P;
  d.cap_seg = s->cap_seg;
  d.near_cap = cfg->near_cap;
  d.edges_cap = cfg->edges_cap;
  d.depth_cap = cfg->depth_cap;
  d.first_batch = f->cfg.first_batch;
  d.row_stride = store_row_stride(s);
  d.max_distance = cfg->max_distance;
  d.tolerance = cfg->tolerance;
  d.utopia_tol = cfg->utopia_tolerance;
  d.x32 = s->x32;
  d.x64 = s->x64;
  d.used = s->d_used;
  const size_t PN = (size_t)P * cfg->near_cap, PD = (size_t)P * D;
  const size_t nA = (size_t)P * d.first_batch;
  double *d_card = nullptr, *d_goals = nullptr;
#define FA(ptr, count, ...)                                  \
  if ((rc = forest_alloc(f, &(ptr), count, ##__VA_ARGS__)) < 0) \
    return bail(rc);
  FA(d.nodes, (size_t)P * s->cap_seg);
  FA(d.prob, P);
  FA(d.path_cost, P);
  FA(f->d_rrt, P);
  FA(d_card, (size_t)s->cap_seg + 1);
  FA(d.changed, P, true);
  FA(d_goals, PD);
  FA(f->d_starts, PD);
  FA(f->d_lb, D);
  FA(f->d_ub, D);
  FA(f->d_samples, PD);
  FA(d.seg_q, P);
  FA(d.seg_near, P);
  FA(d.radius, P);
  FA(d.closest, P);
  FA(d.dist, P);
  FA(d.c1, PD);
  FA(d.c2, PD);
  FA(d.ok_ext, P);
  FA(d.near_idx, PN);
  FA(d.near_dist, PN);
  FA(d.near_count, P);
  FA(d.c2near, PN);
  FA(d.par_cand, PN);
  FA(d.par_total, PN);
  FA(d.tmp_pos, PN);
  FA(d.tmp_total, PN);
  FA(d.chi_edge, PN);
  FA(d.eA1, nA * D);
  FA(d.eA2, nA * D);
  FA(d.okA, nA);
  FA(d.eB1, (size_t)cfg->edges_cap * D);
  FA(d.eB2, (size_t)cfg->edges_cap * D);
  FA(d.okB, cfg->edges_cap);
  FA(d.ctr, 8, true);
  FA(f->d_probe, 4, true);
  FA(d.chg_p, P);
  FA(d.stats, 16, true);
  FA(d.sol_slot, (size_t)P * cfg->depth_cap);
  FA(d.sol_cost, (size_t)P * cfg->depth_cap);
#undef FA
  d.r_rrt = f->d_rrt;
  d.card_table = d_card;
  d.goals = d_goals;
  d.samples = f->d_samples;
  const size_t shared_bytes = 64 + align_up(sizeof(uint32_t) * P, 64) + 2 * sizeof(double) * (size_t)P;
  cudaError_t e = cudaHostAlloc(&f->h_shared, shared_bytes, cudaHostAllocMapped | cudaHostAllocPortable);
  if (e == cudaSuccess)
  {
    memset(f->h_shared, 0, shared_bytes);
    void* dp = nullptr;
    e = cudaHostGetDevicePointer(&dp, f->h_shared, 0);
    char* base = static_cast<char*>(dp);
    d.hs_ctl = reinterpret_cast<uint32_t*>(base);
    d.hs_p = reinterpret_cast<uint32_t*>(base + 64);
    d.hs_cost = reinterpret_cast<double*>(base + 64 + align_up(sizeof(uint32_t) * P, 64));
    d.hs_rrt = d.hs_cost + P;
    if (e == cudaSuccess && dp != f->h_shared)
    {
      gc::set_error("gc_forest_create: mapped pinned memory is not addressable at one address on this system");
      return bail(GC_E_UNSUPPORTED);
    }
  }
  std::vector<double> h_path_cost(P), h_rrt(P);
  if (e != cudaSuccess)
  {
    gc::set_error("gc_forest_create: %s", cudaGetErrorString(e));
    return bail(GC_E_CUDA);
  }
  // host images of the initial trees
  std::vector<FProb> prob(P);
  std::vector<FNode> nodes;
  std::vector<double> card((size_t)s->cap_seg + 1);
  card[0] = 0.0;
  for (uint32_t n = 1; n <= s->cap_seg; n++)
  {
    const double cardDbl = n + 1.0;
    card[n] = std::pow(std::log(cardDbl) / cardDbl, 1.0 / dimension);
  }
  uint32_t max_used = 0;
  double max_abs = 0.0;
  for (int i = 0; i < D; i++)
    max_abs = fmax(max_abs, fmax(fabs(lb[i]), fabs(ub[i])));
  for (uint32_t p = 0; p < P; p++)
  {
    const uint32_t n = node_first[p + 1] - node_first[p];
    if (n != s->used[p] || n < 1)
    {
      gc::set_error("gc_forest_create: problem %u has %u nodes but segment %u holds %u", p, n, p, s->used[p]);
      return bail(GC_E_INVALID);
    }
    max_used = n > max_used ? n : max_used;
    FProb& pr = prob[p];
    memset(&pr, 0, sizeof(pr));
    f->focii[p] = host_dist(starts + (size_t)p * D, goals + (size_t)p * D, D);
    pr.goal_cost = 0.0;  // NullGoalCostFunction
    pr.best_utopia = pr.goal_cost + f->focii[p];
    pr.solved = solved[p];
    pr.cost = solved[p] ? path_cost[p] + pr.goal_cost : std::numeric_limits<double>::infinity();
    pr.goal_slot = goal_slot[p];
    pr.tree_size = n;
    pr.can_improve = 1;
    h_path_cost[p] = solved[p] ? path_cost[p] : std::numeric_limits<double>::infinity();
    h_rrt[p] = rrt_radius_unit_volume(f, p, h_path_cost[p]);
    for (int i = 0; i < D; i++)
      max_abs = fmax(max_abs, fmax(fabs(starts[(size_t)p * D + i]), fabs(goals[(size_t)p * D + i])));
  }
  // every later node is a sample or lies between a node and a sample: the fp32 slack bound of the scans holds
  for (uint32_t g = 0; g < s->nseg; g++)
    s->max_abs[g] = fmax(s->max_abs[g], max_abs);
  e = cudaMemcpy(d.prob, prob.data(), sizeof(FProb) * P, cudaMemcpyHostToDevice);
  for (uint32_t p = 0; p < P && e == cudaSuccess; p++)
  {
    const uint32_t n = node_first[p + 1] - node_first[p];
    nodes.resize(n);
    for (uint32_t i = 0; i < n; i++)
    {
      nodes[i].parent = node_parent[node_first[p] + i];
      nodes[i].pcost = node_pcost[node_first[p] + i];
      nodes[i].pad = 0;
    }
    e = cudaMemcpy(d.nodes + (size_t)p * s->cap_seg, nodes.data(), sizeof(FNode) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && solved[p])
    {
      // the solution the solver holds: connection costs root -> goal
      std::vector<uint32_t> ss;
      std::vector<double> sc;
      for (int32_t t = goal_slot[p]; t > 0; t = nodes[t].parent)
      {
        ss.insert(ss.begin(), (uint32_t)t);
        sc.insert(sc.begin(), nodes[t].pcost);
      }
      if (ss.size() <= cfg->depth_cap)
      {
        prob[p].sol_len = (uint32_t)ss.size();
        e = cudaMemcpy(d.sol_slot + (size_t)p * cfg->depth_cap, ss.data(), sizeof(uint32_t) * ss.size(),
                       cudaMemcpyHostToDevice);
        if (e == cudaSuccess)
          e = cudaMemcpy(d.sol_cost + (size_t)p * cfg->depth_cap, sc.data(), sizeof(double) * sc.size(),
                         cudaMemcpyHostToDevice);
      }
    }
  }
  if (e == cudaSuccess)
    e = cudaMemcpy(d.prob, prob.data(), sizeof(FProb) * P, cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(d.path_cost, h_path_cost.data(), sizeof(double) * P, cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(f->d_rrt, h_rrt.data(), sizeof(double) * P, cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(d_card, card.data(), sizeof(double) * card.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(d_goals, goals, sizeof(double) * PD, cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(f->d_starts, starts, sizeof(double) * PD, cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(f->d_lb, lb, sizeof(double) * D, cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(f->d_ub, ub, sizeof(double) * D, cudaMemcpyHostToDevice);
  if (e != cudaSuccess)
  {
    gc::set_error("gc_forest_create: %s", cudaGetErrorString(e));
    return bail(GC_E_CUDA);
  }
  // scratch of the scans and the edge checks at their final size: no reallocation inside a step
  const size_t max_edges = std::max<size_t>(std::max<size_t>(nA, cfg->edges_cap), P);
  if ((rc = s->d_part.reserve((size_t)(s->sms * 4 + P + 8) * (sizeof(double) + sizeof(uint32_t)))) < 0 ||
      (rc = w->d_plan.reserve(sizeof(double) * 2 * max_edges)) < 0 ||
      (rc = w->d_offs.reserve(sizeof(uint32_t) * (4 * (max_edges + 1) + max_edges + 1 + 8))) < 0)
    return bail(rc);
  f->used_at_sync = max_used;
  s->used_bound = max_used;
  for (double nn = 2; cfg->max_distance > nn * cfg->min_distance && f->per_edge_states < (1u << 20); nn *= 2)
    f->per_edge_states += (uint64_t)(nn / 2);
  f->edgesB_per_step = (double)P * 4.0;
  f->snap = prob;
  f->snap_path_cost = h_path_cost;
  f->snapshot_valid = true;
  f->radius_thread = std::thread(forest_radius_loop, f);
  f->graph_enabled = getenv("GCB200_NO_GRAPH") == nullptr;
  if (cudaMallocHost(&f->h_probe, 4 * sizeof(uint32_t)) != cudaSuccess ||
      cudaEventCreateWithFlags(&f->probe_ev, cudaEventDisableTiming) != cudaSuccess)
  {
    gc::set_error("gc_forest_create: probe buffers");
    return bail(GC_E_CUDA);
  }
  *out = f;
  return GC_OK;
}

int gc_forest_destroy(gc_forest_t* f)
{
  if (!f)
    return GC_OK;
  GC_SET_DEVICE(f->store->device);
  cudaStreamSynchronize(f->store->stream);
  refresh_store_mirror(f->store);
  f->store->used_bound = 0;
  return forest_free(f);
}

// the launches of one step after the samples are in place (d.samples); everything on `st`, nothing synchronises
static int forest_enqueue(gc_forest* f, const ForestDev& d, cudaStream_t st)
{
  gc_store* s = f->store;
  gc_world* w = f->world;
  const uint32_t P = f->P;
  forest_radius_wait_kernel<<<1, 128, 0, st>>>(d);
  GC_CUDA(cudaGetLastError());
  forest_prologue_kernel<<<(P + 127) / 128, 128, 0, st>>>(d);
  GC_CUDA(cudaGetLastError());
  count_launch(2);
  GC_TRY(nn1_dev(s, d.samples, d.seg_q, P, d.closest, d.dist, false));
  GC_TRY(steer_launch(s, d.samples, d.seg_q, d.closest, d.dist, P, f->cfg.max_distance, d.c1, d.c2, st));
  GC_TRY(check_edges_dev(w, d.c1, d.c2, nullptr, P, f->cfg.min_distance, GC_EDGE_BISECT, d.ok_ext, nullptr,
                         f->per_edge_states * P, st));
  forest_gate_kernel<<<(P + 255) / 256, 256, 0, st>>>(d);
  GC_CUDA(cudaGetLastError());
  count_launch();
  GC_TRY(radius_dev(s, d.c2, d.seg_near, d.radius, P, d.near_cap, d.near_idx, d.near_dist, d.near_count));
  forest_extend_kernel<<<P, 128, 0, st>>>(d);
  GC_CUDA(cudaGetLastError());
  count_launch();
  const uint32_t nA = P * d.first_batch;
  GC_TRY(check_edges_dev(w, d.eA1, d.eA2, nullptr, nA, f->cfg.min_distance, GC_EDGE_BISECT, d.okA, nullptr,
                         f->per_edge_states * std::max<uint64_t>(1, (uint64_t)(nA * 0.6)), st, d.ctr + 0));
  forest_resolve_kernel<<<P, 128, 0, st>>>(d);
  GC_CUDA(cudaGetLastError());
  count_launch();
  GC_TRY(check_edges_dev(w, d.eB1, d.eB2, nullptr, d.edges_cap, f->cfg.min_distance, GC_EDGE_BISECT, d.okB, nullptr,
                         f->per_edge_states * std::max<uint64_t>(1, (uint64_t)f->graph_hintB), st, d.ctr + 1));
  forest_replay_kernel<<<(P + 3) / 4, 128, 0, st>>>(d);
  GC_CUDA(cudaGetLastError());
  count_launch();
  return GC_OK;
}

int gc_forest_step(gc_forest_t* f, const double* samples, const double* d_samples, uint64_t seed, uint64_t first_index)
{
  GC_REQUIRE(f != nullptr, "gc_forest_step: NULL forest");
  GC_REQUIRE(!(samples && d_samples), "gc_forest_step: give host samples or device samples, not both");
  gc_store* s = f->store;
  gc_world* w = f->world;
  GC_SET_DEVICE(s->device);
  cudaStream_t st = s->stream;
  ForestDev d = f->dev;
  const uint32_t P = f->P;
  const int D = f->D;
  // The step is replayed from a CUDA graph (one launch instead of ~20, no gaps between the kernels) unless the world
  // records per-call timing events or GCB200_NO_GRAPH is set; the samples then go through the forest's own block.
  const bool use_graph = f->graph_enabled && !w->timing;
  if (samples)
  {
    const size_t bytes = sizeof(double) * (size_t)P * D;
    if (!f->copy_stream)
    {
      GC_CUDA(cudaStreamCreateWithFlags(&f->copy_stream, cudaStreamNonBlocking));
      for (int b = 0; b < 2; b++)
      {
        GC_TRY(forest_alloc(f, &f->d_stage[b], (size_t)P * D));
        GC_CUDA(cudaEventCreateWithFlags(&f->ev_ready[b], cudaEventDisableTiming));
        GC_CUDA(cudaEventCreateWithFlags(&f->ev_consumed[b], cudaEventDisableTiming));
      }
    }
    const uint32_t b = f->stage_next;
    f->stage_next ^= 1u;
    if (f->stage_used[b])  // the step that last read this staging buffer must have moved it out
      GC_CUDA(cudaStreamWaitEvent(f->copy_stream, f->ev_consumed[b], 0));
    GC_CUDA(cudaMemcpyAsync(f->d_stage[b], samples, bytes, cudaMemcpyHostToDevice, f->copy_stream));
    GC_CUDA(cudaEventRecord(f->ev_ready[b], f->copy_stream));
    GC_CUDA(cudaStreamWaitEvent(st, f->ev_ready[b], 0));
    GC_CUDA(cudaMemcpyAsync(f->d_samples, f->d_stage[b], bytes, cudaMemcpyDeviceToDevice, st));
    GC_CUDA(cudaEventRecord(f->ev_consumed[b], st));
    f->stage_used[b] = true;
  }
  else if (d_samples)
  {
    if (use_graph)
      GC_CUDA(cudaMemcpyAsync(f->d_samples, d_samples, sizeof(double) * (size_t)P * D, cudaMemcpyDeviceToDevice, st));
    else
      d.samples = d_samples;
  }
  else
    GC_TRY(sample_informed_launch(D, P, f->d_lb, f->d_ub, f->d_starts, d.goals, d.path_cost, seed, first_index,
                                  f->d_samples, st));
  if (f->probe_pending && cudaEventQuery(f->probe_ev) == cudaSuccess)
  {
    f->probe_pending = false;
    if (f->probe_step >= f->steps_at_sync)
    {
      f->used_at_sync = f->h_probe[0];
      f->steps_at_sync = f->probe_step;
    }
    const uint64_t edges = ((uint64_t)f->h_probe[3] << 32) | f->h_probe[2], nsteps = f->h_probe[1];
    if (nsteps > f->probe_steps_prev)
      f->edgesB_per_step = std::max(64.0, (double)(edges - f->probe_edges_prev) / (double)(nsteps - f->probe_steps_prev));
    f->probe_edges_prev = edges;
    f->probe_steps_prev = nsteps;
  }
  // every tree grows by at most one node per step, plus the goal once
  {
    const uint64_t bound = (uint64_t)f->used_at_sync + (f->steps - f->steps_at_sync) + 3u;
    s->used_bound = (uint32_t)std::min<uint64_t>(bound, s->cap_seg);
  }
  // grid-sizing hint of the second edge batch: sticky, so that the captured graph stays valid
  if (f->graph_hintB <= 0.0 || f->edgesB_per_step > 1.5 * f->graph_hintB || f->edgesB_per_step < 0.5 * f->graph_hintB)
    f->graph_hintB = f->edgesB_per_step;
  if (!use_graph)
    GC_TRY(forest_enqueue(f, d, st));
  else
  {
    // launch geometry depends on: the stream, the tile count of the scans, the small-segment switch, the hint
    const uint64_t sig = ((uint64_t)(s->used_bound <= 1024u ? 0u : (s->used_bound + 1023u) / 1024u) << 32) ^
                         (uint64_t)(f->graph_hintB) ^ ((uint64_t)(uintptr_t)st << 20);
    if (!f->graph_exec || sig != f->graph_sig)
    {
      if (f->graph_exec)
        cudaGraphExecDestroy(f->graph_exec);
      f->graph_exec = nullptr;
      uint64_t l0 = 0, l1 = 0;
      gc_launch_count(&l0);
      GC_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
      const int rc = forest_enqueue(f, d, st);
      cudaGraph_t g = nullptr;
      const cudaError_t e = cudaStreamEndCapture(st, &g);
      gc_launch_count(&l1);
      if (rc < 0 || e != cudaSuccess || !g)
      {
        if (g)
          cudaGraphDestroy(g);
        if (rc >= 0)
          gc::set_error("gc_forest_step: stream capture failed: %s", cudaGetErrorString(e));
        return rc < 0 ? rc : GC_E_CUDA;
      }
      const cudaError_t e2 = cudaGraphInstantiate(&f->graph_exec, g, 0);
      cudaGraphDestroy(g);
      if (e2 != cudaSuccess)
      {
        f->graph_exec = nullptr;
        gc::set_error("gc_forest_step: cudaGraphInstantiate: %s", cudaGetErrorString(e2));
        return GC_E_CUDA;
      }
      f->graph_sig = sig;
      f->graph_launches = (unsigned)(l1 - l0);
    }
    else
      count_launch(f->graph_launches);
    GC_CUDA(cudaGraphLaunch(f->graph_exec, st));
  }
  f->steps++;
  f->snapshot_valid = false;
  if (!f->probe_pending && (f->steps & 15u) == 0)
  {
    forest_probe_kernel<<<1, 32, 0, st>>>(d, f->d_probe);
    GC_CUDA(cudaGetLastError());
    count_launch();
    GC_CUDA(cudaMemcpyAsync(f->h_probe, f->d_probe, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    GC_CUDA(cudaEventRecord(f->probe_ev, st));
    f->probe_pending = true;
    f->probe_step = f->steps;
  }
  return GC_OK;
}

int gc_forest_read_costs_async(gc_forest_t* f, double* costs)
{
  GC_REQUIRE(f && costs, "gc_forest_read_costs_async: NULL argument");
  GC_SET_DEVICE(f->store->device);
  // path cost + goal cost with a null goal cost function: the path costs themselves
  GC_CUDA(cudaMemcpyAsync(costs, f->dev.path_cost, sizeof(double) * f->P, cudaMemcpyDeviceToHost, f->store->stream));
  return GC_OK;
}

int gc_forest_sync(gc_forest_t* f)
{
  GC_REQUIRE(f != nullptr, "gc_forest_sync: NULL forest");
  gc_store* s = f->store;
  GC_SET_DEVICE(s->device);
  GC_CUDA(cudaStreamSynchronize(s->stream));
  uint32_t ctr[4];
  GC_CUDA(cudaMemcpy(ctr, f->dev.ctr, sizeof(ctr), cudaMemcpyDeviceToHost));
  const unsigned long long edges_before = f->stats[2], steps_before = f->stats[6];
  GC_CUDA(cudaMemcpy(f->stats, f->dev.stats, sizeof(f->stats), cudaMemcpyDeviceToHost));
  f->snap.resize(f->P);
  GC_CUDA(cudaMemcpy(f->snap.data(), f->dev.prob, sizeof(FProb) * f->P, cudaMemcpyDeviceToHost));
  f->snap_path_cost.resize(f->P);
  GC_CUDA(cudaMemcpy(f->snap_path_cost.data(), f->dev.path_cost, sizeof(double) * f->P, cudaMemcpyDeviceToHost));
  GC_TRY(refresh_store_mirror(s));
  uint32_t mu = 0;
  for (uint32_t v : s->used)
    mu = v > mu ? v : mu;
  f->used_at_sync = mu;
  f->steps_at_sync = f->steps;
  s->used_bound = mu ? mu : 1;
  if (f->stats[6] > steps_before)
  {
    // grid sizing hint of the second edge batch: what the last stretch of steps needed (first batch excluded)
    const double per_step = (double)(f->stats[2] - edges_before) / (double)(f->stats[6] - steps_before);
    f->edgesB_per_step = std::max(64.0, per_step);
  }
  f->snapshot_valid = true;
  f->err_bits |= ctr[2];
  if (f->err_bits)
  {
    gc::set_error("gc_forest: capacity exceeded since creation (%s%s%s%s%s): the trees are no longer the reference's",
                  (f->err_bits & kErrNearOverflow) ? " near list > near_cap" : "",
                  (f->err_bits & kErrEdgeOverflow) ? " rewiring edges > edges_cap" : "",
                  (f->err_bits & kErrSegmentFull) ? " tree reached the segment capacity" : "",
                  (f->err_bits & kErrDepthOverflow) ? " solution longer than depth_cap" : "",
                  (f->err_bits & kErrSpecMiss) ? " edge outside the speculative set" : "");
    return GC_E_CAPACITY;
  }
  return GC_OK;
}

int gc_forest_stats(gc_forest_t* f, uint64_t out[16])
{
  GC_REQUIRE(f && out, "gc_forest_stats: NULL argument");
  GC_TRY(forest_snapshot(f));
  for (int i = 0; i < 16; i++)
    out[i] = f->stats[i];
  uint64_t improvable = 0, solved = 0;
  for (const FProb& pr : f->snap)
  {
    improvable += pr.can_improve ? 1 : 0;
    solved += pr.solved ? 1 : 0;
  }
  out[7] = improvable;
  out[8] = solved;
  out[10] = f->radius_updates.load();
  return GC_OK;
}

int gc_forest_info(gc_forest_t* f, uint32_t p, int32_t* solved, int32_t* can_improve, double* cost, double* radius,
                   uint32_t* num_nodes, uint32_t* tree_size)
{
  GC_REQUIRE(f && p < f->P, "gc_forest_info: bad argument");
  GC_TRY(forest_snapshot(f));
  const FProb& pr = f->snap[p];
  if (solved)
    *solved = pr.solved;
  if (can_improve)
    *can_improve = pr.can_improve;
  if (cost)
    *cost = pr.cost;
  if (radius)
    *radius = pr.r_rewire;
  if (num_nodes)
    *num_nodes = pr.tree_size + (pr.goal_slot < 0 ? 1u : 0u);
  if (tree_size)
    *tree_size = pr.tree_size;
  return GC_OK;
}

// slot -> id in the reference's creation order (0 start, 1 goal, 2.. new nodes)
static inline int32_t forest_id_of_slot(int32_t slot, int32_t goal_slot)
{
  if (slot <= 0)
    return slot;
  if (goal_slot >= 0 && slot == goal_slot)
    return 1;
  return (goal_slot < 0 || slot < goal_slot) ? slot + 1 : slot;
}

int gc_forest_dump(gc_forest_t* f, uint32_t p, int32_t* parents, double* pcost, double* conf)
{
  GC_REQUIRE(f && p < f->P && parents && pcost && conf, "gc_forest_dump: bad argument");
  GC_TRY(forest_snapshot(f));
  GC_SET_DEVICE(f->store->device);
  const FProb& pr = f->snap[p];
  const int D = f->D;
  const uint32_t n = pr.tree_size;
  std::vector<FNode> nodes(n);
  std::vector<double> rows((size_t)n * D);
  GC_CUDA(cudaMemcpy(nodes.data(), f->dev.nodes + (size_t)p * f->store->cap_seg, sizeof(FNode) * n,
                     cudaMemcpyDeviceToHost));
  GC_CUDA(cudaMemcpy(rows.data(), f->store->x64 + (size_t)p * f->store->cap_seg * D, sizeof(double) * n * D,
                     cudaMemcpyDeviceToHost));
  if (pr.goal_slot < 0)
  {
    parents[1] = -1;
    pcost[1] = 0.0;
    GC_CUDA(cudaMemcpy(conf + D, f->dev.goals + (size_t)p * D, sizeof(double) * D, cudaMemcpyDeviceToHost));
  }
  for (uint32_t sl = 0; sl < n; sl++)
  {
    const int32_t id = forest_id_of_slot((int32_t)sl, pr.goal_slot);
    parents[id] = sl == 0 ? -1 : forest_id_of_slot(nodes[sl].parent, pr.goal_slot);
    pcost[id] = sl == 0 ? 0.0 : nodes[sl].pcost;
    memcpy(conf + (size_t)id * D, &rows[(size_t)sl * D], sizeof(double) * D);
  }
  return GC_OK;
}

int gc_forest_solution(gc_forest_t* f, uint32_t p, int32_t* ids, uint32_t cap)
{
  GC_REQUIRE(f && p < f->P && (ids || cap == 0), "gc_forest_solution: bad argument");
  GC_TRY(forest_snapshot(f));
  GC_SET_DEVICE(f->store->device);
  const FProb& pr = f->snap[p];
  if (!pr.solved)
    return 0;
  std::vector<uint32_t> ss(pr.sol_len);
  if (pr.sol_len)
    GC_CUDA(cudaMemcpy(ss.data(), f->dev.sol_slot + (size_t)p * f->cfg.depth_cap, sizeof(uint32_t) * pr.sol_len,
                       cudaMemcpyDeviceToHost));
  if (cap > 0)
    ids[0] = 0;
  for (uint32_t i = 0; i < pr.sol_len && i + 1 < cap; i++)
    ids[i + 1] = forest_id_of_slot((int32_t)ss[i], pr.goal_slot);
  return (int)pr.sol_len + 1;
}

}  // extern "C"
