// anytime.cu -- lock-step AnytimeRRT (src/graph_core/solvers/anytime_rrt.cpp, BASELINE config 5): P independent
// queries that share one collision world advance together, one planner update per problem and step.
//
// What a problem does is the reference's AnytimeRRT::solve, restated as a per-problem state machine:
//   set-up      TreeSolver::setProblem (tree_solver.cpp:55-105): the direct ray start -> goal (Tree::connectToNode,
//               tree.cpp:304-327), whose nodes stay in the start tree when the ray is blocked;
//   RRT phase   up to FAILED_ITER rounds of RRT::solve = TreeSolver::solve (tree_solver.cpp:107-130): max_iter times
//               RRT::update (rrt.cpp:116-157) -- Tree::connect / Tree::extend towards a sample of the solver's own
//               sampler, then the goal connection when the new node is within max_distance of the goal;
//   improve     rounds of AnytimeRRT::improve (:269-330): cost2beat = (1 - cost_impr) * path cost, bias -= delta, a
//               fresh tree rooted at the start, an informed sampler of cost cost2beat, max_iter times improveUpdate
//               (:356-427) = Tree::informedExtend (tree.cpp:235-302) + the goal connection when the new solution is
//               cheaper; FAILED_ITER failed rounds in a row end the query (:134-171).
// NOTE (reference as written): RRT::update leaves can_improve_ = false when it connects the goal (rrt.cpp:150), so the
// improve loop of AnytimeRRT::solve (:134) is never entered from solve() itself; callers reach the rounds through the
// public improve().  cfg.improve_in_solve = 1 runs that loop with improve() called unconditionally (the anytime
// behaviour the class is named after), 0 reproduces solve() literally.
//
// Where the work runs.  Device: the samples (informed Philox sampler, one counter per problem so that a problem's
// k-th sample does not depend on what the other problems do), nearest node + steer + extension edge for the RRT rows
// (nn1_dev / steer / check_edges_dev), Tree::informedExtend for the improve rows (informed.cu: all-nodes k-nearest,
// cost-to-come walks, heuristic order, edge rounds), the goal edges, the node store and the parent / connection-cost
// arrays.  Host: the per-problem control flow above (a few branches per problem and step) and the parent / cost
// mirror from which solution costs are summed in the reference's order.  Per step: one H2D of the row lists, one D2H
// of the row results, a handful of launches over all problems.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <limits>
#include <string>
#include <vector>

#include "gc_common.cuh"

namespace gc
{
constexpr int kAnyFailedIter = 3;  // anytime_rrt.h:35 FAILED_ITER
constexpr uint8_t kAnyInit = 0, kAnyRrt = 1, kAnyImprove = 2, kAnyDone = 3;

__global__ void any_gather_kernel(const double* __restrict__ src, const uint32_t* __restrict__ rows, uint32_t n, int dim,
                                  double* __restrict__ dst)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * (uint32_t)dim)
    dst[i] = src[(size_t)rows[i / dim] * dim + i % dim];
}

// parent slot and connection cost of freshly appended nodes (global slot = segment * capacity + slot)
__global__ void any_links_kernel(const uint32_t* __restrict__ gslot, const int32_t* __restrict__ parent,
                                 const double* __restrict__ pcost, const double* __restrict__ c2n, uint32_t n,
                                 int32_t* d_parent, double* d_pcost, double* d_c2n)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
  {
    d_parent[gslot[i]] = parent[i];
    d_pcost[gslot[i]] = pcost[i];
    d_c2n[gslot[i]] = c2n[i];
  }
}

// Tree::costToNode (tree.cpp:767-789): the connection costs summed from the node up to the root
static double cost_to_node(const std::vector<int32_t>& parent, const std::vector<double>& pcost, uint32_t slot)
{
  double c = 0.0;
  for (int32_t x = (int32_t)slot; x != 0;)
  {
    const int32_t par = parent[(size_t)x];
    if (par < 0)
      return std::numeric_limits<double>::infinity();
    c += pcost[(size_t)x];
    x = par;
  }
  return c;
}

// EuclideanMetrics::cost / (a - b).norm(): sequential, individually rounded (no contraction in host code)
static double dist64(const double* a, const double* b, int d)
{
  double s = 0.0;
  for (int i = 0; i < d; i++)
  {
    const double t = a[i] - b[i];
    s += t * t;
  }
  return std::sqrt(s);
}

struct AnyProb
{
  uint8_t phase = kAnyInit;
  bool solved = false, can_improve = true;
  bool ray = false;         // a Tree::connect ray continues with the same target in the next step
  bool need_round = false;  // improve phase: the next round has to be set up before a sample is drawn
  int n_failed = 0;
  uint32_t iter = 0;        // updates of the current round
  uint64_t k = 0;           // samples drawn so far
  double bias = 0.9, cost2beat = 0.0, path_cost = 0.0, cost = 0.0, best_utopia = 0.0;
  // parent slot / connection cost of the tree that is growing (slot order = creation order, root = slot 0)
  std::vector<int32_t> parent;
  std::vector<double> pcost;
  // current solution: way points start -> goal and the connection costs (Path keeps its Connection objects)
  std::vector<double> sol_conf, sol_costs;
  // tree that holds the current solution (kept when cfg.keep_trees): the goal is its last node
  std::vector<int32_t> t_parent;
  std::vector<double> t_pcost, t_conf;
  std::vector<double> log;  // rows of 8: phase, iterations, success, path cost, tree nodes, bias, cost2beat, set up
};
}  // namespace gc

using namespace gc;

struct gc_anytime
{
  int device = 0, D = 0;
  uint32_t P = 0, cap = 0;
  gc_anytime_config_t cfg{};
  double tolerance = 1e-06, utopia_tol = 1.01, k_rrt = 0.0;
  gc_store* store = nullptr;
  gc_world* world = nullptr;
  std::vector<double> starts, goals, cost0;
  double *d_lb = nullptr, *d_ub = nullptr, *d_start = nullptr, *d_goal = nullptr, *d_samples = nullptr;
  int32_t* d_parent = nullptr;
  double* d_pcost = nullptr;
  double* d_basis = nullptr;  // informed-sampler basis of every problem (its foci never change)
  double* d_c2n = nullptr;  // Tree::costToNode of every node, summed node -> root when it was appended
  DevArena arena;
  PinBuf h_in, h_out;
  PinBuf h_links;  // staging of the appended nodes' links: its upload may still be in flight when the next step fills h_in
  std::vector<AnyProb> pr;
  std::vector<uint64_t> stream_index;  // Philox counter of sample k of problem p: k * stream_stride + stream_index[p]
  uint64_t stream_stride = 0;
  uint32_t active = 0;
  double t_phase[8] = {};  // host wall clock: control | RRT rows | improve rows | wait | decisions | appends | goals
  uint64_t stats[8] = {};  // steps, samples, RRT rows, improve rows, nodes appended, goal edges, improve edges, solved
  // per-step scratch (host)
  std::vector<uint32_t> rowsS, rowsR, rowsI, app_seg, app_gslot, goal_row;
  std::vector<int32_t> app_parent;
  std::vector<double> app_conf, app_pcost, app_c2n, goal_c1, goal_c2, goal_cost;
  std::vector<uint8_t> goal_kind, goal_ok;
  std::vector<uint32_t> app_slots, goal_new;
};

namespace gc
{
static void any_log(AnyProb& p, int phase, uint32_t iters, bool ok, double tree_nodes, bool setup)
{
  const double row[8] = { (double)phase, (double)iters, ok ? 1.0 : 0.0, p.path_cost, tree_nodes, p.bias,
                          p.cost2beat,   setup ? 1.0 : 0.0 };
  p.log.insert(p.log.end(), row, row + 8);
}

static void any_finish(gc_anytime* a, AnyProb& p)
{
  if (p.phase != kAnyDone)
  {
    p.phase = kAnyDone;
    a->active--;
  }
}

// way points / costs of the chain root -> node (+ goal) from the device store; snapshot of the tree when asked for
static int any_make_solution(gc_anytime* a, uint32_t pi, uint32_t last_slot, double cost_to_goal)
{
  AnyProb& p = a->pr[pi];
  const int D = a->D;
  gc_store* s = a->store;
  const uint32_t n = s->used[pi];
  std::vector<double> conf((size_t)n * D);
  GC_CUDA(cudaMemcpyAsync(conf.data(), s->x64 + (size_t)pi * s->cap_seg * D, sizeof(double) * (size_t)n * D,
                          cudaMemcpyDeviceToHost, s->stream));
  GC_CUDA(cudaStreamSynchronize(s->stream));
  std::vector<uint32_t> chain;
  for (int32_t x = (int32_t)last_slot; x > 0; x = p.parent[(size_t)x])
    chain.push_back((uint32_t)x);
  std::reverse(chain.begin(), chain.end());
  p.sol_conf.assign(conf.begin(), conf.begin() + D);  // root
  p.sol_costs.clear();
  for (uint32_t x : chain)
  {
    p.sol_conf.insert(p.sol_conf.end(), conf.begin() + (size_t)x * D, conf.begin() + (size_t)(x + 1) * D);
    p.sol_costs.push_back(p.pcost[x]);
  }
  const double* goal = a->goals.data() + (size_t)pi * D;
  if (cost_to_goal >= 0.0)  // the goal is a node of its own (connectGoal); < 0: last_slot IS the goal (direct ray)
  {
    p.sol_conf.insert(p.sol_conf.end(), goal, goal + D);
    p.sol_costs.push_back(cost_to_goal);
  }
  double c = 0.0;  // Path::cost (path.cpp:285-291): the stored connection costs, start -> goal
  for (double v : p.sol_costs)
    c += v;
  p.path_cost = c;
  p.cost = c;  // NullGoalCostFunction
  p.solved = true;
  if (a->cfg.keep_trees)
  {
    p.t_parent = p.parent;
    p.t_pcost = p.pcost;
    p.t_conf = conf;
    if (cost_to_goal >= 0.0)
    {
      p.t_parent.push_back((int32_t)last_slot);
      p.t_pcost.push_back(cost_to_goal);
      p.t_conf.insert(p.t_conf.end(), goal, goal + D);
    }
  }
  return GC_OK;
}

// the part of AnytimeRRT::solve between two samples of problem pi: round bookkeeping, improve() set-up.
// Returns with the problem ready to take part in the next step, or done.  truncate receives the segments whose tree
// restarts from the root.
static void any_prepare(gc_anytime* a, uint32_t pi, std::vector<uint32_t>& truncate)
{
  AnyProb& p = a->pr[pi];
  const uint32_t max_iter = a->cfg.max_iter;
  const int D = a->D;
  while (true)
  {
    if (p.phase == kAnyDone || p.phase == kAnyInit)
      return;
    if (p.phase == kAnyRrt)
    {
      if (p.ray)
        return;
      if (p.iter >= max_iter)  // RRT::solve returned false (:98-104)
      {
        any_log(p, 0, p.iter, false, (double)p.parent.size(), true);
        p.n_failed++;
        p.iter = 0;
        if (p.n_failed >= kAnyFailedIter)
        {
          any_finish(a, p);
          return;
        }
      }
      return;
    }
    // improve phase
    if (p.need_round)
    {
      if (!((a->cfg.improve_in_solve || p.can_improve) && p.n_failed < kAnyFailedIter))
      {
        any_finish(a, p);
        return;
      }
      // AnytimeRRT::improve (:260-315)
      const double cost2beat = (1.0 - a->cfg.cost_impr) * p.path_cost;
      const double utopia = dist64(a->goals.data() + (size_t)pi * D, a->starts.data() + (size_t)pi * D, D);
      p.can_improve = true;
      bool early = false;
      if (p.cost <= a->utopia_tol * utopia)
      {
        p.can_improve = false;
        early = true;
      }
      else if (cost2beat <= utopia)
        early = true;
      if (early)
      {
        any_log(p, 1, 0, false, 0.0, false);
        p.n_failed++;
        if (p.cost <= a->utopia_tol * p.best_utopia)  // :160-165
        {
          p.can_improve = false;
          any_finish(a, p);
          return;
        }
        continue;
      }
      p.cost2beat = cost2beat;
      p.bias = p.bias - a->cfg.delta;
      if (p.bias < 0.1)
        p.bias = 0.1;
      p.parent.assign(1, -1);
      p.pcost.assign(1, 0.0);
      truncate.push_back(pi);
      p.iter = 0;
      p.need_round = false;
    }
    if (p.iter >= max_iter)  // improve() returned false (:316-329)
    {
      any_log(p, 1, p.iter, false, (double)p.parent.size(), true);
      p.n_failed++;
      p.need_round = true;
      if (p.cost <= a->utopia_tol * p.best_utopia)
      {
        p.can_improve = false;
        any_finish(a, p);
        return;
      }
      continue;
    }
    return;
  }
}

// after the RRT phase found a solution (AnytimeRRT::solve :106-131)
static void any_after_first_solution(gc_anytime* a, AnyProb& p)
{
  p.ray = false;
  if (p.cost <= a->utopia_tol * p.best_utopia)
  {
    p.can_improve = false;
    any_finish(a, p);
    return;
  }
  p.n_failed = 0;
  p.phase = kAnyImprove;
  p.need_round = true;
}

static int any_step(gc_anytime* a)
{
  gc_store* s = a->store;
  gc_world* w = a->world;
  const int D = a->D;
  const uint32_t P = a->P;
  cudaStream_t st = s->stream;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto secs = [](std::chrono::steady_clock::time_point a0, std::chrono::steady_clock::time_point a1) {
    return std::chrono::duration<double>(a1 - a0).count();
  };
  auto t0 = now();
  // ---- control: bring every live problem to its next sample
  std::vector<uint32_t> truncate;
  a->rowsS.clear();
  a->rowsR.clear();
  a->rowsI.clear();
  for (uint32_t pi = 0; pi < P; pi++)
  {
    AnyProb& p = a->pr[pi];
    if (p.phase == kAnyDone)
      continue;
    any_prepare(a, pi, truncate);
    if (p.phase == kAnyDone)
      continue;
    if (p.phase == kAnyImprove)
    {
      a->rowsS.push_back(pi);
      a->rowsI.push_back(pi);
    }
    else
    {
      if (!p.ray)
        a->rowsS.push_back(pi);
      a->rowsR.push_back(pi);
    }
  }
  if (!truncate.empty())
  {
    for (uint32_t pi : truncate)
    {
      s->used[pi] = 1;  // the root (slot 0) holds the start configuration in every round
      s->live[pi] = 1;
    }
    GC_CUDA(cudaMemcpyAsync(s->d_used, s->used.data(), sizeof(uint32_t) * s->nseg, cudaMemcpyHostToDevice, st));
  }
  const uint32_t nS = (uint32_t)a->rowsS.size(), nR = (uint32_t)a->rowsR.size(), nI = (uint32_t)a->rowsI.size();
  if (nR + nI == 0)
    return GC_OK;
  a->stats[0]++;
  a->stats[1] += nS;
  a->stats[2] += nR;
  a->stats[3] += nI;
  a->arena.reset();
  // ---- H2D: row lists and per-row parameters, one staging block
  //   u32 rowsS[nS] | u32 rowsR[nR] | u32 rowsI[nI] | (pad) u64 ctr[nS] | f64 cost[nS] | f64 c2b[nI] | f64 bias[nI]
  const size_t off_R = sizeof(uint32_t) * nS, off_I = off_R + sizeof(uint32_t) * nR;
  const size_t off_ctr = align_up(off_I + sizeof(uint32_t) * nI, 16), off_cost = off_ctr + sizeof(uint64_t) * nS;
  const size_t off_c2b = off_cost + sizeof(double) * nS, off_bias = off_c2b + sizeof(double) * nI;
  const size_t in_bytes = off_bias + sizeof(double) * nI;
  GC_TRY(a->h_in.reserve(in_bytes));
  char* hin = static_cast<char*>(a->h_in.p);
  {
    uint32_t* r = reinterpret_cast<uint32_t*>(hin);
    std::copy(a->rowsS.begin(), a->rowsS.end(), r);
    std::copy(a->rowsR.begin(), a->rowsR.end(), r + nS);
    std::copy(a->rowsI.begin(), a->rowsI.end(), r + nS + nR);
    uint64_t* ctr = reinterpret_cast<uint64_t*>(hin + off_ctr);
    double* cost = reinterpret_cast<double*>(hin + off_cost);
    for (uint32_t i = 0; i < nS; i++)
    {
      AnyProb& p = a->pr[a->rowsS[i]];
      ctr[i] = p.k * a->stream_stride + a->stream_index[a->rowsS[i]];  // sample k of problem p: counter k * P + p
      p.k++;
      cost[i] = (p.phase == kAnyImprove) ? p.cost2beat : a->cost0[a->rowsS[i]];
    }
    double* c2b = reinterpret_cast<double*>(hin + off_c2b);
    double* bias = reinterpret_cast<double*>(hin + off_bias);
    for (uint32_t i = 0; i < nI; i++)
    {
      c2b[i] = a->pr[a->rowsI[i]].cost2beat;
      bias[i] = a->pr[a->rowsI[i]].bias;
    }
  }
  char* d_in;
  GC_CUDA(a->arena.get(&d_in, in_bytes));
  GC_CUDA(cudaMemcpyAsync(d_in, hin, in_bytes, cudaMemcpyHostToDevice, st));
  const uint32_t* d_rowsS = reinterpret_cast<const uint32_t*>(d_in);
  const uint32_t* d_rowsR = d_rowsS + nS;
  const uint32_t* d_rowsI = d_rowsR + nR;
  // ---- samples (rows of d_samples; a continuing ray keeps its row)
  if (nS)
    GC_TRY(sample_informed_rows_launch(D, nS, a->d_lb, a->d_ub, a->d_start, a->d_goal,
                                       reinterpret_cast<const double*>(d_in + off_cost), a->cfg.seed, d_rowsS,
                                       reinterpret_cast<const unsigned long long*>(d_in + off_ctr), a->d_samples, st,
                                       a->d_basis));
  // ---- D2H staging: R: q[nR][D] | next[nR][D] | near[nR][D] | dist[nR] | closest[nR] | ok[nR]
  //                   I: new[nI][D] | node conf -> not needed | ecost[nI] | node[nI] | ok[nI]
  const size_t o_q = 0, o_next = o_q + sizeof(double) * (size_t)nR * D, o_near = o_next + sizeof(double) * (size_t)nR * D;
  const size_t o_dist = o_near + sizeof(double) * (size_t)nR * D, o_new = o_dist + sizeof(double) * nR;
  const size_t o_ecost = o_new + sizeof(double) * (size_t)nI * D, o_closest = o_ecost + sizeof(double) * nI;
  const size_t o_node = o_closest + sizeof(uint32_t) * nR, o_okR = o_node + sizeof(uint32_t) * nI, o_okI = o_okR + nR;
  GC_TRY(a->h_out.reserve(o_okI + nI + 16));
  char* hout = static_cast<char*>(a->h_out.p);
  auto t1 = now();
  a->t_phase[0] += secs(t0, t1);
  if (nR)
  {
    double *d_q, *d_c1, *d_c2, *d_dist;
    uint32_t* d_closest;
    uint8_t* d_ok;
    GC_CUDA(a->arena.get(&d_q, (size_t)nR * D));
    GC_CUDA(a->arena.get(&d_c1, (size_t)nR * D));
    GC_CUDA(a->arena.get(&d_c2, (size_t)nR * D));
    GC_CUDA(a->arena.get(&d_dist, nR));
    GC_CUDA(a->arena.get(&d_closest, nR));
    GC_CUDA(a->arena.get(&d_ok, nR));
    any_gather_kernel<<<(nR * (uint32_t)D + 255) / 256, 256, 0, st>>>(a->d_samples, d_rowsR, nR, D, d_q);
    GC_CUDA(cudaGetLastError());
    count_launch();
    GC_TRY(nn1_dev(s, d_q, d_rowsR, nR, d_closest, d_dist, false));  // segment of problem p = p
    GC_TRY(steer_launch(s, d_q, d_rowsR, d_closest, d_dist, nR, a->cfg.max_distance, d_c1, d_c2, st));
    uint64_t per_edge = 2;
    for (double nn = 2; a->cfg.max_distance > nn * a->cfg.min_distance && per_edge < (1u << 20); nn *= 2)
      per_edge += (uint64_t)(nn / 2);
    GC_TRY(check_edges_dev(w, d_c1, d_c2, nullptr, nR, a->cfg.min_distance, GC_EDGE_BISECT, d_ok, nullptr,
                           per_edge * nR, st));
    GC_CUDA(cudaMemcpyAsync(hout + o_q, d_q, sizeof(double) * (size_t)nR * D, cudaMemcpyDeviceToHost, st));
    GC_CUDA(cudaMemcpyAsync(hout + o_next, d_c2, sizeof(double) * (size_t)nR * D, cudaMemcpyDeviceToHost, st));
    GC_CUDA(cudaMemcpyAsync(hout + o_near, d_c1, sizeof(double) * (size_t)nR * D, cudaMemcpyDeviceToHost, st));
    GC_CUDA(cudaMemcpyAsync(hout + o_dist, d_dist, sizeof(double) * nR, cudaMemcpyDeviceToHost, st));
    GC_CUDA(cudaMemcpyAsync(hout + o_closest, d_closest, sizeof(uint32_t) * nR, cudaMemcpyDeviceToHost, st));
    GC_CUDA(cudaMemcpyAsync(hout + o_okR, d_ok, nR, cudaMemcpyDeviceToHost, st));
  }
  auto t2 = now();
  a->t_phase[1] += secs(t1, t2);
  if (nI)
  {
    double *d_q, *d_g;
    GC_CUDA(a->arena.get(&d_q, (size_t)nI * D));
    GC_CUDA(a->arena.get(&d_g, (size_t)nI * D));
    any_gather_kernel<<<(nI * (uint32_t)D + 255) / 256, 256, 0, st>>>(a->d_samples, d_rowsI, nI, D, d_q);
    any_gather_kernel<<<(nI * (uint32_t)D + 255) / 256, 256, 0, st>>>(a->d_goal, d_rowsI, nI, D, d_g);
    GC_CUDA(cudaGetLastError());
    count_launch(2);
    IeOut o{};
    uint64_t edges = 0;
    GC_TRY(informed_extend_core(s, w, d_q, d_rowsI, a->rowsI.data(), d_g, reinterpret_cast<const double*>(d_in + off_c2b),
                                reinterpret_cast<const double*>(d_in + off_bias), nI, a->k_rrt, a->cfg.max_distance,
                                a->cfg.min_distance, a->tolerance, a->d_parent, a->d_pcost, a->arena, &o, &edges,
                                a->d_c2n));
    a->stats[6] += edges;
    GC_CUDA(cudaMemcpyAsync(hout + o_new, o.d_new, sizeof(double) * (size_t)nI * D, cudaMemcpyDeviceToHost, st));
    GC_CUDA(cudaMemcpyAsync(hout + o_ecost, o.d_ecost, sizeof(double) * nI, cudaMemcpyDeviceToHost, st));
    GC_CUDA(cudaMemcpyAsync(hout + o_node, o.d_node, sizeof(uint32_t) * nI, cudaMemcpyDeviceToHost, st));
    GC_CUDA(cudaMemcpyAsync(hout + o_okI, o.d_ok, nI, cudaMemcpyDeviceToHost, st));
  }
  auto t3 = now();
  a->t_phase[2] += secs(t2, t3);
  GC_CUDA(cudaStreamSynchronize(st));
  auto t4 = now();
  a->t_phase[3] += secs(t3, t4);
  // ---- host: the decisions of RRT::update / improveUpdate on the row results
  const double* h_q = reinterpret_cast<const double*>(hout + o_q);
  const double* h_next = reinterpret_cast<const double*>(hout + o_next);
  const double* h_near = reinterpret_cast<const double*>(hout + o_near);
  const double* h_dist = reinterpret_cast<const double*>(hout + o_dist);
  const double* h_new = reinterpret_cast<const double*>(hout + o_new);
  const uint32_t* h_closest = reinterpret_cast<const uint32_t*>(hout + o_closest);
  const uint32_t* h_node = reinterpret_cast<const uint32_t*>(hout + o_node);
  const uint8_t* h_okR = reinterpret_cast<const uint8_t*>(hout + o_okR);
  const uint8_t* h_okI = reinterpret_cast<const uint8_t*>(hout + o_okI);
  a->app_seg.clear();
  a->app_conf.clear();
  a->app_parent.clear();
  a->app_pcost.clear();
  a->app_c2n.clear();
  a->goal_row.clear();
  a->goal_c1.clear();
  a->goal_c2.clear();
  a->goal_cost.clear();
  a->goal_kind.clear();
  a->goal_new.clear();
  const double max_distance = a->cfg.max_distance;
  std::vector<double> nodeconf(D);
  for (uint32_t i = 0; i < nR; i++)
  {
    const uint32_t pi = a->rowsR[i];
    AnyProb& p = a->pr[pi];
    const double* goal = a->goals.data() + (size_t)pi * D;
    const bool init = (p.phase == kAnyInit);
    const bool ok = (h_closest[i] != GC_NO_NODE) && ((h_dist[i] < a->tolerance) ? true : (h_okR[i] != 0));  // tree.cpp:69-81
    if (!ok)
    {
      // the extension failed: Tree::connect / connectToNode / extend return false
      p.ray = false;
      if (init)
      {
        p.phase = kAnyRrt;  // setProblem found no direct connection; the ray's nodes stay in the tree
        p.path_cost = std::numeric_limits<double>::infinity();
        p.cost = p.path_cost;
      }
      else
        p.iter++;
      continue;
    }
    const double* q = h_q + (size_t)i * D;
    const double* next = h_next + (size_t)i * D;
    const double* nearc = h_near + (size_t)i * D;
    const double* newc = next;
    // Tree::extendToNode (tree.cpp:194-215): within TOLERANCE of the goal the goal node itself joins the tree
    if (init && dist64(next, goal, D) < a->tolerance)
      newc = goal;
    const uint32_t slot = (uint32_t)p.parent.size();
    if (slot >= a->cap)
    {
      gc::set_error("gc_anytime: tree of problem %u reached the capacity (%u nodes)", pi, a->cap);
      return GC_E_CAPACITY;
    }
    const double cost = dist64(nearc, newc, D);  // metrics_->cost(closest_node, new_node)
    p.parent.push_back((int32_t)h_closest[i]);
    p.pcost.push_back(cost);
    a->app_seg.push_back(pi);
    a->app_conf.insert(a->app_conf.end(), newc, newc + D);
    a->app_parent.push_back((int32_t)h_closest[i]);
    a->app_pcost.push_back(cost);
    a->app_c2n.push_back(0.0);  // only the improve rounds rank by cost-to-come
    const bool reached = dist64(newc, q, D) < a->tolerance;
    if (init)
    {
      if (reached)  // connectToNode returned true: the direct connection is the solution
      {
        p.ray = false;
        a->goal_row.push_back(pi);
        a->goal_kind.push_back(2);
        a->goal_new.push_back(slot);
        a->goal_cost.push_back(-1.0);
      }
      else
        p.ray = true;
      continue;
    }
    const bool update_done = a->cfg.extend ? true : reached;
    if (!update_done)
    {
      p.ray = true;
      continue;
    }
    p.ray = false;
    p.iter++;
    // RRT::update (rrt.cpp:137-153): goal within max_distance of the new node -> check the connection
    const double dg = dist64(newc, goal, D);
    if (dg <= max_distance)
    {
      a->goal_row.push_back(pi);
      a->goal_kind.push_back(0);
      a->goal_new.push_back(slot);
      a->goal_cost.push_back(dg);
      a->goal_c1.insert(a->goal_c1.end(), newc, newc + D);
      a->goal_c2.insert(a->goal_c2.end(), goal, goal + D);
    }
  }
  for (uint32_t i = 0; i < nI; i++)
  {
    const uint32_t pi = a->rowsI[i];
    AnyProb& p = a->pr[pi];
    p.iter++;
    if (!h_okI[i])
      continue;
    const double* goal = a->goals.data() + (size_t)pi * D;
    const double* newc = h_new + (size_t)i * D;
    const uint32_t slot = (uint32_t)p.parent.size();
    if (slot >= a->cap)
    {
      gc::set_error("gc_anytime: tree of problem %u reached the capacity (%u nodes)", pi, a->cap);
      return GC_E_CAPACITY;
    }
    const double cost = reinterpret_cast<const double*>(hout + o_ecost)[i];  // |node - new| (extendOnly, tree.cpp:132)
    p.parent.push_back((int32_t)h_node[i]);
    p.pcost.push_back(cost);
    a->app_seg.push_back(pi);
    a->app_conf.insert(a->app_conf.end(), newc, newc + D);
    a->app_parent.push_back((int32_t)h_node[i]);
    a->app_pcost.push_back(cost);
    a->app_c2n.push_back(cost_to_node(p.parent, p.pcost, slot));
    // improveUpdate (:370-388)
    const double dg = dist64(newc, goal, D);
    if (dg <= max_distance)
    {
      // getConnectionToNode: the chain root -> new node; the sum starts from the goal connection
      std::vector<double> chain;
      for (int32_t x = (int32_t)slot; x > 0; x = p.parent[(size_t)x])
        chain.push_back(p.pcost[(size_t)x]);
      double new_cost = dg;
      for (size_t j = chain.size(); j-- > 0;)
        new_cost += chain[j];
      double cur = 0.0;  // solution_->cost()
      for (double v : p.sol_costs)
        cur += v;
      if (new_cost < cur)
      {
        a->goal_row.push_back(pi);
        a->goal_kind.push_back(1);
        a->goal_new.push_back(slot);
        a->goal_cost.push_back(dg);
        a->goal_c1.insert(a->goal_c1.end(), newc, newc + D);
        a->goal_c2.insert(a->goal_c2.end(), goal, goal + D);
      }
    }
  }
  auto t5 = now();
  a->t_phase[4] += secs(t4, t5);
  // ---- appends: node store + parent / cost arrays
  const uint32_t nA = (uint32_t)a->app_seg.size();
  if (nA)
  {
    a->app_slots.resize(nA);
    GC_TRY(gc_store_append_multi(s, a->app_seg.data(), a->app_conf.data(), nA, a->app_slots.data()));
    const size_t lb = align_up(sizeof(uint32_t) * nA, 16), lp = align_up(sizeof(int32_t) * nA, 16);
    const size_t lc = sizeof(double) * nA, lbytes = lb + lp + 2 * lc;
    GC_CUDA(cudaStreamSynchronize(st));  // the previous step's upload from h_links has completed
    GC_TRY(a->h_links.reserve(lbytes));
    char* h = static_cast<char*>(a->h_links.p);
    uint32_t* g = reinterpret_cast<uint32_t*>(h);
    for (uint32_t i = 0; i < nA; i++)
      g[i] = a->app_seg[i] * s->cap_seg + a->app_slots[i];
    memcpy(h + lb, a->app_parent.data(), sizeof(int32_t) * nA);
    memcpy(h + lb + lp, a->app_pcost.data(), lc);
    memcpy(h + lb + lp + lc, a->app_c2n.data(), lc);
    char* d_l;
    GC_CUDA(a->arena.get(&d_l, lbytes));
    GC_CUDA(cudaMemcpyAsync(d_l, h, lbytes, cudaMemcpyHostToDevice, st));
    any_links_kernel<<<(nA + 255) / 256, 256, 0, st>>>(
        reinterpret_cast<const uint32_t*>(d_l), reinterpret_cast<const int32_t*>(d_l + lb),
        reinterpret_cast<const double*>(d_l + lb + lp), reinterpret_cast<const double*>(d_l + lb + lp + lc), nA,
        a->d_parent, a->d_pcost, a->d_c2n);
    GC_CUDA(cudaGetLastError());
    count_launch();
    a->stats[4] += nA;
  }
  auto t6 = now();
  a->t_phase[5] += secs(t5, t6);
  // ---- goal connections
  const uint32_t nG = (uint32_t)a->goal_row.size();
  if (nG)
  {
    const uint32_t nE = (uint32_t)(a->goal_c1.size() / D);
    a->goal_ok.assign(nE ? nE : 1, 0);
    if (nE)
    {
      GC_TRY(gc_check_edges(w, a->goal_c1.data(), a->goal_c2.data(), nE, a->cfg.min_distance, GC_EDGE_BISECT,
                            a->goal_ok.data(), nullptr));
      a->stats[5] += nE;
    }
    uint32_t e = 0;
    for (uint32_t i = 0; i < nG; i++)
    {
      const uint32_t pi = a->goal_row[i];
      AnyProb& p = a->pr[pi];
      const uint8_t kind = a->goal_kind[i];
      if (kind == 2)  // direct connection found by setProblem
      {
        GC_TRY(any_make_solution(a, pi, a->goal_new[i], -1.0));
        p.phase = kAnyRrt;
        a->stats[7]++;
        any_after_first_solution(a, p);
        continue;
      }
      const bool free_edge = a->goal_ok[e++] != 0;
      if (!free_edge)
        continue;
      GC_TRY(any_make_solution(a, pi, a->goal_new[i], a->goal_cost[i]));
      if (kind == 0)
      {
        p.can_improve = false;  // rrt.cpp:150
        any_log(p, 0, p.iter, true, (double)p.parent.size() + 1.0, true);
        a->stats[7]++;
        any_after_first_solution(a, p);
      }
      else
      {
        // improveUpdate :392-417: the round's tree becomes the start tree
        any_log(p, 1, p.iter, true, (double)p.parent.size() + 1.0, true);
        p.n_failed = 0;
        p.need_round = true;
        if (p.cost <= a->utopia_tol * p.best_utopia)
        {
          p.can_improve = false;
          any_finish(a, p);
        }
      }
    }
  }
  a->t_phase[6] += secs(t6, now());
  return GC_OK;
}
}  // namespace gc

extern "C" {

int gc_anytime_destroy(gc_anytime_t* a)
{
  if (!a)
    return GC_OK;
  GC_SET_DEVICE(a->device);
  if (a->store)
    cudaStreamSynchronize(a->store->stream);
  for (void* p : { (void*)a->d_lb, (void*)a->d_start, (void*)a->d_goal, (void*)a->d_samples, (void*)a->d_parent,
                   (void*)a->d_pcost, (void*)a->d_c2n, (void*)a->d_basis })
    if (p)
      cudaFree(p);
  a->h_in.release();
  a->h_out.release();
  a->h_links.release();
  if (a->world)
    gc_world_destroy(a->world);
  if (a->store)
    gc_store_destroy(a->store);
  delete a;
  return GC_OK;
}

int gc_anytime_create(gc_world_t* w, const gc_anytime_config_t* cfg, const double* lb, const double* ub,
                      uint32_t n_problems, const double* starts, const double* goals, const double* sampler_cost0,
                      gc_anytime_t** out)
{
  GC_REQUIRE(w && cfg && lb && ub && starts && goals && out, "gc_anytime_create: NULL argument");
  *out = nullptr;
  GC_REQUIRE(n_problems > 0, "gc_anytime_create: no problems");
  GC_REQUIRE(cfg->max_distance > 0.0 && cfg->min_distance > 0.0 && cfg->max_iter > 0,
             "gc_anytime_create: distances and max_iter must be positive");
  GC_REQUIRE(cfg->capacity >= 2 && cfg->capacity <= 8192,
             "gc_anytime_create: capacity %u: Tree::informedExtend sorts at most 8192 nodes per tree", cfg->capacity);
  GC_SET_DEVICE(w->device);
  gc_anytime* a = new gc_anytime();
  a->device = w->device;
  a->D = w->dim;
  a->P = n_problems;
  a->cfg = *cfg;
  a->tolerance = cfg->tolerance > 0.0 ? cfg->tolerance : 1e-06;
  // tree_solver.cpp:40-46: negative tolerances count as 0, then 1 is added
  a->utopia_tol = 1.0 + (cfg->utopia_tolerance > 0.0 ? cfg->utopia_tolerance : 0.0);
  const int D = a->D;
  // Tree::Tree (tree.cpp:52-53)
  a->k_rrt = 1.1 * std::pow(2.0, (double)D + 1.0) * std::exp(1.0) * (1.0 + 1.0 / (double)D);
  int rc = gc_store_create(D, cfg->capacity, n_problems, w->device, &a->store);
  if (rc == GC_OK)
    rc = gc_world_clone(w, &a->world);
  if (rc != GC_OK)
  {
    gc_anytime_destroy(a);
    return rc;
  }
  a->cap = a->store->cap_seg;
  gc_world_set_stream(a->world, a->store->stream);
  cudaStream_t st = a->store->stream;
  const size_t nd = (size_t)n_problems * D;
  a->starts.assign(starts, starts + nd);
  a->goals.assign(goals, goals + nd);
  if (sampler_cost0)
    a->cost0.assign(sampler_cost0, sampler_cost0 + n_problems);
  else
    a->cost0.assign(n_problems, std::numeric_limits<double>::infinity());
  auto fail = [&](cudaError_t e, const char* what) {
    gc::set_error("gc_anytime_create: %s: %s", what, cudaGetErrorString(e));
    gc_anytime_destroy(a);
    return GC_E_CUDA;
  };
  cudaError_t e;
  if ((e = cudaMalloc(&a->d_lb, sizeof(double) * 2 * D)) != cudaSuccess)
    return fail(e, "bounds");
  a->d_ub = a->d_lb + D;
  if ((e = cudaMalloc(&a->d_start, sizeof(double) * nd)) != cudaSuccess)
    return fail(e, "starts");
  if ((e = cudaMalloc(&a->d_goal, sizeof(double) * nd)) != cudaSuccess)
    return fail(e, "goals");
  if ((e = cudaMalloc(&a->d_samples, sizeof(double) * nd)) != cudaSuccess)
    return fail(e, "samples");
  if ((e = cudaMalloc(&a->d_parent, sizeof(int32_t) * a->store->cap_total)) != cudaSuccess)
    return fail(e, "parent slots");
  if ((e = cudaMalloc(&a->d_pcost, sizeof(double) * a->store->cap_total)) != cudaSuccess)
    return fail(e, "connection costs");
  if ((e = cudaMalloc(&a->d_c2n, sizeof(double) * a->store->cap_total)) != cudaSuccess)
    return fail(e, "costs to come");
  cudaMemsetAsync(a->d_c2n, 0, sizeof(double) * a->store->cap_total, st);  // roots: cost-to-come 0
  cudaMemcpyAsync(a->d_lb, lb, sizeof(double) * D, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(a->d_ub, ub, sizeof(double) * D, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(a->d_start, starts, sizeof(double) * nd, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(a->d_goal, goals, sizeof(double) * nd, cudaMemcpyHostToDevice, st);
  // the set-up ray aims at the goal: the "sample" of every problem is its goal configuration
  cudaMemcpyAsync(a->d_samples, goals, sizeof(double) * nd, cudaMemcpyHostToDevice, st);
  cudaMemsetAsync(a->d_parent, 0xFF, sizeof(int32_t) * a->store->cap_total, st);  // -1: root / not connected
  cudaMemsetAsync(a->d_pcost, 0, sizeof(double) * a->store->cap_total, st);
  if ((e = cudaMalloc(&a->d_basis, sizeof(double) * (size_t)n_problems * (D * D + D + 1))) != cudaSuccess)
    return fail(e, "sampler bases");
  if (informed_basis_launch(D, n_problems, a->d_start, a->d_goal, a->d_basis, st) != GC_OK)
    return fail(cudaGetLastError(), "sampler bases");
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess)
    return fail(e, "uploads");
  // roots: RRT::addStart (rrt.cpp:58-78)
  std::vector<uint32_t> segs(n_problems);
  for (uint32_t i = 0; i < n_problems; i++)
    segs[i] = i;
  rc = gc_store_append_multi(a->store, segs.data(), starts, n_problems, nullptr);
  if (rc != GC_OK)
  {
    gc_anytime_destroy(a);
    return rc;
  }
  a->pr.resize(n_problems);
  a->stream_stride = n_problems;
  a->stream_index.resize(n_problems);
  for (uint32_t i = 0; i < n_problems; i++)
    a->stream_index[i] = i;
  a->active = n_problems;
  for (uint32_t i = 0; i < n_problems; i++)
  {
    AnyProb& p = a->pr[i];
    p.phase = kAnyInit;
    p.ray = true;
    p.parent.assign(1, -1);
    p.pcost.assign(1, 0.0);
    p.best_utopia = dist64(starts + (size_t)i * D, goals + (size_t)i * D, D);  // tree_solver.cpp:66-67
    p.path_cost = std::numeric_limits<double>::infinity();
    p.cost = p.path_cost;
  }
  // TreeSolver::setProblem: the direct rays, all problems in lock step, until every ray ended
  while (true)
  {
    bool any_init = false;
    for (const AnyProb& p : a->pr)
      any_init |= (p.phase == kAnyInit);
    if (!any_init)
      break;
    // only the set-up rows take part: park the others
    rc = [&]() {
      std::vector<uint8_t> saved(n_problems);
      for (uint32_t i = 0; i < n_problems; i++)
      {
        saved[i] = a->pr[i].phase;
        if (a->pr[i].phase != kAnyInit && a->pr[i].phase != kAnyDone)
          a->pr[i].phase = kAnyDone;
      }
      const int r = any_step(a);
      for (uint32_t i = 0; i < n_problems; i++)
        if (saved[i] != kAnyInit && saved[i] != kAnyDone)
          a->pr[i].phase = saved[i];
      return r;
    }();
    if (rc != GC_OK)
    {
      gc_anytime_destroy(a);
      return rc;
    }
  }
  for (uint64_t& v : a->stats)
    v = 0;
  *out = a;
  return GC_OK;
}

int gc_anytime_set_streams(gc_anytime_t* a, uint64_t stride, const uint64_t* index)
{
  GC_REQUIRE(a != nullptr && stride > 0, "gc_anytime_set_streams: NULL handle or zero stride");
  for (const AnyProb& p : a->pr)
    GC_REQUIRE(p.k == 0, "gc_anytime_set_streams: samples were already drawn");
  a->stream_stride = stride;
  for (uint32_t i = 0; i < a->P; i++)
    a->stream_index[i] = index ? index[i] : i;
  return GC_OK;
}

int gc_anytime_step(gc_anytime_t* a)
{
  GC_REQUIRE(a != nullptr, "gc_anytime_step: NULL handle");
  GC_SET_DEVICE(a->device);
  GC_TRY(any_step(a));
  return (int)std::min<uint32_t>(a->active, 0x7fffffffu);
}

int gc_anytime_run(gc_anytime_t* a, uint64_t max_steps)
{
  GC_REQUIRE(a != nullptr, "gc_anytime_run: NULL handle");
  GC_SET_DEVICE(a->device);
  for (uint64_t i = 0; (max_steps == 0 || i < max_steps) && a->active > 0; i++)
    GC_TRY(any_step(a));
  return (int)std::min<uint32_t>(a->active, 0x7fffffffu);
}

int gc_anytime_stats(gc_anytime_t* a, uint64_t out[8])
{
  GC_REQUIRE(a && out, "gc_anytime_stats: NULL argument");
  for (int i = 0; i < 8; i++)
    out[i] = a->stats[i];
  return GC_OK;
}

int gc_anytime_timing(gc_anytime_t* a, double out[8])
{
  GC_REQUIRE(a && out, "gc_anytime_timing: NULL argument");
  for (int i = 0; i < 8; i++)
    out[i] = a->t_phase[i];
  return GC_OK;
}

int gc_anytime_info(gc_anytime_t* a, uint32_t p, int32_t* solved, int32_t* done, double* cost, double* bias,
                    double* cost2beat, uint64_t* samples, uint32_t* tree_nodes, uint32_t* log_rows)
{
  GC_REQUIRE(a && p < a->P, "gc_anytime_info: bad handle or problem index");
  const AnyProb& q = a->pr[p];
  if (solved)
    *solved = q.solved ? 1 : 0;
  if (done)
    *done = (q.phase == kAnyDone) ? 1 : 0;
  if (cost)
    *cost = q.cost;
  if (bias)
    *bias = q.bias;
  if (cost2beat)
    *cost2beat = q.cost2beat;
  if (samples)
    *samples = q.k;
  if (tree_nodes)
    *tree_nodes = (uint32_t)q.parent.size();
  if (log_rows)
    *log_rows = (uint32_t)(q.log.size() / 8);
  return GC_OK;
}

int gc_anytime_read_solved(gc_anytime_t* a, uint8_t* solved, double* cost)
{
  GC_REQUIRE(a && solved && cost, "gc_anytime_read_solved: NULL argument");
  for (uint32_t i = 0; i < a->P; i++)
  {
    solved[i] = a->pr[i].solved ? 1 : 0;
    cost[i] = a->pr[i].cost;
  }
  return GC_OK;
}

int gc_anytime_log(gc_anytime_t* a, uint32_t p, double* rows, uint32_t cap_rows)
{
  GC_REQUIRE(a && p < a->P, "gc_anytime_log: bad handle or problem index");
  const AnyProb& q = a->pr[p];
  const uint32_t n = (uint32_t)(q.log.size() / 8);
  if (rows)
    memcpy(rows, q.log.data(), sizeof(double) * 8 * std::min(n, cap_rows));
  return (int)n;
}

int gc_anytime_solution(gc_anytime_t* a, uint32_t p, double* waypoints, double* costs, uint32_t cap)
{
  GC_REQUIRE(a && p < a->P, "gc_anytime_solution: bad handle or problem index");
  const AnyProb& q = a->pr[p];
  if (!q.solved)
    return 0;
  const uint32_t n = (uint32_t)(q.sol_conf.size() / a->D);
  for (uint32_t i = 0; i < n && i < cap; i++)
  {
    if (waypoints)
      memcpy(waypoints + (size_t)i * a->D, q.sol_conf.data() + (size_t)i * a->D, sizeof(double) * a->D);
    if (costs && i > 0)
      costs[i - 1] = q.sol_costs[i - 1];
  }
  return (int)n;
}

int gc_anytime_dump(gc_anytime_t* a, uint32_t p, int which, int32_t* parents, double* pcost, double* conf, uint32_t cap)
{
  GC_REQUIRE(a && p < a->P, "gc_anytime_dump: bad handle or problem index");
  GC_SET_DEVICE(a->device);
  const AnyProb& q = a->pr[p];
  const int D = a->D;
  if (which == 1)
  {
    const uint32_t n = (uint32_t)q.t_parent.size();
    for (uint32_t i = 0; i < n && i < cap; i++)
    {
      if (parents)
        parents[i] = q.t_parent[i];
      if (pcost)
        pcost[i] = q.t_pcost[i];
      if (conf)
        memcpy(conf + (size_t)i * D, q.t_conf.data() + (size_t)i * D, sizeof(double) * D);
    }
    return (int)n;
  }
  const uint32_t n = (uint32_t)q.parent.size();
  const uint32_t m = std::min(n, cap);
  for (uint32_t i = 0; i < m; i++)
  {
    if (parents)
      parents[i] = q.parent[i];
    if (pcost)
      pcost[i] = q.pcost[i];
  }
  if (conf && m)
  {
    gc_store* s = a->store;
    GC_CUDA(cudaMemcpyAsync(conf, s->x64 + (size_t)p * s->cap_seg * D, sizeof(double) * (size_t)m * D,
                            cudaMemcpyDeviceToHost, s->stream));
    GC_CUDA(cudaStreamSynchronize(s->stream));
  }
  return (int)n;
}

}  // extern "C"
