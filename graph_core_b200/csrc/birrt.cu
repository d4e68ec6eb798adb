// birrt.cu -- device-resident lock-step BiRRT / RRT (include/gcb200.h "gc_birrt"): BASELINE config 2.
//
//   BiRRT::update                src/graph_core/solvers/birrt.cpp:93-152
//   RRT::update                  src/graph_core/solvers/rrt.cpp:119-157
//   Tree::connect                src/graph_core/graph/tree.cpp:217-233
//   Tree::extend / tryExtend     src/graph_core/graph/tree.cpp:56-81, 130-161
//   Tree::selectNextConfiguration  src/graph_core/graph/tree.cpp:110-128
//   Path::cost                   src/graph_core/graph/path.cpp:285-291
//
// One launch advances P independent planning problems by one update(): one warp per tree (two per problem).  A warp
// finds the nearest node of ITS tree by an exact fp64 scan (lowest slot on ties, like Vector / KdTree), then runs the
// whole `connect` ray -- steer, bisection edge check with the lanes as states and a ballot early exit, append -- without
// leaving the kernel.  Inside a ray the next nearest node is the node just appended: every older node is at least as far
// from the sample as the previous nearest, and the new node is strictly closer (checked on the computed distances, with
// a re-scan as fall-back), so Tree::connect's repeated findClosestNode costs one scan per ray.  Warp 0 of a problem then
// tests whether the two rays met (birrt.cpp:118-126) and assembles the solution the way BiRRT::update flips and
// re-parents the goal branch.  Parent slots, connection costs and configurations live in HBM; nothing returns to the
// host between iterations.  Light worlds only (cube, boxes, C-space spheres): their single-state check is a few dozen
// flops, so the ray is latency-bound and the many-problems axis is what fills the machine.
#include <math_constants.h>

#include <cstring>
#include <limits>
#include <vector>

#include "edge_device.cuh"
#include "gc_common.cuh"

namespace gc
{
enum : uint32_t
{
  kBirrtErrSegmentFull = 1u,
  kBirrtErrDepth = 2u,
};
constexpr uint32_t kGoalTreeBit = 0x80000000u;

struct __align__(16) BNode
{
  double pcost;
  int32_t parent;
  int32_t pad;
};

struct BProb
{
  double cost;           // path cost (+inf while unsolved)
  uint32_t solved_iter;  // 1-based iteration that solved the problem
  uint32_t sol_len;      // connections of the solution
  uint32_t solved;
  uint32_t pad;
};

struct BirrtDev
{
  int D;
  uint32_t P, cap, depth_cap;
  int bidirectional, extend;
  double max_distance, min_d, tolerance;
  double* conf;   // [2P][cap][D]   tree 2p = start tree of problem p, 2p+1 = goal tree
  BNode* nodes;   // [2P][cap]
  uint32_t* used; // [2P]
  BProb* prob;    // [P]
  const double* goals;    // [P][D]
  const double* samples;  // [P][D]
  uint32_t* sol_slot;     // [P][depth_cap + 1] way points: slot | kGoalTreeBit
  double* sol_cost;       // [P][depth_cap]
  uint32_t* err;
  unsigned long long* stats;  // [0] updates of unsolved problems [1] nodes added [2] edges checked [3] states
                              // [4] (state, object) tests [5] scans [6] re-scans inside a ray [7] problems solved
  uint32_t iteration;
  int setup;  // 1: TreeSolver::setProblem's direct connection start -> goal instead of an update
  // world
  int kind;
  uint32_t nobj;
  double cube_thr;
  const double* lo;
  const double* hi;
  const float* sc;
  const float* sr2;
};

struct RayStats
{
  unsigned long long added = 0, edges = 0, states = 0, pairs = 0, scans = 0, rescans = 0;
};

// checkConnection(c1, c2) (collision_checker_base.h:207-237) by one warp: lane = state, ballot exit per 32 states
template <int D>
__device__ __forceinline__ bool warp_edge_free(const LightWorld& w, const double* c1, const double* c2, double min_d,
                                               int lane, RayStats& st)
{
  const EdgeGeom g = edge_geom<D, GC_EDGE_BISECT>(c1, c2, c1, min_d);
  st.edges += (lane == 0);
  for (uint32_t s0 = 0; s0 < g.nstates; s0 += 32u)
  {
    const uint32_t s = s0 + lane;
    bool bad = false;
    if (s < g.nstates)
    {
      double conf[D];
      if (state_conf<D, GC_EDGE_BISECT>(c1, c2, c1, g, s, conf))
      {
        bad = !light_check<D>(w, conf, st.pairs);
        st.states++;
      }
    }
    if (__ballot_sync(0xffffffffu, bad))
      return false;
  }
  return true;
}

// nearest node of the tree to q: exact fp64 distances, lowest slot on ties
template <int D>
__device__ __forceinline__ void warp_nearest(const double* __restrict__ conf, uint32_t n, const double* q, int lane,
                                             double& best, uint32_t& bslot)
{
  best = CUDART_INF;
  bslot = GC_NO_NODE;
  for (uint32_t i = lane; i < n; i += 32u)
  {
    const double d = norm_diff<D>(conf + (size_t)i * D, q);
    if (d < best)
    {
      best = d;
      bslot = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    const double od = __shfl_xor_sync(0xffffffffu, best, o);
    const uint32_t os = __shfl_xor_sync(0xffffffffu, bslot, o);
    if (od < best || (od == best && os < bslot))
    {
      best = od;
      bslot = os;
    }
  }
}

// Tree::connect (tree.cpp:217-233) or, with b.extend, one Tree::extend (tree.cpp:148-161).  Returns the reference's
// boolean; new_node = the last node appended (valid when true).
template <int D>
__device__ bool warp_connect(const BirrtDev& b, const LightWorld& w, uint32_t tree, const double* q, int lane,
                             int32_t& new_node, RayStats& st, bool force_connect)
{
  double* conf = b.conf + (size_t)tree * b.cap * D;
  BNode* nodes = b.nodes + (size_t)tree * b.cap;
  uint32_t n = b.used[tree];
  double dist;
  uint32_t closest;
  warp_nearest<D>(conf, n, q, lane, dist, closest);
  st.scans += (lane == 0);
  bool result = false;
  for (;;)
  {
    double cn[D], nx[D];
#pragma unroll
    for (int i = 0; i < D; i++)
    {
      cn[i] = conf[(size_t)closest * D + i];
      // selectNextConfiguration (tree.cpp:110-128): the sample itself within max_distance, else one step towards it
      nx[i] = (dist <= b.max_distance) ? q[i] :
                                         __dadd_rn(cn[i], __dmul_rn(__ddiv_rn(__dsub_rn(q[i], cn[i]), dist), b.max_distance));
    }
    // tryExtendFromNode (tree.cpp:69-81)
    const bool ok = (dist < b.tolerance) ? true : warp_edge_free<D>(w, cn, nx, b.min_d, lane, st);
    if (!ok)
      break;
    if (n >= b.cap)
    {
      if (lane == 0)
        atomicOr(b.err, kBirrtErrSegmentFull);
      break;
    }
    const uint32_t slot = n++;
    if (lane == 0)
    {
#pragma unroll
      for (int i = 0; i < D; i++)
        conf[(size_t)slot * D + i] = nx[i];
      BNode nd;
      nd.parent = (int32_t)closest;
      nd.pcost = norm_diff<D>(cn, nx);  // extendOnly: metrics_->cost(closest_node, new_node) (tree.cpp:132)
      nd.pad = 0;
      nodes[slot] = nd;
      st.added++;
    }
    __syncwarp();
    new_node = (int32_t)slot;
    if (b.extend && !force_connect)
    {
      result = true;
      break;
    }
    const double dnew = norm_diff<D>(nx, q);
    if (dnew < b.tolerance)  // tree.cpp:228
    {
      result = true;
      break;
    }
    if (dnew < dist)
    {
      closest = slot;  // strictly closer than the previous nearest, hence than every older node
      dist = dnew;
    }
    else
    {
      warp_nearest<D>(conf, n, q, lane, dist, closest);
      st.rescans += (lane == 0);
    }
  }
  if (lane == 0)
    b.used[tree] = n;
  return result;
}

template <int D>
__global__ void __launch_bounds__(256) birrt_step_kernel(BirrtDev b)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LightWorld w;
  light_stage<D>(w, smem_raw, b.kind, b.nobj, b.cube_thr, b.lo, b.hi, b.sc, b.sr2);
  __shared__ int s_added[8];
  __shared__ int32_t s_new[8];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const uint32_t p = blockIdx.x * 4u + (uint32_t)(wib >> 1);
  const int t = wib & 1;
  const bool live = p < b.P && !b.prob[p].solved;
  RayStats st;
  double q[D];
  int32_t new_node = -1;
  bool added = false;
  const bool setup = b.setup != 0;
  if (live && (t == 0 || (b.bidirectional && !setup)))
  {
    // setup: Tree::connectToNode(goal) (tree.cpp:304-327, through extendToNode :192-215) -- the same ray as connect;
    // the node that lands on the goal configuration IS the goal node
    const double* src = setup ? b.goals : b.samples;
#pragma unroll
    for (int i = 0; i < D; i++)
      q[i] = src[(size_t)p * D + i];
    added = warp_connect<D>(b, w, 2u * p + (uint32_t)t, q, lane, new_node, st, setup);
  }
  if (lane == 0)
  {
    s_added[wib] = added ? 1 : 0;
    s_new[wib] = new_node;
  }
  __syncthreads();
  if (live && t == 0)
  {
    double* conf_s = b.conf + (size_t)(2u * p) * b.cap * D;
    BNode* nodes_s = b.nodes + (size_t)(2u * p) * b.cap;
    const int32_t new_s = s_new[wib];
    bool solved = false;
    uint32_t* way = b.sol_slot + (size_t)p * (b.depth_cap + 1);
    double* wcost = b.sol_cost + (size_t)p * b.depth_cap;
    uint32_t len = 0;
    if (setup)
    {
      // TreeSolver::setProblem (tree_solver.cpp:84-98): a direct connection is the solution
      if (s_added[wib])
      {
        solved = true;
        if (lane == 0)
        {
          uint32_t ls = 0;
          for (int32_t x = new_s; x != 0; x = nodes_s[x].parent)
            ls++;
          len = ls;
          if (len > b.depth_cap)
          {
            atomicOr(b.err, kBirrtErrDepth);
            len = 0;
          }
          else
          {
            uint32_t k = ls;
            way[0] = 0u;
            for (int32_t x = new_s; x != 0;)
            {
              const BNode nd = nodes_s[x];
              way[k] = (uint32_t)x;
              wcost[k - 1] = nd.pcost;
              k--;
              x = nd.parent;
            }
          }
        }
      }
    }
    else if (b.bidirectional)
    {
      const double* conf_g = b.conf + (size_t)(2u * p + 1u) * b.cap * D;
      const BNode* nodes_g = b.nodes + (size_t)(2u * p + 1u) * b.cap;
      const int32_t new_g = s_new[wib + 1];
      if (s_added[wib] && s_added[wib + 1] &&
          norm_diff<D>(conf_s + (size_t)new_s * D, conf_g + (size_t)new_g * D) < b.tolerance)  // birrt.cpp:118-126
      {
        solved = true;
        if (lane == 0)
        {
          // start branch root -> new_start_node, then new_start_node -> parent of new_goal_node with that connection's
          // cost, then the goal branch flipped down to the goal (birrt.cpp:128-141)
          uint32_t ls = 0, lg = 0;
          for (int32_t x = new_s; x != 0; x = nodes_s[x].parent)
            ls++;
          const int32_t gp = nodes_g[new_g].parent;
          for (int32_t x = gp; x != 0; x = nodes_g[x].parent)
            lg++;
          len = ls + 1u + lg;
          if (len > b.depth_cap)
          {
            atomicOr(b.err, kBirrtErrDepth);
            len = 0;
          }
          else
          {
            uint32_t k = ls;
            way[0] = 0u;
            for (int32_t x = new_s; x != 0;)
            {
              const BNode nd = nodes_s[x];
              way[k] = (uint32_t)x;
              wcost[k - 1] = nd.pcost;
              k--;
              x = nd.parent;
            }
            way[ls + 1] = (uint32_t)gp | kGoalTreeBit;
            wcost[ls] = nodes_g[new_g].pcost;
            uint32_t idx = ls + 1;
            for (int32_t x = gp; x != 0;)
            {
              const BNode nd = nodes_g[x];
              wcost[idx] = nd.pcost;
              idx++;
              way[idx] = (uint32_t)nd.parent | kGoalTreeBit;
              x = nd.parent;
            }
          }
        }
      }
    }
    else if (s_added[wib])
    {
      // RRT::update (rrt.cpp:137-153): connect the new node to the goal when it is within max_distance and free
      const double* goal = b.goals + (size_t)p * D;
      double nn[D], gg[D];
#pragma unroll
      for (int i = 0; i < D; i++)
      {
        nn[i] = conf_s[(size_t)new_s * D + i];
        gg[i] = goal[i];
      }
      const double dg = norm_diff<D>(nn, gg);
      if (dg <= b.max_distance && warp_edge_free<D>(w, nn, gg, b.min_d, lane, st))
      {
        uint32_t n = b.used[2u * p];
        if (n >= b.cap)
        {
          if (lane == 0)
            atomicOr(b.err, kBirrtErrSegmentFull);
        }
        else
        {
          solved = true;
          if (lane == 0)
          {
#pragma unroll
            for (int i = 0; i < D; i++)
              conf_s[(size_t)n * D + i] = gg[i];
            BNode nd;
            nd.parent = new_s;
            nd.pcost = dg;  // metrics_->cost(new_start_node, goal_node_)
            nd.pad = 0;
            nodes_s[n] = nd;
            b.used[2u * p] = n + 1;  // start_tree_->addNode(goal_node_)
            uint32_t ls = 0;
            for (int32_t x = (int32_t)n; x != 0; x = nodes_s[x].parent)
              ls++;
            len = ls;
            if (len > b.depth_cap)
            {
              atomicOr(b.err, kBirrtErrDepth);
              len = 0;
            }
            else
            {
              uint32_t k = ls;
              way[0] = 0u;
              for (int32_t x = (int32_t)n; x != 0;)
              {
                const BNode nd2 = nodes_s[x];
                way[k] = (uint32_t)x;
                wcost[k - 1] = nd2.pcost;
                k--;
                x = nd2.parent;
              }
            }
          }
        }
      }
    }
    if (solved && lane == 0)
    {
      double c = 0.0;  // Path::cost: connection costs summed start -> goal
      for (uint32_t k = 0; k < len; k++)
        c = __dadd_rn(c, wcost[k]);
      BProb& pr = b.prob[p];
      pr.cost = c;
      pr.sol_len = len;
      pr.solved_iter = b.iteration;
      pr.solved = 1u;
      atomicAdd(&b.stats[7], 1ull);
    }
    if (lane == 0 && !setup)
      atomicAdd(&b.stats[0], 1ull);
  }
  // statistics: one atomic per warp and counter
  unsigned long long v[6] = { st.added, st.edges, st.states, st.pairs, st.scans, st.rescans };
#pragma unroll
  for (int k = 0; k < 6; k++)
  {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
    if (lane == 0 && v[k])
      atomicAdd(&b.stats[1 + k], v[k]);
  }
}

#define GC_BIRRT_DISPATCH(dim, ...)                         \
  switch (dim)                                              \
  {                                                         \
    case 1: { constexpr int DD = 1; __VA_ARGS__; } break;   \
    case 2: { constexpr int DD = 2; __VA_ARGS__; } break;   \
    case 3: { constexpr int DD = 3; __VA_ARGS__; } break;   \
    case 4: { constexpr int DD = 4; __VA_ARGS__; } break;   \
    case 5: { constexpr int DD = 5; __VA_ARGS__; } break;   \
    case 6: { constexpr int DD = 6; __VA_ARGS__; } break;   \
    case 7: { constexpr int DD = 7; __VA_ARGS__; } break;   \
    case 8: { constexpr int DD = 8; __VA_ARGS__; } break;   \
    case 9: { constexpr int DD = 9; __VA_ARGS__; } break;   \
    case 10: { constexpr int DD = 10; __VA_ARGS__; } break; \
    case 11: { constexpr int DD = 11; __VA_ARGS__; } break; \
    case 12: { constexpr int DD = 12; __VA_ARGS__; } break; \
    case 13: { constexpr int DD = 13; __VA_ARGS__; } break; \
    case 14: { constexpr int DD = 14; __VA_ARGS__; } break; \
    case 15: { constexpr int DD = 15; __VA_ARGS__; } break; \
    case 16: { constexpr int DD = 16; __VA_ARGS__; } break; \
    default: gc::set_error("dimension %d not supported", dim); return GC_E_INVALID; \
  }
}  // namespace gc

using namespace gc;

struct gc_birrt
{
  int device = 0;
  gc_world* world = nullptr;
  BirrtDev d{};
  size_t smem = 0;
  double* d_samples = nullptr;  // [P][D] staging for host samples
  gc::PinBuf h_samples;
  std::vector<void*> allocs;
};

extern "C" {

int gc_birrt_destroy(gc_birrt_t* b)
{
  if (!b)
    return GC_OK;
  GC_SET_DEVICE(b->device);
  if (b->world)
    cudaStreamSynchronize(b->world->stream);
  for (void* p : b->allocs)
    cudaFree(p);
  b->h_samples.release();
  if (b->world)
    gc_world_destroy(b->world);
  delete b;
  return GC_OK;
}

int gc_birrt_create(gc_world_t* w, const gc_birrt_config_t* cfg, uint32_t n_problems, const double* starts,
                    const double* goals, gc_birrt_t** out)
{
  GC_REQUIRE(w && cfg && starts && goals && out, "gc_birrt_create: NULL argument");
  *out = nullptr;
  GC_REQUIRE(n_problems > 0 && cfg->capacity >= 2, "gc_birrt_create: empty problem set or capacity");
  GC_REQUIRE(w->kind == GC_WORLD_CUBE || w->kind == GC_WORLD_BOXES || w->kind == GC_WORLD_CSPHERES,
             "gc_birrt_create: the in-kernel ray needs a light world (cube, boxes, C-space spheres)");
  GC_REQUIRE(cfg->max_distance > 0.0 && cfg->min_distance > 0.0, "gc_birrt_create: distances must be positive");
  GC_SET_DEVICE(w->device);
  gc_birrt* b = new gc_birrt();
  b->device = w->device;
  gc_world_retain(w);
  b->world = w;
  BirrtDev& d = b->d;
  const int D = w->dim;
  d.D = D;
  d.P = n_problems;
  d.cap = cfg->capacity;
  d.depth_cap = cfg->depth_cap ? cfg->depth_cap : 4096u;
  d.bidirectional = cfg->bidirectional ? 1 : 0;
  d.extend = cfg->extend ? 1 : 0;
  d.max_distance = cfg->max_distance;
  d.min_d = cfg->min_distance;
  d.tolerance = cfg->tolerance > 0.0 ? cfg->tolerance : 1e-06;
  d.kind = w->kind;
  d.nobj = w->nobj;
  d.cube_thr = w->cube_thr;
  const gc_world* root = w->parent ? w->parent : w;
  d.lo = root->d_lo;
  d.hi = root->d_hi;
  d.sc = root->d_sc;
  d.sr2 = root->d_sr2;
  b->smem = light_smem_bytes(w->kind, w->nobj, D);
  const size_t trees = 2 * (size_t)n_problems;
  auto alloc = [&](void** p, size_t bytes) -> cudaError_t {
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaSuccess)
    {
      b->allocs.push_back(*p);
      e = cudaMemsetAsync(*p, 0, bytes, w->stream);
    }
    return e;
  };
  double* d_goals = nullptr;
  cudaError_t e = alloc((void**)&d.conf, sizeof(double) * trees * d.cap * D);
  if (e == cudaSuccess)
    e = alloc((void**)&d.nodes, sizeof(BNode) * trees * d.cap);
  if (e == cudaSuccess)
    e = alloc((void**)&d.used, sizeof(uint32_t) * trees);
  if (e == cudaSuccess)
    e = alloc((void**)&d.prob, sizeof(BProb) * n_problems);
  if (e == cudaSuccess)
    e = alloc((void**)&d_goals, sizeof(double) * (size_t)n_problems * D);
  if (e == cudaSuccess)
    e = alloc((void**)&b->d_samples, sizeof(double) * (size_t)n_problems * D);
  if (e == cudaSuccess)
    e = alloc((void**)&d.sol_slot, sizeof(uint32_t) * (size_t)n_problems * (d.depth_cap + 1));
  if (e == cudaSuccess)
    e = alloc((void**)&d.sol_cost, sizeof(double) * (size_t)n_problems * d.depth_cap);
  if (e == cudaSuccess)
    e = alloc((void**)&d.err, sizeof(uint32_t) * 4);
  if (e == cudaSuccess)
    e = alloc((void**)&d.stats, sizeof(unsigned long long) * 8);
  if (e != cudaSuccess)
  {
    gc::set_error("gc_birrt_create: %s", cudaGetErrorString(e));
    gc_birrt_destroy(b);
    return GC_E_CUDA;
  }
  d.goals = d_goals;
  // roots: slot 0 of tree 2p = start, of tree 2p+1 = goal (birrt.cpp:65-70); parent -1
  std::vector<BNode> roots(trees);
  std::vector<uint32_t> used(trees, 1u);
  std::vector<BProb> prob(n_problems);
  for (auto& r : roots)
  {
    r.parent = -1;
    r.pcost = 0.0;
    r.pad = 0;
  }
  for (auto& pr : prob)
  {
    pr.cost = std::numeric_limits<double>::infinity();
    pr.solved = 0;
    pr.solved_iter = 0;
    pr.sol_len = 0;
    pr.pad = 0;
  }
  cudaStream_t st = w->stream;
  for (uint32_t p = 0; p < n_problems && e == cudaSuccess; p++)
  {
    e = cudaMemcpyAsync(d.conf + (size_t)(2 * p) * d.cap * D, starts + (size_t)p * D, sizeof(double) * D,
                        cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(d.conf + (size_t)(2 * p + 1) * d.cap * D, goals + (size_t)p * D, sizeof(double) * D,
                          cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(d.nodes + (size_t)(2 * p) * d.cap, &roots[0], sizeof(BNode), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(d.nodes + (size_t)(2 * p + 1) * d.cap, &roots[0], sizeof(BNode), cudaMemcpyHostToDevice, st);
  }
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(d.used, used.data(), sizeof(uint32_t) * trees, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(d.prob, prob.data(), sizeof(BProb) * n_problems, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(d_goals, goals, sizeof(double) * (size_t)n_problems * D, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(st);
  if (e != cudaSuccess)
  {
    gc::set_error("gc_birrt_create: %s", cudaGetErrorString(e));
    gc_birrt_destroy(b);
    return GC_E_CUDA;
  }
  // TreeSolver::setProblem (tree_solver.cpp:84-104): the direct connection start -> goal, which also leaves its ray in
  // the start tree when it is blocked half way
  {
    BirrtDev ds = b->d;
    ds.setup = 1;
    ds.iteration = 0;
    if (b->smem > 48 * 1024)
      GC_BIRRT_DISPATCH(ds.D, GC_CUDA(cudaFuncSetAttribute(birrt_step_kernel<DD>,
                                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem)));
    GC_BIRRT_DISPATCH(ds.D, (birrt_step_kernel<DD><<<(ds.P + 3u) / 4u, 256, b->smem, st>>>(ds)));
    e = cudaGetLastError();
    count_launch();
    if (e == cudaSuccess)
      e = cudaStreamSynchronize(st);
    if (e != cudaSuccess)
    {
      gc::set_error("gc_birrt_create: %s", cudaGetErrorString(e));
      gc_birrt_destroy(b);
      return GC_E_CUDA;
    }
  }
  *out = b;
  return GC_OK;
}

int gc_birrt_step(gc_birrt_t* b, const double* samples, const double* d_samples)
{
  GC_REQUIRE(b != nullptr, "gc_birrt_step: NULL handle");
  GC_REQUIRE((samples != nullptr) != (d_samples != nullptr), "gc_birrt_step: give host samples or device samples");
  GC_SET_DEVICE(b->device);
  cudaStream_t st = b->world->stream;
  BirrtDev d = b->d;
  if (samples)
  {
    const size_t bytes = sizeof(double) * (size_t)d.P * d.D;
    GC_TRY(b->h_samples.reserve(bytes));
    GC_CUDA(cudaStreamSynchronize(st));  // the staging buffer of the previous step has been consumed
    std::memcpy(b->h_samples.p, samples, bytes);
    GC_CUDA(cudaMemcpyAsync(b->d_samples, b->h_samples.p, bytes, cudaMemcpyHostToDevice, st));
    d.samples = b->d_samples;
  }
  else
    d.samples = d_samples;
  d.iteration = ++b->d.iteration;
  d.setup = 0;
  if (b->smem > 48 * 1024)
    GC_BIRRT_DISPATCH(d.D, GC_CUDA(cudaFuncSetAttribute(birrt_step_kernel<DD>,
                                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem)));
  const uint32_t blocks = (d.P + 3u) / 4u;
  GC_BIRRT_DISPATCH(d.D, (birrt_step_kernel<DD><<<blocks, 256, b->smem, st>>>(d)));
  GC_CUDA(cudaGetLastError());
  count_launch();
  return GC_OK;
}

int gc_birrt_sync(gc_birrt_t* b)
{
  GC_REQUIRE(b != nullptr, "gc_birrt_sync: NULL handle");
  GC_SET_DEVICE(b->device);
  uint32_t err = 0;
  GC_CUDA(cudaMemcpyAsync(&err, b->d.err, sizeof(uint32_t), cudaMemcpyDeviceToHost, b->world->stream));
  GC_CUDA(cudaStreamSynchronize(b->world->stream));
  if (err)
  {
    gc::set_error("gc_birrt: capacity exceeded since creation (%s%s): the trees are no longer the reference's",
                  (err & kBirrtErrSegmentFull) ? " tree capacity" : "", (err & kBirrtErrDepth) ? " solution depth" : "");
    return GC_E_CAPACITY;
  }
  return GC_OK;
}

int gc_birrt_stats(gc_birrt_t* b, uint64_t out[8])
{
  GC_REQUIRE(b && out, "gc_birrt_stats: NULL argument");
  GC_SET_DEVICE(b->device);
  GC_CUDA(cudaMemcpyAsync(out, b->d.stats, sizeof(uint64_t) * 8, cudaMemcpyDeviceToHost, b->world->stream));
  GC_CUDA(cudaStreamSynchronize(b->world->stream));
  return GC_OK;
}

int gc_birrt_info(gc_birrt_t* b, uint32_t p, int32_t* solved, uint32_t* solved_iteration, double* cost,
                  uint32_t* size_start, uint32_t* size_goal)
{
  GC_REQUIRE(b && p < b->d.P, "gc_birrt_info: bad handle or problem");
  GC_SET_DEVICE(b->device);
  BProb pr;
  uint32_t used[2];
  GC_CUDA(cudaMemcpyAsync(&pr, b->d.prob + p, sizeof(BProb), cudaMemcpyDeviceToHost, b->world->stream));
  GC_CUDA(cudaMemcpyAsync(used, b->d.used + 2 * (size_t)p, sizeof(used), cudaMemcpyDeviceToHost, b->world->stream));
  GC_CUDA(cudaStreamSynchronize(b->world->stream));
  if (solved)
    *solved = (int32_t)pr.solved;
  if (solved_iteration)
    *solved_iteration = pr.solved_iter;
  if (cost)
    *cost = pr.cost;
  if (size_start)
    *size_start = used[0];
  if (size_goal)
    *size_goal = used[1];
  return GC_OK;
}

int gc_birrt_read_solved(gc_birrt_t* b, uint8_t* solved, double* cost)
{
  GC_REQUIRE(b && solved, "gc_birrt_read_solved: NULL argument");
  GC_SET_DEVICE(b->device);
  std::vector<BProb> pr(b->d.P);
  GC_CUDA(cudaMemcpyAsync(pr.data(), b->d.prob, sizeof(BProb) * b->d.P, cudaMemcpyDeviceToHost, b->world->stream));
  GC_CUDA(cudaStreamSynchronize(b->world->stream));
  for (uint32_t p = 0; p < b->d.P; p++)
  {
    solved[p] = (uint8_t)pr[p].solved;
    if (cost)
      cost[p] = pr[p].cost;
  }
  return GC_OK;
}

int gc_birrt_dump(gc_birrt_t* b, uint32_t p, int tree, int32_t* parents, double* pcost, double* conf)
{
  GC_REQUIRE(b && p < b->d.P && (tree == 0 || tree == 1), "gc_birrt_dump: bad handle, problem or tree");
  GC_SET_DEVICE(b->device);
  cudaStream_t st = b->world->stream;
  const size_t t = 2 * (size_t)p + tree;
  uint32_t n = 0;
  GC_CUDA(cudaMemcpyAsync(&n, b->d.used + t, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  GC_CUDA(cudaStreamSynchronize(st));
  std::vector<BNode> nodes(n);
  GC_CUDA(cudaMemcpyAsync(nodes.data(), b->d.nodes + t * b->d.cap, sizeof(BNode) * n, cudaMemcpyDeviceToHost, st));
  if (conf)
    GC_CUDA(cudaMemcpyAsync(conf, b->d.conf + t * b->d.cap * b->d.D, sizeof(double) * (size_t)n * b->d.D,
                            cudaMemcpyDeviceToHost, st));
  GC_CUDA(cudaStreamSynchronize(st));
  for (uint32_t i = 0; i < n; i++)
  {
    if (parents)
      parents[i] = nodes[i].parent;
    if (pcost)
      pcost[i] = nodes[i].pcost;
  }
  return (int)n;
}

int gc_birrt_solution(gc_birrt_t* b, uint32_t p, double* waypoints, double* costs, uint32_t cap)
{
  GC_REQUIRE(b && p < b->d.P, "gc_birrt_solution: bad handle or problem");
  GC_SET_DEVICE(b->device);
  cudaStream_t st = b->world->stream;
  BProb pr;
  GC_CUDA(cudaMemcpyAsync(&pr, b->d.prob + p, sizeof(BProb), cudaMemcpyDeviceToHost, st));
  GC_CUDA(cudaStreamSynchronize(st));
  if (!pr.solved || pr.sol_len == 0)
    return 0;
  const uint32_t nw = pr.sol_len + 1;
  std::vector<uint32_t> way(nw);
  std::vector<double> wc(pr.sol_len);
  GC_CUDA(cudaMemcpyAsync(way.data(), b->d.sol_slot + (size_t)p * (b->d.depth_cap + 1), sizeof(uint32_t) * nw,
                          cudaMemcpyDeviceToHost, st));
  GC_CUDA(cudaMemcpyAsync(wc.data(), b->d.sol_cost + (size_t)p * b->d.depth_cap, sizeof(double) * pr.sol_len,
                          cudaMemcpyDeviceToHost, st));
  GC_CUDA(cudaStreamSynchronize(st));
  const int D = b->d.D;
  for (uint32_t i = 0; i < nw && i < cap; i++)
  {
    const size_t t = 2 * (size_t)p + ((way[i] & kGoalTreeBit) ? 1 : 0);
    const uint32_t slot = way[i] & ~kGoalTreeBit;
    if (waypoints)
      GC_CUDA(cudaMemcpyAsync(waypoints + (size_t)i * D, b->d.conf + (t * b->d.cap + slot) * D, sizeof(double) * D,
                              cudaMemcpyDeviceToHost, st));
    if (costs && i < pr.sol_len)
      costs[i] = wc[i];
  }
  GC_CUDA(cudaStreamSynchronize(st));
  return (int)nw;
}

}  // extern "C"
