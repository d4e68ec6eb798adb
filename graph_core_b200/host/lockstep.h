// lockstep.h -- host side of the B200 planner path: many independent graph_core planners advanced in
// lock step, their hot-path operators batched into a few CUDA launches per iteration.
//
// Each planner is an exact restatement of the reference's solver loop with the two hot paths
// replaced by calls through the C ABI (include/gcb200.h):
//   RRTStar::update(configuration, solution)   src/graph_core/solvers/rrt_star.cpp:89-163
//   RRTStar::updateRewireRadius                src/graph_core/solvers/rrt_star.cpp:278-286
//   Tree::rewire / extend / rewireOnly         src/graph_core/graph/tree.cpp:718-726, 148-161, 381-513
//   Tree::costToNode / getConnectionToNode     src/graph_core/graph/tree.cpp:767-813
//   TreeSolver::setProblem                     src/graph_core/solvers/tree_solver.cpp:55-105
//   InformedSampler::setCost (specific volume) src/graph_core/samplers/informed_sampler.cpp:181-215
// Decisions that depend on earlier decisions (rewireOnly mutates cost_to_node while it walks the
// near set) are replayed sequentially on the host; the device evaluates a superset of
// the edges those decisions can ask for, all at once.  With samples injected (tree_solver.h:338)
// every planner builds exactly the tree the reference would build.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "gcb200.h"

namespace gcb200
{
struct SolverConfig
{
  int dim = 0;
  double max_distance = 1.0;       // tree_solver.cpp:37
  double min_distance = 0.01;      // checker resolution (collision_checker_base.h:55)
  double rewire_factor = 1.1;      // rrt_star.cpp:46
  double utopia_tolerance = 0.01;  // tree_solver.cpp:40 (1 is added in config)
  std::vector<double> lb, ub;      // sampler bounds
};

// parent candidates per problem in the first edge batch of a step (see Problem::par_cand); GCB200_FIRST_BATCH
// overrides it for experiments
constexpr size_t kFirstBatchCandidates = 3;

struct StepStats
{
  uint64_t iterations = 0;   // planner updates executed (samples consumed)
  uint64_t extended = 0;     // updates that added a node
  uint64_t edges_device = 0; // edges evaluated on the device (speculative superset included)
  uint64_t edges_used = 0;   // edge results the replay actually consumed
  uint64_t edges_fallback = 0;
  uint64_t near_hits = 0;
  uint64_t gpu_calls = 0;    // C-ABI calls that launch kernels
  uint64_t h2d_bytes = 0;    // payload bytes handed to / received from the C ABI (host buffers)
  uint64_t d2h_bytes = 0;
  // host wall-clock per phase of step(): prologue | fused device pass | candidate sets | append + edge
  // batch | replay
  double t_phase[5] = { 0, 0, 0, 0, 0 };
};

class ThreadPool;

// P independent RRT* problems that share one collision world; problem p owns store segment p.
class LockstepRRTStar
{
public:
  // device_trees (default): parent slots and connection costs live on the GPU (gc_forest, include/gcb200.h) and a
  // step is one stream-ordered sequence of launches without a host round trip; accessors synchronise on demand.
  // device_trees = false keeps the host replay of round 1 (the same trees; used by the tests as a cross-check).
  LockstepRRTStar(const SolverConfig& cfg, gc_world_t* world, int device, uint32_t nproblems, const double* starts,
                  const double* goals, uint32_t capacity_per_tree, int threads, bool device_trees = true);
  ~LockstepRRTStar();

  // one RRTStar::update per still-improvable problem; samples is [nproblems][dim] (row p is used when
  // problem p is active).  Returns the number of problems that consumed a sample, <0 on error.
  int step(const double* samples);
  // same, with the [nproblems][dim] sample block resident in device memory (e.g. written by the Philox sampler)
  int stepDeviceSamples(const double* d_samples);
  // device trees only: the sample of problem p comes from the device informed sampler (InformedSampler::sample with the
  // problem's current path cost; Philox key `seed`, counter first_index + p) -- rrt_star.cpp:147-161
  int stepInformed(uint64_t seed, uint64_t first_index);
  // device trees only: wait for the enqueued steps; <0 when a capacity was exceeded (error() says which)
  int sync();
  // device trees only: enqueue a D2H copy of the nproblems current path costs into `costs` (pinned memory)
  int readCostsAsync(double* costs);
  bool deviceTrees() const { return forest_ != nullptr; }
  // run the store's kernels on an externally owned cudaStream_t (the world keeps its own setting)
  int setStream(void* cuda_stream);
  const char* error() const { return err_.c_str(); }

  uint32_t size() const { return (uint32_t)P_.size(); }
  bool solved(uint32_t p) const;
  bool canImprove(uint32_t p) const;
  double cost(uint32_t p) const;
  double radius(uint32_t p) const;
  uint32_t numNodes(uint32_t p) const;
  uint32_t treeSize(uint32_t p) const;
  // ids follow the reference's creation order: 0 start, 1 goal, 2.. new nodes
  void dump(uint32_t p, int32_t* parents, double* pcost, double* conf) const;
  std::vector<int32_t> solution(uint32_t p) const;
  const StepStats& stats() const;
  gc_store_t* store() { return store_; }
  // informed-sampler parameters of problem p (cost = inf until solved)
  double pathCost(uint32_t p) const { return P_[p].path_cost; }

private:
  struct Problem
  {
    uint32_t seg = 0;
    std::vector<double> conf;      // [id][dim]
    std::vector<int32_t> parent;   // by id
    std::vector<double> pcost;     // by id
    std::vector<int32_t> id_of_slot;
    std::vector<int32_t> slot_of_id;
    uint32_t tree_size = 0;        // nodes_->size()
    bool solved = false, can_improve = true;
    double goal_cost = 0, best_utopia = 0, path_cost = 0, cost = 0, r_rewire = 0;
    double specific_volume = 0, focii_distance = 0;
    std::vector<int32_t> solution;   // child ids root -> goal
    std::vector<double> sol_costs;   // Path keeps its Connection objects: costs frozen at creation
    // per-step scratch
    bool active = false, extended = false;
    int32_t new_node = -1;
    uint32_t row = 0;                // row in the batched device call
    uint32_t near_n = 0;
    double cost_to_node0 = 0;
    // Parent phase of rewireOnly: the walk ends with the collision-free candidate of least cost-to-come (first in walk
    // order on ties), so the candidates are tried in THAT order, the two best in a first edge batch and the rest --
    // only for the problems both failed -- in a second one together with the children and goal edges.
    std::vector<int32_t> par_cand;   // positions in the near list, sorted by (cost through the candidate, position)
    std::vector<double> par_total;   // cost through the candidate, same order
    std::vector<uint8_t> par_ok;     // verdicts, same order (valid up to par_known)
    uint32_t par_known = 0;
    int32_t par_winner = -1;         // index into par_cand, -1 = none free
    bool par_resolved = false;
    double cost_to_node1 = 0;        // cost-to-come after the parent phase (a lower bound while unresolved)
    uint32_t edgeA_first = 0, edgeA_count = 0;  // slices of the two batched edge lists
    uint32_t edge_first = 0, edge_count = 0;
    uint32_t parB_count = 0;         // parent candidates in the second batch (they lead the slice)
    std::vector<int32_t> chi_cand;   // positions in the near list, ascending
    std::vector<uint32_t> sort_scratch;
    std::vector<double> c2near;      // cost-to-come of every near node at the start of the step (walked once)
    int32_t goal_edge = -1;
  };

  int stepImpl(const double* samples, const double* d_samples);
  double costToNode(const Problem& pr, int32_t node) const;
  void setSamplerCost(Problem& pr, double cost);
  void makeSolution(Problem& pr);
  double solutionCost(const Problem& pr) const;
  int initialConnect();
  int createForest(const double* starts, const double* goals, uint32_t capacity_per_tree);
  int addNode(Problem& pr, const double* conf);  // host mirror only; returns id
  bool fail(const char* what);

  SolverConfig cfg_;
  size_t first_batch_ = kFirstBatchCandidates;
  int D_;
  gc_world_t* world_;
  gc_store_t* store_ = nullptr;
  gc_forest_t* forest_ = nullptr;
  std::vector<Problem> P_;
  std::unique_ptr<ThreadPool> pool_;
  mutable StepStats stats_;
  std::string err_;
  double utopia_tol_;
  // batched buffers
  std::vector<double> q_, radius_, next_, dist_, near_dist_, e1_, e2_, app_rows_;
  std::vector<uint32_t> seg_, closest_, near_idx_, near_count_, app_seg_, app_slots_;
  std::vector<uint8_t> ok_, eok_;
  uint32_t near_cap_ = 256;
};

}  // namespace gcb200

// ---- C interface for Python (ctypes) / other hosts ------------------------------------------------
extern "C" {
typedef struct gcp_rrtstar gcp_rrtstar_t;
int gcp_rrtstar_create(int dim, double max_distance, double min_distance, double rewire_factor,
                       double utopia_tolerance, const double* lb, const double* ub, gc_world_t* world, int device,
                       uint32_t nproblems, const double* starts, const double* goals, uint32_t capacity_per_tree,
                       int threads, gcp_rrtstar_t** out);
int gcp_rrtstar_destroy(gcp_rrtstar_t* s);
int gcp_rrtstar_step(gcp_rrtstar_t* s, const double* samples);
int gcp_rrtstar_step_dsamples(gcp_rrtstar_t* s, const double* d_samples);
int gcp_rrtstar_step_informed(gcp_rrtstar_t* s, uint64_t seed, uint64_t first_index);
int gcp_rrtstar_sync(gcp_rrtstar_t* s);
int gcp_rrtstar_read_costs_async(gcp_rrtstar_t* s, double* costs);
int gcp_rrtstar_create2(int dim, double max_distance, double min_distance, double rewire_factor,
                        double utopia_tolerance, const double* lb, const double* ub, gc_world_t* world, int device,
                        uint32_t nproblems, const double* starts, const double* goals, uint32_t capacity_per_tree,
                        int threads, int device_trees, gcp_rrtstar_t** out);
int gcp_rrtstar_set_stream(gcp_rrtstar_t* s, void* cuda_stream);
// runs n lock-step iterations; samples is [n][nproblems][dim]
int gcp_rrtstar_run(gcp_rrtstar_t* s, const double* samples, uint32_t n);
int gcp_rrtstar_info(gcp_rrtstar_t* s, uint32_t p, int32_t* solved, int32_t* can_improve, double* cost,
                     double* radius, uint32_t* num_nodes, uint32_t* tree_size);
int gcp_rrtstar_dump(gcp_rrtstar_t* s, uint32_t p, int32_t* parents, double* pcost, double* conf);
int gcp_rrtstar_solution(gcp_rrtstar_t* s, uint32_t p, int32_t* ids, uint32_t cap);
int gcp_rrtstar_stats(gcp_rrtstar_t* s, uint64_t out[16]);
const char* gcp_last_error(void);
}
