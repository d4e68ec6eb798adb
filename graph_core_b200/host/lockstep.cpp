// lockstep.cpp -- see lockstep.h.  Host C++ above the C ABI; all hot-path work happens in libgcb200.so.
#include "lockstep.h"

#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <limits>
#include <mutex>
#include <string>
#include <thread>

namespace gcb200
{
static const double kInf = std::numeric_limits<double>::infinity();
static const double TOLERANCE = 1e-06;  // include/graph_core/util.h:80

// Eigen (a-b).norm(): sequential sum of squares, individually rounded (built with -ffp-contract=off)
static inline double dist(const double* a, const double* b, int d)
{
  double s = 0.0;
  for (int i = 0; i < d; i++)
  {
    const double t = a[i] - b[i];
    s = s + t * t;
  }
  return std::sqrt(s);
}

// ---- minimal persistent thread pool: parallel_for over problem indices --------------------------
class ThreadPool
{
public:
  explicit ThreadPool(int n) : stop_(false), gen_(0), pending_(0)
  {
    for (int i = 0; i < n; i++)
      workers_.emplace_back([this, i, n] { loop(i, n); });
  }
  ~ThreadPool()
  {
    {
      std::unique_lock<std::mutex> lk(m_);
      stop_ = true;
      gen_++;
    }
    cv_.notify_all();
    for (auto& t : workers_)
      t.join();
  }
  // runs fn(i) for i in [0,n); the calling thread takes part
  void parallel_for(uint32_t n, const std::function<void(uint32_t)>& fn)
  {
    if (workers_.empty() || n < 8)
    {
      for (uint32_t i = 0; i < n; i++)
        fn(i);
      return;
    }
    {
      std::unique_lock<std::mutex> lk(m_);
      fn_ = &fn;
      n_ = n;
      next_.store(0);
      pending_ = (int)workers_.size();
      gen_++;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

private:
  void work()
  {
    const uint32_t chunk = 4;
    for (;;)
    {
      const uint32_t b = next_.fetch_add(chunk);
      if (b >= n_)
        break;
      const uint32_t e = std::min(n_, b + chunk);
      for (uint32_t i = b; i < e; i++)
        (*fn_)(i);
    }
  }
  void loop(int, int)
  {
    uint64_t seen = 0;
    for (;;)
    {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_)
          return;
      }
      work();
      {
        std::unique_lock<std::mutex> lk(m_);
        if (--pending_ == 0)
          done_.notify_one();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  bool stop_;
  uint64_t gen_;
  int pending_;
  const std::function<void(uint32_t)>* fn_ = nullptr;
  uint32_t n_ = 0;
  std::atomic<uint32_t> next_{ 0 };
};

// =================================================================================================
LockstepRRTStar::LockstepRRTStar(const SolverConfig& cfg, gc_world_t* world, int device, uint32_t nproblems,
                                 const double* starts, const double* goals, uint32_t capacity_per_tree, int threads,
                                 bool device_trees)
  : cfg_(cfg), D_(cfg.dim), world_(world)
{
  if (const char* hr = std::getenv("GCB200_HOST_REPLAY"))
    if (std::atoi(hr) != 0)
      device_trees = false;
  if (const char* fb = std::getenv("GCB200_FIRST_BATCH"))
    first_batch_ = (size_t)std::max(1, std::atoi(fb));
  // TreeSolver::config (tree_solver.cpp:40-47)
  utopia_tol_ = cfg.utopia_tolerance;
  if (utopia_tol_ <= 0.0)
    utopia_tol_ = 0.0;
  utopia_tol_ += 1.0;
  // device trees: one spare, always empty segment absorbs the queries of problems that stopped improving
  if (gc_store_create(D_, capacity_per_tree, nproblems + (device_trees ? 1u : 0u), device, &store_) != GC_OK)
  {
    err_ = gc_last_error();
    store_ = nullptr;
    return;
  }
  if (threads > 1)
    pool_.reset(new ThreadPool(threads - 1));
  else
    pool_.reset(new ThreadPool(0));
  P_.resize(nproblems);
  std::vector<double> rows((size_t)nproblems * D_);
  std::vector<uint32_t> segs(nproblems), slots(nproblems);
  for (uint32_t p = 0; p < nproblems; p++)
  {
    Problem& pr = P_[p];
    pr.seg = p;
    // ids: 0 = start (tree root, RRT::addStart rrt.cpp:58-78), 1 = goal (RRT::addGoal rrt.cpp:34-56)
    addNode(pr, starts + (size_t)p * D_);
    addNode(pr, goals + (size_t)p * D_);
    pr.slot_of_id[0] = 0;
    pr.id_of_slot.push_back(0);
    pr.tree_size = 1;
    std::memcpy(&rows[(size_t)p * D_], starts + (size_t)p * D_, sizeof(double) * D_);
    segs[p] = p;
    // InformedSampler(start, goal, lb, ub): focii distance, specific volume at infinite cost
    pr.focii_distance = dist(starts + (size_t)p * D_, goals + (size_t)p * D_, D_);
    setSamplerCost(pr, kInf);
    // TreeSolver::setProblem (tree_solver.cpp:55-72)
    pr.goal_cost = 0.0;  // NullGoalCostFunction
    pr.best_utopia = pr.goal_cost + dist(&pr.conf[0], &pr.conf[D_], D_);
    pr.path_cost = kInf;
    pr.cost = kInf;
  }
  if (gc_store_append_multi(store_, segs.data(), rows.data(), nproblems, slots.data()) != GC_OK)
  {
    err_ = gc_last_error();
    return;
  }
  stats_.gpu_calls++;
  if (initialConnect() < 0)
    return;
  if (device_trees)
    createForest(starts, goals, capacity_per_tree);
}

// hands the trees as they stand after TreeSolver::setProblem to the device-resident forest
int LockstepRRTStar::createForest(const double* starts, const double* goals, uint32_t capacity_per_tree)
{
  const uint32_t n = (uint32_t)P_.size();
  gc_forest_config_t fc{};
  fc.max_distance = cfg_.max_distance;
  fc.min_distance = cfg_.min_distance;
  fc.tolerance = TOLERANCE;
  fc.rewire_factor = cfg_.rewire_factor;
  fc.utopia_tolerance = utopia_tol_;
  fc.near_cap = 512;
  fc.edges_cap = std::max<uint32_t>(4096u, n * 48u);
  fc.depth_cap = std::min<uint32_t>(capacity_per_tree, 1024u);
  fc.first_batch = (uint32_t)first_batch_;
  if (const char* v = std::getenv("GCB200_NEAR_CAP"))
    fc.near_cap = (uint32_t)std::max(1, std::atoi(v));
  if (const char* v = std::getenv("GCB200_EDGES_PER_PROBLEM"))
    fc.edges_cap = std::max<uint32_t>(64u, n * (uint32_t)std::max(1, std::atoi(v)));
  std::vector<uint32_t> first(n + 1, 0);
  std::vector<int32_t> par, goal_slot(n);
  std::vector<double> pc, path_cost(n);
  std::vector<uint8_t> solved(n);
  for (uint32_t p = 0; p < n; p++)
  {
    const Problem& pr = P_[p];
    first[p + 1] = first[p] + pr.tree_size;
    for (uint32_t sl = 0; sl < pr.tree_size; sl++)
    {
      const int32_t id = pr.id_of_slot[sl];
      par.push_back(pr.parent[id] < 0 ? -1 : pr.slot_of_id[pr.parent[id]]);
      pc.push_back(pr.pcost[id]);
    }
    goal_slot[p] = pr.slot_of_id[1];
    solved[p] = pr.solved ? 1 : 0;
    path_cost[p] = pr.path_cost;
  }
  if (gc_forest_create(store_, world_, &fc, cfg_.lb.data(), cfg_.ub.data(), n, starts, goals, first.data(), par.data(),
                       pc.data(), goal_slot.data(), solved.data(), path_cost.data(), &forest_) != GC_OK)
  {
    forest_ = nullptr;
    return fail("gc_forest_create") ? 0 : -1;
  }
  return 0;
}

LockstepRRTStar::~LockstepRRTStar()
{
  if (forest_)
    gc_forest_destroy(forest_);
  if (store_)
    gc_store_destroy(store_);
}

bool LockstepRRTStar::fail(const char* what)
{
  err_ = std::string(what) + ": " + gc_last_error();
  return false;
}

int LockstepRRTStar::addNode(Problem& pr, const double* conf)
{
  pr.conf.insert(pr.conf.end(), conf, conf + D_);
  pr.parent.push_back(-1);
  pr.pcost.push_back(0.0);
  pr.slot_of_id.push_back(-1);
  return (int)pr.parent.size() - 1;
}

// Tree::costToNode (tree.cpp:767-789)
double LockstepRRTStar::costToNode(const Problem& pr, int32_t node) const
{
  double cost = 0;
  while (node != 0)
  {
    if (node < 0 || pr.parent[node] < 0)
      return kInf;
    cost += pr.pcost[node];
    node = pr.parent[node];
  }
  return cost;
}

// InformedSampler::setCost (informed_sampler.cpp:181-215), scale = 1
void LockstepRRTStar::setSamplerCost(Problem& pr, double cost)
{
  double c = cost;
  const bool inf_cost = c >= kInf;
  double min_radius;
  if (c < pr.focii_distance)
  {
    c = pr.focii_distance;
    min_radius = 0.0;
  }
  else
    min_radius = 0.5 * std::sqrt(std::pow(c, 2) - std::pow(pr.focii_distance, 2));
  const double max_radius = 0.5 * c;
  if (inf_cost)
  {
    double v = std::tgamma(((double)D_) * 0.5 + 1.0) / std::pow(M_PI, (double)D_ * 0.5);
    for (int i = 0; i < D_; i++)
      v *= (cfg_.ub[i] - cfg_.lb[i]);
    pr.specific_volume = v;
  }
  else
    pr.specific_volume = max_radius * std::pow(min_radius, D_ - 1);
}

// Path(getConnectionToNode(goal)) (tree.cpp:791-813, path.cpp:34-39): the Path holds the Connection
// objects, whose costs never change afterwards -> freeze them
void LockstepRRTStar::makeSolution(Problem& pr)
{
  pr.solution.clear();
  int32_t t = 1;
  while (t != 0)
  {
    pr.solution.push_back(t);
    t = pr.parent[t];
  }
  std::reverse(pr.solution.begin(), pr.solution.end());
  pr.sol_costs.clear();
  for (int32_t id : pr.solution)
    pr.sol_costs.push_back(pr.pcost[id]);
}

// Path::cost() -> computeCost (path.cpp:285-291): root-to-goal summation order
double LockstepRRTStar::solutionCost(const Problem& pr) const
{
  double c = 0;
  for (double v : pr.sol_costs)
    c += v;
  return c;
}

// TreeSolver::setProblem's direct connection attempt: Tree::connectToNode(goal) (tree.cpp:304-327),
// i.e. repeated Tree::extendToNode (tree.cpp:194-215), all problems in lock step
int LockstepRRTStar::initialConnect()
{
  const uint32_t n = (uint32_t)P_.size();
  std::vector<uint32_t> act(n);
  for (uint32_t p = 0; p < n; p++)
    act[p] = p;
  std::vector<double> q, next, d;
  std::vector<uint32_t> seg, closest, slots;
  std::vector<uint8_t> ok;
  while (!act.empty())
  {
    const uint32_t m = (uint32_t)act.size();
    q.resize((size_t)m * D_);
    next.resize((size_t)m * D_);
    d.resize(m);
    seg.resize(m);
    closest.resize(m);
    ok.resize(m);
    for (uint32_t i = 0; i < m; i++)
    {
      std::memcpy(&q[(size_t)i * D_], &P_[act[i]].conf[D_], sizeof(double) * D_);
      seg[i] = act[i];
    }
    if (gc_extend_batch(store_, world_, q.data(), seg.data(), m, cfg_.max_distance, cfg_.min_distance, TOLERANCE,
                        nullptr, 0, closest.data(), d.data(), next.data(), ok.data(), nullptr, nullptr, nullptr) < 0)
      return fail("gc_extend_batch") ? 0 : -1;
    stats_.gpu_calls++;
    std::vector<uint32_t> keep, app_seg;
    std::vector<double> app_rows;
    for (uint32_t i = 0; i < m; i++)
    {
      Problem& pr = P_[act[i]];
      if (!ok[i])
        continue;  // blocked: no direct solution (tree_solver.cpp:98-101)
      const double* nx = &next[(size_t)i * D_];
      const int32_t par = pr.id_of_slot[closest[i]];
      int32_t id;
      const bool reached = dist(nx, &pr.conf[D_], D_) < TOLERANCE;
      if (reached)
        id = 1;  // the goal node itself joins the tree
      else
        id = addNode(pr, nx);
      pr.parent[id] = par;
      pr.pcost[id] = dist(&pr.conf[(size_t)par * D_], &pr.conf[(size_t)id * D_], D_);
      pr.slot_of_id[id] = (int32_t)pr.tree_size;
      pr.id_of_slot.push_back(id);
      pr.tree_size++;
      app_seg.push_back(pr.seg);
      app_rows.insert(app_rows.end(), &pr.conf[(size_t)id * D_], &pr.conf[(size_t)id * D_] + D_);
      if (reached)
      {
        makeSolution(pr);
        pr.path_cost = solutionCost(pr);
        setSamplerCost(pr, pr.path_cost);
        pr.solved = true;
        pr.cost = pr.path_cost + pr.goal_cost;
      }
      else
        keep.push_back(act[i]);
    }
    if (!app_seg.empty())
    {
      slots.resize(app_seg.size());
      if (gc_store_append_multi(store_, app_seg.data(), app_rows.data(), (uint32_t)app_seg.size(), slots.data()) != GC_OK)
        return fail("gc_store_append_multi") ? 0 : -1;
      stats_.gpu_calls++;
    }
    act.swap(keep);
  }
  return 0;
}

int LockstepRRTStar::step(const double* samples) { return stepImpl(samples, nullptr); }
int LockstepRRTStar::stepDeviceSamples(const double* d_samples) { return stepImpl(nullptr, d_samples); }

int LockstepRRTStar::setStream(void* cuda_stream)
{
  if (gc_store_set_stream(store_, cuda_stream) != GC_OK)
    return fail("gc_store_set_stream") ? 0 : -1;
  return 0;
}

static inline double secs_since(std::chrono::steady_clock::time_point t0)
{
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int LockstepRRTStar::stepInformed(uint64_t seed, uint64_t first_index)
{
  if (!forest_)
  {
    err_ = "stepInformed needs device trees";
    return -1;
  }
  if (gc_forest_step(forest_, nullptr, nullptr, seed, first_index) != GC_OK)
    return fail("gc_forest_step") ? 0 : -1;
  return (int)P_.size();
}

int LockstepRRTStar::sync()
{
  if (!forest_)
    return 0;
  if (gc_forest_sync(forest_) != GC_OK)
    return fail("gc_forest_sync") ? 0 : -1;
  return 0;
}

int LockstepRRTStar::readCostsAsync(double* costs)
{
  if (!forest_)
  {
    err_ = "readCostsAsync needs device trees";
    return -1;
  }
  if (gc_forest_read_costs_async(forest_, costs) != GC_OK)
    return fail("gc_forest_read_costs_async") ? 0 : -1;
  return 0;
}

// accessors: the host mirror, or the forest's snapshot (synchronises when steps are in flight)
bool LockstepRRTStar::solved(uint32_t p) const
{
  if (!forest_)
    return P_[p].solved;
  int32_t v = 0;
  gc_forest_info(forest_, p, &v, nullptr, nullptr, nullptr, nullptr, nullptr);
  return v != 0;
}
bool LockstepRRTStar::canImprove(uint32_t p) const
{
  if (!forest_)
    return P_[p].can_improve;
  int32_t v = 0;
  gc_forest_info(forest_, p, nullptr, &v, nullptr, nullptr, nullptr, nullptr);
  return v != 0;
}
double LockstepRRTStar::cost(uint32_t p) const
{
  if (!forest_)
    return P_[p].cost;
  double v = 0;
  gc_forest_info(forest_, p, nullptr, nullptr, &v, nullptr, nullptr, nullptr);
  return v;
}
double LockstepRRTStar::radius(uint32_t p) const
{
  if (!forest_)
    return P_[p].r_rewire;
  double v = 0;
  gc_forest_info(forest_, p, nullptr, nullptr, nullptr, &v, nullptr, nullptr);
  return v;
}
uint32_t LockstepRRTStar::numNodes(uint32_t p) const
{
  if (!forest_)
    return (uint32_t)P_[p].parent.size();
  uint32_t v = 0;
  gc_forest_info(forest_, p, nullptr, nullptr, nullptr, nullptr, &v, nullptr);
  return v;
}
uint32_t LockstepRRTStar::treeSize(uint32_t p) const
{
  if (!forest_)
    return P_[p].tree_size;
  uint32_t v = 0;
  gc_forest_info(forest_, p, nullptr, nullptr, nullptr, nullptr, nullptr, &v);
  return v;
}
const StepStats& LockstepRRTStar::stats() const
{
  if (forest_)
  {
    uint64_t o[16] = {};
    gc_forest_stats(forest_, o);
    stats_.iterations = o[0];
    stats_.extended = o[1];
    stats_.edges_device = o[2];
    stats_.edges_used = o[3];
    stats_.edges_fallback = o[4];
    stats_.near_hits = o[5];
  }
  return stats_;
}

int LockstepRRTStar::stepImpl(const double* samples, const double* d_samples)
{
  if (!store_)
    return -1;
  if (forest_)
  {
    // device trees: everything is enqueued on the store's stream; nothing comes back here
    if (gc_forest_step(forest_, samples, d_samples, 0, 0) != GC_OK)
      return fail("gc_forest_step") ? 0 : -1;
    stats_.gpu_calls++;
    if (samples)
      stats_.h2d_bytes += sizeof(double) * (uint64_t)P_.size() * D_;
    stats_.h2d_bytes += sizeof(double) * (uint64_t)P_.size();           // rewire-radius factors
    stats_.d2h_bytes += (sizeof(double) + 1) * (uint64_t)P_.size();     // path costs + changed flags
    return (int)P_.size();
  }
  auto tp = std::chrono::steady_clock::now();
  const uint32_t n = (uint32_t)P_.size();
  // ---- A: RRTStar::update prologue (rrt_star.cpp:93-108) ----------------------------------------
  uint32_t rows = 0;
  q_.resize((size_t)n * D_);
  seg_.resize(n);
  radius_.resize(n);
  std::vector<uint32_t> act;
  act.reserve(n);
  for (uint32_t p = 0; p < n; p++)
  {
    Problem& pr = P_[p];
    pr.active = false;
    if (!pr.can_improve)
      continue;
    if (pr.cost <= utopia_tol_ * pr.best_utopia)
    {
      pr.can_improve = false;  // "Solution already optimal"
      continue;
    }
    // updateRewireRadius (rrt_star.cpp:278-286)
    const double dimension = (double)D_;
    const double r_rrt = cfg_.rewire_factor * std::pow(2.0 * (1.0 + 1.0 / dimension), 1.0 / dimension) *
                         std::pow(pr.specific_volume, 1.0 / dimension);
    const double cardDbl = pr.tree_size + 1.0;
    pr.r_rewire = (r_rrt * std::pow(std::log(cardDbl) / cardDbl, 1.0 / dimension));
    pr.active = true;
    pr.row = rows;
    if (samples)
      std::memcpy(&q_[(size_t)rows * D_], samples + (size_t)p * D_, sizeof(double) * D_);
    seg_[rows] = pr.seg;
    radius_[rows] = pr.r_rewire;
    act.push_back(p);
    rows++;
  }
  if (rows == 0)
    return 0;
  stats_.t_phase[0] += secs_since(tp);
  tp = std::chrono::steady_clock::now();
  stats_.iterations += rows;
  // ---- B: nearest + steer + edge check + radius-near, one stream-ordered device pass ------------
  closest_.resize(rows);
  dist_.resize(rows);
  next_.resize((size_t)rows * D_);
  ok_.resize(rows);
  near_count_.resize(rows);
  for (;;)
  {
    near_idx_.resize((size_t)rows * near_cap_);
    near_dist_.resize((size_t)rows * near_cap_);
    const int rc =
        samples ? gc_extend_batch(store_, world_, q_.data(), seg_.data(), rows, cfg_.max_distance, cfg_.min_distance,
                                  TOLERANCE, radius_.data(), near_cap_, closest_.data(), dist_.data(), next_.data(),
                                  ok_.data(), near_idx_.data(), near_dist_.data(), near_count_.data())
                : gc_extend_batch_dsamples(store_, world_, d_samples, rows == n ? nullptr : act.data(), seg_.data(),
                                           rows, cfg_.max_distance, cfg_.min_distance, TOLERANCE, radius_.data(),
                                           near_cap_, closest_.data(), dist_.data(), next_.data(), ok_.data(),
                                           near_idx_.data(), near_dist_.data(), near_count_.data());
    stats_.gpu_calls++;
    if (rc < 0)
      return fail("gc_extend_batch") ? 0 : -1;
    // samples (host path only) + segment ids + radii in; closest, dist, next, ok and the near lists out
    stats_.h2d_bytes += (uint64_t)rows * ((samples ? sizeof(double) * D_ : 0) + sizeof(uint32_t) + sizeof(double));
    {
      // the library fetches only as many near-list columns as the longest list needs (about)
      uint32_t mxc = 0;
      for (uint32_t i = 0; i < rows; i++)
        mxc = std::max(mxc, std::min(near_count_[i], near_cap_));
      stats_.d2h_bytes +=
          (uint64_t)rows * (sizeof(uint32_t) + sizeof(double) + sizeof(double) * D_ + 1 + sizeof(uint32_t)) +
          (uint64_t)rows * mxc * (sizeof(uint32_t) + sizeof(double));
    }
    if (rc != GC_W_TRUNCATED)
      break;
    uint32_t mx = 0;
    for (uint32_t i = 0; i < rows; i++)
      mx = std::max(mx, near_count_[i]);
    while (near_cap_ < mx)
      near_cap_ *= 2;
  }
  stats_.t_phase[1] += secs_since(tp);
  tp = std::chrono::steady_clock::now();
  // ---- C: extension bookkeeping + speculative edge sets (per problem, in parallel) --------------
  pool_->parallel_for(rows, [&](uint32_t r) {
    Problem& pr = P_[act[r]];
    pr.extended = false;
    pr.par_cand.clear();
    pr.chi_cand.clear();
    pr.goal_edge = -1;
    pr.edge_count = 0;
    if (!ok_[r])
      return;  // Tree::extend failed (tree.cpp:152-157)
    // Tree::extend -> extendOnly (tree.cpp:130-140, 159-160)
    const double* nx = &next_[(size_t)r * D_];
    const int32_t closest = pr.id_of_slot[closest_[r]];
    const int32_t id = addNode(pr, nx);
    pr.parent[id] = closest;
    pr.pcost[id] = dist(&pr.conf[(size_t)closest * D_], &pr.conf[(size_t)id * D_], D_);
    pr.slot_of_id[id] = (int32_t)pr.tree_size;
    pr.id_of_slot.push_back(id);
    pr.tree_size++;
    pr.new_node = id;
    pr.extended = true;
    // rewireOnly (tree.cpp:419-513), parent phase (:428-463): every near node that would lower the cost-to-come of
    // the new node is a candidate; the reference walks them in near order and keeps the best collision-free one.
    pr.near_n = std::min(near_count_[r], near_cap_);
    const uint32_t* nidx = &near_idx_[(size_t)r * near_cap_];
    const double* ndist = &near_dist_[(size_t)r * near_cap_];
    const double c2n0 = costToNode(pr, id);
    pr.cost_to_node0 = c2n0;
    pr.par_total.clear();
    pr.par_known = 0;
    pr.par_winner = -1;
    // costToNode walks the parent chain (tree.cpp:767-789): once per near node and step; the values stay valid until
    // the children phase of the replay rewires something
    pr.c2near.resize(pr.near_n);
    for (uint32_t i = 0; i < pr.near_n; i++)
    {
      const int32_t nn = pr.id_of_slot[nidx[i]];
      pr.c2near[i] = (nn == id) ? c2n0 : costToNode(pr, nn);
    }
    for (uint32_t i = 0; i < pr.near_n; i++)
    {
      const int32_t nn = pr.id_of_slot[nidx[i]];
      if (nn == closest || nn == id)
        continue;
      const double c2near = pr.c2near[i];
      if (c2near >= c2n0)
        continue;
      if ((c2near + ndist[i]) >= c2n0)
        continue;
      pr.par_cand.push_back((int32_t)i);
      pr.par_total.push_back(c2near + ndist[i]);
    }
    const size_t nc = pr.par_cand.size();
    if (nc > 1)
    {
      // by (total, near position): a later candidate replaces the parent only if strictly cheaper
      std::vector<uint32_t>& ord = pr.sort_scratch;
      ord.resize(nc);
      for (size_t k = 0; k < nc; k++)
        ord[k] = (uint32_t)k;
      std::sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) {
        return pr.par_total[a] < pr.par_total[b] || (pr.par_total[a] == pr.par_total[b] && a < b);
      });
      std::vector<int32_t> ci(nc);
      std::vector<double> ct(nc);
      for (size_t k = 0; k < nc; k++)
      {
        ci[k] = pr.par_cand[ord[k]];
        ct[k] = pr.par_total[ord[k]];
      }
      pr.par_cand.swap(ci);
      pr.par_total.swap(ct);
    }
    pr.par_ok.assign(nc, 0);
    pr.par_resolved = (nc == 0);
    pr.cost_to_node1 = c2n0;
    pr.edgeA_count = (uint32_t)std::min<size_t>(nc, first_batch_);
  });
  stats_.t_phase[2] += secs_since(tp);
  tp = std::chrono::steady_clock::now();
  // ---- D: append the new nodes; first edge batch = the best parent candidates of every problem ---
  uint32_t nedgesA = 0, napp = 0;
  for (uint32_t r = 0; r < rows; r++)
  {
    Problem& pr = P_[act[r]];
    pr.edgeA_first = nedgesA;
    if (pr.extended)
    {
      nedgesA += pr.edgeA_count;
      napp++;
    }
    else
      pr.edgeA_count = 0;
  }
  if (napp)
  {
    app_seg_.resize(napp);
    app_slots_.resize(napp);
    app_rows_.resize((size_t)napp * D_);
    uint32_t k = 0;
    for (uint32_t r = 0; r < rows; r++)
    {
      Problem& pr = P_[act[r]];
      if (!pr.extended)
        continue;
      app_seg_[k] = pr.seg;
      std::memcpy(&app_rows_[(size_t)k * D_], &pr.conf[(size_t)pr.new_node * D_], sizeof(double) * D_);
      k++;
    }
    if (gc_store_append_multi(store_, app_seg_.data(), app_rows_.data(), napp, app_slots_.data()) != GC_OK)
      return fail("gc_store_append_multi") ? 0 : -1;
    stats_.gpu_calls++;
    stats_.extended += napp;
    stats_.h2d_bytes += (uint64_t)napp * (sizeof(double) * D_ + sizeof(uint32_t)) + sizeof(uint32_t) * n;
  }
  auto run_edges = [&](uint32_t count) -> bool {
    if (gc_check_edges(world_, e1_.data(), e2_.data(), count, cfg_.min_distance, GC_EDGE_BISECT, eok_.data(), nullptr) !=
        GC_OK)
      return false;
    stats_.gpu_calls++;
    stats_.edges_device += count;
    stats_.h2d_bytes += (uint64_t)count * 2 * sizeof(double) * D_;
    stats_.d2h_bytes += count;
    return true;
  };
  if (nedgesA)
  {
    e1_.resize((size_t)nedgesA * D_);
    e2_.resize((size_t)nedgesA * D_);
    eok_.resize(nedgesA);
    pool_->parallel_for(rows, [&](uint32_t r) {
      Problem& pr = P_[act[r]];
      if (!pr.edgeA_count)
        return;
      const uint32_t* nidx = &near_idx_[(size_t)r * near_cap_];
      const double* me = &pr.conf[(size_t)pr.new_node * D_];
      for (uint32_t k = 0; k < pr.edgeA_count; k++)  // n -> node
      {
        const int32_t nn = pr.id_of_slot[nidx[pr.par_cand[k]]];
        std::memcpy(&e1_[(size_t)(pr.edgeA_first + k) * D_], &pr.conf[(size_t)nn * D_], sizeof(double) * D_);
        std::memcpy(&e2_[(size_t)(pr.edgeA_first + k) * D_], me, sizeof(double) * D_);
      }
    });
    if (!run_edges(nedgesA))
      return fail("gc_check_edges") ? 0 : -1;
  }
  // ---- D2: resolve what the first batch decided; second batch = the remaining parent candidates of the problems
  // it left open + the children-phase candidates (tree.cpp:471-510) + the goal edge ------------------------------
  pool_->parallel_for(rows, [&](uint32_t r) {
    Problem& pr = P_[act[r]];
    pr.chi_cand.clear();
    pr.goal_edge = -1;
    pr.edge_count = 0;
    pr.parB_count = 0;
    if (!pr.extended)
      return;
    const uint32_t* nidx = &near_idx_[(size_t)r * near_cap_];
    const double* ndist = &near_dist_[(size_t)r * near_cap_];
    const int32_t id = pr.new_node;
    for (uint32_t k = 0; k < pr.edgeA_count; k++)
    {
      pr.par_ok[k] = eok_[pr.edgeA_first + k];
      if (pr.par_ok[k] && pr.par_winner < 0)
        pr.par_winner = (int32_t)k;
    }
    pr.par_known = pr.edgeA_count;
    const size_t nc = pr.par_cand.size();
    if (pr.par_winner >= 0)
    {
      pr.par_resolved = true;
      pr.cost_to_node1 = pr.par_total[pr.par_winner];
    }
    else if (pr.par_known == nc)
      pr.par_resolved = true;  // nothing free: the nearest node stays the parent
    else
    {
      pr.parB_count = (uint32_t)(nc - pr.par_known);
      pr.cost_to_node1 = pr.par_total[pr.par_known];  // the best the open candidates can still reach
    }
    // children phase: exact cost-to-come when resolved, its lower bound otherwise (a superset either way: the
    // cost of a near node only falls while the phase runs)
    const double c2n1 = pr.cost_to_node1;
    for (uint32_t i = 0; i < pr.near_n; i++)
    {
      const int32_t nn = pr.id_of_slot[nidx[i]];
      if (nn == id || nn == 0)
        continue;
      const double c2near = pr.c2near[i];
      if (c2n1 >= c2near)
        continue;
      if ((c2n1 + ndist[i]) >= c2near)
        continue;
      pr.chi_cand.push_back((int32_t)i);
    }
    pr.edge_count = pr.parB_count + (uint32_t)pr.chi_cand.size();
    if (!pr.solved && dist(&pr.conf[(size_t)id * D_], &pr.conf[D_], D_) < cfg_.max_distance)
    {
      pr.goal_edge = (int32_t)pr.edge_count;
      pr.edge_count++;
    }
  });
  uint32_t nedges = 0;
  for (uint32_t r = 0; r < rows; r++)
  {
    Problem& pr = P_[act[r]];
    pr.edge_first = nedges;
    nedges += pr.edge_count;
  }
  if (nedges)
  {
    e1_.resize((size_t)nedges * D_);
    e2_.resize((size_t)nedges * D_);
    eok_.resize(nedges);
    pool_->parallel_for(rows, [&](uint32_t r) {
      Problem& pr = P_[act[r]];
      if (!pr.edge_count)
        return;
      const uint32_t* nidx = &near_idx_[(size_t)r * near_cap_];
      const double* me = &pr.conf[(size_t)pr.new_node * D_];
      uint32_t e = pr.edge_first;
      for (uint32_t k = 0; k < pr.parB_count; k++)  // n -> node
      {
        const int32_t nn = pr.id_of_slot[nidx[pr.par_cand[pr.par_known + k]]];
        std::memcpy(&e1_[(size_t)e * D_], &pr.conf[(size_t)nn * D_], sizeof(double) * D_);
        std::memcpy(&e2_[(size_t)e * D_], me, sizeof(double) * D_);
        e++;
      }
      for (int32_t i : pr.chi_cand)  // node -> n
      {
        const int32_t nn = pr.id_of_slot[nidx[i]];
        std::memcpy(&e1_[(size_t)e * D_], me, sizeof(double) * D_);
        std::memcpy(&e2_[(size_t)e * D_], &pr.conf[(size_t)nn * D_], sizeof(double) * D_);
        e++;
      }
      if (pr.goal_edge >= 0)  // node -> goal
      {
        std::memcpy(&e1_[(size_t)e * D_], me, sizeof(double) * D_);
        std::memcpy(&e2_[(size_t)e * D_], &pr.conf[D_], sizeof(double) * D_);
      }
    });
    if (!run_edges(nedges))
      return fail("gc_check_edges") ? 0 : -1;
  }
  stats_.t_phase[3] += secs_since(tp);
  tp = std::chrono::steady_clock::now();
  // ---- E: sequential replay of rewireOnly + the solver epilogue (per problem, in parallel) -------
  std::vector<uint8_t> goal_joined(rows, 0);
  std::vector<uint32_t> used(rows, 0), fallback(rows, 0), hits(rows, 0);
  std::mutex fb_mutex;
  pool_->parallel_for(rows, [&](uint32_t r) {
    Problem& pr = P_[act[r]];
    if (!pr.extended)
      return;  // update() returns false
    const uint32_t* nidx = &near_idx_[(size_t)r * near_cap_];
    const double* ndist = &near_dist_[(size_t)r * near_cap_];
    const int32_t node = pr.new_node;
    hits[r] = pr.near_n;
    auto edge_result = [&](const std::vector<int32_t>& cand, size_t& cursor, uint32_t base, int32_t i,
                           const double* a, const double* b) -> bool {
      while (cursor < cand.size() && cand[cursor] < i)
        cursor++;
      used[r]++;
      if (cursor < cand.size() && cand[cursor] == i)
        return eok_[pr.edge_first + base + cursor] != 0;
      // not in the speculative set (cannot happen by construction; kept as a safety net)
      std::lock_guard<std::mutex> lk(fb_mutex);
      uint8_t res = 0;
      gc_check_edges(world_, a, b, 1, cfg_.min_distance, GC_EDGE_BISECT, &res, nullptr);
      fallback[r]++;
      return res != 0;
    };
    double cost_to_node = pr.cost_to_node0;
    bool improved = false;
    // parent phase (tree.cpp:428-463): the first collision-free candidate in (cost, near position) order is where the
    // reference's walk ends up
    if (!pr.par_resolved)
    {
      for (uint32_t k = 0; k < pr.parB_count; k++)
      {
        pr.par_ok[pr.par_known + k] = eok_[pr.edge_first + k];
        if (pr.par_ok[pr.par_known + k] && pr.par_winner < 0)
          pr.par_winner = (int32_t)(pr.par_known + k);
      }
      pr.par_known += pr.parB_count;
    }
    used[r] += (pr.par_winner >= 0) ? (uint32_t)pr.par_winner + 1u : pr.par_known;
    if (pr.par_winner >= 0)
    {
      const int32_t i = pr.par_cand[pr.par_winner];
      pr.parent[node] = pr.id_of_slot[nidx[i]];
      pr.pcost[node] = ndist[i];  // metrics_->cost(n, node) == the exact fp64 near distance
      cost_to_node = pr.par_total[pr.par_winner];
      improved = true;
    }
    size_t cur = 0;
    const int32_t par = pr.parent[node];
    bool rewired = false;
    cur = 0;
    for (uint32_t i = 0; i < pr.near_n; i++)  // children phase (tree.cpp:471-510)
    {
      const int32_t nn = pr.id_of_slot[nidx[i]];
      if (nn == par || nn == node || nn == 0)
        continue;
      const double c2near = rewired ? costToNode(pr, nn) : pr.c2near[i];
      if (cost_to_node >= c2near)
        continue;
      const double c = ndist[i];
      if ((cost_to_node + c) >= c2near)
        continue;
      if (!edge_result(pr.chi_cand, cur, pr.parB_count, (int32_t)i, &pr.conf[(size_t)node * D_],
                       &pr.conf[(size_t)nn * D_]))
        continue;
      pr.parent[nn] = node;
      pr.pcost[nn] = c;
      improved = true;
      rewired = true;  // descendants of nn just got cheaper: walk from here on
    }
    // RRTStar::update epilogue (rrt_star.cpp:110-162); Tree::rewire returns rewireOnly's `improved`
    if (!pr.solved)
    {
      if (!improved)
        return;
      if (dist(&pr.conf[(size_t)node * D_], &pr.conf[D_], D_) < cfg_.max_distance)
      {
        bool free_edge;
        used[r]++;
        if (pr.goal_edge >= 0)
          free_edge = eok_[pr.edge_first + pr.goal_edge] != 0;
        else
        {
          std::lock_guard<std::mutex> lk(fb_mutex);
          uint8_t res = 0;
          gc_check_edges(world_, &pr.conf[(size_t)node * D_], &pr.conf[D_], 1, cfg_.min_distance, GC_EDGE_BISECT, &res,
                         nullptr);
          fallback[r]++;
          free_edge = res != 0;
        }
        if (free_edge)
        {
          pr.parent[1] = node;
          pr.pcost[1] = dist(&pr.conf[(size_t)node * D_], &pr.conf[D_], D_);
          makeSolution(pr);
          pr.slot_of_id[1] = (int32_t)pr.tree_size;  // start_tree_->addNode(goal_node_)
          pr.id_of_slot.push_back(1);
          pr.tree_size++;
          goal_joined[r] = 1;
          pr.path_cost = solutionCost(pr);
          pr.cost = pr.path_cost + pr.goal_cost;
          setSamplerCost(pr, pr.path_cost);
          pr.solved = true;
        }
      }
    }
    else if (improved)
    {
      if (costToNode(pr, 1) >= (solutionCost(pr) - 1e-8))
        return;
      makeSolution(pr);
      pr.path_cost = solutionCost(pr);
      pr.cost = pr.path_cost + pr.goal_cost;
      setSamplerCost(pr, pr.path_cost);
    }
  });
  // goal nodes that joined a tree this step enter the device store too
  uint32_t ng = 0;
  for (uint32_t r = 0; r < rows; r++)
  {
    stats_.edges_used += used[r];
    stats_.edges_fallback += fallback[r];
    stats_.near_hits += hits[r];
    ng += goal_joined[r];
  }
  if (ng)
  {
    app_seg_.resize(ng);
    app_slots_.resize(ng);
    app_rows_.resize((size_t)ng * D_);
    uint32_t k = 0;
    for (uint32_t r = 0; r < rows; r++)
      if (goal_joined[r])
      {
        Problem& pr = P_[act[r]];
        app_seg_[k] = pr.seg;
        std::memcpy(&app_rows_[(size_t)k * D_], &pr.conf[D_], sizeof(double) * D_);
        k++;
      }
    if (gc_store_append_multi(store_, app_seg_.data(), app_rows_.data(), ng, app_slots_.data()) != GC_OK)
      return fail("gc_store_append_multi") ? 0 : -1;
    stats_.gpu_calls++;
  }
  stats_.t_phase[4] += secs_since(tp);
  return (int)rows;
}

void LockstepRRTStar::dump(uint32_t p, int32_t* parents, double* pcost, double* conf) const
{
  if (forest_)
  {
    gc_forest_dump(forest_, p, parents, pcost, conf);
    return;
  }
  const Problem& pr = P_[p];
  const size_t n = pr.parent.size();
  std::memcpy(parents, pr.parent.data(), sizeof(int32_t) * n);
  std::memcpy(pcost, pr.pcost.data(), sizeof(double) * n);
  std::memcpy(conf, pr.conf.data(), sizeof(double) * n * D_);
}

std::vector<int32_t> LockstepRRTStar::solution(uint32_t p) const
{
  std::vector<int32_t> out;
  if (forest_)
  {
    out.resize(1);
    const int n = gc_forest_solution(forest_, p, out.data(), 1);
    if (n <= 0)
      return std::vector<int32_t>();
    out.resize((size_t)n);
    gc_forest_solution(forest_, p, out.data(), (uint32_t)n);
    return out;
  }
  if (!P_[p].solved)
    return out;
  out.push_back(0);
  out.insert(out.end(), P_[p].solution.begin(), P_[p].solution.end());
  return out;
}

}  // namespace gcb200

// =================================================================================================
//  C interface
// =================================================================================================
using gcb200::LockstepRRTStar;
static thread_local std::string g_perr;

struct gcp_rrtstar
{
  std::unique_ptr<LockstepRRTStar> impl;
};

extern "C" {

const char* gcp_last_error(void) { return g_perr.c_str(); }

int gcp_rrtstar_create(int dim, double max_distance, double min_distance, double rewire_factor,
                       double utopia_tolerance, const double* lb, const double* ub, gc_world_t* world, int device,
                       uint32_t nproblems, const double* starts, const double* goals, uint32_t capacity_per_tree,
                       int threads, gcp_rrtstar_t** out)
{
  return gcp_rrtstar_create2(dim, max_distance, min_distance, rewire_factor, utopia_tolerance, lb, ub, world, device,
                             nproblems, starts, goals, capacity_per_tree, threads, 1, out);
}

int gcp_rrtstar_step_informed(gcp_rrtstar_t* s, uint64_t seed, uint64_t first_index)
{
  if (!s)
    return GC_E_INVALID;
  const int rc = s->impl->stepInformed(seed, first_index);
  if (rc < 0)
    g_perr = s->impl->error();
  return rc;
}

int gcp_rrtstar_sync(gcp_rrtstar_t* s)
{
  if (!s)
    return GC_E_INVALID;
  const int rc = s->impl->sync();
  if (rc < 0)
    g_perr = s->impl->error();
  return rc;
}

int gcp_rrtstar_read_costs_async(gcp_rrtstar_t* s, double* costs)
{
  if (!s || !costs)
    return GC_E_INVALID;
  const int rc = s->impl->readCostsAsync(costs);
  if (rc < 0)
    g_perr = s->impl->error();
  return rc;
}

int gcp_rrtstar_create2(int dim, double max_distance, double min_distance, double rewire_factor,
                        double utopia_tolerance, const double* lb, const double* ub, gc_world_t* world, int device,
                        uint32_t nproblems, const double* starts, const double* goals, uint32_t capacity_per_tree,
                        int threads, int device_trees, gcp_rrtstar_t** out)
{
  if (!out || !lb || !ub || !world || !starts || !goals || nproblems == 0 || dim < 1 || dim > GC_MAX_DIM)
  {
    g_perr = "gcp_rrtstar_create: invalid argument";
    return GC_E_INVALID;
  }
  gcb200::SolverConfig cfg;
  cfg.dim = dim;
  cfg.max_distance = max_distance;
  cfg.min_distance = min_distance;
  cfg.rewire_factor = rewire_factor;
  cfg.utopia_tolerance = utopia_tolerance;
  cfg.lb.assign(lb, lb + dim);
  cfg.ub.assign(ub, ub + dim);
  if (threads <= 0)
    threads = (int)std::max(1u, std::thread::hardware_concurrency());
  gcp_rrtstar* s = new gcp_rrtstar();
  s->impl.reset(new LockstepRRTStar(cfg, world, device, nproblems, starts, goals, capacity_per_tree, threads,
                                    device_trees != 0));
  if (s->impl->error()[0] != 0)
  {
    g_perr = s->impl->error();
    delete s;
    *out = nullptr;
    return GC_E_CUDA;
  }
  *out = s;
  return GC_OK;
}

int gcp_rrtstar_destroy(gcp_rrtstar_t* s)
{
  delete s;
  return GC_OK;
}

int gcp_rrtstar_step(gcp_rrtstar_t* s, const double* samples)
{
  if (!s || !samples)
    return GC_E_INVALID;
  const int rc = s->impl->step(samples);
  if (rc < 0)
    g_perr = s->impl->error();
  return rc;
}

int gcp_rrtstar_step_dsamples(gcp_rrtstar_t* s, const double* d_samples)
{
  if (!s || !d_samples)
    return GC_E_INVALID;
  const int rc = s->impl->stepDeviceSamples(d_samples);
  if (rc < 0)
    g_perr = s->impl->error();
  return rc;
}

int gcp_rrtstar_set_stream(gcp_rrtstar_t* s, void* cuda_stream)
{
  if (!s)
    return GC_E_INVALID;
  const int rc = s->impl->setStream(cuda_stream);
  if (rc < 0)
    g_perr = s->impl->error();
  return rc;
}

int gcp_rrtstar_run(gcp_rrtstar_t* s, const double* samples, uint32_t n)
{
  if (!s || !samples)
    return GC_E_INVALID;
  const size_t stride = (size_t)s->impl->size() * (size_t)gc_store_dim(s->impl->store());
  long total = 0;
  for (uint32_t i = 0; i < n; i++)
  {
    const int rc = s->impl->step(samples + (size_t)i * stride);
    if (rc < 0)
    {
      g_perr = s->impl->error();
      return rc;
    }
    total += rc;
    if (rc == 0)
      break;
  }
  if (s->impl->deviceTrees() && s->impl->sync() < 0)
  {
    g_perr = s->impl->error();
    return GC_E_CAPACITY;
  }
  return (int)std::min<long>(total, 0x7fffffff);
}

int gcp_rrtstar_info(gcp_rrtstar_t* s, uint32_t p, int32_t* solved, int32_t* can_improve, double* cost,
                     double* radius, uint32_t* num_nodes, uint32_t* tree_size)
{
  if (!s || p >= s->impl->size())
    return GC_E_INVALID;
  if (solved) *solved = s->impl->solved(p) ? 1 : 0;
  if (can_improve) *can_improve = s->impl->canImprove(p) ? 1 : 0;
  if (cost) *cost = s->impl->cost(p);
  if (radius) *radius = s->impl->radius(p);
  if (num_nodes) *num_nodes = s->impl->numNodes(p);
  if (tree_size) *tree_size = s->impl->treeSize(p);
  return GC_OK;
}

int gcp_rrtstar_dump(gcp_rrtstar_t* s, uint32_t p, int32_t* parents, double* pcost, double* conf)
{
  if (!s || p >= s->impl->size() || !parents || !pcost || !conf)
    return GC_E_INVALID;
  s->impl->dump(p, parents, pcost, conf);
  return GC_OK;
}

int gcp_rrtstar_solution(gcp_rrtstar_t* s, uint32_t p, int32_t* ids, uint32_t cap)
{
  if (!s || p >= s->impl->size())
    return GC_E_INVALID;
  const std::vector<int32_t> sol = s->impl->solution(p);
  for (size_t i = 0; i < sol.size() && i < cap; i++)
    ids[i] = sol[i];
  return (int)sol.size();
}

int gcp_rrtstar_stats(gcp_rrtstar_t* s, uint64_t out[16])
{
  if (!s || !out)
    return GC_E_INVALID;
  const gcb200::StepStats& st = s->impl->stats();
  out[0] = st.iterations;
  out[1] = st.extended;
  out[2] = st.edges_device;
  out[3] = st.edges_used;
  out[4] = st.edges_fallback;
  out[5] = st.near_hits;
  out[6] = st.gpu_calls;
  out[7] = st.h2d_bytes;
  out[8] = st.d2h_bytes;
  out[9] = 0;
  for (int i = 0; i < 5; i++)
    out[10 + i] = (uint64_t)(st.t_phase[i] * 1e9);
  return GC_OK;
}

}  // extern "C"
