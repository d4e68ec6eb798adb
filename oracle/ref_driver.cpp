/*
 * ref_driver.cpp -- C entry points over the UNMODIFIED graph_core reference, compiled here from its
 * own sources where they lie under /root/reference (never copied into this repository).
 *
 * TEST INFRASTRUCTURE (oracle/): the resulting oracle/_ref/libgc_ref.so is the executable form of the
 * reference that pins oracle/gc_oracle.cpp (tests/test_oracle_vs_ref.py) and mints the golden vectors
 * under tests/golden/ (tests/golden/make_ref_golden.py).  Nothing under graph_core_b200/ links it.
 *
 * What is the reference's own code here: every graph::core class used below --
 *   KdTree / Vector                         src/graph_core/datastructure/{kdtree,vector}.cpp
 *   Node / Connection                       src/graph_core/graph/{node,connection}.cpp
 *   CollisionCheckerBase::checkConnection*  include/graph_core/collision_checkers/collision_checker_base.h:178-313
 *   Cube3dCollisionChecker                  include/graph_core/collision_checkers/cube_3d_collision_checker.h
 *   Tree, Path, RRT, RRTStar, BiRRT, AnytimeRRT, samplers, EuclideanMetrics
 * What is a stand-in: Eigen, cnr_logger, cnr_param, yaml-cpp (oracle/ref_shim/, see mini_eigen.hpp for
 * the arithmetic conventions) -- this image has none of them.
 *
 * The synthetic worlds of BASELINE.json (boxes, C-space spheres, sphere arm) are not part of the
 * reference; CallbackChecker::check() forwards single states to a caller-supplied function (the
 * oracle's orc_check_states), so the reference's OWN checkConnection / checkConnFromConf loops decide
 * which states are visited and in which order.
 */
#include <graph_core/collision_checkers/cube_3d_collision_checker.h>
// kdtree.h / vector.h arrive through tree.h (vector.h has no include guard upstream: include it once only)
#include <graph_core/metrics/euclidean_metrics.h>
#include <graph_core/samplers/informed_sampler.h>
#include <graph_core/samplers/uniform_sampler.h>
#include <graph_core/solvers/anytime_rrt.h>
#include <graph_core/solvers/birrt.h>
#include <graph_core/solvers/rrt.h>
#include <graph_core/solvers/path_optimizers/path_local_optimizer.h>
#include <graph_core/solvers/rrt_star.h>

#include <cstring>
#include <unordered_map>

using namespace graph::core;

namespace
{
cnr_logger::TraceLoggerPtr g_logger = std::make_shared<cnr_logger::TraceLogger>();

Eigen::VectorXd toVec(const double* q, int d)
{
  Eigen::VectorXd v(d);
  for (int i = 0; i < d; i++)
    v(i) = q[i];
  return v;
}

typedef void (*check_states_fn)(void* ctx, const double* q, int n, unsigned char* ok);

class CallbackChecker : public CollisionCheckerBase
{
public:
  CallbackChecker(check_states_fn fn, void* ctx, double min_distance)
    : CollisionCheckerBase(g_logger, min_distance), fn_(fn), ctx_(ctx)
  {
  }
  bool check(const Eigen::VectorXd& configuration) override
  {
    unsigned char ok = 0;
    fn_(ctx_, configuration.data(), 1, &ok);
    states++;
    return ok != 0;
  }
  CollisionCheckerPtr clone() override { return std::make_shared<CallbackChecker>(fn_, ctx_, min_distance_); }
  unsigned long long states = 0;

private:
  check_states_fn fn_;
  void* ctx_;
};

struct RefNN
{
  int dim;
  NearestNeighborsPtr nn;
  std::vector<NodePtr> nodes;  // id = creation order
  std::unordered_map<Node*, int> id_of;
};

int copyMap(RefNN* o, const std::multimap<double, NodePtr>& m, int* ids, double* dists, int cap)
{
  int n = 0;
  for (const auto& p : m)
  {
    if (n < cap)
    {
      if (ids)
        ids[n] = o->id_of.at(p.second.get());
      if (dists)
        dists[n] = p.first;
    }
    n++;
  }
  return n;
}

struct RefSolver
{
  int kind, dim;
  CollisionCheckerPtr checker;
  MetricsPtr metrics;
  SamplerPtr sampler;
  TreeSolverPtr solver;
  NodePtr start, goal;
  PathPtr solution;
};
}  // namespace

extern "C" {

const char* ref_about()
{
  return "graph_core reference sources compiled against oracle/ref_shim stand-ins (Eigen, cnr_logger, cnr_param, yaml-cpp)";
}

// ---- NearestNeighbors: kind 0 = Vector, 1 = KdTree -----------------------------------------------
void* ref_nn_new(int kind, int dim)
{
  RefNN* o = new RefNN();
  o->dim = dim;
  if (kind == 1)
    o->nn = std::make_shared<KdTree>(g_logger);
  else
    o->nn = std::make_shared<Vector>(g_logger);
  return o;
}
void ref_nn_free(void* h) { delete (RefNN*)h; }
void ref_nn_insert_many(void* h, const double* q, int n)
{
  RefNN* o = (RefNN*)h;
  for (int i = 0; i < n; i++)
  {
    NodePtr node = std::make_shared<Node>(toVec(q + (size_t)i * o->dim, o->dim), g_logger);
    o->id_of[node.get()] = (int)o->nodes.size();
    o->nodes.push_back(node);
    o->nn->insert(node);
  }
}
void ref_nn_reinsert(void* h, int id) { ((RefNN*)h)->nn->insert(((RefNN*)h)->nodes[id]); }
int ref_nn_delete(void* h, int id) { return ((RefNN*)h)->nn->deleteNode(((RefNN*)h)->nodes[id], false) ? 1 : 0; }
int ref_nn_restore(void* h, int id) { return ((RefNN*)h)->nn->restoreNode(((RefNN*)h)->nodes[id]) ? 1 : 0; }
int ref_nn_find(void* h, int id) { return ((RefNN*)h)->nn->findNode(((RefNN*)h)->nodes[id]) ? 1 : 0; }
int ref_nn_clear(void* h) { return ((RefNN*)h)->nn->clear() ? 1 : 0; }
unsigned int ref_nn_size(void* h) { return ((RefNN*)h)->nn->size(); }
void ref_nn_set_deleted_threshold(void* h, unsigned int t)
{
  KdTree* k = dynamic_cast<KdTree*>(((RefNN*)h)->nn.get());
  if (k)
    k->deletedNodesThreshold(t);
}
int ref_nn_get_nodes(void* h, int* ids, int cap)
{
  RefNN* o = (RefNN*)h;
  std::vector<NodePtr> v = o->nn->getNodes();
  for (int i = 0; i < (int)v.size() && i < cap; i++)
    ids[i] = o->id_of.at(v[i].get());
  return (int)v.size();
}
void ref_nn_nearest(void* h, const double* q, int nq, int* ids, double* dists)
{
  RefNN* o = (RefNN*)h;
  for (int i = 0; i < nq; i++)
  {
    NodePtr best;
    double bd = std::numeric_limits<double>::infinity();
    o->nn->nearestNeighbor(toVec(q + (size_t)i * o->dim, o->dim), best, bd);
    ids[i] = best ? o->id_of.at(best.get()) : -1;
    dists[i] = bd;
  }
}
void ref_nn_near(void* h, const double* q, int nq, const double* radius, int cap, int* ids, double* dists,
                 int* counts)
{
  RefNN* o = (RefNN*)h;
  for (int i = 0; i < nq; i++)
  {
    std::multimap<double, NodePtr> m = o->nn->near(toVec(q + (size_t)i * o->dim, o->dim), radius[i]);
    counts[i] = copyMap(o, m, ids + (size_t)i * cap, dists + (size_t)i * cap, cap);
  }
}
void ref_nn_knn(void* h, const double* q, int nq, unsigned long long k, int cap, int* ids, double* dists,
                int* counts)
{
  RefNN* o = (RefNN*)h;
  for (int i = 0; i < nq; i++)
  {
    std::multimap<double, NodePtr> m = o->nn->kNearestNeighbors(toVec(q + (size_t)i * o->dim, o->dim), (size_t)k);
    counts[i] = copyMap(o, m, ids + (size_t)i * cap, dists + (size_t)i * cap, cap);
  }
}

// ---- collision checkers ---------------------------------------------------------------------------
struct RefChecker
{
  CollisionCheckerPtr c;
  int dim;
};
void* ref_checker_cube(int dim, double threshold, double min_distance)
{
  return new RefChecker{ std::make_shared<Cube3dCollisionChecker>(g_logger, threshold, min_distance), dim };
}
// fn has the signature of the oracle's orc_check_states; ctx is the oracle world handle
void* ref_checker_callback(int dim, void* fn, void* ctx, double min_distance)
{
  return new RefChecker{ std::make_shared<CallbackChecker>((check_states_fn)fn, ctx, min_distance), dim };
}
void ref_checker_free(void* h) { delete (RefChecker*)h; }
unsigned long long ref_checker_states(void* h)
{
  CallbackChecker* c = dynamic_cast<CallbackChecker*>(((RefChecker*)h)->c.get());
  return c ? c->states : 0ull;
}
void ref_check_states(void* h, const double* q, int n, unsigned char* ok)
{
  RefChecker* r = (RefChecker*)h;
  for (int i = 0; i < n; i++)
    ok[i] = r->c->check(toVec(q + (size_t)i * r->dim, r->dim)) ? 1 : 0;
}
// mode 0: checkConnection(c1,c2) (:207-237).  mode 1: checkConnection(c1,c2,conf) (:178-205); last_conf[i] receives
// conf (pre-filled with NaN so "left untouched" is visible).
void ref_check_edges(void* h, const double* c1, const double* c2, int n, int mode, unsigned char* ok, double* last_conf)
{
  RefChecker* r = (RefChecker*)h;
  const int d = r->dim;
  for (int i = 0; i < n; i++)
  {
    Eigen::VectorXd a = toVec(c1 + (size_t)i * d, d), b = toVec(c2 + (size_t)i * d, d);
    if (mode == 1)
    {
      Eigen::VectorXd conf(d);
      conf.setConstant(std::numeric_limits<double>::quiet_NaN());
      ok[i] = r->c->checkConnection(a, b, conf) ? 1 : 0;
      if (last_conf)
        for (int k = 0; k < d; k++)
          last_conf[(size_t)i * d + k] = conf(k);
    }
    else
      ok[i] = r->c->checkConnection(a, b) ? 1 : 0;
  }
}
// checkConnFromConf (:264-313): 1 free, 0 collision, 255 the reference threw std::invalid_argument
void ref_check_from_conf(void* h, const double* parent, const double* child, const double* conf, int n,
                         unsigned char* result)
{
  RefChecker* r = (RefChecker*)h;
  const int d = r->dim;
  for (int i = 0; i < n; i++)
  {
    NodePtr p = std::make_shared<Node>(toVec(parent + (size_t)i * d, d), g_logger);
    NodePtr c = std::make_shared<Node>(toVec(child + (size_t)i * d, d), g_logger);
    ConnectionPtr conn = std::make_shared<Connection>(p, c, g_logger);
    conn->add();
    try
    {
      result[i] = r->c->checkConnFromConf(conn, toVec(conf + (size_t)i * d, d)) ? 1 : 0;
    }
    catch (const std::invalid_argument&)
    {
      result[i] = 255;
    }
    conn->remove();
  }
}

// ---- solvers: kind 0 RRT, 1 RRTStar, 2 BiRRT, 3 AnytimeRRT ------------------------------------------
void* ref_solver_new(int kind, int dim, void* checker, const double* lb, const double* ub, const double* start,
                     const double* goal, double max_distance, double rewire_factor, double utopia_tolerance,
                     int use_kdtree, int extend, int informed_sampler)
{
  RefSolver* s = new RefSolver();
  s->kind = kind;
  s->dim = dim;
  s->checker = ((RefChecker*)checker)->c;
  s->metrics = std::make_shared<EuclideanMetrics>(g_logger);
  Eigen::VectorXd vlb = toVec(lb, dim), vub = toVec(ub, dim), vs = toVec(start, dim), vg = toVec(goal, dim);
  if (informed_sampler)
    s->sampler = std::make_shared<InformedSampler>(vs, vg, vlb, vub, g_logger);
  else
    s->sampler = std::make_shared<UniformSampler>(vlb, vub, g_logger);
  switch (kind)
  {
    case 0:
      s->solver = std::make_shared<RRT>(s->metrics, s->checker, s->sampler, g_logger);
      break;
    case 1:
      s->solver = std::make_shared<RRTStar>(s->metrics, s->checker, s->sampler, g_logger);
      break;
    case 2:
      s->solver = std::make_shared<BiRRT>(s->metrics, s->checker, s->sampler, g_logger);
      break;
    default:
      s->solver = std::make_shared<AnytimeRRT>(s->metrics, s->checker, s->sampler, g_logger);
      break;
  }
  const std::string ns = "/ref_driver";
  cnr::param::shim_set(ns + "/max_distance", max_distance);
  cnr::param::shim_set(ns + "/use_kdtree", use_kdtree ? 1.0 : 0.0);
  cnr::param::shim_set(ns + "/extend", extend ? 1.0 : 0.0);
  cnr::param::shim_set(ns + "/utopia_tolerance", utopia_tolerance);
  cnr::param::shim_set(ns + "/rewire_factor", rewire_factor);
  s->solver->config(ns);
  s->start = std::make_shared<Node>(vs, g_logger);
  s->goal = std::make_shared<Node>(vg, g_logger);
  s->solver->addStart(s->start);
  s->solver->addGoal(s->goal);
  return s;
}
void ref_solver_free(void* h) { delete (RefSolver*)h; }
// feeds n injected samples through update(configuration, solution); same stopping rule as orc_solver_run
int ref_solver_run(void* h, const double* samples, int n)
{
  RefSolver* s = (RefSolver*)h;
  int i = 0;
  for (; i < n; i++)
  {
    bool r = s->solver->update(toVec(samples + (size_t)i * s->dim, s->dim), s->solution);
    if (r && !s->solver->canImprove())
    {
      i++;
      break;
    }
  }
  return i;
}
int ref_solver_update(void* h, const double* q)
{
  RefSolver* s = (RefSolver*)h;
  return s->solver->update(toVec(q, s->dim), s->solution) ? 1 : 0;
}
// AnytimeRRT / generic: the solver's own solve loop with its own (std::rand) sampler
int ref_solver_solve(void* h, unsigned int max_iter, double max_time)
{
  RefSolver* s = (RefSolver*)h;
  return s->solver->solve(s->solution, max_iter, max_time) ? 1 : 0;
}
int ref_solver_solved(void* h) { return ((RefSolver*)h)->solver->solved() ? 1 : 0; }
int ref_solver_can_improve(void* h) { return ((RefSolver*)h)->solver->canImprove() ? 1 : 0; }
double ref_solver_cost(void* h) { return ((RefSolver*)h)->solver->cost(); }
double ref_solver_specific_volume(void* h) { return ((RefSolver*)h)->sampler->getSpecificVolume(); }
int ref_solver_tree_size(void* h)
{
  TreePtr t = ((RefSolver*)h)->solver->getStartTree();
  return t ? (int)t->getNumberOfNodes() : 0;
}
// every node of the start tree, in NearestNeighbors::getNodes() order: conf[n][dim], parent_conf[n][dim]
// (NaN for the root), cost of the parent connection (0 for the root).  Returns n; fills at most cap rows.
int ref_solver_dump(void* h, int cap, double* conf, double* parent_conf, double* pcost)
{
  RefSolver* s = (RefSolver*)h;
  TreePtr t = s->solver->getStartTree();
  if (!t)
    return 0;
  std::vector<NodePtr> nodes = t->getNodes();
  const int d = s->dim;
  for (int i = 0; i < (int)nodes.size() && i < cap; i++)
  {
    for (int k = 0; k < d; k++)
      conf[(size_t)i * d + k] = nodes[i]->getConfiguration()(k);
    if (nodes[i]->getParentConnectionsSize() == 1)
    {
      ConnectionPtr c = nodes[i]->parentConnection(0);
      for (int k = 0; k < d; k++)
        parent_conf[(size_t)i * d + k] = c->getParent()->getConfiguration()(k);
      pcost[i] = c->getCost();
    }
    else
    {
      for (int k = 0; k < d; k++)
        parent_conf[(size_t)i * d + k] = std::numeric_limits<double>::quiet_NaN();
      pcost[i] = 0.0;
    }
  }
  return (int)nodes.size();
}
// waypoints of the current solution, start first; returns their number
int ref_solver_solution(void* h, int cap, double* conf)
{
  RefSolver* s = (RefSolver*)h;
  PathPtr p = s->solver->getSolution();
  if (!p || !s->solver->solved())
    return 0;
  std::vector<Eigen::VectorXd> wp = p->getWaypoints();
  for (int i = 0; i < (int)wp.size() && i < cap; i++)
    for (int k = 0; k < s->dim; k++)
      conf[(size_t)i * s->dim + k] = wp[i](k);
  return (int)wp.size();
}
double ref_solver_path_cost(void* h)
{
  PathPtr p = ((RefSolver*)h)->solver->getSolution();
  return p ? p->cost() : std::numeric_limits<double>::infinity();
}

// ---- AnytimeRRT step by step (anytime_rrt.cpp:269-427) with injected samples ---------------------------------------
// improve(start, goal, solution, max_iter = 0) runs the reference's own round set-up (utopia tests, cost2beat, bias,
// new tree, :269-315) without drawing a sample; improveUpdate(point, solution) is then one iteration of its loop.
// Returns 1 when a round tree was set up, 0 when improve() returned before (nothing changed but can_improve_).
int ref_anytime_improve_begin(void* h)
{
  RefSolver* s = (RefSolver*)h;
  AnytimeRRTPtr a = std::dynamic_pointer_cast<AnytimeRRT>(s->solver);
  if (!a)
    return -1;
  TreePtr before = a->getNewTree();
  NodePtr tmp_start = std::make_shared<Node>(s->start->getConfiguration(), g_logger);
  NodePtr tmp_goal = std::make_shared<Node>(s->goal->getConfiguration(), g_logger);
  a->improve(tmp_start, tmp_goal, s->solution, 0u, std::numeric_limits<double>::infinity());
  return (a->getNewTree() != before) ? 1 : 0;
}
int ref_anytime_improve_update(void* h, const double* q)
{
  RefSolver* s = (RefSolver*)h;
  AnytimeRRTPtr a = std::dynamic_pointer_cast<AnytimeRRT>(s->solver);
  if (!a)
    return -1;
  return a->improveUpdate(toVec(q, s->dim), s->solution) ? 1 : 0;
}
// out: bias, cost2beat, path_cost, cost, can_improve, solved, nodes in the round tree (0: none)
void ref_anytime_state(void* h, double* out)
{
  RefSolver* s = (RefSolver*)h;
  AnytimeRRTPtr a = std::dynamic_pointer_cast<AnytimeRRT>(s->solver);
  out[0] = a->getBias();
  out[1] = a->getCost2Beat();
  out[2] = a->getPathCost();
  out[3] = a->cost();
  out[4] = a->canImprove() ? 1.0 : 0.0;
  out[5] = a->solved() ? 1.0 : 0.0;
  out[6] = a->getNewTree() ? (double)a->getNewTree()->getNumberOfNodes() : 0.0;
}
// AnytimeRRT::solve (anytime_rrt.cpp:73-236) without the wall clock, samples injected: sample k of this problem comes
// from `sampler` (the oracle's orc_sample_informed) with Philox counter k * P + p, foci start/goal, cost = cost0 during
// the RRT phase and cost2beat during an improve round.  Every planner decision is the reference's own update() /
// improve(max_iter = 0) / improveUpdate(); the loop around them restates solve()'s (see ref_py.anytime_solve).
// out: rounds, updates, improved rounds.  Returns solved.
typedef void (*informed_sampler_fn)(int dim, int n, const double* lb, const double* ub, const double* f1, const double* f2,
                                    const double* cost, unsigned long long seed, unsigned long long first, double* out,
                                    unsigned int* attempts);
int ref_anytime_solve_injected(void* h, unsigned int max_iter, int improve_in_solve, informed_sampler_fn sampler,
                               const double* lb, const double* ub, double cost0, unsigned long long seed,
                               unsigned long long p, unsigned long long P, double utopia_tolerance, double* out)
{
  RefSolver* s = (RefSolver*)h;
  AnytimeRRTPtr a = std::dynamic_pointer_cast<AnytimeRRT>(s->solver);
  if (!a)
    return -1;
  const int d = s->dim;
  std::vector<double> f1(d), f2(d), q(d);
  for (int i = 0; i < d; i++)
  {
    f1[i] = s->start->getConfiguration()(i);
    f2[i] = s->goal->getConfiguration()(i);
  }
  const double best_utopia = (s->goal->getConfiguration() - s->start->getConfiguration()).norm();
  const double tol = (1.0 + (utopia_tolerance > 0.0 ? utopia_tolerance : 0.0)) * best_utopia;
  unsigned long long k = 0, rounds = 0, updates = 0, improved_rounds = 0;
  auto draw = [&](double cost) {
    sampler(d, 1, lb, ub, f1.data(), f2.data(), &cost, seed, k * P + p, q.data(), nullptr);
    k++;
    updates++;
    return toVec(q.data(), d);
  };
  auto finish = [&]() {
    if (out)
    {
      out[0] = (double)rounds;
      out[1] = (double)updates;
      out[2] = (double)improved_rounds;
    }
    return a->solved() ? 1 : 0;
  };
  if (a->solved() && a->cost() <= tol)
    return finish();
  int n_failed = 0;
  while (!a->solved() && n_failed < FAILED_ITER)
  {
    bool ok = false;
    rounds++;
    for (unsigned int it = 0; it < max_iter; it++)
      if (s->solver->update(draw(cost0), s->solution))  // RRT::update (AnytimeRRT hides the overload by name)
      {
        ok = true;
        break;
      }
    if (!ok)
      n_failed++;
  }
  if (!a->solved() || a->cost() <= tol)
    return finish();
  n_failed = 0;
  while ((improve_in_solve || a->canImprove()) && n_failed < FAILED_ITER)
  {
    bool improved = false;
    rounds++;
    if (ref_anytime_improve_begin(h) == 1)
    {
      const double c2b = a->getCost2Beat();
      for (unsigned int it = 0; it < max_iter; it++)
        if (a->improveUpdate(draw(c2b), s->solution))
        {
          improved = true;
          break;
        }
    }
    if (improved)
    {
      n_failed = 0;
      improved_rounds++;
    }
    else
      n_failed++;
    if (a->cost() <= tol)
      break;
  }
  return finish();
}
void ref_anytime_set_parameters(void* h, double delta, double cost_impr)
{
  AnytimeRRTPtr a = std::dynamic_pointer_cast<AnytimeRRT>(((RefSolver*)h)->solver);
  if (a)
    a->setParameters(delta, cost_impr);
}

// ---- Path validity and local optimisation with the SOLVER'S OWN checker ------------
// Path::isValid (path.cpp:1230-1260 -> isValidFromConn): every connection of the solution is re-checked, invalid
// ones get cost +inf.  costs[i] receives the cost of connection i afterwards; returns 1 valid, 0 invalid, -1 no path.
// other_checker (a RefChecker*, may be NULL) replaces the path's checker, e.g. the same scene with new obstacles.
int ref_solution_is_valid(void* h, void* other_checker, int cap, double* costs)
{
  RefSolver* s = (RefSolver*)h;
  PathPtr p = s->solver->getSolution();
  if (!p || !s->solver->solved())
    return -1;
  CollisionCheckerPtr chk = other_checker ? ((RefChecker*)other_checker)->c : CollisionCheckerPtr();
  const bool valid = p->isValid(chk);
  const std::vector<ConnectionPtr>& conns = p->getConnectionsConst();
  for (int i = 0; i < (int)conns.size() && i < cap; i++)
    costs[i] = conns[i]->getCost();
  return valid ? 1 : 0;
}
// Path::isValidFromConf (path.cpp:1262-1407) from the point at fraction t of connection conn_idx
int ref_solution_is_valid_from(void* h, void* other_checker, int conn_idx, double t)
{
  RefSolver* s = (RefSolver*)h;
  PathPtr p = s->solver->getSolution();
  if (!p || !s->solver->solved())
    return -1;
  CollisionCheckerPtr chk = other_checker ? ((RefChecker*)other_checker)->c : CollisionCheckerPtr();
  const std::vector<ConnectionPtr>& conns = p->getConnectionsConst();
  if (conn_idx < 0 || conn_idx >= (int)conns.size())
    return -1;
  const Eigen::VectorXd a = conns[conn_idx]->getParent()->getConfiguration();
  const Eigen::VectorXd b = conns[conn_idx]->getChild()->getConfiguration();
  const Eigen::VectorXd conf = a + (b - a) * t;
  return p->isValidFromConf(conf, (size_t)conn_idx, chk) ? 1 : 0;
}
// PathLocalOptimizer (path_local_optimizer.cpp:59-204) on a CLONE of the solution: mode 0 warp, 1 simplify, 2 solve();
// the optimised waypoints go to conf, returns their number (0 = no path)
int ref_solution_optimize(void* h, int mode, int max_iter, int cap, double* conf, double* cost_out)
{
  RefSolver* s = (RefSolver*)h;
  PathPtr p = s->solver->getSolution();
  if (!p || !s->solver->solved())
    return 0;
  PathPtr q = p->clone();
  PathLocalOptimizer opt(s->checker, s->metrics, g_logger);
  opt.config("/ref_driver/path_optimizer");
  opt.setPath(q);
  if (mode == 0)
    for (int i = 0; i < max_iter; i++)
      opt.warp();
  else if (mode == 1)
    for (int i = 0; i < max_iter; i++)
      opt.simplify();
  else
    opt.solve((unsigned int)max_iter, 1.0e9);
  std::vector<Eigen::VectorXd> wp = q->getWaypoints();
  for (int i = 0; i < (int)wp.size() && i < cap; i++)
    for (int k = 0; k < s->dim; k++)
      conf[(size_t)i * s->dim + k] = wp[i](k);
  if (cost_out)
    *cost_out = q->cost();
  return (int)wp.size();
}
// Tree::recheckCollision (tree.cpp:1405-1429): re-validates the whole start tree with the solver's checker
int ref_tree_recheck(void* h)
{
  TreePtr t = ((RefSolver*)h)->solver->getStartTree();
  return (t && t->recheckCollision()) ? 1 : 0;
}

// Tree::checkPathToNode (tree.cpp:335-373) towards the tree node closest to `conf` (pass a node's own configuration),
// with the tree's checker or -- for the duration of the call -- other_checker.  costs[] receives the cost of every
// connection root -> node afterwards, *n_path their number, *n_checked how many the call actually checked (the
// "recently checked" ones are skipped, :345-352).  keep_flags == 0 clears the recently-checked marks again.
int ref_check_path_to_node(void* h, void* other_checker, const double* conf, int keep_flags, int cap, double* costs,
                           int* n_checked, int* n_path)
{
  RefSolver* s = (RefSolver*)h;
  TreePtr t = s->solver->getStartTree();
  if (!t)
    return -1;
  NodePtr node = t->findClosestNode(toVec(conf, s->dim));
  CollisionCheckerPtr old = t->getChecker();
  if (other_checker)
    t->setChecker(((RefChecker*)other_checker)->c);
  std::vector<ConnectionPtr> checked, path;
  const bool ok = t->checkPathToNode(node, checked, path);
  t->setChecker(old);
  for (int i = 0; i < (int)path.size() && i < cap; i++)
    costs[i] = path[i]->getCost();
  *n_checked = (int)checked.size();
  *n_path = (int)path.size();
  if (!keep_flags)
    for (const ConnectionPtr& c : checked)
      c->setRecentlyChecked(false);
  return ok ? 1 : 0;
}

// Tree::rewireOnlyWithPathCheck (tree.cpp:522-679) around the tree node at `conf`, with `other_checker` (may be NULL) as
// the tree's checker for the call.  Returns the reference's boolean; n_checked = connections it flagged.
int ref_rewire_only_with_path_check(void* h, void* other_checker, const double* conf, double r_rewire, int what_rewire,
                                    int keep_flags, int* n_checked)
{
  RefSolver* s = (RefSolver*)h;
  TreePtr t = s->solver->getStartTree();
  if (!t)
    return -1;
  NodePtr node = t->findClosestNode(toVec(conf, s->dim));
  CollisionCheckerPtr old = t->getChecker();
  if (other_checker)
    t->setChecker(((RefChecker*)other_checker)->c);
  std::vector<ConnectionPtr> checked;
  const bool ok = t->rewireOnlyWithPathCheck(node, checked, r_rewire, what_rewire);
  t->setChecker(old);
  *n_checked = (int)checked.size();
  if (!keep_flags)
    for (const ConnectionPtr& c : checked)
      c->setRecentlyChecked(false);
  return ok ? 1 : 0;
}

// Tree::informedExtend (tree.cpp:235-302) on a tree given as arrays: conf[n][dim], parent[n] (index, -1 for the root =
// node 0) and the connection costs pcost[n].  Nodes enter the NearestNeighbors structure in index order.  Returns the
// reference's boolean; new_conf / parent_index describe the node it appended.
int ref_informed_extend(int dim, void* checker, int n, const double* conf, const int* parent, const double* pcost,
                        double max_distance, int use_kdtree, const double* q, const double* goal, double cost2beat,
                        double bias, double* new_conf, int* parent_index)
{
  RefChecker* rc = (RefChecker*)checker;
  MetricsPtr metrics = std::make_shared<EuclideanMetrics>(g_logger);
  std::vector<NodePtr> nodes;
  for (int i = 0; i < n; i++)
    nodes.push_back(std::make_shared<Node>(toVec(conf + (size_t)i * dim, dim), g_logger));
  TreePtr tree = std::make_shared<Tree>(nodes[0], max_distance, rc->c, metrics, g_logger, use_kdtree != 0);
  for (int i = 1; i < n; i++)
  {
    ConnectionPtr c = std::make_shared<Connection>(nodes[parent[i]], nodes[i], g_logger);
    c->setCost(pcost[i]);
    c->add();
  }
  for (int i = 1; i < n; i++)
    tree->addNode(nodes[i], false);
  NodePtr new_node;
  const bool ok = tree->informedExtend(toVec(q, dim), new_node, toVec(goal, dim), cost2beat, bias);
  *parent_index = -1;
  if (ok && new_node)
  {
    for (int k = 0; k < dim; k++)
      new_conf[k] = new_node->getConfiguration()(k);
    NodePtr par = new_node->getParents().at(0);
    for (int i = 0; i < n; i++)
      if (nodes[i] == par)
        *parent_index = i;
  }
  // break the shared_ptr cycles of the reference's graph objects
  for (NodePtr& nd : nodes)
    nd->disconnect();
  if (new_node)
    new_node->disconnect();
  return ok ? 1 : 0;
}

// ---- Tree YAML (tree.cpp:1188-1227 toYAML, :1272-1389 fromYAML): the document as arrays ---------------------
// Tree::toYAML() of the start tree: nodes[n][dim] in document order, connections[m][2] = (parent index, child
// index) in document order.  Returns n (or -1); *nconn = m.
int ref_tree_to_yaml(void* h, int cap_nodes, double* conf, int cap_conn, int* conn, int* nconn, double* max_distance,
                     int* use_kdtree)
{
  RefSolver* s = (RefSolver*)h;
  TreePtr t = s->solver->getStartTree();
  if (!t)
    return -1;
  YAML::Node y = t->toYAML();
  YAML::Node nodes = y["nodes"], cs = y["connections"];
  *max_distance = y["max_distance"].as<double>();
  *use_kdtree = y["use_kdtree"].as<bool>() ? 1 : 0;
  for (std::size_t i = 0; i < nodes.size() && (int)i < cap_nodes; i++)
    for (int k = 0; k < s->dim; k++)
      conf[i * s->dim + k] = nodes[i][k].as<double>();
  const std::size_t m = cs.IsSequence() ? cs.size() : 0;
  for (std::size_t i = 0; i < m && (int)i < cap_conn; i++)
  {
    conn[2 * i] = cs[i][0].as<int>();
    conn[2 * i + 1] = cs[i][1].as<int>();
  }
  *nconn = (int)m;
  return (int)nodes.size();
}
// Tree::fromYAML on a document given as arrays (max_distance NaN = field absent).  Dumps the loaded tree in
// getNodes() order like ref_solver_dump; *max_distance_out = Tree::getMaximumDistance().  Returns the node count, -1
// when the reference refuses the document.
int ref_tree_from_yaml(int dim, int n, const double* conf, int m, const int* conn, double max_distance, int use_kdtree,
                       int cap, double* out_conf, double* out_parent_conf, double* out_pcost, double* max_distance_out)
{
  YAML::Node y, nodes, cs;
  for (int i = 0; i < n; i++)
  {
    YAML::Node q;
    q.SetStyle(YAML::EmitterStyle::Flow);
    for (int k = 0; k < dim; k++)
      q.push_back(conf[(size_t)i * dim + k]);
    nodes.push_back(q);
  }
  for (int i = 0; i < m; i++)
  {
    YAML::Node c;
    c.SetStyle(YAML::EmitterStyle::Flow);
    c.push_back(conn[2 * i]);
    c.push_back(conn[2 * i + 1]);
    cs.push_back(c);
  }
  if (max_distance == max_distance)
    y["max_distance"] = max_distance;
  y["use_kdtree"] = (use_kdtree != 0);
  y["nodes"] = nodes;
  y["connections"] = cs;
  MetricsPtr metrics = std::make_shared<EuclideanMetrics>(g_logger);
  CollisionCheckerPtr checker = std::make_shared<Cube3dCollisionChecker>(g_logger, 0.0, 0.01);
  TreePtr t = Tree::fromYAML(y, metrics, checker, g_logger);
  if (!t)
    return -1;
  *max_distance_out = t->getMaximumDistance();
  std::vector<NodePtr> nv = t->getNodes();
  for (int i = 0; i < (int)nv.size() && i < cap; i++)
  {
    for (int k = 0; k < dim; k++)
      out_conf[(size_t)i * dim + k] = nv[i]->getConfiguration()(k);
    if (nv[i]->getParentConnectionsSize() == 1)
    {
      ConnectionPtr c = nv[i]->parentConnection(0);
      for (int k = 0; k < dim; k++)
        out_parent_conf[(size_t)i * dim + k] = c->getParent()->getConfiguration()(k);
      out_pcost[i] = c->getCost();
    }
    else
    {
      for (int k = 0; k < dim; k++)
        out_parent_conf[(size_t)i * dim + k] = std::numeric_limits<double>::quiet_NaN();
      out_pcost[i] = 0.0;
    }
  }
  return (int)nv.size();
}

// ---- samplers: the reference's own sampling code on the shim's seedable stream ------------------------
// seeds the one random stream the reference draws from in this library (oracle/ref_shim/shim_random.hpp)
void ref_srand(unsigned int seed)
{
  std::srand(seed);
  gc_shim_rng::seed(seed);
}
void ref_sample_informed(int dim, const double* lb, const double* ub, const double* f1, const double* f2, double cost,
                         int n, double* out)
{
  InformedSampler smp(toVec(f1, dim), toVec(f2, dim), toVec(lb, dim), toVec(ub, dim), g_logger, cost);
  for (int i = 0; i < n; i++)
  {
    Eigen::VectorXd q = smp.sample();
    for (int k = 0; k < dim; k++)
      out[(size_t)i * dim + k] = q(k);
  }
}
// ellipse parameters after setCost: out = [specific_volume, focii_distance, min_radius, max_radius]... only what the
// public API exposes: specific volume and in-bounds tests
double ref_informed_specific_volume(int dim, const double* lb, const double* ub, const double* f1, const double* f2,
                                    double cost)
{
  InformedSampler smp(toVec(f1, dim), toVec(f2, dim), toVec(lb, dim), toVec(ub, dim), g_logger, cost);
  return smp.getSpecificVolume();
}
// RRTStar::updateRewireRadius is protected; the k of Tree::nearK and Metrics::cost are reachable
double ref_metrics_cost(int dim, const double* a, const double* b)
{
  EuclideanMetrics m(g_logger);
  return m.cost(toVec(a, dim), toVec(b, dim));
}

}  // extern "C"
