/*
 * ref_dropin.cpp -- the UNMODIFIED graph_core solvers (RRT, RRTStar, BiRRT, AnytimeRRT) running with this
 * repository's drop-in plugin classes in place of their CPU hot paths:
 *   graph::core::CudaCollisionChecker behind CollisionCheckerBase   (every solver)
 *   graph::core::CudaNearestNeighbors behind NearestNeighbors       (RRT / RRTStar, through addStartTree)
 *
 * TEST INFRASTRUCTURE (oracle/): graph_core_b200/host/cuda_plugins.cpp is compiled here with
 * -DGCB200_WITH_GRAPH_CORE against graph_core's GENUINE headers (plus the oracle/ref_shim stand-ins for Eigen,
 * cnr_logger, cnr_param, yaml-cpp, cnr_class_loader) and linked with the reference's own objects and libgcb200.so
 * into oracle/_ref/libgc_ref_dropin.so.  tests/test_ref_dropin_gpu.py then checks that the reference's planners
 * build the SAME trees and paths with the CUDA plugins as with their CPU KdTree + per-state checker.
 * Everything of ref_driver.cpp (ref_* entry points) is available in this library too.
 */
#include "ref_driver.cpp"

#include "cuda_batched_helpers.h"
#include "cuda_plugins.h"

#include <atomic>

namespace
{
// The reference creates its NearestNeighbors inside Tree's constructor (tree.cpp:42-50) and offers no factory
// hook; a Tree subclass can swap the protected nodes_ member right after construction.
class CudaTree : public Tree
{
public:
  CudaTree(const NodePtr& root, const double& max_distance, const CollisionCheckerPtr& checker, const MetricsPtr& metrics,
           const cnr_logger::TraceLoggerPtr& logger, int device)
    : Tree(root, max_distance, checker, metrics, logger, true)
  {
    nodes_ = std::make_shared<CudaNearestNeighbors>(logger, device);
    nodes_->insert(root);
  }
};
}  // namespace

extern "C" {

// same arguments as ref_solver_new, but `world` is a gc_world_t* of libgcb200 and the NearestNeighbors structure of
// the start tree is the CUDA one when use_cuda_nn != 0
void* dropin_solver_new(int kind, int dim, void* world, const double* lb, const double* ub, const double* start,
                        const double* goal, double max_distance, double min_distance, double rewire_factor,
                        double utopia_tolerance, int extend, int informed_sampler, int use_cuda_nn, int device)
{
  RefSolver* s = new RefSolver();
  s->kind = kind;
  s->dim = dim;
  try
  {
    s->checker = std::make_shared<CudaCollisionChecker>(g_logger, (gc_world_t*)world, min_distance);
  }
  catch (const std::exception&)
  {
    delete s;
    return nullptr;
  }
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
  const std::string ns = "/ref_dropin";
  cnr::param::shim_set(ns + "/max_distance", max_distance);
  cnr::param::shim_set(ns + "/use_kdtree", 1.0);
  cnr::param::shim_set(ns + "/extend", extend ? 1.0 : 0.0);
  cnr::param::shim_set(ns + "/utopia_tolerance", utopia_tolerance);
  cnr::param::shim_set(ns + "/rewire_factor", rewire_factor);
  s->solver->config(ns);
  s->start = std::make_shared<Node>(vs, g_logger);
  s->goal = std::make_shared<Node>(vg, g_logger);
  if (use_cuda_nn)
    s->solver->addStartTree(std::make_shared<CudaTree>(s->start, max_distance, s->checker, s->metrics, g_logger, device));
  else
    s->solver->addStart(s->start);
  s->solver->addGoal(s->goal);
  return s;
}

// a RefChecker (as ref_checker_* make them) backed by a CUDA world: usable wherever the ref_* entry points take one
void* dropin_checker_new(int dim, void* world, double min_distance)
{
  try
  {
    return new RefChecker{ std::make_shared<CudaCollisionChecker>(g_logger, (gc_world_t*)world, min_distance), dim };
  }
  catch (const std::exception&)
  {
    return nullptr;
  }
}

// device calls made by the checker of a drop-in solver (proof that the CUDA path answered)
unsigned long long dropin_checker_device_calls(void* h)
{
  RefSolver* s = (RefSolver*)h;
  CudaCollisionChecker* c = dynamic_cast<CudaCollisionChecker*>(s->checker.get());
  return c ? c->deviceCalls() : 0ull;
}
#ifdef GCB200_TREE_FACTORY_HOOK
// with the factory hook (applied to a build-time copy of tree.h / tree.cpp, see
// oracle/ref_build.py: patched_tree_sources) EVERY Tree the solvers create -- BiRRT's goal tree (birrt.cpp:67),
// AnytimeRRT's per-round trees (anytime_rrt.cpp:298), RRT's start tree (rrt.cpp:73) -- owns a CudaNearestNeighbors.
// device < 0 removes the factory (stock KdTree / Vector again).
static std::atomic<unsigned long long> g_factory_calls{ 0 };
void dropin_set_nn_factory(int device)
{
  if (device < 0)
  {
    Tree::setNearestNeighborsFactory(nullptr);
    return;
  }
  Tree::setNearestNeighborsFactory([device](const cnr_logger::TraceLoggerPtr& lg) -> NearestNeighborsPtr {
    g_factory_calls++;
    return std::make_shared<CudaNearestNeighbors>(lg, device);
  });
}
unsigned long long dropin_nn_factory_calls() { return g_factory_calls.load(); }
#endif

// ---- SURVEY 8(f) N1-N3 with the edge checks batched (graph_core_b200/host/cuda_batched_helpers.h) -------------
static std::shared_ptr<CudaCollisionChecker> cudaCheckerOf(RefSolver* s, void* other_checker)
{
  CollisionCheckerPtr c = other_checker ? ((RefChecker*)other_checker)->c : s->checker;
  return std::dynamic_pointer_cast<CudaCollisionChecker>(c);
}
unsigned long long dropin_device_calls_of(void* h, void* other_checker)
{
  auto c = cudaCheckerOf((RefSolver*)h, other_checker);
  return c ? c->deviceCalls() : 0ull;
}
// ref_solution_is_valid through cuda_batched::isValid: one device call for the whole path
int dropin_solution_is_valid_batched(void* h, void* other_checker, int cap, double* costs)
{
  RefSolver* s = (RefSolver*)h;
  PathPtr p = s->solver->getSolution();
  auto chk = cudaCheckerOf(s, other_checker);
  if (!p || !s->solver->solved() || !chk)
    return -1;
  const bool valid = cuda_batched::isValid(p, chk);
  const std::vector<ConnectionPtr>& conns = p->getConnectionsConst();
  for (int i = 0; i < (int)conns.size() && i < cap; i++)
    costs[i] = conns[i]->getCost();
  return valid ? 1 : 0;
}
int dropin_solution_is_valid_from_batched(void* h, void* other_checker, int conn_idx, double t)
{
  RefSolver* s = (RefSolver*)h;
  PathPtr p = s->solver->getSolution();
  auto chk = cudaCheckerOf(s, other_checker);
  if (!p || !s->solver->solved() || !chk)
    return -1;
  const std::vector<ConnectionPtr>& conns = p->getConnectionsConst();
  if (conn_idx < 0 || conn_idx >= (int)conns.size())
    return -1;
  const Eigen::VectorXd a = conns[conn_idx]->getParent()->getConfiguration();
  const Eigen::VectorXd b = conns[conn_idx]->getChild()->getConfiguration();
  const Eigen::VectorXd conf = a + (b - a) * t;
  int pos = -1;
  return cuda_batched::isValidFromConf(p, conf, (size_t)conn_idx, pos, chk) ? 1 : 0;
}
int dropin_tree_recheck_batched(void* h)
{
  RefSolver* s = (RefSolver*)h;
  TreePtr t = s->solver->getStartTree();
  auto chk = cudaCheckerOf(s, nullptr);
  return (t && chk && cuda_batched::recheckCollision(t, chk)) ? 1 : 0;
}
int dropin_check_path_to_node_batched(void* h, void* other_checker, const double* conf, int keep_flags, int cap,
                                      double* costs, int* n_checked, int* n_path)
{
  RefSolver* s = (RefSolver*)h;
  TreePtr t = s->solver->getStartTree();
  auto chk = cudaCheckerOf(s, other_checker);
  if (!t || !chk)
    return -1;
  NodePtr node = t->findClosestNode(toVec(conf, s->dim));
  CollisionCheckerPtr old = t->getChecker();
  t->setChecker(chk);
  std::vector<ConnectionPtr> checked, path;
  const bool ok = cuda_batched::checkPathToNode(t, node, checked, path, chk);
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
// ref_rewire_only_with_path_check through cuda_batched::rewireOnlyWithPathCheck: one device call for the near set's
// edges and every connection on the way to those nodes
int dropin_rewire_only_with_path_check_batched(void* h, void* other_checker, const double* conf, double r_rewire,
                                               int what_rewire, int keep_flags, int* n_checked)
{
  RefSolver* s = (RefSolver*)h;
  TreePtr t = s->solver->getStartTree();
  auto chk = cudaCheckerOf(s, other_checker);
  if (!t || !chk)
    return -1;
  NodePtr node = t->findClosestNode(toVec(conf, s->dim));
  CollisionCheckerPtr old = t->getChecker();
  t->setChecker(chk);
  std::vector<ConnectionPtr> checked;
  std::vector<NodePtr> white_list;
  const bool ok = cuda_batched::rewireOnlyWithPathCheck(t, node, checked, r_rewire, white_list, what_rewire, chk);
  t->setChecker(old);
  *n_checked = (int)checked.size();
  if (!keep_flags)
    for (const ConnectionPtr& c : checked)
      c->setRecentlyChecked(false);
  return ok ? 1 : 0;
}
// ref_solution_optimize with cuda_batched::SpeculativePathLocalOptimizer
int dropin_solution_optimize_speculative(void* h, int mode, int max_iter, int cap, double* conf, double* cost_out)
{
  RefSolver* s = (RefSolver*)h;
  PathPtr p = s->solver->getSolution();
  auto chk = cudaCheckerOf(s, nullptr);
  if (!p || !s->solver->solved() || !chk)
    return 0;
  PathPtr q = p->clone();
  cuda_batched::SpeculativePathLocalOptimizer opt(chk, s->metrics, g_logger);
  opt.config("/ref_driver/path_optimizer");
  opt.setPath(q);
  if (mode == 0)
    for (int i = 0; i < max_iter; i++)
      opt.warpSpeculative();
  else if (mode == 1)
    for (int i = 0; i < max_iter; i++)
      opt.simplifySpeculative();
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

int dropin_uses_cuda_nn(void* h)
{
  RefSolver* s = (RefSolver*)h;
  TreePtr t = s->solver->getStartTree();
  return (t && dynamic_cast<CudaTree*>(t.get())) ? 1 : 0;
}

}  // extern "C"
