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

#include "cuda_plugins.h"

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
int dropin_uses_cuda_nn(void* h)
{
  RefSolver* s = (RefSolver*)h;
  TreePtr t = s->solver->getStartTree();
  return (t && dynamic_cast<CudaTree*>(t.get())) ? 1 : 0;
}

}  // extern "C"
