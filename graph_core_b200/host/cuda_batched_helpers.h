// cuda_batched_helpers.h -- SURVEY.md 8(f) rows N1-N3: the reference's path / tree re-validation and local path
// optimisation with their edge checks gathered into device batches.
//
// Header only and compiled ONLY against the genuine graph_core headers (-DGCB200_WITH_GRAPH_CORE): it uses Tree, Path
// and PathLocalOptimizer, which host/compat/graph_core_min.h does not re-declare.  Every helper keeps the
// reference's own function as the decision maker: it first hands the edges that function is going to ask for to
// CudaCollisionChecker::prefetchConnections (ONE gc_check_edges call), then calls the unmodified reference function,
// whose checkConnection(conn) / checkConnection(c1, c2) calls are answered from the checker's cache (keyed on the exact
// bit patterns of the ordered end points).  An edge the speculation did not foresee simply misses the cache and costs
// one device call, so verdicts, connection costs, flags and early returns are the reference's by construction.
//
//   N1  Tree::checkPathToNode          src/graph_core/graph/tree.cpp:335-373
//       Tree::recheckCollision         src/graph_core/graph/tree.cpp:1405-1429
//       Tree::rewireOnlyWithPathCheck  src/graph_core/graph/tree.cpp:522-679
//   N2  Path::isValid / isValidFromConn / isValidFromConf   src/graph_core/graph/path.cpp:1230-1407
//   N3  PathLocalOptimizer::bisection / warp / simplify / step
//       src/graph_core/solvers/path_optimizers/path_local_optimizer.cpp:59-246
#pragma once
#ifdef GCB200_WITH_GRAPH_CORE

#include <algorithm>
#include <map>

#include <graph_core/graph/path.h>
#include <graph_core/graph/tree.h>
#include <graph_core/solvers/path_optimizers/path_local_optimizer.h>

#include "cuda_plugins.h"

namespace graph
{
namespace core
{
namespace cuda_batched
{
using CudaCheckerPtr = std::shared_ptr<CudaCollisionChecker>;

inline void prefetch(const CudaCheckerPtr& chk, const std::vector<ConnectionPtr>& conns)
{
  std::vector<const double*> from, to;
  from.reserve(conns.size());
  to.reserve(conns.size());
  for (const ConnectionPtr& c : conns)
  {
    from.push_back(c->getParent()->getConfiguration().data());
    to.push_back(c->getChild()->getConfiguration().data());
  }
  chk->prefetchConnections(from, to);
}

// N2: Path::isValid (path.cpp:1230-1253) -- every connection is checked whatever the others answer (no early exit
// across connections), i.e. one batch
inline bool isValid(const PathPtr& path, const CudaCheckerPtr& chk)
{
  prefetch(chk, path->getConnectionsConst());
  return path->isValid(chk);
}

// N2: Path::isValidFromConf (path.cpp:1290-1398): the connections after conn_idx go as one batch, the partial
// connection conf -> child as the single checkConnFromConf call the reference makes
inline bool isValidFromConf(const PathPtr& path, const Eigen::VectorXd& conf, const size_t& conn_idx,
                            int& pos_closest_obs_from_goal, const CudaCheckerPtr& chk)
{
  const std::vector<ConnectionPtr>& conns = path->getConnectionsConst();
  std::vector<ConnectionPtr> rest;
  const bool from_parent = conn_idx < conns.size() && conf == conns[conn_idx]->getParent()->getConfiguration();
  for (size_t i = from_parent ? conn_idx : conn_idx + 1; i < conns.size(); i++)
    rest.push_back(conns[i]);
  prefetch(chk, rest);
  return path->isValidFromConf(conf, conn_idx, pos_closest_obs_from_goal, chk);
}

// N1: Tree::recheckCollision (tree.cpp:1405-1429) walks the tree depth first and stops at the first colliding
// connection; the batch holds every connection of the tree (a superset of what the walk asks for)
inline bool recheckCollision(const TreePtr& tree, const CudaCheckerPtr& chk)
{
  std::vector<ConnectionPtr> all;
  std::vector<NodePtr> stack{ tree->getRoot() };
  while (!stack.empty())
  {
    NodePtr n = stack.back();
    stack.pop_back();
    for (const ConnectionPtr& c : n->getChildConnections())
    {
      all.push_back(c);
      stack.push_back(c->getChild());
    }
  }
  prefetch(chk, all);
  return tree->recheckCollision();
}

// N1: Tree::checkPathToNode (tree.cpp:335-373): the connections root -> node that are not flagged "recently checked"
inline bool checkPathToNode(const TreePtr& tree, const NodePtr& node, std::vector<ConnectionPtr>& checked_connections,
                            std::vector<ConnectionPtr>& path_connections, const CudaCheckerPtr& chk)
{
  std::vector<ConnectionPtr> todo;
  for (const ConnectionPtr& c : tree->getConnectionToNode(node))
  {
    if (!c->isRecentlyChecked())
      todo.push_back(c);
    else if (c->getCost() == std::numeric_limits<double>::infinity())
    {
      todo.clear();  // the reference returns false before checking anything (tree.cpp:345-349)
      break;
    }
  }
  prefetch(chk, todo);
  return tree->checkPathToNode(node, checked_connections, path_connections);
}

// N1: Tree::rewireOnlyWithPathCheck (tree.cpp:522-679): rewiring around `node` on a tree whose connections may have
// become invalid (replanning).  The function asks, one at a time and depending on the costs met so far, for
//   * the connections root -> node and root -> n of the near nodes n it considers (checkPathToNode, not yet flagged),
//   * the edge n -> node of every parent candidate and the edge node -> n of every child candidate.
// All of them are known before the walk starts -- the near set, both directions of every (node, near node) pair and
// every unflagged connection on the way to those nodes -- so they go to the device as ONE batch; the reference's own
// function then runs unchanged on the cached verdicts.  r_rewire <= 0 selects nearK as in the reference (:557-560).
inline bool rewireOnlyWithPathCheck(const TreePtr& tree, NodePtr& node, std::vector<ConnectionPtr>& checked_connections,
                                    double r_rewire, const std::vector<NodePtr>& white_list, const int& what_rewire,
                                    const CudaCheckerPtr& chk)
{
  const std::multimap<double, NodePtr> near_nodes = (r_rewire <= 0) ? tree->nearK(node) : tree->near(node, r_rewire);
  std::vector<const double*> from, to;
  std::vector<ConnectionPtr> on_paths;
  std::vector<Connection*> seen;
  auto addPath = [&](const NodePtr& n) {
    if (n == tree->getRoot())
      return;
    for (const ConnectionPtr& c : tree->getConnectionToNode(n))
      if (!c->isRecentlyChecked() && std::find(seen.begin(), seen.end(), c.get()) == seen.end())
      {
        seen.push_back(c.get());
        on_paths.push_back(c);
      }
  };
  addPath(node);
  const double* nc = node->getConfiguration().data();
  for (const std::pair<const double, NodePtr>& p : near_nodes)
  {
    const NodePtr& n = p.second;
    if (n == node)
      continue;
    addPath(n);
    if (what_rewire != 2)  // parent phase: checkConnection(n, node)  (:589)
    {
      from.push_back(n->getConfiguration().data());
      to.push_back(nc);
    }
    if (what_rewire != 1)  // children phase: checkConnection(node, n)  (:661)
    {
      from.push_back(nc);
      to.push_back(n->getConfiguration().data());
    }
  }
  for (const ConnectionPtr& c : on_paths)
  {
    from.push_back(c->getParent()->getConfiguration().data());
    to.push_back(c->getChild()->getConfiguration().data());
  }
  chk->prefetchConnections(from, to);
  return tree->rewireOnlyWithPathCheck(node, checked_connections, r_rewire, white_list, what_rewire);
}

// N3: PathLocalOptimizer with speculative batches.  bisection() (path_local_optimizer.cpp:60-136) tries at most five
// mid points; which one comes next depends on the previous verdict, so the candidates form a binary tree of depth
// five over [min_distance, max_distance] whose 31 abscissae depend only on the arguments.  All of them (2 edges each:
// parent -> p and p -> child) are checked in one device call before the reference's own bisection() runs.
class SpeculativePathLocalOptimizer : public PathLocalOptimizer
{
public:
  SpeculativePathLocalOptimizer(const CudaCheckerPtr& checker, const MetricsPtr& metrics,
                                const cnr_logger::TraceLoggerPtr& logger)
    : PathLocalOptimizer(checker, metrics, logger), cuda_checker_(checker)
  {
  }

  // PathLocalOptimizer::warp (path_local_optimizer.cpp:138-172) with the speculation in front of every bisection
  bool warpSpeculative(const double& min_conn_length = 0.1, const double min_step_size = 0.01,
                       const double& max_time = std::numeric_limits<double>::infinity())
  {
    if (max_time > 0)
    {
      std::vector<ConnectionPtr> connections = path_->getConnections();
      auto tic = graph_time::now();
      for (unsigned int idx = 1; idx < connections.size(); idx++)
      {
        if (connections.at(idx - 1)->norm() > min_conn_length && connections.at(idx)->norm() > min_conn_length)
        {
          if (change_warp_.at(idx - 1) || change_warp_.at(idx))
          {
            Eigen::VectorXd center = 0.5 * (connections.at(idx - 1)->getParent()->getConfiguration() +
                                            connections.at(idx)->getChild()->getConfiguration());
            Eigen::VectorXd direction = connections.at(idx - 1)->getChild()->getConfiguration() - center;
            double max_distance = direction.norm();
            double min_distance = 0;
            direction.normalize();
            speculate(idx, center, direction, min_step_size, max_distance, min_distance);
            change_warp_.at(idx) = bisection(idx, center, direction, min_step_size, max_distance, min_distance);
          }
        }
        if (toSeconds(graph_time::now(), tic) >= 0.98 * max_time)
          break;
      }
    }
    return std::any_of(change_warp_.cbegin(), change_warp_.cend(), [](bool i) { return i; });
  }

  // PathLocalOptimizer::simplify (path_local_optimizer.cpp:174-228): the short-cut edges of the path as it stands go
  // as one batch; a short cut taken changes the neighbours' candidates, which then miss the cache (one call each)
  bool simplifySpeculative(const double& min_conn_length = 0.1)
  {
    const std::vector<ConnectionPtr>& c = path_->getConnectionsConst();
    std::vector<const double*> from, to;
    for (size_t ic = 1; ic < c.size(); ic++)
    {
      from.push_back(c[ic - 1]->getParent()->getConfiguration().data());
      to.push_back(c[ic]->getChild()->getConfiguration().data());
    }
    cuda_checker_->prefetchConnections(from, to);
    return simplify(min_conn_length);
  }

  // PathLocalOptimizer::step (path_local_optimizer.cpp:230-258)
  bool step() override
  {
    if (not path_)
      return false;
    bool improved = false;
    if (solved_)
      return false;
    double cost = path_->cost();
    bool solved = not warpSpeculative(warp_min_conn_length_, warp_min_step_size_);
    if (cost <= (1.001 * path_->cost()))
    {
      if (stall_gen_ == 0)
        simplifySpeculative(simplify_max_conn_length_) ? stall_gen_++ : solved = false;
      else
        stall_gen_++;
    }
    else
    {
      improved = true;
      stall_gen_ = 0;
    }
    solved_ = solved || (stall_gen_ >= max_stall_gen_);
    return improved;
  }

protected:
  void candidates(double min_distance, double max_distance, unsigned int iter, double min_step_size,
                  std::vector<double>& out) const
  {
    // while ((iter++ < 5) && ((max_distance - min_distance) > min_step_size))   (path_local_optimizer.cpp:80)
    if (!(iter < 5) || !((max_distance - min_distance) > min_step_size))
      return;
    const double distance = 0.5 * (max_distance + min_distance);
    out.push_back(distance);
    candidates(distance, max_distance, iter + 1, min_step_size, out);  // rejected: min_distance = distance
    candidates(min_distance, distance, iter + 1, min_step_size, out);  // accepted: max_distance = distance
  }

  void speculate(const size_t& connection_idx, const Eigen::VectorXd& center, const Eigen::VectorXd& direction,
                 const double min_step_size, double max_distance, double min_distance)
  {
    const std::vector<ConnectionPtr>& connections = path_->getConnectionsConst();
    const Eigen::VectorXd& parent = connections.at(connection_idx - 1)->getParent()->getConfiguration();
    const Eigen::VectorXd& child = connections.at(connection_idx)->getChild()->getConfiguration();
    std::vector<double> ds;
    candidates(min_distance, max_distance, 0, min_step_size, ds);
    points_.clear();
    points_.reserve(ds.size());
    for (double distance : ds)
    {
      Eigen::VectorXd p = center + direction * distance;  // the expression of path_local_optimizer.cpp:87
      points_.push_back(p);
    }
    std::vector<const double*> from, to;
    for (const Eigen::VectorXd& p : points_)
    {
      from.push_back(parent.data());
      to.push_back(p.data());
      from.push_back(p.data());
      to.push_back(child.data());
    }
    cuda_checker_->prefetchConnections(from, to);
  }

  CudaCheckerPtr cuda_checker_;
  std::vector<Eigen::VectorXd> points_;
};

}  // namespace cuda_batched
}  // namespace core
}  // namespace graph

#endif  // GCB200_WITH_GRAPH_CORE
