// plugin_probe.cpp -- exercises the drop-in plugin classes the way graph_core's Tree and solvers use
// NearestNeighbors and CollisionCheckerBase (src/graph_core/graph/tree.cpp:56-59, 751-765, 815-825;
// collision_checker_base.h:166-313), through the cnr_class_loader factory names.
// Exit code 0 and "PROBE OK" on success.  Needs a CUDA device: without one it must fail loudly.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "cuda_plugins.h"

using namespace graph::core;

static int g_fail = 0;
#define CHECK(cond, ...)                          \
  do                                              \
  {                                               \
    if (!(cond))                                  \
    {                                             \
      g_fail++;                                   \
      std::printf("FAIL %s:%d: ", __FILE__, __LINE__); \
      std::printf(__VA_ARGS__);                   \
      std::printf("\n");                          \
    }                                             \
  } while (0)

static unsigned long long g_rng = 88172645463325252ull;
static double urand()
{
  g_rng ^= g_rng << 13;
  g_rng ^= g_rng >> 7;
  g_rng ^= g_rng << 17;
  return (double)(g_rng >> 11) * (1.0 / 9007199254740992.0);
}
static Eigen::VectorXd randomConf(int d, double lo, double hi)
{
  Eigen::VectorXd v(d);
  for (int i = 0; i < d; i++)
    v(i) = lo + (hi - lo) * urand();
  return v;
}
static double norm(const Eigen::VectorXd& a, const Eigen::VectorXd& b)
{
  double s = 0.0;
  for (std::ptrdiff_t i = 0; i < a.size(); i++)
  {
    const double t = a(i) - b(i);
    s = s + t * t;
  }
  return std::sqrt(s);
}

// the reference's state set of checkConnection(c1,c2) in its visiting order (collision_checker_base.h:207-237),
// evaluated through the public single-state check()
static bool bisectByStates(CollisionCheckerBase& ch, const Eigen::VectorXd& c1, const Eigen::VectorXd& c2, double min_d)
{
  if (!ch.check(c1) || !ch.check(c2))
    return false;
  const double distance = norm(c2, c1);
  if (distance < min_d)
    return true;
  const int d = (int)c1.size();
  double n = 2;
  while (distance > n * min_d)
  {
    for (double idx = 1; idx < n; idx += 2)
    {
      Eigen::VectorXd conf(d);
      for (int i = 0; i < d; i++)
        conf(i) = c1(i) + ((c2(i) - c1(i)) * idx) / n;
      if (!ch.check(conf))
        return false;
    }
    n *= 2;
  }
  return true;
}

static void testNearestNeighbors()
{
  cnr_logger::TraceLoggerPtr logger;
  cnr::param::set("/probe/nn/capacity", 64.0);  // small on purpose: the store has to grow
  cnr::param::set("/probe/nn/device_id", 0.0);
  auto plugin = cnr_class_loader::createInstance<NearestNeighborsPlugin>("graph::core::CudaNearestNeighborsPlugin");
  CHECK(plugin != nullptr, "factory name not registered");
  if (!plugin)
    return;
  CHECK(plugin->init("/probe/nn", logger), "plugin init failed");
  NearestNeighborsPtr nn = plugin->getNearestNeighbors();
  if (!nn)
    return;
  const int d = 7, N = 3000;
  NodePtr best;
  double bd = 0.0;
  nn->nearestNeighbor(randomConf(d, -3, 3), best, bd);
  CHECK(!best && std::isinf(bd), "empty structure must give (nullptr, +inf)");  // kdtree.cpp:356-358
  std::vector<NodePtr> nodes;
  for (int i = 0; i < N; i++)
  {
    nodes.push_back(std::make_shared<Node>(randomConf(d, -3.14, 3.14), logger));
    nn->insert(nodes.back());
  }
  nodes.push_back(std::make_shared<Node>(nodes[10]->getConfiguration(), logger));  // exact duplicate of node 10
  nn->insert(nodes.back());
  CHECK(nn->size() == (unsigned)N + 1, "size %u", nn->size());
  for (int t = 0; t < 40; t++)
  {
    const Eigen::VectorXd q = (t == 0) ? nodes[10]->getConfiguration() : randomConf(d, -3.14, 3.14);
    // brute force in insertion order, strict '<' (vector.cpp:55-67)
    NodePtr want;
    double wd = std::numeric_limits<double>::infinity();
    std::multimap<double, NodePtr> all;
    for (const NodePtr& n : nodes)
    {
      const double dist = norm(n->getConfiguration(), q);
      if (dist < wd)
      {
        wd = dist;
        want = n;
      }
      all.insert(std::make_pair(dist, n));
    }
    nn->nearestNeighbor(q, best, bd);
    CHECK(best == want && bd == wd, "nearest mismatch at query %d (%.17g vs %.17g)", t, bd, wd);
    CHECK(nn->nearestNeighbor(q) == want, "convenience overload mismatch");
    const double radius = 2.2;
    std::multimap<double, NodePtr> got = nn->near(q, radius);
    std::multimap<double, NodePtr> exp;
    for (const NodePtr& n : nodes)
    {
      const double dist = norm(n->getConfiguration(), q);
      if (dist < radius)
        exp.insert(std::make_pair(dist, n));
    }
    CHECK(got.size() == exp.size(), "near size %zu vs %zu", got.size(), exp.size());
    auto a = got.begin();
    auto b = exp.begin();
    for (; a != got.end() && b != exp.end(); ++a, ++b)
      CHECK(a->first == b->first && a->second == b->second, "near order mismatch");
    const size_t k = (t % 2) ? 9 : 100000;  // Tree::nearK asks for huge k (tree.cpp:761-765)
    std::multimap<double, NodePtr> kg = nn->kNearestNeighbors(q, k);
    CHECK(kg.size() == std::min(k, all.size()), "knn size");
    a = kg.begin();
    b = all.begin();
    for (; a != kg.end(); ++a, ++b)
      CHECK(a->first == b->first && a->second == b->second, "knn order mismatch");
  }
  // lazy deletion (kdtree.cpp:395-428)
  CHECK(nn->findNode(nodes[5]), "findNode");
  CHECK(nn->deleteNode(nodes[5], true), "deleteNode");
  CHECK(nodes[5]->disconnected_, "disconnect_node flag must reach the node");
  CHECK(!nn->deleteNode(nodes[5]), "second delete of a tombstone is false");
  CHECK(!nn->findNode(nodes[5]) && nn->size() == (unsigned)N, "tombstone bookkeeping");
  nn->nearestNeighbor(nodes[5]->getConfiguration(), best, bd);
  CHECK(best != nodes[5] && bd > 0.0, "a tombstoned node must not be returned");
  CHECK(nn->getNodes().size() == (size_t)N, "getNodes skips tombstones");
  nn->insert(nodes[5]);  // KdNode::insert of a stored pointer restores it (kdtree.cpp:93-97)
  nn->nearestNeighbor(nodes[5]->getConfiguration(), best, bd);
  CHECK(best == nodes[5] && bd == 0.0 && nn->size() == (unsigned)N + 1, "re-insert restores");
  std::vector<NodePtr> white = { nodes[0] };
  nn->disconnectNodes(white);
  CHECK(!nodes[0]->disconnected_ && nodes[1]->disconnected_, "disconnectNodes honours the white list");
  CHECK(nn->clear() && nn->size() == 0 && nn->getNodes().empty(), "clear");
  nn->nearestNeighbor(randomConf(d, -3, 3), best = nullptr, bd);
  CHECK(!best && std::isinf(bd), "cleared structure");
}

static void testCollisionChecker()
{
  cnr_logger::TraceLoggerPtr logger;
  // the reference's own test world: Cube3dCollisionChecker, threshold 1.0 (cube_3d_collision_checker.h:61-64)
  cnr::param::set("/probe/cc/world_kind", std::string("cube"));
  cnr::param::set("/probe/cc/dimension", 3.0);
  cnr::param::set("/probe/cc/threshold", 1.0);
  auto plugin =
      cnr_class_loader::createInstance<CollisionCheckerBasePlugin>("graph::core::CudaCollisionCheckerPlugin");
  CHECK(plugin != nullptr, "factory name not registered");
  if (!plugin)
    return;
  const double min_d = 0.01;
  CHECK(plugin->init("/probe/cc", logger, min_d), "plugin init failed");
  CollisionCheckerPtr ch = plugin->getCollisionChecker();
  if (!ch)
    return;
  int nfree = 0, nblocked = 0;
  for (int t = 0; t < 60; t++)
  {
    const Eigen::VectorXd q = randomConf(3, -2.5, 2.5);
    double m = 0;
    for (int i = 0; i < 3; i++)
      m = std::max(m, std::fabs(q(i)));
    CHECK(ch->check(q) == (m > 1.0), "cube check");
    const Eigen::VectorXd c1 = randomConf(3, -2.5, 2.5), c2 = randomConf(3, -2.5, 2.5);
    const bool want = bisectByStates(*ch, c1, c2, min_d);
    const bool got = ch->checkConnection(c1, c2);
    CHECK(got == want, "checkConnection(c1,c2) vs state-by-state walk");
    (got ? nfree : nblocked)++;
    // linear variant (:178-205)
    Eigen::VectorXd conf({ 99.0, 99.0, 99.0 });
    const bool lin = ch->checkConnection(c1, c2, conf);
    bool wlin = ch->check(c1);
    Eigen::VectorXd wconf({ 99.0, 99.0, 99.0 });
    if (wlin)
    {
      const double dist = norm(c2, c1);
      const unsigned int npnt = (unsigned int)std::ceil(dist / min_d);
      if (npnt > 1)
      {
        for (unsigned int ip = 1; ip < npnt && wlin; ip++)
        {
          Eigen::VectorXd p(3);
          for (int i = 0; i < 3; i++)
            p(i) = c1(i) + ((c2(i) - c1(i)) / (double)npnt) * (double)ip;
          wconf = p;
          if (!ch->check(p))
          {
            for (int i = 0; i < 3; i++)
              wconf(i) = c1(i) + ((c2(i) - c1(i)) / (double)npnt) * (double)(ip - 1);
            wlin = false;
          }
        }
      }
      if (wlin && !ch->check(c2))
        wlin = false;
    }
    CHECK(lin == wlin, "linear checkConnection boolean");
    if (!lin)
      for (int i = 0; i < 3; i++)
        CHECK(conf(i) == wconf(i), "linear checkConnection last feasible configuration");
    // connection overloads
    NodePtr p = std::make_shared<Node>(c1, logger), c = std::make_shared<Node>(c2, logger);
    ConnectionPtr conn = std::make_shared<Connection>(p, c, logger);
    CHECK(ch->checkConnection(conn) == want, "checkConnection(ConnectionPtr)");
    Eigen::VectorXd mid(3);
    for (int i = 0; i < 3; i++)
      mid(i) = c1(i) + (c2(i) - c1(i)) * 0.5;
    const bool fc = ch->checkConnFromConf(conn, mid);
    if (want)
      CHECK(fc, "a free connection is free from any configuration on it");
  }
  CHECK(nfree > 5 && nblocked > 5, "both outcomes must occur (%d free, %d blocked)", nfree, nblocked);
  // checkConnections: AND over a chain (:249-255) -- tests/path.yaml's first waypoints
  std::vector<Eigen::VectorXd> wp = { { -1.5, -1.5, -1.5 },
                                      { -1.368304943282731, -0.869310126335348, -0.7325118657666546 },
                                      { -1.113252883580279, -0.09815786150562764, 0.5658478923234942 },
                                      { -0.986172691213979, 0.05930640853909153, 1.050255684438281 } };
  std::vector<ConnectionPtr> chain;
  std::vector<NodePtr> keep;
  for (auto& v : wp)
    keep.push_back(std::make_shared<Node>(v, logger));
  for (size_t i = 0; i + 1 < keep.size(); i++)
    chain.push_back(std::make_shared<Connection>(keep[i], keep[i + 1], logger));
  CHECK(ch->checkConnections(chain), "reference path.yaml prefix must be collision free");
  NodePtr inside = std::make_shared<Node>(Eigen::VectorXd({ 0.0, 0.0, 0.0 }), logger);
  chain.push_back(std::make_shared<Connection>(keep.back(), inside, logger));
  CHECK(!ch->checkConnections(chain), "a chain entering the cube is rejected");
  // clone shares the device world (:319)
  CollisionCheckerPtr cl = ch->clone();
  CHECK(cl->check(wp[0]) && !cl->check(Eigen::VectorXd({ 0.1, 0.2, 0.3 })), "clone");
}

static void testSpeculativeRewire()
{
  cnr_logger::TraceLoggerPtr logger;
  // small sphere-arm world: 7 joints, 8 robot spheres, 200 obstacles
  CudaWorldDescription w;
  w.kind = GC_WORLD_SPHERE_ARM;
  w.dim = 7;
  w.arm_base_tf = { 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0 };
  for (int j = 0; j < 7; j++)
  {
    const float sg = (j % 2) ? -1.f : 1.f;
    const float L = (j % 2) ? 0.f : 0.25f;
    const float f[12] = { 1, 0, 0, 0, 0, -sg, 0, sg, 0, 0, 0, L };
    w.arm_joint_tf.insert(w.arm_joint_tf.end(), f, f + 12);
  }
  for (int k = 0; k < 8; k++)
  {
    w.arm_sphere_link.push_back(k);
    const float s[4] = { 0.f, 0.f, 0.08f, 0.06f };
    w.arm_sphere_local.insert(w.arm_sphere_local.end(), s, s + 4);
  }
  for (int o = 0; o < 200; o++)
  {
    float c[4] = { (float)(-1 + 2 * urand()), (float)(-1 + 2 * urand()), (float)(-1 + 2 * urand()), 0.04f };
    if (std::fabs(c[0]) < 0.3f && std::fabs(c[1]) < 0.3f)
      c[2] = 1.5f;  // keep the base column free
    w.obstacles.insert(w.obstacles.end(), c, c + 4);
  }
  auto checker = std::make_shared<CudaCollisionChecker>(logger, w, 0, 0.01);
  auto nn = std::make_shared<CudaNearestNeighbors>(logger, 0, 256);
  std::vector<NodePtr> nodes;
  for (int i = 0; i < 400; i++)
  {
    nodes.push_back(std::make_shared<Node>(randomConf(7, -1.5, 1.5), logger));
    nn->insert(nodes.back());
  }
  const Eigen::VectorXd q = randomConf(7, -1.0, 1.0);
  std::multimap<double, NodePtr> plain = nn->near(q, 1.6);
  std::vector<bool> direct_to, direct_from;
  for (auto& h : plain)  // what Tree::rewireOnly does one edge at a time (tree.cpp:450, 498)
  {
    direct_to.push_back(checker->checkConnection(h.second->getConfiguration(), q));
    direct_from.push_back(checker->checkConnection(q, h.second->getConfiguration()));
  }
  const uint64_t calls_plain = checker->deviceCalls();
  nn->setSpeculativeChecker(checker);
  std::multimap<double, NodePtr> spec = nn->near(q, 1.6);
  CHECK(spec.size() == plain.size() && plain.size() > 4, "near set size %zu", plain.size());
  const uint64_t calls_before = checker->deviceCalls();
  size_t i = 0;
  for (auto& h : spec)
  {
    CHECK(checker->checkConnection(h.second->getConfiguration(), q) == direct_to[i], "cached n->q differs");
    CHECK(checker->checkConnection(q, h.second->getConfiguration()) == direct_from[i], "cached q->n differs");
    i++;
  }
  CHECK(checker->deviceCalls() == calls_before, "the rewire edge checks must be answered from the batch");
  CHECK(checker->cacheHits() == 2 * spec.size(), "cache hits %llu", (unsigned long long)checker->cacheHits());
  CHECK(calls_plain >= 2 * plain.size(), "without speculation every edge is a device call");
  // batched states == single states
  std::vector<Eigen::VectorXd> qs;
  for (int t = 0; t < 100; t++)
    qs.push_back(randomConf(7, -3.14, 3.14));
  std::vector<bool> free;
  checker->checkStates(qs, free);
  int nf = 0;
  for (int t = 0; t < 100; t++)
  {
    CHECK(free[t] == checker->check(qs[t]), "batched vs single state check");
    nf += free[t];
  }
  CHECK(nf > 0 && nf < 100, "both outcomes (%d free)", nf);
}

// Tree::nearK on a large high-dimensional tree (tree.cpp:52-53, 761-765): k = ceil(k_rrt * ln(N + 1)) >= N, so
// kNearestNeighbors returns every node in ascending (distance, insertion) order -- lists far beyond one CTA's sort
// capacity -- and near() with a radius that catches most of the tree does the same.
static void testLargeResultSets()
{
  cnr_logger::TraceLoggerPtr logger;
  cnr::param::set("/probe/nnl/capacity", 65536.0);
  cnr::param::set("/probe/nnl/device_id", 0.0);
  auto plugin = cnr_class_loader::createInstance<NearestNeighborsPlugin>("graph::core::CudaNearestNeighborsPlugin");
  if (!plugin || !plugin->init("/probe/nnl", logger))
  {
    CHECK(false, "plugin init failed");
    return;
  }
  NearestNeighborsPtr nn = plugin->getNearestNeighbors();
  const int d = 12, N = 40000;
  std::vector<NodePtr> nodes;
  for (int i = 0; i < N; i++)
  {
    nodes.push_back(std::make_shared<Node>(randomConf(d, -3.14, 3.14), logger));
    nn->insert(nodes.back());
  }
  const Eigen::VectorXd q = randomConf(d, -3.14, 3.14);
  const double k_rrt = 1.1 * std::pow(2.0, d + 1) * std::exp(1.0) * (1.0 + 1.0 / d);
  const size_t k = (size_t)std::ceil(k_rrt * std::log(N + 1.0));
  std::multimap<double, NodePtr> all = nn->kNearestNeighbors(q, k);
  CHECK(all.size() == (size_t)N, "k >= N returns every node (%zu)", all.size());
  std::vector<std::pair<double, int>> want;
  for (int i = 0; i < N; i++)
    want.emplace_back(norm(nodes[i]->getConfiguration(), q), i);
  std::sort(want.begin(), want.end());
  size_t j = 0;
  bool same = all.size() == (size_t)N;
  for (auto it = all.begin(); same && it != all.end(); ++it, ++j)
    same = it->first == want[j].first && it->second == nodes[want[j].second];
  CHECK(same, "k >= N order mismatch at rank %zu", j);
  const double radius = want[30000].first;  // strict <: exactly the 30000 nearer nodes (no ties in random data)
  std::multimap<double, NodePtr> nr = nn->near(q, radius);
  CHECK(nr.size() == 30000, "near() with %zu hits", nr.size());
  j = 0;
  same = nr.size() == 30000;
  for (auto it = nr.begin(); same && it != nr.end(); ++it, ++j)
    same = it->first == want[j].first && it->second == nodes[want[j].second];
  CHECK(same, "large near() order mismatch at rank %zu", j);
}

int main()
{
  try
  {
    testNearestNeighbors();
    testCollisionChecker();
    testSpeculativeRewire();
    testLargeResultSets();
  }
  catch (const std::exception& e)
  {
    std::printf("PROBE EXCEPTION: %s\n", e.what());
    return 2;
  }
  if (g_fail)
  {
    std::printf("PROBE FAILED: %d checks\n", g_fail);
    return 1;
  }
  std::printf("PROBE OK\n");
  return 0;
}
