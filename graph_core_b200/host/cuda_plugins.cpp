// cuda_plugins.cpp -- see cuda_plugins.h.  Host logic only; every distance and collision test runs in
// libgcb200.so (CUDA) through the C ABI.
#include "cuda_plugins.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace graph
{
namespace core
{
namespace
{
[[noreturn]] void raise(const char* what)
{
  throw std::runtime_error(std::string(what) + ": " + gc_last_error());
}
void copyConf(const Eigen::VectorXd& v, double* out)
{
  for (std::ptrdiff_t i = 0; i < v.size(); i++)
    out[i] = v(i);
}
}  // namespace

// =================================================================================================
//  CudaNearestNeighbors
// =================================================================================================
CudaNearestNeighbors::CudaNearestNeighbors(const cnr_logger::TraceLoggerPtr& logger, int device, uint32_t capacity)
  : NearestNeighbors(logger), device_(device), capacity_(std::max<uint32_t>(capacity, 32))
{
  int ndev = 0;
  if (gc_device_count(&ndev) != GC_OK || device < 0 || device >= ndev)
    raise("CudaNearestNeighbors: no usable CUDA device (there is no CPU fallback)");
}

CudaNearestNeighbors::~CudaNearestNeighbors()
{
  if (store_)
    gc_store_destroy(store_);
}

void CudaNearestNeighbors::ensureStore(int dim)
{
  if (store_)
  {
    if (dim != dim_)
      throw std::invalid_argument("CudaNearestNeighbors: configuration size differs from the stored nodes");
    return;
  }
  dim_ = dim;
  if (gc_store_create(dim_, capacity_, 1, device_, &store_) != GC_OK)
    raise("gc_store_create");
}

void CudaNearestNeighbors::grow()
{
  const uint32_t used = (uint32_t)node_of_slot_.size();
  std::vector<double> rows((size_t)used * dim_);
  if (used && gc_store_read(store_, 0, 0, used, rows.data()) != GC_OK)
    raise("gc_store_read");
  gc_store_t* bigger = nullptr;
  capacity_ *= 2;
  if (gc_store_create(dim_, capacity_, 1, device_, &bigger) != GC_OK)
    raise("gc_store_create (grow)");
  uint32_t first = 0;
  if (used && gc_store_append(bigger, 0, rows.data(), used, &first) != GC_OK)
    raise("gc_store_append (grow)");
  for (uint32_t s = 0; s < used; s++)
    if (dead_[s] && gc_store_tombstone(bigger, 0, s) < 0)
      raise("gc_store_tombstone (grow)");
  gc_store_destroy(store_);
  store_ = bigger;
}

// KdTree::insert (kdtree.cpp:327-342): re-inserting a tombstoned node restores it
void CudaNearestNeighbors::insert(const NodePtr& node)
{
  ensureStore((int)node->getConfiguration().size());
  auto it = slot_of_.find(node.get());
  if (it != slot_of_.end())
  {
    if (dead_[it->second])
    {
      if (gc_store_restore(store_, 0, it->second) < 0)
        raise("gc_store_restore");
      dead_[it->second] = 0;
      size_++;
      deleted_nodes_--;
    }
    else
      CNR_WARN(logger_, "node already stored, insert ignored");
    return;
  }
  if (node_of_slot_.size() >= capacity_)
    grow();
  uint32_t slot = 0;
  if (gc_store_append(store_, 0, node->getConfiguration().data(), 1, &slot) != GC_OK)
    raise("gc_store_append");
  node_of_slot_.push_back(node);
  dead_.push_back(0);
  slot_of_[node.get()] = slot;
  size_++;
}

bool CudaNearestNeighbors::clear()
{
  size_ = 0;
  deleted_nodes_ = 0;
  node_of_slot_.clear();
  dead_.clear();
  slot_of_.clear();
  if (store_ && gc_store_clear(store_, GC_NO_NODE) != GC_OK)
    raise("gc_store_clear");
  return true;
}

void CudaNearestNeighbors::nearestNeighbor(const Eigen::VectorXd& configuration, NodePtr& best, double& best_distance)
{
  best_distance = std::numeric_limits<double>::infinity();  // kdtree.cpp:356
  if (!store_ || size_ == 0)
    return;
  uint32_t idx = GC_NO_NODE;
  double d = best_distance;
  if (gc_nn1(store_, configuration.data(), nullptr, 1, &idx, &d) != GC_OK)
    raise("gc_nn1");
  if (idx != GC_NO_NODE)
  {
    best = node_of_slot_[idx];
    best_distance = d;
  }
}

void CudaNearestNeighbors::nearestNeighbors(const std::vector<Eigen::VectorXd>& configurations,
                                            std::vector<NodePtr>& best, std::vector<double>& best_distance)
{
  const uint32_t n = (uint32_t)configurations.size();
  best.assign(n, nullptr);
  best_distance.assign(n, std::numeric_limits<double>::infinity());
  if (!store_ || size_ == 0 || n == 0)
    return;
  std::vector<double> q((size_t)n * dim_);
  for (uint32_t i = 0; i < n; i++)
    copyConf(configurations[i], &q[(size_t)i * dim_]);
  std::vector<uint32_t> idx(n);
  if (gc_nn1(store_, q.data(), nullptr, n, idx.data(), best_distance.data()) != GC_OK)
    raise("gc_nn1");
  for (uint32_t i = 0; i < n; i++)
    if (idx[i] != GC_NO_NODE)
      best[i] = node_of_slot_[idx[i]];
}

// the device lists are already in (distance, slot) order = the order a multimap filled in scan order has
std::multimap<double, NodePtr> CudaNearestNeighbors::toMap(uint32_t count, const uint32_t* idx, const double* dist)
{
  std::multimap<double, NodePtr> out;
  for (uint32_t i = 0; i < count; i++)
    out.emplace_hint(out.end(), dist[i], node_of_slot_[idx[i]]);
  return out;
}

std::multimap<double, NodePtr> CudaNearestNeighbors::near(const Eigen::VectorXd& configuration, const double& radius)
{
  std::multimap<double, NodePtr> out;
  if (!store_ || size_ == 0)
    return out;
  uint32_t cap = std::max<uint32_t>(64, (uint32_t)idx_.size());
  for (;;)
  {
    idx_.resize(cap);
    dist_.resize(cap);
    uint32_t count = 0;
    const int rc = gc_radius(store_, configuration.data(), nullptr, &radius, 1, cap, idx_.data(), dist_.data(), &count);
    if (rc < 0)
      raise("gc_radius");
    if (rc != GC_W_TRUNCATED)
    {
      out = toMap(count, idx_.data(), dist_.data());
      break;
    }
    while (cap < count)
      cap *= 2;
  }
  if (spec_)
    speculate(configuration, out);
  return out;
}

std::multimap<double, NodePtr> CudaNearestNeighbors::kNearestNeighbors(const Eigen::VectorXd& configuration,
                                                                       const size_t& k)
{
  std::multimap<double, NodePtr> out;
  if (!store_ || size_ == 0 || k == 0)
    return out;
  const uint32_t cap = (uint32_t)std::min<size_t>(k, size_);
  idx_.resize(std::max<size_t>(idx_.size(), cap));
  dist_.resize(std::max<size_t>(dist_.size(), cap));
  uint32_t count = 0;
  if (gc_knn(store_, configuration.data(), nullptr, 1, (uint64_t)k, cap, idx_.data(), dist_.data(), &count) < 0)
    raise("gc_knn");
  out = toMap(std::min(count, cap), idx_.data(), dist_.data());
  if (spec_)
    speculate(configuration, out);
  return out;
}

void CudaNearestNeighbors::speculate(const Eigen::VectorXd& q, const std::multimap<double, NodePtr>& hits)
{
  std::vector<const double*> from, to;
  from.reserve(2 * hits.size());
  to.reserve(2 * hits.size());
  for (const auto& h : hits)
  {
    const double* n = h.second->getConfiguration().data();
    if (h.first == 0.0)
      continue;  // the query node itself
    from.push_back(n);  // parent phase of rewireOnly: n -> q   (tree.cpp:450)
    to.push_back(q.data());
    from.push_back(q.data());  // children phase: q -> n           (tree.cpp:498)
    to.push_back(n);
  }
  spec_->prefetchConnections(from, to);
}

bool CudaNearestNeighbors::findNode(const NodePtr& node)
{
  auto it = slot_of_.find(node.get());
  return it != slot_of_.end() && !dead_[it->second];  // a tombstoned node is "not found" (kdtree.cpp:236-239)
}

// KdTree::deleteNode (kdtree.cpp:395-428): lazy deletion; false when the node is absent or already deleted
bool CudaNearestNeighbors::deleteNode(const NodePtr& node, const bool& disconnect_node)
{
  auto it = slot_of_.find(node.get());
  if (it == slot_of_.end() || dead_[it->second])
    return false;
  if (gc_store_tombstone(store_, 0, it->second) < 0)
    raise("gc_store_tombstone");
  dead_[it->second] = 1;
  size_--;
  deleted_nodes_++;
  if (disconnect_node)
    node->disconnect();
  return true;
}

// The reference never restores (KdTree::restoreNode searches with findNode, which skips tombstones,
// kdtree.cpp:447-456; Vector::restoreNode returns false, vector.cpp:121-124) and nothing in graph_core
// calls it.  Here a tombstoned node does come back, which is what the interface documents.
bool CudaNearestNeighbors::restoreNode(const NodePtr& node)
{
  auto it = slot_of_.find(node.get());
  if (it == slot_of_.end() || !dead_[it->second])
    return false;
  if (gc_store_restore(store_, 0, it->second) < 0)
    raise("gc_store_restore");
  dead_[it->second] = 0;
  size_++;
  deleted_nodes_--;
  return true;
}

std::vector<NodePtr> CudaNearestNeighbors::getNodes()
{
  std::vector<NodePtr> out;
  out.reserve(size_);
  for (size_t s = 0; s < node_of_slot_.size(); s++)
    if (!dead_[s])
      out.push_back(node_of_slot_[s]);
  return out;
}

void CudaNearestNeighbors::disconnectNodes(const std::vector<NodePtr>& white_list)
{
  for (size_t s = 0; s < node_of_slot_.size(); s++)
    if (!dead_[s] && std::find(white_list.begin(), white_list.end(), node_of_slot_[s]) == white_list.end())
      node_of_slot_[s]->disconnect();
}

// =================================================================================================
//  CudaCollisionChecker
// =================================================================================================
CudaCollisionChecker::CudaCollisionChecker(const cnr_logger::TraceLoggerPtr& logger, const CudaWorldDescription& w,
                                           int device, const double& min_distance)
  : CollisionCheckerBase(logger, min_distance), dim_(w.dim)
{
  int rc = GC_E_INVALID;
  switch (w.kind)
  {
    case GC_WORLD_CUBE:
      rc = gc_world_create_cube(w.dim, w.cube_threshold, device, &world_);
      break;
    case GC_WORLD_BOXES:
      rc = gc_world_create_boxes(w.dim, (uint32_t)(w.boxes_lo.size() / std::max(1, w.dim)), w.boxes_lo.data(),
                                 w.boxes_hi.data(), device, &world_);
      break;
    case GC_WORLD_CSPHERES:
      rc = gc_world_create_cspheres(w.dim, (uint32_t)w.sphere_radii.size(), w.sphere_centers.data(),
                                    w.sphere_radii.data(), device, &world_);
      break;
    case GC_WORLD_SPHERE_ARM:
      rc = gc_world_create_sphere_arm(w.dim, w.arm_base_tf.data(), w.arm_joint_tf.data(),
                                      (uint32_t)w.arm_sphere_link.size(), w.arm_sphere_link.data(),
                                      w.arm_sphere_local.data(), (uint32_t)(w.obstacles.size() / 4), w.obstacles.data(),
                                      device, &world_);
      break;
  }
  if (rc != GC_OK)
    raise("CudaCollisionChecker: world creation failed (there is no CPU fallback)");
}

CudaCollisionChecker::CudaCollisionChecker(const cnr_logger::TraceLoggerPtr& logger, gc_world_t* world,
                                           const double& min_distance)
  : CollisionCheckerBase(logger, min_distance), world_(world), dim_(gc_world_dim(world))
{
  if (!world || gc_world_retain(world) != GC_OK)
    raise("CudaCollisionChecker: invalid world handle");
}

CudaCollisionChecker::~CudaCollisionChecker()
{
  if (world_)
    gc_world_destroy(world_);
}

// CollisionCheckerBase::clone (:319): an independent host object; the immutable device world is shared
CollisionCheckerPtr CudaCollisionChecker::clone()
{
  gc_world_t* c = nullptr;
  if (gc_world_clone(world_, &c) != GC_OK)
    raise("gc_world_clone");
  auto out = std::make_shared<CudaCollisionChecker>(logger_, c, min_distance_);  // retains c
  gc_world_destroy(c);                                                           // drop the creation reference
  return out;
}

bool CudaCollisionChecker::check(const Eigen::VectorXd& configuration)
{
  uint8_t ok = 0;
  device_calls_++;
  if (gc_check_states(world_, configuration.data(), 1, &ok) != GC_OK)
    raise("gc_check_states");
  return ok != 0;
}

void CudaCollisionChecker::checkStates(const std::vector<Eigen::VectorXd>& configurations, std::vector<bool>& free)
{
  const uint32_t n = (uint32_t)configurations.size();
  std::vector<double> q((size_t)n * dim_);
  for (uint32_t i = 0; i < n; i++)
    copyConf(configurations[i], &q[(size_t)i * dim_]);
  std::vector<uint8_t> ok(n);
  device_calls_++;
  if (n && gc_check_states(world_, q.data(), n, ok.data()) != GC_OK)
    raise("gc_check_states");
  free.resize(n);
  for (uint32_t i = 0; i < n; i++)
    free[i] = ok[i] != 0;
}

CudaCollisionChecker::Key CudaCollisionChecker::makeKey(const double* a, const double* b) const
{
  Key k;
  k.bits.resize(2 * (size_t)dim_);
  std::memcpy(k.bits.data(), a, sizeof(double) * dim_);
  std::memcpy(k.bits.data() + dim_, b, sizeof(double) * dim_);
  return k;
}

void CudaCollisionChecker::checkConnectionsBatch(const double* c1, const double* c2, uint32_t n, uint8_t* free)
{
  device_calls_++;
  if (n && gc_check_edges(world_, c1, c2, n, min_distance_, GC_EDGE_BISECT, free, nullptr) != GC_OK)
    raise("gc_check_edges");
}

void CudaCollisionChecker::prefetchConnections(const std::vector<const double*>& from,
                                               const std::vector<const double*>& to)
{
  cache_.clear();  // only the latest near set is worth keeping
  const uint32_t n = (uint32_t)from.size();
  if (n == 0)
    return;
  std::vector<double> c1((size_t)n * dim_), c2((size_t)n * dim_);
  for (uint32_t i = 0; i < n; i++)
  {
    std::memcpy(&c1[(size_t)i * dim_], from[i], sizeof(double) * dim_);
    std::memcpy(&c2[(size_t)i * dim_], to[i], sizeof(double) * dim_);
  }
  std::vector<uint8_t> ok(n);
  checkConnectionsBatch(c1.data(), c2.data(), n, ok.data());
  for (uint32_t i = 0; i < n; i++)
    cache_[makeKey(from[i], to[i])] = ok[i];
}

// checkConnection(c1, c2) (:207-237)
bool CudaCollisionChecker::checkConnection(const Eigen::VectorXd& configuration1, const Eigen::VectorXd& configuration2)
{
  if (!cache_.empty())
  {
    auto it = cache_.find(makeKey(configuration1.data(), configuration2.data()));
    if (it != cache_.end())
    {
      cache_hits_++;
      return it->second != 0;
    }
  }
  uint8_t ok = 0;
  checkConnectionsBatch(configuration1.data(), configuration2.data(), 1, &ok);
  return ok != 0;
}

// checkConnection(c1, c2, conf&) (:178-205): conf receives the last feasible configuration
bool CudaCollisionChecker::checkConnection(const Eigen::VectorXd& configuration1,
                                           const Eigen::VectorXd& configuration2, Eigen::VectorXd& conf)
{
  uint8_t ok = 0;
  int64_t first_bad = -1;
  device_calls_++;
  if (gc_check_edges(world_, configuration1.data(), configuration2.data(), 1, min_distance_, GC_EDGE_LINEAR, &ok,
                     &first_bad) != GC_OK)
    raise("gc_check_edges");
  if (ok)
    return true;
  if (first_bad == 0)
    return false;  // configuration1 itself collides: conf is left untouched (:181-182)
  double s = 0.0;
  for (int i = 0; i < dim_; i++)
  {
    const double t = configuration2(i) - configuration1(i);
    s = s + t * t;
  }
  const double dist = std::sqrt(s);
  const unsigned int npnt = (unsigned int)std::ceil(dist / min_distance_);
  if (npnt > 1)
  {
    // interior failure at ipnt = first_bad -> conf = c1 + step*(ipnt-1); failure at configuration2
    // (first_bad == npnt) leaves the last interior point c1 + step*(npnt-1) in conf (:188-202)
    const unsigned int k = (unsigned int)first_bad - 1u;
    conf.resize(dim_);
    for (int i = 0; i < dim_; i++)
    {
      const double step = (configuration2(i) - configuration1(i)) / (double)npnt;
      conf(i) = configuration1(i) + step * (double)k;
    }
  }
  return false;
}

bool CudaCollisionChecker::checkConnection(const ConnectionPtr& conn)
{
  return checkConnection(conn->getParent()->getConfiguration(), conn->getChild()->getConfiguration());
}

// checkConnections (:249-255): true iff every connection is free -- one device batch
bool CudaCollisionChecker::checkConnections(const std::vector<ConnectionPtr>& connections)
{
  const uint32_t n = (uint32_t)connections.size();
  if (n == 0)
    return true;
  std::vector<double> c1((size_t)n * dim_), c2((size_t)n * dim_);
  for (uint32_t i = 0; i < n; i++)
  {
    copyConf(connections[i]->getParent()->getConfiguration(), &c1[(size_t)i * dim_]);
    copyConf(connections[i]->getChild()->getConfiguration(), &c2[(size_t)i * dim_]);
  }
  std::vector<uint8_t> ok(n);
  checkConnectionsBatch(c1.data(), c2.data(), n, ok.data());
  return std::all_of(ok.begin(), ok.end(), [](uint8_t v) { return v != 0; });
}

// checkConnFromConf (:264-313)
bool CudaCollisionChecker::checkConnFromConf(const ConnectionPtr& conn, const Eigen::VectorXd& this_conf)
{
  uint8_t res = 0;
  device_calls_++;
  if (gc_check_edges_from_conf(world_, conn->getParent()->getConfiguration().data(),
                               conn->getChild()->getConfiguration().data(), this_conf.data(), 1, min_distance_,
                               &res) != GC_OK)
    raise("gc_check_edges_from_conf");
  if (res == 255)
  {
    CNR_ERROR(logger_, "The conf is not on the connection between parent and child");
    throw std::invalid_argument("The conf is not on the connection between parent and child");  // :273-277
  }
  return res != 0;
}

// =================================================================================================
//  plugins
// =================================================================================================
namespace
{
template <typename T>
T paramOr(const std::string& ns, const std::string& name, const T& def)
{
  std::string what;
  T v = def;
  if (cnr::param::has(ns + "/" + name, what))
    cnr::param::get(ns + "/" + name, v, what);
  return v;
}
template <typename T>
std::vector<T> castVec(const std::vector<double>& v)
{
  return std::vector<T>(v.begin(), v.end());
}
}  // namespace

bool CudaNearestNeighborsPlugin::init(const std::string& param_ns, const cnr_logger::TraceLoggerPtr& logger)
{
  const int device = paramOr<int>(param_ns, "device_id", 0);
  const int capacity = paramOr<int>(param_ns, "capacity", 4096);
  try
  {
    nearest_neighbors_ = std::make_shared<CudaNearestNeighbors>(logger, device, (uint32_t)capacity);
  }
  catch (const std::exception& e)
  {
    CNR_ERROR(logger, e.what());
    return false;
  }
  return true;
}

bool CudaCollisionCheckerPlugin::init(const std::string& param_ns, const cnr_logger::TraceLoggerPtr& logger,
                                      const double& min_distance)
{
  CudaWorldDescription w;
  const std::string kind = paramOr<std::string>(param_ns, "world_kind", "cube");
  w.dim = paramOr<int>(param_ns, "dimension", 3);
  const int device = paramOr<int>(param_ns, "device_id", 0);
  const std::vector<double> none;
  if (kind == "cube")
  {
    w.kind = GC_WORLD_CUBE;
    w.cube_threshold = paramOr<double>(param_ns, "threshold", 1.0);
  }
  else if (kind == "boxes")
  {
    w.kind = GC_WORLD_BOXES;
    w.boxes_lo = paramOr<std::vector<double>>(param_ns, "boxes_lo", none);
    w.boxes_hi = paramOr<std::vector<double>>(param_ns, "boxes_hi", none);
  }
  else if (kind == "cspheres")
  {
    w.kind = GC_WORLD_CSPHERES;
    w.sphere_centers = castVec<float>(paramOr<std::vector<double>>(param_ns, "centers", none));
    w.sphere_radii = castVec<float>(paramOr<std::vector<double>>(param_ns, "radii", none));
  }
  else if (kind == "sphere_arm")
  {
    w.kind = GC_WORLD_SPHERE_ARM;
    w.arm_base_tf = castVec<float>(paramOr<std::vector<double>>(param_ns, "base_tf", none));
    w.arm_joint_tf = castVec<float>(paramOr<std::vector<double>>(param_ns, "joint_tf", none));
    w.arm_sphere_link = castVec<int32_t>(paramOr<std::vector<double>>(param_ns, "sphere_link", none));
    w.arm_sphere_local = castVec<float>(paramOr<std::vector<double>>(param_ns, "sphere_local", none));
    w.obstacles = castVec<float>(paramOr<std::vector<double>>(param_ns, "obstacles", none));
    if (w.arm_base_tf.size() != 12 || w.arm_joint_tf.size() != 12u * (size_t)w.dim)
    {
      CNR_ERROR(logger, "sphere_arm world: base_tf needs 12 values and joint_tf 12 per joint");
      return false;
    }
  }
  else
  {
    CNR_ERROR(logger, "unknown world_kind " << kind);
    return false;
  }
  try
  {
    collision_checker_ = std::make_shared<CudaCollisionChecker>(logger, w, device, min_distance);
  }
  catch (const std::exception& e)
  {
    CNR_ERROR(logger, e.what());
    return false;
  }
  return true;
}

}  // namespace core
}  // namespace graph

CLASS_LOADER_REGISTER_CLASS(graph::core::CudaNearestNeighborsPlugin, graph::core::NearestNeighborsPlugin)
CLASS_LOADER_REGISTER_CLASS(graph::core::CudaCollisionCheckerPlugin, graph::core::CollisionCheckerBasePlugin)
