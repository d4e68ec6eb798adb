// cuda_plugins.h -- the two drop-in operator classes and their cnr_class_loader plugins.
//
//   graph::core::CudaNearestNeighbors : graph::core::NearestNeighbors
//       overrides every pure virtual of include/graph_core/datastructure/nearest_neighbors.h:59-173;
//       semantic twin of graph::core::Vector (src/graph_core/datastructure/vector.cpp:38-138: scan order =
//       insertion order, strict '<', ties to the earliest node) with KdTree's lazy deletion
//       (src/graph_core/datastructure/kdtree.cpp:261-278, 395-428).
//   graph::core::CudaCollisionChecker : graph::core::CollisionCheckerBase
//       overrides check (:166), clone (:319) and -- for speed -- the already-virtual checkConnection x3
//       (:178, :207, :239), checkConnections (:249) and checkConnFromConf (:264).
//   graph::core::CudaCollisionCheckerPlugin : graph::core::CollisionCheckerBasePlugin
//       (include/graph_core/plugins/collision_checkers/collision_checker_base_plugin.h:45-96)
//   graph::core::NearestNeighborsPlugin / CudaNearestNeighborsPlugin
//       graph_core has no plugin base for NearestNeighbors; this one follows the
//       collision-checker plugin pattern and is what INTEGRATION.md's Tree factory hook loads.
//
// All device work goes through the C ABI of include/gcb200.h.  There is no CPU fallback: construction
// throws std::runtime_error when no CUDA device is usable.
#pragma once
#include <cstdint>
#include <memory>
#include <unordered_map>
#include <vector>

#include "compat/graph_core_min.h"
#include "gcb200.h"

namespace graph
{
namespace core
{
class CudaCollisionChecker;
typedef std::shared_ptr<CudaCollisionChecker> CudaCollisionCheckerPtr;

class CudaNearestNeighbors : public NearestNeighbors
{
public:
  // capacity grows on demand (the store is re-created at twice the size and refilled on the device)
  CudaNearestNeighbors(const cnr_logger::TraceLoggerPtr& logger, int device = 0, uint32_t capacity = 4096);
  ~CudaNearestNeighbors();  // the upstream base has no virtual destructor: objects live in shared_ptr

  void insert(const NodePtr& node) override;
  bool clear() override;
  void nearestNeighbor(const Eigen::VectorXd& configuration, NodePtr& best, double& best_distance) override;
  using NearestNeighbors::nearestNeighbor;
  std::multimap<double, NodePtr> near(const Eigen::VectorXd& configuration, const double& radius) override;
  std::multimap<double, NodePtr> kNearestNeighbors(const Eigen::VectorXd& configuration, const size_t& k) override;
  bool findNode(const NodePtr& node) override;
  bool deleteNode(const NodePtr& node, const bool& disconnect_node = false) override;
  bool restoreNode(const NodePtr& node) override;
  std::vector<NodePtr> getNodes() override;
  void disconnectNodes(const std::vector<NodePtr>& white_list) override;

  // batched entry point for callers that have many queries (not part of the reference interface)
  void nearestNeighbors(const std::vector<Eigen::VectorXd>& configurations, std::vector<NodePtr>& best,
                        std::vector<double>& best_distance);
  // Zero-patch batching of Tree::rewireOnly: when set, every near()/kNearestNeighbors()
  // result is handed to the checker, which evaluates both ordered edges (hit -> q, q -> hit) in one
  // batch and answers the reference's following one-at-a-time checkConnection calls from that cache.
  void setSpeculativeChecker(const CudaCollisionCheckerPtr& checker) { spec_ = checker; }
  unsigned int deletedNodes() const { return deleted_nodes_; }
  gc_store_t* store() { return store_; }

private:
  void ensureStore(int dim);
  void grow();
  std::multimap<double, NodePtr> toMap(uint32_t count, const uint32_t* idx, const double* dist);
  void speculate(const Eigen::VectorXd& q, const std::multimap<double, NodePtr>& hits);

  int device_;
  uint32_t capacity_;
  int dim_ = 0;
  gc_store_t* store_ = nullptr;
  std::vector<NodePtr> node_of_slot_;
  std::vector<uint8_t> dead_;
  std::unordered_map<Node*, uint32_t> slot_of_;
  std::vector<uint32_t> idx_;
  std::vector<double> dist_;
  CudaCollisionCheckerPtr spec_;
};
typedef std::shared_ptr<CudaNearestNeighbors> CudaNearestNeighborsPtr;

// description of a collision world, the argument of CudaCollisionChecker's constructor
struct CudaWorldDescription
{
  int kind = GC_WORLD_CUBE;  // gc_world_kind
  int dim = 3;
  double cube_threshold = 1.0;
  std::vector<double> boxes_lo, boxes_hi;   // [n][dim]
  std::vector<float> sphere_centers;        // [n][dim]
  std::vector<float> sphere_radii;          // [n]
  std::vector<float> arm_base_tf;           // 12
  std::vector<float> arm_joint_tf;          // [dim][12]
  std::vector<int32_t> arm_sphere_link;     // [nrs]
  std::vector<float> arm_sphere_local;      // [nrs][4]
  std::vector<float> obstacles;             // [nobs][4]
};

class CudaCollisionChecker : public CollisionCheckerBase
{
public:
  CudaCollisionChecker(const cnr_logger::TraceLoggerPtr& logger, const CudaWorldDescription& world, int device = 0,
                       const double& min_distance = 0.01);
  // shares an existing device world (one more reference is taken)
  CudaCollisionChecker(const cnr_logger::TraceLoggerPtr& logger, gc_world_t* world, const double& min_distance = 0.01);
  ~CudaCollisionChecker();  // as above

  bool check(const Eigen::VectorXd& configuration) override;
  bool checkConnection(const Eigen::VectorXd& configuration1, const Eigen::VectorXd& configuration2,
                       Eigen::VectorXd& conf) override;
  bool checkConnection(const Eigen::VectorXd& configuration1, const Eigen::VectorXd& configuration2) override;
  bool checkConnection(const ConnectionPtr& conn) override;
  bool checkConnections(const std::vector<ConnectionPtr>& connections) override;
  bool checkConnFromConf(const ConnectionPtr& conn, const Eigen::VectorXd& this_conf) override;
  CollisionCheckerPtr clone() override;
  std::string getGroupName() override { return "cuda_world"; }

  // batched entry points (not part of the reference interface)
  void checkStates(const std::vector<Eigen::VectorXd>& configurations, std::vector<bool>& free);
  void checkConnectionsBatch(const double* c1, const double* c2, uint32_t n, uint8_t* free);
  // evaluates a -> b for every listed pair in one device batch and remembers the answers for the
  // checkConnection(c1, c2) calls that follow (keyed on the exact bit patterns of the ordered pair)
  void prefetchConnections(const std::vector<const double*>& from, const std::vector<const double*>& to);
  uint64_t cacheHits() const { return cache_hits_; }
  uint64_t deviceCalls() const { return device_calls_; }
  gc_world_t* world() { return world_; }
  int dim() const { return dim_; }

private:
  struct Key
  {
    std::vector<uint64_t> bits;
    bool operator==(const Key& o) const { return bits == o.bits; }
  };
  struct KeyHash
  {
    size_t operator()(const Key& k) const
    {
      uint64_t h = 1469598103934665603ull;
      for (uint64_t b : k.bits)
        h = (h ^ b) * 1099511628211ull;
      return (size_t)h;
    }
  };
  Key makeKey(const double* a, const double* b) const;

  gc_world_t* world_ = nullptr;
  int dim_ = 0;
  std::unordered_map<Key, uint8_t, KeyHash> cache_;
  uint64_t cache_hits_ = 0, device_calls_ = 0;
};

// ---- plugins ---------------------------------------------------------------------------------------
class NearestNeighborsPlugin;
typedef std::shared_ptr<NearestNeighborsPlugin> NearestNeighborsPluginPtr;
class NearestNeighborsPlugin : std::enable_shared_from_this<NearestNeighborsPlugin>
{
protected:
  NearestNeighborsPtr nearest_neighbors_;

public:
  NearestNeighborsPlugin() { nearest_neighbors_ = nullptr; }
  virtual ~NearestNeighborsPlugin() { nearest_neighbors_ = nullptr; }
  NearestNeighborsPtr getNearestNeighbors() { return nearest_neighbors_; }
  virtual bool init(const std::string& param_ns, const cnr_logger::TraceLoggerPtr& logger) = 0;
};

// parameters (cnr_param, namespace param_ns): device_id (0), capacity (4096)
class CudaNearestNeighborsPlugin : public NearestNeighborsPlugin
{
public:
  bool init(const std::string& param_ns, const cnr_logger::TraceLoggerPtr& logger) override;
};

// parameters: device_id (0), world_kind ("cube" | "boxes" | "cspheres" | "sphere_arm"), dimension,
// threshold | boxes_lo, boxes_hi | centers, radii | base_tf, joint_tf, sphere_link, sphere_local, obstacles
class CudaCollisionCheckerPlugin : public CollisionCheckerBasePlugin
{
public:
  bool init(const std::string& param_ns, const cnr_logger::TraceLoggerPtr& logger,
            const double& min_distance = 0.01) override;
};

}  // namespace core
}  // namespace graph
