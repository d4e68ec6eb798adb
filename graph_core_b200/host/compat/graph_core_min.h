// graph_core_min.h -- the slice of graph_core's public interface that the two drop-in plugins derive
// from, for builds where graph_core itself is not installed.
//
// With -DGCB200_WITH_GRAPH_CORE the real headers are included instead and this file declares nothing:
// the plugin classes then derive from the genuine graph::core::NearestNeighbors /
// CollisionCheckerBase and register with the genuine cnr_class_loader.  Without it
// (this container has no Eigen / cnr_logger / cnr_param / cnr_class_loader) the stand-ins
// below carry the SAME names, signatures and member variables so that the plugin sources compile
// unchanged:
//   graph::core::NearestNeighbors           include/graph_core/datastructure/nearest_neighbors.h:41-195
//   graph::core::CollisionCheckerBase       include/graph_core/collision_checkers/collision_checker_base.h:49-320
//   graph::core::CollisionCheckerBasePlugin include/graph_core/plugins/collision_checkers/collision_checker_base_plugin.h:45-96
//   graph::core::Node / Connection          only getConfiguration(), disconnect(), getParent(), getChild()
// The stand-in base classes hold no algorithm: the reference's CPU loops are not restated here.
#pragma once

#ifdef GCB200_WITH_GRAPH_CORE
#include <cnr_class_loader/register_macro.hpp>
#include <graph_core/collision_checkers/collision_checker_base.h>
#include <graph_core/datastructure/nearest_neighbors.h>
#include <graph_core/plugins/collision_checkers/collision_checker_base_plugin.h>
#else

#include <cmath>
#include <cstddef>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

// ---- Eigen::VectorXd: just the members the plugins touch -----------------------------------------
namespace Eigen
{
class VectorXd
{
public:
  VectorXd() {}
  explicit VectorXd(std::ptrdiff_t n) : v_((size_t)n, 0.0) {}
  VectorXd(std::initializer_list<double> l) : v_(l) {}
  std::ptrdiff_t size() const { return (std::ptrdiff_t)v_.size(); }
  std::ptrdiff_t rows() const { return size(); }
  void resize(std::ptrdiff_t n) { v_.resize((size_t)n); }
  double& operator()(std::ptrdiff_t i) { return v_[(size_t)i]; }
  const double& operator()(std::ptrdiff_t i) const { return v_[(size_t)i]; }
  double& operator[](std::ptrdiff_t i) { return v_[(size_t)i]; }
  const double& operator[](std::ptrdiff_t i) const { return v_[(size_t)i]; }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }

private:
  std::vector<double> v_;
};
}  // namespace Eigen
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

// ---- cnr_logger: a logger that drops everything ---------------------------------------------------
namespace cnr_logger
{
class TraceLogger
{
};
typedef std::shared_ptr<TraceLogger> TraceLoggerPtr;
}  // namespace cnr_logger
#define CNR_TRACE(logger, ...) do { (void)(logger); } while (0)
#define CNR_DEBUG(logger, ...) do { (void)(logger); } while (0)
#define CNR_INFO(logger, ...) do { (void)(logger); } while (0)
#define CNR_WARN(logger, ...) do { (void)(logger); } while (0)
#define CNR_ERROR(logger, ...) do { (void)(logger); } while (0)
#define CNR_FATAL(logger, ...) do { (void)(logger); } while (0)

// ---- cnr_param: an in-process key/value store with the has/get entry points util.h wraps -----------
namespace cnr
{
namespace param
{
struct Value
{
  std::vector<double> num;
  std::string str;
};
inline std::map<std::string, Value>& store()
{
  static std::map<std::string, Value> s;
  return s;
}
inline void set(const std::string& key, double v) { store()[key].num = { v }; }
inline void set(const std::string& key, const std::vector<double>& v) { store()[key].num = v; }
inline void set(const std::string& key, const std::string& v) { store()[key].str = v; }
inline bool has(const std::string& key, std::string& what)
{
  if (store().count(key))
    return true;
  what = "parameter " + key + " not found";
  return false;
}
inline bool get(const std::string& key, double& v, std::string& what)
{
  if (!has(key, what) || store()[key].num.empty())
    return false;
  v = store()[key].num[0];
  return true;
}
inline bool get(const std::string& key, int& v, std::string& what)
{
  double d;
  if (!get(key, d, what))
    return false;
  v = (int)d;
  return true;
}
inline bool get(const std::string& key, std::vector<double>& v, std::string& what)
{
  if (!has(key, what))
    return false;
  v = store()[key].num;
  return true;
}
inline bool get(const std::string& key, std::string& v, std::string& what)
{
  if (!has(key, what))
    return false;
  v = store()[key].str;
  return true;
}
}  // namespace param
}  // namespace cnr

// ---- cnr_class_loader: name -> factory registry behind the same registration macro ------------------
namespace cnr_class_loader
{
inline std::map<std::string, std::function<std::shared_ptr<void>()>>& registry()
{
  static std::map<std::string, std::function<std::shared_ptr<void>()>> r;
  return r;
}
template <typename Base>
std::shared_ptr<Base> createInstance(const std::string& derived_name)
{
  auto it = registry().find(derived_name);
  if (it == registry().end())
    return nullptr;
  return std::static_pointer_cast<Base>(it->second());
}
struct Registrar
{
  Registrar(const char* name, std::function<std::shared_ptr<void>()> f) { registry()[name] = f; }
};
}  // namespace cnr_class_loader
#define GCB200_CAT2(a, b) a##b
#define GCB200_CAT(a, b) GCB200_CAT2(a, b)
#define CLASS_LOADER_REGISTER_CLASS(Derived, Base)                                                        \
  static cnr_class_loader::Registrar GCB200_CAT(gcb200_registrar_, __LINE__)(                             \
      #Derived, [] { return std::static_pointer_cast<void>(std::shared_ptr<Base>(new Derived())); });

namespace graph
{
namespace core
{
class Node;
class Connection;
typedef std::shared_ptr<Node> NodePtr;
typedef std::shared_ptr<Connection> ConnectionPtr;
static const double TOLERANCE = 1e-06;  // include/graph_core/util.h:80

class Node : public std::enable_shared_from_this<Node>
{
public:
  explicit Node(const Eigen::VectorXd& configuration) : configuration_(configuration) {}
  Node(const Eigen::VectorXd& configuration, const cnr_logger::TraceLoggerPtr&) : configuration_(configuration) {}
  const Eigen::VectorXd& getConfiguration() { return configuration_; }
  void disconnect() { disconnected_ = true; }  // the real Node drops its parent/child connections here
  bool disconnected_ = false;

protected:
  Eigen::VectorXd configuration_;
};

class Connection : public std::enable_shared_from_this<Connection>
{
public:
  Connection(const NodePtr& parent, const NodePtr& child, const cnr_logger::TraceLoggerPtr& = nullptr)
    : parent_(parent), child_(child)
  {
  }
  NodePtr getParent() const { return parent_.lock(); }
  NodePtr getChild() const { return child_; }

protected:
  std::weak_ptr<Node> parent_;
  NodePtr child_;
};

class NearestNeighbors : public std::enable_shared_from_this<NearestNeighbors>
{
public:
  NearestNeighbors(const cnr_logger::TraceLoggerPtr& logger) : logger_(logger)
  {
    deleted_nodes_ = 0;
    size_ = 0;
  }
  virtual ~NearestNeighbors() {}
  virtual void insert(const NodePtr& node) = 0;
  virtual bool clear() = 0;
  virtual void nearestNeighbor(const Eigen::VectorXd& configuration, NodePtr& best, double& best_distance) = 0;
  virtual NodePtr nearestNeighbor(const Eigen::VectorXd& configuration)
  {
    double best_distance = std::numeric_limits<double>::infinity();
    NodePtr best;
    nearestNeighbor(configuration, best, best_distance);
    return best;
  }
  virtual std::multimap<double, NodePtr> near(const Eigen::VectorXd& configuration, const double& radius) = 0;
  virtual std::multimap<double, NodePtr> kNearestNeighbors(const Eigen::VectorXd& configuration, const size_t& k) = 0;
  virtual bool findNode(const NodePtr& node) = 0;
  virtual bool deleteNode(const NodePtr& node, const bool& disconnect_node = false) = 0;
  virtual bool restoreNode(const NodePtr& node) = 0;
  virtual unsigned int size() { return size_; }
  virtual std::vector<NodePtr> getNodes() = 0;
  virtual void disconnectNodes(const std::vector<NodePtr>& white_list) = 0;

protected:
  unsigned int size_;
  unsigned int deleted_nodes_;
  cnr_logger::TraceLoggerPtr logger_;
};
typedef std::shared_ptr<NearestNeighbors> NearestNeighborsPtr;

class CollisionCheckerBase;
typedef std::shared_ptr<CollisionCheckerBase> CollisionCheckerPtr;

class CollisionCheckerBase : public std::enable_shared_from_this<CollisionCheckerBase>
{
protected:
  double min_distance_;
  bool verbose_ = false;
  bool initialized_;
  cnr_logger::TraceLoggerPtr logger_;

public:
  CollisionCheckerBase()
  {
    initialized_ = false;
    verbose_ = false;
  }
  CollisionCheckerBase(const cnr_logger::TraceLoggerPtr& logger, const double& min_distance = 0.01)
    : min_distance_(min_distance), logger_(logger)
  {
    initialized_ = true;
    verbose_ = false;
  }
  virtual ~CollisionCheckerBase() {}
  virtual bool init(const cnr_logger::TraceLoggerPtr& logger, const double& min_distance = 0.01)
  {
    if (initialized_)
      return false;
    logger_ = logger;
    min_distance_ = min_distance;
    initialized_ = true;
    return true;
  }
  void setVerbose(const bool& verbose) { verbose_ = verbose; }
  virtual std::string getGroupName() { return ""; }
  bool getInitialized() { return initialized_; }
  virtual bool check(const Eigen::VectorXd& configuration) = 0;
  // In graph_core these five carry the CPU discretisation loops; the CUDA checker overrides all of them.
  virtual bool checkConnection(const Eigen::VectorXd& configuration1, const Eigen::VectorXd& configuration2,
                               Eigen::VectorXd& conf) = 0;
  virtual bool checkConnection(const Eigen::VectorXd& configuration1, const Eigen::VectorXd& configuration2) = 0;
  virtual bool checkConnection(const ConnectionPtr& conn) = 0;
  virtual bool checkConnections(const std::vector<ConnectionPtr>& connections) = 0;
  virtual bool checkConnFromConf(const ConnectionPtr& conn, const Eigen::VectorXd& this_conf) = 0;
  virtual CollisionCheckerPtr clone() = 0;
};

class CollisionCheckerBasePlugin;
typedef std::shared_ptr<CollisionCheckerBasePlugin> CollisionCheckerPluginPtr;
class CollisionCheckerBasePlugin : std::enable_shared_from_this<CollisionCheckerBasePlugin>
{
protected:
  graph::core::CollisionCheckerPtr collision_checker_;

public:
  CollisionCheckerBasePlugin() { collision_checker_ = nullptr; }
  virtual ~CollisionCheckerBasePlugin() { collision_checker_ = nullptr; }
  graph::core::CollisionCheckerPtr getCollisionChecker() { return collision_checker_; }
  virtual bool init(const std::string& param_ns, const cnr_logger::TraceLoggerPtr& logger,
                    const double& min_distance = 0.01) = 0;
};

}  // namespace core
}  // namespace graph
#endif  // GCB200_WITH_GRAPH_CORE
