/*
 * gc_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain C++ restatement of the two graph_core hot paths and of the callers
 * that drive them.  It exists so that the CUDA path can be checked bit for bit
 * and so that bench.py has a host-CPU baseline ("port").  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; nothing under graph_core_b200/ links or calls it.
 *
 * PARITY PINNING: the reference's own tests hold no numerical vector for either
 * path.  The restatement is pinned against the REFERENCE ITSELF:
 * oracle/ref_build.py compiles graph_core's own translation units (kdtree.cpp,
 * vector.cpp, node.cpp, connection.cpp, tree.cpp, path.cpp, rrt.cpp, rrt_star.cpp,
 * samplers, collision_checker_base.h ...) where they lie under /root/reference
 * against stand-in Eigen / cnr_logger / cnr_param / yaml-cpp headers
 * (oracle/ref_shim/) into oracle/_ref/libgc_ref.so (driver: oracle/ref_driver.cpp).
 *   tests/test_oracle_vs_ref.py      restatement == reference, bit for bit: NN ids,
 *                                    orderings and distances (Vector and KdTree, ties,
 *                                    tombstones, rebuild), all three checkConnection
 *                                    forms in every world, complete RRT / RRT* trees
 *   tests/golden/ref_golden.json     outputs of the reference committed as fixtures
 *                                    (tests/golden/make_ref_golden.py), replayed by
 *                                    tests/test_oracle_cpu.py and, through the CUDA
 *                                    path, by tests/test_golden_gpu.py
 * plus the single fixture of the reference's tests, tests/path.yaml.
 *
 * Reference citations are relative to /root/reference/graph_core/ :
 *   inc/ = include/graph_core/,  src/ = src/graph_core/
 *
 * Arithmetic conventions (documented in DESIGN.md, mirrored by the CUDA code):
 *   - Euclidean norm = sqrt(sum_i (a_i-b_i)^2), i ascending, every operation
 *     individually rounded in fp64 (build with -ffp-contract=off).
 *   - synthetic worlds (not in the reference) follow include/gcb200_spec.h.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <vector>

#include "../include/gcb200_spec.h"

namespace orc
{
static const double kInf = std::numeric_limits<double>::infinity();
static const double TOLERANCE = 1e-06;  // inc/util.h:80

// Eigen (a-b).norm() restated: sequential sum of squares, then sqrt.
static inline double dist(const double* a, const double* b, int d)
{
  double s = 0.0;
  for (int i = 0; i < d; i++)
  {
    double t = a[i] - b[i];
    s = s + t * t;
  }
  return std::sqrt(s);
}

// ============================================================================
//  Node set: ids are creation order; configurations immutable (src/graph/node.cpp:34-48)
// ============================================================================
struct NodePool
{
  int dim;
  std::vector<double> q;  // [n][dim]
  explicit NodePool(int d) : dim(d) {}
  int add(const double* c)
  {
    q.insert(q.end(), c, c + dim);
    return (int)(q.size() / dim) - 1;
  }
  const double* conf(int id) const { return &q[(size_t)id * dim]; }
  int size() const { return (int)(q.size() / dim); }
};

typedef std::multimap<double, int> NearMap;  // multimap<double,NodePtr> of the reference

// ============================================================================
//  NearestNeighbors contract (inc/datastructure/nearest_neighbors.h:41-195)
// ============================================================================
struct NearestNeighbors
{
  NodePool* pool;
  unsigned int size_ = 0;
  unsigned int deleted_nodes_ = 0;
  explicit NearestNeighbors(NodePool* p) : pool(p) {}
  virtual ~NearestNeighbors() {}
  virtual void insert(int node) = 0;
  virtual bool clear() = 0;
  virtual void nearestNeighbor(const double* q, int& best, double& best_distance) = 0;
  virtual NearMap near(const double* q, double radius) = 0;
  virtual NearMap kNearestNeighbors(const double* q, size_t k) = 0;
  virtual bool findNode(int node) = 0;
  virtual bool deleteNode(int node) = 0;
  virtual bool restoreNode(int node) = 0;
  virtual std::vector<int> getNodes() = 0;
};

// ---- Vector: brute force (src/datastructure/vector.cpp:38-138) ---------------
struct VectorNN : NearestNeighbors
{
  std::vector<int> nodes_;
  explicit VectorNN(NodePool* p) : NearestNeighbors(p) {}
  void insert(int node) override  // vector.cpp:38-43
  {
    nodes_.push_back(node);
    size_++;
  }
  bool clear() override  // :45-53
  {
    size_ = 0;
    deleted_nodes_ = 0;
    nodes_.clear();
    return true;
  }
  void nearestNeighbor(const double* q, int& best, double& best_distance) override  // :55-67
  {
    best_distance = kInf;
    for (int n : nodes_)
    {
      double d = dist(pool->conf(n), q, pool->dim);
      if (d < best_distance)
      {
        best = n;
        best_distance = d;
      }
    }
  }
  NearMap near(const double* q, double radius) override  // :69-81
  {
    NearMap out;
    for (int n : nodes_)
    {
      double d = dist(pool->conf(n), q, pool->dim);
      if (d < radius)
        out.insert(std::make_pair(d, n));
    }
    return out;
  }
  NearMap kNearestNeighbors(const double* q, size_t k) override  // :83-96
  {
    NearMap all;
    for (int n : nodes_)
      all.insert(std::make_pair(dist(pool->conf(n), q, pool->dim), n));
    if (all.size() < k)
      return all;
    return NearMap(all.begin(), std::next(all.begin(), k));
  }
  bool findNode(int node) override { return std::find(nodes_.begin(), nodes_.end(), node) != nodes_.end(); }
  bool deleteNode(int node) override  // :103-119 (physical erase)
  {
    auto it = std::find(nodes_.begin(), nodes_.end(), node);
    if (it == nodes_.end())
      return false;
    size_--;
    deleted_nodes_++;
    nodes_.erase(it);
    return true;
  }
  bool restoreNode(int) override { return false; }  // :121-124
  std::vector<int> getNodes() override { return nodes_; }
};

// ---- KdTree: unbalanced incremental k-d tree (src/datastructure/kdtree.cpp:91-472) ----
struct KdNode
{
  int node;
  int dimension;
  bool deleted = false;
  int left = -1, right = -1;
};

struct KdTreeNN : NearestNeighbors
{
  std::vector<KdNode> kd;  // kd[0] is the root when non-empty
  unsigned int deleted_nodes_threshold_ = std::numeric_limits<unsigned int>::max();  // kdtree.cpp:319
  explicit KdTreeNN(NodePool* p) : NearestNeighbors(p) {}

  // KdNode::insert (kdtree.cpp:91-125); returns false when an existing node is restored
  bool kdInsert(int cur, int node)
  {
    for (;;)
    {
      if (node == kd[cur].node)
      {
        kd[cur].deleted = false;
        return false;
      }
      int dimn = kd[cur].dimension;
      int next_dim = (dimn == pool->dim - 1) ? 0 : dimn + 1;
      if (pool->conf(node)[dimn] >= pool->conf(kd[cur].node)[dimn])  // goRight
      {
        if (kd[cur].right < 0)
        {
          KdNode nn;
          nn.node = node;
          nn.dimension = next_dim;
          kd.push_back(nn);
          kd[cur].right = (int)kd.size() - 1;
          return true;
        }
        cur = kd[cur].right;
      }
      else
      {
        if (kd[cur].left < 0)
        {
          KdNode nn;
          nn.node = node;
          nn.dimension = next_dim;
          kd.push_back(nn);
          kd[cur].left = (int)kd.size() - 1;
          return true;
        }
        cur = kd[cur].left;
      }
    }
  }
  void insert(int node) override  // kdtree.cpp:327-342
  {
    size_++;
    if (kd.empty())
    {
      KdNode r;
      r.node = node;
      r.dimension = 0;
      kd.push_back(r);
      return;
    }
    if (!kdInsert(0, node))
      deleted_nodes_--;
  }
  bool clear() override  // :344-352
  {
    size_ = 0;
    deleted_nodes_ = 0;
    kd.clear();
    return true;
  }
  // KdNode::nearestNeighbor (:127-161)
  void nnRec(int cur, const double* q, int& best, double& best_distance)
  {
    const KdNode& k = kd[cur];
    const double* c = pool->conf(k.node);
    double distance = dist(q, c, pool->dim);
    if (!k.deleted && distance < best_distance)
    {
      best_distance = distance;
      best = k.node;
    }
    int dm = k.dimension;
    int l = k.left, r = k.right;
    bool right_first = q[dm] > c[dm];
    if (!right_first)
    {
      if (l >= 0 && (q[dm] - best_distance) <= c[dm])
        nnRec(l, q, best, best_distance);
      if (r >= 0 && (q[dm] + best_distance) >= c[dm])
        nnRec(r, q, best, best_distance);
    }
    else
    {
      if (r >= 0 && (q[dm] + best_distance) >= c[dm])
        nnRec(r, q, best, best_distance);
      if (l >= 0 && (q[dm] - best_distance) <= c[dm])
        nnRec(l, q, best, best_distance);
    }
  }
  void nearestNeighbor(const double* q, int& best, double& best_distance) override  // :354-360
  {
    best_distance = kInf;
    if (kd.empty())
      return;
    nnRec(0, q, best, best_distance);
  }
  // KdNode::near (:163-180)
  void nearRec(int cur, const double* q, double radius, NearMap& out)
  {
    const KdNode& k = kd[cur];
    const double* c = pool->conf(k.node);
    double distance = dist(q, c, pool->dim);
    if (!k.deleted && distance < radius)
      out.insert(std::make_pair(distance, k.node));
    int dm = k.dimension;
    int l = k.left, r = k.right;
    if (l >= 0 && (q[dm] - radius) <= c[dm])
      nearRec(l, q, radius, out);
    if (r >= 0 && (q[dm] + radius) >= c[dm])
      nearRec(r, q, radius, out);
  }
  NearMap near(const double* q, double radius) override  // :362-370
  {
    NearMap out;
    if (kd.empty())
      return out;
    nearRec(0, q, radius, out);
    return out;
  }
  // KdNode::kNearestNeighbors (:182-229)
  void knnRec(int cur, const double* q, size_t k, NearMap& nodes)
  {
    const KdNode& kn = kd[cur];
    const double* c = pool->conf(kn.node);
    double last_distance = nodes.empty() ? kInf : std::prev(nodes.end())->first;
    double distance = dist(q, c, pool->dim);
    if (!kn.deleted && nodes.size() < k)
    {
      nodes.insert(std::make_pair(distance, kn.node));
      last_distance = std::prev(nodes.end())->first;
    }
    else if (!kn.deleted && distance < last_distance)
    {
      nodes.erase(std::prev(nodes.end()));
      nodes.insert(std::make_pair(distance, kn.node));
      last_distance = std::prev(nodes.end())->first;
    }
    int dm = kn.dimension;
    int l = kn.left, r = kn.right;
    bool right_first = q[dm] > c[dm];
    if (!right_first)
    {
      if (l >= 0 && (q[dm] - last_distance) <= c[dm])
        knnRec(l, q, k, nodes);
      if (r >= 0 && (q[dm] + last_distance) >= c[dm])
        knnRec(r, q, k, nodes);
    }
    else
    {
      if (r >= 0 && (q[dm] + last_distance) >= c[dm])
        knnRec(r, q, k, nodes);
      if (l >= 0 && (q[dm] - last_distance) <= c[dm])
        knnRec(l, q, k, nodes);
    }
  }
  NearMap kNearestNeighbors(const double* q, size_t k) override  // :372-380
  {
    NearMap out;
    if (kd.empty())
      return out;
    knnRec(0, q, k, out);
    return out;
  }
  // KdNode::findNode (:231-259): descends by coordinates, fails on tombstones
  int findKd(int node)
  {
    if (kd.empty())
      return -1;
    int cur = 0;
    for (;;)
    {
      if (kd[cur].node == node)
        return kd[cur].deleted ? -1 : cur;
      int dm = kd[cur].dimension;
      if (pool->conf(node)[dm] >= pool->conf(kd[cur].node)[dm])
      {
        if (kd[cur].right < 0)
          return -1;
        cur = kd[cur].right;
      }
      else
      {
        if (kd[cur].left < 0)
          return -1;
        cur = kd[cur].left;
      }
    }
  }
  bool findNode(int node) override { return findKd(node) >= 0; }
  void getNodesRec(int cur, std::vector<int>& out)  // :280-288 (preorder, skips tombstones)
  {
    if (!kd[cur].deleted)
      out.push_back(kd[cur].node);
    if (kd[cur].left >= 0)
      getNodesRec(kd[cur].left, out);
    if (kd[cur].right >= 0)
      getNodesRec(kd[cur].right, out);
  }
  std::vector<int> getNodes() override
  {
    std::vector<int> out;
    if (!kd.empty())
      getNodesRec(0, out);
    return out;
  }
  bool deleteNode(int node) override  // :395-428
  {
    int k = findKd(node);
    if (k < 0)
      return false;
    if (!kd[k].deleted)
    {
      kd[k].deleted = true;
      size_--;
      deleted_nodes_++;
      if (deleted_nodes_ > deleted_nodes_threshold_ && size_ > 0)
      {
        bool root_was_deleted = kd[0].deleted;
        int root_node = kd[0].node;
        kd[0].deleted = false;
        std::vector<int> nodes = getNodes();
        clear();
        for (int n : nodes)
          insert(n);
        if (root_was_deleted)
          deleteNode(root_node);
      }
    }
    return true;
  }
  bool restoreNode(int node) override  // :447-456 (findNode fails on tombstones => false for deleted nodes)
  {
    int k = findKd(node);
    if (k < 0)
      return false;
    size_++;
    deleted_nodes_--;
    kd[k].deleted = false;
    return true;
  }
};

// ============================================================================
//  Collision worlds: check(q) -> true = collision free (collision_checker_base.h:166)
// ============================================================================
struct World
{
  int kind = 0;
  int dim = 0;
  // cube
  double cube_thr = 1.0;
  // boxes: lo/hi [n][dim]
  int nobj = 0;
  std::vector<double> lo, hi;
  // cspheres: centres fp32 [n][dim], radii^2 fp32
  std::vector<float> sc, sr2;
  // sphere arm
  int njoints = 0, nrs = 0, nobs = 0;
  std::vector<float> base_tf;   // 12: R row-major (9) + t (3)
  std::vector<float> joint_tf;  // [J][12]
  std::vector<int> rs_link;     // [nrs] in [0, J]
  std::vector<float> rs_local;  // [nrs][4] x y z r
  std::vector<float> obs;       // [nobs][4]
  std::vector<float> obs_na;    // [nobs] ro^2 - |o|^2 (fp32, spec order)
  unsigned long long pair_tests = 0;  // executed sphere-pair tests (for flop accounting)
  unsigned long long state_checks = 0;
  // Optional uniform-grid culling (orc_world_enable_grid): the same idea the CUDA path uses, restated here so that the
  // CPU arm of the benchmark is not denied the algorithm -- cell -> obstacles closer than rs_max + r_o + m to the cell
  // box (m = E/256, see DESIGN.md "grid culling"); every other pair is provably "apart" for the spec'd fp32 test, so
  // the verdict is the one of the plain loop above (tests/test_oracle_cpu.py checks that on random and grazing states).
  bool grid_on = false;
  std::vector<unsigned> grid_start;
  std::vector<unsigned short> grid_items;
  float grid_lo[3] = { 0, 0, 0 };
  float grid_inv_h = 0;
  int grid_n[3] = { 0, 0, 0 };
};

static inline void sincos_spec(float q, float& s, float& c)
{
  float j = rintf(q * GC_TWO_OVER_PI);
  float r = __builtin_fmaf(-j, GC_PIO2_1, q);
  r = __builtin_fmaf(-j, GC_PIO2_2, r);
  r = __builtin_fmaf(-j, GC_PIO2_3, r);
  float z = r * r;
  float ps = __builtin_fmaf(z, GC_SIN_S3, GC_SIN_S2);
  ps = __builtin_fmaf(z, ps, GC_SIN_S1);
  float rz = r * z;
  float sn = __builtin_fmaf(rz, ps, r);
  float pc = __builtin_fmaf(z, GC_COS_C3, GC_COS_C2);
  pc = __builtin_fmaf(z, pc, GC_COS_C1);
  float zz = z * z;
  float cs = __builtin_fmaf(z, -0.5f, 1.0f);
  cs = __builtin_fmaf(zz, pc, cs);
  int qd = ((int)j) & 3;
  switch (qd)
  {
    case 0: s = sn; c = cs; break;
    case 1: s = cs; c = -sn; break;
    case 2: s = -sn; c = -cs; break;
    default: s = -cs; c = sn; break;
  }
}

// 3x3 * 3x3 with the spec's operation order: out[r][c] = fma(a[r][2],b[2][c], fma(a[r][1],b[1][c], a[r][0]*b[0][c]))
static inline void mat3_mul(const float* a, const float* b, float* o)
{
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++)
    {
      float v = a[3 * r + 0] * b[0 + c];
      v = __builtin_fmaf(a[3 * r + 1], b[3 + c], v);
      v = __builtin_fmaf(a[3 * r + 2], b[6 + c], v);
      o[3 * r + c] = v;
    }
}
// o = R*p + t
static inline void mat3_vec(const float* R, const float* t, const float* p, float* o)
{
  for (int r = 0; r < 3; r++)
  {
    float v = R[3 * r + 0] * p[0];
    v = __builtin_fmaf(R[3 * r + 1], p[1], v);
    v = __builtin_fmaf(R[3 * r + 2], p[2], v);
    o[r] = v + t[r];
  }
}

static bool check_sphere_arm(World& w, const double* q)
{
  const int J = w.njoints;
  // link frames B_j (rotation 9 + translation 3), frame J = tool frame A_J
  float fr[(GC_MAX_JOINTS + 1) * 12];
  float AR[9], At[3];
  for (int i = 0; i < 9; i++) AR[i] = w.base_tf[i];
  for (int i = 0; i < 3; i++) At[i] = w.base_tf[9 + i];
  for (int j = 0; j < J; j++)
  {
    float s, c;
    sincos_spec((float)q[j], s, c);
    // B = A * Rz(q): col0' = c*col0 + s*col1 ; col1' = c*col1 - s*col0 ; col2' = col2
    float BR[9];
    for (int r = 0; r < 3; r++)
    {
      float x = AR[3 * r + 0], y = AR[3 * r + 1];
      BR[3 * r + 0] = __builtin_fmaf(s, y, c * x);
      BR[3 * r + 1] = __builtin_fmaf(-s, x, c * y);
      BR[3 * r + 2] = AR[3 * r + 2];
    }
    float* f = &fr[j * 12];
    for (int i = 0; i < 9; i++) f[i] = BR[i];
    for (int i = 0; i < 3; i++) f[9 + i] = At[i];
    // A_{j+1} = B * F_j
    const float* F = &w.joint_tf[j * 12];
    float NR[9], Nt[3];
    mat3_mul(BR, F, NR);
    mat3_vec(BR, At, F + 9, Nt);
    for (int i = 0; i < 9; i++) AR[i] = NR[i];
    for (int i = 0; i < 3; i++) At[i] = Nt[i];
  }
  {
    float* f = &fr[J * 12];
    for (int i = 0; i < 9; i++) f[i] = AR[i];
    for (int i = 0; i < 3; i++) f[9 + i] = At[i];
  }
  bool collide = false;
  for (int k = 0; k < w.nrs && !collide; k++)
  {
    const float* f = &fr[w.rs_link[k] * 12];
    float c[3];
    mat3_vec(f, f + 9, &w.rs_local[4 * k], c);
    float rs = w.rs_local[4 * k + 3];
    // sphere pair test of include/gcb200_spec.h: |o-c|^2 < (ro+rs)^2 expanded to
    //   B + ox*(-2cx) + oy*(-2cy) + oz*(-2cz) + ro*(-2rs) < ro^2 - |o|^2 ,  B = |c|^2 - rs^2
    const float c2x = -2.0f * c[0], c2y = -2.0f * c[1], c2z = -2.0f * c[2], rs2 = -2.0f * rs;
    float n2 = c[0] * c[0];
    n2 = __builtin_fmaf(c[1], c[1], n2);
    n2 = __builtin_fmaf(c[2], c[2], n2);
    const float B = n2 - rs * rs;
    if (w.grid_on)
    {
      const float fx = (c[0] - w.grid_lo[0]) * w.grid_inv_h;
      const float fy = (c[1] - w.grid_lo[1]) * w.grid_inv_h;
      const float fz = (c[2] - w.grid_lo[2]) * w.grid_inv_h;
      if (!(fx >= 0.0f && fy >= 0.0f && fz >= 0.0f && fx < (float)w.grid_n[0] && fy < (float)w.grid_n[1] &&
            fz < (float)w.grid_n[2]))
        continue;  // no obstacle within reach of this sphere
      const unsigned cell = ((unsigned)fx * (unsigned)w.grid_n[1] + (unsigned)fy) * (unsigned)w.grid_n[2] + (unsigned)fz;
      for (unsigned i = w.grid_start[cell]; i < w.grid_start[cell + 1]; i++)
      {
        const int o = w.grid_items[i];
        const float* ob = &w.obs[4 * o];
        const float na = w.obs_na[o];
        float t = __builtin_fmaf(ob[0], c2x, B);
        t = __builtin_fmaf(ob[1], c2y, t);
        t = __builtin_fmaf(ob[2], c2z, t);
        t = __builtin_fmaf(ob[3], rs2, t);
        w.pair_tests++;
        if (t < na)
        {
          collide = true;
          break;
        }
      }
      continue;
    }
    for (int o = 0; o < w.nobs; o++)
    {
      const float* ob = &w.obs[4 * o];
      const float na = w.obs_na[o];
      float t = __builtin_fmaf(ob[0], c2x, B);
      t = __builtin_fmaf(ob[1], c2y, t);
      t = __builtin_fmaf(ob[2], c2z, t);
      t = __builtin_fmaf(ob[3], rs2, t);
      w.pair_tests++;
      if (t < na)
      {
        collide = true;
        break;
      }
    }
  }
  return !collide;
}

static bool world_check(World& w, const double* q)
{
  w.state_checks++;
  switch (w.kind)
  {
    case GC_WORLD_CUBE:  // inc/collision_checkers/cube_3d_collision_checker.h:61-64
    {
      double m = 0.0;
      for (int i = 0; i < w.dim; i++)
        m = std::max(m, std::fabs(q[i]));
      return m > w.cube_thr;
    }
    case GC_WORLD_BOXES:  // collide iff inside any box (closed), pure fp64 comparisons
    {
      for (int b = 0; b < w.nobj; b++)
      {
        bool inside = true;
        for (int i = 0; i < w.dim; i++)
        {
          double v = q[i];
          if (v < w.lo[(size_t)b * w.dim + i] || v > w.hi[(size_t)b * w.dim + i])
          {
            inside = false;
            break;
          }
        }
        if (inside)
          return false;
      }
      return true;
    }
    case GC_WORLD_CSPHERES:  // collide iff sum (float(q_i)-c_i)^2 < r^2 in fp32 (fma chain, i ascending)
    {
      float qf[GC_MAX_DIM];
      for (int i = 0; i < w.dim; i++)
        qf[i] = (float)q[i];
      for (int s = 0; s < w.nobj; s++)
      {
        const float* c = &w.sc[(size_t)s * w.dim];
        float t0 = qf[0] - c[0];
        float acc = t0 * t0;
        for (int i = 1; i < w.dim; i++)
        {
          float t = qf[i] - c[i];
          acc = __builtin_fmaf(t, t, acc);
        }
        w.pair_tests++;
        if (acc < w.sr2[s])
          return false;
      }
      return true;
    }
    case GC_WORLD_SPHERE_ARM:
      return check_sphere_arm(w, q);
  }
  return false;
}

// ============================================================================
//  CollisionCheckerBase discretisation logic (inc/collision_checkers/collision_checker_base.h)
// ============================================================================
// checkConnection(c1,c2)  :207-237 -- bisection order
static bool checkConnectionBisect(World& w, const double* c1, const double* c2, double min_distance)
{
  const int d = w.dim;
  if (!world_check(w, c1))
    return false;
  if (!world_check(w, c2))
    return false;
  double distance = dist(c2, c1, d);
  if (distance < min_distance)
    return true;
  double conf[GC_MAX_DIM];
  double n = 2;
  while (distance > n * min_distance)
  {
    for (double idx = 1; idx < n; idx += 2)
    {
      for (int i = 0; i < d; i++)
        conf[i] = c1[i] + ((c2[i] - c1[i]) * idx) / n;
      if (!world_check(w, conf))
        return false;
    }
    n *= 2;
  }
  return true;
}

// checkConnection(c1,c2,conf&) :178-205 -- linear order; returns first failing state index via *first_bad
// (0 = c1, 1..npnt-1 interior, npnt = c2, -1 = none); conf receives the reference's "last feasible" vector.
static bool checkConnectionLinear(World& w, const double* c1, const double* c2, double min_distance, double* conf,
                                  long long* first_bad)
{
  const int d = w.dim;
  *first_bad = -1;
  if (!world_check(w, c1))
  {
    *first_bad = 0;
    return false;
  }
  double dst = dist(c2, c1, d);
  unsigned int npnt = (unsigned int)std::ceil(dst / min_distance);
  if (npnt > 1)
  {
    double step[GC_MAX_DIM];
    for (int i = 0; i < d; i++)
      step[i] = (c2[i] - c1[i]) / (double)npnt;
    for (unsigned int ipnt = 1; ipnt < npnt; ipnt++)
    {
      for (int i = 0; i < d; i++)
        conf[i] = c1[i] + step[i] * (double)ipnt;
      if (!world_check(w, conf))
      {
        for (int i = 0; i < d; i++)
          conf[i] = c1[i] + step[i] * (double)(ipnt - 1);
        *first_bad = ipnt;
        return false;
      }
    }
  }
  if (!world_check(w, c2))
  {
    *first_bad = (npnt > 1) ? (long long)npnt : 1;
    return false;
  }
  return true;
}

// checkConnFromConf(conn,this_conf) :264-313 ; returns -1 when the reference throws std::invalid_argument
static int checkConnFromConf(World& w, const double* parent, const double* child, const double* this_conf,
                             double min_distance)
{
  const int d = w.dim;
  double dist_child = dist(this_conf, child, d);
  double dist_parent = dist(parent, this_conf, d);
  double dst = dist(parent, child, d);
  if ((dst - dist_child - dist_parent) > 1e-04)
    return -1;
  if (!world_check(w, this_conf))
    return 0;
  if (!world_check(w, child))
    return 0;
  double distance = dist(this_conf, child, d);
  if (distance < min_distance)
    return 1;
  double this_abscissa = dist(parent, this_conf, d) / dist(parent, child, d);
  double n = 2;
  double conf[GC_MAX_DIM];
  while (distance > n * min_distance)
  {
    for (double idx = 1; idx < n; idx += 2)
    {
      double abscissa = idx / n;
      if (abscissa >= this_abscissa)
      {
        for (int i = 0; i < d; i++)
          conf[i] = parent[i] + (child[i] - parent[i]) * abscissa;
        if (!world_check(w, conf))
          return 0;
      }
    }
    n *= 2;
  }
  return 1;
}

// ============================================================================
//  Samplers' specific volume (inc/samplers/sampler_base.h:135-153, src/samplers/informed_sampler.cpp:181-233)
// ============================================================================
struct InformedSamplerState
{
  int ndof = 0;
  std::vector<double> lb, ub, f1, f2;
  double cost_ = kInf, focii_distance_ = 0, min_radius_ = 0, max_radius_ = 0, specific_volume_ = 0;
  bool inf_cost_ = true;
  void init(int d, const double* lower, const double* upper, const double* focus1, const double* focus2)
  {
    ndof = d;
    lb.assign(lower, lower + d);
    ub.assign(upper, upper + d);
    f1.assign(focus1, focus1 + d);
    f2.assign(focus2, focus2 + d);
    focii_distance_ = dist(f1.data(), f2.data(), d);  // informed_sampler.cpp:86 (scale == 1)
    setCost(kInf);
  }
  void setCost(double cost)  // informed_sampler.cpp:181-215
  {
    cost_ = cost;
    inf_cost_ = cost_ >= kInf;
    if (cost_ < focii_distance_)
    {
      cost_ = focii_distance_;
      min_radius_ = 0.0;
    }
    else
      min_radius_ = 0.5 * std::sqrt(std::pow(cost_, 2) - std::pow(focii_distance_, 2));
    max_radius_ = 0.5 * cost_;
    if (inf_cost_)
    {
      specific_volume_ = std::tgamma(((double)ndof) * 0.5 + 1.0) / std::pow(M_PI, (double)ndof * 0.5);
      for (int i = 0; i < ndof; i++)
        specific_volume_ *= (ub[i] - lb[i]);
    }
    else
      specific_volume_ = max_radius_ * std::pow(min_radius_, ndof - 1);
  }
  bool inBoundsBox(const double* q) const
  {
    for (int i = 0; i < ndof; i++)
      if (q[i] > ub[i] || q[i] < lb[i])
        return false;
    return true;
  }
};

// ============================================================================
//  Tree (src/graph/tree.cpp) -- parent-array restatement of Node/Connection bookkeeping
// ============================================================================
struct Counters
{
  unsigned long long nn1 = 0, near = 0, knn = 0, edges = 0, near_hits = 0;
};

struct Tree
{
  NodePool* pool;
  std::unique_ptr<NearestNeighbors> nodes_;
  World* checker_;
  double min_distance_;
  double max_distance_;
  int root_;
  double k_rrt_;
  std::vector<int> parent;     // parent[id] = parent node id (-1: none)
  std::vector<double> pcost;   // cost of the parent connection
  Counters cnt;

  Tree(NodePool* p, int root, double max_distance, World* checker, double min_distance, bool use_kdtree)
    : pool(p), checker_(checker), min_distance_(min_distance), max_distance_(max_distance), root_(root)
  {  // tree.cpp:34-54
    if (use_kdtree)
      nodes_.reset(new KdTreeNN(p));
    else
      nodes_.reset(new VectorNN(p));
    nodes_->insert(root);
    double dimension = p->dim;
    k_rrt_ = 1.1 * std::pow(2.0, dimension + 1) * std::exp(1) * (1.0 + 1.0 / dimension);
  }
  void ensure(int id)
  {
    if ((int)parent.size() <= id)
    {
      parent.resize(id + 1, -1);
      pcost.resize(id + 1, 0.0);
    }
  }
  void setParent(int child, int par, double cost)
  {
    ensure(std::max(child, par));
    parent[child] = par;
    pcost[child] = cost;
  }
  bool checkConnection(const double* a, const double* b)
  {
    cnt.edges++;
    return checkConnectionBisect(*checker_, a, b, min_distance_);
  }
  int findClosestNode(const double* q)  // tree.cpp:56-59
  {
    int best = -1;
    double bd;
    cnt.nn1++;
    nodes_->nearestNeighbor(q, best, bd);
    return best;
  }
  // tree.cpp:110-128
  double selectNextConfiguration(const double* q, double* next, int node)
  {
    const int d = pool->dim;
    const double* nc = pool->conf(node);
    double distance = dist(nc, q, d);
    if (distance <= max_distance_)
      for (int i = 0; i < d; i++) next[i] = q[i];
    else
      for (int i = 0; i < d; i++) next[i] = nc[i] + (q[i] - nc[i]) / distance * max_distance_;
    return distance;
  }
  bool tryExtendFromNode(const double* q, double* next, int node)  // tree.cpp:69-81
  {
    double distance = selectNextConfiguration(q, next, node);
    if (distance < TOLERANCE)
      return true;
    return checkConnection(pool->conf(node), next);
  }
  bool extend(const double* q, int& new_node)  // tree.cpp:148-161 + extendOnly :130-140
  {
    int closest = findClosestNode(q);
    double next[GC_MAX_DIM];
    if (!tryExtendFromNode(q, next, closest))
    {
      new_node = -1;
      return false;
    }
    new_node = pool->add(next);
    double cost = dist(pool->conf(closest), pool->conf(new_node), pool->dim);  // metrics_->cost
    setParent(new_node, closest, cost);
    nodes_->insert(new_node);
    return true;
  }
  bool isInTree(int node) { return nodes_->findNode(node); }
  bool extendToNode(int node, int& new_node)  // tree.cpp:194-215
  {
    if (isInTree(node))
      return true;
    int closest = findClosestNode(pool->conf(node));
    double next[GC_MAX_DIM];
    if (!tryExtendFromNode(pool->conf(node), next, closest))
      return false;
    if (dist(next, pool->conf(node), pool->dim) < TOLERANCE)
      new_node = node;
    else
      new_node = pool->add(next);
    nodes_->insert(new_node);
    double cost = dist(pool->conf(closest), pool->conf(new_node), pool->dim);
    setParent(new_node, closest, cost);
    return true;
  }
  bool connect(const double* q, int& new_node)  // tree.cpp:217-233
  {
    bool success = true;
    while (success)
    {
      int tmp = -1;
      success = extend(q, tmp);
      if (success)
      {
        new_node = tmp;
        if (dist(pool->conf(new_node), q, pool->dim) < TOLERANCE)
          return true;
      }
    }
    return false;
  }
  bool connectToNode(int node, int& new_node)  // tree.cpp:304-327 (no time limit)
  {
    bool success = true;
    while (success)
    {
      int tmp = -1;
      success = extendToNode(node, tmp);
      if (success)
      {
        if (tmp < 0)  // node already in tree: the reference leaves tmp_node null and would crash; stop
          return true;
        new_node = tmp;
        if (dist(pool->conf(new_node), pool->conf(node), pool->dim) < TOLERANCE)
          return true;
      }
    }
    return false;
  }
  double costToNode(int node)  // tree.cpp:767-789
  {
    double cost = 0;
    while (node != root_)
    {
      if (node < 0 || node >= (int)parent.size() || parent[node] < 0)
        return kInf;
      cost += pcost[node];
      node = parent[node];
    }
    return cost;
  }
  std::vector<int> getConnectionToNode(int node)  // tree.cpp:791-813: returns child ids root->node
  {
    std::vector<int> chain;
    int t = node;
    while (t != root_)
    {
      chain.push_back(t);
      t = parent[t];
    }
    std::reverse(chain.begin(), chain.end());
    return chain;
  }
  double pathCost(const std::vector<int>& chain)  // Path::computeCost path.cpp:285-291
  {
    double c = 0;
    for (int id : chain)
      c += pcost[id];
    return c;
  }
  NearMap nearK(const double* q)  // tree.cpp:761-765
  {
    size_t k = (size_t)std::ceil(k_rrt_ * std::log(nodes_->size_ + 1));
    cnt.knn++;
    return nodes_->kNearestNeighbors(q, k);
  }
  // tree.cpp:381-513 (what_rewire = 0, empty white list)
  bool rewireOnly(int node, double r_rewire)
  {
    const int d = pool->dim;
    bool rewire_parent = (node != root_);
    NearMap near_nodes;
    if (r_rewire <= 0)
      near_nodes = nearK(pool->conf(node));
    else
    {
      cnt.near++;
      near_nodes = nodes_->near(pool->conf(node), r_rewire);
    }
    cnt.near_hits += near_nodes.size();
    double cost_to_node = costToNode(node);
    bool improved = false;
    if (rewire_parent)
    {
      int nearest_node = parent[node];
      for (const auto& p : near_nodes)
      {
        int n = p.second;
        if (n == nearest_node)
          continue;
        if (n == node)
          continue;
        double cost_to_near = costToNode(n);
        if (cost_to_near >= cost_to_node)
          continue;
        double cost_near_to_node = dist(pool->conf(n), pool->conf(node), d);
        if ((cost_to_near + cost_near_to_node) >= cost_to_node)
          continue;
        if (!checkConnection(pool->conf(n), pool->conf(node)))
          continue;
        setParent(node, n, cost_near_to_node);
        cost_to_node = cost_to_near + cost_near_to_node;
        improved = true;
      }
    }
    int par = (node == root_) ? -1 : parent[node];
    for (const auto& p : near_nodes)
    {
      int n = p.second;
      if (n == par)
        continue;
      if (n == node)
        continue;
      if (n == root_)
        continue;
      double cost_to_near = costToNode(n);
      if (cost_to_node >= cost_to_near)
        continue;
      double cost_node_to_near = dist(pool->conf(node), pool->conf(n), d);
      if ((cost_to_node + cost_node_to_near) >= cost_to_near)
        continue;
      if (!checkConnection(pool->conf(node), pool->conf(n)))
        continue;
      setParent(n, node, cost_node_to_near);
      improved = true;
    }
    return improved;
  }
  bool rewire(const double* q, double r_rewire, int& new_node)  // tree.cpp:718-726
  {
    if (!extend(q, new_node))
      return false;
    return rewireOnly(new_node, r_rewire);
  }
  // tree.cpp:235-302
  bool informedExtend(const double* q, int& new_node, const double* goal, double cost2beat, double bias)
  {
    const int d = pool->dim;
    NearMap closest = nearK(q);
    if (closest.empty())
      throw std::runtime_error("closest nodes map is empty");
    struct Ext
    {
      int tree_node;
      std::vector<double> new_conf;
      double distance;
    };
    std::multimap<double, Ext> best;
    double nc[GC_MAX_DIM];
    for (const auto& n : closest)
    {
      double distance = selectNextConfiguration(q, nc, n.second);
      double cost2node = costToNode(n.second);
      double heuristic = bias * distance + (1 - bias) * cost2node;
      if ((cost2node + distance + dist(goal, nc, d)) < cost2beat)
      {
        Ext e;
        e.tree_node = n.second;
        e.new_conf.assign(nc, nc + d);
        e.distance = distance;
        best.insert(std::make_pair(heuristic, e));
      }
    }
    bool ok = false;
    Ext ext;
    for (const auto& n : best)
    {
      ext = n.second;
      if (ext.distance < TOLERANCE)
      {
        ok = true;
        break;
      }
      if (checkConnection(pool->conf(ext.tree_node), ext.new_conf.data()))
      {
        ok = true;
        break;
      }
    }
    if (!ok)
      return false;
    new_node = pool->add(ext.new_conf.data());
    double cost = dist(pool->conf(ext.tree_node), pool->conf(new_node), d);
    setParent(new_node, ext.tree_node, cost);
    nodes_->insert(new_node);
    return true;
  }
};

// ============================================================================
//  Solvers (src/solvers/{tree_solver,rrt,rrt_star}.cpp), driven by injected samples (tree_solver.h:338)
// ============================================================================
struct Solver
{
  enum Kind { RRT = 0, RRTSTAR = 1 };
  int kind;
  NodePool pool;
  World* world;
  std::unique_ptr<Tree> start_tree_;
  InformedSamplerState sampler_;
  int goal_node_ = -1;
  double min_distance_, max_distance_, rewire_factor_, utopia_tolerance_;
  bool use_kdtree_, extend_;
  bool solved_ = false, can_improve_ = true, problem_set_ = false;
  double goal_cost_ = 0, best_utopia_ = 0, path_cost_ = kInf, cost_ = kInf, r_rewire_ = 0;
  std::vector<int> solution_;  // child ids root->goal
  unsigned long long iterations = 0;

  Solver(int k, int dim, World* w, const double* lb, const double* ub, const double* start, const double* goal,
         double max_distance, double min_distance, double rewire_factor, double utopia_tolerance, bool use_kdtree,
         bool extend)
    : kind(k), pool(dim), world(w), min_distance_(min_distance), max_distance_(max_distance),
      rewire_factor_(rewire_factor), use_kdtree_(use_kdtree), extend_(extend)
  {
    // tree_solver.cpp:34-53 config
    utopia_tolerance_ = utopia_tolerance;
    if (utopia_tolerance_ <= 0.0)
      utopia_tolerance_ = 0.0;
    utopia_tolerance_ += 1.0;
    sampler_.init(dim, lb, ub, start, goal);
    // addStart (rrt.cpp:58-78)
    int s = pool.add(start);
    start_tree_.reset(new Tree(&pool, s, max_distance_, world, min_distance_, use_kdtree_));
    // addGoal (rrt.cpp:34-56) -> setProblem (tree_solver.cpp:55-105)
    goal_node_ = pool.add(goal);
    setProblem();
  }
  double solutionCost() { return start_tree_->pathCost(solution_); }  // Path::cost() re-sums stored connection costs
  void makeSolution()
  {
    solution_ = start_tree_->getConnectionToNode(goal_node_);
    // Path keeps the Connection objects alive: their costs never change afterwards, cache them
    sol_costs_.clear();
    for (int id : solution_)
      sol_costs_.push_back(start_tree_->pcost[id]);
  }
  std::vector<double> sol_costs_;
  double cachedSolutionCost()
  {
    double c = 0;
    for (double v : sol_costs_)
      c += v;
    return c;
  }
  void setProblem()
  {
    const int d = pool.dim;
    can_improve_ = true;
    goal_cost_ = 0.0;  // NullGoalCostFunction
    best_utopia_ = goal_cost_ + dist(pool.conf(start_tree_->root_), pool.conf(goal_node_), d);
    problem_set_ = true;
    int new_node = -1;
    if (start_tree_->connectToNode(goal_node_, new_node))
    {
      makeSolution();
      path_cost_ = cachedSolutionCost();
      sampler_.setCost(path_cost_);
      solved_ = true;
    }
    else
      path_cost_ = kInf;
    cost_ = path_cost_ + goal_cost_;
  }
  void updateRewireRadius()  // rrt_star.cpp:278-286
  {
    double dimension = (double)pool.dim;
    double r_rrt = rewire_factor_ * std::pow(2.0 * (1.0 + 1.0 / dimension), 1.0 / dimension) *
                   std::pow(sampler_.specific_volume_, 1.0 / dimension);
    double cardDbl = start_tree_->nodes_->size_ + 1.0;
    r_rewire_ = (r_rrt * std::pow(std::log(cardDbl) / cardDbl, 1.0 / dimension));
  }
  void connectGoal(int new_node)
  {
    const int d = pool.dim;
    start_tree_->setParent(goal_node_, new_node, dist(pool.conf(new_node), pool.conf(goal_node_), d));
    makeSolution();
    start_tree_->nodes_->insert(goal_node_);  // addNode(goal) with check_if_present: not present yet
    path_cost_ = cachedSolutionCost();
    cost_ = path_cost_ + goal_cost_;
    sampler_.setCost(path_cost_);
    solved_ = true;
  }
  // RRTStar::update(configuration, solution)  rrt_star.cpp:89-163
  bool updateRRTStar(const double* q)
  {
    const int d = pool.dim;
    if (!problem_set_)
      return false;
    if (cost_ <= utopia_tolerance_ * best_utopia_)
    {
      can_improve_ = false;
      return true;
    }
    updateRewireRadius();
    if (!solved_)
    {
      int new_node = -1;
      if (start_tree_->rewire(q, r_rewire_, new_node))
      {
        if (dist(pool.conf(new_node), pool.conf(goal_node_), d) < max_distance_)
        {
          if (start_tree_->checkConnection(pool.conf(new_node), pool.conf(goal_node_)))
          {
            connectGoal(new_node);
            return true;
          }
        }
      }
      return false;
    }
    else
    {
      int nn = -1;
      bool improved = start_tree_->rewire(q, r_rewire_, nn);
      if (improved)
      {
        if (start_tree_->costToNode(goal_node_) >= (cachedSolutionCost() - 1e-8))
          return false;
        makeSolution();
        path_cost_ = cachedSolutionCost();
        cost_ = path_cost_ + goal_cost_;
        sampler_.setCost(path_cost_);
      }
      return improved;
    }
  }
  // RRT::update(configuration, solution)  rrt.cpp:116-157
  bool updateRRT(const double* q)
  {
    const int d = pool.dim;
    if (solved_)
    {
      can_improve_ = false;
      return true;
    }
    int new_node = -1;
    bool added = extend_ ? start_tree_->extend(q, new_node) : start_tree_->connect(q, new_node);
    if (added)
    {
      if (dist(pool.conf(new_node), pool.conf(goal_node_), d) <= max_distance_)
      {
        if (start_tree_->checkConnection(pool.conf(new_node), pool.conf(goal_node_)))
        {
          connectGoal(new_node);
          can_improve_ = false;
          return true;
        }
      }
    }
    return false;
  }
  bool update(const double* q)
  {
    iterations++;
    return kind == RRTSTAR ? updateRRTStar(q) : updateRRT(q);
  }
};


// ============================================================================
//  AnytimeRRT (src/graph_core/solvers/anytime_rrt.cpp), BASELINE config 5.
//  The RRT phase is RRT::solve = TreeSolver::solve (tree_solver.cpp:107-130) over RRT::update (rrt.cpp:116-157);
//  the improve rounds are AnytimeRRT::improve (:269-330) over improveUpdate (:356-427) and Tree::informedExtend.
//  NOTE (reference as written): RRT::update leaves can_improve_ = false when it connects the goal (rrt.cpp:150), so the
//  improve loop of AnytimeRRT::solve (:134) is never entered from solve(); the rounds are reached through the public
//  improve().  `improve_in_solve` runs the loop of :134-171 with improve() called unconditionally, which is what a
//  caller who wants the anytime behaviour writes on top of the public API.
// ============================================================================
struct AnytimeSolver
{
  static const int kFailedIter = 3;  // anytime_rrt.h:35
  Solver base;
  double bias_ = 0.9, delta_ = 0.1, cost_impr_ = 0.1, cost2beat_ = 0.0;  // setParameters (anytime_rrt.h:145-150)
  std::unique_ptr<Tree> new_tree_;
  int tmp_goal_node_ = -1;
  std::vector<double> start_conf, goal_conf;
  // per round: phase, iterations, success, path cost afterwards, nodes in the round's tree, bias, cost2beat, set up
  std::vector<double> log;
  unsigned long long samples_drawn = 0;

  AnytimeSolver(int dim, World* w, const double* lb, const double* ub, const double* start, const double* goal,
                double max_distance, double min_distance, double utopia_tolerance, bool use_kdtree, bool extend,
                double delta, double cost_impr)
    : base(Solver::RRT, dim, w, lb, ub, start, goal, max_distance, min_distance, 1.1, utopia_tolerance, use_kdtree,
           extend),
      delta_(delta), cost_impr_(cost_impr), start_conf(start, start + dim), goal_conf(goal, goal + dim)
  {
  }
  void pushLog(int phase, unsigned iters, bool ok, double tree_nodes, bool setup)
  {
    const double row[8] = { (double)phase, (double)iters, ok ? 1.0 : 0.0, base.path_cost_, tree_nodes, bias_,
                            cost2beat_,   setup ? 1.0 : 0.0 };
    log.insert(log.end(), row, row + 8);
  }
  // AnytimeRRT::improve(start, goal, solution, cost2beat, max_iter) up to the loop (:269-315); false: returned early
  bool improveBegin()
  {
    const int d = base.pool.dim;
    const double cost2beat = (1.0 - cost_impr_) * base.path_cost_;  // :262
    const double utopia = dist(goal_conf.data(), start_conf.data(), d);
    base.can_improve_ = true;
    if (base.cost_ <= base.utopia_tolerance_ * utopia)
    {
      base.can_improve_ = false;
      return false;
    }
    if (cost2beat <= utopia)
      return false;
    const int tmp_start = base.pool.add(start_conf.data());
    tmp_goal_node_ = base.pool.add(goal_conf.data());
    new_tree_.reset(new Tree(&base.pool, tmp_start, base.max_distance_, base.world, base.min_distance_,
                             base.use_kdtree_));
    cost2beat_ = cost2beat;
    bias_ = bias_ - delta_;
    if (bias_ < 0.1)
      bias_ = 0.1;
    return true;
  }
  // AnytimeRRT::improveUpdate(point, solution)  :356-427
  bool improveUpdate(const double* q)
  {
    const int d = base.pool.dim;
    if (!base.can_improve_)
      return true;
    int new_node = -1;
    if (new_tree_->informedExtend(q, new_node, goal_conf.data(), cost2beat_, bias_))
    {
      if (dist(base.pool.conf(new_node), goal_conf.data(), d) <= base.max_distance_)
      {
        const std::vector<int> chain = new_tree_->getConnectionToNode(new_node);
        const double cost_node2goal = dist(base.pool.conf(new_node), goal_conf.data(), d);
        double new_solution_cost = cost_node2goal;
        for (int id : chain)
          new_solution_cost += new_tree_->pcost[id];
        if (new_solution_cost < base.cachedSolutionCost())
        {
          if (new_tree_->checkConnection(base.pool.conf(new_node), goal_conf.data()))
          {
            base.goal_node_ = tmp_goal_node_;
            new_tree_->setParent(base.goal_node_, new_node, cost_node2goal);
            new_tree_->nodes_->insert(base.goal_node_);
            base.start_tree_ = std::move(new_tree_);
            base.makeSolution();
            base.path_cost_ = base.cachedSolutionCost();
            base.cost_ = base.path_cost_ + base.goal_cost_;
            return true;
          }
        }
      }
    }
    return false;
  }
  template <class Sampler>
  bool rrtRound(unsigned max_iter, Sampler&& sample)  // TreeSolver::solve
  {
    double q[GC_MAX_DIM];
    unsigned it = 0;
    bool ok = false;
    for (; it < max_iter; it++)
    {
      if (base.solved_)  // RRT::update(PathPtr&) :102-108
      {
        base.can_improve_ = false;
        ok = true;
        break;
      }
      sample(kInf0(), q);
      if (base.update(q))
      {
        ok = true;
        it++;
        break;
      }
    }
    pushLog(0, it, ok, (double)base.start_tree_->nodes_->size_, true);
    return ok;
  }
  double sampler_cost0 = kInf;  // cost bound of the solver's own sampler during the RRT phase
  double kInf0() const { return sampler_cost0; }
  template <class Sampler>
  bool improveRound(unsigned max_iter, Sampler&& sample)
  {
    if (!improveBegin())
    {
      pushLog(1, 0, false, 0.0, false);
      return false;
    }
    double q[GC_MAX_DIM];
    unsigned it = 0;
    bool ok = false;
    for (; it < max_iter; it++)
    {
      sample(cost2beat_, q);
      if (improveUpdate(q))
      {
        ok = true;
        it++;
        break;
      }
    }
    pushLog(1, it, ok, ok ? (double)base.start_tree_->nodes_->size_ : (double)new_tree_->nodes_->size_, true);
    return ok;
  }
  // AnytimeRRT::solve (:73-236) without the wall clock
  template <class Sampler>
  bool solve(unsigned max_iter, bool improve_in_solve, Sampler&& sample)
  {
    if (base.solved_ && base.cost_ <= base.utopia_tolerance_ * base.best_utopia_)
    {
      base.can_improve_ = false;
      return true;
    }
    int n_failed = 0;
    while (!base.solved_ && n_failed < kFailedIter)
      if (!rrtRound(max_iter, sample))
        n_failed++;
    if (!base.solved_)
      return false;
    if (base.cost_ <= base.utopia_tolerance_ * base.best_utopia_)
    {
      base.can_improve_ = false;
      return true;
    }
    n_failed = 0;
    while ((improve_in_solve || base.can_improve_) && n_failed < kFailedIter)
    {
      const bool improved = improveRound(max_iter, sample);
      improved ? (n_failed = 0) : (n_failed++);
      if (base.cost_ <= base.utopia_tolerance_ * base.best_utopia_)
      {
        base.can_improve_ = false;
        break;
      }
    }
    return base.solved_;
  }
};

}  // namespace orc

// ============================================================================
//  Informed sampling with the repository's Philox stream (include/gcb200.h gc_sample_informed).
//  Algorithm: src/samplers/informed_sampler.cpp:107-143 (rotation basis), :145-180 (sample),
//  :181-207 (ellipse radii).  The random stream and u^(1/D) (64-step bisection) are this
//  repository's specification, restated here independently of csrc/planner.cu.
// ============================================================================
namespace orc
{
static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out)
{
  for (int r = 0; r < 10; r++)
  {
    const uint64_t p0 = (uint64_t)GC_PHILOX_M0 * c0, p1 = (uint64_t)GC_PHILOX_M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1,
                   n3 = (uint32_t)p0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += GC_PHILOX_W0;
    k1 += GC_PHILOX_W1;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}
struct PhiloxDraws
{
  uint64_t ctr, seed;
  double operator()(uint32_t j) const
  {
    uint32_t r[4];
    philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), j >> 1, 1u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const uint64_t x = (j & 1u) ? (((uint64_t)r[2] << 32) | r[3]) : (((uint64_t)r[0] << 32) | r[1]);
    return (double)(x >> 11) * 1.1102230246251565404e-16;
  }
};
static double dth_root(double u, int D)
{
  double lo = 0.0, hi = 1.0;
  for (int it = 0; it < 64; it++)
  {
    const double mid = 0.5 * (lo + hi);
    double p = mid;
    for (int k = 1; k < D; k++)
      p = p * mid;
    if (p < u)
      lo = mid;
    else
      hi = mid;
  }
  return 0.5 * (lo + hi);
}
// returns the attempts code of gc_sample_informed
static unsigned informed_sample(int D, const double* lb, const double* ub, const double* f1, const double* f2,
                                double cost, const PhiloxDraws& draw, double* out)
{
  auto uniform = [&](uint32_t j0) {
    for (int d = 0; d < D; d++)
    {
      const double u = draw(j0 + d);
      out[d] = 0.5 * (lb[d] + ub[d]) + (2.0 * u - 1.0) * (0.5 * (lb[d] - ub[d]));
    }
  };
  if (!(cost < kInf))
  {
    uniform(0);
    return 0;
  }
  std::vector<double> v(D), centre(D), axis(D), R((size_t)D * D, 0.0);  // R[row * D + col]
  double s2 = 0.0;
  for (int d = 0; d < D; d++)
  {
    v[d] = f1[d] - f2[d];
    s2 += v[d] * v[d];
    centre[d] = 0.5 * (f1[d] + f2[d]);
  }
  const double fd = std::sqrt(s2);
  for (int d = 0; d < D; d++)
    v[d] = v[d] / fd;
  for (int d = 0; d < D; d++)
    R[(size_t)d * D + d] = 1.0;
  bool standard = false;
  for (int ic = 0; ic < D && !standard; ic++)
    if (std::fabs(v[ic]) > 0.999)  // main_versor . e_ic (:118)
    {
      standard = true;
      for (int r = 0; r < D; r++)
        std::swap(R[(size_t)r * D + ic], R[(size_t)r * D]);
    }
  if (!standard)
  {
    for (int r = 0; r < D; r++)
      R[(size_t)r * D] = v[r];
    for (int ic = 1; ic < D; ic++)
    {
      for (int il = 0; il < ic; il++)
      {
        double dot = 0.0;
        for (int r = 0; r < D; r++)
          dot += R[(size_t)r * D + ic] * R[(size_t)r * D + il];
        for (int r = 0; r < D; r++)
          R[(size_t)r * D + ic] = R[(size_t)r * D + ic] - dot * R[(size_t)r * D + il];
      }
      double n2 = 0.0;
      for (int r = 0; r < D; r++)
        n2 += R[(size_t)r * D + ic] * R[(size_t)r * D + ic];
      const double nn = std::sqrt(n2);
      for (int r = 0; r < D; r++)
        R[(size_t)r * D + ic] = R[(size_t)r * D + ic] / nn;
    }
  }
  double rmin;
  if (cost < fd)
  {
    cost = fd;
    rmin = 0.0;
  }
  else
    rmin = 0.5 * std::sqrt(cost * cost - fd * fd);
  const double rmax = 0.5 * cost;
  for (int d = 0; d < D; d++)
    axis[d] = d == 0 ? rmax : rmin;
  std::vector<double> ball(D), q(D);
  for (unsigned t = 0; t < 100; t++)
  {
    const uint32_t j0 = t * (uint32_t)(D + 1);
    double b2 = 0.0;
    for (int d = 0; d < D; d++)
    {
      ball[d] = 2.0 * draw(j0 + d) - 1.0;
      b2 += ball[d] * ball[d];
    }
    const double scale = dth_root(draw(j0 + D), D) / std::sqrt(b2);
    bool inside = true;
    for (int r = 0; r < D; r++)
    {
      double acc = 0.0;
      for (int c = 0; c < D; c++)
        acc += (R[(size_t)r * D + c] * axis[c]) * (ball[c] * scale);
      q[r] = acc + centre[r];
      if (q[r] > ub[r] || q[r] < lb[r])
        inside = false;
    }
    if (inside)
    {
      for (int d = 0; d < D; d++)
        out[d] = q[d];
      return t + 1;
    }
  }
  uniform(100u * (uint32_t)(D + 1));
  return 101;
}
}  // namespace orc

// ============================================================================
//  C interface (ctypes)
// ============================================================================
using namespace orc;

struct OrcNN
{
  NodePool pool;
  std::unique_ptr<NearestNeighbors> nn;
  OrcNN(int kind, int dim) : pool(dim)
  {
    if (kind == 1)
      nn.reset(new KdTreeNN(&pool));
    else
      nn.reset(new VectorNN(&pool));
  }
};

static int copy_map(const NearMap& m, int* ids, double* dists, int cap)
{
  int n = 0;
  for (const auto& p : m)
  {
    if (n < cap)
    {
      if (ids) ids[n] = p.second;
      if (dists) dists[n] = p.first;
    }
    n++;
  }
  return n;
}

extern "C" {

void* orc_nn_new(int kind, int dim) { return new OrcNN(kind, dim); }
void orc_nn_free(void* h) { delete (OrcNN*)h; }
int orc_nn_insert(void* h, const double* q)
{
  OrcNN* o = (OrcNN*)h;
  int id = o->pool.add(q);
  o->nn->insert(id);
  return id;
}
void orc_nn_insert_many(void* h, const double* q, int n)
{
  OrcNN* o = (OrcNN*)h;
  for (int i = 0; i < n; i++)
    o->nn->insert(o->pool.add(q + (size_t)i * o->pool.dim));
}
void orc_nn_reinsert(void* h, int id) { ((OrcNN*)h)->nn->insert(id); }
int orc_nn_delete(void* h, int id) { return ((OrcNN*)h)->nn->deleteNode(id) ? 1 : 0; }
int orc_nn_restore(void* h, int id) { return ((OrcNN*)h)->nn->restoreNode(id) ? 1 : 0; }
int orc_nn_find(void* h, int id) { return ((OrcNN*)h)->nn->findNode(id) ? 1 : 0; }
int orc_nn_clear(void* h) { return ((OrcNN*)h)->nn->clear() ? 1 : 0; }
unsigned int orc_nn_size(void* h) { return ((OrcNN*)h)->nn->size_; }
unsigned int orc_nn_deleted(void* h) { return ((OrcNN*)h)->nn->deleted_nodes_; }
void orc_nn_set_deleted_threshold(void* h, unsigned int t)
{
  KdTreeNN* k = dynamic_cast<KdTreeNN*>(((OrcNN*)h)->nn.get());
  if (k && t > 1)
    k->deleted_nodes_threshold_ = t;
}
int orc_nn_get_nodes(void* h, int* ids, int cap)
{
  std::vector<int> v = ((OrcNN*)h)->nn->getNodes();
  for (int i = 0; i < (int)v.size() && i < cap; i++)
    ids[i] = v[i];
  return (int)v.size();
}
// nq queries; id = -1 and dist = +inf on an empty structure (kdtree.cpp:356-358)
void orc_nn_nearest(void* h, const double* q, int nq, int* ids, double* dists)
{
  OrcNN* o = (OrcNN*)h;
  for (int i = 0; i < nq; i++)
  {
    int best = -1;
    double bd = kInf;
    o->nn->nearestNeighbor(q + (size_t)i * o->pool.dim, best, bd);
    ids[i] = best;
    dists[i] = bd;
  }
}
// each query writes up to cap entries at ids[i*cap..]; counts[i] = full hit count
void orc_nn_near(void* h, const double* q, int nq, const double* radius, int cap, int* ids, double* dists,
                 int* counts)
{
  OrcNN* o = (OrcNN*)h;
  for (int i = 0; i < nq; i++)
  {
    NearMap m = o->nn->near(q + (size_t)i * o->pool.dim, radius[i]);
    counts[i] = copy_map(m, ids + (size_t)i * cap, dists + (size_t)i * cap, cap);
  }
}
void orc_nn_knn(void* h, const double* q, int nq, unsigned long long k, int cap, int* ids, double* dists,
                int* counts)
{
  OrcNN* o = (OrcNN*)h;
  for (int i = 0; i < nq; i++)
  {
    NearMap m = o->nn->kNearestNeighbors(q + (size_t)i * o->pool.dim, (size_t)k);
    counts[i] = copy_map(m, ids + (size_t)i * cap, dists + (size_t)i * cap, cap);
  }
}

// ---- worlds -----------------------------------------------------------------
void* orc_world_cube(int dim, double thr)
{
  World* w = new World();
  w->kind = GC_WORLD_CUBE;
  w->dim = dim;
  w->cube_thr = thr;
  return w;
}
void* orc_world_boxes(int dim, int n, const double* lo, const double* hi)
{
  World* w = new World();
  w->kind = GC_WORLD_BOXES;
  w->dim = dim;
  w->nobj = n;
  w->lo.assign(lo, lo + (size_t)n * dim);
  w->hi.assign(hi, hi + (size_t)n * dim);
  return w;
}
void* orc_world_cspheres(int dim, int n, const float* centers, const float* radii)
{
  World* w = new World();
  w->kind = GC_WORLD_CSPHERES;
  w->dim = dim;
  w->nobj = n;
  w->sc.assign(centers, centers + (size_t)n * dim);
  w->sr2.resize(n);
  for (int i = 0; i < n; i++)
    w->sr2[i] = radii[i] * radii[i];
  return w;
}
void* orc_world_sphere_arm(int njoints, const float* base_tf, const float* joint_tf, int nrs, const int* rs_link,
                           const float* rs_local, int nobs, const float* obs)
{
  World* w = new World();
  w->kind = GC_WORLD_SPHERE_ARM;
  w->dim = njoints;
  w->njoints = njoints;
  w->nrs = nrs;
  w->nobs = nobs;
  w->base_tf.assign(base_tf, base_tf + 12);
  w->joint_tf.assign(joint_tf, joint_tf + (size_t)12 * njoints);
  w->rs_link.assign(rs_link, rs_link + nrs);
  w->rs_local.assign(rs_local, rs_local + (size_t)4 * nrs);
  w->obs.assign(obs, obs + (size_t)4 * nobs);
  w->obs_na.resize(nobs);
  for (int o = 0; o < nobs; o++)
  {
    const float* ob = obs + (size_t)4 * o;
    float on2 = ob[0] * ob[0];
    on2 = __builtin_fmaf(ob[1], ob[1], on2);
    on2 = __builtin_fmaf(ob[2], ob[2], on2);
    w->obs_na[o] = ob[3] * ob[3] - on2;
  }
  return w;
}
// Builds the culling grid of a sphere-arm world (see World::grid_on); returns the number of cells, 0 = not built.
int orc_world_enable_grid(void* wh, int on)
{
  World* w = (World*)wh;
  w->grid_on = false;
  if (!on || w->kind != GC_WORLD_SPHERE_ARM || w->nobs < 1 || w->nobs > 65535)
    return 0;
  auto norm3 = [](const float* t) { return std::sqrt((double)t[0] * t[0] + (double)t[1] * t[1] + (double)t[2] * t[2]); };
  double olo[3] = { 1e300, 1e300, 1e300 }, ohi[3] = { -1e300, -1e300, -1e300 }, omax = 0, romax = 0, rsmax = 0;
  for (int i = 0; i < w->nobs; i++)
  {
    for (int d = 0; d < 4; d++)
      if (!(std::fabs((double)w->obs[4 * (size_t)i + d]) < 1e18))
        return 0;
    for (int d = 0; d < 3; d++)
    {
      olo[d] = std::min(olo[d], (double)w->obs[4 * (size_t)i + d]);
      ohi[d] = std::max(ohi[d], (double)w->obs[4 * (size_t)i + d]);
    }
    omax = std::max(omax, norm3(&w->obs[4 * (size_t)i]));
    romax = std::max(romax, std::fabs((double)w->obs[4 * (size_t)i + 3]));
  }
  double reach = norm3(&w->base_tf[9]), smax = 0;
  for (int j = 0; j < w->njoints; j++)
    reach += norm3(&w->joint_tf[12 * (size_t)j + 9]);
  for (int k = 0; k < w->nrs; k++)
  {
    smax = std::max(smax, norm3(&w->rs_local[4 * (size_t)k]));
    rsmax = std::max(rsmax, std::fabs((double)w->rs_local[4 * (size_t)k + 3]));
  }
  reach += smax;
  if (!(reach < 1e18))
    return 0;
  const double E = reach + omax + std::max(rsmax, romax), m = E / 256.0;
  const double pad = rsmax + romax + m * (1.0 + 1e-6);
  double lo[3], hi[3], ext = 0;
  for (int d = 0; d < 3; d++)
  {
    lo[d] = std::max(olo[d] - pad, -reach) - 1e-6 * E;
    hi[d] = std::min(ohi[d] + pad, reach) + 1e-6 * E;
    if (!(hi[d] > lo[d]))
      hi[d] = lo[d] + 1e-3 * (E + 1.0);
    ext = std::max(ext, hi[d] - lo[d]);
  }
  double h = std::cbrt((hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]) / (16.0 * std::max(w->nobs, 64)));
  h = std::max(h, ext / 96.0);
  const float inv_h = (float)(1.0 / h);
  const double he = 1.0 / (double)inv_h;
  size_t cells = 1;
  for (int d = 0; d < 3; d++)
  {
    w->grid_lo[d] = std::nextafter((float)lo[d], -INFINITY);
    w->grid_n[d] = std::max(1, (int)std::ceil((hi[d] - (double)w->grid_lo[d]) / he));
    cells *= (size_t)w->grid_n[d];
  }
  if (cells > (1u << 21))
    return 0;
  w->grid_inv_h = inv_h;
  const double slop = 1e-4 * he + 1e-6 * E;
  std::vector<std::vector<unsigned short>> lists(cells);
  for (int i = 0; i < w->nobs; i++)
  {
    const double thr = rsmax + std::fabs((double)w->obs[4 * (size_t)i + 3]) + m * (1.0 + 1e-6);
    int c0[3], c1[3];
    bool empty = false;
    for (int d = 0; d < 3; d++)
    {
      const double oc = w->obs[4 * (size_t)i + d];
      c0[d] = std::max((int)std::floor((oc - thr - slop - (double)w->grid_lo[d]) / he), 0);
      c1[d] = std::min((int)std::floor((oc + thr + slop - (double)w->grid_lo[d]) / he), w->grid_n[d] - 1);
      empty = empty || c0[d] > c1[d];
    }
    if (empty)
      continue;
    for (int x = c0[0]; x <= c1[0]; x++)
      for (int y = c0[1]; y <= c1[1]; y++)
        for (int z = c0[2]; z <= c1[2]; z++)
        {
          const int ci[3] = { x, y, z };
          double d2 = 0;
          for (int d = 0; d < 3; d++)
          {
            const double oc = w->obs[4 * (size_t)i + d];
            const double bl = (double)w->grid_lo[d] + ci[d] * he - slop, bh = (double)w->grid_lo[d] + (ci[d] + 1) * he + slop;
            const double t = oc < bl ? bl - oc : (oc > bh ? oc - bh : 0.0);
            d2 += t * t;
          }
          if (std::sqrt(d2) < thr)
            lists[((size_t)x * w->grid_n[1] + y) * w->grid_n[2] + z].push_back((unsigned short)i);
        }
  }
  w->grid_start.assign(cells + 1, 0);
  w->grid_items.clear();
  for (size_t c = 0; c < cells; c++)
  {
    w->grid_start[c] = (unsigned)w->grid_items.size();
    w->grid_items.insert(w->grid_items.end(), lists[c].begin(), lists[c].end());
  }
  w->grid_start[cells] = (unsigned)w->grid_items.size();
  w->grid_on = true;
  return (int)cells;
}

void orc_world_free(void* w) { delete (World*)w; }
unsigned long long orc_world_pair_tests(void* w) { return ((World*)w)->pair_tests; }
unsigned long long orc_world_state_checks(void* w) { return ((World*)w)->state_checks; }
void orc_world_reset_counters(void* w)
{
  ((World*)w)->pair_tests = 0;
  ((World*)w)->state_checks = 0;
}

void orc_check_states(void* wh, const double* q, int n, unsigned char* ok)
{
  World* w = (World*)wh;
  for (int i = 0; i < n; i++)
    ok[i] = world_check(*w, q + (size_t)i * w->dim) ? 1 : 0;
}
// mode BISECT / LINEAR; first_bad and last_conf ([n][dim]) only written in LINEAR mode (may be NULL otherwise)
void orc_check_edges(void* wh, const double* c1, const double* c2, int n, double min_distance, int mode,
                     unsigned char* ok, long long* first_bad, double* last_conf)
{
  World* w = (World*)wh;
  const int d = w->dim;
  for (int i = 0; i < n; i++)
  {
    const double* a = c1 + (size_t)i * d;
    const double* b = c2 + (size_t)i * d;
    if (mode == GC_EDGE_LINEAR)
    {
      double tmp[GC_MAX_DIM];
      double* conf = last_conf ? last_conf + (size_t)i * d : tmp;
      long long fb = -1;
      ok[i] = checkConnectionLinear(*w, a, b, min_distance, conf, &fb) ? 1 : 0;
      if (first_bad)
        first_bad[i] = fb;
    }
    else
      ok[i] = checkConnectionBisect(*w, a, b, min_distance) ? 1 : 0;
  }
}
// result[i]: 1 free, 0 collision, 255 = the reference throws invalid_argument
void orc_check_from_conf(void* wh, const double* parent, const double* child, const double* conf, int n,
                         double min_distance, unsigned char* result)
{
  World* w = (World*)wh;
  const int d = w->dim;
  for (int i = 0; i < n; i++)
  {
    int r = checkConnFromConf(*w, parent + (size_t)i * d, child + (size_t)i * d, conf + (size_t)i * d, min_distance);
    result[i] = r < 0 ? 255 : (unsigned char)r;
  }
}

// ---- samplers ------------------------------------------------------------------
void orc_sample_informed(int dim, int n, const double* lb, const double* ub, const double* f1, const double* f2,
                         const double* cost, unsigned long long seed, unsigned long long first, double* out,
                         unsigned int* attempts)
{
  for (int i = 0; i < n; i++)
  {
    PhiloxDraws d{ first + (unsigned long long)i, seed };
    const unsigned a = informed_sample(dim, lb, ub, f1 + (size_t)i * dim, f2 + (size_t)i * dim, cost[i], d,
                                       out + (size_t)i * dim);
    if (attempts)
      attempts[i] = a;
  }
}

// ---- solvers ------------------------------------------------------------------
void* orc_solver_new(int kind, int dim, void* world, const double* lb, const double* ub, const double* start,
                     const double* goal, double max_distance, double min_distance, double rewire_factor,
                     double utopia_tolerance, int use_kdtree, int extend)
{
  return new Solver(kind, dim, (World*)world, lb, ub, start, goal, max_distance, min_distance, rewire_factor,
                    utopia_tolerance, use_kdtree != 0, extend != 0);
}
void orc_solver_free(void* s) { delete (Solver*)s; }
// feeds n injected samples; stops early when the solver reports it cannot improve; returns samples consumed
int orc_solver_run(void* sh, const double* samples, int n)
{
  Solver* s = (Solver*)sh;
  int i = 0;
  for (; i < n; i++)
  {
    bool r = s->update(samples + (size_t)i * s->pool.dim);
    if (r && !s->can_improve_)
    {
      i++;
      break;
    }
  }
  return i;
}
int orc_solver_update(void* sh, const double* q) { return ((Solver*)sh)->update(q) ? 1 : 0; }
int orc_solver_solved(void* sh) { return ((Solver*)sh)->solved_ ? 1 : 0; }
int orc_solver_can_improve(void* sh) { return ((Solver*)sh)->can_improve_ ? 1 : 0; }
double orc_solver_cost(void* sh) { return ((Solver*)sh)->cost_; }
double orc_solver_radius(void* sh) { return ((Solver*)sh)->r_rewire_; }
double orc_solver_specific_volume(void* sh) { return ((Solver*)sh)->sampler_.specific_volume_; }
int orc_solver_num_nodes(void* sh) { return ((Solver*)sh)->pool.size(); }
int orc_solver_tree_size(void* sh) { return (int)((Solver*)sh)->start_tree_->nodes_->size_; }
int orc_solver_goal_id(void* sh) { return ((Solver*)sh)->goal_node_; }
// parents[n], pcost[n], conf[n*dim] for all created nodes (ids = creation order)
void orc_solver_dump(void* sh, int* parents, double* pcost, double* conf)
{
  Solver* s = (Solver*)sh;
  int n = s->pool.size();
  s->start_tree_->ensure(n - 1);
  for (int i = 0; i < n; i++)
  {
    parents[i] = s->start_tree_->parent[i];
    pcost[i] = s->start_tree_->pcost[i];
  }
  std::memcpy(conf, s->pool.q.data(), sizeof(double) * s->pool.q.size());
}
int orc_solver_solution(void* sh, int* ids, int cap)
{
  Solver* s = (Solver*)sh;
  if (!s->solved_)
    return 0;
  int n = 0;
  if (cap > 0)
    ids[n] = s->start_tree_->root_;
  n++;
  for (int id : s->solution_)
  {
    if (n < cap)
      ids[n] = id;
    n++;
  }
  return n;
}
// ---- AnytimeRRT: sample k of problem p comes from the informed Philox sampler with counter k * P + p ------------
struct OrcAnytime
{
  std::unique_ptr<AnytimeSolver> s;
  std::vector<double> lb, ub;
};
void* orc_anytime_new(int dim, void* world, const double* lb, const double* ub, const double* start, const double* goal,
                      double max_distance, double min_distance, double utopia_tolerance, int use_kdtree, int extend,
                      double delta, double cost_impr, double sampler_cost0)
{
  OrcAnytime* a = new OrcAnytime();
  a->lb.assign(lb, lb + dim);
  a->ub.assign(ub, ub + dim);
  a->s.reset(new AnytimeSolver(dim, (World*)world, lb, ub, start, goal, max_distance, min_distance, utopia_tolerance,
                               use_kdtree != 0, extend != 0, delta, cost_impr));
  a->s->sampler_cost0 = sampler_cost0;
  return a;
}
void orc_anytime_free(void* h) { delete (OrcAnytime*)h; }
void* orc_anytime_base(void* h) { return &((OrcAnytime*)h)->s->base; }  // for the orc_solver_* getters
int orc_anytime_solve(void* h, unsigned int max_iter, int improve_in_solve, unsigned long long seed,
                      unsigned long long p, unsigned long long P)
{
  OrcAnytime* a = (OrcAnytime*)h;
  AnytimeSolver* s = a->s.get();
  const int d = s->base.pool.dim;
  auto sample = [&](double cost, double* q) {
    PhiloxDraws draw{ s->samples_drawn * P + p, seed };
    informed_sample(d, a->lb.data(), a->ub.data(), s->start_conf.data(), s->goal_conf.data(), cost, draw, q);
    s->samples_drawn++;
  };
  return s->solve(max_iter, improve_in_solve != 0, sample) ? 1 : 0;
}
// step-wise access (the harness that drives the reference's own AnytimeRRT uses the same three calls)
int orc_anytime_rrt_update(void* h, const double* q) { return ((OrcAnytime*)h)->s->base.update(q) ? 1 : 0; }
int orc_anytime_improve_begin(void* h) { return ((OrcAnytime*)h)->s->improveBegin() ? 1 : 0; }
int orc_anytime_improve_update(void* h, const double* q) { return ((OrcAnytime*)h)->s->improveUpdate(q) ? 1 : 0; }
// out: bias, cost2beat, path_cost, cost, can_improve, solved, nodes in the round tree (0: none), samples drawn
void orc_anytime_state(void* h, double* out)
{
  AnytimeSolver* s = ((OrcAnytime*)h)->s.get();
  out[0] = s->bias_;
  out[1] = s->cost2beat_;
  out[2] = s->base.path_cost_;
  out[3] = s->base.cost_;
  out[4] = s->base.can_improve_ ? 1.0 : 0.0;
  out[5] = s->base.solved_ ? 1.0 : 0.0;
  out[6] = s->new_tree_ ? (double)s->new_tree_->nodes_->size_ : 0.0;
  out[7] = (double)s->samples_drawn;
}
int orc_anytime_log(void* h, int cap_rows, double* rows)
{
  AnytimeSolver* s = ((OrcAnytime*)h)->s.get();
  const int n = (int)(s->log.size() / 8);
  for (int i = 0; i < n && i < cap_rows; i++)
    std::memcpy(rows + (size_t)i * 8, s->log.data() + (size_t)i * 8, sizeof(double) * 8);
  return n;
}
// the solution as configurations root -> goal; returns the number of waypoints
int orc_anytime_solution(void* h, int cap, double* conf)
{
  AnytimeSolver* s = ((OrcAnytime*)h)->s.get();
  if (!s->base.solved_)
    return 0;
  const int d = s->base.pool.dim;
  int n = 0;
  if (cap > 0)
    std::memcpy(conf, s->base.pool.conf(s->base.start_tree_->root_), sizeof(double) * d);
  n++;
  for (int id : s->base.solution_)
  {
    if (n < cap)
      std::memcpy(conf + (size_t)n * d, s->base.pool.conf(id), sizeof(double) * d);
    n++;
  }
  return n;
}

// counters: nn1, near, knn, edges, near_hits, state_checks, pair_tests
void orc_solver_counters(void* sh, unsigned long long* out)
{
  Solver* s = (Solver*)sh;
  out[0] = s->start_tree_->cnt.nn1;
  out[1] = s->start_tree_->cnt.near;
  out[2] = s->start_tree_->cnt.knn;
  out[3] = s->start_tree_->cnt.edges;
  out[4] = s->start_tree_->cnt.near_hits;
  out[5] = s->world->state_checks;
  out[6] = s->world->pair_tests;
}

}  // extern "C"
