// informed.cu -- Tree::informedExtend (src/graph_core/graph/tree.cpp:235-302) for many trees at once: the extension
// step of AnytimeRRT's improve rounds (src/graph_core/solvers/anytime_rrt.cpp:356-427, BASELINE config 5).
//
// Per query (one sample against one tree = one segment of the node store), exactly the reference's steps:
//   1. nearK (tree.cpp:761-765): the k = ceil(k_rrt * ln(N + 1)) nearest nodes in (distance, slot) order -- in seven
//      dimensions k exceeds N for every tree below ~10^4 nodes, i.e. all of them;
//   2. for each: next configuration (selectNextConfiguration, tree.cpp:110-128), cost-to-come (costToNode,
//      tree.cpp:767-789: connection costs summed from the node up to the root), heuristic = bias * distance +
//      (1 - bias) * cost-to-come; kept only when cost-to-come + distance + |goal - next| < cost2beat;
//   3. the kept candidates in (heuristic, rank of step 1) order -- the order of the reference's two multimaps;
//   4. the first candidate whose distance is below TOLERANCE or whose edge node -> next is collision free wins.
// Step 4 is sequential with an early exit in the reference; here the edges are checked in rounds of 4, 32, 256, ...
// candidates per still-undecided query (one gc_check_edges launch per round over all queries), and the first free one in
// order is taken, so the decision is the reference's while almost every query is settled by the first round.
#include <math_constants.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "edge_device.cuh"
#include "gc_common.cuh"

namespace gc
{
constexpr uint32_t kIeMaxNodes = 8192;  // nodes per tree the ranking kernel sorts in shared memory

__device__ __forceinline__ void ie_bitonic(double* key, uint32_t* val, uint32_t P)
{
  for (uint32_t k = 2; k <= P; k <<= 1)
    for (uint32_t j = k >> 1; j > 0; j >>= 1)
    {
      for (uint32_t i = threadIdx.x; i < P; i += blockDim.x)
      {
        const uint32_t ixj = i ^ j;
        if (ixj > i)
        {
          const double ka = key[i], kb = key[ixj];
          const uint32_t va = val[i], vb = val[ixj];
          const bool gt = (ka > kb) || (ka == kb && va > vb);
          if (gt == ((i & k) == 0))
          {
            key[i] = kb;
            key[ixj] = ka;
            val[i] = vb;
            val[ixj] = va;
          }
        }
      }
      __syncthreads();
    }
}

struct IeArgs
{
  const float* x32;
  const double* x64;
  const uint32_t* used;
  uint32_t cap_seg;
  const double* q;          // [n][D]
  const uint32_t* seg;      // [n]
  const double* goal;       // [n][D]
  const double* cost2beat;  // [n]
  const double* bias;       // [n]
  const unsigned long long* k;  // [n] neighbours asked for
  const int32_t* parent;    // [segments * cap_seg]
  const double* pcost;
  const double* c2n;        // optional [segments * cap_seg]: Tree::costToNode of every node, summed node -> root (the
                            // reference's order) when the node was appended; trees that only grow never change it
  double max_distance, tolerance;
  uint32_t P2;              // scratch row length (power of two >= the largest tree)
  double* cand_d;           // [n][P2] distances in nearK order
  uint32_t* cand_slot;      // [n][P2]
  uint32_t* order;          // [n][P2] positions in nearK order, by (heuristic, position)
  uint32_t* n_adm;          // [n]
};

template <int D>
__device__ __forceinline__ void ie_next_conf(const double* node, const double* q, double dist, double max_distance,
                                             double* nx)
{
#pragma unroll
  for (int i = 0; i < D; i++)
    nx[i] = (dist <= max_distance) ? q[i] : __dadd_rn(node[i], __dmul_rn(__ddiv_rn(__dsub_rn(q[i], node[i]), dist), max_distance));
}

template <int D>
__global__ void __launch_bounds__(512) ie_rank_kernel(IeArgs a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* key = reinterpret_cast<double*>(smem_raw);
  uint32_t* val = reinterpret_cast<uint32_t*>(key + a.P2);
  __shared__ double qs[D], gs[D];
  __shared__ uint32_t s_live, s_adm;
  const uint32_t qi = blockIdx.x;
  const uint32_t seg = a.seg[qi];
  const uint32_t n_s = a.used[seg];
  const size_t base = (size_t)seg * a.cap_seg;
  if (threadIdx.x < D)
  {
    qs[threadIdx.x] = a.q[(size_t)qi * D + threadIdx.x];
    gs[threadIdx.x] = a.goal[(size_t)qi * D + threadIdx.x];
  }
  if (threadIdx.x == 0)
  {
    s_live = 0;
    s_adm = 0;
  }
  __syncthreads();
  if (a.k[qi] >= (unsigned long long)n_s)
  {
    // Every live node is a candidate (k >= N: always so in seven dimensions, tree.cpp:761-765).  The reference orders
    // the kept candidates by (heuristic, position in (distance, slot) order); with all nodes taking part that is the
    // order of (heuristic, distance, slot), so the distance sort is not needed: evaluate every node, compact the kept
    // ones and sort those alone.  Positions are the slots themselves (cand_d / cand_slot indexed by slot).
    double* dd = reinterpret_cast<double*>(val + a.P2);  // distance by slot, for ties of the heuristic
    double* cd = a.cand_d + (size_t)qi * a.P2;
    uint32_t* cs = a.cand_slot + (size_t)qi * a.P2;
    const double bias = a.bias[qi], c2b = a.cost2beat[qi];
    for (uint32_t i = threadIdx.x; i < n_s; i += blockDim.x)
    {
      double d = CUDART_INF;
      if (a.x32[base + i] < CUDART_INF_F)
      {
        double node[D], nx[D];
#pragma unroll
        for (int c = 0; c < D; c++)
          node[c] = a.x64[(base + i) * D + c];
        d = norm_diff<D>(node, qs);
        ie_next_conf<D>(node, qs, d, a.max_distance, nx);
        double c2n = 0.0;  // Tree::costToNode
        if (a.c2n)
          c2n = a.c2n[base + i];
        else
          for (int32_t x = (int32_t)i; x != 0;)
          {
            const int32_t par = a.parent[base + x];
            if (par < 0)
            {
              c2n = CUDART_INF;
              break;
            }
            c2n = __dadd_rn(c2n, a.pcost[base + x]);
            x = par;
          }
        const double heur = __dadd_rn(__dmul_rn(bias, d), __dmul_rn(__dsub_rn(1.0, bias), c2n));
        if (__dadd_rn(__dadd_rn(c2n, d), norm_diff<D>(gs, nx)) < c2b)
        {
          const uint32_t pos = atomicAdd(&s_adm, 1u);
          key[pos] = heur;
          val[pos] = i;
        }
      }
      dd[i] = d;
      cd[i] = d;
      cs[i] = i;
    }
    __syncthreads();
    const uint32_t na = s_adm;
    uint32_t q2 = 2;
    while (q2 < na)
      q2 <<= 1;
    for (uint32_t i = na + threadIdx.x; i < q2; i += blockDim.x)
    {
      key[i] = CUDART_INF;
      val[i] = GC_NO_NODE;
    }
    __syncthreads();
    if (na > 1)
      for (uint32_t k = 2; k <= q2; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1)
        {
          for (uint32_t i = threadIdx.x; i < q2; i += blockDim.x)
          {
            const uint32_t ixj = i ^ j;
            if (ixj > i)
            {
              const double ka = key[i], kb = key[ixj];
              const uint32_t va = val[i], vb = val[ixj];
              bool gt = ka > kb;
              if (ka == kb)
              {
                const double da = (va < n_s) ? dd[va] : CUDART_INF, db = (vb < n_s) ? dd[vb] : CUDART_INF;
                gt = (da > db) || (da == db && va > vb);
              }
              if (gt == ((i & k) == 0))
              {
                key[i] = kb;
                key[ixj] = ka;
                val[i] = vb;
                val[ixj] = va;
              }
            }
          }
          __syncthreads();
        }
    uint32_t* od = a.order + (size_t)qi * a.P2;
    for (uint32_t j = threadIdx.x; j < na; j += blockDim.x)
      od[j] = val[j];
    if (threadIdx.x == 0)
      a.n_adm[qi] = na;
    return;
  }
  uint32_t p2 = 2;
  while (p2 < n_s)
    p2 <<= 1;
  uint32_t my_live = 0;
  for (uint32_t i = threadIdx.x; i < p2; i += blockDim.x)
  {
    double d = CUDART_INF;
    uint32_t v = GC_NO_NODE;
    if (i < n_s && a.x32[base + i] < CUDART_INF_F)  // row 0 of the fp32 SoA carries the tombstone mark
    {
      d = norm_diff<D>(a.x64 + (base + i) * D, qs);
      v = i;
      my_live++;
    }
    key[i] = d;
    val[i] = v;
  }
  if (my_live)
    atomicAdd(&s_live, my_live);
  __syncthreads();
  ie_bitonic(key, val, p2);
  const uint32_t live = s_live;
  const unsigned long long kq = a.k[qi];
  const uint32_t kk = (kq < (unsigned long long)live) ? (uint32_t)kq : live;
  double* cd = a.cand_d + (size_t)qi * a.P2;
  uint32_t* cs = a.cand_slot + (size_t)qi * a.P2;
  const double bias = a.bias[qi], c2b = a.cost2beat[qi];
  // positions [0, kk): the nearK multimap in order; evaluate, keep, and key by the heuristic
  double hk[16];
  uint32_t cnt = 0;  // this thread handles positions threadIdx.x + m * blockDim.x
  for (uint32_t r = threadIdx.x; r < p2; r += blockDim.x, cnt++)
  {
    double h = CUDART_INF;
    if (r < kk)
    {
      const uint32_t slot = val[r];
      const double d = key[r];
      cd[r] = d;
      cs[r] = slot;
      double node[D], nx[D];
#pragma unroll
      for (int i = 0; i < D; i++)
        node[i] = a.x64[(base + slot) * D + i];
      ie_next_conf<D>(node, qs, d, a.max_distance, nx);
      double c2n = 0.0;  // Tree::costToNode
      for (int32_t x = (int32_t)slot; x != 0;)
      {
        const int32_t par = a.parent[base + x];
        if (par < 0)
        {
          c2n = CUDART_INF;
          break;
        }
        c2n = __dadd_rn(c2n, a.pcost[base + x]);
        x = par;
      }
      const double heur = __dadd_rn(__dmul_rn(bias, d), __dmul_rn(__dsub_rn(1.0, bias), c2n));
      if (__dadd_rn(__dadd_rn(c2n, d), norm_diff<D>(gs, nx)) < c2b)
        h = heur;
    }
    hk[cnt & 15] = h;
  }
  __syncthreads();  // every thread has read key/val
  cnt = 0;
  uint32_t my_adm = 0;
  for (uint32_t r = threadIdx.x; r < p2; r += blockDim.x, cnt++)
  {
    const double h = hk[cnt & 15];
    key[r] = h;
    val[r] = r;
    my_adm += (h < CUDART_INF) ? 1u : 0u;
  }
  if (my_adm)
    atomicAdd(&s_adm, my_adm);
  __syncthreads();
  ie_bitonic(key, val, p2);
  const uint32_t na = s_adm;
  uint32_t* od = a.order + (size_t)qi * a.P2;
  for (uint32_t j = threadIdx.x; j < na; j += blockDim.x)
    od[j] = val[j];
  if (threadIdx.x == 0)
    a.n_adm[qi] = na;
}

struct IeRound
{
  IeArgs a;
  const uint32_t* unres;  // [U] queries still undecided
  uint32_t U, lo, width;
  double* c1;             // [U * width][D]
  double* c2;
  int32_t* edge_of;       // [U][width]  >= 0 edge index, -1 none, -2 free without a check (distance < TOLERANCE)
  uint32_t* ctr;          // [0] edges of this round, [1] queries undecided after it
  const uint8_t* ok_edge;
  uint32_t* next_unres;
  // results
  uint8_t* ok;
  uint32_t* node;
  double* new_conf;
  double* dist;
  double* edge_cost;
};

template <int D>
__global__ void ie_edges_kernel(IeRound r)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= r.U * r.width)
    return;
  const uint32_t u = t / r.width, jj = t % r.width;
  const uint32_t qi = r.unres[u];
  const uint32_t j = r.lo + jj;
  int32_t e = -1;
  if (j < r.a.n_adm[qi])
  {
    const uint32_t pos = r.a.order[(size_t)qi * r.a.P2 + j];
    const double d = r.a.cand_d[(size_t)qi * r.a.P2 + pos];
    if (d < r.a.tolerance)
      e = -2;
    else
    {
      const uint32_t slot = r.a.cand_slot[(size_t)qi * r.a.P2 + pos];
      const size_t base = (size_t)r.a.seg[qi] * r.a.cap_seg;
      double node[D], qv[D], nx[D];
#pragma unroll
      for (int i = 0; i < D; i++)
      {
        node[i] = r.a.x64[(base + slot) * D + i];
        qv[i] = r.a.q[(size_t)qi * D + i];
      }
      ie_next_conf<D>(node, qv, d, r.a.max_distance, nx);
      e = (int32_t)atomicAdd(&r.ctr[0], 1u);
#pragma unroll
      for (int i = 0; i < D; i++)
      {
        r.c1[(size_t)e * D + i] = node[i];
        r.c2[(size_t)e * D + i] = nx[i];
      }
    }
  }
  r.edge_of[t] = e;
}

template <int D>
__global__ void ie_pick_kernel(IeRound r)
{
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= r.U)
    return;
  const uint32_t qi = r.unres[u];
  const uint32_t na = r.a.n_adm[qi];
  int win = -1;
  for (uint32_t jj = 0; jj < r.width && r.lo + jj < na; jj++)
  {
    const int32_t e = r.edge_of[(size_t)u * r.width + jj];
    if (e == -2 || (e >= 0 && r.ok_edge[e]))
    {
      win = (int)jj;
      break;
    }
  }
  if (win >= 0)
  {
    const uint32_t pos = r.a.order[(size_t)qi * r.a.P2 + r.lo + (uint32_t)win];
    const double d = r.a.cand_d[(size_t)qi * r.a.P2 + pos];
    const uint32_t slot = r.a.cand_slot[(size_t)qi * r.a.P2 + pos];
    const size_t base = (size_t)r.a.seg[qi] * r.a.cap_seg;
    double node[D], qv[D], nx[D];
#pragma unroll
    for (int i = 0; i < D; i++)
    {
      node[i] = r.a.x64[(base + slot) * D + i];
      qv[i] = r.a.q[(size_t)qi * D + i];
    }
    ie_next_conf<D>(node, qv, d, r.a.max_distance, nx);
    r.ok[qi] = 1;
    r.node[qi] = slot;
    r.dist[qi] = d;
    r.edge_cost[qi] = norm_diff<D>(node, nx);  // extendOnly: metrics_->cost(closest_node, new_node) (tree.cpp:132)
#pragma unroll
    for (int i = 0; i < D; i++)
      r.new_conf[(size_t)qi * D + i] = nx[i];
  }
  else if (r.lo + r.width < na)
    r.next_unres[atomicAdd(&r.ctr[1], 1u)] = qi;  // more candidates to try
  else
  {
    r.ok[qi] = 0;  // every admissible candidate collides (or there was none)
    r.node[qi] = GC_NO_NODE;
  }
}

#define GC_IE_DISPATCH(dim, ...)                           \
  switch (dim)                                             \
  {                                                        \
    case 1: { constexpr int DD = 1; __VA_ARGS__; } break;  \
    case 2: { constexpr int DD = 2; __VA_ARGS__; } break;  \
    case 3: { constexpr int DD = 3; __VA_ARGS__; } break;  \
    case 4: { constexpr int DD = 4; __VA_ARGS__; } break;  \
    case 5: { constexpr int DD = 5; __VA_ARGS__; } break;  \
    case 6: { constexpr int DD = 6; __VA_ARGS__; } break;  \
    case 7: { constexpr int DD = 7; __VA_ARGS__; } break;  \
    case 8: { constexpr int DD = 8; __VA_ARGS__; } break;  \
    case 9: { constexpr int DD = 9; __VA_ARGS__; } break;  \
    case 10: { constexpr int DD = 10; __VA_ARGS__; } break; \
    case 11: { constexpr int DD = 11; __VA_ARGS__; } break; \
    case 12: { constexpr int DD = 12; __VA_ARGS__; } break; \
    case 13: { constexpr int DD = 13; __VA_ARGS__; } break; \
    case 14: { constexpr int DD = 14; __VA_ARGS__; } break; \
    case 15: { constexpr int DD = 15; __VA_ARGS__; } break; \
    case 16: { constexpr int DD = 16; __VA_ARGS__; } break; \
    default: gc::set_error("dimension %d not supported", dim); return GC_E_INVALID; \
  }

__global__ void ie_iota_kernel(uint32_t* v, uint32_t n)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    v[i] = i;
}
}  // namespace gc

using namespace gc;

namespace gc
{
// DevArena (gc_common.cuh): cached device blocks, handed out best-fit and returned by reset()
DevArena::~DevArena()
{
  for (Block& b : blocks)
    cudaFree(b.p);
}
cudaError_t DevArena::get_bytes(void** out, size_t bytes)
{
  bytes = std::max<size_t>(align_up(bytes, 256), 256);
  int best = -1;
  for (size_t i = 0; i < blocks.size(); i++)
    if (!blocks[i].in_use && blocks[i].bytes >= bytes && (best < 0 || blocks[i].bytes < blocks[(size_t)best].bytes))
      best = (int)i;
  if (best >= 0)
  {
    blocks[(size_t)best].in_use = true;
    *out = blocks[(size_t)best].p;
    return cudaSuccess;
  }
  void* p = nullptr;
  // grow with head-room so that slowly growing requests (tree sizes) do not reallocate every call
  const size_t want = bytes + bytes / 2;
  cudaError_t e = cudaMalloc(&p, want);
  size_t got = want;
  if (e != cudaSuccess)
  {
    cudaGetLastError();
    e = cudaMalloc(&p, bytes);
    got = bytes;
  }
  if (e != cudaSuccess)
    return e;
  blocks.push_back(Block{ p, got, true });
  *out = p;
  return cudaSuccess;
}
void DevArena::reset()
{
  for (Block& b : blocks)
    b.in_use = false;
}

// Tree::informedExtend for n queries, every input and output in device memory (see gc_informed_extend_batch for the
// contract).  h_seg is the host copy of d_seg (k per query is derived from the host mirror of the store).  The
// stream is synchronised between the candidate rounds (the number of undecided queries decides the next launch).
int informed_extend_core(gc_store* s, gc_world* w, const double* d_q, const uint32_t* d_seg, const uint32_t* h_seg,
                         const double* d_goal, const double* d_c2b, const double* d_bias, uint32_t n, double k_rrt,
                         double max_distance, double min_distance, double tolerance, const int32_t* d_parent,
                         const double* d_pcost, DevArena& arena, IeOut* out, uint64_t* edges_checked,
                         const double* d_c2n)
{
  const int D = s->dim;
  cudaStream_t st = s->stream;
  uint32_t mu = 1;
  std::vector<unsigned long long> kq(n);
  for (uint32_t i = 0; i < n; i++)
  {
    GC_REQUIRE(h_seg[i] < s->nseg, "gc_informed_extend_batch: query %u: segment %u of %u", i, h_seg[i], s->nseg);
    mu = std::max(mu, s->used[h_seg[i]]);
    // Tree::nearK (tree.cpp:761-765): k = ceil(k_rrt * ln(size + 1)), size = live nodes
    const double kd = (k_rrt > 0.0) ? std::ceil(k_rrt * std::log((double)s->live[h_seg[i]] + 1.0)) : 1.0e18;
    kq[i] = (kd >= 1.0e18) ? ~0ull : (unsigned long long)kd;
  }
  GC_REQUIRE(mu <= kIeMaxNodes, "gc_informed_extend_batch: trees of more than %u nodes are not supported (%u)",
             kIeMaxNodes, mu);
  uint32_t P2 = 2;
  while (P2 < mu)
    P2 <<= 1;
  IeArgs a{};
  unsigned long long* d_k;
  GC_CUDA(arena.get(&d_k, n));
  GC_CUDA(arena.get(&a.cand_d, (size_t)n * P2));
  GC_CUDA(arena.get(&a.cand_slot, (size_t)n * P2));
  GC_CUDA(arena.get(&a.order, (size_t)n * P2));
  GC_CUDA(arena.get(&a.n_adm, n));
  GC_CUDA(cudaMemcpyAsync(d_k, kq.data(), sizeof(unsigned long long) * n, cudaMemcpyHostToDevice, st));
  a.x32 = s->x32;
  a.x64 = s->x64;
  a.used = s->d_used;
  a.cap_seg = s->cap_seg;
  a.q = d_q;
  a.seg = d_seg;
  a.goal = d_goal;
  a.cost2beat = d_c2b;
  a.bias = d_bias;
  a.k = d_k;
  a.parent = d_parent;
  a.pcost = d_pcost;
  a.c2n = d_c2n;
  a.max_distance = max_distance;
  a.tolerance = tolerance;
  a.P2 = P2;
  const size_t smem = (size_t)P2 * (2 * sizeof(double) + sizeof(uint32_t));  // key | val | distance by slot
  GC_REQUIRE(P2 <= 16u * 512u, "gc_informed_extend_batch: internal: per-thread heuristic buffer too small");
  GC_IE_DISPATCH(D, if (smem > 48 * 1024) GC_CUDA(cudaFuncSetAttribute(
                        ie_rank_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                 (ie_rank_kernel<DD><<<n, 512, smem, st>>>(a)));
  GC_CUDA(cudaGetLastError());
  count_launch();
  // results
  uint8_t* d_ok;
  uint32_t* d_node;
  double *d_new, *d_dist, *d_ecost;
  uint32_t *d_unres[2], *d_ctr;
  GC_CUDA(arena.get(&d_ok, n));
  GC_CUDA(arena.get(&d_node, n));
  GC_CUDA(arena.get(&d_new, (size_t)n * D));
  GC_CUDA(arena.get(&d_dist, n));
  GC_CUDA(arena.get(&d_ecost, n));
  GC_CUDA(arena.get(&d_unres[0], n));
  GC_CUDA(arena.get(&d_unres[1], n));
  GC_CUDA(arena.get(&d_ctr, 4));
  GC_CUDA(cudaMemsetAsync(d_ok, 0, n, st));
  GC_CUDA(cudaMemsetAsync(d_dist, 0, sizeof(double) * n, st));
  GC_CUDA(cudaMemsetAsync(d_ecost, 0, sizeof(double) * n, st));
  GC_CUDA(cudaMemsetAsync(d_new, 0, sizeof(double) * (size_t)n * D, st));
  ie_iota_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_unres[0], n);
  GC_CUDA(cudaGetLastError());
  count_launch();
  // (kq is pageable memory: cudaMemcpyAsync has staged it before returning, no synchronisation needed)
  uint32_t U = n, lo = 0, width = 4;
  int cur = 0;
  uint64_t edges_total = 0;
  // states of an edge of length max_distance (grid sizing of the check kernels only)
  uint64_t per_edge = 2;
  for (double nn = 2; max_distance > nn * min_distance && per_edge < (1u << 20); nn *= 2)
    per_edge += (uint64_t)(nn / 2);
  while (U > 0)
  {
    const size_t ne = (size_t)U * width;
    IeRound r{};
    r.a = a;
    r.unres = d_unres[cur];
    r.next_unres = d_unres[cur ^ 1];
    r.U = U;
    r.lo = lo;
    r.width = width;
    uint8_t* d_oke;
    void* round_blocks[4];
    GC_CUDA(arena.get(&r.c1, ne * D));
    GC_CUDA(arena.get(&r.c2, ne * D));
    GC_CUDA(arena.get(&r.edge_of, ne));
    GC_CUDA(arena.get(&d_oke, ne));
    round_blocks[0] = r.c1;
    round_blocks[1] = r.c2;
    round_blocks[2] = r.edge_of;
    round_blocks[3] = d_oke;
    r.ctr = d_ctr;
    r.ok_edge = d_oke;
    r.ok = d_ok;
    r.node = d_node;
    r.new_conf = d_new;
    r.dist = d_dist;
    r.edge_cost = d_ecost;
    GC_CUDA(cudaMemsetAsync(d_ctr, 0, sizeof(uint32_t) * 4, st));
    GC_IE_DISPATCH(D, (ie_edges_kernel<DD><<<(uint32_t)((ne + 255) / 256), 256, 0, st>>>(r)));
    GC_CUDA(cudaGetLastError());
    count_launch();
    uint32_t hctr[2] = { 0, 0 };
    GC_CUDA(cudaMemcpyAsync(hctr, d_ctr, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    GC_CUDA(cudaStreamSynchronize(st));
    if (hctr[0] > 0)
    {
      gc_world_set_stream(w, st);
      GC_TRY(check_edges_dev(w, r.c1, r.c2, nullptr, hctr[0], min_distance, GC_EDGE_BISECT, d_oke, nullptr,
                             per_edge * hctr[0], st));
      edges_total += hctr[0];
    }
    GC_IE_DISPATCH(D, (ie_pick_kernel<DD><<<(U + 127) / 128, 128, 0, st>>>(r)));
    GC_CUDA(cudaGetLastError());
    count_launch();
    GC_CUDA(cudaMemcpyAsync(hctr, d_ctr, sizeof(uint32_t) * 2, cudaMemcpyDeviceToHost, st));
    GC_CUDA(cudaStreamSynchronize(st));
    for (void* b : round_blocks)
      arena.put(b);
    U = hctr[1];
    cur ^= 1;
    lo += width;
    width = std::min<uint32_t>(width * 8u, 4096u);  // 4, 32, 256, 2048, 4096: the late rounds hold few queries
  }
  out->d_ok = d_ok;
  out->d_node = d_node;
  out->d_new = d_new;
  out->d_dist = d_dist;
  out->d_ecost = d_ecost;
  if (edges_checked)
    *edges_checked = edges_total;
  return GC_OK;
}
}  // namespace gc

extern "C" {

int gc_informed_extend_batch(gc_store_t* s, gc_world_t* w, const double* q, const uint32_t* seg, const double* goal,
                             const double* cost2beat, const double* bias, uint32_t n, double k_rrt,
                             double max_distance, double min_distance, double tolerance, const int32_t* d_parent,
                             const double* d_pcost, uint8_t* ok, uint32_t* tree_node, double* new_conf, double* distance,
                             double* edge_cost, uint64_t* edges_checked)
{
  GC_REQUIRE(s && w && q && seg && goal && cost2beat && bias && d_parent && d_pcost && ok && tree_node && new_conf,
             "gc_informed_extend_batch: NULL argument");
  GC_REQUIRE(s->dim == w->dim && s->device == w->device, "gc_informed_extend_batch: store and world do not match");
  if (edges_checked)
    *edges_checked = 0;
  if (n == 0)
    return GC_OK;
  GC_SET_DEVICE(s->device);
  const int D = s->dim;
  cudaStream_t st = s->stream;
  GC_CUDA(cudaStreamSynchronize(st));
  DevArena arena;  // freed on return
  double *d_q, *d_goal, *d_c2b, *d_bias;
  uint32_t* d_seg;
  GC_CUDA(arena.get(&d_q, (size_t)n * D));
  GC_CUDA(arena.get(&d_goal, (size_t)n * D));
  GC_CUDA(arena.get(&d_c2b, n));
  GC_CUDA(arena.get(&d_bias, n));
  GC_CUDA(arena.get(&d_seg, n));
  GC_CUDA(cudaMemcpyAsync(d_q, q, sizeof(double) * (size_t)n * D, cudaMemcpyHostToDevice, st));
  GC_CUDA(cudaMemcpyAsync(d_goal, goal, sizeof(double) * (size_t)n * D, cudaMemcpyHostToDevice, st));
  GC_CUDA(cudaMemcpyAsync(d_c2b, cost2beat, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  GC_CUDA(cudaMemcpyAsync(d_bias, bias, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  GC_CUDA(cudaMemcpyAsync(d_seg, seg, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
  IeOut o{};
  GC_TRY(informed_extend_core(s, w, d_q, d_seg, seg, d_goal, d_c2b, d_bias, n, k_rrt, max_distance, min_distance,
                              tolerance, d_parent, d_pcost, arena, &o, edges_checked, nullptr));
  GC_CUDA(cudaMemcpyAsync(ok, o.d_ok, n, cudaMemcpyDeviceToHost, st));
  GC_CUDA(cudaMemcpyAsync(tree_node, o.d_node, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
  GC_CUDA(cudaMemcpyAsync(new_conf, o.d_new, sizeof(double) * (size_t)n * D, cudaMemcpyDeviceToHost, st));
  if (distance)
    GC_CUDA(cudaMemcpyAsync(distance, o.d_dist, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  if (edge_cost)
    GC_CUDA(cudaMemcpyAsync(edge_cost, o.d_ecost, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  GC_CUDA(cudaStreamSynchronize(st));
  return GC_OK;
}

}  // extern "C"
