// sharded.cu -- a node-sharded tree behind the C ABI (gc_comm_init, gc_store_create_sharded):
// ONE process drives the GPUs of a box, as a graph_core planner does.  Node i lives on shard i mod G at local slot
// i div G (round robin keeps the shards balanced under incremental insertion and keeps "lowest global slot" recoverable
// as local * G + shard).  A query batch is broadcast to the shards (one H2D copy per device), every shard answers on its
// slice with the single-store kernels of nn.cu on its own stream, and the exchange is fused into the merge: the merge
// kernel on shard 0's device loads the G per-shard lists straight from the other GPUs through peer-mapped pointers
// (NVLink / NVSwitch loads, cudaDeviceEnablePeerAccess) -- no collective library, no gathered copy.  Results are ordered
// by (fp64 distance, global slot), i.e. exactly what one Vector / KdTree over all the nodes returns.
// Several shards may sit on the same device (testing on one GPU; oversubscription).
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "gc_common.cuh"

namespace gc
{
__global__ void nn1_counts_kernel(const uint32_t* __restrict__ idx, uint32_t nq, uint32_t* __restrict__ count)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nq)
    count[i] = (idx[i] != GC_NO_NODE) ? 1u : 0u;
}
}  // namespace gc

using namespace gc;

struct gc_comm
{
  std::vector<int> devices;
};

// One persistent host thread per shard (beyond shard 0, which the caller's thread serves): the per-shard query calls
// enqueue several launches each and the k-nearest path ends with a one-word read-back, so issuing them from one thread
// would serialise the shards.
struct ShardWorker
{
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  std::function<int()> job;
  bool has = false, stop = false, done = true;
  int rc = 0;
  std::string err;
  ShardWorker() { th = std::thread([this] { run(); }); }
  ~ShardWorker()
  {
    {
      std::lock_guard<std::mutex> lk(m);
      stop = true;
    }
    cv.notify_all();
    th.join();
  }
  void run()
  {
    for (;;)
    {
      std::unique_lock<std::mutex> lk(m);
      cv.wait(lk, [&] { return has || stop; });
      if (stop)
        return;
      std::function<int()> j = std::move(job);
      has = false;
      lk.unlock();
      const int r = j();
      std::string e = (r < 0) ? std::string(gc_last_error()) : std::string();
      lk.lock();
      rc = r;
      err = std::move(e);
      done = true;
      cv.notify_all();
    }
  }
  void submit(std::function<int()> f)
  {
    {
      std::lock_guard<std::mutex> lk(m);
      job = std::move(f);
      has = true;
      done = false;
    }
    cv.notify_all();
  }
  int wait()
  {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [&] { return done; });
    return rc;
  }
};

struct gc_shard
{
  int device = 0;
  gc_store_t* store = nullptr;
  cudaEvent_t done = nullptr;
  gc::DevBuf d_q, d_radius, d_idx, d_dist, d_count;
  std::unique_ptr<ShardWorker> worker;  // shards 1 .. G-1
};

struct gc_sharded_store
{
  int dim = 0;
  uint64_t n_global = 0;
  uint64_t capacity = 0;
  std::vector<gc_shard> shards;
  // root = shard 0's device
  cudaStream_t root_stream = nullptr;
  gc::DevBuf d_out_idx, d_out_dist, d_out_count;
  gc::PinBuf h_q, h_out, h_cnt;
  std::vector<double> rows;  // append staging
};

extern "C" {

int gc_comm_init(int ndev, const int* devices, gc_comm_t** out)
{
  GC_REQUIRE(out != nullptr, "gc_comm_init: out is NULL");
  *out = nullptr;
  GC_REQUIRE(ndev >= 1 && ndev <= 16, "gc_comm_init: %d shards (1..16)", ndev);
  int have = 0;
  GC_TRY(gc_device_count(&have));
  GC_REQUIRE(have > 0, "gc_comm_init: no CUDA device (no CPU fallback exists)");
  gc_comm* c = new gc_comm();
  for (int i = 0; i < ndev; i++)
  {
    const int d = devices ? devices[i] : i;
    if (d < 0 || d >= have)
    {
      gc::set_error("gc_comm_init: device %d of %d", d, have);
      delete c;
      return GC_E_INVALID;
    }
    c->devices.push_back(d);
  }
  // the merge kernel on shard 0's device reads every other shard's result lists: peer access from devices[0]
  gc::CtxGuard guard;
  const int root = c->devices[0];
  for (int d : c->devices)
  {
    if (d == root)
      continue;
    int can = 0;
    cudaError_t e = cudaDeviceCanAccessPeer(&can, root, d);
    if (e != cudaSuccess || !can)
    {
      gc::set_error("gc_comm_init: device %d cannot access device %d as a peer", root, d);
      delete c;
      return GC_E_UNSUPPORTED;
    }
    e = cudaSetDevice(root);
    if (e == cudaSuccess)
      e = cudaDeviceEnablePeerAccess(d, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled)
    {
      cudaGetLastError();
      e = cudaSuccess;
    }
    if (e != cudaSuccess)
    {
      gc::set_error("gc_comm_init: cudaDeviceEnablePeerAccess(%d -> %d): %s", root, d, cudaGetErrorString(e));
      delete c;
      return GC_E_CUDA;
    }
  }
  *out = c;
  return GC_OK;
}

int gc_comm_size(gc_comm_t* c) { return c ? (int)c->devices.size() : GC_E_INVALID; }

int gc_comm_destroy(gc_comm_t* c)
{
  delete c;
  return GC_OK;
}

int gc_sharded_store_destroy(gc_sharded_store_t* s)
{
  if (!s)
    return GC_OK;
  gc::CtxGuard guard;
  for (gc_shard& sh : s->shards)
    sh.worker.reset();
  for (gc_shard& sh : s->shards)
  {
    cudaSetDevice(sh.device);
    if (sh.store)
      gc_store_destroy(sh.store);
    if (sh.done)
      cudaEventDestroy(sh.done);
    sh.d_q.release();
    sh.d_radius.release();
    sh.d_idx.release();
    sh.d_dist.release();
    sh.d_count.release();
  }
  if (!s->shards.empty())
    cudaSetDevice(s->shards[0].device);
  s->d_out_idx.release();
  s->d_out_dist.release();
  s->d_out_count.release();
  s->h_q.release();
  s->h_out.release();
  s->h_cnt.release();
  delete s;
  return GC_OK;
}

int gc_store_create_sharded(gc_comm_t* c, int dim, uint64_t capacity_total, gc_sharded_store_t** out)
{
  GC_REQUIRE(c && out, "gc_store_create_sharded: NULL argument");
  *out = nullptr;
  GC_REQUIRE(capacity_total > 0, "gc_store_create_sharded: empty capacity");
  const size_t G = c->devices.size();
  const uint64_t per = (capacity_total + G - 1) / G + 1;
  GC_REQUIRE(per < 0xFFFFFFF0ull / G, "gc_store_create_sharded: %llu nodes exceed the 32-bit global slot space",
             (unsigned long long)capacity_total);
  gc::CtxGuard guard;
  gc_sharded_store* s = new gc_sharded_store();
  s->dim = dim;
  s->capacity = capacity_total;
  s->shards.resize(G);
  for (size_t g = 0; g < G; g++)
  {
    gc_shard& sh = s->shards[g];
    sh.device = c->devices[g];
    int rc = gc_store_create(dim, (uint32_t)per, 1, sh.device, &sh.store);
    if (rc == GC_OK && (cudaSetDevice(sh.device) != cudaSuccess ||
                        cudaEventCreateWithFlags(&sh.done, cudaEventDisableTiming) != cudaSuccess))
    {
      gc::set_error("gc_store_create_sharded: event creation failed on device %d", sh.device);
      rc = GC_E_CUDA;
    }
    if (rc != GC_OK)
    {
      gc_sharded_store_destroy(s);
      return rc;
    }
  }
  for (size_t g = 1; g < G; g++)
    s->shards[g].worker.reset(new ShardWorker());
  s->root_stream = s->shards[0].store->stream;
  *out = s;
  return GC_OK;
}

int gc_sharded_store_size(gc_sharded_store_t* s, uint64_t* nodes, uint32_t* shards)
{
  GC_REQUIRE(s != nullptr, "gc_sharded_store_size: NULL store");
  if (nodes)
    *nodes = s->n_global;
  if (shards)
    *shards = (uint32_t)s->shards.size();
  return GC_OK;
}

// insert (NearestNeighbors::insert, kdtree.cpp:327-352 / vector.cpp:38-43) of n nodes: global slots first .. first+n-1
int gc_sharded_store_append(gc_sharded_store_t* s, const double* q, uint32_t n, uint64_t* first_global_slot)
{
  GC_REQUIRE(s && q, "gc_sharded_store_append: NULL argument");
  GC_REQUIRE(s->n_global + n <= s->capacity, "gc_sharded_store_append: %llu + %u nodes exceed the capacity %llu",
             (unsigned long long)s->n_global, n, (unsigned long long)s->capacity);
  if (first_global_slot)
    *first_global_slot = s->n_global;
  const size_t G = s->shards.size();
  const int D = s->dim;
  for (size_t g = 0; g < G; g++)
  {
    // rows i with (n_global + i) mod G == g, in order: their local slots are consecutive
    s->rows.clear();
    const uint64_t r0 = (g + G - s->n_global % G) % G;
    for (uint64_t i = r0; i < n; i += G)
      s->rows.insert(s->rows.end(), q + i * D, q + (i + 1) * D);
    const uint32_t cnt = (uint32_t)(s->rows.size() / D);
    if (cnt == 0)
      continue;
    uint32_t first = 0;
    GC_TRY(gc_store_append(s->shards[g].store, 0, s->rows.data(), cnt, &first));
    if ((uint64_t)first != (s->n_global + r0) / G)
    {
      gc::set_error("gc_sharded_store_append: shard %zu is out of step (local slot %u)", g, first);
      return GC_E_INVALID;
    }
  }
  s->n_global += n;
  return GC_OK;
}

// mode 0 nearest, 1 k-nearest, 2 radius
static int sharded_query(gc_sharded_store* s, int mode, const double* q, uint32_t nq, uint64_t k, const double* radius,
                         uint32_t cap, uint32_t* idx, double* dist, uint32_t* count)
{
  if (nq == 0)
    return GC_OK;
  gc::CtxGuard guard;
  const uint32_t G = (uint32_t)s->shards.size();
  const int D = s->dim;
  const uint32_t cap_shard = (mode == 0) ? 1u : ((mode == 1) ? (uint32_t)std::min<uint64_t>(k, cap) : cap);
  const size_t qb = sizeof(double) * (size_t)nq * D;
  GC_CUDA(cudaSetDevice(s->shards[0].device));
  GC_TRY(s->h_q.reserve(qb + sizeof(double) * nq));
  std::memcpy(s->h_q.p, q, qb);
  double* h_rad = reinterpret_cast<double*>(static_cast<char*>(s->h_q.p) + qb);
  if (mode == 2)
    std::memcpy(h_rad, radius, sizeof(double) * nq);
  const void* idx_p[16];
  const void* dist_p[16];
  const void* cnt_p[16];
  auto shard_job = [=](uint32_t g) -> int {
    gc_shard& sh = s->shards[g];
    gc::CtxGuard job_guard;
    GC_CUDA(cudaSetDevice(sh.device));
    cudaStream_t st = sh.store->stream;
    GC_TRY(sh.d_q.reserve(qb));
    GC_TRY(sh.d_idx.reserve(sizeof(uint32_t) * (size_t)nq * cap_shard));
    GC_TRY(sh.d_dist.reserve(sizeof(double) * (size_t)nq * cap_shard));
    GC_TRY(sh.d_count.reserve(sizeof(uint32_t) * nq));
    GC_CUDA(cudaMemcpyAsync(sh.d_q.p, s->h_q.p, qb, cudaMemcpyHostToDevice, st));
    if (mode == 0)
    {
      GC_TRY(gc_nn1_dev(sh.store, sh.d_q.as<double>(), nullptr, nq, sh.d_idx.as<uint32_t>(), sh.d_dist.as<double>()));
      nn1_counts_kernel<<<(nq + 255) / 256, 256, 0, st>>>(sh.d_idx.as<uint32_t>(), nq, sh.d_count.as<uint32_t>());
      GC_CUDA(cudaGetLastError());
      count_launch();
    }
    else if (mode == 1)
      GC_TRY(gc_knn_dev(sh.store, sh.d_q.as<double>(), nullptr, nq, k, cap_shard, sh.d_idx.as<uint32_t>(),
                        sh.d_dist.as<double>(), sh.d_count.as<uint32_t>()));
    else
    {
      GC_TRY(sh.d_radius.reserve(sizeof(double) * nq));
      GC_CUDA(cudaMemcpyAsync(sh.d_radius.p, h_rad, sizeof(double) * nq, cudaMemcpyHostToDevice, st));
      GC_TRY(gc_radius_dev(sh.store, sh.d_q.as<double>(), nullptr, sh.d_radius.as<double>(), nq, cap_shard,
                           sh.d_idx.as<uint32_t>(), sh.d_dist.as<double>(), sh.d_count.as<uint32_t>()));
    }
    GC_CUDA(cudaEventRecord(sh.done, st));
    return GC_OK;
  };
  for (uint32_t g = 1; g < G; g++)
    s->shards[g].worker->submit([=] { return shard_job(g); });
  int rc0 = shard_job(0);
  for (uint32_t g = 1; g < G; g++)
  {
    const int r = s->shards[g].worker->wait();
    if (r < 0 && rc0 >= 0)
    {
      gc::set_error("shard %u: %s", g, s->shards[g].worker->err.c_str());
      rc0 = r;
    }
  }
  if (rc0 < 0)
    return rc0;
  for (uint32_t g = 0; g < G; g++)
  {
    idx_p[g] = s->shards[g].d_idx.p;
    dist_p[g] = s->shards[g].d_dist.p;
    cnt_p[g] = s->shards[g].d_count.p;
  }
  // merge on shard 0's device, reading the other shards' lists through peer pointers
  gc_shard& root = s->shards[0];
  GC_CUDA(cudaSetDevice(root.device));
  cudaStream_t rs = s->root_stream;
  for (uint32_t g = 1; g < G; g++)
    GC_CUDA(cudaStreamWaitEvent(rs, s->shards[g].done, 0));
  const uint64_t k_out = (mode == 0) ? 1 : ((mode == 1) ? k : (uint64_t)G * cap_shard);
  const size_t no = (size_t)nq * cap;
  GC_TRY(s->d_out_idx.reserve(sizeof(uint32_t) * no));
  GC_TRY(s->d_out_dist.reserve(sizeof(double) * no));
  GC_TRY(s->d_out_count.reserve(sizeof(uint32_t) * nq));
  GC_TRY(gc_merge_shard_lists_p2p_dev(root.device, rs, G, nq, cap_shard, idx_p, dist_p, cnt_p, k_out, cap,
                                      s->d_out_idx.as<uint32_t>(), s->d_out_dist.as<double>(),
                                      s->d_out_count.as<uint32_t>()));
  GC_TRY(s->h_out.reserve(sizeof(double) * no + sizeof(uint32_t) * no + sizeof(uint32_t) * nq));
  double* hd = s->h_out.as<double>();
  uint32_t* hi = reinterpret_cast<uint32_t*>(hd + no);
  uint32_t* hc = hi + no;
  GC_CUDA(cudaMemcpyAsync(hd, s->d_out_dist.p, sizeof(double) * no, cudaMemcpyDeviceToHost, rs));
  GC_CUDA(cudaMemcpyAsync(hi, s->d_out_idx.p, sizeof(uint32_t) * no, cudaMemcpyDeviceToHost, rs));
  GC_CUDA(cudaMemcpyAsync(hc, s->d_out_count.p, sizeof(uint32_t) * nq, cudaMemcpyDeviceToHost, rs));
  // radius: the raw per-shard counts tell whether a shard list was cut (gc_radius reports the true size)
  uint32_t* hraw = nullptr;
  if (mode == 2)
  {
    GC_TRY(s->h_cnt.reserve(sizeof(uint32_t) * (size_t)nq * G));
    hraw = s->h_cnt.as<uint32_t>();
    for (uint32_t g = 0; g < G; g++)
    {
      gc_shard& sh = s->shards[g];
      GC_CUDA(cudaSetDevice(sh.device));
      GC_CUDA(cudaMemcpyAsync(hraw + (size_t)g * nq, sh.d_count.p, sizeof(uint32_t) * nq, cudaMemcpyDeviceToHost,
                              sh.store->stream));
      GC_CUDA(cudaStreamSynchronize(sh.store->stream));
    }
    GC_CUDA(cudaSetDevice(root.device));
  }
  GC_CUDA(cudaStreamSynchronize(rs));
  int rc = GC_OK;
  for (uint32_t i = 0; i < nq; i++)
  {
    uint32_t c = hc[i];
    if (mode == 2)
    {
      uint64_t total = 0;
      bool cut = false;
      for (uint32_t g = 0; g < G; g++)
      {
        total += hraw[(size_t)g * nq + i];
        cut = cut || hraw[(size_t)g * nq + i] > cap_shard;
      }
      if (cut || total > cap)
        rc = GC_W_TRUNCATED;
      c = (uint32_t)std::min<uint64_t>(total, 0xFFFFFFFFull);
    }
    const uint32_t m = std::min(c, cap);
    if (idx)
      std::memcpy(idx + (size_t)i * cap, hi + (size_t)i * cap, sizeof(uint32_t) * m);
    if (dist)
      std::memcpy(dist + (size_t)i * cap, hd + (size_t)i * cap, sizeof(double) * m);
    if (count)
      count[i] = c;
  }
  return rc;
}

int gc_sharded_nn1(gc_sharded_store_t* s, const double* q, uint32_t nq, uint32_t* idx, double* dist)
{
  GC_REQUIRE(s && q && idx && dist, "gc_sharded_nn1: NULL argument");
  std::vector<uint32_t> cnt(nq);
  for (uint32_t i = 0; i < nq; i++)
  {
    idx[i] = GC_NO_NODE;
    dist[i] = std::numeric_limits<double>::infinity();
  }
  return sharded_query(s, 0, q, nq, 1, nullptr, 1, idx, dist, cnt.data());
}

int gc_sharded_knn(gc_sharded_store_t* s, const double* q, uint32_t nq, uint64_t k, uint32_t cap, uint32_t* idx,
                   double* dist, uint32_t* count)
{
  GC_REQUIRE(s && q && idx && dist && count && cap > 0, "gc_sharded_knn: NULL argument or cap = 0");
  return sharded_query(s, 1, q, nq, k, nullptr, cap, idx, dist, count);
}

int gc_sharded_radius(gc_sharded_store_t* s, const double* q, const double* radius, uint32_t nq, uint32_t cap,
                      uint32_t* idx, double* dist, uint32_t* count)
{
  GC_REQUIRE(s && q && radius && idx && dist && count && cap > 0, "gc_sharded_radius: NULL argument or cap = 0");
  return sharded_query(s, 2, q, nq, 0, radius, cap, idx, dist, count);
}

}  // extern "C"
