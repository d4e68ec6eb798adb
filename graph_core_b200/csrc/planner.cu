// planner.cu -- fused device part of Tree::extend for lock-step trees, and the Philox samplers.
//
//   Tree::findClosestNode          src/graph_core/graph/tree.cpp:56-59
//   Tree::selectNextConfiguration  src/graph_core/graph/tree.cpp:110-128  (steer)
//   Tree::tryExtendFromNode        src/graph_core/graph/tree.cpp:69-81    (edge check unless dist < TOLERANCE)
//   Tree::near                     src/graph_core/graph/tree.cpp:751-754
//   UniformSampler::sample         src/graph_core/samplers/uniform_sampler.cpp:34-39
//
// All kernels of one gc_extend_batch call are enqueued on the store's stream without any host
// round trip in between: nearest -> steer -> edge plan/check -> radius-near -> sort.
#include <math_constants.h>

#include <cmath>

#include "gc_common.cuh"

namespace gc
{
// next = q                      when |node - q| <= max_distance   (same bits as q)
//      = node + (q - node)/dist*max_distance  otherwise          (individually rounded fp64 ops)
__global__ void steer_kernel(const double* __restrict__ x64, const double* __restrict__ q,
                             const uint32_t* __restrict__ seg, const uint32_t* __restrict__ closest,
                             const double* __restrict__ dist, uint32_t n, int dim, uint32_t cap_seg,
                             double max_distance, double* __restrict__ c1, double* __restrict__ c2)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const uint32_t slot = closest[i];
  const uint32_t sg = seg ? seg[i] : 0u;
  const double dd = dist[i];
  for (int d = 0; d < dim; d++)
  {
    const double qv = q[(size_t)i * dim + d];
    double nc = qv, nx = qv;
    if (slot != GC_NO_NODE)
    {
      nc = x64[((size_t)sg * cap_seg + slot) * dim + d];
      if (!(dd <= max_distance))
        nx = __dadd_rn(nc, __dmul_rn(__ddiv_rn(__dsub_rn(qv, nc), dd), max_distance));
    }
    c1[(size_t)i * dim + d] = nc;
    c2[(size_t)i * dim + d] = nx;
  }
}

__global__ void gather_rows_kernel(const double* __restrict__ src, const uint32_t* __restrict__ rows, uint32_t n,
                                   int dim, double* __restrict__ dst)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * (uint32_t)dim)
    dst[i] = src[(size_t)rows[i / dim] * dim + i % dim];
}

// ---- Philox4x32-10 ------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t* out)
{
#pragma unroll
  for (int r = 0; r < 10; r++)
  {
    const uint32_t hi0 = __umulhi(GC_PHILOX_M0, c0), lo0 = GC_PHILOX_M0 * c0;
    const uint32_t hi1 = __umulhi(GC_PHILOX_M1, c2), lo1 = GC_PHILOX_M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
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

__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo)
{
  const unsigned long long x = ((unsigned long long)hi << 32) | lo;
  return (double)(x >> 11) * 1.1102230246251565404e-16;  // 2^-53
}

// x_d = 0.5*(lb+ub) + (2u-1) * 0.5*(lb-ub)   (the reference's formula with Random() in [-1,1))
__global__ void sample_uniform_kernel(const double* __restrict__ lb, const double* __restrict__ ub, int dim,
                                      unsigned long long seed, unsigned long long first, uint32_t n,
                                      double* __restrict__ out)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t pairs = (uint32_t)(dim + 1) / 2;
  if (t >= n * pairs)
    return;
  const uint32_t i = t / pairs, b = t % pairs;
  const unsigned long long ctr = first + i;
  uint32_t r[4];
  philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), b, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  for (int h = 0; h < 2; h++)
  {
    const int d = 2 * (int)b + h;
    if (d < dim)
    {
      const double u = u01(r[2 * h], r[2 * h + 1]);
      const double c = __dmul_rn(0.5, __dadd_rn(lb[d], ub[d]));
      const double hw = __dmul_rn(0.5, __dsub_rn(lb[d], ub[d]));
      const double rr = __dsub_rn(__dmul_rn(2.0, u), 1.0);
      out[(size_t)i * dim + d] = __dadd_rn(c, __dmul_rn(rr, hw));
    }
  }
}

// ---- informed sampler -----------------------------------------------------------------------------------
// InformedSampler::sample (informed_sampler.cpp:145-180) for n independent problems: problem i has its own foci
// and cost.  cost = +inf -> uniform in the box (:149).  Otherwise up to 100 attempts (:154-173): a point of the
// unit ball -- direction = D uniforms in [-1,1] divided by their 2-norm (the reference's setRandom()/norm(), a
// cube-normalised direction, reproduced on purpose), radius = u^(1/D) -- is stretched by the ellipse axes
// (max_radius = cost/2 along the focal line, min_radius = sqrt(cost^2 - focii^2)/2 elsewhere, :181-207), rotated by
// the Gram-Schmidt basis of computeRotationMatrix (:107-143), moved to the centre and accepted when inside the
// bounds; after 100 rejections the uniform fallback (:178).
// Random stream of sample i: Philox4x32-10, counter (i, draw/2, 1), key = seed; draws are consumed in the order
// written above (D direction draws, then the radius draw, per attempt).  u^(1/D) is a 64-step bisection built from
// multiplications only, so that the oracle's CPU restatement produces the same bits (pow() is not portable).
__device__ __forceinline__ double informed_root(double u, int D)
{
  double lo = 0.0, hi = 1.0;
  for (int it = 0; it < 64; it++)
  {
    const double mid = __dmul_rn(0.5, __dadd_rn(lo, hi));
    double p = mid;
    for (int k = 1; k < D; k++)
      p = __dmul_rn(p, mid);
    if (p < u)
      lo = mid;
    else
      hi = mid;
  }
  return __dmul_rn(0.5, __dadd_rn(lo, hi));
}

struct InformedArgs
{
  int dim;
  uint32_t n;
  const double* lb;
  const double* ub;
  const double* f1;    // [n][dim]
  const double* f2;    // [n][dim]
  const double* cost;  // [n]
  unsigned long long seed, first;
  double* out;         // [n][dim]
  uint32_t* attempts;  // optional [n]: rejection rounds used (101 = fallback, 0 = infinite cost)
  // optional (lock-step AnytimeRRT, anytime.cu): row i works for problem prob[i] -- foci and output are rows prob[i] of
  // f1 / f2 / out -- and draws from Philox counter ctr[i] instead of first + i
  const uint32_t* prob;
  const unsigned long long* ctr;
  const double* basis;        // optional [rows][D*D + D + 1]: informed_basis_kernel's output for rows prob[i] (or i)
  uint32_t lanes_per_sample;  // 0 / 1: one thread per sample; 32: a warp per sample, its lanes share the rejection rounds
};

// rotation basis of the prolate hyperspheroid (informed_sampler.cpp:107-143): R[row][col], centre, distance of the foci
__device__ __forceinline__ void informed_basis(int D, const double* f1, const double* f2, double (*R)[GC_MAX_DIM],
                                               double* centre, double& fd)
{
  double v[GC_MAX_DIM];
  double s2 = 0.0;
  for (int d = 0; d < D; d++)
  {
    v[d] = __dsub_rn(f1[d], f2[d]);
    s2 = __dadd_rn(s2, __dmul_rn(v[d], v[d]));
    centre[d] = __dmul_rn(0.5, __dadd_rn(f1[d], f2[d]));
  }
  fd = __dsqrt_rn(s2);
  for (int d = 0; d < D; d++)
    v[d] = __ddiv_rn(v[d], fd);
  for (int r0 = 0; r0 < D; r0++)
    for (int c0 = 0; c0 < D; c0++)
      R[r0][c0] = (r0 == c0) ? 1.0 : 0.0;
  bool standard = false;
  for (int ic = 0; ic < D; ic++)
    if (fabs(v[ic]) > 0.999)
    {
      standard = true;  // swap columns ic and 0 of the identity
      for (int r0 = 0; r0 < D; r0++)
      {
        const double t = R[r0][ic];
        R[r0][ic] = R[r0][0];
        R[r0][0] = t;
      }
      break;
    }
  if (!standard)
  {
    for (int r0 = 0; r0 < D; r0++)
      R[r0][0] = v[r0];
    for (int ic = 1; ic < D; ic++)
    {
      for (int il = 0; il < ic; il++)
      {
        double dot = 0.0;
        for (int r0 = 0; r0 < D; r0++)
          dot = __dadd_rn(dot, __dmul_rn(R[r0][ic], R[r0][il]));
        for (int r0 = 0; r0 < D; r0++)
          R[r0][ic] = __dsub_rn(R[r0][ic], __dmul_rn(dot, R[r0][il]));
      }
      double n2 = 0.0;
      for (int r0 = 0; r0 < D; r0++)
        n2 = __dadd_rn(n2, __dmul_rn(R[r0][ic], R[r0][ic]));
      const double nn = __dsqrt_rn(n2);
      for (int r0 = 0; r0 < D; r0++)
        R[r0][ic] = __ddiv_rn(R[r0][ic], nn);
    }
  }
}

// the basis depends on the foci only: lock-step planners whose foci are fixed per problem compute it once
// (basis[p] = R row-major [D][D] | centre[D] | fd)
__global__ void informed_basis_kernel(int D, uint32_t n, const double* f1, const double* f2, double* basis)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  double R[GC_MAX_DIM][GC_MAX_DIM], centre[GC_MAX_DIM], fd;
  informed_basis(D, f1 + (size_t)i * D, f2 + (size_t)i * D, R, centre, fd);
  double* b = basis + (size_t)i * (D * D + D + 1);
  for (int r0 = 0; r0 < D; r0++)
    for (int c0 = 0; c0 < D; c0++)
      b[r0 * D + c0] = R[r0][c0];
  for (int d = 0; d < D; d++)
    b[D * D + d] = centre[d];
  b[D * D + D] = fd;
}

__global__ void sample_informed_kernel(InformedArgs a)
{
  const uint32_t lps = a.lanes_per_sample ? a.lanes_per_sample : 1u;  // 1 or 32
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = gid / lps, sub = gid % lps;
  if (i >= a.n)
    return;
  const int D = a.dim;
  const unsigned long long ctr = a.ctr ? a.ctr[i] : a.first + i;
  const size_t row = a.prob ? a.prob[i] : i;
  uint32_t cached_block = 0xFFFFFFFFu;
  uint32_t r[4];
  auto draw = [&](uint32_t j) {
    const uint32_t b = j >> 1;
    if (b != cached_block)
    {
      philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), b, 1u, (uint32_t)a.seed, (uint32_t)(a.seed >> 32), r);
      cached_block = b;
    }
    return (j & 1u) ? u01(r[2], r[3]) : u01(r[0], r[1]);
  };
  auto uniform = [&](uint32_t j0) {
    for (int d = 0; d < D; d++)
    {
      const double u = draw(j0 + (uint32_t)d);
      const double c = __dmul_rn(0.5, __dadd_rn(a.lb[d], a.ub[d]));
      const double hw = __dmul_rn(0.5, __dsub_rn(a.lb[d], a.ub[d]));
      a.out[row * D + d] = __dadd_rn(c, __dmul_rn(__dsub_rn(__dmul_rn(2.0, u), 1.0), hw));
    }
  };
  double cost = a.cost[i];
  if (!(cost < CUDART_INF))
  {
    uniform(0u);
    if (a.attempts)
      a.attempts[i] = 0;
    return;
  }
  const double* f1 = a.f1 + row * D;
  const double* f2 = a.f2 + row * D;
  double R[GC_MAX_DIM][GC_MAX_DIM];  // R[row][col]
  double centre[GC_MAX_DIM], axis[GC_MAX_DIM];
  double fd;
  if (a.basis)
  {
    const double* b = a.basis + row * (size_t)(D * D + D + 1);
    for (int r0 = 0; r0 < D; r0++)
      for (int c0 = 0; c0 < D; c0++)
        R[r0][c0] = b[r0 * D + c0];
    for (int d = 0; d < D; d++)
      centre[d] = b[D * D + d];
    fd = b[D * D + D];
  }
  else
    informed_basis(D, f1, f2, R, centre, fd);
  double rmin;
  if (cost < fd)
  {
    cost = fd;
    rmin = 0.0;
  }
  else
    rmin = __dmul_rn(0.5, __dsqrt_rn(__dsub_rn(__dmul_rn(cost, cost), __dmul_rn(fd, fd))));
  const double rmax = __dmul_rn(0.5, cost);
  for (int d = 0; d < D; d++)
    axis[d] = (d == 0) ? rmax : rmin;
  // rejection rounds: with lps == 32 the 32 lanes that share a sample try 32 consecutive rounds at once and the
  // first accepted round in order wins -- the same draws and arithmetic per round as the one-thread form
  for (uint32_t tb = 0; tb < 100u; tb += lps)
  {
    const uint32_t t = tb + sub;
    bool inside = false;
    double q[GC_MAX_DIM];
    if (t < 100u)
    {
      const uint32_t j0 = t * (uint32_t)(D + 1);
      double ball[GC_MAX_DIM];
      double b2 = 0.0;
      for (int d = 0; d < D; d++)
      {
        ball[d] = __dsub_rn(__dmul_rn(2.0, draw(j0 + (uint32_t)d)), 1.0);
        b2 = __dadd_rn(b2, __dmul_rn(ball[d], ball[d]));
      }
      const double scale = __ddiv_rn(informed_root(draw(j0 + (uint32_t)D), D), __dsqrt_rn(b2));
      inside = true;
      for (int r0 = 0; r0 < D; r0++)
      {
        double acc = 0.0;
        for (int c0 = 0; c0 < D; c0++)
          acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(R[r0][c0], axis[c0]), __dmul_rn(ball[c0], scale)));
        q[r0] = __dadd_rn(acc, centre[r0]);
        if (q[r0] > a.ub[r0] || q[r0] < a.lb[r0])
          inside = false;
      }
    }
    const uint32_t mask = (lps == 32u) ? __ballot_sync(0xffffffffu, inside) : (inside ? 1u : 0u);
    if (mask)
    {
      const uint32_t winner = (uint32_t)__ffs((int)mask) - 1u;
      if (sub == winner)
      {
        for (int d = 0; d < D; d++)
          a.out[row * D + d] = q[d];
        if (a.attempts)
          a.attempts[i] = tb + winner + 1u;
      }
      return;
    }
  }
  if (sub != 0u)
    return;
  uniform(100u * (uint32_t)(D + 1));
  if (a.attempts)
    a.attempts[i] = 101u;
}
// launchers for other translation units (gc_forest): everything stays on `st`, no synchronisation
int steer_launch(const gc_store* s, const double* d_q, const uint32_t* d_seg, const uint32_t* d_closest,
                 const double* d_dist, uint32_t n, double max_distance, double* d_c1, double* d_c2, cudaStream_t st)
{
  steer_kernel<<<(n + 127) / 128, 128, 0, st>>>(s->x64, d_q, d_seg, d_closest, d_dist, n, s->dim, s->cap_seg,
                                                 max_distance, d_c1, d_c2);
  GC_CUDA(cudaGetLastError());
  count_launch();
  return GC_OK;
}

int sample_informed_launch(int dim, uint32_t n, const double* d_lb, const double* d_ub, const double* d_f1,
                           const double* d_f2, const double* d_cost, unsigned long long seed,
                           unsigned long long first, double* d_out, cudaStream_t st)
{
  InformedArgs a{};
  a.dim = dim;
  a.n = n;
  a.lb = d_lb;
  a.ub = d_ub;
  a.f1 = d_f1;
  a.f2 = d_f2;
  a.cost = d_cost;
  a.seed = seed;
  a.first = first;
  a.out = d_out;
  a.attempts = nullptr;
  sample_informed_kernel<<<(n + 127) / 128, 128, 0, st>>>(a);
  GC_CUDA(cudaGetLastError());
  count_launch();
  return GC_OK;
}

int sample_informed_rows_launch(int dim, uint32_t n, const double* d_lb, const double* d_ub, const double* d_f1,
                                const double* d_f2, const double* d_cost, unsigned long long seed,
                                const uint32_t* d_prob, const unsigned long long* d_ctr, double* d_out, cudaStream_t st,
                                const double* d_basis)
{
  InformedArgs a{};
  a.basis = d_basis;
  a.dim = dim;
  a.n = n;
  a.lb = d_lb;
  a.ub = d_ub;
  a.f1 = d_f1;
  a.f2 = d_f2;
  a.cost = d_cost;
  a.seed = seed;
  a.first = 0;
  a.out = d_out;
  a.attempts = nullptr;
  a.prob = d_prob;
  a.ctr = d_ctr;
  // few samples per call and up to 100 rejection rounds each: the max over the batch sets the kernel time
  a.lanes_per_sample = 32;
  sample_informed_kernel<<<(n + 3) / 4, 128, 0, st>>>(a);
  GC_CUDA(cudaGetLastError());
  count_launch();
  return GC_OK;
}

int informed_basis_launch(int dim, uint32_t n, const double* d_f1, const double* d_f2, double* d_basis, cudaStream_t st)
{
  informed_basis_kernel<<<(n + 127) / 128, 128, 0, st>>>(dim, n, d_f1, d_f2, d_basis);
  GC_CUDA(cudaGetLastError());
  count_launch();
  return GC_OK;
}
}  // namespace gc

using namespace gc;

extern "C" {

// q: host samples [n][D]  -- or --  d_samples: device samples, row src_row[i] (identity when NULL) feeds sample i
static int extend_batch_impl(gc_store_t* s, gc_world_t* w, const double* q, const double* d_samples,
                             const uint32_t* src_row, const uint32_t* seg, uint32_t n, double max_distance,
                             double min_distance, double tolerance, const double* radius, uint32_t near_cap,
                             uint32_t* closest, double* dist, double* next, uint8_t* ok, uint32_t* near_idx,
                             double* near_dist, uint32_t* near_count)
{
  GC_REQUIRE(s && w && (q || d_samples) && closest && dist && next && ok, "gc_extend_batch: NULL argument");
  GC_REQUIRE(s->dim == w->dim, "gc_extend_batch: store dim %d != world dim %d", s->dim, w->dim);
  GC_REQUIRE(s->device == w->device, "gc_extend_batch: store and world live on different devices");
  GC_REQUIRE(!radius || (near_cap > 0 && near_idx && near_dist && near_count), "gc_extend_batch: near outputs missing");
  GC_REQUIRE(max_distance > 0.0 && min_distance > 0.0, "gc_extend_batch: distances must be > 0");
  if (n == 0)
    return GC_OK;
  GC_SET_DEVICE(s->device);
  const int D = s->dim;
  cudaStream_t st = s->stream;
  const size_t qb = sizeof(double) * (size_t)n * D;
  const size_t sb = seg ? align_up(sizeof(uint32_t) * n, 16) : 0;
  const size_t rb = radius ? sizeof(double) * n : 0;
  const size_t lists = radius ? (size_t)n * near_cap : 0;
  GC_TRY(s->h_in.reserve(qb + sb + rb));
  GC_TRY(s->d_q.reserve(qb));
  if (seg)
    GC_TRY(s->d_seg.reserve(sb));
  if (radius)
    GC_TRY(s->d_radius.reserve(rb));
  GC_TRY(s->d_misc.reserve(sizeof(double) * n + sizeof(uint32_t) * n));  // dist | closest
  GC_TRY(w->d_c1.reserve(qb));
  GC_TRY(w->d_c2.reserve(qb));
  GC_TRY(w->d_ok.reserve(n));
  if (radius)
  {
    GC_TRY(s->d_idx.reserve(sizeof(uint32_t) * lists));
    GC_TRY(s->d_dist.reserve(sizeof(double) * lists));
    GC_TRY(s->d_count.reserve(sizeof(uint32_t) * n));
  }
  // host output staging: dist[n] | next[n*D] | near_dist[lists] | closest[n] | near_idx[lists] | near_count[n] | ok[n]
  const size_t out_bytes =
      sizeof(double) * (n + (size_t)n * D + lists) + sizeof(uint32_t) * (n + lists + n) + n + 64;
  GC_TRY(s->h_out.reserve(out_bytes));
  GC_CUDA(cudaStreamSynchronize(st));
  for (uint32_t i = 0; seg && i < n; i++)
    GC_REQUIRE(seg[i] < s->nseg, "sample %u: segment %u of %u", i, seg[i], s->nseg);
  char* h = static_cast<char*>(s->h_in.p);
  const double* d_q = s->d_q.as<double>();
  if (q)
  {
    memcpy(h, q, qb);
    GC_CUDA(cudaMemcpyAsync(s->d_q.p, h, qb, cudaMemcpyHostToDevice, st));
  }
  else if (src_row)
  {
    // samples stay resident in HBM: gather the rows of the active problems on the device
    GC_TRY(s->d_slots.reserve(sizeof(uint32_t) * n));
    memcpy(h, src_row, sizeof(uint32_t) * n);
    GC_CUDA(cudaMemcpyAsync(s->d_slots.p, h, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    const uint32_t total = n * (uint32_t)D;
    gather_rows_kernel<<<(total + 255) / 256, 256, 0, st>>>(d_samples, s->d_slots.as<uint32_t>(), n, D,
                                                             s->d_q.as<double>());
    GC_CUDA(cudaGetLastError());
    count_launch();
  }
  else
    d_q = d_samples;
  if (seg)
  {
    memcpy(h + qb, seg, sizeof(uint32_t) * n);
    GC_CUDA(cudaMemcpyAsync(s->d_seg.p, h + qb, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
  }
  if (radius)
  {
    memcpy(h + qb + sb, radius, rb);
    GC_CUDA(cudaMemcpyAsync(s->d_radius.p, h + qb + sb, rb, cudaMemcpyHostToDevice, st));
  }
  const uint32_t* d_seg = seg ? s->d_seg.as<uint32_t>() : nullptr;
  double* d_dist = s->d_misc.as<double>();
  uint32_t* d_closest = reinterpret_cast<uint32_t*>(d_dist + n);
  GC_TRY(nn1_dev(s, d_q, d_seg, n, d_closest, d_dist, seg == nullptr));
  steer_kernel<<<(n + 127) / 128, 128, 0, st>>>(s->x64, d_q, d_seg, d_closest, d_dist, n, D, s->cap_seg,
                                                 max_distance, w->d_c1.as<double>(), w->d_c2.as<double>());
  GC_CUDA(cudaGetLastError());
  count_launch();
  // states per edge for an edge of length max_distance (grid sizing only)
  uint64_t per_edge = 2;
  for (double nn = 2; max_distance > nn * min_distance && per_edge < (1u << 20); nn *= 2)
    per_edge += (uint64_t)(nn / 2);
  GC_TRY(check_edges_dev(w, w->d_c1.as<double>(), w->d_c2.as<double>(), nullptr, n, min_distance, GC_EDGE_BISECT,
                         w->d_ok.as<uint8_t>(), nullptr, per_edge * n, st));
  if (radius)
    GC_TRY(radius_dev(s, w->d_c2.as<double>(), d_seg, s->d_radius.as<double>(), n, near_cap, s->d_idx.as<uint32_t>(),
                      s->d_dist.as<double>(), s->d_count.as<uint32_t>()));
  double* h_dist = s->h_out.as<double>();
  double* h_next = h_dist + n;
  double* h_ndist = h_next + (size_t)n * D;
  uint32_t* h_closest = reinterpret_cast<uint32_t*>(h_ndist + lists);
  uint32_t* h_nidx = h_closest + n;
  uint32_t* h_ncount = h_nidx + lists;
  uint8_t* h_ok = reinterpret_cast<uint8_t*>(h_ncount + n);
  GC_CUDA(cudaMemcpyAsync(h_dist, d_dist, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  GC_CUDA(cudaMemcpyAsync(h_closest, d_closest, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
  GC_CUDA(cudaMemcpyAsync(h_next, w->d_c2.p, qb, cudaMemcpyDeviceToHost, st));
  GC_CUDA(cudaMemcpyAsync(h_ok, w->d_ok.p, n, cudaMemcpyDeviceToHost, st));
  uint32_t width = 0;
  if (radius)
  {
    // the lists are [n][near_cap] on the device but mostly short: fetch only the first `width` columns
    // (a guess from the previous call) and the rest only if some list turned out longer
    width = s->near_width_hint < near_cap ? s->near_width_hint : near_cap;
    GC_CUDA(cudaMemcpyAsync(h_ncount, s->d_count.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
    GC_CUDA(cudaMemcpy2DAsync(h_ndist, sizeof(double) * near_cap, s->d_dist.p, sizeof(double) * near_cap,
                              sizeof(double) * width, n, cudaMemcpyDeviceToHost, st));
    GC_CUDA(cudaMemcpy2DAsync(h_nidx, sizeof(uint32_t) * near_cap, s->d_idx.p, sizeof(uint32_t) * near_cap,
                              sizeof(uint32_t) * width, n, cudaMemcpyDeviceToHost, st));
  }
  GC_CUDA(cudaStreamSynchronize(st));
  if (radius)
  {
    uint32_t mx = 0;
    for (uint32_t i = 0; i < n; i++)
      mx = h_ncount[i] > mx ? h_ncount[i] : mx;
    const uint32_t need = mx < near_cap ? mx : near_cap;
    if (need > width)
    {
      const uint32_t extra = need - width;
      GC_CUDA(cudaMemcpy2DAsync(h_ndist + width, sizeof(double) * near_cap, s->d_dist.as<double>() + width,
                                sizeof(double) * near_cap, sizeof(double) * extra, n, cudaMemcpyDeviceToHost, st));
      GC_CUDA(cudaMemcpy2DAsync(h_nidx + width, sizeof(uint32_t) * near_cap, s->d_idx.as<uint32_t>() + width,
                                sizeof(uint32_t) * near_cap, sizeof(uint32_t) * extra, n, cudaMemcpyDeviceToHost, st));
      GC_CUDA(cudaStreamSynchronize(st));
    }
    uint32_t hint = mx + mx / 4 + 8;
    s->near_width_hint = hint < 32 ? 32 : hint;
  }
  int rc = GC_OK;
  memcpy(dist, h_dist, sizeof(double) * n);
  memcpy(closest, h_closest, sizeof(uint32_t) * n);
  memcpy(next, h_next, qb);
  for (uint32_t i = 0; i < n; i++)
  {
    if (h_closest[i] == GC_NO_NODE)
      ok[i] = 0;  // empty tree: nothing to extend from
    else
      ok[i] = (h_dist[i] < tolerance) ? 1 : h_ok[i];  // tree.cpp:76-77
  }
  if (radius)
  {
    memcpy(near_count, h_ncount, sizeof(uint32_t) * n);
    for (uint32_t i = 0; i < n; i++)
    {
      const uint32_t m = h_ncount[i] < near_cap ? h_ncount[i] : near_cap;
      if (h_ncount[i] > near_cap)
        rc = GC_W_TRUNCATED;
      memcpy(near_idx + (size_t)i * near_cap, h_nidx + (size_t)i * near_cap, sizeof(uint32_t) * m);
      memcpy(near_dist + (size_t)i * near_cap, h_ndist + (size_t)i * near_cap, sizeof(double) * m);
    }
  }
  return rc;
}

int gc_extend_batch(gc_store_t* s, gc_world_t* w, const double* q, const uint32_t* seg, uint32_t n,
                    double max_distance, double min_distance, double tolerance, const double* radius,
                    uint32_t near_cap, uint32_t* closest, double* dist, double* next, uint8_t* ok,
                    uint32_t* near_idx, double* near_dist, uint32_t* near_count)
{
  GC_REQUIRE(q != nullptr, "gc_extend_batch: NULL samples");
  return extend_batch_impl(s, w, q, nullptr, nullptr, seg, n, max_distance, min_distance, tolerance, radius, near_cap,
                           closest, dist, next, ok, near_idx, near_dist, near_count);
}

int gc_extend_batch_dsamples(gc_store_t* s, gc_world_t* w, const double* d_samples, const uint32_t* src_row,
                             const uint32_t* seg, uint32_t n, double max_distance, double min_distance,
                             double tolerance, const double* radius, uint32_t near_cap, uint32_t* closest,
                             double* dist, double* next, uint8_t* ok, uint32_t* near_idx, double* near_dist,
                             uint32_t* near_count)
{
  GC_REQUIRE(d_samples != nullptr, "gc_extend_batch_dsamples: NULL samples");
  return extend_batch_impl(s, w, nullptr, d_samples, src_row, seg, n, max_distance, min_distance, tolerance, radius,
                           near_cap, closest, dist, next, ok, near_idx, near_dist, near_count);
}

// device-resident variant of gc_sample_uniform: fills d_out[n][dim] on `cuda_stream` (NULL: default stream)
int gc_sample_uniform_dev(int device, int dim, const double* lb, const double* ub, uint64_t seed,
                          uint64_t first_index, uint32_t n, double* d_out, void* cuda_stream)
{
  GC_REQUIRE(lb && ub && d_out, "gc_sample_uniform_dev: NULL argument");
  GC_REQUIRE(dim >= 1 && dim <= GC_MAX_DIM, "gc_sample_uniform_dev: dim %d", dim);
  if (n == 0)
    return GC_OK;
  GC_SET_DEVICE(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  double* d_b = nullptr;
  GC_CUDA(cudaMalloc(&d_b, sizeof(double) * 2 * dim));
  GC_CUDA(cudaMemcpyAsync(d_b, lb, sizeof(double) * dim, cudaMemcpyHostToDevice, st));
  GC_CUDA(cudaMemcpyAsync(d_b + dim, ub, sizeof(double) * dim, cudaMemcpyHostToDevice, st));
  const uint32_t total = n * (uint32_t)((dim + 1) / 2);
  sample_uniform_kernel<<<(total + 255) / 256, 256, 0, st>>>(d_b, d_b + dim, dim, seed, first_index, n, d_out);
  cudaError_t e = cudaGetLastError();
  count_launch();
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(st);
  cudaFree(d_b);
  if (e != cudaSuccess)
  {
    gc::set_error("gc_sample_uniform_dev: %s", cudaGetErrorString(e));
    return GC_E_CUDA;
  }
  return GC_OK;
}

int gc_sample_uniform(int device, int dim, const double* lb, const double* ub, uint64_t seed, uint64_t first_index,
                      uint32_t n, double* out)
{
  GC_REQUIRE(lb && ub && out, "gc_sample_uniform: NULL argument");
  GC_REQUIRE(dim >= 1 && dim <= GC_MAX_DIM, "gc_sample_uniform: dim %d", dim);
  if (n == 0)
    return GC_OK;
  int ndev = 0;
  GC_TRY(gc_device_count(&ndev));
  GC_REQUIRE(device >= 0 && device < ndev, "gc_sample_uniform: device %d of %d", device, ndev);
  GC_SET_DEVICE(device);
  double *d_b = nullptr, *d_out = nullptr;
  GC_CUDA(cudaMalloc(&d_b, sizeof(double) * 2 * dim));
  cudaError_t e = cudaMalloc(&d_out, sizeof(double) * (size_t)n * dim);
  if (e != cudaSuccess)
  {
    cudaFree(d_b);
    gc::set_error("gc_sample_uniform: %s", cudaGetErrorString(e));
    return GC_E_CUDA;
  }
  cudaMemcpy(d_b, lb, sizeof(double) * dim, cudaMemcpyHostToDevice);
  cudaMemcpy(d_b + dim, ub, sizeof(double) * dim, cudaMemcpyHostToDevice);
  const uint32_t total = n * (uint32_t)((dim + 1) / 2);
  sample_uniform_kernel<<<(total + 255) / 256, 256>>>(d_b, d_b + dim, dim, seed, first_index, n, d_out);
  count_launch();
  e = cudaGetLastError();
  if (e == cudaSuccess)
    e = cudaMemcpy(out, d_out, sizeof(double) * (size_t)n * dim, cudaMemcpyDeviceToHost);
  cudaFree(d_b);
  cudaFree(d_out);
  if (e != cudaSuccess)
  {
    gc::set_error("gc_sample_uniform: %s", cudaGetErrorString(e));
    return GC_E_CUDA;
  }
  return GC_OK;
}

// informed sampling for n problems (foci f1/f2 [n][dim], cost [n]); all pointers DEVICE pointers except lb/ub (host)
int gc_sample_informed_dev(int device, int dim, const double* lb, const double* ub, const double* d_f1,
                           const double* d_f2, const double* d_cost, uint64_t seed, uint64_t first_index, uint32_t n,
                           double* d_out, uint32_t* d_attempts, void* cuda_stream)
{
  GC_REQUIRE(lb && ub && d_f1 && d_f2 && d_cost && d_out, "gc_sample_informed_dev: NULL argument");
  GC_REQUIRE(dim >= 1 && dim <= GC_MAX_DIM, "gc_sample_informed_dev: dim %d", dim);
  if (n == 0)
    return GC_OK;
  GC_SET_DEVICE(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  double* d_b = nullptr;
  GC_CUDA(cudaMalloc(&d_b, sizeof(double) * 2 * dim));
  GC_CUDA(cudaMemcpyAsync(d_b, lb, sizeof(double) * dim, cudaMemcpyHostToDevice, st));
  GC_CUDA(cudaMemcpyAsync(d_b + dim, ub, sizeof(double) * dim, cudaMemcpyHostToDevice, st));
  InformedArgs a{};
  a.dim = dim;
  a.n = n;
  a.lb = d_b;
  a.ub = d_b + dim;
  a.f1 = d_f1;
  a.f2 = d_f2;
  a.cost = d_cost;
  a.seed = seed;
  a.first = first_index;
  a.out = d_out;
  a.attempts = d_attempts;
  sample_informed_kernel<<<(n + 127) / 128, 128, 0, st>>>(a);
  cudaError_t e = cudaGetLastError();
  count_launch();
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(st);
  cudaFree(d_b);
  if (e != cudaSuccess)
  {
    gc::set_error("gc_sample_informed_dev: %s", cudaGetErrorString(e));
    return GC_E_CUDA;
  }
  return GC_OK;
}

int gc_sample_informed(int device, int dim, const double* lb, const double* ub, const double* f1, const double* f2,
                       const double* cost, uint64_t seed, uint64_t first_index, uint32_t n, double* out,
                       uint32_t* attempts)
{
  GC_REQUIRE(lb && ub && f1 && f2 && cost && out, "gc_sample_informed: NULL argument");
  GC_REQUIRE(dim >= 1 && dim <= GC_MAX_DIM, "gc_sample_informed: dim %d", dim);
  if (n == 0)
    return GC_OK;
  int ndev = 0;
  GC_TRY(gc_device_count(&ndev));
  GC_REQUIRE(device >= 0 && device < ndev, "gc_sample_informed: device %d of %d", device, ndev);
  GC_SET_DEVICE(device);
  const size_t nd = (size_t)n * dim;
  double* d_buf = nullptr;  // f1 | f2 | out | cost
  uint32_t* d_att = nullptr;
  GC_CUDA(cudaMalloc(&d_buf, sizeof(double) * (3 * nd + n)));
  cudaError_t e = cudaMalloc(&d_att, sizeof(uint32_t) * n);
  if (e == cudaSuccess)
    e = cudaMemcpy(d_buf, f1, sizeof(double) * nd, cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(d_buf + nd, f2, sizeof(double) * nd, cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(d_buf + 3 * nd, cost, sizeof(double) * n, cudaMemcpyHostToDevice);
  int rc = GC_OK;
  if (e == cudaSuccess)
    rc = gc_sample_informed_dev(device, dim, lb, ub, d_buf, d_buf + nd, d_buf + 3 * nd, seed, first_index, n,
                                d_buf + 2 * nd, d_att, cudaStreamLegacy);
  if (e == cudaSuccess && rc == GC_OK)
    e = cudaMemcpy(out, d_buf + 2 * nd, sizeof(double) * nd, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && rc == GC_OK && attempts)
    e = cudaMemcpy(attempts, d_att, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost);
  cudaFree(d_buf);
  cudaFree(d_att);
  if (e != cudaSuccess)
  {
    gc::set_error("gc_sample_informed: %s", cudaGetErrorString(e));
    return GC_E_CUDA;
  }
  return rc;
}

}  // extern "C"
