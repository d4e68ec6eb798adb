// nn.cu -- device-resident node store and the brute-force nearest / radius / k-nearest kernels.
//
// Replaces the reference's CPU structures behind graph::core::NearestNeighbors
//   KdTree  src/graph_core/datastructure/kdtree.cpp:91-472
//   Vector  src/graph_core/datastructure/vector.cpp:38-138   (the semantic twin: scan order = slot order,
//                                                             strict '<', lowest slot wins ties)
//
// Design (DESIGN.md "NN path"): every query is a tiled brute-force scan of the segment's fp32
// dimension-major SoA with coalesced float4 loads; the fp32 distance only PRE-SELECTS: any node
// whose fp32 distance is within the rounding slack of the decision threshold is re-evaluated
// from the fp64 master copy with exactly the oracle's arithmetic (sequential sum of squares,
// individually rounded, then sqrt) and the reference's comparisons are applied to that value.
#include <math_constants.h>

#include <cfloat>
#include <cmath>

#include "gc_common.cuh"

namespace gc
{
constexpr int kScanThreads = 256;
constexpr int kNodesPerThread = 4;
constexpr int kTile = kScanThreads * kNodesPerThread;  // 1024 nodes per tile
constexpr int kQT = 8;                                 // queries sharing one pass over a tile
constexpr int kRowPad = kStoreRowPad;                  // floats of padding after each SoA row
static_assert(kRowPad == kTile, "the scans read one tile past the last row");
constexpr uint32_t kSortMax = 16384;                   // single-CTA bitonic capacity

__device__ __forceinline__ float u2f(uint32_t u) { return __uint_as_float(u); }
__device__ __forceinline__ uint32_t f2u(float f) { return __float_as_uint(f); }

// exact fp64 Euclidean distance, operation order fixed by DESIGN.md (sequential, individually rounded)
template <int D>
__device__ __forceinline__ double dist64(const double* __restrict__ row, const double* __restrict__ q)
{
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < D; i++)
  {
    double t = __dsub_rn(row[i], q[i]);
    s = __dadd_rn(s, __dmul_rn(t, t));
  }
  return __dsqrt_rn(s);
}

// upper bound (squared) of the fp32 distance of any node whose fp64 distance is <= the fp64
// distance corresponding to fp32 value m (or < radius m); see include/gcb200_spec.h
__device__ __forceinline__ float slack_thr2(float m, float slack_abs)
{
  float b = m * (1.0f + 4.0e-6f) + 3.0f * slack_abs;
  return b * b * (1.0f + 1.0e-6f);
}

// ---------------------------------------------------------------------------------------------
// store maintenance kernels
// ---------------------------------------------------------------------------------------------
// rows[n][dim] (fp64) -> x64 master rows + fp32 SoA columns, gslot[i] = global slot of row i
__global__ void store_scatter_kernel(const double* __restrict__ rows, const uint32_t* __restrict__ gslot, uint32_t n,
                                     int dim, size_t row_stride, float* __restrict__ x32, double* __restrict__ x64)
{
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * (uint32_t)dim)
    return;
  uint32_t r = i / dim, d = i % dim;
  double v = rows[i];
  size_t slot = gslot[r];
  x64[slot * dim + d] = v;
  x32[(size_t)d * row_stride + slot] = (float)v;
}

__global__ void store_fill_slots_kernel(uint32_t* gslot, uint32_t first, uint32_t n)
{
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    gslot[i] = first + i;
}

// ---------------------------------------------------------------------------------------------
// tiled scan: MODE 0 = nearest (exact, lexicographic (dist64, slot) minimum per chunk)
//             MODE 1 = radius  (exact strict dist64 < radius, unordered append to per-query lists)
// grid = (chunks, query groups).  A group holds up to kQT queries of one segment.
// ---------------------------------------------------------------------------------------------
struct ScanArgs
{
  const float* x32;
  const double* x64;
  const uint32_t* used;    // [nseg]
  const double* q;         // [nq][D]
  const uint32_t* seg;     // [nq] or nullptr (segment 0)
  const double* radius;    // MODE 1
  const float* thr2;       // MODE 2: per-query squared fp32 candidate threshold
  uint32_t nq;
  uint32_t group_size;     // queries per group (kQT when seg == nullptr, else 1)
  uint32_t cap_seg;
  size_t row_stride;
  float slack_abs;
  float xn_max;            // streaming scan: upper bound of |x|^2 over the store (fp32)
  uint32_t* gmin;          // streaming scan MODE 0: [nq] shared running fp32 minima (order-preserving keys)
  uint32_t* ctl;           // streaming scan: [groups][2] tile hand-out and finished-CTA counters (self-resetting)
  float* out_c2;           // streaming scan MODE 2: candidates keep their fp32 pre-selection value (no fp64 here)
  uint32_t tile_begin;     // streaming scan: first tile and number of tiles of this launch (0, 0 = the whole store)
  uint32_t tile_count;
  int thr_raw;             // MODE 2: thr2 already is the final expanded-form threshold (second collecting phase)
  // MODE 0 outputs: partials [chunks][nq]
  double* part_d;
  uint32_t* part_i;
  // MODE 1 outputs
  uint32_t cap;
  uint32_t* out_idx;
  double* out_dist;
  uint32_t* out_count;
};

template <int D, int MODE>
__global__ void __launch_bounds__(kScanThreads, 2) scan_kernel(ScanArgs a)
{
  __shared__ float qs32[kQT][D];
  __shared__ double qs64[kQT][D];
  __shared__ uint32_t cur[kQT];                 // running fp32 minimum (bits) of this CTA
  __shared__ unsigned long long best64[kQT];    // running exact minimum (bits)
  __shared__ unsigned long long prev64[kQT];
  __shared__ uint32_t bestidx[kQT];
  __shared__ float thr_s[kQT];

  const int tid = threadIdx.x;
  const uint32_t q0 = blockIdx.y * a.group_size;
  const int qcount = min(a.group_size, a.nq - q0);
  const uint32_t seg = a.seg ? a.seg[q0] : 0u;
  const uint32_t n_s = a.used[seg];
  const size_t seg_base = (size_t)seg * a.cap_seg;

  for (int i = tid; i < qcount * D; i += kScanThreads)
  {
    double v = a.q[(size_t)q0 * D + i];
    qs64[i / D][i % D] = v;
    qs32[i / D][i % D] = (float)v;
  }
  if (tid < kQT)
  {
    cur[tid] = 0x7f800000u;  // +inf
    best64[tid] = 0x7ff0000000000000ull;
    prev64[tid] = 0x7ff0000000000000ull;
    bestidx[tid] = GC_NO_NODE;
    if (MODE == 1 && tid < qcount)
    {
      double r = a.radius[q0 + tid];
      float rf = (r > 0.0) ? (float)fmin(r, 1.0e18) : -1.0f;
      thr_s[tid] = (rf < 0.0f) ? -1.0f : slack_thr2(rf * (1.0f + 1.0e-6f), a.slack_abs);
    }
    if (MODE == 2 && tid < qcount)
      thr_s[tid] = a.thr2[q0 + tid];  // squared fp32 candidate threshold (already widened by the slack)
  }
  __syncthreads();

  const uint32_t ntiles = (n_s + kTile - 1) / kTile;
  for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
  {
    const uint32_t local0 = tile * kTile + tid * kNodesPerThread;  // first of this thread's 4 slots
    float x[D][kNodesPerThread];
#pragma unroll
    for (int d = 0; d < D; d++)
    {
      const float4 v = __ldg(reinterpret_cast<const float4*>(a.x32 + (size_t)d * a.row_stride + seg_base + local0));
      x[d][0] = v.x;
      x[d][1] = v.y;
      x[d][2] = v.z;
      x[d][3] = v.w;
    }
    float d2[kQT][kNodesPerThread];
#pragma unroll
    for (int q = 0; q < kQT; q++)
    {
      if (q < qcount)
      {
        float acc[kNodesPerThread];
#pragma unroll
        for (int r = 0; r < kNodesPerThread; r++)
          acc[r] = 0.0f;
#pragma unroll
        for (int d = 0; d < D; d++)
        {
          const float qv = qs32[q][d];
#pragma unroll
          for (int r = 0; r < kNodesPerThread; r++)
          {
            const float t = x[d][r] - qv;
            acc[r] = fmaf(t, t, acc[r]);
          }
        }
#pragma unroll
        for (int r = 0; r < kNodesPerThread; r++)
        {
          // slots beyond the segment's used range, tombstones (x32[0] = +inf) and NaNs never qualify
          const bool valid = (local0 + r < n_s) && (acc[r] < CUDART_INF_F);
          d2[q][r] = valid ? acc[r] : CUDART_INF_F;
        }
        if (MODE == 0)
        {
          const float tmin = fminf(fminf(d2[q][0], d2[q][1]), fminf(d2[q][2], d2[q][3]));
          const uint32_t wmin = __reduce_min_sync(0xffffffffu, f2u(tmin));
          if ((tid & 31) == 0 && wmin < cur[q])
            atomicMin(&cur[q], wmin);
        }
      }
    }
    if (MODE == 0)
    {
      __syncthreads();
      bool pass = false;
      float thr[kQT];
#pragma unroll
      for (int q = 0; q < kQT; q++)
      {
        thr[q] = -1.0f;
        if (q < qcount)
        {
          const float m2 = u2f(cur[q]);
          thr[q] = (m2 < CUDART_INF_F) ? slack_thr2(sqrtf(m2) * (1.0f + 1.0e-6f), a.slack_abs) : -1.0f;
#pragma unroll
          for (int r = 0; r < kNodesPerThread; r++)
            pass |= (d2[q][r] <= thr[q]);
        }
      }
      if (__syncthreads_or(pass))
      {
        // step A: exact distances of the candidates -> running exact minimum
#pragma unroll
        for (int q = 0; q < kQT; q++)
          if (q < qcount)
#pragma unroll
            for (int r = 0; r < kNodesPerThread; r++)
              if (d2[q][r] <= thr[q])
              {
                const double e = dist64<D>(a.x64 + (seg_base + local0 + r) * D, qs64[q]);
                atomicMin(&best64[q], (unsigned long long)__double_as_longlong(e));
              }
        __syncthreads();
        // step B: a strictly smaller exact minimum invalidates the slot kept for the old one
        if (tid < qcount && best64[tid] != prev64[tid])
        {
          bestidx[tid] = GC_NO_NODE;
          prev64[tid] = best64[tid];
        }
        __syncthreads();
        // step C: lowest slot among the candidates that attain the exact minimum
#pragma unroll
        for (int q = 0; q < kQT; q++)
          if (q < qcount)
#pragma unroll
            for (int r = 0; r < kNodesPerThread; r++)
              if (d2[q][r] <= thr[q])
              {
                const double e = dist64<D>(a.x64 + (seg_base + local0 + r) * D, qs64[q]);
                if ((unsigned long long)__double_as_longlong(e) == best64[q])
                  atomicMin(&bestidx[q], local0 + r);
              }
        __syncthreads();
      }
    }
    else
    {
#pragma unroll
      for (int q = 0; q < kQT; q++)
        if (q < qcount)
#pragma unroll
          for (int r = 0; r < kNodesPerThread; r++)
            if (d2[q][r] <= thr_s[q])
            {
              const double e = dist64<D>(a.x64 + (seg_base + local0 + r) * D, qs64[q]);
              // MODE 1: strict radius test (vector.cpp:75 / kdtree.cpp:167); MODE 2: every candidate is kept
              if (MODE == 2 || e < a.radius[q0 + q])
              {
                const uint32_t pos = atomicAdd(&a.out_count[q0 + q], 1u);
                if (pos < a.cap)
                {
                  a.out_idx[(size_t)(q0 + q) * a.cap + pos] = local0 + r;
                  a.out_dist[(size_t)(q0 + q) * a.cap + pos] = e;
                }
              }
            }
    }
  }
  if (MODE == 0)
  {
    __syncthreads();
    if (tid < qcount)
    {
      a.part_d[(size_t)blockIdx.x * a.nq + q0 + tid] = __longlong_as_double((long long)best64[tid]);
      a.part_i[(size_t)blockIdx.x * a.nq + q0 + tid] = bestidx[tid];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// streaming scan for LARGE single-segment stores (BASELINE config 4: 2^20 nodes x 12 DoF):
//   - node tiles of kStTile slots x D dimensions land in shared memory through 1-D bulk async copies
//     (cp.async.bulk ... mbarrier::complete_tx, the TMA copy engine; kStStages tiles in flight per CTA,
//     several CTAs per SM) and are consumed into registers at once, so HBM latency is hidden by the
//     copy engine instead of by load instructions;
//   - up to kStGQ queries share every tile: the node data is read once per query GROUP, the inner
//     loop is the expanded form  c2 = |x|^2 + |q|^2 - 2 x.q  = D fused multiply-adds per (node, query)
//     pair (half the instructions of sum (x-q)^2);
//   - as everywhere in this file the fp32 value only PRE-SELECTS: its absolute error is bounded by
//     e2 (see stream_e2), every node within the widened threshold goes to the fp64 decision.
// MODE 0 nearest (per-CTA exact (dist, slot) minimum -> nn1_finalize_kernel), MODE 1 radius,
// MODE 2 collect under a given per-query threshold (large-N k-nearest).
// ---------------------------------------------------------------------------------------------
constexpr int kStThreads = 128;
constexpr int kStTile = kStThreads * kNodesPerThread;  // 512 nodes per tile
constexpr int kStStages = 2;
constexpr int kStGQ = 64;     // queries per group
constexpr int kStList = 512;  // candidate list entries per CTA (MODE 0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "GC_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra GC_MBAR_DONE;\n"
      "bra GC_MBAR_WAIT;\n"
      "GC_MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// order-preserving map float -> uint32 (negative values included: c2 can round slightly below zero)
__device__ __forceinline__ uint32_t fkey(float f)
{
  const uint32_t b = __float_as_uint(f);
  return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
}
// candidate threshold for a running minimum cmin of the expanded-form value c2 (absolute error <= e2 each):
// a node whose fp64 distance is <= that of the argmin node has c2 <= (sqrt(cmin + e2) + 2 s_abs)^2 + e2
__device__ __forceinline__ float stream_thr(float cmin, float e2, float slack_abs)
{
  const float b = sqrtf(fmaxf(cmin + e2, 0.0f)) * (1.0f + 1.0e-6f) + 3.0f * slack_abs;
  return b * b * (1.0f + 1.0e-6f) + e2 * (1.0f + 1.0e-6f);
}

__device__ __forceinline__ float funkey(uint32_t k)
{
  return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}

constexpr int kStQPT = 4;              // queries per trip of the sweep (independent accumulator chains)
constexpr uint32_t kStNoTile = 0xFFFFFFFFu;

struct StreamSmem
{
  float* xs;            // [kStStages][D][kStTile]
  uint64_t* bars;       // [kStStages]
  double* q64;          // [kStGQ][D]
  double* le;           // [kStList]  candidate: fp32 pre-selection value, then its exact distance
  unsigned long long* best64;  // [kStGQ]
  unsigned long long* prev64;  // [kStGQ]
  float* qm2;           // [kStGQ][DP]  -2 * q (fp32)
  float* qn;            // [kStGQ]      |q|^2 (fp32)
  float* e2;            // [kStGQ]
  float* thr;           // [kStGQ]
  uint32_t* curkey;     // [kStGQ]
  uint32_t* bestidx;    // [kStGQ]
  uint32_t* lq;         // [kStList]
  uint32_t* ls;         // [kStList]
  uint32_t* cnt;        // [0] list length, [1] "this CTA finished last", [2..2+kStStages) tile of each stage
};
template <int D>
__host__ __device__ constexpr size_t stream_smem_bytes()
{
  constexpr int DP = (D + 3) & ~3;
  return sizeof(float) * kStStages * D * kStTile + 8 * kStStages + sizeof(double) * kStGQ * D + 8 * kStList +
         16 * kStGQ + sizeof(float) * kStGQ * DP + 4 * 5 * kStGQ + 8 * kStList + 32;
}

// Tiles are handed out dynamically (one counter per query group), so CTAs that start late or run on a
// busier SM simply take fewer tiles; the CTA that finishes last for its group resets the counters and, in
// MODE 0, reduces the per-CTA (dist, slot) minima to the final answer (no separate finalize launch).
template <int D, int MODE>
__global__ void __launch_bounds__(kStThreads, 3) scan_stream_kernel(ScanArgs a)
{
  constexpr int DP = (D + 3) & ~3;
  extern __shared__ __align__(128) unsigned char st_raw[];
  StreamSmem m;
  {
    unsigned char* p = st_raw;
    m.xs = reinterpret_cast<float*>(p);
    p += sizeof(float) * kStStages * D * kStTile;
    m.bars = reinterpret_cast<uint64_t*>(p);
    p += 8 * kStStages;
    m.q64 = reinterpret_cast<double*>(p);
    p += sizeof(double) * kStGQ * D;
    m.le = reinterpret_cast<double*>(p);
    p += 8 * kStList;
    m.best64 = reinterpret_cast<unsigned long long*>(p);
    p += 8 * kStGQ;
    m.prev64 = reinterpret_cast<unsigned long long*>(p);
    p += 8 * kStGQ;
    m.qm2 = reinterpret_cast<float*>(p);
    p += sizeof(float) * kStGQ * DP;
    m.qn = reinterpret_cast<float*>(p);
    p += 4 * kStGQ;
    m.e2 = reinterpret_cast<float*>(p);
    p += 4 * kStGQ;
    m.thr = reinterpret_cast<float*>(p);
    p += 4 * kStGQ;
    m.curkey = reinterpret_cast<uint32_t*>(p);
    p += 4 * kStGQ;
    m.bestidx = reinterpret_cast<uint32_t*>(p);
    p += 4 * kStGQ;
    m.lq = reinterpret_cast<uint32_t*>(p);
    p += 4 * kStList;
    m.ls = reinterpret_cast<uint32_t*>(p);
    p += 4 * kStList;
    m.cnt = reinterpret_cast<uint32_t*>(p);
  }
  const int tid = threadIdx.x;
  const uint32_t q0 = blockIdx.y * a.group_size;
  const int gq = (int)min(a.group_size, a.nq - q0);
  const int gqp = (gq + kStQPT - 1) / kStQPT * kStQPT;  // padded with inert queries up to a whole trip
  const uint32_t n_s = a.used[0];
  const uint32_t all_tiles = (n_s + kStTile - 1) / kStTile;
  const uint32_t tile0 = min(a.tile_begin, all_tiles);
  const uint32_t ntiles = a.tile_count ? min(a.tile_count, all_tiles - tile0) : all_tiles - tile0;  // of this launch
  uint32_t* tile_ctr = a.ctl + 2 * blockIdx.y;
  uint32_t* done_ctr = tile_ctr + 1;
  volatile uint32_t* tile_of = m.cnt + 2;

  // take the next tile of this group and put its D rows in flight (thread 0 only)
  auto fetch = [&](int stage) {
    const uint32_t k = atomicAdd(tile_ctr, 1u);
    if (k < ntiles)
    {
      const uint32_t t = tile0 + k;
      tile_of[stage] = t;
      mbar_expect_tx(&m.bars[stage], (uint32_t)(D * kStTile * sizeof(float)));
#pragma unroll
      for (int d = 0; d < D; d++)
        bulk_g2s(m.xs + ((size_t)stage * D + d) * kStTile, a.x32 + (size_t)d * a.row_stride + (size_t)t * kStTile,
                 (uint32_t)(kStTile * sizeof(float)), &m.bars[stage]);
    }
    else
      tile_of[stage] = kStNoTile;
  };
  if (tid == 0)
  {
    for (int s = 0; s < kStStages; s++)
      mbar_init(&m.bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    m.cnt[0] = 0;
    m.cnt[1] = 0;
    for (int s = 0; s < kStStages; s++)  // the node data starts moving before the queries are even read
      fetch(s);
  }
  for (int i = tid; i < gqp * DP; i += kStThreads)
  {
    const int q = i / DP, d = i % DP;
    const double v = (q < gq && d < D) ? a.q[(size_t)(q0 + q) * D + d] : 0.0;
    if (d < D)
      m.q64[q * D + d] = v;
    m.qm2[i] = -2.0f * (float)v;
  }
  __syncthreads();
  if (tid < gqp)
  {
    float qn = 0.0f;
#pragma unroll
    for (int d = 0; d < D; d++)
    {
      const float v = (float)m.q64[tid * D + d];
      qn = fmaf(v, v, qn);
    }
    // absolute error bound of the expanded form: each of its ~3D+2 roundings is <= 2^-24 times a partial sum that
    // never exceeds 2 (|x|^2 + |q|^2); a factor 1.5 on top
    const float e2 = (float)(3 * D + 8) * 5.9604645e-8f * 1.5f * (a.xn_max + qn) * 2.0f + FLT_MIN;
    m.e2[tid] = e2;
    m.curkey[tid] = fkey(CUDART_INF_F);
    m.best64[tid] = 0x7ff0000000000000ull;
    m.prev64[tid] = 0x7ff0000000000000ull;
    m.bestidx[tid] = GC_NO_NODE;
    float thr = CUDART_INF_F;
    if (MODE == 1 && tid < gq)
    {
      const double r = a.radius[q0 + tid];
      const float rf = (r > 0.0) ? (float)fmin(r, 1.0e18) : -1.0f;
      if (rf < 0.0f)
        thr = -CUDART_INF_F;
      else
      {
        const float b = rf * (1.0f + 1.0e-6f) + 3.0f * a.slack_abs;
        thr = b * b * (1.0f + 1.0e-6f) + e2 * (1.0f + 1.0e-6f);
      }
    }
    if (MODE == 2 && tid < gq)
      thr = a.thr_raw ? a.thr2[q0 + tid] : a.thr2[q0 + tid] * (1.0f + 8.0e-6f) + e2 * (1.0f + 1.0e-6f);
    if (tid >= gq)  // inert query: |q|^2 = +inf makes every value +inf, which never passes "< +inf"
    {
      qn = CUDART_INF_F;
      thr = -CUDART_INF_F;
    }
    m.qn[tid] = qn;
    m.thr[tid] = thr;
  }
  __syncthreads();

  volatile float* thr_v = m.thr;
  // MODE 0: the running fp32 minimum is shared between the CTAs of a query group through a.gmin (relaxed: a stale
  // value only makes the candidate threshold looser)
  auto exchange = [&]() {
    if (tid < gq)
    {
      const uint32_t mine = m.curkey[tid];
      uint32_t g = *reinterpret_cast<volatile uint32_t*>(a.gmin + q0 + tid);
      if (mine < g)
        g = min(mine, atomicMin(a.gmin + q0 + tid, mine));
      if (g < mine)
        m.curkey[tid] = g;
      if (g != fkey(CUDART_INF_F) && g <= mine)
        thr_v[tid] = stream_thr(funkey(g), m.e2[tid], a.slack_abs);
    }
  };
  for (uint32_t it = 0;; it++)
  {
    const int stage = (int)(it % kStStages);
    const uint32_t tile = tile_of[stage];
    if (tile == kStNoTile)  // CTA-uniform
      break;
    mbar_wait(&m.bars[stage], (it / kStStages) & 1u);
    const uint32_t local0 = tile * kStTile + tid * kNodesPerThread;
    float x[D][kNodesPerThread];
#pragma unroll
    for (int d = 0; d < D; d++)
    {
      const float4 v = *reinterpret_cast<const float4*>(m.xs + ((size_t)stage * D + d) * kStTile + tid * kNodesPerThread);
      x[d][0] = v.x;
      x[d][1] = v.y;
      x[d][2] = v.z;
      x[d][3] = v.w;
    }
    __syncthreads();  // every thread holds its nodes in registers: the stage can be refilled
    if (tid == 0)
      fetch(stage);
    float xn[kNodesPerThread];
#pragma unroll
    for (int r = 0; r < kNodesPerThread; r++)
    {
      float s2 = 0.0f;
#pragma unroll
      for (int d = 0; d < D; d++)
        s2 = fmaf(x[d][r], x[d][r], s2);
      // slots beyond the used range, tombstones (x32[0] = +inf) and NaNs never qualify: |x|^2 = +inf, x = 0
      const bool valid = (local0 + r < n_s) && (s2 < CUDART_INF_F);
      xn[r] = valid ? s2 : CUDART_INF_F;
      if (!valid)
      {
#pragma unroll
        for (int d = 0; d < D; d++)
          x[d][r] = 0.0f;
      }
    }
    // expanded-form values of this thread's four nodes against queries q .. q+kStQPT-1.  The D fused multiply-adds
    // per pair run as packed FFMA2 (fma.rn.f32x2, sm_100): two NODES per 64-bit register pair, the query coordinate
    // as a scalar operand broadcast to both halves -- 1.5 register operands per FMA instead of 2 (the register
    // file's operand delivery, not the issue rate, bounds the scalar form; tools/ffma_forms4.cu).  Each half is an
    // ordinary IEEE fma: the values are bit-identical to the scalar loop's.
    const float2 xn01 = make_float2(xn[0], xn[1]), xn23 = make_float2(xn[2], xn[3]);
    auto trip_values = [&](int q, float (&acc)[kStQPT][kNodesPerThread]) {
      const float4 qn4 = *reinterpret_cast<const float4*>(m.qn + q);
      const float qna[4] = { qn4.x, qn4.y, qn4.z, qn4.w };
      float2 lo[kStQPT], hi[kStQPT];
#pragma unroll
      for (int h = 0; h < kStQPT; h++)
      {
        lo[h] = __fadd2_rn(xn01, make_float2(qna[h], qna[h]));
        hi[h] = __fadd2_rn(xn23, make_float2(qna[h], qna[h]));
      }
#pragma unroll
      for (int d4 = 0; d4 < DP; d4 += 4)
      {
        float qa[kStQPT][4];
#pragma unroll
        for (int h = 0; h < kStQPT; h++)
        {
          const float4 u = *reinterpret_cast<const float4*>(m.qm2 + (q + h) * DP + d4);
          qa[h][0] = u.x;
          qa[h][1] = u.y;
          qa[h][2] = u.z;
          qa[h][3] = u.w;
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (d4 + j < D)
          {
            const float2 x01 = make_float2(x[d4 + j][0], x[d4 + j][1]), x23 = make_float2(x[d4 + j][2], x[d4 + j][3]);
#pragma unroll
            for (int h = 0; h < kStQPT; h++)
            {
              const float2 qq = make_float2(qa[h][j], qa[h][j]);
              lo[h] = __ffma2_rn(x01, qq, lo[h]);
              hi[h] = __ffma2_rn(x23, qq, hi[h]);
            }
          }
      }
#pragma unroll
      for (int h = 0; h < kStQPT; h++)
      {
        acc[h][0] = lo[h].x;
        acc[h][1] = lo[h].y;
        acc[h][2] = hi[h].x;
        acc[h][3] = hi[h].y;
      }
    };
    if (MODE == 0 && it == 0)
    {
      // priming pass over the CTA's first tile: only the running minima (warp reduction, one atomic per warp)
      for (int q = 0; q < gqp; q += kStQPT)
      {
        float acc[kStQPT][kNodesPerThread];
        trip_values(q, acc);
#pragma unroll
        for (int h = 0; h < kStQPT; h++)
        {
          const float tmin = fminf(fminf(acc[h][0], acc[h][1]), fminf(acc[h][2], acc[h][3]));
          const uint32_t wmin = __reduce_min_sync(0xffffffffu, fkey(tmin));  // NaN keys sort above +inf
          if ((tid & 31) == 0 && wmin < m.curkey[q + h])
            atomicMin(&m.curkey[q + h], wmin);
        }
      }
      __syncthreads();
      exchange();
      __syncthreads();
    }
    for (int q = 0; q < gqp; q += kStQPT)
    {
      float acc[kStQPT][kNodesPerThread];
      trip_values(q, acc);
      const float4 th4 = *reinterpret_cast<const float4*>(m.thr + q);
      const float tha[4] = { th4.x, th4.y, th4.z, th4.w };
      float tmin[kStQPT];
      bool any = false;
#pragma unroll
      for (int h = 0; h < kStQPT; h++)
      {
        tmin[h] = fminf(fminf(acc[h][0], acc[h][1]), fminf(acc[h][2], acc[h][3]));
        any |= (tmin[h] <= tha[h]);
      }
      if (any)  // rare: at least one of this thread's 16 values is a candidate
      {
#pragma unroll
        for (int h = 0; h < kStQPT; h++)
        {
          float thr = tha[h];
          if (tmin[h] <= thr && tmin[h] < CUDART_INF_F)
          {
            const int qq = q + h;
            if (MODE == 0)
            {
              const uint32_t key = fkey(tmin[h]);
              if (key < m.curkey[qq] && key < atomicMin(&m.curkey[qq], key))
              {
                thr = stream_thr(tmin[h], m.e2[qq], a.slack_abs);
                thr_v[qq] = thr;  // a racing larger value is still a valid (looser) bound
              }
#pragma unroll
              for (int r = 0; r < kNodesPerThread; r++)
                if (acc[h][r] <= thr)
                {
                  const uint32_t pos = atomicAdd(m.cnt, 1u);
                  if (pos < (uint32_t)kStList)
                  {
                    m.lq[pos] = (uint32_t)qq;
                    m.ls[pos] = local0 + r;
                    m.le[pos] = (double)acc[h][r];
                  }
                }
            }
            else
            {
#pragma unroll
              for (int r = 0; r < kNodesPerThread; r++)
                if (acc[h][r] <= thr && acc[h][r] < CUDART_INF_F)
                {
                  if (MODE == 2)
                  {
                    // k-nearest collect: the exact distance is only needed for the few candidates next to the
                    // k-th one (select_topk_f32_kernel); here the candidate keeps its fp32 value
                    const uint32_t pos = atomicAdd(&a.out_count[q0 + qq], 1u);
                    if (pos < a.cap)
                    {
                      a.out_idx[(size_t)(q0 + qq) * a.cap + pos] = local0 + r;
                      a.out_c2[(size_t)(q0 + qq) * a.cap + pos] = acc[h][r];
                    }
                  }
                  else
                  {
                    const double e = dist64<D>(a.x64 + (size_t)(local0 + r) * D, m.q64 + qq * D);
                    if (e < a.radius[q0 + qq])
                    {
                      const uint32_t pos = atomicAdd(&a.out_count[q0 + qq], 1u);
                      if (pos < a.cap)
                      {
                        a.out_idx[(size_t)(q0 + qq) * a.cap + pos] = local0 + r;
                        a.out_dist[(size_t)(q0 + qq) * a.cap + pos] = e;
                      }
                    }
                  }
                }
            }
          }
        }
      }
    }
    if (MODE == 0)
    {
      __syncthreads();
      const uint32_t cnt = m.cnt[0];
      const bool last = (tile_of[(it + 1) % kStStages] == kStNoTile);
      if (it < 4 || (it & 7u) == 7u || last)
        exchange();
      // the candidate list is kept across tiles and resolved when it is half full or at the end: by then the
      // (shared) thresholds are much tighter than when the entries were made and most of them drop out
      if (cnt > (uint32_t)kStList / 2 || (last && cnt > 0))  // CTA-uniform
      {
        __syncthreads();
        const uint32_t nl = min(cnt, (uint32_t)kStList);
        // A: exact distances of the surviving candidates -> running exact minimum per query
        for (uint32_t i = tid; i < nl; i += kStThreads)
        {
          const uint32_t q = m.lq[i];
          double e = CUDART_INF;
          if ((float)m.le[i] <= thr_v[q])
          {
            e = dist64<D>(a.x64 + (size_t)m.ls[i] * D, m.q64 + q * D);
            atomicMin(&m.best64[q], (unsigned long long)__double_as_longlong(e));
          }
          else
            m.ls[i] = GC_NO_NODE;
          m.le[i] = e;
        }
        double eo[kNodesPerThread];
        bool co[kNodesPerThread];
        if (cnt > (uint32_t)kStList)
        {
          // the list overflowed inside this tile (massively duplicated points): exact pass over the tile's
          // nodes, query by query (entries that did fit are handled twice, harmlessly)
          for (int q = 0; q < gq; q++)
          {
            const float qn = m.qn[q];
            const float thr = thr_v[q];
#pragma unroll
            for (int r = 0; r < kNodesPerThread; r++)
            {
              float acc = xn[r] + qn;
#pragma unroll
              for (int d = 0; d < D; d++)
                acc = fmaf(x[d][r], m.qm2[q * DP + d], acc);
              co[r] = (acc <= thr) && (acc < CUDART_INF_F);
              eo[r] = 0.0;
              if (co[r])
              {
                eo[r] = dist64<D>(a.x64 + (size_t)(local0 + r) * D, m.q64 + q * D);
                atomicMin(&m.best64[q], (unsigned long long)__double_as_longlong(eo[r]));
              }
            }
            __syncthreads();
            if (tid == 0 && m.best64[q] != m.prev64[q])
            {
              m.bestidx[q] = GC_NO_NODE;
              m.prev64[q] = m.best64[q];
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < kNodesPerThread; r++)
              if (co[r] && (unsigned long long)__double_as_longlong(eo[r]) == m.best64[q])
                atomicMin(&m.bestidx[q], local0 + r);
          }
        }
        __syncthreads();
        // B: a strictly smaller exact minimum invalidates the slot kept for the old one
        if (tid < gq && m.best64[tid] != m.prev64[tid])
        {
          m.bestidx[tid] = GC_NO_NODE;
          m.prev64[tid] = m.best64[tid];
        }
        __syncthreads();
        // C: lowest slot among the candidates that attain the exact minimum
        for (uint32_t i = tid; i < nl; i += kStThreads)
        {
          const uint32_t q = m.lq[i];
          if (m.ls[i] != GC_NO_NODE && (unsigned long long)__double_as_longlong(m.le[i]) == m.best64[q])
            atomicMin(&m.bestidx[q], m.ls[i]);
        }
        __syncthreads();
        if (tid == 0)
          m.cnt[0] = 0;
        // the next tile's first list append happens after that tile's own barriers
      }
    }
  }
  // ---- epilogue: per-CTA partials, last-CTA-done reduction, counter reset ---------------------------------
  __syncthreads();
  if (MODE == 0 && tid < gq)
  {
    a.part_d[(size_t)blockIdx.x * a.nq + q0 + tid] = __longlong_as_double((long long)m.best64[tid]);
    a.part_i[(size_t)blockIdx.x * a.nq + q0 + tid] = m.bestidx[tid];
  }
  __threadfence();
  __syncthreads();
  if (tid == 0)
    m.cnt[1] = (atomicAdd(done_ctr, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (m.cnt[1])
  {
    __threadfence();
    if (MODE == 0)
    {
      // lexicographic (dist, slot) minimum over the CTAs of this group: one warp per query, lanes over CTAs
      const int lane = tid & 31, warp = tid >> 5;
      for (int q = warp; q < gq; q += kStThreads / 32)
      {
        double bd = CUDART_INF;
        uint32_t bi = GC_NO_NODE;
        for (uint32_t c = lane; c < gridDim.x; c += 32)
        {
          const double d = __ldcg(a.part_d + (size_t)c * a.nq + q0 + q);
          const uint32_t i = __ldcg(a.part_i + (size_t)c * a.nq + q0 + q);
          if (i != GC_NO_NODE && (d < bd || (d == bd && i < bi)))
          {
            bd = d;
            bi = i;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
          const double od = __shfl_xor_sync(0xffffffffu, bd, o);
          const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (oi != GC_NO_NODE && (od < bd || (od == bd && oi < bi)))
          {
            bd = od;
            bi = oi;
          }
        }
        if (lane == 0)
        {
          a.out_idx[q0 + q] = bi;
          a.out_dist[q0 + q] = bd;
          a.gmin[q0 + q] = 0xFFFFFFFFu;  // ready for the next launch
        }
      }
    }
    if (tid == 0)
    {
      *tile_ctr = 0;
      *done_ctr = 0;
    }
  }
}

// lexicographic (dist, slot) minimum over the chunk partials
__global__ void nn1_finalize_kernel(const double* __restrict__ part_d, const uint32_t* __restrict__ part_i,
                                    uint32_t chunks, uint32_t nq, uint32_t* __restrict__ idx,
                                    double* __restrict__ dist)
{
  // one warp per query: lanes stride over the chunks, then a shuffle reduction (independent loads in flight)
  const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= nq)
    return;
  double bd = CUDART_INF;
  uint32_t bi = GC_NO_NODE;
  for (uint32_t c = lane; c < chunks; c += 32)
  {
    const double d = part_d[(size_t)c * nq + q];
    const uint32_t i = part_i[(size_t)c * nq + q];
    if (i != GC_NO_NODE && (d < bd || (d == bd && i < bi)))
    {
      bd = d;
      bi = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    const double od = __shfl_xor_sync(0xffffffffu, bd, o);
    const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi != GC_NO_NODE && (od < bd || (od == bd && oi < bi)))
    {
      bd = od;
      bi = oi;
    }
  }
  if (lane == 0)
  {
    idx[q] = bi;
    dist[q] = bd;
  }
}

// ---------------------------------------------------------------------------------------------
// small segments (the lock-step planner regime: thousands of trees of 10^2..10^3 nodes): one warp per query walks its
// segment with the exact fp64 distance straight away -- a tile of 1024 slots per CTA (scan_kernel) would leave most
// lanes idle and pay a block-wide fp32 -> fp64 hand-over per query.  Same results by construction: nearest = the
// lexicographic (dist64, slot) minimum over the live slots, radius = every live slot with dist64 < r, the list ordered
// by (dist64, slot) -- the order sort_lists_kernel produces.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kSmallSegMax = 1024;   // segments up to this many slots take the warp-per-query kernels
constexpr uint32_t kSmallCapMax = 1024;   // radius lists up to this many entries are ordered inside the kernel
constexpr int kSmallWarps = 4;

template <int D, int MODE>
__global__ void __launch_bounds__(kSmallWarps * 32) small_scan_kernel(ScanArgs a)
{
  extern __shared__ __align__(16) unsigned char small_smem[];
  const uint32_t wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t q = blockIdx.x * kSmallWarps + wid;
  if (q >= a.nq)
    return;
  const uint32_t seg = a.seg ? a.seg[q] : 0u;
  const uint32_t n = a.used[seg];
  const size_t base = (size_t)seg * a.cap_seg;
  double qv[D];
#pragma unroll
  for (int d = 0; d < D; d++)
    qv[d] = a.q[(size_t)q * D + d];
  const float* x0 = a.x32 + base;  // first coordinate row: +inf marks a tombstone
  if (MODE == 0)
  {
    double bd = CUDART_INF;
    uint32_t bi = GC_NO_NODE;
    for (uint32_t slot = lane; slot < n; slot += 32)
    {
      if (!(x0[slot] < CUDART_INF_F))
        continue;
      const double e = dist64<D>(a.x64 + (base + slot) * D, qv);
      if (e < bd)
      {
        bd = e;
        bi = slot;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      const double od = __shfl_xor_sync(0xffffffffu, bd, o);
      const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi != GC_NO_NODE && (od < bd || (od == bd && oi < bi)))
      {
        bd = od;
        bi = oi;
      }
    }
    if (lane == 0)
    {
      a.out_idx[q] = bi;
      a.out_dist[q] = bd;
    }
    return;
  }
  // radius: hits staged in slot order, then rank-sorted by (dist, slot) into the caller's list
  double* sd = reinterpret_cast<double*>(small_smem) + (size_t)wid * a.cap;
  uint32_t* si = reinterpret_cast<uint32_t*>(reinterpret_cast<double*>(small_smem) + (size_t)kSmallWarps * a.cap) +
                 (size_t)wid * a.cap;
  const double r = a.radius[q];
  uint32_t count = 0;
  for (uint32_t s0 = 0; s0 < n; s0 += 32)
  {
    const uint32_t slot = s0 + lane;
    bool hit = false;
    double e = 0.0;
    if (slot < n && x0[slot] < CUDART_INF_F)
    {
      e = dist64<D>(a.x64 + (base + slot) * D, qv);
      hit = e < r;  // strict (vector.cpp:75 / kdtree.cpp:167)
    }
    const uint32_t m = __ballot_sync(0xffffffffu, hit);
    const uint32_t pos = count + __popc(m & ((1u << lane) - 1u));
    if (hit && pos < a.cap)
    {
      sd[pos] = e;
      si[pos] = slot;
    }
    count += __popc(m);
  }
  __syncwarp();
  const uint32_t nk = min(count, a.cap);
  for (uint32_t i = lane; i < nk; i += 32)
  {
    const double di = sd[i];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < nk; j++)
    {
      const double dj = sd[j];
      rank += (dj < di || (dj == di && j < i)) ? 1u : 0u;  // staged in ascending slot order
    }
    a.out_idx[(size_t)q * a.cap + rank] = si[i];
    a.out_dist[(size_t)q * a.cap + rank] = di;
  }
  if (lane == 0)
    a.out_count[q] = count;
}

// ---------------------------------------------------------------------------------------------
// block-wide bitonic sort of (key, slot) pairs, ascending lexicographic; P = power of two
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bitonic_sort(double* key, uint32_t* val, uint32_t P)
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
          const bool up = ((i & k) == 0);
          if (gt == up)
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

// sorts the first min(count, cap) entries of every query's list in place
__global__ void sort_lists_kernel(uint32_t* __restrict__ idx, double* __restrict__ dist,
                                  const uint32_t* __restrict__ count, uint32_t cap, uint32_t P)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* key = reinterpret_cast<double*>(smem_raw);
  uint32_t* val = reinterpret_cast<uint32_t*>(key + P);
  const uint32_t q = blockIdx.x;
  const uint32_t n = min(count[q], cap);
  if (n <= 1)
    return;
  uint32_t p2 = 2;
  while (p2 < n)
    p2 <<= 1;
  for (uint32_t i = threadIdx.x; i < p2; i += blockDim.x)
  {
    key[i] = (i < n) ? dist[(size_t)q * cap + i] : CUDART_INF;
    val[i] = (i < n) ? idx[(size_t)q * cap + i] : GC_NO_NODE;
  }
  __syncthreads();
  bitonic_sort(key, val, p2);
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
  {
    dist[(size_t)q * cap + i] = key[i];
    idx[(size_t)q * cap + i] = val[i];
  }
}

// ---------------------------------------------------------------------------------------------
// Lists longer than one CTA can sort (kSortMax entries): runs of kSortMax entries are sorted in shared
// memory, then merged pairwise in global memory.  A merge pass is rank based: entry i of the left run
// lands at i + (entries of the right run strictly before it), entry j of the right run at j + (entries
// of the left run not after it), each found by a binary search -- every entry is independent, so a
// pass is one launch over all lists however long they are.  Order: ascending (dist, slot), the
// multimap order of KdTree / Vector results (kdtree.cpp:362-370, vector.cpp:83-96).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool pair_before(double da, uint32_t sa, double db, uint32_t sb)
{
  return da < db || (da == db && sa < sb);
}

// `count` == nullptr: every list holds `cap` entries.  `stride` = entries between consecutive lists.
__global__ void __launch_bounds__(1024, 1)
    sort_runs_kernel(uint32_t* __restrict__ idx, double* __restrict__ dist, const uint32_t* __restrict__ count,
                     uint32_t cap, size_t stride)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* key = reinterpret_cast<double*>(smem_raw);
  uint32_t* val = reinterpret_cast<uint32_t*>(key + kSortMax);
  const uint32_t q = blockIdx.y;
  const uint32_t n = count ? min(count[q], cap) : cap;
  const uint32_t r0 = blockIdx.x * kSortMax;
  if (r0 >= n)
    return;
  const uint32_t m = min(kSortMax, n - r0);
  uint32_t p2 = 2;
  while (p2 < m)
    p2 <<= 1;
  uint32_t* li = idx + (size_t)q * stride + r0;
  double* ld = dist + (size_t)q * stride + r0;
  for (uint32_t i = threadIdx.x; i < p2; i += blockDim.x)
  {
    key[i] = (i < m) ? ld[i] : CUDART_INF;
    val[i] = (i < m) ? li[i] : GC_NO_NODE;
  }
  __syncthreads();
  bitonic_sort(key, val, p2);
  for (uint32_t i = threadIdx.x; i < m; i += blockDim.x)
  {
    ld[i] = key[i];
    li[i] = val[i];
  }
}

__global__ void __launch_bounds__(256)
    merge_runs_kernel(const uint32_t* __restrict__ idx_in, const double* __restrict__ dist_in,
                      uint32_t* __restrict__ idx_out, double* __restrict__ dist_out,
                      const uint32_t* __restrict__ count, uint32_t cap, size_t stride, uint32_t L)
{
  const uint32_t q = blockIdx.y;
  const uint32_t n = count ? min(count[q], cap) : cap;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const uint32_t* li = idx_in + (size_t)q * stride;
  const double* ld = dist_in + (size_t)q * stride;
  const double d = ld[i];
  const uint32_t sl = li[i];
  const uint32_t run = i / L;
  const uint32_t pair_base = (run & ~1u) * L;
  const bool left = (run & 1u) == 0;
  const uint32_t sib_base = left ? pair_base + L : pair_base;
  uint32_t rank = 0;
  if (sib_base < n)
  {
    const uint32_t sib_n = min(L, n - sib_base);
    uint32_t lo = 0, hi = sib_n;
    while (lo < hi)
    {
      const uint32_t mid = (lo + hi) >> 1;
      const double dm = ld[sib_base + mid];
      const uint32_t sm = li[sib_base + mid];
      // left entries go before equal right entries, so the pass is a permutation even with repeated keys
      const bool sib_first = left ? pair_before(dm, sm, d, sl) : !pair_before(d, sl, dm, sm);
      if (sib_first)
        lo = mid + 1;
      else
        hi = mid;
    }
    rank = lo;
  }
  const size_t o = (size_t)q * stride + pair_base + (i - run * L) + rank;
  idx_out[o] = sl;
  dist_out[o] = d;
}

// exact fp64 distance of every slot of the query's segment (tombstones and the padding: +inf, GC_NO_NODE); the
// k >= N "all nodes, sorted" answer of Tree::nearK (tree.cpp:761-765) and k > kSortMax on large trees
template <int D>
__global__ void __launch_bounds__(256)
    dist_all_kernel(const float* __restrict__ x32, const double* __restrict__ x64, const uint32_t* __restrict__ used,
                    const double* __restrict__ qd, const uint32_t* __restrict__ segs, uint32_t cap_seg, uint32_t npad,
                    uint32_t* __restrict__ out_idx, double* __restrict__ out_dist, uint32_t* __restrict__ live_count)
{
  __shared__ double qs[D];
  const uint32_t q = blockIdx.y;
  if (threadIdx.x < D)
    qs[threadIdx.x] = qd[(size_t)q * D + threadIdx.x];
  __syncthreads();
  const uint32_t seg = segs ? segs[q] : 0u;
  const uint32_t n_s = used[seg];
  const size_t seg_base = (size_t)seg * cap_seg;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool live = false;
  double e = CUDART_INF;
  if (i < n_s && x32[seg_base + i] < CUDART_INF_F)
  {
    e = dist64<D>(x64 + (seg_base + i) * D, qs);
    live = true;
  }
  if (i < npad)
  {
    out_idx[(size_t)q * npad + i] = live ? i : GC_NO_NODE;
    out_dist[(size_t)q * npad + i] = e;
  }
  const uint32_t m = __ballot_sync(0xffffffffu, live);
  if ((threadIdx.x & 31) == 0 && m)
    atomicAdd(&live_count[q], (uint32_t)__popc(m));
}

// first min(k, have, out_cap) entries of every sorted list; out_count = min(k, have)
__global__ void take_first_kernel(const uint32_t* __restrict__ idx, const double* __restrict__ dist,
                                  const uint32_t* __restrict__ have, size_t stride, unsigned long long k,
                                  uint32_t out_cap, uint32_t* __restrict__ out_idx, double* __restrict__ out_dist,
                                  uint32_t* __restrict__ out_count)
{
  const uint32_t q = blockIdx.y;
  const uint32_t h = have[q];
  const uint32_t want = (k < (unsigned long long)h) ? (uint32_t)k : h;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < min(want, out_cap))
  {
    out_idx[(size_t)q * out_cap + i] = idx[(size_t)q * stride + i];
    out_dist[(size_t)q * out_cap + i] = dist[(size_t)q * stride + i];
  }
  if (i == 0)
    out_count[q] = want;
}

// shard lists [G][nq][cap] -> one list per query with global slots (local * G + g), `total` = entries gathered
__global__ void gather_shard_lists_kernel(const uint32_t* __restrict__ idx, const double* __restrict__ dist,
                                          const uint32_t* __restrict__ count, uint32_t G, uint32_t nq, uint32_t cap,
                                          uint32_t* __restrict__ out_idx, double* __restrict__ out_dist,
                                          uint32_t* __restrict__ total)
{
  const uint32_t q = blockIdx.y;
  uint32_t base = 0;
  for (uint32_t g = 0; g < G; g++)
  {
    const uint32_t n = min(count[(size_t)g * nq + q], cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
      const uint32_t local = idx[((size_t)g * nq + q) * cap + i];
      out_idx[(size_t)q * G * cap + base + i] = (local == GC_NO_NODE) ? GC_NO_NODE : local * G + g;
      out_dist[(size_t)q * G * cap + base + i] = dist[((size_t)g * nq + q) * cap + i];
    }
    base += n;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    total[q] = base;
}

// k-nearest for segments of at most kSortMax slots (planner regime; also the k >= N "return all,
// sorted" case of Tree::nearK, src/graph_core/graph/tree.cpp:761-765): exact fp64 distance of
// every live slot, block-wide sort, first min(k, live) entries.
template <int D>
__global__ void __launch_bounds__(1024, 1)
    knn_all_kernel(const float* __restrict__ x32, const double* __restrict__ x64, const uint32_t* __restrict__ used,
                   const double* __restrict__ qd, const uint32_t* __restrict__ segs, uint32_t cap_seg,
                   size_t row_stride, unsigned long long k, uint32_t cap, uint32_t* __restrict__ out_idx,
                   double* __restrict__ out_dist, uint32_t* __restrict__ out_count, uint32_t P)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* key = reinterpret_cast<double*>(smem_raw);
  uint32_t* val = reinterpret_cast<uint32_t*>(key + P);
  __shared__ double qs[D];
  __shared__ uint32_t live_s;
  const uint32_t q = blockIdx.x;
  const uint32_t seg = segs ? segs[q] : 0u;
  const uint32_t n_s = used[seg];
  const size_t seg_base = (size_t)seg * cap_seg;
  if (threadIdx.x < D)
    qs[threadIdx.x] = qd[(size_t)q * D + threadIdx.x];
  if (threadIdx.x == 0)
    live_s = 0;
  __syncthreads();
  uint32_t p2 = 2;
  while (p2 < n_s)
    p2 <<= 1;
  uint32_t my_live = 0;
  for (uint32_t i = threadIdx.x; i < p2; i += blockDim.x)
  {
    double e = CUDART_INF;
    uint32_t v = GC_NO_NODE;
    if (i < n_s && x32[seg_base + i] < CUDART_INF_F)  // row 0 of the SoA carries the tombstone mark
    {
      e = dist64<D>(x64 + (seg_base + i) * D, qs);
      v = i;
      my_live++;
    }
    key[i] = e;
    val[i] = v;
  }
  if (my_live)
    atomicAdd(&live_s, my_live);
  __syncthreads();
  bitonic_sort(key, val, p2);
  const uint32_t live = live_s;
  const uint32_t want = (k < (unsigned long long)live) ? (uint32_t)k : live;
  for (uint32_t i = threadIdx.x; i < min(want, cap); i += blockDim.x)
  {
    out_idx[(size_t)q * cap + i] = val[i];
    out_dist[(size_t)q * cap + i] = key[i];
  }
  if (threadIdx.x == 0)
    out_count[q] = want;
}

// ---------------------------------------------------------------------------------------------
// k-nearest over segments larger than one CTA can sort (BASELINE config 4: 2^20 nodes):
//   1. knn_sample_kernel : exact k-th smallest fp32 distance over a 16384-node SAMPLE of the segment.
//      The sample is a subset, so its k-th smallest is an upper bound of the segment's k-th smallest:
//      every true k-nearest node lies under the (slack-widened) threshold.
//   2. scan_kernel<MODE 2>: one brute-force pass collecting the nodes under the threshold with their
//      exact fp64 distances (a few k entries per query instead of N).
//   3. select_topk_kernel: exact radix select of the k smallest (dist64, slot) keys, then a sort.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bitonic_sort_keys(float* key, uint32_t P)
{
  for (uint32_t k = 2; k <= P; k <<= 1)
    for (uint32_t j = k >> 1; j > 0; j >>= 1)
    {
      for (uint32_t i = threadIdx.x; i < P; i += blockDim.x)
      {
        const uint32_t ixj = i ^ j;
        if (ixj > i)
        {
          const float ka = key[i], kb = key[ixj];
          if ((ka > kb) == ((i & k) == 0))
          {
            key[i] = kb;
            key[ixj] = ka;
          }
        }
      }
      __syncthreads();
    }
}

// sample slot i of a segment with n_s used slots: everything when n_s <= kSortMax, else 16 evenly
// spread contiguous blocks of 1024 slots (coalesced reads, spread over the insertion order)
__device__ __forceinline__ uint32_t sample_slot(uint32_t i, uint32_t n_s)
{
  if (n_s <= kSortMax)
    return i;
  const uint32_t b = i >> 10, o = i & 1023u;
  return b * (n_s / 16u) + o;
}

template <int D>
__global__ void __launch_bounds__(1024, 1)
    knn_sample_kernel(const float* __restrict__ x32, const uint32_t* __restrict__ used, const double* __restrict__ qd,
                      const uint32_t* __restrict__ segs, uint32_t cap_seg, size_t row_stride, unsigned long long k,
                      float slack_abs, float* __restrict__ thr2_out)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* key = reinterpret_cast<float*>(smem_raw);
  __shared__ float qs[D];
  __shared__ uint32_t live_s;
  const uint32_t q = blockIdx.x;
  const uint32_t seg = segs ? segs[q] : 0u;
  const uint32_t n_s = used[seg];
  const size_t seg_base = (size_t)seg * cap_seg;
  if (threadIdx.x < D)
    qs[threadIdx.x] = (float)qd[(size_t)q * D + threadIdx.x];
  if (threadIdx.x == 0)
    live_s = 0;
  __syncthreads();
  const uint32_t m = n_s < kSortMax ? n_s : kSortMax;
  uint32_t p2 = 2;
  while (p2 < m)
    p2 <<= 1;
  uint32_t my_live = 0;
  for (uint32_t i = threadIdx.x; i < p2; i += blockDim.x)
  {
    float acc = CUDART_INF_F;
    if (i < m)
    {
      const size_t slot = seg_base + sample_slot(i, n_s);
      acc = 0.0f;
#pragma unroll
      for (int d = 0; d < D; d++)
      {
        const float t = x32[(size_t)d * row_stride + slot] - qs[d];
        acc = fmaf(t, t, acc);
      }
      if (acc < CUDART_INF_F)
        my_live++;
      else
        acc = CUDART_INF_F;  // tombstone (x32[0] = +inf) or NaN
    }
    key[i] = acc;
  }
  if (my_live)
    atomicAdd(&live_s, my_live);
  __syncthreads();
  bitonic_sort_keys(key, p2);
  if (threadIdx.x == 0)
  {
    float thr = FLT_MAX;  // fewer than k live nodes in the sample: no bound, collect everything
    if (k >= 1 && k <= (unsigned long long)live_s)
      thr = slack_thr2(sqrtf(key[k - 1]) * (1.0f + 1.0e-6f), slack_abs);
    thr2_out[q] = thr;
  }
}

// block-wide exact radix select: returns (in shared result slots) the value of the k-th smallest key
// among the items i < n for which pred(i) holds; keys are unsigned 64-bit.  k is 1-based, k <= #pred.
template <typename KeyFn, typename PredFn>
__device__ __forceinline__ unsigned long long radix_select(uint32_t n, unsigned long long k, int bytes, KeyFn keyf,
                                                           PredFn pred, uint32_t* hist, unsigned long long* sh,
                                                           unsigned long long* k_rest)
{
  unsigned long long prefix = 0, mask = 0;
  for (int r = bytes - 1; r >= 0; r--)
  {
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x)
      hist[i] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
      if (pred(i))
      {
        const unsigned long long key = keyf(i);
        if ((key & mask) == prefix)
          atomicAdd(&hist[(uint32_t)(key >> (8 * r)) & 255u], 1u);
      }
    __syncthreads();
    // bin holding the k-th key: inclusive scan of the 256 counts by the first 8 warps (was a 255-step serial walk by one
    // thread, the largest single cost of the single-CTA select kernels); blockDim >= 256 at every call site
    {
      __shared__ uint32_t warp_tot[8];
      const uint32_t tid = threadIdx.x, lane = tid & 31u;
      uint32_t c = 0, incl = 0;
      if (tid < 256)
      {
        c = hist[tid];
        incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
          const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= (uint32_t)o)
            incl += v;
        }
        if (lane == 31)
          warp_tot[tid >> 5] = incl;
      }
      __syncthreads();
      if (tid < 256)
      {
        unsigned long long off = 0;
        for (uint32_t w = 0; w < (tid >> 5); w++)
          off += warp_tot[w];
        const unsigned long long in = off + incl, ex = in - c;
        // the last bin takes whatever is left (k <= number of eligible keys by contract)
        if (ex < k && (k <= in || tid == 255))
        {
          sh[0] = (unsigned long long)tid;
          sh[1] = k - ex;
        }
      }
    }
    __syncthreads();
    prefix |= sh[0] << (8 * r);
    mask |= 0xffull << (8 * r);
    k = sh[1];
    __syncthreads();
  }
  *k_rest = k;  // how many items with exactly this key are needed
  return prefix;
}

// lists [nq][cap] with counts -> the k smallest (dist, slot) per query, ascending, into [nq][out_cap]
__global__ void __launch_bounds__(1024, 1)
    select_topk_kernel(const uint32_t* __restrict__ idx, const double* __restrict__ dist,
                       const uint32_t* __restrict__ count, uint32_t cap, unsigned long long k, uint32_t out_cap,
                       uint32_t* __restrict__ out_idx, double* __restrict__ out_dist, uint32_t* __restrict__ out_count,
                       uint32_t* __restrict__ err_flag, uint32_t P)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* skey = reinterpret_cast<double*>(smem_raw);
  uint32_t* sval = reinterpret_cast<uint32_t*>(skey + P);
  __shared__ uint32_t hist[256];
  __shared__ unsigned long long sh[2];
  __shared__ uint32_t npicked;
  const uint32_t q = blockIdx.x;
  uint32_t n = count[q];
  if (n > cap)
  {
    if (threadIdx.x == 0)
      atomicOr(err_flag, 1u);  // the candidate list overflowed: the caller retries with a larger one
    n = cap;
  }
  const uint32_t* li = idx + (size_t)q * cap;
  const double* ld = dist + (size_t)q * cap;
  const uint32_t want = (k < (unsigned long long)n) ? (uint32_t)k : n;
  uint32_t m = 0;  // entries staged in shared memory
  if (n <= P)
  {
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
    {
      skey[i] = ld[i];
      sval[i] = li[i];
    }
    m = n;
  }
  else
  {
    // exact k-th smallest distance, then the slot rank among equal distances
    unsigned long long need_eq = 0, need_slot = 0;
    const unsigned long long kth = radix_select(
        n, want, 8, [&](uint32_t i) { return (unsigned long long)__double_as_longlong(ld[i]); },
        [&](uint32_t) { return true; }, hist, sh, &need_eq);
    const unsigned long long slot_max = radix_select(
        n, need_eq, 4, [&](uint32_t i) { return (unsigned long long)li[i]; },
        [&](uint32_t i) { return (unsigned long long)__double_as_longlong(ld[i]) == kth; }, hist, sh, &need_slot);
    if (threadIdx.x == 0)
      npicked = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
    {
      const unsigned long long key = (unsigned long long)__double_as_longlong(ld[i]);
      if (key < kth || (key == kth && (unsigned long long)li[i] <= slot_max))
      {
        const uint32_t pos = atomicAdd(&npicked, 1u);
        if (pos < P)
        {
          skey[pos] = ld[i];
          sval[pos] = li[i];
        }
      }
    }
    __syncthreads();
    m = npicked < P ? npicked : P;
  }
  uint32_t p2 = 2;
  while (p2 < m)
    p2 <<= 1;
  for (uint32_t i = m + threadIdx.x; i < p2; i += blockDim.x)
  {
    skey[i] = CUDART_INF;
    sval[i] = GC_NO_NODE;
  }
  __syncthreads();
  bitonic_sort(skey, sval, p2);
  for (uint32_t i = threadIdx.x; i < (want < out_cap ? want : out_cap); i += blockDim.x)
  {
    out_idx[(size_t)q * out_cap + i] = sval[i];
    out_dist[(size_t)q * out_cap + i] = skey[i];
  }
  if (threadIdx.x == 0)
    out_count[q] = want;
}

// k-th smallest fp32 distance of the sample by radix select (4 histogram passes) instead of a full sort
template <int D>
__global__ void __launch_bounds__(1024, 1)
    knn_sample_select_kernel(const float* __restrict__ x32, const uint32_t* __restrict__ used,
                             const double* __restrict__ qd, const uint32_t* __restrict__ segs, uint32_t cap_seg,
                             size_t row_stride, unsigned long long k, float slack_abs, float* __restrict__ thr2_out)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* key = reinterpret_cast<float*>(smem_raw);
  __shared__ float qs[D];
  __shared__ uint32_t live_s;
  __shared__ uint32_t hist[256];
  __shared__ unsigned long long sh[2];
  const uint32_t q = blockIdx.x;
  const uint32_t seg = segs ? segs[q] : 0u;
  const uint32_t n_s = used[seg];
  const size_t seg_base = (size_t)seg * cap_seg;
  if (threadIdx.x < D)
    qs[threadIdx.x] = (float)qd[(size_t)q * D + threadIdx.x];
  if (threadIdx.x == 0)
    live_s = 0;
  __syncthreads();
  const uint32_t m = n_s < kSortMax ? n_s : kSortMax;
  uint32_t my_live = 0;
  for (uint32_t i = threadIdx.x; i < m; i += blockDim.x)
  {
    const size_t slot = seg_base + sample_slot(i, n_s);
    float acc = 0.0f;
#pragma unroll
    for (int d = 0; d < D; d++)
    {
      const float t = x32[(size_t)d * row_stride + slot] - qs[d];
      acc = fmaf(t, t, acc);
    }
    if (acc < CUDART_INF_F)
      my_live++;
    else
      acc = CUDART_INF_F;  // tombstone (x32[0] = +inf) or NaN
    key[i] = acc;
  }
  if (my_live)
    atomicAdd(&live_s, my_live);
  __syncthreads();
  const uint32_t live = live_s;
  float thr = FLT_MAX;  // fewer than k live nodes in the sample: no bound, collect everything
  if (k >= 1 && k <= (unsigned long long)live)
  {
    unsigned long long rest = 0;
    const unsigned long long kk = radix_select(
        m, k, 4, [&](uint32_t i) { return (unsigned long long)fkey(key[i]); },
        [&](uint32_t i) { return key[i] < CUDART_INF_F; }, hist, sh, &rest);
    thr = slack_thr2(sqrtf(funkey((uint32_t)kk)) * (1.0f + 1.0e-6f), slack_abs);
  }
  if (threadIdx.x == 0)
    thr2_out[q] = thr;
}

// Candidate lists with fp32 values (scan_stream_kernel MODE 2) -> the k nearest, exactly.
// The k-th smallest fp32 value c_k bounds the true k-th distance from above (some node among the k smallest values
// is at least the k-th in truth), so every true k-nearest node has a value <= stream_thr(c_k): only those -- k plus
// a handful -- get their fp64 distance, are sorted by (dist64, slot) and the first k are the answer.
template <int D>
__global__ void __launch_bounds__(1024, 1)
    select_topk_f32_kernel(const uint32_t* __restrict__ idx, const float* __restrict__ c2,
                           const uint32_t* __restrict__ count, uint32_t cap, const double* __restrict__ x64,
                           const double* __restrict__ qd, unsigned long long k, float xn_max, float slack_abs,
                           uint32_t out_cap, uint32_t* __restrict__ out_idx, double* __restrict__ out_dist,
                           uint32_t* __restrict__ out_count, uint32_t* __restrict__ err_flag, uint32_t P)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* skey = reinterpret_cast<double*>(smem_raw);
  uint32_t* sval = reinterpret_cast<uint32_t*>(skey + P);
  __shared__ uint32_t hist[256];
  __shared__ unsigned long long sh[2];
  __shared__ uint32_t npicked;
  __shared__ double qs[D];
  const uint32_t q = blockIdx.x;
  uint32_t n = count[q];
  if (n > cap)
  {
    if (threadIdx.x == 0)
      atomicOr(err_flag, 1u);  // the candidate list overflowed: the caller retries with a larger one
    n = cap;
  }
  const uint32_t* li = idx + (size_t)q * cap;
  const float* lc = c2 + (size_t)q * cap;
  if (threadIdx.x < D)
    qs[threadIdx.x] = qd[(size_t)q * D + threadIdx.x];
  if (threadIdx.x == 0)
    npicked = 0;
  __syncthreads();
  const uint32_t want = (k < (unsigned long long)n) ? (uint32_t)k : n;
  float thr = CUDART_INF_F;
  if (want < n)
  {
    unsigned long long rest = 0;
    const unsigned long long kk = radix_select(
        n, want, 4, [&](uint32_t i) { return (unsigned long long)fkey(lc[i]); }, [&](uint32_t) { return true; }, hist,
        sh, &rest);
    float qn = 0.0f;
#pragma unroll
    for (int d = 0; d < D; d++)
    {
      const float v = (float)qs[d];
      qn = fmaf(v, v, qn);
    }
    const float e2 = (float)(3 * D + 8) * 5.9604645e-8f * 1.5f * (xn_max + qn) * 2.0f + FLT_MIN;
    thr = stream_thr(funkey((uint32_t)kk), e2, slack_abs);
  }
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
    if (lc[i] <= thr)
    {
      const uint32_t pos = atomicAdd(&npicked, 1u);
      if (pos < P)
      {
        skey[pos] = dist64<D>(x64 + (size_t)li[i] * D, qs);
        sval[pos] = li[i];
      }
    }
  __syncthreads();
  uint32_t m = npicked;
  if (m > P)
  {
    if (threadIdx.x == 0)
      atomicOr(err_flag, 2u);  // more near-ties than the sort buffer holds: the caller takes the all-fp64 path
    m = P;
  }
  uint32_t p2 = 2;
  while (p2 < m)
    p2 <<= 1;
  for (uint32_t i = m + threadIdx.x; i < p2; i += blockDim.x)
  {
    skey[i] = CUDART_INF;
    sval[i] = GC_NO_NODE;
  }
  __syncthreads();
  bitonic_sort(skey, sval, p2);
  for (uint32_t i = threadIdx.x; i < (want < out_cap ? want : out_cap); i += blockDim.x)
  {
    out_idx[(size_t)q * out_cap + i] = sval[i];
    out_dist[(size_t)q * out_cap + i] = skey[i];
  }
  if (threadIdx.x == 0)
    out_count[q] = want;
}

// Second-phase threshold of the large-N k-nearest: after the candidates of the first N/16 nodes have been collected
// under the (loose) sample threshold, the k-th smallest of THEIR fp32 values bounds the true k-th distance much more
// tightly -- it is the k-th of a 16 times larger sample -- and the remaining 15/16 of the store is scanned under
// stream_thr() of it.  Fewer than k candidates so far: the old threshold stays.
template <int D>
__global__ void __launch_bounds__(1024, 1)
    knn_rethreshold_kernel(const float* __restrict__ c2, const uint32_t* __restrict__ count, uint32_t cap,
                           const double* __restrict__ qd, unsigned long long k, float xn_max, float slack_abs,
                           float* __restrict__ thr_io)
{
  __shared__ uint32_t hist[256];
  __shared__ unsigned long long sh[2];
  const uint32_t q = blockIdx.x;
  const uint32_t n = min(count[q], cap);
  const float* lc = c2 + (size_t)q * cap;
  float qn = 0.0f;
#pragma unroll
  for (int d = 0; d < D; d++)
  {
    const float v = (float)qd[(size_t)q * D + d];
    qn = fmaf(v, v, qn);
  }
  const float e2 = (float)(3 * D + 8) * 5.9604645e-8f * 1.5f * (xn_max + qn) * 2.0f + FLT_MIN;
  // the first phase ran under the sample threshold in its expanded-form version
  float thr = thr_io[q] * (1.0f + 8.0e-6f) + e2 * (1.0f + 1.0e-6f);
  if (k >= 1 && k <= (unsigned long long)n && count[q] <= cap)
  {
    unsigned long long rest = 0;
    const unsigned long long kk = radix_select(
        n, k, 4, [&](uint32_t i) { return (unsigned long long)fkey(lc[i]); }, [&](uint32_t) { return true; }, hist, sh,
        &rest);
    thr = fminf(thr, stream_thr(funkey((uint32_t)kk), e2, slack_abs));
  }
  if (threadIdx.x == 0)
    thr_io[q] = thr;
}

// Node-sharded trees: node i lives on shard i mod G at local slot i div G.  After the
// all-gather of the per-shard lists this merges them: global slot = local * G + g, ascending (dist, slot),
// first k_out entries.  parts are [G][nq][cap] / [G][nq].
__global__ void __launch_bounds__(1024, 1)
    merge_shard_lists_kernel(const uint32_t* __restrict__ idx, const double* __restrict__ dist,
                             const uint32_t* __restrict__ count, uint32_t G, uint32_t nq, uint32_t cap,
                             unsigned long long k_out, uint32_t out_cap, uint32_t* __restrict__ out_idx,
                             double* __restrict__ out_dist, uint32_t* __restrict__ out_count, uint32_t P)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* skey = reinterpret_cast<double*>(smem_raw);
  uint32_t* sval = reinterpret_cast<uint32_t*>(skey + P);
  __shared__ uint32_t total_s;
  const uint32_t q = blockIdx.x;
  if (threadIdx.x == 0)
    total_s = 0;
  __syncthreads();
  for (uint32_t g = 0; g < G; g++)
  {
    const uint32_t n = min(count[(size_t)g * nq + q], cap);
    __shared__ uint32_t base_s;
    if (threadIdx.x == 0)
    {
      base_s = total_s;
      total_s += n;
    }
    __syncthreads();
    const uint32_t base = base_s;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
      if (base + i < P)
      {
        const uint32_t local = idx[((size_t)g * nq + q) * cap + i];
        skey[base + i] = dist[((size_t)g * nq + q) * cap + i];
        sval[base + i] = (local == GC_NO_NODE) ? GC_NO_NODE : local * G + g;
      }
    __syncthreads();
  }
  const uint32_t m = min(total_s, P);
  uint32_t p2 = 2;
  while (p2 < m)
    p2 <<= 1;
  for (uint32_t i = m + threadIdx.x; i < p2; i += blockDim.x)
  {
    skey[i] = CUDART_INF;
    sval[i] = GC_NO_NODE;
  }
  __syncthreads();
  bitonic_sort(skey, sval, p2);
  const uint32_t want = (k_out < (unsigned long long)m) ? (uint32_t)k_out : m;
  for (uint32_t i = threadIdx.x; i < min(want, out_cap); i += blockDim.x)
  {
    out_idx[(size_t)q * out_cap + i] = sval[i];
    out_dist[(size_t)q * out_cap + i] = skey[i];
  }
  if (threadIdx.x == 0)
    out_count[q] = want;
}

// The same merge with the exchange fused in: shard g's lists are read straight from shard g's own result buffers
// through peer-mapped pointers (NVLink / NVSwitch loads), so there is no all-gather and no gathered copy.  The
// callers put one cross-GPU barrier between the producing kernels and this one.
struct PeerLists
{
  const uint32_t* idx[16];
  const double* dist[16];
  const uint32_t* count[16];
};
__global__ void __launch_bounds__(1024, 1)
    merge_shard_lists_p2p_kernel(PeerLists pl, uint32_t G, uint32_t nq, uint32_t cap, unsigned long long k_out,
                                 uint32_t out_cap, uint32_t* __restrict__ out_idx, double* __restrict__ out_dist,
                                 uint32_t* __restrict__ out_count, uint32_t P)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* skey = reinterpret_cast<double*>(smem_raw);
  uint32_t* sval = reinterpret_cast<uint32_t*>(skey + P);
  __shared__ uint32_t base_s[17];
  const uint32_t q = blockIdx.x;
  if (threadIdx.x == 0)
  {
    uint32_t tot = 0;
    for (uint32_t g = 0; g < G; g++)
    {
      base_s[g] = tot;
      tot += min(__ldcg(pl.count[g] + q), cap);
    }
    base_s[G] = tot;
  }
  __syncthreads();
  for (uint32_t g = 0; g < G; g++)
  {
    const uint32_t base = base_s[g], n = base_s[g + 1] - base_s[g];
    const uint32_t* li = pl.idx[g] + (size_t)q * cap;
    const double* ld = pl.dist[g] + (size_t)q * cap;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
      if (base + i < P)
      {
        const uint32_t local = __ldcg(li + i);
        skey[base + i] = __ldcg(ld + i);
        sval[base + i] = (local == GC_NO_NODE) ? GC_NO_NODE : local * G + g;
      }
  }
  __syncthreads();
  const uint32_t m = min(base_s[G], P);
  uint32_t p2 = 2;
  while (p2 < m)
    p2 <<= 1;
  for (uint32_t i = m + threadIdx.x; i < p2; i += blockDim.x)
  {
    skey[i] = CUDART_INF;
    sval[i] = GC_NO_NODE;
  }
  __syncthreads();
  bitonic_sort(skey, sval, p2);
  const uint32_t want = (k_out < (unsigned long long)m) ? (uint32_t)k_out : m;
  for (uint32_t i = threadIdx.x; i < min(want, out_cap); i += blockDim.x)
  {
    out_idx[(size_t)q * out_cap + i] = sval[i];
    out_dist[(size_t)q * out_cap + i] = skey[i];
  }
  if (threadIdx.x == 0)
    out_count[q] = want;
}

// the same gather with the shard lists read through peer-mapped pointers
__global__ void gather_shard_lists_p2p_kernel(PeerLists pl, uint32_t G, uint32_t cap, uint32_t* __restrict__ out_idx,
                                              double* __restrict__ out_dist, uint32_t* __restrict__ total)
{
  const uint32_t q = blockIdx.y;
  uint32_t base = 0;
  for (uint32_t g = 0; g < G; g++)
  {
    const uint32_t n = min(__ldcg(pl.count[g] + q), cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
      const uint32_t local = __ldcg(pl.idx[g] + (size_t)q * cap + i);
      out_idx[(size_t)q * G * cap + base + i] = (local == GC_NO_NODE) ? GC_NO_NODE : local * G + g;
      out_dist[(size_t)q * G * cap + base + i] = __ldcg(pl.dist[g] + (size_t)q * cap + i);
    }
    base += n;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    total[q] = base;
}

// ---------------------------------------------------------------------------------------------
// k-nearest of a FEW queries over a large single-segment store in two launches (the latency path of BASELINE config 4:
// one planner asking for its 16 nearest nodes among 2^20).  The sample -> collect -> re-threshold -> collect -> select
// pipeline above costs five launches, three of them single-CTA (120 us for one query); here
//   1. knn_small_scan_kernel: every CTA owns a contiguous strip of nodes, computes the fp32 distance of each node to
//      each query once, finds the exact k-th smallest fp32 value OF ITS STRIP by a radix select in shared memory and
//      keeps the nodes under the error-widened bound of it (stream_thr).  A node among the true k nearest can be pushed
//      out of a strip's list only by k strip nodes that are strictly nearer in fp64 -- impossible -- so the union of the
//      lists holds the answer;
//   2. knn_small_merge_kernel: per query, the k-th smallest fp32 value of the union (radix select), the exact fp64
//      distance only for the entries under its widened bound (k + a handful), sort by (d64, slot).
// Lists that overflow their fixed capacity (massive ties) raise a flag and the call falls back to the pipeline above.
// ---------------------------------------------------------------------------------------------
constexpr int kSmThreads = 512;
constexpr uint32_t kSmMaxQ = 8;
constexpr uint32_t kSmMaxK = 64;
constexpr uint32_t kSmRefine = 2048;

// k-th smallest (1-based) of n unsigned keys in shared or global memory, block-wide (blockDim >= 256); hist[256], sh[16]
__device__ __forceinline__ uint32_t block_kth_u32(const uint32_t* keys, uint32_t n, uint32_t k, uint32_t* hist,
                                                  uint32_t* sh)
{
  uint32_t prefix = 0, mask = 0;
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  for (int r = 3; r >= 0; r--)
  {
    if (tid < 256)
      hist[tid] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < n; i += blockDim.x)
    {
      const uint32_t key = keys[i];
      if ((key & mask) == prefix)
        atomicAdd(&hist[(key >> (8 * r)) & 255u], 1u);
    }
    __syncthreads();
    uint32_t c = 0, incl = 0;
    if (tid < 256)
    {
      c = hist[tid];
      incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
      {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o)
          incl += v;
      }
      if (lane == 31)
        sh[tid >> 5] = incl;
    }
    __syncthreads();
    if (tid < 256)
    {
      uint32_t off = 0;
      for (uint32_t w = 0; w < (tid >> 5); w++)
        off += sh[w];
      incl += off;
      const uint32_t excl = incl - c;
      if (excl < k && k <= incl)
      {
        sh[8] = tid;
        sh[9] = k - excl;
      }
    }
    __syncthreads();
    prefix |= sh[8] << (8 * r);
    mask |= 0xffu << (8 * r);
    k = sh[9];
    __syncthreads();
  }
  return prefix;
}

struct SmallKnnArgs
{
  const float* x32;
  const double* x64;
  const uint32_t* used;
  size_t row_stride;
  const double* q;  // [nq][D]
  uint32_t nq, k, G, chunk, capC;
  float xn_max, slack_abs;
  float* cand_c2;       // [nq][G][capC]
  uint32_t* cand_slot;  // [nq][G][capC]
  uint32_t* cand_cnt;   // [nq][G]
  uint32_t* err;
  uint32_t out_cap;
  uint32_t* out_idx;
  double* out_dist;
  uint32_t* out_count;
};

template <int D>
__global__ void __launch_bounds__(kSmThreads) knn_small_scan_kernel(SmallKnnArgs a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* vals = reinterpret_cast<float*>(smem_raw);  // [nq][chunk]
  __shared__ float qf[kSmMaxQ][D];
  __shared__ float e2s[kSmMaxQ];
  __shared__ uint32_t nfin[kSmMaxQ];
  __shared__ uint32_t hist[256];
  __shared__ uint32_t sh[16];
  __shared__ uint32_t scnt;
  const uint32_t tid = threadIdx.x;
  const uint32_t n = a.used[0];
  const uint32_t lo = blockIdx.x * a.chunk;
  const uint32_t len = (lo < n) ? min(a.chunk, n - lo) : 0u;
  if (tid < a.nq)
  {
    float qn = 0.0f;
    for (int d = 0; d < D; d++)
    {
      const float v = (float)a.q[(size_t)tid * D + d];
      qf[tid][d] = v;
      qn = fmaf(v, v, qn);
    }
    e2s[tid] = (float)(3 * D + 8) * 5.9604645e-8f * 1.5f * (a.xn_max + qn) * 2.0f + FLT_MIN;
    nfin[tid] = 0;
  }
  __syncthreads();
  uint32_t fin[kSmMaxQ];
#pragma unroll
  for (uint32_t qq = 0; qq < kSmMaxQ; qq++)
    fin[qq] = 0;
  for (uint32_t i = tid; i < len; i += kSmThreads)
  {
    float x[D];
#pragma unroll
    for (int d = 0; d < D; d++)
      x[d] = __ldg(a.x32 + (size_t)d * a.row_stride + lo + i);
#pragma unroll
    for (uint32_t qq = 0; qq < kSmMaxQ; qq++)
      if (qq < a.nq)
      {
        float acc = 0.0f;
#pragma unroll
        for (int d = 0; d < D; d++)
        {
          const float t = x[d] - qf[qq][d];
          acc = fmaf(t, t, acc);
        }
        vals[(size_t)qq * a.chunk + i] = acc;  // +inf for a tombstone (row 0 carries the mark)
        fin[qq] += (acc < CUDART_INF_F) ? 1u : 0u;
      }
  }
#pragma unroll
  for (uint32_t qq = 0; qq < kSmMaxQ; qq++)
    if (qq < a.nq && fin[qq])
      atomicAdd(&nfin[qq], fin[qq]);
  __syncthreads();
  for (uint32_t qq = 0; qq < a.nq; qq++)
  {
    const uint32_t kk = min(a.k, nfin[qq]);
    const float* v = vals + (size_t)qq * a.chunk;
    float thr = -1.0f;
    if (kk > 0)
    {
      const uint32_t kth = block_kth_u32(reinterpret_cast<const uint32_t*>(v), len, kk, hist, sh);
      thr = stream_thr(__uint_as_float(kth), e2s[qq], a.slack_abs);
    }
    if (tid == 0)
      scnt = 0;
    __syncthreads();
    float* oc = a.cand_c2 + ((size_t)qq * a.G + blockIdx.x) * a.capC;
    uint32_t* os = a.cand_slot + ((size_t)qq * a.G + blockIdx.x) * a.capC;
    for (uint32_t i = tid; i < len; i += kSmThreads)
    {
      const float c = v[i];
      if (c <= thr && c < CUDART_INF_F)
      {
        const uint32_t pos = atomicAdd(&scnt, 1u);
        if (pos < a.capC)
        {
          oc[pos] = c;
          os[pos] = lo + i;
        }
      }
    }
    __syncthreads();
    for (uint32_t pos = min(scnt, a.capC) + tid; pos < a.capC; pos += kSmThreads)
      oc[pos] = CUDART_INF_F;  // unused entries sort behind every candidate in the merge kernel's select
    if (tid == 0)
    {
      a.cand_cnt[(size_t)qq * a.G + blockIdx.x] = min(scnt, a.capC);
      if (scnt > a.capC)
        atomicOr(a.err, 1u);
    }
    __syncthreads();
  }
}

template <int D>
__global__ void __launch_bounds__(1024, 1) knn_small_merge_kernel(SmallKnnArgs a)
{
  __shared__ double rkey[kSmRefine];
  __shared__ uint32_t rval[kSmRefine];
  __shared__ double qs[D];
  __shared__ uint32_t hist[256];
  __shared__ uint32_t sh[16];
  __shared__ uint32_t tot, nref;
  const uint32_t qq = blockIdx.x, tid = threadIdx.x;
  if (tid < D)
    qs[tid] = a.q[(size_t)qq * D + tid];
  if (tid == 0)
  {
    tot = 0;
    nref = 0;
  }
  __syncthreads();
  // the union of the strip lists, read in place (L2): unused entries hold +inf
  const uint32_t ncand = a.G * a.capC;
  const float* cc = a.cand_c2 + (size_t)qq * ncand;
  const uint32_t* cs = a.cand_slot + (size_t)qq * ncand;
  const uint32_t* cn = a.cand_cnt + (size_t)qq * a.G;
  uint32_t mine = 0;
  for (uint32_t c = tid; c < a.G; c += blockDim.x)
    mine += cn[c];
  if (mine)
    atomicAdd(&tot, mine);
  __syncthreads();
  const uint32_t T = tot;
  const uint32_t want = min(a.k, T);  // every strip handed over min(k, its live nodes): T >= min(k, live nodes)
  uint32_t m = 0;
  if (want > 0)
  {
    float qn = 0.0f;
    for (int d = 0; d < D; d++)
    {
      const float v = (float)qs[d];
      qn = fmaf(v, v, qn);
    }
    const float e2 = (float)(3 * D + 8) * 5.9604645e-8f * 1.5f * (a.xn_max + qn) * 2.0f + FLT_MIN;
    // stage the values once: the four histogram passes of the select then run out of shared memory
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sk = reinterpret_cast<float*>(smem_raw);
    for (uint32_t i = tid; i < ncand; i += blockDim.x)
      sk[i] = __ldg(cc + i);
    __syncthreads();
    const uint32_t kth = block_kth_u32(reinterpret_cast<const uint32_t*>(sk), ncand, want, hist, sh);
    const float thr = stream_thr(__uint_as_float(kth), e2, a.slack_abs);
    for (uint32_t i = tid; i < ncand; i += blockDim.x)
    {
      const float c = sk[i];
      if (c <= thr && c < CUDART_INF_F)
      {
        const uint32_t pos = atomicAdd(&nref, 1u);
        if (pos < kSmRefine)
        {
          const uint32_t slot = cs[i];
          rkey[pos] = dist64<D>(a.x64 + (size_t)slot * D, qs);
          rval[pos] = slot;
        }
      }
    }
    __syncthreads();
    if (nref > kSmRefine)
    {
      if (tid == 0)
        atomicOr(a.err, 2u);
      m = kSmRefine;
    }
    else
      m = nref;
  }
  uint32_t p2 = 2;
  while (p2 < m)
    p2 <<= 1;
  for (uint32_t i = m + tid; i < p2; i += blockDim.x)
  {
    rkey[i] = CUDART_INF;
    rval[i] = GC_NO_NODE;
  }
  __syncthreads();
  bitonic_sort(rkey, rval, p2);
  for (uint32_t i = tid; i < min(want, a.out_cap); i += blockDim.x)
  {
    a.out_idx[(size_t)qq * a.out_cap + i] = rval[i];
    a.out_dist[(size_t)qq * a.out_cap + i] = rkey[i];
  }
  if (tid == 0)
    a.out_count[qq] = want;
}

// ---------------------------------------------------------------------------------------------
// dispatch helpers
// ---------------------------------------------------------------------------------------------
#define GC_DISPATCH_DIM(dim, ...)                                  \
  switch (dim)                                                       \
  {                                                                  \
    case 1: { constexpr int DD = 1; __VA_ARGS__; } break;                   \
    case 2: { constexpr int DD = 2; __VA_ARGS__; } break;                   \
    case 3: { constexpr int DD = 3; __VA_ARGS__; } break;                   \
    case 4: { constexpr int DD = 4; __VA_ARGS__; } break;                   \
    case 5: { constexpr int DD = 5; __VA_ARGS__; } break;                   \
    case 6: { constexpr int DD = 6; __VA_ARGS__; } break;                   \
    case 7: { constexpr int DD = 7; __VA_ARGS__; } break;                   \
    case 8: { constexpr int DD = 8; __VA_ARGS__; } break;                   \
    case 9: { constexpr int DD = 9; __VA_ARGS__; } break;                   \
    case 10: { constexpr int DD = 10; __VA_ARGS__; } break;                 \
    case 11: { constexpr int DD = 11; __VA_ARGS__; } break;                 \
    case 12: { constexpr int DD = 12; __VA_ARGS__; } break;                 \
    case 13: { constexpr int DD = 13; __VA_ARGS__; } break;                 \
    case 14: { constexpr int DD = 14; __VA_ARGS__; } break;                 \
    case 15: { constexpr int DD = 15; __VA_ARGS__; } break;                 \
    case 16: { constexpr int DD = 16; __VA_ARGS__; } break;                 \
    default: gc::set_error("dimension %d not in [1,%d]", dim, GC_MAX_DIM); return GC_E_INVALID; \
  }

double store_slack(const gc_store* s)
{
  double m = 0.0;
  for (double v : s->max_abs)
    m = fmax(m, v);
  return GC_SCAN_SLACK_REL * sqrt((double)s->dim) * m;
}

int refresh_store_mirror(gc_store* s)
{
  GC_CUDA(cudaStreamSynchronize(s->stream));
  std::vector<uint32_t> now(s->nseg);
  GC_CUDA(cudaMemcpy(now.data(), s->d_used, sizeof(uint32_t) * s->nseg, cudaMemcpyDeviceToHost));
  for (uint32_t g = 0; g < s->nseg; g++)
  {
    s->live[g] += now[g] - s->used[g];  // device-side appends never tombstone
    s->used[g] = now[g];
  }
  return GC_OK;
}

static uint32_t max_used(const gc_store* s)
{
  if (s->used_bound)
    return s->used_bound;
  uint32_t m = 0;
  for (uint32_t v : s->used)
    m = v > m ? v : m;
  return m;
}

static void scan_grid(const gc_store* s, uint32_t nq, bool per_query_seg, uint32_t* groups, uint32_t* group_size,
                      uint32_t* chunks)
{
  *group_size = per_query_seg ? 1u : (uint32_t)kQT;
  *groups = (nq + *group_size - 1) / *group_size;
  uint32_t ntiles = (max_used(s) + kTile - 1) / kTile;
  if (ntiles == 0)
    ntiles = 1;
  uint32_t target = (uint32_t)s->sms * 4u;
  uint32_t c = (target + *groups - 1) / *groups;
  if (c > ntiles)
    c = ntiles;
  if (c < 1)
    c = 1;
  *chunks = c;
}

// streaming scan (scan_stream_kernel): single-segment stores from kStreamMinNodes slots up
constexpr uint32_t kStreamMinNodes = 32768;
static bool use_stream(const gc_store* s, const uint32_t* d_seg)
{
  return d_seg == nullptr && s->used[0] >= kStreamMinNodes;
}
// control words of the streaming scan: gmin[nq] (0xFF..) then ctl[groups][2] (0).  The kernels leave them in
// their initial state, so they are only (re)initialised when the buffer has to grow.
static int ensure_stream_ctl(gc_store* s, uint32_t nq, uint32_t groups, uint32_t** gmin, uint32_t** ctl)
{
  const size_t need_q = std::max<size_t>(nq, 4096), need_g = std::max<size_t>(groups, 4096);
  if (s->stream_ctl_q < need_q || s->stream_ctl_g < need_g)
  {
    GC_CUDA(cudaStreamSynchronize(s->stream));
    GC_TRY(s->d_gmin.reserve(sizeof(uint32_t) * (need_q + 2 * need_g)));
    GC_CUDA(cudaMemsetAsync(s->d_gmin.p, 0xFF, sizeof(uint32_t) * need_q, s->stream));
    GC_CUDA(cudaMemsetAsync(s->d_gmin.as<uint32_t>() + need_q, 0, sizeof(uint32_t) * 2 * need_g, s->stream));
    s->stream_ctl_q = need_q;
    s->stream_ctl_g = need_g;
  }
  *gmin = s->d_gmin.as<uint32_t>();
  *ctl = s->d_gmin.as<uint32_t>() + s->stream_ctl_q;
  return GC_OK;
}

// MODE 0 writes the final (idx, dist) itself: a.out_idx / a.out_dist must point at the [nq] result arrays
template <int MODE>
static int launch_stream(gc_store* s, ScanArgs& a, uint32_t nq)
{
  const uint32_t groups = (nq + kStGQ - 1) / kStGQ;
  a.group_size = (nq + groups - 1) / groups;
  const uint32_t ntiles = (s->used[0] + kStTile - 1) / kStTile;
  double m = 0.0;
  for (double v : s->max_abs)
    m = fmax(m, v);
  a.xn_max = (float)fmin((double)s->dim * m * m * 1.000001, 3.0e38);
  GC_TRY(ensure_stream_ctl(s, nq, groups, &a.gmin, &a.ctl));
  int* occ = s->stream_occ;  // per store, hence per device: the opt-in below is a per-context attribute
  int blocks = 0;
  GC_DISPATCH_DIM(
      s->dim, constexpr size_t smem = stream_smem_bytes<DD>(); if (!occ[DD]) {
        GC_CUDA(cudaFuncSetAttribute(scan_stream_kernel<DD, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GC_CUDA(cudaFuncSetAttribute(scan_stream_kernel<DD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GC_CUDA(cudaFuncSetAttribute(scan_stream_kernel<DD, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int b = 0;
        GC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, scan_stream_kernel<DD, 0>, kStThreads, smem));
        occ[DD] = b > 0 ? b : 1;
      } blocks = occ[DD];
      // CTAs per group: enough to fill every SM slot; tiles are handed out dynamically, so a few CTAs more than
      // slots cost nothing (late starters find the counter exhausted)
      uint32_t chunks = ((uint32_t)(s->sms * blocks) + groups - 1) / groups; if (chunks > ntiles) chunks = ntiles;
      if (chunks < 1) chunks = 1; if (MODE == 0) {
        GC_TRY(s->d_part.reserve((size_t)chunks * nq * (sizeof(double) + sizeof(uint32_t))));
        a.part_d = s->d_part.as<double>();
        a.part_i = reinterpret_cast<uint32_t*>(a.part_d + (size_t)chunks * nq);
      } dim3 grid(chunks, groups);
      (scan_stream_kernel<DD, MODE><<<grid, kStThreads, smem, s->stream>>>(a)));
  GC_CUDA(cudaGetLastError());
  count_launch();
  return GC_OK;
}

int nn1_dev(gc_store* s, const double* d_q, const uint32_t* d_seg, uint32_t nq, uint32_t* d_idx, double* d_dist,
            bool)
{
  if (nq == 0)
    return GC_OK;
  uint32_t groups, gsz, chunks;
  scan_grid(s, nq, d_seg != nullptr, &groups, &gsz, &chunks);
  GC_TRY(s->d_part.reserve((size_t)chunks * nq * (sizeof(double) + sizeof(uint32_t))));
  ScanArgs a{};
  a.x32 = s->x32;
  a.x64 = s->x64;
  a.used = s->d_used;
  a.q = d_q;
  a.seg = d_seg;
  a.nq = nq;
  a.group_size = gsz;
  a.cap_seg = s->cap_seg;
  a.row_stride = s->cap_total + kRowPad;
  a.slack_abs = (float)(store_slack(s) * 1.0000001) + FLT_MIN;
  a.part_d = s->d_part.as<double>();
  a.part_i = reinterpret_cast<uint32_t*>(a.part_d + (size_t)chunks * nq);
  if (max_used(s) <= kSmallSegMax && !getenv("GCB200_NO_SMALL_SCAN"))
  {
    a.out_idx = d_idx;
    a.out_dist = d_dist;
    GC_DISPATCH_DIM(s->dim, (small_scan_kernel<DD, 0><<<(nq + kSmallWarps - 1) / kSmallWarps, kSmallWarps * 32, 0,
                                                        s->stream>>>(a)));
    GC_CUDA(cudaGetLastError());
    count_launch();
    return GC_OK;
  }
  if (use_stream(s, d_seg))
  {
    a.out_idx = d_idx;
    a.out_dist = d_dist;
    return launch_stream<0>(s, a, nq);
  }
  dim3 grid(chunks, groups);
  GC_DISPATCH_DIM(s->dim, (scan_kernel<DD, 0><<<grid, kScanThreads, 0, s->stream>>>(a)));
  GC_CUDA(cudaGetLastError());
  nn1_finalize_kernel<<<(nq + 3) / 4, 128, 0, s->stream>>>(a.part_d, a.part_i, chunks, nq, d_idx, d_dist);
  GC_CUDA(cudaGetLastError());
  count_launch(2);
  return GC_OK;
}

// Sorts nq lists of up to `cap` entries (count == nullptr: exactly cap), `stride` entries apart, of any length:
// sort_runs_kernel then ceil(log2(cap / kSortMax)) rank-merge passes between the lists and (tmp_idx, tmp_dist).
static int sort_large(cudaStream_t stream, uint32_t nq, uint32_t cap, size_t stride, uint32_t* d_idx, double* d_dist,
                      const uint32_t* d_count, uint32_t* tmp_idx, double* tmp_dist)
{
  if (nq == 0 || cap == 0)
    return GC_OK;
  const size_t smem = (size_t)kSortMax * (sizeof(double) + sizeof(uint32_t));
  GC_CUDA(cudaFuncSetAttribute(sort_runs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const uint32_t runs = (cap + kSortMax - 1) / kSortMax;
  for (uint32_t q0 = 0; q0 < nq; q0 += 65535u)  // grid.y limit
  {
    const uint32_t nqc = std::min(65535u, nq - q0);
    uint32_t* li = d_idx + (size_t)q0 * stride;
    double* ld = d_dist + (size_t)q0 * stride;
    uint32_t* ti = tmp_idx + (size_t)q0 * stride;
    double* td = tmp_dist + (size_t)q0 * stride;
    const uint32_t* cnt = d_count ? d_count + q0 : nullptr;
    sort_runs_kernel<<<dim3(runs, nqc), 1024, smem, stream>>>(li, ld, cnt, cap, stride);
    GC_CUDA(cudaGetLastError());
    count_launch();
    bool in_tmp = false;
    for (uint64_t L = kSortMax; L < cap; L <<= 1)
    {
      merge_runs_kernel<<<dim3((cap + 255) / 256, nqc), 256, 0, stream>>>(in_tmp ? ti : li, in_tmp ? td : ld,
                                                                          in_tmp ? li : ti, in_tmp ? ld : td, cnt, cap,
                                                                          stride, (uint32_t)L);
      GC_CUDA(cudaGetLastError());
      count_launch();
      in_tmp = !in_tmp;
    }
    if (in_tmp)
    {
      // only the first min(count, cap) entries of a list were written to the other buffer: copy whole rows back
      const size_t rows = (size_t)(nqc - 1) * stride + cap;
      GC_CUDA(cudaMemcpyAsync(li, ti, sizeof(uint32_t) * rows, cudaMemcpyDeviceToDevice, stream));
      GC_CUDA(cudaMemcpyAsync(ld, td, sizeof(double) * rows, cudaMemcpyDeviceToDevice, stream));
    }
  }
  return GC_OK;
}

static int sort_lists(gc_store* s, uint32_t nq, uint32_t cap, uint32_t* d_idx, double* d_dist, uint32_t* d_count)
{
  uint32_t P = 2;
  while (P < cap)
    P <<= 1;
  if (P > kSortMax)
  {
    // cap entries per list, runs of kSortMax merged pairwise through the candidate scratch of the store
    GC_TRY(s->d_cand_idx.reserve(sizeof(uint32_t) * (size_t)nq * cap));
    GC_TRY(s->d_cand_dist.reserve(sizeof(double) * (size_t)nq * cap));
    return sort_large(s->stream, nq, cap, cap, d_idx, d_dist, d_count, s->d_cand_idx.as<uint32_t>(),
                      s->d_cand_dist.as<double>());
  }
  size_t smem = (size_t)P * (sizeof(double) + sizeof(uint32_t));
  if (!s->sort_attr_set)
  {
    GC_CUDA(cudaFuncSetAttribute(sort_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(kSortMax * (sizeof(double) + sizeof(uint32_t)))));
    s->sort_attr_set = true;
  }
  int threads = P >= 2048 ? 1024 : (P >= 512 ? 256 : 64);
  sort_lists_kernel<<<nq, threads, smem, s->stream>>>(d_idx, d_dist, d_count, cap, P);
  GC_CUDA(cudaGetLastError());
  count_launch();
  return GC_OK;
}

int radius_dev(gc_store* s, const double* d_q, const uint32_t* d_seg, const double* d_radius, uint32_t nq,
               uint32_t cap, uint32_t* d_idx, double* d_dist, uint32_t* d_count)
{
  if (nq == 0)
    return GC_OK;
  GC_REQUIRE(cap > 0, "gc_radius: cap must be > 0");
  uint32_t groups, gsz, chunks;
  scan_grid(s, nq, d_seg != nullptr, &groups, &gsz, &chunks);
  GC_CUDA(cudaMemsetAsync(d_count, 0, sizeof(uint32_t) * nq, s->stream));
  ScanArgs a{};
  a.x32 = s->x32;
  a.x64 = s->x64;
  a.used = s->d_used;
  a.q = d_q;
  a.seg = d_seg;
  a.radius = d_radius;
  a.nq = nq;
  a.group_size = gsz;
  a.cap_seg = s->cap_seg;
  a.row_stride = s->cap_total + kRowPad;
  a.slack_abs = (float)(store_slack(s) * 1.0000001) + FLT_MIN;
  a.cap = cap;
  a.out_idx = d_idx;
  a.out_dist = d_dist;
  a.out_count = d_count;
  if (max_used(s) <= kSmallSegMax && cap <= kSmallCapMax && !getenv("GCB200_NO_SMALL_SCAN"))
  {
    const size_t smem = (size_t)kSmallWarps * cap * (sizeof(double) + sizeof(uint32_t));
    GC_DISPATCH_DIM(s->dim,
                    if (smem > 48 * 1024) GC_CUDA(cudaFuncSetAttribute(small_scan_kernel<DD, 1>,
                                                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                                       (int)smem));
                    (small_scan_kernel<DD, 1><<<(nq + kSmallWarps - 1) / kSmallWarps, kSmallWarps * 32, smem,
                                               s->stream>>>(a)));
    GC_CUDA(cudaGetLastError());
    count_launch();
    return GC_OK;
  }
  if (use_stream(s, d_seg))
  {
    GC_TRY(launch_stream<1>(s, a, nq));
    return sort_lists(s, nq, cap, d_idx, d_dist, d_count);
  }
  dim3 grid(chunks, groups);
  GC_DISPATCH_DIM(s->dim, (scan_kernel<DD, 1><<<grid, kScanThreads, 0, s->stream>>>(a)));
  GC_CUDA(cudaGetLastError());
  count_launch();
  return sort_lists(s, nq, cap, d_idx, d_dist, d_count);
}

// k above kSortMax on a segment larger than kSortMax slots -- what Tree::nearK asks for in high dimensions, where
// k_rrt * ln(N + 1) exceeds N (tree.cpp:52-53, 761-765) and the answer is "every node, sorted": fp64 distance of every
// slot, full sort (sort_large), first min(k, live) entries.  Queries go in chunks that keep the scratch under ~1 GiB.
static int knn_full_sort_dev(gc_store* s, const double* d_q, const uint32_t* d_seg, uint32_t nq, uint64_t k, uint32_t cap,
                             uint32_t* d_idx, double* d_dist, uint32_t* d_count)
{
  const uint32_t npad = max_used(s);
  const uint32_t qchunk =
      (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(std::min<uint32_t>(nq, 65535u), (1ull << 30) / ((uint64_t)npad * 24ull)));
  const size_t n = (size_t)qchunk * npad;
  GC_TRY(s->d_cand_idx.reserve(sizeof(uint32_t) * n * 2));
  GC_TRY(s->d_cand_dist.reserve(sizeof(double) * n * 2));
  GC_TRY(s->d_cand_count.reserve(sizeof(uint32_t) * ((size_t)qchunk + 4)));
  uint32_t* li = s->d_cand_idx.as<uint32_t>();
  double* ld = s->d_cand_dist.as<double>();
  uint32_t* live = s->d_cand_count.as<uint32_t>();
  for (uint32_t q0 = 0; q0 < nq; q0 += qchunk)
  {
    const uint32_t nqc = std::min(qchunk, nq - q0);
    GC_CUDA(cudaMemsetAsync(live, 0, sizeof(uint32_t) * nqc, s->stream));
    GC_DISPATCH_DIM(s->dim, (dist_all_kernel<DD><<<dim3((npad + 255) / 256, nqc), 256, 0, s->stream>>>(
                                s->x32, s->x64, s->d_used, d_q + (size_t)q0 * s->dim, d_seg ? d_seg + q0 : nullptr,
                                s->cap_seg, npad, li, ld, live)));
    GC_CUDA(cudaGetLastError());
    count_launch();
    GC_TRY(sort_large(s->stream, nqc, npad, npad, li, ld, nullptr, li + n, ld + n));
    const uint32_t width = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(k, cap), npad);
    take_first_kernel<<<dim3((std::max(width, 1u) + 255) / 256, nqc), 256, 0, s->stream>>>(
        li, ld, live, npad, (unsigned long long)k, cap, d_idx + (size_t)q0 * cap, d_dist + (size_t)q0 * cap,
        d_count + q0);
    GC_CUDA(cudaGetLastError());
    count_launch();
  }
  return GC_OK;
}

// sample threshold -> one collecting scan -> exact select (see the kernels above).  Queries are processed in
// chunks so that the candidate lists stay under ~256 MiB.  Returns GC_E_CAPACITY after a synchronisation
// when a candidate list overflowed its estimate (the caller retries with s->knn_cand_scale doubled).
static int knn_large_dev(gc_store* s, const double* d_q, const uint32_t* d_seg, uint32_t nq, uint64_t k, uint32_t cap,
                         uint32_t* d_idx, double* d_dist, uint32_t* d_count)
{
  if (k > kSortMax)
    return knn_full_sort_dev(s, d_q, d_seg, nq, k, cap, d_idx, d_dist, d_count);
  if (k == 0)
  {
    GC_CUDA(cudaMemsetAsync(d_count, 0, sizeof(uint32_t) * nq, s->stream));
    return GC_OK;
  }
  const uint32_t mu = max_used(s);
  // expected candidates per query: k * N / sample size; scale leaves room for sampling noise
  uint64_t est = (uint64_t)k * ((mu + kSortMax - 1) / kSortMax) * s->knn_cand_scale + 1024u;
  if (est > mu)
    est = mu;
  const uint32_t capB = (uint32_t)est;
  uint32_t qchunk = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(nq, (256ull << 20) / ((uint64_t)capB * 12ull)));
  GC_TRY(s->d_cand_idx.reserve(sizeof(uint32_t) * (size_t)qchunk * capB));
  GC_TRY(s->d_cand_dist.reserve(sizeof(double) * (size_t)qchunk * capB));
  GC_TRY(s->d_cand_count.reserve(sizeof(uint32_t) * ((size_t)qchunk + 4)));
  GC_TRY(s->d_thr.reserve(sizeof(float) * nq));
  uint32_t* d_err = s->d_cand_count.as<uint32_t>() + qchunk;
  GC_CUDA(cudaMemsetAsync(d_err, 0, sizeof(uint32_t), s->stream));
  const size_t row_stride = s->cap_total + kRowPad;
  const float slack = (float)(store_slack(s) * 1.0000001) + FLT_MIN;
  const size_t smem_sample = sizeof(float) * kSortMax;
  // streaming stores keep fp32 candidates and refine only around the k-th (select_topk_f32_kernel); if that ever
  // meets more near-ties than its sort buffer holds the call is redone on the all-fp64 path
  const bool f32_path = use_stream(s, d_seg) && !s->knn_force_fp64;
  uint32_t P32 = 2048;
  while (P32 < 4 * k && P32 < kSortMax)
    P32 <<= 1;
  const size_t smem_sel32 = (size_t)P32 * (sizeof(double) + sizeof(uint32_t));
  double xm = 0.0;
  for (double v : s->max_abs)
    xm = fmax(xm, v);
  const float xn_max = (float)fmin((double)s->dim * xm * xm * 1.000001, 3.0e38);
  GC_DISPATCH_DIM(s->dim,
                  GC_CUDA(cudaFuncSetAttribute(knn_sample_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)smem_sample));
                  GC_CUDA(cudaFuncSetAttribute(knn_sample_select_kernel<DD>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sample));
                  GC_CUDA(cudaFuncSetAttribute(select_topk_f32_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)(kSortMax * (sizeof(double) + sizeof(uint32_t)))));
                  if (f32_path) knn_sample_select_kernel<DD><<<nq, 1024, smem_sample, s->stream>>>(
                      s->x32, s->d_used, d_q, d_seg, s->cap_seg, row_stride, (unsigned long long)k, slack,
                      s->d_thr.as<float>());
                  else knn_sample_kernel<DD><<<nq, 1024, smem_sample, s->stream>>>(
                      s->x32, s->d_used, d_q, d_seg, s->cap_seg, row_stride, (unsigned long long)k, slack,
                      s->d_thr.as<float>()));
  GC_CUDA(cudaGetLastError());
  count_launch();
  const size_t smem_sel = (size_t)kSortMax * (sizeof(double) + sizeof(uint32_t));
  GC_CUDA(cudaFuncSetAttribute(select_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sel));
  for (uint32_t q0 = 0; q0 < nq; q0 += qchunk)
  {
    const uint32_t nqc = std::min(qchunk, nq - q0);
    GC_CUDA(cudaMemsetAsync(s->d_cand_count.p, 0, sizeof(uint32_t) * nqc, s->stream));
    uint32_t groups, gsz, chunks;
    scan_grid(s, nqc, d_seg != nullptr, &groups, &gsz, &chunks);
    ScanArgs a{};
    a.x32 = s->x32;
    a.x64 = s->x64;
    a.used = s->d_used;
    a.q = d_q + (size_t)q0 * s->dim;
    a.seg = d_seg ? d_seg + q0 : nullptr;
    a.thr2 = s->d_thr.as<float>() + q0;
    a.nq = nqc;
    a.group_size = gsz;
    a.cap_seg = s->cap_seg;
    a.row_stride = row_stride;
    a.slack_abs = slack;
    a.cap = capB;
    a.out_idx = s->d_cand_idx.as<uint32_t>();
    a.out_dist = s->d_cand_dist.as<double>();
    a.out_count = s->d_cand_count.as<uint32_t>();
    if (f32_path)
    {
      a.out_c2 = reinterpret_cast<float*>(s->d_cand_dist.p);
      // phase 1: the first 1/16 of the tiles under the sample threshold; re-threshold; phase 2: the rest
      const uint32_t all_tiles = (s->used[0] + kStTile - 1) / kStTile;
      const uint32_t first = std::max<uint32_t>(1u, all_tiles / 16u);
      if (first < all_tiles)
      {
        a.tile_begin = 0;
        a.tile_count = first;
        GC_TRY(launch_stream<2>(s, a, nqc));
        GC_DISPATCH_DIM(s->dim, (knn_rethreshold_kernel<DD><<<nqc, 1024, 0, s->stream>>>(
                                    a.out_c2, a.out_count, capB, a.q, (unsigned long long)k, xn_max, slack,
                                    s->d_thr.as<float>() + q0)));
        GC_CUDA(cudaGetLastError());
        count_launch();
        a.tile_begin = first;
        a.tile_count = all_tiles - first;
        a.thr_raw = 1;
      }
      GC_TRY(launch_stream<2>(s, a, nqc));
      GC_DISPATCH_DIM(s->dim, (select_topk_f32_kernel<DD><<<nqc, 1024, smem_sel32, s->stream>>>(
                                  a.out_idx, a.out_c2, a.out_count, capB, s->x64, a.q, (unsigned long long)k, xn_max,
                                  slack, cap, d_idx + (size_t)q0 * cap, d_dist + (size_t)q0 * cap, d_count + q0, d_err,
                                  P32)));
      GC_CUDA(cudaGetLastError());
      count_launch();
      continue;
    }
    {
      dim3 grid(chunks, groups);
      GC_DISPATCH_DIM(s->dim, (scan_kernel<DD, 2><<<grid, kScanThreads, 0, s->stream>>>(a)));
      GC_CUDA(cudaGetLastError());
    }
    select_topk_kernel<<<nqc, 1024, smem_sel, s->stream>>>(a.out_idx, a.out_dist, a.out_count, capB,
                                                           (unsigned long long)k, cap, d_idx + (size_t)q0 * cap,
                                                           d_dist + (size_t)q0 * cap, d_count + q0, d_err, kSortMax);
    GC_CUDA(cudaGetLastError());
    count_launch(2);
  }
  // overflow check (one word): the estimate is generous, so this almost never fires
  uint32_t err = 0;
  GC_CUDA(cudaMemcpyAsync(&err, d_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
  GC_CUDA(cudaStreamSynchronize(s->stream));
  if (err & 2u)
  {
    s->knn_force_fp64 = true;  // this store holds massive near-ties: stay on the all-fp64 path from now on
    return GC_E_CAPACITY;
  }
  if (err)
  {
    if (capB >= mu)
    {
      gc::set_error("gc_knn: candidate list overflow with a full-size list (internal error)");
      return GC_E_INVALID;
    }
    s->knn_cand_scale *= 2;
    return GC_E_CAPACITY;
  }
  return GC_OK;
}

// the two-launch latency path (see knn_small_scan_kernel); GC_E_CAPACITY = a list overflowed, use the general pipeline
static int knn_small_dev(gc_store* s, const double* d_q, uint32_t nq, uint32_t k, uint32_t cap, uint32_t* d_idx,
                         double* d_dist, uint32_t* d_count)
{
  SmallKnnArgs a{};
  const uint32_t n = s->used[0];
  // strips: a multiple of the SM count, more of them when several queries share a strip's shared memory
  uint32_t mult = 2u;
  if (const char* e = getenv("GCB200_SMALL_KNN_STRIPS"))
    mult = (uint32_t)std::max(1, atoi(e));
  for (;;)
  {
    a.G = (uint32_t)s->sms * mult;
    a.chunk = (n + a.G - 1) / a.G;
    a.chunk = (a.chunk + 3u) & ~3u;
    if (sizeof(float) * (size_t)nq * a.chunk <= 64 * 1024 || mult >= 16u)
      break;
    mult *= 2u;
  }
  a.G = (n + a.chunk - 1) / a.chunk;
  a.capC = k + 16u;
  const size_t smem1 = sizeof(float) * (size_t)nq * a.chunk;
  const size_t ncand = (size_t)a.G * a.capC;
  const size_t smem2 = sizeof(float) * ncand;
  if (smem1 > 100 * 1024 || smem2 > 160 * 1024)
    return GC_E_CAPACITY;
  GC_TRY(s->d_cand_dist.reserve(sizeof(float) * nq * ncand));
  GC_TRY(s->d_cand_idx.reserve(sizeof(uint32_t) * nq * ncand));
  GC_TRY(s->d_cand_count.reserve(sizeof(uint32_t) * ((size_t)nq * a.G + 4)));
  a.x32 = s->x32;
  a.x64 = s->x64;
  a.used = s->d_used;
  a.row_stride = s->cap_total + kRowPad;
  a.q = d_q;
  a.nq = nq;
  a.k = k;
  double xm = 0.0;
  for (double v : s->max_abs)
    xm = fmax(xm, v);
  a.xn_max = (float)fmin((double)s->dim * xm * xm * 1.000001, 3.0e38);
  a.slack_abs = (float)(store_slack(s) * 1.0000001) + FLT_MIN;
  a.cand_c2 = reinterpret_cast<float*>(s->d_cand_dist.p);
  a.cand_slot = s->d_cand_idx.as<uint32_t>();
  a.cand_cnt = s->d_cand_count.as<uint32_t>();
  a.err = a.cand_cnt + (size_t)nq * a.G;
  a.out_cap = cap;
  a.out_idx = d_idx;
  a.out_dist = d_dist;
  a.out_count = d_count;
  if (s->small_knn_err_words != (size_t)nq * a.G + 4)
  {
    // the error word sits behind the counts: cleared when the layout changes and after it was raised, not per call
    GC_CUDA(cudaMemsetAsync(a.err, 0, sizeof(uint32_t), s->stream));
    s->small_knn_err_words = (size_t)nq * a.G + 4;
  }
  GC_DISPATCH_DIM(s->dim,
                  if (smem1 > 48 * 1024) GC_CUDA(cudaFuncSetAttribute(
                      knn_small_scan_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
                  // (static + dynamic shared memory of the merge kernel exceeds the 48 KB default long before smem2 does)
                  GC_CUDA(cudaFuncSetAttribute(knn_small_merge_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)(160 * 1024)));
                  (knn_small_scan_kernel<DD><<<a.G, kSmThreads, smem1, s->stream>>>(a));
                  (knn_small_merge_kernel<DD><<<nq, 1024, smem2, s->stream>>>(a)));
  GC_CUDA(cudaGetLastError());
  count_launch(2);
  uint32_t err = 0;
  GC_CUDA(cudaMemcpyAsync(&err, a.err, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
  GC_CUDA(cudaStreamSynchronize(s->stream));
  if (getenv("GCB200_DEBUG"))
    fprintf(stderr, "knn_small_dev: n=%u nq=%u k=%u G=%u chunk=%u capC=%u smem1=%zu smem2=%zu err=%u\n", n, nq, k, a.G,
            a.chunk, a.capC, smem1, smem2, err);
  if (err)
    s->small_knn_err_words = 0;  // cleared by the next call
  return err ? GC_E_CAPACITY : GC_OK;
}

int knn_dev(gc_store* s, const double* d_q, const uint32_t* d_seg, uint32_t nq, uint64_t k, uint32_t cap,
            uint32_t* d_idx, double* d_dist, uint32_t* d_count)
{
  if (nq == 0)
    return GC_OK;
  GC_REQUIRE(cap > 0, "gc_knn: cap must be > 0");
  const uint32_t mu = max_used(s);
  if (mu > kSortMax)
  {
    int rc;
    if (nq <= kSmMaxQ && k >= 1 && k <= kSmMaxK && use_stream(s, d_seg) && !s->knn_force_fp64 &&
        !getenv("GCB200_NO_SMALL_KNN"))
    {
      rc = knn_small_dev(s, d_q, nq, (uint32_t)k, cap, d_idx, d_dist, d_count);
      if (rc != GC_E_CAPACITY)
        return rc;
    }
    do
      rc = knn_large_dev(s, d_q, d_seg, nq, k, cap, d_idx, d_dist, d_count);
    while (rc == GC_E_CAPACITY);
    return rc;
  }
  uint32_t P = 2;
  while (P < mu)
    P <<= 1;
  size_t smem = (size_t)P * (sizeof(double) + sizeof(uint32_t));
  bool* attr_set = s->knn_attr_set;
  const size_t row_stride = s->cap_total + kRowPad;
  GC_DISPATCH_DIM(
      s->dim,
      if (!attr_set[DD]) {
        GC_CUDA(cudaFuncSetAttribute(knn_all_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(kSortMax * (sizeof(double) + sizeof(uint32_t)))));
        attr_set[DD] = true;
      } knn_all_kernel<DD><<<nq, 1024, smem, s->stream>>>(s->x32, s->x64, s->d_used, d_q, d_seg, s->cap_seg,
                                                           row_stride, (unsigned long long)k, cap, d_idx, d_dist,
                                                           d_count, P));
  GC_CUDA(cudaGetLastError());
  count_launch();
  return GC_OK;
}

}  // namespace gc

// =================================================================================================
//  C ABI: store
// =================================================================================================
using namespace gc;

extern "C" {

int gc_store_create(int dim, uint32_t capacity_per_segment, uint32_t n_segments, int device, gc_store_t** out)
{
  GC_REQUIRE(out != nullptr, "gc_store_create: out is NULL");
  *out = nullptr;
  GC_REQUIRE(dim >= 1 && dim <= GC_MAX_DIM, "gc_store_create: dim %d not in [1,%d]", dim, GC_MAX_DIM);
  GC_REQUIRE(capacity_per_segment > 0 && n_segments > 0, "gc_store_create: empty capacity");
  int ndev = 0;
  GC_TRY(gc_device_count(&ndev));
  GC_REQUIRE(device >= 0 && device < ndev, "gc_store_create: device %d of %d (no CPU fallback exists)", device, ndev);
  GC_SET_DEVICE(device);
  gc_store* s = new gc_store();
  s->device = device;
  s->dim = dim;
  s->cap_seg = (uint32_t)align_up(capacity_per_segment, 32);
  s->nseg = n_segments;
  s->cap_total = (size_t)s->cap_seg * n_segments;
  if (s->cap_total > 0xFFFFFFF0ull)
  {
    gc::set_error("gc_store_create: %zu slots exceed the 32-bit slot space", s->cap_total);
    delete s;
    return GC_E_INVALID;
  }
  s->sms = num_sms(device);
  const size_t row_stride = s->cap_total + kRowPad;
  cudaError_t e = cudaMalloc(&s->x32, sizeof(float) * row_stride * dim);
  if (e == cudaSuccess)
    e = cudaMalloc(&s->x64, sizeof(double) * (s->cap_total + 8) * dim);
  if (e == cudaSuccess)
    e = cudaMalloc(&s->d_used, sizeof(uint32_t) * n_segments);
  if (e == cudaSuccess)
    e = cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess)
  {
    gc::set_error("gc_store_create: %s", cudaGetErrorString(e));
    gc_store_destroy(s);
    return GC_E_CUDA;
  }
  s->stream = s->own_stream;
  s->used.assign(n_segments, 0);
  s->live.assign(n_segments, 0);
  s->max_abs.assign(n_segments, 0.0);
  GC_CUDA(cudaMemsetAsync(s->x32, 0, sizeof(float) * row_stride * dim, s->stream));
  GC_CUDA(cudaMemsetAsync(s->d_used, 0, sizeof(uint32_t) * n_segments, s->stream));
  GC_CUDA(cudaStreamSynchronize(s->stream));
  *out = s;
  return GC_OK;
}

int gc_store_destroy(gc_store_t* s)
{
  if (!s)
    return GC_OK;
  gc::CtxGuard gc_ctx_guard_;
  cudaSetDevice(s->device);
  if (s->own_stream)
    cudaStreamSynchronize(s->own_stream);
  cudaFree(s->x32);
  cudaFree(s->x64);
  cudaFree(s->d_used);
  s->d_q.release(); s->d_seg.release(); s->d_radius.release(); s->d_idx.release(); s->d_dist.release();
  s->d_count.release(); s->d_part.release(); s->d_misc.release(); s->d_rows.release(); s->d_slots.release();
  s->h_in.release(); s->h_out.release();
  s->d_cand_idx.release(); s->d_cand_dist.release(); s->d_cand_count.release(); s->d_thr.release();
  s->d_gmin.release();
  if (s->own_stream)
    cudaStreamDestroy(s->own_stream);
  delete s;
  return GC_OK;
}

int gc_store_set_stream(gc_store_t* s, void* cuda_stream)
{
  GC_REQUIRE(s != nullptr, "gc_store_set_stream: NULL store");
  GC_CUDA(cudaStreamSynchronize(s->stream));
  s->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : s->own_stream;
  return GC_OK;
}

int gc_store_dim(gc_store_t* s) { return s ? s->dim : GC_E_INVALID; }

static void track_max_abs(gc_store* s, uint32_t seg, const double* q, size_t count)
{
  double m = s->max_abs[seg];
  for (size_t i = 0; i < count; i++)
  {
    double v = fabs(q[i]);
    if (v > m)
      m = v;
  }
  s->max_abs[seg] = m;
}

int gc_store_append(gc_store_t* s, uint32_t segment, const double* q, uint32_t n, uint32_t* first_slot)
{
  GC_REQUIRE(s && q, "gc_store_append: NULL argument");
  GC_REQUIRE(segment < s->nseg, "gc_store_append: segment %u of %u", segment, s->nseg);
  if (n == 0)
    return GC_OK;
  if ((size_t)s->used[segment] + n > s->cap_seg)
  {
    gc::set_error("gc_store_append: segment %u full (%u + %u > %u)", segment, s->used[segment], n, s->cap_seg);
    return GC_E_CAPACITY;
  }
  GC_SET_DEVICE(s->device);
  const size_t bytes = sizeof(double) * (size_t)n * s->dim;
  GC_TRY(s->h_in.reserve(bytes));
  GC_TRY(s->d_rows.reserve(bytes));
  GC_TRY(s->d_slots.reserve(sizeof(uint32_t) * n));
  // the pinned staging buffer may still feed an earlier asynchronous copy
  GC_CUDA(cudaStreamSynchronize(s->stream));
  memcpy(s->h_in.p, q, bytes);
  track_max_abs(s, segment, q, (size_t)n * s->dim);
  const uint32_t first = s->used[segment];
  const uint32_t gfirst = segment * s->cap_seg + first;
  GC_CUDA(cudaMemcpyAsync(s->d_rows.p, s->h_in.p, bytes, cudaMemcpyHostToDevice, s->stream));
  store_fill_slots_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(s->d_slots.as<uint32_t>(), gfirst, n);
  const uint32_t total = n * (uint32_t)s->dim;
  store_scatter_kernel<<<(total + 255) / 256, 256, 0, s->stream>>>(s->d_rows.as<double>(), s->d_slots.as<uint32_t>(), n,
                                                                    s->dim, s->cap_total + kRowPad, s->x32, s->x64);
  GC_CUDA(cudaGetLastError());
  count_launch(2);
  s->used[segment] = first + n;
  s->live[segment] += n;
  GC_CUDA(cudaMemcpyAsync(s->d_used + segment, &s->used[segment], sizeof(uint32_t), cudaMemcpyHostToDevice, s->stream));
  GC_CUDA(cudaStreamSynchronize(s->stream));
  if (first_slot)
    *first_slot = first;
  return GC_OK;
}

int gc_store_append_multi(gc_store_t* s, const uint32_t* segments, const double* q, uint32_t n, uint32_t* slots)
{
  GC_REQUIRE(s && q && segments, "gc_store_append_multi: NULL argument");
  if (n == 0)
    return GC_OK;
  GC_SET_DEVICE(s->device);
  const size_t bytes = sizeof(double) * (size_t)n * s->dim;
  const size_t sbytes = align_up(sizeof(uint32_t) * n, 16);
  GC_TRY(s->h_in.reserve(bytes + sbytes));
  GC_TRY(s->d_rows.reserve(bytes));
  GC_TRY(s->d_slots.reserve(sbytes));
  GC_CUDA(cudaStreamSynchronize(s->stream));
  uint32_t* hs = reinterpret_cast<uint32_t*>(static_cast<char*>(s->h_in.p) + bytes);
  for (uint32_t i = 0; i < n; i++)  // validate every segment id before the host mirror is touched
    GC_REQUIRE(segments[i] < s->nseg, "gc_store_append_multi: segment %u of %u", segments[i], s->nseg);
  for (uint32_t i = 0; i < n; i++)
  {
    const uint32_t seg = segments[i];
    if (s->used[seg] >= s->cap_seg)
    {
      // roll back the rows already accounted for
      for (uint32_t j = 0; j < i; j++)
      {
        s->used[segments[j]]--;
        s->live[segments[j]]--;
      }
      gc::set_error("gc_store_append_multi: segment %u full (%u slots)", seg, s->cap_seg);
      return GC_E_CAPACITY;
    }
    const uint32_t slot = s->used[seg]++;
    s->live[seg]++;
    hs[i] = seg * s->cap_seg + slot;
    if (slots)
      slots[i] = slot;
    track_max_abs(s, seg, q + (size_t)i * s->dim, s->dim);
  }
  memcpy(s->h_in.p, q, bytes);
  GC_CUDA(cudaMemcpyAsync(s->d_rows.p, s->h_in.p, bytes, cudaMemcpyHostToDevice, s->stream));
  GC_CUDA(cudaMemcpyAsync(s->d_slots.p, hs, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, s->stream));
  const uint32_t total = n * (uint32_t)s->dim;
  store_scatter_kernel<<<(total + 255) / 256, 256, 0, s->stream>>>(s->d_rows.as<double>(), s->d_slots.as<uint32_t>(), n,
                                                                    s->dim, s->cap_total + kRowPad, s->x32, s->x64);
  GC_CUDA(cudaGetLastError());
  count_launch();
  // the host mirror `used` outlives the copy; the pinned staging buffer is protected by the
  // synchronisation every entry point performs before it refills it, so no trailing sync is needed
  GC_CUDA(cudaMemcpyAsync(s->d_used, s->used.data(), sizeof(uint32_t) * s->nseg, cudaMemcpyHostToDevice, s->stream));
  return GC_OK;
}

static int poke_row0(gc_store* s, uint32_t segment, uint32_t slot, bool tombstone)
{
  GC_REQUIRE(s != nullptr, "NULL store");
  GC_REQUIRE(segment < s->nseg && slot < s->used[segment], "slot %u of segment %u is not in use", slot, segment);
  GC_SET_DEVICE(s->device);
  const size_t g = (size_t)segment * s->cap_seg + slot;
  float v;
  if (tombstone)
    v = INFINITY;
  else
  {
    double d0;
    GC_CUDA(cudaMemcpyAsync(&d0, s->x64 + g * s->dim, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    GC_CUDA(cudaStreamSynchronize(s->stream));
    v = (float)d0;
  }
  float cur;
  GC_CUDA(cudaMemcpyAsync(&cur, s->x32 + g, sizeof(float), cudaMemcpyDeviceToHost, s->stream));
  GC_CUDA(cudaStreamSynchronize(s->stream));
  const bool was_dead = std::isinf(cur) && cur > 0;
  if (tombstone == was_dead)
    return GC_W_NOOP;  // nothing to change
  GC_CUDA(cudaMemcpyAsync(s->x32 + g, &v, sizeof(float), cudaMemcpyHostToDevice, s->stream));
  GC_CUDA(cudaStreamSynchronize(s->stream));
  if (tombstone)
    s->live[segment]--;
  else
    s->live[segment]++;
  return GC_OK;
}

int gc_store_tombstone(gc_store_t* s, uint32_t segment, uint32_t slot) { return poke_row0(s, segment, slot, true); }
int gc_store_restore(gc_store_t* s, uint32_t segment, uint32_t slot) { return poke_row0(s, segment, slot, false); }

int gc_store_clear(gc_store_t* s, uint32_t segment)
{
  GC_REQUIRE(s != nullptr, "gc_store_clear: NULL store");
  GC_REQUIRE(segment == GC_NO_NODE || segment < s->nseg, "gc_store_clear: segment %u of %u", segment, s->nseg);
  GC_SET_DEVICE(s->device);
  for (uint32_t g = 0; g < s->nseg; g++)
    if (segment == GC_NO_NODE || segment == g)
    {
      s->used[g] = 0;
      s->live[g] = 0;
      s->max_abs[g] = 0.0;
    }
  GC_CUDA(cudaMemcpyAsync(s->d_used, s->used.data(), sizeof(uint32_t) * s->nseg, cudaMemcpyHostToDevice, s->stream));
  GC_CUDA(cudaStreamSynchronize(s->stream));
  return GC_OK;
}

int gc_store_size(gc_store_t* s, uint32_t segment, uint32_t* slots_used, uint32_t* live)
{
  GC_REQUIRE(s != nullptr && segment < s->nseg, "gc_store_size: bad store/segment");
  if (slots_used)
    *slots_used = s->used[segment];
  if (live)
    *live = s->live[segment];
  return GC_OK;
}

int gc_store_read(gc_store_t* s, uint32_t segment, uint32_t first_slot, uint32_t n, double* q_out)
{
  GC_REQUIRE(s && q_out && segment < s->nseg, "gc_store_read: bad argument");
  GC_REQUIRE((size_t)first_slot + n <= s->used[segment], "gc_store_read: slots [%u,%u) beyond %u", first_slot,
             first_slot + n, s->used[segment]);
  if (n == 0)
    return GC_OK;
  GC_SET_DEVICE(s->device);
  const size_t g = (size_t)segment * s->cap_seg + first_slot;
  GC_CUDA(cudaMemcpyAsync(q_out, s->x64 + g * s->dim, sizeof(double) * (size_t)n * s->dim, cudaMemcpyDeviceToHost,
                          s->stream));
  GC_CUDA(cudaStreamSynchronize(s->stream));
  return GC_OK;
}

// =================================================================================================
//  C ABI: queries
// =================================================================================================
int gc_nn1_dev(gc_store_t* s, const double* d_q, const uint32_t* d_seg, uint32_t nq, uint32_t* d_idx, double* d_dist)
{
  GC_REQUIRE(s && d_q && d_idx && d_dist, "gc_nn1_dev: NULL argument");
  GC_SET_DEVICE(s->device);
  return nn1_dev(s, d_q, d_seg, nq, d_idx, d_dist, false);
}
int gc_radius_dev(gc_store_t* s, const double* d_q, const uint32_t* d_seg, const double* d_radius, uint32_t nq,
                  uint32_t cap, uint32_t* d_idx, double* d_dist, uint32_t* d_count)
{
  GC_REQUIRE(s && d_q && d_radius && d_idx && d_dist && d_count, "gc_radius_dev: NULL argument");
  GC_SET_DEVICE(s->device);
  return radius_dev(s, d_q, d_seg, d_radius, nq, cap, d_idx, d_dist, d_count);
}
int gc_knn_dev(gc_store_t* s, const double* d_q, const uint32_t* d_seg, uint32_t nq, uint64_t k, uint32_t cap,
               uint32_t* d_idx, double* d_dist, uint32_t* d_count)
{
  GC_REQUIRE(s && d_q && d_idx && d_dist && d_count, "gc_knn_dev: NULL argument");
  GC_SET_DEVICE(s->device);
  return knn_dev(s, d_q, d_seg, nq, k, cap, d_idx, d_dist, d_count);
}

int gc_merge_shard_lists_dev(int device, void* cuda_stream, uint32_t n_shards, uint32_t nq, uint32_t cap,
                             const uint32_t* d_idx, const double* d_dist, const uint32_t* d_count, uint64_t k_out,
                             uint32_t out_cap, uint32_t* d_out_idx, double* d_out_dist, uint32_t* d_out_count)
{
  GC_REQUIRE(d_idx && d_dist && d_count && d_out_idx && d_out_dist && d_out_count, "gc_merge_shard_lists_dev: NULL");
  GC_REQUIRE(n_shards >= 1 && cap >= 1 && out_cap >= 1, "gc_merge_shard_lists_dev: empty shape");
  if (nq == 0)
    return GC_OK;
  uint32_t P = 2;
  while (P < n_shards * cap)
    P <<= 1;
  GC_SET_DEVICE(device);
  if (P > kSortMax)
  {
    // more entries than one CTA sorts: gather with global slots, full sort, first k_out (stream-ordered scratch)
    cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
    const size_t W = (size_t)n_shards * cap, n = (size_t)nq * W;
    void* scratch = nullptr;
    GC_CUDA(cudaMallocAsync(&scratch, n * 24 + sizeof(uint32_t) * ((size_t)nq + 4), st));
    double* ld = static_cast<double*>(scratch);
    double* td = ld + n;
    uint32_t* li = reinterpret_cast<uint32_t*>(td + n);
    uint32_t* ti = li + n;
    uint32_t* total = ti + n;
    int rc = GC_OK;
    for (uint32_t q0 = 0; q0 < nq && rc == GC_OK; q0 += 65535u)
    {
      const uint32_t nqc = std::min(65535u, nq - q0);
      // the shard-major layout is addressed with the full nq, so the kernel is given the chunk through offsets
      gather_shard_lists_kernel<<<dim3(32, nqc), 256, 0, st>>>(d_idx + (size_t)q0 * cap, d_dist + (size_t)q0 * cap,
                                                              d_count + q0, n_shards, nq, cap, li + (size_t)q0 * W,
                                                              ld + (size_t)q0 * W, total + q0);
      if (cudaGetLastError() != cudaSuccess)
        rc = GC_E_CUDA;
      count_launch();
    }
    if (rc == GC_OK)
      rc = sort_large(st, nq, (uint32_t)W, W, li, ld, total, ti, td);
    for (uint32_t q0 = 0; q0 < nq && rc == GC_OK; q0 += 65535u)
    {
      const uint32_t nqc = std::min(65535u, nq - q0);
      const uint32_t width = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(k_out, out_cap), W);
      take_first_kernel<<<dim3((std::max(width, 1u) + 255) / 256, nqc), 256, 0, st>>>(
          li + (size_t)q0 * W, ld + (size_t)q0 * W, total + q0, W, (unsigned long long)k_out, out_cap,
          d_out_idx + (size_t)q0 * out_cap, d_out_dist + (size_t)q0 * out_cap, d_out_count + q0);
      if (cudaGetLastError() != cudaSuccess)
        rc = GC_E_CUDA;
      count_launch();
    }
    cudaFreeAsync(scratch, st);
    if (rc != GC_OK)
      gc::set_error("gc_merge_shard_lists_dev: large merge failed");
    return rc;
  }
  const size_t smem = (size_t)P * (sizeof(double) + sizeof(uint32_t));
  GC_CUDA(cudaFuncSetAttribute(merge_shard_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)(kSortMax * (sizeof(double) + sizeof(uint32_t)))));
  const int threads = P >= 2048 ? 1024 : (P >= 512 ? 256 : 64);
  merge_shard_lists_kernel<<<nq, threads, smem, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(
      d_idx, d_dist, d_count, n_shards, nq, cap, (unsigned long long)k_out, out_cap, d_out_idx, d_out_dist,
      d_out_count, P);
  GC_CUDA(cudaGetLastError());
  count_launch();
  return GC_OK;
}

int gc_merge_shard_lists_p2p_dev(int device, void* cuda_stream, uint32_t n_shards, uint32_t nq, uint32_t cap,
                                 const void* const* d_idx_of_shard, const void* const* d_dist_of_shard,
                                 const void* const* d_count_of_shard, uint64_t k_out, uint32_t out_cap,
                                 uint32_t* d_out_idx, double* d_out_dist, uint32_t* d_out_count)
{
  GC_REQUIRE(d_idx_of_shard && d_dist_of_shard && d_count_of_shard && d_out_idx && d_out_dist && d_out_count,
             "gc_merge_shard_lists_p2p_dev: NULL argument");
  GC_REQUIRE(n_shards >= 1 && n_shards <= 16, "gc_merge_shard_lists_p2p_dev: %u shards (1..16)", n_shards);
  if (nq == 0)
    return GC_OK;
  GC_SET_DEVICE(device);
  PeerLists pl{};
  for (uint32_t g = 0; g < n_shards; g++)
  {
    GC_REQUIRE(d_idx_of_shard[g] && d_dist_of_shard[g] && d_count_of_shard[g], "gc_merge_shard_lists_p2p_dev: shard %u", g);
    pl.idx[g] = static_cast<const uint32_t*>(d_idx_of_shard[g]);
    pl.dist[g] = static_cast<const double*>(d_dist_of_shard[g]);
    pl.count[g] = static_cast<const uint32_t*>(d_count_of_shard[g]);
  }
  if ((uint64_t)n_shards * cap > kSortMax)
  {
    GC_REQUIRE(nq <= 65535u, "gc_merge_shard_lists_p2p_dev: %u queries with lists above %u entries (max 65535)", nq,
               kSortMax);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
    const size_t W = (size_t)n_shards * cap, n = (size_t)nq * W;
    void* scratch = nullptr;
    GC_CUDA(cudaMallocAsync(&scratch, n * 24 + sizeof(uint32_t) * ((size_t)nq + 4), st));
    double* ld = static_cast<double*>(scratch);
    double* td = ld + n;
    uint32_t* li = reinterpret_cast<uint32_t*>(td + n);
    uint32_t* ti = li + n;
    uint32_t* total = ti + n;
    gather_shard_lists_p2p_kernel<<<dim3(32, nq), 256, 0, st>>>(pl, n_shards, cap, li, ld, total);
    int rc = cudaGetLastError() == cudaSuccess ? GC_OK : GC_E_CUDA;
    count_launch();
    if (rc == GC_OK)
      rc = sort_large(st, nq, (uint32_t)W, W, li, ld, total, ti, td);
    if (rc == GC_OK)
    {
      const uint32_t width = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(k_out, out_cap), W);
      take_first_kernel<<<dim3((std::max(width, 1u) + 255) / 256, nq), 256, 0, st>>>(
          li, ld, total, W, (unsigned long long)k_out, out_cap, d_out_idx, d_out_dist, d_out_count);
      if (cudaGetLastError() != cudaSuccess)
        rc = GC_E_CUDA;
      count_launch();
    }
    cudaFreeAsync(scratch, st);
    if (rc != GC_OK)
      gc::set_error("gc_merge_shard_lists_p2p_dev: large merge failed");
    return rc;
  }
  uint32_t P = 2;
  while (P < n_shards * cap)
    P <<= 1;
  const size_t smem = (size_t)P * (sizeof(double) + sizeof(uint32_t));
  GC_CUDA(cudaFuncSetAttribute(merge_shard_lists_p2p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)(kSortMax * (sizeof(double) + sizeof(uint32_t)))));
  const int threads = P >= 2048 ? 1024 : (P >= 512 ? 256 : 64);
  merge_shard_lists_p2p_kernel<<<nq, threads, smem, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(
      pl, n_shards, nq, cap, (unsigned long long)k_out, out_cap, d_out_idx, d_out_dist, d_out_count, P);
  GC_CUDA(cudaGetLastError());
  count_launch();
  return GC_OK;
}

// stage host queries (and optional per-query segments / radii) on the device
// Queries (and their segments / radii) as the kernels will read them: copied to device scratch, or -- for the small
// batches of a planner that asks one question at a time through the plugin -- left in the pinned staging buffer, which
// the kernels read directly (unified addressing): two or three copy calls fewer on the latency path.
struct StagedQueries
{
  const double* q = nullptr;
  const uint32_t* seg = nullptr;
  const double* radius = nullptr;
};
static constexpr size_t kZeroCopyBytes = 4096;

static int stage_queries(gc_store* s, const double* q, const uint32_t* seg, const double* radius, uint32_t nq,
                         StagedQueries* out)
{
  const size_t qb = sizeof(double) * (size_t)nq * s->dim;
  const size_t sb = seg ? align_up(sizeof(uint32_t) * nq, 16) : 0;
  const size_t rb = radius ? sizeof(double) * nq : 0;
  GC_TRY(s->h_in.reserve(qb + sb + rb));
  GC_TRY(s->d_q.reserve(qb));
  if (seg)
    GC_TRY(s->d_seg.reserve(sb));
  if (radius)
    GC_TRY(s->d_radius.reserve(rb));
  GC_CUDA(cudaStreamSynchronize(s->stream));
  char* h = static_cast<char*>(s->h_in.p);
  const bool zero_copy = qb + sb + rb <= kZeroCopyBytes;
  memcpy(h, q, qb);
  out->q = zero_copy ? reinterpret_cast<const double*>(h) : s->d_q.as<double>();
  if (!zero_copy)
    GC_CUDA(cudaMemcpyAsync(s->d_q.p, h, qb, cudaMemcpyHostToDevice, s->stream));
  if (seg)
  {
    for (uint32_t i = 0; i < nq; i++)
      GC_REQUIRE(seg[i] < s->nseg, "query %u: segment %u of %u", i, seg[i], s->nseg);
    memcpy(h + qb, seg, sizeof(uint32_t) * nq);
    out->seg = zero_copy ? reinterpret_cast<const uint32_t*>(h + qb) : s->d_seg.as<uint32_t>();
    if (!zero_copy)
      GC_CUDA(cudaMemcpyAsync(s->d_seg.p, h + qb, sizeof(uint32_t) * nq, cudaMemcpyHostToDevice, s->stream));
  }
  if (radius)
  {
    memcpy(h + qb + sb, radius, rb);
    out->radius = zero_copy ? reinterpret_cast<const double*>(h + qb + sb) : s->d_radius.as<double>();
    if (!zero_copy)
      GC_CUDA(cudaMemcpyAsync(s->d_radius.p, h + qb + sb, rb, cudaMemcpyHostToDevice, s->stream));
  }
  return GC_OK;
}

int gc_nn1(gc_store_t* s, const double* q, const uint32_t* seg, uint32_t nq, uint32_t* idx, double* dist)
{
  GC_REQUIRE(s && q && idx && dist, "gc_nn1: NULL argument");
  if (nq == 0)
    return GC_OK;
  GC_SET_DEVICE(s->device);
  StagedQueries in;
  GC_TRY(stage_queries(s, q, seg, nullptr, nq, &in));
  GC_TRY(s->d_idx.reserve(sizeof(uint32_t) * nq));
  GC_TRY(s->d_dist.reserve(sizeof(double) * nq));
  const size_t ob = sizeof(double) * nq + sizeof(uint32_t) * nq;
  GC_TRY(s->h_out.reserve(ob));
  double* hd = s->h_out.as<double>();
  uint32_t* hi = reinterpret_cast<uint32_t*>(hd + nq);
  if (ob <= kZeroCopyBytes)
  {
    // latency path: the final (idx, dist) are stored by the kernel straight into the pinned result buffer
    GC_TRY(nn1_dev(s, in.q, in.seg, nq, hi, hd, seg == nullptr));
    GC_CUDA(cudaStreamSynchronize(s->stream));
  }
  else
  {
    GC_TRY(nn1_dev(s, in.q, in.seg, nq, s->d_idx.as<uint32_t>(), s->d_dist.as<double>(), seg == nullptr));
    GC_CUDA(cudaMemcpyAsync(hd, s->d_dist.p, sizeof(double) * nq, cudaMemcpyDeviceToHost, s->stream));
    GC_CUDA(cudaMemcpyAsync(hi, s->d_idx.p, sizeof(uint32_t) * nq, cudaMemcpyDeviceToHost, s->stream));
    GC_CUDA(cudaStreamSynchronize(s->stream));
  }
  memcpy(dist, hd, sizeof(double) * nq);
  memcpy(idx, hi, sizeof(uint32_t) * nq);
  return GC_OK;
}

static int fetch_lists(gc_store* s, uint32_t nq, uint32_t cap, uint32_t* idx, double* dist, uint32_t* count)
{
  const size_t n = (size_t)nq * cap;
  GC_TRY(s->h_out.reserve(sizeof(double) * n + sizeof(uint32_t) * n + sizeof(uint32_t) * nq));
  double* hd = s->h_out.as<double>();
  uint32_t* hi = reinterpret_cast<uint32_t*>(hd + n);
  uint32_t* hc = hi + n;
  GC_CUDA(cudaMemcpyAsync(hc, s->d_count.p, sizeof(uint32_t) * nq, cudaMemcpyDeviceToHost, s->stream));
  GC_CUDA(cudaMemcpyAsync(hd, s->d_dist.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s->stream));
  GC_CUDA(cudaMemcpyAsync(hi, s->d_idx.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, s->stream));
  GC_CUDA(cudaStreamSynchronize(s->stream));
  memcpy(count, hc, sizeof(uint32_t) * nq);
  int rc = GC_OK;
  for (uint32_t i = 0; i < nq; i++)
  {
    const uint32_t m = hc[i] < cap ? hc[i] : cap;
    if (hc[i] > cap)
      rc = GC_W_TRUNCATED;
    memcpy(idx + (size_t)i * cap, hi + (size_t)i * cap, sizeof(uint32_t) * m);
    memcpy(dist + (size_t)i * cap, hd + (size_t)i * cap, sizeof(double) * m);
  }
  return rc;
}

int gc_radius(gc_store_t* s, const double* q, const uint32_t* seg, const double* radius, uint32_t nq, uint32_t cap,
              uint32_t* idx, double* dist, uint32_t* count)
{
  GC_REQUIRE(s && q && radius && idx && dist && count, "gc_radius: NULL argument");
  if (nq == 0)
    return GC_OK;
  GC_SET_DEVICE(s->device);
  StagedQueries in;
  GC_TRY(stage_queries(s, q, seg, radius, nq, &in));
  const size_t n = (size_t)nq * cap;
  GC_TRY(s->d_idx.reserve(sizeof(uint32_t) * n));
  GC_TRY(s->d_dist.reserve(sizeof(double) * n));
  GC_TRY(s->d_count.reserve(sizeof(uint32_t) * nq));
  GC_TRY(radius_dev(s, in.q, in.seg, in.radius, nq, cap, s->d_idx.as<uint32_t>(), s->d_dist.as<double>(),
                    s->d_count.as<uint32_t>()));
  return fetch_lists(s, nq, cap, idx, dist, count);
}

int gc_knn(gc_store_t* s, const double* q, const uint32_t* seg, uint32_t nq, uint64_t k, uint32_t cap, uint32_t* idx,
           double* dist, uint32_t* count)
{
  GC_REQUIRE(s && q && idx && dist && count, "gc_knn: NULL argument");
  if (nq == 0)
    return GC_OK;
  GC_SET_DEVICE(s->device);
  StagedQueries in;
  GC_TRY(stage_queries(s, q, seg, nullptr, nq, &in));
  const size_t n = (size_t)nq * cap;
  GC_TRY(s->d_idx.reserve(sizeof(uint32_t) * n));
  GC_TRY(s->d_dist.reserve(sizeof(double) * n));
  GC_TRY(s->d_count.reserve(sizeof(uint32_t) * nq));
  GC_TRY(knn_dev(s, in.q, in.seg, nq, k, cap, s->d_idx.as<uint32_t>(), s->d_dist.as<double>(),
                 s->d_count.as<uint32_t>()));
  return fetch_lists(s, nq, cap, idx, dist, count);
}

}  // extern "C"
