// edge_device.cuh -- device functions shared by the edge kernels (edge.cu) and the device-resident bidirectional
// planner (birrt.cu): the reference's state enumeration of an edge (collision_checker_base.h:178-313) and the
// single-state check of the light worlds (cube, boxes, C-space spheres; include/gcb200_spec.h).
#pragma once
#include <math_constants.h>

#include "gc_common.cuh"

namespace gc
{
// =================================================================================================
//  state enumeration
// =================================================================================================
struct EdgeGeom
{
  double dist;      // BISECT/LINEAR: |c2-c1| ; FROM_CONF: |conf-child|
  double aux;       // LINEAR: (double)npnt ; FROM_CONF: this_abscissa
  uint32_t nstates; // states in the reference's visiting order (0: reference throws)
};

template <int D>
__device__ __forceinline__ double norm_diff(const double* a, const double* b)
{
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < GC_MAX_DIM; i++)
    if (i < D)
    {
      const double t = __dsub_rn(a[i], b[i]);
      s = __dadd_rn(s, __dmul_rn(t, t));
    }
  return __dsqrt_rn(s);
}

// number of interior states of the bisection loop: n = 2; while (dist > n*min_d) { n/2 states; n *= 2; }
__device__ __forceinline__ uint32_t bisect_interior(double dist, double min_d)
{
  double n = 2.0;
  uint32_t count = 0;
  while (dist > __dmul_rn(n, min_d) && count < 0x3fffffffu)
  {
    count += (uint32_t)(n * 0.5);
    n *= 2.0;
  }
  return count;
}

// c1/c2/c3 are thread-local copies.  MODE FROM_CONF: c1 = parent, c2 = child, c3 = this_conf
template <int D, int MODE>
__device__ __forceinline__ EdgeGeom edge_geom(const double* c1, const double* c2, const double* c3, double min_d)
{
  EdgeGeom g;
  g.aux = 0.0;
  if (MODE == GC_EDGE_BISECT)
  {
    g.dist = norm_diff<D>(c2, c1);
    g.nstates = 2u + ((g.dist < min_d) ? 0u : bisect_interior(g.dist, min_d));
  }
  else if (MODE == GC_EDGE_LINEAR)
  {
    g.dist = norm_diff<D>(c2, c1);
    const double np = ceil(__ddiv_rn(g.dist, min_d));
    const uint32_t npnt = (np >= 1073741824.0) ? 0x3fffffffu : (uint32_t)np;
    g.aux = (double)npnt;
    g.nstates = (npnt > 1u ? npnt : 1u) + 1u;
  }
  else
  {
    const double dist_child = norm_diff<D>(c3, c2);
    const double dist_parent = norm_diff<D>(c1, c3);
    const double dist = norm_diff<D>(c1, c2);
    if (__dsub_rn(__dsub_rn(dist, dist_child), dist_parent) > 1e-04)
    {
      g.dist = 0.0;
      g.nstates = 0u;  // reference throws std::invalid_argument (:273-277)
      return g;
    }
    g.dist = dist_child;
    g.aux = __ddiv_rn(dist_parent, dist);
    g.nstates = 2u + ((g.dist < min_d) ? 0u : bisect_interior(g.dist, min_d));
  }
  return g;
}

// configuration of state s; returns false when the reference does not evaluate this state
template <int D, int MODE>
__device__ __forceinline__ bool state_conf(const double* c1, const double* c2, const double* c3, const EdgeGeom& g,
                                           uint32_t s, double* conf)
{
  if (MODE == GC_EDGE_LINEAR)
  {
    const uint32_t last = g.nstates - 1u;
    if (s == 0u || s == last)
    {
#pragma unroll
      for (int i = 0; i < GC_MAX_DIM; i++)
        if (i < D)
          conf[i] = (s == 0u) ? c1[i] : c2[i];
      return true;
    }
#pragma unroll
    for (int i = 0; i < GC_MAX_DIM; i++)
      if (i < D)
      {
        const double step = __ddiv_rn(__dsub_rn(c2[i], c1[i]), g.aux);
        conf[i] = __dadd_rn(c1[i], __dmul_rn(step, (double)s));
      }
    return true;
  }
  if (s < 2u)
  {
#pragma unroll
    for (int i = 0; i < GC_MAX_DIM; i++)
      if (i < D)
        conf[i] = (MODE == GC_EDGE_FROM_CONF) ? ((s == 0u) ? c3[i] : c2[i]) : ((s == 0u) ? c1[i] : c2[i]);
    return true;
  }
  const uint32_t v = s - 1u;                 // 1-based interior index in visiting order
  const int L = 31 - __clz((int)v);          // level: n = 2^(L+1)
  const double n = (double)(2u << L);
  const double idx = (double)(2u * (v - (1u << L)) + 1u);
  if (MODE == GC_EDGE_FROM_CONF)
  {
    const double abscissa = __ddiv_rn(idx, n);
    if (!(abscissa >= g.aux))
      return false;
#pragma unroll
    for (int i = 0; i < GC_MAX_DIM; i++)
      if (i < D)
        conf[i] = __dadd_rn(c1[i], __dmul_rn(__dsub_rn(c2[i], c1[i]), abscissa));
    return true;
  }
#pragma unroll
  for (int i = 0; i < GC_MAX_DIM; i++)
    if (i < D)
      conf[i] = __dadd_rn(c1[i], __ddiv_rn(__dmul_rn(__dsub_rn(c2[i], c1[i]), idx), n));
  return true;
}

// =================================================================================================
//  light worlds: true = collision free
// =================================================================================================
// C-space spheres are staged as records of light_rec_floats(D) floats: centre[0..D-1], radius^2, padding to a multiple
// of four floats, so that a state reads a sphere with 128-bit shared-memory loads; the list is padded to a multiple of
// four spheres with radius^2 = -1 (never hit).
__host__ __device__ constexpr int light_rec_floats(int D) { return ((D + 1 + 3) / 4) * 4; }
__host__ __device__ inline uint32_t light_padded_spheres(uint32_t n) { return (n + 3u) & ~3u; }

struct LightWorld
{
  int kind;
  uint32_t nobj;
  double cube_thr;
  const double* lo;  // smem copies
  const double* hi;
  const float* rec;  // C-space spheres: packed records (see above)
};

inline size_t light_smem_bytes(int kind, uint32_t nobj, int D)
{
  if (kind == GC_WORLD_BOXES)
    return 2 * sizeof(double) * (size_t)nobj * D;
  if (kind == GC_WORLD_CSPHERES)
    return sizeof(float) * (size_t)light_padded_spheres(nobj) * light_rec_floats(D);
  return 0;
}

// every thread of the CTA calls this once; the caller synchronises afterwards
template <int D>
__device__ __forceinline__ void light_stage(LightWorld& w, unsigned char* smem_raw, int kind, uint32_t nobj,
                                            double cube_thr, const double* g_lo, const double* g_hi, const float* g_sc,
                                            const float* g_sr2)
{
  w.kind = kind;
  w.nobj = nobj;
  w.cube_thr = cube_thr;
  double* slo = reinterpret_cast<double*>(smem_raw);
  double* shi = slo + (kind == GC_WORLD_BOXES ? (size_t)nobj * D : 0);
  float* rec = reinterpret_cast<float*>(smem_raw);
  if (kind == GC_WORLD_BOXES)
    for (uint32_t i = threadIdx.x; i < nobj * D; i += blockDim.x)
    {
      slo[i] = g_lo[i];
      shi[i] = g_hi[i];
    }
  if (kind == GC_WORLD_CSPHERES)
  {
    constexpr int R = light_rec_floats(D);
    const uint32_t np = light_padded_spheres(nobj);
    for (uint32_t i = threadIdx.x; i < np * R; i += blockDim.x)
    {
      const uint32_t sp = i / R, k = i % R;
      float v = 0.0f;
      if (sp < nobj)
        v = (k < (uint32_t)D) ? g_sc[(size_t)sp * D + k] : (k == (uint32_t)D ? g_sr2[sp] : 0.0f);
      else if (k == (uint32_t)D)
        v = -1.0f;
      rec[i] = v;
    }
  }
  w.lo = slo;
  w.hi = shi;
  w.rec = rec;
}

template <int D>
__device__ __forceinline__ bool light_check(const LightWorld& w, const double* q, unsigned long long& pairs)
{
  if (w.kind == GC_WORLD_CUBE)
  {
    double m = 0.0;
#pragma unroll
    for (int i = 0; i < GC_MAX_DIM; i++)
      if (i < D)
        m = fmax(m, fabs(q[i]));
    return m > w.cube_thr;
  }
  if (w.kind == GC_WORLD_BOXES)
  {
    for (uint32_t b = 0; b < w.nobj; b++)
    {
      bool inside = true;
#pragma unroll
      for (int i = 0; i < GC_MAX_DIM; i++)
        if (i < D)
          inside = inside && !(q[i] < w.lo[b * D + i] || q[i] > w.hi[b * D + i]);
      if (inside)
        return false;
    }
    return true;
  }
  // C-space hyperspheres, fp32 (include/gcb200_spec.h): acc = t0*t0, then fma(t_i, t_i, acc); hit = acc < r^2.
  // Four spheres per trip (independent chains), early exit between trips.
  constexpr int R = light_rec_floats(D);
  float qf[D];
#pragma unroll
  for (int i = 0; i < D; i++)
    qf[i] = (float)q[i];
  bool hit = false;
  uint32_t s = 0;
  const uint32_t np = light_padded_spheres(w.nobj);
  for (; s < np && !hit; s += 4)
  {
#pragma unroll
    for (int u = 0; u < 4; u++)
    {
      float r[R];
      const float4* rp = reinterpret_cast<const float4*>(w.rec + (size_t)(s + u) * R);
#pragma unroll
      for (int k = 0; k < R / 4; k++)
      {
        const float4 v = rp[k];
        r[4 * k + 0] = v.x;
        r[4 * k + 1] = v.y;
        r[4 * k + 2] = v.z;
        r[4 * k + 3] = v.w;
      }
      const float t0 = __fsub_rn(qf[0], r[0]);
      float acc = __fmul_rn(t0, t0);
#pragma unroll
      for (int i = 1; i < D; i++)
      {
        const float t = __fsub_rn(qf[i], r[i]);
        acc = __fmaf_rn(t, t, acc);
      }
      hit = hit || (acc < r[D]);
    }
  }
  pairs += min(s, w.nobj);
  return !hit;
}
}  // namespace gc
