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
struct LightWorld
{
  int kind;
  uint32_t nobj;
  double cube_thr;
  const double* lo;  // smem copies
  const double* hi;
  const float* sc;
  const float* sr2;
};

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
  // C-space hyperspheres, fp32 (include/gcb200_spec.h)
  float qf[D];
#pragma unroll
  for (int i = 0; i < D; i++)
    qf[i] = (float)q[i];
  bool hit = false;
  uint32_t s = 0;
  for (; s < w.nobj && !hit; s++)
  {
    const float* c = w.sc + (size_t)s * D;
    const float t0 = __fsub_rn(qf[0], c[0]);
    float acc = __fmul_rn(t0, t0);
#pragma unroll
    for (int i = 1; i < D; i++)
    {
      const float t = __fsub_rn(qf[i], c[i]);
      acc = __fmaf_rn(t, t, acc);
    }
    hit = acc < w.sr2[s];
  }
  pairs += s;
  return !hit;
}
}  // namespace gc
