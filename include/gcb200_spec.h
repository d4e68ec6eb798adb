/*
 * gcb200_spec.h -- numeric SPECIFICATION shared by the CUDA product and the CPU oracle.
 *
 * Only constants and enums live here (no algorithms): the synthetic collision
 * worlds of BASELINE.json (C-space boxes / hyperspheres, sphere-model FK arm)
 * do not exist in the reference, so their check() is defined by
 * this repository.  For collision booleans to be bit-identical on CPU and GPU
 * both sides evaluate the SAME formulas in the SAME operation order with
 * IEEE-754 round-to-nearest single operations and explicit fused multiply-adds
 * (never compiler-contracted ones).  DESIGN.md section "World definitions" spells
 * the formulas out; the product implements them in graph_core_b200/csrc/edge.cu,
 * the oracle independently in oracle/gc_oracle.cpp.
 */
#ifndef GCB200_SPEC_H_
#define GCB200_SPEC_H_

#ifdef __cplusplus
extern "C" {
#endif

/* ---- limits ---------------------------------------------------------------- */
#define GC_MAX_DIM 16          /* state dimension D, reference configs use 2..14 */
#define GC_MAX_JOINTS 16       /* sphere-arm joints (== D of that world)         */
#define GC_MAX_ROBOT_SPHERES 32

/* ---- world kinds ------------------------------------------------------------ */
enum gc_world_kind
{
  GC_WORLD_CUBE = 0,       /* reference Cube3dCollisionChecker (cube_3d_collision_checker.h:61-64), any D */
  GC_WORLD_BOXES = 1,      /* axis-aligned boxes in C-space (BASELINE config 1)                          */
  GC_WORLD_CSPHERES = 2,   /* hyperspheres in C-space (BASELINE config 2)                                */
  GC_WORLD_SPHERE_ARM = 3  /* serial revolute arm, sphere model, vs sphere cloud (BASELINE config 3)     */
};

/* ---- edge-check modes (collision_checker_base.h) ---------------------------- */
enum gc_edge_mode
{
  GC_EDGE_BISECT = 0,    /* checkConnection(c1,c2)        :207-237 */
  GC_EDGE_LINEAR = 1,    /* checkConnection(c1,c2,conf&)  :178-205 */
  GC_EDGE_FROM_CONF = 2  /* checkConnFromConf(conn,conf)  :264-313 */
};

/* ---- fp32 sine/cosine used by the sphere-arm forward kinematics ------------- */
/* q -> j = rintf(q * GC_TWO_OVER_PI); r = fma(-j,P1,q); r = fma(-j,P2,r); r = fma(-j,P3,r)
 * s = r + r*z*(S1 + z*(S2 + z*S3)),  c = 1 - z/2 + z*z*(C1 + z*(C2 + z*C3)),  z = r*r
 * (Horner steps are fused multiply-adds); quadrant = j & 3 swaps/negates.        */
#define GC_TWO_OVER_PI 0.636619772367581343f
#define GC_PIO2_1 1.5703125f
#define GC_PIO2_2 4.837512969970703125e-4f
#define GC_PIO2_3 7.54978995489188216e-8f
#define GC_SIN_S1 (-1.6666654611e-1f)
#define GC_SIN_S2 (8.3321608736e-3f)
#define GC_SIN_S3 (-1.9515295891e-4f)
#define GC_COS_C1 (4.166664568298827e-2f)
#define GC_COS_C2 (-1.388731625493765e-3f)
#define GC_COS_C3 (2.443315711809948e-5f)

/* ---- sphere pair test of the sphere-arm world (all fp32) --------------------------
 * robot sphere (c, rs) vs obstacle sphere (o, ro) collide iff |o-c|^2 < (ro+rs)^2, evaluated in the
 * expanded form (4 fused multiply-adds + 1 compare per pair):
 *   per obstacle : NA = ro*ro - fma(oz,oz, fma(oy,oy, ox*ox))
 *   per sphere   : B  = fma(cz,cz, fma(cy,cy, cx*cx)) - rs*rs ;  c2 = -2*c ;  rs2 = -2*rs
 *   per pair     : t = fma(ro,rs2, fma(oz,c2z, fma(oy,c2y, fma(ox,c2x, B)))) ;  collide iff t < NA
 * 9 flop per pair (8 in the FMAs + the compare) is the unit of bench.py's FP32 roofline.          */
#define GC_FLOP_PER_PAIR 9

/* ---- Philox4x32-10 (Salmon et al. 2011) ------------------------------------- */
#define GC_PHILOX_M0 0xD2511F53u
#define GC_PHILOX_M1 0xCD9E8D57u
#define GC_PHILOX_W0 0x9E3779B9u
#define GC_PHILOX_W1 0xBB67AE85u
/* uniform double in [0,1): ((hi<<32 | lo) >> 11) * 2^-53 */

/* ---- fp32 scan safety margin -------------------------------------------------- */
/* The fp32 distance scan only pre-selects candidates; decisions are made on the
 * fp64 distance.  |d32 - d64| <= GC_SCAN_SLACK_REL * sqrt(D) * max|coordinate|
 * (+ relative 2^-20 of d) is a loose bound of the fp32 rounding error (see
 * DESIGN.md "fp32 scan, fp64 decisions").                                         */
#define GC_SCAN_SLACK_REL 1.0e-6

#ifdef __cplusplus
}
#endif
#endif
