// gc_common.cuh -- shared declarations of the sm_100a implementation behind include/gcb200.h
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/gcb200.h"

namespace gc
{
// ---- error plumbing: C ABI never throws (SURVEY.md 8b "Error conventions") ------------------
void set_error(const char* fmt, ...);
#define GC_CUDA(call)                                                                          \
  do                                                                                           \
  {                                                                                            \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
    {                                                                                          \
      gc::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return GC_E_CUDA;                                                                        \
    }                                                                                          \
  } while (0)
#define GC_REQUIRE(cond, ...)        \
  do                                 \
  {                                  \
    if (!(cond))                     \
    {                                \
      gc::set_error(__VA_ARGS__);    \
      return GC_E_INVALID;           \
    }                                \
  } while (0)
#define GC_TRY(expr)   \
  do                   \
  {                    \
    int r_ = (expr);   \
    if (r_ < 0)        \
      return r_;       \
  } while (0)

// ---- current-context guard ---------------------------------------------------------------------------
// libgcb200.so links the CUDA runtime statically; a host process usually also carries another runtime
// instance (PyTorch's).  Both share the driver's per-thread "current context", so a cudaSetDevice() in
// here could silently redirect the other runtime's next allocation to our device.  Every C-ABI entry
// therefore restores the caller's driver context on exit (cuCtxGetCurrent / cuCtxSetCurrent looked up
// with dlsym, so the library still loads on a machine without libcuda for symbol checks).
struct CtxGuard
{
  void* prev = nullptr;
  bool have = false;
  CtxGuard();
  ~CtxGuard();
};
#define GC_SET_DEVICE(dev)     \
  gc::CtxGuard gc_ctx_guard_;  \
  GC_CUDA(::cudaSetDevice(dev))

// ---- grow-only device / pinned scratch ---------------------------------------------------------
struct DevBuf
{
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes);
  void release();
  template <typename T>
  T* as() { return reinterpret_cast<T*>(p); }
};
struct PinBuf
{
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes);
  void release();
  template <typename T>
  T* as() { return reinterpret_cast<T*>(p); }
};

// cached device blocks for call sequences whose scratch sizes vary from call to call (informed.cu, anytime.cu):
// get() hands out the smallest free cached block that fits (cudaMalloc otherwise), put() / reset() return blocks
struct DevArena
{
  struct Block
  {
    void* p;
    size_t bytes;
    bool in_use;
  };
  std::vector<Block> blocks;
  ~DevArena();
  cudaError_t get_bytes(void** out, size_t bytes);
  template <typename T>
  cudaError_t get(T** out, size_t count)
  {
    void* p = nullptr;
    const cudaError_t e = get_bytes(&p, sizeof(T) * (count ? count : 1));
    *out = static_cast<T*>(p);
    return e;
  }
  void put(void* p)
  {
    for (Block& b : blocks)
      if (b.p == p)
        b.in_use = false;
  }
  void reset();
};

int num_sms(int device);

// every kernel launch site bumps this (bench.py reports it as "gpu_launches")
void count_launch(unsigned n = 1);

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
}  // namespace gc

// ---- node store ----------------------------------------------------------------------------------
// HBM layout (DESIGN.md "Data layout"):
//   x32 : float  [dim][cap_total]   dimension-major SoA, scanned with coalesced float4 loads
//   x64 : double [cap_total][dim]   row-major master copy, gathered for the exact fp64 decisions
//   used: uint32 [n_segments]       slots in use per segment
// segment s owns slots [s*cap_seg, (s+1)*cap_seg); cap_seg is a multiple of 32 so every
// segment row starts on a 128-byte boundary.  A tombstoned slot has x32[0][slot] = +inf.
struct gc_store
{
  int device = 0;
  int dim = 0;
  uint32_t cap_seg = 0;
  uint32_t nseg = 0;
  size_t cap_total = 0;
  float* x32 = nullptr;
  double* x64 = nullptr;
  uint32_t* d_used = nullptr;
  std::vector<uint32_t> used;      // host mirror of d_used
  uint32_t used_bound = 0;         // > 0: a device-side owner (gc_forest) appends on the device; the host mirror is
                                   // stale until gc_forest_sync and this is an upper bound of every segment's count
  std::vector<uint32_t> live;      // live (non tombstoned) nodes per segment
  std::vector<double> max_abs;     // per segment max |coordinate| ever appended (fp32 slack bound)
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  int sms = 148;
  uint32_t near_width_hint = 64;   // columns of the near lists fetched eagerly by gc_extend_batch
  uint32_t knn_cand_scale = 3;     // head-room factor of the large-N k-nearest candidate lists
  bool knn_force_fp64 = false;     // large-N k-nearest: candidates carry fp64 distances (massive near-ties met)
  gc::DevBuf d_cand_idx, d_cand_dist, d_cand_count, d_thr, d_gmin;
  size_t stream_ctl_q = 0, stream_ctl_g = 0;  // initialised extent of d_gmin (streaming-scan control words)
  // one-time kernel attribute opt-ins (dynamic shared memory) and occupancy look-ups: per store = per device context
  int stream_occ[GC_MAX_DIM + 1] = {};
  bool sort_attr_set = false;
  bool knn_attr_set[GC_MAX_DIM + 1] = {};
  size_t small_knn_err_words = 0;  // layout for which the error word of the two-launch k-nearest path is known clear
  // scratch
  gc::DevBuf d_q, d_seg, d_radius, d_idx, d_dist, d_count, d_part, d_misc, d_rows, d_slots;
  gc::PinBuf h_in, h_out;
};

// ---- collision world ---------------------------------------------------------------------------
struct gc_world
{
  int device = 0;
  int kind = 0;
  int dim = 0;
  int refs = 1;
  gc_world* parent = nullptr;  // gc_world_clone: the world that owns the immutable device data
  // cube
  double cube_thr = 1.0;
  // generic object count (boxes / cspheres / obstacle spheres)
  uint32_t nobj = 0;
  double* d_lo = nullptr;  // boxes [n][dim]
  double* d_hi = nullptr;
  float* d_sc = nullptr;   // cspheres centres [n][dim]
  float* d_sr2 = nullptr;  // cspheres radius^2 [n]
  // sphere arm
  int njoints = 0;
  uint32_t nrs = 0;
  float* d_arm = nullptr;  // base_tf[12] | joint_tf[J][12] | rs_local[nrs][4] | rs_link[nrs] (as float bits)
  float* d_obs = nullptr;  // [nobs][4], then [nblk][4] bounding spheres of the blocks of 16 (see edge.cu, culling)
  uint32_t nblk = 0;       // 0 = no block culling for this world
  // uniform grid over the part of the workspace where a robot sphere can touch an obstacle (edge.cu, "grid culling"):
  // cell -> the obstacles no sphere centred in the cell can rule out.  grid_cells == 0: no grid for this world
  uint32_t* d_grid_start = nullptr;  // [grid_cells][2] cell records (count + up to three inline indices, or the list offset)
  uint16_t* d_grid_items = nullptr;  // obstacle indices (positions in d_obs)
  uint32_t grid_cells = 0;
  int grid_n[3] = { 0, 0, 0 };
  float grid_lo[3] = { 0, 0, 0 };
  float grid_inv_h = 0.0f;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  int sms = 148;
  // counters (device) : [0] states evaluated, [1] pair tests
  unsigned long long* d_counters = nullptr;
  uint64_t launches = 0;
  bool timing = false;
  bool inputs_in_host_memory = false;  // set by the host latency path around check_edges_dev (see edge.cu)
  uint64_t kernel_ns = 0;
  std::vector<cudaEvent_t> ev_pool;  // (begin, end) pairs recorded around the check kernels
  size_t ev_used = 0;
  // scratch
  gc::DevBuf d_c1, d_c2, d_c3, d_ok, d_first, d_plan, d_offs, d_blocksum;
  gc::PinBuf h_in, h_out;
};

namespace gc
{
// internal device-pointer entry points shared by the host wrappers and the fused planner step
int nn1_dev(gc_store* s, const double* d_q, const uint32_t* d_seg, uint32_t nq, uint32_t* d_idx, double* d_dist,
            bool single_segment_host_hint);
int radius_dev(gc_store* s, const double* d_q, const uint32_t* d_seg, const double* d_radius, uint32_t nq,
               uint32_t cap, uint32_t* d_idx, double* d_dist, uint32_t* d_count);
int knn_dev(gc_store* s, const double* d_q, const uint32_t* d_seg, uint32_t nq, uint64_t k, uint32_t cap,
            uint32_t* d_idx, double* d_dist, uint32_t* d_count);
int check_edges_dev(gc_world* w, const double* d_c1, const double* d_c2, const double* d_c3, uint32_t n,
                    double min_distance, int mode, uint8_t* d_ok, int64_t* d_first_bad, uint64_t states_hint,
                    cudaStream_t stream, const uint32_t* d_n = nullptr);
int check_states_dev(gc_world* w, const double* d_q, uint32_t n, uint8_t* d_ok, cudaStream_t stream);
double store_slack(const gc_store* s);
int steer_launch(const gc_store* s, const double* d_q, const uint32_t* d_seg, const uint32_t* d_closest,
                 const double* d_dist, uint32_t n, double max_distance, double* d_c1, double* d_c2, cudaStream_t st);
int sample_informed_launch(int dim, uint32_t n, const double* d_lb, const double* d_ub, const double* d_f1,
                           const double* d_f2, const double* d_cost, unsigned long long seed,
                           unsigned long long first, double* d_out, cudaStream_t st);
// Tree::informedExtend for n queries with device inputs and outputs (informed.cu); the outputs live in `arena`
struct IeOut
{
  uint8_t* d_ok;
  uint32_t* d_node;
  double *d_new, *d_dist, *d_ecost;
};
int informed_extend_core(gc_store* s, gc_world* w, const double* d_q, const uint32_t* d_seg, const uint32_t* h_seg,
                         const double* d_goal, const double* d_c2b, const double* d_bias, uint32_t n, double k_rrt,
                         double max_distance, double min_distance, double tolerance, const int32_t* d_parent,
                         const double* d_pcost, DevArena& arena, IeOut* out, uint64_t* edges_checked,
                         const double* d_c2n = nullptr);
// row i samples for problem d_prob[i] (rows d_prob[i] of d_f1 / d_f2 / d_out, cost d_cost[i]) from Philox counter d_ctr[i]
int sample_informed_rows_launch(int dim, uint32_t n, const double* d_lb, const double* d_ub, const double* d_f1,
                                const double* d_f2, const double* d_cost, unsigned long long seed,
                                const uint32_t* d_prob, const unsigned long long* d_ctr, double* d_out, cudaStream_t st,
                                const double* d_basis = nullptr);
// d_basis[n][dim * dim + dim + 1]: the rotation basis, centre and focal distance of n (f1, f2) pairs
int informed_basis_launch(int dim, uint32_t n, const double* d_f1, const double* d_f2, double* d_basis, cudaStream_t st);
// floats between consecutive rows of the fp32 SoA (cap_total + the padding the scans read past the last tile)
constexpr int kStoreRowPad = 1024;
static inline size_t store_row_stride(const gc_store* s) { return s->cap_total + kStoreRowPad; }
int refresh_store_mirror(gc_store* s);  // d_used -> used/live after device-side appends (synchronises)
}  // namespace gc
