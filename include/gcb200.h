/*
 * gcb200.h -- C ABI of the B200-native graph_core hot paths.
 *
 * The reference (JRL-CARI-CNR-UNIBS/graph_core) has no FFI: its two hot paths sit
 * behind C++ virtual interfaces.  This header is the NEW boundary underneath the
 * drop-in plugin classes (graph_core_b200/host/):
 *
 *   graph::core::NearestNeighbors      include/graph_core/datastructure/nearest_neighbors.h:41-195
 *     -> CudaNearestNeighbors          -> gc_store_* / gc_nn1 / gc_radius / gc_knn
 *   graph::core::CollisionCheckerBase  include/graph_core/collision_checkers/collision_checker_base.h:118-319
 *     -> CudaCollisionChecker          -> gc_world_* / gc_check_states / gc_check_edges*
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (gc_last_error() explains),
 *     >0 for a non-fatal condition (GC_W_TRUNCATED).  Nothing throws or aborts.
 *   - plain pointers and sizes only.  Unless the name ends in _dev all buffers are
 *     HOST memory owned by the caller; the call returns after the results are in them.
 *   - *_dev entry points take DEVICE pointers, enqueue on the handle's stream and do
 *     not synchronise (for callers that keep data resident in HBM).
 *   - a handle is not thread-safe; distinct handles are independent.
 *   - there is NO CPU fallback: without a CUDA device every create() fails.
 *   - node slots are 32-bit and stable: slot = insertion order inside a segment.
 *     "lowest slot wins" on exactly equal fp64 distances (Vector semantics,
 *     src/graph_core/datastructure/vector.cpp:55-67).
 */
#ifndef GCB200_H_
#define GCB200_H_

#include <stddef.h>
#include <stdint.h>

#include "gcb200_spec.h"

#ifdef __cplusplus
extern "C" {
#endif

#define GC_OK 0
#define GC_W_TRUNCATED 1        /* an output list was longer than the caller's cap; counts hold the true sizes */
#define GC_W_NOOP 2             /* tombstone/restore of a slot already in that state                         */
#define GC_E_INVALID (-1)       /* bad argument                                                              */
#define GC_E_CUDA (-2)          /* CUDA runtime error (message in gc_last_error)                             */
#define GC_E_CAPACITY (-3)      /* store segment full                                                        */
#define GC_E_UNSUPPORTED (-4)   /* combination not implemented on the device path                            */
#define GC_E_NOT_ON_SEGMENT (-5)/* checkConnFromConf: conf not on the parent-child segment (reference throws) */

#define GC_NO_NODE 0xFFFFFFFFu

typedef struct gc_store gc_store_t; /* device-resident SoA node store (one or many trees = segments) */
typedef struct gc_world gc_world_t; /* device-resident immutable collision world, ref-counted       */

/* ---- library -------------------------------------------------------------------- */
int gc_version(void);
const char* gc_last_error(void);
int gc_device_count(int* count);
/* name[0..len) receives the device name; sms and cc (10*major+minor) may be NULL */
int gc_device_info(int device, char* name, int len, int* sms, int* cc);
/* measurement helpers used by bench.py: kernels launched by this library so far; a dependent-chain FFMA
 * micro-benchmark (the FP32 roofline denominator, with the SM clock it implies); an L2 flush that
 * overwrites `bytes` of scratch (bytes > L2 size) */
int gc_launch_count(uint64_t* count);
int gc_measure_fp32_peak(int device, double* tflops, double* sm_mhz_estimate);
int gc_flush_l2(int device, size_t bytes);

/* ---- node store: replaces KdTree/Vector storage (kdtree.cpp:327-342, vector.cpp:38-43) ---- */
int gc_store_create(int dim, uint32_t capacity_per_segment, uint32_t n_segments, int device, gc_store_t** out);
int gc_store_destroy(gc_store_t* s);
/* use an externally owned cudaStream_t (e.g. torch's current stream); NULL restores the private stream.
 * To run on CUDA's legacy default stream pass cudaStreamLegacy ((cudaStream_t)0x1), not NULL. */
int gc_store_set_stream(gc_store_t* s, void* cuda_stream);
/* NearestNeighbors::insert -- appends n rows (row-major, dim doubles each) to one segment */
int gc_store_append(gc_store_t* s, uint32_t segment, const double* q, uint32_t n, uint32_t* first_slot);
/* one row per listed segment (lock-step planners); slots[i] receives the slot of row i */
int gc_store_append_multi(gc_store_t* s, const uint32_t* segments, const double* q, uint32_t n, uint32_t* slots);
/* lazy deletion as KdNode::deleteNode / restoreNode (kdtree.cpp:261-278) */
int gc_store_tombstone(gc_store_t* s, uint32_t segment, uint32_t slot);
int gc_store_restore(gc_store_t* s, uint32_t segment, uint32_t slot);
/* NearestNeighbors::clear; segment == GC_NO_NODE clears every segment */
int gc_store_clear(gc_store_t* s, uint32_t segment);
int gc_store_size(gc_store_t* s, uint32_t segment, uint32_t* slots_used, uint32_t* live);
int gc_store_read(gc_store_t* s, uint32_t segment, uint32_t first_slot, uint32_t n, double* q_out);
int gc_store_dim(gc_store_t* s);

/* ---- queries.  Query i runs against segment seg[i] (seg == NULL: segment 0). ------------- */
/* NearestNeighbors::nearestNeighbor (kdtree.cpp:354-360 / vector.cpp:55-67).
 * Empty segment: idx = GC_NO_NODE, dist = +inf.                                              */
int gc_nn1(gc_store_t* s, const double* q, const uint32_t* seg, uint32_t nq, uint32_t* idx, double* dist);
/* NearestNeighbors::near (kdtree.cpp:362-370 / vector.cpp:69-81): strict dist < radius[i].
 * Query i writes min(count[i], cap) entries at idx/dist[i*cap ...], ascending (dist, slot). */
int gc_radius(gc_store_t* s, const double* q, const uint32_t* seg, const double* radius, uint32_t nq, uint32_t cap,
              uint32_t* idx, double* dist, uint32_t* count);
/* NearestNeighbors::kNearestNeighbors (vector.cpp:83-96): the k closest live nodes, all of
 * them when the segment holds fewer than k.  Same output layout as gc_radius.               */
int gc_knn(gc_store_t* s, const double* q, const uint32_t* seg, uint32_t nq, uint64_t k, uint32_t cap, uint32_t* idx,
           double* dist, uint32_t* count);
/* device-pointer variants (asynchronous on the store's stream) */
int gc_nn1_dev(gc_store_t* s, const double* d_q, const uint32_t* d_seg, uint32_t nq, uint32_t* d_idx, double* d_dist);
int gc_radius_dev(gc_store_t* s, const double* d_q, const uint32_t* d_seg, const double* d_radius, uint32_t nq,
                  uint32_t cap, uint32_t* d_idx, double* d_dist, uint32_t* d_count);
int gc_knn_dev(gc_store_t* s, const double* d_q, const uint32_t* d_seg, uint32_t nq, uint64_t k, uint32_t cap,
               uint32_t* d_idx, double* d_dist, uint32_t* d_count);

/* Node-sharded trees over G GPUs (one process per GPU): node i lives on shard i mod G at local slot
 * i div G; every shard answers the same queries on its slice (gc_nn1_dev with cap 1, gc_knn_dev,
 * gc_radius_dev).  After an all-gather of the per-shard lists ([G][nq][cap] and [G][nq], e.g.
 * torch.distributed.all_gather_into_tensor over NCCL) this merges them on the device: global slot =
 * local*G + shard, ascending (dist, global slot), first k_out entries per query.  G*cap <= 16384.   */
int gc_merge_shard_lists_dev(int device, void* cuda_stream, uint32_t n_shards, uint32_t nq, uint32_t cap,
                             const uint32_t* d_idx, const double* d_dist, const uint32_t* d_count, uint64_t k_out,
                             uint32_t out_cap, uint32_t* d_out_idx, double* d_out_dist, uint32_t* d_out_count);

/* The same merge with the exchange fused in: entry g of the three host arrays is a DEVICE pointer to shard g's own
 * result buffers ([nq][cap], [nq][cap], [nq]), peer-mapped into this process (CUDA IPC / VMM, e.g. torch symmetric
 * memory): the kernel loads the lists over NVLink itself, there is no all-gather and no gathered copy.  The caller
 * places one cross-GPU barrier between the producing kernels and this call, and one before the buffers are reused. */
int gc_merge_shard_lists_p2p_dev(int device, void* cuda_stream, uint32_t n_shards, uint32_t nq, uint32_t cap,
                                 const void* const* d_idx_of_shard, const void* const* d_dist_of_shard,
                                 const void* const* d_count_of_shard, uint64_t k_out, uint32_t out_cap,
                                 uint32_t* d_out_idx, double* d_out_dist, uint32_t* d_out_count);

/* ---- collision worlds (check(q) of collision_checker_base.h:166) --------------------------- */
int gc_world_create_cube(int dim, double threshold, int device, gc_world_t** out);
int gc_world_create_boxes(int dim, uint32_t n, const double* lo, const double* hi, int device, gc_world_t** out);
int gc_world_create_cspheres(int dim, uint32_t n, const float* centers, const float* radii, int device,
                             gc_world_t** out);
/* base_tf[12], joint_tf[njoints][12]: 3x3 rotation row-major then translation; robot sphere k sits at
 * rs_local[k][0..2] (radius rs_local[k][3]) in link frame rs_link[k] in [0, njoints]; obs[nobs][4] = x y z r */
int gc_world_create_sphere_arm(int njoints, const float* base_tf, const float* joint_tf, uint32_t nrs,
                               const int32_t* rs_link, const float* rs_local, uint32_t nobs, const float* obs,
                               int device, gc_world_t** out);
int gc_world_retain(gc_world_t* w);  /* one more owner of the SAME handle (not for concurrent use)         */
/* CollisionCheckerBase::clone (:319): a new handle with its own stream, scratch and counters that shares the
 * immutable device world; handle and clone may be used from different host threads */
int gc_world_clone(gc_world_t* w, gc_world_t** out);
int gc_world_destroy(gc_world_t* w); /* drops one reference */
int gc_world_set_stream(gc_world_t* w, void* cuda_stream);
int gc_world_dim(gc_world_t* w);
/* out[0] states evaluated, out[1] sphere-pair tests executed, out[2] edge-kernel launches,
 * out[3] edge-kernel device time in ns (only while timing is enabled) */
int gc_world_counters(gc_world_t* w, uint64_t out[4]);
int gc_world_reset_counters(gc_world_t* w);
int gc_world_enable_timing(gc_world_t* w, int on);

/* CollisionCheckerBase::check on n configurations: ok[i] = 1 collision free */
int gc_check_states(gc_world_t* w, const double* q, uint32_t n, uint8_t* ok);
/* checkConnection on n segments c1[i]->c2[i].
 * mode GC_EDGE_BISECT (:207-237): ok only.
 * mode GC_EDGE_LINEAR (:178-205): first_bad[i] = index of the first colliding state in the
 *   reference's visiting order (0 = c1, 1..npnt-1 interior, npnt = c2), -1 when free.       */
int gc_check_edges(gc_world_t* w, const double* c1, const double* c2, uint32_t n, double min_distance, int mode,
                   uint8_t* ok, int64_t* first_bad);
/* checkConnFromConf (:264-313): result[i] = 1 free, 0 collision, 255 conf not on the segment */
int gc_check_edges_from_conf(gc_world_t* w, const double* parent, const double* child, const double* conf, uint32_t n,
                             double min_distance, uint8_t* result);
int gc_check_states_dev(gc_world_t* w, const double* d_q, uint32_t n, uint8_t* d_ok);
int gc_check_edges_dev(gc_world_t* w, const double* d_c1, const double* d_c2, uint32_t n, double min_distance,
                       int mode, uint8_t* d_ok, int64_t* d_first_bad);

/* ---- fused planner step: Tree::extend's device part for n lock-step trees ------------------
 * For sample i against segment seg[i]: nearest node (tree.cpp:56-59), steer to at most
 * max_distance (tree.cpp:110-128), bisection edge check closest->next unless dist < tolerance
 * (tree.cpp:69-81) and, when radius != NULL, the radius-near set around next (tree.cpp:751-754).
 * Outputs: closest[i], dist[i] (unclamped), next[i][dim], ok[i]; near lists as gc_radius.     */
int gc_extend_batch(gc_store_t* s, gc_world_t* w, const double* q, const uint32_t* seg, uint32_t n,
                    double max_distance, double min_distance, double tolerance, const double* radius,
                    uint32_t near_cap, uint32_t* closest, double* dist, double* next, uint8_t* ok,
                    uint32_t* near_idx, double* near_dist, uint32_t* near_count);

/* same step with the samples resident in HBM: sample i is row src_row[i] of d_samples (src_row == NULL:
 * row i).  src_row and seg are host arrays; everything else as gc_extend_batch.                      */
int gc_extend_batch_dsamples(gc_store_t* s, gc_world_t* w, const double* d_samples, const uint32_t* src_row,
                             const uint32_t* seg, uint32_t n, double max_distance, double min_distance,
                             double tolerance, const double* radius, uint32_t near_cap, uint32_t* closest,
                             double* dist, double* next, uint8_t* ok, uint32_t* near_idx, double* near_dist,
                             uint32_t* near_count);

/* ---- device-resident lock-step RRT* trees (parent index + cost-to-come on the GPU) ----------
 * A forest advances P independent RRT* problems, problem p owning segment p of `store` (which must have P + 1
 * segments: the last one stays empty and absorbs the queries of inactive problems).  One gc_forest_step is one
 * RRTStar::update (src/graph_core/solvers/rrt_star.cpp:89-163) of every still-improvable problem and runs entirely
 * stream-ordered on the store's stream: nearest -> steer -> extension edge -> radius-near -> Tree::rewireOnly
 * (src/graph_core/graph/tree.cpp:381-513: parent phase in cost order in two edge batches, children phase replayed in
 * near order with Tree::costToNode walks, :767-789) -> goal connection / solution update.  Parent slots and
 * connection costs live in device memory; the call returns without synchronising.  The only host work per step is a
 * stream-ordered callback that evaluates RRTStar::updateRewireRadius' pow() terms (rrt_star.cpp:278-286,
 * informed_sampler.cpp:181-215) with the host libm for the problems whose path cost changed, so radii are bit-identical
 * to the reference's. */
typedef struct gc_forest gc_forest_t;
typedef struct gc_forest_config
{
  double max_distance;     /* tree_solver.cpp:37 */
  double min_distance;     /* checker resolution */
  double tolerance;        /* util.h:80 TOLERANCE */
  double rewire_factor;    /* rrt_star.cpp:46 */
  double utopia_tolerance; /* tree_solver.cpp:40-47, ALREADY 1 + the parameter */
  uint32_t near_cap;       /* longest radius-near list (entries); overflow is reported by gc_forest_sync */
  uint32_t edges_cap;      /* rewiring edges of one step over all problems (second batch) */
  uint32_t depth_cap;      /* longest solution path kept verbatim (nodes) */
  uint32_t first_batch;    /* parent candidates per problem in the first edge batch (0 = 3) */
} gc_forest_config_t;
/* Trees as they stand after TreeSolver::setProblem: problem p has nodes [node_first[p], node_first[p+1]) = its store
 * slots 0.., with parent slot (-1 for the root) and connection cost each; goal_slot[p] = slot of the goal node or -1
 * while it is not in the tree; solved/path_cost as the solver holds them (path_cost = +inf when unsolved). */
int gc_forest_create(gc_store_t* store, gc_world_t* world, const gc_forest_config_t* cfg, const double* lb,
                     const double* ub, uint32_t nproblems, const double* starts, const double* goals,
                     const uint32_t* node_first, const int32_t* node_parent, const double* node_pcost,
                     const int32_t* goal_slot, const uint8_t* solved, const double* path_cost, gc_forest_t** out);
int gc_forest_destroy(gc_forest_t* f);
/* one lock-step iteration; exactly one sample source: host samples [P][dim] (copied H2D on the stream; a pinned buffer
 * must stay untouched until the next gc_forest_sync), d_samples (device) or, both NULL, the forest's own informed
 * sampler (InformedSampler::sample per problem with its current path cost; Philox key `seed`, counter
 * first_index + p).  Does not synchronise. */
int gc_forest_step(gc_forest_t* f, const double* samples, const double* d_samples, uint64_t seed, uint64_t first_index);
/* enqueues a D2H copy of the P current costs (path cost + goal cost, +inf when unsolved) into `costs` (pinned) */
int gc_forest_read_costs_async(gc_forest_t* f, double* costs);
/* waits for the stream; GC_E_CAPACITY (message in gc_last_error) when a near list, the edge list or a segment
 * overflowed since the last call -- the trees are then no longer the reference's */
int gc_forest_sync(gc_forest_t* f);
/* out[0..9]: iterations, extended, edges_device, edges_used, speculation misses, near hits, steps, problems still
 * improvable, problems solved, host-callback radius updates */
int gc_forest_stats(gc_forest_t* f, uint64_t out[16]);
int gc_forest_info(gc_forest_t* f, uint32_t p, int32_t* solved, int32_t* can_improve, double* cost, double* radius,
                   uint32_t* num_nodes, uint32_t* tree_size);
/* ids follow the reference's creation order: 0 start, 1 goal, 2.. new nodes; arrays sized num_nodes */
int gc_forest_dump(gc_forest_t* f, uint32_t p, int32_t* parents, double* pcost, double* conf);
/* node ids root -> goal of the current solution (0 when unsolved); returns the length */
int gc_forest_solution(gc_forest_t* f, uint32_t p, int32_t* ids, uint32_t cap);

/* ---- node-sharded tree over the GPUs of one box ------------------------------------------
 * One process (a graph_core planner is one) drives ndev shards: node i lives on shard i mod ndev at local slot
 * i div ndev.  Queries are broadcast, every shard answers on its slice with the single-store kernels on its own
 * stream, and the merge kernel on shard 0's device reads the other shards' lists through peer-mapped pointers
 * (cudaDeviceEnablePeerAccess: NVLink / NVSwitch loads, no collective library).  Results carry GLOBAL slots and are
 * ordered by (fp64 distance, global slot): what one KdTree / Vector over all nodes returns (kdtree.cpp:354-380,
 * vector.cpp:55-96).  devices == NULL: shard g on device g; several shards may share a device. */
typedef struct gc_comm gc_comm_t;
typedef struct gc_sharded_store gc_sharded_store_t;
int gc_comm_init(int ndev, const int* devices, gc_comm_t** out);
int gc_comm_size(gc_comm_t* c);
int gc_comm_destroy(gc_comm_t* c);
int gc_store_create_sharded(gc_comm_t* c, int dim, uint64_t capacity_total, gc_sharded_store_t** out);
int gc_sharded_store_destroy(gc_sharded_store_t* s);
int gc_sharded_store_size(gc_sharded_store_t* s, uint64_t* nodes, uint32_t* shards);
/* insert n nodes; they receive the global slots *first_global_slot .. + n - 1 (insertion order, as Vector's indices) */
int gc_sharded_store_append(gc_sharded_store_t* s, const double* q, uint32_t n, uint64_t* first_global_slot);
int gc_sharded_nn1(gc_sharded_store_t* s, const double* q, uint32_t nq, uint32_t* idx, double* dist);
int gc_sharded_knn(gc_sharded_store_t* s, const double* q, uint32_t nq, uint64_t k, uint32_t cap, uint32_t* idx,
                   double* dist, uint32_t* count);
/* GC_W_TRUNCATED when a query has more than cap neighbours (count holds the true number), as gc_radius */
int gc_sharded_radius(gc_sharded_store_t* s, const double* q, const double* radius, uint32_t nq, uint32_t cap,
                      uint32_t* idx, double* dist, uint32_t* count);

/* ---- device-resident lock-step BiRRT / RRT (BASELINE config 2) -------------------------------------------------
 * P independent problems advance by one BiRRT::update (src/graph_core/solvers/birrt.cpp:93-152) -- or, with
 * bidirectional = 0, one RRT::update (src/graph_core/solvers/rrt.cpp:119-157) -- per gc_birrt_step: ONE kernel launch,
 * one warp per tree, the whole Tree::connect ray (src/graph_core/graph/tree.cpp:217-233: nearest node, steer, bisection
 * edge check, append, repeat) inside the kernel.  Trees (configurations, parent slots, connection costs) live in device
 * memory owned by the handle; light worlds only (cube, boxes, C-space spheres).  With the same samples every problem
 * builds the reference's two trees, meets at the same iteration and returns the same path and cost. */
typedef struct gc_birrt gc_birrt_t;
typedef struct gc_birrt_config
{
  double max_distance;  /* tree_solver.cpp:37 */
  double min_distance;  /* checker resolution */
  double tolerance;     /* util.h:80 TOLERANCE (0 = 1e-6) */
  int32_t bidirectional; /* 1 BiRRT, 0 RRT (goal connection of rrt.cpp:137-153) */
  int32_t extend;       /* tree_solver.cpp `extend` parameter: 1 = one Tree::extend per tree and update, 0 = connect */
  uint32_t capacity;    /* nodes per tree */
  uint32_t depth_cap;   /* longest solution kept (connections, 0 = 4096) */
} gc_birrt_config_t;
int gc_birrt_create(gc_world_t* world, const gc_birrt_config_t* cfg, uint32_t n_problems, const double* starts,
                    const double* goals, gc_birrt_t** out);
int gc_birrt_destroy(gc_birrt_t* b);
/* one update of every unsolved problem with sample p for problem p: samples[n_problems][dim] on the host (copied
 * through pinned memory) or d_samples on the device; exactly one of the two.  Stream-ordered on the world's stream. */
int gc_birrt_step(gc_birrt_t* b, const double* samples, const double* d_samples);
/* waits for the stream; GC_E_CAPACITY when a tree or a solution outgrew its capacity since creation */
int gc_birrt_sync(gc_birrt_t* b);
/* out[0] updates of unsolved problems, [1] nodes added, [2] edges checked, [3] states evaluated, [4] (state, object)
 * tests, [5] nearest-neighbour scans, [6] re-scans inside a ray, [7] problems solved */
int gc_birrt_stats(gc_birrt_t* b, uint64_t out[8]);
int gc_birrt_info(gc_birrt_t* b, uint32_t p, int32_t* solved, uint32_t* solved_iteration, double* cost,
                  uint32_t* size_start, uint32_t* size_goal);
int gc_birrt_read_solved(gc_birrt_t* b, uint8_t* solved, double* cost);
/* tree 0 = start tree, 1 = goal tree of problem p in slot (= creation) order; returns the number of nodes */
int gc_birrt_dump(gc_birrt_t* b, uint32_t p, int tree, int32_t* parents, double* pcost, double* conf);
/* way points start -> goal and the connection costs between them; returns the number of way points (0 = unsolved) */
int gc_birrt_solution(gc_birrt_t* b, uint32_t p, double* waypoints, double* costs, uint32_t cap);

/* ---- Tree::informedExtend for n trees at once (src/graph_core/graph/tree.cpp:235-302; AnytimeRRT's improve rounds,
 * src/graph_core/solvers/anytime_rrt.cpp:356-427) ---------------------------------------------------------------
 * Query i: sample q[i] against segment seg[i] of `store` with goal[i], cost2beat[i], bias[i].  The candidates are the
 * nearK neighbours (k = ceil(k_rrt * ln(live nodes + 1)), tree.cpp:761-765; k_rrt <= 0: every node) in (distance, slot)
 * order; each is steered towards the sample (tree.cpp:110-128), kept when cost-to-come + distance + |goal - next| <
 * cost2beat and ranked by bias * distance + (1 - bias) * cost-to-come; the first candidate in that order whose
 * distance is below `tolerance` or whose edge is collision free wins.  d_parent / d_pcost are DEVICE arrays indexed by
 * global slot (segment * capacity + slot): parent slot (-1 root / not connected) and the cost of the connection to it,
 * from which Tree::costToNode (tree.cpp:767-789) is walked.  Outputs (host): ok[i], tree_node[i] (slot of the chosen
 * node), new_conf[i][dim], distance[i] (sample to node, unclamped), edge_cost[i] (= |node - new|, the connection cost of
 * extendOnly); the caller appends the node.  Trees of at most 8 192 nodes. */
int gc_informed_extend_batch(gc_store_t* store, gc_world_t* world, const double* q, const uint32_t* seg,
                             const double* goal, const double* cost2beat, const double* bias, uint32_t n, double k_rrt,
                             double max_distance, double min_distance, double tolerance, const int32_t* d_parent,
                             const double* d_pcost, uint8_t* ok, uint32_t* tree_node, double* new_conf,
                             double* distance, double* edge_cost, uint64_t* edges_checked);

/* ---- lock-step AnytimeRRT (src/graph_core/solvers/anytime_rrt.cpp; BASELINE config 5) ---------------------------
 * P independent queries over one collision world, problem p owning segment p of a node store the handle creates.
 * Each gc_anytime_step is one planner update of every live problem: RRT::update (rrt.cpp:116-157: Tree::connect or
 * Tree::extend towards a sample, then the goal connection) during the first phase -- up to FAILED_ITER = 3 rounds of
 * max_iter updates, TreeSolver::solve (tree_solver.cpp:107-130) -- and AnytimeRRT::improveUpdate (:356-427:
 * Tree::informedExtend on the round's tree, goal connection when cheaper) during the improve rounds
 * (AnytimeRRT::improve :269-330: cost2beat = (1 - cost_impr) * path cost, bias -= delta down to 0.1, informed sampler
 * of cost cost2beat, fresh tree; three failed rounds in a row end the query, :134-171).  Creation runs
 * TreeSolver::setProblem (tree_solver.cpp:55-105) including the direct start -> goal ray, whose nodes stay in the tree.
 * Samples: problem p draws its k-th sample from the informed Philox sampler (gc_sample_informed) with counter
 * k * P + p, foci start/goal, cost sampler_cost0[p] (NULL or +inf: uniform in [lb, ub]) in the first phase and
 * cost2beat in a round -- a problem's stream does not depend on the other problems.  With the same stream every
 * problem makes the reference's decisions: rounds, iterations, tree sizes, path and costs are bit-identical.
 * improve_in_solve: the reference's RRT::update leaves can_improve_ = false once the goal is connected (rrt.cpp:150), so
 * AnytimeRRT::solve never enters its own improve loop; 0 reproduces that, 1 runs the loop (improve() is public and is
 * what callers use).  Trees of at most 8192 nodes (capacity). */
typedef struct gc_anytime gc_anytime_t;
typedef struct gc_anytime_config
{
  double max_distance;     /* tree_solver.cpp:37 */
  double min_distance;     /* checker resolution */
  double tolerance;        /* util.h:80 TOLERANCE (0 = 1e-6) */
  double utopia_tolerance; /* tree_solver.cpp:40 (1 is added, as TreeSolver::config does) */
  double delta;            /* anytime_rrt.h:145 (0.1) */
  double cost_impr;        /* anytime_rrt.h:145 (0.1) */
  uint64_t seed;           /* Philox key */
  uint32_t max_iter;       /* updates per round (solve()'s max_iter) */
  uint32_t capacity;       /* nodes per tree, <= 8192 */
  int32_t extend;          /* tree_solver.cpp `extend`: 1 = Tree::extend per update, 0 = Tree::connect */
  int32_t improve_in_solve;
  int32_t keep_trees;      /* keep a copy of the tree that holds each problem's solution (gc_anytime_dump which = 1) */
  int32_t reserved;
} gc_anytime_config_t;
int gc_anytime_create(gc_world_t* world, const gc_anytime_config_t* cfg, const double* lb, const double* ub,
                      uint32_t n_problems, const double* starts, const double* goals, const double* sampler_cost0,
                      gc_anytime_t** out);
int gc_anytime_destroy(gc_anytime_t* a);
/* queries that are a shard of a larger set (problem j on GPU j mod G): sample k of local problem i then uses Philox
 * counter k * stride + index[i] (stride = size of the whole set, index[i] = its position there), so that a query's
 * result does not depend on how the set is sharded.  Default: stride = n_problems, index[i] = i.  Before any step. */
int gc_anytime_set_streams(gc_anytime_t* a, uint64_t stride, const uint64_t* index);
/* one update of every live problem; returns the number of problems still live (< 0: error) */
int gc_anytime_step(gc_anytime_t* a);
/* steps until every query ended or max_steps were made (0 = no limit); returns the number still live */
int gc_anytime_run(gc_anytime_t* a, uint64_t max_steps);
/* out: steps, samples drawn, RRT rows, improve rows, nodes appended, goal edges checked, edges checked inside
 * informedExtend, first solutions found */
int gc_anytime_stats(gc_anytime_t* a, uint64_t out[8]);
/* host wall-clock seconds spent per part of gc_anytime_step since creation: control flow, enqueueing the RRT rows,
 * Tree::informedExtend for the improve rows (includes its round synchronisations), waiting for the device, the
 * decisions on the row results, appends, goal connections */
int gc_anytime_timing(gc_anytime_t* a, double out[8]);
int gc_anytime_info(gc_anytime_t* a, uint32_t p, int32_t* solved, int32_t* done, double* cost, double* bias,
                    double* cost2beat, uint64_t* samples, uint32_t* tree_nodes, uint32_t* log_rows);
int gc_anytime_read_solved(gc_anytime_t* a, uint8_t* solved, double* cost);
/* one row of 8 doubles per finished round: phase (0 RRT, 1 improve), updates used, success, path cost afterwards, nodes
 * in the round's tree (goal included when it succeeded), bias, cost2beat, 1 when the round was set up (0: improve()
 * returned before building a tree); returns the number of rows */
int gc_anytime_log(gc_anytime_t* a, uint32_t p, double* rows, uint32_t cap_rows);
/* way points start -> goal and the connection costs between them; returns the number of way points (0 = unsolved) */
int gc_anytime_solution(gc_anytime_t* a, uint32_t p, double* waypoints, double* costs, uint32_t cap);
/* which = 0: the tree that is growing; 1: the tree that holds the solution (keep_trees).  Slot order = creation order,
 * root first; returns the number of nodes */
int gc_anytime_dump(gc_anytime_t* a, uint32_t p, int which, int32_t* parents, double* pcost, double* conf, uint32_t cap);

/* ---- Philox samplers (uniform_sampler.cpp:34-39, informed_sampler.cpp:145-180) -------------- */
/* n uniform samples in [lb, ub]; sample i uses Philox counter (first_index + i) */
int gc_sample_uniform(int device, int dim, const double* lb, const double* ub, uint64_t seed, uint64_t first_index,
                      uint32_t n, double* out);
/* same samples written to device memory d_out[n][dim] (lb/ub are host arrays) */
int gc_sample_uniform_dev(int device, int dim, const double* lb, const double* ub, uint64_t seed,
                          uint64_t first_index, uint32_t n, double* d_out, void* cuda_stream);

/* InformedSampler::sample (informed_sampler.cpp:145-180) for n problems at once: problem i samples the prolate
 * hyperspheroid with foci f1[i], f2[i] and transverse diameter cost[i] (cost = +inf: uniform in the box), rejecting
 * points outside [lb, ub] up to 100 times.  Philox counter of problem i = first_index + i.  attempts (may be NULL)
 * receives the rounds used: 0 infinite cost, 1..100 accepted round, 101 uniform fallback.                       */
int gc_sample_informed(int device, int dim, const double* lb, const double* ub, const double* f1, const double* f2,
                       const double* cost, uint64_t seed, uint64_t first_index, uint32_t n, double* out,
                       uint32_t* attempts);
/* d_f1, d_f2, d_cost, d_out, d_attempts are device pointers; lb/ub host arrays */
int gc_sample_informed_dev(int device, int dim, const double* lb, const double* ub, const double* d_f1,
                           const double* d_f2, const double* d_cost, uint64_t seed, uint64_t first_index, uint32_t n,
                           double* d_out, uint32_t* d_attempts, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif
