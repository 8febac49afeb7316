/* ipcb200.h — C ABI of the B200-native per-step contact pipeline.
 *
 * This is the drop-in boundary for ONE path of ipc-sim/ipc-toolkit (v1.6.0):
 *
 *   CollisionMesh -> BroadPhase -> Candidates -> NormalCollisions::build
 *                 -> BarrierPotential {(), gradient, hessian}
 *                 -> compute_collision_free_stepsize
 *
 * Every entry point names the reference interface (file:line under the
 * reference tree's src/ipc/) it replaces.  The signatures hold only plain
 * pointers and sizes: no Eigen, no torch, no C++ types.
 *
 * Conventions
 *  - Matrices are COLUMN-major like Eigen::MatrixXd / MatrixXi:
 *    V(i,k) = V[i + k*ld], ld >= rows ("ConstRef" semantics,
 *    utils/eigen_ext.hpp:17).  Index type is int32 (config.hpp.in:30-34).
 *  - The global DOF of vertex v, axis k is 3*v+k (VERTEX_DERIVATIVE_LAYOUT =
 *    RowMajor, config.hpp.in:45).
 *  - Pair lists are returned row-major: pairs[2*i+0], pairs[2*i+1].
 *  - Functions return 0 on success, non-zero on error; ipcb_last_error()
 *    gives the thread-local message (the C++ adapters rethrow it as
 *    std::runtime_error, like log_and_throw_error, utils/logger.cpp:41-45).
 *  - Functions ending in _dev take DEVICE pointers (same layout) and never
 *    touch host memory except for scalar outputs that are documented as host.
 *    The host variants stage through the context's device buffers.
 *  - All calls are synchronous at the API edge (one CUDA stream per context),
 *    matching the blocking TBB-parallel reference entry points.
 *
 * There is NO CPU fallback in this library: ipcb_ctx_create fails when no
 * CUDA device is usable.  The CPU restatement lives in oracle/ (test
 * infrastructure, prefix ipco_) and exports the host half of this ABI.
 */
#ifndef IPCB200_H
#define IPCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef IPCB_PREFIX
#define IPCB_PREFIX ipcb_
#endif
#define IPCB_CAT2(a, b) a##b
#define IPCB_CAT(a, b) IPCB_CAT2(a, b)
#define IPCB_FN(name) IPCB_CAT(IPCB_PREFIX, name)

typedef struct ipcb_ctx ipcb_ctx;
typedef struct ipcb_collision_set ipcb_collision_set;

/* candidate / collision kinds (candidates/candidates.hpp:240-244) */
enum { IPCB_VV = 0, IPCB_EV = 1, IPCB_EE = 2, IPCB_FV = 3, IPCB_EF = 4, IPCB_FF = 5 };

/* PSDProjectionMethod (utils/eigen_ext.hpp:202-206) */
enum { IPCB_PSD_NONE = 0, IPCB_PSD_CLAMP = 1, IPCB_PSD_ABS = 2 };

/* NarrowPhaseCCD implementations (ccd/tight_inclusion_ccd.hpp, ccd/additive_ccd.hpp) */
enum { IPCB_CCD_TIGHT_INCLUSION = 0, IPCB_CCD_ADDITIVE = 1 };

/* broad phase predicate: the default LBVH's outward-rounded FLOAT boxes (broad_phase/lbvh.cpp:29-41, lbvh.hpp:66-70),
 * the only predicate this library implements.  (BruteForce / HashGrid compare double boxes, broad_phase/aabb.cpp:29-33;
 * the oracle can switch to them as a cross-check; this library answers any other value with an error.) */
enum { IPCB_BOXES_FLOAT = 0 };

/* broad-phase method of a context (north_star subsystem 1): the LBVH (default) or a sweep-and-prune over the boxes sorted
 * along the longest scene axis (reference semantics broad_phase/sweep_and_prune.cpp:106-119).  Same predicate, hence the
 * same candidate sets; only the cost differs (bench.py --broad sap). */
enum { IPCB_BROAD_LBVH = 0, IPCB_BROAD_SAP = 1 };

/* flags for collisions_build (collisions/normal/normal_collisions.hpp:191-194: use_area_weighting; :29-39
 * CollisionSetType).  IPCB_SET_IMPROVED_MAX_APPROX selects CollisionSetType::IMPROVED_MAX_APPROX (the negative /
 * positive correction collisions of normal_collisions.cpp:84-128, normal_collisions_builder.cpp:340-543).  Not
 * available on a sharded context (ctx_set_shard with world > 1): the call fails. */
enum { IPCB_USE_AREA_WEIGHTING = 1, IPCB_SET_IMPROVED_MAX_APPROX = 2, IPCB_DEFER_CORRECTIONS = 4 };

/* flags for collisions_merge: the appended builders worked on DISJOINT candidate shards (the ranks of a sharded
 * build), so their edge-edge and face-vertex records are unique across builders and only vertex-vertex /
 * edge-vertex records need uniting.  The result is the same set; the product then skips the global sort of the
 * EE / FV records until somebody asks for them in canonical order (collisions_fetch, collisions_dev_ptrs). */
enum { IPCB_MERGE_DISJOINT_SHARDS = 1 };

/* CCD parameters; defaults are the reference's (tight_inclusion_ccd.hpp:11-19,
 * additive_ccd.hpp:21-26).  A value <= 0 for tolerance / conservative_rescaling
 * or == 0 for max_iterations selects the default of `kind`. */
typedef struct ipcb_ccd_params {
    int32_t kind;                  /* IPCB_CCD_* */
    double tolerance;              /* TI only; default 1e-6 */
    int64_t max_iterations;        /* default 10'000'000; < 0 = unlimited */
    double conservative_rescaling; /* default 0.8 (TI) / 0.9 (Additive) */
} ipcb_ccd_params;

/* BarrierPotential(dhat, stiffness, use_physical_barrier)
 * (potentials/barrier_potential.hpp, barrier_potential.cpp:62-99) */
typedef struct ipcb_barrier_params {
    double dhat;
    double stiffness;
    int32_t use_physical_barrier;
} ipcb_barrier_params;

/* ---- context ------------------------------------------------------------ */
int IPCB_FN(ctx_create)(int device, ipcb_ctx** out);
void IPCB_FN(ctx_destroy)(ipcb_ctx* ctx);
const char* IPCB_FN(last_error)(void);
/* "cuda-sm100a" for the product, "oracle-cpu" for the oracle. */
const char* IPCB_FN(backend_name)(void);
/* CUDA stream handle of the context (cudaStream_t as void*), NULL for oracle. */
void* IPCB_FN(ctx_stream)(ipcb_ctx* ctx);

/* ---- CollisionMesh (collision_mesh.cpp:15-127) -------------------------- */
/* Builds faces_to_edges (:510-543, error "Unable to find edge!"), codim
 * vertex / edge lists (:145-183) and vertex / edge areas (:309-374) on the
 * host and keeps device mirrors.  E is nE x 2, F is nF x 3. */
int IPCB_FN(mesh_set)(ipcb_ctx* ctx, int32_t nV, const double* rest_positions, int32_t ld_rest,
                      int32_t nE, const int32_t* E, int32_t ldE, int32_t nF, const int32_t* F, int32_t ldF);
/* CollisionMesh::can_collide (collision_mesh.hpp:338, CollisionFilter collision_filter.hpp:30-111): which pairs of
 * vertices — and the primitives containing them (broad_phase.cpp:127-202: no shared vertex AND some pair of their
 * vertices can collide) — may become candidates.  The descriptor is the INTERSECTION of the factories of
 * collision_filter.hpp:113-143 that are data, not code:
 *   patch_ids (nV labels, or NULL = off): make_vertex_patches_filter / make_connected_components_filter —
 *       vertices with equal labels never collide;
 *   n_dynamic (< 0 = off): make_static_obstacle_filter — pairs of vertices with index >= n_dynamic never collide.
 * Arbitrary callables, unions and negations cannot cross a C ABI: the toolkit adapter applies those as a host
 * post-filter on the fetched candidates (cpp/ipc_toolkit_adapter.hpp).  The filter applies to every later broad-phase
 * / candidate / collision / step-size build on this context; mesh_set resets it to accept-all.
 * Like the reference, the codimensional vertex-vertex and edge-vertex passes of Candidates::build evaluate the filter
 * on the ids of their re-indexed vertex subsets (candidates.cpp:61,66-77,83-108). */
int IPCB_FN(mesh_set_collision_filter)(ipcb_ctx* ctx, const int32_t* patch_ids /* nV or NULL */, int32_t n_dynamic);
int IPCB_FN(mesh_num_codim_vertices)(ipcb_ctx* ctx, int32_t* n);
int IPCB_FN(mesh_num_codim_edges)(ipcb_ctx* ctx, int32_t* n);
int IPCB_FN(mesh_faces_to_edges)(ipcb_ctx* ctx, int32_t* f2e /* nF x 3 col-major, ld = nF */);
int IPCB_FN(mesh_areas)(ipcb_ctx* ctx, double* vertex_areas /* nV */, double* edge_areas /* nE */);

/* ---- BroadPhase (broad_phase/broad_phase.hpp:19-133) -------------------- */
/* build(V,E,F,r) :12-23 and build(V0,V1,E,F,r) :25-41 of broad_phase.cpp on
 * the context's mesh; detect_* replaces the six pure virtuals (:71-97).
 * `boxes`: IPCB_BOXES_FLOAT (anything else is refused). */
int IPCB_FN(broad_build_static)(ipcb_ctx* ctx, const double* V, int32_t ld, double inflation_radius, int32_t boxes);
int IPCB_FN(broad_build_swept)(ipcb_ctx* ctx, const double* V0, const double* V1, int32_t ld,
                               double inflation_radius, int32_t boxes);
int IPCB_FN(broad_detect)(ipcb_ctx* ctx, int32_t kind, int64_t* count);
int IPCB_FN(broad_fetch)(ipcb_ctx* ctx, int32_t kind, int32_t* pairs /* count x 2 */);
/* vertex boxes of the last build as 6 floats per vertex (min xyz, max xyz) (aabb.cpp:35-87, lbvh.cpp:29-41) */
int IPCB_FN(broad_vertex_boxes)(ipcb_ctx* ctx, void* boxes /* nV x 6 row-major */);

/* ---- Candidates (candidates/candidates.cpp:43-222) ---------------------- */
/* 3D: EE + FV from the main broad phase, VV between codim vertices, EV
 * between codim edges and codim vertices.  counts[k], k = IPCB_VV..IPCB_FV.
 * Pair order: VV (v0,v1), EV (edge,vertex), EE (ea,eb), FV (face,vertex). */
int IPCB_FN(candidates_build_static)(ipcb_ctx* ctx, const double* V, int32_t ld, double inflation_radius,
                                     int64_t counts[4]);
int IPCB_FN(candidates_build_swept)(ipcb_ctx* ctx, const double* V0, const double* V1, int32_t ld,
                                    double inflation_radius, int64_t counts[4]);
int IPCB_FN(candidates_fetch)(ipcb_ctx* ctx, int32_t kind, int32_t* pairs);
/* Replace the resident candidates by caller-provided ones (the reference's
 * public-data Candidates container, candidates.hpp:240-244). */
int IPCB_FN(candidates_set)(ipcb_ctx* ctx, int32_t kind, int64_t count, const int32_t* pairs);

/* ---- NormalCollisions::build (collisions/normal/normal_collisions.cpp) --- */
/* build(mesh,V,dhat,dmin,bp) :20-36 — runs candidates_build_static with
 * r = 0.5*(dhat+dmin), then the candidate overload. */
int IPCB_FN(collisions_build)(ipcb_ctx* ctx, const double* V, int32_t ld, double dhat, double dmin, int32_t flags,
                              int64_t counts[4]);
/* build(candidates,mesh,V,dhat,dmin) :38-158 on the RESIDENT candidates
 * (IPC set type; dedup with weight accumulation, builder.cpp:547-689). */
int IPCB_FN(collisions_build_from_candidates)(ipcb_ctx* ctx, const double* V, int32_t ld, double dhat, double dmin,
                                              int32_t flags, int64_t counts[4]);
/* CollisionSetType::IMPROVED_MAX_APPROX built by SEVERAL builders (ranks, or host threads with their own contexts) over
 * disjoint candidate shards (normal_collisions.cpp:84-128 at builder granularity).  A sub-element pair (candidates.cpp:584-695)
 * can be derived from candidates of several builders while its correction records (normal_collisions_builder.cpp:340-543)
 * must be added once.  collisions_build* with IPCB_SET_IMPROVED_MAX_APPROX | IPCB_DEFER_CORRECTIONS stops after the
 * classification and the builder's unique sub-element pairs (zero counts are returned); then
 *   corrections_keys   n[k] = pairs in the builder's list k (0: VV of EV, 1: EV of EE, 2: EV of FV, 3: VV of FV candidates)
 *   corrections_pack   the four lists, one after the other, as 64-bit keys (opaque: only to be handed to corrections_apply)
 *   corrections_apply  keys: for each list the keys of ALL builders (n[k] of them, duplicates allowed), list after list; the
 *                      builder unites them, adds the corrections of slice `rank` of `world` of every united list and merges
 *                      its records.  The builders' sets are then united with collisions_clear / _append / _merge(flags = 0). */
int IPCB_FN(collisions_corrections_keys)(ipcb_ctx* ctx, int64_t n[4]);
int IPCB_FN(collisions_corrections_pack)(ipcb_ctx* ctx, uint64_t* keys);
int IPCB_FN(collisions_corrections_apply)(ipcb_ctx* ctx, const uint64_t* keys, const int64_t n[4], int32_t rank, int32_t world,
                                          int64_t counts[4]);
/* ids: count x 2 (VV (v0<v1), EV (edge,vertex), EE (ea<eb), FV (face,vertex)),
 * sorted lexicographically; weight: count; eps_x, dtype: EE only (may be NULL) */
int IPCB_FN(collisions_fetch)(ipcb_ctx* ctx, int32_t kind, int32_t* ids, double* weight, double* eps_x,
                              uint8_t* dtype);
/* compute_minimum_distance (normal_collisions.cpp:209-233): min squared distance, +inf if empty */
int IPCB_FN(collisions_min_distance)(ipcb_ctx* ctx, const double* V, int32_t ld, double* min_dist_sqr);
/* The reference's collision containers are public data (normal_collisions.hpp:177-189: vv_collisions ..
 * fv_collisions) and NormalCollisions::build fills them by merging per-thread builders
 * (NormalCollisionsBuilder::merge, builder.cpp:547-689: equal collisions are united and their weights
 * added, weight == 0 dropped).  These three calls expose exactly that step: start an empty set, append the
 * records of any number of builders (= of the other ranks of a sharded build), merge.  The merged set
 * becomes the context's resident set, in canonical order (sorted ids), as after collisions_build.
 * Records: ids count x 2 row-major (as collisions_fetch returns them), weight, and for EE eps_x and dtype. */
int IPCB_FN(collisions_clear)(ipcb_ctx* ctx);
int IPCB_FN(collisions_append)(ipcb_ctx* ctx, int32_t kind, int64_t count, const int32_t* ids, const double* weight,
                               const double* eps_x, const uint8_t* dtype);
int IPCB_FN(collisions_merge)(ipcb_ctx* ctx, double dmin, int32_t flags, int64_t counts[4]);
/* Independent collision sets on one mesh.  A context works on ONE resident set; the reference's NormalCollisions are plain
 * containers of which a caller may hold several per mesh (a lagged set for friction, the sets of a line search).  A
 * collision-set object parks a set outside the context: collisions_swap exchanges the context's resident set with the
 * object's content in O(1) (device buffers change owner, nothing is copied; an empty object receives the resident set and
 * leaves the context without one).  The host mirrors give every NormalCollisions its own object and swap it in when the
 * potential is evaluated on it. */
int IPCB_FN(collision_set_create)(ipcb_ctx* ctx, ipcb_collision_set** out);
void IPCB_FN(collision_set_destroy)(ipcb_collision_set* set);
int IPCB_FN(collisions_swap)(ipcb_ctx* ctx, ipcb_collision_set* set, int64_t counts[4] /* of the now-resident set */);
/* Sharding of the potential over ranks, for contexts that all hold the SAME collision set:
 *  - collision range: energy and gradient only visit the slice [rank*n/world, (rank+1)*n/world) of every
 *    kind's collisions (the results need a sum all-reduce);
 *  - row block: the Hessian is assembled only for the DOF rows (== columns, the matrix is symmetric) of the
 *    vertices [v_begin, v_end); local Hessians are computed for every collision touching such a vertex and
 *    only the owned rows are kept, so the rank matrices tile the global matrix without any exchange.
 *    outer keeps its global length 3nV+1 (empty outside the block).  v_end < 0 selects all vertices. */
int IPCB_FN(ctx_set_collision_range)(ipcb_ctx* ctx, int32_t rank, int32_t world);
int IPCB_FN(ctx_set_row_block)(ipcb_ctx* ctx, int32_t v_begin, int32_t v_end);
/* Row-block boundaries that balance the Hessian work: bounds[0..world] ascending with bounds[0] = 0,
 * bounds[world] = nV, such that every block holds about the same number of 3x3 block contributions of
 * the resident collision set (deterministic: identical on every rank holding the same set). */
int IPCB_FN(hessian_balanced_row_blocks)(ipcb_ctx* ctx, int32_t world, int32_t* bounds /* world + 1 */);

/* ---- BarrierPotential (potentials/potential.cpp:36-222) ----------------- */
int IPCB_FN(barrier_energy)(ipcb_ctx* ctx, const double* V, int32_t ld, const ipcb_barrier_params* bp, double* energy);
int IPCB_FN(barrier_gradient)(ipcb_ctx* ctx, const double* V, int32_t ld, const ipcb_barrier_params* bp,
                              double* grad /* 3*nV */);
/* Assembles H (3nV x 3nV, compressed column == compressed row of the
 * symmetric matrix, indices ascending, duplicates summed, exact-zero local
 * entries skipped: utils/local_to_global.hpp:263-305) and returns nnz. */
int IPCB_FN(barrier_hessian)(ipcb_ctx* ctx, const double* V, int32_t ld, const ipcb_barrier_params* bp,
                             int32_t psd_mode, int64_t* nnz);
int IPCB_FN(barrier_hessian_fetch)(ipcb_ctx* ctx, int32_t* outer /* 3nV+1 */, int32_t* inner /* nnz */,
                                   double* values /* nnz */);

/* ipc::has_intersections(mesh, vertices) (ipc.cpp:105-166), 3D: edge-face candidates of a broad phase inflated by 1e-6 of
 * the world bounding-box diagonal, then is_edge_intersecting_triangle (geometry/intersection.cpp:115-145) with an exact
 * orientation test.  *result = 1 if some edge intersects some triangle (sharing no vertex, passing the collision filter). */
int IPCB_FN(has_intersections)(ipcb_ctx* ctx, const double* V, int32_t ld, int32_t* result);

/* ---- Friction ----------------------------------------- */
/* TangentialCollisions::build(mesh, vertices, collisions, normal_potential, mu_s, mu_k)
 * (collisions/tangential/tangential_collisions.cpp:62-171) from the RESIDENT normal collision set: per collision the
 * lagged closest point, tangent basis (tangent/ *.cpp), normal force magnitude N = -kappa b'(d^2) 2 d
 * (barrier/barrier_force_magnitude.cpp:7-15) and the blended coefficients (default_blend_mu: the average).  Edge-edge
 * collisions that are close to parallel (cross^2 < eps_x) are skipped like in the reference.  mu_s / mu_k: per vertex.
 * Isotropic coefficients only (the anisotropic "matchstick" lagging and the force Jacobians are outside this path). */
int IPCB_FN(tangential_build)(ipcb_ctx* ctx, const double* V, int32_t ld, const ipcb_barrier_params* normal_potential, const double* mu_s,
                              const double* mu_k, int64_t counts[4]);
/* ids count x 2 like collisions_fetch; closest_point count x 2, tangent_basis count x 6 (column 0 then column 1), row-major;
 * records keep the order of the normal collisions; any pointer may be NULL */
int IPCB_FN(tangential_fetch)(ipcb_ctx* ctx, int32_t kind, int32_t* ids, double* weight, double* normal_force, double* mu_s, double* mu_k,
                              double* closest_point, double* tangent_basis);
/* FrictionPotential(eps_v) operator() / gradient / hessian over the resident tangential set (potentials/potential.cpp:36-222
 * with potentials/tangential_potential.cpp:162-325); `velocities` is the potential's argument (nV x 3, column-major).
 * The Hessian lands in the same resident CSR as the barrier Hessian (barrier_hessian_fetch / _dev_ptrs). */
int IPCB_FN(friction_energy)(ipcb_ctx* ctx, const double* velocities, int32_t ld, double eps_v, double* energy);
int IPCB_FN(friction_gradient)(ipcb_ctx* ctx, const double* velocities, int32_t ld, double eps_v, double* grad /* 3*nV */);
int IPCB_FN(friction_hessian)(ipcb_ctx* ctx, const double* velocities, int32_t ld, double eps_v, int32_t psd_mode, int64_t* nnz);

/* ---- CCD (ipc.cpp:45-101, candidates.cpp:252-292) ----------------------- */
/* compute_collision_free_stepsize(mesh,V0,V1,min_distance,bp,ccd):
 * candidates_build_swept with r = 0.5*min_distance, then the earliest time of
 * impact over all candidates (1.0 if there are none). */
int IPCB_FN(ccd_stepsize)(ipcb_ctx* ctx, const double* V0, const double* V1, int32_t ld, double min_distance,
                          const ipcb_ccd_params* ccd, double* step);
/* Candidates::compute_collision_free_stepsize on the RESIDENT candidates */
int IPCB_FN(ccd_stepsize_from_candidates)(ipcb_ctx* ctx, const double* V0, const double* V1, int32_t ld,
                                          double min_distance, const ipcb_ccd_params* ccd, double* step);
/* Candidates::compute_noncandidate_conservative_stepsize (candidates.cpp:294-338) on the RESIDENT candidates:
 * 0.5 * dhat / max |displacement| over the vertices that belong to some candidate; 1 when there are no candidates (not
 * clamped otherwise, like the reference).  displacements: nV x 3 column-major. */
int IPCB_FN(candidates_noncandidate_stepsize)(ipcb_ctx* ctx, const double* displacements, int32_t ld, double dhat, double* step);
/* Candidates::compute_cfl_stepsize (candidates.cpp:340-363) on the RESIDENT candidates: alpha_C = their collision-free
 * step size, alpha_F = the non-candidate bound for V1 - V0; if alpha_F < alpha_C / 2 the full
 * compute_collision_free_stepsize (new swept broad phase; the resident candidates are replaced), else min of the two. */
int IPCB_FN(candidates_cfl_stepsize)(ipcb_ctx* ctx, const double* V0, const double* V1, int32_t ld, double dhat, double min_distance,
                                     const ipcb_ccd_params* ccd, double* step);
/* NarrowPhaseCCD batch (ccd/narrow_phase_ccd.hpp:8-119): n independent
 * queries; x_t0 / x_t1 hold 12 doubles per query (4 points xyz: EE ea0 ea1 eb0
 * eb1; FV p t0 t1 t2; EV p e0 e1 -; VV p0 p1 - -).  hit[i] in {0,1}, toi[i]
 * valid when hit (+inf otherwise). tmax in [0,1]. */
int IPCB_FN(ccd_narrow_phase)(ipcb_ctx* ctx, int32_t kind, int64_t n, const double* x_t0, const double* x_t1,
                              double min_distance, double tmax, const ipcb_ccd_params* ccd, uint8_t* hit,
                              double* toi);

#ifndef IPCB_ORACLE
/* ---- friction, device-resident (see the host forms above) */
/* device-resident forms (no reference equivalent, like the other _dev calls): positions / velocities, the per-vertex
 * coefficients, energy (1 double) and gradient (3 nV doubles) are device pointers; the Hessian stays resident */
int IPCB_FN(tangential_build_dev)(ipcb_ctx* ctx, const double* dV, int32_t ld, const ipcb_barrier_params* normal_potential,
                                  const double* d_mu_s, const double* d_mu_k, int64_t counts[4]);
int IPCB_FN(friction_energy_dev)(ipcb_ctx* ctx, const double* d_velocities, int32_t ld, double eps_v, double* d_energy);
int IPCB_FN(friction_gradient_dev)(ipcb_ctx* ctx, const double* d_velocities, int32_t ld, double eps_v, double* d_grad);
int IPCB_FN(friction_hessian_dev)(ipcb_ctx* ctx, const double* d_velocities, int32_t ld, double eps_v, int32_t psd_mode, int64_t* nnz);

/* ---- device-resident variants (product only) ---------------------------- */
/* V*, grad are DEVICE pointers (col-major, ld); scalar results are written to
 * HOST memory unless the name says _devout, in which case the pointer is a
 * device pointer and the call does not synchronise the stream. */
int IPCB_FN(collisions_build_dev)(ipcb_ctx* ctx, const double* dV, int32_t ld, double dhat, double dmin,
                                  int32_t flags, int64_t counts[4]);
int IPCB_FN(collisions_build_from_candidates_dev)(ipcb_ctx* ctx, const double* dV, int32_t ld, double dhat,
                                                  double dmin, int32_t flags, int64_t counts[4]);
/* device pointers to the resident collision set of one kind (ids: int32 count x 2; eps_x / dtype NULL unless EE);
 * valid until the next collision build / merge on this context */
int IPCB_FN(collisions_dev_ptrs)(ipcb_ctx* ctx, int32_t kind, int64_t* count, const int32_t** d_ids, const double** d_weight,
                                 const double** d_eps_x, const uint8_t** d_dtype);
/* collisions_append with DEVICE arrays (e.g. the all-gathered records of the other ranks) */
int IPCB_FN(collisions_append_dev)(ipcb_ctx* ctx, int32_t kind, int64_t count, const int32_t* d_ids, const double* d_weight,
                                   const double* d_eps_x, const uint8_t* d_dtype);
int IPCB_FN(barrier_energy_dev)(ipcb_ctx* ctx, const double* dV, int32_t ld, const ipcb_barrier_params* bp,
                                double* d_energy /* device, 1 double */);
int IPCB_FN(barrier_gradient_dev)(ipcb_ctx* ctx, const double* dV, int32_t ld, const ipcb_barrier_params* bp,
                                  double* d_grad /* device, 3*nV */);
int IPCB_FN(barrier_hessian_dev)(ipcb_ctx* ctx, const double* dV, int32_t ld, const ipcb_barrier_params* bp,
                                 int32_t psd_mode, int64_t* nnz);
/* device pointers to the resident CSR of the last barrier_hessian (valid until the next one) */
int IPCB_FN(barrier_hessian_dev_ptrs)(ipcb_ctx* ctx, const int32_t** d_outer, const int32_t** d_inner,
                                      const double** d_values);
/* CollisionSetType::IMPROVED_MAX_APPROX on a SHARDED context (normal_collisions.cpp:84-128 at rank granularity).  A
 * sub-element pair (candidates.cpp:584-695) can be derived from candidates of several ranks and its correction records
 * (normal_collisions_builder.cpp:340-543) must be added once, so collisions_build*_dev(…, IPCB_SET_IMPROVED_MAX_APPROX) on a
 * sharded context stops after the rank's unique sub-element keys (it returns zero counts; the host-buffer calls need the
 * explicit IPCB_DEFER_CORRECTIONS flag and have their own corrections_* entry points above).  The ranks then exchange the keys:
 *   corrections_keys_dev   n[k] = keys in the rank's list k (0: VV of EV, 1: EV of EE, 2: EV of FV, 3: VV of FV candidates)
 *   corrections_pack_dev   the four lists, one after the other, into a device buffer of sum(n) 64-bit keys
 *   (all-gather)
 *   corrections_apply_dev  d_keys: for each list the keys of ALL ranks (n[k] of them, duplicates allowed), list after list;
 *                          every rank unites them, adds the corrections of ITS slice of every united list and merges its
 *                          records.  Afterwards the context holds the rank's part of the set: exchange it with
 *                          collisions_pack_dev / _append_packed_dev and unite with collisions_merge(flags = 0) — correction
 *                          records of different ranks can coincide, so the disjoint-shard shortcut does not apply. */
int IPCB_FN(collisions_corrections_keys_dev)(ipcb_ctx* ctx, int64_t n[4]);
int IPCB_FN(collisions_corrections_pack_dev)(ipcb_ctx* ctx, void* d_keys);
int IPCB_FN(collisions_corrections_apply_dev)(ipcb_ctx* ctx, const void* d_keys, const int64_t n[4], int64_t counts[4]);

/* The rank-to-rank exchange format of a sharded build: all records of the resident set in ONE device buffer,
 * so that one all-gather moves them.  Layout for counts n[VV..FV], every array 8-byte aligned, in this order:
 *   ids_vv (8 n0) | w_vv (8 n0) | ids_ev | w_ev | ids_ee | w_ee | eps_ee (8 n2) | ids_fv | w_fv | dtype_ee (n2, padded to 8)
 * = 16 (n0 + n1 + n3) + 24 n2 + pad8(n2) bytes.  pack copies on the context's stream and does not synchronise;
 * append_packed is collisions_append_dev for each kind of one packed buffer. */
int IPCB_FN(collisions_pack_dev)(ipcb_ctx* ctx, void* d_buffer, int64_t capacity_bytes, int64_t* bytes);
int IPCB_FN(collisions_append_packed_dev)(ipcb_ctx* ctx, const void* d_buffer, const int64_t counts[4]);
int IPCB_FN(candidates_build_swept_dev)(ipcb_ctx* ctx, const double* dV0, const double* dV1, int32_t ld,
                                        double inflation_radius, int64_t counts[4]);
int IPCB_FN(ccd_stepsize_dev)(ipcb_ctx* ctx, const double* dV0, const double* dV1, int32_t ld, double min_distance,
                              const ipcb_ccd_params* ccd, double* d_step /* device, 1 double */);
int IPCB_FN(ccd_stepsize_from_candidates_dev)(ipcb_ctx* ctx, const double* dV0, const double* dV1, int32_t ld,
                                              double min_distance, const ipcb_ccd_params* ccd, double* d_step);
/* Multi-GPU sharding: this context processes only the slice
 * [rank*n/world, (rank+1)*n/world) of the Morton-ordered query leaves in the
 * broad phase, hence a disjoint shard of candidates / collisions.  Energy,
 * gradient and step size then need a sum / sum / min all-reduce by the caller
 * (NCCL), the Hessian is the rank's additive contribution. */
int IPCB_FN(ctx_set_shard)(ipcb_ctx* ctx, int32_t rank, int32_t world);
/* IPCB_BROAD_LBVH / IPCB_BROAD_SAP for every later broad-phase build on this context */
int IPCB_FN(ctx_set_broad_phase_method)(ipcb_ctx* ctx, int32_t method);
/* number of kernels this context has launched so far (bench.py gpu_launches) */
int IPCB_FN(ctx_launch_count)(ipcb_ctx* ctx, int64_t* n);
/* per-stage device time of the last call in ms, by stage name; returns the number of stages filled.
 * Stage timing synchronises the stream at every stage boundary, so it is OFF by default: switch it on
 * with ctx_enable_stage_timing(ctx, 1) for a profiling pass (bench.py does, outside its timed region). */
int IPCB_FN(ctx_enable_stage_timing)(ipcb_ctx* ctx, int32_t on);
int IPCB_FN(ctx_stage_times)(ipcb_ctx* ctx, int32_t max_stages, const char** names, float* ms);
/* roofline denominators measured on the context's device (bench.py): the FP64 FMA rate of a register-only kernel with
 * eight independent chains per thread (TFLOP/s, 2 flop per FMA) and the bandwidth of a device-to-device copy kernel
 * over `bytes` (GB/s, read + write); best of `repeats` launches, CUDA events on the context's stream */
int IPCB_FN(measure_fp64_peak)(ipcb_ctx* ctx, int32_t repeats, double* tflops);
int IPCB_FN(measure_copy_bandwidth)(ipcb_ctx* ctx, int64_t bytes, int32_t repeats, double* gbs);
#endif

#ifdef __cplusplus
}
#endif
#endif /* IPCB200_H */
