/* shamb200.h — C ABI of the B200-native SPH timestep backend (libshamb200.so).
 *
 * The reference (Shamrock, /root/reference, commit 3ddd3ab4) has no FFI for this path: it is
 * extended by C++ templates compiled in, by pybind11 modules and by solvergraph INode subclasses
 * (SURVEY.md §8b).  This header is therefore the boundary a Shamrock maintainer would bind
 * against (INTEGRATION.md shows the INode / pybind stubs).  Every entry point cites the reference
 * function it replaces (paths relative to /root/reference/src).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no C++ or torch types.
 *  - "d_" pointers are DEVICE pointers owned by the caller; the library never frees them.
 *    Output arrays reachable through shamb200_tree / shamb200_csr views are owned by the context
 *    arena and stay valid until the next call that rebuilds them or the context is destroyed.
 *  - vec3 fields are arrays of doubles with a stride given in doubles (3 = packed 24 B as in
 *    the reference's serialised form, 4 = 32 B like sycl::vec<f64,3>).
 *  - every function returns 0 on success, a negative code on error; shamb200_last_error()
 *    returns the message (thread local).  No exception crosses the boundary.  There is no CPU
 *    fallback: without a CUDA device every compute entry point fails with SHAMB200_ERR_CUDA.
 *  - one context per GPU, driven by one host thread; all work of a context is enqueued on the
 *    context's stream (shamb200_ctx_stream) unless a function says it synchronises.
 */
#ifndef SHAMB200_H
#define SHAMB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHAMB200_OK 0
#define SHAMB200_ERR_INVALID -1
#define SHAMB200_ERR_CUDA -2
#define SHAMB200_ERR_OVERFLOW -3
#define SHAMB200_ERR_NCCL -4
#define SHAMB200_ERR_RUNTIME -5

typedef struct shamb200_ctx shamb200_ctx;
typedef struct shamb200_model shamb200_model;

/* SPH kernels: shammath/include/shammath/sphkernels.hpp:29-82 (M4), :265-346 (M6) */
enum { SHAMB200_KERNEL_M4 = 0, SHAMB200_KERNEL_M6 = 1 };
/* sort backends of shamalgs::algorithm::sort_by_key (shamalgs/src/primitives/sort_by_keys.cpp:36-44):
 * BITONIC reproduces the reference's network bit for bit (incl. the order inside equal-key runs);
 * RADIX is a stable LSD radix sort (same keys, ties in input order). */
enum { SHAMB200_SORT_BITONIC = 0, SHAMB200_SORT_RADIX = 1 };
enum { SHAMB200_EOS_ADIABATIC = 0, SHAMB200_EOS_ISOTHERMAL = 1, SHAMB200_EOS_LOCALLY_ISOTHERMAL_LP07 = 2 };
enum { SHAMB200_AV_NONE = 0, SHAMB200_AV_CONSTANT = 1, SHAMB200_AV_MM97 = 2, SHAMB200_AV_CD10 = 3, SHAMB200_AV_CONSTANT_DISC = 4 };
enum { SHAMB200_BC_FREE = 0, SHAMB200_BC_PERIODIC = 1 };
/* floating-point contract of the SPH loops of the model path.  STRICT: the reference's expressions in
 * the reference's order, no FMA contraction: every float64 output is bit-identical to a CPU run of the
 * reference algorithms (the oracle).  FAST: same algorithm, pair math restructured (per-particle
 * reciprocals, rsqrt, FMA, several lanes per particle): within 1e-10 relative per particle.  Integer
 * outputs (Morton codes, sort permutation, tree, neighbour lists) are exact in both modes. */
enum { SHAMB200_FP_STRICT = 0, SHAMB200_FP_FAST = 1 };

const char *shamb200_last_error(void);
/* library / build information ("sm_100a", strict-fp flag, ...) */
const char *shamb200_build_info(void);
/* number of CUDA kernels launched by this library since the last reset (process wide) */
uint64_t shamb200_launch_count(void);
void shamb200_reset_launch_count(void);

/* ---- context --------------------------------------------------------------------------------
 * replaces: shamsys::instance (device + queue selection, shamsys/src/NodeInstance.cpp) and
 * sham::DeviceScheduler for this path.  `stream` may be NULL (the context creates its own). */
int shamb200_ctx_create(int device, void *cuda_stream, shamb200_ctx **out);
int shamb200_ctx_destroy(shamb200_ctx *ctx);
void *shamb200_ctx_stream(shamb200_ctx *ctx);
int shamb200_ctx_synchronize(shamb200_ctx *ctx);

/* roofline denominators measured on this device (pattern of the reference's own micro-benchmarks,
 * shamsys/src/MicroBenchmark.cpp:51-77): what = 0 FP64 FMA chains [TFLOP/s, FMA = 2 flop],
 * what = 1 streaming copy [GB/s, read + write]; what = 10 + p / 20 + p: warp-wide gathers of 32-byte records
 * (one 256-bit load per lane, the access pattern of the SPH neighbour loops) from an L1- / L2-resident table
 * [G records/s]: p = 0 coalesced, 1 random record per lane, 2 random line per lane with bank group = lane & 3,
 * 3 random line per lane in one bank group, 4 four lanes per random line. */
int shamb200_microbench(shamb200_ctx *ctx, int what, double *out);

/* ---- tree (shamtree::CompressedLeafBVH<u32, f64_3, 3>) ----------------------------------------
 * Device-resident output of rebuild_from_positions; contract of SURVEY.md §3.3. */
typedef struct shamb200_tree {
    uint32_t obj_cnt;      /* M                                    */
    uint32_t morton_count; /* P2 = roundup_pow2(M)                 */
    uint32_t leaf_count;   /* L                                    */
    uint32_t int_count;    /* I = L - 1                            */
    double bmin[3], bmax[3];
    const uint32_t *d_sorted_morton;   /* [P2] (pad = 0xFFFFFFFF)  */
    const uint32_t *d_sort_index_map;  /* [P2] map_morton_id_to_obj_id */
    const uint32_t *d_reduc_index_map; /* [L+2] leaf starts, then M, then 0 */
    const uint32_t *d_reduced_morton;  /* [L]                      */
    const uint32_t *d_lchild_id, *d_rchild_id, *d_endrange; /* [I] */
    const uint8_t *d_lchild_flag, *d_rchild_flag;           /* [I] 1 = leaf */
    const double *d_aabb_min, *d_aabb_max; /* [(I+L)*3] packed, internal cells first */
} shamb200_tree;

/* replaces shamtree::CompressedLeafBVH::rebuild_from_positions (shamtree/src/CompressedLeafBVH.cpp:79-95):
 * Morton codes (MortonCodeSet.cpp:61-128) -> key/value sort (MortonCodeSortedSet.cpp:24-49) ->
 * leaf compression (MortonReducedSet.cpp:24-72, kernels/reduction_alg.cpp) -> Karras tree
 * (KarrasRadixTree.cpp:46-183) -> AABBs (KarrasRadixTreeAABB.cpp:93-135).
 * Synchronises once (the leaf count is needed on the host, as in the reference). */
int shamb200_tree_build(shamb200_ctx *ctx, const double *d_xyz, size_t stride_dbl, uint32_t obj_cnt,
                        const double bmin[3], const double bmax[3], uint32_t reduction_level,
                        int sort_mode, shamb200_tree *out);
/* same, with the bounding box computed like modules::BuildTrees::build_merged_pos_trees
 * (shammodels/sph/src/modules/BuildTrees.cpp:37-56): min/max of the positions widened by one ulp */
int shamb200_tree_build_auto_bbox(shamb200_ctx *ctx, const double *d_xyz, size_t stride_dbl,
                                  uint32_t obj_cnt, uint32_t reduction_level, int sort_mode,
                                  shamb200_tree *out);
/* replaces shamtree::compute_tree_field_max_field<f64> (KarrasRadixTreeField.hpp:189-222) followed by
 * the `*= htol` of Solver::compute_presteps_rint (shammodels/sph/src/Solver.cpp:1322-1356).
 * d_out: [(I+L)] doubles, caller owned. */
int shamb200_tree_field_max(shamb200_ctx *ctx, const shamb200_tree *tree, const double *d_field,
                            double scale, double *d_out);

/* ---- neighbour cache (shamrock::tree::ObjectCache, shamtree/include/shamtree/TreeTraversal.hpp:375-485) */
typedef struct shamb200_csr {
    uint32_t obj_cnt;            /* N                         */
    uint32_t sum_neigh_cnt;      /* K (u32 like the reference) */
    const uint32_t *d_cnt_neigh;   /* [N] */
    const uint32_t *d_scanned_cnt; /* [N] exclusive scan */
    const uint32_t *d_index_neigh_map; /* [K] */
} shamb200_csr;

/* replaces modules::NeighbourCache::start_neighbors_cache_2stages / start_neighbors_cache
 * (shammodels/sph/src/modules/NeighbourCache.cpp:223-604 / :30-220).  d_rint = output of
 * shamb200_tree_field_max(h, htol).  Lists are bit-identical to the reference's (ascending rank in
 * the sorted Morton array).  Synchronises (list sizing).  Fails with SHAMB200_ERR_OVERFLOW when the
 * total count does not fit u32 (the reference's sum_neigh_cnt is u32). */
int shamb200_neigh_cache_build(shamb200_ctx *ctx, const shamb200_tree *tree, const double *d_xyz,
                               size_t stride_dbl, const double *d_hpart, const double *d_rint,
                               uint32_t obj_cnt, double Rkern, double h_tolerance, int two_stage,
                               shamb200_csr *out);
/* How the last two-stage shamb200_neigh_cache_build of this context went: out = {sum of the list lengths,
 * (particle, candidate) accept tests, attempts (1 = every internal capacity was large enough; more = the
 * candidate / list arrays or the tree-walk frontier were regrown and the search repeated), leaf groups walked
 * with a frontier in global memory (objects with a very large h), shared-memory frontier capacity, capacity of
 * the candidate-entry array}.  Diagnostics of the B200 search (no reference counterpart). */
int shamb200_neigh_cache_stats(shamb200_ctx *ctx, uint64_t out[6]);

/* ---- smoothing length, density ------------------------------------------------------------------
 * replaces modules::IterateSmoothingLengthDensity::_impl_evaluate_internal (one Newton sweep,
 * shammodels/sph/src/modules/IterateSmoothingLengthDensity.cpp:28-120).  d_h_new / d_eps in place. */
int shamb200_h_iterate(shamb200_ctx *ctx, int kernel, const shamb200_csr *csr, const double *d_xyz,
                       size_t stride_dbl, const double *d_h_old, double *d_h_new, double *d_eps,
                       double gpart_mass, double h_evol_max, double h_evol_iter_max);
/* replaces modules::LoopSmoothingLengthIter (LoopSmoothingLengthIter.cpp:29-84): up to max_sweeps
 * sweeps with the max-eps test after each; out3 = {max_eps, min_eps, sweeps done}. */
int shamb200_h_iterate_loop(shamb200_ctx *ctx, int kernel, const shamb200_csr *csr, const double *d_xyz,
                            size_t stride_dbl, const double *d_h_old, double *d_h_new, double *d_eps,
                            double gpart_mass, double h_evol_max, double h_evol_iter_max,
                            double epsilon_h, uint32_t max_sweeps, double out3[3]);
/* replaces modules::NodeComputeOmega (shammodels/sph/src/modules/ComputeOmega.cpp:25-74) */
int shamb200_compute_omega(shamb200_ctx *ctx, int kernel, const shamb200_csr *csr, const double *d_xyz,
                           size_t stride_dbl, const double *d_hpart, double *d_omega, double gpart_mass);

/* ---- SPH modules on merged patch data (stage level) ------------------------------------------------
 * The modules Solver::evolve_once runs between the h iteration and the corrector, one entry point per
 * reference module, on the arrays the reference's solver graph hands them: the MERGED fields of one
 * patch (its N real objects first, then the ghosts: M objects, BasicSPHGhosts.hpp:294-514 /
 * Solver.cpp:1394-1633) and the patch's ObjectCache.  Device pointers, caller owned.  Results are
 * bit-identical to the reference's expressions evaluated without FMA contraction (the oracle);
 * shamb200_model_evolve_once runs the fused Morton-ordered versions of the same loops.
 * A NULL input pointer stands for a field the module does not read (documented per function). */
typedef struct shamb200_merged_fields {
    uint32_t obj_cnt;          /* M, merged objects                                       */
    uint32_t real_cnt;         /* N <= M, the patch's own objects (outputs have N entries) */
    const double *d_xyz;       /* [M * stride_dbl]                                         */
    size_t stride_dbl;         /* 3 or 4                                                   */
    const double *d_hpart;     /* [M]                                                      */
    const double *d_vxyz;      /* [M * 3] packed                                           */
    const double *d_uint;      /* [M]                                                      */
    const double *d_axyz;      /* [M * 3] packed (only d(div v)/dt reads it)               */
    const double *d_omega;     /* [M]                                                      */
    const double *d_pressure;  /* [M]                                                      */
    const double *d_soundspeed;/* [M]                                                      */
    const double *d_alpha_AV;  /* [M] (MM97 / CD10); NULL with a constant alpha            */
} shamb200_merged_fields;

/* replaces modules::ComputeEos::compute_eos (shammodels/sph/src/modules/ComputeEos.cpp:1138-1308;
 * adiabatic :147-248, isothermal :54-130, locally isothermal LP07 :724-800) over the merged range.
 * Reads xyz (LP07), hpart, uint (adiabatic).  d_pressure / d_soundspeed: [M]. */
int shamb200_compute_eos(shamb200_ctx *ctx, int kernel, int eos, const shamb200_merged_fields *f,
                         double gpart_mass, double gamma, double cs0, double eos_q, double eos_r0,
                         double *d_pressure, double *d_soundspeed);
/* replaces modules::DiffOperators::update_divv / update_curlv (DiffOperator.cpp:25-146, :148-260).
 * Reads xyz, hpart, vxyz, omega.  d_divv [N]; d_curlv [N * 3] or NULL. */
int shamb200_update_divv_curlv(shamb200_ctx *ctx, int kernel, const shamb200_csr *csr,
                               const shamb200_merged_fields *f, double gpart_mass, double *d_divv,
                               double *d_curlv);
/* replaces modules::DiffOperatorDtDivv::update_dtdivv (DiffOperatorDtDivv.cpp:29-353).  Reads xyz, hpart,
 * vxyz, axyz.  also_divv_curlv != 0: the combined variant that also writes div v and curl v from the
 * same velocity-gradient matrix (combined_dtdiv_divcurlv_compute, SolverConfig.hpp:602). */
int shamb200_update_dtdivv(shamb200_ctx *ctx, int kernel, const shamb200_csr *csr,
                           const shamb200_merged_fields *f, double gpart_mass, int also_divv_curlv,
                           double *d_divv, double *d_curlv, double *d_dtdivv);
/* replaces modules::UpdateViscosity::update_artificial_viscosity (UpdateViscosity.cpp:52-119 MM97,
 * :122-222 CD10).  All arrays [N] (d_curlv [N * 3]); d_dtdivv / d_curlv are only read for CD10. */
int shamb200_update_viscosity(shamb200_ctx *ctx, int av, uint32_t real_cnt, double dt, double sigma_decay,
                              double alpha_min, double alpha_max, const double *d_divv,
                              const double *d_curlv, const double *d_dtdivv, const double *d_soundspeed,
                              const double *d_hpart, const double *d_alpha_AV, double *d_alpha_AV_updated);
/* replaces modules::UpdateDerivs::update_derivs (UpdateDerivs.cpp:89-288 constant alpha, :580-780 disc;
 * NodeUpdateDerivsVaryingAlphaAV.cpp:26-137 MM97 / CD10; math/forces.hpp:27-226, math/q_ab.hpp:37-62).
 * Reads xyz, hpart, vxyz, uint, omega, pressure, soundspeed and alpha_AV (MM97 / CD10).
 * d_axyz [N * 3] = pressure + viscosity forces + d_axyz_ext (NULL = 0); d_duint [N]. */
int shamb200_update_derivs(shamb200_ctx *ctx, int kernel, int av, const shamb200_csr *csr,
                           const shamb200_merged_fields *f, double gpart_mass, double alpha_u,
                           double alpha_AV, double beta_AV, const double *d_axyz_ext, double *d_axyz,
                           double *d_duint);
/* replaces the signal-velocity loop and the CFL time step of Solver::evolve_once (Solver.cpp:2677-2840,
 * :2895-3119; modules/ComputeCFLCourant.hpp, ComputeCFLForce.hpp).  Reads xyz, hpart, vxyz, soundspeed
 * and d_axyz [N * 3] (the new accelerations).  d_vsig, d_cfl_dt: [N]; *dt_min = min over the patch.
 * Synchronises. */
int shamb200_vsig_cfl(shamb200_ctx *ctx, int kernel, const shamb200_csr *csr, const shamb200_merged_fields *f,
                      const double *d_axyz, double C_cour, double C_force, double *d_vsig, double *d_cfl_dt,
                      double *dt_min);
/* replaces the leapfrog predictor of Solver::do_predictor_leapfrog (Solver.cpp:390-524,
 * shamrock/src/math/integrators.cpp:88-119): v += dt/2 a; u += dt/2 du; x += dt v; v += dt/2 a;
 * u += dt/2 du, in place over n objects. */
int shamb200_leapfrog_predict(shamb200_ctx *ctx, uint32_t n, double dt, double *d_xyz, double *d_vxyz,
                              const double *d_axyz, double *d_uint, const double *d_duint);
/* replaces the corrector of Solver::apply_corrector / the eps_v test (Solver.cpp:2513-2616):
 * v += half_dt (a - a_old); u += half_dt (du - du_old); out2 = {max |dv|^2, sum |v|^2}.  Synchronises. */
int shamb200_leapfrog_correct(shamb200_ctx *ctx, uint32_t n, double half_dt, double *d_vxyz,
                              const double *d_axyz, const double *d_axyz_old, double *d_uint,
                              const double *d_duint, const double *d_duint_old, double out2[2]);

/* ---- patch decomposition / ghost-zone planning (host only, no CUDA call) ---------------------------
 * Pure functions of replicated metadata: every rank computes the same plan, so the NCCL send/recv
 * pairs of the ghost exchange match without negotiation.
 * replaces: the static part of PatchScheduler (shamrock/src/scheduler/PatchScheduler.cpp; patches on
 * the 2^21 integer grid, Patch.hpp:63-72) and BasicSPHGhostHandler::find_interfaces
 * (shammodels/sph/src/BasicSPHGhosts.cpp:261-509) with the exchange order of
 * shambase::DistributedDataShared (multimap on (sender, receiver), DistributedDataShared.hpp:54). */
typedef struct shamb200_iface {
    uint32_t sender, receiver; /* patch indices                                         */
    int32_t ioff[3];           /* periodic image (-1, 0, 1 per axis)                    */
    double offset[3];          /* added to the sender's positions                       */
    double cut_lo[3], cut_hi[3]; /* sender-frame box [lo, hi) whose particles are ghosts  */
} shamb200_iface;
/* boxes: [np*6] (lo[3], hi[3]) per patch, x fastest; owner: [np] rank of each patch */
int shamb200_plan_patch_grid(const double bmin[3], const double bmax[3], uint32_t nx, uint32_t ny, uint32_t nz,
                             int world_size, double *boxes, int32_t *owner);
/* interact_r[np] = max(h)*htol*Rkern per patch, pcount[np] particle counts.  Writes up to cap
 * interfaces (exchange order) and the total number found into *n_found. */
int shamb200_plan_interfaces(uint32_t npatch, const double *boxes, const double bmin[3], const double bmax[3],
                             int periodic, const double *interact_r, const uint32_t *pcount, uint32_t cap,
                             shamb200_iface *out, uint32_t *n_found);

/* Hilbert index of a cell of the 2^21-per-axis patch grid (shamrock::sfc::HilbertCurve<u64, 3>::icoord_to_hilbert,
 * shammath/include/shammath/sfc/hilbert.hpp:36-87) */
uint64_t shamb200_hilbert_index(uint64_t x, uint64_t y, uint64_t z);
/* replaces the owner table of shamrock::scheduler::HilbertLoadBalance<u64>::make_change_list
 * (shamrock/src/scheduler/HilbertLoadBalance.cpp:46-75): patches ordered by the Hilbert index of their
 * coord_min [npatch * 3], loads = particle counts (ComputeLoadBalanceValue.cpp:23-30), and
 * shamrock::scheduler::load_balance (loadbalance/LoadBalanceStrategy.hpp:274-312: parallel sweep against
 * round robin, the smaller maximum rank load wins, round robin favoured by 0.95).
 * owner [npatch]; *strategy (may be NULL): 0 parallel sweep, 1 round robin. */
int shamb200_plan_load_balance(uint32_t npatch, const uint64_t *coord_min, const uint64_t *load, int world_size,
                               int32_t *owner, int *strategy);

/* ---- model (shammodels::sph::Model<f64_3, Kernel> / Solver::evolve_once) -------------------------
 * Host-side drop-in: owns the patch data on the device and runs the whole step on the GPU.
 * Mirrors shammodels/sph/include/shammodels/sph/Model.hpp:55-1076 and Solver.cpp:1942-3272 for
 * the configuration subset of SURVEY.md §8.  All pointers here are HOST pointers. */
typedef struct shamb200_solver_config {
    int32_t kernel;  /* SHAMB200_KERNEL_*                                  */
    int32_t eos;     /* SHAMB200_EOS_*                                     */
    int32_t av;      /* SHAMB200_AV_*                                      */
    int32_t bc;      /* SHAMB200_BC_*                                      */
    double gpart_mass;
    double gamma, cs0, eos_q, eos_r0;
    double alpha_u, alpha_AV, beta_AV, alpha_min, alpha_max, sigma_decay;
    double cfl_cour, cfl_force, cfl_multiplier_stiffness;
    double htol_up_coarse_cycle, htol_up_fine_cycle, epsilon_h;
    uint32_t h_iter_per_subcycles, h_max_subcycles_count, tree_reduction_level;
    int32_t use_two_stage_search;
    int32_t combined_dtdiv_divcurlv_compute;
    int32_t sort_mode;   /* SHAMB200_SORT_*                                */
    int32_t has_point_mass;
    double pm_mass, pm_racc, constant_G;
    int32_t n_kill_spheres;
    int32_t keep_step_data; /* keep per-step intermediates for shamb200_model_get (tests) */
    int32_t fp_mode;        /* SHAMB200_FP_*                                   */
    int32_t enable_particle_reordering; /* SolverConfig.hpp:621: Morton-reorder the patch data ...     */
    double kill_center[4][3];
    double kill_radius[4];
    uint64_t particle_reordering_step_freq; /* ... at every step whose index is a multiple of this (1000) */
} shamb200_solver_config;

/* defaults of SolverConfig (shammodels/sph/include/shammodels/sph/SolverConfig.hpp:584-630,
 * config/AVConfig.hpp:46-140) */
void shamb200_solver_config_default(shamb200_solver_config *cfg);

int shamb200_model_create(shamb200_ctx *ctx, const shamb200_solver_config *cfg, shamb200_model **out);
int shamb200_model_destroy(shamb200_model *m);
int shamb200_model_set_config(shamb200_model *m, const shamb200_solver_config *cfg);
/* Model::resize_simulation_box + a static patch grid (nx*ny*nz, powers of two) on the 2^21 integer
 * patch grid of PatchScheduler; patches are dealt to ranks in id order, contiguously. */
int shamb200_model_set_box(shamb200_model *m, const double bmin[3], const double bmax[3], uint32_t nx,
                           uint32_t ny, uint32_t nz);
/* patch -> rank table (one entry per patch of the grid, patch id order), e.g. from shamb200_plan_load_balance
 * on the particle counts of the setup; allowed while no particle has been pushed (patches do not migrate
 * between ranks afterwards).  Every rank passes the same table. */
int shamb200_model_set_patch_owners(shamb200_model *m, uint32_t npatch, const int32_t *owner);
/* coord_min [npatch * 3]: the patches' lower corners on the 2^21 integer grid (Patch::coord_min,
 * shamrock/include/shamrock/patch/Patch.hpp:63-72), the input of shamb200_plan_load_balance */
int shamb200_model_patch_coords(shamb200_model *m, uint32_t npatch, uint64_t *coord_min);
/* multi-GPU: rank/size of this process and the NCCL unique id (128 bytes, from
 * shamb200_nccl_unique_id on rank 0, broadcast by the caller).  Optional (single GPU otherwise). */
int shamb200_nccl_unique_id(void *out128);
int shamb200_model_init_comm(shamb200_model *m, int rank, int world_size, const void *nccl_id128);
/* append particles (host arrays; vxyz / uint may be NULL = 0).  Each rank passes particles of any
 * patch; only those owned by a local patch are kept (setup generators call this with the same
 * deterministic stream on every rank). */
int shamb200_model_push_particles(shamb200_model *m, uint64_t n, const double *xyz, const double *vxyz,
                                  const double *hpart, const double *uint_);
/* ---- checkpoint / restart ------------------------------------------------------------------
 * Model::dump / Model::load_from_dump (shammodels/sph/include/shammodels/sph/Model.hpp:906-995) on the container of
 * shamrock::write_shamrock_dump / load_shamrock_dump (shamrock/src/io/ShamrockDump.cpp:25-274): three
 * length-prefixed JSON headers (user metadata = solver configuration, time, next dt, cfl multiplier; patch
 * metadata = patch list, simulation box, scheduler criteria, layout; table = pids / bytecounts / offsets), then one
 * blob per patch, written by the rank that owns it into ONE file.  load_dump replaces the configuration, the
 * patches and the state of `m` (create it with any configuration; call init_comm first for several ranks): owners
 * are folded onto the ranks present (owner % world), so a dump restarts on another number of GPUs.  A restarted
 * model continues bit-identically.  The blob encoding is this library's ("shamb200-1"), not the reference's
 * SerializeHelper byte stream.  Collective: every rank calls it with the same file name (shared file system). */
int shamb200_model_dump(shamb200_model *m, const char *fname);
int shamb200_model_load_dump(shamb200_model *m, const char *fname);
/* ---- Phantom dumps and legacy VTK files ---------------------------------------------------------
 * phantom_dump = Model::make_phantom_dump().save_dump(fname) (shammodels/sph/src/Model.cpp:1491-1638,
 *   io/PhantomDump.cpp:276-316): the Fortran-record container of PhantomDump::gen_file with the reference's header
 *   tables (same tags, types and order) and block 0 = fort_real x y z vx vy vz u, f32 h [alpha divv].  Sinks are
 *   outside this library: nptmass = 0, no block 1.  The reference leaves isink / polyk2 (and polyk / RK2 for the
 *   adiabatic EOS) of its EOS header uninitialised; 0 is written here.  Collective; one file, every rank writes
 *   its particles (rank by rank, patch by patch).
 * init_from_phantom_dump = Model::init_from_phantom_dump(dump, hpart_fact_load) (Model.cpp:1225-1429): box from
 *   xmin..zmax of the header, else the positions' bounding box grown by 1.2; time from the header; every particle
 *   of block 0 with h >= 0 goes to the patch that contains it (xyz vxyz hpart * hpart_fact_load, uint, alpha_AV).
 *   Every rank reads the file; *kept = particles inside the box with h >= 0.  A model without a box gets one patch.
 * phantom_gen_config = Model::gen_config_from_phantom_dump (Model.cpp:1203-1222, io/Phantom2Shamrock.cpp:27-67,
 *   :128-136, :186-197): gpart_mass = massoftype[0], C_cour, C_force, ieos 1 / 2 / 3 -> isothermal / adiabatic /
 *   LP07 (other values: error unless bypass_error), CD10 viscosity (0, 1, 0.1, alphau, 2), periodic iff xmin is in
 *   the header.  Fields of *cfg it does not name are left as they are.  No device needed.
 * phantom_copy = PhantomDump::from_file + gen_file + write_to_file (the reference's own test of its reader /
 *   writer, src/tests/phantom_read_test.cpp:20-34: the copy must be byte-identical).  phantom_header_* =
 *   PhantomDump::read_header_float / read_header_int / has_header_entry; phantom_compare = compare_phantom_dumps
 *   (number of header entries that differ, are missing or extra).
 * vtk_dump = Model::do_vtk_dump(fname, add_patch_world_id) (modules/io/VTKDump.cpp:36-178,
 *   shamrock/include/shamrock/io/LegacyVtkWriter.hpp): legacy BINARY unstructured grid, points + one FIELD section
 *   ([patchid world_rank] h u v a [alpha_AV divv] [dtdivv curlv] [soundspeed] rho), every value big-endian f32 /
 *   i32, converted on the device.  Collective. */
int shamb200_model_phantom_dump(shamb200_model *m, const char *fname);
int shamb200_model_init_from_phantom_dump(shamb200_model *m, const char *fname, double hpart_fact_load, uint64_t *kept);
int shamb200_model_vtk_dump(shamb200_model *m, const char *fname, int add_patch_world_id);
int shamb200_phantom_gen_config(const char *fname, int bypass_error, shamb200_solver_config *cfg);
int shamb200_phantom_copy(const char *fname_in, const char *fname_out);
int shamb200_phantom_header_float(const char *fname, const char *key, double *out, int *found);
int shamb200_phantom_header_int(const char *fname, const char *key, int64_t *out, int *found);
int shamb200_phantom_compare(const char *fname_a, const char *fname_b, uint64_t *offenses);
/* ---- patch scheduler ---------------------------------------------------------------------
 * PatchScheduler (shamrock/src/scheduler/PatchScheduler.cpp:308-500): patches on the 2^21 integer grid are split
 * into their eight children above crit_split objects, an octet of sibling leaves is merged below crit_merge,
 * and the patches are dealt to the ranks along the Hilbert curve of their coordinates by object count
 * (HilbertLoadBalance.cpp:46-75); a patch that changes owner moves with all its fields over NCCL send / recv.
 * The patch list is replicated, every rank calls these functions with the same arguments (collective).
 * init_scheduler = Model::init_scheduler(crit_split, crit_merge) (Model.hpp:82-98); step_freq > 0 runs
 *   scheduler_step(true, true) at the start of every step_freq-th evolve_once (Solver.cpp:1970-1976), 0 = only
 *   when called.  Ids and list order follow the reference: child 0 keeps the parent's id and place, children
 *   1..7 (child c = 4 ix + 2 iy + iz) get fresh ids at the end of the list; a merged patch keeps child 0's id.
 * patch_info: out = {id, coord_min[3], coord_max[3], owner rank}, box = [lo, hi) in simulation coordinates.
 * scheduler_log (last scheduler_step): {splits, merges, patches moved, objects moved, patch count, largest rank
 *   load, mean rank load, imbalance = max / mean - 1}. */
int shamb200_model_init_scheduler(shamb200_model *m, uint64_t crit_split, uint64_t crit_merge, uint32_t step_freq);
int shamb200_model_scheduler_step(shamb200_model *m, int do_split_merge, int do_load_balancing);
int shamb200_model_split_patch(shamb200_model *m, uint32_t ip);
int shamb200_model_merge_patches(shamb200_model *m, uint32_t ip0);
int shamb200_model_migrate_patch(shamb200_model *m, uint32_t ip, int new_owner);
int shamb200_model_patch_info(shamb200_model *m, uint32_t ip, uint64_t out[8], double box[6]);
int shamb200_model_scheduler_log(shamb200_model *m, double out[8]);
/* ---- initial conditions generated on the device ---------------------------------------------
 * The reference builds its initial conditions on the host (SPHSetup::apply_setup, shammodels/sph/src/modules/
 * SPHSetup.cpp:112-267, with GeneratorLatticeHCP / GeneratorMCDisc) and edits fields with host loops
 * (Model::set_value_in_a_box / set_value_in_sphere / add_kernel_value / get_sum, Model.hpp:669-785).  Here every
 * rank generates the objects of ITS patches in device memory; the counts returned are global.
 * add_lattice_hcp: shammath::LatticeHCP (crystalLattice.hpp:52-290) points r with box_min <= r < box_max, hpart =
 *   dr, every other field 0, appended to the owning patch in the reference's iteration order (x index fastest,
 *   :244-248) - bit-identical to the host generator.
 * add_disc_mc: Monte-Carlo disc (GeneratorMCDisc.cpp): Sigma ~ r^-p between r_in and r_out, H / r = H_r_in
 *   (r / r_in)^(1/2 - q), Keplerian velocities around the configured point mass, h from the local density; object
 *   i draws from its own counter-based stream (seed, i), so the disc does not depend on the patch / rank layout.
 *   Needs the particle mass (shamb200_model_set_particle_mass or the config). */
int shamb200_model_add_lattice_hcp(shamb200_model *m, double dr, const double box_min[3], const double box_max[3],
                                   uint64_t *added);
/* add_disc_lattice: the HCP lattice cut to a flared disc (r_in < R < r_out, |z| < zcut R), Keplerian velocities,
 * h = hfact (4 sqrt 2)^(1/3) dr: a regular stand-in for the Monte-Carlo disc (whose Poisson clumps leave objects
 * with a non-converging h iteration, in the reference as well) for large benchmark runs. */
int shamb200_model_add_disc_lattice(shamb200_model *m, double dr, double r_in, double r_out, double zcut,
                                    uint64_t *added);
int shamb200_model_add_disc_mc(shamb200_model *m, uint64_t npart, uint64_t seed, double r_in, double r_out, double p,
                               double q, double H_r_in, double disc_mass, uint64_t *added);
int shamb200_model_set_value_in_a_box(shamb200_model *m, const char *field, int ivar, double val,
                                      const double box_min[3], const double box_max[3]);
int shamb200_model_set_value_in_sphere(shamb200_model *m, const char *field, double val, const double center[3],
                                       double radius);
int shamb200_model_add_kernel_value(shamb200_model *m, const char *field, double val, const double center[3],
                                    double h_ker);
int shamb200_model_get_sum(shamb200_model *m, const char *field, double out[3]); /* all ranks */
int shamb200_model_total_part_count(shamb200_model *m, uint64_t *out);           /* Model::get_total_part_count */
int shamb200_model_set_particle_mass(shamb200_model *m, double gpart_mass);      /* Model::set_particle_mass */
uint32_t shamb200_model_patch_count(shamb200_model *m);       /* global number of patches   */
int shamb200_model_patch_is_local(shamb200_model *m, uint32_t ip);
uint32_t shamb200_model_patch_size(shamb200_model *m, uint32_t ip); /* 0 for remote patches  */
/* field access by name, patch by patch (ctx.collect_data() equivalent).  Names: main layout
 * (SolverConfig.cpp:24-121) xyz vxyz axyz axyz_ext hpart uint duint alpha_AV divv dtdivv curlv
 * soundspeed; with keep_step_data also step.* / tree.* / cache.* (see DESIGN.md).
 * get: returns the byte size, copies when cap_bytes is large enough; -1 if unknown. */
int64_t shamb200_model_get(shamb200_model *m, uint32_t ip, const char *name, void *out, int64_t cap_bytes);
int shamb200_model_set_field(shamb200_model *m, uint32_t ip, const char *name, const double *in, uint64_t count);
/* modules::ParticleReordering::reorder_particles (shammodels/sph/src/modules/ParticleReordering.cpp:22-51),
 * the call SPHSetup::apply_setup(part_reordering = true) makes after the particles are in place
 * (SPHSetup.cpp:202-205): every local patch is permuted into the Morton order of its positions over the
 * patch box (RadixTreeMortonBuilder.cpp:68-107; sort_mode BITONIC reproduces the reference's order inside
 * runs of equal codes).  evolve_once does the same at the steps selected by enable_particle_reordering /
 * particle_reordering_step_freq (Solver.cpp:2043-2048). */
int shamb200_model_reorder_particles(shamb200_model *m);
/* Solver::evolve_once (Solver.cpp:1942).  Runs one full step; synchronises at the end. */
int shamb200_model_evolve_once(shamb200_model *m);
/* Solver::evolve_once on HOST-resident patch data: the call a host code that keeps its PatchDataLayer
 * fields in its own memory makes once per step (the reference's fields live in sham::DeviceBuffer /
 * PatchDataField, shamrock/include/shamrock/patch/PatchDataField.hpp; main layout SolverConfig.cpp:24-121).
 * `in`: the step's inputs (xyz vxyz axyz hpart uint duint; alpha_AV and soundspeed for MM97/CD10 — the AV
 * switch reads the previous step's soundspeed, UpdateViscosity.cpp:52-222); the other main-layout fields are
 * outputs of the step and are never read (NULL allowed everywhere in `in` = keep the device copy; fields of a
 * patch the library has just grown start at 0, like the reference's PatchDataField).  `out`: every non-NULL pointer receives the field after the step; out->n is the capacity
 * (objects) of the out arrays on entry and the object count of the patch on return.
 * Copies run on two copy streams and overlap with the kernels: the positions are drifted and the tree /
 * neighbour cache built while uint, duint, alpha_AV are still uploading; xyz, hpart, axyz_ext go back
 * during the CD10 operators, divv curlv dtdivv alpha_AV during the force loop; only vxyz uint axyz duint
 * soundspeed follow the corrector.  Host memory should be page-locked (shamb200_host_register) for the
 * copies to be asynchronous.  One local patch per call (`ip`); a model with several local patches,
 * kill spheres, a point mass or free boundaries takes the same call without the overlap.
 * With enable_particle_reordering the objects of a patch change places at the reordering steps: the fields that
 * come back are in the new order, so a host that keeps its own copy reads back every field it keeps. */
typedef struct shamb200_host_patchdata {
    uint64_t n;
    double *xyz, *vxyz, *axyz, *axyz_ext; /* 3 doubles per object */
    double *hpart, *uint_, *duint, *alpha_AV, *divv, *dtdivv;
    double *curlv; /* 3 doubles per object */
    double *soundspeed;
} shamb200_host_patchdata;
int shamb200_model_evolve_once_host(shamb200_model *m, uint32_t ip, const shamb200_host_patchdata *in,
                                    shamb200_host_patchdata *out);
/* page-lock / unlock caller memory (cudaHostRegister) so that the copies above are asynchronous */
int shamb200_host_register(void *p, uint64_t bytes);
int shamb200_host_unregister(void *p);
/* neighbour search of the last step on this rank: out[0] = sum of the list lengths (ObjectCache::sum_neigh_cnt,
 * TreeTraversal.hpp:378), out[1] = (particle, candidate) accept tests done to build them */
int shamb200_model_search_stats(shamb200_model *m, uint64_t out[2]);
/* bytes moved by the last shamb200_model_evolve_once_host: out[0] host->device, out[1] device->host */
int shamb200_model_host_traffic(shamb200_model *m, uint64_t out[2]);
/* Neighbour-list tolerance of the model's own step (not of shamb200_neigh_cache_build, which takes the caller's).
 * The reference builds the lists of a step with the radius R h htol, htol = htol_up_coarse_cycle = 1.1
 * (Solver.cpp:1322-1386, NeighbourCache.cpp:482-520), so that h may grow inside the step.  With fp_mode FAST the
 * step builds them with a tolerance fitted to the h growth of the previous step, verifies after the h iteration
 * that no h_iterate / h_old exceeded it (then the loops saw every pair inside a kernel support: same sums as with
 * the full lists) and otherwise redoes the sub-cycle with htol.  STRICT mode, keep_step_data and epsilon_h != 1e-6
 * always use htol.  out = {tolerance of the last step's lists, largest h_iterate / h_old of the last step (all
 * ranks), tolerance the next step will start with, number of steps that had to fall back so far}.
 * Environment: SHAMB200_LIST_TOL=0 always htol; a value in (1, htol] fixes the starting tolerance. */
int shamb200_model_list_tolerance(shamb200_model *m, double out[4]);
/* how the last shamb200_model_evolve_once_host ran: out[0] = number of id ranges its operator / force / corrector
 * passes were cut into so that finished ranges travel to the host while the next is computed (0: one launch per
 * pass), out[1] = objects of the patch whose successor by id lay farther away than 8 h — above 1 % of the patch the
 * ranges are not used: consecutive ids must be neighbours in space (shamb200_model_reorder_particles,
 * ParticleReordering.hpp:38-120, makes them so) */
int shamb200_model_host_step_info(shamb200_model *m, uint32_t ip, uint64_t out[2]);
/* state: {time, next dt, cfl_multiplier, eps_v, h_subcycles, h_iters_last, corrector_iter,
 *         npart(global), t_step seconds (host wall), rate(part/s, this rank), K (local neighbour count)} */
int shamb200_model_state(shamb200_model *m, double out[12]);
/* modules::ConservativeCheck::check_conservation (shammodels/sph/src/modules/ConservativeCheck.cpp:26-190), the
 * sums the reference logs at every step right before the corrector (Solver.cpp:2503-2504), over all ranks:
 * out = {m sum v (x, y, z), m sum a (x, y, z), m sum (u + v.v / 2), m sum (v.a + du/dt)}.  They are reduced inside the
 * corrector kernel of the last evolve_once (no extra pass over the fields). */
int shamb200_model_conservation(shamb200_model *m, double out[8]);
int shamb200_model_set_next_dt(shamb200_model *m, double dt);
int shamb200_model_set_time(shamb200_model *m, double t);
int shamb200_model_set_cfl_multiplier(shamb200_model *m, double v);
/* per-stage device time of the last step (CUDA events), names separated by ';' in *names */
int shamb200_model_stage_times(shamb200_model *m, const char **names, const double **ms, uint32_t *count);

#ifdef __cplusplus
}
#endif
#endif /* SHAMB200_H */
