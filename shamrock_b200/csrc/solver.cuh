// solver.cuh — host side of the B200 SPH step: device-resident patch data + Solver::evolve_once.
// Mirrors shammodels::sph::Solver / Model (shammodels/sph/include/shammodels/sph/{Solver,Model}.hpp)
// for the configuration subset of SURVEY.md §8.
#pragma once
#include "../../include/shamb200.h"
#include "ghost_plan.hpp"
#include "neigh.cuh"
#include "neigh2.cuh"
#include "sph.cuh"
#include "sph2.cuh"
#include "stream_kernels.cuh"
#include "tree.cuh"
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace sb {

struct Ctx {
    int device          = 0;
    cudaStream_t stream = nullptr;
    bool own_stream     = false;
    // arenas of the stage-level C ABI (shamb200_tree_build / shamb200_neigh_cache_build)
    TreeBuffers api_tree;
    NeighBuffers api_nb;
    SearchBuffers api_srch;
    DevBuf<u64> red;
    PinnedBuf<u64> h_red;
    DevBuf<Pack4> api_A, api_B, api_C, api_D; ///< merged records of the stage-level SPH modules
    DevBuf<f64> api_tmp;
};

/// main patch data layout (SolverConfig.cpp:24-121), device resident, packed vec3 (3 doubles)
struct PatchFields {
    u32 n = 0;
    DevBuf<f64> xyz, vxyz, axyz, axyz_ext, curlv;                         // 3n
    DevBuf<f64> hpart, uint_, duint, alpha_AV, divv, dtdivv, soundspeed; // n
    struct Ref {
        const char *name;
        DevBuf<f64> *buf;
        int nvar;
    };
    std::vector<Ref> all();
    void reserve(u32 cap, cudaStream_t s); ///< grow keeping the first n objects
};

struct Iface {
    u32 sender, receiver; ///< indices in Model::patches
    f64 offset[3];
    i32 ioff[3];
    f64 cut_lo[3], cut_hi[3];
    u32 count   = 0;
    u32 dst_off = 0; ///< offset inside the receiver's ghost range
    size_t stage_off = 0; ///< offset (objects) in the send staging when the receiver is on another rank
    u32 *ids = nullptr;   ///< ascending sender-local ids (inside the sender patch's ids pool; local senders only)
};

struct PatchStep {
    u32 n = 0, m = 0;
    DevBuf<Pack4> A;          ///< merged (x,y,z,h) by merged id (real first, then ghosts): tree-build input
    DevBuf<Pack4> SB, SC, SD; ///< Morton-sorted records: (v,u), (P,omega,cs,alpha), (a,0); (x,y,z,h) is srch.SA
    DevBuf<Pack4> SE, SF;     ///< fast fp mode: per-particle derived factors (sph2_fast.cu)
    DevBuf<double2> SG;       ///< fast fp mode, adiabatic EOS: (1/(rho² Ω), α c_s), the 16-byte neighbour record
    TreeBuffers tree;
    DevBuf<f64> rint;
    SearchBuffers srch;
    DevBuf<f64> omega, alpha_updated, vsig, cfl_dt, eps, h_old, a_old, du_old;
    DevBuf<f64> mh_snapshot; ///< pre-iteration merged h (keep_step_data only)
    // ghost selection scratch of this patch as a sender (stream_kernels.cu: ghost_select_*)
    DevBuf<u64> gmask, gbase;
    DevBuf<u32> gblock, gtotals, ids_pool;
};

struct PatchD : PatchBox {
    PatchFields f;
    PatchStep st;
};

struct StageTimer {
    std::vector<cudaEvent_t> pool;
    std::vector<std::pair<std::string, int>> marks; // (name, event index)
    std::map<std::string, double> acc;
    std::vector<std::string> order;
    std::string names_joined;
    std::vector<double> values;
    void begin_step();
    void mark(cudaStream_t s, const char *name);
    void end_step(cudaStream_t s);
    ~StageTimer();
};

struct Model {
    Ctx *ctx;
    shamb200_solver_config cfg;
    f64 box_min[3] = {0, 0, 0}, box_max[3] = {1, 1, 1};
    std::vector<PatchD> patches;
    int rank = 0, world = 1;
    void *nccl_comm = nullptr;
    cudaStream_t comm_s = nullptr; ///< stream of the ncclSend / ncclRecv groups (solver_comm.cu)
    cudaEvent_t ev_comm_begin = nullptr, ev_comm_done = nullptr;
    f64 time = 0, dt = 0, cfl_multiplier = 1e-2; // Solver.hpp:131-147
    // log of the last step
    f64 eps_v = 0, t_step = 0;
    u32 h_subcycles = 0, h_iters_last = 0, corrector_iter = 0;
    u64 npart_all = 0, K_local = 0, pair_tests_local = 0;
    u64 step_count = 0; ///< SolverLog::step_count: steps done so far
    TreeBuffers reorder_tree; ///< Morton codes / permutation of reorder_particles
    std::vector<Iface> ifaces;
    StageTimer timer;
    // scratch
    DevBuf<u8> flag;
    DevBuf<u32> pos, scan_tmp, owner_tmp;
    DevBuf<u64> red;
    PinnedBuf<u64> h_red;
    DevBuf<f64> field_tmp, d_boxes;
    DevBuf<u32> box_counts;
    std::vector<f64> host_tmp;
    // multi-GPU staging (ghost / migration sends) and host<->device scratch of the scalar allreduces
    DevBuf<Pack4> send_stage, recv_stage;
    size_t send_total = 0; ///< ghosts this rank sends to other ranks (objects)
    DevBuf<f64> send_stage_f, recv_stage_f;
    DevBuf<u64> comm_dev;
    PinnedBuf<u64> comm_host;
    PinnedBuf<u32> h_counts; ///< small count read-backs (d2h_small needs page-locked memory)
    DevBuf<f64> step_sc;       ///< the step's cross-rank scalars (stream_kernels.cu: step_scalars)
    PinnedBuf<f64> h_step_sc;
    /// modules::ConservativeCheck of the last step: m Σ v (3), m Σ a (3), m Σ (u + v²/2), m Σ (v·a + du/dt)
    f64 conservation[8] = {0, 0, 0, 0, 0, 0, 0, 0};

    /// host-resident patch data of the running step (evolve_once_host): copy streams + hand-over events
    struct HostPipe {
        bool active = false, early_out = false, defer_in2 = false;
        bool sliced_out = false; ///< this step ran its loops over id ranges (Model::pipe_slices)
        u32 nslices     = 0;
        DevBuf<u64> far_dev; ///< objects whose successor by id is far away (count_far_successors) ...
        PinnedBuf<u64> far_host; ///< ... read back with the synchronisation of the neighbour search
        u32 ip = 0;
        const shamb200_host_patchdata *in = nullptr;
        shamb200_host_patchdata *out      = nullptr;
        cudaStream_t h2d = nullptr, d2h = nullptr;
        cudaEvent_t ev_in1 = nullptr, ev_in2 = nullptr, ev_stage = nullptr;
        u64 bytes_h2d = 0, bytes_d2h = 0, out_cap = 0;
    } pipe;
    void pipe_download(const char *const *names, int count, u32 first = 0, u32 n = 0xffffffffu);
    std::vector<std::pair<u32, u32>> pipe_slices();

    explicit Model(Ctx *c, const shamb200_solver_config &cf) : ctx(c), cfg(cf) {}
    /// SHAMB200_VERBOSE=1|2: progress of the prestep on stderr (sub-cycles; 2: every patch)
    static int verbose() {
        static int v = [] {
            const char *e = getenv("SHAMB200_VERBOSE");
            return e ? atoi(e) : 0;
        }();
        return v;
    }
    cudaStream_t s() const { return ctx->stream; }
    bool is_local(const PatchD &p) const { return p.owner == rank; }

    void set_box(const f64 bmin[3], const f64 bmax[3], u32 nx, u32 ny, u32 nz);
    void push_particles(u64 n, const f64 *xyz, const f64 *vxyz, const f64 *h, const f64 *u, const f64 *alpha = nullptr);
    // patch scheduler (scheduler.cu): PatchScheduler::scheduler_step and its pieces
    u64 next_patch_id = 0;            ///< SchedulerPatchList::_next_patch_id
    u64 crit_split = 0, crit_merge = 0; ///< Model::init_scheduler(crit_split, crit_merge); 0 = never
    u32 scheduler_freq = 0;           ///< run scheduler_step inside evolve_once every this many steps (0 = never)
    std::vector<u64> patch_counts;    ///< replicated object count per patch (refresh_counts)
    struct SchedLog {
        u32 splits = 0, merges = 0, moves = 0, npatch = 0;
        u64 moved_objects = 0, max_rank_load = 0;
        f64 mean_rank_load = 0;
    } sched_log;
    void refresh_boxes();
    void refresh_counts();
    std::vector<u32> sibling_octet(u32 ip0) const;
    void split_patch(u32 ip);
    void merge_patches(u32 ip0);
    void migrate_patch(u32 ip, int new_owner);
    void scheduler_step(bool do_split_merge, bool do_load_balancing);
    // checkpoint / restart (dump.cu): write_shamrock_dump / load_shamrock_dump container
    void dump(const std::string &fname);
    void load_dump(const std::string &fname);
    // Phantom dumps and legacy VTK files (io_formats.cu)
    void phantom_dump(const std::string &fname);                             ///< make_phantom_dump + save_dump
    u64 init_from_phantom_dump(const std::string &fname, f64 hpart_fact_load); ///< returns the particles kept (all ranks' input)
    void vtk_dump(const std::string &fname, bool add_patch_world_id);        ///< Model::do_vtk_dump
    void evolve_once();
    void evolve_once_host(u32 ip, const shamb200_host_patchdata *in, shamb200_host_patchdata *out);
    int64_t get(u32 ip, const std::string &name, void *out, int64_t cap);
    void set_field(u32 ip, const std::string &name, const f64 *in, u64 count);
    // initial conditions generated on the device (setup.cu); counts are global (all ranks)
    u64 add_lattice_hcp(f64 dr, const f64 bmin[3], const f64 bmax[3]);
    u64 add_disc_lattice(f64 dr, f64 r_in, f64 r_out, f64 zcut);
    u64 add_lattice_impl(f64 dr, const f64 bmin[3], const f64 bmax[3], const f64 *disc);
    u64 add_disc_mc(u64 npart, u64 seed, f64 r_in, f64 r_out, f64 p_exp, f64 q_exp, f64 H_r_in, f64 disc_mass);
    void set_value_in_a_box(const std::string &name, int ivar, f64 val, const f64 bmin[3], const f64 bmax[3]);
    void set_value_in_sphere(const std::string &name, f64 val, const f64 center[3], f64 radius);
    void add_kernel_value(const std::string &name, f64 val, const f64 center[3], f64 h_ker);
    void get_sum(const std::string &name, f64 out[3]);
    u64 total_part_count();

    // pieces of the step (names follow the reference's Solver methods)
    void keep_flagged(PatchD &p, u32 *out_kept = nullptr, const u8 *flg = nullptr);
    void remove_in_sphere(const f64 center[3], f64 radius, int mode);
    void point_mass_accrete_particles();
    void kill_particles();
    void compute_ext_forces_indep_v();
    void apply_position_boundary();
    void reorder_particles();
    void reattribute_patch_objects();
    void build_ghost_cache();
    void merge_position_ghost();
    void build_merged_pos_trees(f64 tol);
    void compute_presteps_rint(f64 tol);
    void start_neighbors_cache(f64 tol);
    void sph_prestep();
    void communicate_merge_ghosts_fields();
    void exchange_alpha_ghosts(bool with_omega);
    /// fast fp mode with a varying-alpha switch: the Ω sum rides in the CD10 / MM97 operator pass instead of
    /// costing a pass of its own after the h iteration (sph2_fast.cu: av_operators_fast_kernel<..., OMEGA>)
    bool omega_in_av_pass() const {
        return cfg.fp_mode == SHAMB200_FP_FAST && (cfg.av == SHAMB200_AV_MM97 || cfg.av == SHAMB200_AV_CD10);
    }
    // ---- list tolerance of the fast fp mode (solver.cu: Model::sph_prestep) ----
    f64 list_tol_next  = 0;  ///< tolerance the next step's lists are built with (0: htol_up_coarse_cycle)
    f64 list_tol_last  = 0;  ///< what the last step's lists were built with
    f64 h_growth_last  = 1;  ///< max over all ranks, particles and sweeps of h_iterate / h_old in the last step
    u64 list_fallbacks = 0;  ///< steps whose h outgrew the tight lists and were redone with the reference's tolerance
    /// SHAMB200_LIST_TOL: 0 = always the reference's tolerance; a value in (1, htol]: that tolerance in every step
    /// (still with the fallback); unset: adaptive
    static f64 list_tol_override() {
        const char *e = getenv("SHAMB200_LIST_TOL");
        return e ? atof(e) : -1.;
    }
    /// SHAMB200_FUSED_RINT=0: the interaction radii in their own tree pass (A / B runs)
    static bool no_fused_rint() {
        const char *e = getenv("SHAMB200_FUSED_RINT");
        return e && atoi(e) == 0;
    }
    bool wrapped_in_drift = false; ///< this step's predictor already applied the periodic wrap
    void reset_red();
    void read_red(int n);
};

f64 microbench(Ctx &c, int what); ///< microbench.cu

// Phantom dump files (io_formats.cu; host only)
void phantom_gen_config(const std::string &fname, bool bypass_error, shamb200_solver_config &cfg);
void phantom_copy(const std::string &in, const std::string &out); ///< from_file + gen_file + write_to_file
int phantom_header(const std::string &fname, const std::string &key, int want_int, f64 *fval, i64 *ival);
u64 phantom_compare(const std::string &fa, const std::string &fb); ///< compare_phantom_dumps: number of offenses

// NCCL plumbing (solver_comm.cu)
void comm_unique_id(void *out128);
void comm_init(Model &m, int rank, int world, const void *id128);
void comm_allreduce_f64(Model &m, f64 *d_buf, size_t n, int op); ///< op: 0 sum, 1 max, 2 min
void comm_allreduce_u64(Model &m, u64 *d_buf, size_t n, int op);
/// blocking allreduce of a few HOST scalars, in place (no-op when world == 1)
void comm_allreduce_host_f64(Model &m, f64 *vals, size_t n, int op);
void comm_allreduce_host_u64(Model &m, u64 *vals, size_t n, int op);
void comm_group_start(Model &m);
void comm_group_end(Model &m);
void comm_wait(Model &m);
void comm_send(Model &m, const void *d, size_t bytes, int peer);
void comm_recv(Model &m, void *d, size_t bytes, int peer);
void comm_destroy(Model &m);

} // namespace sb

namespace sb {
/// message returned by shamb200_last_error (thread local; capi.cu)
void set_last_error(const char *msg);
/// throws std::invalid_argument unless `c` is a live context handle (capi.cu)
void require_live(struct ::shamb200_ctx *c);
} // namespace sb

struct shamb200_ctx {
    sb::Ctx c;
};
struct shamb200_model {
    sb::Model m;
    shamb200_model(sb::Ctx *c, const shamb200_solver_config &cf) : m(c, cf) {}
};
