// solver_comm.cu — multi-GPU plumbing: NCCL (loaded at run time from the process, i.e. the copy
// torch.distributed already uses) for the ghost-zone exchange (C1-C3 of SURVEY.md §2.4) and the
// step scalars (C5, C7, C8).  One process per GPU.
#include "solver.cuh"
#include <dlfcn.h>
#include <cstring>

namespace sb {

// minimal NCCL ABI (nccl.h 2.x): only what this library calls
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess_ = 0 };
enum { ncclInt8_ = 0, ncclChar_ = 0, ncclUint8_ = 1, ncclInt32_ = 2, ncclUint32_ = 3, ncclInt64_ = 4, ncclUint64_ = 5, ncclFloat64_ = 8 };
enum { ncclSum_ = 0, ncclProd_ = 1, ncclMax_ = 2, ncclMin_ = 3 };

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId *)                                                         = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int)                                  = nullptr;
    int (*CommDestroy)(ncclComm_t)                                                             = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t)         = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t)              = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t)                      = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t)                            = nullptr;
    int (*GroupStart)()                                                                        = nullptr;
    int (*GroupEnd)()                                                                          = nullptr;
    const char *(*GetErrorString)(int)                                                         = nullptr;
};

static NcclApi &nccl() {
    static NcclApi api;
    if (api.handle)
        return api;
    // prefer the NCCL already mapped in the process (torch's), then the system one
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h             = nullptr;
    for (auto n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (h)
            break;
    }
    if (!h)
        for (auto n : names) {
            h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (h)
                break;
        }
    if (!h)
        throw std::runtime_error(std::string("cannot load NCCL: ") + dlerror());
    api.handle = h;
    auto sym   = [&](const char *n) {
        void *p = dlsym(h, n);
        if (!p)
            throw std::runtime_error(std::string("NCCL symbol missing: ") + n);
        return p;
    };
    api.GetUniqueId    = (decltype(api.GetUniqueId)) sym("ncclGetUniqueId");
    api.CommInitRank   = (decltype(api.CommInitRank)) sym("ncclCommInitRank");
    api.CommDestroy    = (decltype(api.CommDestroy)) sym("ncclCommDestroy");
    api.AllReduce      = (decltype(api.AllReduce)) sym("ncclAllReduce");
    api.AllGather      = (decltype(api.AllGather)) sym("ncclAllGather");
    api.Send           = (decltype(api.Send)) sym("ncclSend");
    api.Recv           = (decltype(api.Recv)) sym("ncclRecv");
    api.GroupStart     = (decltype(api.GroupStart)) sym("ncclGroupStart");
    api.GroupEnd       = (decltype(api.GroupEnd)) sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString)) sym("ncclGetErrorString");
    return api;
}

struct NcclError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
#define SB_NCCL_CHECK(expr)                                                                      \
    do {                                                                                         \
        int _r = (expr);                                                                         \
        if (_r != 0)                                                                             \
            throw NcclError(std::string(#expr) + " failed: " + nccl().GetErrorString(_r));       \
    } while (0)

void comm_unique_id(void *out128) {
    ncclUniqueId id;
    SB_NCCL_CHECK(nccl().GetUniqueId(&id));
    std::memcpy(out128, &id, 128);
}

void comm_init(Model &m, int rank, int world, const void *id128) {
    if (world < 1 || rank < 0 || rank >= world)
        throw std::invalid_argument("invalid rank / world size");
    // set_box deals the patches to the ranks known at that time: a communicator that arrives later re-deals
    // them (same contiguous id blocks as plan_patch_grid) — allowed only while no patch holds objects
    if (!m.patches.empty() && (world != m.world || rank != m.rank)) {
        for (auto &p : m.patches)
            if (p.f.n)
                throw std::invalid_argument("init_comm: the patches already hold objects (call it before push_particles)");
        const size_t np = m.patches.size();
        for (size_t k = 0; k < np; k++)
            m.patches[k].owner = int((uint64_t(k) * uint64_t(world)) / np);
    }
    m.rank  = rank;
    m.world = world;
    if (world == 1)
        return;
    SB_CUDA_CHECK(cudaSetDevice(m.ctx->device));
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    ncclComm_t c = nullptr;
    SB_NCCL_CHECK(nccl().CommInitRank(&c, world, id, rank));
    m.nccl_comm = c;
    // point-to-point exchanges run on their own stream, tied to the compute stream by two events: the work that does
    // not need the arriving data is queued between comm_group_end and comm_wait and overlaps the transfer
    if (!m.comm_s) {
        SB_CUDA_CHECK(cudaStreamCreateWithFlags(&m.comm_s, cudaStreamNonBlocking));
        SB_CUDA_CHECK(cudaEventCreateWithFlags(&m.ev_comm_begin, cudaEventDisableTiming));
        SB_CUDA_CHECK(cudaEventCreateWithFlags(&m.ev_comm_done, cudaEventDisableTiming));
    }
}

void comm_allreduce_f64(Model &m, f64 *d_buf, size_t n, int op) {
    if (m.world == 1)
        return;
    int nop = op == 0 ? ncclSum_ : (op == 1 ? ncclMax_ : ncclMin_);
    SB_NCCL_CHECK(nccl().AllReduce(d_buf, d_buf, n, ncclFloat64_, nop, (ncclComm_t) m.nccl_comm, m.s()));
}
void comm_allreduce_u64(Model &m, u64 *d_buf, size_t n, int op) {
    if (m.world == 1)
        return;
    int nop = op == 0 ? ncclSum_ : (op == 1 ? ncclMax_ : ncclMin_);
    SB_NCCL_CHECK(nccl().AllReduce(d_buf, d_buf, n, ncclUint64_, nop, (ncclComm_t) m.nccl_comm, m.s()));
}
void comm_allreduce_host_u64(Model &m, u64 *vals, size_t n, int op) {
    if (m.world == 1 || n == 0)
        return;
    m.comm_dev.ensure(n);
    m.comm_host.ensure(n);
    std::memcpy(m.comm_host.p, vals, n * sizeof(u64));
    h2d_small(m.s(), m.comm_dev.p, m.comm_host.p, n * sizeof(u64));
    comm_allreduce_u64(m, m.comm_dev.p, n, op);
    d2h_small(m.s(), m.comm_host.p, m.comm_dev.p, n * sizeof(u64));
    SB_CUDA_CHECK(cudaStreamSynchronize(m.s()));
    std::memcpy(vals, m.comm_host.p, n * sizeof(u64));
}
void comm_allreduce_host_f64(Model &m, f64 *vals, size_t n, int op) {
    if (m.world == 1 || n == 0)
        return;
    m.comm_dev.ensure(n);
    m.comm_host.ensure(n);
    std::memcpy(m.comm_host.p, vals, n * sizeof(f64));
    h2d_small(m.s(), m.comm_dev.p, m.comm_host.p, n * sizeof(f64));
    comm_allreduce_f64(m, reinterpret_cast<f64 *>(m.comm_dev.p), n, op);
    d2h_small(m.s(), m.comm_host.p, m.comm_dev.p, n * sizeof(f64));
    SB_CUDA_CHECK(cudaStreamSynchronize(m.s()));
    std::memcpy(vals, m.comm_host.p, n * sizeof(f64));
}
/// what the compute stream has queued so far (the staging of the outgoing messages) precedes the group
void comm_group_start(Model &m) {
    if (m.world > 1) {
        SB_CUDA_CHECK(cudaEventRecord(m.ev_comm_begin, m.s()));
        SB_CUDA_CHECK(cudaStreamWaitEvent(m.comm_s, m.ev_comm_begin, 0));
        SB_NCCL_CHECK(nccl().GroupStart());
    }
}
/// the group is in flight on the communication stream; the compute stream goes on until comm_wait
void comm_group_end(Model &m) {
    if (m.world > 1) {
        SB_NCCL_CHECK(nccl().GroupEnd());
        SB_CUDA_CHECK(cudaEventRecord(m.ev_comm_done, m.comm_s));
    }
}
/// the compute stream waits for the last group (arrived data may be read, sent buffers reused)
void comm_wait(Model &m) {
    if (m.world > 1)
        SB_CUDA_CHECK(cudaStreamWaitEvent(m.s(), m.ev_comm_done, 0));
}
void comm_send(Model &m, const void *d, size_t bytes, int peer) {
    if (!bytes)
        return;
    SB_NCCL_CHECK(nccl().Send(d, bytes, ncclUint8_, peer, (ncclComm_t) m.nccl_comm, m.comm_s));
}
void comm_recv(Model &m, void *d, size_t bytes, int peer) {
    if (!bytes)
        return;
    SB_NCCL_CHECK(nccl().Recv(d, bytes, ncclUint8_, peer, (ncclComm_t) m.nccl_comm, m.comm_s));
}
void comm_destroy(Model &m) {
    if (m.nccl_comm) {
        nccl().CommDestroy((ncclComm_t) m.nccl_comm);
        m.nccl_comm = nullptr;
    }
    if (m.comm_s) {
        cudaStreamSynchronize(m.comm_s);
        cudaEventDestroy(m.ev_comm_begin);
        cudaEventDestroy(m.ev_comm_done);
        cudaStreamDestroy(m.comm_s);
        m.comm_s = nullptr;
    }
}

} // namespace sb

