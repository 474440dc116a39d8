// io_formats.cu — the on-disk formats either side of the step: Phantom dumps and legacy VTK.
//
// Phantom dump (shammodels/sph/src/io/PhantomDump.cpp:34-340, shammodels/sph/include/shammodels/sph/io/
// PhantomDump.hpp, shambase/include/shambase/fortran_io.hpp): a Fortran unformatted file — every record is
// [i32 byte count][payload][i32 byte count] —
//   (i1, r1, i2, iversion, i3)  magic numbers 60769 / 60878 / 690706, r1 = i2
//   fileid                      100 characters
//   8 header tables             fort_int, i8, i16, i32, i64, fort_real, f32, f64: [nvars] and, when nvars > 0,
//                               [nvars tags of 16 characters][nvars values]
//   nblocks, then per block     [i64 tot_count, i32 counts[8]]  (arrays of each of the 8 types)
//   per block, per type, per array  [tag of 16 characters][tot_count values]
// Model::make_phantom_dump (Model.cpp:1491-1638), add_pdat_to_phantom_block (:1432-1488), the EOS / boundary / unit
// header entries (io/PhantomDumpEOSUtils.cpp:43-60,170-247, io/Phantom2Shamrock.cpp:147-233), and the way back:
// gen_config_from_phantom_dump (Model.cpp:1203-1222) and init_from_phantom_dump (:1225-1429).
//
// Legacy VTK (shamrock/include/shamrock/io/LegacyVtkWriter.hpp:160-420, shammodels/common/include/shammodels/
// common/io/VTKDumpUtils.hpp:42-160, shammodels/sph/src/modules/io/VTKDump.cpp:36-178): "BINARY" unstructured grid
// with points only, every value converted to big-endian f32 (integers: big-endian i32), the fields of the main
// layout as one FIELD section of the point data plus rho = m (hfact / h)^3.
//
// Several ranks: the particle order of both files is rank by rank, inside a rank patch by patch in list order
// (what the reference's MPI file views / allgather give); every rank writes its own slice of ONE file.
#include "solver.cuh"
#include "sphkern.cuh"
#include <array>
#include <cmath>
#include <cstring>
#include <fcntl.h>
#include <functional>
#include <limits>
#include <map>
#include <sys/stat.h>
#include <unistd.h>

namespace sb {

namespace {

void pwrite_all(int fd, const void *buf, size_t n, u64 off) {
    const char *p = static_cast<const char *>(buf);
    while (n) {
        ssize_t w = ::pwrite(fd, p, n, off_t(off));
        if (w <= 0)
            throw std::runtime_error("write failed");
        p += w, n -= size_t(w), off += u64(w);
    }
}

std::string pad16(const std::string &s) { // sham::format("{:16s}", s): left aligned, space padded, never cut
    std::string o = s;
    if (o.size() < 16)
        o.resize(16, ' ');
    return o;
}

// ---------------------------------------------------------------------------------------------
// Phantom container
// ---------------------------------------------------------------------------------------------
/// the 8 element types of a Phantom dump, in file order
enum PhType { PH_INT = 0, PH_I8, PH_I16, PH_I32, PH_I64, PH_REAL, PH_F32, PH_F64 };
constexpr size_t kPhSize[8] = {4, 1, 2, 4, 8, 8, 4, 8};

f64 ph_as_f64(int t, const unsigned char *p) {
    switch (t) {
    case PH_INT:
    case PH_I32: { i32 v; std::memcpy(&v, p, 4); return f64(v); }
    case PH_I8: { signed char v; std::memcpy(&v, p, 1); return f64(v); }
    case PH_I16: { short v; std::memcpy(&v, p, 2); return f64(v); }
    case PH_I64: { i64 v; std::memcpy(&v, p, 8); return f64(v); }
    case PH_F32: { float v; std::memcpy(&v, p, 4); return f64(v); }
    default: { f64 v; std::memcpy(&v, p, 8); return v; }
    }
}
i64 ph_as_i64(int t, const unsigned char *p) {
    switch (t) {
    case PH_INT:
    case PH_I32: { i32 v; std::memcpy(&v, p, 4); return v; }
    case PH_I8: { signed char v; std::memcpy(&v, p, 1); return v; }
    case PH_I16: { short v; std::memcpy(&v, p, 2); return v; }
    case PH_I64: { i64 v; std::memcpy(&v, p, 8); return v; }
    default: return i64(ph_as_f64(t, p));
    }
}
void ph_store(int t, f64 v, unsigned char *p) {
    switch (t) {
    case PH_INT:
    case PH_I32: { i32 x = i32(v); std::memcpy(p, &x, 4); break; }
    case PH_I8: { signed char x = (signed char) v; std::memcpy(p, &x, 1); break; }
    case PH_I16: { short x = short(v); std::memcpy(p, &x, 2); break; }
    case PH_I64: { i64 x = i64(v); std::memcpy(p, &x, 8); break; }
    case PH_F32: { float x = float(v); std::memcpy(p, &x, 4); break; }
    default: std::memcpy(p, &v, 8);
    }
}

struct PhTable { ///< PhantomDumpTableHeader<T>: (tag, value) pairs in file order, values as raw bytes
    std::vector<std::string> tags;
    std::vector<unsigned char> vals;
};
struct PhArray { ///< PhantomDumpBlockArray<T>
    std::string tag;
    std::vector<unsigned char> vals;
};
struct PhBlock { ///< PhantomDumpBlock
    i64 tot_count = 0;
    std::array<std::vector<PhArray>, 8> arrays;
};

struct PhDump { ///< PhantomDump
    i32 i1 = 60769, i2 = 60878, iversion = 1, i3 = 690706;
    f64 r1 = 60878;
    std::string fileid;
    std::array<PhTable, 8> tables;
    std::vector<PhBlock> blocks;

    void add(int t, const std::string &tag, f64 v) { // PhantomDumpTableHeader::add
        tables[t].tags.push_back(pad16(tag));
        size_t o = tables[t].vals.size();
        tables[t].vals.resize(o + kPhSize[t]);
        ph_store(t, v, tables[t].vals.data() + o);
    }
    /// PhantomDumpTableHeader::fetch: the LAST entry with this tag
    const unsigned char *fetch(int t, const std::string &tag16) const {
        const unsigned char *r = nullptr;
        for (size_t k = 0; k < tables[t].tags.size(); k++)
            if (tables[t].tags[k] == tag16)
                r = tables[t].vals.data() + k * kPhSize[t];
        return r;
    }
    bool has_header_entry(const std::string &s) const {
        for (int t = 0; t < 8; t++)
            if (fetch(t, pad16(s)))
                return true;
        return false;
    }
    f64 read_header_float(const std::string &s) const { // fort_real, f32, f64 in this order
        for (int t : {PH_REAL, PH_F32, PH_F64})
            if (auto p = fetch(t, pad16(s)))
                return ph_as_f64(t, p);
        throw std::runtime_error("the entry cannot be found : " + pad16(s));
    }
    i64 read_header_int(const std::string &s) const { // fort_int, i8, i16, i32, i64 in this order
        for (int t : {PH_INT, PH_I8, PH_I16, PH_I32, PH_I64})
            if (auto p = fetch(t, pad16(s)))
                return ph_as_i64(t, p);
        throw std::runtime_error("the entry cannot be found");
    }
    std::vector<f64> read_header_floats(const std::string &s) const {
        std::vector<f64> v;
        for (int t : {PH_REAL, PH_F32, PH_F64})
            for (size_t k = 0; k < tables[t].tags.size(); k++)
                if (tables[t].tags[k] == pad16(s))
                    v.push_back(ph_as_f64(t, tables[t].vals.data() + k * kPhSize[t]));
        return v;
    }
    /// PhantomDumpBlock::fill_vec: every array of the block with this tag, whatever its type, appended
    void fill_vec(size_t iblock, const std::string &name, std::vector<f64> &out) const {
        if (iblock >= blocks.size())
            return;
        const std::string tag = pad16(name);
        for (int t = 0; t < 8; t++)
            for (auto &a : blocks[iblock].arrays[t])
                if (a.tag == tag) {
                    const size_t n = a.vals.size() / kPhSize[t];
                    for (size_t k = 0; k < n; k++)
                        out.push_back(ph_as_f64(t, a.vals.data() + k * kPhSize[t]));
                }
    }
};

/// FortranIOFile, write side: records appended to a byte vector
struct FortranOut {
    std::vector<unsigned char> d;
    void raw(const void *p, size_t n) {
        const unsigned char *c = static_cast<const unsigned char *>(p);
        d.insert(d.end(), c, c + n);
    }
    void record(const void *p, size_t n) {
        if (n > size_t(std::numeric_limits<i32>::max()))
            throw std::overflow_error("phantom dump: a record longer than 2^31 bytes");
        i32 len = i32(n);
        raw(&len, 4), raw(p, n), raw(&len, 4);
    }
};

/// FortranIOFile, read side: records of a file held in memory, every byte count checked on both sides
struct FortranIn {
    const std::vector<unsigned char> &d;
    size_t pos = 0;
    explicit FortranIn(const std::vector<unsigned char> &b) : d(b) {}
    i32 marker() {
        if (pos + 4 > d.size())
            throw std::runtime_error("phantom dump: truncated file");
        i32 v;
        std::memcpy(&v, d.data() + pos, 4);
        pos += 4;
        return v;
    }
    const unsigned char *record(size_t expect) {
        i32 n = marker();
        if (n < 0 || size_t(n) != expect)
            throw std::runtime_error("the byte count is not correct");
        if (pos + expect + 4 > d.size())
            throw std::runtime_error("phantom dump: truncated file");
        const unsigned char *p = d.data() + pos;
        pos += expect;
        if (marker() != n)
            throw std::runtime_error("fortran 4 bytes invalid");
        return p;
    }
    bool finished() const { return pos == d.size(); }
};

void ph_write_header(const PhDump &ph, FortranOut &o) { // PhantomDump::gen_file up to the block table
    unsigned char first[24];
    std::memcpy(first, &ph.i1, 4), std::memcpy(first + 4, &ph.r1, 8), std::memcpy(first + 12, &ph.i2, 4);
    std::memcpy(first + 16, &ph.iversion, 4), std::memcpy(first + 20, &ph.i3, 4);
    o.record(first, 24);
    std::string id = ph.fileid;
    id.resize(100, ' ');
    o.record(id.data(), 100);
    for (int t = 0; t < 8; t++) {
        const PhTable &tb = ph.tables[t];
        i32 nvars         = i32(tb.tags.size());
        o.record(&nvars, 4);
        if (nvars == 0)
            continue;
        std::string tags;
        for (auto &s : tb.tags)
            tags += s.substr(0, 16);
        o.record(tags.data(), tags.size());
        o.record(tb.vals.data(), tb.vals.size());
    }
}
void ph_write_block_table(const std::vector<std::pair<i64, std::array<i32, 8>>> &bt, FortranOut &o) {
    i32 nblocks = i32(bt.size());
    o.record(&nblocks, 4);
    for (auto &b : bt) {
        unsigned char rec[40];
        std::memcpy(rec, &b.first, 8);
        std::memcpy(rec + 8, b.second.data(), 32);
        o.record(rec, 40);
    }
}

std::vector<unsigned char> ph_serialize(const PhDump &ph) { // PhantomDump::gen_file
    FortranOut o;
    ph_write_header(ph, o);
    std::vector<std::pair<i64, std::array<i32, 8>>> bt;
    for (auto &b : ph.blocks) {
        std::array<i32, 8> c;
        for (int t = 0; t < 8; t++)
            c[t] = i32(b.arrays[t].size());
        bt.push_back({b.tot_count, c});
    }
    ph_write_block_table(bt, o);
    for (auto &b : ph.blocks)
        for (int t = 0; t < 8; t++)
            for (auto &a : b.arrays[t]) {
                if (a.vals.size() < size_t(b.tot_count) * kPhSize[t]) // write_val_array: val count higher than vec size
                    throw std::invalid_argument("val count is higher than vec size");
                std::string tag = a.tag;
                tag.resize(16, ' ');
                o.record(tag.data(), 16);
                o.record(a.vals.data(), size_t(b.tot_count) * kPhSize[t]);
            }
    return std::move(o.d);
}

PhDump ph_parse(const std::vector<unsigned char> &bytes) { // PhantomDump::from_file
    FortranIn in(bytes);
    PhDump ph;
    const unsigned char *f = in.record(24);
    std::memcpy(&ph.i1, f, 4), std::memcpy(&ph.r1, f + 4, 8), std::memcpy(&ph.i2, f + 12, 4);
    std::memcpy(&ph.iversion, f + 16, 4), std::memcpy(&ph.i3, f + 20, 4);
    if (ph.i1 != 60769 || ph.i2 != 60878 || ph.i3 != 690706 || ph.r1 != f64(ph.i2)) // check_magic_numbers
        throw std::runtime_error("phantom dump: wrong magic numbers");
    ph.fileid.assign(reinterpret_cast<const char *>(in.record(100)), 100);
    for (int t = 0; t < 8; t++) {
        i32 nvars;
        std::memcpy(&nvars, in.record(4), 4);
        if (nvars == 0)
            continue;
        if (nvars < 0)
            throw std::runtime_error("phantom dump: negative header length");
        const char *tags = reinterpret_cast<const char *>(in.record(size_t(nvars) * 16));
        for (i32 k = 0; k < nvars; k++)
            ph.tables[t].tags.emplace_back(tags + size_t(k) * 16, 16);
        const unsigned char *v = in.record(size_t(nvars) * kPhSize[t]);
        ph.tables[t].vals.assign(v, v + size_t(nvars) * kPhSize[t]);
    }
    i32 nblocks;
    std::memcpy(&nblocks, in.record(4), 4);
    if (nblocks < 0)
        throw std::runtime_error("phantom dump: negative block count");
    std::vector<std::array<i32, 8>> counts(nblocks);
    ph.blocks.resize(nblocks);
    for (i32 b = 0; b < nblocks; b++) {
        const unsigned char *r = in.record(40);
        std::memcpy(&ph.blocks[b].tot_count, r, 8);
        std::memcpy(counts[b].data(), r + 8, 32);
        if (ph.blocks[b].tot_count < 0)
            throw std::runtime_error("phantom dump: negative block length");
    }
    for (i32 b = 0; b < nblocks; b++)
        for (int t = 0; t < 8; t++)
            for (i32 j = 0; j < counts[b][t]; j++) {
                PhArray a;
                a.tag.assign(reinterpret_cast<const char *>(in.record(16)), 16);
                const size_t bytes_a   = size_t(ph.blocks[b].tot_count) * kPhSize[t];
                const unsigned char *v = in.record(bytes_a);
                a.vals.assign(v, v + bytes_a);
                ph.blocks[b].arrays[t].push_back(std::move(a));
            }
    if (!in.finished())
        fprintf(stderr, "[PhantomReader] some data was not read\n");
    return ph;
}

std::vector<unsigned char> read_file(const std::string &fname) {
    int fd = ::open(fname.c_str(), O_RDONLY);
    if (fd < 0)
        throw std::runtime_error("cannot open " + fname);
    struct stat st;
    if (::fstat(fd, &st) != 0) {
        ::close(fd);
        throw std::runtime_error("cannot stat " + fname);
    }
    std::vector<unsigned char> b(size_t(st.st_size));
    size_t got = 0;
    while (got < b.size()) {
        ssize_t r = ::read(fd, b.data() + got, b.size() - got);
        if (r <= 0) {
            ::close(fd);
            throw std::runtime_error("read failed: " + fname);
        }
        got += size_t(r);
    }
    ::close(fd);
    return b;
}
void write_file(const std::string &fname, const std::vector<unsigned char> &b) {
    int fd = ::open(fname.c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0644);
    if (fd < 0)
        throw std::runtime_error("cannot create " + fname);
    try {
        pwrite_all(fd, b.data(), b.size(), 0);
    } catch (...) {
        ::close(fd);
        throw;
    }
    ::close(fd);
}

/// get_shamrock_eosconfig / read_headeropts_eos (Phantom2Shamrock.cpp:27-67, PhantomDumpEOSUtils.cpp:62-135)
void ph_eos_to_config(const PhDump &ph, bool bypass_error, shamb200_solver_config &cfg) {
    const i64 ieos = ph.read_header_int("ieos");
    if (ieos == 1 || ieos == 2 || ieos == 3) {
        const f64 gamma = ph.read_header_float("gamma");
        const f64 polyk = 2.0 / 3.0 * ph.read_header_float("RK2");
        (void) ph.read_header_float("polyk2"); // read_headeropts_eos needs these entries to exist
        const f64 qfacdisc = ph.read_header_float("qfacdisc");
        (void) ph.read_header_float("qfacdisc2");
        (void) ph.read_header_int("isink");
        if (ieos == 1) {
            cfg.eos = SHAMB200_EOS_ISOTHERMAL, cfg.cs0 = std::sqrt(polyk);
        } else if (ieos == 2) {
            cfg.eos = SHAMB200_EOS_ADIABATIC, cfg.gamma = gamma;
        } else {
            cfg.eos = SHAMB200_EOS_LOCALLY_ISOTHERMAL_LP07, cfg.cs0 = std::sqrt(polyk), cfg.eos_q = qfacdisc, cfg.eos_r0 = 1;
        }
        return;
    }
    const std::string msg = "loading phantom ieos=" + std::to_string(ieos) + " is not implemented in shamrock";
    if (!bypass_error)
        throw std::runtime_error(msg);
    fprintf(stderr, "[SPH] warning: %s\n", msg.c_str());
}

/// write_shamrock_eos_in_phantom_dump + write_headeropts_eos (Phantom2Shamrock.cpp:69-117,
/// PhantomDumpEOSUtils.cpp:43-60).  The reference leaves isink / polyk / polyk2 of its EOSPhConfig
/// uninitialised where an EOS does not set them; this writer puts 0 there.
void ph_write_eos(PhDump &ph, const shamb200_solver_config &cfg) {
    int ieos     = 0;
    f64 gamma    = 1, polyk = 0, qfacdisc = 0.75;
    if (cfg.eos == SHAMB200_EOS_ISOTHERMAL) {
        ieos = 1, polyk = cfg.cs0 * cfg.cs0;
    } else if (cfg.eos == SHAMB200_EOS_ADIABATIC) {
        ieos = 2, gamma = cfg.gamma;
    } else if (cfg.eos == SHAMB200_EOS_LOCALLY_ISOTHERMAL_LP07) {
        ieos = 3, polyk = cfg.cs0 * cfg.cs0 / (cfg.eos_r0 * cfg.eos_r0), qfacdisc = cfg.eos_q;
    } else {
        throw std::runtime_error("The current shamrock EOS is not implemented in phantom dump conversion");
    }
    ph.add(PH_I32, "ieos", ieos);
    ph.add(PH_I32, "isink", 0);
    ph.add(PH_REAL, "gamma", gamma);
    ph.add(PH_REAL, "RK2", 1.5 * polyk);
    ph.add(PH_REAL, "polyk2", 0);
    ph.add(PH_REAL, "qfacdisc", qfacdisc);
    ph.add(PH_REAL, "qfacdisc2", 0.75);
}

bool cfg_has_alpha(const shamb200_solver_config &c) { return c.av == SHAMB200_AV_MM97 || c.av == SHAMB200_AV_CD10; }

// ---------------------------------------------------------------------------------------------
// device side of the VTK writer
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 bswap32(u32 v) { return __byte_perm(v, 0, 0x0123); }

/// shamrock::details::to_vtk_buf_type<f32>: value -> f32 -> big endian (io/details/bufToVtkBuf.hpp:26-70)
__global__ void __launch_bounds__(256) vtk_f32_kernel(u64 n, const f64 *__restrict__ in, u32 *__restrict__ out) {
    u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n)
        out[i] = bswap32(__float_as_uint(float(in[i])));
}
/// rho = m (hfact / h)^3 (VTKDump.cpp:52-77, sph/math/density.hpp:23-41), converted like every other field
__global__ void __launch_bounds__(256) vtk_rho_kernel(u64 n, const f64 *__restrict__ h, f64 mass, f64 hfact, u32 *__restrict__ out) {
    u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n)
        out[i] = bswap32(__float_as_uint(float(rho_h(mass, h[i], hfact))));
}
__global__ void __launch_bounds__(256) vtk_const_i32_kernel(u64 n, i32 v, u32 *__restrict__ out) {
    u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n)
        out[i] = bswap32(u32(v));
}

/// particle counts of all ranks (every rank's local patches, list order) and this rank's offset
struct RankView {
    std::vector<u64> per_rank;
    u64 total = 0, my_off = 0, my_cnt = 0;
};
RankView rank_view(Model &m) {
    RankView v;
    v.per_rank.assign(size_t(m.world), 0);
    for (auto &p : m.patches)
        if (m.is_local(p))
            v.per_rank[m.rank] += p.f.n;
    v.my_cnt = v.per_rank[m.rank];
    comm_allreduce_host_u64(m, v.per_rank.data(), v.per_rank.size(), 0);
    for (int r = 0; r < m.world; r++) {
        if (r == m.rank)
            v.my_off = v.total;
        v.total += v.per_rank[r];
    }
    return v;
}

/// rank 0 creates the file, the others open it once it exists (the dump.cu protocol)
int open_shared(Model &m, const std::string &fname) {
    int fd = -1;
    if (m.rank == 0) {
        fd = ::open(fname.c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0644);
        if (fd < 0)
            throw std::runtime_error("cannot create " + fname);
    }
    u64 ready = 1;
    comm_allreduce_host_u64(m, &ready, 1, 0);
    if (m.rank != 0) {
        fd = ::open(fname.c_str(), O_WRONLY);
        if (fd < 0)
            throw std::runtime_error("cannot open " + fname);
    }
    return fd;
}
void close_shared(Model &m, int fd) {
    ::fsync(fd);
    ::close(fd);
    u64 done = 1;
    comm_allreduce_host_u64(m, &done, 1, 0);
}

/// this rank's values of one field, patch by patch in list order (host copy)
std::vector<f64> gather_local(Model &m, const char *name, int nvar, u64 my_cnt) {
    std::vector<f64> out(size_t(my_cnt) * nvar);
    size_t o = 0;
    for (auto &p : m.patches) {
        if (!m.is_local(p) || p.f.n == 0)
            continue;
        for (auto &r : p.f.all())
            if (std::string(r.name) == name) {
                SB_CUDA_CHECK(cudaMemcpyAsync(
                    out.data() + o, r.buf->p, size_t(p.f.n) * nvar * sizeof(f64), cudaMemcpyDeviceToHost, m.s()));
                o += size_t(p.f.n) * nvar;
            }
    }
    SB_CUDA_CHECK(cudaStreamSynchronize(m.s()));
    return out;
}

} // namespace

// ---------------------------------------------------------------------------------------------
// Model::make_phantom_dump + PhantomDump::gen_file + write_to_file
// ---------------------------------------------------------------------------------------------
void Model::phantom_dump(const std::string &fname) {
    SB_CUDA_CHECK(cudaSetDevice(ctx->device));
    SB_CUDA_CHECK(cudaStreamSynchronize(s()));
    const RankView rv = rank_view(*this);
    const u64 Ntot    = rv.total;
    if (Ntot * 8 > u64(std::numeric_limits<i32>::max()))
        throw std::overflow_error("phantom dump: more than 2^28 particles do not fit a Fortran record");
    PhDump ph;
    ph.fileid = "FT:Phantom Shamrock writer";
    // Model.cpp:1506-1541 (u32 Ntot in the reference)
    for (int t : {PH_INT, PH_I64}) {
        ph.add(t, "nparttot", f64(Ntot));
        ph.add(t, "ntypes", 8);
        ph.add(t, "npartoftype", f64(Ntot));
        for (int k = 0; k < 7; k++)
            ph.add(t, "npartoftype", 0);
    }
    ph.add(PH_INT, "nblocks", 1);
    ph.add(PH_INT, "nptmass", 0); // sinks are outside the path (out of scope)
    ph.add(PH_INT, "ndustlarge", 0);
    ph.add(PH_INT, "ndustsmall", 0);
    ph.add(PH_INT, "idust", 7);
    ph.add(PH_INT, "idtmax_n", 1);
    ph.add(PH_INT, "idtmax_frac", 0);
    ph.add(PH_INT, "idumpfile", 0);
    ph.add(PH_INT, "majorv", 2023);
    ph.add(PH_INT, "minorv", 0);
    ph.add(PH_INT, "microv", 0);
    ph.add(PH_INT, "isink", 0);
    ph.add(PH_I32, "iexternalforce", 0);
    ph_write_eos(ph, cfg);
    const f64 hfact = cfg.kernel == SHAMB200_KERNEL_M6 ? KM6::hfactd : KM4::hfactd;
    ph.add(PH_REAL, "time", time);
    ph.add(PH_REAL, "dtmax", dt);
    ph.add(PH_REAL, "rhozero", 0);
    ph.add(PH_REAL, "hfact", hfact);
    ph.add(PH_REAL, "tolh", 0.0001);
    ph.add(PH_REAL, "C_cour", cfg.cfl_cour);
    ph.add(PH_REAL, "C_force", cfg.cfl_force);
    ph.add(PH_REAL, "alpha", 0);
    ph.add(PH_REAL, "alphau", 1);
    ph.add(PH_REAL, "alphaB", 1);
    ph.add(PH_REAL, "massoftype", cfg.gpart_mass);
    for (int k = 0; k < 7; k++)
        ph.add(PH_REAL, "massoftype", 0);
    ph.add(PH_REAL, "Bextx", 0);
    ph.add(PH_REAL, "Bexty", 0);
    ph.add(PH_REAL, "Bextz", 0);
    ph.add(PH_REAL, "dum", 0);
    if (cfg.bc == SHAMB200_BC_PERIODIC) { // Phantom2Shamrock.cpp:203-209: ymax and zmax are written as bmax.x() there
        ph.add(PH_REAL, "xmin", box_min[0]);
        ph.add(PH_REAL, "xmax", box_max[0]);
        ph.add(PH_REAL, "ymin", box_min[1]);
        ph.add(PH_REAL, "ymax", box_max[0]);
        ph.add(PH_REAL, "zmin", box_min[2]);
        ph.add(PH_REAL, "zmax", box_max[0]);
    }
    ph.add(PH_REAL, "get_conserv", -1);
    ph.add(PH_REAL, "etot_in", 0.59762);
    ph.add(PH_REAL, "angtot_in", 0.0189694);
    ph.add(PH_REAL, "totmom_in", 0.0306284);
    // no unit system in this library's configuration: the reference's "no units are set, defaulting to SI" branch
    ph.add(PH_F64, "udist", 1);
    ph.add(PH_F64, "umass", 1);
    ph.add(PH_F64, "utime", 1);
    ph.add(PH_F64, "umagfd", 3.54491);

    // block 0 (add_pdat_to_phantom_block): fort_real x y z vx vy vz u, f32 h [alpha] [divv]
    const bool has_alpha = cfg_has_alpha(cfg);
    std::vector<std::string> real_tags = {"x", "y", "z", "vx", "vy", "vz", "u"};
    std::vector<std::string> f32_tags  = {"h"};
    if (has_alpha)
        f32_tags.push_back("alpha"), f32_tags.push_back("divv");
    FortranOut head;
    ph_write_header(ph, head);
    std::array<i32, 8> counts{};
    counts[PH_REAL] = i32(real_tags.size());
    counts[PH_F32]  = i32(f32_tags.size());
    ph_write_block_table({{i64(Ntot), counts}}, head);

    const int fd = open_shared(*this, fname);
    try {
        u64 off = 0;
        if (rank == 0)
            pwrite_all(fd, head.d.data(), head.d.size(), 0);
        off = head.d.size();
        // one array: [16][tag][16] [bytes][values][bytes]; the values of rank r sit at my_off inside the record
        auto write_array = [&](const std::string &tag, size_t esz, const void *mine) {
            const i32 len = i32(Ntot * esz);
            if (rank == 0) {
                FortranOut t;
                const std::string tg = pad16(tag);
                t.record(tg.data(), 16);
                t.raw(&len, 4);
                pwrite_all(fd, t.d.data(), t.d.size(), off);
                pwrite_all(fd, &len, 4, off + 28 + Ntot * esz);
            }
            if (rv.my_cnt)
                pwrite_all(fd, mine, size_t(rv.my_cnt) * esz, off + 28 + rv.my_off * esz);
            off += 28 + Ntot * esz + 4;
        };
        std::vector<f64> comp(size_t(rv.my_cnt));
        {
            const std::vector<f64> xyz = gather_local(*this, "xyz", 3, rv.my_cnt);
            for (int c = 0; c < 3; c++) {
                for (u64 i = 0; i < rv.my_cnt; i++)
                    comp[i] = xyz[3 * i + c];
                write_array(real_tags[c], 8, comp.data());
            }
        }
        {
            const std::vector<f64> v = gather_local(*this, "vxyz", 3, rv.my_cnt);
            for (int c = 0; c < 3; c++) {
                for (u64 i = 0; i < rv.my_cnt; i++)
                    comp[i] = v[3 * i + c];
                write_array(real_tags[3 + c], 8, comp.data());
            }
        }
        {
            const std::vector<f64> u = gather_local(*this, "uint", 1, rv.my_cnt);
            write_array("u", 8, u.data());
        }
        std::vector<float> f32v(size_t(rv.my_cnt));
        auto write_f32 = [&](const std::string &tag, const char *field) {
            const std::vector<f64> v = gather_local(*this, field, 1, rv.my_cnt);
            for (u64 i = 0; i < rv.my_cnt; i++)
                f32v[i] = float(v[i]);
            write_array(tag, 4, f32v.data());
        };
        write_f32("h", "hpart");
        if (has_alpha) {
            write_f32("alpha", "alpha_AV");
            write_f32("divv", "divv");
        }
    } catch (...) {
        ::close(fd);
        throw;
    }
    close_shared(*this, fd);
}

// ---------------------------------------------------------------------------------------------
// Model::gen_config_from_phantom_dump / Model::init_from_phantom_dump
// ---------------------------------------------------------------------------------------------
void phantom_gen_config(const std::string &fname, bool bypass_error, shamb200_solver_config &cfg) {
    const PhDump ph = ph_parse(read_file(fname));
    const std::vector<f64> massoftype = ph.read_header_floats("massoftype");
    if (massoftype.empty())
        throw std::runtime_error("the entry cannot be found : massoftype");
    cfg.gpart_mass = massoftype[0];
    cfg.cfl_cour   = ph.read_header_float("C_cour");
    cfg.cfl_force  = ph.read_header_float("C_force");
    ph_eos_to_config(ph, bypass_error, cfg);
    // get_shamrock_avconfig: set_varying_cd10(0, 1, 0.1, alphau, 2)
    cfg.av          = SHAMB200_AV_CD10;
    cfg.alpha_min   = 0;
    cfg.alpha_max   = 1;
    cfg.sigma_decay = 0.1;
    cfg.alpha_u     = ph.read_header_float("alphau");
    cfg.beta_AV     = 2;
    // get_shamrock_units needs the four unit entries (this library's configuration carries no unit system)
    (void) ph.read_header_float("udist"), (void) ph.read_header_float("umass");
    (void) ph.read_header_float("utime"), (void) ph.read_header_float("umagfd");
    // xmin ... zmax are in the header only in periodic mode in phantom
    cfg.bc = ph.has_header_entry("xmin") ? SHAMB200_BC_PERIODIC : SHAMB200_BC_FREE;
}

u64 Model::init_from_phantom_dump(const std::string &fname, f64 hpart_fact_load) {
    SB_CUDA_CHECK(cudaSetDevice(ctx->device));
    const PhDump ph = ph_parse(read_file(fname));
    std::vector<f64> x, y, z, vx, vy, vz, h, u, alpha;
    ph.fill_vec(0, "x", x), ph.fill_vec(0, "y", y), ph.fill_vec(0, "z", z);
    if (x.size() != y.size() || x.size() != z.size())
        throw std::runtime_error("phantom dump: x, y and z have different lengths");
    f64 bmin[3], bmax[3];
    if (ph.has_header_entry("xmin")) {
        bmin[0] = ph.read_header_float("xmin"), bmax[0] = ph.read_header_float("xmax");
        bmin[1] = ph.read_header_float("ymin"), bmax[1] = ph.read_header_float("ymax");
        bmin[2] = ph.read_header_float("zmin"), bmax[2] = ph.read_header_float("zmax");
    } else { // the bounding box of the positions, grown by 20 % about its centre
        if (x.empty())
            throw std::runtime_error("phantom dump: no particles and no box in the header");
        const std::vector<f64> *c[3] = {&x, &y, &z};
        for (int d = 0; d < 3; d++) {
            f64 lo = (*c[d])[0], hi = lo;
            for (f64 v : *c[d])
                lo = std::min(lo, v), hi = std::max(hi, v);
            const f64 center = (lo + hi) * 0.5, half = (hi - lo) * 0.5 * 1.2;
            bmin[d] = center - half, bmax[d] = center + half;
        }
    }
    // Model::resize_simulation_box: the integer patch coordinates stay, the simulation box changes
    if (patches.empty()) {
        set_box(bmin, bmax, 1, 1, 1);
    } else {
        for (int d = 0; d < 3; d++)
            box_min[d] = bmin[d], box_max[d] = bmax[d];
        for (auto &p : patches)
            for (int d = 0; d < 3; d++) {
                const f64 fact = (box_max[d] - box_min[d]) / f64(kPatchGrid);
                p.lo[d]        = f64(p.cmin[d]) * fact + box_min[d];
                p.hi[d]        = f64(p.cmax[d] + 1) * fact + box_min[d];
            }
        refresh_boxes();
    }
    ph.fill_vec(0, "h", h);
    ph.fill_vec(0, "vx", vx), ph.fill_vec(0, "vy", vy), ph.fill_vec(0, "vz", vz);
    ph.fill_vec(0, "u", u), ph.fill_vec(0, "alpha", alpha);
    const size_t n = x.size();
    if (h.size() != n || vx.size() != n || vy.size() != n || vz.size() != n || (!u.empty() && u.size() != n)
        || (!alpha.empty() && alpha.size() != n))
        throw std::runtime_error("phantom dump: the arrays of block 0 have different lengths");
    time = ph.read_header_float("time"); // solver.set_time
    // a particle goes to the patch that contains it, if its h is not negative (dead particles of phantom)
    std::vector<f64> pxyz, pv, ph_, pu, pa;
    for (size_t i = 0; i < n; i++) {
        const f64 r[3] = {x[i], y[i], z[i]};
        bool inside    = h[i] >= 0;
        for (int d = 0; d < 3; d++)
            inside = inside && box_min[d] <= r[d] && r[d] < box_max[d];
        if (!inside)
            continue;
        pxyz.insert(pxyz.end(), {x[i], y[i], z[i]});
        pv.insert(pv.end(), {vx[i], vy[i], vz[i]});
        ph_.push_back(h[i] * hpart_fact_load);
        if (!u.empty())
            pu.push_back(u[i]);
        if (!alpha.empty())
            pa.push_back(alpha[i]);
    }
    const u64 kept = ph_.size();
    if (kept)
        push_particles(kept, pxyz.data(), pv.data(), ph_.data(), pu.empty() ? nullptr : pu.data(),
                       pa.empty() ? nullptr : pa.data());
    refresh_counts();
    return kept;
}

// ---- file-level helpers of the C ABI (no device needed) ----
void phantom_copy(const std::string &in, const std::string &out) { write_file(out, ph_serialize(ph_parse(read_file(in)))); }
int phantom_header(const std::string &fname, const std::string &key, int want_int, f64 *fval, i64 *ival) {
    const PhDump ph = ph_parse(read_file(fname));
    if (!ph.has_header_entry(key))
        return 0;
    if (want_int)
        *ival = ph.read_header_int(key);
    else
        *fval = ph.read_header_float(key);
    return 1;
}
/// compare_phantom_dumps (PhantomDump.cpp:389-467): the header entries as one (tag -> value) map per dump, the
/// number of missing / extra / different keys
u64 phantom_compare(const std::string &fa, const std::string &fb) {
    auto load = [](const std::string &f) {
        const PhDump ph = ph_parse(read_file(f));
        std::map<std::string, f64> h;
        for (int t = 0; t < 8; t++)
            for (size_t k = 0; k < ph.tables[t].tags.size(); k++)
                h[ph.tables[t].tags[k]] = ph_as_f64(t, ph.tables[t].vals.data() + k * kPhSize[t]);
        return h;
    };
    const auto a = load(fa), b = load(fb);
    u64 offenses = 0;
    for (auto &kv : a) {
        auto it = b.find(kv.first);
        if (it == b.end() || it->second != kv.second)
            offenses++;
    }
    for (auto &kv : b)
        if (!a.count(kv.first))
            offenses++;
    return offenses;
}

// ---------------------------------------------------------------------------------------------
// Model::do_vtk_dump (modules::VTKDump::do_dump)
// ---------------------------------------------------------------------------------------------
void Model::vtk_dump(const std::string &fname, bool add_patch_world_id) {
    SB_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (fname.find(".vtk") == std::string::npos)
        throw std::invalid_argument("the extension should be .vtk");
    SB_CUDA_CHECK(cudaStreamSynchronize(s()));
    const RankView rv = rank_view(*this);
    if (rv.total > 0xFFFFFFFFull / 3)
        throw std::overflow_error("vtk dump: the reference's writer counts values in u32");
    const bool has_alpha = cfg_has_alpha(cfg), has_cd10 = cfg.av == SHAMB200_AV_CD10;
    const bool has_cs = has_alpha || cfg.eos == SHAMB200_EOS_LOCALLY_ISOTHERMAL_LP07;
    const f64 hfact   = cfg.kernel == SHAMB200_KERNEL_M6 ? KM6::hfactd : KM4::hfactd;
    const int fd      = open_shared(*this, fname);
    DevBuf<u32> conv;
    std::vector<u32> host;
    try {
        u64 head = 0;
        auto text = [&](const std::string &t) { // write_header_raw: rank 0 writes, everybody advances
            if (rank == 0)
                pwrite_all(fd, t.data(), t.size(), head);
            head += t.size();
        };
        // converted values of this rank at its offset inside a section of `total * nvar` 4-byte values
        auto section = [&](int nvar, const std::function<void(PatchD &, u32 *)> &convert) {
            conv.ensure(size_t(rv.my_cnt) * nvar + 1);
            size_t o = 0;
            for (auto &p : patches) {
                if (!is_local(p) || p.f.n == 0)
                    continue;
                convert(p, conv.p + o);
                o += size_t(p.f.n) * nvar;
            }
            host.resize(o);
            if (o) {
                SB_LAUNCH_CHECK();
                SB_CUDA_CHECK(cudaMemcpyAsync(host.data(), conv.p, o * 4, cudaMemcpyDeviceToHost, s()));
                SB_CUDA_CHECK(cudaStreamSynchronize(s()));
                pwrite_all(fd, host.data(), o * 4, head + rv.my_off * nvar * 4);
            }
            head += rv.total * nvar * 4;
        };
        auto field_f64 = [&](const char *name, int nvar) {
            return [this, name, nvar](PatchD &p, u32 *out) {
                for (auto &r : p.f.all())
                    if (std::string(r.name) == name) {
                        const u64 cnt = u64(p.f.n) * nvar;
                        vtk_f32_kernel<<<grid_for(cnt, 256), 256, 0, s()>>>(cnt, r.buf->p, out);
                        SB_COUNT_LAUNCH();
                    }
            };
        };
        auto named = [&](const std::string &name, int nvar, const char *type) {
            text("\n" + name + " " + std::to_string(nvar) + " " + std::to_string(rv.total) + " " + type + "\n");
        };
        text("# vtk DataFile Version 4.2\nvtk output\nBINARY\nDATASET UNSTRUCTURED_GRID");
        text("\n\nPOINTS " + std::to_string(rv.total) + " float\n");
        section(3, field_f64("xyz", 3));
        text("\n\nPOINT_DATA " + std::to_string(rv.total));
        u32 fnum = 5 + (add_patch_world_id ? 2 : 0) + (has_alpha ? 2 : 0) + (has_cd10 ? 2 : 0) + (has_cs ? 1 : 0);
        text("\nFIELD FieldData " + std::to_string(fnum));
        if (add_patch_world_id) {
            named("patchid", 1, "int");
            section(1, [this](PatchD &p, u32 *out) {
                vtk_const_i32_kernel<<<grid_for(p.f.n, 256), 256, 0, s()>>>(p.f.n, i32(p.id), out);
                SB_COUNT_LAUNCH();
            });
            named("world_rank", 1, "int");
            section(1, [this](PatchD &p, u32 *out) {
                vtk_const_i32_kernel<<<grid_for(p.f.n, 256), 256, 0, s()>>>(p.f.n, i32(rank), out);
                SB_COUNT_LAUNCH();
            });
        }
        named("h", 1, "float"), section(1, field_f64("hpart", 1));
        named("u", 1, "float"), section(1, field_f64("uint", 1));
        named("v", 3, "float"), section(3, field_f64("vxyz", 3));
        named("a", 3, "float"), section(3, field_f64("axyz", 3));
        if (has_alpha) {
            named("alpha_AV", 1, "float"), section(1, field_f64("alpha_AV", 1));
            named("divv", 1, "float"), section(1, field_f64("divv", 1));
        }
        if (has_cd10) {
            named("dtdivv", 1, "float"), section(1, field_f64("dtdivv", 1));
            named("curlv", 3, "float"), section(3, field_f64("curlv", 3));
        }
        if (has_cs)
            named("soundspeed", 1, "float"), section(1, field_f64("soundspeed", 1));
        named("rho", 1, "float");
        section(1, [this, hfact](PatchD &p, u32 *out) {
            vtk_rho_kernel<<<grid_for(p.f.n, 256), 256, 0, s()>>>(p.f.n, p.f.hpart.p, cfg.gpart_mass, hfact, out);
            SB_COUNT_LAUNCH();
        });
    } catch (...) {
        ::close(fd);
        throw;
    }
    close_shared(*this, fd);
}

} // namespace sb
