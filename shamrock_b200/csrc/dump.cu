// dump.cu — checkpoint / restart of the model state.
//
// Container of the reference's dump (shamrock/src/io/ShamrockDump.cpp:25-274, shamalgs/collective/io.hpp:96-147):
//   [size_t len][user metadata JSON]  [size_t len][patch metadata JSON]  [size_t len][table JSON]  [patch blobs]
// where the table holds {"pids", "bytecounts", "offsets"} (offsets relative to the end of the headers) and every
// rank writes the blobs of the patches it owns at their offsets of ONE file.  The user metadata carries what
// Model::dump stores (solver configuration, time, next dt, cfl multiplier: shammodels/sph/include/shammodels/sph/
// Model.hpp:906-995), the patch metadata the patch list (id, integer coordinates, owner), the simulation box
// and the scheduler criteria.  On load the owners are folded onto the ranks present (owner % world_size,
// ShamrockDump.cpp:206-208), so a dump restarts on a different number of GPUs.
// The blob of a patch: for every field of the main layout, in layout order, [u64 object count][values, f64
// packed].  This encoding is this library's own ("shamb200-1"): the reference's blobs are written by its
// SerializeHelper, whose byte layout cannot be checked here (the reference does not build in this image), so
// byte compatibility with .sham files is not claimed — the container, the metadata split and the restart
// semantics are the reference's.
#include "solver.cuh"
#include <cctype>
#include <cstring>
#include <fcntl.h>
#include <sstream>
#include <sys/stat.h>
#include <unistd.h>

namespace sb {

namespace {

// ---- a minimal JSON value (objects, arrays, numbers, strings): enough for the headers this file writes ----
struct JVal {
    enum Kind { NUM, STR, ARR, OBJ } kind = NUM;
    double num = 0;
    std::string str, raw; ///< raw: the number's text (u64 values do not survive a double)
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;
    const JVal &at(const std::string &k) const {
        for (auto &kv : obj)
            if (kv.first == k)
                return kv.second;
        throw std::runtime_error("dump: missing key " + k);
    }
    bool has(const std::string &k) const {
        for (auto &kv : obj)
            if (kv.first == k)
                return true;
        return false;
    }
    u64 as_u64() const { return std::strtoull(raw.c_str(), nullptr, 10); }
    i64 as_i64() const { return std::strtoll(raw.c_str(), nullptr, 10); }
    f64 as_f64() const { return std::strtod(raw.c_str(), nullptr); }
};
struct JParser {
    const std::string &s;
    size_t i = 0;
    explicit JParser(const std::string &t) : s(t) {}
    void ws() {
        while (i < s.size() && std::isspace((unsigned char) s[i]))
            i++;
    }
    JVal parse() {
        ws();
        if (i >= s.size())
            throw std::runtime_error("dump: truncated JSON header");
        JVal v;
        if (s[i] == '{') {
            v.kind = JVal::OBJ;
            i++;
            ws();
            if (s[i] == '}') {
                i++;
                return v;
            }
            for (;;) {
                ws();
                JVal k = parse();
                ws();
                if (s[i++] != ':')
                    throw std::runtime_error("dump: malformed JSON header");
                v.obj.emplace_back(k.str, parse());
                ws();
                if (s[i] == ',') {
                    i++;
                    continue;
                }
                if (s[i++] != '}')
                    throw std::runtime_error("dump: malformed JSON header");
                return v;
            }
        }
        if (s[i] == '[') {
            v.kind = JVal::ARR;
            i++;
            ws();
            if (s[i] == ']') {
                i++;
                return v;
            }
            for (;;) {
                v.arr.push_back(parse());
                ws();
                if (s[i] == ',') {
                    i++;
                    continue;
                }
                if (s[i++] != ']')
                    throw std::runtime_error("dump: malformed JSON header");
                return v;
            }
        }
        if (s[i] == '"') {
            v.kind = JVal::STR;
            i++;
            while (i < s.size() && s[i] != '"')
                v.str += s[i++];
            i++;
            return v;
        }
        size_t j = i;
        while (j < s.size() && (std::isdigit((unsigned char) s[j]) || std::strchr("+-.eE", s[j]) || std::isalpha((unsigned char) s[j])))
            j++;
        v.raw = s.substr(i, j - i);
        v.num = std::strtod(v.raw.c_str(), nullptr);
        i     = j;
        return v;
    }
};

std::string f64_text(f64 v) {
    char b[40];
    snprintf(b, sizeof(b), "%.17g", v);
    return b;
}
std::string hex_of(const void *p, size_t n) {
    static const char *d = "0123456789abcdef";
    std::string o;
    const unsigned char *c = static_cast<const unsigned char *>(p);
    for (size_t k = 0; k < n; k++) {
        o += d[c[k] >> 4];
        o += d[c[k] & 15];
    }
    return o;
}
void unhex(const std::string &h, void *p, size_t n) {
    if (h.size() != 2 * n)
        throw std::runtime_error("dump: the solver configuration of the dump has another size (other library version)");
    auto v = [](char c) { return c <= '9' ? c - '0' : c - 'a' + 10; };
    unsigned char *o = static_cast<unsigned char *>(p);
    for (size_t k = 0; k < n; k++)
        o[k] = (unsigned char) ((v(h[2 * k]) << 4) | v(h[2 * k + 1]));
}

void pwrite_all(int fd, const void *buf, size_t n, u64 off) {
    const char *p = static_cast<const char *>(buf);
    while (n) {
        ssize_t w = ::pwrite(fd, p, n, off_t(off));
        if (w <= 0)
            throw std::runtime_error("dump: write failed");
        p += w, n -= size_t(w), off += u64(w);
    }
}
void pread_all(int fd, void *buf, size_t n, u64 off) {
    char *p = static_cast<char *>(buf);
    while (n) {
        ssize_t r = ::pread(fd, p, n, off_t(off));
        if (r <= 0)
            throw std::runtime_error("dump: the file is shorter than its table says");
        p += r, n -= size_t(r), off += u64(r);
    }
}
/// shamalgs::collective::write_header: [size_t length][bytes] (io.hpp:96-147); only rank 0 writes
void write_header(int fd, bool writer, const std::string &s, u64 &head) {
    size_t len = s.size();
    if (writer) {
        pwrite_all(fd, &len, sizeof(len), head);
        pwrite_all(fd, s.data(), len, head + sizeof(len));
    }
    head += sizeof(len) + len;
}
std::string read_header(int fd, u64 &head) {
    size_t len = 0;
    pread_all(fd, &len, sizeof(len), head);
    if (len > (size_t(1) << 32))
        throw std::runtime_error("dump: not a shamb200 dump (header length)");
    std::string s(len, '\0');
    pread_all(fd, s.data(), len, head + sizeof(len));
    head += sizeof(len) + len;
    return s;
}

} // namespace

void Model::dump(const std::string &fname) {
    SB_CUDA_CHECK(cudaSetDevice(ctx->device));
    SB_CUDA_CHECK(cudaStreamSynchronize(s()));
    refresh_counts();
    const size_t np = patches.size();
    const auto refs = patches.empty() ? std::vector<PatchFields::Ref>{} : patches[0].f.all();
    // ---- headers (identical on every rank: pure functions of replicated state)
    std::ostringstream user;
    user << "{\"format\": \"shamb200-1\", \"solver_config_hex\": \"" << hex_of(&cfg, sizeof(cfg)) << "\", \"time\": "
         << f64_text(time) << ", \"dt\": " << f64_text(dt) << ", \"cfl_multiplier\": " << f64_text(cfl_multiplier)
         << ", \"step_count\": " << step_count << "}";
    std::ostringstream pm;
    pm << "{\"crit_patch_split\": " << crit_split << ", \"crit_patch_merge\": " << crit_merge
       << ", \"scheduler_freq\": " << scheduler_freq << ", \"next_patch_id\": " << next_patch_id << ", \"sim_box\": ["
       << f64_text(box_min[0]) << ", " << f64_text(box_min[1]) << ", " << f64_text(box_min[2]) << ", "
       << f64_text(box_max[0]) << ", " << f64_text(box_max[1]) << ", " << f64_text(box_max[2]) << "], \"patchdata_layout\": [";
    for (size_t r = 0; r < refs.size(); r++)
        pm << (r ? ", " : "") << "{\"field_name\": \"" << refs[r].name << "\", \"nvar\": " << refs[r].nvar << ", \"type\": \"f64\"}";
    pm << "], \"patchlist\": [";
    for (size_t k = 0; k < np; k++) {
        const PatchD &p = patches[k];
        pm << (k ? ", " : "") << "{\"id_patch\": " << p.id << ", \"load_value\": " << patch_counts[k] << ", \"coord_min\": ["
           << p.cmin[0] << ", " << p.cmin[1] << ", " << p.cmin[2] << "], \"coord_max\": [" << p.cmax[0] << ", " << p.cmax[1]
           << ", " << p.cmax[2] << "], \"node_owner_id\": " << p.owner << "}";
    }
    pm << "]}";
    size_t per_obj = 0;
    for (auto &r : refs)
        per_obj += size_t(r.nvar) * sizeof(f64);
    std::vector<u64> bytecounts(np), offsets(np);
    u64 run = 0;
    for (size_t k = 0; k < np; k++) {
        bytecounts[k] = refs.size() * sizeof(u64) + patch_counts[k] * per_obj;
        offsets[k]    = run;
        run += bytecounts[k];
    }
    std::ostringstream tb;
    tb << "{\"pids\": [";
    for (size_t k = 0; k < np; k++)
        tb << (k ? ", " : "") << patches[k].id;
    tb << "], \"bytecounts\": [";
    for (size_t k = 0; k < np; k++)
        tb << (k ? ", " : "") << bytecounts[k];
    tb << "], \"offsets\": [";
    for (size_t k = 0; k < np; k++)
        tb << (k ? ", " : "") << offsets[k];
    tb << "]}";
    // ---- file: rank 0 creates it, everybody writes its own patches (shamcomm::open_reset_file + write_at_large)
    int fd = -1;
    if (rank == 0) {
        fd = ::open(fname.c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0644);
        if (fd < 0)
            throw std::runtime_error("dump: cannot create " + fname);
    }
    u64 ready = 1;
    comm_allreduce_host_u64(*this, &ready, 1, 0); // the file exists before the other ranks open it
    if (rank != 0) {
        fd = ::open(fname.c_str(), O_WRONLY);
        if (fd < 0)
            throw std::runtime_error("dump: cannot open " + fname);
    }
    u64 head = 0;
    write_header(fd, rank == 0, user.str(), head);
    write_header(fd, rank == 0, pm.str(), head);
    write_header(fd, rank == 0, tb.str(), head);
    std::vector<unsigned char> blob;
    for (size_t k = 0; k < np; k++) {
        PatchD &p = patches[k];
        if (!is_local(p))
            continue;
        blob.resize(bytecounts[k]);
        size_t o = 0;
        for (auto &r : p.f.all()) {
            const u64 cnt = p.f.n;
            std::memcpy(blob.data() + o, &cnt, sizeof(cnt));
            o += sizeof(cnt);
            const size_t bytes = size_t(cnt) * r.nvar * sizeof(f64);
            if (bytes)
                SB_CUDA_CHECK(cudaMemcpyAsync(blob.data() + o, r.buf->p, bytes, cudaMemcpyDeviceToHost, s()));
            o += bytes;
        }
        SB_CUDA_CHECK(cudaStreamSynchronize(s()));
        pwrite_all(fd, blob.data(), blob.size(), head + offsets[k]);
    }
    ::fsync(fd);
    ::close(fd);
    u64 done = 1;
    comm_allreduce_host_u64(*this, &done, 1, 0);
}

void Model::load_dump(const std::string &fname) {
    SB_CUDA_CHECK(cudaSetDevice(ctx->device));
    int fd = ::open(fname.c_str(), O_RDONLY);
    if (fd < 0)
        throw std::runtime_error("dump: cannot open " + fname);
    try {
        u64 head            = 0;
        const std::string u = read_header(fd, head), pmeta = read_header(fd, head), tab = read_header(fd, head);
        JVal ju = JParser(u).parse(), jp = JParser(pmeta).parse(), jt = JParser(tab).parse();
        if (!ju.has("format") || ju.at("format").str != "shamb200-1")
            throw std::runtime_error("dump: not a shamb200-1 dump");
        unhex(ju.at("solver_config_hex").str, &cfg, sizeof(cfg));
        time           = ju.at("time").as_f64();
        dt             = ju.at("dt").as_f64();
        cfl_multiplier = ju.at("cfl_multiplier").as_f64();
        step_count     = ju.at("step_count").as_u64();
        crit_split     = jp.at("crit_patch_split").as_u64();
        crit_merge     = jp.at("crit_patch_merge").as_u64();
        scheduler_freq = u32(jp.at("scheduler_freq").as_u64());
        next_patch_id  = jp.at("next_patch_id").as_u64();
        for (int d = 0; d < 3; d++) {
            box_min[d] = jp.at("sim_box").arr.at(d).as_f64();
            box_max[d] = jp.at("sim_box").arr.at(3 + d).as_f64();
        }
        const auto &layout = jp.at("patchdata_layout").arr;
        const auto &plist  = jp.at("patchlist").arr;
        patches.clear();
        patches.resize(plist.size());
        for (size_t k = 0; k < plist.size(); k++) {
            PatchD &p = patches[k];
            p.id      = plist[k].at("id_patch").as_u64();
            for (int d = 0; d < 3; d++) {
                p.cmin[d]  = plist[k].at("coord_min").arr.at(d).as_u64();
                p.cmax[d]  = plist[k].at("coord_max").arr.at(d).as_u64();
                f64 fact   = (box_max[d] - box_min[d]) / f64(kPatchGrid);
                p.lo[d]    = f64(p.cmin[d]) * fact + box_min[d];
                p.hi[d]    = f64(p.cmax[d] + 1) * fact + box_min[d];
            }
            p.owner = int(plist[k].at("node_owner_id").as_i64() % world); // ShamrockDump.cpp:206-208
        }
        {
            auto refs = patches.empty() ? std::vector<PatchFields::Ref>{} : patches[0].f.all();
            if (layout.size() != refs.size())
                throw std::runtime_error("dump: the patch data layout of the dump is not the main layout of this library");
            for (size_t r = 0; r < refs.size(); r++)
                if (layout[r].at("field_name").str != refs[r].name || int(layout[r].at("nvar").as_u64()) != refs[r].nvar)
                    throw std::runtime_error("dump: the patch data layout of the dump is not the main layout of this library");
        }
        const auto &pids = jt.at("pids").arr;
        const auto &bcs  = jt.at("bytecounts").arr;
        const auto &offs = jt.at("offsets").arr;
        if (pids.size() != patches.size())
            throw std::runtime_error("dump: patch table and patch list disagree");
        std::vector<unsigned char> blob;
        for (size_t k = 0; k < patches.size(); k++) {
            PatchD &p = patches[k];
            if (pids[k].as_u64() != p.id)
                throw std::runtime_error("dump: patch table and patch list disagree");
            if (!is_local(p))
                continue;
            blob.resize(bcs[k].as_u64());
            if (!blob.empty())
                pread_all(fd, blob.data(), blob.size(), head + offs[k].as_u64());
            size_t o  = 0;
            bool first = true;
            for (auto &r : p.f.all()) {
                u64 cnt = 0;
                std::memcpy(&cnt, blob.data() + o, sizeof(cnt));
                o += sizeof(cnt);
                if (first) {
                    if (cnt > 0xFFFFFFF0ull)
                        throw std::runtime_error("dump: patch too large");
                    p.f.n = 0;
                    p.f.reserve(u32(cnt), s());
                    p.f.n = u32(cnt);
                    first = false;
                } else if (cnt != p.f.n) {
                    throw std::runtime_error("dump: the fields of a patch have different lengths");
                }
                const size_t bytes = size_t(cnt) * r.nvar * sizeof(f64);
                if (o + bytes > blob.size())
                    throw std::runtime_error("dump: patch blob shorter than its fields");
                if (bytes)
                    SB_CUDA_CHECK(cudaMemcpyAsync(r.buf->p, blob.data() + o, bytes, cudaMemcpyHostToDevice, s()));
                o += bytes;
            }
            SB_CUDA_CHECK(cudaStreamSynchronize(s()));
        }
    } catch (...) {
        ::close(fd);
        throw;
    }
    ::close(fd);
    refresh_boxes();
    refresh_counts();
}

} // namespace sb
