// setup.cu — initial conditions generated ON THE DEVICE, straight into the patches.
//
// Replaces, for this path, the host-side setup of the reference (paths relative to /root/reference/src):
//   shammath/include/shammath/crystalLattice.hpp:52-290   LatticeHCP: generator (:68-80), index bounds
//                                                           (:140-163), iteration order x fastest (:244-248)
//   shammodels/sph/include/shammodels/sph/modules/setup/GeneratorLatticeHCP.hpp:39-120
//   shammodels/sph/src/modules/SPHSetup.cpp:112-267        apply_setup: every generated object goes to the
//                                                           patch that owns its position, generation order kept
//   shammodels/sph/src/modules/setup/GeneratorMCDisc.cpp   Monte-Carlo disc, one random draw per object
//   shammodels/sph/include/shammodels/sph/Model.hpp:669-785 set_value_in_a_box / set_value_in_sphere /
//                                                           add_kernel_value / get_sum (host loops there)
// Every rank generates only the objects of its own patches: a 64 Mi - 512 Mi particle setup costs a few
// milliseconds of kernels per patch instead of a minute of host numpy and PCIe.  The lattice is bit-identical to
// the host restatement (shamrock_b200/lattice.py, itself pinned on the reference's script geometry): same
// expressions, no FMA contraction (-fmad=false), same order inside every patch.
#include "solver.cuh"
#include "sphkern.cuh"
#include <cmath>
#include <cstring>

namespace sb {

namespace {

struct IdxBox {
    i64 lo[3], n[3]; ///< first lattice index and count per axis
};

/// LatticeHCP::get_box_index_bounds (crystalLattice.hpp:140-163): i32 truncation, -1 / +1 margins
IdxBox hcp_index_box(f64 dr, const f64 bmin[3], const f64 bmax[3]) {
    const f64 sc[3] = {2.0, std::sqrt(3.0), 2 * std::sqrt(6.0) / 3};
    IdxBox b;
    for (int d = 0; d < 3; d++) {
        f64 cmin = (bmin[d] / sc[d]) / dr, cmax = (bmax[d] / sc[d]) / dr;
        i64 imin = i64(i32(cmin)) - 1, imax = i64(i32(cmax)) + 1;
        b.lo[d]  = imin;
        b.n[d]   = imax > imin ? imax - imin : 0;
    }
    return b;
}

/// LatticeHCP::generator (crystalLattice.hpp:68-80)
__device__ __forceinline__ void hcp_point(f64 dr, i64 i, i64 j, i64 k, f64 &x, f64 &y, f64 &z) {
    const i64 ajk = (j + k) < 0 ? -(j + k) : (j + k);
    const i64 ak  = k < 0 ? -k : k;
    x = dr * f64(2 * i + (ajk % 2));
    y = dr * (sqrt(3.) * (f64(j) + (1. / 3.) * f64(ak % 2)));
    z = dr * (((2 * sqrt(6.)) * f64(k)) / 3);
}

struct LatticeArgs {
    f64 dr;
    f64 gen_lo[3], gen_hi[3]; ///< generator box: lower <= r < upper
    f64 pat_lo[3], pat_hi[3]; ///< patch box
    IdxBox ib;                ///< index sub-box scanned for this patch (a superset of its lattice points)
    u64 first;                ///< first flat index of this launch (x fastest)
    u32 count;                ///< flat indices in this launch
    // flared-disc filter (add_disc_lattice): keep r_in < R < r_out, |z| < zcut R; Keplerian speed sqrt(GM / R)
    int disc = 0;
    f64 r_in = 0, r_out = 0, zcut = 0, GM = 0, h_disc = 0;
};

__device__ __forceinline__ bool lattice_point(const LatticeArgs &a, u32 t, f64 &x, f64 &y, f64 &z) {
    const u64 f = a.first + t;
    const i64 i = a.ib.lo[0] + i64(f % u64(a.ib.n[0]));
    const u64 r = f / u64(a.ib.n[0]);
    const i64 j = a.ib.lo[1] + i64(r % u64(a.ib.n[1]));
    const i64 k = a.ib.lo[2] + i64(r / u64(a.ib.n[1]));
    hcp_point(a.dr, i, j, k, x, y, z);
    if (a.disc) {
        const f64 R = sqrt(x * x + y * y);
        if (!(R > a.r_in && R < a.r_out && fabs(z) < a.zcut * R))
            return false;
    }
    return a.gen_lo[0] <= x && x < a.gen_hi[0] && a.gen_lo[1] <= y && y < a.gen_hi[1] && a.gen_lo[2] <= z
           && z < a.gen_hi[2] && a.pat_lo[0] <= x && x < a.pat_hi[0] && a.pat_lo[1] <= y && y < a.pat_hi[1]
           && a.pat_lo[2] <= z && z < a.pat_hi[2];
}

__global__ void __launch_bounds__(256) lattice_flag_kernel(LatticeArgs a, u8 *__restrict__ flag) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.count)
        return;
    f64 x, y, z;
    flag[t] = lattice_point(a, t, x, y, z) ? 1 : 0;
}
__global__ void __launch_bounds__(256) lattice_scatter_kernel(
    LatticeArgs a, const u8 *__restrict__ flag, const u32 *__restrict__ pos, u32 base, f64 *__restrict__ xyz,
    f64 *__restrict__ hpart, f64 *__restrict__ vxyz) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.count || !flag[t])
        return;
    f64 x, y, z;
    lattice_point(a, t, x, y, z);
    const u64 o    = u64(base) + pos[t];
    xyz[3 * o]     = x;
    xyz[3 * o + 1] = y;
    xyz[3 * o + 2] = z;
    hpart[o]       = a.disc ? a.h_disc : a.dr; // GeneratorLatticeHCP: hpart = dr
    if (a.disc) {
        const f64 R  = sqrt(x * x + y * y);
        const f64 vk = sqrt(a.GM / R);
        vxyz[3 * o]     = -vk * (y / R);
        vxyz[3 * o + 1] = vk * (x / R);
        vxyz[3 * o + 2] = 0;
    }
}

__global__ void __launch_bounds__(256) set_in_box_kernel(
    u32 n, const f64 *__restrict__ xyz, f64 *__restrict__ f, int nvar, int ivar, f64 val, f64 l0, f64 l1, f64 l2, f64 h0,
    f64 h1, f64 h2) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    f64 x = xyz[3 * u64(i)], y = xyz[3 * u64(i) + 1], z = xyz[3 * u64(i) + 2];
    if (l0 <= x && x < h0 && l1 <= y && y < h1 && l2 <= z && z < h2) // BBAA::is_coord_in_range
        f[u64(i) * nvar + ivar] = val;
}
__global__ void __launch_bounds__(256) set_in_sphere_kernel(
    u32 n, const f64 *__restrict__ xyz, f64 *__restrict__ f, f64 val, f64 c0, f64 c1, f64 c2, f64 r2) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    f64 x = xyz[3 * u64(i)] - c0, y = xyz[3 * u64(i) + 1] - c1, z = xyz[3 * u64(i) + 2] - c2;
    if (x * x + y * y + z * z < r2)
        f[i] = val;
}
template<class K>
__global__ void __launch_bounds__(256) add_kernel_value_kernel(
    u32 n, const f64 *__restrict__ xyz, f64 *__restrict__ f, f64 val, f64 c0, f64 c1, f64 c2, f64 h_ker) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    f64 x = xyz[3 * u64(i)] - c0, y = xyz[3 * u64(i) + 1] - c1, z = xyz[3 * u64(i) + 2] - c2;
    f64 r = sqrt(x * x + y * y + z * z);
    f[i] += val * Kern<K>::W_3d(r, h_ker);
}
/// sums of the nvar components of a field: block partial sums, then atomics (setup / diagnostics only)
__global__ void __launch_bounds__(256) field_sum_kernel(u32 n, const f64 *__restrict__ f, int nvar, f64 *__restrict__ out) {
    __shared__ f64 ss[8][3];
    f64 s[3] = {0, 0, 0};
    for (u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += u64(gridDim.x) * blockDim.x)
        for (int c = 0; c < nvar; c++)
            s[c] += f[i * nvar + c];
    for (int c = 0; c < nvar; c++)
        s[c] = warp_sum(s[c]);
    if ((threadIdx.x & 31) == 0)
        for (int c = 0; c < nvar; c++)
            ss[threadIdx.x >> 5][c] = s[c];
    __syncthreads();
    if (threadIdx.x < unsigned(nvar)) {
        f64 t = 0;
        for (int w = 0; w < 8; w++)
            t += ss[w][threadIdx.x];
        atomicAdd(out + threadIdx.x, t);
    }
}

// ---- Monte-Carlo disc ---------------------------------------------------------------------------------
/// counter-based draws: object `idx` of the disc owns the stream hash(seed, idx, draw number) — the same object
/// gets the same draws whichever rank or launch generates it ("1 part = 1 random draw", GeneratorMCDisc.cpp:28)
__device__ __forceinline__ u64 splitmix64(u64 x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ f64 draw01(u64 seed, u64 idx, u32 k) {
    u64 v = splitmix64(splitmix64(seed ^ (idx * 0xD1342543DE82EF95ull)) + k);
    return f64(v >> 11) * (1.0 / 9007199254740992.0); // [0, 1)
}

struct DiscArgs {
    u64 seed, first;
    u32 count;
    f64 r_in, r_out, p, q, H_r_in, part_mass, disc_mass, central_mass, G, hfact;
    f64 pat_lo[3], pat_hi[3];
};
/// Σ(r) ∝ r^-p between r_in and r_out (rejection sampling of f(r) = r Σ(r), GeneratorMCDisc.cpp:30-43),
/// H(r) = H_r_in r_in (r / r_in)^(3/2 - q) ... (locally isothermal disc with c_s ∝ r^-q: H = c_s / Ω),
/// z = H · Gauss (Box-Muller), ρ = Σ / (sqrt(2π) H) exp(-z² / 2H²), h from ρ (h_rho), Keplerian velocity
__device__ __forceinline__ bool disc_point(const DiscArgs &a, u32 t, f64 (&pos)[3], f64 (&vel)[3], f64 &h) {
    const u64 idx = a.first + t;
    u32 k         = 0;
    const f64 theta = 6.283185307179586 * draw01(a.seed, idx, k++);
    const f64 g1 = draw01(a.seed, idx, k++), g2 = draw01(a.seed, idx, k++);
    const f64 gauss = sqrt(-2. * log(1. - g1)) * cos(6.283185307179586 * g2);
    auto f_func     = [&](f64 r) { return r * pow(r / a.r_in, -a.p); };
    const f64 fmx   = fmax(f_func(a.r_in), f_func(a.r_out));
    f64 r           = a.r_in;
    for (int it = 0; it < 256; it++) {
        f64 u2 = fmx * draw01(a.seed, idx, k++);
        r      = a.r_in + (a.r_out - a.r_in) * draw01(a.seed, idx, k++);
        if (u2 < f_func(r))
            break;
    }
    const f64 H = a.H_r_in * a.r_in * pow(r / a.r_in, 1.5 - a.q);
    const f64 z = H * gauss;
    pos[0] = r * cos(theta), pos[1] = r * sin(theta), pos[2] = z;
    // Σ normalisation: disc_mass = ∫ 2π r Σ0 (r / r_in)^-p dr
    const f64 e2   = 2. - a.p;
    const f64 integ = fabs(e2) > 1e-12 ? (pow(a.r_out, e2) - pow(a.r_in, e2)) / (e2 * pow(a.r_in, -a.p)) : log(a.r_out / a.r_in) * a.r_in * a.r_in;
    const f64 sigma0 = a.disc_mass / (6.283185307179586 * integ);
    const f64 sigma  = sigma0 * pow(r / a.r_in, -a.p);
    const f64 rho    = sigma / (2.5066282746310002 * H) * exp(-z * z / (2 * H * H));
    h                = a.hfact * cbrt(a.part_mass / rho); // h_rho (math/density.hpp)
    const f64 vk     = sqrt(a.G * a.central_mass / r);
    vel[0] = -vk * sin(theta), vel[1] = vk * cos(theta), vel[2] = 0;
    return a.pat_lo[0] <= pos[0] && pos[0] < a.pat_hi[0] && a.pat_lo[1] <= pos[1] && pos[1] < a.pat_hi[1]
           && a.pat_lo[2] <= pos[2] && pos[2] < a.pat_hi[2];
}
__global__ void __launch_bounds__(256) disc_flag_kernel(DiscArgs a, u8 *__restrict__ flag) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.count)
        return;
    f64 p[3], v[3], h;
    flag[t] = disc_point(a, t, p, v, h) ? 1 : 0;
}
__global__ void __launch_bounds__(256) disc_scatter_kernel(
    DiscArgs a, const u8 *__restrict__ flag, const u32 *__restrict__ pos, u32 base, f64 *__restrict__ xyz,
    f64 *__restrict__ vxyz, f64 *__restrict__ hpart) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.count || !flag[t])
        return;
    f64 p[3], v[3], h;
    disc_point(a, t, p, v, h);
    const u64 o = u64(base) + pos[t];
    for (int c = 0; c < 3; c++) {
        xyz[3 * o + c]  = p[c];
        vxyz[3 * o + c] = v[c];
    }
    hpart[o] = h;
}

constexpr u32 CHUNK = 1u << 26; ///< flat indices per launch (flags + scan positions: 5 bytes each)

} // namespace

/// zero the main-layout fields of objects [from, to) of a patch (PatchDataLayer::fields_raz)
static void zero_tail(cudaStream_t s, PatchFields &f, u32 from, u32 to) {
    for (auto &r : f.all())
        if (to > from)
            SB_CUDA_CHECK(cudaMemsetAsync(r.buf->p + size_t(from) * r.nvar, 0, size_t(to - from) * r.nvar * sizeof(f64), s));
}

u64 Model::add_lattice_hcp(f64 dr, const f64 bmin[3], const f64 bmax[3]) {
    return add_lattice_impl(dr, bmin, bmax, nullptr);
}
/// HCP lattice cut to a flared disc (r_in < R < r_out, |z| < zcut R) on Keplerian orbits: a regular (relaxed)
/// stand-in for the Monte-Carlo disc whose Poisson clumps leave objects with a non-converging h iteration
u64 Model::add_disc_lattice(f64 dr, f64 r_in, f64 r_out, f64 zcut) {
    const f64 zmax    = zcut * r_out;
    const f64 lo[3]   = {-r_out, -r_out, -zmax}, hi[3] = {r_out, r_out, zmax};
    const f64 hfact   = cfg.kernel == SHAMB200_KERNEL_M4 ? 1.2 : 1.0;
    const f64 disc[5] = {r_in, r_out, zcut, cfg.constant_G * (cfg.has_point_mass ? cfg.pm_mass : 1.),
                         hfact * std::cbrt(4 * std::sqrt(2.0)) * dr};
    return add_lattice_impl(dr, lo, hi, disc);
}
u64 Model::add_lattice_impl(f64 dr, const f64 bmin[3], const f64 bmax[3], const f64 *disc) {
    if (patches.empty())
        throw std::runtime_error("the box size is not set, please resize the box to the domain size");
    if (!(dr > 0))
        throw std::invalid_argument("lattice spacing must be positive");
    SB_CUDA_CHECK(cudaSetDevice(ctx->device));
    u64 added = 0;
    for (auto &p : patches) {
        if (!is_local(p))
            continue;
        // the lattice indices that can fall into (generator box ∩ patch box)
        f64 lo[3], hi[3];
        bool empty = false;
        for (int d = 0; d < 3; d++) {
            lo[d] = std::fmax(bmin[d], p.lo[d]);
            hi[d] = std::fmin(bmax[d], p.hi[d]);
            empty = empty || !(lo[d] < hi[d]);
        }
        if (empty)
            continue;
        LatticeArgs a;
        a.dr = dr;
        for (int d = 0; d < 3; d++) {
            a.gen_lo[d] = bmin[d], a.gen_hi[d] = bmax[d];
            a.pat_lo[d] = p.lo[d], a.pat_hi[d] = p.hi[d];
        }
        a.ib          = hcp_index_box(dr, lo, hi);
        if (disc) {
            a.disc = 1;
            a.r_in = disc[0], a.r_out = disc[1], a.zcut = disc[2], a.GM = disc[3], a.h_disc = disc[4];
        }
        const u64 tot = u64(a.ib.n[0]) * u64(a.ib.n[1]) * u64(a.ib.n[2]);
        for (u64 first = 0; first < tot; first += CHUNK) {
            a.first = first;
            a.count = u32(std::min<u64>(CHUNK, tot - first));
            flag.ensure(a.count);
            pos.ensure(a.count);
            lattice_flag_kernel<<<grid_for(a.count, 256), 256, 0, s()>>>(a, flag.p);
            SB_COUNT_LAUNCH();
            red.ensure(8 + 256);
            h_red.ensure(8 + 256);
            exclusive_scan<u8>(s(), flag.p, pos.p, a.count, scan_tmp, red.p + 5);
            SB_CUDA_CHECK(cudaMemcpyAsync(h_red.p + 5, red.p + 5, sizeof(u64), cudaMemcpyDeviceToHost, s()));
            SB_CUDA_CHECK(cudaStreamSynchronize(s()));
            const u64 kept = h_red.p[5];
            if (!kept)
                continue;
            if (u64(p.f.n) + kept > 0xFFFFFFF0ull)
                throw std::overflow_error("patch object count overflows u32: use more patches");
            const u32 n0 = p.f.n;
            p.f.reserve(u32(n0 + kept), s());
            zero_tail(s(), p.f, n0, u32(n0 + kept));
            lattice_scatter_kernel<<<grid_for(a.count, 256), 256, 0, s()>>>(
                a, flag.p, pos.p, n0, p.f.xyz.p, p.f.hpart.p, p.f.vxyz.p);
            SB_COUNT_LAUNCH();
            p.f.n = u32(n0 + kept);
            added += kept;
        }
    }
    SB_LAUNCH_CHECK();
    SB_CUDA_CHECK(cudaStreamSynchronize(s()));
    comm_allreduce_host_u64(*this, &added, 1, 0);
    return added;
}

u64 Model::add_disc_mc(u64 npart, u64 seed, f64 r_in, f64 r_out, f64 p_exp, f64 q_exp, f64 H_r_in, f64 disc_mass) {
    if (patches.empty())
        throw std::runtime_error("the box size is not set, please resize the box to the domain size");
    if (!(cfg.gpart_mass > 0))
        throw std::invalid_argument("the disc generator needs the particle mass (h from rho): set gpart_mass first");
    SB_CUDA_CHECK(cudaSetDevice(ctx->device));
    const f64 hfact = cfg.kernel == SHAMB200_KERNEL_M4 ? 1.2 : 1.0;
    u64 added       = 0;
    for (auto &p : patches) {
        if (!is_local(p))
            continue;
        DiscArgs a;
        a.seed = seed, a.r_in = r_in, a.r_out = r_out, a.p = p_exp, a.q = q_exp, a.H_r_in = H_r_in;
        a.part_mass = cfg.gpart_mass, a.disc_mass = disc_mass, a.central_mass = cfg.has_point_mass ? cfg.pm_mass : 1.;
        a.G = cfg.constant_G, a.hfact = hfact;
        for (int d = 0; d < 3; d++)
            a.pat_lo[d] = p.lo[d], a.pat_hi[d] = p.hi[d];
        for (u64 first = 0; first < npart; first += CHUNK) {
            a.first = first;
            a.count = u32(std::min<u64>(CHUNK, npart - first));
            flag.ensure(a.count);
            pos.ensure(a.count);
            disc_flag_kernel<<<grid_for(a.count, 256), 256, 0, s()>>>(a, flag.p);
            SB_COUNT_LAUNCH();
            red.ensure(8 + 256);
            h_red.ensure(8 + 256);
            exclusive_scan<u8>(s(), flag.p, pos.p, a.count, scan_tmp, red.p + 5);
            SB_CUDA_CHECK(cudaMemcpyAsync(h_red.p + 5, red.p + 5, sizeof(u64), cudaMemcpyDeviceToHost, s()));
            SB_CUDA_CHECK(cudaStreamSynchronize(s()));
            const u64 kept = h_red.p[5];
            if (!kept)
                continue;
            if (u64(p.f.n) + kept > 0xFFFFFFF0ull)
                throw std::overflow_error("patch object count overflows u32: use more patches");
            const u32 n0 = p.f.n;
            p.f.reserve(u32(n0 + kept), s());
            zero_tail(s(), p.f, n0, u32(n0 + kept));
            disc_scatter_kernel<<<grid_for(a.count, 256), 256, 0, s()>>>(
                a, flag.p, pos.p, n0, p.f.xyz.p, p.f.vxyz.p, p.f.hpart.p);
            SB_COUNT_LAUNCH();
            p.f.n = u32(n0 + kept);
            added += kept;
        }
    }
    SB_LAUNCH_CHECK();
    SB_CUDA_CHECK(cudaStreamSynchronize(s()));
    comm_allreduce_host_u64(*this, &added, 1, 0);
    return added;
}

static PatchFields::Ref field_ref(PatchFields &f, const std::string &name) {
    for (auto &r : f.all())
        if (name == r.name)
            return r;
    throw std::invalid_argument("unknown field " + name);
}

void Model::set_value_in_a_box(const std::string &name, int ivar, f64 val, const f64 bmin[3], const f64 bmax[3]) {
    SB_CUDA_CHECK(cudaSetDevice(ctx->device));
    for (auto &p : patches) {
        if (!is_local(p) || !p.f.n)
            continue;
        auto r = field_ref(p.f, name);
        if (ivar < 0 || ivar >= r.nvar)
            throw std::invalid_argument(
                "You are trying to set value in a box for field (" + name + ") with ivar >= f.get_nvar");
        set_in_box_kernel<<<grid_for(p.f.n, 256), 256, 0, s()>>>(
            p.f.n, p.f.xyz.p, r.buf->p, r.nvar, ivar, val, bmin[0], bmin[1], bmin[2], bmax[0], bmax[1], bmax[2]);
        SB_COUNT_LAUNCH();
    }
    SB_LAUNCH_CHECK();
    SB_CUDA_CHECK(cudaStreamSynchronize(s()));
}
void Model::set_value_in_sphere(const std::string &name, f64 val, const f64 center[3], f64 radius) {
    SB_CUDA_CHECK(cudaSetDevice(ctx->device));
    for (auto &p : patches) {
        if (!is_local(p) || !p.f.n)
            continue;
        auto r = field_ref(p.f, name);
        if (r.nvar != 1)
            throw std::invalid_argument("set_value_in_sphere: scalar fields only");
        set_in_sphere_kernel<<<grid_for(p.f.n, 256), 256, 0, s()>>>(
            p.f.n, p.f.xyz.p, r.buf->p, val, center[0], center[1], center[2], radius * radius);
        SB_COUNT_LAUNCH();
    }
    SB_LAUNCH_CHECK();
    SB_CUDA_CHECK(cudaStreamSynchronize(s()));
}
void Model::add_kernel_value(const std::string &name, f64 val, const f64 center[3], f64 h_ker) {
    SB_CUDA_CHECK(cudaSetDevice(ctx->device));
    for (auto &p : patches) {
        if (!is_local(p) || !p.f.n)
            continue;
        auto r = field_ref(p.f, name);
        if (r.nvar != 1)
            throw std::invalid_argument("add_kernel_value: scalar fields only");
        if (cfg.kernel == SHAMB200_KERNEL_M4)
            add_kernel_value_kernel<KM4><<<grid_for(p.f.n, 256), 256, 0, s()>>>(
                p.f.n, p.f.xyz.p, r.buf->p, val, center[0], center[1], center[2], h_ker);
        else
            add_kernel_value_kernel<KM6><<<grid_for(p.f.n, 256), 256, 0, s()>>>(
                p.f.n, p.f.xyz.p, r.buf->p, val, center[0], center[1], center[2], h_ker);
        SB_COUNT_LAUNCH();
    }
    SB_LAUNCH_CHECK();
    SB_CUDA_CHECK(cudaStreamSynchronize(s()));
}
void Model::get_sum(const std::string &name, f64 out[3]) {
    SB_CUDA_CHECK(cudaSetDevice(ctx->device));
    field_tmp.ensure(4);
    SB_CUDA_CHECK(cudaMemsetAsync(field_tmp.p, 0, 4 * sizeof(f64), s()));
    int nvar = 1;
    for (auto &p : patches) {
        if (!is_local(p) || !p.f.n)
            continue;
        auto r = field_ref(p.f, name);
        nvar   = r.nvar;
        unsigned nb = (unsigned) std::min<u64>(u64(kNumSM) * 8, (u64(p.f.n) + 255) / 256);
        field_sum_kernel<<<nb, 256, 0, s()>>>(p.f.n, r.buf->p, r.nvar, field_tmp.p);
        SB_COUNT_LAUNCH();
    }
    SB_LAUNCH_CHECK();
    f64 h[3] = {0, 0, 0};
    SB_CUDA_CHECK(cudaMemcpyAsync(h, field_tmp.p, 3 * sizeof(f64), cudaMemcpyDeviceToHost, s()));
    SB_CUDA_CHECK(cudaStreamSynchronize(s()));
    comm_allreduce_host_f64(*this, h, 3, 0);
    for (int c = 0; c < 3; c++)
        out[c] = c < nvar ? h[c] : 0.;
}

u64 Model::total_part_count() {
    u64 n = 0;
    for (auto &p : patches)
        if (is_local(p))
            n += p.f.n;
    comm_allreduce_host_u64(*this, &n, 1, 0);
    return n;
}

} // namespace sb
