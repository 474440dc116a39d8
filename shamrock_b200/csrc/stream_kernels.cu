// stream_kernels.cu — the streaming (non neighbour-list) kernels of the SPH step: leapfrog
// predictor / corrector, periodic wrap, point-mass force, ghost selection / gather, pack building,
// order-preserving compaction of the patch fields.
//
// Reference behaviour restated (paths relative to /root/reference/src):
//   shammodels/common/include/shammodels/common/modules/ForwardEuler.hpp:62-69 (predictor nodes,
//   sequence shammodels/sph/src/Solver.cpp:390-524), shamrock/src/math/integrators.cpp:88-119
//   (corrector), :207-236 (position modulo), shammodels/common/src/modules/
//   AddForceCentralGravPotential.cpp:38-48, shammodels/sph/src/BasicSPHGhosts.cpp:526-531
//   (get_ids_where on the cut volume), shammodels/sph/include/shammodels/sph/BasicSPHGhosts.hpp:294-321
//   (append_subset_to + apply_offset), shamrock/src/patch/PatchDataField.cpp:300-331 (remove_ids).
// Compiled with -fmad=false (bit-exact with the oracle).
#include "stream_kernels.cuh"

namespace sb {

// ---- leapfrog predictor: v+=½dt a ; u+=½dt du ; x+=dt v ; v+=½dt a ; u+=½dt du -----------------
/// periodic wrap of one coordinate into [lo, hi) (BasicSPHGhosts / apply_position_boundary of the reference)
__device__ __forceinline__ f64 wrap_coord(f64 x, f64 lo, f64 hi) {
    f64 delt = hi - lo;
    f64 r    = x - lo;
    r        = fmod(r, delt);
    r += delt;
    r = fmod(r, delt);
    r += lo;
    return r;
}
struct WrapBox {
    f64 lo[3], hi[3];
};
/// WRAP: the periodic position boundary applied to the drifted position in the same pass (nothing reads the
/// positions between the drift and the wrap in such a step: Model::evolve_once)
template<bool WRAP>
__global__ void __launch_bounds__(256) predictor_kernel(
    u32 n_pos, u32 n_u, f64 dt, f64 dt_half, f64 *__restrict__ xyz, f64 *__restrict__ vxyz,
    const f64 *__restrict__ axyz, f64 *__restrict__ uint_, const f64 *__restrict__ duint, WrapBox wb) {
    u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < u64(n_pos) * 3) {
        f64 a  = axyz[i];
        f64 v  = vxyz[i];
        v      = v + dt_half * a;
        f64 x  = xyz[i] + dt * v;
        if (WRAP) {
            const int c = int(i % 3);
            x = wrap_coord(x, c == 0 ? wb.lo[0] : (c == 1 ? wb.lo[1] : wb.lo[2]),
                           c == 0 ? wb.hi[0] : (c == 1 ? wb.hi[1] : wb.hi[2]));
        }
        xyz[i] = x;
        v      = v + dt_half * a;
        vxyz[i] = v;
    }
    if (i < n_u) {
        f64 du = duint[i];
        f64 u  = uint_[i];
        u      = u + dt_half * du;
        u      = u + dt_half * du;
        uint_[i] = u;
    }
}
static WrapBox wrap_box(const f64 *bmin, const f64 *bmax) {
    WrapBox wb{{0, 0, 0}, {0, 0, 0}};
    for (int d = 0; bmin && d < 3; d++)
        wb.lo[d] = bmin[d], wb.hi[d] = bmax[d];
    return wb;
}
void leapfrog_predictor(
    cudaStream_t s, u32 n, f64 dt, f64 *xyz, f64 *vxyz, const f64 *axyz, f64 *uint_, const f64 *duint,
    const f64 *wrap_min, const f64 *wrap_max) {
    if (!n)
        return;
    if (wrap_min)
        predictor_kernel<true><<<grid_for(u64(n) * 3, 256), 256, 0, s>>>(
            n, n, dt, dt / 2, xyz, vxyz, axyz, uint_, duint, wrap_box(wrap_min, wrap_max));
    else
        predictor_kernel<false><<<grid_for(u64(n) * 3, 256), 256, 0, s>>>(
            n, n, dt, dt / 2, xyz, vxyz, axyz, uint_, duint, wrap_box(nullptr, nullptr));
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}
/// the two halves of the predictor on their own (host-resident patch data: `uint` may still be in flight
/// while the positions are already being drifted; the updates are independent element by element)
void leapfrog_predictor_pos(
    cudaStream_t s, u32 n, f64 dt, f64 *xyz, f64 *vxyz, const f64 *axyz, const f64 *wrap_min, const f64 *wrap_max) {
    if (!n)
        return;
    if (wrap_min)
        predictor_kernel<true><<<grid_for(u64(n) * 3, 256), 256, 0, s>>>(
            n, 0, dt, dt / 2, xyz, vxyz, axyz, nullptr, nullptr, wrap_box(wrap_min, wrap_max));
    else
        predictor_kernel<false><<<grid_for(u64(n) * 3, 256), 256, 0, s>>>(
            n, 0, dt, dt / 2, xyz, vxyz, axyz, nullptr, nullptr, wrap_box(nullptr, nullptr));
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}
void leapfrog_predictor_u(cudaStream_t s, u32 n, f64 dt, f64 *uint_, const f64 *duint) {
    if (!n)
        return;
    predictor_kernel<false><<<grid_for(n, 256), 256, 0, s>>>(
        0, n, dt, dt / 2, nullptr, nullptr, nullptr, uint_, duint, wrap_box(nullptr, nullptr));
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

// ---- leapfrog corrector + eps_v² max + Σ v·v (+ the sums of modules::ConservativeCheck) -------------------
/// cons (optional, 8 doubles, atomically accumulated): the quantities ConservativeCheck::check_conservation
/// (ConservativeCheck.cpp:26-190) reduces right before the corrector — Σ v (3), Σ a (3), Σ (u + v²/2),
/// Σ (v·a + du/dt), without the particle mass — from the values this kernel reads anyway.
__global__ void __launch_bounds__(256) corrector_kernel(
    u32 n, f64 hdt, f64 *__restrict__ vxyz, const f64 *__restrict__ axyz, const f64 *__restrict__ axyz_old,
    f64 *__restrict__ uint_, const f64 *__restrict__ duint, const f64 *__restrict__ duint_old, u64 *red_max,
    f64 *red_sum, f64 *cons) {
    u32 i     = blockIdx.x * blockDim.x + threadIdx.x;
    f64 epsv2 = -INFINITY, vsq = 0;
    f64 c[8]  = {0, 0, 0, 0, 0, 0, 0, 0};
    if (i < n) {
        f64 ax = axyz[3 * u64(i)], ay = axyz[3 * u64(i) + 1], az = axyz[3 * u64(i) + 2];
        f64 ix = hdt * (ax - axyz_old[3 * u64(i)]);
        f64 iy = hdt * (ay - axyz_old[3 * u64(i) + 1]);
        f64 iz = hdt * (az - axyz_old[3 * u64(i) + 2]);
        f64 v0x = vxyz[3 * u64(i)], v0y = vxyz[3 * u64(i) + 1], v0z = vxyz[3 * u64(i) + 2];
        f64 vx = v0x + ix, vy = v0y + iy, vz = v0z + iz;
        vxyz[3 * u64(i)]     = vx;
        vxyz[3 * u64(i) + 1] = vy;
        vxyz[3 * u64(i) + 2] = vz;
        epsv2    = ix * ix + iy * iy + iz * iz;
        vsq      = vx * vx + vy * vy + vz * vz;
        f64 du   = duint[i], u0 = uint_[i];
        f64 incu = hdt * (du - duint_old[i]);
        uint_[i] = u0 + incu;
        if (cons) {
            c[0] = v0x, c[1] = v0y, c[2] = v0z, c[3] = ax, c[4] = ay, c[5] = az;
            c[6] = u0 + 0.5 * (v0x * v0x + v0y * v0y + v0z * v0z);
            c[7] = (v0x * ax + v0y * ay + v0z * az) + du;
        }
    }
    __shared__ f64 smax[8], ssum[8], scons[8][8];
    f64 m = warp_max(epsv2), q = warp_sum(vsq);
    if (cons) {
#pragma unroll
        for (int k = 0; k < 8; k++)
            c[k] = warp_sum(c[k]);
    }
    if ((threadIdx.x & 31) == 0) {
        smax[threadIdx.x >> 5] = m;
        ssum[threadIdx.x >> 5] = q;
        if (cons) {
#pragma unroll
            for (int k = 0; k < 8; k++)
                scons[threadIdx.x >> 5][k] = c[k];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        f64 a = smax[0], b = ssum[0];
        for (int k = 1; k < 8; k++) {
            a = fmax(a, smax[k]);
            b += ssum[k];
        }
        atomicMax((unsigned long long *) red_max, (unsigned long long) f64_to_ordered(a));
        atomicAdd(red_sum, b);
    }
    if (cons && threadIdx.x < 8) {
        f64 t = 0;
        for (int w = 0; w < 8; w++)
            t += scons[w][threadIdx.x];
        atomicAdd(cons + threadIdx.x, t);
    }
}
void leapfrog_corrector(
    cudaStream_t s, u32 n, f64 hdt, f64 *vxyz, const f64 *axyz, const f64 *axyz_old, f64 *uint_, const f64 *duint,
    const f64 *duint_old, u64 *red_max, f64 *red_sum, f64 *cons) {
    if (!n)
        return;
    corrector_kernel<<<grid_for(n, 256), 256, 0, s>>>(n, hdt, vxyz, axyz, axyz_old, uint_, duint, duint_old, red_max, red_sum, cons);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

// ---- do consecutive ids lie next to each other? ---------------------------------------------------
/// counts the objects whose successor by id lies farther away than 8 h: a handful per thousand when the patch
/// data follows a space-filling curve (ParticleReordering: the jumps of the curve), nearly all of them otherwise
__global__ void __launch_bounds__(256) far_successor_kernel(
    u32 n, const f64 *__restrict__ xyz, const f64 *__restrict__ h, unsigned long long *__restrict__ out) {
    u32 i    = blockIdx.x * blockDim.x + threadIdx.x;
    bool far = false;
    if (i + 1 < n) {
        f64 dx = xyz[3 * u64(i) + 3] - xyz[3 * u64(i)], dy = xyz[3 * u64(i) + 4] - xyz[3 * u64(i) + 1],
            dz = xyz[3 * u64(i) + 5] - xyz[3 * u64(i) + 2];
        f64 hh = h[i];
        far    = dx * dx + dy * dy + dz * dz > 64. * hh * hh;
    }
    const u32 b = __ballot_sync(0xffffffffu, far);
    if (b && (threadIdx.x & 31) == 0)
        atomicAdd(out, (unsigned long long) __popc(b));
}
void count_far_successors(cudaStream_t s, u32 n, const f64 *xyz, const f64 *h, u64 *out) {
    SB_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(u64), s));
    if (n < 2)
        return;
    far_successor_kernel<<<grid_for(n, 256), 256, 0, s>>>(n, xyz, h, reinterpret_cast<unsigned long long *>(out));
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

/// the scalars a step reduces over the ranks, gathered into one f64 vector on the device:
/// sc[0] = Σ v² and sc[1..8] (the conservation sums, already there) are summed; sc[9] = max eps_v²,
/// sc[10] = -min dt are maximised — two back-to-back NCCL all-reduces and ONE host synchronisation
__global__ void step_scalars_kernel(const u64 *__restrict__ red, f64 *__restrict__ sc) {
    if (threadIdx.x == 0) {
        sc[0]  = __longlong_as_double((long long) red[3]);
        sc[9]  = red[2] ? ordered_to_f64(red[2]) : 0.;                       // a rank without objects: neutral
        sc[10] = red[4] != 0xFFFFFFFFFFFFFFFFull ? -ordered_to_f64(red[4]) : -INFINITY;
    }
}
void step_scalars(cudaStream_t s, const u64 *red, f64 *sc) {
    step_scalars_kernel<<<1, 32, 0, s>>>(red, sc);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

// ---- periodic wrap ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wrap_kernel(u32 n, f64 *__restrict__ xyz, f64 b0x, f64 b0y, f64 b0z, f64 b1x, f64 b1y, f64 b1z) {
    u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= u64(n) * 3)
        return;
    int c    = int(i % 3);
    f64 lo   = c == 0 ? b0x : (c == 1 ? b0y : b0z);
    f64 hi   = c == 0 ? b1x : (c == 1 ? b1y : b1z);
    xyz[i] = wrap_coord(xyz[i], lo, hi);
}
void periodic_wrap(cudaStream_t s, u32 n, f64 *xyz, const f64 bmin[3], const f64 bmax[3]) {
    if (!n)
        return;
    wrap_kernel<<<grid_for(u64(n) * 3, 256), 256, 0, s>>>(n, xyz, bmin[0], bmin[1], bmin[2], bmax[0], bmax[1], bmax[2]);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

// ---- external force: central point mass -----------------------------------------------------------
__global__ void __launch_bounds__(256) point_mass_kernel(u32 n, const f64 *__restrict__ xyz, f64 *__restrict__ axyz_ext, f64 mGM) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    f64 x = xyz[3 * u64(i)] - 0., y = xyz[3 * u64(i) + 1] - 0., z = xyz[3 * u64(i) + 2] - 0.;
    f64 abs_ra   = sqrt(x * x + y * y + z * z);
    f64 abs_ra_3 = abs_ra * abs_ra * abs_ra;
    axyz_ext[3 * u64(i)] += (mGM * x) / abs_ra_3;
    axyz_ext[3 * u64(i) + 1] += (mGM * y) / abs_ra_3;
    axyz_ext[3 * u64(i) + 2] += (mGM * z) / abs_ra_3;
}
void ext_force_point_mass(cudaStream_t s, u32 n, const f64 *xyz, f64 *axyz_ext, f64 central_mass, f64 G) {
    if (!n)
        return;
    point_mass_kernel<<<grid_for(n, 256), 256, 0, s>>>(n, xyz, axyz_ext, -central_mass * G);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

// ---- selection flags --------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) flag_in_box_kernel(
    u32 n, const f64 *__restrict__ xyz, f64 l0, f64 l1, f64 l2, f64 h0, f64 h1, f64 h2, u8 *__restrict__ flag) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    f64 x = xyz[3 * u64(i)], y = xyz[3 * u64(i) + 1], z = xyz[3 * u64(i) + 2];
    flag[i] = ((l0 <= x) && (x < h0) && (l1 <= y) && (y < h1) && (l2 <= z) && (z < h2)) ? 1 : 0;
}
void flag_in_box(cudaStream_t s, u32 n, const f64 *xyz, const f64 lo[3], const f64 hi[3], u8 *flag) {
    if (!n)
        return;
    flag_in_box_kernel<<<grid_for(n, 256), 256, 0, s>>>(n, xyz, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], flag);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

/// mode 0: keep if |r - c|² > rad² (accretion: ExternalForces.cpp:651-655, rad = Racc)
/// mode 1: keep if !(|r - c| > rad)  (kill sphere: GetParticlesOutsideSphere.cpp:34-36)
__global__ void __launch_bounds__(256) flag_sphere_kernel(
    u32 n, const f64 *__restrict__ xyz, f64 cx, f64 cy, f64 cz, f64 rad, int mode, u8 *__restrict__ flag) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    f64 x = xyz[3 * u64(i)] - cx, y = xyz[3 * u64(i) + 1] - cy, z = xyz[3 * u64(i) + 2] - cz;
    f64 d2 = x * x + y * y + z * z;
    if (mode == 0)
        flag[i] = (d2 > rad * rad) ? 1 : 0;
    else
        flag[i] = (sqrt(d2) > rad) ? 0 : 1;
}
void flag_sphere(cudaStream_t s, u32 n, const f64 *xyz, const f64 c[3], f64 rad, int mode, u8 *flag) {
    if (!n)
        return;
    flag_sphere_kernel<<<grid_for(n, 256), 256, 0, s>>>(n, xyz, c[0], c[1], c[2], rad, mode, flag);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

/// owner patch of each particle: index of the patch whose [lo,hi) box contains it, 0xFFFFFFFF if none
__global__ void __launch_bounds__(256) patch_owner_kernel(
    u32 n, const f64 *__restrict__ xyz, u32 npatch, const f64 *__restrict__ boxes /*npatch*6*/, u32 self,
    u8 *__restrict__ stay_flag, u32 *__restrict__ owner) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    f64 x = xyz[3 * u64(i)], y = xyz[3 * u64(i) + 1], z = xyz[3 * u64(i) + 2];
    u32 own = 0xFFFFFFFFu;
    {   // nearly every object is still in its patch (patch boxes are disjoint: the order of the tests is free)
        const f64 *b = boxes + 6 * self;
        if (self < npatch && b[0] <= x && x < b[3] && b[1] <= y && y < b[4] && b[2] <= z && z < b[5])
            own = self;
    }
    for (u32 k = 0; k < npatch && own == 0xFFFFFFFFu; k++) {
        const f64 *b = boxes + 6 * k;
        if (b[0] <= x && x < b[3] && b[1] <= y && y < b[4] && b[2] <= z && z < b[5])
            own = k;
    }
    owner[i]     = own;
    stay_flag[i] = (own == self) ? 1 : 0;
}
void patch_owner(cudaStream_t s, u32 n, const f64 *xyz, u32 npatch, const f64 *d_boxes, u32 self, u8 *stay_flag, u32 *owner) {
    if (!n)
        return;
    patch_owner_kernel<<<grid_for(n, 256), 256, 0, s>>>(n, xyz, npatch, d_boxes, self, stay_flag, owner);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}
__global__ void __launch_bounds__(256) flag_eq_kernel(u32 n, const u32 *__restrict__ v, u32 val, u8 *__restrict__ flag) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        flag[i] = (v[i] == val) ? 1 : 0;
}
void flag_equal(cudaStream_t s, u32 n, const u32 *v, u32 val, u8 *flag) {
    if (!n)
        return;
    flag_eq_kernel<<<grid_for(n, 256), 256, 0, s>>>(n, v, val, flag);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

/// ids[pos[i]] = i where flag[i]  (stream compaction, ascending ids)
__global__ void __launch_bounds__(256) scatter_ids_kernel(u32 n, const u8 *__restrict__ flag, const u32 *__restrict__ pos, u32 *__restrict__ ids) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i])
        ids[pos[i]] = i;
}
void scatter_ids(cudaStream_t s, u32 n, const u8 *flag, const u32 *pos, u32 *ids) {
    if (!n)
        return;
    scatter_ids_kernel<<<grid_for(n, 256), 256, 0, s>>>(n, flag, pos, ids);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

// ---- ghost selection: all interfaces of one sender patch in two passes ---------------------------------
// replaces one get_ids_where (flag + stream compaction, BasicSPHGhosts.cpp:526-531) PER INTERFACE: pass A
// classifies every particle against up to 64 cut boxes at once (bit mask) and counts per 256-particle
// block, a scan per interface turns the block counts into offsets, pass B writes the ids — ascending
// inside every interface, as the reference's stream compaction does.
constexpr int GS_THREADS = 256;

__global__ void __launch_bounds__(GS_THREADS) ghost_mask_kernel(
    u32 n, const f64 *__restrict__ xyz, u32 nbox, const f64 *__restrict__ boxes, u64 *__restrict__ mask,
    u32 nblocks, u32 *__restrict__ block_counts) {
    __shared__ f64 sbox[64 * 6];
    __shared__ u32 scnt[64];
    for (u32 j = threadIdx.x; j < nbox * 6; j += GS_THREADS)
        sbox[j] = boxes[j];
    if (threadIdx.x < 64)
        scnt[threadIdx.x] = 0;
    __syncthreads();
    u32 i  = blockIdx.x * GS_THREADS + threadIdx.x;
    u64 mk = 0;
    if (i < n) {
        f64 x = xyz[3 * u64(i)], y = xyz[3 * u64(i) + 1], z = xyz[3 * u64(i) + 2];
        for (u32 b = 0; b < nbox; b++) {
            const f64 *q = sbox + 6 * b;
            if (q[0] <= x && x < q[3] && q[1] <= y && y < q[4] && q[2] <= z && z < q[5])
                mk |= (1ull << b);
        }
        mask[i] = mk;
    }
    u32 lo = __reduce_or_sync(0xffffffffu, u32(mk)), hi = __reduce_or_sync(0xffffffffu, u32(mk >> 32));
    u64 any = (u64(hi) << 32) | lo;
    while (any) {
        int b    = __ffsll((long long) any) - 1;
        any &= any - 1;
        u32 bal = __ballot_sync(0xffffffffu, (mk >> b) & 1ull);
        if ((threadIdx.x & 31) == 0)
            atomicAdd(&scnt[b], __popc(bal));
    }
    __syncthreads();
    if (threadIdx.x < nbox && scnt[threadIdx.x])
        block_counts[u64(threadIdx.x) * nblocks + blockIdx.x] = scnt[threadIdx.x];
}

/// one CTA per interface: exclusive scan of its block counts (in place), total -> totals[b]
__global__ void __launch_bounds__(1024) ghost_scan_kernel(u32 *block_counts, u32 nblocks, u32 *totals) {
    __shared__ u32 warp_s[32];
    __shared__ u32 carry_s;
    u32 *c = block_counts + u64(blockIdx.x) * nblocks;
    if (threadIdx.x == 0)
        carry_s = 0;
    __syncthreads();
    for (u32 base = 0; base < nblocks; base += 1024) {
        u32 i   = base + threadIdx.x;
        u32 v   = i < nblocks ? c[i] : 0u;
        u32 inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((threadIdx.x & 31) >= o)
                inc += t;
        }
        if ((threadIdx.x & 31) == 31)
            warp_s[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            u32 w = warp_s[threadIdx.x], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                u32 t = __shfl_up_sync(0xffffffffu, wi, o);
                if (threadIdx.x >= o)
                    wi += t;
            }
            warp_s[threadIdx.x] = wi - w;
        }
        __syncthreads();
        u32 excl = carry_s + warp_s[threadIdx.x >> 5] + (inc - v);
        if (i < nblocks)
            c[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023)
            carry_s = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        totals[blockIdx.x] = carry_s;
}

/// pass B: ids_pool[base[b] + offset] = i for every particle i inside box b, ascending in i
__global__ void __launch_bounds__(GS_THREADS) ghost_scatter_kernel(
    u32 n, const u64 *__restrict__ mask, u32 nbox, u32 nblocks, const u32 *__restrict__ block_offsets,
    const u64 *__restrict__ base, u32 *__restrict__ ids_pool) {
    __shared__ u32 wcnt[GS_THREADS / 32][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    u32 i  = blockIdx.x * GS_THREADS + threadIdx.x;
    u64 mk = i < n ? mask[i] : 0ull;
    u32 lo = __reduce_or_sync(0xffffffffu, u32(mk)), hi = __reduce_or_sync(0xffffffffu, u32(mk >> 32));
    u64 anyw = (u64(hi) << 32) | lo;
    if (__syncthreads_or(anyw != 0) == 0)
        return; // no ghost in this block
    for (int b = lane; b < 64; b += 32)
        wcnt[warp][b] = 0;
    __syncwarp();
    u64 any = anyw;
    while (any) {
        int b = __ffsll((long long) any) - 1;
        any &= any - 1;
        u32 bal = __ballot_sync(0xffffffffu, (mk >> b) & 1ull);
        if (lane == 0)
            wcnt[warp][b] = __popc(bal);
    }
    __syncthreads();
    any = anyw;
    while (any) {
        int b = __ffsll((long long) any) - 1;
        any &= any - 1;
        u32 bal = __ballot_sync(0xffffffffu, (mk >> b) & 1ull);
        if ((mk >> b) & 1ull) {
            u32 off = block_offsets[u64(b) * nblocks + blockIdx.x];
            for (int w = 0; w < warp; w++)
                off += wcnt[w][b];
            ids_pool[base[b] + off + __popc(bal & ((1u << lane) - 1u))] = i;
        }
    }
}

void ghost_select_count(
    cudaStream_t s, u32 n, const f64 *xyz, u32 nbox, const f64 *d_boxes, u64 *mask, u32 *block_counts, u32 *d_totals) {
    if (!n || !nbox)
        return;
    if (nbox > 64)
        throw std::invalid_argument("ghost_select: at most 64 boxes per pass");
    u32 nblocks = grid_for(n, GS_THREADS);
    SB_CUDA_CHECK(cudaMemsetAsync(block_counts, 0, size_t(nbox) * nblocks * sizeof(u32), s));
    ghost_mask_kernel<<<nblocks, GS_THREADS, 0, s>>>(n, xyz, nbox, d_boxes, mask, nblocks, block_counts);
    SB_COUNT_LAUNCH();
    ghost_scan_kernel<<<nbox, 1024, 0, s>>>(block_counts, nblocks, d_totals);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}
void ghost_select_scatter(
    cudaStream_t s, u32 n, const u64 *mask, u32 nbox, const u32 *block_offsets, const u64 *d_base, u32 *ids_pool) {
    if (!n || !nbox)
        return;
    u32 nblocks = grid_for(n, GS_THREADS);
    ghost_scatter_kernel<<<nblocks, GS_THREADS, 0, s>>>(n, mask, nbox, nblocks, block_offsets, d_base, ids_pool);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

// ---- gathers ------------------------------------------------------------------------------------------
/// dst[k*nvar + c] = src[ids[k]*nvar + c]   (append_subset_to / keep_ids)
__global__ void __launch_bounds__(256) gather_field_kernel(u32 cnt, int nvar, const u32 *__restrict__ ids, const f64 *__restrict__ src, f64 *__restrict__ dst) {
    u64 t = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= u64(cnt) * nvar)
        return;
    u32 k = u32(t / nvar);
    int c = int(t % nvar);
    dst[t] = src[u64(ids[k]) * nvar + c];
}
void gather_field(cudaStream_t s, u32 cnt, int nvar, const u32 *ids, const f64 *src, f64 *dst) {
    if (!cnt)
        return;
    gather_field_kernel<<<grid_for(u64(cnt) * nvar, 256), 256, 0, s>>>(cnt, nvar, ids, src, dst);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

__global__ void __launch_bounds__(256) rows_gather_kernel(u32 cnt, const u32 *__restrict__ ids, RowTable t, u32 src_off, u32 dst_off) {
    u64 q = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (q >= u64(cnt) * t.nf)
        return;
    const u32 k = u32(q / t.nf);
    const int f = int(q % t.nf);
    const u64 i = ids ? u64(ids[k]) : u64(src_off) + k;
    const int nv = t.nvar[f];
    const f64 *a = t.src[f] + i * nv;
    f64 *b       = t.dst[f] + (u64(dst_off) + k) * nv;
    for (int c = 0; c < nv; c++)
        b[c] = a[c];
}
void rows_gather(cudaStream_t s, u32 cnt, const u32 *ids, const RowTable &t, u32 src_off, u32 dst_off) {
    if (!cnt)
        return;
    rows_gather_kernel<<<grid_for(u64(cnt) * t.nf, 256), 256, 0, s>>>(cnt, ids, t, src_off, dst_off);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

__global__ void __launch_bounds__(256) split_ids_kernel(
    u32 n, const u8 *__restrict__ flag, const u32 *__restrict__ pos, u32 *__restrict__ set_ids, u32 *__restrict__ cleared_ids) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    if (flag[i])
        set_ids[pos[i]] = i;
    else
        cleared_ids[i - pos[i]] = i;
}
void split_ids(cudaStream_t s, u32 n, const u8 *flag, const u32 *pos, u32 *set_ids, u32 *cleared_ids) {
    if (!n)
        return;
    split_ids_kernel<<<grid_for(n, 256), 256, 0, s>>>(n, flag, pos, set_ids, cleared_ids);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}
__global__ void __launch_bounds__(1024) select_equal_kernel(
    u32 n, const u32 *__restrict__ ids, const u32 *__restrict__ key, u32 val, u32 *__restrict__ out) {
    __shared__ u32 wsum[32];
    __shared__ u32 tile_total;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 base = 0;
    for (u32 t0 = 0; t0 < n; t0 += 1024) {
        const u32 j  = t0 + threadIdx.x;
        const u32 id = j < n ? ids[j] : 0u;
        const bool f = j < n && key[id] == val;
        const u32 b  = __ballot_sync(0xffffffffu, f);
        if (lane == 0)
            wsum[warp] = __popc(b);
        __syncthreads();
        if (warp == 0) {
            const u32 v = wsum[lane];
            u32 inc     = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                u32 t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= u32(o))
                    inc += t;
            }
            wsum[lane] = inc - v;
            if (lane == 31)
                tile_total = inc;
        }
        __syncthreads();
        if (f)
            out[base + wsum[warp] + __popc(b & ((1u << lane) - 1u))] = id;
        base += tile_total;
        __syncthreads();
    }
}
void select_equal(cudaStream_t s, u32 n, const u32 *ids, const u32 *key, u32 val, u32 *out) {
    if (!n)
        return;
    select_equal_kernel<<<1, 1024, 0, s>>>(n, ids, key, val, out);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

// ---- packs ---------------------------------------------------------------------------------------------
/// A[i] = (xyz_i, h_i) for the real particles
__global__ void __launch_bounds__(256) pack_xyzh_kernel(u32 n, const f64 *__restrict__ xyz, const f64 *__restrict__ h, Pack4 *__restrict__ A) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    A[i] = Pack4{xyz[3 * u64(i)], xyz[3 * u64(i) + 1], xyz[3 * u64(i) + 2], h[i]};
}
void pack_xyzh(cudaStream_t s, u32 n, const f64 *xyz, const f64 *h, Pack4 *A) {
    if (!n)
        return;
    pack_xyzh_kernel<<<grid_for(n, 256), 256, 0, s>>>(n, xyz, h, A);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}
/// ghost positions: A_dst[k] = (xyz[ids[k]] + offset, h[ids[k]])
__global__ void __launch_bounds__(256) ghost_xyzh_kernel(
    u32 cnt, const u32 *__restrict__ ids, const f64 *__restrict__ xyz, const f64 *__restrict__ h, f64 ox, f64 oy,
    f64 oz, Pack4 *__restrict__ A_dst) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt)
        return;
    u32 id   = ids[k];
    A_dst[k] = Pack4{xyz[3 * u64(id)] + ox, xyz[3 * u64(id) + 1] + oy, xyz[3 * u64(id) + 2] + oz, h[id]};
}
void ghost_xyzh(cudaStream_t s, u32 cnt, const u32 *ids, const f64 *xyz, const f64 *h, const f64 off[3], Pack4 *A_dst) {
    if (!cnt)
        return;
    ghost_xyzh_kernel<<<grid_for(cnt, 256), 256, 0, s>>>(cnt, ids, xyz, h, off[0], off[1], off[2], A_dst);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

// ---- batched interface kernels (stream_kernels.cuh) ----------------------------------------------------
template<class Job>
__device__ __forceinline__ int batch_job_of_block(const JobBatch<Job> &b, u32 blk) {
    int lo = 0, hi = b.n - 1; // last job whose first block is <= blk
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (b.first_block[mid] <= blk)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}
__global__ void __launch_bounds__(256) ghost_xyzh_batch_kernel(const __grid_constant__ JobBatch<GhostXyzhJob> b) {
    const int j           = batch_job_of_block(b, blockIdx.x);
    const GhostXyzhJob &J = b.job[j];
    const u32 k           = (blockIdx.x - b.first_block[j]) * 256 + threadIdx.x;
    if (k >= J.count)
        return;
    if (J.ids) {
        const u32 id = J.ids[k];
        J.dst[k] = Pack4{J.xyz[3 * u64(id)] + J.ox, J.xyz[3 * u64(id) + 1] + J.oy, J.xyz[3 * u64(id) + 2] + J.oz, J.h[id]};
    } else { // the patch's own objects (pack_xyzh): no offset arithmetic, not even + 0
        J.dst[k] = Pack4{J.xyz[3 * u64(k)], J.xyz[3 * u64(k) + 1], J.xyz[3 * u64(k) + 2], J.h[k]};
    }
}
__global__ void __launch_bounds__(256) pack_fields_batch_kernel(const __grid_constant__ JobBatch<PackFieldsJob> b) {
    const int j            = batch_job_of_block(b, blockIdx.x);
    const PackFieldsJob &J = b.job[j];
    const u32 k            = (blockIdx.x - b.first_block[j]) * 256 + threadIdx.x;
    if (k >= J.count)
        return;
    const u32 id = J.ids ? J.ids[k] : k;
    const u32 o  = J.dst_map ? J.dst_map[k] : k;
    J.A[o].d = J.h[id];
    J.B[o]   = Pack4{J.vxyz[3 * u64(id)], J.vxyz[3 * u64(id) + 1], J.vxyz[3 * u64(id) + 2], J.uint_[id]};
    J.C[o].b = J.omega[id];
    if (J.axyz)
        J.D[o] = Pack4{J.axyz[3 * u64(id)], J.axyz[3 * u64(id) + 1], J.axyz[3 * u64(id) + 2], 0.};
}
__global__ void __launch_bounds__(256) pack_alpha_batch_kernel(const __grid_constant__ JobBatch<PackAlphaJob> b) {
    const int j           = batch_job_of_block(b, blockIdx.x);
    const PackAlphaJob &J = b.job[j];
    const u32 k           = (blockIdx.x - b.first_block[j]) * 256 + threadIdx.x;
    if (k >= J.count)
        return;
    const u32 src = J.ids ? J.ids[k] : k;
    Pack4 &c      = J.C[J.dst_map ? J.dst_map[k] : k];
    c.d           = J.alpha[src];
    if (J.omega)
        c.b = J.omega[src];
}
__global__ void __launch_bounds__(256) unpack_ghost_batch_kernel(const __grid_constant__ JobBatch<UnpackGhostJob> b) {
    const int j             = batch_job_of_block(b, blockIdx.x);
    const UnpackGhostJob &J = b.job[j];
    const u32 k             = (blockIdx.x - b.first_block[j]) * 256 + threadIdx.x;
    if (k >= J.count)
        return;
    const u32 o = J.dst_map ? J.dst_map[k] : k;
    J.A[o].d = J.sA[k].d;
    J.B[o]   = J.sB[k];
    J.C[o].b = J.sC[k].b;
    if (J.sD)
        J.D[o] = J.sD[k];
}
template<class Job>
static void launch_batch(cudaStream_t s, const JobBatch<Job> &b, u32 blocks);
template<>
void launch_batch<GhostXyzhJob>(cudaStream_t s, const JobBatch<GhostXyzhJob> &b, u32 blocks) {
    ghost_xyzh_batch_kernel<<<blocks, 256, 0, s>>>(b);
}
template<>
void launch_batch<PackFieldsJob>(cudaStream_t s, const JobBatch<PackFieldsJob> &b, u32 blocks) {
    pack_fields_batch_kernel<<<blocks, 256, 0, s>>>(b);
}
template<>
void launch_batch<PackAlphaJob>(cudaStream_t s, const JobBatch<PackAlphaJob> &b, u32 blocks) {
    pack_alpha_batch_kernel<<<blocks, 256, 0, s>>>(b);
}
template<>
void launch_batch<UnpackGhostJob>(cudaStream_t s, const JobBatch<UnpackGhostJob> &b, u32 blocks) {
    unpack_ghost_batch_kernel<<<blocks, 256, 0, s>>>(b);
}
template<class Job>
void Batcher<Job>::flush() {
    if (!b.n)
        return;
    b.first_block[b.n] = blocks;
    launch_batch<Job>(s, b, blocks);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
    b.n    = 0;
    blocks = 0;
}
template struct Batcher<GhostXyzhJob>;
template struct Batcher<PackFieldsJob>;
template struct Batcher<PackAlphaJob>;
template struct Batcher<UnpackGhostJob>;

/// field packs.  ids == nullptr: identity (real particles of the patch itself)
/// A.d = h ; B = (v, u) ; C.b = omega ; D = (a, 0) when `axyz` is given
/// dst_map (optional): record k goes to slot dst_map[k] (the Morton rank of merged index base + k)
__global__ void __launch_bounds__(256) pack_fields_kernel(
    u32 cnt, const u32 *__restrict__ ids, const u32 *__restrict__ dst_map, const f64 *__restrict__ h,
    const f64 *__restrict__ vxyz, const f64 *__restrict__ uint_, const f64 *__restrict__ omega,
    const f64 *__restrict__ axyz, Pack4 *__restrict__ A, Pack4 *__restrict__ B, Pack4 *__restrict__ C,
    Pack4 *__restrict__ D) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt)
        return;
    u32 id = ids ? ids[k] : k;
    u32 o  = dst_map ? dst_map[k] : k;
    A[o].d = h[id];
    B[o]   = Pack4{vxyz[3 * u64(id)], vxyz[3 * u64(id) + 1], vxyz[3 * u64(id) + 2], uint_[id]};
    C[o].b = omega[id];
    if (axyz)
        D[o] = Pack4{axyz[3 * u64(id)], axyz[3 * u64(id) + 1], axyz[3 * u64(id) + 2], 0.};
}
void pack_fields(
    cudaStream_t s, u32 cnt, const u32 *ids, const f64 *h, const f64 *vxyz, const f64 *uint_, const f64 *omega,
    const f64 *axyz, Pack4 *A, Pack4 *B, Pack4 *C, Pack4 *D, const u32 *dst_map) {
    if (!cnt)
        return;
    pack_fields_kernel<<<grid_for(cnt, 256), 256, 0, s>>>(cnt, ids, dst_map, h, vxyz, uint_, omega, axyz, A, B, C, D);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}
/// received ghost-field blocks → merged arrays: A.d = h, B = (v,u), C.b = omega, D = (a,0)
__global__ void __launch_bounds__(256) unpack_ghost_fields_kernel(
    u32 cnt, const Pack4 *__restrict__ sA, const Pack4 *__restrict__ sB, const Pack4 *__restrict__ sC,
    const Pack4 *__restrict__ sD, const u32 *__restrict__ dst_map, Pack4 *__restrict__ A, Pack4 *__restrict__ B,
    Pack4 *__restrict__ C, Pack4 *__restrict__ D) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt)
        return;
    u32 o  = dst_map ? dst_map[k] : k;
    A[o].d = sA[k].d;
    B[o]   = sB[k];
    C[o].b = sC[k].b;
    if (sD)
        D[o] = sD[k];
}
void unpack_ghost_fields(
    cudaStream_t s, u32 cnt, const Pack4 *sA, const Pack4 *sB, const Pack4 *sC, const Pack4 *sD, Pack4 *A, Pack4 *B,
    Pack4 *C, Pack4 *D, const u32 *dst_map) {
    if (!cnt)
        return;
    unpack_ghost_fields_kernel<<<grid_for(cnt, 256), 256, 0, s>>>(cnt, sA, sB, sC, sD, dst_map, A, B, C, D);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}
/// C[k].d = alpha[ids ? ids[k] : k]; with omega (fast fp mode: Ω comes out of the CD10 / MM97 operator pass
/// and travels with alpha) also C[k].b = omega[...]
__global__ void __launch_bounds__(256) pack_alpha_kernel(
    u32 cnt, const u32 *__restrict__ ids, const u32 *__restrict__ dst_map, const f64 *__restrict__ alpha,
    const f64 *__restrict__ omega, Pack4 *__restrict__ C) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt)
        return;
    const u32 src = ids ? ids[k] : k;
    Pack4 &c      = C[dst_map ? dst_map[k] : k];
    c.d           = alpha[src];
    if (omega)
        c.b = omega[src];
}
void pack_alpha(cudaStream_t s, u32 cnt, const u32 *ids, const f64 *alpha, Pack4 *C, const u32 *dst_map,
                const f64 *omega) {
    if (!cnt)
        return;
    pack_alpha_kernel<<<grid_for(cnt, 256), 256, 0, s>>>(cnt, ids, dst_map, alpha, omega, C);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}
/// out[i] = C[i].c (sound speed of the real particles → main field, Solver.cpp:3129-3161)
__global__ void __launch_bounds__(256) unpack_cs_kernel(
    u32 n, const Pack4 *__restrict__ C, const u32 *__restrict__ src_map, f64 *__restrict__ cs) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        cs[i] = C[src_map ? src_map[i] : i].c;
}
void unpack_cs(cudaStream_t s, u32 n, const Pack4 *C, f64 *cs, const u32 *src_map) {
    if (!n)
        return;
    unpack_cs_kernel<<<grid_for(n, 256), 256, 0, s>>>(n, C, src_map, cs);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}
/// generic component extraction for inspection (tests): out[i*nc + c] = P[i].comp(first + c)
__global__ void __launch_bounds__(256) unpack_comp_kernel(
    u32 n, const Pack4 *__restrict__ P, const u32 *__restrict__ src_map, int first, int nc, f64 *__restrict__ out) {
    u64 t = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= u64(n) * nc)
        return;
    u32 i = u32(t / nc);
    int c = first + int(t % nc);
    const f64 *q = reinterpret_cast<const f64 *>(P + (src_map ? src_map[i] : i));
    out[t]       = q[c];
}
void unpack_comp(cudaStream_t s, u32 n, const Pack4 *P, int first, int nc, f64 *out, const u32 *src_map) {
    if (!n)
        return;
    unpack_comp_kernel<<<grid_for(u64(n) * nc, 256), 256, 0, s>>>(n, P, src_map, first, nc, out);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

// ---- max reduction of a field (interactR_patch = max(h)·htol·Rkern) ---------------------------------------
__global__ void __launch_bounds__(256) max_reduce_kernel(u32 n, const f64 *__restrict__ v, u64 *red) {
    f64 m = -INFINITY;
    for (u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += u64(gridDim.x) * blockDim.x)
        m = fmax(m, v[i]);
    __shared__ f64 sm[8];
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0)
        sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        f64 a = sm[0];
        for (int k = 1; k < 8; k++)
            a = fmax(a, sm[k]);
        atomicMax((unsigned long long *) red, (unsigned long long) f64_to_ordered(a));
    }
}
void max_reduce(cudaStream_t s, u32 n, const f64 *v, u64 *red) {
    if (!n)
        return;
    unsigned nb = (unsigned) std::min<u64>(u64(kNumSM) * 8, (u64(n) + 255) / 256);
    max_reduce_kernel<<<nb, 256, 0, s>>>(n, v, red);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

} // namespace sb
