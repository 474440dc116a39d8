// ghost_plan.hpp — host-side (CUDA-free) planning of the patch decomposition and of the ghost-zone
// interfaces.  Pure functions of replicated metadata, so every rank computes the same plan and the
// NCCL send/recv pairs match without any negotiation.
//
// Reference behaviour restated (paths relative to /root/reference/src):
//   shamrock/include/shamrock/patch/Patch.hpp:63-72 + PatchCoord.hpp:135-139 (patches on the 2^21 integer
//   grid, range = [coord_min, coord_max + 1)), shamrock/src/scheduler/PatchScheduler.cpp (patch→rank; here
//   a static contiguous deal of patch ids instead of Hilbert load balancing),
//   shammodels/sph/include/shammodels/sph/SPHUtilities.hpp:74-103 (interaction radius per patch),
//   shammodels/sph/src/BasicSPHGhosts.cpp:261-509 (find_interfaces: 27 periodic images, x→y→z offset
//   loop, intersection of the sender box with the receiver box grown by its interaction radius),
//   shambase/include/shambase/DistributedDataShared.hpp:54 (multimap order: (sender id, receiver id),
//   equal keys in insertion order).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <vector>

namespace sb {

constexpr uint64_t kPatchGrid = 1ull << 21; // PatchScheduler::max_axis_patch_coord_length

struct PatchBox {
    uint64_t id = 0;
    uint64_t cmin[3], cmax[3];
    double lo[3], hi[3];
    int owner = 0;
};

struct IfaceCand {
    uint32_t sender, receiver; ///< indices in the patch list
    int32_t ioff[3];
    double offset[3];
    double cut_lo[3], cut_hi[3]; ///< sender-frame volume whose particles become ghosts of the receiver
};

/// static nx*ny*nz grid of patches (x fastest), dealt to `world` ranks in contiguous id blocks
inline std::vector<PatchBox> plan_patch_grid(
    const double bmin[3], const double bmax[3], uint32_t nx, uint32_t ny, uint32_t nz, int world) {
    auto pow2 = [](uint32_t v) { return v && !(v & (v - 1)); };
    if (!pow2(nx) || !pow2(ny) || !pow2(nz))
        throw std::invalid_argument("the patch grid must be made of powers of two");
    if (world < 1)
        throw std::invalid_argument("invalid world size");
    uint32_t nn[3] = {nx, ny, nz};
    uint32_t np    = nx * ny * nz;
    std::vector<PatchBox> out(np);
    for (uint32_t z = 0; z < nz; z++)
        for (uint32_t y = 0; y < ny; y++)
            for (uint32_t x = 0; x < nx; x++) {
                uint32_t k  = x + nx * (y + ny * z);
                PatchBox &p = out[k];
                p.id        = k;
                uint32_t c[3] = {x, y, z};
                for (int d = 0; d < 3; d++) {
                    uint64_t sz = kPatchGrid / nn[d];
                    p.cmin[d]   = sz * c[d];
                    p.cmax[d]   = sz * (c[d] + 1) - 1;
                    // CoordRangeTransform<u64_3,f64_3> "multiply" mode: obj = f64(pc) * fact + bmin
                    double fact = (bmax[d] - bmin[d]) / double(kPatchGrid);
                    p.lo[d]     = double(p.cmin[d]) * fact + bmin[d];
                    p.hi[d]     = double(p.cmax[d] + 1) * fact + bmin[d];
                }
                p.owner = int((uint64_t(k) * uint64_t(world)) / np);
            }
    return out;
}

/// candidate interfaces in exchange order.  interactR[k] = max(h)·htol·Rkern of patch k, pcount[k] its
/// particle count (patches with 0 particles neither send nor receive).
inline std::vector<IfaceCand> plan_interfaces(
    const std::vector<PatchBox> &patches, const double box_min[3], const double box_max[3], bool periodic,
    const std::vector<double> &interactR, const std::vector<uint32_t> &pcount) {
    const size_t np = patches.size();
    double bsize[3] = {box_max[0] - box_min[0], box_max[1] - box_min[1], box_max[2] - box_min[2]};
    int rep         = periodic ? 1 : 0;
    std::vector<IfaceCand> cand;
    // pruning that does not change the result (hundreds of patches: the 27 np² loop must stay cheap): only the
    // patches that hold objects take part, and a shifted sender that does not even touch the hull of the grown
    // receivers (most periodic images of most patches) is dropped before the loop over the receivers
    std::vector<size_t> live;
    for (size_t k = 0; k < np; k++)
        if (pcount[k])
            live.push_back(k);
    double hull_lo[3], hull_hi[3];
    for (int d = 0; d < 3; d++) {
        hull_lo[d] = std::numeric_limits<double>::infinity();
        hull_hi[d] = -std::numeric_limits<double>::infinity();
    }
    for (size_t k : live)
        for (int d = 0; d < 3; d++) {
            hull_lo[d] = std::fmin(hull_lo[d], patches[k].lo[d] - interactR[k]);
            hull_hi[d] = std::fmax(hull_hi[d], patches[k].hi[d] + interactR[k]);
        }
    for (int32_t xoff = -rep; xoff <= rep; xoff++)
        for (int32_t yoff = -rep; yoff <= rep; yoff++)
            for (int32_t zoff = -rep; zoff <= rep; zoff++) {
                double off[3] = {xoff * bsize[0], yoff * bsize[1], zoff * bsize[2]};
                for (size_t sd : live) {
                    const PatchBox &S = patches[sd];
                    bool reach        = true; // does the shifted sender touch the hull of the grown receivers?
                    for (int d = 0; d < 3; d++)
                        reach = reach && (S.hi[d] + off[d] >= hull_lo[d]) && (S.lo[d] + off[d] <= hull_hi[d]);
                    if (!reach)
                        continue;
                    for (size_t rc : live) {
                        if (rc == sd && xoff == 0 && yoff == 0 && zoff == 0)
                            continue;
                        const PatchBox &R = patches[rc];
                        double Rr         = interactR[rc];
                        bool ok           = true;
                        IfaceCand itf;
                        for (int d = 0; d < 3; d++) {
                            double elo = R.lo[d] - Rr, ehi = R.hi[d] + Rr;
                            double so_lo = S.lo[d] + off[d], so_hi = S.hi[d] + off[d];
                            double ilo = std::fmax(elo, so_lo), ihi = std::fmin(ehi, so_hi);
                            if (!(ihi >= ilo))
                                ok = false;
                            double moff   = -off[d];
                            itf.cut_lo[d] = std::fmax(S.lo[d], elo + moff);
                            itf.cut_hi[d] = std::fmin(S.hi[d], ehi + moff);
                            itf.offset[d] = off[d];
                        }
                        if (!ok)
                            continue;
                        itf.sender   = uint32_t(sd);
                        itf.receiver = uint32_t(rc);
                        itf.ioff[0]  = xoff;
                        itf.ioff[1]  = yoff;
                        itf.ioff[2]  = zoff;
                        cand.push_back(itf);
                    }
                }
            }
    std::stable_sort(cand.begin(), cand.end(), [&](const IfaceCand &a, const IfaceCand &b) {
        if (patches[a.sender].id != patches[b.sender].id)
            return patches[a.sender].id < patches[b.sender].id;
        return patches[a.receiver].id < patches[b.receiver].id;
    });
    return cand;
}

} // namespace sb
