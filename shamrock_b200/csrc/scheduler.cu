// scheduler.cu — the patch scheduler of the B200 path: split / merge of patches, Hilbert-curve load balancing
// and the migration of whole patches between ranks (NCCL send / recv of every field of the main layout).
//
// Reference behaviour restated (paths relative to /root/reference/src/shamrock):
//   src/scheduler/PatchScheduler.cpp:308-500          scheduler_step: split requests (load > crit_split), merge
//                                                      requests (an octet of sibling leaves with summed load
//                                                      < crit_merge), load-balance change list, apply, merge
//   include/shamrock/patch/PatchCoord.hpp:36-120       split coordinate, child c = 4 ix + 2 iy + iz
//   src/scheduler/scheduler_patch_list.cpp:109-185     child 0 keeps the parent's id and place, children 1..7
//                                                      get fresh ids at the end of the list; the merged patch
//                                                      keeps the id of child 0
//   src/scheduler/SchedulerPatchData.cpp:302-420       split_patchdata (order kept inside every child),
//                                                      merge_patchdata (children appended in child order),
//                                                      apply_change_list (whole patches move between ranks)
//   src/scheduler/HilbertLoadBalance.cpp:46-75,
//   include/shamrock/scheduler/loadbalance/LoadBalanceStrategy.hpp:62-312   (csrc/load_balance.hpp)
//   src/shammodels/sph/src/Solver.cpp:1970-1976         evolve_once runs the scheduler step first
// The patch list (ids, integer coordinates, owners, loads) is replicated on every rank and every decision is
// a pure function of it, so the ranks agree on the operations and on the matching send / recv pairs without
// negotiation; the loads are all-reduced once per scheduler step.
#include "load_balance.hpp"
#include "solver.cuh"
#include <algorithm>
#include <cstring>

namespace sb {

/// CoordRangeTransform<u64_3, f64_3> "multiply" mode (ghost_plan.hpp: plan_patch_grid)
static void box_from_coords(PatchBox &p, const f64 bmin[3], const f64 bmax[3]) {
    for (int d = 0; d < 3; d++) {
        f64 fact = (bmax[d] - bmin[d]) / f64(kPatchGrid);
        p.lo[d]  = f64(p.cmin[d]) * fact + bmin[d];
        p.hi[d]  = f64(p.cmax[d] + 1) * fact + bmin[d];
    }
}

void Model::refresh_boxes() {
    const size_t np = patches.size();
    std::vector<f64> hb(np * 6);
    for (size_t k = 0; k < np; k++)
        for (int d = 0; d < 3; d++) {
            hb[6 * k + d]     = patches[k].lo[d];
            hb[6 * k + 3 + d] = patches[k].hi[d];
        }
    d_boxes.ensure(hb.size() + 6);
    SB_CUDA_CHECK(cudaMemcpyAsync(d_boxes.p, hb.data(), hb.size() * sizeof(f64), cudaMemcpyHostToDevice, s()));
    SB_CUDA_CHECK(cudaStreamSynchronize(s()));
}

/// PatchScheduler::split_patches for one patch (every rank updates the list, the owner splits the data)
void Model::split_patch(u32 ip) {
    SB_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (ip >= patches.size())
        throw std::invalid_argument("split_patch: no such patch");
    PatchD &par = patches[ip];
    u64 sp[3];
    for (int d = 0; d < 3; d++) {
        if (par.cmax[d] == par.cmin[d])
            throw std::invalid_argument("split_patch: the patch cannot be split any further");
        sp[d] = ((par.cmax[d] - par.cmin[d]) + 1) / 2 - 1 + par.cmin[d];
    }
    std::vector<PatchD> ch(8);
    for (int c = 0; c < 8; c++) {
        const int up[3] = {(c >> 2) & 1, (c >> 1) & 1, c & 1};
        ch[c].id        = c == 0 ? par.id : next_patch_id++;
        ch[c].owner     = par.owner;
        for (int d = 0; d < 3; d++) {
            ch[c].cmin[d] = up[d] ? sp[d] + 1 : par.cmin[d];
            ch[c].cmax[d] = up[d] ? par.cmax[d] : sp[d];
        }
        box_from_coords(ch[c], box_min, box_max);
    }
    if (is_local(par) && par.f.n) {
        const u32 n = par.f.n;
        std::vector<f64> hb(8 * 6);
        for (int c = 0; c < 8; c++)
            for (int d = 0; d < 3; d++) {
                hb[6 * c + d]     = ch[c].lo[d];
                hb[6 * c + 3 + d] = ch[c].hi[d];
            }
        DevBuf<u32> child, ids;
        DevBuf<f64> boxes8;
        boxes8.ensure(48);
        SB_CUDA_CHECK(cudaMemcpyAsync(boxes8.p, hb.data(), 48 * sizeof(f64), cudaMemcpyHostToDevice, s()));
        SB_CUDA_CHECK(cudaStreamSynchronize(s()));
        child.ensure(n);
        ids.ensure(n);
        flag.ensure(n);
        pos.ensure(n);
        red.ensure(8 + 256);
        h_red.ensure(8 + 256);
        patch_owner(s(), n, par.f.xyz.p, 8, boxes8.p, 0xFFFFFFFFu, flag.p, child.p);
        u64 placed = 0;
        for (int c = 0; c < 8; c++) {
            flag_equal(s(), n, child.p, u32(c), flag.p);
            exclusive_scan<u8>(s(), flag.p, pos.p, n, scan_tmp, red.p + 5);
            SB_CUDA_CHECK(cudaMemcpyAsync(h_red.p + 5, red.p + 5, sizeof(u64), cudaMemcpyDeviceToHost, s()));
            SB_CUDA_CHECK(cudaStreamSynchronize(s()));
            const u32 cnt = u32(h_red.p[5]);
            placed += cnt;
            if (!cnt)
                continue;
            scatter_ids(s(), n, flag.p, pos.p, ids.p);
            ch[c].f.reserve(cnt, s());
            auto src = par.f.all();
            auto dst = ch[c].f.all();
            for (size_t r = 0; r < src.size(); r++)
                gather_field(s(), cnt, src[r].nvar, ids.p, src[r].buf->p, dst[r].buf->p);
            ch[c].f.n = cnt;
        }
        SB_CUDA_CHECK(cudaStreamSynchronize(s()));
        if (placed != n)
            throw std::runtime_error("split_patchdata: an object is outside of the patch that is split");
    }
    patches[ip] = std::move(ch[0]);
    for (int c = 1; c < 8; c++)
        patches.push_back(std::move(ch[c]));
    refresh_boxes();
}

/// index of the patch with exactly these integer coordinates, or -1
static int find_patch(const std::vector<PatchD> &patches, const u64 cmin[3], const u64 cmax[3]) {
    for (size_t k = 0; k < patches.size(); k++) {
        bool ok = true;
        for (int d = 0; d < 3; d++)
            ok = ok && patches[k].cmin[d] == cmin[d] && patches[k].cmax[d] == cmax[d];
        if (ok)
            return int(k);
    }
    return -1;
}

/// the eight siblings whose child 0 is patch ip0 (list indices, child order), empty if the octet is incomplete
std::vector<u32> Model::sibling_octet(u32 ip0) const {
    const PatchD &p0 = patches.at(ip0);
    u64 ext[3];
    for (int d = 0; d < 3; d++) {
        ext[d] = p0.cmax[d] - p0.cmin[d] + 1;
        if ((p0.cmin[d] / ext[d]) % 2 != 0) // child 0 sits on an even cell of its level on every axis
            return {};
        if (2 * ext[d] > kPatchGrid)
            return {};
    }
    std::vector<u32> sib(8);
    for (int c = 0; c < 8; c++) {
        const int up[3] = {(c >> 2) & 1, (c >> 1) & 1, c & 1};
        u64 cmin[3], cmax[3];
        for (int d = 0; d < 3; d++) {
            cmin[d] = p0.cmin[d] + up[d] * ext[d];
            cmax[d] = p0.cmax[d] + up[d] * ext[d];
        }
        int k = find_patch(patches, cmin, cmax);
        if (k < 0)
            return {};
        sib[c] = u32(k);
    }
    return sib;
}

/// PatchScheduler::merge_patches for the octet whose child 0 is patch ip0: the siblings are brought to the
/// owner of child 0 first (the reference's pack_node_index makes the load balancer do that), then appended
void Model::merge_patches(u32 ip0) {
    SB_CUDA_CHECK(cudaSetDevice(ctx->device));
    std::vector<u32> sib = sibling_octet(ip0);
    if (sib.empty())
        throw std::invalid_argument("merge_patches: the octet of siblings is not complete");
    refresh_counts();
    for (int c = 1; c < 8; c++)
        if (patches[sib[c]].owner != patches[ip0].owner)
            migrate_patch(sib[c], patches[ip0].owner);
    PatchD &p0 = patches[ip0];
    if (is_local(p0)) {
        u64 total = p0.f.n;
        for (int c = 1; c < 8; c++)
            total += patches[sib[c]].f.n;
        if (total > 0xFFFFFFF0ull)
            throw std::overflow_error("merge_patches: the merged patch holds more than 2^32 objects");
        p0.f.reserve(u32(total), s());
        for (int c = 1; c < 8; c++) {
            PatchD &q = patches[sib[c]];
            if (!q.f.n)
                continue;
            auto src = q.f.all();
            auto dst = p0.f.all();
            for (size_t r = 0; r < src.size(); r++)
                SB_CUDA_CHECK(cudaMemcpyAsync(
                    dst[r].buf->p + size_t(p0.f.n) * dst[r].nvar, src[r].buf->p,
                    size_t(q.f.n) * src[r].nvar * sizeof(f64), cudaMemcpyDeviceToDevice, s()));
            p0.f.n += q.f.n;
        }
        SB_CUDA_CHECK(cudaStreamSynchronize(s()));
    }
    u64 ext[3];
    for (int d = 0; d < 3; d++)
        ext[d] = p0.cmax[d] - p0.cmin[d] + 1;
    for (int d = 0; d < 3; d++)
        p0.cmax[d] = p0.cmin[d] + 2 * ext[d] - 1;
    box_from_coords(p0, box_min, box_max);
    std::vector<u32> dead(sib.begin() + 1, sib.end());
    std::sort(dead.begin(), dead.end());
    for (size_t q = dead.size(); q-- > 0;)
        patches.erase(patches.begin() + dead[q]);
    refresh_boxes();
}

/// replicated object count of every patch (one all-reduce); counts[k] for k in list order
void Model::refresh_counts() {
    const size_t np = patches.size();
    patch_counts.assign(np, 0);
    for (size_t k = 0; k < np; k++)
        if (is_local(patches[k]))
            patch_counts[k] = patches[k].f.n;
    comm_allreduce_host_u64(*this, patch_counts.data(), np, 0);
}

/// SchedulerPatchData::apply_change_list for one patch: all fields travel as NCCL messages, device to device
void Model::migrate_patch(u32 ip, int new_owner) {
    SB_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (ip >= patches.size())
        throw std::invalid_argument("migrate_patch: no such patch");
    if (new_owner < 0 || new_owner >= world)
        throw std::invalid_argument("migrate_patch: owner outside the world");
    PatchD &p = patches[ip];
    const int old_owner = p.owner;
    if (old_owner == new_owner)
        return;
    if (patch_counts.size() != patches.size())
        refresh_counts();
    const u32 n = u32(patch_counts[ip]);
    if (rank == new_owner) {
        p.f.n = 0;
        p.f.reserve(n, s());
    }
    if (n && (rank == old_owner || rank == new_owner)) {
        comm_group_start(*this);
        for (auto &r : p.f.all()) {
            const size_t bytes = size_t(n) * r.nvar * sizeof(f64);
            if (rank == old_owner)
                comm_send(*this, r.buf->p, bytes, new_owner);
            else
                comm_recv(*this, r.buf->p, bytes, old_owner);
        }
        comm_group_end(*this);
        comm_wait(*this);
        SB_CUDA_CHECK(cudaStreamSynchronize(s()));
    }
    if (rank == new_owner)
        p.f.n = n;
    if (rank == old_owner) { // the data left: give the memory back to the pool
        p.f.n = 0;
        for (auto &r : p.f.all())
            r.buf->release();
        p.st = PatchStep{};
    }
    p.owner = new_owner;
}

/// PatchScheduler::scheduler_step(do_split_merge, do_load_balancing)
void Model::scheduler_step(bool do_split_merge, bool do_load_balancing) {
    SB_CUDA_CHECK(cudaSetDevice(ctx->device));
    refresh_counts();
    sched_log = SchedLog{};
    if (do_split_merge && crit_split > 0) {
        // split requests: leaves above crit_split (PatchTree::get_split_request); a patch is split once per step
        const size_t np0 = patches.size();
        std::vector<u32> rq;
        for (size_t k = 0; k < np0; k++)
            if (patch_counts[k] > crit_split && patches[k].cmax[0] > patches[k].cmin[0])
                rq.push_back(u32(k));
        for (u32 k : rq)
            split_patch(k);
        sched_log.splits = u32(rq.size());
        if (!rq.empty())
            refresh_counts();
        // merge requests: octets of sibling leaves whose summed load is below crit_merge
        // (PatchTree::get_merge_request); patches created by a split of this step are left alone
        if (crit_merge > 0) {
            bool merged = true;
            while (merged) {
                merged = false;
                for (u32 k = 0; k < patches.size(); k++) {
                    std::vector<u32> sib = sibling_octet(k);
                    if (sib.empty())
                        continue;
                    u64 tot = 0;
                    for (u32 q : sib)
                        tot += patch_counts[q];
                    if (tot >= crit_merge)
                        continue;
                    merge_patches(k);
                    refresh_counts();
                    sched_log.merges++;
                    merged = true;
                    break;
                }
            }
        }
    }
    if (do_load_balancing && world > 1) {
        const size_t np = patches.size();
        std::vector<u64> cmin3(np * 3), load(np);
        for (size_t k = 0; k < np; k++) {
            for (int d = 0; d < 3; d++)
                cmin3[3 * k + d] = patches[k].cmin[d];
            load[k] = patch_counts[k]; // ComputeLoadBalanceValue.cpp:23-30: the object count
        }
        int strategy = 0;
        std::vector<int32_t> owner = hilbert_load_balance(cmin3, load, world, &strategy);
        for (size_t k = 0; k < np; k++)
            if (owner[k] != patches[k].owner) {
                migrate_patch(u32(k), owner[k]);
                sched_log.moves++;
                sched_log.moved_objects += patch_counts[k];
            }
    }
    u64 mx = 0, tot = 0;
    std::vector<u64> per_rank(size_t(world), 0);
    for (size_t k = 0; k < patches.size(); k++)
        per_rank[size_t(patches[k].owner)] += patch_counts[k];
    for (u64 v : per_rank) {
        mx = std::max(mx, v);
        tot += v;
    }
    sched_log.max_rank_load  = mx;
    sched_log.mean_rank_load = world ? f64(tot) / world : 0;
    sched_log.npatch         = u32(patches.size());
}

} // namespace sb
