// -------------------------------------------------------------------------------------------
// shamrock_oracle.hpp — CPU restatement of the Shamrock SPH-timestep hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may build, link or call it.  The product path
// (shamrock_b200/csrc) never includes this header and has no CPU fallback.
//
// Parity status: PINNED.  The tree chain (Morton → bitonic sort → leaf compression → Karras →
// AABB/field → traversal order) and the smoothing-length iteration are checked against every
// golden vector the reference's own unit tests hold for this path (tests/golden/
// reference_goldens.json, extracted by tests/golden/extract_goldens.py; see
// tests/test_oracle_golden.py).  The force / CD10 operators have no unit-level golden in the
// reference; they are restated expression by expression from the cited lines.
//
// The reference itself (C++20 + SYCL 2020 + MPI) cannot be compiled in this image (no SYCL
// compiler, no MPI), so there is no oracle/_ref build; see DESIGN.md.
//
// Build flags that matter: -ffp-contract=off -fno-fast-math (IEEE, no FMA contraction), so the
// strict CUDA build (also no FMA contraction, same expression order) is bit-identical.
//
// All "ref:" comments cite files under /root/reference (commit 3ddd3ab4).
// -------------------------------------------------------------------------------------------
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <vector>

namespace oracle {

using u8  = uint8_t;
using u16 = uint16_t;
using u32 = uint32_t;
using u64 = uint64_t;
using i32 = int32_t;
using f64 = double;

struct vec3 {
    f64 x, y, z;
};
inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, f64 s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(f64 s, vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline vec3 operator/(vec3 a, f64 s) { return {a.x / s, a.y / s, a.z / s}; }
inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }
inline vec3 &operator+=(vec3 &a, vec3 b) {
    a = a + b;
    return a;
}
inline vec3 &operator-=(vec3 &a, vec3 b) {
    a = a - b;
    return a;
}
inline vec3 operator+(vec3 a, f64 s) { return {a.x + s, a.y + s, a.z + s}; }
inline vec3 operator-(vec3 a, f64 s) { return {a.x - s, a.y - s, a.z - s}; }
/// sycl::dot on a 3-vector: x*x + y*y + z*z, left to right (no FMA)
inline f64 dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline f64 length(vec3 a) { return std::sqrt(dot(a, a)); }
inline vec3 vmin(vec3 a, vec3 b) { return {std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)}; }
inline vec3 vmax(vec3 a, vec3 b) { return {std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)}; }

// ===========================================================================================
// Morton codes.  ref: src/shammath/include/shammath/sfc/bmi.hpp:29-50,
//                     src/shammath/include/shammath/sfc/morton.hpp:44-64,113-130,200-230
// ===========================================================================================
inline u64 expand_bits_u64_2(u64 x) {
    x &= 0x1fffffULL;
    x = (x | x << 32ULL) & 0x1f00000000ffffULL;
    x = (x | x << 16ULL) & 0x1f0000ff0000ffULL;
    x = (x | x << 8ULL) & 0x100f00f00f00f00fULL;
    x = (x | x << 4ULL) & 0x10c30c30c30c30c3ULL;
    x = (x | x << 2ULL) & 0x1249249249249249ULL;
    return x;
}
inline u32 expand_bits_u32_2(u32 x) {
    x &= 0x3ffU;
    x = (x | x << 16U) & 0x30000ffU;
    x = (x | x << 8U) & 0x300f00fU;
    x = (x | x << 4U) & 0x30c30c3U;
    x = (x | x << 2U) & 0x9249249U;
    return x;
}

template<class Tm>
struct MortonTraits;
template<>
struct MortonTraits<u32> {
    using ipos_t                    = u16;
    static constexpr u32 val_count  = 1024;
    static constexpr u32 max_val    = 1023;
    static constexpr u32 sig_bits   = 30;
    static constexpr u32 err_code   = 0xFFFFFFFFu;
    static constexpr u32 bitsize    = 32;
    static u32 icoord_to_morton(u32 x, u32 y, u32 z) {
        return expand_bits_u32_2(x) * 4 + expand_bits_u32_2(y) * 2 + expand_bits_u32_2(z);
    }
    static int clz(u32 v) { return v == 0 ? 32 : __builtin_clz(v); }
};
template<>
struct MortonTraits<u64> {
    using ipos_t                    = u32;
    static constexpr u32 val_count  = 2097152;
    static constexpr u32 max_val    = 2097151;
    static constexpr u32 sig_bits   = 63;
    static constexpr u64 err_code   = 0xFFFFFFFFFFFFFFFFull;
    static constexpr u32 bitsize    = 64;
    static u64 icoord_to_morton(u64 x, u64 y, u64 z) {
        return expand_bits_u64_2(x) * 4 + expand_bits_u64_2(y) * 2 + expand_bits_u64_2(z);
    }
    static int clz(u64 v) { return v == 0 ? 64 : __builtin_clzll(v); }
};

inline f64 clamp(f64 v, f64 lo, f64 hi) { return std::fmin(std::fmax(v, lo), hi); }

/// ref: src/shamtree/src/MortonCodeSet.cpp:61-128 (+ CoordRangeTransform.cpp:169-184 "multiply"
/// mode: fact = (bmax-bmin)/val_count; reverse_transform = convert<int>((r-bmin)/fact) + 0)
template<class Tm>
inline void morton_code_set_from_positions(
    const f64 *xyz, size_t stride_dbl, u32 cnt_obj, const f64 bmin[3], const f64 bmax[3],
    u32 morton_count, Tm *out) {
    using MT = MortonTraits<Tm>;
    using ip = typename MT::ipos_t;
    if (morton_count < cnt_obj)
        throw std::invalid_argument("MortonCodeSet: morton_count < cnt_obj");
    f64 fact[3];
    for (int d = 0; d < 3; d++)
        fact[d] = (bmax[d] - bmin[d]) / f64(MT::val_count);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t) cnt_obj; i++) {
        u32 ic[3];
        for (int d = 0; d < 3; d++) {
            f64 r  = clamp(xyz[i * stride_dbl + d], bmin[d], bmax[d]);
            ip c   = static_cast<ip>((r - bmin[d]) / fact[d]);
            ip mx  = ip(MT::max_val);
            ip zer = 0;
            c      = std::min(std::max(c, zer), mx);
            ic[d]  = c;
        }
        out[i] = MT::icoord_to_morton(ic[0], ic[1], ic[2]);
    }
    for (u32 i = cnt_obj; i < morton_count; i++)
        out[i] = MT::err_code;
}

inline u32 roundup_pow2(u32 v) {
    if (v <= 1)
        return 1; // ref roundup_pow2_clz: 0 -> 1, 1 -> 1
    if ((v & (v - 1)) == 0)
        return v;
    return 1u << (32 - __builtin_clz(v));
}

// ===========================================================================================
// Bitonic key/value sort.  ref: src/shamalgs/src/details/algorithm/bitonicSort_updated_usm.cpp
// :29-39 (swap rule `reverse ^ (a < b)`), :285-396 (schedule: for length, for inc).  The fused
// stencil kernels (B16/B8/B4/B2) evaluate exactly the compare-exchanges of the plain network,
// in an order that respects the data dependencies, so the sequential emulation is identical.
// ===========================================================================================
template<class Tk>
inline void sort_by_key_bitonic(Tk *m, u32 *id, u32 len) {
    if (len & (len - 1))
        throw std::invalid_argument("bitonic sort needs a power-of-two length");
    for (u32 length = 1; length < len; length <<= 1) {
        for (u32 inc = length; inc > 0; inc >>= 1) {
            u32 dir = length << 1;
#pragma omp parallel for schedule(static) if (len > (1u << 16))
            for (int64_t t = 0; t < (int64_t) (len >> 1); t++) {
                u32 low      = u32(t) & (inc - 1);
                u32 i        = (u32(t) << 1) - low; // insert 0 at bit `inc`
                bool reverse = ((dir & i) == 0);
                Tk a = m[i], b = m[i + inc];
                bool swap = reverse ^ (a < b);
                if (swap) {
                    m[i]       = b;
                    m[i + inc] = a;
                    u32 va     = id[i];
                    id[i]       = id[i + inc];
                    id[i + inc] = va;
                }
            }
        }
    }
}

// ===========================================================================================
// Leaf compression.  ref: src/shamtree/src/kernels/reduction_alg.cpp:63-85 (split table),
// :197-244 (iteration, NEW_BEHAVIOR: OFFSET empty), :274-295 (index map tail {M, 0}),
// :414-456 (ping-pong), :565-585 (remap).  karras_delta: src/shambackends/.../math.hpp:783-828
// ===========================================================================================
template<class Tm>
inline i32 karras_delta(i32 x, i32 y, u32 morton_length, const Tm *m) {
    // (y > morton_length - 1 || y < 0) is evaluated in unsigned arithmetic for the first term
    return ((u32(y) > morton_length - 1 || y < 0) ? -1
                                                   : int(MortonTraits<Tm>::clz(m[x] ^ m[y])));
}

template<class Tm>
inline void reduction_alg(
    const Tm *m, u32 morton_count, u32 reduction_level, std::vector<u32> &reduc_index_map,
    u32 &leaf_count) {
    std::vector<u32> s1(morton_count), s2(morton_count);
    for (u32 i = 0; i < morton_count; i++)
        s1[i] = (i > 0) ? (m[i - 1] != m[i] ? 1 : 0) : 1;

    auto iteration = [&](const std::vector<u32> &in, std::vector<u32> &out) {
#pragma omp parallel for schedule(static)
        for (int64_t ii = 0; ii < (int64_t) morton_count; ii++) {
            u32 i       = u32(ii);
            u32 before1 = i - 1;
            while (before1 <= morton_count - 1 && !in[before1])
                before1--;
            u32 before2 = before1 - 1;
            while (before2 <= morton_count - 1 && !in[before2])
                before2--;
            u32 next1 = i + 1;
            while (next1 <= morton_count - 1 && !in[next1])
                next1++;
            int delt_0  = karras_delta<Tm>(i32(i), i32(next1), morton_count, m);
            int delt_m  = karras_delta<Tm>(i32(i), i32(before1), morton_count, m);
            int delt_mm = karras_delta<Tm>(i32(before1), i32(before2), morton_count, m);
            out[i]      = (!(delt_0 < delt_m && delt_mm < delt_m) && in[i]) ? 1 : 0;
        }
    };
    for (u32 iter = 1; iter <= reduction_level; iter++) {
        if (iter % 2 == 0)
            iteration(s2, s1);
        else
            iteration(s1, s2);
    }
    const std::vector<u32> &split = (reduction_level % 2 == 0) ? s1 : s2;
    reduc_index_map.clear();
    for (u32 i = 0; i < morton_count; i++)
        if (split[i])
            reduc_index_map.push_back(i);
    leaf_count = u32(reduc_index_map.size());
    reduc_index_map.push_back(morton_count);
    reduc_index_map.push_back(0);
}

// ===========================================================================================
// Karras 2012 radix tree (+ endrange).  ref: src/shamtree/src/KarrasRadixTree.cpp:46-143
// (note the `float div` / ceil quirk at :107-121 — kept).
// ===========================================================================================
template<class Tm>
inline void karras_alg(
    const Tm *morton, u32 internal_cell_count, u32 *lchild_id, u32 *rchild_id, u8 *lchild_flag,
    u8 *rchild_flag, u32 *end_range_cell) {
    if (internal_cell_count == 0)
        return;
    const u32 morton_length = internal_cell_count + 1;
#pragma omp parallel for schedule(static)
    for (int64_t ii = 0; ii < (int64_t) internal_cell_count; ii++) {
        i32 i      = i32(ii);
        auto DELTA = [&](i32 x, i32 y) { return karras_delta<Tm>(x, y, morton_length, morton); };
        int ddelta = DELTA(i, i + 1) - DELTA(i, i - 1);
        int d      = (ddelta == 0) ? 0 : ((ddelta > 0) ? 1 : -1);
        int delta_min = DELTA(i, i - d);
        int lmax      = 2;
        while (DELTA(i, i + lmax * d) > delta_min)
            lmax *= 2;
        int l = 0;
        int t = lmax / 2;
        while (t > 0) {
            if (DELTA(i, i + (l + t) * d) > delta_min)
                l = l + t;
            t = t / 2;
        }
        u32 j             = u32(i + l * d);
        end_range_cell[i] = j;
        int delta_node    = DELTA(i, i32(j));
        int s             = 0;
        float div         = 2;
        t                 = int(std::ceil(float(l) / div));
        while (true) {
            int tmp_ = i + (s + t) * d;
            if (DELTA(i, tmp_) > delta_node)
                s = s + t;
            if (t <= 1)
                break;
            div *= 2;
            t = int(std::ceil(float(l) / div));
        }
        int gamma      = i + s * d + std::min(d, 0);
        lchild_id[i]   = u32(gamma);
        lchild_flag[i] = (std::min(i, i32(j)) == gamma) ? 1 : 0;
        rchild_id[i]   = u32(gamma + 1);
        rchild_flag[i] = (std::max(i, i32(j)) == gamma + 1) ? 1 : 0;
    }
}

// ===========================================================================================
// Compressed-leaf BVH container.  ref: src/shamtree/src/CompressedLeafBVH.cpp:32-95 and the
// output contract of SURVEY.md §3.3.
// ===========================================================================================
template<class Tm>
struct Tree {
    u32 obj_cnt      = 0; ///< M
    u32 morton_count = 0; ///< P2
    u32 leaf_count   = 0; ///< L
    u32 int_count    = 0; ///< I = L - 1
    f64 bmin[3], bmax[3];
    std::vector<Tm> sorted_morton;    ///< [P2]
    std::vector<u32> sort_index_map;  ///< [P2] map_morton_id_to_obj_id
    std::vector<u32> reduc_index_map; ///< [L+2]
    std::vector<Tm> reduced_morton;   ///< [L]
    std::vector<u32> lchild_id, rchild_id, endrange; ///< [I]
    std::vector<u8> lchild_flag, rchild_flag;        ///< [I]
    std::vector<vec3> aabb_min, aabb_max;            ///< [I+L], internal first

    u32 left_child(u32 id) const { return lchild_id[id] + int_count * u32(lchild_flag[id]); }
    u32 right_child(u32 id) const { return rchild_id[id] + int_count * u32(rchild_flag[id]); }
    bool is_leaf(u32 id) const { return id >= int_count; }

    template<class F>
    void for_each_in_leaf_cell(u32 cell_id, F &&f) const {
        u32 a = reduc_index_map[cell_id], b = reduc_index_map[cell_id + 1];
        for (u32 s = a; s < b; s++)
            f(sort_index_map[s]);
    }
};

/// Bottom-up max/min propagation.  The reference does `tree_depth` (=bitsizeof(Tmorton))
/// brute-force passes over all internal cells (KarrasRadixTreeAABB.cpp:33-74,
/// KarrasRadixTreeField.hpp:134-166); min/max are exact so a post-order evaluation gives the
/// same values as soon as the number of passes ≥ tree height (always true, height ≤ sig_bits).
template<class Tm, class T, class Comb>
inline void propagate_up(const Tree<Tm> &t, std::vector<T> &field, Comb comb) {
    if (t.int_count == 0)
        return;
    // iterative post-order from the root (node 0)
    std::vector<u32> stack;
    std::vector<u8> state(t.int_count, 0);
    stack.push_back(0);
    while (!stack.empty()) {
        u32 n = stack.back();
        u32 l = t.left_child(n), r = t.right_child(n);
        if (state[n] == 0) {
            state[n] = 1;
            if (!t.is_leaf(r))
                stack.push_back(r);
            if (!t.is_leaf(l))
                stack.push_back(l);
        } else {
            field[n] = comb(field[l], field[r]);
            stack.pop_back();
        }
    }
}

/// ref: src/shamtree/src/KarrasRadixTreeAABB.cpp:93-135
template<class Tm>
inline void compute_tree_aabb_from_positions(Tree<Tm> &t, const f64 *xyz, size_t stride_dbl) {
    u32 tot = t.int_count + t.leaf_count;
    t.aabb_min.assign(tot, vec3{0, 0, 0});
    t.aabb_max.assign(tot, vec3{0, 0, 0});
    const f64 mx = std::numeric_limits<f64>::max();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t) t.leaf_count; i++) {
        vec3 mn{mx, mx, mx}, mxv{-mx, -mx, -mx};
        t.for_each_in_leaf_cell(u32(i), [&](u32 id) {
            vec3 r{xyz[id * stride_dbl], xyz[id * stride_dbl + 1], xyz[id * stride_dbl + 2]};
            mn  = vmin(mn, r);
            mxv = vmax(mxv, r);
        });
        t.aabb_min[t.int_count + i] = mn;
        t.aabb_max[t.int_count + i] = mxv;
    }
    propagate_up(t, t.aabb_min, [](vec3 a, vec3 b) { return vmin(a, b); });
    propagate_up(t, t.aabb_max, [](vec3 a, vec3 b) { return vmax(a, b); });
}

/// ref: src/shamtree/include/shamtree/KarrasRadixTreeField.hpp:189-222
template<class Tm>
inline std::vector<f64> compute_tree_field_max_field(const Tree<Tm> &t, const f64 *field) {
    std::vector<f64> out(t.int_count + t.leaf_count);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t) t.leaf_count; i++) {
        f64 v = std::numeric_limits<f64>::lowest(); // shambase::get_min<f64>() is -max
        t.for_each_in_leaf_cell(u32(i), [&](u32 id) { v = std::fmax(v, field[id]); });
        out[t.int_count + i] = v;
    }
    propagate_up(t, out, [](f64 a, f64 b) { return std::fmax(a, b); });
    return out;
}

/// ref: src/shamtree/src/CompressedLeafBVH.cpp:32-95 (rebuild_from_positions)
template<class Tm>
inline Tree<Tm> rebuild_from_positions(
    const f64 *xyz, size_t stride_dbl, u32 obj_cnt, const f64 bmin[3], const f64 bmax[3],
    u32 compression_level, bool with_aabb = true, u32 morton_count_override = 0) {
    if (obj_cnt == 0)
        throw std::invalid_argument("obj_cnt is 0, cannot build a CompressedLeafBVH");
    Tree<Tm> t;
    t.obj_cnt      = obj_cnt;
    t.morton_count = morton_count_override ? morton_count_override : roundup_pow2(obj_cnt);
    for (int d = 0; d < 3; d++) {
        t.bmin[d] = bmin[d];
        t.bmax[d] = bmax[d];
    }
    t.sorted_morton.resize(t.morton_count);
    morton_code_set_from_positions<Tm>(
        xyz, stride_dbl, obj_cnt, bmin, bmax, t.morton_count, t.sorted_morton.data());
    t.sort_index_map.resize(t.morton_count);
    for (u32 i = 0; i < t.morton_count; i++)
        t.sort_index_map[i] = i;
    sort_by_key_bitonic<Tm>(t.sorted_morton.data(), t.sort_index_map.data(), t.morton_count);
    reduction_alg<Tm>(
        t.sorted_morton.data(), obj_cnt, compression_level, t.reduc_index_map, t.leaf_count);
    if (t.leaf_count == 0)
        throw std::runtime_error("0 leaf tree cannot exists");
    t.reduced_morton.resize(t.leaf_count);
    for (u32 i = 0; i < t.leaf_count; i++)
        t.reduced_morton[i] = t.sorted_morton[t.reduc_index_map[i]];
    t.int_count = t.leaf_count - 1;
    t.lchild_id.resize(t.int_count);
    t.rchild_id.resize(t.int_count);
    t.lchild_flag.resize(t.int_count);
    t.rchild_flag.resize(t.int_count);
    t.endrange.resize(t.int_count);
    karras_alg<Tm>(
        t.reduced_morton.data(), t.int_count, t.lchild_id.data(), t.rchild_id.data(),
        t.lchild_flag.data(), t.rchild_flag.data(), t.endrange.data());
    if (with_aabb)
        compute_tree_aabb_from_positions(t, xyz, stride_dbl);
    return t;
}

// ===========================================================================================
// Traversal.  ref: src/shamtree/include/shamtree/KarrasTreeTraverser.hpp:71-118 (stack DFS,
// pushes right then left → left child visited first), CLBVHObjectIterator.hpp:127-139
// ===========================================================================================
template<class Tm, class Cond, class OnLeaf>
inline void rtree_for(const Tree<Tm> &t, Cond &&cond, OnLeaf &&on_leaf) {
    constexpr u32 depth = MortonTraits<Tm>::sig_bits + 1;
    u32 stack[depth];
    u32 cursor      = depth - 1;
    stack[cursor]   = 0;
    while (cursor < depth) {
        u32 cur = stack[cursor];
        cursor++;
        if (cond(cur, t.aabb_min[cur], t.aabb_max[cur])) {
            if (t.is_leaf(cur)) {
                on_leaf(cur);
            } else {
                u32 l = t.left_child(cur), r = t.right_child(cur);
                stack[cursor - 1] = r;
                cursor--;
                stack[cursor - 1] = l;
                cursor--;
            }
        }
    }
}

/// ref: src/shamtree/include/shamtree/kernels/geometry_utils.hpp:126-135
inline bool cella_neigh_b(vec3 amin, vec3 amax, vec3 bmin, vec3 bmax) {
    return (std::fmax(amin.x, bmin.x) <= std::fmin(amax.x, bmax.x))
           && (std::fmax(amin.y, bmin.y) <= std::fmin(amax.y, bmax.y))
           && (std::fmax(amin.z, bmin.z) <= std::fmin(amax.z, bmax.z));
}
/// ref: geometry_utils.hpp:73-80
inline bool is_coord_in_range_incl_max(vec3 p, vec3 mn, vec3 mx) {
    return (mn.x <= p.x) && (p.x <= mx.x) && (mn.y <= p.y) && (p.y <= mx.y) && (mn.z <= p.z)
           && (p.z <= mx.z);
}

/// Neighbour CSR.  ref: src/shamtree/include/shamtree/TreeTraversal.hpp:375-485
struct ObjectCache {
    std::vector<u32> cnt_neigh, scanned_cnt, index_neigh_map;
    u32 sum_neigh_cnt = 0;
};
inline void prepare_object_cache(ObjectCache &c) {
    size_t n = c.cnt_neigh.size();
    c.scanned_cnt.resize(n);
    u64 acc = 0;
    for (size_t i = 0; i < n; i++) {
        c.scanned_cnt[i] = u32(acc);
        acc += c.cnt_neigh[i];
    }
    if (acc > 0xFFFFFFFFull)
        throw std::overflow_error("neighbour count overflows u32 (TreeTraversal.hpp:378)");
    c.sum_neigh_cnt = u32(acc);
    c.index_neigh_map.assign(acc, 0);
}

struct xyzh_view {
    const f64 *xyz;
    size_t stride; // in doubles
    const f64 *h;
    vec3 r(u32 i) const { return {xyz[i * stride], xyz[i * stride + 1], xyz[i * stride + 2]}; }
};

/// Two-stage search.  ref: src/shammodels/sph/src/modules/NeighbourCache.cpp:223-604.
/// rint = per-node max(h)*htol (Solver.cpp:1322-1356).  Optionally returns the leaf-leaf cache
/// and the per-particle owner leaf.
template<class Tm>
inline ObjectCache neighbour_cache_2stages(
    const Tree<Tm> &t, xyzh_view P, u32 obj_cnt, const std::vector<f64> &rint_tree, f64 Rkern,
    f64 h_tolerance, ObjectCache *leaf_cache_out = nullptr,
    std::vector<u32> *leaf_owner_out = nullptr) {
    const u32 offset_leaf = t.int_count;
    const u32 L           = t.leaf_count;
    ObjectCache lc;
    lc.cnt_neigh.resize(L);
    auto leaf_pass = [&](bool fill) {
#pragma omp parallel for schedule(dynamic, 64)
        for (int64_t g = 0; g < (int64_t) L; g++) {
            f64 a_rint = rint_tree[offset_leaf + g] * Rkern;
            vec3 amin = t.aabb_min[offset_leaf + g], amax = t.aabb_max[offset_leaf + g];
            vec3 amin_ext = amin - a_rint, amax_ext = amax + a_rint;
            u32 cnt = fill ? lc.scanned_cnt[g] : 0;
            rtree_for(
                t,
                [&](u32 node, vec3 nmin, vec3 nmax) {
                    f64 r         = rint_tree[node] * Rkern;
                    vec3 ext_bmin = nmin - r, ext_bmax = nmax + r;
                    return cella_neigh_b(amin, amax, ext_bmin, ext_bmax)
                           || cella_neigh_b(amin_ext, amax_ext, nmin, nmax);
                },
                [&](u32 leaf_b) {
                    if (fill)
                        lc.index_neigh_map[cnt] = leaf_b;
                    cnt++;
                });
            if (!fill)
                lc.cnt_neigh[g] = cnt;
        }
    };
    leaf_pass(false);
    prepare_object_cache(lc);
    leaf_pass(true);

    std::vector<u32> owner(obj_cnt);
#pragma omp parallel for schedule(static)
    for (int64_t a = 0; a < (int64_t) obj_cnt; a++) {
        vec3 r_a  = P.r(u32(a));
        u32 found = 0x7fffffff;
        rtree_for(
            t,
            [&](u32, vec3 nmin, vec3 nmax) { return is_coord_in_range_incl_max(r_a, nmin, nmax); },
            [&](u32 leaf_b) { found = leaf_b - offset_leaf; });
        owner[a] = found;
    }

    const f64 Rker2 = Rkern * Rkern;
    ObjectCache pc;
    pc.cnt_neigh.resize(obj_cnt);
    auto part_pass = [&](bool fill) {
#pragma omp parallel for schedule(dynamic, 256)
        for (int64_t a = 0; a < (int64_t) obj_cnt; a++) {
            f64 rint_a = P.h[a] * h_tolerance;
            vec3 xyz_a = P.r(u32(a));
            u32 cnt    = fill ? pc.scanned_cnt[a] : 0;
            u32 own    = owner[a];
            u32 s0 = lc.scanned_cnt[own], s1 = s0 + lc.cnt_neigh[own];
            for (u32 k = s0; k < s1; k++) {
                u32 leaf_b = lc.index_neigh_map[k];
                t.for_each_in_leaf_cell(leaf_b - offset_leaf, [&](u32 id_b) {
                    vec3 dr    = xyz_a - P.r(id_b);
                    f64 rab2   = dot(dr, dr);
                    f64 rint_b = P.h[id_b] * h_tolerance;
                    bool no_interact
                        = rab2 > rint_a * rint_a * Rker2 && rab2 > rint_b * rint_b * Rker2;
                    if (!no_interact) {
                        if (fill)
                            pc.index_neigh_map[cnt] = id_b;
                        cnt++;
                    }
                });
            }
            if (!fill)
                pc.cnt_neigh[a] = cnt;
        }
    };
    part_pass(false);
    prepare_object_cache(pc);
    part_pass(true);
    if (leaf_cache_out)
        *leaf_cache_out = std::move(lc);
    if (leaf_owner_out)
        *leaf_owner_out = std::move(owner);
    return pc;
}

/// One-stage search.  ref: NeighbourCache.cpp:30-220 and
/// src/shamtree/include/shamtree/RadixTree.hpp:802-816 (sph_radix_cell_crit)
template<class Tm>
inline ObjectCache neighbour_cache_1stage(
    const Tree<Tm> &t, xyzh_view P, u32 obj_cnt, const std::vector<f64> &rint_tree, f64 Rkern,
    f64 h_tolerance) {
    const u32 offset_leaf = t.int_count;
    const f64 Rker2       = Rkern * Rkern;
    ObjectCache pc;
    pc.cnt_neigh.resize(obj_cnt);
    auto pass = [&](bool fill) {
#pragma omp parallel for schedule(dynamic, 256)
        for (int64_t a = 0; a < (int64_t) obj_cnt; a++) {
            f64 rint_a  = P.h[a] * h_tolerance;
            vec3 xyz_a  = P.r(u32(a));
            vec3 boxmin = xyz_a - rint_a * Rkern;
            vec3 boxmax = xyz_a + rint_a * Rkern;
            u32 cnt     = fill ? pc.scanned_cnt[a] : 0;
            rtree_for(
                t,
                [&](u32 node, vec3 nmin, vec3 nmax) {
                    f64 r      = rint_tree[node] * Rkern;
                    vec3 ibmin = nmin - r, ibmax = nmax + r;
                    return cella_neigh_b(boxmin, boxmax, nmin, nmax)
                           || cella_neigh_b(xyz_a, xyz_a, ibmin, ibmax);
                },
                [&](u32 leaf_b) {
                    t.for_each_in_leaf_cell(leaf_b - offset_leaf, [&](u32 id_b) {
                        vec3 dr    = xyz_a - P.r(id_b);
                        f64 rab2   = dot(dr, dr);
                        f64 rint_b = P.h[id_b] * h_tolerance;
                        bool no_interact
                            = rab2 > rint_a * rint_a * Rker2 && rab2 > rint_b * rint_b * Rker2;
                        if (!no_interact) {
                            if (fill)
                                pc.index_neigh_map[cnt] = id_b;
                            cnt++;
                        }
                    });
                });
            if (!fill)
                pc.cnt_neigh[a] = cnt;
        }
    };
    pass(false);
    prepare_object_cache(pc);
    pass(true);
    return pc;
}

// ===========================================================================================
// SPH kernels.  ref: src/shammath/include/shammath/sphkernels.hpp:29-82 (M4), :265-346 (M6),
// :2286-2343 (SPHKernelGen W_3d / dW_3d / dhW_3d)
// ===========================================================================================
constexpr f64 PI = 3.14159265358979323846264338327950288;

struct KernelM4 {
    static constexpr f64 Rkern   = 2;
    static constexpr f64 hfactd  = 1.2;
    static constexpr f64 norm_3d = 1 / PI;
    static f64 f(f64 q) {
        f64 t1 = 2 - q, t2 = 1 - q;
        t1 = t1 * t1 * t1;
        t2 = t2 * t2 * t2;
        t1 *= (1. / 4.);
        t2 *= -1;
        if (q < 1)
            return t1 + t2;
        else if (q < 2)
            return t1;
        return 0;
    }
    static f64 df(f64 q) {
        constexpr f64 div9_4 = 9. / 4., div3_4 = 3. / 4.;
        if (q < 1)
            return -3 * q + div9_4 * q * q;
        else if (q < 2)
            return -3 + 3 * q - div3_4 * q * q;
        return 0;
    }
};
struct KernelM6 {
    static constexpr f64 Rkern   = 3;
    static constexpr f64 hfactd  = 1.0;
    static constexpr f64 norm_3d = 1 / (120 * PI);
    static f64 f(f64 q) {
        f64 t1 = 3 - q, t2 = 2 - q, t3 = 1 - q;
        f64 t1_2 = t1 * t1, t2_2 = t2 * t2, t3_2 = t3 * t3;
        t1 = t1 * t1_2 * t1_2;
        t2 = t2 * t2_2 * t2_2;
        t3 = t3 * t3_2 * t3_2;
        t1 *= 1;
        t2 *= -6;
        t3 *= 15;
        if (q < 1.)
            return t1 + t2 + t3;
        else if (q < 2.)
            return t1 + t2;
        else if (q < 3.)
            return t1;
        return 0;
    }
    static f64 df(f64 q) {
        f64 t1 = 3 - q, t2 = 2 - q, t3 = 1 - q;
        f64 t1_2 = t1 * t1, t2_2 = t2 * t2, t3_2 = t3 * t3;
        t1 = t1_2 * t1_2;
        t2 = t2_2 * t2_2;
        t3 = t3_2 * t3_2;
        t1 *= (1) * (-5);
        t2 *= (-6) * (-5);
        t3 *= (15) * (-5);
        if (q < 1.)
            return t1 + t2 + t3;
        else if (q < 2.)
            return t1 + t2;
        else if (q < 3.)
            return t1;
        return 0;
    }
};
template<class K>
struct SPHKernel {
    static constexpr f64 Rkern  = K::Rkern;
    static constexpr f64 hfactd = K::hfactd;
    static f64 W_3d(f64 r, f64 h) { return K::norm_3d * K::f(r / h) / (h * h * h); }
    static f64 dW_3d(f64 r, f64 h) { return K::norm_3d * K::df(r / h) / (h * h * h * h); }
    static f64 dhW_3d(f64 r, f64 h) {
        return -(K::norm_3d) * (3 * K::f(r / h) + (r / h) * K::df(r / h)) / (h * h * h * h);
    }
};

// ref: src/shammodels/sph/include/shammodels/sph/math/density.hpp:23-41
inline f64 rho_h(f64 m, f64 h, f64 hfact) { return m * (hfact / h) * (hfact / h) * (hfact / h); }
inline f64 newton_iterate_new_h(f64 rho_ha, f64 rho_sum, f64 sumdWdh, f64 h_a) {
    f64 f_iter  = rho_sum - rho_ha;
    f64 df_iter = sumdWdh + 3 * rho_ha / h_a;
    return h_a - f_iter / df_iter;
}
// ref: src/shambackends/include/shambackends/math.hpp:838-873
inline f64 inv_sat_positive(f64 v, f64 minvsat = 1e-9, f64 satval = 0.) {
    return (v >= minvsat) ? 1. / v : satval;
}
inline f64 inv_sat_zero(f64 v, f64 satval = 0.) { return (v != 0. && v == v) ? 1. / v : satval; }

/// One Newton sweep.  ref: src/shammodels/sph/src/modules/IterateSmoothingLengthDensity.cpp:52-119
template<class K>
inline void iterate_smoothing_length_density(
    const ObjectCache &c, const f64 *xyz, size_t stride, u32 n, const f64 *h_old, f64 *h_new,
    f64 *eps, f64 gpart_mass, f64 h_evol_max, f64 h_evol_iter_max) {
    using Kern = SPHKernel<K>;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t ia = 0; ia < (int64_t) n; ia++) {
        u32 id_a               = u32(ia);
        f64 part_mass          = gpart_mass;
        f64 h_max_tot_max_evol = h_evol_max;
        f64 h_max_evol_p       = h_evol_iter_max;
        f64 h_max_evol_m       = 1 / h_evol_iter_max;
        if (eps[id_a] > 1e-6) {
            vec3 xyz_a{xyz[id_a * stride], xyz[id_a * stride + 1], xyz[id_a * stride + 2]};
            f64 h_a     = h_new[id_a];
            f64 dint    = h_a * h_a * Kern::Rkern * Kern::Rkern;
            f64 rho_sum = 0, sumdWdh = 0;
            u32 s0 = c.scanned_cnt[id_a], s1 = s0 + c.cnt_neigh[id_a];
            for (u32 k = s0; k < s1; k++) {
                u32 id_b = c.index_neigh_map[k];
                vec3 dr  = xyz_a
                          - vec3{xyz[id_b * stride], xyz[id_b * stride + 1], xyz[id_b * stride + 2]};
                f64 rab2 = dot(dr, dr);
                if (rab2 > dint)
                    continue;
                f64 rab = std::sqrt(rab2);
                rho_sum += part_mass * Kern::W_3d(rab, h_a);
                sumdWdh += part_mass * Kern::dhW_3d(rab, h_a);
            }
            f64 rho_ha = rho_h(part_mass, h_a, Kern::hfactd);
            f64 new_h  = newton_iterate_new_h(rho_ha, rho_sum, sumdWdh, h_a);
            if (new_h < h_a * h_max_evol_m)
                new_h = h_max_evol_m * h_a;
            if (new_h > h_a * h_max_evol_p)
                new_h = h_max_evol_p * h_a;
            f64 ha_0 = h_old[id_a];
            if (new_h < ha_0 * h_max_tot_max_evol) {
                h_new[id_a] = new_h;
                eps[id_a]   = std::fabs(new_h - h_a) / ha_0;
            } else {
                h_new[id_a] = ha_0 * h_max_tot_max_evol;
                eps[id_a]   = -1;
            }
        }
    }
}

/// ref: src/shammodels/sph/src/modules/ComputeOmega.cpp:36-73
template<class K>
inline void compute_omega(
    const ObjectCache &c, const f64 *xyz, size_t stride, u32 n, const f64 *hpart, f64 *omega,
    f64 part_mass) {
    using Kern = SPHKernel<K>;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t ia = 0; ia < (int64_t) n; ia++) {
        u32 id_a = u32(ia);
        vec3 xyz_a{xyz[id_a * stride], xyz[id_a * stride + 1], xyz[id_a * stride + 2]};
        f64 h_a     = hpart[id_a];
        f64 dint    = h_a * h_a * Kern::Rkern * Kern::Rkern;
        f64 rho_sum = 0, part_omega_sum = 0;
        u32 s0 = c.scanned_cnt[id_a], s1 = s0 + c.cnt_neigh[id_a];
        for (u32 k = s0; k < s1; k++) {
            u32 id_b = c.index_neigh_map[k];
            vec3 dr
                = xyz_a - vec3{xyz[id_b * stride], xyz[id_b * stride + 1], xyz[id_b * stride + 2]};
            f64 rab2 = dot(dr, dr);
            if (rab2 > dint)
                continue;
            f64 rab = std::sqrt(rab2);
            rho_sum += part_mass * Kern::W_3d(rab, h_a);
            part_omega_sum += part_mass * Kern::dhW_3d(rab, h_a);
        }
        f64 rho_ha  = rho_h(part_mass, h_a, Kern::hfactd);
        omega[id_a] = 1 + (h_a / (3 * rho_ha)) * part_omega_sum;
    }
}

} // namespace oracle
