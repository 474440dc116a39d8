// load_balance.hpp — host-side (CUDA-free) patch → rank assignment along the Hilbert curve.
//
// Reference behaviour restated (paths relative to /root/reference/src):
//   shammath/include/shammath/sfc/hilbert.hpp:36-72, :85-87 (Skilling's transpose algorithm on 21 bits per
//   axis, the transposed words interleaved x-first), shamrock/src/scheduler/HilbertLoadBalance.cpp:46-75
//   (one tile per patch: Hilbert code of its coord_min, load = load_value),
//   shamrock/include/shamrock/scheduler/loadbalance/LoadBalanceStrategy.hpp:62-125 (parallel sweep: tiles in
//   curve order, exclusive running load / (last running load / world) → owner, clamped), :136-205 (round
//   robin: the same with unit loads), :274-312 (the one with the smaller maximum rank load wins, round robin
//   favoured by a factor 0.95).
// Pure functions of replicated metadata: every rank computes the same table.
#pragma once
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <stdexcept>
#include <vector>

namespace sb {

/// 63-bit Hilbert index of a cell of the 2^21-per-axis patch grid
inline uint64_t hilbert_index_3d(uint64_t x, uint64_t y, uint64_t z) {
    constexpr int bits = 21;
    uint64_t w[3]      = {x, y, z};
    const uint64_t top = uint64_t(1) << (bits - 1);
    // undo the excess work of the transpose form, most significant bit first
    for (uint64_t q = top; q > 1; q >>= 1) {
        const uint64_t low = q - 1;
        for (int a = 0; a < 3; a++) {
            if (w[a] & q) {
                w[0] ^= low;
            } else {
                const uint64_t swap = (w[0] ^ w[a]) & low;
                w[0] ^= swap;
                w[a] ^= swap;
            }
        }
    }
    // Gray code
    w[1] ^= w[0];
    w[2] ^= w[1];
    uint64_t flip = 0;
    for (uint64_t q = top; q > 1; q >>= 1)
        if (w[2] & q)
            flip ^= q - 1;
    for (int a = 0; a < 3; a++)
        w[a] ^= flip;
    // interleave: bit b of w[a] lands at 3 b + (2 - a)
    auto spread = [](uint64_t v) {
        v &= 0x1fffff;
        v = (v | v << 32) & 0x1f00000000ffffull;
        v = (v | v << 16) & 0x1f0000ff0000ffull;
        v = (v | v << 8) & 0x100f00f00f00f00full;
        v = (v | v << 4) & 0x10c30c30c30c30c3ull;
        v = (v | v << 2) & 0x1249249249249249ull;
        return v;
    };
    return (spread(w[0]) << 2) + (spread(w[1]) << 1) + spread(w[2]);
}

/// owners from a sweep along the curve.  unit == true: every tile weighs one (round robin)
inline std::vector<int32_t> lb_sweep(
    const std::vector<uint64_t> &order, const std::vector<uint64_t> &load, int world, bool unit) {
    const size_t n = order.size();
    std::vector<size_t> by_curve(n);
    std::iota(by_curve.begin(), by_curve.end(), size_t(0));
    std::stable_sort(by_curve.begin(), by_curve.end(), [&](size_t a, size_t b) { return order[a] < order[b]; });
    std::vector<uint64_t> before(n); // load of the tiles ahead on the curve
    uint64_t run = 0;
    for (size_t k = 0; k < n; k++) {
        before[k] = run;
        run += unit ? 1 : load[by_curve[k]];
    }
    std::vector<int32_t> owner(n, 0);
    if (n == 0)
        return owner;
    const double per_rank = double(before[n - 1]) / world; // the reference divides the LAST running value
    for (size_t k = 0; k < n; k++) {
        int32_t o = 0;
        if (per_rank != 0) {
            o = int32_t(double(before[k]) / per_rank);
            o = std::min(std::max(o, 0), world - 1);
        }
        owner[by_curve[k]] = o;
    }
    return owner;
}

inline uint64_t lb_max_rank_load(const std::vector<uint64_t> &load, const std::vector<int32_t> &owner, int world) {
    std::vector<uint64_t> per(world, 0);
    for (size_t i = 0; i < load.size(); i++)
        per[owner[i]] += load[i];
    return *std::max_element(per.begin(), per.end());
}

/// shamrock::scheduler::load_balance.  *strategy (optional): 0 parallel sweep, 1 round robin
inline std::vector<int32_t> load_balance(
    const std::vector<uint64_t> &order, const std::vector<uint64_t> &load, int world, int *strategy = nullptr) {
    if (world < 1)
        throw std::invalid_argument("invalid world size");
    if (order.size() != load.size())
        throw std::invalid_argument("load balance: one load per tile");
    auto sweep = lb_sweep(order, load, world, false);
    auto robin = lb_sweep(order, load, world, true);
    const double m_sweep = double(lb_max_rank_load(load, sweep, world)) * 1.0;
    const double m_robin = double(lb_max_rank_load(load, robin, world)) * 0.95;
    const bool take_robin = m_robin < m_sweep;
    if (strategy)
        *strategy = take_robin ? 1 : 0;
    return take_robin ? robin : sweep;
}

/// HilbertLoadBalance<u64>: tiles = patches, curve position = Hilbert index of coord_min
inline std::vector<int32_t> hilbert_load_balance(
    const std::vector<uint64_t> &coord_min3, const std::vector<uint64_t> &load, int world, int *strategy = nullptr) {
    if (coord_min3.size() != 3 * load.size())
        throw std::invalid_argument("load balance: three coordinates per patch");
    std::vector<uint64_t> order(load.size());
    for (size_t i = 0; i < load.size(); i++)
        order[i] = hilbert_index_3d(coord_min3[3 * i], coord_min3[3 * i + 1], coord_min3[3 * i + 2]);
    return load_balance(order, load, world, strategy);
}

} // namespace sb
