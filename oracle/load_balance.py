"""Oracle (TEST INFRASTRUCTURE, never imported by shamrock_b200): the reference's patch load balancing restated
in plain Python, function by function.

ref (paths relative to /root/reference/src):
  shammath/include/shammath/sfc/hilbert.hpp:36-72      details::compute_hilbert_index_3d<21>
  shammath/include/shammath/sfc/bmi.hpp                expand_bits<u64, 2>
  shamrock/include/shamrock/scheduler/loadbalance/LoadBalanceStrategy.hpp
      :62-125 lb_startegy_parallel_sweep, :136-205 lb_startegy_roundrobin, :222-258 compute_LB_metric,
      :274-312 load_balance
  shamrock/src/scheduler/HilbertLoadBalance.cpp:46-75   one tile per patch (Hilbert code of coord_min, load_value)
Pinned on the reference's own known answer (src/tests/shamrock/patch/legacy/scheduler/test_hilbert_sfc.cpp:
compute_hilbert_index_3d<21>(2^21 - 1, 0, 0) == 2^63 - 1) by tests/test_load_balance.py."""
MASK64 = (1 << 64) - 1


def expand_bits_u64_2(x):
    """two zero bits between the 21 low bits of x (bmi.hpp, expand_bits<u64, 2>)"""
    r = 0
    for b in range(21):
        if (x >> b) & 1:
            r |= 1 << (3 * b)
    return r


def compute_hilbert_index_3d(x, y, z, bits=21):
    n = 3
    X = [x, y, z]
    M = 1 << (bits - 1)
    Q = M
    while Q > 1:  # inverse undo
        P = Q - 1
        for i in range(n):
            if X[i] & Q:
                X[0] ^= P
            else:
                t = (X[0] ^ X[i]) & P
                X[0] ^= t
                X[i] ^= t
        Q >>= 1
    for i in range(1, n):  # Gray encode
        X[i] ^= X[i - 1]
    t = 0
    Q = M
    while Q > 1:
        if X[n - 1] & Q:
            t ^= Q - 1
        Q >>= 1
    for i in range(n):
        X[i] ^= t
    return ((expand_bits_u64_2(X[0]) << 2) + (expand_bits_u64_2(X[1]) << 1) + expand_bits_u64_2(X[2])) & MASK64


def _sweep(order, load, wsize, unit):
    idx = sorted(range(len(order)), key=lambda i: order[i])  # apply_ordering (distinct codes: any sort)
    acc, accum = [], 0
    for i in idx:
        acc.append(accum)
        accum += 1 if unit else load[i]
    owners = [0] * len(order)
    if not idx:
        return owners
    target = float(acc[-1]) / wsize
    for k, i in enumerate(idx):
        owners[i] = 0 if target == 0 else min(max(int(acc[k] / target), 0), wsize - 1)
    return owners


def lb_startegy_parallel_sweep(order, load, wsize):
    return _sweep(order, load, wsize, False)


def lb_startegy_roundrobin(order, load, wsize):
    return _sweep(order, load, wsize, True)


def metric_max(load, owners, wsize, weight):
    per = [0] * wsize
    for l, o in zip(load, owners):
        per[o] += l
    return float(max(per)) * weight


def load_balance(order, load, wsize):
    a = lb_startegy_parallel_sweep(order, load, wsize)
    b = lb_startegy_roundrobin(order, load, wsize)
    if metric_max(load, b, wsize, 0.95) < metric_max(load, a, wsize, 1.0):
        return b, "round robin"
    return a, "psweep"


def hilbert_load_balance(coord_min, load, wsize):
    order = [compute_hilbert_index_3d(int(c[0]), int(c[1]), int(c[2])) for c in coord_min]
    return load_balance(order, [int(l) for l in load], wsize)
