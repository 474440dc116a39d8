// sphkern.cuh — SPH kernel shape functions (M4, M6) and small math helpers shared by the SPH loops.
// Restates shammath/include/shammath/sphkernels.hpp:29-82 (M4), :265-346 (M6), :2286-2343 (W_3d, dW_3d,
// dhW_3d), shammodels/sph/include/shammodels/sph/math/density.hpp:23-41 (rho_h) and
// shambackends/include/shambackends/math.hpp (inv_sat_positive / inv_sat_zero) in the reference's
// evaluation order.
#pragma once
#include "common.cuh"

namespace sb {

constexpr f64 PI_D = 3.14159265358979323846264338327950288;

struct KM4 {
    static constexpr f64 Rkern   = 2;
    static constexpr f64 hfactd  = 1.2;
    static constexpr f64 norm_3d = 1 / PI_D;
    __device__ static __forceinline__ f64 f(f64 q) {
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
    __device__ static __forceinline__ f64 df(f64 q) {
        constexpr f64 div9_4 = 9. / 4., div3_4 = 3. / 4.;
        if (q < 1)
            return -3 * q + div9_4 * q * q;
        else if (q < 2)
            return -3 + 3 * q - div3_4 * q * q;
        return 0;
    }
};
struct KM6 {
    static constexpr f64 Rkern   = 3;
    static constexpr f64 hfactd  = 1.0;
    static constexpr f64 norm_3d = 1 / (120 * PI_D);
    __device__ static __forceinline__ f64 f(f64 q) {
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
    __device__ static __forceinline__ f64 df(f64 q) {
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
struct Kern {
    static constexpr f64 Rkern  = K::Rkern;
    static constexpr f64 hfactd = K::hfactd;
    __device__ static __forceinline__ f64 W_3d(f64 r, f64 h) { return K::norm_3d * K::f(r / h) / (h * h * h); }
    __device__ static __forceinline__ f64 dW_3d(f64 r, f64 h) { return K::norm_3d * K::df(r / h) / (h * h * h * h); }
    __device__ static __forceinline__ f64 dhW_3d(f64 r, f64 h) {
        return -(K::norm_3d) * (3 * K::f(r / h) + (r / h) * K::df(r / h)) / (h * h * h * h);
    }
};

__device__ __forceinline__ f64 rho_h(f64 m, f64 h, f64 hfact) { return m * (hfact / h) * (hfact / h) * (hfact / h); }
__device__ __forceinline__ f64 inv_sat_positive(f64 v) { return (v >= 1e-9) ? 1. / v : 0.; }
__device__ __forceinline__ f64 inv_sat_zero(f64 v) { return (v != 0. && v == v) ? 1. / v : 0.; }


} // namespace sb
