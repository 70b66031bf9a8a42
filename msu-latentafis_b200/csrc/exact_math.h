// Bit-exact float math shared by the CUDA kernels and their host-side unit tests.
//
// The reference evaluates the line angle between two minutiae with libm's atan2f
// (matching/matcher.cpp:1516, :1524 — `atan2(float,float)` resolves to the float overload) and
// then takes hard decisions on it (<= PI/6).  CUDA's atan2f is not bit-identical to glibc's, so the
// device path carries its own implementation of the algorithm glibc 2.39 uses for atan2f/atanf on
// x86-64 (the Sun fdlibm single-precision kernels, sysdeps/ieee754/flt-32/e_atan2f.c and
// s_atanf.c): pure fp32 add/sub/mul/div with no fused operations, which IEEE-754 makes
// reproducible on any conforming machine.  tests/test_exact_math.py checks it against libm on the
// whole integer-difference domain the matcher can produce and on random inputs.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define LAFIS_HD __host__ __device__ __forceinline__
#else
#define LAFIS_HD static inline
#endif

namespace lafis {

// Unfused fp32 primitives.  On the device the intrinsics forbid FMA contraction; host builds of
// this header must use -ffp-contract=off (the baseline x86-64 target has no FMA anyway).
LAFIS_HD float f_add(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
LAFIS_HD float f_sub(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
LAFIS_HD float f_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
LAFIS_HD float f_div(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
LAFIS_HD uint32_t f_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
LAFIS_HD float f_from_bits(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

// atanf for finite, non-NaN x (the matcher never produces anything else here).
LAFIS_HD float atanf_fdlibm(float x) {
    const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f,
                aT3 = -1.1111110449e-01f, aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f,
                aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f, aT8 = 4.9768779427e-02f,
                aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
    const uint32_t hx = f_bits(x);
    const uint32_t ix = hx & 0x7fffffffu;
    const bool neg = (hx >> 31) != 0;
    if (ix >= 0x4c000000u) {  // |x| >= 2^25
        const float r = f_add(atanhi[3], atanlo[3]);
        return neg ? -r : r;
    }
    int id;
    if (ix < 0x3ee00000u) {            // |x| < 0.4375
        if (ix < 0x31000000u) return x;  // |x| < 2^-29
        id = -1;
    } else {
        x = f_from_bits(ix);  // fabsf
        if (ix < 0x3f980000u) {      // |x| < 1.1875
            if (ix < 0x3f300000u) {  // 7/16 <= |x| < 11/16
                id = 0;
                x = f_div(f_sub(f_mul(2.0f, x), 1.0f), f_add(2.0f, x));
            } else {  // 11/16 <= |x| < 19/16
                id = 1;
                x = f_div(f_sub(x, 1.0f), f_add(x, 1.0f));
            }
        } else {
            if (ix < 0x401c0000u) {  // |x| < 2.4375
                id = 2;
                x = f_div(f_sub(x, 1.5f), f_add(1.0f, f_mul(1.5f, x)));
            } else {  // 2.4375 <= |x| < 2^25
                id = 3;
                x = f_div(-1.0f, x);
            }
        }
    }
    const float z = f_mul(x, x);
    const float w = f_mul(z, z);
    // break sum from i=0 to 10 aT[i]z**(i+1) into odd and even poly
    float s1 = f_mul(w, aT10);
    s1 = f_mul(w, f_add(aT8, s1));
    s1 = f_mul(w, f_add(aT6, s1));
    s1 = f_mul(w, f_add(aT4, s1));
    s1 = f_mul(w, f_add(aT2, s1));
    s1 = f_mul(z, f_add(aT0, s1));
    float s2 = f_mul(w, aT9);
    s2 = f_mul(w, f_add(aT7, s2));
    s2 = f_mul(w, f_add(aT5, s2));
    s2 = f_mul(w, f_add(aT3, s2));
    s2 = f_mul(w, f_add(aT1, s2));
    if (id < 0) return f_sub(x, f_mul(x, f_add(s1, s2)));
    const float r = f_sub(atanhi[id], f_sub(f_sub(f_mul(x, f_add(s1, s2)), atanlo[id]), x));
    return neg ? -r : r;
}

// atan2f(y, x) for finite, non-NaN arguments.
LAFIS_HD float atan2f_fdlibm(float y, float x) {
    const float pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    const uint32_t hx = f_bits(x), hy = f_bits(y);
    const uint32_t ix = hx & 0x7fffffffu, iy = hy & 0x7fffffffu;
    if (hx == 0x3f800000u) return atanf_fdlibm(y);  // x == 1.0
    const int m = (int)((hy >> 31) & 1u) | (int)((hx >> 30) & 2u);  // 2*sign(x) + sign(y)
    if (iy == 0) {  // y == 0
        switch (m) {
            case 0:
            case 1: return y;
            case 2: return pi;
            default: return -pi;
        }
    }
    if (ix == 0) return (hy >> 31) ? -pi_o_2 : pi_o_2;  // x == 0
    const int k = ((int)iy - (int)ix) >> 23;
    float z;
    if (k > 60) z = f_add(pi_o_2, f_mul(0.5f, pi_lo));
    else if ((hx >> 31) && k < -60) z = 0.0f;
    else z = atanf_fdlibm(f_from_bits(f_bits(f_div(y, x)) & 0x7fffffffu));
    switch (m) {
        case 0: return z;
        case 1: return f_from_bits(f_bits(z) ^ 0x80000000u);
        case 2: return f_sub(pi, f_sub(z, pi_lo));
        default: return f_sub(f_sub(z, pi_lo), pi);
    }
}

}  // namespace lafis
