/*
 * evo_math.h -- transcendental functions shared by the CPU oracle and the CUDA kernels.
 *
 * glibc's exp() and CUDA's exp() are both accurate to < 1 ulp but not bit-identical to each other;
 * the FAS parity tests compare residual histories bit for bit, so both sides evaluate the same
 * operation sequence (no FMA contraction on either side): Cody-Waite reduction x = k ln2 + r,
 * |r| <= ln2/2, degree-13 Taylor polynomial in Horner form, scaling by 2^k through the exponent
 * bits.  Measured against glibc over [-30, 30]: max relative error 2.2e-16 (1 ulp).
 */
#ifndef EVO_MATH_H
#define EVO_MATH_H
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define EVO_HD __host__ __device__ __forceinline__
#else
#define EVO_HD static inline
#endif

EVO_HD double evo_exp(double x)
{
    if (x != x) return x;
    if (x > 709.78) return 1e308 * 1e308; /* +inf */
    if (x < -708.0) return 0.0;
    const double inv_ln2 = 1.4426950408889634074;
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    double kf = x * inv_ln2;
    kf = kf + (kf >= 0.0 ? 0.5 : -0.5);
    const int k = (int)kf;
    const double kd = (double)k;
    double r = x - kd * ln2_hi;
    r = r - kd * ln2_lo;
    double p = 1.0 / 6227020800.0;
    p = p * r + 1.0 / 479001600.0;
    p = p * r + 1.0 / 39916800.0;
    p = p * r + 1.0 / 3628800.0;
    p = p * r + 1.0 / 362880.0;
    p = p * r + 1.0 / 40320.0;
    p = p * r + 1.0 / 5040.0;
    p = p * r + 1.0 / 720.0;
    p = p * r + 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;
    p = p * r + 0.5;
    p = p * r + 1.0;
    p = p * r + 1.0;
    /* scale by 2^k in two steps so that the exponent field never overflows for |k| <= 1023 */
    const int k1 = k / 2, k2 = k - k1;
    uint64_t b1 = (uint64_t)(1023 + k1) << 52, b2 = (uint64_t)(1023 + k2) << 52;
    double s1, s2;
    memcpy(&s1, &b1, 8);
    memcpy(&s2, &b2, 8);
    return p * s1 * s2;
}
#endif
