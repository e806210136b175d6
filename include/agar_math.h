/*
 * agar_math.h — the "portable" fp64 math of the agar.io step.
 *
 * The reference calls libm through Python's math module (atan2, cos, sin, pow; src/model/cell.py:47-57,
 * 96-103,246-248; src/model/player.py:163-167).  glibc and CUDA's libdevice do not round those functions
 * identically, so a GPU port that calls them drifts from the CPU by an ulp here and there — harmless for the
 * physics, fatal for BIT-EXACT state comparison.  This header defines the transcendental part of the step
 * with nothing but IEEE-754 basic operations (+ - * / sqrt fma rint), which round identically on x86-64
 * (compiled with -ffp-contract=off) and on sm_100a (compiled with -fmad=false).  The CUDA kernels always use
 * it; the CPU oracle uses libm by default (bit-exact against the Python reference) and this header when
 * built with -DAGAR_PORTABLE_MATH (bit-exact against the GPU).  tests/ bound the distance between the two
 * oracle builds.
 *
 * Accuracy: every function is within a few ulp of the correctly rounded result on the domain the game uses
 * (masses 1..22500, radii 0.5..85, |coordinates| <= 1000); see tests/test_portable_math.py.
 */
#ifndef AGAR_MATH_H
#define AGAR_MATH_H

#include <stdint.h>
#include <string.h>
#include <math.h>

#ifdef __CUDACC__
#define AGAR_HD __host__ __device__ __forceinline__
#else
#define AGAR_HD static inline
#endif

AGAR_HD double agar_bits_to_double(uint64_t u) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, 8);
    return d;
#endif
}
AGAR_HD uint64_t agar_double_to_bits(double d) {
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    memcpy(&u, &d, 8);
    return u;
#endif
}

/* natural log of a positive normal double:  x = 2^e * m, m in [sqrt(1/2), sqrt(2));
 * log m = 2 atanh(s), s = (m-1)/(m+1), |s| <= 0.1716 -> odd series to s^21 (next term < 1.2e-19). */
AGAR_HD double agar_log(double x) {
    uint64_t u = agar_double_to_bits(x);
    int e = (int)((u >> 52) & 0x7ff) - 1023;
    uint64_t mant = (u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m = agar_bits_to_double(mant); /* [1,2) */
    if (m > 1.4142135623730951) {
        m = m * 0.5;
        e += 1;
    }
    double s = (m - 1.0) / (m + 1.0);
    double z = s * s;
    double p = 1.0 / 21.0;
    p = fma(p, z, 1.0 / 19.0);
    p = fma(p, z, 1.0 / 17.0);
    p = fma(p, z, 1.0 / 15.0);
    p = fma(p, z, 1.0 / 13.0);
    p = fma(p, z, 1.0 / 11.0);
    p = fma(p, z, 1.0 / 9.0);
    p = fma(p, z, 1.0 / 7.0);
    p = fma(p, z, 1.0 / 5.0);
    p = fma(p, z, 1.0 / 3.0);
    /* log m = 2s + 2 s^3 p */
    double lm = fma(2.0 * s * z, p, 2.0 * s);
    /* e*ln2 in two parts (hi has 32 significant bits -> e*hi exact) */
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    double de = (double)e;
    return fma(de, LN2_HI, fma(de, LN2_LO, lm));
}

/* exp for |z| < 700:  z = k ln2 + r, |r| <= 0.3466 -> Taylor to r^13 (next term / result < 5e-18). */
AGAR_HD double agar_exp(double z) {
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    double k = rint(z * 1.44269504088896338700e+00);
    double r = fma(-k, LN2_HI, z);
    r = fma(-k, LN2_LO, r);
    double p = 1.0 / 6227020800.0;
    p = fma(p, r, 1.0 / 479001600.0);
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    int ik = (int)k;
    return p * agar_bits_to_double((uint64_t)(ik + 1023) << 52);
}

/* x^y for x > 0 (the game only raises masses / radii to fixed exponents) */
AGAR_HD double agar_pow(double x, double y) { return agar_exp(y * agar_log(x)); }

/* (cos a, sin a) for a = atan2(dy, dx), without forming the angle.  atan2(0, 0) = 0 -> (1, 0). */
AGAR_HD void agar_dir(double dy, double dx, double* c, double* s) {
    double h2 = dx * dx + dy * dy;
    if (h2 == 0.0) {
        *c = 1.0;
        *s = 0.0;
        return;
    }
    double h = sqrt(h2);
    *c = dx / h;
    *s = dy / h;
}

/* Python's round(x, nd) for |x| * 10^nd < 2^51 (bot.py:16-20,449): the multiple of 10^-nd nearest to the
 * EXACT value of x, ties to the even multiple, then correctly rounded to double.  scale = 10^nd (exact). */
AGAR_HD double agar_round_dec(double x, double scale) {
    double p = x * scale;
    double e = fma(x, scale, -p); /* x*scale == p + e exactly */
    double n = rint(p);           /* ties-to-even on p */
    double d = p - n;             /* exact, |d| <= 0.5; p, n and 0.5 are multiples of ulp(p), |e| <= ulp(p)/2, */
    if (d == 0.5) {               /* so only an exact tie of p can be moved across the half by e               */
        if (e > 0.0) n += 1.0;    /* true value above the tie (rint went down to the even neighbour)           */
    } else if (d == -0.5) {
        if (e < 0.0) n -= 1.0;
    }
    return n / scale;
}

#endif /* AGAR_MATH_H */
