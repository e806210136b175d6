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
#include "agar_libm_tables.h"

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

/* x^y, BIT-IDENTICAL to the pow() CPython's math.pow calls on this image (glibc 2.39, the FMA variant its ifunc
 * selects on every AVX2+FMA host): the ARM optimized-routines algorithm — log x as a double-double from a 128-entry
 * table ({1/c, log c} per subinterval, r = x/c - 1 exact in one fma, degree-7 polynomial), times y as a double-double,
 * then exp from a 128-entry table of 2^(i/128) with a degree-5 polynomial.  It is < 1 ULP but not correctly rounded,
 * so the operation ORDER and every fma contraction below follow the compiled libm (disassembly of __pow_fma, main
 * path), not a textbook formula; tables from tools/gen_libm_tables.py.  tests/test_portable_math.py: equal to libm pow
 * on > 1e8 inputs.  Domain (all the path needs): x positive normal finite; y = 0 or 2^-65 <= |y| < 2^63; |y log x| < 512
 * is exact-parity, beyond that the result saturates to inf / 0 without glibc's subnormal rounding.
 * Reference call sites: cell.py:246-248 (mass^-0.35), player.py:163-167 (r^0.475, n^0.32), replay_buffer.py:166-199. */
#ifdef __CUDACC__
static __device__ const double agar_pow_log_tab_d[128][3] = {AGAR_POW_LOG_TAB};
static __device__ const unsigned long long agar_exp_tab_d[256] = {AGAR_EXP_TAB};
static __device__ const double agar_atan_tab_d[241][7] = {AGAR_ATAN_TAB};
static __device__ const double agar_sincos_tab_d[440] = {AGAR_SINCOS_TAB};
#endif
static const double agar_pow_log_tab_h[128][3] = {AGAR_POW_LOG_TAB};
static const unsigned long long agar_exp_tab_h[256] = {AGAR_EXP_TAB};
static const double agar_atan_tab_h[241][7] = {AGAR_ATAN_TAB};
static const double agar_sincos_tab_h[440] = {AGAR_SINCOS_TAB};
#ifdef __CUDA_ARCH__
#define AGAR_POW_LOG_T(i, j) __ldg(&agar_pow_log_tab_d[i][j])
#define AGAR_EXP_T(i) __ldg(&agar_exp_tab_d[i])
#define AGAR_ATAN_T(i, j) __ldg(&agar_atan_tab_d[i][j])
#define AGAR_SINCOS_T(i) __ldg(&agar_sincos_tab_d[i])
#else
#define AGAR_POW_LOG_T(i, j) agar_pow_log_tab_h[i][j]
#define AGAR_EXP_T(i) agar_exp_tab_h[i]
#define AGAR_ATAN_T(i, j) agar_atan_tab_h[i][j]
#define AGAR_SINCOS_T(i) agar_sincos_tab_h[i]
#endif

AGAR_HD double agar_pow(double x, double y) {
    if (y == 0.0) return 1.0;
    const uint64_t ix = agar_double_to_bits(x);
    /* log_inline: x = 2^k z, z in [0x1.69555p-1, 0x1.69555p0), c = centre of z's subinterval */
    const uint64_t tmp = ix - 0x3fe6955500000000ULL;
    const int i = (int)((tmp >> 45) & 127);
    const int k = (int)((int64_t)tmp >> 52);
    const double z = agar_bits_to_double(ix - (tmp & 0xfff0000000000000ULL));
    const double kd = (double)k;
    const double invc = AGAR_POW_LOG_T(i, 0), logc = AGAR_POW_LOG_T(i, 1), logctail = AGAR_POW_LOG_T(i, 2);
    const double r = fma(z, invc, -1.0);
    const double t1 = fma(kd, AGAR_POW_LN2HI, logc);
    const double lo1 = fma(kd, AGAR_POW_LN2LO, logctail);
    const double t2 = t1 + r;
    const double lo2 = (t1 - t2) + r;
    const double ar = AGAR_POW_A0 * r;
    const double ar2 = r * ar;
    const double ar3 = r * ar2;
    const double hi = t2 + ar2;
    const double lo3 = fma(ar, r, -ar2);
    const double lo4 = (t2 - hi) + ar2;
    const double q = fma(ar2, fma(fma(r, AGAR_POW_A6, AGAR_POW_A5), ar2, fma(r, AGAR_POW_A4, AGAR_POW_A3)),
                         fma(r, AGAR_POW_A2, AGAR_POW_A1));
    const double lo = fma(ar3, q, ((lo1 + lo2) + lo3) + lo4);
    const double lhi = hi + lo;
    const double llo = (hi - lhi) + lo;
    /* y * log x as ehi + elo */
    const double ehi = y * lhi;
    const double elo = fma(y, llo, fma(y, lhi, -ehi));
    /* exp_inline(ehi, elo) */
    const uint32_t abstop = (uint32_t)(agar_double_to_bits(ehi) >> 52) & 0x7ff;
    if (abstop - 0x3c9u >= 0x3fu) {
        if (abstop < 0x3c9u) return 1.0 + ehi;      /* |y log x| < 2^-54 */
        return ehi > 0 ? (double)INFINITY : 0.0;     /* |y log x| >= 512: outside the path's domain */
    }
    double kz = fma(ehi, AGAR_EXP_INVLN2N, AGAR_EXP_SHIFT);
    const uint64_t ki = agar_double_to_bits(kz);
    kz -= AGAR_EXP_SHIFT;
    double rr = fma(kz, AGAR_EXP_NEGLN2HIN, ehi);
    rr = fma(kz, AGAR_EXP_NEGLN2LON, rr);
    rr = elo + rr;
    const int idx = 2 * (int)(ki & 127);
    const double tail = agar_bits_to_double(AGAR_EXP_T(idx));
    const uint64_t sbits = AGAR_EXP_T(idx + 1) + (ki << 45);
    const double r2 = rr * rr;
    const double p23 = fma(rr, AGAR_EXP_C3, AGAR_EXP_C2);
    const double p45 = fma(rr, AGAR_EXP_C5, AGAR_EXP_C4);
    double t = fma(p23, r2, rr + tail);
    t = fma(p45, r2 * r2, t);
    const double scale = agar_bits_to_double(sbits);
    return fma(t, scale, scale);
}

/* ---- atan2, sin, cos: BIT-IDENTICAL to glibc 2.39's FMA builds (what math.atan2 / math.cos / math.sin call on this image).
 * e_atan2.c (IBM Accurate Mathematical Library): u = min(|y|,|x|) / max(|y|,|x|) with its division residual du; u < 1/16 ->
 * odd polynomial to u^13, else the table cij[i] = {x_i, atan x_i, expansion of atan around x_i}; the octant decides how the
 * result is assembled from pi/2 or pi as a double-double.  s_sin.c: |x| < 0.126 -> Taylor; else x = x_k + r with
 * {sin x_k, cos x_k} from __sincostab (x_k ~ k/128) and short polynomials in r; 0.855 < |x| < 2.426 goes through pi/2 - |x|,
 * larger arguments through the 4-part Cody-Waite reduction by pi/2.  Operation order and fma contractions follow the compiled
 * code (disassembly of __atan2_fma / __sin_fma / __cos_fma), e.g. `s = x + x*xx*(sn3 + xx*sn5)` is ONE fma in do_cos.
 * Domain: finite arguments; atan2 additionally |x|, |y| in {0} or [2^-500, 2^500] (no rescaling step, no subnormal results);
 * sin / cos |x| < 105414350 (the path only passes angles in [-pi, pi]).  tests/test_portable_math.py: equal to libm on
 * > 1e8 inputs each.  Reference call sites: cell.py:49-57,96-103 (atan2 -> cos, sin), field.py:363-366 (integer degrees). */
#define AGAR_HPI 0x1.921fb54442d18p+0
#define AGAR_HPI1 0x1.1a62633145c07p-54
#define AGAR_OPI 0x1.921fb54442d18p+1
#define AGAR_OPI1 0x1.1a62633145c07p-53

AGAR_HD double agar_copysign(double mag, double sgn) {
    return agar_bits_to_double((agar_double_to_bits(mag) & 0x7fffffffffffffffULL) | (agar_double_to_bits(sgn) & 0x8000000000000000ULL));
}

AGAR_HD double agar_atan2(double y, double x) {
    const uint64_t bx = agar_double_to_bits(x), by = agar_double_to_bits(y);
    const int xneg = (int)(bx >> 63);
    if ((by << 1) == 0) { /* y = +-0: +-0 for x >= +0, +-pi otherwise */
        if (!xneg) return y;
        return (by >> 63) ? -AGAR_OPI : AGAR_OPI;
    }
    if (x == 0.0) return (by >> 63) ? -AGAR_HPI : AGAR_HPI;
    double ax = fabs(x), ay = fabs(y);
    const int de = (int)((uint32_t)(by >> 32) & 0x7ff00000u) - (int)((uint32_t)(bx >> 32) & 0x7ff00000u);
    if (de >= 59768832) return y > 0 ? AGAR_HPI : -AGAR_HPI;      /* |y / x| > 2^57 */
    if (de <= -59768832) {                                          /* |y / x| < 2^-57 */
        if (x > 0) return agar_copysign(ay / ax, y);
        return y > 0 ? AGAR_OPI : -AGAR_OPI;
    }
    /* u = min / max with the residual of the division; the two octant halves run the same operations on swapped operands, so the
     * operands are selected and the divisions exist once (lanes of a warp that sit in different octants stay converged) */
    const int steep = !(ay < ax);
    const double num = steep ? ax : ay, den = steep ? ay : ax;
    const double u = num / den;
    double du;
    {
        const double v = u * den, vv = fma(u, den, -v);
        du = ((num - v) - vv) / den;
    }
    double z;
    const int case_i = x > 0 && !steep; /* (i) atan(ay / ax) */
    /* (ii) x > 0: pi/2 - atan(ax/ay)   (iii) x < 0, ax < ay: pi/2 + atan(ax/ay)   (iv) x < 0: pi - atan(ay/ax);
     * a subtraction is the addition of the exactly negated operand, so one code path serves all three */
    const int iii = x < 0 && ax < ay;
    const double K0 = (x > 0 || iii) ? AGAR_HPI : AGAR_OPI, K1 = (x > 0 || iii) ? AGAR_HPI1 : AGAR_OPI1;
    if (u < 0.0625) {
        const double v = u * u;
        const double P = fma(v, fma(v, fma(v, fma(v, fma(v, 0x1.375f08b31cbcep-4, -0x1.7458022b13c25p-4), 0x1.c71c6e5129a3bp-4),
                                           -0x1.24924923f7603p-3), 0x1.99999999997fdp-3), -0x1.5555555555555p-2);
        if (case_i) {
            z = u + fma(u * v, P, du);
        } else {
            const double zz = (u * v) * P;
            const double su = iii ? u : -u, sdu = iii ? du : -du, szz = iii ? zz : -zz;
            const double t2 = K0 + su;
            const double cor = (K0 - t2) + su;
            z = (((cor + K1) + sdu) + szz) + t2;
        }
    } else {
        const int i = (int)(fma(u, 256.0, 4503599627370496.0) - 4503599627370496.0) - 16;
        if (case_i) {
            const double t3 = u - AGAR_ATAN_T(i, 0);
            const double v = du + t3;
            const double dv = fabs(t3) > fabs(du) ? (t3 - v) + du : (du - v) + t3;
            const double t2 = AGAR_ATAN_T(i, 2);
            const double poly = fma(v, fma(v, fma(v, AGAR_ATAN_T(i, 6), AGAR_ATAN_T(i, 5)), AGAR_ATAN_T(i, 4)), AGAR_ATAN_T(i, 3));
            z = fma(v, t2, fma(dv, t2, (v * v) * poly)) + AGAR_ATAN_T(i, 1);
        } else {
            const double v = (u - AGAR_ATAN_T(i, 0)) + du;
            const double q = fma(v, fma(v, fma(v, fma(v, AGAR_ATAN_T(i, 6), AGAR_ATAN_T(i, 5)), AGAR_ATAN_T(i, 4)), AGAR_ATAN_T(i, 3)),
                                 AGAR_ATAN_T(i, 2));
            const double c1 = AGAR_ATAN_T(i, 1);
            z = (K0 + (iii ? c1 : -c1)) + fma(iii ? v : -v, q, K1);
        }
    }
    return agar_copysign(z, y);
}

/* s_sin.c helpers; T = __sincostab */
AGAR_HD double agar__taylor_sin(double a, double da) {
    const double xx = a * a;
    const double p = fma(xx, fma(xx, fma(xx, fma(xx, -0x1.addffc2fcdf59p-26, 0x1.71de27b9a7ed9p-19), -0x1.a01a019db08b8p-13),
                                 0x1.1111111110ecep-7), -0x1.5555555555555p-3);
    return a + fma(xx, fma(p, a, -(0.5 * da)), da);
}
AGAR_HD double agar__do_sin(double a, double da) {
    const double dxs = a <= 0 ? -da : da, aa = fabs(a);
    const double u = 0x1.8p45 + aa;
    const double xr = aa - (u - 0x1.8p45);
    const int k = (int)((uint32_t)agar_double_to_bits(u) << 2);
    const double xx = xr * xr;
    const double s = xr + fma(xr * xx, fma(xx, 0x1.11110e829872fp-7, -0x1.5555555555515p-3), dxs);
    const double c = fma(xr, dxs, xx * fma(xx, fma(xx, 0x1.6c16bedd9e239p-10, -0x1.5555555555535p-5), 0.5));
    const double sn = AGAR_SINCOS_T(k), ssn = AGAR_SINCOS_T(k + 1), cs = AGAR_SINCOS_T(k + 2), ccs = AGAR_SINCOS_T(k + 3);
    const double cor = fma(s, cs, fma(-c, sn, fma(s, ccs, ssn)));
    return agar_copysign(sn + cor, a);
}
AGAR_HD double agar__do_cos(double a, double da) {
    const double dxs = a < 0 ? -da : da, aa = fabs(a);
    const double u = 0x1.8p45 + aa;
    const double xr = (aa - (u - 0x1.8p45)) + dxs;
    const int k = (int)((uint32_t)agar_double_to_bits(u) << 2);
    const double xx = xr * xr;
    const double s = fma(xr * xx, fma(xx, 0x1.11110e829872fp-7, -0x1.5555555555515p-3), xr);
    const double c = xx * fma(xx, fma(xx, 0x1.6c16bedd9e239p-10, -0x1.5555555555535p-5), 0.5);
    const double sn = AGAR_SINCOS_T(k), ssn = AGAR_SINCOS_T(k + 1), cs = AGAR_SINCOS_T(k + 2), ccs = AGAR_SINCOS_T(k + 3);
    const double cor = fma(-s, sn, fma(-c, cs, fma(-s, ssn, ccs)));
    return cs + cor;
}
/* reduce_sincos: x = n pi/2 + (a + da), |a| <= pi/4; returns n mod 4 */
AGAR_HD int agar__reduce_sincos(double x, double* a, double* da) {
    const double t = fma(x, 0x1.45f306dc9c883p-1, 0x1.8p52);
    const double xn = t - 0x1.8p52;
    const int n = (int)(agar_double_to_bits(t) & 3);
    const double y = fma(-xn, -0x1.dde973c000000p-27, fma(-xn, 0x1.921fb58000000p+0, x));
    const double t2 = fma(-xn, -0x1.cb3b398000000p-55, y);
    const double d0 = fma(-xn, -0x1.cb3b398000000p-55, y - t2);
    const double b = fma(-xn, -0x1.d747f23e32ed7p-83, t2);
    *a = b;
    *da = d0 + fma(-xn, -0x1.d747f23e32ed7p-83, t2 - b);
    return n;
}
AGAR_HD double agar__sin_small(double a, double da) { return fabs(a) < 0.126 ? agar__taylor_sin(a, da) : agar__do_sin(a, da); }

AGAR_HD double agar_sin(double x) {
    const uint32_t k = (uint32_t)(agar_double_to_bits(x) >> 32) & 0x7fffffffu;
    if (k < 0x3e500000u) return x;
    if (k < 0x3feb6000u) return agar__sin_small(x, 0.0);
    if (k < 0x400368fdu) return agar_copysign(agar__do_cos(0x1.921fb54442d18p+0 - fabs(x), 0x1.1a62633145c07p-54), x);
    double a, da;
    const int n = agar__reduce_sincos(x, &a, &da);
    const double r = (n & 1) ? agar__do_cos(a, da) : agar__sin_small(a, da);
    return (n & 2) ? -r : r;
}
AGAR_HD double agar_cos(double x) {
    const uint32_t k = (uint32_t)(agar_double_to_bits(x) >> 32) & 0x7fffffffu;
    if (k < 0x3e400000u) return 1.0;
    if (k < 0x3feb6000u) return agar__do_cos(x, 0.0);
    if (k < 0x400368fdu) {
        const double y = 0x1.921fb54442d18p+0 - fabs(x);
        const double a = y + 0x1.1a62633145c07p-54;
        return agar__sin_small(a, (y - a) + 0x1.1a62633145c07p-54);
    }
    double a, da;
    const int n = agar__reduce_sincos(x, &a, &da) + 1;
    const double r = (n & 1) ? agar__do_cos(a, da) : agar__sin_small(a, da);
    return (n & 2) ? -r : r;
}

/* sin x and cos x together, the SAME arithmetic as agar_sin / agar_cos.  In every branch of s_sin.c exactly one of the two goes
 * through the cos-like leaf (do_cos) and the other through the sin-like leaf (Taylor / do_sin), on arguments that differ only in
 * how the reduction hands them over — so the pair costs ONE evaluation of each leaf, and on a GPU all lanes of a warp run the
 * cos-like leaf together whatever branch their angle took (calling agar_sin and agar_cos separately runs up to three leaves per
 * call on different lane subsets). */
AGAR_HD void agar_sincos(double x, double* sn, double* cs) {
    const uint32_t k = (uint32_t)(agar_double_to_bits(x) >> 32) & 0x7fffffffu;
    double ac, dac, as, das;   /* arguments of the cos-like and of the sin-like leaf */
    int cos_leaf_is_sin;       /* the cos-like leaf's value is sin x (and the sin-like leaf's is cos x) */
    int neg_s = 0, neg_c = 0, csign = 0;
    if (k < 0x3feb6000u) {             /* |x| < 0.855469: sin = sin_small(x, 0), cos = do_cos(x, 0) */
        ac = x, dac = 0.0, as = x, das = 0.0, cos_leaf_is_sin = 0;
    } else if (k < 0x400368fdu) {      /* |x| < 2.426265: sin = copysign(do_cos(pi/2 - |x|, hp1), x), cos = sin_small(a, da) */
        const double y = 0x1.921fb54442d18p+0 - fabs(x);
        ac = y, dac = 0x1.1a62633145c07p-54;
        as = y + 0x1.1a62633145c07p-54;
        das = (y - as) + 0x1.1a62633145c07p-54;
        cos_leaf_is_sin = 1, csign = 1;
    } else {                            /* reduce_sincos: sin takes quadrant n, cos quadrant n + 1 */
        double a, da;
        const int n = agar__reduce_sincos(x, &a, &da);
        ac = a, dac = da, as = a, das = da;
        cos_leaf_is_sin = n & 1;        /* n odd: sin x = +-do_cos(a, da), cos x = +-sin_small(a, da) */
        neg_s = n & 2, neg_c = (n + 1) & 2;
    }
    double vc = agar__do_cos(ac, dac);
    double vs = agar__sin_small(as, das);
    if (csign) vc = agar_copysign(vc, x);
    double s_val = cos_leaf_is_sin ? vc : vs, c_val = cos_leaf_is_sin ? vs : vc;
    if (neg_s) s_val = -s_val;
    if (neg_c) c_val = -c_val;
    if (k < 0x3e500000u) s_val = x;   /* |x| < 2^-26 */
    if (k < 0x3e400000u) c_val = 1.0; /* |x| < 2^-27 */
    *sn = s_val, *cs = c_val;
}

/* (cos a, sin a) for a = atan2(dy, dx) — cell.py:49-57: the angle is formed and rounded, then cos and sin are taken of it. */
AGAR_HD void agar_dir(double dy, double dx, double* c, double* s) {
    const double a = agar_atan2(dy, dx);
    agar_sincos(a, s, c);
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
