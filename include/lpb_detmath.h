/* lpb_detmath.h -- deterministic elementary functions for lpopc-b200 functors.
 *
 * Every function here is built from IEEE-754 binary64 + - * / sqrt floor and
 * integer bit moves only, so that g++ (with -ffp-contract=off) and nvcc (with
 * --fmad=false) produce bit-identical results.  The reference's examples call
 * libm through Armadillo (exp/pow/acos in Lpopc/example/launch/Launch.cpp:589-738);
 * glibc and CUDA libdevice differ in the last ulp, and a forward difference
 * (LpFiniteDifferenceDerive.cpp:208-259, h ~ 1e-6) amplifies one ulp of f to
 * ~1e-10 relative in the Jacobian -- above the 1e-12 parity bound.  Functors
 * therefore use these instead of libm.  Accuracy is a few ulp (checked against
 * numpy in tests/test_detmath.py); NaN propagates; no fast-math assumptions.
 *
 * The polynomial kernels follow the classic published fdlibm algorithms
 * (argument reduction + minimax polynomial); constants are written as decimal
 * literals so both compilers parse the same binary64 values.
 */
#ifndef LPB_DETMATH_H
#define LPB_DETMATH_H

#include <math.h>
#include <string.h>

#ifndef LPB_HD
#if defined(__CUDACC__)
#define LPB_HD __host__ __device__ __forceinline__
#else
#define LPB_HD inline
#endif
#endif

LPB_HD double lpb_from_bits(long long b)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(b);
#else
    double d;
    memcpy(&d, &b, sizeof d);
    return d;
#endif
}

LPB_HD long long lpb_to_bits(double d)
{
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(d);
#else
    long long b;
    memcpy(&b, &d, sizeof b);
    return b;
#endif
}

/* 2^k for -1022 <= k <= 1023 */
LPB_HD double lpb_pow2i(long long k) { return lpb_from_bits((k + 1023LL) << 52); }

LPB_HD double lpb_det_nan(double x)
{
    double z = x - x; /* 0, or NaN for inf/NaN */
    return z / z;
}

/* ---- exp ---------------------------------------------------------------- */
LPB_HD double lpb_det_exp(double x)
{
    const double ln2HI = 6.93147180369123816490e-01;
    const double ln2LO = 1.90821492927058770002e-10;
    const double invln2 = 1.44269504088896338700e+00;
    const double P1 = 1.66666666666666019037e-01;
    const double P2 = -2.77777777770155933842e-03;
    const double P3 = 6.61375632143793436117e-05;
    const double P4 = -1.65339022054652515390e-06;
    const double P5 = 4.13813679705723846039e-08;
    const bool isnan_ = x != x;
    const bool big = x > 709.782712893384, small = x < -745.2;
    const double xr = (isnan_ || big || small) ? 0.0 : x;
    double kd = floor(xr * invln2 + 0.5);
    double hi = xr - kd * ln2HI;
    double lo = kd * ln2LO;
    double r = hi - lo;
    double t = r * r;
    double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
    long long k = (long long)kd;
    long long k1 = k / 2;
    long long k2 = k - k1;
    double v = (y * lpb_pow2i(k1)) * lpb_pow2i(k2);
    v = small ? 0.0 : v;
    v = big ? lpb_from_bits(0x7ff0000000000000LL) : v;
    return isnan_ ? x : v;
}

/* ---- tanh --------------------------------------------------------------- */
/* tanh(x) = e1/(e1+2), e1 = expm1(2|x|); expm1 by ln2 argument reduction and a
 * degree-15 Taylor kernel on |r| <= ln2/2.  Branch-free apart from the clamp so
 * a warp never diverges on it. */
LPB_HD double lpb_det_tanh(double x)
{
    const double ln2HI = 6.93147180369123816490e-01;
    const double ln2LO = 1.90821492927058770002e-10;
    const double invln2 = 1.44269504088896338700e+00;
    double ax = fabs(x);
    ax = (ax > 20.0) ? 20.0 : ax; /* NaN stays NaN (comparison false) */
    double y = ax + ax;
    double kd = floor(y * invln2 + 0.5);
    double hi = y - kd * ln2HI;
    double lo = kd * ln2LO;
    double r = hi - lo;
    double p = 7.647163731819816e-13;
    p = 1.1470745597729725e-11 + r * p;
    p = 1.6059043836821613e-10 + r * p;
    p = 2.08767569878681e-09 + r * p;
    p = 2.505210838544172e-08 + r * p;
    p = 2.755731922398589e-07 + r * p;
    p = 2.7557319223985893e-06 + r * p;
    p = 2.48015873015873e-05 + r * p;
    p = 0.0001984126984126984 + r * p;
    p = 0.001388888888888889 + r * p;
    p = 0.008333333333333333 + r * p;
    p = 0.041666666666666664 + r * p;
    p = 0.16666666666666666 + r * p;
    p = 0.5 + r * p;
    p = 1.0 + r * p;
    p = r * p; /* expm1(r) */
    /* kd is NaN when x is NaN: keep the conversion defined on both sides */
    long long k = (kd == kd) ? (long long)kd : 0LL;
    double twok = lpb_pow2i(k);
    double e1 = twok * p + (twok - 1.0);
    double th = e1 / (e1 + 2.0);
    return (x < 0.0) ? -th : th;
}

/* ---- sin / cos ---------------------------------------------------------- */
LPB_HD double lpb_ksin(double r)
{
    const double S1 = -1.66666666666666324348e-01;
    const double S2 = 8.33333333332248946124e-03;
    const double S3 = -1.98412698298579493134e-04;
    const double S4 = 2.75573137070700676789e-06;
    const double S5 = -2.50507602534068634195e-08;
    const double S6 = 1.58969099521155010221e-10;
    double z = r * r;
    double v = z * r;
    double q = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    return r + v * (S1 + z * q);
}

LPB_HD double lpb_kcos(double r)
{
    const double C1 = 4.16666666666666019037e-02;
    const double C2 = -1.38888888888741095749e-03;
    const double C3 = 2.48015872894767294178e-05;
    const double C4 = -2.75573143513906633035e-07;
    const double C5 = 2.08757232129817482790e-09;
    const double C6 = -1.13596475577881948265e-11;
    double z = r * r;
    double q = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    return 1.0 - (0.5 * z - z * q);
}

/* Cody-Waite reduction to [-pi/4, pi/4]; valid for |x| < 1e6, NaN beyond. */
LPB_HD double lpb_reduce_pio2(double x, long long* quadrant)
{
    const double invpio2 = 6.36619772367581382433e-01;
    const double pio2_1 = 1.57079632673412561417e+00;
    const double pio2_2 = 6.07710050630396597660e-11;
    const double pio2_3 = 2.02226624871116645580e-21;
    const double pio2_3t = 8.47842766036889956997e-32;
    double kd = floor(x * invpio2 + 0.5);
    double r = ((x - kd * pio2_1) - kd * pio2_2) - kd * pio2_3;
    r = r - kd * pio2_3t;
    *quadrant = ((long long)kd) & 3LL;
    return r;
}

/* sin/cos are straight-line code (selects, no branches): a perturbed dae() evaluation
 * whose angle argument is unchanged is then a literal common subexpression of the base
 * evaluation, which the compile-time colour unrolling of the Jacobian kernel relies on. */
LPB_HD double lpb_det_sin(double x)
{
    const bool ok = fabs(x) < 1.0e6; /* false for NaN */
    long long q;
    double r = lpb_reduce_pio2(ok ? x : 0.0, &q);
    double s = lpb_ksin(r);
    double c = lpb_kcos(r);
    double v = (q & 1LL) ? c : s;
    v = (q & 2LL) ? -v : v;
    return ok ? v : lpb_from_bits(0x7ff8000000000000LL);
}

LPB_HD double lpb_det_cos(double x)
{
    const bool ok = fabs(x) < 1.0e6;
    long long q;
    double r = lpb_reduce_pio2(ok ? x : 0.0, &q);
    double s = lpb_ksin(r);
    double c = lpb_kcos(r);
    double v = (q & 1LL) ? s : c;
    v = (((q + 1LL) & 2LL) != 0LL) ? -v : v;
    return ok ? v : lpb_from_bits(0x7ff8000000000000LL);
}

/* ---- acos --------------------------------------------------------------- */
LPB_HD double lpb_acos_R(double z)
{
    const double pS0 = 1.66666666666666657415e-01;
    const double pS1 = -3.25565818622400915405e-01;
    const double pS2 = 2.01212532134862925881e-01;
    const double pS3 = -4.00555345006794114027e-02;
    const double pS4 = 7.91534994289814532176e-04;
    const double pS5 = 3.47933107596021167570e-05;
    const double qS1 = -2.40339491173441421878e+00;
    const double qS2 = 2.02094576023350569471e+00;
    const double qS3 = -6.88283971605453293030e-01;
    const double qS4 = 7.70381505559019352791e-02;
    double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    return p / q;
}

LPB_HD double lpb_det_acos(double x)
{
    const double pio2_hi = 1.57079632679489655800e+00;
    const double pio2_lo = 6.12323399573676603587e-17;
    const double pi = 3.14159265358979311600e+00;
    if (x != x) return x;
    double ax = fabs(x);
    if (ax > 1.0) return lpb_det_nan(lpb_from_bits(0x7ff0000000000000LL));
    if (ax == 1.0) return (x > 0.0) ? 0.0 : pi + 2.0 * pio2_lo;
    if (ax < 0.5) {
        double z = x * x;
        double r = lpb_acos_R(z);
        return pio2_hi - (x - (pio2_lo - x * r));
    }
    if (x < 0.0) {
        double z = (1.0 + x) * 0.5;
        double s = sqrt(z);
        double r = lpb_acos_R(z);
        double w = r * s - pio2_lo;
        return pi - 2.0 * (s + w);
    }
    {
        double z = (1.0 - x) * 0.5;
        double s = sqrt(z);
        double df = lpb_from_bits(lpb_to_bits(s) & (long long)0xffffffff00000000ULL);
        double c = (z - df * df) / (s + df);
        double r = lpb_acos_R(z);
        double w = r * s + c;
        return 2.0 * (df + w);
    }
}

#endif /* LPB_DETMATH_H */
