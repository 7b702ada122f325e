/*
 * cpm_detmath.h -- deterministic fp32 elementary functions shared by the CUDA kernels
 * and by the CPU oracle.
 *
 * Why this exists: the reference tracer calls OpenCL `native_log`, `native_sin/cos`,
 * `acos`, `atan2` (ppm/cl/transmittance.cl:135, Inviwo shading/transformations headers,
 * ppm/photondata.cpp:100-117).  Those are implementation-defined, and one flipped
 * accept/reject in Woodcock tracking sends a photon down a different path.  To make
 * "same seed -> same photon" a bit-exact statement between the sm_100a kernels and the
 * CPU oracle, every transcendental on the path is defined HERE, once, in terms of IEEE-754
 * correctly rounded fp32 operations only (+, -, *, /, sqrt, fma, floor/rint, int<->float
 * conversions and bit casts).  Those operations give identical bits on an x86 host
 * (gcc -ffp-contract=off, fmaf -> vfmadd with -mfma) and on the GPU (nvcc -fmad=false,
 * default -prec-div=true -prec-sqrt=true -ftz=false).
 *
 * Accuracy (checked in tests/test_detmath.py against float64 libm): <= 2 ulp (cpm_expf_sym 3, cpm_powf 6) on the
 * argument ranges the path uses.  Coefficients: Taylor for log/sin/cos (error bound in
 * comments), least-squares Chebyshev fits from tools/fit_detmath.py for asin/atan.
 *
 * Everything is `static inline` and usable from C99, C++ and CUDA.
 */
#ifndef CPM_DETMATH_H
#define CPM_DETMATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define CPM_HD __host__ __device__ __forceinline__
#else
#define CPM_HD static inline
#endif

#define CPM_PI_F 3.14159274101257324219f      /* float(pi)            0x1.921fb6p+1 */
#define CPM_2PI_F 6.28318548202514648438f     /* float(2*pi)          0x1.921fb6p+2 */
#define CPM_PIO2_HI 1.57079637050628662109f   /* float(pi/2)          0x1.921fb6p+0 */
#define CPM_PIO2_LO -4.37113882867379118e-8f  /* float(pi/2 - PIO2_HI) */
#define CPM_PIO4_F 0.78539818525314331055f    /* float(pi/4) */
#define CPM_2OPI_F 0.63661974668502807617f    /* float(2/pi) */
#define CPM_LN2_HI 0.693145751953125f         /* 0x1.62e3p-1, 16 significant bits */
#define CPM_LN2_LO 1.42860677e-06f            /* ln2 - LN2_HI */
#define CPM_INV_4PI_F 0.07957747154594767f    /* 1/(4 pi) */

CPM_HD uint32_t cpm_f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
CPM_HD float cpm_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

/* min/max with the IEEE "return the non-NaN operand" rule, written with compares so the
 * host and the device agree bit for bit (fminf/fmaxf may differ in the sign of zero). */
CPM_HD float cpm_fmin(float a, float b) { return (a != a) ? b : ((b < a) ? b : a); }
CPM_HD float cpm_fmax(float a, float b) { return (a != a) ? b : ((b > a) ? b : a); }
/* clamp that maps NaN to lo */
CPM_HD float cpm_clamp(float x, float lo, float hi) {
    if (!(x >= lo)) x = lo;
    if (x > hi) x = hi;
    return x;
}

/* Natural logarithm for x >= 0.  x == 0 -> -inf, x == 1 -> 0 exactly.
 * x = 2^e * m, m in [sqrt(1/2), sqrt(2)), f = m - 1:
 *   log m = f - f^2/2 + f^3 P(f),  P = the degree-8 polynomial of Cephes logf (S. Moshier),
 * all Horner steps are fma, no division (the tracer evaluates this once per delta-tracking
 * step).  Measured against float64 log on 6M arguments of the form k * 2^-32 and a dense
 * sweep of [0.4, 1]: max 0.83 ulp.  The exponent becomes a float through the 1.5 * 2^23 bit
 * trick (exact for |e| < 2^22) instead of an int->float conversion.
 * Zero, subnormal, infinite and NaN arguments leave through one rarely taken branch. */
CPM_HD float cpm_logf(float x) {
    uint32_t ix = cpm_f2u(x);
    float eadj = 0.0f;
    if (ix - 0x00800000u >= 0x7f000000u) {          /* not a positive normal number */
        if ((ix << 1) == 0u) return cpm_u2f(0xff800000u); /* log(+-0) = -inf */
        if (ix >= 0x7f800000u) return x + x;        /* inf, nan, negative (not reached on the path) */
        x = x * 33554432.0f;                        /* subnormal: scale by 2^25 */
        ix = cpm_f2u(x);
        eadj = -25.0f;
    }
    /* bring the mantissa into [sqrt(1/2), sqrt(2)) */
    ix += 0x3f800000u - 0x3f3504f3u;
    uint32_t eb = (ix >> 23) + (0x4B400000u - 127u);        /* bits of 12582912 + e */
    float fe = (cpm_u2f(eb) - 12582912.0f) + eadj;
    ix = (ix & 0x007fffffu) + 0x3f3504f3u;
    float f = cpm_u2f(ix) - 1.0f;
    float z = f * f;
    float p = 7.0376836292E-2f;
    p = fmaf(p, f, -1.1514610310E-1f);
    p = fmaf(p, f, 1.1676998740E-1f);
    p = fmaf(p, f, -1.2420140846E-1f);
    p = fmaf(p, f, 1.4249322787E-1f);
    p = fmaf(p, f, -1.6668057665E-1f);
    p = fmaf(p, f, 2.0000714765E-1f);
    p = fmaf(p, f, -2.4999993993E-1f);
    p = fmaf(p, f, 3.3333331174E-1f);
    float y = (f * z) * p;
    y = fmaf(fe, -2.12194440e-4f, y);               /* ln2 - 0.693359375 */
    y = fmaf(-0.5f, z, y);
    float r = f + y;
    return fmaf(fe, 0.693359375f, r);
}

/* native_log (ppm/cl/transmittance.cl:135 `t += -native_log(random_01(...)) * ...`): the delta-tracking step evaluates one
 * logarithm per collision test, so this one is built for instruction count.  x = 2^e * m with m in [0.75, 1.5)
 * (adding 0x00400000 to the bits moves mantissas >= 1.5 to the next exponent), the top five bits of the shifted
 * mantissa field select one of 32 intervals with a tabulated reciprocal rc of its centre and lc = -log(rc) (exact
 * reciprocal, so r = m * rc - 1 carries no table error; one fma), and
 *     log x = e ln2 + lc + (r - r^2/2 + r^3/3 - r^4/4 + r^5/5),     |r| < 2^-5: next term < 5e-9 relative.
 * The two intervals touching m = 1 use rc = 1, lc = 0, so log(x) -> x - 1 keeps its relative accuracy near 1 and
 * log(1) == 0 exactly.  Measured against float64 log on the tracer's arguments k * 2^-32: <= 2 ulp.
 * `tab` = CPM_NLOG_TABLE (64 floats: rc, lc per interval): a static array on the host, shared memory in the tracer.
 * Zero, subnormal, infinite, NaN and negative arguments take cpm_logf's exits. */
#define CPM_NLOG_TABLE \
    0x1.51d07e0000000p+0f, -0x1.1bf9940000000p-2f, 0x1.4afd6a0000000p+0f, -0x1.0713860000000p-2f, \
    0x1.446f860000000p+0f, -0x1.e530ee0000000p-3f, 0x1.3e22cc0000000p+0f, -0x1.bd08740000000p-3f, \
    0x1.3813820000000p+0f, -0x1.95a5b20000000p-3f, 0x1.323e340000000p+0f, -0x1.6f01240000000p-3f, \
    0x1.2c9fb40000000p+0f, -0x1.4913d20000000p-3f, 0x1.27350c0000000p+0f, -0x1.23d7160000000p-3f, \
    0x1.21fb780000000p+0f, -0x1.fe89120000000p-4f, 0x1.1cf06a0000000p+0f, -0x1.b6ac7c0000000p-4f, \
    0x1.1811820000000p+0f, -0x1.700d3e0000000p-4f, 0x1.135c820000000p+0f, -0x1.2aa0580000000p-4f, \
    0x1.0ecf560000000p+0f, -0x1.ccb7260000000p-5f, 0x1.0a68100000000p+0f, -0x1.466ada0000000p-5f, \
    0x1.0624de0000000p+0f, -0x1.8492860000000p-6f, 0x1.0000000000000p+0f, 0x0p+0f, \
    0x1.0000000000000p+0f, 0x0p+0f, 0x1.e9131a0000000p-1f, 0x1.77459c0000000p-5f, \
    0x1.dae6080000000p-1f, 0x1.341d740000000p-4f, 0x1.cd85680000000p-1f, 0x1.a926d80000000p-4f, \
    0x1.c0e0700000000p-1f, 0x1.0d77e80000000p-3f, 0x1.b4e81c0000000p-1f, 0x1.44d2b40000000p-3f, \
    0x1.a98ef60000000p-1f, 0x1.7ab8900000000p-3f, 0x1.9ec8ea0000000p-1f, 0x1.af3c920000000p-3f, \
    0x1.948b100000000p-1f, 0x1.e270760000000p-3f, 0x1.8acb900000000p-1f, 0x1.0a32500000000p-2f, \
    0x1.8181820000000p-1f, 0x1.22941e0000000p-2f, 0x1.78a4c80000000p-1f, 0x1.3a64c60000000p-2f, \
    0x1.702e060000000p-1f, 0x1.51aad80000000p-2f, 0x1.6816820000000p-1f, 0x1.686c800000000p-2f, \
    0x1.6058160000000p-1f, 0x1.7eaf840000000p-2f, 0x1.58ed240000000p-1f, 0x1.94793e0000000p-2f

CPM_HD float cpm_native_logf_tab(float x, const float* tab) {
    uint32_t ix = cpm_f2u(x);
    if (ix - 0x00800000u >= 0x7f000000u) return cpm_logf(x);
    uint32_t jx = ix + 0x00400000u;
    float fe = cpm_u2f((jx >> 23) + (0x4B400000u - 127u)) - 12582912.0f;
    uint32_t i = (jx >> 18) & 31u;
    float m = cpm_u2f(ix + 0x3f800000u - (jx & 0xff800000u));
    float r = fmaf(m, tab[2 * i], -1.0f);
    float q = fmaf(r, 0.2f, -0.25f);
    q = fmaf(r, q, 0.3333333432674407958984375f);
    q = fmaf(r, q, -0.5f);
    float lp = fmaf(r * r, q, r);
    return fmaf(fe, 0.693147182464599609375f, tab[2 * i + 1]) + lp;
}
/* host form (the oracle, the OpenCL-on-host shim of oracle/_ref) */
static const float cpm_nlog_table_host[64] = {CPM_NLOG_TABLE};
static inline float cpm_native_logf(float x) { return cpm_native_logf_tab(x, cpm_nlog_table_host); }

/* exp(x) for x in [-87, 0] (the path uses it for the transmittance of one ray-march step).
 * k = rint(x log2 e), r = x - k ln2 (two fma steps with the 16-bit-exact LN2_HI), degree-6 Taylor kernel
 * on |r| <= 0.347 (next term r^7/5040 < 1.2e-7 relative), scaled by 2^k through the exponent field.
 * x < -87 returns 0, x > 0 is clamped to 0 (returns 1). */
CPM_HD float cpm_expf(float x) {
    if (!(x < 0.0f)) return 1.0f;
    if (x < -87.0f) return 0.0f;
    float kf = rintf(x * 1.44269502162933349609f);
    float r = fmaf(-kf, CPM_LN2_HI, x);
    r = fmaf(-kf, CPM_LN2_LO, r);
    float p = 1.3888889225e-3f;                 /* 1/720 */
    p = fmaf(p, r, 8.3333337680e-3f);           /* 1/120 */
    p = fmaf(p, r, 4.1666667908e-2f);           /* 1/24  */
    p = fmaf(p, r, 1.6666667163e-1f);           /* 1/6   */
    p = fmaf(p, r, 0.5f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    int k = (int)kf;                            /* -126 <= k <= 0 */
    return p * cpm_u2f((uint32_t)(k + 127) << 23);
}

/* exp(x) for |x| <= 87: the same kernel as cpm_expf with both signs of k (used by cpm_powf); <= 3 ulp. */
CPM_HD float cpm_expf_sym(float x) {
    if (x < -87.0f) return 0.0f;
    if (x > 87.0f) x = 87.0f;
    float kf = rintf(x * 1.44269502162933349609f);
    float r = fmaf(-kf, CPM_LN2_HI, x);
    r = fmaf(-kf, CPM_LN2_LO, r);
    float p = 1.3888889225e-3f;
    p = fmaf(p, r, 8.3333337680e-3f);
    p = fmaf(p, r, 4.1666667908e-2f);
    p = fmaf(p, r, 1.6666667163e-1f);
    p = fmaf(p, r, 0.5f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    int k = (int)kf;                            /* -126 <= k <= 126 */
    return p * cpm_u2f((uint32_t)(k + 127) << 23);
}

/* x^p for x > 0 (x <= 0 returns 0): exp(p log x) with the product's rounding error carried into the
 * result (y + e = p*log x exactly, exp(y + e) = exp(y)(1 + e)).  Used by the sRGB -> linear step of
 * rgb2lab (p = 2.4, x in (0.09, 1.06]); measured there against float64 pow: <= 6 ulp. */
CPM_HD float cpm_powf(float x, float p) {
    if (!(x > 0.0f)) return 0.0f;
    float l = cpm_logf(x);
    float y = p * l;
    float e = fmaf(p, l, -y);
    float r = cpm_expf_sym(y);
    return fmaf(r, e, r);
}

/* cube root for x >= 0 (negative and NaN inputs are returned unchanged: not on the path).
 * Exponent/3 bit estimate (5 % error), two Halley steps t (2x + t^3) / (x + 2 t^3), one Newton step on the
 * fma residual t^3 - x.  Measured against float64 cbrt on [1e-30, 1e30]: <= 1 ulp.  Arguments above
 * 1e37 overflow in t^3-scale intermediates and are not on the path. */
CPM_HD float cpm_cbrtf(float x) {
    if (!(x > 0.0f)) return x;
    float s = 1.0f;
    if (x < 1.17549435e-38f) {                  /* subnormal: scale by 2^24, result by 2^-8 */
        x = x * 16777216.0f;
        s = 0.00390625f;
    }
    float t = cpm_u2f(cpm_f2u(x) / 3u + 0x2a5137a0u);
    float r = (t * t) * t;
    t = t * (((x + x) + r) / (x + (r + r)));
    r = (t * t) * t;
    t = t * (((x + x) + r) / (x + (r + r)));
    float t2 = t * t;
    float res = fmaf(t2, t, -x);
    t = t - res / (3.0f * t2);
    return t * s;
}

/* sin and cos for |x| <= ~16 (path uses [-pi, 2pi]).  Cody-Waite reduction with two fma
 * steps, Taylor kernels on [-pi/4, pi/4] (next omitted term < 2.3e-9 relative). */
CPM_HD void cpm_sincosf(float x, float* sn, float* cs) {
    float kf = rintf(x * CPM_2OPI_F);
    float r = fmaf(-kf, CPM_PIO2_HI, x);
    r = fmaf(-kf, CPM_PIO2_LO, r);
    int k = (int)kf;
    float z = r * r;
    float ps = 2.7557319223985893e-06f;            /*  1/9! */
    ps = fmaf(ps, z, -1.9841269841269841e-04f);    /* -1/7! */
    ps = fmaf(ps, z, 8.3333333333333333e-03f);     /*  1/5! */
    ps = fmaf(ps, z, -1.6666666666666667e-01f);    /* -1/3! */
    float s = fmaf(r * z, ps, r);
    float pc = -2.7557319223985888e-07f;           /* -1/10! */
    pc = fmaf(pc, z, 2.4801587301587302e-05f);     /*  1/8!  */
    pc = fmaf(pc, z, -1.3888888888888889e-03f);    /* -1/6!  */
    pc = fmaf(pc, z, 4.1666666666666667e-02f);     /*  1/4!  */
    float c = fmaf(z * z, pc, fmaf(-0.5f, z, 1.0f));
    float so = (k & 1) ? c : s;
    float co = (k & 1) ? s : c;
    if (k & 2) so = -so;
    if ((k + 1) & 2) co = -co;
    *sn = so;
    *cs = co;
}

/* asin on [0, 0.5]: x + x z Q(z).  Coefficients: tools/fit_detmath.py (max rel err 2.7e-9). */
CPM_HD float cpm_asin_core(float x) {
    float z = x * x;
    float q = 0x1.14f3d2p-5f;
    q = fmaf(q, z, 0x1.17c832p-6f);
    q = fmaf(q, z, 0x1.fdd000p-6f);
    q = fmaf(q, z, 0x1.6d58c6p-5f);
    q = fmaf(q, z, 0x1.33343cp-4f);
    q = fmaf(q, z, 0x1.555554p-3f);
    return fmaf(x * z, q, x);
}

/* acos for x in [-1, 1] (callers clamp first, as ppm/photondata.cpp:109 does). */
CPM_HD float cpm_acosf(float x) {
    float ax = fabsf(x);
    if (ax <= 0.5f) {
        float a = cpm_asin_core(ax);
        a = (x < 0.0f) ? -a : a;
        return (CPM_PIO2_HI - a) + CPM_PIO2_LO;
    }
    float h = (1.0f - ax) * 0.5f;
    float r = sqrtf(h);
    float a2 = 2.0f * cpm_asin_core(r);
    return (x < 0.0f) ? (CPM_PI_F - a2) : a2;
}

/* atan for t >= 0, result in [0, pi/2]. */
CPM_HD float cpm_atan_pos(float t) {
    /* reduce to |u| <= tan(pi/8) */
    float base = 0.0f;
    float u = t;
    if (t > 2.41421356237309515f) { /* tan(3pi/8) */
        base = CPM_PIO2_HI;
        u = -1.0f / t;
    } else if (t > 0.41421356237309503f) { /* tan(pi/8) */
        base = CPM_PIO4_F;
        u = (t - 1.0f) / (t + 1.0f);
    }
    float z = u * u;
    float p = 0x1.9dffb4p-5f;
    p = fmaf(p, z, -0x1.615d26p-4f);
    p = fmaf(p, z, 0x1.c57f60p-4f);
    p = fmaf(p, z, -0x1.248a34p-3f);
    p = fmaf(p, z, 0x1.99997cp-3f);
    p = fmaf(p, z, -0x1.555556p-2f);
    float a = fmaf(u * z, p, u);
    return base + a;
}

/* atan2(y, x) with the C99 special cases that can occur for finite inputs. */
CPM_HD float cpm_atan2f(float y, float x) {
    float ay = fabsf(y), ax = fabsf(x);
    float a;
    if (ax == 0.0f && ay == 0.0f) {
        /* atan2(+-0, +0) = +-0 ; atan2(+-0, -0) = +-pi */
        a = (cpm_f2u(x) >> 31) ? CPM_PI_F : 0.0f;
        return (cpm_f2u(y) >> 31) ? -a : a;
    }
    if (ax == 0.0f) {
        a = CPM_PIO2_HI;
    } else {
        a = cpm_atan_pos(ay / ax);
        if (cpm_f2u(x) >> 31) a = CPM_PI_F - a;
    }
    return (cpm_f2u(y) >> 31) ? -a : a;
}

#endif /* CPM_DETMATH_H */
