/* orc_common.h -- small vector helpers of the oracle.  TEST INFRASTRUCTURE. */
#ifndef ORC_COMMON_H
#define ORC_COMMON_H
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "cpm_detmath.h"
#include "cpm_oracle.h"

typedef struct { float x, y, z; } v3;

static inline v3 v3_make(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline float v3_dot(v3 a, v3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
static inline v3 v3_cross(v3 a, v3 b) {
    return v3_make(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
/* origin + t*direction, one fused multiply-add per component (the oracle's own marchers: gather, ray caster) */
static inline v3 v3_madd(v3 o, float t, v3 d) { return v3_make(fmaf(t, d.x, o.x), fmaf(t, d.y, o.y), fmaf(t, d.z, o.z)); }
/* `origin + t*direction` as the reference's kernels WRITE it: a rounded product, then a rounded sum per component
 * (built with -ffp-contract=off).  This is what oracle/_ref/libcl_ref.so -- the reference's .cl files compiled with
 * strict IEEE evaluation -- computes, so the oracle matches it bit for bit (tests/test_ref_kernels.py). */
static inline v3 v3_ray(v3 o, float t, v3 d) { return v3_make(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z); }

static inline float lerpf_(float p, float q, float a) { return fmaf(a, q - p, p); }

/* MWC64X step + random_01, rng/cl/random.cl:58-95 */
typedef struct { uint32_t x, c; } random_state;
static inline float random_01(random_state* s) {
    uint32_t res = s->x ^ s->c;
    orc_rng_step(&s->x, &s->c);
    return res / 4294967295.0f;
}

/* decodeDirection / encodeDirection: host twin ppm/photondata.cpp:100-117 */
static inline v3 decodeDirection(float theta, float phi) {
    float st, ct, sp, cp;
    cpm_sincosf(theta, &st, &ct);
    cpm_sincosf(phi, &sp, &cp);
    return v3_make(st * cp, st * sp, ct);
}
static inline void encodeDirection(v3 d, float* theta, float* phi) {
    *theta = cpm_acosf(cpm_clamp(d.z, -1.0f, 1.0f));
    *phi = cpm_atan2f(d.y, d.x);
}

/* uniformSampleSphere (Inviwo shading/shadingmath.cl, restated) */
static inline v3 uniformSampleSphere(float u1, float u2) {
    float z = fmaf(-2.0f, u1, 1.0f);
    float r = sqrtf(cpm_fmax(0.0f, fmaf(-z, z, 1.0f)));
    float s, c;
    cpm_sincosf(CPM_2PI_F * u2, &s, &c);
    return v3_make(r * c, r * s, z);
}

/* rayBoxIntersection (Inviwo intersection/rayboxintersection.cl, restated): slab test */
static inline int rayBoxIntersection(const float* bmin, const float* bmax, v3 o, v3 d, float* t0, float* t1) {
    float ix = 1.0f / d.x, iy = 1.0f / d.y, iz = 1.0f / d.z;
    float ax = (bmin[0] - o.x) * ix, bx = (bmax[0] - o.x) * ix;
    float ay = (bmin[1] - o.y) * iy, by = (bmax[1] - o.y) * iy;
    float az = (bmin[2] - o.z) * iz, bz = (bmax[2] - o.z) * iz;
    float n = cpm_fmax(cpm_fmax(cpm_fmin(ax, bx), cpm_fmin(ay, by)), cpm_fmin(az, bz));
    float f = cpm_fmin(cpm_fmin(cpm_fmax(ax, bx), cpm_fmax(ay, by)), cpm_fmax(az, bz));
    *t0 = cpm_fmax(*t0, n);
    *t1 = cpm_fmin(*t1, f);
    return *t0 < *t1;
}
#endif
