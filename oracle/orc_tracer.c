/* orc_tracer.c -- oracle restatement of the delta-tracking photon tracer.
 * TEST INFRASTRUCTURE (see cpm_oracle.h).
 *
 * Follows ppm/cl/photontracer.cl:69-216 and ppm/cl/transmittance.cl:126-144 statement by
 * statement.  Image reads follow OpenCL 1.2 section 8.2 (normalised coordinates, clamp to edge,
 * linear filter) because Inviwo's samplers.cl is not vendored. */
#include <omp.h>

#include "orc_common.h"

int orc_num_threads(void) { return omp_get_max_threads(); }
/* launchers such as torchrun export OMP_NUM_THREADS=1: the CPU arm of bench.py sets its thread count explicitly */
void orc_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

static float fetch(const orc_volume* v, int i, int j, int k) {
    size_t idx = ((size_t)k * v->dims[1] + (size_t)j) * v->dims[0] + (size_t)i;
    switch (v->format) {
        case 0: return (float)((const uint8_t*)v->data)[idx] / 255.0f;   /* CL_UNORM_INT8  */
        case 1: return (float)((const uint16_t*)v->data)[idx] / 65535.0f; /* CL_UNORM_INT16 */
        default: return ((const float*)v->data)[idx];
    }
}

static int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* getNormalizedVoxel(volumeTex, volumeParams, pos).x */
float orc_sample_volume(const orc_volume* V, float px, float py, float pz) {
    float fx = (float)V->dims[0], fy = (float)V->dims[1], fz = (float)V->dims[2];
    float u = fmaf(px, fx, -0.5f), v = fmaf(py, fy, -0.5f), w = fmaf(pz, fz, -0.5f);
    float fu = floorf(u), fv = floorf(v), fw = floorf(w);
    float a = u - fu, b = v - fv, c = w - fw;
    int i0 = (int)cpm_clamp(fu, -1.0f, fx - 1.0f);
    int j0 = (int)cpm_clamp(fv, -1.0f, fy - 1.0f);
    int k0 = (int)cpm_clamp(fw, -1.0f, fz - 1.0f);
    int i1 = clampi(i0 + 1, 0, V->dims[0] - 1), j1 = clampi(j0 + 1, 0, V->dims[1] - 1),
        k1 = clampi(k0 + 1, 0, V->dims[2] - 1);
    i0 = clampi(i0, 0, V->dims[0] - 1);
    j0 = clampi(j0, 0, V->dims[1] - 1);
    k0 = clampi(k0, 0, V->dims[2] - 1);
    float x00 = lerpf_(fetch(V, i0, j0, k0), fetch(V, i1, j0, k0), a);
    float x10 = lerpf_(fetch(V, i0, j1, k0), fetch(V, i1, j1, k0), a);
    float x01 = lerpf_(fetch(V, i0, j0, k1), fetch(V, i1, j0, k1), a);
    float x11 = lerpf_(fetch(V, i0, j1, k1), fetch(V, i1, j1, k1), a);
    float y0 = lerpf_(x00, x10, b), y1 = lerpf_(x01, x11, b);
    float val = lerpf_(y0, y1, c);
    return (val + V->offset) * V->scale;
}

/* read_imagef(tfData, smpNormClampEdgeLinear, (float2)(v, 0.5f)).w */
float orc_sample_tf_alpha(const float* tf, int width, float v) {
    float fw = (float)width;
    float u = fmaf(v, fw, -0.5f);
    float fu = floorf(u);
    float a = u - fu;
    int i0 = (int)cpm_clamp(fu, -1.0f, fw - 1.0f);
    int i1 = clampi(i0 + 1, 0, width - 1);
    i0 = clampi(i0, 0, width - 1);
    return lerpf_(tf[4 * i0 + 3], tf[4 * i1 + 3], a);
}

/* ppm/cl/transmittance.cl:126-144 */
static float woodcockTracking(const orc_volume* vol, const float* tf, int tfw, v3 origin, v3 direction, float tStart,
                              float tEnd, float tauMax, random_state* rs, unsigned long long* tests) {
    float invTauMaxSampleBaseInterval = 1.f / (tauMax * 150.f);
    float invTauMax = 1.f / tauMax;
    float t = tStart;
    float opacity;
    float r;
    do {
        t += -cpm_native_logf(random_01(rs)) * invTauMaxSampleBaseInterval;
        v3 pos = v3_ray(origin, t, direction);
        float volumeSample = orc_sample_volume(vol, pos.x, pos.y, pos.z);
        opacity = orc_sample_tf_alpha(tf, tfw, volumeSample);
        r = random_01(rs);
        ++*tests;
    } while (r >= opacity * invTauMax && t <= tEnd);
    return t;
}

/* sampleShadingFunction for the two phase functions the oracle restates */
static v3 samplePhase(int phase, const float* material, v3 wi, float u1, float u2) {
    if (phase != 1) return uniformSampleSphere(u1, u2);
    float g = material[0], ct;
    if (fabsf(g) < 1e-3f) {
        ct = fmaf(-2.0f, u1, 1.0f);
    } else {
        float q = (1.0f - g * g) / fmaf(2.0f * g, u1, 1.0f - g);
        ct = (1.0f + g * g - q * q) / (2.0f * g);
    }
    ct = cpm_clamp(ct, -1.0f, 1.0f);
    float st = sqrtf(cpm_fmax(0.0f, fmaf(-ct, ct, 1.0f)));
    float sp, cp;
    cpm_sincosf(CPM_2PI_F * u2, &sp, &cp);
    v3 v2_;
    if (fabsf(wi.x) > fabsf(wi.y)) {
        float inv = 1.0f / sqrtf(fmaf(wi.x, wi.x, wi.z * wi.z));
        v2_ = v3_make(-wi.z * inv, 0.0f, wi.x * inv);
    } else {
        float inv = 1.0f / sqrtf(fmaf(wi.y, wi.y, wi.z * wi.z));
        v2_ = v3_make(0.0f, wi.z * inv, -wi.y * inv);
    }
    v3 v3_ = v3_cross(wi, v2_);
    float a = st * cp, b = st * sp;
    return v3_make(fmaf(a, v2_.x, fmaf(b, v3_.x, ct * wi.x)), fmaf(a, v2_.y, fmaf(b, v3_.y, ct * wi.y)),
                   fmaf(a, v2_.z, fmaf(b, v3_.z, ct * wi.z)));
}

static void writePhoton(float* photons, size_t id, v3 o, float pr, float pg, float pb, float th, float ph) {
    float* p = photons + 8 * id;
    p[0] = o.x; p[1] = o.y; p[2] = o.z; p[3] = pr; p[4] = pg; p[5] = pb; p[6] = th; p[7] = ph;
}

static unsigned long long trace_one(const orc_volume* vol, const float* tf, int tfw, const orc_trace_params* P,
                                    const float* lightSamples, const float* isect, int threadId, float* photons,
                                    uint32_t* rng) {
    unsigned long long tests = 0;
    random_state rs = {rng[2 * (size_t)(P->photon_offset + threadId)], rng[2 * (size_t)(P->photon_offset + threadId) + 1]};
    unsigned nInteractions = 0;
    const unsigned maxInteractions = (unsigned)P->max_interactions;
    const float* ls = lightSamples + 8 * (size_t)threadId;
    v3 origin = v3_make(ls[0], ls[1], ls[2]);
    float fmaxi = (float)maxInteractions;
    float pr = ls[3] / fmaxi, pg = ls[4] / fmaxi, pb = ls[5] / fmaxi; /* photontracer.cl:129 */
    v3 direction = decodeDirection(ls[6], ls[7]);
    float tStart = isect[2 * (size_t)threadId], tEnd = isect[2 * (size_t)threadId + 1];
    int scatterEvent = tStart < tEnd;

    if (P->flags & 2u) { /* NO_SINGLE_SCATTERING, photontracer.cl:143-157 */
        float t = woodcockTracking(vol, tf, tfw, origin, direction, tStart, tEnd, 1.f, &rs, &tests);
        if (scatterEvent) {
            origin = v3_ray(origin, t, direction);
            tStart = 0.f;
            tEnd = FLT_MAX;
            float u1 = random_01(&rs), u2 = random_01(&rs);
            direction = samplePhase(P->phase_function, P->material, direction, u1, u2);
            scatterEvent = rayBoxIntersection(P->aabb_min, P->aabb_max, origin, direction, &tStart, &tEnd);
            pr = pr / CPM_INV_4PI_F; pg = pg / CPM_INV_4PI_F; pb = pb / CPM_INV_4PI_F;
            tStart += 0.5f * P->step_size;
        }
    }
    while (scatterEvent) {
        float t = woodcockTracking(vol, tf, tfw, origin, direction, tStart, tEnd, 1.f, &rs, &tests);
        scatterEvent = t <= tEnd;
        if (scatterEvent) {
            origin = v3_ray(origin, t, direction);
            size_t photonId = (size_t)P->photon_offset + (size_t)nInteractions * P->total_photons + threadId;
            float th, ph;
            encodeDirection(direction, &th, &ph);
            float volumeSample = orc_sample_volume(vol, origin.x, origin.y, origin.z);
            float colorW = orc_sample_tf_alpha(tf, tfw, volumeSample);
            float scatteringW = colorW; /* same layer bound twice: ppm/photontracercl.cpp:150-151 */
            float scatteringAlbedo = scatteringW / (scatteringW + colorW);
            float den = cpm_fmax(colorW, 0.01f);
            pr = pr / den; pg = pg / den; pb = pb / den;
            ++nInteractions;
            if (nInteractions < maxInteractions && random_01(&rs) < scatteringAlbedo) {
                pr *= scatteringAlbedo; pg *= scatteringAlbedo; pb *= scatteringAlbedo;
                writePhoton(photons, photonId, origin, pr, pg, pb, th, ph);
                tStart = 0.f;
                tEnd = FLT_MAX;
                float u1 = random_01(&rs), u2 = random_01(&rs);
                direction = samplePhase(P->phase_function, P->material, direction, u1, u2);
                scatterEvent = rayBoxIntersection(P->aabb_min, P->aabb_max, origin, direction, &tStart, &tEnd);
                tStart += 0.5f * P->step_size;
            } else {
                writePhoton(photons, photonId, origin, pr, pg, pb, th, ph);
                pr = pg = pb = FLT_MAX;
                scatterEvent = 0;
            }
        }
    }
    float th, ph;
    encodeDirection(direction, &th, &ph);
    for (unsigned i = nInteractions; i < maxInteractions; ++i) {
        size_t photonId = (size_t)P->photon_offset + (size_t)i * P->total_photons + threadId;
        float* p = photons + 8 * photonId;
        p[0] = FLT_MAX; p[1] = FLT_MAX; p[2] = FLT_MAX; p[3] = pr; p[4] = FLT_MAX; p[5] = FLT_MAX; p[6] = th; p[7] = ph;
    }
    if (P->flags & 1u) { /* PROGRESSIVE_PHOTON_MAPPING */
        rng[2 * (size_t)(P->photon_offset + threadId)] = rs.x;
        rng[2 * (size_t)(P->photon_offset + threadId) + 1] = rs.c;
    }
    return tests;
}

unsigned long long orc_trace_photons(const orc_volume* vol, const float* tf, int tfw, const orc_trace_params* P,
                                     const float* lightSamples, const float* isect, const uint32_t* recompute,
                                     int n_recompute, float* photons, uint32_t* rng, int n_threads) {
    unsigned long long total = 0;
    int n_work = recompute ? n_recompute : P->n_light_samples;
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : total) num_threads(n_threads)
    for (int g = 0; g < n_work; ++g) {
        int threadId = g;
        if (recompute) {
            threadId = (int)recompute[g] - P->photon_offset;
            if (threadId < 0 || threadId >= P->n_light_samples) continue;
        }
        total += trace_one(vol, tf, tfw, P, lightSamples, isect, threadId, photons, rng);
    }
    return total;
}
