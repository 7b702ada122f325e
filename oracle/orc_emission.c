/* orc_emission.c -- oracle restatement of sample generation, light sampling and the CPU light
 * plane fit.  TEST INFRASTRUCTURE (see cpm_oracle.h). */
#include <stdlib.h>
#include <string.h>

#include "orc_common.h"

/* isc/cl/uniformsamplegenerator2d.cl:35-52 */
void orc_sample_uniform2d(float nx, float ny, int n, float* out) {
    for (int threadId = 0; threadId < n; ++threadId) {
        float cx = fmodf((float)threadId, nx);
        float cy = (float)threadId / nx; /* not floored, as in the reference */
        out[4 * threadId + 0] = (0.5f + cx) / nx;
        out[4 * threadId + 1] = (0.5f + cy) / ny;
        out[4 * threadId + 2] = 0.0f;
        out[4 * threadId + 3] = 1.0f; /* pdf */
    }
}

/* lcl/cl/directionallightsampler.cl:38-63 + writeLightSample (lcl/cl/datastructures/lightsample.cl:80-89) */
void orc_light_sample_directional(const float* samples, const float radiance[3], const float dir[3],
                                  const float origin[3], const float u[3], const float v[3], float area, int n,
                                  float* out) {
    float theta, phi;
    encodeDirection(v3_make(dir[0], dir[1], dir[2]), &theta, &phi);
    for (int i = 0; i < n; ++i) {
        const float* s = samples + 4 * i;
        float* o = out + 8 * (size_t)i;
        /* planeOrigin + planeTangentU * s.x + planeTangentV * s.y, evaluated as written (rounded products and sums) */
        for (int k = 0; k < 3; ++k) o[k] = (origin[k] + u[k] * s[0]) + v[k] * s[1];
        float pdf = s[3] / area;
        for (int k = 0; k < 3; ++k) o[3 + k] = radiance[k] / pdf;
        o[6] = theta;
        o[7] = phi;
    }
}

/* isc/cl/light/light.cl:84-92 (LIGHT_POINT branch of sampleLight) */
void orc_light_sample_point(const float* samples, const float radiance[3], const float pos[3], int n, float* out) {
    for (int i = 0; i < n; ++i) {
        const float* s = samples + 4 * i;
        float* o = out + 8 * (size_t)i;
        v3 d = uniformSampleSphere(s[0], s[1]);
        d = v3_make(-d.x, -d.y, -d.z);
        const float pdf = CPM_INV_4PI_F;
        for (int k = 0; k < 3; ++k) o[k] = pos[k];
        for (int k = 0; k < 3; ++k) o[3 + k] = radiance[k] / pdf;
        encodeDirection(d, &o[6], &o[7]);
    }
}

/* lcl/cl/intersection/lightsamplemeshintersection.cl:37-59 with rayMeshIntersection restated as
 * Moeller-Trumbore over the index triples (Inviwo intersection/raymeshintersection.cl is not
 * vendored: parity unpinned). */
void orc_light_mesh_intersect(const float* vertices, const int32_t* indices, int n_indices,
                              const float* light_samples, int n, float* out) {
    int n_tri = n_indices / 3;
    for (int i = 0; i < n; ++i) {
        const float* ls = light_samples + 8 * (size_t)i;
        v3 o = v3_make(ls[0], ls[1], ls[2]);
        v3 d = decodeDirection(ls[6], ls[7]);
        float tn = FLT_MAX, tf = -FLT_MAX;
        int hit = 0;
        for (int t = 0; t < n_tri; ++t) {
            const float* p0 = vertices + 3 * (size_t)indices[3 * t];
            const float* p1 = vertices + 3 * (size_t)indices[3 * t + 1];
            const float* p2 = vertices + 3 * (size_t)indices[3 * t + 2];
            v3 v0 = v3_make(p0[0], p0[1], p0[2]);
            v3 e1 = v3_make(p1[0] - v0.x, p1[1] - v0.y, p1[2] - v0.z);
            v3 e2 = v3_make(p2[0] - v0.x, p2[1] - v0.y, p2[2] - v0.z);
            v3 p = v3_cross(d, e2);
            float det = v3_dot(e1, p);
            if (fabsf(det) < 1e-12f) continue;
            float inv = 1.0f / det;
            v3 tv = v3_sub(o, v0);
            float u = v3_dot(tv, p) * inv;
            if (u < 0.0f || u > 1.0f) continue;
            v3 q = v3_cross(tv, e1);
            float v = v3_dot(d, q) * inv;
            if (v < 0.0f || u + v > 1.0f) continue;
            float tt = v3_dot(e2, q) * inv;
            tn = cpm_fmin(tn, tt);
            tf = cpm_fmax(tf, tt);
            hit = 1;
        }
        float t0 = 0.0f, t1 = FLT_MAX;
        if (hit) {
            t0 = cpm_fmax(t0, tn);
            t1 = cpm_fmin(t1, tf);
            hit = t0 < t1;
        }
        if (!hit) {
            t0 = 0.0f;
            t1 = -1.0f;
        }
        out[2 * (size_t)i] = t0;
        out[2 * (size_t)i + 1] = t1;
    }
}

/* ---- light plane fit (CPU in the reference too) ------------------------------------------ */
typedef struct { float x, y; } v2;

static int cmp_v2(const void* a, const void* b) {
    const v2 *p = (const v2*)a, *q = (const v2*)b;
    if (p->x != q->x) return p->x < q->x ? -1 : 1;
    if (p->y != q->y) return p->y < q->y ? -1 : 1;
    return 0;
}
/* lcl/convexhull2d.cpp:53-55 */
static float isPointLeftOfLine(v2 p0, v2 p1, v2 pt) { return (p1.x - p0.x) * (pt.y - p0.y) - (pt.x - p0.x) * (p1.y - p0.y); }

/* lcl/convexhull2d.cpp:38-130.  hull must hold 2*n+2 points; returns the hull size. */
static int convexHull2D(v2* points, int n, v2* hull) {
    qsort(points, (size_t)n, sizeof(v2), cmp_v2);
    if (n < 4) {
        memcpy(hull, points, (size_t)n * sizeof(v2));
        return n;
    }
    int minXMinYId = 0, minXMaxYId = 1;
    for (; minXMaxYId < n; ++minXMaxYId)
        if (points[0].x != points[minXMaxYId].x) break;
    --minXMaxYId;
    int h = 0;
    if (minXMaxYId == n - 1) { /* all x equal */
        hull[h++] = points[minXMinYId];
        if (points[minXMaxYId].y != points[minXMinYId].y) hull[h++] = points[minXMaxYId];
        hull[h++] = points[minXMinYId];
        return h;
    }
    int maxXMinYId = n - 1, maxXMaxYId = n - 2;
    for (; maxXMaxYId >= 0; --maxXMaxYId)
        if (points[n - 1].x > points[maxXMaxYId].x) break;
    ++maxXMaxYId;
    hull[h++] = points[minXMinYId];
    for (int i = minXMaxYId + 1; i <= maxXMinYId; ++i) {
        if (isPointLeftOfLine(points[minXMinYId], points[maxXMinYId], points[i]) >= 0 && i < maxXMinYId) continue;
        while (h >= 2) {
            if (isPointLeftOfLine(hull[h - 2], hull[h - 1], points[i]) > 0) break;
            --h;
        }
        hull[h++] = points[i];
    }
    if (maxXMaxYId != maxXMinYId) hull[h++] = points[maxXMaxYId];
    int bottomHull = h - 1;
    for (int i = maxXMaxYId; i > minXMaxYId; --i) {
        if (isPointLeftOfLine(points[maxXMaxYId], points[minXMaxYId], points[i]) >= 0 && i > minXMaxYId) continue;
        while (h - bottomHull >= 2) {
            if (isPointLeftOfLine(hull[h - 2], hull[h - 1], points[i]) > 0) break;
            --h;
        }
        hull[h++] = points[i];
    }
    if (minXMaxYId != minXMinYId) hull[h++] = points[maxXMinYId];
    return h;
}

/* glm::normalize(v) = v * inversesqrt(dot(v, v)), inversesqrt(x) = 1 / sqrt(x) (GLM 0.9.7, Inviwo's ext/glm) */
/* exposed for the golden-vector test of the hull alone; hull_out holds 2*n+2 points, returns the hull size */
int orc_convex_hull2d(const float* pts, int n, float* hull_out) {
    v2* p = (v2*)malloc(sizeof(v2) * (size_t)(n > 0 ? n : 1));
    memcpy(p, pts, sizeof(v2) * (size_t)n);
    int h = convexHull2D(p, n, (v2*)hull_out);
    free(p);
    return h;
}

static v3 v3_normalize(v3 a) {
    float inv = 1.0f / sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
    return v3_make(a.x * inv, a.y * inv, a.z * inv);
}
static float dot_glm(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static v3 cross_glm(v3 a, v3 b) { return v3_make(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
/* Inviwo Plane::projectPoint: p - dot(p - point, n) * n */
static v3 plane_project(v3 point, v3 normal, v3 p) {
    float d = dot_glm(v3_sub(p, point), normal);
    return v3_make(p.x - d * normal.x, p.y - d * normal.y, p.z - d * normal.z);
}

/* lcl/orientedboundingbox2d.cpp:80-100 (fitPlaneAlignedOrientedBoundingBox2D) with
 * mimumBoundingRectangle (:40-78) and projectPointsOnPlane (lcl/pointplaneprojection.cpp:39-54). */
void orc_fit_light_plane(const float* pts, int n_points, const float plane_point[3], const float plane_normal[3],
                         float out[9]) {
    v3 P = v3_make(plane_point[0], plane_point[1], plane_point[2]);
    v3 N = v3_make(plane_normal[0], plane_normal[1], plane_normal[2]);
    v3 u, v;
    if (fabsf(N.x) > fabsf(N.y))
        u = v3_normalize(v3_sub(plane_project(P, N, v3_make(1.f, 0.f, 0.f)), P));
    else
        u = v3_normalize(v3_sub(plane_project(P, N, v3_make(0.f, 1.f, 0.f)), P));
    v = v3_normalize(cross_glm(N, u));
    v2* proj = (v2*)malloc(sizeof(v2) * (size_t)n_points);
    v2* hull = (v2*)malloc(sizeof(v2) * (size_t)(2 * n_points + 4));
    float d = dot_glm(N, P);
    for (int i = 0; i < n_points; ++i) {
        v3 e = v3_make(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
        float dist = dot_glm(N, e) - d;
        v3 pp = v3_make(e.x - dist * N.x, e.y - dist * N.y, e.z - dist * N.z);
        v3 op = v3_sub(pp, P);
        proj[i].x = dot_glm(u, op);
        proj[i].y = dot_glm(v, op);
    }
    int nh = convexHull2D(proj, n_points, hull);
    float minArea = FLT_MAX;
    v2 origin = {0, 0}, bu = {0, 0}, bv = {0, 0};
    for (int i = 0, j = nh - 1; i < nh; j = i, ++i) {
        float ex = hull[i].x - hull[j].x, ey = hull[i].y - hull[j].y;
        float inv = 1.0f / sqrtf(ex * ex + ey * ey); /* glm::normalize */
        v2 e0 = {ex * inv, ey * inv};
        if (e0.x != e0.x || e0.y != e0.y) continue;
        v2 e1 = {-e0.y, e0.x};
        float min0 = 0.f, min1 = 0.f, max0 = 0.f, max1 = 0.f;
        for (int k = 0; k < nh; ++k) {
            float dx = hull[k].x - hull[j].x, dy = hull[k].y - hull[j].y;
            float dot = dx * e0.x + dy * e0.y;
            min0 = fminf(min0, dot);
            max0 = fmaxf(max0, dot);
            dot = dx * e1.x + dy * e1.y;
            min1 = fminf(min1, dot);
            max1 = fmaxf(max1, dot);
        }
        float area = (max0 - min0) * (max1 - min1);
        if (area < minArea) {
            minArea = area;
            float m0 = fminf(min0, 0.f), m1 = fminf(min1, 0.f);
            origin.x = hull[j].x + m0 * e0.x + m1 * e1.x;
            origin.y = hull[j].y + m0 * e0.y + m1 * e1.y;
            bu.x = e0.x * (max0 - min0); bu.y = e0.y * (max0 - min0);
            bv.x = e1.x * (max1 - min1); bv.y = e1.y * (max1 - min1);
        }
    }
    const float uu[3] = {u.x, u.y, u.z}, vv[3] = {v.x, v.y, v.z};
    for (int k = 0; k < 3; ++k) {
        out[k] = plane_point[k] + origin.x * uu[k] + origin.y * vv[k];
        out[3 + k] = bu.x * uu[k] + bu.y * vv[k];
        out[6 + k] = bv.x * uu[k] + bv.y * vv[k];
    }
    free(proj);
    free(hull);
}

/* host evaluation of include/cpm_detmath.h for the host-vs-device bit-equality test */
void orc_selftest_math(int fn, const float* x, const float* y, float* out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        float a = x[i], b = y ? y[i] : 0.0f, r, s, c;
        switch (fn) {
            case 0: r = cpm_logf(a); break;
            case 1: cpm_sincosf(a, &s, &c); r = s; break;
            case 2: cpm_sincosf(a, &s, &c); r = c; break;
            case 3: r = cpm_acosf(a); break;
            case 4: r = cpm_atan2f(a, b); break;
            case 5: r = a / 255.0f; break;
            case 6: r = a / 65535.0f; break;
            case 7: r = cpm_powf(a, b); break;
            case 8: r = cpm_cbrtf(a); break;
            case 9: r = cpm_expf_sym(a); break;
            case 10: r = cpm_native_logf(a); break;
            default: r = 0.0f;
        }
        out[i] = r;
    }
}
