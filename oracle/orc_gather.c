/* orc_gather.c -- oracle restatement of the photon-map gather path.  TEST INFRASTRUCTURE (see cpm_oracle.h).
 *
 * PARITY UNPINNED: no launched kernel of the reference computes cell keys of photons, cell ranges or a
 * ray-march gather.  What follows the reference is the estimator:
 * Epanechnikov kernel (ppm/cl/densityestimationkernel.cl:56-60), power * isotropicPhase * scale
 * (ppm/cl/photonstolightvolume.cl:160-165), evaluated per point as in the disabled
 * photonsToLightVolumeKernel (ppm/cl/photonstolightvolume.cl:81-134).  The oracle is written
 * independently of the CUDA path: it bins photons with its own counting sort instead of radix sort +
 * cell ranges + reorder, and the tests compare images with a stated tolerance. */
#include <stdlib.h>
#include <string.h>

#include "orc_common.h"

static int cell_coord(float p, float g, int n) { return (int)cpm_clamp(truncf(p * g), 0.0f, (float)(n - 1)); }

void orc_photon_cell_keys(const float* photons, size_t n, const int grid_dims[3], uint32_t* keys) {
    const int gx = grid_dims[0], gy = grid_dims[1], gz = grid_dims[2];
    for (size_t i = 0; i < n; ++i) {
        const float* p = photons + 8 * i;
        if (p[0] == FLT_MAX || p[1] == FLT_MAX || p[2] == FLT_MAX) {
            keys[i] = (uint32_t)gx * (uint32_t)gy * (uint32_t)gz;
        } else {
            int cx = cell_coord(p[0], (float)gx, gx), cy = cell_coord(p[1], (float)gy, gy), cz = cell_coord(p[2], (float)gz, gz);
            keys[i] = (uint32_t)cx + (uint32_t)gx * ((uint32_t)cy + (uint32_t)gy * (uint32_t)cz);
        }
    }
}

/* photon bins: offsets[c]..offsets[c+1] index `order`, photons of a cell in ascending record id */
typedef struct {
    uint32_t* offsets;
    uint32_t* order;
    uint32_t n_cells;
} bins_t;

static bins_t make_bins(const float* photons, size_t n, const int grid_dims[3]) {
    bins_t b;
    b.n_cells = (uint32_t)grid_dims[0] * (uint32_t)grid_dims[1] * (uint32_t)grid_dims[2];
    uint32_t* keys = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    orc_photon_cell_keys(photons, n, grid_dims, keys);
    b.offsets = (uint32_t*)calloc((size_t)b.n_cells + 2, sizeof(uint32_t));
    b.order = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    for (size_t i = 0; i < n; ++i) b.offsets[keys[i] + 1]++;
    for (uint32_t c = 0; c <= b.n_cells; ++c) b.offsets[c + 1] += b.offsets[c];
    uint32_t* cursor = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)b.n_cells + 1));
    memcpy(cursor, b.offsets, sizeof(uint32_t) * ((size_t)b.n_cells + 1));
    for (size_t i = 0; i < n; ++i) b.order[cursor[keys[i]]++] = (uint32_t)i;
    free(cursor);
    free(keys);
    return b;
}

static void free_bins(bins_t* b) {
    free(b->offsets);
    free(b->order);
}

static void gather_point(const bins_t* b, const float* photons, const int g[3], float radius, float x, float y, float z,
                         float e[3]) {
    int x0 = cell_coord(x - radius, (float)g[0], g[0]), x1 = cell_coord(x + radius, (float)g[0], g[0]);
    int y0 = cell_coord(y - radius, (float)g[1], g[1]), y1 = cell_coord(y + radius, (float)g[1], g[1]);
    int z0 = cell_coord(z - radius, (float)g[2], g[2]), z1 = cell_coord(z + radius, (float)g[2], g[2]);
    for (int cz = z0; cz <= z1; ++cz)
        for (int cy = y0; cy <= y1; ++cy)
            for (int cx = x0; cx <= x1; ++cx) {
                uint32_t c = (uint32_t)cx + (uint32_t)g[0] * ((uint32_t)cy + (uint32_t)g[1] * (uint32_t)cz);
                for (uint32_t j = b->offsets[c]; j < b->offsets[c + 1]; ++j) {
                    const float* p = photons + 8 * (size_t)b->order[j];
                    float dx = p[0] - x, dy = p[1] - y, dz = p[2] - z;
                    float dist = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
                    float xk = dist / radius;
                    if (xk <= 1.0f) {
                        float w = 0.75f * (1.0f - xk * xk);
                        e[0] = fmaf(p[3], w, e[0]);
                        e[1] = fmaf(p[4], w, e[1]);
                        e[2] = fmaf(p[5], w, e[2]);
                    }
                }
            }
}

void orc_gather_points(const orc_gather_params* P, const float* photons, size_t n_records, const float* points,
                       int n_points, float* irradiance) {
    bins_t b = make_bins(photons, n_records, P->grid_dims);
    const float s = CPM_INV_4PI_F * P->scale;
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n_points; ++i) {
        float e[3] = {0.f, 0.f, 0.f};
        gather_point(&b, photons, P->grid_dims, P->radius, points[3 * i], points[3 * i + 1], points[3 * i + 2], e);
        irradiance[3 * i] = e[0] * s;
        irradiance[3 * i + 1] = e[1] * s;
        irradiance[3 * i + 2] = e[2] * s;
    }
    free_bins(&b);
}

static void sample_tf_rgba(const float* tf, int width, float v, float c[4]) {
    float fwidth = (float)width;
    float u = fmaf(v, fwidth, -0.5f);
    float fu = floorf(u);
    float a = u - fu;
    int i0 = (int)cpm_clamp(fu, -1.0f, fwidth - 1.0f);
    int i1 = i0 + 1 < width - 1 ? i0 + 1 : width - 1;
    if (i0 < 0) i0 = 0;
    for (int k = 0; k < 4; ++k) c[k] = lerpf_(tf[4 * i0 + k], tf[4 * i1 + k], a);
}

void orc_gather_raymarch(const orc_volume* vol, const float* tf_rgba, int tf_width, const orc_gather_params* P,
                         const float* photons, size_t n_records, float* image) {
    bins_t b = make_bins(photons, n_records, P->grid_dims);
    const float s = CPM_INV_4PI_F * P->scale;
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < P->height; ++py)
        for (int px = 0; px < P->width; ++px) {
            float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
            float dx = fmaf(fy, P->cam_dv[0], fmaf(fx, P->cam_du[0], P->cam_dir00[0]));
            float dy = fmaf(fy, P->cam_dv[1], fmaf(fx, P->cam_du[1], P->cam_dir00[1]));
            float dz = fmaf(fy, P->cam_dv[2], fmaf(fx, P->cam_du[2], P->cam_dir00[2]));
            float inv = 1.0f / sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
            v3 d = v3_make(dx * inv, dy * inv, dz * inv);
            v3 o = v3_make(P->cam_origin[0], P->cam_origin[1], P->cam_origin[2]);
            float t0 = 0.0f, t1 = FLT_MAX;
            float L[3] = {0.f, 0.f, 0.f}, T = 1.0f;
            if (rayBoxIntersection(P->aabb_min, P->aabb_max, o, d, &t0, &t1)) {
                int k = 0;
                for (float t = fmaf(0.5f, P->step, t0); t < t1; ++k, t = fmaf((float)k + 0.5f, P->step, t0)) {
                    v3 x = v3_madd(o, t, d);
                    float v = orc_sample_volume(vol, x.x, x.y, x.z);
                    float c[4];
                    sample_tf_rgba(tf_rgba, tf_width, v, c);
                    if (c[3] > 0.0f) {
                        float e[3] = {0.f, 0.f, 0.f};
                        gather_point(&b, photons, P->grid_dims, P->radius, x.x, x.y, x.z, e);
                        float Ts = cpm_expf(-(c[3] * P->sigma_scale) * P->step);
                        float wgt = T * (1.0f - Ts);
                        for (int ch = 0; ch < 3; ++ch) L[ch] = fmaf(wgt * c[ch], e[ch] * s, L[ch]);
                        T *= Ts;
                        if (T < 1e-4f) break;
                    }
                }
            }
            float* out = image + 4 * ((size_t)py * P->width + px);
            out[0] = L[0]; out[1] = L[1]; out[2] = L[2]; out[3] = 1.0f - T;
        }
    free_bins(&b);
}

/* ---- final image from the light volume ------------------------------------------------------------------
 * The role Inviwo's LightingRaycaster plays after PhotonToLightVolumeProcessorCL in the workspace network
 * (ws:1178-1271; an Inviwo core processor, not part of the reference tree: PARITY UNPINNED).  Same camera,
 * samples and emission-absorption compositing as orc_gather_raymarch; the in-scattered radiance of a sample is
 * TF colour x light volume (trilinear, normalised coordinates, clamp-to-edge) instead of a photon gather. */
static float sample_lv(const float* lv, const int d[3], int nch, int ch, float px, float py, float pz) {
    float fx = (float)d[0], fy = (float)d[1], fz = (float)d[2];
    float u = fmaf(px, fx, -0.5f), v = fmaf(py, fy, -0.5f), w = fmaf(pz, fz, -0.5f);
    float fu = floorf(u), fv = floorf(v), fw = floorf(w);
    float a = u - fu, b = v - fv, c = w - fw;
    int i0 = (int)cpm_clamp(fu, -1.0f, fx - 1.0f), j0 = (int)cpm_clamp(fv, -1.0f, fy - 1.0f), k0 = (int)cpm_clamp(fw, -1.0f, fz - 1.0f);
    int i1 = i0 + 1 < d[0] - 1 ? i0 + 1 : d[0] - 1, j1 = j0 + 1 < d[1] - 1 ? j0 + 1 : d[1] - 1, k1 = k0 + 1 < d[2] - 1 ? k0 + 1 : d[2] - 1;
    if (i0 < 0) i0 = 0;
    if (j0 < 0) j0 = 0;
    if (k0 < 0) k0 = 0;
#define LV(i, j, k) lv[((size_t)(i) + (size_t)d[0] * ((size_t)(j) + (size_t)d[1] * (size_t)(k))) * nch + ch]
    float x00 = lerpf_(LV(i0, j0, k0), LV(i1, j0, k0), a), x10 = lerpf_(LV(i0, j1, k0), LV(i1, j1, k0), a);
    float x01 = lerpf_(LV(i0, j0, k1), LV(i1, j0, k1), a), x11 = lerpf_(LV(i0, j1, k1), LV(i1, j1, k1), a);
#undef LV
    float y0 = lerpf_(x00, x10, b), y1 = lerpf_(x01, x11, b);
    return lerpf_(y0, y1, c);
}

void orc_raycast_light_volume(const orc_volume* vol, const float* tf_rgba, int tf_width, const orc_gather_params* P,
                              const float* light_volume, const int lv_dims[3], int channels, float* image) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < P->height; ++py)
        for (int px = 0; px < P->width; ++px) {
            float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
            float dx = fmaf(fy, P->cam_dv[0], fmaf(fx, P->cam_du[0], P->cam_dir00[0]));
            float dy = fmaf(fy, P->cam_dv[1], fmaf(fx, P->cam_du[1], P->cam_dir00[1]));
            float dz = fmaf(fy, P->cam_dv[2], fmaf(fx, P->cam_du[2], P->cam_dir00[2]));
            float inv = 1.0f / sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
            v3 d = v3_make(dx * inv, dy * inv, dz * inv);
            v3 o = v3_make(P->cam_origin[0], P->cam_origin[1], P->cam_origin[2]);
            float t0 = 0.0f, t1 = FLT_MAX;
            float L[3] = {0.f, 0.f, 0.f}, T = 1.0f;
            if (rayBoxIntersection(P->aabb_min, P->aabb_max, o, d, &t0, &t1)) {
                int k = 0;
                for (float t = fmaf(0.5f, P->step, t0); t < t1; ++k, t = fmaf((float)k + 0.5f, P->step, t0)) {
                    v3 x = v3_madd(o, t, d);
                    float v = orc_sample_volume(vol, x.x, x.y, x.z);
                    float c[4];
                    sample_tf_rgba(tf_rgba, tf_width, v, c);
                    if (c[3] > 0.0f) {
                        float e[3];
                        e[0] = sample_lv(light_volume, lv_dims, channels, 0, x.x, x.y, x.z);
                        e[1] = channels == 4 ? sample_lv(light_volume, lv_dims, channels, 1, x.x, x.y, x.z) : e[0];
                        e[2] = channels == 4 ? sample_lv(light_volume, lv_dims, channels, 2, x.x, x.y, x.z) : e[0];
                        float Ts = cpm_expf(-(c[3] * P->sigma_scale) * P->step);
                        float wgt = T * (1.0f - Ts);
                        for (int ch = 0; ch < 3; ++ch) L[ch] = fmaf(wgt * c[ch], e[ch], L[ch]);
                        T *= Ts;
                        if (T < 1e-4f) break;
                    }
                }
            }
            float* out = image + 4 * ((size_t)py * P->width + px);
            out[0] = L[0]; out[1] = L[1]; out[2] = L[2]; out[3] = 1.0f - T;
        }
}
