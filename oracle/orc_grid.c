/* orc_grid.c -- oracle restatement of the uniform-grid operators: min/max bricks, inter-step
 * difference bricks, importance classification, the DDA-based re-computation detector, the
 * light-sample cell hash and the cell-range build.  TEST INFRASTRUCTURE (see cpm_oracle.h). */
#include <stdlib.h>
#include <string.h>

#include "orc_common.h"

static float voxel_norm(const orc_volume* v, size_t idx) {
    float t;
    switch (v->format) {
        case 0: t = (float)((const uint8_t*)v->data)[idx] / 255.0f; break;
        case 1: t = (float)((const uint16_t*)v->data)[idx] / 65535.0f; break;
        default: t = ((const float*)v->data)[idx];
    }
    return (t + v->offset) * v->scale; /* getNormalizedVoxelUnorm */
}

/* writeImageVec2UInt16f is un-vendored: restated as convert_ushort_sat_rte(v * 65535) */
static uint16_t to_u16(float v) { return (uint16_t)rintf(cpm_clamp(v * 65535.0f, 0.0f, 65535.0f)); }

/* ugc/cl/uniformgrid/volumeminmax.cl:33-61; outDim = ceil(dim / region)
 * (ugc/processors/volumeminmaxclprocessor.cpp:151) */
void orc_volume_minmax(const orc_volume* vol, int region, uint16_t* out /* ushort2 per cell */) {
    int nx = vol->dims[0], ny = vol->dims[1], nz = vol->dims[2];
    int ox = (nx + region - 1) / region, oy = (ny + region - 1) / region, oz = (nz + region - 1) / region;
#pragma omp parallel for collapse(2) schedule(static)
    for (int gz = 0; gz < oz; ++gz)
        for (int gy = 0; gy < oy; ++gy)
            for (int gx = 0; gx < ox; ++gx) {
                float mn = FLT_MAX, mx = 0.0f; /* (float2)(FLT_MAX, 0) */
                int ex = gx * region + region < nx ? gx * region + region : nx;
                int ey = gy * region + region < ny ? gy * region + region : ny;
                int ez = gz * region + region < nz ? gz * region + region : nz;
                for (int z = gz * region; z < ez; ++z)
                    for (int y = gy * region; y < ey; ++y)
                        for (int x = gx * region; x < ex; ++x) {
                            float value = voxel_norm(vol, ((size_t)z * ny + y) * nx + x);
                            mn = cpm_fmin(mn, value);
                            mx = cpm_fmax(mx, value);
                        }
                size_t o = ((size_t)gz * oy + gy) * ox + gx;
                out[2 * o] = to_u16(mn);
                out[2 * o + 1] = to_u16(mx);
            }
}

/* ugc/processors/dynamicvolumedifferenceanalysis.h:96-151 (+ .cpp:60-104): per brick
 * (sum |scaling * (B - A)| / region^3 - range.x) / (range.y - range.x), double accumulation,
 * divides by the FULL region^3 even for clipped edge bricks (.h:147). */
void orc_volume_diff_bricks(const orc_volume* a, const orc_volume* b, int region, double data_scaling,
                            double range_min, double range_max, float* out) {
    int nx = a->dims[0], ny = a->dims[1], nz = a->dims[2];
    int ox = (nx + region - 1) / region, oy = (ny + region - 1) / region, oz = (nz + region - 1) / region;
#pragma omp parallel for collapse(2) schedule(static)
    for (int gz = 0; gz < oz; ++gz)
        for (int gy = 0; gy < oy; ++gy)
            for (int gx = 0; gx < ox; ++gx) {
                int ex = gx * region + region < nx ? gx * region + region : nx;
                int ey = gy * region + region < ny ? gy * region + region : ny;
                int ez = gz * region + region < nz ? gz * region + region : nz;
                double absDiffSum = 0.0;
                for (int z = gz * region; z < ez; ++z)
                    for (int y = gy * region; y < ey; ++y)
                        for (int x = gx * region; x < ex; ++x) {
                            size_t p = ((size_t)z * ny + y) * nx + x;
                            double va, vb;
                            switch (a->format) {
                                case 0: va = ((const uint8_t*)a->data)[p]; vb = ((const uint8_t*)b->data)[p]; break;
                                case 1: va = ((const uint16_t*)a->data)[p]; vb = ((const uint16_t*)b->data)[p]; break;
                                default: va = ((const float*)a->data)[p]; vb = ((const float*)b->data)[p];
                            }
                            absDiffSum += fabs(data_scaling * (vb - va));
                        }
                double r3 = (double)region * region * region;
                out[((size_t)gz * oy + gy) * ox + gx] = (float)((absDiffSum / r3 - range_min) / (range_max - range_min));
            }
}

/* ---- importance classification: isc/cl/minmaxuniformgrid3dimportance.cl ---------------------- */
typedef struct { float x, y, z, w; } v4;
static v4 v4_mix(v4 a, v4 b, float t) { /* mix(x,y,a) = x + (y-x)*a */
    v4 r = {fmaf(b.x - a.x, t, a.x), fmaf(b.y - a.y, t, a.y), fmaf(b.z - a.z, t, a.z), fmaf(b.w - a.w, t, a.w)};
    return r;
}
static v4 v4_min(v4 a, v4 b) { v4 r = {cpm_fmin(a.x, b.x), cpm_fmin(a.y, b.y), cpm_fmin(a.z, b.z), cpm_fmin(a.w, b.w)}; return r; }
static v4 v4_max(v4 a, v4 b) { v4 r = {cpm_fmax(a.x, b.x), cpm_fmax(a.y, b.y), cpm_fmax(a.z, b.z), cpm_fmax(a.w, b.w)}; return r; }

/* rgb2lab (Inviwo colorconversion.cl, un-vendored): sRGB (D65) -> XYZ -> CIE L*a*b*.  pow and cbrt come from
 * include/cpm_detmath.h (shared with the CUDA kernel), so this branch is compared bit for bit as well. */
static void rgb2lab(const float rgb[3], float lab[3]) {
    float lin[3];
    for (int k = 0; k < 3; ++k) {
        float c = rgb[k];
        lin[k] = c > 0.04045f ? cpm_powf((c + 0.055f) / 1.055f, 2.4f) : c / 12.92f;
    }
    float X = 0.4124564f * lin[0] + 0.3575761f * lin[1] + 0.1804375f * lin[2];
    float Y = 0.2126729f * lin[0] + 0.7151522f * lin[1] + 0.0721750f * lin[2];
    float Z = 0.0193339f * lin[0] + 0.1191920f * lin[1] + 0.9503041f * lin[2];
    float xyz[3] = {X / 0.95047f, Y / 1.0f, Z / 1.08883f};
    float f[3];
    for (int k = 0; k < 3; ++k) f[k] = xyz[k] > 0.008856f ? cpm_cbrtf(xyz[k]) : 7.787f * xyz[k] + 16.0f / 116.0f;
    lab[0] = 116.0f * f[1] - 16.0f;
    lab[1] = 500.0f * (f[0] - f[1]);
    lab[2] = 200.0f * (f[1] - f[2]);
}

/* :164-183 */
static float tfPointsImportance(v4 color, v4 nextColor, const float w[4], int incremental) {
    if (incremental) return nextColor.x + nextColor.y + nextColor.z + nextColor.w;
    float importance = 0.f;
    if (color.w > 0.f || nextColor.w > 0.f) {
        float a[3] = {color.x, color.y, color.z}, b[3] = {nextColor.x, nextColor.y, nextColor.z}, la[3], lb[3];
        rgb2lab(a, la);
        rgb2lab(b, lb);
        float dl[3] = {lb[0] - la[0], lb[1] - la[1], lb[2] - la[2]};
        /* length(): the same definition as everywhere else in the oracle (fma chain from x) */
        float lenN = sqrtf(fmaf(lb[2], lb[2], fmaf(lb[1], lb[1], lb[0] * lb[0])));
        float lenC = sqrtf(fmaf(la[2], la[2], fmaf(la[1], la[1], la[0] * la[0])));
        float lenD = sqrtf(fmaf(dl[2], dl[2], fmaf(dl[1], dl[1], dl[0] * dl[0])));
        float opacityDiff = nextColor.w - color.w;
        /* weights: colorWeight, colorDiffWeight, opacityDiffWeight, opacityWeight */
        importance = w[0] * fmaxf(lenN, lenC) + w[1] * lenD + w[2] * fabsf(opacityDiff) + w[3] * fmaxf(color.w, nextColor.w);
    }
    return importance;
}

/* :186-228 */
static float importanceForRangeTF(float lo, float hi, const float* positions, const v4* colors, int nPoints,
                                  const float w[4], int incremental) {
    int i = 0;
    while (i < nPoints - 1 && lo > positions[i + 1]) ++i;
    v4 color = v4_mix(colors[i], colors[i + 1], (lo - positions[i]) / (positions[i + 1] - positions[i]));
    v4 minColor = color, maxColor = color;
    if (hi <= positions[i + 1]) {
        v4 nextColor = v4_mix(colors[i], colors[i + 1], (hi - positions[i]) / (positions[i + 1] - positions[i]));
        minColor = v4_min(minColor, nextColor);
        maxColor = v4_max(maxColor, nextColor);
        return tfPointsImportance(minColor, maxColor, w, incremental);
    } else {
        v4 nextColor = colors[i + 1];
        minColor = v4_min(minColor, nextColor);
        maxColor = v4_max(maxColor, nextColor);
        ++i;
    }
    while (i < nPoints - 1 && hi > positions[i + 1]) {
        v4 nextColor = colors[i + 1];
        minColor = v4_min(minColor, nextColor);
        maxColor = v4_max(maxColor, nextColor);
        ++i;
    }
    if (i < nPoints - 1) {
        color = v4_mix(colors[i], colors[i + 1], (hi - positions[i]) / (positions[i + 1] - positions[i]));
        minColor = v4_min(minColor, color);
        maxColor = v4_max(maxColor, color);
    }
    return tfPointsImportance(minColor, maxColor, w, incremental);
}

/* classifyMinMaxUniformGrid3DImportanceKernel (:269-289) when prev == NULL, else
 * classifyTimeVaryingMinMaxUniformGrid3DImportanceKernel (:291-330). */
void orc_classify_importance(const uint16_t* minmax, const uint16_t* prev_minmax, const float* diff, int n,
                             const float* positions, const float* colors, int n_points, const float weights[4],
                             int incremental, float* out) {
    for (int i = 0; i < n; ++i) {
        uint16_t mn = minmax[2 * i], mx = minmax[2 * i + 1];
        if (prev_minmax) {
            if (prev_minmax[2 * i] < mn) mn = prev_minmax[2 * i];
            if (prev_minmax[2 * i + 1] > mx) mx = prev_minmax[2 * i + 1];
        }
        float lo = (1.f / 65535.f) * (float)mn, hi = (1.f / 65535.f) * (float)mx;
        float imp = importanceForRangeTF(lo, hi, positions, (const v4*)colors, n_points, weights, incremental);
        out[i] = prev_minmax ? diff[i] * imp : imp;
    }
}

/* ---- DDA traversal: ugc/cl/uniformgrid/uniformgrid.cl:38-69 and :147-197 (OPTIMIZE_STEP_FOR_SIMD) -- */
typedef struct { int x, y, z; } i3;

static float uniformGridImportance(v3 x1, v3 x2, v3 cellDim, const float* grid, const int dims[3]) {
    /* setupUniformGridTraversal */
    float mx[3] = {(float)(dims[0] - 1), (float)(dims[1] - 1), (float)(dims[2] - 1)};
    float a1[3] = {x1.x, x1.y, x1.z}, a2[3] = {x2.x, x2.y, x2.z}, cd[3] = {cellDim.x, cellDim.y, cellDim.z};
    float cellCoordf[3], dt[3], deltatx[3];
    int cellCoord[3], cellCoordEnd[3], di[3];
    for (int k = 0; k < 3; ++k) {
        cellCoordf[k] = cpm_clamp(floorf(a1[k] / cd[k]), 0.0f, mx[k]);
        cellCoord[k] = (int)cellCoordf[k];
        /* clamp(convert_int3(x2/cellDim), 0, maxCells-1): clamp in float first so that the
         * conversion is defined for any input (truncation toward zero, as convert_int) */
        cellCoordEnd[k] = (int)cpm_clamp(truncf(a2[k] / cd[k]), 0.0f, mx[k]);
        di[k] = (a1[k] < a2[k]) ? 1 : ((a1[k] > a2[k]) ? -1 : 0);
        float invAbsDir = 1.f / fabsf(a2[k] - a1[k]);
        float minx = cd[k] * cellCoordf[k];
        float maxx = minx + cd[k];
        dt[k] = ((a1[k] > a2[k]) ? (a1[k] - minx) : (maxx - a1[k])) * invAbsDir;
        deltatx[k] = cd[k] * invAbsDir;
    }
    int continueTraversal = 1;
    float importance = 0.f, dt1 = 0.f;
    while (continueTraversal) {
        float val = grid[cellCoord[0] + cellCoord[1] * dims[0] + cellCoord[2] * dims[0] * dims[1]];
        float dt0 = dt1;
        /* stepToNextCellNextHit */
        int ax = (dt[0] <= dt[1] && dt[0] <= dt[2]);
        int ay = (dt[0] > dt[1] && dt[1] <= dt[2]);
        if (ax) ay = 0;
        int az = !(ax || ay);
        int axis = ax ? 0 : (ay ? 1 : 2);
        (void)az;
        dt1 = dt[axis];
        if (cellCoord[axis] == cellCoordEnd[axis]) {
            continueTraversal = 0;
        } else {
            dt[axis] += deltatx[axis];
            cellCoord[axis] += di[axis];
        }
        importance += val * (cpm_fmin(1.f, dt1) - dt0);
    }
    v3 dvec = v3_sub(x2, x1);
    float len = sqrtf(fmaf(dvec.z, dvec.z, fmaf(dvec.y, dvec.y, dvec.x * dvec.x)));
    return importance * len;
}

static v3 transformPoint(const float m[16], v3 p) { /* column-major float16 */
    return v3_make(fmaf(m[8], p.z, fmaf(m[4], p.y, fmaf(m[0], p.x, m[12]))),
                   fmaf(m[9], p.z, fmaf(m[5], p.y, fmaf(m[1], p.x, m[13]))),
                   fmaf(m[10], p.z, fmaf(m[6], p.y, fmaf(m[2], p.x, m[14]))));
}

static uint32_t convert_uint_sat_rtp(float v) {
    if (!(v > 0.0f)) return 0u; /* negative, zero, NaN */
    float c = ceilf(v);
    if (c >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)c;
}

/* ppm/cl/photonrecomputationdetector.cl:92-157 (fix_exit = 0 reproduces :128 `exit = tEnd*direction`)
 * and :160-194 (equal importance). */
void orc_detect_invalid(const float* grid, const int grid_dims[3], const float cell_size[3], const float tex2idx[16],
                        const float* photons, int photon_offset, const float* light_samples, const float* isect,
                        int n_light_samples, int max_interactions, int total_photons, uint32_t* importances,
                        int equal_importance, int percentage, int iteration, int fix_exit) {
#pragma omp parallel for schedule(dynamic, 1024)
    for (int threadId = 0; threadId < n_light_samples; ++threadId) {
        float recomputationImportance = 0.f;
        if (equal_importance) {
            int photonId = photon_offset + threadId;
            int div = percentage > 0 && percentage <= 100 ? 100 / percentage : 1; /* guard the reference's /0 */
            if ((photonId + iteration) % div == 0) recomputationImportance = 1.f;
        } else {
            const float* ls = light_samples + 8 * (size_t)threadId;
            v3 origin = v3_make(ls[0], ls[1], ls[2]);
            v3 direction = decodeDirection(ls[6], ls[7]);
            float tStart = isect[2 * (size_t)threadId], tEnd = isect[2 * (size_t)threadId + 1];
            if (tStart < tEnd) {
                v3 entry = v3_ray(origin, tStart, direction);
                for (int interaction = 0; interaction < max_interactions; ++interaction) {
                    size_t photonId = (size_t)photon_offset + (size_t)interaction * total_photons + threadId;
                    const float* ph = photons + 8 * photonId;
                    v3 exit = v3_make(ph[0], ph[1], ph[2]);
                    if (ph[0] == FLT_MAX || ph[1] == FLT_MAX || ph[2] == FLT_MAX) {
                        if (interaction == 0) {
                            exit = fix_exit ? v3_ray(origin, tEnd, direction)
                                            : v3_make(tEnd * direction.x, tEnd * direction.y, tEnd * direction.z);
                        } else if (entry.x == FLT_MAX || entry.y == FLT_MAX || entry.z == FLT_MAX) {
                            break;
                        } else {
                            const float bmin[3] = {0.f, 0.f, 0.f}, bmax[3] = {1.f, 1.f, 1.f};
                            float t0 = 0.f, t1 = FLT_MAX;
                            v3 pd = decodeDirection(ph[6], ph[7]);
                            if (ph[3] != FLT_MAX && rayBoxIntersection(bmin, bmax, entry, pd, &t0, &t1)) {
                                if (fix_exit) {
                                    exit = v3_ray(entry, t1, pd); /* the evidently intended exit point */
                                } else {
                                    /* reference (:137): `exit += photonDirection*tEnd` with exit == photon.xyz ==
                                     * (FLT_MAX, FLT_MAX, FLT_MAX).  The sum stays FLT_MAX (or overflows), the index
                                     * coordinate x2 = textureToIndex * exit is +inf on every axis, the DDA then walks
                                     * along x with dt == 0 (every cell contributes val * 0) and multiplies the sum by
                                     * length(x2 - x1) == inf: the segment's importance is 0 * inf = NaN, the photon's
                                     * total is NaN, and convert_uint_sat_rtp(NaN) == 0 -- the photon is NOT flagged.
                                     * Checked against the reference's own kernel (oracle/_ref/libcl_ref.so). */
                                    recomputationImportance = NAN;
                                    break;
                                }
                            } else {
                                break;
                            }
                        }
                    }
                    v3 x1 = transformPoint(tex2idx, entry), x2 = transformPoint(tex2idx, exit);
                    x1 = v3_make(x1.x + 0.5f, x1.y + 0.5f, x1.z + 0.5f);
                    x2 = v3_make(x2.x + 0.5f, x2.y + 0.5f, x2.z + 0.5f);
                    recomputationImportance +=
                        uniformGridImportance(x1, x2, v3_make(cell_size[0], cell_size[1], cell_size[2]), grid, grid_dims);
                    entry = v3_make(ph[0], ph[1], ph[2]);
                }
            }
        }
        uint32_t v = convert_uint_sat_rtp(100.f * recomputationImportance);
        if (v > 2147483647u) v = 2147483647u; /* clamp(0u, 2147483647u, v) behaves as min */
        importances[photon_offset + threadId] -= v;
    }
}

/* ppm/cl/hashlightsample.cl:38-66 */
void orc_hash_light_samples(const float* light_samples, const float* isect, int n_light_source_samples,
                            const uint32_t* ids, int n_ids, const float cell_size[3], const int n_blocks[3],
                            uint32_t* which_bucket, int out_offset) {
    for (int g = 0; g < n_ids; ++g) {
        uint32_t id = ids[g];
        if (id < (uint32_t)out_offset || id >= (uint32_t)n_light_source_samples) continue;
        const float* ls = light_samples + 8 * (size_t)id;
        v3 origin = v3_make(ls[0], ls[1], ls[2]);
        v3 direction = decodeDirection(ls[6], ls[7]);
        float tStart = isect[2 * (size_t)id];
        v3 pos = v3_ray(origin, tStart, direction);
        /* convert_uint3: truncation; clamp in float so the conversion is defined */
        uint32_t hx = (uint32_t)cpm_clamp(truncf(pos.x * cell_size[0]), 0.f, 4294967040.f);
        uint32_t hy = (uint32_t)cpm_clamp(truncf(pos.y * cell_size[1]), 0.f, 4294967040.f);
        uint32_t hz = (uint32_t)cpm_clamp(truncf(pos.z * cell_size[2]), 0.f, 4294967040.f);
        which_bucket[out_offset + g] = hz * (uint32_t)n_blocks[0] * (uint32_t)n_blocks[1] + hy * (uint32_t)n_blocks[0] + hx;
    }
}

/* cell ranges over sorted keys (north-star subsystem 6; nothing in the reference):
 * start[c] = lower_bound(keys, c), end[c] = upper_bound(keys, c), defined for empty cells too. */
void orc_build_cell_ranges(const uint32_t* sorted_keys, size_t n, uint32_t n_cells, uint32_t* cell_start,
                           uint32_t* cell_end) {
    size_t i = 0;
    for (uint32_t c = 0; c < n_cells; ++c) {
        while (i < n && sorted_keys[i] < c) ++i;
        cell_start[c] = (uint32_t)i;
        size_t j = i;
        while (j < n && sorted_keys[j] == c) ++j;
        cell_end[c] = (uint32_t)j;
        i = j;
    }
}
