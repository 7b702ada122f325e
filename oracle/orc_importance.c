/* orc_importance.c -- oracle restatement of the view/light importance image and the importance-driven
 * 2-D sample generator.  TEST INFRASTRUCTURE (see cpm_oracle.h).
 *
 * orc_view_importance follows isc/cl/minmaxuniformgrid3dimportance.cl:336-378 (uniformGridImportanceKernel),
 * :86-133 (uniformGridImportance), :42-68 (stepToNextCell2) and ugc/cl/uniformgrid/uniformgrid.cl:38-69
 * (setupUniformGridTraversal).  orc_sample_importance2d has no counterpart in the reference (only the
 * SampleGenerator2DCL interface, lcl/samplegenerator2dcl.h:53-88): PARITY UNPINNED, own restatement. */
#include <stdlib.h>

#include "orc_common.h"

static v3 xf(const float m[16], v3 p) {
    return v3_make(fmaf(m[8], p.z, fmaf(m[4], p.y, fmaf(m[0], p.x, m[12]))),
                   fmaf(m[9], p.z, fmaf(m[5], p.y, fmaf(m[1], p.x, m[13]))),
                   fmaf(m[10], p.z, fmaf(m[6], p.y, fmaf(m[2], p.x, m[14]))));
}

void orc_view_importance(const uint16_t* minmax, const int dims[3], const float cellDim[3], const float tex2idx[16],
                         const float idx2tex[16], const float* entry, const float* exit, int width, int height,
                         float tf_min, float tf_max, float* out) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int id = 0; id < width * height; ++id) {
        v3 x1 = xf(tex2idx, v3_make(entry[4 * id], entry[4 * id + 1], entry[4 * id + 2]));
        v3 x2 = xf(tex2idx, v3_make(exit[4 * id], exit[4 * id + 1], exit[4 * id + 2]));
        x1 = v3_make(x1.x + 0.5f, x1.y + 0.5f, x1.z + 0.5f);
        x2 = v3_make(x2.x + 0.5f, x2.y + 0.5f, x2.z + 0.5f);
        if (x1.x == x2.x && x1.y == x2.y && x1.z == x2.z) {
            out[id] = 0.0f;
            continue;
        }
        float a1[3] = {x1.x, x1.y, x1.z}, a2[3] = {x2.x, x2.y, x2.z};
        float dt[3], deltatx[3];
        int cell[3], cellEnd[3], di[3];
        for (int k = 0; k < 3; ++k) {
            float mx = (float)(dims[k] - 1);
            float cf = cpm_clamp(floorf(a1[k] / cellDim[k]), 0.0f, mx);
            cell[k] = (int)cf;
            cellEnd[k] = (int)cpm_clamp(truncf(a2[k] / cellDim[k]), 0.0f, mx);
            di[k] = (a1[k] < a2[k]) ? 1 : ((a1[k] > a2[k]) ? -1 : 0);
            float invAbs = 1.0f / fabsf(a2[k] - a1[k]);
            float minx = cellDim[k] * cf;
            float maxx = minx + cellDim[k];
            dt[k] = ((a1[k] > a2[k]) ? (a1[k] - minx) : (maxx - a1[k])) * invAbs;
            deltatx[k] = cellDim[k] * invAbs;
        }
        v3 t1 = xf(idx2tex, x1), t2 = xf(idx2tex, x2);
        v3 dl = v3_sub(t2, t1);
        float len = sqrtf(fmaf(dl.z, dl.z, fmaf(dl.y, dl.y, dl.x * dl.x)));
        float importance = 0.0f, dt1 = 0.0f;
        int go = 1;
        while (go) {
            const uint16_t* mm = minmax + 2 * ((size_t)cell[0] + (size_t)cell[1] * dims[0] + (size_t)cell[2] * dims[0] * dims[1]);
            float lo = (1.0f / 65535.0f) * (float)mm[0], hi = (1.0f / 65535.0f) * (float)mm[1];
            float dt0 = dt1;
            int k;
            if (dt[0] <= dt[1] && dt[0] <= dt[2]) k = 0;
            else if (dt[1] <= dt[0] && dt[1] <= dt[2]) k = 1;
            else k = 2;
            dt1 = dt[k];
            if (cell[k] == cellEnd[k]) {
                go = 0;
            } else {
                dt[k] += deltatx[k];
                cell[k] += di[k];
            }
            if (!(hi < tf_min || lo > tf_max)) importance += cpm_fmin(1.0f, dt1) - dt0;
        }
        out[id] = importance * len;
    }
}

static int upper_cell(const float* c, int n, float t) {
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (c[mid] <= t) lo = mid; else hi = mid;
    }
    return lo;
}

void orc_sample_importance2d(const float* importance, int w, int h, float floor_value, const float* uniform_samples,
                             int n, float* out) {
    float* cdf = (float*)malloc(sizeof(float) * ((size_t)h * (w + 1) + h + 1));
    float* marg = cdf + (size_t)h * (w + 1);
    for (int y = 0; y < h; ++y) {
        float* row = cdf + (size_t)y * (w + 1);
        float acc = 0.0f;
        row[0] = 0.0f;
        for (int x = 0; x < w; ++x) {
            acc += cpm_fmax(importance[(size_t)y * w + x], 0.0f) + floor_value;
            row[x + 1] = acc;
        }
    }
    float acc = 0.0f;
    marg[0] = 0.0f;
    for (int y = 0; y < h; ++y) {
        acc += cdf[(size_t)y * (w + 1) + w];
        marg[y + 1] = acc;
    }
    const float total = marg[h], one_m = 0.99999994f;
    for (int i = 0; i < n; ++i) {
        const float* s = uniform_samples + 4 * (size_t)i;
        float tv = cpm_clamp(s[1], 0.0f, one_m) * total;
        int y = upper_cell(marg, h, tv);
        float rowsum = marg[y + 1] - marg[y];
        float dv = rowsum > 0.0f ? cpm_clamp((tv - marg[y]) / rowsum, 0.0f, one_m) : 0.5f;
        const float* row = cdf + (size_t)y * (w + 1);
        float tu = cpm_clamp(s[0], 0.0f, one_m) * row[w];
        int x = upper_cell(row, w, tu);
        float f = row[x + 1] - row[x];
        float du = f > 0.0f ? cpm_clamp((tu - row[x]) / f, 0.0f, one_m) : 0.5f;
        float pdf = total > 0.0f ? f * ((float)w * (float)h) / total : 1.0f;
        float* o = out + 4 * (size_t)i;
        o[0] = ((float)x + du) / (float)w;
        o[1] = ((float)y + dv) / (float)h;
        o[2] = s[2];
        o[3] = pdf * s[3];
    }
    free(cdf);
}
