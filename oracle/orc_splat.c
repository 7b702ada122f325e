/* orc_splat.c -- oracle restatement of photon density estimation by splatting.
 * TEST INFRASTRUCTURE (see cpm_oracle.h).
 *
 * ppm/cl/photonstolightvolume.cl:31-79 (splatPhoton), :139-166, :168-202 and
 * ppm/cl/densityestimationkernel.cl:56-60.  The reference adds with a float CAS loop in
 * arbitrary order; the oracle accumulates in DOUBLE so that it is the order-free value the
 * float sums scatter around. */
#include "orc_common.h"

static v3 xform(const float m[16], v3 p) {
    return v3_make(fmaf(m[8], p.z, fmaf(m[4], p.y, fmaf(m[0], p.x, m[12]))),
                   fmaf(m[9], p.z, fmaf(m[5], p.y, fmaf(m[1], p.x, m[13]))),
                   fmaf(m[10], p.z, fmaf(m[6], p.y, fmaf(m[2], p.x, m[14]))));
}

static void splatPhoton(double* vol, int channels, const float tex2idx[16], const float idx2tex[16],
                        const int outDim[3], const float ph[8], float pr, float pg, float pb, float radius) {
    if (ph[0] == FLT_MAX || ph[1] == FLT_MAX || ph[2] == FLT_MAX) return;
    v3 p = v3_make(ph[0], ph[1], ph[2]);
    v3 lo = xform(tex2idx, v3_make(p.x - radius, p.y - radius, p.z - radius));
    v3 hi = xform(tex2idx, v3_make(p.x + radius, p.y + radius, p.z + radius));
    float l[3] = {lo.x, lo.y, lo.z}, h[3] = {hi.x + 1.f, hi.y + 1.f, hi.z + 1.f};
    int s[3], e[3];
    for (int k = 0; k < 3; ++k) {
        /* max(0, convert_int3(..)), min(convert_int3(.. + 1), outDim): truncation; clamped in float first */
        s[k] = (int)cpm_clamp(truncf(l[k]), 0.f, 2147483520.f);
        e[k] = (int)cpm_clamp(truncf(h[k]), -2147483520.f, (float)outDim[k]);
    }
    for (int z = s[2]; z < e[2]; ++z)
        for (int y = s[1]; y < e[1]; ++y)
            for (int x = s[0]; x < e[0]; ++x) {
                size_t voxelIndex = (size_t)x + (size_t)y * outDim[0] + (size_t)z * outDim[0] * outDim[1];
                v3 c = xform(idx2tex, v3_make((float)x, (float)y, (float)z));
                v3 d = v3_sub(c, p);
                float dist = sqrtf(fmaf(d.z, d.z, fmaf(d.y, d.y, d.x * d.x)));
                float xk = dist / radius;
                float weight = xk <= 1.f ? 0.75f * (1.f - xk * xk) : 0.f; /* Epanechnikov */
                float fr = pr * weight, fg = pg * weight, fb = pb * weight;
                /* photons are splatted by several threads: atomic double adds (exact enough that the
                 * order does not show at the tolerances the tests use) */
                if (channels == 1) {
                    if (fr != 0.f) {
#pragma omp atomic
                        vol[voxelIndex] += fr;
                    }
                } else {
                    if (fr != 0.f) {
#pragma omp atomic
                        vol[voxelIndex * 4] += fr;
                    }
                    if (fg != 0.f) {
#pragma omp atomic
                        vol[voxelIndex * 4 + 1] += fg;
                    }
                    if (fb != 0.f) {
#pragma omp atomic
                        vol[voxelIndex * 4 + 2] += fb;
                    }
                }
            }
}

/* indices == NULL: splatPhotonsToLightVolumeKernel over photon ids [0, n) (the host passes
 * n = getNumberOfPhotons(), so only interaction 0 is covered -- reproduced by the caller);
 * otherwise splatSelectedPhotonsToLightVolumeKernel: n indices x n_interactions, x multiplier. */
void orc_splat(double* vol, int channels, const float tex2idx[16], const float idx2tex[16], const int outDim[3],
               const float* photons, const uint32_t* indices, int n, int photons_per_interaction, int n_interactions,
               float radius, float relative_irradiance_scale, float multiplier) {
    const float phase = CPM_INV_4PI_F; /* isotropicPhaseFunction() */
#pragma omp parallel for schedule(dynamic, 1024)
    for (int g = 0; g < n; ++g) {
        if (!indices) {
            const float* ph = photons + 8 * (size_t)g;
            float s = phase * relative_irradiance_scale;
            splatPhoton(vol, channels, tex2idx, idx2tex, outDim, ph, ph[3] * s, ph[4] * s, ph[5] * s, radius);
        } else {
            uint32_t photonId = indices[g];
            for (int interaction = 0; interaction < n_interactions; ++interaction) {
                const float* ph = photons + 8 * ((size_t)interaction * photons_per_interaction + photonId);
                float s = phase * relative_irradiance_scale;
                float pr = ph[3] * s * multiplier, pg = ph[4] * s * multiplier, pb = ph[5] * s * multiplier;
                splatPhoton(vol, channels, tex2idx, idx2tex, outDim, ph, pr, pg, pb, radius);
            }
        }
    }
}
