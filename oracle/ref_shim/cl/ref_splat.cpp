// ref_splat.cpp -- ppm/cl/photonstolightvolume.cl + densityestimationkernel.cl of the reference on the host.
// Compiled twice: -DVOLUME_OUTPUT_SINGLE_CHANNEL=1 (REF_ENTRY=ref_splat_1) and without (ref_splat_4).  The adds are
// the reference's own float compare-and-swap loop, executed in work-item order.  TEST INFRASTRUCTURE.
#include "ref_common.h"
namespace {
#include "photonstolightvolume.cl"
}  // namespace

REF_API void REF_ENTRY(float* volume_out, const float tex2idx[16], const float idx2tex[16], const int out_dims[3],
                       const float* photons, const uint32_t* indices, int n, int photons_per_interaction,
                       int n_interactions, float radius, float relative_irradiance_scale, float multiplier) {
    VolumeParameters vp;
    memset(&vp, 0, sizeof(vp));
    vp.textureToIndex = ref_mat(tex2idx);
    vp.indexToTexture = ref_mat(idx2tex);
    int4 od = make_int4(out_dims[0], out_dims[1], out_dims[2], 0);
#ifdef VOLUME_OUTPUT_SINGLE_CHANNEL
    float* out = volume_out;
#else
    float4* out = (float4*)volume_out;
#endif
    if (!indices) {
        REF_FOR_EACH_WORK_ITEM(n, splatPhotonsToLightVolumeKernel(nullptr, &vp, out, &vp, od, (float8*)photons, n, radius,
                                                                  relative_irradiance_scale));
    } else {
        REF_FOR_EACH_WORK_ITEM(n, splatSelectedPhotonsToLightVolumeKernel(out, &vp, od, (float8*)photons, (int*)indices, n, radius,
                                                                          relative_irradiance_scale, multiplier,
                                                                          photons_per_interaction, n_interactions));
    }
}
