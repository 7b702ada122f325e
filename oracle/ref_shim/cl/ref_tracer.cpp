// ref_tracer.cpp -- ppm/cl/photontracer.cl (+ transmittance.cl, photon.cl, rng/cl/random.cl, lcl lightsample.cl,
// isc light.cl) of the reference on the host.  Compiled once per kernel variant the reference builds
// (ppm/photontracercl.cpp:201-209): REF_ENTRY names the exported function, -DPHOTON_RECOMPUTATION /
// -DPROGRESSIVE_PHOTON_MAPPING / -DNO_SINGLE_SCATTERING select the variant.  TEST INFRASTRUCTURE.
#include "ref_common.h"
namespace {
#include "photontracer.cl"
}  // namespace

// the oracle's argument structs (cpm_oracle.h); returns nothing: collision tests are not counted by the reference
REF_API void REF_ENTRY(const orc_volume* vol, const float* tf, int tfw, const orc_trace_params* P, const float* lightSamples,
                       const float* isect, const uint32_t* recompute, int n_recompute, float* photons, uint32_t* rng) {
    clc_image volImg = ref_image3d(vol);
    clc_image2d tfImg = ref_image_tf(tf, tfw);
    VolumeParameters vp;
    memset(&vp, 0, sizeof(vp));
    vp.formatScaling = vol->scale;
    vp.formatOffset = vol->offset;
    BBox box;
    box.pMin = make_float3(P->aabb_min[0], P->aabb_min[1], P->aabb_min[2]);
    box.pMax = make_float3(P->aabb_max[0], P->aabb_max[1], P->aabb_max[2]);
    float4 material = make_float4(P->material[0], P->material[1], P->material[2], P->material[3]);
#ifdef PHOTON_RECOMPUTATION
    const int n = n_recompute;
#else
    const int n = P->n_light_samples;
#endif
    REF_FOR_EACH_WORK_ITEM(n, photonTracerKernel(
#ifdef PHOTON_RECOMPUTATION
        recompute, n_recompute,
#endif
        &volImg, &vp, &box, &tfImg, &tfImg /* the same layer is bound as scattering TF: ppm/photontracercl.cpp:150-151 */,
        material, rng, P->step_size, (float8*)photons, 1, P->photon_offset, 0, (const float8*)lightSamples,
        (const float2*)isect, P->n_light_samples, (uint)P->max_interactions, (ShadingType)P->phase_function, 0,
        P->total_photons));
}
