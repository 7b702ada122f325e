// ref_detector.cpp -- ppm/cl/photonrecomputationdetector.cl + ugc/cl/uniformgrid/uniformgrid.cl (the DDA) of the
// reference on the host.  TEST INFRASTRUCTURE.
#include "ref_common.h"
namespace {
#include "photonrecomputationdetector.cl"
}  // namespace

REF_API void ref_detect_invalid(const float* grid, const int grid_dims[3], const float cell_size[3], const float tex2idx[16],
                                const float* photons, int photon_offset, const float* light_samples, const float* isect,
                                int n_light_samples, int max_interactions, int total_photons, uint32_t* importances,
                                int equal_importance, int percentage, int iteration) {
    int4 gd = make_int4(grid_dims[0], grid_dims[1], grid_dims[2], 0);
    float3 cs = make_float3(cell_size[0], cell_size[1], cell_size[2]);
    float16 t2i = ref_mat(tex2idx), i2t = ref_mat(tex2idx) /* indexToTextureMat is not used by the kernel */;
    if (equal_importance) {
        REF_FOR_EACH_WORK_ITEM(n_light_samples, photonRecomputationDetectorEqualImportanceKernel(
            grid, gd, cs, t2i, i2t, (float8*)photons, photon_offset, (const float8*)light_samples, (const float2*)isect,
            n_light_samples, (uint)max_interactions, total_photons, importances, percentage, iteration));
    } else {
        REF_FOR_EACH_WORK_ITEM(n_light_samples, photonRecomputationDetectorKernel(
            grid, gd, cs, t2i, i2t, (float8*)photons, photon_offset, (const float8*)light_samples, (const float2*)isect,
            n_light_samples, (uint)max_interactions, total_photons, importances));
    }
}
// the DDA alone: importance of the segment x1 -> x2 (index coordinates + 0.5, as the kernel passes them)
REF_API float ref_uniform_grid_importance(const float x1[3], const float x2[3], const float cell_size[3], const float* grid,
                                          const int grid_dims[3]) {
    float tHit;
    return uniformGridImportance(make_float3(x1[0], x1[1], x1[2]), make_float3(x2[0], x2[1], x2[2]),
                                 make_float3(cell_size[0], cell_size[1], cell_size[2]), grid,
                                 make_int4(grid_dims[0], grid_dims[1], grid_dims[2], 0), &tHit);
}
