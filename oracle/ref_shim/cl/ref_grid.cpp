// ref_grid.cpp -- ugc/cl/uniformgrid/volumeminmax.cl and isc/cl/minmaxuniformgrid3dimportance.cl of the reference on
// the host.  Compiled twice: with -DINCREMENTAL_TF_IMPORTANCE (the static kernel's build, isc/minmaxuniformgrid3dimportancecl
// .cpp:99-101; REF_SUFFIX=_incremental) and without (the time-varying kernel's, REF_SUFFIX=_lab).  TEST INFRASTRUCTURE.
#include "ref_common.h"
namespace {
#include "uniformgrid/volumeminmax.cl"
#include "minmaxuniformgrid3dimportance.cl"
}  // namespace
#define REF_CAT2(a, b) a##b
#define REF_CAT(a, b) REF_CAT2(a, b)

REF_API void REF_CAT(ref_volume_minmax, REF_SUFFIX)(const orc_volume* vol, int region, uint16_t* out) {
    clc_image im = ref_image3d(vol);
    VolumeParameters vp;
    memset(&vp, 0, sizeof(vp));
    vp.formatScaling = vol->scale;
    vp.formatOffset = vol->offset;
    int ox = (vol->dims[0] + region - 1) / region, oy = (vol->dims[1] + region - 1) / region, oz = (vol->dims[2] + region - 1) / region;
    int4 outDim = make_int4(ox, oy, oz, 0), reg = make_int4(region, region, region, 0);
    clc::wi().gsize[0] = ox; clc::wi().gsize[1] = oy; clc::wi().gsize[2] = oz;
    for (int z = 0; z < oz; ++z)
        for (int y = 0; y < oy; ++y)
            for (int x = 0; x < ox; ++x) {
                clc::wi().gid[0] = x; clc::wi().gid[1] = y; clc::wi().gid[2] = z;
                volumeMinMaxKernel(&im, &vp, (ushort2*)out, outDim, reg);
            }
}
// prev_minmax == NULL: classifyMinMaxUniformGrid3DImportanceKernel, else the time-varying kernel
REF_API void REF_CAT(ref_classify_importance, REF_SUFFIX)(const uint16_t* minmax, const uint16_t* prev_minmax, const float* diff,
                                                         int n, const float* positions, const float* colors, int n_points,
                                                         const float w[4], float* out) {
    if (!prev_minmax) {
        REF_FOR_EACH_WORK_ITEM(n, classifyMinMaxUniformGrid3DImportanceKernel((const ushort2*)minmax, n, positions, (const float4*)colors,
                                                                              n_points, w[0], w[1], w[2], w[3], out));
    } else {
        REF_FOR_EACH_WORK_ITEM(n, classifyTimeVaryingMinMaxUniformGrid3DImportanceKernel(
            (const ushort2*)minmax, (const ushort2*)prev_minmax, diff, n, positions, (const float4*)colors, n_points, w[0], w[1],
            w[2], w[3], out));
    }
}
