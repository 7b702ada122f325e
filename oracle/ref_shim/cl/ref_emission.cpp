// ref_emission.cpp -- isc/cl/uniformsamplegenerator2d.cl, lcl/cl/directionallightsampler.cl and
// lcl/cl/intersection/lightsamplemeshintersection.cl of the reference on the host.  TEST INFRASTRUCTURE.
#include "ref_common.h"
namespace {
#include "uniformsamplegenerator2d.cl"
#include "directionallightsampler.cl"
#include "intersection/lightsamplemeshintersection.cl"
}  // namespace

REF_API void ref_sample_uniform2d(float nx, float ny, int n, float* out) {
    REF_FOR_EACH_WORK_ITEM(n, uniformSampleGenerator2DKernel(make_float2(nx, ny), n, (float4*)out));
}
REF_API void ref_light_sample_directional(const float* samples, const float radiance[3], const float dir[3],
                                          const float origin[3], const float u[3], const float v[3], float area, int n,
                                          float* out) {
    REF_FOR_EACH_WORK_ITEM(n, directionalLightSamplerKernel((const float4*)samples, make_float4(radiance[0], radiance[1], radiance[2], 1.f),
                                                            make_float4(dir[0], dir[1], dir[2], 0.f),
                                                            make_float4(origin[0], origin[1], origin[2], 1.f),
                                                            make_float4(u[0], u[1], u[2], 0.f), make_float4(v[0], v[1], v[2], 0.f),
                                                            area, n, (float8*)out));
}
REF_API void ref_light_mesh_intersect(const float* vertices, const int* indices, int n_indices, const float* ls, int n,
                                      float* out) {
    REF_FOR_EACH_WORK_ITEM(n, lightSampleMeshIntersectionKernel(vertices, indices, n_indices, (const float8*)ls, n, (float2*)out));
}
