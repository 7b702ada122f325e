// geometry_ref_driver.cpp -- C entry points over the reference's own CPU light-plane fit
// (lcl/convexhull2d.cpp, lcl/orientedboundingbox2d.cpp, lcl/pointplaneprojection.cpp, compiled where they lie;
// GLM / Inviwo core stand-ins under ref_shim/host).  TEST INFRASTRUCTURE; built only into oracle/_ref/.
#include <modules/lightcl/convexhull2d.h>
#include <modules/lightcl/orientedboundingbox2d.h>
#include <modules/lightcl/pointplaneprojection.h>
#define REF_API extern "C" __attribute__((visibility("default")))
using namespace inviwo;

// hull_out holds up to 2 * n + 2 points; returns the hull size
REF_API int ref_convex_hull2d(const float* pts, int n, float* hull_out) {
    std::vector<vec2> p;
    for (int i = 0; i < n; ++i) p.emplace_back(pts[2 * i], pts[2 * i + 1]);
    auto h = geometry::convexHull2D(p);
    for (size_t i = 0; i < h.size(); ++i) { hull_out[2 * i] = h[i].x; hull_out[2 * i + 1] = h[i].y; }
    return (int)h.size();
}
// out = origin.xy, u.xy, v.xy
REF_API void ref_minimum_bounding_rectangle(const float* hull, int n, float out[6]) {
    std::vector<vec2> h;
    for (int i = 0; i < n; ++i) h.emplace_back(hull[2 * i], hull[2 * i + 1]);
    auto b = geometry::mimumBoundingRectangle(h);
    out[0] = b.origin.x; out[1] = b.origin.y; out[2] = b.u.x; out[3] = b.u.y; out[4] = b.v.x; out[5] = b.v.y;
}
REF_API void ref_project_points_on_plane(const float* pts, int n, const float P[3], const float N[3], const float u[3],
                                         const float v[3], float* out) {
    std::vector<vec3> p;
    for (int i = 0; i < n; ++i) p.emplace_back(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    std::vector<vec2> q;
    geometry::projectPointsOnPlane(p, Plane(vec3(P[0], P[1], P[2]), vec3(N[0], N[1], N[2])), vec3(u[0], u[1], u[2]),
                                   vec3(v[0], v[1], v[2]), q);
    for (int i = 0; i < n; ++i) { out[2 * i] = q[i].x; out[2 * i + 1] = q[i].y; }
}
// out = origin[3], u[3], v[3]
REF_API void ref_fit_plane_aligned_obb2d(const float* pts, int n, const float P[3], const float N[3], float out[9]) {
    std::vector<vec3> p;
    for (int i = 0; i < n; ++i) p.emplace_back(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    auto r = geometry::fitPlaneAlignedOrientedBoundingBox2D(p, Plane(vec3(P[0], P[1], P[2]), vec3(N[0], N[1], N[2])));
    vec3 o = std::get<0>(r), u = std::get<1>(r), v = std::get<2>(r);
    out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = u.x; out[4] = u.y; out[5] = u.z; out[6] = v.x; out[7] = v.y; out[8] = v.z;
}
