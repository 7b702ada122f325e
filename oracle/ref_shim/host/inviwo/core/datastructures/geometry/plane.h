// stand-in for inviwo/core/datastructures/geometry/plane.h (Inviwo core at commit 989dc16e, un-vendored):
//   distance(p) = dot(p - point, normal);  projectPoint(p) = p - distance(p) * normal
#pragma once
#include <inviwo/core/common/inviwo.h>
namespace inviwo {
class Plane {
public:
    Plane(vec3 point, vec3 normal) : point_(point), normal_(normal) {}
    const vec3& getPoint() const { return point_; }
    const vec3& getNormal() const { return normal_; }
    float distance(const vec3& p) const { return glm::dot(p - point_, normal_); }
    vec3 projectPoint(const vec3& p) const { return p - distance(p) * normal_; }
private:
    vec3 point_, normal_;
};
}  // namespace inviwo
