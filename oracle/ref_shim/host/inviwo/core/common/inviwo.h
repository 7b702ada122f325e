// stand-in for Inviwo's umbrella header (un-vendored): the std headers and glm names the lightcl geometry files use
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <tuple>
#include <vector>
#include <glm/glm.hpp>
namespace inviwo {
using glm::vec2;
using glm::vec3;
}  // namespace inviwo
