// stand-in for Inviwo's umbrella header (un-vendored): the std headers and glm names the lightcl geometry files use
#pragma once
#ifndef _USE_MATH_DEFINES
#define _USE_MATH_DEFINES
#endif
#include <algorithm>
#include <string>
#include <cfloat>
#include <cmath>
#include <tuple>
#include <vector>
#include <glm/glm.hpp>
namespace inviwo {
using glm::vec2;
using glm::vec3;
using glm::vec4;
using glm::uvec3;
}  // namespace inviwo
