// glm.hpp -- stand-in for the part of GLM (g-truc/glm 0.9.7.x, vendored by Inviwo commit 989dc16e under ext/glm;
// NOT part of /root/reference) that the reference's lightcl geometry files use.  The definitions follow GLM's
// published ones operation for operation, because they fix the rounding of the light-plane fit:
//   dot(a, b)      = tmp = a * b; tmp.x + tmp.y (+ tmp.z)                      (detail/func_geometric.inl, compute_dot)
//   length(v)      = sqrt(dot(v, v))
//   inversesqrt(x) = 1 / sqrt(x)                                              (detail/func_exponential.inl)
//   normalize(v)   = v * inversesqrt(dot(v, v))
//   cross(x, y)    = (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)
// TEST INFRASTRUCTURE: lets oracle/Makefile compile lcl/convexhull2d.cpp, orientedboundingbox2d.cpp and
// pointplaneprojection.cpp where they lie.
#pragma once
#include <algorithm>
#include <cmath>
namespace glm {
struct vec2 {
    float x, y;
    vec2() : x(0.f), y(0.f) {}
    vec2(float a, float b) : x(a), y(b) {}
    explicit vec2(float a) : x(a), y(a) {}
};
struct vec3 {
    float x, y, z;
    vec3() : x(0.f), y(0.f), z(0.f) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit vec3(float a) : x(a), y(a), z(a) {}
};
struct vec4 {
    float x, y, z, w;
    vec4() : x(0.f), y(0.f), z(0.f), w(0.f) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    explicit vec4(float a) : x(a), y(a), z(a), w(a) {}
};
struct uvec3 {
    unsigned x, y, z;
    uvec3(unsigned a, unsigned b, unsigned c) : x(a), y(b), z(c) {}
};
struct bvec2 { bool x, y; };
inline vec2 operator+(vec2 a, vec2 b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(vec2 a, vec2 b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator*(vec2 a, vec2 b) { return vec2(a.x * b.x, a.y * b.y); }
inline vec2 operator*(vec2 a, float s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator*(float s, vec2 a) { return vec2(s * a.x, s * a.y); }
inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(vec3 a, vec3 b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, vec3 a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline float dot(vec2 a, vec2 b) { vec2 t(a * b); return t.x + t.y; }
inline float dot(vec3 a, vec3 b) { vec3 t(a * b); return t.x + t.y + t.z; }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline float length(vec2 v) { return std::sqrt(dot(v, v)); }
inline float length(vec3 v) { return std::sqrt(dot(v, v)); }
inline vec2 normalize(vec2 v) { return v * inversesqrt(dot(v, v)); }
inline vec3 normalize(vec3 v) { return v * inversesqrt(dot(v, v)); }
inline vec3 cross(vec3 x, vec3 y) { return vec3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
// component-wise std functions on vec2 (func_trigonometric.inl: functor1 over the scalar function)
inline vec2 cos(vec2 v) { return vec2(std::cos(v.x), std::cos(v.y)); }
inline vec2 sin(vec2 v) { return vec2(std::sin(v.x), std::sin(v.y)); }
inline float clamp(float x, float lo, float hi) { return std::min(std::max(x, lo), hi); }   // min(max(x, minVal), maxVal)
inline bvec2 isnan(vec2 v) { bvec2 r; r.x = std::isnan(v.x); r.y = std::isnan(v.y); return r; }
inline bool any(bvec2 b) { return b.x || b.y; }
}  // namespace glm
