// clc.h -- the part of OpenCL C 1.2 that the reference's kernels use, as C++14, so that the kernel files of
// /root/reference compile for the host WHERE THEY LIE (oracle/Makefile `refcl`).  TEST INFRASTRUCTURE: only
// oracle/_ref/libcl_ref.so is built from it; nothing of the product includes it.
//
// What is here and what is not:
//  * vector types float2/3/4/8/16, int2/3/4, uint2/3/4, ushort2 with the component names, swizzles, operators,
//    relational results (-1 / 0 per component), select / any, conversions and the geometric / common built-ins the
//    kernels call -- each defined as OpenCL 1.2 section 6.12 states it (citations at the definitions);
//  * work-item functions over a thread-local id (the drivers run one work-item at a time);
//  * address-space and access qualifiers as empty macros, __constant as const;
//  * native_log / native_exp / native_sin / native_cos / acos / atan2 are implementation-defined in OpenCL: they are
//    mapped to include/cpm_detmath.h, the definition the oracle and the CUDA kernels share;
//  * the ONE textual rewrite the Makefile applies to a kernel file is "(typeN)(" -> "make_typeN(" (C++ has no
//    vector-literal syntax) plus, in uniformgrid.cl, the two vector-condition ternaries (C++ cannot overload ?:).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

#include "cpm_detmath.h"

typedef uint8_t uchar;
typedef uint16_t ushort;
typedef uint32_t uint;
typedef uint64_t ulong;

#define __kernel
#define __global
#define __local
#define __private
#define __constant const
#define __read_only
#define __write_only
#define read_only
#define write_only
#define global
#define restrict __restrict
#define CLC_INLINE inline
#ifndef M_PI_F
#define M_PI_F 3.14159274101257324219f
#endif

// ---- work-item functions (section 6.12.1) ----------------------------------------------------------------------
namespace clc {
struct WorkItem { size_t gid[3], gsize[3]; };
inline WorkItem& wi() { static thread_local WorkItem w = {{0, 0, 0}, {1, 1, 1}}; return w; }
}  // namespace clc
inline size_t get_global_id(uint d) { return clc::wi().gid[d]; }
inline size_t get_global_size(uint d) { return clc::wi().gsize[d]; }

// ---- swizzle proxies ----------------------------------------------------------------------------------------------
// `v.xyz`, `v.s345` ... are members of a union with the component array; they convert to / assign from the vector type
// they name.  N = number of stored components of the parent.
template <class T, class V, int N, int A, int B>
struct Swz2 {
    T d[N];
    operator V() const { return V(d[A], d[B]); }
    Swz2& operator=(const V& v) { d[A] = v.s[0]; d[B] = v.s[1]; return *this; }
    Swz2& operator+=(const V& v) { return *this = V(*this) + v; }
    Swz2& operator-=(const V& v) { return *this = V(*this) - v; }
    Swz2& operator*=(const V& v) { return *this = V(*this) * v; }
    Swz2& operator*=(T s) { return *this = V(*this) * s; }
    Swz2& operator/=(T s) { return *this = V(*this) / s; }
};
template <class T, class V, int N, int A, int B, int C>
struct Swz3 {
    T d[N];
    operator V() const { return V(d[A], d[B], d[C]); }
    Swz3& operator=(const V& v) { d[A] = v.s[0]; d[B] = v.s[1]; d[C] = v.s[2]; return *this; }
    Swz3& operator+=(const V& v) { return *this = V(*this) + v; }
    Swz3& operator-=(const V& v) { return *this = V(*this) - v; }
    Swz3& operator*=(const V& v) { return *this = V(*this) * v; }
    Swz3& operator*=(T s) { return *this = V(*this) * s; }
    Swz3& operator/=(T s) { return *this = V(*this) / s; }
};

// ---- vector types -----------------------------------------------------------------------------------------------------
#define CLC_VEC2(NAME, T)                                                       \
    struct NAME {                                                               \
        union {                                                                 \
            T s[2];                                                             \
            struct { T x, y; };                                                 \
            struct { T s0, s1; };                                               \
            Swz2<T, NAME, 2, 0, 1> xy;                                          \
        };                                                                      \
        NAME() : s{0, 0} {}                                                     \
        NAME(T a) : s{a, a} {}                                                  \
        NAME(T a, T b) : s{a, b} {}                                             \
    };
#define CLC_VEC3(NAME, T, V2)                                                   \
    struct NAME {                                                               \
        union {                                                                 \
            T s[4];                                                             \
            struct { T x, y, z; };                                              \
            struct { T s0, s1, s2; };                                           \
            Swz2<T, V2, 4, 0, 1> xy;                                            \
            Swz3<T, NAME, 4, 0, 1, 2> xyz;                                      \
        };                                                                      \
        NAME() : s{0, 0, 0, 0} {}                                               \
        NAME(T a) : s{a, a, a, 0} {}                                            \
        NAME(T a, T b, T c) : s{a, b, c, 0} {}                                  \
    };
#define CLC_VEC4(NAME, T, V2, V3)                                               \
    struct NAME {                                                               \
        union {                                                                 \
            T s[4];                                                             \
            struct { T x, y, z, w; };                                           \
            struct { T s0, s1, s2, s3; };                                       \
            Swz2<T, V2, 4, 0, 1> xy;                                            \
            Swz3<T, V3, 4, 0, 1, 2> xyz;                                        \
        };                                                                      \
        NAME() : s{0, 0, 0, 0} {}                                               \
        NAME(T a) : s{a, a, a, a} {}                                            \
        NAME(T a, T b, T c, T d) : s{a, b, c, d} {}                             \
    };

CLC_VEC2(float2, float)
CLC_VEC2(int2, int)
CLC_VEC2(uint2, uint)
CLC_VEC2(ushort2, ushort)
CLC_VEC3(float3, float, float2)
CLC_VEC3(int3, int, int2)
CLC_VEC3(uint3, uint, uint2)
CLC_VEC4(float4, float, float2, float3)
CLC_VEC4(int4, int, int2, int3)
CLC_VEC4(uint4, uint, uint2, uint3)

struct float8 {
    union {
        float s[8];
        struct { float x, y, z, w; };
        struct { float s0, s1, s2, s3, s4, s5, s6, s7; };
        Swz3<float, float3, 8, 0, 1, 2> xyz;
        Swz3<float, float3, 8, 0, 1, 2> s012;
        Swz3<float, float3, 8, 3, 4, 5> s345;
        Swz2<float, float2, 8, 6, 7> s67;
    };
    float8() : s{0, 0, 0, 0, 0, 0, 0, 0} {}
    float8(float a, float b, float c, float d, float e, float f, float g, float h) : s{a, b, c, d, e, f, g, h} {}
};
struct float16 {
    float s[16];   // column-major 4 x 4 when it holds a matrix (glm layout, as the host uploads it)
};

// vector literals: "(typeN)(...)" is rewritten to "make_typeN(...)" by the Makefile (section 6.1.6: a single scalar
// is replicated, otherwise the operands are concatenated).  make_typeN is a MACRO over a braced initialiser so that the
// operands are evaluated LEFT TO RIGHT (C++ [dcl.init.list]/4) -- `(float2)(random_01(s), random_01(s))`
// (ppm/cl/photontracer.cl:53) must draw its first number into .x, as OpenCL compilers (clang) do; a plain function
// call would leave the order to the host compiler (gcc evaluates arguments right to left).
#define CLC_MK2(NAME, T)                          \
    struct mk_##NAME {                            \
        NAME v;                                   \
        mk_##NAME(T a) : v(a) {}                  \
        mk_##NAME(T a, T b) : v(a, b) {}          \
        mk_##NAME(const NAME& a) : v(a) {}        \
    };
CLC_MK2(float2, float)
CLC_MK2(int2, int)
CLC_MK2(uint2, uint)
CLC_MK2(ushort2, ushort)
#define CLC_MK3(NAME, T, V2)                                          \
    struct mk_##NAME {                                                \
        NAME v;                                                       \
        mk_##NAME(T a) : v(a) {}                                      \
        mk_##NAME(T a, T b, T c) : v(a, b, c) {}                      \
        mk_##NAME(const V2& a, T c) : v(a.s[0], a.s[1], c) {}         \
        mk_##NAME(const NAME& a) : v(a) {}                            \
    };
CLC_MK3(float3, float, float2)
CLC_MK3(int3, int, int2)
CLC_MK3(uint3, uint, uint2)
#define CLC_MK4(NAME, T, V2, V3)                                                  \
    struct mk_##NAME {                                                            \
        NAME v;                                                                   \
        mk_##NAME(T a) : v(a) {}                                                  \
        mk_##NAME(T a, T b, T c, T d) : v(a, b, c, d) {}                          \
        mk_##NAME(const V3& a, T d) : v(a.s[0], a.s[1], a.s[2], d) {}             \
        mk_##NAME(const V2& a, T c, T d) : v(a.s[0], a.s[1], c, d) {}             \
        mk_##NAME(const NAME& a) : v(a) {}                                        \
    };
CLC_MK4(float4, float, float2, float3)
CLC_MK4(int4, int, int2, int3)
CLC_MK4(uint4, uint, uint2, uint3)
struct mk_float8 {
    float8 v;
    mk_float8(float a, float b, float c, float d, float e, float f, float g, float h) : v(a, b, c, d, e, f, g, h) {}
    mk_float8(const float3& a, const float3& b, const float2& c) : v(a.s[0], a.s[1], a.s[2], b.s[0], b.s[1], b.s[2], c.s[0], c.s[1]) {}
};
#define make_float2(...) (mk_float2{__VA_ARGS__}.v)
#define make_int2(...) (mk_int2{__VA_ARGS__}.v)
#define make_uint2(...) (mk_uint2{__VA_ARGS__}.v)
#define make_ushort2(...) (mk_ushort2{__VA_ARGS__}.v)
#define make_float3(...) (mk_float3{__VA_ARGS__}.v)
#define make_int3(...) (mk_int3{__VA_ARGS__}.v)
#define make_uint3(...) (mk_uint3{__VA_ARGS__}.v)
#define make_float4(...) (mk_float4{__VA_ARGS__}.v)
#define make_int4(...) (mk_int4{__VA_ARGS__}.v)
#define make_uint4(...) (mk_uint4{__VA_ARGS__}.v)
#define make_float8(...) (mk_float8{__VA_ARGS__}.v)

// ---- operators (section 6.3): component-wise; relational operators give -1 (true) / 0 per component ---------------
#define CLC_ARITH(V, T, N)                                                                                         \
    inline V operator+(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] + b.s[i]; return r; } \
    inline V operator-(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] - b.s[i]; return r; } \
    inline V operator*(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] * b.s[i]; return r; } \
    inline V operator/(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] / b.s[i]; return r; } \
    inline V operator+(const V& a, T b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] + b; return r; }         \
    inline V operator-(const V& a, T b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] - b; return r; }         \
    inline V operator/(const V& a, T b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] / b; return r; }         \
    inline V operator+(T a, const V& b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a + b.s[i]; return r; }         \
    inline V operator-(T a, const V& b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a - b.s[i]; return r; }         \
    inline V operator/(T a, const V& b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a / b.s[i]; return r; }         \
    inline V operator-(const V& a) { V r; for (int i = 0; i < N; ++i) r.s[i] = -a.s[i]; return r; }                 \
    inline V& operator+=(V& a, const V& b) { return a = a + b; }                                                    \
    inline V& operator-=(V& a, const V& b) { return a = a - b; }                                                    \
    inline V& operator*=(V& a, const V& b) { return a = a * b; }                                                    \
    inline V& operator/=(V& a, const V& b) { return a = a / b; }                                                    \
    inline V& operator+=(V& a, T b) { return a = a + b; }                                                           \
    inline V& operator-=(V& a, T b) { return a = a - b; }                                                           \
    inline V& operator*=(V& a, T b) { return a = a * b; }                                                           \
    inline V& operator/=(V& a, T b) { return a = a / b; }
#define CLC_SCALE(V, T, N)                                                                                      \
    inline V operator*(const V& a, T b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] * b; return r; }      \
    inline V operator*(T a, const V& b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a * b.s[i]; return r; }
#define CLC_REL(V, IV, N)                                                                                              \
    inline IV operator<(const V& a, const V& b) { IV r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] < b.s[i] ? -1 : 0; return r; }   \
    inline IV operator>(const V& a, const V& b) { IV r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] > b.s[i] ? -1 : 0; return r; }   \
    inline IV operator<=(const V& a, const V& b) { IV r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] <= b.s[i] ? -1 : 0; return r; } \
    inline IV operator>=(const V& a, const V& b) { IV r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] >= b.s[i] ? -1 : 0; return r; } \
    inline IV operator==(const V& a, const V& b) { IV r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] == b.s[i] ? -1 : 0; return r; } \
    inline IV operator!=(const V& a, const V& b) { IV r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] != b.s[i] ? -1 : 0; return r; }
CLC_ARITH(float2, float, 2)
CLC_ARITH(float3, float, 3)
CLC_ARITH(float4, float, 4)
CLC_SCALE(float2, float, 2)
CLC_SCALE(float4, float, 4)
CLC_SCALE(int2, int, 2)
CLC_SCALE(int3, int, 3)
CLC_SCALE(int4, int, 4)
CLC_SCALE(uint3, uint, 3)
#ifndef CLC_CONTRACT
CLC_SCALE(float3, float, 3)
#else
// FP_CONTRACT ON (the OpenCL C default, section 6.12.2... #pragma OPENCL FP_CONTRACT): a product that directly feeds a
// sum may be evaluated as one fused multiply-add.  Scalar expressions are contracted by the host compiler
// (-ffp-contract=fast); for `vector + scalar * vector` the fusion is spelled out here, because the compiler does not see
// through the vector classes reliably: scalar * float3 yields a lazy product that + / - / += consume with fmaf.
struct Prod3 {
    float3 v;
    float k;
    operator float3() const { return float3(v.s[0] * k, v.s[1] * k, v.s[2] * k); }
};
inline Prod3 operator*(float k, const float3& v) { Prod3 p; p.v = v; p.k = k; return p; }
inline Prod3 operator*(const float3& v, float k) { Prod3 p; p.v = v; p.k = k; return p; }
inline float3 operator*(const Prod3& p, float k) { return float3(p) * float3(k); }
inline float3 operator/(const Prod3& p, float k) { return float3(p) / k; }
inline float3 operator-(const Prod3& p) { return -float3(p); }
inline float3 operator+(const float3& a, const Prod3& p) { return float3(fmaf(p.k, p.v.s[0], a.s[0]), fmaf(p.k, p.v.s[1], a.s[1]), fmaf(p.k, p.v.s[2], a.s[2])); }
inline float3 operator+(const Prod3& p, const float3& a) { return a + p; }
inline float3 operator-(const float3& a, const Prod3& p) { return float3(fmaf(-p.k, p.v.s[0], a.s[0]), fmaf(-p.k, p.v.s[1], a.s[1]), fmaf(-p.k, p.v.s[2], a.s[2])); }
inline float3& operator+=(float3& a, const Prod3& p) { return a = a + p; }
inline float3& operator-=(float3& a, const Prod3& p) { return a = a - p; }
#endif
CLC_ARITH(int2, int, 2)
CLC_ARITH(int3, int, 3)
CLC_ARITH(int4, int, 4)
CLC_ARITH(uint3, uint, 3)
CLC_REL(float2, int2, 2)
CLC_REL(float3, int3, 3)
CLC_REL(float4, int4, 4)
CLC_REL(int2, int2, 2)
CLC_REL(int3, int3, 3)
CLC_REL(int4, int4, 4)
// logical operators on vectors (section 6.3.g): per component, -1 / 0
inline int3 operator&&(const int3& a, const int3& b) { int3 r; for (int i = 0; i < 3; ++i) r.s[i] = (a.s[i] && b.s[i]) ? -1 : 0; return r; }
inline int3 operator||(const int3& a, const int3& b) { int3 r; for (int i = 0; i < 3; ++i) r.s[i] = (a.s[i] || b.s[i]) ? -1 : 0; return r; }

// ---- relational built-ins (section 6.12.6) -----------------------------------------------------------------------
// any(x): 1 if the most significant bit of any component is set.  NOTE the scalar case: a scalar comparison yields
// 1, whose MSB is clear -- so `any(a >= b)` on scalars is always 0 (photonstolightvolume.cl:152 relies on nothing else).
inline int any(int x) { return x < 0; }
inline int any(const int2& v) { return (v.s[0] | v.s[1]) < 0; }
inline int any(const int3& v) { return (v.s[0] | v.s[1] | v.s[2]) < 0; }
inline int any(const int4& v) { return (v.s[0] | v.s[1] | v.s[2] | v.s[3]) < 0; }
// select(a, b, c): vector c -> component MSB set ? b : a
inline float3 select(const float3& a, const float3& b, const int3& c) { float3 r; for (int i = 0; i < 3; ++i) r.s[i] = c.s[i] < 0 ? b.s[i] : a.s[i]; return r; }
inline int3 select(const int3& a, const int3& b, const int3& c) { int3 r; for (int i = 0; i < 3; ++i) r.s[i] = c.s[i] < 0 ? b.s[i] : a.s[i]; return r; }
// `c ? a : b` with a vector condition (section 6.3.i) = select(b, a, c); used by the Makefile's rewrite of the two
// vector ternaries in uniformgrid.cl
inline float3 vternary(const int3& c, const float3& a, const float3& b) { return select(b, a, c); }
inline int3 vternary(const int3& c, const int3& a, const int3& b) { return select(b, a, c); }

// ---- conversions (section 6.2.3) ------------------------------------------------------------------------------------
// float -> int: round toward zero.  Out-of-range / NaN inputs are undefined in OpenCL; the saturating behaviour of
// NVIDIA's cvt.rzi (NaN -> 0) is used, which is also what a clamp in float before the conversion gives (the oracle).
inline int clc_f2i(float v) {
    if (!(v == v)) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return (int)0x80000000;
    return (int)v;
}
inline uint clc_f2u(float v) {
    if (!(v > 0.0f)) return 0u;
    if (v >= 4294967296.0f) return 0xffffffffu;
    return (uint)v;
}
inline float convert_float(int v) { return (float)v; }
inline float convert_float(uint v) { return (float)v; }
inline float convert_float(float v) { return v; }
inline float2 convert_float2(const float2& v) { return v; }
inline float2 convert_float2(const int2& v) { return float2((float)v.s[0], (float)v.s[1]); }
inline float2 convert_float2(const ushort2& v) { return float2((float)v.s[0], (float)v.s[1]); }
inline float3 convert_float3(const int3& v) { return float3((float)v.s[0], (float)v.s[1], (float)v.s[2]); }
inline float3 convert_float3(const float3& v) { return v; }
inline int3 convert_int3(const float3& v) { return int3(clc_f2i(v.s[0]), clc_f2i(v.s[1]), clc_f2i(v.s[2])); }
inline uint3 convert_uint3(const float3& v) { return uint3(clc_f2u(v.s[0]), clc_f2u(v.s[1]), clc_f2u(v.s[2])); }
inline uint convert_uint(bool v) { return v ? 1u : 0u; }
inline uint convert_uint(int v) { return (uint)v; }
inline uint convert_uint_sat_rtp(float v) {   // saturate, round toward +infinity
    if (!(v > 0.0f)) return 0u;
    float c = ceilf(v);
    if (c >= 4294967296.0f) return 0xffffffffu;
    return (uint)c;
}
inline ushort convert_ushort_sat_rte(float v) { return (ushort)rintf(cpm_clamp(v, 0.0f, 65535.0f)); }
// as_typen: reinterpretation (section 6.2.4); a float3 occupies a float4
inline float4 as_float4(const float3& v) { return float4(v.s[0], v.s[1], v.s[2], v.s[3]); }
inline float4 as_float4(const float4& v) { return v; }
inline int4 as_int4(const int3& v) { return int4(v.s[0], v.s[1], v.s[2], v.s[3]); }
inline int4 as_int4(const int4& v) { return v; }

// ---- integer built-ins (section 6.12.3) -------------------------------------------------------------------------------
inline uint mad_hi(uint a, uint b, uint c) { return (uint)(((uint64_t)a * b) >> 32) + c; }
inline int min(int a, int b) { return b < a ? b : a; }
inline int max(int a, int b) { return a < b ? b : a; }
inline uint min(uint a, uint b) { return b < a ? b : a; }
inline uint max(uint a, uint b) { return a < b ? b : a; }
inline uint clamp(uint x, uint lo, uint hi) { return min(max(x, lo), hi); }
inline int clamp(int x, int lo, int hi) { return min(max(x, lo), hi); }
#define CLC_IMINMAX(V, N)                                                                                     \
    inline V min(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.s[i] = min(a.s[i], b.s[i]); return r; } \
    inline V max(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.s[i] = max(a.s[i], b.s[i]); return r; } \
    inline V clamp(const V& x, const V& lo, const V& hi) { return min(max(x, lo), hi); }
CLC_IMINMAX(int3, 3)
CLC_IMINMAX(int4, 4)

// ---- common / math built-ins (sections 6.12.2, 6.12.4) ---------------------------------------------------------
// fmin / fmax semantics for min / max on floats (6.12.4: "min returns y if y < x, otherwise x"); NaN handling as
// cpm_fmin / cpm_fmax (the non-NaN operand), the definition shared with the oracle.
inline float min(float a, float b) { return cpm_fmin(a, b); }
inline float max(float a, float b) { return cpm_fmax(a, b); }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline float mix(float x, float y, float a) { return fmaf(y - x, a, x); }   // x + (y - x) a
inline float step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
inline float native_log(float x) { return cpm_native_logf(x); }
inline float native_exp(float x) { return cpm_expf_sym(x); }
#define CLC_FCOMMON(V, N)                                                                                         \
    inline V min(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.s[i] = min(a.s[i], b.s[i]); return r; } \
    inline V max(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.s[i] = max(a.s[i], b.s[i]); return r; } \
    inline V clamp(const V& x, const V& lo, const V& hi) { return min(max(x, lo), hi); }                            \
    inline V mix(const V& x, const V& y, float a) { V r; for (int i = 0; i < N; ++i) r.s[i] = mix(x.s[i], y.s[i], a); return r; } \
    inline V floor(const V& a) { V r; for (int i = 0; i < N; ++i) r.s[i] = floorf(a.s[i]); return r; }              \
    inline V fabs(const V& a) { V r; for (int i = 0; i < N; ++i) r.s[i] = fabsf(a.s[i]); return r; }                \
    inline V step(const V& e, const V& x) { V r; for (int i = 0; i < N; ++i) r.s[i] = step(e.s[i], x.s[i]); return r; } \
    inline V step(const V& e, float x) { V r; for (int i = 0; i < N; ++i) r.s[i] = step(e.s[i], x); return r; }     \
    inline V native_log(const V& a) { V r; for (int i = 0; i < N; ++i) r.s[i] = native_log(a.s[i]); return r; }
CLC_FCOMMON(float2, 2)
CLC_FCOMMON(float3, 3)
CLC_FCOMMON(float4, 4)

// ---- geometric built-ins (section 6.12.5) -------------------------------------------------------------------------
// The OpenCL specification leaves the evaluation order and contraction of dot / length to the implementation.  The
// order below is the one the oracle and the CUDA kernels use (fma chain from x), so that a comparison of the
// reference's kernel logic is not blurred by a choice the reference does not make.
inline float dot(const float3& a, const float3& b) { return fmaf(a.s[2], b.s[2], fmaf(a.s[1], b.s[1], a.s[0] * b.s[0])); }
inline float dot(const float2& a, const float2& b) { return fmaf(a.s[1], b.s[1], a.s[0] * b.s[0]); }
inline float length(const float3& a) { return sqrtf(dot(a, a)); }
inline float length(const float2& a) { return sqrtf(dot(a, a)); }
inline float distance(const float3& a, const float3& b) { return length(a - b); }
inline float3 normalize(const float3& a) { float l = length(a); return float3(a.s[0] / l, a.s[1] / l, a.s[2] / l); }
inline float3 cross(const float3& a, const float3& b) {
    return float3(fmaf(a.s[1], b.s[2], -(a.s[2] * b.s[1])), fmaf(a.s[2], b.s[0], -(a.s[0] * b.s[2])),
                  fmaf(a.s[0], b.s[1], -(a.s[1] * b.s[0])));
}

// ---- atomics (section 6.12.11): the drivers run work-items one after the other ---------------------------------
inline uint atomic_cmpxchg(volatile uint* p, uint cmp, uint val) {
    uint old = *p;
    if (old == cmp) *p = val;
    return old;
}

// ---- images (section 6.12.14): only what the kernels touch; the sampling arithmetic itself lives in the Inviwo
// stand-in samplers.cl (Inviwo's header is un-vendored) ---------------------------------------------------------------
struct clc_image {
    const void* data;
    int dims[3];      // width, height, depth
    int format;       // 0 = UNORM_INT8, 1 = UNORM_INT16, 2 = FLOAT (1 channel); 3 = RGBA FLOAT (transfer functions)
};
struct clc_image2d : clc_image {};
typedef const clc_image* image3d_t;
typedef const clc_image2d* image2d_t;
typedef int sampler_t;
inline int4 get_image_dim(image3d_t img) { return int4(img->dims[0], img->dims[1], img->dims[2], 0); }
inline int2 get_image_dim(image2d_t img) { return int2(img->dims[0], img->dims[1]); }
inline int get_image_width(image2d_t img) { return img->dims[0]; }
inline int get_image_height(image2d_t img) { return img->dims[1]; }
