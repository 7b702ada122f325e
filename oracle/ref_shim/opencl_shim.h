// opencl_shim.h -- the handful of OpenCL C names that rng/cl/*.cl of the reference use,
// so that those files compile as C++ for the host.  TEST INFRASTRUCTURE.
#pragma once
#include <stddef.h>
#include <stdint.h>
typedef uint32_t uint;
typedef uint64_t ulong;
struct uint2 { uint x, y; };
static inline uint2 make_uint2(uint a, uint b) { uint2 r; r.x = a; r.y = b; return r; }
struct int2 { int x, y; };
static inline int2 make_int2(int a, int b) { int2 r; r.x = a; r.y = b; return r; }
#define __kernel
#define __global
#define global
#define write_only
typedef void* image2d_t;
static thread_local size_t g_global_id = 0;
static inline size_t get_global_id(int) { return g_global_id; }
static inline uint mad_hi(uint a, uint b, uint c) { return (uint)(((uint64_t)a * b) >> 32) + c; }
// image writes are not exercised by the checker
static inline void write_imagef(image2d_t, int2, float) {}
