/*
 * cpm_oracle.h -- CPU restatement of the reference's correlated photon-mapping hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (the package directory, include/)
 * links, loads or calls this library.  It is used by tests/, by __graft_entry__.smoke() as
 * the checker, and by bench.py's cpu_baseline / --impl reference legs as the timed CPU arm.
 *
 * Parity status: the reference ships no tests, golden vectors or fixtures,
 * and its OpenCL kernels cannot run here (no OpenCL implementation, 13 un-vendored Inviwo
 * headers).  Pinning therefore is:
 *   - MWC64X: PINNED against the reference's own source.  oracle/Makefile compiles
 *     rng/cl/skip_mwc.cl, random.cl, randstategen.cl and randomnumbergenerator.cl (where
 *     they lie under /root/reference) into oracle/_ref/libmwc64x_ref.so; tests compare this
 *     restatement with it and with golden vectors generated from it (tests/golden/).
 *   - sort / count / iota: pinned by definition (unique stable permutation).
 *   - everything that passes through un-vendored Inviwo headers (volume sampling, ray-box,
 *     phase functions, direction codec, mesh intersection, 16-bit min/max write):
 *     PARITY UNPINNED -- restated from the OpenCL 1.2 specification and the call sites.
 *
 * All float arithmetic is fp32 with explicit fmaf and no contraction (-ffp-contract=off);
 * transcendentals come from include/cpm_detmath.h, the shared definition of native_log etc.
 */
#ifndef CPM_ORACLE_H
#define CPM_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* --- MWC64X: rng/cl/random.cl:44-95, rng/cl/skip_mwc.cl:40-105 ------------------------- */
void orc_rng_host_base_offsets(uint32_t seed, uint32_t* state, size_t n); /* uses libc srand/rand */
void orc_rng_seed_streams(uint32_t* state, size_t n, uint64_t gap, uint64_t first_stream);
void orc_rng_uniform(uint32_t* state, size_t n, int per_stream, float* out);
void orc_rng_step(uint32_t* x, uint32_t* c);

/* --- emission ---------------------------------------------------------------------------- */
void orc_sample_uniform2d(float nx, float ny, int n, float* out);
void orc_light_sample_directional(const float* samples, const float radiance[3], const float dir[3],
                                  const float origin[3], const float u[3], const float v[3], float area,
                                  int n, float* out);
void orc_light_sample_point(const float* samples, const float radiance[3], const float pos[3], int n,
                            float* out);
void orc_light_mesh_intersect(const float* vertices, const int32_t* indices, int n_indices,
                              const float* light_samples, int n, float* out);
/* CPU fit of the light plane: lcl/orientedboundingbox2d.cpp:40-100, lcl/convexhull2d.cpp:38-130,
 * lcl/pointplaneprojection.cpp:39-54.  out = origin[3], u[3], v[3]. */
void orc_fit_light_plane(const float* points, int n_points, const float plane_point[3],
                         const float plane_normal[3], float out[9]);

/* --- volume sampling (restated Inviwo samplers.cl) ------------------------------------------ */
typedef struct orc_volume {
    const void* data;
    int dims[3];
    int format; /* 0 u8, 1 u16, 2 f32 */
    float scale, offset;
} orc_volume;
float orc_sample_volume(const orc_volume* v, float px, float py, float pz);
float orc_sample_tf_alpha(const float* tf_rgba, int width, float v);

/* --- tracer: ppm/cl/photontracer.cl:69-216, ppm/cl/transmittance.cl:126-144 -------------------- */
typedef struct orc_trace_params {
    float aabb_min[3], aabb_max[3];
    float material[4];
    int32_t phase_function;
    float step_size;
    int32_t max_interactions;
    int32_t photon_offset;
    int32_t total_photons;
    int32_t n_light_samples;
    uint32_t flags;
} orc_trace_params;
/* returns the number of delta-tracking collision tests executed; n_threads <= 0: all cores */
unsigned long long orc_trace_photons(const orc_volume* vol, const float* tf_rgba, int tf_width,
                                     const orc_trace_params* p, const float* light_samples,
                                     const float* isect, const uint32_t* recompute, int n_recompute,
                                     float* photons, uint32_t* rng, int n_threads);

void orc_selftest_math(int fn, const float* x, const float* y, float* out, size_t n);
int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
