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

/* --- selection stage: clogs radix sort semantics, threshold + reduce ------------------------- */
void orc_radix_sort_u32(uint32_t* keys, uint32_t* values, size_t n, unsigned max_bits);
void orc_merge_sort_u32(uint32_t* keys, uint32_t* values, size_t n);
long long orc_count_below(const uint32_t* data, size_t n, uint32_t threshold);

/* --- uniform grids, detector, hash, cell ranges (orc_grid.c) ---------------------------------- */
void orc_volume_minmax(const orc_volume* vol, int region, uint16_t* out);
void orc_volume_diff_bricks(const orc_volume* a, const orc_volume* b, int region, double data_scaling,
                            double range_min, double range_max, float* out);
void orc_classify_importance(const uint16_t* minmax, const uint16_t* prev_minmax, const float* diff, int n,
                             const float* positions, const float* colors, int n_points, const float weights[4],
                             int incremental, float* out);
void orc_detect_invalid(const float* grid, const int grid_dims[3], const float cell_size[3], const float tex2idx[16],
                        const float* photons, int photon_offset, const float* light_samples, const float* isect,
                        int n_light_samples, int max_interactions, int total_photons, uint32_t* importances,
                        int equal_importance, int percentage, int iteration, int fix_exit);
void orc_hash_light_samples(const float* light_samples, const float* isect, int n_light_source_samples,
                            const uint32_t* ids, int n_ids, const float cell_size[3], const int n_blocks[3],
                            uint32_t* which_bucket, int out_offset);
void orc_build_cell_ranges(const uint32_t* sorted_keys, size_t n, uint32_t n_cells, uint32_t* cell_start,
                           uint32_t* cell_end);
/* --- splat density estimation (orc_splat.c), double accumulation ------------------------------ */
void orc_splat(double* vol, int channels, const float tex2idx[16], const float idx2tex[16], const int outDim[3],
               const float* photons, const uint32_t* indices, int n, int photons_per_interaction, int n_interactions,
               float radius, float relative_irradiance_scale, float multiplier);

/* --- photon-map gather (orc_gather.c): parity unpinned, own restatement ------------------------ */
typedef struct orc_gather_params {
    int32_t width, height;
    float cam_origin[3], cam_dir00[3], cam_du[3], cam_dv[3];
    float aabb_min[3], aabb_max[3];
    float step, radius, scale, sigma_scale;
    int32_t grid_dims[3];
} orc_gather_params;
void orc_photon_cell_keys(const float* photons, size_t n, const int grid_dims[3], uint32_t* keys);
void orc_gather_points(const orc_gather_params* P, const float* photons, size_t n_records, const float* points,
                       int n_points, float* irradiance);
void orc_gather_raymarch(const orc_volume* vol, const float* tf_rgba, int tf_width, const orc_gather_params* P,
                         const float* photons, size_t n_records, float* image);

/* final image from the light volume (the LightingRaycaster step of the workspace network): parity unpinned */
void orc_raycast_light_volume(const orc_volume* vol, const float* tf_rgba, int tf_width, const orc_gather_params* P,
                              const float* light_volume, const int lv_dims[3], int channels, float* image);

/* --- view importance + importance-driven sample generator (orc_importance.c) ---------------------- */
void orc_view_importance(const uint16_t* minmax, const int dims[3], const float cellDim[3], const float tex2idx[16],
                         const float idx2tex[16], const float* entry, const float* exit, int width, int height,
                         float tf_min, float tf_max, float* out);
void orc_sample_importance2d(const float* importance, int w, int h, float floor_value, const float* uniform_samples,
                             int n, float* out);

void orc_selftest_math(int fn, const float* x, const float* y, float* out, size_t n);
int orc_convex_hull2d(const float* pts, int n, float* hull_out);
int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
