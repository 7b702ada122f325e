/*
 * cpm_b200.h -- C ABI of libcpm_b200.so: the B200-native (sm_100a) implementation of the
 * correlated progressive photon-mapping hot path.
 *
 * The reference has no C ABI of its own (it is six Inviwo C++/OpenCL modules); the natural
 * cut is its layer of kernel-launcher classes, whose methods take raw device buffers and
 * plain scalars.  Every entry point below names the launcher method (and the OpenCL kernel
 * behind it) that it replaces.  Paths are relative to /modules of the reference:
 *   ppm/ = progressivephotonmapping/   lcl/ = lightcl/   isc/ = importancesamplingcl/
 *   ugc/ = uniformgridcl/   rsc/ = radixsortcl/   rng/ = rndgenmwc64x/
 *
 * Conventions
 *  - Every pointer argument is a DEVICE pointer owned by the caller unless its name ends
 *    in `_host`.  Nothing is retained after the call returns unless documented.
 *  - Calls are asynchronous on the context's CUDA stream (like enqueueNDRangeKernel on the
 *    reference's in-order queue) except the ones documented as synchronous.
 *  - Return value: 0 (CPM_OK) or a negative CPM_E_* code; cpm_last_error() gives the text.
 *    No exception crosses the ABI.
 *  - One context per GPU; a context is not thread-safe (the reference runs on Inviwo's
 *    single evaluation thread).
 *  - There is no CPU fallback: without a CUDA device cpm_ctx_create fails with
 *    CPM_E_NO_DEVICE.
 */
#ifndef CPM_B200_H
#define CPM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define CPM_API
#else
#define CPM_API __attribute__((visibility("default")))
#endif

/* ---- error codes ------------------------------------------------------------------- */
enum {
    CPM_OK = 0,
    CPM_E_INVALID = -1,     /* bad argument (mirrors cl::Error CL_INVALID_VALUE) */
    CPM_E_NO_DEVICE = -2,   /* no CUDA device / wrong architecture */
    CPM_E_CUDA = -3,        /* a CUDA runtime call failed */
    CPM_E_NOMEM = -4,       /* device allocation failed */
    CPM_E_UNSUPPORTED = -5, /* valid in the reference but not built here */
    CPM_E_COMM = -6         /* NCCL failure */
};

typedef struct cpm_ctx cpm_ctx;
typedef struct cpm_volume cpm_volume;

/* ---- context ----------------------------------------------------------------------- */
/* stream: a cudaStream_t cast to void*, or NULL to let the context create its own
 * non-blocking stream.  Replaces OpenCL::getPtr()->getQueue(). */
CPM_API int cpm_ctx_create(int device, void* stream, cpm_ctx** out);
CPM_API void cpm_ctx_destroy(cpm_ctx* ctx);
CPM_API void* cpm_ctx_stream(cpm_ctx* ctx);
CPM_API int cpm_ctx_sync(cpm_ctx* ctx); /* cudaStreamSynchronize */
CPM_API const char* cpm_last_error(cpm_ctx* ctx);
CPM_API const char* cpm_version(void);
/* number of kernel launches issued through this context since creation / last reset */
CPM_API uint64_t cpm_ctx_launch_count(cpm_ctx* ctx, int reset);

/* ---- stage timing ------------------------------------------------------------------- */
/* CUDA events on the context stream: what IVW_OPENCL_PROFILING / cl::Event timestamps are in the
 * reference (ppm/processor/progressivephotontracercl.cpp:562-598).  cpm_event_record is
 * asynchronous; cpm_event_elapsed_ms waits for `end` and returns the device time between the two. */
typedef struct cpm_event cpm_event;
CPM_API int cpm_event_create(cpm_ctx* ctx, cpm_event** out);
CPM_API int cpm_event_record(cpm_ctx* ctx, cpm_event* ev);
CPM_API int cpm_event_elapsed_ms(cpm_ctx* ctx, cpm_event* begin, cpm_event* end, float* ms_host);
CPM_API void cpm_event_destroy(cpm_ctx* ctx, cpm_event* ev);

/* ---- (1) MWC64X per-photon RNG streams -------------------------------------------- */
/* Fill the per-stream base offsets exactly as MWC64XSeedGenerator::generateRandomSeeds
 * does on the host (rng/mwc64xseedgenerator.cpp:56-64): srand(seed); state[i].x = rand().
 * Only .x is written by the reference; we also zero .y.  Host-side, synchronous.
 * glibc's rand() is reproduced bit-exactly (TYPE_3 additive feedback generator). */
CPM_API int cpm_rng_host_base_offsets(uint32_t seed, uint32_t* state_host /* 2*n */, size_t n);
/* The same sequence starting at its element `first`: the base offsets of photons
 * [first, first + n) of a larger photon set, for a GPU that owns that photon range. */
CPM_API int cpm_rng_host_base_offsets_range(uint32_t seed, uint64_t first, uint32_t* state_host /* 2*n */,
                                            size_t n);

/* MWC64X_GenerateRandomState (rng/cl/randstategen.cl:39-47): in place,
 * state[i] = SeedStreams(baseOffset = state[i].x, perStreamOffset = stream_gap) for stream
 * index first_stream + i.  stream_gap = 2^40 reproduces the reference kernel;
 * MWC64X_GeneratePerStreamRandomState (:52-60) is the same with a caller-supplied gap.
 * first_stream lets a GPU seed its own shard of a larger stream set. */
CPM_API int cpm_rng_seed_streams(cpm_ctx* ctx, uint32_t* state /* uint2[n] */, size_t n,
                                 uint64_t stream_gap, uint64_t first_stream);

/* randomNumberGeneratorKernel (rng/cl/randomnumbergenerator.cl:34-49): one random_01 per
 * stream, state advanced and saved.  samples_per_stream generalises N_NUMBERS_PER_THREAD. */
CPM_API int cpm_rng_uniform(cpm_ctx* ctx, uint32_t* state, size_t n, int samples_per_stream,
                            float* out /* n*samples_per_stream */);

/* ---- (2) emission ------------------------------------------------------------------ */
/* uniformSampleGenerator2DKernel (isc/cl/uniformsamplegenerator2d.cl:35-52).
 * out[i] = ((0.5 + fmod(i, nx)) / nx, (0.5 + i / nx) / ny, 0, 1)  (y is NOT floored). */
CPM_API int cpm_sample_uniform2d(cpm_ctx* ctx, float nx, float ny, int n_elements,
                                 float* out /* float4[n] */);

/* directionalLightSamplerKernel (lcl/cl/directionallightsampler.cl:38-63), host side
 * DirectionalLightSamplerCL::sampleLightSource (lcl/directionallightsamplercl.cpp:114-130).
 * light_samples is the StoredLightSample float8 buffer {origin.xyz, power.rgb, theta, phi}. */
CPM_API int cpm_light_sample_directional(cpm_ctx* ctx, const float* samples /* float4[n] */,
                                         const float radiance[3], const float direction[3],
                                         const float plane_origin[3], const float plane_u[3],
                                         const float plane_v[3], float plane_area, int n,
                                         float* light_samples /* float8[n] */);

/* Point-light emission restated from sampleLight's LIGHT_POINT branch
 * (isc/cl/light/light.cl:84-92): origin = position, wi = -uniformSampleSphere(uv),
 * power = radiance / (1/4pi).  No reference processor launches it ("parity unpinned"). */
CPM_API int cpm_light_sample_point(cpm_ctx* ctx, const float* samples /* float4[n] */,
                                   const float radiance[3], const float position[3], int n,
                                   float* light_samples /* float8[n] */);

/* lightSampleMeshIntersectionKernel (lcl/cl/intersection/lightsamplemeshintersection.cl:37-59),
 * host side LightSampleMeshIntersectionCL::meshSampleIntersection
 * (lcl/lightsamplemeshintersectioncl.cpp:87-99).  Miss -> (0, -1). */
CPM_API int cpm_light_mesh_intersect(cpm_ctx* ctx, const float* vertices /* float3 packed */,
                                     const int32_t* indices, int n_indices,
                                     const float* light_samples, int n,
                                     float* intersections /* float2[n] */);

/* View (or light) importance image: uniformGridImportanceKernel
 * (isc/cl/minmaxuniformgrid3dimportance.cl:336-378 with uniformGridImportance :86-133); host side
 * MinMaxUniformGrid3DImportanceCL::computeImportance (isc/minmaxuniformgrid3dimportancecl.cpp:99-132).  Per
 * pixel: length of the segment entry -> exit (texture coordinates, float4 per pixel) that lies in min-max
 * bricks overlapping [tf_min, tf_max].  minmax = ushort2 per brick as written by cpm_volume_minmax. */
CPM_API int cpm_view_importance(cpm_ctx* ctx, const uint16_t* minmax, const int grid_dims[3], const float cell_size[3],
                                const float texture_to_index[16], const float index_to_texture[16],
                                const float* entry /* float4[w*h] */, const float* exit /* float4[w*h] */, int width,
                                int height, float tf_min, float tf_max, float* importance /* w*h */);
/* Importance-driven 2-D sample generator behind the SampleGenerator2DCL interface
 * (lcl/samplegenerator2dcl.h:53-88; the reference ships only the uniform implementation): warps uniform
 * samples (u, v, w, pdf) through the inverse CDF of the piecewise-constant density
 * max(importance, 0) + floor_value on a width x height grid; the output pdf slot is multiplied by the density
 * (integral 1 over the unit square), which cpm_light_sample_directional divides the power by.
 * cdf_scratch: cpm_sample_importance2d_scratch_floats(width, height) floats. */
CPM_API size_t cpm_sample_importance2d_scratch_floats(int width, int height);
CPM_API int cpm_sample_importance2d(cpm_ctx* ctx, const float* importance, int width, int height, float floor_value,
                                    const float* uniform_samples /* float4[n] */, int n, float* cdf_scratch,
                                    float* samples_out /* float4[n] */);

/* ---- volumes ----------------------------------------------------------------------- */
enum { CPM_FMT_U8 = 0, CPM_FMT_U16 = 1, CPM_FMT_F32 = 2 };
enum {
    CPM_VOLUME_LINEAR = 0, /* sample straight from the caller's x-fastest linear buffer */
    CPM_VOLUME_TEXTURE = 1 /* copy into a 2-D layered CUDA array; the tracer then fetches
                              bilinear footprints with tld4 (unfiltered texels) */
};
/* Replaces Volume::getRepresentation<VolumeCL>() + VolumeCLBase::getVolumeStruct
 * (ppm/photontracercl.cpp:108-119).  A voxel read returns
 *   (normalise(v) + format_offset) * format_scale
 * where normalise is v/255, v/65535 or identity ("Scaling for 12-bit data",
 * ugc/processors/volumeminmaxclprocessor.cpp:134-136).  The LINEAR layout keeps a pointer
 * to `data`; the caller must keep it alive.  TEXTURE copies. */
CPM_API int cpm_volume_create(cpm_ctx* ctx, const void* data, const int dims[3], int format,
                              float format_scale, float format_offset, int layout,
                              cpm_volume** out);
/* new voxel data, same dims/format (a time step change) */
CPM_API int cpm_volume_update(cpm_ctx* ctx, cpm_volume* vol, const void* data);
CPM_API void cpm_volume_destroy(cpm_ctx* ctx, cpm_volume* vol);

/* ---- (3) photon tracer -------------------------------------------------------------- */
enum {
    CPM_TRACE_PROGRESSIVE = 1,          /* -D PROGRESSIVE_PHOTON_MAPPING: save RNG state */
    CPM_TRACE_NO_SINGLE_SCATTERING = 2, /* -D NO_SINGLE_SCATTERING */
    CPM_TRACE_STATS = 4,                /* collision_tests points to TWO counters: [0] collision tests,
                                           [1] tests that fetched voxels (== [0] without an opacity bound) */
    CPM_TRACE_LANE_REFILL = 8           /* scheduling only (same photons): the wavefront form -- set-up kernel, persistent
                                           walk kernel whose idle lanes take the next walk from a cursor, interaction
                                           kernel, one round per scattering event.  On C4 0.594 vs 0.559 ms for the
                                           one-kernel form; a tested option for long walks.  Needs an
                                           opacity bound; ignored with NO_SINGLE_SCATTERING */
};
enum { CPM_PHASE_ISOTROPIC = 0, CPM_PHASE_HENYEY_GREENSTEIN = 1 };

typedef struct cpm_trace_params {
    float aabb_min[3];     /* clip box in texture space (axisAlignedBoundingBoxCL_,       */
    float aabb_max[3];     /*   ppm/processor/progressivephotontracercl.cpp:192-195,678-684) */
    float material[4];     /* AdvancedMaterialProperty::getCombinedMaterialParameters();
                              [0] = anisotropy g for Henyey-Greenstein */
    int32_t phase_function; /* CPM_PHASE_* (material.getPhaseFunctionEnum()) */
    float step_size;       /* samplingRate * min voxel spacing (:236-239) */
    int32_t max_interactions; /* maxScatteringEvents, 1..16 */
    int32_t photon_offset; /* first photon id of this light source */
    int32_t total_photons; /* photonData->getNumberOfPhotons(): stride between interactions */
    int32_t n_light_samples;
    uint32_t flags;        /* CPM_TRACE_* */
    /* Optional per-cell opacity bound written by cpm_opacity_bound for THIS volume and THIS transfer function
     * (NULL = test every sample like the reference).  Results are identical either way; with a bound, tests whose
     * second random number is >= the bound of their cell are decided without fetching voxels. */
    const float* opacity_bound;
    int32_t bound_cell_log2; /* the cell_log2 the bound grid was built with */
    int32_t reserved_;
    /* Optional: the same bound grid as a point-sampled 3-D texture (cpm_bound_tex_create / _update; bound_cell_log2
     * applies).  When set it is used instead of opacity_bound: one TEX per collision test instead of the index
     * arithmetic and the load.  Results are identical. */
    const struct cpm_bound_tex* opacity_bound_tex;
} cpm_trace_params;

/* photonTracerKernel (ppm/cl/photontracer.cl:69-216) incl. woodcockTracking
 * (ppm/cl/transmittance.cl:126-144); host side PhotonTracerCL::tracePhotons
 * (ppm/photontracercl.cpp:136-174).
 *  tf_rgba         the transfer function rasterised to tf_width RGBA float texels (both
 *                  tfData and tfScattering, as the reference binds the same layer twice,
 *                  ppm/photontracercl.cpp:150-151)
 *  recompute_index NULL for the plain kernel; otherwise the -D PHOTON_RECOMPUTATION build:
 *                  n_recompute photon ids, ids outside [photon_offset,
 *                  photon_offset + n_light_samples) are skipped (photontracer.cl:99-106)
 *  photons         float8 records, index photon_offset + k*total_photons + i
 *  rng_state       uint2 per photon of the WHOLE photon set (indexed photon_offset + i)
 *  collision_tests optional device uint64 counter (may be NULL): incremented by the number
 *                  of delta-tracking collision tests executed -- the benchmark's
 *                  "photon-interaction" unit.  Not in the reference. */
CPM_API int cpm_trace_photons(cpm_ctx* ctx, const cpm_volume* vol, const float* tf_rgba,
                              int tf_width, const cpm_trace_params* params,
                              const float* light_samples, const float* intersections,
                              const uint32_t* recompute_index, int n_recompute, float* photons,
                              uint32_t* rng_state, unsigned long long* collision_tests);

/* ---- (3b) opacity-bound grid for the tracer ("L2-resident bricks") ------------------------ */
/* Not in the reference, which fetches 8 voxels + the transfer function for every collision test
 * (ppm/cl/transmittance.cl:134-140).  A test rejects when u2 >= opacity; with an upper bound of the opacity over
 * the cell the sample falls in, u2 >= bound decides the same without the fetch.  Random stream, positions and
 * photons are unchanged (tests compare the bounded tracer bit for bit with the unbounded one and the oracle).
 *
 * Cells have 2^cell_log2 voxels per axis; grid dims = (dims >> cell_log2) + 1 (cpm_bound_grid_dims).  Cell c
 * holds the trilinear footprints whose lower tap i0 has (i0 + 1) >> cell_log2 == c.
 * cpm_volume_value_range: range[c] = float2 (min, max) of the normalised voxel values (v/255, v/65535, v) those
 * footprints touch; NaN pair when a voxel is NaN/inf.  Needs a LINEAR volume.  One pass per volume / time step. */
CPM_API int cpm_bound_grid_dims(const int dims[3], int cell_log2, int out_dims[3]);
CPM_API int cpm_volume_value_range(cpm_ctx* ctx, const cpm_volume* vol, int cell_log2, float* range /* float2 per cell */,
                                   int out_dims[3]);
/* bound[c] >= alpha of tf_rgba at every value (v + format_offset) * format_scale, v in range[c], including the
 * rounding of the fp32 trilinear / TF blends; +inf (always fetch) where NaN/inf makes that impossible.  Re-run
 * when the volume (range) or the transfer function changes. */
CPM_API int cpm_opacity_bound(cpm_ctx* ctx, const float* range, size_t n_cells, float format_scale, float format_offset,
                              const float* tf_rgba, int tf_width, float* bound);

/* The bound grid as a point-sampled 3-D texture for cpm_trace_params::opacity_bound_tex.  grid_dims =
 * cpm_bound_grid_dims; cpm_bound_tex_update copies `bound` (device memory, the layout cpm_opacity_bound writes) into
 * the texture on the context stream -- call it after every cpm_opacity_bound.  Not in the reference. */
typedef struct cpm_bound_tex cpm_bound_tex;
CPM_API int cpm_bound_tex_create(cpm_ctx* ctx, const int grid_dims[3], cpm_bound_tex** out);
CPM_API int cpm_bound_tex_update(cpm_ctx* ctx, cpm_bound_tex* tex, const float* bound);
CPM_API void cpm_bound_tex_destroy(cpm_bound_tex* tex);

/* Clearance of transparent cells, in place: a cell whose bound is exactly 0 (every transfer-function texel it can
 * reach is 0) and whose surrounding cube of radius R >= 1 cells (R <= max_radius) holds only such cells gets the
 * value -R: the next R cell widths of any ray through the cell are transparent.  Optional annotation for ray
 * marchers that want to leave transparent regions in one jump (tracer and gather treat every bound <= 0 as
 * transparent; the tracer does not use R -- measured slower, see csrc/tracer.cu).  Cells with a positive bound
 * are untouched.  grid_dims = cpm_bound_grid_dims. */
CPM_API int cpm_opacity_bound_clearance(cpm_ctx* ctx, float* bound, const int grid_dims[3], int max_radius);

/* ---- (5) selection: threshold / count / iota / radix sort --------------------------- */
/* thresholdKernel (ppm/cl/threshold.cl:33-40): out[i] = data[i] < threshold. */
CPM_API int cpm_threshold_u32(cpm_ctx* ctx, const uint32_t* data, uint32_t threshold, size_t n,
                              uint32_t* out);
/* indexToBufferKernel (ppm/cl/indextobuffer.cl:33-40): out[i] = i. */
CPM_API int cpm_iota_u32(cpm_ctx* ctx, uint32_t* out, size_t n);
/* clogs::Reduce::enqueue with TYPE_INT (rsc/ext/clogs/src/reduce.cpp:262-345) as called by
 * ProgressivePhotonTracerCL::reduceInts (ppm/processor/progressivephotontracercl.cpp:727-741).
 * SYNCHRONOUS: returns the exact sum in *result_host (the reference reads its result before
 * the read-back event completes, :342-345,374 -- fixed here). */
CPM_API int cpm_reduce_sum_i32(cpm_ctx* ctx, const int32_t* data, size_t n, long long* result_host);
/* Fused replacement of threshold + reduce + iota for the re-trace selection
 * (ppm/processor/progressivephotontracercl.cpp:318-356): one 4 B/photon pass that counts
 * data[i] < threshold and, if iota_out != NULL, writes iota_out[i] = i.  SYNCHRONOUS. */
CPM_API int cpm_count_below(cpm_ctx* ctx, const uint32_t* data, size_t n, uint32_t threshold,
                            uint32_t* iota_out, long long* count_host);

/* Stable selection: ids_out[0 .. count) = the indices i < n with data[i] < threshold, ascending; entries from
 * count on are left untouched.  This is what threshold + count + sort-by-importance + keys-only id sort
 * (ppm/processor/progressivephotontracercl.cpp:318-473) produce when the update budget covers every invalid
 * photon and `spatialSorting` is on: all invalid ids, ascending -- one 4 B/photon pass instead of two radix
 * sorts.  SYNCHRONOUS (returns the count). */
CPM_API int cpm_select_below(cpm_ctx* ctx, const uint32_t* data, size_t n, uint32_t threshold, uint32_t* ids_out,
                             long long* count_host);
/* The same in two halves, so that the caller can enqueue work that does not depend on the count (e.g. the tracer's
 * opacity-bound refresh) while the count travels to the host: _begin launches the selection and the copy of the count,
 * _end waits for that copy only -- not for work enqueued after _begin -- and returns the count.  No other call that
 * returns a value synchronously (cpm_count_below, cpm_reduce_sum_i32, cpm_select_below) may come in between. */
CPM_API int cpm_select_below_begin(cpm_ctx* ctx, const uint32_t* data, size_t n, uint32_t threshold, uint32_t* ids_out);
CPM_API int cpm_select_below_end(cpm_ctx* ctx, long long* count_host);

/* clogs::Radixsort::enqueue (rsc/ext/clogs/radixsort.h:227-262, src/radixsort.cpp:169-259) for
 * TYPE_UINT keys with TYPE_UINT values (values == NULL: keys only, as recomputationIndexSorter_).
 * Ascending, stable, in place; max_bits = 0 means all 32 bits; tmp_* are the ping-pong buffers
 * of setTemporaryBuffers (same sizes as keys / values).  Errors as clogs: n == 0 and
 * max_bits > 32 -> CPM_E_INVALID.  Hand-written onesweep (8-bit digits, decoupled look-back). */
CPM_API int cpm_radix_sort_u32(cpm_ctx* ctx, uint32_t* keys, uint32_t* values, size_t n,
                               unsigned max_bits, uint32_t* tmp_keys, uint32_t* tmp_values);
/* bytes of context scratch a sort of n elements uses (status words of the chained scan) */
CPM_API size_t cpm_radix_sort_scratch_bytes(size_t n);

/* ---- (4) temporal-correlation detector ---------------------------------------------- */
enum { CPM_DETECT_FIX_EXIT = 1 /* use origin + tEnd*direction where the reference has
                                  `exit = tEnd*direction` (photonrecomputationdetector.cl:128) */ };
/* photonRecomputationDetectorKernel / ...EqualImportanceKernel
 * (ppm/cl/photonrecomputationdetector.cl:92-157, 160-194); host side
 * PhotonRecomputationDetector::photonRecomputationImportance
 * (ppm/photonrecomputationdetector.cpp:93-121).  importances[photon_offset + i] -=
 * min(2^31-1, ceil(100 * sum over path segments of DDA(importance grid))).
 *  importance_grid   float per cell, id = x + y*dx + z*dx*dy (ugc/uniformgrid3d.h:56-62)
 *  cell_size         UniformGrid3DBase::getCellDimension() as floats
 *  texture_to_index  column-major 4x4 of the ORIGINAL volume (getTextureToIndexMatrix) */
CPM_API int cpm_detect_invalid(cpm_ctx* ctx, const float* importance_grid, const int grid_dims[3],
                               const float cell_size[3], const float texture_to_index[16],
                               const float* photons, int photon_offset, const float* light_samples,
                               const float* intersections, int n_light_samples, int max_interactions,
                               int total_photons, uint32_t* importances, int equal_importance,
                               int percentage, int iteration, uint32_t flags);

/* ---- (6) uniform grids ---------------------------------------------------------------- */
/* volumeMinMaxKernel (ugc/cl/uniformgrid/volumeminmax.cl:33-61), host side
 * VolumeMinMaxCLProcessor::compute (ugc/processors/volumeminmaxclprocessor.cpp:149-184).
 * out: ushort2 (min, max) * 65535 per brick, out_dims = ceil(dims / region) (may be NULL).
 * Needs a LINEAR volume. */
CPM_API int cpm_volume_minmax(cpm_ctx* ctx, const cpm_volume* vol, int region, uint16_t* out,
                              int out_dims[3]);
/* DynamicVolumeDifferenceAnalysis (ugc/processors/dynamicvolumedifferenceanalysis.h:96-151,
 * .cpp:60-104) -- a CPU loop in the reference -- on the GPU: per brick
 * (sum |data_scaling*(b - a)| / region^3 - range_min) / (range_max - range_min). */
CPM_API int cpm_volume_diff_bricks(cpm_ctx* ctx, const cpm_volume* a, const cpm_volume* b, int region,
                                   double data_scaling, double range_min, double range_max,
                                   float* out);
/* classifyMinMaxUniformGrid3DImportanceKernel (prev_minmax == NULL) and
 * classifyTimeVaryingMinMaxUniformGrid3DImportanceKernel
 * (isc/cl/minmaxuniformgrid3dimportance.cl:269-330); host side computeImportance
 * (isc/processors/minmaxuniformgrid3dimportanceclprocessor.cpp:218-297).
 * weights = {colorWeight, colorDiffWeight, opacityDiffWeight, opacityWeight} already
 * normalised by the host; incremental != 0 selects -D INCREMENTAL_TF_IMPORTANCE. */
CPM_API int cpm_classify_importance(cpm_ctx* ctx, const uint16_t* minmax, const uint16_t* prev_minmax,
                                    const float* volume_diff, int n, const float* tf_positions,
                                    const float* tf_colors /* float4[n_points] */, int n_points,
                                    const float weights[4], int incremental, float* out);
/* hashLightSampleKernel (ppm/cl/hashlightsample.cl:38-66): cell index of each listed light
 * sample's entry point, the spatial sort key of the HASH_SORT_PHOTONS build. */
CPM_API int cpm_hash_light_samples(cpm_ctx* ctx, const float* light_samples, const float* intersections,
                                   int n_light_source_samples, const uint32_t* ids, int n_ids,
                                   const float cell_size[3], const int n_blocks[3],
                                   uint32_t* which_bucket, int out_offset);
/* Cell ranges over ascending keys: cell_start[c] = lower_bound(c), cell_end[c] = upper_bound(c)
 * for every c < n_cells (keys >= n_cells are ignored).  Not in the reference;
 * parity is against the oracle's restatement. */
CPM_API int cpm_build_cell_ranges(cpm_ctx* ctx, const uint32_t* sorted_keys, size_t n, uint32_t n_cells,
                                  uint32_t* cell_start, uint32_t* cell_end);

/* ---- (7) density estimation: splat to the light volume ---------------------------------- */
/* splatPhotonsToLightVolumeKernel (indices == NULL; photon ids [0, n)) and
 * splatSelectedPhotonsToLightVolumeKernel (n indices, every interaction, times multiplier)
 * (ppm/cl/photonstolightvolume.cl:139-202); host side executeVolumeOperation /
 * photonsToLightVolume (ppm/processor/photontolightvolumeprocessorcl.cpp:356-472).
 * light_volume: float[dims] (channels = 1) or float4[dims] (channels = 4, alpha untouched).
 * Adds are L2 reductions (red.global.add.f32) in unspecified order, like the reference's CAS adds. */
CPM_API int cpm_splat_photons(cpm_ctx* ctx, float* light_volume, int channels,
                              const float texture_to_index[16], const float index_to_texture[16],
                              const int out_dims[3], const float* photons, const uint32_t* indices,
                              int n, int photons_per_interaction, int n_interactions, float radius,
                              float relative_irradiance_scale, float multiplier);

/* The incremental update of photonsToLightVolume (ppm/processor/photontolightvolumeprocessorcl.cpp:262-274: one
 * splatSelected pass with multiplier -1 over the previous records, one with +1 over the new ones) as ONE pass:
 * per listed id and interaction, remove old_photons' contribution and add new_photons'; records the re-trace
 * reproduced bit for bit are skipped (their two contributions cancel). */
CPM_API int cpm_splat_photons_update(cpm_ctx* ctx, float* light_volume, int channels,
                                     const float texture_to_index[16], const float index_to_texture[16],
                                     const int out_dims[3], const float* old_photons, const float* new_photons,
                                     const uint32_t* indices, int n, int photons_per_interaction, int n_interactions,
                                     float radius, float relative_irradiance_scale);

/* The same update, and afterwards old_photons holds the NEW records of every listed id (all interactions): the copy of
 * the photon buffer that the next incremental update subtracts (ppm/processor/photontolightvolumeprocessorcl.cpp:
 * 488-497 copies the whole buffer after every frame) stays current without a pass over all photons.  Ids must be
 * unique, as the tracer's recomputed-index list is. */
CPM_API int cpm_splat_photons_update_sync(cpm_ctx* ctx, float* light_volume, int channels,
                                          const float texture_to_index[16], const float index_to_texture[16],
                                          const int out_dims[3], float* old_photons, const float* new_photons,
                                          const uint32_t* indices, int n, int photons_per_interaction, int n_interactions,
                                          float radius, float relative_irradiance_scale);

/* copyIndexPhotonsKernel (ppm/cl/photonstolightvolume.cl:225-247; host: photontolightvolumeprocessorcl.cpp:207-244, the
 * `alignChangedPhotons` path): aligned_photons[out_offset + g + k * n] = record k * photons_per_interaction + indices[g]
 * with its power multiplied by `multiplier` (-1 for the previous records, +1 for the new ones), g < n, k < n_interactions.
 * The packed buffer is then splatted as plain records (cpm_splat_photons, indices == NULL). */
CPM_API int cpm_copy_index_photons(cpm_ctx* ctx, const float* photons, const uint32_t* indices, int n, float multiplier,
                                   int photons_per_interaction, int n_interactions, float* aligned_photons, size_t out_offset);

/* ---- multi-GPU exchange (option B) --------------------------------------------------------------- */
/* Sum of the per-GPU light volumes over NVLink peer memory, in place: peer_buffers[r] (r < world; host array of device
 * pointers, this GPU's mapping of rank r's buffer, all n_floats long) each hold one rank's volume on entry and the sum
 * on return.  Rank `rank` reduces the rank-th slice and stores it to every buffer.  multicast != NULL: the NVSwitch
 * multicast mapping of the same buffers -- the reduction and the broadcast are done by the switch (multimem.ld_reduce /
 * multimem.st) and peer_buffers may be NULL; otherwise peer loads in rank order (every rank gets the same bits) and peer
 * stores.  No inter-GPU synchronisation inside: the caller orders "all ranks wrote their buffer" -> this call -> "all
 * ranks finished" with barriers on the same stream.  max_ctas <= 0: default (48).  Replaces nothing in the reference
 * (single GPU); the role of the proposed cpm_allreduce_lightvol. */
#define CPM_MAX_PEERS 8
CPM_API int cpm_allreduce_peer_f32(cpm_ctx* ctx, float* const* peer_buffers, float* multicast, size_t n_floats,
                                   int rank, int world, int max_ctas);

/* ---- multi-GPU communicator (cpm_comm_init / cpm_allreduce_lightvol / cpm_allgather_photons) ----------- */
/* One communicator per process (= per GPU, one cpm_ctx).  A C / C++ host runs the sharded path with these calls alone:
 *   rank 0: cpm_comm_unique_id(id); the application hands the 128 bytes to every rank (MPI, a file, a socket ...);
 *   every rank: cpm_comm_init(ctx, id, rank, world, &comm).
 * NCCL is loaded at run time (libnccl.so.2; CPM_NCCL_LIBRARY overrides the name) -- no link-time dependency, single-GPU
 * users never load it.  cpm_comm_init_nccl adopts a communicator the application already owns (an ncclComm_t).
 * Nothing in the reference corresponds to these (it is a single-GPU program); photon ranges per rank come from its own
 * photonOffset / totalPhotons kernel arguments (ppm/cl/photontracer.cl:123,166), see cpmh_runtime_init.
 * All calls are collective over the communicator and asynchronous on the context's stream. */
typedef struct cpm_comm cpm_comm;
#define CPM_COMM_ID_BYTES 128
CPM_API int cpm_comm_unique_id(void* id_out /* CPM_COMM_ID_BYTES */);
CPM_API int cpm_comm_init(cpm_ctx* ctx, const void* unique_id, int rank, int world, cpm_comm** out);
CPM_API int cpm_comm_init_nccl(cpm_ctx* ctx, void* nccl_comm, int rank, int world, cpm_comm** out);
CPM_API void cpm_comm_destroy(cpm_comm* comm);
CPM_API int cpm_comm_rank(const cpm_comm* comm);
CPM_API int cpm_comm_world(const cpm_comm* comm);
/* "nccl" or "peer kernel over CUDA IPC symmetric memory": what cpm_allreduce_lightvol last used */
CPM_API const char* cpm_comm_transport(const cpm_comm* comm);
/* sum_out = sum over ranks of `local` (n_floats each; sum_out may equal local).  Option B of SURVEY 8e: every rank
 * splats its photon shard into its own light volume, the frame's result is the sum.  When the GPUs can map each other's
 * memory (CUDA IPC; CPM_COMM_TRANSPORT=nccl switches it off) the sum is formed by the library's own kernel over NVLink
 * peer memory -- snapshot into a symmetric staging buffer, flag barrier, cpm_allreduce_peer_f32 (rank r reduces slice r
 * with peer loads in rank order: every rank gets the same bits), flag barrier -- otherwise by ncclAllReduce. */
CPM_API int cpm_allreduce_lightvol(cpm_comm* comm, const float* local, float* sum_out, size_t n_floats);
/* The same in two halves for pipelining (cpm_allreduce_lightvol = _begin + _end): _begin snapshots `local` -- record an
 * event after it; once that event has completed the next frame may overwrite `local` --, _end forms the sum in sum_out.
 * Typical use: a communicator on a side-stream context (cpm_comm_split), so that the exchange runs next to the following
 * frame's detector / re-trace and only that frame's splat waits for the snapshot event. */
CPM_API int cpm_allreduce_lightvol_begin(cpm_comm* comm, const float* local, float* sum_out, size_t n_floats);
CPM_API int cpm_allreduce_lightvol_end(cpm_comm* comm, float* sum_out, size_t n_floats);
/* a second communicator over the same ranks, bound to another context (another stream) of this process; collective */
CPM_API int cpm_comm_split(cpm_comm* comm, cpm_ctx* other_ctx, cpm_comm** out);
/* all_out = every rank's n_floats_per_rank photon-record floats in rank order (option A of SURVEY 8e: the replicated
 * photon map for gathering) */
CPM_API int cpm_allgather_photons(cpm_comm* comm, const float* local, size_t n_floats_per_rank, float* all_out);
/* Sharded ingest of a time step ("broadcast once per time step over NVLink"): on entry rank r holds slab r
 * (bytes [r * slab_bytes, (r + 1) * slab_bytes) of `volume`, e.g. uploaded from its host), on return every rank holds
 * all world * slab_bytes bytes.  In place. */
CPM_API int cpm_allgather_volume(cpm_comm* comm, void* volume, size_t slab_bytes);
/* The same with the upload: this rank copies ITS slab of the host buffer `src_host` (the whole time step, total_bytes) to
 * the same offset of `volume` and the slabs are all-gathered, so a step costs total_bytes / world of PCIe traffic per
 * GPU instead of total_bytes.  on_transfer_stream = 0: on the context stream.  on_transfer_stream = 1: both the copy and
 * the all-gather run on the context's transfer stream (its own NCCL communicator where ncclCommSplit exists) after the
 * work already submitted to the context stream, and *done is an event to hand to cpm_ctx_wait_event before the volume is
 * used -- the pendant of cpm_mem_prefetch_h2d.  total_bytes must split into `world` equal 16-byte aligned slabs. */
CPM_API int cpm_comm_upload_volume_sharded(cpm_comm* comm, void* volume, const void* src_host, size_t total_bytes,
                                           int on_transfer_stream, cpm_event** done);
CPM_API int cpm_comm_barrier(cpm_comm* comm);
/* Host-value all-gather: all_out[r * count + i] = values[i] of rank r (count <= 64).  SYNCHRONOUS (returns the values). */
CPM_API int cpm_comm_allgather_u64(cpm_comm* comm, const unsigned long long* values, int count, unsigned long long* all_out);
/* Global (cross-shard) selection, SURVEY.md 8e: "all-gather the (key, idx) of candidates and select globally".  Every rank
 * holds its own keys sorted ascending (device memory, what cpm_radix_sort_u32 leaves); the global order is (key, rank,
 * local index), i.e. one stable sort over the concatenated shards.  *local_count = how many of THIS rank's sorted keys
 * are among the first `position` elements of that order: the re-trace budget max% * N_total applied to the whole photon
 * set instead of per shard (ppm/processor/progressivephotontracercl.cpp:419-431 on one GPU).  Four rounds of a 256-ary
 * search over the key bits, each one all-gather of 257 counts: no keys travel.  COLLECTIVE and SYNCHRONOUS. */
CPM_API int cpm_comm_select_global(cpm_comm* comm, const uint32_t* sorted_keys, size_t n_local, unsigned long long position,
                                   unsigned long long* local_count);

/* ---- (5)(6)(7) photon map for gathering: cell keys, cell-sorted records, ray-march gather ------ */
/* Not launched anywhere in the reference; the estimator is the reference's
 * (Epanechnikov kernel, ppm/cl/densityestimationkernel.cl:56-60; power * 1/(4 pi) * relativeIrradianceScale,
 * ppm/cl/photonstolightvolume.cl:160-165), the per-point formulation is the one of its disabled
 * photonsToLightVolumeKernel (ppm/cl/photonstolightvolume.cl:81-134).  Parity: the oracle's restatement.
 *
 * Photon map build: cpm_photon_cell_keys -> cpm_radix_sort_u32(keys, ids) -> cpm_build_cell_ranges ->
 * cpm_reorder_photons; then cpm_gather_raymarch / cpm_gather_points. */

/* keys[i] = cell of photon record i in a grid of grid_dims cells over texture space [0,1]^3
 * (x + gx*(y + gy*z)); empty records (FLT_MAX sentinel) get n_cells.  ids (optional) = iota. */
CPM_API int cpm_photon_cell_keys(cpm_ctx* ctx, const float* photons, size_t n_records, const int grid_dims[3],
                                 uint32_t* keys, uint32_t* ids);
/* out[j] = photons[ids[j]] (32 B records), j < n */
CPM_API int cpm_reorder_photons(cpm_ctx* ctx, const float* photons, const uint32_t* ids, size_t n, float* out);
/* the same records, planar: out[4*j..] = floats 0-3 of photons[ids[j]] (position, power.r), out[4*(n+j)..] = floats
 * 4-7 (power.g, power.b, theta, phi).  The gather kernels test candidates against the first half only; pass
 * cpm_gather_params::planar_records = n. */
CPM_API int cpm_reorder_photons_planar(cpm_ctx* ctx, const float* photons, const uint32_t* ids, size_t n, float* out);

typedef struct cpm_gather_params {
    int32_t width, height;
    float cam_origin[3];  /* texture space */
    float cam_dir00[3];   /* ray through pixel (x, y) = normalize(dir00 + (x+0.5) du + (y+0.5) dv) */
    float cam_du[3];
    float cam_dv[3];
    float aabb_min[3];    /* clip box, texture space */
    float aabb_max[3];
    float step;           /* ray-march step, texture units */
    float radius;         /* gather radius, texture units */
    float scale;          /* relative irradiance scale, as cpm_splat_photons */
    float sigma_scale;    /* extinction per unit opacity and unit length: 150 matches the tracer
                             (invTauMaxSampleBaseInterval = 1/(tauMax*150), ppm/cl/transmittance.cl:40,130) */
    int32_t grid_dims[3];
    /* optional (NULL = off): the per-cell opacity bound of cpm_opacity_bound for the SAME volume and transfer
     * function; cells whose bound is exactly zero are stepped over without fetching voxels (identical image) */
    const float* opacity_bound;
    int32_t bound_cell_log2;
    /* Image tiling across GPUs (SURVEY 8e option A: every GPU gathers its own image tiles against the replicated
     * photon map).  The image buffer of a call holds `height` rows made of strips of 4 rows (the kernels' tile
     * height); local strip s is strip `strip_first + s * strip_stride` of the camera's image, i.e. local row py is
     * camera row 4 * (strip_first + (py >> 2) * strip_stride) + (py & 3).  strip_stride <= 1 with strip_first = 0
     * is the whole image.  Interleaved strips balance the very different ray costs between ranks; pixels are
     * bit-identical to the same pixels of a whole-image call. */
    int32_t strip_first;
    int32_t strip_stride;
    /* layout of sorted_photons: 0 = 32-byte records as cpm_reorder_photons writes them; n > 0 = the n records as
     * cpm_reorder_photons_planar writes them (n first halves, then n second halves) */
    int32_t planar_records;
} cpm_gather_params;

/* image[y*width + x] = (radiance rgb, 1 - transmittance): front-to-back emission-absorption ray march,
 * the in-scattered radiance at every sample = TF colour * gathered irradiance (isotropic phase). */
CPM_API int cpm_gather_raymarch(cpm_ctx* ctx, const cpm_volume* vol, const float* tf_rgba, int tf_width,
                                const cpm_gather_params* params, const float* sorted_photons,
                                const uint32_t* cell_start, const uint32_t* cell_end, float* image /* float4[w*h] */);
/* irradiance[3*i..] at points[3*i..] (only radius, scale and grid_dims of params are used): at a light-volume
 * voxel centre this equals what cpm_splat_photons accumulates there. */
CPM_API int cpm_gather_points(cpm_ctx* ctx, const cpm_gather_params* params, const float* sorted_photons,
                              const uint32_t* cell_start, const uint32_t* cell_end, const float* points,
                              int n_points, float* irradiance);

/* Final image from the light volume: image[y*width + x] = (radiance rgb, 1 - transmittance) of a front-to-back
 * emission-absorption ray march of volume x transfer function x light volume -- what the workspace network
 * does with Inviwo's LightingRaycaster after PhotonToLightVolumeProcessorCL (ws:1178-1271; an Inviwo core OpenGL
 * processor outside the reference tree: parity unpinned, checker = the oracle's restatement).  Camera, clip box,
 * step, sigma_scale and the optional opacity bound are taken from params exactly as cpm_gather_raymarch takes
 * them (radius / scale / grid_dims are unused); light_volume is float[lv_dims] (channels = 1: scalar irradiance
 * applied to the three colour channels) or float4[lv_dims] (channels = 4), sampled trilinearly. */
CPM_API int cpm_raycast_light_volume(cpm_ctx* ctx, const cpm_volume* vol, const float* tf_rgba, int tf_width,
                                     const cpm_gather_params* params, const float* light_volume, const int lv_dims[3],
                                     int channels, float* image /* float4[w*h] */);

/* ---- device memory (cl::Buffer, enqueueWrite/Read/Copy/FillBuffer) -------------------- */
/* Lets host code above this ABI (host/: the Inviwo processor mirror) stay free of CUDA headers.
 * Copies and fills are asynchronous on the context stream; use cpm_ctx_sync before reading a
 * d2h destination.  cpm_host_alloc returns page-locked memory (true async copies). */
CPM_API int cpm_mem_alloc(cpm_ctx* ctx, size_t bytes, void** out);
CPM_API int cpm_mem_free(cpm_ctx* ctx, void* ptr);
CPM_API int cpm_host_alloc(cpm_ctx* ctx, size_t bytes, void** out);
CPM_API int cpm_host_free(cpm_ctx* ctx, void* ptr);
CPM_API int cpm_mem_copy_h2d(cpm_ctx* ctx, void* dst, const void* src_host, size_t bytes);
CPM_API int cpm_mem_copy_d2h(cpm_ctx* ctx, void* dst_host, const void* src, size_t bytes);
/* Upload on the context's TRANSFER stream (created on first use), so that the copy of the next time step
 * overlaps the kernels of the current one.  The transfer first waits for everything enqueued on the context
 * stream so far (earlier readers of dst); *done receives a new event (cpm_event_destroy) that
 * cpm_ctx_wait_event makes the context stream wait for before dst is consumed.  src_host must be pinned. */
CPM_API int cpm_mem_prefetch_h2d(cpm_ctx* ctx, void* dst, const void* src_host, size_t bytes, cpm_event** done);
CPM_API int cpm_ctx_wait_event(cpm_ctx* ctx, cpm_event* ev);
/* The mirror image: device -> pinned host on the context's READ-BACK stream (created on first use), after the work
 * already submitted to the context stream; the context stream itself does not wait, so the next frame's kernels run
 * while the result travels.  *done: cpm_event_sync before the host reads dst_host, cpm_ctx_wait_event before `src` is
 * overwritten. */
CPM_API int cpm_mem_readback_d2h(cpm_ctx* ctx, void* dst_host, const void* src, size_t bytes, cpm_event** done);
CPM_API int cpm_event_sync(cpm_ctx* ctx, cpm_event* ev); /* cudaEventSynchronize */
/* the same for an event owned by the caller's runtime (a cudaEvent_t, e.g. torch.cuda.Event.cuda_event) */
CPM_API int cpm_ctx_wait_cuda_event(cpm_ctx* ctx, void* cuda_event);
/* overlapping ranges are allowed when dst < src (the index-list slide-down of
 * ppm/processor/progressivephotontracercl.cpp:389-419) */
CPM_API int cpm_mem_copy_d2d(cpm_ctx* ctx, void* dst, const void* src, size_t bytes);
/* enqueueFillBuffer<unsigned int> (resetPhotonImportance, :607-611) */
CPM_API int cpm_mem_fill_u32(cpm_ctx* ctx, void* dst, uint32_t value, size_t count);
/* dst[indices[i]] = value for i < n: resets the importance keys of the photons just re-traced.
 * (The reference resets a slice of its in-place sorted key array, :529; here keys stay in photon
 * order and are addressed through the sorted id list.) */
CPM_API int cpm_mem_scatter_fill_u32(cpm_ctx* ctx, void* dst, const uint32_t* indices, size_t n,
                                     uint32_t value);

/* mixKernel (ugc/cl/buffermixer.cl:37-48), host side BufferMixerCL::mix (ugc/buffermixercl.cpp:47-85); also the
 * voxel arithmetic of VolumeSequencePlayer's volume_mix.frag (ugc/processors/volumesequenceplayer.cpp:94-143):
 * out[i] = x[i] + (y[i] - x[i]) * a over n scalars of format CPM_FMT_* (float2/3/4 buffers: n = elements *
 * components); u8 / u16 are mixed in float and converted back with round-toward-zero. */
CPM_API int cpm_mix(cpm_ctx* ctx, const void* x, const void* y, float a, size_t n, int format, void* out);
/* the same through normalised textures, as VolumeSequencePlayer's fragment shader sees integer volumes
 * (ugc/glsl/volume_mix.frag:44-52): out = round(clamp(mix(x / max, y / max, a), 0, 1) * max); identical to cpm_mix for f32.
 * OpenGL arithmetic: parity unpinned. */
CPM_API int cpm_mix_unorm(cpm_ctx* ctx, const void* x, const void* y, float a, size_t n, int format, void* out);

/* ---- self test ----------------------------------------------------------------------- */
/* Evaluates one function of include/cpm_detmath.h on the device: fn 0 log, 1 sin, 2 cos,
 * 3 acos, 4 atan2(x, y), 5 v/255, 6 v/65535 (x holds the integer value as float), 7 pow(x, y), 8 cbrt,
 * 9 exp on [-87, 87].  Lets the
 * tests prove that the host and sm_100a compilations of the math layer agree bit for bit. */
CPM_API int cpm_selftest_math(cpm_ctx* ctx, int fn, const float* x, const float* y, float* out,
                              size_t n);

#ifdef __cplusplus
}
#endif
#endif /* CPM_B200_H */
