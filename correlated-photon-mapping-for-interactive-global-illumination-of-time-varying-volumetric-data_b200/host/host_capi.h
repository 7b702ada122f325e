/* host_capi.h -- C entry points that drive the drop-in processor network headless (what Inviwo's
 * ProcessorNetworkEvaluator does when the reference workspace is open).  Used by bench.py (the
 * `e2e` leg: host buffers in, host buffers out) and by the host-layer tests.  The network is the
 * one of workspaces/CorrelatedPhotonMappingSingleVolume.inv (ws:1178-1271):
 *   VolumeSource -> VolumeMinMaxCL -> MinMaxUniformGrid3DImportance -> (importance grid) --+
 *   UniformSampleGenerator2D -> DirectionalLightSamplerCL (x n_lights) -> ProgressivePhotonTracerCL
 *                                                                      -> PhotonToLightVolumeProcessorCL
 */
#ifndef CPM_HOST_CAPI_H
#define CPM_HOST_CAPI_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#define CPMH_API __attribute__((visibility("default")))

typedef struct cpmh_network cpmh_network;

typedef struct cpmh_config {
    int32_t device;
    int32_t dims[3];
    int32_t format;              /* CPM_FMT_* */
    int32_t samples_per_side;    /* UniformSampleGenerator2D nSamples = (s, s) per light */
    int32_t n_lights;
    float light_directions[8][3];
    float light_intensity[8][3];
    int32_t max_scattering_events;
    int32_t light_volume_option; /* volumeSizeOption: 0 radius, 1, 2, 4 */
    int32_t light_volume_channels; /* 1 or 4 */
    int32_t with_importance_grid;  /* connect VolumeMinMax + importance grid => correlated re-tracing */
    int32_t volume_layout;       /* CPM_VOLUME_* used by the tracer */
    float photon_radius_voxels;  /* `radius` property */
    float max_incremental_percent; /* `maxIncrementalPhotonsToUpdate` */
    int32_t clip[6];             /* clipX.min,max, clipY.., clipZ..; all zero = no clipping */
    int32_t reference_full_splat_bound;
    float incremental_threshold_percent; /* `incrementalRecomputationThreshold`; 0 = the reference's default 50 */
    int32_t opacity_bound_cell_log2; /* tracer's per-cell opacity bound (cpm_opacity_bound): 0 = default (8^3-voxel
                                        cells), n > 0 = 2^n voxels per axis, < 0 = off (fetch for every test, as
                                        the reference does); photons are identical in every setting */
} cpmh_config;

/* One context per process.  Optional: call before the first network is created to choose the CUDA
 * stream (a cudaStream_t, NULL = own stream) and this process's photon shard: photon i of this process
 * is photon photon_shard_offset + i of the global photon set (MWC64X stream and host base offset). */
CPMH_API int cpmh_runtime_init(int device, void* stream, uint64_t photon_shard_offset);
CPMH_API int cpmh_runtime_set_photon_shard_offset(uint64_t photon_shard_offset);
/* Multi-GPU (one process per GPU).  cpmh_runtime_ctx() is the process's cpm_ctx*: create the communicator on it with
 * cpm_comm_unique_id / cpm_comm_init (include/cpm_b200.h) and hand it over with cpmh_runtime_set_comm.  sharded_ingest != 0:
 * every volume that arrives from HOST memory (cpmh_network_set_volume_host with pinned memory, _stream_timestep_host,
 * _prefetch_timestep_host; every rank passes a buffer holding the whole time step) is uploaded as this rank's slab only
 * and completed over NVLink -- 1 / world of the PCIe traffic per GPU.  Pass NULL to detach before destroying the
 * communicator. */
CPMH_API void* cpmh_runtime_ctx(void);
CPMH_API int cpmh_runtime_set_comm(void* cpm_comm_handle, int sharded_ingest);
/* on != 0 (and a communicator set): `maxIncrementalPhotonsToUpdate` is a budget over the photons of ALL shards, selected in one
 * global importance order (cpm_comm_select_global); every rank must then evaluate its network the same number of times.
 * Default off: the budget applies per shard (SURVEY.md 8e offers both). */
CPMH_API int cpmh_runtime_set_global_budget(int on);
/* frame result across GPUs: sum over ranks of the networks' light volumes (cpm_allreduce_lightvol) into a device buffer
 * owned by the network (*sum_device, valid until the next call); out_host != NULL: also read back into it (n_floats),
 * synchronously -- typically on rank 0 only */
CPMH_API int cpmh_network_sum_light_volume(cpmh_network* net, float* out_host, size_t n_floats, void** sum_device);
/* The frame result on its way to the host WITHOUT stalling the next frame: the light volume (sum_over_ranks != 0 and a
 * communicator set: its sum over ranks; out_host == NULL: the sum is formed, nothing is read back -- the ranks that do not
 * display) is snapshotted into one of two staging buffers and copied to out_host (pinned) on a read-back stream; the call
 * returns at once.  cpmh_network_wait_readback blocks until the most recent read-back has landed.  Pattern:
 *   evaluate(k); read_light_volume_async(out[k & 1]); ... evaluate(k+1) ...; wait_readback(); use out[k & 1] */
CPMH_API int cpmh_network_read_light_volume_async(cpmh_network* net, float* out_host, size_t n_floats, int sum_over_ranks);
CPMH_API int cpmh_network_wait_readback(cpmh_network* net);
CPMH_API int cpmh_network_create(const cpmh_config* cfg, cpmh_network** out);
CPMH_API void cpmh_network_destroy(cpmh_network* net);
CPMH_API const char* cpmh_last_error(void);
/* transfer function of the tracer AND of the importance processor: n points of (pos, r, g, b, a) */
CPMH_API int cpmh_network_set_transfer_function(cpmh_network* net, const float* points, int n);
/* new voxel data from a HOST buffer (kept by reference until the next evaluate: use pinned memory for
 * asynchronous upload).  Marks the volume port changed, as VolumeSource / a sequence player would. */
CPMH_API int cpmh_network_set_volume_host(cpmh_network* net, const void* voxels_host);
/* a time series: T host buffers; min-max grids and inter-step difference grids are computed on the
 * device for all steps (VolumeMinMaxCL on the sequence + DynamicVolumeDifferenceAnalysis) */
CPMH_API int cpmh_network_set_sequence_host(cpmh_network* net, const void* const* voxels_host, int n_steps);
/* Volume::dataMap_.dataRange of the network's volume(s), e.g. (0, 4095) for 12-bit data in a 16-bit volume ("scaling for
 * 12-bit data", ugc/processors/volumeminmaxclprocessor.cpp:134-136); marks the volume changed */
CPMH_API int cpmh_network_set_data_range(cpmh_network* net, double lo, double hi);
CPMH_API int cpmh_network_set_timestep(cpmh_network* net, int t);
/* layout the tracer samples from now on (CPM_VOLUME_TEXTURE / CPM_VOLUME_LINEAR; same photons either way).  A resident
 * series is best kept as CUDA arrays (built once, 12 % faster walks); a series that streams from the host is best
 * sampled where it lands -- the linear buffer -- which saves the 0.54 ms linear -> array copy of a 512^3 f32 step. */
CPMH_API int cpmh_network_set_volume_layout(cpmh_network* net, int layout);
/* Streaming variant for data that does not stay resident: upload the next time step from a HOST buffer
 * (pinned memory: asynchronous), compute its min-max grid and the per-brick difference to the previously
 * streamed step on the device, and mark volume / grids changed.  The buffer must stay valid until the
 * next evaluate has been synchronised (cpmh_network_read_* or cpmh_network_sync). */
CPMH_API int cpmh_network_stream_timestep_host(cpmh_network* net, const void* voxels_host);
/* Optional: announce the buffer of the NEXT cpmh_network_stream_timestep_host call.  Its upload starts at once
 * on a transfer stream and overlaps the evaluation of the current step (three device volumes rotate: previous,
 * current, incoming).  The buffer must be pinned and stay untouched until that step has been evaluated. */
CPMH_API int cpmh_network_prefetch_timestep_host(cpmh_network* net, const void* voxels_host);
CPMH_API int cpmh_network_sync(cpmh_network* net);
/* device pointer of the light volume (float[dims] or float4[dims]) for zero-copy consumers, e.g. an
 * NCCL all-reduce across the GPUs that each splatted their photon shard */
CPMH_API int cpmh_network_light_volume_device(cpmh_network* net, void** ptr, size_t* n_floats);
/* device pointer of the photon records (float8 x N x maxScatteringEvents) for zero-copy consumers (the
 * photon-map gather) */
/* The next evaluation's first write to the light volume waits for `cuda_event` (a cudaEvent_t recorded on another
 * stream by a consumer that is still reading the volume -- the multi-GPU exchange's snapshot); everything before that
 * write (detector, selection, re-trace) is not held up.  One-shot. */
CPMH_API int cpmh_network_wait_before_light_volume_write(cpmh_network* net, void* cuda_event);
CPMH_API int cpmh_network_photons_device(cpmh_network* net, void** ptr, size_t* n_floats);
/* count delta-tracking collision tests of every trace from now on (device counter); read = sync */
CPMH_API int cpmh_network_count_collision_tests(cpmh_network* net, int on);
CPMH_API unsigned long long cpmh_network_read_collision_tests(cpmh_network* net, int reset);
/* out[0] = collision tests, out[1] = the ones that fetched voxels (the rest were decided by the opacity bound) */
CPMH_API int cpmh_network_read_collision_stats(cpmh_network* net, unsigned long long out[2], int reset);
/* evaluate every invalid processor in network order; returns the number of processors that ran */
CPMH_API int cpmh_network_evaluate(cpmh_network* net);
/* progressive work left (budgeted re-trace batches): call evaluate again while > 0 */
CPMH_API int cpmh_network_remaining_photons(cpmh_network* net);
/* the tracer's 100 ms progressive-refinement timer tick (ProgressivePhotonTracerCL::onTimerEvent,
 * ppm/processor/progressivephotontracercl.cpp:186-190): marks the tracer invalid for reason Progressive */
CPMH_API int cpmh_network_timer_event(cpmh_network* net);
CPMH_API int cpmh_network_n_photons(cpmh_network* net);
CPMH_API int cpmh_network_n_recomputed(cpmh_network* net);
CPMH_API int cpmh_network_light_volume_dims(cpmh_network* net, int dims[3]);
/* device -> host reads (synchronous) */
CPMH_API int cpmh_network_read_light_volume(cpmh_network* net, float* out_host, size_t n_floats);
CPMH_API int cpmh_network_read_photons(cpmh_network* net, float* out_host, size_t n_floats);
CPMH_API int cpmh_network_read_importance_keys(cpmh_network* net, uint32_t* out_host, size_t n);
/* the ids the tracer re-traced in its last evaluation (RecomputedPhotonIndices, ppm/photondata.h:58-63): out_host holds
 * cpmh_network_n_recomputed() entries; returns the number written (0 when the last evaluation traced everything) */
CPMH_API int cpmh_network_read_recomputed_indices(cpmh_network* net, uint32_t* out_host, size_t n);
/* the importance grid the tracer's detector last saw (float per brick of `region` voxels) */
CPMH_API int cpmh_network_read_importance_grid(cpmh_network* net, float* out_host, size_t n);
/* the transfer-function point list the importance classifier last used (host side: the points with end points, or the
 * |new - old| difference list after a transfer-function change): positions[capacity], colors[4 * capacity]; returns the
 * number of points (tfPointImportanceSize_) */
CPMH_API int cpmh_network_importance_tf_points(cpmh_network* net, float* positions, float* colors, int capacity);
/* properties of the network's processors by class id / occurrence / identifier (bool, int, float and option values) */
CPMH_API int cpmh_network_set_property(cpmh_network* net, const char* class_id, int k, const char* property, double value);
/* kernel arguments of light sampler `light` (lcl/directionallightsamplercl.cpp:66-73), for parity checks:
 * out = direction[3], plane point before the fit[3], fitted origin[3], u[3], v[3], radiance[3], area */
CPMH_API int cpmh_network_light_setup(cpmh_network* net, int light, float out[19]);
/* light samples (float8 per sample) and (tStart, tEnd) intersections of light sampler `light`, device -> host */
CPMH_API int cpmh_network_read_light_samples(cpmh_network* net, int light, float* samples_out, float* isect_out, size_t n);
CPMH_API const char* cpmh_network_last_splat_path(cpmh_network* net);
/* stage timing of the tracer's last process(): "detector","count+iota","sort","indexsort","trace" */
CPMH_API int cpmh_network_set_profile(cpmh_network* net, int on);
CPMH_API float cpmh_network_stage_ms(cpmh_network* net, const char* stage);
/* accumulated device time per stage (CUDA events on the context stream) since the last reset:
 * "seed","emission","h2d","texcopy","minmax","voldiff","classify","detector","count+iota","sort",
 * "indexsort","trace","splat","copyprev" */
CPMH_API void cpmh_profile_enable(int on);
/* time only this stage while profiling is on (NULL or "": every stage): two event records per frame instead of two per stage */
CPMH_API void cpmh_profile_only(const char* stage);
CPMH_API void cpmh_profile_reset(void);
CPMH_API double cpmh_profile_total_ms(const char* stage);
CPMH_API int cpmh_profile_count(const char* stage);
CPMH_API const char* cpmh_profile_stages(void);
CPMH_API uint64_t cpmh_network_launch_count(cpmh_network* net, int reset);
CPMH_API void cpmh_transfer_bytes(uint64_t* h2d, uint64_t* d2h, int reset);
CPMH_API void* cpmh_network_ctx(cpmh_network* net);
/* the host layer's CPU light-plane fit (lcl/orientedboundingbox2d.cpp:80-100 + convexhull2d.cpp +
 * pointplaneprojection.cpp), exposed for parity tests: out = origin[3], u[3], v[3].  No device needed. */
/* ".u3d" uniform-grid sequences (ugc/uniformgrid3dreader.cpp:59-183, ugc/uniformgrid3dwriter.cpp:47-102): text
 * header + raw file.  format: 0 = FLOAT32 (importance / difference grids), 1 = Vec2UINT16 (min-max grids).
 * Host-only (no device needed).  dims4 = (x, y, z, number of grids); matrices are column-major 4x4. */
CPMH_API int cpmh_u3d_write(const char* path, int format, const int dims4[4], const int cell[3], const float model[16],
                            const float world[16], const void* data);
CPMH_API int cpmh_u3d_read_info(const char* path, int* format, int dims4[4], int cell[3], float model[16], float world[16]);
CPMH_API int cpmh_u3d_read_data(const char* path, void* out, size_t bytes);
/* UniformGrid3DExport (ugc/processors/uniformgrid3dexport.cpp) for the grids of the resident sequence:
 * which = 0 the per-step min-max grids, 1 the step-to-step difference grids */
CPMH_API int cpmh_network_export_sequence_grids(cpmh_network* net, int which, const char* path);

/* ".inv" workspaces (host only, no device needed for the first two).  The reference ships
 * workspaces/CorrelatedPhotonMappingSingleVolume.inv; the drop-in processors keep the reference's class and property
 * identifiers, so the file configures them headless.
 *   cpmh_workspace_describe      "processor|<class id>|<name>|<stored property paths>" and
 *                                "connection|<name>.<outport>|<name>.<inport>" lines
 *   cpmh_config_from_workspace   fills what the workspace determines -- samples_per_side (UniformSampleGenerator2DCL
 *                                nSamples), n_lights (DirectionalLightSamplerCL processors), light directions and
 *                                intensities (k-th org.inviwo.Directionallightsource: direction = -normalize(position
 *                                in world space), converted to texture space with the volume's world extents `basis`,
 *                                NULL = unit cube; intensity = lightPower * lightDiffuse -- Inviwo base-module
 *                                behaviour, outside the reference tree: parity unpinned), max_scattering_events,
 *                                photon_radius_voxels, max_incremental_percent, clip, light_volume_option / channels,
 *                                incremental_threshold_percent, with_importance_grid (is the tracer's
 *                                recomputationImportance port connected) -- and leaves device, dims, format, layout.
 *   cpmh_network_load_workspace  applies every stored property (transfer functions, weights, options, clip ranges,
 *                                material, camera ...) to the network's processors of the same class id, k-th
 *                                occurrence to k-th instance; returns the number of properties applied (>= 0). */
CPMH_API const char* cpmh_workspace_describe(const char* path);
CPMH_API int cpmh_config_from_workspace(const char* path, const float basis[3], cpmh_config* cfg);
CPMH_API int cpmh_network_load_workspace(cpmh_network* net, const char* path);
/* value of a scalar / bool / option property of the network's processors, for checks: processor = class id
 * ("org.inviwo.ProgressivePhotonTracerCL"), occurrence k, property identifier; options report their selected value */
/* UniformSampleGenerator2DCL nSamples = (n, n): e.g. a smaller photon count than the workspace stores */
CPMH_API int cpmh_network_set_samples_per_side(cpmh_network* net, int n);
CPMH_API int cpmh_network_get_property(cpmh_network* net, const char* class_id, int k, const char* property, double* out);

/* RandomNumberGeneratorCL (ny == 0: nSamples = nx) or RandomNumberGenerator2DCL (nSamples = (nx, ny)) evaluated
 * `evaluations` times with the given seed; the numbers of the last evaluation are read back (nx * max(ny, 1) floats). */
CPMH_API int cpmh_random_numbers(int nx, int ny, int seed, int evaluations, float* out_host);
/* host only: the |new - old| transfer-function point list of MinMaxUniformGrid3DImportanceCLProcessor
 * (isc/processors/minmaxuniformgrid3dimportanceclprocessor.cpp:364-501) for two point sets of (pos, r, g, b, a); returns the
 * number of points written to positions[capacity] / colors[4 * capacity] */
CPMH_API int cpmh_tf_difference_points(const float* cur, int n_cur, const float* prev, int n_prev, float epsilon, int associated,
                                       float* positions, float* colors, int capacity);
/* host only, for parity tests against the reference's ppm/photondata.cpp: PhotonData after setSize / setRadius(radius_rel,
 * scene_radius) / `iterations` x advanceToNextIteration(alpha): out = {getRadius, getRadiusRelativeToSceneSize,
 * getRelativeIrradianceScale, iteration}; and Photon::setDirection / getDirection */
CPMH_API int cpmh_photondata_progress(size_t n_photons, int max_interactions, double radius_rel, double scene_radius, int iterations,
                                      double alpha, double out[4]);
CPMH_API int cpmh_photon_encode_direction(const float dir[3], float out[2]);
CPMH_API int cpmh_photon_decode_direction(const float enc[2], float out[3]);
/* state of the network's PhotonData after the last evaluation: out = {iteration, getRadius, getSceneRadius,
 * getRadiusRelativeToSceneSize, getRelativeIrradianceScale} */
CPMH_API int cpmh_network_photon_state(cpmh_network* net, double out[5]);
CPMH_API int cpmh_fit_light_plane(const float* points, int n_points, const float plane_point[3],
                                  const float plane_normal[3], float out[9]);
/* the 2-D convex hull alone (lcl/convexhull2d.cpp:38-130): hull_out holds up to 2 * n + 2 points; returns the hull size */
CPMH_API int cpmh_convex_hull2d(const float* points_xy, int n_points, float* hull_out);
/* The sequence players, headless.  cpmh_player_clock: host only -- `ticks` timer events of a player over a
 * sequence of n_elements (time and selectedSequenceIndex after each tick).  cpmh_player_grids_f32: UniformGrid3DPlayerProcessor
 * over n_grids float grids of n_cells (host memory, back to back); for every entry of `times`: set time, evaluate, read the
 * interpolated grid back (out_host: n_times x n_cells), report the sequence index and which of the two ping-pong output
 * grids carried it (0 / 1; -1: an input grid passed through).  cpmh_player_volumes: VolumeSequencePlayer likewise. */
CPMH_API int cpmh_player_clock(int n_elements, float time_per_element, int frame_rate, int ticks, float* times_out, int* index_out);
CPMH_API int cpmh_player_grids_f32(const float* grids_host, int n_grids, size_t n_cells, float time_per_element, const float* times,
                                   int n_times, float* out_host, int* out_index, int* out_buffer);
CPMH_API int cpmh_player_volumes(const void* volumes_host, int n_volumes, const int dims[3], int format, float time_per_volume,
                                 const float* times, int n_times, void* out_host, int* out_index);
/* introspection for drop-in checks: "classId|port,port,...|prop,prop,..." per processor, newline separated */
CPMH_API const char* cpmh_describe_processors(void);
#ifdef __cplusplus
}
#endif
#endif
