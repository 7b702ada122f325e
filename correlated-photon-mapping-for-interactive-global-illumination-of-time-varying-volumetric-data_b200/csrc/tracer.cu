// tracer.cu -- delta-tracking photon tracer (north-star subsystem 3 and the re-trace half of 4).
//
// Replaces photonTracerKernel (ppm/cl/photontracer.cl:69-216) with woodcockTracking
// (ppm/cl/transmittance.cl:126-144) and nextInteraction (photontracer.cl:50-58).
//
// B200 design
//  * Each photon owns MWC64X stream `photon_offset + i`, so the result of a photon does not
//    depend on which lane, warp or GPU traces it.  That freedom is used for scheduling.
//  * The woodcock loop needs only the ALPHA channel of the transfer function (colour is never
//    read on this path: photontracer.cl:171-176 use color.w / scattering.w only).  The alpha
//    column (tf_width floats) is staged once per CTA in shared memory; every collision test
//    then costs two LDS instead of two RGBA texture fetches.
//  * Volume taps: exact fp32 trilinear (OpenCL 1.2 section 8.2 arithmetic) over unfiltered
//    texels.  TEXTURE layout: 2 x tld4 on a 2-D layered array (the texture unit does the
//    clamping and fetches a 2x2 footprint per instruction, block-linear layout keeps the
//    footprint in one sector).  LINEAR layout: 8 x LDG from the caller's buffer.
//    Hardware-filtered taps are not used: their 1.8 fixed-point weights flip accept/reject
//    decisions, which breaks the replay property the correlated re-trace depends on.
//  * Photon records are 32 B: written as two 16 B stores (STG.128).
#include <algorithm>

#include "sampling.cuh"

namespace {

struct TraceArgs {
    cpm_trace_params p;
    VolumeView vol;
    const float4* tf;
    int tf_width;
    const float4* light_samples;  // float8 as 2 x float4
    const float2* isect;
    const uint32_t* recompute;
    int n_work;  // n_light_samples or n_recompute
    float4* photons;
    uint2* rng;
    unsigned long long* tests;
    BoundGrid bound;  // per-cell opacity bound (cpm_opacity_bound); bound.g == null: off
    cudaTextureObject_t btex;  // the same grid as a point-sampled 3-D texture (cpm_bound_tex), or 0
    int scan;  // cheap tests a lane scans before the warp reconverges for the candidate fetches
    int zero;  // 0, unknown to the compiler (ScanRegs)
};

__device__ __forceinline__ void store_photon(float4* photons, size_t id, float x, float y, float z, float pr, float pg,
                                             float pb, float th, float ph) {
    photons[2 * id] = make_float4(x, y, z, pr);
    photons[2 * id + 1] = make_float4(pg, pb, th, ph);
}

#define CPM_FLT_MAX 3.402823466e+38f
#ifndef CPM_SCAN
#define CPM_SCAN 8  // cheap tests a lane scans before the warp reconverges for the candidate fetches
#endif
#ifndef CPM_TRACE_THREADS
#define CPM_TRACE_THREADS 128  // threads per CTA of trace_kernel
#endif
#ifndef CPM_TRACE_MIN_CTAS
#define CPM_TRACE_MIN_CTAS (9 * 128 / CPM_TRACE_THREADS)  // __launch_bounds__: register cap 56 (tools/build_variant.sh sweeps it)
#endif
#ifndef CPM_TRACE_WAVEFRONT_DEFAULT
#define CPM_TRACE_WAVEFRONT_DEFAULT 0  // 1: bounded traces take the wavefront form unless CPM_TRACE_WAVEFRONT=0
#endif
#ifndef CPM_TRACE_OPAQUE_RD
#define CPM_TRACE_OPAQUE_RD 0  // 1: also hide d * fc from the compiler's rematerialisation (measured slower: 0.596 vs 0.561 ms)
#endif

// native_log(u) for u = (float)k * 2^-32, k the raw 32-bit draw (what random_01 returns): cpm_native_logf_tab
// (include/cpm_detmath.h) with its table in shared memory, without the rare-argument exits such values never take,
// and with the 2^-32 folded into the bit constants (the product is exact, so bits((float)k * 2^-32) =
// bits((float)k) - 0x10000000 for k != 0); k == 0 becomes a select.  Same bits as the header function.
__device__ __forceinline__ float log_unit(uint32_t k, const float2* __restrict__ s_nlog, float c02 = 0.2f) {
    const uint32_t fb = __float_as_uint(__uint2float_rn(k));
    const uint32_t jx = fb + (0x00400000u - 0x10000000u);
    const float fe = __uint_as_float((jx >> 23) + (0x4B400000u - 127u)) - 12582912.0f;
    const float2 T = s_nlog[(jx >> 18) & 31u];
    const float m = __uint_as_float(fb + (0x3f800000u - 0x10000000u) - (jx & 0xff800000u));
    const float r = fmaf(m, T.x, -1.0f);
    float q = fmaf(r, c02, -0.25f);
    q = fmaf(r, q, 0.3333333432674407958984375f);
    q = fmaf(r, q, -0.5f);
    const float lp = fmaf(r * r, q, r);
    const float y = fmaf(fe, 0.693147182464599609375f, T.y) + lp;
    return k == 0u ? __uint_as_float(0xff800000u) : y;
}
// the table of cpm_native_logf_tab, staged in shared memory by every CTA
__device__ const float g_nlog_table[64] = {CPM_NLOG_TABLE};
#define CPM_SMEM_NLOG_FLOATS 64

// woodcockTracking (ppm/cl/transmittance.cl:126-144).  The step length of test k+1 depends only on the
// random stream, not on the outcome of test k, so the taps of sample k+1 are requested BEFORE sample k is
// blended and tested: two samples are in flight per lane and the texture latency hides behind the
// arithmetic of the previous test.  When test k ends the walk, the speculative draw is dropped (the
// stream is left exactly after the second number of test k), so the result is bit-identical to the
// sequential loop.
template <int FMT, int LAYOUT>
__device__ __forceinline__ float woodcock(const VolumeView& V, const float* s_alpha, const float2* s_nlog, int tfw, float ftfw, float3_ o,
                                          float3_ d, float tStart, float tEnd, cpm_rng& rng, unsigned& tests) {
    // tauMax = 1 (photontracer.cl:160): invTauMaxSampleBaseInterval = 1/(1*150), invTauMax = 1
    const float inv = 1.0f / 150.0f;
    cpm_rng spec = rng;
    float tA = advance_t(tStart, log_unit(cpm_rng_next(spec), s_nlog), inv), tB;
    float3_ pA = ray_at(o, tA, d);
    Taps A = fetch_taps<FMT, LAYOUT>(V, pA.x, pA.y, pA.z), B;
    float t;
    // one test: CUR is in flight since the previous test, NXT is requested here (the loop body is written
    // twice with the roles of A and B swapped so that no register moves are needed)
#define CPM_WOODCOCK_TEST(CUR, TCUR, NXT, TNXT)                                                            \
    {                                                                                                      \
        rng = spec;                 /* commit the first number of this test */                             \
        float u2 = cpm_rng_01(rng); /* second number of this test */                                       \
        spec = rng;                                                                                        \
        TNXT = advance_t(TCUR, log_unit(cpm_rng_next(spec), s_nlog), inv);                                        \
        const float3_ pn = ray_at(o, TNXT, d);                                                             \
        NXT = fetch_taps<FMT, LAYOUT>(V, pn.x, pn.y, pn.z);                                                \
        float opacity = sample_tf_alpha(s_alpha, tfw, ftfw, blend_taps<FMT>(V, CUR));                      \
        ++tests;                                                                                           \
        if (!(u2 >= opacity && TCUR <= tEnd)) {                                                            \
            t = TCUR;                                                                                      \
            break;                                                                                         \
        }                                                                                                  \
    }
    while (true) {
        CPM_WOODCOCK_TEST(A, tA, B, tB)
        CPM_WOODCOCK_TEST(B, tB, A, tA)
    }
#undef CPM_WOODCOCK_TEST
    return t;
}

// The same walk with the per-cell opacity bound (bound.cu).  A test continues the walk iff
// `u2 >= opacity && t <= tEnd`; with m >= opacity known for the cell, `u2 >= m` decides "continue" without the
// voxel and transfer-function fetches, and `!(t <= tEnd)` ends the walk whatever the opacity is.  Only the
// remaining CANDIDATE tests are evaluated in full.  Draws, positions and results are those of the plain loop.
//
// Scheduling: every lane scans through cheap tests (two draws, one log, one L1/L2-resident bound lookup)
// until it holds a candidate, leaves the volume, or has done A.scan tests; the warp then reconverges and all
// lanes holding a candidate fetch their taps together, so the expensive path runs warp-coherently instead of
// once per lane and iteration.
// Values the scan loop wants in REGISTERS.  sm_100 ALU instructions take no constant-bank operands and uniform registers
// do not live across the divergent loop, so whatever the compiler can trace back to a kernel argument or a literal it
// re-loads (LDC / LDCU) or re-builds (d * fc) every test.  Or-ing the bits with a kernel argument that is zero at run
// time hides the origin: the values are computed once per thread and stay put.
struct ScanRegs {
    float fcx, fcy, fcz;   // BoundGrid::fc (texture look-up) or ::fn (linear look-up)
    float c02;             // 0.2f, the leading coefficient of log_unit's polynomial
    float magic;           // 1.5 * 2^23 (bound_at)
};
__device__ __forceinline__ float opaque(float v, int zero) { return __uint_as_float(__float_as_uint(v) | (unsigned)zero); }
template <bool BTEX>
__device__ __forceinline__ ScanRegs scan_regs(const TraceArgs& A) {
    const float* f = BTEX ? A.bound.fc : A.bound.fn;
    return {opaque(f[0], A.zero), opaque(f[1], A.zero), opaque(f[2], A.zero), opaque(0.2f, A.zero), opaque(12582912.0f, A.zero)};
}

// BTEX: the bound comes from the 3-D texture -- the ray in cell coordinates (three FFMA) and one TEX whose unit floors,
// clamps and addresses; otherwise from the linear grid (bound_at).
template <int FMT, int LAYOUT, bool BTEX>
__device__ __forceinline__ float woodcock_bounded(const TraceArgs& A, const ScanRegs& C, const float* s_alpha, const float2* s_nlog,
                                                  int tfw, float ftfw, float3_ o, float3_ d, float tStart, float tEnd, cpm_rng& rng,
                                                  unsigned& tests, unsigned& fetched) {
    const VolumeView& V = A.vol;
    const float inv = 1.0f / 150.0f;
    const float hx = BTEX ? A.bound.hc : A.bound.hn[0], hy = BTEX ? A.bound.hc : A.bound.hn[1], hz = BTEX ? A.bound.hc : A.bound.hn[2];
    const CellRay R = {fmaf(o.x, C.fcx, hx), fmaf(o.y, C.fcy, hy), fmaf(o.z, C.fcz, hz),
#if CPM_TRACE_OPAQUE_RD
                       opaque(d.x * C.fcx, A.zero), opaque(d.y * C.fcy, A.zero), opaque(d.z * C.fcz, A.zero)};
#else
                       d.x * C.fcx, d.y * C.fcy, d.z * C.fcz};
#endif
    float t = tStart;
    unsigned k = 0u;                              // tests of this walk; a scan round ends when k reaches a multiple of A.scan
    const unsigned mask = (unsigned)A.scan - 1u;  // (a power of two)
    while (true) {
        bool cand = false, done = false;
        uint32_t k2 = 0u;
#pragma unroll 1
        while (true) {
            t = advance_t(t, log_unit(cpm_rng_next(rng), s_nlog, C.c02), inv);
            k2 = cpm_rng_next(rng);
            ++k;
            if (!(t <= tEnd)) {
                done = true;
                break;
            }
            // (a bound <= 0 -- transparent cell, negative when cpm_opacity_bound_clearance annotated it -- never
            // makes a candidate.  Using the clearance to skip look-ups was measured: lanes leave their transparent
            // stretches at different tests, the warp splits, and the walk gets slower, not faster.)
            float m = BTEX ? tex3DLod<float>(A.btex, fmaf(t, R.dx, R.ox), fmaf(t, R.dy, R.oy), fmaf(t, R.dz, R.oz), 0.0f)
                           : bound_at(A.bound, R, t, C.magic);
            if (!(cpm_u01(k2) >= m)) {
                cand = true;
                break;
            }
            if ((k & mask) == 0u) break;
        }
        if (cand) {
            ++fetched;
            const float3_ pc = ray_at(o, t, d);
            float v = sample_volume<FMT, LAYOUT>(V, pc.x, pc.y, pc.z);
            float opacity = sample_tf_alpha(s_alpha, tfw, ftfw, v);
            done = !(cpm_u01(k2) >= opacity);
        }
        if (done) break;
    }
    tests += k;
    return t;
}

// (Measured and rejected: the same loop software-pipelined -- draws, logarithm and bound look-up of test k+1 issued
// before test k is decided, the speculative test dropped when k ends the walk or is a candidate.  Bit-identical, but
// the second set of pending values costs 8 registers and ~10 copies per test: 0.562 -> 0.624 ms with the texture
// look-up, 0.592 -> 0.696 ms with the linear one.)

// BOUNDED: 0 = every test fetches (the reference's loop), 1 = opacity bound from the linear grid, 2 = from the texture
template <int FMT, int LAYOUT, int BOUNDED>
__global__ void __launch_bounds__(CPM_TRACE_THREADS, CPM_TRACE_MIN_CTAS) trace_kernel(const TraceArgs A) {
    extern __shared__ float2 s_nlog[];   // 32 x (rc, lc) of native_log, then the alpha column of the transfer function
    float* s_alpha = reinterpret_cast<float*>(s_nlog) + CPM_SMEM_NLOG_FLOATS;
    if (threadIdx.x < CPM_SMEM_NLOG_FLOATS) reinterpret_cast<float*>(s_nlog)[threadIdx.x] = g_nlog_table[threadIdx.x];
    for (int i = threadIdx.x; i < A.tf_width; i += blockDim.x) s_alpha[i] = A.tf[i].w;
    __syncthreads();

    const cpm_trace_params& P = A.p;
    const ScanRegs C = scan_regs<BOUNDED == 2>(A);
    unsigned tests = 0, fetched = 0;
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    int tid = -1;
    if (gid < A.n_work) {
        if (A.recompute) {
            int t = (int)A.recompute[gid] - P.photon_offset;
            if (t >= 0 && t < P.n_light_samples) tid = t;
        } else {
            tid = gid;
        }
    }
    if (tid >= 0) {
        const int tfw = A.tf_width;
        const float ftfw = (float)tfw;
        uint2 st = A.rng[P.photon_offset + tid];
        cpm_rng rng{st.x, st.y};
        float4 l0 = A.light_samples[2 * (size_t)tid], l1 = A.light_samples[2 * (size_t)tid + 1];
        float3_ o = {l0.x, l0.y, l0.z};
        float fmaxi = (float)P.max_interactions;
        float pr = l0.w / fmaxi, pg = l1.x / fmaxi, pb = l1.y / fmaxi;
        float3_ d = decode_direction(l1.z, l1.w);
        float2 ip = A.isect[tid];
        float tStart = ip.x, tEnd = ip.y;
        bool scatter = tStart < tEnd;
        unsigned n = 0;
        const unsigned maxI = (unsigned)P.max_interactions;

        if (P.flags & CPM_TRACE_NO_SINGLE_SCATTERING) {
            // photontracer.cl:143-157: the walk is executed even when the ray misses
            float t = BOUNDED ? woodcock_bounded<FMT, LAYOUT, BOUNDED == 2>(A, C, s_alpha, s_nlog, tfw, ftfw, o, d, tStart, tEnd, rng, tests, fetched)
                              : woodcock<FMT, LAYOUT>(A.vol, s_alpha, s_nlog, tfw, ftfw, o, d, tStart, tEnd, rng, tests);
            if (scatter) {
                o = ray_at(o, t, d);
                tStart = 0.0f;
                tEnd = CPM_FLT_MAX;
                float u1 = cpm_rng_01(rng), u2 = cpm_rng_01(rng);
                d = (P.phase_function == CPM_PHASE_HENYEY_GREENSTEIN)
                        ? sample_henyey_greenstein(d, P.material[0], u1, u2)
                        : uniform_sample_sphere(u1, u2);
                scatter = ray_box(P.aabb_min, P.aabb_max, o, d, tStart, tEnd);
                // power /= pdf, isotropic pdf = 1/(4 pi); HG restated with the same constant
                pr = pr / CPM_INV_4PI_F; pg = pg / CPM_INV_4PI_F; pb = pb / CPM_INV_4PI_F;
                tStart += 0.5f * P.step_size;
            }
        }
        while (scatter) {
            float t = BOUNDED ? woodcock_bounded<FMT, LAYOUT, BOUNDED == 2>(A, C, s_alpha, s_nlog, tfw, ftfw, o, d, tStart, tEnd, rng, tests, fetched)
                              : woodcock<FMT, LAYOUT>(A.vol, s_alpha, s_nlog, tfw, ftfw, o, d, tStart, tEnd, rng, tests);
            scatter = t <= tEnd;
            if (scatter) {
                o = ray_at(o, t, d);
                size_t pid = (size_t)P.photon_offset + (size_t)n * P.total_photons + tid;
                float2 ang = encode_direction(d);
                float vs = sample_volume<FMT, LAYOUT>(A.vol, o.x, o.y, o.z);
                float ca = sample_tf_alpha(s_alpha, tfw, ftfw, vs);  // color.w == scattering.w
                float albedo = ca / (ca + ca);                      // 0.5, or NaN when alpha == 0
                float den = cpm_fmax(ca, 0.01f);
                pr = pr / den; pg = pg / den; pb = pb / den;
                ++n;
                // short-circuit: the random number is drawn only when n < maxInteractions
                if (n < maxI && cpm_rng_01(rng) < albedo) {
                    pr *= albedo; pg *= albedo; pb *= albedo;
                    store_photon(A.photons, pid, o.x, o.y, o.z, pr, pg, pb, ang.x, ang.y);
                    tStart = 0.0f;
                    tEnd = CPM_FLT_MAX;
                    float u1 = cpm_rng_01(rng), u2 = cpm_rng_01(rng);
                    d = (P.phase_function == CPM_PHASE_HENYEY_GREENSTEIN)
                            ? sample_henyey_greenstein(d, P.material[0], u1, u2)
                            : uniform_sample_sphere(u1, u2);
                    scatter = ray_box(P.aabb_min, P.aabb_max, o, d, tStart, tEnd);
                    tStart += 0.5f * P.step_size;
                } else {
                    store_photon(A.photons, pid, o.x, o.y, o.z, pr, pg, pb, ang.x, ang.y);
                    pr = pg = pb = CPM_FLT_MAX;  // "absorbed" marker read by the detector
                    scatter = false;
                }
            }
        }
        float2 ang = encode_direction(d);
        for (unsigned i = n; i < maxI; ++i) {
            size_t pid = (size_t)P.photon_offset + (size_t)i * P.total_photons + tid;
            store_photon(A.photons, pid, CPM_FLT_MAX, CPM_FLT_MAX, CPM_FLT_MAX, pr, CPM_FLT_MAX, CPM_FLT_MAX, ang.x, ang.y);
        }
        if (P.flags & CPM_TRACE_PROGRESSIVE) A.rng[P.photon_offset + tid] = make_uint2(rng.x, rng.c);
    }
    if (A.tests) {
        unsigned long long v = tests;
        for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(A.tests, v);
        if (P.flags & CPM_TRACE_STATS) {
            unsigned long long f = BOUNDED ? fetched : tests;
            for (int off = 16; off; off >>= 1) f += __shfl_xor_sync(0xffffffffu, f, off);
            if ((threadIdx.x & 31) == 0 && f) atomicAdd(A.tests + 1, f);
        }
    }
}

// ---- the bounded walk as a WAVEFRONT: set-up | walk | interaction -------------------------------------------------------
// trace_kernel gives every photon a thread for its whole life: a warp runs until its longest walk ends, and the scan
// loop executes with about half of its lanes on C4 (ncu: 18.8 lanes per instruction over the kernel).  Photons are
// independent (per-photon RNG stream, result independent of lane / warp / GPU), so lanes can be re-used -- but only if
// taking the next photon is CHEAP: with set-up (two sincos, divisions) and the interaction (acos, atan2, voxel taps,
// divisions, stores) inside the refill, those ~800 instructions per photon run at the few lanes that happen to be idle
// and eat the gain (measured: 0.59 -> 0.70 ms).  Here the three phases are three kernels:
//   walk_setup_kernel   one thread per work item, full warps: light sample -> walk entry (origin, direction, t, tEnd,
//                       stream state), 48 bytes; rays that miss the volume are finalised here
//   walk_kernel         persistent warps; an idle lane takes the next entry from a global cursor (a 48-byte load), walks
//                       it with the scan loop of woodcock_bounded and writes back t and the stream state (12 bytes)
//   walk_finish_kernel  one thread per entry, full warps: the interaction at t (store the record, scatter or absorb); a
//                       photon that scatters again rewrites its entry for the next round, up to maxInteractions rounds
// Same draws, same positions, same records as trace_kernel and the oracle (tests/test_bound.py runs both).
// Measured on C4 (2 M re-traced photons): walk_kernel runs with 23.9 lanes per instruction instead of 18.8 and takes
// 0.47 ms, but the two streaming kernels around it move 216 B per photon and cost 0.12 ms: 0.594 ms against 0.559 ms for
// trace_kernel.  OPT-IN (CPM_TRACE_LANE_REFILL / CPM_TRACE_WAVEFRONT=<batch>); it pays where walks are long compared
// with the interaction, i.e. thin media and several scattering events.
struct WalkEntry {
    float4 a;   // origin, t (start of the walk; after walk_kernel: where it ended)
    float4 b;   // direction, tEnd
    uint4 c;    // stream state (x, c), photon id within the light (tid), interactions so far | WALK_ACTIVE
};
constexpr unsigned WALK_ACTIVE = 0x80000000u;
struct WaveArgs {
    WalkEntry* entries;
    float4* power;       // (r, g, b, -) carried between interactions
    unsigned* cursor;    // next unclaimed entry (zeroed before every walk_kernel launch)
    int refill_min;      // idle lanes a warp collects before it fetches entries for them
};

// empty slots n .. maxI-1, and the stream state of a progressive trace (photontracer.cl:199-214)
__device__ __forceinline__ void finalize_photon(const TraceArgs& A, int tid, unsigned n, float pr, float3_ d, cpm_rng rng) {
    const cpm_trace_params& P = A.p;
    const unsigned maxI = (unsigned)P.max_interactions;
    float2 ang = encode_direction(d);
    for (unsigned i = n; i < maxI; ++i) {
        size_t pid = (size_t)P.photon_offset + (size_t)i * P.total_photons + tid;
        store_photon(A.photons, pid, CPM_FLT_MAX, CPM_FLT_MAX, CPM_FLT_MAX, pr, CPM_FLT_MAX, CPM_FLT_MAX, ang.x, ang.y);
    }
    if (P.flags & CPM_TRACE_PROGRESSIVE) A.rng[P.photon_offset + tid] = make_uint2(rng.x, rng.c);
}

__global__ void __launch_bounds__(128) walk_setup_kernel(const TraceArgs A, const WaveArgs W) {
    const cpm_trace_params& P = A.p;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= A.n_work) return;
    int tid = gid;
    if (A.recompute) {
        int t = (int)A.recompute[gid] - P.photon_offset;
        tid = (t >= 0 && t < P.n_light_samples) ? t : -1;   // not this light's photon
    }
    if (tid < 0) {
        W.entries[gid].c = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    uint2 st = A.rng[P.photon_offset + tid];
    float4 l0 = A.light_samples[2 * (size_t)tid], l1 = A.light_samples[2 * (size_t)tid + 1];
    float fmaxi = (float)P.max_interactions;
    float pr = l0.w / fmaxi, pg = l1.x / fmaxi, pb = l1.y / fmaxi;
    float3_ d = decode_direction(l1.z, l1.w);
    float2 ip = A.isect[tid];
    if (ip.x < ip.y) {
        W.entries[gid].a = make_float4(l0.x, l0.y, l0.z, ip.x);
        W.entries[gid].b = make_float4(d.x, d.y, d.z, ip.y);
        W.entries[gid].c = make_uint4(st.x, st.y, (unsigned)tid, WALK_ACTIVE);
        W.power[gid] = make_float4(pr, pg, pb, 0.0f);
    } else {
        W.entries[gid].c = make_uint4(0u, 0u, 0u, 0u);
        finalize_photon(A, tid, 0u, pr, d, cpm_rng{st.x, st.y});   // the ray misses the volume
    }
}

template <int FMT, int LAYOUT, bool BTEX>
__global__ void __launch_bounds__(128, 9) walk_kernel(const TraceArgs A, const WaveArgs W) {
    extern __shared__ float2 s_nlog[];
    float* s_alpha = reinterpret_cast<float*>(s_nlog) + CPM_SMEM_NLOG_FLOATS;
    if (threadIdx.x < CPM_SMEM_NLOG_FLOATS) reinterpret_cast<float*>(s_nlog)[threadIdx.x] = g_nlog_table[threadIdx.x];
    for (int i = threadIdx.x; i < A.tf_width; i += blockDim.x) s_alpha[i] = A.tf[i].w;
    __syncthreads();
    const VolumeView& V = A.vol;
    const ScanRegs C = scan_regs<BTEX>(A);
    const int tfw = A.tf_width;
    const float ftfw = (float)tfw;
    const float inv = 1.0f / 150.0f;
    const float hx = BTEX ? A.bound.hc : A.bound.hn[0], hy = BTEX ? A.bound.hc : A.bound.hn[1], hz = BTEX ? A.bound.hc : A.bound.hn[2];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned n_work = (unsigned)A.n_work;
    bool walking = false, exhausted = false;   // exhausted is warp-uniform: the cursor has passed the last entry
    unsigned g = 0u;                           // entry this lane walks
    float3_ o = {0.f, 0.f, 0.f}, d = {0.f, 0.f, 1.f};
    CellRay R = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float t = 0.f, tEnd = 0.f;
    cpm_rng rng{0u, 0u};
    unsigned tests = 0, fetched = 0;
    while (true) {
        const unsigned idle = __ballot_sync(0xffffffffu, !walking);
        if (!exhausted && ((int)__popc(idle) >= W.refill_min || idle == 0xffffffffu)) {
            // ---- idle lanes take the next entries --------------------------------------------------------------------------
            const int leader = __ffs(idle) - 1;
            unsigned base = 0u;
            if ((int)lane == leader) base = atomicAdd(W.cursor, (unsigned)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, leader);
            exhausted = base + (unsigned)__popc(idle) >= n_work;
            if (!walking) {
                const unsigned e = base + (unsigned)__popc(idle & ((1u << lane) - 1u));
                if (e < n_work) {
                    const uint4 c = W.entries[e].c;
                    if (c.w & WALK_ACTIVE) {
                        const float4 a = W.entries[e].a, b = W.entries[e].b;
                        g = e;
                        o = {a.x, a.y, a.z};
                        d = {b.x, b.y, b.z};
                        t = a.w;
                        tEnd = b.w;
                        rng = cpm_rng{c.x, c.y};
                        R = {fmaf(o.x, C.fcx, hx), fmaf(o.y, C.fcy, hy), fmaf(o.z, C.fcz, hz), d.x * C.fcx, d.y * C.fcy, d.z * C.fcz};
                        walking = true;
                    }
                }
            }
            continue;
        }
        if (idle == 0xffffffffu) break;   // nothing walks and nothing is left
        // ---- one scan round (woodcock_bounded) for the walking lanes; the trip count is warp-uniform -----------------------
        bool live = walking, cand = false, done = false;
        uint32_t k2 = 0u;
#pragma unroll 1
        for (int j = 0; j < A.scan; ++j) {
            if (live) {
                t = advance_t(t, log_unit(cpm_rng_next(rng), s_nlog, C.c02), inv);
                k2 = cpm_rng_next(rng);
                ++tests;
                if (!(t <= tEnd)) {
                    done = true;
                    live = false;
                } else {
                    float m = BTEX ? tex3DLod<float>(A.btex, fmaf(t, R.dx, R.ox), fmaf(t, R.dy, R.oy), fmaf(t, R.dz, R.oz), 0.0f)
                                   : bound_at(A.bound, R, t, C.magic);
                    if (!(cpm_u01(k2) >= m)) {
                        cand = true;
                        live = false;
                    }
                }
            }
        }
        if (cand) {
            ++fetched;
            const float3_ pc = ray_at(o, t, d);
            float v = sample_volume<FMT, LAYOUT>(V, pc.x, pc.y, pc.z);
            float opacity = sample_tf_alpha(s_alpha, tfw, ftfw, v);
            done = !(cpm_u01(k2) >= opacity);
        }
        if (done) {
            W.entries[g].a.w = t;
            *reinterpret_cast<uint2*>(&W.entries[g].c) = make_uint2(rng.x, rng.c);
            walking = false;
        }
    }
    if (A.tests) {
        unsigned long long v = tests;
        for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0 && v) atomicAdd(A.tests, v);
        if (A.p.flags & CPM_TRACE_STATS) {
            unsigned long long f = fetched;
            for (int off = 16; off; off >>= 1) f += __shfl_xor_sync(0xffffffffu, f, off);
            if (lane == 0 && f) atomicAdd(A.tests + 1, f);
        }
    }
}

// photontracer.cl:161-197 for the walk that just ended
template <int FMT, int LAYOUT>
__global__ void __launch_bounds__(128) walk_finish_kernel(const TraceArgs A, const WaveArgs W) {
    extern __shared__ float s_alpha_f[];
    for (int i = threadIdx.x; i < A.tf_width; i += blockDim.x) s_alpha_f[i] = A.tf[i].w;
    __syncthreads();
    const cpm_trace_params& P = A.p;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= A.n_work) return;
    const uint4 c = W.entries[gid].c;
    if (!(c.w & WALK_ACTIVE)) return;
    const float4 a = W.entries[gid].a, b = W.entries[gid].b;
    const float4 pw = W.power[gid];
    const int tfw = A.tf_width;
    const float ftfw = (float)tfw;
    const int tid = (int)c.z;
    unsigned n = c.w & ~WALK_ACTIVE;
    const unsigned maxI = (unsigned)P.max_interactions;
    float3_ o = {a.x, a.y, a.z}, d = {b.x, b.y, b.z};
    const float t = a.w;
    float tEnd = b.w;
    cpm_rng rng{c.x, c.y};
    float pr = pw.x, pg = pw.y, pb = pw.z;
    bool scatter = t <= tEnd;
    if (scatter) {
        o = ray_at(o, t, d);
        size_t pid = (size_t)P.photon_offset + (size_t)n * P.total_photons + tid;
        float2 ang = encode_direction(d);
        float vs = sample_volume<FMT, LAYOUT>(A.vol, o.x, o.y, o.z);
        float ca = sample_tf_alpha(s_alpha_f, tfw, ftfw, vs);  // color.w == scattering.w
        float albedo = ca / (ca + ca);                        // 0.5, or NaN when alpha == 0
        float den = cpm_fmax(ca, 0.01f);
        pr = pr / den; pg = pg / den; pb = pb / den;
        ++n;
        if (n < maxI && cpm_rng_01(rng) < albedo) {
            pr *= albedo; pg *= albedo; pb *= albedo;
            store_photon(A.photons, pid, o.x, o.y, o.z, pr, pg, pb, ang.x, ang.y);
            float tStart = 0.0f;
            tEnd = CPM_FLT_MAX;
            float u1 = cpm_rng_01(rng), u2 = cpm_rng_01(rng);
            d = (P.phase_function == CPM_PHASE_HENYEY_GREENSTEIN) ? sample_henyey_greenstein(d, P.material[0], u1, u2)
                                                                  : uniform_sample_sphere(u1, u2);
            scatter = ray_box(P.aabb_min, P.aabb_max, o, d, tStart, tEnd);
            tStart += 0.5f * P.step_size;
            if (scatter) {   // walks again in the next round
                W.entries[gid].a = make_float4(o.x, o.y, o.z, tStart);
                W.entries[gid].b = make_float4(d.x, d.y, d.z, tEnd);
                W.entries[gid].c = make_uint4(rng.x, rng.c, (unsigned)tid, n | WALK_ACTIVE);
                W.power[gid] = make_float4(pr, pg, pb, 0.0f);
                return;
            }
        } else {
            store_photon(A.photons, pid, o.x, o.y, o.z, pr, pg, pb, ang.x, ang.y);
            pr = pg = pb = CPM_FLT_MAX;  // "absorbed" marker read by the detector
        }
    }
    W.entries[gid].c.w = 0u;
    finalize_photon(A, tid, n, pr, d, rng);
}

template <int FMT, int LAYOUT, int BOUNDED>
int launch2(cpm_ctx* ctx, const TraceArgs& a) {
    size_t smem = ((size_t)a.tf_width + CPM_SMEM_NLOG_FLOATS) * sizeof(float);
    if (smem > 48 * 1024)
        CPM_CUDA(ctx, cudaFuncSetAttribute(trace_kernel<FMT, LAYOUT, BOUNDED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CPM_LAUNCH(ctx, (trace_kernel<FMT, LAYOUT, BOUNDED>), cpm_div_up(a.n_work, CPM_TRACE_THREADS), CPM_TRACE_THREADS, smem, a);
    return CPM_OK;
}
template <int FMT, int LAYOUT, bool BTEX>
int launch_wave2(cpm_ctx* ctx, const TraceArgs& a, const WaveArgs& w) {
    const size_t smem = ((size_t)a.tf_width + CPM_SMEM_NLOG_FLOATS) * sizeof(float), smem_f = (size_t)a.tf_width * sizeof(float);
    if (smem > 48 * 1024) {
        CPM_CUDA(ctx, cudaFuncSetAttribute(walk_kernel<FMT, LAYOUT, BTEX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CPM_CUDA(ctx, cudaFuncSetAttribute(walk_finish_kernel<FMT, LAYOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
    }
    static int per_sm = 0;
    if (!per_sm) {
        CPM_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, walk_kernel<FMT, LAYOUT, BTEX>, 128, smem));
        per_sm = per_sm > 0 ? per_sm : 1;
    }
    // persistent grid: every resident CTA slot of the device, but no more warps than entries / 32
    const unsigned grid = std::max(1u, std::min((unsigned)(ctx->sm_count * per_sm), cpm_div_up(a.n_work, 128)));
    const unsigned blocks = cpm_div_up(a.n_work, 128);
    CPM_LAUNCH(ctx, walk_setup_kernel, blocks, 128, 0, a, w);
    for (int round = 0; round < a.p.max_interactions; ++round) {
        CPM_CUDA(ctx, cudaMemsetAsync(w.cursor, 0, sizeof(unsigned), ctx->stream));
        CPM_LAUNCH(ctx, (walk_kernel<FMT, LAYOUT, BTEX>), grid, 128, smem, a, w);
        CPM_LAUNCH(ctx, (walk_finish_kernel<FMT, LAYOUT>), blocks, 128, smem_f, a, w);
    }
    return CPM_OK;
}
template <int FMT, int LAYOUT>
int launch_wave(cpm_ctx* ctx, const TraceArgs& a, int refill_min) {
    const size_t need = (size_t)a.n_work * (sizeof(WalkEntry) + sizeof(float4));
    if (need > ctx->walk_bytes) {
        CPM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->walk_buf) CPM_CUDA(ctx, cudaFree(ctx->walk_buf));
        ctx->walk_buf = nullptr;
        ctx->walk_bytes = 0;
        CPM_CUDA(ctx, cudaMalloc(&ctx->walk_buf, need));
        ctx->walk_bytes = need;
    }
    if (!ctx->trace_cursor) CPM_CUDA(ctx, cudaMalloc((void**)&ctx->trace_cursor, 256));
    WaveArgs w;
    w.entries = (WalkEntry*)ctx->walk_buf;
    w.power = (float4*)((char*)ctx->walk_buf + (size_t)a.n_work * sizeof(WalkEntry));
    w.cursor = ctx->trace_cursor;
    w.refill_min = refill_min;
    return a.btex ? launch_wave2<FMT, LAYOUT, true>(ctx, a, w) : launch_wave2<FMT, LAYOUT, false>(ctx, a, w);
}
template <int FMT, int LAYOUT>
int launch(cpm_ctx* ctx, const TraceArgs& a) {
    // the wavefront form (set-up | persistent walk with lane refill | interaction): bounded walks that start at the light
    // sample.  NO_SINGLE_SCATTERING walks start differently and stay on trace_kernel.  CPM_TRACE_WAVEFRONT=0 / =<refill
    // batch> in the environment overrides the default (A/B runs); the CPM_TRACE_LANE_REFILL flag asks for it explicitly.
    static const int wave_env = getenv("CPM_TRACE_WAVEFRONT") ? atoi(getenv("CPM_TRACE_WAVEFRONT")) : -1;
    const int refill = wave_env >= 0 ? wave_env : (((a.p.flags & CPM_TRACE_LANE_REFILL) || CPM_TRACE_WAVEFRONT_DEFAULT) ? 8 : 0);
    if (a.bound.g && refill > 0 && !(a.p.flags & CPM_TRACE_NO_SINGLE_SCATTERING))
        return launch_wave<FMT, LAYOUT>(ctx, a, std::min(refill, 32));
    if (a.btex) return launch2<FMT, LAYOUT, 2>(ctx, a);
    return a.bound.g ? launch2<FMT, LAYOUT, 1>(ctx, a) : launch2<FMT, LAYOUT, 0>(ctx, a);
}

}  // namespace

extern "C" int cpm_trace_photons(cpm_ctx* ctx, const cpm_volume* vol, const float* tf_rgba, int tf_width,
                                 const cpm_trace_params* params, const float* light_samples, const float* intersections,
                                 const uint32_t* recompute_index, int n_recompute, float* photons, uint32_t* rng_state,
                                 unsigned long long* collision_tests) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, vol && tf_rgba && params && light_samples && intersections && photons && rng_state, "null argument");
    CPM_REQUIRE(ctx, tf_width >= 1 && tf_width <= 32768, "tf_width out of range");
    CPM_REQUIRE(ctx, params->max_interactions >= 1, "max_interactions must be >= 1");
    CPM_REQUIRE(ctx, params->n_light_samples >= 0 && params->photon_offset >= 0, "negative size");
    CPM_REQUIRE(ctx, params->total_photons >= params->photon_offset + params->n_light_samples,
                "total_photons smaller than photon_offset + n_light_samples");
    CPM_REQUIRE(ctx, recompute_index == nullptr || n_recompute >= 0, "negative n_recompute");
    TraceArgs a;
    a.p = *params;
    a.vol = make_view(vol);
    a.tf = (const float4*)tf_rgba;
    a.tf_width = tf_width;
    a.light_samples = (const float4*)light_samples;
    a.isect = (const float2*)intersections;
    a.recompute = recompute_index;
    a.n_work = recompute_index ? n_recompute : params->n_light_samples;
    a.photons = (float4*)photons;
    a.rng = (uint2*)rng_state;
    a.tests = collision_tests;
    static const int scan_env = getenv("CPM_TRACE_SCAN") ? atoi(getenv("CPM_TRACE_SCAN")) : 0;   // tuning sweeps only
    a.scan = scan_env > 0 ? scan_env : CPM_SCAN;
    while (a.scan & (a.scan - 1)) a.scan &= a.scan - 1;   // a power of two (woodcock_bounded masks with scan - 1)
    a.zero = 0;
    CPM_REQUIRE(ctx, make_bound_grid(a.bound, params->opacity_bound, vol->dims, params->bound_cell_log2),
                "bound_cell_log2 must be in 0..8 and the bound grid smaller than 2^31 cells");
    a.btex = 0;
    if (params->opacity_bound_tex) {
        const cpm_bound_tex* bt = params->opacity_bound_tex;
        // the texture needs the cell geometry only: a placeholder pointer switches the grid parameters on
        CPM_REQUIRE(ctx, make_bound_grid(a.bound, params->opacity_bound ? params->opacity_bound : (const float*)bt, vol->dims,
                                         params->bound_cell_log2), "bound_cell_log2 must be in 0..8");
        int gd[3];
        cpm_bound_grid_dims(vol->dims, params->bound_cell_log2, gd);
        CPM_REQUIRE(ctx, gd[0] == bt->dims[0] && gd[1] == bt->dims[1] && gd[2] == bt->dims[2],
                    "opacity_bound_tex was created for another grid");
        static const bool tex_off = getenv("CPM_TRACE_BOUND_TEX") && atoi(getenv("CPM_TRACE_BOUND_TEX")) == 0;   // A/B runs
        if (!(tex_off && params->opacity_bound)) a.btex = bt->tex;
    }
    if (a.n_work == 0) return CPM_OK;
#define CPM_DISPATCH(F)                                                                  \
    return vol->layout == CPM_VOLUME_TEXTURE ? launch<F, CPM_VOLUME_TEXTURE>(ctx, a)      \
                                             : launch<F, CPM_VOLUME_LINEAR>(ctx, a);
    switch (vol->format) {
        case CPM_FMT_U8: CPM_DISPATCH(CPM_FMT_U8)
        case CPM_FMT_U16: CPM_DISPATCH(CPM_FMT_U16)
        default: CPM_DISPATCH(CPM_FMT_F32)
    }
#undef CPM_DISPATCH
}
