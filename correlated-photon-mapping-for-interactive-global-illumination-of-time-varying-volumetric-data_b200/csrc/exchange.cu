// exchange.cu -- the one exchange step of the sharded photon path (option B): the sum of the per-GPU
// light volumes, as ONE kernel over NVLink peer memory instead of a ring of NCCL kernels.
//
// Every rank holds a snapshot of its light volume in a buffer that all ranks of the node have mapped (symmetric
// memory: the same allocation size on every GPU, peer pointers exchanged once).  Rank r owns the r-th slice of the
// volume; it forms the sum of that slice over all ranks and writes it back into EVERY rank's buffer, so that after
// the call each buffer holds the whole sum (a two-shot all-reduce: (G-1)/G of the volume in, (G-1)/G out per GPU).
//
//  * multicast pointer given (NVSwitch, NVLS): `multimem.ld_reduce` -- the switch adds the G copies and returns one
//    value -- and `multimem.st` -- the switch stores to all G copies: 1/G of the volume over this GPU's links in
//    either direction, no reduction arithmetic on the SMs.
//  * otherwise: plain peer loads in rank order (a fixed order: every rank ends up with the same bits) and peer stores.
//
// The kernel contains no inter-GPU synchronisation: the caller separates "all snapshots written" -> kernel -> "all
// slices stored" with two barriers on the same stream (torch's symmetric-memory signal-pad barrier in
// sharding.PeerLightVolumeExchange).  It runs on a side stream next to the following frame's kernels, so the grid is
// kept small: the transfer is latency-, not SM-bound.
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace {

struct PeerArgs {
    float4* bufs[CPM_MAX_PEERS];
    float4* mc;
    size_t begin, end;   // this rank's slice, in float4 units
    int world;
};

__device__ __forceinline__ float4 mc_ld_reduce(const float4* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float4* p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

template <bool MULTICAST>
__global__ void __launch_bounds__(256) allreduce_peer_kernel(const PeerArgs A) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = A.begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (MULTICAST) {
        // four independent reductions in flight per thread
        for (; i + 3 * stride < A.end; i += 4 * stride) {
            float4 a = mc_ld_reduce(A.mc + i), b = mc_ld_reduce(A.mc + i + stride);
            float4 c = mc_ld_reduce(A.mc + i + 2 * stride), d = mc_ld_reduce(A.mc + i + 3 * stride);
            mc_st(A.mc + i, a);
            mc_st(A.mc + i + stride, b);
            mc_st(A.mc + i + 2 * stride, c);
            mc_st(A.mc + i + 3 * stride, d);
        }
        for (; i < A.end; i += stride) mc_st(A.mc + i, mc_ld_reduce(A.mc + i));
    } else {
        for (; i < A.end; i += 2 * stride) {
            const bool two = i + stride < A.end;
            float4 v[CPM_MAX_PEERS], w[CPM_MAX_PEERS];
#pragma unroll
            for (int r = 0; r < CPM_MAX_PEERS; ++r)
                if (r < A.world) {
                    v[r] = A.bufs[r][i];
                    if (two) w[r] = A.bufs[r][i + stride];
                }
            float4 s = v[0], t = two ? w[0] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 1; r < CPM_MAX_PEERS; ++r)
                if (r < A.world) {   // rank order: the same sum, bit for bit, whoever computes it
                    s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w;
                    if (two) { t.x += w[r].x; t.y += w[r].y; t.z += w[r].z; t.w += w[r].w; }
                }
#pragma unroll
            for (int r = 0; r < CPM_MAX_PEERS; ++r)
                if (r < A.world) {
                    A.bufs[r][i] = s;
                    if (two) A.bufs[r][i + stride] = t;
                }
        }
    }
}

}  // namespace

extern "C" int cpm_allreduce_peer_f32(cpm_ctx* ctx, float* const* peer_buffers, float* multicast, size_t n_floats, int rank,
                                      int world, int max_ctas) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, world >= 1 && world <= CPM_MAX_PEERS && rank >= 0 && rank < world, "rank / world out of range");
    CPM_REQUIRE(ctx, multicast || peer_buffers, "null argument");
    CPM_REQUIRE(ctx, n_floats % 4 == 0, "n_floats must be a multiple of 4");
    if (n_floats == 0 || world == 1) return CPM_OK;
    PeerArgs a;
    memset(&a, 0, sizeof(a));
    a.mc = (float4*)multicast;
    a.world = world;
    CPM_REQUIRE(ctx, (uintptr_t)multicast % 16 == 0, "multicast pointer must be 16-byte aligned");
    if (!multicast)
        for (int r = 0; r < world; ++r) {
            CPM_REQUIRE(ctx, peer_buffers[r] && (uintptr_t)peer_buffers[r] % 16 == 0, "peer buffers must be 16-byte aligned");
            a.bufs[r] = (float4*)peer_buffers[r];
        }
    const size_t n4 = n_floats / 4, per = (n4 + world - 1) / world;
    a.begin = std::min(n4, per * (size_t)rank);
    a.end = std::min(n4, a.begin + per);
    if (a.begin >= a.end) return CPM_OK;
    const unsigned want = cpm_div_up(a.end - a.begin, 256 * 4);
    const unsigned grid = std::max(1u, std::min(want, (unsigned)(max_ctas > 0 ? max_ctas : 48)));
    if (multicast)
        CPM_LAUNCH(ctx, allreduce_peer_kernel<true>, grid, 256, 0, a);
    else
        CPM_LAUNCH(ctx, allreduce_peer_kernel<false>, grid, 256, 0, a);
    return CPM_OK;
}
