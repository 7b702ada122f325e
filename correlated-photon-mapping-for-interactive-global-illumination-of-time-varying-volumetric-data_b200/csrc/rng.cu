// rng.cu -- MWC64X per-photon random streams (north-star subsystem 1).
//
// Replaces rng/cl/randstategen.cl (MWC64X_GenerateRandomState), rng/cl/skip_mwc.cl
// (MWC_SeedImpl_Mod64 and friends) and rng/cl/randomnumbergenerator.cl.
//
// B200 design note: the reference computes A^dist mod M with shift-and-add modular
// multiplication (~64*64 modular additions per multiply, "about 2^15 instructions").  Here a
// modular multiply is one 64x64->128 bit product (IMAD.WIDE chains) followed by an exact
// reduction that uses the special form of the modulus, M = A*2^32 - 1:
//     N = q*(M+1) + r  =>  N mod M = (q + r) mod M ,   (M+1) = A << 32
// so the 128-bit remainder needs only a 96-by-32 bit long division.  The result is the exact
// residue, hence bit-identical to the reference's.  Integer ALU bound, not HBM bound:
// 12 B of traffic per stream against ~10^4 integer instructions.
#include "common.cuh"

namespace {

__host__ __device__ __forceinline__ uint64_t addmod(uint64_t a, uint64_t b) {
    uint64_t v = a + b;
    if (v >= CPM_MWC64X_M || v < a) v -= CPM_MWC64X_M;
    return v;
}

__device__ __forceinline__ uint64_t mulmod(uint64_t a, uint64_t b) {
    const uint64_t A = CPM_MWC64X_A;
    uint64_t lo = a * b;
    uint64_t hi = __umul64hi(a, b);
    // T = N >> 32 (96 bits) = hi : (lo >> 32);   q = T / A, rem = T % A
    uint64_t q_hi = hi / A;
    uint64_t r1 = hi - q_hi * A;                       // < A < 2^32
    uint64_t t2 = (r1 << 32) | (lo >> 32);
    uint64_t q_lo = t2 / A;                            // < 2^32
    uint64_t r2 = t2 - q_lo * A;
    uint64_t q = (q_hi << 32) + q_lo;                  // <= M
    uint64_t r = (r2 << 32) | (lo & 0xffffffffull);    // <= M
    if (q >= CPM_MWC64X_M) q -= CPM_MWC64X_M;
    if (r >= CPM_MWC64X_M) r -= CPM_MWC64X_M;
    return addmod(q, r);
}

__device__ uint64_t powmod(uint64_t a, uint64_t e) {
    uint64_t sqr = a, acc = 1;
    while (e != 0) {
        if (e & 1) acc = mulmod(acc, sqr);
        sqr = mulmod(sqr, sqr);
        e >>= 1;
    }
    return acc;
}

// One thread per stream.  state[2i] holds the host-chosen base offset on entry.
__global__ void __launch_bounds__(256) seed_streams_kernel(uint2* __restrict__ state, size_t n, uint64_t gap,
                                                           uint64_t first_stream) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t BASEID = 4077358422479273989ull;
    uint64_t base = state[i].x;
    uint64_t dist = base + (first_stream + i) * gap;  // wraps mod 2^64 like the reference's ulong
    uint64_t m = powmod(CPM_MWC64X_A, dist);
    uint64_t x = mulmod(BASEID, m);
    state[i] = make_uint2((uint32_t)(x / CPM_MWC64X_A), (uint32_t)(x % CPM_MWC64X_A));
}

__global__ void __launch_bounds__(256) uniform_kernel(uint2* __restrict__ state, size_t n, int per_stream,
                                                      float* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint2 s = state[i];
    cpm_rng r{s.x, s.y};
    for (int k = 0; k < per_stream; ++k) out[i * (size_t)per_stream + k] = cpm_rng_01(r);
    state[i] = make_uint2(r.x, r.c);
}

// glibc random_r TYPE_3 (degree 31, separation 3), the generator behind rand()/srand().
struct glibc_rand {
    int32_t r[34];
    int f, b;  // front / rear indices into a 31-entry ring
    int32_t ring[31];
    void seed(uint32_t s) {
        if (s == 0) s = 1;
        ring[0] = (int32_t)s;
        for (int i = 1; i < 31; ++i) {
            // 16807 * x mod (2^31 - 1) without overflow (Schrage)
            int32_t hi = ring[i - 1] / 127773;
            int32_t lo = ring[i - 1] % 127773;
            int32_t w = 16807 * lo - 2836 * hi;
            if (w < 0) w += 2147483647;
            ring[i] = w;
        }
        f = 3;
        b = 0;
        for (int i = 0; i < 310; ++i) next();
    }
    uint32_t next() {
        uint32_t v = (uint32_t)ring[f] + (uint32_t)ring[b];
        ring[f] = (int32_t)v;
        uint32_t res = v >> 1;
        f = (f + 1) % 31;
        b = (b + 1) % 31;
        return res;
    }
};

}  // namespace

extern "C" {

int cpm_rng_host_base_offsets(uint32_t seed, uint32_t* state_host, size_t n) {
    if (!state_host && n) return CPM_E_INVALID;
    glibc_rand g;
    g.seed(seed);
    for (size_t i = 0; i < n; ++i) {
        state_host[2 * i] = g.next();
        state_host[2 * i + 1] = 0;
    }
    return CPM_OK;
}

int cpm_rng_host_base_offsets_range(uint32_t seed, uint64_t first, uint32_t* state_host, size_t n) {
    if (!state_host && n) return CPM_E_INVALID;
    glibc_rand g;
    g.seed(seed);
    for (uint64_t i = 0; i < first; ++i) g.next();
    for (size_t i = 0; i < n; ++i) {
        state_host[2 * i] = g.next();
        state_host[2 * i + 1] = 0;
    }
    return CPM_OK;
}

int cpm_rng_seed_streams(cpm_ctx* ctx, uint32_t* state, size_t n, uint64_t stream_gap, uint64_t first_stream) {
    if (!ctx) return CPM_E_INVALID;
    if (n == 0) return CPM_OK;
    CPM_REQUIRE(ctx, state != nullptr, "state is NULL");
    CPM_LAUNCH(ctx, seed_streams_kernel, cpm_div_up(n, 256), 256, 0, (uint2*)state, n, stream_gap, first_stream);
    return CPM_OK;
}

int cpm_rng_uniform(cpm_ctx* ctx, uint32_t* state, size_t n, int samples_per_stream, float* out) {
    if (!ctx) return CPM_E_INVALID;
    if (n == 0) return CPM_OK;
    CPM_REQUIRE(ctx, state && out, "null buffer");
    CPM_REQUIRE(ctx, samples_per_stream >= 1, "samples_per_stream must be >= 1");
    CPM_LAUNCH(ctx, uniform_kernel, cpm_div_up(n, 256), 256, 0, (uint2*)state, n, samples_per_stream, out);
    return CPM_OK;
}

}  // extern "C"
