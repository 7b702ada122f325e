/* orc_rng.c -- oracle restatement of MWC64X.  TEST INFRASTRUCTURE (see cpm_oracle.h).
 * Follows the reference line by line, including the shift-and-add modular multiply. */
#include <stdlib.h>

#include "cpm_oracle.h"

#define MWC64X_A 4294883355u                /* rng/cl/random.cl:46 */
#define MWC64X_M 18446383549859758079ull    /* rng/cl/random.cl:47 */
#define MWC_BASEID 4077358422479273989ull   /* rng/cl/skip_mwc.cl:98 */

/* rng/cl/skip_mwc.cl:40-46 */
static uint64_t MWC_AddMod64(uint64_t a, uint64_t b, uint64_t M) {
    uint64_t v = a + b;
    if ((v >= M) || (v < a)) v = v - M;
    return v;
}
/* rng/cl/skip_mwc.cl:54-64 */
static uint64_t MWC_MulMod64(uint64_t a, uint64_t b, uint64_t M) {
    uint64_t r = 0;
    while (a != 0) {
        if (a & 1) r = MWC_AddMod64(r, b, M);
        b = MWC_AddMod64(b, b, M);
        a = a >> 1;
    }
    return r;
}
/* rng/cl/skip_mwc.cl:71-81 */
static uint64_t MWC_PowMod64(uint64_t a, uint64_t e, uint64_t M) {
    uint64_t sqr = a, acc = 1;
    while (e != 0) {
        if (e & 1) acc = MWC_MulMod64(acc, sqr, M);
        sqr = MWC_MulMod64(sqr, sqr, M);
        e = e >> 1;
    }
    return acc;
}

/* rng/cl/random.cl:58-69 */
void orc_rng_step(uint32_t* x, uint32_t* c) {
    uint32_t X = *x, C = *c;
    uint32_t Xn = MWC64X_A * X + C;
    uint32_t carry = (uint32_t)(Xn < C);
    uint32_t Cn = (uint32_t)(((uint64_t)MWC64X_A * X) >> 32) + carry; /* mad_hi(A, X, carry) */
    *x = Xn;
    *c = Cn;
}
/* rng/cl/random.cl:85-90 */
static uint32_t MWC64X_NextUint(uint32_t* x, uint32_t* c) {
    uint32_t res = *x ^ *c;
    orc_rng_step(x, c);
    return res;
}
/* rng/cl/random.cl:92-95 */
static float random_01(uint32_t* x, uint32_t* c) { return MWC64X_NextUint(x, c) / 4294967295.0f; }

/* rng/mwc64xseedgenerator.cpp:56-64: srand(seed); buffer[i].x = rand().  (.y is left
 * uninitialised by the reference and overwritten by the kernel; we write 0.) */
void orc_rng_host_base_offsets(uint32_t seed, uint32_t* state, size_t n) {
    srand(seed);
    for (size_t i = 0; i < n; ++i) {
        state[2 * i] = (uint32_t)rand();
        state[2 * i + 1] = 0;
    }
}

/* rng/cl/randstategen.cl:39-47 with MWC_SeedImpl_Mod64 (rng/cl/skip_mwc.cl:91-105),
 * vecSize = 1, vecOffset = 0; global id = first_stream + i. */
void orc_rng_seed_streams(uint32_t* state, size_t n, uint64_t gap, uint64_t first_stream) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) {
        uint64_t streamBase = state[2 * i];
        uint64_t dist = streamBase + (first_stream + (uint64_t)i) * gap;
        uint64_t m = MWC_PowMod64(MWC64X_A, dist, MWC64X_M);
        uint64_t x = MWC_MulMod64(MWC_BASEID, m, MWC64X_M);
        state[2 * i] = (uint32_t)(x / MWC64X_A);
        state[2 * i + 1] = (uint32_t)(x % MWC64X_A);
    }
}

/* rng/cl/randomnumbergenerator.cl:34-49 */
void orc_rng_uniform(uint32_t* state, size_t n, int per_stream, float* out) {
    for (size_t i = 0; i < n; ++i) {
        uint32_t x = state[2 * i], c = state[2 * i + 1];
        for (int k = 0; k < per_stream; ++k) out[i * (size_t)per_stream + k] = random_01(&x, &c);
        state[2 * i] = x;
        state[2 * i + 1] = c;
    }
}
