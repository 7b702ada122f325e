/* orc_sort.c -- oracle restatement of the selection stage.  TEST INFRASTRUCTURE.
 *
 * clogs promises a stable ascending sort (rsc/ext/clogs/radixsort.h:227-229); a stable sort has
 * exactly one result, so any stable algorithm is a valid oracle.  Two are provided:
 *   orc_radix_sort_u32  -- LSD radix sort with clogs' pass structure (4-bit digits, ping-pong,
 *                          copy back after an odd pass count; src/radixsort.cpp:229-256).  Also
 *                          the timed CPU baseline.
 *   orc_merge_sort_u32  -- top-down merge sort (what std::stable_sort does), independent check.
 */
#include <stdlib.h>
#include <string.h>

#include "cpm_oracle.h"

void orc_radix_sort_u32(uint32_t* keys, uint32_t* values, size_t n, unsigned max_bits) {
    const unsigned radixBits = 4, radix = 16;
    if (max_bits == 0) max_bits = 32;
    uint32_t* tk = (uint32_t*)malloc(n * sizeof(uint32_t));
    uint32_t* tv = values ? (uint32_t*)malloc(n * sizeof(uint32_t)) : NULL;
    uint32_t *ck = keys, *nk = tk, *cv = values, *nv = tv;
    for (unsigned firstBit = 0; firstBit < max_bits; firstBit += radixBits) {
        size_t hist[16] = {0};
        for (size_t i = 0; i < n; ++i) hist[(ck[i] >> firstBit) & (radix - 1)]++;   /* radixsortReduce */
        size_t sum = 0;
        for (unsigned d = 0; d < radix; ++d) { size_t c = hist[d]; hist[d] = sum; sum += c; } /* radixsortScan */
        for (size_t i = 0; i < n; ++i) {                                            /* radixsortScatter */
            size_t p = hist[(ck[i] >> firstBit) & (radix - 1)]++;
            nk[p] = ck[i];
            if (values) nv[p] = cv[i];
        }
        uint32_t* t = ck; ck = nk; nk = t;
        t = cv; cv = nv; nv = t;
    }
    if (ck != keys) {
        memcpy(keys, ck, n * sizeof(uint32_t));
        if (values) memcpy(values, cv, n * sizeof(uint32_t));
    }
    free(tk);
    free(tv);
}

static void merge_rec(uint32_t* k, uint32_t* v, uint32_t* tk, uint32_t* tv, size_t lo, size_t hi) {
    if (hi - lo < 2) return;
    size_t mid = lo + (hi - lo) / 2;
    merge_rec(k, v, tk, tv, lo, mid);
    merge_rec(k, v, tk, tv, mid, hi);
    size_t i = lo, j = mid, o = lo;
    while (i < mid && j < hi) {
        if (k[j] < k[i]) { tk[o] = k[j]; if (v) tv[o] = v[j]; ++j; }
        else { tk[o] = k[i]; if (v) tv[o] = v[i]; ++i; }
        ++o;
    }
    while (i < mid) { tk[o] = k[i]; if (v) tv[o] = v[i]; ++i; ++o; }
    while (j < hi) { tk[o] = k[j]; if (v) tv[o] = v[j]; ++j; ++o; }
    memcpy(k + lo, tk + lo, (hi - lo) * sizeof(uint32_t));
    if (v) memcpy(v + lo, tv + lo, (hi - lo) * sizeof(uint32_t));
}

void orc_merge_sort_u32(uint32_t* keys, uint32_t* values, size_t n) {
    uint32_t* tk = (uint32_t*)malloc(n * sizeof(uint32_t));
    uint32_t* tv = values ? (uint32_t*)malloc(n * sizeof(uint32_t)) : NULL;
    merge_rec(keys, values, tk, tv, 0, n);
    free(tk);
    free(tv);
}

/* ppm/cl/threshold.cl:39 + clogs reduce: count of data[i] < threshold */
long long orc_count_below(const uint32_t* data, size_t n, uint32_t threshold) {
    long long c = 0;
    for (size_t i = 0; i < n; ++i) c += data[i] < threshold;
    return c;
}
