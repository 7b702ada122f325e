// radixsort.cu -- hand-written onesweep LSD radix sort (north-star subsystem 5), no CUB.
//
// Replaces clogs::Radixsort::enqueue (rsc/ext/clogs/src/radixsort.cpp:169-259) and its three
// kernels radixsortReduce / radixsortScan / radixsortScatter (clogs/kernels/radixsort.cl:248,
// 323, 893).  Same contract (clogs/radixsort.h:227-229): ascending, STABLE, keys (and values)
// sorted in place, `max_bits` bounds the significant key bits (0 = all 32).
//
// Traffic: clogs reads the keys twice and moves key+value once per 4-bit pass: 8 x (4+8+8) =
// 160 B per (u32,u32) pair.  Onesweep with 8-bit digits reads the keys once for all digit
// histograms and then moves each element exactly once per pass with a single-pass chained scan
// (decoupled look-back): 4 + 4 x 2 x (4+4) = 68 B per pair, 36 B per key for keys only.
//
// Kernel structure per pass (one CTA = one tile of SORT_TILE = 8192 elements, tiles taken in order from
// an atomic ticket so that look-back never waits on a CTA that has not started):
//   1. warp-striped coalesced load of ITEMS keys (and values) per thread
//   2. EARLY COUNTS: per-warp digit histograms with shared-memory atomics (order-free), summed per
//      digit by thread d -> the tile's count for d is published (FLAG_AGGREGATE) before any ranking,
//      and the first window of predecessor status words is already requested, so that the chained
//      scan's L2 round trips overlap the ranking below
//   3. stable ranking inside each warp: peers = lanes with the same digit, found with 8 ballots (one
//      per digit bit; MATCH.ANY serialises over the distinct values of a warp and was 55 % of the
//      round-1 kernel time); running per-(warp,digit) offsets in shared memory start at the final
//      tile-local position (known from step 2), so every element goes straight to its slot in the
//      tile-sorted staging buffer -- no rank registers
//   4. decoupled look-back over windows of 4 predecessors (4 independent loads in flight per digit),
//      publish FLAG_PREFIX
//   5. coalesced write-out in runs of equal digit
#include "common.cuh"

namespace {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int RADIX_THREADS_MIN = RADIX;
// tuning knobs (tools/sort_variants.sh sweeps them; the defaults are the round-1 winners on B200)
#ifndef CPM_SORT_THREADS
#define CPM_SORT_THREADS 256
#endif
#ifndef CPM_SORT_ITEMS
#define CPM_SORT_ITEMS 24
#endif
#ifndef CPM_SORT_MIN_BLOCKS
#define CPM_SORT_MIN_BLOCKS 3
#endif
#ifndef CPM_SORT_LOOKBACK
#define CPM_SORT_LOOKBACK 4
#endif
constexpr int SORT_THREADS = CPM_SORT_THREADS;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS = CPM_SORT_ITEMS;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;
constexpr int MAX_PASSES = 4;
constexpr int LOOKBACK_WINDOW = CPM_SORT_LOOKBACK;
static_assert(SORT_THREADS >= RADIX_THREADS_MIN && SORT_THREADS % 32 == 0, "one thread per digit is needed");
static_assert((SORT_WARPS * 256) % SORT_THREADS == 0, "counter clear loop");

constexpr uint32_t FLAG_AGGREGATE = 1u << 30;
constexpr uint32_t FLAG_PREFIX = 2u << 30;
constexpr uint32_t FLAG_MASK = 3u << 30;
constexpr uint32_t VALUE_MASK = ~FLAG_MASK;

// ---- pass 0: all digit histograms in one read of the keys -----------------------------------
// 4 B/key of HBM traffic.  Shared-memory atomics; a warp whose 32 keys share a digit (the
// importance keys are ~90 % 0x7FFFFFFF) adds 32 once instead of serialising on one bank.
__global__ void __launch_bounds__(512) histogram_kernel(const uint32_t* __restrict__ keys, size_t n, int passes,
                                                        uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_hist[MAX_PASSES * RADIX];
    for (int i = threadIdx.x; i < MAX_PASSES * RADIX; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const size_t n4 = n / 4;
    const uint4* k4 = reinterpret_cast<const uint4*>(keys);
    const unsigned lane = threadIdx.x & 31;
    auto add = [&](uint32_t k) {
        // all lanes of the warp are active here (uniform loop trip count per warp is enforced below)
        uint32_t k0 = __shfl_sync(0xffffffffu, k, 0);
        bool same = __all_sync(0xffffffffu, k == k0);
        if (same) {
            if (lane < (unsigned)passes) atomicAdd(&s_hist[lane * RADIX + ((k0 >> (lane * RADIX_BITS)) & (RADIX - 1))], 32u);
        } else {
#pragma unroll
            for (int p = 0; p < MAX_PASSES; ++p)
                if (p < passes) atomicAdd(&s_hist[p * RADIX + ((k >> (p * RADIX_BITS)) & (RADIX - 1))], 1u);
        }
    };
    // warp-uniform trip count: iterate over whole-warp chunks, tail handled separately
    size_t warp_global = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
    size_t n_warps = ((size_t)gridDim.x * blockDim.x) / 32;
    size_t full_chunks = n4 / 32;  // chunks of 32 uint4
    for (size_t c = warp_global; c < full_chunks; c += n_warps) {
        uint4 v = k4[c * 32 + lane];
        add(v.x); add(v.y); add(v.z); add(v.w);
    }
    // tail: remaining keys one by one (fewer than 128 + 3), done by block 0
    if (blockIdx.x == 0) {
        for (size_t i = full_chunks * 128 + threadIdx.x; i < n; i += blockDim.x) {
            uint32_t k = keys[i];
            for (int p = 0; p < passes; ++p) atomicAdd(&s_hist[p * RADIX + ((k >> (p * RADIX_BITS)) & (RADIX - 1))], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RADIX; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
}

// one block per pass: digit histogram -> exclusive prefix (the global base of every digit), in place;
// trivial[pass] = 1 when every key has the same digit in this pass (the pass is then the identity
// permutation: the importance keys share their two top bytes, photon ids their top byte)
__global__ void __launch_bounds__(RADIX) hist_scan_kernel(uint32_t* __restrict__ hist, uint32_t n, uint32_t* __restrict__ trivial) {
    __shared__ uint32_t s_warp[RADIX / 32];
    uint32_t* h = hist + blockIdx.x * RADIX;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t v = h[tid], incl = v;
    if (v == n) trivial[blockIdx.x] = 1u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += u;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t off = 0;
#pragma unroll
    for (int w = 0; w < RADIX / 32; ++w)
        if (w < (int)warp) off += s_warp[w];
    h[tid] = off + incl - v;
}

template <bool HAS_VALUES>
struct SortSmem {
    uint32_t warp_count[SORT_WARPS][RADIX];  // per-warp digit counts, then running tile-local offsets
    uint32_t gbase[RADIX];                   // global index of tile-local position 0 of each digit run
    uint32_t keys[SORT_TILE];
    uint32_t vals[HAS_VALUES ? SORT_TILE : 1];
    uint32_t scan_tmp[RADIX / 32];
    uint32_t tile;
};

// lanes of the warp whose digit equals this lane's: one ballot per digit bit (full rate, fixed cost)
__device__ __forceinline__ uint32_t match_digit(uint32_t d) {
    uint32_t peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < RADIX_BITS; ++b) {
        bool bit = (d >> b) & 1u;
        uint32_t m = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? m : ~m;
    }
    return peers;
}

template <bool HAS_VALUES>
__global__ void __launch_bounds__(SORT_THREADS, CPM_SORT_MIN_BLOCKS) onesweep_kernel(const uint32_t* __restrict__ keys_in,
                                                                   const uint32_t* __restrict__ vals_in,
                                                                   uint32_t* __restrict__ keys_out,
                                                                   uint32_t* __restrict__ vals_out, size_t n, int shift,
                                                                   const uint32_t* __restrict__ digit_base,  // this pass
                                                                   volatile uint32_t* status,                // [tiles][RADIX]
                                                                   uint32_t* ticket, const uint32_t* __restrict__ trivial) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SortSmem<HAS_VALUES>& S = *reinterpret_cast<SortSmem<HAS_VALUES>*>(smem_raw);
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (*trivial) {
        // every key has the same digit: a stable pass leaves the order unchanged -- straight tile copy
        // (the ping-pong parity of the passes is kept), no ranking, no chained scan
        const size_t base = (size_t)blockIdx.x * SORT_TILE;
        const uint32_t nv = (uint32_t)min((size_t)SORT_TILE, n - base);
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; ++j) {
            uint32_t i = j * SORT_THREADS + tid;
            if (i < nv) {
                keys_out[base + i] = keys_in[base + i];
                if (HAS_VALUES) vals_out[base + i] = vals_in[base + i];
            }
        }
        return;
    }

    if (tid == 0) S.tile = atomicAdd(ticket, 1u);
    {
        uint32_t* wc0 = &S.warp_count[0][0];
#pragma unroll
        for (int i = 0; i < SORT_WARPS * RADIX / SORT_THREADS; ++i) wc0[i * SORT_THREADS + tid] = 0;
    }
    __syncthreads();
    const uint32_t tile = S.tile;
    const size_t tile_base = (size_t)tile * SORT_TILE;
    const uint32_t n_valid = (uint32_t)min((size_t)SORT_TILE, n - tile_base);

    // 1. load (warp-striped): warp w owns [w*32*ITEMS, (w+1)*32*ITEMS) of the tile
    uint32_t key[SORT_ITEMS], val[HAS_VALUES ? SORT_ITEMS : 1];
    const uint32_t warp_base = warp * 32 * SORT_ITEMS;
    if (n_valid == SORT_TILE) {
        const uint32_t* kp = keys_in + tile_base + warp_base + lane;
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; ++j) key[j] = kp[j * 32];
        if (HAS_VALUES) {
            const uint32_t* vp = vals_in + tile_base + warp_base + lane;
#pragma unroll
            for (int j = 0; j < SORT_ITEMS; ++j) val[j] = vp[j * 32];
        }
    } else {
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; ++j) {
            uint32_t local = warp_base + j * 32 + lane;
            bool ok = local < n_valid;
            key[j] = ok ? keys_in[tile_base + local] : 0xffffffffu;   // padding: last digit, after every real key
            if (HAS_VALUES) val[j] = ok ? vals_in[tile_base + local] : 0u;
        }
    }

    // 2. early counts
    uint32_t* wc = S.warp_count[warp];
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) atomicAdd(&wc[(key[j] >> shift) & (RADIX - 1)], 1u);
    __syncthreads();

    uint32_t total = 0, digit_start = 0, exclusive = 0;
    uint32_t early[LOOKBACK_WINDOW];
    if (tid < RADIX) {
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) total += S.warp_count[w][tid];
        uint32_t real_total = total;
        if (tid == RADIX - 1) real_total -= (SORT_TILE - n_valid);
        status[(size_t)tile * RADIX + tid] = (tile == 0 ? FLAG_PREFIX : FLAG_AGGREGATE) | real_total;
        // request the first look-back window now; it is consumed after the ranking
#pragma unroll
        for (int i = 0; i < LOOKBACK_WINDOW; ++i)
            early[i] = ((int64_t)tile - 1 - i >= 0) ? status[((size_t)tile - 1 - i) * RADIX + tid] : 0u;
        // tile-local start of each digit run: exclusive scan of `total` over the 256 digits
        uint32_t incl = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += v;
        }
        if (lane == 31) S.scan_tmp[warp] = incl;
        digit_start = incl - total;
    }
    __syncthreads();
    if (tid < RADIX) {
#pragma unroll
        for (int w = 0; w < RADIX / 32; ++w)
            if (w < (int)warp) digit_start += S.scan_tmp[w];
        // per-warp running offsets start at the final tile-local position
        uint32_t run = digit_start;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            uint32_t c = S.warp_count[w][tid];
            S.warp_count[w][tid] = run;
            run += c;
        }
    }
    __syncthreads();

    // 3. stable in-warp ranking, elements go straight to their tile-sorted slot
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        uint32_t d = (key[j] >> shift) & (RADIX - 1);
        uint32_t peers = match_digit(d);
        int leader = __ffs(peers) - 1;
        uint32_t before = 0;
        if ((int)lane == leader) {
            before = wc[d];
            wc[d] = before + __popc(peers);
        }
        before = __shfl_sync(0xffffffffu, before, leader);
        uint32_t pos = before + __popc(peers & lt_mask);
        S.keys[pos] = key[j];
        if (HAS_VALUES) S.vals[pos] = val[j];
        __syncwarp();
    }

    // 4. decoupled look-back, windows of LOOKBACK_WINDOW predecessors (value and flag share one word)
    if (tid < RADIX) {
        if (tile != 0) {
            int64_t t = (int64_t)tile - 1;
            uint32_t win[LOOKBACK_WINDOW];
#pragma unroll
            for (int i = 0; i < LOOKBACK_WINDOW; ++i) win[i] = early[i];
            while (true) {
                int used = 0;
                bool done = false;
#pragma unroll
                for (int i = 0; i < LOOKBACK_WINDOW; ++i) {
                    uint32_t f = win[i] & FLAG_MASK;
                    if (f == 0u) break;                      // not published yet: poll again from here
                    exclusive += win[i] & VALUE_MASK;
                    ++used;
                    if (f == FLAG_PREFIX) {                  // tile 0 always publishes a prefix
                        done = true;
                        break;
                    }
                }
                if (done) break;
                t -= used;
#pragma unroll
                for (int i = 0; i < LOOKBACK_WINDOW; ++i) win[i] = (t - i >= 0) ? status[(size_t)(t - i) * RADIX + tid] : 0u;
            }
            uint32_t real_total = total;
            if (tid == RADIX - 1) real_total -= (SORT_TILE - n_valid);
            status[(size_t)tile * RADIX + tid] = FLAG_PREFIX | (exclusive + real_total);
        }
        // tile-local position p of digit d goes to global index gbase[d] + p
        S.gbase[tid] = digit_base[tid] + exclusive - digit_start;
    }
    __syncthreads();

    // 5. coalesced stores in runs of equal digit
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        uint32_t p = j * SORT_THREADS + tid;
        if (p < n_valid) {
            uint32_t k = S.keys[p];
            size_t g = (size_t)S.gbase[(k >> shift) & (RADIX - 1)] + p;
            keys_out[g] = k;
            if (HAS_VALUES) vals_out[g] = S.vals[p];
        }
    }
}

// ---- small stream kernels of the selection stage ----------------------------------------------
__global__ void __launch_bounds__(256) threshold_kernel(const uint32_t* __restrict__ data, uint32_t threshold, size_t n,
                                                        uint32_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = data[i] < threshold ? 1u : 0u;
}
__global__ void __launch_bounds__(256) iota_kernel(uint32_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)i;
}
// sum of ints (clogs Reduce with TYPE_INT) / fused count(data < threshold) + optional iota
template <bool COUNT_BELOW>
__global__ void __launch_bounds__(256) reduce_kernel(const uint32_t* __restrict__ data, uint32_t threshold, size_t n,
                                                     uint32_t* __restrict__ iota_out, unsigned long long* __restrict__ result) {
    unsigned long long acc = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t v = data[i];
        if (COUNT_BELOW) {
            acc += v < threshold ? 1u : 0u;
            if (iota_out) iota_out[i] = (uint32_t)i;
        } else {
            acc += (unsigned long long)(long long)(int32_t)v;
        }
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ unsigned long long s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; ++w) t += s[w];
        if (t) atomicAdd(result, t);
    }
}

// ---- stable selection: ids of the elements below a threshold, ascending ---------------------------------
// One pass: a tile (32 warps x 512 consecutive elements) ranks its hits with ballots (16 coalesced loads per
// lane, order = element order), warp 0 chains the tile totals through a decoupled look-back over status
// words (2 flag bits + 30 value bits), then every warp writes its ids to a contiguous range.
constexpr int SEL_ROUNDS = 16, SEL_WARPS = 32, SEL_TILE = SEL_ROUNDS * 32 * SEL_WARPS;   // 16 Ki elements per tile
constexpr uint32_t SEL_AGG = 1u << 30, SEL_INC = 2u << 30, SEL_VAL = (1u << 30) - 1u;

__global__ void __launch_bounds__(SEL_WARPS * 32) select_kernel(const uint32_t* __restrict__ data, size_t n, uint32_t threshold,
                                                                uint32_t* __restrict__ out, volatile uint32_t* status,
                                                                uint32_t* ticket, unsigned long long* __restrict__ result,
                                                                unsigned n_tiles) {
    __shared__ uint32_t s_tile, s_base, s_warp[SEL_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const size_t base = (size_t)tile * SEL_TILE + (size_t)warp * (SEL_ROUNDS * 32);
    uint32_t masks[SEL_ROUNDS];
    uint32_t wtotal = 0;
#pragma unroll
    for (int r = 0; r < SEL_ROUNDS; ++r) {
        size_t i = base + r * 32 + lane;
        bool hit = i < n && data[i] < threshold;
        masks[r] = __ballot_sync(0xffffffffu, hit);
        wtotal += __popc(masks[r]);
    }
    if (lane == 0) s_warp[warp] = wtotal;
    __syncthreads();
    if (warp == 0) {
        // Warp 0 chains the tile totals.  The look-back reads a WINDOW of 32 predecessors at a time: a tile is done as
        // soon as every predecessor up to the nearest inclusive prefix has published at least its aggregate -- which
        // each does right after counting, so no tile waits for a chain of prefixes (with one thread and one predecessor
        // per read, 1024 tiles of 4 Ki elements took 23 us for 16 MB of keys: 22 ns per link).
        uint32_t v = lane < SEL_WARPS ? s_warp[lane] : 0u;
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const uint32_t total = v;
        uint32_t excl = 0;
        if (tile > 0) {
            if (lane == 0) {
                status[tile] = SEL_AGG | total;
                __threadfence();
            }
            int p = (int)tile - 1;   // the window is tiles p - 31 ... p, lane l looks at p - l
            while (true) {
                const int q = p - lane;
                uint32_t st = 2u << 30;                              // "before tile 0": inclusive prefix 0 (SEL_INC | 0)
                if (q >= 0) st = status[q];
                const unsigned ready = __ballot_sync(0xffffffffu, (st & (SEL_AGG | SEL_INC)) != 0u);
                const unsigned inc = __ballot_sync(0xffffffffu, (st & SEL_INC) != 0u);
                if (inc) {
                    const int f = __ffs(inc) - 1;                    // nearest inclusive prefix
                    const unsigned need = f == 31 ? 0xffffffffu : ((2u << f) - 1u);
                    if ((ready & need) != need) continue;            // someone nearer has not published yet: poll again
                    uint32_t c = lane <= f ? (st & SEL_VAL) : 0u;
#pragma unroll
                    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                    excl += c;
                    break;
                }
                if (ready != 0xffffffffu) continue;
                uint32_t c = st & SEL_VAL;                           // 32 aggregates: take them all, look further back
#pragma unroll
                for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                excl += c;
                p -= 32;
            }
        }
        if (lane == 0) {
            status[tile] = SEL_INC | (excl + total);
            __threadfence();
            s_base = excl;
            if (tile == n_tiles - 1) {
                *result = (unsigned long long)(excl + total);   // device scratch, or the mapped pinned host word
                __threadfence_system();
            }
        }
    }
    __syncthreads();
    uint32_t wbase = s_base;
    for (int w = 0; w < warp; ++w) wbase += s_warp[w];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < SEL_ROUNDS; ++r) {
        uint32_t m = masks[r];
        if ((m >> lane) & 1u) out[wbase + __popc(m & lt)] = (uint32_t)(base + r * 32 + lane);
        wbase += __popc(m);
    }
}

}  // namespace

extern "C" {

size_t cpm_radix_sort_scratch_bytes(size_t n) {
    size_t tiles = (n + SORT_TILE - 1) / SORT_TILE;
    return 4096 /*hist*/ + 256 /*tickets*/ + (size_t)MAX_PASSES * tiles * RADIX * sizeof(uint32_t);
}

int cpm_radix_sort_u32(cpm_ctx* ctx, uint32_t* keys, uint32_t* values, size_t n, unsigned max_bits, uint32_t* tmp_keys,
                       uint32_t* tmp_values) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, n > 0, "elements is zero");                       // clogs: CL_INVALID_GLOBAL_WORK_SIZE
    CPM_REQUIRE(ctx, max_bits <= 32, "maxBits is too large");          // clogs: CL_INVALID_VALUE
    CPM_REQUIRE(ctx, keys && tmp_keys, "keys / temporary key buffer is NULL");
    CPM_REQUIRE(ctx, !values || tmp_values, "temporary value buffer is NULL");
    if (n >= (1ull << 30)) return cpm_fail(ctx, CPM_E_UNSUPPORTED, "cpm_radix_sort_u32: n must be < 2^30");
    if (max_bits == 0) max_bits = 32;
    const int passes = (int)((max_bits + RADIX_BITS - 1) / RADIX_BITS);
    const size_t tiles = (n + SORT_TILE - 1) / SORT_TILE;
    void* scratch;
    size_t status_bytes = (size_t)passes * tiles * RADIX * sizeof(uint32_t);
    int rc = cpm_scratch(ctx, 4096 + 256 + status_bytes, &scratch);
    if (rc != CPM_OK) return rc;
    uint32_t* hist = (uint32_t*)scratch;
    uint32_t* tickets = (uint32_t*)((char*)scratch + 4096);
    uint32_t* status = (uint32_t*)((char*)scratch + 4096 + 256);
    CPM_CUDA(ctx, cudaMemsetAsync(scratch, 0, 4096 + 256 + status_bytes, ctx->stream));

    unsigned hgrid = (unsigned)std::min<size_t>((size_t)ctx->sm_count * 4, std::max<size_t>(1, n / (512 * 4)));
    CPM_LAUNCH(ctx, histogram_kernel, hgrid, 512, 0, keys, n, passes, hist);

    uint32_t* trivial = tickets + 16;   // [MAX_PASSES], zeroed with the tickets
    CPM_LAUNCH(ctx, hist_scan_kernel, passes, RADIX, 0, hist, (uint32_t)n, trivial);

    static bool attr_set = false;
    if (!attr_set) {
        CPM_CUDA(ctx, cudaFuncSetAttribute(onesweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem<true>)));
        CPM_CUDA(ctx, cudaFuncSetAttribute(onesweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem<false>)));
        attr_set = true;
    }
    uint32_t *kin = keys, *kout = tmp_keys, *vin = values, *vout = tmp_values;
    for (int p = 0; p < passes; ++p) {
        if (values) {
            CPM_LAUNCH(ctx, onesweep_kernel<true>, (unsigned)tiles, SORT_THREADS, sizeof(SortSmem<true>), kin, vin, kout, vout, n,
                       p * RADIX_BITS, hist + p * RADIX, status + (size_t)p * tiles * RADIX, tickets + p, trivial + p);
        } else {
            CPM_LAUNCH(ctx, onesweep_kernel<false>, (unsigned)tiles, SORT_THREADS, sizeof(SortSmem<false>), kin, nullptr, kout, nullptr,
                       n, p * RADIX_BITS, hist + p * RADIX, status + (size_t)p * tiles * RADIX, tickets + p, trivial + p);
        }
        std::swap(kin, kout);
        std::swap(vin, vout);
    }
    if (kin != keys) {  // odd number of passes: copy back, as clogs does (radixsort.cpp:241-256)
        CPM_CUDA(ctx, cudaMemcpyAsync(keys, kin, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
        if (values) CPM_CUDA(ctx, cudaMemcpyAsync(values, vin, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return CPM_OK;
}

int cpm_threshold_u32(cpm_ctx* ctx, const uint32_t* data, uint32_t threshold, size_t n, uint32_t* out) {
    if (!ctx) return CPM_E_INVALID;
    if (n == 0) return CPM_OK;
    CPM_REQUIRE(ctx, data && out, "null buffer");
    CPM_LAUNCH(ctx, threshold_kernel, cpm_div_up(n, 256), 256, 0, data, threshold, n, out);
    return CPM_OK;
}

int cpm_iota_u32(cpm_ctx* ctx, uint32_t* out, size_t n) {
    if (!ctx) return CPM_E_INVALID;
    if (n == 0) return CPM_OK;
    CPM_REQUIRE(ctx, out != nullptr, "null buffer");
    CPM_LAUNCH(ctx, iota_kernel, cpm_div_up(n, 256), 256, 0, out, n);
    return CPM_OK;
}

static int reduce_common(cpm_ctx* ctx, bool count_below, const uint32_t* data, uint32_t threshold, size_t n,
                         uint32_t* iota_out, long long* result_host) {
    CPM_REQUIRE(ctx, result_host != nullptr, "result pointer is NULL");
    *result_host = 0;
    if (n == 0) return CPM_OK;
    CPM_REQUIRE(ctx, data != nullptr, "null buffer");
    void* scratch;
    int rc = cpm_scratch(ctx, 4096, &scratch);
    if (rc != CPM_OK) return rc;
    unsigned long long* acc = (unsigned long long*)((char*)scratch + 2048);  // away from the sort's histogram words
    CPM_CUDA(ctx, cudaMemsetAsync(acc, 0, sizeof(*acc), ctx->stream));
    unsigned grid = (unsigned)std::min<size_t>((size_t)ctx->sm_count * 8, cpm_div_up(n, 256));
    if (count_below)
        CPM_LAUNCH(ctx, reduce_kernel<true>, grid, 256, 0, data, threshold, n, iota_out, acc);
    else
        CPM_LAUNCH(ctx, reduce_kernel<false>, grid, 256, 0, data, threshold, n, nullptr, acc);
    unsigned long long* pinned = (unsigned long long*)ctx->pinned;
    CPM_CUDA(ctx, cudaMemcpyAsync(pinned, acc, sizeof(*acc), cudaMemcpyDeviceToHost, ctx->stream));
    CPM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *result_host = (long long)*pinned;
    return CPM_OK;
}

int cpm_select_below_begin(cpm_ctx* ctx, const uint32_t* data, size_t n, uint32_t threshold, uint32_t* ids_out) {
    if (!ctx) return CPM_E_INVALID;
    ctx->select_pending = false;
    if (n == 0) return CPM_OK;
    CPM_REQUIRE(ctx, data && ids_out, "null buffer");
    if (n >= (1ull << 30)) return cpm_fail(ctx, CPM_E_UNSUPPORTED, "cpm_select_below: n must be < 2^30");
    const size_t tiles = (n + SEL_TILE - 1) / SEL_TILE;
    void* scratch;
    int rc = cpm_scratch(ctx, 4096 + tiles * sizeof(uint32_t), &scratch);
    if (rc != CPM_OK) return rc;
    unsigned long long* acc = (unsigned long long*)((char*)scratch + 2048);
    uint32_t* ticket = (uint32_t*)((char*)scratch + 2048 + 8);
    uint32_t* status = (uint32_t*)((char*)scratch + 4096);
    CPM_CUDA(ctx, cudaMemsetAsync((char*)scratch + 2048, 0, 2048 + tiles * sizeof(uint32_t), ctx->stream));
    // The count goes straight into the context's pinned host word: the last tile stores it through the mapped
    // address.  A cudaMemcpyAsync of 8 bytes would queue on the device-to-host copy engine -- behind a 64 MB read-back
    // of the previous frame's light volume when the caller streams results out (1 ms of the frame at N = 8).
    unsigned long long* pinned_dev = nullptr;
    if (cudaHostGetDevicePointer((void**)&pinned_dev, ctx->pinned, 0) != cudaSuccess) {
        (void)cudaGetLastError();
        pinned_dev = nullptr;
    }
    CPM_LAUNCH(ctx, select_kernel, (unsigned)tiles, SEL_WARPS * 32, 0, data, n, threshold, ids_out, status, ticket,
               pinned_dev ? pinned_dev : acc, (unsigned)tiles);
    if (!pinned_dev) {
        unsigned long long* pinned = (unsigned long long*)ctx->pinned;
        CPM_CUDA(ctx, cudaMemcpyAsync(pinned, acc, sizeof(*acc), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (!ctx->select_done) CPM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->select_done, cudaEventDisableTiming));
    CPM_CUDA(ctx, cudaEventRecord(ctx->select_done, ctx->stream));
    ctx->select_pending = true;
    return CPM_OK;
}

int cpm_select_below_end(cpm_ctx* ctx, long long* count_host) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, count_host != nullptr, "result pointer is NULL");
    *count_host = 0;
    if (!ctx->select_pending) return CPM_OK;   // n == 0
    ctx->select_pending = false;
    CPM_CUDA(ctx, cudaEventSynchronize(ctx->select_done));
    *count_host = (long long)*(unsigned long long*)ctx->pinned;
    return CPM_OK;
}

int cpm_select_below(cpm_ctx* ctx, const uint32_t* data, size_t n, uint32_t threshold, uint32_t* ids_out,
                     long long* count_host) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, count_host != nullptr, "result pointer is NULL");
    *count_host = 0;
    int rc = cpm_select_below_begin(ctx, data, n, threshold, ids_out);
    if (rc != CPM_OK) return rc;
    return cpm_select_below_end(ctx, count_host);
}

int cpm_reduce_sum_i32(cpm_ctx* ctx, const int32_t* data, size_t n, long long* result_host) {
    if (!ctx) return CPM_E_INVALID;
    return reduce_common(ctx, false, (const uint32_t*)data, 0, n, nullptr, result_host);
}

int cpm_count_below(cpm_ctx* ctx, const uint32_t* data, size_t n, uint32_t threshold, uint32_t* iota_out,
                    long long* count_host) {
    if (!ctx) return CPM_E_INVALID;
    return reduce_common(ctx, true, data, threshold, n, iota_out, count_host);
}

}  // extern "C"
