// bound.cu -- per-cell opacity bound for the delta-tracking tracer ("L2-resident bricks", north-star 3).
//
// Not in the reference: woodcockTracking (ppm/cl/transmittance.cl:126-144) fetches the volume and the
// transfer function for EVERY collision test.  The accept/reject decision of a test is `u2 >= opacity`; if an
// upper bound m >= opacity is known for the cell the sample falls in, `u2 >= m` already decides "reject"
// and the 8 voxel taps + 2 TF taps are never fetched.  The random stream, the sample positions and every
// stored photon stay exactly those of the reference loop -- the bound only removes memory traffic.
//
//  cpm_volume_value_range   (lo, hi) of the normalised voxel values that a trilinear footprint can touch, per
//                           cell.  Cell c = floor((u + 1) / cell) per axis, u = p * dim - 0.5 the continuous
//                           voxel coordinate of the sample; its footprints have lower tap i0 in
//                           [c*cell - 1, c*cell + cell - 2].  The grid covers one voxel more on either side,
//                           voxels [c*cell - 2, c*cell + cell] clamped to the volume, so that the tracer may
//                           compute c with its own, differently rounded arithmetic (errors << 1 voxel).
//                           One HBM pass over the linear buffer, overlap rows served by L2.
//  cpm_opacity_bound        bound[c] >= alpha(TF((v + offset) * scale)) for every v in [lo, hi], widened by
//                           the worst-case rounding of the fp32 blend (sampling.cuh blend_taps /
//                           sample_tf_alpha); +inf ("always fetch") where a NaN/inf voxel or NaN alpha makes
//                           the arithmetic non-monotone.
#include "sampling.cuh"

namespace {

__device__ __forceinline__ uint32_t okey(float f) {  // monotone float -> uint
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float oval(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

template <int FMT>
struct VoxT;
template <>
struct VoxT<CPM_FMT_U8> {
    typedef unsigned char T;
    static constexpr int PER16 = 16;
    __device__ static float norm(T v) { return unorm8((float)v); }
};
template <>
struct VoxT<CPM_FMT_U16> {
    typedef unsigned short T;
    static constexpr int PER16 = 8;
    __device__ static float norm(T v) { return unorm16((float)v); }
};
template <>
struct VoxT<CPM_FMT_F32> {
    typedef float T;
    static constexpr int PER16 = 4;
    __device__ static float norm(T v) { return v; }
};

// one CTA per (cy, cz): all cells along x of that row.  smem: ncx min keys, ncx max keys, ncx bad flags.
template <int FMT>
__global__ void __launch_bounds__(256) range_kernel(const void* __restrict__ vol, int nx, int ny, int nz, int s, int ncx,
                                                    float2* __restrict__ out, int vec_ok) {
    typedef typename VoxT<FMT>::T T;
    constexpr int K = VoxT<FMT>::PER16;
    extern __shared__ uint32_t s_r[];
    uint32_t *s_min = s_r, *s_max = s_r + ncx, *s_bad = s_r + 2 * ncx;
    const int cy = blockIdx.x, cz = blockIdx.y, cell = 1 << s;
    for (int i = threadIdx.x; i < ncx; i += blockDim.x) {
        s_min[i] = 0xffffffffu;
        s_max[i] = 0u;
        s_bad[i] = 0u;
    }
    __syncthreads();
    const int y0 = max(cy * cell - 2, 0), y1 = min(cy * cell + cell, ny - 1);
    const int z0 = max(cz * cell - 2, 0), z1 = min(cz * cell + cell, nz - 1);
    const int ry = y1 - y0 + 1, rz = z1 - z0 + 1;
    const int rows = ry * rz;
    const T* base = (const T*)vol;
    if (vec_ok) {
        // one warp per row, lanes stride over its 16-byte chunks: no per-item index divisions
        const int chunks = nx / K;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
        for (int r = warp; r < rows; r += n_warps) {
            const int zr = r / ry;
            const int y = y0 + (r - zr * ry), z = z0 + zr;
            const uint4* row = reinterpret_cast<const uint4*>(base + ((size_t)z * ny + y) * nx);
            for (int c = lane; c < chunks; c += 32) {
                uint4 raw = __ldg(row + c);
                const T* e = reinterpret_cast<const T*>(&raw);
                const int x = c * K;
                float v[K];
                bool anybad = false;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    v[k] = VoxT<FMT>::norm(e[k]);
                    anybad = anybad || !(fabsf(v[k]) <= CPM_FLT_MAX_);
                }
                // voxel x is in cell q iff q*cell - 2 <= x <= q*cell + cell: the chunk touches a contiguous cell range
                const int qmin = max(((x + cell - 1) >> s) - 1, 0), qmax = min((x + K + 1) >> s, ncx - 1);
                for (int q = qmin; q <= qmax; ++q) {
                    const int lo = q * cell - 2 - x, hi = q * cell + cell - x;   // relative to the chunk
                    float mn = CPM_FLT_MAX_, mx = -CPM_FLT_MAX_;
                    bool bad = false;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const bool in = k >= lo && k <= hi;
                        mn = in ? fminf(mn, v[k]) : mn;
                        mx = in ? fmaxf(mx, v[k]) : mx;
                        if (anybad) bad = bad || (in && !(fabsf(v[k]) <= CPM_FLT_MAX_));
                    }
                    atomicMin(&s_min[q], okey(mn));
                    atomicMax(&s_max[q], okey(mx));
                    if (bad) s_bad[q] = 1u;
                }
            }
        }
    } else {
        const int items = rows * nx;
        for (int w = threadIdx.x; w < items; w += blockDim.x) {
            int r = w / nx, x = w - r * nx;
            int y = y0 + r % ry, z = z0 + r / ry;
            float v = VoxT<FMT>::norm(base[((size_t)z * ny + y) * nx + x]);
            bool isbad = !(fabsf(v) <= CPM_FLT_MAX_);
            const int qmin = max(((x + cell - 1) >> s) - 1, 0), qmax = min((x + 2) >> s, ncx - 1);
            for (int q = qmin; q <= qmax; ++q) {
                atomicMin(&s_min[q], okey(v));
                atomicMax(&s_max[q], okey(v));
                if (isbad) s_bad[q] = 1u;
            }
        }
    }
    __syncthreads();
    const int ncy = gridDim.x;
    for (int i = threadIdx.x; i < ncx; i += blockDim.x) {
        float2 o = make_float2(oval(s_min[i]), oval(s_max[i]));
        if (s_bad[i]) o.x = o.y = __uint_as_float(0x7fc00000u);
        out[((size_t)cz * ncy + cy) * ncx + i] = o;
    }
}

// ---- cells of 8 voxels (the default), 16-byte aligned rows: streaming version --------------------------------
// Same CTA mapping (one CTA per (cy, cz), 11 x 11 rows of the padded cell row), but a thread owns one 16-byte chunk
// COLUMN and walks down the rows: the cells a chunk feeds are then fixed per thread, min / max stay in registers
// (packed bytes / halves for the integer formats, order keys for f32 so that NaN and inf show up in the extremes),
// four independent loads are in flight per thread, and shared memory sees a handful of atomics per thread instead
// of several per chunk.  Voxel v feeds cell q iff 8q - 2 <= v <= 8q + 8.
template <int FMT>
struct ChunkAcc;

template <>
struct ChunkAcc<CPM_FMT_F32> {   // 4 voxels: x = 8m + 4j .. + 3, j = column parity
    uint32_t mn[4], mx[4];
    __device__ void init() {
#pragma unroll
        for (int k = 0; k < 4; ++k) mn[k] = 0xffffffffu, mx[k] = 0u;
    }
    __device__ void add(const uint4& r) {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t key = w[k] ^ (uint32_t)(((int32_t)w[k] >> 31) | 0x80000000);   // okey()
            mn[k] = min(mn[k], key);
            mx[k] = max(mx[k], key);
        }
    }
    // order keys: +inf / +NaN are >= key(+inf), -inf / -NaN are <= key(-inf)
    __device__ static bool bad(uint32_t kmin, uint32_t kmax) { return kmax >= 0xff800000u || kmin <= 0x007fffffu; }
    __device__ void flush(int c, int ncx, uint32_t* s_min, uint32_t* s_max, uint32_t* s_bad) const {
        const int m = c >> 1, j = c & 1;
        uint32_t a = min(min(mn[0], mn[1]), min(mn[2], mn[3])), b = max(max(mx[0], mx[1]), max(mx[2], mx[3]));
        put(m, ncx, a, b, s_min, s_max, s_bad);
        if (j == 0) {
            if (m >= 1) put(m - 1, ncx, mn[0], mx[0], s_min, s_max, s_bad);                     // voxel 8m = 8(m-1) + 8
        } else {
            put(m + 1, ncx, min(mn[2], mn[3]), max(mx[2], mx[3]), s_min, s_max, s_bad);         // voxels 8m+6, 8m+7
        }
    }
    __device__ static void put(int q, int ncx, uint32_t a, uint32_t b, uint32_t* s_min, uint32_t* s_max, uint32_t* s_bad) {
        if (q >= ncx || a > b) return;
        if (bad(a, b)) s_bad[q] = 1u;
        // the generic path reduces with fminf / fmaxf, which skip NaN: keep the extremes over the non-NaN values
        // comparable by clamping NaN keys to the infinities (the cell is flagged bad either way)
        atomicMin(&s_min[q], max(a, 0x007fffffu));
        atomicMax(&s_max[q], min(b, 0xff800000u));
    }
};

template <>
struct ChunkAcc<CPM_FMT_U16> {   // 8 voxels: x = 8m .. 8m + 7
    uint32_t mn[4], mx[4];
    __device__ void init() {
#pragma unroll
        for (int k = 0; k < 4; ++k) mn[k] = 0xffffffffu, mx[k] = 0u;
    }
    __device__ void add(const uint4& r) {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            mn[k] = __vminu2(mn[k], w[k]);
            mx[k] = __vmaxu2(mx[k], w[k]);
        }
    }
    __device__ static uint32_t lo2(uint32_t v) { return min(v & 0xffffu, v >> 16); }
    __device__ static uint32_t hi2(uint32_t v) { return max(v & 0xffffu, v >> 16); }
    __device__ static void put(int q, int ncx, uint32_t a, uint32_t b, uint32_t* s_min, uint32_t* s_max) {
        if (q < 0 || q >= ncx) return;
        atomicMin(&s_min[q], okey(unorm16((float)a)));
        atomicMax(&s_max[q], okey(unorm16((float)b)));
    }
    __device__ void flush(int c, int ncx, uint32_t* s_min, uint32_t* s_max, uint32_t*) const {
        uint32_t a = lo2(__vminu2(__vminu2(mn[0], mn[1]), __vminu2(mn[2], mn[3])));
        uint32_t b = hi2(__vmaxu2(__vmaxu2(mx[0], mx[1]), __vmaxu2(mx[2], mx[3])));
        put(c, ncx, a, b, s_min, s_max);
        put(c - 1, ncx, mn[0] & 0xffffu, mx[0] & 0xffffu, s_min, s_max);     // voxel 8m
        put(c + 1, ncx, lo2(mn[3]), hi2(mx[3]), s_min, s_max);               // voxels 8m+6, 8m+7
    }
};

template <>
struct ChunkAcc<CPM_FMT_U8> {    // 16 voxels: x = 16m .. 16m + 15 = cells 2m (bytes 0-7) and 2m + 1 (bytes 8-15)
    uint32_t mn[4], mx[4];
    __device__ void init() {
#pragma unroll
        for (int k = 0; k < 4; ++k) mn[k] = 0xffffffffu, mx[k] = 0u;
    }
    __device__ void add(const uint4& r) {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            mn[k] = __vminu4(mn[k], w[k]);
            mx[k] = __vmaxu4(mx[k], w[k]);
        }
    }
    __device__ static uint32_t lo4(uint32_t v) { v = __vminu4(v, v >> 16); return min(v & 0xffu, (v >> 8) & 0xffu); }
    __device__ static uint32_t hi4(uint32_t v) { v = __vmaxu4(v, v >> 16); return max(v & 0xffu, (v >> 8) & 0xffu); }
    __device__ static void put(int q, int ncx, uint32_t a, uint32_t b, uint32_t* s_min, uint32_t* s_max) {
        if (q < 0 || q >= ncx) return;
        atomicMin(&s_min[q], okey(unorm8((float)a)));
        atomicMax(&s_max[q], okey(unorm8((float)b)));
    }
    __device__ void flush(int c, int ncx, uint32_t* s_min, uint32_t* s_max, uint32_t*) const {
        const int q = 2 * c;
        const uint32_t b8n = mn[2] & 0xffu, b8x = mx[2] & 0xffu;                       // byte 8 = voxel 8q + 8
        const uint32_t p67n = min((mn[1] >> 16) & 0xffu, mn[1] >> 24), p67x = max((mx[1] >> 16) & 0xffu, mx[1] >> 24);
        put(q, ncx, min(lo4(__vminu4(mn[0], mn[1])), b8n), max(hi4(__vmaxu4(mx[0], mx[1])), b8x), s_min, s_max);
        put(q + 1, ncx, min(lo4(__vminu4(mn[2], mn[3])), p67n), max(hi4(__vmaxu4(mx[2], mx[3])), p67x), s_min, s_max);
        put(q - 1, ncx, mn[0] & 0xffu, mx[0] & 0xffu, s_min, s_max);                   // byte 0 = voxel 8(q-1) + 8
        put(q + 2, ncx, min((mn[3] >> 16) & 0xffu, mn[3] >> 24), max((mx[3] >> 16) & 0xffu, mx[3] >> 24), s_min, s_max);
    }
};

template <int FMT>
__global__ void __launch_bounds__(256) range8_kernel(const void* __restrict__ vol, int nx, int ny, int nz, int ncx,
                                                     float2* __restrict__ out) {
    constexpr int K = VoxT<FMT>::PER16;
    extern __shared__ uint32_t s_r[];
    uint32_t *s_min = s_r, *s_max = s_r + ncx, *s_bad = s_r + 2 * ncx;
    __shared__ unsigned long long s_row[121];
    const int cy = blockIdx.x, cz = blockIdx.y;
    for (int i = threadIdx.x; i < ncx; i += blockDim.x) {
        s_min[i] = 0xffffffffu;
        s_max[i] = 0u;
        s_bad[i] = 0u;
    }
    const int y0 = max(cy * 8 - 2, 0), y1 = min(cy * 8 + 8, ny - 1);
    const int z0 = max(cz * 8 - 2, 0), z1 = min(cz * 8 + 8, nz - 1);
    const int ry = y1 - y0 + 1, rz = z1 - z0 + 1;
    const int rows = ry * rz;
    const int chunks = nx / K;
    for (int i = threadIdx.x; i < rows; i += blockDim.x) {
        const int zr = i / ry;
        s_row[i] = ((unsigned long long)(z0 + zr) * ny + (y0 + i - zr * ry)) * chunks;
    }
    __syncthreads();
    const uint4* V = reinterpret_cast<const uint4*>(vol);
    const int cols = min(chunks, (int)blockDim.x), groups = blockDim.x / cols;
    const int col = threadIdx.x % cols, rg = threadIdx.x / cols;
    if (rg < groups) {
        for (int c = col; c < chunks; c += cols) {
            ChunkAcc<FMT> acc;
            acc.init();
            int r = rg;
            for (; r + 3 * groups < rows; r += 4 * groups) {
                uint4 a0 = __ldg(V + s_row[r] + c), a1 = __ldg(V + s_row[r + groups] + c);
                uint4 a2 = __ldg(V + s_row[r + 2 * groups] + c), a3 = __ldg(V + s_row[r + 3 * groups] + c);
                acc.add(a0); acc.add(a1); acc.add(a2); acc.add(a3);
            }
            for (; r < rows; r += groups) acc.add(__ldg(V + s_row[r] + c));
            acc.flush(c, ncx, s_min, s_max, s_bad);
        }
    }
    __syncthreads();
    const int ncy = gridDim.x;
    for (int i = threadIdx.x; i < ncx; i += blockDim.x) {
        float2 o = make_float2(oval(s_min[i]), oval(s_max[i]));
        if (s_bad[i]) o.x = o.y = __uint_as_float(0x7fc00000u);
        out[((size_t)cz * ncy + cy) * ncx + i] = o;
    }
}

// alpha column + per-32-texel summaries (max with NaN -> +inf, max |.|)
__global__ void __launch_bounds__(256) tf_summary_kernel(const float4* __restrict__ tf, int w, float* __restrict__ alpha,
                                                         float* __restrict__ bmax, float* __restrict__ babs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float a = i < w ? tf[i].w : -CPM_FLT_MAX_;
    if (i < w) alpha[i] = a;
    float m = (a != a) ? __uint_as_float(0x7f800000u) : a;
    float ab = (i < w) ? fabsf(m) : 0.0f;
    for (int off = 16; off; off >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
        ab = fmaxf(ab, __shfl_xor_sync(0xffffffffu, ab, off));
    }
    if ((threadIdx.x & 31) == 0 && i < w) {
        bmax[i >> 5] = m;
        babs[i >> 5] = ab;
    }
}

__device__ __forceinline__ int tf_index(float v, float fw) {  // i0 of sample_tf_alpha before the max(.,0)
    return (int)cpm_clamp(floorf(fmaf(v, fw, -0.5f)), -1.0f, fw - 1.0f);
}

__global__ void __launch_bounds__(256) bound_kernel(const float2* __restrict__ range, size_t n, float scale, float offset,
                                                    const float* __restrict__ alpha, const float* __restrict__ bmax,
                                                    const float* __restrict__ babs, int w, float* __restrict__ out) {
    size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    const float INF = __uint_as_float(0x7f800000u);
    float2 r = range[id];
    float res = INF;
    if (fabsf(r.x) <= CPM_FLT_MAX_ && fabsf(r.y) <= CPM_FLT_MAX_) {
        // three nested fp32 lerps of values in [lo, hi] stay within a few ulp of max(|lo|, |hi|) of the interval
        float slack = fmaf(fmaxf(fabsf(r.x), fabsf(r.y)), 0x1p-18f, 1e-37f);
        float lo = (r.x - slack + offset) * scale, hi = (r.y + slack + offset) * scale;  // blend_taps: (val + offset) * scale
        if (lo > hi) {
            float t = lo;
            lo = hi;
            hi = t;
        }
        if (fabsf(lo) <= CPM_FLT_MAX_ && fabsf(hi) <= CPM_FLT_MAX_) {
            const float fw = (float)w;
            int ia = max(tf_index(lo, fw), 0), ib = min(tf_index(hi, fw) + 1, w - 1);
            float m = -INF, ab = 0.0f;
            int i = ia;
#define ACC(v)                                  \
    {                                           \
        float v_ = (v);                         \
        m = (v_ != v_) ? INF : fmaxf(m, v_);    \
        ab = fmaxf(ab, fabsf(v_));              \
    }
            while (i <= ib && (i & 31)) {
                ACC(alpha[i]);
                ++i;
            }
            while (i + 31 <= ib) {
                m = fmaxf(m, bmax[i >> 5]);
                ab = fmaxf(ab, babs[i >> 5]);
                i += 32;
            }
            while (i <= ib) {
                ACC(alpha[i]);
                ++i;
            }
#undef ACC
            // sample_tf_alpha's lerp: within a few ulp of the larger magnitude of its two texels
            res = m + fmaf(ab, 0x1p-18f, 1e-37f);
            if (!(res == res)) res = INF;
            // every texel the cell can reach is exactly zero: the opacity is exactly zero (lerp(0, 0, a) = 0), which
            // the ray marchers use to skip the cell altogether
            if (ab == 0.0f) res = 0.0f;
        }
    }
    out[id] = res;
}

// ---- clearance of transparent cells ---------------------------------------------------------------------
// For a cell whose bound is exactly zero: R = radius, in cells, of the largest cube around it that holds only such
// cells (outside the grid counts as transparent: lookups clamp to the border cells).  A cube is the intersection of
// three slabs, so R comes from three 1-D passes: out[c] = max r <= cap with min(in[c - r .. c + r]) >= r.
template <bool FROM_BOUND, bool TO_BOUND>
__global__ void __launch_bounds__(256) clearance_kernel(const signed char* __restrict__ in, signed char* __restrict__ out,
                                                        float* __restrict__ bound, int nx, int ny, int nz, int axis, int cap) {
    size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)nx * ny * nz;
    if (id >= n) return;
    const int x = (int)(id % nx), y = (int)((id / nx) % ny), z = (int)(id / ((size_t)nx * ny));
    const int pos = axis == 0 ? x : (axis == 1 ? y : z), len = axis == 0 ? nx : (axis == 1 ? ny : nz);
    const size_t stride = axis == 0 ? 1 : (axis == 1 ? (size_t)nx : (size_t)nx * ny);
    auto value = [&](size_t i) -> int { return FROM_BOUND ? (bound[i] == 0.0f ? cap : -1) : (int)in[i]; };
    int m = value(id), r = -1;
    if (m >= 0) {
        r = 0;
        while (r < cap) {
            const int rr = r + 1;
            int a = pos - rr >= 0 ? value(id - (size_t)rr * stride) : cap;
            int b = pos + rr < len ? value(id + (size_t)rr * stride) : cap;
            m = min(m, min(a, b));
            if (m < rr) break;
            r = rr;
        }
    }
    if (TO_BOUND) {
        if (r >= 1) bound[id] = -(float)r;
    } else {
        out[id] = (signed char)r;
    }
}

}  // namespace

extern "C" {

int cpm_opacity_bound_clearance(cpm_ctx* ctx, float* bound, const int grid_dims[3], int max_radius) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, bound && grid_dims, "null argument");
    CPM_REQUIRE(ctx, grid_dims[0] > 0 && grid_dims[1] > 0 && grid_dims[2] > 0, "grid dims must be positive");
    CPM_REQUIRE(ctx, max_radius >= 1 && max_radius <= 100, "max_radius must be in 1..100");
    const size_t n = (size_t)grid_dims[0] * grid_dims[1] * grid_dims[2];
    void* scr = nullptr;
    int rc = cpm_scratch(ctx, 2 * n, &scr);
    if (rc != CPM_OK) return rc;
    signed char *a = (signed char*)scr, *b = a + n;
    const unsigned grid = cpm_div_up(n, 256);
    CPM_LAUNCH(ctx, (clearance_kernel<true, false>), grid, 256, 0, nullptr, a, bound, grid_dims[0], grid_dims[1], grid_dims[2], 0,
               max_radius);
    CPM_LAUNCH(ctx, (clearance_kernel<false, false>), grid, 256, 0, a, b, bound, grid_dims[0], grid_dims[1], grid_dims[2], 1,
               max_radius);
    CPM_LAUNCH(ctx, (clearance_kernel<false, true>), grid, 256, 0, b, nullptr, bound, grid_dims[0], grid_dims[1], grid_dims[2], 2,
               max_radius);
    return CPM_OK;
}

int cpm_bound_grid_dims(const int dims[3], int cell_log2, int out_dims[3]) {
    if (!dims || !out_dims || cell_log2 < 0 || cell_log2 > 8) return CPM_E_INVALID;
    for (int k = 0; k < 3; ++k) out_dims[k] = (dims[k] >> cell_log2) + 1;
    return CPM_OK;
}

int cpm_volume_value_range(cpm_ctx* ctx, const cpm_volume* vol, int cell_log2, float* range, int out_dims[3]) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, vol && range, "null argument");
    CPM_REQUIRE(ctx, cell_log2 >= 0 && cell_log2 <= 8, "cell_log2 must be in 0..8");
    CPM_REQUIRE(ctx, vol->linear != nullptr, "needs the LINEAR layout (reads the caller's buffer)");
    const int nx = vol->dims[0], ny = vol->dims[1], nz = vol->dims[2];
    int od[3];
    cpm_bound_grid_dims(vol->dims, cell_log2, od);
    if (out_dims) {
        out_dims[0] = od[0];
        out_dims[1] = od[1];
        out_dims[2] = od[2];
    }
    CPM_REQUIRE(ctx, od[2] <= 65535, "too many cell layers");
    size_t smem = (size_t)3 * od[0] * sizeof(uint32_t);
    CPM_REQUIRE(ctx, smem <= 48 * 1024, "volume too wide for the cell size");
    dim3 grid(od[1], od[2]);
    static const int generic_env = getenv("CPM_GRID_GENERIC") ? atoi(getenv("CPM_GRID_GENERIC")) : 0;   // A/B timing only
#define RG(F)                                                                                             \
    {                                                                                                     \
        int vec_ok = ((uintptr_t)vol->linear % 16 == 0) && (nx % VoxT<F>::PER16 == 0);                    \
        if (vec_ok && cell_log2 == 3 && !generic_env) {                                                   \
            CPM_LAUNCH(ctx, range8_kernel<F>, grid, 256, smem, vol->linear, nx, ny, nz, od[0],            \
                       (float2*)range);                                                                   \
        } else {                                                                                          \
            CPM_LAUNCH(ctx, range_kernel<F>, grid, 256, smem, vol->linear, nx, ny, nz, cell_log2, od[0],  \
                       (float2*)range, vec_ok);                                                           \
        }                                                                                                 \
    }
    switch (vol->format) {
        case CPM_FMT_U8: RG(CPM_FMT_U8) break;
        case CPM_FMT_U16: RG(CPM_FMT_U16) break;
        default: RG(CPM_FMT_F32)
    }
#undef RG
    return CPM_OK;
}

int cpm_opacity_bound(cpm_ctx* ctx, const float* range, size_t n_cells, float format_scale, float format_offset,
                      const float* tf_rgba, int tf_width, float* bound) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, range && tf_rgba && bound, "null argument");
    CPM_REQUIRE(ctx, tf_width >= 1 && tf_width <= 32768, "tf_width out of range");
    if (n_cells == 0) return CPM_OK;
    const int nb = (tf_width + 31) / 32;
    void* scr = nullptr;
    int rc = cpm_scratch(ctx, ((size_t)tf_width + 2 * (size_t)nb) * sizeof(float), &scr);
    if (rc != CPM_OK) return rc;
    float* alpha = (float*)scr;
    float* bmax = alpha + tf_width;
    float* babs = bmax + nb;
    CPM_LAUNCH(ctx, tf_summary_kernel, cpm_div_up(tf_width, 256), 256, 0, (const float4*)tf_rgba, tf_width, alpha, bmax,
               babs);
    CPM_LAUNCH(ctx, bound_kernel, cpm_div_up(n_cells, 256), 256, 0, (const float2*)range, n_cells, format_scale,
               format_offset, alpha, bmax, babs, tf_width, bound);
    return CPM_OK;
}

// ---- the bound grid as a 3-D texture -------------------------------------------------------------------------------------
// One look-up per collision test is the tracer's hottest load.  From a linear buffer it costs, per test, three floor
// conversions, two multiply-adds for the index, the bias, the address and the load -- plus one constant-bank load per
// grid parameter, because sm_100 ALU instructions take no constant operands and uniform registers do not live across the
// divergent scan loop.  As a point-sampled 3-D texture with clamped, unnormalised coordinates it is three FFMA (the ray
// in cell coordinates) and one TEX: the texture unit floors, clamps and addresses.
int cpm_bound_tex_create(cpm_ctx* ctx, const int grid_dims[3], cpm_bound_tex** out) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, grid_dims && out, "null argument");
    CPM_REQUIRE(ctx, grid_dims[0] >= 1 && grid_dims[1] >= 1 && grid_dims[2] >= 1, "grid dims must be positive");
    CPM_REQUIRE(ctx, grid_dims[0] <= 16384 && grid_dims[1] <= 16384 && grid_dims[2] <= 16384, "grid too large for a 3-D texture");
    cpm_bound_tex* t = new cpm_bound_tex();
    for (int k = 0; k < 3; ++k) t->dims[k] = grid_dims[k];
    cudaChannelFormatDesc cd = cudaCreateChannelDesc<float>();
    cudaError_t e = cudaMalloc3DArray(&t->array, &cd, make_cudaExtent(grid_dims[0], grid_dims[1], grid_dims[2]), 0);
    if (e != cudaSuccess) {
        delete t;
        return cpm_fail(ctx, e == cudaErrorMemoryAllocation ? CPM_E_NOMEM : CPM_E_CUDA, "cudaMalloc3DArray: %s", cudaGetErrorString(e));
    }
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = t->array;
    cudaTextureDesc td;
    memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    e = cudaCreateTextureObject(&t->tex, &rd, &td, nullptr);
    if (e != cudaSuccess) {
        cudaFreeArray(t->array);
        delete t;
        return cpm_fail(ctx, CPM_E_CUDA, "cudaCreateTextureObject: %s", cudaGetErrorString(e));
    }
    *out = t;
    return CPM_OK;
}

int cpm_bound_tex_update(cpm_ctx* ctx, cpm_bound_tex* t, const float* bound) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, t && bound, "null argument");
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof(p));
    p.srcPtr = make_cudaPitchedPtr(const_cast<float*>(bound), (size_t)t->dims[0] * sizeof(float), (size_t)t->dims[0], (size_t)t->dims[1]);
    p.dstArray = t->array;
    p.extent = make_cudaExtent(t->dims[0], t->dims[1], t->dims[2]);
    p.kind = cudaMemcpyDeviceToDevice;
    CPM_CUDA(ctx, cudaMemcpy3DAsync(&p, ctx->stream));
    return CPM_OK;
}

void cpm_bound_tex_destroy(cpm_bound_tex* t) {
    if (!t) return;
    if (t->tex) cudaDestroyTextureObject(t->tex);
    if (t->array) cudaFreeArray(t->array);
    delete t;
}

}  // extern "C"
