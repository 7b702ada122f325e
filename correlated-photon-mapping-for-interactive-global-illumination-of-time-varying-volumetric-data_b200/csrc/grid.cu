// grid.cu -- uniform-grid builders (north-star subsystem 6 and the importance-grid feeders of 4).
//
//  cpm_volume_minmax        volumeMinMaxKernel (ugc/cl/uniformgrid/volumeminmax.cl:33-61)
//  cpm_volume_diff_bricks   DynamicVolumeDifferenceAnalysis, a single-thread CPU triple loop in the
//                           reference (ugc/processors/dynamicvolumedifferenceanalysis.h:96-151)
//  cpm_classify_importance  classify[TimeVarying]MinMaxUniformGrid3DImportanceKernel
//                           (isc/cl/minmaxuniformgrid3dimportance.cl:164-228, 269-330)
//  cpm_hash_light_samples   hashLightSampleKernel (ppm/cl/hashlightsample.cl:38-66)
//  cpm_build_cell_ranges    [lower_bound, upper_bound) of every cell over sorted keys
//
// The two volume passes are HBM streaming kernels: one CTA per row of bricks (fixed by, bz),
// 16-byte vector loads along x, per-brick accumulators in shared memory.  The reference reads
// one texel per work-item iteration through the image path (4x4x4 work-groups over bricks).
#include "sampling.cuh"

namespace {

__device__ __forceinline__ uint32_t order_key(float f) {  // monotone float -> uint
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float order_val(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

template <int FMT>
struct Vox;
template <>
struct Vox<CPM_FMT_U8> {
    typedef unsigned char T;
    static constexpr int PER16 = 16;
    __device__ static float norm(T v) { return unorm8((float)v); }
};
template <>
struct Vox<CPM_FMT_U16> {
    typedef unsigned short T;
    static constexpr int PER16 = 8;
    __device__ static float norm(T v) { return unorm16((float)v); }
};
template <>
struct Vox<CPM_FMT_F32> {
    typedef float T;
    static constexpr int PER16 = 4;
    __device__ static float norm(T v) { return v; }
};

__device__ __forceinline__ uint16_t to_u16(float v) { return (uint16_t)rintf(cpm_clamp(v * 65535.0f, 0.0f, 65535.0f)); }

// one CTA per (by, bz): all bricks along x of that row.  smem: 2 * nbx uint32 order keys.
template <int FMT>
__global__ void __launch_bounds__(256) minmax_kernel(const void* __restrict__ vol, int nx, int ny, int nz, int region,
                                                     int nbx, float scale, float offset, ushort2* __restrict__ out,
                                                     int vec_ok) {
    typedef typename Vox<FMT>::T T;
    constexpr int K = Vox<FMT>::PER16;
    extern __shared__ uint32_t s_mm[];  // [0,nbx) min keys, [nbx,2nbx) max keys
    const int by = blockIdx.x, bz = blockIdx.y;
    for (int i = threadIdx.x; i < nbx; i += blockDim.x) {
        s_mm[i] = 0xffffffffu;
        s_mm[nbx + i] = 0u;
    }
    __syncthreads();
    const int y0 = by * region, z0 = bz * region;
    const int ry = min(region, ny - y0), rz = min(region, nz - z0);
    const int rows = ry * rz;
    const T* base = (const T*)vol;
    if (vec_ok) {
        const int chunks = nx / K;  // nx % K == 0 guaranteed by vec_ok
        const int items = rows * chunks;
        for (int w = threadIdx.x; w < items; w += blockDim.x) {
            int r = w / chunks, c = w - r * chunks;
            int y = y0 + r % ry, z = z0 + r / ry;
            const T* row = base + ((size_t)z * ny + y) * nx;
            uint4 raw = __ldg(reinterpret_cast<const uint4*>(row) + c);
            const T* e = reinterpret_cast<const T*>(&raw);
            int x = c * K;
            int cur = x / region;
            float mn = CPM_FLT_MAX_, mx = -CPM_FLT_MAX_;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                int b = (x + k) / region;
                if (b != cur) {
                    atomicMin(&s_mm[cur], order_key(mn));
                    atomicMax(&s_mm[nbx + cur], order_key(mx));
                    cur = b;
                    mn = CPM_FLT_MAX_;
                    mx = -CPM_FLT_MAX_;
                }
                float v = (Vox<FMT>::norm(e[k]) + offset) * scale;
                mn = fminf(mn, v);
                mx = fmaxf(mx, v);
            }
            atomicMin(&s_mm[cur], order_key(mn));
            atomicMax(&s_mm[nbx + cur], order_key(mx));
        }
    } else {
        const int items = rows * nx;
        for (int w = threadIdx.x; w < items; w += blockDim.x) {
            int r = w / nx, x = w - r * nx;
            int y = y0 + r % ry, z = z0 + r / ry;
            float v = (Vox<FMT>::norm(base[((size_t)z * ny + y) * nx + x]) + offset) * scale;
            atomicMin(&s_mm[x / region], order_key(v));
            atomicMax(&s_mm[nbx + x / region], order_key(v));
        }
    }
    __syncthreads();
    const int nby = gridDim.x;
    for (int i = threadIdx.x; i < nbx; i += blockDim.x) {
        // reference initial value is (FLT_MAX, 0): min over {FLT_MAX, values}, max over {0, values}
        float mn = fminf(CPM_FLT_MAX_, order_val(s_mm[i]));
        float mx = fmaxf(0.0f, order_val(s_mm[nbx + i]));
        out[((size_t)bz * nby + by) * nbx + i] = make_ushort2(to_u16(mn), to_u16(mx));
    }
}

// ---- bricks of 8 voxels (the default region), 16-byte aligned rows: streaming versions ---------------------------
// Same CTA mapping, but a thread owns one 16-byte chunk COLUMN and walks down the brick row's 64 voxel rows: the
// brick(s) a chunk feeds are fixed per thread, partial results stay in registers (packed bytes / halves for the
// integer formats: min / max commute with the monotone normalisation), four rows are in flight per thread and
// shared memory sees one or two atomics per thread instead of several per chunk.
template <int FMT>
struct MinMaxAcc;
template <>
struct MinMaxAcc<CPM_FMT_F32> {
    float mn[4], mx[4];
    __device__ void init() {
#pragma unroll
        for (int k = 0; k < 4; ++k) mn[k] = CPM_FLT_MAX_, mx[k] = -CPM_FLT_MAX_;
    }
    __device__ void add(const uint4& r) {
        const float w[4] = {__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z), __uint_as_float(r.w)};
#pragma unroll
        for (int k = 0; k < 4; ++k) mn[k] = fminf(mn[k], w[k]), mx[k] = fmaxf(mx[k], w[k]);   // NaN voxels are skipped
    }
    template <class PUT>
    __device__ void flush(int c, PUT put) const {
        put(c >> 1, fminf(fminf(mn[0], mn[1]), fminf(mn[2], mn[3])), fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])));
    }
};
template <>
struct MinMaxAcc<CPM_FMT_U16> {
    uint32_t mn[4], mx[4];
    __device__ void init() {
#pragma unroll
        for (int k = 0; k < 4; ++k) mn[k] = 0xffffffffu, mx[k] = 0u;
    }
    __device__ void add(const uint4& r) {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) mn[k] = __vminu2(mn[k], w[k]), mx[k] = __vmaxu2(mx[k], w[k]);
    }
    template <class PUT>
    __device__ void flush(int c, PUT put) const {
        uint32_t a = __vminu2(__vminu2(mn[0], mn[1]), __vminu2(mn[2], mn[3]));
        uint32_t b = __vmaxu2(__vmaxu2(mx[0], mx[1]), __vmaxu2(mx[2], mx[3]));
        put(c, unorm16((float)min(a & 0xffffu, a >> 16)), unorm16((float)max(b & 0xffffu, b >> 16)));
    }
};
template <>
struct MinMaxAcc<CPM_FMT_U8> {
    uint32_t mn[4], mx[4];
    __device__ void init() {
#pragma unroll
        for (int k = 0; k < 4; ++k) mn[k] = 0xffffffffu, mx[k] = 0u;
    }
    __device__ void add(const uint4& r) {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) mn[k] = __vminu4(mn[k], w[k]), mx[k] = __vmaxu4(mx[k], w[k]);
    }
    __device__ static uint32_t lo4(uint32_t v) { v = __vminu4(v, v >> 16); return min(v & 0xffu, (v >> 8) & 0xffu); }
    __device__ static uint32_t hi4(uint32_t v) { v = __vmaxu4(v, v >> 16); return max(v & 0xffu, (v >> 8) & 0xffu); }
    template <class PUT>
    __device__ void flush(int c, PUT put) const {
        put(2 * c, unorm8((float)lo4(__vminu4(mn[0], mn[1]))), unorm8((float)hi4(__vmaxu4(mx[0], mx[1]))));
        put(2 * c + 1, unorm8((float)lo4(__vminu4(mn[2], mn[3]))), unorm8((float)hi4(__vmaxu4(mx[2], mx[3]))));
    }
};

// row r of the brick row (by, bz) -> index of its first 16-byte chunk
__device__ __forceinline__ void brick_row_offsets(unsigned long long* s_row, int rows, int ry, int y0, int z0, int ny,
                                                  int chunks) {
    for (int i = threadIdx.x; i < rows; i += blockDim.x) {
        const int zr = i / ry;
        s_row[i] = ((unsigned long long)(z0 + zr) * ny + (y0 + i - zr * ry)) * chunks;
    }
}

template <int FMT>
__global__ void __launch_bounds__(256) minmax8_kernel(const void* __restrict__ vol, int nx, int ny, int nz, int nbx,
                                                      float scale, float offset, ushort2* __restrict__ out) {
    constexpr int K = Vox<FMT>::PER16;
    extern __shared__ uint32_t s_mm[];  // [0,nbx) min keys, [nbx,2nbx) max keys
    __shared__ unsigned long long s_row[64];
    const int by = blockIdx.x, bz = blockIdx.y;
    for (int i = threadIdx.x; i < nbx; i += blockDim.x) {
        s_mm[i] = 0xffffffffu;
        s_mm[nbx + i] = 0u;
    }
    const int y0 = by * 8, z0 = bz * 8;
    const int ry = min(8, ny - y0), rz = min(8, nz - z0);
    const int rows = ry * rz, chunks = nx / K;
    brick_row_offsets(s_row, rows, ry, y0, z0, ny, chunks);
    __syncthreads();
    const uint4* V = reinterpret_cast<const uint4*>(vol);
    const int cols = min(chunks, (int)blockDim.x), groups = blockDim.x / cols;
    const int col = threadIdx.x % cols, rg = threadIdx.x / cols;
    auto put = [&](int b, float lo, float hi) {
        if (b >= nbx || lo > hi) return;
        // the normalisation (v + offset) * scale is monotone (either way) and so is its rounding
        const float a = (lo + offset) * scale, c = (hi + offset) * scale;
        atomicMin(&s_mm[b], order_key(fminf(a, c)));
        atomicMax(&s_mm[nbx + b], order_key(fmaxf(a, c)));
    };
    if (rg < groups) {
        for (int c = col; c < chunks; c += cols) {
            MinMaxAcc<FMT> acc;
            acc.init();
            int r = rg;
            for (; r + 3 * groups < rows; r += 4 * groups) {
                uint4 a0 = __ldg(V + s_row[r] + c), a1 = __ldg(V + s_row[r + groups] + c);
                uint4 a2 = __ldg(V + s_row[r + 2 * groups] + c), a3 = __ldg(V + s_row[r + 3 * groups] + c);
                acc.add(a0); acc.add(a1); acc.add(a2); acc.add(a3);
            }
            for (; r < rows; r += groups) acc.add(__ldg(V + s_row[r] + c));
            acc.flush(c, put);
        }
    }
    __syncthreads();
    const int nby = gridDim.x;
    for (int i = threadIdx.x; i < nbx; i += blockDim.x) {
        float mn = fminf(CPM_FLT_MAX_, order_val(s_mm[i]));
        float mx = fmaxf(0.0f, order_val(s_mm[nbx + i]));
        out[((size_t)bz * nby + by) * nbx + i] = make_ushort2(to_u16(mn), to_u16(mx));
    }
}

// ---- inter-step difference bricks --------------------------------------------------------------------------------
// One THREAD per brick, voxels added in the reference's order (z, then y, then x ascending:
// ugc/processors/dynamicvolumedifferenceanalysis.h:123-138).  The double-precision sum is then the reference's bit
// for bit and the same on every run (a sum combined through atomics is neither).  Lanes of a warp own consecutive
// bricks along x, so a row read by a warp is one contiguous segment of 32 bricks.
template <int FMT>
struct BrickRow;   // the 8 voxels of one brick row
template <>
struct BrickRow<CPM_FMT_F32> {
    uint4 a, b;
    __device__ void load(const void* p) { a = __ldg((const uint4*)p); b = __ldg((const uint4*)p + 1); }
    __device__ double get(int k) const {
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        return (double)__uint_as_float(w[k]);
    }
};
template <>
struct BrickRow<CPM_FMT_U16> {
    uint4 a;
    __device__ void load(const void* p) { a = __ldg((const uint4*)p); }
    __device__ double get(int k) const {
        const uint32_t w[4] = {a.x, a.y, a.z, a.w};
        return (double)((w[k >> 1] >> ((k & 1) * 16)) & 0xffffu);
    }
};
template <>
struct BrickRow<CPM_FMT_U8> {
    uint2 a;
    __device__ void load(const void* p) { a = __ldg((const uint2*)p); }
    __device__ double get(int k) const {
        const uint32_t w[2] = {a.x, a.y};
        return (double)((w[k >> 2] >> ((k & 3) * 8)) & 0xffu);
    }
};

// region 8, nx % 8 == 0, 16-byte aligned buffers.  Block (32, 8): 32 bricks along x, 8 along y; grid.z = brick layer.
template <int FMT>
__global__ void __launch_bounds__(256) diff8_kernel(const void* __restrict__ va, const void* __restrict__ vb, int nx,
                                                    int ny, int nz, int nbx, int nby, double scaling, double rmin,
                                                    double rmax, float* __restrict__ out) {
    typedef typename Vox<FMT>::T T;
    const int bx = blockIdx.x * 32 + threadIdx.x, by = blockIdx.y * 8 + threadIdx.y, bz = blockIdx.z;
    if (bx >= nbx || by >= nby) return;
    const int y0 = by * 8, z0 = bz * 8;
    const int ry = min(8, ny - y0), rz = min(8, nz - z0);
    const T *A = (const T*)va, *B = (const T*)vb;
    double acc = 0.0;
    for (int z = 0; z < rz; ++z) {
        const size_t slab = ((size_t)(z0 + z) * ny + y0) * nx + (size_t)bx * 8;
        for (int y = 0; y < ry; y += 4) {          // four rows of both volumes in flight, added in order
            BrickRow<FMT> ra[4], rb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (y + j < ry) {
                    ra[j].load(A + slab + (size_t)(y + j) * nx);
                    rb[j].load(B + slab + (size_t)(y + j) * nx);
                }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (y + j < ry) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc += fabs(scaling * (rb[j].get(k) - ra[j].get(k)));
                }
        }
    }
    out[((size_t)bz * nby + by) * nbx + bx] = (float)((acc / 512.0 - rmin) / (rmax - rmin));
}

// any region / row length: the same order with scalar loads
template <int FMT>
__global__ void __launch_bounds__(256) diff_kernel(const void* __restrict__ va, const void* __restrict__ vb, int nx,
                                                   int ny, int nz, int region, int nbx, int nby, double scaling,
                                                   double rmin, double rmax, float* __restrict__ out) {
    typedef typename Vox<FMT>::T T;
    const int bx = blockIdx.x * 32 + threadIdx.x, by = blockIdx.y * 8 + threadIdx.y, bz = blockIdx.z;
    if (bx >= nbx || by >= nby) return;
    const int x0 = bx * region, y0 = by * region, z0 = bz * region;
    const int x1 = min(x0 + region, nx), y1 = min(y0 + region, ny), z1 = min(z0 + region, nz);
    const T *A = (const T*)va, *B = (const T*)vb;
    double acc = 0.0;
    for (int z = z0; z < z1; ++z)
        for (int y = y0; y < y1; ++y) {
            const size_t row = ((size_t)z * ny + y) * nx;
            for (int x = x0; x < x1; ++x) acc += fabs(scaling * ((double)B[row + x] - (double)A[row + x]));
        }
    const double r3 = (double)region * region * region;
    out[((size_t)bz * nby + by) * nbx + bx] = (float)((acc / r3 - rmin) / (rmax - rmin));
}

// ---- importance classification ----------------------------------------------------------------
__device__ __forceinline__ float4 mix4(float4 a, float4 b, float t) {
    return make_float4(fmaf(b.x - a.x, t, a.x), fmaf(b.y - a.y, t, a.y), fmaf(b.z - a.z, t, a.z), fmaf(b.w - a.w, t, a.w));
}
__device__ __forceinline__ float4 min4(float4 a, float4 b) {
    return make_float4(cpm_fmin(a.x, b.x), cpm_fmin(a.y, b.y), cpm_fmin(a.z, b.z), cpm_fmin(a.w, b.w));
}
__device__ __forceinline__ float4 max4(float4 a, float4 b) {
    return make_float4(cpm_fmax(a.x, b.x), cpm_fmax(a.y, b.y), cpm_fmax(a.z, b.z), cpm_fmax(a.w, b.w));
}
__device__ void rgb2lab(float r, float g, float b, float lab[3]) {
    float c[3] = {r, g, b}, lin[3];
    for (int k = 0; k < 3; ++k) lin[k] = c[k] > 0.04045f ? cpm_powf((c[k] + 0.055f) / 1.055f, 2.4f) : c[k] / 12.92f;
    float X = 0.4124564f * lin[0] + 0.3575761f * lin[1] + 0.1804375f * lin[2];
    float Y = 0.2126729f * lin[0] + 0.7151522f * lin[1] + 0.0721750f * lin[2];
    float Z = 0.0193339f * lin[0] + 0.1191920f * lin[1] + 0.9503041f * lin[2];
    float xyz[3] = {X / 0.95047f, Y / 1.0f, Z / 1.08883f}, f[3];
    for (int k = 0; k < 3; ++k) f[k] = xyz[k] > 0.008856f ? cpm_cbrtf(xyz[k]) : 7.787f * xyz[k] + 16.0f / 116.0f;
    lab[0] = 116.0f * f[1] - 16.0f;
    lab[1] = 500.0f * (f[0] - f[1]);
    lab[2] = 200.0f * (f[1] - f[2]);
}
__device__ float tf_points_importance(float4 color, float4 next, const float* w, int incremental) {
    if (incremental) return next.x + next.y + next.z + next.w;
    float imp = 0.0f;
    if (color.w > 0.0f || next.w > 0.0f) {
        float la[3], lb[3];
        rgb2lab(color.x, color.y, color.z, la);
        rgb2lab(next.x, next.y, next.z, lb);
        float d0 = lb[0] - la[0], d1 = lb[1] - la[1], d2 = lb[2] - la[2];
        // length(): the definition used everywhere on the path (fma chain from x)
        float lenN = sqrtf(fmaf(lb[2], lb[2], fmaf(lb[1], lb[1], lb[0] * lb[0])));
        float lenC = sqrtf(fmaf(la[2], la[2], fmaf(la[1], la[1], la[0] * la[0])));
        float lenD = sqrtf(fmaf(d2, d2, fmaf(d1, d1, d0 * d0)));
        imp = w[0] * fmaxf(lenN, lenC) + w[1] * lenD + w[2] * fabsf(next.w - color.w) + w[3] * fmaxf(color.w, next.w);
    }
    return imp;
}

struct ClassifyArgs {
    const ushort2* mm;
    const ushort2* prev;
    const float* diff;
    int n;
    const float* pos;
    const float4* col;
    int n_points;
    float w[4];
    int incremental;
    float* out;
};

__global__ void __launch_bounds__(128) classify_kernel(const ClassifyArgs A) {
    extern __shared__ float4 s_col[];                 // n_points colours, then positions
    float* s_pos = (float*)(s_col + A.n_points);
    for (int i = threadIdx.x; i < A.n_points; i += blockDim.x) {
        s_col[i] = A.col[i];
        s_pos[i] = A.pos[i];
    }
    __syncthreads();
    int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= A.n) return;
    ushort2 c = A.mm[id];
    if (A.prev) {
        ushort2 p = A.prev[id];
        c.x = min(c.x, p.x);
        c.y = max(c.y, p.y);
    }
    float lo = (1.0f / 65535.0f) * (float)c.x, hi = (1.0f / 65535.0f) * (float)c.y;
    const int nP = A.n_points;
    int i = 0;
    while (i < nP - 1 && lo > s_pos[i + 1]) ++i;
    float4 color = mix4(s_col[i], s_col[i + 1], (lo - s_pos[i]) / (s_pos[i + 1] - s_pos[i]));
    float4 mn = color, mx = color;
    float imp;
    if (hi <= s_pos[i + 1]) {
        float4 nc = mix4(s_col[i], s_col[i + 1], (hi - s_pos[i]) / (s_pos[i + 1] - s_pos[i]));
        mn = min4(mn, nc);
        mx = max4(mx, nc);
        imp = tf_points_importance(mn, mx, A.w, A.incremental);
    } else {
        float4 nc = s_col[i + 1];
        mn = min4(mn, nc);
        mx = max4(mx, nc);
        ++i;
        while (i < nP - 1 && hi > s_pos[i + 1]) {
            nc = s_col[i + 1];
            mn = min4(mn, nc);
            mx = max4(mx, nc);
            ++i;
        }
        if (i < nP - 1) {
            color = mix4(s_col[i], s_col[i + 1], (hi - s_pos[i]) / (s_pos[i + 1] - s_pos[i]));
            mn = min4(mn, color);
            mx = max4(mx, color);
        }
        imp = tf_points_importance(mn, mx, A.w, A.incremental);
    }
    A.out[id] = A.prev ? A.diff[id] * imp : imp;
}

// ---- light-sample cell hash and cell ranges ---------------------------------------------------------
__global__ void __launch_bounds__(128) hash_kernel(const float4* __restrict__ ls, const float2* __restrict__ isect,
                                                   int n_src, const uint32_t* __restrict__ ids, int n_ids, float cx,
                                                   float cy, float cz, int nbx, int nby, uint32_t* __restrict__ bucket,
                                                   int out_offset) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_ids) return;
    uint32_t id = ids[g];
    if (id < (uint32_t)out_offset || id >= (uint32_t)n_src) return;
    float4 l0 = ls[2 * (size_t)id], l1 = ls[2 * (size_t)id + 1];
    float3_ d = decode_direction(l1.z, l1.w);
    float ts = isect[id].x;
    // lightSample.origin + tStart * lightSample.direction as written (ppm/cl/hashlightsample.cl:56)
    float px = __fadd_rn(l0.x, __fmul_rn(ts, d.x)), py = __fadd_rn(l0.y, __fmul_rn(ts, d.y)), pz = __fadd_rn(l0.z, __fmul_rn(ts, d.z));
    uint32_t hx = (uint32_t)cpm_clamp(truncf(px * cx), 0.f, 4294967040.f);
    uint32_t hy = (uint32_t)cpm_clamp(truncf(py * cy), 0.f, 4294967040.f);
    uint32_t hz = (uint32_t)cpm_clamp(truncf(pz * cz), 0.f, 4294967040.f);
    bucket[out_offset + g] = hz * (uint32_t)nbx * (uint32_t)nby + hy * (uint32_t)nbx + hx;
}

// One thread per cell boundary c in [0, n_cells]: lb = lower_bound(keys, c) is where cell c starts and where cell
// c - 1 ends.  A binary search per cell (22 L2-resident probes for 4 Mi keys) costs the same for dense and for empty
// stretches of the grid; walking the key array instead leaves one thread to fill every cell of a long empty run
// (measured: 0.87 ms of a 1.2 ms photon-map build for the run behind the last occupied cell).
__global__ void __launch_bounds__(256) cell_range_kernel(const uint32_t* __restrict__ keys, size_t n, uint32_t n_cells,
                                                         uint32_t* __restrict__ start, uint32_t* __restrict__ end) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_cells) return;
    size_t lo = 0, hi = n;   // first index with keys[index] >= c
    while (lo < hi) {
        size_t mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) < (uint32_t)c) lo = mid + 1;
        else hi = mid;
    }
    if (c < n_cells) start[c] = (uint32_t)lo;
    if (c > 0) end[c - 1] = (uint32_t)lo;
}

static bool grid_generic() {   // CPM_GRID_GENERIC=1: the generic kernels also for region 8 (A/B timing only)
    static const int v = getenv("CPM_GRID_GENERIC") ? atoi(getenv("CPM_GRID_GENERIC")) : 0;
    return v != 0;
}

}  // namespace

extern "C" {

int cpm_volume_minmax(cpm_ctx* ctx, const cpm_volume* vol, int region, uint16_t* out, int out_dims[3]) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, vol && out, "null argument");
    CPM_REQUIRE(ctx, region >= 1, "region must be >= 1");
    CPM_REQUIRE(ctx, vol->linear != nullptr, "needs the LINEAR layout (reads the caller's buffer)");
    const int nx = vol->dims[0], ny = vol->dims[1], nz = vol->dims[2];
    const int nbx = (nx + region - 1) / region, nby = (ny + region - 1) / region, nbz = (nz + region - 1) / region;
    if (out_dims) {
        out_dims[0] = nbx;
        out_dims[1] = nby;
        out_dims[2] = nbz;
    }
    CPM_REQUIRE(ctx, nbz <= 65535, "too many brick layers");
    size_t smem = (size_t)2 * nbx * sizeof(uint32_t);
    dim3 grid(nby, nbz);
#define MM(F)                                                                                              \
    {                                                                                                      \
        int vec_ok = ((uintptr_t)vol->linear % 16 == 0) && (nx % Vox<F>::PER16 == 0);                      \
        if (vec_ok && region == 8 && !grid_generic()) {                                                    \
            CPM_LAUNCH(ctx, minmax8_kernel<F>, grid, 256, smem, vol->linear, nx, ny, nz, nbx, vol->scale,  \
                       vol->offset, (ushort2*)out);                                                        \
        } else {                                                                                           \
            CPM_LAUNCH(ctx, minmax_kernel<F>, grid, 256, smem, vol->linear, nx, ny, nz, region, nbx, vol->scale, \
                       vol->offset, (ushort2*)out, vec_ok);                                                \
        }                                                                                                  \
    }
    switch (vol->format) {
        case CPM_FMT_U8: MM(CPM_FMT_U8) break;
        case CPM_FMT_U16: MM(CPM_FMT_U16) break;
        default: MM(CPM_FMT_F32)
    }
#undef MM
    return CPM_OK;
}

int cpm_volume_diff_bricks(cpm_ctx* ctx, const cpm_volume* a, const cpm_volume* b, int region, double data_scaling,
                           double range_min, double range_max, float* out) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, a && b && out, "null argument");
    CPM_REQUIRE(ctx, region >= 1, "region must be >= 1");
    CPM_REQUIRE(ctx, a->linear && b->linear, "needs the LINEAR layout");
    CPM_REQUIRE(ctx, a->format == b->format && a->dims[0] == b->dims[0] && a->dims[1] == b->dims[1] && a->dims[2] == b->dims[2],
                "volumes differ in format or size");
    const int nx = a->dims[0], ny = a->dims[1], nz = a->dims[2];
    const int nbx = (nx + region - 1) / region, nby = (ny + region - 1) / region, nbz = (nz + region - 1) / region;
    CPM_REQUIRE(ctx, nbz <= 65535, "too many brick layers");
    dim3 grid(cpm_div_up(nbx, 32), cpm_div_up(nby, 8), nbz), block(32, 8);
#define DF(F)                                                                                                      \
    {                                                                                                              \
        int vec_ok = ((uintptr_t)a->linear % 16 == 0) && ((uintptr_t)b->linear % 16 == 0) && (nx % 8 == 0);        \
        if (vec_ok && region == 8 && !grid_generic()) {                                                            \
            CPM_LAUNCH(ctx, diff8_kernel<F>, grid, block, 0, a->linear, b->linear, nx, ny, nz, nbx, nby, data_scaling, \
                       range_min, range_max, out);                                                                 \
        } else {                                                                                                   \
            CPM_LAUNCH(ctx, diff_kernel<F>, grid, block, 0, a->linear, b->linear, nx, ny, nz, region, nbx, nby,    \
                       data_scaling, range_min, range_max, out);                                                   \
        }                                                                                                          \
    }
    switch (a->format) {
        case CPM_FMT_U8: DF(CPM_FMT_U8) break;
        case CPM_FMT_U16: DF(CPM_FMT_U16) break;
        default: DF(CPM_FMT_F32)
    }
#undef DF
    return CPM_OK;
}

int cpm_classify_importance(cpm_ctx* ctx, const uint16_t* minmax, const uint16_t* prev_minmax, const float* volume_diff,
                            int n, const float* tf_positions, const float* tf_colors, int n_points,
                            const float weights[4], int incremental, float* out) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, n >= 0, "negative n");
    if (n == 0) return CPM_OK;
    CPM_REQUIRE(ctx, minmax && tf_positions && tf_colors && weights && out, "null argument");
    CPM_REQUIRE(ctx, (prev_minmax == nullptr) == (volume_diff == nullptr), "prev_minmax and volume_diff go together");
    CPM_REQUIRE(ctx, n_points >= 2 && n_points <= 2048, "n_points out of range");
    ClassifyArgs a;
    a.mm = (const ushort2*)minmax;
    a.prev = (const ushort2*)prev_minmax;
    a.diff = volume_diff;
    a.n = n;
    a.pos = tf_positions;
    a.col = (const float4*)tf_colors;
    a.n_points = n_points;
    for (int k = 0; k < 4; ++k) a.w[k] = weights[k];
    a.incremental = incremental;
    a.out = out;
    CPM_LAUNCH(ctx, classify_kernel, cpm_div_up(n, 128), 128, (size_t)n_points * 20, a);
    return CPM_OK;
}

int cpm_hash_light_samples(cpm_ctx* ctx, const float* light_samples, const float* intersections,
                           int n_light_source_samples, const uint32_t* ids, int n_ids, const float cell_size[3],
                           const int n_blocks[3], uint32_t* which_bucket, int out_offset) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, n_ids >= 0, "negative n_ids");
    if (n_ids == 0) return CPM_OK;
    CPM_REQUIRE(ctx, light_samples && intersections && ids && cell_size && n_blocks && which_bucket, "null argument");
    CPM_LAUNCH(ctx, hash_kernel, cpm_div_up(n_ids, 128), 128, 0, (const float4*)light_samples, (const float2*)intersections,
               n_light_source_samples, ids, n_ids, cell_size[0], cell_size[1], cell_size[2], n_blocks[0], n_blocks[1],
               which_bucket, out_offset);
    return CPM_OK;
}

int cpm_build_cell_ranges(cpm_ctx* ctx, const uint32_t* sorted_keys, size_t n, uint32_t n_cells, uint32_t* cell_start,
                          uint32_t* cell_end) {
    if (!ctx) return CPM_E_INVALID;
    if (n_cells == 0) return CPM_OK;
    CPM_REQUIRE(ctx, cell_start && cell_end && (sorted_keys || n == 0), "null argument");
    CPM_LAUNCH(ctx, cell_range_kernel, cpm_div_up((size_t)n_cells + 1, 256), 256, 0, sorted_keys, n, n_cells, cell_start, cell_end);
    return CPM_OK;
}

}  // extern "C"
