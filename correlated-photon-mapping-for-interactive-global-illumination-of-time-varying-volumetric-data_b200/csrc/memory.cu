// memory.cu -- device memory and transfer entry points, so that host code above the C ABI
// (the Inviwo processor mirror in host/) needs no CUDA headers.  Counterparts of cl::Buffer,
// enqueueWriteBuffer / enqueueReadBuffer / enqueueCopyBuffer / enqueueFillBuffer.
#include "common.cuh"

extern "C" {

int cpm_mem_alloc(cpm_ctx* ctx, size_t bytes, void** out) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, out != nullptr, "out is NULL");
    *out = nullptr;
    if (bytes == 0) return CPM_OK;
    // stream-ordered allocation from the device's default pool (release threshold raised in
    // cpm_ctx_create): per-frame buffers of the host layer are recycled without a device sync
    CPM_CUDA(ctx, cudaMallocAsync(out, bytes, ctx->stream));
    return CPM_OK;
}

int cpm_mem_free(cpm_ctx* ctx, void* ptr) {
    if (!ctx) return CPM_E_INVALID;
    if (!ptr) return CPM_OK;
    CPM_CUDA(ctx, cudaFreeAsync(ptr, ctx->stream));
    return CPM_OK;
}

int cpm_host_alloc(cpm_ctx* ctx, size_t bytes, void** out) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, out != nullptr, "out is NULL");
    *out = nullptr;
    if (bytes == 0) return CPM_OK;
    CPM_CUDA(ctx, cudaMallocHost(out, bytes));
    return CPM_OK;
}

int cpm_host_free(cpm_ctx* ctx, void* ptr) {
    if (!ctx) return CPM_E_INVALID;
    if (!ptr) return CPM_OK;
    CPM_CUDA(ctx, cudaFreeHost(ptr));
    return CPM_OK;
}

int cpm_mem_copy_h2d(cpm_ctx* ctx, void* dst, const void* src_host, size_t bytes) {
    if (!ctx) return CPM_E_INVALID;
    if (bytes == 0) return CPM_OK;
    CPM_REQUIRE(ctx, dst && src_host, "null pointer");
    CPM_CUDA(ctx, cudaMemcpyAsync(dst, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return CPM_OK;
}

int cpm_mem_copy_d2h(cpm_ctx* ctx, void* dst_host, const void* src, size_t bytes) {
    if (!ctx) return CPM_E_INVALID;
    if (bytes == 0) return CPM_OK;
    CPM_REQUIRE(ctx, dst_host && src, "null pointer");
    CPM_CUDA(ctx, cudaMemcpyAsync(dst_host, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return CPM_OK;
}

int cpm_mem_prefetch_h2d(cpm_ctx* ctx, void* dst, const void* src_host, size_t bytes, cpm_event** done) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, dst && src_host && done && bytes > 0, "null argument");
    if (!ctx->xfer_stream) {
        CPM_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->xfer_stream, cudaStreamNonBlocking));
        CPM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->xfer_fence, cudaEventDisableTiming));
    }
    // earlier readers / writers of dst on the context stream finish first
    CPM_CUDA(ctx, cudaEventRecord(ctx->xfer_fence, ctx->stream));
    CPM_CUDA(ctx, cudaStreamWaitEvent(ctx->xfer_stream, ctx->xfer_fence, 0));
    CPM_CUDA(ctx, cudaMemcpyAsync(dst, src_host, bytes, cudaMemcpyHostToDevice, ctx->xfer_stream));
    cpm_event* e = new cpm_event();
    cudaError_t rc = cudaEventCreateWithFlags(&e->ev, cudaEventDisableTiming);
    if (rc != cudaSuccess) {
        delete e;
        return cpm_fail(ctx, CPM_E_CUDA, "cudaEventCreate: %s", cudaGetErrorString(rc));
    }
    CPM_CUDA(ctx, cudaEventRecord(e->ev, ctx->xfer_stream));
    *done = e;
    return CPM_OK;
}

// The mirror image of cpm_mem_prefetch_h2d: device -> (pinned) host on a read-back stream of its own, after the work
// already submitted to the context stream; the context stream does not wait.  *done: hand it to cpm_event_sync before
// the host reads dst_host, and to cpm_ctx_wait_event before src is overwritten.
int cpm_mem_readback_d2h(cpm_ctx* ctx, void* dst_host, const void* src, size_t bytes, cpm_event** done) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, dst_host && src && done && bytes > 0, "null argument");
    if (!ctx->d2h_stream) {
        CPM_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
        CPM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->d2h_fence, cudaEventDisableTiming));
    }
    CPM_CUDA(ctx, cudaEventRecord(ctx->d2h_fence, ctx->stream));
    CPM_CUDA(ctx, cudaStreamWaitEvent(ctx->d2h_stream, ctx->d2h_fence, 0));
    CPM_CUDA(ctx, cudaMemcpyAsync(dst_host, src, bytes, cudaMemcpyDeviceToHost, ctx->d2h_stream));
    cpm_event* e = new cpm_event();
    cudaError_t rc = cudaEventCreateWithFlags(&e->ev, cudaEventDisableTiming);
    if (rc != cudaSuccess) {
        delete e;
        return cpm_fail(ctx, CPM_E_CUDA, "cudaEventCreate: %s", cudaGetErrorString(rc));
    }
    CPM_CUDA(ctx, cudaEventRecord(e->ev, ctx->d2h_stream));
    *done = e;
    return CPM_OK;
}

int cpm_event_sync(cpm_ctx* ctx, cpm_event* ev) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, ev != nullptr, "null argument");
    CPM_CUDA(ctx, cudaEventSynchronize(ev->ev));
    return CPM_OK;
}

int cpm_ctx_wait_event(cpm_ctx* ctx, cpm_event* ev) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, ev != nullptr, "null argument");
    CPM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev->ev, 0));
    return CPM_OK;
}

int cpm_ctx_wait_cuda_event(cpm_ctx* ctx, void* cuda_event) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, cuda_event != nullptr, "null argument");
    CPM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, (cudaEvent_t)cuda_event, 0));
    return CPM_OK;
}

int cpm_mem_copy_d2d(cpm_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return CPM_E_INVALID;
    if (bytes == 0) return CPM_OK;
    CPM_REQUIRE(ctx, dst && src, "null pointer");
    // memmove semantics are what the reference's chunked self-copy of the index list emulates
    // (ppm/processor/progressivephotontracercl.cpp:389-419); cudaMemcpyAsync D2D needs disjoint ranges
    const char *d = (const char*)dst, *s = (const char*)src;
    if (d < s + bytes && s < d + bytes) {
        size_t gap = d < s ? (size_t)(s - d) : (size_t)(d - s);
        CPM_REQUIRE(ctx, d < s && gap > 0, "overlapping copy must move data toward lower addresses");
        for (size_t off = 0; off < bytes; off += gap) {
            size_t len = bytes - off < gap ? bytes - off : gap;
            CPM_CUDA(ctx, cudaMemcpyAsync((char*)dst + off, s + off, len, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        return CPM_OK;
    }
    CPM_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return CPM_OK;
}

namespace {
__global__ void fill_u32_kernel(uint32_t* p, uint32_t v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void scatter_fill_u32_kernel(uint32_t* p, const uint32_t* idx, size_t n, uint32_t v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[idx[i]] = v;
}
}  // namespace

}  // extern "C"

namespace {
// mixKernel (ugc/cl/buffermixer.cl:37-48): out = mix(x, y, a) = x + (y - x) * a; integer formats go through
// float and back with OpenCL's default float->int conversion (round toward zero)
template <typename T>
__global__ void __launch_bounds__(256) mix_kernel(const T* __restrict__ x, const T* __restrict__ y, float a, size_t n,
                                                  T* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float fx = (float)x[i], fy = (float)y[i];
    float m = fmaf(fy - fx, a, fx);   // mix(): one definition on the whole path (grid.cu mix4, the oracle's v4_mix)
    if (sizeof(T) == 4) {
        out[i] = (T)m;
    } else {
        const float hi = sizeof(T) == 1 ? 255.0f : 65535.0f;
        out[i] = (T)(int)cpm_clamp(truncf(m), 0.0f, hi);
    }
}
// volume_mix.frag of VolumeSequencePlayer: integer volumes are NORMALISED textures there -- the shader mixes v / max and
// the render target converts back with round-to-nearest (OpenGL 4.x 2.3.5.2), unlike mixKernel's truncation
template <typename T>
__global__ void __launch_bounds__(256) mix_unorm_kernel(const T* __restrict__ x, const T* __restrict__ y, float a, size_t n,
                                                        T* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float hi = sizeof(T) == 1 ? 255.0f : 65535.0f;
    float fx = (float)x[i] / hi, fy = (float)y[i] / hi;
    float m = fmaf(fy - fx, a, fx);
    out[i] = (T)(int)rintf(cpm_clamp(m, 0.0f, 1.0f) * hi);
}
}  // namespace

extern "C" {

int cpm_mix_unorm(cpm_ctx* ctx, const void* x, const void* y, float a, size_t n, int format, void* out) {
    if (format == CPM_FMT_F32) return cpm_mix(ctx, x, y, a, n, format, out);
    if (!ctx) return CPM_E_INVALID;
    if (n == 0) return CPM_OK;
    CPM_REQUIRE(ctx, x && y && out, "null pointer");
    unsigned grid = cpm_div_up(n, 256);
    if (format == CPM_FMT_U8)
        CPM_LAUNCH(ctx, mix_unorm_kernel<unsigned char>, grid, 256, 0, (const unsigned char*)x, (const unsigned char*)y, a, n, (unsigned char*)out);
    else if (format == CPM_FMT_U16)
        CPM_LAUNCH(ctx, mix_unorm_kernel<unsigned short>, grid, 256, 0, (const unsigned short*)x, (const unsigned short*)y, a, n, (unsigned short*)out);
    else
        return cpm_fail(ctx, CPM_E_INVALID, "cpm_mix_unorm: unknown format");
    return CPM_OK;
}

int cpm_mix(cpm_ctx* ctx, const void* x, const void* y, float a, size_t n, int format, void* out) {
    if (!ctx) return CPM_E_INVALID;
    if (n == 0) return CPM_OK;
    CPM_REQUIRE(ctx, x && y && out, "null pointer");
    unsigned grid = cpm_div_up(n, 256);
    switch (format) {
        case CPM_FMT_U8:
            CPM_LAUNCH(ctx, mix_kernel<unsigned char>, grid, 256, 0, (const unsigned char*)x, (const unsigned char*)y, a, n,
                       (unsigned char*)out);
            break;
        case CPM_FMT_U16:
            CPM_LAUNCH(ctx, mix_kernel<unsigned short>, grid, 256, 0, (const unsigned short*)x, (const unsigned short*)y, a, n,
                       (unsigned short*)out);
            break;
        case CPM_FMT_F32:
            CPM_LAUNCH(ctx, mix_kernel<float>, grid, 256, 0, (const float*)x, (const float*)y, a, n, (float*)out);
            break;
        default:
            return cpm_fail(ctx, CPM_E_INVALID, "cpm_mix: unknown format");
    }
    return CPM_OK;
}

int cpm_mem_scatter_fill_u32(cpm_ctx* ctx, void* dst, const uint32_t* indices, size_t n, uint32_t value) {
    if (!ctx) return CPM_E_INVALID;
    if (n == 0) return CPM_OK;
    CPM_REQUIRE(ctx, dst && indices, "null pointer");
    CPM_LAUNCH(ctx, scatter_fill_u32_kernel, cpm_div_up(n, 256), 256, 0, (uint32_t*)dst, indices, n, value);
    return CPM_OK;
}

int cpm_mem_fill_u32(cpm_ctx* ctx, void* dst, uint32_t value, size_t count) {
    if (!ctx) return CPM_E_INVALID;
    if (count == 0) return CPM_OK;
    CPM_REQUIRE(ctx, dst != nullptr, "null pointer");
    uint8_t b = (uint8_t)value;
    if (value == (0x01010101u * b)) {
        CPM_CUDA(ctx, cudaMemsetAsync(dst, b, count * 4, ctx->stream));
    } else {
        CPM_LAUNCH(ctx, fill_u32_kernel, cpm_div_up(count, 256), 256, 0, (uint32_t*)dst, value, count);
    }
    return CPM_OK;
}

}  // extern "C"
