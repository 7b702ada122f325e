// ctx.cu -- context lifetime, error strings, volume objects.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

static thread_local std::string g_create_error;

int cpm_fail(cpm_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx)
        ctx->err = buf;
    else
        g_create_error = buf;
    return code;
}

int cpm_scratch(cpm_ctx* ctx, size_t bytes, void** out) {
    if (bytes > ctx->scratch_bytes) {
        // Growth happens only when a larger problem arrives; wait for in-flight users.
        CPM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->scratch) CPM_CUDA(ctx, cudaFree(ctx->scratch));
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
        size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
        CPM_CUDA(ctx, cudaMalloc(&ctx->scratch, want));
        ctx->scratch_bytes = want;
    }
    *out = ctx->scratch;
    return CPM_OK;
}

extern "C" {

const char* cpm_version(void) { return "cpm_b200 0.1 (sm_100a)"; }

int cpm_ctx_create(int device, void* stream, cpm_ctx** out) {
    if (!out) return cpm_fail(nullptr, CPM_E_INVALID, "cpm_ctx_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return cpm_fail(nullptr, CPM_E_NO_DEVICE,
                        "cpm_ctx_create: no CUDA device (%s); this library has no CPU fallback",
                        cudaGetErrorString(e));
    if (device < 0 || device >= count)
        return cpm_fail(nullptr, CPM_E_INVALID, "cpm_ctx_create: device %d out of range", device);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return cpm_fail(nullptr, CPM_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return cpm_fail(nullptr, CPM_E_NO_DEVICE,
                        "cpm_ctx_create: device %d is sm_%d%d; this build contains sm_100a code only",
                        device, prop.major, prop.minor);
    if ((e = cudaSetDevice(device)) != cudaSuccess)
        return cpm_fail(nullptr, CPM_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    cpm_ctx* c = new cpm_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
            delete c;
            return cpm_fail(nullptr, CPM_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
        }
        c->own_stream = true;
    }
    {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        (void)cudaGetLastError();
    }
    if ((e = cudaMallocHost((void**)&c->pinned, 256)) != cudaSuccess) {
        if (c->own_stream) cudaStreamDestroy(c->stream);
        delete c;
        return cpm_fail(nullptr, CPM_E_NOMEM, "cudaMallocHost: %s", cudaGetErrorString(e));
    }
    *out = c;
    return CPM_OK;
}

void cpm_ctx_destroy(cpm_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->select_done) cudaEventDestroy(ctx->select_done);
    if (ctx->trace_cursor) cudaFree(ctx->trace_cursor);
    if (ctx->walk_buf) cudaFree(ctx->walk_buf);
    if (ctx->d2h_stream) {
        cudaStreamSynchronize(ctx->d2h_stream);
        cudaStreamDestroy(ctx->d2h_stream);
        if (ctx->d2h_fence) cudaEventDestroy(ctx->d2h_fence);
    }
    if (ctx->xfer_stream) {
        cudaStreamSynchronize(ctx->xfer_stream);
        cudaStreamDestroy(ctx->xfer_stream);
    }
    if (ctx->xfer_fence) cudaEventDestroy(ctx->xfer_fence);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

void* cpm_ctx_stream(cpm_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int cpm_ctx_sync(cpm_ctx* ctx) {
    if (!ctx) return CPM_E_INVALID;
    CPM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CPM_OK;
}

const char* cpm_last_error(cpm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

uint64_t cpm_ctx_launch_count(cpm_ctx* ctx, int reset) {
    if (!ctx) return 0;
    uint64_t n = ctx->launches;
    if (reset) ctx->launches = 0;
    return n;
}

// ---- events ----------------------------------------------------------------------------
int cpm_event_create(cpm_ctx* ctx, cpm_event** out) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, out, "null argument");
    cpm_event* e = new cpm_event();
    cudaError_t rc = cudaEventCreate(&e->ev);
    if (rc != cudaSuccess) {
        delete e;
        return cpm_fail(ctx, CPM_E_CUDA, "cudaEventCreate: %s", cudaGetErrorString(rc));
    }
    *out = e;
    return CPM_OK;
}

int cpm_event_record(cpm_ctx* ctx, cpm_event* ev) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, ev, "null argument");
    CPM_CUDA(ctx, cudaEventRecord(ev->ev, ctx->stream));
    return CPM_OK;
}

int cpm_event_elapsed_ms(cpm_ctx* ctx, cpm_event* begin, cpm_event* end, float* ms_host) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, begin && end && ms_host, "null argument");
    CPM_CUDA(ctx, cudaEventSynchronize(end->ev));
    CPM_CUDA(ctx, cudaEventElapsedTime(ms_host, begin->ev, end->ev));
    return CPM_OK;
}

void cpm_event_destroy(cpm_ctx* ctx, cpm_event* ev) {
    (void)ctx;
    if (!ev) return;
    cudaEventDestroy(ev->ev);
    delete ev;
}

// ---- volumes ---------------------------------------------------------------------------

static size_t voxel_bytes(int format) { return format == CPM_FMT_U8 ? 1 : (format == CPM_FMT_U16 ? 2 : 4); }

static int volume_upload(cpm_ctx* ctx, cpm_volume* v, const void* data) {
    if (v->layout == CPM_VOLUME_LINEAR) {
        v->linear = data;
        return CPM_OK;
    }
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof(p));
    size_t vb = voxel_bytes(v->format);
    p.srcPtr = make_cudaPitchedPtr(const_cast<void*>(data), (size_t)v->dims[0] * vb, v->dims[0], v->dims[1]);
    p.dstArray = v->array;
    p.extent = make_cudaExtent(v->dims[0], v->dims[1], v->dims[2]);
    p.kind = cudaMemcpyDeviceToDevice;
    CPM_CUDA(ctx, cudaMemcpy3DAsync(&p, ctx->stream));
    return CPM_OK;
}

int cpm_volume_create(cpm_ctx* ctx, const void* data, const int dims[3], int format, float format_scale,
                      float format_offset, int layout, cpm_volume** out) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, out && data && dims, "null argument");
    CPM_REQUIRE(ctx, dims[0] > 0 && dims[1] > 0 && dims[2] > 0, "dims must be positive");
    CPM_REQUIRE(ctx, format == CPM_FMT_U8 || format == CPM_FMT_U16 || format == CPM_FMT_F32, "unknown format");
    CPM_REQUIRE(ctx, layout == CPM_VOLUME_LINEAR || layout == CPM_VOLUME_TEXTURE, "unknown layout");
    cpm_volume* v = new cpm_volume();
    memset(v, 0, sizeof(*v));
    v->dims[0] = dims[0];
    v->dims[1] = dims[1];
    v->dims[2] = dims[2];
    v->format = format;
    v->scale = format_scale;
    v->offset = format_offset;
    v->layout = layout;
    if (layout == CPM_VOLUME_TEXTURE) {
        if (dims[2] > 2048 || dims[0] > 32768 || dims[1] > 32768) {
            delete v;
            return cpm_fail(ctx, CPM_E_INVALID, "cpm_volume_create: TEXTURE layout supports at most 32768x32768x2048");
        }
        cudaChannelFormatDesc cd = format == CPM_FMT_U8    ? cudaCreateChannelDesc<unsigned char>()
                                   : format == CPM_FMT_U16 ? cudaCreateChannelDesc<unsigned short>()
                                                           : cudaCreateChannelDesc<float>();
        // The gather flag is documented for plain 2-D arrays; tld4.a2d itself only needs a layered array.  Drivers
        // that reject the combination do so for every array: ask once per process, then go straight to the plain
        // layered array.
        static bool gather_flag_ok = true;
        cudaError_t e = cudaErrorInvalidValue;
        if (gather_flag_ok)
            e = cudaMalloc3DArray(&v->array, &cd, make_cudaExtent(dims[0], dims[1], dims[2]),
                                  cudaArrayLayered | cudaArrayTextureGather);
        if (e == cudaErrorInvalidValue) {
            if (gather_flag_ok) (void)cudaGetLastError();
            gather_flag_ok = false;
            e = cudaMalloc3DArray(&v->array, &cd, make_cudaExtent(dims[0], dims[1], dims[2]), cudaArrayLayered);
            if (e == cudaSuccess && getenv("CPM_DEBUG")) fprintf(stderr, "cpm: layered array created without gather flag\n");
        }
        if (e != cudaSuccess) {
            delete v;
            return cpm_fail(ctx, e == cudaErrorMemoryAllocation ? CPM_E_NOMEM : CPM_E_CUDA, "cudaMalloc3DArray: %s",
                            cudaGetErrorString(e));
        }
        cudaResourceDesc rd;
        memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = v->array;
        cudaTextureDesc td;
        memset(&td, 0, sizeof(td));
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        e = cudaCreateTextureObject(&v->tex, &rd, &td, nullptr);
        if (e != cudaSuccess) {
            cudaFreeArray(v->array);
            delete v;
            return cpm_fail(ctx, CPM_E_CUDA, "cudaCreateTextureObject: %s", cudaGetErrorString(e));
        }
    }
    int rc = volume_upload(ctx, v, data);
    if (rc != CPM_OK) {
        cpm_volume_destroy(ctx, v);
        return rc;
    }
    *out = v;
    return CPM_OK;
}

int cpm_volume_update(cpm_ctx* ctx, cpm_volume* vol, const void* data) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, vol && data, "null argument");
    return volume_upload(ctx, vol, data);
}

void cpm_volume_destroy(cpm_ctx* ctx, cpm_volume* vol) {
    if (!vol) return;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    if (vol->tex) cudaDestroyTextureObject(vol->tex);
    if (vol->array) cudaFreeArray(vol->array);
    delete vol;
}

}  // extern "C"
