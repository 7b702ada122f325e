// comm.cu -- multi-GPU behind the C ABI: one communicator per process (= per GPU), so that a C or
// C++ host (the Inviwo application with one process per GPU) runs the sharded photon path without Python or torch.
//
//  cpm_comm_unique_id / cpm_comm_init      NCCL bootstrap: rank 0 makes the 128-byte id, the host application hands it
//                                          to every rank by any means it likes, every rank calls cpm_comm_init
//  cpm_comm_init_nccl                      adopt a communicator the host already owns (ncclComm_t)
//  cpm_allreduce_lightvol                  sum of the per-GPU light volumes (option B of 8e), out of place
//  cpm_allgather_photons                   every rank's photon records in rank order (option A of 8e)
//  cpm_allgather_volume                    sharded ingest: rank r uploaded slab r of a time step; afterwards every rank
//                                          holds the whole volume (8e: "broadcast once per time step over NVLink")
//
// Transport.  NCCL is resolved at run time (dlopen("libnccl.so.2")) -- the library has no link-time dependency on it and
// single-GPU users never load it.  For the light-volume sum the communicator additionally sets up SYMMETRIC staging
// buffers (cudaMalloc + CUDA IPC handles exchanged through one NCCL all-gather) and then uses the library's own kernel
// over NVLink peer memory (exchange.cu: rank r reduces slice r with peer loads in rank order and stores it to every
// peer), bracketed by a flag barrier in the same symmetric memory; CPM_COMM_TRANSPORT=nccl keeps everything on NCCL.
// Every call is asynchronous on the context's stream, like the rest of the ABI.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

// the few NCCL declarations used (nccl.h of NCCL 2.x; the ABI of these entry points has been stable since 2.0)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess_ = 0, ncclUint8_ = 1, ncclFloat32_ = 7, ncclSum_ = 0 };

namespace {

struct NcclApi {
    void* so = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*) = nullptr;   // NCCL >= 2.18, optional
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;
};

NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char* names[] = {getenv("CPM_NCCL_LIBRARY"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n || !*n) continue;
        api.so = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.so) break;
    }
    if (!api.so) {
        api.why = "libnccl.so.2 not found (set CPM_NCCL_LIBRARY or LD_LIBRARY_PATH)";
        return api;
    }
#define CPM_SYM(field, name)                                              \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.so, name)); \
    if (!api.field) api.why = std::string("symbol missing in libnccl: ") + name;
    CPM_SYM(GetUniqueId, "ncclGetUniqueId")
    CPM_SYM(CommInitRank, "ncclCommInitRank")
    CPM_SYM(CommDestroy, "ncclCommDestroy")
    CPM_SYM(AllReduce, "ncclAllReduce")
    CPM_SYM(AllGather, "ncclAllGather")
    CPM_SYM(GetErrorString, "ncclGetErrorString")
#undef CPM_SYM
    api.CommSplit = reinterpret_cast<decltype(api.CommSplit)>(dlsym(api.so, "ncclCommSplit"));
    return api;
}

// ---- flag barrier over symmetric memory ---------------------------------------------------------------------------------
// flags[r] of rank q's pad is written by rank r.  One CTA per GPU: thread r < world posts `epoch` to peer r's pad, then
// waits until peer r has posted it here.  Every rank launches the kernel in the same stream position, so all world
// kernels are resident together (one small CTA each on different GPUs).
struct BarrierArgs {
    unsigned* pads[CPM_MAX_PEERS];   // this GPU's mapping of every rank's pad
    int rank, world;
    unsigned epoch;
};
__global__ void barrier_kernel(const BarrierArgs A) {
    const int r = threadIdx.x;
    if (r < A.world) {
        __threadfence_system();   // what this GPU wrote before the barrier is visible to the peers first
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(A.pads[r] + A.rank), "r"(A.epoch) : "memory");
        unsigned v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(A.pads[A.rank] + r) : "memory");
        } while ((int)(v - A.epoch) < 0);
    }
}

// ---- global (cross-shard) selection ------------------------------------------------------------------------------------
// out[b] = number of keys below prefix + b * 2^shift (b = 0 ... 256) in an ascending array: one binary search per thread
__global__ void __launch_bounds__(288) probe_counts_kernel(const uint32_t* __restrict__ sorted, unsigned long long n,
                                                           unsigned long long prefix, int shift, uint32_t* __restrict__ out) {
    const unsigned b = threadIdx.x;
    if (b > 256u) return;
    const unsigned long long T = prefix + ((unsigned long long)b << shift);
    unsigned long long lo = 0, hi = n;
    if (T > 0xffffffffull) {
        lo = n;
    } else {
        const uint32_t t32 = (uint32_t)T;
        while (lo < hi) {
            const unsigned long long mid = (lo + hi) >> 1;
            if (sorted[mid] < t32) lo = mid + 1; else hi = mid;
        }
    }
    out[b] = (uint32_t)lo;
}

}  // namespace

struct cpm_comm {
    cpm_ctx* ctx = nullptr;
    ncclComm_t nccl = nullptr;
    ncclComm_t nccl_ingest = nullptr;   // second communicator for the transfer stream (ncclCommSplit), else == nccl
    bool own_nccl = false;
    int rank = 0, world = 1;
    bool want_peer = true;
    // symmetric staging for the light-volume sum: two buffers (alternating, so that a result stays valid while the next
    // frame is being exchanged is the CALLER's business; here they alternate to keep a late reader of call k away from
    // the writers of call k+1) and one flag pad
    size_t sym_floats = 0;
    float* sym_local[2] = {nullptr, nullptr};
    float* sym_peers[2][CPM_MAX_PEERS] = {};
    unsigned* pad_local = nullptr;
    unsigned* pads[CPM_MAX_PEERS] = {};
    unsigned epoch = 0;
    int turn = 0;
    bool peer_ok = false, peer_tried = false;
    std::string transport = "nccl";
};

namespace {

int nccl_check(cpm_ctx* ctx, ncclResult_t r, const char* what) {
    if (r == ncclSuccess_) return CPM_OK;
    return cpm_fail(ctx, CPM_E_CUDA, "%s: NCCL error %d (%s)", what, r, nccl().GetErrorString ? nccl().GetErrorString(r) : "?");
}

void close_peer(cpm_comm* c) {
    for (int b = 0; b < 2; ++b) {
        for (int r = 0; r < c->world; ++r)
            if (r != c->rank && c->sym_peers[b][r]) cudaIpcCloseMemHandle(c->sym_peers[b][r]);
        if (c->sym_local[b]) cudaFree(c->sym_local[b]);
        c->sym_local[b] = nullptr;
        memset(c->sym_peers[b], 0, sizeof(c->sym_peers[b]));
    }
    for (int r = 0; r < c->world; ++r)
        if (r != c->rank && c->pads[r]) cudaIpcCloseMemHandle(c->pads[r]);
    if (c->pad_local) cudaFree(c->pad_local);
    c->pad_local = nullptr;
    memset(c->pads, 0, sizeof(c->pads));
    c->sym_floats = 0;
    c->peer_ok = false;
}

// allocate one symmetric buffer: cudaMalloc here, IPC handles all-gathered over NCCL, peers opened.
// Collective; returns false (with everything released) when any rank failed.
bool symmetric_alloc(cpm_comm* c, size_t bytes, void** local, void** peers /* [world] */) {
    cpm_ctx* ctx = c->ctx;
    *local = nullptr;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    int ok = cudaMalloc(local, bytes) == cudaSuccess && cudaMemsetAsync(*local, 0, bytes, ctx->stream) == cudaSuccess &&
             cudaIpcGetMemHandle(&mine, *local) == cudaSuccess;
    // [handle | ok flag] per rank through one all-gather of bytes
    const size_t rec = sizeof(cudaIpcMemHandle_t) + 8;
    std::vector<unsigned char> host(rec * c->world, 0);
    unsigned char* dev = nullptr;
    if (cudaMalloc((void**)&dev, rec * c->world) != cudaSuccess) return false;
    unsigned char me[sizeof(cudaIpcMemHandle_t) + 8] = {0};
    memcpy(me, &mine, sizeof(mine));
    me[sizeof(mine)] = (unsigned char)ok;
    cudaMemcpyAsync(dev + rec * c->rank, me, rec, cudaMemcpyHostToDevice, ctx->stream);
    bool good = nccl().AllGather(dev + rec * c->rank, dev, rec, ncclUint8_, c->nccl, ctx->stream) == ncclSuccess_;
    good = good && cudaMemcpyAsync(host.data(), dev, rec * c->world, cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess;
    good = good && cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    cudaFree(dev);
    for (int r = 0; good && r < c->world; ++r) good = host[rec * r + sizeof(mine)] != 0;
    if (good) {
        for (int r = 0; r < c->world; ++r) {
            if (r == c->rank) {
                peers[r] = *local;
                continue;
            }
            cudaIpcMemHandle_t h;
            memcpy(&h, &host[rec * r], sizeof(h));
            if (cudaIpcOpenMemHandle(&peers[r], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                good = false;
                peers[r] = nullptr;
            }
        }
    }
    // agree on the outcome (a rank that could not open a peer must not leave the others using the peer path)
    int* flag = nullptr;
    if (cudaMalloc((void**)&flag, 4 * (c->world + 1)) == cudaSuccess) {
        int v = good ? 1 : 0;
        cudaMemcpyAsync(flag + c->world, &v, 4, cudaMemcpyHostToDevice, ctx->stream);
        std::vector<int> all(c->world, 0);
        if (nccl().AllGather(flag + c->world, flag, 4, ncclUint8_, c->nccl, ctx->stream) == ncclSuccess_ &&
            cudaMemcpyAsync(all.data(), flag, 4 * c->world, cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
            cudaStreamSynchronize(ctx->stream) == cudaSuccess) {
            for (int r = 0; r < c->world; ++r) good = good && all[r] == 1;
        } else {
            good = false;
        }
        cudaFree(flag);
    } else {
        good = false;
    }
    if (!good) {
        for (int r = 0; r < c->world; ++r)
            if (r != c->rank && peers[r]) cudaIpcCloseMemHandle(peers[r]);
        if (*local) cudaFree(*local);
        *local = nullptr;
        for (int r = 0; r < c->world; ++r) peers[r] = nullptr;
        cudaGetLastError();
    }
    return good;
}

bool ensure_peer(cpm_comm* c, size_t n_floats) {
    if (!c->want_peer || c->world > CPM_MAX_PEERS) return false;
    if (c->peer_ok && c->sym_floats >= n_floats) return true;
    if (c->peer_tried && !c->peer_ok) return false;   // set-up failed once on this node: stay on NCCL
    c->peer_tried = true;
    close_peer(c);
    void* peers[CPM_MAX_PEERS] = {};
    void* local = nullptr;
    bool ok = true;
    for (int b = 0; ok && b < 2; ++b) {
        ok = symmetric_alloc(c, n_floats * sizeof(float), &local, peers);
        if (ok) {
            c->sym_local[b] = (float*)local;
            for (int r = 0; r < c->world; ++r) c->sym_peers[b][r] = (float*)peers[r];
        }
    }
    if (ok) {
        ok = symmetric_alloc(c, 256, &local, peers);
        if (ok) {
            c->pad_local = (unsigned*)local;
            for (int r = 0; r < c->world; ++r) c->pads[r] = (unsigned*)peers[r];
        }
    }
    if (!ok) {
        close_peer(c);
        return false;
    }
    c->sym_floats = n_floats;
    c->peer_ok = true;
    c->epoch = 0;
    c->transport = "peer kernel over CUDA IPC symmetric memory";
    return true;
}

int device_barrier(cpm_comm* c) {
    BarrierArgs a;
    memset(&a, 0, sizeof(a));
    for (int r = 0; r < c->world; ++r) a.pads[r] = c->pads[r];
    a.rank = c->rank;
    a.world = c->world;
    a.epoch = ++c->epoch;
    CPM_LAUNCH(c->ctx, barrier_kernel, 1, 32, 0, a);
    return CPM_OK;
}

}  // namespace

extern "C" {

int cpm_comm_unique_id(void* id_out) {
    if (!id_out) return CPM_E_INVALID;
    NcclApi& n = nccl();
    if (!n.GetUniqueId) return CPM_E_UNSUPPORTED;
    ncclUniqueId id;
    if (n.GetUniqueId(&id) != ncclSuccess_) return CPM_E_CUDA;
    memcpy(id_out, &id, sizeof(id));
    return CPM_OK;
}

static int comm_make(cpm_ctx* ctx, ncclComm_t nc, bool own, int rank, int world, cpm_comm** out) {
    cpm_comm* c = new cpm_comm();
    c->ctx = ctx;
    c->nccl = nc;
    c->own_nccl = own;
    c->rank = rank;
    c->world = world;
    const char* t = getenv("CPM_COMM_TRANSPORT");
    c->want_peer = !(t && strcmp(t, "nccl") == 0);
    ctx->comm = c;
    *out = c;
    return CPM_OK;
}

int cpm_comm_init(cpm_ctx* ctx, const void* unique_id, int rank, int world, cpm_comm** out) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, unique_id && out, "null argument");
    CPM_REQUIRE(ctx, world >= 1 && rank >= 0 && rank < world, "rank / world out of range");
    NcclApi& n = nccl();
    if (!n.CommInitRank || !n.why.empty()) return cpm_fail(ctx, CPM_E_UNSUPPORTED, "cpm_comm_init: %s", n.why.c_str());
    CPM_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    ncclComm_t nc = nullptr;
    int rc = nccl_check(ctx, n.CommInitRank(&nc, world, id, rank), "ncclCommInitRank");
    if (rc != CPM_OK) return rc;
    return comm_make(ctx, nc, true, rank, world, out);
}

int cpm_comm_init_nccl(cpm_ctx* ctx, void* nccl_comm, int rank, int world, cpm_comm** out) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, nccl_comm && out, "null argument");
    CPM_REQUIRE(ctx, world >= 1 && rank >= 0 && rank < world, "rank / world out of range");
    NcclApi& n = nccl();
    if (!n.AllReduce || !n.why.empty()) return cpm_fail(ctx, CPM_E_UNSUPPORTED, "cpm_comm_init_nccl: %s", n.why.c_str());
    return comm_make(ctx, (ncclComm_t)nccl_comm, false, rank, world, out);
}

void cpm_comm_destroy(cpm_comm* c) {
    if (!c) return;
    cudaStreamSynchronize(c->ctx->stream);
    close_peer(c);
    if (c->nccl_ingest && c->nccl_ingest != c->nccl && nccl().CommDestroy) nccl().CommDestroy(c->nccl_ingest);
    if (c->own_nccl && c->nccl && nccl().CommDestroy) nccl().CommDestroy(c->nccl);
    if (c->ctx->comm == c) c->ctx->comm = nullptr;
    delete c;
}

int cpm_comm_rank(const cpm_comm* c) { return c ? c->rank : -1; }
int cpm_comm_world(const cpm_comm* c) { return c ? c->world : 0; }
const char* cpm_comm_transport(const cpm_comm* c) { return c ? c->transport.c_str() : ""; }

// The sum in two halves, so that a caller can pipeline it: _begin takes the snapshot of `local` (after the work already
// submitted to the stream; once an event recorded after _begin has completed, `local` may be overwritten), _end forms the
// sum in sum_out.  cpm_allreduce_lightvol = _begin + _end.
int cpm_allreduce_lightvol_begin(cpm_comm* c, const float* local, float* sum_out, size_t n_floats) {
    if (!c) return CPM_E_INVALID;
    cpm_ctx* ctx = c->ctx;
    CPM_REQUIRE(ctx, local && sum_out, "null argument");
    if (n_floats == 0) return CPM_OK;
    float* dst = sum_out;
    if (c->world > 1 && n_floats % 4 == 0 && ensure_peer(c, n_floats)) dst = c->sym_local[c->turn];
    if (dst != local) CPM_CUDA(ctx, cudaMemcpyAsync(dst, local, n_floats * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    return CPM_OK;
}

int cpm_allreduce_lightvol_end(cpm_comm* c, float* sum_out, size_t n_floats) {
    if (!c) return CPM_E_INVALID;
    cpm_ctx* ctx = c->ctx;
    CPM_REQUIRE(ctx, sum_out != nullptr, "null argument");
    if (n_floats == 0 || c->world == 1) return CPM_OK;
    if (n_floats % 4 == 0 && c->peer_ok && c->sym_floats >= n_floats) {
        const int b = c->turn;
        c->turn ^= 1;
        int rc = device_barrier(c);                 // every rank's snapshot is in place
        if (rc != CPM_OK) return rc;
        rc = cpm_allreduce_peer_f32(ctx, c->sym_peers[b], nullptr, n_floats, c->rank, c->world, 0);
        if (rc != CPM_OK) return rc;
        rc = device_barrier(c);                     // every rank's slice is stored everywhere
        if (rc != CPM_OK) return rc;
        CPM_CUDA(ctx, cudaMemcpyAsync(sum_out, c->sym_local[b], n_floats * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        return CPM_OK;
    }
    c->transport = "nccl";
    return nccl_check(ctx, nccl().AllReduce(sum_out, sum_out, n_floats, ncclFloat32_, ncclSum_, c->nccl, ctx->stream), "ncclAllReduce");
}

int cpm_allreduce_lightvol(cpm_comm* c, const float* local, float* sum_out, size_t n_floats) {
    int rc = cpm_allreduce_lightvol_begin(c, local, sum_out, n_floats);
    return rc == CPM_OK ? cpm_allreduce_lightvol_end(c, sum_out, n_floats) : rc;
}

// A second communicator over the same ranks bound to another context (= another stream) of this process: e.g. the
// exchange on a side stream next to the frame's kernels.  Collective.  Needs ncclCommSplit (NCCL >= 2.18).
int cpm_comm_split(cpm_comm* c, cpm_ctx* other_ctx, cpm_comm** out) {
    if (!c || !other_ctx) return CPM_E_INVALID;
    CPM_REQUIRE(other_ctx, out != nullptr, "null argument");
    if (!nccl().CommSplit) return cpm_fail(other_ctx, CPM_E_UNSUPPORTED, "cpm_comm_split: this NCCL has no ncclCommSplit");
    ncclComm_t nc = nullptr;
    int rc = nccl_check(other_ctx, nccl().CommSplit(c->nccl, 0, c->rank, &nc, nullptr), "ncclCommSplit");
    if (rc != CPM_OK) return rc;
    return comm_make(other_ctx, nc, true, c->rank, c->world, out);
}

int cpm_allgather_photons(cpm_comm* c, const float* local, size_t n_floats_per_rank, float* all_out) {
    if (!c) return CPM_E_INVALID;
    cpm_ctx* ctx = c->ctx;
    CPM_REQUIRE(ctx, local && all_out, "null argument");
    if (n_floats_per_rank == 0) return CPM_OK;
    if (c->world == 1) {
        if (all_out != local) CPM_CUDA(ctx, cudaMemcpyAsync(all_out, local, n_floats_per_rank * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        return CPM_OK;
    }
    return nccl_check(ctx, nccl().AllGather(local, all_out, n_floats_per_rank, ncclFloat32_, c->nccl, ctx->stream), "ncclAllGather");
}

int cpm_allgather_volume(cpm_comm* c, void* volume, size_t slab_bytes) {
    if (!c) return CPM_E_INVALID;
    cpm_ctx* ctx = c->ctx;
    CPM_REQUIRE(ctx, volume, "null argument");
    if (slab_bytes == 0 || c->world == 1) return CPM_OK;
    // in place: NCCL accepts sendbuff == recvbuff + rank * count
    unsigned char* base = (unsigned char*)volume;
    return nccl_check(ctx, nccl().AllGather(base + slab_bytes * c->rank, base, slab_bytes, ncclUint8_, c->nccl, ctx->stream),
                      "ncclAllGather");
}

int cpm_comm_upload_volume_sharded(cpm_comm* c, void* volume, const void* src_host, size_t total_bytes, int on_transfer_stream,
                                   cpm_event** done) {
    if (!c) return CPM_E_INVALID;
    cpm_ctx* ctx = c->ctx;
    CPM_REQUIRE(ctx, volume && src_host && total_bytes > 0, "null argument");
    CPM_REQUIRE(ctx, !on_transfer_stream || done, "the transfer-stream variant returns its completion event");
    const size_t slab = total_bytes / (size_t)c->world;
    CPM_REQUIRE(ctx, slab * (size_t)c->world == total_bytes && slab % 16 == 0, "the volume does not split into equal 16-byte aligned slabs");
    cudaStream_t st = ctx->stream;
    ncclComm_t nc = c->nccl;
    if (on_transfer_stream) {
        if (!ctx->xfer_stream) {
            CPM_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->xfer_stream, cudaStreamNonBlocking));
            CPM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->xfer_fence, cudaEventDisableTiming));
        }
        if (!c->nccl_ingest) {
            // its own communicator, so that the ingest all-gather never queues behind (or in front of) the context
            // stream's collectives inside NCCL
            c->nccl_ingest = c->nccl;
            if (c->world > 1 && nccl().CommSplit) {
                ncclComm_t split = nullptr;
                if (nccl().CommSplit(c->nccl, 0, c->rank, &split, nullptr) == ncclSuccess_ && split) c->nccl_ingest = split;
            }
        }
        // earlier readers / writers of the volume on the context stream finish first
        CPM_CUDA(ctx, cudaEventRecord(ctx->xfer_fence, ctx->stream));
        CPM_CUDA(ctx, cudaStreamWaitEvent(ctx->xfer_stream, ctx->xfer_fence, 0));
        st = ctx->xfer_stream;
        nc = c->nccl_ingest;
    }
    unsigned char* base = (unsigned char*)volume;
    const unsigned char* src = (const unsigned char*)src_host;
    CPM_CUDA(ctx, cudaMemcpyAsync(base + slab * c->rank, src + slab * c->rank, slab, cudaMemcpyHostToDevice, st));
    if (c->world > 1) {
        int rc = nccl_check(ctx, nccl().AllGather(base + slab * c->rank, base, slab, ncclUint8_, nc, st), "ncclAllGather (ingest)");
        if (rc != CPM_OK) return rc;
    }
    if (on_transfer_stream) {
        cpm_event* e = new cpm_event();
        cudaError_t rc = cudaEventCreateWithFlags(&e->ev, cudaEventDisableTiming);
        if (rc != cudaSuccess) {
            delete e;
            return cpm_fail(ctx, CPM_E_CUDA, "cudaEventCreate: %s", cudaGetErrorString(rc));
        }
        CPM_CUDA(ctx, cudaEventRecord(e->ev, st));
        *done = e;
    }
    return CPM_OK;
}

int cpm_comm_barrier(cpm_comm* c) {
    if (!c) return CPM_E_INVALID;
    if (c->world == 1) return CPM_OK;
    if (c->peer_ok) return device_barrier(c);
    void* s = nullptr;
    int rc = cpm_scratch(c->ctx, 64, &s);
    if (rc != CPM_OK) return rc;
    return nccl_check(c->ctx, nccl().AllReduce(s, s, 1, ncclFloat32_, ncclSum_, c->nccl, c->ctx->stream), "ncclAllReduce (barrier)");
}

// ---- host-value collectives and the global selection --------------------------------------------------------------------
int cpm_comm_allgather_u64(cpm_comm* c, const unsigned long long* values, int count, unsigned long long* all_out) {
    if (!c) return CPM_E_INVALID;
    cpm_ctx* ctx = c->ctx;
    CPM_REQUIRE(ctx, values && all_out && count >= 1 && count <= 64, "bad argument");
    if (c->world == 1) {
        memcpy(all_out, values, (size_t)count * sizeof(unsigned long long));
        return CPM_OK;
    }
    const size_t bytes = (size_t)count * sizeof(unsigned long long);
    void* scr = nullptr;
    int rc = cpm_scratch(ctx, bytes * (size_t)(c->world + 1), &scr);
    if (rc != CPM_OK) return rc;
    unsigned char* mine = (unsigned char*)scr;
    unsigned char* all = mine + bytes;
    CPM_CUDA(ctx, cudaMemcpyAsync(mine, values, bytes, cudaMemcpyHostToDevice, ctx->stream));
    rc = nccl_check(ctx, nccl().AllGather(mine, all, bytes, ncclUint8_, c->nccl, ctx->stream), "ncclAllGather");
    if (rc != CPM_OK) return rc;
    CPM_CUDA(ctx, cudaMemcpyAsync(all_out, all, bytes * (size_t)c->world, cudaMemcpyDeviceToHost, ctx->stream));
    CPM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CPM_OK;
}

int cpm_comm_select_global(cpm_comm* c, const uint32_t* sorted_keys, size_t n_local, unsigned long long position,
                           unsigned long long* local_count) {
    if (!c) return CPM_E_INVALID;
    cpm_ctx* ctx = c->ctx;
    CPM_REQUIRE(ctx, local_count && (sorted_keys || n_local == 0), "null argument");
    const int W = c->world;
    void* scr = nullptr;
    int rc = cpm_scratch(ctx, 257 * sizeof(uint32_t) * (size_t)(W + 1), &scr);
    if (rc != CPM_OK) return rc;
    uint32_t* mine = (uint32_t*)scr;
    uint32_t* all = mine + 257;
    std::vector<uint32_t> h((size_t)W * 257);
    unsigned long long prefix = 0;
    std::vector<unsigned long long> less((size_t)W, 0), leq((size_t)W, 0);
    // 256-ary descent over the key bits: after the round with `shift`, prefix holds the bits >= shift of the key T of the
    // element at `position` of the global order (the largest T with #{keys < T} <= position)
    for (int shift = 24; shift >= 0; shift -= 8) {
        CPM_LAUNCH(ctx, probe_counts_kernel, 1, 288, 0, sorted_keys, (unsigned long long)n_local, prefix, shift, mine);
        if (W > 1) {
            rc = nccl_check(ctx, nccl().AllGather(mine, all, 257 * sizeof(uint32_t), ncclUint8_, c->nccl, ctx->stream), "ncclAllGather");
            if (rc != CPM_OK) return rc;
        }
        CPM_CUDA(ctx, cudaMemcpyAsync(h.data(), W > 1 ? all : mine, h.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CPM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        int pick = 0;
        for (int b = 0; b <= 255; ++b) {
            unsigned long long tot = 0;
            for (int r = 0; r < W; ++r) tot += h[(size_t)r * 257 + b];
            if (tot <= position) pick = b; else break;
        }
        prefix += (unsigned long long)pick << shift;
        for (int r = 0; r < W; ++r) {
            less[r] = h[(size_t)r * 257 + pick];
            leq[r] = h[(size_t)r * 257 + pick + 1];
        }
    }
    // every key < T is taken; the keys == T fill what is left of `position` in rank order (the global order is
    // (key, rank, local index): what one stable sort over the concatenated shards gives)
    unsigned long long taken = 0;
    for (int r = 0; r < W; ++r) taken += less[r];
    unsigned long long left = position > taken ? position - taken : 0;
    unsigned long long count = 0;
    for (int r = 0; r < W; ++r) {
        const unsigned long long ties = leq[r] - less[r];
        const unsigned long long take = std::min(ties, left);
        if (r == c->rank) count = less[r] + take;
        left -= take;
    }
    *local_count = count;
    return CPM_OK;
}

}  // extern "C"
