// processors.cpp -- implementation of the Inviwo shim, the kernel-launcher classes and the
// drop-in processors.  Every device operation is one call into the C ABI (include/cpm_b200.h).
#include <atomic>
#include <cctype>
#include <sstream>
#include <fstream>
#include "processors.h"

#include <cstdio>
#include <cstdlib>

namespace inviwo {

void LogError(const std::string& msg) { std::fprintf(stderr, "[cpm host] error: %s\n", msg.c_str()); }
void LogInfo(const std::string& msg) {
    if (std::getenv("CPM_VERBOSE")) std::fprintf(stderr, "[cpm host] %s\n", msg.c_str());
}

// ---------------------------------------------------------------------------------- math --------
mat4 inverse(const mat4& M) {
    // general 4x4 inverse (cofactor expansion), double precision internally
    double a[16], inv[16];
    for (int i = 0; i < 16; ++i) a[i] = M.m[i];
    inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
    inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
    inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
    inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
    inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
    inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
    inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
    inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
    inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
    inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
    inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
    inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
    inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
    inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
    inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
    inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
    double det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
    mat4 R;
    for (int i = 0; i < 16; ++i) R.m[i] = (float)(inv[i] / det);
    return R;
}

// ------------------------------------------------------------------------------- runtime --------
static CpmRuntime g_runtime;
CpmRuntime& CpmRuntime::get() {
    if (!g_runtime.ctx_) {
        const char* d = std::getenv("CPM_DEVICE");
        init(d ? std::atoi(d) : 0);
    }
    return g_runtime;
}
void CpmRuntime::init(int device, void* stream) {
    if (g_runtime.ctx_) return;
    cpm_ctx* c = nullptr;
    int rc = cpm_ctx_create(device, stream, &c);
    if (rc != CPM_OK) throw CpmError(rc, cpm_last_error(nullptr));
    g_runtime.ctx_ = c;
}
void CpmRuntime::shutdown() {
    StageProfiler::get().releaseEvents();
    if (g_runtime.ctx_) cpm_ctx_destroy(g_runtime.ctx_);
    g_runtime.ctx_ = nullptr;
}

// ------------------------------------------------------------------------------ profiler --------
static StageProfiler g_profiler;
StageProfiler& StageProfiler::get() { return g_profiler; }
cpm_event* StageProfiler::take() {
    if (!pool_.empty()) {
        cpm_event* e = pool_.back();
        pool_.pop_back();
        return e;
    }
    cpm_event* e = nullptr;
    CPM_CHECK(cpm_event_create(CpmRuntime::get().ctx(), &e));
    return e;
}
void StageProfiler::begin(const char* stage) {
    if (!enabled) return;
    if (!only.empty() && only != stage) {   // filtered out: keep the nesting, record nothing
        open_.emplace_back(stage, nullptr);
        return;
    }
    cpm_event* e = take();
    CPM_CHECK(cpm_event_record(CpmRuntime::get().ctx(), e));
    open_.emplace_back(stage, e);
}
void StageProfiler::end() {
    if (!enabled || open_.empty()) return;
    if (!open_.back().second) {
        open_.pop_back();
        return;
    }
    cpm_event* b = take();
    CPM_CHECK(cpm_event_record(CpmRuntime::get().ctx(), b));
    pending_.push_back({open_.back().first, open_.back().second, b});
    open_.pop_back();
    if (pending_.size() >= 4096) resolve();
}
void StageProfiler::resolve() {
    std::map<std::string, double> thisRound;
    for (auto& p : pending_) {
        float ms = 0.f;
        CPM_CHECK(cpm_event_elapsed_ms(CpmRuntime::get().ctx(), p.a, p.b, &ms));
        Acc& a = acc_[p.stage];
        a.total += ms;
        a.n += 1;
        a.last = ms;
        pool_.push_back(p.a);
        pool_.push_back(p.b);
    }
    pending_.clear();
}
void StageProfiler::reset() {
    resolve();
    for (auto& o : open_)
        if (o.second) pool_.push_back(o.second);   // stages left open by an exception
    open_.clear();
    acc_.clear();
}
double StageProfiler::totalMs(const std::string& s) { resolve(); auto it = acc_.find(s); return it == acc_.end() ? 0.0 : it->second.total; }
double StageProfiler::lastMs(const std::string& s) { resolve(); auto it = acc_.find(s); return it == acc_.end() ? 0.0 : it->second.last; }
int StageProfiler::count(const std::string& s) { resolve(); auto it = acc_.find(s); return it == acc_.end() ? 0 : it->second.n; }
std::string StageProfiler::stages() {
    resolve();
    std::string out;
    for (auto& kv : acc_) out += (out.empty() ? "" : ",") + kv.first;
    return out;
}
void StageProfiler::releaseEvents() {
    cpm_ctx* c = g_runtime.ctx();
    for (auto& p : pending_) { cpm_event_destroy(c, p.a); cpm_event_destroy(c, p.b); }
    pending_.clear();
    for (auto& o : open_) cpm_event_destroy(c, o.second);
    open_.clear();
    for (auto* e : pool_) cpm_event_destroy(c, e);
    pool_.clear();
}

// ------------------------------------------------------------------------------- buffers --------
void BufferBase::releaseDevice() {
    if (dev_ && g_runtime.ctx()) cpm_mem_free(g_runtime.ctx(), dev_);
    dev_ = nullptr;
    devBytes_ = 0;
}
void BufferBase::ensureDevice() {
    size_t bytes = getSizeInBytes();
    if (bytes != devBytes_) {
        releaseDevice();
        if (bytes) CPM_CHECK(cpm_mem_alloc(CpmRuntime::get().ctx(), bytes, &dev_));
        devBytes_ = bytes;
        devValid_ = false;
    }
}
const void* BufferBase::deviceRead() {
    ensureDevice();
    if (!devValid_ && devBytes_) {
        CPM_CHECK(cpm_mem_copy_h2d(CpmRuntime::get().ctx(), dev_, ramPtr(), devBytes_));
        CpmRuntime::get().sync();   // pageable source: complete before the vector can change
        h2dBytes() += devBytes_;
    }
    devValid_ = true;
    return dev_;
}
void* BufferBase::deviceWrite() {
    ensureDevice();
    devValid_ = true;
    ramValid_ = false;
    return dev_;
}
void BufferBase::downloadIfStale() {
    if (!ramValid_ && dev_ && devValid_ && devBytes_ == getSizeInBytes() && devBytes_) {
        CPM_CHECK(cpm_mem_copy_d2h(CpmRuntime::get().ctx(), ramPtr(), dev_, devBytes_));
        CpmRuntime::get().sync();
        d2hBytes() += devBytes_;
    }
    ramValid_ = true;
}

// -------------------------------------------------------------------------------- volume --------
const DataFormatBase* DataFormatBase::get(DataFormatId id) {
    static const DataFormatBase f[4] = {{DataFormatId::UInt8, 1, 1, 255.0},
                                        {DataFormatId::UInt16, 2, 1, 65535.0},
                                        {DataFormatId::Float32, 4, 1, 1.0},
                                        {DataFormatId::Vec4Float32, 16, 4, 1.0}};
    return &f[(int)id];
}

Volume::Volume(size3_t dim, const DataFormatBase* format) : dim_(dim), format_(format) {
    dataMap_.dataRange = {0.0, format->maxValue};
    dataMap_.valueRange = dataMap_.dataRange;
    ram_.assign(getSizeInBytes(), 0);
    touch();
}
Volume::~Volume() {
    cpm_ctx* c = g_runtime.ctx();
    if (prefetchDone_) cpm_event_destroy(c, prefetchDone_);
    if (lin_) cpm_volume_destroy(c, lin_);
    if (tex_) cpm_volume_destroy(c, tex_);
    if (range_ && c) cpm_mem_free(c, range_);
    if (dev_ && c) cpm_mem_free(c, dev_);
}
void Volume::setDimensions(size3_t d) {
    if (d == dim_) return;
    dim_ = d;
    ext_ = nullptr;
    ram_.assign(getSizeInBytes(), 0);
    ramValid_ = true;
    devValid_ = false;
    cpm_ctx* c = g_runtime.ctx();
    if (lin_) { cpm_volume_destroy(c, lin_); lin_ = nullptr; }
    if (tex_) { cpm_volume_destroy(c, tex_); tex_ = nullptr; }
}
void* Volume::getEditableRAMData() {
    getRAMData();
    devValid_ = false;
    texValid_ = false; touch();
    return ext_ ? ext_ : (void*)ram_.data();
}
const void* Volume::getRAMData() {
    void* p = ext_ ? ext_ : (void*)ram_.data();
    if (!ramValid_ && dev_ && devValid_) {
        CPM_CHECK(cpm_mem_copy_d2h(CpmRuntime::get().ctx(), p, dev_, getSizeInBytes()));
        CpmRuntime::get().sync();
        BufferBase::d2hBytes() += getSizeInBytes();
    }
    ramValid_ = true;
    return p;
}
void Volume::setExternalRAMData(void* ptr) {
    ext_ = ptr;
    ramValid_ = true;
    texValid_ = false; touch();
    if (ptr && ptr == prefetched_ && prefetchDone_) {
        // adopt the upload started by prefetchExternalRAMData: the context stream waits for it, nothing is copied
        auto& rt = CpmRuntime::get();
        rt.check(cpm_ctx_wait_event(rt.ctx(), prefetchDone_));
        cpm_event_destroy(rt.ctx(), prefetchDone_);
        prefetchDone_ = nullptr;
        prefetched_ = nullptr;
        devValid_ = true;
        return;
    }
    devValid_ = false;
}
void Volume::prefetchExternalRAMData(void* ptr) {
    ensureDevice();
    auto& rt = CpmRuntime::get();
    if (prefetchDone_) {
        rt.check(cpm_ctx_wait_event(rt.ctx(), prefetchDone_));
        cpm_event_destroy(rt.ctx(), prefetchDone_);
        prefetchDone_ = nullptr;
    }
    if (rt.shardedUpload(devBytes_)) {
        // this rank's slab over PCIe, the other slabs over NVLink, both on the transfer stream
        rt.check(cpm_comm_upload_volume_sharded(rt.comm, dev_, ptr, devBytes_, 1, &prefetchDone_));
        BufferBase::h2dBytes() += devBytes_ / (size_t)cpm_comm_world(rt.comm);
    } else {
        rt.check(cpm_mem_prefetch_h2d(rt.ctx(), dev_, ptr, devBytes_, &prefetchDone_));
        BufferBase::h2dBytes() += devBytes_;
    }
    prefetched_ = ptr;
    devValid_ = false;
    texValid_ = false; touch();
}
void Volume::ensureDevice() {
    size_t bytes = getSizeInBytes();
    if (bytes != devBytes_) {
        cpm_ctx* c = CpmRuntime::get().ctx();
        if (lin_) { cpm_volume_destroy(c, lin_); lin_ = nullptr; }
        if (tex_) { cpm_volume_destroy(c, tex_); tex_ = nullptr; }
        if (dev_) cpm_mem_free(c, dev_);
        dev_ = nullptr;
        if (bytes) CPM_CHECK(cpm_mem_alloc(c, bytes, &dev_));
        devBytes_ = bytes;
        devValid_ = false;
    }
}
const void* Volume::deviceRead() {
    ensureDevice();
    if (!devValid_ && devBytes_) {
        const void* p = ext_ ? ext_ : (const void*)ram_.data();
        ScopedStage st("h2d");
        auto& rt = CpmRuntime::get();
        if (ext_ && rt.shardedUpload(devBytes_)) {
            rt.check(cpm_comm_upload_volume_sharded(rt.comm, dev_, p, devBytes_, 0, nullptr));
            BufferBase::h2dBytes() += devBytes_ / (size_t)cpm_comm_world(rt.comm);
        } else {
            CPM_CHECK(cpm_mem_copy_h2d(rt.ctx(), dev_, p, devBytes_));
            if (!ext_) rt.sync();   // external (pinned) sources stay valid; stream order suffices
            BufferBase::h2dBytes() += devBytes_;
        }
        texValid_ = false; touch();
    }
    devValid_ = true;
    return dev_;
}
void* Volume::deviceWrite() {
    ensureDevice();
    devValid_ = true;
    ramValid_ = false;
    texValid_ = false; touch();
    return dev_;
}
const cpm_volume* Volume::handle(int layout) {
    const void* d = deviceRead();
    cpm_ctx* c = CpmRuntime::get().ctx();
    const int dims[3] = {(int)dim_.x, (int)dim_.y, (int)dim_.z};
    int fmt = format_->id == DataFormatId::UInt8 ? CPM_FMT_U8 : (format_->id == DataFormatId::UInt16 ? CPM_FMT_U16 : CPM_FMT_F32);
    if (format_->id == DataFormatId::Vec4Float32) throw CpmError(CPM_E_UNSUPPORTED, "vec4 volumes cannot be sampled");
    float scale, offset;
    formatScaleOffset(scale, offset);
    if ((lin_ || tex_) && (scale != handleScale_ || offset != handleOffset_)) {
        // dataMap_.dataRange changed since the handles were made (e.g. 12-bit data declared in a 16-bit volume): new
        // handles, and a new data version so that value ranges / opacity bounds derived with the old scaling are rebuilt
        if (lin_) { cpm_volume_destroy(c, lin_); lin_ = nullptr; }
        if (tex_) { cpm_volume_destroy(c, tex_); tex_ = nullptr; }
        touch();
    }
    handleScale_ = scale;
    handleOffset_ = offset;
    if (layout == CPM_VOLUME_LINEAR) {
        if (!lin_) CPM_CHECK(cpm_volume_create(c, d, dims, fmt, scale, offset, CPM_VOLUME_LINEAR, &lin_));
        return lin_;
    }
    if (!tex_) {
        ScopedStage st("texcopy");
        CPM_CHECK(cpm_volume_create(c, d, dims, fmt, scale, offset, CPM_VOLUME_TEXTURE, &tex_));
        texValid_ = true;
    } else if (!texValid_) {
        ScopedStage st("texcopy");
        CPM_CHECK(cpm_volume_update(c, tex_, d));
        texValid_ = true;
    }
    return tex_;
}

void Volume::formatScaleOffset(float& scale, float& offset) const {
    // "Scaling for 12-bit data": normalised = (v/typeMax) * typeMax/(dataRange.y - dataRange.x) - dataRange.x/(range)
    double range = dataMap_.dataRange.y - dataMap_.dataRange.x;
    scale = (float)(format_->maxValue / range);
    offset = (float)(-dataMap_.dataRange.x / format_->maxValue);
}
static std::atomic<uint64_t> g_dataVersion{0};
void Volume::touch() {
    rangeValid_ = false;
    version_ = ++g_dataVersion;
}
const float* Volume::valueRange(int cellLog2, size_t* nCells) {
    const cpm_volume* vh = handle(CPM_VOLUME_LINEAR);   // uploads when the device copy is stale
    cpm_ctx* c = CpmRuntime::get().ctx();
    const int dims[3] = {(int)dim_.x, (int)dim_.y, (int)dim_.z};
    int gd[3];
    CPM_CHECK(cpm_bound_grid_dims(dims, cellLog2, gd));
    const size_t n = (size_t)gd[0] * gd[1] * gd[2];
    if (n != rangeCells_ || cellLog2 != rangeLog2_) {
        if (range_) cpm_mem_free(c, range_);
        range_ = nullptr;
        CPM_CHECK(cpm_mem_alloc(c, n * 2 * sizeof(float), &range_));
        rangeCells_ = n;
        rangeLog2_ = cellLog2;
        rangeValid_ = false;
    }
    if (!rangeValid_) {
        ScopedStage st("range");
        CPM_CHECK(cpm_volume_value_range(c, vh, cellLog2, static_cast<float*>(range_), nullptr));
        rangeValid_ = true;
    }
    if (nCells) *nCells = n;
    return static_cast<const float*>(range_);
}

mat4 StructuredCoordinateTransformer::getTextureToIndexMatrix() const {
    mat4 m;
    size3_t d = v_->getDimensions();
    m[0][0] = (float)d.x; m[1][1] = (float)d.y; m[2][2] = (float)d.z;
    m[3][0] = m[3][1] = m[3][2] = -0.5f;
    return m;
}
mat4 StructuredCoordinateTransformer::getIndexToTextureMatrix() const {
    mat4 m;
    size3_t d = v_->getDimensions();
    m[0][0] = 1.f / (float)d.x; m[1][1] = 1.f / (float)d.y; m[2][2] = 1.f / (float)d.z;
    m[3][0] = 0.5f / (float)d.x; m[3][1] = 0.5f / (float)d.y; m[3][2] = 0.5f / (float)d.z;
    return m;
}
mat4 StructuredCoordinateTransformer::getTextureToWorldMatrix() const { return v_->getWorldMatrix() * v_->getModelMatrix(); }

// ---------------------------------------------------------------------- transfer function --------
void TransferFunction::add(double pos, vec4 color) {
    points_.emplace_back(pos, color);
    std::stable_sort(points_.begin(), points_.end());
    dirty_ = true;
}
void TransferFunction::rasterise() {
    auto* ram = data_.getEditableRAMRepresentation();
    const size_t w = ram->size();
    for (size_t i = 0; i < w; ++i) {
        double x = w > 1 ? (double)i / (double)(w - 1) : 0.0;
        vec4 c(0.f);
        if (!points_.empty()) {
            if (x <= points_.front().getPosition()) {
                c = points_.front().getColor();
            } else if (x >= points_.back().getPosition()) {
                c = points_.back().getColor();
            } else {
                size_t k = 1;
                while (points_[k].getPosition() < x) ++k;
                double p0 = points_[k - 1].getPosition(), p1 = points_[k].getPosition();
                double t = (x - p0) / (p1 - p0);
                vec4 a = points_[k - 1].getColor(), b = points_[k].getColor();
                for (int ch = 0; ch < 4; ++ch) c[ch] = (float)((1.0 - t) * a[ch] + t * b[ch]);
            }
        }
        (*ram)[i] = c;
    }
    dirty_ = false;
    version_ = ++g_dataVersion;
}
const float* TransferFunction::deviceData() {
    if (dirty_) rasterise();
    return static_cast<const float*>(data_.deviceRead());
}
const std::vector<vec4>& TransferFunction::ramData() {
    if (dirty_) rasterise();
    return *data_.getRAMRepresentation();
}

// ------------------------------------------------------------------------ mesh and lights --------
std::shared_ptr<Mesh> Mesh::unitCube() {
    auto m = std::make_shared<Mesh>();
    m->vertices.setSize(8);
    auto* v = m->vertices.getEditableRAMRepresentation();
    int k = 0;
    for (int z = 0; z < 2; ++z)
        for (int y = 0; y < 2; ++y)
            for (int x = 0; x < 2; ++x) (*v)[k++] = vec3((float)x, (float)y, (float)z);
    static const uint32_t idx[36] = {0, 2, 1, 1, 2, 3, 4, 5, 6, 5, 7, 6, 0, 1, 4, 1, 5, 4,
                                     2, 6, 3, 3, 6, 7, 0, 4, 2, 2, 4, 6, 1, 3, 5, 3, 7, 5};
    m->indices.setSize(36);
    auto* i = m->indices.getEditableRAMRepresentation();
    std::copy(idx, idx + 36, i->begin());
    return m;
}
void DirectionalLight::set(vec3 position, vec3 direction) {
    vec3 z = normalize(direction);
    vec3 up = std::fabs(z.y) < 0.99f ? vec3(0, 1, 0) : vec3(1, 0, 0);
    vec3 x = normalize(cross(up, z));
    vec3 y = cross(z, x);
    mat4 m;
    for (int r = 0; r < 3; ++r) {
        m[0][r] = x[r];
        m[1][r] = y[r];
        m[2][r] = z[r];
        m[3][r] = position[r];
    }
    modelToWorld_ = m;
}
void PointLight::setPosition(vec3 p) {
    mat4 m;
    for (int r = 0; r < 3; ++r) m[3][r] = p[r];
    modelToWorld_ = m;
}
PackedLightSource baseLightToPackedLight(const LightSource* light, float radianceScale, const mat4& transformLightMat) {
    PackedLightSource p;
    p.tm = transformLightMat * light->getModelToWorldMatrix();
    vec3 i = light->getIntensity();
    p.radiance = vec4(radianceScale * i.x, radianceScale * i.y, radianceScale * i.z, 1.f);
    p.type = (int)light->getLightSourceType();
    return p;
}

AdvancedMaterialProperty::AdvancedMaterialProperty(std::string id, std::string name)
    : Property(std::move(id), std::move(name))
    , phaseFunctionProp("phaseFunction", "Phase function")
    , indexOfRefractionProp("IOR", "Index of refraction", 1.f, 1.f, 20.f)
    , roughnessProp("roughness", "Roughness", 0.1f, 0.01f, 1.f)
    , anisotropyProp("anisotropy", "Anisotropy (g)", 0.f, -1.f, 1.f)
    , specularColorProp("specularColor", "Specular color", vec4(1.f), vec4(0.f), vec4(1.f)) {
    phaseFunctionProp.addOption("isotropic", "Isotropic", (int)ShadingFunctionKind::Isotropic);
    phaseFunctionProp.addOption("HenyeyGreenstein", "Henyey-Greenstein", (int)ShadingFunctionKind::HenyeyGreenstein);
}
vec4 AdvancedMaterialProperty::getCombinedMaterialParameters() const {
    return vec4(anisotropyProp.get(), roughnessProp.get(), indexOfRefractionProp.get(), 0.f);
}
int AdvancedMaterialProperty::getPhaseFunctionEnum() const {
    return phaseFunctionProp.get() == (int)ShadingFunctionKind::HenyeyGreenstein ? CPM_PHASE_HENYEY_GREENSTEIN : CPM_PHASE_ISOTROPIC;
}

// --------------------------------------------------------------------------- data types ---------
const float PhotonData::defaultRadiusRelativeToSceneRadius = 0.0153866f;
const float PhotonData::defaultSceneRadius = 1.1447142425533318678080422119397f;
const double PhotonData::scaleToMakeLightPowerOfOneVisibleForDirectionalLightSource = 1.0 / 3.14159265358979323846;

void Photon::setDirection(vec3 dir) {
    // ppm/photondata.cpp:100-109 calls unqualified atan2 / acos on float arguments: with GCC's <cmath> those are the C
    // double functions, the result is rounded to float (pinned against the reference's file: tests/test_ref_photondata.py)
    float phi = (float)::atan2((double)dir.y, (double)dir.x);
    float theta = (float)::acos((double)std::min(std::max(dir.z, -1.f), 1.f));
    encodedDirection = vec2{theta, phi};
}
vec3 Photon::getDirection() const {
    float st = std::sin(encodedDirection.x), ct = std::cos(encodedDirection.x);
    float sp = std::sin(encodedDirection.y), cp = std::cos(encodedDirection.y);
    return vec3(st * cp, st * sp, ct);
}
void PhotonData::setSize(size_t numberOfPhotons, int maxPhotonInteractions) {
    maxPhotonInteractions_ = maxPhotonInteractions;
    if (numberOfPhotons > 0) photons_.setSize(numberOfPhotons * 2 * maxPhotonInteractions);
}
void PhotonData::setRadius(double radiusRelativeToSceneSize, double sceneRadius) {
    sceneRadius_ = sceneRadius;
    worldSpaceRadius_ = radiusRelativeToSceneSize * sceneRadius;
}
void PhotonData::advanceToNextIteration(double alpha) {
    setRadius(progressiveSphereRadius(getRadius(), iteration_, alpha));
    iteration_++;
}
double PhotonData::progressiveSphereRadius(double radius, int iteration, double alpha) {
    // Knaus & Zwicker 2011, eq. 20: r_{i+1} = r_i ((i + alpha)/(i + 1))^(1/3)
    return radius * std::pow(((double)iteration + alpha) / (1.0 + (double)iteration), 1.0 / 3.0);
}
double PhotonData::sphereVolume(double radius) { return std::pow(radius, 3) * (3.14159265358979323846 * 4.0 / 3.0); }
double PhotonData::getRelativeIrradianceScale() const {
    double referenceRadiusVolumeScale = sphereVolume(getRadiusRelativeToSceneSize()) / sphereVolume(defaultRadiusRelativeToSceneRadius);
    double nPhotonsScale = (double)getNumberOfPhotons() / (double)defaultNumberOfPhotons;
    return referenceRadiusVolumeScale * nPhotonsScale;
}
void LightSamples::setSize(size_t nSamples) {
    lightSamples_.setSize(nSamples * sizeof(LightSample));
    intersectionPoints_.setSize(nSamples);
}

// ------------------------------------------------------------------------------ launchers -------
void MWC64XSeedGenerator::generateRandomSeeds(Buffer<uvec2>* buffer, unsigned int seed, bool, size_t) {
    auto* ram = buffer->getEditableRAMRepresentation();
    if (ram->empty()) return;
    auto& rt = CpmRuntime::get();
    cpm_rng_host_base_offsets_range(seed, rt.photonShardOffset, reinterpret_cast<uint32_t*>(ram->data()), ram->size());
    uint32_t* dev = static_cast<uint32_t*>(const_cast<void*>(buffer->deviceRead()));
    buffer->deviceWrite();
    ScopedStage st("seed");
    rt.check(cpm_rng_seed_streams(rt.ctx(), dev, ram->size(), 1099511627776ull, rt.photonShardOffset));
}
void MWC64XRandomNumberGenerator::generate(Buffer<float>& out) {
    if (out.getSize() != randomState_.getSize() || dirty_) {
        randomState_.setSize(out.getSize());
        MWC64XSeedGenerator().generateRandomSeeds(&randomState_, seed_, false);
        dirty_ = false;
    }
    auto& rt = CpmRuntime::get();
    uint32_t* st = static_cast<uint32_t*>(const_cast<void*>(randomState_.deviceRead()));
    randomState_.deviceWrite();
    rt.check(cpm_rng_uniform(rt.ctx(), st, out.getSize(), 1, static_cast<float*>(out.deviceWrite())));
}
void UniformSampleGenerator2DCL::generateNextSamples(SampleBuffer& out) {
    auto& rt = CpmRuntime::get();
    float n = (float)(size_t)std::sqrt((double)out.getSize());   // isc/uniformsamplegenerator2dcl.cpp:63
    rt.check(cpm_sample_uniform2d(rt.ctx(), n, n, (int)out.getSize(), static_cast<float*>(out.deviceWrite())));
}

namespace geometry {
static float isPointLeftOfLine(vec2 p0, vec2 p1, vec2 pt) { return (p1.x - p0.x) * (pt.y - p0.y) - (pt.x - p0.x) * (p1.y - p0.y); }

// Andrew-style monotone chain with the reference's treatment of equal-x end points
// (lcl/convexhull2d.cpp:38-130); returns a closed hull (first point repeated when not degenerate).
std::vector<vec2> convexHull2D(std::vector<vec2> pts) {
    std::sort(pts.begin(), pts.end(), [](vec2 a, vec2 b) { return a.x != b.x ? a.x < b.x : a.y < b.y; });
    const int n = (int)pts.size();
    if (n < 4) return pts;
    int loLeft = 0, hiLeft = 1;
    while (hiLeft < n && pts[hiLeft].x == pts[0].x) ++hiLeft;
    --hiLeft;
    std::vector<vec2> hull;
    if (hiLeft == n - 1) {
        hull.push_back(pts[loLeft]);
        if (pts[hiLeft].y != pts[loLeft].y) hull.push_back(pts[hiLeft]);
        hull.push_back(pts[loLeft]);
        return hull;
    }
    int loRight = n - 1, hiRight = n - 2;   // hiRight: first index of the run of maximal x
    while (hiRight >= 0 && !(pts[n - 1].x > pts[hiRight].x)) --hiRight;
    ++hiRight;
    // the reference names these maxXMinY = n-1 and maxXMaxY = first of the max-x run
    const int maxXMinY = loRight, maxXMaxY = hiRight;
    hull.push_back(pts[loLeft]);
    for (int i = hiLeft + 1; i <= maxXMinY; ++i) {
        if (isPointLeftOfLine(pts[loLeft], pts[maxXMinY], pts[i]) >= 0 && i < maxXMinY) continue;
        while (hull.size() >= 2 && !(isPointLeftOfLine(hull[hull.size() - 2], hull.back(), pts[i]) > 0)) hull.pop_back();
        hull.push_back(pts[i]);
    }
    if (maxXMaxY != maxXMinY) hull.push_back(pts[maxXMaxY]);
    const size_t bottom = hull.size() - 1;
    for (int i = maxXMaxY; i > hiLeft; --i) {
        if (isPointLeftOfLine(pts[maxXMaxY], pts[hiLeft], pts[i]) >= 0 && i > hiLeft) continue;
        while (hull.size() - bottom >= 2 && !(isPointLeftOfLine(hull[hull.size() - 2], hull.back(), pts[i]) > 0)) hull.pop_back();
        hull.push_back(pts[i]);
    }
    if (hiLeft != loLeft) hull.push_back(pts[maxXMinY]);
    return hull;
}

PlaneFit fitPlaneAlignedOrientedBoundingBox2D(const std::vector<vec3>& points, vec3 P, vec3 N) {
    auto project = [&](vec3 p) { return p - dot(p - P, N) * N; };
    vec3 u = std::fabs(N.x) > std::fabs(N.y) ? normalize(project(vec3(1, 0, 0)) - P) : normalize(project(vec3(0, 1, 0)) - P);
    vec3 v = normalize(cross(N, u));
    std::vector<vec2> proj;
    const float d = dot(N, P);
    for (const auto& e : points) {
        float dist = dot(N, e) - d;
        vec3 op = (e - dist * N) - P;
        proj.push_back(vec2{dot(u, op), dot(v, op)});
    }
    auto hull = convexHull2D(proj);
    // minimum-area rectangle with one side collinear to a hull edge (rotating callipers, O(h^2))
    float minArea = 3.402823466e+38f;
    vec2 origin, bu, bv;
    const size_t nh = hull.size();
    for (size_t i = 0, j = nh - 1; i < nh; j = i, ++i) {
        float ex = hull[i].x - hull[j].x, ey = hull[i].y - hull[j].y;
        float inv = 1.0f / std::sqrt(ex * ex + ey * ey);   // glm::normalize
        vec2 e0{ex * inv, ey * inv};
        if (e0.x != e0.x || e0.y != e0.y) continue;
        vec2 e1{-e0.y, e0.x};
        float min0 = 0, min1 = 0, max0 = 0, max1 = 0;
        for (size_t k = 0; k < nh; ++k) {
            float dx = hull[k].x - hull[j].x, dy = hull[k].y - hull[j].y;
            float t = dx * e0.x + dy * e0.y;
            min0 = std::min(min0, t);
            max0 = std::max(max0, t);
            t = dx * e1.x + dy * e1.y;
            min1 = std::min(min1, t);
            max1 = std::max(max1, t);
        }
        float area = (max0 - min0) * (max1 - min1);
        if (area < minArea) {
            minArea = area;
            float m0 = std::min(min0, 0.f), m1 = std::min(min1, 0.f);
            origin = vec2{hull[j].x + m0 * e0.x + m1 * e1.x, hull[j].y + m0 * e0.y + m1 * e1.y};
            bu = vec2{e0.x * (max0 - min0), e0.y * (max0 - min0)};
            bv = vec2{e1.x * (max1 - min1), e1.y * (max1 - min1)};
        }
    }
    PlaneFit f;
    f.origin = P + origin.x * u + origin.y * v;
    f.u = bu.x * u + bu.y * v;
    f.v = bv.x * u + bv.y * v;
    return f;
}
}  // namespace geometry

void DirectionalLightSamplerCL::sampleLightSource(const Mesh* mesh, const SampleBuffer* samples, const LightSource* light,
                                                  LightSamples& out) {
    auto* meshNC = const_cast<Mesh*>(mesh);
    const std::vector<vec3>& vertices = *meshNC->vertices.getRAMRepresentation();
    auto* samplesNC = const_cast<SampleBuffer*>(samples);
    if (samples->getSize() != out.getSize()) out.setSize(samples->getSize());
    PackedLightSource lightBase = baseLightToPackedLight(light, 1.f, mesh->getWorldToDataMatrix());
    vec4 d4 = lightBase.tm * vec4(0.f, 0.f, 1.f, 0.f);
    vec3 lightDirection = normalize(vec3(d4.x, d4.y, d4.z));
    vec4 o4 = lightBase.tm * vec4(0.f, 0.f, 0.f, 1.f);
    auto fit = geometry::fitPlaneAlignedOrientedBoundingBox2D(vertices, vec3(o4.x, o4.y, o4.z), lightDirection);
    float area = length(fit.u) * length(fit.v);
    auto& rt = CpmRuntime::get();
    const float rad[3] = {lightBase.radiance.x, lightBase.radiance.y, lightBase.radiance.z};
    lastSetup.direction = lightDirection;
    lastSetup.planePoint = vec3(o4.x, o4.y, o4.z);
    lastSetup.origin = fit.origin;
    lastSetup.u = fit.u;
    lastSetup.v = fit.v;
    lastSetup.radiance = vec3(rad[0], rad[1], rad[2]);
    lastSetup.area = area;
    rt.check(cpm_light_sample_directional(rt.ctx(), static_cast<const float*>(samplesNC->deviceRead()), rad, &lightDirection.x,
                                          &fit.origin.x, &fit.u.x, &fit.v.x, area, (int)samples->getSize(),
                                          static_cast<float*>(out.getLightSamples()->deviceWrite())));
    out.advanceIteration();
}
void LightSampleMeshIntersectionCL::meshSampleIntersection(const Mesh* mesh, LightSamples* samples) {
    auto* m = const_cast<Mesh*>(mesh);
    auto& rt = CpmRuntime::get();
    rt.check(cpm_light_mesh_intersect(rt.ctx(), static_cast<const float*>(m->vertices.deviceRead()),
                                      static_cast<const int32_t*>(m->indices.deviceRead()), (int)m->indices.getSize(),
                                      static_cast<const float*>(samples->getLightSamples()->deviceRead()), (int)samples->getSize(),
                                      static_cast<float*>(samples->getIntersectionPoints()->deviceWrite())));
}

void PhotonTracerCL::setRandomSeedSize(size_t nPhotons) {
    if (nPhotons > 0) {
        randomState_.setSize(nPhotons);
        MWC64XSeedGenerator().generateRandomSeeds(&randomState_, 0, false);   // seed 0: ppm/photontracercl.cpp:180
    }
}
void PhotonTracerCL::prepareOpacityBound(const Volume* volume, TransferFunction& tf) {
    if (!useOpacityBound) return;
    auto& rt = CpmRuntime::get();
    // per-cell opacity bound of (this volume, this transfer function): refreshed when either changed
    size_t nCells = 0;
    const float* range = const_cast<Volume*>(volume)->valueRange(boundCellLog2, &nCells);
    if (opacityBound_.getSize() != nCells || boundVolumeVersion_ != volume->dataVersion() || boundTfVersion_ != tf.version()) {
        ScopedStage st("bound");
        float scale, offset;
        volume->formatScaleOffset(scale, offset);
        opacityBound_.setSize(nCells);
        float* bound = static_cast<float*>(opacityBound_.deviceWrite());
        rt.check(cpm_opacity_bound(rt.ctx(), range, nCells, scale, offset, tf.deviceData(), (int)tf.getTextureSize(), bound));
        if (useBoundTexture) {
            // the tracer reads the grid as a point-sampled 3-D texture (one TEX per collision test)
            const size3_t vd = volume->getDimensions();
            const int vdims[3] = {(int)vd.x, (int)vd.y, (int)vd.z};
            int gd[3];
            cpm_bound_grid_dims(vdims, boundCellLog2, gd);
            if (!boundTex_ || gd[0] != boundTexDims_[0] || gd[1] != boundTexDims_[1] || gd[2] != boundTexDims_[2]) {
                if (boundTex_) cpm_bound_tex_destroy(boundTex_);
                boundTex_ = nullptr;
                rt.check(cpm_bound_tex_create(rt.ctx(), gd, &boundTex_));
                for (int k = 0; k < 3; ++k) boundTexDims_[k] = gd[k];
            }
            rt.check(cpm_bound_tex_update(rt.ctx(), boundTex_, bound));
        }
        boundVolumeVersion_ = volume->dataVersion();
        boundTfVersion_ = tf.version();
    }
}
void PhotonTracerCL::tracePhotons(const Volume* volume, TransferFunction& tf, const vec4 aabb[2],
                                  const AdvancedMaterialProperty& material, float stepSize, const LightSamples* lightSamples,
                                  Buffer<unsigned int>* recomputeIdx, int nInvalidPhotons, int photonOffset, int /*batch*/,
                                  int maxInteractions, PhotonData* photonOutData) {
    if (randomState_.getSize() != photonOutData->getNumberOfPhotons()) setRandomSeedSize(photonOutData->getNumberOfPhotons());
    cpm_trace_params p;
    std::memset(&p, 0, sizeof(p));
    for (int k = 0; k < 3; ++k) {
        p.aabb_min[k] = aabb[0][k];
        p.aabb_max[k] = aabb[1][k];
    }
    vec4 mat = material.getCombinedMaterialParameters();
    for (int k = 0; k < 4; ++k) p.material[k] = mat[k];
    p.phase_function = material.getPhaseFunctionEnum();
    p.step_size = stepSize;
    p.max_interactions = maxInteractions;
    p.photon_offset = photonOffset;
    p.total_photons = (int)photonOutData->getNumberOfPhotons();
    p.n_light_samples = (int)lightSamples->getSize();
    // the reference overwrites instead of appending its -D flags (ppm/photontracercl.cpp:202-207);
    // here the two switches are independent
    p.flags = (progressive_ ? CPM_TRACE_PROGRESSIVE : 0) | (onlyMultipleScattering_ ? CPM_TRACE_NO_SINGLE_SCATTERING : 0);
    if (collisionCounter) p.flags |= CPM_TRACE_STATS;   // the counter buffer holds two counters
    auto& rt = CpmRuntime::get();
    const cpm_volume* vh = const_cast<Volume*>(volume)->handle(volumeLayout);
    uint32_t* rng = static_cast<uint32_t*>(const_cast<void*>(randomState_.deviceRead()));
    if (progressive_) randomState_.deviceWrite();
    const float* ls = static_cast<const float*>(const_cast<LightSamples*>(lightSamples)->getLightSamples()->deviceRead());
    const float* ip = static_cast<const float*>(const_cast<LightSamples*>(lightSamples)->getIntersectionPoints()->deviceRead());
    const uint32_t* idx = recomputeIdx ? static_cast<const uint32_t*>(recomputeIdx->deviceRead()) : nullptr;
    // a re-trace rewrites part of the buffer: the rest must be valid on the device first
    float* photons = recomputeIdx ? static_cast<float*>(const_cast<void*>(photonOutData->photons_.deviceRead())) : nullptr;
    photons = static_cast<float*>(photonOutData->photons_.deviceWrite());
    const float* tfData = tf.deviceData();
    if (useOpacityBound) {
        prepareOpacityBound(volume, tf);
        p.opacity_bound = static_cast<const float*>(opacityBound_.deviceRead());
        p.bound_cell_log2 = boundCellLog2;
        p.opacity_bound_tex = useBoundTexture ? boundTex_ : nullptr;
    }
    rt.check(cpm_trace_photons(rt.ctx(), vh, tfData, (int)tf.getTextureSize(), &p, ls, ip, idx, nInvalidPhotons, photons,
                               rng, collisionCounter));
}

void PhotonRecomputationDetector::photonRecomputationImportance(const PhotonData* photonData, int photonOffset,
                                                                const Volume* origVolume,
                                                                const ImportanceUniformGrid3D* grid,
                                                                const LightSamples& lightSamples,
                                                                Buffer<unsigned int>& importance) {
    auto& rt = CpmRuntime::get();
    size3_t gd = grid->getDimensions(), cd = grid->getCellDimension();
    const int dims[3] = {(int)gd.x, (int)gd.y, (int)gd.z};
    const float cell[3] = {(float)cd.x, (float)cd.y, (float)cd.z};
    mat4 t2i = origVolume->getCoordinateTransformer().getTextureToIndexMatrix();
    auto* pd = const_cast<PhotonData*>(photonData);
    auto& ls = const_cast<LightSamples&>(lightSamples);
    uint32_t* keys = static_cast<uint32_t*>(const_cast<void*>(importance.deviceRead()));
    importance.deviceWrite();
    rt.check(cpm_detect_invalid(rt.ctx(), static_cast<const float*>(grid->data.deviceRead()), dims, cell, t2i.data(),
                                static_cast<const float*>(pd->photons_.deviceRead()), photonOffset,
                                static_cast<const float*>(ls.getLightSamples()->deviceRead()),
                                static_cast<const float*>(ls.getIntersectionPoints()->deviceRead()), (int)lightSamples.getSize(),
                                photonData->getMaxPhotonInteractions(), (int)photonData->getNumberOfPhotons(), keys,
                                equalImportance_ ? 1 : 0, percentage_, iteration_, fixExitPoint_ ? CPM_DETECT_FIX_EXIT : 0));
}

void Radixsort::enqueue(Buffer<unsigned int>& keys, Buffer<unsigned int>* values, size_t elements, unsigned int maxBits) {
    if (elements == 0) throw std::invalid_argument("clogs::Radixsort::enqueue: elements is zero");
    if (keys.getSize() < elements) throw std::invalid_argument("clogs::Radixsort::enqueue: range out of buffer bounds for key");
    if (hasValues_ && (!values || values->getSize() < elements))
        throw std::invalid_argument("clogs::Radixsort::enqueue: range out of buffer bounds for value");
    if (tmpKeys_.getSize() < elements) tmpKeys_.setSize(keys.getSize());
    if (hasValues_ && tmpValues_.getSize() < elements) tmpValues_.setSize(values->getSize());
    auto& rt = CpmRuntime::get();
    uint32_t* k = static_cast<uint32_t*>(const_cast<void*>(keys.deviceRead()));
    keys.deviceWrite();
    uint32_t* v = nullptr;
    if (hasValues_) {
        v = static_cast<uint32_t*>(const_cast<void*>(values->deviceRead()));
        values->deviceWrite();
    }
    rt.check(cpm_radix_sort_u32(rt.ctx(), k, v, elements, maxBits, static_cast<uint32_t*>(tmpKeys_.deviceWrite()),
                                hasValues_ ? static_cast<uint32_t*>(tmpValues_.deviceWrite()) : nullptr));
}

// ============================================================================ processors ========
const ProcessorInfo UniformSampleGenerator2DProcessorCL::processorInfo_{"org.inviwo.UniformSampleGenerator2DCL", "Uniform sample generator 2D",
                                                                        "Sampling", "Experimental", "CL"};
UniformSampleGenerator2DProcessorCL::UniformSampleGenerator2DProcessorCL()
    : samplesPort_("samples")
    , directionalSamplesPort_("DirectionalSamples")
    , nSamples_("nSamples", "N samples", ivec2{256, 256}, ivec2{2, 2}, ivec2{2048, 2048})
    , workGroupSize_("wgsize", "Work group size", ivec2{8, 8}, ivec2{0, 0}, ivec2{256, 256})
    , useGLSharing_("glsharing", "Use OpenGL sharing", true)
    , samples_(std::make_shared<SampleBuffer>())
    , directionalSamples_(std::make_shared<SampleBuffer>()) {
    addPort(samplesPort_);
    addPort(directionalSamplesPort_);
    addProperty(nSamples_);
    addProperty(workGroupSize_);
    addProperty(useGLSharing_);
    samplesPort_.setData(samples_);
    directionalSamplesPort_.setData(directionalSamples_);
}
void UniformSampleGenerator2DProcessorCL::process() {
    size_t n = (size_t)nSamples_.get().x * (size_t)nSamples_.get().y;
    if (n != samples_->getSize()) samples_->setSize(n);
    if (directionalSamples_->getSize() != 0) directionalSamples_->setSize(0);
    sampleGenerator_.generateNextSamples(*samples_);
    samplesPort_.setData(samples_);
}

const ProcessorInfo DirectionalLightSamplerCLProcessor::processorInfo_{"org.inviwo.DirectionalLightSamplerCL", "Directional light sampler",
                                                                       "Light source", "Experimental", "CL"};
DirectionalLightSamplerCLProcessor::DirectionalLightSamplerCLProcessor()
    : boundingVolumeInport_("SceneGeometry")
    , samplesInport_("samples")
    , lightInport_("light")
    , lightSamplesOutport_("LightSamples")
    , workGroupSize_("wgsize", "Work group size", 64, 1, 4096)
    , useGLSharing_("glsharing", "Use OpenGL sharing", true)
    , lightSamples_(std::make_shared<LightSamples>()) {
    addPort(boundingVolumeInport_);
    addPort(samplesInport_);
    addPort(lightInport_);
    addPort(lightSamplesOutport_);
    addProperty(workGroupSize_);
    lightInport_.onChange([this]() { lightSamples_->resetIteration(); });
}
void DirectionalLightSamplerCLProcessor::process() {
    auto samples = samplesInport_.getData();
    auto mesh = boundingVolumeInport_.getData();
    auto light = lightInport_.getData();
    if (!samples || !mesh || !light) return;
    lightSampler_.sampleLightSource(mesh.get(), samples.get(), light.get(), *lightSamples_);
    intersector_.meshSampleIntersection(mesh.get(), lightSamples_.get());
    lightSamplesOutport_.setData(lightSamples_);
}

// ---- ProgressivePhotonTracerCL ---------------------------------------------------------------------
const ProcessorInfo ProgressivePhotonTracerCL::processorInfo_{"org.inviwo.ProgressivePhotonTracerCL", "ProgressivePhotonTracer", "Photons",
                                                              "Experimental", "CL"};
ProgressivePhotonTracerCL::ProgressivePhotonTracerCL()
    : volumePort_("volume")
    , recomputationImportanceGrid_("recomputationImportance")
    , lightSamples_("LightSamples")
    , outport_("photons")
    , recomputedIndicesPort_("recomputedIndices")
    , samplingRate_("samplingRate", "Sampling rate", 1.0f, 1.0f, 15.0f)
    , radius_("radius", "Photon radius (# voxels)", 1.f, 0.00001f, 200.f)
    , sceneRadianceScaling_("radianceScale", "Scene radiance scale", 1.f, 0.01f, 100.f)
    , camera_("camera", "Camera")
    , maxIncrementalPhotonsToUpdate_("maxIncrementalPhotonsToUpdate", "Max photons per update (%)", 100.f, 0.f, 100.f)
    , equalIncrementalImportance_("equalImportance", "Equal importance", false)
    , spatialSorting_("spatialSorting", "Spatial sorting", true)
    , maxScatteringEvents_("maxScatteringEvents", "Max scattering events", 1, 1, 16)
    , noSingleScattering_("noSingleScattering", "No single scattering", false)
    , transferFunction_("transferFunction", "Transfer function")
    , advancedMaterial_("material", "Material")
    , alphaProp_("alpha", "Progressive alpha", 0.5f, 0.0001f, 1.f)
    , workGroupSize_("wgsize", "Work group size", ivec2{8, 8}, ivec2{0, 0}, ivec2{256, 256})
    , useGLSharing_("glsharing", "Use OpenGL sharing", true)
    , invalidateRendering_("invalidate", "Invalidate rendering")
    , enableProgressiveRefinement_("enableRefinement", "Progressive refinement", false)
    , enableProgressivePhotonRecomputation_("enableProgressiveRecomputation", "Progressive recomputation", true)
    , clipX_("clipX", "Clip X Slices", ivec2{0, 256}, ivec2{0, 0}, ivec2{256, 256})
    , clipY_("clipY", "Clip Y Slices", ivec2{0, 256}, ivec2{0, 0}, ivec2{256, 256})
    , clipZ_("clipZ", "Clip Z Slices", ivec2{0, 256}, ivec2{0, 0}, ivec2{256, 256})
    , fixDetectorExitPoint_("fixDetectorExitPoint", "Detector: repaired exit points (not in the reference)", false)
    , photonData_(std::make_shared<PhotonData>())
    , recomputedPhotonIndices_(std::make_shared<RecomputedPhotonIndices>()) {
    using R = PhotonData::InvalidationReason;
    addPort(volumePort_);
    volumePort_.onChange([this]() { invalidateProgressiveRendering(R::Volume); });
    addPort(recomputationImportanceGrid_);
    recomputationImportanceGrid_.setOptional(true);
    recomputationImportanceGrid_.onConnect([this]() {
        invalidateProgressiveRendering(R::All);
        progressiveRefinementChanged();
    });
    addPort(lightSamples_);
    lightSamples_.onChange([this]() {
        for (auto& s : lightSamples_.getVectorData())
            if (s && s->isReset()) invalidateProgressiveRendering(R::Light);
    });
    addPort(outport_);
    addPort(recomputedIndicesPort_);

    addProperty(samplingRate_);
    samplingRate_.onChange([this]() { invalidateProgressiveRendering(R::All); });
    addProperty(radius_);
    radius_.onChange([this]() { invalidateProgressiveRendering(R::All); });
    addProperty(maxScatteringEvents_);
    addProperty(noSingleScattering_);
    noSingleScattering_.onChange([this]() {
        photonTracer_.setNoSingleScattering(noSingleScattering_.get());
        invalidateProgressiveRendering(R::All);
    });
    addProperty(alphaProp_);
    alphaProp_.onChange([this]() { invalidateProgressiveRendering(R::All); });
    addProperty(advancedMaterial_);
    advancedMaterial_.phaseFunctionProp.onChange([this]() { invalidateProgressiveRendering(R::All); invalidate(InvalidationLevel::InvalidOutput); });
    advancedMaterial_.anisotropyProp.onChange([this]() { invalidateProgressiveRendering(R::All); invalidate(InvalidationLevel::InvalidOutput); });
    addProperty(transferFunction_);
    transferFunction_.onChange([this]() { invalidateProgressiveRendering(R::TransferFunction); });
    addProperty(workGroupSize_);
    addProperty(useGLSharing_);
    addProperty(camera_);
    camera_.onChange([this]() {
        invalidateProgressiveRendering(R::Camera);
        photonData_->setIteration(1);
    });
    addProperty(maxIncrementalPhotonsToUpdate_);
    addProperty(equalIncrementalImportance_);
    equalIncrementalImportance_.onChange([this]() { photonRecomputationDetector_.setEqualImportance(equalIncrementalImportance_.get()); });
    addProperty(spatialSorting_);
    addProperty(invalidateRendering_);
    addProperty(enableProgressiveRefinement_);
    addProperty(enableProgressivePhotonRecomputation_);
    addProperty(clipX_);
    addProperty(clipY_);
    addProperty(clipZ_);
    addProperty(fixDetectorExitPoint_);
    fixDetectorExitPoint_.onChange([this]() { photonRecomputationDetector_.setFixExitPoint(fixDetectorExitPoint_.get()); });
    clipX_.setVisible(false);
    clipY_.setVisible(false);
    clipZ_.setVisible(false);
    clipX_.onChange([this]() { onClipChange(); });
    clipY_.onChange([this]() { onClipChange(); });
    clipZ_.onChange([this]() { onClipChange(); });
    enableProgressiveRefinement_.onChange([this]() { progressiveRefinementChanged(); });
    aabb_[0] = vec4(0.f);
    aabb_[1] = vec4(1.f);
    progressiveRefinementChanged();
}

void ProgressivePhotonTracerCL::onTimerEvent() {
    invalidationFlag_ |= PhotonData::InvalidationReason::Progressive;
    invalidateRendering_.pressButton();
}
void ProgressivePhotonTracerCL::progressiveRefinementChanged() {
    // RNG state is saved (new random numbers every iteration) only for pure progressive refinement; with an
    // importance grid connected every re-trace replays the same stream (:642-645) -- the correlation property
    photonTracer_.setProgressive(enableProgressiveRefinement_.get() && !recomputationImportanceGrid_.isConnected());
}
float ProgressivePhotonTracerCL::getSceneRadius() const {
    auto volume = volumePort_.getData();
    float scale = 1.f;
    if (volume) {
        mat4 t2w = volume->getCoordinateTransformer().getTextureToWorldMatrix();
        vec3 ext(length(vec3(t2w[0][0], t2w[0][1], t2w[0][2])), length(vec3(t2w[1][0], t2w[1][1], t2w[1][2])),
                 length(vec3(t2w[2][0], t2w[2][1], t2w[2][2])));
        scale = 0.5f * length(ext);
    }
    return scale;
}
void ProgressivePhotonTracerCL::onClipChange() {
    if (!volumePort_.isReady()) return;
    size3_t d = volumePort_.getData()->getDimensions();
    aabb_[0] = vec4((float)clipX_.get().x / (float)d.x, (float)clipY_.get().x / (float)d.y, (float)clipZ_.get().x / (float)d.z, 1.f);
    aabb_[1] = vec4((float)clipX_.get().y / (float)d.x, (float)clipY_.get().y / (float)d.y, (float)clipZ_.get().y / (float)d.z, 1.f);
    invalidateProgressiveRendering(PhotonData::InvalidationReason::All);
}
void ProgressivePhotonTracerCL::resetPhotonImportance(size_t offset, size_t nPhotons) {
    auto& rt = CpmRuntime::get();
    uint32_t* keys = static_cast<uint32_t*>(photonRecomputationImportance_.deviceWrite());
    rt.check(cpm_mem_fill_u32(rt.ctx(), keys + offset, 2147483647u, nPhotons));
}

void ProgressivePhotonTracerCL::process() {
    using R = PhotonData::InvalidationReason;
    auto lights = lightSamples_.getVectorData();
    auto volumeC = volumePort_.getData();
    if (!volumeC || lights.empty()) return;
    Volume* volume = const_cast<Volume*>(volumeC.get());

    size_t nPhotons = 0;
    for (auto& l : lights) nPhotons += l->getSize();
    if (nPhotons != photonData_->getNumberOfPhotons() || maxScatteringEvents_.get() != photonData_->getMaxPhotonInteractions()) {
        photonData_->setSize(nPhotons, maxScatteringEvents_.get());
        invalidateProgressiveRendering(R::All);
    }
    const float sceneRadius = getSceneRadius();
    mat4 t2i = volume->getCoordinateTransformer().getTextureToIndexMatrix();
    vec3 voxelSpacing(1.f / length(vec3(t2i[0][0], t2i[0][1], t2i[0][2])), 1.f / length(vec3(t2i[1][0], t2i[1][1], t2i[1][2])),
                      1.f / length(vec3(t2i[2][0], t2i[2][1], t2i[2][2])));
    const float stepSize = samplingRate_.get() * std::min(voxelSpacing.x, std::min(voxelSpacing.y, voxelSpacing.z));
    const int maxInteractions = maxScatteringEvents_.get();
    const int flag = static_cast<int>(invalidationFlag_);
    if (flag == 0 || (flag & (static_cast<int>(R::Light) | static_cast<int>(R::Camera) | static_cast<int>(R::TransferFunction) |
                              static_cast<int>(R::Volume)))) {
        photonData_->resetIteration();
    }
    if (photonData_->iteration() == 0) {
        mat4 i2t = volume->getCoordinateTransformer().getIndexToTextureMatrix();
        vec4 r = i2t * vec4(vec3(radius_.get()), 0.f);
        photonData_->setRadius(length(vec3(r.x, r.y, r.z)), sceneRadius);
        photonData_->setIteration(1);
    } else {
        photonData_->advanceToNextIteration(alphaProp_.get());
    }

    size_t nPhotonsToCompute = photonData_->getNumberOfPhotons();
    const bool correlated = !(flag & static_cast<int>(R::Light)) && recomputationImportanceGrid_.isReady();
    if (correlated) {
        const size_t N = photonData_->getNumberOfPhotons();
        if (photonRecomputationImportance_.getSize() != N) {
            photonRecomputationImportance_.setSize(N);
            sortedImportance_.setSize(N);
            resetPhotonImportance(0, N);
        }
        if (recomputedPhotonIndices_->indicesToRecomputedPhotons.getSize() != N)
            recomputedPhotonIndices_->indicesToRecomputedPhotons.setSize(N);
        auto& indices = recomputedPhotonIndices_->indicesToRecomputedPhotons;
        auto& rt = CpmRuntime::get();
        const bool globalBudget = rt.globalBudget && rt.comm && cpm_comm_world(rt.comm) > 1;

        if (flag & (static_cast<int>(R::TransferFunction) | static_cast<int>(R::Volume))) {
            auto grid = dynamic_cast<const ImportanceUniformGrid3D*>(recomputationImportanceGrid_.getData().get());
            if (!grid) {
                LogError("UniformGrid3DInport require ImportanceUniformGrid3D as input");
                return;
            }
            photonRecomputationDetector_.setPercentage(static_cast<int>(maxIncrementalPhotonsToUpdate_.get()));
            photonRecomputationDetector_.setIteration(photonRecomputationDetector_.getIteration() + 1);
            // 1. importance of every stored path (one launch per light, as the reference)
            StageProfiler::get().begin("detector");
            int offset = 0;
            for (auto& l : lights) {
                photonRecomputationDetector_.photonRecomputationImportance(photonData_.get(), offset, volume, grid, *l,
                                                                           photonRecomputationImportance_);
                offset += (int)l->getSize();
            }
            StageProfiler::get().end();
            // 2. exact, synchronous count of the invalid photons (the reference reads its count early) together
            //    with their ids in ascending order
            StageProfiler::get().begin("select");
            long long nInvalid = 0;
            const uint32_t* keys = static_cast<const uint32_t*>(photonRecomputationImportance_.deviceRead());
            rt.check(cpm_select_below_begin(rt.ctx(), keys, N, 2147483647u, static_cast<uint32_t*>(indices.deviceWrite())));
            StageProfiler::get().end();
            // the count is on its way to the host: the device meanwhile refreshes the tracer's opacity bound, which
            // the re-trace needs whatever the count is
            photonTracer_.prepareOpacityBound(volume, transferFunction_.get());
            rt.check(cpm_select_below_end(rt.ctx(), &nInvalid));
            const long long budget = static_cast<long long>((maxIncrementalPhotonsToUpdate_.get() / 100.f) * (float)N);
            selectionIsSorted_ = spatialSorting_.get() && nInvalid <= budget;
            if (globalBudget) {
                // SURVEY 8e "select globally": the budget max% * N_total is applied to the photons of ALL shards in one
                // importance order.  Every decision below derives from all-gathered values, so the ranks stay in lockstep.
                const int W = cpm_comm_world(rt.comm);
                unsigned long long mine[2] = {(unsigned long long)N, (unsigned long long)nInvalid};
                std::vector<unsigned long long> all((size_t)2 * W);
                rt.check(cpm_comm_allgather_u64(rt.comm, mine, 2, all.data()));
                unsigned long long nTotal = 0, invalidTotal = 0;
                for (int r = 0; r < W; ++r) {
                    nTotal += all[2 * r];
                    invalidTotal += all[2 * r + 1];
                }
                globalBudget_ = static_cast<long long>((maxIncrementalPhotonsToUpdate_.get() / 100.f) * (float)nTotal);
                selectionIsSorted_ = spatialSorting_.get() && (long long)invalidTotal <= globalBudget_;
                remainingGlobal_ = (long long)invalidTotal;
                globalOffset_ = 0;
            }
            if (!selectionIsSorted_) {
                // 3. the budget cuts the list (or the importance order is wanted): sort all photon ids by importance key.
                //    The keys are sorted on a COPY so that key[i] keeps belonging to photon i (the reference permutes
                //    them in place)
                StageProfiler::get().begin("sort");
                rt.check(cpm_iota_u32(rt.ctx(), static_cast<uint32_t*>(indices.deviceWrite()), N));
                rt.check(cpm_mem_copy_d2d(rt.ctx(), sortedImportance_.deviceWrite(), keys, N * sizeof(uint32_t)));
                recomputationImportanceSorter_.enqueue(sortedImportance_, &indices, N, 0);
                StageProfiler::get().end();
            }
            // else: every invalid photon fits the budget and the ids are wanted in ascending order -- sorting by
            // importance, cutting at nInvalid and re-sorting by id (:361-363, 467-473) yields exactly the list
            // cpm_select_below just wrote
            remainingPhotonsOffset_ = 0;
            if (remainingPhotonsToUpdate_ < 0 || nInvalid > 0) remainingPhotonsToUpdate_ = (int)nInvalid;
        }
        int maxPhotonsToUpdate = static_cast<int>((maxIncrementalPhotonsToUpdate_.get() / 100.f) * (float)N);
        nPhotonsToCompute = (size_t)std::max(0, std::min(remainingPhotonsToUpdate_, maxPhotonsToUpdate));
        long long takeGlobal = 0;
        if (globalBudget) {
            takeGlobal = std::max(0LL, std::min(remainingGlobal_, globalBudget_));
            if (selectionIsSorted_) {
                nPhotonsToCompute = (size_t)std::max(0, remainingPhotonsToUpdate_);   // nothing is cut: every invalid photon of this shard
            } else {
                // this shard's part of the next `takeGlobal` photons of the global importance order: the sorted id list is
                // consumed front to back, remainingPhotonsOffset_ entries are already done
                unsigned long long upTo = 0;
                rt.check(cpm_comm_select_global(rt.comm, static_cast<const uint32_t*>(sortedImportance_.deviceRead()), N,
                                                (unsigned long long)(globalOffset_ + takeGlobal), &upTo));
                nPhotonsToCompute = (size_t)std::max(0LL, (long long)upTo - (long long)remainingPhotonsOffset_);
            }
        }
        if (remainingPhotonsOffset_ > 0 && nPhotonsToCompute > 0) {
            // continue a budgeted batch: slide the next slice of the sorted id list to the front
            uint32_t* idx = static_cast<uint32_t*>(const_cast<void*>(indices.deviceRead()));
            indices.deviceWrite();
            rt.check(cpm_mem_copy_d2d(rt.ctx(), idx, idx + remainingPhotonsOffset_, nPhotonsToCompute * sizeof(uint32_t)));
        }
        recomputedPhotonIndices_->nRecomputedPhotons = static_cast<int>(nPhotonsToCompute);
        if (nPhotonsToCompute > 0) {
            if (spatialSorting_.get() && !selectionIsSorted_) {
                // keys-only sort of the selected ids: ascending id == raster order on the light plane (:467-473)
                StageProfiler::get().begin("indexsort");
                recomputationIndexSorter_.enqueue(indices, nullptr, nPhotonsToCompute, 0);
                StageProfiler::get().end();
            }
            // the tracer's volume layout and opacity bound are refreshed before the stage clock starts ("texcopy" and
            // "bound" are stages of their own; nested inside "trace" they were counted twice)
            volume->handle(photonTracer_.volumeLayout);
            photonTracer_.prepareOpacityBound(volume, transferFunction_.get());
            StageProfiler::get().begin("trace");
            int offset = 0;
            for (auto& l : lights) {
                photonTracer_.tracePhotons(volume, transferFunction_.get(), aabb_, advancedMaterial_, stepSize, l.get(), &indices,
                                           (int)nPhotonsToCompute, offset, 0, maxInteractions, photonData_.get());
                offset += (int)l->getSize();
            }
            StageProfiler::get().end();
            // reset the keys of the photons just re-traced.  Keys stay in photon order here, so the reset goes
            // through the id list (the reference resets the matching slice of its in-place sorted keys, :529)
            uint32_t* keysW = static_cast<uint32_t*>(const_cast<void*>(photonRecomputationImportance_.deviceRead()));
            photonRecomputationImportance_.deviceWrite();
            rt.check(cpm_mem_scatter_fill_u32(rt.ctx(), keysW, static_cast<const uint32_t*>(indices.deviceRead()), nPhotonsToCompute,
                                              2147483647u));
        }
        remainingPhotonsOffset_ += (int)nPhotonsToCompute;
        remainingPhotonsToUpdate_ -= (int)nPhotonsToCompute;
        if (globalBudget) {
            globalOffset_ += takeGlobal;
            remainingGlobal_ -= takeGlobal;
            // batches continue while ANY shard has photons left (the selection above is a collective)
            remainingPhotonsToUpdate_ = (int)std::min<long long>(remainingGlobal_, 2147483647LL);
        }
        if (remainingPhotonsToUpdate_ > 0 && enableProgressivePhotonRecomputation_.get()) {
            enableProgressiveRefinement_.set(true);
        } else {
            enableProgressiveRefinement_.set(false);
        }
    } else {
        volume->handle(photonTracer_.volumeLayout);
        photonTracer_.prepareOpacityBound(volume, transferFunction_.get());
        StageProfiler::get().begin("trace");
        int offset = 0;
        for (auto& l : lights) {
            photonTracer_.tracePhotons(volume, transferFunction_.get(), aabb_, advancedMaterial_, stepSize, l.get(), nullptr, 0, offset, 0,
                                       maxInteractions, photonData_.get());
            offset += (int)l->getSize();
        }
        StageProfiler::get().end();
        recomputedPhotonIndices_->nRecomputedPhotons = -1;
        remainingPhotonsToUpdate_ = 0;
        remainingPhotonsOffset_ = 0;
        if (photonRecomputationImportance_.getSize() > 0) resetPhotonImportance(0, photonRecomputationImportance_.getSize());
    }
    recomputedIndicesPort_.setData(recomputedPhotonIndices_);
    photonData_->setInvalidationReason(invalidationFlag_);
    invalidationFlag_ = R(0);
    outport_.setData(photonData_);
    if (enableProgressiveRefinement_.get() && remainingPhotonsToUpdate_ > 0) invalidate(InvalidationLevel::InvalidOutput);
}

// ---- .u3d reader / writer ---------------------------------------------------------------------------
namespace {
std::string trimmed(const std::string& s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
std::string lowered(std::string s) {
    for (auto& c : s) c = (char)std::tolower((unsigned char)c);
    return s;
}
std::string parentDir(const std::string& path) {
    size_t k = path.find_last_of('/');
    return k == std::string::npos ? std::string() : path.substr(0, k + 1);
}
void readMatrix(std::istream& ss, mat4& m) {
    // written row by row (the writer transposes glm's column-major matrix first), read back and transposed again
    mat4 t;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) ss >> t[i][j];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) m[i][j] = t[j][i];
}
}  // namespace

std::shared_ptr<UniformGrid3DVector> UniformGrid3DReader::readData(const std::string& filePath) {
    std::ifstream f(filePath);
    if (!f.good()) throw std::invalid_argument("Error: Unable to open file: " + filePath);
    std::string rawFile, formatFlag;
    mat4 modelMatrix, worldMatrix;
    size3_t cellDimensions(0);
    size_t resolution[4] = {0, 0, 0, 0};
    // the header is line oriented except for the matrices, whose four rows follow the key on four lines
    std::string textLine;
    while (std::getline(f, textLine)) {
        textLine = trimmed(textLine);
        if (textLine.empty() || textLine[0] == '#' || textLine[0] == '/') continue;
        std::string noComment = textLine.substr(0, textLine.find('#'));
        size_t colon = noComment.find(':');
        if (colon == std::string::npos || noComment.find(':', colon + 1) != std::string::npos) continue;
        std::string key = lowered(trimmed(noComment.substr(0, colon))), value = trimmed(noComment.substr(colon + 1));
        if (key == "modelmatrix" || key == "worldmatrix") {
            // the reference parses only the key's own line (one row) and leaves the rest of the matrix at identity
            // rows (uniformgrid3dreader.cpp:100-113); its writer, however, emits four lines.  Read all four so that
            // a written file round-trips.
            for (int r = 1; r < 4; ++r) {
                std::streampos pos = f.tellg();
                std::string more;
                if (!std::getline(f, more)) break;
                if (more.find(':') != std::string::npos) {   // next key already: a one-line matrix
                    f.seekg(pos);
                    break;
                }
                value += " " + trimmed(more);
            }
        }
        std::stringstream ss(value);
        if (key == "objectfilename" || key == "rawfile") {
            rawFile = parentDir(filePath) + value;
        } else if (key == "resolution" || key == "dimensions") {
            ss >> resolution[0] >> resolution[1] >> resolution[2] >> resolution[3];
        } else if (key == "format") {
            ss >> formatFlag;
        } else if (key == "modelmatrix") {
            readMatrix(ss, modelMatrix);
        } else if (key == "worldmatrix") {
            readMatrix(ss, worldMatrix);
        } else if (key == "celldimensions") {
            ss >> cellDimensions.x >> cellDimensions.y >> cellDimensions.z;
        }
    }
    if (resolution[0] == 0 && resolution[1] == 0 && resolution[2] == 0 && resolution[3] == 0)
        throw std::invalid_argument("Error: Unable to find \"Resolution\" tag in file: " + filePath);
    if (formatFlag.empty()) throw std::invalid_argument("Error: Unable to find \"Format\" tag in file: " + filePath);
    std::shared_ptr<UniformGrid3DBase> data;
    const size3_t dim(resolution[0], resolution[1], resolution[2]);
    if (formatFlag == "FLOAT32")
        data = std::make_shared<UniformGrid3D<float>>(dim, cellDimensions);
    else if (formatFlag == "Vec2UINT16")
        data = std::make_shared<UniformGrid3D<u16vec2>>(dim, cellDimensions);
    else
        throw std::invalid_argument("Error: Unsupported data fromat \"Format\" tag in file: " + filePath + " (" + formatFlag +
                                    "; this build reads FLOAT32 and Vec2UINT16 grids)");
    data->setModelMatrix(modelMatrix);
    data->setWorldMatrix(worldMatrix);
    auto dataVector = std::make_shared<UniformGrid3DVector>();
    std::ifstream fin(rawFile, std::ios::in | std::ios::binary);
    if (!fin.good()) throw std::invalid_argument("Error: Unable to read from  file: " + rawFile);
    const size_t bytes = data->getSizeInBytes();
    for (size_t t = 0; t < resolution[3]; ++t) {
        dataVector->push_back(t == 0 ? data : data->cloneEmpty());
        fin.read(static_cast<char*>(dataVector->back()->getData()), (std::streamsize)bytes);
        if ((size_t)fin.gcount() != bytes) throw std::invalid_argument("Error: raw file too short: " + rawFile);
    }
    return dataVector;
}

void UniformGrid3DWriter::writeData(const UniformGrid3DVector* vectorData, const std::string& filePath) const {
    if (!vectorData || vectorData->size() < 1) throw std::invalid_argument("Error: Cannot write empty vector");
    std::string rawPath = filePath;
    size_t dot = rawPath.find_last_of('.');
    if (dot != std::string::npos && rawPath.find('/', dot) == std::string::npos) rawPath.erase(dot);
    rawPath += ".raw";
    if (!overwrite_) {
        if (std::ifstream(filePath).good() || std::ifstream(rawPath).good())
            throw std::invalid_argument("Error: file exists and overwrite is off: " + filePath);
    }
    size_t slash = rawPath.find_last_of('/');
    const std::string rawName = slash == std::string::npos ? rawPath : rawPath.substr(slash + 1);
    UniformGrid3DBase* data = vectorData->front().get();
    std::stringstream ss;
    const size3_t dim = data->getDimensions(), cell = data->getCellDimension();
    ss << "RawFile: " << rawName << std::endl;
    ss << "Resolution: " << dim.x << " " << dim.y << " " << dim.z << " " << vectorData->size() << std::endl;
    ss << "Format: " << data->getFormatString() << std::endl;
    const mat4 mats[2] = {data->getModelMatrix(), data->getWorldMatrix()};
    const char* names[2] = {"ModelMatrix", "WorldMatrix"};
    for (int k = 0; k < 2; ++k) {
        ss << names[k] << ":";
        for (int i = 0; i < 4; ++i) {   // row i of the matrix = element [j][i] of the column-major storage
            for (int j = 0; j < 4; ++j) ss << " " << mats[k][j][i];
            ss << std::endl;
        }
    }
    ss << "CellDimensions: " << cell.x << " " << cell.y << " " << cell.z << std::endl;
    std::ofstream f(filePath);
    if (!f.good()) throw std::invalid_argument("Could not write to file: " + filePath);
    f << ss.str();
    f.close();
    std::ofstream fout(rawPath, std::ios::out | std::ios::binary);
    if (!fout.good()) throw std::invalid_argument("Could not write to raw file: " + rawPath);
    for (auto& element : *vectorData)
        fout.write(static_cast<const char*>(element->getData()), (std::streamsize)element->getSizeInBytes());
}

// ---- PhotonToLightVolumeProcessorCL -----------------------------------------------------------------
const ProcessorInfo PhotonToLightVolumeProcessorCL::processorInfo_{"org.inviwo.PhotonToLightVolumeProcessorCL", "Photon to light volume",
                                                                   "Photons", "Experimental", "CL"};
PhotonToLightVolumeProcessorCL::PhotonToLightVolumeProcessorCL()
    : volumeInport_("volume")
    , photons_("photons")
    , recomputedPhotonIndicesPort_("recomputedPhotonIndices")
    , outport_("lightvolume")
    , incrementalRecomputationThreshold_("incrementalRecomputationThreshold", "Max % invalid photons to use add-remove", 50.f, 0.f, 100.f)
    , volumeSizeOption_("volumeSizeOption", "Light Volume Size")
    , volumeDataTypeOption_("volumeDataType", "Output data type")
    , alignChangedPhotons_("alignChangedPhotons", "Mem-align changed photons", false)
    , workGroupSize_("wgsize", "Work group size", 128, 1, 2048)
    , useGLSharing_("glsharing", "Use OpenGL sharing", true)
    , lightVolume_(std::make_shared<Volume>(size3_t(1), DataFormatBase::get(DataFormatId::Float32))) {
    addPort(volumeInport_);
    addPort(photons_);
    recomputedPhotonIndicesPort_.setOptional(true);
    addPort(recomputedPhotonIndicesPort_);
    addPort(outport_);
    volumeInport_.onChange([this]() { volumeSizeOptionChanged(); });
    addProperty(incrementalRecomputationThreshold_);
    volumeSizeOption_.addOption("radius", "Photon radius", 0);
    volumeSizeOption_.addOption("1", "Full of incoming volume", 1);
    volumeSizeOption_.addOption("1/2", "Half of incoming volume", 2);
    volumeSizeOption_.addOption("1/4", "Quarter of incoming volume", 4);
    volumeSizeOption_.onChange([this]() { volumeSizeOptionChanged(); });
    volumeDataTypeOption_.addOption("float32", "float32", 1);
    volumeDataTypeOption_.addOption("4xfloat32", "4 x float32", 4);
    volumeDataTypeOption_.onChange([this]() {
        lightVolume_ = std::make_shared<Volume>(lightVolume_->getDimensions(),
                                                DataFormatBase::get(volumeDataTypeOption_.get() == 1 ? DataFormatId::Float32 : DataFormatId::Vec4Float32));
        prevPhotons_.setSize(0);
    });
    addProperty(volumeSizeOption_);
    addProperty(volumeDataTypeOption_);
    addProperty(alignChangedPhotons_);
    addProperty(workGroupSize_);
    addProperty(useGLSharing_);
    outport_.setData(lightVolume_);
}
void PhotonToLightVolumeProcessorCL::volumeSizeOptionChanged() {
    if (volumeInport_.hasData() && volumeSizeOption_.get() != 0) {
        auto in = volumeInport_.getData();
        size3_t d = in->getDimensions();
        size_t div = (size_t)volumeSizeOption_.get();
        size3_t newSize(d.x / div, d.y / div, d.z / div);
        if (newSize != lightVolume_->getDimensions()) {
            lightVolume_->setDimensions(newSize);
            lightVolume_->setModelMatrix(in->getModelMatrix());
            lightVolume_->setWorldMatrix(in->getWorldMatrix());
            prevPhotons_.setSize(0);
        }
    }
}
void PhotonToLightVolumeProcessorCL::process() {
    auto photonDataC = photons_.getData();
    auto volume = volumeInport_.getData();
    if (!photonDataC || !volume) return;
    PhotonData* photonData = const_cast<PhotonData*>(photonDataC.get());
    if (volumeSizeOption_.get() == 0) {
        double invPhotonRadius = 1.0 / photonData->getRadiusRelativeToSceneSize();
        size3_t dims((size_t)std::ceil(invPhotonRadius));
        if (dims != lightVolume_->getDimensions()) {
            lightVolume_->setDimensions(dims);
            lightVolume_->setModelMatrix(volume->getModelMatrix());
            lightVolume_->setWorldMatrix(volume->getWorldMatrix());
            prevPhotons_.setSize(0);
        }
    }
    const size3_t od = lightVolume_->getDimensions();
    const int outDim[3] = {(int)od.x, (int)od.y, (int)od.z};
    const int channels = (int)lightVolume_->getDataFormat()->components;
    const int N = (int)photonData->getNumberOfPhotons(), I = photonData->getMaxPhotonInteractions();
    const int maxRecomputationPhotons = static_cast<int>((float)N * (incrementalRecomputationThreshold_.get() / incrementalRecomputationThreshold_.getMaxValue()));
    mat4 t2i = lightVolume_->getCoordinateTransformer().getTextureToIndexMatrix();
    mat4 i2t = lightVolume_->getCoordinateTransformer().getIndexToTextureMatrix();
    const float radius = (float)photonData->getRadiusRelativeToSceneSize();
    const double photonVolume = PhotonData::sphereVolume(photonData->getRadiusRelativeToSceneSize());
    const float scale = (float)(PhotonData::scaleToMakeLightPowerOfOneVisibleForDirectionalLightSource / (photonVolume * (double)N));
    auto& rt = CpmRuntime::get();
    auto idxData = recomputedPhotonIndicesPort_.isReady() ? recomputedPhotonIndicesPort_.getData() : nullptr;
    const int nRecomputed = idxData ? idxData->nRecomputedPhotons : -1;
    const float* photonsDev = static_cast<const float*>(photonData->photons_.deviceRead());
    lastPath = "none";
    bool prevInSync = false;
    auto beforeWrite = [&]() {
        if (waitBeforeLightVolumeWrite) {
            rt.check(cpm_ctx_wait_cuda_event(rt.ctx(), waitBeforeLightVolumeWrite));
            waitBeforeLightVolumeWrite = nullptr;
        }
    };
    if (idxData && prevPhotons_.getSize() == photonData->photons_.getSize() && nRecomputed > 0 && nRecomputed < maxRecomputationPhotons) {
        // incremental: remove the old contribution of the re-traced photons, add the new one (:262-274)
        auto* idxBuf = const_cast<Buffer<unsigned int>*>(&idxData->indicesToRecomputedPhotons);
        const uint32_t* idx = static_cast<const uint32_t*>(idxBuf->deviceRead());
        float* lv = static_cast<float*>(const_cast<void*>(lightVolume_->deviceRead()));
        lightVolume_->deviceWrite();
        ScopedStage st("splat");
        beforeWrite();
        if (alignChangedPhotons_.get()) {
            // :207-244 -- gather the old (power * -1) and the new (power * +1) records of the listed ids into one packed
            // buffer (copyIndexPhotonsKernel twice) and splat its 2 * n * I records as plain photons.  The reference sizes
            // the buffer for one interaction (maxRecomputationPhotons * 4 vec4); here it holds every interaction.
            const size_t need = (size_t)nRecomputed * (size_t)I * 4;
            if (changedAlignedPhotons_.getSize() < need) changedAlignedPhotons_.setSize(std::max(need, (size_t)maxRecomputationPhotons * 4));
            float* packed = reinterpret_cast<float*>(changedAlignedPhotons_.deviceWrite());
            rt.check(cpm_copy_index_photons(rt.ctx(), static_cast<const float*>(prevPhotons_.deviceRead()), idx, nRecomputed, -1.f, N, I, packed, 0));
            rt.check(cpm_copy_index_photons(rt.ctx(), photonsDev, idx, nRecomputed, 1.f, N, I, packed, (size_t)nRecomputed * (size_t)I));
            rt.check(cpm_splat_photons(rt.ctx(), lv, channels, t2i.data(), i2t.data(), outDim, packed, nullptr, 2 * nRecomputed * I, N, I,
                                       radius, scale, 1.f));
            lastPath = "incremental-aligned";
        } else {
        // (the kernel leaves prevPhotons_ equal to the new records of the listed ids: no whole-buffer copy below)
        rt.check(cpm_splat_photons_update_sync(rt.ctx(), lv, channels, t2i.data(), i2t.data(), outDim,
                                               static_cast<float*>(prevPhotons_.deviceWrite()), photonsDev, idx, nRecomputed, N, I,
                                               radius, scale));
        lastPath = "incremental";
        prevInSync = true;
        }
    } else if (prevPhotons_.getSize() != photonData->photons_.getSize() || nRecomputed < 0 || nRecomputed >= maxRecomputationPhotons) {
        float* lv = static_cast<float*>(lightVolume_->deviceWrite());
        ScopedStage st("splat");
        beforeWrite();
        rt.check(cpm_mem_fill_u32(rt.ctx(), lv, 0u, od.x * od.y * od.z * (size_t)channels));
        const int n = referenceFullSplatBound ? N : N * I;
        rt.check(cpm_splat_photons(rt.ctx(), lv, channels, t2i.data(), i2t.data(), outDim, photonsDev, nullptr, n, N, I, radius, scale, 1.f));
        lastPath = "full";
    }
    if (idxData && nRecomputed != 0 && !prevInSync) {
        // keep a copy of the photons so that the next incremental update can subtract them (:488-497)
        if (prevPhotons_.getSize() != photonData->photons_.getSize()) prevPhotons_.setSize(photonData->photons_.getSize());
        ScopedStage st("copyprev");
        rt.check(cpm_mem_copy_d2d(rt.ctx(), prevPhotons_.deviceWrite(), photonsDev, photonData->photons_.getSizeInBytes()));
    }
    outport_.setData(lightVolume_);
}

// ---- VolumeMinMaxCLProcessor ------------------------------------------------------------------------------
const ProcessorInfo VolumeMinMaxCLProcessor::processorInfo_{"org.inviwo.VolumeMinMaxCLProcessor", "Volume min max", "Volume Operation",
                                                            "Experimental", "CL"};
VolumeMinMaxCLProcessor::VolumeMinMaxCLProcessor()
    : inport_("volume")
    , vectorInport_("VolumeSequenceInput")
    , outport_("output")
    , vectorOutport_("UniformGrid3DVectorOut")
    , volumeRegionSize_("region", "Region size", 8, 1, 100)
    , workGroupSize_("wgsize", "Work group size", ivec3{4, 4, 4}, ivec3{0, 0, 0}, ivec3{256, 256, 256})
    , useGLSharing_("glsharing", "Use OpenGL sharing", true) {
    inport_.setOptional(true);
    vectorInport_.setOptional(true);
    addPort(inport_);
    addPort(vectorInport_);
    addPort(outport_);
    addPort(vectorOutport_);
    addProperty(volumeRegionSize_);
    addProperty(workGroupSize_);
    addProperty(useGLSharing_);
}
std::unique_ptr<MinMaxUniformGrid3D> VolumeMinMaxCLProcessor::compute(const Volume* volume) {
    const size3_t dim = volume->getDimensions();
    const size_t r = (size_t)volumeRegionSize_.get();
    const size3_t outDim((dim.x + r - 1) / r, (dim.y + r - 1) / r, (dim.z + r - 1) / r);
    std::unique_ptr<MinMaxUniformGrid3D> out(new MinMaxUniformGrid3D(size3_t(r)));
    out->setModelMatrix(volume->getModelMatrix());
    out->setWorldMatrix(volume->getWorldMatrix());
    out->setDimensions(outDim);
    auto& rt = CpmRuntime::get();
    const cpm_volume* vh = const_cast<Volume*>(volume)->handle(CPM_VOLUME_LINEAR);
    ScopedStage st("minmax");
    rt.check(cpm_volume_minmax(rt.ctx(), vh, (int)r,
                               static_cast<uint16_t*>(out->data.deviceWrite()), nullptr));
    return out;
}
void VolumeMinMaxCLProcessor::process() {
    if (vectorInport_.isReady()) {
        outport_.setData(std::shared_ptr<const UniformGrid3DBase>());
        auto volumes = vectorInport_.getData();
        auto output = std::make_shared<UniformGrid3DVector>();
        for (auto& v : *volumes) output->emplace_back(std::shared_ptr<UniformGrid3DBase>(compute(v.get()).release()));
        vectorOutport_.setData(output);
    }
    if (inport_.isReady()) outport_.setData(std::shared_ptr<const UniformGrid3DBase>(compute(inport_.getData().get()).release()));
}

// ---- DynamicVolumeDifferenceAnalysis -----------------------------------------------------------------------
const ProcessorInfo DynamicVolumeDifferenceAnalysis::processorInfo_{"org.inviwo.DynamicVolumeDifferenceAnalysis",
                                                                    "Dynamic Volume Difference Analysis", "Volume", "Experimental", "CPU"};
DynamicVolumeDifferenceAnalysis::DynamicVolumeDifferenceAnalysis()
    : inport_("data"), outport_("DynamicDataInfo"), volumeRegionSize_("region", "Region size", 8, 1, 100) {
    addPort(inport_);
    addPort(outport_);
    addProperty(volumeRegionSize_);
}
std::shared_ptr<DynamicVolumeInfoUniformGrid3D> DynamicVolumeDifferenceAnalysis::difference(Volume* cur, Volume* nxt, size_t r) {
    auto& rt = CpmRuntime::get();
    const size3_t dim = cur->getDimensions();
    const size3_t outDim((dim.x + r - 1) / r, (dim.y + r - 1) / r, (dim.z + r - 1) / r);
    auto out = std::make_shared<DynamicVolumeInfoUniformGrid3D>(size3_t(r));
    out->setModelMatrix(cur->getModelMatrix());
    out->setWorldMatrix(cur->getWorldMatrix());
    out->setDimensions(outDim);
    dvec2 dataRange = cur->dataMap_.dataRange;
    double typeRange = cur->getDataFormat()->maxValue;   // DataMapper(format).dataRange = (0, max)
    double defaultToDataRange = typeRange / (dataRange.y - dataRange.x);
    const cpm_volume* a = cur->handle(CPM_VOLUME_LINEAR);
    const cpm_volume* b = nxt->handle(CPM_VOLUME_LINEAR);
    ScopedStage st("voldiff");
    rt.check(cpm_volume_diff_bricks(rt.ctx(), a, b, (int)r, defaultToDataRange, dataRange.x, dataRange.y,
                                    static_cast<float*>(out->data.deviceWrite())));
    return out;
}
void DynamicVolumeDifferenceAnalysis::process() {
    auto data = inport_.getData();
    if (!data) return;
    auto output = std::make_shared<UniformGrid3DVector>();
    const size_t r = (size_t)volumeRegionSize_.get();
    for (size_t t = 0; t < data->size(); ++t) {
        size_t next = (t + 1) % data->size();
        output->emplace_back(difference((*data)[t].get(), (*data)[next].get(), r));
    }
    outport_.setData(output);
}

// ---- MinMaxUniformGrid3DImportanceCLProcessor ----------------------------------------------------------------
const ProcessorInfo MinMaxUniformGrid3DImportanceCLProcessor::processorInfo_{"org.inviwo.MinMaxUniformGrid3DImportanceCLProcessor",
                                                                             "MinMaxUniformGrid3D Importance", "UniformGrid3D",
                                                                             "Experimental", "CL"};
MinMaxUniformGrid3DImportanceCLProcessor::MinMaxUniformGrid3DImportanceCLProcessor()
    : minMaxUniformGrid3DInport_("minMaxUniformGrid3D")
    , volumeDifferenceInfoInport_("volumeDifferenceInfo")
    , importanceUniformGrid3DOutport_("importanceUniformGrid3D")
    , incrementalImportance("incrementalImportance", "Incremental importance", true)
    , opacityWeight_("constantWeight", "Opacity weight", 1.f, 0.f, 1.f)
    , opacityDiffWeight_("opacityDiffWeight", "Opacity difference weight", 0.f, 0.f, 1.f)
    , colorWeight_("colorWeight", "Color weight", 0.f, 0.f, 1.f)
    , colorDiffWeight_("colorDiffWeight", "Color difference weight", 0.f, 0.f, 1.f)
    , useAssociatedColor_("useAssociatedColor", "Associated color", false)
    , TFPointEpsilon_("TFPointEpsilon", "Minimum change threshold", 1e-4f, 0.f, 1e-2f)
    , transferFunction_("transferfunction", "Transfer function")
    , workGroupSize_("wgsize", "Work group size", 128, 1, 2048)
    , useGLSharing_("glsharing", "Use OpenGL sharing", true)
    , importanceUniformGrid3D_(std::make_shared<ImportanceUniformGrid3D>(size3_t(1))) {
    addPort(minMaxUniformGrid3DInport_);
    volumeDifferenceInfoInport_.setOptional(true);
    addPort(volumeDifferenceInfoInport_);
    addPort(importanceUniformGrid3DOutport_);
    minMaxUniformGrid3DInport_.onChange([this]() { invalidationFlag_ = InvalidationReason((int)invalidationFlag_ | (int)InvalidationReason::Volume); });
    addProperty(incrementalImportance);
    addProperty(opacityWeight_);
    addProperty(opacityDiffWeight_);
    addProperty(colorWeight_);
    addProperty(colorDiffWeight_);
    addProperty(useAssociatedColor_);
    addProperty(TFPointEpsilon_);
    addProperty(transferFunction_);
    transferFunction_.onChange([this]() { invalidationFlag_ = InvalidationReason((int)invalidationFlag_ | (int)InvalidationReason::TransferFunction); });
    addProperty(workGroupSize_);
    addProperty(useGLSharing_);
    importanceUniformGrid3DOutport_.setData(importanceUniformGrid3D_);
}
vec4 MinMaxUniformGrid3DImportanceCLProcessor::tfPointColorDiff(const vec4& p1, const vec4& p2) {
    float a1 = useAssociatedColor_.get() ? p1.w : 1.f, a2 = useAssociatedColor_.get() ? p2.w : 1.f;
    return vec4(std::fabs(p2.x * a2 - p1.x * a1), std::fabs(p2.y * a2 - p1.y * a1), std::fabs(p2.z * a2 - p1.z * a1),
                std::fabs(p2.w * a2 - p1.w * a1));
}
void MinMaxUniformGrid3DImportanceCLProcessor::updateTransferFunctionData() {
    // the TF as a point list over [0,1] with explicit end points (:304-362)
    TransferFunction& tf = transferFunction_.get();
    std::vector<float> pos;
    std::vector<vec4> col;
    auto colorOf = [&](const TFPrimitive& p) {
        vec4 c = p.getColor();
        if (useAssociatedColor_.get()) c = vec4(c.x * c.w, c.y * c.w, c.z * c.w, c.w * c.w);
        return c;
    };
    if (tf.size() == 0) {
        pos = {0.f, 1.f};
        col = {vec4(0.f), vec4(0.f)};
    } else {
        if (tf.get(0).getPosition() > 0.0) { pos.push_back(0.f); col.push_back(colorOf(tf.get(0))); }
        for (size_t i = 0; i < tf.size(); ++i) { pos.push_back((float)tf.get(i).getPosition()); col.push_back(colorOf(tf.get(i))); }
        if (tf.get(tf.size() - 1).getPosition() < 1.0) { pos.push_back(1.f); col.push_back(colorOf(tf.get(tf.size() - 1))); }
    }
    tfPointImportanceSize_ = (int)pos.size();
    if ((int)tfPointPositions_.getSize() < tfPointImportanceSize_) {
        tfPointPositions_.setSize(pos.size());
        tfPointColors_.setSize(pos.size());
    }
    auto* P = tfPointPositions_.getEditableRAMRepresentation();
    auto* C = tfPointColors_.getEditableRAMRepresentation();
    std::copy(pos.begin(), pos.end(), P->begin());
    std::copy(col.begin(), col.end(), C->begin());
}
void MinMaxUniformGrid3DImportanceCLProcessor::updateTransferFunctionDifferenceData() {
    // |TF_new - TF_old| as a point list for the incremental classifier: a merge walk over the break points of both
    // functions, with the reference's rules for what is stored -- a pair of points only where the difference exceeds
    // TFPointEpsilon and one of them is visible, the moved zero-opacity first point, the last point of either function
    // replaced by (1, its colour), zero-difference end points at 0 and 1
    // (isc/processors/minmaxuniformgrid3dimportanceclprocessor.cpp:364-501; oracle/frame.py tf_difference_lists is the
    // checker's restatement, tests/test_host_processors.py compares the two lists).
    const TransferFunction& cur = transferFunction_.get();
    const TransferFunction& prev = prevTransferFunction_;
    const int nC = (int)cur.size(), nP = (int)prev.size();
    std::vector<float> pos;
    std::vector<vec4> col;
    const float eps = TFPointEpsilon_.get();
    auto differs = [eps](const vec4& c) {   // glm::any(glm::epsilonNotEqual(c, vec4(0), eps)): |c_k| >= eps
        return std::fabs(c.x) >= eps || std::fabs(c.y) >= eps || std::fabs(c.z) >= eps || std::fabs(c.w) >= eps;
    };
    // mix(a, b, t): the colour of the segment a -> b at t's position; glm::mix(vec4, vec4, double) = vec4(dvec4(x) + a * dvec4(y - x))
    auto colourAt = [](const TFPrimitive& a, const TFPrimitive& b, const TFPrimitive& t) {
        const double w = (t.getPosition() - a.getPosition()) / (b.getPosition() - a.getPosition());
        const vec4 x = a.getColor(), y = b.getColor();
        vec4 r;
        for (int k = 0; k < 4; ++k) r[k] = (float)((double)x[k] + w * (double)(y[k] - x[k]));
        return r;
    };
    if (nC == 0 || nP == 0) {
        // the reference reads point 0 of both functions; with both empty it stores two zero points at position 0
        pos = {0.f, nC == 0 && nP == 0 ? 0.f : 1.f};
        col = {vec4(0.f), vec4(0.f)};
    } else {
        const TFPrimitive first = cur.get(0), pfirst = prev.get(0);
        TFPrimitive p1(first < pfirst ? first.getPosition() : pfirst.getPosition(), tfPointColorDiff(first.getColor(), pfirst.getColor()));
        TFPrimitive p2 = p1;
        if (first.getPosition() != pfirst.getPosition() && first.getAlpha() == 0.f && pfirst.getAlpha() == 0.f) {
            // a moved first point with zero opacity
            if (first < pfirst) {
                const TFPrimitive a2 = cur.get((size_t)std::min(1, nC - 1));
                p2 = TFPrimitive(pfirst.getPosition(), tfPointColorDiff(pfirst.getColor(), colourAt(first, a2, pfirst)));
            } else {
                const TFPrimitive a2 = prev.get((size_t)std::min(1, nP - 1));
                p2 = TFPrimitive(first.getPosition(), tfPointColorDiff(first.getColor(), colourAt(pfirst, a2, first)));
            }
        }
        pos.push_back(0.f);
        col.push_back(p1.getPosition() > 0.0 && (first.getAlpha() > 0.f || pfirst.getAlpha() > 0.f) && differs(p1.getColor())
                          ? p1.getColor() : vec4(0.f));
        int id = 0, prevId = 0;
        while (id < nC || prevId < nP) {
            if ((differs(p1.getColor()) || differs(p2.getColor())) && (p1.getAlpha() > 0.f || p2.getAlpha() > 0.f)) {
                if (pos.size() == 1) {   // everything before was equal: open the segment
                    pos.push_back((float)p1.getPosition());
                    col.push_back(p1.getColor());
                }
                pos.push_back((float)p2.getPosition());
                col.push_back(p2.getColor());
            }
            const TFPrimitive a1 = cur.get((size_t)std::min(id, nC - 1));
            const TFPrimitive a2 = id + 1 < nC - 1 ? cur.get((size_t)id + 1) : TFPrimitive(1.0, cur.get((size_t)nC - 1).getColor());
            const TFPrimitive b1 = prev.get((size_t)std::min(prevId, nP - 1));
            const TFPrimitive b2 = prevId + 1 < nP - 1 ? prev.get((size_t)prevId + 1) : TFPrimitive(1.0, prev.get((size_t)nP - 1).getColor());
            p1 = p2;
            if (a2 < b2) {
                p2 = TFPrimitive(a2.getPosition(), tfPointColorDiff(a2.getColor(), colourAt(b1, b2, a2)));
                ++id;
            } else if (b2 < a2) {
                p2 = TFPrimitive(b2.getPosition(), tfPointColorDiff(b2.getColor(), colourAt(a1, a2, b2)));
                ++prevId;
            } else {
                p2 = TFPrimitive(a2.getAlpha() < b2.getAlpha() ? b2.getPosition() : a2.getPosition(),
                                 tfPointColorDiff(a2.getColor(), b2.getColor()));
                ++id;
                ++prevId;
            }
        }
        if (p2.getPosition() < 1.0 && p2.getAlpha() > 0.f) {
            pos.push_back((float)p2.getPosition());
            col.push_back(p2.getColor());
        }
        if (pos.back() < 1.f) {
            pos.push_back(1.f);
            col.push_back(vec4(0.f));
        }
    }
    tfPointImportanceSize_ = (int)pos.size();
    if ((int)tfPointPositions_.getSize() < tfPointImportanceSize_) {
        tfPointPositions_.setSize(pos.size());
        tfPointColors_.setSize(pos.size());
    }
    auto* P = tfPointPositions_.getEditableRAMRepresentation();
    auto* C = tfPointColors_.getEditableRAMRepresentation();
    std::copy(pos.begin(), pos.end(), P->begin());
    std::copy(col.begin(), col.end(), C->begin());
}
void MinMaxUniformGrid3DImportanceCLProcessor::process() {
    auto in = minMaxUniformGrid3DInport_.getData();
    auto minMax = dynamic_cast<const MinMaxUniformGrid3D*>(in.get());
    if (!minMax) {
        LogError("minMaxUniformGrid3DInport_ expects MinMaxUniformGrid3D as input");
        return;
    }
    if (minMax->getDimensions() != importanceUniformGrid3D_->getDimensions()) {
        importanceUniformGrid3D_->setDimensions(minMax->getDimensions());
        importanceUniformGrid3D_->setCellDimension(minMax->getCellDimension());
        importanceUniformGrid3D_->setModelMatrix(minMax->getModelMatrix());
        importanceUniformGrid3D_->setWorldMatrix(minMax->getWorldMatrix());
    }
    bool incrementalFormula = true;   // the static kernel is built with -D INCREMENTAL_TF_IMPORTANCE (:99-101)
    if ((int)invalidationFlag_ & (int)InvalidationReason::TransferFunction) {
        if (!prevTransferFunctionValid_ || prevTransferFunction_.size() == 0 || !incrementalImportance.get())
            updateTransferFunctionData();
        else
            updateTransferFunctionDifferenceData();
        prevTransferFunction_ = transferFunction_.get();
        prevTransferFunctionValid_ = true;
    } else if ((int)invalidationFlag_ & (int)InvalidationReason::Volume) {
        updateTransferFunctionData();
    }
    const size3_t dim = importanceUniformGrid3D_->getDimensions();
    const int n = (int)(dim.x * dim.y * dim.z);
    float wn = colorWeight_.get() + colorDiffWeight_.get() + opacityDiffWeight_.get() + opacityWeight_.get();
    if (wn <= 0.f) wn = 1.f;
    const float lab = 1.f / std::sqrt(100.f * 100.f + 500.f * 500.f + 400.f * 400.f);
    const float w[4] = {colorWeight_.get() * lab / wn, colorDiffWeight_.get() * lab / wn, opacityDiffWeight_.get() / wn, opacityWeight_.get() / wn};
    auto& rt = CpmRuntime::get();
    const float* pos = static_cast<const float*>(tfPointPositions_.deviceRead());
    const float* col = static_cast<const float*>(tfPointColors_.deviceRead());
    const uint16_t* mm = static_cast<const uint16_t*>(minMax->data.deviceRead());
    float* out = static_cast<float*>(importanceUniformGrid3D_->data.deviceWrite());
    auto prevMM = dynamic_cast<const MinMaxUniformGrid3D*>(prevMinMaxUniformGrid3D_.get());
    ScopedStage st("classify");
    if (volumeDifferenceInfoInport_.isReady() && prevMM && prevMM != minMax) {
        auto diff = dynamic_cast<const DynamicVolumeInfoUniformGrid3D*>(volumeDifferenceInfoInport_.getData().get());
        if (!diff) {
            LogError("volumeDifferenceInfoInport_ expects DynamicVolumeInfoUniformGrid3D as input");
            return;
        }
        incrementalFormula = false;   // the time-varying kernel is built without the define (Lab formula)
        rt.check(cpm_classify_importance(rt.ctx(), mm, static_cast<const uint16_t*>(prevMM->data.deviceRead()),
                                         static_cast<const float*>(diff->data.deviceRead()), n, pos, col, tfPointImportanceSize_, w,
                                         incrementalFormula ? 1 : 0, out));
    } else {
        rt.check(cpm_classify_importance(rt.ctx(), mm, nullptr, nullptr, n, pos, col, tfPointImportanceSize_, w, 1, out));
    }
    prevMinMaxUniformGrid3D_ = in;
    invalidationFlag_ = InvalidationReason(0);
    importanceUniformGrid3DOutport_.setData(importanceUniformGrid3D_);
}

// ---- RadixSortCL, RandomNumberGeneratorCL ---------------------------------------------------------------------
const ProcessorInfo RadixSortCL::processorInfo_{"org.inviwo.RadixSortCL", "Radix sort", "Sorting", "Experimental", "CL"};
RadixSortCL::RadixSortCL() : keysPort_("unsortedKeys"), inputPort_("unsortedData"), outputPort_("sortedData") {
    addPort(keysPort_);
    addPort(inputPort_);
    addPort(outputPort_);
}
void RadixSortCL::process() {
    auto keys = keysPort_.getData();
    auto data = inputPort_.getData();
    if (!keys || !data) return;
    try {
        radixSort_.enqueue(*const_cast<Buffer<unsigned int>*>(keys.get()), const_cast<Buffer<unsigned int>*>(data.get()), keys->getSize(), 0);
    } catch (std::invalid_argument& e) {
        LogError(e.what());
    } catch (CpmError& e) {
        LogError(e.what());
    }
    outputPort_.setData(data);   // pass-through of the input object, sorted in place (:244-247)
}

const ProcessorInfo RandomNumberGeneratorCL::processorInfo_{"org.inviwo.RandomNumberGeneratorCL", "Random Number Generator", "Random numbers",
                                                            "Experimental", "CL"};
RandomNumberGeneratorCL::RandomNumberGeneratorCL()
    : randomNumbersPort_("samples")
    , nRandomNumbers_("nSamples", "N samples", 256, 1, 100000000)
    , regenerateNumbers_("genRnd", "Regenerate")
    , seed_("seed", "Seed number", 0, 0, 2147483647)
    , workGroupSize_("wgsize", "Work group size", 256, 1, 2048)
    , useGLSharing_("glsharing", "Use OpenGL sharing", true)
    , randomNumbers_(std::make_shared<Buffer<float>>()) {
    addPort(randomNumbersPort_);
    addProperty(nRandomNumbers_);
    addProperty(regenerateNumbers_);
    addProperty(seed_);
    addProperty(workGroupSize_);
    addProperty(useGLSharing_);
    seed_.onChange([this]() { randomNumberGenerator_.setSeed((unsigned)seed_.get()); });
    randomNumbersPort_.setData(randomNumbers_);
}
void RandomNumberGeneratorCL::process() {
    if ((size_t)nRandomNumbers_.get() != randomNumbers_->getSize()) randomNumbers_->setSize((size_t)nRandomNumbers_.get());
    randomNumberGenerator_.generate(*randomNumbers_);
    randomNumbersPort_.setData(randomNumbers_);
}

const ProcessorInfo RandomNumberGenerator2DCL::processorInfo_{"org.inviwo.RandomNumberGenerator2DCL", "Random Number Generator 2D",
                                                              "Random numbers", "Stable", "CL"};
RandomNumberGenerator2DCL::RandomNumberGenerator2DCL()
    : randomNumbersPort_("samples")
    , nRandomNumbers_("nSamples", "N samples", ivec2{128, 128}, ivec2{2, 2}, ivec2{2048, 2048})
    , regenerateNumbers_("genRnd", "Regenerate")
    , seed_("seed", "Seed number", 0, 0, 2147483647)
    , workGroupSize_("wgsize", "Work group size", 256, 1, 2048)
    , useGLSharing_("glsharing", "Use OpenGL sharing", true)
    , image_(std::make_shared<ImageF32>()) {
    addPort(randomNumbersPort_);
    addProperty(nRandomNumbers_);
    addProperty(regenerateNumbers_);
    addProperty(seed_);
    addProperty(workGroupSize_);
    addProperty(useGLSharing_);
    nRandomNumbers_.onChange([this]() { nRandomNumbersChanged(); });
    randomNumbersPort_.setData(image_);
}
void RandomNumberGenerator2DCL::nRandomNumbersChanged() {
    const ivec2 n = nRandomNumbers_.get();
    if (n.x < 2 || n.y < 2 || n.x > 2048 || n.y > 2048) throw std::invalid_argument("nSamples out of range");
    randomState_.setSize((size_t)n.x * (size_t)n.y);
    MWC64XSeedGenerator().generateRandomSeeds(&randomState_, (unsigned)seed_.get(), false);
    image_->dims = n;
    image_->data.setSize((size_t)n.x * (size_t)n.y);
}
void RandomNumberGenerator2DCL::process() {
    if (image_->dims.x != nRandomNumbers_.get().x || image_->dims.y != nRandomNumbers_.get().y || randomState_.getSize() == 0)
        nRandomNumbersChanged();
    auto& rt = CpmRuntime::get();
    uint32_t* st = static_cast<uint32_t*>(const_cast<void*>(randomState_.deviceRead()));
    randomState_.deviceWrite();
    rt.check(cpm_rng_uniform(rt.ctx(), st, image_->data.getSize(), 1, static_cast<float*>(image_->data.deviceWrite())));
    randomNumbersPort_.setData(image_);
}

}  // namespace inviwo
