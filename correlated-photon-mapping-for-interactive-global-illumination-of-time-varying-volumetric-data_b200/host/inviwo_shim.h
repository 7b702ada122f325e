// inviwo_shim.h -- the thin slice of the Inviwo API that the six reference modules use at their
// processor boundary, so that the drop-in processor classes in processors.h compile and run
// headless.  Against a real Inviwo checkout these types are replaced by the real ones; only
// the names and semantics used by the reference's process() methods are provided:
//   glm-like vectors/matrices, Buffer<T> with a RAM and a device representation (the role of
//   BufferRAM / BufferCL), Volume with its coordinate transformer and data map,
//   TransferFunction, Mesh, LightSource, properties, ports, Processor.
// Device memory goes through the C ABI (cpm_mem_*); there is no CUDA header dependency here.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "cpm_b200.h"

namespace inviwo {

// ---- math ------------------------------------------------------------------------------------
struct vec2 { float x = 0, y = 0; };
struct vec3 {
    float x = 0, y = 0, z = 0;
    vec3() = default;
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit vec3(float a) : x(a), y(a), z(a) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
struct vec4 {
    float x = 0, y = 0, z = 0, w = 0;
    vec4() = default;
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(vec3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    explicit vec4(float a) : x(a), y(a), z(a), w(a) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
struct ivec2 { int x = 0, y = 0; };
struct ivec3 { int x = 0, y = 0, z = 0; };
struct size3_t {
    size_t x = 0, y = 0, z = 0;
    size3_t() = default;
    size3_t(size_t a, size_t b, size_t c) : x(a), y(b), z(c) {}
    explicit size3_t(size_t a) : x(a), y(a), z(a) {}
    bool operator==(const size3_t& o) const { return x == o.x && y == o.y && z == o.z; }
    bool operator!=(const size3_t& o) const { return !(*this == o); }
};
struct uvec2 { uint32_t x = 0, y = 0; };
struct dvec2 { double x = 0, y = 1; };

inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, vec3 a) { return a * s; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
// glm::normalize(v) = v * inversesqrt(dot(v, v)), inversesqrt(x) = 1 / sqrt(x) (GLM 0.9.7 func_geometric.inl): the rounding
// of the reference's CPU light-plane fit depends on it (tests/test_ref_geometry.py pins it against the reference's files)
inline vec3 normalize(vec3 a) { float inv = 1.0f / std::sqrt(dot(a, a)); return {a.x * inv, a.y * inv, a.z * inv}; }

// column-major 4x4, m[c][r] like glm
struct mat4 {
    float m[16];
    mat4() { std::memset(m, 0, sizeof(m)); m[0] = m[5] = m[10] = m[15] = 1.f; }
    float* operator[](int c) { return m + 4 * c; }
    const float* operator[](int c) const { return m + 4 * c; }
    const float* data() const { return m; }
};
inline vec4 operator*(const mat4& M, vec4 v) {
    vec4 r;
    for (int i = 0; i < 4; ++i) r[i] = M[0][i] * v.x + M[1][i] * v.y + M[2][i] * v.z + M[3][i] * v.w;
    return r;
}
inline mat4 operator*(const mat4& A, const mat4& B) {
    mat4 R;
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) {
            float s = 0;
            for (int k = 0; k < 4; ++k) s += A[k][r] * B[c][k];
            R[c][r] = s;
        }
    return R;
}
mat4 inverse(const mat4& M);  // processors.cpp

// ---- runtime: the one context per process, the role of OpenCL::getPtr() --------------------------
class CpmError : public std::runtime_error {
public:
    CpmError(int code, const std::string& what) : std::runtime_error(what), code_(code) {}
    int code() const { return code_; }
private:
    int code_;
};

class CpmRuntime {
public:
    static CpmRuntime& get();            // creates the context on first use (device 0 or CPM_DEVICE)
    static void init(int device, void* stream = nullptr);   // explicit initialisation (one process per GPU)
    static void shutdown();
    // First photon id owned by this process when the photon set is sharded over GPUs: photon i of this
    // process uses MWC64X stream photonShardOffset + i (and the matching host base offset), so that the
    // shards of all GPUs together are the streams of one large single-GPU photon set.
    uint64_t photonShardOffset = 0;
    // multi-GPU: the communicator of this process (cpm_comm_init on ctx()) and whether volumes are ingested sharded --
    // rank r uploads slab r of a time step and the slabs are all-gathered over NVLink (cpm_comm_upload_volume_sharded)
    cpm_comm* comm = nullptr;
    bool shardedIngest = false;
    bool globalBudget = false;   // re-trace budget over the photons of all shards (cpm_comm_select_global) instead of per shard
    // slab upload possible for a volume of `bytes`?
    bool shardedUpload(size_t bytes) const {
        if (!comm || !shardedIngest) return false;
        const size_t w = (size_t)cpm_comm_world(comm);
        return w > 1 && bytes % w == 0 && (bytes / w) % 16 == 0;
    }
    cpm_ctx* ctx() { return ctx_; }
    void check(int rc) const {
        if (rc != CPM_OK) throw CpmError(rc, cpm_last_error(ctx_));
    }
    void sync() { check(cpm_ctx_sync(ctx_)); }
private:
    cpm_ctx* ctx_ = nullptr;
};
#define CPM_CHECK(call) ::inviwo::CpmRuntime::get().check(call)

// ---- stage profiler: CUDA events on the context stream, the role of IVW_OPENCL_PROFILING ---------
// Recording is asynchronous; totals are resolved (one event synchronise per pending pair) on query.
class StageProfiler {
public:
    static StageProfiler& get();
    bool enabled = false;
    std::string only;   // non-empty: only this stage is timed (two event records per frame instead of two per stage)
    void begin(const char* stage);
    void end();
    void resolve();
    void reset();
    double totalMs(const std::string& stage);
    int count(const std::string& stage);
    double lastMs(const std::string& stage);
    std::string stages();
    void releaseEvents();
private:
    struct Pending { std::string stage; cpm_event* a; cpm_event* b; };
    cpm_event* take();
    std::vector<cpm_event*> pool_;
    std::vector<Pending> pending_;
    std::vector<std::pair<std::string, cpm_event*>> open_;   // stages nest: an outer stage's time includes its inner stages
    struct Acc { double total = 0, last = 0; int n = 0; };
    std::map<std::string, Acc> acc_;
};
struct ScopedStage {
    explicit ScopedStage(const char* s) { StageProfiler::get().begin(s); }
    ~ScopedStage() { StageProfiler::get().end(); }
};

// ---- Buffer<T>: RAM + device representation with lazy synchronisation ----------------------------
enum class BufferUsage { Static, Dynamic };

class BufferBase {
public:
    virtual ~BufferBase() { releaseDevice(); }
    size_t getSize() const { return size_; }
    virtual size_t getSizeOfElement() const = 0;
    size_t getSizeInBytes() const { return size_ * getSizeOfElement(); }
    // device representation valid for reading (uploads when RAM is newer): getRepresentation<BufferCL>()
    const void* deviceRead();
    // device representation for writing (marks RAM stale): getEditableRepresentation<BufferCL>()
    void* deviceWrite();
    // number of bytes moved host<->device by the lazy synchronisation so far
    static size_t& h2dBytes() { static size_t b = 0; return b; }
    static size_t& d2hBytes() { static size_t b = 0; return b; }
protected:
    virtual void* ramPtr() = 0;
    void ensureDevice();
    void releaseDevice();
    void downloadIfStale();
    size_t size_ = 0;
    void* dev_ = nullptr;
    size_t devBytes_ = 0;
    bool ramValid_ = true, devValid_ = false;
};

template <typename T>
class Buffer : public BufferBase {
public:
    explicit Buffer(size_t n = 0, BufferUsage = BufferUsage::Static) { setSize(n); }
    Buffer(const Buffer&) = delete;
    Buffer& operator=(const Buffer&) = delete;
    size_t getSizeOfElement() const override { return sizeof(T); }
    // destructive resize, like Inviwo's Buffer::setSize
    void setSize(size_t n) {
        if (n == size_) return;
        size_ = n;
        ram_.assign(n, T());
        ramValid_ = true;
        devValid_ = false;
    }
    std::vector<T>* getEditableRAMRepresentation() {
        downloadIfStale();
        devValid_ = false;
        return &ram_;
    }
    const std::vector<T>* getRAMRepresentation() {
        downloadIfStale();
        return &ram_;
    }
protected:
    void* ramPtr() override { return ram_.data(); }
    std::vector<T> ram_;
};

// ---- Volume ---------------------------------------------------------------------------------------
enum class DataFormatId { UInt8, UInt16, Float32, Vec4Float32 };
struct DataFormatBase {
    DataFormatId id;
    size_t size;        // bytes per element
    size_t components;
    double maxValue;    // normalisation range of the type (1 for float)
    static const DataFormatBase* get(DataFormatId id);
    size_t getSize() const { return size; }
};

struct DataMapper {
    dvec2 dataRange, valueRange;
};

class Volume;
class StructuredCoordinateTransformer {
public:
    explicit StructuredCoordinateTransformer(const Volume* v) : v_(v) {}
    mat4 getTextureToIndexMatrix() const;  // index = tex*dim - 0.5
    mat4 getIndexToTextureMatrix() const;
    mat4 getTextureToWorldMatrix() const;  // world * model
    mat4 getWorldToDataMatrix() const { return inverse(getTextureToWorldMatrix()); }
private:
    const Volume* v_;
};

class Volume {
public:
    Volume(size3_t dim, const DataFormatBase* format);
    ~Volume();
    Volume(const Volume&) = delete;
    size3_t getDimensions() const { return dim_; }
    void setDimensions(size3_t d);   // destructive
    const DataFormatBase* getDataFormat() const { return format_; }
    mat4 getModelMatrix() const { return model_; }
    void setModelMatrix(const mat4& m) { model_ = m; }
    mat4 getWorldMatrix() const { return world_; }
    void setWorldMatrix(const mat4& m) { world_ = m; }
    StructuredCoordinateTransformer getCoordinateTransformer() const { return StructuredCoordinateTransformer(this); }
    DataMapper dataMap_;
    // VolumeRAM
    void* getEditableRAMData();          // marks the device copy stale
    const void* getRAMData();            // downloads when the device copy is newer
    // point the RAM representation at caller-owned (e.g. pinned) memory; marks the device copy stale
    void setExternalRAMData(void* ptr);
    // start uploading caller-owned pinned memory on the transfer stream; a later setExternalRAMData(ptr) with
    // the same pointer adopts the upload instead of copying again
    void prefetchExternalRAMData(void* ptr);
    size_t getSizeInBytes() const { return dim_.x * dim_.y * dim_.z * format_->size; }
    // VolumeCL: linear device buffer (+ the cpm_volume handle the kernels take)
    const void* deviceRead();
    void* deviceWrite();
    const cpm_volume* handle(int layout);   // CPM_VOLUME_LINEAR or CPM_VOLUME_TEXTURE; refreshed after uploads
    // (min, max) voxel value per cell of the tracer's opacity-bound grid (cpm_volume_value_range): a derived
    // device representation like the texture copy, recomputed after the voxels changed
    const float* valueRange(int cellLog2, size_t* nCells);
    void formatScaleOffset(float& scale, float& offset) const;
    uint64_t dataVersion() const { return version_; }   // changes whenever the voxels may have changed
private:
    void ensureDevice();
    void touch();
    void* range_ = nullptr;
    size_t rangeCells_ = 0;
    int rangeLog2_ = -1;
    bool rangeValid_ = false;
    uint64_t version_ = 0;
    size3_t dim_;
    const DataFormatBase* format_;
    mat4 model_, world_;
    std::vector<uint8_t> ram_;
    void* ext_ = nullptr;
    void* dev_ = nullptr;
    size_t devBytes_ = 0;
    bool ramValid_ = true, devValid_ = false;
    cpm_volume* lin_ = nullptr;
    cpm_volume* tex_ = nullptr;
    bool texValid_ = false;
    // format scale / offset the handles were created with (dataMap_.dataRange is a public member, as in Inviwo: a change
    // of the data range must reach the sampling arithmetic and everything derived from it, e.g. the opacity bound)
    float handleScale_ = 0.f, handleOffset_ = 0.f;
    void* prefetched_ = nullptr;
    cpm_event* prefetchDone_ = nullptr;
};
using VolumeSequence = std::vector<std::shared_ptr<Volume>>;

// ---- TransferFunction --------------------------------------------------------------------------------
class TFPrimitive {
public:
    TFPrimitive(double pos = 0.0, vec4 color = vec4(0.f)) : pos_(pos), color_(color) {}
    double getPosition() const { return pos_; }
    vec4 getColor() const { return color_; }
    float getAlpha() const { return color_.w; }
    bool operator<(const TFPrimitive& o) const { return pos_ < o.pos_; }
private:
    double pos_;
    vec4 color_;
};

class TransferFunction {
public:
    explicit TransferFunction(size_t textureSize = 1024) : data_(textureSize) {}
    TransferFunction(const TransferFunction& o) : points_(o.points_), data_(o.data_.getSize()) { dirty_ = true; }
    TransferFunction& operator=(const TransferFunction& o) {
        points_ = o.points_;
        dirty_ = true;
        return *this;
    }
    void clear() { points_.clear(); dirty_ = true; }
    void add(double pos, vec4 color);
    size_t size() const { return points_.size(); }
    const TFPrimitive& get(size_t i) const { return points_[i]; }
    size_t getTextureSize() const { return data_.getSize(); }
    // the RGBA float32 layer (piecewise linear between points, constant outside), device side
    const float* deviceData();
    const std::vector<vec4>& ramData();
    uint64_t version() { if (dirty_) rasterise(); return version_; }   // changes with the rasterised texels
private:
    void rasterise();
    uint64_t version_ = 0;
    std::vector<TFPrimitive> points_;
    Buffer<vec4> data_;
    bool dirty_ = true;
};

// ---- Mesh, light source ---------------------------------------------------------------------------------
class Mesh {
public:
    Buffer<vec3> vertices;
    Buffer<uint32_t> indices;
    mat4 modelMatrix, worldMatrix;
    mat4 getWorldToDataMatrix() const { return inverse(worldMatrix * modelMatrix); }
    static std::shared_ptr<Mesh> unitCube();   // the scene proxy geometry (texture-space cube)
};

enum class LightSourceType { LIGHT_AREA = 0, LIGHT_CONE, LIGHT_POINT, LIGHT_DIRECTIONAL };
class LightSource {
public:
    virtual ~LightSource() = default;
    virtual LightSourceType getLightSourceType() const = 0;
    vec3 getIntensity() const { return intensity_; }
    void setIntensity(vec3 i) { intensity_ = i; }
    mat4 getModelToWorldMatrix() const { return modelToWorld_; }
    void setModelToWorldMatrix(const mat4& m) { modelToWorld_ = m; }
protected:
    vec3 intensity_{1.f, 1.f, 1.f};
    mat4 modelToWorld_;
};
class DirectionalLight : public LightSource {
public:
    LightSourceType getLightSourceType() const override { return LightSourceType::LIGHT_DIRECTIONAL; }
    // orients the light's +z axis along `direction` and places its origin at `position` (world space)
    void set(vec3 position, vec3 direction);
};
class PointLight : public LightSource {
public:
    LightSourceType getLightSourceType() const override { return LightSourceType::LIGHT_POINT; }
    void setPosition(vec3 p);
};
struct PackedLightSource {
    mat4 tm;
    vec4 radiance;
    int type;
};
PackedLightSource baseLightToPackedLight(const LightSource* light, float radianceScale, const mat4& transformLightMat);

// ---- properties ------------------------------------------------------------------------------------------
class Property {
public:
    Property(std::string id, std::string name) : id_(std::move(id)), name_(std::move(name)) {}
    virtual ~Property() = default;
    const std::string& getIdentifier() const { return id_; }
    const std::string& getDisplayName() const { return name_; }
    void onChange(std::function<void()> cb) { callbacks_.push_back(std::move(cb)); }
    void setVisible(bool v) { visible_ = v; }
    void propertyModified() {
        for (auto& cb : callbacks_) cb();
        if (owner_) owner_();
    }
    void setOwnerInvalidate(std::function<void()> f) { owner_ = std::move(f); }
protected:
    std::string id_, name_;
    std::vector<std::function<void()>> callbacks_;
    std::function<void()> owner_;
    bool visible_ = true;
};

template <typename T>
class OrdinalProperty : public Property {
public:
    OrdinalProperty(std::string id, std::string name, T value, T minV, T maxV)
        : Property(std::move(id), std::move(name)), value_(value), min_(minV), max_(maxV) {}
    const T& get() const { return value_; }
    operator const T&() const { return value_; }
    void set(const T& v) { value_ = v; propertyModified(); }
    T getMinValue() const { return min_; }
    T getMaxValue() const { return max_; }
    void setMinValue(const T& v) { min_ = v; }
    void setMaxValue(const T& v) { max_ = v; }
    void setReadOnly(bool v) { readOnly_ = v; }
    bool getReadOnly() const { return readOnly_; }
private:
    T value_, min_, max_;
    bool readOnly_ = false;
};
using FloatProperty = OrdinalProperty<float>;
using IntProperty = OrdinalProperty<int>;
using IntVec2Property = OrdinalProperty<ivec2>;
using IntVec3Property = OrdinalProperty<ivec3>;
using IntMinMaxProperty = OrdinalProperty<ivec2>;

class BoolProperty : public Property {
public:
    BoolProperty(std::string id, std::string name, bool v) : Property(std::move(id), std::move(name)), value_(v) {}
    bool get() const { return value_; }
    operator bool() const { return value_; }
    void set(bool v) { if (v != value_) { value_ = v; propertyModified(); } }
private:
    bool value_;
};
class ButtonProperty : public Property {
public:
    using Property::Property;
    void pressButton() { propertyModified(); }
};
class TransferFunctionProperty : public Property {
public:
    TransferFunctionProperty(std::string id, std::string name, TransferFunction tf = TransferFunction())
        : Property(std::move(id), std::move(name)), tf_(tf) {}
    TransferFunction& get() { return tf_; }
    void set(const TransferFunction& tf) { tf_ = tf; propertyModified(); }
private:
    TransferFunction tf_;
};
template <typename T>
class OptionProperty : public Property {
public:
    using Property::Property;
    void addOption(std::string id, std::string name, T value) { options_.push_back({std::move(id), std::move(name), value}); }
    const T& get() const { return options_[selected_].value; }
    const std::string& getSelectedIdentifier() const { return options_[selected_].id; }
    void setSelectedValue(const T& v) {
        for (size_t i = 0; i < options_.size(); ++i)
            if (options_[i].value == v) { selected_ = i; propertyModified(); return; }
        throw std::invalid_argument("no option with that value");
    }
    void setSelectedIdentifier(const std::string& id) {
        for (size_t i = 0; i < options_.size(); ++i)
            if (options_[i].id == id) { selected_ = i; propertyModified(); return; }
        throw std::invalid_argument("unknown option " + id);
    }
private:
    struct Opt { std::string id, name; T value; };
    std::vector<Opt> options_;
    size_t selected_ = 0;
};

// phase function enum values follow Inviwo's ShadingFunctionKind order used by shading.cl
enum class ShadingFunctionKind { HenyeyGreenstein = 0, Schlick, BlinnPhong, Ward, CookTorrance, AbcMicrofacet, Ashikhmin, Mix, Isotropic, None };
class AdvancedMaterialProperty : public Property {
public:
    AdvancedMaterialProperty(std::string id, std::string name);
    OptionProperty<int> phaseFunctionProp;
    FloatProperty indexOfRefractionProp, roughnessProp, anisotropyProp;
    OrdinalProperty<vec4> specularColorProp;
    vec4 getCombinedMaterialParameters() const;   // [0] = anisotropy g (see cpm_trace_params.material)
    int getPhaseFunctionEnum() const;             // CPM_PHASE_*
    void phaseFunctionChanged() {}
};
struct Camera {
    vec3 lookFrom{0, 0, -2}, lookTo{0, 0, 0}, lookUp{0, 1, 0};
};
class CameraProperty : public Property {
public:
    CameraProperty(std::string id, std::string name) : Property(std::move(id), std::move(name)) {}
    const Camera& get() const { return cam_; }
    void set(const Camera& c) { cam_ = c; propertyModified(); }
private:
    Camera cam_;
};

// ---- ports and processors ----------------------------------------------------------------------------------
enum class InvalidationLevel { Valid, InvalidOutput };
class Processor;

class Port {
public:
    explicit Port(std::string id) : id_(std::move(id)) {}
    virtual ~Port() = default;
    const std::string& getIdentifier() const { return id_; }
    Processor* owner = nullptr;
protected:
    std::string id_;
};

template <typename T>
class DataOutport : public Port {
public:
    using Port::Port;
    template <typename U>
    void setData(std::shared_ptr<U> d) { setDataImpl(std::shared_ptr<const T>(std::move(d))); }
    void setDataImpl(std::shared_ptr<const T> d);
    std::shared_ptr<const T> getData() const { return data_; }
    bool hasData() const { return data_ != nullptr; }
    std::vector<std::function<void()>> listeners;   // connected inports
private:
    std::shared_ptr<const T> data_;
};

template <typename T>
class DataInport : public Port {
public:
    using Port::Port;
    void connectTo(DataOutport<T>* out);
    void disconnect() { src_ = nullptr; }
    void setOptional(bool o) { optional_ = o; }
    bool isOptional() const { return optional_; }
    bool isConnected() const { return src_ != nullptr; }
    bool isReady() const { return src_ && src_->hasData(); }
    bool hasData() const { return isReady(); }
    std::shared_ptr<const T> getData() const { return src_ ? src_->getData() : nullptr; }
    void onChange(std::function<void()> cb) { onChange_.push_back(std::move(cb)); }
    void onConnect(std::function<void()> cb) { onConnect_.push_back(std::move(cb)); }
private:
    DataOutport<T>* src_ = nullptr;
    bool optional_ = false;
    std::vector<std::function<void()>> onChange_, onConnect_;
};

template <typename T>
class MultiDataInport : public Port {
public:
    using Port::Port;
    void connectTo(DataOutport<T>* out);
    bool isReady() const {
        if (srcs_.empty()) return false;
        for (auto* s : srcs_) if (!s->hasData()) return false;
        return true;
    }
    std::vector<std::shared_ptr<const T>> getVectorData() const {
        std::vector<std::shared_ptr<const T>> v;
        for (auto* s : srcs_) v.push_back(s->getData());
        return v;
    }
    void onChange(std::function<void()> cb) { onChange_.push_back(std::move(cb)); }
private:
    std::vector<DataOutport<T>*> srcs_;
    std::vector<std::function<void()>> onChange_;
};

struct ProcessorInfo {
    std::string classIdentifier, displayName, category, codeState, tags;
};

class Processor {
public:
    virtual ~Processor() = default;
    virtual void process() = 0;
    virtual const ProcessorInfo getProcessorInfo() const = 0;
    void addPort(Port& p) { p.owner = this; ports_.push_back(&p); }
    void addProperty(Property& p) {
        props_[p.getIdentifier()] = &p;
        p.setOwnerInvalidate([this]() { invalidate(InvalidationLevel::InvalidOutput); });
    }
    Property* getPropertyByIdentifier(const std::string& id) {
        auto it = props_.find(id);
        return it == props_.end() ? nullptr : it->second;
    }
    Port* getPort(const std::string& id) {
        for (auto* p : ports_) if (p->getIdentifier() == id) return p;
        return nullptr;
    }
    std::vector<std::string> getPropertyIdentifiers() const {
        std::vector<std::string> v;
        for (auto& kv : props_) v.push_back(kv.first);
        return v;
    }
    std::vector<std::string> getPortIdentifiers() const {
        std::vector<std::string> v;
        for (auto* p : ports_) v.push_back(p->getIdentifier());
        return v;
    }
    void invalidate(InvalidationLevel) { valid_ = false; }
    bool isValid() const { return valid_; }
    void setValid() { valid_ = true; }
    // evaluate if invalid (what Inviwo's ProcessorNetworkEvaluator does for one processor)
    bool evaluate() {
        if (valid_) return false;
        valid_ = true;   // process() may re-invalidate (progressive refinement)
        process();
        return true;
    }
private:
    std::vector<Port*> ports_;
    std::map<std::string, Property*> props_;
    bool valid_ = false;
};

template <typename T>
void DataOutport<T>::setDataImpl(std::shared_ptr<const T> d) {
    data_ = std::move(d);
    for (auto& l : listeners) l();
}
template <typename T>
void DataInport<T>::connectTo(DataOutport<T>* out) {
    src_ = out;
    out->listeners.push_back([this]() {
        for (auto& cb : onChange_) cb();
        if (owner) owner->invalidate(InvalidationLevel::InvalidOutput);
    });
    for (auto& cb : onConnect_) cb();
    if (owner) owner->invalidate(InvalidationLevel::InvalidOutput);
}
template <typename T>
void MultiDataInport<T>::connectTo(DataOutport<T>* out) {
    srcs_.push_back(out);
    out->listeners.push_back([this]() {
        for (auto& cb : onChange_) cb();
        if (owner) owner->invalidate(InvalidationLevel::InvalidOutput);
    });
    if (owner) owner->invalidate(InvalidationLevel::InvalidOutput);
}

void LogError(const std::string& msg);
void LogInfo(const std::string& msg);

}  // namespace inviwo
