// processors.h -- drop-in host side of the correlated photon-mapping path: the reference's port
// data types, kernel-launcher classes and Inviwo processors, re-implemented on top of the C ABI
// (include/cpm_b200.h).  Class identifiers, port identifiers, property identifiers, defaults and
// ranges are the reference's so that a saved workspace keeps working.
#pragma once
#include <cstdlib>
#include "inviwo_shim.h"

namespace inviwo {

// ======================================================================== port data types ====
// ppm/photondata.h:47-56
struct Photon {
    vec3 pos;
    vec3 power;
    vec2 encodedDirection;   // theta = acos(z), phi = atan2(y, x)
    void setDirection(vec3 dir);
    vec3 getDirection() const;
};
static_assert(sizeof(Photon) == 32, "photon records are float8");

// ppm/photondata.h:58-63
struct RecomputedPhotonIndices {
    Buffer<unsigned int> indicesToRecomputedPhotons;
    int nRecomputedPhotons = -1;   // -1 means uninitialised / "all"
    bool isInitialized() const { return nRecomputedPhotons != -1; }
    void setUninitialized() { nRecomputedPhotons = -1; }
};

// ppm/photondata.h:65-156, ppm/photondata.cpp
class PhotonData {
public:
    enum class InvalidationReason {
        Camera = 1 << 0, TransferFunction = 1 << 1, Light = 1 << 2, Progressive = 1 << 3, Volume = 1 << 4,
        All = Camera | TransferFunction | Light | Progressive | Volume
    };
    void setSize(size_t numberOfPhotons, int maxPhotonInteractions);
    size_t getNumberOfPhotons() const { return photons_.getSize() / (2 * maxPhotonInteractions_); }
    int getMaxPhotonInteractions() const { return maxPhotonInteractions_; }
    void setRadius(double radiusRelativeToSceneSize, double sceneRadius);
    void setRadius(double radius) { worldSpaceRadius_ = radius; }
    void advanceToNextIteration(double alpha = 0.5);
    double getRadiusRelativeToSceneSize() const { return getRadius() / sceneRadius_; }
    double getRadius() const { return worldSpaceRadius_; }
    static double progressiveSphereRadius(double radius, int iteration, double alpha);
    double getSceneRadius() const { return sceneRadius_; }
    void resetIteration() { iteration_ = 0; }
    bool isReset() const { return iteration_ <= 1; }
    int iteration() const { return iteration_; }
    void setIteration(int v) { iteration_ = v; }
    static double sphereVolume(double radius);
    double getRelativeIrradianceScale() const;
    InvalidationReason getInvalidationReason() const { return invalidationFlag_; }
    void setInvalidationReason(InvalidationReason v) { invalidationFlag_ = v; }

    Buffer<vec4> photons_;   // 2 vec4 per photon record, interaction-major
    static const float defaultRadiusRelativeToSceneRadius;
    static const float defaultSceneRadius;
    static const double scaleToMakeLightPowerOfOneVisibleForDirectionalLightSource;
    static const int defaultNumberOfPhotons = 256 * 256;
protected:
    int maxPhotonInteractions_ = 1;
    double sceneRadius_ = 1.0;
    double worldSpaceRadius_ = 0.01;
    int iteration_ = 0;
    InvalidationReason invalidationFlag_ = InvalidationReason::All;
};
inline PhotonData::InvalidationReason operator|(PhotonData::InvalidationReason a, PhotonData::InvalidationReason b) {
    return static_cast<PhotonData::InvalidationReason>(static_cast<int>(a) | static_cast<int>(b));
}
inline PhotonData::InvalidationReason& operator|=(PhotonData::InvalidationReason& a, PhotonData::InvalidationReason b) {
    return a = a | b;
}

// lcl/lightsample.h:52-115.  The host class has a vtable, so the reference sizes its byte buffer
// with 40 B per sample while the device stride is 32 B; the byte size is kept for layout parity.
class LightSample {
public:
    virtual ~LightSample() = default;
    vec3 origin, power;
    vec2 encodedDirection;
};
class LightSamples {
public:
    explicit LightSamples(size_t nSamples = 0) { setSize(nSamples); }
    Buffer<unsigned char>* getLightSamples() { return &lightSamples_; }
    const Buffer<unsigned char>* getLightSamples() const { return &lightSamples_; }
    Buffer<vec2>* getIntersectionPoints() { return &intersectionPoints_; }
    const Buffer<vec2>* getIntersectionPoints() const { return &intersectionPoints_; }
    void setSize(size_t nSamples);
    size_t getSize() const { return lightSamples_.getSize() / sizeof(LightSample); }
    void resetIteration() { iteration_ = 0; }
    void advanceIteration() { ++iteration_; }
    bool isReset() const { return iteration_ <= 1; }
    size_t getIteration() const { return iteration_; }
private:
    mutable Buffer<unsigned char> lightSamples_;
    mutable Buffer<vec2> intersectionPoints_;
    size_t iteration_ = 0;
};
using SampleBuffer = Buffer<vec4>;   // lcl/sample.h:54  (u, v, w, pdf)

// ugc/uniformgrid3d.h:63-198
class UniformGrid3DBase {
public:
    explicit UniformGrid3DBase(size3_t cellDimension = size3_t(1)) : cellDimension_(cellDimension) {}
    virtual ~UniformGrid3DBase() = default;
    virtual size3_t getDimensions() const = 0;
    virtual void setDimensions(const size3_t& dim) = 0;
    virtual size_t getSizeInBytes() const = 0;
    size3_t getCellDimension() const { return cellDimension_; }
    void setCellDimension(size3_t v) { cellDimension_ = v; }
    mat4 getModelMatrix() const { return model_; }
    void setModelMatrix(const mat4& m) { model_ = m; }
    mat4 getWorldMatrix() const { return world_; }
    void setWorldMatrix(const mat4& m) { world_ = m; }
    // what the .u3d reader / writer need (ugc/uniformgrid3d.h:84-98): raw element storage, its format name
    // ("FLOAT32", "Vec2UINT16") and an empty grid of the same type
    virtual void* getData() = 0;
    virtual const char* getFormatString() const = 0;
    virtual std::shared_ptr<UniformGrid3DBase> cloneEmpty() const = 0;
private:
    size3_t cellDimension_;
    mat4 model_, world_;
};
struct u16vec2 { uint16_t x = 0, y = 0; };
template <typename T> struct GridFormatName;
template <> struct GridFormatName<float> { static const char* get() { return "FLOAT32"; } };
template <> struct GridFormatName<u16vec2> { static const char* get() { return "Vec2UINT16"; } };
template <typename T>
class UniformGrid3D : public UniformGrid3DBase {
public:
    explicit UniformGrid3D(size3_t cellDimension = size3_t(1)) : UniformGrid3DBase(cellDimension) {}
    UniformGrid3D(size3_t gridDimensions, size3_t cellDimension) : UniformGrid3DBase(cellDimension) { setDimensions(gridDimensions); }
    size3_t getDimensions() const override { return dimensions_; }
    void setDimensions(const size3_t& dim) override {
        dimensions_ = dim;
        data.setSize(dim.x * dim.y * dim.z);
    }
    size_t getSizeInBytes() const override { return data.getSizeInBytes(); }
    void* getData() override { return data.getEditableRAMRepresentation()->data(); }
    const char* getFormatString() const override { return GridFormatName<T>::get(); }
    std::shared_ptr<UniformGrid3DBase> cloneEmpty() const override {
        auto g = std::make_shared<UniformGrid3D<T>>(dimensions_, getCellDimension());
        g->setModelMatrix(getModelMatrix());
        g->setWorldMatrix(getWorldMatrix());
        return g;
    }
    mutable Buffer<T> data;   // id = x + y*dx + z*dx*dy
private:
    size3_t dimensions_;
};
using MinMaxUniformGrid3D = UniformGrid3D<u16vec2>;                 // ugc/minmaxuniformgrid3d.h:42
using ImportanceUniformGrid3D = UniformGrid3D<float>;              // isc/importanceuniformgrid3d.h:46
using DynamicVolumeInfoUniformGrid3D = UniformGrid3D<float>;       // ugc/processors/dynamicvolumedifferenceanalysis.h:60-61
using UniformGrid3DVector = std::vector<std::shared_ptr<UniformGrid3DBase>>;

// ".u3d" uniform-grid sequences on disk: a text header (RawFile, Resolution x y z t, Format, ModelMatrix,
// WorldMatrix, CellDimensions) next to a raw file holding the t grids back to back.
// ugc/uniformgrid3dreader.cpp:59-183, ugc/uniformgrid3dwriter.cpp:47-102.
class UniformGrid3DReader {
public:
    std::shared_ptr<UniformGrid3DVector> readData(const std::string& filePath);
};
class UniformGrid3DWriter {
public:
    void writeData(const UniformGrid3DVector* data, const std::string& filePath) const;
    void setOverwrite(bool v) { overwrite_ = v; }
private:
    bool overwrite_ = true;
};

// ===================================================================== kernel launchers ======
// rng/mwc64xseedgenerator.h:60
class MWC64XSeedGenerator {
public:
    void generateRandomSeeds(Buffer<uvec2>* buffer, unsigned int seed, bool useGLSharing = true, size_t localWorkGroupSize = 256);
};
// rng/mwc64xrandomnumbergenerator.h
class MWC64XRandomNumberGenerator {
public:
    void setSeed(unsigned int s) { seed_ = s; dirty_ = true; }
    void generate(Buffer<float>& randomNumbersOut);
private:
    Buffer<uvec2> randomState_;
    unsigned int seed_ = 0;
    bool dirty_ = true;
};
// lcl/samplegenerator2dcl.h:53-88 / isc/uniformsamplegenerator2dcl.h
class SampleGenerator2DCL {
public:
    virtual ~SampleGenerator2DCL() = default;
    virtual void reset() = 0;
    virtual void generateNextSamples(SampleBuffer& positionSamplesOut) = 0;
};
class UniformSampleGenerator2DCL : public SampleGenerator2DCL {
public:
    void reset() override {}
    void generateNextSamples(SampleBuffer& positionSamplesOut) override;
};
namespace geometry {
// lcl/orientedboundingbox2d.cpp:80-100 (+ convexhull2d.cpp, pointplaneprojection.cpp): CPU fit of the light plane
struct PlaneFit { vec3 origin, u, v; };
PlaneFit fitPlaneAlignedOrientedBoundingBox2D(const std::vector<vec3>& points, vec3 planePoint, vec3 planeNormal);
std::vector<vec2> convexHull2D(std::vector<vec2> points);
}  // namespace geometry
// lcl/directionallightsamplercl.h:71
class DirectionalLightSamplerCL {
public:
    void sampleLightSource(const Mesh* mesh, const SampleBuffer* samples, const LightSource* light, LightSamples& lightSamplesOut);
    // the kernel arguments of the last call (lcl/directionallightsamplercl.cpp:66-73), kept for parity checks
    struct Setup { vec3 direction, planePoint, origin, u, v, radiance; float area = 0.f; } lastSetup;
};
// lcl/lightsamplemeshintersectioncl.h:64
class LightSampleMeshIntersectionCL {
public:
    void meshSampleIntersection(const Mesh* mesh, LightSamples* samples);
};
// ppm/photontracercl.h:69-104
class PhotonTracerCL {
public:
    // volume layout used for sampling: CPM_VOLUME_TEXTURE (default) or CPM_VOLUME_LINEAR
    int volumeLayout = CPM_VOLUME_TEXTURE;
    void tracePhotons(const Volume* volume, TransferFunction& transferFunction, const vec4 aabb[2],
                      const AdvancedMaterialProperty& material, float stepSize, const LightSamples* lightSamples,
                      Buffer<unsigned int>* photonsToRecomputeIndices, int nInvalidPhotons, int photonOffset, int batch,
                      int maxInteractions, PhotonData* photonOutData);
    // refresh of the per-cell opacity bound of (volume, transfer function) if either changed; tracePhotons calls it,
    // and a caller may do so earlier to overlap it with something else (idempotent)
    void prepareOpacityBound(const Volume* volume, TransferFunction& transferFunction);
    void setRandomSeedSize(size_t nPhotons);
    void setNoSingleScattering(bool v) { onlyMultipleScattering_ = v; }
    void setProgressive(bool v) { progressive_ = v; }
    bool isProgressive() const { return progressive_; }
    bool isValid() const { return true; }
    unsigned long long* collisionCounter = nullptr;   // optional device counter (benchmarks)
    // per-cell opacity bound (cpm_opacity_bound): collision tests it decides skip the voxel fetch, results unchanged
    bool useOpacityBound = true;
    int boundCellLog2 = 3;
    // the bound grid also as a 3-D texture for the tracer (cpm_bound_tex): CPM_BOUND_TEXTURE=0 in the environment
    // keeps the linear look-up (A/B runs)
    bool useBoundTexture = getenv("CPM_BOUND_TEXTURE") ? atoi(getenv("CPM_BOUND_TEXTURE")) != 0 : true;
    ~PhotonTracerCL() { if (boundTex_) cpm_bound_tex_destroy(boundTex_); }
private:
    Buffer<float> opacityBound_;
    cpm_bound_tex* boundTex_ = nullptr;
    int boundTexDims_[3] = {0, 0, 0};
    uint64_t boundVolumeVersion_ = 0, boundTfVersion_ = 0;
    Buffer<uvec2> randomState_;
    bool onlyMultipleScattering_ = false, progressive_ = false;
};
// ppm/photonrecomputationdetector.h:55-87
class PhotonRecomputationDetector {
public:
    void photonRecomputationImportance(const PhotonData* photonData, int photonOffset, const Volume* origVolume,
                                       const ImportanceUniformGrid3D* uniformGridVolume, const LightSamples& lightSamples,
                                       Buffer<unsigned int>& recomputationImportance);
    bool getEqualImportance() const { return equalImportance_; }
    void setEqualImportance(bool v) { equalImportance_ = v; }
    int getPercentage() const { return percentage_; }
    void setPercentage(int v) { percentage_ = v; }
    int getIteration() const { return iteration_; }
    void setIteration(int v) { iteration_ = v; }
    // not in the reference: CPM_DETECT_FIX_EXIT -- exit points of photons that left the volume are computed as intended
    // (origin + tEnd * dir for interaction 0, entry + t * dir later) instead of the reference's
    // ppm/cl/photonrecomputationdetector.cl:128 / :137 arithmetic
    bool getFixExitPoint() const { return fixExitPoint_; }
    void setFixExitPoint(bool v) { fixExitPoint_ = v; }
    bool isValid() const { return true; }
private:
    bool equalImportance_ = false, fixExitPoint_ = false;
    int percentage_ = 100, iteration_ = 0;
};
// clogs::Radixsort (rsc/ext/clogs/radixsort.h) for uint keys / uint-or-no values
class Radixsort {
public:
    explicit Radixsort(bool hasValues) : hasValues_(hasValues) {}
    void enqueue(Buffer<unsigned int>& keys, Buffer<unsigned int>* values, size_t elements, unsigned int maxBits = 0);
private:
    bool hasValues_;
    Buffer<unsigned int> tmpKeys_, tmpValues_;
};

// =========================================================================== processors ======
// org.inviwo.UniformSampleGenerator2DCL -- isc/processors/uniformsamplegenerator2dprocessorcl.cpp:41-96
class UniformSampleGenerator2DProcessorCL : public Processor {
public:
    UniformSampleGenerator2DProcessorCL();
    void process() override;
    const ProcessorInfo getProcessorInfo() const override { return processorInfo_; }
    static const ProcessorInfo processorInfo_;
    DataOutport<SampleBuffer> samplesPort_, directionalSamplesPort_;
    IntVec2Property nSamples_, workGroupSize_;
    BoolProperty useGLSharing_;
private:
    std::shared_ptr<SampleBuffer> samples_, directionalSamples_;
    UniformSampleGenerator2DCL sampleGenerator_;
};

// org.inviwo.DirectionalLightSamplerCL -- lcl/processors/directionallightsamplerclprocessor.cpp:38-89
class DirectionalLightSamplerCLProcessor : public Processor {
public:
    DirectionalLightSamplerCLProcessor();
    void process() override;
    const ProcessorInfo getProcessorInfo() const override { return processorInfo_; }
    static const ProcessorInfo processorInfo_;
    DataInport<Mesh> boundingVolumeInport_;
    DataInport<SampleBuffer> samplesInport_;
    DataInport<LightSource> lightInport_;
    DataOutport<LightSamples> lightSamplesOutport_;
    IntProperty workGroupSize_;
    BoolProperty useGLSharing_;
    const DirectionalLightSamplerCL& sampler() const { return lightSampler_; }
    const LightSamples* lightSamples() const { return lightSamples_.get(); }
private:
    std::shared_ptr<LightSamples> lightSamples_;
    DirectionalLightSamplerCL lightSampler_;
    LightSampleMeshIntersectionCL intersector_;
};

// org.inviwo.ProgressivePhotonTracerCL -- ppm/processor/progressivephotontracercl.cpp:61-746
class ProgressivePhotonTracerCL : public Processor {
public:
    ProgressivePhotonTracerCL();
    void process() override;
    const ProcessorInfo getProcessorInfo() const override { return processorInfo_; }
    static const ProcessorInfo processorInfo_;

    DataInport<Volume> volumePort_;
    DataInport<UniformGrid3DBase> recomputationImportanceGrid_;
    MultiDataInport<LightSamples> lightSamples_;
    DataOutport<PhotonData> outport_;
    DataOutport<RecomputedPhotonIndices> recomputedIndicesPort_;

    FloatProperty samplingRate_, radius_, sceneRadianceScaling_;
    CameraProperty camera_;
    FloatProperty maxIncrementalPhotonsToUpdate_;
    BoolProperty equalIncrementalImportance_, spatialSorting_;
    IntProperty maxScatteringEvents_;
    BoolProperty noSingleScattering_;
    TransferFunctionProperty transferFunction_;
    AdvancedMaterialProperty advancedMaterial_;
    FloatProperty alphaProp_;
    IntVec2Property workGroupSize_;
    BoolProperty useGLSharing_;
    ButtonProperty invalidateRendering_;
    BoolProperty enableProgressiveRefinement_, enableProgressivePhotonRecomputation_;
    IntMinMaxProperty clipX_, clipY_, clipZ_;
    BoolProperty fixDetectorExitPoint_;   // extension (default off = the reference's arithmetic): see PhotonRecomputationDetector

    void invalidateProgressiveRendering(PhotonData::InvalidationReason r) { invalidationFlag_ |= r; }
    void onTimerEvent();    // the reference's 100 ms Timer callback; call it to step progressive work
    PhotonTracerCL& tracer() { return photonTracer_; }
    int remainingPhotonsToUpdate() const { return remainingPhotonsToUpdate_; }
    Buffer<unsigned int>& importanceKeys() { return photonRecomputationImportance_; }   // for parity checks
    // stage timings (detector, count+iota, sort, indexsort, trace), like the reference's
    // IVW_DETAILED_PROFILING log (:562-598), are collected by StageProfiler when it is enabled
private:
    void onClipChange();
    void progressiveRefinementChanged();
    float getSceneRadius() const;
    void resetPhotonImportance(size_t offset, size_t nPhotons);

    std::shared_ptr<PhotonData> photonData_;
    vec4 aabb_[2];
    PhotonTracerCL photonTracer_;
    PhotonRecomputationDetector photonRecomputationDetector_;
    PhotonData::InvalidationReason invalidationFlag_ = PhotonData::InvalidationReason::All;
    std::shared_ptr<RecomputedPhotonIndices> recomputedPhotonIndices_;
    Buffer<unsigned int> photonRecomputationImportance_;
    Buffer<unsigned int> sortedImportance_;     // the keys are sorted on a copy (Appendix A of SURVEY.md)
    Radixsort recomputationImportanceSorter_{true};
    Radixsort recomputationIndexSorter_{false};
    int remainingPhotonsToUpdate_ = -1;
    // global (cross-shard) budget, CpmRuntime::globalBudget: budget and photons left over ALL shards, position in the global order
    long long globalBudget_ = 0, remainingGlobal_ = 0, globalOffset_ = 0;
    int remainingPhotonsOffset_ = 0;
    bool selectionIsSorted_ = false;   // the id list came from cpm_select_below: ascending, complete
};

// org.inviwo.PhotonToLightVolumeProcessorCL -- ppm/processor/photontolightvolumeprocessorcl.cpp:45-509
class PhotonToLightVolumeProcessorCL : public Processor {
public:
    PhotonToLightVolumeProcessorCL();
    void process() override;
    const ProcessorInfo getProcessorInfo() const override { return processorInfo_; }
    static const ProcessorInfo processorInfo_;
    DataInport<Volume> volumeInport_;
    DataInport<PhotonData> photons_;
    DataInport<RecomputedPhotonIndices> recomputedPhotonIndicesPort_;
    // one-shot: a cudaEvent_t the next write to the light volume waits for (a consumer on another stream, e.g. the
    // multi-GPU exchange, is still reading the volume); cleared once waited for
    void* waitBeforeLightVolumeWrite = nullptr;
    DataOutport<Volume> outport_;
    FloatProperty incrementalRecomputationThreshold_;
    OptionProperty<int> volumeSizeOption_;
    OptionProperty<int> volumeDataTypeOption_;   // value = channels
    BoolProperty alignChangedPhotons_;
    IntProperty workGroupSize_;
    BoolProperty useGLSharing_;
    // the reference launches the full splat over N*I work-items but bounds them with N, so only
    // interaction 0 is splatted on the full path (:304,368).  true = reproduce, false = all interactions
    bool referenceFullSplatBound = true;
    std::string lastPath;   // "full", "incremental" or "none" (for tests)
private:
    void volumeSizeOptionChanged();
    std::shared_ptr<Volume> lightVolume_;
    Buffer<vec4> prevPhotons_;
    Buffer<vec4> changedAlignedPhotons_;   // packed -old / +new records of the `alignChangedPhotons` path
};

// org.inviwo.VolumeMinMaxCLProcessor -- ugc/processors/volumeminmaxclprocessor.cpp:45-184
class VolumeMinMaxCLProcessor : public Processor {
public:
    VolumeMinMaxCLProcessor();
    void process() override;
    const ProcessorInfo getProcessorInfo() const override { return processorInfo_; }
    static const ProcessorInfo processorInfo_;
    DataInport<Volume> inport_;
    DataInport<VolumeSequence> vectorInport_;
    DataOutport<UniformGrid3DBase> outport_;
    DataOutport<UniformGrid3DVector> vectorOutport_;
    IntProperty volumeRegionSize_;
    IntVec3Property workGroupSize_;
    BoolProperty useGLSharing_;
    std::unique_ptr<MinMaxUniformGrid3D> compute(const Volume* volume);
};

// org.inviwo.DynamicVolumeDifferenceAnalysis -- ugc/processors/dynamicvolumedifferenceanalysis.cpp:36-104
class DynamicVolumeDifferenceAnalysis : public Processor {
public:
    DynamicVolumeDifferenceAnalysis();
    void process() override;
    const ProcessorInfo getProcessorInfo() const override { return processorInfo_; }
    static const ProcessorInfo processorInfo_;
    DataInport<VolumeSequence> inport_;
    DataOutport<UniformGrid3DVector> outport_;
    IntProperty volumeRegionSize_;
    // one pair of the sequence loop (dynamicvolumedifferenceanalysis.cpp:66-101), on the device
    static std::shared_ptr<DynamicVolumeInfoUniformGrid3D> difference(Volume* cur, Volume* next, size_t region);
};

// org.inviwo.MinMaxUniformGrid3DImportanceCLProcessor -- isc/processors/minmaxuniformgrid3dimportanceclprocessor.cpp:43-524
class MinMaxUniformGrid3DImportanceCLProcessor : public Processor {
public:
    enum class InvalidationReason { TransferFunction = 1 << 0, Volume = 1 << 1, All = 3 };
    MinMaxUniformGrid3DImportanceCLProcessor();
    void process() override;
    const ProcessorInfo getProcessorInfo() const override { return processorInfo_; }
    static const ProcessorInfo processorInfo_;
    DataInport<UniformGrid3DBase> minMaxUniformGrid3DInport_;
    DataInport<UniformGrid3DBase> volumeDifferenceInfoInport_;
    DataOutport<UniformGrid3DBase> importanceUniformGrid3DOutport_;
    BoolProperty incrementalImportance;
    FloatProperty opacityWeight_, opacityDiffWeight_, colorWeight_, colorDiffWeight_;
    BoolProperty useAssociatedColor_;
    FloatProperty TFPointEpsilon_;
    TransferFunctionProperty transferFunction_;
    IntProperty workGroupSize_;
    BoolProperty useGLSharing_;
    // host TF point lists, exposed for tests
    const std::vector<float>& tfPointPositions() { return *tfPointPositions_.getRAMRepresentation(); }
    const std::vector<vec4>& tfPointColors() { return *tfPointColors_.getRAMRepresentation(); }
    int tfPointImportanceSize() const { return tfPointImportanceSize_; }
    // host-only hook for parity tests: the difference list of (cur, prev) as a transfer-function change would build it
    void buildDifferenceLists(const TransferFunction& cur, const TransferFunction& prev) {
        transferFunction_.set(cur);
        prevTransferFunction_ = prev;
        prevTransferFunctionValid_ = true;
        updateTransferFunctionDifferenceData();
    }
private:
    void updateTransferFunctionData();
    void updateTransferFunctionDifferenceData();
    vec4 tfPointColorDiff(const vec4& p1, const vec4& p2);
    std::shared_ptr<ImportanceUniformGrid3D> importanceUniformGrid3D_;
    std::shared_ptr<const UniformGrid3DBase> prevMinMaxUniformGrid3D_;
    TransferFunction prevTransferFunction_;
    bool prevTransferFunctionValid_ = false;
    Buffer<float> tfPointPositions_;
    Buffer<vec4> tfPointColors_;
    int tfPointImportanceSize_ = 0;
    InvalidationReason invalidationFlag_ = InvalidationReason::All;
};

// org.inviwo.RadixSortCL -- rsc/processors/radixsortcl.cpp:40-248 (uint keys, uint data)
class RadixSortCL : public Processor {
public:
    RadixSortCL();
    void process() override;
    const ProcessorInfo getProcessorInfo() const override { return processorInfo_; }
    static const ProcessorInfo processorInfo_;
    DataInport<Buffer<unsigned int>> keysPort_, inputPort_;
    DataOutport<Buffer<unsigned int>> outputPort_;
private:
    Radixsort radixSort_{true};
};

// org.inviwo.RandomNumberGeneratorCL -- rng/processors/randomnumbergeneratorcl.cpp:41-94
class RandomNumberGeneratorCL : public Processor {
public:
    RandomNumberGeneratorCL();
    void process() override;
    const ProcessorInfo getProcessorInfo() const override { return processorInfo_; }
    static const ProcessorInfo processorInfo_;
    DataOutport<Buffer<float>> randomNumbersPort_;
    IntProperty nRandomNumbers_;
    ButtonProperty regenerateNumbers_;
    IntProperty seed_, workGroupSize_;
    BoolProperty useGLSharing_;
private:
    std::shared_ptr<Buffer<float>> randomNumbers_;
    MWC64XRandomNumberGenerator randomNumberGenerator_;
};

// org.inviwo.RandomNumberGenerator2DCL -- rng/processors/randomnumbergenerator2dcl.cpp:44-136.  One MWC64X stream per
// pixel, one number per stream and evaluation (N_NUMBERS_PER_THREAD = 1, rng/cl/randomnumbergenerator.cl:32,51-71): pixel
// (x, y) is number y * width + x of the 1-D generator with the same seed.  As in the reference the streams are seeded
// when nSamples changes (nRandomNumbersChanged), not when `seed` changes.
struct ImageF32 {
    ivec2 dims{0, 0};
    Buffer<float> data;     // row-major, dims.x * dims.y
};
class RandomNumberGenerator2DCL : public Processor {
public:
    RandomNumberGenerator2DCL();
    void process() override;
    const ProcessorInfo getProcessorInfo() const override { return processorInfo_; }
    static const ProcessorInfo processorInfo_;
    DataOutport<ImageF32> randomNumbersPort_;
    IntVec2Property nRandomNumbers_;
    ButtonProperty regenerateNumbers_;
    IntProperty seed_, workGroupSize_;
    BoolProperty useGLSharing_;
private:
    void nRandomNumbersChanged();
    std::shared_ptr<ImageF32> image_;
    Buffer<uvec2> randomState_;
};

}  // namespace inviwo
