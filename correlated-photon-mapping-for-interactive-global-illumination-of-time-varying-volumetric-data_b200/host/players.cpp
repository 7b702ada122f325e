// players.cpp -- see players.h
#include "players.h"

#include <cmath>

namespace inviwo {

template <typename T>
struct MixFormat;
template <> struct MixFormat<float> { static constexpr int fmt = CPM_FMT_F32; static constexpr size_t comps = 1; };
template <> struct MixFormat<u16vec2> { static constexpr int fmt = CPM_FMT_U16; static constexpr size_t comps = 2; };

template <typename T>
void BufferMixerCL::mix(const Buffer<T>& x, const Buffer<T>& y, float a, Buffer<T>& out) {
    if (x.getSize() != y.getSize()) throw std::invalid_argument("BufferMixerCL::mix: buffers differ in size");
    if (out.getSize() != x.getSize()) out.setSize(x.getSize());
    auto& rt = CpmRuntime::get();
    ScopedStage st("mix");
    rt.check(cpm_mix(rt.ctx(), const_cast<Buffer<T>&>(x).deviceRead(), const_cast<Buffer<T>&>(y).deviceRead(), a,
                     x.getSize() * MixFormat<T>::comps, MixFormat<T>::fmt, out.deviceWrite()));
}
template void BufferMixerCL::mix<float>(const Buffer<float>&, const Buffer<float>&, float, Buffer<float>&);
template void BufferMixerCL::mix<u16vec2>(const Buffer<u16vec2>&, const Buffer<u16vec2>&, float, Buffer<u16vec2>&);

// ---- clock ---------------------------------------------------------------------------------------------------------------
SequenceClock::SequenceClock(const char* timePerId, const char* timePerName, const char* rateId)
    : time_("time", "Time", 0.f, 0.f, 0.f)
    , index_("selectedSequenceIndex", "Sequence index", 1, 1, 1)
    , timePerElement_(timePerId, timePerName, 1.f, 0.01f, 10.f)
    , frameRate_(rateId, "Frame rate", 10, 1, 60)
    , playSequence_("playSequence", "Play Sequence", false) {
    index_.setReadOnly(true);
    time_.onChange([this]() { updateIndex(); });
    playSequence_.onChange([this]() { time_.setReadOnly(playSequence_.get()); });
}
void SequenceClock::addTo(Processor& p) {
    p.addProperty(time_);
    p.addProperty(index_);
    p.addProperty(timePerElement_);
    p.addProperty(frameRate_);
    p.addProperty(playSequence_);
}
float SequenceClock::weight() const {
    float whole;
    return std::modf(time_.get() / timePerElement_.get(), &whole);
}
void SequenceClock::updateIndex() {
    float whole;
    std::modf(time_.get() / timePerElement_.get(), &whole);
    const size_t step = static_cast<size_t>(whole) % (size_t)index_.getMaxValue();
    if ((int)step != index_.get() - 1) index_.set((int)step + 1);
}
void SequenceClock::onSequenceTimerEvent() {
    float t = time_.get() + static_cast<float>(1000 / frameRate_.get()) / 1000.f;   // integer milliseconds per tick, as the Timer
    if (t > time_.getMaxValue()) t -= time_.getMaxValue();
    time_.set(t);
    updateIndex();
}
void SequenceClock::onSequenceChange(size_t n) {
    time_.setMaxValue(time_.getMinValue() + static_cast<float>(n - 1) * timePerElement_.get());
    if (time_.get() > time_.getMaxValue()) time_.set(time_.getMinValue());
    index_.setMaxValue((int)n);
    if (index_.get() > index_.getMaxValue()) index_.set(index_.getMinValue());
}

// ---- UniformGrid3DPlayerProcessor -----------------------------------------------------------------------------------------
const ProcessorInfo UniformGrid3DPlayerProcessor::processorInfo_{"org.inviwo.UniformGrid3DPlayerProcessor", "Uniform Grid 3D Player Processor",
                                                                 "UniformGrid3D", "Experimental", "CL"};
UniformGrid3DPlayerProcessor::UniformGrid3DPlayerProcessor()
    : inport_("Sequence"), outport_("InterpolatedData"), clock_("timePerElement", "Time Per element (s)", "frameRate") {
    addPort(inport_);
    addPort(outport_);
    clock_.addTo(*this);
    inport_.onChange([this]() { if (inport_.hasData()) clock_.onSequenceChange(inport_.getData()->size()); });
    clock_.timePerElement_.onChange([this]() { if (inport_.hasData()) clock_.onSequenceChange(inport_.getData()->size()); });
}
void UniformGrid3DPlayerProcessor::process() {
    auto elements = inport_.getData();
    if (!elements || elements->empty()) return;
    const float t = clock_.weight();
    const size_t step = clock_.step(), next = (step + 1) % elements->size();
    if (elements->size() < 2) {
        outport_.setData(std::shared_ptr<const UniformGrid3DBase>((*elements)[step]));
        return;
    }
    // two output grids alternate, so that a consumer that compares "previous" with "current" (the importance
    // processor keeps the grid of its last evaluation) never sees its previous grid overwritten
    std::swap(outData_, outDataPingPong_);
    auto in0 = (*elements)[step], in1 = (*elements)[next];
    if (!outData_ || outData_->getDimensions() != in0->getDimensions() ||
        std::string(outData_->getFormatString()) != in0->getFormatString()) {
        outData_ = in0->cloneEmpty();
    }
    if (auto* a = dynamic_cast<UniformGrid3D<float>*>(in0.get())) {
        bufferMixer_.mix(a->data, dynamic_cast<UniformGrid3D<float>&>(*in1).data, t, dynamic_cast<UniformGrid3D<float>&>(*outData_).data);
    } else if (auto* b = dynamic_cast<UniformGrid3D<u16vec2>*>(in0.get())) {
        bufferMixer_.mix(b->data, dynamic_cast<UniformGrid3D<u16vec2>&>(*in1).data, t, dynamic_cast<UniformGrid3D<u16vec2>&>(*outData_).data);
    } else {
        throw std::invalid_argument("UniformGrid3DPlayerProcessor: unsupported grid format");
    }
    outport_.setData(std::shared_ptr<const UniformGrid3DBase>(outData_));
}

// ---- VolumeSequencePlayer ---------------------------------------------------------------------------------------------------
const ProcessorInfo VolumeSequencePlayer::processorInfo_{"org.inviwo.VolumeSequencePlayer", "Volume Sequence Player", "Volume",
                                                         "Experimental", "GL"};
VolumeSequencePlayer::VolumeSequencePlayer()
    : inport_("volumeSequence"), outport_("InterpolatedVolume"), clock_("timePerVolume", "Time Per Volume (s)", "volumesPerSecond") {
    addPort(inport_);
    addPort(outport_);
    clock_.addTo(*this);
    inport_.onChange([this]() { if (inport_.hasData()) clock_.onSequenceChange(inport_.getData()->size()); });
    clock_.timePerElement_.onChange([this]() { if (inport_.hasData()) clock_.onSequenceChange(inport_.getData()->size()); });
}
void VolumeSequencePlayer::process() {
    auto volumes = inport_.getData();
    if (!volumes || volumes->empty()) return;
    const float t = clock_.weight();
    const size_t step = clock_.step(), next = (step + 1) % volumes->size();
    if (volumes->size() < 2) {
        outport_.setData(std::shared_ptr<const Volume>((*volumes)[step]));
        return;
    }
    Volume* v0 = (*volumes)[step].get();
    Volume* v1 = (*volumes)[next].get();
    if (!outVolume_ || outVolume_->getDimensions() != v0->getDimensions() || outVolume_->getDataFormat() != v0->getDataFormat()) {
        outVolume_ = std::make_shared<Volume>(v0->getDimensions(), v0->getDataFormat());
        outVolume_->setModelMatrix(v0->getModelMatrix());
        outVolume_->setWorldMatrix(v0->getWorldMatrix());
        outVolume_->dataMap_ = v0->dataMap_;
    }
    // volume_mix.frag: result = mix(texture(volume, p), texture(volume1, p), weight) rendered slice by slice into the output
    // texture; here one pass over the voxels (normalised integer textures round to nearest on the way out)
    const DataFormatBase* f = v0->getDataFormat();
    const int fmt = f->id == DataFormatId::UInt8 ? CPM_FMT_U8 : (f->id == DataFormatId::UInt16 ? CPM_FMT_U16 : CPM_FMT_F32);
    const size3_t d = v0->getDimensions();
    auto& rt = CpmRuntime::get();
    ScopedStage st("mix");
    rt.check(cpm_mix_unorm(rt.ctx(), v0->deviceRead(), v1->deviceRead(), t, d.x * d.y * d.z * f->components, fmt, outVolume_->deviceWrite()));
    outport_.setData(std::shared_ptr<const Volume>(outVolume_));
}

}  // namespace inviwo
