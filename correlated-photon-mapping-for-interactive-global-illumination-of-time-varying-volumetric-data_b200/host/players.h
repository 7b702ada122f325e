// players.h -- the sequence players of the uniformgridcl module: time -> (index, weight) bookkeeping and
// the interpolation between two neighbouring elements of a sequence, on the device.
//   org.inviwo.UniformGrid3DPlayerProcessor   ugc/processors/uniformgrid3dplayerprocessor.cpp:36-152  (BufferMixerCL, ping-pong)
//   org.inviwo.VolumeSequencePlayer           ugc/processors/volumesequenceplayer.cpp:36-180          (volume_mix.frag)
// The reference drives `time` from a Timer while "playSequence" is on; headless, onSequenceTimerEvent() is the tick.
#pragma once
#include "processors.h"

namespace inviwo {

// ugc/buffermixercl.h:49-71: out = mix(x, y, a) per element (mixKernel)
class BufferMixerCL {
public:
    template <typename T>
    void mix(const Buffer<T>& x, const Buffer<T>& y, float a, Buffer<T>& out);
};

// the time / index logic both players share (their property identifiers differ)
class SequenceClock {
public:
    SequenceClock(const char* timePerId, const char* timePerName, const char* rateId);
    FloatProperty time_;
    IntProperty index_;
    FloatProperty timePerElement_;
    IntProperty frameRate_;
    BoolProperty playSequence_;
    void addTo(Processor& p);
    void onSequenceTimerEvent();             // one timer tick: time += (1000 / frameRate) / 1000, wrapped at the maximum
    void updateIndex();                      // index = floor(time / timePerElement) % size + 1
    void onSequenceChange(size_t nElements); // ranges of time and index follow the sequence length
    float weight() const;                    // fractional part of time / timePerElement
    size_t step() const { return (size_t)(index_.get() - 1); }
};

class UniformGrid3DPlayerProcessor : public Processor {
public:
    UniformGrid3DPlayerProcessor();
    void process() override;
    const ProcessorInfo getProcessorInfo() const override { return processorInfo_; }
    static const ProcessorInfo processorInfo_;
    DataInport<UniformGrid3DVector> inport_;
    DataOutport<UniformGrid3DBase> outport_;
    SequenceClock clock_;
private:
    std::shared_ptr<UniformGrid3DBase> outData_, outDataPingPong_;
    BufferMixerCL bufferMixer_;
};

class VolumeSequencePlayer : public Processor {
public:
    VolumeSequencePlayer();
    void process() override;
    const ProcessorInfo getProcessorInfo() const override { return processorInfo_; }
    static const ProcessorInfo processorInfo_;
    DataInport<VolumeSequence> inport_;
    DataOutport<Volume> outport_;
    SequenceClock clock_;
private:
    std::shared_ptr<Volume> outVolume_;
};

}  // namespace inviwo
