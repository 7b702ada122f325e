// host_capi.cpp -- headless driver of the drop-in processor network (see host_capi.h).
#include "host_capi.h"
#include "players.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <sstream>

#include "processors.h"
#include "workspace.h"

using namespace inviwo;

static thread_local std::string g_err;

struct cpmh_network {
    cpmh_config cfg;
    std::shared_ptr<Volume> volume;
    DataOutport<Volume> volumeSource{"data"};
    std::shared_ptr<Mesh> proxy;
    DataOutport<Mesh> meshSource{"geometry"};
    std::vector<std::shared_ptr<DirectionalLight>> lights;
    std::vector<std::unique_ptr<DataOutport<LightSource>>> lightSources;
    UniformSampleGenerator2DProcessorCL sampleGen;
    std::vector<std::unique_ptr<DirectionalLightSamplerCLProcessor>> lightSamplers;
    VolumeMinMaxCLProcessor minMax;
    MinMaxUniformGrid3DImportanceCLProcessor importance;
    ProgressivePhotonTracerCL tracer;
    PhotonToLightVolumeProcessorCL toLightVolume;
    // time series plumbing (the reference's sequence players / selectors are GUI processors)
    std::shared_ptr<VolumeSequence> sequence;
    std::vector<std::shared_ptr<UniformGrid3DBase>> seqMinMax, seqDiff;
    DataOutport<UniformGrid3DBase> minMaxSelector{"selectedMinMax"}, diffSelector{"selectedDiff"};
    bool useSequence = false;
    // streaming time steps (cpmh_network_stream_timestep_host): previous / current / incoming volumes rotate
    std::shared_ptr<Volume> streamVol[3];
    std::shared_ptr<MinMaxUniformGrid3D> streamMinMax[3];
    std::shared_ptr<DynamicVolumeInfoUniformGrid3D> streamDiff[3];
    int streamSlot = -1;
    unsigned long long* collisionCounter = nullptr;
    Buffer<float> lightVolumeSum;   // cpmh_network_sum_light_volume
    Buffer<float> readbackStage[2];  // cpmh_network_read_light_volume_async
    cpm_event* readbackDone[2] = {nullptr, nullptr};
    int readbackTurn = 0, readbackLast = -1;
};

template <typename F>
static int guarded(F&& f) {
    try {
        return f();
    } catch (CpmError& e) {
        g_err = e.what();
        return e.code();
    } catch (std::exception& e) {
        g_err = e.what();
        return CPM_E_INVALID;
    }
}

extern "C" {

const char* cpmh_last_error(void) { return g_err.c_str(); }

int cpmh_runtime_init(int device, void* stream, uint64_t photon_shard_offset) {
    return guarded([&]() {
        CpmRuntime::init(device, stream);
        CpmRuntime::get().photonShardOffset = photon_shard_offset;
        return (int)CPM_OK;
    });
}

int cpmh_runtime_set_photon_shard_offset(uint64_t photon_shard_offset) {
    return guarded([&]() {
        CpmRuntime::get().photonShardOffset = photon_shard_offset;
        return (int)CPM_OK;
    });
}

void* cpmh_runtime_ctx(void) {
    void* c = nullptr;
    guarded([&]() { c = CpmRuntime::get().ctx(); return (int)CPM_OK; });
    return c;
}

int cpmh_runtime_set_comm(void* comm, int sharded_ingest) {
    return guarded([&]() {
        auto& rt = CpmRuntime::get();
        rt.sync();
        rt.comm = static_cast<cpm_comm*>(comm);
        rt.shardedIngest = comm != nullptr && sharded_ingest != 0;
        return (int)CPM_OK;
    });
}

int cpmh_runtime_set_global_budget(int on) {
    return guarded([&]() {
        CpmRuntime::get().globalBudget = on != 0;
        return (int)CPM_OK;
    });
}

int cpmh_network_sum_light_volume(cpmh_network* net, float* out_host, size_t n_floats, void** sum_device) {
    return guarded([&]() {
        auto v = std::const_pointer_cast<Volume>(net->toLightVolume.outport_.getData());
        if (!v || v->getSizeInBytes() != n_floats * sizeof(float)) throw std::invalid_argument("light volume size mismatch");
        auto& rt = CpmRuntime::get();
        const float* local = static_cast<const float*>(v->deviceRead());
        const float* result = local;
        if (rt.comm && cpm_comm_world(rt.comm) > 1) {
            if (net->lightVolumeSum.getSize() != n_floats) net->lightVolumeSum.setSize(n_floats);
            float* sum = static_cast<float*>(net->lightVolumeSum.deviceWrite());
            ScopedStage st("exchange");
            rt.check(cpm_allreduce_lightvol(rt.comm, local, sum, n_floats));
            result = sum;
        }
        if (sum_device) *sum_device = const_cast<float*>(result);
        if (out_host) {
            rt.check(cpm_mem_copy_d2h(rt.ctx(), out_host, result, n_floats * sizeof(float)));
            rt.sync();
            BufferBase::d2hBytes() += n_floats * sizeof(float);
        }
        return (int)CPM_OK;
    });
}

int cpmh_network_read_light_volume_async(cpmh_network* net, float* out_host, size_t n_floats, int sum_over_ranks) {
    return guarded([&]() {
        auto v = std::const_pointer_cast<Volume>(net->toLightVolume.outport_.getData());
        if (!v || v->getSizeInBytes() != n_floats * sizeof(float)) throw std::invalid_argument("light volume size mismatch");
        auto& rt = CpmRuntime::get();
        const int i = net->readbackTurn;
        net->readbackTurn ^= 1;
        if (net->readbackDone[i]) {
            // the staging buffer is reused: its previous read-back must have left it
            rt.check(cpm_ctx_wait_event(rt.ctx(), net->readbackDone[i]));
            cpm_event_destroy(rt.ctx(), net->readbackDone[i]);
            net->readbackDone[i] = nullptr;
        }
        if (net->readbackStage[i].getSize() != n_floats) net->readbackStage[i].setSize(n_floats);
        float* stage = static_cast<float*>(net->readbackStage[i].deviceWrite());
        const float* local = static_cast<const float*>(v->deviceRead());
        if (sum_over_ranks && rt.comm && cpm_comm_world(rt.comm) > 1) {
            ScopedStage st("exchange");
            rt.check(cpm_allreduce_lightvol(rt.comm, local, stage, n_floats));
        } else if (out_host) {
            rt.check(cpm_mem_copy_d2d(rt.ctx(), stage, local, n_floats * sizeof(float)));   // the next frame updates `local`
        }
        net->readbackLast = -1;
        if (out_host) {
            rt.check(cpm_mem_readback_d2h(rt.ctx(), out_host, stage, n_floats * sizeof(float), &net->readbackDone[i]));
            BufferBase::d2hBytes() += n_floats * sizeof(float);
            net->readbackLast = i;
        }
        return (int)CPM_OK;
    });
}

int cpmh_network_wait_readback(cpmh_network* net) {
    return guarded([&]() {
        auto& rt = CpmRuntime::get();
        const int i = net->readbackLast;
        if (i >= 0 && net->readbackDone[i]) rt.check(cpm_event_sync(rt.ctx(), net->readbackDone[i]));
        return (int)CPM_OK;
    });
}

int cpmh_network_create(const cpmh_config* cfg, cpmh_network** out) {
    return guarded([&]() {
        if (!cfg || !out) throw std::invalid_argument("null argument");
        if (cfg->n_lights < 1 || cfg->n_lights > 8) throw std::invalid_argument("n_lights must be 1..8");
        CpmRuntime::init(cfg->device);
        auto* n = new cpmh_network();
        n->cfg = *cfg;
        DataFormatId fid = cfg->format == CPM_FMT_U8 ? DataFormatId::UInt8 : (cfg->format == CPM_FMT_U16 ? DataFormatId::UInt16 : DataFormatId::Float32);
        n->volume = std::make_shared<Volume>(size3_t(cfg->dims[0], cfg->dims[1], cfg->dims[2]), DataFormatBase::get(fid));
        n->proxy = Mesh::unitCube();
        n->meshSource.setData(n->proxy);
        n->sampleGen.nSamples_.set(ivec2{cfg->samples_per_side, cfg->samples_per_side});
        for (int l = 0; l < cfg->n_lights; ++l) {
            auto light = std::make_shared<DirectionalLight>();
            vec3 d = normalize(vec3(cfg->light_directions[l][0], cfg->light_directions[l][1], cfg->light_directions[l][2]));
            light->set(vec3(0.5f, 0.5f, 0.5f) - 2.f * d, d);
            light->setIntensity(vec3(cfg->light_intensity[l][0], cfg->light_intensity[l][1], cfg->light_intensity[l][2]));
            n->lights.push_back(light);
            n->lightSources.emplace_back(new DataOutport<LightSource>("light"));
            n->lightSamplers.emplace_back(new DirectionalLightSamplerCLProcessor());
            auto& ls = *n->lightSamplers.back();
            ls.boundingVolumeInport_.connectTo(&n->meshSource);
            ls.samplesInport_.connectTo(&n->sampleGen.samplesPort_);
            ls.lightInport_.connectTo(n->lightSources.back().get());
            n->lightSources.back()->setData(std::shared_ptr<const LightSource>(light));
            n->tracer.lightSamples_.connectTo(&ls.lightSamplesOutport_);
        }
        n->tracer.volumePort_.connectTo(&n->volumeSource);
        n->tracer.maxScatteringEvents_.set(cfg->max_scattering_events);
        n->tracer.radius_.set(cfg->photon_radius_voxels > 0 ? cfg->photon_radius_voxels : 1.f);
        if (cfg->max_incremental_percent > 0) n->tracer.maxIncrementalPhotonsToUpdate_.set(cfg->max_incremental_percent);
        n->tracer.tracer().volumeLayout = cfg->volume_layout;
        n->tracer.tracer().useOpacityBound = cfg->opacity_bound_cell_log2 >= 0;
        if (cfg->opacity_bound_cell_log2 > 0) n->tracer.tracer().boundCellLog2 = cfg->opacity_bound_cell_log2;
        n->toLightVolume.volumeInport_.connectTo(&n->volumeSource);
        n->toLightVolume.photons_.connectTo(&n->tracer.outport_);
        n->toLightVolume.recomputedPhotonIndicesPort_.connectTo(&n->tracer.recomputedIndicesPort_);
        n->toLightVolume.referenceFullSplatBound = cfg->reference_full_splat_bound != 0;
        if (cfg->incremental_threshold_percent > 0) n->toLightVolume.incrementalRecomputationThreshold_.set(cfg->incremental_threshold_percent);
        if (cfg->light_volume_channels == 4) n->toLightVolume.volumeDataTypeOption_.setSelectedIdentifier("4xfloat32");
        const char* opt = cfg->light_volume_option == 1 ? "1" : (cfg->light_volume_option == 2 ? "1/2" : (cfg->light_volume_option == 4 ? "1/4" : "radius"));
        n->toLightVolume.volumeSizeOption_.setSelectedIdentifier(opt);
        if (cfg->with_importance_grid) {
            n->minMax.inport_.connectTo(&n->volumeSource);
            n->importance.minMaxUniformGrid3DInport_.connectTo(&n->minMax.outport_);
            n->tracer.recomputationImportanceGrid_.connectTo(&n->importance.importanceUniformGrid3DOutport_);
        }
        n->volumeSource.setData(n->volume);
        bool anyClip = false;
        for (int k = 0; k < 6; ++k) anyClip |= cfg->clip[k] != 0;
        if (anyClip) {
            n->tracer.clipX_.set(ivec2{cfg->clip[0], cfg->clip[1]});
            n->tracer.clipY_.set(ivec2{cfg->clip[2], cfg->clip[3]});
            n->tracer.clipZ_.set(ivec2{cfg->clip[4], cfg->clip[5]});
        }
        *out = n;
        return (int)CPM_OK;
    });
}

void cpmh_network_destroy(cpmh_network* net) {
    if (!net) return;
    try {
        CpmRuntime::get().sync();
        for (int i = 0; i < 2; ++i)
            if (net->readbackDone[i]) {
                cpm_event_sync(CpmRuntime::get().ctx(), net->readbackDone[i]);
                cpm_event_destroy(CpmRuntime::get().ctx(), net->readbackDone[i]);
            }
    } catch (...) {
    }
    delete net;
}

int cpmh_network_set_transfer_function(cpmh_network* net, const float* pts, int n) {
    return guarded([&]() {
        TransferFunction tf;
        for (int i = 0; i < n; ++i) tf.add(pts[5 * i], vec4(pts[5 * i + 1], pts[5 * i + 2], pts[5 * i + 3], pts[5 * i + 4]));
        net->tracer.transferFunction_.set(tf);
        net->importance.transferFunction_.set(tf);
        return (int)CPM_OK;
    });
}

int cpmh_network_set_volume_host(cpmh_network* net, const void* voxels) {
    return guarded([&]() {
        net->useSequence = false;
        net->volume->setExternalRAMData(const_cast<void*>(voxels));
        net->volumeSource.setData(net->volume);   // notifies connected inports: Volume invalidation
        return (int)CPM_OK;
    });
}

int cpmh_network_set_sequence_host(cpmh_network* net, const void* const* voxels, int T) {
    return guarded([&]() {
        auto seq = std::make_shared<VolumeSequence>();
        const cpmh_config& c = net->cfg;
        DataFormatId fid = c.format == CPM_FMT_U8 ? DataFormatId::UInt8 : (c.format == CPM_FMT_U16 ? DataFormatId::UInt16 : DataFormatId::Float32);
        for (int t = 0; t < T; ++t) {
            auto v = std::make_shared<Volume>(size3_t(c.dims[0], c.dims[1], c.dims[2]), DataFormatBase::get(fid));
            v->setExternalRAMData(const_cast<void*>(voxels[t]));
            v->deviceRead();   // keep the whole series resident in HBM ...
            v->handle(net->tracer.tracer().volumeLayout);   // ... in the layout the tracer samples (no allocation inside a frame)
            if (net->tracer.tracer().useOpacityBound) v->valueRange(net->tracer.tracer().boundCellLog2, nullptr);   // per-step, like the min-max grids
            seq->push_back(v);
        }
        CpmRuntime::get().sync();
        net->sequence = seq;
        net->seqMinMax.clear();
        net->seqDiff.clear();
        // VolumeMinMaxCLProcessor on every step and DynamicVolumeDifferenceAnalysis between steps
        for (int t = 0; t < T; ++t) net->seqMinMax.emplace_back(std::shared_ptr<UniformGrid3DBase>(net->minMax.compute((*seq)[t].get()).release()));
        DynamicVolumeDifferenceAnalysis diff;
        DataOutport<VolumeSequence> src("data");
        diff.inport_.connectTo(&src);
        src.setData(std::shared_ptr<const VolumeSequence>(seq));
        diff.process();
        net->seqDiff = *diff.outport_.getData();
        // re-route the importance processor to the selected step's grids
        net->importance.minMaxUniformGrid3DInport_.connectTo(&net->minMaxSelector);
        net->importance.volumeDifferenceInfoInport_.connectTo(&net->diffSelector);
        net->useSequence = true;
        return (int)CPM_OK;
    });
}

int cpmh_network_set_timestep(cpmh_network* net, int t) {
    return guarded([&]() {
        if (!net->useSequence || !net->sequence || t < 0 || t >= (int)net->sequence->size()) throw std::invalid_argument("bad time step");
        const int T = (int)net->sequence->size();
        // difference between the previous step and this one is stored at index of the earlier step
        net->diffSelector.setData(std::shared_ptr<const UniformGrid3DBase>(net->seqDiff[(t + T - 1) % T]));
        net->minMaxSelector.setData(std::shared_ptr<const UniformGrid3DBase>(net->seqMinMax[t]));
        net->volumeSource.setData((*net->sequence)[t]);
        return (int)CPM_OK;
    });
}

int cpmh_network_set_data_range(cpmh_network* net, double lo, double hi) {
    return guarded([&]() {
        if (!(hi > lo)) throw std::invalid_argument("empty data range");
        auto apply = [&](Volume* v) { if (v) v->dataMap_.dataRange = dvec2{lo, hi}; };
        apply(net->volume.get());
        if (net->sequence)
            for (auto& v : *net->sequence) apply(v.get());
        for (auto& v : net->streamVol) apply(v.get());
        if (!net->useSequence) net->volumeSource.setData(net->volume);   // notifies the connected inports: Volume invalidation
        return (int)CPM_OK;
    });
}

int cpmh_network_set_volume_layout(cpmh_network* net, int layout) {
    return guarded([&]() {
        if (layout != CPM_VOLUME_TEXTURE && layout != CPM_VOLUME_LINEAR) throw std::invalid_argument("unknown volume layout");
        net->tracer.tracer().volumeLayout = layout;
        return (int)CPM_OK;
    });
}

int cpmh_network_stream_timestep_host(cpmh_network* net, const void* voxels) {
    return guarded([&]() {
        if (!voxels) throw std::invalid_argument("null voxel buffer");
        const cpmh_config& c = net->cfg;
        auto& rt = CpmRuntime::get();
        DataFormatId fid = c.format == CPM_FMT_U8 ? DataFormatId::UInt8 : (c.format == CPM_FMT_U16 ? DataFormatId::UInt16 : DataFormatId::Float32);
        const int prev = net->streamSlot;
        const int cur = prev < 0 ? 0 : (prev + 1) % 3;
        if (!net->streamVol[cur]) net->streamVol[cur] = std::make_shared<Volume>(size3_t(c.dims[0], c.dims[1], c.dims[2]), DataFormatBase::get(fid));
        Volume* v = net->streamVol[cur].get();
        v->setExternalRAMData(const_cast<void*>(voxels));
        const size_t r = (size_t)net->minMax.volumeRegionSize_.get();
        // multi-GPU with sharded ingest, opt-in (CPM_SHARD_GRIDS=1): every rank builds the min-max and difference bricks of ITS
        // z-slab only and the slices are all-gathered (2 x ~1 MB).  Measured at N = 8 on C4: the two extra rank
        // synchronisations cost what the 7/8 smaller passes save (min-max 0.117 vs 0.104 ms, difference 0.20 vs 0.16 ms
        // per step), so whole-volume passes on every rank stay the default.
        static const bool shardGridsEnv = std::getenv("CPM_SHARD_GRIDS") && std::atoi(std::getenv("CPM_SHARD_GRIDS")) > 0;
        const int W = rt.comm ? cpm_comm_world(rt.comm) : 1, R = rt.comm ? cpm_comm_rank(rt.comm) : 0;
        const size_t nz = (size_t)c.dims[2];
        const bool shardGrids = shardGridsEnv && c.with_importance_grid && W > 1 && rt.shardedIngest && nz % ((size_t)W * r) == 0 &&
                                (size_t)c.dims[0] % r == 0 && (size_t)c.dims[1] % r == 0 &&
                                (((size_t)c.dims[0] / r) * ((size_t)c.dims[1] / r) * (nz / W / r) * 4) % 16 == 0;
        auto slabView = [&](Volume* vol, cpm_volume** out) {
            const char* base = static_cast<const char*>(vol->deviceRead());
            const size_t slabBytes = vol->getSizeInBytes() / (size_t)W;
            const int dims[3] = {c.dims[0], c.dims[1], (int)(nz / W)};
            float scale, offset;
            vol->formatScaleOffset(scale, offset);
            rt.check(cpm_volume_create(rt.ctx(), base + slabBytes * (size_t)R, dims, c.format, scale, offset, CPM_VOLUME_LINEAR, out));
        };
        const size_t cellsPerSlab = ((size_t)c.dims[0] / r) * ((size_t)c.dims[1] / r) * (nz / (size_t)W / r);
        if (c.with_importance_grid && shardGrids) {
            if (!net->streamMinMax[cur]) {
                const size3_t outDim((size_t)c.dims[0] / r, (size_t)c.dims[1] / r, nz / r);
                auto g = std::make_shared<MinMaxUniformGrid3D>(size3_t(r));
                g->setModelMatrix(v->getModelMatrix());
                g->setWorldMatrix(v->getWorldMatrix());
                g->setDimensions(outDim);
                net->streamMinMax[cur] = g;
            }
            cpm_volume* sv = nullptr;
            slabView(v, &sv);
            uint16_t* mm = static_cast<uint16_t*>(net->streamMinMax[cur]->data.deviceWrite());
            {
                ScopedStage st("minmax");
                rt.check(cpm_volume_minmax(rt.ctx(), sv, (int)r, mm + 2 * cellsPerSlab * (size_t)R, nullptr));
                rt.check(cpm_allgather_volume(rt.comm, mm, cellsPerSlab * 4));
            }
            if (prev >= 0) {
                Volume* pv = net->streamVol[prev].get();
                if (!net->streamDiff[cur]) {
                    auto g = std::make_shared<DynamicVolumeInfoUniformGrid3D>(size3_t(r));
                    g->setModelMatrix(v->getModelMatrix());
                    g->setWorldMatrix(v->getWorldMatrix());
                    g->setDimensions(net->streamMinMax[cur]->getDimensions());
                    net->streamDiff[cur] = g;
                }
                cpm_volume* pvw = nullptr;
                slabView(pv, &pvw);
                dvec2 dataRange = pv->dataMap_.dataRange;
                const double defaultToDataRange = pv->getDataFormat()->maxValue / (dataRange.y - dataRange.x);
                float* df = static_cast<float*>(net->streamDiff[cur]->data.deviceWrite());
                {
                    ScopedStage st("voldiff");
                    rt.check(cpm_volume_diff_bricks(rt.ctx(), pvw, sv, (int)r, defaultToDataRange, dataRange.x, dataRange.y,
                                                    df + cellsPerSlab * (size_t)R));
                    rt.check(cpm_allgather_volume(rt.comm, df, cellsPerSlab * 4));
                }
                cpm_volume_destroy(rt.ctx(), pvw);
                net->diffSelector.setData(std::shared_ptr<const UniformGrid3DBase>(net->streamDiff[cur]));
                net->importance.volumeDifferenceInfoInport_.connectTo(&net->diffSelector);
            }
            cpm_volume_destroy(rt.ctx(), sv);
            net->importance.minMaxUniformGrid3DInport_.connectTo(&net->minMaxSelector);
            net->minMaxSelector.setData(std::shared_ptr<const UniformGrid3DBase>(net->streamMinMax[cur]));
        } else if (c.with_importance_grid) {
            if (!net->streamMinMax[cur]) {
                net->streamMinMax[cur] = std::shared_ptr<MinMaxUniformGrid3D>(net->minMax.compute(v).release());
            } else {
                const cpm_volume* vh = v->handle(CPM_VOLUME_LINEAR);
                ScopedStage st("minmax");
                rt.check(cpm_volume_minmax(rt.ctx(), vh, (int)r, static_cast<uint16_t*>(net->streamMinMax[cur]->data.deviceWrite()), nullptr));
            }
            if (prev >= 0) {
                // difference grid prev -> cur; a fresh grid object every step would churn the allocator, so two are kept
                net->streamDiff[cur] = DynamicVolumeDifferenceAnalysis::difference(net->streamVol[prev].get(), v, r);
                net->diffSelector.setData(std::shared_ptr<const UniformGrid3DBase>(net->streamDiff[cur]));
                net->importance.volumeDifferenceInfoInport_.connectTo(&net->diffSelector);
            }
            net->importance.minMaxUniformGrid3DInport_.connectTo(&net->minMaxSelector);
            net->minMaxSelector.setData(std::shared_ptr<const UniformGrid3DBase>(net->streamMinMax[cur]));
        }
        net->useSequence = true;
        net->streamSlot = cur;
        net->volumeSource.setData(net->streamVol[cur]);
        return (int)CPM_OK;
    });
}

int cpmh_network_prefetch_timestep_host(cpmh_network* net, const void* voxels) {
    return guarded([&]() {
        if (!voxels) throw std::invalid_argument("null voxel buffer");
        const cpmh_config& c = net->cfg;
        DataFormatId fid = c.format == CPM_FMT_U8 ? DataFormatId::UInt8 : (c.format == CPM_FMT_U16 ? DataFormatId::UInt16 : DataFormatId::Float32);
        const int nxt = net->streamSlot < 0 ? 0 : (net->streamSlot + 1) % 3;   // the slot the next stream call will use
        if (!net->streamVol[nxt]) net->streamVol[nxt] = std::make_shared<Volume>(size3_t(c.dims[0], c.dims[1], c.dims[2]), DataFormatBase::get(fid));
        net->streamVol[nxt]->prefetchExternalRAMData(const_cast<void*>(voxels));
        return (int)CPM_OK;
    });
}

int cpmh_network_sync(cpmh_network*) {
    return guarded([&]() {
        CpmRuntime::get().sync();
        return (int)CPM_OK;
    });
}

int cpmh_network_light_volume_device(cpmh_network* net, void** ptr, size_t* n_floats) {
    return guarded([&]() {
        auto v = std::const_pointer_cast<Volume>(net->toLightVolume.outport_.getData());
        if (!v || !ptr) throw std::invalid_argument("no light volume yet");
        const void* p = v->deviceRead();
        *ptr = v->deviceWrite();
        (void)p;
        if (n_floats) *n_floats = v->getSizeInBytes() / sizeof(float);
        return (int)CPM_OK;
    });
}

int cpmh_network_wait_before_light_volume_write(cpmh_network* net, void* cuda_event) {
    if (!net) return CPM_E_INVALID;
    net->toLightVolume.waitBeforeLightVolumeWrite = cuda_event;
    return CPM_OK;
}

int cpmh_network_photons_device(cpmh_network* net, void** ptr, size_t* n_floats) {
    return guarded([&]() {
        auto p = std::const_pointer_cast<PhotonData>(net->tracer.outport_.getData());
        if (!p || !ptr) throw std::invalid_argument("no photons yet");
        *ptr = const_cast<void*>(p->photons_.deviceRead());
        if (n_floats) *n_floats = p->photons_.getSizeInBytes() / sizeof(float);
        return (int)CPM_OK;
    });
}

int cpmh_network_count_collision_tests(cpmh_network* net, int on) {
    return guarded([&]() {
        auto& rt = CpmRuntime::get();
        if (on && !net->collisionCounter) {
            void* p = nullptr;
            rt.check(cpm_mem_alloc(rt.ctx(), 2 * sizeof(unsigned long long), &p));   // [0] tests, [1] tests that fetched voxels
            rt.check(cpm_mem_fill_u32(rt.ctx(), p, 0u, 4));
            net->collisionCounter = static_cast<unsigned long long*>(p);
        }
        net->tracer.tracer().collisionCounter = on ? net->collisionCounter : nullptr;
        return (int)CPM_OK;
    });
}

int cpmh_network_read_collision_stats(cpmh_network* net, unsigned long long out[2], int reset) {
    out[0] = out[1] = 0;
    return guarded([&]() {
        if (!net->collisionCounter) return (int)CPM_OK;
        auto& rt = CpmRuntime::get();
        rt.check(cpm_mem_copy_d2h(rt.ctx(), out, net->collisionCounter, 2 * sizeof(unsigned long long)));
        rt.sync();
        if (reset) rt.check(cpm_mem_fill_u32(rt.ctx(), net->collisionCounter, 0u, 4));
        return (int)CPM_OK;
    });
}
unsigned long long cpmh_network_read_collision_tests(cpmh_network* net, int reset) {
    unsigned long long v[2];
    cpmh_network_read_collision_stats(net, v, reset);
    return v[0];
}

int cpmh_network_evaluate(cpmh_network* net) {
    int ran = 0;
    int rc = guarded([&]() {
        ran += net->sampleGen.evaluate();
        for (auto& ls : net->lightSamplers) ran += ls->evaluate();
        if (net->cfg.with_importance_grid) {
            if (!net->useSequence) ran += net->minMax.evaluate();
            ran += net->importance.evaluate();
        }
        ran += net->tracer.evaluate();
        ran += net->toLightVolume.evaluate();
        return (int)CPM_OK;
    });
    return rc == CPM_OK ? ran : rc;
}

int cpmh_network_timer_event(cpmh_network* net) {
    return guarded([&]() {
        net->tracer.onTimerEvent();
        return (int)CPM_OK;
    });
}
int cpmh_network_remaining_photons(cpmh_network* net) { return net->tracer.remainingPhotonsToUpdate(); }
int cpmh_network_n_photons(cpmh_network* net) {
    auto d = net->tracer.outport_.getData();
    return d ? (int)d->getNumberOfPhotons() : 0;
}
int cpmh_network_n_recomputed(cpmh_network* net) {
    auto d = net->tracer.recomputedIndicesPort_.getData();
    return d ? d->nRecomputedPhotons : -1;
}
int cpmh_network_light_volume_dims(cpmh_network* net, int dims[3]) {
    auto v = net->toLightVolume.outport_.getData();
    if (!v) return CPM_E_INVALID;
    size3_t d = v->getDimensions();
    dims[0] = (int)d.x; dims[1] = (int)d.y; dims[2] = (int)d.z;
    return CPM_OK;
}
int cpmh_network_read_light_volume(cpmh_network* net, float* out, size_t n) {
    return guarded([&]() {
        auto v = std::const_pointer_cast<Volume>(net->toLightVolume.outport_.getData());
        if (!v || v->getSizeInBytes() != n * sizeof(float)) throw std::invalid_argument("light volume size mismatch");
        auto& rt = CpmRuntime::get();
        rt.check(cpm_mem_copy_d2h(rt.ctx(), out, v->deviceRead(), n * sizeof(float)));
        rt.sync();
        BufferBase::d2hBytes() += n * sizeof(float);
        return (int)CPM_OK;
    });
}
int cpmh_network_read_photons(cpmh_network* net, float* out, size_t n) {
    return guarded([&]() {
        auto p = std::const_pointer_cast<PhotonData>(net->tracer.outport_.getData());
        if (!p || p->photons_.getSizeInBytes() != n * sizeof(float)) throw std::invalid_argument("photon buffer size mismatch");
        auto& rt = CpmRuntime::get();
        rt.check(cpm_mem_copy_d2h(rt.ctx(), out, p->photons_.deviceRead(), n * sizeof(float)));
        rt.sync();
        BufferBase::d2hBytes() += n * sizeof(float);
        return (int)CPM_OK;
    });
}
int cpmh_network_read_importance_keys(cpmh_network* net, uint32_t* out, size_t n) {
    return guarded([&]() {
        auto& keys = net->tracer.importanceKeys();
        if (keys.getSize() != n) throw std::invalid_argument("importance key buffer size mismatch");
        auto& rt = CpmRuntime::get();
        rt.check(cpm_mem_copy_d2h(rt.ctx(), out, keys.deviceRead(), n * sizeof(uint32_t)));
        rt.sync();
        return (int)CPM_OK;
    });
}
int cpmh_network_read_recomputed_indices(cpmh_network* net, uint32_t* out, size_t n) {
    return guarded([&]() {
        auto d = std::const_pointer_cast<RecomputedPhotonIndices>(net->tracer.recomputedIndicesPort_.getData());
        if (!d || d->nRecomputedPhotons <= 0) return 0;
        const size_t m = (size_t)d->nRecomputedPhotons;
        if (n < m) throw std::invalid_argument("index buffer too small");
        auto& rt = CpmRuntime::get();
        rt.check(cpm_mem_copy_d2h(rt.ctx(), out, d->indicesToRecomputedPhotons.deviceRead(), m * sizeof(uint32_t)));
        rt.sync();
        return (int)m;
    });
}
int cpmh_network_read_importance_grid(cpmh_network* net, float* out, size_t n) {
    return guarded([&]() {
        auto g = std::dynamic_pointer_cast<const ImportanceUniformGrid3D>(net->importance.importanceUniformGrid3DOutport_.getData());
        if (!g || g->data.getSize() != n) throw std::invalid_argument("importance grid size mismatch");
        auto& rt = CpmRuntime::get();
        rt.check(cpm_mem_copy_d2h(rt.ctx(), out, g->data.deviceRead(), n * sizeof(float)));
        rt.sync();
        return (int)CPM_OK;
    });
}
int cpmh_network_light_setup(cpmh_network* net, int light, float out[19]) {
    return guarded([&]() {
        if (light < 0 || light >= (int)net->lightSamplers.size()) throw std::invalid_argument("no such light");
        const auto& s = net->lightSamplers[light]->sampler().lastSetup;
        const vec3 v[6] = {s.direction, s.planePoint, s.origin, s.u, s.v, s.radiance};
        for (int k = 0; k < 6; ++k) { out[3 * k] = v[k].x; out[3 * k + 1] = v[k].y; out[3 * k + 2] = v[k].z; }
        out[18] = s.area;
        return (int)CPM_OK;
    });
}
int cpmh_network_read_light_samples(cpmh_network* net, int light, float* samples_out, float* isect_out, size_t n) {
    return guarded([&]() {
        if (light < 0 || light >= (int)net->lightSamplers.size()) throw std::invalid_argument("no such light");
        auto* ls = const_cast<LightSamples*>(net->lightSamplers[light]->lightSamples());
        if (!ls || ls->getSize() != n) throw std::invalid_argument("light sample count mismatch");
        auto& rt = CpmRuntime::get();
        if (samples_out) rt.check(cpm_mem_copy_d2h(rt.ctx(), samples_out, ls->getLightSamples()->deviceRead(), n * 8 * sizeof(float)));
        if (isect_out) rt.check(cpm_mem_copy_d2h(rt.ctx(), isect_out, ls->getIntersectionPoints()->deviceRead(), n * 2 * sizeof(float)));
        rt.sync();
        return (int)CPM_OK;
    });
}
const char* cpmh_network_last_splat_path(cpmh_network* net) { return net->toLightVolume.lastPath.c_str(); }
int cpmh_network_set_profile(cpmh_network*, int on) {
    StageProfiler::get().enabled = on != 0;
    return CPM_OK;
}
float cpmh_network_stage_ms(cpmh_network*, const char* stage) {
    float v = 0.f;
    guarded([&]() { v = (float)StageProfiler::get().lastMs(stage); return (int)CPM_OK; });
    return v;
}
void cpmh_profile_enable(int on) { StageProfiler::get().enabled = on != 0; }
void cpmh_profile_only(const char* stage) { StageProfiler::get().only = stage ? stage : ""; }
void cpmh_profile_reset(void) { guarded([&]() { StageProfiler::get().reset(); return (int)CPM_OK; }); }
double cpmh_profile_total_ms(const char* stage) {
    double v = 0;
    guarded([&]() { v = StageProfiler::get().totalMs(stage); return (int)CPM_OK; });
    return v;
}
int cpmh_profile_count(const char* stage) {
    int v = 0;
    guarded([&]() { v = StageProfiler::get().count(stage); return (int)CPM_OK; });
    return v;
}
const char* cpmh_profile_stages(void) {
    static std::string s;
    guarded([&]() { s = StageProfiler::get().stages(); return (int)CPM_OK; });
    return s.c_str();
}
uint64_t cpmh_network_launch_count(cpmh_network*, int reset) { return cpm_ctx_launch_count(CpmRuntime::get().ctx(), reset); }
void cpmh_transfer_bytes(uint64_t* h2d, uint64_t* d2h, int reset) {
    if (h2d) *h2d = BufferBase::h2dBytes();
    if (d2h) *d2h = BufferBase::d2hBytes();
    if (reset) BufferBase::h2dBytes() = BufferBase::d2hBytes() = 0;
}
void* cpmh_network_ctx(cpmh_network*) { return CpmRuntime::get().ctx(); }

int cpmh_u3d_write(const char* path, int format, const int dims4[4], const int cell[3], const float model[16],
                   const float world[16], const void* data) {
    return guarded([&]() {
        if (!path || !dims4 || !cell || !data) throw std::invalid_argument("null argument");
        if (format != 0 && format != 1) throw std::invalid_argument("format must be 0 (FLOAT32) or 1 (Vec2UINT16)");
        UniformGrid3DVector v;
        const size3_t dim(dims4[0], dims4[1], dims4[2]), cd(cell[0], cell[1], cell[2]);
        const size_t elem = 4, bytes = (size_t)dims4[0] * dims4[1] * dims4[2] * elem;
        for (int t = 0; t < dims4[3]; ++t) {
            std::shared_ptr<UniformGrid3DBase> g;
            if (format == 0) g = std::make_shared<UniformGrid3D<float>>(dim, cd);
            else g = std::make_shared<UniformGrid3D<u16vec2>>(dim, cd);
            mat4 m, w;
            for (int k = 0; k < 16; ++k) {
                if (model) m[k / 4][k % 4] = model[k];
                if (world) w[k / 4][k % 4] = world[k];
            }
            g->setModelMatrix(m);
            g->setWorldMatrix(w);
            std::memcpy(g->getData(), static_cast<const char*>(data) + (size_t)t * bytes, bytes);
            v.push_back(g);
        }
        UniformGrid3DWriter().writeData(&v, path);
        return (int)CPM_OK;
    });
}
static thread_local std::shared_ptr<UniformGrid3DVector> g_u3d;
static thread_local std::string g_u3dPath;
static std::shared_ptr<UniformGrid3DVector> u3dLoad(const char* path) {
    if (!g_u3d || g_u3dPath != path) {
        g_u3d = UniformGrid3DReader().readData(path);
        g_u3dPath = path;
    }
    return g_u3d;
}
int cpmh_u3d_read_info(const char* path, int* format, int dims4[4], int cell[3], float model[16], float world[16]) {
    return guarded([&]() {
        if (!path) throw std::invalid_argument("null argument");
        g_u3d.reset();
        auto v = u3dLoad(path);
        UniformGrid3DBase* g = v->front().get();
        if (format) *format = std::string(g->getFormatString()) == "FLOAT32" ? 0 : 1;
        size3_t d = g->getDimensions(), c = g->getCellDimension();
        if (dims4) { dims4[0] = (int)d.x; dims4[1] = (int)d.y; dims4[2] = (int)d.z; dims4[3] = (int)v->size(); }
        if (cell) { cell[0] = (int)c.x; cell[1] = (int)c.y; cell[2] = (int)c.z; }
        mat4 m = g->getModelMatrix(), w = g->getWorldMatrix();
        for (int k = 0; k < 16; ++k) {
            if (model) model[k] = m[k / 4][k % 4];
            if (world) world[k] = w[k / 4][k % 4];
        }
        return (int)CPM_OK;
    });
}
int cpmh_u3d_read_data(const char* path, void* out, size_t bytes) {
    return guarded([&]() {
        if (!path || !out) throw std::invalid_argument("null argument");
        auto v = u3dLoad(path);
        size_t each = v->front()->getSizeInBytes();
        if (bytes != each * v->size()) throw std::invalid_argument("output buffer size does not match the file");
        for (size_t t = 0; t < v->size(); ++t) std::memcpy(static_cast<char*>(out) + t * each, (*v)[t]->getData(), each);
        g_u3d.reset();
        return (int)CPM_OK;
    });
}
int cpmh_network_export_sequence_grids(cpmh_network* net, int which, const char* path) {
    return guarded([&]() {
        if (!net || !path) throw std::invalid_argument("null argument");
        const auto& v = which == 0 ? net->seqMinMax : net->seqDiff;
        if (v.empty()) throw std::invalid_argument("no resident sequence (cpmh_network_set_sequence_host first)");
        UniformGrid3DVector copy(v.begin(), v.end());
        UniformGrid3DWriter().writeData(&copy, path);
        return (int)CPM_OK;
    });
}

static std::vector<std::pair<std::string, std::vector<Processor*>>> network_processors(cpmh_network* n) {
    std::vector<Processor*> samplers;
    for (auto& s : n->lightSamplers) samplers.push_back(s.get());
    return {{"org.inviwo.UniformSampleGenerator2DCL", {&n->sampleGen}},
            {"org.inviwo.DirectionalLightSamplerCL", samplers},
            {"org.inviwo.VolumeMinMaxCLProcessor", {&n->minMax}},
            {"org.inviwo.MinMaxUniformGrid3DImportanceCLProcessor", {&n->importance}},
            {"org.inviwo.ProgressivePhotonTracerCL", {&n->tracer}},
            {"org.inviwo.PhotonToLightVolumeProcessorCL", {&n->toLightVolume}}};
}

const char* cpmh_workspace_describe(const char* path) {
    static thread_local std::string text;
    int rc = guarded([&]() {
        if (!path) throw std::invalid_argument("null argument");
        text = Workspace::load(path).describe();
        return (int)CPM_OK;
    });
    return rc == CPM_OK ? text.c_str() : nullptr;
}

int cpmh_config_from_workspace(const char* path, const float basis[3], cpmh_config* cfg) {
    return guarded([&]() {
        if (!path || !cfg) throw std::invalid_argument("null argument");
        Workspace w = Workspace::load(path);
        double s = 0, v[4];
        int nv = 0;
        std::string sel;
        auto gens = w.ofType("org.inviwo.UniformSampleGenerator2DCL");
        if (!gens.empty() && readVec(gens[0]->property("nSamples"), v, nv) && nv >= 2) {
            if (v[0] != v[1]) throw std::invalid_argument("workspace: nSamples is not square; the headless network takes one side length");
            cfg->samples_per_side = (int)v[0];
        }
        auto samplers = w.ofType("org.inviwo.DirectionalLightSamplerCL");
        if (samplers.size() > 8) throw std::invalid_argument("workspace: more than 8 light samplers");
        if (!samplers.empty()) cfg->n_lights = (int)samplers.size();
        auto sources = w.ofType("org.inviwo.Directionallightsource");
        for (size_t l = 0; l < sources.size() && l < 8; ++l) {
            const XmlNode* pos = sources[l]->property("lightPosition");
            double p[3] = {0, 0, 1};
            bool have = false;
            if (const XmlNode* ws = pos ? pos->child("positionWorldSpace") : nullptr) {
                p[0] = std::atof(ws->attrOr("x", "0").c_str());
                p[1] = std::atof(ws->attrOr("y", "0").c_str());
                p[2] = std::atof(ws->attrOr("z", "0").c_str());
                have = true;
            } else if (readVec(sources[l]->property("lightPosition.position"), v, nv) && nv >= 3) {
                p[0] = v[0]; p[1] = v[1]; p[2] = v[2];
                have = true;
            }
            if (have) {
                // world -> texture space for a direction: divide by the volume's world extents, renormalise
                double d[3], len = 0;
                for (int k = 0; k < 3; ++k) {
                    d[k] = -p[k] / (basis ? (double)basis[k] : 1.0);
                    len += d[k] * d[k];
                }
                len = std::sqrt(len);
                if (!(len > 0)) throw std::invalid_argument("workspace: light source at the origin has no direction");
                for (int k = 0; k < 3; ++k) cfg->light_directions[l][k] = (float)(d[k] / len);
            }
            double power = 1.0, diffuse[3] = {1, 1, 1};
            readScalar(sources[l]->property("lighting.lightPower"), power);
            if (readVec(sources[l]->property("lighting.lightDiffuse"), v, nv) && nv >= 3)
                for (int k = 0; k < 3; ++k) diffuse[k] = v[k];
            for (int k = 0; k < 3; ++k) cfg->light_intensity[l][k] = (float)(power * diffuse[k]);
        }
        auto tracers = w.ofType("org.inviwo.ProgressivePhotonTracerCL");
        if (!tracers.empty()) {
            const WorkspaceProcessor* t = tracers[0];
            if (readScalar(t->property("maxScatteringEvents"), s)) cfg->max_scattering_events = (int)s;
            if (readScalar(t->property("radius"), s)) cfg->photon_radius_voxels = (float)s;
            if (readScalar(t->property("maxIncrementalPhotonsToUpdate"), s)) cfg->max_incremental_percent = (float)s;
            const char* clips[3] = {"clipX", "clipY", "clipZ"};
            for (int k = 0; k < 3; ++k)
                if (readVec(t->property(clips[k]), v, nv) && nv >= 2) {
                    cfg->clip[2 * k] = (int)v[0];
                    cfg->clip[2 * k + 1] = (int)v[1];
                }
            cfg->with_importance_grid = w.connected("org.inviwo.MinMaxUniformGrid3DImportanceCLProcessor",
                                                    "org.inviwo.ProgressivePhotonTracerCL", "recomputationImportance") ? 1 : 0;
        }
        auto lvs = w.ofType("org.inviwo.PhotonToLightVolumeProcessorCL");
        if (!lvs.empty()) {
            if (readSelected(lvs[0]->property("volumeSizeOption"), sel))
                cfg->light_volume_option = sel == "1" ? 1 : (sel == "1/2" ? 2 : (sel == "1/4" ? 4 : 0));
            if (readSelected(lvs[0]->property("volumeDataType"), sel)) cfg->light_volume_channels = sel == "4xfloat32" ? 4 : 1;
            if (readScalar(lvs[0]->property("incrementalRecomputationThreshold"), s)) cfg->incremental_threshold_percent = (float)s;
        }
        return (int)CPM_OK;
    });
}

int cpmh_network_load_workspace(cpmh_network* net, const char* path) {
    return guarded([&]() {
        if (!net || !path) throw std::invalid_argument("null argument");
        Workspace w = Workspace::load(path);
        int applied = 0;
        for (auto& entry : network_processors(net)) {
            auto found = w.ofType(entry.first);
            if (found.size() > entry.second.size())
                throw std::invalid_argument("workspace has " + std::to_string(found.size()) + " x " + entry.first +
                                            ", the network " + std::to_string(entry.second.size()));
            for (size_t k = 0; k < found.size(); ++k) applied += applyWorkspaceProperties(*entry.second[k], *found[k]);
        }
        return applied;
    });
}

int cpmh_network_set_samples_per_side(cpmh_network* net, int n) {
    return guarded([&]() {
        if (!net) throw std::invalid_argument("null argument");
        if (n < net->sampleGen.nSamples_.getMinValue().x || n > net->sampleGen.nSamples_.getMaxValue().x)
            throw std::invalid_argument("nSamples out of range");
        net->sampleGen.nSamples_.set(ivec2{n, n});
        net->cfg.samples_per_side = n;
        return (int)CPM_OK;
    });
}

int cpmh_network_get_property(cpmh_network* net, const char* class_id, int k, const char* property, double* out) {
    return guarded([&]() {
        if (!net || !class_id || !property || !out) throw std::invalid_argument("null argument");
        for (auto& entry : network_processors(net)) {
            if (entry.first != class_id) continue;
            if (k < 0 || k >= (int)entry.second.size()) throw std::invalid_argument("no such processor instance");
            Property* p = entry.second[k]->getPropertyByIdentifier(property);
            if (!p) throw std::invalid_argument(std::string("no property ") + property);
            if (auto* f = dynamic_cast<FloatProperty*>(p)) *out = f->get();
            else if (auto* i = dynamic_cast<IntProperty*>(p)) *out = i->get();
            else if (auto* b = dynamic_cast<BoolProperty*>(p)) *out = b->get() ? 1.0 : 0.0;
            else if (auto* o = dynamic_cast<OptionProperty<int>*>(p)) *out = o->get();
            else if (auto* t = dynamic_cast<TransferFunctionProperty*>(p)) *out = (double)t->get().size();
            else throw std::invalid_argument(std::string("property ") + property + " is not scalar");
            return (int)CPM_OK;
        }
        throw std::invalid_argument(std::string("no processor of class ") + class_id);
    });
}

int cpmh_network_set_property(cpmh_network* net, const char* class_id, int k, const char* property, double value) {
    return guarded([&]() {
        if (!net || !class_id || !property) throw std::invalid_argument("null argument");
        for (auto& entry : network_processors(net)) {
            if (entry.first != class_id) continue;
            if (k < 0 || k >= (int)entry.second.size()) throw std::invalid_argument("no such processor instance");
            Property* p = entry.second[k]->getPropertyByIdentifier(property);
            if (!p) throw std::invalid_argument(std::string("no property ") + property);
            if (auto* f = dynamic_cast<FloatProperty*>(p)) f->set((float)value);
            else if (auto* i = dynamic_cast<IntProperty*>(p)) i->set((int)value);
            else if (auto* b = dynamic_cast<BoolProperty*>(p)) b->set(value != 0.0);
            else if (auto* o = dynamic_cast<OptionProperty<int>*>(p)) o->setSelectedValue((int)value);
            else throw std::invalid_argument(std::string("property ") + property + " is not scalar");
            return (int)CPM_OK;
        }
        throw std::invalid_argument(std::string("no processor of class ") + class_id);
    });
}

int cpmh_network_importance_tf_points(cpmh_network* net, float* positions, float* colors, int capacity) {
    int n = 0;
    int rc = guarded([&]() {
        n = net->importance.tfPointImportanceSize();
        if (n > capacity) throw std::invalid_argument("capacity too small");
        const auto& P = net->importance.tfPointPositions();
        const auto& C = net->importance.tfPointColors();
        for (int i = 0; i < n; ++i) {
            positions[i] = P[i];
            for (int k = 0; k < 4; ++k) colors[4 * i + k] = C[i][k];
        }
        return (int)CPM_OK;
    });
    return rc == CPM_OK ? n : rc;
}

int cpmh_random_numbers(int nx, int ny, int seed, int evaluations, float* out_host) {
    return guarded([&]() {
        if (!out_host || nx < 1 || ny < 0 || evaluations < 1) throw std::invalid_argument("bad argument");
        CpmRuntime::init(0);
        if (ny == 0) {
            RandomNumberGeneratorCL p;
            p.seed_.set(seed);
            p.nRandomNumbers_.set(nx);
            for (int k = 0; k < evaluations; ++k) p.process();
            auto data = p.randomNumbersPort_.getData();
            std::memcpy(out_host, const_cast<Buffer<float>*>(data.get())->getRAMRepresentation()->data(), (size_t)nx * sizeof(float));
        } else {
            RandomNumberGenerator2DCL p;
            p.seed_.set(seed);                       // before nSamples: the streams are seeded when nSamples changes
            p.nRandomNumbers_.set(ivec2{nx, ny});
            for (int k = 0; k < evaluations; ++k) p.process();
            auto img = p.randomNumbersPort_.getData();
            std::memcpy(out_host, const_cast<ImageF32*>(img.get())->data.getRAMRepresentation()->data(), (size_t)nx * ny * sizeof(float));
        }
        return (int)CPM_OK;
    });
}

int cpmh_tf_difference_points(const float* cur, int n_cur, const float* prev, int n_prev, float epsilon, int associated,
                              float* positions, float* colors, int capacity) {
    int n = 0;
    int rc = guarded([&]() {
        TransferFunction a, b;
        for (int i = 0; i < n_cur; ++i) a.add(cur[5 * i], vec4(cur[5 * i + 1], cur[5 * i + 2], cur[5 * i + 3], cur[5 * i + 4]));
        for (int i = 0; i < n_prev; ++i) b.add(prev[5 * i], vec4(prev[5 * i + 1], prev[5 * i + 2], prev[5 * i + 3], prev[5 * i + 4]));
        MinMaxUniformGrid3DImportanceCLProcessor p;
        p.TFPointEpsilon_.set(epsilon);
        p.useAssociatedColor_.set(associated != 0);
        p.buildDifferenceLists(a, b);
        n = p.tfPointImportanceSize();
        if (n > capacity) throw std::invalid_argument("capacity too small");
        const auto& P = p.tfPointPositions();
        const auto& C = p.tfPointColors();
        for (int i = 0; i < n; ++i) {
            positions[i] = P[i];
            for (int k = 0; k < 4; ++k) colors[4 * i + k] = C[i][k];
        }
        return (int)CPM_OK;
    });
    return rc == CPM_OK ? n : rc;
}

int cpmh_photondata_progress(size_t n_photons, int max_interactions, double radius_rel, double scene_radius, int iterations,
                             double alpha, double out[4]) {
    return guarded([&]() {
        // (sizes only: PhotonData::setSize would allocate the record buffer; the arithmetic needs the count alone)
        struct Sized : PhotonData {
            void sizeOnly(size_t n, int I) { maxPhotonInteractions_ = I; count_ = n; }
            size_t count_ = 0;
        } d;
        d.sizeOnly(n_photons, max_interactions);
        d.setRadius(radius_rel, scene_radius);
        for (int i = 0; i < iterations; ++i) d.advanceToNextIteration(alpha);
        out[0] = d.getRadius();
        out[1] = d.getRadiusRelativeToSceneSize();
        // getRelativeIrradianceScale with the photon count given (ppm/photondata.cpp:84-94)
        out[2] = PhotonData::sphereVolume(d.getRadiusRelativeToSceneSize()) / PhotonData::sphereVolume(PhotonData::defaultRadiusRelativeToSceneRadius) *
                 ((double)n_photons / (double)PhotonData::defaultNumberOfPhotons);
        out[3] = d.iteration();
        return (int)CPM_OK;
    });
}
int cpmh_photon_encode_direction(const float dir[3], float out[2]) {
    Photon p;
    p.setDirection(vec3(dir[0], dir[1], dir[2]));
    out[0] = p.encodedDirection.x;
    out[1] = p.encodedDirection.y;
    return CPM_OK;
}
int cpmh_photon_decode_direction(const float enc[2], float out[3]) {
    Photon p;
    p.encodedDirection = vec2{enc[0], enc[1]};
    vec3 d = p.getDirection();
    out[0] = d.x; out[1] = d.y; out[2] = d.z;
    return CPM_OK;
}
int cpmh_network_photon_state(cpmh_network* net, double out[5]) {
    return guarded([&]() {
        auto d = net->tracer.outport_.getData();
        if (!d) throw std::invalid_argument("no photons yet");
        out[0] = d->iteration();
        out[1] = d->getRadius();
        out[2] = d->getSceneRadius();
        out[3] = d->getRadiusRelativeToSceneSize();
        out[4] = d->getRelativeIrradianceScale();
        return (int)CPM_OK;
    });
}

int cpmh_fit_light_plane(const float* points, int n, const float P[3], const float N[3], float out[9]) {
    return guarded([&]() {
        std::vector<vec3> pts;
        for (int i = 0; i < n; ++i) pts.push_back(vec3(points[3 * i], points[3 * i + 1], points[3 * i + 2]));
        auto f = geometry::fitPlaneAlignedOrientedBoundingBox2D(pts, vec3(P[0], P[1], P[2]), vec3(N[0], N[1], N[2]));
        const vec3 v[3] = {f.origin, f.u, f.v};
        for (int k = 0; k < 3; ++k) { out[3 * k] = v[k].x; out[3 * k + 1] = v[k].y; out[3 * k + 2] = v[k].z; }
        return (int)CPM_OK;
    });
}

int cpmh_convex_hull2d(const float* pts, int n, float* hull_out) {
    int m = 0;
    int rc = guarded([&]() {
        std::vector<vec2> p;
        for (int i = 0; i < n; ++i) p.push_back(vec2{pts[2 * i], pts[2 * i + 1]});
        auto h = geometry::convexHull2D(p);
        for (size_t i = 0; i < h.size(); ++i) { hull_out[2 * i] = h[i].x; hull_out[2 * i + 1] = h[i].y; }
        m = (int)h.size();
        return (int)CPM_OK;
    });
    return rc == CPM_OK ? m : rc;
}

int cpmh_player_clock(int n_elements, float time_per_element, int frame_rate, int ticks, float* times_out, int* index_out) {
    return guarded([&]() {
        if (n_elements < 1 || !times_out || !index_out) throw std::invalid_argument("bad argument");
        SequenceClock c("timePerElement", "Time Per element (s)", "frameRate");
        c.timePerElement_.set(time_per_element);
        c.frameRate_.set(frame_rate);
        c.onSequenceChange((size_t)n_elements);
        for (int k = 0; k < ticks; ++k) {
            c.onSequenceTimerEvent();
            times_out[k] = c.time_.get();
            index_out[k] = c.index_.get();
        }
        return (int)CPM_OK;
    });
}

int cpmh_player_grids_f32(const float* grids, int n_grids, size_t n_cells, float time_per_element, const float* times, int n_times,
                          float* out_host, int* out_index, int* out_buffer) {
    return guarded([&]() {
        if (!grids || n_grids < 1 || !times || !out_host) throw std::invalid_argument("bad argument");
        auto seq = std::make_shared<UniformGrid3DVector>();
        for (int g = 0; g < n_grids; ++g) {
            auto grid = std::make_shared<UniformGrid3D<float>>(size3_t(n_cells, 1, 1), size3_t(8));
            std::memcpy(grid->getData(), grids + (size_t)g * n_cells, n_cells * sizeof(float));
            seq->push_back(grid);
        }
        DataOutport<UniformGrid3DVector> src("Sequence");
        UniformGrid3DPlayerProcessor player;
        player.inport_.connectTo(&src);
        src.setData(std::shared_ptr<const UniformGrid3DVector>(seq));
        player.clock_.timePerElement_.set(time_per_element);
        std::vector<const UniformGrid3DBase*> seen;
        for (int k = 0; k < n_times; ++k) {
            player.clock_.time_.set(times[k]);
            player.process();
            auto out = player.outport_.getData();
            auto* g = dynamic_cast<const UniformGrid3D<float>*>(out.get());
            if (!g) throw std::invalid_argument("player produced no float grid");
            const std::vector<float>* ram = const_cast<UniformGrid3D<float>*>(g)->data.getRAMRepresentation();
            std::memcpy(out_host + (size_t)k * n_cells, ram->data(), n_cells * sizeof(float));
            if (out_index) out_index[k] = player.clock_.index_.get();
            if (out_buffer) {
                int id = -1;
                bool isInput = false;
                for (auto& e : *seq) isInput = isInput || e.get() == out.get();
                if (!isInput) {
                    auto it = std::find(seen.begin(), seen.end(), out.get());
                    if (it == seen.end()) { seen.push_back(out.get()); it = seen.end() - 1; }
                    id = (int)(it - seen.begin());
                }
                out_buffer[k] = id;
            }
        }
        return (int)CPM_OK;
    });
}

int cpmh_player_volumes(const void* volumes, int n_volumes, const int dims[3], int format, float time_per_volume, const float* times,
                        int n_times, void* out_host, int* out_index) {
    return guarded([&]() {
        if (!volumes || n_volumes < 1 || !dims || !times || !out_host) throw std::invalid_argument("bad argument");
        DataFormatId fid = format == CPM_FMT_U8 ? DataFormatId::UInt8 : (format == CPM_FMT_U16 ? DataFormatId::UInt16 : DataFormatId::Float32);
        auto seq = std::make_shared<VolumeSequence>();
        const size3_t d(dims[0], dims[1], dims[2]);
        size_t bytes = 0;
        for (int v = 0; v < n_volumes; ++v) {
            auto vol = std::make_shared<Volume>(d, DataFormatBase::get(fid));
            bytes = vol->getSizeInBytes();
            std::memcpy(vol->getEditableRAMData(), static_cast<const char*>(volumes) + (size_t)v * bytes, bytes);
            seq->push_back(vol);
        }
        DataOutport<VolumeSequence> src("volumeSequence");
        VolumeSequencePlayer player;
        player.inport_.connectTo(&src);
        src.setData(std::shared_ptr<const VolumeSequence>(seq));
        player.clock_.timePerElement_.set(time_per_volume);
        for (int k = 0; k < n_times; ++k) {
            player.clock_.time_.set(times[k]);
            player.process();
            auto out = std::const_pointer_cast<Volume>(player.outport_.getData());
            std::memcpy(static_cast<char*>(out_host) + (size_t)k * bytes, out->getRAMData(), bytes);
            if (out_index) out_index[k] = player.clock_.index_.get();
        }
        return (int)CPM_OK;
    });
}

const char* cpmh_describe_processors(void) {
    static std::string s;
    std::ostringstream os;
    auto dump = [&](Processor& p) {
        os << p.getProcessorInfo().classIdentifier << "|";
        bool first = true;
        for (auto& id : p.getPortIdentifiers()) { os << (first ? "" : ",") << id; first = false; }
        os << "|";
        first = true;
        for (auto& id : p.getPropertyIdentifiers()) { os << (first ? "" : ",") << id; first = false; }
        os << "\n";
    };
    { UniformSampleGenerator2DProcessorCL p; dump(p); }
    { DirectionalLightSamplerCLProcessor p; dump(p); }
    { ProgressivePhotonTracerCL p; dump(p); }
    { PhotonToLightVolumeProcessorCL p; dump(p); }
    { VolumeMinMaxCLProcessor p; dump(p); }
    { DynamicVolumeDifferenceAnalysis p; dump(p); }
    { MinMaxUniformGrid3DImportanceCLProcessor p; dump(p); }
    { RadixSortCL p; dump(p); }
    { RandomNumberGeneratorCL p; dump(p); }
    { RandomNumberGenerator2DCL p; dump(p); }
    { UniformGrid3DPlayerProcessor p; dump(p); }
    { VolumeSequencePlayer p; dump(p); }
    s = os.str();
    return s.c_str();
}

}  // extern "C"
