"""Drop-in host layer: processor identifiers / ports / properties of the reference, and the network
evaluated headless (full trace, correlated re-trace, incremental light-volume update)."""
import importlib

import numpy as np
import pytest

import scenes
from conftest import PKG_NAME
from test_tracer import oracle_trace

FLT_MAX = np.float32(3.4028234663852886e38)

# class id -> (ports, properties that the reference adds)
CONTRACT = {
    "org.inviwo.ProgressivePhotonTracerCL": (
        {"volume", "recomputationImportance", "LightSamples", "photons", "recomputedIndices"},
        {"samplingRate", "radius", "maxScatteringEvents", "noSingleScattering", "alpha", "material", "transferFunction",
         "wgsize", "glsharing", "camera", "maxIncrementalPhotonsToUpdate", "equalImportance", "spatialSorting",
         "invalidate", "enableRefinement", "enableProgressiveRecomputation", "clipX", "clipY", "clipZ"}),
    "org.inviwo.PhotonToLightVolumeProcessorCL": (
        {"volume", "photons", "recomputedPhotonIndices", "lightvolume"},
        {"incrementalRecomputationThreshold", "volumeSizeOption", "volumeDataType", "alignChangedPhotons", "wgsize",
         "glsharing"}),
    "org.inviwo.DirectionalLightSamplerCL": ({"SceneGeometry", "samples", "light", "LightSamples"}, {"wgsize"}),
    "org.inviwo.UniformSampleGenerator2DCL": ({"samples", "DirectionalSamples"}, {"nSamples", "wgsize", "glsharing"}),
    "org.inviwo.MinMaxUniformGrid3DImportanceCLProcessor": (
        {"minMaxUniformGrid3D", "volumeDifferenceInfo", "importanceUniformGrid3D"},
        {"incrementalImportance", "constantWeight", "opacityDiffWeight", "colorWeight", "colorDiffWeight",
         "useAssociatedColor", "TFPointEpsilon", "transferfunction", "wgsize", "glsharing"}),
    "org.inviwo.VolumeMinMaxCLProcessor": ({"volume", "VolumeSequenceInput", "output", "UniformGrid3DVectorOut"},
                                           {"region", "wgsize", "glsharing"}),
    "org.inviwo.DynamicVolumeDifferenceAnalysis": ({"data", "DynamicDataInfo"}, {"region"}),
    "org.inviwo.RadixSortCL": ({"unsortedKeys", "unsortedData", "sortedData"}, set()),
    "org.inviwo.RandomNumberGeneratorCL": ({"samples"}, {"nSamples", "genRnd", "seed", "wgsize", "glsharing"}),
    "org.inviwo.RandomNumberGenerator2DCL": ({"samples"}, {"nSamples", "genRnd", "seed", "wgsize", "glsharing"}),
}


@pytest.fixture(scope="module")
def host():
    return importlib.import_module(PKG_NAME + ".host")


def test_processor_contract(host):
    got = host.describe_processors()
    for cid, (ports, props) in CONTRACT.items():
        assert cid in got, cid
        assert got[cid][0] == ports, (cid, got[cid][0] ^ ports)
        assert props <= got[cid][1], (cid, props - got[cid][1])


@pytest.mark.gpu
def test_network_full_trace_matches_oracle(host, cpm, orc, synth, torch_cuda):
    dims, ns, I = (64, 64, 64), 96, 2
    d = (0.3, -0.5, 0.8)
    vol = synth.volume_u8(dims, 8)
    net = host.Network(dims, cpm.CPM_FMT_U8, ns, [d], max_scattering_events=I, light_volume_option=2)
    net.set_transfer_function(synth.WS_TF_POINTS)
    net.set_volume_host(vol)
    assert net.evaluate() >= 4
    ph = net.read_photons(I)
    # the oracle with the same set-up: light samples from the kernel arguments the host layer derived (its CPU plane
    # fit equals the oracle's and the reference's own files bit for bit), the transfer function rasterised as the host does
    from oracle import frame
    from test_configs import oracle_lights_from_host
    L = dict(oracle_lights_from_host(orc, synth, net, ns)[0], n=ns * ns)
    tf = frame.rasterise_tf(synth.WS_TF_POINTS)
    want, _, _ = oracle_trace(orc, vol, tf, L, max_interactions=I, step_size=1.0 / 64)
    assert (want[:, 0] != FLT_MAX).sum() > 1000
    assert np.array_equal(ph.view(np.uint32), want.view(np.uint32))     # through the plug-in: every photon bit for bit
    assert net.last_splat_path == "full"
    lv = net.read_light_volume()
    assert lv.shape[0] == 32 * 32 * 32 and lv.sum() > 0 and np.isfinite(lv).all()
    assert net.evaluate() == 0          # nothing invalid: nothing runs
    net.close()


@pytest.mark.gpu
def test_network_correlated_retrace(host, cpm, orc, synth, torch_cuda):
    """TF change with the importance grid connected: only photons whose paths cross changed cells are
    re-traced, the light volume is updated by -old/+new splats and equals a from-scratch result."""
    dims, ns, I = (64, 64, 64), 128, 2
    d = (0.2, 0.3, 0.9)
    vol = synth.volume_u8(dims, 8)
    kw = dict(max_scattering_events=I, light_volume_option=2, with_importance_grid=True, reference_full_splat_bound=False,
              incremental_threshold=100.0)
    net = host.Network(dims, cpm.CPM_FMT_U8, ns, [d], **kw)
    net.set_transfer_function(synth.WS_TF_POINTS)
    net.set_volume_host(vol)
    net.evaluate()
    assert net.n_recomputed == -1 and net.last_splat_path == "full"
    before = net.read_photons(I).copy()
    # change only the top of the transfer function (dense material): few cells are affected
    pts = list(synth.WS_TF_POINTS)
    pts[-1] = (pts[-1][0], (0.1, 0.6, 0.65, 0.9))
    net.set_transfer_function(pts)
    net.evaluate()
    n = net.n_photons
    nrec = net.n_recomputed
    assert 0 < nrec < n, nrec
    assert net.last_splat_path == "incremental"
    after = net.read_photons(I)
    changed = np.any(before.view(np.uint32) != after.view(np.uint32), axis=1).reshape(I, n).any(axis=0)
    assert changed.sum() <= nrec
    lv_inc = net.read_light_volume().astype(np.float64)
    # from-scratch network with the new transfer function
    ref = host.Network(dims, cpm.CPM_FMT_U8, ns, [d], **kw)
    ref.set_transfer_function(pts)
    ref.set_volume_host(vol)
    ref.evaluate()
    full = ref.read_photons(I)
    # replay property: every photon equals the from-scratch trace (re-traced ones are new, the others
    # were provably unaffected) unless the detector's conservative grid missed nothing
    same = np.all(full.view(np.uint32) == after.view(np.uint32), axis=1)
    assert same.mean() > 0.999, same.mean()
    lv_full = ref.read_light_volume().astype(np.float64)
    rmse = np.sqrt(((lv_inc - lv_full) ** 2).mean()) / np.sqrt((lv_full ** 2).mean())
    assert rmse < 2e-3, rmse
    net.close(); ref.close()


@pytest.mark.gpu
def test_network_consecutive_incremental_updates_do_not_drift(host, cpm, synth, torch_cuda):
    """Four transfer-function changes in a row, each an incremental -old/+new update: the copy of the previous records
    that the update subtracts is kept current by the update itself (cpm_splat_photons_update_sync), so after the last
    change the light volume still equals a from-scratch network's (a stale copy would subtract the wrong records)."""
    dims, ns, I = (64, 64, 64), 96, 2
    d = (0.2, 0.3, 0.9)
    vol = synth.volume_u8(dims, 8)
    kw = dict(max_scattering_events=I, light_volume_option=2, with_importance_grid=True, reference_full_splat_bound=False,
              incremental_threshold=100.0)
    net = host.Network(dims, cpm.CPM_FMT_U8, ns, [d], **kw)
    net.set_transfer_function(synth.WS_TF_POINTS)
    net.set_volume_host(vol)
    net.evaluate()
    pts = list(synth.WS_TF_POINTS)
    paths = []
    for alpha in (0.9, 0.3, 0.7, 0.45):
        pts[-1] = (pts[-1][0], (0.1, 0.6, 0.65, alpha))
        net.set_transfer_function(pts)
        net.evaluate()
        paths.append(net.last_splat_path)
        assert 0 < net.n_recomputed < net.n_photons
    assert paths == ["incremental"] * 4
    lv_inc = net.read_light_volume().astype(np.float64)
    ref = host.Network(dims, cpm.CPM_FMT_U8, ns, [d], **kw)
    ref.set_transfer_function(pts)
    ref.set_volume_host(vol)
    ref.evaluate()
    lv_full = ref.read_light_volume().astype(np.float64)
    same = np.all(ref.read_photons(I).view(np.uint32) == net.read_photons(I).view(np.uint32), axis=1)
    assert same.mean() > 0.999, same.mean()
    rmse = np.sqrt(((lv_inc - lv_full) ** 2).mean()) / np.sqrt((lv_full ** 2).mean())
    assert rmse < 4e-3, rmse
    net.close(); ref.close()


@pytest.mark.gpu
def test_network_budgeted_batches(host, cpm, synth, torch_cuda):
    """maxIncrementalPhotonsToUpdate < 100: the re-trace is spread over several evaluations"""
    dims, ns = (48, 48, 48), 64
    vol = synth.volume_u8(dims, 8)
    net = host.Network(dims, cpm.CPM_FMT_U8, ns, [(0, 0, 1)], with_importance_grid=True, max_incremental_percent=10.0)
    net.set_transfer_function(synth.WS_TF_POINTS)
    net.set_volume_host(vol)
    net.evaluate()
    pts = [(p, (c[0], c[1], c[2], min(1.0, c[3] * 1.5))) for p, c in synth.WS_TF_POINTS]
    net.set_transfer_function(pts)
    total, rounds = 0, 0
    net.evaluate()
    total += max(net.n_recomputed, 0)
    while net.remaining_photons > 0 and rounds < 50:
        assert net.evaluate() >= 1
        total += max(net.n_recomputed, 0)
        rounds += 1
    assert rounds >= 1 and total > 0
    assert net.n_recomputed <= int(0.10 * net.n_photons)
    net.close()


@pytest.mark.gpu
def test_random_number_generator_processors_match_oracle(host, orc, torch_cuda):
    """RandomNumberGeneratorCL / RandomNumberGenerator2DCL: stream i seeded from the host base offset of `seed`
    (rng/mwc64xseedgenerator.cpp:56-64), one number per stream and evaluation; pixel (x, y) of the 2-D generator is
    number y * width + x (rng/cl/randomnumbergenerator.cl:51-71).  Bit-exact."""
    for seed, evals in ((0, 1), (7, 3)):
        n = 96 * 40
        st = orc.rng_seed_streams(orc.rng_host_base_offsets(seed, n))
        want = None
        for _ in range(evals):
            want = orc.rng_uniform(st, 1)
        got1 = host.random_numbers(n, 0, seed, evals)
        got2 = host.random_numbers(96, 40, seed, evals)
        assert np.array_equal(got1.view(np.uint32), np.asarray(want, np.float32).reshape(-1).view(np.uint32))
        assert got2.shape == (40, 96) and np.array_equal(got2.reshape(-1).view(np.uint32), got1.view(np.uint32))
