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


# ---------------------------------------------------------------------------------------------------------------
# ProgressivePhotonTracerCL::process / MinMaxUniformGrid3DImportanceCLProcessor::process against the oracle's network
def _host_tf_difference(host, cur, prev, eps=1e-4, assoc=False):
    import ctypes as C
    a = np.array([[p, *c] for p, c in cur], np.float32)
    b = np.array([[p, *c] for p, c in prev], np.float32)
    pos, col = np.zeros(64, np.float32), np.zeros((64, 4), np.float32)
    n = host.lib().cpmh_tf_difference_points(a.ctypes.data_as(C.c_void_p), len(cur), b.ctypes.data_as(C.c_void_p), len(prev),
                                             C.c_float(eps), int(assoc), pos.ctypes.data_as(C.c_void_p),
                                             col.ctypes.data_as(C.c_void_p), 64)
    assert n >= 0
    return pos[:n], col[:n]


def _same_bits(a, b):
    return a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)) and \
        np.array_equal(a[~np.isnan(a)].view(np.uint32), b[~np.isnan(b)].view(np.uint32))


def test_tf_difference_lists_host_equals_oracle_restatement(host):
    """updateTransferFunctionDifferenceData (isc/processors/minmaxuniformgrid3dimportanceclprocessor.cpp:364-501): the host
    layer's merge walk and the checker's independent restatement (oracle/frame.py) give the same point lists bit for bit
    on 1500 random transfer-function pairs -- equal break points, moved zero-opacity first points, points at 0 and 1,
    associated colours, coincident positions (division by zero propagates as in the reference).  No GPU needed."""
    from oracle import frame
    rng = np.random.default_rng(3)

    def rnd(n, it):
        ps = np.sort(rng.uniform(0, 1, n))
        if it % 5 == 0:
            ps[0] = 0.0
        if it % 7 == 0:
            ps[-1] = 1.0
        pts = []
        for p in ps:
            c = rng.uniform(0, 1, 4)
            if rng.random() < 0.3:
                c[3] = 0.0
            pts.append((float(np.float32(p)), tuple(float(np.float32(x)) for x in c)))
        return pts

    for it in range(1500):
        cur = rnd(int(rng.integers(1, 7)), it)
        if it % 3 == 0:      # a small edit of the same function: the usual interactive case
            prev = list(cur)
            k = int(rng.integers(0, len(prev)))
            c = list(prev[k][1])
            c[int(rng.integers(0, 4))] = float(np.float32(rng.uniform(0, 1)))
            prev[k] = (prev[k][0], tuple(c))
            if it % 6 == 0:
                prev[0] = (float(np.float32(prev[0][0] * 0.5)), prev[0][1])
        else:
            prev = rnd(int(rng.integers(1, 7)), it + 1)
        assoc = it % 4 == 0
        hp, hc = _host_tf_difference(host, cur, prev, assoc=assoc)
        op, oc = frame.tf_difference_lists(cur, prev, associated=assoc)
        assert _same_bits(hp, op) and _same_bits(hc, oc), (it, cur, prev)
    # the documented quirk: the last point of a function enters the walk as (1, its colour), so a change of the last
    # point's colour ramps up to position 1 instead of ending at the point
    pts = list(importlib.import_module(PKG_NAME + ".synth").WS_TF_POINTS)
    new = list(pts)
    new[-1] = (new[-1][0], (0.1, 0.6, 0.65, 0.9))
    p, c = frame.tf_difference_lists(new, pts)
    assert p.tolist() == [0.0, float(np.float32(pts[-2][0])), 1.0] and c[2, 3] > 0.3 and not c[:2].any()


def _oracle_for(host, cpm, orc, synth, net, dims, ns, I, **kw):
    from oracle import frame
    from test_configs import oracle_lights_from_host
    return frame.OracleNetwork(dims, oracle_lights_from_host(orc, synth, net, ns), frame.rasterise_tf(synth.WS_TF_POINTS),
                               synth.WS_TF_POINTS, max_interactions=I, **kw)


def _rel_rmse(got, want):
    return float(np.sqrt(((got.astype(np.float64) - want) ** 2).mean()) / max(np.sqrt((want ** 2).mean()), 1e-300))


@pytest.mark.gpu
def test_network_tf_change_matches_oracle(host, cpm, orc, synth, torch_cuda):
    """Transfer-function changes through the plug-in (call stack D with R::TransferFunction): the importance processor
    builds the |new - old| point list, classifies with the incremental formula, the tracer detects / selects / re-traces,
    the light volume is updated by -old/+new.  Three consecutive edits, each frame compared with the oracle's network:
    point lists, importance grid, re-traced ids, photon records (bit for bit), light volume (1e-5 relative RMSE)."""
    dims, ns, I = (64, 64, 64), 128, 2
    vol = synth.volume_u8(dims, 8)
    net = host.Network(dims, cpm.CPM_FMT_U8, ns, [(0.2, 0.3, 0.9)], max_scattering_events=I, light_volume_option=2,
                       with_importance_grid=True, reference_full_splat_bound=False)
    net.set_transfer_function(synth.WS_TF_POINTS)
    net.set_volume_host(vol)
    net.evaluate()
    O = _oracle_for(host, cpm, orc, synth, net, dims, ns, I)
    O.first_frame(vol)
    assert np.array_equal(net.read_photons(I).view(np.uint32), O.photons.view(np.uint32))
    mm = orc.volume_minmax(vol, 8)
    pts = list(synth.WS_TF_POINTS)
    for k, alpha in enumerate((0.9, 0.3, 0.7)):
        pts[-1] = (pts[-1][0], (0.1, 0.6, 0.65, alpha))
        if k == 1:
            pts[2] = (pts[2][0], (0.9, 0.5, 0.4, 0.05))      # a second point, previously transparent, becomes visible
        net.set_transfer_function(pts)
        net.evaluate()
        pos, col = O.set_transfer_function(pts)
        hp, hc = net.importance_tf_points()
        assert _same_bits(hp, pos) and _same_bits(hc, col), k
        imp = O.importance_static(mm, pos, col)
        assert np.array_equal(net.read_importance_grid(imp.size).view(np.uint32), imp.view(np.uint32)), k
        ids = O.frame(vol, imp)
        assert 0 < ids.size < O.n
        assert net.n_recomputed == ids.size and np.array_equal(net.read_recomputed_indices(), ids), k
        assert np.array_equal(net.read_photons(I).view(np.uint32), O.photons.view(np.uint32)), k
        assert net.last_splat_path == O.last_splat_path, (k, net.last_splat_path, O.last_splat_path)
        assert _rel_rmse(net.read_light_volume(), O.lightvol) <= 1e-5, k
    net.close()


@pytest.mark.gpu
@pytest.mark.parametrize("equal", [False, True])
def test_network_budgeted_batches_match_oracle(host, cpm, orc, synth, torch_cuda, equal):
    """maxIncrementalPhotonsToUpdate = 10 %: the invalid photons are re-traced in importance order, one budget-sized batch
    per evaluation (progressivephotontracercl.cpp:361-363, 425-540); with equalImportance every (100 / 10)-th photon is
    flagged instead (photonrecomputationdetector.cl:160-194).  Every batch: the same ids, in the same order, and the same
    records as the oracle's network; the keys of photons still waiting are equal too."""
    dims, ns, I = (48, 48, 48), 96, 1
    vols = [synth.volume_f32(dims, 4, t / 8.0) for t in range(2)]
    net = host.Network(dims, cpm.CPM_FMT_F32, ns, [(0.0, 0.2, 1.0)], max_scattering_events=I, with_importance_grid=True,
                       max_incremental_percent=10.0, reference_full_splat_bound=False)
    net.set_transfer_function(synth.WS_TF_POINTS)
    if equal:
        net.set_property("org.inviwo.ProgressivePhotonTracerCL", "equalImportance", True)
    net.set_sequence_host(vols)
    net.set_timestep(0)
    net.evaluate()
    O = _oracle_for(host, cpm, orc, synth, net, dims, ns, I, budget_percent=10.0)
    O.first_frame(vols[0])
    mm = [orc.volume_minmax(v, 8) for v in vols]
    diff = orc.volume_diff_bricks(vols[0], vols[1], 8)
    net.set_timestep(1)
    net.evaluate()
    imp = O.importance_time_varying(mm[1], mm[0], diff)
    O.detect(imp, equal_importance=equal)
    n_inv = O.select()
    assert n_inv > O.n // 10 or equal
    rounds = 0
    while True:
        ids = O.retrace_batch(vols[1])
        assert net.n_recomputed == ids.size, (rounds, net.n_recomputed, ids.size)
        assert np.array_equal(net.read_recomputed_indices(), ids), rounds
        assert np.array_equal(net.read_photons(I).view(np.uint32), O.photons.view(np.uint32)), rounds
        assert np.array_equal(net.read_importance_keys(), O.keys), rounds
        assert net.remaining_photons == O.remaining, rounds
        rounds += 1
        if net.remaining_photons <= 0 or rounds > 40:
            break
        assert net.evaluate() >= 1
    assert rounds >= (1 if equal else 2)
    assert _rel_rmse(net.read_light_volume(), O.lightvol) <= 1e-5
    net.close()


@pytest.mark.gpu
def test_network_progressive_refinement_without_importance_grid(host, cpm, orc, synth, torch_cuda):
    """enableRefinement with no importance grid connected (call stack C): every evaluation traces ALL photons with the RNG
    state carried over (new random numbers, -D PROGRESSIVE_PHOTON_MAPPING), the photon radius follows Knaus & Zwicker
    (ppm/photondata.cpp:70-80) and the records equal the oracle's progressive traces bit for bit."""
    from oracle import frame
    from test_configs import oracle_lights_from_host
    dims, ns, I = (48, 48, 48), 64, 2
    vol = synth.volume_u8(dims, 5)
    net = host.Network(dims, cpm.CPM_FMT_U8, ns, [(0.1, -0.2, 1.0)], max_scattering_events=I, with_importance_grid=False)
    net.set_transfer_function(synth.WS_TF_POINTS)
    net.set_property("org.inviwo.ProgressivePhotonTracerCL", "enableRefinement", True)
    net.set_volume_host(vol)
    net.evaluate()
    L = oracle_lights_from_host(orc, synth, net, ns)[0]
    tf = frame.rasterise_tf(synth.WS_TF_POINTS)
    n = ns * ns
    rng = scenes.rng_states(n)
    p = orc.trace_params(n_light_samples=n, max_interactions=I, step_size=1.0 / 48, flags=1)
    want = np.zeros((n * I, 8), np.float32)
    orc.trace_photons(orc.volume(vol), tf, p, L["light_samples"], L["isect"], want, rng)
    assert np.array_equal(net.read_photons(I).view(np.uint32), want.view(np.uint32))
    s0 = net.photon_state()
    assert s0["iteration"] == 1
    radii = [s0["radius"]]
    for it in range(1, 4):
        net.tracer_timer_event()
        assert net.evaluate() >= 1
        orc.trace_photons(orc.volume(vol), tf, p, L["light_samples"], L["isect"], want, rng)     # rng advanced by the last call
        assert np.array_equal(net.read_photons(I).view(np.uint32), want.view(np.uint32)), it
        s = net.photon_state()
        assert s["iteration"] == it + 1
        radii.append(s["radius"])
        assert abs(radii[-1] / radii[-2] - ((it + 0.5) / (it + 1.0)) ** (1.0 / 3.0)) < 1e-12, (it, radii)
    net.close()


@pytest.mark.gpu
def test_network_data_range_change_reaches_sampling_and_bound(host, cpm, orc, synth, torch_cuda):
    """12-bit data in a 16-bit volume: dataMap_.dataRange = (0, 4095) set AFTER the first evaluation.  The tracer's
    volume handle, the value-range grid and the opacity bound were built with the old scaling; all must follow, or the
    bounded tracer skips real collisions.  Records equal the oracle's with scale 65535 / 4095, bit for bit."""
    from oracle import frame
    from test_configs import oracle_lights_from_host
    dims, ns, I = (48, 48, 48), 64, 2
    vol = (synth.volume_u16(dims, 6) >> 4).astype(np.uint16)          # values 0 ... 4095
    net = host.Network(dims, cpm.CPM_FMT_U16, ns, [(0.1, -0.2, 1.0)], max_scattering_events=I)
    net.set_transfer_function(synth.WS_TF_POINTS)
    net.set_volume_host(vol)
    net.evaluate()
    L = dict(oracle_lights_from_host(orc, synth, net, ns)[0], n=ns * ns)
    tf = frame.rasterise_tf(synth.WS_TF_POINTS)
    p = orc.trace_params(n_light_samples=ns * ns, max_interactions=I, step_size=1.0 / 48)
    for rng_hi, scale in ((65535.0, 1.0), (4095.0, float(np.float32(65535.0 / 4095.0)))):
        if rng_hi != 65535.0:
            net.set_data_range(0.0, rng_hi)
            assert net.evaluate() >= 1
        want = np.zeros((ns * ns * I, 8), np.float32)
        orc.trace_photons(orc.volume(vol, scale=scale, offset=0.0), tf, p, L["light_samples"], L["isect"], want, scenes.rng_states(ns * ns))
        got = net.read_photons(I)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), rng_hi
    stored = (want[:, 0] != FLT_MAX).sum()
    assert stored > 500
    net.close()


@pytest.mark.gpu
def test_network_align_changed_photons_path(host, cpm, synth, torch_cuda):
    """`alignChangedPhotons` (photontolightvolumeprocessorcl.cpp:207-244): the incremental update through the packed
    -old / +new buffer gives the light volume of the default (index-list) update up to the order of the float adds, and
    keeps doing so over consecutive updates (the copy of the previous records is refreshed after every frame)."""
    dims, ns, I = (64, 64, 64), 96, 2
    vol = synth.volume_u8(dims, 8)
    kw = dict(max_scattering_events=I, light_volume_option=2, with_importance_grid=True, reference_full_splat_bound=False,
              incremental_threshold=100.0)
    nets = [host.Network(dims, cpm.CPM_FMT_U8, ns, [(0.2, 0.3, 0.9)], **kw) for _ in range(2)]
    nets[1].set_property("org.inviwo.PhotonToLightVolumeProcessorCL", "alignChangedPhotons", 1)
    for net in nets:
        net.set_transfer_function(synth.WS_TF_POINTS)
        net.set_volume_host(vol)
        net.evaluate()
    pts = list(synth.WS_TF_POINTS)
    for alpha in (0.9, 0.35, 0.6):
        pts[-1] = (pts[-1][0], (0.1, 0.6, 0.65, alpha))
        for net in nets:
            net.set_transfer_function(pts)
            net.evaluate()
        assert nets[0].last_splat_path == "incremental" and nets[1].last_splat_path == "incremental-aligned"
        assert 0 < nets[0].n_recomputed == nets[1].n_recomputed
        a, b = (net.read_light_volume().astype(np.float64) for net in nets)
        rmse = np.sqrt(((a - b) ** 2).mean()) / np.sqrt((a ** 2).mean())
        assert rmse < 1e-5, rmse
    for net in nets:
        net.close()
