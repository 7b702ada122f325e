"""Delta-tracking photon tracer: oracle sanity (closed forms) and CUDA bit-exact parity."""
import numpy as np
import pytest

import scenes

FLT_MAX = np.float32(3.4028234663852886e38)


def oracle_trace(orc, vol_np, tf, L, max_interactions=1, flags=0, phase=0, material=(0, 0, 0, 0), rng=None,
                 recompute=None, photons=None, step_size=1.0 / 64, aabb=((0, 0, 0), (1, 1, 1)),
                 total_photons=None, photon_offset=0):
    n = L["n"]
    total = total_photons or n
    rng = scenes.rng_states(total) if rng is None else rng
    if photons is None:
        photons = np.zeros((total * max_interactions, 8), np.float32)
    p = orc.trace_params(n_light_samples=n, total_photons=total, photon_offset=photon_offset,
                         max_interactions=max_interactions, flags=flags, phase=phase, material=material,
                         step_size=step_size, aabb_min=aabb[0], aabb_max=aabb[1])
    tests = orc.trace_photons(orc.volume(vol_np), tf, p, L["light_samples"], L["isect"], photons, rng,
                              recompute, 0 if recompute is None else len(recompute))
    return photons, rng, tests


# ------------------------------------------------------------------ CPU: closed-form checks --
def test_homogeneous_slab_transmission(orc, synth):
    """Constant opacity sigma: P(no collision across the slab) = exp(-150 * sigma * L)
    (SAMPLING_BASE_INTERVAL_RCP = 150, ppm/cl/transmittance.cl:40,130)."""
    vol = np.full((8, 8, 8), 128, np.uint8)
    sigma = 0.02
    tf = synth.dense_tf(sigma, 64)
    L = scenes.directional_light(128, (0.0, 0.0, 1.0))
    photons, _, tests = oracle_trace(orc, vol, tf, L)
    hit = L["isect"][:, 0] < L["isect"][:, 1]
    escaped = photons[:L["n"], 0] == FLT_MAX
    frac = escaped[hit].mean()
    want = np.exp(-150.0 * sigma * 1.0)
    assert abs(frac - want) < 4 * np.sqrt(want * (1 - want) / hit.sum()) + 1e-3
    # first-collision depth ~ Exp(150 sigma) truncated to the slab
    z = photons[:L["n"], 2][hit & ~escaped]
    lam = 150.0 * sigma
    mean_want = 1 / lam - np.exp(-lam) / (1 - np.exp(-lam))
    assert abs(z.mean() - mean_want) < 0.01
    assert tests > hit.sum()


def test_sentinel_and_power_conventions(orc, synth):
    """slot padding (ppm/cl/photontracer.cl:199-209) and power bookkeeping (:129,176,180)"""
    vol = scenes.make_volume((32, 32, 32), "u8", 11)
    tf = synth.rasterise_tf(width=256)
    L = scenes.directional_light(64)
    I = 4
    photons, _, _ = oracle_trace(orc, vol, tf, L, max_interactions=I)
    n = L["n"]
    ph = photons.reshape(I, n, 8)
    stored = ph[:, :, 0] != FLT_MAX
    # stored slots form a prefix per photon
    assert (np.diff(stored.astype(int), axis=0) <= 0).all()
    # empty slots: xyz and power.gb are FLT_MAX; power.r is FLT_MAX iff the photon was absorbed
    empty = ~stored
    assert (ph[empty][:, [0, 1, 2, 4, 5]] == FLT_MAX).all()
    miss = L["isect"][:, 0] >= L["isect"][:, 1]
    assert (ph[0, miss, 0] == FLT_MAX).all()
    # a ray that never interacts keeps power/maxInteractions in slot 0's power.r
    never = empty[0] & ~miss
    if never.any():
        assert np.allclose(ph[0, never, 3], L["light_samples"][never, 3] / I)
    # stored positions are inside the volume
    pos = ph[stored][:, 0:3]
    assert pos.min() >= -1e-3 and pos.max() <= 1 + 1e-3


def test_replay_property(orc, synth):
    """Correlation: re-tracing with the same RNG state reproduces the photon exactly."""
    vol = scenes.make_volume((32, 32, 32), "f32", 5)
    tf = synth.rasterise_tf(width=256)
    L = scenes.directional_light(32)
    a, rng_a, _ = oracle_trace(orc, vol, tf, L, max_interactions=2)
    idx = np.arange(0, L["n"], 3, dtype=np.uint32)
    b = a.copy()
    b[:] = -7.0
    oracle_trace(orc, vol, tf, L, max_interactions=2, photons=b, recompute=idx)
    n = L["n"]
    for k in range(2):
        assert np.array_equal(b[k * n + idx].view(np.uint32), a[k * n + idx].view(np.uint32))
    untouched = np.setdiff1d(np.arange(n), idx)
    assert (b[untouched] == -7.0).all()


# ------------------------------------------------------------------ GPU: CUDA vs oracle ------
def cuda_trace(cpm, ctx, torch, vol_np, tf, L, layout, max_interactions=1, flags=0, phase=0, material=(0, 0, 0, 0),
               rng=None, recompute=None, photons=None, step_size=1.0 / 64, aabb=((0, 0, 0), (1, 1, 1)),
               total_photons=None, photon_offset=0, opacity_bound=None, bound_cell_log2=3, opacity_bound_tex=None):
    """returns (photons, rng, collision tests) and, with an opacity bound, the number of tests that fetched voxels"""
    n = L["n"]
    total = total_photons or n
    rng = scenes.rng_states(total) if rng is None else rng
    fmt = {np.dtype(np.uint8): cpm.CPM_FMT_U8, np.dtype(np.uint16): cpm.CPM_FMT_U16,
           np.dtype(np.float32): cpm.CPM_FMT_F32}[vol_np.dtype]
    raw = vol_np.view(np.int16) if vol_np.dtype == np.uint16 else vol_np
    dvol = torch.from_numpy(raw).cuda()
    dims = (vol_np.shape[2], vol_np.shape[1], vol_np.shape[0])
    V = ctx.volume_create(dvol, dims, fmt, layout=layout)
    dtf = torch.from_numpy(tf).cuda()
    dls = torch.from_numpy(L["light_samples"]).cuda()
    dis = torch.from_numpy(L["isect"]).cuda()
    drng = torch.from_numpy(rng.copy().view(np.int32)).cuda()
    if photons is None:
        dph = torch.zeros(total * max_interactions * 8, dtype=torch.float32, device="cuda")
    else:
        dph = torch.from_numpy(photons.copy()).cuda()
    dcount = torch.zeros(2, dtype=torch.int64, device="cuda")
    p = cpm.make_trace_params(n, total_photons=total, photon_offset=photon_offset, max_interactions=max_interactions,
                              flags=flags | cpm.CPM_TRACE_STATS, phase=phase, material=material, step_size=step_size,
                              aabb_min=aabb[0], aabb_max=aabb[1], opacity_bound=opacity_bound,
                              bound_cell_log2=bound_cell_log2, opacity_bound_tex=opacity_bound_tex)
    didx = None if recompute is None else torch.from_numpy(recompute.view(np.int32)).cuda()
    ctx.trace_photons(V, dtf, p, dls, dis, dph, drng, didx, 0 if recompute is None else len(recompute), dcount)
    ctx.sync()
    counts = dcount.cpu().numpy()
    out = dph.cpu().numpy().reshape(-1, 8), drng.cpu().numpy().view(np.uint32), int(counts[0])
    if opacity_bound is None and opacity_bound_tex is None:
        assert counts[1] == counts[0]
    else:
        out = out + (int(counts[1]),)
    V.destroy()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["linear", "texture"])
@pytest.mark.parametrize("fmt", ["u8", "u16", "f32"])
def test_cuda_tracer_bit_exact(cpm, orc, ctx, torch_cuda, synth, fmt, layout):
    lay = cpm.CPM_VOLUME_LINEAR if layout == "linear" else cpm.CPM_VOLUME_TEXTURE
    vol = scenes.make_volume((48, 40, 36), fmt, 21)
    tf = synth.rasterise_tf(width=512)
    L = scenes.directional_light(96, (0.25, -0.4, 0.85))
    for I, flags in ((1, 0), (4, cpm.CPM_TRACE_PROGRESSIVE)):
        want, want_rng, want_tests = oracle_trace(orc, vol, tf, L, max_interactions=I, flags=flags)
        got, got_rng, got_tests = cuda_trace(cpm, ctx, torch_cuda, vol, tf, L, lay, max_interactions=I, flags=flags)
        assert got_tests == want_tests
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        assert np.array_equal(got_rng, want_rng)


@pytest.mark.gpu
def test_cuda_tracer_variants_bit_exact(cpm, orc, ctx, torch_cuda, synth):
    """point light, clip box, Henyey-Greenstein, NO_SINGLE_SCATTERING, multi-light offsets"""
    vol = scenes.make_volume((64, 64, 64), "u8", 8)
    tf = synth.rasterise_tf(width=1024)
    Lp = scenes.point_light(80)
    Ld = scenes.directional_light(64)
    aabb = ((73 / 512, 7 / 512, 0.0), (1.0, 1.0, 1.0))   # the workspace clip box (ws:740-757)
    cases = [
        dict(L=Lp, max_interactions=3),
        dict(L=Ld, max_interactions=4, aabb=aabb),
        dict(L=Ld, max_interactions=5, phase=1, material=(0.6, 0, 0, 0)),
        dict(L=Ld, max_interactions=3, flags=cpm.CPM_TRACE_NO_SINGLE_SCATTERING),
        dict(L=Ld, max_interactions=2, total_photons=3 * Ld["n"], photon_offset=Ld["n"]),
    ]
    for kw in cases:
        L = kw.pop("L")
        want, want_rng, wt = oracle_trace(orc, vol, tf, L, **kw)
        for lay in (cpm.CPM_VOLUME_LINEAR, cpm.CPM_VOLUME_TEXTURE):
            got, got_rng, gt = cuda_trace(cpm, ctx, torch_cuda, vol, tf, L, lay, **kw)
            assert gt == wt, kw
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), kw
            assert np.array_equal(got_rng, want_rng)


@pytest.mark.gpu
def test_cuda_retrace_index_list(cpm, orc, ctx, torch_cuda, synth):
    """-D PHOTON_RECOMPUTATION: only listed photons are rewritten, ids of other lights are skipped"""
    vol = scenes.make_volume((40, 40, 40), "f32", 4)
    tf = synth.rasterise_tf(width=256)
    L = scenes.directional_light(50)
    n = L["n"]
    total, off = 2 * n, n
    rngs = scenes.rng_states(total)
    idx = np.concatenate([np.arange(3, n, 5), np.arange(n + 1, 2 * n, 7)]).astype(np.uint32)  # both lights' ids
    base = np.full((total * 2, 8), -3.0, np.float32)
    want, _, wt = oracle_trace(orc, vol, tf, L, max_interactions=2, rng=rngs, recompute=idx, photons=base.copy(),
                               total_photons=total, photon_offset=off)
    got, _, gt = cuda_trace(cpm, ctx, torch_cuda, vol, tf, L, cpm.CPM_VOLUME_TEXTURE, max_interactions=2, rng=rngs,
                            recompute=idx, photons=base.copy(), total_photons=total, photon_offset=off)
    assert gt == wt
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    touched = np.zeros(total * 2, bool)
    sel = idx[idx >= off]
    touched[sel] = True
    touched[total + sel] = True
    assert (got[~touched] == -3.0).all() and (got[touched] != -3.0).any()


@pytest.mark.gpu
def test_cuda_tracer_argument_errors(cpm, ctx, torch_cuda):
    torch = torch_cuda
    v = torch.zeros(8, dtype=torch.uint8, device="cuda")
    with pytest.raises(cpm.CpmError):
        ctx.volume_create(v, (0, 2, 2), cpm.CPM_FMT_U8)
    V = ctx.volume_create(v, (2, 2, 2), cpm.CPM_FMT_U8)
    tf = torch.zeros(16, dtype=torch.float32, device="cuda")
    p = cpm.make_trace_params(4, max_interactions=0)
    buf = torch.zeros(64, dtype=torch.float32, device="cuda")
    with pytest.raises(cpm.CpmError) as e:
        ctx.trace_photons(V, tf, p, buf, buf, buf, buf.view(torch.int32))
    assert e.value.code == -1
