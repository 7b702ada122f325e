"""Temporal-correlation detector and splat density estimation."""
import numpy as np
import pytest

import scenes
from test_grid import tf_points
from test_tracer import oracle_trace

FLT_MAX = np.float32(3.4028234663852886e38)


def make_case(orc, synth, cpm_mod, I=3, n_side=64, dims=(64, 64, 64)):
    vol = synth.volume_u8(dims, 8)
    tf = synth.rasterise_tf(width=512)
    L = scenes.directional_light(n_side)
    photons, _, _ = oracle_trace(orc, vol, tf, L, max_interactions=I)
    mm = orc.volume_minmax(vol, 8)
    pos, col = tf_points(synth)
    grid = orc.classify_importance(mm, pos, col, (0, 0, 0, 1), True)
    gd = (mm.shape[2], mm.shape[1], mm.shape[0])
    return dict(vol=vol, tf=tf, L=L, photons=photons, grid=grid, grid_dims=gd, cell=(8.0, 8.0, 8.0),
                tex2idx=cpm_mod.capi.texture_to_index_matrix(dims), I=I, dims=dims)


def test_oracle_detector_zero_grid_leaves_keys(orc, synth, cpm):
    c = make_case(orc, synth, cpm)
    n = c["L"]["n"]
    keys = np.full(n, 0x7FFFFFFF, np.uint32)
    orc.detect_invalid(np.zeros_like(c["grid"]), c["grid_dims"], c["cell"], c["tex2idx"], c["photons"], 0,
                       c["L"]["light_samples"], c["L"]["isect"], n, c["I"], n, keys)
    assert (keys == 0x7FFFFFFF).all()
    # constant grid: importance = ceil(100 * path length in voxels * value) for hit rays, 0 for misses
    keys[:] = 0x7FFFFFFF
    orc.detect_invalid(np.ones_like(c["grid"]), c["grid_dims"], c["cell"], c["tex2idx"], c["photons"], 0,
                       c["L"]["light_samples"], c["L"]["isect"], n, c["I"], n, keys, fix_exit=True)
    miss = c["L"]["isect"][:, 0] >= c["L"]["isect"][:, 1]
    assert (keys[miss] == 0x7FFFFFFF).all()
    assert (keys[~miss] < 0x7FFFFFFF).all()
    # single-interaction absorbed photon: one segment entry -> photon, importance ~ 100 * voxel distance
    ph = c["photons"].reshape(c["I"], n, 8)
    one = (~miss) & (ph[0, :, 0] != FLT_MAX) & (ph[1, :, 0] == FLT_MAX) & (ph[1, :, 3] == FLT_MAX)
    ls = c["L"]["light_samples"]
    entry = ls[one, 0:3] + c["L"]["isect"][one, 0:1] * c["L"]["dir"]
    dist = np.linalg.norm((ph[0, one, 0:3] - entry) * 64.0, axis=1)
    got = (0x7FFFFFFF - keys[one]).astype(np.float64)
    assert np.allclose(got, np.ceil(100 * dist), atol=2 + 1e-3 * 100 * dist.max())


def test_oracle_equal_importance_round_robin(orc):
    """ppm/cl/photonrecomputationdetector.cl:186: (id + iteration) % (100/percentage) == 0"""
    keys = np.full(20, 0x7FFFFFFF, np.uint32)
    orc.detect_invalid(None, (1, 1, 1), (1, 1, 1), [0] * 16, None, 0, None, None, 20, 1, 20, keys,
                       equal_importance=True, percentage=25, iteration=3)
    sel = np.nonzero(keys < 0x7FFFFFFF)[0]
    assert sel.tolist() == [i for i in range(20) if (i + 3) % 4 == 0]
    assert (keys[sel] == 0x7FFFFFFF - 100).all()


def test_oracle_splat_energy(orc, synth, cpm):
    """sum over voxels == sum over photons of power * sum of kernel weights; indexed -1/+1 cancels"""
    c = make_case(orc, synth, cpm, I=2)
    n = c["L"]["n"]
    od = (32, 32, 32)
    t2i, i2t = cpm.capi.texture_to_index_matrix(od), cpm.capi.index_to_texture_matrix(od)
    vol = np.zeros(od[0] * od[1] * od[2], np.float64)
    orc.splat(vol, 1, t2i, i2t, od, c["photons"], None, n, n, 2, 1.5 / 32, 1.0)
    assert vol.sum() > 0 and vol.min() >= 0
    idx = np.arange(0, n, 3, dtype=np.uint32)
    v2 = np.zeros_like(vol)
    orc.splat(v2, 1, t2i, i2t, od, c["photons"], idx, idx.size, n, 2, 1.5 / 32, 1.0, 1.0)
    orc.splat(v2, 1, t2i, i2t, od, c["photons"], idx, idx.size, n, 2, 1.5 / 32, 1.0, -1.0)
    assert np.abs(v2).max() < 1e-12 * max(1.0, vol.max())


@pytest.mark.gpu
@pytest.mark.parametrize("fix_exit", [False, True])
def test_cuda_detector_bit_exact(cpm, orc, ctx, torch_cuda, synth, fix_exit):
    torch = torch_cuda
    c = make_case(orc, synth, cpm, I=3, n_side=80)
    n = c["L"]["n"]
    total, off = 2 * n, n     # second light of two
    photons = np.zeros((total * c["I"], 8), np.float32)
    ph_src = c["photons"].reshape(c["I"], n, 8)
    view = photons.reshape(c["I"], total, 8)
    view[:, off:off + n] = ph_src
    want = np.full(total, 0x7FFFFFFF, np.uint32)
    orc.detect_invalid(c["grid"], c["grid_dims"], c["cell"], c["tex2idx"], photons, off, c["L"]["light_samples"],
                       c["L"]["isect"], n, c["I"], total, want, fix_exit=fix_exit)
    dgrid = torch.from_numpy(c["grid"]).cuda()
    dph = torch.from_numpy(photons).cuda()
    dls, dis = torch.from_numpy(c["L"]["light_samples"]).cuda(), torch.from_numpy(c["L"]["isect"]).cuda()
    keys = torch.from_numpy(np.full(total, 0x7FFFFFFF, np.uint32).view(np.int32)).cuda()
    ctx.detect_invalid(dgrid, c["grid_dims"], c["cell"], c["tex2idx"], dph, off, dls, dis, n, c["I"], total, keys,
                       flags=cpm.capi.CPM_DETECT_FIX_EXIT if fix_exit else 0)
    ctx.sync()
    got = keys.cpu().numpy().view(np.uint32)
    assert np.array_equal(got, want)
    assert (got[:off] == 0x7FFFFFFF).all() and (got[off:] < 0x7FFFFFFF).any()
    # equal-importance variant
    want2 = np.full(total, 0x7FFFFFFF, np.uint32)
    orc.detect_invalid(None, (1, 1, 1), (1, 1, 1), [0] * 16, None, off, None, None, n, c["I"], total, want2,
                       equal_importance=True, percentage=10, iteration=7)
    keys2 = torch.from_numpy(np.full(total, 0x7FFFFFFF, np.uint32).view(np.int32)).cuda()
    ctx.detect_invalid(None, (1, 1, 1), (1, 1, 1), [0] * 16, None, off, None, None, n, c["I"], total, keys2,
                       equal_importance=True, percentage=10, iteration=7)
    ctx.sync()
    assert np.array_equal(keys2.cpu().numpy().view(np.uint32), want2)


@pytest.mark.gpu
@pytest.mark.parametrize("channels", [1, 4])
def test_cuda_splat_matches_double_oracle(cpm, orc, ctx, torch_cuda, synth, channels):
    torch = torch_cuda
    c = make_case(orc, synth, cpm, I=2, n_side=96)
    n = c["L"]["n"]
    od = (32, 32, 32)
    nvox = od[0] * od[1] * od[2]
    t2i, i2t = cpm.capi.texture_to_index_matrix(od), cpm.capi.index_to_texture_matrix(od)
    radius, scale = 1.3 / 32, 0.37
    want = np.zeros(nvox * channels, np.float64)
    orc.splat(want, channels, t2i, i2t, od, c["photons"], None, n, n, 2, radius, scale)
    dph = torch.from_numpy(c["photons"]).cuda()
    lv = torch.zeros(nvox * channels, dtype=torch.float32, device="cuda")
    ctx.splat_photons(lv, channels, t2i, i2t, od, dph, None, n, n, 2, radius, scale)
    ctx.sync()
    got = lv.cpu().numpy().astype(np.float64)
    rmse = np.sqrt(((got - want) ** 2).mean()) / np.sqrt((want ** 2).mean())
    assert rmse < 1e-5, rmse
    assert abs(got.sum() - want.sum()) < 1e-5 * want.sum()
    # incremental update: remove old contribution of a subset (-1), add it back (+1)
    idx = np.arange(1, n, 4, dtype=np.uint32)
    didx = torch.from_numpy(idx.view(np.int32)).cuda()
    w2 = want.copy()
    orc.splat(w2, channels, t2i, i2t, od, c["photons"], idx, idx.size, n, 2, radius, scale, -1.0)
    ctx.splat_photons(lv, channels, t2i, i2t, od, dph, didx, idx.size, n, 2, radius, scale, -1.0)
    ctx.sync()
    g2 = lv.cpu().numpy().astype(np.float64)
    assert np.sqrt(((g2 - w2) ** 2).mean()) / np.sqrt((want ** 2).mean()) < 1e-5
    ctx.splat_photons(lv, channels, t2i, i2t, od, dph, didx, idx.size, n, 2, radius, scale, 1.0)
    ctx.sync()
    g3 = lv.cpu().numpy().astype(np.float64)
    assert np.sqrt(((g3 - want) ** 2).mean()) / np.sqrt((want ** 2).mean()) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("channels", [1, 4])
def test_cuda_splat_update_equals_remove_then_add(cpm, orc, ctx, torch_cuda, synth, channels):
    """cpm_splat_photons_update == splatSelected(-1, old records) + splatSelected(+1, new records)
    (ppm/processor/photontolightvolumeprocessorcl.cpp:262-274); unchanged records are skipped"""
    torch = torch_cuda
    c = make_case(orc, synth, cpm, I=2, n_side=96)
    n = c["L"]["n"]
    od = (32, 32, 32)
    nvox = od[0] * od[1] * od[2]
    t2i, i2t = cpm.capi.texture_to_index_matrix(od), cpm.capi.index_to_texture_matrix(od)
    radius, scale = 1.3 / 32, 0.37
    old = c["photons"]
    new = old.copy()
    ph = new.reshape(2, n, 8)
    moved = np.arange(0, n, 3)
    stored = ph[0, moved, 0] != np.float32(3.4028234663852886e38)
    ph[0, moved[stored], 0:3] = np.clip(ph[0, moved[stored], 0:3] + np.float32(0.03), 0, 1)    # a third of the photons move
    ph[0, moved[stored], 3:6] *= np.float32(1.5)
    idx = np.arange(0, n, 2, dtype=np.uint32)                                               # half are listed
    want = np.zeros(nvox * channels, np.float64)
    orc.splat(want, channels, t2i, i2t, od, old, None, n, n, 2, radius, scale)
    base = want.copy()
    orc.splat(want, channels, t2i, i2t, od, old, idx, idx.size, n, 2, radius, scale, -1.0)
    orc.splat(want, channels, t2i, i2t, od, new, idx, idx.size, n, 2, radius, scale, 1.0)
    lv = torch.from_numpy(base.astype(np.float32)).cuda()
    ctx.splat_photons_update(lv, channels, t2i, i2t, od, torch.from_numpy(old).cuda(), torch.from_numpy(new).cuda(),
                             torch.from_numpy(idx.view(np.int32)).cuda(), idx.size, n, 2, radius, scale)
    ctx.sync()
    got = lv.cpu().numpy().astype(np.float64)
    assert np.abs(want - base).max() > 0
    assert np.sqrt(((got - want) ** 2).mean()) / np.sqrt((want ** 2).mean()) < 1e-5
    # the _sync variant: same light volume, and the old-record buffer ends up holding the new records of the listed
    # ids (both interactions) and its own records everywhere else
    lv3 = torch.from_numpy(base.astype(np.float32)).cuda()
    dold = torch.from_numpy(old).cuda()
    ctx.splat_photons_update(lv3, channels, t2i, i2t, od, dold, torch.from_numpy(new).cuda(),
                             torch.from_numpy(idx.view(np.int32)).cuda(), idx.size, n, 2, radius, scale, sync=True)
    ctx.sync()
    g3 = lv3.cpu().numpy().astype(np.float64)
    assert np.sqrt(((g3 - want) ** 2).mean()) / np.sqrt((want ** 2).mean()) < 1e-5
    expect = old.reshape(2, n, 8).copy()
    expect[:, idx] = new.reshape(2, n, 8)[:, idx]
    assert np.array_equal(dold.cpu().numpy().reshape(2, n, 8).view(np.uint32), expect.view(np.uint32))
    # no listed record changed: the light volume is untouched bit for bit
    lv2 = torch.from_numpy(base.astype(np.float32)).cuda()
    ctx.splat_photons_update(lv2, channels, t2i, i2t, od, torch.from_numpy(old).cuda(), torch.from_numpy(old).cuda(),
                             torch.from_numpy(idx.view(np.int32)).cuda(), idx.size, n, 2, radius, scale)
    ctx.sync()
    assert np.array_equal(lv2.cpu().numpy(), base.astype(np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("channels", [1, 4])
def test_cuda_copy_index_photons_then_splat_equals_update(cpm, orc, ctx, torch_cuda, synth, channels):
    """copyIndexPhotonsKernel (ppm/cl/photonstolightvolume.cl:225-247): packed[off + g + k * n] = record (k, ids[g]) with
    its power times the multiplier, bit for bit; splatting the packed -old / +new records as plain photons is the
    `alignChangedPhotons` update (photontolightvolumeprocessorcl.cpp:207-244)"""
    torch = torch_cuda
    c = make_case(orc, synth, cpm, I=2, n_side=64)
    n = c["L"]["n"]
    old = c["photons"]
    new = old.copy()
    ph = new.reshape(2, n, 8)
    moved = np.arange(1, n, 3)
    stored = ph[0, moved, 0] != np.float32(3.4028234663852886e38)
    ph[0, moved[stored], 0:3] = np.clip(ph[0, moved[stored], 0:3] - np.float32(0.02), 0, 1)
    idx = np.arange(1, n, 2, dtype=np.uint32)
    m = idx.size
    packed = torch.zeros(4 * m * 8, dtype=torch.float32, device="cuda")
    d_idx = torch.from_numpy(idx.view(np.int32)).cuda()
    ctx.copy_index_photons(torch.from_numpy(old).cuda(), d_idx, m, -1.0, n, 2, packed, 0)
    ctx.copy_index_photons(torch.from_numpy(new).cuda(), d_idx, m, 1.0, n, 2, packed, 2 * m)
    ctx.sync()
    got = packed.cpu().numpy().reshape(2, 2, m, 8)                 # (old/new, interaction, listed id, field)
    for half, (src, mul) in enumerate(((old, np.float32(-1)), (new, np.float32(1)))):
        want = src.reshape(2, n, 8)[:, idx].copy()
        want[:, :, 3:6] *= mul
        assert np.array_equal(got[half].view(np.uint32), want.view(np.uint32))
    od = (32, 32, 32)
    nvox = od[0] * od[1] * od[2]
    t2i, i2t = cpm.capi.texture_to_index_matrix(od), cpm.capi.index_to_texture_matrix(od)
    radius, scale = 1.3 / 32, 0.37
    want = np.zeros(nvox * channels, np.float64)
    orc.splat(want, channels, t2i, i2t, od, old, None, n, n, 2, radius, scale)
    base = want.copy()
    orc.splat(want, channels, t2i, i2t, od, old, idx, m, n, 2, radius, scale, -1.0)
    orc.splat(want, channels, t2i, i2t, od, new, idx, m, n, 2, radius, scale, 1.0)
    lv = torch.from_numpy(base.astype(np.float32)).cuda()
    ctx.splat_photons(lv, channels, t2i, i2t, od, packed, None, 4 * m, n, 2, radius, scale)
    ctx.sync()
    g = lv.cpu().numpy().astype(np.float64)
    assert np.abs(want - base).max() > 0
    assert np.sqrt(((g - want) ** 2).mean()) / np.sqrt((want ** 2).mean()) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("light", ["oblique", "axis", "point"])
def test_cuda_detector_sparse_grids_bit_exact(cpm, orc, ctx, torch_cuda, synth, light):
    """Importance grids that are mostly zero -- isolated cells, a slab, cells on the grid's faces, a NaN cell, negative
    zeros -- under an oblique, an axis-parallel (segments inside cell faces: the NaN-parameter path of the DDA) and a
    point light, several interactions, both exit-point semantics, a non-cubic grid: keys bit for bit."""
    torch = torch_cuda
    dims = (64, 48, 40)
    vol = synth.volume_u8(dims, 8)
    tf = synth.rasterise_tf(width=512)
    L = {"oblique": lambda: scenes.directional_light(72, (0.3, -0.5, 0.8)), "axis": lambda: scenes.directional_light(72, (0.0, 0.0, 1.0)),
         "point": lambda: scenes.point_light(72)}[light]()
    I = 3
    photons, _, _ = oracle_trace(orc, vol, tf, L, max_interactions=I)
    n = L["n"]
    gd = (8, 6, 5)
    t2i = cpm.capi.texture_to_index_matrix(dims)
    rng = np.random.default_rng(11)
    grids = []
    g = np.zeros(gd[::-1], np.float32); g[2, 3, 4] = 0.7; g[0, 0, 0] = 0.2; g[4, 5, 7] = 1.5
    grids.append(g)
    g = np.zeros(gd[::-1], np.float32); g[:, 2, :] = 0.05
    grids.append(g)
    g = np.full(gd[::-1], -0.0, np.float32); g[3, 1, 6] = np.nan; g[1, 4, 2] = 0.3
    grids.append(g)
    g = (rng.random(gd[::-1]) < 0.06).astype(np.float32) * rng.random(gd[::-1]).astype(np.float32)
    grids.append(g)
    grids.append(np.zeros(gd[::-1], np.float32))
    dph = torch.from_numpy(photons).cuda()
    dls, dis = torch.from_numpy(L["light_samples"]).cuda(), torch.from_numpy(L["isect"]).cuda()
    flagged = 0
    for gi, g in enumerate(grids):
        g = np.ascontiguousarray(g.reshape(-1))
        for fix_exit in (False, True):
            want = np.full(n, 0x7FFFFFFF, np.uint32)
            orc.detect_invalid(g, gd, (8.0, 8.0, 8.0), t2i, photons, 0, L["light_samples"], L["isect"], n, I, n, want, fix_exit=fix_exit)
            keys = torch.from_numpy(np.full(n, 0x7FFFFFFF, np.uint32).view(np.int32)).cuda()
            ctx.detect_invalid(torch.from_numpy(g).cuda(), gd, (8.0, 8.0, 8.0), t2i, dph, 0, dls, dis, n, I, n, keys,
                               flags=cpm.capi.CPM_DETECT_FIX_EXIT if fix_exit else 0)
            ctx.sync()
            got = keys.cpu().numpy().view(np.uint32)
            assert np.array_equal(got, want), (light, gi, fix_exit, int((got != want).sum()))
            flagged += int((want < 0x7FFFFFFF).sum())
    assert flagged > 0


@pytest.mark.gpu
@pytest.mark.parametrize("channels", [1, 4])
def test_cuda_splat_sheared_light_volume(cpm, orc, ctx, torch_cuda, synth, channels):
    """a light volume whose index-to-texture matrix is NOT diagonal (the general row loop of splat_kernel; axis-aligned
    volumes take the shorter one): same estimate as the double-precision oracle"""
    torch = torch_cuda
    c = make_case(orc, synth, cpm, I=2, n_side=64)
    n = c["L"]["n"]
    od = (32, 28, 24)
    A = np.eye(4)
    A[:3, :3] = np.array([[1 / 32, 0.0021, 0.0], [0.0, 1 / 28, 0.0013], [0.0017, 0.0, 1 / 24]])
    A[:3, 3] = A[:3, :3] @ np.array([0.5, 0.5, 0.5]) + np.array([0.01, -0.02, 0.015])
    i2t = [float(np.float32(v)) for v in A.T.reshape(-1)]                       # column-major, as the C ABI takes it
    t2i = [float(np.float32(v)) for v in np.linalg.inv(A).T.reshape(-1)]
    nvox = od[0] * od[1] * od[2]
    radius, scale = 1.4 / 28, 0.21
    want = np.zeros(nvox * channels, np.float64)
    orc.splat(want, channels, t2i, i2t, od, c["photons"], None, n, n, 2, radius, scale)
    lv = torch.zeros(nvox * channels, dtype=torch.float32, device="cuda")
    ctx.splat_photons(lv, channels, t2i, i2t, od, torch.from_numpy(c["photons"]).cuda(), None, n, n, 2, radius, scale)
    ctx.sync()
    got = lv.cpu().numpy().astype(np.float64)
    assert want.sum() > 0
    assert np.sqrt(((got - want) ** 2).mean()) / np.sqrt((want ** 2).mean()) < 1e-5
