"""Emission: sample generator, directional / point light sampling, proxy-mesh intersection."""
import numpy as np
import pytest

import scenes


def test_uniform2d_quirk_y_not_floored(orc):
    """isc/cl/uniformsamplegenerator2d.cl:46-47: y uses id/dims.x without floor"""
    s = orc.sample_uniform2d(4.0, 4.0, 16)
    assert s[5, 0] == np.float32((0.5 + 1.0) / 4.0)
    assert s[5, 1] == np.float32((np.float32(0.5) + np.float32(5) / np.float32(4)) / np.float32(4))
    assert (s[:, 2] == 0).all() and (s[:, 3] == 1).all()


def test_light_plane_fit_covers_cube(orc, synth):
    """the fitted rectangle's projection contains every proxy vertex (lcl/orientedboundingbox2d.cpp)"""
    for d in [(0, 0, 1), (0.3, -0.5, 0.8), (-1, 0.2, 0.1), (0.577, 0.577, 0.577)]:
        L = scenes.directional_light(4, d)
        u, v, o = L["u"].astype(np.float64), L["v"].astype(np.float64), L["origin"].astype(np.float64)
        assert abs(np.dot(u, v)) < 1e-5 and abs(np.dot(u, L["dir"])) < 1e-5
        for p in synth.CUBE_VERTICES.astype(np.float64):
            q = p - o
            a, b = np.dot(q, u) / np.dot(u, u), np.dot(q, v) / np.dot(v, v)
            assert -1e-4 <= a <= 1 + 1e-4 and -1e-4 <= b <= 1 + 1e-4


def test_directional_rays_hit_or_miss_consistently(orc):
    L = scenes.directional_light(64)
    t = L["isect"]
    hit = t[:, 0] < t[:, 1]
    assert 0.2 < hit.mean() < 1.0
    assert (t[~hit] == np.array([0.0, -1.0], np.float32)).all()
    # entry and exit points of hits lie on the unit cube surface
    ls = L["light_samples"]
    p0 = ls[hit, 0:3] + t[hit, 0:1] * L["dir"]
    p1 = ls[hit, 0:3] + t[hit, 1:2] * L["dir"]
    for p in (p0, p1):
        on_face = np.minimum(np.abs(p), np.abs(p - 1)).min(axis=1)
        assert on_face.max() < 1e-4
        assert p.min() > -1e-4 and p.max() < 1 + 1e-4


@pytest.mark.gpu
def test_cuda_emission_bit_exact(cpm, orc, ctx, torch_cuda, synth):
    torch = torch_cuda
    for n_side, d in [(64, (0.3, -0.5, 0.8)), (37, (0.0, 0.0, 1.0)), (50, (-0.7, 0.1, -0.2))]:
        L = scenes.directional_light(n_side, d)
        n = L["n"]
        s = torch.empty(n * 4, dtype=torch.float32, device="cuda")
        ctx.sample_uniform2d(float(n_side), float(n_side), n, s)
        ls = torch.empty(n * 8, dtype=torch.float32, device="cuda")
        ctx.light_sample_directional(s, L["radiance"], L["dir"], L["origin"], L["u"], L["v"], L["area"], n, ls)
        verts = torch.from_numpy(synth.CUBE_VERTICES).cuda()
        idx = torch.from_numpy(synth.CUBE_INDICES).cuda()
        it = torch.empty(n * 2, dtype=torch.float32, device="cuda")
        ctx.light_mesh_intersect(verts, idx, idx.numel(), ls, n, it)
        ctx.sync()
        assert np.array_equal(s.cpu().numpy().view(np.uint32), L["samples"].reshape(-1).view(np.uint32))
        assert np.array_equal(ls.cpu().numpy().view(np.uint32), L["light_samples"].reshape(-1).view(np.uint32))
        assert np.array_equal(it.cpu().numpy().view(np.uint32), L["isect"].reshape(-1).view(np.uint32))


@pytest.mark.gpu
def test_cuda_point_light_bit_exact(cpm, orc, ctx, torch_cuda, synth):
    torch = torch_cuda
    L = scenes.point_light(48)
    n = L["n"]
    s = torch.from_numpy(L["samples"]).cuda()
    ls = torch.empty(n * 8, dtype=torch.float32, device="cuda")
    ctx.light_sample_point(s, L["radiance"], L["position"], n, ls)
    verts = torch.from_numpy(synth.CUBE_VERTICES).cuda()
    idx = torch.from_numpy(synth.CUBE_INDICES).cuda()
    it = torch.empty(n * 2, dtype=torch.float32, device="cuda")
    ctx.light_mesh_intersect(verts, idx, idx.numel(), ls, n, it)
    ctx.sync()
    assert np.array_equal(ls.cpu().numpy().view(np.uint32), L["light_samples"].reshape(-1).view(np.uint32))
    assert np.array_equal(it.cpu().numpy().view(np.uint32), L["isect"].reshape(-1).view(np.uint32))


@pytest.mark.gpu
def test_empty_inputs_are_noops(ctx, torch_cuda):
    ctx.sample_uniform2d(4.0, 4.0, 0, None)
    ctx.sync()
