"""Importance-map-driven emission (north-star 2): the view/light importance image
(isc/cl/minmaxuniformgrid3dimportance.cl:336-378) and the importance-driven 2-D sample generator behind the
SampleGenerator2DCL interface (new; parity unpinned)."""
import numpy as np
import pytest

import scenes


def _entry_exit(L, n_side):
    """entry / exit points (texture space) of the light's own sample rays: the light is the 'camera'"""
    ls, it = L["light_samples"], L["isect"]
    th, ph = ls[:, 6].astype(np.float64), ls[:, 7].astype(np.float64)
    d = np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], axis=1)
    hit = it[:, 0] < it[:, 1]
    e = ls[:, :3] + d * it[:, 0:1]
    x = ls[:, :3] + d * it[:, 1:2]
    e[~hit] = 0.0
    x[~hit] = 0.0       # x1 == x2: importance 0 (the kernel's any(x1 != x2) test)
    pad = np.zeros((ls.shape[0], 1))
    entry = np.concatenate([e, pad], axis=1).astype(np.float32).reshape(n_side, n_side, 4)
    exit_ = np.concatenate([x, pad], axis=1).astype(np.float32).reshape(n_side, n_side, 4)
    return np.ascontiguousarray(entry), np.ascontiguousarray(exit_), hit


def _scene(orc, cpm, synth, ns=64, dims=(64, 64, 64)):
    vol = synth.volume_u8(dims, 6)
    mm = orc.volume_minmax(vol, 8)
    L = scenes.directional_light(ns, (0.3, -0.5, 0.8))
    entry, exit_, hit = _entry_exit(L, ns)
    gd = (mm.shape[2], mm.shape[1], mm.shape[0])
    t2i, i2t = cpm.capi.texture_to_index_matrix(dims), cpm.capi.index_to_texture_matrix(dims)
    return vol, mm, L, entry, exit_, hit, gd, t2i, i2t


# ------------------------------------------------------------------------------- CPU ---------
def test_oracle_view_importance_bounds(orc, cpm, synth):
    vol, mm, L, entry, exit_, hit, gd, t2i, i2t = _scene(orc, cpm, synth)
    # every brick overlaps [0, 1]: importance == chord length inside the grid (texture units), 0 for misses
    full = orc.view_importance(mm, gd, (8, 8, 8), t2i, i2t, entry, exit_, 0.0, 1.0)
    chord = np.linalg.norm(exit_[..., :3].astype(np.float64) - entry[..., :3].astype(np.float64), axis=2)
    h = hit.reshape(full.shape)
    assert np.all(full[~h] == 0)
    assert np.allclose(full[h], chord[h], rtol=2e-3, atol=2e-3)
    # a narrower TF range can only shorten the visible length
    part = orc.view_importance(mm, gd, (8, 8, 8), t2i, i2t, entry, exit_, 0.3, 1.0)
    assert np.all(part <= full + 1e-6) and part.sum() < full.sum() and part.max() > 0


def test_oracle_importance_sampler_is_a_density(orc, synth):
    w, h, ns = 32, 24, 128
    imp = (synth.uniform01(3, w * h).reshape(h, w) ** 4).astype(np.float32)
    imp[:, :8] = 0.0
    uni = orc.sample_uniform2d(ns, ns, ns * ns)
    out = orc.sample_importance2d(imp, 0.01, uni)
    assert out[:, :2].min() >= 0.0 and out[:, :2].max() < 1.0
    # unbiasedness of the emission: mean(1 / pdf) estimates the integral of 1 over the unit square
    assert abs((1.0 / out[:, 3].astype(np.float64)).mean() - 1.0) < 3e-2
    # samples land in cells in proportion to the density
    xi = np.minimum((out[:, 0] * w).astype(int), w - 1)
    yi = np.minimum((out[:, 1] * h).astype(int), h - 1)
    hist = np.zeros((h, w))
    np.add.at(hist, (yi, xi), 1)
    f = imp.astype(np.float64) + 0.01
    want = f / f.sum() * ns * ns
    assert np.abs(hist - want).max() < 0.02 * ns * ns / 10 + 3 * np.sqrt(want.max())
    # uniform importance leaves the stratified samples where they are, pdf 1
    flat = orc.sample_importance2d(np.ones((h, w), np.float32), 0.0, uni)
    # (the uniform generator's un-floored y reaches slightly past 1, isc/cl/uniformsamplegenerator2d.cl:46-47;
    # the warp clamps its inputs to [0, 1))
    clamped = np.minimum(uni[:, :2], np.float32(0.99999994))
    assert np.allclose(flat[:, :2], clamped, atol=2e-6) and np.allclose(flat[:, 3], 1.0, rtol=1e-5)


# ------------------------------------------------------------------------------- GPU ---------
@pytest.mark.gpu
def test_cuda_view_importance_bit_exact(cpm, orc, synth, ctx, torch_cuda):
    torch = torch_cuda
    vol, mm, L, entry, exit_, hit, gd, t2i, i2t = _scene(orc, cpm, synth)
    ns = entry.shape[0]
    for lo, hi in ((0.0, 1.0), (0.3, 1.0), (0.0737, 0.6)):
        want = orc.view_importance(mm, gd, (8, 8, 8), t2i, i2t, entry, exit_, lo, hi)
        out = torch.zeros(ns * ns, dtype=torch.float32, device="cuda")
        ctx.view_importance(torch.from_numpy(mm.reshape(-1).view(np.int16)).cuda(), gd, (8, 8, 8), t2i, i2t,
                            torch.from_numpy(entry.reshape(-1)).cuda(), torch.from_numpy(exit_.reshape(-1)).cuda(), ns, ns,
                            lo, hi, out)
        ctx.sync()
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want.reshape(-1).view(np.uint32)), (lo, hi)


@pytest.mark.gpu
def test_cuda_importance_driven_emission(cpm, orc, synth, ctx, torch_cuda):
    """light importance image -> warped samples -> light samples: bit-exact vs the oracle, more photons where the
    volume is visible through the TF, and the total emitted power is preserved (power carries 1 / pdf)"""
    torch = torch_cuda
    vol, mm, L, entry, exit_, hit, gd, t2i, i2t = _scene(orc, cpm, synth)
    ns = entry.shape[0]
    imp = orc.view_importance(mm, gd, (8, 8, 8), t2i, i2t, entry, exit_, 0.0737, 1.0)
    n_side = 192
    n = n_side * n_side
    uni = orc.sample_uniform2d(n_side, n_side, n)
    floor_value = 0.05 * float(imp.mean())
    want = orc.sample_importance2d(imp, floor_value, uni)
    d_imp, d_uni = torch.from_numpy(imp.reshape(-1)).cuda(), torch.from_numpy(uni.reshape(-1)).cuda()
    scratch = torch.zeros(cpm.capi.sample_importance2d_scratch_floats(ns, ns), dtype=torch.float32, device="cuda")
    out = torch.zeros(n * 4, dtype=torch.float32, device="cuda")
    ctx.sample_importance2d(d_imp, ns, ns, floor_value, d_uni, n, scratch, out)
    ctx.sync()
    got = out.cpu().numpy().reshape(n, 4)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # emission through the unchanged light sampler: power = radiance * area / pdf
    ls_w = torch.zeros(n * 8, dtype=torch.float32, device="cuda")
    ctx.light_sample_directional(out, L["radiance"], L["dir"], L["origin"], L["u"], L["v"], L["area"], n, ls_w)
    ls_u = torch.zeros(n * 8, dtype=torch.float32, device="cuda")
    ctx.light_sample_directional(d_uni, L["radiance"], L["dir"], L["origin"], L["u"], L["v"], L["area"], n, ls_u)
    ctx.sync()
    pw, pu = ls_w.cpu().numpy().reshape(n, 8)[:, 3].astype(np.float64), ls_u.cpu().numpy().reshape(n, 8)[:, 3].astype(np.float64)
    # (a stratified estimate of the integral of 1 over the square: 0.988 at 192^2 samples, 0.999 at 256^2)
    assert abs(pw.sum() / pu.sum() - 1.0) < 0.03
    # the share of samples whose ray sees TF-visible material rises
    xi = np.minimum((got[:, 0] * ns).astype(int), ns - 1)
    yi = np.minimum((got[:, 1] * ns).astype(int), ns - 1)
    frac_imp = (imp[yi, xi] > 0).mean()
    assert frac_imp > (imp > 0).mean() + 0.1
