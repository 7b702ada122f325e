"""Final image from the light volume (csrc/raycast.cu): the LightingRaycaster step of the workspace network
(ws:1178-1271).  Inviwo's LightingRaycaster is outside the reference tree, so the checker is the oracle's own
restatement ("parity unpinned"); the closed-form test pins the compositing model."""
import numpy as np
import pytest

import scenes
from test_tracer import oracle_trace


def _params(orc, cpm, **kw):
    return cpm.capi.make_gather_params(cls=orc.GatherParams, **kw)


def test_oracle_raycast_homogeneous_closed_form(orc, cpm, synth):
    """constant opacity a, constant TF colour, constant light L: pixel = colour * L * (1 - T), T = exp(-a sigma step n)"""
    vol = np.zeros((16, 16, 16), np.uint8)
    a, sig, step, Lc = 0.2, 10.0, 1.0 / 64, 0.7
    tf = synth.dense_tf(a, 64)
    tf[:, :3] = (0.9, 0.5, 0.25)
    lv = np.full(8 * 8 * 8, Lc, np.float32)
    P = _params(orc, cpm, width=16, height=12, eye=(0.5, 0.5, -1.5), look_at=(0.5, 0.5, 0.5), fov_deg=30.0, step=step,
                sigma_scale=sig)
    img = orc.raycast_light_volume(orc.volume(vol), tf, P, lv, (8, 8, 8), 1)
    hit = img[..., 3] > 0
    assert hit.mean() > 0.2
    for ch, col in enumerate((0.9, 0.5, 0.25)):
        assert np.allclose(img[..., ch][hit], col * Lc * img[..., 3][hit], rtol=2e-5)
    assert abs(img[6, 8, 3] - (1 - np.exp(-a * sig * step * 64))) < 0.02
    assert np.all(img[~hit] == 0)


def _scene(orc, cpm, synth, channels):
    dims = (48, 48, 48)
    vol = synth.volume_u8(dims, 8)
    tf = synth.rasterise_tf(width=1024)
    L = scenes.directional_light(96, (0.3, -0.5, 0.8), radiance=(1.0, 0.9, 0.8))
    ph, _, _ = oracle_trace(orc, vol, tf, L, max_interactions=2, step_size=1.0 / dims[0])
    n = L["n"]
    lvd = (24, 24, 24)
    acc = np.zeros(lvd[0] * lvd[1] * lvd[2] * channels, np.float64)
    t2i, i2t = cpm.capi.texture_to_index_matrix(lvd), cpm.capi.index_to_texture_matrix(lvd)
    orc.splat(acc, channels, t2i, i2t, lvd, ph, np.arange(n, dtype=np.uint32), n, n, 2, 2.0 / 48, 50.0)
    return dims, vol, tf, lvd, acc.astype(np.float32)


@pytest.mark.gpu
@pytest.mark.parametrize("channels", [1, 4])
@pytest.mark.parametrize("layout", ["linear", "texture"])
def test_cuda_raycast_matches_oracle(cpm, orc, synth, ctx, torch_cuda, layout, channels):
    torch = torch_cuda
    dims, vol, tf, lvd, lv = _scene(orc, cpm, synth, channels)
    kw = dict(width=96, height=64, eye=(1.6, 1.3, -1.2), look_at=(0.5, 0.5, 0.5), fov_deg=35.0, step=0.5 / 48,
              sigma_scale=150.0)
    want = orc.raycast_light_volume(orc.volume(vol), tf, _params(orc, cpm, **kw), lv, lvd, channels)
    assert want[..., :3].max() > 0 and (want[..., 3] > 0).mean() > 0.2
    dvol = torch.from_numpy(vol).cuda()
    V = ctx.volume_create(dvol, dims, cpm.CPM_FMT_U8,
                          layout=cpm.CPM_VOLUME_LINEAR if layout == "linear" else cpm.CPM_VOLUME_TEXTURE)
    dtf, dlv = torch.from_numpy(tf.reshape(-1)).cuda(), torch.from_numpy(lv).cuda()
    P = cpm.capi.make_gather_params(**kw)
    img = torch.zeros(96 * 64 * 4, dtype=torch.float32, device="cuda")
    ctx.raycast_light_volume(V, dtf, P, dlv, lvd, channels, img)
    ctx.sync()
    got = img.cpu().numpy().reshape(64, 96, 4)
    # same samples, same operation order: bit-identical is expected; the stated criterion is PSNR > 100 dB
    mse = ((got.astype(np.float64) - want.astype(np.float64)) ** 2).mean()
    assert 10 * np.log10(float(want.max()) ** 2 / max(mse, 1e-300)) > 100.0
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # with the opacity bound transparent cells are stepped over: exactly the same image
    Vl = V if layout == "linear" else ctx.volume_create(dvol, dims, cpm.CPM_FMT_U8)
    for s in (2, 3):
        gd = cpm.capi.bound_grid_dims(dims, s)
        nc = gd[0] * gd[1] * gd[2]
        rng = torch.zeros(2 * nc, dtype=torch.float32, device="cuda")
        ctx.volume_value_range(Vl, s, rng)
        bound = torch.zeros(nc, dtype=torch.float32, device="cuda")
        ctx.opacity_bound(rng, nc, dtf, bound)
        if s == 3:
            ctx.opacity_bound_clearance(bound, gd, 4)
        Pb = cpm.capi.make_gather_params(opacity_bound=bound, bound_cell_log2=s, **kw)
        img2 = torch.zeros(96 * 64 * 4, dtype=torch.float32, device="cuda")
        ctx.raycast_light_volume(V, dtf, Pb, dlv, lvd, channels, img2)
        ctx.sync()
        assert np.array_equal(img2.cpu().numpy().reshape(64, 96, 4).view(np.uint32), got.view(np.uint32)), s
    if Vl is not V:
        Vl.destroy()
    V.destroy()


@pytest.mark.gpu
def test_cuda_raycast_argument_errors(cpm, ctx, torch_cuda):
    torch = torch_cuda
    v = torch.zeros(8, dtype=torch.uint8, device="cuda")
    V = ctx.volume_create(v, (2, 2, 2), cpm.CPM_FMT_U8)
    tf = torch.zeros(16, dtype=torch.float32, device="cuda")
    buf = torch.zeros(64, dtype=torch.float32, device="cuda")
    P = cpm.capi.make_gather_params(2, 2, (0.5, 0.5, -2), (0.5, 0.5, 0.5))
    with pytest.raises(cpm.CpmError) as e:
        ctx.raycast_light_volume(V, tf, P, buf, (2, 2, 2), 3, buf)
    assert e.value.code == -1
    V.destroy()
