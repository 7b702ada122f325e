"""Opacity-bound grid of the tracer (csrc/bound.cu; not in the reference): the value-range grid against a numpy
restatement, the bound against the oracle's own sampler at random points, and -- the real contract -- the bounded
tracer bit for bit against the oracle's plain loop (ppm/cl/transmittance.cl:126-144)."""
import ctypes as C

import numpy as np
import pytest

import scenes
from test_tracer import cuda_trace, oracle_trace


def np_value_range(vol, s):
    """(lo, hi) of the normalised voxels of cell c = voxels [c*cell - 2, c*cell + cell] per axis, clamped"""
    cell = 1 << s
    if vol.dtype == np.uint8:
        f = vol.astype(np.float32) / np.float32(255.0)
    elif vol.dtype == np.uint16:
        f = vol.astype(np.float32) / np.float32(65535.0)
    else:
        f = vol

    def axis_reduce(a, axis, fn):
        n = a.shape[axis]
        nc = (n >> s) + 1
        out = []
        for c in range(nc):
            a0, a1 = max(c * cell - 2, 0), min(c * cell + cell, n - 1)
            out.append(fn(np.take(a, np.arange(a0, a1 + 1), axis=axis), axis=axis, keepdims=True))
        return np.concatenate(out, axis=axis)

    lo, hi = f, f
    for ax in (0, 1, 2):
        lo = axis_reduce(lo, ax, np.min)
        hi = axis_reduce(hi, ax, np.max)
    return lo, hi


def gpu_range(cpm, ctx, torch, vol, s):
    fmt = {np.dtype(np.uint8): cpm.CPM_FMT_U8, np.dtype(np.uint16): cpm.CPM_FMT_U16,
           np.dtype(np.float32): cpm.CPM_FMT_F32}[vol.dtype]
    raw = vol.view(np.int16) if vol.dtype == np.uint16 else vol
    dvol = torch.from_numpy(raw).cuda()
    dims = (vol.shape[2], vol.shape[1], vol.shape[0])
    V = ctx.volume_create(dvol, dims, fmt)
    gd = cpm.capi.bound_grid_dims(dims, s)
    rng = torch.zeros(2 * gd[0] * gd[1] * gd[2], dtype=torch.float32, device="cuda")
    assert ctx.volume_value_range(V, s, rng) == gd
    ctx.sync()
    V.destroy()
    return rng, gd


def test_bound_grid_dims(cpm):
    assert cpm.capi.bound_grid_dims((512, 512, 96), 3) == (65, 65, 13)
    assert cpm.capi.bound_grid_dims((5, 8, 9), 3) == (1, 2, 2)
    assert cpm.capi.bound_grid_dims((5, 8, 9), 0) == (6, 9, 10)
    with pytest.raises(cpm.CpmError):
        cpm.capi.bound_grid_dims((5, 8, 9), 9)


@pytest.mark.gpu
@pytest.mark.parametrize("fmt,dims", [("u8", (48, 40, 36)), ("u8", (37, 21, 19)), ("u16", (40, 24, 17)),
                                      ("f32", (36, 31, 20)), ("f32", (64, 16, 8)), ("u8", (4112, 9, 8)),
                                      ("u16", (2056, 9, 8)), ("f32", (1028, 9, 11))])
def test_cuda_value_range_exact(cpm, ctx, torch_cuda, fmt, dims):
    """(the rows of the three wide volumes have more 16-byte chunks than the streaming kernel has threads)"""
    vol = scenes.make_volume(dims, fmt, 31)
    for s in ((0, 1, 3, 6) if dims[0] < 1000 else (3,)):
        got, gd = gpu_range(cpm, ctx, torch_cuda, vol, s)
        got = got.cpu().numpy().reshape(gd[2], gd[1], gd[0], 2)
        lo, hi = np_value_range(vol, s)
        assert np.array_equal(got[..., 0], lo), (fmt, dims, s)
        assert np.array_equal(got[..., 1], hi), (fmt, dims, s)


@pytest.mark.gpu
@pytest.mark.parametrize("s", [2, 3])
def test_cuda_value_range_flags_nonfinite(cpm, ctx, torch_cuda, s):
    """(s = 3: the streaming kernel for cells of 8 voxels, which finds NaN / inf through the order keys)"""
    vol = scenes.make_volume((32, 16, 16), "f32", 2).copy()
    vol[5, 7, 9] = np.nan      # voxel (x=9, y=7, z=5)
    vol[12, 3, 31] = np.inf
    vol[8, 14, 16] = -np.inf
    vol[2, 2, 6] = -np.nan
    got, gd = gpu_range(cpm, ctx, torch_cuda, vol, s)
    got = got.cpu().numpy().reshape(gd[2], gd[1], gd[0], 2)
    bad = np.isnan(got[..., 0])
    # voxel x belongs to the cells q with q*cell - 2 <= x <= q*cell + cell
    cell = 1 << s
    want = np.zeros_like(bad)
    for (x, y, z) in ((9, 7, 5), (31, 3, 12), (16, 14, 8), (6, 2, 2)):
        qs = [[q for q in range(gd[k]) if q * cell - 2 <= c <= q * cell + cell] for k, c in enumerate((x, y, z))]
        for qz in qs[2]:
            for qy in qs[1]:
                for qx in qs[0]:
                    want[qz, qy, qx] = True
    assert np.array_equal(bad, want) and want.sum() >= 6 and not np.isnan(got[~bad]).any()


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["u8", "f32"])
def test_cuda_opacity_bound_dominates_sampler(cpm, orc, ctx, torch_cuda, synth, fmt):
    """bound[cell of p] >= alpha(TF(volume(p))) as the oracle's sampler computes it, at random points, also
    outside the volume (clamp-to-edge), for a scaled/offset volume and a non-monotone transfer function"""
    torch = torch_cuda
    dims = (40, 33, 27)
    vol = scenes.make_volume(dims, fmt, 17)
    tf = synth.rasterise_tf(width=256)
    tf[:, 3] = (0.5 + 0.5 * np.sin(np.arange(256) * 0.21)) * tf[:, 3] + 0.02 * (np.arange(256) % 7 == 0)
    lib = orc.lib()
    for scale, offset in ((1.0, 0.0), (1.7, -0.05)):
        for s in (0, 2, 3):
            rng, gd = gpu_range(cpm, ctx, torch, vol, s)
            n_cells = gd[0] * gd[1] * gd[2]
            bound = torch.zeros(n_cells, dtype=torch.float32, device="cuda")
            ctx.opacity_bound(rng, n_cells, torch.from_numpy(tf).cuda(), bound, scale=scale, offset=offset)
            ctx.sync()
            b = bound.cpu().numpy().reshape(gd[2], gd[1], gd[0])
            assert np.isfinite(b).all()
            ov = orc.volume(vol, scale=scale, offset=offset)
            pts = (synth.uniform01(100 + s, 3 * 4000).reshape(-1, 3) * 1.2 - 0.1).astype(np.float32)
            worst = -1.0
            for p in pts:
                v = lib.orc_sample_volume(C.byref(ov), C.c_float(p[0]), C.c_float(p[1]), C.c_float(p[2]))
                a = lib.orc_sample_tf_alpha(tf.ctypes.data_as(C.c_void_p), 256, C.c_float(v))
                # the tracer picks cell floor((p * dim + 0.5) / cell) with its own rounding: accept either neighbour
                # when p is within 0.01 voxel of a cell boundary, and require the bound of each to dominate
                u1 = [float(p[k]) * dims[k] + 0.5 for k in range(3)]
                cs = [sorted({int(np.clip(np.floor((u + e) / (1 << s)), 0, gd[k] - 1)) for e in (-0.01, 0.01)})
                      for k, u in enumerate(u1)]
                for cz in cs[2]:
                    for cy in cs[1]:
                        for cx in cs[0]:
                            assert a <= b[cz, cy, cx], (p, a, b[cz, cy, cx])
                worst = max(worst, a / max(b[cs[2][0], cs[1][0], cs[0][0]], 1e-30))
            if s == 0:
                assert worst > 0.5   # the bound stays useful for the smallest cells


def _bounded_trace(cpm, ctx, torch, vol, tf, L, layout, s, clearance=True, **kw):
    """cuda_trace with an opacity bound built on the device for this volume / TF"""
    rng, gd = gpu_range(cpm, ctx, torch, vol, s)
    n_cells = gd[0] * gd[1] * gd[2]
    bound = torch.zeros(n_cells, dtype=torch.float32, device="cuda")
    ctx.opacity_bound(rng, n_cells, torch.from_numpy(tf).cuda(), bound)
    if clearance:
        ctx.opacity_bound_clearance(bound, gd, 6)
    out = cuda_trace(cpm, ctx, torch, vol, tf, L, layout, opacity_bound=bound, bound_cell_log2=s, **kw)
    # the same grid as a point-sampled 3-D texture (cpm_bound_tex: one TEX per test instead of the index arithmetic):
    # same photons, same random states, same counters -- with and without the linear grid beside it
    btex = ctx.bound_texture(gd, bound)
    for lin in (bound, None):
        kw2 = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
        out_t = cuda_trace(cpm, ctx, torch, vol, tf, L, layout, opacity_bound=lin, bound_cell_log2=s, opacity_bound_tex=btex, **kw2)
        assert np.array_equal(out_t[0].view(np.uint32), out[0].view(np.uint32)), "texture bound: photons differ"
        assert np.array_equal(out_t[1], out[1]) and out_t[2:] == out[2:], "texture bound: rng / counters differ"
    btex.close()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["linear", "texture"])
@pytest.mark.parametrize("fmt", ["u8", "u16", "f32"])
def test_cuda_bounded_tracer_bit_exact(cpm, orc, ctx, torch_cuda, synth, fmt, layout):
    lay = cpm.CPM_VOLUME_LINEAR if layout == "linear" else cpm.CPM_VOLUME_TEXTURE
    vol = scenes.make_volume((48, 40, 36), fmt, 21)
    tf = synth.rasterise_tf(width=512)
    L = scenes.directional_light(96, (0.25, -0.4, 0.85))
    for I, flags in ((1, 0), (4, cpm.CPM_TRACE_PROGRESSIVE)):
        want, want_rng, want_tests = oracle_trace(orc, vol, tf, L, max_interactions=I, flags=flags)
        for s, refill in ((0, 0), (2, 0), (3, 0), (6, 0), (3, cpm.capi.CPM_TRACE_LANE_REFILL), (1, cpm.capi.CPM_TRACE_LANE_REFILL)):
            # (refill: the persistent-warp scheduling of trace_refill_kernel -- same photons, same counters)
            got, got_rng, got_tests, fetched = _bounded_trace(cpm, ctx, torch_cuda, vol, tf, L, lay, s,
                                                              max_interactions=I, flags=flags | refill)
            assert got_tests == want_tests
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (I, s, refill)
            assert np.array_equal(got_rng, want_rng)
            assert 0 < fetched < want_tests


@pytest.mark.gpu
def test_cuda_bounded_tracer_variants_bit_exact(cpm, orc, ctx, torch_cuda, synth):
    """point light, clip box, Henyey-Greenstein, NO_SINGLE_SCATTERING, light offsets, index list -- all bounded"""
    vol = scenes.make_volume((64, 64, 64), "u8", 8)
    tf = synth.rasterise_tf(width=1024)
    Lp = scenes.point_light(80)
    Ld = scenes.directional_light(64)
    aabb = ((73 / 512, 7 / 512, 0.0), (1.0, 1.0, 1.0))
    n = Ld["n"]
    idx = np.concatenate([np.arange(3, n, 5), np.arange(n + 1, 2 * n, 7)]).astype(np.uint32)
    cases = [
        dict(L=Lp, max_interactions=3),
        dict(L=Ld, max_interactions=4, aabb=aabb),
        dict(L=Ld, max_interactions=5, phase=1, material=(0.6, 0, 0, 0)),
        dict(L=Ld, max_interactions=3, flags=cpm.CPM_TRACE_NO_SINGLE_SCATTERING),
        dict(L=Ld, max_interactions=2, total_photons=3 * n, photon_offset=n),
        dict(L=Ld, max_interactions=2, total_photons=2 * n, photon_offset=n, recompute=idx,
             photons=np.full((2 * n * 2, 8), -3.0, np.float32), rng=scenes.rng_states(2 * n)),
        # the same index-list re-trace and the clip box / Henyey-Greenstein walk with lane refill
        dict(L=Ld, max_interactions=2, total_photons=2 * n, photon_offset=n, recompute=idx, cuda_flags=cpm.capi.CPM_TRACE_LANE_REFILL,
             photons=np.full((2 * n * 2, 8), -3.0, np.float32), rng=scenes.rng_states(2 * n)),
        dict(L=Ld, max_interactions=5, phase=1, material=(0.6, 0, 0, 0), aabb=aabb, cuda_flags=cpm.capi.CPM_TRACE_LANE_REFILL),
    ]
    for kw in cases:
        L = kw.pop("L")
        cuda_flags = kw.pop("cuda_flags", 0)
        want, want_rng, wt = oracle_trace(orc, vol, tf, L, **{k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in kw.items()})
        if cuda_flags:
            kw["flags"] = kw.get("flags", 0) | cuda_flags
        got, got_rng, gt, fetched = _bounded_trace(cpm, ctx, torch_cuda, vol, tf, L, cpm.CPM_VOLUME_TEXTURE, 3, **kw)
        assert gt == wt, kw
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), kw
        assert np.array_equal(got_rng, want_rng)
        assert fetched < wt


@pytest.mark.gpu
def test_cuda_bounded_tracer_dense_and_empty_media(cpm, orc, ctx, torch_cuda, synth):
    """homogeneous media: opacity 0 (nothing is ever fetched), opacity 1 (every first test is a candidate), and a
    NaN voxel (its cells must fall back to fetching so that the NaN opacity still ends the walk)"""
    L = scenes.directional_light(48, (0.1, 0.2, 0.97))
    vol = np.full((16, 16, 16), 128, np.uint8)
    for alpha in (0.0, 1.0, 0.02):
        tf = synth.dense_tf(alpha, 64)
        want, want_rng, wt = oracle_trace(orc, vol, tf, L, max_interactions=2)
        got, got_rng, gt, fetched = _bounded_trace(cpm, ctx, torch_cuda, vol, tf, L, cpm.CPM_VOLUME_TEXTURE, 2,
                                                   max_interactions=2)
        assert gt == wt and np.array_equal(got.view(np.uint32), want.view(np.uint32)), alpha
        if alpha == 0.0:
            assert fetched == 0
    volf = scenes.make_volume((24, 24, 24), "f32", 9).copy()
    volf[10:13, 11, 12] = np.nan
    tf = synth.rasterise_tf(width=128)
    want, want_rng, wt = oracle_trace(orc, volf, tf, L, max_interactions=2)
    for lay in (cpm.CPM_VOLUME_LINEAR, cpm.CPM_VOLUME_TEXTURE):
        got, got_rng, gt, fetched = _bounded_trace(cpm, ctx, torch_cuda, volf, tf, L, lay, 2, max_interactions=2)
        assert gt == wt
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_cuda_clearance_matches_brute_force(cpm, ctx, torch_cuda, synth):
    """-R for transparent cells whose cube of radius R (cells; outside the grid counts as transparent) is all
    transparent, R capped; cells with a positive bound untouched"""
    torch = torch_cuda
    gd = (19, 14, 11)
    u = synth.uniform01(77, gd[0] * gd[1] * gd[2]).reshape(gd[2], gd[1], gd[0])
    zz, yy, xx = np.meshgrid(np.arange(gd[2]), np.arange(gd[1]), np.arange(gd[0]), indexing="ij")
    solid = (u < 0.004) | (((xx - 9) ** 2 + (yy - 7) ** 2 + (zz - 5) ** 2) < 5)
    b0 = np.where(solid, 0.25 + u, 0.0).astype(np.float32)
    for cap in (1, 3, 8):
        d = torch.from_numpy(b0.copy()).cuda()
        ctx.opacity_bound_clearance(d, gd, cap)
        ctx.sync()
        got = d.cpu().numpy()
        want = b0.copy()
        pad = np.pad(~solid, cap, constant_values=True)
        for z in range(gd[2]):
            for y in range(gd[1]):
                for x in range(gd[0]):
                    if solid[z, y, x]:
                        continue
                    r = 0
                    while r < cap and pad[z + cap - r - 1:z + cap + r + 2, y + cap - r - 1:y + cap + r + 2,
                                          x + cap - r - 1:x + cap + r + 2].all():
                        r += 1
                    want[z, y, x] = -float(r) if r >= 1 else 0.0
        assert np.array_equal(got, want), cap


@pytest.mark.gpu
def test_cuda_bounded_tracer_with_and_without_clearance(cpm, orc, ctx, torch_cuda, synth):
    """sparse medium (most of the volume transparent): clearance on/off, every cell size -- same photons as the oracle"""
    vol = scenes.make_volume((64, 56, 48), "f32", 13).copy()
    vol[vol < 0.35] = 0.0          # large exactly-transparent regions under the workspace TF
    tf = synth.rasterise_tf(width=1024)
    L = scenes.directional_light(80, (0.35, 0.2, 0.9))
    want, want_rng, wt = oracle_trace(orc, vol, tf, L, max_interactions=3)
    for s in (1, 2, 3):
        for clearance in (False, True):
            got, got_rng, gt, fetched = _bounded_trace(cpm, ctx, torch_cuda, vol, tf, L, cpm.CPM_VOLUME_TEXTURE, s,
                                                       clearance=clearance, max_interactions=3)
            assert gt == wt
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (s, clearance)
            assert np.array_equal(got_rng, want_rng)


@pytest.mark.gpu
def test_bound_texture_errors(cpm, ctx, torch_cuda, synth):
    """cpm_bound_tex_*: a texture made for another grid is refused by the tracer (CPM_E_INVALID, message names the reason),
    non-positive dims are refused at creation"""
    torch = torch_cuda
    with pytest.raises(cpm.capi.CpmError):
        ctx.bound_texture((0, 4, 4))
    vol = scenes.make_volume((32, 32, 32), "u8", 3)
    tf = synth.rasterise_tf(width=64)
    L = scenes.directional_light(16)
    gd = cpm.capi.bound_grid_dims((32, 32, 32), 3)
    wrong = ctx.bound_texture((gd[0] + 1, gd[1], gd[2]), torch.zeros((gd[0] + 1) * gd[1] * gd[2], dtype=torch.float32, device="cuda"))
    with pytest.raises(cpm.capi.CpmError) as e:
        cuda_trace(cpm, ctx, torch, vol, tf, L, cpm.CPM_VOLUME_LINEAR, opacity_bound_tex=wrong, bound_cell_log2=3)
    assert "another grid" in str(e.value)
    wrong.close()
