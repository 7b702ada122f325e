"""Oracle pinned against the reference's OWN kernels.

oracle/_ref/libcl_ref.so is every OpenCL kernel file of the reference on the hot path -- ppm/cl/{photontracer,
transmittance,photon,photonrecomputationdetector,photonstolightvolume,densityestimationkernel,threshold,indextobuffer,
hashlightsample}.cl, rng/cl/{random,skip_mwc}.cl, lcl/cl/{directionallightsampler,datastructures/lightsample,
intersection/lightsamplemeshintersection}.cl, isc/cl/{uniformsamplegenerator2d,minmaxuniformgrid3dimportance,light/light}.cl,
ugc/cl/{uniformgrid/uniformgrid,uniformgrid/volumeminmax,buffermixer}.cl -- compiled for the host where the files lie
(oracle/Makefile `refcl`: OpenCL C on C++ through oracle/ref_shim/cl/clc.h, strict IEEE evaluation, stand-ins for the 13
Inviwo headers the reference does not ship).  tests/golden/ref_kernels.npz holds its outputs on the seeded cases of
tests/ref_cases.py (tools/make_golden.py kernels).

  test_oracle_equals_reference_golden    the oracle reproduces every golden array: bit for bit for all kernels whose result is
                                         order-free, fp32-sum tolerance for the splat's atomic accumulation
  test_live_*                            the same against the library itself on other inputs (skipped where oracle/_ref was
                                         not built, e.g. a checkout without /root/reference)
  test_golden_matches_library            the committed fixture is what the generator writes
The CUDA kernels are compared bit for bit with the oracle in the -m gpu tests, which closes the chain
reference kernels == oracle == sm_100a kernels."""
import ctypes as C

import numpy as np
import pytest

import ref_cases as rc
from conftest import GOLDEN
from oracle import frame, orc

FLT_MAX = np.float32(3.4028234663852886e38)


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32) if a.dtype == np.float32 else np.ascontiguousarray(a)


def oracle_outputs():
    """the oracle on the cases of ref_cases.ref_outputs, same keys"""
    S = rc.scene()
    L = S["L"]
    N, NS = rc.N, rc.NS
    out = {}
    out["uniform2d"] = orc.sample_uniform2d(float(NS), float(NS), N)
    out["uniform2d_ragged"] = orc.sample_uniform2d(33.0, 31.0, 1000)
    out["light_samples"] = L["light_samples"]
    out["isect"] = L["isect"]
    for name, I, flags, phase, material, entry, aabb in rc.TRACE_VARIANTS:
        out["trace_" + name], out["rng_" + name] = rc.oracle_trace(S, I, flags, phase, material, aabb)
    out["trace_recompute_I2"], _ = rc.oracle_trace(S, 2, 0, 0, (0, 0, 0, 0), ((0, 0, 0), (1, 1, 1)), rc.recompute_ids())
    t2i = frame.texture_to_index(rc.DIMS)
    for name in ("plain_I1", "plain_I3", "hg_I3"):
        I = int(name[-1])
        keys = np.full(N, 0x7FFFFFFF, np.uint32)
        orc.detect_invalid(S["grid"], S["gd"], (rc.REGION,) * 3, t2i, out["trace_" + name], 0, L["light_samples"], L["isect"], N, I,
                           N, keys)
        out["detect_" + name] = keys
    keys = np.full(N, 0x7FFFFFFF, np.uint32)
    orc.detect_invalid(S["grid"], S["gd"], (rc.REGION,) * 3, t2i, out["trace_plain_I1"], 0, L["light_samples"], L["isect"], N, 1, N,
                       keys, equal_importance=True, percentage=25, iteration=3)
    out["detect_equal_importance"] = keys
    out["threshold"] = (out["detect_plain_I1"] < 0x7FFFFFFF).astype(np.uint32)     # ppm/cl/threshold.cl:39
    out["iota"] = np.arange(N, dtype=np.uint32)                                   # ppm/cl/indextobuffer.cl:39
    out["minmax_u8"] = orc.volume_minmax(S["vol"], rc.REGION)
    out["minmax_f32"] = orc.volume_minmax(S["volf"], rc.REGION)
    mm, prev = orc.volume_minmax(S["vol"], rc.REGION), orc.volume_minmax(S["vol_prev"], rc.REGION)
    diff = orc.volume_diff_bricks(S["vol_prev"], S["vol"], rc.REGION, 1.0, 0.0, 255.0)
    pos, col = frame.tf_point_lists(rc.synth.WS_TF_POINTS)
    for k, w in enumerate(rc.classify_weights()):
        out[f"classify_static_{k}"] = orc.classify_importance(mm, pos, col, w, True)
        out[f"classify_timevarying_{k}"] = orc.classify_importance(mm, pos, col, w, False, prev=prev, diff=diff.reshape(-1))
    ids = (rc.synth.splitmix64(3, 700) % np.uint64(N + 50)).astype(np.uint32)
    hb = np.zeros(700, np.uint32)
    orc.hash_light_samples(L["light_samples"], L["isect"], N, ids, (8, 8, 8), (8, 8, 8), hb)
    out["hash"] = hb
    t2, i2 = frame.texture_to_index(rc.LV), frame.index_to_texture(rc.LV)
    radius, scale = float(np.float32(1.7 / 24)), 1e-3
    ph = out["trace_plain_I3"]
    nv = rc.LV[0] * rc.LV[1] * rc.LV[2]
    stored = np.where(ph[:, 0] != FLT_MAX)[0][:6]
    for k, g in enumerate(stored):
        v = np.zeros(nv, np.float64)
        orc.splat(v, 1, t2, i2, rc.LV, np.ascontiguousarray(ph[g:g + 1]), None, 1, 1, 1, radius, scale)
        out[f"splat_single_{k}"] = v.astype(np.float32)      # one contribution per voxel: the float64 sum is that float
    v = np.zeros(nv, np.float64)
    orc.splat(v, 1, t2, i2, rc.LV, ph, None, N * 3, N, 3, radius, scale)
    out["splat_all"] = v
    sel = np.arange(0, N, 3, dtype=np.uint32)
    v4 = np.zeros(nv * 4, np.float64)
    orc.splat(v4, 4, t2, i2, rc.LV, ph, sel, int(sel.size), N, 3, radius, scale, -1.0)
    out["splat_selected_rgba_minus"] = v4
    return out


SUMS = ("splat_all", "splat_selected_rgba_minus")       # fp32 atomic sums: order-dependent in the reference itself


def _compare(got, want, label):
    for key in sorted(want.keys()):
        if key.startswith("mix_"):
            continue                                     # the oracle has no mixer; the CUDA kernel is checked against this array (test_grid.py)
        g, w = got[key], want[key]
        if key in SUMS:
            rel = np.sqrt(((g.astype(np.float64) - w.astype(np.float64)) ** 2).mean()) / np.sqrt((w.astype(np.float64) ** 2).mean())
            assert rel <= 1e-6, (label, key, rel)
            assert ((g != 0) == (w != 0)).all(), (label, key)
        else:
            assert g.shape == w.shape and g.dtype == w.dtype, (label, key, g.shape, w.shape, g.dtype, w.dtype)
            assert np.array_equal(_bits(g), _bits(w)), (label, key, float((_bits(g) != _bits(w)).mean()))


def test_oracle_equals_reference_golden():
    want = dict(np.load(GOLDEN / "ref_kernels.npz"))
    got = oracle_outputs()
    assert len(want) >= 39
    # the cases exercise what they claim to
    assert (want["trace_plain_I3"][:, 0] != FLT_MAX).sum() > 300
    assert ((want["detect_plain_I3"] != 0x7FFFFFFF).sum() > 50) and ((want["detect_plain_I1"] != 0x7FFFFFFF).sum() > 50)
    assert want["classify_timevarying_1"].max() > 0
    _compare(got, want, "golden")


def test_detector_exit_quirks_are_the_reference_s():
    """Two documented quirks, pinned by the reference's own kernel (golden) and reproduced by the oracle:
    interaction 0 of an escaped photon uses exit = tEnd * direction (no origin, :128); an escape after k > 0 interactions
    adds to a FLT_MAX position (:137), which makes the photon's importance NaN -> key untouched."""
    want = dict(np.load(GOLDEN / "ref_kernels.npz"))
    S = rc.scene()
    ph = want["trace_plain_I3"].reshape(3, rc.N, 8)
    # photons that scattered at least once and then escaped with the "not absorbed" marker
    escaped_late = (ph[0, :, 0] != FLT_MAX) & (ph[1, :, 0] == FLT_MAX) & (ph[1, :, 3] != FLT_MAX)
    assert escaped_late.sum() > 20
    assert (want["detect_plain_I3"][escaped_late] == 0x7FFFFFFF).all()
    # with the repair (CPM_DETECT_FIX_EXIT) the oracle flags some of them: the quirk is observable
    keys = np.full(rc.N, 0x7FFFFFFF, np.uint32)
    orc.detect_invalid(S["grid"], S["gd"], (rc.REGION,) * 3, frame.texture_to_index(rc.DIMS), want["trace_plain_I3"], 0,
                       S["L"]["light_samples"], S["L"]["isect"], rc.N, 3, rc.N, keys, fix_exit=True)
    assert (keys[escaped_late] != 0x7FFFFFFF).any()


def _ref():
    ref = orc.ref_lib("cl_ref")
    if ref is None:
        pytest.skip("oracle/_ref/libcl_ref.so not built (reference tree absent)")
    return ref


def test_golden_matches_library():
    ref = _ref()
    want = dict(np.load(GOLDEN / "ref_kernels.npz"))
    got = rc.ref_outputs(ref)
    assert set(got) == set(want)
    for k in want:
        assert np.array_equal(_bits(got[k]), _bits(want[k])), k


@pytest.mark.parametrize("seed,fmt", [(11, "u8"), (12, "u16"), (13, "f32")])
def test_live_tracer_and_detector_on_other_volumes(seed, fmt):
    """fresh volumes of every voxel format, another light direction, 4 interactions, index list: photons, RNG states and
    detector keys of the oracle == the reference kernels', bit for bit"""
    ref = _ref()
    P, N = rc.P, rc.N
    dims = (40, 33, 29)
    vol = {"u8": rc.synth.volume_u8, "u16": rc.synth.volume_u16}.get(fmt, lambda d, s: rc.synth.volume_f32(d, s, 0.25))(dims, seed)
    tf = frame.rasterise_tf(rc.synth.WS_TF_POINTS, 256)
    L = frame.directional_light(rc.NS, (-0.6, 0.2, 0.77), radiance=(0.7, 1.0, 0.4))
    I = 4
    for flags, entry, rec in ((0, "ref_trace_photons", None), (0, "ref_trace_photons_recompute", rc.recompute_ids()),
                              (1, "ref_trace_photons_progressive", None), (2, "ref_trace_photons_nss", None)):
        p = orc.trace_params(n_light_samples=N, max_interactions=I, step_size=1.0 / 40, flags=flags)
        rng0 = orc.rng_seed_streams(orc.rng_host_base_offsets(0, N))
        ph_o, ph_r = np.full((N * I, 8), 3.0, np.float32), np.full((N * I, 8), 3.0, np.float32)
        r_o, r_r = rng0.copy(), rng0.copy()
        kw = {} if rec is None else dict(recompute=rec, n_recompute=int(rec.size))
        orc.trace_photons(orc.volume(vol), tf, p, L["light_samples"], L["isect"], ph_o, r_o, **kw)
        V = orc.volume(vol)
        getattr(ref, entry)(C.byref(V), P(tf), 256, C.byref(p), P(L["light_samples"]), P(L["isect"]), P(rec),
                            0 if rec is None else int(rec.size), P(ph_r), P(r_r))
        assert np.array_equal(_bits(ph_o), _bits(ph_r)), (fmt, entry)
        assert np.array_equal(r_o, r_r), (fmt, entry)
        if rec is None and flags == 0:
            gd = tuple(-(-d // 8) for d in dims)
            rs = np.random.default_rng(seed)
            grid = (rs.random(gd[0] * gd[1] * gd[2]) * (rs.random(gd[0] * gd[1] * gd[2]) < 0.4)).astype(np.float32)
            k_o, k_r = np.full(N, 0x7FFFFFFF, np.uint32), np.full(N, 0x7FFFFFFF, np.uint32)
            t2i = frame.texture_to_index(dims)
            orc.detect_invalid(grid, gd, (8, 8, 8), t2i, ph_o, 0, L["light_samples"], L["isect"], N, I, N, k_o)
            ref.ref_detect_invalid(P(grid), rc.I3(gd), rc.F3((8, 8, 8)), rc.F16(t2i), P(ph_o), 0, P(L["light_samples"]), P(L["isect"]),
                                   N, I, N, P(k_r), 0, 100, 0)
            assert np.array_equal(k_o, k_r), fmt
            assert (k_o != 0x7FFFFFFF).sum() > 20


def test_live_dda_on_random_segments():
    """the DDA alone (ugc/cl/uniformgrid/uniformgrid.cl:38-69,147-197 through ppm/cl/photonrecomputationdetector.cl:55-90):
    random segments incl. axis-aligned ones, points on cell faces and outside the grid; the oracle's importance per segment
    equals the reference's bit for bit (the oracle is driven through its detector with one synthetic photon per segment)"""
    ref = _ref()
    rs = np.random.default_rng(5)
    gd, cell = (7, 5, 6), (8.0, 8.0, 8.0)
    grid = rs.random(gd[0] * gd[1] * gd[2]).astype(np.float32)
    n = 4000
    x1 = rs.uniform(-4, 60, (n, 3)).astype(np.float32)
    x2 = rs.uniform(-4, 60, (n, 3)).astype(np.float32)
    x2[::7, 0] = x1[::7, 0]                       # axis-parallel
    x2[::11, 1:] = x1[::11, 1:]
    x1[::5] = np.round(x1[::5] / 8) * 8           # on cell faces
    want = np.array([ref.ref_uniform_grid_importance(rc.F3(a), rc.F3(b), rc.F3(cell), rc.P(grid), rc.I3(gd)) for a, b in zip(x1, x2)])
    # the oracle's DDA is static: reach it through orc_detect_invalid with an identity texture-to-index matrix shifted by
    # -0.5 (the kernel adds 0.5), a light sample whose entry point is x1 and a stored photon at x2
    ident = np.zeros(16, np.float32)
    ident[0] = ident[5] = ident[10] = ident[15] = 1.0
    ident[12:15] = -0.5
    ls = np.zeros((n, 8), np.float32)
    ls[:, 0:3] = x1
    ls[:, 6], ls[:, 7] = 0.0, 0.0                 # direction (0, 0, 1); tStart = 0 => entry = origin + 0 * dir = x1 exactly
    isect = np.zeros((n, 2), np.float32)
    isect[:, 1] = 1.0
    ph = np.zeros((n, 8), np.float32)
    ph[:, 0:3] = x2
    keys = np.full(n, 0x7FFFFFFF, np.uint32)
    orc.detect_invalid(grid, gd, cell, ident, ph, 0, ls, isect, n, 1, n, keys)
    ref.ref_uniform_grid_importance.restype = C.c_float
    want = np.array([ref.ref_uniform_grid_importance(rc.F3(a), rc.F3(b), rc.F3(cell), rc.P(grid), rc.I3(gd)) for a, b in zip(x1, x2)],
                    np.float32)
    v = np.ceil(np.float32(100.0) * want)
    v = np.where(np.isnan(v) | (v <= 0), 0, np.minimum(v, 2147483647)).astype(np.uint32)
    assert np.array_equal(np.uint32(0x7FFFFFFF) - keys, v)


# ------------------------------------------------------------------------------------------------------------------
# GPU: the sm_100a kernels against the reference kernels' golden outputs directly (through the C ABI)
@pytest.mark.gpu
def test_cuda_equals_reference_golden(cpm, ctx, torch_cuda):
    """Emission, mesh intersection, every tracer variant (both volume layouts, with and without the opacity-bound grid),
    the detector (incl. both exit-point quirks and equal importance), threshold / iota, min-max bricks, both importance
    classifiers and the light-sample hash: bit for bit what the reference's own OpenCL kernels produce.  Splat: the
    single-photon contributions exactly, the accumulated volumes within the fp32 atomic-sum tolerance."""
    torch = torch_cuda
    want = dict(np.load(GOLDEN / "ref_kernels.npz"))
    S = rc.scene()
    L = S["L"]
    N, NS = rc.N, rc.NS
    dev = "cuda"
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a.view(np.int32) if a.dtype == np.uint32 else a)).to(dev)   # noqa: E731

    def eq(t, key, view=np.float32):
        g = t.cpu().numpy().view(view).reshape(want[key].shape)
        assert np.array_equal(_bits(g), _bits(want[key])), (key, float((_bits(g) != _bits(want[key])).mean()))

    s = torch.empty(N * 4, dtype=torch.float32, device=dev)
    ctx.sample_uniform2d(float(NS), float(NS), N, s)
    eq(s, "uniform2d")
    s2 = torch.empty(4000, dtype=torch.float32, device=dev)
    ctx.sample_uniform2d(33.0, 31.0, 1000, s2)
    eq(s2, "uniform2d_ragged")
    ls = torch.empty(N * 8, dtype=torch.float32, device=dev)
    ctx.light_sample_directional(s, (1.0, 0.9, 0.8), L["dir"], L["origin"], L["u"], L["v"], float(L["area"]), N, ls)
    eq(ls, "light_samples")
    it = torch.empty(N * 2, dtype=torch.float32, device=dev)
    verts, idx = T(rc.synth.CUBE_VERTICES), T(rc.synth.CUBE_INDICES)
    ctx.light_mesh_intersect(verts, idx, 36, ls, N, it)
    eq(it, "isect")
    ctx.sync()

    dvol, dtf = T(S["vol"]), T(S["tf"])
    dims = rc.DIMS
    Vlin = ctx.volume_create(dvol, dims, cpm.CPM_FMT_U8)
    gd = cpm.capi.bound_grid_dims(dims, 3)
    ncell = gd[0] * gd[1] * gd[2]
    vrange = torch.zeros(2 * ncell, dtype=torch.float32, device=dev)
    ctx.volume_value_range(Vlin, 3, vrange)
    bound = torch.zeros(ncell, dtype=torch.float32, device=dev)
    ctx.opacity_bound(vrange, ncell, dtf, bound)
    rng0 = T(orc.rng_seed_streams(orc.rng_host_base_offsets(0, N)))
    photons = {}
    for layout in (cpm.CPM_VOLUME_LINEAR, cpm.CPM_VOLUME_TEXTURE):
        V = ctx.volume_create(dvol, dims, cpm.CPM_FMT_U8, layout=layout)
        for bounded in (False, True):
            for name, I, flags, phase, material, entry, aabb in rc.TRACE_VARIANTS:
                p = cpm.make_trace_params(N, max_interactions=I, step_size=1.0 / max(dims), aabb_min=aabb[0], aabb_max=aabb[1],
                                          phase=phase, material=material, flags=flags,
                                          opacity_bound=bound if bounded else None, bound_cell_log2=3)
                ph = torch.zeros(N * I * 8, dtype=torch.float32, device=dev)
                rng = rng0.clone()
                ctx.trace_photons(V, dtf, p, ls, it, ph, rng)
                ctx.sync()
                eq(ph, "trace_" + name)
                eq(rng, "rng_" + name, np.uint32)
                photons[name] = ph
            rec = T(rc.recompute_ids())
            p = cpm.make_trace_params(N, max_interactions=2, step_size=1.0 / max(dims),
                                      opacity_bound=bound if bounded else None, bound_cell_log2=3)
            ph = torch.full((N * 2 * 8,), 7.0, dtype=torch.float32, device=dev)
            ctx.trace_photons(V, dtf, p, ls, it, ph, rng0.clone(), rec, rec.numel())
            ctx.sync()
            eq(ph, "trace_recompute_I2")
        V.destroy()

    t2i = frame.texture_to_index(dims)
    grid = T(S["grid"])
    for name in ("plain_I1", "plain_I3", "hg_I3"):
        I = int(name[-1])
        keys = torch.full((N,), 0x7FFFFFFF, dtype=torch.int32, device=dev)
        ctx.detect_invalid(grid, S["gd"], (rc.REGION,) * 3, t2i, photons[name], 0, ls, it, N, I, N, keys)
        ctx.sync()
        eq(keys, "detect_" + name, np.uint32)
        if name == "plain_I1":
            thr = torch.zeros(N, dtype=torch.int32, device=dev)
            ctx.threshold(keys, 0x7FFFFFFF, thr)
            ctx.sync()
            eq(thr, "threshold", np.uint32)
    keys = torch.full((N,), 0x7FFFFFFF, dtype=torch.int32, device=dev)
    ctx.detect_invalid(grid, S["gd"], (rc.REGION,) * 3, t2i, photons["plain_I1"], 0, ls, it, N, 1, N, keys, equal_importance=True,
                       percentage=25, iteration=3)
    eq(keys, "detect_equal_importance", np.uint32)
    io = torch.full((N,), 99, dtype=torch.int32, device=dev)
    ctx.iota(io)
    eq(io, "iota", np.uint32)

    n_cells = S["gd"][0] * S["gd"][1] * S["gd"][2]
    for nm, v, fmt in (("u8", S["vol"], cpm.CPM_FMT_U8), ("f32", S["volf"], cpm.CPM_FMT_F32)):
        dv = T(v)
        V = ctx.volume_create(dv, dims, fmt)
        mm = torch.zeros(n_cells * 2, dtype=torch.int16, device=dev)
        ctx.volume_minmax(V, rc.REGION, mm)
        ctx.sync()
        eq(mm, "minmax_" + nm, np.uint16)
        V.destroy()
    mm_np, prev_np = orc.volume_minmax(S["vol"], rc.REGION), orc.volume_minmax(S["vol_prev"], rc.REGION)
    diff_np = orc.volume_diff_bricks(S["vol_prev"], S["vol"], rc.REGION, 1.0, 0.0, 255.0)
    pos, col = frame.tf_point_lists(rc.synth.WS_TF_POINTS)
    dmm, dprev = T(mm_np.reshape(-1).view(np.int16)), T(prev_np.reshape(-1).view(np.int16))
    ddiff, dpos, dcol = T(diff_np.reshape(-1)), T(pos), T(col)
    for k, w in enumerate(rc.classify_weights()):
        o = torch.zeros(n_cells, dtype=torch.float32, device=dev)
        ctx.classify_importance(dmm, n_cells, dpos, dcol, len(pos), w, True, o)
        eq(o, f"classify_static_{k}")
        o = torch.zeros(n_cells, dtype=torch.float32, device=dev)
        ctx.classify_importance(dmm, n_cells, dpos, dcol, len(pos), w, False, o, prev=dprev, diff=ddiff)
        eq(o, f"classify_timevarying_{k}")
    ids = T((rc.synth.splitmix64(3, 700) % np.uint64(N + 50)).astype(np.uint32))
    hb = torch.zeros(700, dtype=torch.int32, device=dev)
    ctx.hash_light_samples(ls, it, N, ids, 700, (8, 8, 8), (8, 8, 8), hb)
    eq(hb, "hash", np.uint32)

    # splat
    t2, i2 = frame.texture_to_index(rc.LV), frame.index_to_texture(rc.LV)
    radius, scale = float(np.float32(1.7 / 24)), 1e-3
    nv = rc.LV[0] * rc.LV[1] * rc.LV[2]
    ph = want["trace_plain_I3"]
    stored = np.where(ph[:, 0] != FLT_MAX)[0][:6]
    for k, g in enumerate(stored):
        v = torch.zeros(nv, dtype=torch.float32, device=dev)
        ctx.splat_photons(v, 1, t2, i2, rc.LV, T(ph[g:g + 1]), None, 1, 1, 1, radius, scale)
        ctx.sync()
        got = v.cpu().numpy()
        w = want[f"splat_single_{k}"]
        # the kernel weights with the squared distance (1 - d^2 / r^2, no sqrt / division per voxel) where the reference
        # has 1 - (d / r)^2: a few ulp of the photon's peak contribution per voxel, relatively more towards the rim
        assert ((got != 0) == (w != 0)).mean() > 0.999, k
        assert np.abs(got - w).max() <= 4e-6 * np.abs(w).max(), (k, np.abs(got - w).max(), np.abs(w).max())
    dph = T(ph)
    v = torch.zeros(nv, dtype=torch.float32, device=dev)
    ctx.splat_photons(v, 1, t2, i2, rc.LV, dph, None, N * 3, N, 3, radius, scale)
    sel = T(np.arange(0, N, 3, dtype=np.uint32))
    v4 = torch.zeros(nv * 4, dtype=torch.float32, device=dev)
    ctx.splat_photons(v4, 4, t2, i2, rc.LV, dph, sel, sel.numel(), N, 3, radius, scale, -1.0)
    ctx.sync()
    for got, key in ((v, "splat_all"), (v4, "splat_selected_rgba_minus")):
        g, w = got.cpu().numpy().astype(np.float64), want[key].astype(np.float64)
        assert np.sqrt(((g - w) ** 2).mean()) / np.sqrt((w ** 2).mean()) <= 1e-5, key
    Vlin.destroy()
