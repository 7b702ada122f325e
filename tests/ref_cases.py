"""Seeded inputs of the reference-kernel parity cases, shared by tools/make_golden.py (which runs the reference's own
kernels, oracle/_ref/libcl_ref.so, and writes tests/golden/ref_kernels.npz) and tests/test_ref_kernels.py."""
import ctypes as C
import importlib

import numpy as np

from conftest import PKG_NAME
from oracle import frame, orc

synth = importlib.import_module(PKG_NAME + ".synth")
P = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)   # noqa: E731
F3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])               # noqa: E731
F16 = lambda v: (C.c_float * 16)(*[float(x) for x in v])             # noqa: E731
I3 = lambda v: (C.c_int * 3)(*[int(x) for x in v])                   # noqa: E731

NS = 40                      # light samples per side
N = NS * NS
DIMS = (48, 40, 36)          # nx, ny, nz: not a multiple of the brick size along z
REGION = 8
LV = (24, 20, 18)


def scene():
    vol = synth.volume_u8(DIMS, 3)
    vol_prev = synth.volume_u8(DIMS, 2)
    volf = synth.volume_f32(DIMS, 4, 1.0 / 32)
    tf = frame.rasterise_tf(synth.WS_TF_POINTS)
    L = frame.directional_light(NS, (0.3, -0.5, 0.8), radiance=(1.0, 0.9, 0.8))
    rs = np.random.default_rng(1)
    gd = tuple(-(-d // REGION) for d in DIMS)
    grid = (rs.random(gd[0] * gd[1] * gd[2]) * (rs.random(gd[0] * gd[1] * gd[2]) < 0.3)).astype(np.float32)
    return dict(vol=vol, vol_prev=vol_prev, volf=volf, tf=tf, L=L, grid=grid, gd=gd)


TRACE_VARIANTS = [  # name, max interactions, flags, phase, material, entry point of libcl_ref.so, aabb
    ("plain_I1", 1, 0, 0, (0, 0, 0, 0), "ref_trace_photons", ((0, 0, 0), (1, 1, 1))),
    ("plain_I3", 3, 0, 0, (0, 0, 0, 0), "ref_trace_photons", ((0, 0, 0), (1, 1, 1))),
    ("clip_I2", 2, 0, 0, (0, 0, 0, 0), "ref_trace_photons", ((0.15, 0.02, 0.0), (1.0, 1.0, 0.9))),
    ("hg_I3", 3, 0, 1, (0.6, 0.1, 1.0, 0.0), "ref_trace_photons", ((0, 0, 0), (1, 1, 1))),
    ("nss_I2", 2, 2, 0, (0, 0, 0, 0), "ref_trace_photons_nss", ((0, 0, 0), (1, 1, 1))),
    ("progressive_I2", 2, 1, 0, (0, 0, 0, 0), "ref_trace_photons_progressive", ((0, 0, 0), (1, 1, 1))),
]


def trace_params(I, flags, phase, material, aabb, n=N, offset=0, total=None):
    return orc.trace_params(n_light_samples=n, max_interactions=I, step_size=1.0 / max(DIMS), flags=flags, phase=phase,
                            material=material, aabb_min=aabb[0], aabb_max=aabb[1], photon_offset=offset, total_photons=total)


def oracle_trace(S, I, flags, phase, material, aabb, recompute=None):
    rng = orc.rng_seed_streams(orc.rng_host_base_offsets(0, N))
    ph = np.zeros((N * I, 8), np.float32)
    if recompute is not None:
        ph[:] = 7.0          # untouched records must stay untouched
    kw = {} if recompute is None else dict(recompute=recompute, n_recompute=int(recompute.size))
    orc.trace_photons(orc.volume(S["vol"]), S["tf"], trace_params(I, flags, phase, material, aabb), S["L"]["light_samples"],
                      S["L"]["isect"], ph, rng, **kw)
    return ph, rng


def ref_trace(ref, S, I, flags, phase, material, aabb, entry, recompute=None):
    rng = orc.rng_seed_streams(orc.rng_host_base_offsets(0, N))
    ph = np.zeros((N * I, 8), np.float32)
    if recompute is not None:
        ph[:] = 7.0
    V = orc.volume(S["vol"])
    p = trace_params(I, flags, phase, material, aabb)
    getattr(ref, entry)(C.byref(V), P(S["tf"]), int(S["tf"].shape[0]), C.byref(p), P(S["L"]["light_samples"]), P(S["L"]["isect"]),
                        P(recompute), 0 if recompute is None else int(recompute.size), P(ph), P(rng))
    return ph, rng


def recompute_ids():
    """an index list as the re-trace gets it: ascending, with entries outside this light's range"""
    ids = np.arange(3, N, 7, dtype=np.uint32)
    return np.ascontiguousarray(np.concatenate([ids, np.array([N + 5, N + 99], np.uint32)]))


def classify_weights():
    return [(0.0, 0.0, 0.0, 1.0), tuple(float(x) for x in frame.importance_weights(0.5, 0.5, 0.5, 0.5))]


def ref_outputs(ref):
    """every parity case evaluated by the reference's own kernels -> dict of arrays"""
    S = scene()
    L = S["L"]
    out = {}
    a = np.zeros((N, 4), np.float32)
    ref.ref_sample_uniform2d(C.c_float(NS), C.c_float(NS), N, P(a))
    out["uniform2d"] = a
    b = np.zeros((1000, 4), np.float32)            # n that is not nx * ny: the un-floored uv.y quirk
    ref.ref_sample_uniform2d(C.c_float(33), C.c_float(31), 1000, P(b))
    out["uniform2d_ragged"] = b
    ls = np.zeros((N, 8), np.float32)
    ref.ref_light_sample_directional(P(a), F3((1.0, 0.9, 0.8)), F3(L["dir"]), F3(L["origin"]), F3(L["u"]), F3(L["v"]),
                                     C.c_float(float(L["area"])), N, P(ls))
    out["light_samples"] = ls
    it = np.zeros((N, 2), np.float32)
    ref.ref_light_mesh_intersect(P(synth.CUBE_VERTICES), P(synth.CUBE_INDICES), 36, P(ls), N, P(it))
    out["isect"] = it
    for name, I, flags, phase, material, entry, aabb in TRACE_VARIANTS:
        ph, rng = ref_trace(ref, S, I, flags, phase, material, aabb, entry)
        out["trace_" + name], out["rng_" + name] = ph, rng
    ph, _ = ref_trace(ref, S, 2, 0, 0, (0, 0, 0, 0), ((0, 0, 0), (1, 1, 1)), "ref_trace_photons_recompute", recompute_ids())
    out["trace_recompute_I2"] = ph
    # detector on the photons of two trace variants (I = 1: the interaction-0 exit quirk; I = 3: the FLT_MAX add)
    t2i = frame.texture_to_index(DIMS)
    for name in ("plain_I1", "plain_I3", "hg_I3"):
        I = int(name[-1])
        keys = np.full(N, 0x7FFFFFFF, np.uint32)
        ref.ref_detect_invalid(P(S["grid"]), I3(S["gd"]), F3((REGION,) * 3), F16(t2i), P(out["trace_" + name]), 0, P(ls), P(it),
                               N, I, N, P(keys), 0, 100, 0)
        out["detect_" + name] = keys
    keys = np.full(N, 0x7FFFFFFF, np.uint32)
    ref.ref_detect_invalid(P(S["grid"]), I3(S["gd"]), F3((REGION,) * 3), F16(t2i), P(out["trace_plain_I1"]), 0, P(ls), P(it), N, 1,
                           N, P(keys), 1, 25, 3)
    out["detect_equal_importance"] = keys
    thr = np.zeros(N, np.uint32)
    ref.ref_threshold(P(out["detect_plain_I1"]), C.c_uint32(0x7FFFFFFF), N, P(thr))
    out["threshold"] = thr
    idx = np.full(N, 99, np.uint32)
    ref.ref_index_to_buffer(P(idx), N)
    out["iota"] = idx
    for nm, v in (("u8", S["vol"]), ("f32", S["volf"])):
        mm = np.zeros(S["gd"][::-1] + (2,), np.uint16)
        V = orc.volume(v)
        ref.ref_volume_minmax_lab(C.byref(V), REGION, P(mm))
        out["minmax_" + nm] = mm
    mm, prev = orc.volume_minmax(S["vol"], REGION), orc.volume_minmax(S["vol_prev"], REGION)
    diff = orc.volume_diff_bricks(S["vol_prev"], S["vol"], REGION, 1.0, 0.0, 255.0)
    pos, col = frame.tf_point_lists(synth.WS_TF_POINTS)
    n_cells = mm.size // 2
    for k, w in enumerate(classify_weights()):
        ww = (C.c_float * 4)(*w)
        o = np.zeros(n_cells, np.float32)
        ref.ref_classify_importance_incremental(P(mm), None, None, n_cells, P(pos), P(col), len(pos), ww, P(o))
        out[f"classify_static_{k}"] = o
        o = np.zeros(n_cells, np.float32)
        ref.ref_classify_importance_lab(P(mm), P(prev), P(diff), n_cells, P(pos), P(col), len(pos), ww, P(o))
        out[f"classify_timevarying_{k}"] = o
    ids = (synth.splitmix64(3, 700) % np.uint64(N + 50)).astype(np.uint32)
    hb = np.zeros(700, np.uint32)
    ref.ref_hash_light_samples(P(ls), P(it), N, P(ids), 700, F3((8, 8, 8)), I3((8, 8, 8)), P(hb), 0)
    out["hash"] = hb
    # splat: single photons (exact per-voxel contributions) and the whole set (fp32 adds in work-item order)
    t2, i2 = frame.texture_to_index(LV), frame.index_to_texture(LV)
    radius, scale = float(np.float32(1.7 / 24)), 1e-3
    ph = out["trace_plain_I3"]
    stored = np.where(ph[:, 0] != np.float32(3.4028234663852886e38))[0][:6]
    for k, g in enumerate(stored):
        v = np.zeros(LV[0] * LV[1] * LV[2], np.float32)
        one = np.ascontiguousarray(ph[g:g + 1])
        ref.ref_splat_1(P(v), F16(t2), F16(i2), I3(LV), P(one), None, 1, 1, 1, C.c_float(radius), C.c_float(scale), C.c_float(1.0))
        out[f"splat_single_{k}"] = v
    v = np.zeros(LV[0] * LV[1] * LV[2], np.float32)
    ref.ref_splat_1(P(v), F16(t2), F16(i2), I3(LV), P(ph), None, N * 3, N, 3, C.c_float(radius), C.c_float(scale), C.c_float(1.0))
    out["splat_all"] = v
    sel = np.arange(0, N, 3, dtype=np.uint32)
    v4 = np.zeros(LV[0] * LV[1] * LV[2] * 4, np.float32)
    ref.ref_splat_4(P(v4), F16(t2), F16(i2), I3(LV), P(ph), P(sel), int(sel.size), N, 3, C.c_float(radius), C.c_float(scale),
                    C.c_float(-1.0))
    out["splat_selected_rgba_minus"] = v4
    x = synth.volume_f32((8, 8, 8), 1).reshape(-1)
    y = synth.volume_f32((8, 8, 8), 2).reshape(-1)
    m = np.zeros_like(x)
    ref.ref_mix_f32(P(x), P(y), C.c_float(0.3), x.size, P(m))
    out["mix_f32"] = m
    bx = (synth.splitmix64(3, 4096) & np.uint64(255)).astype(np.uint8)
    by = (synth.splitmix64(4, 4096) & np.uint64(255)).astype(np.uint8)
    for k, a in enumerate((0.0, 0.3, 0.5, 1.0)):
        mb = np.zeros_like(bx)
        ref.ref_mix_u8(P(bx), P(by), C.c_float(a), bx.size, P(mb))
        out[f"mix_u8_{k}"] = mb
    return out
