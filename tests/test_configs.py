"""BASELINE.json configs at (or near) full size on the GPU: parity against the oracle on a bounded photon sample
plus size-independent properties (replay determinism, sortedness + stability + permutation checksum,
energy bookkeeping).  C1, C2, C3 and C5 call the C ABI directly; C4 (bench.py's workload) goes through the drop-in
host network and is compared frame by frame with the oracle's network (oracle/frame.py)."""
import importlib

import numpy as np
import pytest

import scenes
from conftest import PKG_NAME
from test_tracer import oracle_trace

FLT_MAX = np.float32(3.4028234663852886e38)


def _emit(cpm, ctx, torch, synth, orc, n_side, direction, radiance=(1.0, 1.0, 1.0)):
    """device light samples + intersections of one directional light (and the oracle's host copies)"""
    L = scenes.directional_light(n_side, direction, radiance=radiance)
    ls = torch.from_numpy(L["light_samples"].reshape(-1)).cuda()
    it = torch.from_numpy(L["isect"].reshape(-1)).cuda()
    return L, ls, it


def _device_rng(cpm, ctx, torch, total):
    """MWC64X states of photons [0, total) seeded on the device (bit-exact vs the oracle: tests/test_rng.py)"""
    st = torch.from_numpy(cpm.capi.rng_host_base_offsets(0, total).view(np.int32).reshape(-1).copy()).cuda()
    ctx.rng_seed_streams(st, total)
    ctx.sync()
    return st


def _trace(cpm, ctx, torch, V, tf, ls, it, n, total, offset, I, step, rng, photons, aabb=((0, 0, 0), (1, 1, 1)), cnt=None):
    p = cpm.make_trace_params(n, total_photons=total, photon_offset=offset, max_interactions=I, step_size=step,
                              aabb_min=aabb[0], aabb_max=aabb[1])
    ctx.trace_photons(V, tf, p, ls, it, photons, rng, None, 0, cnt)


@pytest.mark.gpu
def test_c1_workspace_demo_two_lights_clip(cpm, orc, synth, ctx, torch_cuda):
    """C1: 512 x 512 x 96 u8 stand-in for Subclavia.pvm, the workspace's TF, two directional lights with 256^2
    samples each, clip box (73,7,0)-(512,512,96)/dims (ws:740-757); every photon bit-exact vs the oracle.  The
    workspace runs I = 1; I = 2 is used here so that the clip box (applied when a scattered ray is re-intersected,
    ppm/cl/photontracer.cl:56) is exercised too."""
    torch = torch_cuda
    dims = (512, 512, 96)
    vol = synth.volume_field_torch(dims, 1, device="cuda")
    vol8 = torch.round(vol * 255.0).to(torch.uint8).contiguous()
    vol_np = vol8.cpu().numpy()
    tf_np = synth.rasterise_tf(width=1024)
    tf = torch.from_numpy(tf_np).cuda()
    aabb = ((73 / 512, 7 / 512, 0.0), (1.0, 1.0, 1.0))
    ns = 256
    n = ns * ns
    dirs = [(-0.36, 0.48, 0.8), (0.5, -0.3, 0.81)]
    total = 2 * n
    rng = _device_rng(cpm, ctx, torch, total)
    rng_np = rng.cpu().numpy().view(np.uint32).reshape(total, 2)
    I = 2
    photons = torch.zeros(total * I * 8, dtype=torch.float32, device="cuda")
    want = np.zeros((total * I, 8), np.float32)
    V = ctx.volume_create(vol8, dims, cpm.CPM_FMT_U8, layout=cpm.CPM_VOLUME_TEXTURE)
    step = 1.0 / 512
    for l, d in enumerate(dirs):
        L, ls, it = _emit(cpm, ctx, torch, synth, orc, ns, d)
        _trace(cpm, ctx, torch, V, tf, ls, it, n, total, l * n, I, step, rng, photons, aabb)
        oracle_trace(orc, vol_np, tf_np, L, max_interactions=I, rng=rng_np.copy(), photons=want, step_size=step, aabb=aabb,
                     total_photons=total, photon_offset=l * n)
    ctx.sync()
    got = photons.cpu().numpy().reshape(total * I, 8)
    assert (want[:, 0] != FLT_MAX).sum() > 10_000
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    second = got[total:][got[total:, 0] != FLT_MAX]
    assert second.shape[0] > 1000 and second[:, 0].min() >= 73 / 512 - 1e-4   # scattered rays respect the clip box
    V.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("logn", [26, 28])
def test_c2_sort_sweep_properties(cpm, ctx, torch_cuda, logn):
    """C2 at 2^26 / 2^28 pairs: ascending keys, stability (values ascending inside equal keys), and the values
    are a permutation (sum and xor checksums) -- importance-like and uniform keys; then the keys-only sort of a
    10 % prefix"""
    torch = torch_cuda
    n = 1 << logn
    g = torch.Generator(device="cuda").manual_seed(logn)
    for dist in ("uniform", "importance"):
        if dist == "uniform":
            keys = torch.randint(0, 2 ** 31 - 1, (n,), dtype=torch.int32, device="cuda", generator=g)
            keys ^= torch.randint(0, 2, (n,), dtype=torch.int32, device="cuda", generator=g) << 31   # all 32 bits used
        else:
            keys = torch.full((n,), 0x7FFFFFFF, dtype=torch.int32, device="cuda")
            sel = torch.rand(n, device="cuda", generator=g) < 0.1
            keys[sel] -= (-torch.log(torch.rand(int(sel.sum()), device="cuda", generator=g)) * 300).ceil().int()
        vals = torch.arange(n, dtype=torch.int32, device="cuda")
        tk, tv = torch.empty_like(keys), torch.empty_like(vals)
        src = keys.clone()
        ctx.radix_sort(keys, vals, tk, tv)
        ctx.sync()
        ku = keys.view(torch.int64) if False else keys.to(torch.int64) & 0xFFFFFFFF       # unsigned order
        dk = ku[1:] - ku[:-1]
        assert bool((dk >= 0).all()), dist
        same = dk == 0
        assert bool((vals[1:][same] > vals[:-1][same]).all()), f"{dist}: not stable"
        assert int(vals.to(torch.int64).sum()) == n * (n - 1) // 2
        assert bool((src[vals.long()] == keys).all()), f"{dist}: values do not carry their keys"
        del ku, dk, same
        m = n // 10
        idx = vals[:m].clone()
        t2 = torch.empty_like(idx)
        ctx.radix_sort(idx, None, t2, None)
        ctx.sync()
        assert bool((idx[1:] > idx[:-1]).all())
        assert int(idx.to(torch.int64).sum()) == int(vals[:m].to(torch.int64).sum())
        del keys, vals, tk, tv, src, idx, t2
        torch.cuda.empty_cache()


@pytest.mark.gpu
@pytest.mark.parametrize("I", [1, 4])
def test_c3_256cube_u8_one_million_photons(cpm, orc, synth, ctx, torch_cuda, I):
    """C3: 256^3 u8, 1024^2 photons.  Full-size properties: two traces are bit-identical (replay), both volume
    layouts agree bit for bit, stored slots form a prefix; parity: photons [0, 65536) equal the oracle's."""
    torch = torch_cuda
    dims = (256, 256, 256)
    vol_np = synth.volume_u8(dims, 3)
    tf_np = synth.rasterise_tf(width=1024)
    dvol, tf = torch.from_numpy(vol_np).cuda(), torch.from_numpy(tf_np).cuda()
    ns = 1024
    n = ns * ns
    L, ls, it = _emit(cpm, ctx, torch, synth, orc, ns, (0.3, -0.5, 0.8))
    rng0 = _device_rng(cpm, ctx, torch, n)
    out = {}
    for layout in (cpm.CPM_VOLUME_TEXTURE, cpm.CPM_VOLUME_LINEAR):
        V = ctx.volume_create(dvol, dims, cpm.CPM_FMT_U8, layout=layout)
        for rep in range(2 if layout == cpm.CPM_VOLUME_TEXTURE else 1):
            rng = rng0.clone()
            ph = torch.zeros(n * I * 8, dtype=torch.float32, device="cuda")
            cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
            _trace(cpm, ctx, torch, V, tf, ls, it, n, n, 0, I, 1.0 / 256, rng, ph, cnt=cnt)
            ctx.sync()
            out[(layout, rep)] = (ph, int(cnt.item()))
        V.destroy()
    a, b, c = out[(cpm.CPM_VOLUME_TEXTURE, 0)], out[(cpm.CPM_VOLUME_TEXTURE, 1)], out[(cpm.CPM_VOLUME_LINEAR, 0)]
    assert torch.equal(a[0].view(torch.int32), b[0].view(torch.int32)) and a[1] == b[1]
    assert torch.equal(a[0].view(torch.int32), c[0].view(torch.int32)) and a[1] == c[1]
    ph = a[0].view(I, n, 8)
    stored = ph[:, :, 0] != float(FLT_MAX)
    assert bool((stored[1:].int() <= stored[:-1].int()).all())
    # oracle on the first 65536 photons (same stream ids: photon_offset + i)
    m = 65536
    sub = dict(L)
    sub["n"] = m
    sub["light_samples"] = np.ascontiguousarray(L["light_samples"][:m])
    sub["isect"] = np.ascontiguousarray(L["isect"][:m])
    want, _, _ = oracle_trace(orc, vol_np, tf_np, sub, max_interactions=I, rng=scenes.rng_states(m), step_size=1.0 / 256)
    got = ph[:, :m, :].cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.reshape(I, m, 8).view(np.uint32))


@pytest.mark.gpu
def test_c5_1024cube_four_lights_sharded_ranges(cpm, orc, synth, ctx, torch_cuda):
    """C5 (single-GPU part): 1024^3 u8 (1 GiB), 4 directional lights x 2048^2 = 16 Mi photons traced as 8 photon
    ranges the way 8 ranks would own them; properties: every range is bit-identical to the same range of the
    unsharded trace; parity: 16384 photons of light 2 equal the oracle's."""
    torch = torch_cuda
    dims = (1024, 1024, 1024)
    vol = synth.volume_field_torch(dims, 5, device="cuda")
    vol8 = torch.round(vol * 255.0).to(torch.uint8).contiguous()
    del vol
    tf_np = synth.rasterise_tf(width=1024)
    tf = torch.from_numpy(tf_np).cuda()
    ns = 2048
    n = ns * ns
    dirs = [(0.3, -0.5, 0.8), (-0.6, 0.2, 0.77), (0.1, 0.9, -0.42), (-0.5, -0.5, -0.7)]
    total = 4 * n
    V = ctx.volume_create(vol8, dims, cpm.CPM_FMT_U8, layout=cpm.CPM_VOLUME_TEXTURE)
    base = torch.from_numpy(cpm.capi.rng_host_base_offsets(0, total).view(np.int32).reshape(-1)).cuda()
    ctx.rng_seed_streams(base, total)
    photons = torch.zeros(total * 8, dtype=torch.float32, device="cuda")
    lights = []
    for l, d in enumerate(dirs):
        L, ls, it = _emit(cpm, ctx, torch, synth, orc, ns, d)
        lights.append((L, ls, it))
        _trace(cpm, ctx, torch, V, tf, ls, it, n, total, l * n, 1, 1.0 / 1024, base, photons)
    ctx.sync()
    full = photons.view(total, 8)
    assert int((full[:, 0] != float(FLT_MAX)).sum()) > total // 10
    # 8 shards of 2 Mi photons: half a light each, traced from their own state slices into their own buffers
    shard_n = total // 8
    for r in (0, 3, 5, 7):
        l, half = divmod(r, 2)
        L, ls, it = lights[l]
        lo = half * shard_n
        st = torch.zeros(shard_n * 2, dtype=torch.int32, device="cuda")
        hb = np.zeros((shard_n, 2), np.uint32)
        import ctypes as C
        cpm.lib().cpm_rng_host_base_offsets_range(C.c_uint32(0), C.c_uint64(r * shard_n), hb.ctypes.data_as(C.c_void_p),
                                                  C.c_size_t(shard_n))
        st.copy_(torch.from_numpy(hb.view(np.int32).reshape(-1)))
        ctx.rng_seed_streams(st, shard_n, first_stream=r * shard_n)
        out = torch.zeros(shard_n * 8, dtype=torch.float32, device="cuda")
        _trace(cpm, ctx, torch, V, tf, ls[lo * 8:(lo + shard_n) * 8], it[lo * 2:(lo + shard_n) * 2], shard_n, shard_n, 0, 1,
               1.0 / 1024, st, out)
        ctx.sync()
        assert torch.equal(out.view(torch.int32), full[r * shard_n:(r + 1) * shard_n].reshape(-1).view(torch.int32)), r
    # oracle parity on 16384 photons of light 2 (needs the 1 GiB volume on the host)
    m = 16384
    vol_np = vol8.cpu().numpy()
    L = lights[2][0]
    sub = dict(L)
    sub["n"] = m
    sub["light_samples"] = np.ascontiguousarray(L["light_samples"][:m])
    sub["isect"] = np.ascontiguousarray(L["isect"][:m])
    rng_np = orc.rng_seed_streams(orc.rng_host_base_offsets(0, 2 * n + m)[2 * n:].copy(), first_stream=2 * n)
    want, _, _ = oracle_trace(orc, vol_np, tf_np, sub, max_interactions=1, rng=rng_np, step_size=1.0 / 1024)
    got = full[2 * n:2 * n + m].cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    V.destroy()


# ---------------------------------------------------------------------------------------------------------------
# C4: time-varying f32 volume, correlated re-tracing through the drop-in host network, against the oracle's network
def oracle_lights_from_host(orc, synth, net, ns):
    """The oracle's light samples built from the kernel arguments the host layer derived (direction and plane point
    through the light's transform matrix, lcl/directionallightsamplercl.cpp:66-73).  Asserts on the way that the host's
    CPU plane fit equals the oracle's (== the reference's own files, tests/test_ref_geometry.py) and that the emitted
    light samples / intersections are bit-identical."""
    lights = []
    f = np.float32
    for l in range(net.cfg.n_lights):
        S = net.light_setup(l)
        o, u, v = orc.fit_light_plane(synth.CUBE_VERTICES, S["plane_point"], S["dir"])
        for got, want in ((S["origin"], o), (S["u"], u), (S["v"], v)):
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (l, got, want)
        lu = np.sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2])
        lv = np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
        assert f(lu * lv) == S["area"]
        samples = orc.sample_uniform2d(float(ns), float(ns), ns * ns)
        ls = orc.light_sample_directional(samples, S["radiance"], S["dir"], o, u, v, float(S["area"]))
        isect = orc.light_mesh_intersect(synth.CUBE_VERTICES, synth.CUBE_INDICES, ls)
        got_ls, got_it = net.read_light_samples(l)
        assert np.array_equal(got_ls.view(np.uint32), ls.view(np.uint32)), l
        assert np.array_equal(got_it.view(np.uint32), isect.view(np.uint32)), l
        lights.append(dict(light_samples=ls, isect=isect))
    return lights


def _rel_rmse(got, want):
    return float(np.sqrt(((got.astype(np.float64) - want) ** 2).mean()) / max(np.sqrt((want ** 2).mean()), 1e-300))


@pytest.mark.gpu
@pytest.mark.parametrize("I,n_lights", [(1, 1), (2, 2)])
def test_c4_correlated_frames_match_oracle(cpm, orc, synth, torch_cuda, I, n_lights):
    """C4-shaped case (f32 time series, time-varying Lab classify + volume difference + detector + selection +
    indexed re-trace + -old/+new splat) through libcpm_host.so, four time-step changes incl. the wrap-around:
    per frame equal re-trace count, equal id list, bit-equal importance grid and photon records, light volume
    within 1e-5 relative RMSE of the oracle's float64 accumulation."""
    from oracle import frame
    host = importlib.import_module(PKG_NAME + ".host")
    D, ns, T = 128, 256, 4
    dims = (D, D, D)
    vols = [synth.volume_f32(dims, 4, t / 32.0) for t in range(T)]
    dirs = [(0.3, -0.5, 0.8), (-0.6, 0.2, 0.77)][:n_lights]
    net = host.Network(dims, cpm.CPM_FMT_F32, ns, dirs, max_scattering_events=I, light_volume_option=2,
                       with_importance_grid=True, reference_full_splat_bound=False)
    net.set_transfer_function(synth.WS_TF_POINTS)
    net.set_sequence_host(vols)
    net.set_timestep(0)
    net.evaluate()
    assert net.n_recomputed == -1
    O = frame.OracleNetwork(dims, oracle_lights_from_host(orc, synth, net, ns), frame.rasterise_tf(synth.WS_TF_POINTS),
                            synth.WS_TF_POINTS, max_interactions=I)
    O.first_frame(vols[0])
    assert np.array_equal(net.read_photons(I).view(np.uint32), O.photons.view(np.uint32))
    assert _rel_rmse(net.read_light_volume(), O.lightvol) <= 1e-5
    mm = [orc.volume_minmax(v, 8) for v in vols]
    diff = [orc.volume_diff_bricks(vols[t], vols[(t + 1) % T], 8) for t in range(T)]
    n_cells = mm[0].size // 2
    total = 0
    for step in range(1, T + 2):              # 1, 2, 3, 0 (wrap: (t + 1) % T, dynamicvolumedifferenceanalysis.cpp:64), 1
        t, tp = step % T, (step - 1) % T
        net.set_timestep(t)
        net.evaluate()
        imp = O.importance_time_varying(mm[t], mm[tp], diff[tp])
        assert np.array_equal(net.read_importance_grid(n_cells).view(np.uint32), imp.view(np.uint32)), step
        ids = O.frame(vols[t], imp)
        assert net.n_recomputed == ids.size, (step, net.n_recomputed, ids.size)
        assert np.array_equal(net.read_recomputed_indices(), ids), step
        assert np.array_equal(net.read_photons(I).view(np.uint32), O.photons.view(np.uint32)), step
        assert np.array_equal(net.read_importance_keys(), O.keys), step
        assert net.last_splat_path == O.last_splat_path, (step, net.last_splat_path, O.last_splat_path)
        assert _rel_rmse(net.read_light_volume(), O.lightvol) <= 1e-5, step
        total += ids.size
    assert 0 < total < (T + 1) * O.n          # something was re-traced, and not everything
    net.close()


@pytest.mark.gpu
def test_c4_full_size_retrace_counts_match_oracle(cpm, orc, synth, torch_cuda):
    """C4 at BASELINE size (512^3 f32, 2048^2 photons): the number of photons the network re-traces per time step and
    the re-traced id lists equal the oracle's for two consecutive steps (the records themselves are compared at
    128^3 above and, sampled, here)."""
    from oracle import frame
    torch = torch_cuda
    host = importlib.import_module(PKG_NAME + ".host")
    D, ns, T = 512, 2048, 3
    dims = (D, D, D)
    vols = [synth.volume_field_torch(dims, 4, t / 32.0, device="cuda").cpu().numpy() for t in range(T)]
    net = host.Network(dims, cpm.CPM_FMT_F32, ns, [(0.3, -0.5, 0.8)], max_scattering_events=1, light_volume_option=2,
                       with_importance_grid=True, reference_full_splat_bound=False)
    net.set_transfer_function(synth.WS_TF_POINTS)
    net.set_sequence_host(vols)
    net.set_timestep(0)
    net.evaluate()
    O = frame.OracleNetwork(dims, oracle_lights_from_host(orc, synth, net, ns), frame.rasterise_tf(synth.WS_TF_POINTS),
                            synth.WS_TF_POINTS, max_interactions=1)
    O.first_frame(vols[0])
    mm = [orc.volume_minmax(v, 8) for v in vols]
    for t in (1, 2):
        net.set_timestep(t)
        net.evaluate()
        imp = O.importance_time_varying(mm[t], mm[t - 1], orc.volume_diff_bricks(vols[t - 1], vols[t], 8))
        ids = O.frame(vols[t], imp)
        assert net.n_recomputed == ids.size, (t, net.n_recomputed, ids.size)
        assert 0.05 * O.n < ids.size < 0.95 * O.n
        assert np.array_equal(net.read_recomputed_indices(), ids), t
        assert np.array_equal(net.read_photons(1).view(np.uint32), O.photons.view(np.uint32)), t
    net.close()
