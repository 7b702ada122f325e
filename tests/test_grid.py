"""Uniform-grid operators: min/max bricks, inter-step difference, importance classification,
light-sample hash, cell ranges."""
import numpy as np
import pytest

import scenes


def _dev_vol(cpm, ctx, torch, vol_np):
    fmt = {np.dtype(np.uint8): cpm.CPM_FMT_U8, np.dtype(np.uint16): cpm.CPM_FMT_U16,
           np.dtype(np.float32): cpm.CPM_FMT_F32}[vol_np.dtype]
    raw = vol_np.view(np.int16) if vol_np.dtype == np.uint16 else vol_np
    d = torch.from_numpy(np.ascontiguousarray(raw)).cuda()
    return ctx.volume_create(d, (vol_np.shape[2], vol_np.shape[1], vol_np.shape[0]), fmt), d


def tf_points(synth):
    """host TF point list as updateTransferFunctionData builds it
    (isc/processors/minmaxuniformgrid3dimportanceclprocessor.cpp:304-362): end points added at 0 and 1"""
    pts = synth.WS_TF_POINTS
    pos = [0.0] + [p[0] for p in pts] + [1.0]
    col = [pts[0][1]] + [p[1] for p in pts] + [pts[-1][1]]
    return np.array(pos, np.float32), np.ascontiguousarray(np.array(col, np.float32))


# ------------------------------------------------------------------------------- CPU ---------
def test_oracle_minmax_against_numpy(orc, synth):
    vol = synth.volume_u8((40, 24, 19), 5)          # ragged: not multiples of the region
    mm = orc.volume_minmax(vol, 8)
    assert mm.shape == (3, 3, 5, 2)
    for (gz, gy, gx) in [(0, 0, 0), (2, 2, 4), (1, 2, 3)]:
        blk = vol[gz * 8:(gz + 1) * 8, gy * 8:(gy + 1) * 8, gx * 8:(gx + 1) * 8].astype(np.float64) / 255.0
        assert mm[gz, gy, gx, 0] == int(np.rint(np.float32(blk.min()) * np.float32(65535)))
        assert mm[gz, gy, gx, 1] == int(np.rint(np.float32(blk.max()) * np.float32(65535)))


def test_oracle_diff_bricks_edge_divides_by_full_region(orc, synth):
    """ugc/processors/dynamicvolumedifferenceanalysis.h:147"""
    a = np.zeros((4, 4, 12), np.uint8)
    b = np.full((4, 4, 12), 10, np.uint8)
    d = orc.volume_diff_bricks(a, b, 8, 1.0, 0.0, 255.0)
    assert d.shape == (1, 1, 2)
    assert np.isclose(d[0, 0, 0], 10.0 * (8 * 4 * 4) / 512 / 255)
    assert np.isclose(d[0, 0, 1], 10.0 * (4 * 4 * 4) / 512 / 255)


def test_oracle_cell_ranges(orc):
    keys = np.array([1, 1, 3, 3, 3, 7], np.uint32)
    s, e = orc.build_cell_ranges(keys, 9)
    assert s.tolist() == [0, 0, 2, 2, 5, 5, 5, 5, 6]
    assert e.tolist() == [0, 2, 2, 5, 5, 5, 5, 6, 6]
    s, e = orc.build_cell_ranges(np.zeros(0, np.uint32), 3)
    assert s.tolist() == [0, 0, 0] and e.tolist() == [0, 0, 0]


def test_oracle_classify_incremental_is_sum_of_max_colour(orc, synth):
    pos, col = tf_points(synth)
    mm = np.array([[0, 65535], [0, 100], [30000, 30000]], np.uint16)
    imp = orc.classify_importance(mm, pos, col, (0, 0, 0, 1), True)
    assert np.isclose(imp[0], col.max(axis=0).sum(), rtol=1e-6)
    assert imp[1] == np.float32(col[0].sum())          # range entirely left of the first real point
    assert imp[2] > 0


# ------------------------------------------------------------------------------- GPU ---------
@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["u8", "u16", "f32"])
@pytest.mark.parametrize("dims,region", [((64, 48, 40), 8), ((50, 33, 17), 8), ((48, 33, 17), 8), ((528, 9, 8), 8),
                                         ((64, 64, 64), 5), ((32, 32, 32), 1)])
def test_cuda_minmax_bit_exact(cpm, orc, ctx, torch_cuda, fmt, dims, region):
    torch = torch_cuda
    vol = scenes.make_volume(dims, fmt, 9)
    want = orc.volume_minmax(vol, region)
    V, keep = _dev_vol(cpm, ctx, torch, vol)
    out = torch.zeros(want.size, dtype=torch.int16, device="cuda")
    od = ctx.volume_minmax(V, region, out)
    ctx.sync()
    assert od == (want.shape[2], want.shape[1], want.shape[0])
    assert np.array_equal(out.cpu().numpy().view(np.uint16), want.reshape(-1))
    V.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["u8", "u16", "f32"])
@pytest.mark.parametrize("dims,region", [((72, 40, 36), 8), ((80, 37, 19), 8), ((1040, 8, 9), 8), ((75, 41, 19), 8),
                                         ((64, 48, 40), 5), ((33, 20, 17), 16)])
def test_cuda_diff_bricks(cpm, orc, ctx, torch_cuda, synth, fmt, dims, region):
    """One thread per brick adds the voxels in the reference's order (z, y, x; dynamicvolumedifferenceanalysis.h:123-138):
    the double sum -- and with it the float result -- equals the oracle's bit for bit for every format.
    (rows that are a multiple of 8 voxels: the vector-load kernel, ragged y / z; 75 / 33 and regions 5 / 16: the scalar one)"""
    torch = torch_cuda
    if fmt == "u8":
        a, b = synth.volume_u8(dims, 1), synth.volume_u8(dims, 2)
        rng = (1.0, 0.0, 255.0)
    elif fmt == "u16":
        a, b = synth.volume_u16(dims, 1), synth.volume_u16(dims, 2)
        rng = (65535.0 / 4095.0, 0.0, 4095.0)      # 12-bit data in a 16-bit volume: a scaling that is not 1
    else:
        a, b = synth.volume_f32(dims, 4, 0.0), synth.volume_f32(dims, 4, 1.0 / 32)
        rng = (1.0, 0.0, 1.0)
    want = orc.volume_diff_bricks(a, b, region, *rng)
    Va, ka = _dev_vol(cpm, ctx, torch, a)
    Vb, kb = _dev_vol(cpm, ctx, torch, b)
    out = torch.zeros(want.size, dtype=torch.float32, device="cuda")
    ctx.volume_diff_bricks(Va, Vb, region, *rng, out)
    ctx.sync()
    got = out.cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.reshape(-1).view(np.uint32))
    Va.destroy(); Vb.destroy()


@pytest.mark.gpu
def test_cuda_classify_importance(cpm, orc, ctx, torch_cuda, synth):
    torch = torch_cuda
    vol_a, vol_b = synth.volume_u8((64, 64, 64), 1), synth.volume_u8((64, 64, 64), 2)
    mm, prev = orc.volume_minmax(vol_a, 8), orc.volume_minmax(vol_b, 8)
    diff = orc.volume_diff_bricks(vol_b, vol_a, 8, 1.0, 0.0, 255.0)
    pos, col = tf_points(synth)
    n = mm.size // 2
    dmm = torch.from_numpy(mm.reshape(-1).view(np.int16)).cuda()
    dprev = torch.from_numpy(prev.reshape(-1).view(np.int16)).cuda()
    ddiff = torch.from_numpy(diff.reshape(-1)).cuda()
    dpos, dcol = torch.from_numpy(pos).cuda(), torch.from_numpy(col).cuda()
    out = torch.zeros(n, dtype=torch.float32, device="cuda")
    # static, incremental formula: pure fp32 arithmetic -> bit exact
    w = (0.0, 0.0, 0.0, 1.0)
    ctx.classify_importance(dmm, n, dpos, dcol, len(pos), w, True, out)
    ctx.sync()
    want = orc.classify_importance(mm, pos, col, w, True)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))
    # time-varying, Lab formula: pow / cbrt of rgb2lab come from include/cpm_detmath.h on both sides -> bit exact
    norm = 1.0 / np.sqrt(100.0 ** 2 + 500.0 ** 2 + 400.0 ** 2)
    w = (0.5 * norm / 2.0, 0.5 * norm / 2.0, 0.5 / 2.0, 0.5 / 2.0)   # ws:472-483, normalised as the host does
    ctx.classify_importance(dmm, n, dpos, dcol, len(pos), w, False, out, prev=dprev, diff=ddiff)
    ctx.sync()
    want = orc.classify_importance(mm, pos, col, w, False, prev=prev, diff=diff.reshape(-1))
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))
    assert want.max() > 0


@pytest.mark.gpu
def test_cuda_hash_and_cell_ranges(cpm, orc, ctx, torch_cuda, synth):
    torch = torch_cuda
    L = scenes.directional_light(96)
    n = L["n"]
    ids = (synth.splitmix64(3, 5000) % np.uint64(n + 50)).astype(np.uint32)   # some ids out of range
    nb = (32, 32, 32)
    want = np.full(5000, 0xDEADBEEF, np.uint32)
    orc.hash_light_samples(L["light_samples"], L["isect"], n, ids, nb, nb, want)
    dls, dis = torch.from_numpy(L["light_samples"]).cuda(), torch.from_numpy(L["isect"]).cuda()
    dids = torch.from_numpy(ids.view(np.int32)).cuda()
    out = torch.from_numpy(np.full(5000, 0xDEADBEEF, np.uint32).view(np.int32)).cuda()
    ctx.hash_light_samples(dls, dis, n, dids, 5000, nb, nb, out)
    ctx.sync()
    got = out.cpu().numpy().view(np.uint32)
    assert np.array_equal(got, want)
    # cell ranges over the sorted valid keys
    keys = np.sort(got[got != 0xDEADBEEF])
    ncell = nb[0] * nb[1] * nb[2]
    for kk, nc in ((keys, ncell), (keys[:1], ncell), (np.zeros(0, np.uint32), 7), (keys, 100)):
        ws, we = orc.build_cell_ranges(kk, nc)
        dk = torch.from_numpy(kk.view(np.int32)).cuda() if kk.size else None
        s = torch.full((nc,), -1, dtype=torch.int32, device="cuda")
        e = torch.full((nc,), -1, dtype=torch.int32, device="cuda")
        ctx.build_cell_ranges(dk, kk.size, nc, s, e)
        ctx.sync()
        assert np.array_equal(s.cpu().numpy().view(np.uint32), ws)
        assert np.array_equal(e.cpu().numpy().view(np.uint32), we)


@pytest.mark.gpu
def test_cuda_mix_kernel(cpm, ctx, torch_cuda, synth):
    """mixKernel (ugc/cl/buffermixer.cl:37-48): x + (y - x) * a; integer formats truncate toward zero.  First against the
    reference's own kernel (tests/golden/ref_kernels.npz: float and uchar buffers), then at size against numpy."""
    torch = torch_cuda
    from conftest import GOLDEN
    gold = np.load(GOLDEN / "ref_kernels.npz")
    x = synth.volume_f32((8, 8, 8), 1).reshape(-1)
    y = synth.volume_f32((8, 8, 8), 2).reshape(-1)
    out = torch.zeros(x.size, dtype=torch.float32, device="cuda")
    ctx.mix(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), 0.3, x.size, cpm.CPM_FMT_F32, out)
    ctx.sync()
    assert np.array_equal(out.cpu().numpy().view(np.uint32), gold["mix_f32"].view(np.uint32))
    bx = (synth.splitmix64(3, 4096) & np.uint64(255)).astype(np.uint8)
    by = (synth.splitmix64(4, 4096) & np.uint64(255)).astype(np.uint8)
    for k, a in enumerate((0.0, 0.3, 0.5, 1.0)):
        ob = torch.zeros(4096, dtype=torch.uint8, device="cuda")
        ctx.mix(torch.from_numpy(bx).cuda(), torch.from_numpy(by).cuda(), a, 4096, cpm.CPM_FMT_U8, ob)
        ctx.sync()
        assert np.array_equal(ob.cpu().numpy(), gold[f"mix_u8_{k}"]), a
    n = 100_003
    for a in (0.0, 0.25, 1.0):
        fx = synth.uniform01(1, n).astype(np.float32)
        fy = synth.uniform01(2, n).astype(np.float32)
        out = torch.zeros(n, dtype=torch.float32, device="cuda")
        ctx.mix(torch.from_numpy(fx).cuda(), torch.from_numpy(fy).cuda(), a, n, cpm.CPM_FMT_F32, out)
        ctx.sync()
        want = (fx.astype(np.float64) + (fy - fx).astype(np.float64) * np.float64(np.float32(a))).astype(np.float32)   # fma
        assert np.abs(out.cpu().numpy().view(np.int32) - want.view(np.int32)).max() <= 1      # (double rounding: <= 1 ulp)
        bx = (synth.splitmix64(3, n) & np.uint64(255)).astype(np.uint8)
        by = (synth.splitmix64(4, n) & np.uint64(255)).astype(np.uint8)
        ob = torch.zeros(n, dtype=torch.uint8, device="cuda")
        ctx.mix(torch.from_numpy(bx).cuda(), torch.from_numpy(by).cuda(), a, n, cpm.CPM_FMT_U8, ob)
        ctx.sync()
        wb = np.trunc(bx.astype(np.float64) + (by.astype(np.float64) - bx.astype(np.float64)) * np.float64(np.float32(a))).astype(np.uint8)
        assert np.array_equal(ob.cpu().numpy(), wb)
