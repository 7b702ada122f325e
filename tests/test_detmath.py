"""include/cpm_detmath.h: accuracy against float64 libm (CPU) and host == device bits (GPU)."""
import numpy as np
import pytest


def _inputs():
    rng = np.random.default_rng(3)
    u = (rng.integers(1, 2 ** 32, 200_000, dtype=np.uint64).astype(np.float64) / 4294967295.0).astype(np.float32)
    ang = rng.uniform(-3.2, 6.4, 200_000).astype(np.float32)
    cz = np.clip(rng.uniform(-1.05, 1.05, 200_000), -1, 1).astype(np.float32)
    y = rng.uniform(-1, 1, 200_000).astype(np.float32)
    x = rng.uniform(-1, 1, 200_000).astype(np.float32)
    x[::7] *= 1e-4
    y[::11] *= 1e-5
    y[::1000] = 0.0
    x[::1500] = 0.0
    return u, ang, cz, y, x


def _ulps(got, want64):
    w = want64.astype(np.float32)
    ulp = np.abs(np.nextafter(w, np.float32(np.inf)) - w).astype(np.float64)
    return np.abs(got.astype(np.float64) - want64) / np.maximum(ulp, 1e-45)


def test_accuracy_against_libm(orc):
    u, ang, cz, y, x = _inputs()
    assert _ulps(orc.selftest_math(0, u), np.log(u.astype(np.float64))).max() <= 2.0
    # native_log of the delta-tracking step (table-driven cpm_native_logf)
    assert _ulps(orc.selftest_math(10, u), np.log(u.astype(np.float64))).max() <= 2.0
    nl = orc.selftest_math(10, np.array([0.0, 1.0], np.float32))
    assert np.isneginf(nl[0]) and nl[1] == 0.0
    assert np.abs(orc.selftest_math(1, ang) - np.sin(ang.astype(np.float64))).max() <= 1.5e-7
    assert np.abs(orc.selftest_math(2, ang) - np.cos(ang.astype(np.float64))).max() <= 1.5e-7
    assert np.abs(orc.selftest_math(3, cz) - np.arccos(cz.astype(np.float64))).max() <= 4e-7
    got = orc.selftest_math(4, y, x)
    assert np.abs(got - np.arctan2(y.astype(np.float64), x.astype(np.float64))).max() <= 5e-7
    # special values the path relies on
    sp = orc.selftest_math(0, np.array([0.0, 1.0], np.float32))
    assert np.isneginf(sp[0]) and sp[1] == 0.0
    assert orc.selftest_math(3, np.array([1.0], np.float32))[0] == 0.0
    # pow / cbrt / exp of rgb2lab (isc/cl/minmaxuniformgrid3dimportance.cl:174-175 through Inviwo's colorconversion.cl)
    rng = np.random.default_rng(5)
    base = rng.uniform(0.05, 1.2, 200_000).astype(np.float32)
    p24 = np.full_like(base, 2.4)
    assert _ulps(orc.selftest_math(7, base, p24), np.power(base.astype(np.float64), np.float64(np.float32(2.4)))).max() <= 8.0
    cb = np.concatenate([rng.uniform(1e-6, 8.0, 200_000), 10.0 ** rng.uniform(-30, 30, 20_000)]).astype(np.float32)
    assert _ulps(orc.selftest_math(8, cb), np.cbrt(cb.astype(np.float64))).max() <= 1.0
    ex = rng.uniform(-87, 87, 200_000).astype(np.float32)
    assert _ulps(orc.selftest_math(9, ex), np.exp(ex.astype(np.float64))).max() <= 3.0
    assert orc.selftest_math(8, np.array([0.0, 1.0, 8.0, 27.0], np.float32)).tolist() == [0.0, 1.0, 2.0, 3.0]


@pytest.mark.gpu
def test_host_and_device_agree_bitwise(orc, ctx, torch_cuda):
    torch = torch_cuda
    u, ang, cz, y, x = _inputs()
    cases = [(0, u, None), (10, u, None), (10, np.concatenate([np.float32(10.0) ** np.linspace(-44, 38, 50_000, dtype=np.float32), np.zeros(3, np.float32)]), None),
             (1, ang, None), (2, ang, None), (3, cz, None), (4, y, x),
             (5, np.arange(256, dtype=np.float32), None), (6, np.arange(65536, dtype=np.float32), None)]
    rng = np.random.default_rng(5)
    base = rng.uniform(0.0, 1.5, 200_000).astype(np.float32)
    cases += [(7, base, np.full_like(base, 2.4)), (7, base, rng.uniform(-3, 3, 200_000).astype(np.float32)),
              (8, np.concatenate([rng.uniform(0, 8.0, 200_000), 10.0 ** rng.uniform(-44, 36, 20_000)]).astype(np.float32), None),
              (9, rng.uniform(-90, 90, 200_000).astype(np.float32), None)]
    for fn, a, b in cases:
        da = torch.from_numpy(a).cuda()
        db = torch.from_numpy(b).cuda() if b is not None else None
        out = torch.empty_like(da)
        ctx.selftest_math(fn, da, db, out)
        ctx.sync()
        want = orc.selftest_math(fn, a, b)
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32)), f"fn {fn}"
