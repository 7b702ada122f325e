"""The CPU light-plane fit: three implementations that must agree bit for bit --
  * the reference's own lcl/convexhull2d.cpp, orientedboundingbox2d.cpp, pointplaneprojection.cpp compiled where they lie
    (oracle/_ref/libgeometry_ref.so; golden vectors of it: tests/golden/lightplane.json, tools/make_golden.py geometry),
  * the oracle's restatement (oracle/orc_emission.c),
  * the drop-in host layer's (host/processors.cpp: geometry::), which the product path runs.
No GPU needed: the fit is CPU code in the reference too."""
import ctypes as C
import importlib
import json
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, PKG_NAME, ROOT

sys.path.insert(0, str(ROOT / "tools"))


def _unhex(xs, shape):
    return np.array([float.fromhex(x) for x in xs], np.float32).reshape(shape)


def _cases():
    g = json.loads((GOLDEN / "lightplane.json").read_text())
    for c in g["cases"]:
        yield (_unhex(c["points"], (-1, 3)), _unhex(c["plane_point"], 3), _unhex(c["normal"], 3), _unhex(c["fit"], 9),
               _unhex(c["hull_xy"], (-1, 2)))


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def host():
    return importlib.import_module(PKG_NAME + ".host")


def _host_fit(host, pts, pp, d):
    out = (C.c_float * 9)()
    p = np.ascontiguousarray(pts, np.float32)
    rc = host.lib().cpmh_fit_light_plane(p.ctypes.data_as(C.c_void_p), len(p), (C.c_float * 3)(*pp), (C.c_float * 3)(*d), out)
    assert rc == 0
    return np.array(out[:], np.float32)


def _host_hull(host, xy):
    p = np.ascontiguousarray(xy, np.float32)
    hull = np.zeros((2 * len(p) + 2, 2), np.float32)
    n = host.lib().cpmh_convex_hull2d(p.ctypes.data_as(C.c_void_p), len(p), hull.ctypes.data_as(C.c_void_p))
    assert n >= 0
    return hull[:n]


def test_oracle_fit_equals_reference_golden(orc):
    n = 0
    for pts, pp, d, fit, hull in _cases():
        o, u, v = orc.fit_light_plane(pts, pp, d)
        got = np.concatenate([o, u, v])
        assert np.array_equal(_bits(got), _bits(fit)) or (np.isnan(got) == np.isnan(fit)).all() and np.array_equal(
            _bits(got)[~np.isnan(fit)], _bits(fit)[~np.isnan(fit)]), n
        assert np.array_equal(_bits(orc.convex_hull2d(pts[:, :2])), _bits(hull)), n
        n += 1
    assert n >= 50


def test_host_fit_equals_reference_golden(host):
    n = 0
    for pts, pp, d, fit, hull in _cases():
        got = _host_fit(host, pts, pp, d)
        ok = ~np.isnan(fit)
        assert (np.isnan(got) == np.isnan(fit)).all() and np.array_equal(_bits(got)[ok], _bits(fit)[ok]), n
        assert np.array_equal(_bits(_host_hull(host, pts[:, :2])), _bits(hull)), n
        n += 1


def test_live_reference_library_on_fresh_inputs(orc, host):
    """oracle == host == the reference build on 1000 fresh random inputs (skipped where oracle/_ref was not built)"""
    ref = orc.ref_lib("geometry_ref")
    if ref is None:
        pytest.skip("oracle/_ref/libgeometry_ref.so not built (reference tree absent)")
    rng = np.random.default_rng(23)
    cube = np.array([[x, y, z] for z in (0.0, 1.0) for y in (0.0, 1.0) for x in (0.0, 1.0)], np.float32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    for it in range(1000):
        if it % 2 == 0:
            pts = cube
        else:
            pts = np.ascontiguousarray(rng.uniform(-1, 2, (int(rng.integers(3, 40)), 3)).astype(np.float32))
        d = rng.normal(size=3)
        d = (d / np.linalg.norm(d)).astype(np.float32)
        pp = (np.float32([0.5, 0.5, 0.5]) - 2 * d).astype(np.float32)
        want = np.zeros(9, np.float32)
        ref.ref_fit_plane_aligned_obb2d(P(pts), len(pts), P(pp), P(d), P(want))
        o, u, v = orc.fit_light_plane(pts, pp, d)
        assert np.array_equal(_bits(np.concatenate([o, u, v])), _bits(want)), it
        assert np.array_equal(_bits(_host_fit(host, pts, pp, d)), _bits(want)), it


def test_golden_file_is_what_the_generator_writes(orc):
    """tests/golden/lightplane.json is reproducible from the committed generator (where the reference build exists)"""
    if orc.ref_lib("geometry_ref") is None:
        pytest.skip("oracle/_ref/libgeometry_ref.so not built (reference tree absent)")
    import make_golden
    ref = orc.ref_lib("geometry_ref")
    cases = list(_cases())
    P = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    for (pts, pp, d), (gp, gpp, gd, gfit, _) in zip(make_golden.geometry_cases(), cases):
        assert np.array_equal(_bits(pts), _bits(gp)) and np.array_equal(_bits(d), _bits(gd))
        fit = np.zeros(9, np.float32)
        ref.ref_fit_plane_aligned_obb2d(P(np.ascontiguousarray(pts)), len(pts), P(pp), P(d), P(fit))
        ok = ~np.isnan(gfit)
        assert np.array_equal(_bits(fit)[ok], _bits(gfit)[ok])
