"""PhotonData parameters (progressive radius, scene radius, irradiance scale) and the host twin of the photon
direction codec (8a-1): the drop-in host layer against the reference's own ppm/photondata.cpp compiled for the host
(oracle/_ref/libphotondata_ref.so; golden vectors tests/golden/photondata.json, tools/make_golden.py photondata).
Double-precision libm results (pow) are compared bit for bit: both sides call the same std::pow on this machine and on the
GPU box; the float acos / atan2 / sin / cos of the codec likewise."""
import ctypes as C
import importlib
import json

import numpy as np
import pytest

from conftest import GOLDEN, PKG_NAME


@pytest.fixture(scope="module")
def host():
    return importlib.import_module(PKG_NAME + ".host")


def _gold():
    return json.loads((GOLDEN / "photondata.json").read_text())


def test_photondata_progress_equals_reference_golden(host):
    g = _gold()
    lib = host.lib()
    for c in g["progress"]:
        n, I, rr, sr, it, al = c["in"]
        out = (C.c_double * 4)()
        assert lib.cpmh_photondata_progress(int(n), int(I), float.fromhex(rr), float.fromhex(sr), int(it), float.fromhex(al), out) == 0
        assert [float(x).hex() for x in out] == c["out"], c["in"]
    # Knaus & Zwicker eq. 20 holds for what is pinned: r_{i+1} = r_i ((i + alpha) / (i + 1))^(1/3)
    out0, out1 = (C.c_double * 4)(), (C.c_double * 4)()
    lib.cpmh_photondata_progress(65536, 1, 0.02, 2.0, 5, 0.7, out0)
    lib.cpmh_photondata_progress(65536, 1, 0.02, 2.0, 6, 0.7, out1)
    assert abs(out1[0] / out0[0] - ((5 + 0.7) / 6.0) ** (1.0 / 3.0)) < 1e-15


def test_photon_direction_codec_equals_reference_golden(host):
    g = _gold()
    lib = host.lib()
    for d in g["directions"]:
        v = [float.fromhex(x) for x in d["dir"]]
        enc, dec = (C.c_float * 2)(), (C.c_float * 3)()
        lib.cpmh_photon_encode_direction((C.c_float * 3)(*v), enc)
        lib.cpmh_photon_decode_direction(enc, dec)
        assert [float(x).hex() for x in enc] == d["encoded"], d["dir"]
        assert [float(x).hex() for x in dec] == d["decoded"], d["dir"]


def test_reference_constants():
    k = [float.fromhex(x) for x in _gold()["constants"]]
    assert k[0] == float(np.float32(0.0153866)) and k[3] == 65536
    assert abs(k[2] - 1.0 / np.pi) < 1e-16                      # scaleToMakeLightPowerOfOneVisibleForDirectionalLightSource
    assert abs(k[1] - 0.5 * np.sqrt(12.0) * (1.1447142425533318678080422119397 / np.sqrt(3.0))) < 1e-6


def test_live_reference_library(host, orc):
    ref = orc.ref_lib("photondata_ref")
    if ref is None:
        pytest.skip("oracle/_ref/libphotondata_ref.so not built (reference tree absent)")
    ref.ref_photondata_progress.argtypes = [C.c_size_t, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.POINTER(C.c_double)]
    rng = np.random.default_rng(99)
    lib = host.lib()
    for _ in range(500):
        args = (int(rng.integers(1, 1 << 24)), int(rng.integers(1, 17)), float(rng.uniform(1e-4, 0.5)), float(rng.uniform(0.1, 1e3)),
                int(rng.integers(0, 200)), float(rng.uniform(1e-4, 1.0)))
        a, b = (C.c_double * 4)(), (C.c_double * 4)()
        ref.ref_photondata_progress(*args, a)
        lib.cpmh_photondata_progress(*args, b)
        assert list(a) == list(b), args
