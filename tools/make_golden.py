"""Generate tests/golden/*.json.

mwc64x.json comes from the REFERENCE's own MWC64X sources compiled for the host
(oracle/_ref/libmwc64x_ref.so, built by oracle/Makefile from /root/reference).  Run in the
build container (the reference tree does not exist on the GPU box); the JSON is committed.
"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import orc  # noqa: E402


def mwc64x():
    ref = orc.ref()
    if ref is None:
        raise SystemExit("oracle/_ref/libmwc64x_ref.so missing: run `make -C oracle ref` first")
    out = {"source": "reference rng/cl/{skip_mwc,random,randstategen,randomnumbergenerator}.cl compiled via oracle/Makefile"}
    # (a) the reference pipeline: srand(0); base_i = rand(); GenerateRandomState; 3 x random_01
    n = 64
    base = orc.rng_host_base_offsets(0, n)
    state = base.copy()
    ref.ref_generate_random_state(state.ctypes.data_as(C.c_void_p), n)
    out["seed0_base"] = base[:, 0].tolist()
    out["seed0_state"] = state.tolist()
    s = state.copy()
    draws = []
    for _ in range(3):
        o = np.zeros(n, np.float32)
        ref.ref_random_number_generator(s.ctypes.data_as(C.c_void_p), n, o.ctypes.data_as(C.c_void_p))
        draws.append([float(x).hex() for x in o])
    out["seed0_random01_hex"] = draws
    out["seed0_state_after3"] = s.tolist()
    # (b) per-stream gap variant with awkward bases
    bases = np.array([[0, 0], [1, 0], [2147483647, 0], [4294967295, 0], [123456789, 0]], np.uint32)
    for gap in (1, 12345, 1 << 40, (1 << 63) + 5):
        st = bases.copy()
        ref.ref_generate_per_stream_random_state(st.ctypes.data_as(C.c_void_p), C.c_uint64(gap), len(st))
        out[f"gap_{gap}"] = st.tolist()
    out["gap_bases"] = bases[:, 0].tolist()
    # (c) raw steps from corner states
    steps = []
    for x, c in ((0, 1), (1, 0), (0xFFFFFFFF, 0xFFFEB81A), (0x12345678, 0x9ABCDEF0 % 4294883355)):
        xx, cc = C.c_uint32(x), C.c_uint32(c)
        seq = []
        for _ in range(4):
            ref.ref_step(C.byref(xx), C.byref(cc))
            seq.append([xx.value, cc.value])
        steps.append({"x": x, "c": c, "seq": seq})
    out["steps"] = steps
    (ROOT / "tests" / "golden").mkdir(parents=True, exist_ok=True)
    (ROOT / "tests" / "golden" / "mwc64x.json").write_text(json.dumps(out, indent=1))
    print("wrote tests/golden/mwc64x.json")


def _hex(a):
    return [float(x).hex() for x in np.asarray(a, np.float32).reshape(-1)]


def geometry_cases():
    """seeded inputs of the light-plane fit: the unit-cube proxy under many directions (incl. axis-aligned ones, where
    hull points share x), random point clouds, degenerate sets (collinear, duplicates, < 4 points)"""
    rng = np.random.default_rng(11)
    cube = np.array([[x, y, z] for z in (0.0, 1.0) for y in (0.0, 1.0) for x in (0.0, 1.0)], np.float32)
    cases = []
    dirs = [(0, 0, 1), (1, 0, 0), (0, 1, 0), (0, -1, 0), (0.3, -0.5, 0.8), (-0.36, 0.48, 0.8), (0.5, -0.3, 0.81), (1, 1, 0),
            (1, 1, 1), (-1, 2, -3)]
    for k in range(30):
        d = np.array(dirs[k], np.float64) if k < len(dirs) else rng.normal(size=3)
        d = (d / np.linalg.norm(d)).astype(np.float32)
        cases.append((cube, (np.float32([0.5, 0.5, 0.5]) - 2 * d).astype(np.float32), d))
    for k in range(20):
        n = int(rng.integers(3, 24))
        pts = rng.uniform(-1, 2, (n, 3)).astype(np.float32)
        if k % 5 == 0:
            pts[n // 2:] = pts[: n - n // 2]            # duplicates
        if k % 7 == 0:
            pts = (pts[:1] + np.outer(np.linspace(0, 1, n), [1, 2, 3])).astype(np.float32)   # collinear
        d = rng.normal(size=3)
        d = (d / np.linalg.norm(d)).astype(np.float32)
        cases.append((np.ascontiguousarray(pts), rng.uniform(-1, 1, 3).astype(np.float32), d))
    return cases


def geometry():
    """tests/golden/lightplane.json from the reference's own lcl/{convexhull2d,orientedboundingbox2d,pointplaneprojection}.cpp
    (oracle/_ref/libgeometry_ref.so): plane fit, plus the 2-D hull and minimum rectangle of the projected points"""
    ref = orc.ref_lib("geometry_ref")
    if ref is None:
        raise SystemExit("oracle/_ref/libgeometry_ref.so missing: run `make -C oracle ref` first")
    P = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    out = {"source": "reference lcl/convexhull2d.cpp, lcl/orientedboundingbox2d.cpp, lcl/pointplaneprojection.cpp compiled "
                     "via oracle/Makefile (GLM / Inviwo Plane stand-ins: oracle/ref_shim/host)", "cases": []}
    for pts, pp, d in geometry_cases():
        fit = np.zeros(9, np.float32)
        ref.ref_fit_plane_aligned_obb2d(P(pts), len(pts), P(pp), P(d), P(fit))
        # the hull / rectangle of an arbitrary 2-D set: the points' xy
        xy = np.ascontiguousarray(pts[:, :2])
        hull = np.zeros((2 * len(pts) + 2, 2), np.float32)
        nh = ref.ref_convex_hull2d(P(xy), len(xy), P(hull))
        rect = np.zeros(6, np.float32)
        ref.ref_minimum_bounding_rectangle(P(hull), nh, P(rect))
        out["cases"].append({"points": _hex(pts), "plane_point": _hex(pp), "normal": _hex(d), "fit": _hex(fit),
                             "hull_xy": _hex(hull[:nh]), "rect_xy": _hex(rect)})
    (ROOT / "tests" / "golden" / "lightplane.json").write_text(json.dumps(out, indent=0))
    print("wrote tests/golden/lightplane.json")


def kernels():
    """tests/golden/ref_kernels.npz: the outputs of the reference's own OpenCL kernels (oracle/_ref/libcl_ref.so: the .cl
    files compiled for the host, strict IEEE evaluation) on the seeded cases of tests/ref_cases.py"""
    ref = orc.ref_lib("cl_ref")
    if ref is None:
        raise SystemExit("oracle/_ref/libcl_ref.so missing: run `make -C oracle ref` first")
    sys.path.insert(0, str(ROOT / "tests"))
    import ref_cases
    out = ref_cases.ref_outputs(ref)
    np.savez_compressed(ROOT / "tests" / "golden" / "ref_kernels.npz", **out)
    print("wrote tests/golden/ref_kernels.npz:", len(out), "arrays,", (ROOT / "tests" / "golden" / "ref_kernels.npz").stat().st_size, "bytes")


def photondata_cases():
    rng = np.random.default_rng(17)
    cases = []
    for k in range(40):
        n = int(rng.choice([256 * 256, 1024 * 1024, 2048 * 2048, 131072, 4 * 2048 * 2048]))
        cases.append((n, int(rng.integers(1, 9)), float(rng.uniform(1e-3, 0.1)), float(rng.uniform(0.3, 400.0)), int(rng.integers(0, 60)),
                      float(rng.choice([0.5, 0.7, 0.9, 1e-4, 1.0]))))
    cases.append((256 * 256, 1, float(np.float32(0.0153866)), float(np.float32(1.1447142425533318678080422119397)), 0, 0.5))
    dirs = rng.normal(size=(64, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dirs = np.vstack([dirs, [[0, 0, 1], [0, 0, -1], [1, 0, 0], [0, -1, 0], [0, 0, 1.0000001]]]).astype(np.float32)
    return cases, dirs


def photondata():
    """tests/golden/photondata.json from the reference's own ppm/photondata.cpp (oracle/_ref/libphotondata_ref.so)"""
    ref = orc.ref_lib("photondata_ref")
    if ref is None:
        raise SystemExit("oracle/_ref/libphotondata_ref.so missing: run `make -C oracle ref` first")
    ref.ref_photondata_progress.argtypes = [C.c_size_t, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.POINTER(C.c_double)]
    cases, dirs = photondata_cases()
    out = {"source": "reference ppm/photondata.cpp compiled via oracle/Makefile (GLM / Inviwo Buffer stand-ins: oracle/ref_shim/host)",
           "progress": [], "directions": []}
    for c in cases:
        o = (C.c_double * 4)()
        ref.ref_photondata_progress(*c, o)
        out["progress"].append({"in": [c[0], c[1], c[2].hex(), c[3].hex(), c[4], c[5].hex()], "out": [float(x).hex() for x in o]})
    k = (C.c_double * 4)()
    ref.ref_photondata_constants(k)
    out["constants"] = [float(x).hex() for x in k]
    for d in dirs:
        enc, dec = (C.c_float * 2)(), (C.c_float * 3)()
        ref.ref_photon_encode_direction((C.c_float * 3)(*d), enc)
        ref.ref_photon_decode_direction(enc, dec)
        out["directions"].append({"dir": _hex(d), "encoded": _hex(enc[:]), "decoded": _hex(dec[:])})
    (ROOT / "tests" / "golden" / "photondata.json").write_text(json.dumps(out, indent=0))
    print("wrote tests/golden/photondata.json")


if __name__ == "__main__":
    which = sys.argv[1:] or ["mwc64x", "geometry", "kernels", "photondata"]
    for w in which:
        {"mwc64x": mwc64x, "geometry": geometry, "kernels": kernels, "photondata": photondata}[w]()
