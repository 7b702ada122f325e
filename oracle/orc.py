"""ctypes loader of the CPU oracle (oracle/liboracle.so) and of the reference-derived MWC64X
checker (oracle/_ref/libmwc64x_ref.so).  TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_lib = None
_ref = None


class OrcVolume(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dims", C.c_int * 3), ("format", C.c_int), ("scale", C.c_float),
                ("offset", C.c_float)]


class OrcTraceParams(C.Structure):
    _fields_ = [
        ("aabb_min", C.c_float * 3), ("aabb_max", C.c_float * 3), ("material", C.c_float * 4),
        ("phase_function", C.c_int32), ("step_size", C.c_float), ("max_interactions", C.c_int32),
        ("photon_offset", C.c_int32), ("total_photons", C.c_int32), ("n_light_samples", C.c_int32),
        ("flags", C.c_uint32),
    ]


def build():
    subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        so = HERE / "liboracle.so"
        if not so.exists():
            build()
        _lib = C.CDLL(str(so))
        _lib.orc_sample_volume.restype = C.c_float
        _lib.orc_sample_tf_alpha.restype = C.c_float
        _lib.orc_trace_photons.restype = C.c_ulonglong
    return _lib


def ref():
    """The reference's own MWC64X kernels compiled for the host, or None when not built."""
    global _ref
    if _ref is None:
        so = HERE / "_ref" / "libmwc64x_ref.so"
        if not so.exists():
            return None
        _ref = C.CDLL(str(so))
    return _ref


_ref_libs = {}


def ref_lib(name):
    """oracle/_ref/lib<name>.so -- the reference's own sources compiled for the host (oracle/Makefile `ref`), or None
    when it was not built (no /root/reference at build time and no prebuilt copy)"""
    if name not in _ref_libs:
        so = HERE / "_ref" / f"lib{name}.so"
        _ref_libs[name] = C.CDLL(str(so)) if so.exists() else None
    return _ref_libs[name]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def volume(data: np.ndarray, scale=1.0, offset=0.0) -> OrcVolume:
    """data: (nz, ny, nx) C-contiguous array of uint8 / uint16 / float32"""
    fmt = {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1, np.dtype(np.float32): 2}[data.dtype]
    v = OrcVolume()
    v.data = data.ctypes.data
    v.dims[:] = [data.shape[2], data.shape[1], data.shape[0]]
    v.format = fmt
    v.scale = scale
    v.offset = offset
    v._keep = data
    return v


def rng_host_base_offsets(seed, n):
    out = np.zeros((n, 2), np.uint32)
    lib().orc_rng_host_base_offsets(C.c_uint32(seed), _ptr(out), C.c_size_t(n))
    return out


def rng_seed_streams(state, gap=1 << 40, first_stream=0):
    lib().orc_rng_seed_streams(_ptr(state), C.c_size_t(state.shape[0]), C.c_uint64(gap), C.c_uint64(first_stream))
    return state


def rng_uniform(state, per_stream=1):
    out = np.empty((state.shape[0], per_stream), np.float32)
    lib().orc_rng_uniform(_ptr(state), C.c_size_t(state.shape[0]), int(per_stream), _ptr(out))
    return out


def sample_uniform2d(nx, ny, n):
    out = np.empty((n, 4), np.float32)
    lib().orc_sample_uniform2d(C.c_float(nx), C.c_float(ny), int(n), _ptr(out))
    return out


def light_sample_directional(samples, radiance, direction, origin, u, v, area):
    n = samples.shape[0]
    out = np.empty((n, 8), np.float32)
    lib().orc_light_sample_directional(_ptr(samples), _f3(radiance), _f3(direction), _f3(origin), _f3(u), _f3(v),
                                       C.c_float(area), n, _ptr(out))
    return out


def light_sample_point(samples, radiance, pos):
    n = samples.shape[0]
    out = np.empty((n, 8), np.float32)
    lib().orc_light_sample_point(_ptr(samples), _f3(radiance), _f3(pos), n, _ptr(out))
    return out


def light_mesh_intersect(vertices, indices, light_samples):
    n = light_samples.shape[0]
    out = np.empty((n, 2), np.float32)
    lib().orc_light_mesh_intersect(_ptr(vertices), _ptr(indices), int(indices.size), _ptr(light_samples), n, _ptr(out))
    return out


def fit_light_plane(points, plane_point, plane_normal):
    out = np.empty(9, np.float32)
    pts = np.ascontiguousarray(points, np.float32)
    lib().orc_fit_light_plane(_ptr(pts), int(pts.shape[0]), _f3(plane_point), _f3(plane_normal), _ptr(out))
    return out[0:3].copy(), out[3:6].copy(), out[6:9].copy()


def convex_hull2d(points_xy):
    pts = np.ascontiguousarray(points_xy, np.float32)
    hull = np.zeros((2 * len(pts) + 2, 2), np.float32)
    n = lib().orc_convex_hull2d(_ptr(pts), int(len(pts)), _ptr(hull))
    return hull[:n].copy()


def trace_params(**kw) -> OrcTraceParams:
    p = OrcTraceParams()
    p.aabb_min[:] = [float(x) for x in kw.get("aabb_min", (0, 0, 0))]
    p.aabb_max[:] = [float(x) for x in kw.get("aabb_max", (1, 1, 1))]
    p.material[:] = [float(x) for x in kw.get("material", (0, 0, 0, 0))]
    p.phase_function = kw.get("phase", 0)
    p.step_size = kw.get("step_size", 1.0 / 256)
    p.max_interactions = kw.get("max_interactions", 1)
    p.photon_offset = kw.get("photon_offset", 0)
    p.n_light_samples = kw["n_light_samples"]
    p.total_photons = kw.get("total_photons", None) or kw["n_light_samples"]
    p.flags = kw.get("flags", 0)
    return p


def trace_photons(vol: OrcVolume, tf_rgba, params: OrcTraceParams, light_samples, isect, photons, rng,
                  recompute=None, n_recompute=0, n_threads=0) -> int:
    return int(lib().orc_trace_photons(C.byref(vol), _ptr(tf_rgba), int(tf_rgba.shape[0]), C.byref(params),
                                       _ptr(light_samples), _ptr(isect), _ptr(recompute), int(n_recompute),
                                       _ptr(photons), _ptr(rng), int(n_threads)))


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> int:
    """OpenMP threads of every later oracle call (torchrun exports OMP_NUM_THREADS=1); returns the count in effect"""
    lib().orc_set_num_threads(int(n))
    return num_threads()


def selftest_math(fn, x, y=None):
    out = np.empty_like(x)
    lib().orc_selftest_math(int(fn), _ptr(x), _ptr(y), _ptr(out), C.c_size_t(x.size))
    return out


def radix_sort(keys, values=None, max_bits=0):
    lib().orc_radix_sort_u32(_ptr(keys), _ptr(values), C.c_size_t(keys.size), C.c_uint(max_bits))


def merge_sort(keys, values=None):
    lib().orc_merge_sort_u32(_ptr(keys), _ptr(values), C.c_size_t(keys.size))


def count_below(data, threshold):
    lib().orc_count_below.restype = C.c_longlong
    return int(lib().orc_count_below(_ptr(data), C.c_size_t(data.size), C.c_uint32(threshold)))


def _fN(v, n):
    return (C.c_float * n)(*[float(x) for x in v])


def _i3(v):
    return (C.c_int * 3)(*[int(x) for x in v])


def volume_minmax(vol_np, region, scale=1.0, offset=0.0):
    nz, ny, nx = vol_np.shape
    od = [-(-nx // region), -(-ny // region), -(-nz // region)]
    out = np.empty((od[2], od[1], od[0], 2), np.uint16)
    v = volume(vol_np, scale, offset)
    lib().orc_volume_minmax(C.byref(v), int(region), _ptr(out))
    return out


def volume_diff_bricks(a_np, b_np, region, scaling=1.0, rmin=0.0, rmax=1.0):
    nz, ny, nx = a_np.shape
    od = [-(-nx // region), -(-ny // region), -(-nz // region)]
    out = np.empty((od[2], od[1], od[0]), np.float32)
    va, vb = volume(a_np), volume(b_np)
    lib().orc_volume_diff_bricks(C.byref(va), C.byref(vb), int(region), C.c_double(scaling), C.c_double(rmin),
                                 C.c_double(rmax), _ptr(out))
    return out


def classify_importance(minmax, positions, colors, weights, incremental, prev=None, diff=None):
    n = minmax.size // 2
    out = np.empty(n, np.float32)
    lib().orc_classify_importance(_ptr(minmax), _ptr(prev), _ptr(diff), n, _ptr(positions), _ptr(colors),
                                  int(positions.size), _fN(weights, 4), int(bool(incremental)), _ptr(out))
    return out


def detect_invalid(grid, grid_dims, cell_size, tex2idx, photons, photon_offset, light_samples, isect, n_light_samples,
                   max_interactions, total_photons, importances, equal_importance=False, percentage=100, iteration=0,
                   fix_exit=False):
    lib().orc_detect_invalid(_ptr(grid), _i3(grid_dims), _f3(cell_size), _fN(tex2idx, 16), _ptr(photons),
                             int(photon_offset), _ptr(light_samples), _ptr(isect), int(n_light_samples),
                             int(max_interactions), int(total_photons), _ptr(importances), int(bool(equal_importance)),
                             int(percentage), int(iteration), int(bool(fix_exit)))


def hash_light_samples(light_samples, isect, n_src, ids, cell_size, n_blocks, out, out_offset=0):
    lib().orc_hash_light_samples(_ptr(light_samples), _ptr(isect), int(n_src), _ptr(ids), int(ids.size), _f3(cell_size),
                                 _i3(n_blocks), _ptr(out), int(out_offset))


def build_cell_ranges(keys, n_cells):
    s, e = np.empty(n_cells, np.uint32), np.empty(n_cells, np.uint32)
    lib().orc_build_cell_ranges(_ptr(keys), C.c_size_t(keys.size), C.c_uint32(n_cells), _ptr(s), _ptr(e))
    return s, e


def splat(vol64, channels, tex2idx, idx2tex, out_dims, photons, indices, n, per_interaction, n_interactions, radius,
          scale, multiplier=1.0):
    lib().orc_splat(_ptr(vol64), int(channels), _fN(tex2idx, 16), _fN(idx2tex, 16), _i3(out_dims), _ptr(photons),
                    _ptr(indices), int(n), int(per_interaction), int(n_interactions), C.c_float(radius), C.c_float(scale),
                    C.c_float(multiplier))


# -- photon-map gather (parity unpinned: own restatement) ----------------------------------------------
class GatherParams(C.Structure):
    """layout shared by cpm_gather_params (include/cpm_b200.h) and orc_gather_params"""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("cam_origin", C.c_float * 3), ("cam_dir00", C.c_float * 3),
                ("cam_du", C.c_float * 3), ("cam_dv", C.c_float * 3), ("aabb_min", C.c_float * 3),
                ("aabb_max", C.c_float * 3), ("step", C.c_float), ("radius", C.c_float), ("scale", C.c_float),
                ("sigma_scale", C.c_float), ("grid_dims", C.c_int32 * 3)]


def photon_cell_keys(photons, grid_dims):
    n = photons.shape[0]
    keys = np.empty(n, np.uint32)
    lib().orc_photon_cell_keys(_ptr(photons), C.c_size_t(n), _i3(grid_dims), _ptr(keys))
    return keys


def gather_points(params, photons, points):
    pts = np.ascontiguousarray(points, np.float32)
    out = np.empty_like(pts)
    lib().orc_gather_points(C.byref(params), _ptr(photons), C.c_size_t(photons.shape[0]), _ptr(pts), int(pts.shape[0]),
                            _ptr(out))
    return out


def gather_raymarch(vol, tf_rgba, params, photons):
    img = np.empty((params.height, params.width, 4), np.float32)
    lib().orc_gather_raymarch(C.byref(vol), _ptr(tf_rgba), int(tf_rgba.shape[0]), C.byref(params), _ptr(photons),
                              C.c_size_t(photons.shape[0]), _ptr(img))
    return img


def raycast_light_volume(vol, tf_rgba, params, light_volume, lv_dims, channels=1):
    img = np.empty((params.height, params.width, 4), np.float32)
    lib().orc_raycast_light_volume(C.byref(vol), _ptr(tf_rgba), int(tf_rgba.shape[0]), C.byref(params), _ptr(light_volume),
                                   _i3(lv_dims), int(channels), _ptr(img))
    return img


# -- view importance + importance-driven sample generator ------------------------------------------------
def view_importance(minmax, grid_dims, cell_size, tex2idx, idx2tex, entry, exit_, tf_min, tf_max):
    h, w = entry.shape[0], entry.shape[1]
    out = np.empty((h, w), np.float32)
    lib().orc_view_importance(_ptr(minmax), _i3(grid_dims), _f3(cell_size), _fN(tex2idx, 16), _fN(idx2tex, 16), _ptr(entry),
                              _ptr(exit_), int(w), int(h), C.c_float(tf_min), C.c_float(tf_max), _ptr(out))
    return out


def sample_importance2d(importance, floor_value, uniform_samples):
    h, w = importance.shape
    out = np.empty_like(uniform_samples)
    lib().orc_sample_importance2d(_ptr(importance), int(w), int(h), C.c_float(floor_value), _ptr(uniform_samples),
                                  int(uniform_samples.shape[0]), _ptr(out))
    return out
