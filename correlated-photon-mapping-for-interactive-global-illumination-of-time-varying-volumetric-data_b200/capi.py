"""ctypes binding of libcpm_b200.so (the C ABI declared in include/cpm_b200.h).

PyTorch is used only as the owner of device memory: every function takes CUDA tensors and
passes ``data_ptr()`` through.  There is no CPU path here -- if the shared library is missing
or no B200 is visible the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

_HERE = Path(__file__).resolve().parent
ROOT = _HERE.parent
LIB_PATH = Path(os.environ.get("CPM_B200_LIB", _HERE / "libcpm_b200.so"))   # override: tuning sweeps only
HEADER = ROOT / "include" / "cpm_b200.h"

CPM_FMT_U8, CPM_FMT_U16, CPM_FMT_F32 = 0, 1, 2
CPM_VOLUME_LINEAR, CPM_VOLUME_TEXTURE = 0, 1
CPM_TRACE_PROGRESSIVE, CPM_TRACE_NO_SINGLE_SCATTERING, CPM_TRACE_STATS = 1, 2, 4
CPM_PHASE_ISOTROPIC, CPM_PHASE_HENYEY_GREENSTEIN = 0, 1
FLT_MAX = 3.4028234663852886e38


class CpmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"cpm error {code}: {msg}")
        self.code = code


class TraceParams(C.Structure):
    _fields_ = [
        ("aabb_min", C.c_float * 3),
        ("aabb_max", C.c_float * 3),
        ("material", C.c_float * 4),
        ("phase_function", C.c_int32),
        ("step_size", C.c_float),
        ("max_interactions", C.c_int32),
        ("photon_offset", C.c_int32),
        ("total_photons", C.c_int32),
        ("n_light_samples", C.c_int32),
        ("flags", C.c_uint32),
        ("opacity_bound", C.c_void_p),
        ("bound_cell_log2", C.c_int32),
        ("reserved_", C.c_int32),
        ("opacity_bound_tex", C.c_void_p),
    ]


_lib = None


def declared_symbols() -> list[str]:
    """Every CPM_API function name declared in include/cpm_b200.h."""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"CPM_API\s+[\w\s\*]+?\b(cpm_\w+)\s*\(", text)))


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        _lib = C.CDLL(str(LIB_PATH))
        _lib.cpm_last_error.restype = C.c_char_p
        _lib.cpm_last_error.argtypes = [C.c_void_p]
        _lib.cpm_version.restype = C.c_char_p
        _lib.cpm_ctx_stream.restype = C.c_void_p
        _lib.cpm_ctx_stream.argtypes = [C.c_void_p]
        _lib.cpm_ctx_launch_count.restype = C.c_uint64
        _lib.cpm_ctx_launch_count.argtypes = [C.c_void_p, C.c_int]
        _lib.cpm_ctx_destroy.restype = None
        _lib.cpm_ctx_destroy.argtypes = [C.c_void_p]
        _lib.cpm_volume_destroy.restype = None
        _lib.cpm_volume_destroy.argtypes = [C.c_void_p, C.c_void_p]
        _lib.cpm_comm_destroy.restype = None
        _lib.cpm_comm_destroy.argtypes = [C.c_void_p]
        _lib.cpm_comm_transport.restype = C.c_char_p
        _lib.cpm_comm_transport.argtypes = [C.c_void_p]
        _lib.cpm_comm_rank.argtypes = [C.c_void_p]
        _lib.cpm_comm_world.argtypes = [C.c_void_p]
        _lib.cpm_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        _lib.cpm_comm_split.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        _lib.cpm_allreduce_lightvol.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        _lib.cpm_allreduce_lightvol_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        _lib.cpm_allreduce_lightvol_end.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        _lib.cpm_allgather_photons.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        _lib.cpm_allgather_volume.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        _lib.cpm_comm_barrier.argtypes = [C.c_void_p]
    return _lib


def _p(t):
    """device (or host) pointer of a tensor / None"""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class Volume:
    def __init__(self, ctx, handle, data, dims, fmt, layout):
        self.ctx, self.handle, self.data, self.dims, self.fmt, self.layout = ctx, handle, data, dims, fmt, layout

    def update(self, data):
        self.ctx._check(lib().cpm_volume_update(self.ctx.h, self.handle, _p(data)))
        self.data = data

    def destroy(self):
        if self.handle:
            lib().cpm_volume_destroy(self.ctx.h, self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Context:
    """One per GPU.  Mirrors the role of OpenCL::getPtr()->getQueue() in the reference."""

    def __init__(self, device: int = 0, stream=None):
        self.h = C.c_void_p()
        rc = lib().cpm_ctx_create(int(device), C.c_void_p(stream or 0), C.byref(self.h))
        if rc != 0:
            raise CpmError(rc, lib().cpm_last_error(None).decode())
        self.device = device

    def close(self):
        if self.h:
            lib().cpm_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise CpmError(rc, lib().cpm_last_error(self.h).decode())

    def sync(self):
        self._check(lib().cpm_ctx_sync(self.h))

    @property
    def stream(self):
        return lib().cpm_ctx_stream(self.h)

    def launch_count(self, reset=False):
        return int(lib().cpm_ctx_launch_count(self.h, int(reset)))

    # -- RNG ---------------------------------------------------------------------------
    def rng_seed_streams(self, state, n=None, gap=1 << 40, first_stream=0):
        n = state.numel() // 2 if n is None else n
        self._check(lib().cpm_rng_seed_streams(self.h, _p(state), C.c_size_t(n), C.c_uint64(gap), C.c_uint64(first_stream)))

    def rng_uniform(self, state, out, per_stream=1):
        n = state.numel() // 2
        self._check(lib().cpm_rng_uniform(self.h, _p(state), C.c_size_t(n), int(per_stream), _p(out)))

    # -- emission ----------------------------------------------------------------------
    def sample_uniform2d(self, nx, ny, n, out):
        self._check(lib().cpm_sample_uniform2d(self.h, C.c_float(nx), C.c_float(ny), int(n), _p(out)))

    def light_sample_directional(self, samples, radiance, direction, origin, u, v, area, n, out):
        self._check(lib().cpm_light_sample_directional(self.h, _p(samples), _f3(radiance), _f3(direction), _f3(origin),
                                                       _f3(u), _f3(v), C.c_float(area), int(n), _p(out)))

    def light_sample_point(self, samples, radiance, position, n, out):
        self._check(lib().cpm_light_sample_point(self.h, _p(samples), _f3(radiance), _f3(position), int(n), _p(out)))

    def light_mesh_intersect(self, vertices, indices, n_indices, light_samples, n, out):
        self._check(lib().cpm_light_mesh_intersect(self.h, _p(vertices), _p(indices), int(n_indices), _p(light_samples),
                                                   int(n), _p(out)))

    # -- volumes -----------------------------------------------------------------------
    def volume_create(self, data, dims, fmt, scale=1.0, offset=0.0, layout=CPM_VOLUME_LINEAR) -> Volume:
        h = C.c_void_p()
        d = (C.c_int * 3)(*[int(x) for x in dims])
        self._check(lib().cpm_volume_create(self.h, _p(data), d, int(fmt), C.c_float(scale), C.c_float(offset),
                                            int(layout), C.byref(h)))
        return Volume(self, h, data, tuple(dims), fmt, layout)

    # -- opacity-bound grid of the tracer -----------------------------------------------
    def volume_value_range(self, vol: Volume, cell_log2, out):
        """out: float tensor of 2 * prod(bound_grid_dims(vol.dims, cell_log2)); returns the grid dims"""
        od = (C.c_int * 3)()
        self._check(lib().cpm_volume_value_range(self.h, vol.handle, int(cell_log2), _p(out), od))
        return tuple(od)

    def opacity_bound(self, value_range, n_cells, tf_rgba, out, scale=1.0, offset=0.0):
        self._check(lib().cpm_opacity_bound(self.h, _p(value_range), C.c_size_t(n_cells), C.c_float(scale),
                                            C.c_float(offset), _p(tf_rgba), int(tf_rgba.numel() // 4), _p(out)))

    def opacity_bound_clearance(self, bound, grid_dims, max_radius=8):
        self._check(lib().cpm_opacity_bound_clearance(self.h, _p(bound), (C.c_int * 3)(*[int(x) for x in grid_dims]),
                                                      int(max_radius)))

    def bound_texture(self, grid_dims, bound=None):
        """the bound grid as a 3-D texture (cpm_bound_tex_*): returns a BoundTexture, filled from `bound` if given"""
        t = BoundTexture(self, grid_dims)
        if bound is not None:
            t.update(bound)
        return t

    # -- tracer ------------------------------------------------------------------------
    def trace_photons(self, vol: Volume, tf_rgba, params: TraceParams, light_samples, intersections, photons,
                      rng_state, recompute_index=None, n_recompute=0, collision_tests=None):
        self._check(lib().cpm_trace_photons(self.h, vol.handle, _p(tf_rgba), int(tf_rgba.numel() // 4), C.byref(params),
                                            _p(light_samples), _p(intersections), _p(recompute_index),
                                            int(n_recompute), _p(photons), _p(rng_state), _p(collision_tests)))


class BoundTexture:
    def __init__(self, ctx, grid_dims):
        self.ctx = ctx
        self.handle = C.c_void_p()
        ctx._check(lib().cpm_bound_tex_create(ctx.h, (C.c_int * 3)(*[int(x) for x in grid_dims]), C.byref(self.handle)))

    def update(self, bound):
        self.ctx._check(lib().cpm_bound_tex_update(self.ctx.h, self.handle, _p(bound)))

    def close(self):
        if self.handle:
            lib().cpm_bound_tex_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def rng_host_base_offsets(seed: int, n: int):
    """glibc srand(seed)/rand() base offsets as a (n,2) uint32 numpy array (host, synchronous)."""
    import numpy as np
    out = np.zeros((n, 2), dtype=np.uint32)
    rc = lib().cpm_rng_host_base_offsets(C.c_uint32(seed), out.ctypes.data_as(C.c_void_p), C.c_size_t(n))
    if rc != 0:
        raise CpmError(rc, "cpm_rng_host_base_offsets")
    return out


def make_trace_params(n_light_samples, total_photons=None, photon_offset=0, max_interactions=1, step_size=1.0 / 256,
                      aabb_min=(0, 0, 0), aabb_max=(1, 1, 1), phase=CPM_PHASE_ISOTROPIC, material=(0, 0, 0, 0),
                      flags=0, opacity_bound=None, bound_cell_log2=3, opacity_bound_tex=None) -> TraceParams:
    p = TraceParams()
    p.aabb_min[:] = [float(x) for x in aabb_min]
    p.aabb_max[:] = [float(x) for x in aabb_max]
    p.material[:] = [float(x) for x in material]
    p.phase_function = phase
    p.step_size = step_size
    p.max_interactions = max_interactions
    p.photon_offset = photon_offset
    p.total_photons = n_light_samples if total_photons is None else total_photons
    p.n_light_samples = n_light_samples
    p.flags = flags
    p.opacity_bound = opacity_bound.data_ptr() if opacity_bound is not None else None
    p.bound_cell_log2 = bound_cell_log2
    p.opacity_bound_tex = opacity_bound_tex.handle if opacity_bound_tex is not None else None
    return p


def bound_grid_dims(dims, cell_log2):
    """dims of the opacity-bound grid of the tracer: (dims >> cell_log2) + 1"""
    od = (C.c_int * 3)()
    rc = lib().cpm_bound_grid_dims((C.c_int * 3)(*[int(x) for x in dims]), int(cell_log2), od)
    if rc != 0:
        raise CpmError(rc, "cpm_bound_grid_dims")
    return tuple(od)


def _selftest_math(self, fn, x, y, out):
    self._check(lib().cpm_selftest_math(self.h, int(fn), _p(x), _p(y), _p(out), C.c_size_t(x.numel())))


Context.selftest_math = _selftest_math


# -- selection / sort --------------------------------------------------------------------------
def _threshold(self, data, threshold, out):
    self._check(lib().cpm_threshold_u32(self.h, _p(data), C.c_uint32(threshold), C.c_size_t(data.numel()), _p(out)))


def _iota(self, out):
    self._check(lib().cpm_iota_u32(self.h, _p(out), C.c_size_t(out.numel())))


def _reduce_sum_i32(self, data):
    r = C.c_longlong(0)
    self._check(lib().cpm_reduce_sum_i32(self.h, _p(data), C.c_size_t(data.numel()), C.byref(r)))
    return r.value


def _select_below(self, data, threshold, ids_out, n=None):
    r = C.c_longlong(0)
    n = data.numel() if n is None else n
    self._check(lib().cpm_select_below(self.h, _p(data), C.c_size_t(n), C.c_uint32(threshold), _p(ids_out), C.byref(r)))
    return r.value


def _count_below(self, data, threshold, iota_out=None, n=None):
    r = C.c_longlong(0)
    n = data.numel() if n is None else n
    self._check(lib().cpm_count_below(self.h, _p(data), C.c_size_t(n), C.c_uint32(threshold), _p(iota_out), C.byref(r)))
    return r.value


def _radix_sort(self, keys, values, tmp_keys, tmp_values, n=None, max_bits=0):
    n = keys.numel() if n is None else n
    self._check(lib().cpm_radix_sort_u32(self.h, _p(keys), _p(values), C.c_size_t(n), C.c_uint(max_bits), _p(tmp_keys),
                                         _p(tmp_values)))


Context.threshold = _threshold
Context.iota = _iota
Context.reduce_sum_i32 = _reduce_sum_i32
Context.count_below = _count_below
Context.select_below = _select_below
Context.radix_sort = _radix_sort


# -- detector / grids / splat ------------------------------------------------------------------
CPM_DETECT_FIX_EXIT = 1
CPM_TRACE_LANE_REFILL = 8


def _fN(v, n):
    return (C.c_float * n)(*[float(x) for x in v])


def _i3(v):
    return (C.c_int * 3)(*[int(x) for x in v])


def texture_to_index_matrix(dims):
    """column-major 4x4: index = tex * dim - 0.5 (Inviwo StructuredCoordinateTransformer)"""
    m = [0.0] * 16
    m[0], m[5], m[10], m[15] = float(dims[0]), float(dims[1]), float(dims[2]), 1.0
    m[12] = m[13] = m[14] = -0.5
    return m


def index_to_texture_matrix(dims):
    m = [0.0] * 16
    m[0], m[5], m[10], m[15] = 1.0 / dims[0], 1.0 / dims[1], 1.0 / dims[2], 1.0
    m[12], m[13], m[14] = 0.5 / dims[0], 0.5 / dims[1], 0.5 / dims[2]
    import numpy as np
    return [float(np.float32(x)) for x in m]


def _detect_invalid(self, grid, grid_dims, cell_size, tex2idx, photons, photon_offset, light_samples, intersections,
                    n_light_samples, max_interactions, total_photons, importances, equal_importance=False,
                    percentage=100, iteration=0, flags=0):
    self._check(lib().cpm_detect_invalid(self.h, _p(grid), _i3(grid_dims), _f3(cell_size), _fN(tex2idx, 16), _p(photons),
                                         int(photon_offset), _p(light_samples), _p(intersections), int(n_light_samples),
                                         int(max_interactions), int(total_photons), _p(importances),
                                         int(bool(equal_importance)), int(percentage), int(iteration), C.c_uint32(flags)))


def _volume_minmax(self, vol, region, out):
    od = (C.c_int * 3)()
    self._check(lib().cpm_volume_minmax(self.h, vol.handle, int(region), _p(out), od))
    return tuple(od)


def _volume_diff_bricks(self, a, b, region, scaling, rmin, rmax, out):
    self._check(lib().cpm_volume_diff_bricks(self.h, a.handle, b.handle, int(region), C.c_double(scaling),
                                             C.c_double(rmin), C.c_double(rmax), _p(out)))


def _classify_importance(self, minmax, n, positions, colors, n_points, weights, incremental, out, prev=None, diff=None):
    self._check(lib().cpm_classify_importance(self.h, _p(minmax), _p(prev), _p(diff), int(n), _p(positions), _p(colors),
                                              int(n_points), _fN(weights, 4), int(bool(incremental)), _p(out)))


def _hash_light_samples(self, light_samples, intersections, n_src, ids, n_ids, cell_size, n_blocks, out, out_offset=0):
    self._check(lib().cpm_hash_light_samples(self.h, _p(light_samples), _p(intersections), int(n_src), _p(ids), int(n_ids),
                                             _f3(cell_size), _i3(n_blocks), _p(out), int(out_offset)))


def _build_cell_ranges(self, keys, n, n_cells, start, end):
    self._check(lib().cpm_build_cell_ranges(self.h, _p(keys), C.c_size_t(n), C.c_uint32(n_cells), _p(start), _p(end)))


def _splat_photons(self, light_volume, channels, tex2idx, idx2tex, out_dims, photons, indices, n, per_interaction,
                   n_interactions, radius, scale, multiplier=1.0):
    self._check(lib().cpm_splat_photons(self.h, _p(light_volume), int(channels), _fN(tex2idx, 16), _fN(idx2tex, 16),
                                        _i3(out_dims), _p(photons), _p(indices), int(n), int(per_interaction),
                                        int(n_interactions), C.c_float(radius), C.c_float(scale), C.c_float(multiplier)))


def _splat_photons_update(self, light_volume, channels, tex2idx, idx2tex, out_dims, old_photons, new_photons, indices, n,
                          per_interaction, n_interactions, radius, scale, sync=False):
    """sync=True: cpm_splat_photons_update_sync (old_photons is left holding the new records of the listed ids)"""
    f = lib().cpm_splat_photons_update_sync if sync else lib().cpm_splat_photons_update
    self._check(f(self.h, _p(light_volume), int(channels), _fN(tex2idx, 16), _fN(idx2tex, 16),
        _i3(out_dims), _p(old_photons), _p(new_photons), _p(indices), int(n),
        int(per_interaction), int(n_interactions), C.c_float(radius), C.c_float(scale)))


def _copy_index_photons(self, photons, indices, n, multiplier, per_interaction, n_interactions, out, out_offset=0):
    self._check(lib().cpm_copy_index_photons(self.h, _p(photons), _p(indices), int(n), C.c_float(multiplier), int(per_interaction),
                                             int(n_interactions), _p(out), C.c_size_t(out_offset)))


Context.copy_index_photons = _copy_index_photons
Context.detect_invalid = _detect_invalid
Context.splat_photons_update = _splat_photons_update
Context.volume_minmax = _volume_minmax
Context.volume_diff_bricks = _volume_diff_bricks
Context.classify_importance = _classify_importance
Context.hash_light_samples = _hash_light_samples
Context.build_cell_ranges = _build_cell_ranges
Context.splat_photons = _splat_photons


# -- photon map for gathering --------------------------------------------------------------------------
class GatherParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("cam_origin", C.c_float * 3), ("cam_dir00", C.c_float * 3),
                ("cam_du", C.c_float * 3), ("cam_dv", C.c_float * 3), ("aabb_min", C.c_float * 3),
                ("aabb_max", C.c_float * 3), ("step", C.c_float), ("radius", C.c_float), ("scale", C.c_float),
                ("sigma_scale", C.c_float), ("grid_dims", C.c_int32 * 3), ("opacity_bound", C.c_void_p),
                ("bound_cell_log2", C.c_int32), ("strip_first", C.c_int32), ("strip_stride", C.c_int32),
                ("planar_records", C.c_int32)]


def make_gather_params(width, height, eye, look_at, up=(0, 1, 0), fov_deg=60.0, step=1.0 / 256, radius=1.0 / 64,
                       scale=1.0, sigma_scale=150.0, grid_dims=(32, 32, 32), aabb_min=(0, 0, 0), aabb_max=(1, 1, 1),
                       cls=None, opacity_bound=None, bound_cell_log2=3):
    """pinhole camera in texture space -> the (dir00, du, dv) ray basis of cpm_gather_params"""
    import numpy as np
    eye, look_at, up = (np.asarray(v, np.float64) for v in (eye, look_at, up))
    f = look_at - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, up)
    r /= np.linalg.norm(r)
    u = np.cross(r, f)
    h = np.tan(np.radians(fov_deg) / 2.0)
    w = h * width / height
    du = 2.0 * w * r / width
    dv = -2.0 * h * u / height
    dir00 = f - w * r + h * u
    p = (cls or GatherParams)()
    p.width, p.height = int(width), int(height)
    for name, v in (("cam_origin", eye), ("cam_dir00", dir00), ("cam_du", du), ("cam_dv", dv), ("aabb_min", aabb_min),
                    ("aabb_max", aabb_max)):
        getattr(p, name)[:] = [float(np.float32(x)) for x in v]
    p.step, p.radius, p.scale, p.sigma_scale = float(step), float(radius), float(scale), float(sigma_scale)
    p.grid_dims[:] = [int(g) for g in grid_dims]
    if opacity_bound is not None:
        p.opacity_bound = opacity_bound.data_ptr()
        p.bound_cell_log2 = bound_cell_log2
    return p


def _photon_cell_keys(self, photons, n_records, grid_dims, keys, ids=None):
    self._check(lib().cpm_photon_cell_keys(self.h, _p(photons), C.c_size_t(n_records), _i3(grid_dims), _p(keys), _p(ids)))


def _reorder_photons(self, photons, ids, n, out, planar=False):
    f = lib().cpm_reorder_photons_planar if planar else lib().cpm_reorder_photons
    self._check(f(self.h, _p(photons), _p(ids), C.c_size_t(n), _p(out)))


def _raycast_light_volume(self, vol, tf_rgba, params, light_volume, lv_dims, channels, image):
    self._check(lib().cpm_raycast_light_volume(self.h, vol.handle, _p(tf_rgba), int(tf_rgba.numel() // 4), C.byref(params),
                                               _p(light_volume), _i3(lv_dims), int(channels), _p(image)))


def _gather_raymarch(self, vol, tf_rgba, params, sorted_photons, cell_start, cell_end, image):
    self._check(lib().cpm_gather_raymarch(self.h, vol.handle, _p(tf_rgba), int(tf_rgba.numel() // 4), C.byref(params),
                                          _p(sorted_photons), _p(cell_start), _p(cell_end), _p(image)))


def _gather_points(self, params, sorted_photons, cell_start, cell_end, points, n_points, out):
    self._check(lib().cpm_gather_points(self.h, C.byref(params), _p(sorted_photons), _p(cell_start), _p(cell_end),
                                        _p(points), int(n_points), _p(out)))


def _build_photon_map(self, photons, n_records, grid_dims, torch, planar=False):
    """cell keys -> radix sort (keys, ids) -> cell ranges -> records in cell order (planar: first halves, then second
    halves -- set cpm_gather_params.planar_records = n_records).  Returns (sorted_photons, cell_start, cell_end, keys)."""
    dev = photons.device
    n_cells = int(grid_dims[0]) * int(grid_dims[1]) * int(grid_dims[2])
    keys = torch.empty(n_records, dtype=torch.int32, device=dev)
    ids = torch.empty_like(keys)
    self.photon_cell_keys(photons, n_records, grid_dims, keys, ids)
    tk, tv = torch.empty_like(keys), torch.empty_like(ids)
    bits = max(1, int(n_cells).bit_length())     # keys <= n_cells
    self.radix_sort(keys, ids, tk, tv, max_bits=bits)
    start = torch.zeros(n_cells, dtype=torch.int32, device=dev)
    end = torch.zeros(n_cells, dtype=torch.int32, device=dev)
    self.build_cell_ranges(keys, n_records, n_cells, start, end)
    out = torch.empty_like(photons)
    self.reorder_photons(photons, ids, n_records, out, planar=planar)
    return out, start, end, keys


Context.photon_cell_keys = _photon_cell_keys
Context.reorder_photons = _reorder_photons
Context.gather_raymarch = _gather_raymarch
Context.raycast_light_volume = _raycast_light_volume
Context.gather_points = _gather_points
Context.build_photon_map = _build_photon_map


# -- view importance + importance-driven sample generator ------------------------------------------------
def _view_importance(self, minmax, grid_dims, cell_size, tex2idx, idx2tex, entry, exit_, width, height, tf_min, tf_max, out):
    self._check(lib().cpm_view_importance(self.h, _p(minmax), _i3(grid_dims), _f3(cell_size), _fN(tex2idx, 16),
                                          _fN(idx2tex, 16), _p(entry), _p(exit_), int(width), int(height), C.c_float(tf_min),
                                          C.c_float(tf_max), _p(out)))


def _sample_importance2d(self, importance, width, height, floor_value, uniform_samples, n, scratch, out):
    self._check(lib().cpm_sample_importance2d(self.h, _p(importance), int(width), int(height), C.c_float(floor_value),
                                              _p(uniform_samples), int(n), _p(scratch), _p(out)))


def sample_importance2d_scratch_floats(width, height):
    lib().cpm_sample_importance2d_scratch_floats.restype = C.c_size_t
    return int(lib().cpm_sample_importance2d_scratch_floats(int(width), int(height)))


Context.view_importance = _view_importance
Context.sample_importance2d = _sample_importance2d


def _mix(self, x, y, a, n, fmt, out):
    self._check(lib().cpm_mix(self.h, _p(x), _p(y), C.c_float(a), C.c_size_t(n), int(fmt), _p(out)))


Context.mix = _mix


# -- multi-GPU communicator (cpm_comm_*: NCCL resolved at run time, own peer kernel for the light-volume sum) ---------
CPM_COMM_ID_BYTES = 128


def comm_unique_id() -> bytes:
    """rank 0: the 128-byte id every rank passes to Comm() (hand it over by any transport)"""
    buf = C.create_string_buffer(CPM_COMM_ID_BYTES)
    rc = lib().cpm_comm_unique_id(buf)
    if rc != 0:
        raise CpmError(rc, "cpm_comm_unique_id: NCCL not available")
    return buf.raw


class Comm:
    """One per process / GPU: cpm_comm_init on a Context.  All methods are collective and asynchronous on the context's
    stream; tensors are device tensors."""

    def __init__(self, ctx: Context, unique_id: bytes = None, rank: int = 0, world: int = 1, handle=None):
        self.ctx = ctx
        self.h = C.c_void_p(handle) if handle is not None else C.c_void_p()
        if handle is None:
            ctx._check(lib().cpm_comm_init(ctx.h, C.c_char_p(unique_id), int(rank), int(world), C.byref(self.h)))

    def split(self, other_ctx: Context) -> "Comm":
        """a second communicator over the same ranks bound to another context (stream) of this process"""
        h = C.c_void_p()
        other_ctx._check(lib().cpm_comm_split(self.h, other_ctx.h, C.byref(h)))
        return Comm(other_ctx, handle=h.value)

    @property
    def rank(self):
        return lib().cpm_comm_rank(self.h)

    @property
    def world(self):
        return lib().cpm_comm_world(self.h)

    @property
    def transport(self):
        return lib().cpm_comm_transport(self.h).decode()

    def allreduce_lightvol(self, local, out):
        self.ctx._check(lib().cpm_allreduce_lightvol(self.h, _p(local), _p(out), C.c_size_t(local.numel())))
        return out

    def allreduce_lightvol_begin(self, local, out):
        self.ctx._check(lib().cpm_allreduce_lightvol_begin(self.h, _p(local), _p(out), C.c_size_t(local.numel())))

    def allreduce_lightvol_end(self, out):
        self.ctx._check(lib().cpm_allreduce_lightvol_end(self.h, _p(out), C.c_size_t(out.numel())))

    def allgather_photons(self, local, out):
        self.ctx._check(lib().cpm_allgather_photons(self.h, _p(local), C.c_size_t(local.numel()), _p(out)))
        return out

    def allgather_volume(self, volume, slab_bytes):
        self.ctx._check(lib().cpm_allgather_volume(self.h, _p(volume), C.c_size_t(slab_bytes)))

    def barrier(self):
        self.ctx._check(lib().cpm_comm_barrier(self.h))

    def allgather_u64(self, values):
        """host values of every rank: list of `world` lists (synchronous)"""
        n = len(values)
        mine = (C.c_ulonglong * n)(*[int(v) for v in values])
        out = (C.c_ulonglong * (n * self.world))()
        self.ctx._check(lib().cpm_comm_allgather_u64(self.h, mine, n, out))
        return [list(out[r * n:(r + 1) * n]) for r in range(self.world)]

    def select_global(self, sorted_keys, n_local, position):
        """how many of this rank's ascending keys are among the first `position` of the global (key, rank, index) order"""
        c = C.c_ulonglong()
        self.ctx._check(lib().cpm_comm_select_global(self.h, _p(sorted_keys), C.c_size_t(n_local), C.c_ulonglong(int(position)), C.byref(c)))
        return int(c.value)

    def close(self):
        if self.h:
            lib().cpm_comm_destroy(self.h)
            self.h = C.c_void_p()
