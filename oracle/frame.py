"""frame.py -- the reference's per-frame algorithm (call stacks C and D) executed by the CPU
oracle: first frame = full trace + full splat; a time-step (or transfer-function) change = importance classify ->
photon re-computation detector -> count -> key/value sort -> cut at the budget -> index sort -> re-trace of the
selected photons -> key reset -> (-old / +new) or full splat.

TEST INFRASTRUCTURE (see cpm_oracle.h): the checker of the drop-in host network (tests/test_configs.py,
tests/test_host_processors.py) and the CPU arm of bench.py.  Follows, statement by statement,
ppm/processor/progressivephotontracercl.cpp:219-605 (process), ppm/processor/photontolightvolumeprocessorcl.cpp:131-344
and isc/processors/minmaxuniformgrid3dimportanceclprocessor.cpp:120-300, with the two documented repairs of SURVEY.md
appendix A (keys sorted on a copy; exact synchronous count)."""
from __future__ import annotations

import math

import numpy as np

from . import orc

KEY_VALID = 0x7FFFFFFF


def texture_to_index(dims):
    """StructuredCoordinateTransformer::getTextureToIndexMatrix, column-major float16: p * dim - 0.5"""
    m = np.zeros(16, np.float32)
    m[0], m[5], m[10], m[15] = dims[0], dims[1], dims[2], 1.0
    m[12] = m[13] = m[14] = -0.5
    return m


def index_to_texture(dims):
    m = np.zeros(16, np.float32)
    d = [np.float32(x) for x in dims]
    m[0], m[5], m[10], m[15] = np.float32(1) / d[0], np.float32(1) / d[1], np.float32(1) / d[2], 1.0
    m[12], m[13], m[14] = np.float32(0.5) / d[0], np.float32(0.5) / d[1], np.float32(0.5) / d[2]
    return m


def rasterise_tf(points, width=1024):
    """(width, 4) float32 RGBA exactly as the host layer's TransferFunction rasterises its points (Inviwo's
    TransferFunction::calcTransferValues is un-vendored): constant outside the points, (1 - t) a + t b in double
    between them, texel i at x = i / (width - 1); positions and colours are the float32 values the property stores"""
    pos = [float(np.float32(p)) for p, _ in points]
    col = [[float(np.float32(c)) for c in cc] for _, cc in points]
    out = np.zeros((width, 4), np.float32)
    for i in range(width):
        x = i / (width - 1) if width > 1 else 0.0
        if not pos:
            continue
        if x <= pos[0]:
            c = col[0]
        elif x >= pos[-1]:
            c = col[-1]
        else:
            k = 1
            while pos[k] < x:
                k += 1
            t = (x - pos[k - 1]) / (pos[k] - pos[k - 1])
            c = [(1.0 - t) * col[k - 1][ch] + t * col[k][ch] for ch in range(4)]
        out[i] = np.array(c, np.float64).astype(np.float32)
    return out


def tf_point_lists(points):
    """updateTransferFunctionData (isc/processors/minmaxuniformgrid3dimportanceclprocessor.cpp:304-362): the TF points
    with explicit end points at 0 and 1"""
    pos, col = [], []
    if not points:
        return np.array([0.0, 1.0], np.float32), np.zeros((2, 4), np.float32)
    if points[0][0] > 0.0:
        pos.append(0.0); col.append(points[0][1])
    for p, c in points:
        pos.append(p); col.append(c)
    if points[-1][0] < 1.0:
        pos.append(1.0); col.append(points[-1][1])
    return np.array(pos, np.float32), np.ascontiguousarray(np.array(col, np.float32))


def tf_difference_lists(cur_points, prev_points, eps=1e-4, associated=False):
    """updateTransferFunctionDifferenceData (isc/processors/minmaxuniformgrid3dimportanceclprocessor.cpp:364-501): the
    |new - old| point list the incremental classifier gets after a transfer-function change -- a merge walk over the break
    points of both functions.  Restated statement by statement (the checker's copy; the host layer has its own).  Points:
    [(pos, (r, g, b, a))], positions double, colours float32."""
    f = np.float32

    def col(c):
        return np.array([f(x) for x in c], np.float32)

    def diff(c1, c2):   # tfPointColorDiff: |p2 * (assoc ? p2.w : 1) - p1 * (assoc ? p1.w : 1)| in fp32
        a1, a2 = (c1[3], c2[3]) if associated else (f(1), f(1))
        return np.abs(c2 * a2 - c1 * a1).astype(np.float32)

    def differs(c):     # glm::any(glm::epsilonNotEqual(c, 0, eps))
        return bool((np.abs(c) >= f(eps)).any())

    def colour_at(a, b, t):   # mix(a, b, t): vec4(dvec4(x) + w * dvec4(y - x)), w in double
        with np.errstate(all="ignore"):      # coincident points divide by zero, as in the reference (inf / NaN propagate)
            w = np.float64(t[0] - a[0]) / np.float64(b[0] - a[0])
            return (a[1].astype(np.float64) + w * (b[1] - a[1]).astype(np.float64)).astype(np.float32)

    cur = [(float(f(p)), col(c)) for p, c in cur_points]
    prev = [(float(f(p)), col(c)) for p, c in prev_points]
    nC, nP = len(cur), len(prev)
    zero = np.zeros(4, np.float32)
    if nC == 0 or nP == 0:
        return np.array([0.0, 0.0 if (nC == 0 and nP == 0) else 1.0], np.float32), np.zeros((2, 4), np.float32)
    first, pfirst = cur[0], prev[0]
    p1 = (first[0] if first[0] < pfirst[0] else pfirst[0], diff(first[1], pfirst[1]))
    p2 = p1
    if first[0] != pfirst[0] and first[1][3] == 0 and pfirst[1][3] == 0:
        if first[0] < pfirst[0]:
            a2 = cur[min(1, nC - 1)]
            p2 = (pfirst[0], diff(pfirst[1], colour_at(first, a2, pfirst)))
        else:
            a2 = prev[min(1, nP - 1)]
            p2 = (first[0], diff(first[1], colour_at(pfirst, a2, first)))
    pos, cols = [0.0], []
    cols.append(p1[1] if (p1[0] > 0 and (first[1][3] > 0 or pfirst[1][3] > 0) and differs(p1[1])) else zero)
    i = j = 0
    while i < nC or j < nP:
        if (differs(p1[1]) or differs(p2[1])) and (p1[1][3] > 0 or p2[1][3] > 0):
            if len(pos) == 1:
                pos.append(p1[0]); cols.append(p1[1])
            pos.append(p2[0]); cols.append(p2[1])
        a1 = cur[min(i, nC - 1)]
        a2 = cur[i + 1] if i + 1 < nC - 1 else (1.0, cur[nC - 1][1])
        b1 = prev[min(j, nP - 1)]
        b2 = prev[j + 1] if j + 1 < nP - 1 else (1.0, prev[nP - 1][1])
        p1 = p2
        if a2[0] < b2[0]:
            p2 = (a2[0], diff(a2[1], colour_at(b1, b2, a2)))
            i += 1
        elif b2[0] < a2[0]:
            p2 = (b2[0], diff(b2[1], colour_at(a1, a2, b2)))
            j += 1
        else:
            p2 = (b2[0] if a2[1][3] < b2[1][3] else a2[0], diff(a2[1], b2[1]))
            i += 1
            j += 1
    if p2[0] < 1.0 and p2[1][3] > 0:
        pos.append(p2[0]); cols.append(p2[1])
    if f(pos[-1]) < 1:
        pos.append(1.0); cols.append(zero)
    return np.array(pos, np.float32), np.ascontiguousarray(np.array(cols, np.float32))


def importance_weights(color=0.0, color_diff=0.0, opacity_diff=0.0, opacity=1.0):
    """the four kernel weights as MinMaxUniformGrid3DImportanceCLProcessor::process normalises them (fp32)"""
    f = np.float32
    wn = f(color) + f(color_diff) + f(opacity_diff) + f(opacity)
    if not wn > 0:
        wn = f(1)
    lab = f(1) / np.sqrt(f(100) * f(100) + f(500) * f(500) + f(400) * f(400))
    return (f(color) * lab / wn, f(color_diff) * lab / wn, f(opacity_diff) / wn, f(opacity) / wn)


def _glm_normalize(v):
    """glm::normalize in fp32: v * (1 / sqrt(x*x + y*y + z*z))"""
    f = np.float32
    v = [f(x) for x in v]
    inv = f(1) / np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    return np.array([v[0] * inv, v[1] * inv, v[2] * inv], np.float32)


def directional_light(n_side, direction, radiance=(1.0, 1.0, 1.0), proxy_vertices=None, proxy_indices=None):
    """Light samples + intersections of one directional light the way the reference's host code sets the sampler up
    (lcl/directionallightsamplercl.cpp:66-73): a light placed at centre - 2 d looking along d (the headless network's
    convention), direction = normalize(tm * (0,0,1,0)) -- the third normalisation of the same vector, fp32 --, plane
    point = tm * (0,0,0,1), the CPU plane fit around the proxy mesh, then the oracle's emission kernels."""
    f = np.float32
    if proxy_vertices is None:
        proxy_vertices = np.array([[x, y, z] for z in (0.0, 1.0) for y in (0.0, 1.0) for x in (0.0, 1.0)], np.float32)
        proxy_indices = np.array([0, 2, 1, 1, 2, 3, 4, 5, 6, 5, 7, 6, 0, 1, 4, 1, 5, 4, 2, 6, 3, 3, 6, 7, 0, 4, 2, 2, 4, 6,
                                  1, 3, 5, 3, 7, 5], np.int32)
    d = _glm_normalize(direction)
    plane_point = np.array([f(0.5) - f(2) * d[k] for k in range(3)], np.float32)
    z = _glm_normalize(_glm_normalize(d))        # DirectionalLight::set normalises again, the sampler once more
    o, u, v = orc.fit_light_plane(proxy_vertices, plane_point, z)
    lu = np.sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2])
    lv = np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    area = f(lu * lv)
    n = n_side * n_side
    samples = orc.sample_uniform2d(float(n_side), float(n_side), n)
    ls = orc.light_sample_directional(samples, radiance, z, o, u, v, float(area))
    isect = orc.light_mesh_intersect(proxy_vertices, proxy_indices, ls)
    return dict(light_samples=ls, isect=isect, dir=z, plane_point=plane_point, origin=o, u=u, v=v, area=area, n=n)


class OracleNetwork:
    """lights: list of dicts with "light_samples" (n, 8) and "isect" (n, 2) -- e.g. scenes.directional_light() or the
    arrays built from the host layer's light set-up; photons of light l occupy ids [sum(n_<l), sum(n_<=l))."""

    def __init__(self, dims, lights, tf_rgba, tf_points, max_interactions=1, region=8, light_volume_divisor=2,
                 radius_voxels=1.0, budget_percent=100.0, incremental_threshold_percent=50.0, weights=None,
                 reference_full_splat_bound=False, spatial_sorting=True, aabb=((0, 0, 0), (1, 1, 1)), sampling_rate=1.0,
                 rng_seed=0, first_stream=0):
        self.dims = tuple(int(d) for d in dims)
        self.lights = lights
        self.n = int(sum(L["light_samples"].shape[0] for L in lights))
        self.I = int(max_interactions)
        self.tf = tf_rgba
        self.tfpos, self.tfcol = tf_point_lists(tf_points)
        self.tf_points_ = list(tf_points)
        self.weights = weights if weights is not None else importance_weights()
        self.region = int(region)
        self.gd = tuple(-(-d // self.region) for d in self.dims)
        self.lvdims = tuple(d // light_volume_divisor for d in self.dims)
        self.budget = float(budget_percent)
        self.threshold = float(incremental_threshold_percent)
        self.full_bound = bool(reference_full_splat_bound)
        self.spatial_sorting = bool(spatial_sorting)
        self.aabb = aabb
        # stepSize = samplingRate * min voxel spacing (progressivephotontracercl.cpp:246-250)
        self.step = float(np.float32(sampling_rate) * min(np.float32(1) / np.float32(d) for d in self.dims))
        base = orc.rng_host_base_offsets(rng_seed, first_stream + self.n)[first_stream:].copy()
        self.rng = orc.rng_seed_streams(base, first_stream=first_stream)
        self.photons = np.zeros((self.n * self.I, 8), np.float32)
        self.prev = None
        self.keys = np.full(self.n, KEY_VALID, np.uint32)
        self.lightvol = np.zeros(self.lvdims[0] * self.lvdims[1] * self.lvdims[2], np.float64)
        self.t2i_vol = texture_to_index(self.dims)
        self.t2i_lv = texture_to_index(self.lvdims)
        self.i2t_lv = index_to_texture(self.lvdims)
        # photon radius: |indexToTexture * (r, r, r, 0)| relative to the scene (progressivephotontracercl.cpp:253-262,
        # ppm/photondata.cpp:53-94); unit model / world matrices => scene radius = sqrt(3) / 2
        i2t = index_to_texture(self.dims)
        r = [i2t[0] * np.float32(radius_voxels), i2t[5] * np.float32(radius_voxels), i2t[10] * np.float32(radius_voxels)]
        rel = float(np.sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]))
        scene = float(np.float32(0.5) * np.sqrt(np.float32(3.0)))
        self.radius_rel = (rel * scene) / scene
        self.radius = float(np.float32(self.radius_rel))
        vol = self.radius_rel ** 3 * (math.pi * 4.0 / 3.0)
        self.scale = float(np.float32((1.0 / math.pi) / (vol * self.n)))
        self.remaining, self.remaining_offset = -1, 0
        self.sorted_ids = None
        self.last_splat_path = "none"
        self.tests = 0

    # -- tracer ----------------------------------------------------------------------------------------------
    def _params(self, n_light, offset):
        return orc.trace_params(n_light_samples=n_light, max_interactions=self.I, step_size=self.step, photon_offset=offset,
                                total_photons=self.n, aabb_min=self.aabb[0], aabb_max=self.aabb[1])

    def _trace(self, vol_np, ids=None):
        tests, offset = 0, 0
        V = orc.volume(vol_np)
        for L in self.lights:
            nl = L["light_samples"].shape[0]
            kw = {} if ids is None else dict(recompute=ids, n_recompute=int(ids.size))
            tests += orc.trace_photons(V, self.tf, self._params(nl, offset), L["light_samples"], L["isect"], self.photons,
                                       self.rng, **kw)   # rng is not advanced: the replay property (:642-645)
            offset += nl
        return tests

    def _splat_full(self):
        self.lightvol[:] = 0
        # the reference bounds the full splat with N work-items, i.e. interaction 0 only
        # (photontolightvolumeprocessorcl.cpp:304,368); the repaired path covers all N * I slots
        n = self.n if self.full_bound else self.n * self.I
        orc.splat(self.lightvol, 1, self.t2i_lv, self.i2t_lv, self.lvdims, self.photons, None, n, self.n, self.I, self.radius,
                  self.scale)
        self.last_splat_path = "full"

    def first_frame(self, vol_np):
        self.tests = self._trace(vol_np)
        self.keys[:] = KEY_VALID
        self.remaining, self.remaining_offset = 0, 0
        self._splat_full()
        self.prev = self.photons.copy()
        return self.n

    # -- importance grid ---------------------------------------------------------------------------------------
    def importance_time_varying(self, mm, prev_mm, diff):
        """classifyTimeVaryingMinMaxUniformGrid3DImportanceKernel (Lab formula)"""
        return orc.classify_importance(mm, self.tfpos, self.tfcol, self.weights, False, prev=prev_mm, diff=diff.reshape(-1))

    def importance_static(self, mm, positions=None, colors=None):
        """classifyMinMaxUniformGrid3DImportanceKernel, built with INCREMENTAL_TF_IMPORTANCE"""
        return orc.classify_importance(mm, self.tfpos if positions is None else positions,
                                       self.tfcol if colors is None else colors, self.weights, True)

    # -- the correlated branch -----------------------------------------------------------------------------------
    def detect(self, importance_grid, equal_importance=False):
        """one detector launch per light (progressivephotontracercl.cpp:306-325); equal importance: every
        (100 / percentage)-th photon, shifted by the detector's iteration counter, gets importance 1"""
        offset = 0
        self.detector_iteration = getattr(self, "detector_iteration", 0) + 1
        for L in self.lights:
            nl = L["light_samples"].shape[0]
            orc.detect_invalid(importance_grid, self.gd, (self.region,) * 3, self.t2i_vol, self.photons, offset,
                               L["light_samples"], L["isect"], nl, self.I, self.n, self.keys,
                               equal_importance=equal_importance, percentage=int(self.budget), iteration=self.detector_iteration)
            offset += nl

    def select(self):
        """count + (keys, ids) stable sort on a copy; returns the number of invalid photons"""
        n_inv = orc.count_below(self.keys, KEY_VALID)
        ids = np.arange(self.n, dtype=np.uint32)
        sk = self.keys.copy()
        orc.radix_sort(sk, ids)
        self.sorted_ids = ids
        self.remaining_offset = 0
        if self.remaining < 0 or n_inv > 0:
            self.remaining = n_inv
        return n_inv

    def retrace_batch(self, vol_np):
        """one evaluation of the tracer after select(): re-trace min(remaining, budget) photons; returns their ids"""
        max_update = int(np.float32(self.budget) / np.float32(100.0) * np.float32(self.n))
        m = max(0, min(self.remaining, max_update))
        ids = np.ascontiguousarray(self.sorted_ids[self.remaining_offset:self.remaining_offset + m])
        if m:
            if self.spatial_sorting:
                orc.radix_sort(ids, None)
            self.tests = self._trace(vol_np, ids)
            self.keys[ids] = KEY_VALID
        self.remaining_offset += m
        self.remaining -= m
        self._splat_update(ids)
        return ids

    def _splat_update(self, ids):
        m = int(ids.size)
        max_rec = int(np.float32(self.n) * (np.float32(self.threshold) / np.float32(100.0)))
        if 0 < m < max_rec:
            orc.splat(self.lightvol, 1, self.t2i_lv, self.i2t_lv, self.lvdims, self.prev, ids, m, self.n, self.I, self.radius,
                      self.scale, -1.0)
            orc.splat(self.lightvol, 1, self.t2i_lv, self.i2t_lv, self.lvdims, self.photons, ids, m, self.n, self.I,
                      self.radius, self.scale, 1.0)
            self.last_splat_path = "incremental"
        elif m >= max_rec:
            self._splat_full()
        else:
            self.last_splat_path = "none"
        if m:
            self.prev[:] = self.photons

    def frame(self, vol_np, importance_grid, equal_importance=False):
        """one change handled in one evaluation (budget permitting): returns the re-traced ids (ascending when
        spatial sorting is on)"""
        self.detect(importance_grid, equal_importance)
        self.select()
        return self.retrace_batch(vol_np)

    def set_transfer_function(self, tf_points, width=1024):
        """a transfer-function change: new raster for the tracer, the |new - old| point list for the importance classifier
        (returned as (positions, colours) for importance_static)"""
        lists = tf_difference_lists(tf_points, self.tf_points_)
        self.tf = rasterise_tf(tf_points, width)
        self.tf_points_ = list(tf_points)
        self.tfpos, self.tfcol = tf_point_lists(tf_points)
        return lists
