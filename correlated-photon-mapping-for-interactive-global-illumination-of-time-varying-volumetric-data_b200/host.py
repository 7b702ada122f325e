"""ctypes binding of libcpm_host.so: the headless driver of the drop-in Inviwo processor network
(host/host_capi.h).  This is the reference-facing call path: host buffers in, host buffers out."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
HOST_LIB_PATH = Path(os.environ.get("CPM_HOST_LIB", _HERE / "libcpm_host.so"))   # override: tuning sweeps only
_lib = None


class HostConfig(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("dims", C.c_int32 * 3), ("format", C.c_int32), ("samples_per_side", C.c_int32),
        ("n_lights", C.c_int32), ("light_directions", (C.c_float * 3) * 8), ("light_intensity", (C.c_float * 3) * 8),
        ("max_scattering_events", C.c_int32), ("light_volume_option", C.c_int32), ("light_volume_channels", C.c_int32),
        ("with_importance_grid", C.c_int32), ("volume_layout", C.c_int32), ("photon_radius_voxels", C.c_float),
        ("max_incremental_percent", C.c_float), ("clip", C.c_int32 * 6), ("reference_full_splat_bound", C.c_int32),
        ("incremental_threshold_percent", C.c_float), ("opacity_bound_cell_log2", C.c_int32),
    ]


def lib():
    global _lib
    if _lib is None:
        if not HOST_LIB_PATH.exists():
            raise FileNotFoundError(f"{HOST_LIB_PATH} missing: run __graft_entry__.build()")
        _lib = C.CDLL(str(HOST_LIB_PATH))
        _lib.cpmh_last_error.restype = C.c_char_p
        _lib.cpmh_network_last_splat_path.restype = C.c_char_p
        _lib.cpmh_network_last_splat_path.argtypes = [C.c_void_p]
        _lib.cpmh_network_stage_ms.restype = C.c_float
        _lib.cpmh_network_stage_ms.argtypes = [C.c_void_p, C.c_char_p]
        _lib.cpmh_network_launch_count.restype = C.c_uint64
        _lib.cpmh_network_launch_count.argtypes = [C.c_void_p, C.c_int]
        _lib.cpmh_describe_processors.restype = C.c_char_p
        _lib.cpmh_network_ctx.restype = C.c_void_p
        _lib.cpmh_network_ctx.argtypes = [C.c_void_p]
        _lib.cpmh_network_destroy.argtypes = [C.c_void_p]
        _lib.cpmh_network_destroy.restype = None
        _lib.cpmh_runtime_init.argtypes = [C.c_int, C.c_void_p, C.c_uint64]
        _lib.cpmh_runtime_set_photon_shard_offset.argtypes = [C.c_uint64]
        _lib.cpmh_network_stream_timestep_host.argtypes = [C.c_void_p, C.c_void_p]
        _lib.cpmh_network_sync.argtypes = [C.c_void_p]
        _lib.cpmh_network_prefetch_timestep_host.argtypes = [C.c_void_p, C.c_void_p]
        _lib.cpmh_network_light_volume_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        _lib.cpmh_network_photons_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        _lib.cpmh_network_count_collision_tests.argtypes = [C.c_void_p, C.c_int]
        _lib.cpmh_network_read_collision_tests.argtypes = [C.c_void_p, C.c_int]
        _lib.cpmh_network_read_collision_tests.restype = C.c_ulonglong
        _lib.cpmh_network_read_collision_stats.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.c_int]
        _lib.cpmh_u3d_write.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float),
                                        C.POINTER(C.c_float), C.c_void_p]
        _lib.cpmh_u3d_read_info.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                            C.POINTER(C.c_float), C.POINTER(C.c_float)]
        _lib.cpmh_u3d_read_data.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t]
        _lib.cpmh_network_export_sequence_grids.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        _lib.cpmh_profile_enable.argtypes = [C.c_int]
        _lib.cpmh_profile_enable.restype = None
        _lib.cpmh_profile_reset.restype = None
        _lib.cpmh_profile_total_ms.argtypes = [C.c_char_p]
        _lib.cpmh_profile_total_ms.restype = C.c_double
        _lib.cpmh_profile_count.argtypes = [C.c_char_p]
        _lib.cpmh_profile_stages.restype = C.c_char_p
        _lib.cpmh_workspace_describe.restype = C.c_char_p
        _lib.cpmh_workspace_describe.argtypes = [C.c_char_p]
        _lib.cpmh_config_from_workspace.argtypes = [C.c_char_p, C.POINTER(C.c_float), C.POINTER(HostConfig)]
        _lib.cpmh_network_load_workspace.argtypes = [C.c_void_p, C.c_char_p]
        _lib.cpmh_network_get_property.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_char_p, C.POINTER(C.c_double)]
        _lib.cpmh_runtime_ctx.restype = C.c_void_p
        _lib.cpmh_runtime_set_comm.argtypes = [C.c_void_p, C.c_int]
        _lib.cpmh_network_sum_light_volume.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        _lib.cpmh_network_read_light_volume_async.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        _lib.cpmh_network_wait_readback.argtypes = [C.c_void_p]
        _lib.cpmh_network_importance_tf_points.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib.cpmh_network_set_property.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_char_p, C.c_double]
        _lib.cpmh_photondata_progress.argtypes = [C.c_size_t, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.POINTER(C.c_double)]
        _lib.cpmh_network_photon_state.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        _lib.cpmh_network_set_data_range.argtypes = [C.c_void_p, C.c_double, C.c_double]
        _lib.cpmh_network_read_importance_keys.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        _lib.cpmh_network_read_recomputed_indices.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        _lib.cpmh_network_read_importance_grid.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        _lib.cpmh_network_light_setup.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        _lib.cpmh_network_read_light_samples.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    return _lib


def runtime_init(device=0, stream=None, photon_shard_offset=0):
    """one context per process; stream = cudaStream_t handle (int) or None; see cpmh_runtime_init"""
    rc = lib().cpmh_runtime_init(int(device), C.c_void_p(stream or 0), C.c_uint64(photon_shard_offset))
    if rc < 0:
        raise HostError(f"cpm host error {rc}: {lib().cpmh_last_error().decode()}")


def set_photon_shard_offset(offset: int):
    """first global photon id of this process (used by the next RNG seeding); see cpmh_runtime_init"""
    rc = lib().cpmh_runtime_set_photon_shard_offset(C.c_uint64(offset))
    if rc < 0:
        raise HostError(f"cpm host error {rc}: {lib().cpmh_last_error().decode()}")


def runtime_ctx() -> int:
    """the process's cpm_ctx* (create the multi-GPU communicator on it: capi.Comm(handle=...) / cpm_comm_init)"""
    return lib().cpmh_runtime_ctx()


def runtime_set_comm(comm_handle, sharded_ingest=True):
    """hand the process's cpm_comm* to the host layer; sharded_ingest: host volumes are uploaded as this rank's slab and
    completed over NVLink (see cpmh_runtime_set_comm).  None detaches."""
    rc = lib().cpmh_runtime_set_comm(C.c_void_p(comm_handle or 0), int(bool(sharded_ingest)))
    if rc < 0:
        raise HostError(f"cpm host error {rc}: {lib().cpmh_last_error().decode()}")


def runtime_set_global_budget(on=True):
    """re-trace budget over the photons of all shards (cpm_comm_select_global) instead of per shard"""
    rc = lib().cpmh_runtime_set_global_budget(int(bool(on)))
    if rc != 0:
        raise RuntimeError(lib().cpmh_last_error().decode())


def profile_enable(on=True):
    lib().cpmh_profile_enable(int(on))


def profile_only(stage=None):
    """while profiling is on, time only this stage (None: every stage)"""
    lib().cpmh_profile_only(stage.encode() if stage else None)


def profile_reset():
    lib().cpmh_profile_reset()


def profile_total_ms(stage):
    return float(lib().cpmh_profile_total_ms(stage.encode()))


def profile_count(stage):
    return int(lib().cpmh_profile_count(stage.encode()))


def profile_stages():
    s = lib().cpmh_profile_stages().decode()
    return [x for x in s.split(",") if x]


class HostError(RuntimeError):
    pass


def workspace_describe(path) -> dict:
    """{"processors": [(class id, name, [stored property paths])], "connections": [(out, in)]} of an .inv file"""
    t = lib().cpmh_workspace_describe(str(path).encode())
    if t is None:
        raise HostError(f"cpm host error: {lib().cpmh_last_error().decode()}")
    out = {"processors": [], "connections": []}
    for line in t.decode().splitlines():
        f = line.split("|")
        if f[0] == "processor":
            out["processors"].append((f[1], f[2], [p for p in f[3].split(",") if p]))
        elif f[0] == "connection":
            out["connections"].append((f[1], f[2]))
    return out


def workspace_config(path, basis=None) -> HostConfig:
    """HostConfig with the fields an .inv workspace determines (see cpmh_config_from_workspace)"""
    cfg = HostConfig()
    cfg.n_lights, cfg.samples_per_side, cfg.max_scattering_events = 1, 256, 1
    cfg.light_volume_option, cfg.light_volume_channels = 0, 1
    cfg.photon_radius_voxels, cfg.max_incremental_percent = 1.0, 100.0
    cfg.light_directions[0][:] = [0.0, 0.0, 1.0]
    for i in range(8):
        cfg.light_intensity[i][:] = [1.0, 1.0, 1.0]
    b = None if basis is None else (C.c_float * 3)(*[float(x) for x in basis])
    rc = lib().cpmh_config_from_workspace(str(path).encode(), b, C.byref(cfg))
    if rc < 0:
        raise HostError(f"cpm host error {rc}: {lib().cpmh_last_error().decode()}")
    return cfg


def random_numbers(nx, ny=0, seed=0, evaluations=1):
    """RandomNumberGeneratorCL (ny = 0) / RandomNumberGenerator2DCL evaluated `evaluations` times: the last numbers"""
    out = np.empty(nx * max(ny, 1), np.float32)
    rc = lib().cpmh_random_numbers(int(nx), int(ny), int(seed), int(evaluations), out.ctypes.data_as(C.c_void_p))
    if rc < 0:
        raise HostError(f"cpm host error {rc}: {lib().cpmh_last_error().decode()}")
    return out.reshape(ny, nx) if ny else out


def describe_processors() -> dict:
    """{classIdentifier: (set(port ids), set(property ids))} of the drop-in processors"""
    out = {}
    for line in lib().cpmh_describe_processors().decode().strip().split("\n"):
        cid, ports, props = line.split("|")
        out[cid] = (set(filter(None, ports.split(","))), set(filter(None, props.split(","))))
    return out


class Network:
    """The workspace network of the reference (ws:1178-1271), evaluated headless."""

    def __init__(self, dims, fmt, samples_per_side, light_directions, max_scattering_events=1, light_volume_option=2,
                 light_volume_channels=1, with_importance_grid=False, volume_layout=1, photon_radius_voxels=1.0,
                 max_incremental_percent=100.0, clip=None, device=0, light_intensity=None,
                 reference_full_splat_bound=True, incremental_threshold=0.0, opacity_bound_cell_log2=0):
        cfg = HostConfig()
        cfg.device = device
        cfg.dims[:] = [int(d) for d in dims]
        cfg.format = fmt
        cfg.samples_per_side = samples_per_side
        cfg.n_lights = len(light_directions)
        for i, d in enumerate(light_directions):
            cfg.light_directions[i][:] = [float(x) for x in d]
            it = (1.0, 1.0, 1.0) if light_intensity is None else light_intensity[i]
            cfg.light_intensity[i][:] = [float(x) for x in it]
        cfg.max_scattering_events = max_scattering_events
        cfg.light_volume_option = light_volume_option
        cfg.light_volume_channels = light_volume_channels
        cfg.with_importance_grid = int(with_importance_grid)
        cfg.volume_layout = volume_layout
        cfg.photon_radius_voxels = photon_radius_voxels
        cfg.max_incremental_percent = max_incremental_percent
        if clip:
            cfg.clip[:] = [int(c) for c in clip]
        cfg.reference_full_splat_bound = int(reference_full_splat_bound)
        cfg.incremental_threshold_percent = incremental_threshold
        cfg.opacity_bound_cell_log2 = opacity_bound_cell_log2   # 0 default (8^3 cells), < 0 off
        self.h = C.c_void_p()
        self._keep = []
        self._check(lib().cpmh_network_create(C.byref(cfg), C.byref(self.h)))
        self.cfg = cfg

    @classmethod
    def from_workspace(cls, path, dims, fmt, basis=None, volume_layout=1, device=0, samples_per_side=None,
                       opacity_bound_cell_log2=0, reference_full_splat_bound=True):
        """The network an ".inv" workspace describes (host/workspace.h): topology parameters through
        cpmh_config_from_workspace, then every stored property through cpmh_network_load_workspace.  The volume
        itself is the caller's (dims, fmt); `samples_per_side` overrides the workspace's nSamples."""
        cfg = workspace_config(path, basis)
        n = cfg.n_lights
        net = cls(dims, fmt, samples_per_side or cfg.samples_per_side, [tuple(cfg.light_directions[i]) for i in range(n)],
                  max_scattering_events=cfg.max_scattering_events, light_volume_option=cfg.light_volume_option,
                  light_volume_channels=cfg.light_volume_channels, with_importance_grid=bool(cfg.with_importance_grid),
                  volume_layout=volume_layout, photon_radius_voxels=cfg.photon_radius_voxels,
                  max_incremental_percent=cfg.max_incremental_percent, clip=list(cfg.clip), device=device,
                  light_intensity=[tuple(cfg.light_intensity[i]) for i in range(n)],
                  reference_full_splat_bound=reference_full_splat_bound,
                  incremental_threshold=cfg.incremental_threshold_percent, opacity_bound_cell_log2=opacity_bound_cell_log2)
        net.properties_applied = net.load_workspace(path)
        if samples_per_side:
            net.set_samples_per_side(samples_per_side)
        return net

    def load_workspace(self, path) -> int:
        """apply every stored property of the workspace's processors to this network; returns how many"""
        return self._check(lib().cpmh_network_load_workspace(self.h, str(path).encode()))

    def get_property(self, class_id, prop, k=0) -> float:
        v = C.c_double()
        self._check(lib().cpmh_network_get_property(self.h, class_id.encode(), int(k), prop.encode(), C.byref(v)))
        return v.value

    def set_property(self, class_id, prop, value, k=0):
        """scalar / bool / option property of the k-th processor of a class, e.g.
        ("org.inviwo.ProgressivePhotonTracerCL", "equalImportance", True)"""
        self._check(lib().cpmh_network_set_property(self.h, class_id.encode(), int(k), prop.encode(), C.c_double(float(value))))

    def importance_tf_points(self):
        """(positions (n,), colors (n, 4)) of the point list the importance classifier last used"""
        pos, col = np.zeros(256, np.float32), np.zeros((256, 4), np.float32)
        n = self._check(lib().cpmh_network_importance_tf_points(self.h, pos.ctypes.data_as(C.c_void_p),
                                                                col.ctypes.data_as(C.c_void_p), 256))
        return pos[:n].copy(), col[:n].copy()

    def tracer_timer_event(self):
        """one tick of the tracer's progressive-refinement timer (onTimerEvent)"""
        self._check(lib().cpmh_network_timer_event(self.h))

    def photon_state(self) -> dict:
        """PhotonData after the last evaluation: iteration, radius (world), scene radius, relative radius, irradiance scale"""
        out = (C.c_double * 5)()
        self._check(lib().cpmh_network_photon_state(self.h, out))
        return dict(iteration=int(out[0]), radius=out[1], scene_radius=out[2], radius_rel=out[3], irradiance_scale=out[4])

    def set_samples_per_side(self, n):
        self._check(lib().cpmh_network_set_samples_per_side(self.h, int(n)))

    def _check(self, rc):
        if rc < 0:
            raise HostError(f"cpm host error {rc}: {lib().cpmh_last_error().decode()}")
        return rc

    def close(self):
        if self.h:
            lib().cpmh_network_destroy(self.h)
            self.h = C.c_void_p()

    def set_transfer_function(self, points):
        """points: iterable of (pos, (r, g, b, a))"""
        flat = np.array([[p, *c] for p, c in points], np.float32)
        self._check(lib().cpmh_network_set_transfer_function(self.h, flat.ctypes.data_as(C.c_void_p), len(flat)))

    def set_volume_host(self, ptr_or_array):
        if isinstance(ptr_or_array, np.ndarray):
            self._keep = [ptr_or_array]
            ptr = ptr_or_array.ctypes.data
        elif hasattr(ptr_or_array, "data_ptr"):
            self._keep = [ptr_or_array]
            ptr = ptr_or_array.data_ptr()
        else:
            ptr = int(ptr_or_array)
        self._check(lib().cpmh_network_set_volume_host(self.h, C.c_void_p(ptr)))

    def set_sequence_host(self, arrays):
        self._seq = list(arrays)
        ptrs = (C.c_void_p * len(arrays))(*[a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data for a in arrays])
        self._check(lib().cpmh_network_set_sequence_host(self.h, ptrs, len(arrays)))

    def set_data_range(self, lo, hi):
        """Volume::dataMap_.dataRange, e.g. (0, 4095) for 12-bit data in a 16-bit volume"""
        self._check(lib().cpmh_network_set_data_range(self.h, float(lo), float(hi)))

    def set_volume_layout(self, layout):
        """CPM_VOLUME_TEXTURE / CPM_VOLUME_LINEAR for the tracer from now on (see cpmh_network_set_volume_layout)"""
        self._check(lib().cpmh_network_set_volume_layout(self.h, int(layout)))

    def set_timestep(self, t):
        self._check(lib().cpmh_network_set_timestep(self.h, int(t)))

    def stream_timestep_host(self, array):
        """upload the next time step from a (pinned) host array / tensor; see cpmh_network_stream_timestep_host"""
        ptr = array.data_ptr() if hasattr(array, "data_ptr") else array.ctypes.data
        self._stream_keep = array
        self._check(lib().cpmh_network_stream_timestep_host(self.h, C.c_void_p(ptr)))

    def prefetch_timestep_host(self, array):
        """announce the next stream_timestep_host buffer: its upload overlaps the current evaluation"""
        ptr = array.data_ptr() if hasattr(array, "data_ptr") else array.ctypes.data
        self._prefetch_keep = array
        self._check(lib().cpmh_network_prefetch_timestep_host(self.h, C.c_void_p(ptr)))

    def sync(self):
        self._check(lib().cpmh_network_sync(self.h))

    def light_volume_device(self):
        """(device pointer, number of floats) of the light volume"""
        p, n = C.c_void_p(), C.c_size_t()
        self._check(lib().cpmh_network_light_volume_device(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def wait_before_light_volume_write(self, event):
        """the next write to the light volume waits for `event` (torch.cuda.Event recorded on another stream)"""
        self._keep_event = event
        self._check(lib().cpmh_network_wait_before_light_volume_write(self.h, C.c_void_p(event.cuda_event)))

    def photons_device(self):
        """(device pointer, number of floats) of the photon records"""
        p, n = C.c_void_p(), C.c_size_t()
        self._check(lib().cpmh_network_photons_device(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def count_collision_tests(self, on=True):
        self._check(lib().cpmh_network_count_collision_tests(self.h, int(on)))

    def read_collision_tests(self, reset=False):
        return int(lib().cpmh_network_read_collision_tests(self.h, int(reset)))

    def read_collision_stats(self, reset=False):
        """(collision tests, tests that fetched voxels); the rest were decided by the opacity bound"""
        out = (C.c_ulonglong * 2)()
        self._check(lib().cpmh_network_read_collision_stats(self.h, out, int(reset)))
        return int(out[0]), int(out[1])

    def evaluate(self) -> int:
        return self._check(lib().cpmh_network_evaluate(self.h))

    @property
    def remaining_photons(self):
        return lib().cpmh_network_remaining_photons(self.h)

    @property
    def n_photons(self):
        return lib().cpmh_network_n_photons(self.h)

    @property
    def n_recomputed(self):
        return lib().cpmh_network_n_recomputed(self.h)

    @property
    def light_volume_dims(self):
        d = (C.c_int * 3)()
        self._check(lib().cpmh_network_light_volume_dims(self.h, d))
        return tuple(d)

    def read_light_volume(self, out=None):
        d = self.light_volume_dims
        n = d[0] * d[1] * d[2] * self.cfg.light_volume_channels
        if out is None:
            out = np.empty(n, np.float32)
        ptr = out.data_ptr() if hasattr(out, "data_ptr") else out.ctypes.data
        self._check(lib().cpmh_network_read_light_volume(self.h, C.c_void_p(ptr), C.c_size_t(n)))
        return out

    def sum_light_volume(self, out_host=None):
        """sum over ranks of the light volumes (cpm_allreduce_lightvol through the host layer's communicator); returns the
        device pointer of the sum; out_host (pinned tensor / array): also read back into it, synchronously"""
        d = self.light_volume_dims
        n = d[0] * d[1] * d[2] * self.cfg.light_volume_channels
        ptr = C.c_void_p()
        hp = 0 if out_host is None else (out_host.data_ptr() if hasattr(out_host, "data_ptr") else out_host.ctypes.data)
        self._check(lib().cpmh_network_sum_light_volume(self.h, C.c_void_p(hp), C.c_size_t(n), C.byref(ptr)))
        return ptr.value

    def read_light_volume_async(self, out_host=None, sum_over_ranks=True):
        """start the read-back of the frame result (summed over ranks when a communicator is set) into a pinned buffer on
        the read-back stream and return at once; out_host None: only form the sum (ranks that do not display)"""
        d = self.light_volume_dims
        n = d[0] * d[1] * d[2] * self.cfg.light_volume_channels
        hp = 0 if out_host is None else (out_host.data_ptr() if hasattr(out_host, "data_ptr") else out_host.ctypes.data)
        self._keep_readback = out_host
        self._check(lib().cpmh_network_read_light_volume_async(self.h, C.c_void_p(hp), C.c_size_t(n), int(bool(sum_over_ranks))))

    def wait_readback(self):
        """block until the most recent read_light_volume_async has landed in its host buffer"""
        self._check(lib().cpmh_network_wait_readback(self.h))

    def read_photons(self, max_interactions):
        n = self.n_photons * max_interactions * 8
        out = np.empty(n, np.float32)
        self._check(lib().cpmh_network_read_photons(self.h, out.ctypes.data_as(C.c_void_p), C.c_size_t(n)))
        return out.reshape(-1, 8)

    def read_importance_keys(self):
        """the tracer's per-photon importance keys (0x7FFFFFFF = valid), device -> host"""
        out = np.empty(self.n_photons, np.uint32)
        self._check(lib().cpmh_network_read_importance_keys(self.h, out.ctypes.data_as(C.c_void_p), C.c_size_t(out.size)))
        return out

    def read_recomputed_indices(self):
        """ids re-traced by the last evaluation (empty when it traced everything)"""
        n = max(self.n_recomputed, 0)
        out = np.empty(n, np.uint32)
        if n:
            m = self._check(lib().cpmh_network_read_recomputed_indices(self.h, out.ctypes.data_as(C.c_void_p), C.c_size_t(n)))
            out = out[:m]
        return out

    def read_importance_grid(self, n_cells):
        out = np.empty(int(n_cells), np.float32)
        self._check(lib().cpmh_network_read_importance_grid(self.h, out.ctypes.data_as(C.c_void_p), C.c_size_t(out.size)))
        return out

    def light_setup(self, light=0) -> dict:
        """kernel arguments of one directional light sampler: direction, plane point (before the fit), fitted origin, u, v,
        radiance (float32 triples) and area"""
        out = (C.c_float * 19)()
        self._check(lib().cpmh_network_light_setup(self.h, int(light), out))
        a = np.array(out[:], np.float32)
        return dict(dir=a[0:3].copy(), plane_point=a[3:6].copy(), origin=a[6:9].copy(), u=a[9:12].copy(), v=a[12:15].copy(),
                    radiance=a[15:18].copy(), area=np.float32(a[18]))

    def read_light_samples(self, light=0):
        """(light samples (n, 8), intersections (n, 2)) of one light sampler, device -> host"""
        n = self.cfg.samples_per_side ** 2
        ls, it = np.empty((n, 8), np.float32), np.empty((n, 2), np.float32)
        self._check(lib().cpmh_network_read_light_samples(self.h, int(light), ls.ctypes.data_as(C.c_void_p),
                                                          it.ctypes.data_as(C.c_void_p), C.c_size_t(n)))
        return ls, it

    @property
    def last_splat_path(self):
        return lib().cpmh_network_last_splat_path(self.h).decode()

    def set_profile(self, on=True):
        lib().cpmh_network_set_profile(self.h, int(on))

    def stage_ms(self, stage):
        return float(lib().cpmh_network_stage_ms(self.h, stage.encode()))

    def launch_count(self, reset=False):
        return int(lib().cpmh_network_launch_count(self.h, int(reset)))

    @staticmethod
    def transfer_bytes(reset=False):
        a, b = C.c_uint64(), C.c_uint64()
        lib().cpmh_transfer_bytes(C.byref(a), C.byref(b), int(reset))
        return a.value, b.value


# ---- ".u3d" uniform-grid sequences (host only) -------------------------------------------------------------------
U3D_FLOAT32, U3D_VEC2UINT16 = 0, 1


def u3d_write(path, grids, cell=(8, 8, 8), model=None, world=None):
    """grids: float32 array (t, z, y, x) or uint16 array (t, z, y, x, 2) -> path (.u3d header) + path's .raw"""
    import numpy as np
    g = np.ascontiguousarray(grids)
    if g.dtype == np.float32 and g.ndim == 4:
        fmt = U3D_FLOAT32
    elif g.dtype == np.uint16 and g.ndim == 5 and g.shape[4] == 2:
        fmt = U3D_VEC2UINT16
    else:
        raise ValueError("grids must be float32 (t,z,y,x) or uint16 (t,z,y,x,2)")
    dims4 = (C.c_int * 4)(g.shape[3], g.shape[2], g.shape[1], g.shape[0])
    ident = np.eye(4, dtype=np.float32)
    m = np.ascontiguousarray(ident if model is None else np.asarray(model, np.float32)).reshape(-1)
    w = np.ascontiguousarray(ident if world is None else np.asarray(world, np.float32)).reshape(-1)
    rc = lib().cpmh_u3d_write(str(path).encode(), fmt, dims4, (C.c_int * 3)(*[int(c) for c in cell]),
                              m.ctypes.data_as(C.POINTER(C.c_float)), w.ctypes.data_as(C.POINTER(C.c_float)),
                              g.ctypes.data_as(C.c_void_p))
    if rc < 0:
        raise HostError(f"cpm host error {rc}: {lib().cpmh_last_error().decode()}")


def u3d_read(path):
    """-> (grids, cell, model, world); grids shaped like u3d_write's input, matrices column-major (4,4)"""
    import numpy as np
    fmt, dims4, cell = C.c_int(0), (C.c_int * 4)(), (C.c_int * 3)()
    m, w = np.zeros(16, np.float32), np.zeros(16, np.float32)
    rc = lib().cpmh_u3d_read_info(str(path).encode(), C.byref(fmt), dims4, cell, m.ctypes.data_as(C.POINTER(C.c_float)),
                                  w.ctypes.data_as(C.POINTER(C.c_float)))
    if rc < 0:
        raise HostError(f"cpm host error {rc}: {lib().cpmh_last_error().decode()}")
    shape = (dims4[3], dims4[2], dims4[1], dims4[0])
    out = np.empty(shape, np.float32) if fmt.value == U3D_FLOAT32 else np.empty(shape + (2,), np.uint16)
    rc = lib().cpmh_u3d_read_data(str(path).encode(), out.ctypes.data_as(C.c_void_p), C.c_size_t(out.nbytes))
    if rc < 0:
        raise HostError(f"cpm host error {rc}: {lib().cpmh_last_error().decode()}")
    return out, tuple(cell), m.reshape(4, 4), w.reshape(4, 4)


def _export_sequence_grids(self, which, path):
    self._check(lib().cpmh_network_export_sequence_grids(self.h, int(which), str(path).encode()))


Network.export_sequence_grids = _export_sequence_grids
