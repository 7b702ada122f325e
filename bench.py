#!/usr/bin/env python
"""bench.py -- the contract benchmark of the correlated photon-mapping hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[3], "C4" of SURVEY.md section 8d): a synthetic 512^3 float32 time-varying
volume with 32 time steps, 2048^2 = 4 194 304 photons from one directional light, correlated re-tracing.
A *step* is one time-step change handled the way the reference handles it (call stack D):
importance classify -> photon re-computation detector -> count -> key/value radix sort -> index sort ->
re-trace of the invalidated photons -> -old/+new splat into the light volume (with the default budget the count,
both sorts and the index sort collapse into one stable compaction).  Everything goes through the
reference-facing host layer (libcpm_host.so, the drop-in Inviwo processor mirror) on top of the C ABI.

  value     photons (re)traced per second, whole job, the time series resident in HBM
  e2e       the same metric with the volume of every step uploaded from pinned HOST memory inside the timed
            region (its min-max and difference grids computed on the device) and the light volume read back
  roofline  the dominant kernel (trace) against the HBM peak, SURVEY.md section 8(d) byte accounting
  cpu_baseline / --impl reference
            the CPU oracle (oracle/, a restatement of the reference's OpenCL kernels; the reference itself
            needs Inviwo + an OpenCL device, neither exists here) on all host cores, on a bounded photon sample

  gather    gathered frames/s: photon-map build + view-ray-march gather (north-star 5-7), and the light-volume
            ray caster (the reference network's own final image step)

N > 1 (torchrun, one rank per GPU): weak scaling -- every GPU traces its own 4 Mi photons on disjoint MWC64X
substreams against the replicated volume and the per-GPU light volumes are summed on a side stream by one kernel
over NVLink peer memory (cpm_allreduce_peer_f32; NCCL all-reduce as the fallback and as --exchange nccl).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
PKG = "correlated-photon-mapping-for-interactive-global-illumination-of-time-varying-volumetric-data_b200"
FALLBACK_HBM_GBS = 6650.0       # /opt/skills/guides/B200_PROFILING.md, used when MEASURED_PEAKS.json is absent
LIGHT_DIR = (0.3, -0.5, 0.8)
METRIC = "photons_traced_per_sec"
UNIT = "photons/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dims", type=int, default=512)
    ap.add_argument("--photons-side", type=int, default=2048)
    ap.add_argument("--timesteps", type=int, default=32)
    ap.add_argument("--max-interactions", type=int, default=1)
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--ref-seconds", type=float, default=150.0, help="budget of the whole --impl reference run")
    ap.add_argument("--bound-log2", type=int, default=0,
                    help="tracer opacity-bound cells: 0 = default (8^3 voxels), n = 2^n voxels per axis, -1 = off")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-pageable", action="store_true",
                    help="e2e leg from PAGEABLE host buffers (an unregistered Inviwo VolumeRAM) instead of pinned ones: a data point, "
                         "not the default line")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gather", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="N > 1: do not bind the rank to the CPUs next to its GPU (A/B)")
    ap.add_argument("--exchange", choices=["auto", "multimem", "p2p", "cabi", "nccl"], default="auto",
                    help="N > 1: how the per-rank light volumes are summed (auto: peer kernel with multimem where available; "
                         "cabi: the C-ABI communicator alone -- cpm_comm_split + cpm_allreduce_lightvol_begin/_end)")
    ap.add_argument("--no-sharded-ingest", action="store_true",
                    help="N > 1: every rank uploads the whole time step over PCIe instead of its slab + NVLink all-gather (A/B)")
    ap.add_argument("--exchange-ctas", type=int, default=0, help="grid of the peer exchange kernel (0 = default)")
    ap.add_argument("--exchange-eager-wait", action="store_true",
                    help="the launch stream waits for the snapshot copy right after the frame (A/B; default: only the next splat waits)")
    ap.add_argument("--no-check-exchange", action="store_true",
                    help="N > 1: skip the comparison of the first exchanged sum with NCCL's (untimed, on by default)")
    ap.add_argument("--gather-planar", action="store_true", help="photon map as planar halves instead of 32-byte records (A/B)")
    ap.add_argument("--gather-grid-scale", type=float, default=1.0, help="photon-map cells per 2r along an axis (tuning sweeps)")
    ap.add_argument("--view", type=int, default=1024, help="side of the gathered view image")
    ap.add_argument("--e2e-layout", choices=["auto", "texture", "linear"], default="auto",
                    help="layout the tracer samples in the e2e leg (auto: linear -- a streamed step is sampled where it lands)")
    ap.add_argument("--volume-layout", choices=["texture", "linear"], default="texture",
                    help="layout the tracer samples: 2-D layered CUDA array (tld4) or the caller's linear buffer (A/B)")
    return ap.parse_args()


def workload_name(a):
    return (f"C4: synthetic {a.dims}^3 float32 time-varying volume, {a.timesteps} time steps, "
            f"{a.photons_side}^2 photons per GPU, correlated re-tracing, maxScatteringEvents={a.max_interactions}")


# ---------------------------------------------------------------------------------------------------------
# clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs (B200_PROFILING.md's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def kernel_sources_sha():
    """sha256 over the CUDA sources of the library: ties a committed ncu capture to the code it measured"""
    import hashlib
    h = hashlib.sha256()
    for f in sorted((ROOT / PKG / "csrc").glob("*.cu*")) + [ROOT / "include" / "cpm_detmath.h", ROOT / "Makefile"]:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


def profile_counters():
    """DRAM traffic and issue counters of trace_kernel from the committed ncu capture (profiles/roofline_traffic.json,
    written by tools/update_roofline_traffic.py from a `ncu --set full` run of this benchmark).  Only reported when the
    capture was taken from the kernel sources being run; otherwise null + the reason."""
    tp = ROOT / "profiles" / "roofline_traffic.json"
    try:
        d = json.loads(tp.read_text())
    except Exception:   # noqa: BLE001
        return {"traffic": None, "counters": None, "traffic_source": "no profiles/roofline_traffic.json"}
    sha = kernel_sources_sha()
    if d.get("kernel_sources_sha") != sha:
        return {"traffic": None, "counters": None,
                "traffic_source": f"stale: capture of sources {d.get('kernel_sources_sha')}, running {sha}"}
    return {"traffic": d.get("trace_kernel_dram_bytes_per_launch"), "counters": d.get("counters"),
            "traffic_source": d.get("source")}


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            for k in ("hbm_gbs", "hbm_gbps", "hbm_GBs"):
                if k in d:
                    return float(d[k]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle's correlated frame on a photon sample
class CpuArm:
    """The reference's per-frame algorithm executed by the CPU oracle with OpenMP
    (oracle/frame.py: the same class the parity tests check the GPU network against).  Light set-up, transfer-function
    rasterisation and photon streams are the GPU arm's, so on the full photon grid both arms re-trace the same photons."""

    def __init__(self, a, volumes, n_side):
        from oracle import frame, orc
        self.orc, self.a, self.vols = orc, a, volumes
        synth = importlib.import_module(PKG + ".synth")
        dims = (a.dims,) * 3
        light = frame.directional_light(n_side, LIGHT_DIR)
        self.net = frame.OracleNetwork(dims, [light], frame.rasterise_tf(synth.WS_TF_POINTS), synth.WS_TF_POINTS,
                                       max_interactions=a.max_interactions)
        self.n = self.net.n
        self.region = 8
        self.mm, self.diff = {}, {}

    def prepare(self, t):
        """untimed, as in the GPU arm's resident mode: min-max grid of step t and difference grid (t-1 -> t)"""
        T = len(self.vols)
        for k in (t % T, (t - 1) % T):
            if k not in self.mm:
                self.mm[k] = self.orc.volume_minmax(self.vols[k], self.region)
        if (t - 1) % T not in self.diff:
            self.diff[(t - 1) % T] = self.orc.volume_diff_bricks(self.vols[(t - 1) % T], self.vols[t % T], self.region)

    def first_frame(self, t=0):
        self.net.first_frame(self.vols[t])

    def frame(self, t):
        """one time-step change; returns (photons re-traced, collision tests)"""
        T = len(self.vols)
        imp = self.net.importance_time_varying(self.mm[t % T], self.mm[(t - 1) % T], self.diff[(t - 1) % T])
        ids = self.net.frame(self.vols[t % T], imp)
        return int(ids.size), (self.net.tests if ids.size else 0)


def host_threads():
    """host cores this process may use (the cpuset, not what OMP_NUM_THREADS says: torchrun exports 1)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_run(a, volumes, steps, warmup, budget_s, grow_steps=False, full_grid=False):
    """times `steps` CPU frames after `warmup`, the photon sample sized to fit budget_s (full_grid: always the
    whole light-sample grid, so that the arm is the same job for every --gpus N)"""
    from oracle import orc
    cores = orc.set_num_threads(host_threads())
    side = a.photons_side
    if not full_grid:
        # calibrate: trace rate on a 256^2 sample
        cal = CpuArm(a, volumes, 256)
        cal.first_frame(0)
        cal.prepare(1)
        t0 = time.perf_counter()
        cal.frame(1)
        rate = cal.n / max(time.perf_counter() - t0, 1e-6)          # photons in the map per second of frame time
        frames = steps + warmup + 1
        side = int(np.sqrt(max(rate * budget_s / frames, 64.0 * 64.0)))
        side = int(min(a.photons_side, max(64, side // 32 * 32)))
    arm = CpuArm(a, volumes, side)
    arm.first_frame(0)
    for t in range(1, warmup + 1):
        arm.prepare(t)
        arm.frame(t)
    traced = tests = 0
    elapsed = 0.0
    t, done = warmup + 1, 0
    cap = min(a.timesteps, len(volumes) - warmup - 2) if grow_steps else steps
    while done < steps or (grow_steps and elapsed < 0.6 * budget_s and done < cap):
        arm.prepare(t)
        t0 = time.perf_counter()
        n_inv, nt = arm.frame(t)
        elapsed += time.perf_counter() - t0
        traced += n_inv
        tests += nt
        t += 1
        done += 1
    steps = done
    return {"value": traced / elapsed if elapsed > 0 else 0.0, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{side}x{side} = {side * side} photons of the {a.photons_side}^2 light-sample grid, {steps} correlated "
                      f"frames ({elapsed:.2f} s CPU) on the full {a.dims}^3 volumes; min-max / difference grids "
                      f"precomputed untimed as in the resident GPU run",
            "ms_per_step": elapsed / steps * 1e3, "retrace_fraction": traced / (steps * arm.n),
            "n_recomputed_total": int(traced), "photon_grid": f"{side}x{side}",
            "collision_tests_per_sec": tests / elapsed if elapsed > 0 else 0.0, "frames_per_sec": steps / elapsed}


def probe_opencl():
    """the preferred CPU baseline is the reference's own OpenCL kernels on a PoCL CPU device.  That route
    needs (a) an OpenCL ICD with a CPU device and (b) an Inviwo checkout (commit 989dc16e) for the 13 un-vendored .cl
    headers; this probe records what the box has, and the arm falls back to the oracle port (always the case so far)."""
    import ctypes.util
    import glob
    import shutil
    out = {"clinfo": shutil.which("clinfo") is not None, "icd_files": sorted(glob.glob("/etc/OpenCL/vendors/*.icd")),
           "libOpenCL": ctypes.util.find_library("OpenCL"), "libpocl": ctypes.util.find_library("pocl"),
           "cpu_device": None, "inviwo_cl_headers": False}
    if out["clinfo"]:
        try:
            txt = subprocess.run(["clinfo", "-l"], capture_output=True, text=True, timeout=20).stdout
            out["cpu_device"] = any(k in txt.lower() for k in ("pocl", "cpu", "portable computing"))
        except Exception:   # noqa: BLE001
            out["cpu_device"] = False
    for root in (os.environ.get("INVIWO_HOME"), "/opt/inviwo", "/usr/share/inviwo"):
        if root and os.path.exists(os.path.join(root, "modules", "opencl", "cl", "samplers.cl")):
            out["inviwo_cl_headers"] = True
    usable = bool(out["cpu_device"]) and out["inviwo_cl_headers"]
    out["route"] = "reference OpenCL kernels on PoCL" if usable else "oracle port (no CPU OpenCL device and/or no Inviwo headers)"
    return out


def make_volumes_numpy(a, n_steps, device):
    synth = importlib.import_module(PKG + ".synth")
    T = a.timesteps
    return [synth.volume_field_torch((a.dims,) * 3, 4, t / T, device=device).cpu().numpy() for t in range(min(n_steps, T))]


def run_reference(a):
    """--impl reference: the CPU oracle on this box's host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    vols = make_volumes_numpy(a, a.steps + a.warmup + 2, dev)
    # always the whole light-sample grid and every host core, whatever the launcher's OMP_NUM_THREADS says: the arm is
    # the same job (cores, sample) for every --gpus N
    r = cpu_run(a, vols, a.steps, a.warmup, a.ref_seconds, full_grid=True)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "note": "CPU oracle restatement of the reference's OpenCL kernels "
                       "(reference needs Inviwo + OpenCL: unbuildable here), OpenMP over photons"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "frames_per_sec": r["frames_per_sec"], "retrace_fraction": r["retrace_fraction"],
            "n_recomputed_total": r["n_recomputed_total"], "photon_grid": r["photon_grid"],
            "collision_tests_per_sec": r["collision_tests_per_sec"], "gpu_launches": 0, "opencl_probe": probe_opencl()}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
class DevTensorView:
    """__cuda_array_interface__ over a raw device pointer (the light volume owned by the host layer)"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}


def gather_leg(a, cpm, torch, stream, net, vol_host, dev, sharding, rank=0, world=1, reps=5):
    """gathered frames/s: photon-map build (cell keys, onesweep sort, cell ranges, reorder) + view ray march of
    the network's current photon records; CUDA events on the launch stream, median of `reps`.
    world > 1 (SURVEY 8e option A): the ranks' photon records are all-gathered over NCCL, every rank builds the
    replicated map and marches its own strips of the image, the strips are all-gathered into the whole image;
    per-repetition times are the max over ranks."""
    ctx = cpm.Context(dev.index, stream.cuda_stream)
    D, I, n = a.dims, a.max_interactions, a.photons_side ** 2
    ptr, nf = net.photons_device()
    photons = torch.as_tensor(DevTensorView(ptr, nf), device=dev)
    dvol = vol_host.to(dev, non_blocking=True)
    V = ctx.volume_create(dvol, (D, D, D), cpm.CPM_FMT_F32, layout=cpm.CPM_VOLUME_TEXTURE)
    synth = importlib.import_module(PKG + ".synth")
    tf = torch.from_numpy(synth.rasterise_tf(width=1024)).to(dev)
    radius = float(np.float32(np.sqrt(3.0) / D))                       # the tracer's 1-voxel photon radius
    g = int(min(512, max(1, int(a.gather_grid_scale / (2.0 * radius)))))   # scale 1: cell edge >= 2 r (at most 8 cells per point)
    scale = float((1.0 / np.pi) / (4.0 / 3.0 * np.pi * radius ** 3 * n * world))
    # per-cell opacity bound of (this volume, this TF), as the tracer keeps it: the value-range grid is per-step data
    # like the min-max grid (untimed), the TF classification of its cells belongs to the frame (timed with the build)
    bs = 3 if a.bound_log2 <= 0 else a.bound_log2
    bound = None
    if a.bound_log2 >= 0:
        Vl = ctx.volume_create(dvol, (D, D, D), cpm.CPM_FMT_F32)
        gd = cpm.capi.bound_grid_dims((D, D, D), bs)
        ncell = gd[0] * gd[1] * gd[2]
        vrange = torch.empty(2 * ncell, dtype=torch.float32, device=dev)
        ctx.volume_value_range(Vl, bs, vrange)
        bound = torch.empty(ncell, dtype=torch.float32, device=dev)
    P = cpm.capi.make_gather_params(a.view, a.view, (1.7, 1.4, -1.3), (0.5, 0.5, 0.5), fov_deg=40.0, step=0.5 / D,
                                    radius=radius, scale=scale, sigma_scale=150.0, grid_dims=(g, g, g),
                                    opacity_bound=bound, bound_cell_log2=bs)
    P.planar_records = n * I * world if a.gather_planar else 0
    # ---- (world > 1, first) photon-sharded gather: own map, whole image, rgb all-reduce -- no photon exchange
    sharded = None
    if world > 1:
        img_w = torch.empty(a.view * a.view * 4, dtype=torch.float32, device=dev)
        P.planar_records = 0
        tb, tm, tt = [], [], []
        for it in range(reps + 1):
            torch.distributed.barrier()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            e[0].record(stream)
            sp, start, end, _ = ctx.build_photon_map(photons, n * I, (g, g, g), torch)
            if bound is not None:
                ctx.opacity_bound(vrange, ncell, tf, bound)
            e[1].record(stream)
            ctx.gather_raymarch(V, tf, P, sp, start, end, img_w)
            e[2].record(stream)
            full_s = sharding.allreduce_image(img_w.view(-1, 4))
            e[3].record(stream)
            e[3].synchronize()
            if it:
                t = sharding.max_over_ranks([e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[0].elapsed_time(e[3])], device=dev)
                tb.append(t[0]); tm.append(t[1]); tt.append(t[2])
        sharded = {"frames_per_sec": 1e3 / float(np.median(tt)), "frame_ms": float(np.median(tt)),
                   "photon_map_build_ms": float(np.median(tb)), "raymarch_ms": float(np.median(tm)),
                   "image_allreduce_bytes": a.view * a.view * 16, "photons_in_image": n * I * world,
                   "how": "every rank gathers the whole image against the map of its own photon shard (estimator scale for "
                          "the total photon count) and the rgb channels are summed with an NCCL all-reduce: radiance is "
                          "linear in the photon set, the per-rank map stays L2 resident; times are max over ranks"}
        P.planar_records = n * I * world if a.gather_planar else 0
    first, stride, rows = sharding.image_strips(rank, world, a.view)
    if world > 1:
        P.height, P.strip_first, P.strip_stride = rows, first, stride
    img = torch.empty(rows * a.view * 4, dtype=torch.float32, device=dev)
    allp = None
    xchg_ms, build_ms, march_ms, total_ms = [], [], [], []
    for it in range(reps + 1):
        if world > 1:
            torch.distributed.barrier()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        e[0].record(stream)
        allp = sharding.allgather_photons(photons, allp)
        e[1].record(stream)
        sp, start, end, _ = ctx.build_photon_map(allp, n * I * world, (g, g, g), torch, planar=a.gather_planar)
        if bound is not None:
            ctx.opacity_bound(vrange, ncell, tf, bound)
        e[2].record(stream)
        ctx.gather_raymarch(V, tf, P, sp, start, end, img)
        e[3].record(stream)
        full = sharding.allgather_image(img, a.view, a.view)
        e[4].record(stream)
        e[4].synchronize()
        if it:
            t = [e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]), e[0].elapsed_time(e[4])]
            t = sharding.max_over_ranks(t, device=dev)
            xchg_ms.append(t[0]); build_ms.append(t[1]); march_ms.append(t[2]); total_ms.append(t[3])
    # final image from the light volume (the LightingRaycaster step of the workspace network), same camera
    lptr, lnf = net.light_volume_device()
    lvol = torch.as_tensor(DevTensorView(lptr, lnf), device=dev)
    lvd = net.light_volume_dims
    img2 = torch.empty_like(img)
    cast_ms = []
    for it in range(reps + 1):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record(stream)
        ctx.raycast_light_volume(V, tf, P, lvol, lvd, 1, img2)
        e[1].record(stream)
        e[1].synchronize()
        if it:
            cast_ms.append(e[0].elapsed_time(e[1]))
    full = full.reshape(-1, 4)
    cover = float((full[:, 3] > 0).float().mean().item())
    lit = float((full[:, :3].sum(dim=1) > 0).float().mean().item())
    V.destroy()
    ctx.close()
    b, m, x, tot = (float(np.median(v)) for v in (build_ms, march_ms, xchg_ms, total_ms))
    out = {"frames_per_sec": 1e3 / tot if world > 1 else 1e3 / (b + m), "photon_map_build_ms": b, "raymarch_ms": m,
           "light_volume_raycast_ms": float(np.median(cast_ms)), "view": f"{a.view}x{a.view}",
           "grid": f"{g}^3 cells", "pixels_hitting_volume": cover, "pixels_lit": lit,
           "note": "not part of `value`: build = cell keys + onesweep (keys, ids) + cell ranges + reorder of "
                   f"{n * I * world} photon records (+ TF classification of the opacity-bound cells); march = step 0.5 voxel, "
                   "Epanechnikov gather r = 1 voxel, transparent cells stepped over, 12 samples gathered per photon pass"}
    if world > 1:
        out = {"frames_per_sec": sharded["frames_per_sec"], "photon_sharded": sharded, "replicated_map": out,
               "light_volume_raycast_ms": out["light_volume_raycast_ms"], "view": out["view"],
               "note": "frames_per_sec = the photon-sharded gather (SURVEY 8e: no photon exchange, image all-reduce); "
                       "replicated_map = option A (photon all-gather, image strips per rank): the map of all ranks' photons "
                       "outgrows the 126 MB L2, which the latency-bound gather pays for"}
        out["replicated_map"].update({"photon_allgather_ms": x, "frame_ms": tot, "photons_in_map": n * I * world,
                    "tiling": f"strips of 4 rows dealt round robin over {world} ranks, NCCL all-gather of photon records "
                              f"({32 * n * I * world} B) and of the image strips; times are max over ranks"})
    return out


def bind_to_gpu_numa(torch, local):
    """Multi-GPU runs: keep this rank's threads (and with them its pinned host buffers, first-touch) on the CPUs next
    to its GPU -- /sys/bus/pci/devices/<bdf>/local_cpulist -- so that eight ranks uploading 512 MB per step do not all
    cross the same socket link.  Returns the cpulist applied, or None (no sysfs entry, restricted cpuset, ...)."""
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        text = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        cpus = set()
        for part in text.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return text
    except Exception:   # noqa: BLE001 -- best effort
        return None


def run_b200(a):
    import ctypes as C

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a B200 (there is no CPU fallback; use --impl reference for the CPU arm)"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    cpulist = bind_to_gpu_numa(torch, local) if (world > 1 and not a.no_numa_bind) else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cpm = importlib.import_module(PKG)
    host = importlib.import_module(PKG + ".host")
    synth = importlib.import_module(PKG + ".synth")
    sharding = importlib.import_module(PKG + ".sharding")
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(device=dev)
    T, D, I = a.timesteps, a.dims, a.max_interactions
    n_photons = a.photons_side ** 2
    host.runtime_init(local, stream.cuda_stream, sharding.photon_shard(rank, world, n_photons)[0])
    hcomm = None
    if world > 1:
        # the C-ABI communicator of this process (cpm_comm_init on the host layer's context; torch.distributed only carries
        # the NCCL id): sharded ingest of host volumes and the frame sum of the e2e leg go through it
        hcomm = sharding.bootstrap_comm(cpm, host.runtime_ctx())
        host.runtime_set_comm(hcomm.h.value, not a.no_sharded_ingest)

    # ---- the time series: generated on the device, staged in pinned host memory (the e2e source) ----
    pinned = []
    for t in range(T):
        v = synth.volume_field_torch((D, D, D), 4, t / T, device=dev)
        h = torch.empty(v.shape, dtype=torch.float32, pin_memory=True)
        h.copy_(v)
        pinned.append(h)
        del v
    torch.cuda.synchronize()
    torch.cuda.empty_cache()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def time_loop(net, n_steps, first_t, step_fn, drain=None):
        """barrier + sync, CUDA events on the launch stream around n_steps, max over ranks"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record(stream)
        traced = 0
        for k in range(n_steps):
            traced += step_fn(first_t + k)
        if drain is not None:
            drain()                      # the launch stream waits for work still in flight on side streams
        e1.record(stream)
        e1.synchronize()
        torch.cuda.synchronize()
        wall = time.perf_counter() - w0
        ms = e0.elapsed_time(e1)
        t = sharding.max_over_ranks([ms, wall * 1e3], device=dev)
        tr = sharding.sum_over_ranks([traced], device=dev)
        return t[0], t[1], tr[0]

    with torch.cuda.stream(stream):
        net = host.Network((D, D, D), cpm.CPM_FMT_F32, a.photons_side, [LIGHT_DIR], max_scattering_events=I,
                           light_volume_option=2, with_importance_grid=True,
                           volume_layout=cpm.CPM_VOLUME_TEXTURE if a.volume_layout == "texture" else cpm.CPM_VOLUME_LINEAR,
                           reference_full_splat_bound=False, device=local, opacity_bound_cell_log2=a.bound_log2)
        net.set_transfer_function(synth.WS_TF_POINTS)
        net.count_collision_tests(True)
        lv_view = {}

        def allreduce_light_volume():
            """frame result = sum of the per-rank light volumes, out of place (the local one is updated
            incrementally by the next frame); returns the tensor holding the result"""
            ptr, n = net.light_volume_device()
            if lv_view.get("ptr") != ptr:
                lv_view["ptr"], lv_view["t"] = ptr, torch.as_tensor(DevTensorView(ptr, n), device=dev)
                lv_view["sum"] = torch.empty_like(lv_view["t"]) if world > 1 else None
            return sharding.allreduce_light_volume(lv_view["t"], lv_view["sum"])

        # ---------------- resident leg (value) ----------------
        net.set_sequence_host(pinned)          # uploads all T steps once, min-max + difference grids on the device
        net.set_timestep(0)
        net.evaluate()                         # first frame: full trace + full splat

        exchange, exchange_kind = sharding.LightVolumeExchange(), "NCCL all-reduce"
        if world > 1 and a.exchange == "cabi":
            lvd0 = net.light_volume_dims
            exchange = sharding.CommLightVolumeExchange(cpm, hcomm, lvd0[0] * lvd0[1] * lvd0[2], dev)
            exchange_kind = "cpm_allreduce_lightvol_begin/_end on a cpm_comm_split side-stream communicator"
        elif world > 1 and a.exchange != "nccl":
            # one kernel over NVLink peer memory (csrc/exchange.cu); NCCL stays the fallback where symmetric memory
            # cannot be set up
            try:
                lvd0 = net.light_volume_dims
                exchange = sharding.PeerLightVolumeExchange(cpm, lvd0[0] * lvd0[1] * lvd0[2], dev, max_ctas=a.exchange_ctas,
                                                            use_multicast=(a.exchange != "p2p"))
                exchange_kind = ("cpm_allreduce_peer_f32, NVSwitch multimem.ld_reduce/st" if exchange.multicast
                                 else "cpm_allreduce_peer_f32, peer loads / stores")
            except Exception as ex:   # noqa: BLE001 -- any set-up failure means: use NCCL
                sys.stderr.write(f"[rank {rank}] peer exchange unavailable ({type(ex).__name__}: {ex}); using NCCL\n")
            ok = torch.tensor([1.0 if exchange_kind.startswith("cpm") else 0.0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() == 0.0 and exchange_kind.startswith("cpm"):
                exchange, exchange_kind = sharding.LightVolumeExchange(), "NCCL all-reduce"

        def step_resident(t):
            net.set_timestep(t % T)
            net.evaluate()
            if world > 1:
                # frame result = sum of the per-rank light volumes: snapshot + all-reduce on a side stream, overlapping
                # the next frame's kernels; the last one is waited for inside the timed region (time_loop's `drain`)
                ptr, n = net.light_volume_device()
                if lv_view.get("ptr") != ptr:
                    lv_view["ptr"], lv_view["t"] = ptr, torch.as_tensor(DevTensorView(ptr, n), device=dev)
                    lv_view["sum"] = torch.empty_like(lv_view["t"])
                exchange.submit(lv_view["t"], None if a.exchange_eager_wait else net.wait_before_light_volume_write)
            return max(net.n_recomputed, 0) if net.n_recomputed >= 0 else net.n_photons

        def drain_resident():
            if world > 1:
                exchange.result()

        for k in range(a.warmup):
            step_resident(1 + k)
        drain_resident()
        exchange_check = None
        if world > 1 and not a.no_check_exchange:
            # untimed: the sum the exchange kernel produced for the last warm-up frame against NCCL's all-reduce
            got = exchange.result().clone()
            want = sharding.allreduce_light_volume(lv_view["t"])
            torch.cuda.synchronize()
            err = float((got - want).abs().max().item())
            ref = float(want.abs().max().item())
            worst = sharding.max_over_ranks([err], device=dev)[0]
            exchange_check = {"max_abs_diff": worst, "max_abs_value": ref, "against": "NCCL all-reduce of the same per-rank volumes"}
            assert err <= 1e-5 * ref + 1e-12, "peer exchange differs from the NCCL sum"
        net.read_collision_tests(reset=True)
        # inside the timed region only the dominant stage is timed (CUDA events around the trace launches: the roofline's
        # live kernel time); an event pair around every stage costs ~20 us of the 1.1 ms frame
        host.profile_only("trace")
        host.profile_enable(True)
        host.profile_reset()
        net.launch_count(reset=True)
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        ms, wall_ms, traced = time_loop(net, a.steps, 1 + a.warmup, step_resident, drain_resident)
        clk = clocks.stop() if rank == 0 else None
        launches = net.launch_count()
        tests, fetched = net.read_collision_stats(reset=True)
        trace_ms_timed, trace_n_timed = host.profile_total_ms("trace"), host.profile_count("trace")
        # the per-stage breakdown: the same steps once more, untimed, with an event pair around every stage
        host.profile_only(None)
        host.profile_reset()
        n_extra = min(a.steps, 8)
        for k in range(n_extra):
            step_resident(1 + a.warmup + a.steps + k)
        drain_resident()
        torch.cuda.synchronize()
        stages = {s: (host.profile_total_ms(s) * a.steps / n_extra, host.profile_count(s)) for s in host.profile_stages()}
        stages["trace"] = (trace_ms_timed, trace_n_timed)        # (the timed region's own figure)
        host.profile_enable(False)
        net.read_collision_stats(reset=True)
        tests_t = torch.tensor([float(tests)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tests_t, op=dist.ReduceOp.SUM)
        value = traced / (ms * 1e-3)

        # ---------------- e2e leg: host volume in, light volume out, every step ----------------
        e2e = None
        # ---------------- gathered frames: photon-map build + view ray march (north-star 5-7) ----------------
        gather = None
        if not a.no_gather:
            # the resident leg ended on time step warmup + steps + n_extra: its photons and its volume
            gather = gather_leg(a, cpm, torch, stream, net, pinned[(a.warmup + a.steps + n_extra) % T], dev, sharding, rank, world)
        if not a.no_e2e:
            lvd = net.light_volume_dims
            out_host = torch.empty(lvd[0] * lvd[1] * lvd[2], dtype=torch.float32, pin_memory=True)

            # a streamed step is sampled where it lands (the linear buffer): no linear -> CUDA-array copy per step
            e2e_layout = "linear" if a.e2e_layout == "auto" else a.e2e_layout
            net.set_volume_layout(cpm.CPM_VOLUME_LINEAR if e2e_layout == "linear" else cpm.CPM_VOLUME_TEXTURE)
            out_hosts = [out_host, torch.empty_like(out_host).pin_memory()]
            e2e_k = [0]

            src = pinned
            if a.e2e_pageable:
                src = [torch.empty(h.shape, dtype=h.dtype).copy_(h) for h in pinned]   # plain malloc'ed memory

            def step_e2e(t):
                net.stream_timestep_host(src[t % T])            # adopts the upload announced one step earlier
                net.prefetch_timestep_host(src[(t + 1) % T])    # next step's 512 MB (N > 1: this rank's slab + NVLink all-gather)
                net.evaluate()
                # frame result -> host: the sum over ranks (cpm_allreduce_lightvol through the host layer's communicator) is
                # read back by rank 0 -- the process that shows the image -- on the read-back stream; the copy of frame k
                # travels while frame k + 1 is computed and is waited for before the buffer pair is reused
                net.wait_readback()
                net.read_light_volume_async(out_hosts[e2e_k[0] & 1] if rank == 0 else None)
                e2e_k[0] += 1
                return max(net.n_recomputed, 0) if net.n_recomputed >= 0 else net.n_photons

            d2h_extra = [0]
            first = 1 + a.warmup + a.steps
            for k in range(max(a.warmup, 2)):
                step_e2e(first + k)
            first += max(a.warmup, 2)
            host.Network.transfer_bytes(reset=True)
            d2h_extra[0] = 0
            host.profile_enable(True)
            host.profile_reset()
            ms_e, wall_e, traced_e = time_loop(net, a.steps, first, step_e2e, net.wait_readback)   # the last result lands inside the timed region
            st_e = {s: (host.profile_total_ms(s), host.profile_count(s)) for s in host.profile_stages()}
            host.profile_enable(False)
            h2d, d2h = host.Network.transfer_bytes()
            d2h += d2h_extra[0]
            # the per-step grid builders run once per uploaded volume here: live HBM figures for them
            vox_bytes = D ** 3 * 4
            grid_kernels = {}
            for stage, kname, nbytes in (("minmax", "minmax8_kernel (volumeMinMaxKernel)", vox_bytes),
                                         ("voldiff", "diff8_kernel (DynamicVolumeDifferenceAnalysis)", 2 * vox_bytes),
                                         ("range", "range8_kernel (opacity-bound value range)", vox_bytes)):
                tot, cnt = st_e.get(stage, (0.0, 0))
                if cnt:
                    gbs = nbytes / (tot / cnt * 1e-3) / 1e9
                    grid_kernels[stage] = {"kernel": kname, "avg_launch_ms": tot / cnt, "algorithmic_bytes_per_launch": nbytes,
                                           "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak()[0]}
            if rank != 0:
                stream.synchronize()
            tot = sharding.sum_over_ranks([h2d, d2h], device=dev)
            e2e = {"value": traced_e / (wall_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d // a.steps,
                   "d2h_bytes_per_step": d2h // a.steps, "ms_per_step": wall_e / a.steps,
                   "h2d_bytes_per_step_all_ranks": int(tot[0]) // a.steps, "d2h_bytes_per_step_all_ranks": int(tot[1]) // a.steps,
                   "ingest": ("sharded: rank r uploads slab r of the step (1/N of it) from pinned host memory, the slabs are "
                              "all-gathered over NVLink on the transfer stream (cpm_comm_upload_volume_sharded); rank 0 reads "
                              "the summed light volume back" if (world > 1 and not a.no_sharded_ingest) else
                              "every rank uploads the whole step from pinned host memory" if world > 1 else
                              "whole step from pinned (cudaHostAlloc) host memory") +
                             (" -- THIS RUN: --e2e-pageable, plain pageable source buffers" if a.e2e_pageable else ""),
                   "frames_per_sec": a.steps / (wall_e * 1e-3),
                   "h2d_gbs": (h2d / a.steps) / (wall_e / a.steps * 1e-3) / 1e9,   # per rank: the step is PCIe bound
                   "stages_ms_per_step": {s: v[0] / a.steps for s, v in sorted(st_e.items())},
                   "grid_kernels": grid_kernels,
                   "cpu_affinity": cpulist, "volume_layout": e2e_layout,
                   "path": "libcpm_host.so: cpmh_network_stream_timestep_host(pinned host volume; the next step's "
                           "upload is announced with cpmh_network_prefetch_timestep_host and overlaps this step) -> "
                           "cpmh_network_evaluate -> cpmh_network_read_light_volume_async(pinned host buffer; lands while "
                           "the next step is computed, cpmh_network_wait_readback before the buffer pair is reused and at "
                           "the end of the timed region)"}
        net.close()
        if hcomm is not None:
            if hasattr(exchange, "close"):
                exchange.close()
            host.runtime_set_comm(None, False)
            hcomm.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel ----------------
    peak, peak_src = hbm_peak()
    trace_ms, trace_n = stages.get("trace", (0.0, 0))
    dom = max(stages.items(), key=lambda kv: kv[1][0])[0] if stages else "trace"
    tests_rank0 = float(tests)
    traced_rank0 = traced / world
    fetched_rank0 = float(fetched)
    # HBM-algorithmic bytes, SURVEY.md 8(d): stream part 48 + 32 I + 4 (index) B per traced photon + 8 * sizeof(voxel)
    # per collision test THAT FETCHES VOXELS.  The 4 B opacity-bound look-up every test makes hits a ~1 MB table that
    # lives in L1 / L2: reported separately (l2_lookup_bytes_per_launch), not as HBM traffic.  The 2 x 4 B
    # transfer-function taps come from shared memory.
    hbm_bytes = traced_rank0 * (48 + 32 * I + 4) + fetched_rank0 * 8 * 4
    ref_bytes = traced_rank0 * (48 + 32 * I + 4) + tests_rank0 * 8 * 4     # what the reference's loop touches
    ach = hbm_bytes / (trace_ms * 1e-3) / 1e9 if trace_ms > 0 else None
    roof = {"bound": "issue", "kernel": "trace_kernel (photonTracerKernel -D PHOTON_RECOMPUTATION)",
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None, "peak_source": peak_src,
            "traffic": None, "launches": trace_n, "avg_launch_ms": trace_ms / trace_n if trace_n else None,
            "algorithmic_bytes_per_launch": hbm_bytes / trace_n if trace_n else None,
            "l2_lookup_bytes_per_launch": (tests_rank0 * 4 / trace_n) if (trace_n and a.bound_log2 >= 0) else 0,
            "reference_algorithm_bytes_per_launch": ref_bytes / trace_n if trace_n else None,
            "collision_tests_per_launch": tests_rank0 / trace_n if trace_n else None,
            "dominant_stage_by_time": dom,
            "note": "frac = HBM-algorithmic bytes (84 B per re-traced photon + 32 B per collision test that fetches "
                    "voxels) / launch time / measured HBM peak.  The kernel is instruction-issue bound, not HBM bound: "
                    "the opacity-bound grid decides ~98 % of the reference's per-test fetches from an L1/L2-resident "
                    "table (reference_algorithm_bytes_per_launch is what the reference loop would move), so HBM is "
                    "the wrong roof and `counters` (ncu, same kernel sources) carries the binding ones: issue slots "
                    "busy, active lanes per instruction, resident warps"}
    roof.update(profile_counters())
    cpu = None
    if world == 1 and not a.no_cpu:
        vols = [p.numpy() for p in pinned]
        r = cpu_run(a, vols, 3, 1, a.cpu_seconds, grow_steps=True)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "photons_total": n_photons * world,
                       "light_volume": f"{D // 2}^3 f32", "volume_layout": "2-D layered CUDA array (tld4)" if a.volume_layout == "texture" else "linear buffer (8 x LDG)",
                       "l2": "inputs larger than L2: a different 512 MB volume every step, 128 MB photon records",
                       "parallelism": f"photon shards x{world}, light volumes summed on a side stream (overlaps the next frame): {exchange_kind}" if world > 1 else "1 GPU"},
            "frames_per_sec": a.steps / (ms * 1e-3), "retrace_fraction": traced / (a.steps * n_photons * world),
            "n_recomputed_total": int(traced), "photon_grid": f"{a.photons_side}x{a.photons_side}" + (f" x {world} shards" if world > 1 else ""),
            "exchange_check": exchange_check,
            "collision_tests_per_sec": float(tests_t[0]) / (ms * 1e-3),
            "tests_fetching_voxels": fetched / tests if tests else None,
            "wall_ms_per_step": wall_ms / a.steps,
            "stages_ms_per_step": {s: v[0] / a.steps for s, v in sorted(stages.items())},
            "view_frames_per_sec": (1e3 / (ms / a.steps + gather["light_volume_raycast_ms"])) if gather else None,
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gather": gather, "gpu_launches": int(launches), "clocks": clk}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
