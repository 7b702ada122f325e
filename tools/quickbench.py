"""Development timing script (not the contract bench): tracer and sort throughput on one GPU."""
import importlib
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
PKG = "correlated-photon-mapping-for-interactive-global-illumination-of-time-varying-volumetric-data_b200"
cpm = importlib.import_module(PKG)
synth = importlib.import_module(PKG + ".synth")
from oracle import orc  # noqa


def timed(fn, stream, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), ts


def main():
    what = sys.argv[1:] or ["trace", "sort"]
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = cpm.Context(0, stream.cuda_stream)
        if "trace" in what:
            for dims, fmt, nside in (((256, 256, 256), "u8", 1024), ((512, 512, 512), "f32", 2048)):
                t0 = time.time()
                vol = synth.volume_u8(dims, 3) if fmt == "u8" else synth.volume_f32(dims, 4)
                tf = synth.rasterise_tf(width=1024)
                d = synth.normalize((0.3, -0.5, 0.8))
                o, u, v = orc.fit_light_plane(synth.CUBE_VERTICES, np.float32([0.5, 0.5, 0.5]) - 2 * d, d)
                area = float(np.float32(np.linalg.norm(u)) * np.float32(np.linalg.norm(v)))
                n = nside * nside
                print(f"[{dims} {fmt}] host setup {time.time()-t0:.1f}s, opaque voxels frac(alpha>0)=",
                      float((vol.astype(np.float32) / (255 if fmt == 'u8' else 1) > 0.0737).mean()))
                st = torch.from_numpy(cpm.capi.rng_host_base_offsets(0, n).view(np.int32)).cuda()
                ms, _ = timed(lambda: ctx.rng_seed_streams(st, n), stream, reps=3, warm=1)
                print(f"  seed_streams n={n}: {ms:.3f} ms  ({n/ms/1e6:.2f} Gstreams/s)")
                st = torch.from_numpy(cpm.capi.rng_host_base_offsets(0, n).view(np.int32)).cuda()
                ctx.rng_seed_streams(st, n)
                s = torch.empty(n * 4, dtype=torch.float32, device="cuda")
                ls = torch.empty(n * 8, dtype=torch.float32, device="cuda")
                it = torch.empty(n * 2, dtype=torch.float32, device="cuda")
                verts, idx = torch.from_numpy(synth.CUBE_VERTICES).cuda(), torch.from_numpy(synth.CUBE_INDICES).cuda()

                def emit():
                    ctx.sample_uniform2d(nside, nside, n, s)
                    ctx.light_sample_directional(s, (1, 1, 1), d, o, u, v, area, n, ls)
                    ctx.light_mesh_intersect(verts, idx, idx.numel(), ls, n, it)
                ms, _ = timed(emit, stream)
                print(f"  emission (3 kernels): {ms:.3f} ms")
                dvol, dtf = torch.from_numpy(vol).cuda(), torch.from_numpy(tf).cuda()
                for I in (1, 4):
                    for layout, lname in ((cpm.CPM_VOLUME_LINEAR, "linear"), (cpm.CPM_VOLUME_TEXTURE, "texture")):
                        V = ctx.volume_create(dvol, dims, cpm.CPM_FMT_U8 if fmt == "u8" else cpm.CPM_FMT_F32, layout=layout)
                        ph = torch.zeros(n * I * 8, dtype=torch.float32, device="cuda")
                        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
                        p = cpm.make_trace_params(n, max_interactions=I, step_size=1.0 / dims[0])
                        ctx.trace_photons(V, dtf, p, ls, it, ph, st, None, 0, cnt)
                        ctx.sync()
                        tests = int(cnt.item())
                        ms, all_ = timed(lambda: ctx.trace_photons(V, dtf, p, ls, it, ph, st, None, 0, None), stream)
                        stored = int((ph.view(-1, 8)[:, 0] < 1e38).sum().item())
                        print(f"  trace I={I} {lname:8s}: {ms:.3f} ms  photons/s={n/ms*1e3:.3e}  tests={tests} "
                              f"({tests/n:.1f}/photon) tests/s={tests/ms*1e3:.3e} stored={stored}  runs={['%.3f'%x for x in all_]}")
                        V.destroy()
        if "grids" in what or "grids512" in what:
            # per-volume grid builders on the C3 / C4 / C5 volumes: min-max bricks, value range of the bound cells
            cases = (((256, 256, 256), "u8"), ((512, 512, 512), "f32"), ((1024, 1024, 1024), "u8"))
            if "grids512" in what:
                cases = cases[1:2]
            for dims, fmt in cases:
                n = dims[0] * dims[1] * dims[2]
                if fmt == "u8":
                    dvol = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda")
                else:
                    dvol = torch.rand(n, dtype=torch.float32, device="cuda")
                V = ctx.volume_create(dvol, dims, cpm.CPM_FMT_U8 if fmt == "u8" else cpm.CPM_FMT_F32)
                nb = [-(-x // 8) for x in dims]
                mm = torch.empty(nb[0] * nb[1] * nb[2] * 2, dtype=torch.int16, device="cuda")
                ms, _ = timed(lambda: ctx.volume_minmax(V, 8, mm), stream)
                bytes_ = n * (1 if fmt == "u8" else 4)
                print(f"  [{dims[0]}^3 {fmt}] volume_minmax: {ms:.3f} ms = {bytes_/ms/1e6:.0f} GB/s")
                for s_ in (2, 3, 4):
                    gd = cpm.capi.bound_grid_dims(dims, s_)
                    rg = torch.empty(2 * gd[0] * gd[1] * gd[2], dtype=torch.float32, device="cuda")
                    ms, _ = timed(lambda: ctx.volume_value_range(V, s_, rg), stream)
                    print(f"  [{dims[0]}^3 {fmt}] volume_value_range cell {1 << s_}: {ms:.3f} ms = {bytes_/ms/1e6:.0f} GB/s")
                dvol2 = dvol.flip(0).contiguous()
                V2 = ctx.volume_create(dvol2, dims, cpm.CPM_FMT_U8 if fmt == "u8" else cpm.CPM_FMT_F32)
                df = torch.empty(nb[0] * nb[1] * nb[2], dtype=torch.float32, device="cuda")
                ms, _ = timed(lambda: ctx.volume_diff_bricks(V, V2, 8, 1.0, 0.0, 1.0, df), stream)
                print(f"  [{dims[0]}^3 {fmt}] volume_diff_bricks: {ms:.3f} ms = {2*bytes_/ms/1e6:.0f} GB/s")
                V2.destroy()
                V.destroy()
                del dvol, dvol2
        if "sort26" in what:
            what = what + ["sort"]
        if "sort" in what:
            for logn in ((26,) if "sort26" in what else (20, 24, 26, 28)):
                n = 1 << logn
                keys = torch.randint(0, 2**31 - 1, (n,), dtype=torch.int32, device="cuda")
                imp = torch.full((n,), 0x7FFFFFFF, dtype=torch.int32, device="cuda")
                sel = torch.rand(n, device="cuda") < 0.1
                imp[sel] -= (torch.rand(int(sel.sum()), device="cuda") * 3000).int()
                for name, src in (("uniform", keys), ("importance", imp)):
                    k, v = src.clone(), torch.arange(n, dtype=torch.int32, device="cuda")
                    tk, tv = torch.empty_like(k), torch.empty_like(v)

                    def run():
                        k.copy_(src)
                        ctx.radix_sort(k, v, tk, tv)
                    def copy_only():
                        k.copy_(src)
                    ms_c, _ = timed(copy_only, stream)
                    ms, _ = timed(run, stream)
                    ms -= ms_c
                    ok = bool((k[1:] >= k[:-1]).all().item())
                    print(f"  sort kv 2^{logn} {name:10s}: {ms:.3f} ms  {n/ms/1e6:.2f} Gpairs/s  {68*n/ms/1e6:.0f} GB/s(68B/pair) sorted={ok}")
                    def run_k():
                        k.copy_(src)
                        ctx.radix_sort(k, None, tk, None)
                    ms, _ = timed(run_k, stream)
                    ms -= ms_c
                    print(f"  sort k  2^{logn} {name:10s}: {ms:.3f} ms  {n/ms/1e6:.2f} Gkeys/s  {36*n/ms/1e6:.0f} GB/s(36B/key)")
                del keys, imp, k, v, tk, tv
        ctx.close()


if __name__ == "__main__":
    main()
