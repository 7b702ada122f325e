"""Multi-GPU checks, one process per GPU (launched by tests/test_multigpu.py through torch.distributed.run, or by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_worker.py
Everything goes through the C ABI communicator (cpm_comm_*); torch.distributed only carries the 128-byte NCCL id and
serves as the independent checker (all_gather of the inputs -> rank-ordered fp32 sum)."""
import importlib
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
PKG = "correlated-photon-mapping-for-interactive-global-illumination-of-time-varying-volumetric-data_b200"


def rank_ordered_sum(t):
    """sum over ranks in rank order, fp32 (what the peer kernel computes)"""
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t)
    s = parts[0].clone()
    for p in parts[1:]:
        s += p
    return s, parts


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cpm = importlib.import_module(PKG)
    host = importlib.import_module(PKG + ".host")
    sharding = importlib.import_module(PKG + ".sharding")
    synth = importlib.import_module(PKG + ".synth")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    log = []

    # ---- 1. the communicator on a plain context --------------------------------------------------------------------
    ctx = cpm.Context(local, stream.cuda_stream)
    comm = sharding.bootstrap_comm(cpm, ctx)
    assert comm.rank == rank and comm.world == world
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    for n in (4 * 1024 * 1024, 262144 + 4 * 37, 1000):      # multiples of 4 (peer kernel), the last one tiny
        x = torch.rand(n, device=dev, generator=g) * (1.0 + rank)
        out = torch.zeros_like(x)
        comm.allreduce_lightvol(x, out)
        ctx.sync()
        want, _ = rank_ordered_sum(x)
        assert torch.equal(out.view(torch.int32), want.view(torch.int32)), ("allreduce (peer, rank order)", n)
        # in place
        y = x.clone()
        comm.allreduce_lightvol(y, y)
        ctx.sync()
        assert torch.equal(y.view(torch.int32), want.view(torch.int32)), ("allreduce in place", n)
    log.append(f"allreduce transport: {comm.transport}")
    assert world == 1 or "peer" in comm.transport, comm.transport
    x = torch.rand(1003, device=dev, generator=g)            # not a multiple of 4: NCCL
    out = torch.zeros_like(x)
    comm.allreduce_lightvol(x, out)
    ctx.sync()
    want, _ = rank_ordered_sum(x)
    assert torch.allclose(out, want, rtol=1e-6, atol=0), "allreduce (nccl)"
    ph = torch.rand(4096 * 8, device=dev, generator=g)
    allp = torch.zeros(world * ph.numel(), device=dev)
    comm.allgather_photons(ph, allp)
    ctx.sync()
    _, parts = rank_ordered_sum(ph)
    assert torch.equal(allp, torch.cat(parts)), "allgather photons"
    # volume slabs: in place all-gather, then the upload variants from pinned host memory
    slab = 1 << 20
    full = torch.randint(0, 255, (world * slab,), dtype=torch.uint8, device="cpu", generator=torch.Generator().manual_seed(7))
    pinned = full.pin_memory()
    vol = torch.zeros(world * slab, dtype=torch.uint8, device=dev)
    vol[rank * slab:(rank + 1) * slab] = full[rank * slab:(rank + 1) * slab].to(dev)
    comm.allgather_volume(vol, slab)
    ctx.sync()
    assert torch.equal(vol.cpu(), full), "allgather volume"
    import ctypes as C
    for on_xfer in (0, 1):
        vol.zero_()
        ev = C.c_void_p()
        rc = cpm.lib().cpm_comm_upload_volume_sharded(comm.h, C.c_void_p(vol.data_ptr()), C.c_void_p(pinned.data_ptr()),
                                                      C.c_size_t(vol.numel()), on_xfer, C.byref(ev))
        assert rc == 0, cpm.lib().cpm_last_error(ctx.h)
        if on_xfer:
            assert cpm.lib().cpm_ctx_wait_event(ctx.h, ev) == 0
        ctx.sync()
        assert torch.equal(vol.cpu(), full), ("sharded upload", on_xfer)
        if on_xfer:
            cpm.lib().cpm_event_destroy(ctx.h, ev)
    comm.barrier()
    ctx.sync()

    # ---- 2. pipelined exchange on the C ABI (side-stream communicator) and the multimem kernel ------------------------
    n = 128 ** 3
    lv = torch.rand(n, device=dev, generator=g)
    want, _ = rank_ordered_sum(lv)
    ex = sharding.CommLightVolumeExchange(cpm, comm, n, dev)
    for _ in range(3):
        ex.submit(lv)
        got = ex.result()
        torch.cuda.current_stream().synchronize()
        assert torch.equal(got.view(torch.int32), want.view(torch.int32)), "CommLightVolumeExchange"
    log.append(f"pipelined C-ABI exchange transport: {ex.transport}")
    ex.close()
    try:
        pex = sharding.PeerLightVolumeExchange(cpm, n, dev, use_multicast=True)
        kinds = [("multimem" if pex.multicast else "peer loads (no multicast support on this node)", pex)]
        if pex.multicast:
            kinds.append(("peer loads", sharding.PeerLightVolumeExchange(cpm, n, dev, use_multicast=False)))
        for kind, e in kinds:
            for _ in range(2):
                e.submit(lv)
                got = e.result()
                torch.cuda.current_stream().synchronize()
                if kind == "multimem" and world > 2:
                    # the switch adds the copies in its own order: fp32 sum of `world` terms, any order
                    assert torch.allclose(got, want, rtol=float(world) * 1.2e-7, atol=0), kind
                else:
                    assert torch.equal(got.view(torch.int32), want.view(torch.int32)), kind
            log.append(f"symmetric-memory exchange: {kind} ok")
            e.close()
    except (RuntimeError, ImportError) as e:       # no symmetric memory on this node: reported, not a failure of the ABI
        log.append(f"torch symmetric memory unavailable: {type(e).__name__}: {e}")
    # ---- 2b. global (cross-shard) selection: this rank's share of the first P elements of the (key, rank, index) order --
    gk = torch.Generator().manual_seed(500 + rank)
    n_loc = 5000 + 137 * rank
    raw = torch.cat([torch.randint(0, 40, (n_loc - 900,), generator=gk), torch.randint(0, 2 ** 31 - 1, (600,), generator=gk),
                     torch.full((300,), 2 ** 31 - 1, dtype=torch.int64)])           # many ties, a tail of "valid" photons
    keys_loc = torch.sort(raw).values.to(torch.int32).to(dev)
    sizes = comm.allgather_u64([n_loc])
    assert [v[0] for v in sizes] == [5000 + 137 * r for r in range(world)], "allgather_u64"
    padded = torch.full((5000 + 137 * world,), -1, dtype=torch.int32, device=dev)
    padded[:n_loc] = keys_loc
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    glob = np.concatenate([np.stack([p.cpu().numpy()[:sizes[r][0]].astype(np.int64), np.full(sizes[r][0], r), np.arange(sizes[r][0])], 1)
                           for r, p in enumerate(parts)])
    order = np.lexsort((glob[:, 2], glob[:, 1], glob[:, 0]))
    ranks_in_order = glob[order, 1]
    total = len(order)
    for P in (0, 1, 17, total // 3, total // 2, total - 301 * world, total - 1, total, total + 5):
        want_c = int((ranks_in_order[:min(P, total)] == rank).sum())
        got_c = comm.select_global(keys_loc, n_loc, P)
        assert got_c == want_c, ("select_global", P, got_c, want_c)
    log.append("global selection == one stable sort over the concatenated shards")
    comm.close()
    ctx.close()

    # ---- 3. the host network with sharded ingest == the same network with full uploads -----------------------------
    host.runtime_init(local, stream.cuda_stream, sharding.photon_shard(rank, world, 128 * 128)[0])
    hcomm = sharding.bootstrap_comm(cpm, host.runtime_ctx())
    D, T = 64, 4
    vols = [torch.from_numpy(synth.volume_f32((D, D, D), 4, t / 32.0)).pin_memory() for t in range(T)]
    results = {}
    for sharded in (False, True):
        host.runtime_set_comm(hcomm.h.value, sharded)
        net = host.Network((D, D, D), cpm.CPM_FMT_F32, 128, [(0.3, -0.5, 0.8)], light_volume_option=2, with_importance_grid=True,
                           reference_full_splat_bound=False, device=local)
        net.set_transfer_function(synth.WS_TF_POINTS)
        host.Network.transfer_bytes(reset=True)
        recs = []
        for t in range(T + 1):
            net.stream_timestep_host(vols[t % T])
            net.prefetch_timestep_host(vols[(t + 1) % T])
            net.evaluate()
            recs.append((net.n_recomputed, net.read_photons(1).copy()))
        out = torch.empty(32 ** 3, dtype=torch.float32).pin_memory()
        net.sum_light_volume(out if rank == 0 else None)
        net.sync()
        h2d, _ = host.Network.transfer_bytes()
        results[sharded] = (recs, out.clone(), net.read_light_volume().copy(), h2d)
        net.close()
    (ra, sa, la, ha), (rb, sb, lb, hb) = results[False], results[True]
    for (na, pa), (nb, pb) in zip(ra, rb):
        assert na == nb and np.array_equal(pa.view(np.uint32), pb.view(np.uint32)), "sharded ingest changes the photons"
    # (the splat adds with fp32 atomics: two runs agree to rounding, not bit for bit)
    assert np.sqrt(((la.astype(np.float64) - lb) ** 2).mean()) <= 1e-5 * np.sqrt((la.astype(np.float64) ** 2).mean())
    assert hb < ha and abs(hb * world - ha) <= 0.05 * ha + 4 * 16 * 1024 * 1024, (ha, hb)
    # the summed light volume = rank-ordered sum of the per-rank volumes
    mine = torch.from_numpy(lb).to(dev)
    want, _ = rank_ordered_sum(mine)
    if rank == 0:
        assert torch.equal(sb.to(dev).view(torch.int32), want.view(torch.int32)), "cpmh_network_sum_light_volume"
    log.append(f"sharded ingest: h2d bytes {ha} -> {hb} per rank over {T + 1} steps, photons identical")
    # ---- 4. global budget: max% of the photons of ALL shards per evaluation, most important first --------------------
    for mode in (False, True):
        host.runtime_set_comm(hcomm.h.value, False)
        host.runtime_set_global_budget(mode)
        net = host.Network((D, D, D), cpm.CPM_FMT_F32, 128, [(0.3, -0.5, 0.8)], light_volume_option=2, with_importance_grid=True,
                           reference_full_splat_bound=False, device=local, max_incremental_percent=10.0)
        net.set_transfer_function(synth.WS_TF_POINTS)
        net.set_volume_host(vols[0].numpy())
        net.evaluate()
        pts = [(p, (c[0], c[1], c[2], min(1.0, c[3] * (1.3 + 0.4 * rank)))) for p, c in synth.WS_TF_POINTS]   # shards differ in how much changes
        net.set_transfer_function(pts)
        n_tot = net.n_photons * world
        budget = int(np.float32(0.10) * np.float32(n_tot))
        done_ids, rounds, left, keys_prev = [], 0, None, None
        VALID = 2 ** 31 - 1
        while True:
            net.evaluate()
            rounds += 1
            n_rec = max(net.n_recomputed, 0)
            ids = net.read_recomputed_indices()[:n_rec]
            keys_after = net.read_importance_keys().copy()        # re-traced photons are valid again, the rest keep their keys
            if mode:
                cnt = torch.tensor([n_rec, int((keys_after < VALID).sum())], device=dev)
                dist.all_reduce(cnt)          # (lockstep: in global mode every rank evaluates the same number of times)
                if left is None:
                    left = int(cnt[0]) + int(cnt[1])
                assert int(cnt[0]) == min(left, budget), ("global budget per evaluation", int(cnt[0]), left, budget)
                left -= int(cnt[0])
                assert left == int(cnt[1])
                if keys_prev is not None:
                    # importance order across shards: nothing selected now is less important than anything left anywhere
                    still = keys_after < VALID
                    ext = torch.tensor([int(keys_prev[ids].max()) if n_rec else -1, -(int(keys_prev[still].min()) if still.any() else VALID)],
                                       device=dev)
                    dist.all_reduce(ext, op=dist.ReduceOp.MAX)
                    assert int(ext[0]) <= -int(ext[1]), ("global importance order", int(ext[0]), -int(ext[1]))
            else:
                assert n_rec <= int(np.float32(0.10) * np.float32(net.n_photons))
            keys_prev = keys_after
            done_ids.append(ids.copy())
            if net.remaining_photons <= 0 or rounds > 60:
                break
        allids = np.concatenate(done_ids)
        assert len(np.unique(allids)) == len(allids), "a photon was re-traced twice"
        log.append(f"{'global' if mode else 'per-shard'} budget: {rounds} evaluations")
        net.close()
    host.runtime_set_global_budget(False)
    host.runtime_set_comm(None, False)
    hcomm.close()
    dist.barrier()
    if rank == 0:
        print("MGPU_OK world=%d | " % world + " | ".join(log), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
