"""Multi-GPU host logic on CPU: world_size-2 gloo processes for the exchange helpers,
host-only checks of the per-shard RNG base offsets, and (GPU) shards == one large photon set."""
import ctypes as C
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG_NAME


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = importlib.import_module(PKG_NAME + ".sharding")
    local = torch.full((1000,), float(rank + 1))
    keep = local.clone()
    out = sh.allreduce_light_volume(local)
    ok = bool(torch.equal(local, keep)) and bool((out == sum(range(1, world + 1))).all()) and out.data_ptr() != local.data_ptr()
    # a second frame re-uses the caller's buffer and must not accumulate the previous global sum
    out2 = sh.allreduce_light_volume(local, out)
    ok = ok and out2.data_ptr() == out.data_ptr() and bool((out2 == sum(range(1, world + 1))).all())
    # the pipelined exchange: results alternate between two buffers and never accumulate
    ex = sh.LightVolumeExchange()
    ex.submit(local)
    r1 = ex.result()
    ex.submit(local * 2)
    r2 = ex.result()
    ok = ok and bool((r1 == sum(range(1, world + 1))).all()) and bool((r2 == 2 * sum(range(1, world + 1))).all())
    ok = ok and r1.data_ptr() != r2.data_ptr() and bool(torch.equal(local, keep))
    # option A: photon records of all ranks in rank order; image strips dealt round robin and reassembled
    rec = torch.arange(48, dtype=torch.float32) + 100 * rank
    allp = sh.allgather_photons(rec)
    ok = ok and bool(torch.equal(allp, torch.cat([torch.arange(48, dtype=torch.float32) + 100 * r for r in range(world)])))
    W, H = 5, 22                                    # 6 strips (the last one partial) over `world` ranks
    first, stride, rows = sh.image_strips(rank, world, H)
    cam_row = torch.tensor([4 * (first + (py >> 2) * stride) + (py & 3) for py in range(rows)], dtype=torch.float32)
    mine = cam_row[:, None, None].expand(rows, W, 4).contiguous()
    img = sh.allgather_image(mine, W, H)
    ok = ok and img.shape == (H, W, 4) and bool(torch.equal(img[:, 0, 0], torch.arange(H, dtype=torch.float32)))
    # photon-sharded gather: rgb summed over ranks, the opacity channel stays the rank's own
    part = torch.tensor([[1.0, 2.0, 3.0, 0.5], [0.0, 0.25, 0.0, 0.75]]) * torch.tensor([rank + 1.0, rank + 1.0, rank + 1.0, 1.0])
    tot = sh.allreduce_image(part)
    k = float(sum(range(1, world + 1)))
    ok = ok and bool(torch.equal(tot, torch.tensor([[k, 2 * k, 3 * k, 0.5], [0.0, 0.25 * k, 0.0, 0.75]]))) and tot.data_ptr() != part.data_ptr()
    # global (cross-shard) selection: this rank's share of the first P elements of the (key, rank, index) order
    g = torch.Generator().manual_seed(40 + rank)
    n_loc = 3000 + 111 * rank
    raw = torch.cat([torch.randint(0, 30, (n_loc - 500,), generator=g), torch.randint(0, 2 ** 32 - 1, (300,), generator=g, dtype=torch.int64),
                     torch.full((200,), 2 ** 31 - 1, dtype=torch.int64)])
    mykeys = torch.sort(raw).values
    sizes = [3000 + 111 * r for r in range(world)]
    padded = torch.full((max(sizes),), -1, dtype=torch.int64)
    padded[:n_loc] = mykeys
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    glob = np.concatenate([np.stack([p.numpy()[:sizes[r]], np.full(sizes[r], r), np.arange(sizes[r])], 1) for r, p in enumerate(parts)])
    order = np.lexsort((glob[:, 2], glob[:, 1], glob[:, 0]))
    ranks_in_order = glob[order, 1]
    total = len(order)
    for P in (0, 1, 29, total // 3, total // 2, total - 250, total - 1, total, total + 7):
        want_c = int((ranks_in_order[:min(P, total)] == rank).sum())
        ok = ok and sh.select_global(mykeys, P) == want_c
    mx = sh.max_over_ranks([float(rank), 5.0 - rank])
    sm = sh.sum_over_ranks([float(rank + 1)])
    first, count = sh.photon_shard(rank, world, 4096)
    q.put((rank, ok, mx, sm, first, count))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_exchange_helpers():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, mx, sm, first, count in res:
        assert ok
        assert mx == [1.0, 5.0] and sm == [3.0]
        assert (first, count) == (rank * 4096, 4096)


def test_image_strips_single_process_assembly():
    sh = importlib.import_module(PKG_NAME + ".sharding")
    for world in (1, 2, 3, 8):
        W, H = 3, 37
        parts = []
        for r in range(world):
            first, stride, rows = sh.image_strips(r, world, H)
            assert rows % sh.STRIP_ROWS == 0
            cam = torch.tensor([4 * (first + (py >> 2) * stride) + (py & 3) for py in range(rows)], dtype=torch.float32)
            parts.append(cam[:, None, None].expand(rows, W, 4))
        img = sh.assemble_image(torch.stack(parts), world, W, H)
        assert torch.equal(img[:, 1, 2], torch.arange(H, dtype=torch.float32))
    with pytest.raises(ValueError):
        sh.image_strips(2, 2, 16)


def test_shard_ranges_cover_the_photon_set():
    sh = importlib.import_module(PKG_NAME + ".sharding")
    for total, world in ((10, 3), (4194304, 8), (7, 8), (1, 1)):
        spans = [sh.strong_shard(r, world, total) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        for (f0, c0), (f1, _) in zip(spans, spans[1:]):
            assert f0 + c0 == f1
    with pytest.raises(ValueError):
        sh.photon_shard(2, 2, 10)


def test_host_base_offsets_of_a_shard_are_a_slice_of_the_whole(cpm):
    """cpm_rng_host_base_offsets_range(first, n) == cpm_rng_host_base_offsets(first + n)[first:]"""
    whole = cpm.capi.rng_host_base_offsets(0, 5000)
    for first, n in ((0, 5000), (1234, 1000), (4999, 1)):
        out = np.zeros((n, 2), np.uint32)
        rc = cpm.lib().cpm_rng_host_base_offsets_range(C.c_uint32(0), C.c_uint64(first), out.ctypes.data_as(C.c_void_p),
                                                       C.c_size_t(n))
        assert rc == 0 and np.array_equal(out, whole[first:first + n])


@pytest.mark.gpu
def test_photon_shards_equal_one_large_photon_set(cpm, orc, synth, torch_cuda):
    """Two ranks with one light each == one GPU with the same light twice (2N photons): the photon records of
    shard r are bit-identical to photons [rN, (r+1)N) of the large set, and the per-shard light volumes
    (each normalised by its own N) sum to twice the large set's (normalised by 2N)."""
    host = importlib.import_module(PKG_NAME + ".host")
    dims, ns, I = (48, 48, 48), 64, 2
    n = ns * ns
    d = (0.3, -0.5, 0.8)
    vol = synth.volume_u8(dims, 8)

    def run(lights, shard_offset):
        host.set_photon_shard_offset(shard_offset)
        net = host.Network(dims, cpm.CPM_FMT_U8, ns, lights, max_scattering_events=I, light_volume_option=2,
                           reference_full_splat_bound=False)
        net.set_transfer_function(synth.WS_TF_POINTS)
        net.set_volume_host(vol)
        net.evaluate()
        ph, lv = net.read_photons(I).copy(), net.read_light_volume().astype(np.float64)
        net.close()
        return ph, lv

    try:
        big_ph, big_lv = run([d, d], 0)
        big = big_ph.reshape(I, 2 * n, 8)
        lv_sum = np.zeros_like(big_lv)
        for r in range(2):
            ph, lv = run([d], r * n)
            assert np.array_equal(ph.reshape(I, n, 8).view(np.uint32), big[:, r * n:(r + 1) * n].view(np.uint32)), r
            lv_sum += lv
    finally:
        host.set_photon_shard_offset(0)
    rel = np.sqrt(((lv_sum - 2.0 * big_lv) ** 2).mean()) / np.sqrt(((2.0 * big_lv) ** 2).mean())
    assert rel < 1e-5, rel


@pytest.mark.gpu
@pytest.mark.parametrize("world,n", [(2, 4096), (3, 1000 * 4), (8, 4 * 12345)])
def test_cuda_peer_allreduce_kernel_single_gpu(cpm, ctx, torch_cuda, world, n):
    """cpm_allreduce_peer_f32 (csrc/exchange.cu), peer-load path, with all `world` buffers on one GPU and the ranks'
    calls issued one after the other: every rank reduces its slice in rank order and stores it to every buffer, so
    afterwards all buffers hold the same sum, bit for bit the rank-ordered fp32 sum.  (The NVSwitch multimem path and
    real peer mappings are exercised by bench.py --gpus N --check-exchange.)"""
    torch = torch_cuda
    g = torch.Generator(device="cuda").manual_seed(5)
    bufs = [torch.rand(n, dtype=torch.float32, device="cuda", generator=g) * (r + 1) for r in range(world)]
    want = bufs[0].clone()
    for r in range(1, world):
        want = want + bufs[r]
    ptrs = (C.c_void_p * world)(*[b.data_ptr() for b in bufs])
    for r in range(world):
        rc = cpm.lib().cpm_allreduce_peer_f32(ctx.h, ptrs, C.c_void_p(0), C.c_size_t(n), r, world, 0)
        assert rc == 0
    ctx.sync()
    for b in bufs:
        assert torch.equal(b, want)
    # argument errors
    assert cpm.lib().cpm_allreduce_peer_f32(ctx.h, ptrs, C.c_void_p(0), C.c_size_t(n + 1), 0, world, 0) < 0
    assert cpm.lib().cpm_allreduce_peer_f32(ctx.h, ptrs, C.c_void_p(0), C.c_size_t(n), world, world, 0) < 0
    assert cpm.lib().cpm_allreduce_peer_f32(ctx.h, ptrs, C.c_void_p(0), C.c_size_t(n), 0, 9, 0) < 0
