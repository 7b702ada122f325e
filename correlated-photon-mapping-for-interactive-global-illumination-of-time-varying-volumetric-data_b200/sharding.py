"""Multi-GPU plumbing of the photon path: one process per GPU, photons sharded by
index range, one exchange step.  torch.distributed is the transport (NCCL on GPUs, gloo in the CPU tests);
no compute lives here.

Photon i of rank r is photon r * photons_per_gpu + i of the global photon set: it uses that MWC64X stream and
that host base offset (cpmh_runtime_init / cpm_rng_host_base_offsets_range), so the union of all ranks'
photon records is bit-identical to one GPU tracing the whole set.  Every rank splats its own photons into its
own light volume; the frame's result is the SUM over ranks, produced out of place -- the per-rank volume must
stay local because the next frame updates it incrementally (-old/+new of the rank's re-traced photons).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def photon_shard(rank: int, world: int, photons_per_gpu: int):
    """(first global photon id, count) owned by `rank` under weak scaling"""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return rank * photons_per_gpu, photons_per_gpu


def strong_shard(rank: int, world: int, total: int):
    """(first, count) of a fixed photon set split as evenly as possible (ranks < total % world get one more)"""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def allreduce_light_volume(local: torch.Tensor, out: torch.Tensor | None = None, group=None) -> torch.Tensor:
    """out = sum over ranks of `local`; `local` is left untouched.  Single process: returns `local`."""
    if not is_distributed():
        return local
    if out is None or out.shape != local.shape or out.dtype != local.dtype or out.device != local.device:
        out = torch.empty_like(local)
    out.copy_(local)
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


class LightVolumeExchange:
    """The same sum, pipelined: submit() snapshots the rank's light volume and starts the all-reduce on a side
    stream, so that the NEXT frame's detector / re-trace / splat run while NVLink moves this frame's volume;
    result() makes the caller's stream wait for the most recent sum.  Two result buffers alternate, so a result
    stays valid while the following frame is being exchanged.  On CPU tensors (gloo tests) and in a single process
    it degrades to the synchronous allreduce_light_volume."""

    def __init__(self, group=None):
        self.group = group
        self.bufs = [None, None]
        self.turn = 0
        self.pending = None     # (work, buffer) of the most recent submit
        self.side = None
        self.copied = None

    def submit(self, local: torch.Tensor, defer_wait=None) -> None:
        """defer_wait(event): instead of making the launch stream wait for the snapshot copy right away, hand the event
        to the only writer of `local` (host.Network.wait_before_light_volume_write): the next frame's detector and
        re-trace then start at once and only its splat waits -- by then the copy is long done."""
        if not is_distributed():
            self.pending = (None, local)
            return
        i = self.turn
        self.turn ^= 1
        b = self.bufs[i]
        if b is None or b.shape != local.shape or b.dtype != local.dtype or b.device != local.device:
            b = self.bufs[i] = torch.empty_like(local)
        if local.device.type != "cuda":
            b.copy_(local)
            dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group)
            self.pending = (None, b)
            return
        cur = torch.cuda.current_stream(local.device)
        if self.side is None:
            self.side = torch.cuda.Stream(device=local.device)
            self.copied = torch.cuda.Event()
        if self.pending is not None and self.pending[0] is not None:
            with torch.cuda.stream(self.side):
                self.pending[0].wait()          # collectives stay ordered on the side stream
        self.side.wait_stream(cur)              # this frame's splat has to be complete
        with torch.cuda.stream(self.side):
            b.copy_(local, non_blocking=True)
            self.copied.record(self.side)
            work = dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        if defer_wait is not None:
            defer_wait(self.copied)
        else:
            cur.wait_event(self.copied)         # the next frame may touch `local` once the snapshot is taken
        b.record_stream(self.side)
        self.pending = (work, b)

    def result(self) -> torch.Tensor:
        """the most recently submitted sum; the current stream is made to wait for it"""
        if self.pending is None:
            raise RuntimeError("LightVolumeExchange.result() before submit()")
        work, b = self.pending
        if work is not None:
            work.wait()                          # current stream waits for the collective
            self.pending = (None, b)
        return b


def bootstrap_comm(cpm, ctx_or_handle, group=None):
    """The C-ABI communicator (cpm_comm_init) of this process: rank 0 makes the NCCL id, torch.distributed carries the 128
    bytes to the other ranks (a C++ host would use MPI or a socket for this one message), every rank initialises.
    ctx_or_handle: a cpm.Context, or the raw cpm_ctx* of the host layer (host.runtime_ctx())."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [cpm.capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    if isinstance(ctx_or_handle, int):
        class _Raw:                      # a Context-like view of a cpm_ctx* owned elsewhere
            def __init__(self, h):
                import ctypes as C
                self.h = C.c_void_p(h)

            def _check(self, rc):
                if rc != 0:
                    raise cpm.capi.CpmError(rc, cpm.lib().cpm_last_error(self.h).decode())
        ctx_or_handle = _Raw(ctx_or_handle)
    return cpm.capi.Comm(ctx_or_handle, box[0], rank, world)


class CommLightVolumeExchange:
    """LightVolumeExchange on the C ABI alone (what a C++ host does): a side-stream context with its own communicator
    (cpm_comm_split); submit() = wait for the frame's splat, cpm_allreduce_lightvol_begin (snapshot), event,
    cpm_allreduce_lightvol_end (the library's peer kernel over CUDA IPC symmetric memory, or NCCL) on the side stream."""

    def __init__(self, cpm, comm, n_floats: int, device):
        self.side = torch.cuda.Stream(device=device)
        self.ctx = cpm.Context(device.index, self.side.cuda_stream)
        self.comm = comm.split(self.ctx)
        self.bufs = [torch.empty(n_floats, dtype=torch.float32, device=device) for _ in range(2)]
        self.copied = torch.cuda.Event()
        self.done = [torch.cuda.Event(), torch.cuda.Event()]
        self.turn, self.pending = 0, None

    @property
    def transport(self):
        return self.comm.transport

    def submit(self, local: torch.Tensor, defer_wait=None) -> None:
        i = self.turn
        self.turn ^= 1
        cur = torch.cuda.current_stream(local.device)
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            self.comm.allreduce_lightvol_begin(local.reshape(-1), self.bufs[i])
            self.copied.record(self.side)
            self.comm.allreduce_lightvol_end(self.bufs[i])
            self.done[i].record(self.side)
        if defer_wait is not None:
            defer_wait(self.copied)
        else:
            cur.wait_event(self.copied)
        self.pending = i

    def result(self) -> torch.Tensor:
        if self.pending is None:
            raise RuntimeError("CommLightVolumeExchange.result() before submit()")
        i = self.pending
        torch.cuda.current_stream(self.bufs[i].device).wait_event(self.done[i])
        return self.bufs[i]

    def close(self):
        torch.cuda.synchronize()
        self.comm.close()
        self.ctx.close()


class PeerLightVolumeExchange:
    """LightVolumeExchange with the sum done by cpm_allreduce_peer_f32 (csrc/exchange.cu) instead of NCCL: the snapshot
    buffers are symmetric memory (torch.distributed._symmetric_memory: one allocation per rank, peer and NVSwitch
    multicast mappings exchanged once), and one small kernel per rank reduces its slice over NVLink -- in the switch
    (multimem.ld_reduce / multimem.st) where the node has multicast support, with peer loads in rank order otherwise.
    On the side stream: snapshot copy -> barrier (all snapshots written) -> kernel -> barrier (all slices stored);
    the barriers are the symmetric-memory signal-pad barriers, the kernel itself never waits on another GPU.
    Same interface and result semantics as LightVolumeExchange; raises at construction if symmetric memory cannot be
    set up on this node (the caller falls back to NCCL)."""

    def __init__(self, cpm, n_floats: int, device, group=None, max_ctas: int = 0, use_multicast: bool = True):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        if not is_distributed():
            raise RuntimeError("PeerLightVolumeExchange needs an initialised process group with world size > 1")
        if n_floats % 4:
            raise ValueError("the light volume must hold a multiple of 4 floats")
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 8:
            raise RuntimeError("at most 8 peers (one NVLink domain)")
        self.n = n_floats
        self.max_ctas = max_ctas
        self.side = torch.cuda.Stream(device=device)
        self.copied = torch.cuda.Event()
        self.done = [torch.cuda.Event(), torch.cuda.Event()]
        self.ctx = cpm.Context(device.index, self.side.cuda_stream)      # launches on the side stream
        self.bufs, self.hdls, self.peer_arrays, self.mc = [], [], [], []
        for _ in range(2):
            b = symm_mem.empty(n_floats, dtype=torch.float32, device=device)
            h = symm_mem.rendezvous(b, self.group.group_name)
            ptrs = (C.c_void_p * self.world)(*[int(p) for p in h.buffer_ptrs])
            mc = int(h.multicast_ptr) if (use_multicast and h.has_multicast_support and h.multicast_ptr) else 0
            self.bufs.append(b); self.hdls.append(h); self.peer_arrays.append(ptrs); self.mc.append(mc)
        self.multicast = all(m != 0 for m in self.mc)
        self.turn = 0
        self.pending = None
        self._lib = cpm.lib()
        self._C = C

    def submit(self, local: torch.Tensor, defer_wait=None) -> None:
        i = self.turn
        self.turn ^= 1
        b, h = self.bufs[i], self.hdls[i]
        cur = torch.cuda.current_stream(local.device)
        self.side.wait_stream(cur)              # this frame's splat has to be complete
        with torch.cuda.stream(self.side):
            b.copy_(local.reshape(-1), non_blocking=True)
            self.copied.record(self.side)
            h.barrier(channel=i)                # every rank's snapshot is in place
            rc = self._lib.cpm_allreduce_peer_f32(self.ctx.h, self.peer_arrays[i], self._C.c_void_p(self.mc[i] if self.multicast else 0),
                                                  self._C.c_size_t(self.n), self.rank, self.world, int(self.max_ctas))
            if rc != 0:
                raise RuntimeError(f"cpm_allreduce_peer_f32 failed: {rc}")
            h.barrier(channel=i)                # every rank's slice is stored everywhere
            self.done[i].record(self.side)
        if defer_wait is not None:
            defer_wait(self.copied)             # only the next write to `local` waits (see LightVolumeExchange.submit)
        else:
            cur.wait_event(self.copied)         # the next frame may touch `local` once the snapshot is taken
        self.pending = i

    def result(self) -> torch.Tensor:
        """the most recently submitted sum; the current stream is made to wait for it"""
        if self.pending is None:
            raise RuntimeError("PeerLightVolumeExchange.result() before submit()")
        i = self.pending
        torch.cuda.current_stream(self.bufs[i].device).wait_event(self.done[i])
        return self.bufs[i]

    def close(self):
        torch.cuda.synchronize()
        self.ctx.close()


# ---- option A of SURVEY 8e: replicated photon map, image tiles per GPU --------------------------------------------
def allgather_photons(local: torch.Tensor, out: torch.Tensor | None = None, group=None) -> torch.Tensor:
    """Every rank's photon records (float32, 8 per record, any number of interactions) concatenated in rank order:
    the replicated photon set the map for gathering is built from (cell keys treat records independently, so the
    interaction-major order inside a rank's slice does not matter).  All ranks must hold the same number of
    records (weak scaling).  Single process: returns `local`."""
    if not is_distributed():
        return local
    world = dist.get_world_size(group)
    n = local.numel()
    if out is None or out.numel() != n * world or out.dtype != local.dtype or out.device != local.device:
        out = torch.empty(n * world, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.reshape(-1), group=group)
    return out


STRIP_ROWS = 4      # tile height of the view ray marchers (cpm_gather_params::strip_first / strip_stride)


def image_strips(rank: int, world: int, height: int):
    """(strip_first, strip_stride, local_rows) of `rank`: strips of 4 image rows are dealt round robin, which
    balances rays that miss the volume against rays that cross it.  Every rank renders the same number of strips
    (ceil(strips / world)); strips past the image are camera rows outside it and are cropped by assemble_image."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    strips = (height + STRIP_ROWS - 1) // STRIP_ROWS
    per_rank = (strips + world - 1) // world
    return rank, world, per_rank * STRIP_ROWS


def assemble_image(parts: torch.Tensor, world: int, width: int, height: int) -> torch.Tensor:
    """parts: [world, local_rows, width, 4] as image_strips dealt them -> [height, width, 4]"""
    local_rows = parts.shape[1]
    s = local_rows // STRIP_ROWS
    full = parts.reshape(world, s, STRIP_ROWS, width, 4).permute(1, 0, 2, 3, 4).reshape(s * world * STRIP_ROWS, width, 4)
    return full[:height]


def allgather_image(local: torch.Tensor, width: int, height: int, group=None) -> torch.Tensor:
    """local: this rank's [local_rows, width, 4] strips -> the whole [height, width, 4] image on every rank"""
    if not is_distributed():
        return assemble_image(local.reshape(1, -1, width, 4), 1, width, height)
    world = dist.get_world_size(group)
    flat = local.reshape(-1).contiguous()
    parts = torch.empty(world * flat.numel(), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(parts, flat, group=group)
    return assemble_image(parts.reshape(world, -1, width, 4), world, width, height)


# ---- photon-sharded gathering: the estimator is linear in the photon set -------------------------------------------------
def allreduce_image(local: torch.Tensor, group=None) -> torch.Tensor:
    """Every rank gathers the WHOLE image against the map of its own photon shard (cpm_gather_params.scale set for the
    total photon count); radiance is a sum over photons -- compositing weights, opacity and early ray termination depend
    on the volume and the transfer function only -- so the image of the union is the sum of the ranks' images in the
    rgb channels, and the opacity channel is every rank's own.  local: [..., 4]; returns a new tensor."""
    if not is_distributed():
        return local
    out = local.clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    out.view(-1, 4)[:, 3] = local.reshape(-1, 4)[:, 3]
    return out


def max_over_ranks(values, device="cpu"):
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if is_distributed():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def sum_over_ranks(values, device="cpu"):
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if is_distributed():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t]


# ---- global (cross-shard) selection: the algorithm of cpm_comm_select_global over torch.distributed -----------------------
def select_global(sorted_keys: torch.Tensor, position: int, group=None) -> int:
    """How many of THIS rank's ascending uint32 keys (an int64 / int32 tensor holding 0 .. 2^32-1) are among the first
    `position` elements of the global (key, rank, local index) order -- one stable sort over the concatenated shards.
    The same four rounds of a 256-ary search over the key bits as csrc/comm.cu (`cpm_comm_select_global`), each one
    all-gather of 257 counts; no keys travel.  Used by the gloo tests and as the checker of the C entry point."""
    keys = sorted_keys.to(torch.int64) & 0xFFFFFFFF
    world = dist.get_world_size(group) if is_distributed() else 1
    rank = dist.get_rank(group) if is_distributed() else 0
    prefix, less, leq = 0, [0] * world, [0] * world
    for shift in (24, 16, 8, 0):
        probes = prefix + (torch.arange(257, dtype=torch.int64) << shift)
        mine = torch.searchsorted(keys.cpu().contiguous(), probes, right=False).to(torch.int64)
        mine[probes > 0xFFFFFFFF] = keys.numel()
        if world > 1:
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine, group=group)
        else:
            parts = [mine]
        total = torch.stack(parts).sum(0)
        pick = int((total[:256] <= position).nonzero().max())     # total[0] <= position by induction
        prefix += pick << shift
        less = [int(p[pick]) for p in parts]
        leq = [int(p[pick + 1]) for p in parts]
    left = max(position - sum(less), 0)
    count = 0
    for r in range(world):
        take = min(leq[r] - less[r], left)
        if r == rank:
            count = less[r] + take
        left -= take
    return count
