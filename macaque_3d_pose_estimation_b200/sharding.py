"""Frame sharding of the hot path over the GPUs of one box.

Every joint-instance is independent, so ranks never exchange data while computing.  The recording
is cut into TILES of ``tile_frames`` consecutive frames that are dealt round-robin (tile t belongs
to rank t % world): the per-point cost of the subset RANSAC is heavy-tailed and clusters in time
(occlusions), so interleaving balances the ranks, and one ROUND (one tile of every rank) is a
contiguous block of ``world`` tiles of the recording.

The only exchange is the delivery of the 3D results (p3d + err = 32 B per joint-instance) to one
rank, into frame-ordered arrays of the whole recording (round j lands at rows [j * world * tile, ...)
— no reorder pass).  Two implementations:

* ``PeerResults`` (NCCL runs, the default on the B200 box): the destination rank's arrays are a
  CUDA IPC window every other rank maps over NVLink peer access; a rank's COPY ENGINE writes each
  finished tile straight to its rows while the next tile's kernels run.  No SM takes part, no rank
  waits for another per tile; one tiny all-reduce at the end of the step orders the destination
  behind everybody's last copy.
* ``RoundGather`` (any ``torch.distributed`` backend; gloo in the CPU tests): one asynchronous
  ``dist.gather`` per round, received in place.

A rank only ever holds the observations of its own tiles.
"""
import ctypes

import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None

RANSAC_TILE_FRAMES = 256


class TilePlan:
    """Round-robin deal of tiles of ``tile_frames`` frames (``per`` joint-instances per frame)."""

    def __init__(self, n_frames, world, tile_frames, per):
        self.n_frames = int(n_frames)
        self.world = int(world)
        self.per = int(per)
        # a tile never needs to be larger than an even split: keeps every rank busy on short clips
        even = -(-self.n_frames // self.world) if self.n_frames else 1
        self.tile = max(1, min(int(tile_frames), even))
        self.n_tiles = -(-self.n_frames // self.tile) if self.n_frames else 0
        self.rounds = -(-self.n_tiles // self.world) if self.n_tiles else 0

    def tile_span(self, t):
        """Frame range [lo, hi) of tile t (empty when t is past the end)."""
        lo = min(t * self.tile, self.n_frames)
        return lo, min(lo + self.tile, self.n_frames)

    def frames_of(self, rank):
        """Frames of ``rank`` in the order it processes them."""
        spans = [self.tile_span(j * self.world + rank) for j in range(self.rounds)]
        parts = [np.arange(a, b, dtype=np.int64) for a, b in spans if b > a]
        return np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)

    def local_offset(self, rank, j):
        """Row offset (joint-instances) of round j inside rank's local arrays."""
        done = 0
        for i in range(j):
            a, b = self.tile_span(i * self.world + rank)
            done += b - a
        return done * self.per


def frame_range(n_frames, rank, world_size):
    """Contiguous, balanced frame range [lo, hi) of ``rank`` (first n_frames % world ranks get
    one extra frame) — the layout of the DLT path, which has uniform cost per point."""
    base, rem = divmod(int(n_frames), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def tile_frames_of(n_frames, rank, world_size, tile):
    """Frames of ``rank`` under the round-robin deal of tiles of ``tile`` frames."""
    return TilePlan(n_frames, world_size, tile, 1).frames_of(rank)


def shard_points(points, n_frames, rank, world_size, tile=None):
    """Slice a (C, F*per, 2) observation array (frame-major point order) to this rank's frames:
    one contiguous range (``tile`` None) or round-robin tiles of ``tile`` frames.  Convenience for
    callers that hold the whole recording; the sharded run itself only needs the local slice."""
    n = points.shape[1]
    assert n_frames == 0 or n % n_frames == 0, "point count is not a multiple of the frame count"
    per = n // n_frames if n_frames else 0
    if tile is None:
        lo, hi = frame_range(n_frames, rank, world_size)
        return points[:, lo * per:hi * per]
    fr = TilePlan(n_frames, world_size, tile, per).frames_of(rank)
    idx = (fr[:, None] * per + np.arange(per)[None, :]).reshape(-1)
    if torch is not None and isinstance(points, torch.Tensor):
        return points[:, torch.as_tensor(idx, device=points.device)]
    return points[:, idx]


class RoundGather:
    """The per-round gather of result rows to ``dst`` in global frame order.

    ``add(j, tensors)`` is called by every rank after it has queued the kernels of round j;
    ``tensors`` are this rank's result arrays of that round (rows = joint-instances of its tile,
    possibly fewer than a full tile, possibly none).  ``finish()`` waits for all gathers and
    returns the global arrays on ``dst`` (None elsewhere)."""

    def __init__(self, plan, row_shapes, dtype, device, dst=0, group=None):
        self.plan = plan
        self.group = group
        self.dst = dst
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.tile_rows = plan.tile * plan.per
        self.works = []
        self.fixups = []
        self.out = None
        self.device = device
        self.dtype = dtype
        self.row_shapes = [tuple(s) for s in row_shapes]
        if self.rank == dst:
            total = plan.n_frames * plan.per
            self.out = [torch.empty((total,) + s, dtype=dtype, device=device) for s in self.row_shapes]
        # staging for a short / missing tile: every rank contributes exactly tile_rows rows
        self._pad = None

    def _padded(self, t, shape):
        if t is not None and t.shape[0] == self.tile_rows:
            return t.contiguous()
        buf = torch.zeros((self.tile_rows,) + shape, dtype=self.dtype, device=self.device)
        if t is not None and t.shape[0] > 0:
            buf[:t.shape[0]].copy_(t)
        return buf

    def add(self, j, tensors):
        plan = self.plan
        for k, shape in enumerate(self.row_shapes):
            send = self._padded(tensors[k] if tensors is not None else None, shape)
            if self.rank == self.dst:
                bufs = []
                for r in range(self.world):
                    a, b = plan.tile_span(j * self.world + r)
                    rows = (b - a) * plan.per
                    if rows == self.tile_rows:      # full tile: receive in place, already in frame order
                        bufs.append(self.out[k][a * plan.per:b * plan.per])
                    else:                           # ragged end of the recording
                        tmp = torch.empty((self.tile_rows,) + shape, dtype=self.dtype, device=self.device)
                        bufs.append(tmp)
                        if rows > 0:
                            self.fixups.append((k, a * plan.per, rows, tmp))
                self.works.append(dist.gather(send, bufs, dst=self.dst, group=self.group, async_op=True))
            else:
                self.works.append(dist.gather(send, None, dst=self.dst, group=self.group, async_op=True))

    def finish(self):
        for w in self.works:
            w.wait()
        self.works = []
        for k, off, rows, tmp in self.fixups:
            self.out[k][off:off + rows].copy_(tmp[:rows])
        self.fixups = []
        return self.out


class _DeviceBytes:
    """A raw device allocation as seen by ``torch.as_tensor`` (the CUDA array interface)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}


class PeerResults:
    """Frame-ordered result arrays of the whole recording on rank ``dst``, written by every rank's
    copy engine through a peer-mapped window (``m3d_peer_*``, csrc/m3d_peer.cu).

    ``push(j, tensors)`` is called by every rank after it has queued the kernels of round j; the
    copies run on a side stream behind those kernels.  ``row_ptrs(j)`` gives the window addresses of this
    rank's tile of round j (the kernels of ``dst`` write them directly instead of pushing).  ``finish()`` orders the calling stream behind the pushes of ALL ranks and returns the
    global arrays on ``dst`` (None elsewhere).  NCCL only: the window is device memory."""

    def __init__(self, plan, row_shapes, dtype, device, dst=0, group=None):
        from . import _lib
        self.lib = _lib.require_gpu()
        self._check = _lib.check
        self.plan, self.group, self.dst = plan, group, dst
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device(device)
        self.dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.dtype = dtype
        self.row_shapes = [tuple(s) for s in row_shapes]
        item = torch.empty((), dtype=dtype).element_size()
        total = plan.n_frames * plan.per
        self.row_bytes = [item * int(np.prod(s, dtype=np.int64)) for s in self.row_shapes]
        self.offsets, off = [], 0
        for rb in self.row_bytes:                      # 256-byte aligned sections of ONE allocation
            self.offsets.append(off)
            off += -(-(total * rb) // 256) * 256
        self.nbytes = max(off, 256)
        self.base = ctypes.c_void_p()
        self.owner = self.rank == dst
        handle = [None]
        if self.owner:
            buf = (ctypes.c_uint8 * 64)()
            self._check(self.lib.m3d_peer_alloc(self.dev_index, self.nbytes, ctypes.byref(self.base), buf),
                        "m3d_peer_alloc")
            handle[0] = bytes(buf)
        dist.broadcast_object_list(handle, src=dist.get_global_rank(group, dst) if group is not None else dst,
                                   group=group)
        if not self.owner:
            buf = (ctypes.c_uint8 * 64).from_buffer_copy(handle[0])
            self._check(self.lib.m3d_peer_open(self.dev_index, buf, ctypes.byref(self.base)), "m3d_peer_open")
        # the window as tensors: on the owner only (elsewhere the mapping is used through raw pointers, so
        # that torch never has to decide which device an IPC mapping belongs to)
        self.window = None
        if self.owner:
            raw = torch.as_tensor(_DeviceBytes(self.base.value, self.nbytes), device=self.device)
            self.window = [raw[o:o + total * rb].view(dtype).view((total,) + s)
                           for o, rb, s in zip(self.offsets, self.row_bytes, self.row_shapes)]
        self.out = self.window
        self.side = torch.cuda.Stream(self.device)
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._keep = []

    def row_ptrs(self, j, rank=None):
        """Device pointers (valid on THIS rank) of the window rows of the tile of ``rank`` (default: this
        rank) in round j, one per result array: a kernel may write its results there directly."""
        r = self.rank if rank is None else rank
        a, _ = self.plan.tile_span(j * self.world + r)
        return [self.base.value + o + a * self.plan.per * rb for o, rb in zip(self.offsets, self.row_bytes)]

    def push(self, j, tensors):
        a, b = self.plan.tile_span(j * self.world + self.rank)
        rows = (b - a) * self.plan.per
        if rows == 0 or tensors is None:
            return
        cur = torch.cuda.current_stream(self.device)
        self.side.wait_stream(cur)                     # behind the kernels that produce the tile
        for k, rb in enumerate(self.row_bytes):
            t = tensors[k]
            assert t.shape[0] == rows and t.is_contiguous() and t.dtype == self.dtype
            self._keep.append(t)                       # alive until finish()
            self._check(self.lib.m3d_peer_push(self.base.value + self.offsets[k] + a * self.plan.per * rb,
                                               t.data_ptr(), rows * rb, self.side.cuda_stream), "m3d_peer_push")

    def finish(self):
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.side)
        # every rank contributes after its own last copy: on return the calling stream of `dst` is
        # ordered behind the copies (and direct window writes) of all ranks
        dist.all_reduce(self._flag, group=self.group)
        self._keep = []
        return self.out

    def close(self):
        if self.base is not None and self.base.value:
            torch.cuda.synchronize(self.device)
            self.window = self.out = None
            if self.owner:
                dist.barrier(group=self.group)         # nobody writes any more
                self.lib.m3d_peer_free(self.dev_index, self.base)
            else:
                self.lib.m3d_peer_close(self.dev_index, self.base)
                dist.barrier(group=self.group)
            self.base = ctypes.c_void_p()


def triangulate_sharded(cgroup, local_points, n_frames, ransac=False, min_cams=2, gather=True, group=None,
                        tile_frames="auto", exchange="auto"):
    """Run this rank's frames through the GPU CameraGroup and gather (p3d, err) to rank 0.

    ``local_points`` (C, n_local, 2): the observations of THIS rank's frames only, in the order of
    ``TilePlan(...).frames_of(rank)`` (``shard_points`` cuts them out of a full recording);
    ``n_frames``: frames of the whole recording.  ``tile_frames``: an int = tile size of the
    round-robin deal; "auto" = about 16 tiles per rank (at least RANSAC_TILE_FRAMES frames each) for the
    subset RANSAC and one even split per rank for the DLT path.  Returns (p3d, err) of the whole recording in frame order on rank 0, (None,
    None) on the other ranks; with ``gather=False`` every rank gets its local results.
    ``exchange``: "peer" = ``PeerResults`` (copy-engine writes into rank 0's window over NVLink; NCCL
    runs only), "gather" = ``RoundGather`` (one ``dist.gather`` per round), "auto" = "peer" under NCCL."""
    world = dist.get_world_size(group) if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    per_guess = 0
    if tile_frames == "auto" or tile_frames is None:
        even = max(1, -(-int(n_frames) // world))
        # RANSAC: interleave, but keep a step at ~16 rounds (a launch + a gather per round)
        tile = max(RANSAC_TILE_FRAMES, -(-even // 16)) if ransac else even
    else:
        tile = int(tile_frames)
    # joint-instances per frame from GLOBAL data: a rank may own no frame at all
    plan0 = TilePlan(n_frames, world, tile, 1)
    mine = plan0.frames_of(rank).size
    n_local = local_points.shape[1]
    if world > 1:
        t = torch.tensor([n_local, mine], dtype=torch.int64)
        dev = None
        if dist.get_backend(group) == "nccl":
            dev = torch.device("cuda", torch.cuda.current_device())
            t = t.to(dev)
        dist.all_reduce(t, group=group)
        tot_points, tot_frames = int(t[0]), int(t[1])
    else:
        tot_points, tot_frames = n_local, mine
    assert tot_frames == int(n_frames), "tile plan does not cover the recording"
    per_guess = tot_points // int(n_frames) if n_frames else 0
    assert per_guess * int(n_frames) == tot_points and n_local == mine * per_guess, \
        "local point count does not match this rank's frames"
    plan = TilePlan(n_frames, world, tile, per_guess)

    def run(pts):
        if ransac:
            p3d, _, _, err = cgroup.triangulate_ransac(pts, min_cams=min_cams)
        else:
            p3d, err = cgroup.triangulate_with_error(pts)
        return p3d, err

    if world == 1 or not gather:
        return run(local_points)

    as_numpy = isinstance(local_points, np.ndarray)
    backend = dist.get_backend(group)
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    if exchange == "auto":
        exchange = "peer" if backend == "nccl" else "gather"
    assert exchange in ("peer", "gather") and (exchange == "gather" or backend == "nccl"), \
        "the peer window needs device memory (NCCL runs)"
    if exchange == "peer":
        pr = PeerResults(plan, [(3,), ()], torch.float64, device, dst=0, group=group)
        for j in range(plan.rounds):
            a, b = plan.tile_span(j * world + rank)
            if b > a:
                off = plan.local_offset(rank, j)
                p3d, err = run(local_points[:, off:off + (b - a) * plan.per])
                pr.push(j, [torch.as_tensor(p3d).to(device).contiguous(), torch.as_tensor(err).to(device).contiguous()])
        out = pr.finish()
        res = (None, None)
        if rank == 0:
            res = (out[0].clone(), out[1].clone())     # the window is released below
        pr.close()
        if rank == 0 and as_numpy:
            return res[0].cpu().numpy(), res[1].cpu().numpy()
        return res
    rg = RoundGather(plan, [(3,), ()], torch.float64, device, dst=0, group=group)
    for j in range(plan.rounds):
        a, b = plan.tile_span(j * world + rank)
        res = None
        if b > a:
            off = plan.local_offset(rank, j)
            p3d, err = run(local_points[:, off:off + (b - a) * plan.per])
            res = [torch.as_tensor(p3d).to(device), torch.as_tensor(err).to(device)]
        rg.add(j, res)
    out = rg.finish()
    if rank != 0:
        return None, None
    if as_numpy:
        return out[0].cpu().numpy(), out[1].cpu().numpy()
    return out[0], out[1]
