"""Frame-chunk sharding of the hot path over the GPUs of one box.

Every joint-instance is independent, so ranks own contiguous frame ranges and never
exchange data while computing; the only collective is the final gather of the 3D results
(p3d + err = 32 B per joint-instance; ``picked`` optionally) to one rank.  Works with any
``torch.distributed`` backend (NCCL over NVLink on the B200 box, gloo in the CPU tests).
"""
import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def frame_range(n_frames, rank, world_size):
    """Contiguous, balanced frame range [lo, hi) of ``rank`` (first n_frames % world ranks get
    one extra frame)."""
    base, rem = divmod(int(n_frames), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def tile_frames_of(n_frames, rank, world_size, tile):
    """Frames of ``rank`` when tiles of ``tile`` consecutive frames are dealt round-robin
    (tile t belongs to rank t % world_size): the layout for the subset RANSAC, whose per-point
    cost is heavy-tailed and clusters in time (occlusions), SURVEY.md §8e."""
    tile = max(1, int(tile))
    starts = np.arange(rank * tile, int(n_frames), tile * int(world_size))
    if starts.size == 0:
        return np.zeros(0, dtype=np.int64)
    return np.concatenate([np.arange(a, min(a + tile, int(n_frames))) for a in starts]).astype(np.int64)


def _frames_of(n_frames, rank, world_size, tile):
    if tile is None:
        lo, hi = frame_range(n_frames, rank, world_size)
        return np.arange(lo, hi, dtype=np.int64)
    return tile_frames_of(n_frames, rank, world_size, tile)


def shard_points(points, n_frames, rank, world_size, tile=None):
    """Slice the (C, F*P, 2) observation array (frame-major point order) to this rank's frames:
    one contiguous range (``tile`` None) or round-robin tiles of ``tile`` frames."""
    n = points.shape[1]
    assert n % n_frames == 0, "point count is not a multiple of the frame count"
    per = n // n_frames
    if tile is None:
        lo, hi = frame_range(n_frames, rank, world_size)
        return points[:, lo * per:hi * per]
    fr = tile_frames_of(n_frames, rank, world_size, tile)
    idx = (fr[:, None] * per + np.arange(per)[None, :]).reshape(-1)
    if torch is not None and isinstance(points, torch.Tensor):
        return points[:, torch.as_tensor(idx, device=points.device)]
    return points[:, idx]


def gather_results(tensors, n_frames, dst=0, group=None, tile=None):
    """Gather per-rank result tensors (first dim = this rank's joint-instances, frame-major)
    to ``dst`` in global frame order.  Returns the concatenated tensors on ``dst`` and None
    elsewhere.  Uneven shards are handled by padding to the largest shard; round-robin tiles
    (``tile``) are put back into frame order on ``dst``."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    frames = [_frames_of(n_frames, r, world, tile) for r in range(world)]
    outs = []
    for t in tensors:
        per = t.shape[0] // max(1, frames[rank].size)
        sizes = [f.size * per for f in frames]
        mx = max(sizes)
        pad = t
        if t.shape[0] < mx:
            pad = torch.cat([t, t.new_zeros((mx - t.shape[0],) + tuple(t.shape[1:]))])
        pad = pad.contiguous()
        if rank == dst:
            bufs = [torch.empty_like(pad) for _ in range(world)]
            dist.gather(pad, bufs, dst=dst, group=group)
            res = torch.cat([b[:s] for b, s in zip(bufs, sizes)])
            if tile is not None:
                order = np.argsort(np.concatenate(frames), kind="stable")          # rank-major -> frame order
                idx = (order[:, None] * per + np.arange(per)[None, :]).reshape(-1)
                res = res[torch.as_tensor(idx, device=res.device)]
            outs.append(res)
        else:
            dist.gather(pad, None, dst=dst, group=group)
            outs.append(None)
    return outs


RANSAC_TILE_FRAMES = 256


def triangulate_sharded(cgroup, points, n_frames, ransac=False, min_cams=2, gather=True, group=None,
                        tile_frames="auto"):
    """Run this rank's frames of ``points`` (C, F*P, 2, the full array on every rank) through the
    GPU CameraGroup and gather (p3d, err) to rank 0.  Returns (p3d, err) on rank 0 in global
    frame order (or the local shard when ``gather`` is False).  ``tile_frames``: None = one
    contiguous frame range per rank; an int = round-robin tiles of that many frames; "auto" =
    contiguous for the DLT path, tiles of RANSAC_TILE_FRAMES for the subset RANSAC."""
    world = dist.get_world_size(group) if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    tile = (RANSAC_TILE_FRAMES if ransac else None) if tile_frames == "auto" else tile_frames
    if world == 1:
        tile = None
    local = shard_points(points, n_frames, rank, world, tile)
    if ransac:
        p3d, _, _, err = cgroup.triangulate_ransac(local, min_cams=min_cams)
    else:
        p3d, err = cgroup.triangulate_with_error(local)
    if world == 1 or not gather:
        return p3d, err
    as_t = [torch.as_tensor(p3d), torch.as_tensor(err)]
    g = gather_results(as_t, n_frames, dst=0, group=group, tile=tile)
    if rank != 0:
        return None, None
    if isinstance(p3d, np.ndarray):
        return g[0].cpu().numpy(), g[1].cpu().numpy()
    return g[0], g[1]
