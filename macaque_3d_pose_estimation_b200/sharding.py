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


def shard_points(points, n_frames, rank, world_size):
    """Slice the (C, F*P, 2) observation array (frame-major point order) to this rank's frames."""
    n = points.shape[1]
    assert n % n_frames == 0, "point count is not a multiple of the frame count"
    per = n // n_frames
    lo, hi = frame_range(n_frames, rank, world_size)
    return points[:, lo * per:hi * per]


def gather_results(tensors, n_frames, dst=0, group=None):
    """Gather per-rank result tensors (first dim = this rank's joint-instances, frame-major)
    to ``dst`` in global frame order.  Returns the concatenated tensors on ``dst`` and None
    elsewhere.  Uneven shards are handled by padding to the largest shard."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    outs = []
    for t in tensors:
        per = t.shape[0] // max(1, (frame_range(n_frames, rank, world)[1] - frame_range(n_frames, rank, world)[0]))
        sizes = [(frame_range(n_frames, r, world)[1] - frame_range(n_frames, r, world)[0]) * per for r in range(world)]
        mx = max(sizes)
        pad = t
        if t.shape[0] < mx:
            pad = torch.cat([t, t.new_zeros((mx - t.shape[0],) + tuple(t.shape[1:]))])
        pad = pad.contiguous()
        if rank == dst:
            bufs = [torch.empty_like(pad) for _ in range(world)]
            dist.gather(pad, bufs, dst=dst, group=group)
            outs.append(torch.cat([b[:s] for b, s in zip(bufs, sizes)]))
        else:
            dist.gather(pad, None, dst=dst, group=group)
            outs.append(None)
    return outs


def triangulate_sharded(cgroup, points, n_frames, ransac=False, min_cams=2, gather=True, group=None):
    """Run this rank's frame chunk of ``points`` (C, F*P, 2; the full array on every rank or a
    callable ``points(lo_pt, hi_pt)`` producing the shard) through the GPU CameraGroup and
    gather (p3d, err) to rank 0.  Returns (p3d, err) on rank 0 (or the local shard when
    ``gather`` is False)."""
    world = dist.get_world_size(group) if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    local = shard_points(points, n_frames, rank, world)
    if ransac:
        p3d, _, _, err = cgroup.triangulate_ransac(local, min_cams=min_cams)
    else:
        p3d, err = cgroup.triangulate_with_error(local)
    if world == 1 or not gather:
        return p3d, err
    as_t = [torch.as_tensor(p3d), torch.as_tensor(err)]
    g = gather_results(as_t, n_frames, dst=0, group=group)
    if rank != 0:
        return None, None
    if isinstance(p3d, np.ndarray):
        return g[0].cpu().numpy(), g[1].cpu().numpy()
    return g[0], g[1]
