"""CPU tier: the N > 1 path — frame-chunk sharding and the result gather — with
world_size-2 gloo process groups.  The per-rank compute is a stand-in (the oracle) because
this container has no GPU; the sharding / gather logic under test is the product's
(macaque_3d_pose_estimation_b200/sharding.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from macaque_3d_pose_estimation_b200 import sharding, synth
from oracle import cameragroup as og
from oracle import fixtures


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleGroup:
    """Stand-in for the GPU CameraGroup on a CPU-only box (test infrastructure)."""

    def __init__(self, dicts):
        self.cameras = fixtures.cams_from_dicts(dicts)

    def triangulate_with_error(self, pts):
        p3d = og.triangulate(self.cameras, pts)
        return p3d, og.reprojection_error(self.cameras, p3d, pts, mean=True)

    def triangulate_ransac(self, pts, min_cams=2):
        return og.triangulate_ransac(self.cameras, pts, min_cams=min_cams)


def _worker(rank, world, port, n_frames, ransac, tmp, tile="auto"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dicts = synth.make_rig(4, "pinhole", seed=31)
        cg = OracleGroup(dicts)
        X = synth.make_tracks(n_frames, 2, n_joints=5, seed=31).reshape(-1, 3)
        p2 = synth.corrupt(og.project(cg.cameras, X), seed=31, p_outlier=0.2 if ransac else 0.0, p_missing=0.1)
        # every rank is handed the observations of its own frames only
        tl = tile
        if tl == "auto":
            even = max(1, -(-n_frames // world))
            tl = max(sharding.RANSAC_TILE_FRAMES, -(-even // 16)) if ransac else even
        local = sharding.shard_points(p2, n_frames, rank, world, tl)
        p3d, err = sharding.triangulate_sharded(cg, local, n_frames, ransac=ransac, tile_frames=tile)
        if rank == 0:
            np.savez(tmp, p3d=p3d, err=err, p2=p2)
        else:
            assert p3d is None and err is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames,ransac,tile", [(11, False, "auto"), (6, True, "auto"), (7, True, 2), (9, False, 1),
                                                  (1, True, 256), (5, True, 3), (0, False, "auto")])
def test_sharded_matches_single_process(tmp_path, n_frames, ransac, tile):
    tmp = str(tmp_path / "out.npz")
    mp.spawn(_worker, args=(2, _free_port(), n_frames, ransac, tmp, tile), nprocs=2, join=True)
    r = np.load(tmp)
    cg = OracleGroup(synth.make_rig(4, "pinhole", seed=31))
    if ransac:
        p3d, _, _, err = cg.triangulate_ransac(r["p2"])
    else:
        p3d, err = cg.triangulate_with_error(r["p2"])
    assert np.array_equal(r["p3d"], p3d, equal_nan=True)     # uneven shards, global frame order
    assert np.array_equal(r["err"], err, equal_nan=True)


def test_round_robin_tiles_partition():
    for n in (0, 1, 7, 1000):
        for w in (1, 2, 3, 8):
            for t in (1, 3, 256):
                fr = [sharding.tile_frames_of(n, r, w, t) for r in range(w)]
                assert np.array_equal(np.sort(np.concatenate(fr)), np.arange(n))
                plan = sharding.TilePlan(n, w, t, 3)
                # a round is a contiguous block of the recording; a short clip still feeds every rank
                assert plan.tile <= max(1, -(-n // w))
                for r in range(w):
                    offs = [plan.local_offset(r, j) for j in range(plan.rounds + 1)]
                    assert offs[-1] == 3 * fr[r].size
                assert all((f[1:] > f[:-1]).all() for f in fr if f.size > 1)


def test_frame_ranges_partition():
    for n in (0, 1, 7, 8, 1000001):
        for w in (1, 2, 3, 8):
            spans = [sharding.frame_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
