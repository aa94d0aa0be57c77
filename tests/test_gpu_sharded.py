"""GPU tier, N > 1: the frame-sharded subset RANSAC (one process per GPU, round-robin tiles, results
delivered into rank 0's frame-ordered arrays — by copy-engine writes into its peer window over NVLink,
or by a per-round NCCL gather) returns bit-for-bit what one GPU returns for the whole recording.  Needs two visible GPUs (the driver's single-GPU box skips it; run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu`)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, tile, ransac, exchange, tmp):
    import torch
    import torch.distributed as dist
    from macaque_3d_pose_estimation_b200 import sharding, synth
    from macaque_3d_pose_estimation_b200.cameras import CameraGroup
    from oracle import cameragroup as og
    from oracle import fixtures
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        dicts = synth.make_rig(8, "pinhole", seed=20261020)
        cg = CameraGroup.from_dicts(dicts)
        cg.device = rank
        cams = fixtures.cams_from_dicts(dicts)
        X = synth.make_tracks(n_frames, 2, seed=3).reshape(-1, 3)
        p2 = synth.corrupt(og.project(cams, X), seed=3, p_outlier=0.2 if ransac else 0.0, p_missing=0.1)
        tl = tile
        if tl == "auto":
            even = max(1, -(-n_frames // world))
            tl = max(sharding.RANSAC_TILE_FRAMES, -(-even // 16)) if ransac else even
        local = torch.from_numpy(np.ascontiguousarray(sharding.shard_points(p2, n_frames, rank, world, tl))).cuda()
        p3d, err = sharding.triangulate_sharded(cg, local, n_frames, ransac=ransac, tile_frames=tile, exchange=exchange)
        if rank == 0:
            full = torch.from_numpy(p2).cuda()
            if ransac:
                r3, _, _, re_ = cg.triangulate_ransac(full)
            else:
                r3, re_ = cg.triangulate_with_error(full)
            np.savez(tmp, ok3=bool(torch.equal(p3d.nan_to_num(), r3.nan_to_num())),
                     oke=bool(torch.equal(err.nan_to_num(), re_.nan_to_num())), n=int(p3d.shape[0]))
        else:
            assert p3d is None and err is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["peer", "gather"])
@pytest.mark.parametrize("n_frames,tile,ransac", [(3000, 97, True), (700, "auto", True), (1000, "auto", False), (5, 256, True),
                                                  (1, 256, True)])
def test_nccl_sharded_equals_single_gpu(tmp_path, n_frames, tile, ransac, exchange):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 visible GPUs (this box has %d): NCCL multi-rank parity is run with gpurun --gpus 2"
                    % torch.cuda.device_count())
    import __graft_entry__ as ge
    ge.build_library()
    tmp = str(tmp_path / "out.npz")
    mp.spawn(_worker, args=(2, _free_port(), n_frames, tile, ransac, exchange, tmp), nprocs=2, join=True)
    r = np.load(tmp)
    assert r["n"] == n_frames * 34
    assert bool(r["ok3"]) and bool(r["oke"]), "sharded result differs from the single-GPU result"


def _span_case():
    from macaque_3d_pose_estimation_b200 import synth
    from macaque_3d_pose_estimation_b200.cameras import CameraGroup
    from oracle import cameragroup as og
    from oracle import fixtures
    import __graft_entry__ as ge
    ge.build_library()
    dicts = synth.make_rig(8, "pinhole", seed=20261020)
    cams = fixtures.cams_from_dicts(dicts)
    X = synth.make_tracks(700, 2, seed=9).reshape(-1, 3)
    p2 = synth.corrupt(og.project(cams, X), seed=9, p_outlier=0.2, p_missing=0.1)
    return CameraGroup.from_dicts(dicts), p2


def test_host_span_entry_points_cover_the_arrays():
    """m3d_triangulate_*_host_span: disjoint spans of one set of host arrays, processed by separate calls
    (two threads on one GPU here), give exactly the single-call result."""
    cg, p2 = _span_case()
    cg.device = 0
    ref3, refe = cg.triangulate_with_error(p2)
    rr = cg.triangulate_ransac(p2, min_cams=2, return_stats=True)
    cg2, _ = _span_case()
    cg2._host_devices = lambda n: [0, 0, 0]                    # three spans, one device
    got3, gote = cg2.triangulate_with_error(p2)
    assert np.array_equal(got3, ref3, equal_nan=True) and np.array_equal(gote, refe, equal_nan=True)
    gr = cg2.triangulate_ransac(p2, min_cams=2, return_stats=True)
    for a, b in zip(rr, gr):
        assert np.array_equal(a, b, equal_nan=True)


def test_single_process_spreads_numpy_input_over_all_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 visible GPUs (this box has %d)" % torch.cuda.device_count())
    cg, p2 = _span_case()
    cg.device = 0
    ref3, refe = cg.triangulate_with_error(p2)
    cg2, _ = _span_case()
    cg2.MULTI_GPU_MIN_POINTS = 1000
    assert cg2._host_devices(p2.shape[1]) == list(range(torch.cuda.device_count()))
    got3, gote = cg2.triangulate_with_error(p2)
    assert np.array_equal(got3, ref3, equal_nan=True) and np.array_equal(gote, refe, equal_nan=True)
    rr = cg.triangulate_ransac(p2, return_stats=True)
    gr = cg2.triangulate_ransac(p2, return_stats=True)
    for a, b in zip(rr, gr):
        assert np.array_equal(a, b, equal_nan=True)
    assert len(cg2._rig_cache[1]) == torch.cuda.device_count()   # one rig per GPU


def test_peer_window_api_on_one_gpu():
    """m3d_peer_alloc / push / free on a single GPU (the owner's side of the result window; the mapping side
    needs a second process and is covered by the 2-GPU test above): the window is exported, copies land
    where they are pushed, and opening one's own handle fails loudly instead of aliasing."""
    import ctypes
    import torch
    from macaque_3d_pose_estimation_b200 import _lib
    import __graft_entry__ as ge
    ge.build_library()
    lib = _lib.require_gpu()
    base = ctypes.c_void_p()
    handle = (ctypes.c_uint8 * 64)()
    n = 100000
    _lib.check(lib.m3d_peer_alloc(0, 8 * n, ctypes.byref(base), handle), "m3d_peer_alloc")
    try:
        assert base.value and any(bytes(handle))
        src = torch.arange(n, dtype=torch.float64, device="cuda:0")
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.m3d_peer_push(ctypes.c_void_p(base.value + 8 * 1000), ctypes.c_void_p(src.data_ptr()),
                                     8 * (n - 1000), ctypes.c_void_p(st)), "m3d_peer_push")
        from macaque_3d_pose_estimation_b200.sharding import _DeviceBytes
        win = torch.as_tensor(_DeviceBytes(base.value, 8 * n), device="cuda:0").view(torch.float64)
        torch.cuda.synchronize()
        assert torch.equal(win[1000:], src[:n - 1000])
        assert lib.m3d_peer_push(None, ctypes.c_void_p(src.data_ptr()), 8, ctypes.c_void_p(st)) != 0
        other = ctypes.c_void_p()
        rc = lib.m3d_peer_open(0, handle, ctypes.byref(other))             # own handle: CUDA refuses
        if rc == 0:
            _lib.check(lib.m3d_peer_close(0, other), "m3d_peer_close")
        else:
            assert "cudaIpcOpenMemHandle" in _lib.last_error()
        # a refused mapping leaves no stale error behind: the next launch reports its own status
        from macaque_3d_pose_estimation_b200 import synth
        from macaque_3d_pose_estimation_b200.cameras import CameraGroup
        cg = CameraGroup.from_dicts(synth.make_rig(8, "pinhole", seed=1))
        p3d = cg.triangulate(torch.zeros((8, 4, 2), dtype=torch.float64, device="cuda:0"))
        assert p3d.shape == (4, 3)
        del win
    finally:
        _lib.check(lib.m3d_peer_free(0, base), "m3d_peer_free")
