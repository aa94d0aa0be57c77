#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 multi-view 3D reconstruction hot path.

    python bench.py --gpus N --steps K --warmup W [--workload dlt|ransac] [--impl reference]

metric  : triangulated joint-instances/sec (BASELINE.json)
workload: "dlt"    = BASELINE config 2: 8-view undistort + DLT triangulate + mean reprojection
                     error, 4 macaques x 17 joints x 1M frames = 6.8e7 joint-instances per GPU
          "ransac" = BASELINE config 3: 8-view triangulate_ransac (all subsets, min 2 views),
                     20 % outlier detections, frame-sharded
A step is one pass of the hot path over the whole resident batch.  Inputs (8.7 GB for
"dlt") are far larger than the 126 MB L2, so no flush is needed between iterations.
N > 1: one process per GPU (torchrun), weak scaling (each rank owns its own frame chunk of
the same size), no data-path collective; the NCCL gather of the 3D results to rank 0 is
timed separately ("gather").

`--impl reference` times the CPU port of the reference's NumPy/OpenCV path (oracle/) on
the host cores for the same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HBM_FALLBACK_GBS = 6650.0
BYTES_PER_INSTANCE = {"dlt": lambda C: 16 * C + 32, "ransac": lambda C: 33 * C + 32}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self, t0, t1):
        """wall-clock window of the timed region: only samples inside it are reported"""
        self.window = (t0, t1)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        win = getattr(self, "window", None)
        inside = [ln for (t, ln) in self.lines if win and win[0] - 0.02 <= t <= win[1] + 0.02]
        use = inside if len(inside) >= 2 else [ln for (_, ln) in self.lines]
        for ln in use:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# synthetic workload, generated on the device
# ------------------------------------------------------------------------------------------

def make_device_workload(cg, n_frames, n_animals, n_joints, seed, workload, device):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    f64 = torch.float64
    root0 = torch.rand((n_animals, 3), generator=g, device=device, dtype=f64)
    # animal roots stay inside the volume every camera of the ring rig sees (the cage centre)
    lo = torch.tensor([-600.0, -600.0, 0.0], device=device, dtype=f64)
    hi = torch.tensor([600.0, 600.0, 800.0], device=device, dtype=f64)
    root0 = lo + root0 * (hi - lo)
    steps = torch.randn((n_frames, n_animals, 3), generator=g, device=device, dtype=f64) * 15.0
    root = root0[None] + torch.cumsum(steps, dim=0)
    del steps
    # keep the animals in the cage: reflect the walk into the box
    span = hi - lo
    root = lo + (span - ((root - lo) % (2 * span) - span).abs()).abs()
    skel = torch.randn((n_animals, n_joints, 3), generator=g, device=device, dtype=f64) * 120.0
    X = (root[:, :, None, :] + skel[None]).reshape(-1, 3).contiguous()
    del root
    xy = cg.project(X)                                         # our own projection kernel
    C, N = xy.shape[0], xy.shape[1]
    W, H = synth_image_size()
    for c in range(C):                                         # plane by plane: bounded temporaries
        # a camera does not detect what falls outside its frame (also keeps the polynomial
        # distortion model inside its monotone range)
        off = (xy[c, :, 0] < 0) | (xy[c, :, 0] > W) | (xy[c, :, 1] < 0) | (xy[c, :, 1] > H)
        xy[c][off] = float("nan")
        xy[c] += torch.randn((N, 2), generator=g, device=device, dtype=f64) * 0.3
        if workload == "ransac":
            o = torch.rand((N,), generator=g, device=device) < 0.2
            xy[c] += o[:, None] * torch.randn((N, 2), generator=g, device=device, dtype=f64) * 60.0
        m = torch.rand((N,), generator=g, device=device) < 0.1
        xy[c][m] = float("nan")
    return X, xy


# ------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference's NumPy/OpenCV path)
# ------------------------------------------------------------------------------------------

def _cpu_chunk(args):
    dicts, p2d, workload = args
    import numpy as np
    from oracle import cameragroup as og
    from oracle import fixtures
    cams = fixtures.cams_from_dicts(dicts)
    if workload == "dlt":
        p3d = og.triangulate_loops(cams, p2d)
        og.reprojection_error_loops(cams, p3d, p2d, mean=True)
    else:
        og.triangulate_ransac_loops(cams, p2d, min_cams=2)
    return p2d.shape[1]


def cpu_workload(workload, n_points, seed, n_cams=8):
    import numpy as np
    from macaque_3d_pose_estimation_b200 import synth
    from oracle import cameragroup as og
    from oracle import fixtures
    dicts = synth.make_rig(n_cams, "pinhole", seed=seed)
    cams = fixtures.cams_from_dicts(dicts)
    n_frames = max(1, n_points // (2 * 17))
    X = synth.make_tracks(n_frames, 2, seed=seed).reshape(-1, 3)[:n_points]
    p2d = synth.corrupt(og.project(cams, X), seed=seed, p_outlier=0.2 if workload == "ransac" else 0.0,
                        p_missing=0.1)
    return dicts, p2d


def time_cpu(workload, n_points, procs, seed=20261018, n_cams=8):
    """joint-instances/s of the loop-faithful port on `procs` host processes."""
    import numpy as np
    dicts, p2d = cpu_workload(workload, n_points, seed, n_cams)
    n = p2d.shape[1]
    if procs <= 1:
        _cpu_chunk((dicts, p2d[:, :8], workload))              # warm caches / imports
        t0 = time.perf_counter()
        _cpu_chunk((dicts, p2d, workload))
        return n / (time.perf_counter() - t0), n
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    chunks = np.array_split(np.arange(n), procs)
    jobs = [(dicts, np.ascontiguousarray(p2d[:, c]), workload) for c in chunks if c.size]
    with ctx.Pool(procs) as pool:
        pool.map(_cpu_chunk, [(dicts, p2d[:, :8].copy(), workload)] * procs)   # spawn + import warm-up
        t0 = time.perf_counter()
        pool.map(_cpu_chunk, jobs)
        dt = time.perf_counter() - t0
    return n / dt, n


def run_reference(args):
    """--impl reference: the CPU port on all host cores, same metric / config keys."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    sample = 2 * 17 * (1000 if args.workload == "dlt" else 6) * max(1, procs // 2)
    vals = []
    t_all = time.perf_counter()
    for i in range(args.steps):
        v, n = time_cpu(args.workload, sample, procs, n_cams=args.cameras)
        vals.append(v)
        if time.perf_counter() - t_all > 240:
            break
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "triangulated joint-instances/sec", "value": value,
        "unit": "joint-instances/s", "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
        "ms_per_step": 1e3 * n / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, args.cameras, args.frames * 4 * 17),     # the workload of our arm; see cpu_baseline.sample
        "cpu_baseline": {"value": value, "unit": "joint-instances/s", "cores": procs, "kind": "port",
                         "sample": "%d joint-instances per step (cfg-1 rig: pinhole cameras, 2 animals x 17 "
                                   "joints), loop-faithful NumPy/OpenCV port of the reference in oracle/, "
                                   "%d spawn processes" % (n, procs)},
        "e2e": {"value": value, "unit": "joint-instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def config_dict(args, C, n_per_gpu):
    names = {"dlt": "cfg2: %d-view undistort + DLT triangulate + mean reprojection_error, 4 macaques x 17 joints",
             "ransac": "cfg3: %d-view triangulate_ransac (all camera subsets, min_cams=2), 20%% outlier detections"}
    return {"workload": names[args.workload] % C, "cameras": C, "camera_model": "pinhole(5 coeff)",
            "joint_instances_per_gpu": int(n_per_gpu), "frames_per_gpu": int(args.frames),
            "animals": 4, "joints": 17, "missing_views": 0.1,
            "outliers": 0.2 if args.workload == "ransac" else 0.0,
            "l2_policy": "inputs larger than L2 (no flush needed)", "parallelism": "frame-sharded dp%d" % args.gpus}


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------

def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.build()
    from macaque_3d_pose_estimation_b200 import _lib, synth
    from macaque_3d_pose_estimation_b200.cameras import CameraGroup

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.require_gpu()

    C, A, J = args.cameras, 4, 17
    F = args.frames
    seed = 20261018 + 2 + rank
    cg = CameraGroup.from_dicts(synth.make_rig(C, "pinhole", seed=20261018 + 2))
    cg.device = local
    X, xy = make_device_workload(cg, F, A, J, seed, args.workload, device)
    del X
    N = xy.shape[1]
    p3d = torch.empty((N, 3), dtype=torch.float64, device=device)
    err = torch.empty((N,), dtype=torch.float64, device=device)
    picked = torch.empty((C, N), dtype=torch.uint8, device=device) if args.workload == "ransac" else None
    xyp = torch.empty((C, N, 2), dtype=torch.float64, device=device) if args.workload == "ransac" else None
    nev = torch.empty((N,), dtype=torch.int32, device=device) if args.workload == "ransac" else None
    rig = cg._rig(local)
    stream = torch.cuda.current_stream(device)
    sp = lambda t: None if t is None else t.data_ptr()

    def step():
        if args.workload == "dlt":
            rc = lib.m3d_triangulate_error(rig.handle, xy.data_ptr(), N, 1, p3d.data_ptr(), err.data_ptr(),
                                           stream.cuda_stream)
        else:
            rc = lib.m3d_triangulate_ransac(rig.handle, xy.data_ptr(), N, 1, 2, 0.5, 200.0, p3d.data_ptr(),
                                            sp(picked), sp(xyp), err.data_ptr(), None, sp(nev),
                                            stream.cuda_stream)
        _lib.check(rc, "bench step")

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = lib.m3d_launch_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    t_wall0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    sampler.mark(t_wall0, time.time())
    ms = e0.elapsed_time(e1)
    launches = lib.m3d_launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / args.steps
    value = world * N / (ms_per_step * 1e-3)

    # gather of the 3D results to rank 0 (the only collective of the path), timed separately
    gather_ms = None
    if world > 1:
        from macaque_3d_pose_estimation_b200 import sharding
        small = torch.zeros((world * 8,), dtype=torch.float64, device=device)
        dist.all_reduce(small)                                 # communicator set-up outside the timing
        out = sharding.gather_results([p3d, err], F, dst=0)    # one untimed gather (NCCL warm-up)
        del out
        torch.cuda.synchronize()
        dist.barrier()
        g0 = torch.cuda.Event(enable_timing=True)
        g1 = torch.cuda.Event(enable_timing=True)
        g0.record()
        out = sharding.gather_results([p3d, err], F, dst=0)
        g1.record()
        torch.cuda.synchronize()
        tg = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=device)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gather_ms = float(tg.item())
        del out

    peak, peak_kind = measured_peaks()
    bpi = BYTES_PER_INSTANCE[args.workload](C)
    achieved = bpi * N / (ms_per_step * 1e-3) / 1e9            # per-GPU kernel, GB/s
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_kind": peak_kind + " copy bandwidth",
                "kernel": "k_triangulate<undistort,err>" if args.workload == "dlt" else "k_ransac",
                "algorithmic_bytes_per_instance": bpi,
                "note": "path is fp64-ALU bound at reference precision (SURVEY 7); see fp64"}
    try:   # DRAM traffic per launch from the committed ncu capture (profiles/traffic.json)
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[args.workload]
        roofline["traffic"] = tj["dram_bytes_per_instance"] * N
        roofline["traffic_source"] = tj["source"]
    except Exception:
        pass
    tf = ctypes_double()
    if rank == 0 and lib.m3d_probe_fp64_tflops(local, ctypes_byref(tf)) == 0:
        roofline["fp64_peak_tflops_measured"] = tf.value
        try:   # the binding bound: fp64 pipe slots (one per fp64 instruction and lane) from the committed ncu capture
            per = float(tj["fp64_thread_instr_per_instance"])
            ach = per * N / (ms_per_step * 1e-3) / 1e12
            roofline["fp64"] = {"thread_instr_per_instance": per, "achieved": ach, "peak": tf.value / 2.0,
                                "unit": "T fp64 instr/s (per GPU)", "frac": ach / (tf.value / 2.0),
                                "source": tj.get("fp64_source")}
        except Exception:
            pass
    extra = {"mean_valid_views": float((~torch.isnan(xy[:, :, 0])).double().mean().item() * C)}
    if args.workload == "ransac":
        extra["mean_subsets_per_point"] = float(nev.double().mean().item())
        extra["selected_fraction"] = float((~torch.isnan(p3d[:, 0])).double().mean().item())

    # ---- e2e: host buffers through the C-ABI host pipeline (H2D + kernel + D2H per step), every
    # rank streams its own shard concurrently ----
    # N = 1: the whole workload; N > 1: a 2e7-instance slice per rank (bounds the pinned host
    # memory of 8 concurrent ranks), reported in joint_instances_per_gpu
    n_e2e = min(N, args.e2e_points) if args.e2e_points > 0 else (N if world == 1 else min(N, 20000000))
    e2e = None
    try:
        if args.no_e2e:
            raise RuntimeError("skipped (--no-e2e)")
        numa = bind_to_gpu_numa_node(local)     # host buffers on the memory next to this rank's GPU
        h_xy = torch.empty((C, n_e2e, 2), dtype=torch.float64, pin_memory=True)
        h_xy.copy_(xy[:, :n_e2e])
        h_p3d = torch.empty((n_e2e, 3), dtype=torch.float64, pin_memory=True)
        h_err = torch.empty((n_e2e,), dtype=torch.float64, pin_memory=True)
        h_pick = torch.empty((C, n_e2e), dtype=torch.uint8, pin_memory=True) if args.workload == "ransac" else None
        h_xyp = torch.empty((C, n_e2e, 2), dtype=torch.float64, pin_memory=True) if args.workload == "ransac" else None

        def e2e_step():
            if args.workload == "dlt":
                rc = lib.m3d_triangulate_error_host(rig.handle, h_xy.data_ptr(), n_e2e, 1, h_p3d.data_ptr(),
                                                    h_err.data_ptr())
            else:
                rc = lib.m3d_triangulate_ransac_host(rig.handle, h_xy.data_ptr(), n_e2e, 1, 2, 0.5, 200.0,
                                                     h_p3d.data_ptr(), sp(h_pick), sp(h_xyp), h_err.data_ptr(),
                                                     None, None)
            _lib.check(rc, "e2e step")
        e2e_step()
        torch.cuda.synchronize()
        ke = max(1, min(args.steps, 5))
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            e2e_step()
        dt = (time.perf_counter() - t0) / ke
        if world > 1:
            td = torch.tensor([dt], dtype=torch.float64, device=device)
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
            dt = float(td.item())
        # the pipeline's output equals the resident run
        assert torch.equal(h_p3d.nan_to_num(), p3d[:n_e2e].cpu().nan_to_num())
        d2h = n_e2e * 32 + (n_e2e * C * 17 if args.workload == "ransac" else 0)
        e2e = {"value": world * n_e2e / dt, "unit": "joint-instances/s", "h2d_bytes_per_step": n_e2e * C * 16,
               "d2h_bytes_per_step": d2h, "joint_instances_per_gpu": n_e2e, "ms_per_step": dt * 1e3,
               "api": "m3d_triangulate_%s_host (pinned host buffers, 3-slot H2D/kernel/D2H pipeline)"
                      % ("error" if args.workload == "dlt" else "ransac"), "n_gpus": world,
               "host_numa_node": numa}
        del h_xy, h_p3d, h_err
    except Exception as ex:  # pragma: no cover
        e2e = {"value": None, "unit": "joint-instances/s", "error": str(ex)[:200]}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- secondary line: BASELINE config 3 (subset RANSAC, 20 % outliers) on a 1e5-frame slice ----
    if args.workload == "dlt" and not args.no_ransac_extra:
        try:
            del xy
            torch.cuda.empty_cache()
            Fr = 100000
            _, xyr = make_device_workload(cg, Fr, A, J, seed + 100, "ransac", device)
            Nr = xyr.shape[1]
            r3 = torch.empty((Nr, 3), dtype=torch.float64, device=device)
            re_ = torch.empty((Nr,), dtype=torch.float64, device=device)
            rp = torch.empty((C, Nr), dtype=torch.uint8, device=device)
            rx = torch.empty((C, Nr, 2), dtype=torch.float64, device=device)
            rn = torch.empty((Nr,), dtype=torch.int32, device=device)

            def rstep():
                _lib.check(lib.m3d_triangulate_ransac(rig.handle, xyr.data_ptr(), Nr, 1, 2, 0.5, 200.0, r3.data_ptr(),
                                                      rp.data_ptr(), rx.data_ptr(), re_.data_ptr(), None, rn.data_ptr(),
                                                      stream.cuda_stream), "ransac step")
            for _ in range(3):
                rstep()
            torch.cuda.synchronize()
            r0 = torch.cuda.Event(enable_timing=True)
            r1 = torch.cuda.Event(enable_timing=True)
            r0.record(stream)
            for _ in range(5):
                rstep()
            r1.record(stream)
            torch.cuda.synchronize()
            rms = r0.elapsed_time(r1) / 5
            extra["ransac"] = {
                "workload": "cfg3: 8-view triangulate_ransac (all camera subsets, min_cams=2), 20% outlier detections",
                "value": Nr / (rms * 1e-3), "unit": "joint-instances/s (one GPU)", "joint_instances": Nr,
                "ms_per_step": rms, "mean_subsets_per_point": float(rn.double().mean().item()),
                "roofline_frac_hbm": 296 * Nr / (rms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_instance": 296}
        except Exception as ex:  # pragma: no cover
            extra["ransac"] = {"error": str(ex)[:200]}
        # ---- step-4 2D Viterbi filter on the series of the same recording shape ----
        try:
            from macaque_3d_pose_estimation_b200 import filter2d
            Sv, Fv = A * C * J, 20000
            det = torch.from_numpy(np.ascontiguousarray(
                synth.make_detection_series(Fv, 32, 1, seed).transpose(1, 0, 2, 3))).to(device)
            det = det.repeat((Sv + 31) // 32, 1, 1, 1)[:Sv].contiguous()
            for _ in range(2):
                filter2d.viterbi_series(det, 3, 25.0, 0.3)
            torch.cuda.synchronize()
            v0 = torch.cuda.Event(enable_timing=True)
            v1 = torch.cuda.Event(enable_timing=True)
            v0.record()
            for _ in range(3):
                filter2d.viterbi_series(det, 3, 25.0, 0.3)
            v1.record()
            torch.cuda.synchronize()
            vms = v0.elapsed_time(v1) / 3
            extra["viterbi_filter"] = {
                "workload": "step-4 2D Viterbi filter (n_back 3, offset_threshold 25, score_threshold 0.3), %d series "
                            "(animals x cameras x joints) x %d frames" % (Sv, Fv),
                "value": Sv * Fv / (vms * 1e-3), "unit": "series-frames/s (one GPU)", "ms_per_step": vms}
        except Exception as ex:  # pragma: no cover
            extra["viterbi_filter"] = {"error": str(ex)[:200]}

    # ---- CPU baseline: loop-faithful port, one core, bounded sample ---------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        n_cpu = 34000 if args.workload == "dlt" else 340
        v, n = time_cpu(args.workload, n_cpu, 1, n_cams=C)
        cpu = {"value": v, "unit": "joint-instances/s", "cores": 1, "kind": "port",
               "sample": "%d joint-instances of the cfg-1 rig (8 pinhole cameras), loop-faithful NumPy/OpenCV "
                         "port of the reference (oracle/cameragroup.py *_loops)" % n}

    line = {
        "metric": "triangulated joint-instances/sec", "value": value, "unit": "joint-instances/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, C, N), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if gather_ms is not None:
        line["gather"] = {"ms": gather_ms, "bytes_to_rank0": int((world - 1) * N * 32),
                          "what": "torch.distributed.gather (NCCL) of p3d+err to rank 0"}
    line.update(extra)
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def synth_image_size():
    from macaque_3d_pose_estimation_b200 import synth
    return synth.IMG_SIZE


def ctypes_double():
    import ctypes
    return ctypes.c_double(0.0)


def ctypes_byref(x):
    import ctypes
    return ctypes.byref(x)


def bind_to_gpu_numa_node(local):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that the pinned host
    buffers allocated afterwards are local to that GPU's PCIe root (with 8 ranks the H2D streams
    otherwise cross the socket interconnect).  Returns the node or None; never fails."""
    try:
        p = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", choices=["dlt", "ransac"], default="dlt")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU (default 1e6 dlt, 2e5 ransac)")
    ap.add_argument("--e2e-points", type=int, default=0, help="joint-instances of the e2e run (0 = all)")
    ap.add_argument("--cameras", type=int, default=8, help="cameras of the synthetic ring rig (BASELINE: 8; config 5: 16)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ransac-extra", action="store_true")
    args = ap.parse_args()
    if args.frames <= 0:
        args.frames = 1000000 if args.workload == "dlt" else 200000
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
