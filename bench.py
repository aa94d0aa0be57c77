#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 multi-view 3D reconstruction hot path.

    python bench.py --gpus N --steps K --warmup W [--workload ransac|dlt] [--impl reference]

metric  : triangulated joint-instances/sec (BASELINE.json)
workload: "ransac" (default, the configuration the north-star target is quoted on) = BASELINE
                     config 3: 8-view triangulate_ransac (all camera subsets, min 2 views), 20 %
                     outlier 2D detections, 4 macaques x 17 joints x 1M frames per GPU, frame-sharded
                     in round-robin tiles; the NCCL gather of the 3D results to rank 0 runs per tile
                     round INSIDE the timed step, overlapped with the next round's kernels
          "dlt"    = BASELINE config 2: 8-view undistort + DLT triangulate + mean reprojection
                     error, 4 macaques x 17 joints x 1M frames per GPU (no collective on this path:
                     every rank returns its own frame range)
The other workload is measured in the same run and reported as a complete second object
("cfg2" / "cfg3": value, roofline, e2e, cpu_baseline).
A step is one pass of the hot path over the whole resident batch (inputs of 8.7 GB per GPU are far
larger than the 126 MB L2, so no flush is needed between iterations).  N > 1: one process per GPU
(torchrun), weak scaling (each rank owns the same number of frames).

`--impl reference` times the loop-faithful CPU port of the reference's NumPy/OpenCV path (oracle/)
on all host cores for the same metric and workload; its `config` states the sample it really ran.
"""
import argparse
import ctypes
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
ANIMALS, JOINTS = 4, 17
NAMES = {"dlt": "cfg2: %d-view undistort + DLT triangulate + mean reprojection_error, 4 macaques x 17 joints",
         "ransac": "cfg3: %d-view triangulate_ransac (all camera subsets, min_cams=2), 20%% outlier detections, "
                   "4 macaques x 17 joints"}


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
        self.windows = []

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
        """wall-clock window of a timed region: only samples inside the windows are reported"""
        self.windows.append((t0, t1))

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
        inside = [ln for (t, ln) in self.lines if any(a - 0.02 <= t <= b + 0.02 for a, b in self.windows)]
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
    from macaque_3d_pose_estimation_b200 import synth
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
    del X
    C, N = xy.shape[0], xy.shape[1]
    W, H = synth.IMG_SIZE
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
    return xy


# ------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference's NumPy/OpenCV path)
# ------------------------------------------------------------------------------------------

def _cpu_chunk(args):
    dicts, p2d, workload = args
    from oracle import cameragroup as og
    from oracle import fixtures
    cams = fixtures.cams_from_dicts(dicts)
    if workload == "dlt":
        p3d = og.triangulate_loops(cams, p2d)
        og.reprojection_error_loops(cams, p3d, p2d, mean=True)
    else:
        og.triangulate_ransac_loops(cams, p2d, min_cams=2)
    return p2d.shape[1]


def cpu_workload(workload, n_frames, seed, n_cams=8):
    """(rig dicts, (C, n_frames * 4 * 17, 2)): the numpy twin of make_device_workload (same rig, same
    walk / skeleton / in-frame / noise / outlier / missing-view model), generated on the host."""
    import numpy as np
    from macaque_3d_pose_estimation_b200 import synth
    from oracle import cameragroup as og
    from oracle import fixtures
    dicts = synth.make_rig(n_cams, "pinhole", seed=20261018 + 2)
    cams = fixtures.cams_from_dicts(dicts)
    rng = np.random.default_rng(seed)
    lo, hi = np.array([-600.0, -600.0, 0.0]), np.array([600.0, 600.0, 800.0])
    span = hi - lo
    root = lo + rng.random((ANIMALS, 3)) * span + np.cumsum(rng.normal(0, 15.0, (n_frames, ANIMALS, 3)), axis=0)
    root = lo + np.abs(span - np.abs((root - lo) % (2 * span) - span))
    X = (root[:, :, None, :] + rng.normal(0, 120.0, (ANIMALS, JOINTS, 3))[None]).reshape(-1, 3)
    xy = og.project(cams, X)
    W, H = synth.IMG_SIZE
    N = xy.shape[1]
    for c in range(n_cams):
        off = (xy[c, :, 0] < 0) | (xy[c, :, 0] > W) | (xy[c, :, 1] < 0) | (xy[c, :, 1] > H)
        xy[c][off] = np.nan
        xy[c] += rng.normal(0, 0.3, (N, 2))
        if workload == "ransac":
            o = rng.random(N) < 0.2
            xy[c] += o[:, None] * rng.normal(0, 60.0, (N, 2))
        xy[c][rng.random(N) < 0.1] = np.nan
    return dicts, xy


class CpuPool:
    """A persistent spawn pool for the all-core timings (created once, warmed once)."""

    def __init__(self, procs, workload, n_cams):
        import multiprocessing as mp
        self.procs = procs
        self.pool = mp.get_context("spawn").Pool(procs)
        dicts, p2d = cpu_workload(workload, 1, 1, n_cams)
        self.pool.map(_cpu_chunk, [(dicts, p2d[:, :8].copy(), workload)] * procs)   # spawn + import warm-up

    def run(self, dicts, p2d, workload):
        import numpy as np
        chunks = np.array_split(np.arange(p2d.shape[1]), self.procs)
        jobs = [(dicts, np.ascontiguousarray(p2d[:, c]), workload) for c in chunks if c.size]
        t0 = time.perf_counter()
        self.pool.map(_cpu_chunk, jobs)
        return p2d.shape[1] / (time.perf_counter() - t0)

    def close(self):
        self.pool.close()
        self.pool.join()


def time_cpu_single(workload, n_frames, seed=20261018, n_cams=8):
    dicts, p2d = cpu_workload(workload, n_frames, seed, n_cams)
    _cpu_chunk((dicts, p2d[:, :8], workload))                  # warm caches / imports
    t0 = time.perf_counter()
    _cpu_chunk((dicts, p2d, workload))
    return p2d.shape[1] / (time.perf_counter() - t0), p2d.shape[1]


CPU_SAMPLE_FRAMES = {"dlt": 500, "ransac": 25}                 # single core: ~1.5 s / ~18 s
CPU_POOL_FRAMES_PER_PROC = {"dlt": 250, "ransac": 8}           # all cores: ~1 s / ~6 s per step


def cpu_baseline(workload, n_cams, with_all_cores):
    v, n = time_cpu_single(workload, CPU_SAMPLE_FRAMES[workload], n_cams=n_cams)
    out = {"value": v, "unit": "joint-instances/s", "cores": 1, "kind": "port",
           "sample": "%d joint-instances (%d frames x 4 animals x 17 joints, the bench rig and corruption model), "
                     "loop-faithful NumPy/OpenCV port of the reference (oracle/cameragroup.py *_loops)"
                     % (n, CPU_SAMPLE_FRAMES[workload])}
    if with_all_cores:
        procs = max(1, min(os.cpu_count() or 1, 64))
        pool = CpuPool(procs, workload, n_cams)
        try:
            dicts, p2d = cpu_workload(workload, CPU_POOL_FRAMES_PER_PROC[workload] * procs, 20261018 + 7, n_cams)
            out["all_cores"] = {"value": pool.run(dicts, p2d, workload), "cores": procs,
                                "sample": "%d joint-instances over %d spawn processes" % (p2d.shape[1], procs)}
        finally:
            pool.close()
    return out


def optim_line(cx, args):
    """Step-4 refinement (CameraGroup.optim_points, the default branch of the reference's 3D stage): one animal,
    8 views, the template's constraints and weights; GPU solver against the port of the reference's
    scipy.optimize.least_squares call on the same data (a shorter clip: its cost grows faster than linearly)."""
    import numpy as np
    from macaque_3d_pose_estimation_b200 import synth
    from oracle import cameragroup as og
    from oracle import fixtures
    from oracle import make_golden as mg
    from oracle import optim as oopt

    def clip(F, seed):
        dicts = synth.make_rig(args.cameras, "pinhole", seed=20261018 + 2)
        cams = fixtures.cams_from_dicts(dicts)
        rng = np.random.default_rng(seed)
        X = synth.make_tracks(F, 1, seed=seed)[:, 0] * np.array([0.6, 0.6, 0.5])
        X = X + 30.0 * np.sin(np.arange(F)[:, None, None] / 7.0 + rng.uniform(0, 6, (1, X.shape[1], 3)))
        p2 = synth.corrupt(og.project(cams, X.reshape(-1, 3)), seed=seed, noise=0.8, p_outlier=0.03,
                           sigma_outlier=25.0, p_missing=0.15)
        return cams, p2.reshape(args.cameras, F, X.shape[1], 2)
    kw = dict(constraints=mg.MACAQUE_CONSTRAINTS, constraints_weak=mg.MACAQUE_CONSTRAINTS_WEAK, scale_smooth=3,
              scale_length=5, scale_length_weak=2, n_deriv_smooth=2, reproj_error_threshold=3)
    out = {}
    try:
        F = 20000
        cams, pts = clip(F, 77)
        init = cx.cg.triangulate(pts.reshape(args.cameras, -1, 2)).reshape(F, -1, 3)
        cx.cg.optim_points(pts[:, :200], init[:200], **kw)                    # warm-up
        dt = None
        for _ in range(2):                                                     # best of two full-size runs (the
            t0 = time.perf_counter()                                           # first one also pays the allocations)
            new, jl, info = cx.cg.optim_points(pts, init, return_info=True, **kw)
            d1 = time.perf_counter() - t0
            dt = d1 if dt is None else min(dt, d1)
        out = {"workload": "optim_points: 1 animal x 17 joints x %d frames, %d views, 20 strong + 11 weak limb "
                           "constraints, n_deriv_smooth 2 (config_tmpl.toml defaults)" % (F, args.cameras),
               "value": F / dt, "unit": "frames/s (one GPU, host arrays in and out)", "seconds": dt,
               "cost0": info["cost0"], "cost": info["cost"], "lm_steps": info["lm_steps"],
               "cg_iterations": info["cg_iterations"], "residual_evaluations": info["evaluations"]}
        if not args.no_cpu:
            Fc = 200
            t0 = time.perf_counter()
            pn, pj, pcost = oopt.optim_points_port(cams, pts[:, :Fc], init[:Fc], **kw)
            dtc = time.perf_counter() - t0
            g2, gj, ginfo = cx.cg.optim_points(pts[:, :Fc], init[:Fc], return_info=True, **kw)
            out["cpu_baseline"] = {"value": Fc / dtc, "unit": "frames/s", "cores": 1, "kind": "port", "seconds": dtc,
                                   "sample": "%d frames of the same clip through scipy.optimize.least_squares exactly as "
                                             "the reference calls it (oracle/optim.py optim_points_port)" % Fc,
                                   "final_cost": pcost, "gpu_final_cost_same_clip": ginfo["cost"]}
    except Exception as ex:  # pragma: no cover
        out = {"error": str(ex)[:300]}
    return out


def cfg4_line(cx, args):
    """BASELINE config 4: cross-view affinity + SVT association + triangulation, 6 macaques x 8 views,
    --cfg4-frames keyframes (default 100k) through crossview.associate_batch in chunks of 10k keyframes
    (host arrays in, host arrays out)."""
    import numpy as np
    from macaque_3d_pose_estimation_b200 import crossview, synth
    out = {}
    try:
        C, A, J = args.cameras, 6, 17
        F = args.cfg4_frames
        rng = np.random.default_rng(404)
        X = synth.make_tracks(F, A, seed=404) * np.array([0.6, 0.6, 0.5])              # (F, A, J, 3)
        proj = cx.cg.project(X.reshape(-1, 3)).reshape(C, F, A, J, 2)
        M = C * A
        kp = np.empty((F, M, J, 3))
        kp[..., :2] = proj.transpose(1, 0, 2, 3, 4).reshape(F, M, J, 2) + rng.normal(0, 0.4, size=(F, M, J, 2))
        sc = rng.uniform(0.3, 1.0, size=(F, M, J))
        sc[rng.random((F, M, J)) < 0.1] = 0.0
        kp[..., 2] = sc
        dim = np.tile(np.arange(C + 1, dtype=np.int32) * A, (F, 1))
        owner = np.tile(np.tile(np.arange(A), C), (F, 1))
        cid = np.where(rng.random((F, M)) < 0.6, owner, -1).astype(np.int32)
        chunk = 10000
        crossview.associate_batch(cx.cg, kp[:2000], dim[:2000], cid[:2000])              # warm-up
        stages = {}
        t0 = time.perf_counter()
        n_person, ok = 0, 0
        for a in range(0, F, chunk):
            res = crossview.associate_batch(cx.cg, kp[a:a + chunk], dim[a:a + chunk], cid[a:a + chunk], timing=stages)
            n_person += len(res["frame"])
            mem = res["members"]
            own = np.where(mem >= 0, owner[0][np.where(mem >= 0, mem, 0)], -1)
            first = own.max(axis=1)
            ok += int(((own == first[:, None]) | (own < 0)).all(axis=1).sum())
        dt = time.perf_counter() - t0
        out = {"workload": "cfg4: cross-view ray affinity + SVT association + LS triangulation, 6 macaques x %d "
                           "views (M = %d detections per keyframe), %d keyframes" % (C, M, F),
               "value": F / dt, "unit": "keyframes/s (one GPU, host arrays in and out)", "seconds": dt,
               "joint_instances_per_s": F * A * J / dt, "persons_found_per_frame": n_person / F,
               "pure_clusters_fraction": ok / max(1, n_person),
               "stages_s": {k: round(v, 4) for k, v in stages.items()}}
        # the same with the detections already on the device (what a GPU pose network upstream would hand over)
        import torch
        nd = min(F, 2 * chunk)
        d_kp = torch.from_numpy(kp[:nd]).to(cx.device)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for a in range(0, nd, chunk):
            crossview.associate_batch(cx.cg, d_kp[a:a + chunk], dim[a:a + chunk], cid[a:a + chunk])
        torch.cuda.synchronize()
        out["device_resident_input"] = {"value": nd / (time.perf_counter() - t0), "unit": "keyframes/s",
                                        "keyframes": nd}
        del d_kp
        if not args.no_cpu:
            from oracle import crossview as ocv
            from oracle import fixtures
            cams = fixtures.cams_from_dicts(synth.make_rig(C, "pinhole", seed=20261018 + 2))
            nf = 6
            t0 = time.perf_counter()
            for f in range(nf):
                ocv.associate_frame(cams, kp[f], dim[f], cid[f], np.arange(M))
            dtc = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": nf / dtc, "unit": "keyframes/s", "cores": 1, "kind": "port",
                                   "sample": "%d keyframes through the loop-faithful restatement of "
                                             "MultiEstimator.predict_data (oracle/crossview.py associate_frame; the "
                                             "reference's own affinity loop is ~100x slower than this numpy form: "
                                             "0.86 s + 1.1 s per keyframe measured, SURVEY.md 8a)" % nf}
    except Exception as ex:  # pragma: no cover
        out = {"error": str(ex)[:300]}
    return out


def config_dict(workload, C, frames_per_gpu, n_gpus, extra=None):
    d = {"workload": NAMES[workload] % C, "cameras": C, "camera_model": "pinhole(5 coeff)",
         "joint_instances_per_gpu": int(frames_per_gpu * ANIMALS * JOINTS), "frames_per_gpu": int(frames_per_gpu),
         "animals": ANIMALS, "joints": JOINTS, "missing_views": 0.1,
         "outliers": 0.2 if workload == "ransac" else 0.0,
         "l2_policy": "inputs larger than L2 (no flush needed)", "parallelism": "frame-sharded dp%d" % n_gpus}
    if extra:
        d.update(extra)
    return d


def run_reference(args):
    """--impl reference: the CPU port on all host cores, same metric and workload; every step is a
    bounded sample of the workload (stated in `config` and `cpu_baseline.sample`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    procs = max(1, min(os.cpu_count() or 1, 64))
    frames = CPU_POOL_FRAMES_PER_PROC[wl] * procs
    pool = CpuPool(procs, wl, args.cameras)
    vals = []
    n = 0
    t_all = time.perf_counter()
    try:
        for i in range(args.warmup + args.steps):
            dicts, p2d = cpu_workload(wl, frames, 20261018 + 11 + i, args.cameras)
            n = p2d.shape[1]
            v = pool.run(dicts, p2d, wl)
            if i >= args.warmup:
                vals.append(v)
            if time.perf_counter() - t_all > 200 and vals:
                break
    finally:
        pool.close()
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "triangulated joint-instances/sec", "value": value,
        "unit": "joint-instances/s", "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
        "ms_per_step": 1e3 * n / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(wl, args.cameras, frames, args.gpus,
                              {"parallelism": "%d host processes" % procs,
                               "l2_policy": "n/a (CPU)",
                               "note": "bounded sample of the workload of the GPU arm: same rig, corruption model, "
                                       "animals x joints; %d frames per step instead of %d per GPU"
                                       % (frames, default_frames(wl))}),
        "cpu_baseline": {"value": value, "unit": "joint-instances/s", "cores": procs, "kind": "port",
                         "sample": "%d joint-instances per step, loop-faithful NumPy/OpenCV port of the reference "
                                   "(oracle/), %d spawn processes (one persistent pool)" % (n, procs)},
        "e2e": {"value": value, "unit": "joint-instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def default_frames(workload):
    return 1000000


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------

class Ctx:
    pass


def bind_to_gpu_numa_node(local):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that the pinned host
    buffers allocated afterwards are local to that GPU's PCIe root.  Returns the node or None
    (platform does not expose one); never fails."""
    try:
        import torch
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


def sync_all(cx):
    import torch
    torch.cuda.synchronize()
    if cx.world > 1:
        cx.dist.barrier()
        torch.cuda.synchronize()


def max_over_ranks(cx, x):
    import torch
    if cx.world == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=cx.device)
    cx.dist.all_reduce(t, op=cx.dist.ReduceOp.MAX)
    return float(t.item())


def timed_steps(cx, step, steps, warmup):
    """W warm-up steps, then exactly K steps between barrier + synchronize, CUDA events on the launch
    stream, max over ranks.  Returns (ms per step, launches in the timed region)."""
    import torch
    for _ in range(max(warmup, 3)):
        step()
    sync_all(cx)
    launches0 = cx.lib.m3d_launch_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record(cx.stream)
    for _ in range(steps):
        step()
    e1.record(cx.stream)
    torch.cuda.synchronize()
    if cx.sampler is not None:
        cx.sampler.mark(t_wall0, time.time())
    ms = e0.elapsed_time(e1)
    launches = cx.lib.m3d_launch_count() - launches0
    ms = max_over_ranks(cx, ms)
    if cx.world > 1:
        cx.dist.barrier()
    return ms / steps, int(launches)


def kernel_breakdown(cx, step, reps=2):
    """Per-kernel device time of `reps` extra steps (CUDA events around every launch, on the launch
    stream, inside the library: m3d_profile_enable / m3d_profile_read)."""
    import torch
    buf = ctypes.create_string_buffer(1 << 14)
    cx.lib.m3d_profile_enable(1)
    cx.lib.m3d_profile_read(buf, len(buf))                     # clear
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    rc = cx.lib.m3d_profile_read(buf, len(buf))
    cx.lib.m3d_profile_enable(0)
    if rc != 0:
        return {}
    d = json.loads(buf.value.decode())
    return {k: {"launches_per_step": v["launches"] / reps, "ms_per_launch": v["ms"] / max(1, v["launches"]),
                "ms_per_step": v["ms"] / reps} for k, v in d.items()}


def roofline_of(cx, workload, C, n_per_step, ms_per_step, kernels):
    peak, peak_kind = measured_peaks()
    bpi = BYTES_PER_INSTANCE[workload](C)
    roof = {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_kind": peak_kind + " copy bandwidth",
            "algorithmic_bytes_per_instance": bpi, "traffic": None}
    # the whole path (all kernels of a step): what the metric sees
    path = bpi * n_per_step / (ms_per_step * 1e-3) / 1e9
    if kernels:
        name = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
        kd = kernels[name]
        n_launch = n_per_step / max(1.0, kd["launches_per_step"])
        ach = bpi * n_launch / (kd["ms_per_launch"] * 1e-3) / 1e9
        roof.update({"kernel": name, "achieved": ach, "frac": ach / peak,
                     "kernel_ms_per_launch": kd["ms_per_launch"], "instances_per_launch": n_launch,
                     "kernel_share_of_step": kd["ms_per_step"] / sum(k["ms_per_step"] for k in kernels.values())})
    else:
        roof.update({"kernel": "whole step", "achieved": path, "frac": path / peak})
    roof["path"] = {"achieved": path, "frac": path / peak,
                    "what": "algorithmic bytes of a step / step time (all kernels, collective included)"}
    roof["kernels"] = kernels
    roof["note"] = ("the path is fp64-ALU bound at reference precision (SURVEY 7): 60 % of HBM would need more fp64 "
                    "instructions per second than the B200 issues")
    try:   # DRAM traffic per launch of the dominant kernel from the committed ncu capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[workload]
        if tj.get("kernel") == roof.get("kernel"):
            roof["traffic"] = tj["dram_bytes_per_instance"] * roof.get("instances_per_launch", n_per_step)
        roof["traffic_source"] = tj["source"]
        tf = ctypes.c_double(0.0)
        if cx.rank == 0 and cx.lib.m3d_probe_fp64_tflops(cx.local, ctypes.byref(tf)) == 0:
            roof["fp64_peak_tflops_measured"] = tf.value
            per = float(tj["fp64_thread_instr_per_instance"])
            ach = per * n_per_step / (ms_per_step * 1e-3) / 1e12
            roof["fp64"] = {"thread_instr_per_instance": per, "achieved": ach, "peak": tf.value / 2.0,
                            "unit": "T fp64 instr/s (per GPU)", "frac": ach / (tf.value / 2.0),
                            "source": tj.get("fp64_source")}
    except Exception:
        pass
    return roof


def copy_ceiling(cx, nbytes=1 << 30, reps=3):
    """Plain pinned cudaMemcpyAsync H2D + D2H at once on two streams, every rank concurrently: the
    ceiling the e2e numbers are read against.  Aggregate GB/s over all ranks."""
    import torch
    try:
        h_in = torch.empty((nbytes,), dtype=torch.uint8, pin_memory=True)
        h_out = torch.empty((nbytes,), dtype=torch.uint8, pin_memory=True)
        d_in = torch.empty((nbytes,), dtype=torch.uint8, device=cx.device)
        d_out = torch.empty((nbytes,), dtype=torch.uint8, device=cx.device)
        s1, s2 = torch.cuda.Stream(cx.device), torch.cuda.Stream(cx.device)
        res = {}
        for mode in ("h2d", "d2h", "duplex"):
            best = None
            for _ in range(reps + 1):
                sync_all(cx)
                t0 = time.perf_counter()
                if mode in ("h2d", "duplex"):
                    with torch.cuda.stream(s1):
                        d_in.copy_(h_in, non_blocking=True)
                if mode in ("d2h", "duplex"):
                    with torch.cuda.stream(s2):
                        h_out.copy_(d_out, non_blocking=True)
                torch.cuda.synchronize()
                dt = max_over_ranks(cx, time.perf_counter() - t0)
                best = dt if best is None else min(best, dt)
            mult = 2 if mode == "duplex" else 1
            res[mode + "_gbs"] = cx.world * mult * nbytes / best / 1e9
        res["what"] = ("pinned cudaMemcpyAsync of 1 GiB per direction and rank, all ranks at once, aggregate over "
                       "ranks (duplex: both directions at once, sum of both)")
        return res
    except Exception as ex:  # pragma: no cover
        return {"error": str(ex)[:200]}


def e2e_run(cx, workload, xy_dev, n_e2e, ref_p3d, steps, f32, stage=False):
    """Same metric through the C-ABI host pipeline: pinned HOST buffers in, HOST buffers out, H2D +
    kernels + D2H inside every timed call, every rank streaming its own shard at once."""
    import torch
    C = xy_dev.shape[0]
    lib, rig = cx.lib, cx.rig
    from macaque_3d_pose_estimation_b200 import _lib
    sp = lambda t: None if t is None else t.data_ptr()
    in_dt = torch.float32 if f32 else torch.float64
    h_xy = torch.empty((C, n_e2e, 2), dtype=in_dt, pin_memory=True)
    h_xy.copy_(xy_dev[:, :n_e2e])
    h_p3d = torch.empty((n_e2e, 3), dtype=torch.float64, pin_memory=True)
    h_err = torch.empty((n_e2e,), dtype=torch.float64, pin_memory=True)
    h_pick = h_xyp = None
    if workload == "ransac":
        h_pick = torch.empty((C, n_e2e), dtype=torch.uint8, pin_memory=True)
        # stage = what the 3D stage asks for (pipeline3d.reconstruct: outputs="picked"): no masked copy of the input
        h_xyp = None if stage else torch.empty((C, n_e2e, 2), dtype=in_dt, pin_memory=True)
    sfx = "_f32" if f32 else ""

    def call():
        if workload == "dlt":
            fn = getattr(lib, "m3d_triangulate_error_host" + sfx)
            rc = fn(rig.handle, h_xy.data_ptr(), n_e2e, 1, h_p3d.data_ptr(), h_err.data_ptr())
        else:
            fn = getattr(lib, "m3d_triangulate_ransac_host" + sfx)
            rc = fn(rig.handle, h_xy.data_ptr(), n_e2e, 1, 2, 0.5, 200.0, h_p3d.data_ptr(), sp(h_pick), sp(h_xyp),
                    h_err.data_ptr(), None, None)
        _lib.check(rc, "e2e step")
    call()
    torch.cuda.synchronize()
    ke = max(1, min(steps, 5))
    sync_all(cx)
    t0 = time.perf_counter()
    for _ in range(ke):
        call()
    dt = max_over_ranks(cx, (time.perf_counter() - t0) / ke)
    if ref_p3d is not None:    # the pipeline's output equals the resident run (f64 input only)
        assert torch.equal(h_p3d.nan_to_num(), ref_p3d[:n_e2e].cpu().nan_to_num()), "e2e result differs"
    isz = 4 if f32 else 8
    d2h = n_e2e * 32 + (n_e2e * C * (1 + (0 if stage else 2 * isz)) if workload == "ransac" else 0)
    return {"value": cx.world * n_e2e / dt, "unit": "joint-instances/s", "h2d_bytes_per_step": n_e2e * C * 2 * isz,
            "d2h_bytes_per_step": d2h, "joint_instances_per_gpu": n_e2e, "ms_per_step": dt * 1e3,
            "input_dtype": "f32 (widened on the device)" if f32 else "f64",
            "api": "m3d_triangulate_%s_host%s (pinned host buffers, 3-slot H2D/kernel/D2H pipeline)"
                   % ("error" if workload == "dlt" else "ransac", sfx), "n_gpus": cx.world}


def measure(cx, args, workload):
    """One workload end to end: device-resident value, kernel breakdown / roofline, e2e, CPU baseline."""
    import torch
    from macaque_3d_pose_estimation_b200 import _lib, sharding
    lib, rig, device, world, rank = cx.lib, cx.rig, cx.device, cx.world, cx.rank
    C = args.cameras
    F = args.frames if args.frames > 0 else default_frames(workload)
    per = ANIMALS * JOINTS
    N = F * per
    seed = 20261018 + (2 if workload == "dlt" else 3) + 17 * rank
    xy = make_device_workload(cx.cg, F, ANIMALS, JOINTS, seed, workload, device)
    p3d = torch.empty((N, 3), dtype=torch.float64, device=device)
    err = torch.empty((N,), dtype=torch.float64, device=device)
    sp = lambda t: None if t is None else t.data_ptr()
    extra = {"mean_valid_views": float((~torch.isnan(xy[:, :, 0])).double().mean().item() * C)}
    out = {}

    if workload == "dlt":
        flag = [1]                                             # M3D_UNDISTORT; 3 = M3D_UNDISTORT_FAST (opt-in)

        def step():
            _lib.check(lib.m3d_triangulate_error(rig.handle, xy.data_ptr(), N, flag[0], p3d.data_ptr(),
                                                 err.data_ptr(), cx.stream.cuda_stream), "bench step")
        collective = None
    else:
        # this rank's tiles of the round-robin deal, stored tile by tile
        rounds = max(1, args.rounds)
        tile_frames = -(-F // rounds)
        plan = sharding.TilePlan(F * world, world, tile_frames, per)
        tiles = []
        off = 0
        for j in range(plan.rounds):
            a, b = plan.tile_span(j * world + rank)
            n = (b - a) * per
            t = {"n": n, "off": off, "xy": xy[:, off:off + n].contiguous(),
                 "picked": torch.empty((C, n), dtype=torch.uint8, device=device),
                 "xyp": torch.empty((C, n, 2), dtype=torch.float64, device=device)}
            tiles.append(t)
            off += n
        assert off == N
        nev = torch.empty((N,), dtype=torch.int32, device=device)
        # delivery of p3d + err to rank 0 (sharding.py): "peer" = every rank's copy engine writes its
        # finished tile into rank 0's window over NVLink (the headline), "direct" = the search kernels
        # store into the window themselves, "gather" = one NCCL gather per round, "none" = no exchange
        pr = sharding.PeerResults(plan, [(3,), ()], torch.float64, device, dst=0, group=None) if world > 1 else None
        rg = sharding.RoundGather(plan, [(3,), ()], torch.float64, device, dst=0, group=None) if world > 1 else None
        mode = ["peer" if world > 1 else "none"]
        if pr is not None and rank == 0:
            for w in pr.out:
                w.fill_(float("nan"))

        def step():
            m = mode[0]
            for j, t in enumerate(tiles):
                o, n = t["off"], t["n"]
                p3d_ptr, err_ptr = p3d[o:o + n].data_ptr(), err[o:o + n].data_ptr()
                direct = m == "direct" or (m == "peer" and rank == 0)
                if direct:
                    p3d_ptr, err_ptr = pr.row_ptrs(j)
                if n:
                    _lib.check(lib.m3d_triangulate_ransac(
                        rig.handle, t["xy"].data_ptr(), n, 1, 2, 0.5, 200.0, p3d_ptr,
                        t["picked"].data_ptr(), t["xyp"].data_ptr(), err_ptr, None,
                        nev[o:o + n].data_ptr(), cx.stream.cuda_stream), "bench step")
                if m == "gather":
                    rg.add(j, [p3d[o:o + n], err[o:o + n]])
                elif m == "peer" and not direct:
                    pr.push(j, [p3d[o:o + n], err[o:o + n]])
            if m == "gather":
                rg.finish()
            elif m in ("peer", "direct"):
                pr.finish()
        collective = None if world == 1 else {
            "what": "p3d + err of every tile delivered to rank 0's frame-ordered arrays INSIDE the timed step: each "
                    "rank's copy engine writes its finished tile into rank 0's CUDA-IPC window over NVLink "
                    "(m3d_peer_push) while the next tile's kernels run; one 4-byte NCCL all-reduce ends the step",
            "rounds_per_step": plan.rounds, "bytes_to_rank0_per_step": int((world - 1) * N * 32)}

    ms_per_step, launches = timed_steps(cx, step, args.steps, args.warmup)
    value = world * N / (ms_per_step * 1e-3)
    if workload == "ransac" and world > 1:
        def window_check(tag):
            # the delivered rows == this rank's own tiles, wherever rank 0 can see both
            torch.cuda.synchronize()
            if rank == 0:       # err is never NaN (cameras.py:675) and the window was NaN-filled: every row arrived
                assert bool(torch.isfinite(pr.out[1]).all()), tag + ": rows missing"
            # every rank: its first tile as delivered, compared on rank 0 with the rank's own copy
            j = 0
            a, b = plan.tile_span(j * world + rank)
            got = [torch.empty(((b - a) * per,), dtype=torch.float64, device=device) for _ in range(world)] \
                if rank == 0 else None
            t = tiles[j]
            mine = err[t["off"]:t["off"] + t["n"]].contiguous()
            dist_ = torch.distributed
            dist_.gather(mine, got, dst=0)
            if rank == 0:
                for r in range(1, world):
                    a, b = plan.tile_span(j * world + r)
                    assert torch.equal(pr.out[1][a * per:b * per], got[r]), tag + ": delivered rows differ"
        window_check("peer")
        collective["ms_per_step_with_exchange"] = ms_per_step
        # the same step with the alternatives, and with no exchange at all: what the delivery costs
        for m, key in (("none", "ms_per_step_without_exchange"), ("gather", "ms_per_step_nccl_gather_per_round"),
                       ("direct", "ms_per_step_kernel_stores_into_window")):
            mode[0] = m
            if m == "direct" and rank == 0:
                pr.out[1].fill_(float("nan"))
            sync_all(cx)
            collective[key], _ = timed_steps(cx, step, max(3, args.steps // 2), 3)
            if m == "direct":
                window_check("direct")
            if m == "gather" and rank == 0:
                t = tiles[0]
                a, b = plan.tile_span(rank)
                assert torch.equal(rg.out[1][a * per:b * per], err[t["off"]:t["off"] + t["n"]]), "gathered rows differ"
        mode[0] = "peer"
    kernels = kernel_breakdown(cx, step) if rank == 0 or world > 1 else {}
    if workload == "dlt":
        # the opt-in north-star-tolerance path (first three undistortion iterations in float32)
        strict3d = p3d.clone()
        flag[0] = 3
        ms_fast, _ = timed_steps(cx, step, max(3, args.steps // 2), 3)
        dmax = float((p3d - strict3d).nan_to_num().abs().max().item())
        peak, _pk = measured_peaks()
        out["fast_undistort"] = {
            "value": world * N / (ms_fast * 1e-3), "unit": "joint-instances/s", "ms_per_step": ms_fast,
            "roofline_frac": BYTES_PER_INSTANCE["dlt"](C) * N / (ms_fast * 1e-3) / 1e9 / peak,
            "max_abs_p3d_diff_vs_strict_mm": dmax,
            "what": "M3D_UNDISTORT_FAST (opt-in): float32 in the first three undistortion iterations; inside "
                    "BASELINE.json's tolerances, not the default"}
        flag[0] = 1
        step()                                                 # p3d back to the strict result for the e2e check
        del strict3d
    out["value"] = value
    out["ms_per_step"] = ms_per_step
    out["gpu_launches"] = launches
    out["roofline"] = roofline_of(cx, workload, C, N, ms_per_step, kernels)
    if workload == "ransac":
        extra["mean_subsets_per_point"] = float(nev.double().mean().item())
        extra["selected_fraction"] = float((~torch.isnan(p3d[:, 0])).double().mean().item())
    if collective:
        out["collective"] = collective

    # ---- e2e through host buffers: the same per-GPU size at every N -------------------------
    n_e2e = min(N, args.e2e_points)
    if args.no_e2e:
        out["e2e"] = {"value": None, "unit": "joint-instances/s", "error": "skipped (--no-e2e)"}
    else:
        try:
            xy_flat = xy if workload == "dlt" else torch.cat([t["xy"] for t in tiles], dim=1)
            out["e2e"] = e2e_run(cx, workload, xy_flat, n_e2e, p3d, args.steps, f32=False)
            out["e2e"]["host_numa_node"] = cx.numa
            out["e2e_f32"] = e2e_run(cx, workload, xy_flat, n_e2e, None, args.steps, f32=True)
            if workload == "ransac":
                # the call the 3D stage makes (step4:296-300 needs points_3d, picked_vals, errors)
                out["e2e_stage"] = e2e_run(cx, workload, xy_flat, n_e2e, p3d, args.steps, f32=False, stage=True)
                out["e2e_stage"]["what"] = ("m3d_triangulate_ransac_host without the points_2d output (NULL): what "
                                            "pipeline3d.reconstruct / CameraGroup.triangulate_ransac(outputs='picked') "
                                            "move; points_2d is the input masked by picked_vals")
            del xy_flat
        except Exception as ex:  # pragma: no cover
            out["e2e"] = {"value": None, "unit": "joint-instances/s", "error": str(ex)[:300]}
    out["config"] = config_dict(workload, C, F, world)
    out["extra"] = extra
    if workload == "ransac" and pr is not None:
        pr.close()
    del xy, p3d, err
    torch.cuda.empty_cache()
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.build_library()
    from macaque_3d_pose_estimation_b200 import _lib, synth
    from macaque_3d_pose_estimation_b200.cameras import CameraGroup

    cx = Ctx()
    cx.world = int(os.environ.get("WORLD_SIZE", "1"))
    cx.rank = int(os.environ.get("RANK", "0"))
    cx.local = int(os.environ.get("LOCAL_RANK", "0"))
    cx.dist = dist
    torch.cuda.set_device(cx.local)
    cx.device = torch.device("cuda", cx.local)
    cx.numa = bind_to_gpu_numa_node(cx.local)     # before any pinned allocation
    if cx.world > 1:
        dist.init_process_group("nccl", device_id=cx.device)
        warm = torch.zeros((cx.world * 8,), dtype=torch.float64, device=cx.device)
        dist.all_reduce(warm)                     # communicator set-up outside every timing
    cx.lib = _lib.require_gpu()
    cx.cg = CameraGroup.from_dicts(synth.make_rig(args.cameras, "pinhole", seed=args.rig_seed))
    cx.cg.device = cx.local
    cx.rig = cx.cg._rig(cx.local)
    cx.stream = torch.cuda.current_stream(cx.device)
    cx.sampler = ClockSampler(cx.local) if cx.rank == 0 else None
    if cx.sampler:
        cx.sampler.start()
        time.sleep(0.25)

    other = "dlt" if args.workload == "ransac" else "ransac"
    head = measure(cx, args, args.workload)
    second = None if args.only else measure(cx, args, other)
    ceiling = None if args.no_e2e else copy_ceiling(cx)
    clocks = cx.sampler.stop() if cx.sampler else None

    if cx.rank != 0:
        if cx.world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- CPU baselines (rank 0, N = 1 only): loop-faithful port, one core + all cores ---------
    if cx.world == 1 and not args.no_cpu:
        head["cpu_baseline"] = cpu_baseline(args.workload, args.cameras, True)
        if second is not None:
            second["cpu_baseline"] = cpu_baseline(other, args.cameras, True)

    for obj in (head, second):
        if obj is not None and ceiling is not None and obj["e2e"].get("value"):
            obj["e2e"]["pcie_ceiling"] = ceiling
            bytes_per_inst = (obj["e2e"]["h2d_bytes_per_step"] + obj["e2e"]["d2h_bytes_per_step"]) / \
                obj["e2e"]["joint_instances_per_gpu"]
            obj["e2e"]["moved_gbs"] = obj["e2e"]["value"] * bytes_per_inst / 1e9
            if ceiling.get("duplex_gbs"):
                obj["e2e"]["frac_of_duplex_ceiling"] = obj["e2e"]["moved_gbs"] / ceiling["duplex_gbs"]

    line = {
        "metric": "triangulated joint-instances/sec", "value": head["value"], "unit": "joint-instances/s",
        "n_gpus": cx.world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": head["config"], "roofline": head["roofline"], "cpu_baseline": head.get("cpu_baseline"),
        "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "clocks": clocks,
    }
    for k in ("e2e_f32", "e2e_stage", "collective", "fast_undistort"):
        if k in head:
            line[k] = head[k]
    line.update(head["extra"])
    if cx.world == 1 and not args.only:
        line["optim_points"] = optim_line(cx, args)
        line["cfg4"] = cfg4_line(cx, args)
    if second is not None:
        key = "cfg2" if other == "dlt" else "cfg3"
        sec = {k: second[k] for k in ("value", "ms_per_step", "gpu_launches", "config", "roofline", "e2e") if k in second}
        sec["unit"] = "joint-instances/s"
        for k in ("e2e_f32", "e2e_stage", "collective", "cpu_baseline", "fast_undistort"):
            if k in second:
                sec[k] = second[k]
        sec.update(second["extra"])
        line[key] = sec
    print(json.dumps(line))
    if cx.world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", choices=["dlt", "ransac"], default="ransac")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU (default 1e6)")
    ap.add_argument("--rounds", type=int, default=8, help="tile rounds per step of the sharded RANSAC")
    ap.add_argument("--e2e-points", type=int, default=20000000,
                    help="joint-instances per GPU of the e2e run (the same at every N)")
    ap.add_argument("--cameras", type=int, default=8, help="cameras of the synthetic ring rig (BASELINE: 8; config 5: 16)")
    ap.add_argument("--cfg4-frames", type=int, default=100000, help="keyframes of the config-4 association line")
    ap.add_argument("--rig-seed", type=int, default=20261018 + 2, help="seed of the synthetic ring rig")
    ap.add_argument("--only", action="store_true", help="measure the headline workload only")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
