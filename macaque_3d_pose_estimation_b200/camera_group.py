"""``CameraGroup`` of the drop-in API (reference aniposelib/cameras.py:558-783, 1883-1889,
1966-2013) on the GPU kernels of libm3d.so."""
import ctypes

import numpy as np

from . import _lib
from ._device import (_RigHandle, _default_device, _is_torch, _np_ptr, _ptr, _ret, _stream, _to_dev, torch)
from .camera_models import Camera, FisheyeCamera, OmnidirCamera

_OUT_OF_SCOPE = ("%s is outside the accelerated hot path (SURVEY.md §8f); "
                 "reference: aniposelib/cameras.py:%s")


class CameraGroup:
    """Reference cameras.py:558-783 on the GPU."""

    def __init__(self, cameras, metadata={}, device=None):
        self.cameras = cameras
        self.metadata = metadata
        self.device = device
        self._rig_cache = None
        # opt-in (not part of the reference API): run the first three of OpenCV's five
        # undistortion iterations in float32 inside triangulate / triangulate_with_error
        # (M3D_UNDISTORT_FAST of include/m3d.h: results within BASELINE.json's tolerances, not
        # bit-identical to the float64 path).  The subset RANSAC ignores it.
        self.fast_undistort = False

    # -- bookkeeping --
    def _dev(self):
        return self.device if self.device is not None else _default_device()

    def _rig(self, device=None):
        device = self._dev() if device is None else device
        fp = tuple(cam._fingerprint() for cam in self.cameras)
        if self._rig_cache is None or self._rig_cache[0] != fp:
            self._rig_cache = (fp, {})
        rigs = self._rig_cache[1]
        if device not in rigs:
            rigs[device] = _RigHandle(self.cameras, device)
        return rigs[device]

    # numpy inputs of at least this many joint-instances are spread over all visible GPUs (one
    # process, one thread per GPU, each streaming its own span of the caller's arrays); set
    # ``device`` on the group to pin it to one GPU
    MULTI_GPU_MIN_POINTS = 4000000

    def _host_devices(self, n):
        if self.device is not None or n < self.MULTI_GPU_MIN_POINTS or torch is None:
            return None
        k = torch.cuda.device_count()
        return list(range(k)) if k > 1 else None

    def _host_spans(self, devices, n, call):
        """Run ``call(rig, first, count)`` for one contiguous span per device, concurrently (ctypes
        releases the GIL for the duration of the C call)."""
        import threading
        k = len(devices)
        bounds = [n * i // k for i in range(k + 1)]
        errs = []

        def work(dev, a, b):
            try:
                call(self._rig(dev), a, b - a)
            except Exception as e:  # noqa: BLE001 - re-raised below
                errs.append(e)
        rigs = [self._rig(d) for d in devices]            # create the handles in this thread
        ths = [threading.Thread(target=work, args=(d, bounds[i], bounds[i + 1])) for i, d in enumerate(devices)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        del rigs
        if errs:
            raise errs[0]

    def subset_cameras(self, indices):
        """New group over deep copies of the selected cameras (the reference copies as well)."""
        return CameraGroup([self.cameras[i].copy() for i in indices], self.metadata, self.device)

    def subset_cameras_names(self, names):
        """Select cameras by name; unknown names raise IndexError with the reference's message."""
        known = self.get_names()
        position = {nm: i for i, nm in enumerate(known)}
        missing = [nm for nm in names if nm not in position]
        if missing:
            raise IndexError("name {} not part of camera names: {}".format(missing[0], known))
        return self.subset_cameras([position[nm] for nm in names])

    def get_names(self):
        return [c.get_name() for c in self.cameras]

    def set_names(self, names):
        for c, nm in zip(self.cameras, names):
            c.set_name(nm)

    def get_rotations(self):
        return np.array([c.get_rotation() for c in self.cameras])

    def get_translations(self):
        return np.array([c.get_translation() for c in self.cameras])

    def set_rotations(self, rvecs):
        for c, r in zip(self.cameras, rvecs):
            c.set_rotation(r)

    def set_translations(self, tvecs):
        for c, t in zip(self.cameras, tvecs):
            c.set_translation(t)

    def resize_cameras(self, scale):
        for c in self.cameras:
            c.resize_camera(scale)

    def get_extrinsics_mats(self):
        """make_M for every camera -> (C,4,4) (cameras.py:621)."""
        rig = self._rig()
        M = np.empty((max(len(self.cameras), 1), 4, 4))
        _lib.check(rig._lib.m3d_rig_extrinsics(rig.handle, _np_ptr(M)), "m3d_rig_extrinsics")
        return M[:len(self.cameras)]

    # -- helpers --
    def _device_of(self, *arrays):
        for a in arrays:
            if _is_torch(a) and a.device.type == "cuda":
                return a.device.index, True
        return self._dev(), any(_is_torch(a) for a in arrays)

    def _assert_cams(self, points):
        assert points.shape[0] == len(self.cameras), \
            "Invalid points shape, first dim should be equal to" \
            " number of cameras ({}), but shape is {}".format(
                len(self.cameras), tuple(points.shape) if _is_torch(points) else points.shape)

    # -- hot path --
    def undistort_points(self, points):
        """Batched form of the per-camera loop at cameras.py:608-614: (C,N,2) -> (C,N,2)."""
        self._assert_cams(points)
        device, like_torch = self._device_of(points)
        rig = self._rig(device)
        src = _to_dev(points, device).reshape(len(self.cameras), -1, 2)
        out = torch.empty_like(src)
        _lib.check(rig._lib.m3d_undistort(rig.handle, _ptr(src), src.shape[1], _ptr(out),
                                          _stream(device)), "m3d_undistort")
        return _ret(out.reshape(tuple(points.shape)), like_torch)

    def project(self, points):
        """Given an Nx3 array of points, this returns an CxNx2 array of 2D points
        (cameras.py:580-591)."""
        device, like_torch = self._device_of(points)
        rig = self._rig(device)
        src = _to_dev(points, device).reshape(-1, 3)
        n = src.shape[0]
        out = torch.empty((len(self.cameras), n, 2), dtype=torch.float64, device=src.device)
        _lib.check(rig._lib.m3d_project(rig.handle, _ptr(src), n, _ptr(out), _stream(device)),
                   "m3d_project")
        return _ret(out, like_torch)

    def triangulate(self, points, undistort=True, progress=False):
        """Given an CxNx2 array, this returns an Nx3 array of points (cameras.py:593-637).
        ``progress`` is accepted for compatibility (the reference only drives tqdm with it)."""
        self._assert_cams(points)
        one_point = False
        if len(points.shape) == 2:
            points = points.reshape(-1, 1, 2)
            one_point = True
        out, _ = self._triangulate_error(points, undistort, with_err=False)
        if one_point:
            out = out[0]
        return out

    def triangulate_with_error(self, points, undistort=True):
        """Fused triangulate + reprojection_error(mean=True) in one pass over the input
        (the plain branch of the 3D stage, step4_aniposefiltering.py:306-309).
        Returns (p3d (N,3), err (N,))."""
        self._assert_cams(points)
        return self._triangulate_error(points, undistort, with_err=True)

    def _undistort_flag(self, undistort):
        if not undistort:
            return 0
        return 3 if self.fast_undistort else 1          # M3D_UNDISTORT_FAST / M3D_UNDISTORT

    def _triangulate_error(self, points, undistort, with_err):
        C = len(self.cameras)
        n = points.shape[1]
        device, like_torch = self._device_of(points)
        rig = self._rig(device)
        if not like_torch:
            # host buffers: chunked H2D -> kernel -> D2H pipeline inside the library
            src = np.ascontiguousarray(points, dtype=np.float64)
            p3d = np.empty((n, 3))
            err = np.empty(n) if with_err else None
            devs = self._host_devices(n)
            if devs:
                flag = self._undistort_flag(undistort)
                self._host_spans(devs, n, lambda r, a, cnt: _lib.check(r._lib.m3d_triangulate_error_host_span(
                    r.handle, _np_ptr(src), n, a, cnt, flag, _np_ptr(p3d), _np_ptr(err)),
                    "m3d_triangulate_error_host_span"))
                return p3d, err
            _lib.check(rig._lib.m3d_triangulate_error_host(rig.handle, _np_ptr(src), n, self._undistort_flag(undistort),
                                                           _np_ptr(p3d), _np_ptr(err)),
                       "m3d_triangulate_error_host")
            return p3d, err
        src = _to_dev(points, device)
        p3d = torch.empty((n, 3), dtype=torch.float64, device=src.device)
        err = torch.empty((n,), dtype=torch.float64, device=src.device) if with_err else None
        _lib.check(rig._lib.m3d_triangulate_error(rig.handle, _ptr(src), n, self._undistort_flag(undistort),
                                                  _ptr(p3d), _ptr(err), _stream(device)),
                   "m3d_triangulate_error")
        return p3d, err

    def triangulate_possible(self, points, undistort=True, min_cams=2, progress=False,
                             threshold=0.5, return_stats=False):
        """Given an CxNxPx2 array, triangulate all combinations of one candidate (or none) per
        camera and pick the one with the best reprojection error (cameras.py:639-724).  P == 1 is
        the camera-subset search of triangulate_ransac (k_ransac_search8/16); P > 1 runs the
        mixed-radix product on k_possible (cameras * P <= 32).  Returns (points_3d (N,3),
        picked_vals (C,N,P) bool, points_2d (C,N,2), errors (N,)); with return_stats also the
        index of the accepted combination in itertools.product order and the number evaluated."""
        self._assert_cams(points)
        n_cams, n_points, n_possible, _ = points.shape
        if n_possible == 1:
            pts = points.reshape(n_cams, n_points, 2)
            return self._ransac(pts, undistort, min_cams, threshold, 200.0, return_stats)
        device, like_torch = self._device_of(points)
        rig = self._rig(device)
        x = _to_dev(points, device)
        p3d = torch.empty((n_points, 3), dtype=torch.float64, device=x.device)
        picked = torch.empty((n_cams, n_points, n_possible), dtype=torch.uint8, device=x.device)
        xyp = torch.empty((n_cams, n_points, 2), dtype=torch.float64, device=x.device)
        err = torch.empty((n_points,), dtype=torch.float64, device=x.device)
        idx = torch.empty((n_points,), dtype=torch.int32, device=x.device)
        nev = torch.empty((n_points,), dtype=torch.int32, device=x.device)
        with torch.cuda.device(device):
            _lib.check(rig._lib.m3d_triangulate_possible(
                rig.handle, _ptr(x), n_points, int(n_possible), int(bool(undistort)), int(min_cams),
                float(threshold), 200.0, _ptr(p3d), _ptr(picked), _ptr(xyp), _ptr(err), _ptr(idx), _ptr(nev),
                _stream(device)), "m3d_triangulate_possible")
        res = (p3d, picked.bool(), xyp, err) + ((idx, nev) if return_stats else ())
        return res if like_torch else tuple(t.cpu().numpy() for t in res)

    def triangulate_ransac(self, points, undistort=True, min_cams=2, progress=False,
                           return_stats=False, outputs="all"):
        """Given an CxNx2 array, this returns (points_3d (N,3), picked_vals (C,N,1) bool,
        points_2d (C,N,2), errors (N,)) (cameras.py:726-743).  With return_stats also
        (subset_index (N,) int32, n_evaluated (N,) int32).

        ``outputs`` (not in the reference): "all" = the reference's four arrays; "picked" = points_2d is
        not produced (None in its place) — it is the input masked by ``picked_vals``, 2/3 of the bytes
        a host caller gets back, and step 4 only looks at its NaN pattern (step4_aniposefiltering.py:
        299-300); "points_3d" = only points_3d and errors (step4:237)."""
        self._assert_cams(points)
        assert outputs in ("all", "picked", "points_3d"), "outputs must be 'all', 'picked' or 'points_3d'"
        return self._ransac(points, undistort, min_cams, 0.5, 200.0, return_stats, outputs)

    def _ransac(self, points, undistort, min_cams, threshold, init_best, return_stats, outputs="all"):
        C = len(self.cameras)
        n = points.shape[1]
        device, like_torch = self._device_of(points)
        rig = self._rig(device)
        want_pick = outputs in ("all", "picked")
        want_xyp = outputs == "all"
        if not like_torch:
            src = np.ascontiguousarray(points, dtype=np.float64)
            p3d = np.empty((n, 3))
            picked = np.empty((C, n, 1), dtype=np.uint8) if want_pick else None
            xyp = np.empty((C, n, 2)) if want_xyp else None
            err = np.empty(n)
            sub = np.empty(n, dtype=np.int32) if return_stats else None
            nev = np.empty(n, dtype=np.int32) if return_stats else None
            opt = lambda a: None if a is None else _np_ptr(a)
            devs = self._host_devices(n)
            if devs:
                self._host_spans(devs, n, lambda r, a, cnt: _lib.check(r._lib.m3d_triangulate_ransac_host_span(
                    r.handle, _np_ptr(src), n, a, cnt, int(bool(undistort)), int(min_cams), float(threshold),
                    float(init_best), _np_ptr(p3d), opt(picked), opt(xyp), _np_ptr(err), opt(sub),
                    opt(nev)), "m3d_triangulate_ransac_host_span"))
            else:
                _lib.check(rig._lib.m3d_triangulate_ransac_host(
                    rig.handle, _np_ptr(src), n, int(bool(undistort)), int(min_cams), float(threshold),
                    float(init_best), _np_ptr(p3d), opt(picked), opt(xyp), _np_ptr(err),
                    opt(sub), opt(nev)), "m3d_triangulate_ransac_host")
            res = (p3d, picked.view(np.bool_) if want_pick else None, xyp, err)
            return res + (sub, nev) if return_stats else res
        src = _to_dev(points, device)
        dev = src.device
        p3d = torch.empty((n, 3), dtype=torch.float64, device=dev)
        picked = torch.empty((C, n, 1), dtype=torch.uint8, device=dev) if want_pick else None
        xyp = torch.empty((C, n, 2), dtype=torch.float64, device=dev) if want_xyp else None
        err = torch.empty((n,), dtype=torch.float64, device=dev)
        sub = torch.empty((n,), dtype=torch.int32, device=dev) if return_stats else None
        nev = torch.empty((n,), dtype=torch.int32, device=dev) if return_stats else None
        _lib.check(rig._lib.m3d_triangulate_ransac(
            rig.handle, _ptr(src), n, int(bool(undistort)), int(min_cams), float(threshold),
            float(init_best), _ptr(p3d), _ptr(picked), _ptr(xyp), _ptr(err), _ptr(sub), _ptr(nev),
            _stream(device)), "m3d_triangulate_ransac")
        res = (p3d, picked.view(torch.bool) if want_pick else None, xyp, err)
        return res + (sub, nev) if return_stats else res

    def reprojection_error(self, p3ds, p2ds, mean=False):
        """Given an Nx3 array of 3D points and an CxNx2 array of 2D points, this returns an
        CxNx2 array of errors; mean=True averages the residual norms over cameras and
        returns an array of length N (cameras.py:746-783)."""
        one_point = False
        if len(p3ds.shape) == 1 and len(p2ds.shape) == 2:
            p3ds = p3ds.reshape(1, 3)
            p2ds = p2ds.reshape(-1, 1, 2)
            one_point = True
        n_cams, n_points, _ = p2ds.shape
        assert tuple(p3ds.shape) == (n_points, 3), \
            "shapes of 2D and 3D points are not consistent: " \
            "2D={}, 3D={}".format(tuple(p2ds.shape), tuple(p3ds.shape))
        assert n_cams == len(self.cameras), \
            "Invalid points shape, first dim should be equal to" \
            " number of cameras ({}), but shape is {}".format(len(self.cameras), tuple(p2ds.shape))
        device, like_torch = self._device_of(p3ds, p2ds)
        rig = self._rig(device)
        X = _to_dev(p3ds, device)
        P = _to_dev(p2ds, device)
        if mean:
            out = torch.empty((n_points,), dtype=torch.float64, device=X.device)
        else:
            out = torch.empty((n_cams, n_points, 2), dtype=torch.float64, device=X.device)
        _lib.check(rig._lib.m3d_reproj_error(rig.handle, _ptr(X), _ptr(P), n_points, int(bool(mean)),
                                             _ptr(out), _stream(device)), "m3d_reproj_error")
        errors = _ret(out, like_torch)
        if one_point:
            if mean:
                errors = float(errors[0])
            else:
                errors = errors.reshape(-1, 2)
        return errors

    def average_error(self, p2ds, median=False):
        """cameras.py:1883-1889."""
        p3ds, errors = self.triangulate_with_error(p2ds)
        if _is_torch(errors):
            return torch.median(errors) if median else torch.mean(errors)
        return np.median(errors) if median else np.mean(errors)

    # -- (de)serialisation: same dict / TOML layout as the reference ------------------------------
    def get_dicts(self):
        return [c.get_dict() for c in self.cameras]

    @staticmethod
    def from_dicts(arr):
        """'fisheye' -> FisheyeCamera, 'omnidir' -> OmnidirCamera, else pinhole Camera."""
        def build(d):
            if d.get('fisheye'):
                return FisheyeCamera.from_dict(d)
            if d.get('omnidir'):
                return OmnidirCamera.from_dict(d)
            return Camera.from_dict(d)
        return CameraGroup([build(d) for d in arr])

    @staticmethod
    def from_names(names, fisheye=False):
        kind = FisheyeCamera if fisheye else Camera
        return CameraGroup([kind(name=nm) for nm in names])

    def load_dicts(self, arr):
        for c, d in zip(self.cameras, arr):
            c.load_dict(d)

    def dump(self, fname):
        """calibration.toml: one table cam_<i> per camera plus 'metadata'."""
        import toml
        doc = {'cam_{}'.format(i): d for i, d in enumerate(self.get_dicts())}
        doc['metadata'] = self.metadata
        with open(fname, 'w') as f:
            toml.dump(doc, f, encoder=toml.TomlNumpyEncoder())

    @staticmethod
    def load(fname):
        import toml
        doc = toml.load(fname)
        cgroup = CameraGroup.from_dicts([doc[k] for k in sorted(doc) if k != 'metadata'])
        if 'metadata' in doc:
            cgroup.metadata = doc['metadata']
        return cgroup

    # -- temporal / skeletal refinement --
    def optim_points(self, points, p3ds, constraints=[], constraints_weak=[], scale_smooth=4, scale_length=2,
                     scale_length_weak=0.5, reproj_error_threshold=15, reproj_loss='soft_l1', n_deriv_smooth=1,
                     scores=None, verbose=False, **solver):
        """Take in an array of 2D points of shape CxNxJx2, an array of 3D points of shape NxJx3 and
        constraints of shape Kx2; returns the optimised NxJx3 points and the limb lengths
        (cameras.py:1116-1190) — GPU Levenberg-Marquardt, see optim.py."""
        from . import optim
        return optim.optim_points(self, points, p3ds, constraints, constraints_weak, scale_smooth, scale_length,
                                  scale_length_weak, reproj_error_threshold, reproj_loss, n_deriv_smooth, scores,
                                  verbose, None, **solver)

    def optim_points_jointlenfix(self, points, p3ds, joint_len, constraints=[], constraints_weak=[], scale_smooth=4,
                                 scale_length=2, scale_length_weak=0.5, reproj_error_threshold=15,
                                 reproj_loss='soft_l1', n_deriv_smooth=1, scores=None, verbose=False, **solver):
        """optim_points with the limb lengths held at ``joint_len`` (cameras.py:1192-1270)."""
        from . import optim
        return optim.optim_points(self, points, p3ds, constraints, constraints_weak, scale_smooth, scale_length,
                                  scale_length_weak, reproj_error_threshold, reproj_loss, n_deriv_smooth, scores,
                                  verbose, np.asarray(joint_len, dtype=np.float64), **solver)

    # -- out of scope --

    def bundle_adjust(self, *a, **k):
        raise NotImplementedError(_OUT_OF_SCOPE % ("bundle_adjust", "860-946"))

    def bundle_adjust_iter(self, *a, **k):
        raise NotImplementedError(_OUT_OF_SCOPE % ("bundle_adjust_iter", "786-858"))

    def calibrate_rows(self, *a, **k):
        raise NotImplementedError(_OUT_OF_SCOPE % ("calibrate_rows", "1891-1940"))

    def calibrate_videos(self, *a, **k):
        raise NotImplementedError(_OUT_OF_SCOPE % ("calibrate_videos", "1951-1964"))
