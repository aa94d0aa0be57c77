"""Drop-in camera classes for the reference's ``aniposelib.cameras`` hot path, running on
hand-written sm_100a kernels through libm3d.so (include/m3d.h).

Mirrored reference API (/root/reference/src/third_party/aniposelib/cameras.py):
  Camera :173-337, FisheyeCamera :339-426, OmnidirCamera :429-555,
  CameraGroup :558-783 (subset_cameras, project, triangulate, triangulate_possible,
  triangulate_ransac, reprojection_error), average_error :1883, get_dicts/from_dicts/
  from_names/load_dicts/dump/load :1966-2013.

Same names, argument meaning, shapes, dtypes (float64 / bool), NaN conventions and
assertion messages.  numpy in -> numpy out (host buffers are streamed through the GPU by
the library's chunked H2D/kernel/D2H pipeline); torch CUDA tensors in -> torch CUDA
tensors out with no host round trip.  There is no CPU implementation behind these
classes: without libm3d.so and a GPU every compute method raises.

Out of scope: bundle adjustment, optim_points*, calibration from
videos/boards — those methods raise NotImplementedError naming the reference line.
"""
import ctypes

import numpy as np

from . import _lib

try:  # torch is the device-memory / stream plumbing (never the compute path)
    import torch
except Exception:  # pragma: no cover
    torch = None


# ----------------------------------------------------------------------------------
# device plumbing
# ----------------------------------------------------------------------------------

def _default_device():
    if torch is None or not torch.cuda.is_available():
        _lib.require_gpu()
        raise RuntimeError("torch.cuda is not available; the B200 path has no CPU fallback")
    return torch.cuda.current_device()


def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


def _to_dev(x, device):
    """float64 contiguous CUDA tensor view/copy of a numpy array or tensor."""
    if _is_torch(x):
        t = x
        if t.device.type != "cuda":
            t = t.to("cuda:%d" % device)
        return t.to(torch.float64).contiguous()
    a = np.ascontiguousarray(x, dtype=np.float64)
    return torch.from_numpy(a).to("cuda:%d" % device)


def _ret(t, like_torch):
    return t if like_torch else t.cpu().numpy()


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else ctypes.c_void_p(0)


def _np_ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None and a.size > 0 else ctypes.c_void_p(0)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _RigHandle:
    """Owns one m3d_rig (immutable camera group on one GPU)."""

    def __init__(self, cameras, device):
        lib = _lib.require_gpu()
        n = len(cameras)
        if n > _lib.MAX_CAMS:
            raise ValueError("at most %d cameras are supported, got %d" % (_lib.MAX_CAMS, n))
        arr = (_lib.M3DCam * max(n, 1))()
        for i, cam in enumerate(cameras):
            cam._fill_struct(arr[i])
        h = ctypes.c_void_p()
        _lib.check(lib.m3d_rig_create(arr, n, int(device), ctypes.byref(h)), "m3d_rig_create")
        self._lib = lib
        self.handle = h
        self.device = int(device)
        self.n_cams = n

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self._lib.m3d_rig_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


# ----------------------------------------------------------------------------------
# cameras
# ----------------------------------------------------------------------------------

class Camera:
    """Pinhole camera with OpenCV distortion (reference cameras.py:173-337)."""

    _MODEL = _lib.MODEL_PINHOLE

    def __init__(self, matrix=np.eye(3), dist=np.zeros(5), size=None, rvec=np.zeros(3),
                 tvec=np.zeros(3), name=None, extra_dist=False):
        self.set_camera_matrix(matrix)
        self.set_distortions(dist)
        self.set_size(size)
        self.set_rotation(rvec)
        self.set_translation(tvec)
        self.set_name(name)
        self.extra_dist = extra_dist
        self._rig_cache = None

    # -- (de)serialisation (cameras.py:191-212) --
    def get_dict(self):
        return {
            'name': self.get_name(),
            'size': list(self.get_size()),
            'matrix': self.get_camera_matrix().tolist(),
            'distortions': self.get_distortions().tolist(),
            'rotation': self.get_rotation().tolist(),
            'translation': self.get_translation().tolist(),
        }

    def load_dict(self, d):
        self.set_camera_matrix(d['matrix'])
        self.set_rotation(d['rotation'])
        self.set_translation(d['translation'])
        self.set_distortions(d['distortions'])
        self.set_name(d['name'])
        self.set_size(d['size'])

    @classmethod
    def from_dict(cls, d):
        cam = cls()
        cam.load_dict(d)
        return cam

    # -- accessors (cameras.py:214-267) --
    def get_camera_matrix(self):
        return self.matrix

    def get_distortions(self):
        return self.dist

    def set_camera_matrix(self, matrix):
        self.matrix = np.array(matrix, dtype='float64')

    def set_focal_length(self, fx, fy=None):
        if fy is None:
            fy = fx
        self.matrix[0, 0] = fx
        self.matrix[1, 1] = fy

    def get_focal_length(self, both=False):
        fx = self.matrix[0, 0]
        fy = self.matrix[1, 1]
        if both:
            return (fx, fy)
        return (fx + fy) / 2.0

    def set_distortions(self, dist):
        self.dist = np.array(dist, dtype='float64').ravel()

    def set_rotation(self, rvec):
        self.rvec = np.array(rvec, dtype='float64').ravel()

    def get_rotation(self):
        return self.rvec

    def set_translation(self, tvec):
        self.tvec = np.array(tvec, dtype='float64').ravel()

    def get_translation(self):
        return self.tvec

    def get_name(self):
        return self.name

    def set_name(self, name):
        self.name = str(name)

    def set_size(self, size):
        """set size as (width, height)"""
        self.size = size

    def get_size(self):
        """get size as (width, height)"""
        return self.size

    def resize_camera(self, scale):
        """resize the camera by scale factor, updating intrinsics to match (cameras.py:269)"""
        size = self.get_size()
        new_size = size[0] * scale, size[1] * scale
        new_matrix = self.get_camera_matrix() * scale
        new_matrix[2, 2] = 1
        self.set_size(new_size)
        self.set_camera_matrix(new_matrix)

    def get_params(self):
        params = np.zeros(8 + self.extra_dist, dtype='float64')
        params[0:3] = self.get_rotation()
        params[3:6] = self.get_translation()
        params[6] = self.get_focal_length()
        dist = self.get_distortions()
        params[7] = dist[0]
        if self.extra_dist:
            params[8] = dist[1]
        return params

    def set_params(self, params):
        self.set_rotation(params[0:3])
        self.set_translation(params[3:6])
        self.set_focal_length(params[6])
        dist = np.zeros(self._N_PARAM_DIST, dtype='float64')
        dist[0] = params[7]
        if self.extra_dist:
            dist[1] = params[8]
        self.set_distortions(dist)

    _N_PARAM_DIST = 5

    def copy(self):
        return type(self)(matrix=self.get_camera_matrix().copy(), dist=self.get_distortions().copy(),
                          size=self.get_size(), rvec=self.get_rotation().copy(),
                          tvec=self.get_translation().copy(), name=self.get_name(),
                          extra_dist=self.extra_dist)

    # -- C-ABI record --
    def _intrinsics(self):
        """(K, distortion vector, xi) handed to the kernels."""
        return self.matrix, self.dist, 0.0

    def _fill_struct(self, s):
        K, dist, xi = self._intrinsics()
        K = np.asarray(K, dtype=np.float64).reshape(3, 3)
        dist = np.asarray(dist, dtype=np.float64).ravel()
        if dist.size > 14:
            raise ValueError("distortion vector longer than 14 entries")
        s.model = self._MODEL
        s.n_dist = int(dist.size)
        for i in range(9):
            s.K[i] = float(K.flat[i])
        for i in range(14):
            s.dist[i] = float(dist[i]) if i < dist.size else 0.0
        for i in range(3):
            s.rvec[i] = float(self.rvec[i])
            s.tvec[i] = float(self.tvec[i])
        s.xi = float(xi)

    def _fingerprint(self):
        K, dist, xi = self._intrinsics()
        return (self._MODEL, np.asarray(K, dtype=np.float64).tobytes(),
                np.asarray(dist, dtype=np.float64).tobytes(), float(xi),
                self.rvec.tobytes(), self.tvec.tobytes())

    def _rig(self, device):
        key = (device, self._fingerprint())
        cache = getattr(self, "_rig_cache", None)
        if cache is None or cache[0] != key:
            self._rig_cache = (key, _RigHandle([self], device))
        return self._rig_cache[1]

    def get_extrinsics_mat(self):
        """4x4 [R|t] with R = cv2.Rodrigues(rvec) (cameras.py:252, utils.py:9-15)."""
        rig = self._rig(_default_device())
        M = np.empty((1, 4, 4))
        _lib.check(rig._lib.m3d_rig_extrinsics(rig.handle, _np_ptr(M)), "m3d_rig_extrinsics")
        return M[0]

    # -- point maps (cameras.py:301-327) --
    def _map2(self, fn_name, points):
        like_torch = _is_torch(points)
        device = points.device.index if like_torch and points.device.type == "cuda" else _default_device()
        shape = tuple(points.shape)
        rig = self._rig(device)
        src = _to_dev(points, device).reshape(-1, 2)
        out = torch.empty_like(src)
        fn = getattr(rig._lib, fn_name)
        _lib.check(fn(rig.handle, 0, _ptr(src), src.shape[0], _ptr(out), _stream(device)), fn_name)
        return _ret(out.reshape(shape), like_torch)

    def distort_points(self, points):
        return self._map2("m3d_distort_cam", points)

    def undistort_points(self, points):
        return self._map2("m3d_undistort_cam", points)

    def project(self, points):
        like_torch = _is_torch(points)
        device = points.device.index if like_torch and points.device.type == "cuda" else _default_device()
        rig = self._rig(device)
        src = _to_dev(points, device).reshape(-1, 3)
        out = torch.empty((src.shape[0], 1, 2), dtype=torch.float64, device=src.device)
        _lib.check(rig._lib.m3d_project_cam(rig.handle, 0, _ptr(src), src.shape[0], _ptr(out),
                                            _stream(device)), "m3d_project_cam")
        return _ret(out, like_torch)

    def reprojection_error(self, p3d, p2d):
        proj = self.project(p3d).reshape(p2d.shape)
        return p2d - proj


class FisheyeCamera(Camera):
    """Kannala-Brandt fisheye camera (reference cameras.py:339-426)."""

    _MODEL = _lib.MODEL_FISHEYE
    _N_PARAM_DIST = 4

    def __init__(self, matrix=np.eye(3), dist=np.zeros(4), size=None, rvec=np.zeros(3),
                 tvec=np.zeros(3), name=None, extra_dist=False):
        super().__init__(matrix, dist, size, rvec, tvec, name, extra_dist)

    def get_dict(self):
        d = super().get_dict()
        d['fisheye'] = True
        return d


class OmnidirCamera(Camera):
    """Mei unified omnidirectional camera (reference cameras.py:429-555, the lab's
    default model).  Uses K / xi / D, not matrix / dist, exactly like the reference."""

    _MODEL = _lib.MODEL_OMNIDIR
    _N_PARAM_DIST = 4

    def __init__(self, matrix=np.eye(3), dist=np.zeros(4), size=None, rvec=np.zeros(3),
                 tvec=np.zeros(3), xi=np.zeros(1), K=np.zeros([3, 3]), D=np.zeros(4), name=None,
                 extra_dist=False):
        super().__init__(matrix, dist, size, rvec, tvec, name, extra_dist)
        self.set_xi(xi)
        self.set_K(K)
        self.set_D(D)

    def set_xi(self, xi):
        self.xi = np.array(xi, dtype='float64').ravel()

    def get_xi(self):
        return self.xi

    def set_K(self, K):
        self.K = np.array(K, dtype='float64')

    def get_K(self):
        return self.K

    def set_D(self, D):
        self.D = np.array(D, dtype='float64').ravel()

    def get_D(self):
        return self.D

    def load_dict(self, d):
        super().load_dict(d)
        self.set_xi(d['xi'])
        self.set_K(d['K'])
        self.set_D(d['D'])

    def get_dict(self):
        d = super().get_dict()
        d['Omnidir'] = True   # sic: the reference writes 'Omnidir' but reads 'omnidir' (:481 vs :1977)
        d['xi'] = self.get_xi()
        d['K'] = self.get_K()
        d['D'] = self.get_D()
        return d

    def copy(self):
        return OmnidirCamera(matrix=self.get_camera_matrix().copy(), dist=self.get_distortions().copy(),
                             size=self.get_size(), rvec=self.get_rotation().copy(),
                             tvec=self.get_translation().copy(), xi=self.get_xi().copy(),
                             K=self.get_K().copy(), D=self.get_D().copy(), name=self.get_name(),
                             extra_dist=self.extra_dist)

    def _intrinsics(self):
        return self.K, self.D, float(self.xi[0])


# ----------------------------------------------------------------------------------
# camera group
# ----------------------------------------------------------------------------------

_OUT_OF_SCOPE = ("%s is outside the accelerated hot path (SURVEY.md §8f); "
                 "reference: aniposelib/cameras.py:%s")


class CameraGroup:
    """Reference cameras.py:558-783 on the GPU."""

    def __init__(self, cameras, metadata={}, device=None):
        self.cameras = cameras
        self.metadata = metadata
        self.device = device
        self._rig_cache = None

    # -- bookkeeping --
    def _dev(self):
        return self.device if self.device is not None else _default_device()

    def _rig(self, device=None):
        device = self._dev() if device is None else device
        key = (device, tuple(cam._fingerprint() for cam in self.cameras))
        if self._rig_cache is None or self._rig_cache[0] != key:
            self._rig_cache = (key, _RigHandle(self.cameras, device))
        return self._rig_cache[1]

    def subset_cameras(self, indices):
        cams = [self.cameras[ix].copy() for ix in indices]
        return CameraGroup(cams, self.metadata, self.device)

    def subset_cameras_names(self, names):
        cur_names = self.get_names()
        cur_names_dict = dict(zip(cur_names, range(len(cur_names))))
        indices = []
        for name in names:
            if name not in cur_names_dict:
                raise IndexError(
                    "name {} not part of camera names: {}".format(name, cur_names))
            indices.append(cur_names_dict[name])
        return self.subset_cameras(indices)

    def get_names(self):
        return [cam.get_name() for cam in self.cameras]

    def set_names(self, names):
        for cam, name in zip(self.cameras, names):
            cam.set_name(name)

    def get_rotations(self):
        return np.array([cam.get_rotation() for cam in self.cameras])

    def get_translations(self):
        return np.array([cam.get_translation() for cam in self.cameras])

    def set_rotations(self, rvecs):
        for cam, rvec in zip(self.cameras, rvecs):
            cam.set_rotation(rvec)

    def set_translations(self, tvecs):
        for cam, tvec in zip(self.cameras, tvecs):
            cam.set_translation(tvec)

    def resize_cameras(self, scale):
        for cam in self.cameras:
            cam.resize_camera(scale)

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
            _lib.check(rig._lib.m3d_triangulate_error_host(rig.handle, _np_ptr(src), n, int(bool(undistort)),
                                                           _np_ptr(p3d), _np_ptr(err)),
                       "m3d_triangulate_error_host")
            return p3d, err
        src = _to_dev(points, device)
        p3d = torch.empty((n, 3), dtype=torch.float64, device=src.device)
        err = torch.empty((n,), dtype=torch.float64, device=src.device) if with_err else None
        _lib.check(rig._lib.m3d_triangulate_error(rig.handle, _ptr(src), n, int(bool(undistort)),
                                                  _ptr(p3d), _ptr(err), _stream(device)),
                   "m3d_triangulate_error")
        return p3d, err

    def triangulate_possible(self, points, undistort=True, min_cams=2, progress=False,
                             threshold=0.5, return_stats=False):
        """Given an CxNxPx2 array, triangulate all camera subsets and pick the one with the
        best reprojection error (cameras.py:639-724).  Implemented for P == 1 (one candidate
        per camera), which is the only form the reference's callers use
        (triangulate_ransac, cameras.py:738-743)."""
        self._assert_cams(points)
        n_cams, n_points, n_possible, _ = points.shape
        if n_possible != 1:
            raise NotImplementedError(
                "triangulate_possible with %d candidates per camera: only P == 1 "
                "(triangulate_ransac) is on the accelerated path" % n_possible)
        pts = points.reshape(n_cams, n_points, 2)
        return self._ransac(pts, undistort, min_cams, threshold, 200.0, return_stats)

    def triangulate_ransac(self, points, undistort=True, min_cams=2, progress=False,
                           return_stats=False):
        """Given an CxNx2 array, this returns (points_3d (N,3), picked_vals (C,N,1) bool,
        points_2d (C,N,2), errors (N,)) (cameras.py:726-743).  With return_stats also
        (subset_index (N,) int32, n_evaluated (N,) int32)."""
        self._assert_cams(points)
        n_cams, n_points, _ = points.shape
        return self._ransac(points, undistort, min_cams, 0.5, 200.0, return_stats)

    def _ransac(self, points, undistort, min_cams, threshold, init_best, return_stats):
        C = len(self.cameras)
        n = points.shape[1]
        device, like_torch = self._device_of(points)
        rig = self._rig(device)
        if not like_torch:
            src = np.ascontiguousarray(points, dtype=np.float64)
            p3d = np.empty((n, 3))
            picked = np.empty((C, n, 1), dtype=np.uint8)
            xyp = np.empty((C, n, 2))
            err = np.empty(n)
            sub = np.empty(n, dtype=np.int32)
            nev = np.empty(n, dtype=np.int32)
            _lib.check(rig._lib.m3d_triangulate_ransac_host(
                rig.handle, _np_ptr(src), n, int(bool(undistort)), int(min_cams), float(threshold),
                float(init_best), _np_ptr(p3d), _np_ptr(picked), _np_ptr(xyp), _np_ptr(err),
                _np_ptr(sub), _np_ptr(nev)), "m3d_triangulate_ransac_host")
            res = (p3d, picked.view(np.bool_), xyp, err)
            return res + (sub, nev) if return_stats else res
        src = _to_dev(points, device)
        dev = src.device
        p3d = torch.empty((n, 3), dtype=torch.float64, device=dev)
        picked = torch.empty((C, n, 1), dtype=torch.uint8, device=dev)
        xyp = torch.empty((C, n, 2), dtype=torch.float64, device=dev)
        err = torch.empty((n,), dtype=torch.float64, device=dev)
        sub = torch.empty((n,), dtype=torch.int32, device=dev)
        nev = torch.empty((n,), dtype=torch.int32, device=dev)
        _lib.check(rig._lib.m3d_triangulate_ransac(
            rig.handle, _ptr(src), n, int(bool(undistort)), int(min_cams), float(threshold),
            float(init_best), _ptr(p3d), _ptr(picked), _ptr(xyp), _ptr(err), _ptr(sub), _ptr(nev),
            _stream(device)), "m3d_triangulate_ransac")
        res = (p3d, picked.view(torch.bool), xyp, err)
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

    # -- (de)serialisation (cameras.py:1966-2013) --
    def get_dicts(self):
        return [cam.get_dict() for cam in self.cameras]

    @staticmethod
    def from_dicts(arr):
        cameras = []
        for d in arr:
            if 'fisheye' in d and d['fisheye']:
                cam = FisheyeCamera.from_dict(d)
            elif 'omnidir' in d and d['omnidir']:
                cam = OmnidirCamera.from_dict(d)
            else:
                cam = Camera.from_dict(d)
            cameras.append(cam)
        return CameraGroup(cameras)

    @staticmethod
    def from_names(names, fisheye=False):
        cameras = []
        for name in names:
            cam = FisheyeCamera(name=name) if fisheye else Camera(name=name)
            cameras.append(cam)
        return CameraGroup(cameras)

    def load_dicts(self, arr):
        for cam, d in zip(self.cameras, arr):
            cam.load_dict(d)

    def dump(self, fname):
        import toml
        dicts = self.get_dicts()
        names = ['cam_{}'.format(i) for i in range(len(dicts))]
        master_dict = dict(zip(names, dicts))
        master_dict['metadata'] = self.metadata
        with open(fname, 'w') as f:
            toml.dump(master_dict, f, encoder=toml.TomlNumpyEncoder())

    @staticmethod
    def load(fname):
        import toml
        master_dict = toml.load(fname)
        keys = sorted(master_dict.keys())
        items = [master_dict[k] for k in keys if k != 'metadata']
        cgroup = CameraGroup.from_dicts(items)
        if 'metadata' in master_dict:
            cgroup.metadata = master_dict['metadata']
        return cgroup

    # -- out of scope --
    def optim_points(self, *a, **k):
        raise NotImplementedError(_OUT_OF_SCOPE % ("optim_points", "1116-1190"))

    def optim_points_jointlenfix(self, *a, **k):
        raise NotImplementedError(_OUT_OF_SCOPE % ("optim_points_jointlenfix", "1192-1270"))

    def bundle_adjust(self, *a, **k):
        raise NotImplementedError(_OUT_OF_SCOPE % ("bundle_adjust", "860-946"))

    def bundle_adjust_iter(self, *a, **k):
        raise NotImplementedError(_OUT_OF_SCOPE % ("bundle_adjust_iter", "786-858"))

    def calibrate_rows(self, *a, **k):
        raise NotImplementedError(_OUT_OF_SCOPE % ("calibrate_rows", "1891-1940"))

    def calibrate_videos(self, *a, **k):
        raise NotImplementedError(_OUT_OF_SCOPE % ("calibrate_videos", "1951-1964"))
