"""Camera models of the drop-in API: pinhole ``Camera``, ``FisheyeCamera``, ``OmnidirCamera``.

Public surface = the reference's aniposelib/cameras.py:173-555 (constructor arguments,
get_* / set_* accessors, get_dict / load_dict / from_dict, get_params / set_params,
resize_camera, copy, distort_points / undistort_points / project / reprojection_error,
get_extrinsics_mat).  The accessors are generated from one field table; every point map runs
in libm3d.so on the GPU.
"""
import numpy as np

from . import _lib
from ._device import (_RigHandle, _default_device, _is_torch, _np_ptr, _ptr, _ret, _stream, _to_dev, torch)


def _vec(v):
    return np.array(v, dtype='float64').ravel()


def _mat(v):
    return np.array(v, dtype='float64')


# public name, attribute, converter applied by the setter, key in get_dict()
_FIELDS = (
    ("camera_matrix", "matrix", _mat, "matrix"),
    ("distortions", "dist", _vec, "distortions"),
    ("rotation", "rvec", _vec, "rotation"),
    ("translation", "tvec", _vec, "translation"),
    ("name", "name", str, "name"),
    ("size", "size", lambda v: v, "size"),       # (width, height)
)


class Camera:
    """Pinhole camera with OpenCV distortion coefficients."""

    _MODEL = _lib.MODEL_PINHOLE
    _N_PARAM_DIST = 5        # length of the distortion vector rebuilt by set_params
    _EXTRA = ()              # extra (name, attribute, converter, dict key) fields of subclasses
    _DICT_FLAG = None        # model marker written by get_dict

    def __init__(self, matrix=np.eye(3), dist=None, size=None, rvec=np.zeros(3), tvec=np.zeros(3),
                 name=None, extra_dist=False, **extra):
        if dist is None:
            dist = np.zeros(self._N_PARAM_DIST)
        for (pub, _, _, _), val in zip(_FIELDS, (matrix, dist, rvec, tvec, name, size)):
            getattr(self, "set_" + pub)(val)
        for pub, _, _, _ in self._EXTRA:
            getattr(self, "set_" + pub)(extra.pop(pub))
        if extra:
            raise TypeError("unexpected arguments: " + ", ".join(sorted(extra)))
        self.extra_dist = extra_dist
        self._rig_cache = None

    # -- serialisation ---------------------------------------------------------------------
    def get_dict(self):
        d = {key: getattr(self, "get_" + pub)() for pub, _, _, key in _FIELDS}
        d = {k: (list(v) if k == "size" else v.tolist() if isinstance(v, np.ndarray) else v) for k, v in d.items()}
        if self._DICT_FLAG:
            d[self._DICT_FLAG] = True
        for pub, _, _, key in self._EXTRA:
            d[key] = getattr(self, "get_" + pub)()
        return d

    def load_dict(self, d):
        for pub, _, _, key in _FIELDS + tuple(self._EXTRA):
            getattr(self, "set_" + pub)(d[key])

    @classmethod
    def from_dict(cls, d):
        cam = cls()
        cam.load_dict(d)
        return cam

    def copy(self):
        kw = {pub: np.copy(getattr(self, attr)) for pub, attr, _, _ in self._EXTRA}
        return type(self)(matrix=self.matrix.copy(), dist=self.dist.copy(), size=self.size,
                          rvec=self.rvec.copy(), tvec=self.tvec.copy(), name=self.name,
                          extra_dist=self.extra_dist, **kw)

    # -- focal length / optimiser parameter vector -----------------------------------------------
    def set_focal_length(self, fx, fy=None):
        self.matrix[0, 0] = fx
        self.matrix[1, 1] = fx if fy is None else fy

    def get_focal_length(self, both=False):
        fx, fy = self.matrix[0, 0], self.matrix[1, 1]
        return (fx, fy) if both else (fx + fy) / 2.0

    def resize_camera(self, scale):
        """Scale the image size and the intrinsics with it."""
        w, h = self.size
        m = self.matrix * scale
        m[2, 2] = 1
        self.set_size((w * scale, h * scale))
        self.set_camera_matrix(m)

    def get_params(self):
        n = 8 + self.extra_dist
        out = np.zeros(n, dtype='float64')
        out[0:3], out[3:6], out[6] = self.rvec, self.tvec, self.get_focal_length()
        out[7:n] = self.dist[:n - 7]
        return out

    def set_params(self, params):
        self.set_rotation(params[0:3])
        self.set_translation(params[3:6])
        self.set_focal_length(params[6])
        dist = np.zeros(self._N_PARAM_DIST, dtype='float64')
        n = 8 + self.extra_dist
        dist[:n - 7] = params[7:n]
        self.set_distortions(dist)

    # -- C-ABI record ---------------------------------------------------------------------------
    def _intrinsics(self):
        """(K, distortion vector, xi) handed to the kernels."""
        return self.matrix, self.dist, 0.0

    def _fill_struct(self, s):
        K, dist, xi = self._intrinsics()
        K = np.asarray(K, dtype=np.float64).reshape(3, 3)
        dist = np.asarray(dist, dtype=np.float64).ravel()
        if dist.size > 14:
            raise ValueError("distortion vector longer than 14 entries")
        s.model, s.n_dist, s.xi = self._MODEL, int(dist.size), float(xi)
        s.K[:] = [float(v) for v in K.flat]
        s.dist[:] = [float(dist[i]) if i < dist.size else 0.0 for i in range(14)]
        s.rvec[:] = [float(v) for v in self.rvec]
        s.tvec[:] = [float(v) for v in self.tvec]

    def _fingerprint(self):
        K, dist, xi = self._intrinsics()
        return (self._MODEL, np.asarray(K, dtype=np.float64).tobytes(),
                np.asarray(dist, dtype=np.float64).tobytes(), float(xi), self.rvec.tobytes(), self.tvec.tobytes())

    def _rig(self, device):
        key = (device, self._fingerprint())
        cache = getattr(self, "_rig_cache", None)
        if cache is None or cache[0] != key:
            self._rig_cache = (key, _RigHandle([self], device))
        return self._rig_cache[1]

    def get_extrinsics_mat(self):
        """4x4 [R|t], R = Rodrigues(rvec)."""
        rig = self._rig(_default_device())
        M = np.empty((1, 4, 4))
        _lib.check(rig._lib.m3d_rig_extrinsics(rig.handle, _np_ptr(M)), "m3d_rig_extrinsics")
        return M[0]

    # -- point maps -------------------------------------------------------------------------------
    def _map(self, fn_name, points, n_in, out_shape):
        like_torch = _is_torch(points)
        device = points.device.index if like_torch and points.device.type == "cuda" else _default_device()
        rig = self._rig(device)
        src = _to_dev(points, device).reshape(-1, n_in)
        out = torch.empty((src.shape[0], 2), dtype=torch.float64, device=src.device)
        fn = getattr(rig._lib, fn_name)
        _lib.check(fn(rig.handle, 0, _ptr(src), src.shape[0], _ptr(out), _stream(device)), fn_name)
        return _ret(out.reshape(out_shape(src.shape[0])), like_torch)

    def distort_points(self, points):
        shape = tuple(points.shape)
        return self._map("m3d_distort_cam", points, 2, lambda n: shape)

    def undistort_points(self, points):
        shape = tuple(points.shape)
        return self._map("m3d_undistort_cam", points, 2, lambda n: shape)

    def project(self, points):
        return self._map("m3d_project_cam", points, 3, lambda n: (n, 1, 2))

    def reprojection_error(self, p3d, p2d):
        return p2d - self.project(p3d).reshape(p2d.shape)


def _install_accessors(cls, fields):
    for pub, attr, conv, _ in fields:
        setattr(cls, "get_" + pub, (lambda a: lambda self: getattr(self, a))(attr))
        setattr(cls, "set_" + pub, (lambda a, c: lambda self, value: setattr(self, a, c(value)))(attr, conv))


_install_accessors(Camera, _FIELDS)


class FisheyeCamera(Camera):
    """Kannala-Brandt fisheye camera (four distortion coefficients)."""
    _MODEL = _lib.MODEL_FISHEYE
    _N_PARAM_DIST = 4
    _DICT_FLAG = "fisheye"


class OmnidirCamera(Camera):
    """Mei unified omnidirectional camera: uses K / xi / D (not matrix / dist) for its point maps.
    get_dict writes the marker 'Omnidir' while from_dicts looks for 'omnidir' — the reference's
    own mismatch, kept so that dump / load round-trips behave identically."""
    _MODEL = _lib.MODEL_OMNIDIR
    _N_PARAM_DIST = 4
    _DICT_FLAG = "Omnidir"
    _EXTRA = (("xi", "xi", _vec, "xi"), ("K", "K", _mat, "K"), ("D", "D", _vec, "D"))

    def __init__(self, matrix=np.eye(3), dist=None, size=None, rvec=np.zeros(3), tvec=np.zeros(3),
                 xi=np.zeros(1), K=np.zeros([3, 3]), D=np.zeros(4), name=None, extra_dist=False):
        super().__init__(matrix, dist, size, rvec, tvec, name, extra_dist, xi=xi, K=K, D=D)

    def _intrinsics(self):
        return self.K, self.D, float(self.xi[0])


_install_accessors(OmnidirCamera, OmnidirCamera._EXTRA)
