"""ctypes binding of libm3d.so (include/m3d.h).

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a) and
is the ONLY compute path of this package: if it is missing or no GPU is visible the
entry points raise — there is no CPU fallback.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libm3d.so")

MAX_CAMS = 16
MAX_DETS = 128
MAX_JOINTS = 32
MODEL_PINHOLE, MODEL_FISHEYE, MODEL_OMNIDIR = 0, 1, 2

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_uint8_p = ctypes.POINTER(ctypes.c_uint8)


class M3DCam(ctypes.Structure):
    """struct m3d_cam of include/m3d.h."""
    _fields_ = [
        ("model", ctypes.c_int32),
        ("n_dist", ctypes.c_int32),
        ("K", ctypes.c_double * 9),
        ("dist", ctypes.c_double * 14),
        ("rvec", ctypes.c_double * 3),
        ("tvec", ctypes.c_double * 3),
        ("xi", ctypes.c_double),
    ]


# name -> (restype, argtypes); every symbol include/m3d.h declares
_P = ctypes.c_void_p
_I = ctypes.c_int32
_L = ctypes.c_int64
_D = ctypes.c_double
SIGNATURES = {
    "m3d_version": (ctypes.c_int, []),
    "m3d_last_error": (ctypes.c_char_p, []),
    "m3d_device_count": (ctypes.c_int, []),
    "m3d_rig_create": (ctypes.c_int, [ctypes.POINTER(M3DCam), _I, _I, ctypes.POINTER(_P)]),
    "m3d_rig_destroy": (None, [_P]),
    "m3d_rig_num_cams": (_I, [_P]),
    "m3d_rig_device": (_I, [_P]),
    "m3d_rig_extrinsics": (ctypes.c_int, [_P, _P]),
    "m3d_rig_set_ransac_mode": (ctypes.c_int, [_P, _I]),
    "m3d_rig_certified_mask": (_I, [_P]),
    "m3d_undistort_cam": (ctypes.c_int, [_P, _I, _P, _L, _P, _P]),
    "m3d_project_cam": (ctypes.c_int, [_P, _I, _P, _L, _P, _P]),
    "m3d_distort_cam": (ctypes.c_int, [_P, _I, _P, _L, _P, _P]),
    "m3d_undistort": (ctypes.c_int, [_P, _P, _L, _P, _P]),
    "m3d_project": (ctypes.c_int, [_P, _P, _L, _P, _P]),
    "m3d_triangulate": (ctypes.c_int, [_P, _P, _L, _I, _P, _P]),
    "m3d_reproj_error": (ctypes.c_int, [_P, _P, _P, _L, _I, _P, _P]),
    "m3d_triangulate_error": (ctypes.c_int, [_P, _P, _L, _I, _P, _P, _P]),
    "m3d_triangulate_ransac": (ctypes.c_int, [_P, _P, _L, _I, _I, _D, _D, _P, _P, _P, _P, _P, _P, _P]),
    "m3d_triangulate_possible": (ctypes.c_int, [_P, _P, _L, _I, _I, _I, _D, _D, _P, _P, _P, _P, _P, _P, _P]),
    "m3d_triangulate_error_host": (ctypes.c_int, [_P, _P, _L, _I, _P, _P]),
    "m3d_triangulate_ransac_host": (ctypes.c_int, [_P, _P, _L, _I, _I, _D, _D, _P, _P, _P, _P, _P, _P]),
    "m3d_triangulate_error_host_f32": (ctypes.c_int, [_P, _P, _L, _I, _P, _P]),
    "m3d_triangulate_ransac_host_f32": (ctypes.c_int, [_P, _P, _L, _I, _I, _D, _D, _P, _P, _P, _P, _P, _P]),
    "m3d_triangulate_error_host_span": (ctypes.c_int, [_P, _P, _L, _L, _L, _I, _P, _P]),
    "m3d_triangulate_ransac_host_span": (ctypes.c_int, [_P, _P, _L, _L, _L, _I, _I, _D, _D, _P, _P, _P, _P, _P, _P]),
    "m3d_host_register": (ctypes.c_int, [_P, _L]),
    "m3d_host_unregister": (ctypes.c_int, [_P]),
    "m3d_ray_affinity": (ctypes.c_int, [_P, _P, _P, _I, _I, _I, _D, _P, _P, _P]),
    "m3d_triangulate_ls": (ctypes.c_int, [_P, _P, _P, _L, _P, _P]),
    "m3d_match_svt": (ctypes.c_int, [_P, _P, _I, _I, _I, _D, _D, _D, _D, _I, _P, _P, _I, _P]),
    "m3d_association_weights": (ctypes.c_int, [_P, _P, _P, _I, _I, _I, _D, _P, _I, _P]),
    "m3d_match_clusters": (ctypes.c_int, [_P, _P, _I, _I, _I, _P, _I, _P]),
    "m3d_viterbi_filter": (ctypes.c_int, [_P, _L, _L, _I, _I, _D, _D, _D, _P, _P, _I, _P]),
    "m3d_optim_points": (ctypes.c_int, [_P, _P, _P, _I, _I, _P, _I, _P, _I, _D, _D, _D, _D, _I, _I, _I, _D, _I, _I,
                                        _P, _P, _P, _P]),
    "m3d_undistort_detections": (ctypes.c_int, [_P, _P, _P, _L, _I, _I, _P, _P]),
    "m3d_cluster_members": (ctypes.c_int, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _P]),
    "m3d_triangulate_ls_members": (ctypes.c_int, [_P, _P, _P, _P, _L, _I, _I, _D, _P, _P]),
    "m3d_peer_alloc": (ctypes.c_int, [_I, _L, ctypes.POINTER(_P), _P]),
    "m3d_peer_open": (ctypes.c_int, [_I, _P, ctypes.POINTER(_P)]),
    "m3d_peer_push": (ctypes.c_int, [_P, _P, _L, _P]),
    "m3d_peer_close": (ctypes.c_int, [_I, _P]),
    "m3d_peer_free": (ctypes.c_int, [_I, _P]),
    "m3d_launch_count": (_L, []),
    "m3d_profile_enable": (ctypes.c_int, [_I]),
    "m3d_profile_read": (ctypes.c_int, [ctypes.c_char_p, _L]),
    "m3d_probe_fp64_tflops": (ctypes.c_int, [_I, ctypes.POINTER(_D)]),
}

_lib = None


def load():
    """Load libm3d.so and set the prototypes.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libm3d.so not found at %s — build it with `python -c \"import __graft_entry__ as g; "
            "g.build()\"` (nvcc, sm_100a). This package has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    msg = load().m3d_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("libm3d %s failed (%d): %s" % (what, rc, last_error()))


def require_gpu():
    """Fail loudly when the CUDA path cannot run (no fallback exists)."""
    lib = load()
    if lib.m3d_device_count() <= 0:
        raise RuntimeError("libm3d: no CUDA device visible; the B200 path has no CPU fallback")
    return lib
