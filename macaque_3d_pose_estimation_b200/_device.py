"""Device-memory / stream plumbing shared by the host classes (PyTorch is used for device
buffers and streams only — never as a compute path)."""
import ctypes

import numpy as np

from . import _lib

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


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


