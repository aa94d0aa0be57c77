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
from ._device import (_RigHandle, _default_device, _is_torch, _np_ptr, _ptr, _ret, _stream, _to_dev,  # noqa: F401
                      torch)
from .camera_group import CameraGroup  # noqa: F401
from .camera_models import Camera, FisheyeCamera, OmnidirCamera  # noqa: F401

__all__ = ["Camera", "FisheyeCamera", "OmnidirCamera", "CameraGroup"]
