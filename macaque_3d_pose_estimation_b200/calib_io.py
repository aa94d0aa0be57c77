"""Calibration assembly of the 3D stage (reference src/pipeline/step4_aniposefiltering.py:101-138 and
src/pipeline/step2_crossviewmatching.py:35-75): the lab keeps intrinsics in ``cam_intrinsic.h5``
(per camera id: mtx, dist, K, xi, D) and extrinsics in ``cam_extrinsic_optim.h5`` (rvec, tvec); step 4
turns them into ``calibration.toml``, step 2 into the ``camparam`` dict.

The readers here work on any mapping with the HDF5 access pattern ``f[cam_id][field][()]`` — an open
``h5py.File`` when h5py is installed, or plain nested dicts of arrays (tests, other stores).  h5py itself is
only imported by ``open_h5`` and is not a dependency of the package."""
import os

import numpy as np

IMAGE_SIZE = [2048, 1536]          # configs/calibration_tmpl.toml:3


def _get(store, cam_id, field):
    node = store[str(cam_id)][field]
    try:
        return np.asarray(node[()], dtype=np.float64)      # h5py dataset
    except (TypeError, IndexError, KeyError):
        return np.asarray(node, dtype=np.float64)          # plain array


def open_h5(path):
    """Open an HDF5 calibration file (needs h5py, which the reference requires as well)."""
    try:
        import h5py
    except ImportError as e:  # pragma: no cover
        raise ImportError("reading %s needs h5py; pass dicts of arrays to assemble_calibration / read_camparam "
                          "instead" % path) from e
    return h5py.File(path, "r")


def assemble_calibration(intrinsic, extrinsic, camera_ids, metadata=None):
    """The dict step 4 writes to calibration.toml (:107-138): one ``cam_<i>`` table per camera id with the
    template's keys (calibration_tmpl.toml: name, size, matrix, distortions, rotation, translation, fisheye,
    omnidir) plus xi / K / D.  Like the reference, the first two rows of ``mtx`` are halved (:119) — the 2D
    stage runs on half-resolution frames — while the omnidirectional ``K`` is stored as is."""
    calib = {}
    for i, cid in enumerate(str(c) for c in camera_ids):
        mtx = _get(intrinsic, cid, "mtx").copy()
        mtx[:2, :] /= 2
        calib["cam_%d" % i] = {
            "name": cid, "size": list(IMAGE_SIZE), "matrix": mtx.tolist(),
            "distortions": _get(intrinsic, cid, "dist").ravel().tolist(),
            "rotation": _get(extrinsic, cid, "rvec").ravel().tolist(),
            "translation": _get(extrinsic, cid, "tvec").ravel().tolist(),
            "fisheye": False, "omnidir": True,
            "xi": _get(intrinsic, cid, "xi").ravel().tolist(), "K": _get(intrinsic, cid, "K").tolist(),
            "D": _get(intrinsic, cid, "D").ravel().tolist()}
    calib["metadata"] = dict(metadata) if metadata is not None else {"adjusted": False, "error": 0.0}
    return calib


def write_calibration(result_dir, intrinsic, extrinsic, camera_ids, metadata=None):
    """assemble_calibration -> ``<result_dir>/calibration.toml`` (:138); returns the dict."""
    import toml
    calib = assemble_calibration(intrinsic, extrinsic, camera_ids, metadata)
    with open(os.path.join(result_dir, "calibration.toml"), "w") as f:
        toml.dump(calib, f)
    return calib


def camera_group_from_calibration(calib):
    """CameraGroup from the assembled dict — what ``CameraGroup.load(calibration.toml)`` returns (:212),
    without the file round trip."""
    from .cameras import CameraGroup
    keys = sorted((k for k in calib if k != "metadata"), key=lambda k: int(k.split("_")[1]))
    cg = CameraGroup.from_dicts([calib[k] for k in keys])
    cg.metadata = calib.get("metadata", {})
    return cg


def read_camparam(intrinsic, extrinsic, camera_ids):
    """step2's get_camparam (:35-75): K, xi, D, rvecs, tvecs and pmat = [R | t] per camera; additionally the
    pinhole ``mtx`` / ``dist`` that mct.undistortPoints(omnidir=False) reads (multicam_toolbox.py:423-425)."""
    from .csrc_host import rodrigues
    out = {"camera_id": list(camera_ids), "K": [], "xi": [], "D": [], "rvecs": [], "tvecs": [], "pmat": [],
           "mtx": [], "dist": []}
    for cid in camera_ids:
        out["K"].append(_get(intrinsic, cid, "K"))
        out["xi"].append(_get(intrinsic, cid, "xi"))
        out["D"].append(_get(intrinsic, cid, "D"))
        rvec, tvec = _get(extrinsic, cid, "rvec"), _get(extrinsic, cid, "tvec")
        out["rvecs"].append(rvec)
        out["tvecs"].append(tvec)
        out["pmat"].append(np.hstack([rodrigues(rvec.ravel()), tvec.reshape(3, 1)]))
        try:
            out["mtx"].append(_get(intrinsic, cid, "mtx"))
            out["dist"].append(_get(intrinsic, cid, "dist"))
        except KeyError:
            pass
    if len(out["mtx"]) != len(out["K"]):
        del out["mtx"], out["dist"]
    return out
