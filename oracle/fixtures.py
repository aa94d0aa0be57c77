"""Oracle (test infrastructure): load tests/golden/*.npz written by make_golden.py."""
import os

import numpy as np

from . import camera_math as cm

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def cams_from_arrays(g):
    cams = []
    for i in range(g["rig_model"].shape[0]):
        n = int(g["rig_ndist"][i])
        cams.append(cm.CamSpec(int(g["rig_model"][i]), g["rig_K"][i], g["rig_dist"][i, :n],
                               g["rig_rvec"][i], g["rig_tvec"][i], g["rig_xi"][i],
                               str(g["rig_names"][i])))
    return cams


def cams_from_dicts(dicts):
    """CamSpec list from camera dicts with the reference's get_dict keys."""
    cams = []
    for d in dicts:
        model = cm.MODEL_FISHEYE if d.get("fisheye") else (cm.MODEL_OMNIDIR if d.get("omnidir") else cm.MODEL_PINHOLE)
        if model == cm.MODEL_OMNIDIR:
            cams.append(cm.CamSpec(model, d["K"], d["D"], d["rotation"], d["translation"], d["xi"], d["name"]))
        else:
            cams.append(cm.CamSpec(model, d["matrix"], d["distortions"], d["rotation"], d["translation"],
                                   0.0, d["name"]))
    return cams


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    return g, cams_from_arrays(g)


def golden_names(prefix=""):
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith(prefix))
