"""Calibration assembly (step4_aniposefiltering.py:101-138) and camparam reading (step2:35-75) from HDF5-like
stores: checked against an expectation built by EXECUTING the reference's own assembly code on an h5py stand-in
when /root/reference is present (build container), and structurally everywhere."""
import os

import numpy as np
import pytest

from macaque_3d_pose_estimation_b200 import calib_io, csrc_host, synth


class FakeDataset:
    def __init__(self, a):
        self.a = np.asarray(a)

    def __getitem__(self, k):
        assert k == ()
        return self.a.copy()


def stores(seed=3, n=4):
    rng = np.random.default_rng(seed)
    rig = synth.make_rig(n, "omnidir", seed=seed)
    intr, extr, ids = {}, {}, []
    for d in rig:
        cid = str(int(d["name"]) + 10)
        ids.append(int(cid))
        intr[cid] = {"mtx": FakeDataset(np.array(d["matrix"]) * 2.0), "dist": FakeDataset(rng.normal(0, 0.1, (1, 5))),
                     "K": FakeDataset(d["K"]), "xi": FakeDataset(np.array(d["xi"]).reshape(1, 1)),
                     "D": FakeDataset(np.array(d["D"]).reshape(1, 4))}
        extr[cid] = {"rvec": FakeDataset(np.array(d["rotation"]).reshape(3, 1)),
                     "tvec": FakeDataset(np.array(d["translation"]).reshape(3, 1))}
    return intr, extr, ids


def test_assemble_calibration_layout_and_roundtrip(tmp_path):
    intr, extr, ids = stores()
    calib = calib_io.write_calibration(str(tmp_path), intr, extr, ids)
    assert sorted(k for k in calib if k != "metadata") == ["cam_%d" % i for i in range(len(ids))]
    c0 = calib["cam_0"]
    assert c0["name"] == str(ids[0]) and c0["omnidir"] is True and c0["fisheye"] is False
    assert c0["size"] == [2048, 1536]
    mtx = intr[str(ids[0])]["mtx"][()]
    assert np.allclose(np.array(c0["matrix"])[:2], mtx[:2] / 2) and np.allclose(np.array(c0["matrix"])[2], mtx[2])
    assert np.allclose(c0["K"], intr[str(ids[0])]["K"][()])             # the omnidir K is NOT halved
    import toml
    again = toml.load(os.path.join(str(tmp_path), "calibration.toml"))
    assert np.allclose(again["cam_1"]["rotation"], extr[str(ids[1])]["rvec"][()].ravel())
    assert again["cam_1"]["omnidir"] is True and len(again["cam_1"]["D"]) == 4


def test_read_camparam_matches_rodrigues():
    import cv2
    intr, extr, ids = stores(5)
    cp = calib_io.read_camparam(intr, extr, ids)
    assert cp["camera_id"] == ids and len(cp["pmat"]) == len(ids) and "mtx" in cp
    for i, cid in enumerate(ids):
        R, _ = cv2.Rodrigues(extr[str(cid)]["rvec"][()])
        assert np.abs(cp["pmat"][i][:, :3] - R).max() <= 1e-15
        assert np.array_equal(cp["pmat"][i][:, 3], extr[str(cid)]["tvec"][()].ravel())
        assert np.abs(csrc_host.rodrigues(extr[str(cid)]["rvec"][()]) - R).max() <= 1e-15


@pytest.mark.skipif(not os.path.exists("/root/reference/src/pipeline/step4_aniposefiltering.py"),
                    reason="the reference tree is only present in the build container")
def test_assembly_equals_executed_reference_code(tmp_path):
    """Run the reference's OWN assembly statements (step4_aniposefiltering.py:107-138, sliced out of its source
    and executed with an h5py stand-in that serves the fake stores) and compare the resulting calibration dict."""
    import toml
    intr, extr, ids = stores(7)
    src = open("/root/reference/src/pipeline/step4_aniposefiltering.py").read().splitlines()
    a = next(i for i, l in enumerate(src) if "calib = toml.load(open('./configs/calibration_tmpl.toml'))" in l)
    b = next(i for i, l in enumerate(src) if "toml.dump(calib, open(result_dir + '/calibration.toml'" in l)
    block = [l[4:] for l in src[a + 1:b]]                                    # de-indent the function body
    block = [l for l in block if "yaml.safe_load" not in l and "with open(config_path" not in l]

    class FakeFile:
        def __init__(self, store):
            self.store = store

        def __enter__(self):
            return self.store

        def __exit__(self, *a):
            return False

    class FakeH5:
        @staticmethod
        def File(path, mode="r"):
            return FakeFile(intr if "intrinsic" in path else extr)
    calib = toml.load(open("/root/reference/configs/calibration_tmpl.toml"))
    calib = {k: v for k, v in calib.items() if k == "metadata" or int(k.split("_")[1]) < len(ids)}
    env = {"calib": calib, "h5py": FakeH5, "os": os, "np": np, "config_path": "/x/config.yaml", "cfg": {"camera_id": list(ids)}}
    exec("\n".join(block), env)
    ref = env["calib"]
    mine = calib_io.assemble_calibration(intr, extr, ids, metadata=ref["metadata"])
    assert set(mine) == set(ref)
    for k in ref:
        if k == "metadata":
            continue
        assert set(mine[k]) == set(ref[k]), (k, set(mine[k]) ^ set(ref[k]))
        for f in ref[k]:
            if isinstance(ref[k][f], (list, tuple)):
                assert np.allclose(np.array(mine[k][f], dtype=float), np.array(ref[k][f], dtype=float)), (k, f)
            else:
                assert mine[k][f] == ref[k][f], (k, f)
