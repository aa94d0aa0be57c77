"""CPU oracle for the multi-view 3D reconstruction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker or as the timed CPU baseline.  The product path
(``macaque_3d_pose_estimation_b200``) never imports this package and has no CPU
fallback.

The oracle is a numpy restatement of the algorithms the reference runs through
OpenCV / numba / LAPACK:

* ``camera_math``  – OpenCV pinhole / fisheye / omnidir point maps and
  ``cv2.Rodrigues`` (reference: src/third_party/aniposelib/cameras.py:301-516,
  utils.py:9-15).
* ``cameragroup``  – ``CameraGroup.triangulate / reprojection_error /
  triangulate_possible / triangulate_ransac`` (cameras.py:20-32, 593-783), in a
  vectorised form (fast checker) and a loop-faithful form (CPU baseline with
  the reference's cost structure).
* ``crossview``    – ``geometry_affinity2``, ``calc_dist_btw_lines``,
  ``matchSVT`` (src/pipeline/step2_crossviewmatching.py:130-216, 327-432) and
  ``mct.triangulatePoints`` (src/utils/multicam_toolbox.py:433-486).

Parity pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md §4), so the oracle is pinned by EXECUTION of the reference itself in
the build container: ``oracle/make_golden.py`` imports the unmodified reference
from /root/reference, runs it on seeded synthetic rigs and writes
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every oracle
function against those vectors.  Exception: the omnidir (Mei) camera model —
``cv2.omnidir`` is not installed here, so that model is "parity unpinned"
(restated from the published opencv_contrib ccalib algorithm only).
"""
