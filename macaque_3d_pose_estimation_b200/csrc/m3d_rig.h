// m3d_rig.h — host-side construction of the device rig record from the C-ABI camera
// structs (include/m3d.h).  Plain C++ (no CUDA) so that the test harness can reuse it.
//
// Reference: Camera.__init__ / set_* (cameras.py:174-250), OmnidirCamera (cameras.py:429-472),
// make_M (utils.py:9-15) -> cv2.Rodrigues.
#pragma once
#include <cmath>
#include <cstring>
#include <string>

#include "../../include/m3d.h"
#include "m3d_math.cuh"

namespace m3d {

// cv2.Rodrigues(rvec): theta < DBL_EPSILON -> I, else
// R = cos(t) I + (1 - cos t) r r^T + sin(t) [r]x   (same operation order as OpenCV)
inline void rodrigues(const double rvec[3], double R[9]) {
  const double theta = std::sqrt(rvec[0] * rvec[0] + rvec[1] * rvec[1] + rvec[2] * rvec[2]);
  if (theta < 2.220446049250313e-16) {
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    std::memcpy(R, I, sizeof(I));
    return;
  }
  const double c = std::cos(theta), s = std::sin(theta), c1 = 1.0 - c;
  const double it = 1.0 / theta;
  const double rx = rvec[0] * it, ry = rvec[1] * it, rz = rvec[2] * it;
  const double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
  const double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 9; ++i) R[i] = c * I[i] + c1 * rrt[i] + s * r_x[i];
}

// Returns an empty string on success, else the reason the rig is rejected.
inline std::string build_rig(const m3d_cam* cams, int n_cams, RigDev* rig) {
  if (!cams || !rig) return "null camera array";
  if (n_cams < 0 || n_cams > M3D_MAX_CAMS)
    return "number of cameras must be in [0, " + std::to_string(M3D_MAX_CAMS) + "]";
  std::memset(rig, 0, sizeof(RigDev));
  rig->n_cams = n_cams;
  int flags = 0;
  for (int i = 0; i < n_cams; ++i) {
    const m3d_cam& in = cams[i];
    CamDev& c = rig->cam[i];
    if (in.model != M3D_MODEL_PINHOLE && in.model != M3D_MODEL_FISHEYE && in.model != M3D_MODEL_OMNIDIR)
      return "camera " + std::to_string(i) + ": unknown model";
    c.model = in.model;
    c.fx = in.K[0];
    c.skew = in.K[1];
    c.cx = in.K[2];
    c.fy = in.K[4];
    c.cy = in.K[5];
    c.ifx = 1.0 / c.fx;
    c.ify = 1.0 / c.fy;
    c.xi = in.xi;
    const int nd = in.n_dist;
    if (in.model == M3D_MODEL_PINHOLE) {
      if (!(nd == 4 || nd == 5 || nd == 8 || nd == 12 || nd == 14))
        return "camera " + std::to_string(i) + ": pinhole distortion vector must have 4, 5, 8, 12 or 14 entries";
      if (nd == 14 && (in.dist[12] != 0.0 || in.dist[13] != 0.0))
        return "camera " + std::to_string(i) + ": tilted sensor model (tauX, tauY) is not supported";
      for (int j = 0; j < 12 && j < nd; ++j) c.k[j] = in.dist[j];
      if (c.k[5] != 0.0 || c.k[6] != 0.0 || c.k[7] != 0.0) flags |= RIG_HAS_RATIONAL;
      if (c.k[8] != 0.0 || c.k[9] != 0.0 || c.k[10] != 0.0 || c.k[11] != 0.0) flags |= RIG_HAS_PRISM;
    } else {
      if (nd < 4) return "camera " + std::to_string(i) + ": fisheye / omnidir need 4 distortion coefficients";
      for (int j = 0; j < 4; ++j) c.k[j] = in.dist[j];
      flags |= RIG_HAS_NONPINHOLE;
    }
    for (int j = 0; j < 5; ++j) c.kf[j] = (float)c.k[j];
    c.kf[5] = 0.0f;
    rodrigues(in.rvec, c.R);
    c.t[0] = in.tvec[0];
    c.t[1] = in.tvec[1];
    c.t[2] = in.tvec[2];
  }
  rig->flags = flags;
  return std::string();
}

}  // namespace m3d
