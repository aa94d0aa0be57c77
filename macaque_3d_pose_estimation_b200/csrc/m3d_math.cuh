// m3d_math.cuh — per-point arithmetic of the hot path (camera maps, DLT solve,
// reprojection error).  Every function is __host__ __device__ so that the test-only
// host harness (tests/host_harness.cpp) can compile the SAME source with g++ and check
// it against the oracle without a GPU; the product only ever runs the device build.
//
// Reference semantics (file:line under /root/reference/src/third_party/aniposelib/):
//   Camera.undistort_points  cameras.py:310  cv2.undistortPoints, 5 fixed-point iterations
//   Camera.project           cameras.py:318  cv2.projectPoints
//   FisheyeCamera.*          cameras.py:376,384
//   OmnidirCamera.*          cameras.py:498,509  (opencv_contrib ccalib; parity unpinned)
//   triangulate_simple       cameras.py:20-32  DLT rows + smallest right singular vector
//   reprojection_error       cameras.py:746-783
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define M3D_HD __host__ __device__ __forceinline__
#define M3D_HD_NOINLINE inline __host__ __device__ __noinline__
#else
#define M3D_HD inline
#define M3D_HD_NOINLINE inline
#endif

#define M3D_MAXC 16

namespace m3d {

enum { PINHOLE = 0, FISHEYE = 1, OMNIDIR = 2 };
// rig-level feature flags (select leaner kernel instantiations)
enum { RIG_HAS_RATIONAL = 1, RIG_HAS_PRISM = 2, RIG_HAS_NONPINHOLE = 4 };

struct CamDev {
  double fx, fy, cx, cy, skew;
  double ifx, ify;  // OpenCV multiplies by 1/fx, 1/fy in undistortPoints
  double k[12];     // k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4 | fisheye k1..k4 | omnidir k1 k2 p1 p2
  double R[9];      // row-major rotation (cv2.Rodrigues of rvec)
  double t[3];
  double xi;
  int32_t model;
  int32_t pad;
  float kf[6];      // k1 k2 p1 p2 k3 in float32 (undistort_pinhole_fast), padding
};

struct RigDev {
  int32_t n_cams;
  int32_t flags;
  CamDev cam[M3D_MAXC];
};

M3D_HD double qnan() {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(0x7ff8000000000000LL);
#else
  return NAN;
#endif
}

// Reciprocal / square root.  CUDA's IEEE-rounded double division and sqrt expand to a
// MUFU seed, Newton steps and a subroutine CALL on every use (measured: 46 CALLs per
// joint-instance in the DLT kernel).  On the device we take the hardware seed
// (MUFU.RCP64H / MUFU.RSQ64H, ~2^-20) and two Newton steps (2^-40, 2^-80), branch-free:
// the result is within 1 ulp of the IEEE value, far inside the parity budget of the path
// (the reference's own LAPACK solve carries ~1e-10 px).  Special operands: NaN -> NaN as
// in IEEE; 0, inf and denormal operands give NaN instead of inf / 0 — every call site
// either guards them (z == 0 in projectPoints, zero residual) or is already in
// garbage-in territory where the reference's own result is inf/NaN noise.
M3D_HD double rcp(double a) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double e = fma(-a, y, 1.0);
  y = fma(y, e, y);
  e = fma(-a, y, 1.0);
  y = fma(y, e, y);
  return y;
#else
  return 1.0 / a;
#endif
}

// One Newton step on the hardware seed: relative error ~2^-40.  For quantities that only steer an
// iteration (the Newton step of dlt_solve_warp), never for a value that is returned.
M3D_HD double rcp_coarse(double a) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double e = fma(-a, y, 1.0);
  return fma(y, e, y);
#else
  return 1.0 / a;
#endif
}

M3D_HD double sqrt_fast(double a) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double g = a * y;        // ~ sqrt(a)
  double h = 0.5 * y;      // ~ 1 / (2 sqrt(a))
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  const double d = fma(-g, g, a);
  g = fma(d, h, g);
  return (a == 0.0) ? 0.0 : g;   // exact zero residual stays zero
#else
  return sqrt(a);
#endif
}

// ---------------------------------------------------------------------------------------
// undistortion: pixel -> normalised coordinates
// ---------------------------------------------------------------------------------------

// cv2.undistortPoints(pts, K, dist): x0 = (u-cx)/fx; exactly 5 iterations of the
// fixed-point update, bail-out to (x0, y0) when icdist < 0 (OpenCV
// cvUndistortPointsInternal, TermCriteria(MAX_ITER, 5)).  FULL = rational (k4..k6) and
// thin-prism (s1..s4) terms present.
template <bool FULL>
M3D_HD void undistort_pinhole_exact(const CamDev& c, double u, double v, double& xo, double& yo) {
  // literal transcription, taken only when the icdist < 0 bail-out fires
  const double x0 = (u - c.cx) * c.ifx;
  const double y0 = (v - c.cy) * c.ify;
  double x = x0, y = y0;
  for (int j = 0; j < 5; ++j) {
    const double r2 = x * x + y * y;
    double icdist;
    if (FULL) {
      icdist = (1.0 + ((c.k[7] * r2 + c.k[6]) * r2 + c.k[5]) * r2) /
               (1.0 + ((c.k[4] * r2 + c.k[1]) * r2 + c.k[0]) * r2);
    } else {
      icdist = 1.0 / (1.0 + ((c.k[4] * r2 + c.k[1]) * r2 + c.k[0]) * r2);
    }
    if (icdist < 0.0) {
      x = x0;
      y = y0;
      break;
    }
    double dx = 2.0 * c.k[2] * x * y + c.k[3] * (r2 + 2.0 * x * x);
    double dy = c.k[2] * (r2 + 2.0 * y * y) + 2.0 * c.k[3] * x * y;
    if (FULL) {
      dx += c.k[8] * r2 + c.k[9] * r2 * r2;
      dy += c.k[10] * r2 + c.k[11] * r2 * r2;
    }
    x = (x0 - dx) * icdist;
    y = (y0 - dy) * icdist;
  }
  xo = x;
  yo = y;
}

// The five fixed-point iterations without the bail-out: returns a negative value when some icdist had its
// sign bit set (icdist < 0, or -0 / negative NaN, which the literal transcription handles identically to
// this loop) — the caller then replays the point through undistort_pinhole_exact.  Straight-line: several
// cameras of one point can be in flight at once (cert_undistort).
template <bool FULL>
M3D_HD int undistort_pinhole_core(const CamDev& c, double u, double v, double& xo, double& yo) {
  const double x0 = (u - c.cx) * c.ifx;
  const double y0 = (v - c.cy) * c.ify;
  double x = x0, y = y0;
  // OpenCV: "if (icdist < 0) { x = x0; y = y0; break; }".  The sign bits of the five icdist
  // values are OR-ed on the integer pipe.
  int neg = 0;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const double r2 = x * x + y * y;
    double icdist = rcp(1.0 + ((c.k[4] * r2 + c.k[1]) * r2 + c.k[0]) * r2);
    if (FULL) icdist *= 1.0 + ((c.k[7] * r2 + c.k[6]) * r2 + c.k[5]) * r2;
#if defined(__CUDA_ARCH__)
    neg |= __double2hiint(icdist);
#else
    neg |= (icdist < 0.0) ? -1 : 0;
#endif
    const double x2 = x + x, y2 = y + y;
    const double xy2 = x2 * y;
    double dx = c.k[2] * xy2 + c.k[3] * (x2 * x + r2);
    double dy = c.k[3] * xy2 + c.k[2] * (y2 * y + r2);
    if (FULL) {
      dx += c.k[8] * r2 + c.k[9] * r2 * r2;
      dy += c.k[10] * r2 + c.k[11] * r2 * r2;
    }
    x = (x0 - dx) * icdist;
    y = (y0 - dy) * icdist;
  }
  xo = x;
  yo = y;
  return neg;
}

// undistort_pinhole_core for G cameras of one point, written iteration by iteration ACROSS the cameras: the G
// dependency chains (r2 -> polynomial -> reciprocal -> update, ~25 dependent float64 operations per iteration)
// are adjacent in program order, which is what lets the hardware overlap them within one thread.  Per camera
// the operations and their order are those of undistort_pinhole_core: identical results.
template <bool FULL, int G>
M3D_HD void undistort_pinhole_core_group(const CamDev* c, const double* u, const double* v, double* xo, double* yo,
                                         int* neg) {
  double x0[G], y0[G], x[G], y[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    x0[g] = (u[g] - c[g].cx) * c[g].ifx;
    y0[g] = (v[g] - c[g].cy) * c[g].ify;
    x[g] = x0[g], y[g] = y0[g];
    neg[g] = 0;
  }
#pragma unroll
  for (int j = 0; j < 5; ++j) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const double r2 = x[g] * x[g] + y[g] * y[g];
      double icdist = rcp(1.0 + ((c[g].k[4] * r2 + c[g].k[1]) * r2 + c[g].k[0]) * r2);
      if (FULL) icdist *= 1.0 + ((c[g].k[7] * r2 + c[g].k[6]) * r2 + c[g].k[5]) * r2;
#if defined(__CUDA_ARCH__)
      neg[g] |= __double2hiint(icdist);
#else
      neg[g] |= (icdist < 0.0) ? -1 : 0;
#endif
      const double x2 = x[g] + x[g], y2 = y[g] + y[g];
      const double xy2 = x2 * y[g];
      double dx = c[g].k[2] * xy2 + c[g].k[3] * (x2 * x[g] + r2);
      double dy = c[g].k[3] * xy2 + c[g].k[2] * (y2 * y[g] + r2);
      if (FULL) {
        dx += c[g].k[8] * r2 + c[g].k[9] * r2 * r2;
        dy += c[g].k[10] * r2 + c[g].k[11] * r2 * r2;
      }
      x[g] = (x0[g] - dx) * icdist;
      y[g] = (y0[g] - dy) * icdist;
    }
  }
#pragma unroll
  for (int g = 0; g < G; ++g) xo[g] = x[g], yo[g] = y[g];
}

// undistort_pinhole through undistort_pinhole_core: the replay of a flagged point in cert_undistort
template <bool FULL>
M3D_HD void undistort_pinhole_replay(const CamDev& c, double u, double v, double& xo, double& yo) {
  double x, y;
  const int neg = undistort_pinhole_core<FULL>(c, u, v, x, y);
  if (neg < 0) undistort_pinhole_exact<FULL>(c, u, v, x, y);  // rare (k1 << 0 at image corners)
  xo = x;
  yo = y;
}

// One camera, one point (every kernel but the straight-line form of cert_undistort).
template <bool FULL>
M3D_HD void undistort_pinhole(const CamDev& c, double u, double v, double& xo, double& yo) {
  const double x0 = (u - c.cx) * c.ifx;
  const double y0 = (v - c.cy) * c.ify;
  double x = x0, y = y0;
  // OpenCV: "if (icdist < 0) { x = x0; y = y0; break; }".  The sign bits of the five icdist
  // values are OR-ed on the integer pipe; a set bit (icdist < 0, or -0 / negative NaN, which
  // the literal transcription handles identically to this loop) replays the point through it.
  int neg = 0;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const double r2 = x * x + y * y;
    double icdist = rcp(1.0 + ((c.k[4] * r2 + c.k[1]) * r2 + c.k[0]) * r2);
    if (FULL) icdist *= 1.0 + ((c.k[7] * r2 + c.k[6]) * r2 + c.k[5]) * r2;
#if defined(__CUDA_ARCH__)
    neg |= __double2hiint(icdist);
#else
    neg |= (icdist < 0.0) ? -1 : 0;
#endif
    const double x2 = x + x, y2 = y + y;
    const double xy2 = x2 * y;
    double dx = c.k[2] * xy2 + c.k[3] * (x2 * x + r2);
    double dy = c.k[3] * xy2 + c.k[2] * (y2 * y + r2);
    if (FULL) {
      dx += c.k[8] * r2 + c.k[9] * r2 * r2;
      dy += c.k[10] * r2 + c.k[11] * r2 * r2;
    }
    x = (x0 - dx) * icdist;
    y = (y0 - dy) * icdist;
  }
  if (neg < 0) undistort_pinhole_exact<FULL>(c, u, v, x, y);  // rare (k1 << 0 at image corners)
  xo = x;
  yo = y;
}

// North-star-tolerance variant of undistort_pinhole (opt-in, M3D_UNDISTORT_FAST): the first three
// of OpenCV's five fixed-point iterations run in float32 (FMA pipe, twice the rate of the fp64
// pipe and issued beside it), the last two in float64 from the exact (x0, y0).  The iteration
// contracts (factor <= ~0.25 inside the image for |k1| <= 0.25), so the float32 rounding of the
// third iterate (~1e-7 relative) reaches the result damped by two more steps: measured <= 2e-8 in
// normalised units = 3e-5 px (tests/test_gpu_parity.py asserts the north-star tolerances 1e-4 rel /
// 0.01 mm / 1e-3 px against the reference goldens).  Plain 5-coefficient model only; the icdist < 0
// bail-out replays the point through the literal float64 transcription exactly like the strict path.
template <int NF32>
M3D_HD void undistort_pinhole_fast(const CamDev& c, double u, double v, double& xo, double& yo) {
  const double x0 = (u - c.cx) * c.ifx;
  const double y0 = (v - c.cy) * c.ify;
  const float x0f = (float)x0, y0f = (float)y0;
  float xf = x0f, yf = y0f;
  int neg = 0;
#pragma unroll
  for (int j = 0; j < NF32; ++j) {
    const float r2 = xf * xf + yf * yf;
    const float den = 1.0f + ((c.kf[4] * r2 + c.kf[1]) * r2 + c.kf[0]) * r2;
#if defined(__CUDA_ARCH__)
    float icd;  // MUFU.RCP (1 ulp): the contraction of the iteration absorbs it
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(icd) : "f"(den));
    neg |= __float_as_int(icd);
#else
    const float icd = 1.0f / den;
    neg |= (icd < 0.0f) ? -1 : 0;
#endif
    const float x2 = xf + xf, y2 = yf + yf;
    const float xy2 = x2 * yf;
    const float dx = c.kf[2] * xy2 + c.kf[3] * (x2 * xf + r2);
    const float dy = c.kf[3] * xy2 + c.kf[2] * (y2 * yf + r2);
    xf = (x0f - dx) * icd;
    yf = (y0f - dy) * icd;
  }
  double x = (double)xf, y = (double)yf;
#pragma unroll
  for (int j = 0; j < 5 - NF32; ++j) {
    const double r2 = x * x + y * y;
    const double icdist = rcp(1.0 + ((c.k[4] * r2 + c.k[1]) * r2 + c.k[0]) * r2);
#if defined(__CUDA_ARCH__)
    neg |= __double2hiint(icdist);
#else
    neg |= (icdist < 0.0) ? -1 : 0;
#endif
    const double x2 = x + x, y2 = y + y;
    const double xy2 = x2 * y;
    const double dx = c.k[2] * xy2 + c.k[3] * (x2 * x + r2);
    const double dy = c.k[3] * xy2 + c.k[2] * (y2 * y + r2);
    x = (x0 - dx) * icdist;
    y = (y0 - dy) * icdist;
  }
  // bail-out, or a float32 overflow on a finite input (|x0| > 1e6: far outside any image)
  if (neg < 0 || (!(x == x) && u == u && v == v)) undistort_pinhole_exact<false>(c, u, v, x, y);
  xo = x;
  yo = y;
}

// cv2.fisheye.undistortPoints(pts, K, D): Newton on theta, <= 10 iterations, stop when
// |fix| < 1e-8; non-converged or sign-flipped -> (-1e6, -1e6) (OpenCV >= 4.5).
M3D_HD void undistort_fisheye(const CamDev& c, double u, double v, double& xo, double& yo) {
  const double pw1 = (v - c.cy) / c.fy;
  double pw0 = (u - c.cx) / c.fx;
  const double alpha = c.skew / c.fx;
  if (alpha != 0.0) pw0 -= alpha * pw1;
  const double half_pi = 1.5707963267948966;
  double theta_d = sqrt(pw0 * pw0 + pw1 * pw1);
  // std::min(std::max(-pi/2, theta_d), pi/2): NaN collapses to -pi/2
  theta_d = (-half_pi < theta_d) ? theta_d : -half_pi;
  theta_d = (theta_d < half_pi) ? theta_d : half_pi;
  bool converged = false;
  double theta = theta_d, scale = 0.0;
  if (fabs(theta_d) > 1e-8) {
    for (int j = 0; j < 10; ++j) {
      const double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t6 * t2;
      const double k0 = c.k[0] * t2, k1 = c.k[1] * t4, k2 = c.k[2] * t6, k3 = c.k[3] * t8;
      const double fix = (theta * (1.0 + k0 + k1 + k2 + k3) - theta_d) /
                         (1.0 + 3.0 * k0 + 5.0 * k1 + 7.0 * k2 + 9.0 * k3);
      theta -= fix;
      if (fabs(fix) < 1e-8) {
        converged = true;
        break;
      }
    }
    scale = tan(theta) / theta_d;
  } else {
    converged = true;
  }
  const bool flipped = (theta_d < 0.0 && theta > 0.0) || (theta_d > 0.0 && theta < 0.0);
  if (converged && !flipped) {
    xo = pw0 * scale;
    yo = pw1 * scale;
  } else {
    xo = -1000000.0;
    yo = -1000000.0;
  }
}

// cv2.omnidir.undistortPoints(pts, K, D, xi, R = I) (Mei model; restated from the
// published opencv_contrib ccalib algorithm, parity unpinned).
M3D_HD void undistort_omnidir(const CamDev& c, double u, double v, double& xo, double& yo) {
  const double k1 = c.k[0], k2 = c.k[1], p1 = c.k[2], p2 = c.k[3], xi = c.xi;
  const double ppx = (u * c.fy - c.cx * c.fy - c.skew * (v - c.cy)) / (c.fx * c.fy);
  const double ppy = (v - c.cy) / c.fy;
  double pux = ppx, puy = ppy;
  for (int j = 0; j < 20; ++j) {
    const double r2 = pux * pux + puy * puy;
    const double r4 = r2 * r2;
    const double den = 1.0 + k1 * r2 + k2 * r4;
    pux = (ppx - 2.0 * p1 * pux * puy - p2 * (r2 + 2.0 * pux * pux)) / den;
    puy = (ppy - 2.0 * p2 * pux * puy - p1 * (r2 + 2.0 * puy * puy)) / den;
  }
  const double r2 = pux * pux + puy * puy;
  const double a = r2 + 1.0;
  const double b = 2.0 * xi * r2;
  const double cc = r2 * xi * xi - 1.0;
  const double Zs = (-b + sqrt(b * b - 4.0 * a * cc)) / (2.0 * a);
  const double X0 = pux * (Zs + xi), X1 = puy * (Zs + xi), X2 = Zs;
  const double nrm = sqrt(X0 * X0 + X1 * X1 + X2 * X2);
  const double s0 = X0 / nrm, s1 = X1 / nrm, s2 = X2 / nrm;
  xo = s0 / s2;
  yo = s1 / s2;
}

template <bool FULL, bool PINHOLE_ONLY>
M3D_HD void undistort_point(const CamDev& c, double u, double v, double& x, double& y) {
  if (PINHOLE_ONLY || c.model == PINHOLE) {
    undistort_pinhole<FULL>(c, u, v, x, y);
  } else if (c.model == FISHEYE) {
    undistort_fisheye(c, u, v, x, y);
  } else {
    undistort_omnidir(c, u, v, x, y);
  }
}

// ---------------------------------------------------------------------------------------
// projection: 3D world point -> pixel
// ---------------------------------------------------------------------------------------

M3D_HD void to_camera(const CamDev& c, double X, double Y, double Z, double& xc, double& yc,
                      double& zc) {
  xc = c.R[0] * X + c.R[1] * Y + c.R[2] * Z + c.t[0];
  yc = c.R[3] * X + c.R[4] * Y + c.R[5] * Z + c.t[1];
  zc = c.R[6] * X + c.R[7] * Y + c.R[8] * Z + c.t[2];
}

// normalised (x, y) -> pixel with the pinhole distortion model (second half of
// cv2.projectPoints)
template <bool FULL>
M3D_HD void distort_pinhole(const CamDev& c, double x, double y, double& u, double& v) {
  const double r2 = x * x + y * y;
  const double r4 = r2 * r2;
  const double r6 = r4 * r2;
  const double a1 = 2.0 * x * y;
  const double a2 = r2 + 2.0 * x * x;
  const double a3 = r2 + 2.0 * y * y;
  double cdist = 1.0 + c.k[0] * r2 + c.k[1] * r4 + c.k[4] * r6;
  double xd, yd;
  if (FULL) {
    const double icdist2 = rcp(1.0 + c.k[5] * r2 + c.k[6] * r4 + c.k[7] * r6);
    xd = x * cdist * icdist2 + c.k[2] * a1 + c.k[3] * a2 + c.k[8] * r2 + c.k[9] * r4;
    yd = y * cdist * icdist2 + c.k[2] * a3 + c.k[3] * a1 + c.k[10] * r2 + c.k[11] * r4;
  } else {
    xd = x * cdist + c.k[2] * a1 + c.k[3] * a2;
    yd = y * cdist + c.k[2] * a3 + c.k[3] * a1;
  }
  u = xd * c.fx + c.cx;
  v = yd * c.fy + c.cy;
}

// cv2.projectPoints: 1/z with z == 0 -> 1; points behind the camera are not rejected.
template <bool FULL>
M3D_HD void project_pinhole(const CamDev& c, double X, double Y, double Z, double& u, double& v) {
  double xc, yc, zc;
  to_camera(c, X, Y, Z, xc, yc, zc);
  const double iz = (zc != 0.0) ? rcp(zc) : 1.0;
  distort_pinhole<FULL>(c, xc * iz, yc * iz, u, v);
}

M3D_HD void distort_fisheye_cam(const CamDev& c, double xc, double yc, double zc, double& u,
                                double& v) {
  const double x = xc / zc, y = yc / zc;
  const double r2 = x * x + y * y;
  const double r = sqrt(r2);
  const double theta = atan(r);
  const double t2 = theta * theta, t3 = t2 * theta, t5 = t3 * t2, t7 = t5 * t2, t9 = t7 * t2;
  const double theta_d = theta + c.k[0] * t3 + c.k[1] * t5 + c.k[2] * t7 + c.k[3] * t9;
  const double inv_r = r > 1e-8 ? 1.0 / r : 1.0;
  double cdist = r > 1e-8 ? theta_d * inv_r : 1.0;
  if (r != r) cdist = r;  // NaN propagates
  const double xd1 = x * cdist, xd2 = y * cdist;
  const double alpha = c.skew / c.fx;
  u = c.fx * (xd1 + alpha * xd2) + c.cx;
  v = c.fy * xd2 + c.cy;
}

M3D_HD void distort_omnidir_cam(const CamDev& c, double xc, double yc, double zc, double& u,
                                double& v) {
  const double k1 = c.k[0], k2 = c.k[1], p1 = c.k[2], p2 = c.k[3];
  const double nrm = sqrt(xc * xc + yc * yc + zc * zc);
  const double s0 = xc / nrm, s1 = yc / nrm, s2 = zc / nrm;
  const double xu = s0 / (s2 + c.xi), yu = s1 / (s2 + c.xi);
  const double r2 = xu * xu + yu * yu, r4 = r2 * r2;
  const double rad = 1.0 + k1 * r2 + k2 * r4;
  const double xd = xu * rad + 2.0 * p1 * xu * yu + p2 * (r2 + 2.0 * xu * xu);
  const double yd = yu * rad + p1 * (r2 + 2.0 * yu * yu) + 2.0 * p2 * xu * yu;
  u = c.fx * xd + c.skew * yd + c.cx;
  v = c.fy * yd + c.cy;
}

template <bool FULL, bool PINHOLE_ONLY>
M3D_HD void project_point(const CamDev& c, double X, double Y, double Z, double& u, double& v) {
  if (PINHOLE_ONLY || c.model == PINHOLE) {
    project_pinhole<FULL>(c, X, Y, Z, u, v);
  } else {
    double xc, yc, zc;
    to_camera(c, X, Y, Z, xc, yc, zc);
    if (c.model == FISHEYE) {
      distort_fisheye_cam(c, xc, yc, zc, u, v);
    } else {
      distort_omnidir_cam(c, xc, yc, zc, u, v);
    }
  }
}

// Camera.distort_points: projectPoints of (x, y, 1) with identity extrinsics.
template <bool FULL>
M3D_HD void distort_point(const CamDev& c, double x, double y, double& u, double& v) {
  if (c.model == PINHOLE) {
    distort_pinhole<FULL>(c, x, y, u, v);
  } else if (c.model == FISHEYE) {
    distort_fisheye_cam(c, x, y, 1.0, u, v);
  } else {
    distort_omnidir_cam(c, x, y, 1.0, u, v);
  }
}

// residual norm exactly as np.linalg.norm(axis=2): sqrt(ex*ex + ey*ey)
M3D_HD double residual_norm(double ex, double ey) { return sqrt_fast(ex * ex + ey * ey); }

// ---------------------------------------------------------------------------------------
// DLT normal equations.  For camera rows a1 = x*M[2]-M[0], a2 = y*M[2]-M[1] (M = [R|t])
// the 4x4 Gram matrix G = sum a a^T is kept in block form
//   G = [ H  g ]   H: 3x3 symmetric (h[0..5] = xx xy xz yy yz zz), g: 3-vector, w: scalar
//       [ g' w ]
// ---------------------------------------------------------------------------------------
struct Gram {
  double h[6];
  double g[3];
  double w;
};

M3D_HD void gram_zero(Gram& G) {
#pragma unroll
  for (int i = 0; i < 6; ++i) G.h[i] = 0.0;
  G.g[0] = G.g[1] = G.g[2] = 0.0;
  G.w = 0.0;
}

M3D_HD void gram_add_row(Gram& G, double a0, double a1, double a2, double a3) {
  G.h[0] += a0 * a0;
  G.h[1] += a0 * a1;
  G.h[2] += a0 * a2;
  G.h[3] += a1 * a1;
  G.h[4] += a1 * a2;
  G.h[5] += a2 * a2;
  G.g[0] += a0 * a3;
  G.g[1] += a1 * a3;
  G.g[2] += a2 * a3;
  G.w += a3 * a3;
}

// G += (s a)(a)^T for a signed row (sa = sign * a)
M3D_HD void gram_add_row(Gram& G, double s0, double s1, double s2, double s3, double a0, double a1,
                         double a2, double a3) {
  G.h[0] += s0 * a0;
  G.h[1] += s0 * a1;
  G.h[2] += s0 * a2;
  G.h[3] += s1 * a1;
  G.h[4] += s1 * a2;
  G.h[5] += s2 * a2;
  G.g[0] += s0 * a3;
  G.g[1] += s1 * a3;
  G.g[2] += s2 * a3;
  G.w += s3 * a3;
}

// rows of camera c for the undistorted observation (x, y)   (cameras.py:27-28)
M3D_HD void gram_add_camera(Gram& G, const CamDev& c, double x, double y) {
  gram_add_row(G, x * c.R[6] - c.R[0], x * c.R[7] - c.R[1], x * c.R[8] - c.R[2],
               x * c.t[2] - c.t[0]);
  gram_add_row(G, y * c.R[6] - c.R[3], y * c.R[7] - c.R[4], y * c.R[8] - c.R[5],
               y * c.t[2] - c.t[1]);
}

// signed accumulation (sign = +1 / -1): toggling a camera in or out of a running Gram sum
M3D_HD void gram_acc_camera(Gram& G, const CamDev& c, double x, double y, double sign) {
  const double a0 = x * c.R[6] - c.R[0], a1 = x * c.R[7] - c.R[1], a2 = x * c.R[8] - c.R[2],
               a3 = x * c.t[2] - c.t[0];
  const double b0 = y * c.R[6] - c.R[3], b1 = y * c.R[7] - c.R[4], b2 = y * c.R[8] - c.R[5],
               b3 = y * c.t[2] - c.t[1];
  gram_add_row(G, sign * a0, sign * a1, sign * a2, sign * a3, a0, a1, a2, a3);
  gram_add_row(G, sign * b0, sign * b1, sign * b2, sign * b3, b0, b1, b2, b3);
}

M3D_HD void gram_add(Gram& G, const Gram& B) {
#pragma unroll
  for (int i = 0; i < 6; ++i) G.h[i] += B.h[i];
  G.g[0] += B.g[0];
  G.g[1] += B.g[1];
  G.g[2] += B.g[2];
  G.w += B.w;
}

// Fallback eigen-solver: cyclic Jacobi on the full 4x4 Gram matrix, returns the
// eigenvector of the smallest eigenvalue dehomogenised.  Only reached for degenerate
// geometry (H singular, point at infinity, lambda_1 not separated).
M3D_HD_NOINLINE void dlt_solve_jacobi(const Gram& G, double& X, double& Y, double& Z) {
  double A[4][4] = {{G.h[0], G.h[1], G.h[2], G.g[0]},
                    {G.h[1], G.h[3], G.h[4], G.g[1]},
                    {G.h[2], G.h[4], G.h[5], G.g[2]},
                    {G.g[0], G.g[1], G.g[2], G.w}};
  double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < 4; ++i) {
      diag += A[i][i] * A[i][i];
      for (int j = i + 1; j < 4; ++j) off += A[i][j] * A[i][j];
    }
    if (!(off > 1e-60 * diag)) break;
    for (int p = 0; p < 3; ++p) {
      for (int q = p + 1; q < 4; ++q) {
        const double apq = A[p][q];
        if (apq == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
        for (int r = 0; r < 4; ++r) {
          const double arp = A[r][p], arq = A[r][q];
          A[r][p] = cs * arp - sn * arq;
          A[r][q] = sn * arp + cs * arq;
        }
        for (int r = 0; r < 4; ++r) {
          const double apr = A[p][r], aqr = A[q][r];
          A[p][r] = cs * apr - sn * aqr;
          A[q][r] = sn * apr + cs * aqr;
        }
        for (int r = 0; r < 4; ++r) {
          const double vrp = V[r][p], vrq = V[r][q];
          V[r][p] = cs * vrp - sn * vrq;
          V[r][q] = sn * vrp + cs * vrq;
        }
      }
    }
  }
  int m = 0;
  for (int i = 1; i < 4; ++i)
    if (A[i][i] < A[m][m]) m = i;
  X = V[0][m] / V[3][m];
  Y = V[1][m] / V[3][m];
  Z = V[2][m] / V[3][m];
}

// Smallest eigenvector of G, dehomogenised (= vh[-1][:3]/vh[-1][3] of the SVD of the DLT
// matrix, cameras.py:29-31).  With v = (X, 1) the eigen-equations are
//   (H - lam I) X = -g ,   f(lam) := w - lam + g.X(lam) = 0 ,
// f is concave and decreasing on [0, lam_min(H)), and its only root there is lam_min(G)
// (Cauchy interlacing).  Newton on f from lam = 0 is Rayleigh-quotient iteration in this
// block form: lam += f / (1 + |X|^2).  The 3x3 systems are solved with the adjugate;
// H (entries O(k)) is well conditioned, the O(|t|^2) dynamic range of G stays in g and w,
// which is why this is more accurate than an eigen-decomposition of G itself
// (measured: <= 2e-11 px in mean reprojection error vs extended precision; LAPACK's SVD
// of the DLT matrix: ~1e-10 px).
M3D_HD void dlt_solve(const Gram& G, double& X, double& Y, double& Z) {
  double lam = 0.0, lam_pd = 0.0;  // lam_pd: last shift at which H - lam I was positive definite
  double x0 = 0.0, x1 = 0.0, x2 = 0.0;
  bool ok = false;
  const double tr = G.h[0] + G.h[3] + G.h[5];
  const double itr2 = rcp(tr * tr);
#pragma unroll 1
  for (int it = 0; it < 16; ++it) {
    const double a = G.h[0] - lam, b = G.h[1], c = G.h[2], d = G.h[3] - lam, e = G.h[4],
                 f = G.h[5] - lam;
    const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
    const double c11 = a * f - c * c, c12 = b * c - a * e, c22 = a * d - b * b;
    const double det = a * c00 + b * c01 + c * c02;
    // positive definite (Sylvester) <=> lam < lam_min(H): we are on the right branch.  The
    // first Newton step (from the left of the root of a concave f) overshoots; with gross
    // outliers in the subset it can jump past the pole lam_min(H): halve it back.
    if (!(a > 0.0 && c22 > 0.0 && det > 0.0)) {
      if (it == 0 || !(lam > lam_pd)) break;  // H itself is not positive definite: degenerate
      lam = 0.5 * (lam + lam_pd);
      continue;
    }
    lam_pd = lam;
    // p = adj(H - lam I) g = -det X.  The Newton step f / (1 + |X|^2) needs one reciprocal
    // in this form (and no 1/det on the dependency chain):
    //   f = w - lam - (g.p)/det ,  1 + |X|^2 = (det^2 + |p|^2)/det^2
    const double p0 = c00 * G.g[0] + c01 * G.g[1] + c02 * G.g[2];
    const double p1 = c01 * G.g[0] + c11 * G.g[1] + c12 * G.g[2];
    const double p2 = c02 * G.g[0] + c12 * G.g[1] + c22 * G.g[2];
    const double q = G.g[0] * p0 + G.g[1] * p1 + G.g[2] * p2;
    const double pp = p0 * p0 + p1 * p1 + p2 * p2;
    const double num = (G.w - lam) * det - q;
    const double iden = rcp(det * det + pp);
    const double dl = det * num * iden;
    // lam_min(H - lam I) >= det / tr^2 ; a step below 1e-7 of that changes X by < 1e-14 |X|
    // beyond the first-order correction applied here.  A step at the rounding-noise level of
    // its own numerator (near-parallel rays: lam_min(H) -> 0) cannot be improved in fp64
    // either: finish there as well.
    const double mu_lb = det * itr2;
    const double noise = 4e-16 * (fabs(G.w - lam) * det + fabs(q)) * det * iden;
    if (fabs(dl) <= 1e-7 * mu_lb || fabs(dl) <= noise) {
      const double idet = rcp(det);
      x0 = -p0 * idet;
      x1 = -p1 * idet;
      x2 = -p2 * idet;
      // first-order update X(lam + dl) = X + dl (H - lam I)^-1 X
      const double y0 = (c00 * x0 + c01 * x1 + c02 * x2) * idet;
      const double y1 = (c01 * x0 + c11 * x1 + c12 * x2) * idet;
      const double y2 = (c02 * x0 + c12 * x1 + c22 * x2) * idet;
      x0 += dl * y0;
      x1 += dl * y1;
      x2 += dl * y2;
      ok = (x0 == x0);
      break;
    }
    lam += dl;
    if (!(lam >= 0.0)) break;  // also catches NaN
  }
  if (ok) {
    X = x0;
    Y = x1;
    Z = x2;
  } else {
    const Gram tmp = G;  // the copy (not G) is what lives in local memory for the call
    dlt_solve_jacobi(tmp, X, Y, Z);
  }
}

#if defined(__CUDACC__)
// dlt_solve for a whole warp (device only): the same Newton / Rayleigh-quotient iteration per
// lane, but warp-convergent and without per-lane control flow.  A lane that has stopped keeps
// lam fixed and therefore recomputes identical values (so "converged" can be read off the last
// trip); lanes without work (active = false) ride along on whatever G they were given.  The rare
// lanes that leave the fast path (step past the pole lam_min(H), degenerate H, no convergence in
// 8 trips: ~0.2 % of the solves) are redone by the general dlt_solve, which halves such steps
// back and ends in the Jacobi eigen-solver.
__device__ __forceinline__ void dlt_solve_warp(const Gram& G, bool active, double& X, double& Y, double& Z) {
  constexpr unsigned FULLM = 0xffffffffu;
  double lam = 0.0;
  bool done = !active;
  const double tr = G.h[0] + G.h[3] + G.h[5];
  const double tol = 1e-7 * rcp_coarse(tr * tr);  // lam_min(H - lam I) >= det / tr^2
  double c00, c01, c02, c11, c12, c22, det, p0, p1, p2, dl;
  bool pd, crit;
  const double b = G.h[1], c = G.h[2], e = G.h[4];
#pragma unroll 1
  for (int it = 0; it < 8; ++it) {
    const double a = G.h[0] - lam, d = G.h[3] - lam, f = G.h[5] - lam;
    c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
    c11 = a * f - c * c, c12 = b * c - a * e, c22 = a * d - b * b;
    det = a * c00 + b * c01 + c * c02;
    pd = a > 0.0 && c22 > 0.0 && det > 0.0;  // Sylvester: lam < lam_min(H)
    p0 = c00 * G.g[0] + c01 * G.g[1] + c02 * G.g[2];
    p1 = c01 * G.g[0] + c11 * G.g[1] + c12 * G.g[2];
    p2 = c02 * G.g[0] + c12 * G.g[1] + c22 * G.g[2];
    const double q = G.g[0] * p0 + G.g[1] * p1 + G.g[2] * p2;
    const double pp = p0 * p0 + p1 * p1 + p2 * p2;
    const double wl = G.w - lam;
    const double num = wl * det - q;
    // 2^-40 is plenty for the step: its error is second order in what is left of the iteration
    const double iden = rcp_coarse(det * det + pp);
    dl = det * num * iden;
    crit = fabs(dl) <= tol * det;
    if (it >= 2 && !crit) {
      // a step at the rounding-noise level of its own numerator (near-parallel rays) cannot be
      // improved in fp64 either
      const double noise = 4e-16 * (fabs(wl) * det + fabs(q)) * det * iden;
      crit = fabs(dl) <= noise;
    }
    const double ln = lam + dl;
    const bool go = !done && pd && !crit && ln >= 0.0;
    if (go) lam = ln;
    done = !go;
    if (!__any_sync(FULLM, go)) break;
  }
  bool ok = active && pd && crit;
  X = Y = Z = qnan();
  if (ok) {
    const double idet = rcp(det);
    double x0 = -p0 * idet, x1 = -p1 * idet, x2 = -p2 * idet;
    // first-order update X(lam + dl) = X + dl (H - lam I)^-1 X
    const double y0 = (c00 * x0 + c01 * x1 + c02 * x2) * idet;
    const double y1 = (c01 * x0 + c11 * x1 + c12 * x2) * idet;
    const double y2 = (c02 * x0 + c12 * x1 + c22 * x2) * idet;
    x0 += dl * y0;
    x1 += dl * y1;
    x2 += dl * y2;
    ok = (x0 == x0);
    X = x0;
    Y = x1;
    Z = x2;
  }
  if (active && !ok) dlt_solve(G, X, Y, Z);
}
#endif  // __CUDACC__

// Rank-deficient normal matrix (rays that do not pin the point down: two coincident cameras, one
// ray seen twice): np.linalg.pinv zeroes the singular values below 1e-15 sigma_max and returns the
// MINIMUM-NORM solution.  Same result through the eigen-decomposition of H = A^T A (cyclic Jacobi):
// X = - sum_{lambda_i > 1e-12 lambda_max} v_i (v_i . g) / lambda_i  (the squared cut-off is floored at
// the rounding noise of forming H).
M3D_HD_NOINLINE void ls_solve_pinv(const Gram& G, double& X, double& Y, double& Z) {
  double A[3][3] = {{G.h[0], G.h[1], G.h[2]}, {G.h[1], G.h[3], G.h[4]}, {G.h[2], G.h[4], G.h[5]}};
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 30; ++sweep) {
    const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    const double dg = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
    if (!(off > 1e-60 * dg)) break;
    for (int p = 0; p < 2; ++p) {
      for (int q = p + 1; q < 3; ++q) {
        const double apq = A[p][q];
        if (apq == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
        for (int r = 0; r < 3; ++r) {
          const double arp = A[r][p], arq = A[r][q];
          A[r][p] = cs * arp - sn * arq;
          A[r][q] = sn * arp + cs * arq;
        }
        for (int r = 0; r < 3; ++r) {
          const double apr = A[p][r], aqr = A[q][r];
          A[p][r] = cs * apr - sn * aqr;
          A[q][r] = sn * apr + cs * aqr;
        }
        for (int r = 0; r < 3; ++r) {
          const double vrp = V[r][p], vrq = V[r][q];
          V[r][p] = cs * vrp - sn * vrq;
          V[r][q] = sn * vrp + cs * vrq;
        }
      }
    }
  }
  double lmax = A[0][0];
  if (A[1][1] > lmax) lmax = A[1][1];
  if (A[2][2] > lmax) lmax = A[2][2];
  X = Y = Z = 0.0;
  for (int i = 0; i < 3; ++i) {
    if (A[i][i] > 1e-12 * lmax) {
      const double c = -(V[0][i] * G.g[0] + V[1][i] * G.g[1] + V[2][i] * G.g[2]) / A[i][i];
      X += c * V[0][i];
      Y += c * V[1][i];
      Z += c * V[2][i];
    }
  }
  if (!(lmax > 0.0)) X = Y = Z = qnan();
}

// Inhomogeneous least squares X = -pinv(A[:, :3]) A[:, 3] (multicam_toolbox.py:476-484): -H^-1 g
// through the adjugate for a well-conditioned H, the pseudo-inverse (ls_solve_pinv) when H is
// numerically rank-deficient.
M3D_HD void ls_solve(const Gram& G, double& X, double& Y, double& Z) {
  const double a = G.h[0], b = G.h[1], c = G.h[2], d = G.h[3], e = G.h[4], f = G.h[5];
  const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
  const double c11 = a * f - c * c, c12 = b * c - a * e, c22 = a * d - b * b;
  const double det = a * c00 + b * c01 + c * c02;
  const double tr = a + d + f;
  // det = l1 l2 l3 <= (tr/3)^3: a ratio below 1e-11 means lambda_min / lambda_max < ~1e-10
  if (!(det > 1e-11 * tr * tr * tr)) {
    const Gram tmp = G;
    ls_solve_pinv(tmp, X, Y, Z);
    return;
  }
  const double idet = 1.0 / det;
  X = -(c00 * G.g[0] + c01 * G.g[1] + c02 * G.g[2]) * idet;
  Y = -(c01 * G.g[0] + c11 * G.g[1] + c12 * G.g[2]) * idet;
  Z = -(c02 * G.g[0] + c12 * G.g[1] + c22 * G.g[2]) * idet;
}

// ---------------------------------------------------------------------------------------
// camera-subset enumeration of triangulate_possible (cameras.py:689-692)
// ---------------------------------------------------------------------------------------

// Camera mask of enumeration step s for valid-camera mask vmask (k = popcount):
// the j-th valid camera (ascending) is included iff bit (k-1-j) of s is 0.
M3D_HD uint32_t subset_mask(uint32_t vmask, int k, uint32_t s) {
  uint32_t m = 0;
  int j = 0;
  for (uint32_t rest = vmask; rest; rest &= rest - 1, ++j) {
    const uint32_t low = rest & (0u - rest);
    if (!((s >> (k - 1 - j)) & 1u)) m |= low;
  }
  return m;
}

}  // namespace m3d
