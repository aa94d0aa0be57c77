// m3d_cert.h — host-side construction of the PAIR CERTIFICATE tables of the subset search
// (k_ransac_cert, m3d_ransac_cert.cuh).  Plain C++ (no CUDA) so that the CPU test tier can
// compile and check it (tests/host_harness.cpp).
//
// What is certified (proof: DESIGN.md section 3.2c).  For two cameras a, b of a subset S that the
// reference would accept at threshold T (mean reprojection error < T, cameras.py:701-713) the
// residuals satisfy  e_a(X) + e_b(X) < rho := T |S|  at the subset's DLT point X.  For a camera
// whose distortion map D is STRONGLY MONOTONE on the whole normalised plane with modulus mu
// (<D(n) - D(m), n - m> >= mu |n - m|^2) the true normalised projection n_c(X) then lies within
//   r_c = (e_c + delta_c) / (mu_c fmin_c)
// of the centre xh_c (the 5-iteration undistorted observation; delta_c = |raw_c - K D(xh_c)| is
// evaluated, not assumed), and the epipolar form  n_b~^T E n_a~ = 0  of the two true projections
// gives
//   |xh_b~^T E xh_a~| <= r_b |(E xh_a~)_xy| + r_a |(E^T xh_b~)_xy| + r_a r_b |E_2x2| .
// A pair that violates this inequality for every split e_a + e_b < rho cannot be part of an
// accepted subset: every subset containing it is rejected WITHOUT a solve.  The test never
// rejects a subset the reference accepts; when it is inconclusive the subset is solved as before.
//
// This file certifies mu_c > 0 for the plain pinhole model (k1 k2 p1 p2 k3, cv2.projectPoints,
// cameras.py:318-323) and builds the normalised essential matrices of all camera pairs.  A camera
// that cannot be certified (rational / thin-prism terms, fisheye / omnidir models, a radial
// polynomial that folds back such as k1 < 0 alone) gets inv_mf = 0: none of its pairs is ever used.
#pragma once
#include <cmath>
#include <cstring>

#include "m3d_math.cuh"

namespace m3d {

constexpr int CERT_MAX_PAIRS = M3D_MAXC * (M3D_MAXC - 1) / 2;  // 120

struct CertDev {
  int32_t n_cams;
  int32_t ok_mask;            // bit c: camera c is certified
  double inv_mf[M3D_MAXC];    // 1 / (mu_c * min(|fx|, |fy|)), 0 = not certified
  // pair (a < b) at index pair_index(a, b, C): row-major E (normalised to unit Frobenius norm) with
  // X_b^T E X_a = 0 for camera-frame coordinates of the same world point, then
  // gam = |E[0:2,0:2]|_2 * inv_mf[a] * inv_mf[b]
  double E[CERT_MAX_PAIRS][10];
  // float32 forms of the same tables for the pair test (m3d_ransac_cert.cuh cert_pairs): E entries rounded
  // to nearest (their rounding is part of the test's error budget), gam / 4 and inv_mf rounded UP
  float Ef[CERT_MAX_PAIRS][10];
  float inv_mf_f[M3D_MAXC];
};

// smallest float >= x * (1 + 1e-6) for x >= 0 (upper bounds that survive the conversion)
inline float cert_up_f32(double x) {
  if (!(x > 0.0)) return 0.0f;
  float f = (float)(x * (1.0 + 1e-6));
  while ((double)f < x) f = std::nextafter(f, INFINITY);
  return f;
}

M3D_HD int pair_index(int a, int b, int C) { return a * (2 * C - a - 1) / 2 + (b - a - 1); }

// Certified lower bound mu of the smallest eigenvalue of the (symmetric) Jacobian of
//   D(n) = n s(u) + tau(n),  u = |n|^2,  s(u) = 1 + k1 u + k2 u^2 + k3 u^3,
//   tau = (2 p1 x y + p2 (u + 2 x^2), p1 (u + 2 y^2) + 2 p2 x y)
// over the WHOLE plane, or 0 when no positive bound can be certified.
//   J = s I + 2 s' n n^T + J_tau : eigenvalues of the radial part are s(u) (tangential direction)
//   and q(u) = s + 2 u s' = 1 + 3 k1 u + 5 k2 u^2 + 7 k3 u^3 (radial direction); J_tau is symmetric
//   with |J_tau|_2 <= 6 (|p1| + |p2|) sqrt(u).  Hence lambda_min(J) >= h(u) := min(s, q) - 6 P sqrt(u).
// [0, U]: interval scan with Lipschitz bounds of s and q;  [U, inf): for k3 >= 0 and
// c2 = k2 + k3 U >= 0, c1 = k1 + U c2 >= 0 one has s(u) >= 1 + c1 u (same for q with the
// coefficients 3 k1, 5 k2, 7 k3), and 1 + c u - 6 P sqrt(u) is increasing once c >= 3 P / sqrt(U).
inline double certify_pinhole5(const CamDev& c) {
  for (int j = 5; j < 12; ++j)
    if (c.k[j] != 0.0) return 0.0;
  const double k1 = c.k[0], k2 = c.k[1], p1 = c.k[2], p2 = c.k[3], k3 = c.k[4];
  if (!(std::isfinite(k1) && std::isfinite(k2) && std::isfinite(k3) && std::isfinite(p1) && std::isfinite(p2)))
    return 0.0;
  const double P = std::fabs(p1) + std::fabs(p2);
  if (!(k3 >= 0.0)) return 0.0;
  const double U = 400.0;
  const double c2 = k2 + k3 * U, c1 = k1 + U * c2;
  const double c2q = 5.0 * k2 + 7.0 * k3 * U, c1q = 3.0 * k1 + U * c2q;
  if (!(c2 >= 0.0 && c1 >= 0.0 && c2q >= 0.0 && c1q >= 0.0)) return 0.0;
  const double cm = c1 < c1q ? c1 : c1q;
  if (!(cm >= 3.0 * P / std::sqrt(U))) return 0.0;
  double mu = 1.0 + cm * U - 6.0 * P * std::sqrt(U);
  double u0 = 0.0;
  while (u0 < U) {
    const double du = u0 < 4.0 ? 1e-3 : u0 * 2.5e-4;
    double u1 = u0 + du;
    if (u1 > U) u1 = U;
    const double s0 = 1.0 + u0 * (k1 + u0 * (k2 + u0 * k3));
    const double q0 = 1.0 + u0 * (3.0 * k1 + u0 * (5.0 * k2 + u0 * 7.0 * k3));
    const double Ls = std::fabs(k1) + 2.0 * std::fabs(k2) * u1 + 3.0 * std::fabs(k3) * u1 * u1;
    const double Lq = 3.0 * std::fabs(k1) + 10.0 * std::fabs(k2) * u1 + 21.0 * std::fabs(k3) * u1 * u1;
    const double a = s0 - Ls * (u1 - u0), b = q0 - Lq * (u1 - u0);
    const double lb = (a < b ? a : b) - 6.0 * P * std::sqrt(u1);
    if (lb < mu) mu = lb;
    u0 = u1;
  }
  mu *= 1.0 - 1e-9;  // rounding of the scan itself
  return (mu >= 0.05) ? mu : 0.0;
}

// spectral norm of a 2x2 matrix
inline double norm2x2(double a, double b, double c, double d) {
  const double f = a * a + b * b + c * c + d * d;
  const double det = a * d - b * c;
  const double disc = f * f - 4.0 * det * det;
  return std::sqrt(0.5 * (f + std::sqrt(disc > 0.0 ? disc : 0.0)));
}

inline void build_cert(const RigDev& rig, CertDev* cert) {
  std::memset(cert, 0, sizeof(CertDev));
  const int C = rig.n_cams;
  cert->n_cams = C;
  for (int c = 0; c < C; ++c) {
    const CamDev& cam = rig.cam[c];
    double mu = 0.0;
    if (cam.model == PINHOLE) mu = certify_pinhole5(cam);
    const double fmin = std::fabs(cam.fx) < std::fabs(cam.fy) ? std::fabs(cam.fx) : std::fabs(cam.fy);
    if (mu > 0.0 && std::isfinite(fmin) && fmin > 0.0) {
      cert->inv_mf[c] = (1.0 + 1e-12) / (mu * fmin);
      cert->inv_mf_f[c] = cert_up_f32(cert->inv_mf[c]);
      cert->ok_mask |= 1 << c;
    }
  }
  for (int a = 0; a < C; ++a) {
    for (int b = a + 1; b < C; ++b) {
      double* E = cert->E[pair_index(a, b, C)];
      const double* Ra = rig.cam[a].R;
      const double* Rb = rig.cam[b].R;
      const double* ta = rig.cam[a].t;
      const double* tb = rig.cam[b].t;
      // X_b = Rba X_a + tba,  Rba = Rb Ra^T,  tba = tb - Rba ta ;  E = [tba]x Rba
      double Rba[9], tba[3];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
          Rba[3 * i + j] = Rb[3 * i] * Ra[3 * j] + Rb[3 * i + 1] * Ra[3 * j + 1] + Rb[3 * i + 2] * Ra[3 * j + 2];
      for (int i = 0; i < 3; ++i)
        tba[i] = tb[i] - (Rba[3 * i] * ta[0] + Rba[3 * i + 1] * ta[1] + Rba[3 * i + 2] * ta[2]);
      const double tx[9] = {0, -tba[2], tba[1], tba[2], 0, -tba[0], -tba[1], tba[0], 0};
      double nrm = 0.0;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          E[3 * i + j] = tx[3 * i] * Rba[j] + tx[3 * i + 1] * Rba[3 + j] + tx[3 * i + 2] * Rba[6 + j];
          nrm += E[3 * i + j] * E[3 * i + j];
        }
      nrm = std::sqrt(nrm);
      if (!(nrm > 0.0) || !std::isfinite(nrm)) {  // coincident centres: the form vanishes, never "bad"
        for (int i = 0; i < 10; ++i) E[i] = 0.0;
        continue;
      }
      for (int i = 0; i < 9; ++i) E[i] /= nrm;
      E[9] = norm2x2(E[0], E[1], E[3], E[4]) * (1.0 + 1e-12) * cert->inv_mf[a] * cert->inv_mf[b];
      float* Ef = cert->Ef[pair_index(a, b, C)];
      for (int i = 0; i < 9; ++i) Ef[i] = (float)E[i];
      Ef[9] = cert_up_f32(0.25 * E[9]);
    }
  }
}

// cumb[i][j] = sum_{t <= j} C(i, t), i, j <= 16
struct CumBinom {
  uint32_t v[17][17];
};

inline CumBinom make_cumbinom() {
  CumBinom t;
  uint32_t c[17][17] = {};
  for (int i = 0; i <= 16; ++i) {
    c[i][0] = 1;
    for (int j = 1; j <= i; ++j) c[i][j] = c[i - 1][j - 1] + (j <= i - 1 ? c[i - 1][j] : 0);
  }
  for (int i = 0; i <= 16; ++i) {
    uint32_t acc = 0;
    for (int j = 0; j <= 16; ++j) {
      if (j <= i) acc += c[i][j];
      t.v[i][j] = acc;
    }
  }
  return t;
}

}  // namespace m3d
