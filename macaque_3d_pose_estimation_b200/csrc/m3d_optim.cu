// m3d_optim.cu — CameraGroup.optim_points / optim_points_jointlenfix on the GPU
// (reference: aniposelib/cameras.py:1116-1270; residual _error_fun_triangulation :1560-1620;
// caller step4_aniposefiltering.py:247-271, the default `optim = true` branch of the 3D stage).
//
// The reference minimises, over all 3D points of one animal P (F, J, 3) and the limb lengths L,
//   r1  rho(|score * (p2d - project_c(P[f,j]))|)           per camera, frame, joint, x / y  (NaN p2d skipped)
//   r2  scale_smooth * diff_n(P, axis = frames)             n-th temporal difference
//   r3  scale_length      * 100 (|P[f,a] - P[f,b]| - L_k) / L_k   strong limb constraints
//   r4  scale_length_weak * 100 (|P[f,a] - P[f,b]| - Lw_k) / Lw_k weak limb constraints
// with scipy.optimize.least_squares(method 'trf', 2-point finite-difference Jacobian, lsmr, ftol 1e-3).
// Here: the same residual vector (m3d_optim_residual, checked value by value against the executed
// reference), EXACT Jacobian blocks (forward-mode dual numbers through the camera model, closed
// forms for r2..r4), and Levenberg-Marquardt with a block-preconditioned conjugate-gradient inner
// solve that never forms J^T J:  every product is local — per point (2C x 3 block), per frame pair
// (differences), per (constraint, frame) — so all frames of the recording are processed at once.
// The solver starts from the reference's own x0 and is run to a tighter tolerance than the
// reference's ftol = 1e-3, so its final cost is <= the reference's (tests/test_optim.py).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/m3d.h"
#include "m3d_handle.h"
#include "m3d_math.cuh"

using namespace m3d;
typedef M3dDeviceGuard DeviceGuard;
static int fail(int code, const std::string& msg) { return m3d_fail(code, msg); }

namespace {

// ---- forward-mode dual numbers: value + gradient with respect to (X, Y, Z) -------------------------
struct Dual {
  double v, d0, d1, d2;
};
__host__ __device__ inline Dual mk(double v) { return {v, 0.0, 0.0, 0.0}; }
__host__ __device__ inline Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d0 + b.d0, a.d1 + b.d1, a.d2 + b.d2}; }
__host__ __device__ inline Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d0 - b.d0, a.d1 - b.d1, a.d2 - b.d2}; }
__host__ __device__ inline Dual operator*(Dual a, Dual b) {
  return {a.v * b.v, a.d0 * b.v + a.v * b.d0, a.d1 * b.v + a.v * b.d1, a.d2 * b.v + a.v * b.d2};
}
__host__ __device__ inline Dual operator+(Dual a, double b) { return {a.v + b, a.d0, a.d1, a.d2}; }
__host__ __device__ inline Dual operator+(double b, Dual a) { return {a.v + b, a.d0, a.d1, a.d2}; }
__host__ __device__ inline Dual operator*(Dual a, double b) { return {a.v * b, a.d0 * b, a.d1 * b, a.d2 * b}; }
__host__ __device__ inline Dual operator*(double b, Dual a) { return {a.v * b, a.d0 * b, a.d1 * b, a.d2 * b}; }
__host__ __device__ inline Dual operator/(Dual a, Dual b) {
  const double ib = 1.0 / b.v, q = a.v * ib;
  return {q, (a.d0 - q * b.d0) * ib, (a.d1 - q * b.d1) * ib, (a.d2 - q * b.d2) * ib};
}
__host__ __device__ inline Dual dsqrt(Dual a) {
  const double s = sqrt(a.v), h = 0.5 / s;
  return {s, a.d0 * h, a.d1 * h, a.d2 * h};
}
__host__ __device__ inline Dual datan(Dual a) {
  const double h = 1.0 / (1.0 + a.v * a.v);
  return {atan(a.v), a.d0 * h, a.d1 * h, a.d2 * h};
}

// cv2.projectPoints / cv2.fisheye.projectPoints / cv2.omnidir.projectPoints with derivatives
// (the arithmetic of m3d_math.cuh project_point, written on duals)
__device__ void project_dual(const CamDev& c, double X, double Y, double Z, Dual& u, Dual& v) {
  const Dual xc = {c.R[0] * X + c.R[1] * Y + c.R[2] * Z + c.t[0], c.R[0], c.R[1], c.R[2]};
  const Dual yc = {c.R[3] * X + c.R[4] * Y + c.R[5] * Z + c.t[1], c.R[3], c.R[4], c.R[5]};
  const Dual zc = {c.R[6] * X + c.R[7] * Y + c.R[8] * Z + c.t[2], c.R[6], c.R[7], c.R[8]};
  if (c.model == PINHOLE) {
    const Dual iz = (zc.v != 0.0) ? mk(1.0) / zc : mk(1.0);
    const Dual x = xc * iz, y = yc * iz;
    const Dual r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
    const Dual a1 = 2.0 * (x * y), a2 = r2 + 2.0 * (x * x), a3 = r2 + 2.0 * (y * y);
    Dual cd = 1.0 + c.k[0] * r2 + c.k[1] * r4 + c.k[4] * r6;
    const Dual den = 1.0 + c.k[5] * r2 + c.k[6] * r4 + c.k[7] * r6;
    cd = cd / den;
    const Dual xd = x * cd + c.k[2] * a1 + c.k[3] * a2 + c.k[8] * r2 + c.k[9] * r4;
    const Dual yd = y * cd + c.k[2] * a3 + c.k[3] * a1 + c.k[10] * r2 + c.k[11] * r4;
    u = xd * c.fx + c.cx;
    v = yd * c.fy + c.cy;
  } else if (c.model == FISHEYE) {
    const Dual x = xc / zc, y = yc / zc;
    const Dual r = dsqrt(x * x + y * y);
    const Dual th = datan(r);
    const Dual t2 = th * th;
    const Dual thd = th * (1.0 + t2 * (c.k[0] + t2 * (c.k[1] + t2 * (c.k[2] + t2 * c.k[3]))));
    const Dual cd = (r.v > 1e-8) ? thd / r : mk(1.0);
    const Dual xd1 = x * cd, xd2 = y * cd;
    const double alpha = c.skew / c.fx;
    u = (xd1 + alpha * xd2) * c.fx + c.cx;
    v = xd2 * c.fy + c.cy;
  } else {
    const Dual nrm = dsqrt(xc * xc + yc * yc + zc * zc);
    const Dual s0 = xc / nrm, s1 = yc / nrm, s2 = zc / nrm;
    const Dual den = s2 + c.xi;
    const Dual xu = s0 / den, yu = s1 / den;
    const Dual r2 = xu * xu + yu * yu, r4 = r2 * r2;
    const Dual rad = 1.0 + c.k[0] * r2 + c.k[1] * r4;
    const Dual xd = xu * rad + 2.0 * c.k[2] * (xu * yu) + c.k[3] * (r2 + 2.0 * (xu * xu));
    const Dual yd = yu * rad + c.k[2] * (r2 + 2.0 * (yu * yu)) + 2.0 * c.k[3] * (xu * yu);
    u = xd * c.fx + c.skew * yd + c.cx;
    v = yd * c.fy + c.cy;
  }
}

enum { LOSS_LINEAR = 0, LOSS_SOFT_L1 = 1, LOSS_HUBER = 2 };

// rho(a) for a = |e| >= 0 and its derivative (cameras.py:1591-1597)
__device__ inline void loss_fn(int kind, double a, double rp, double& r, double& dr) {
  if (kind == LOSS_SOFT_L1) {
    const double s = sqrt(1.0 + a / rp);
    r = rp * 2.0 * (s - 1.0);
    dr = 1.0 / s;
  } else if (kind == LOSS_HUBER && a > rp) {
    const double s = sqrt(a / rp);
    r = rp * (2.0 * s - 1.0);
    dr = 1.0 / s;
  } else {
    r = a;
    dr = 1.0;
  }
}

struct Dims {
  int C, F, J, K, Kw, n_deriv;
  int64_t n_pts;  // F * J
};

// r1 and its Jacobian block.  Thread = (frame, joint).  res (C, F, J, 2): NaN where p2d is NaN.
// jac (n_pts, C, 2, 3): zero where masked.
template <bool WITH_JAC>
__global__ void __launch_bounds__(128)
k_opt_reproj(const __grid_constant__ RigDev rig, Dims dm, const double* __restrict__ p2d,
             const double* __restrict__ scores, const double* __restrict__ P, double rp, int loss,
             double* __restrict__ res, double* __restrict__ jac) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < dm.n_pts;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double X = P[3 * i], Y = P[3 * i + 1], Z = P[3 * i + 2];
    for (int c = 0; c < dm.C; ++c) {
      const int64_t o = (int64_t)c * dm.n_pts + i;
      const double px = p2d[2 * o], py = p2d[2 * o + 1];
      const double sc = scores ? scores[o] : 1.0;
      Dual u, v;
      project_dual(rig.cam[c], X, Y, Z, u, v);
      const double obs[2] = {px, py};
      const Dual pr[2] = {u, v};
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        double r = qnan(), j0 = 0.0, j1 = 0.0, j2 = 0.0;
        if (obs[a] == obs[a]) {
          const double e = (obs[a] - pr[a].v) * sc;
          double dr;
          loss_fn(loss, fabs(e), rp, r, dr);
          const double sg = (e > 0.0) ? 1.0 : ((e < 0.0) ? -1.0 : 0.0);
          const double w = -dr * sg * sc;  // d r / d proj
          j0 = w * pr[a].d0;
          j1 = w * pr[a].d1;
          j2 = w * pr[a].d2;
        }
        res[2 * o + a] = r;
        if (WITH_JAC) {
          double* jj = jac + ((i * dm.C + c) * 2 + a) * 3;
          jj[0] = j0;
          jj[1] = j1;
          jj[2] = j2;
        }
      }
    }
  }
}

// binomial coefficients of the n-th forward difference: sum_k coef[k] P[f + k]
__host__ __device__ inline double diff_coef(int n, int k) {
  double c = 1.0;
  for (int i = 0; i < k; ++i) c = c * (double)(n - i) / (double)(i + 1);
  return ((n - k) & 1) ? -c : c;
}

// r2: (F - n, J, 3)
__global__ void __launch_bounds__(256)
k_opt_smooth(Dims dm, const double* __restrict__ P, double s, double* __restrict__ res) {
  const int64_t row = (int64_t)dm.J * 3;
  const int64_t n = (int64_t)(dm.F - dm.n_deriv) * row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double acc = 0.0;
    // np.diff applies first differences n times: identical to the binomial sum up to rounding order;
    // evaluate it the same way (repeated differencing) for n <= 3
    if (dm.n_deriv == 1) acc = P[i + row] - P[i];
    else if (dm.n_deriv == 2) acc = (P[i + 2 * row] - P[i + row]) - (P[i + row] - P[i]);
    else if (dm.n_deriv == 3)
      acc = ((P[i + 3 * row] - P[i + 2 * row]) - (P[i + 2 * row] - P[i + row])) -
            ((P[i + 2 * row] - P[i + row]) - (P[i + row] - P[i]));
    else
      for (int k = 0; k <= dm.n_deriv; ++k) acc += diff_coef(dm.n_deriv, k) * P[i + k * row];
    res[i] = acc * s;
  }
}

// r3 / r4: (K + Kw, F).  cons (K + Kw, 2) joint pairs; L (K + Kw) expected lengths.
// gu (K + Kw, F, 3): d r / d P[f, a] (= - d r / d P[f, b]);  dL (K + Kw, F): d r / d L_k.
__global__ void __launch_bounds__(256)
k_opt_len(Dims dm, const double* __restrict__ P, const double* __restrict__ L, const int* __restrict__ cons,
          double sc_strong, double sc_weak, double* __restrict__ res, double* __restrict__ gu,
          double* __restrict__ dL) {
  const int KA = dm.K + dm.Kw;
  const int64_t n = (int64_t)KA * dm.F;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i / dm.F);
    const int f = (int)(i % dm.F);
    const int a = cons[2 * k], b = cons[2 * k + 1];
    const double* pa = P + ((int64_t)f * dm.J + a) * 3;
    const double* pb = P + ((int64_t)f * dm.J + b) * 3;
    const double dx = pa[0] - pb[0], dy = pa[1] - pb[1], dz = pa[2] - pb[2];
    const double len = sqrt(dx * dx + dy * dy + dz * dz);
    const double e = L[k];
    const double sc = (k < dm.K) ? sc_strong : sc_weak;
    res[i] = 100.0 * (len - e) / e * sc;
    if (gu) {
      const double g = (len > 0.0) ? 100.0 * sc / (e * len) : 0.0;
      gu[3 * i] = g * dx;
      gu[3 * i + 1] = g * dy;
      gu[3 * i + 2] = g * dz;
      dL[i] = -100.0 * sc * len / (e * e);
    }
  }
}

// ---- products with J and J^T ------------------------------------------------------------------------
// residual-space vector layout of the SOLVER: [r1 (n_pts, C, 2) | r2 ((F-n), J, 3) | r34 (KA, F)];
// parameter-space: [P (n_pts, 3) | L (KA)] (L part absent / ignored when the lengths are fixed)
struct Ptrs {
  const double* jac;   // (n_pts, C, 2, 3)
  const double* gu;    // (KA, F, 3)
  const double* dL;    // (KA, F)
  const int* cons;     // (KA, 2)
  const int* adj_off;  // (J + 1)
  const int* adj_k;    // constraint index
  const int* adj_sgn;  // +1: the joint is `a`, -1: `b`
};

// w = J v
__global__ void __launch_bounds__(128)
k_opt_Jv(Dims dm, Ptrs p, const double* __restrict__ v, const double* __restrict__ vL, double s,
         double* __restrict__ w1, double* __restrict__ w2, double* __restrict__ w34) {
  const int64_t n1 = dm.n_pts;
  const int64_t row = (int64_t)dm.J * 3;
  const int64_t n2 = (int64_t)(dm.F - dm.n_deriv) * row;
  const int KA = dm.K + dm.Kw;
  const int64_t n3 = (int64_t)KA * dm.F;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2 + n3;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (i < n1) {
      const double v0 = v[3 * i], v1 = v[3 * i + 1], v2 = v[3 * i + 2];
      const double* jj = p.jac + i * dm.C * 6;
      for (int r = 0; r < 2 * dm.C; ++r) w1[i * 2 * dm.C + r] = jj[3 * r] * v0 + jj[3 * r + 1] * v1 + jj[3 * r + 2] * v2;
    } else if (i < n1 + n2) {
      const int64_t q = i - n1;
      double acc = 0.0;
      for (int k = 0; k <= dm.n_deriv; ++k) acc += diff_coef(dm.n_deriv, k) * v[q + k * row];
      w2[q] = acc * s;
    } else {
      const int64_t q = i - n1 - n2;
      const int k = (int)(q / dm.F);
      const int f = (int)(q % dm.F);
      const int a = p.cons[2 * k], b = p.cons[2 * k + 1];
      const double* va = v + ((int64_t)f * dm.J + a) * 3;
      const double* vb = v + ((int64_t)f * dm.J + b) * 3;
      double acc = p.gu[3 * q] * (va[0] - vb[0]) + p.gu[3 * q + 1] * (va[1] - vb[1]) + p.gu[3 * q + 2] * (va[2] - vb[2]);
      if (vL) acc += p.dL[q] * vL[k];
      w34[q] = acc;
    }
  }
}

// g = J^T w (point part): thread = (frame, joint)
__global__ void __launch_bounds__(128)
k_opt_Jtw(Dims dm, Ptrs p, const double* __restrict__ w1, const double* __restrict__ w2,
          const double* __restrict__ w34, double s, double* __restrict__ g) {
  const int64_t row = (int64_t)dm.J * 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < dm.n_pts;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(i / dm.J), j = (int)(i % dm.J);
    double g0 = 0.0, g1 = 0.0, g2 = 0.0;
    const double* jj = p.jac + i * dm.C * 6;
    for (int r = 0; r < 2 * dm.C; ++r) {
      const double w = w1[i * 2 * dm.C + r];
      g0 += jj[3 * r] * w;
      g1 += jj[3 * r + 1] * w;
      g2 += jj[3 * r + 2] * w;
    }
    // smoothness: row (f - k) uses coef[k] on frame f
    for (int k = 0; k <= dm.n_deriv; ++k) {
      const int fr = f - k;
      if (fr >= 0 && fr < dm.F - dm.n_deriv) {
        const double cf = diff_coef(dm.n_deriv, k) * s;
        const int64_t q = (int64_t)fr * row + (int64_t)j * 3;
        g0 += cf * w2[q];
        g1 += cf * w2[q + 1];
        g2 += cf * w2[q + 2];
      }
    }
    for (int e = p.adj_off[j]; e < p.adj_off[j + 1]; ++e) {
      const int64_t q = (int64_t)p.adj_k[e] * dm.F + f;
      const double w = w34[q] * (double)p.adj_sgn[e];
      g0 += p.gu[3 * q] * w;
      g1 += p.gu[3 * q + 1] * w;
      g2 += p.gu[3 * q + 2] * w;
    }
    g[3 * i] = g0;
    g[3 * i + 1] = g1;
    g[3 * i + 2] = g2;
  }
}

// block reduction helper
__device__ inline double block_sum(double v) {
  __shared__ double sh[32];
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = (threadIdx.x < (blockDim.x + 31) / 32) ? sh[threadIdx.x] : 0.0;
    for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
  }
  __syncthreads();
  return t;  // valid in thread 0
}

// g_L[k] = sum_f dL[k, f] w34[k, f]   (one block per constraint)
__global__ void __launch_bounds__(256) k_opt_JtwL(Dims dm, Ptrs p, const double* __restrict__ w34, double* __restrict__ gL) {
  const int k = blockIdx.x;
  double acc = 0.0;
  for (int f = threadIdx.x; f < dm.F; f += blockDim.x) acc += p.dL[(int64_t)k * dm.F + f] * w34[(int64_t)k * dm.F + f];
  acc = block_sum(acc);
  if (threadIdx.x == 0) gL[k] = acc;
}

// 3x3 diagonal blocks of J^T J per point (6 unique entries) and the diagonal for L
__global__ void __launch_bounds__(128)
k_opt_blockdiag(Dims dm, Ptrs p, double s, double* __restrict__ B) {
  double c2 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < dm.n_pts;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(i / dm.J), j = (int)(i % dm.J);
    double b[6] = {0, 0, 0, 0, 0, 0};
    const double* jj = p.jac + i * dm.C * 6;
    for (int r = 0; r < 2 * dm.C; ++r) {
      const double a0 = jj[3 * r], a1 = jj[3 * r + 1], a2 = jj[3 * r + 2];
      b[0] += a0 * a0, b[1] += a0 * a1, b[2] += a0 * a2, b[3] += a1 * a1, b[4] += a1 * a2, b[5] += a2 * a2;
    }
    c2 = 0.0;
    for (int k = 0; k <= dm.n_deriv; ++k) {
      const int fr = f - k;
      if (fr >= 0 && fr < dm.F - dm.n_deriv) {
        const double cf = diff_coef(dm.n_deriv, k) * s;
        c2 += cf * cf;
      }
    }
    b[0] += c2, b[3] += c2, b[5] += c2;
    for (int e = p.adj_off[j]; e < p.adj_off[j + 1]; ++e) {
      const int64_t q = (int64_t)p.adj_k[e] * dm.F + f;
      const double a0 = p.gu[3 * q], a1 = p.gu[3 * q + 1], a2 = p.gu[3 * q + 2];
      b[0] += a0 * a0, b[1] += a0 * a1, b[2] += a0 * a2, b[3] += a1 * a1, b[4] += a1 * a2, b[5] += a2 * a2;
    }
    for (int t = 0; t < 6; ++t) B[6 * i + t] = b[t];
  }
}

__global__ void __launch_bounds__(256) k_opt_diagL(Dims dm, Ptrs p, double* __restrict__ BL) {
  const int k = blockIdx.x;
  double acc = 0.0;
  for (int f = threadIdx.x; f < dm.F; f += blockDim.x) {
    const double d = p.dL[(int64_t)k * dm.F + f];
    acc += d * d;
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) BL[k] = acc;
}

// z = M^-1 r with M = blockdiag(B + lam diag(B));  also q = (A v) damping term helper
__global__ void __launch_bounds__(128)
k_opt_precond(int64_t n_pts, int nL, const double* __restrict__ B, const double* __restrict__ BL, double lam,
              const double* __restrict__ r, double* __restrict__ z) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pts + nL;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (i >= n_pts) {
      const int k = (int)(i - n_pts);
      const double d = BL[k] * (1.0 + lam);
      z[3 * n_pts + k] = (d > 0.0) ? r[3 * n_pts + k] / d : r[3 * n_pts + k];
      continue;
    }
    const double* b = B + 6 * i;
    const double a = b[0] * (1.0 + lam) + 1e-300, bb = b[1], c = b[2], d = b[3] * (1.0 + lam) + 1e-300, e = b[4],
                 f = b[5] * (1.0 + lam) + 1e-300;
    const double c00 = d * f - e * e, c01 = c * e - bb * f, c02 = bb * e - c * d;
    const double c11 = a * f - c * c, c12 = bb * c - a * e, c22 = a * d - bb * bb;
    const double det = a * c00 + bb * c01 + c * c02;
    const double r0 = r[3 * i], r1 = r[3 * i + 1], r2 = r[3 * i + 2];
    if (det > 0.0) {
      const double id = 1.0 / det;
      z[3 * i] = (c00 * r0 + c01 * r1 + c02 * r2) * id;
      z[3 * i + 1] = (c01 * r0 + c11 * r1 + c12 * r2) * id;
      z[3 * i + 2] = (c02 * r0 + c12 * r1 + c22 * r2) * id;
    } else {
      z[3 * i] = r0 / a;
      z[3 * i + 1] = r1 / d;
      z[3 * i + 2] = r2 / f;
    }
  }
}

// out += lam * diag(B) .* v   (Marquardt damping term of A v)
__global__ void __launch_bounds__(256)
k_opt_damp(int64_t n_pts, int nL, const double* __restrict__ B, const double* __restrict__ BL, double lam,
           const double* __restrict__ v, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 3 * n_pts + nL;
       i += (int64_t)gridDim.x * blockDim.x) {
    double d;
    if (i < 3 * n_pts) {
      const int64_t pt = i / 3;
      const int a = (int)(i % 3);
      d = B[6 * pt + (a == 0 ? 0 : (a == 1 ? 3 : 5))];
    } else {
      d = BL[i - 3 * n_pts];
    }
    out[i] += lam * d * v[i];
  }
}

// ---- vector kernels ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dot(const double* __restrict__ a, const double* __restrict__ b, int64_t n,
                                            double* __restrict__ out, int skip_nan) {
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double x = a[i] * b[i];
    if (!skip_nan || x == x) acc += x;
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) out[1 + blockIdx.x] = acc;  // per-block partials, summed in a fixed order below
}
// out[0] = sum of the nb per-block partials out[1 .. nb], always in the same order: the solver is
// deterministic from run to run (no floating-point atomics anywhere on its path)
__global__ void __launch_bounds__(256) k_dot_final(double* __restrict__ out, int nb) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) acc += out[1 + i];
  acc = block_sum(acc);
  if (threadIdx.x == 0) out[0] = acc;
}
// y = alpha x + beta y
__global__ void __launch_bounds__(256) k_axpby(double alpha, const double* __restrict__ x, double beta,
                                              double* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = alpha * x[i] + beta * y[i];
}
// solver-layout r1 (n_pts, C, 2) from the dense reference-layout residual (C, n_pts, 2), NaN -> 0
__global__ void __launch_bounds__(256) k_opt_pack_r1(Dims dm, const double* __restrict__ res, double* __restrict__ w1) {
  const int64_t n = dm.n_pts * dm.C * 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pt = i / (2 * dm.C);
    const int r = (int)(i % (2 * dm.C));
    const int c = r >> 1, a = r & 1;
    const double x = res[2 * ((int64_t)c * dm.n_pts + pt) + a];
    w1[i] = (x == x) ? x : 0.0;
  }
}
// dense reference-layout (C, n_pts, 2) from solver-layout (n_pts, C, 2), NaN where p2d is NaN
__global__ void __launch_bounds__(256) k_opt_unpack_r1(Dims dm, const double* __restrict__ w1,
                                                      const double* __restrict__ p2d, double* __restrict__ res) {
  const int64_t n = dm.n_pts * dm.C * 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pt = i / (2 * dm.C);
    const int r = (int)(i % (2 * dm.C));
    const int c = r >> 1, a = r & 1;
    const int64_t o = 2 * ((int64_t)c * dm.n_pts + pt) + a;
    res[o] = (p2d[o] == p2d[o]) ? w1[i] : qnan();
  }
}

inline int blocks_for(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}

// Everything one optimisation needs on the device.
struct Optim {
  const m3d_rig* rig;
  cudaStream_t st;
  Dims dm;
  int nL;             // free length parameters (0: fixed lengths)
  double s_smooth, sc_len, sc_weak, rp;
  int loss;
  const double* p2d;
  const double* scores;
  // device buffers
  double *P = nullptr, *L = nullptr;                       // parameters
  double *res1 = nullptr, *r1 = nullptr, *r2 = nullptr, *r34 = nullptr;
  double *jac = nullptr, *gu = nullptr, *dL = nullptr, *B = nullptr, *BL = nullptr;
  int *cons = nullptr, *adj_off = nullptr, *adj_k = nullptr, *adj_sgn = nullptr;
  double* scal = nullptr;                                  // scalar scratch
  std::vector<void*> owned;
  int64_t n1, n2, n3, np;

  template <class T>
  cudaError_t alloc(T** p, size_t n) {
    cudaError_t e = cudaMalloc((void**)p, sizeof(T) * (n ? n : 1));
    if (e == cudaSuccess) owned.push_back(*p);
    return e;
  }
  ~Optim() {
    for (void* p : owned) cudaFree(p);
  }
  Ptrs ptrs() const { return {jac, gu, dL, cons, adj_off, adj_k, adj_sgn}; }

  double dot(const double* a, const double* b, int64_t n, bool skip_nan = false) {
    const int nb = n > 0 ? blocks_for(n, 256) : 0;
    if (nb > 0) k_dot<<<nb, 256, 0, st>>>(a, b, n, scal, skip_nan ? 1 : 0);
    k_dot_final<<<1, 256, 0, st>>>(scal, nb);
    double h = 0.0;
    cudaMemcpyAsync(&h, scal, sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    return h;
  }
  // residuals (and Jacobian blocks) at (Px, Lx); returns the cost 0.5 |r|^2
  double evaluate(const double* Px, const double* Lx, bool with_jac) {
    const int g1 = blocks_for(dm.n_pts, 128);
    if (with_jac)
      k_opt_reproj<true><<<g1, 128, 0, st>>>(rig->dev, dm, p2d, scores, Px, rp, loss, res1, jac);
    else
      k_opt_reproj<false><<<g1, 128, 0, st>>>(rig->dev, dm, p2d, scores, Px, rp, loss, res1, nullptr);
    k_opt_pack_r1<<<blocks_for(n1, 256), 256, 0, st>>>(dm, res1, r1);
    if (n2 > 0) k_opt_smooth<<<blocks_for(n2, 256), 256, 0, st>>>(dm, Px, s_smooth, r2);
    if (n3 > 0)
      k_opt_len<<<blocks_for(n3, 256), 256, 0, st>>>(dm, Px, Lx, cons, sc_len, sc_weak, r34, with_jac ? gu : nullptr, dL);
    double c = dot(r1, r1, n1);
    if (n2 > 0) c += dot(r2, r2, n2);
    if (n3 > 0) c += dot(r34, r34, n3);
    return 0.5 * c;
  }
};

}  // namespace

extern "C" {

// see include/m3d.h
int m3d_optim_points(const m3d_rig* rig, const double* p2d_dev, const double* scores_dev, int32_t F, int32_t J,
                     const int32_t* constraints, int32_t K, const int32_t* constraints_weak, int32_t Kw,
                     double scale_smooth, double scale_length, double scale_length_weak,
                     double reproj_error_threshold, int32_t loss, int32_t n_deriv_smooth, int32_t fix_lengths,
                     double ftol, int32_t max_iter, int32_t mode, double* params_dev, double* out_dev,
                     double* info_host, void* stream) {
  if (!rig) return fail(M3D_ERR_INVALID, "m3d_optim_points: rig is NULL");
  if (F < 1 || J < 1 || K < 0 || Kw < 0 || n_deriv_smooth < 1 || n_deriv_smooth > 8)
    return fail(M3D_ERR_INVALID, "m3d_optim_points: bad dimensions");
  if (!p2d_dev || !params_dev) return fail(M3D_ERR_INVALID, "m3d_optim_points: NULL buffer");
  if (loss < LOSS_LINEAR || loss > LOSS_HUBER) return fail(M3D_ERR_INVALID, "m3d_optim_points: unknown loss");
  if ((K > 0 && !constraints) || (Kw > 0 && !constraints_weak))
    return fail(M3D_ERR_INVALID, "m3d_optim_points: NULL constraint list");
  DeviceGuard guard(rig->device);
  Optim o;
  o.rig = rig;
  o.st = (cudaStream_t)stream;
  o.dm.C = rig->dev.n_cams;
  o.dm.F = F;
  o.dm.J = J;
  o.dm.K = K;
  o.dm.Kw = Kw;
  o.dm.n_deriv = n_deriv_smooth;
  o.dm.n_pts = (int64_t)F * J;
  const int KA = K + Kw;
  o.nL = fix_lengths ? 0 : KA;
  o.s_smooth = scale_smooth;
  o.sc_len = scale_length;
  o.sc_weak = scale_length_weak;
  o.rp = reproj_error_threshold;
  o.loss = loss;
  o.p2d = p2d_dev;
  o.scores = scores_dev;
  o.n1 = o.dm.n_pts * o.dm.C * 2;
  o.n2 = F > n_deriv_smooth ? (int64_t)(F - n_deriv_smooth) * J * 3 : 0;
  o.n3 = (int64_t)KA * F;
  o.np = 3 * o.dm.n_pts + o.nL;
  const int64_t n3p = 3 * o.dm.n_pts;
  // constraint tables
  std::vector<int> cons(2 * (KA > 0 ? KA : 1), 0), adj_off(J + 1, 0), adj_k, adj_sgn;
  for (int k = 0; k < KA; ++k) {
    const int32_t* src = (k < K) ? constraints + 2 * k : constraints_weak + 2 * (k - K);
    if (src[0] < 0 || src[0] >= J || src[1] < 0 || src[1] >= J)
      return fail(M3D_ERR_INVALID, "m3d_optim_points: constraint joint index out of range");
    cons[2 * k] = src[0];
    cons[2 * k + 1] = src[1];
  }
  for (int j = 0; j < J; ++j) {
    adj_off[j] = (int)adj_k.size();
    for (int k = 0; k < KA; ++k) {
      if (cons[2 * k] == j) adj_k.push_back(k), adj_sgn.push_back(1);
      if (cons[2 * k + 1] == j) adj_k.push_back(k), adj_sgn.push_back(-1);
    }
  }
  adj_off[J] = (int)adj_k.size();
#define OPT_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess) return fail(M3D_ERR_CUDA, std::string("m3d_optim_points: ") + cudaGetErrorString(e__)); \
  } while (0)
  OPT_CUDA(o.alloc(&o.res1, (size_t)o.n1));
  OPT_CUDA(o.alloc(&o.r1, (size_t)o.n1));
  OPT_CUDA(o.alloc(&o.r2, (size_t)o.n2));
  OPT_CUDA(o.alloc(&o.r34, (size_t)o.n3));
  OPT_CUDA(o.alloc(&o.jac, (size_t)o.n1 * 3));
  OPT_CUDA(o.alloc(&o.gu, (size_t)o.n3 * 3));
  OPT_CUDA(o.alloc(&o.dL, (size_t)o.n3));
  OPT_CUDA(o.alloc(&o.B, (size_t)o.dm.n_pts * 6));
  OPT_CUDA(o.alloc(&o.BL, (size_t)KA));
  OPT_CUDA(o.alloc(&o.cons, cons.size()));
  OPT_CUDA(o.alloc(&o.adj_off, adj_off.size()));
  OPT_CUDA(o.alloc(&o.adj_k, adj_k.size()));
  OPT_CUDA(o.alloc(&o.adj_sgn, adj_sgn.size()));
  OPT_CUDA(o.alloc(&o.scal, 148 * 16 + 4));
  OPT_CUDA(cudaMemcpyAsync(o.cons, cons.data(), sizeof(int) * cons.size(), cudaMemcpyHostToDevice, o.st));
  OPT_CUDA(cudaMemcpyAsync(o.adj_off, adj_off.data(), sizeof(int) * adj_off.size(), cudaMemcpyHostToDevice, o.st));
  if (!adj_k.empty()) {
    OPT_CUDA(cudaMemcpyAsync(o.adj_k, adj_k.data(), sizeof(int) * adj_k.size(), cudaMemcpyHostToDevice, o.st));
    OPT_CUDA(cudaMemcpyAsync(o.adj_sgn, adj_sgn.data(), sizeof(int) * adj_sgn.size(), cudaMemcpyHostToDevice, o.st));
  }
  OPT_CUDA(cudaStreamSynchronize(o.st));  // the host vectors go out of use
  double* P = params_dev;                 // [P | L]; with fixed lengths L is read-only
  double* Lp = params_dev + n3p;
  const Dims dm = o.dm;

  // ---- mode 1: residual vector in the reference's dense layout; mode 2: J v -----------------------
  if (mode == 1 || mode == 2) {
    if (!out_dev) return fail(M3D_ERR_INVALID, "m3d_optim_points: out_dev is NULL");
    const double cost = o.evaluate(P, Lp, mode == 2);
    if (mode == 1) {
      // [r1 (C, F, J, 2) with NaN where p2d is NaN | r2 | r3 | r4]
      OPT_CUDA(cudaMemcpyAsync(out_dev, o.res1, sizeof(double) * o.n1, cudaMemcpyDeviceToDevice, o.st));
      if (o.n2) OPT_CUDA(cudaMemcpyAsync(out_dev + o.n1, o.r2, sizeof(double) * o.n2, cudaMemcpyDeviceToDevice, o.st));
      if (o.n3) OPT_CUDA(cudaMemcpyAsync(out_dev + o.n1 + o.n2, o.r34, sizeof(double) * o.n3, cudaMemcpyDeviceToDevice, o.st));
    } else {
      // out_dev holds v (np values) on entry and J v (dense layout) on return
      double* v = nullptr;
      OPT_CUDA(o.alloc(&v, (size_t)(n3p + KA)));
      OPT_CUDA(cudaMemsetAsync(v, 0, sizeof(double) * (n3p + KA), o.st));
      OPT_CUDA(cudaMemcpyAsync(v, out_dev, sizeof(double) * o.np, cudaMemcpyDeviceToDevice, o.st));
      double* w = nullptr;
      OPT_CUDA(o.alloc(&w, (size_t)(o.n1 + o.n2 + o.n3)));
      k_opt_Jv<<<blocks_for(dm.n_pts + o.n2 + o.n3, 128), 128, 0, o.st>>>(dm, o.ptrs(), v, o.nL ? v + n3p : nullptr,
                                                                        o.s_smooth, w, w + o.n1, w + o.n1 + o.n2);
      k_opt_unpack_r1<<<blocks_for(o.n1, 256), 256, 0, o.st>>>(dm, w, p2d_dev, out_dev);
      if (o.n2 + o.n3)
        OPT_CUDA(cudaMemcpyAsync(out_dev + o.n1, w + o.n1, sizeof(double) * (o.n2 + o.n3), cudaMemcpyDeviceToDevice, o.st));
    }
    OPT_CUDA(cudaStreamSynchronize(o.st));
    if (info_host) info_host[0] = cost;
    return M3D_OK;
  }

  // ---- mode 0: Levenberg-Marquardt, preconditioned CG on (J^T J + lam D) d = -J^T r ----------------
  double *g = nullptr, *d = nullptr, *rr = nullptr, *z = nullptr, *pp = nullptr, *Ap = nullptr, *w = nullptr,
         *xt = nullptr;
  const int64_t npad = n3p + KA;
  OPT_CUDA(o.alloc(&g, (size_t)npad));
  OPT_CUDA(o.alloc(&d, (size_t)npad));
  OPT_CUDA(o.alloc(&rr, (size_t)npad));
  OPT_CUDA(o.alloc(&z, (size_t)npad));
  OPT_CUDA(o.alloc(&pp, (size_t)npad));
  OPT_CUDA(o.alloc(&Ap, (size_t)npad));
  OPT_CUDA(o.alloc(&xt, (size_t)npad));
  OPT_CUDA(o.alloc(&w, (size_t)(o.n1 + o.n2 + o.n3)));
  OPT_CUDA(cudaMemcpyAsync(xt, P, sizeof(double) * npad, cudaMemcpyDeviceToDevice, o.st));
  const int64_t np = o.np;
  auto apply_A = [&](const double* v, double* out, double lam) {
    k_opt_Jv<<<blocks_for(dm.n_pts + o.n2 + o.n3, 128), 128, 0, o.st>>>(dm, o.ptrs(), v, o.nL ? v + n3p : nullptr,
                                                                      o.s_smooth, w, w + o.n1, w + o.n1 + o.n2);
    k_opt_Jtw<<<blocks_for(dm.n_pts, 128), 128, 0, o.st>>>(dm, o.ptrs(), w, w + o.n1, w + o.n1 + o.n2, o.s_smooth, out);
    if (o.nL) k_opt_JtwL<<<KA, 256, 0, o.st>>>(dm, o.ptrs(), w + o.n1 + o.n2, out + n3p);
    k_opt_damp<<<blocks_for(np, 256), 256, 0, o.st>>>(dm.n_pts, o.nL, o.B, o.BL, lam, v, out);
  };
  double cost = o.evaluate(P, Lp, true);
  const double cost0 = cost;
  double lam = 1e-3;
  int it = 0, n_cg_total = 0, n_eval = 1;
  int status = 0;
  for (; it < max_iter; ++it) {
    // gradient g = J^T r and the block diagonal of J^T J
    k_opt_Jtw<<<blocks_for(dm.n_pts, 128), 128, 0, o.st>>>(dm, o.ptrs(), o.r1, o.r2, o.r34, o.s_smooth, g);
    if (o.nL) k_opt_JtwL<<<KA, 256, 0, o.st>>>(dm, o.ptrs(), o.r34, g + n3p);
    k_opt_blockdiag<<<blocks_for(dm.n_pts, 128), 128, 0, o.st>>>(dm, o.ptrs(), o.s_smooth, o.B);
    if (o.nL) k_opt_diagL<<<KA, 256, 0, o.st>>>(dm, o.ptrs(), o.BL);
    const double gnorm2 = o.dot(g, g, np);
    if (!(gnorm2 > 0.0)) {
      status = 1;  // stationary (or NaN)
      break;
    }
    bool accepted = false;
    for (int tries = 0; tries < 12 && !accepted; ++tries) {
      // PCG: d = 0, rr = -g
      OPT_CUDA(cudaMemsetAsync(d, 0, sizeof(double) * npad, o.st));
      k_axpby<<<blocks_for(np, 256), 256, 0, o.st>>>(-1.0, g, 0.0, rr, np);
      k_opt_precond<<<blocks_for(dm.n_pts + o.nL, 128), 128, 0, o.st>>>(dm.n_pts, o.nL, o.B, o.BL, lam, rr, z);
      OPT_CUDA(cudaMemcpyAsync(pp, z, sizeof(double) * np, cudaMemcpyDeviceToDevice, o.st));
      double rz = o.dot(rr, z, np);
      const double rz0 = rz;
      for (int cg = 0; cg < 200 && rz > 1e-6 * rz0 && rz > 0.0; ++cg) {
        apply_A(pp, Ap, lam);
        const double pAp = o.dot(pp, Ap, np);
        if (!(pAp > 0.0)) break;
        const double alpha = rz / pAp;
        k_axpby<<<blocks_for(np, 256), 256, 0, o.st>>>(alpha, pp, 1.0, d, np);
        k_axpby<<<blocks_for(np, 256), 256, 0, o.st>>>(-alpha, Ap, 1.0, rr, np);
        k_opt_precond<<<blocks_for(dm.n_pts + o.nL, 128), 128, 0, o.st>>>(dm.n_pts, o.nL, o.B, o.BL, lam, rr, z);
        const double rz_new = o.dot(rr, z, np);
        k_axpby<<<blocks_for(np, 256), 256, 0, o.st>>>(1.0, z, rz_new / rz, pp, np);
        rz = rz_new;
        ++n_cg_total;
      }
      // trial point
      OPT_CUDA(cudaMemcpyAsync(xt, P, sizeof(double) * npad, cudaMemcpyDeviceToDevice, o.st));
      k_axpby<<<blocks_for(np, 256), 256, 0, o.st>>>(1.0, d, 1.0, xt, np);
      // keep the Jacobian blocks of the current point: evaluate the trial without them
      const double c_new = o.evaluate(xt, o.nL ? xt + n3p : Lp, false);
      ++n_eval;
      if (c_new < cost) {
        const double dF = cost - c_new;
        OPT_CUDA(cudaMemcpyAsync(P, xt, sizeof(double) * np, cudaMemcpyDeviceToDevice, o.st));
        const double prev = cost;
        cost = o.evaluate(P, Lp, true);  // residuals + Jacobian blocks at the accepted point
        ++n_eval;
        lam = lam / 3.0 > 1e-9 ? lam / 3.0 : 1e-9;
        accepted = true;
        if (dF < ftol * prev) status = 2;  // relative cost reduction below ftol
      } else {
        lam *= 4.0;
        if (lam > 1e12) {
          status = 3;
          break;
        }
      }
    }
    if (!accepted) {
      if (!status) status = 3;
      // restore the residuals of the current point (the last trial overwrote them)
      cost = o.evaluate(P, Lp, true);
      break;
    }
    if (status == 2) {
      ++it;
      break;
    }
  }
  OPT_CUDA(cudaStreamSynchronize(o.st));
  if (info_host) {
    info_host[0] = cost;
    info_host[1] = cost0;
    info_host[2] = (double)it;
    info_host[3] = (double)n_cg_total;
    info_host[4] = (double)n_eval;
    info_host[5] = (double)status;
  }
#undef OPT_CUDA
  return M3D_OK;
}

}  // extern "C"
