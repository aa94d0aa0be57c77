// m3d_point.cuh — per-point building blocks shared by the kernels (m3d_kernels.cu) and
// the test-only host harness: mean reprojection error with pruning, and the evaluation of
// one camera subset of triangulate_possible (cameras.py:689-713).
#pragma once
#include "m3d_math.cuh"

namespace m3d {

#if defined(__CUDA_ARCH__)
#define M3D_POPC(x) __popc(x)
#define M3D_CLZ(x) __clz((int)(x))
#else
#define M3D_POPC(x) __builtin_popcount(x)
#define M3D_CLZ(x) __builtin_clz(x)
#endif

M3D_HD double pos_inf() {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(0x7ff0000000000000LL);
#else
  return INFINITY;
#endif
}

// Mean reprojection error of (X,Y,Z) against the RAW pixels raw[2c], raw[2c+1] of the
// cameras in cmask: residual norms that are NaN drop out of sum and count, fewer than two
// contributing cameras -> NaN (cameras.py:769-775).  If the running sum exceeds `limit`
// the subset cannot reach mean < limit / popc(cmask) any more and +inf is returned
// (sum / count >= sum / popc(cmask)); pass limit = +inf for the exact value.
template <bool FULL, bool PO>
M3D_HD double mean_reproj_error(const RigDev& rig, const double* raw, uint32_t cmask, double X,
                                double Y, double Z, double limit) {
  double sum = 0.0;
  int m = 0;
  for (uint32_t rest = cmask; rest; rest &= rest - 1) {
#if defined(__CUDA_ARCH__)
    const int c = __ffs(rest) - 1;
#else
    const int c = __builtin_ctz(rest);
#endif
    double u, v;
    project_point<FULL, PO>(rig.cam[c], X, Y, Z, u, v);
    const double e = residual_norm(raw[2 * c] - u, raw[2 * c + 1] - v);
    if (e == e) {
      sum += e;
      ++m;
    }
    if (sum > limit) return pos_inf();
  }
  return (m >= 2) ? sum / (double)m : qnan();
}

// One step of the subset search: triangulate from the usable cameras of cmask (Gram
// contributions gc[c]) and score against the raw pixels of all cameras of cmask.
// Returns the mean reprojection error, NaN when it is undefined, +inf when pruned
// against `T` (the subset provably has err >= T).
template <bool FULL, bool PO>
M3D_HD double eval_subset(const RigDev& rig, const double* raw, const Gram* gc, uint32_t cmask,
                          uint32_t umask, double T, double& X, double& Y, double& Z) {
  const uint32_t ucm = cmask & umask;
  if (M3D_POPC(ucm) < 2) {
    X = Y = Z = qnan();
    return qnan();
  }
  Gram G;
  gram_zero(G);
  for (uint32_t rest = ucm; rest; rest &= rest - 1) {
#if defined(__CUDA_ARCH__)
    const int c = __ffs(rest) - 1;
#else
    const int c = __builtin_ctz(rest);
#endif
    gram_add(G, gc[c]);
  }
  dlt_solve(G, X, Y, Z);
  // prune only when clearly above T * |S| (the tiny slack keeps the test conservative
  // against the rounding of the product and of the final division)
  const double limit = T * (double)M3D_POPC(cmask) * (1.0 + 1e-12);
  return mean_reproj_error<FULL, PO>(rig, raw, cmask, X, Y, Z, limit);
}

}  // namespace m3d
