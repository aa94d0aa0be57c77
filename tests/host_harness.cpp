// tests/host_harness.cpp — TEST INFRASTRUCTURE ONLY (never loaded by the product).
//
// Compiles the per-point arithmetic of the CUDA library (csrc/m3d_math.cuh,
// csrc/m3d_point.cuh — the very same source the kernels use) with g++ so that the
// CPU-only test tier can compare it with the oracle when no GPU is present.  The RANSAC
// driver below replays the kernel's schedule (32 subsets per step, pass 1 / pass 2)
// sequentially.  Built by tests/conftest.py into tests/_build/.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>

#include "../macaque_3d_pose_estimation_b200/csrc/m3d_math.cuh"
#include "../macaque_3d_pose_estimation_b200/csrc/m3d_point.cuh"
#include "../macaque_3d_pose_estimation_b200/csrc/m3d_rig.h"
#include "../macaque_3d_pose_estimation_b200/csrc/m3d_cert.h"
#include "../macaque_3d_pose_estimation_b200/csrc/m3d_ransac_cert.cuh"

using namespace m3d;

static thread_local std::string g_err;

template <bool FULL, bool PO>
static void undistort_all(const RigDev& rig, const double* xy, int64_t N, double* out) {
  for (int c = 0; c < rig.n_cams; ++c)
    for (int64_t n = 0; n < N; ++n) {
      const int64_t i = (int64_t)c * N + n;
      undistort_point<FULL, PO>(rig.cam[c], xy[2 * i], xy[2 * i + 1], out[2 * i], out[2 * i + 1]);
    }
}

template <bool FULL, bool PO>
static void project_all(const RigDev& rig, const double* p3d, int64_t N, double* out) {
  for (int c = 0; c < rig.n_cams; ++c)
    for (int64_t n = 0; n < N; ++n) {
      const int64_t i = (int64_t)c * N + n;
      project_point<FULL, PO>(rig.cam[c], p3d[3 * n], p3d[3 * n + 1], p3d[3 * n + 2], out[2 * i],
                              out[2 * i + 1]);
    }
}

template <bool FULL, bool PO>
static void triangulate_all(const RigDev& rig, const double* xy, int64_t N, int undistort, double* p3d,
                            double* err) {
  const int C = rig.n_cams;
  for (int64_t n = 0; n < N; ++n) {
    Gram G;
    gram_zero(G);
    int cnt = 0;
    for (int c = 0; c < C; ++c) {
      const double px = xy[2 * ((int64_t)c * N + n)], py = xy[2 * ((int64_t)c * N + n) + 1];
      double x = px, y = py;
      if (undistort) undistort_point<FULL, PO>(rig.cam[c], px, py, x, y);
      if (x == x) {
        gram_add_camera(G, rig.cam[c], x, y);
        ++cnt;
      }
    }
    double X = qnan(), Y = qnan(), Z = qnan();
    if (cnt >= 2) dlt_solve(G, X, Y, Z);
    p3d[3 * n] = X;
    p3d[3 * n + 1] = Y;
    p3d[3 * n + 2] = Z;
    if (err) {
      double sum = 0.0;
      int m = 0;
      for (int c = 0; c < C; ++c) {
        const double px = xy[2 * ((int64_t)c * N + n)], py = xy[2 * ((int64_t)c * N + n) + 1];
        double u, v;
        project_point<FULL, PO>(rig.cam[c], X, Y, Z, u, v);
        const double e = residual_norm(px - u, py - v);
        if (e == e) {
          sum += e;
          ++m;
        }
      }
      err[n] = (m >= 2) ? sum / (double)m : qnan();
    }
  }
}

template <bool FULL, bool PO>
static void ransac_all(const RigDev& rig, const double* xy, int64_t N, int undistort, int min_cams,
                       double thr, double init_best, double* p3d, uint8_t* picked, double* xy_picked,
                       double* err_out, int32_t* subset_out, int32_t* neval_out) {
  const int C = rig.n_cams;
  const double T1 = thr < init_best ? thr : init_best;
  for (int64_t n = 0; n < N; ++n) {
    double raw[2 * M3D_MAXC];
    Gram gc[M3D_MAXC];
    uint32_t vmask = 0, umask = 0;
    Gram G;
    gram_zero(G);
    for (int c = 0; c < C; ++c) {
      const double px = xy[2 * ((int64_t)c * N + n)], py = xy[2 * ((int64_t)c * N + n) + 1];
      raw[2 * c] = px;
      raw[2 * c + 1] = py;
      gram_zero(gc[c]);
      if (px == px) {
        vmask |= 1u << c;
        double x = px, y = py;
        if (undistort) undistort_point<FULL, PO>(rig.cam[c], px, py, x, y);
        if (x == x) {
          umask |= 1u << c;
          gram_add_camera(gc[c], rig.cam[c], x, y);
          gram_add_camera(G, rig.cam[c], x, y);
        }
      }
    }
    const int k = __builtin_popcount(vmask);
    double best_err = init_best, bx = qnan(), by = qnan(), bz = qnan();
    int32_t best_s = -1, neval = 1;
    uint32_t best_mask = 0;
    bool done = false;
    if (__builtin_popcount(umask) >= 2) {
      double X, Y, Z;
      dlt_solve(G, X, Y, Z);
      const double e0 = mean_reproj_error<FULL, PO>(rig, raw, vmask, X, Y, Z, pos_inf());
      if (e0 < best_err) {
        best_err = e0;
        best_s = 0;
        best_mask = vmask;
        bx = X;
        by = Y;
        bz = Z;
        if (e0 < thr) done = true;
      }
    }
    if (k < 2 || k <= min_cams) done = true;
    if (!done) {
      const uint32_t n_sub = 1u << k;
      bool found = false;
      for (uint32_t s = 1; s < n_sub && !found; ++s) {
        const uint32_t cm = subset_mask(vmask, k, s);
        const int cnt = __builtin_popcount(cm);
        if (!(cnt >= min_cams || cnt == k)) continue;
        ++neval;
        double X, Y, Z;
        const double e = eval_subset<FULL, PO>(rig, raw, gc, cm, umask, T1, X, Y, Z);
        if (e < T1) {
          best_err = e;
          best_s = (int32_t)s;
          best_mask = cm;
          bx = X;
          by = Y;
          bz = Z;
          found = true;
        }
      }
      if (!found) {
        double rb = best_err;
        for (uint32_t base = 0; base < n_sub; base += 32) {
          double ce = pos_inf(), cx = 0, cy = 0, cz = 0;
          uint32_t cmk = 0;
          int32_t cs = -1;
          for (uint32_t s = base; s < base + 32 && s < n_sub; ++s) {
            if (s < 1) continue;
            const uint32_t cm = subset_mask(vmask, k, s);
            const int cnt = __builtin_popcount(cm);
            if (!(cnt >= min_cams || cnt == k)) continue;
            double X, Y, Z;
            double e = eval_subset<FULL, PO>(rig, raw, gc, cm, umask, rb, X, Y, Z);
            if (!(e < rb)) continue;
            if (e < ce) {
              ce = e;
              cs = (int32_t)s;
              cmk = cm;
              cx = X;
              cy = Y;
              cz = Z;
            }
          }
          if (ce < rb) {
            rb = ce;
            best_err = ce;
            best_s = cs;
            best_mask = cmk;
            bx = cx;
            by = cy;
            bz = cz;
          }
        }
      }
    }
    p3d[3 * n] = bx;
    p3d[3 * n + 1] = by;
    p3d[3 * n + 2] = bz;
    err_out[n] = (best_s >= 0) ? best_err : 0.0;
    if (subset_out) subset_out[n] = best_s;
    if (neval_out) neval_out[n] = neval;
    for (int c = 0; c < C; ++c) {
      const bool in = (best_mask >> c) & 1u;
      if (picked) picked[(int64_t)c * N + n] = in ? 1 : 0;
      if (xy_picked) {
        xy_picked[2 * ((int64_t)c * N + n)] = in ? raw[2 * c] : qnan();
        xy_picked[2 * ((int64_t)c * N + n) + 1] = in ? raw[2 * c + 1] : qnan();
      }
    }
  }
}

#define DISPATCH(rig, CALL)                                                      \
  do {                                                                           \
    const bool full__ = ((rig).flags & (RIG_HAS_RATIONAL | RIG_HAS_PRISM)) != 0; \
    const bool po__ = ((rig).flags & RIG_HAS_NONPINHOLE) == 0;                   \
    if (full__) {                                                                \
      if (po__) { CALL(true, true); } else { CALL(true, false); }                \
    } else {                                                                     \
      if (po__) { CALL(false, true); } else { CALL(false, false); }              \
    }                                                                            \
  } while (0)

extern "C" {

const char* hh_last_error() { return g_err.c_str(); }

int hh_extrinsics(const m3d_cam* cams, int C, double* M) {
  RigDev rig;
  g_err = build_rig(cams, C, &rig);
  if (!g_err.empty()) return -1;
  for (int c = 0; c < C; ++c) {
    double* o = M + 16 * c;
    for (int r = 0; r < 3; ++r) {
      for (int j = 0; j < 3; ++j) o[4 * r + j] = rig.cam[c].R[3 * r + j];
      o[4 * r + 3] = rig.cam[c].t[r];
    }
    o[12] = o[13] = o[14] = 0;
    o[15] = 1;
  }
  return 0;
}

int hh_undistort(const m3d_cam* cams, int C, const double* xy, int64_t N, double* out) {
  RigDev rig;
  g_err = build_rig(cams, C, &rig);
  if (!g_err.empty()) return -1;
#define CALL(F, P) undistort_all<F, P>(rig, xy, N, out)
  DISPATCH(rig, CALL);
#undef CALL
  return 0;
}

int hh_project(const m3d_cam* cams, int C, const double* p3d, int64_t N, double* out) {
  RigDev rig;
  g_err = build_rig(cams, C, &rig);
  if (!g_err.empty()) return -1;
#define CALL(F, P) project_all<F, P>(rig, p3d, N, out)
  DISPATCH(rig, CALL);
#undef CALL
  return 0;
}

int hh_triangulate_error(const m3d_cam* cams, int C, const double* xy, int64_t N, int undistort,
                         double* p3d, double* err) {
  RigDev rig;
  g_err = build_rig(cams, C, &rig);
  if (!g_err.empty()) return -1;
#define CALL(F, P) triangulate_all<F, P>(rig, xy, N, undistort, p3d, err)
  DISPATCH(rig, CALL);
#undef CALL
  return 0;
}

int hh_ransac(const m3d_cam* cams, int C, const double* xy, int64_t N, int undistort, int min_cams,
              double thr, double init_best, double* p3d, uint8_t* picked, double* xy_picked,
              double* err, int32_t* subset, int32_t* neval) {
  RigDev rig;
  g_err = build_rig(cams, C, &rig);
  if (!g_err.empty()) return -1;
#define CALL(F, P) \
  ransac_all<F, P>(rig, xy, N, undistort, min_cams, thr, init_best, p3d, picked, xy_picked, err, subset, neval)
  DISPATCH(rig, CALL);
#undef CALL
  return 0;
}

// pair-certificate tables of the pruned subset search (csrc/m3d_cert.h): inv_mf (C), E (pairs, 10)
int hh_cert(const m3d_cam* cams, int C, int32_t* ok_mask, double* inv_mf, double* E) {
  RigDev rig;
  std::string why = build_rig(cams, C, &rig);
  if (!why.empty()) {
    g_err = why;
    return -1;
  }
  static CertDev cert;
  build_cert(rig, &cert);
  *ok_mask = cert.ok_mask;
  for (int c = 0; c < C; ++c) inv_mf[c] = cert.inv_mf[c];
  const int np = C * (C - 1) / 2;
  for (int p = 0; p < np; ++p)
    for (int j = 0; j < 10; ++j) E[10 * p + j] = cert.E[p][j];
  return 0;
}

// the pruned subset search of k_ransac_cert, point by point (the same ransac_cert_point the kernel runs)
int hh_ransac_cert(const m3d_cam* cams, int C, const double* xy, int64_t N, int undistort, int min_cams,
                   double threshold, double init_best, int use_cert, double* p3d, uint8_t* picked,
                   double* xy_picked, double* err, int32_t* subset, int32_t* neval, int32_t* n_solved) {
  RigDev rig;
  std::string why = build_rig(cams, C, &rig);
  if (!why.empty()) {
    g_err = why;
    return -1;
  }
  if (rig.flags & (RIG_HAS_RATIONAL | RIG_HAS_PRISM)) {
    g_err = "hh_ransac_cert: rational / thin-prism rigs are not handled by the pruned search";
    return -1;
  }
  static CertDev cert;
  build_cert(rig, &cert);
  static const CumBinom cumb = make_cumbinom();
  const bool po = (rig.flags & RIG_HAS_NONPINHOLE) == 0;
  for (int64_t n = 0; n < N; ++n) {
    XY raw[M3D_MAXC];
    for (int c = 0; c < C; ++c) {
      raw[c].x = xy[2 * ((int64_t)c * N + n)];
      raw[c].y = xy[2 * ((int64_t)c * N + n) + 1];
    }
    CertOut o;
    // 8-camera pinhole rigs: the compile-time-count instantiation the kernels of the headline config run
    // (straight-line undistortion / half budgets of cert_undistort and cert_pairs)
    // use_cert: bit 0 = prune by the pair certificates, bit 1 = force the run-time-count instantiation
    const bool prune = (use_cert & 1) != 0, general = (use_cert & 2) != 0;
    if (po && C == 8 && !general) ransac_cert_point<true, 8>(rig, cert, &cumb.v[0][0], raw, undistort, min_cams, threshold, init_best, prune, o);
    else if (po) ransac_cert_point<true, 0>(rig, cert, &cumb.v[0][0], raw, undistort, min_cams, threshold, init_best, prune, o);
    else ransac_cert_point<false, 0>(rig, cert, &cumb.v[0][0], raw, undistort, min_cams, threshold, init_best, prune, o);
    p3d[3 * n] = o.X;
    p3d[3 * n + 1] = o.Y;
    p3d[3 * n + 2] = o.Z;
    err[n] = o.err;
    subset[n] = o.s_sel;
    neval[n] = o.neval;
    if (n_solved) n_solved[n] = o.n_solved;
    for (int c = 0; c < C; ++c) {
      const bool in = (o.sel >> (C - 1 - c)) & 1u;
      picked[(int64_t)c * N + n] = in ? 1 : 0;
      xy_picked[2 * ((int64_t)c * N + n)] = in ? raw[c].x : NAN;
      xy_picked[2 * ((int64_t)c * N + n) + 1] = in ? raw[c].y : NAN;
    }
  }
  return 0;
}

}  // extern "C"
