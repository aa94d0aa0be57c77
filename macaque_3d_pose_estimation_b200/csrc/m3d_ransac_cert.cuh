// m3d_ransac_cert.cuh — K4c: camera-subset RANSAC (CameraGroup.triangulate_possible with one
// candidate per camera, cameras.py:639-743) with EXACT PRE-SOLVE PRUNING by pair certificates.
//
// The reference visits the subsets of the k valid cameras in itertools.product order and stops at
// the first one whose mean reprojection error is < T1 = min(threshold, init_best).  Almost all of
// the subsets in front of that one contain a grossly wrong detection; the exhaustive kernels
// (m3d_ransac8.cuh / m3d_ransac16.cuh) reject each of them with a DLT solve and a projection.  Here
// a subset is rejected without any arithmetic when it contains a PAIR of cameras whose two
// observations cannot both lie within the accepted residual budget of ANY 3D point (m3d_cert.h;
// proof in DESIGN.md 3.2c).  The test is conservative: a subset it rejects has
// mean error >= T1 for certain, a subset it cannot reject is solved and scored exactly as before,
// so the selected subset, its error and the number of subsets the reference would have evaluated
// are unchanged.
//
// One THREAD = one point, one kernel, no scratch:
//   * 8-camera rigs: raw and undistorted observations live in registers (camera loops unrolled,
//     camera parameters and the 28 essential matrices are constant-bank operands); other rigs
//     (2..16 cameras) run the same code with run-time loops;
//   * cameras are numbered by the bit of the enumeration step that drops them in a fixed frame
//     (bit C - 1 - c), so that "next subset" is one add on the dropped-camera mask
//     d' = ((d | ~v) + 1) & v and ascending d is the reference's ascending step index s;
//   * pruning is a JUMP: if the kept set contains a certified-bad pair whose lower bit is p, every
//     step up to the next one that sets bit p keeps that pair, so the search continues at
//     ((d >> p) | 1) << p.  The pair masks are computed once per point at the weakest radius
//     rho = T1 (k - 1) (valid for every subset size after the full set);
//   * subsets that survive are solved (dlt_solve) and scored against the raw pixels exactly like
//     the full set; pass 2 (no subset under T1: strict arg-min over all admissible subsets) runs
//     without pruning, because a pruned subset may still be the arg-min;
//   * the number of subsets the reference would have triangulated follows combinatorially from the
//     stopping step (cumulative binomials), no counting loop.
#pragma once
#include <type_traits>
#ifndef M3D_PREP_ILP
#define M3D_PREP_ILP 3  // developer switch, bits: 1 = straight-line undistortion of all cameras (cert_undistort), 2 = straight-line
                        // half budgets (cert_pairs); 0 = per-camera branches in both (round-2e form)
#endif
#ifndef M3D_PREP_GROUP
#define M3D_PREP_GROUP 4  // cameras whose undistortion iterations are interleaved in cert_undistort
#endif
#include "m3d_cert.h"
#include "m3d_math.cuh"
#include "m3d_point.cuh"
#if defined(__CUDACC__)
#include "m3d_ransac.cuh"
#endif

namespace m3d {

// #{ 1 <= s <= n : popc(s) <= m }  (n < 2^16); cumb = CumBinom::v flattened
M3D_HD int count_adm(const uint32_t* cumb, uint32_t n, int m) {
  if (m < 0) return 0;
  int ones = 0, total = 0;
  // the set bits of n from the top down (one trip per set bit: the common stop at the full set, n = 0, costs none)
  for (uint32_t rest = n & 0xffffu; rest != 0;) {
    const int i = 31 - M3D_CLZ(rest);
    rest ^= 1u << i;
    const int rem = m - ones;
    if (rem < 0) break;
    total += (int)cumb[i * 17 + (rem < i ? rem : i)];
    ++ones;
  }
  if (ones <= m) total += 1;  // n itself (the loop was not cut short)
  return total - 1;           // s = 0 is not counted here
}

struct XY {
  double x, y;
};

struct CertOut {
  double X, Y, Z, err;  // err: 0.0 when nothing was selected (cameras.py:675)
  uint32_t sel;         // cameras of the selection, bit C - 1 - c
  int32_t s_sel;        // step index of the selection, -1 when none
  int32_t neval;        // subsets the reference would have triangulated
  int32_t n_solved;     // subsets this search solved (diagnostic of the CPU test tier)
  int32_t n_visited;    // steps of cert_advance (diagnostic)
};

// ---- building blocks (shared by the reference single-thread form below, the kernels and the CPU
// test tier).  NC > 0: camera count known at compile time (loops unroll, arrays live in registers);
// NC == 0: run-time count up to M3D_MAXC.  Bit position of camera c: C - 1 - c.

// per-camera undistortion (cameras.py:608-614) and the two validity masks
template <bool PO, int NC>
M3D_HD void cert_undistort(const RigDev& rig, const XY* raw, int undistort, XY* xh, uint32_t& v, uint32_t& u) {
  constexpr int CC = NC > 0 ? NC : M3D_MAXC;
  const int C = NC > 0 ? NC : rig.n_cams;
  const uint32_t TOP = (uint32_t)(C - 1);
  v = 0, u = 0;
#if (M3D_PREP_ILP & 1)
  if constexpr (PO && NC > 0) {
    // Rigs of pinhole cameras, count known at compile time: the cameras' iterations as ONE straight-line
    // block (no validity branch, no bail-out branch between them), so that the scheduler can keep several
    // cameras' dependency chains in flight per thread; an invalid view runs on NaN and is masked afterwards,
    // the icdist < 0 replay of all cameras is one rare branch at the end.  Same arithmetic per camera, same
    // results bit for bit.
    if (undistort) {
      double ux[CC], uy[CC];
      int neg = 0;
      constexpr int G = (CC % M3D_PREP_GROUP == 0) ? M3D_PREP_GROUP : 1;  // cameras in flight per thread
#pragma unroll
      for (int c0 = 0; c0 < CC; c0 += G) {
        double ru[G], rv[G];
        int ng[G];
#pragma unroll
        for (int g = 0; g < G; ++g) ru[g] = raw[c0 + g].x, rv[g] = raw[c0 + g].y;
        undistort_pinhole_core_group<false, G>(&rig.cam[c0], ru, rv, &ux[c0], &uy[c0], ng);
#pragma unroll
        for (int g = 0; g < G; ++g) neg |= (ru[g] == ru[g]) ? ng[g] : 0;
      }
      if (neg < 0) {  // rare; unrolled (a run-time camera index would put ux / uy in local memory)
#pragma unroll
        for (int c = 0; c < CC; ++c)
          if (raw[c].x == raw[c].x) undistort_pinhole_replay<false>(rig.cam[c], raw[c].x, raw[c].y, ux[c], uy[c]);
      }
#pragma unroll
      for (int c = 0; c < CC; ++c) {
        const bool valid = raw[c].x == raw[c].x;  // raw x not NaN (cameras.py:658-659)
        const double x = valid ? ux[c] : raw[c].x, y = valid ? uy[c] : raw[c].y;
        v |= valid ? (1u << (TOP - c)) : 0u;
        u |= (valid && x == x) ? (1u << (TOP - c)) : 0u;  // usable inside triangulate (cameras.py:630)
        xh[c].x = x;
        xh[c].y = y;
      }
      return;
    }
  }
#endif
#pragma unroll
  for (int c = 0; c < CC; ++c) {
    if (c < C) {
      double x = raw[c].x, y = raw[c].y;
      if (x == x) {  // valid: raw x not NaN (cameras.py:658-659)
        v |= 1u << (TOP - c);
        if (undistort) undistort_point<false, PO>(rig.cam[c], raw[c].x, raw[c].y, x, y);
        if (x == x) u |= 1u << (TOP - c);  // usable inside triangulate (cameras.py:630)
      }
      xh[c].x = x;
      xh[c].y = y;
    }
  }
}

// triangulate from the usable cameras of `kept`, score against the raw pixels of all of `kept`
// (cameras.py:697-701): mean reprojection error, NaN when undefined
// (RV / XV: anything indexable by camera that yields an XY — plain arrays, or the shared-memory views
// of k_cert_search)
template <bool PO, int NC, class RV, class XV>
M3D_HD double cert_eval(const RigDev& rig, const RV& raw, const XV& xh, uint32_t kept, uint32_t uc, double& X,
                        double& Y, double& Z) {
  constexpr int CC = NC > 0 ? NC : M3D_MAXC;
  const int C = NC > 0 ? NC : rig.n_cams;
  const uint32_t TOP = (uint32_t)(C - 1);
  X = Y = Z = qnan();
  if (M3D_POPC(uc) < 2) return qnan();
  // Branch-free over the cameras: every lane of a warp holds a different subset, so a branch per camera
  // serialises eight partially filled blocks and keeps the compiler from interleaving them.  Instead every
  // camera is processed by every lane and a 0 / 1 weight removes what the subset does not contain: the rows of
  // an excluded camera are built from a finite stand-in observation and scaled by 0 (adds exactly +0), those of
  // an included one by 1 (exact) — the sums are bit-identical to the branched form.
  Gram G;
  gram_zero(G);
#pragma unroll
  for (int c = 0; c < CC; ++c) {
    if (c < C) {
      const bool in = ((uc >> (TOP - c)) & 1u) != 0;
      const XY q = xh[c];
      const double w = in ? 1.0 : 0.0;
      const double x = in ? q.x : 0.0, y = in ? q.y : 0.0;
      const CamDev& cam = rig.cam[c];
      gram_add_row(G, w * (x * cam.R[6] - cam.R[0]), w * (x * cam.R[7] - cam.R[1]), w * (x * cam.R[8] - cam.R[2]),
                   w * (x * cam.t[2] - cam.t[0]));
      gram_add_row(G, w * (y * cam.R[6] - cam.R[3]), w * (y * cam.R[7] - cam.R[4]), w * (y * cam.R[8] - cam.R[5]),
                   w * (y * cam.t[2] - cam.t[1]));
    }
  }
  dlt_solve(G, X, Y, Z);
  double sum = 0.0;
  int m = 0;
#pragma unroll
  for (int c = 0; c < CC; ++c) {
    if (c < C) {
      double pu, pv;
      project_point<false, PO>(rig.cam[c], X, Y, Z, pu, pv);
      const XY q = raw[c];
      const double e = residual_norm(q.x - pu, q.y - pv);
      const bool use = (((kept >> (TOP - c)) & 1u) != 0) && (e == e);
      sum += use ? e : 0.0;  // e >= 0: adding +0 leaves the sum as it is
      m += use ? 1 : 0;
    }
  }
  return (m >= 2) ? sum / (double)m : qnan();
}

// f(integral_constant<a>, integral_constant<b>) for every camera pair a < b < CC, indices known at compile time
template <int NC, int A, int B, class Fn>
M3D_HD void cert_for_pairs(Fn&& f) {
  constexpr int CC = NC > 0 ? NC : M3D_MAXC;
  if constexpr (A < CC - 1) {
    f(std::integral_constant<int, A>{}, std::integral_constant<int, B>{});
    if constexpr (B + 1 < CC) cert_for_pairs<NC, A, B + 1>(f);
    else cert_for_pairs<NC, A + 1, A + 2>(f);
  }
}

// pair certificates at residual budget rho (m3d_cert.h): badrow[p] = certified-bad partners of bit
// position p at HIGHER positions.  rho_full > 0: also report (full_bad) whether some pair is flagged
// at that larger budget — the budget of the FULL set, |S| = k — so that the full-set solve itself
// can be skipped.  Straight-line per pair (no branch): a camera that cannot take part (no
// certificate, unusable view, non-finite or far-off centre) carries an infinite half-budget, which makes
// the right-hand side infinite; a NaN form compares false.
//
// The per-camera part (delta_c, half budgets) is float64.  The 28 (120) pair forms run in FLOAT32 on the
// otherwise idle FMA pipe, made conservative by an explicit error budget ("Floating
// point"): with unit-Frobenius E and n_c = |(x_c, y_c, 1)|, every float32 sum of products below is within
// 7 * 2^-24 * n_a n_b (the form) resp. 5 * 2^-24 * n_c (the gradient rows) of its exact value — inputs
// rounded to nearest, one rounding per fma — so that
//     F^ - 2e-6 n_a n_b  >  ((max(A^, B^) with + 2e-6 n / (mu f) inside) + g^ D^) D^ (1 + 2e-5)
// implies the exact inequality F > (max(A, B) + g D) D that the proof needs; g^, 1 / (mu f), the half
// budgets and n_c enter rounded UP.  A pair is flagged on px-level margins, the float32 slack is ~1e-5 of
// the right-hand side: the flags are, for all practical purposes, those of the float64 evaluation.
template <int NC>
M3D_HD void cert_pairs(const RigDev& rig, const CertDev& cert, const XY* raw, const XY* xh, uint32_t u,
                       double rho, uint32_t* badrow, double rho_full = 0.0, bool* full_bad = nullptr) {
  constexpr int CC = NC > 0 ? NC : M3D_MAXC;
  const int C = NC > 0 ? NC : rig.n_cams;
  const uint32_t TOP = (uint32_t)(C - 1);
  // half budgets per camera: D = h[a] + h[b] = (rho + delta_a + delta_b) (1 + 1e-9); the slack on D
  // covers the comparison (the right-hand side grows faster than linearly in D)
  float h[CC], hf[CC], xf[CC], yf[CC], nf[CC];
#pragma unroll
  for (int c = 0; c < CC; ++c) {
    h[c] = hf[c] = (float)pos_inf();
    xf[c] = yf[c] = 0.0f;
    nf[c] = 1.0f;
    badrow[c] = 0;
    if constexpr ((M3D_PREP_ILP & 2) != 0 && NC > 0) {
      // straight-line over the cameras (the per-camera branches of the general form below serialise the eight
      // distortion + square-root chains): computed for every camera, kept by a select — a view that takes no
      // part (NaN centre, no certificate) fails the comparisons and keeps the infinite half-budget
      double pu, pv;
      distort_pinhole<false>(rig.cam[c], xh[c].x, xh[c].y, pu, pv);
      const double e = residual_norm(raw[c].x - pu, raw[c].y - pv);
      const double x = xh[c].x, y = xh[c].y;
      const bool ok = ((u >> (TOP - c)) & 1u) && cert.inv_mf[c] > 0.0 && e < 1e3 && fabs(x) <= 100.0 && fabs(y) <= 100.0;
      const float hc = (float)((e + 0.5 * rho) * (1.0 + 1e-6));
      const float hfc = (float)((e + 0.5 * rho_full) * (1.0 + 1e-6));
      const float nfc = (float)(sqrt_fast(fma(x, x, fma(y, y, 1.0))) * (1.0 + 1e-6));
      h[c] = ok ? hc : h[c];
      hf[c] = ok ? hfc : hf[c];
      xf[c] = ok ? (float)x : 0.0f;
      yf[c] = ok ? (float)y : 0.0f;
      nf[c] = ok ? nfc : 1.0f;
    } else {
      if (c < C && ((u >> (TOP - c)) & 1u) && cert.inv_mf[c] > 0.0) {
        double pu, pv;
        distort_pinhole<false>(rig.cam[c], xh[c].x, xh[c].y, pu, pv);
        const double e = residual_norm(raw[c].x - pu, raw[c].y - pv);
        const double x = xh[c].x, y = xh[c].y;
        // finite centre inside |x|, |y| <= 100 (float32 range of the forms), finite raw pixel
        if (e < 1e3 && fabs(x) <= 100.0 && fabs(y) <= 100.0) {
          h[c] = (float)((e + 0.5 * rho) * (1.0 + 1e-6));        // rounded to nearest of a value inflated by 1e-6: >= exact
          hf[c] = (float)((e + 0.5 * rho_full) * (1.0 + 1e-6));
          xf[c] = (float)x;
          yf[c] = (float)y;
          nf[c] = (float)(sqrt_fast(fma(x, x, fma(y, y, 1.0))) * (1.0 + 1e-6));
        }
      }
    }
  }
  bool fb = false;
  // every pair with COMPILE-TIME camera indices (cert_for_pairs): a nested `#pragma unroll` leaves the inner loop
  // rolled (28 / 120 bodies), which turns the per-camera arrays above into local memory
  cert_for_pairs<NC, 0, 1>([&](auto a_, auto b_) {
    constexpr int a = decltype(a_)::value, b = decltype(b_)::value;
    if (b >= C) return;
    const float* E = cert.Ef[pair_index(a, b, C)];
    const float ax = xf[a], ay = yf[a], qx = xf[b], qy = yf[b];
    const float ea0 = fmaf(E[0], ax, fmaf(E[1], ay, E[2]));
    const float ea1 = fmaf(E[3], ax, fmaf(E[4], ay, E[5]));
    const float ea2 = fmaf(E[6], ax, fmaf(E[7], ay, E[8]));
    const float F = fabsf(fmaf(qx, ea0, fmaf(qy, ea1, ea2)));
    const float tb0 = fmaf(E[0], qx, fmaf(E[3], qy, E[6]));
    const float tb1 = fmaf(E[1], qx, fmaf(E[4], qy, E[7]));
    const float A = (fabsf(tb0) + fabsf(tb1) + 2e-6f * nf[b]) * cert.inv_mf_f[a];
    const float B = (fabsf(ea0) + fabsf(ea1) + 2e-6f * nf[a]) * cert.inv_mf_f[b];
    const float M = A > B ? A : B;
    const float g = E[9];  // gam / 4, rounded up
    const float lhs = fmaf(-2e-6f * nf[a], nf[b], F);
    const float D = h[a] + h[b];
    const bool bad = lhs > fmaf(g, D, M) * D * 1.00002f;
    badrow[TOP - b] |= bad ? (1u << (TOP - a)) : 0u;
    if (full_bad) {
      const float Df = hf[a] + hf[b];
      fb = fb || (lhs > fmaf(g, Df, M) * Df * 1.00002f);
    }
  });
  if (full_bad) *full_bad = fb;
}

// residual budget of the pair certificates: the weakest radius any subset after the full set can
// have (|S| <= k - 1), plus 1e-6 px for the rounding of the reference's own projections
M3D_HD double cert_rho(double T1, int k) { return T1 * (double)(k - 1) * (1.0 + 1e-9) + 1e-6; }

M3D_HD uint32_t cert_next(uint32_t v, uint32_t d) { return ((d | ~v) + 1u) & v; }

// Smallest dropped-camera mask >= d0 in enumeration order (0 = d0 wrapped around: exhausted) whose
// kept set contains no certified-bad pair and at least min_cams cameras; 0 when there is none.
// One pass from the top bit down: follow d0 until a camera it keeps conflicts with a camera kept
// above it — that camera has to go, which makes the mask larger than d0 at the highest possible
// position; from there on keep every camera that does not conflict (lexicographically smallest
// completion).  Every mask in between keeps a conflicting pair, i.e. is certified-bad.
template <int NC>
M3D_HD uint32_t cert_advance(uint32_t v, uint32_t d0, const uint32_t* badrow, int min_cams) {
  constexpr int CC = NC > 0 ? NC : M3D_MAXC;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  while (d0 != 0) {
    uint32_t kept = 0, dn = 0;
    bool departed = false;
#pragma unroll
    for (int p = CC - 1; p >= 0; --p) {
      const uint32_t bit = 1u << p;
      if (v & bit) {
        const bool conflict = (badrow[p] & kept) != 0;
        const bool follow = !departed && ((d0 & bit) != 0);
        if (conflict && !follow) departed = true;
        if (conflict || follow) dn |= bit;
        else kept |= bit;
      }
    }
    if (M3D_POPC(kept) >= min_cams) return dn;
    d0 = cert_next(v, dn);
  }
  return 0;
}

// step index s of a dropped-camera mask: its bits at the valid positions, compacted
template <int NC>
M3D_HD uint32_t cert_step_index(uint32_t v, uint32_t d) {
  constexpr int CC = NC > 0 ? NC : M3D_MAXC;
  uint32_t s = 0;
  int j = 0;
#pragma unroll
  for (int p = 0; p < CC; ++p) {
    if ((v >> p) & 1u) {
      s |= ((d >> p) & 1u) << j;
      ++j;
    }
  }
  return s;
}

// One point of the pruned subset search, start to end in one thread: the reference form of the
// algorithm (CPU test tier, tests/host_harness.cpp) — the kernels below run the same blocks, split
// over two launches so that the lanes of a warp stay busy.
template <bool PO, int NC>
M3D_HD void ransac_cert_point(const RigDev& rig, const CertDev& cert, const uint32_t* cumb, const XY* raw,
                              int undistort, int min_cams, double thr, double init_best, bool use_cert,
                              CertOut& out) {
  constexpr int CC = NC > 0 ? NC : M3D_MAXC;
  const double T1 = thr < init_best ? thr : init_best;
  XY xh[CC];
  uint32_t v, u;
  cert_undistort<PO, NC>(rig, raw, undistort, xh, v, u);
  const int k = M3D_POPC(v);
  uint32_t badrow[CC];
#pragma unroll
  for (int p = 0; p < CC; ++p) badrow[p] = 0;
  uint32_t d = 0, best_d = 0;
  int pass = 0;
  bool have = false, ran_out = false;
  double best_err = init_best, bx = qnan(), by = qnan(), bz = qnan();
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  int n_solved = 0;
  for (;;) {
    const uint32_t kept = v & ~d;
    double X, Y, Z;
    const double err = cert_eval<PO, NC>(rig, raw, xh, kept, kept & u, X, Y, Z);
    ++n_solved;
    // ---- accept / continue (cameras.py:703-713)
    bool stop = false;
    if (pass == 0) {
      if (err < best_err) {
        best_err = err, best_d = 0, have = true, bx = X, by = Y, bz = Z;
        stop = err < thr;
      }
      if (k < 2 || k <= min_cams) stop = true;  // every smaller subset would be skipped (:691)
      if (!stop && use_cert) cert_pairs<NC>(rig, cert, raw, xh, u, cert_rho(T1, k), badrow);
      pass = 1;
    } else if (pass == 1) {
      if (err < T1) {
        best_err = err, best_d = d, have = true, bx = X, by = Y, bz = Z;
        stop = true;
      }
    } else {
      if (err < best_err) best_err = err, best_d = d, have = true, bx = X, by = Y, bz = Z;
    }
    if (stop) break;
    d = cert_advance<NC>(v, cert_next(v, d), badrow, min_cams);
    if (d == 0 && pass == 1 && T1 < best_err) {
      // nothing under T1: rescan everything for the strict arg-min, without pruning (a pruned
      // subset has err >= T1 but may still be the arg-min)
      pass = 2;
#pragma unroll
      for (int p = 0; p < CC; ++p) badrow[p] = 0;
      d = cert_advance<NC>(v, cert_next(v, 0u), badrow, min_cams);
    }
    if (d == 0) {
      ran_out = true;
      break;
    }
  }
  const uint32_t s_sel = cert_step_index<NC>(v, best_d);
  out.X = bx;
  out.Y = by;
  out.Z = bz;
  out.err = have ? best_err : 0.0;
  out.s_sel = have ? (int32_t)s_sel : -1;
  out.sel = have ? (v & ~best_d) : 0u;
  int ne = 1;
  if (ran_out) ne += count_adm(cumb, (1u << k) - 1u, k - min_cams);  // never stopped
  else if (best_d != 0) ne += count_adm(cumb, s_sel, k - min_cams);  // stopped at s_sel in pass 1
  out.neval = ne;
  out.n_solved = n_solved * (pass == 2 ? -1 : 1);  // negative: the point needed the arg-min pass
  out.n_visited = 0;
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------
// kernels.  Per-point cost is 1 evaluation for ~20 % of the points, 2 for most, 10+ for a tail:
// run start-to-end per thread, a warp waits for its slowest lane (measured: 13 evaluations per
// warp for 2.4 per point).  Hence two launches:
//   k_cert_setup   thread = point, convergent: undistort, full-set evaluation, decision, pair
//                  certificates; writes the result slot of every point and, for the undecided
//                  ones, a record (raw + undistorted views, pair masks) into a compact queue
//   k_cert_search  persistent lanes: every lane owns one queued point and evaluates one
//                  surviving subset per trip; a lane that finishes takes the next record at once,
//                  so all 32 lanes evaluate on (almost) every trip
//   k_cert_overflow  the few searches that outlast `lane_limit` evaluations in a lane (points
//                  without a clean subset: up to 2^k evaluations) are parked by k_cert_search and
//                  finished here warp-cooperatively, 32 surviving subsets per round
//   k_ransac_emit  (m3d_ransac.cuh) expands the slots into the reference's outputs
// Record q: fields of 16 bytes, field f of record q at ((q / 32) * F + f) * 32 + q % 32 (so that 32
// consecutive records are read / written with full sectors):
//   f <  C       raw (x, y) of camera f          f < 2C   undistorted (x, y) of camera f - C
//   f == 2C      v | u << 16, point index, -, -  f == 2C + 1, 2C + 2   best_err, bx | by, bz
//   f >= 2C + 3  pair masks, 16 bits per bit position
// ---------------------------------------------------------------------------------------
M3D_HD int cert_record_fields(int C) { return 2 * C + 3 + (C + 7) / 8; }

template <bool PO, int NC, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_cert_setup(const __grid_constant__ RigDev rig, const __grid_constant__ CertDev cert,
             const double* __restrict__ xy, int64_t ld, int64_t n0, int64_t n, int undistort, int min_cams,
             double thr, double init_best, RansacSlot* __restrict__ slots, double2* __restrict__ rec,
             unsigned int* __restrict__ n_rec) {
  constexpr int CC = NC > 0 ? NC : M3D_MAXC;
  constexpr unsigned FULLM = 0xffffffffu;
  const int C = NC > 0 ? NC : rig.n_cams;
  const int F = cert_record_fields(C);
  const double T1 = thr < init_best ? thr : init_best;
  const int lane = threadIdx.x & 31;
  // whole warps iterate together (the queue append is warp-aggregated)
  const int64_t n_round = (n + 31) & ~(int64_t)31;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round;
       i += (int64_t)gridDim.x * blockDim.x) {
    const bool inb = i < n;
    XY raw[CC], xh[CC];
#pragma unroll
    for (int c = 0; c < CC; ++c) {
      if (c < C) {
        double2 q = make_double2(qnan(), qnan());
        if (inb) q = ld_xy(xy, (int64_t)c * ld + n0 + i);
        raw[c].x = q.x;
        raw[c].y = q.y;
      }
    }
    uint32_t v, u;
    cert_undistort<PO, NC>(rig, raw, undistort, xh, v, u);
    const int k = __popc(v);
    double X, Y, Z;
    const double err = cert_eval<PO, NC>(rig, raw, xh, v, u, X, Y, Z);
    bool stop = false, have = false;
    double best_err = init_best;
    if (err < best_err) {
      best_err = err, have = true;
      stop = err < thr;
    }
    if (k < 2 || k <= min_cams) stop = true;  // every smaller subset would be skipped (:691)
    if (inb) {
      RansacSlot sl;
      sl.best_err = best_err;
      sl.bx = have ? X : qnan();
      sl.by = have ? Y : qnan();
      sl.bz = have ? Z : qnan();
      sl.ord = 0;
      sl.masks = __brev(v) >> (32 - C);  // physical camera numbering for k_ransac_emit
      sl.vlist = 0;
      sl.best_s = have ? 0 : -1;
      sl.neval = 1;
      sl.decided = stop ? 1 : 0;
      sl.uml = 0;
      slots[i] = sl;
    }
    const bool queue = inb && !stop;
    const uint32_t qb = __ballot_sync(FULLM, queue);
    if (qb) {
      uint32_t badrow[CC];
      if (queue) cert_pairs<NC>(rig, cert, raw, xh, u, cert_rho(T1, k), badrow);
      unsigned int base = 0;
      if (lane == __ffs(qb) - 1) base = atomicAdd(n_rec, (unsigned int)__popc(qb));
      base = __shfl_sync(FULLM, base, __ffs(qb) - 1);
      if (queue) {
        const unsigned int q = base + (unsigned int)__popc(qb & ((1u << lane) - 1u));
        double2* r = rec + ((size_t)(q >> 5) * F) * 32 + (q & 31u);
#pragma unroll
        for (int c = 0; c < CC; ++c) {
          if (c < C) {
            r[(size_t)c * 32] = make_double2(raw[c].x, raw[c].y);
            r[(size_t)(C + c) * 32] = make_double2(xh[c].x, xh[c].y);
          }
        }
        uint4 m0 = make_uint4(v | (u << 16), (uint32_t)i, 0u, 0u);
        reinterpret_cast<uint4*>(r)[(size_t)(2 * C) * 32] = m0;
        r[(size_t)(2 * C + 1) * 32] = make_double2(best_err, have ? X : qnan());
        r[(size_t)(2 * C + 2) * 32] = make_double2(have ? Y : qnan(), have ? Z : qnan());
#pragma unroll
        for (int g = 0; g < (CC + 7) / 8; ++g) {
          if (8 * g < C) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int p0 = 8 * g + 2 * j, p1 = p0 + 1;
              w[j] = (p0 < CC ? (badrow[p0 < CC ? p0 : 0] & 0xffffu) : 0u) |
                     ((p1 < CC ? (badrow[p1 < CC ? p1 : 0] & 0xffffu) : 0u) << 16);
            }
            reinterpret_cast<uint4*>(r)[(size_t)(2 * C + 3 + g) * 32] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
    }
  }
}

// v2 split: NO solve here.  Undistort, pair flags (at the budget of the subsets after the full set, and
// whether the full set itself is already excluded at ITS budget), one record per point for EVERY point.
// All solves — the full set's included — then run in k_cert_search with all 32 lanes busy; for the
// ~75 % of the points whose full set holds a flagged pair the full-set solve disappears altogether.
template <bool PO, int NC, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_cert_prep(const __grid_constant__ RigDev rig, const __grid_constant__ CertDev cert,
            const double* __restrict__ xy, int64_t ld, int64_t n0, int64_t n, int undistort, double thr,
            double init_best, double2* __restrict__ rec, unsigned int* __restrict__ n_rec) {
  constexpr int CC = NC > 0 ? NC : M3D_MAXC;
  const int C = NC > 0 ? NC : rig.n_cams;
  const int F = cert_record_fields(C);
  const double T1 = thr < init_best ? thr : init_best;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    XY raw[CC], xh[CC];
#pragma unroll
    for (int c = 0; c < CC; ++c) {
      if (c < C) {
        const double2 q = ld_xy(xy, (int64_t)c * ld + n0 + i);
        raw[c].x = q.x;
        raw[c].y = q.y;
      }
    }
    uint32_t v, u;
    cert_undistort<PO, NC>(rig, raw, undistort, xh, v, u);
    const int k = __popc(v);
    // record i of this launch (the queue is the identity: every point has one).  The views are stored BEFORE
    // the pair forms: 64 registers of float64 observations are dead while the 28 forms run.
    const unsigned int q = (unsigned int)i;
    double2* r = rec + ((size_t)(q >> 5) * F) * 32 + (q & 31u);
#pragma unroll
    for (int c = 0; c < CC; ++c) {
      if (c < C) {
        r[(size_t)c * 32] = make_double2(raw[c].x, raw[c].y);
        r[(size_t)(C + c) * 32] = make_double2(xh[c].x, xh[c].y);
      }
    }
    uint32_t badrow[CC];
    bool full_bad = false;
    cert_pairs<NC>(rig, cert, raw, xh, u, cert_rho(T1, k), badrow, cert_rho(T1, k + 1), &full_bad);
    // m0.z: bit 0 = the full set is still to be visited, bit 1 = it holds no flagged pair (solve it),
    // bit 2 = record of the v2 split (the full set's error has not been recorded anywhere)
    reinterpret_cast<uint4*>(r)[(size_t)(2 * C) * 32] =
        make_uint4(v | (u << 16), (uint32_t)i, 1u | (full_bad ? 0u : 2u) | 4u, 0u);
    r[(size_t)(2 * C + 1) * 32] = make_double2(init_best, qnan());
    r[(size_t)(2 * C + 2) * 32] = make_double2(qnan(), qnan());
#pragma unroll
    for (int g = 0; g < (CC + 7) / 8; ++g) {
      if (8 * g < C) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int p0 = 8 * g + 2 * j, p1 = p0 + 1;
          w[j] = (p0 < CC ? (badrow[p0 < CC ? p0 : 0] & 0xffffu) : 0u) |
                 ((p1 < CC ? (badrow[p1 < CC ? p1 : 0] & 0xffffu) : 0u) << 16);
        }
        reinterpret_cast<uint4*>(r)[(size_t)(2 * C + 3 + g) * 32] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    if (i == 0) *n_rec = (unsigned int)n;
  }
}

// The reference's outputs of one launch (cameras.py:715-724); NULL members are not written.  When `p3d`
// is NULL the kernels fill the result slot instead and k_ransac_emit expands it.
struct CertOutputs {
  double* p3d;
  uint8_t* picked;
  double* xy_picked;
  double* err;
  int32_t* subset;
  int32_t* neval;
  int64_t ld;  // points per camera plane of picked / xy_picked
  int64_t n0;  // first point of this launch
};

// Result of one point straight into the output arrays (v2 split: queue order is point order, so the
// stores of neighbouring lanes and warps meet in L2 before they reach DRAM).
template <int NC, class RV>
__device__ __forceinline__ void cert_write_point(const CertOutputs& o, int C, uint32_t idx, uint32_t v, uint32_t best_d,
                                                 bool have, double best_err, double bx, double by, double bz,
                                                 int32_t s_sel, int32_t ne, const RV& raw) {
  constexpr int CC = NC > 0 ? NC : M3D_MAXC;
  const int64_t i = o.n0 + (int64_t)idx;
  o.p3d[3 * i] = bx;
  o.p3d[3 * i + 1] = by;
  o.p3d[3 * i + 2] = bz;
  o.err[i] = have ? best_err : 0.0;  // errors default to 0.0 (cameras.py:675)
  if (o.subset) o.subset[i] = have ? s_sel : -1;
  if (o.neval) o.neval[i] = ne;
  const uint32_t sel = have ? (v & ~best_d) : 0u;
#pragma unroll
  for (int c = 0; c < CC; ++c) {
    if (c < C) {
      const bool in = (sel >> (C - 1 - c)) & 1u;
      if (o.picked) o.picked[(int64_t)c * o.ld + i] = in ? 1 : 0;
      if (o.xy_picked) {
        XY q = raw[c];
        st_xy(o.xy_picked, (int64_t)c * o.ld + i, in ? q.x : qnan(), in ? q.y : qnan());
      }
    }
  }
}

template <bool PO, int NC, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_cert_search(const __grid_constant__ RigDev rig, const uint32_t* __restrict__ cumb, int min_cams, double thr,
              double init_best, RansacSlot* __restrict__ slots, double2* __restrict__ rec,
              const unsigned int* __restrict__ n_rec, unsigned int* __restrict__ counter,
              unsigned int* __restrict__ over, unsigned int* __restrict__ n_over, int lane_limit,
              const CertOutputs outs) {
  constexpr int CC = NC > 0 ? NC : M3D_MAXC;
  constexpr unsigned FULLM = 0xffffffffu;
  const int C = NC > 0 ? NC : rig.n_cams;
  const int F = cert_record_fields(C);
  const double T1 = thr < init_best ? thr : init_best;
  const int lane = threadIdx.x & 31;
  const unsigned int total = *n_rec;
  // the observations of a lane's point live in shared memory ([field][thread], 16-byte accesses by
  // consecutive lanes: conflict-free): 64 registers less per thread, a quarter more resident warps
  extern __shared__ __align__(16) unsigned char smem_search[];
  struct View {
    XY* base;
    __device__ __forceinline__ XY operator[](int c) const { return base[c * 128]; }
  };
  XY* sbase = reinterpret_cast<XY*>(smem_search) + threadIdx.x;
  const View raw{sbase}, xh{sbase + (size_t)C * 128};
  uint32_t badrow[CC];
  uint32_t v = 0, u = 0, d = 0, best_d = 0, idx = 0, qrec = 0;
  int pass = 1, n_done = 0;
  bool active = false, drained = false, have = false;
  bool first = false, full_ok = false, is_v2 = false;  // v2 records: the full set is the first candidate (if not excluded)
  double best_err = 0.0, bx = 0.0, by = 0.0, bz = 0.0;
#pragma unroll 1
  for (;;) {
    // ---- lanes without a point take the next records of the queue
    const bool need = !active && !drained;
    const uint32_t nb = __ballot_sync(FULLM, need);
    if (nb) {
      unsigned int base = 0;
      if (lane == __ffs(nb) - 1) base = atomicAdd(counter, (unsigned int)__popc(nb));
      base = __shfl_sync(FULLM, base, __ffs(nb) - 1);
      if (need) {
        const unsigned int q = base + (unsigned int)__popc(nb & ((1u << lane) - 1u));
        if (q < total) {
          const double2* r = rec + ((size_t)(q >> 5) * F) * 32 + (q & 31u);
#pragma unroll
          for (int c = 0; c < CC; ++c) {
            if (c < C) {
              const double2 a = r[(size_t)c * 32], b = r[(size_t)(C + c) * 32];
              reinterpret_cast<double2*>(raw.base)[c * 128] = a;
              reinterpret_cast<double2*>(xh.base)[c * 128] = b;
            }
          }
          const uint4 m0 = reinterpret_cast<const uint4*>(r)[(size_t)(2 * C) * 32];
          v = m0.x & 0xffffu;
          u = m0.x >> 16;
          idx = m0.y;
          const double2 b0 = r[(size_t)(2 * C + 1) * 32], b1 = r[(size_t)(2 * C + 2) * 32];
          best_err = b0.x, bx = b0.y, by = b1.x, bz = b1.y;
          have = bx == bx;  // the full set was accepted (its error is >= T1, else the point was decided)
#pragma unroll
          for (int g = 0; g < (CC + 7) / 8; ++g) {
            if (8 * g < C) {
              const uint4 w = reinterpret_cast<const uint4*>(r)[(size_t)(2 * C + 3 + g) * 32];
              const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (8 * g + 2 * j < CC) badrow[8 * g + 2 * j] = ww[j] & 0xffffu;
                if (8 * g + 2 * j + 1 < CC) badrow[8 * g + 2 * j + 1] = ww[j] >> 16;
              }
            }
          }
          best_d = 0;
          pass = 1;
          active = true;
          qrec = q;
          n_done = 0;
          d = 0;
          first = (m0.z & 1u) != 0;
          full_ok = (m0.z & 2u) != 0;
          is_v2 = (m0.z & 4u) != 0;
        } else {
          drained = true;
        }
      }
    }
    if (!__any_sync(FULLM, active)) break;
    if (active) {
      // ---- every lane moves to its next surviving subset (one convergent pass per trip)
      bool stop = false, ran_out = false;
      bool at_full = false;  // the candidate of this trip is the full set (d == 0, always admissible: :691)
      if (first && full_ok) {
        at_full = true;
      } else {
        d = cert_advance<NC>(v, cert_next(v, d), badrow, min_cams);
      }
      first = false;
      if (!at_full && d == 0 && pass == 1 && T1 < best_err) {
        // nothing under T1: rescan everything for the strict arg-min, without pruning.  v2 records
        // restart AT the full set (its error was never recorded), v1 records after it.
        pass = 2;
#pragma unroll
        for (int p = 0; p < CC; ++p) badrow[p] = 0;
        if (is_v2) at_full = true;
        else d = cert_advance<NC>(v, cert_next(v, 0u), badrow, min_cams);
      }
      if (at_full) d = 0;
      if (!at_full && d == 0) {
        stop = ran_out = true;
      } else if (!at_full && n_done >= lane_limit) {
        // a long search (a point without a clean subset, a point far outside the images) would
        // keep this lane — in the end this whole warp — busy for hundreds of trips: park its
        // state in the record and hand it to the warp-cooperative kernel
        double2* r = rec + ((size_t)(qrec >> 5) * F) * 32 + (qrec & 31u);
        reinterpret_cast<uint4*>(r)[(size_t)(2 * C) * 32] =
            make_uint4(v | (u << 16), idx, d, (uint32_t)pass | (have ? 256u : 0u) | (is_v2 ? 512u : 0u) | (best_d << 16));
        r[(size_t)(2 * C + 1) * 32] = make_double2(best_err, bx);
        r[(size_t)(2 * C + 2) * 32] = make_double2(by, bz);
        over[atomicAdd(n_over, 1u)] = qrec;
        active = false;
      } else {
        const uint32_t kept = v & ~d;
        double X, Y, Z;
        const double err = cert_eval<PO, NC>(rig, raw, xh, kept, kept & u, X, Y, Z);
        ++n_done;
        if (pass == 1) {
          if (err < T1) {
            best_err = err, best_d = d, have = true, bx = X, by = Y, bz = Z;
            stop = true;
          }
        } else if (err < best_err) {
          best_err = err, best_d = d, have = true, bx = X, by = Y, bz = Z;
        }
      }
      if (stop) {
        const int k = __popc(v);
        const uint32_t s_sel = cert_step_index<NC>(v, best_d);
        int ne = 1;
        if (ran_out) ne += count_adm(cumb, (1u << k) - 1u, k - min_cams);
        else ne += count_adm(cumb, s_sel, k - min_cams);
        if (outs.p3d) {
          cert_write_point<NC>(outs, C, idx, v, best_d, have, best_err, bx, by, bz, (int32_t)s_sel, ne, raw);
        } else {
          RansacSlot* sl = slots + idx;
          sl->best_err = best_err;
          sl->bx = bx;
          sl->by = by;
          sl->bz = bz;
          sl->masks = __brev(v) >> (32 - C);  // physical camera numbering for k_ransac_emit
          sl->best_s = have ? (int32_t)s_sel : -1;
          sl->neval = ne;
        }
        active = false;
      }
    }
  }
}

// Warp-cooperative continuation of the searches k_cert_search parked: one warp = one point, lane j
// takes the j-th surviving subset after the current one (the subsets of a round are evaluated at
// once); pass 1 stops at the first lane (= first subset in enumeration order) under T1, pass 2
// reduces the strict arg-min (first on ties) over all lanes at the end.
template <bool PO, int NC>
__global__ void __launch_bounds__(128, 2)
k_cert_overflow(const __grid_constant__ RigDev rig, const uint32_t* __restrict__ cumb, int min_cams, double thr,
                double init_best, RansacSlot* __restrict__ slots, const double2* __restrict__ rec,
                const unsigned int* __restrict__ over, const unsigned int* __restrict__ n_over,
                unsigned int* __restrict__ counter, const CertOutputs outs) {
  constexpr int CC = NC > 0 ? NC : M3D_MAXC;
  constexpr unsigned FULLM = 0xffffffffu;
  const int C = NC > 0 ? NC : rig.n_cams;
  const int F = cert_record_fields(C);
  const double T1 = thr < init_best ? thr : init_best;
  const int lane = threadIdx.x & 31;
  const unsigned int total = *n_over;
#pragma unroll 1
  for (;;) {
    unsigned int i = 0;
    if (lane == 0) i = atomicAdd(counter, 1u);
    i = __shfl_sync(FULLM, i, 0);
    if (i >= total) break;
    const unsigned int q = over[i];
    const double2* r = rec + ((size_t)(q >> 5) * F) * 32 + (q & 31u);
    XY raw[CC], xh[CC];
    uint32_t badrow[CC];
#pragma unroll
    for (int c = 0; c < CC; ++c) {
      if (c < C) {
        const double2 a = r[(size_t)c * 32], b = r[(size_t)(C + c) * 32];
        raw[c].x = a.x, raw[c].y = a.y;
        xh[c].x = b.x, xh[c].y = b.y;
      }
    }
    const uint4 m0 = reinterpret_cast<const uint4*>(r)[(size_t)(2 * C) * 32];
    const uint32_t v = m0.x & 0xffffu, u = m0.x >> 16, idx = m0.y;
    uint32_t dcur = m0.z;
    int pass = (int)(m0.w & 255u);
    bool have = (m0.w & 256u) != 0;
    const bool is_v2 = (m0.w & 512u) != 0;
    uint32_t best_d = m0.w >> 16;
    const double2 b0 = r[(size_t)(2 * C + 1) * 32], b1 = r[(size_t)(2 * C + 2) * 32];
    double best_err = b0.x, bx = b0.y, by = b1.x, bz = b1.y;
#pragma unroll
    for (int g = 0; g < (CC + 7) / 8; ++g) {
      if (8 * g < C) {
        const uint4 w = reinterpret_cast<const uint4*>(r)[(size_t)(2 * C + 3 + g) * 32];
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (8 * g + 2 * j < CC) badrow[8 * g + 2 * j] = pass == 2 ? 0u : (ww[j] & 0xffffu);
          if (8 * g + 2 * j + 1 < CC) badrow[8 * g + 2 * j + 1] = pass == 2 ? 0u : (ww[j] >> 16);
        }
      }
    }
    // lane-local best of pass 2 (every lane starts from the point's running best)
    double l_err = best_err, lx = bx, ly = by, lz = bz;
    uint32_t l_d = best_d;
    bool l_have = false, stopped = false, ran_out = false;
#pragma unroll 1
    for (;;) {
      if (dcur == 0) {
        if (pass == 1 && T1 < best_err) {
          pass = 2;
#pragma unroll
          for (int p = 0; p < CC; ++p) badrow[p] = 0;
          if (is_v2) {
            // v2 records never recorded the full set's error: it is the first candidate of the arg-min scan
            double X0, Y0, Z0;
            const double e0 = cert_eval<PO, NC>(rig, raw, xh, v, u, X0, Y0, Z0);
            if (lane == 0 && e0 < l_err) l_err = e0, l_d = 0, lx = X0, ly = Y0, lz = Z0, l_have = true;
          }
          dcur = cert_advance<NC>(v, cert_next(v, 0u), badrow, min_cams);
          if (dcur != 0) continue;
        }
        ran_out = true;
        break;
      }
      // lane j: the j-th surviving subset from dcur on (0 once the enumeration is exhausted)
      uint32_t mine = dcur;
      if (pass == 2) {  // no pruning: the next admissible mask
#pragma unroll 1
        for (int t = 0; t < lane && mine != 0; ++t) {
          do {
            mine = cert_next(v, mine);
          } while (mine != 0 && __popc(v & ~mine) < min_cams);
        }
      } else {
#pragma unroll 1
        for (int t = 0; t < lane; ++t)
          if (mine != 0) mine = cert_advance<NC>(v, cert_next(v, mine), badrow, min_cams);
      }
      double X = qnan(), Y = qnan(), Z = qnan(), err = qnan();
      if (mine != 0) {
        const uint32_t kept = v & ~mine;
        err = cert_eval<PO, NC>(rig, raw, xh, kept, kept & u, X, Y, Z);
      }
      if (pass == 1) {
        const uint32_t okb = __ballot_sync(FULLM, mine != 0 && err < T1);
        if (okb) {  // first subset in enumeration order under T1: the reference stops here
          const int src = __ffs(okb) - 1;
          best_err = __shfl_sync(FULLM, err, src);
          bx = __shfl_sync(FULLM, X, src);
          by = __shfl_sync(FULLM, Y, src);
          bz = __shfl_sync(FULLM, Z, src);
          best_d = __shfl_sync(FULLM, mine, src);
          have = true;
          stopped = true;
          break;
        }
      } else if (mine != 0 && err < l_err) {
        l_err = err, l_d = mine, lx = X, ly = Y, lz = Z, l_have = true;
      }
      const uint32_t last = __shfl_sync(FULLM, mine, 31);
      dcur = last != 0 ? cert_advance<NC>(v, cert_next(v, last), badrow, min_cams) : 0u;
    }
    if (!stopped && pass == 2) {
      // strict arg-min over the lanes, the lower step index (= lower d) on ties
      double e = l_have ? l_err : pos_inf();
      uint32_t dd = l_have ? l_d : 0xffffffffu;
      int src = lane;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double e2 = __shfl_xor_sync(FULLM, e, off);
        const uint32_t d2 = __shfl_xor_sync(FULLM, dd, off);
        const int s2 = __shfl_xor_sync(FULLM, src, off);
        if (e2 < e || (e2 == e && d2 < dd)) e = e2, dd = d2, src = s2;
      }
      if (e < best_err) {
        best_err = e;
        best_d = dd;
        bx = __shfl_sync(FULLM, lx, src);
        by = __shfl_sync(FULLM, ly, src);
        bz = __shfl_sync(FULLM, lz, src);
        have = true;
      }
    }
    if (lane == 0) {
      const int k = __popc(v);
      const uint32_t s_sel = cert_step_index<NC>(v, best_d);
      int ne = 1;
      if (ran_out) ne += count_adm(cumb, (1u << k) - 1u, k - min_cams);
      else ne += count_adm(cumb, s_sel, k - min_cams);
      if (outs.p3d) {
        cert_write_point<NC>(outs, C, idx, v, best_d, have, best_err, bx, by, bz, (int32_t)s_sel, ne, raw);
      } else {
        RansacSlot* sl = slots + idx;
        sl->best_err = best_err;
        sl->bx = bx;
        sl->by = by;
        sl->bz = bz;
        sl->masks = __brev(v) >> (32 - C);
        sl->best_s = have ? (int32_t)s_sel : -1;
        sl->neval = ne;
      }
    }
  }
}

#endif  // __CUDACC__

}  // namespace m3d
