// m3d_ransac.cuh — K4: camera-subset RANSAC, i.e. CameraGroup.triangulate_possible with one
// candidate per camera (cameras.py:639-743).
//
// Semantics: visit the subsets of the k valid cameras in
// itertools.product order (step s drops camera V[j] iff bit k-1-j of s is set), skip the
// ones smaller than min_cams unless they are the full set, accept when err < best (best
// starts at init_best), stop when best < thr.  Equivalent closed form used here:
//   T1 = min(thr, init_best);  s* = first admissible s with err(s) < T1 if one exists,
//   else the strict arg-min of err over admissible s (first on ties) if below init_best.
//
// Three kernels per launch (per-point cost varies by two orders of magnitude, so the
// search itself runs on persistent warps that pull work from a global counter):
//   k_ransac_full   thread = point.  Undistort every view once (written to scratch, camera
//                   planes, coalesced), solve the full set (s = 0), rank the cameras by
//                   their residual there ("suspicion order"), write the per-point slot.
//                   ~20 % of the points are decided here.
//   k_ransac_search8 / k_ransac_search16 (m3d_ransac8.cuh, m3d_ransac16.cuh; rigs of <= 8 /
//                   9..16 cameras)  persistent warps, one warp = one point, lane = subset (32
//                   consecutive s per step), points from an atomic counter in batches of 32.
//                   Per step every lane sums the Gram blocks of its subset from per-point tables
//                   and solves; then ONE projection on the subset's most suspicious camera
//                   prunes every subset whose residual already exceeds T * |S| (exact: the
//                   mean cannot come back under T).  The few survivors are scored
//                   cooperatively — lanes = cameras, fixed butterfly sum — in ascending s,
//                   which keeps the sequential accept / stop rule of the reference.
//                   Pass 2 (no subset under T1, rare) rescans with the running best as T.
//   k_ransac_emit   thread = point.  Expands the slots into the reference's outputs
//                   (p3d, errors, picked, points_2d, ...), every plane coalesced.
// Every pruning decision is made on converged fp64 values, so the selected subset is the
// reference's unless an error lands within ~1e-10 px of a threshold (LAPACK's own noise).
#pragma once
#include "m3d_math.cuh"
#include "m3d_point.cuh"

namespace m3d {

constexpr int RANSAC_WARPS = 4;
constexpr int RANSAC_THREADS = RANSAC_WARPS * 32;

__host__ __device__ inline size_t ransac_rig_bytes() { return (sizeof(RigDev) + 15) & ~size_t(15); }

// per-point record handed from k_ransac_full to k_ransac_search to k_ransac_emit
struct RansacSlot {
  double best_err, bx, by, bz;
  // suspicion order (cameras by decreasing residual at the full-set solution), 4 bits each.
  //   C > 8 : nibble r = camera of rank r
  //   C <= 8: low word  "ordl":  nibble r = LOCAL index of the valid camera of rank r,
  //           high word "lrank": nibble b = rank of local camera b
  // (local index b of a valid camera = the bit of the enumeration step s that drops it)
  unsigned long long ord;
  uint32_t masks;  // vmask | umask << 16
  uint32_t vlist;  // C <= 8: nibble b = camera dropped by bit b of s  (= V[k-1-b])
  int32_t best_s, neval;
  int32_t decided;  // 1: nothing left to search
  uint32_t uml;     // C <= 8: usable (post-undistortion) cameras as a mask over local indices
};
static_assert(sizeof(RansacSlot) == 64, "RansacSlot layout");

// Local camera numbering of one point for the <= 8-camera search kernel: local index b of a
// valid camera is the bit of s that drops it (b = 0 is the LAST valid camera).  ord: nibble
// r = physical camera of rank r (all C cameras, the invalid ones last or anywhere).
__device__ __forceinline__ void local_lists(uint32_t vmask, uint32_t umask, unsigned long long ord, int C,
                                            uint32_t& vlist, uint32_t& ordl, uint32_t& lrank, uint32_t& uml) {
  vlist = 0;
  ordl = 0;
  lrank = 0;
  uml = 0;
  int b = 0;
#pragma unroll
  for (int c = 7; c >= 0; --c) {
    if (c < C && ((vmask >> c) & 1u)) {
      vlist |= (uint32_t)c << (4 * b);
      uml |= ((umask >> c) & 1u) << b;
      ++b;
    }
  }
  int rr = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int c = (int)((ord >> (4 * r)) & 15ull);
    if (r < C && ((vmask >> c) & 1u)) {
      const int lb = __popc(vmask >> (c + 1));
      ordl |= (uint32_t)lb << (4 * rr);
      lrank |= (uint32_t)rr << (4 * lb);
      ++rr;
    }
  }
}

// ---------------------------------------------------------------------------------------
// full-set pass: thread = point
// ---------------------------------------------------------------------------------------
template <bool FULL, bool PO, int NC>
__global__ void __launch_bounds__(256, 2)
k_ransac_full(const __grid_constant__ RigDev rig, const double* __restrict__ xy, int64_t ld, int64_t n0,
              int64_t n, int undistort, int min_cams, double thr, double init_best,
              double* __restrict__ U, RansacSlot* __restrict__ slots) {
  // xy: (C, ld, 2) planes, this launch covers points [n0, n0 + n); U: (C, n, 2) receives the
  // undistorted views; slots: (n)
  const int C = NC > 0 ? NC : rig.n_cams;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t vmask = 0, umask = 0;
    unsigned long long ord = 0;
    double best_err = init_best, bx = qnan(), by = qnan(), bz = qnan();
    int32_t best_s = -1;
    bool done = false;
    Gram G;
    gram_zero(G);
    if (NC > 0) {
      double2 raw[NC > 0 ? NC : 1];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        raw[c] = ld_xy(xy, (int64_t)c * ld + n0 + i);
        if (raw[c].x == raw[c].x) vmask |= 1u << c;  // validity on the RAW x (cameras.py:658-659)
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        double x = raw[c].x, y = raw[c].y;
        if ((vmask >> c) & 1u) {
          if (undistort) undistort_point<FULL, PO>(rig.cam[c], raw[c].x, raw[c].y, x, y);
          if (x == x) {  // survives inside triangulate (cameras.py:630)
            umask |= 1u << c;
            gram_add_camera(G, rig.cam[c], x, y);
          }
        }
        st_xy(U, (int64_t)c * n + i, x, y);
      }
      unsigned long long key[NC > 0 ? NC : 1];
#pragma unroll
      for (int c = 0; c < NC; ++c) key[c] = (unsigned long long)c;
      if (__popc(umask) >= 2) {
        double X, Y, Z;
        dlt_solve(G, X, Y, Z);
        double sum = 0.0;
        int m = 0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          if ((vmask >> c) & 1u) {
            double u, v;
            project_point<FULL, PO>(rig.cam[c], X, Y, Z, u, v);
            const double e = residual_norm(raw[c].x - u, raw[c].y - v);
            if (e == e) {
              sum += e;
              ++m;
              // sortable key: residual bits (non-negative double) with the camera id in the
              // four lowest mantissa bits
              key[c] = ((unsigned long long)__double_as_longlong(e) & ~15ull) | (unsigned long long)c;
            }
          }
        }
        const double e0 = (m >= 2) ? sum / (double)m : qnan();
        if (e0 < best_err) {
          best_err = e0;
          best_s = 0;
          bx = X;
          by = Y;
          bz = Z;
          if (e0 < thr) done = true;
        }
      }
      // Batcher odd-even merge sort of the 8 keys, descending
#define M3D_CE(a, b)                                                   \
  {                                                                    \
    const unsigned long long lo__ = key[a] < key[b] ? key[a] : key[b]; \
    const unsigned long long hi__ = key[a] < key[b] ? key[b] : key[a]; \
    key[a] = hi__;                                                     \
    key[b] = lo__;                                                     \
  }
      if (NC == 8) {
        M3D_CE(0, 1) M3D_CE(2, 3) M3D_CE(4, 5) M3D_CE(6, 7)
        M3D_CE(0, 2) M3D_CE(1, 3) M3D_CE(4, 6) M3D_CE(5, 7)
        M3D_CE(1, 2) M3D_CE(5, 6)
        M3D_CE(0, 4) M3D_CE(1, 5) M3D_CE(2, 6) M3D_CE(3, 7)
        M3D_CE(2, 4) M3D_CE(3, 5)
        M3D_CE(1, 2) M3D_CE(3, 4) M3D_CE(5, 6)
      }
#undef M3D_CE
#pragma unroll
      for (int c = 0; c < NC; ++c) ord |= (key[c] & 15ull) << (4 * c);
    } else {
#pragma unroll 1
      for (int c = 0; c < C; ++c) {
        const double2 p = ld_xy(xy, (int64_t)c * ld + n0 + i);
        double x = p.x, y = p.y;
        if (p.x == p.x) {
          vmask |= 1u << c;
          if (undistort) undistort_point<FULL, PO>(rig.cam[c], p.x, p.y, x, y);
          if (x == x) {
            umask |= 1u << c;
            gram_add_camera(G, rig.cam[c], x, y);
          }
        }
        st_xy(U, (int64_t)c * n + i, x, y);
        ord |= (unsigned long long)c << (4 * c);  // identity order
      }
      if (__popc(umask) >= 2) {
        double X, Y, Z;
        dlt_solve(G, X, Y, Z);
        double sum = 0.0;
        int m = 0;
        for (uint32_t rest = vmask; rest; rest &= rest - 1) {
          const int c = __ffs(rest) - 1;
          const double2 p = ld_xy(xy, (int64_t)c * ld + n0 + i);
          double u, v;
          project_point<FULL, PO>(rig.cam[c], X, Y, Z, u, v);
          const double e = residual_norm(p.x - u, p.y - v);
          if (e == e) {
            sum += e;
            ++m;
          }
        }
        const double e0 = (m >= 2) ? sum / (double)m : qnan();
        if (e0 < best_err) {
          best_err = e0;
          best_s = 0;
          bx = X;
          by = Y;
          bz = Z;
          if (e0 < thr) done = true;
        }
      }
    }
    const int k = __popc(vmask);
    if (k < 2 || k <= min_cams) done = true;  // every smaller subset would be skipped
    RansacSlot sl;
    sl.best_err = best_err;
    sl.bx = bx;
    sl.by = by;
    sl.bz = bz;
    if (C <= 8) {
      uint32_t vlist, ordl, lrank, uml;
      local_lists(vmask, umask, ord, NC > 0 ? NC : C, vlist, ordl, lrank, uml);
      sl.ord = (unsigned long long)ordl | ((unsigned long long)lrank << 32);
      sl.vlist = vlist;
      sl.uml = uml;
    } else {
      sl.ord = ord;
      sl.vlist = 0;
      sl.uml = 0;
    }
    sl.masks = vmask | (umask << 16);
    sl.best_s = best_s;
    sl.neval = 1;  // the full set is always tried (cameras.py:691)
    sl.decided = done ? 1 : 0;
    slots[i] = sl;
  }
}

// ---------------------------------------------------------------------------------------
// outputs: thread = point, every plane coalesced
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_ransac_emit(int C, const double* __restrict__ xy, int64_t ld, int64_t n0, int64_t n,
              const RansacSlot* __restrict__ slots, double* __restrict__ p3d, uint8_t* __restrict__ picked,
              double* __restrict__ xy_picked, double* __restrict__ err_out,
              int32_t* __restrict__ subset_out, int32_t* __restrict__ neval_out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const RansacSlot sl = slots[i];
    const int64_t o = n0 + i;
    const uint32_t vmask = sl.masks & 0xffffu;
    const uint32_t best_mask = sl.best_s >= 0 ? subset_mask(vmask, __popc(vmask), (uint32_t)sl.best_s) : 0u;
    p3d[3 * o] = sl.bx;
    p3d[3 * o + 1] = sl.by;
    p3d[3 * o + 2] = sl.bz;
    err_out[o] = (sl.best_s >= 0) ? sl.best_err : 0.0;  // errors default to 0.0 (cameras.py:675)
    if (subset_out) subset_out[o] = sl.best_s;
    if (neval_out) neval_out[o] = sl.neval;
    if (picked || xy_picked) {
#pragma unroll 1
      for (int c = 0; c < C; ++c) {
        const bool in = (best_mask >> c) & 1u;
        if (picked) picked[(int64_t)c * ld + o] = in ? 1 : 0;
        if (xy_picked) {
          double2 q = make_double2(qnan(), qnan());
          if (in) q = ld_xy(xy, (int64_t)c * ld + o);
          st_xy(xy_picked, (int64_t)c * ld + o, q.x, q.y);
        }
      }
    }
  }
}

}  // namespace m3d
