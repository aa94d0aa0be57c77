// m3d_ransac.cuh — K4: camera-subset RANSAC, i.e. CameraGroup.triangulate_possible with one
// candidate per camera (cameras.py:639-743), as one fused kernel.
//
// Semantics: visit the subsets of the k valid cameras in
// itertools.product order (step s drops camera V[j] iff bit k-1-j of s is set), skip the
// ones smaller than min_cams unless they are the full set, accept when err < best (best
// starts at init_best), stop when best < thr.  Equivalent closed form used here:
//   T1 = min(thr, init_best);  s* = first admissible s with err(s) < T1 if one exists,
//   else the strict arg-min of err over admissible s (first on ties) if below init_best.
//
// Schedule
//   phase A  lane = point: undistort every view once (kept in shared memory), solve the
//            full set (s = 0), rank the cameras by their residual at that solution
//            ("suspicion order").  ~20 % of the points finish here.
//   phase B  a GROUP of GS lanes (8 for rigs of up to 8 cameras) = one point, lane = subset
//            (GS consecutive s per step), 32 / GS points in flight per warp; a group that
//            finishes its point takes the next unfinished one of the warp's tile.  Per step
//            every lane sums the per-camera Gram blocks of its subset (low log2(GS) cameras
//            pre-summed per lane, the rest group-uniform) and solves; then ONE projection on
//            the subset's most suspicious camera prunes every subset whose residual already
//            exceeds T * |S| (exact: the mean cannot come back under T).  The few survivors
//            are scored cooperatively — lane = camera, fixed butterfly sum — in ascending
//            s, which keeps the sequential accept / stop rule of the reference.  Pass 2 (no
//            subset under T1, rare) repeats the scan with the running best as T.
// Every pruning decision is made on converged fp64 values, so the selected subset is the
// reference's unless an error lands within ~1e-10 px of a threshold (LAPACK's own noise).
#pragma once
#include "m3d_math.cuh"
#include "m3d_point.cuh"

namespace m3d {

constexpr int RANSAC_WARPS = 4;
constexpr int RANSAC_THREADS = RANSAC_WARPS * 32;

__host__ __device__ inline size_t ransac_rig_bytes() { return (sizeof(RigDev) + 15) & ~size_t(15); }

// per-point record shared between phase A (lane = point) and phase B (group = point)
struct RansacSlot {
  double best_err, bx, by, bz;
  unsigned long long ord;
  uint32_t vmask, umask, best_mask;
  int32_t best_s, neval, pad;
};
static_assert(sizeof(RansacSlot) == 64, "RansacSlot layout");

// per warp: U[C][32] double2 | slots[32] | per group (32/GS of them): raw[C][2] | gc[C] | glow[10][GS]
__host__ __device__ inline size_t ransac_group_bytes(int C, int GS) {
  return (size_t)C * 16 + (size_t)C * sizeof(Gram) + (size_t)10 * GS * 8;
}
__host__ __device__ inline size_t ransac_warp_bytes(int C, int GS) {
  return (size_t)C * 32 * 16 + 32 * sizeof(RansacSlot) + (size_t)(32 / GS) * ransac_group_bytes(C, GS);
}
inline size_t ransac_smem_bytes(int C, int GS) {
  return ransac_rig_bytes() + RANSAC_WARPS * ransac_warp_bytes(C, GS);
}

// next camera of subset cm in suspicion order, starting at position pos (returns -1 when the
// subset is exhausted)
__device__ __forceinline__ int next_member(unsigned long long ord, int C, uint32_t cm, int& pos) {
  while (pos < C) {
    const int c = (int)((ord >> (4 * pos)) & 15ull);
    ++pos;
    if ((cm >> c) & 1u) return c;
  }
  return -1;
}

template <bool FULL, bool PO, int NC, int GS, int MINB>
__global__ void __launch_bounds__(RANSAC_THREADS, MINB)
k_ransac(const __grid_constant__ RigDev rig, const RigDev* __restrict__ rig_g,
         const double* __restrict__ xy, int64_t N, int undistort, int min_cams, double thr,
         double init_best, double* __restrict__ p3d, uint8_t* __restrict__ picked,
         double* __restrict__ xy_picked, double* __restrict__ err_out,
         int32_t* __restrict__ subset_out, int32_t* __restrict__ neval_out) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr unsigned FULLM = 0xffffffffu;
  constexpr int NG = 32 / GS;            // points in flight per warp
  constexpr int LOGGS = GS == 8 ? 3 : (GS == 16 ? 4 : 5);
  const int C = NC > 0 ? NC : rig.n_cams;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / GS, j = lane % GS;  // group and lane-in-group
  const int gshift = g * GS;
  constexpr uint32_t GM = GS == 32 ? 0xffffffffu : ((1u << GS) - 1u);
  RigDev& srig = *reinterpret_cast<RigDev*>(smem);
  unsigned char* wbase = smem + ransac_rig_bytes() + (size_t)warp * ransac_warp_bytes(C, GS);
  double2* Us = reinterpret_cast<double2*>(wbase);                              // [C][32] undistorted
  RansacSlot* slots = reinterpret_cast<RansacSlot*>(wbase + (size_t)C * 512);   // [32]
  unsigned char* gbase = wbase + (size_t)C * 512 + 32 * sizeof(RansacSlot) + (size_t)g * ransac_group_bytes(C, GS);
  double* raws = reinterpret_cast<double*>(gbase);                              // [C][2] raw, group's point
  Gram* gcs = reinterpret_cast<Gram*>(gbase + (size_t)C * 16);                  // [C]
  double* glow = reinterpret_cast<double*>(gbase + (size_t)C * 16 + (size_t)C * sizeof(Gram));  // [10][GS]

  // rig copy for per-lane camera indexing (constant-bank reads with lane-varying addresses
  // would serialise)
  {
    const double* src = reinterpret_cast<const double*>(rig_g);
    double* dst = reinterpret_cast<double*>(&srig);
    for (int i = threadIdx.x; i < (int)(sizeof(RigDev) / 8); i += RANSAC_THREADS) dst[i] = src[i];
  }
  __syncthreads();

  const int64_t tile0 = ((int64_t)blockIdx.x * RANSAC_WARPS + warp) * 32;
  if (tile0 >= N) return;
  const int64_t n = tile0 + lane;
  const bool inb = n < N;
  const double T1 = thr < init_best ? thr : init_best;

  // ---- phase A ---------------------------------------------------------------------------
  uint32_t vmask = 0, umask = 0;
  unsigned long long ord = 0;
  double best_err = init_best, bx = qnan(), by = qnan(), bz = qnan();
  int32_t best_s = -1, neval = 0;
  uint32_t best_mask = 0;
  bool done = !inb;
  if (NC > 0) {
    double2 raw[NC > 0 ? NC : 1];
    Gram G;
    gram_zero(G);
    if (inb) {
#pragma unroll
      for (int c = 0; c < NC; ++c) raw[c] = ld_xy(xy, (int64_t)c * N + n);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        double x = raw[c].x, y = raw[c].y;
        if (raw[c].x == raw[c].x) {  // validity on the RAW x (cameras.py:658-659)
          vmask |= 1u << c;
          if (undistort) undistort_point<FULL, PO>(rig.cam[c], raw[c].x, raw[c].y, x, y);
          if (x == x) {  // survives inside triangulate (cameras.py:630)
            umask |= 1u << c;
            gram_add_camera(G, rig.cam[c], x, y);
          }
        }
        Us[c * 32 + lane] = make_double2(x, y);
      }
      neval = 1;  // the full set is always tried (cameras.py:691)
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
          best_mask = vmask;
          bx = X;
          by = Y;
          bz = Z;
          if (e0 < thr) done = true;
        }
      }
      const int k = __popc(vmask);
      if (k < 2 || k <= min_cams) done = true;  // every smaller subset would be skipped
      if (!done) {
        // Batcher odd-even merge sort, descending
#define M3D_CE(a, b)                         \
  {                                          \
    const unsigned long long lo__ = key[a] < key[b] ? key[a] : key[b]; \
    const unsigned long long hi__ = key[a] < key[b] ? key[b] : key[a]; \
    key[a] = hi__;                           \
    key[b] = lo__;                           \
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
        for (int i = 0; i < NC; ++i) ord |= (key[i] & 15ull) << (4 * i);
      }
    }
  } else {
    if (inb) {
      Gram G;
      gram_zero(G);
#pragma unroll 1
      for (int c = 0; c < C; ++c) {
        const double2 p = ld_xy(xy, (int64_t)c * N + n);
        double x = p.x, y = p.y;
        if (p.x == p.x) {
          vmask |= 1u << c;
          if (undistort) undistort_point<FULL, PO>(rig.cam[c], p.x, p.y, x, y);
          if (x == x) {
            umask |= 1u << c;
            gram_add_camera(G, rig.cam[c], x, y);
          }
        }
        Us[c * 32 + lane] = make_double2(x, y);
        ord |= (unsigned long long)c << (4 * c);  // identity order
      }
      neval = 1;
      if (__popc(umask) >= 2) {
        double X, Y, Z;
        dlt_solve(G, X, Y, Z);
        double sum = 0.0;
        int m = 0;
        for (uint32_t rest = vmask; rest; rest &= rest - 1) {
          const int c = __ffs(rest) - 1;
          const double2 p = ld_xy(xy, (int64_t)c * N + n);
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
          best_mask = vmask;
          bx = X;
          by = Y;
          bz = Z;
          if (e0 < thr) done = true;
        }
      }
      const int k = __popc(vmask);
      if (k < 2 || k <= min_cams) done = true;
    }
  }
  {
    RansacSlot sl;
    sl.best_err = best_err;
    sl.bx = bx;
    sl.by = by;
    sl.bz = bz;
    sl.ord = ord;
    sl.vmask = vmask;
    sl.umask = umask;
    sl.best_mask = best_mask;
    sl.best_s = best_s;
    sl.neval = neval;
    sl.pad = 0;
    slots[lane] = sl;
  }
  __syncwarp();

  // ---- phase B ---------------------------------------------------------------------------
  uint32_t todo = __ballot_sync(FULLM, !done);
  // group state, replicated in the lanes of the group
  int cur = -1;            // point (lane index in the tile) the group works on
  bool fresh = false;
  uint32_t vm = 0, um = 0, vhigh = 0, cm_low = 0, base = 0, n_sub = 0;
  unsigned long long ordp = 0;
  int k = 0, khigh = 0, pass = 1;
  int32_t ne = 0;
  double rb = T1;
#pragma unroll 1
  for (;;) {
    // hand unfinished points to idle groups (warp-uniform)
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) {
      const int cg = __shfl_sync(FULLM, cur, gi * GS);
      if (cg < 0 && todo) {
        const int p = __ffs(todo) - 1;
        todo &= todo - 1;
        if (g == gi) {
          cur = p;
          fresh = true;
        }
      }
    }
    if (!__any_sync(FULLM, cur >= 0)) break;
    if (__any_sync(FULLM, fresh)) {
      if (fresh) {
        const RansacSlot& sl = slots[cur];
        vm = sl.vmask;
        um = sl.umask;
        ordp = sl.ord;
        k = __popc(vm);
        n_sub = 1u << k;
        if (j < C) {  // lane = camera: raw pixels and Gram block of the group's point
          const double2 q = ld_xy(xy, (int64_t)j * N + tile0 + cur);
          raws[2 * j] = q.x;
          raws[2 * j + 1] = q.y;
          Gram gg;
          gram_zero(gg);
          if ((um >> j) & 1u) {
            const double2 u = Us[j * 32 + cur];
            gram_add_camera(gg, srig.cam[j], u.x, u.y);
          }
          gcs[j] = gg;
        }
      }
      __syncwarp();
      if (fresh) {
        // two-level subset structure of one GS-subset step: the low log2(GS) bits of s (the
        // LAST valid cameras) vary across the lanes of the group, the high bits are shared
        const int klow = k < LOGGS ? k : LOGGS;
        uint32_t vlow = vm;
        for (int i = 0; i < k - klow; ++i) vlow &= vlow - 1;
        vhigh = vm & ~vlow;
        khigh = k - klow;
        cm_low = subset_mask(vlow, klow, (uint32_t)j & ((1u << klow) - 1u));
        Gram gg;
        gram_zero(gg);
        for (uint32_t rest = cm_low & um; rest; rest &= rest - 1) gram_add(gg, gcs[__ffs(rest) - 1]);
#pragma unroll
        for (int i = 0; i < 6; ++i) glow[i * GS + j] = gg.h[i];
        glow[6 * GS + j] = gg.g[0];
        glow[7 * GS + j] = gg.g[1];
        glow[8 * GS + j] = gg.g[2];
        glow[9 * GS + j] = gg.w;
        base = 0;
        pass = 1;
        rb = T1;
        ne = 0;
        fresh = false;
      }
      __syncwarp();
    }

    // ---- one step: GS consecutive subsets per active group
    const bool act = cur >= 0;
    const uint32_t s = base + (uint32_t)j;
    uint32_t cm = 0;
    bool adm = false;
    if (act && s >= 1 && s < n_sub) {
      cm = subset_mask(vhigh, khigh, base >> LOGGS) | cm_low;
      const int cnt = __popc(cm);
      adm = (cnt >= min_cams) || (cnt == k);
    }
    const uint32_t admb = (__ballot_sync(FULLM, adm) >> gshift) & GM;
    if (pass == 1) ne += __popc(admb);
    double X = qnan(), Y = qnan(), Z = qnan();
    bool alive = adm && (__popc(cm & um) >= 2);
    if (alive) {
      Gram G;
#pragma unroll
      for (int i = 0; i < 6; ++i) G.h[i] = glow[i * GS + j];
      G.g[0] = glow[6 * GS + j];
      G.g[1] = glow[7 * GS + j];
      G.g[2] = glow[8 * GS + j];
      G.w = glow[9 * GS + j];
      for (uint32_t rest = cm & ~cm_low & um; rest; rest &= rest - 1) gram_add(G, gcs[__ffs(rest) - 1]);
      dlt_solve(G, X, Y, Z);
      alive = (X == X);
    }
    // one pruning round on the most suspicious camera of the subset
    if (alive) {
      int pos = 0;
      const int c = next_member(ordp, C, cm, pos);
      double u, v;
      project_point<FULL, PO>(srig.cam[c], X, Y, Z, u, v);
      const double e = residual_norm(raws[2 * c] - u, raws[2 * c + 1] - v);
      const double limit = rb * (double)__popc(cm) * (1.0 + 1e-12);
      if (e > limit) alive = false;  // mean >= e / |S| > T: can never be accepted
    }
    // survivors of every group, in ascending s: exact mean with lane-in-group = camera
    uint32_t cand = (__ballot_sync(FULLM, alive) >> gshift) & GM;
    bool finished = false;  // the group's point is decided
    while (__any_sync(FULLM, cand != 0)) {
      const bool has = cand != 0;
      const int l = has ? __ffs(cand) - 1 : 0;
      cand &= cand - 1;
      const double Xl = __shfl_sync(FULLM, X, gshift + l), Yl = __shfl_sync(FULLM, Y, gshift + l),
                   Zl = __shfl_sync(FULLM, Z, gshift + l);
      const uint32_t cml = __shfl_sync(FULLM, cm, gshift + l);
      double e = qnan();
      if (has && j < C && ((cml >> j) & 1u)) {
        double u, v;
        project_point<FULL, PO>(srig.cam[j], Xl, Yl, Zl, u, v);
        e = residual_norm(raws[2 * j] - u, raws[2 * j + 1] - v);
      }
      // fixed-shape butterfly over the camera lanes: NaN residuals count as 0 and drop out of
      // the denominator (cameras.py:771-775); for 8 cameras this is numpy's pairwise order
      const int m = __popc((__ballot_sync(FULLM, e == e) >> gshift) & GM);
      double sum = (e == e) ? e : 0.0;
#pragma unroll
      for (int off = 1; off < GS; off <<= 1) sum += __shfl_xor_sync(FULLM, sum, off);
      const double el = (m >= 2) ? sum / (double)m : qnan();
      if (has && el < rb) {
        if (j == 0) {
          RansacSlot& sl = slots[cur];
          sl.best_err = el;
          sl.best_s = (int32_t)(base + l);
          sl.best_mask = cml;
          sl.bx = Xl;
          sl.by = Yl;
          sl.bz = Zl;
        }
        if (pass == 1) {
          // first subset under T1: the reference stops here; later subsets of this step were
          // never evaluated by it
          ne -= __popc(admb & ~(0xffffffffu >> (31 - l)));
          finished = true;
          cand = 0;
        } else {
          rb = el;  // pass 2: sequential arg-min
        }
      }
    }
    if (act && !finished) {
      base += GS;
      if (base >= n_sub) {
        if (pass == 1) {  // nothing under T1: rescan for the strict arg-min
          pass = 2;
          base = 0;
          rb = slots[cur].best_err;
        } else {
          finished = true;
        }
      }
    }
    if (act && finished) {
      if (j == 0) slots[cur].neval += ne;
      cur = -1;
    }
    __syncwarp();
  }
  __syncwarp();

  // ---- outputs: lane = point again, coalesced per plane ---------------------------------------
  if (inb) {
    const RansacSlot sl = slots[lane];
    p3d[3 * n] = sl.bx;
    p3d[3 * n + 1] = sl.by;
    p3d[3 * n + 2] = sl.bz;
    err_out[n] = (sl.best_s >= 0) ? sl.best_err : 0.0;  // errors default to 0.0 (cameras.py:675)
    if (subset_out) subset_out[n] = sl.best_s;
    if (neval_out) neval_out[n] = sl.neval;
    if (picked || xy_picked) {
#pragma unroll 1
      for (int c = 0; c < C; ++c) {
        const bool in = (sl.best_mask >> c) & 1u;
        if (picked) picked[(int64_t)c * N + n] = in ? 1 : 0;
        if (xy_picked) {
          double2 q = make_double2(qnan(), qnan());
          if (in) q = ld_xy(xy, (int64_t)c * N + n);
          st_xy(xy_picked, (int64_t)c * N + n, q.x, q.y);
        }
      }
    }
  }
}

}  // namespace m3d
