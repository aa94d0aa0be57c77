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
//   k_ransac_search (rigs of 9..16 cameras; rigs of <= 8 cameras use k_ransac_search8,
//                   m3d_ransac8.cuh)  persistent warps; a GROUP of GS = 16 lanes (>= cameras) =
//                   one point, lane = subset (GS consecutive s per step), 32 / GS points in
//                   flight per warp; idle groups take the next undecided point of the
//                   warp's current 32-point batch, batches come from an atomic counter.
//                   Per step every lane sums the per-camera Gram blocks of its subset (low
//                   log2(GS) cameras pre-summed per lane, the rest group-uniform) and
//                   solves; then ONE projection on the subset's most suspicious camera
//                   prunes every subset whose residual already exceeds T * |S| (exact: the
//                   mean cannot come back under T).  The few survivors are scored
//                   cooperatively — lane = camera, fixed butterfly sum — in ascending s,
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

// shared memory of k_ransac_search: rig | per warp, per group: raw[C][2] | gc[C] | glow[10][GS]
// (+64 B so that consecutive groups start 16 banks apart: a 64-bit access of 4 groups x 8
// lanes then takes the minimum two wavefronts)
__host__ __device__ inline size_t ransac_group_bytes(int C, int GS) {
  size_t b = (size_t)C * 16 + (size_t)C * sizeof(Gram) + (size_t)10 * GS * 8;
  b = (b + 127) & ~size_t(127);
  return b + 64;
}
inline size_t ransac_smem_bytes(int C, int GS) {
  return ransac_rig_bytes() + (size_t)RANSAC_WARPS * (32 / GS) * ransac_group_bytes(C, GS);
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
// subset search: persistent warps, group of GS lanes = point, lane = subset
// ---------------------------------------------------------------------------------------
template <bool FULL, bool PO, int GS, int MINB>
__global__ void __launch_bounds__(RANSAC_THREADS, MINB)
k_ransac_search(const RigDev* __restrict__ rig_g, const double* __restrict__ xy, int64_t ld, int64_t n0,
                int64_t n, int min_cams, double thr, double init_best, const double* __restrict__ U,
                RansacSlot* __restrict__ slots, unsigned long long* __restrict__ counter) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr unsigned FULLM = 0xffffffffu;
  constexpr int NG = 32 / GS;  // points in flight per warp
  constexpr int LOGGS = GS == 8 ? 3 : (GS == 16 ? 4 : 5);
  constexpr uint32_t GM = GS == 32 ? 0xffffffffu : ((1u << GS) - 1u);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / GS, j = lane % GS;  // group and lane-in-group
  const int gshift = g * GS;
  RigDev& srig = *reinterpret_cast<RigDev*>(smem);
  {
    // rig copy for per-lane camera indexing (constant-bank reads with lane-varying addresses
    // would serialise)
    const double* src = reinterpret_cast<const double*>(rig_g);
    double* dst = reinterpret_cast<double*>(&srig);
    for (int i = threadIdx.x; i < (int)(sizeof(RigDev) / 8); i += RANSAC_THREADS) dst[i] = src[i];
  }
  __syncthreads();
  const int C = srig.n_cams;
  unsigned char* gbase = smem + ransac_rig_bytes() + (size_t)(warp * NG + g) * ransac_group_bytes(C, GS);
  double* raws = reinterpret_cast<double*>(gbase);                              // [C][2] raw pixels
  Gram* gcs = reinterpret_cast<Gram*>(gbase + (size_t)C * 16);                  // [C] Gram blocks
  double* glow = reinterpret_cast<double*>(gbase + (size_t)C * 16 + (size_t)C * sizeof(Gram));  // [10][GS]
  const double T1 = thr < init_best ? thr : init_best;

  uint32_t todo = 0;      // undecided points of the current batch
  int64_t batch0 = 0;     // first point of the current batch
  bool exhausted = false;
  // group state, replicated in the lanes of the group
  int64_t cur = -1;       // point the group works on
  bool fresh = false;
  uint32_t vm = 0, um = 0, vhigh = 0, cm_low = 0, base = 0, n_sub = 0;
  unsigned long long ordp = 0, hcam = 0;  // hcam: nibble b = camera dropped by bit b of s >> log2(GS)
  int k = 0, khigh = 0, pass = 1;
  int32_t ne = 0;
  double rb = T1;
#pragma unroll 1
  for (;;) {
    // hand undecided points to idle groups; fetch a new batch of 32 points when the current
    // one is used up (warp-uniform)
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) {
      const bool idle = __shfl_sync(FULLM, cur < 0 ? 1 : 0, gi * GS) != 0;
      if (!idle) continue;
      while (!todo && !exhausted) {
        unsigned long long b = 0;
        if (lane == 0) b = atomicAdd(counter, 32ull);
        b = __shfl_sync(FULLM, b, 0);
        if ((int64_t)b >= n) {
          exhausted = true;
        } else {
          batch0 = (int64_t)b;
          const int64_t i = batch0 + lane;
          const bool open = (i < n) && (slots[i].decided == 0);
          todo = __ballot_sync(FULLM, open);
        }
      }
      if (todo) {
        const int p = __ffs(todo) - 1;
        todo &= todo - 1;
        if (g == gi) {
          cur = batch0 + p;
          fresh = true;
        }
      }
    }
    if (!__any_sync(FULLM, cur >= 0)) break;
    if (__any_sync(FULLM, fresh)) {
      if (fresh) {
        const RansacSlot* sl = slots + cur;
        vm = sl->masks & 0xffffu;
        um = sl->masks >> 16;
        ordp = sl->ord;
        k = __popc(vm);
        n_sub = 1u << k;
        if (j < C) {  // lane = camera: raw pixels and Gram block of the group's point
          const double2 q = ld_xy(xy, (int64_t)j * ld + n0 + cur);
          raws[2 * j] = q.x;
          raws[2 * j + 1] = q.y;
          Gram gg;
          gram_zero(gg);
          if ((um >> j) & 1u) {
            const double2 u = ld_xy(U, (int64_t)j * n + cur);
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
        hcam = 0;
        {
          int jj = 0;  // V[jj] (ascending) is dropped by bit khigh-1-jj
          for (uint32_t rest = vhigh; rest; rest &= rest - 1, ++jj)
            hcam |= (unsigned long long)(__ffs(rest) - 1) << (4 * (khigh - 1 - jj));
        }
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
      uint32_t dropped = 0;
      for (uint32_t h = base >> LOGGS, b = 0; h; h >>= 1, ++b)
        if (h & 1u) dropped |= 1u << (uint32_t)((hcam >> (4 * b)) & 15ull);
      cm = (vhigh & ~dropped) | cm_low;
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
          RansacSlot* sl = slots + cur;
          sl->best_err = el;
          sl->best_s = (int32_t)(base + l);
          sl->bx = Xl;
          sl->by = Yl;
          sl->bz = Zl;
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
          rb = slots[cur].best_err;  // init_best or the full-set error (>= T1)
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
