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
//   phase B  warp = point, lane = subset (32 consecutive s per step).  Every lane sums the
//            per-camera Gram blocks of its subset and solves; then ONE projection round on
//            the most suspicious camera of the subset prunes every subset whose partial
//            residual sum already exceeds T * |S| (exact: the mean cannot come back under
//            T).  The few survivors are scored cooperatively — lane = camera, ordered
//            shuffle sum — in ascending s, which keeps the sequential accept / stop rule
//            of the reference.  Pass 2 (no subset under T1, rare) repeats the scan with
//            the running best as T.
// Every pruning decision is made on converged fp64 values, so the selected subset is the
// reference's unless an error lands within ~1e-10 px of a threshold (LAPACK's own noise).
#pragma once
#include "m3d_math.cuh"
#include "m3d_point.cuh"

namespace m3d {

constexpr int RANSAC_WARPS = 4;
constexpr int RANSAC_THREADS = RANSAC_WARPS * 32;

__host__ __device__ inline size_t ransac_rig_bytes() { return (sizeof(RigDev) + 15) & ~size_t(15); }
// per warp: U[C][32] double2 | raw[C][2] | gc[C] Gram | glow[10][32] doubles
__host__ __device__ inline size_t ransac_warp_bytes(int C) {
  return (size_t)C * 32 * 16 + (size_t)C * 16 + (size_t)C * sizeof(Gram) + 10 * 32 * 8;
}
inline size_t ransac_smem_bytes(int C) { return ransac_rig_bytes() + RANSAC_WARPS * ransac_warp_bytes(C); }

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

template <bool FULL, bool PO, int NC, int MINB>
__global__ void __launch_bounds__(RANSAC_THREADS, MINB)
k_ransac(const __grid_constant__ RigDev rig, const RigDev* __restrict__ rig_g,
         const double* __restrict__ xy, int64_t N, int undistort, int min_cams, double thr,
         double init_best, double* __restrict__ p3d, uint8_t* __restrict__ picked,
         double* __restrict__ xy_picked, double* __restrict__ err_out,
         int32_t* __restrict__ subset_out, int32_t* __restrict__ neval_out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const unsigned FULLM = 0xffffffffu;
  const int C = NC > 0 ? NC : rig.n_cams;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  RigDev& srig = *reinterpret_cast<RigDev*>(smem);
  unsigned char* wbase = smem + ransac_rig_bytes() + (size_t)warp * ransac_warp_bytes(C);
  double2* Us = reinterpret_cast<double2*>(wbase);                    // [C][32] undistorted
  double* raws = reinterpret_cast<double*>(wbase + (size_t)C * 512);  // [C][2] raw, current point
  Gram* gcs = reinterpret_cast<Gram*>(wbase + (size_t)C * 512 + (size_t)C * 16);  // [C]
  double* glow = reinterpret_cast<double*>(wbase + (size_t)C * 512 + (size_t)C * 16 + (size_t)C * sizeof(Gram));

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
  __syncwarp();

  // ---- phase B ---------------------------------------------------------------------------
  uint32_t todo = __ballot_sync(FULLM, !done);
  while (todo) {
    const int p = __ffs(todo) - 1;
    todo &= todo - 1;
    const int64_t np = tile0 + p;
    const uint32_t vm = __shfl_sync(FULLM, vmask, p);
    const uint32_t um = __shfl_sync(FULLM, umask, p);
    const unsigned long long ordp = __shfl_sync(FULLM, ord, p);
    const int k = __popc(vm);
    const uint32_t n_sub = 1u << k;
    __syncwarp();
    if (lane < C) {
      const double2 q = ld_xy(xy, (int64_t)lane * N + np);
      raws[2 * lane] = q.x;
      raws[2 * lane + 1] = q.y;
      Gram g;
      gram_zero(g);
      if ((um >> lane) & 1u) {
        const double2 u = Us[lane * 32 + p];
        gram_add_camera(g, srig.cam[lane], u.x, u.y);
      }
      gcs[lane] = g;
    }
    __syncwarp();

    // Two-level subset structure of one 32-subset step: the low min(k,5) bits of s (the
    // LAST valid cameras) vary across lanes, the high bits (the first valid cameras) are
    // shared by the whole step.  Per lane, once per point: camera mask and Gram block of
    // the low part.
    const int klow = k < 5 ? k : 5;
    uint32_t vlow = 0;  // the klow highest-index valid cameras
    {
      uint32_t rest = vm;
      for (int j = 0; j < k - klow; ++j) rest &= rest - 1;
      vlow = rest;
    }
    const uint32_t vhigh = vm & ~vlow;
    const int khigh = k - klow;
    const uint32_t cm_low = subset_mask(vlow, klow, (uint32_t)lane & ((1u << klow) - 1u));
    {
      Gram g;
      gram_zero(g);
      for (uint32_t rest = cm_low & um; rest; rest &= rest - 1) gram_add(g, gcs[__ffs(rest) - 1]);
#pragma unroll
      for (int i = 0; i < 6; ++i) glow[i * 32 + lane] = g.h[i];
      glow[6 * 32 + lane] = g.g[0];
      glow[7 * 32 + lane] = g.g[1];
      glow[8 * 32 + lane] = g.g[2];
      glow[9 * 32 + lane] = g.w;
    }
    __syncwarp();

    double rb = T1;  // pass 1: fixed threshold T1; pass 2: running best
    bool found = false;
    int32_t ne = 0;
#pragma unroll 1
    for (int pass = 1; pass <= 2 && !found; ++pass) {
      if (pass == 2) rb = __shfl_sync(FULLM, best_err, p);
#pragma unroll 1
      for (uint32_t base = 0; base < n_sub && !found; base += 32) {
        const uint32_t s = base + lane;
        const uint32_t cm_high = subset_mask(vhigh, khigh, base >> 5);  // warp-uniform
        uint32_t cm = 0;
        bool adm = false;
        if (s >= 1 && s < n_sub) {
          cm = cm_high | cm_low;
          const int cnt = __popc(cm);
          adm = (cnt >= min_cams) || (cnt == k);
        }
        if (pass == 1) ne += __popc(__ballot_sync(FULLM, adm));
        // solve
        double X = qnan(), Y = qnan(), Z = qnan();
        bool alive = adm && (__popc(cm & um) >= 2);
        if (alive) {
          Gram G;
#pragma unroll
          for (int i = 0; i < 6; ++i) G.h[i] = glow[i * 32 + lane];
          G.g[0] = glow[6 * 32 + lane];
          G.g[1] = glow[7 * 32 + lane];
          G.g[2] = glow[8 * 32 + lane];
          G.w = glow[9 * 32 + lane];
          for (uint32_t rest = cm_high & um; rest; rest &= rest - 1) gram_add(G, gcs[__ffs(rest) - 1]);
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
        // survivors, in ascending s: exact mean with lane = camera
        uint32_t cand = __ballot_sync(FULLM, alive);
        while (cand) {
          const int l = __ffs(cand) - 1;
          cand &= cand - 1;
          const double Xl = __shfl_sync(FULLM, X, l), Yl = __shfl_sync(FULLM, Y, l),
                       Zl = __shfl_sync(FULLM, Z, l);
          const uint32_t cml = __shfl_sync(FULLM, cm, l);
          double e = qnan();
          if (lane < C && ((cml >> lane) & 1u)) {
            double u, v;
            project_point<FULL, PO>(srig.cam[lane], Xl, Yl, Zl, u, v);
            e = residual_norm(raws[2 * lane] - u, raws[2 * lane + 1] - v);
          }
          // fixed-shape butterfly over the camera lanes: NaN residuals count as 0 and drop out
          // of the denominator (cameras.py:771-775); for 8 cameras this is numpy's pairwise
          // order ((e0+e1)+(e2+e3))+((e4+e5)+(e6+e7))
          const int m = __popc(__ballot_sync(FULLM, e == e));
          double sum = (e == e) ? e : 0.0;
#pragma unroll
          for (int off = 1; off < M3D_MAXC; off <<= 1) sum += __shfl_xor_sync(FULLM, sum, off);
          sum = __shfl_sync(FULLM, sum, 0);
          const double el = (m >= 2) ? sum / (double)m : qnan();
          if (el < rb) {
            if (lane == p) {
              best_err = el;
              best_s = (int32_t)(base + l);
              best_mask = cml;
              bx = Xl;
              by = Yl;
              bz = Zl;
            }
            if (pass == 1) {
              // first subset under T1: the reference stops here; later lanes of this step
              // were never evaluated by it
              const uint32_t admb = __ballot_sync(FULLM, adm);
              ne -= __popc(admb & ~(0xffffffffu >> (31 - l)));
              found = true;
              break;
            }
            rb = el;  // pass 2: sequential arg-min
          }
        }
      }
    }
    if (lane == p) neval += ne;
  }

  // ---- outputs: lane = point again, coalesced per plane ---------------------------------------
  if (inb) {
    p3d[3 * n] = bx;
    p3d[3 * n + 1] = by;
    p3d[3 * n + 2] = bz;
    err_out[n] = (best_s >= 0) ? best_err : 0.0;  // errors default to 0.0 (cameras.py:675)
    if (subset_out) subset_out[n] = best_s;
    if (neval_out) neval_out[n] = neval;
    if (picked || xy_picked) {
#pragma unroll 1
      for (int c = 0; c < C; ++c) {
        const bool in = (best_mask >> c) & 1u;
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
