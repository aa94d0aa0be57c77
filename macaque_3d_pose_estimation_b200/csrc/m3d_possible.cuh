// m3d_possible.cuh — CameraGroup.triangulate_possible with P > 1 candidates per camera
// (cameras.py:639-724): per point, itertools.product over the cameras that have a valid
// candidate (ascending), each offering its valid candidates (ascending) and then "none";
// combinations with fewer than min_cams cameras are skipped unless they use every such
// camera; accept when err < best (best starts at init_best), stop when best < threshold.
// The reference has no caller for P > 1 (triangulate_ransac is P = 1 and runs on
// k_ransac_search8/16), so this kernel favours simplicity over the table tricks of those:
// one warp = one point, lane = combination (32 consecutive indices of the mixed-radix product
// per step), digits decoded per lane, Gram blocks of the chosen candidates summed from shared
// memory, the warp-convergent DLT solve of m3d_ransac8.cuh, the exact mean reprojection error
// per lane.  Every combination is scored exactly, so one pass yields both the first
// combination under the threshold and the running strict arg-min.
#pragma once
#include "m3d_ransac8.cuh"

namespace m3d {

constexpr int POSS_WARPS = 4;
constexpr int POSS_SLOTS = 32;  // cameras * candidates <= 32

struct PossWarp {
  double gc[POSS_SLOTS][10];
  double raw[POSS_SLOTS][2];
  unsigned char radix[M3D_MAXC];   // valid candidates + 1 (1 for a camera without candidates)
  unsigned short cmask[M3D_MAXC];  // valid candidates of the camera, bit p
};

inline size_t possible_smem_bytes() { return ransac_rig_bytes() + POSS_WARPS * sizeof(PossWarp); }

template <bool FULL, bool PO>
__global__ void __launch_bounds__(POSS_WARPS * 32)
k_possible(const RigDev* __restrict__ rig_g, const double* __restrict__ xy, int64_t N, int P, int undistort,
           int min_cams, double thr, double init_best, double* __restrict__ p3d, uint8_t* __restrict__ picked,
           double* __restrict__ xy_picked, double* __restrict__ err_out, int32_t* __restrict__ index_out,
           int32_t* __restrict__ neval_out) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr unsigned FULLM = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  RigDev& srig = *reinterpret_cast<RigDev*>(smem);
  {
    const double* src = reinterpret_cast<const double*>(rig_g);
    double* dst = reinterpret_cast<double*>(&srig);
    for (int i = threadIdx.x; i < (int)(sizeof(RigDev) / 8); i += POSS_WARPS * 32) dst[i] = src[i];
  }
  __syncthreads();
  PossWarp& w = *reinterpret_cast<PossWarp*>(smem + ransac_rig_bytes() + warp * sizeof(PossWarp));
  const int C = srig.n_cams;
  const int mc = lane / P, mp = lane - mc * P;  // this lane's (camera, candidate) slot
  const double T1 = thr < init_best ? thr : init_best;
  for (int64_t n = (int64_t)blockIdx.x * POSS_WARPS + warp; n < N; n += (int64_t)gridDim.x * POSS_WARPS) {
    // ---- candidates of the point: validity on the RAW x (:658), Gram block of the undistorted one
    bool valid = false, usable = false;
    __syncwarp();
    if (lane < C * P) {
      const double2 q = __ldg(reinterpret_cast<const double2*>(xy) + ((int64_t)mc * N + n) * P + mp);
      valid = (q.x == q.x);
      double ux = q.x, uy = q.y;
      if (valid && undistort) undistort_point<FULL, PO>(srig.cam[mc], q.x, q.y, ux, uy);
      usable = valid && (ux == ux);
      Gram gg;
      gram_zero(gg);
      if (usable) gram_add_camera(gg, srig.cam[mc], ux, uy);
#pragma unroll
      for (int i = 0; i < 6; ++i) w.gc[lane][i] = gg.h[i];
      w.gc[lane][6] = gg.g[0];
      w.gc[lane][7] = gg.g[1];
      w.gc[lane][8] = gg.g[2];
      w.gc[lane][9] = gg.w;
      w.raw[lane][0] = q.x;
      w.raw[lane][1] = q.y;
    }
    const unsigned vbits = __ballot_sync(FULLM, valid), ubits = __ballot_sync(FULLM, usable);
    if (lane < C) {
      const unsigned m = (vbits >> (lane * P)) & ((1u << P) - 1u);
      w.cmask[lane] = (unsigned short)m;
      w.radix[lane] = (unsigned char)(__popc(m) + 1);
    }
    __syncwarp();
    int k = 0;            // cameras with at least one candidate (n_cams_max, :687)
    unsigned total = 1;   // combinations
    for (int c = 0; c < C; ++c) {
      k += w.radix[c] > 1;
      total *= w.radix[c];
    }
    // ---- the product, 32 combinations per step
    double best_err = init_best, bx = qnan(), by = qnan(), bz = qnan();
    unsigned long long bpick = ~0ull;  // nibble c = chosen candidate of camera c, 15 = none
    int32_t best_ix = -1, ne = 0;
    bool finished = false;
    for (unsigned base = 0; base < total && !finished; base += 32) {
      const unsigned s = base + (unsigned)lane;
      unsigned rem = s;
      unsigned long long pick = ~0ull;
      int cnt = 0, ucnt = 0;
      Gram G;
      gram_zero(G);
      for (int c = C - 1; c >= 0; --c) {  // last camera = least significant digit
        const unsigned r = w.radix[c];
        const unsigned d = rem % r;
        rem /= r;
        if (d + 1 < r) {                  // digit d = the d-th valid candidate (r - 1 = none)
          unsigned m = w.cmask[c];
          for (unsigned i = 0; i < d; ++i) m &= m - 1;
          const int p = __ffs(m) - 1, l = c * P + p;
          ++cnt;
          pick = (pick & ~(15ull << (4 * c))) | ((unsigned long long)p << (4 * c));
          if ((ubits >> l) & 1u) {
            ++ucnt;
            const double* g = w.gc[l];
#pragma unroll
            for (int i = 0; i < 6; ++i) G.h[i] += g[i];
            G.g[0] += g[6];
            G.g[1] += g[7];
            G.g[2] += g[8];
            G.w += g[9];
          }
        }
      }
      const bool adm = s < total && (cnt >= min_cams || cnt == k);   // :691
      const unsigned admb = __ballot_sync(FULLM, adm);
      ne += __popc(admb);
      const bool alive = adm && ucnt >= 2;
      double X, Y, Z;
      dlt_solve_warp(G, alive, X, Y, Z);
      // mean reprojection error over the chosen cameras, ascending (:701, :769-775)
      double el = qnan();
      if (alive && X == X) {
        double sum = 0.0;
        int m = 0;
        for (int c = 0; c < C; ++c) {
          const int p = (int)((pick >> (4 * c)) & 15ull);
          if (p != 15) {
            const int l = c * P + p;
            double u, v;
            project_point<FULL, PO>(srig.cam[c], X, Y, Z, u, v);
            const double e = residual_norm(w.raw[l][0] - u, w.raw[l][1] - v);
            if (e == e) {
              sum += e;
              ++m;
            }
          }
        }
        if (m >= 2) el = sum / (double)m;
      }
      // sequential accept rule over the step: the first combination under T1 ends the search;
      // otherwise the running best is the strict minimum, first occurrence
      const unsigned okb = __ballot_sync(FULLM, el < T1);
      int win = -1;
      if (okb) {
        win = __ffs(okb) - 1;
        ne -= __popc(admb & ~(0xffffffffu >> (31 - win)));
        finished = true;
      } else {
        double mv = (el == el) ? el : pos_inf();
        int ml = lane;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const double ov = __shfl_xor_sync(FULLM, mv, off);
          const int ol = __shfl_xor_sync(FULLM, ml, off);
          if (ov < mv || (ov == mv && ol < ml)) {
            mv = ov;
            ml = ol;
          }
        }
        if (mv < best_err) win = ml;
      }
      if (win >= 0) {
        best_err = __shfl_sync(FULLM, el, win);
        bx = __shfl_sync(FULLM, X, win);
        by = __shfl_sync(FULLM, Y, win);
        bz = __shfl_sync(FULLM, Z, win);
        bpick = __shfl_sync(FULLM, pick, win);
        best_ix = (int32_t)(base + (unsigned)win);
      }
    }
    // ---- outputs (:715-722)
    const bool sel = best_ix >= 0;
    if (lane == 0) {
      p3d[3 * n] = sel ? bx : qnan();
      p3d[3 * n + 1] = sel ? by : qnan();
      p3d[3 * n + 2] = sel ? bz : qnan();
      err_out[n] = sel ? best_err : 0.0;
      if (index_out) index_out[n] = best_ix;
      if (neval_out) neval_out[n] = ne;
    }
    if (lane < C * P && picked)
      picked[((int64_t)mc * N + n) * P + mp] = (sel && (int)((bpick >> (4 * mc)) & 15ull) == mp) ? 1 : 0;
    if (lane < C && xy_picked) {
      const int p = (int)((bpick >> (4 * lane)) & 15ull);
      double2 q = make_double2(qnan(), qnan());
      if (sel && p != 15) q = make_double2(w.raw[lane * P + p][0], w.raw[lane * P + p][1]);
      reinterpret_cast<double2*>(xy_picked)[(int64_t)lane * N + n] = q;
    }
  }
}

}  // namespace m3d
