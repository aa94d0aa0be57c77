// m3d_viterbi.cu — 2D keypoint Viterbi filter of the step-4 stage:
// anipose/filter_pose.py viterbi_path :48-120 (with remove_dups :26-46 and the score threshold
// of filter_pose_viterbi :157), every (animal, camera, joint) series of a recording in one
// launch.  The reference runs one Python loop per series and frame (cdist + scipy logcdf +
// logsumexp on <= 3x3 matrices, ~1 ms per frame); here ONE WARP owns a series (general kernel
// k_viterbi below; k_viterbi_small further down is the block-of-32-frames form for <= 4 particles):
//   forward pass  per frame: lanes = (age, candidate) build the particle list by ballot
//                 compaction, lanes = (b, a) transition pairs evaluate
//                 log(Phi((d+2)/s) - Phi((d-2)/s)) in the closed form the reference's scipy
//                 calls reduce to, lane b takes the first-max over a (numpy NaN rules) and
//                 stores (back pointer, source) as one 16-bit code per particle;
//   backtrace     tiles of 64 frames of codes are staged in shared memory, lane 0 walks the
//                 chain inside the tile, then all lanes gather the chosen candidates.
// The filter is sequential in time; parallelism is across series (and lanes within a frame).
#include <cuda_runtime.h>

#include <cstdlib>
#include <string>

#include "../../include/m3d.h"
#include "m3d_internal.h"

namespace {

constexpr int VT_WARPS = 4;
constexpr int VT_MAXP = 32;   // particles per frame: n_back * n_possible <= 32
constexpr int VT_TILE = 64;   // frames per backtrace tile
constexpr unsigned FULLM = 0xffffffffu;

__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xfff0000000000000LL); }

// scipy.special.log_ndtr (xsf): log1p(-erfc(z / sqrt 2) / 2) for z >= -1, else the erfcx form
__device__ __forceinline__ double log_ndtr_d(double z) {
  const double t = __dmul_rn(z, 0.70710678118654757);
  if (z < -1.0) return __dadd_rn(log(erfcx(-t) / 2.0), -__dmul_rn(t, t));
  return log1p(-erfc(t) / 2.0);
}

// numpy max / argmax ordering: NaN beats everything, first occurrence wins
__device__ __forceinline__ bool np_better(double v, double best) {
  return (v > best) || (v != v && best == best);
}

// Backtrace (:112-118) of one series by its warp: tiles of VT_TILE frames of (back, source) codes
// are staged in shared memory, lane 0 walks the chain inside the tile, then all lanes gather the
// chosen candidates.  cur = particle chosen at the last frame.
__device__ __forceinline__ void viterbi_backtrace(const double* __restrict__ cs, const unsigned short* cd, int NP,
                                                  int P, int64_t F, int64_t s, int cur, unsigned short* tile,
                                                  int* choice, double* __restrict__ out,
                                                  int32_t* __restrict__ choice_out) {
  const int lane = threadIdx.x & 31;
  for (int64_t i1 = F; i1 > 0; i1 -= VT_TILE) {
    const int64_t i0 = i1 > VT_TILE ? i1 - VT_TILE : 0;
    const int nt = (int)(i1 - i0);
    __syncwarp();
    for (int q = lane; q < nt * NP; q += 32) tile[q] = cd[(size_t)i0 * NP + q];
    __syncwarp();
    if (lane == 0) {
      for (int t = nt - 1; t >= 0; --t) {
        const unsigned short code = tile[t * NP + cur];
        choice[t] = code >> 8;  // source of the particle chosen at frame i0 + t
        cur = code & 0xff;      // particle of frame i0 + t - 1
      }
    }
    cur = __shfl_sync(FULLM, cur, 0);
    __syncwarp();
    for (int t = lane; t < nt; t += 32) {
      const int src = choice[t];
      const int64_t i = i0 + t;
      double x = -1.0, y = -1.0, sc = 0.001;
      if (src != 255) {
        const int a = src / P, p = src - a * P;
        const double* c = cs + ((size_t)(i - a) * P + p) * 3;
        x = c[0];
        y = c[1];
        sc = __dmul_rn(c[2], __longlong_as_double((long long)(1023 - a) << 52));
      }
      double* o = out + ((size_t)s * F + i) * 3;
      o[0] = x;
      o[1] = y;
      o[2] = sc;
      if (choice_out) choice_out[(size_t)s * F + i] = src == 255 ? -1 : src;
    }
  }
}

// transition log-probability between two particles (:89-99)
__device__ __forceinline__ double viterbi_transition(double ax, double ay, double bx, double by, double scale,
                                                     double log_missing) {
  const double dx = ax - bx, dy = ay - by;
  const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
  const double hi = log_ndtr_d(__ddiv_rn(__dadd_rn(d, 2.0), scale));
  const double lo = log_ndtr_d(__ddiv_rn(__dadd_rn(d, -2.0), scale));
  // logsumexp([hi, lo], b = [1, -1]) of scipy >= 1.15
  double pt = (hi == lo) ? neg_inf() : __dadd_rn(log1p(-exp(__dadd_rn(lo, -hi))), hi);
  if (pt < -100.0) pt = -100.0;
  if (bx == -1.0 || ax == -1.0) pt = log_missing;
  return pt;
}

struct VtWarp {
  double pax[VT_MAXP], pay[VT_MAXP], Tp[VT_MAXP];
  double pbx[VT_MAXP], pby[VT_MAXP], sb[VT_MAXP], Tb[VT_MAXP];
  unsigned char srcb[VT_MAXP];
  union {
    double poss[VT_MAXP * 8];                    // transition chunk: 8 rows b x 32 a
    unsigned short codes[VT_TILE * VT_MAXP];     // backtrace tile
  };
  int choice[VT_TILE];
};

__global__ void __launch_bounds__(VT_WARPS * 32)
k_viterbi(const double* __restrict__ cand, int64_t S, int64_t F, int P, int n_back, double scale,
          double score_thr, double dup_thr2, double log_missing, unsigned short* __restrict__ codes,
          double* __restrict__ out, int32_t* __restrict__ choice_out) {
  __shared__ VtWarp sm[VT_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t s = (int64_t)blockIdx.x * VT_WARPS + warp;
  if (s >= S) return;
  VtWarp& w = sm[warp];
  const int NP = n_back * P;  // <= VT_MAXP
  const double* cs = cand + (size_t)s * F * P * 3;
  unsigned short* cd = codes + (size_t)s * F * NP;
  const int age = lane / P, pidx = lane - age * P;  // particle-building role of this lane
  unsigned vring[8];  // valid-candidate masks of the last n_back frames (n_back <= 8)
#pragma unroll
  for (int j = 0; j < 8; ++j) vring[j] = 0;
  int va = 0;
  for (int64_t i = 0; i < F; ++i) {
    // ---- valid candidates of frame i: score threshold, then remove_dups within the frame
    {
      double x = 0.0, y = 0.0;
      bool nanx = true;
      if (lane < P) {
        const double* c = cs + ((size_t)i * P + lane) * 3;
        const double sc = c[2];
        x = c[0];
        y = c[1];
        if (sc < score_thr) x = y = __longlong_as_double(0x7ff8000000000000LL);
        nanx = (x != x);
        // non-finite coordinates move to 1e9 for the duplicate search (filter_pose.py:31)
        if (!(fabs(x) <= 1.7976931348623157e308)) x = 1e9;
        if (!(fabs(y) <= 1.7976931348623157e308)) y = 1e9;
      }
      bool dup = false;
      for (int p2 = 0; p2 < P; ++p2) {
        const double x2 = __shfl_sync(FULLM, x, p2), y2 = __shfl_sync(FULLM, y, p2);
        const double dx = x - x2, dy = y - y2;
        if (p2 < lane && __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) <= dup_thr2) dup = true;
      }
      const unsigned vm = __ballot_sync(FULLM, lane < P && !nanx && !dup);
#pragma unroll
      for (int j = 7; j > 0; --j) vring[j] = vring[j - 1];
      vring[0] = vm;
    }
    // ---- particles of frame i: candidates of frames i, i-1, .. in that order (:58-70)
    bool have = false;
    double px = 0.0, py = 0.0, ps = 0.0;
    if (lane < NP && i - age >= 0) {
      unsigned m = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j == age) m = vring[j];
      if ((m >> pidx) & 1u) {
        const double* c = cs + ((size_t)(i - age) * P + pidx) * 3;
        px = c[0];
        py = c[1];
        ps = __dmul_rn(c[2], __longlong_as_double((long long)(1023 - age) << 52));  // score * 2^-age
        have = true;
      }
    }
    const unsigned hb = __ballot_sync(FULLM, have);
    int vb = __popc(hb);
    __syncwarp();
    if (have) {
      const int idx = __popc(hb & ((1u << lane) - 1u));
      w.pbx[idx] = px;
      w.pby[idx] = py;
      w.sb[idx] = ps;
      w.srcb[idx] = (unsigned char)lane;
    }
    if (vb == 0) {  // missing point (:71-73)
      if (lane == 0) {
        w.pbx[0] = -1.0;
        w.pby[0] = -1.0;
        w.sb[0] = 0.001;
        w.srcb[0] = 255;
      }
      vb = 1;
    }
    __syncwarp();
    // ---- Viterbi step (:85-107)
    int backp = 0;
    double Tnew = neg_inf();
    if (i == 0) {
      if (lane < vb) Tnew = log(w.sb[lane]);
    } else {
      // rows b in chunks of 8: lanes = (b, a) pairs
      double best = 0.0;
      for (int b0 = 0; b0 < vb; b0 += 8) {
        const int nb = (vb - b0) < 8 ? (vb - b0) : 8;
        for (int q = lane; q < nb * va; q += 32) {
          const int bl = q / va, a = q - bl * va, b = b0 + bl;
          const double pt = viterbi_transition(w.pax[a], w.pay[a], w.pbx[b], w.pby[b], scale, log_missing);
          w.poss[bl * 32 + a] = __dadd_rn(w.Tp[a], pt);
        }
        __syncwarp();
        if (lane >= b0 && lane < b0 + nb) {
          const double* row = w.poss + (lane - b0) * 32;
          best = row[0];
          backp = 0;
          for (int a = 1; a < va; ++a)
            if (np_better(row[a], best)) {
              best = row[a];
              backp = a;
            }
        }
        __syncwarp();
      }
      if (lane < vb) Tnew = __dadd_rn(best, log(w.sb[lane]));
    }
    if (lane < NP)
      cd[(size_t)i * NP + lane] = lane < vb ? (unsigned short)(backp | ((int)w.srcb[lane] << 8)) : (unsigned short)0xff00;
    // current frame becomes the previous one
    if (lane < vb) {
      w.pax[lane] = w.pbx[lane];
      w.pay[lane] = w.pby[lane];
      w.Tp[lane] = Tnew;
    }
    va = vb;
    __syncwarp();
  }
  if (F == 0) return;
  // ---- last frame: first arg-max of T (:109-110)
  int cur = 0;
  {
    double best = w.Tp[0];
    for (int a = 1; a < va; ++a)
      if (np_better(w.Tp[a], best)) {
        best = w.Tp[a];
        cur = a;
      }
  }
  __threadfence_block();
  viterbi_backtrace(cs, cd, NP, P, F, s, cur, w.codes, w.choice, out, choice_out);
}

// ---------------------------------------------------------------------------------------------
// Fast path for n_back * n_possible <= 4 particles per frame (step 4: one candidate, n_back = 3).
// Nothing but the max-plus recursion itself is sequential in time, so a warp takes its series in
// blocks of 32 frames:
//   phase A (lane = frame)  valid candidates, particle list of the frame, and all (<= 16)
//                           transition log-probabilities + log scores, staged in shared memory;
//   phase B (sequential)    32 recursion steps of a few adds / compares each (lanes = particles,
//                           previous scores by shuffle), codes staged and written coalesced.
// Same arithmetic, same codes and the same backtrace as k_viterbi.
// ---------------------------------------------------------------------------------------------
struct VtSmall {
  double tr[32][16];             // tr[f][b * 4 + a]
  double ls[32][4];              // log score of particle b of frame f
  unsigned char src[33][4];      // particle sources; slot 0 = last frame of the previous block
  unsigned char cnt[33];
  unsigned short blk[32 * 4];    // codes of the block
  unsigned short codes[VT_TILE * 4];
  int choice[VT_TILE];
  double dm[36][4];              // P == 1: dm[4 + f][k - 1] = transition between the candidates of
                                 // frames f and f - k; rows 0..3 = last four frames of the previous block
};

__device__ __forceinline__ void vt_particle(const double* __restrict__ cs, int64_t frame, int P, int src, double& x,
                                            double& y, double& sc) {
  if (src == 255) {
    x = -1.0;
    y = -1.0;
    sc = 0.001;
  } else {
    const int a = src / P, p = src - a * P;
    const double* c = cs + ((size_t)(frame - a) * P + p) * 3;
    x = c[0];
    y = c[1];
    sc = __dmul_rn(c[2], __longlong_as_double((long long)(1023 - a) << 52));
  }
}

__global__ void __launch_bounds__(VT_WARPS * 32)
k_viterbi_small(const double* __restrict__ cand, int64_t S, int64_t F, int P, int n_back, double scale,
                double score_thr, double dup_thr2, double log_missing, unsigned short* __restrict__ codes,
                double* __restrict__ out, int32_t* __restrict__ choice_out) {
  __shared__ VtSmall sm[VT_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t s = (int64_t)blockIdx.x * VT_WARPS + warp;
  if (s >= S) return;
  VtSmall& w = sm[warp];
  const int NP = n_back * P;  // <= 4
  const double* cs = cand + (size_t)s * F * P * 3;
  unsigned short* cd = codes + (size_t)s * F * NP;
  const double t_same = viterbi_transition(0.0, 0.0, 0.0, 0.0, scale, log_missing);  // d = 0
  unsigned carry[3] = {0u, 0u, 0u};  // valid masks of frames i0 - 1, i0 - 2, i0 - 3
  double T = neg_inf();              // lane b: score of particle b of the previous frame
  int va = 0;
  for (int64_t i0 = 0; i0 < F; i0 += 32) {
    const int nf = (F - i0) < 32 ? (int)(F - i0) : 32;
    const int64_t i = i0 + lane;
    // ---- phase A: lane = frame
    unsigned vm = 0;
    if (lane < nf) {
      double qx[4], qy[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        if (p < P) {
          const double* c = cs + ((size_t)i * P + p) * 3;
          double x = c[0], y = c[1];
          if (c[2] < score_thr) x = y = __longlong_as_double(0x7ff8000000000000LL);
          const bool nanx = (x != x);
          if (!(fabs(x) <= 1.7976931348623157e308)) x = 1e9;
          if (!(fabs(y) <= 1.7976931348623157e308)) y = 1e9;
          qx[p] = x;
          qy[p] = y;
          bool dup = false;
#pragma unroll
          for (int p2 = 0; p2 < 4; ++p2)
            if (p2 < p) {
              const double dx = x - qx[p2], dy = y - qy[p2];
              if (__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) <= dup_thr2) dup = true;
            }
          if (!nanx && !dup) vm |= 1u << p;
        }
      }
    }
    unsigned pm[4];  // valid masks of frames i, i - 1, i - 2, i - 3
    pm[0] = vm;
#pragma unroll
    for (int j = 1; j < 4; ++j) {
      unsigned t = __shfl_up_sync(FULLM, vm, j);
      if (lane < j) t = carry[j - lane - 1];
      pm[j] = t;
    }
    int vb = 0;
    unsigned char mysrc[4] = {255, 255, 255, 255};
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int p = 0; p < 4; ++p)
        if (j < n_back && p < P && ((pm[j] >> p) & 1u)) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (q == vb) mysrc[q] = (unsigned char)(j * P + p);
          ++vb;
        }
    if (vb == 0) vb = 1;  // the missing-point particle (source 255)
    __syncwarp();
    if (lane < nf) {
#pragma unroll
      for (int q = 0; q < 4; ++q) w.src[lane + 1][q] = mysrc[q];
      w.cnt[lane + 1] = (unsigned char)vb;
    }
    __syncwarp();
    if (lane < nf) {
      double bx[4], by[4];
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (b < vb) {
          double sc;
          vt_particle(cs, i, P, mysrc[b], bx[b], by[b], sc);
          w.ls[lane][b] = log(sc);
        }
      if (P == 1) {
        // One candidate per frame: a particle IS the candidate of frame i - age, so the transition of
        // a pair depends only on the two candidate frames.  Of the (up to) 9 pairs of a frame only
        // the ones that involve candidate i are new — compute those (<= n_back values) here, the
        // rest are the neighbouring frames' values (or the d = 0 constant), picked up below.
        if (vm & 1u) {
          const double* c = cs + (size_t)i * 3;
          const double cx = c[0], cy = c[1];
#pragma unroll
          for (int k = 1; k <= 3; ++k)
            if (k <= n_back && i - k >= 0 && (pm[k] & 1u)) {
              const double* o = cs + (size_t)(i - k) * 3;
              w.dm[4 + lane][k - 1] = viterbi_transition(o[0], o[1], cx, cy, scale, log_missing);
            }
          if (n_back >= 4 && i - 4 >= 0) {  // the fourth frame back is outside the pm[] window: test it here
            const double* o = cs + (size_t)(i - 4) * 3;
            double ox = o[0];
            if (o[2] < score_thr) ox = __longlong_as_double(0x7ff8000000000000LL);
            if (ox == ox) w.dm[4 + lane][3] = viterbi_transition(o[0], o[1], cx, cy, scale, log_missing);
          }
        }
      }
    }
    __syncwarp();
    if (lane < nf) {
      if (i > 0) {
        const int pa = w.cnt[lane];
        if (P == 1) {
#pragma unroll 1
          for (int a = 0; a < pa; ++a) {
            const int sa = w.src[lane][a];  // age of particle a at frame i - 1 (255 = missing)
#pragma unroll
            for (int b = 0; b < 4; ++b)
              if (b < vb) {
                const int sb = mysrc[b];
                double pt = log_missing;
                if (sa != 255 && sb != 255) {
                  const int fb = lane - sb, fa = lane - 1 - sa;  // candidate frames relative to i0
                  pt = (fb == fa) ? t_same : (fb > fa ? w.dm[4 + fb][fb - fa - 1] : w.dm[4 + fa][fa - fb - 1]);
                }
                w.tr[lane][b * 4 + a] = pt;
              }
          }
        } else {
          double bx[4], by[4];
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (b < vb) {
              double sc;
              vt_particle(cs, i, P, mysrc[b], bx[b], by[b], sc);
            }
#pragma unroll 1
          for (int a = 0; a < pa; ++a) {
            double ax, ay, asc;
            vt_particle(cs, i - 1, P, w.src[lane][a], ax, ay, asc);
#pragma unroll
            for (int b = 0; b < 4; ++b)
              if (b < vb) w.tr[lane][b * 4 + a] = viterbi_transition(ax, ay, bx[b], by[b], scale, log_missing);
          }
        }
      }
    }
    __syncwarp();
    // ---- phase B: the recursion over the frames of the block (lanes = particles)
    for (int f = 0; f < nf; ++f) {
      const int nb = w.cnt[f + 1];
      double Tn = neg_inf();
      int backp = 0;
      if (i0 + f == 0) {
        if (lane < nb) Tn = w.ls[0][lane];
      } else {
        double best = 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const double ta = __shfl_sync(FULLM, T, a);
          if (a < va && lane < nb) {
            const double v = __dadd_rn(ta, w.tr[f][lane * 4 + a]);
            if (a == 0 || np_better(v, best)) {
              best = v;
              backp = a;
            }
          }
        }
        if (lane < nb) Tn = __dadd_rn(best, w.ls[f][lane]);
      }
      if (lane < NP)
        w.blk[f * NP + lane] = lane < nb ? (unsigned short)(backp | ((int)w.src[f + 1][lane] << 8)) : (unsigned short)0xff00;
      T = Tn;
      va = nb;
    }
    __syncwarp();
    for (int q = lane; q < nf * NP; q += 32) cd[(size_t)i0 * NP + q] = w.blk[q];
    // carry to the next block
    if (P == 1 && lane < 16) {
      const double v = w.dm[nf + (lane >> 2)][lane & 3];   // rows nf .. nf + 3 = the last four frames
      __syncwarp(0xffffu);
      w.dm[lane >> 2][lane & 3] = v;
    }
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) w.src[0][q] = w.src[nf][q];
      w.cnt[0] = w.cnt[nf];
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) carry[q] = __shfl_sync(FULLM, vm, 31 - q);
    __syncwarp();
  }
  if (F == 0) return;
  // ---- last frame: first arg-max of T (:109-110)
  int cur = 0;
  {
    double best = __shfl_sync(FULLM, T, 0);
#pragma unroll
    for (int a = 1; a < 4; ++a) {
      const double ta = __shfl_sync(FULLM, T, a);
      if (a < va && np_better(ta, best)) {
        best = ta;
        cur = a;
      }
    }
  }
  __threadfence_block();
  viterbi_backtrace(cs, cd, NP, P, F, s, cur, w.codes, w.choice, out, choice_out);
}

}  // namespace

extern "C" int m3d_viterbi_filter(const double* cand_dev, int64_t S, int64_t F, int32_t P, int32_t n_back,
                                  double thres_dist, double score_threshold, double dup_thres,
                                  double* out_dev, int32_t* choice_dev, int32_t device, void* stream) {
  if (S < 0 || F < 0 || P < 1 || n_back < 1)
    return m3d_fail(M3D_ERR_INVALID, "m3d_viterbi_filter: bad sizes");
  if (n_back > 8 || (int64_t)n_back * P > VT_MAXP)
    return m3d_fail(M3D_ERR_INVALID, "m3d_viterbi_filter: n_back * n_possible must be <= 32 and n_back <= 8");
  if (S == 0 || F == 0) return M3D_OK;
  if (!cand_dev || !out_dev) return m3d_fail(M3D_ERR_INVALID, "m3d_viterbi_filter: NULL buffer");
  M3dDeviceGuard guard(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned short* codes = nullptr;
  const size_t bytes = sizeof(unsigned short) * (size_t)S * F * n_back * P;
  int cur_dev = device;
  cudaGetDevice(&cur_dev);
  cudaMemPool_t pool = m3d_scratch_pool(cur_dev);
  cudaError_t e = pool ? cudaMallocFromPoolAsync(&codes, bytes, pool, st) : cudaMallocAsync(&codes, bytes, st);
  if (e != cudaSuccess) return m3d_fail(M3D_ERR_CUDA, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
  const unsigned blocks = (unsigned)((S + VT_WARPS - 1) / VT_WARPS);
  static const bool force_general = getenv("M3D_VITERBI_GENERAL") != nullptr;  // developer A/B switch
  const double log_missing = -6.907755278982137;  // np.log(0.001)
  if (n_back * P <= 4 && !force_general)
    k_viterbi_small<<<blocks, VT_WARPS * 32, 0, st>>>(cand_dev, S, F, P, n_back, thres_dist, score_threshold,
                                                      dup_thres * dup_thres, log_missing, codes, out_dev, choice_dev);
  else
    k_viterbi<<<blocks, VT_WARPS * 32, 0, st>>>(cand_dev, S, F, P, n_back, thres_dist, score_threshold,
                                                dup_thres * dup_thres, log_missing, codes, out_dev, choice_dev);
  const int rc = m3d_check_launch("k_viterbi");
  cudaFreeAsync(codes, st);
  if (pool && bytes > (size_t(1) << 30)) cudaMemPoolTrimTo(pool, size_t(1) << 30);  // keep at most 1 GiB cached
  return rc;
}
