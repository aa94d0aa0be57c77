// m3d_ransac8.cuh — subset search of K4 for rigs of at most 8 cameras (the lab rig and every
// BASELINE 8-view configuration; m3d_ransac16.cuh adds one table level for 9..16 cameras).
// Middle kernel of the three-kernel search described in m3d_ransac.cuh.  Its 16-lane-group
// predecessor (tools/experiments/) spent 60 % of its instructions on bookkeeping around the
// solves; here:
//   * one WARP = one point, lane = subset: a step evaluates s = 32 * hi + lane.  All control
//     flow of the search is warp-uniform.
//   * cameras are renumbered per point by the bit of the enumeration step s that drops them
//     (local index b; k_ransac_full writes the nibble lists), so the camera set of step s is
//     simply ~s — no per-step mask loops.
//   * two shared-memory tables built once per point hold the Gram sums of the cameras kept by
//     the low five bits (indexed by the lane, conflict-free) and by the high bits (at most 8
//     entries, broadcast): a subset Gram is ten 16-byte loads + 10 adds.  (Keeping the low sum in
//     registers instead costs 20 registers and a quarter of the resident warps.)
//   * the rank mask (suspicion order) of the dropped cameras is split the same way, so the
//     most suspicious member of a subset is one shuffle and one find-first-set.
//   * the Newton iteration of the DLT solve runs warp-convergent (every lane iterates until the
//     slowest one has converged; a converged lane recomputes the same values), without the
//     divergence bookkeeping of a per-lane loop.
//   * survivors of the pruning round are scored exactly four at a time (8 lanes = cameras each).
#pragma once
#include "m3d_ransac.cuh"

namespace m3d {

// shared memory of one warp: gc[8][10] | raw[8][2] | ghigh[8][10] | glow[5][32] double2   (local camera order;
// the two tables hold a Gram as five 16-byte pairs (h0 h1)(h2 h3)(h4 h5)(g0 g1)(g2 w))
constexpr int R8_RAW = 80, R8_GH = 96, R8_GL = 176, R8_WARP_DOUBLES = 176 + 320;
constexpr int R8_ZEROS = 16;
inline size_t ransac8_smem_bytes() {
  return ransac_rig_bytes() + (size_t)(R8_ZEROS + RANSAC_WARPS * R8_WARP_DOUBLES) * sizeof(double);
}

template <bool FULL, bool PO, int MINB>
__global__ void __launch_bounds__(RANSAC_THREADS, MINB)
k_ransac_search8(const RigDev* __restrict__ rig_g, const double* __restrict__ xy, int64_t ld, int64_t n0,
                 int64_t n, int min_cams, double thr, double init_best, const double* __restrict__ U,
                 RansacSlot* __restrict__ slots, unsigned long long* __restrict__ counter) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr unsigned FULLM = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  RigDev& srig = *reinterpret_cast<RigDev*>(smem);
  double* zeros = reinterpret_cast<double*>(smem + ransac_rig_bytes());
  {
    const double* src = reinterpret_cast<const double*>(rig_g);
    double* dst = reinterpret_cast<double*>(&srig);
    for (int i = threadIdx.x; i < (int)(sizeof(RigDev) / 8); i += RANSAC_THREADS) dst[i] = src[i];
    if (threadIdx.x < R8_ZEROS) zeros[threadIdx.x] = 0.0;
  }
  __syncthreads();
  double* wrec = zeros + R8_ZEROS + warp * R8_WARP_DOUBLES;  // gc[8][10] | raw[8][2]
  double* ghigh = wrec + R8_GH;                              // [8][10]     sums kept by the high bits
  double2* glow = reinterpret_cast<double2*>(wrec + R8_GL) + lane;  // [5][32] x 2 sums kept by the low bits (= lane)
  const double T1 = thr < init_best ? thr : init_best;
  const int lc = lane & 7, lq = lane >> 3;  // scoring: camera / candidate slot of this lane

  uint32_t todo = 0;   // undecided points of the current batch of 32
  int64_t batch0 = 0;
#pragma unroll 1
  for (;;) {
    // ---- next undecided point (warp-uniform)
    bool exhausted = false;
    while (!todo) {
      unsigned long long b = 0;
      if (lane == 0) b = atomicAdd(counter, 32ull);
      b = __shfl_sync(FULLM, b, 0);
      if ((int64_t)b >= n) {
        exhausted = true;
        break;
      }
      batch0 = (int64_t)b;
      const int64_t i = batch0 + lane;
      todo = __ballot_sync(FULLM, (i < n) && (slots[i].decided == 0));
    }
    if (exhausted) break;
    const int64_t cur = batch0 + (__ffs(todo) - 1);
    todo &= todo - 1;

    // ---- per-point setup
    RansacSlot* sl = slots + cur;
    const uint32_t masks = sl->masks;
    const unsigned long long o = sl->ord;
    const uint32_t vlist = sl->vlist, uml = sl->uml;
    const uint32_t ordl = (uint32_t)o, lrank = (uint32_t)(o >> 32);
    const int k = __popc(masks & 0xffffu);
    const uint32_t n_sub = 1u << k;
    const int klow = k < 5 ? k : 5, khigh = k - klow;
    __syncwarp();  // the previous point's readers are done
    if (lane < k) {  // lane = local camera: raw pixels and Gram block (zero when unusable)
      const int c = (int)((vlist >> (4 * lane)) & 15u);
      const double2 q = ld_xy(xy, (int64_t)c * ld + n0 + cur);
      Gram gg;
      gram_zero(gg);
      if ((uml >> lane) & 1u) {
        const double2 u = ld_xy(U, (int64_t)c * n + cur);
        gram_add_camera(gg, srig.cam[c], u.x, u.y);
      }
      double2* d = reinterpret_cast<double2*>(wrec + 10 * lane);
      d[0] = make_double2(gg.h[0], gg.h[1]);
      d[1] = make_double2(gg.h[2], gg.h[3]);
      d[2] = make_double2(gg.h[4], gg.h[5]);
      d[3] = make_double2(gg.g[0], gg.g[1]);
      d[4] = make_double2(gg.g[2], gg.w);
      reinterpret_cast<double2*>(wrec + R8_RAW)[lane] = q;
    }
    __syncwarp();
    uint32_t dlow = 0, dhigh = 0;  // rank masks dropped by low bits = lane / by high bits = lane & 7
    {
      Gram gl;  // Gram sum of the cameras kept by the low bits (= lane)
      gram_zero(gl);
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;  // high table: entry lc, values lq, lq + 4, lq + 8
#pragma unroll
      for (int b = 0; b < 5; ++b) {
        const bool bit = ((lane >> b) & 1) != 0;
        const bool in = b < klow;
        const double2* s = reinterpret_cast<const double2*>((in && !bit) ? wrec + 10 * b : zeros);
        const double2 t0 = s[0], t1 = s[1], t2 = s[2], t3 = s[3], t4 = s[4];
        gl.h[0] += t0.x;
        gl.h[1] += t0.y;
        gl.h[2] += t1.x;
        gl.h[3] += t1.y;
        gl.h[4] += t2.x;
        gl.h[5] += t2.y;
        gl.g[0] += t3.x;
        gl.g[1] += t3.y;
        gl.g[2] += t4.x;
        gl.w += t4.y;
        if (in && bit) dlow |= 1u << ((lrank >> (4 * b)) & 15u);
      }
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const bool bit = ((lc >> b) & 1) != 0;
        const bool in = b < khigh;
        const double* s = (in && !bit) ? wrec + 10 * (klow + b) : zeros;
        a0 += s[lq];
        a1 += s[lq + 4];
        a2 += s[lq + 8];  // lq >= 2: reads past the block, never stored
        if (in && bit) dhigh |= 1u << ((lrank >> (4 * (klow + b))) & 15u);
      }
      ghigh[lc * 10 + lq] = a0;
      ghigh[lc * 10 + lq + 4] = a1;
      if (lq < 2) ghigh[lc * 10 + lq + 8] = a2;
      glow[0] = make_double2(gl.h[0], gl.h[1]);
      glow[32] = make_double2(gl.h[2], gl.h[3]);
      glow[64] = make_double2(gl.h[4], gl.h[5]);
      glow[96] = make_double2(gl.g[0], gl.g[1]);
      glow[128] = make_double2(gl.g[2], gl.w);
      __syncwarp();
    }

    // ---- steps of 32 consecutive subsets
    uint32_t base = 0;
    int pass = 1;
    int32_t ne = 0;
    double rb = T1;
#pragma unroll 1
    for (;;) {
      const uint32_t s = base + (uint32_t)lane;
      const uint32_t hi = base >> 5;
      const uint32_t kept = ~s & (n_sub - 1u);
      const int cnt = __popc(kept);
      const bool adm = s >= 1u && s < n_sub && cnt >= min_cams;  // the full set (s = 0) was done
      const uint32_t admb = __ballot_sync(FULLM, adm);
      if (pass == 1) ne += __popc(admb);
      const uint32_t dh = __shfl_sync(FULLM, dhigh, (int)hi);
      bool alive = adm && (__popc(kept & uml) >= 2);
      double X, Y, Z;
      {
        Gram G;
        const double2* t = reinterpret_cast<const double2*>(ghigh + 10 * hi);
        const double2 l0 = glow[0], l1 = glow[32], l2 = glow[64], l3 = glow[96], l4 = glow[128];
        const double2 h0 = t[0], h1 = t[1], h2 = t[2], h3 = t[3], h4 = t[4];
        G.h[0] = l0.x + h0.x;
        G.h[1] = l0.y + h0.y;
        G.h[2] = l1.x + h1.x;
        G.h[3] = l1.y + h1.y;
        G.h[4] = l2.x + h2.x;
        G.h[5] = l2.y + h2.y;
        G.g[0] = l3.x + h3.x;
        G.g[1] = l3.y + h3.y;
        G.g[2] = l4.x + h4.x;
        G.w = l4.y + h4.y;
        dlt_solve_warp(G, alive, X, Y, Z);
        alive = alive && (X == X);
      }
      // one pruning round on the most suspicious camera of the subset
      if (alive) {
        const uint32_t ranks = ~(dlow | dh) & (n_sub - 1u);  // ranks of the cameras kept
        const int r = __ffs(ranks) - 1;
        const int b = (int)((ordl >> (4 * r)) & 15u);
        const int c = (int)((vlist >> (4 * b)) & 15u);
        double u, v;
        project_point<FULL, PO>(srig.cam[c], X, Y, Z, u, v);
        const double e = residual_norm(wrec[R8_RAW + 2 * b] - u, wrec[R8_RAW + 2 * b + 1] - v);
        const double limit = rb * (double)cnt * (1.0 + 1e-12);
        if (e > limit) alive = false;  // mean >= e / |S| > T: can never be accepted
      }
      // survivors in ascending s, four at a time: exact mean, 8 lanes = cameras per candidate
      uint32_t cand = __ballot_sync(FULLM, alive);
      bool finished = false;
      while (cand) {
        const uint32_t c1 = cand & (cand - 1), c2 = c1 & (c1 - 1), c3 = c2 & (c2 - 1);
        const uint32_t mine = lq == 0 ? cand : (lq == 1 ? c1 : (lq == 2 ? c2 : c3));
        const bool has = mine != 0;
        const int l = has ? __ffs(mine) - 1 : 0;
        cand = c3 & (c3 - 1);
        const double Xl = __shfl_sync(FULLM, X, l), Yl = __shfl_sync(FULLM, Y, l), Zl = __shfl_sync(FULLM, Z, l);
        const uint32_t keptl = ~(base + (uint32_t)l) & (n_sub - 1u);
        double e = qnan();
        if (has && ((keptl >> lc) & 1u)) {
          double u, v;
          project_point<FULL, PO>(srig.cam[(vlist >> (4 * lc)) & 15u], Xl, Yl, Zl, u, v);
          e = residual_norm(wrec[R8_RAW + 2 * lc] - u, wrec[R8_RAW + 2 * lc + 1] - v);
        }
        // fixed-shape butterfly over the 8 camera lanes: NaN residuals count as 0 and drop out
        // of the denominator (cameras.py:771-775)
        const int m = __popc((__ballot_sync(FULLM, e == e) >> (lq * 8)) & 0xffu);
        double sum = (e == e) ? e : 0.0;
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) sum += __shfl_xor_sync(FULLM, sum, off);
        const double el = (m >= 2) ? sum * rcp((double)m) : qnan();
        // sequential accept rule over the (up to) four candidates, ascending s
        if (pass == 1) {
          // first subset under T1: the reference stops here; later subsets of this step were
          // never evaluated by it
          const uint32_t okb = __ballot_sync(FULLM, has && el < rb) & 0x01010101u;
          if (okb) {
            const int src = __ffs(okb) - 1;  // lane 8 * q of the first accepted candidate
            const int lw = __shfl_sync(FULLM, l, src);
            if (lane == src) {
              sl->best_err = el;
              sl->best_s = (int32_t)(base + (uint32_t)lw);
              sl->bx = Xl;
              sl->by = Yl;
              sl->bz = Zl;
            }
            ne -= __popc(admb & ~(0xffffffffu >> (31 - lw)));
            finished = true;
          }
        } else {
#pragma unroll 1
          for (int qi = 0; qi < 4; ++qi) {  // pass 2 (rare): sequential arg-min
            const double elq = __shfl_sync(FULLM, el, qi * 8);
            const int lqi = __shfl_sync(FULLM, has ? l : -1, qi * 8);
            if (lqi >= 0 && elq < rb) {
              rb = elq;
              if (lane == qi * 8) {
                sl->best_err = el;
                sl->best_s = (int32_t)(base + (uint32_t)l);
                sl->bx = Xl;
                sl->by = Yl;
                sl->bz = Zl;
              }
            }
          }
        }
        if (finished) break;
      }
      if (finished) break;
      base += 32;
      if (base >= n_sub) {
        if (pass == 2) break;
        pass = 2;  // nothing under T1: rescan for the strict arg-min
        base = 0;
        rb = sl->best_err;  // init_best or the full-set error (>= T1)
      }
    }
    if (lane == 0) sl->neval = 1 + ne;
  }
}

}  // namespace m3d
