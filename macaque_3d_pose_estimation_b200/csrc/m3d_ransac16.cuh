// m3d_ransac16.cuh — subset search of K4 for rigs of 9..16 cameras: k_ransac_search8
// (m3d_ransac8.cuh) with one more table level.  One warp = one point, lane = subset; a step
// evaluates s = 32 * hi + lane with hi = 32 * top + mid:
//   glow[5][32] x2  Gram sum of the cameras kept by bits 0..4   (index = lane)
//   gmid[32][10]    ... by bits 5..9                            (index = mid)
//   gtop[64][10]    ... by bits 10..15                          (index = top)
// (a Gram = five 16-byte pairs (h0 h1)(h2 h3)(h4 h5)(g0 g1)(g2 w)), so a subset Gram is 15
// 16-byte shared loads + 20 adds.  The rank masks of the dropped cameras
// (suspicion order) are split the same way (dlow in a register, dmid by shuffle, dtop in shared
// memory).  The local camera numbering (local index b = the bit of s that drops the camera) is
// derived here from the slot's valid mask and physical suspicion order: with up to 65,519 subsets
// per point the per-point setup is noise.  Solve, pruning and exact scoring are those of the
// 8-camera kernel (scoring: 16 lanes = cameras per candidate, two candidates at a time).
#pragma once
#include "m3d_ransac8.cuh"

namespace m3d {

// shared memory of one warp (doubles): gc[16][10] | raw[16][2] | glow[5][32]x2 | gmid[32][10] |
// gtop[64][10] | dtop[64] (as uint32, 32 doubles)
constexpr int R16_RAW = 160, R16_GL = 192, R16_GM = 512, R16_GT = 832, R16_DT = 1472,
              R16_WARP_DOUBLES = 1472 + 32;
inline size_t ransac16_smem_bytes() {
  return ransac_rig_bytes() + (size_t)(R8_ZEROS + RANSAC_WARPS * R16_WARP_DOUBLES) * sizeof(double);
}

__device__ __forceinline__ void gram_acc10(Gram& g, const double* s) {
  const double2* q = reinterpret_cast<const double2*>(s);
  const double2 t0 = q[0], t1 = q[1], t2 = q[2], t3 = q[3], t4 = q[4];
  g.h[0] += t0.x;
  g.h[1] += t0.y;
  g.h[2] += t1.x;
  g.h[3] += t1.y;
  g.h[4] += t2.x;
  g.h[5] += t2.y;
  g.g[0] += t3.x;
  g.g[1] += t3.y;
  g.g[2] += t4.x;
  g.w += t4.y;
}

// a Gram as five 16-byte pairs, `stride` pairs apart
__device__ __forceinline__ void gram_store_pairs(double2* t, int stride, const Gram& g) {
  t[0] = make_double2(g.h[0], g.h[1]);
  t[stride] = make_double2(g.h[2], g.h[3]);
  t[2 * stride] = make_double2(g.h[4], g.h[5]);
  t[3 * stride] = make_double2(g.g[0], g.g[1]);
  t[4 * stride] = make_double2(g.g[2], g.w);
}

template <bool FULL, bool PO, int MINB>
__global__ void __launch_bounds__(RANSAC_THREADS, MINB)
k_ransac_search16(const RigDev* __restrict__ rig_g, const double* __restrict__ xy, int64_t ld, int64_t n0,
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
  const int C = srig.n_cams;
  double* wrec = zeros + R8_ZEROS + warp * R16_WARP_DOUBLES;  // gc[16][10] | raw[16][2]
  double2* glow = reinterpret_cast<double2*>(wrec + R16_GL) + lane;  // [5][32] pairs
  double2* gmid = reinterpret_cast<double2*>(wrec + R16_GM);         // [32][5] pairs
  double2* gtop = reinterpret_cast<double2*>(wrec + R16_GT);         // [64][5] pairs
  uint32_t* dtop = reinterpret_cast<uint32_t*>(wrec + R16_DT);
  const double T1 = thr < init_best ? thr : init_best;
  const int lc = lane & 15, lq = lane >> 4;  // scoring: camera / candidate slot of this lane

  uint32_t todo = 0;
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

    // ---- per-point setup: local camera numbering from (vmask, physical suspicion order)
    RansacSlot* sl = slots + cur;
    const uint32_t masks = sl->masks;
    const uint32_t vmask = masks & 0xffffu, umask = masks >> 16;
    const unsigned long long ord = sl->ord;  // nibble r = physical camera of rank r
    const int k = __popc(vmask);
    const uint32_t n_sub = 1u << k;
    const int klow = k < 5 ? k : 5, kmid = (k - klow) < 5 ? (k - klow) : 5, ktop = k - klow - kmid;
    unsigned long long vlist = 0, ordl = 0, lrank = 0;  // see RansacSlot
    uint32_t uml = 0;
    {
      int b = 0;
      for (int c = C - 1; c >= 0; --c)
        if ((vmask >> c) & 1u) {
          vlist |= (unsigned long long)c << (4 * b);
          uml |= ((umask >> c) & 1u) << b;
          ++b;
        }
      int rr = 0;
      for (int r = 0; r < C; ++r) {
        const int c = (int)((ord >> (4 * r)) & 15ull);
        if ((vmask >> c) & 1u) {
          const int lb = __popc(vmask >> (c + 1));
          ordl |= (unsigned long long)lb << (4 * rr);
          lrank |= (unsigned long long)rr << (4 * lb);
          ++rr;
        }
      }
    }
    __syncwarp();  // the previous point's readers are done
    if (lane < k) {  // lane = local camera: raw pixels and Gram block (zero when unusable)
      const int c = (int)((vlist >> (4 * lane)) & 15ull);
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
      reinterpret_cast<double2*>(wrec + R16_RAW)[lane] = q;
    }
    __syncwarp();
    uint32_t dlow = 0, dmid = 0;  // rank masks dropped by bits 0..4 = lane / by bits 5..9 = lane
    {
      Gram gl, gm, gt0, gt1;
      gram_zero(gl);
      gram_zero(gm);
      gram_zero(gt0);
      gram_zero(gt1);
      uint32_t dt0 = 0, dt1 = 0;
#pragma unroll
      for (int b = 0; b < 5; ++b) {
        const bool bit = ((lane >> b) & 1) != 0;
        gram_acc10(gl, (b < klow && !bit) ? wrec + 10 * b : zeros);
        if (b < klow && bit) dlow |= 1u << ((lrank >> (4 * b)) & 15ull);
        gram_acc10(gm, (b < kmid && !bit) ? wrec + 10 * (klow + b) : zeros);
        if (b < kmid && bit) dmid |= 1u << ((lrank >> (4 * (klow + b))) & 15ull);
      }
#pragma unroll
      for (int b = 0; b < 6; ++b) {  // top table: entries lane and lane + 32
        const bool in = b < ktop;
        const bool bit0 = ((lane >> b) & 1) != 0, bit1 = (((lane + 32) >> b) & 1) != 0;
        const double* src = wrec + 10 * (klow + kmid + b);
        const uint32_t rm = in ? 1u << ((lrank >> (4 * (klow + kmid + b))) & 15ull) : 0u;
        gram_acc10(gt0, (in && !bit0) ? src : zeros);
        gram_acc10(gt1, (in && !bit1) ? src : zeros);
        if (bit0) dt0 |= rm;
        if (bit1) dt1 |= rm;
      }
      gram_store_pairs(glow, 32, gl);
      gram_store_pairs(gmid + 5 * lane, 1, gm);
      gram_store_pairs(gtop + 5 * lane, 1, gt0);
      gram_store_pairs(gtop + 5 * (lane + 32), 1, gt1);
      dtop[lane] = dt0;
      dtop[lane + 32] = dt1;
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
      const uint32_t mid = (base >> 5) & 31u, top = base >> 10;
      const uint32_t kept = ~s & (n_sub - 1u);
      const int cnt = __popc(kept);
      const bool adm = s >= 1u && s < n_sub && cnt >= min_cams;  // the full set (s = 0) was done
      const uint32_t admb = __ballot_sync(FULLM, adm);
      if (pass == 1) ne += __popc(admb);
      const uint32_t dh = __shfl_sync(FULLM, dmid, (int)mid) | dtop[top];
      bool alive = adm && (__popc(kept & uml) >= 2);
      double X, Y, Z;
      {
        Gram G;
        const double2* tm = gmid + 5 * mid;
        const double2* tt = gtop + 5 * top;
        const double2 l0 = glow[0], l1 = glow[32], l2 = glow[64], l3 = glow[96], l4 = glow[128];
        const double2 m0 = tm[0], m1 = tm[1], m2 = tm[2], m3 = tm[3], m4 = tm[4];
        const double2 t0 = tt[0], t1 = tt[1], t2 = tt[2], t3 = tt[3], t4 = tt[4];
        G.h[0] = (l0.x + m0.x) + t0.x;
        G.h[1] = (l0.y + m0.y) + t0.y;
        G.h[2] = (l1.x + m1.x) + t1.x;
        G.h[3] = (l1.y + m1.y) + t1.y;
        G.h[4] = (l2.x + m2.x) + t2.x;
        G.h[5] = (l2.y + m2.y) + t2.y;
        G.g[0] = (l3.x + m3.x) + t3.x;
        G.g[1] = (l3.y + m3.y) + t3.y;
        G.g[2] = (l4.x + m4.x) + t4.x;
        G.w = (l4.y + m4.y) + t4.y;
        dlt_solve_warp(G, alive, X, Y, Z);
        alive = alive && (X == X);
      }
      // one pruning round on the most suspicious camera of the subset
      if (alive) {
        const uint32_t ranks = ~(dlow | dh) & (n_sub - 1u);  // ranks of the cameras kept
        const int r = __ffs(ranks) - 1;
        const int b = (int)((ordl >> (4 * r)) & 15ull);
        const int c = (int)((vlist >> (4 * b)) & 15ull);
        double u, v;
        project_point<FULL, PO>(srig.cam[c], X, Y, Z, u, v);
        const double e = residual_norm(wrec[R16_RAW + 2 * b] - u, wrec[R16_RAW + 2 * b + 1] - v);
        const double limit = rb * (double)cnt * (1.0 + 1e-12);
        if (e > limit) alive = false;  // mean >= e / |S| > T: can never be accepted
      }
      // survivors in ascending s, two at a time: exact mean, 16 lanes = cameras per candidate
      uint32_t cand = __ballot_sync(FULLM, alive);
      bool finished = false;
      while (cand) {
        const uint32_t c1 = cand & (cand - 1);
        const uint32_t mine = lq == 0 ? cand : c1;
        const bool has = mine != 0;
        const int l = has ? __ffs(mine) - 1 : 0;
        cand = c1 & (c1 - 1);
        const double Xl = __shfl_sync(FULLM, X, l), Yl = __shfl_sync(FULLM, Y, l), Zl = __shfl_sync(FULLM, Z, l);
        const uint32_t keptl = ~(base + (uint32_t)l) & (n_sub - 1u);
        double e = qnan();
        if (has && ((keptl >> lc) & 1u)) {
          double u, v;
          project_point<FULL, PO>(srig.cam[(vlist >> (4 * lc)) & 15ull], Xl, Yl, Zl, u, v);
          e = residual_norm(wrec[R16_RAW + 2 * lc] - u, wrec[R16_RAW + 2 * lc + 1] - v);
        }
        // fixed-shape butterfly over the 16 camera lanes: NaN residuals count as 0 and drop out
        // of the denominator (cameras.py:771-775)
        const int m = __popc((__ballot_sync(FULLM, e == e) >> (lq * 16)) & 0xffffu);
        double sum = (e == e) ? e : 0.0;
#pragma unroll
        for (int off = 1; off < 16; off <<= 1) sum += __shfl_xor_sync(FULLM, sum, off);
        const double el = (m >= 2) ? sum * rcp((double)m) : qnan();
        if (pass == 1) {
          // first subset under T1: the reference stops here; later subsets of this step were
          // never evaluated by it
          const uint32_t okb = __ballot_sync(FULLM, has && el < rb) & 0x00010001u;
          if (okb) {
            const int src = __ffs(okb) - 1;  // lane 16 * q of the first accepted candidate
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
          for (int qi = 0; qi < 2; ++qi) {  // pass 2 (rare): sequential arg-min
            const double elq = __shfl_sync(FULLM, el, qi * 16);
            const int lqi = __shfl_sync(FULLM, has ? l : -1, qi * 16);
            if (lqi >= 0 && elq < rb) {
              rb = elq;
              if (lane == qi * 16) {
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
