// m3d_crossview.cu — cross-view association kernels (step2 of the reference pipeline):
//   k_ray_affinity : geometry_affinity2 (step2_crossviewmatching.py:373-432) with
//                    deproject (:327-355) and calc_dist_btw_lines (:359-369) fused, one CTA
//                    per frame, rays staged in shared memory;
//   k_match_svt    : matchSVT (step2_crossviewmatching.py:130-216), one CTA per frame, the
//                    ADMM iterate and a one-sided Jacobi SVD of it held in shared memory.
#include <cuda_runtime.h>

#include <cstdlib>
#include <string>

#include "../../include/m3d.h"
#include "m3d_internal.h"
#include "m3d_math.cuh"

using namespace m3d;

// ---------------------------------------------------------------------------------------
// K5: ray affinity
// ---------------------------------------------------------------------------------------
// dynamic shared memory layout per CTA:
//   dir   [M][J][3]  unit ray directions (world frame)
//   score [M][J]
//   cen   [C][3]     camera centres R^-1 (0 - t)
//   cam   [M]        camera index of each detection (-1 = padding)
//   red   [32]       block-reduction scratch
template <int THREADS>
__device__ double block_sum(double v, double* red) {
  __syncthreads();
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = (threadIdx.x < THREADS / 32) ? red[threadIdx.x] : 0.0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  t = red[0];
  __syncthreads();
  return t;
}

constexpr int AFF_THREADS = 256;

__global__ void __launch_bounds__(AFF_THREADS)
k_ray_affinity(const __grid_constant__ RigDev rig, const double* __restrict__ kp,
               const int32_t* __restrict__ dim, int M, int J, double thr_kp,
               double* __restrict__ aff, double* __restrict__ dist_out) {
  extern __shared__ double smem[];
  const int C = rig.n_cams;
  double* dir = smem;                      // M*J*3
  double* score = dir + (size_t)M * J * 3;  // M*J
  double* cen = score + (size_t)M * J;      // C*3
  double* red = cen + 3 * M3D_MAXC;         // 32
  int* camof = reinterpret_cast<int*>(red + 32);  // M
  const int f = blockIdx.x;
  const double* kpf = kp + (size_t)f * M * J * 3;
  const int32_t* dg = dim + (size_t)f * (C + 1);
  double* D = (dist_out ? dist_out : aff) + (size_t)f * M * M;  // distances staged in the output
  double* A = aff + (size_t)f * M * M;
  const int tid = threadIdx.x;

  // camera of each detection: searchsorted(dimGroup, i, side='right') - 1  (step2:395-397)
  for (int i = tid; i < M; i += AFF_THREADS) {
    int c = -1;
    for (int j = 0; j <= C; ++j)
      if (dg[j] <= i) c = j;
    camof[i] = (c >= 0 && c < C && i < dg[C]) ? c : -1;
  }
  // camera centres: R^-1 (0 - t) = -R^T t   (step2:343-354, depth 0)
  for (int c = tid; c < C; c += AFF_THREADS) {
    const CamDev& cam = rig.cam[c];
    cen[3 * c + 0] = -(cam.R[0] * cam.t[0] + cam.R[3] * cam.t[1] + cam.R[6] * cam.t[2]);
    cen[3 * c + 1] = -(cam.R[1] * cam.t[0] + cam.R[4] * cam.t[1] + cam.R[7] * cam.t[2]);
    cen[3 * c + 2] = -(cam.R[2] * cam.t[0] + cam.R[5] * cam.t[1] + cam.R[8] * cam.t[2]);
  }
  __syncthreads();
  // unit direction of every keypoint ray: R^T [x, y, 1] normalised (far - near at depth 1000)
  for (int e = tid; e < M * J; e += AFF_THREADS) {
    const int i = e / J;
    const int c = camof[i];
    const double x = kpf[3 * e], y = kpf[3 * e + 1];
    score[e] = kpf[3 * e + 2];
    double dx = 0, dy = 0, dz = 0;
    if (c >= 0) {
      const CamDev& cam = rig.cam[c];
      const double vx = cam.R[0] * x + cam.R[3] * y + cam.R[6];
      const double vy = cam.R[1] * x + cam.R[4] * y + cam.R[7];
      const double vz = cam.R[2] * x + cam.R[5] * y + cam.R[8];
      const double inv = 1.0 / sqrt(vx * vx + vy * vy + vz * vz);
      dx = vx * inv;
      dy = vy * inv;
      dz = vz * inv;
    }
    dir[3 * e] = dx;
    dir[3 * e + 1] = dy;
    dir[3 * e + 2] = dz;
  }
  __syncthreads();
  // mean line-line distance of every cross-camera pair (step2:411-424)
  for (int e = tid; e < M * M; e += AFF_THREADS) {
    const int i = e / M, j = e % M;
    if (i > j) continue;
    double d = 300.0;  // Dth2 * 2
    if (camof[i] < 0 || camof[j] < 0) {
      d = 300.0;  // padding rows / columns: outside the frame's (n x n) matrix
    } else if (i == j) {
      d = 0.0;
    } else {
      const int ci = camof[i], cj = camof[j];
      if (ci >= 0 && cj >= 0 && ci != cj) {
        const double px = cen[3 * cj] - cen[3 * ci], py = cen[3 * cj + 1] - cen[3 * ci + 1],
                     pz = cen[3 * cj + 2] - cen[3 * ci + 2];
        double sum = 0.0;
        int cnt = 0;
        for (int k = 0; k < J; ++k) {
          if (score[i * J + k] > thr_kp && score[j * J + k] > thr_kp) {
            const double* a = dir + 3 * (i * J + k);
            const double* b = dir + 3 * (j * J + k);
            const double cx = a[1] * b[2] - a[2] * b[1];
            const double cy = a[2] * b[0] - a[0] * b[2];
            const double cz = a[0] * b[1] - a[1] * b[0];
            sum += fabs(px * cx + py * cy + pz * cz) / sqrt(cx * cx + cy * cy + cz * cz);
            ++cnt;
          }
        }
        if (cnt >= 3) d = sum / (double)cnt;
      }
    }
    D[i * M + j] = d;
    D[j * M + i] = d;
  }
  __syncthreads();
  // statistics over entries < 300 (diagonal included; step2:426-428), two-pass like np.std
  double s = 0.0, cntv = 0.0;
  for (int e = tid; e < M * M; e += AFF_THREADS) {
    const double d = D[e];
    if (d < 300.0) {  // padding entries are 300 and never counted
      s += d;
      cntv += 1.0;
    }
  }
  const double tot = block_sum<AFF_THREADS>(s, red);
  const double nv = block_sum<AFF_THREADS>(cntv, red);
  const double mean = tot / nv;
  double v = 0.0;
  for (int e = tid; e < M * M; e += AFF_THREADS) {
    const double d = D[e];
    if (d < 300.0) v += (d - mean) * (d - mean);
  }
  const double sd = sqrt(block_sum<AFF_THREADS>(v, red) / nv);
  for (int e = tid; e < M * M; e += AFF_THREADS) {
    const double d = D[e];
    const double z = -(d - mean) / sd;
    double a = 1.0 / (1.0 + exp(-5.0 * z));
    if (d > 150.0) a = 0.0;
    A[e] = a;
  }
}

// ---------------------------------------------------------------------------------------
// K6: matchSVT — ADMM with singular-value thresholding, one CTA per frame (persistent CTAs
// loop over frames).  The SVD of the M x M iterate is a one-sided (Hestenes) Jacobi on its
// columns, held in shared memory: A V = U S with orthogonal columns, so the reference's
// Q = U max(S - lambda/mu, 0) V^T is sum_j max(s_j - tau, 0)/s_j * (A v_j) v_j^T.
// ---------------------------------------------------------------------------------------
constexpr int SVT_THREADS = 256;
constexpr int SVT_MAX_M = 112;

__global__ void __launch_bounds__(SVT_THREADS)
k_match_svt(const double* __restrict__ Wall, const int32_t* __restrict__ dim, int F, int M, int C,
            double alpha, double lambda, double mu0, double tol, int max_iter,
            uint8_t* __restrict__ match, int32_t* __restrict__ iters, double* __restrict__ wsall, int ld) {
  extern __shared__ double sm[];
  double* B = sm;                          // [Me][ld] columns of the iterate (col-major)
  double* V = B + (size_t)ld * (M + 1);    // [Me][ld] accumulated rotations
  double* sig = V + (size_t)ld * (M + 1);  // [Me] singular values -> shrink weights
  double* red = sig + (M + 1);             // [>=32] reduction scratch
  int* flag = reinterpret_cast<int*>(red + 40);
  double* nrm = red + 48;                  // [Me] running squared column norms of B
  const int tid = threadIdx.x;
  double* X = wsall + (size_t)blockIdx.x * 5 * M * M;
  double* X0 = X + (size_t)M * M;
  double* Y = X0 + (size_t)M * M;
  double* Wm = Y + (size_t)M * M;
  double* Q = Wm + (size_t)M * M;

  for (int f = blockIdx.x; f < F; f += gridDim.x) {
    const int32_t* dg = dim + (size_t)f * (C + 1);
    const int n = dg[C] < M ? dg[C] : M;  // real detections of this frame
    const double* Wf = Wall + (size_t)f * M * M;
    uint8_t* out = match + (size_t)f * M * M;
    for (int e = tid; e < M * M; e += SVT_THREADS) out[e] = 0;
    if (n <= 0) {
      if (iters && tid == 0) iters[f] = 0;
      continue;
    }
    const int ne = n + (n & 1);  // even size for the tournament schedule (dummy column = n)
    const int npairs = ne / 2;
    // lanes per column pair: as many as fit (the round is latency-bound: measured 35.6e3 frames/s
    // with 8 lanes per pair at M = 48, 24.4e3 with 4)
    int g = 32;
    while (g > 1 && g * npairs > SVT_THREADS) g >>= 1;
    const int pair_of = tid / g, sub = tid % g;
    const bool jactive = pair_of < npairs;
    // S = W with zero diagonal, symmetrised; X = S; Y = 0; Wm = alpha - S   (step2:150-157)
    for (int e = tid; e < n * n; e += SVT_THREADS) {
      const int i = e / n, j = e % n;
      const double a = (i == j) ? 0.0 : Wf[i * M + j];
      const double b = (i == j) ? 0.0 : Wf[j * M + i];
      const double sv = (a + b) / 2.0;
      X[e] = sv;
      Y[e] = 0.0;
      Wm[e] = alpha - sv;
    }
    __syncthreads();
    double mu = mu0;
    int it = 0;
#ifdef M3D_DEBUG_SWEEPS
    int dbg_sweeps = 0;
#endif
    for (it = 0; it < max_iter; ++it) {
      // A = Y/mu + X.  First iteration: B = A (column j of the iterate in B[j]), V = I.  Later
      // iterations warm-start the Jacobi process from the previous iteration's rotations:
      // B = A V with V kept — the columns of A V are already nearly orthogonal (the ADMM iterate
      // moves little), so 2-3 sweeps replace ~9 from the identity.
      if (it == 0) {
        for (int e = tid; e < ne * ne; e += SVT_THREADS) {
          const int j = e / ne, i = e % ne;
          double a = 0.0;
          if (i < n && j < n) a = Y[i * n + j] / mu + X[i * n + j];
          B[j * ld + i] = a;
          V[j * ld + i] = (i == j) ? 1.0 : 0.0;
        }
      } else {
        for (int e = tid; e < n * n; e += SVT_THREADS) Q[e] = Y[e] / mu + X[e];  // Q is free here
        __syncthreads();
        for (int e = tid; e < ne * ne; e += SVT_THREADS) {
          const int j = e / ne, i = e % ne;
          double a = 0.0;
          if (i < n) {
            const double* ar = Q + (size_t)i * n;
            const double* vj = V + (size_t)j * ld;
            for (int k = 0; k < n; ++k) a += ar[k] * vj[k];
          }
          B[j * ld + i] = a;
        }
      }
      for (int e = tid; e < n * n; e += SVT_THREADS) X0[e] = X[e];
      __syncthreads();
      // one-sided Jacobi sweeps
      for (int sweep = 0; sweep < 40; ++sweep) {
        if (tid == 0) *flag = 0;
        __syncthreads();
        // exact squared column norms at the start of every sweep; inside the sweep they follow the
        // rotations (aa' = aa - t ab, bb' = bb + t ab), so a pair costs ONE dot product
        for (int j = tid; j < ne; j += SVT_THREADS) {
          const double* bj = B + (size_t)j * ld;
          double ss = 0.0;
          for (int i = 0; i < ne; ++i) ss += bj[i] * bj[i];
          nrm[j] = ss;
        }
        __syncthreads();
        for (int r = 0; r < ne - 1; ++r) {
          const unsigned gmask = __ballot_sync(0xffffffffu, jactive);  // lanes that own a column pair
          if (jactive) {
            int p, q;
            if (pair_of == 0) {
              p = ne - 1;
              q = r;
            } else {
              p = r + pair_of;  // (r + pair_of) mod (ne - 1), pair_of < ne / 2
              if (p >= ne - 1) p -= ne - 1;
              q = r - pair_of;
              if (q < 0) q += ne - 1;
            }
            // 16-byte accesses: lane `sub` of the pair holds elements 2 sub, 2 sub + 1 (+ 2 g k)
            double2* bp = reinterpret_cast<double2*>(B + (size_t)p * ld);
            double2* bq = reinterpret_cast<double2*>(B + (size_t)q * ld);
            const int nh = ne >> 1;
            double ab = 0.0;
            for (int i = sub; i < nh; i += g) {
              const double2 x = bp[i], y = bq[i];
              ab += x.x * y.x;
              ab += x.y * y.y;
            }
            // read the running norms BEFORE the group's shuffles: lane sub == 0 overwrites them below,
            // and only the shuffle keeps it from running ahead of a lane that has not read them yet
            double aa = nrm[p], bb = nrm[q];
            for (int off = g >> 1; off > 0; off >>= 1) ab += __shfl_xor_sync(gmask, ab, off, 32);
            aa = aa > 0.0 ? aa : 0.0;
            bb = bb > 0.0 ? bb : 0.0;
            // rotate when |ab| > 1e-15 sqrt(aa bb); MUFU-seeded reciprocal / sqrt (m3d_math.cuh)
            if (ab * ab > 1e-30 * (aa * bb) && fabs(ab) > 1e-290) {
              const double zeta = (bb - aa) * rcp(2.0 * ab);
              const double t = (zeta >= 0.0 ? 1.0 : -1.0) * rcp(fabs(zeta) + sqrt_fast(1.0 + zeta * zeta));
              const double cs = rcp(sqrt_fast(1.0 + t * t)), sn = cs * t;
              double2* vp = reinterpret_cast<double2*>(V + (size_t)p * ld);
              double2* vq = reinterpret_cast<double2*>(V + (size_t)q * ld);
              for (int i = sub; i < nh; i += g) {
                const double2 x = bp[i], y = bq[i];
                bp[i] = make_double2(cs * x.x - sn * y.x, cs * x.y - sn * y.y);
                bq[i] = make_double2(sn * x.x + cs * y.x, sn * x.y + cs * y.y);
                const double2 vx = vp[i], vy = vq[i];
                vp[i] = make_double2(cs * vx.x - sn * vy.x, cs * vx.y - sn * vy.y);
                vq[i] = make_double2(sn * vx.x + cs * vy.x, sn * vx.y + cs * vy.y);
              }
              if (sub == 0) {
                nrm[p] = aa - t * ab;
                nrm[q] = bb + t * ab;
                // bit 0: some rotation; bit 1: one that was not yet in the quadratic end phase
                atomicOr(flag, (ab * ab > 1e-16 * (aa * bb)) ? 3 : 1);
              }
            }
          }
          __syncthreads();
        }
        const int rotated = *flag;
        __syncthreads();
#ifdef M3D_DEBUG_SWEEPS
        ++dbg_sweeps;
#endif
        // no rotation, or only rotations of relative size <= 1e-8: what they leave behind is of
        // second order (<= 1e-16), the next sweep would not rotate
        if (!(rotated & 2)) break;
      }
      // shrink weights  max(s_j - lambda/mu, 0) / s_j
      const double tau = lambda / mu;
      for (int j = tid; j < ne; j += SVT_THREADS) {
        double ss = 0.0;
        for (int i = 0; i < ne; ++i) ss += B[j * ld + i] * B[j * ld + i];
        const double sv = sqrt(ss);
        sig[j] = (sv > tau) ? (sv - tau) / sv : 0.0;
      }
      __syncthreads();
      // Q = sum_j w_j b_j v_j^T ; X = Q - (Wm + Y)/mu, zero same-camera blocks, diag = 1, clip
      for (int e = tid; e < n * n; e += SVT_THREADS) {
        const int i = e / n, j = e % n;
        double q = 0.0;
        for (int c = 0; c < ne; ++c) {
          const double w = sig[c];
          if (w != 0.0) q += w * B[c * ld + i] * V[c * ld + j];
        }
        Q[e] = q;
        double x = q - (Wm[e] + Y[e]) / mu;
        int ci = -1, cj = -1;
        for (int c = 0; c <= C; ++c) {
          if (dg[c] <= i) ci = c;
          if (dg[c] <= j) cj = c;
        }
        if (ci == cj) x = 0.0;             // X[i0:i1, i0:i1] = 0  (step2:170-172)
        if (i == j) x = 1.0;               // pselect == 1
        x = x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x);
        X[e] = x;
      }
      __syncthreads();
      // X = (X + X^T)/2 ; Y += mu (X - Q) ; residuals
      double pr = 0.0, dr = 0.0;
      for (int e = tid; e < n * n; e += SVT_THREADS) {
        const int i = e / n, j = e % n;
        if (i <= j) {
          const double xs = (X[i * n + j] + X[j * n + i]) / 2.0;
          // stage the symmetric value in the unused half of B (shared) to avoid a race
          B[(size_t)i * ld + j] = xs;
        }
      }
      __syncthreads();
      for (int e = tid; e < n * n; e += SVT_THREADS) {
        const int i = e / n, j = e % n;
        const double xs = (i <= j) ? B[(size_t)i * ld + j] : B[(size_t)j * ld + i];
        const double q = Q[e];
        X[e] = xs;
        Y[e] += mu * (xs - q);
        pr += (xs - q) * (xs - q);
        const double d = xs - X0[e];
        dr += d * d;
      }
      const double pRes = sqrt(block_sum<SVT_THREADS>(pr, red)) / (double)n;
      const double dRes = mu * sqrt(block_sum<SVT_THREADS>(dr, red)) / (double)n;
      if (pRes < tol && dRes < tol) break;
      if (pRes > 10.0 * dRes) mu *= 2.0;
      else if (dRes > 10.0 * pRes) mu /= 2.0;
    }
    __syncthreads();
    for (int e = tid; e < n * n; e += SVT_THREADS) {
      const int i = e / n, j = e % n;
      const double xs = (X[i * n + j] + X[j * n + i]) / 2.0;
      out[i * M + j] = xs > 0.5 ? 1 : 0;
    }
#ifdef M3D_DEBUG_SWEEPS
    if (iters && tid == 0) iters[f] = dbg_sweeps;
#else
    if (iters && tid == 0) iters[f] = it < max_iter ? it : max_iter - 1;
#endif
    __syncthreads();
  }
}

// W = alpha_id * [same identity, different cameras] + (1 - alpha_id) * aff, zero where aff <= 0 or NaN
// (MultiEstimator.predict_data, step2_crossviewmatching.py:557-575); thread = matrix entry
__global__ void __launch_bounds__(256)
k_assoc_weights(const double* __restrict__ aff, const int32_t* __restrict__ cid, const int32_t* __restrict__ dim,
                int F, int M, int C, double alpha_id, double* __restrict__ W) {
  const int64_t n = (int64_t)F * M * M;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(t / ((int64_t)M * M));
    const int r = (int)(t % ((int64_t)M * M));
    const int i = r / M, j = r % M;
    const int32_t* dg = dim + (int64_t)f * (C + 1);
    const int nd = dg[C];
    double w = 0.0;
    if (i < nd && j < nd) {
      const double a = aff[t];
      if (a > 0.0) {  // also drops NaN (np.nan_to_num after the mask, :574-575)
        int ci = 0, cj = 0;
        for (int c = 1; c <= C; ++c) {
          ci += (i >= dg[c]);
          cj += (j >= dg[c]);
        }
        const int32_t idi = cid[(int64_t)f * M + i], idj = cid[(int64_t)f * M + j];
        const double same = (ci != cj && idi >= 0 && idi == idj) ? 1.0 : 0.0;
        w = alpha_id * same + (1.0 - alpha_id) * a;
      }
    }
    W[t] = w;
  }
}

// Person clusters of a match matrix (step2_crossviewmatching.py:598-607): columns whose sum is > 1.9
// are persons; a detection belongs to the first such column it is matched to.  One warp per frame.
__global__ void __launch_bounds__(128)
k_match_clusters(const uint8_t* __restrict__ match, const int32_t* __restrict__ dim, int F, int M, int C,
                 int32_t* __restrict__ label) {
  const int lane = threadIdx.x & 31;
  const int warps = (blockDim.x >> 5) * gridDim.x;
  for (int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); f < F; f += warps) {
    const int nd = dim[(int64_t)f * (C + 1) + C];
    const uint8_t* m = match + (int64_t)f * M * M;
    // person columns as bit masks (M <= 128): lane handles columns lane, lane + 32, ...
    uint32_t is_person[4] = {0, 0, 0, 0};
    for (int j = lane; j < nd; j += 32) {
      int sum = 0;
      for (int i = 0; i < nd; ++i) sum += m[(int64_t)i * M + j];
      if (sum >= 2) is_person[j >> 5] = 1u;  // lane-local flag for column j = 32 * (j >> 5) + lane
    }
    uint32_t pm[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) pm[w] = __ballot_sync(0xffffffffu, is_person[w] != 0);
    for (int i = lane; i < M; i += 32) {
      int lab = -1;
      if (i < nd) {
        for (int j = 0; j < nd && lab < 0; ++j)
          if (((pm[j >> 5] >> (j & 31)) & 1u) && m[(int64_t)i * M + j]) lab = j;
      }
      label[(int64_t)f * M + i] = lab;
    }
  }
}

// Persons of every keyframe as member tables (step2_crossviewmatching.py:598-607, 697-698): a person is a
// cluster (label column) with detections from at least two cameras; members[p][c] = the detection of camera
// c, -1 = none.  A frame with a cluster that holds TWO detections of one camera (get_best_comb territory,
// :610-657) is flagged in `dup` and contributes no person here - the host resolves those frames.
// One warp per frame, lane = label column (j, j + 32, ...).  Two passes of the same kernel: offsets == NULL
// counts (count[f], dup[f]); with the exclusive prefix sums of the counts it writes frame / column / members
// in (frame, column) order.
__global__ void __launch_bounds__(128)
k_cluster_members(const int32_t* __restrict__ label, const int32_t* __restrict__ dim, int F, int M, int C,
                  const int64_t* __restrict__ offsets, int32_t* __restrict__ count, uint8_t* __restrict__ dup,
                  int32_t* __restrict__ frame, int32_t* __restrict__ column, int32_t* __restrict__ members) {
  const int lane = threadIdx.x & 31;
  const int warps = (blockDim.x >> 5) * gridDim.x;
  for (int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); f < F; f += warps) {
    const int32_t* dg = dim + (int64_t)f * (C + 1);
    const int32_t* lab = label + (int64_t)f * M;
    const int nd = dg[C];
    if (offsets && dup[f]) continue;  // warp-uniform
    int total = 0;
    bool any_dup = false;
    const int64_t base = offsets ? offsets[f] : 0;
    for (int j0 = 0; j0 < nd; j0 += 32) {
      const int j = j0 + lane;
      int mem[M3D_MAX_CAMS];
#pragma unroll
      for (int c = 0; c < M3D_MAX_CAMS; ++c) mem[c] = -1;
      int ncam = 0;
      bool d = false;
      if (j < nd) {
        int c = 0;
        for (int i = 0; i < nd; ++i) {
          while (c < C && i >= dg[c + 1]) ++c;  // camera of detection i (detections are grouped by camera)
          if (lab[i] == j) {
#pragma unroll
            for (int k = 0; k < M3D_MAX_CAMS; ++k) {
              if (k == c) {
                if (mem[k] >= 0) d = true; else { mem[k] = i; ++ncam; }
              }
            }
          }
        }
      }
      const bool person = ncam >= 2;
      const unsigned pb = __ballot_sync(0xffffffffu, person);
      any_dup = any_dup || __any_sync(0xffffffffu, d);
      if (offsets && person) {
        const int64_t p = base + total + __popc(pb & ((1u << lane) - 1u));
        frame[p] = f;
        column[p] = j;
#pragma unroll
        for (int c = 0; c < M3D_MAX_CAMS; ++c)
          if (c < C) members[p * C + c] = mem[c];
      }
      total += __popc(pb);
    }
    if (!offsets && lane == 0) {
      count[f] = any_dup ? 0 : total;
      dup[f] = any_dup ? 1 : 0;
    }
  }
}

extern "C" {

int m3d_cluster_members(const int32_t* label, const int32_t* dim, int32_t F, int32_t M, int32_t C,
                        const int64_t* offsets, int32_t* count, uint8_t* dup, int32_t* frame, int32_t* column,
                        int32_t* members, int32_t device, void* stream) {
  if (F < 0 || M < 0 || C < 0 || C > M3D_MAX_CAMS) return m3d_fail(M3D_ERR_INVALID, "m3d_cluster_members: bad size");
  if (F == 0) return M3D_OK;
  if (!label || !dim || !dup || (!offsets && !count) || (offsets && (!frame || !column || !members)))
    return m3d_fail(M3D_ERR_INVALID, "m3d_cluster_members: NULL buffer");
  M3dDeviceGuard guard(device);
  const int blocks = (F + 3) / 4;
  k_cluster_members<<<blocks, 128, 0, (cudaStream_t)stream>>>(label, dim, F, M, C, offsets, count, dup, frame, column,
                                                             members);
  return m3d_check_launch("k_cluster_members");
}

int m3d_association_weights(const double* aff, const int32_t* cid, const int32_t* dim, int32_t F, int32_t M,
                            int32_t C, double alpha_id, double* W, int32_t device, void* stream) {
  if (F < 0 || M < 0 || C < 0 || C > M3D_MAX_CAMS) return m3d_fail(M3D_ERR_INVALID, "m3d_association_weights: bad size");
  if (F == 0 || M == 0) return M3D_OK;
  if (!aff || !cid || !dim || !W) return m3d_fail(M3D_ERR_INVALID, "m3d_association_weights: NULL buffer");
  M3dDeviceGuard guard(device);
  int64_t blocks = ((int64_t)F * M * M + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  k_assoc_weights<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(aff, cid, dim, F, M, C, alpha_id, W);
  return m3d_check_launch("k_assoc_weights");
}

int m3d_match_clusters(const uint8_t* match, const int32_t* dim, int32_t F, int32_t M, int32_t C, int32_t* label,
                       int32_t device, void* stream) {
  if (F < 0 || M < 0 || C < 0 || C > M3D_MAX_CAMS || M > M3D_MAX_DETS)
    return m3d_fail(M3D_ERR_INVALID, "m3d_match_clusters: bad size");
  if (F == 0 || M == 0) return M3D_OK;
  if (!match || !dim || !label) return m3d_fail(M3D_ERR_INVALID, "m3d_match_clusters: NULL buffer");
  M3dDeviceGuard guard(device);
  int blocks = (F + 3) / 4;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_match_clusters<<<blocks, 128, 0, (cudaStream_t)stream>>>(match, dim, F, M, C, label);
  return m3d_check_launch("k_match_clusters");
}

int m3d_ray_affinity(const m3d_rig* rig, const double* kp, const int32_t* dim, int32_t F, int32_t M,
                     int32_t J, double thr_kp, double* aff, double* dist, void* stream) {
  if (!rig) return m3d_fail(M3D_ERR_INVALID, "m3d_ray_affinity: rig is NULL");
  if (F < 0 || M < 0 || J < 0) return m3d_fail(M3D_ERR_INVALID, "m3d_ray_affinity: negative size");
  if (M > M3D_MAX_DETS || J > M3D_MAX_JOINTS)
    return m3d_fail(M3D_ERR_INVALID, "m3d_ray_affinity: M or J above the compiled limits");
  if (F == 0 || M == 0) return M3D_OK;
  if (!kp || !dim || !aff) return m3d_fail(M3D_ERR_INVALID, "m3d_ray_affinity: NULL buffer");
  const RigDev* dev = m3d_rig_dev(rig);
  M3dDeviceGuard guard(m3d_rig_device(rig));
  const size_t smem = sizeof(double) * ((size_t)M * J * 4 + 3 * M3D_MAXC + 32) + sizeof(int) * M;
  cudaError_t e = cudaFuncSetAttribute(k_ray_affinity, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return m3d_fail(M3D_ERR_CUDA, std::string("k_ray_affinity smem: ") + cudaGetErrorString(e));
  k_ray_affinity<<<F, AFF_THREADS, smem, (cudaStream_t)stream>>>(*dev, kp, dim, M, J, thr_kp, aff, dist);
  return m3d_check_launch("k_ray_affinity");
}

int m3d_match_svt(const double* W, const int32_t* dim, int32_t F, int32_t M, int32_t C, double alpha,
                  double lambda, double mu, double tol, int32_t max_iter, uint8_t* match,
                  int32_t* iters, int32_t device, void* stream) {
  if (F < 0 || M < 0 || C < 0) return m3d_fail(M3D_ERR_INVALID, "m3d_match_svt: negative size");
  if (M > SVT_MAX_M)
    return m3d_fail(M3D_ERR_INVALID, "m3d_match_svt: more than " + std::to_string(SVT_MAX_M) +
                                         " detections per frame are not supported");
  if (C > M3D_MAX_CAMS) return m3d_fail(M3D_ERR_INVALID, "m3d_match_svt: too many cameras");
  if (F == 0 || M == 0) return M3D_OK;
  if (!W || !dim || !match) return m3d_fail(M3D_ERR_INVALID, "m3d_match_svt: NULL buffer");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return m3d_fail(M3D_ERR_NO_GPU, "m3d_match_svt: no CUDA device visible; libm3d has no CPU fallback");
  M3dDeviceGuard guard(device);
  cudaStream_t st = (cudaStream_t)stream;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int ld = (M + 2) & ~1;  // even column stride: 16-byte aligned columns for the double2 accesses of the Jacobi sweeps
  const size_t smem = sizeof(double) * (2 * (size_t)ld * (M + 1) + 3 * (size_t)(M + 1) + 64);
  cudaError_t e = cudaFuncSetAttribute(k_match_svt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return m3d_fail(M3D_ERR_CUDA, std::string("k_match_svt smem: ") + cudaGetErrorString(e));
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  static const int cap = [] { const char* e = getenv("M3D_SVT_CTAS"); return e ? atoi(e) : 5; }();
  if (per_sm > cap) per_sm = cap;
  int grid = sms * per_sm;
  if (grid > F) grid = F;
  double* ws = nullptr;  // per-CTA global scratch: X, X0, Y, Wm, Q
  cudaMemPool_t pool = m3d_scratch_pool(device);
  e = pool ? cudaMallocFromPoolAsync(&ws, sizeof(double) * 5 * (size_t)M * M * grid, pool, st)
           : cudaMallocAsync(&ws, sizeof(double) * 5 * (size_t)M * M * grid, st);
  if (e != cudaSuccess) return m3d_fail(M3D_ERR_CUDA, std::string("m3d_match_svt workspace: ") + cudaGetErrorString(e));
  k_match_svt<<<grid, SVT_THREADS, smem, st>>>(W, dim, F, M, C, alpha, lambda, mu, tol, max_iter, match, iters, ws, ld);
  int rc = m3d_check_launch("k_match_svt");
  cudaFreeAsync(ws, st);
  return rc;
}

}  // extern "C"
