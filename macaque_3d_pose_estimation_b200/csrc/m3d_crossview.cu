// m3d_crossview.cu — cross-view association kernels (step2 of the reference pipeline):
//   k_ray_affinity : geometry_affinity2 (step2_crossviewmatching.py:373-432) with
//                    deproject (:327-355) and calc_dist_btw_lines (:359-369) fused, one CTA
//                    per frame, rays staged in shared memory;
//   k_match_svt    : matchSVT (step2_crossviewmatching.py:130-216), one CTA per frame, the
//                    ADMM iterate and a cyclic-Jacobi symmetric eigensolver held in shared
//                    memory (the iterate Y/mu + X stays symmetric, so the reference's SVD
//                    shrinkage U max(s - lambda/mu, 0) V^T equals V sign(L) max(|L| - lambda/mu, 0) V^T).
#include <cuda_runtime.h>

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
    if (i == j) {
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
    if (d < 300.0) {
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

extern "C" {

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
  (void)W; (void)dim; (void)F; (void)M; (void)C; (void)alpha; (void)lambda; (void)mu; (void)tol;
  (void)max_iter; (void)match; (void)iters; (void)device; (void)stream;
  return m3d_fail(M3D_ERR_INVALID, "m3d_match_svt: not implemented yet");
}

}  // extern "C"
