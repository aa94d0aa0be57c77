// m3d_handle.h — the rig handle and helpers shared by the translation units of libm3d.so
// (m3d_kernels.cu: per-point maps, triangulation, host pipelines; m3d_ransac.cu: subset search).
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <string>

#include "../../include/m3d.h"
#include "m3d_cert.h"
#include "m3d_internal.h"
#include "m3d_math.cuh"

#define M3D_CUDA(expr)                                                                    \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess)                                                               \
      return m3d_fail(M3D_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));     \
  } while (0)

struct m3d_rig {
  m3d::RigDev dev;
  m3d::RigDev* dev_g = nullptr;  // device-resident copy (kernels that index cameras per lane)
  // pair certificates of the subset search (m3d_cert.h); cert_all: every camera is certified, the
  // rig takes the pruned search k_ransac_cert instead of the exhaustive kernels
  m3d::CertDev cert;
  bool cert_all = false;
  int ransac_mode = 0;  // M3D_RANSAC_AUTO / M3D_RANSAC_EXHAUSTIVE
  uint32_t* cumb_g = nullptr;  // cumulative binomials (m3d_ransac_cert.cuh)
  int device;
  // stream-ordered scratch of the RANSAC launches comes from a private pool that keeps its
  // memory between launches (the default pool trims at every synchronisation)
  cudaMemPool_t pool = nullptr;
  std::mutex pool_mutex;
  // workspace of the *_host pipelines (lazily allocated, guarded by ws_mutex)
  std::mutex ws_mutex;
  static const int kSlots = 3;
  int64_t ws_chunk = 0;
  int ws_cams = 0;
  double* ws_xy[kSlots] = {nullptr, nullptr, nullptr};
  double* ws_p3d[kSlots] = {nullptr, nullptr, nullptr};
  double* ws_err[kSlots] = {nullptr, nullptr, nullptr};
  double* ws_xyp[kSlots] = {nullptr, nullptr, nullptr};
  float* ws_xy32[kSlots] = {nullptr, nullptr, nullptr};   // float32 staging of the *_host_f32 entry points
  float* ws_xyp32[kSlots] = {nullptr, nullptr, nullptr};
  uint8_t* ws_picked[kSlots] = {nullptr, nullptr, nullptr};
  int32_t* ws_subset[kSlots] = {nullptr, nullptr, nullptr};
  int32_t* ws_neval[kSlots] = {nullptr, nullptr, nullptr};
  cudaStream_t ws_stream[kSlots] = {nullptr, nullptr, nullptr};
  bool ws_ransac = false;
};


// 16-byte read-only load / store of one (x, y) observation
__device__ __forceinline__ double2 ld_xy(const double* __restrict__ xy, int64_t idx) {
  return __ldg(reinterpret_cast<const double2*>(xy) + idx);
}
__device__ __forceinline__ void st_xy(double* out, int64_t idx, double x, double y) {
  reinterpret_cast<double2*>(out)[idx] = make_double2(x, y);
}

// grid-stride kernels: enough blocks to cover N, capped at 32 waves of resident blocks
inline int m3d_grid_for(int64_t N, int threads, int sm_count) {
  int64_t blocks = (N + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count * 8 * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}
int m3d_sm_count(int device);

#define M3D_DISPATCH_MODEL(rig, CALL)                                      \
  do {                                                                     \
    const bool full__ = ((rig)->dev.flags & (m3d::RIG_HAS_RATIONAL | m3d::RIG_HAS_PRISM)) != 0; \
    const bool po__ = ((rig)->dev.flags & m3d::RIG_HAS_NONPINHOLE) == 0;   \
    if (full__) {                                                          \
      if (po__) { CALL(true, true); } else { CALL(true, false); }          \
    } else {                                                               \
      if (po__) { CALL(false, true); } else { CALL(false, false); }        \
    }                                                                      \
  } while (0)

// subset search (m3d_ransac.cu)
int m3d_launch_ransac(const m3d_rig* rig, const double* xy, int64_t N, int undistort, int min_cams,
                      double threshold, double init_best, double* p3d, uint8_t* picked,
                      double* xy_picked, double* err, int32_t* subset, int32_t* neval, cudaStream_t st);
