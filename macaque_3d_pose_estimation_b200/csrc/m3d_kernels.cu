// m3d_kernels.cu — sm_100a kernels and the C ABI (include/m3d.h) of the multi-view 3D
// reconstruction hot path.  No tensor cores (largest matrix on the path is 4x4 per
// point); the work is fp64 ALU + HBM streaming:
//   * one thread per joint-instance for undistort / DLT / reprojection error, 16-byte
//     vector loads of the (C,N,2) observation planes (each camera plane is read fully
//     coalesced: 512 B per warp per camera);
//   * the camera rig (<= 16 cameras, 4.2 KB) travels as a __grid_constant__ kernel
//     parameter, i.e. it sits in the constant bank and is read with uniform LDC /
//     constant operands — no global state, no per-launch upload;
//   * subset RANSAC: lane-per-point evaluation of the full camera set, then one warp per
//     surviving point with one camera subset per lane (32 subsets per step), __ballot_sync
//     + __ffs for "first subset under the threshold", shuffle arg-min otherwise.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/m3d.h"
#include "m3d_handle.h"
#include "m3d_internal.h"
#include "m3d_math.cuh"
#include "m3d_point.cuh"
#include "m3d_rig.h"

using namespace m3d;

// ---------------------------------------------------------------------------------------
// error handling / bookkeeping
// ---------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static std::atomic<int64_t> g_launches{0};

// ---- optional per-kernel timing ------------------------------------------------------------
static std::atomic<bool> g_prof{false};
struct ProfRec {
  const char* name;
  cudaEvent_t e0, e1;
};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof_recs;

M3dKernelTimer::M3dKernelTimer(const char* nm, cudaStream_t s) : st(s), name(nm) {
  if (!g_prof.load(std::memory_order_relaxed)) return;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
    if (e0) cudaEventDestroy(e0);
    e0 = e1 = nullptr;
    return;
  }
  cudaEventRecord(e0, st);
}

M3dKernelTimer::~M3dKernelTimer() {
  if (!e0) return;
  cudaEventRecord(e1, st);
  std::lock_guard<std::mutex> lock(g_prof_mu);
  g_prof_recs.push_back({name, e0, e1});
}

int m3d_fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
static int fail(int code, const std::string& msg) { return m3d_fail(code, msg); }

// Stream-ordered scratch of the rig-less entry points (m3d_viterbi_filter, m3d_match_svt): one
// private pool per device that keeps its memory between calls (the default pool trims at every
// synchronisation: measured 11 ms per call for the 261 MB code array of the Viterbi filter).
cudaMemPool_t m3d_scratch_pool(int device) {
  static std::mutex mu;
  static cudaMemPool_t pools[64] = {};
  if (device < 0 || device >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!pools[device]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    if (cudaMemPoolCreate(&pools[device], &props) != cudaSuccess) {
      pools[device] = nullptr;
      return nullptr;
    }
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pools[device], cudaMemPoolAttrReleaseThreshold, &keep);
  }
  return pools[device];
}

typedef M3dDeviceGuard DeviceGuard;

const RigDev* m3d_rig_dev(const m3d_rig* rig) { return &rig->dev; }

// ---------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------
// K1: undistort (C,N,2) -> (C,N,2); blockIdx.y selects the camera
// ---------------------------------------------------------------------------------------
template <bool FULL, bool PO>
__global__ void __launch_bounds__(256)
k_undistort(const __grid_constant__ RigDev rig, int cam0, const double* __restrict__ xy,
            int64_t N, double* __restrict__ out) {
  const int c = cam0 + blockIdx.y;
  const int64_t plane = (int64_t)blockIdx.y * N;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N;
       n += (int64_t)gridDim.x * blockDim.x) {
    const double2 p = ld_xy(xy, plane + n);
    double x, y;
    undistort_point<FULL, PO>(rig.cam[c], p.x, p.y, x, y);
    st_xy(out, plane + n, x, y);
  }
}

// Camera.distort_points for one camera: normalised (n,2) -> pixels (n,2)
template <bool FULL>
__global__ void __launch_bounds__(256)
k_distort(const __grid_constant__ RigDev rig, int cam, const double* __restrict__ xy, int64_t N,
          double* __restrict__ out) {
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N;
       n += (int64_t)gridDim.x * blockDim.x) {
    const double2 p = ld_xy(xy, n);
    double u, v;
    distort_point<FULL>(rig.cam[cam], p.x, p.y, u, v);
    st_xy(out, n, u, v);
  }
}

// ---------------------------------------------------------------------------------------
// K-project: (N,3) -> (n_out_cams,N,2); one thread per point, loop over cameras
// ---------------------------------------------------------------------------------------
template <bool FULL, bool PO>
__global__ void __launch_bounds__(256)
k_project(const __grid_constant__ RigDev rig, int cam0, int ncam, const double* __restrict__ p3d,
          int64_t N, double* __restrict__ out) {
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N;
       n += (int64_t)gridDim.x * blockDim.x) {
    const double X = __ldg(p3d + 3 * n), Y = __ldg(p3d + 3 * n + 1), Z = __ldg(p3d + 3 * n + 2);
#pragma unroll 1
    for (int j = 0; j < ncam; ++j) {
      double u, v;
      project_point<FULL, PO>(rig.cam[cam0 + j], X, Y, Z, u, v);
      st_xy(out, (int64_t)j * N + n, u, v);
    }
  }
}

// ---------------------------------------------------------------------------------------
// K2: fused undistort + DLT triangulation (+ mean reprojection error)
// ---------------------------------------------------------------------------------------
// NC > 0: camera count known at compile time — the camera loops unroll completely, every
// camera parameter becomes a constant-bank operand of the DFMA that uses it (no LDC, no
// register), and the raw observations stay in registers for the error pass.
// NC == 0: run-time camera count (any rig up to M3D_MAX_CAMS).
// FAST: undistort_pinhole_fast (float32 in the first three iterations; opt-in, pinhole-only rigs
// without rational / thin-prism terms).
template <bool FULL, bool PO, bool UNDISTORT, bool WITH_ERR, int NC, int MINB, int FAST = 0>
__global__ void __launch_bounds__(256, MINB)
k_triangulate(const __grid_constant__ RigDev rig, const double* __restrict__ xy, int64_t N,
              double* __restrict__ p3d, double* __restrict__ err) {
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N;
       n += (int64_t)gridDim.x * blockDim.x) {
    Gram G;
    gram_zero(G);
    int cnt = 0;
    double X = qnan(), Y = qnan(), Z = qnan();
    if (NC > 0) {
      double2 raw[NC > 0 ? NC : 1];
#pragma unroll
      for (int c = 0; c < NC; ++c) raw[c] = ld_xy(xy, (int64_t)c * N + n);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        double x = raw[c].x, y = raw[c].y;
        if (UNDISTORT) {
          if (FAST) undistort_pinhole_fast<FAST>(rig.cam[c], raw[c].x, raw[c].y, x, y);
          else undistort_point<FULL, PO>(rig.cam[c], raw[c].x, raw[c].y, x, y);
        }
        if (x == x) {  // validity on x only (cameras.py:630)
          gram_add_camera(G, rig.cam[c], x, y);
          ++cnt;
        }
      }
      if (cnt >= 2) dlt_solve(G, X, Y, Z);
      if (WITH_ERR) {
        double sum = 0.0;
        int m = 0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          double u, v;
          project_point<FULL, PO>(rig.cam[c], X, Y, Z, u, v);
          const double e = residual_norm(raw[c].x - u, raw[c].y - v);
          if (e == e) {
            sum += e;
            ++m;
          }
        }
        err[n] = (m >= 2) ? sum / (double)m : qnan();
      }
    } else {
      const int C = rig.n_cams;
#pragma unroll 1
      for (int c = 0; c < C; ++c) {
        const double2 p = ld_xy(xy, (int64_t)c * N + n);
        double x = p.x, y = p.y;
        if (UNDISTORT) {
          if (FAST) undistort_pinhole_fast<FAST>(rig.cam[c], p.x, p.y, x, y);
          else undistort_point<FULL, PO>(rig.cam[c], p.x, p.y, x, y);
        }
        if (x == x) {
          gram_add_camera(G, rig.cam[c], x, y);
          ++cnt;
        }
      }
      if (cnt >= 2) dlt_solve(G, X, Y, Z);
      if (WITH_ERR) {
        double sum = 0.0;
        int m = 0;
#pragma unroll 1
        for (int c = 0; c < C; ++c) {
          const double2 p = ld_xy(xy, (int64_t)c * N + n);
          double u, v;
          project_point<FULL, PO>(rig.cam[c], X, Y, Z, u, v);
          const double e = residual_norm(p.x - u, p.y - v);
          if (e == e) {
            sum += e;
            ++m;
          }
        }
        err[n] = (m >= 2) ? sum / (double)m : qnan();
      }
    }
    p3d[3 * n] = X;
    p3d[3 * n + 1] = Y;
    p3d[3 * n + 2] = Z;
  }
}

// ---------------------------------------------------------------------------------------
// K3: reprojection error, full residuals (C,N,2) or mean (N)
// ---------------------------------------------------------------------------------------
template <bool FULL, bool PO, bool MEAN>
__global__ void __launch_bounds__(256)
k_reproj(const __grid_constant__ RigDev rig, const double* __restrict__ p3d,
         const double* __restrict__ xy, int64_t N, double* __restrict__ out) {
  const int C = rig.n_cams;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N;
       n += (int64_t)gridDim.x * blockDim.x) {
    const double X = __ldg(p3d + 3 * n), Y = __ldg(p3d + 3 * n + 1), Z = __ldg(p3d + 3 * n + 2);
    double sum = 0.0;
    int m = 0;
#pragma unroll 1
    for (int c = 0; c < C; ++c) {
      const double2 p = ld_xy(xy, (int64_t)c * N + n);
      double u, v;
      project_point<FULL, PO>(rig.cam[c], X, Y, Z, u, v);
      const double ex = p.x - u, ey = p.y - v;
      if (MEAN) {
        const double e = residual_norm(ex, ey);
        if (e == e) {
          sum += e;
          ++m;
        }
      } else {
        st_xy(out, (int64_t)c * N + n, ex, ey);
      }
    }
    if (MEAN) out[n] = (m >= 2) ? sum / (double)m : qnan();
  }
}

// ---------------------------------------------------------------------------------------
// K7: inhomogeneous least-squares triangulation (mct.triangulatePoints)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_triangulate_ls(const __grid_constant__ RigDev rig, const double* __restrict__ xy,
                 const uint8_t* __restrict__ use, int64_t N, double* __restrict__ p3d) {
  const int C = rig.n_cams;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N;
       n += (int64_t)gridDim.x * blockDim.x) {
    Gram G;
    gram_zero(G);
    int cnt = 0;
#pragma unroll 1
    for (int c = 0; c < C; ++c) {
      if (use[(int64_t)c * N + n]) {
        const double2 p = ld_xy(xy, (int64_t)c * N + n);
        gram_add_camera(G, rig.cam[c], p.x, p.y);
        ++cnt;
      }
    }
    double X = qnan(), Y = qnan(), Z = qnan();
    if (cnt >= 2) ls_solve(G, X, Y, Z);
    p3d[3 * n] = X;
    p3d[3 * n + 1] = Y;
    p3d[3 * n + 2] = Z;
  }
}

// K7b: undistortion of every detection of F keyframes (step2_crossviewmatching.py:306-325 applied to all
// detections of a frame, :520-530): thread = keypoint; the camera of a detection slot follows from the
// frame's cumulative counts; padding slots give NaN.  The score is copied.
template <bool FULL, bool PO>
__global__ void __launch_bounds__(256)
k_undistort_dets(const RigDev* __restrict__ rig, const double* __restrict__ kp_raw, const int32_t* __restrict__ dim,
                 int64_t F, int M, int J, double* __restrict__ kp_und) {
  const int C = rig->n_cams;
  const int64_t n = F * M * J;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t fm = t / J;
    const int64_t f = fm / M;
    const int m = (int)(fm - f * M);
    const int32_t* dg = dim + f * (C + 1);
    int c = 0;
    for (int k = 1; k <= C; ++k) c += (m >= dg[k]);
    const double u = kp_raw[3 * t], v = kp_raw[3 * t + 1];
    double x = qnan(), y = qnan();
    if (c < C) undistort_point<FULL, PO>(rig->cam[c], u, v, x, y);
    kp_und[3 * t] = x;
    kp_und[3 * t + 1] = y;
    kp_und[3 * t + 2] = kp_raw[3 * t + 2];
  }
}

// np.nan_to_num of one value
__device__ __forceinline__ double nan_to_num(double v) {
  if (v != v) return 0.0;
  if (v > 1.7976931348623157e308) return 1.7976931348623157e308;
  if (v < -1.7976931348623157e308) return -1.7976931348623157e308;
  return v;
}

// K7c: calc_3dpose (step2_crossviewmatching.py:436-461) of P persons given as member tables: thread =
// (person, keypoint); camera c contributes the keypoint of detection members[p][c] of frame[p] when its
// undistorted x is not NaN and its score is >= thr_kp.
__global__ void __launch_bounds__(256)
k_triangulate_ls_members(const RigDev* __restrict__ rig, const double* __restrict__ kp_und,
                         const int32_t* __restrict__ frame, const int32_t* __restrict__ members, int64_t P, int M,
                         int J, double thr_kp, double* __restrict__ p3d) {
  const int C = rig->n_cams;
  const int64_t n = P * J;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = t / J;
    const int j = (int)(t - p * J);
    const int64_t f = frame[p];
    Gram G;
    gram_zero(G);
    int cnt = 0;
#pragma unroll 1
    for (int c = 0; c < C; ++c) {
      const int m = members[p * C + c];
      if (m < 0) continue;
      const double* q = kp_und + 3 * ((f * M + m) * J + j);
      const double x = q[0], y = q[1], sc = q[2];
      if (x == x && sc >= thr_kp) {  // ~(isnan(x) | sc < thr | isnan(sc))
        gram_add_camera(G, rig->cam[c], nan_to_num(x), nan_to_num(y));
        ++cnt;
      }
    }
    double X = qnan(), Y = qnan(), Z = qnan();
    if (cnt >= 2) ls_solve(G, X, Y, Z);
    p3d[3 * t] = X;
    p3d[3 * t + 1] = Y;
    p3d[3 * t + 2] = Z;
  }
}

// ---------------------------------------------------------------------------------------
// fp64 FMA peak probe (the second roofline of this path)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dfma_probe(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, b = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, b);
    a1 = fma(a1, m, b);
    a2 = fma(a2, m, b);
    a3 = fma(a3, m, b);
    a4 = fma(a4, m, b);
    a5 = fma(a5, m, b);
    a6 = fma(a6, m, b);
    a7 = fma(a7, m, b);
  }
  out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ---------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------
static int grid_for(int64_t N, int threads, int sm_count) { return m3d_grid_for(N, threads, sm_count); }

int m3d_sm_count(int device) {
  static int cached[64] = {0};
  if (device >= 0 && device < 64 && cached[device]) return cached[device];
  int v = 148;
  cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
  if (device >= 0 && device < 64) cached[device] = v;
  return v;
}

int m3d_check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(M3D_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return M3D_OK;
}
static int check_launch(const char* what) { return m3d_check_launch(what); }
static int sm_count_of(int device) { return m3d_sm_count(device); }

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" {

int m3d_version(void) { return 100; }

const char* m3d_last_error(void) { return g_last_error.c_str(); }

int m3d_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int64_t m3d_launch_count(void) { return g_launches.load(); }

int m3d_profile_enable(int32_t on) {
  g_prof.store(on != 0);
  return M3D_OK;
}

int m3d_profile_read(char* buf, int64_t cap) {
  if (!buf || cap < 3) return fail(M3D_ERR_INVALID, "m3d_profile_read: buffer too small");
  std::vector<ProfRec> recs;
  {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    recs.swap(g_prof_recs);
  }
  struct Acc {
    std::string name;
    int64_t n = 0;
    double ms = 0.0;
  };
  std::vector<Acc> acc;
  for (const ProfRec& r : recs) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      size_t i = 0;
      for (; i < acc.size(); ++i)
        if (acc[i].name == r.name) break;
      if (i == acc.size()) {
        acc.emplace_back();
        acc[i].name = r.name;
      }
      acc[i].n += 1;
      acc[i].ms += ms;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  std::string out = "{";
  for (size_t i = 0; i < acc.size(); ++i) {
    char tmp[256];
    snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"launches\": %lld, \"ms\": %.6f}", i ? ", " : "", acc[i].name.c_str(),
             (long long)acc[i].n, acc[i].ms);
    out += tmp;
  }
  out += "}";
  if ((int64_t)out.size() + 1 > cap) return fail(M3D_ERR_INVALID, "m3d_profile_read: buffer too small");
  memcpy(buf, out.c_str(), out.size() + 1);
  return M3D_OK;
}

int m3d_rig_create(const m3d_cam* cams, int32_t n_cams, int32_t device, m3d_rig** out) {
  if (!out) return fail(M3D_ERR_INVALID, "m3d_rig_create: out is NULL");
  *out = nullptr;
  int ndev = m3d_device_count();
  if (ndev <= 0)
    return fail(M3D_ERR_NO_GPU, "m3d_rig_create: no CUDA device visible; libm3d has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(M3D_ERR_INVALID, "m3d_rig_create: bad device index");
  m3d_rig* rig = new m3d_rig();
  std::string why = build_rig(cams, n_cams, &rig->dev);
  if (!why.empty()) {
    delete rig;
    return fail(M3D_ERR_INVALID, "m3d_rig_create: " + why);
  }
  rig->device = device;
  {
    DeviceGuard g(device);
    cudaError_t e = cudaMalloc(&rig->dev_g, sizeof(RigDev));
    if (e == cudaSuccess) e = cudaMemcpy(rig->dev_g, &rig->dev, sizeof(RigDev), cudaMemcpyHostToDevice);
    static const CumBinom cumb_host = make_cumbinom();
    if (e == cudaSuccess) e = cudaMalloc(&rig->cumb_g, sizeof(CumBinom));
    if (e == cudaSuccess) e = cudaMemcpy(rig->cumb_g, &cumb_host, sizeof(CumBinom), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      cudaFree(rig->dev_g);
      cudaFree(rig->cumb_g);
      delete rig;
      return fail(M3D_ERR_CUDA, std::string("m3d_rig_create: ") + cudaGetErrorString(e));
    }
  }
  build_cert(rig->dev, &rig->cert);
  // pruned search when at most one camera lacks a certificate (an uncertified camera simply never
  // takes part in a certified-bad pair: the search stays exact, its subsets are solved)
  int n_cert = 0;
  for (int c = 0; c < n_cams; ++c) n_cert += (rig->cert.ok_mask >> c) & 1;
  rig->cert_all = n_cams >= 2 && n_cert >= n_cams - 1 && n_cert >= 2 &&
                  (rig->dev.flags & (RIG_HAS_RATIONAL | RIG_HAS_PRISM)) == 0;
  *out = rig;
  return M3D_OK;
}

static void free_workspace(m3d_rig* rig) {
  for (int i = 0; i < m3d_rig::kSlots; ++i) {
    cudaFree(rig->ws_xy[i]);
    cudaFree(rig->ws_p3d[i]);
    cudaFree(rig->ws_err[i]);
    cudaFree(rig->ws_xyp[i]);
    cudaFree(rig->ws_xy32[i]);
    cudaFree(rig->ws_xyp32[i]);
    rig->ws_xy32[i] = rig->ws_xyp32[i] = nullptr;
    cudaFree(rig->ws_picked[i]);
    cudaFree(rig->ws_subset[i]);
    cudaFree(rig->ws_neval[i]);
    rig->ws_xy[i] = rig->ws_p3d[i] = rig->ws_err[i] = rig->ws_xyp[i] = nullptr;
    rig->ws_picked[i] = nullptr;
    rig->ws_subset[i] = rig->ws_neval[i] = nullptr;
    if (rig->ws_stream[i]) cudaStreamDestroy(rig->ws_stream[i]);
    rig->ws_stream[i] = nullptr;
  }
  rig->ws_chunk = 0;
}

void m3d_rig_destroy(m3d_rig* rig) {
  if (!rig) return;
  {
    DeviceGuard g(rig->device);
    free_workspace(rig);
    cudaFree(rig->dev_g);
    cudaFree(rig->cumb_g);
    if (rig->pool) cudaMemPoolDestroy(rig->pool);
  }
  delete rig;
}

int m3d_rig_set_ransac_mode(m3d_rig* rig, int32_t mode) {
  if (!rig) return fail(M3D_ERR_INVALID, "m3d_rig_set_ransac_mode: rig is NULL");
  if (mode != M3D_RANSAC_AUTO && mode != M3D_RANSAC_EXHAUSTIVE)
    return fail(M3D_ERR_INVALID, "m3d_rig_set_ransac_mode: unknown mode");
  rig->ransac_mode = mode;
  return M3D_OK;
}

int32_t m3d_rig_certified_mask(const m3d_rig* rig) { return rig ? rig->cert.ok_mask : 0; }

int32_t m3d_rig_num_cams(const m3d_rig* rig) { return rig ? rig->dev.n_cams : -1; }
int32_t m3d_rig_device(const m3d_rig* rig) { return rig ? rig->device : -1; }

int m3d_rig_extrinsics(const m3d_rig* rig, double* M) {
  if (!rig || !M) return fail(M3D_ERR_INVALID, "m3d_rig_extrinsics: NULL argument");
  for (int c = 0; c < rig->dev.n_cams; ++c) {
    const CamDev& cam = rig->dev.cam[c];
    double* o = M + 16 * c;
    for (int r = 0; r < 3; ++r) {
      o[4 * r + 0] = cam.R[3 * r + 0];
      o[4 * r + 1] = cam.R[3 * r + 1];
      o[4 * r + 2] = cam.R[3 * r + 2];
      o[4 * r + 3] = cam.t[r];
    }
    o[12] = o[13] = o[14] = 0.0;
    o[15] = 1.0;
  }
  return M3D_OK;
}

#define M3D_CHECK_RIG(name)                                              \
  if (!rig) return fail(M3D_ERR_INVALID, name ": rig is NULL");          \
  if (N < 0) return fail(M3D_ERR_INVALID, name ": negative point count"); \
  DeviceGuard guard__(rig->device);                                      \
  cudaStream_t st = (cudaStream_t)stream;                                \
  const int sms = sm_count_of(rig->device);                              \
  (void)sms;

int m3d_undistort_cam(const m3d_rig* rig, int32_t cam, const double* xy, int64_t N, double* out,
                      void* stream) {
  M3D_CHECK_RIG("m3d_undistort_cam");
  if (cam < 0 || cam >= rig->dev.n_cams) return fail(M3D_ERR_INVALID, "m3d_undistort_cam: bad camera index");
  if (N == 0) return M3D_OK;
  if (!xy || !out) return fail(M3D_ERR_INVALID, "m3d_undistort_cam: NULL buffer");
  dim3 grid(grid_for(N, 256, sms), 1);
#define CALL(F, P) k_undistort<F, P><<<grid, 256, 0, st>>>(rig->dev, cam, xy, N, out)
  M3D_DISPATCH_MODEL(rig, CALL);
#undef CALL
  return check_launch("k_undistort");
}

int m3d_undistort(const m3d_rig* rig, const double* xy, int64_t N, double* out, void* stream) {
  M3D_CHECK_RIG("m3d_undistort");
  if (N == 0 || rig->dev.n_cams == 0) return M3D_OK;
  if (!xy || !out) return fail(M3D_ERR_INVALID, "m3d_undistort: NULL buffer");
  dim3 grid(grid_for(N, 256, sms), rig->dev.n_cams);
#define CALL(F, P) k_undistort<F, P><<<grid, 256, 0, st>>>(rig->dev, 0, xy, N, out)
  M3D_DISPATCH_MODEL(rig, CALL);
#undef CALL
  return check_launch("k_undistort");
}

int m3d_distort_cam(const m3d_rig* rig, int32_t cam, const double* xy, int64_t N, double* out,
                    void* stream) {
  M3D_CHECK_RIG("m3d_distort_cam");
  if (cam < 0 || cam >= rig->dev.n_cams) return fail(M3D_ERR_INVALID, "m3d_distort_cam: bad camera index");
  if (N == 0) return M3D_OK;
  if (!xy || !out) return fail(M3D_ERR_INVALID, "m3d_distort_cam: NULL buffer");
  const int grid = grid_for(N, 256, sms);
  if (rig->dev.flags & (RIG_HAS_RATIONAL | RIG_HAS_PRISM))
    k_distort<true><<<grid, 256, 0, st>>>(rig->dev, cam, xy, N, out);
  else
    k_distort<false><<<grid, 256, 0, st>>>(rig->dev, cam, xy, N, out);
  return check_launch("k_distort");
}

int m3d_project_cam(const m3d_rig* rig, int32_t cam, const double* p3d, int64_t N, double* out,
                    void* stream) {
  M3D_CHECK_RIG("m3d_project_cam");
  if (cam < 0 || cam >= rig->dev.n_cams) return fail(M3D_ERR_INVALID, "m3d_project_cam: bad camera index");
  if (N == 0) return M3D_OK;
  if (!p3d || !out) return fail(M3D_ERR_INVALID, "m3d_project_cam: NULL buffer");
  const int grid = grid_for(N, 256, sms);
#define CALL(F, P) k_project<F, P><<<grid, 256, 0, st>>>(rig->dev, cam, 1, p3d, N, out)
  M3D_DISPATCH_MODEL(rig, CALL);
#undef CALL
  return check_launch("k_project");
}

int m3d_project(const m3d_rig* rig, const double* p3d, int64_t N, double* out, void* stream) {
  M3D_CHECK_RIG("m3d_project");
  if (N == 0 || rig->dev.n_cams == 0) return M3D_OK;
  if (!p3d || !out) return fail(M3D_ERR_INVALID, "m3d_project: NULL buffer");
  const int grid = grid_for(N, 256, sms);
#define CALL(F, P) k_project<F, P><<<grid, 256, 0, st>>>(rig->dev, 0, rig->dev.n_cams, p3d, N, out)
  M3D_DISPATCH_MODEL(rig, CALL);
#undef CALL
  return check_launch("k_project");
}

static int launch_triangulate(const m3d_rig* rig, const double* xy, int64_t N, int undistort,
                              double* p3d, double* err, cudaStream_t st, int sms) {
  const int grid = grid_for(N, 256, sms);
  const int C = rig->dev.n_cams;
  M3dKernelTimer timer__("k_triangulate", st);
  // opt-in north-star-tolerance path (M3D_UNDISTORT_FAST): plain pinhole rigs only, else strict
  const bool fast = (undistort & 2) != 0 && (rig->dev.flags & (RIG_HAS_RATIONAL | RIG_HAS_PRISM | RIG_HAS_NONPINHOLE)) == 0;
  if (fast) {
    // float32 iterations of the five (developer switch M3D_FAST_ITERS = 3 | 4 | 5; default 4)
    static const int fast_iters = [] { const char* e = getenv("M3D_FAST_ITERS"); return e ? atoi(e) : 4; }();
#define CALLF(NC, MB, NF)                                                                                          \
  do {                                                                                                             \
    if (err) k_triangulate<false, true, true, true, NC, MB, NF><<<grid, 256, 0, st>>>(rig->dev, xy, N, p3d, err);  \
    else k_triangulate<false, true, true, false, NC, MB, NF><<<grid, 256, 0, st>>>(rig->dev, xy, N, p3d, err);     \
  } while (0)
    if (C == 8) {
      if (fast_iters == 3) CALLF(8, 2, 3);
      else if (fast_iters == 5) CALLF(8, 2, 5);
      else CALLF(8, 2, 4);
    } else {
      CALLF(0, 3, 4);
    }
#undef CALLF
    return check_launch("k_triangulate");
  }
#define CALLV(F, P, NC, MB)                                                                         \
  do {                                                                                              \
    if (undistort) {                                                                                \
      if (err) k_triangulate<F, P, true, true, NC, MB><<<grid, 256, 0, st>>>(rig->dev, xy, N, p3d, err); \
      else k_triangulate<F, P, true, false, NC, MB><<<grid, 256, 0, st>>>(rig->dev, xy, N, p3d, err);    \
    } else {                                                                                        \
      if (err) k_triangulate<F, P, false, true, NC, MB><<<grid, 256, 0, st>>>(rig->dev, xy, N, p3d, err); \
      else k_triangulate<F, P, false, false, NC, MB><<<grid, 256, 0, st>>>(rig->dev, xy, N, p3d, err);   \
    }                                                                                               \
  } while (0)
  // measured on B200 (profiles/): the 8-camera unrolled kernel is fastest at 128 registers
  // (2 CTAs/SM, no spills); the run-time-C kernel at 80 registers (3 CTAs/SM)
#define CALL(F, P)                                          \
  do {                                                      \
    if (C == 8) CALLV(F, P, 8, 2);                          \
    else CALLV(F, P, 0, 3);                                 \
  } while (0)
  M3D_DISPATCH_MODEL(rig, CALL);
#undef CALL
#undef CALLV
  return check_launch("k_triangulate");
}

int m3d_triangulate(const m3d_rig* rig, const double* xy, int64_t N, int32_t undistort, double* p3d,
                    void* stream) {
  M3D_CHECK_RIG("m3d_triangulate");
  if (N == 0) return M3D_OK;
  if (!p3d || (!xy && rig->dev.n_cams > 0)) return fail(M3D_ERR_INVALID, "m3d_triangulate: NULL buffer");
  return launch_triangulate(rig, xy, N, undistort, p3d, nullptr, st, sms);
}

int m3d_triangulate_error(const m3d_rig* rig, const double* xy, int64_t N, int32_t undistort,
                          double* p3d, double* err, void* stream) {
  M3D_CHECK_RIG("m3d_triangulate_error");
  if (N == 0) return M3D_OK;
  if (!p3d || (!xy && rig->dev.n_cams > 0))
    return fail(M3D_ERR_INVALID, "m3d_triangulate_error: NULL buffer");
  return launch_triangulate(rig, xy, N, undistort, p3d, err, st, sms);
}

int m3d_reproj_error(const m3d_rig* rig, const double* p3d, const double* xy, int64_t N,
                     int32_t mean, double* out, void* stream) {
  M3D_CHECK_RIG("m3d_reproj_error");
  if (N == 0) return M3D_OK;
  if (!p3d || !out || (!xy && rig->dev.n_cams > 0))
    return fail(M3D_ERR_INVALID, "m3d_reproj_error: NULL buffer");
  if (!mean && rig->dev.n_cams == 0) return M3D_OK;
  const int grid = grid_for(N, 256, sms);
#define CALL(F, P)                                                                     \
  do {                                                                                 \
    if (mean) k_reproj<F, P, true><<<grid, 256, 0, st>>>(rig->dev, p3d, xy, N, out);   \
    else k_reproj<F, P, false><<<grid, 256, 0, st>>>(rig->dev, p3d, xy, N, out);       \
  } while (0)
  M3D_DISPATCH_MODEL(rig, CALL);
#undef CALL
  return check_launch("k_reproj");
}

int m3d_triangulate_ls(const m3d_rig* rig, const double* xy, const uint8_t* use, int64_t N,
                       double* p3d, void* stream) {
  M3D_CHECK_RIG("m3d_triangulate_ls");
  if (N == 0) return M3D_OK;
  if (!p3d || ((!xy || !use) && rig->dev.n_cams > 0))
    return fail(M3D_ERR_INVALID, "m3d_triangulate_ls: NULL buffer");
  k_triangulate_ls<<<grid_for(N, 256, sms), 256, 0, st>>>(rig->dev, xy, use, N, p3d);
  return check_launch("k_triangulate_ls");
}

int m3d_undistort_detections(const m3d_rig* rig, const double* kp_raw, const int32_t* dim, int64_t N, int32_t M,
                             int32_t J, double* kp_und, void* stream) {
  M3D_CHECK_RIG("m3d_undistort_detections");
  if (M < 0 || J < 0) return fail(M3D_ERR_INVALID, "m3d_undistort_detections: bad size");
  if (N == 0 || M == 0 || J == 0) return M3D_OK;
  if (!kp_raw || !dim || !kp_und) return fail(M3D_ERR_INVALID, "m3d_undistort_detections: NULL buffer");
  const int grid = grid_for(N * M * J, 256, sms);
#define CALL(F, P) k_undistort_dets<F, P><<<grid, 256, 0, st>>>(rig->dev_g, kp_raw, dim, N, M, J, kp_und)
  M3D_DISPATCH_MODEL(rig, CALL);
#undef CALL
  return check_launch("k_undistort_dets");
}

int m3d_triangulate_ls_members(const m3d_rig* rig, const double* kp_und, const int32_t* frame,
                               const int32_t* members, int64_t N, int32_t M, int32_t J, double thr_kp,
                               double* p3d, void* stream) {
  M3D_CHECK_RIG("m3d_triangulate_ls_members");
  if (M < 0 || J < 0) return fail(M3D_ERR_INVALID, "m3d_triangulate_ls_members: bad size");
  if (N == 0 || J == 0) return M3D_OK;
  if (!kp_und || !frame || !members || !p3d) return fail(M3D_ERR_INVALID, "m3d_triangulate_ls_members: NULL buffer");
  k_triangulate_ls_members<<<grid_for(N * J, 256, sms), 256, 0, st>>>(rig->dev_g, kp_und, frame, members, N, M, J,
                                                                      thr_kp, p3d);
  return check_launch("k_triangulate_ls_members");
}

// ---- host pipelines ------------------------------------------------------------------------
static int ensure_workspace(m3d_rig* rig, int64_t chunk, bool ransac) {
  const int C = rig->dev.n_cams > 0 ? rig->dev.n_cams : 1;
  if (rig->ws_chunk >= chunk && rig->ws_cams == C && (rig->ws_ransac || !ransac)) return M3D_OK;
  free_workspace(rig);
  for (int i = 0; i < m3d_rig::kSlots; ++i) {
    M3D_CUDA(cudaStreamCreateWithFlags(&rig->ws_stream[i], cudaStreamNonBlocking));
    M3D_CUDA(cudaMalloc(&rig->ws_xy[i], sizeof(double) * 2 * C * chunk));
    M3D_CUDA(cudaMalloc(&rig->ws_xy32[i], sizeof(float) * 2 * C * chunk));
    M3D_CUDA(cudaMalloc(&rig->ws_p3d[i], sizeof(double) * 3 * chunk));
    M3D_CUDA(cudaMalloc(&rig->ws_err[i], sizeof(double) * chunk));
    if (ransac) {
      M3D_CUDA(cudaMalloc(&rig->ws_xyp[i], sizeof(double) * 2 * C * chunk));
      M3D_CUDA(cudaMalloc(&rig->ws_xyp32[i], sizeof(float) * 2 * C * chunk));
      M3D_CUDA(cudaMalloc(&rig->ws_picked[i], (size_t)C * chunk));
      M3D_CUDA(cudaMalloc(&rig->ws_subset[i], sizeof(int32_t) * chunk));
      M3D_CUDA(cudaMalloc(&rig->ws_neval[i], sizeof(int32_t) * chunk));
    }
  }
  rig->ws_chunk = chunk;
  rig->ws_cams = C;
  rig->ws_ransac = ransac;
  return M3D_OK;
}

static const int64_t kHostChunk = 1 << 19;  // joint-instances per pipeline stage

// float32 <-> float64 conversion of observation planes (the *_host_f32 entry points)
__global__ void __launch_bounds__(256) k_widen(const float2* __restrict__ in, double2* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 q = in[i];
    out[i] = make_double2((double)q.x, (double)q.y);
  }
}
__global__ void __launch_bounds__(256) k_narrow(const double2* __restrict__ in, float2* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 q = in[i];
    out[i] = make_float2((float)q.x, (float)q.y);
  }
}

// xy / xy_picked: float64 host arrays, or (xy32 / xy_picked32) their float32 forms
// The arrays are laid out for `ld` points; this call processes points [first, first + N) of them (the
// *_span entry points: several devices / threads share one set of host arrays).
static int host_pipeline(m3d_rig* rig, const double* xy, const float* xy32, int64_t N, int undistort, bool ransac,
                         int min_cams, double threshold, double init_best, double* p3d,
                         uint8_t* picked, double* xy_picked, float* xy_picked32, double* err, int32_t* subset,
                         int32_t* neval, int64_t ld = -1, int64_t first = 0) {
  if (N == 0) return M3D_OK;
  if (ld < 0) ld = N;
  if (xy) xy += 2 * first;
  if (xy32) xy32 += 2 * first;
  p3d += 3 * first;
  if (err) err += first;
  if (picked) picked += first;
  if (xy_picked) xy_picked += 2 * first;
  if (xy_picked32) xy_picked32 += 2 * first;
  if (subset) subset += first;
  if (neval) neval += first;
  DeviceGuard guard(rig->device);
  std::lock_guard<std::mutex> lock(rig->ws_mutex);
  const int C = rig->dev.n_cams;
  const int64_t chunk = N < kHostChunk ? N : kHostChunk;
  int rc = ensure_workspace(rig, chunk, ransac);
  if (rc) return rc;
  const int sms = sm_count_of(rig->device);
  const bool want_xyp = xy_picked || xy_picked32;
  cudaError_t ce = cudaSuccess;
#define PIPE_CUDA(expr)                          \
  do {                                           \
    ce = (expr);                                 \
    if (ce != cudaSuccess) goto pipe_failed;     \
  } while (0)
  {
    int64_t done = 0;
    for (int it = 0; done < N; ++it) {
      const int slot = it % m3d_rig::kSlots;
      cudaStream_t st = rig->ws_stream[slot];
      const int64_t n = (N - done) < chunk ? (N - done) : chunk;
      // the slot's previous D2H copies must have left its buffers (same stream => ordered)
      if (C > 0) {
        if (xy32) {
          PIPE_CUDA(cudaMemcpy2DAsync(rig->ws_xy32[slot], sizeof(float) * 2 * n, xy32 + 2 * done,
                                      sizeof(float) * 2 * ld, sizeof(float) * 2 * n, C, cudaMemcpyHostToDevice, st));
          k_widen<<<grid_for((int64_t)C * n, 256, sms), 256, 0, st>>>(
              reinterpret_cast<const float2*>(rig->ws_xy32[slot]), reinterpret_cast<double2*>(rig->ws_xy[slot]),
              (int64_t)C * n);
          rc = check_launch("k_widen");
          if (rc) goto pipe_failed;
        } else {
          PIPE_CUDA(cudaMemcpy2DAsync(rig->ws_xy[slot], sizeof(double) * 2 * n, xy + 2 * done,
                                      sizeof(double) * 2 * ld, sizeof(double) * 2 * n, C, cudaMemcpyHostToDevice, st));
        }
      }
      if (!ransac) {
        rc = launch_triangulate(rig, rig->ws_xy[slot], n, undistort, rig->ws_p3d[slot],
                                err ? rig->ws_err[slot] : nullptr, st, sms);
      } else {
        rc = m3d_launch_ransac(rig, rig->ws_xy[slot], n, undistort, min_cams, threshold, init_best,
                               rig->ws_p3d[slot], picked ? rig->ws_picked[slot] : nullptr,
                               want_xyp ? rig->ws_xyp[slot] : nullptr, rig->ws_err[slot],
                               subset ? rig->ws_subset[slot] : nullptr, neval ? rig->ws_neval[slot] : nullptr, st);
      }
      if (rc) goto pipe_failed;
      PIPE_CUDA(cudaMemcpyAsync(p3d + 3 * done, rig->ws_p3d[slot], sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, st));
      if (err)
        PIPE_CUDA(cudaMemcpyAsync(err + done, rig->ws_err[slot], sizeof(double) * n, cudaMemcpyDeviceToHost, st));
      if (ransac) {
        if (picked && C > 0)
          PIPE_CUDA(cudaMemcpy2DAsync(picked + done, (size_t)ld, rig->ws_picked[slot], (size_t)n, (size_t)n, C,
                                      cudaMemcpyDeviceToHost, st));
        if (xy_picked && C > 0)
          PIPE_CUDA(cudaMemcpy2DAsync(xy_picked + 2 * done, sizeof(double) * 2 * ld, rig->ws_xyp[slot],
                                      sizeof(double) * 2 * n, sizeof(double) * 2 * n, C, cudaMemcpyDeviceToHost, st));
        if (xy_picked32 && C > 0) {
          k_narrow<<<grid_for((int64_t)C * n, 256, sms), 256, 0, st>>>(
              reinterpret_cast<const double2*>(rig->ws_xyp[slot]), reinterpret_cast<float2*>(rig->ws_xyp32[slot]),
              (int64_t)C * n);
          rc = check_launch("k_narrow");
          if (rc) goto pipe_failed;
          PIPE_CUDA(cudaMemcpy2DAsync(xy_picked32 + 2 * done, sizeof(float) * 2 * ld, rig->ws_xyp32[slot],
                                      sizeof(float) * 2 * n, sizeof(float) * 2 * n, C, cudaMemcpyDeviceToHost, st));
        }
        if (subset)
          PIPE_CUDA(cudaMemcpyAsync(subset + done, rig->ws_subset[slot], sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
        if (neval)
          PIPE_CUDA(cudaMemcpyAsync(neval + done, rig->ws_neval[slot], sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
      }
      done += n;
    }
  }
  for (int i = 0; i < m3d_rig::kSlots; ++i) PIPE_CUDA(cudaStreamSynchronize(rig->ws_stream[i]));
  return M3D_OK;
pipe_failed:
  // earlier chunks may still be copying into the caller's buffers: drain every slot before the
  // caller is told (and frees them)
  for (int i = 0; i < m3d_rig::kSlots; ++i) cudaStreamSynchronize(rig->ws_stream[i]);
#undef PIPE_CUDA
  if (rc) return rc;
  return fail(M3D_ERR_CUDA, std::string("host pipeline: ") + cudaGetErrorString(ce));
}

int m3d_triangulate_error_host(const m3d_rig* rig, const double* xy, int64_t N, int32_t undistort,
                               double* p3d, double* err) {
  if (!rig) return fail(M3D_ERR_INVALID, "m3d_triangulate_error_host: rig is NULL");
  if (N < 0) return fail(M3D_ERR_INVALID, "m3d_triangulate_error_host: negative point count");
  if (N > 0 && (!p3d || (!xy && rig->dev.n_cams > 0)))
    return fail(M3D_ERR_INVALID, "m3d_triangulate_error_host: NULL buffer");
  return host_pipeline(const_cast<m3d_rig*>(rig), xy, nullptr, N, undistort, false, 0, 0, 0, p3d, nullptr,
                       nullptr, nullptr, err, nullptr, nullptr);
}

int m3d_triangulate_error_host_f32(const m3d_rig* rig, const float* xy, int64_t N, int32_t undistort,
                                   double* p3d, double* err) {
  if (!rig) return fail(M3D_ERR_INVALID, "m3d_triangulate_error_host_f32: rig is NULL");
  if (N < 0) return fail(M3D_ERR_INVALID, "m3d_triangulate_error_host_f32: negative point count");
  if (N > 0 && (!p3d || (!xy && rig->dev.n_cams > 0)))
    return fail(M3D_ERR_INVALID, "m3d_triangulate_error_host_f32: NULL buffer");
  return host_pipeline(const_cast<m3d_rig*>(rig), nullptr, xy, N, undistort, false, 0, 0, 0, p3d, nullptr,
                       nullptr, nullptr, err, nullptr, nullptr);
}

int m3d_triangulate_ransac_host(const m3d_rig* rig, const double* xy, int64_t N, int32_t undistort,
                                int32_t min_cams, double threshold, double init_best, double* p3d,
                                uint8_t* picked, double* xy_picked, double* err, int32_t* subset,
                                int32_t* neval) {
  if (!rig) return fail(M3D_ERR_INVALID, "m3d_triangulate_ransac_host: rig is NULL");
  if (N < 0) return fail(M3D_ERR_INVALID, "m3d_triangulate_ransac_host: negative point count");
  if (N > 0 && (!p3d || !err || (!xy && rig->dev.n_cams > 0)))
    return fail(M3D_ERR_INVALID, "m3d_triangulate_ransac_host: NULL buffer");
  return host_pipeline(const_cast<m3d_rig*>(rig), xy, nullptr, N, undistort, true, min_cams, threshold,
                       init_best, p3d, picked, xy_picked, nullptr, err, subset, neval);
}

int m3d_triangulate_ransac_host_f32(const m3d_rig* rig, const float* xy, int64_t N, int32_t undistort,
                                    int32_t min_cams, double threshold, double init_best, double* p3d,
                                    uint8_t* picked, float* xy_picked, double* err, int32_t* subset,
                                    int32_t* neval) {
  if (!rig) return fail(M3D_ERR_INVALID, "m3d_triangulate_ransac_host_f32: rig is NULL");
  if (N < 0) return fail(M3D_ERR_INVALID, "m3d_triangulate_ransac_host_f32: negative point count");
  if (N > 0 && (!p3d || !err || (!xy && rig->dev.n_cams > 0)))
    return fail(M3D_ERR_INVALID, "m3d_triangulate_ransac_host_f32: NULL buffer");
  return host_pipeline(const_cast<m3d_rig*>(rig), nullptr, xy, N, undistort, true, min_cams, threshold,
                       init_best, p3d, picked, nullptr, xy_picked, err, subset, neval);
}

int m3d_triangulate_error_host_span(const m3d_rig* rig, const double* xy, int64_t N_total, int64_t first,
                                    int64_t count, int32_t undistort, double* p3d, double* err) {
  if (!rig) return fail(M3D_ERR_INVALID, "m3d_triangulate_error_host_span: rig is NULL");
  if (first < 0 || count < 0 || first + count > N_total)
    return fail(M3D_ERR_INVALID, "m3d_triangulate_error_host_span: span outside the arrays");
  if (count > 0 && (!p3d || (!xy && rig->dev.n_cams > 0)))
    return fail(M3D_ERR_INVALID, "m3d_triangulate_error_host_span: NULL buffer");
  return host_pipeline(const_cast<m3d_rig*>(rig), xy, nullptr, count, undistort, false, 0, 0, 0, p3d, nullptr,
                       nullptr, nullptr, err, nullptr, nullptr, N_total, first);
}

int m3d_triangulate_ransac_host_span(const m3d_rig* rig, const double* xy, int64_t N_total, int64_t first,
                                     int64_t count, int32_t undistort, int32_t min_cams, double threshold,
                                     double init_best, double* p3d, uint8_t* picked, double* xy_picked, double* err,
                                     int32_t* subset, int32_t* neval) {
  if (!rig) return fail(M3D_ERR_INVALID, "m3d_triangulate_ransac_host_span: rig is NULL");
  if (first < 0 || count < 0 || first + count > N_total)
    return fail(M3D_ERR_INVALID, "m3d_triangulate_ransac_host_span: span outside the arrays");
  if (count > 0 && (!p3d || !err || (!xy && rig->dev.n_cams > 0)))
    return fail(M3D_ERR_INVALID, "m3d_triangulate_ransac_host_span: NULL buffer");
  return host_pipeline(const_cast<m3d_rig*>(rig), xy, nullptr, count, undistort, true, min_cams, threshold,
                       init_best, p3d, picked, xy_picked, nullptr, err, subset, neval, N_total, first);
}

int m3d_host_register(void* ptr, int64_t bytes) {
  if (!ptr || bytes <= 0) return fail(M3D_ERR_INVALID, "m3d_host_register: bad buffer");
  M3D_CUDA(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
  return M3D_OK;
}

int m3d_host_unregister(void* ptr) {
  if (!ptr) return fail(M3D_ERR_INVALID, "m3d_host_unregister: NULL");
  M3D_CUDA(cudaHostUnregister(ptr));
  return M3D_OK;
}

int m3d_probe_fp64_tflops(int32_t device, double* tflops_out) {
  if (!tflops_out) return fail(M3D_ERR_INVALID, "m3d_probe_fp64_tflops: NULL");
  if (m3d_device_count() <= 0) return fail(M3D_ERR_NO_GPU, "m3d_probe_fp64_tflops: no CUDA device");
  DeviceGuard guard(device);
  const int sms = sm_count_of(device);
  const int blocks = sms * 8, threads = 256, iters = 1 << 14;
  double* buf = nullptr;
  M3D_CUDA(cudaMalloc(&buf, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  M3D_CUDA(cudaEventCreate(&e0));
  M3D_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    M3D_CUDA(cudaEventRecord(e0));
    k_dfma_probe<<<blocks, threads>>>(buf, iters);
    g_launches.fetch_add(1);
    M3D_CUDA(cudaEventRecord(e1));
    M3D_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    M3D_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
  *tflops_out = flops / (best * 1e-3) / 1e12;
  return M3D_OK;
}

}  // extern "C"
