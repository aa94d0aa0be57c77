// m3d_ransac.cu — K4: camera-subset RANSAC (CameraGroup.triangulate_possible / triangulate_ransac,
// cameras.py:639-743): kernels (m3d_ransac*.cuh, m3d_possible.cuh), their launch logic and the two
// C-ABI entry points of include/m3d.h that expose them on device buffers.
#include <cuda_runtime.h>

#include <cstdlib>
#include <mutex>
#include <string>

#include "../../include/m3d.h"
#include "m3d_handle.h"
#include "m3d_math.cuh"
#include "m3d_point.cuh"

using namespace m3d;
typedef M3dDeviceGuard DeviceGuard;

static int fail(int code, const std::string& msg) { return m3d_fail(code, msg); }
static int check_launch(const char* what) { return m3d_check_launch(what); }
static int sm_count_of(int device) { return m3d_sm_count(device); }
static int grid_for(int64_t N, int threads, int sm_count) { return m3d_grid_for(N, threads, sm_count); }

#include "m3d_ransac.cuh"
#include "m3d_ransac8.cuh"
#include "m3d_ransac16.cuh"
#include "m3d_possible.cuh"
#include "m3d_ransac_cert.cuh"

// joint-instances per internal ransac launch: bounds the scratch (undistorted views + slots,
// 16 C + 64 bytes per instance) to ~0.8 GB at C = 8
static const int64_t kRansacChunk = 1 << 22;

// scratch of one ransac launch, returned to the rig's pool on every exit path
struct PoolScratch {
  cudaStream_t st;
  void* ptr[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int n = 0;
  explicit PoolScratch(cudaStream_t s) : st(s) {}
  cudaError_t alloc(void** out, size_t bytes, cudaMemPool_t pool) {
    cudaError_t e = cudaMallocFromPoolAsync(out, bytes, pool, st);
    if (e == cudaSuccess) ptr[n++] = *out;
    return e;
  }
  ~PoolScratch() {
    for (int i = 0; i < n; ++i) cudaFreeAsync(ptr[i], st);
  }
};

static int ensure_pool(const m3d_rig* rig) {
  m3d_rig* mrig = const_cast<m3d_rig*>(rig);
  std::lock_guard<std::mutex> lock(mrig->pool_mutex);
  if (!mrig->pool) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = rig->device;
    M3D_CUDA(cudaMemPoolCreate(&mrig->pool, &props));
    unsigned long long keep = ~0ull;
    M3D_CUDA(cudaMemPoolSetAttribute(mrig->pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  return M3D_OK;
}

// joint-instances per internal launch of the pruned search: bounds the queue of undecided points
// (304 bytes per record at C = 8) + slots to 6 GB
static const int64_t kCertChunk = 1 << 24;

// Pruned subset search (m3d_ransac_cert.cuh): setup -> persistent search -> emit, per chunk.
static int launch_ransac_cert(const m3d_rig* rig, const double* xy, int64_t N, int undistort, int min_cams,
                              double threshold, double init_best, double* p3d, uint8_t* picked,
                              double* xy_picked, double* err, int32_t* subset, int32_t* neval,
                              cudaStream_t st) {
  const int C = rig->dev.n_cams;
  const int sms = sm_count_of(rig->device);
  int rc = ensure_pool(rig);
  if (rc) return rc;
  const int64_t chunk = N < kCertChunk ? N : kCertChunk;
  const int F = cert_record_fields(C);
  const size_t rec_bytes = (size_t)((chunk + 31) / 32) * 32 * F * 16;
  PoolScratch scratch(st);
  RansacSlot* slots = nullptr;
  double2* rec = nullptr;
  unsigned int* counters = nullptr;  // [0] records queued, [1] records taken, [2] parked, [3] parked taken
  unsigned int* over = nullptr;      // queue indices of the parked searches
  M3D_CUDA(scratch.alloc((void**)&slots, sizeof(RansacSlot) * (size_t)chunk, rig->pool));
  M3D_CUDA(scratch.alloc((void**)&rec, rec_bytes, rig->pool));
  M3D_CUDA(scratch.alloc((void**)&counters, 4 * sizeof(unsigned int), rig->pool));
  M3D_CUDA(scratch.alloc((void**)&over, sizeof(unsigned int) * (size_t)chunk, rig->pool));
  const bool po = (rig->dev.flags & RIG_HAS_NONPINHOLE) == 0;
  // developer switch: CTAs per SM of the persistent search kernel (2 = 255 registers, 3 = 168)
  // developer switch: M3D_CERT_V1=1 runs the round-2a split (full-set solve in the thread-per-point kernel)
  static const bool cert_v1 = [] { const char* e = getenv("M3D_CERT_V1"); return e && atoi(e) != 0; }();
  // developer switch: M3D_CERT_EMIT=1 keeps the result slots + k_ransac_emit pass in the v2 split
  static const bool cert_emit = [] { const char* e = getenv("M3D_CERT_EMIT"); return e && atoi(e) != 0; }();
  static const int setup_ctas = [] { const char* e = getenv("M3D_CERT_SETUP_CTAS"); return e ? atoi(e) : 0; }();
  static const int search_ctas = [] { const char* e = getenv("M3D_CERT_CTAS"); return e ? atoi(e) : 0; }();
  // evaluations a lane of k_cert_search spends on one point before it parks the search for
  // k_cert_overflow (developer switch M3D_CERT_LANE_LIMIT; 0 = never park)
  static const int lane_limit = [] {
    const char* e = getenv("M3D_CERT_LANE_LIMIT");
    const int v = e ? atoi(e) : 32;
    return v > 0 ? v : 0x7fffffff;
  }();
  for (int64_t n0 = 0; n0 < N; n0 += chunk) {
    const int64_t n = (N - n0) < chunk ? (N - n0) : chunk;
    M3D_CUDA(cudaMemsetAsync(counters, 0, 4 * sizeof(unsigned int), st));
    // v2 split: the search kernels write the reference's outputs themselves (no slots, no emit pass)
    CertOutputs outs = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, N, n0};
    if (!cert_v1 && !cert_emit) outs = CertOutputs{p3d, picked, xy_picked, err, subset, neval, N, n0};
    int64_t blocksA = (n + 127) / 128;
    const int64_t cap = (int64_t)sms * 5 * 32;
    if (blocksA > cap) blocksA = cap;
    if (!cert_v1) {
#define CALLP(PO, NC, MB)                                                                                    \
  M3dKernelTimer timer__("k_cert_prep", st);                                                                 \
  k_cert_prep<PO, NC, MB><<<(unsigned)blocksA, 128, 0, st>>>(rig->dev, rig->cert, xy, N, n0, n, undistort,    \
                                                             threshold, init_best, rec, counters)
      if (C == 8) {
        // 4 CTAs / SM (128 registers): the straight-line undistortion and half budgets (M3D_PREP_ILP) keep several
        // cameras' chains in flight per thread and want the registers — measured per 6.8e7 points: 9.39 ms at
        // 4 CTAs, 10.05 at 5 (600 bytes spilled); the branched round-2e form: 9.69 at 5 CTAs (tools/ab_prep.sh)
        if (setup_ctas == 3) { if (po) { CALLP(true, 8, 3); } else { CALLP(false, 8, 3); } }
        else if (setup_ctas == 5) { if (po) { CALLP(true, 8, 5); } else { CALLP(false, 8, 5); } }
        else if (setup_ctas == 6) { if (po) { CALLP(true, 8, 6); } else { CALLP(false, 8, 6); } }
        else { if (po) { CALLP(true, 8, 4); } else { CALLP(false, 8, 4); } }
      } else {
        if (po) { CALLP(true, 0, 2); } else { CALLP(false, 0, 2); }
      }
#undef CALLP
    } else {
#define CALLA(PO, NC, MB)                                                                                    \
    M3dKernelTimer timer__("k_cert_setup", st);                                                                \
    k_cert_setup<PO, NC, MB><<<(unsigned)blocksA, 128, 0, st>>>(rig->dev, rig->cert, xy, N, n0, n, undistort,   \
                                                                min_cams, threshold, init_best, slots, rec, counters)
      if (C == 8) {
        if (setup_ctas == 4) { if (po) { CALLA(true, 8, 4); } else { CALLA(false, 8, 4); } }
        else if (setup_ctas == 2) { if (po) { CALLA(true, 8, 2); } else { CALLA(false, 8, 2); } }
        else { if (po) { CALLA(true, 8, 3); } else { CALLA(false, 8, 3); } }
      } else {
        if (po) { CALLA(true, 0, 2); } else { CALLA(false, 0, 2); }
      }
  #undef CALLA
  }
    rc = check_launch("k_cert_setup");
    if (rc) return rc;
#define CALLB(PO, NC, MB)                                                                                    \
  do {                                                                                                       \
    auto kfn = k_cert_search<PO, NC, MB>;                                                                    \
    M3dKernelTimer timer__("k_cert_search", st);                                                             \
    const size_t smem_s = (size_t)2 * C * 128 * 16;                                                          \
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s);                     \
    int per_sm = 0;                                                                                          \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, 128, smem_s);                                \
    if (per_sm < 1) per_sm = 1;                                                                              \
    int64_t blocks = (int64_t)sms * per_sm;                                                                  \
    const int64_t need = (n + 127) / 128;                                                                    \
    if (blocks > need) blocks = need;                                                                        \
    kfn<<<(unsigned)blocks, 128, smem_s, st>>>(rig->dev, rig->cumb_g, min_cams, threshold, init_best, slots, rec, \
                                          counters, counters + 1, over, counters + 2, lane_limit, outs);     \
  } while (0)
    if (C == 8) {
      // 3 CTAs / SM (149 registers, no spills) is the measured optimum of the branch-free evaluation:
      // 13.6 ms per 6.8e7 points against 14.4 at 4 CTAs (128 registers, 50 bytes spilled)
      if (search_ctas == 2) { if (po) CALLB(true, 8, 2); else CALLB(false, 8, 2); }
      else if (search_ctas == 4) { if (po) CALLB(true, 8, 4); else CALLB(false, 8, 4); }
      else { if (po) CALLB(true, 8, 3); else CALLB(false, 8, 3); }
    } else {
      if (po) CALLB(true, 0, 2); else CALLB(false, 0, 2);
    }
#undef CALLB
    rc = check_launch("k_cert_search");
    if (rc) return rc;
    {
      M3dKernelTimer timer__("k_cert_overflow", st);
      const unsigned ob = (unsigned)(sms * 2);
#define CALLO(PO, NC)                                                                                        \
  k_cert_overflow<PO, NC><<<ob, 128, 0, st>>>(rig->dev, rig->cumb_g, min_cams, threshold, init_best, slots, rec, \
                                              over, counters + 2, counters + 3, outs)
      if (C == 8) {
        if (po) CALLO(true, 8); else CALLO(false, 8);
      } else {
        if (po) CALLO(true, 0); else CALLO(false, 0);
      }
#undef CALLO
    }
    rc = check_launch("k_cert_overflow");
    if (rc) return rc;
    if (!outs.p3d) {
      {
        M3dKernelTimer timer__("k_ransac_emit", st);
        k_ransac_emit<<<grid_for(n, 256, sms), 256, 0, st>>>(C, xy, N, n0, n, slots, p3d, picked, xy_picked, err,
                                                             subset, neval);
      }
      rc = check_launch("k_ransac_emit");
      if (rc) return rc;
    }
  }
  return M3D_OK;
}

int m3d_launch_ransac(const m3d_rig* rig, const double* xy, int64_t N, int undistort, int min_cams,
                         double threshold, double init_best, double* p3d, uint8_t* picked,
                         double* xy_picked, double* err, int32_t* subset, int32_t* neval,
                         cudaStream_t st) {
  const int C = rig->dev.n_cams;
  const int sms = sm_count_of(rig->device);
  // certified rigs of 8 cameras: the pruned one-kernel search (m3d_ransac_cert.cuh).
  // M3D_RANSAC_EXHAUSTIVE=1 (developer switch) forces the exhaustive kernels for A/B checks.
  static const bool force_exhaustive = [] { const char* e = getenv("M3D_RANSAC_EXHAUSTIVE"); return e && atoi(e) != 0; }();
  if (C >= 2 && rig->cert_all && !force_exhaustive && rig->ransac_mode == M3D_RANSAC_AUTO)
    return launch_ransac_cert(rig, xy, N, undistort, min_cams, threshold, init_best, p3d, picked, xy_picked, err,
                              subset, neval, st);
  const int64_t chunk = N < kRansacChunk ? N : kRansacChunk;
  const bool small_rig = C <= 8;  // table-driven search on per-point records (m3d_ransac8.cuh)
  double* U = nullptr;
  RansacSlot* slots = nullptr;
  unsigned long long* counter = nullptr;
  {
    m3d_rig* mrig = const_cast<m3d_rig*>(rig);
    std::lock_guard<std::mutex> lock(mrig->pool_mutex);
    if (!mrig->pool) {
      cudaMemPoolProps props = {};
      props.allocType = cudaMemAllocationTypePinned;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = rig->device;
      M3D_CUDA(cudaMemPoolCreate(&mrig->pool, &props));
      unsigned long long keep = ~0ull;
      M3D_CUDA(cudaMemPoolSetAttribute(mrig->pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
  }
  M3D_CUDA(cudaMallocFromPoolAsync(&U, sizeof(double) * 2 * (size_t)(C > 0 ? C : 1) * chunk, rig->pool, st));
  M3D_CUDA(cudaMallocFromPoolAsync(&slots, sizeof(RansacSlot) * (size_t)chunk, rig->pool, st));
  M3D_CUDA(cudaMallocFromPoolAsync(&counter, sizeof(unsigned long long), rig->pool, st));
  int rc = M3D_OK;
  for (int64_t n0 = 0; n0 < N && rc == M3D_OK; n0 += chunk) {
    const int64_t n = (N - n0) < chunk ? (N - n0) : chunk;
    const int gridA = grid_for(n, 256, sms);
#define CALLA(F, P, NC) \
  k_ransac_full<F, P, NC><<<gridA, 256, 0, st>>>(rig->dev, xy, N, n0, n, undistort, min_cams, threshold, init_best, U, slots)
#define CALL(F, P)                   \
  do {                               \
    if (C == 8) CALLA(F, P, 8);      \
    else CALLA(F, P, 0);             \
  } while (0)
    M3D_DISPATCH_MODEL(rig, CALL);
#undef CALL
#undef CALLA
    rc = check_launch("k_ransac_full");
    if (rc) break;
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) {
      rc = fail(M3D_ERR_CUDA, std::string("cudaMemsetAsync: ") + cudaGetErrorString(e));
      break;
    }
#define CALLC(F, P, MB)                                                                                   \
  do {                                                                                                    \
    auto kfn = k_ransac_search8<F, P, MB>;                                                                \
    const size_t smem8 = ransac8_smem_bytes();                                                            \
    int per_sm = 0;                                                                                       \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, RANSAC_THREADS, smem8);                   \
    if (per_sm < 1) per_sm = 1;                                                                           \
    int64_t blocks = (int64_t)sms * per_sm;                                                               \
    const int64_t need = (n + 32 * RANSAC_WARPS - 1) / (32 * RANSAC_WARPS);                               \
    if (blocks > need) blocks = need;                                                                     \
    kfn<<<(unsigned)blocks, RANSAC_THREADS, smem8, st>>>(rig->dev_g, xy, N, n0, n, min_cams, threshold,   \
                                                         init_best, U, slots, counter);                   \
  } while (0)
#define CALLD(F, P, MB)                                                                                   \
  do {                                                                                                    \
    auto kfn = k_ransac_search16<F, P, MB>;                                                               \
    const size_t smem16 = ransac16_smem_bytes();                                                          \
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16);                  \
    int per_sm = 0;                                                                                       \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, RANSAC_THREADS, smem16);                  \
    if (per_sm < 1) per_sm = 1;                                                                           \
    int64_t blocks = (int64_t)sms * per_sm;                                                               \
    const int64_t need = (n + 32 * RANSAC_WARPS - 1) / (32 * RANSAC_WARPS);                               \
    if (blocks > need) blocks = need;                                                                     \
    kfn<<<(unsigned)blocks, RANSAC_THREADS, smem16, st>>>(rig->dev_g, xy, N, n0, n, min_cams, threshold,  \
                                                          init_best, U, slots, counter);                  \
  } while (0)
    // rigs of at most 8 cameras: k_ransac_search8; 9..16 cameras: k_ransac_search16 (one more table
    // level; 16 warps / SM at 128 registers, 52 KB shared memory per CTA).
    // measured on B200 (cfg 3): 16 / 20 / 24 / 32 warps per SM (128 / 96 / 80 / 64 registers) run at
    // 4.27 / 4.2 / 4.58 / 4.50e8 inst/s — 24 warps is the default, M3D_RANSAC_VARIANT=4|8 the others
    static const int dev_variant = [] { const char* e = getenv("M3D_RANSAC_VARIANT"); return e ? atoi(e) : 0; }();
#define CALL(F, P)                                   \
  do {                                               \
    if (!small_rig && dev_variant == 3) CALLD(F, P, 3); \
    else if (!small_rig) CALLD(F, P, 4);             \
    else if (dev_variant == 4) CALLC(F, P, 4);       \
    else if (dev_variant == 8) CALLC(F, P, 8);       \
    else CALLC(F, P, 6);                             \
  } while (0)
    M3D_DISPATCH_MODEL(rig, CALL);
#undef CALL
#undef CALLC
#undef CALLD
    rc = check_launch("k_ransac_search");
    if (rc) break;
    k_ransac_emit<<<grid_for(n, 256, sms), 256, 0, st>>>(C, xy, N, n0, n, slots, p3d, picked, xy_picked, err,
                                                         subset, neval);
    rc = check_launch("k_ransac_emit");
  }
  cudaFreeAsync(U, st);
  cudaFreeAsync(slots, st);
  cudaFreeAsync(counter, st);
  return rc;
}

#define M3D_CHECK_RIG(name)                                              \
  if (!rig) return fail(M3D_ERR_INVALID, name ": rig is NULL");          \
  if (N < 0) return fail(M3D_ERR_INVALID, name ": negative point count"); \
  DeviceGuard guard__(rig->device);                                      \
  cudaStream_t st = (cudaStream_t)stream;                                \
  const int sms = sm_count_of(rig->device);                              \
  (void)sms;

extern "C" {

int m3d_triangulate_ransac(const m3d_rig* rig, const double* xy, int64_t N, int32_t undistort,
                           int32_t min_cams, double threshold, double init_best, double* p3d,
                           uint8_t* picked, double* xy_picked, double* err, int32_t* subset,
                           int32_t* neval, void* stream) {
  M3D_CHECK_RIG("m3d_triangulate_ransac");
  if (N == 0) return M3D_OK;
  if (!p3d || !err || (!xy && rig->dev.n_cams > 0))
    return fail(M3D_ERR_INVALID, "m3d_triangulate_ransac: NULL buffer");
  return m3d_launch_ransac(rig, xy, N, undistort, min_cams, threshold, init_best, p3d, picked, xy_picked,
                       err, subset, neval, st);
}

int m3d_triangulate_possible(const m3d_rig* rig, const double* xy, int64_t N, int32_t P, int32_t undistort,
                             int32_t min_cams, double threshold, double init_best, double* p3d,
                             uint8_t* picked, double* xy_picked, double* err, int32_t* index,
                             int32_t* neval, void* stream) {
  M3D_CHECK_RIG("m3d_triangulate_possible");
  const int C = rig->dev.n_cams;
  if (P < 1 || C * P > POSS_SLOTS)
    return fail(M3D_ERR_INVALID, "m3d_triangulate_possible: cameras * candidates must be between 1 and 32");
  if (P > 15)  // the chosen candidate of a camera is kept in 4 bits, 15 = "none" (m3d_possible.cuh)
    return fail(M3D_ERR_INVALID, "m3d_triangulate_possible: at most 15 candidates per camera");
  if (N == 0) return M3D_OK;
  if (!p3d || !err || (!xy && C > 0)) return fail(M3D_ERR_INVALID, "m3d_triangulate_possible: NULL buffer");
  const size_t smem = possible_smem_bytes();
  int64_t blocks = (N + POSS_WARPS - 1) / POSS_WARPS;
  const int64_t cap = (int64_t)sms * 8;
  if (blocks > cap) blocks = cap;
#define CALL(F, Pm)                                                                                   \
  do {                                                                                                \
    auto kfn = k_possible<F, Pm>;                                                                     \
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                \
    kfn<<<(unsigned)blocks, POSS_WARPS * 32, smem, st>>>(rig->dev_g, xy, N, P, undistort, min_cams,   \
                                                         threshold, init_best, p3d, picked, xy_picked, \
                                                         err, index, neval);                          \
  } while (0)
  M3D_DISPATCH_MODEL(rig, CALL);
#undef CALL
  return check_launch("k_possible");
}


}  // extern "C"
