// m3d_peer.cu — result window of the frame-sharded run (no reference counterpart: the
// reference is single-process NumPy).
//
// One rank owns the frame-ordered result arrays of the whole recording (cudaMalloc, exported as a CUDA
// IPC handle); every other rank maps them into its own address space over NVLink peer access and its COPY
// ENGINE writes each finished tile straight to the rows it belongs to.  No SM of either side takes part,
// no rank waits for another per tile (a gather is a rendezvous of all ranks per round), and the receiving
// GPU's NVLink ingress carries nothing but payload.
#include <cstring>

#include "m3d_handle.h"

static_assert(sizeof(cudaIpcMemHandle_t) == M3D_PEER_HANDLE_BYTES, "handle size of include/m3d.h");

// like M3D_CUDA, and clears the runtime's last-error slot: a refused mapping must not surface later as the
// result of an unrelated kernel launch
#define M3D_PEER_CUDA(expr)                                                                  \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      cudaGetLastError();                                                                    \
      return m3d_fail(M3D_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));    \
    }                                                                                        \
  } while (0)

extern "C" {

int m3d_peer_alloc(int32_t device, int64_t bytes, void** dptr_out, uint8_t* handle_out) {
  if (!dptr_out || !handle_out || bytes <= 0) return m3d_fail(M3D_ERR_INVALID, "m3d_peer_alloc: bad arguments");
  if (m3d_device_count() <= 0) return m3d_fail(M3D_ERR_NO_GPU, "m3d_peer_alloc: no CUDA device");
  M3dDeviceGuard guard(device);
  void* p = nullptr;
  M3D_PEER_CUDA(cudaMalloc(&p, (size_t)bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    cudaGetLastError();
    return m3d_fail(M3D_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  }
  std::memcpy(handle_out, &h, sizeof(h));
  *dptr_out = p;
  return M3D_OK;
}

int m3d_peer_free(int32_t device, void* dptr) {
  if (!dptr) return M3D_OK;
  M3dDeviceGuard guard(device);
  M3D_PEER_CUDA(cudaFree(dptr));
  return M3D_OK;
}

int m3d_peer_open(int32_t device, const uint8_t* handle, void** dptr_out) {
  if (!handle || !dptr_out) return m3d_fail(M3D_ERR_INVALID, "m3d_peer_open: bad arguments");
  if (m3d_device_count() <= 0) return m3d_fail(M3D_ERR_NO_GPU, "m3d_peer_open: no CUDA device");
  M3dDeviceGuard guard(device);
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  // maps the exporter's allocation for `device`; peer access to the exporting GPU is enabled on demand
  M3D_PEER_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dptr_out = p;
  return M3D_OK;
}

int m3d_peer_close(int32_t device, void* dptr) {
  if (!dptr) return M3D_OK;
  M3dDeviceGuard guard(device);
  M3D_PEER_CUDA(cudaIpcCloseMemHandle(dptr));
  return M3D_OK;
}

int m3d_peer_push(void* dst_window, const void* src_dev, int64_t bytes, void* stream) {
  if (bytes == 0) return M3D_OK;
  if (!dst_window || !src_dev || bytes < 0) return m3d_fail(M3D_ERR_INVALID, "m3d_peer_push: bad arguments");
  M3D_PEER_CUDA(cudaMemcpyAsync(dst_window, src_dev, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return M3D_OK;
}

}  // extern "C"
