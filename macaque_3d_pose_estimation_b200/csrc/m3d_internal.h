// m3d_internal.h — declarations shared by the translation units of libm3d.so.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "../../include/m3d.h"
#include "m3d_math.cuh"

// records the message for m3d_last_error() and returns `code`
int m3d_fail(int code, const std::string& msg);
// counts the launch (m3d_launch_count) and converts cudaGetLastError() into a return code
int m3d_check_launch(const char* what);
// private stream-ordered memory pool of a device for scratch buffers (nullptr on failure)
cudaMemPool_t m3d_scratch_pool(int device);
// device-side camera record / device index of a rig handle
const m3d::RigDev* m3d_rig_dev(const m3d_rig* rig);

// Optional per-kernel timing (m3d_profile_enable): brackets one launch with CUDA events on its own
// stream; m3d_profile_read() sums the elapsed times per kernel name.  A no-op unless enabled.
struct M3dKernelTimer {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaStream_t st;
  const char* name;
  M3dKernelTimer(const char* name, cudaStream_t st);
  ~M3dKernelTimer();
};

struct M3dDeviceGuard {
  int prev = -1;
  bool changed = false;
  explicit M3dDeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) changed = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~M3dDeviceGuard() {
    if (changed) cudaSetDevice(prev);
  }
};
