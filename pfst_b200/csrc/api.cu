// Library-level entry points: version, error strings, device check.
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace pfst {

static thread_local char g_last_error[512] = "";

void set_last_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", where,
           cudaGetErrorName(e), cudaGetErrorString(e));
}

}  // namespace pfst

extern "C" {

const char* pfst_version(void) { return "pfst_sm100 0.1.0 (sm_100a)"; }

const char* pfst_error_string(int code) {
  switch (code) {
    case PFST_OK: return "ok";
    case PFST_ERR_INVALID_ARG: return "invalid argument";
    case PFST_ERR_UNSUPPORTED: return "unsupported configuration";
    case PFST_ERR_CUDA: return "CUDA error (see pfst_last_cuda_error)";
    case PFST_ERR_NO_DEVICE: return "no sm_100 (B200) device is current";
    default: return "unknown error code";
  }
}

const char* pfst_last_cuda_error(void) { return pfst::g_last_error; }

int pfst_copy_async(void* dst, const void* src, int64_t bytes, void* stream) {
  if (!dst || !src || bytes < 0) return PFST_ERR_INVALID_ARG;
  if (bytes == 0) return PFST_OK;
  PFST_CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)),
                "pfst_copy_async");
  return PFST_OK;
}

int pfst_device_check(void) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    pfst::set_last_cuda_error(e, "pfst_device_check/cudaGetDevice");
    cudaGetLastError();
    return PFST_ERR_NO_DEVICE;
  }
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) {
    pfst::set_last_cuda_error(e, "pfst_device_check/cudaDeviceGetAttribute");
    cudaGetLastError();
    return PFST_ERR_NO_DEVICE;
  }
  return major == 10 ? PFST_OK : PFST_ERR_NO_DEVICE;
}

}  // extern "C"
