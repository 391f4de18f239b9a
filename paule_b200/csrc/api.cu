// Library-level entry points: version, error strings, device check.
#include "common.cuh"

extern "C" int paule_version(void) { return 100; }  // 0.1.0

extern "C" const char* paule_error_string(int code) {
  switch (code) {
    case PAULE_OK: return "ok";
    case PAULE_ERR_ARG: return "invalid argument (shape, null pointer or flag combination)";
    case PAULE_ERR_CUDA: return "CUDA runtime error (see paule_last_cuda_error)";
    case PAULE_ERR_UNSUPPORTED: return "shape not supported by the tensor-core kernels";
    case PAULE_ERR_NO_DEVICE: return "no sm_100 (B200) device is current; paule_b200 has no CPU fallback";
    default: return "unknown error";
  }
}

extern "C" const char* paule_last_cuda_error(void) { return paule::g_last_cuda_error; }

extern "C" int paule_device_check(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return PAULE_ERR_NO_DEVICE; }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return PAULE_ERR_NO_DEVICE;
  }
  return major == 10 ? PAULE_OK : PAULE_ERR_NO_DEVICE;
}
