// Shared helpers for the paule_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/paule_b200.h"

namespace paule {

extern thread_local char g_last_cuda_error[256];

inline int record_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return PAULE_OK;
  snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s: %s", what, cudaGetErrorString(e));
  return PAULE_ERR_CUDA;
}

#define PAULE_CUDA(call)                                       \
  do {                                                         \
    int _rc = ::paule::record_cuda((call), #call);             \
    if (_rc != PAULE_OK) return _rc;                           \
  } while (0)

#define PAULE_LAUNCH_CHECK(name)                               \
  do {                                                         \
    int _rc = ::paule::record_cuda(cudaGetLastError(), name);  \
    if (_rc != PAULE_OK) return _rc;                           \
  } while (0)

#define PAULE_TRY(expr)                                        \
  do {                                                         \
    int _rc = (expr);                                          \
    if (_rc != PAULE_OK) return _rc;                           \
  } while (0)

#define PAULE_REQUIRE(cond)                                    \
  do {                                                         \
    if (!(cond)) return PAULE_ERR_ARG;                         \
  } while (0)

inline cudaStream_t as_stream(paule_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

__host__ __device__ static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// paule/paule.py:592-597
constexpr float kMelWeight = 5.0f;
constexpr float kVelWeight = 80.0f;
constexpr float kJerkWeight = 400.0f;
constexpr float kSemWeight = 10.0f;
constexpr float kLocalLinearWeight = 100000.0f;
constexpr float kClassifierWeight = 0.1f;   // SPEECH_CLASSIFIER_WEIGHT, paule/paule.py:596

int sm_count();

// true the first time it is called on the current device with this mask (per-kernel cudaFuncSetAttribute calls are
// per device: a process that drives several GPUs must repeat them on each)
inline bool once_per_device(unsigned long long& mask) {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d > 63) return true;
  if ((mask >> d) & 1ull) return false;
  mask |= 1ull << d;
  return true;
}

}  // namespace paule
