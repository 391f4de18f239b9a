// Placeholder while the tcgen05 batched GEMM is being brought up.
#include "common.cuh"
extern "C" int paule_tc_gemm_nt(const float*, const float*, const float*, float*, int64_t, int64_t, int64_t, int, int, paule_stream_t) { return PAULE_ERR_UNSUPPORTED; }
