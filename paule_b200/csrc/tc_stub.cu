// Placeholder for the tcgen05 path while it is being brought up.
#include "common.cuh"
extern "C" size_t paule_tc_packed_lstm_bytes(int64_t, int64_t) { return 0; }
extern "C" int paule_tc_pack_lstm(const float*, const float*, void*, int64_t, int64_t, paule_stream_t) { return PAULE_ERR_UNSUPPORTED; }
extern "C" int paule_tc_gemm_nt(const float*, const float*, const float*, float*, int64_t, int64_t, int64_t, int, int, paule_stream_t) { return PAULE_ERR_UNSUPPORTED; }
extern "C" size_t paule_tc_rnn_xchg_bytes(int64_t) { return 0; }
extern "C" int paule_tc_lstm_seq_fwd(float*, const void*, float*, float*, void*, int64_t, int64_t, int, paule_stream_t) { return PAULE_ERR_UNSUPPORTED; }
extern "C" int paule_tc_lstm_seq_bwd(float*, const float*, const void*, const float*, int, const float*, void*, int64_t, int64_t, int, paule_stream_t) { return PAULE_ERR_UNSUPPORTED; }
