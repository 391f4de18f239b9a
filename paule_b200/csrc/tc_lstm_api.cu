// C-ABI entry points of the persistent tcgen05 recurrent kernels (tc_lstm_fwd2.cu / tc_lstm_bwd2.cu): argument checks,
// operand packing and buffer sizes.  (The first-generation kernels -- W_hh slices as the shared-memory B operand, counters
// + TMA for the exchange, 3.6 / 3.9 us per cell step -- were deleted in round 2; their numbers are kept in DESIGN.md.)
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_lstm.cuh"

using namespace paule;
using namespace paule::tc;

extern "C" size_t paule_tc_packed_lstm_bytes(int64_t H, int64_t I) {
  (void)I;
  if (H != kH) return 0;
  return kPackedBytes;
}

extern "C" int paule_tc_pack_lstm(const float* w_ih, const float* w_hh, void* packed, int64_t H, int64_t I,
                                  paule_stream_t stream) {
  PAULE_REQUIRE(w_hh && packed);
  if (H != kH) return PAULE_ERR_UNSUPPORTED;
  uint8_t* img = reinterpret_cast<uint8_t*>(packed);
  PAULE_REQUIRE(reinterpret_cast<uintptr_t>(img) % 16 == 0);
  return pack_v2(w_ih, w_hh, I, img, as_stream(stream));
}

extern "C" int paule_tc_rnn_pass_plan(int64_t B, int backward, int32_t* nq_out, int32_t* words_out, int cap) {
  if (B <= 0 || cap < 0 || (cap > 0 && (!nq_out || !words_out))) return -1;
  return backward ? paule::tc::bwd2_pass_plan(B, nq_out, words_out, cap) : paule::tc::fwd2_pass_plan(B, nq_out, words_out, cap);
}

extern "C" size_t paule_tc_rnn_xchg_bytes(int64_t B) {
  (void)B;
  return (size_t)kXchgHeader + kLLBytes;   // header (status word, trace words) + the exchange blocks of one launch
}

extern "C" size_t paule_tc_x_image_bytes(int64_t T, int64_t B) {
  if (T <= 0 || B <= 0) return 0;
  return (size_t)T * (size_t)((B + kWq - 1) / kWq) * kXBlockBytes;
}

extern "C" int paule_tc_x_image(const float* x, void* img, int64_t T, int64_t B, int64_t I, paule_stream_t stream) {
  PAULE_REQUIRE(x && img && T >= 0 && B > 0 && I >= 1 && I <= kXK);
  PAULE_REQUIRE(reinterpret_cast<uintptr_t>(img) % 16 == 0);
  return x_image(x, img, T, B, I, as_stream(stream));
}

extern "C" int paule_tc_lstm_seq_fwd_x(float* gates, const void* packed, const float* bias, const void* x_img, float* h,
                                       float* c, void* xchg, void* h_img_seq, int64_t T, int64_t B, int math,
                                       paule_stream_t stream) {
  PAULE_REQUIRE(gates && packed && bias && x_img && c && xchg && T >= 0 && B > 0);
  PAULE_REQUIRE(h != nullptr || h_img_seq != nullptr);   // h may be NULL when only its bf16 images are consumed
  PAULE_REQUIRE(math == PAULE_MATH_BF16);
  if (T == 0) return PAULE_OK;
  PAULE_REQUIRE(reinterpret_cast<uintptr_t>(xchg) % 16 == 0 && reinterpret_cast<uintptr_t>(x_img) % 16 == 0);
  PAULE_REQUIRE(reinterpret_cast<uintptr_t>(h_img_seq) % 16 == 0);
  return lstm_seq_fwd2x(gates, packed, bias, x_img, h, c, xchg, h_img_seq, T, B, as_stream(stream));
}

extern "C" size_t paule_tc_img_seq_bytes(int64_t T, int64_t B, int64_t images_per_step) {
  if (T <= 0 || B <= 0 || images_per_step <= 0) return 0;
  return (size_t)((B + kRows - 1) / kRows) * (size_t)T * (size_t)images_per_step * kXchgImageBytes;
}

extern "C" int paule_tc_lstm_seq_fwd(float* gates, const void* packed, float* h, float* c, void* xchg, void* h_img_seq,
                                     int64_t T, int64_t B, int math, paule_stream_t stream) {
  PAULE_REQUIRE(gates && packed && h && c && xchg && T >= 0 && B > 0);
  PAULE_REQUIRE(math == PAULE_MATH_BF16);
  if (T == 0) return PAULE_OK;
  PAULE_REQUIRE(reinterpret_cast<uintptr_t>(xchg) % 16 == 0);   // 16-byte exchange loads / bulk copies
  PAULE_REQUIRE(reinterpret_cast<uintptr_t>(h_img_seq) % 16 == 0);
  return lstm_seq_fwd2(gates, packed, h, c, xchg, h_img_seq, T, B, as_stream(stream));
}

static int seq_bwd(float* gates, const float* c, const void* packed, const float* dh_seq, int dh_mode, const float* dh_last,
                   void* xchg, void* da_img_seq, int64_t T, int64_t B, int math, int keep_da, paule_stream_t stream) {
  PAULE_REQUIRE(gates && c && packed && xchg && T >= 0 && B > 0);
  PAULE_REQUIRE(dh_mode == 0 || ((dh_mode == 1 || dh_mode == 2) && dh_seq));
  PAULE_REQUIRE(math == PAULE_MATH_BF16);
  if (T == 0) return PAULE_OK;
  PAULE_REQUIRE(reinterpret_cast<uintptr_t>(xchg) % 16 == 0);
  return lstm_seq_bwd2(gates, c, packed, dh_seq, dh_mode, dh_last, xchg, da_img_seq, T, B, keep_da, as_stream(stream));
}

extern "C" int paule_tc_lstm_seq_bwd(float* gates, const float* c, const void* packed, const float* dh_seq,
                                     int dh_mode, const float* dh_last, void* xchg, void* da_img_seq, int64_t T,
                                     int64_t B, int math, paule_stream_t stream) {
  return seq_bwd(gates, c, packed, dh_seq, dh_mode, dh_last, xchg, da_img_seq, T, B, math, 1, stream);
}

// same, but the fp32 d(pre-activation) is NOT written over `gates`: only the bf16 images in da_img_seq (required) are produced
extern "C" int paule_tc_lstm_seq_bwd_img(float* gates, const float* c, const void* packed, const float* dh_seq,
                                         int dh_mode, const float* dh_last, void* xchg, void* da_img_seq, int64_t T,
                                         int64_t B, int math, paule_stream_t stream) {
  PAULE_REQUIRE(da_img_seq != nullptr);
  return seq_bwd(gates, c, packed, dh_seq, dh_mode, dh_last, xchg, da_img_seq, T, B, math, 0, stream);
}
