// Internal (non-ABI) entry points shared by plan_kernels.cu and plan_step.cu.
#pragma once
#include "common.cuh"

extern "C" size_t paule_plan_loss_scratch_floats(int64_t T, int64_t B);
extern "C" int paule_step_tick(int32_t* step_count, paule_stream_t stream);

namespace paule {

// paule_plan_loss_f32 with the terms written to a ring of log slots: slot = (*step_count - 1) % slots
// (step_count may be NULL: slot 0).
int plan_loss_logged(const float* mel, const float* tmel, const float* sv, const float* tsv, const float* cp,
                     float* terms, const int32_t* step_count, int slots, float* dmel, float* dsv, float* dcp_smooth,
                     float* scratch, int64_t T, int64_t Tm, const int32_t* word_T, int64_t B, int64_t C, int64_t Cm,
                     int64_t S, int objective, paule_stream_t stream, const float* cls_w = nullptr,
                     const float* cls_b = nullptr, const float* extra_terms = nullptr, float* aux_log = nullptr, int which = 0);
// `which`: 0 both kernels; 1 only the smoothness kernel (cp -> dcp_smooth + per-tile partial sums: independent of the models'
// forward pass, so the planning step runs it beside the forward pipeline); 2 only the per-word loss kernel (needs 1's partials)

// ragged batches (word_T[b] cp frames per word): out[b,:] = seq[word_T[b]/2 - 1, b, :], and its adjoint into a zero-filled
// [Tm,B,H] sequence
int gather_last(const float* seq, const int32_t* word_T, float* out, int64_t B, int64_t H, paule_stream_t stream);
int scatter_last(const float* rows, const int32_t* word_T, float* seq, int64_t Tm, int64_t B, int64_t H, paule_stream_t stream);

// paule_adam_clamp_f32 that can also emit the summed gradient (log_gradients, paule.py:1062-1063).
int adam_clamp_logged(float* cp, const float* g_a, const float* g_b, float* m, float* v, const int32_t* step_count,
                      float lr, float beta1, float beta2, float eps, float clamp, int smiling, const float* past_cp,
                      int64_t past_T, float* grad_out, int64_t T, int64_t B, int64_t C, paule_stream_t stream,
                      const float* g_c = nullptr);

}  // namespace paule
