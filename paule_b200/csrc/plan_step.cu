// The fused inner planning step: everything between optimizer.zero_grad() and the clamp of
// /root/reference/paule/paule.py:911-1211 (minus the logging / VocalTractLab block), for B words at once.
//
//   cp [T,B,30] --ForwardModel LSTM--> h_f --pool+Linear--> pred_mel [Tm,B,60] --EmbeddingModel 2xLSTM--> h_1
//      --Linear at the last frame--> pred_semvec [B,300] --criterion--> per-word loss terms
//   and back: dsv -> BPTT(l1) -> dX GEMM -> BPTT(l0) -> dX GEMM (+dmel of the mel loss) -> post_linear^T, un-pool
//      -> BPTT(forward model) -> dX GEMM -> d(cp);  + smoothness gradients;  Adam + clamp (+ smiling, past_cp)
//
// math = FP32 : every GEMM / recurrence on the FFMA kernels (parity anchor).
// math = BF16 : the three recurrences run as persistent tcgen05 kernels, the K = 720 input projection of embedder
//               layer 1 and all dX GEMMs (K = 2880) as tcgen05 GEMMs fed by the bf16 images the recurrent kernels
//               leave behind; only the skinny K = 30 / 60 / 300 projections stay on the FFMA kernel (they would be
//               pure HBM traffic on tensor cores, and keeping cp / mel in fp32 there costs nothing).
//
// All launches go to one stream; there is no host synchronisation, no allocation and no host-visible state,
// so the whole step can be captured in a CUDA graph and replayed.
#include <stdlib.h>

#include "common.cuh"
#include "plan_internal.cuh"
#include "tc_common.cuh"
#include "tc_lstm.cuh"

using namespace paule;

namespace {

struct Workspace {
  float *gates_f, *h_f, *c_f;
  float *gates_0, *h_0, *c_0;
  float *gates_1, *h_1, *c_1;
  float *dmel, *dsv, *dh1_last, *dh0, *dhp, *dcp_lstm, *dcp_smooth, *dc, *partial;
  void *xchg, *h_img, *hf_img, *da_img;   // tensor-core path only; the image buffers must have been zero-filled once
  void *x_img_f, *x_img_0;                // operand blocks of the fused input projections (cps, mel)
  // layer wavefront (forward model || pooled post_linear || embedder layer 0 as one pipeline): a second exchange buffer for
  // the co-resident recurrence and the per-step arrival counters
  void* xchg2;
  unsigned int *wf_img_flags, *wf_x_flags, *wf_target;
  size_t floats;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

Workspace carve(void* base, int64_t B, int64_t T, int64_t H, int64_t C, int64_t Cm, int64_t S, int math) {
  Workspace w{};
  const int64_t Tm = T / 2;
  size_t off = 0;
  auto take = [&](int64_t n) -> float* {
    float* p = base ? reinterpret_cast<float*>(base) + off : nullptr;
    off += align_up((size_t)n, 64);  // 256-byte granularity keeps every buffer vector / bulk-copy aligned
    return p;
  };
  auto take_bytes = [&](size_t nbytes) -> void* { return nbytes ? (void*)take((int64_t)((nbytes + 3) / 4)) : nullptr; };
  w.gates_f = take(T * B * 4 * H); w.h_f = take(T * B * H); w.c_f = take(T * B * H);
  w.gates_0 = take(Tm * B * 4 * H); w.h_0 = take(Tm * B * H); w.c_0 = take(Tm * B * H);
  w.gates_1 = take(Tm * B * 4 * H); w.h_1 = take(Tm * B * H); w.c_1 = take(Tm * B * H);
  w.dmel = take(Tm * B * Cm); w.dsv = take(B * S); w.dh1_last = take(B * H);
  w.dh0 = take(Tm * B * H); w.dhp = take(Tm * B * H);
  w.dcp_lstm = take(T * B * C); w.dcp_smooth = take(T * B * C);
  w.dc = take(B * H);
  w.partial = take((int64_t)paule_plan_loss_scratch_floats(T, B));
  if (math != PAULE_MATH_FP32) {
    w.xchg = take_bytes(paule_tc_rnn_xchg_bytes(B));
    w.h_img = take_bytes(paule_tc_img_seq_bytes(Tm, B, 1));    // h_0 of every mel frame: A operand of Xp1 = h_0 W_ih1^T
    w.hf_img = take_bytes(paule_tc_img_seq_bytes(T, B, 1));    // forward model's h_t: A operand of the pooled post_linear
    w.da_img = take_bytes(paule_tc_img_seq_bytes(T, B, 4));    // dA of the layer being back-propagated (reused by all three)
    w.x_img_f = take_bytes(paule_tc_x_image_bytes(T, B));      // cps as hi/lo bf16 operand blocks
    w.x_img_0 = take_bytes(paule_tc_x_image_bytes(Tm, B));     // predicted mel as bf16 operand blocks
    const int64_t n_groups = (B + 63) / 64;
    w.xchg2 = take_bytes(paule_tc_rnn_xchg_bytes(B));
    w.wf_img_flags = reinterpret_cast<unsigned int*>(take(n_groups * T));
    w.wf_x_flags = reinterpret_cast<unsigned int*>(take(n_groups * ((Tm + 1) / 2)));
    w.wf_target = reinterpret_cast<unsigned int*>(take(n_groups));
  }
  w.floats = off;
  return w;
}

int check_plan(const paule_plan* p) {
  PAULE_REQUIRE(p != nullptr);
  PAULE_REQUIRE(p->B > 0 && p->T >= 13 && p->H > 0 && p->C > 0 && p->Cm > 0 && p->S > 0);
  PAULE_REQUIRE(p->objective >= 0 && p->objective <= 2);
  PAULE_REQUIRE(p->math == PAULE_MATH_FP32 || p->math == PAULE_MATH_BF16);
  PAULE_REQUIRE(p->fwd.w_ih && p->fwd.w_hh && p->fwd.w_ih_t && p->fwd.w_hh_t && p->fwd.bias);
  PAULE_REQUIRE(p->emb0.w_ih && p->emb0.w_hh && p->emb0.w_ih_t && p->emb0.w_hh_t && p->emb0.bias);
  PAULE_REQUIRE(p->emb1.w_ih && p->emb1.w_hh && p->emb1.w_ih_t && p->emb1.w_hh_t && p->emb1.bias);
  PAULE_REQUIRE(p->fwd.input_size == p->C && p->emb0.input_size == p->Cm && p->emb1.input_size == p->H);
  PAULE_REQUIRE(p->post_w && p->post_w_t && p->post_b && p->head_w && p->head_w_t && p->head_b);
  PAULE_REQUIRE(p->cp && p->target_mel && p->pred_mel && p->pred_sv && p->workspace);
  PAULE_REQUIRE(reinterpret_cast<uintptr_t>(p->workspace) % 256 == 0);
  PAULE_REQUIRE(p->workspace_bytes >= paule_plan_workspace_bytes(p->B, p->T, p->H, p->C, p->Cm, p->S, p->math));
  if (p->math != PAULE_MATH_FP32) {
    if (p->H != 720) return PAULE_ERR_UNSUPPORTED;
    PAULE_REQUIRE(p->fwd.packed && p->emb0.packed && p->emb1.packed);
    PAULE_REQUIRE(p->emb1.packed_ih && p->fwd.packed_ih_t && p->emb0.packed_ih_t && p->emb1.packed_ih_t);
  }
  return PAULE_OK;
}

inline bool tc(const paule_plan* p) { return p->math != PAULE_MATH_FP32; }

// recurrence of one layer on time-major data; gates already holds x W_ih^T + b
int recur_forward(const paule_plan* p, const paule_lstm_layer& L, int64_t steps, float* gates, float* h, float* c,
                  const Workspace& w, void* h_img, paule_stream_t s) {
  if (tc(p)) return paule_tc_lstm_seq_fwd(gates, L.packed, h, c, w.xchg, h_img, steps, p->B, p->math, s);
  return paule_lstm_seq_fwd_f32(gates, L.w_hh, h, c, steps, p->B, p->H, s);
}

// BPTT of one layer followed by dX = dA W_ih (dA [M,4H] x w_ih_t [I,4H]^T)
int layer_backward(const paule_plan* p, const paule_lstm_layer& L, float* gates, const float* c, const float* dh_seq,
                   int dh_mode, const float* dh_last, int64_t steps, float* dx, int accumulate, const Workspace& w,
                   paule_stream_t s) {
  const int64_t B = p->B, H = p->H, I = L.input_size;
  if (tc(p)) {
    // only the bf16 dA images are consumed (dX GEMM on tcgen05): the fp32 copy over the stash is not written
    PAULE_TRY(paule_tc_lstm_seq_bwd_img(gates, c, L.packed, dh_seq, dh_mode, dh_last, w.xchg, w.da_img, steps, B, p->math, s));
    return tc::gemm_img(w.da_img, L.packed_ih_t, nullptr, dx, steps, B, I, 4, accumulate,
                        reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(w.xchg) + tc::kXchgErrOff), as_stream(s));
  }
  PAULE_TRY(paule_lstm_seq_bwd_f32(gates, c, L.w_hh_t, dh_seq, dh_mode, dh_last, w.dc, steps, B, H, s));
  return paule_linear_f32(gates, L.w_ih_t, nullptr, dx, steps * B, I, 4 * H, 1, 4 * H, 0, 0, 1, I, 0, accumulate, s);
}

// Input projection + recurrence of a layer fed by a narrow input (the cps: 30, the mel: 60 channels).  Tensor-core path:
// the projection runs INSIDE the recurrent kernel (paule_tc_lstm_seq_fwd_x), the [steps,B,2880] pre-activations never
// touch HBM; PAULE_NO_FUSED_X=1 keeps the separate FFMA projection (A/B timing, bisecting).
int project_and_recur(const paule_plan* p, const paule_lstm_layer& L, const float* x, int64_t steps, float* gates, float* h,
                      float* c, const Workspace& w, void* x_img, void* h_img, paule_stream_t s) {
  const int64_t B = p->B, H = p->H, I = L.input_size;
  static const bool no_fuse = getenv("PAULE_NO_FUSED_X") != nullptr;
  if (tc(p) && I <= 64 && !no_fuse) {
    PAULE_TRY(paule_tc_x_image(x, x_img, steps, B, I, s));
    // nobody reads this layer's fp32 h when it leaves bf16 images (post_linear / the next layer's projection run on them)
    return paule_tc_lstm_seq_fwd_x(gates, L.packed, L.bias, x_img, h_img ? nullptr : h, c, w.xchg, h_img, steps, B, p->math, s);
  }
  PAULE_TRY(paule_linear_f32(x, L.w_ih, L.bias, gates, steps * B, 4 * H, I, 1, I, 0, 0, 1, 4 * H, 0, 0, s));
  return recur_forward(p, L, steps, gates, h, c, w, h_img, s);
}

// ---- layer wavefront of the forward pass ------------------------------------------------------------------
// Three kernels run CONCURRENTLY on three streams (fork / join with events, capturable in a CUDA graph):
//   stream s   forward-model recurrence (T steps); every epilogue warp release-increments wf_img_flags[group][t] once its
//              h_t image stores are out
//   side 1     pooled post_linear as a STREAMING tcgen05 GEMM: waits for the four h images of a pair of mel frames, writes
//              pred_mel (fp32) and the bf16 operand blocks of embedder layer 0, release-increments wf_x_flags[group][pair]
//   side 2     embedder layer-0 recurrence with the fused input projection: its x-fetching warp waits for wf_x_flags
// The dependency chain is acyclic (s -> side 1 -> side 2), so partial residency of a later kernel can only delay, never
// deadlock; all three fit the device together by construction (CTA budget below).  Critical path: T steps + a short tail
// instead of T + T/2 steps + a GEMM.
struct SideStreams {
  cudaStream_t s1 = nullptr, s2 = nullptr;
  cudaEvent_t fork = nullptr, join1 = nullptr, join2 = nullptr;
};

int side_streams(SideStreams** out) {
  static SideStreams per_device[64];
  int d = 0;
  PAULE_CUDA(cudaGetDevice(&d));
  PAULE_REQUIRE(d >= 0 && d < 64);
  SideStreams& ss = per_device[d];
  if (ss.s1 == nullptr) {   // first use on this device (BatchPlanner warms up outside graph capture)
    PAULE_CUDA(cudaStreamCreateWithFlags(&ss.s1, cudaStreamNonBlocking));
    PAULE_CUDA(cudaStreamCreateWithFlags(&ss.s2, cudaStreamNonBlocking));
    PAULE_CUDA(cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming));
    PAULE_CUDA(cudaEventCreateWithFlags(&ss.join1, cudaEventDisableTiming));
    PAULE_CUDA(cudaEventCreateWithFlags(&ss.join2, cudaEventDisableTiming));
  }
  *out = &ss;
  return PAULE_OK;
}

// arrivals that complete one step of 64-word group g: 23 unit groups x 8 epilogue warps per non-empty word quarter
__global__ void wave_targets_kernel(unsigned int* target, int B, int n_groups) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n_groups) {
    const int words = min(64, B - 64 * g);
    target[g] = (unsigned int)(tc::kWaveArrivalsPerQuarter * ((words + tc::kWq - 1) / tc::kWq));
  }
}
// the co-resident recurrence used its own exchange buffer: fold its status word into the plan's (sticky) one
__global__ void merge_status_kernel(int* dst, const int* src) {
  const int v = *src;
  if (v != 0) atomicCAS(dst, 0, v);
}

// CTA budget of the wavefront: forward model in its usual layout + streaming GEMM + embedder layer 0 in the smallest layout
// that still fits the remaining SMs.  Returns 0 when the three do not fit together (large batches: no wavefront).
int wavefront_emb0_ctas(const paule_plan* p) {
  static const bool off = getenv("PAULE_NO_WAVEFRONT") != nullptr;
  if (off || !tc(p) || p->post_packed == nullptr || (p->T % 2) != 0 || p->T < 8) return 0;
  const int n_f = tc::fwd2_ctas(p->B, 0);
  if (n_f == 0) return 0;
  // the forward model keeps its latency-optimal layout: fewest quarters per CTA that fit one launch
  const int n_m = tc::gemm_stream_ctas(p->B, p->Cm, 2);
  const int left = sm_count() - n_f - n_m - 2;
  return left > 0 ? tc::fwd2_ctas(p->B, left) : 0;
}

int forward_wavefront(const paule_plan* p, const Workspace& w, int emb0_ctas, paule_stream_t stream) {
  const int64_t B = p->B, T = p->T, Tm = T / 2, Cm = p->Cm;
  const int n_groups = (int)((B + 63) / 64), x_pairs = (int)((Tm + 1) / 2);
  cudaStream_t s = as_stream(stream);
  SideStreams* ss = nullptr;
  PAULE_TRY(side_streams(&ss));
  int* status = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(w.xchg) + tc::kXchgErrOff);
  int* status2 = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(w.xchg2) + tc::kXchgErrOff);
  PAULE_CUDA(cudaMemsetAsync(w.wf_img_flags, 0, sizeof(unsigned int) * (size_t)n_groups * T, s));
  PAULE_CUDA(cudaMemsetAsync(w.wf_x_flags, 0, sizeof(unsigned int) * (size_t)n_groups * x_pairs, s));
  wave_targets_kernel<<<(n_groups + 63) / 64, 64, 0, s>>>(w.wf_target, (int)B, n_groups);
  PAULE_LAUNCH_CHECK("wave_targets_kernel");
  PAULE_TRY(paule_tc_x_image(p->cp, w.x_img_f, T, B, p->C, stream));
  PAULE_CUDA(cudaEventRecord(ss->fork, s));
  PAULE_CUDA(cudaStreamWaitEvent(ss->s1, ss->fork, 0));
  PAULE_CUDA(cudaStreamWaitEvent(ss->s2, ss->fork, 0));
  // (1) forward-model recurrence, T steps, announces every h_t image
  PAULE_TRY(tc::lstm_seq_fwd2x(w.gates_f, p->fwd.packed, p->fwd.bias, w.x_img_f, nullptr, w.c_f, w.xchg, w.hf_img, T, B, s,
                               tc::WaveFlags{w.wf_img_flags, nullptr, 0}, 0));
  // (2) pooled post_linear, streaming: pred_mel + operand blocks of embedder layer 0
  PAULE_TRY(tc::gemm_img_stream(w.hf_img, p->post_packed, p->post_b, p->pred_mel, Tm, B, Cm, 2, w.wf_img_flags, w.wf_target, 2,
                                w.wf_x_flags, w.x_img_0, 2, status, ss->s1));
  // (3) embedder layer 0, fed step by step
  PAULE_TRY(tc::lstm_seq_fwd2x(w.gates_0, p->emb0.packed, p->emb0.bias, w.x_img_0, nullptr, w.c_0, w.xchg2, w.h_img, Tm, B,
                               ss->s2, tc::WaveFlags{nullptr, w.wf_x_flags, x_pairs}, emb0_ctas));
  PAULE_CUDA(cudaEventRecord(ss->join1, ss->s1));
  PAULE_CUDA(cudaEventRecord(ss->join2, ss->s2));
  PAULE_CUDA(cudaStreamWaitEvent(s, ss->join1, 0));
  PAULE_CUDA(cudaStreamWaitEvent(s, ss->join2, 0));
  merge_status_kernel<<<1, 1, 0, s>>>(status, status2);
  PAULE_LAUNCH_CHECK("merge_status_kernel");
  return PAULE_OK;
}

// embedder layer 1 + head on the images / stash layer 0 left behind
int embed_top(const paule_plan* p, const Workspace& w, float* sv, paule_stream_t s) {
  const int64_t B = p->B, H = p->H, Tm = p->T / 2, S = p->S;
  if (tc(p)) {   // the gate GEMM over all time steps: Xp1 = h_0 W_ih1^T + b on tcgen05, A = the images layer 0 left
    PAULE_TRY(tc::gemm_img(w.h_img, p->emb1.packed_ih, p->emb1.bias, w.gates_1, Tm, B, 4 * H, 1, 0,
                           reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(w.xchg) + tc::kXchgErrOff), as_stream(s)));
  } else {
    PAULE_TRY(paule_linear_f32(w.h_0, p->emb1.w_ih, p->emb1.bias, w.gates_1, Tm * B, 4 * H, H, 1, H, 0, 0, 1, 4 * H, 0, 0, s));
  }
  PAULE_TRY(recur_forward(p, p->emb1, Tm, w.gates_1, w.h_1, w.c_1, w, nullptr, s));
  const float* h_last = w.h_1 + (Tm - 1) * B * H;
  if (p->word_frames) {   // ragged: every word's own last mel frame (models.py:442); dh1_last is free until the backward
    PAULE_TRY(gather_last(w.h_1, p->word_frames, w.dh1_last, B, H, s));
    h_last = w.dh1_last;
  }
  return paule_linear_f32(h_last, p->head_w, p->head_b, sv, B, S, H, 1, H, 0, 0, 1, S, 0, 0, s);
}

// EmbeddingModel (models.py:440-448) on a time-major mel [Tm,B,Cm] -> sv [B,S]; lens = Tm for every word (paule.py:922-924)
// or the word's own last frame (ragged batches)
int embed_models(const paule_plan* p, const Workspace& w, const float* mel, float* sv, paule_stream_t s) {
  const int64_t Tm = p->T / 2;
  PAULE_TRY(project_and_recur(p, p->emb0, mel, Tm, w.gates_0, w.h_0, w.c_0, w, w.x_img_0, w.h_img, s));
  return embed_top(p, w, sv, s);
}

int forward_models(const paule_plan* p, const Workspace& w, bool need_semvec, paule_stream_t s) {
  const int64_t B = p->B, T = p->T, H = p->H, Tm = T / 2, Cm = p->Cm;
  if (need_semvec) {
    // batches that leave room for it: forward model, pooled post_linear and embedder layer 0 as ONE pipeline
    const int emb0_ctas = wavefront_emb0_ctas(p);
    if (emb0_ctas > 0) {
      PAULE_TRY(forward_wavefront(p, w, emb0_ctas, s));
      return embed_top(p, w, p->pred_sv, s);
    }
  }
  // ForwardModel (models.py:348-356): K = 30 input projection + recurrence
  const bool tc_post = tc(p) && p->post_packed != nullptr && (T % 2 == 0);
  PAULE_TRY(project_and_recur(p, p->fwd, p->cp, T, w.gates_f, w.h_f, w.c_f, w, w.x_img_f, tc_post ? w.hf_img : nullptr, s));
  // post_linear + AvgPool1d(2,2) (the pool commutes with the Linear)
  if (tc_post) {
    // on tcgen05: frames 2k and 2k+1 are consecutive images = two K segments of row (k, b); weights [0.5 W | 0.5 W]
    PAULE_TRY(tc::gemm_img(w.hf_img, p->post_packed, p->post_b, p->pred_mel, Tm, B, Cm, 2, 0,
                           reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(w.xchg) + tc::kXchgErrOff), as_stream(s)));
  } else {   // pool the pair of frames on load
    PAULE_TRY(paule_linear_f32(w.h_f, p->post_w, p->post_b, p->pred_mel, Tm * B, Cm, H, B, 2 * B * H, H, B * H, 1, Cm,
                               0, 0, s));
  }
  if (!need_semvec) return PAULE_OK;
  return embed_models(p, w, p->pred_mel, p->pred_sv, s);
}

}  // namespace

extern "C" size_t paule_plan_workspace_bytes(int64_t B, int64_t T, int64_t H, int64_t C, int64_t Cm, int64_t S,
                                             int math) {
  if (B <= 0 || T <= 0 || H <= 0) return 0;
  return carve(nullptr, B, T, H, C, Cm, S, math).floats * sizeof(float);
}

// Byte offset, inside the workspace, of the int32 status word of the persistent kernels (0 = ok, 1 / 2 = a watchdog fired:
// an exchange / barrier wait exceeded 4 s, results are invalid).  SIZE_MAX when the configuration has no such word (fp32).
extern "C" size_t paule_plan_status_offset(int64_t B, int64_t T, int64_t H, int64_t C, int64_t Cm, int64_t S, int math) {
  if (B <= 0 || T <= 0 || H <= 0 || math == PAULE_MATH_FP32) return (size_t)-1;
  const Workspace w = carve(reinterpret_cast<void*>(uintptr_t(256)), B, T, H, C, Cm, S, math);
  return (size_t)(reinterpret_cast<uintptr_t>(w.xchg) - 256) + tc::kXchgErrOff;
}

// Byte offset inside the workspace of the model-path gradient d(mel + semvec terms)/d(cp) [T,B,C] fp32 of the last
// paule_plan_step (the BPTT result before the smoothness gradients are added in the Adam kernel).
extern "C" size_t paule_plan_grad_lstm_offset(int64_t B, int64_t T, int64_t H, int64_t C, int64_t Cm, int64_t S, int math) {
  if (B <= 0 || T <= 0 || H <= 0) return (size_t)-1;
  const Workspace w = carve(reinterpret_cast<void*>(uintptr_t(256)), B, T, H, C, Cm, S, math);
  return (size_t)(reinterpret_cast<uintptr_t>(w.dcp_lstm) - 256);
}

// semvec of a time-major mel [Tm,B,Cm] through the plan's embedder (target semvec, paule.py:533-535; produced mels, :1139)
extern "C" int paule_plan_embed(const paule_plan* p, const float* mel, float* sv, paule_stream_t stream) {
  PAULE_TRY(check_plan(p));
  PAULE_REQUIRE(mel && sv);
  const Workspace w = carve(p->workspace, p->B, p->T, p->H, p->C, p->Cm, p->S, p->math);
  return embed_models(p, w, mel, sv, stream);
}

extern "C" int paule_plan_forward(const paule_plan* p, paule_stream_t stream) {
  PAULE_TRY(check_plan(p));
  const Workspace w = carve(p->workspace, p->B, p->T, p->H, p->C, p->Cm, p->S, p->math);
  return forward_models(p, w, true, stream);
}

extern "C" int paule_plan_step(const paule_plan* p, paule_stream_t s) {
  PAULE_TRY(check_plan(p));
  PAULE_REQUIRE(p->adam_m && p->adam_v && p->step_count && p->loss_log && p->log_slot_count > 0);
  PAULE_REQUIRE(p->objective == PAULE_OBJ_ACOUSTIC || p->target_sv);
  const int64_t B = p->B, T = p->T, H = p->H, Tm = T / 2, C = p->C, Cm = p->Cm, S = p->S;
  const Workspace w = carve(p->workspace, B, T, H, C, Cm, S, p->math);
  const bool use_sem = p->objective != PAULE_OBJ_ACOUSTIC;
  const bool need_sv = use_sem || (p->log_semantics && p->target_sv);

  PAULE_TRY(paule_step_tick(p->step_count, s));                                   // optimizer step counter
  PAULE_TRY(forward_models(p, w, need_sv, s));                                    // paule.py:913, :924
  PAULE_TRY(plan_loss_logged(p->pred_mel, p->target_mel, need_sv ? p->pred_sv : nullptr,
                             need_sv ? p->target_sv : nullptr, p->cp, p->loss_log, p->step_count, p->log_slot_count,
                             w.dmel, w.dsv, w.dcp_smooth, w.partial, T, Tm, p->word_frames, B, C, Cm, S, p->objective,
                             s, p->cls_w, p->cls_b, p->extra_terms, p->aux_log));  // :986
  if (use_sem) {                                                                   // discrepancy.backward(), :1052
    // head: dh1[Tm-1] = dsv W_head
    PAULE_TRY(paule_linear_f32(w.dsv, p->head_w_t, nullptr, w.dh1_last, B, H, S, 1, S, 0, 0, 1, H, 0, 0, s));
    if (p->word_frames) {
      // ragged: the semvec gradient enters layer 1 at every word's own last frame -- a full external-gradient sequence that
      // is zero elsewhere (dhp is free until post_linear^T), so the recurrent kernels stay unaware of word lengths
      PAULE_TRY(scatter_last(w.dh1_last, p->word_frames, w.dhp, Tm, B, H, s));
      PAULE_TRY(layer_backward(p, p->emb1, w.gates_1, w.c_1, w.dhp, 1, nullptr, Tm, w.dh0, 0, w, s));
    } else {
      PAULE_TRY(layer_backward(p, p->emb1, w.gates_1, w.c_1, nullptr, 0, w.dh1_last, Tm, w.dh0, 0, w, s));
    }
    PAULE_TRY(layer_backward(p, p->emb0, w.gates_0, w.c_0, w.dh0, 1, nullptr, Tm, w.dmel, 1, w, s));  // += d(mel loss)/dmel
  }
  // post_linear^T; the un-pooling (x0.5 to both frames of a pair) is folded into the BPTT's dh load
  PAULE_TRY(paule_linear_f32(w.dmel, p->post_w_t, nullptr, w.dhp, Tm * B, H, Cm, 1, Cm, 0, 0, 1, H, 0, 0, s));
  PAULE_TRY(layer_backward(p, p->fwd, w.gates_f, w.c_f, w.dhp, 2, nullptr, T, w.dcp_lstm, 0, w, s));
  // optimizer.step() + clamp + smiling + past_cp (paule.py:1199-1211)
  PAULE_TRY(adam_clamp_logged(p->cp, w.dcp_lstm, w.dcp_smooth, p->adam_m, p->adam_v, p->step_count, p->lr, p->beta1,
                              p->beta2, p->eps, p->clamp, p->smiling, p->past_cp, p->past_T, p->grad_out, T, B, C, s,
                              p->extra_grad));
  return PAULE_OK;
}
