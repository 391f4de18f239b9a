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
  void *xchg2, *xchg3;
  void *da_img_0, *da_img_1;              // dA images of the embedder layers when the three BPTT kernels run concurrently
  void *dmel_img;                         // d(loss)/d(pred_mel) as bf16 A images of the narrow-K post_linear^T GEMM
  // arrival counters [word group][step or pair of steps] of every hand-over, one contiguous block (zeroed once per step)
  unsigned int *wf_flags;                 // start of the block
  size_t wf_flag_count;
  unsigned int *f_hf, *f_x0, *f_h0, *f_g1, *f_da1, *f_dh0, *f_da0, *f_dhp;
  unsigned int *wf_target, *wf_target_b;  // arrivals that complete one step: forward / backward recurrences
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
    const int64_t G = (B + 63) / 64, P = (Tm + 1) / 2;
    w.xchg2 = take_bytes(paule_tc_rnn_xchg_bytes(B));
    w.xchg3 = take_bytes(paule_tc_rnn_xchg_bytes(B));
    // the backward wavefront is only taken by batches of at most 64 words: the extra dA image buffers are sized for that
    const int64_t Bw = B <= 64 ? B : 0;
    w.da_img_0 = take_bytes(paule_tc_img_seq_bytes(Tm, Bw, 4));
    w.da_img_1 = take_bytes(paule_tc_img_seq_bytes(Tm, Bw, 4));
    w.wf_flag_count = (size_t)(G * (T + 3 * Tm + 4 * P));
    w.wf_flags = reinterpret_cast<unsigned int*>(take((int64_t)w.wf_flag_count));
    if (w.wf_flags) {
      unsigned int* f = w.wf_flags;
      w.f_hf = f;  f += G * T;
      w.f_h0 = f;  f += G * Tm;
      w.f_da1 = f; f += G * Tm;
      w.f_da0 = f; f += G * Tm;
      w.f_x0 = f;  f += G * P;
      w.f_g1 = f;  f += G * P;
      w.f_dh0 = f; f += G * P;
      w.f_dhp = f;
    }
    w.wf_target = reinterpret_cast<unsigned int*>(take(G));
    w.wf_target_b = reinterpret_cast<unsigned int*>(take(G));
    w.dmel_img = Cm <= 64 ? take_bytes(paule_tc_a_image_bytes(Tm, B)) : nullptr;
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

inline bool use_tc(const paule_plan* p) { return p->math != PAULE_MATH_FP32; }
inline bool Cm_small(const paule_plan* p) { return p->Cm <= 64; }

// recurrence of one layer on time-major data; gates already holds x W_ih^T + b
int recur_forward(const paule_plan* p, const paule_lstm_layer& L, int64_t steps, float* gates, float* h, float* c,
                  const Workspace& w, void* h_img, paule_stream_t s) {
  if (use_tc(p)) return paule_tc_lstm_seq_fwd(gates, L.packed, h, c, w.xchg, h_img, steps, p->B, p->math, s);
  return paule_lstm_seq_fwd_f32(gates, L.w_hh, h, c, steps, p->B, p->H, s);
}

// BPTT of one layer followed by dX = dA W_ih (dA [M,4H] x w_ih_t [I,4H]^T)
int layer_backward(const paule_plan* p, const paule_lstm_layer& L, float* gates, const float* c, const float* dh_seq,
                   int dh_mode, const float* dh_last, int64_t steps, float* dx, int accumulate, const Workspace& w,
                   paule_stream_t s) {
  const int64_t B = p->B, H = p->H, I = L.input_size;
  if (use_tc(p)) {
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
  if (use_tc(p) && I <= 64 && !no_fuse) {
    PAULE_TRY(paule_tc_x_image(x, x_img, steps, B, I, s));
    // nobody reads this layer's fp32 h when it leaves bf16 images (post_linear / the next layer's projection run on them)
    return paule_tc_lstm_seq_fwd_x(gates, L.packed, L.bias, x_img, h_img ? nullptr : h, c, w.xchg, h_img, steps, B, p->math, s);
  }
  PAULE_TRY(paule_linear_f32(x, L.w_ih, L.bias, gates, steps * B, 4 * H, I, 1, I, 0, 0, 1, 4 * H, 0, 0, s));
  return recur_forward(p, L, steps, gates, h, c, w, h_img, s);
}

// ---- layer wavefront -----------------------------------------------------------------------------------------
// The layers of a pass run CONCURRENTLY as one pipeline: every kernel on its own stream (event fork / join, capturable in a
// CUDA graph), handing its output to the next stage step by step through release / acquire arrival counters:
//
//   forward   forward-model recurrence  ->  pooled post_linear (streaming tcgen05 GEMM: pred_mel + bf16 operand blocks)
//             ->  embedder layer 0 (fused input projection)  [->  gate GEMM of layer 1 (streaming)  ->  embedder layer 1]
//   backward  [BPTT layer 1  ->  dX GEMM (streaming, reverse time)  ->]  BPTT layer 0  ->  dX . post_linear^T GEMM (streaming,
//             combined weights, accumulates onto the mel-loss part)  ->  BPTT forward model
//
// A recurrent kernel release-increments counter[group][t] from every epilogue warp once its bf16 image stores of step t are out;
// a streaming GEMM polls the counter of the last step its tile reads, writes its tile and release-increments counter[group][pair];
// the consuming recurrence polls that one before it fetches the step's input.  The dependency chain is acyclic, all kernels of a
// pipeline fit the device together by construction (CTA budget below, one CTA per SM by a shared-memory floor: every kernel
// allocates the whole tensor memory), so partial residency can only delay, never deadlock.  The critical path of a pass falls from
// the SUM of the layers' steps (T + T/2 + T/2) to the longest layer (T) plus a short tail.  Bracketed stages join when the batch
// leaves room for them (<= 48 words); batches beyond 64 words (128 forward) run the serial schedule.
constexpr int kSide = 4;
struct SideStreams {
  cudaStream_t s[kSide] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[kSide] = {nullptr, nullptr, nullptr, nullptr};
};

int side_streams(SideStreams** out) {
  static SideStreams per_device[64];
  int d = 0;
  PAULE_CUDA(cudaGetDevice(&d));
  PAULE_REQUIRE(d >= 0 && d < 64);
  SideStreams& ss = per_device[d];
  if (ss.fork == nullptr) {   // first use on this device (BatchPlanner warms up outside graph capture)
    for (int i = 0; i < kSide; ++i) {
      PAULE_CUDA(cudaStreamCreateWithFlags(&ss.s[i], cudaStreamNonBlocking));
      PAULE_CUDA(cudaEventCreateWithFlags(&ss.join[i], cudaEventDisableTiming));
    }
    PAULE_CUDA(cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming));
  }
  *out = &ss;
  return PAULE_OK;
}
int fork_streams(SideStreams* ss, cudaStream_t s, int n) {
  PAULE_CUDA(cudaEventRecord(ss->fork, s));
  for (int i = 0; i < n; ++i) PAULE_CUDA(cudaStreamWaitEvent(ss->s[i], ss->fork, 0));
  return PAULE_OK;
}
int join_streams(SideStreams* ss, cudaStream_t s, int n) {
  for (int i = 0; i < n; ++i) {
    PAULE_CUDA(cudaEventRecord(ss->join[i], ss->s[i]));
    PAULE_CUDA(cudaStreamWaitEvent(s, ss->join[i], 0));
  }
  return PAULE_OK;
}

// arrivals that complete one step of 64-word group g: (CTAs per word quarter) x 8 epilogue warps per non-empty quarter
__global__ void wave_targets_kernel(unsigned int* target_f, unsigned int* target_b, int B, int n_groups) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n_groups) {
    const int words = min(64, B - 64 * g), quarters = (words + tc::kWq - 1) / tc::kWq;
    target_f[g] = (unsigned int)(tc::kWaveArrivalsPerQuarter * quarters);
    target_b[g] = (unsigned int)(tc::kWaveArrivalsPerQuarterBwd * quarters);
  }
}
// co-resident recurrences use their own exchange buffers: fold their status words into the plan's (sticky) one
__global__ void merge_status_kernel(int* dst, const int* a, const int* b) {
  const int va = *a, vb = *b;
  if (va != 0) atomicCAS(dst, 0, va);
  if (vb != 0) atomicCAS(dst, 0, vb);
  if (a[1] != 0 || b[1] != 0) dst[1] = 1;   // the informational clamp word sits right behind the status word
}
inline int* status_of(void* xchg) { return reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(xchg) + tc::kXchgErrOff); }

struct WavePlan {
  int fwd = 0;          // 0 serial, 1 forward model || post_linear || embedder l0, 2 + gate GEMM || embedder l1
  int f_ctas = 0;       // CTA limit of the forward-model recurrence (0: its latency-optimal layout)
  int e0_ctas = 0, e1_ctas = 0;
  int m_par = 2, g1_ct = 1;   // CTAs per column tile of the streaming post_linear, column tiles per CTA of the streaming gate GEMM
  int bwd = 0;          // 0 serial, 2 BPTT l0 || dX GEMM || BPTT forward model (l1 before, alone), 3 all three BPTT kernels
  int b_nq = 0, g_par = 0;   // quarters per CTA of the co-resident BPTT kernels (one common layout), GEMM CTAs per column tile
};

// CTA budget: the longest layer (forward model) keeps its latency-optimal layout, the others take the smallest layout that
// still fits the remaining SMs (clusters of 4: at most 132 co-resident CTAs of the BPTT kernels).
WavePlan plan_wavefront(const paule_plan* p, bool use_sem) {
  static const bool off = getenv("PAULE_NO_WAVEFRONT") != nullptr;
  static const int fwd_max = getenv("PAULE_WAVEFRONT_FWD") ? atoi(getenv("PAULE_WAVEFRONT_FWD")) : 2;
  static const int bwd_max = getenv("PAULE_WAVEFRONT_BWD") ? atoi(getenv("PAULE_WAVEFRONT_BWD")) : 3;
  WavePlan wp;
  if (off || !use_tc(p) || p->post_packed == nullptr || (p->T % 2) != 0 || p->T < 8) return wp;
  const int sm = sm_count() - 2;
  const int64_t B = p->B;
  const int n_f = tc::fwd2_ctas(B, 0);
  if (n_f > 0 && fwd_max >= 1) {
    const int n_m = tc::gemm_stream_ctas(B, p->Cm, 2), n_g1 = tc::gemm_stream_ctas(B, 4 * p->H, 1);
    // (a budget <= 0 must not reach fwd2_ctas: 0 means "no limit" there)
    const int left2 = (sm - n_f - n_m - n_g1) / 2, left1 = sm - n_f - n_m;
    const int e2 = left2 > 0 ? tc::fwd2_ctas(B, left2) : 0;
    if (fwd_max >= 2 && e2 > 0) {
      wp.fwd = 2; wp.e0_ctas = wp.e1_ctas = e2;
    } else {
      const int e1 = left1 > 0 ? tc::fwd2_ctas(B, left1) : 0;
      if (e1 > 0) { wp.fwd = 1; wp.e0_ctas = e1; }
      // 64 words: the forward model's latency layout (92 CTAs) leaves no room for embedder layer 1.  With TWO quarters per
      // CTA (46 CTAs, 3.25 instead of 2.2 us per step) all five kernels fit, and T x 3.25 us beats T x 2.2 + the serial gate
      // GEMM and layer 1 (T/2 x 2.2 us).  PAULE_WAVEFRONT_DENSE=0 keeps the three-kernel pipeline.
      static const bool dense = getenv("PAULE_WAVEFRONT_DENSE") == nullptr || atoi(getenv("PAULE_WAVEFRONT_DENSE")) != 0;
      const int n_f2 = tc::fwd2_ctas(B, n_f - 1);      // the next denser layout
      if (dense && fwd_max >= 2 && n_f2 > 0) {
        const int left = (sm - n_f2 - n_m - n_g1) / 2;
        const int e = left > 0 ? tc::fwd2_ctas(B, left) : 0;
        const float t_three = (float)p->T * 2.2f + 110.f + (float)(p->T / 2) * 2.2f, t_five = (float)p->T * 3.25f + 40.f;
        if (e > 0 && t_five < t_three) { wp.fwd = 2; wp.f_ctas = n_f2; wp.e0_ctas = wp.e1_ctas = e; }
      }
      // (c) all five kernels with the forward model in its LATENCY layout: the gate GEMM on half as many CTAs (two column tiles
      // each: 0.86 MB per pair of steps through one SM, 8.6 us against the 11.6 us an embedder step pair takes), the pooled
      // post_linear on one, the embedder layers at four quarters per CTA -- at 64 words that is 92 + 1 + 23 + 9 + 23 = all 148
      // SMs.  The embedder layers (T/2 steps at 5.79 us) then set the pace instead of the forward model at two quarters per CTA
      // (T steps at 3.25 us).  No margin is needed for progress: kernels are launched in dependency order and none waits for a
      // later one, so a kernel that does not fit starts late instead of dead-locking.  PAULE_WAVEFRONT_FULL=0 disables it.
      static const bool full = getenv("PAULE_WAVEFRONT_FULL") == nullptr || atoi(getenv("PAULE_WAVEFRONT_FULL")) != 0;
      if (full && fwd_max >= 2) {
        const int n_m1 = tc::gemm_stream_ctas(B, p->Cm, 1), n_g1h = tc::gemm_stream_ctas(B, 4 * p->H, 1, 2);
        const int left = (sm_count() - n_f - n_m1 - n_g1h) / 2;
        const int e = left > 0 ? tc::fwd2_ctas(B, left) : 0;
        if (e > 0 && n_g1h < n_g1) {
          const int64_t quarters = (B + tc::kWq - 1) / tc::kWq;
          const int nq_e = (int)((quarters * tc::kFwd2Groups + e - 1) / e);                     // quarters per CTA of the embedder layers
          const float t_full = (float)(p->T / 2) * tc::kFwdStepUs[nq_e < 1 ? 1 : (nq_e > 4 ? 4 : nq_e)] + 40.f;
          const float t_now = wp.fwd == 2 ? (float)p->T * 3.25f + 40.f : (float)p->T * 2.2f + 110.f + (float)(p->T / 2) * 2.2f;
          if (t_full < t_now) { wp.fwd = 2; wp.f_ctas = 0; wp.e0_ctas = wp.e1_ctas = e; wp.m_par = 1; wp.g1_ct = 2; }
        }
      }
    }
  }
  if (use_sem && p->bwd_fused_packed != nullptr && p->word_frames == nullptr && B <= 64 && bwd_max >= 2) {
    // BPTT kernels that run side by side must be ONE instantiation of the kernel template (measured: two different
    // instantiations of the cluster kernel running concurrently fault with "unspecified launch failure", the same one runs
    // fine) -- so the pipeline picks one common number of quarters per CTA, by estimated time of the pass.  us per step of the
    // BPTT kernel by quarters per CTA (tools/rnn_time.py, <= 64 words per group):
    static const float us_step[5] = {0.f, 2.85f, 3.4f, 5.5f, 8.0f};   // 3 / 4 quarters: as measured side by side with two more kernels
    const float serial = (float)(p->T + 2 * (p->T / 2)) * us_step[tc::bwd2_default_nq(B)];
    float best = serial * 0.97f;   // a pipeline has to be worth its fork / join
    for (int mode = min(bwd_max, 3); mode >= 2; --mode) {
      for (int nq = 1; nq <= 4; ++nq) {
        const int n_b = tc::bwd2_ctas(B, nq), kernels = mode;   // mode 3: l1, l0, forward model; mode 2: l0, forward model
        if (n_b == 0 || kernels * n_b > 132) continue;
        int par = 0;
        for (int c = 4; c >= 2 && par == 0; c -= 2)
          if (kernels * n_b + (mode - 1) * tc::gemm_stream_ctas(B, p->H, c) <= sm) par = c;
        if (par == 0) continue;
        const float est = (float)p->T * us_step[nq] + (mode == 2 ? (float)(p->T / 2) * us_step[tc::bwd2_default_nq(B)] : 0.f);
        if (est < best) { best = est; wp.bwd = mode; wp.b_nq = nq; wp.g_par = par; }
      }
    }
  }
  // debugging overrides of the backward pipeline's layout
  static const int force_nq = getenv("PAULE_WAVE_BNQ") ? atoi(getenv("PAULE_WAVE_BNQ")) : 0;
  static const int force_par = getenv("PAULE_WAVE_PAR") ? atoi(getenv("PAULE_WAVE_PAR")) : 0;
  if (wp.bwd != 0 && force_nq > 0) wp.b_nq = force_nq;
  if (wp.bwd != 0 && force_par > 0) wp.g_par = force_par;
  return wp;
}

int zero_wave_flags(const paule_plan* p, const Workspace& w, cudaStream_t s, bool targets = true) {
  const int n_groups = (int)((p->B + 63) / 64);
  PAULE_CUDA(cudaMemsetAsync(w.wf_flags, 0, sizeof(unsigned int) * w.wf_flag_count, s));
  if (!targets) return PAULE_OK;   // the targets only depend on B: the forward pipeline of the same step wrote them
  wave_targets_kernel<<<(n_groups + 63) / 64, 64, 0, s>>>(w.wf_target, w.wf_target_b, (int)p->B, n_groups);
  PAULE_LAUNCH_CHECK("wave_targets_kernel");
  return PAULE_OK;
}

// smooth_early: the planning step's smoothness kernel (velocity / jerk / local-linear terms of the cps and their gradient) does
// not depend on the models' forward pass -- it is launched on the first side stream, in front of the streaming post_linear, and
// runs beside the start of the pipeline
int forward_wavefront(const paule_plan* p, const Workspace& w, const WavePlan& wp, paule_stream_t stream, bool smooth_early = false) {
  const int64_t B = p->B, T = p->T, Tm = T / 2, Cm = p->Cm, H = p->H;
  const int P = (int)((Tm + 1) / 2);
  cudaStream_t s = as_stream(stream);
  SideStreams* ss = nullptr;
  PAULE_TRY(side_streams(&ss));
  const int n_side = wp.fwd == 2 ? 4 : 2;
  PAULE_TRY(zero_wave_flags(p, w, s));
  PAULE_TRY(paule_tc_x_image(p->cp, w.x_img_f, T, B, p->C, stream));
  PAULE_TRY(fork_streams(ss, s, n_side));
  // (1) forward-model recurrence, T steps, announces every h_t image
  PAULE_TRY(tc::lstm_seq_fwd2x(w.gates_f, p->fwd.packed, p->fwd.bias, w.x_img_f, nullptr, w.c_f, w.xchg, w.hf_img, T, B, s,
                               tc::WaveFlags{w.f_hf, nullptr, 0, 0u}, wp.f_ctas));
  if (smooth_early)
    PAULE_TRY(plan_loss_logged(p->pred_mel, p->target_mel, nullptr, nullptr, p->cp, p->loss_log, p->step_count, p->log_slot_count,
                               w.dmel, w.dsv, w.dcp_smooth, w.partial, T, Tm, p->word_frames, B, p->C, Cm, p->S, PAULE_OBJ_ACOUSTIC,
                               reinterpret_cast<paule_stream_t>(ss->s[0]), nullptr, nullptr, nullptr, nullptr, 1));
  // (2) pooled post_linear, streaming: pred_mel + operand blocks of embedder layer 0
  PAULE_TRY(tc::gemm_img_stream(w.hf_img, p->post_packed, p->post_b, p->pred_mel, Tm, B, Cm, 2, w.f_hf, w.wf_target, 2, w.f_x0,
                                w.x_img_0, wp.m_par, status_of(w.xchg), ss->s[0]));
  // (3) embedder layer 0, fed step by step; announces its h images when layer 1 runs in the pipeline too
  PAULE_TRY(tc::lstm_seq_fwd2x(w.gates_0, p->emb0.packed, p->emb0.bias, w.x_img_0, nullptr, w.c_0, w.xchg2, w.h_img, Tm, B,
                               ss->s[1], tc::WaveFlags{wp.fwd == 2 ? w.f_h0 : nullptr, w.f_x0, P, tc::gemm_stream_arrivals(Cm)},
                               wp.e0_ctas));
  if (wp.fwd == 2) {
    // (4) the gate GEMM over all time steps, streaming: Xp1 = h_0 W_ih1^T + b
    PAULE_TRY(tc::gemm_img_stream(w.h_img, p->emb1.packed_ih, p->emb1.bias, w.gates_1, Tm, B, 4 * H, 1, w.f_h0, w.wf_target, 1,
                                  w.f_g1, nullptr, 1, status_of(w.xchg), ss->s[2], 0, 0, wp.g1_ct));
    // (5) embedder layer 1: its cell warps wait for the pre-activations of the step
    PAULE_TRY(tc::lstm_seq_fwd2(w.gates_1, p->emb1.packed, w.h_1, w.c_1, w.xchg3, nullptr, Tm, B, ss->s[3],
                                tc::WaveFlags{nullptr, w.f_g1, P, tc::gemm_stream_arrivals(4 * H)}, wp.e1_ctas));
  }
  PAULE_TRY(join_streams(ss, s, n_side));
  merge_status_kernel<<<1, 1, 0, s>>>(status_of(w.xchg), status_of(w.xchg2), status_of(w.xchg3));
  PAULE_LAUNCH_CHECK("merge_status_kernel");
  return PAULE_OK;
}

// BPTT of the embedder (both layers, or layer 0 only) and of the forward model as one reverse-time pipeline.  dhp first receives
// the mel-loss part d(mel term)/d(pooled h) = dmel W_post; the streaming GEMM adds dA_0 (W_ih0 W_post) step by step.
int backward_wavefront(const paule_plan* p, const Workspace& w, const WavePlan& wp, paule_stream_t stream) {
  const int64_t B = p->B, T = p->T, Tm = T / 2, H = p->H;
  const int P = (int)((Tm + 1) / 2);
  cudaStream_t s = as_stream(stream);
  SideStreams* ss = nullptr;
  PAULE_TRY(side_streams(&ss));
  const unsigned int g_arr = tc::gemm_stream_arrivals(H);
  PAULE_TRY(zero_wave_flags(p, w, s, wp.fwd == 0));   // (the forward pipeline of this step already wrote the targets)
  // mel-loss part of d/d(pooled h): dhp = dmel W_post (the un-pooling, x0.5 to both frames of a pair, is folded into the BPTT's dh
  // load); only the forward model's BPTT reads it
  auto mel_part = [&](cudaStream_t st) {
    return paule_linear_f32(w.dmel, p->post_w_t, nullptr, w.dhp, Tm * B, H, p->Cm, 1, p->Cm, 0, 0, 1, H, 0, 0,
                            reinterpret_cast<paule_stream_t>(st));
  };
  if (wp.bwd == 2) {
    // layer 1 first, alone among the BPTT kernels, in its usual layout (96 of the 148 SMs at 64 words).  Beside it, on the SMs it
    // leaves free: the mel part, and its own dX GEMM in streaming mode (dh0 = dA_1 W_ih1 follows the recurrence step by step and
    // is complete a few microseconds after it instead of a batch GEMM later).  PAULE_WAVE_B1_BATCH=1: the batch GEMM (A/B timing).
    static const bool b1_batch = getenv("PAULE_WAVE_B1_BATCH") != nullptr;
    if (b1_batch) {
      PAULE_TRY(fork_streams(ss, s, 1));
      PAULE_TRY(mel_part(ss->s[0]));
      PAULE_TRY(layer_backward(p, p->emb1, w.gates_1, w.c_1, nullptr, 0, w.dh1_last, Tm, w.dh0, 0, w, stream));
      PAULE_TRY(join_streams(ss, s, 1));
    } else {
      PAULE_TRY(fork_streams(ss, s, 2));
      PAULE_TRY(tc::lstm_seq_bwd2(w.gates_1, w.c_1, p->emb1.packed, nullptr, 0, w.dh1_last, w.xchg, w.da_img_1, Tm, B, 0, s,
                                  tc::WaveFlags{w.f_da1, nullptr, 0, 0u}, 0));
      PAULE_TRY(tc::gemm_img_stream(w.da_img_1, p->emb1.packed_ih_t, nullptr, w.dh0, Tm, B, H, 4, w.f_da1, w.wf_target_b, 1, w.f_dh0,
                                    nullptr, 4, status_of(w.xchg), ss->s[1], 1, 0));
      PAULE_TRY(mel_part(ss->s[0]));
      PAULE_TRY(join_streams(ss, s, 2));
    }
  } else {
    PAULE_TRY(mel_part(s));
  }
  // PAULE_WAVE_SERIALIZE=1 (debugging): the same kernels with the same flags, one after the other on one stream in dependency
  // order -- every wait finds its counter complete
  static const bool serialize = getenv("PAULE_WAVE_SERIALIZE") != nullptr;
  const int n_side = wp.bwd == 3 ? 4 : 2;
  cudaStream_t sBF = s, sB0 = serialize ? s : ss->s[0], sG0 = serialize ? s : ss->s[1], sB1 = serialize ? s : ss->s[2],
               sG1 = serialize ? s : ss->s[3];
  if (!serialize) PAULE_TRY(fork_streams(ss, s, n_side));
  auto run_bf = [&]() {   // forward model (the longest layer), usual layout; waits for dhp pair by pair
    return tc::lstm_seq_bwd2(w.gates_f, w.c_f, p->fwd.packed, w.dhp, 2, nullptr, w.xchg, w.da_img, T, B, 0, sBF,
                             tc::WaveFlags{nullptr, w.f_dhp, P, g_arr}, wp.b_nq);
  };
  auto run_b0 = [&]() {   // embedder layer 0: announces its dA images; in the three-kernel pipeline it waits for dh0 pair by pair
    return tc::lstm_seq_bwd2(w.gates_0, w.c_0, p->emb0.packed, w.dh0, 1, nullptr, w.xchg2, w.da_img_0, Tm, B, 0, sB0,
                             tc::WaveFlags{w.f_da0, wp.bwd == 3 ? w.f_dh0 : nullptr, P, g_arr}, wp.b_nq);
  };
  auto run_g0 = [&]() {   // dhp += dA_0 (W_ih0 W_post): streaming, reverse time, accumulating onto the mel-loss part
    return tc::gemm_img_stream(w.da_img_0, p->bwd_fused_packed, nullptr, w.dhp, Tm, B, H, 4, w.f_da0, w.wf_target_b, 1, w.f_dhp,
                               nullptr, wp.g_par, status_of(w.xchg), sG0, 1, 1);
  };
  auto run_b1 = [&]() {
    return tc::lstm_seq_bwd2(w.gates_1, w.c_1, p->emb1.packed, nullptr, 0, w.dh1_last, w.xchg3, w.da_img_1, Tm, B, 0, sB1,
                             tc::WaveFlags{w.f_da1, nullptr, 0, 0u}, wp.b_nq);
  };
  auto run_g1 = [&]() {   // dh0 = dA_1 W_ih1: streaming, reverse time
    return tc::gemm_img_stream(w.da_img_1, p->emb1.packed_ih_t, nullptr, w.dh0, Tm, B, H, 4, w.f_da1, w.wf_target_b, 1, w.f_dh0,
                               nullptr, wp.g_par, status_of(w.xchg), sG1, 1, 0);
  };
  if (serialize) {
    if (wp.bwd == 3) { PAULE_TRY(run_b1()); PAULE_TRY(run_g1()); }
    PAULE_TRY(run_b0()); PAULE_TRY(run_g0()); PAULE_TRY(run_bf());
  } else {
    // launched in dependency order: a tool that serialises kernels (Nsight Compute's kernel replay) then still finds every
    // counter complete when its waiter starts
    if (wp.bwd == 3) { PAULE_TRY(run_b1()); PAULE_TRY(run_g1()); }
    PAULE_TRY(run_b0()); PAULE_TRY(run_g0()); PAULE_TRY(run_bf());
    PAULE_TRY(join_streams(ss, s, n_side));
  }
  merge_status_kernel<<<1, 1, 0, s>>>(status_of(w.xchg), status_of(w.xchg2), status_of(w.xchg3));
  PAULE_LAUNCH_CHECK("merge_status_kernel");
  // d(cp) = dA_f W_ih (batch GEMM over the images the forward model's BPTT left)
  return tc::gemm_img(w.da_img, p->fwd.packed_ih_t, nullptr, w.dcp_lstm, T, B, p->fwd.input_size, 4, 0, status_of(w.xchg), s);
}

// head of the embedder: semvec from layer 1's output at the last frame (every word's own last frame in ragged batches)
int embed_head(const paule_plan* p, const Workspace& w, float* sv, paule_stream_t s) {
  const int64_t B = p->B, H = p->H, Tm = p->T / 2, S = p->S;
  const float* h_last = w.h_1 + (Tm - 1) * B * H;
  if (p->word_frames) {   // ragged: every word's own last mel frame (models.py:442); dh1_last is free until the backward
    PAULE_TRY(gather_last(w.h_1, p->word_frames, w.dh1_last, B, H, s));
    h_last = w.dh1_last;
  }
  return paule_linear_f32(h_last, p->head_w, p->head_b, sv, B, S, H, 1, H, 0, 0, 1, S, 0, 0, s);
}

// embedder layer 1 + head on the images / stash layer 0 left behind
int embed_top(const paule_plan* p, const Workspace& w, float* sv, paule_stream_t s) {
  const int64_t B = p->B, H = p->H, Tm = p->T / 2;
  if (use_tc(p)) {   // the gate GEMM over all time steps: Xp1 = h_0 W_ih1^T + b on tcgen05, A = the images layer 0 left
    PAULE_TRY(tc::gemm_img(w.h_img, p->emb1.packed_ih, p->emb1.bias, w.gates_1, Tm, B, 4 * H, 1, 0, status_of(w.xchg), as_stream(s)));
  } else {
    PAULE_TRY(paule_linear_f32(w.h_0, p->emb1.w_ih, p->emb1.bias, w.gates_1, Tm * B, 4 * H, H, 1, H, 0, 0, 1, 4 * H, 0, 0, s));
  }
  PAULE_TRY(recur_forward(p, p->emb1, Tm, w.gates_1, w.h_1, w.c_1, w, nullptr, s));
  return embed_head(p, w, sv, s);
}

// EmbeddingModel (models.py:440-448) on a time-major mel [Tm,B,Cm] -> sv [B,S]; lens = Tm for every word (paule.py:922-924)
// or the word's own last frame (ragged batches)
int embed_models(const paule_plan* p, const Workspace& w, const float* mel, float* sv, paule_stream_t s) {
  const int64_t Tm = p->T / 2;
  PAULE_TRY(project_and_recur(p, p->emb0, mel, Tm, w.gates_0, w.h_0, w.c_0, w, w.x_img_0, w.h_img, s));
  return embed_top(p, w, sv, s);
}

// smooth_done (optional, planning step only): set when the smoothness kernel of the criterion was launched beside the forward pass
int forward_models(const paule_plan* p, const Workspace& w, bool need_semvec, paule_stream_t s, bool* smooth_done = nullptr) {
  const int64_t B = p->B, T = p->T, H = p->H, Tm = T / 2, Cm = p->Cm;
  if (need_semvec) {
    // batches that leave room for it: the layers of the forward pass as ONE pipeline
    const WavePlan wp = plan_wavefront(p, false);
    if (wp.fwd == 2) {
      // (not in the every-SM-taken plan: measured 1.82 -> 1.96 ms per step at 64 words -- there is no SM left for a bystander)
      const bool early = smooth_done != nullptr && wp.g1_ct == 1;
      PAULE_TRY(forward_wavefront(p, w, wp, s, early));
      if (early) *smooth_done = true;
      return embed_head(p, w, p->pred_sv, s);
    }
    if (wp.fwd == 1) {
      PAULE_TRY(forward_wavefront(p, w, wp, s, smooth_done != nullptr));
      if (smooth_done) *smooth_done = true;
      return embed_top(p, w, p->pred_sv, s);
    }
  }
  // ForwardModel (models.py:348-356): K = 30 input projection + recurrence
  const bool tc_post = use_tc(p) && p->post_packed != nullptr && (T % 2 == 0);
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

// Kernel launches of the library that one paule_plan_step issues for this plan (memsets and event nodes not counted): the
// same branch conditions as the code below, for callers that report a launch count (bench.py's `gpu_launches`).
extern "C" int64_t paule_plan_step_launches(const paule_plan* p) {
  if (check_plan(p) != PAULE_OK) return -1;
  const int64_t B = p->B, T = p->T, Tm = T / 2;
  const bool use_sem = p->objective != PAULE_OBJ_ACOUSTIC;
  const bool need_sv = use_sem || (p->log_semantics && p->target_sv);
  const int64_t ragged = p->word_frames ? 1 : 0;
  int64_t n = 1 + 2 + 1;   // step tick, criterion (2 kernels), Adam
  if (!use_tc(p)) {
    n += (1 + T) + 1;                                                    // forward model: projection + T steps, post_linear
    if (need_sv) n += 2 * (1 + Tm) + ragged + 1;                         // embedder: 2 x (projection + steps), head
    if (use_sem) n += 1 + ragged + 2 * (Tm + 1);                         // head^T, BPTT + dX of both embedder layers
    n += 1 + (T + 1);                                                    // post_linear^T, BPTT + dX of the forward model
    return n;
  }
  const int64_t fpass = tc::fwd2_passes(B), bpass = tc::bwd2_passes(B);   // launches per recurrent layer
  const WavePlan wf = need_sv ? plan_wavefront(p, false) : WavePlan{};
  if (wf.fwd != 0) {
    n += 1 + 1 + fpass + 1 + 1 + 1;                                      // targets, x image, forward model, GEMM, embedder l0, status
    n += (wf.fwd == 2 ? 2 : 1 + fpass) + ragged + 1;                     // gate GEMM + embedder l1, head
  } else {
    n += 1 + fpass + 1;                                                  // x image, forward model, post_linear
    if (need_sv) n += 1 + fpass + 1 + fpass + ragged + 1;                // x image, l0, gate GEMM, l1, head
  }
  const WavePlan wb = plan_wavefront(p, use_sem);
  if (use_sem && wb.bwd != 0) {
    n += 2 + (wb.fwd == 0 ? 1 : 0) + (wb.bwd == 2 ? bpass + 1 : 2) + 3 + 1 + 1;   // head^T, post^T, (targets), l1 (+ dX), BF, l0, GEMM, status, dX
  } else {
    if (use_sem) n += 1 + ragged + 2 * (bpass + 1);
    n += 1 + bpass + 1;
    if (p->post_t_packed != nullptr && Cm_small(p)) n += 1;               // dmel -> operand images of the narrow-K GEMM
  }
  return n;
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
  bool smooth_done = false;
  PAULE_TRY(forward_models(p, w, need_sv, s, &smooth_done));                      // paule.py:913, :924
  PAULE_TRY(plan_loss_logged(p->pred_mel, p->target_mel, need_sv ? p->pred_sv : nullptr,
                             need_sv ? p->target_sv : nullptr, p->cp, p->loss_log, p->step_count, p->log_slot_count,
                             w.dmel, w.dsv, w.dcp_smooth, w.partial, T, Tm, p->word_frames, B, C, Cm, S, p->objective,
                             s, p->cls_w, p->cls_b, p->extra_terms, p->aux_log, smooth_done ? 2 : 0));  // :986
  const WavePlan wp = plan_wavefront(p, use_sem);
  if (use_sem && wp.bwd != 0) {                                                    // discrepancy.backward(), :1052, as a pipeline
    PAULE_TRY(paule_linear_f32(w.dsv, p->head_w_t, nullptr, w.dh1_last, B, H, S, 1, S, 0, 0, 1, H, 0, 0, s));   // dh1[Tm-1] = dsv W_head
    PAULE_TRY(backward_wavefront(p, w, wp, s));   // (computes the mel-loss part dhp = dmel W_post itself)
    PAULE_TRY(adam_clamp_logged(p->cp, w.dcp_lstm, w.dcp_smooth, p->adam_m, p->adam_v, p->step_count, p->lr, p->beta1,
                                p->beta2, p->eps, p->clamp, p->smiling, p->past_cp, p->past_T, p->grad_out, T, B, C, s,
                                p->extra_grad));
    return PAULE_OK;
  }
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
  if (use_tc(p) && p->post_t_packed != nullptr && w.dmel_img != nullptr) {
    // on tcgen05: dmel -> bf16 operand images (one k-block per step and 64 words), then the narrow-K batch GEMM (K = Cm <= 64)
    // -- the FFMA kernel's 2 Tm B H Cm FLOP were 2 % of the step at 1024 words (0.71 ms of 35.6)
    PAULE_TRY(paule_tc_a_image(w.dmel, w.dmel_img, Tm, B, Cm, s));
    PAULE_TRY(tc::gemm_img_kb(w.dmel_img, p->post_t_packed, nullptr, w.dhp, Tm, B, H, 1, 0, status_of(w.xchg), as_stream(s)));
  } else {
    PAULE_TRY(paule_linear_f32(w.dmel, p->post_w_t, nullptr, w.dhp, Tm * B, H, Cm, 1, Cm, 0, 0, 1, H, 0, 0, s));
  }
  PAULE_TRY(layer_backward(p, p->fwd, w.gates_f, w.c_f, w.dhp, 2, nullptr, T, w.dcp_lstm, 0, w, s));
  // optimizer.step() + clamp + smiling + past_cp (paule.py:1199-1211)
  PAULE_TRY(adam_clamp_logged(p->cp, w.dcp_lstm, w.dcp_smooth, p->adam_m, p->adam_v, p->step_count, p->lr, p->beta1,
                              p->beta2, p->eps, p->clamp, p->smiling, p->past_cp, p->past_T, p->grad_out, T, B, C, s,
                              p->extra_grad));
  return PAULE_OK;
}
