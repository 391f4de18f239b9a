/* paule_b200.h -- C ABI of the B200-native PAULE planning hot path.
 *
 * Drop-in boundary for the inner loop of paule.paule.Paule.plan_resynth
 * (/root/reference/paule/paule.py:910-1211) and the model forwards it calls
 * (/root/reference/paule/models.py:326-356, :413-448, :177-247).
 *
 * Conventions
 *   - every function returns 0 (PAULE_OK) on success, non-zero = error code (the convention the
 *     reference already uses for its only native library, VocalTractLab: paule/util.py:32-34);
 *     the Python wrapper turns non-zero into RuntimeError(paule_error_string(code)).
 *   - all pointers are DEVICE pointers owned by the caller (PyTorch); the library never allocates,
 *     frees or synchronises; every kernel is enqueued on `stream` (a cudaStream_t).
 *   - fp32 storage everywhere unless a parameter says otherwise; row-major.
 *   - internal sequence layout is TIME-MAJOR: [T, B, C] (step t of all words is contiguous), the
 *     reference's user-facing tensors are batch-first [B, T, C] (models.py:345 batch_first=True);
 *     paule_transpose_btc converts.
 *   - LSTM gate order is torch.nn.LSTM's i, f, g, o; h0 = c0 = 0 (models.py:349, :441).
 */
#ifndef PAULE_B200_H
#define PAULE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* paule_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define PAULE_API __attribute__((visibility("default")))
#else
#define PAULE_API
#endif

enum {
  PAULE_OK = 0,
  PAULE_ERR_ARG = 1,         /* bad shape / null pointer / unsupported flag combination */
  PAULE_ERR_CUDA = 2,        /* a CUDA runtime call failed; see paule_last_cuda_error() */
  PAULE_ERR_UNSUPPORTED = 3, /* shape outside what the tensor-core kernels were built for */
  PAULE_ERR_NO_DEVICE = 4    /* no sm_100 device: there is NO CPU fallback */
};

/* objective of the criterion closure, paule/paule.py:602-776 */
enum { PAULE_OBJ_ACOUSTIC_SEMVEC = 0, PAULE_OBJ_ACOUSTIC = 1, PAULE_OBJ_SEMVEC = 2 };

/* arithmetic of the recurrent / gate GEMMs */
enum {
  PAULE_MATH_FP32 = 0,   /* FFMA, fp32 operands: the parity anchor */
  PAULE_MATH_BF16 = 1,   /* tcgen05 kind::f16 bf16 operands, fp32 accumulate in TMEM, fp32 cell state */
  PAULE_MATH_BF16X3 = 2  /* reserved: hi/lo-split bf16 operands (3 MMAs), ~fp32 accuracy; not implemented */
};

PAULE_API int paule_version(void);
PAULE_API const char* paule_error_string(int code);
PAULE_API const char* paule_last_cuda_error(void);
/* 0 if a compute-capability-10.x device is current, PAULE_ERR_NO_DEVICE otherwise. */
PAULE_API int paule_device_check(void);

/* ---------------------------------------------------------------------------------------------
 * Generic fp32 operators (any shape).  These replace the ATen calls the reference modules make.
 * ------------------------------------------------------------------------------------------- */

/* y = x W^T + b  -- replaces torch.nn.Linear (models.py:346,350; :437,446) and the LSTM input
 * projection x_t W_ih^T + b_ih + b_hh inside torch.nn.LSTM (models.py:345,349,431,441).
 *   C[cmap(r), n] (+)= sum_k A[amap(r), k] * W[n, k] + bias[n]      r < M, n < N
 * A rows / C rows are addressed through a 2-level map so that batch-first <-> time-major
 * conversion, pooling of row pairs and strided selections need no extra pass:
 *   amap(r) = (r / a_inner) * a_outer_stride + (r % a_inner) * a_inner_stride   (element offsets)
 * a_pair_stride != 0 pools two rows on load: A'[r,k] = 0.5*(A[amap(r)+k] + A[amap(r)+a_pair_stride+k])
 * (AvgPool1d(2,2) commutes with the Linear: models.py:350-354).
 * bias may be NULL.  accumulate != 0 adds into C. */
PAULE_API int paule_linear_f32(const float* A, const float* W, const float* bias, float* C,
                     int64_t M, int64_t N, int64_t K,
                     int64_t a_inner, int64_t a_outer_stride, int64_t a_inner_stride, int64_t a_pair_stride,
                     int64_t c_inner, int64_t c_outer_stride, int64_t c_inner_stride,
                     int accumulate, paule_stream_t stream);

/* Weight gradients of the continue-learning step (paule/paule.py:1372-1377: autograd through aten::lstm / aten::linear):
 *   C[m, n] (+)= sum_r A[r, m] B[r, n]     A [R, lda] (d loss / d pre-activation), B [R, ldb] (layer input or h_{t-1}), C [M, N]
 *   out[m]  (+)= sum_r A[r, m]             (bias gradient)
 * accumulate != 0 adds into the output. */
PAULE_API int paule_gemm_tn_f32(const float* A, const float* B, float* C, int64_t R, int64_t M, int64_t N, int64_t lda,
                      int64_t ldb, int accumulate, paule_stream_t stream);
PAULE_API int paule_colsum_f32(const float* A, float* out, int64_t R, int64_t M, int64_t lda, int accumulate,
                     paule_stream_t stream);

/* [B,T,C] <-> [T,B,C]: out[t,b,:] = in[b,t,:] with (B,T) = (n_outer,n_inner) of `in`. */
PAULE_API int paule_transpose_btc(const float* in, float* out, int64_t n_outer, int64_t n_inner, int64_t C,
                        paule_stream_t stream);

/* One LSTM layer over a whole sequence, forward -- replaces torch.nn.LSTM's recurrence
 * (aten::lstm; models.py:349, :441).
 *   gates [T,B,4H]  in : x_t W_ih^T + b_ih + b_hh (from paule_linear_f32)
 *                   out: the ACTIVATED gates sigma(i), sigma(f), tanh(g), sigma(o) (stash for BPTT)
 *   w_hh  [4H,H]    torch layout (weight_hh_l{k})
 *   h, c  [T,B,H]   outputs (c is part of the stash) */
PAULE_API int paule_lstm_seq_fwd_f32(float* gates, const float* w_hh, float* h, float* c,
                           int64_t T, int64_t B, int64_t H, paule_stream_t stream);

/* Reverse-time BPTT of one layer, input gradients only -- replaces autograd through aten::lstm
 * (discrepancy.backward(), paule/paule.py:1052); weight gradients are not needed to plan.
 *   gates [T,B,4H]  in : activated gates from the forward;  out: d loss / d pre-activation
 *   c     [T,B,H]   cell states from the forward
 *   w_hh_t [H,4H]   TRANSPOSE of weight_hh (row j = column j of w_hh)
 *   dh_seq          gradient wrt the layer's output sequence, or NULL:
 *                     dh_mode 1: [T,B,H];  dh_mode 2: [T/2,B,H], frame t receives 0.5*dh_seq[t/2]
 *                     (adjoint of the pair pooling; an odd last frame receives 0)
 *   dh_last [B,H]   extra gradient on the LAST step's output only, or NULL (embedder head, models.py:442)
 *   scratch         >= B*H floats (running dc)
 * d/dx follows as paule_linear_f32(gates, w_ih_t, ...). */
PAULE_API int paule_lstm_seq_bwd_f32(float* gates, const float* c, const float* w_hh_t,
                           const float* dh_seq, int dh_mode, const float* dh_last,
                           float* scratch, int64_t T, int64_t B, int64_t H, paule_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Planning-loop operators
 * ------------------------------------------------------------------------------------------- */

/* The criterion closure (paule/paule.py:647-662, :705-717, :760-773) with per-word reductions and
 * its analytic gradient -- replaces ~40 ATen ops + their autograd (util.py:570-572, :600, :613-614,
 * :634-636; paule.py:75-88).
 *   mel, tmel  [Tm,B,Cm] time-major;  sv, tsv [B,S];  cp [T,B,C] time-major
 *   terms [B,6] = total, mel, semvec, velocity, jerk, local_linear (weighted; logged BEFORE the update, paule.py:988)
 *   dmel [Tm,B,Cm] = d(5*RMSE_mel)/dmel (0 for the semvec objective);  dsv [B,S] likewise (0 for acoustic)
 *   dcp_smooth [T,B,C] = d(80 MSE(vel) + 400 MSE(jerk) + 1e5 MSE(local_linear))/dcp
 *   scratch    >= paule_plan_loss_scratch_floats(T,B) floats (per-tile partial sums, reduced in fixed order)
 * sv/tsv/dsv may be NULL for the acoustic objective.  Requires T >= 13 (three nested valid 5-point
 * stencils, util.py:634-636) and C <= 32. */
PAULE_API size_t paule_plan_loss_scratch_floats(int64_t T, int64_t B);
PAULE_API int paule_plan_loss_f32(const float* mel, const float* tmel, const float* sv, const float* tsv,
                        const float* cp, float* terms, float* dmel, float* dsv, float* dcp_smooth,
                        float* scratch, int64_t T, int64_t Tm, int64_t B, int64_t C, int64_t Cm, int64_t S,
                        int objective, paule_stream_t stream);

/* torch.optim.Adam([cp]).step() + clamp + smiling + past_cp (paule/paule.py:797, :1199-1211;
 * arithmetic of torch/optim/adam.py::_single_tensor_adam, no weight decay / amsgrad).
 *   cp, m, v [T,B,C] updated in place; g_a, g_b gradients summed on load (g_b may be NULL)
 *   step_count: DEVICE int32 holding the 1-based optimiser step (advance it with paule_step_tick before
 *               the call); it lives on the device so that a captured CUDA graph can be replayed
 *   smiling != 0: cp[:,:,4] = -1, cp[:,:,1] = +1 after the clamp
 *   past_cp [past_T,B,C] or NULL: cp[0:past_T] = past_cp after the update */
PAULE_API int paule_step_tick(int32_t* step_count, paule_stream_t stream); /* *step_count += 1 */
PAULE_API int paule_adam_clamp_f32(float* cp, const float* g_a, const float* g_b, float* m, float* v,
                         const int32_t* step_count, float lr, float beta1, float beta2, float eps, float clamp,
                         int smiling, const float* past_cp, int64_t past_T,
                         int64_t T, int64_t B, int64_t C, paule_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Inverse-model stencils (InverseModelMelTimeSmoothResidual, models.py:177-247); batch-first.
 * ------------------------------------------------------------------------------------------- */

/* One MelChannelConv1D block + residual (models.py:152-169, :224-228): per mel channel c an untied
 * 3(mel) x 5(time) zero-padded convolution.  x,y [B,Tm,Cm]; w [Cm,3,5] with w[c][dm][dt] applied to
 * x[c+dm-1, t+dt-2]; bias [Cm].  y = conv(x) + x. */
PAULE_API int paule_melconv_res_f32(const float* x, const float* w, const float* bias, float* y,
                          int64_t B, int64_t Tm, int64_t Cm, paule_stream_t stream);
/* add_vel_and_acc_info (models.py:47-61): x [B,Tm,Cm] -> y [B,Tm,3Cm]. */
PAULE_API int paule_vel_acc_f32(const float* x, float* y, int64_t B, int64_t Tm, int64_t Cm, paule_stream_t stream);
/* double_sequence + 5 TimeConvResBlocks + resid_weighting (models.py:63-81, :131-139, :236-244).
 * x [B,Tm,C] -> y [B,2Tm,C].  res_w [n_blocks,2,C,5], res_b [n_blocks,2,C], mix_w [C,2,5]
 * (mix_w[c][0] on the smoothed, [c][1] on the raw up-sampled signal), mix_b [C].
 * scratch >= 3*B*2Tm*C floats. */
PAULE_API int paule_upsample_smooth_f32(const float* x, const float* res_w, const float* res_b, int n_blocks,
                              const float* mix_w, const float* mix_b, float* y, float* scratch,
                              int64_t B, int64_t Tm, int64_t C, paule_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Tensor-core path (sm_100a tcgen05 / TMEM / TMA), hidden size 720 only.
 * ------------------------------------------------------------------------------------------- */

/* "UMMA image": a bf16 matrix [rows, K] stored as K/64 k-blocks of rows x 128 B with the canonical SWIZZLE_128B
 * K-major layout (see csrc/tc_common.cuh).  The tcgen05 kernels exchange activations through such images so that a
 * consumer fetches a k-block with one TMA bulk copy and feeds it to tcgen05.mma unchanged. */

/* Bytes of the packed W_hh images (forward slices + backward slices) paule_tc_pack_lstm writes for one layer. */
PAULE_API size_t paule_tc_packed_lstm_bytes(int64_t H, int64_t I);
/* Repack one layer's fp32 W_hh into the bf16 UMMA images the persistent kernels keep resident in shared memory.
 * Required whenever the weights change (continue-learning, paule.py:1372-1377).  w_ih is unused (see
 * paule_tc_gemm_pack for the batched GEMM operands). */
PAULE_API int paule_tc_pack_lstm(const float* w_ih, const float* w_hh, void* packed, int64_t H, int64_t I,
                       paule_stream_t stream);

/* Batched gate GEMMs over all time steps on tcgen05 (K5 / K8 of SURVEY 2.2):
 *   C[(t,b), n] (+)= sum_k A[(t,b), k] W[n, k] + bias[n],  C fp32 [steps, B, N] time-major.
 * A = the image sequence a persistent kernel wrote: [ceil(B/64)][steps][nseg*12 k-blocks][64 rows][128 B]
 *     (nseg = 1: h_t, K = 720 padded to 768; nseg = 4: the four gate images of dA_t, K = 4 x 768).
 * W = fp32 [N, nseg*720] row-major, packed once by paule_tc_gemm_pack into bf16 images (N padded to 16). */
PAULE_API size_t paule_tc_gemm_packed_bytes(int64_t N, int64_t nseg);
PAULE_API int paule_tc_gemm_pack(const float* W, void* packed, int64_t N, int64_t nseg, paule_stream_t stream);
PAULE_API int paule_tc_gemm_img(const void* a_img, const void* packed_b, const float* bias, float* C,
                      int64_t steps, int64_t B, int64_t N, int64_t nseg, int accumulate, paule_stream_t stream);
/* Narrow-K variant (K <= 64: one k-block): C [steps,B,N] (+)= X W^T + bias with X [steps,B,K] fp32 converted to bf16 operand
 * images first.  Replaces the FFMA post_linear^T of the backward pass (dh = dmel W_post, K = 60, N = 720: 17.7 GFLOP at 1024
 * words x 200 mel frames) by the tcgen05 GEMM; `img` >= paule_tc_a_image_bytes(steps, B) bytes, ZERO-FILLED once by the owner. */
PAULE_API size_t paule_tc_gemm_packed_bytes_k64(int64_t N);
PAULE_API int paule_tc_gemm_pack_k64(const float* W /* [N,K] */, void* packed, int64_t N, int64_t K, paule_stream_t stream);
PAULE_API size_t paule_tc_a_image_bytes(int64_t steps, int64_t B);
PAULE_API int paule_tc_a_image(const float* x, void* img, int64_t steps, int64_t B, int64_t I, paule_stream_t stream);
PAULE_API int paule_tc_gemm_img_k64(const void* a_img, const void* packed_b, const float* bias, float* C, int64_t steps,
                          int64_t B, int64_t N, int accumulate, paule_stream_t stream);

/* Persistent-RNN forward / backward of one H=720 layer: W_hh slices stay resident in shared memory
 * for the whole sequence, one cooperative launch per layer and 64-word group, grid barrier per time step.
 * Same contract as paule_lstm_seq_fwd_f32 / _bwd_f32 plus:
 *   packed      image from paule_tc_pack_lstm
 *   xchg        >= paule_tc_rnn_xchg_bytes(B) bytes of scratch (exchange blocks + status word), ZERO-FILLED once by the
 *               caller: int32 at byte 2048 is the sticky status (0 ok, non-zero = a watchdog fired, results invalid)
 *   h_img_seq / da_img_seq   NULL, or >= paule_tc_img_seq_bytes(T, B, 1 / 4) bytes that were ZERO-FILLED once by
 *               the owner: the kernel then keeps the bf16 image of every step there (A operand of
 *               paule_tc_gemm_img) instead of ping-ponging inside xchg. */
PAULE_API size_t paule_tc_rnn_xchg_bytes(int64_t B);
PAULE_API size_t paule_tc_img_seq_bytes(int64_t T, int64_t B, int64_t images_per_step);
/* How paule_tc_lstm_seq_fwd* (backward = 0) / paule_tc_lstm_seq_bwd* (backward = 1) cut a batch of B words into launches: pass i
 * covers words_out[i] consecutive words with nq_out[i] word quarters (16 words) per CTA.  One pass up to 384 / 320 words;
 * beyond that the passes minimise the summed per-step time of the layouts (host-side only, no device work).  Returns the
 * number of passes (<= cap entries are written), or -1 on bad arguments. */
PAULE_API int paule_tc_rnn_pass_plan(int64_t B, int backward, int32_t* nq_out, int32_t* words_out, int cap);
PAULE_API int paule_tc_lstm_seq_fwd(float* gates, const void* packed, float* h, float* c, void* xchg, void* h_img_seq,
                          int64_t T, int64_t B, int math, paule_stream_t stream);
/* Fused input projection (layers whose input size is <= 64: the ForwardModel LSTM fed by the cps, embedder layer 0 fed by the
 * mel).  paule_tc_x_image turns x [T,B,I] into bf16 operand blocks (I <= 32: hi/lo split, so x itself is not rounded);
 * paule_tc_lstm_seq_fwd_x computes a_t = W_hh h_{t-1} + W_ih x_t + bias inside the recurrence (the W_ih slice comes from the
 * same paule_tc_pack_lstm image) -- `gates` is OUTPUT only (the activated-gate stash); replaces the x W_ih^T GEMM of
 * torch.nn.LSTM (paule/models.py:349,441) and its [T,B,4H] round trip through HBM.  x_img must be zero-filled once.
 * h may be NULL when h_img_seq is given (the caller only consumes the bf16 images of h). */
PAULE_API size_t paule_tc_x_image_bytes(int64_t T, int64_t B);
PAULE_API int paule_tc_x_image(const float* x, void* img, int64_t T, int64_t B, int64_t I, paule_stream_t stream);
PAULE_API int paule_tc_lstm_seq_fwd_x(float* gates, const void* packed, const float* bias, const void* x_img, float* h,
                          float* c, void* xchg, void* h_img_seq, int64_t T, int64_t B, int math, paule_stream_t stream);
PAULE_API int paule_tc_lstm_seq_bwd(float* gates, const float* c, const void* packed,
                          const float* dh_seq, int dh_mode, const float* dh_last, void* xchg, void* da_img_seq,
                          int64_t T, int64_t B, int math, paule_stream_t stream);

/* paule_tc_lstm_seq_bwd without the fp32 d(pre-activation) written over `gates`: only the bf16 images (da_img_seq, required),
 * which is all the tcgen05 dX GEMM consumes. */
PAULE_API int paule_tc_lstm_seq_bwd_img(float* gates, const float* c, const void* packed,
                          const float* dh_seq, int dh_mode, const float* dh_last, void* xchg, void* da_img_seq,
                          int64_t T, int64_t B, int math, paule_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * The fused inner step: everything between optimizer.zero_grad() and the clamp
 * (paule/paule.py:911-1211 without the logging / VTL block) for B words at once.
 * ------------------------------------------------------------------------------------------- */
typedef struct paule_lstm_layer {
  const float* w_ih;    /* [4H, I]  */
  const float* w_hh;    /* [4H, H]  */
  const float* w_ih_t;  /* [I, 4H]  */
  const float* w_hh_t;  /* [H, 4H]  */
  const float* bias;    /* [4H] = b_ih + b_hh */
  const void* packed;   /* paule_tc_pack_lstm image or NULL (fp32 math) */
  const void* packed_ih;   /* paule_tc_gemm_pack(w_ih, 4H, 1) if input_size == H, else NULL */
  const void* packed_ih_t; /* paule_tc_gemm_pack(w_ih_t, input_size, 4) or NULL */
  int64_t input_size;
} paule_lstm_layer;

typedef struct paule_plan {
  /* shapes */
  int64_t B, T, H;            /* words, cp frames, hidden size; Tm = T/2 */
  int64_t C, Cm, S;           /* 30, 60, 300 */
  int32_t objective, math, smiling, log_slot_count;
  int32_t log_semantics;      /* acoustic objective only: still run the embedder to log the semvec loss (paule.py:952-958) */
  int32_t reserved0;
  /* models */
  paule_lstm_layer fwd;       /* ForwardModel.lstm            */
  const float* post_w;        /* [Cm, H] post_linear.weight   */
  const float* post_w_t;      /* [H, Cm]                      */
  const float* post_b;        /* [Cm]                         */
  const void* post_packed;    /* paule_tc_gemm_pack([0.5 W_post | 0.5 W_post] as [Cm, 2*720], Cm, 2): pooled post_linear
                                 on tcgen05 (the two frames of a pair are two K segments), or NULL */
  paule_lstm_layer emb0, emb1;/* EmbeddingModel.lstm l0, l1   */
  const float* head_w;        /* [S, H] linear_mapping.weight */
  const float* head_w_t;      /* [H, S]                       */
  const float* head_b;        /* [S]                          */
  /* state (time-major) */
  float* cp;                  /* [T,B,C]  planned trajectory, updated in place */
  float* adam_m;              /* [T,B,C] */
  float* adam_v;              /* [T,B,C] */
  int32_t* step_count;        /* device int32 */
  const float* target_mel;    /* [Tm,B,Cm] */
  const float* target_sv;     /* [B,S]     */
  const float* past_cp;       /* [past_T,B,C] or NULL */
  int64_t past_T;
  float lr, beta1, beta2, eps, clamp;
  /* outputs */
  float* loss_log;            /* [log_slot_count, B, 6]; slot = (step-1) % log_slot_count, written before the update */
  float* pred_mel;            /* [Tm,B,Cm] (also the forward's working buffer) */
  float* pred_sv;             /* [B,S] */
  float* grad_out;            /* [T,B,C] total d loss / d cp of this step (log_gradients), or NULL */
  /* workspace: >= paule_plan_workspace_bytes() */
  void* workspace;
  size_t workspace_bytes;
  /* ragged batches (the reference plans one word per call, so a word never sees another word's frames):
     device int32 [B], cp frames of every word, 13 <= word_frames[b] <= T; word b has word_frames[b]/2 mel frames
     (paule/models.py:353 AvgPool1d(2,2)), its semvec is taken at that last frame (models.py:442), every loss term is a
     mean over the word's own frames, and frames beyond it are padding that receives a zero gradient.  NULL: all T. */
  const int32_t* word_frames;
  /* optional loss branches, all NULL = the plain criterion.
     Speech classifier (paule/paule.py:210-225,603-622: LinearClassifier, paule/models.py:887-911): z_b = mean over the word's
     mel frames of (cls_w . pred_mel[t,b,:]) + cls_b[0]; term 0.1 * BCEWithLogits(z_b, 0) = 0.1 softplus(z_b), fused into the
     criterion kernel together with its gradient 0.1 sigmoid(z_b) cls_w / Tm_b on every mel frame. */
  const float* cls_w;         /* [Cm] */
  const float* cls_b;         /* [1]  */
  /* Loss terms evaluated outside the fused step (somatosensory branch, paule/paule.py:227-273,916-931,624-645:
     cp -> tube -> mel / semvec through three more LSTM models on the same kernels): their per-word values, already
     weighted, are added to the logged total and their gradient d(terms)/d(cp) to the step's gradient before Adam. */
  const float* extra_terms;   /* [B,2]: tube_mel_loss, tube_semvec_loss, or NULL */
  const float* extra_grad;    /* [T,B,C] time-major, or NULL */
  float* aux_log;             /* [log_slot_count, B, 3]: speech_classifier, tube_mel, tube_semvec terms, or NULL */
  /* backward layer wavefront (tensor-core math, <= 64 words): paule_tc_gemm_pack of (W_ih0 W_post)^T, i.e. of the [H, 4H] matrix
     post_linear.weight^T x EmbeddingModel.lstm.weight_ih_l0^T, with (N = H, nseg = 4): embedder layer 0's dA goes to
     d/d(pooled forward-model h) in ONE streaming GEMM while both BPTT kernels run.  NULL: serial backward. */
  const void* bwd_fused_packed;
  /* tensor-core math: paule_tc_gemm_pack_k64 of post_linear.weight^T ([H, Cm], N = H, K = Cm <= 64) -- the serial backward then
     computes d/d(pooled forward-model h) = dmel W_post on the tcgen05 GEMM instead of the FFMA kernel.  NULL: FFMA. */
  const void* post_t_packed;
} paule_plan;

PAULE_API size_t paule_plan_workspace_bytes(int64_t B, int64_t T, int64_t H, int64_t C, int64_t Cm, int64_t S, int math);
/* Byte offset inside the workspace of the int32 status word of the persistent kernels: 0 = ok, non-zero = a watchdog fired
 * (an inter-CTA wait exceeded 4 s) and the results are invalid -- the library never hangs the GPU, the caller must check
 * this word before trusting results.  (size_t)-1 when the configuration has no such word (fp32 math). */
PAULE_API size_t paule_plan_status_offset(int64_t B, int64_t T, int64_t H, int64_t C, int64_t Cm, int64_t S, int math);
/* Byte offset inside the workspace of d(mel + semvec terms)/d(cp), [T,B,C] fp32 time-major, as the last paule_plan_step left
 * it: the BPTT result alone (xx_new.grad of paule.py:1052 minus the smoothness terms' gradient, which the Adam kernel adds). */
PAULE_API size_t paule_plan_grad_lstm_offset(int64_t B, int64_t T, int64_t H, int64_t C, int64_t Cm, int64_t S, int math);
/* EmbeddingModel forward on a time-major mel [Tm,B,Cm] -> sv [B,S] (target semvec, paule.py:533-535), in the plan's math. */
PAULE_API int paule_plan_embed(const paule_plan* p, const float* mel, float* sv, paule_stream_t stream);
/* Kernel launches one paule_plan_step issues for this plan (-1: invalid plan); what callers report as their launch count. */
PAULE_API int64_t paule_plan_step_launches(const paule_plan* p);
/* forward only (no_grad predictions: paule.py:822-824, :1460-1464): fills pred_mel, pred_sv. */
PAULE_API int paule_plan_forward(const paule_plan* p, paule_stream_t stream);
/* one full inner step. */
PAULE_API int paule_plan_step(const paule_plan* p, paule_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PAULE_B200_H */
