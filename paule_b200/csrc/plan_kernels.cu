// Fused loss / analytic-gradient / Adam kernels of the planning loop (HBM-bound, fp32).
//
// Reference arithmetic being replaced (all /root/reference):
//   criterion          paule/paule.py:647-662 (+ :705-717 acoustic, :760-773 semvec), weights :592-597
//   RMSE (eps = 0)     paule/util.py:570-572, instance paule/paule.py:68
//   5-point stencil    paule/util.py:600, nested 1x/2x/3x in get_vel_acc_jerk :634-636
//   local_linear       paule/util.py:613-614
//   Adam + clamp       paule/paule.py:797,1199-1211; torch/optim/adam.py::_single_tensor_adam
//
// Layout: time-major [T,B,C]; a word's frames are B*C floats apart, so every stencil tap of a warp
// is a coalesced row segment.
#include "common.cuh"
#include "plan_internal.cuh"

namespace paule {

constexpr int kTileT = 64;   // output frames per CTA of the smoothness kernel
constexpr int kHalo = 12;    // three nested 5-point stencils reach 6 frames, their adjoints another 6
constexpr int kMaxC = 32;    // channels (30 in PAULE)

// d5(v)[i] = (-v[i+4] + 8 v[i+3] - 8 v[i+1] + v[i]) / 12   -- same operation order as util.py:600, no FMA contraction of the
// sum so that it matches the reference's separately rounded tensor ops.  The division by the constant 12 is a multiply
// by RN(1/12) followed by one FMA residual correction (q + (s - 12 q) r): the correctly rounded quotient except for
// measure-zero corner cases, in 3 instructions -- __fdiv_rn is ~35 and made this kernel instruction-bound (223 M warp
// instructions at 1024 words x 400 frames).
__device__ __forceinline__ float div12(float s) {
  constexpr float r = 1.0f / 12.0f;
  const float q = __fmul_rn(s, r);
  return __fmaf_rn(__fmaf_rn(-12.0f, q, s), r, q);
}
__device__ __forceinline__ float d5(float v0, float v1, float v3, float v4) {
  float s = __fadd_rn(-v4, __fmul_rn(8.0f, v3));
  s = __fsub_rn(s, __fmul_rn(8.0f, v1));
  s = __fadd_rn(s, v0);
  return div12(s);
}

// Smoothness terms of one (word, time tile): vel/jerk/local-linear partial sums and their gradient.
//   x -> vel (valid: T-4) -> acc (T-8) -> jerk (T-12);  r3 = s_j*jerk; r2 = D^T r3; r1 = D^T r2 + s_v*vel; g = D^T r1 + ll part
// Shared-memory tiles hold frames [t0-12, t0+kTileT+12) of the word as [88 frames][32 channel slots].  Thread = (channel =
// lane, run of 11 consecutive frames = warp): every stage pulls its run plus a 4-frame apron into registers (15 shared
// loads for 11 results, conflict-free: a warp reads one 128-byte row per load) and writes 11 results back -- no index
// division, the nested stencils keep the reference's stage-by-stage rounding (d5 of d5 of d5, util.py:634-636).
constexpr int kSmW = kTileT + 2 * kHalo;   // 88 frames
constexpr int kRun = kSmW / 8;             // 11 frames per warp
static_assert(kRun * 8 == kSmW, "the frame window splits evenly over the 8 warps");

__global__ void __launch_bounds__(256)
smooth_terms_kernel(const float* __restrict__ cp, float* __restrict__ dcp_smooth, float* __restrict__ partial,
                    int64_t Tmax, const int32_t* __restrict__ word_T, int64_t B, int64_t C, int n_tiles) {
  __shared__ float sx[kSmW * 32];   // x
  __shared__ float sv[kSmW * 32];   // vel
  __shared__ float s1[kSmW * 32];   // acc, later r2
  __shared__ float s2[kSmW * 32];   // r3, later r1
  __shared__ float sred[3][8];
  const int64_t b = blockIdx.y;
  const int64_t t0 = (int64_t)blockIdx.x * kTileT;
  const int tid = threadIdx.x, c = tid & 31, wq = tid >> 5;
  const int f0 = wq * kRun;                      // first frame of this thread's run inside the window
  const int64_t i0 = t0 - kHalo + f0;            // ... and its global frame index
  const bool cl = c < (int)C;
  // ragged batches: word b has T frames; frames [T, Tmax) are padding and get a zero gradient
  const int64_t T = word_T ? (int64_t)word_T[b] : Tmax;
  const float kv = 2.0f * kVelWeight / (float)((T - 4) * C);
  const float kj = 2.0f * kJerkWeight / (float)((T - 12) * C);
  const float kl = 2.0f * kLocalLinearWeight / (float)((T - 2) * C);
  auto own = [&](int f) { return f >= kHalo && f < kHalo + kTileT; };   // frames this CTA sums over and writes

  // ---- stage 0: x.  Unconditional loads from clamped addresses, all issued before the first use: written as
  // `cond ? load : 0` per frame the loads stay behind their branches and the 11 DRAM round trips serialise (17 us per CTA).
  {
    float xin[kRun];
    const int cc = cl ? c : 0;
#pragma unroll
    for (int u = 0; u < kRun; ++u) {
      int64_t t = i0 + u;
      t = t < 0 ? 0 : (t >= T ? T - 1 : t);
      xin[u] = __ldg(cp + (t * B + b) * C + cc);
    }
#pragma unroll
    for (int u = 0; u < kRun; ++u) {
      const int64_t t = i0 + u;
      sx[(f0 + u) * 32 + c] = (cl && t >= 0 && t < T) ? xin[u] : 0.f;
    }
  }
  __syncthreads();
  // a run plus its forward apron (frames f0 .. f0+14; beyond the window: 0)
  auto load_fwd = [&](const float* sm, float (&r)[kRun + 4]) {
#pragma unroll
    for (int u = 0; u < kRun + 4; ++u) r[u] = (f0 + u < kSmW) ? sm[(f0 + u) * 32 + c] : 0.f;
  };
  // a run plus its backward apron (frames f0-4 .. f0+10; before the window: 0); r[u] = value at frame f0 - 4 + u
  auto load_bwd = [&](const float* sm, float (&r)[kRun + 4]) {
#pragma unroll
    for (int u = 0; u < kRun + 4; ++u) r[u] = (f0 - 4 + u >= 0) ? sm[(f0 - 4 + u) * 32 + c] : 0.f;
  };
  // adjoint of d5: (D^T r)[k] = (r[k] - 8 r[k-1] + 8 r[k-3] - r[k-4]) / 12 with r = 0 outside its range
  auto adj = [](float r0, float r1, float r3, float r4) { return (r0 - 8.0f * r1 + 8.0f * r3 - r4) * (1.0f / 12.0f); };

  float pv = 0.f, pj = 0.f, pl = 0.f;
  float xr[kRun + 4];
  load_fwd(sx, xr);
  // ---- stage 1: vel[i] = d5(x[i..i+4]), i in [0, T-4); local-linear partial sums (util.py:614) on the owned frames
#pragma unroll
  for (int u = 0; u < kRun; ++u) {
    const int64_t i = i0 + u;
    const float v = (i >= 0 && i < T - 4 && f0 + u + 4 < kSmW) ? d5(xr[u], xr[u + 1], xr[u + 3], xr[u + 4]) : 0.f;
    sv[(f0 + u) * 32 + c] = v;
    if (own(f0 + u)) {
      if (i < T - 4) pv += v * v;
      if (i >= 1 && i < T - 1) {   // ll centred at frame i = (2 x[i] - x[i-1] - x[i+1]) / 2
        const float xm = (u > 0) ? xr[u - 1] : sx[(f0 - 1) * 32 + c];   // an owned frame is never the window's first
        const float l = __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(2.0f, xr[u]), xm), xr[u + 1]), 0.5f)   /* / 2 is exact */;
        pl += l * l;
      }
    }
  }
  __syncthreads();
  // ---- stage 2: acc[i] = d5(vel[i..i+4]), i in [0, T-8)
  {
    float r[kRun + 4];
    load_fwd(sv, r);
#pragma unroll
    for (int u = 0; u < kRun; ++u) {
      const int64_t i = i0 + u;
      s1[(f0 + u) * 32 + c] = (i >= 0 && i < T - 8 && f0 + u + 4 < kSmW) ? d5(r[u], r[u + 1], r[u + 3], r[u + 4]) : 0.f;
    }
  }
  __syncthreads();
  // ---- stage 3: jerk[i] = d5(acc[i..i+4]), i in [0, T-12); r3 = kj * jerk
  {
    float r[kRun + 4];
    load_fwd(s1, r);
#pragma unroll
    for (int u = 0; u < kRun; ++u) {
      const int64_t i = i0 + u;
      const float jk = (i >= 0 && i < T - 12 && f0 + u + 4 < kSmW) ? d5(r[u], r[u + 1], r[u + 3], r[u + 4]) : 0.f;
      if (own(f0 + u)) pj += jk * jk;
      s2[(f0 + u) * 32 + c] = kj * jk;
    }
  }
  __syncthreads();
  // ---- stage 4: r2 = D^T r3, k in [0, T-8)
  {
    float r[kRun + 4];
    load_bwd(s2, r);
#pragma unroll
    for (int u = 0; u < kRun; ++u) {
      const int64_t k = i0 + u;
      s1[(f0 + u) * 32 + c] = (k >= 0 && k < T - 8) ? adj(r[u + 4], r[u + 3], r[u + 1], r[u]) : 0.f;
    }
  }
  __syncthreads();
  // ---- stage 5: r1 = D^T r2 + kv * vel, k in [0, T-4)
  {
    float r[kRun + 4];
    load_bwd(s1, r);
#pragma unroll
    for (int u = 0; u < kRun; ++u) {
      const int64_t k = i0 + u;
      s2[(f0 + u) * 32 + c] = (k >= 0 && k < T - 4) ? adj(r[u + 4], r[u + 3], r[u + 1], r[u]) + kv * sv[(f0 + u) * 32 + c] : 0.f;
    }
  }
  __syncthreads();
  // ---- stage 6: g = D^T r1 + local-linear adjoint, owned frames only
  {
    float r[kRun + 4];
    load_bwd(s2, r);
    // ll centred at frame index fi (global centre t): 0 outside [1, T-1)
    auto ll_c = [&](int64_t t, int fi) -> float {
      if (t < 1 || t >= T - 1 || fi < 1 || fi + 1 >= kSmW) return 0.f;
      return __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(2.0f, sx[fi * 32 + c]), sx[(fi - 1) * 32 + c]), sx[(fi + 1) * 32 + c]), 0.5f);
    };
#pragma unroll
    for (int u = 0; u < kRun; ++u) {
      const int f = f0 + u;
      const int64_t t = i0 + u;
      if (!own(f) || t >= Tmax || !cl) continue;
      float g = 0.f;
      if (t < T) {
        // g_ll[t] = kl * (ll[t] - ll[t-1] / 2 - ll[t+1] / 2) with ll indexed by its centre frame
        g = adj(r[u + 4], r[u + 3], r[u + 1], r[u]) + kl * (ll_c(t, f) - 0.5f * ll_c(t - 1, f - 1) - 0.5f * ll_c(t + 1, f + 1));
      }
      dcp_smooth[(t * B + b) * C + c] = g;
    }
  }
  // block reduce the three partial sums (fixed order -> deterministic; idle lanes hold zeros)
  for (int o = 16; o > 0; o >>= 1) {
    pv += __shfl_down_sync(0xffffffffu, pv, o);
    pj += __shfl_down_sync(0xffffffffu, pj, o);
    pl += __shfl_down_sync(0xffffffffu, pl, o);
  }
  if ((tid & 31) == 0) { sred[0][tid >> 5] = pv; sred[1][tid >> 5] = pj; sred[2][tid >> 5] = pl; }
  __syncthreads();
  if (tid < 3) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += sred[tid][w];
    partial[(b * n_tiles + blockIdx.x) * 3 + tid] = s;
  }
}

__device__ __forceinline__ float block_sum_256(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < 8; ++w) s += sh[w];
  return s;
}

// One CTA per word: RMSE terms, the six logged loss terms, dmel and dsv.
__global__ void __launch_bounds__(256)
word_loss_kernel(const float* __restrict__ mel, const float* __restrict__ tmel, const float* __restrict__ sv,
                 const float* __restrict__ tsv, const float* __restrict__ partial, int n_tiles,
                 float* __restrict__ terms, const int32_t* __restrict__ step_count, int slots,
                 float* __restrict__ dmel, float* __restrict__ dsv, int64_t Tmax, int64_t Tm_max,
                 const int32_t* __restrict__ word_T, int64_t B, int64_t C, int64_t Cm, int64_t S, int objective,
                 const float* __restrict__ cls_w, const float* __restrict__ cls_b,
                 const float* __restrict__ extra_terms, float* __restrict__ aux_log) {
  __shared__ float sh[8];
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x;
  const int64_t T = word_T ? (int64_t)word_T[b] : Tmax;        // this word's cp frames
  const int64_t Tm = word_T ? (int64_t)(word_T[b] / 2) : Tm_max;  // ... and mel frames
  const bool use_mel = objective != PAULE_OBJ_SEMVEC, use_sem = objective != PAULE_OBJ_ACOUSTIC;
  float sm = 0.f, sz = 0.f;   // sz: speech-classifier logit (LinearClassifier, models.py:887-911), summed over frames
  const int64_t nm = Tm * Cm;
  // thread = (channel tid % 64, frame tid / 64 + 4 i): a warp reads one 240-byte mel row segment per step, no division
  const int mc = tid & 63;
  const bool mcl = mc < (int)Cm && Cm <= 64;
  if (Cm <= 64) {
    if (mcl) {
      const float wz = cls_w ? __ldg(cls_w + mc) : 0.f;
#pragma unroll 8
      for (int64_t t = tid >> 6; t < Tm; t += 4) {   // independent loads: eight frames in flight per thread
        const int64_t off = (t * B + b) * Cm + mc;
        const float pm = __ldg(mel + off);
        const float d = pm - __ldg(tmel + off);
        sm += d * d;
        sz = fmaf(wz, pm, sz);
      }
    }
  } else {
    for (int64_t e = tid; e < nm; e += 256) {
      const int64_t t = e / Cm, c = e % Cm;
      const int64_t off = (t * B + b) * Cm + c;
      const float d = mel[off] - tmel[off];
      sm += d * d;
      if (cls_w) sz = fmaf(cls_w[c], mel[off], sz);
    }
  }
  sm = block_sum_256(sm, sh);
  float l_cls = 0.f, g_cls = 0.f;
  if (cls_w != nullptr) {   // 0.1 * BCEWithLogits(z, 0) = 0.1 softplus(z); d/dz = 0.1 sigmoid(z)   (paule.py:596,610-620)
    const float z = block_sum_256(sz, sh) / (float)Tm + (cls_b ? __ldg(cls_b) : 0.f);
    l_cls = kClassifierWeight * (fmaxf(z, 0.f) + log1pf(expf(-fabsf(z))));
    g_cls = kClassifierWeight / (1.f + expf(-z)) / (float)Tm;
  }
  float ss = 0.f;
  if (sv != nullptr)
    for (int64_t e = tid; e < S; e += 256) {
      const float d = sv[b * S + e] - tsv[b * S + e];
      ss += d * d;
    }
  ss = block_sum_256(ss, sh);
  float pv = 0.f, pj = 0.f, pl = 0.f;
  for (int i = 0; i < n_tiles; ++i) {  // fixed order
    pv += partial[(b * n_tiles + i) * 3 + 0];
    pj += partial[(b * n_tiles + i) * 3 + 1];
    pl += partial[(b * n_tiles + i) * 3 + 2];
  }
  const float rmse_m = sqrtf(sm / (float)nm);
  const float rmse_s = sqrtf(ss / (float)S);
  const float l_mel = kMelWeight * rmse_m;
  const float l_sem = (sv != nullptr) ? kSemWeight * rmse_s : 0.f;
  const float l_vel = kVelWeight * (pv / (float)((T - 4) * C));
  const float l_jerk = kJerkWeight * (pj / (float)((T - 12) * C));
  const float l_ll = kLocalLinearWeight * (pl / (float)((T - 2) * C));
  if (tid == 0) {
    float total;
    if (cls_w != nullptr) {   // summation order of the classifier variants, paule.py:621,681,736
      if (objective == PAULE_OBJ_ACOUSTIC_SEMVEC) total = l_mel + l_vel + l_jerk + l_sem + l_cls + l_ll;
      else if (objective == PAULE_OBJ_ACOUSTIC) total = l_mel + l_vel + l_jerk + l_ll + l_cls;
      else total = l_vel + l_jerk + l_sem + l_cls + l_ll;
    } else if (objective == PAULE_OBJ_ACOUSTIC_SEMVEC) total = l_mel + l_vel + l_jerk + l_sem + l_ll;  // paule.py:660
    else if (objective == PAULE_OBJ_ACOUSTIC) total = l_mel + l_vel + l_jerk + l_ll;            // :715
    else total = l_vel + l_jerk + l_sem + l_ll;                                                  // :771
    float x0 = 0.f, x1 = 0.f;
    if (extra_terms != nullptr) {   // somatosensory terms, evaluated by the caller: ... + tube_mel + tube_semvec (paule.py:643)
      x0 = extra_terms[b * 2 + 0]; x1 = extra_terms[b * 2 + 1];
      total = total + x0 + x1;
    }
    const int64_t slot = step_count ? (int64_t)((*step_count - 1) % slots + slots) % slots : 0;
    float* o = terms + (slot * B + b) * 6;
    o[0] = total; o[1] = l_mel; o[2] = l_sem; o[3] = l_vel; o[4] = l_jerk; o[5] = l_ll;
    if (aux_log != nullptr) {
      float* a = aux_log + (slot * B + b) * 3;
      a[0] = l_cls; a[1] = x0; a[2] = x1;
    }
  }
  // d(w*sqrt(mean(e^2)))/de = w*e/(N*rmse); eps = 0 -> NaN at zero error, as in the reference (paule.py:68)
  const float gm = use_mel ? kMelWeight / ((float)nm * rmse_m) : 0.f;
  if (Cm <= 64) {
    if (mcl) {
#pragma unroll 8
      for (int64_t t = tid >> 6; t < Tm_max; t += 4) {   // padded mel frames get a zero gradient
        const int64_t off = (t * B + b) * Cm + mc;
        float gval = (use_mel && t < Tm) ? gm * (__ldg(mel + off) - __ldg(tmel + off)) : 0.f;
        if (cls_w != nullptr && t < Tm) gval = fmaf(g_cls, __ldg(cls_w + mc), gval);
        dmel[off] = gval;
      }
    }
  } else {
    for (int64_t e = tid; e < Tm_max * Cm; e += 256) {
      const int64_t t = e / Cm, c = e % Cm;
      const int64_t off = (t * B + b) * Cm + c;
      float gval = (use_mel && t < Tm) ? gm * (mel[off] - tmel[off]) : 0.f;
      if (cls_w != nullptr && t < Tm) gval = fmaf(g_cls, cls_w[c], gval);
      dmel[off] = gval;
    }
  }
  if (dsv != nullptr) {
    const float gs = (use_sem && sv != nullptr) ? kSemWeight / ((float)S * rmse_s) : 0.f;
    for (int64_t e = tid; e < S; e += 256)
      dsv[b * S + e] = (use_sem && sv != nullptr) ? gs * (sv[b * S + e] - tsv[b * S + e]) : 0.f;
  }
}

__global__ void step_tick_kernel(int32_t* step_count) { *step_count += 1; }

// ragged batches: row b of `out` [B,H] = seq[word_T[b]/2 - 1, b, :]  (EmbeddingModel takes h at lens[b]-1, models.py:442)
__global__ void gather_last_kernel(const float* __restrict__ seq, const int32_t* __restrict__ word_T, float* __restrict__ out,
                                   int64_t B, int64_t H) {
  const int64_t b = blockIdx.x;
  const int64_t t = (int64_t)(word_T[b] / 2) - 1;
  for (int64_t j = threadIdx.x; j < H; j += blockDim.x) out[b * H + j] = seq[(t * B + b) * H + j];
}
// ... and its adjoint: seq[word_T[b]/2 - 1, b, :] = rows[b, :] in a zero-filled [Tm,B,H] buffer
__global__ void scatter_last_kernel(const float* __restrict__ rows, const int32_t* __restrict__ word_T, float* __restrict__ seq,
                                    int64_t B, int64_t H) {
  const int64_t b = blockIdx.x;
  const int64_t t = (int64_t)(word_T[b] / 2) - 1;
  for (int64_t j = threadIdx.x; j < H; j += blockDim.x) seq[(t * B + b) * H + j] = rows[b * H + j];
}

// torch/optim/adam.py::_single_tensor_adam (foreach/fused variants are arithmetic-equivalent):
//   m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, 1-b2); x.addcdiv_(m, sqrt(v)/sqrt(bc2) + eps, -lr/bc1)
__global__ void __launch_bounds__(256)
adam_clamp_kernel(float* __restrict__ cp, const float* __restrict__ g_a, const float* __restrict__ g_b,
                  const float* __restrict__ g_c, float* __restrict__ m, float* __restrict__ v, const int32_t* __restrict__ step_count, float lr,
                  float beta1, float beta2, float eps, float clampv, int smiling, const float* __restrict__ past_cp,
                  int64_t past_n, float* __restrict__ grad_out, int64_t n, int C) {
  __shared__ float s_step_size, s_bc2_sqrt;
  if (threadIdx.x == 0) {
    const int step = *step_count;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    s_step_size = (float)((double)lr / bc1);
    s_bc2_sqrt = (float)sqrt(bc2);
  }
  __syncthreads();
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  const float w1 = 1.0f - beta1, w2 = 1.0f - beta2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t e0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; e0 < n; e0 += stride) {
    float x[4], g[4], mm[4], vv[4];
    const bool full = (e0 + 4 <= n);
    if (full) {
      const float4 X = *reinterpret_cast<const float4*>(cp + e0);
      const float4 G = *reinterpret_cast<const float4*>(g_a + e0);
      const float4 M = *reinterpret_cast<const float4*>(m + e0);
      const float4 V = *reinterpret_cast<const float4*>(v + e0);
      x[0] = X.x; x[1] = X.y; x[2] = X.z; x[3] = X.w;
      g[0] = G.x; g[1] = G.y; g[2] = G.z; g[3] = G.w;
      mm[0] = M.x; mm[1] = M.y; mm[2] = M.z; mm[3] = M.w;
      vv[0] = V.x; vv[1] = V.y; vv[2] = V.z; vv[3] = V.w;
      if (g_b) {
        const float4 G2 = *reinterpret_cast<const float4*>(g_b + e0);
        g[0] += G2.x; g[1] += G2.y; g[2] += G2.z; g[3] += G2.w;
      }
      if (g_c) {   // gradient of loss branches evaluated outside the fused step (somatosensory, paule.py:624-645)
        const float4 G3 = *reinterpret_cast<const float4*>(g_c + e0);
        g[0] += G3.x; g[1] += G3.y; g[2] += G3.z; g[3] += G3.w;
      }
    } else {
      for (int i = 0; i < 4; ++i) {
        const bool ok = e0 + i < n;
        x[i] = ok ? cp[e0 + i] : 0.f;
        g[i] = ok ? g_a[e0 + i] + (g_b ? g_b[e0 + i] : 0.f) + (g_c ? g_c[e0 + i] : 0.f) : 0.f;
        mm[i] = ok ? m[e0 + i] : 0.f;
        vv[i] = ok ? v[e0 + i] : 0.f;
      }
    }
    if (grad_out != nullptr)
      for (int i = 0; i < 4 && e0 + i < n; ++i) grad_out[e0 + i] = g[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      mm[i] = mm[i] + w1 * (g[i] - mm[i]);
      vv[i] = vv[i] * beta2 + (w2 * g[i]) * g[i];
      const float denom = sqrtf(vv[i]) / bc2_sqrt + eps;
      float xn = x[i] - step_size * (mm[i] / denom);
      xn = fminf(fmaxf(xn, -clampv), clampv);
      const int64_t e = e0 + i;
      if (smiling) {
        const int c = (int)(e % C);
        if (c == 4) xn = -1.0f;   // "LP"  paule.py:1207
        if (c == 1) xn = 1.0f;    // "HY"  paule.py:1208
      }
      if (past_cp != nullptr && e < past_n) xn = past_cp[e];
      x[i] = xn;
    }
    if (full) {
      *reinterpret_cast<float4*>(cp + e0) = make_float4(x[0], x[1], x[2], x[3]);
      *reinterpret_cast<float4*>(m + e0) = make_float4(mm[0], mm[1], mm[2], mm[3]);
      *reinterpret_cast<float4*>(v + e0) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    } else {
      for (int i = 0; i < 4 && e0 + i < n; ++i) { cp[e0 + i] = x[i]; m[e0 + i] = mm[i]; v[e0 + i] = vv[i]; }
    }
  }
}

}  // namespace paule

using namespace paule;

extern "C" size_t paule_plan_loss_scratch_floats(int64_t T, int64_t B) {
  return (size_t)(B * ceil_div(T, (int64_t)kTileT) * 3);
}

namespace paule {

int gather_last(const float* seq, const int32_t* word_T, float* out, int64_t B, int64_t H, paule_stream_t stream) {
  gather_last_kernel<<<(unsigned)B, 256, 0, as_stream(stream)>>>(seq, word_T, out, B, H);
  PAULE_LAUNCH_CHECK("gather_last_kernel");
  return PAULE_OK;
}

int scatter_last(const float* rows, const int32_t* word_T, float* seq, int64_t Tm, int64_t B, int64_t H, paule_stream_t stream) {
  PAULE_CUDA(cudaMemsetAsync(seq, 0, (size_t)(Tm * B * H) * sizeof(float), as_stream(stream)));
  scatter_last_kernel<<<(unsigned)B, 256, 0, as_stream(stream)>>>(rows, word_T, seq, B, H);
  PAULE_LAUNCH_CHECK("scatter_last_kernel");
  return PAULE_OK;
}

int plan_loss_logged(const float* mel, const float* tmel, const float* sv, const float* tsv, const float* cp,
                     float* terms, const int32_t* step_count, int slots, float* dmel, float* dsv, float* dcp_smooth,
                     float* scratch, int64_t T, int64_t Tm, const int32_t* word_T, int64_t B, int64_t C, int64_t Cm,
                     int64_t S, int objective, paule_stream_t stream, const float* cls_w, const float* cls_b,
                     const float* extra_terms, float* aux_log, int which) {
  PAULE_REQUIRE(mel && tmel && cp && terms && dmel && dcp_smooth && scratch && slots >= 1);
  PAULE_REQUIRE((sv == nullptr) == (tsv == nullptr));
  PAULE_REQUIRE(T >= 13 && Tm >= 1 && B >= 1 && C >= 1 && C <= kMaxC && Cm >= 1 && S >= 1);
  PAULE_REQUIRE(objective >= 0 && objective <= 2);
  if (objective != PAULE_OBJ_ACOUSTIC) PAULE_REQUIRE(sv && tsv && dsv);
  const int n_tiles = (int)ceil_div(T, (int64_t)kTileT);
  if (which != 2) {
    smooth_terms_kernel<<<dim3(n_tiles, (unsigned)B), 256, 0, as_stream(stream)>>>(cp, dcp_smooth, scratch, T, word_T, B, C,
                                                                                   n_tiles);
    PAULE_LAUNCH_CHECK("smooth_terms_kernel");
  }
  if (which == 1) return PAULE_OK;
  word_loss_kernel<<<(unsigned)B, 256, 0, as_stream(stream)>>>(mel, tmel, sv, tsv, scratch, n_tiles, terms,
                                                               step_count, slots, dmel, dsv, T, Tm, word_T, B, C, Cm,
                                                               S, objective, cls_w, cls_b, extra_terms, aux_log);
  PAULE_LAUNCH_CHECK("word_loss_kernel");
  return PAULE_OK;
}

int adam_clamp_logged(float* cp, const float* g_a, const float* g_b, float* m, float* v, const int32_t* step_count,
                      float lr, float beta1, float beta2, float eps, float clamp, int smiling, const float* past_cp,
                      int64_t past_T, float* grad_out, int64_t T, int64_t B, int64_t C, paule_stream_t stream,
                      const float* g_c) {
  PAULE_REQUIRE(cp && g_a && m && v && step_count && T >= 0 && B > 0 && C > 0);
  PAULE_REQUIRE(past_T >= 0 && past_T <= T && (past_T == 0 || past_cp));
  PAULE_REQUIRE(!smiling || C > 4);
  const int64_t n = T * B * C;
  if (n == 0) return PAULE_OK;
  PAULE_REQUIRE((reinterpret_cast<uintptr_t>(cp) | reinterpret_cast<uintptr_t>(g_a) | reinterpret_cast<uintptr_t>(m) |
                 reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(g_b) | reinterpret_cast<uintptr_t>(g_c)) % 16 == 0);
  int64_t blocks = ceil_div(n, (int64_t)256 * 4);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  adam_clamp_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(cp, g_a, g_b, g_c, m, v, step_count, lr, beta1, beta2,
                                                                     eps, clamp, smiling, past_T ? past_cp : nullptr,
                                                                     past_T * B * C, grad_out, n, (int)C);
  PAULE_LAUNCH_CHECK("adam_clamp_kernel");
  return PAULE_OK;
}

}  // namespace paule

extern "C" int paule_plan_loss_f32(const float* mel, const float* tmel, const float* sv, const float* tsv,
                                   const float* cp, float* terms, float* dmel, float* dsv, float* dcp_smooth,
                                   float* scratch, int64_t T, int64_t Tm, int64_t B, int64_t C, int64_t Cm,
                                   int64_t S, int objective, paule_stream_t stream) {
  return plan_loss_logged(mel, tmel, sv, tsv, cp, terms, nullptr, 1, dmel, dsv, dcp_smooth, scratch, T, Tm, nullptr, B, C,
                          Cm, S, objective, stream);
}

extern "C" int paule_step_tick(int32_t* step_count, paule_stream_t stream) {
  PAULE_REQUIRE(step_count);
  step_tick_kernel<<<1, 1, 0, as_stream(stream)>>>(step_count);
  PAULE_LAUNCH_CHECK("step_tick_kernel");
  return PAULE_OK;
}

extern "C" int paule_adam_clamp_f32(float* cp, const float* g_a, const float* g_b, float* m, float* v,
                                    const int32_t* step_count, float lr, float beta1, float beta2, float eps,
                                    float clamp, int smiling, const float* past_cp, int64_t past_T, int64_t T,
                                    int64_t B, int64_t C, paule_stream_t stream) {
  return adam_clamp_logged(cp, g_a, g_b, m, v, step_count, lr, beta1, beta2, eps, clamp, smiling, past_cp, past_T,
                           nullptr, T, B, C, stream);
}
