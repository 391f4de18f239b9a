// fp32 LSTM recurrence, one kernel per time step (forward and reverse-time BPTT).
//
// This is the parity anchor: FFMA with fp32 operands, the reference's gate order (i,f,g,o) and
// cell equations (torch.nn.LSTM; /root/reference/paule/models.py:349, :441).  Each step is a skinny
// GEMM  [B,K] x [N,K]^T  (forward K = H, backward K = 4H) followed by the pointwise cell; to keep
// enough warps in flight at small B the K range is split across KG thread groups inside the CTA
// and reduced through shared memory before the pointwise epilogue.
#include "common.cuh"

namespace paule {

constexpr int kRows = 32;   // batch rows per CTA
constexpr int kUnits = 16;  // hidden units per CTA
constexpr int kBK = 16;     // K chunk

// acc[r][c] += sum_{k in [kbeg,kend)} A[b0+2ty+r, k] * W[wrow(c), k]
// 128 threads per K-group laid out as ty (16) x tx (8); NC columns per thread.
template <int NC, class WRow>
__device__ __forceinline__ void group_gemm(const float* __restrict__ A, int64_t lda, const float* __restrict__ W,
                                           int64_t ldw, WRow wrow, int64_t b0, int64_t B, int64_t kbeg,
                                           int64_t kend, float* As, float* Ws, int gtid, float (&acc)[2][NC]) {
  constexpr int NCOLS = 8 * NC;
  constexpr int LDA_S = kRows + 2;
  constexpr int LDW_S = NCOLS + 4;
  const int tx = gtid % 8, ty = gtid / 8;
  for (int64_t k0 = kbeg; k0 < kend; k0 += kBK) {
    for (int e = gtid; e < kRows * kBK; e += 128) {
      const int r = e / kBK, kk = e % kBK;
      const int64_t b = b0 + r, k = k0 + kk;
      As[kk * LDA_S + r] = (b < B && k < kend) ? __ldg(A + b * lda + k) : 0.f;
    }
    for (int e = gtid; e < NCOLS * kBK; e += 128) {
      const int c = e / kBK, kk = e % kBK;
      const int64_t k = k0 + kk;
      const int64_t wr = wrow(c);
      Ws[kk * LDW_S + c] = (wr >= 0 && k < kend) ? __ldg(W + wr * ldw + k) : 0.f;
    }
    asm volatile("bar.sync %0, 128;" ::"r"(1 + (int)(threadIdx.x / 128)));
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      const float2 a = *reinterpret_cast<const float2*>(As + kk * LDA_S + 2 * ty);
      float w[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) w[c] = Ws[kk * LDW_S + tx * NC + c];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        acc[0][c] = fmaf(a.x, w[c], acc[0][c]);
        acc[1][c] = fmaf(a.y, w[c], acc[1][c]);
      }
    }
    asm volatile("bar.sync %0, 128;" ::"r"(1 + (int)(threadIdx.x / 128)));
  }
}

// ------------------------------------------------------------------------------------------------
// forward step: gates_t <- act(gates_t + h_prev W_hh^T); c_t, h_t
// ------------------------------------------------------------------------------------------------
template <int KG>
__global__ void __launch_bounds__(128 * KG)
lstm_step_fwd_f32(float* __restrict__ gates_t, const float* __restrict__ h_prev, const float* __restrict__ c_prev,
                  const float* __restrict__ w_hh, float* __restrict__ h_out, float* __restrict__ c_out, int64_t B,
                  int64_t H) {
  constexpr int NC = 8;  // 2 units x 4 gates per thread; column = unit*4 + gate
  constexpr int NCOLS = 64;
  constexpr int TILE = kBK * (kRows + 2) + kBK * (NCOLS + 4);
  constexpr int RED = kRows * NCOLS;
  __shared__ __align__(16) float smem[(KG * TILE > KG * RED) ? KG * TILE : KG * RED];
  const int grp = threadIdx.x / 128, gtid = threadIdx.x % 128;
  const int64_t j0 = (int64_t)blockIdx.x * kUnits, b0 = (int64_t)blockIdx.y * kRows;

  float acc[2][NC];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[r][c] = 0.f;

  if (h_prev != nullptr) {
    const int64_t per = ceil_div(H, (int64_t)KG);
    const int64_t kbeg = grp * per, kend = (kbeg + per < H) ? kbeg + per : H;
    auto wrow = [=](int c) -> int64_t {
      const int64_t j = j0 + c / 4;
      return (j < H) ? (int64_t)(c % 4) * H + j : -1;
    };
    float* As = smem + grp * TILE;
    float* Ws = As + kBK * (kRows + 2);
    group_gemm<NC>(h_prev, H, w_hh, H, wrow, b0, B, kbeg, kend, As, Ws, gtid, acc);
  }
  __syncthreads();
  {
    const int tx = gtid % 8, ty = gtid / 8;
    float* red = smem + grp * RED;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < NC; ++c) red[(2 * ty + r) * NCOLS + tx * NC + c] = acc[r][c];
  }
  __syncthreads();
  // pointwise: one (row, unit) cell per thread
  for (int cell = threadIdx.x; cell < kRows * kUnits; cell += 128 * KG) {
    const int r = cell / kUnits, u = cell % kUnits;
    const int64_t b = b0 + r, j = j0 + u;
    if (b >= B || j >= H) continue;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int g = 0; g < KG; ++g) {
      const float4 p = *reinterpret_cast<const float4*>(smem + g * RED + r * NCOLS + u * 4);
      a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
    }
    float* grow = gates_t + b * 4 * H;
    const float gi = sigmoidf_acc(a.x + grow[j]);
    const float gf = sigmoidf_acc(a.y + grow[H + j]);
    const float gg = tanhf(a.z + grow[2 * H + j]);
    const float go = sigmoidf_acc(a.w + grow[3 * H + j]);
    const float cp = c_prev ? c_prev[b * H + j] : 0.f;
    const float c = gf * cp + gi * gg;
    grow[j] = gi; grow[H + j] = gf; grow[2 * H + j] = gg; grow[3 * H + j] = go;
    c_out[b * H + j] = c;
    h_out[b * H + j] = go * tanhf(c);
  }
}

// ------------------------------------------------------------------------------------------------
// backward step t:  dh = dh_ext + da_{t+1} W_hh ;  (da_t, dc) <- cell adjoint
// ------------------------------------------------------------------------------------------------
template <int KG>
__global__ void __launch_bounds__(128 * KG)
lstm_step_bwd_f32(float* __restrict__ gates_t, const float* __restrict__ da_next, const float* __restrict__ c_t,
                  const float* __restrict__ c_prev, const float* __restrict__ w_hh_t,
                  const float* __restrict__ dh_ext, float dh_scale, const float* __restrict__ dh_last,
                  float* __restrict__ dc, int dc_is_zero, int64_t B, int64_t H) {
  constexpr int NC = 2;  // 2 units per thread
  constexpr int NCOLS = 16;
  constexpr int TILE = kBK * (kRows + 2) + kBK * (NCOLS + 4);
  constexpr int RED = kRows * NCOLS;
  __shared__ __align__(16) float smem[(KG * TILE > KG * RED) ? KG * TILE : KG * RED];
  const int grp = threadIdx.x / 128, gtid = threadIdx.x % 128;
  const int64_t j0 = (int64_t)blockIdx.x * kUnits, b0 = (int64_t)blockIdx.y * kRows;
  const int64_t K = 4 * H;

  float acc[2][NC];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[r][c] = 0.f;

  if (da_next != nullptr) {
    const int64_t per = ceil_div(K, (int64_t)KG);
    const int64_t kbeg = grp * per, kend = (kbeg + per < K) ? kbeg + per : K;
    auto wrow = [=](int c) -> int64_t { return (j0 + c < H) ? j0 + c : -1; };
    float* As = smem + grp * TILE;
    float* Ws = As + kBK * (kRows + 2);
    group_gemm<NC>(da_next, K, w_hh_t, K, wrow, b0, B, kbeg, kend, As, Ws, gtid, acc);
  }
  __syncthreads();
  {
    const int tx = gtid % 8, ty = gtid / 8;
    float* red = smem + grp * RED;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < NC; ++c) red[(2 * ty + r) * NCOLS + tx * NC + c] = acc[r][c];
  }
  __syncthreads();
  for (int cell = threadIdx.x; cell < kRows * kUnits; cell += 128 * KG) {
    const int r = cell / kUnits, u = cell % kUnits;
    const int64_t b = b0 + r, j = j0 + u;
    if (b >= B || j >= H) continue;
    float dh = 0.f;
#pragma unroll
    for (int g = 0; g < KG; ++g) dh += smem[g * RED + r * NCOLS + u];
    if (dh_ext) dh += dh_scale * dh_ext[b * H + j];
    if (dh_last) dh += dh_last[b * H + j];
    float* grow = gates_t + b * 4 * H;
    const float gi = grow[j], gf = grow[H + j], gg = grow[2 * H + j], go = grow[3 * H + j];
    const float tc = tanhf(c_t[b * H + j]);
    const float cp = c_prev ? c_prev[b * H + j] : 0.f;
    const float d_o = dh * tc;
    const float dct = (dc_is_zero ? 0.f : dc[b * H + j]) + dh * go * (1.f - tc * tc);
    grow[j] = dct * gg * gi * (1.f - gi);
    grow[H + j] = dct * cp * gf * (1.f - gf);
    grow[2 * H + j] = dct * gi * (1.f - gg * gg);
    grow[3 * H + j] = d_o * go * (1.f - go);
    dc[b * H + j] = dct * gf;
  }
}

}  // namespace paule

using namespace paule;

extern "C" int paule_lstm_seq_fwd_f32(float* gates, const float* w_hh, float* h, float* c, int64_t T, int64_t B,
                                      int64_t H, paule_stream_t stream) {
  PAULE_REQUIRE(gates && w_hh && h && c && T >= 0 && B > 0 && H > 0);
  constexpr int KG = 4;
  dim3 grid((unsigned)ceil_div(H, kUnits), (unsigned)ceil_div(B, kRows));
  for (int64_t t = 0; t < T; ++t) {
    const float* hp = t ? h + (t - 1) * B * H : nullptr;
    const float* cp = t ? c + (t - 1) * B * H : nullptr;
    lstm_step_fwd_f32<KG><<<grid, 128 * KG, 0, as_stream(stream)>>>(gates + t * B * 4 * H, hp, cp, w_hh,
                                                                    h + t * B * H, c + t * B * H, B, H);
  }
  PAULE_LAUNCH_CHECK("lstm_step_fwd_f32");
  return PAULE_OK;
}

extern "C" int paule_lstm_seq_bwd_f32(float* gates, const float* c, const float* w_hh_t, const float* dh_seq,
                                      int dh_mode, const float* dh_last, float* scratch, int64_t T, int64_t B,
                                      int64_t H, paule_stream_t stream) {
  PAULE_REQUIRE(gates && c && w_hh_t && scratch && T >= 0 && B > 0 && H > 0);
  PAULE_REQUIRE(dh_mode == 0 || ((dh_mode == 1 || dh_mode == 2) && dh_seq));
  constexpr int KG = 8;
  dim3 grid((unsigned)ceil_div(H, kUnits), (unsigned)ceil_div(B, kRows));
  for (int64_t t = T - 1; t >= 0; --t) {
    const float* da_next = (t + 1 < T) ? gates + (t + 1) * B * 4 * H : nullptr;
    const float* cprev = t ? c + (t - 1) * B * H : nullptr;
    const float* dext = nullptr;
    float scale = 1.f;
    if (dh_mode == 1) {
      dext = dh_seq + t * B * H;
    } else if (dh_mode == 2) {
      if (t / 2 < T / 2) { dext = dh_seq + (t / 2) * B * H; scale = 0.5f; }
    }
    lstm_step_bwd_f32<KG><<<grid, 128 * KG, 0, as_stream(stream)>>>(
        gates + t * B * 4 * H, da_next, c + t * B * H, cprev, w_hh_t, dext, scale,
        (t == T - 1) ? dh_last : nullptr, scratch, (t == T - 1) ? 1 : 0, B, H);
  }
  PAULE_LAUNCH_CHECK("lstm_step_bwd_f32");
  return PAULE_OK;
}
