// Generic fp32 operators: y = x W^T + b with mapped rows, [B,T,C] <-> [T,B,C] transposes.
// These are the parity anchor (FFMA, fp32 everywhere) and serve every shape the tcgen05 kernels
// are not specialised for (skinny K = 30 / 60 input projections, heads, arbitrary hidden sizes).
#include "common.cuh"

namespace paule {

thread_local char g_last_cuda_error[256] = "";

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

struct RowMap {
  int64_t inner, outer_stride, inner_stride;
  __device__ __forceinline__ int64_t operator()(int64_t r) const {
    return (r / inner) * outer_stride + (r % inner) * inner_stride;
  }
};

// C tile BM x BN, K chunk BK, 256 threads, each thread a TM x TN register tile.
template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
linear_f32_kernel(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias,
                  float* __restrict__ C, int64_t M, int64_t N, int64_t K, RowMap amap, int64_t a_pair,
                  RowMap cmap, int accumulate) {
  constexpr int NT = (BM / TM) * (BN / TN);
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = 0; k0 < K; k0 += BK) {
    // A tile: consecutive threads walk k (contiguous in memory)
    for (int e = tid; e < BM * BK; e += NT) {
      const int r = e / BK, kk = e % BK;
      const int64_t gr = m0 + r, gk = k0 + kk;
      float v = 0.f;
      if (gr < M && gk < K) {
        const int64_t off = amap(gr) + gk;
        v = __ldg(A + off);
        if (a_pair != 0) v = 0.5f * (v + __ldg(A + off + a_pair));
      }
      As[kk][r] = v;
    }
    for (int e = tid; e < BN * BK; e += NT) {
      const int r = e / BK, kk = e % BK;
      const int64_t gn = n0 + r, gk = k0 + kk;
      Ws[kk][r] = (gn < N && gk < K) ? __ldg(W + gn * K + gk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], w[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) w[j] = Ws[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t gr = m0 + ty * TM + i;
    if (gr >= M) continue;
    float* crow = C + cmap(gr);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int64_t gn = n0 + tx * TN + j;
      if (gn >= N) continue;
      float v = acc[i][j] + (bias ? __ldg(bias + gn) : 0.f);
      if (accumulate) v += crow[gn];
      crow[gn] = v;
    }
  }
}

// Skinny-K projection (K <= 64, contiguous rows): y[M,N] = x[M,K] W[N,K]^T + b.  The whole K extent of a 128 x 128 output
// tile sits in shared memory (transposed, so that a thread's 8 rows / 8 columns are two 16-byte reads each), every
// thread owns an 8 x 8 register tile: 64 FFMA per 4 LDS.128, rows written as 32-byte segments.  Used for the K = 30 / 60
// input projections (147 MB of output per launch at B=64, T=200) and the K = 60 post_linear^T.
constexpr int kSkM = 128, kSkN = 128, kSkPad = 4;
__global__ void __launch_bounds__(256)
linear_skinny_f32_kernel(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias,
                         float* __restrict__ C, int64_t M, int64_t N, int K, int accumulate) {
  extern __shared__ __align__(16) float sk_smem[];
  float* As = sk_smem;                                   // [K][128 + 4]
  float* Ws = sk_smem + (size_t)K * (kSkM + kSkPad);     // [K][128 + 4]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * kSkM, n0 = (int64_t)blockIdx.x * kSkN;
  const int64_t a_rows = (M - m0 < kSkM) ? (M - m0) : kSkM, w_rows = (N - n0 < kSkN) ? (N - n0) : kSkN;
  const float* a_src = A + m0 * K;
  const float* w_src = W + n0 * K;
  for (int e = tid; e < kSkM * K; e += 256) {
    const int r = e / K, kk = e - r * K;
    As[kk * (kSkM + kSkPad) + r] = (r < a_rows) ? __ldg(a_src + e) : 0.f;
  }
  for (int e = tid; e < kSkN * K; e += 256) {
    const int r = e / K, kk = e - r * K;
    Ws[kk * (kSkN + kSkPad) + r] = (r < w_rows) ? __ldg(w_src + e) : 0.f;
  }
  __syncthreads();
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 4
  for (int kk = 0; kk < K; ++kk) {
    const float4 a0 = *reinterpret_cast<const float4*>(As + kk * (kSkM + kSkPad) + ty * 8);
    const float4 a1 = *reinterpret_cast<const float4*>(As + kk * (kSkM + kSkPad) + ty * 8 + 4);
    const float4 w0 = *reinterpret_cast<const float4*>(Ws + kk * (kSkN + kSkPad) + tx * 8);
    const float4 w1 = *reinterpret_cast<const float4*>(Ws + kk * (kSkN + kSkPad) + tx * 8 + 4);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
  }
  const int64_t gn = n0 + tx * 8;
  const bool vec = ((N & 3) == 0) && (gn + 8 <= N);
  float bv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bv[j] = (bias != nullptr && gn + j < N) ? __ldg(bias + gn + j) : 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t gr = m0 + ty * 8 + i;
    if (gr >= M) continue;
    float* crow = C + gr * N + gn;
    if (vec) {
      float4 o0 = make_float4(acc[i][0] + bv[0], acc[i][1] + bv[1], acc[i][2] + bv[2], acc[i][3] + bv[3]);
      float4 o1 = make_float4(acc[i][4] + bv[4], acc[i][5] + bv[5], acc[i][6] + bv[6], acc[i][7] + bv[7]);
      if (accumulate) {
        const float4 p0 = *reinterpret_cast<const float4*>(crow), p1 = *reinterpret_cast<const float4*>(crow + 4);
        o0.x += p0.x; o0.y += p0.y; o0.z += p0.z; o0.w += p0.w;
        o1.x += p1.x; o1.y += p1.y; o1.z += p1.z; o1.w += p1.w;
      }
      *reinterpret_cast<float4*>(crow) = o0;
      *reinterpret_cast<float4*>(crow + 4) = o1;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (gn + j >= N) continue;
        float v = acc[i][j] + bv[j];
        if (accumulate) v += crow[j];
        crow[j] = v;
      }
    }
  }
}

// Few-row projection (M <= 64 rows, e.g. the heads: one row per word): a CTA owns 8 output columns x 8 rows, keeps the
// weight rows in shared memory and gives every warp ONE row; the row's dot products are split over the lanes along K
// (coalesced, unrolled reads of x) and reduced with shuffles.  grid = (N / 8, M / 8): a single wave of short CTAs -- the
// generic tile kernel needs 45 dependent load/sync rounds on 5 CTAs here.
__global__ void __launch_bounds__(256)
linear_fewrows_f32_kernel(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias,
                          float* __restrict__ C, int M, int64_t N, int K, int accumulate) {
  extern __shared__ __align__(16) float fr_smem[];   // [8][K]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n0 = (int64_t)blockIdx.x * 8;
  const int m = blockIdx.y * 8 + warp;
  // this warp's row of x: issued before the weight tile so that both loads overlap
  const float* arow = A + (size_t)(m < M ? m : 0) * K;
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = (m < M && lane + 32 * i < K) ? __ldg(arow + lane + 32 * i) : 0.f;
  for (int e = tid; e < 8 * K; e += 256) {
    const int j = e / K, kk = e - j * K;
    fr_smem[e] = (n0 + j < N) ? __ldg(W + (n0 + j) * K + kk) : 0.f;
  }
  __syncthreads();
  if (m >= M) return;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += 256) {   // 8 x 32 elements of the row per round, prefetched one round ahead
    float an[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) an[i] = (k0 + 256 + lane + 32 * i < K) ? __ldg(arow + k0 + 256 + lane + 32 * i) : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int kk = k0 + lane + 32 * i;
      if (kk < K) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(a[i], fr_smem[j * K + kk], acc[j]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = an[i];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  if (lane < 8 && n0 + lane < N) {
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) v = (lane == j) ? acc[j] : v;
    v += bias ? __ldg(bias + n0 + lane) : 0.f;
    float* dst = C + (size_t)m * N + n0 + lane;
    if (accumulate) v += *dst;
    *dst = v;
  }
}

__global__ void transpose_btc_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n_outer,
                                     int64_t n_inner, int64_t C) {
  // out[i, o, :] = in[o, i, :]; one thread per element, coalesced on the store side (C contiguous on both).
  const int64_t total = n_outer * n_inner * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = e % C, rest = e / C;
    const int64_t o = rest % n_outer, i = rest / n_outer;
    out[e] = __ldg(in + (o * n_inner + i) * C + c);
  }
}

}  // namespace paule

using namespace paule;

extern "C" int paule_linear_f32(const float* A, const float* W, const float* bias, float* C, int64_t M, int64_t N,
                                int64_t K, int64_t a_inner, int64_t a_outer_stride, int64_t a_inner_stride,
                                int64_t a_pair_stride, int64_t c_inner, int64_t c_outer_stride,
                                int64_t c_inner_stride, int accumulate, paule_stream_t stream) {
  PAULE_REQUIRE(A && W && C);
  PAULE_REQUIRE(M >= 0 && N > 0 && K > 0 && a_inner > 0 && c_inner > 0);
  if (M == 0) return PAULE_OK;
  RowMap am{a_inner, a_outer_stride, a_inner_stride}, cm{c_inner, c_outer_stride, c_inner_stride};
  // contiguous rows on both sides (row r of x at r*K, row r of y at r*N) and no pooled pair: the specialised kernels apply
  const bool plain = a_pair_stride == 0 && (a_inner == 1 ? a_outer_stride == K : (a_inner_stride == K && a_outer_stride == a_inner * K)) &&
                     (c_inner == 1 ? c_outer_stride == N : (c_inner_stride == N && c_outer_stride == c_inner * N));
  if (plain && M <= 64 && N >= 64 && K <= 2048) {
    const size_t smem = (size_t)8 * K * sizeof(float);
    static unsigned long long attr_set = 0ull;
    if (once_per_device(attr_set)) {
      PAULE_CUDA(cudaFuncSetAttribute(linear_fewrows_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2048 * 4));
    }
    linear_fewrows_f32_kernel<<<dim3((unsigned)ceil_div(N, 8), (unsigned)ceil_div(M, 8)), 256, smem, as_stream(stream)>>>(
        A, W, bias, C, (int)M, N, (int)K, accumulate);
  } else if (plain && K <= 64 && M >= 256 && N >= 64) {
    const size_t smem = (size_t)K * (kSkM + kSkPad + kSkN + kSkPad) * sizeof(float);
    static unsigned long long attr_set = 0ull;
    if (once_per_device(attr_set)) {
      PAULE_CUDA(cudaFuncSetAttribute(linear_skinny_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      64 * (kSkM + kSkPad + kSkN + kSkPad) * 4));
    }
    dim3 grid((unsigned)ceil_div(N, kSkN), (unsigned)ceil_div(M, kSkM));
    linear_skinny_f32_kernel<<<grid, 256, smem, as_stream(stream)>>>(A, W, bias, C, M, N, (int)K, accumulate);
  } else if (M <= 64 && N > 32) {
    // few rows (heads: M = words): narrow column tiles so that the grid still covers many SMs
    constexpr int BM = 64, BN = 16, BK = 16, TM = 4, TN = 1;
    dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, BM));
    linear_f32_kernel<BM, BN, BK, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, as_stream(stream)>>>(
        A, W, bias, C, M, N, K, am, a_pair_stride, cm, accumulate);
  } else if (N <= 32) {
    constexpr int BM = 128, BN = 32, BK = 16, TM = 4, TN = 4;
    dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, BM));
    linear_f32_kernel<BM, BN, BK, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, as_stream(stream)>>>(
        A, W, bias, C, M, N, K, am, a_pair_stride, cm, accumulate);
  } else {
    constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
    dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, BM));
    linear_f32_kernel<BM, BN, BK, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, as_stream(stream)>>>(
        A, W, bias, C, M, N, K, am, a_pair_stride, cm, accumulate);
  }
  PAULE_LAUNCH_CHECK("linear_f32_kernel");
  return PAULE_OK;
}

extern "C" int paule_transpose_btc(const float* in, float* out, int64_t n_outer, int64_t n_inner, int64_t C,
                                   paule_stream_t stream) {
  PAULE_REQUIRE(in && out && n_outer >= 0 && n_inner >= 0 && C > 0);
  const int64_t total = n_outer * n_inner * C;
  if (total == 0) return PAULE_OK;
  const int threads = 256;
  int64_t blocks = ceil_div(total, threads);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  transpose_btc_kernel<<<(unsigned)blocks, threads, 0, as_stream(stream)>>>(in, out, n_outer, n_inner, C);
  PAULE_LAUNCH_CHECK("transpose_btc_kernel");
  return PAULE_OK;
}
