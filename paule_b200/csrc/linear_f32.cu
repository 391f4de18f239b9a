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
  if (M <= 64 && N > 32) {
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
