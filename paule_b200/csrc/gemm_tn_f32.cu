// Weight gradients of the continue-learning step (reference: autograd through aten::lstm / aten::linear in
// pred_loss.backward(), /root/reference/paule/paule.py:1376): a reduction over ALL (time step, word) rows
//
//     dW[m, n] (+)= sum_r dA[r, m] X[r, n]        dA [R, M] (d loss / d pre-activation), X [R, N] (layer input or h_{t-1})
//     db[m]    (+)= sum_r dA[r, m]
//
// fp32 FFMA, 64 x 64 output tiles, 16-row chunks of the reduction staged in shared memory, split over the rows (gridDim.z) with
// an atomic epilogue when the output has few tiles.  Outer-loop work (a handful of small batches between planning rounds), so
// the kernel aims at "own, correct and not slow", not at the tensor roofline.
#include "common.cuh"

namespace paule {

constexpr int kTnTile = 64, kTnChunk = 16;

__global__ void __launch_bounds__(256) gemm_tn_f32_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                          float* __restrict__ C, int64_t R, int M, int N, int64_t lda,
                                                          int64_t ldb, int rows_per_split, int atomic) {
  __shared__ float As[kTnChunk][kTnTile + 4];
  __shared__ float Bs[kTnChunk][kTnTile + 4];
  const int m0 = blockIdx.x * kTnTile, n0 = blockIdx.y * kTnTile;
  const int64_t r_begin = (int64_t)blockIdx.z * rows_per_split;
  const int64_t r_end = (r_begin + rows_per_split < R) ? r_begin + rows_per_split : R;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // thread's 4 x 4 block: rows 4 ty.., columns 4 tx..
  float acc[4][4] = {};
  for (int64_t r0 = r_begin; r0 < r_end; r0 += kTnChunk) {
    for (int e = threadIdx.x; e < kTnChunk * kTnTile; e += 256) {
      const int rr = e / kTnTile, cc = e % kTnTile;
      const int64_t r = r0 + rr;
      As[rr][cc] = (r < r_end && m0 + cc < M) ? A[r * lda + m0 + cc] : 0.f;
      Bs[rr][cc] = (r < r_end && n0 + cc < N) ? B[r * ldb + n0 + cc] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < kTnChunk; ++rr) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[rr][4 * ty + i]; b[i] = Bs[rr][4 * tx + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + 4 * ty + i, n = n0 + 4 * tx + j;
      if (m < M && n < N) {
        if (atomic) atomicAdd(&C[(int64_t)m * N + n], acc[i][j]);
        else C[(int64_t)m * N + n] = acc[i][j];
      }
    }
}

// column sums: out[m] (+)= sum_r A[r, m]; one thread per column and row split, coalesced along m
__global__ void colsum_f32_kernel(const float* __restrict__ A, float* __restrict__ out, int64_t R, int M, int64_t lda,
                                  int rows_per_split) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r_end = (r_begin + rows_per_split < R) ? r_begin + rows_per_split : R;
  float s = 0.f;
  for (int64_t r = r_begin; r < r_end; ++r) s += A[r * lda + m];
  atomicAdd(&out[m], s);
}

}  // namespace paule

using namespace paule;

extern "C" int paule_gemm_tn_f32(const float* A, const float* B, float* C, int64_t R, int64_t M, int64_t N, int64_t lda,
                                 int64_t ldb, int accumulate, paule_stream_t stream) {
  PAULE_REQUIRE(A && B && C && R >= 0 && M > 0 && N > 0 && lda >= M && ldb >= N);
  cudaStream_t s = as_stream(stream);
  const int tiles = (int)(ceil_div(M, kTnTile) * ceil_div(N, kTnTile));
  // enough CTAs for the machine: split the reduction when the output has few tiles
  int splits = 1;
  if (R > 0 && tiles < 2 * sm_count()) splits = (int)((2 * sm_count() + tiles - 1) / tiles);
  const int64_t max_splits = ceil_div(R > 0 ? R : 1, 4 * kTnChunk);
  if (splits > max_splits) splits = (int)max_splits;
  if (splits < 1) splits = 1;
  const int atomic = (splits > 1 || accumulate) ? 1 : 0;
  if (atomic && !accumulate) PAULE_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * (size_t)N, s));
  if (R == 0) {
    if (!accumulate && !atomic) PAULE_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * (size_t)N, s));
    return PAULE_OK;
  }
  const int rows_per_split = (int)(ceil_div(ceil_div(R, splits), kTnChunk) * kTnChunk);
  dim3 grid((unsigned)ceil_div(M, kTnTile), (unsigned)ceil_div(N, kTnTile), (unsigned)ceil_div(R, rows_per_split));
  gemm_tn_f32_kernel<<<grid, 256, 0, s>>>(A, B, C, R, (int)M, (int)N, lda, ldb, rows_per_split, atomic);
  PAULE_LAUNCH_CHECK("gemm_tn_f32_kernel");
  return PAULE_OK;
}

extern "C" int paule_colsum_f32(const float* A, float* out, int64_t R, int64_t M, int64_t lda, int accumulate,
                                paule_stream_t stream) {
  PAULE_REQUIRE(A && out && R >= 0 && M > 0 && lda >= M);
  cudaStream_t s = as_stream(stream);
  if (!accumulate) PAULE_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)M, s));
  if (R == 0) return PAULE_OK;
  int splits = (int)ceil_div(R, 256);
  if (splits > 64) splits = 64;
  const int rows_per_split = (int)ceil_div(R, splits);
  dim3 grid((unsigned)ceil_div(M, 128), (unsigned)ceil_div(R, rows_per_split));
  colsum_f32_kernel<<<grid, 128, 0, s>>>(A, out, R, (int)M, lda, rows_per_split);
  PAULE_LAUNCH_CHECK("colsum_f32_kernel");
  return PAULE_OK;
}
