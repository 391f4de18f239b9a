// Microbenchmark: issue / completion cost of short tcgen05.mma chains (kind::f16, bf16) with the A operand in shared
// memory (SS) or in tensor memory (TS), plus a numerical check of the TS operand layout.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_bench tools/mma_bench.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../paule_b200/csrc/tc_common.cuh"

using namespace paule::tc;

struct Smem {
  uint8_t a[12][128 * 128];   // A k-blocks [128 rows][128 B] (192 KB)
  uint8_t b[64 * 128];        // one B k-block, up to 64 rows (8 KB)
  uint64_t bar[4];
  uint32_t tmem_base;
};

// out: [0..M*N) SS result, [M*N .. 2*M*N) TS result (K = 64); timing: cyc[]
__global__ void __launch_bounds__(512, 1) bench_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ Bm,
                                                       float* __restrict__ out, long long* __restrict__ cyc, int M, int N) {
  extern __shared__ uint8_t raw[];
  Smem& S = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  // operands: A [M,64] and B [N,64] (K-major) into every k-block slot (same data 12 times)
  for (int e = tid; e < 128 * 64; e += blockDim.x) {
    const int r = e / 64, k = e % 64;
    const __nv_bfloat16 v = r < M ? A[r * 64 + k] : __float2bfloat16(0.f);
    for (int kb = 0; kb < 12; ++kb) *reinterpret_cast<__nv_bfloat16*>(S.a[kb] + umma_offset(128, r, k)) = v;
  }
  for (int e = tid; e < 64 * 64; e += blockDim.x) {
    const int r = e / 64, k = e % 64;
    *reinterpret_cast<__nv_bfloat16*>(S.b + umma_offset(64, r, k)) = r < N ? Bm[r * 64 + k] : __float2bfloat16(0.f);
  }
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&S.bar[i], i == 3 ? 12 : 1);
    fence_mbar_init();
  }
  fence_proxy_async_shared();
  if (warp == 0) tmem_alloc<512>(&S.tmem_base);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = S.tmem_base;
  volatile int err = 0;
  // A into TMEM columns 128..: lane = row m, column c holds k = 2c (low half) and 2c+1 (high half); 32 columns per k-block,
  // 12 k-blocks -> 384 columns
  if (warp < 4) {
    const int m = warp * 32 + lane;
    for (int kb = 0; kb < 12; ++kb)
      for (int c0 = 0; c0 < 32; c0 += 8) {
        uint32_t r[8];
        for (int i = 0; i < 8; ++i) {
          const int k = 2 * (c0 + i);
          const __nv_bfloat16 lo = m < M ? A[m * 64 + k] : __float2bfloat16(0.f), hi = m < M ? A[m * 64 + k + 1] : __float2bfloat16(0.f);
          r[i] = (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
        }
        tmem_st_x8(tmem + ((uint32_t)(warp * 32) << 16) + 128 + kb * 32 + c0, make_uint4(r[0], r[1], r[2], r[3]),
                   make_uint4(r[4], r[5], r[6], r[7]));
      }
    tmem_st_wait();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t idesc = make_idesc_bf16(M, N);
  const uint64_t db = make_smem_desc_sw128(smem_u32(S.b));
  uint32_t ph[4] = {0, 0, 0, 0};
  // ---- numerical check: D_ss -> columns 0.., D_ts -> columns 64..
  if (warp == 0) {
    if (elect_one_sync()) {
      const uint64_t da = make_smem_desc_sw128(smem_u32(S.a[0]));
      for (int k = 0; k < 4; ++k) umma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, k > 0);
      for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem + 64, tmem + 128 + 8 * k, db + 2 * k, idesc, k > 0);
      umma_commit(&S.bar[0]);
    }
    __syncwarp();
  }
  mbar_wait(&S.bar[0], ph[0], (volatile int*)&err); ph[0] ^= 1;
  tcgen05_fence_after();
  if (warp < 4) {
    for (int half = 0; half < 2; ++half)
      for (int c0 = 0; c0 < N; c0 += 8) {
        float v[8];
        tmem_ld_x8(tmem + ((uint32_t)(warp * 32) << 16) + half * 64 + c0, v);
        const int m = warp * 32 + lane;
        if (m < M) for (int i = 0; i < 8; ++i) out[(size_t)half * 128 * 64 + m * 64 + c0 + i] = v[i];
      }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  // ---- timing: one warp, chains of n MMAs (same accumulator), SS and TS
  int slot = 0;
  for (int mode = 0; mode < 2; ++mode)
    for (int n : {4, 12, 48}) {
      long long t0 = 0, t1 = 0, t2 = 0;
      if (warp == 0) {
        __syncwarp();
        t0 = clock64();
        if (elect_one_sync()) {
          for (int i = 0; i < n; ++i) {
            const int kb = (i >> 2) % 12, k = i & 3;
            if (mode == 0) umma_bf16(tmem, make_smem_desc_sw128(smem_u32(S.a[kb])) + 2 * k, db + 2 * k, idesc, 1u);
            else umma_bf16_ts(tmem, tmem + 128 + kb * 32 + 8 * k, db + 2 * k, idesc, 1u);
          }
          umma_commit(&S.bar[1]);
        }
        __syncwarp();
        t1 = clock64();
        mbar_wait(&S.bar[1], ph[1], (volatile int*)&err);
        t2 = clock64();
        if (lane == 0) { cyc[slot * 2] = t1 - t0; cyc[slot * 2 + 1] = t2 - t0; }
      }
      ph[1] ^= 1;
      ++slot;
      __syncthreads();
    }
  // ---- timing: 12 warps x 4 MMAs into one accumulator (the recurrent step), SS and TS
  for (int mode = 0; mode < 2; ++mode) {
    __syncthreads();
    const long long t0 = clock64();
    if (warp >= 4 && warp < 16) {
      const int kb = warp - 4;
      if (elect_one_sync()) {
        for (int k = 0; k < 4; ++k) {
          if (mode == 0) umma_bf16(tmem, make_smem_desc_sw128(smem_u32(S.a[kb])) + 2 * k, db + 2 * k, idesc, 1u);
          else umma_bf16_ts(tmem, tmem + 128 + kb * 32 + 8 * k, db + 2 * k, idesc, 1u);
        }
        umma_commit(&S.bar[3]);
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    mbar_wait(&S.bar[3], ph[3], (volatile int*)&err);
    const long long t2 = clock64();
    ph[3] ^= 1;
    if (tid == 4 * 32) { cyc[slot * 2] = t1 - t0; cyc[slot * 2 + 1] = t2 - t0; }
    ++slot;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  const int shapes[][2] = {{128, 16}, {128, 32}, {128, 64}, {64, 16}, {64, 32}, {64, 64}};
  std::vector<__nv_bfloat16> hA(128 * 64), hB(64 * 64);
  srand(1);
  for (auto& v : hA) v = __float2bfloat16((rand() % 255 - 127) / 128.f);
  for (auto& v : hB) v = __float2bfloat16((rand() % 255 - 127) / 128.f);
  __nv_bfloat16 *dA, *dB; float* dOut; long long* dCyc;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dOut, 2 * 128 * 64 * 4); cudaMalloc(&dCyc, 64 * 8);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  const int smem = (int)sizeof(Smem) + 1024;
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (auto& sh : shapes) {
    const int M = sh[0], N = sh[1];
    cudaMemset(dOut, 0, 2 * 128 * 64 * 4); cudaMemset(dCyc, 0, 64 * 8);
    for (int rep = 0; rep < 2; ++rep) bench_kernel<<<1, 512, smem>>>(dA, dB, dOut, dCyc, M, N);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("M=%d N=%d: %s\n", M, N, cudaGetErrorString(e)); return 1; }
    std::vector<float> out(2 * 128 * 64); std::vector<long long> cyc(64);
    cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(cyc.data(), dCyc, 64 * 8, cudaMemcpyDeviceToHost);
    double e_ss = 0, e_ts = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < 64; ++k) ref += (double)__bfloat162float(hA[m * 64 + k]) * __bfloat162float(hB[n * 64 + k]);
        e_ss = fmax(e_ss, fabs(out[m * 64 + n] - ref));
        e_ts = fmax(e_ts, fabs(out[128 * 64 + m * 64 + n] - ref));
      }
    printf("M=%3d N=%2d  max err SS %.2e  TS %.2e |", M, N, e_ss, e_ts);
    const char* names[] = {"SS n=4", "SS n=12", "SS n=48", "TS n=4", "TS n=12", "TS n=48", "SS 12x4", "TS 12x4"};
    for (int i = 0; i < 8; ++i) printf("  %s: issue %lld done %lld", names[i], cyc[2 * i], cyc[2 * i + 1]);
    printf("\n");
  }
  return 0;
}
