// Persistent reverse-time BPTT (input gradients only) of one H=720 LSTM layer on tcgen05 (sm_100a).
//
//   dh_{t}[b,j] = dh_ext_t[b,j] + sum_{k<2880} da_{t+1}[b,k] W_hh[k,j];   da_t = cell adjoint(dh_t, dc, stash_t)
//
//   The recurrent GEMM has K = 4H = 2880.  Streaming all of da_{t+1} (368 KB bf16) through every CTA would be
//   L2- and shared-memory-bound, so K is split over the four gates inside a 4-CTA thread-block cluster:
//     grid = 23 unit groups (32 hidden units) x 4 gates = 92 CTAs, cluster (4,1,1); CTA (ug, g) keeps
//     B = W_hh[g*720 + k, 32ug + n] ([N=32, K=768] bf16) resident in shared memory and, per step, bulk-copies only
//     gate g's image of da_{t+1} (96 KB), runs 48 tcgen05.mma (M=64, N=32, K=16; issued by 12 warps in parallel,
//     one k-block and one TMEM accumulator tile each) and pushes its partial [64 x 32] to its three siblings through
//     distributed shared memory (each sibling finalises 8 of the 32 units); after one cluster barrier every CTA
//     sums the four partials for its 8 units, applies the cell adjoint (dc lives in registers for the whole
//     sequence), writes da_t as bf16 into the four gate images of the exchange buffer, signals the grid barrier and
//     only then overwrites the fp32 stash with da_t (the operand of the dX GEMM).
//   roles = warps 0-7 cell adjoint, warp 8 producer (grid barrier + TMA bulk copies), warps 9-20 MMA issuers.
//   Reference operator replaced: autograd through aten::lstm (discrepancy.backward(), paule/paule.py:1052).
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_lstm.cuh"

namespace paule {
namespace tc {

constexpr int kBwdEpiThreads = 256;
constexpr int kBwdMmaWarps = 12;
constexpr int kBwdThreads = kBwdEpiThreads + 32 + 32 * kBwdMmaWarps;

struct BwdSmem {
  uint8_t w[kBwdSliceBytes];            // B operand, resident (48 KB)
  uint8_t a[kNumKB][kRows * 128];       // A operand: gate g's image of da_{t+1} (96 KB)
  float red[4][kRows][8];               // partial sums from the 4 gate CTAs for this CTA's 8 units (8 KB)
  uint64_t full[2][kNumKB];             // two sets so that the next step's barrier is armed early
  uint64_t mma_done;
  uint64_t w_ready;
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kBwdThreads, 1)
tc_lstm_bwd_kernel(float* __restrict__ gates, const float* __restrict__ c_seq, const uint8_t* __restrict__ packed_bwd,
                   const float* __restrict__ dh_seq, int dh_mode, const float* __restrict__ dh_last,
                   uint8_t* __restrict__ xchg, uint8_t* __restrict__ img_seq, int T, int B, int Bs) {
  extern __shared__ uint8_t smem_raw[];
  BwdSmem& S = *reinterpret_cast<BwdSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = (int)cluster_ctarank();            // gate handled by this CTA's K slice
  const int ug = blockIdx.x >> 2;                  // unit group: hidden units [32ug, 32ug+32)
  unsigned int* counters = reinterpret_cast<unsigned int*>(xchg);      // counters[32 * kb]: one 128-byte line per k-block
  volatile int* err = reinterpret_cast<volatile int*>(xchg + kXchgErrOff);
  // dA_t images, four (one per gate) per step: kept for every step when the caller wants them (A operand of the dX
  // GEMM), otherwise two ping-pong sets inside the exchange buffer
  uint8_t* img = img_seq ? img_seq : xchg + kXchgHeader;
  const int img_mask = img_seq ? 0x7fffffff : 1;

  if (tid == 0) {
    for (int i = 0; i < kNumKB; ++i) { mbar_init(&S.full[0][i], 1); mbar_init(&S.full[1][i], 1); }
    mbar_init(&S.mma_done, kBwdMmaWarps);
    mbar_init(&S.w_ready, 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<32>(&S.tmem_base);     // ONE accumulator tile that all 12 issuers add into
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = S.tmem_base;
  cluster_sync_all();   // every CTA of the cluster is resident before any DSMEM traffic

  if (warp == 8) {
    // ===================== producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(&S.w_ready, kBwdSliceBytes);
      bulk_g2s(S.w, packed_bwd + (size_t)blockIdx.x * kBwdSliceBytes, kBwdSliceBytes, &S.w_ready);
    }
    const int kb = lane;
    // k-block kb of every gate image is written by the CTAs of unit groups 2kb and 2kb+1 (the last one by group 22 only)
    const unsigned int owners = (unsigned int)((2 * kb + 1 < kBwdGroups) ? 8 : 4);
    if (lane < kNumKB && T > 1) mbar_arrive_expect_tx(&S.full[0][kb], kRows * 128);   // armed one step ahead of the copy
    for (int it = 1; it < T; ++it) {
      const int t = T - 1 - it;
      if (lane < kNumKB) {
        grid_wait(counters + 32 * kb, (unsigned int)it * owners, err);   // k-block kb of da_{t+1} is complete
        fence_proxy_async_global();                                            // generic-proxy writes -> async-proxy read
        const uint8_t* src = img + (size_t)(((t + 1) & img_mask) * 4 + g) * kXchgImageBytes;
        bulk_g2s(S.a[kb], src + (size_t)kb * kRows * 128, kRows * 128, &S.full[(it - 1) & 1][kb]);
        if (it + 1 < T) mbar_arrive_expect_tx(&S.full[it & 1][kb], kRows * 128);   // next step's barrier (other set)
      }
      __syncwarp();
      cluster_sync_all();
    }
  } else if (warp >= 9) {
    // ===================== MMA issuers: warp 9+kb adds k-block kb into the shared accumulator tile =====================
    const int kb = warp - 9;
    const uint32_t idesc = make_idesc_bf16(kRows, kBwdN);
    if (lane == 0) mbar_wait(&S.w_ready, 0, err);
    __syncwarp();
    for (int it = 1; it < T; ++it) {
      if (lane == 0) {
        mbar_wait(&S.full[(it - 1) & 1][kb], (uint32_t)(((it - 1) >> 1) & 1), err);
        tcgen05_fence_after();
        const uint64_t da = make_smem_desc_sw128(smem_u32(S.a[kb]));
        const uint64_t db = make_smem_desc_sw128(smem_u32(S.w + (size_t)kb * kBwdN * 128));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, 1u);   // the tile was pre-zeroed
        umma_commit(&S.mma_done);
      }
      __syncwarp();
      cluster_sync_all();
    }
  } else {
    // ===================== cell adjoint =====================
    // TMEM read role: lane group (warp%4) = rows 16*(warp%4)..+15 in lanes 0..15; column half (warp/4) = 16 units
    const int rowgrp = warp & 3, hf = warp >> 2;
    const int prow = rowgrp * 16 + (lane & 15);
    const uint32_t taddr = tmem + ((uint32_t)(rowgrp * 32) << 16) + (uint32_t)(hf * 16);
    // remote destinations of this thread's two 8-unit groups: sibling CTAs 2hf and 2hf+1, slot [g][prow][0..7]
    const uint32_t red_local = smem_u32(&S.red[g][prow][0]);
    const uint32_t red_dst0 = mapa_shared(red_local, (uint32_t)(2 * hf));
    const uint32_t red_dst1 = mapa_shared(red_local, (uint32_t)(2 * hf + 1));
    // cell role: row = tid/4, unit pair = tid%4 of this CTA's 8 units
    const int row = tid >> 2, up = tid & 3;
    const int j = ug * kBwdN + g * 8 + up * 2;
    const bool valid = (row < B) && (j < kH);
    const size_t xo = umma_offset(kRows, row, j);
    float dc[2] = {0.f, 0.f};
    unsigned int* my_counter = counters + 32 * (ug >> 1);   // this CTA's units live in k-block ug / 2
    tmem_zero_x16(taddr);
    tmem_st_wait();
    tcgen05_fence_before();

    for (int it = 0; it < T; ++it) {
      const int t = T - 1 - it;
      // (1) everything that does not depend on da_{t+1}: stash, cell states, external gradient
      float2 s_i, s_f, s_g, s_o, ct, cp, dh;
      s_i = s_f = s_g = s_o = ct = cp = dh = make_float2(0.f, 0.f);
      float* grow = gates + ((size_t)t * Bs + (valid ? row : 0)) * (4 * kH);
      if (valid) {
        s_i = *reinterpret_cast<const float2*>(grow + 0 * kH + j);
        s_f = *reinterpret_cast<const float2*>(grow + 1 * kH + j);
        s_g = *reinterpret_cast<const float2*>(grow + 2 * kH + j);
        s_o = *reinterpret_cast<const float2*>(grow + 3 * kH + j);
        ct = *reinterpret_cast<const float2*>(c_seq + ((size_t)t * Bs + row) * kH + j);
        if (t > 0) cp = *reinterpret_cast<const float2*>(c_seq + ((size_t)(t - 1) * Bs + row) * kH + j);
        if (dh_mode == 1) {
          dh = *reinterpret_cast<const float2*>(dh_seq + ((size_t)t * Bs + row) * kH + j);
        } else if (dh_mode == 2 && (t >> 1) < (T >> 1)) {
          const float2 v = *reinterpret_cast<const float2*>(dh_seq + ((size_t)(t >> 1) * Bs + row) * kH + j);
          dh = make_float2(0.5f * v.x, 0.5f * v.y);
        }
        if (dh_last != nullptr && t == T - 1) {
          const float2 v = *reinterpret_cast<const float2*>(dh_last + (size_t)row * kH + j);
          dh.x += v.x; dh.y += v.y;
        }
      }
      if (it > 0) {
        mbar_wait(&S.mma_done, (uint32_t)((it - 1) & 1), err);
        tcgen05_fence_after();
        float p[16];
        tmem_ld_x16(taddr, p);
        tmem_zero_x16(taddr);   // re-arm the accumulator: every MMA of the next step adds into it
        tmem_st_wait();
        if (lane < 16) {   // push the partial sums to the siblings that finalise these units
          st_cluster_v4(red_dst0, make_float4(p[0], p[1], p[2], p[3]));
          st_cluster_v4(red_dst0 + 16, make_float4(p[4], p[5], p[6], p[7]));
          st_cluster_v4(red_dst1, make_float4(p[8], p[9], p[10], p[11]));
          st_cluster_v4(red_dst1 + 16, make_float4(p[12], p[13], p[14], p[15]));
        }
        tcgen05_fence_before();
        cluster_sync_all();
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const float2 v = *reinterpret_cast<const float2*>(&S.red[s][row][up * 2]);
          dh.x += v.x; dh.y += v.y;
        }
      }
      // (2) cell adjoint (oracle: manual_lstm_backward_input)
      float2 d_i, d_f, d_g, d_o;
      {
        const float tc0 = fast_tanh(fminf(fmaxf(ct.x, -15.f), 15.f)), tc1 = fast_tanh(fminf(fmaxf(ct.y, -15.f), 15.f));
        const float do0 = dh.x * tc0, do1 = dh.y * tc1;
        const float dc0 = dc[0] + dh.x * s_o.x * (1.f - tc0 * tc0), dc1 = dc[1] + dh.y * s_o.y * (1.f - tc1 * tc1);
        d_i = make_float2(dc0 * s_g.x * s_i.x * (1.f - s_i.x), dc1 * s_g.y * s_i.y * (1.f - s_i.y));
        d_f = make_float2(dc0 * cp.x * s_f.x * (1.f - s_f.x), dc1 * cp.y * s_f.y * (1.f - s_f.y));
        d_g = make_float2(dc0 * s_i.x * (1.f - s_g.x * s_g.x), dc1 * s_i.y * (1.f - s_g.y * s_g.y));
        d_o = make_float2(do0 * s_o.x * (1.f - s_o.x), do1 * s_o.y * (1.f - s_o.y));
        dc[0] = dc0 * s_f.x;
        dc[1] = dc1 * s_f.y;
      }
      if (t > 0 || img_seq != nullptr) {
        if (valid) {   // bf16 da_t into the four gate images: on the critical path of the next step
          uint8_t* dst = img + (size_t)((t & img_mask) * 4) * kXchgImageBytes + xo;
          *reinterpret_cast<__nv_bfloat162*>(dst + 0 * (size_t)kXchgImageBytes) = __floats2bfloat162_rn(d_i.x, d_i.y);
          *reinterpret_cast<__nv_bfloat162*>(dst + 1 * (size_t)kXchgImageBytes) = __floats2bfloat162_rn(d_f.x, d_f.y);
          *reinterpret_cast<__nv_bfloat162*>(dst + 2 * (size_t)kXchgImageBytes) = __floats2bfloat162_rn(d_g.x, d_g.y);
          *reinterpret_cast<__nv_bfloat162*>(dst + 3 * (size_t)kXchgImageBytes) = __floats2bfloat162_rn(d_o.x, d_o.y);
        }
      }
      if (t > 0) {
        named_bar_sync(1, kBwdEpiThreads);
        if (tid == 0) grid_arrive(my_counter);   // release-add, cumulative over the stores the barrier ordered
      }
      if (valid) {   // fp32 da_t over the stash (operand of the dX GEMM): off the critical path
        *reinterpret_cast<float2*>(grow + 0 * kH + j) = d_i;
        *reinterpret_cast<float2*>(grow + 1 * kH + j) = d_f;
        *reinterpret_cast<float2*>(grow + 2 * kH + j) = d_g;
        *reinterpret_cast<float2*>(grow + 3 * kH + j) = d_o;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<32>(tmem);
  cluster_sync_all();   // no CTA exits while a sibling may still address its shared memory
}

}  // namespace tc
}  // namespace paule

using namespace paule;
using namespace paule::tc;

static int seq_bwd(float* gates, const float* c, const void* packed, const float* dh_seq, int dh_mode, const float* dh_last,
                   void* xchg, void* da_img_seq, int64_t T, int64_t B, int math, int keep_da, paule_stream_t stream);

extern "C" int paule_tc_lstm_seq_bwd(float* gates, const float* c, const void* packed, const float* dh_seq,
                                     int dh_mode, const float* dh_last, void* xchg, void* da_img_seq, int64_t T,
                                     int64_t B, int math, paule_stream_t stream) {
  return seq_bwd(gates, c, packed, dh_seq, dh_mode, dh_last, xchg, da_img_seq, T, B, math, 1, stream);
}

// same, but the fp32 d(pre-activation) is NOT written over `gates`: only the bf16 images in da_img_seq (required) are produced
extern "C" int paule_tc_lstm_seq_bwd_img(float* gates, const float* c, const void* packed, const float* dh_seq,
                                         int dh_mode, const float* dh_last, void* xchg, void* da_img_seq, int64_t T,
                                         int64_t B, int math, paule_stream_t stream) {
  PAULE_REQUIRE(da_img_seq != nullptr);
  return seq_bwd(gates, c, packed, dh_seq, dh_mode, dh_last, xchg, da_img_seq, T, B, math, 0, stream);
}

static int seq_bwd(float* gates, const float* c, const void* packed, const float* dh_seq, int dh_mode, const float* dh_last,
                   void* xchg, void* da_img_seq, int64_t T, int64_t B, int math, int keep_da, paule_stream_t stream) {
  PAULE_REQUIRE(gates && c && packed && xchg && T >= 0 && B > 0);
  PAULE_REQUIRE(dh_mode == 0 || ((dh_mode == 1 || dh_mode == 2) && dh_seq));
  PAULE_REQUIRE(math == PAULE_MATH_BF16);
  if (T == 0) return PAULE_OK;
  PAULE_REQUIRE(reinterpret_cast<uintptr_t>(xchg) % 16 == 0);   // bulk copies need 16-byte aligned global addresses
  cudaStream_t s = as_stream(stream);
  if (!use_v1_bwd()) return lstm_seq_bwd2(gates, c, packed, dh_seq, dh_mode, dh_last, xchg, da_img_seq, T, B, keep_da, s);
  static bool attr_set = false;
  const int smem = (int)sizeof(BwdSmem) + 1024;
  if (!attr_set) {
    PAULE_CUDA(cudaFuncSetAttribute(tc_lstm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const uint8_t* pk = reinterpret_cast<const uint8_t*>(packed) + (size_t)kFwdCtas * kFwdSliceBytes;
  // cooperative + cluster launch accepted by this driver?  (Nsight Compute rejects the combination with LaunchFailed;
  // PAULE_NO_COOP_CLUSTER=1 launches with the cluster attribute only -- 92 CTAs always fit the 148 SMs.)
  static int coop_ok = getenv("PAULE_NO_COOP_CLUSTER") ? 0 : 1;
  // words are independent: batches larger than the UMMA M tile run as consecutive 64-word groups
  for (int64_t r0 = 0; r0 < B; r0 += kRows) {
  PAULE_CUDA(cudaMemsetAsync(xchg, 0, (size_t)kXchgHeader + (size_t)8 * kXchgImageBytes, s));
  const int Bi = (int)((B - r0 < kRows) ? (B - r0) : kRows);
  float* gp = gates + r0 * 4 * kH;
  const float* cp = c + r0 * kH;
  const float* dsp = dh_seq ? dh_seq + r0 * kH : nullptr;
  const float* dlp = dh_last ? dh_last + r0 * kH : nullptr;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kBwdCtas);
  cfg.blockDim = dim3(kBwdThreads);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = s;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = 4;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeCooperative;
  attrs[1].val.cooperative = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = coop_ok ? 2 : 1;
  uint8_t* xc = reinterpret_cast<uint8_t*>(xchg);
  uint8_t* is = da_img_seq ? reinterpret_cast<uint8_t*>(da_img_seq) + (size_t)(r0 / kRows) * (size_t)T * 4 * kXchgImageBytes
                           : nullptr;
  cudaError_t e = cudaLaunchKernelEx(&cfg, tc_lstm_bwd_kernel, gp, cp, pk, dsp, dh_mode, dlp, xc, is, (int)T, Bi, (int)B);
  if (e != cudaSuccess && coop_ok) {
    // some driver/toolkit combinations reject cooperative + cluster launches: 92 CTAs (23 clusters) fit the 148 SMs
    // at one CTA per SM, so all CTAs are co-resident on an otherwise idle stream; the watchdog guards the rest.
    cudaGetLastError();
    coop_ok = 0;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, tc_lstm_bwd_kernel, gp, cp, pk, dsp, dh_mode, dlp, xc, is, (int)T, Bi, (int)B);
  }
  PAULE_CUDA(e);
  }
  return PAULE_OK;
}
