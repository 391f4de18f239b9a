// Persistent-RNN forward of one H=720 LSTM layer on tcgen05 (sm_100a).
//
//   grid  = 90 CTAs (one per SM), CTA q owns hidden units [8q, 8q+8): its 32 gate rows of W_hh (bf16, K padded to
//           768) stay resident in shared memory for the whole sequence as the B operand [N=32, K=768].
//   roles = warps 0-7 epilogue (cell), warp 8 producer (grid barrier + TMA bulk copies), warp 9 MMA issuer.
//   step t: the producer waits on the grid barrier and pulls h_{t-1} (all 64 batch rows x 768, bf16, 96 KB) from the
//           L2-resident exchange image with 12 bulk copies (one per 64-wide k-block, each landing on its own
//           mbarrier); the MMA thread issues 48 tcgen05.mma (M=64, N=32, K=16) as the k-blocks arrive, accumulating
//           [64 x 32] fp32 in TMEM, and commits to an mbarrier; the epilogue warps read TMEM, add the pre-computed
//           input projection x_t W_ih^T + b (prefetched before the wait), apply the cell, store h_t / c_t /
//           activated gates (fp32 stash for BPTT) and write h_t as bf16 straight into the UMMA image of the other
//           exchange buffer; a release-add on a global counter is the grid barrier for the next step.
//   Reference operator replaced: torch.nn.LSTM's recurrence (/root/reference/paule/models.py:349, :441).
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_lstm.cuh"

#ifdef PAULE_TC_TRACE
#define TRACE_DECL uint64_t tr_last = globaltimer_ns(); uint64_t tr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define TRACE(i) { const uint64_t _n = globaltimer_ns(); tr_acc[i] += _n - tr_last; tr_last = _n; }
#define TRACE_DUMP(base) { uint64_t* _o = reinterpret_cast<uint64_t*>(xchg + 64) + (base); for (int _i = 0; _i < 8; ++_i) _o[_i] = tr_acc[_i]; }
#else
#define TRACE_DECL
#define TRACE(i) {}
#define TRACE_DUMP(base) {}
#endif

#ifndef PAULE_FWD_ARRIVALS
#define PAULE_FWD_ARRIVALS 1   // grid-barrier arrivals per CTA and step: 1 (after a CTA barrier) or 8 (one per epilogue warp)
#endif

namespace paule {
namespace tc {

// ------------------------------------------------------------------------------------------------------------
// weight packing: fp32 torch layout -> bf16 UMMA images
// ------------------------------------------------------------------------------------------------------------
// forward image: slice q (48 KB): B operand [N=32, K=768]; row n = half*16 + gate*4 + uu  <->  W_hh[gate*H + 8q + half*4 + uu, :]
__global__ void pack_fwd_kernel(const float* __restrict__ w_hh, uint8_t* __restrict__ img) {
  const int q = blockIdx.x;
  for (int e = threadIdx.x; e < kFwdN * kKPad; e += blockDim.x) {
    const int n = e / kKPad, k = e % kKPad;
    const int half = n >> 4, gate = (n >> 2) & 3, uu = n & 3;
    const int row = gate * kH + q * kFwdUnits + half * 4 + uu;
    const float v = (k < kH) ? w_hh[(size_t)row * kH + k] : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(img + (size_t)q * kFwdSliceBytes + umma_offset(kFwdN, n, k)) = __float2bfloat16_rn(v);
  }
}

// backward image: slice (ug, g) (48 KB): B operand [N=32 units, K=768]; B[n, k] = W_hh[g*H + k, 32*ug + n]
__global__ void pack_bwd_kernel(const float* __restrict__ w_hh, uint8_t* __restrict__ img) {
  const int ug = blockIdx.x / 4, g = blockIdx.x % 4;
  for (int e = threadIdx.x; e < kBwdN * kKPad; e += blockDim.x) {
    const int k = e / kBwdN, n = e % kBwdN;   // n fastest: coalesced reads of a W_hh row
    const int j = ug * kBwdN + n;
    const float v = (k < kH && j < kH) ? w_hh[(size_t)(g * kH + k) * kH + j] : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(img + (size_t)blockIdx.x * kBwdSliceBytes + umma_offset(kBwdN, n, k)) =
        __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------------------------------
// forward kernel
// ------------------------------------------------------------------------------------------------------------
struct FwdSmem {
  uint8_t w[kFwdSliceBytes];            // B operand, resident        (48 KB, 1024-aligned)
  uint8_t a[kNumKB][kRows * 128];       // A operand k-blocks of h_{t-1} (12 x 8 KB)
  uint8_t a_slack[kRows * 128];         // an M=128 descriptor on the last k-block reads 64 rows past it (ignored rows)
  uint64_t full[kNumKB];                // k-block landed
  uint64_t mma_done;                    // accumulator ready
  uint64_t w_ready;
  uint32_t tmem_base;
};

constexpr int kEpiThreads = 256;
// One tcgen05.mma issue costs ~60 ns of the issuing thread regardless of tile size (measured), so a 48-instruction
// K loop from one thread is a 2.9 us serial chain.  Issue from 12 warps in parallel instead: warp 9+m owns k-block
// (m + q) % 12 and its own fp32 accumulator tile in TMEM; the epilogue sums the 12 tiles.
constexpr int kMmaWarps = 12;
constexpr int kFwdThreads = kEpiThreads + 32 + 32 * kMmaWarps;   // + producer warp + MMA warps

__global__ void __launch_bounds__(kFwdThreads, 1)
tc_lstm_fwd_kernel(float* __restrict__ gates, const uint8_t* __restrict__ packed, float* __restrict__ h_out,
                   float* __restrict__ c_out, uint8_t* __restrict__ xchg, uint8_t* __restrict__ img_seq, int T, int B,
                   int Bs) {
  extern __shared__ uint8_t smem_raw[];
  FwdSmem& S = *reinterpret_cast<FwdSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int q = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned int* counter = reinterpret_cast<unsigned int*>(xchg);
  volatile int* err = reinterpret_cast<volatile int*>(xchg + 4);
  // h_t images [12][64][128 B]: one per time step when the caller keeps them (they are the A operand of the next
  // layer's input-projection GEMM), otherwise two ping-pong images inside the exchange buffer
  uint8_t* hbuf = img_seq ? img_seq : xchg + kXchgHeader;
  const int img_mask = img_seq ? 0x7fffffff : 1;

  if (tid == 0) {
    for (int i = 0; i < kNumKB; ++i) mbar_init(&S.full[i], 1);
    mbar_init(&S.mma_done, kMmaWarps);
    mbar_init(&S.w_ready, 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<512>(&S.tmem_base);   // 12 tiles x 32 columns, rounded up to a power of two
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = S.tmem_base;

  if (warp == 8) {
    // ===================== producer: grid barrier + bulk copies of h_{t-1} =====================
    // lane l < 12 owns k-block (l + q) % 12: the 12 copies are issued in parallel, and the 90 CTAs start on
    // different k-blocks so that they do not all hit the same L2 lines at the same moment.
    if (lane == 0) {
      mbar_arrive_expect_tx(&S.w_ready, kFwdSliceBytes);      // resident weights: one 48 KB bulk copy
      bulk_g2s(S.w, packed + (size_t)q * kFwdSliceBytes, kFwdSliceBytes, &S.w_ready);
    }
    const int kb = (lane + q) % kNumKB;
    TRACE_DECL
    for (int t = 1; t < T; ++t) {
      // all 90 slices of h_{t-1} are in the image
      if (lane == 0) grid_wait(counter, (unsigned int)(t * gridDim.x * PAULE_FWD_ARRIVALS), err);
      __syncwarp();
      TRACE(0)
      if (lane < kNumKB) {
        fence_proxy_async();
        const uint8_t* src = hbuf + (size_t)((t - 1) & img_mask) * kXchgImageBytes;
        mbar_arrive_expect_tx(&S.full[kb], kRows * 128);
        bulk_g2s(S.a[kb], src + (size_t)kb * kRows * 128, kRows * 128, &S.full[kb]);
      }
      TRACE(1)
    }
    if (q == 0 && lane == 0) TRACE_DUMP(0)
    __syncwarp();
  } else if (warp >= 9) {
    // ===================== MMA issuers: warp 9+m takes k-blocks i = m, m+4, m+8 into accumulator tile m =====================
    const int mw = warp - 9;
    if (lane == 0) {
#ifndef EXP_M
#define EXP_M kRows
#endif
      const uint32_t idesc = make_idesc_bf16(EXP_M, kFwdN);
      mbar_wait(&S.w_ready, 0, err);
      TRACE_DECL
      for (int t = 1; t < T; ++t) {
        const uint32_t par = (uint32_t)((t - 1) & 1);
#ifdef PAULE_TC_TRACE_SPLIT
        for (int i = 0; i < kNumKB; ++i) {
          mbar_wait(&S.full[(i + q) % kNumKB], par, err);
          if (i == 0) TRACE(0)
        }
        TRACE(2)   // all k-blocks landed
#endif
#pragma unroll 1
        for (int i = mw; i < kNumKB; i += kMmaWarps) {
          const int kb = (i + q) % kNumKB;     // same rotation as the producer
          mbar_wait(&S.full[kb], par, err);
#ifndef PAULE_TC_TRACE_SPLIT
          if (i == 0) TRACE(0)
#endif
          tcgen05_fence_after();
          const uint64_t da = make_smem_desc_sw128(smem_u32(S.a[kb]));
          const uint64_t db = make_smem_desc_sw128(smem_u32(S.w + (size_t)kb * kFwdN * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem + (uint32_t)(mw * 32), da + 2 * k, db + 2 * k, idesc, ((i - mw) | k) ? 1u : 0u);
        }
        umma_commit(&S.mma_done);
        TRACE(1)
#ifdef PAULE_TC_TRACE_SPLIT
        mbar_wait(&S.mma_done, par, err);
        TRACE(3)   // MMA execution after the last issue
#endif
      }
      if (q == 0 && mw == 0) TRACE_DUMP(8)
    }
    __syncwarp();
  } else {
    // ===================== epilogue: the LSTM cell =====================
    // TMEM lane group (warp % 4) holds rows 16*(warp%4) .. +15 in its lanes 0..15 (UMMA M=64 layout);
    // warps 0-3 take columns 0..15 (units 0-3 of the slice), warps 4-7 columns 16..31 (units 4-7);
    // lanes 16..31 take over two of the four units from lane-16 (shuffle) so that all 32 lanes run the cell.
    const int rowgrp = warp & 3, half = warp >> 2;
    const int row = rowgrp * 16 + (lane & 15);
    const int upair = lane >> 4;                           // 0: units 0,1 of the half; 1: units 2,3
    const int j = q * kFwdUnits + half * 4 + upair * 2;    // first of this thread's two hidden units
    const bool valid = row < B;
    const uint32_t taddr = tmem + ((uint32_t)(rowgrp * 32) << 16) + (uint32_t)(half * 16);
    const size_t xo = umma_offset(kRows, row, j);          // position of (row, j) in an exchange image
    float c_prev[2] = {0.f, 0.f};
    TRACE_DECL

    for (int t = 0; t < T; ++t) {
      // input projection of this step (independent of h_{t-1}): issue the loads before waiting
      float2 xp[4];
      float* grow = gates + ((size_t)t * Bs + (valid ? row : 0)) * (4 * kH);
#pragma unroll
      for (int g = 0; g < 4; ++g)
        xp[g] = valid ? *reinterpret_cast<const float2*>(grow + g * kH + j) : make_float2(0.f, 0.f);

      float acc[16];
      if (t > 0) {
        mbar_wait(&S.mma_done, (uint32_t)((t - 1) & 1), err);
        TRACE(0)
        tcgen05_fence_after();
        tmem_ld_sum_x16<kMmaWarps>(taddr, acc);   // the K range is split over kMmaWarps accumulator tiles
        tcgen05_fence_before();
        TRACE(1)
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.f;
      }
      // columns of this half: gate*4 + uu.  Lanes 16..31 fetch units 2,3 from lane-16.
      float pre[4][2];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float a0 = acc[g * 4 + 0], a1 = acc[g * 4 + 1], a2 = acc[g * 4 + 2], a3 = acc[g * 4 + 3];
        const float s2 = __shfl_sync(0xffffffffu, a2, lane & 15), s3 = __shfl_sync(0xffffffffu, a3, lane & 15);
        pre[g][0] = (upair ? s2 : a0) + xp[g].x;
        pre[g][1] = (upair ? s3 : a1) + xp[g].y;
      }
      float hv[2], gi[2], gf[2], gg[2], go[2], cn[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        gi[u] = fast_sigmoid(fminf(fmaxf(pre[0][u], -30.f), 30.f));
        gf[u] = fast_sigmoid(fminf(fmaxf(pre[1][u], -30.f), 30.f));
        gg[u] = fast_tanh(fminf(fmaxf(pre[2][u], -15.f), 15.f));
        go[u] = fast_sigmoid(fminf(fmaxf(pre[3][u], -30.f), 30.f));
        cn[u] = gf[u] * c_prev[u] + gi[u] * gg[u];
        hv[u] = go[u] * fast_tanh(fminf(fmaxf(cn[u], -15.f), 15.f));
        c_prev[u] = cn[u];
      }
      if (valid) {
        // h_t as bf16 into the UMMA image the next step bulk-copies: first, it is on the critical path
        *reinterpret_cast<__nv_bfloat162*>(hbuf + (size_t)(t & img_mask) * kXchgImageBytes + xo) =
            __floats2bfloat162_rn(hv[0], hv[1]);
      }
      TRACE(2)
      if (t + 1 < T) {
#if PAULE_FWD_ARRIVALS == 8
        // publish per warp: no CTA-wide barrier, each warp releases its own 32 rows x 2 units as soon as they are out
        __syncwarp();
        if (lane == 0) {
#ifndef PAULE_NO_WRITER_PROXY_FENCE
          fence_proxy_async();
#endif
          grid_arrive(counter);          // red.release.gpu: cumulative over the warp's stores ordered by __syncwarp
        }
#else
        // publish: CTA-wide barrier, then ONE gpu-scope release (cumulative over the CTA's stores)
        named_bar_sync(1, kEpiThreads);
        if (tid == 0) {
#ifndef PAULE_NO_WRITER_PROXY_FENCE
          fence_proxy_async();
#endif
          grid_arrive(counter);          // red.release.gpu: cumulative over the stores ordered by the barrier above
        }
#endif
      }
      TRACE(3)
      if (valid) {   // stash + outputs: off the critical path (the next step's barrier is already signalled)
        *reinterpret_cast<float2*>(grow + 0 * kH + j) = make_float2(gi[0], gi[1]);
        *reinterpret_cast<float2*>(grow + 1 * kH + j) = make_float2(gf[0], gf[1]);
        *reinterpret_cast<float2*>(grow + 2 * kH + j) = make_float2(gg[0], gg[1]);
        *reinterpret_cast<float2*>(grow + 3 * kH + j) = make_float2(go[0], go[1]);
        const size_t o = ((size_t)t * Bs + row) * kH + j;
        *reinterpret_cast<float2*>(c_out + o) = make_float2(cn[0], cn[1]);
        *reinterpret_cast<float2*>(h_out + o) = make_float2(hv[0], hv[1]);
      }
      TRACE(4)
    }
    if (q == 0 && tid == 0) TRACE_DUMP(16)
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem);
}

}  // namespace tc
}  // namespace paule

using namespace paule;
using namespace paule::tc;

extern "C" size_t paule_tc_packed_lstm_bytes(int64_t H, int64_t I) {
  (void)I;
  if (H != kH) return 0;
  return (size_t)kFwdCtas * kFwdSliceBytes + (size_t)kBwdCtas * kBwdSliceBytes;
}

extern "C" int paule_tc_pack_lstm(const float* w_ih, const float* w_hh, void* packed, int64_t H, int64_t I,
                                  paule_stream_t stream) {
  (void)w_ih; (void)I;
  PAULE_REQUIRE(w_hh && packed);
  if (H != kH) return PAULE_ERR_UNSUPPORTED;
  uint8_t* img = reinterpret_cast<uint8_t*>(packed);
  PAULE_REQUIRE(reinterpret_cast<uintptr_t>(img) % 16 == 0);
  pack_fwd_kernel<<<kFwdCtas, 256, 0, as_stream(stream)>>>(w_hh, img);
  pack_bwd_kernel<<<kBwdCtas, 256, 0, as_stream(stream)>>>(w_hh, img + (size_t)kFwdCtas * kFwdSliceBytes);
  PAULE_LAUNCH_CHECK("pack kernels");
  return PAULE_OK;
}

extern "C" size_t paule_tc_rnn_xchg_bytes(int64_t B) {
  (void)B;
  // header (counter, error flag) + forward: 2 h images; backward: 2 x 4 gate images (sized for the larger user)
  return (size_t)kXchgHeader + (size_t)8 * kXchgImageBytes;
}

extern "C" size_t paule_tc_img_seq_bytes(int64_t T, int64_t B, int64_t images_per_step) {
  if (T <= 0 || B <= 0 || images_per_step <= 0) return 0;
  return (size_t)((B + kRows - 1) / kRows) * (size_t)T * (size_t)images_per_step * kXchgImageBytes;
}

extern "C" int paule_tc_lstm_seq_fwd(float* gates, const void* packed, float* h, float* c, void* xchg, void* h_img_seq,
                                     int64_t T, int64_t B, int math, paule_stream_t stream) {
  PAULE_REQUIRE(gates && packed && h && c && xchg && T >= 0 && B > 0);
  PAULE_REQUIRE(math == PAULE_MATH_BF16);
  if (T == 0) return PAULE_OK;
  PAULE_REQUIRE(reinterpret_cast<uintptr_t>(xchg) % 16 == 0);   // bulk copies need 16-byte aligned global addresses
  cudaStream_t s = as_stream(stream);
  static bool attr_set = false;
  const int smem = (int)sizeof(FwdSmem) + 1024;
  if (!attr_set) {
    PAULE_CUDA(cudaFuncSetAttribute(tc_lstm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  // words are independent: batches larger than the UMMA M tile run as consecutive 64-word groups
  for (int64_t r0 = 0; r0 < B; r0 += kRows) {
    // zero the barrier words and both exchange images (pad columns / pad rows must be exact zeros)
    PAULE_CUDA(cudaMemsetAsync(xchg, 0, (size_t)kXchgHeader + (size_t)2 * kXchgImageBytes, s));
    int Ti = (int)T, Bi = (int)((B - r0 < kRows) ? (B - r0) : kRows), Bsi = (int)B;
    float* gp = gates + r0 * 4 * kH;
    float* hp = h + r0 * kH;
    float* cp = c + r0 * kH;
    uint8_t* xc = reinterpret_cast<uint8_t*>(xchg);
    uint8_t* is = h_img_seq ? reinterpret_cast<uint8_t*>(h_img_seq) + (size_t)(r0 / kRows) * (size_t)T * kXchgImageBytes
                            : nullptr;
    const uint8_t* pk = reinterpret_cast<const uint8_t*>(packed);
    void* args[] = {&gp, &pk, &hp, &cp, &xc, &is, &Ti, &Bi, &Bsi};
    PAULE_CUDA(cudaLaunchCooperativeKernel((void*)tc_lstm_fwd_kernel, dim3(kFwdCtas), dim3(kFwdThreads), args,
                                           (size_t)smem, s));
  }
  return PAULE_OK;
}
