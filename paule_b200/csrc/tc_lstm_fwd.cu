// Persistent-RNN forward of one H=720 LSTM layer on tcgen05 (sm_100a).
//
//   grid  = 90 CTAs (one per SM, cooperative launch), CTA q owns hidden units [8q, 8q+8): its 32 gate rows of W_hh
//           (bf16, K padded to 768) stay resident in shared memory for the whole sequence as the B operand
//           [N=32, K=768].
//   roles = warps 0-7 epilogue (the LSTM cell), warp 8 producer (barrier polling + TMA bulk copies),
//           warps 9-20 MMA issuers (one per 64-wide k-block).
//   step t: producer lane kb polls the arrival counter of k-block kb (the 8 CTAs that own hidden units 64kb..64kb+63)
//           and pulls that k-block of h_{t-1} (64 batch rows x 64 units, bf16, 8 KB) from the L2-resident exchange
//           image with one TMA bulk copy onto its own mbarrier; MMA warp kb issues the four tcgen05.mma
//           (M=64, N=32, K=16) of its k-block as soon as it lands -- a tcgen05.mma issue costs ~60 ns of the issuing
//           thread (measured), so the 48 instructions of a step are issued from 12 threads in parallel -- all adding
//           into ONE fp32 accumulator tile [64 x 32] in TMEM that the epilogue re-zeroes after reading it; the epilogue
//           warps add the pre-computed input projection x_t W_ih^T + b (prefetched before the wait), apply the cell,
//           write h_t as bf16 straight into the UMMA image of the other exchange buffer, signal the k-block's counter
//           (one release-add per CTA) and only then store h_t / c_t / the activated gates (fp32 stash for BPTT).
//   Reference operator replaced: torch.nn.LSTM's recurrence (/root/reference/paule/models.py:349, :441).
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_lstm.cuh"

#ifdef PAULE_TC_TRACE
#define TRACE_DECL uint64_t tr_last = globaltimer_ns(); uint64_t tr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define TRACE(i) { const uint64_t _n = globaltimer_ns(); tr_acc[i] += _n - tr_last; tr_last = _n; }
#define TRACE_DUMP(base) { uint64_t* _o = reinterpret_cast<uint64_t*>(xchg + kXchgTraceOff) + (base); for (int _i = 0; _i < 8; ++_i) _o[_i] = tr_acc[_i]; }
#else
#define TRACE_DECL
#define TRACE(i) {}
#define TRACE_DUMP(base) {}
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
  uint64_t full[2][kNumKB];             // k-block landed; two sets so that the next step's barrier is armed early
  uint64_t mma_done;                    // accumulator complete (one arrival per MMA warp)
  uint64_t w_ready;
  uint32_t tmem_base;
};

constexpr int kEpiThreads = 256;
constexpr int kMmaWarps = kNumKB;                                 // one issuer per k-block
constexpr int kFwdThreads = kEpiThreads + 32 + 32 * kMmaWarps;   // + producer warp + MMA warps = 672

__global__ void __launch_bounds__(kFwdThreads, 1)
tc_lstm_fwd_kernel(float* __restrict__ gates, const uint8_t* __restrict__ packed, float* __restrict__ h_out,
                   float* __restrict__ c_out, uint8_t* __restrict__ xchg, uint8_t* __restrict__ img_seq, int T, int B,
                   int Bs) {
  extern __shared__ uint8_t smem_raw[];
  FwdSmem& S = *reinterpret_cast<FwdSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int q = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned int* counters = reinterpret_cast<unsigned int*>(xchg);          // counters[32 * kb]: one 128-byte line each
  volatile int* err = reinterpret_cast<volatile int*>(xchg + kXchgErrOff);
  // h_t images [12][64][128 B]: one per time step when the caller keeps them (they are the A operand of the next
  // layer's input-projection GEMM), otherwise two ping-pong images inside the exchange buffer
  uint8_t* hbuf = img_seq ? img_seq : xchg + kXchgHeader;
  const int img_mask = img_seq ? 0x7fffffff : 1;

  if (tid == 0) {
    for (int i = 0; i < kNumKB; ++i) { mbar_init(&S.full[0][i], 1); mbar_init(&S.full[1][i], 1); }
    mbar_init(&S.mma_done, kMmaWarps);
    mbar_init(&S.w_ready, 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<32>(&S.tmem_base);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = S.tmem_base;

  if (warp == 8) {
    // ===================== producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(&S.w_ready, kFwdSliceBytes);      // resident weights: one 48 KB bulk copy
      bulk_g2s(S.w, packed + (size_t)q * kFwdSliceBytes, kFwdSliceBytes, &S.w_ready);
    }
    if (lane < kNumKB) {
      const int kb = lane;
      // CTAs 8kb .. 8kb+7 own the hidden units of k-block kb (the last k-block has only CTAs 88, 89)
      const unsigned int owners = (unsigned int)(((kb + 1) * 8 <= kFwdCtas) ? 8 : kFwdCtas - kb * 8);
      if (T > 1) mbar_arrive_expect_tx(&S.full[0][kb], kRows * 128);        // armed one step ahead of the copy
      TRACE_DECL
      for (int t = 1; t < T; ++t) {
        grid_wait(counters + 32 * kb, (unsigned int)t * owners, err);     // k-block kb of h_{t-1} is complete
        TRACE(0)
        fence_proxy_async_global();                                            // generic-proxy writes -> async-proxy read
        const uint8_t* src = hbuf + (size_t)((t - 1) & img_mask) * kXchgImageBytes;
        bulk_g2s(S.a[kb], src + (size_t)kb * kRows * 128, kRows * 128, &S.full[(t - 1) & 1][kb]);
        if (t + 1 < T) mbar_arrive_expect_tx(&S.full[t & 1][kb], kRows * 128);   // next step's barrier (other set)
        TRACE(1)
      }
      if (q == 0 && lane == 0) TRACE_DUMP(0)
    }
    __syncwarp();
  } else if (warp >= 9) {
    // ===================== MMA issuers: warp 9+kb adds k-block kb into the shared accumulator tile =====================
    const int kb = warp - 9;
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kRows, kFwdN);
      mbar_wait(&S.w_ready, 0, err);
      const uint64_t da = make_smem_desc_sw128(smem_u32(S.a[kb]));
      const uint64_t db = make_smem_desc_sw128(smem_u32(S.w + (size_t)kb * kFwdN * 128));
      TRACE_DECL
      for (int t = 1; t < T; ++t) {
        mbar_wait(&S.full[(t - 1) & 1][kb], (uint32_t)(((t - 1) >> 1) & 1), err);
        TRACE(0)
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, 1u);   // the tile was pre-zeroed
        umma_commit(&S.mma_done);
        TRACE(1)
      }
      if (q == 0 && kb == 0) TRACE_DUMP(8)
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
    unsigned int* my_counter = counters + 32 * (q >> 3);   // this CTA's units live in k-block q / 8
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
        tmem_ld_x16(taddr, acc);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.f;
      }
      if (t + 1 < T) {   // re-arm the accumulator: every MMA of the next step adds into it
        tmem_zero_x16(taddr);
        tmem_st_wait();
      }
      tcgen05_fence_before();
      TRACE(1)
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
        // publish: CTA-wide barrier, then ONE gpu-scope release-add (cumulative over the stores the barrier ordered).
        // The consumer side issues fence.proxy.async between its acquire and its TMA read.
        named_bar_sync(1, kEpiThreads);
        if (tid == 0) grid_arrive(my_counter);
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
  if (warp == 9) tmem_dealloc<32>(tmem);
}

}  // namespace tc
}  // namespace paule

using namespace paule;
using namespace paule::tc;

extern "C" size_t paule_tc_packed_lstm_bytes(int64_t H, int64_t I) {
  (void)I;
  if (H != kH) return 0;
  return kPackedBytes;
}

extern "C" int paule_tc_pack_lstm(const float* w_ih, const float* w_hh, void* packed, int64_t H, int64_t I,
                                  paule_stream_t stream) {
  PAULE_REQUIRE(w_hh && packed);
  if (H != kH) return PAULE_ERR_UNSUPPORTED;
  uint8_t* img = reinterpret_cast<uint8_t*>(packed);
  PAULE_REQUIRE(reinterpret_cast<uintptr_t>(img) % 16 == 0);
  pack_fwd_kernel<<<kFwdCtas, 256, 0, as_stream(stream)>>>(w_hh, img);
  pack_bwd_kernel<<<kBwdCtas, 256, 0, as_stream(stream)>>>(w_hh, img + (size_t)kFwdCtas * kFwdSliceBytes);
  PAULE_LAUNCH_CHECK("pack kernels");
  return pack_v2(w_ih, w_hh, I, img, as_stream(stream));
}

extern "C" size_t paule_tc_rnn_xchg_bytes(int64_t B) {
  (void)B;
  // header (barrier counters, error flag) + forward: 2 h images; backward: 2 x 4 gate images (sized for the larger user)
  const size_t v1 = (size_t)8 * kXchgImageBytes;
  return (size_t)kXchgHeader + (v1 > kLLBytes ? v1 : kLLBytes);
}

extern "C" size_t paule_tc_x_image_bytes(int64_t T, int64_t B) {
  if (T <= 0 || B <= 0) return 0;
  return (size_t)T * (size_t)((B + kWq - 1) / kWq) * kXBlockBytes;
}

extern "C" int paule_tc_x_image(const float* x, void* img, int64_t T, int64_t B, int64_t I, paule_stream_t stream) {
  PAULE_REQUIRE(x && img && T >= 0 && B > 0 && I >= 1 && I <= kXK);
  PAULE_REQUIRE(reinterpret_cast<uintptr_t>(img) % 16 == 0);
  return x_image(x, img, T, B, I, as_stream(stream));
}

extern "C" int paule_tc_lstm_seq_fwd_x(float* gates, const void* packed, const float* bias, const void* x_img, float* h,
                                       float* c, void* xchg, void* h_img_seq, int64_t T, int64_t B, int math,
                                       paule_stream_t stream) {
  PAULE_REQUIRE(gates && packed && bias && x_img && c && xchg && T >= 0 && B > 0);
  PAULE_REQUIRE(h != nullptr || h_img_seq != nullptr);   // h may be NULL when only its bf16 images are consumed
  PAULE_REQUIRE(math == PAULE_MATH_BF16);
  if (T == 0) return PAULE_OK;
  PAULE_REQUIRE(reinterpret_cast<uintptr_t>(xchg) % 16 == 0 && reinterpret_cast<uintptr_t>(x_img) % 16 == 0);
  PAULE_REQUIRE(reinterpret_cast<uintptr_t>(h_img_seq) % 16 == 0);
  return lstm_seq_fwd2x(gates, packed, bias, x_img, h, c, xchg, h_img_seq, T, B, as_stream(stream));
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
  PAULE_REQUIRE(reinterpret_cast<uintptr_t>(h_img_seq) % 16 == 0);
  cudaStream_t s = as_stream(stream);
  if (!use_v1_fwd()) return lstm_seq_fwd2(gates, packed, h, c, xchg, h_img_seq, T, B, s);
  static bool attr_set = false;
  const int smem = (int)sizeof(FwdSmem) + 1024;
  if (!attr_set) {
    PAULE_CUDA(cudaFuncSetAttribute(tc_lstm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  // words are independent: batches larger than the UMMA M tile run as consecutive 64-word groups
  for (int64_t r0 = 0; r0 < B; r0 += kRows) {
    // zero the barrier words (and, without an image sequence, both ping-pong images: pad rows / columns must be 0)
    PAULE_CUDA(cudaMemsetAsync(xchg, 0, (size_t)kXchgHeader + (h_img_seq ? 0 : (size_t)2 * kXchgImageBytes), s));
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
