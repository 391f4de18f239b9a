// Persistent-RNN forward of one H=720 LSTM layer on tcgen05 (sm_100a), latency-optimised layout ("v2").
//
// What bounds a recurrent step is not arithmetic but the chain  h_t leaves an SM -> every SM holds h_t in shared
// memory -> MMA -> cell.  Two choices shorten it against tc_lstm_fwd.cu:
//
//   * operand roles are swapped: the resident W_hh slice is the A operand (M = 128 gate rows = 4 gates x 32 hidden
//     units) and the words are the N dimension (16 per CTA).  A CTA therefore ingests
//     16 x 768 values of h_{t-1} per step instead of 64 x 768; the grid is 23 unit groups x ceil(words / 16).
//   * h_t is exchanged as self-validating bf16 values: |h| < 1, so bit 14 of the encoding is free and carries a phase bit
//     (tc_common.cuh).  A reader that finds the expected phase holds the value, so there is no release fence on the
//     writer, no arrival counter, no acquire poll followed by a copy -- the 12 loader warps poll the data itself, drop
//     it into the swizzled UMMA layout in shared memory and issue their k-block's four tcgen05.mma (M=128, N=16, K=16)
//     into one shared TMEM accumulator.
//   * the resident W_hh slice lives in TENSOR memory (384 of the 512 columns), not shared memory: streaming a 192 KB A
//     operand out of shared memory every step costs 1.1 us, with A in TMEM the 48 MMAs of a step finish in 0.34 us
//     (tools/mma_bench.cu).
//
//   TMEM accumulator [128 lanes = gate rows, 16 columns = words]; lanes 4u..4u+3 are the i,f,g,o rows of unit u, so a
//   quad of epilogue threads transposes with three shuffles and every thread runs two complete cells.
//   Reference operator replaced: torch.nn.LSTM's recurrence (/root/reference/paule/models.py:349, :441).
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_lstm.cuh"

namespace paule {
namespace tc {

// A operand image of unit group ug (192 KB) in the tcgen05.st order [48 column octets][128 rows][8 x u32];
//   row m = 4 * u_local + gate  <->  W_hh[gate * H + 32 ug + u_local, :]  (rows of units >= H are zero)
__global__ void pack_fwd2_kernel(const float* __restrict__ w_hh, uint8_t* __restrict__ img) {
  const int ug = blockIdx.x;
  uint32_t* out = reinterpret_cast<uint32_t*>(img + (size_t)ug * kV2SliceBytes);
  for (int e = threadIdx.x; e < kV2M * kV2WCols; e += blockDim.x) {
    const int c8 = e / (kV2M * 8), m = (e / 8) % kV2M, i = e % 8;
    const int k = 2 * (c8 * 8 + i);
    const int u = ug * 32 + (m >> 2), gate = m & 3;
    const float* row = w_hh + (size_t)(gate * kH + (u < kH ? u : 0)) * kH;
    const float lo = (u < kH && k < kH) ? row[k] : 0.f, hi = (u < kH && k + 1 < kH) ? row[k + 1] : 0.f;
    out[e] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(lo)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(hi)) << 16);
  }
}

// A operand image of (unit group ugb, gate g) for the backward kernel, same order: row m = hidden unit 128 ugb + m,
// column pair k = unit index inside gate g:  A[m, k] = W_hh[g * H + k, 128 ugb + m]
__global__ void pack_bwd2_kernel(const float* __restrict__ w_hh, uint8_t* __restrict__ img) {
  const int ugb = blockIdx.x >> 2, g = blockIdx.x & 3;
  uint32_t* out = reinterpret_cast<uint32_t*>(img + (size_t)blockIdx.x * kV2SliceBytes);
  for (int e = threadIdx.x; e < kV2M * kV2WCols; e += blockDim.x) {
    const int c = e / kV2M, m = e % kV2M;   // m fastest: coalesced reads of a W_hh row
    const int k = 2 * c, j = ugb * kV2M + m;
    const float lo = (j < kH && k < kH) ? w_hh[(size_t)(g * kH + k) * kH + j] : 0.f;
    const float hi = (j < kH && k + 1 < kH) ? w_hh[(size_t)(g * kH + k + 1) * kH + j] : 0.f;
    out[(size_t)((c >> 3) * kV2M + m) * 8 + (c & 7)] =
        (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(lo)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(hi)) << 16);
  }
}

template <int NQ>
struct Fwd2Smem {
  uint8_t b[NQ][kV2BBytes];   // B operand per quarter: h_{t-1} of its 16 words, [12 kb][16 rows][128 B]
  uint64_t mma_done[NQ];      // accumulator of quarter q complete: one commit per loader warp
  uint64_t acc_free[NQ];      // accumulator of quarter q read and re-zeroed: one arrival per epilogue warp
  uint32_t tmem_base;
};

constexpr int kF2EpiWarps = 8;
constexpr int kF2Threads = 32 * (kF2EpiWarps + kNumKB);   // 640

// NQ = word quarters (16 words each) per CTA.  The quarters of a CTA are INDEPENDENT recurrences that share the resident
// weights: each has its own accumulator, barriers, operand buffer and exchange blocks, and the loader / epilogue warps
// visit them round-robin, so one quarter's cell runs while another quarter's h_t is in flight between the SMs.  NQ = 1
// is the latency-optimal layout for batches that fit one launch (<= 96 words); NQ = 2 / 4 fill the waiting time.
template <int NQ>
__global__ void __launch_bounds__(kF2Threads, 1)
tc_lstm_fwd2_kernel(float* __restrict__ gates, const uint8_t* __restrict__ packed, float* __restrict__ h_out,
                    float* __restrict__ c_out, uint8_t* __restrict__ xchg, uint8_t* __restrict__ img_seq, int T, int Bv,
                    int Bs, int w0) {
  extern __shared__ uint8_t smem_raw[];
  using Smem = Fwd2Smem<NQ>;
  Smem& S = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kGW = kWq * NQ;                        // words per CTA group
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int ug = blockIdx.x % kFwd2Groups, grp = blockIdx.x / kFwd2Groups;
  volatile int* err = reinterpret_cast<volatile int*>(xchg + kXchgErrOff);
  uint8_t* ll = xchg + kXchgHeader + (size_t)grp * NQ * 2 * kLLBlockBytes;   // [quarter][parity] blocks of this group

  if (tid == 0) {
    for (int q = 0; q < NQ; ++q) { mbar_init(&S.mma_done[q], kNumKB); mbar_init(&S.acc_free[q], kF2EpiWarps); }
    fence_mbar_init();
  }
  for (int i = tid; i < NQ * kV2BBytes / 16; i += kF2Threads) reinterpret_cast<uint4*>(&S.b[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_shared();
  if (warp == kF2EpiWarps) tmem_alloc<512>(&S.tmem_base);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = S.tmem_base;
  // resident weights: this CTA's [128 gate rows, 768] slice of W_hh goes into tensor memory once (384 columns)
  if (warp < 4) load_weights_to_tmem(packed + (size_t)ug * kV2SliceBytes, tmem, warp, lane);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();

  if (warp >= kF2EpiWarps) {
    // ===================== loader + MMA issuer of k-block kb =====================
    const int kb = warp - kF2EpiWarps;
    const uint32_t idesc = make_idesc_bf16(kV2M, kWq);
    const uint32_t ta = tmem + (uint32_t)(kV2WCol + kb * 32);   // A operand: 32 columns per k-block, 8 per K = 16
    // probes: lane p < 16 watches writer warp (CTA p>>3 of the k-block's two, lane group (p>>1)&3, column half p&1):
    // its lane 0 publishes row 8 (p&1), units 32 (p>>3) + 8 ((p>>1)&3) of the k-block
    const uint32_t probe_off = (uint32_t)(((lane & 1) * 8 * 64 + ((lane >> 3) & 1) * 32 + ((lane >> 1) & 3) * 8) * 2);
    TRACE_DECL
    for (int t = 1; t < T; ++t) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int rows = min(kWq, Bv - (grp * NQ + q) * kWq);   // valid words of this quarter: only their rows travel
        if (rows <= 0) continue;
        const bool prober = lane < 16 && (kb < kNumKB - 1 || lane < 8) && ((lane & 1) * 8 < rows);
        uint8_t* bdst = &S.b[q][(size_t)kb * kWq * 128];
        const uint64_t db = make_smem_desc_sw128(smem_u32(bdst));
        const uint8_t* src = ll + (size_t)(q * 2 + ((t - 1) & 1)) * kLLBlockBytes + (size_t)kb * (kWq * 128);
#ifdef PAULE_TC_TRACE
        while ((xchg_load(src) & kPhaseMask) != phase_bits(t - 1)) {}   // split the fetch: until the first value is visible
        TRACE(0)
        uint64_t ftr[2] = {0, 0};
        if (!xchg_fetch_kblock<1>(src, bdst, kb, phase_bits(t - 1), lane, probe_off, prober, rows, err, ftr)) break;
        tr_acc[6] += ftr[0] - tr_last;   // probe phase
        tr_acc[7] += ftr[1] * 1000;      // bulk passes (x1000 so that the printout shows passes per step)
#else
        if (!xchg_fetch_kblock<1>(src, bdst, kb, phase_bits(t - 1), lane, probe_off, prober, rows, err)) break;
#endif
        TRACE(1)
        fence_proxy_async_shared();   // generic-proxy shared-memory writes -> async-proxy (tensor core) reads
        __syncwarp();
        TRACE(2)
        mbar_wait(&S.acc_free[q], (uint32_t)((t - 1) & 1), err);   // long complete by now: the tile was zeroed a step ago
        tcgen05_fence_after();
        TRACE(3)
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem + (uint32_t)(kV2AccCol + q * kWq), ta + 8 * k, db + 2 * k, idesc, 1u);
          umma_commit(&S.mma_done[q]);
        }
        __syncwarp();
        TRACE(4)
#ifdef PAULE_TC_TRACE
        mbar_wait(&S.mma_done[q], (uint32_t)((t - 1) & 1), err);
        TRACE(5)
#endif
      }
    }
    if (blockIdx.x == 0 && kb == 3 && lane == 0) TRACE_DUMP(0)
  } else {
    // ===================== epilogue: the LSTM cell =====================
    const int lg = warp & 3, ch = warp >> 2;        // TMEM lane group (32 gate rows = 8 units), column half (8 words)
    const int gq = lane & 3, ul = lane >> 2;        // position in the quad = gate row held after the load; unit in the warp
    const int u = ug * 32 + lg * 8 + ul;
    const bool uvalid = u < kH;
    const int wl0 = ch * 8 + gq * 2;                // this thread's two words inside a quarter: wl0, wl0 + 1
    const uint32_t taddr = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(kV2AccCol + ch * 8);
    // value pair this thread publishes: even units take word 0 of the pair, odd units word 1
    const int e = ul & 1;
    const int prow = wl0 + e;
    const size_t ll_off = ((size_t)((ug >> 1) * kWq + prow) * 64 + (size_t)((ug & 1) * 32 + lg * 8 + (ul & ~1))) * 2;
    const bool b1 = (gq & 2) != 0, b0 = (gq & 1) != 0;
    float c_prev[NQ][2];
#pragma unroll
    for (int q = 0; q < NQ; ++q) c_prev[q][0] = c_prev[q][1] = 0.f;

#pragma unroll
    for (int q = 0; q < NQ; ++q) tmem_zero_x8(taddr + (uint32_t)(q * kWq));
    tmem_st_wait();
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0)
      for (int q = 0; q < NQ; ++q) mbar_arrive(&S.acc_free[q]);

    TRACE_DECL
    for (int t = 0; t < T; ++t) {
      // input projection of this step (independent of h_{t-1}): quarter 0's loads are issued before the wait, quarter
      // q + 1's while quarter q is computed
      float xp[4][2];
      auto load_xp = [&](int q) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int wp = grp * kGW + q * kWq + wl0 + k;
          const bool ok = uvalid && wp < Bv;
          const float* grow = gates + ((size_t)t * Bs + (ok ? wp : 0)) * (4 * kH) + (uvalid ? u : 0);
#pragma unroll
          for (int g = 0; g < 4; ++g) xp[g][k] = ok ? __ldg(grow + g * kH) : 0.f;
        }
      };
      load_xp(0);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        if ((grp * NQ + q) * kWq >= Bv) continue;   // empty quarter (uniform over the CTA)
        float pre[4][2];
        if (t > 0) {
          float acc[8];
          mbar_wait(&S.mma_done[q], (uint32_t)((t - 1) & 1), err);
          TRACE(0)
          tcgen05_fence_after();
          tmem_ld_x8(taddr + (uint32_t)(q * kWq), acc);
          TRACE(1)
          if (t + 1 < T) {   // re-arm the accumulator: every MMA of the next step adds into it
            tmem_zero_x8(taddr + (uint32_t)(q * kWq));
            tmem_st_wait();
          }
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&S.acc_free[q]);
          TRACE(2)
          // quad transpose (branch-free butterfly): lane gq holds gate gq for word pairs 0..3; afterwards it holds word
          // pair gq of all four gates.  Round 1 (xor 2) swaps the pair halves, round 2 (xor 1) the pairs inside a half.
          float k0[2], k1[2], r0[2], r1[2];   // kept / received pair-half after round 1: pair indices 2*b1 + {0, 1}
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            k0[x] = b1 ? acc[4 + x] : acc[0 + x];
            k1[x] = b1 ? acc[6 + x] : acc[2 + x];
            r0[x] = __shfl_xor_sync(0xffffffffu, b1 ? acc[0 + x] : acc[4 + x], 2);
            r1[x] = __shfl_xor_sync(0xffffffffu, b1 ? acc[2 + x] : acc[6 + x], 2);
          }
          float s[4][2];   // s[h] = pair gq of gate (gq ^ h)
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            s[0][x] = b0 ? k1[x] : k0[x];
            s[2][x] = b0 ? r1[x] : r0[x];
            s[1][x] = __shfl_xor_sync(0xffffffffu, b0 ? k0[x] : k1[x], 1);
            s[3][x] = __shfl_xor_sync(0xffffffffu, b0 ? r0[x] : r1[x], 1);
          }
#pragma unroll
          for (int x = 0; x < 2; ++x) {   // gate g sits in slot g ^ gq
            pre[0][x] = b1 ? (b0 ? s[3][x] : s[2][x]) : (b0 ? s[1][x] : s[0][x]);
            pre[1][x] = b1 ? (b0 ? s[2][x] : s[3][x]) : (b0 ? s[0][x] : s[1][x]);
            pre[2][x] = b1 ? (b0 ? s[1][x] : s[0][x]) : (b0 ? s[3][x] : s[2][x]);
            pre[3][x] = b1 ? (b0 ? s[0][x] : s[1][x]) : (b0 ? s[2][x] : s[3][x]);
          }
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) pre[g][0] = pre[g][1] = 0.f;
        }
        float hv[2], gi[2], gf[2], gg[2], go[2], cn[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          gi[k] = fast_sigmoid(fminf(fmaxf(pre[0][k] + xp[0][k], -30.f), 30.f));
          gf[k] = fast_sigmoid(fminf(fmaxf(pre[1][k] + xp[1][k], -30.f), 30.f));
          gg[k] = fast_tanh(fminf(fmaxf(pre[2][k] + xp[2][k], -15.f), 15.f));
          go[k] = fast_sigmoid(fminf(fmaxf(pre[3][k] + xp[3][k], -30.f), 30.f));
          cn[k] = gf[k] * c_prev[q][k] + gi[k] * gg[k];
          hv[k] = go[k] * fast_tanh(fminf(fmaxf(cn[k], -15.f), 15.f));
          c_prev[q][k] = cn[k];
        }
        if (q + 1 < NQ) load_xp(q + 1);   // next quarter's input projection while this one's results go out
        // pair neighbouring units (lane ^ 4) so that one thread owns {h[u], h[u+1]} of one word
        const float other = __shfl_xor_sync(0xffffffffu, e ? hv[0] : hv[1], 4);
        const __nv_bfloat162 pr = e ? __floats2bfloat162_rn(other, hv[1]) : __floats2bfloat162_rn(hv[0], other);
        const uint32_t payload = *reinterpret_cast<const uint32_t*>(&pr);
        const int wq = grp * kGW + q * kWq;           // first word of this quarter inside the launch
        const bool pvalid = wq + prow < Bv;
        if (t + 1 < T && pvalid)                      // critical path: the next step's operand
          xchg_store(ll + (size_t)(q * 2 + (t & 1)) * kLLBlockBytes + ll_off, payload | phase_bits(t));
        TRACE(3)
        // everything below is off the critical path: the next step is already fed
        if (img_seq != nullptr && uvalid && pvalid) {
          const int wg = w0 + wq + prow;              // global word (image addressing)
          *reinterpret_cast<uint32_t*>(img_seq + ((size_t)(wg / kRows) * (size_t)T + t) * kXchgImageBytes +
                                       umma_offset(kRows, wg % kRows, u & ~1)) = payload;
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int wp = wq + wl0 + k;
          if (uvalid && wp < Bv) {
            float* grow = gates + ((size_t)t * Bs + wp) * (4 * kH) + u;
            grow[0 * kH] = gi[k];
            grow[1 * kH] = gf[k];
            grow[2 * kH] = gg[k];
            grow[3 * kH] = go[k];
            const size_t o = ((size_t)t * Bs + wp) * kH + u;
            c_out[o] = cn[k];
            h_out[o] = hv[k];
          }
        }
        TRACE(4)
      }
    }
    if (blockIdx.x == 0 && tid == 0) TRACE_DUMP(8)
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kF2EpiWarps) tmem_dealloc<512>(tmem);
}

}  // namespace tc
}  // namespace paule

using namespace paule;
using namespace paule::tc;

namespace paule {
namespace tc {

int pack_v2(const float* w_hh, uint8_t* packed, cudaStream_t s) {
  pack_fwd2_kernel<<<kFwd2Groups, 256, 0, s>>>(w_hh, packed + kPackedFwd2Off);
  pack_bwd2_kernel<<<kBwd2Groups * 4, 256, 0, s>>>(w_hh, packed + kPackedBwd2Off);
  PAULE_LAUNCH_CHECK("pack v2 kernels");
  return PAULE_OK;
}

template <int NQ>
int launch_fwd2(float* gates, const void* packed, float* h, float* c, void* xchg, void* h_img_seq, int64_t T, int64_t B,
                cudaStream_t s) {
  static bool attr_set = false;
  const int smem = (int)sizeof(Fwd2Smem<NQ>) + 1024;
  if (!attr_set) {
    PAULE_CUDA(cudaFuncSetAttribute(tc_lstm_fwd2_kernel<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int64_t gw = (int64_t)kWq * NQ, pw = pass_words(B, kMaxQ, NQ);
  // words are independent: batches larger than one launch run as consecutive, balanced passes
  for (int64_t r0 = 0; r0 < B; r0 += pw) {
    const int Bv = (int)((B - r0 < pw) ? (B - r0) : pw);
    const int ng = (int)((Bv + gw - 1) / gw);
    // error flag; exchange blocks of this pass start with the phase bit set (0x4040 per value): never a valid first step
    PAULE_CUDA(cudaMemsetAsync(xchg, 0, (size_t)kXchgHeader, s));
    PAULE_CUDA(cudaMemsetAsync(reinterpret_cast<uint8_t*>(xchg) + kXchgHeader, 0x40, (size_t)ng * 2 * NQ * kLLBlockBytes, s));
    int Ti = (int)T, Bsi = (int)B, w0 = (int)r0, Bvi = Bv;
    float* gp = gates + r0 * 4 * kH;
    float* hp = h + r0 * kH;
    float* cp = c + r0 * kH;
    uint8_t* xc = reinterpret_cast<uint8_t*>(xchg);
    uint8_t* is = reinterpret_cast<uint8_t*>(h_img_seq);
    const uint8_t* pk = reinterpret_cast<const uint8_t*>(packed) + kPackedFwd2Off;
    void* args[] = {&gp, &pk, &hp, &cp, &xc, &is, &Ti, &Bvi, &Bsi, &w0};
    PAULE_CUDA(cudaLaunchCooperativeKernel((void*)tc_lstm_fwd2_kernel<NQ>, dim3(kFwd2Groups * ng), dim3(kF2Threads), args,
                                           (size_t)smem, s));
  }
  return PAULE_OK;
}

int lstm_seq_fwd2(float* gates, const void* packed, float* h, float* c, void* xchg, void* h_img_seq, int64_t T, int64_t B,
                  cudaStream_t s) {
  switch (choose_nq(B, kMaxQ)) {
    case 1: return launch_fwd2<1>(gates, packed, h, c, xchg, h_img_seq, T, B, s);
    case 2: return launch_fwd2<2>(gates, packed, h, c, xchg, h_img_seq, T, B, s);
    case 3: return launch_fwd2<3>(gates, packed, h, c, xchg, h_img_seq, T, B, s);
    default: return launch_fwd2<4>(gates, packed, h, c, xchg, h_img_seq, T, B, s);
  }
}

}  // namespace tc
}  // namespace paule
