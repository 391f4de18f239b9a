// Persistent reverse-time BPTT (input gradients only) of one H=720 LSTM layer on tcgen05 (sm_100a), latency-optimised
// layout ("v2"; see tc_lstm_fwd2.cu for the two ideas: weights as the resident A operand with 16 words as N, and the
// recurrent operand exchanged as self-validating 8-byte {bf16x2, step tag} elements).
//
//   dh_t[b,j] = dh_ext_t[b,j] + sum_{g<4} sum_{k<720} da^g_{t+1}[b,k] W_hh[g*720 + k, j];   da_t = cell adjoint(dh_t, ...)
//
//   grid = (6 unit groups of 128 hidden units) x (4 gates) x ceil(words / 16) CTAs in clusters of 4: K = 2880 is split
//   over the gates inside a cluster.  CTA (ugb, g) keeps A[m, k] = W_hh[g*720 + k, 128 ugb + m] resident (192 KB), pulls
//   gate g's da_{t+1} of its 16 words out of the exchange (24 KB of payload), runs 48 tcgen05.mma (M=128, N=16, K=16;
//   12 issuers, one TMEM accumulator) and owns the finalisation of hidden units 128 ugb + 32 g .. +31: TMEM lane group
//   lg of every CTA holds the partial sums of the units that sibling lg finalises, so each epilogue warp pushes its
//   lanes straight into sibling lg's shared memory with st.async, whose bytes complete on the DESTINATION's mbarrier
//   (tx-count): no release fence on the sender, no cluster-wide barrier in the loop.
//   Reference operator replaced: autograd through aten::lstm (discrepancy.backward(), paule/paule.py:1052).
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_lstm.cuh"

namespace paule {
namespace tc {

constexpr int kRedStride = 20;                      // floats per unit row: 16 words + pad (16-byte aligned, spreads banks)
constexpr uint32_t kRedBytes = 4 * 32 * kWq * 4;    // payload per step, quarter and CTA

template <int NQ>
struct Bwd2Smem {
  uint8_t b[NQ][kV2BBytes];          // B operand per quarter: gate g's da_{t+1} of its 16 words
  float red[NQ][4][32][kRedStride];  // partial sums from the 4 gate CTAs for this CTA's 32 units: [quarter][source][unit][word]
  uint64_t mma_done[NQ];
  uint64_t acc_free[NQ];
  uint64_t red_full[NQ];             // per step and quarter: 8 KB of st.async payload from the four CTAs of the cluster
  uint32_t tmem_base;
};

constexpr int kB2EpiWarps = 8;                            // cell warps of ONE epilogue group
__host__ __device__ constexpr int b2_threads(int EG) { return 32 * (kB2EpiWarps * EG + kNumKB); }   // 640 / 896

__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// remote (or local) 16-byte store whose completion is signalled on the DESTINATION CTA's mbarrier (complete_tx 16 bytes):
// the sender needs no fence and no separate arrive, the receiver sees the data once its barrier phase completes
__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, float a, float b, float c, float d, uint32_t cluster_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                   cluster_addr),
               "f"(a), "f"(b), "f"(c), "f"(d), "r"(cluster_mbar)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait_cluster(uint64_t* bar, uint32_t parity, volatile int* err_flag) {
  uint64_t t0 = 0;
  for (unsigned int spin = 0;; ++spin) {
    if (mbar_try_wait_cluster(bar, parity)) return true;
    if ((spin & 63u) == 63u) {
      if (t0 == 0) t0 = globaltimer_ns();
      if (*err_flag != 0) return false;
      if (globaltimer_ns() - t0 > kWatchdogNs) { *err_flag = 2; return false; }
    }
  }
}
// bf16x2 of a gradient pair for the exchange: |x| < 2 keeps bit 14 of the encoding free for the phase bit (a NaN / Inf /
// huge gradient is clamped on the recurrent path only; the fp32 stash and the dX images keep the raw value)
__device__ __forceinline__ uint32_t xchg_clamped(float2 v) {
  const float a = fabsf(v.x) < 1.5f ? v.x : copysignf(1.5f, v.x), b = fabsf(v.y) < 1.5f ? v.y : copysignf(1.5f, v.y);
  const __nv_bfloat162 r = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ void cluster_sync_all2() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// EG: epilogue groups (see tc_lstm_fwd2.cu): with EG = 2 a second group of 8 warps owns the second half of the CTA's quarters
// (contiguous halves, in the loaders' visiting order: an interleaved split delays quarter q's publication behind the MMAs of
// quarter q + 2 and measured 20 % slower), so two quarters' reductions and cell adjoints run concurrently.
// QS: lock-step quarters as in the forward kernel (the NQ quarters form NQ / QS slots; a slot's QS quarters share ONE loader
// visit, operand buffer and MMA of N = 16 QS, and epilogue group j owns quarter j of every slot).  Measured in round 2 for
// <4, 2, 2> and <2, 2, 2>: SLOWER than independent quarters (8.09 against 6.37 us per step at 256 words, 4.86 against 3.59 at 128)
// -- unlike the forward kernel, whose cell pass dominates a visit, here the reduction hop between the four gate CTAs does, and a
// pair's two reductions queue behind one MMA.  Only QS = 1 is instantiated (and tested).
template <int NQ, int EG, int QS = 1>
__global__ void __launch_bounds__(b2_threads(EG), 1)
tc_lstm_bwd2_kernel(float* __restrict__ gates, const float* __restrict__ c_seq, const uint8_t* __restrict__ packed,
                    const float* __restrict__ dh_seq, int dh_mode, const float* __restrict__ dh_last,
                    uint8_t* __restrict__ xchg, uint8_t* __restrict__ img_seq, int T, int Bv, int Bs, int w0, int keep_da,
                    WaveFlags wf) {
  extern __shared__ uint8_t smem_raw[];
  using Smem = Bwd2Smem<NQ>;
  Smem& S = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kGW = kWq * NQ;                        // words per CTA group (NQ independent quarters, see tc_lstm_fwd2.cu)
  constexpr int kEpiW = kB2EpiWarps * EG, kB2Threads = b2_threads(EG);
  static_assert(EG == 1 || EG == 2, "one or two epilogue groups");
  constexpr int kQPG = (NQ + EG - 1) / EG;             // quarters per epilogue group: group eg owns [eg kQPG, (eg + 1) kQPG)
  static_assert(NQ % QS == 0 && (QS == 1 || QS == EG), "lock-step quarters: one epilogue group per quarter of a slot");
  constexpr int NS = NQ / QS;                          // slots (independent recurrences as far as the loaders are concerned)
  constexpr int kSW = kWq * QS;                        // words per slot (the MMA N)
  constexpr bool kByQ = QS > 1;                        // group eg owns quarter eg of every slot
  constexpr int kOwn = kByQ ? NS : kQPG;               // quarters per epilogue group
  constexpr int kSlotBlk = QS * kLLBlockBytes;         // exchange block of one (slot, parity, gate): [12 kb][16 QS rows][128 B]
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int g = (int)cluster_ctarank_u32();                 // gate handled by this CTA's K slice
  const int cl = blockIdx.x >> 2;
  const int ugb = cl % kBwd2Groups, grp = cl / kBwd2Groups; // unit group (128 hidden units), word group
  volatile int* err = reinterpret_cast<volatile int*>(xchg + kXchgErrOff);
  uint8_t* ll = xchg + kXchgHeader + (size_t)grp * NQ * 2 * 4 * kLLBlockBytes;   // [quarter][parity][gate] blocks of this group

  if (tid == 0) {
    for (int q = 0; q < NQ; ++q) {
      mbar_init(&S.mma_done[q], kNumKB);                           // (indexed by slot: the first NS entries are used)
      mbar_init(&S.acc_free[q], kB2EpiWarps * (kByQ ? EG : 1));
      mbar_init(&S.red_full[q], 1);
    }
    fence_mbar_init();
  }
  for (int i = tid; i < NQ * kV2BBytes / 16; i += kB2Threads) reinterpret_cast<uint4*>(&S.b[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_shared();
  if (warp == kEpiW) tmem_alloc<512>(&S.tmem_base);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = S.tmem_base;
  // resident weights: A[m, k] = W_hh[g*720 + k, 128 ugb + m] goes into tensor memory once (384 columns)
  if (warp < 4) load_weights_to_tmem(packed + (size_t)(ugb * 4 + g) * kV2SliceBytes, tmem, warp, lane);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  cluster_sync_all2();   // every sibling's mbarriers are initialised before any st.async targets them

  if (warp >= kEpiW) {
    // ===================== loader + MMA issuer of k-block kb =====================
    const int kb = warp - kEpiW;
    const uint32_t idesc = make_idesc_bf16(kV2M, kSW);
    const uint32_t ta = tmem + (uint32_t)(kV2WCol + kb * 32);
    // probes: lane p < 16 watches cell warp p&7 (rows 2(p&7), 2(p&7)+1) of writer CTA p>>3 of the k-block's two
    const uint32_t probe_off = (uint32_t)(((lane & 7) * 2 * 64 + ((lane >> 3) & 1) * 32) * 2);
    // back-off between failed probes (as in the forward kernel): 400 / 500 ns measured within noise of none (2.81-2.86 us per
    // step at 64 words), 800 ns 3.16 us -- off by default
#ifndef PAULE_BWD_BACKOFF_NS
#define PAULE_BWD_BACKOFF_NS 0
#endif
    const unsigned int bwd_backoff = (NQ == 1 && gridDim.x >= 4 * 4 * kBwd2Groups) ? (unsigned int)PAULE_BWD_BACKOFF_NS : 0u;
    TL_DECL(0u, kTlLoaderEvents)
    TRACE_DECL
    for (int it = 1; it < T; ++it) {
#pragma unroll
      for (int q = 0; q < NS; ++q) {                              // q = slot (= quarter when QS == 1)
        const int rows = min(kSW, Bv - (grp * NQ + q * QS) * kWq);   // valid words of this slot: only their rows travel
        if (rows <= 0) continue;
        // the probe sits in the last quarter of the slot that has this lane's row (the quarters of a slot are published concurrently)
        const int prow = (lane & 7) * 2;
        const int pq = prow < rows ? min(QS - 1, (rows - 1 - prow) / kWq) : 0;
        const bool prober = lane < 16 && (kb < kNumKB - 1 || lane < 8) && (prow < rows);
        uint8_t* bdst = &S.b[q * QS][(size_t)kb * kSW * 128];
        const uint64_t db = make_smem_desc_sw128(smem_u32(bdst));
        const uint8_t* src = ll + (size_t)((q * 2 + ((it - 1) & 1)) * 4 + g) * kSlotBlk + (size_t)kb * (kSW * 128);
        const uint32_t poff = probe_off + (uint32_t)(pq * kWq * 64 * 2);
        if (kb == 3 && lane == 0) TL(1, q, it)   // loader: quarter visit starts
#ifdef PAULE_TC_TRACE
        uint64_t ftr[2] = {0, 0};
        if (!xchg_fetch_kblock<QS>(src, bdst, kb, phase_bits(it - 1), lane, poff, prober, rows, err, ftr, bwd_backoff)) break;
        tr_acc[0] += ftr[0] - tr_last;   // until the probes pass
        tr_last = ftr[0];
        TRACE(1)
#else
        if (!xchg_fetch_kblock<QS>(src, bdst, kb, phase_bits(it - 1), lane, poff, prober, rows, err, nullptr, bwd_backoff)) break;
#endif
        fence_proxy_async_shared();
        __syncwarp();
        if (kb == 3 && lane == 0) TL(2, q, it)   // loader: k-block fetched
        mbar_wait(&S.acc_free[q], (uint32_t)((it - 1) & 1), err);
        tcgen05_fence_after();
        TRACE(2)
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem + (uint32_t)(kV2AccCol + q * kSW), ta + 8 * k, db + 2 * k, idesc, 1u);
          umma_commit(&S.mma_done[q]);
        }
        __syncwarp();
        if (kb == 3 && lane == 0) TL(3, q, it)   // loader: MMAs issued
        TRACE(3)
#ifdef PAULE_TC_TRACE
        mbar_wait(&S.mma_done[q], (uint32_t)((it - 1) & 1), err);
        TRACE(4)
#endif
      }
    }
    if (blockIdx.x == 0 && kb == 3 && lane == 0) TRACE_DUMP(0)
  } else {
    // ===================== cell adjoint =====================
    // TMEM read role: lane group lg = the 32 units sibling lg finalises, column half ch = 8 words of a quarter
    const int eg = warp >> 3, w8 = warp & 7;        // epilogue group (owns the quarters q / kQPG == eg), warp inside the group
    const bool leader = w8 == 0 && lane == 0;       // arms the group's reduction barriers
    const bool tl0 = leader && eg == 0;             // timeline / trace thread
    const int lg = w8 & 3, ch = w8 >> 2;
    const uint32_t taddr = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(kV2AccCol + ch * 8);
    const uint32_t red_dst = mapa_u32(smem_u32(&S.red[0][g][lane][ch * 8]), (uint32_t)lg);
    const uint32_t bar_dst = mapa_u32(smem_u32(&S.red_full[0]), (uint32_t)lg);
    // cell role: unit pair a (units j, j+1) of word w of every quarter
    const int a = lane & 15, w = (lane >> 4) + 2 * w8;
    const int j = ugb * kV2M + g * 32 + 2 * a;
    const bool jvalid = j < kH;
    // byte offset of this thread's unit pair inside a slot's exchange block, quarter 0 of the slot (+ 16 rows per quarter)
    const size_t ll_off = ((size_t)((j >> 6) * kSW + w) * 64 + (size_t)(j & 63)) * 2;
    auto own = [&](int q) { return kByQ ? (q % QS == eg) : (q / kQPG == eg); };
    auto own_q = [&](int i) { return kByQ ? (i * QS + eg) : (eg * kQPG + i); };   // i-th quarter of this epilogue group
    const bool jpublish = j < kH + 16;   // k-block 11: units 704..735 are read, zeros beyond H
    float dc[NQ][2];
#pragma unroll
    for (int q = 0; q < NQ; ++q) dc[q][0] = dc[q][1] = 0.f;
#pragma unroll
    for (int q = 0; q < NQ; ++q)
      if (own(q)) tmem_zero_x8(taddr + (uint32_t)(q * kWq));
    tmem_st_wait();
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0)
      for (int q = 0; q < NQ; ++q)
        if (own(q)) mbar_arrive(&S.acc_free[q / QS]);

    TL_DECL(kTlLoaderEvents, kTlCellEvents)
    TRACE_DECL
    for (int it = 0; it < T; ++it) {
      const int t = T - 1 - it;
      // (1) everything that does not depend on da_{t+1}: stash, cell states, external gradient of one quarter
      // Two register sets: the loads of a quarter are issued a whole pipeline stage before they are consumed (the quarter after
      // next of this group is fetched while the current one is finalised), so the HBM latency of the stash reads is hidden in
      // the multi-quarter layouts too (ncu, 4 quarters per CTA: the cell warps sat behind these loads, the loaders idle-polled)
      struct QIn { float2 s_i, s_f, s_g, s_o, ct, cp, dh; };
      QIn qbuf[2];
      auto load_q = [&](int q, QIn& in) {
        float2 s_i, s_f, s_g, s_o, ct, cp, dh;
        s_i = s_f = s_g = s_o = ct = cp = dh = make_float2(0.f, 0.f);
        const int wp = grp * kGW + q * kWq + w;
        if (jvalid && wp < Bv) {
          const float* grow = gates + ((size_t)t * Bs + wp) * (4 * kH);
          s_i = *reinterpret_cast<const float2*>(grow + 0 * kH + j);
          s_f = *reinterpret_cast<const float2*>(grow + 1 * kH + j);
          s_g = *reinterpret_cast<const float2*>(grow + 2 * kH + j);
          s_o = *reinterpret_cast<const float2*>(grow + 3 * kH + j);
          ct = *reinterpret_cast<const float2*>(c_seq + ((size_t)t * Bs + wp) * kH + j);
          if (t > 0) cp = *reinterpret_cast<const float2*>(c_seq + ((size_t)(t - 1) * Bs + wp) * kH + j);
          if (wf.x_flags != nullptr && dh_mode != 0 && (dh_mode == 1 || (t >> 1) < (T >> 1))) {
            // layer wavefront: the external gradient of this step is produced while this kernel runs (streaming dX GEMM over the
            // dA images of the layer above, reverse time) -- wait until its pair of steps has been released for this word group
            const int src = dh_mode == 1 ? t : (t >> 1);
            const unsigned int* f = wf.x_flags + (size_t)((w0 + wp) / kRows) * wf.x_pairs + (src >> 1);
            uint64_t wd0 = 0;
            for (unsigned int spin = 0; ld_acquire_u32(f) < wf.x_target; ++spin) {
              __nanosleep(100);
              if ((spin & 255u) == 255u) {
                if (wd0 == 0) wd0 = globaltimer_ns();
                if (*err != 0) break;
                if (globaltimer_ns() - wd0 > kWatchdogNs) { *err = 1; break; }
              }
            }
          }
          if (dh_mode == 1) {
            dh = *reinterpret_cast<const float2*>(dh_seq + ((size_t)t * Bs + wp) * kH + j);
          } else if (dh_mode == 2 && (t >> 1) < (T >> 1)) {
            const float2 v = *reinterpret_cast<const float2*>(dh_seq + ((size_t)(t >> 1) * Bs + wp) * kH + j);
            dh = make_float2(0.5f * v.x, 0.5f * v.y);
          }
          if (dh_last != nullptr && t == T - 1) {
            const float2 v = *reinterpret_cast<const float2*>(dh_last + (size_t)wp * kH + j);
            dh.x += v.x; dh.y += v.y;
          }
        }
        in.s_i = s_i; in.s_f = s_f; in.s_g = s_g; in.s_o = s_o; in.ct = ct; in.cp = cp; in.dh = dh;
      };
      // stage 1 of quarter q: partial sums out of TMEM and over to the sibling that finalises them (st.async: the bytes
      // complete on the sibling's mbarrier)
      auto push_q = [&](int q) {
        float p[8];
        if (leader) mbar_arrive_expect_tx(&S.red_full[q], kRedBytes);   // arm this step's phase (count 1 + 8 KB of tx)
        TRACE(0)
        if (tl0) TL(10, q, it)   // epilogue: starts waiting for the accumulator
        mbar_wait(&S.mma_done[q / QS], (uint32_t)((it - 1) & 1), err);
        if (tl0) TL(11, q, it)   // epilogue: accumulator complete
        TRACE(1)
        tcgen05_fence_after();
        tmem_ld_x8(taddr + (uint32_t)(q * kWq), p);
        if (it + 1 < T) {   // re-arm the accumulator: every MMA of the next step adds into it
          tmem_zero_x8(taddr + (uint32_t)(q * kWq));
          tmem_st_wait();
        }
        tcgen05_fence_before();
        TRACE(2)
        st_async_v4(red_dst + (uint32_t)(q * 4 * 32 * kRedStride * 4), p[0], p[1], p[2], p[3], bar_dst + (uint32_t)(q * 8));
        st_async_v4(red_dst + (uint32_t)(q * 4 * 32 * kRedStride * 4) + 16, p[4], p[5], p[6], p[7], bar_dst + (uint32_t)(q * 8));
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.acc_free[q / QS]);
        if (tl0) TL(14, q, it)   // epilogue: partial sums pushed
        TRACE(3)
      };
      // stage 2 of quarter q: sum the four partials, cell adjoint, publish da_t
      auto finalize_q = [&](int q, const QIn& in) {
        const float2 s_i = in.s_i, s_f = in.s_f, s_g = in.s_g, s_o = in.s_o, ct = in.ct, cp = in.cp;
        float2 dh = in.dh;
        // cell adjoint (oracle: manual_lstm_backward_input).  It is linear in (dh, dc): every factor that only depends on the
        // stash is computed BEFORE the wait for the partial sums, so that the critical path after the wait is a handful of FMAs:
        //   do = dh * k_o;  dc' = dc + dh * k_c;  d_i = dc' * k_i;  d_f = dc' * k_f;  d_g = dc' * k_g;  dc <- dc' * s_f
        const float tc0 = fast_tanh(fminf(fmaxf(ct.x, -15.f), 15.f)), tc1 = fast_tanh(fminf(fmaxf(ct.y, -15.f), 15.f));
        const float2 k_o = make_float2(tc0 * s_o.x * (1.f - s_o.x), tc1 * s_o.y * (1.f - s_o.y));
        const float2 k_c = make_float2(s_o.x * (1.f - tc0 * tc0), s_o.y * (1.f - tc1 * tc1));
        const float2 k_i = make_float2(s_g.x * s_i.x * (1.f - s_i.x), s_g.y * s_i.y * (1.f - s_i.y));
        const float2 k_f = make_float2(cp.x * s_f.x * (1.f - s_f.x), cp.y * s_f.y * (1.f - s_f.y));
        const float2 k_g = make_float2(s_i.x * (1.f - s_g.x * s_g.x), s_i.y * (1.f - s_g.y * s_g.y));
        if (it > 0) {
          mbar_wait_cluster(&S.red_full[q], (uint32_t)((it - 1) & 1), err);
          if (tl0) TL(15, q, it)   // epilogue: partial sums of all four siblings arrived
          TRACE(4)
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            dh.x += S.red[q][s][2 * a][w];
            dh.y += S.red[q][s][2 * a + 1][w];
          }
        }
        float2 d_i, d_f, d_g, d_o;
        {
          const float dc0 = dc[q][0] + dh.x * k_c.x, dc1 = dc[q][1] + dh.y * k_c.y;
          d_i = make_float2(dc0 * k_i.x, dc1 * k_i.y);
          d_f = make_float2(dc0 * k_f.x, dc1 * k_f.y);
          d_g = make_float2(dc0 * k_g.x, dc1 * k_g.y);
          d_o = make_float2(dh.x * k_o.x, dh.y * k_o.y);
          dc[q][0] = dc0 * s_f.x;
          dc[q][1] = dc1 * s_f.y;
        }
        const int wp = grp * kGW + q * kWq + w;
        const bool wvalid = wp < Bv;
        if (t > 0 && jpublish && wvalid) {   // da_t into the four gate blocks of the exchange: critical path of the next step
          uint8_t* dst = ll + (size_t)(((q / QS) * 2 + (it & 1)) * 4) * kSlotBlk + ll_off + (size_t)(q % QS) * (kWq * 64 * 2);
          const uint32_t ph = phase_bits(it);
          xchg_store(dst + 0 * (size_t)kSlotBlk, xchg_clamped(d_i) | ph);
          xchg_store(dst + 1 * (size_t)kSlotBlk, xchg_clamped(d_f) | ph);
          xchg_store(dst + 2 * (size_t)kSlotBlk, xchg_clamped(d_g) | ph);
          xchg_store(dst + 3 * (size_t)kSlotBlk, xchg_clamped(d_o) | ph);
          // off the critical path: a gradient beyond the +-1.5 exchange bound (or NaN / Inf) was clamped above -- record it
          // in the header's informational word (NOT the status word: the wait loops abort on that one) so that the caller learns
          // its results deviate
          const float mx = fmaxf(fmaxf(fmaxf(fabsf(d_i.x), fabsf(d_i.y)), fmaxf(fabsf(d_f.x), fabsf(d_f.y))),
                                 fmaxf(fmaxf(fabsf(d_g.x), fabsf(d_g.y)), fmaxf(fabsf(d_o.x), fabsf(d_o.y))));
          const bool nan = (d_i.x != d_i.x) || (d_i.y != d_i.y) || (d_f.x != d_f.x) || (d_f.y != d_f.y) || (d_g.x != d_g.x) ||
                           (d_g.y != d_g.y) || (d_o.x != d_o.x) || (d_o.y != d_o.y);
          if (!(mx < 1.5f) || nan) *reinterpret_cast<volatile int*>(xchg + kXchgClampOff) = 1;
        }
        if (tl0) TL(12, q, it)   // epilogue: quarter published
        TRACE(5)
        if (jvalid && wvalid) {   // off the critical path: bf16 images (A operand of the dX GEMM) and fp32 da_t over the stash
          if (img_seq != nullptr) {
            const int wg = w0 + wp;
            const __nv_bfloat162 b_i = __floats2bfloat162_rn(d_i.x, d_i.y), b_f = __floats2bfloat162_rn(d_f.x, d_f.y);
            const __nv_bfloat162 b_g = __floats2bfloat162_rn(d_g.x, d_g.y), b_o = __floats2bfloat162_rn(d_o.x, d_o.y);
            uint8_t* d = img_seq + ((size_t)(wg / kRows) * (size_t)T + t) * 4 * kXchgImageBytes + umma_offset(kRows, wg % kRows, j);
            *reinterpret_cast<uint32_t*>(d + 0 * (size_t)kXchgImageBytes) = *reinterpret_cast<const uint32_t*>(&b_i);
            *reinterpret_cast<uint32_t*>(d + 1 * (size_t)kXchgImageBytes) = *reinterpret_cast<const uint32_t*>(&b_f);
            *reinterpret_cast<uint32_t*>(d + 2 * (size_t)kXchgImageBytes) = *reinterpret_cast<const uint32_t*>(&b_g);
            *reinterpret_cast<uint32_t*>(d + 3 * (size_t)kXchgImageBytes) = *reinterpret_cast<const uint32_t*>(&b_o);
          }
          if (keep_da) {   // fp32 da_t over the stash; skipped when the caller only consumes the bf16 images (dX on tcgen05)
            float* grow = gates + ((size_t)t * Bs + wp) * (4 * kH);
            *reinterpret_cast<float2*>(grow + 0 * kH + j) = d_i;
            *reinterpret_cast<float2*>(grow + 1 * kH + j) = d_f;
            *reinterpret_cast<float2*>(grow + 2 * kH + j) = d_g;
            *reinterpret_cast<float2*>(grow + 3 * kH + j) = d_o;
          }
        }
        if (wf.img_flags != nullptr && (grp * NQ + q) * kWq < Bv) {
          // layer wavefront: this warp's dA image stores of step t are out -- one release-arrival per (step, quarter, warp)
          __syncwarp();
          if (lane == 0) red_release_add_u32(wf.img_flags + (size_t)((w0 + (grp * NQ + q) * kWq) / kRows) * (size_t)T + t, 1u);
        }
        if (tl0) TL(13, q, it)   // epilogue: image / stash stores issued
        TRACE(6)
      };
      // software pipeline over this group's quarters: quarter q's push overlaps quarter q-1's wait for its partial sums; the
      // first two quarters of the group are loaded here, quarter q + 2 right after quarter q has been finalised (register set
      // q & 1 is free again)
#pragma unroll
      for (int i = 0; i < 2 && i < kOwn; ++i)
        if (own_q(i) < NQ) load_q(own_q(i), qbuf[i & 1]);
#pragma unroll
      for (int i = 0; i <= kOwn; ++i) {
        if (i < kOwn && own_q(i) < NQ && it > 0) {
          if ((grp * NQ + own_q(i)) * kWq < Bv) {
            push_q(own_q(i));
          } else if (kByQ && (grp * NQ + (own_q(i) / QS) * QS) * kWq < Bv) {
            // an empty quarter of a slot that holds words: the loaders still wait for this group's release of the accumulator
            if (lane == 0) mbar_arrive(&S.acc_free[own_q(i) / QS]);
          }
        }
        if (i >= 1 && own_q(i - 1) < NQ) {
          if ((grp * NQ + own_q(i - 1)) * kWq < Bv) finalize_q(own_q(i - 1), qbuf[(i - 1) & 1]);
          if (i + 1 < kOwn && own_q(i + 1) < NQ) load_q(own_q(i + 1), qbuf[(i + 1) & 1]);
        }
      }
    }
    if (blockIdx.x == 0 && tl0) TRACE_DUMP(8)
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kEpiW) tmem_dealloc<512>(tmem);
  cluster_sync_all2();   // no CTA exits while a sibling may still address its shared memory
}

// B = words of the whole batch (row stride of the time-major tensors); the launch covers words [seg0, seg0 + seg_words)
// (seg_words = 0: all of them), in balanced passes of this layout when they exceed one launch
template <int NQ, int EG, int QS = 1>
int launch_bwd2(float* gates, const float* c, const void* packed, const float* dh_seq, int dh_mode, const float* dh_last,
                void* xchg, void* da_img_seq, int64_t T, int64_t B, int keep_da, cudaStream_t s, WaveFlags wf, int64_t seg0 = 0,
                int64_t seg_words = 0) {
  static unsigned long long attr_set = 0ull;
  const int smem_own = (int)sizeof(Bwd2Smem<NQ>) + 1024;
  const int smem = smem_own > kExclusiveSmemBytes ? smem_own : kExclusiveSmemBytes;   // one CTA per SM, whatever runs beside it
  if (once_per_device(attr_set)) {
    PAULE_CUDA(cudaFuncSetAttribute(tc_lstm_bwd2_kernel<NQ, EG, QS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  // Cluster launch WITHOUT the cooperative attribute: at most 120 CTAs (30 clusters of 4, one CTA per SM by shared memory and
  // TMEM) always fit the 132 cluster-schedulable SMs, so every CTA is co-resident without the co-residency check -- and
  // Nsight Compute rejects the cooperative + cluster combination (round 1: the driver's ncu pass over smoke() died here).
  const int64_t seg_end = seg_words > 0 ? seg0 + seg_words : B;
  const int64_t gw = (int64_t)kWq * NQ, pw = pass_words(seg_end - seg0, kMaxQBwd, NQ);
  for (int64_t r0 = seg0; r0 < seg_end; r0 += pw) {
    const int Bv = (int)((seg_end - r0 < pw) ? (seg_end - r0) : pw);
    const int ng = (int)((Bv + gw - 1) / gw);
    // (the status word in the header is sticky: it starts at zero and is never cleared by a launch)
    PAULE_CUDA(cudaMemsetAsync(reinterpret_cast<uint8_t*>(xchg) + kXchgHeader, 0x40, (size_t)ng * 2 * 4 * NQ * kLLBlockBytes, s));
    float* gp = gates + r0 * 4 * kH;
    const float* cp = c + r0 * kH;
    const float* dsp = dh_seq ? dh_seq + r0 * kH : nullptr;
    const float* dlp = dh_last ? dh_last + r0 * kH : nullptr;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(kBwd2Groups * 4 * ng);
    cfg.blockDim = dim3(b2_threads(EG));
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = s;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = 4;
    attrs[0].val.clusterDim.y = 1;
    attrs[0].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    uint8_t* xc = reinterpret_cast<uint8_t*>(xchg);
    uint8_t* is = reinterpret_cast<uint8_t*>(da_img_seq);
    const uint8_t* pk = reinterpret_cast<const uint8_t*>(packed) + kPackedBwd2Off;
    PAULE_CUDA(cudaLaunchKernelEx(&cfg, tc_lstm_bwd2_kernel<NQ, EG, QS>, gp, cp, pk, dsp, dh_mode, dlp, xc, is, (int)T, Bv, (int)B,
                                  (int)r0, keep_da, wf));
  }
  return PAULE_OK;
}

// CTAs of one launch that holds B words with nq quarters per CTA; 0 = does not fit one launch (at most kMaxQBwd word groups)
int bwd2_ctas(int64_t B, int nq) {
  if (nq < 1 || nq > 4) return 0;
  const int64_t quarters = (B + kWq - 1) / kWq, groups = (quarters + nq - 1) / nq;
  return groups <= kMaxQBwd ? (int)groups * kBwd2Groups * 4 : 0;
}
int bwd2_default_nq(int64_t B) { return choose_nq(B, kMaxQBwd); }
int bwd2_pass_plan(int64_t B, int32_t* nq_out, int32_t* words_out, int cap_out) {
  static const bool balanced = getenv("PAULE_RNN_BALANCED") != nullptr || getenv("PAULE_RNN_NQ") != nullptr;
  const int64_t cap = (int64_t)kMaxQBwd * kWq * 4;
  PassPlan pp;
  if (!(B > cap && !balanced && plan_passes(B, kMaxQBwd, kBwdStepUs, &pp))) {   // one layout, balanced passes
    const int nq = choose_nq(B, kMaxQBwd);
    const int64_t pw = pass_words(B, kMaxQBwd, nq);
    pp.n = 0;
    for (int64_t r0 = 0; r0 < B && pp.n < 96; r0 += pw) { pp.nq[pp.n] = nq; pp.words[pp.n] = (int)(B - r0 < pw ? B - r0 : pw); ++pp.n; }
  }
  for (int i = 0; i < pp.n && i < cap_out; ++i) { nq_out[i] = pp.nq[i]; words_out[i] = pp.words[i]; }
  return pp.n;
}
int bwd2_passes(int64_t B) { return bwd2_pass_plan(B, nullptr, nullptr, 0); }

static int bwd2_segment(float* gates, const float* c, const void* packed, const float* dh_seq, int dh_mode, const float* dh_last,
                        void* xchg, void* da_img_seq, int64_t T, int64_t B, int keep_da, cudaStream_t s, WaveFlags wf, int nq,
                        int64_t seg0, int64_t seg_words);

int lstm_seq_bwd2(float* gates, const float* c, const void* packed, const float* dh_seq, int dh_mode, const float* dh_last,
                  void* xchg, void* da_img_seq, int64_t T, int64_t B, int keep_da, cudaStream_t s, WaveFlags wf, int force_nq) {
  // more words than one launch holds: cut into passes by summed step time (tc_lstm.cuh, plan_passes); PAULE_RNN_BALANCED=1 or a
  // forced layout keeps the balanced passes of one layout
  static const bool balanced = getenv("PAULE_RNN_BALANCED") != nullptr || getenv("PAULE_RNN_NQ") != nullptr;
  if (force_nq <= 0 && !balanced && B > (int64_t)kMaxQBwd * kWq * 4) {
    PassPlan pp;
    if (plan_passes(B, kMaxQBwd, kBwdStepUs, &pp)) {
      int64_t r0 = 0;
      for (int i = 0; i < pp.n; ++i) {
        const int rc = bwd2_segment(gates, c, packed, dh_seq, dh_mode, dh_last, xchg, da_img_seq, T, B, keep_da, s, wf, pp.nq[i], r0,
                                    pp.words[i]);
        if (rc != PAULE_OK) return rc;
        r0 += pp.words[i];
      }
      return PAULE_OK;
    }
  }
  // two quarters per CTA: two epilogue groups, one quarter each (4.57 -> 3.62 us per step at 128 words).  With three or four
  // quarters a second group measured 4-6 % SLOWER (contiguous halves; 20 % slower interleaved) and is not used.
  // PAULE_RNN_EG=1 restores one group everywhere (A/B timing).
  int nq = choose_nq(B, kMaxQBwd);
  if (force_nq > 0) {   // layer wavefront: BPTT kernels that run side by side must be the SAME instantiation (see plan_step.cu)
    if (bwd2_ctas(B, force_nq) == 0) return PAULE_ERR_ARG;
    nq = force_nq;
  }
  return bwd2_segment(gates, c, packed, dh_seq, dh_mode, dh_last, xchg, da_img_seq, T, B, keep_da, s, wf, nq, 0, 0);
}

static int bwd2_segment(float* gates, const float* c, const void* packed, const float* dh_seq, int dh_mode, const float* dh_last,
                        void* xchg, void* da_img_seq, int64_t T, int64_t B, int keep_da, cudaStream_t s, WaveFlags wf, int nq,
                        int64_t seg0, int64_t seg_words) {
  static const bool one_group = getenv("PAULE_RNN_EG") != nullptr && atoi(getenv("PAULE_RNN_EG")) == 1;
#define PAULE_BWD_CASE(NQ_, EG_) \
  return launch_bwd2<NQ_, EG_>(gates, c, packed, dh_seq, dh_mode, dh_last, xchg, da_img_seq, T, B, keep_da, s, wf, seg0, seg_words)
  if (nq == 1) PAULE_BWD_CASE(1, 1);
  if (nq == 2 && !one_group) PAULE_BWD_CASE(2, 2);
  if (nq == 2) PAULE_BWD_CASE(2, 1);
  if (nq == 3) PAULE_BWD_CASE(3, 1);
  PAULE_BWD_CASE(4, 1);
#undef PAULE_BWD_CASE
}

}  // namespace tc
}  // namespace paule
