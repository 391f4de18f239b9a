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

// Input-projection weights of unit group ug (16 KB), same row order and tcgen05.st layout as pack_fwd2_kernel, K = 64.
// input_size <= 32 (the cps): columns [0,32) and [32,64) both hold W_ih -- the operand carries x as a hi/lo bf16 pair, so
// the planned variable itself is not rounded.  32 < input_size <= 64 (the mel): columns [0, input_size) hold W_ih.
__global__ void pack_x_kernel(const float* __restrict__ w_ih, int I, uint8_t* __restrict__ img) {
  const int ug = blockIdx.x;
  const bool split = I <= 32;
  uint32_t* out = reinterpret_cast<uint32_t*>(img + (size_t)ug * kXSliceBytes);
  for (int e = threadIdx.x; e < kV2M * kXCols; e += blockDim.x) {
    const int c8 = e / (kV2M * 8), m = (e / 8) % kV2M, i = e % 8;
    const int k = 2 * (c8 * 8 + i);
    const int u = ug * 32 + (m >> 2), gate = m & 3;
    float v[2];
#pragma unroll
    for (int x = 0; x < 2; ++x) {
      const int kk = split ? ((k + x) & 31) : (k + x);
      v[x] = (u < kH && kk < I) ? w_ih[(size_t)(gate * kH + u) * I + kk] : 0.f;
    }
    out[e] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[0])) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[1])) << 16);
  }
}

// x [T,B,I] fp32 -> operand images [T][ceil(B/16)][16 rows][128 B] (bf16, SWIZZLE_128B): one 2 KB block per (step, word
// quarter), fetched by the recurrent kernel with one TMA bulk copy.  Split mode: column k < 32 = bf16(x_k), column 32 + k =
// bf16(x_k - float(bf16(x_k))).  Rows of words >= B and unused columns stay zero (the buffer is zero-filled once).
__global__ void x_image_kernel(const float* __restrict__ x, uint8_t* __restrict__ img, int64_t T, int64_t B, int I) {
  const bool split = I <= 32;
  const int64_t Q = (B + kWq - 1) / kWq;
  const int64_t total = T * B * I;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(e % I);
    const int64_t tb = e / I, b = tb % B, t = tb / B;
    const int row = (int)(b % kWq);
    uint8_t* blk = img + ((size_t)t * Q + (size_t)(b / kWq)) * kXBlockBytes + (size_t)row * 128;
    const float v = x[e];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    *reinterpret_cast<__nv_bfloat16*>(blk + (((k >> 3) ^ (row & 7)) << 4) + (k & 7) * 2) = hi;
    if (split) {
      const int k2 = 32 + k;
      *reinterpret_cast<__nv_bfloat16*>(blk + (((k2 >> 3) ^ (row & 7)) << 4) + (k2 & 7) * 2) =
          __float2bfloat16_rn(v - __bfloat162float(hi));
    }
  }
}

template <int NS, int QS>
struct Fwd2Smem {
  uint8_t b[NS][QS * kV2BBytes];   // B operand per slot: h_{t-1} of its 16 QS words, [12 kb][16 QS rows][128 B]
  uint64_t mma_done[NS];           // accumulator of slot s complete: one commit per loader warp
  uint64_t acc_free[NS];           // accumulator of slot s read and re-zeroed: one arrival per epilogue warp
  uint64_t xfull[2];               // fused input projection: x_t blocks of all quarters landed (TMA tx-count)
  uint32_t tmem_base;
  alignas(1024) uint8_t xb[2][NS * QS][kXBlockBytes];   // x_t / x_{t+1} operand blocks (double-buffered)
};

constexpr int kF2EpiWarps = 8;                            // cell warps of ONE epilogue group: 4 TMEM lane groups x 2 column halves
__host__ __device__ constexpr int f2_threads(int EG) { return 32 * (kF2EpiWarps * EG + kNumKB); }   // 640 (one group) / 896 (two)

// A CTA carries NS x QS word quarters (16 words each) behind its resident weights.  The NS SLOTS are INDEPENDENT
// recurrences (own accumulator, barriers, operand buffer, exchange blocks) visited round-robin by the loader and epilogue
// warps, so one slot's cell runs while another slot's h_t is in flight between the SMs; the QS quarters of a slot advance in
// lock-step as one MMA of N = 16 QS, so one probe + fetch latency covers all of them.  (1,1) is the latency-optimal layout
// for batches that fit one launch (<= 96 words); (2,1), (3,1), (2,2) put 32 / 48 / 64 words on a CTA.
//
// FUSED: the layer's input projection runs inside the recurrence.  W_ih (K <= 64) sits in 32 more TMEM columns, x_t of the
// CTA's quarters arrives as 2 KB operand blocks by TMA one step ahead, and loader warp 11 -- whose k-block holds only 16
// real hidden units, i.e. one useful MMA -- issues the four x MMAs into the same accumulator; the epilogue adds the bias
// from registers.  `gates` is then output only (the activated-gate stash): the [T,B,2880] pre-activation tensor is never
// written to or read from HBM, and the x MMAs do not wait for the exchange.
//
// EG: epilogue groups.  A quarter's cell pass is a dependent chain (TMEM load -> quad transpose -> activations -> publish ->
// stash stores, ~1.25 us) that one group of 8 warps runs for the CTA's quarters one after the other; in the multi-slot layouts
// that chain, not the tensor core or the exchange, bounds the CTA.  With EG = 2 a second group of 8
// warps runs concurrently: it takes the SECOND quarter of every lock-step pair (QS = 2: both quarters of a slot are then
// published after one cell pass instead of two, which is what the other CTAs wait for), or the odd slots when QS = 1.
template <int NS, int QS, bool FUSED, int EG>
__global__ void __launch_bounds__(f2_threads(EG), 1)
tc_lstm_fwd2_kernel(float* __restrict__ gates, const uint8_t* __restrict__ packed, float* __restrict__ h_out,
                    float* __restrict__ c_out, uint8_t* __restrict__ xchg, uint8_t* __restrict__ img_seq, int T, int Bv,
                    int Bs, int w0, const uint8_t* __restrict__ packed_x, const uint8_t* __restrict__ x_img,
                    const float* __restrict__ bias, int Qtot, WaveFlags wf) {
  extern __shared__ uint8_t smem_raw[];
  using Smem = Fwd2Smem<NS, QS>;
  Smem& S = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kSW = kWq * QS;                        // words per slot (the MMA N)
  constexpr int kGW = kSW * NS;                        // words per CTA group
  constexpr int kBlk = QS * kLLBlockBytes;             // exchange block of one (slot, parity): [12 kb][16 QS rows][128 B]
  constexpr int kEpiW = kF2EpiWarps * EG;              // epilogue warps of the CTA; loader warp kb is warp kEpiW + kb
  constexpr int kF2Threads = f2_threads(EG);
  static_assert(EG == 1 || EG == 2, "one or two epilogue groups");
  constexpr bool kByQ = EG > 1 && QS == EG;            // group eg owns quarter eg of every slot (else: the slots s % EG == eg)
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int ug = blockIdx.x % kFwd2Groups, grp = blockIdx.x / kFwd2Groups;
  volatile int* err = reinterpret_cast<volatile int*>(xchg + kXchgErrOff);
  uint8_t* ll = xchg + kXchgHeader + (size_t)grp * NS * 2 * kBlk;   // [slot][parity] blocks of this group

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&S.mma_done[s], kNumKB); mbar_init(&S.acc_free[s], kF2EpiWarps * (kByQ ? EG : 1)); }
    mbar_init(&S.xfull[0], 1);
    mbar_init(&S.xfull[1], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < NS * QS * kV2BBytes / 16; i += kF2Threads) reinterpret_cast<uint4*>(&S.b[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_shared();
  if (warp == kEpiW) tmem_alloc<512>(&S.tmem_base);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = S.tmem_base;
  // resident weights: this CTA's [128 gate rows, 768] slice of W_hh goes into tensor memory once (384 columns)
  if (warp < 4) {
    load_weights_to_tmem(packed + (size_t)ug * kV2SliceBytes, tmem, warp, lane);
    if (FUSED) load_weights_to_tmem(packed_x + (size_t)ug * kXSliceBytes, tmem, warp, lane, kXWCol, kXCols);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();

  if (warp >= kEpiW) {
    // ===================== loader + MMA issuer of k-block kb =====================
    const int kb = warp - kEpiW;
    const uint32_t idesc = make_idesc_bf16(kV2M, kSW);
    const uint32_t ta = tmem + (uint32_t)(kV2WCol + kb * 32);   // A operand: 32 columns per k-block, 8 per K = 16
    const bool xwarp = FUSED && kb == kNumKB - 1;   // this warp also feeds the fused input projection
    // four or more single-slot word quarters polling at once: 300 ns between failed probes (tc_lstm.cuh); sweep at 64 words:
    // 0 / 200 ns 2.27 / 2.30 us per step, 300 / 400 / 500 ns 2.20, 600 ns 2.37
#ifndef PAULE_FWD_BACKOFF_NS
#define PAULE_FWD_BACKOFF_NS 300
#endif
    const unsigned int backoff = (NS == 1 && gridDim.x >= 4 * kFwd2Groups) ? (unsigned int)PAULE_FWD_BACKOFF_NS : 0u;
    // a single word quarter with at most four words (one 16-byte load per lane covers the k-block): no probes, the block read
    // itself is the poll (tc_lstm.cuh)
    const bool solo = NS == 1 && gridDim.x == kFwd2Groups && Bv <= 4;
    const unsigned int reread = solo ? 300u : 0u;
    const int nk = (kb == kNumKB - 1) ? 1 : 4;       // k-block 11 holds 16 real units: one K = 16 step, the rest is padding
    const int q_first = w0 / kWq + grp * NS * QS;    // first global word quarter of this CTA (x image addressing)
    int nvq = 0;                                     // quarters of this CTA that hold words (a prefix)
#pragma unroll
    for (int q = 0; q < NS * QS; ++q) nvq += (grp * kGW + q * kWq < Bv) ? 1 : 0;
    auto fetch_x = [&](int t) {   // x_t blocks of every valid quarter: one bulk copy each onto xfull[t & 1]
      if (wf.x_flags != nullptr) {
        // layer wavefront: x_t is produced while this kernel runs -- wait until the streaming GEMM has released the pair of
        // steps that holds x_t for every 64-word group this CTA's quarters belong to (4 arrivals: its epilogue warps)
        uint64_t wd0 = 0;
        for (int g64 = q_first / 4; g64 <= (q_first + nvq - 1) / 4; ++g64) {
          const unsigned int* f = wf.x_flags + (size_t)g64 * wf.x_pairs + (t >> 1);
          for (unsigned int spin = 0; ld_acquire_u32(f) < wf.x_target; ++spin) {
            __nanosleep(100);
            if ((spin & 255u) == 255u) {
              if (wd0 == 0) wd0 = globaltimer_ns();
              if (*err != 0) break;
              if (globaltimer_ns() - wd0 > kWatchdogNs) { *err = 1; break; }
            }
          }
        }
        fence_proxy_async_global();   // acquire (generic proxy) -> TMA reads (async proxy)
      }
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&S.xfull[t & 1], (uint32_t)(nvq * kXBlockBytes));
        for (int q = 0; q < nvq; ++q)
          bulk_g2s(S.xb[t & 1][q], x_img + ((size_t)t * Qtot + (size_t)(q_first + q)) * kXBlockBytes, kXBlockBytes, &S.xfull[t & 1]);
      }
      __syncwarp();
    };
    if (xwarp) fetch_x(0);
    TL_DECL(0u, kTlLoaderEvents)
    TRACE_DECL
    for (int t = FUSED ? 0 : 1; t < T; ++t) {
      const uint32_t par = (uint32_t)((FUSED ? t : t - 1) & 1);
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const int rows = min(kSW, Bv - (grp * kGW + s * kSW));   // valid words of this slot: only their rows travel
        if (rows <= 0) continue;
        uint8_t* bdst = &S.b[s][(size_t)kb * kSW * 128];
        const uint64_t db = make_smem_desc_sw128(smem_u32(bdst));
        if (kb == 3 && lane == 0) TL(1, s, t)   // loader: slot visit starts
        if (t > 0) {
          // probes: lane p < 16 watches writer warp (CTA p>>3 of the k-block's two, lane group (p>>1)&3, column half p&1): its
          // lane 0 publishes row 8 (p&1) of every quarter, units 32 (p>>3) + 8 ((p>>1)&3) of the k-block; quarters are
          // published in order, so the probe sits in the last quarter of the slot that has this row
          // (with two epilogue groups the quarters of a lock-step pair are published concurrently; probing BOTH quarters --
          // 32 probing lanes -- measured slower, 6.3 against 5.6 us at 384 words: the fetch validates every value anyway)
          const int prow = (lane & 1) * 8;
          const int pq = prow < rows ? min(QS - 1, (rows - 1 - prow) / kWq) : 0;
          const uint32_t probe_off = (uint32_t)(((pq * kWq + prow) * 64 + ((lane >> 3) & 1) * 32 + ((lane >> 1) & 3) * 8) * 2);
          const bool prober = !solo && lane < 16 && (kb < kNumKB - 1 || lane < 8) && (prow < rows);
          const uint8_t* src = ll + (size_t)(s * 2 + ((t - 1) & 1)) * kBlk + (size_t)kb * (kSW * 128);
#ifdef PAULE_TC_TRACE
          while ((xchg_load(src) & kPhaseMask) != phase_bits(t - 1)) {}   // split the fetch: until the first value is visible
          TRACE(0)
          uint64_t ftr[2] = {0, 0};
          if (!xchg_fetch_kblock<QS>(src, bdst, kb, phase_bits(t - 1), lane, probe_off, prober, rows, err, ftr, backoff, reread)) break;
          tr_acc[6] += ftr[0] - tr_last;   // probe phase
          tr_acc[7] += ftr[1] * 1000;      // bulk passes (x1000 so that the printout shows passes per step)
#else
          if (!xchg_fetch_kblock<QS>(src, bdst, kb, phase_bits(t - 1), lane, probe_off, prober, rows, err, nullptr, backoff, reread)) break;
#endif
          TRACE(1)
          fence_proxy_async_shared();   // generic-proxy shared-memory writes -> async-proxy (tensor core) reads
          __syncwarp();
          if (kb == 3 && lane == 0) TL(2, s, t)   // loader: k-block fetched
        }
        TRACE(2)
        mbar_wait(&S.acc_free[s], par, err);   // long complete by now: the tile was zeroed a step ago
        if (xwarp) mbar_wait(&S.xfull[t & 1], (uint32_t)((t >> 1) & 1), err);   // x_t landed (issued a step ago)
        tcgen05_fence_after();
        TRACE(3)
        if (elect_one_sync()) {
          const uint32_t d = tmem + (uint32_t)(kV2AccCol + s * kSW);
          if (t > 0) {
            if (nk == 4) {
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16_ts(d, ta + 8 * k, db + 2 * k, idesc, 1u);
            } else {
              umma_bf16_ts(d, ta, db, idesc, 1u);
            }
          }
          if (xwarp) {
            const uint64_t dx = make_smem_desc_sw128(smem_u32(S.xb[t & 1][s * QS]));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ts(d, tmem + (uint32_t)(kXWCol + 8 * k), dx + 2 * k, idesc, 1u);
          }
          umma_commit(&S.mma_done[s]);
        }
        __syncwarp();
        if (kb == 3 && lane == 0) TL(3, s, t)   // loader: MMAs issued
        TRACE(4)
#ifdef PAULE_TC_TRACE
        mbar_wait(&S.mma_done[s], par, err);
        TRACE(5)
#endif
      }
      // x_{t+1}: its buffer was last read by the x MMAs of step t-1, which completed before any h_{t-1} could be fetched
      if (xwarp && t + 1 < T) fetch_x(t + 1);
    }
    if (blockIdx.x == 0 && kb == 3 && lane == 0) TRACE_DUMP(0)
  } else {
    // ===================== epilogue: the LSTM cell =====================
    const int eg = warp >> 3;                       // epilogue group: owns the slots s with s % EG == eg
    const int lg = warp & 3, ch = (warp >> 2) & 1;  // TMEM lane group (32 gate rows = 8 units), column half (8 words)
    const bool tl0 = (warp & 7) == 0 && lane == 0 && eg == 0;   // timeline / trace thread
    const int gq = lane & 3, ul = lane >> 2;        // position in the quad = gate row held after the load; unit in the warp
    const int u = ug * 32 + lg * 8 + ul;
    const bool uvalid = u < kH;
    const int wl0 = ch * 8 + gq * 2;                // this thread's two words inside a quarter: wl0, wl0 + 1
    const uint32_t taddr = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(kV2AccCol + ch * 8);
    // value pair this thread publishes: even units take word 0 of the pair, odd units word 1
    const int e = ul & 1;
    const int prow = wl0 + e;
    const size_t ll_unit = (size_t)((ug & 1) * 32 + lg * 8 + (ul & ~1)) * 2;   // byte offset of the unit pair inside a row
    const bool b1 = (gq & 2) != 0, b0 = (gq & 1) != 0;
    float c_prev[NS * QS][2];
#pragma unroll
    for (int q = 0; q < NS * QS; ++q) c_prev[q][0] = c_prev[q][1] = 0.f;
    float bias4[4] = {0.f, 0.f, 0.f, 0.f};   // FUSED: b_ih + b_hh of this thread's unit, one per gate
    if (FUSED && uvalid)
      for (int g = 0; g < 4; ++g) bias4[g] = __ldg(bias + g * kH + u);

    auto own = [&](int q) { return kByQ ? (q % QS == eg) : ((q / QS) % EG == eg); };   // quarters of this epilogue group
#pragma unroll
    for (int q = 0; q < NS * QS; ++q)
      if (own(q)) tmem_zero_x8(taddr + (uint32_t)(q * kWq));
    tmem_st_wait();
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0)
      for (int s = 0; s < NS; ++s)
        if (kByQ || s % EG == eg) mbar_arrive(&S.acc_free[s]);

    TL_DECL(kTlLoaderEvents, kTlCellEvents)
    TRACE_DECL
    for (int t = 0; t < T; ++t) {
      // input projection of this step (independent of h_{t-1}): quarter 0's loads are issued before the wait, quarter
      // q + 1's while quarter q is computed
      float xp[4][2];
      auto load_xp = [&](int q) {
        if (FUSED) {
#pragma unroll
          for (int g = 0; g < 4; ++g) xp[g][0] = xp[g][1] = bias4[g];
          return;
        }
        if (wf.x_flags != nullptr && grp * kGW + q * kWq < Bv) {
          // layer wavefront: the pre-activations x_t W_ih^T + b are produced while this kernel runs (streaming gate GEMM over the
          // images of the layer below) -- wait until the pair of steps that holds step t has been released for this quarter's
          // 64-word group
          const unsigned int* f = wf.x_flags + (size_t)((w0 + grp * kGW + q * kWq) / kRows) * wf.x_pairs + (t >> 1);
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
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int wp = grp * kGW + q * kWq + wl0 + k;
          const bool ok = uvalid && wp < Bv;
          const float* grow = gates + ((size_t)t * Bs + (ok ? wp : 0)) * (4 * kH) + (uvalid ? u : 0);
#pragma unroll
          for (int g = 0; g < 4; ++g) xp[g][k] = ok ? __ldcg(grow + g * kH) : 0.f;   // L2-coherent: may be written by a co-resident kernel
        }
      };
      load_xp(kByQ ? eg : eg * QS);                // this group's first quarter
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        if (EG > 1 && !kByQ && s % EG != eg) continue;   // the other epilogue group's slot
        if (grp * kGW + s * kSW >= Bv) continue;   // empty slot (uniform over the CTA)
        const bool has_acc = FUSED || t > 0;
        if (tl0) TL(10, s, t)   // epilogue: starts waiting for the slot's accumulator
        if (has_acc) {
          mbar_wait(&S.mma_done[s], (uint32_t)((FUSED ? t : t - 1) & 1), err);
          if (tl0) TL(11, s, t)   // epilogue: accumulator complete
          TRACE(0)
          tcgen05_fence_after();
        }
#pragma unroll
        for (int j = 0; j < QS; ++j) {
          if (kByQ && j != eg) continue;                // the other epilogue group's quarter
          const int q = s * QS + j;                     // quarter inside the CTA
          const int wq = grp * kGW + q * kWq;           // first word of this quarter inside the launch
          float pre[4][2];
          if (has_acc) {
            float acc[8];
            tmem_ld_x8(taddr + (uint32_t)(q * kWq), acc);
            TRACE(1)
            if (t + 1 < T) tmem_zero_x8(taddr + (uint32_t)(q * kWq));   // re-arm: every MMA of the next step adds into it
            if (kByQ || j == QS - 1) {                  // this group's last read of the slot's accumulator
              tmem_st_wait();
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&S.acc_free[s]);
            }
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
            float sl[4][2];   // sl[h] = pair gq of gate (gq ^ h)
#pragma unroll
            for (int x = 0; x < 2; ++x) {
              sl[0][x] = b0 ? k1[x] : k0[x];
              sl[2][x] = b0 ? r1[x] : r0[x];
              sl[1][x] = __shfl_xor_sync(0xffffffffu, b0 ? k0[x] : k1[x], 1);
              sl[3][x] = __shfl_xor_sync(0xffffffffu, b0 ? r0[x] : r1[x], 1);
            }
#pragma unroll
            for (int x = 0; x < 2; ++x) {   // gate g sits in slot g ^ gq
              pre[0][x] = b1 ? (b0 ? sl[3][x] : sl[2][x]) : (b0 ? sl[1][x] : sl[0][x]);
              pre[1][x] = b1 ? (b0 ? sl[2][x] : sl[3][x]) : (b0 ? sl[0][x] : sl[1][x]);
              pre[2][x] = b1 ? (b0 ? sl[1][x] : sl[0][x]) : (b0 ? sl[3][x] : sl[2][x]);
              pre[3][x] = b1 ? (b0 ? sl[0][x] : sl[1][x]) : (b0 ? sl[2][x] : sl[3][x]);
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
          // next quarter's input projection (this group's next one) while this one's results go out
          if (kByQ) { if (s + 1 < NS) load_xp(q + QS); }
          else if (j + 1 < QS) load_xp(q + 1);
          else if (s + EG < NS) load_xp((s + EG) * QS);
          // pair neighbouring units (lane ^ 4) so that one thread owns {h[u], h[u+1]} of one word
          const float other = __shfl_xor_sync(0xffffffffu, e ? hv[0] : hv[1], 4);
          const __nv_bfloat162 pr = e ? __floats2bfloat162_rn(other, hv[1]) : __floats2bfloat162_rn(hv[0], other);
          const uint32_t payload = *reinterpret_cast<const uint32_t*>(&pr);
          const bool pvalid = wq + prow < Bv;
          if (t + 1 < T && pvalid)                      // critical path: the next step's operand
            xchg_store(ll + (size_t)(s * 2 + (t & 1)) * kBlk + ((size_t)((ug >> 1) * kSW + j * kWq + prow) * 64) * 2 + ll_unit,
                       payload | phase_bits(t));
          if (tl0) TL(12, q, t)   // epilogue: quarter published
          TRACE(3)
          // everything below is off the critical path: the next step is already fed
          if (img_seq != nullptr && uvalid && pvalid) {
            const int wg = w0 + wq + prow;              // global word (image addressing)
            *reinterpret_cast<uint32_t*>(img_seq + ((size_t)(wg / kRows) * (size_t)T + t) * kXchgImageBytes +
                                         umma_offset(kRows, wg % kRows, u & ~1)) = payload;
          }
          if (wf.img_flags != nullptr && wq < Bv) {
            // layer wavefront: this warp's image stores of step t are out -- one release-arrival per (step, quarter, warp).
            // BEFORE the stash stores: the release orders everything the warp wrote so far, and the consumer (a streaming GEMM)
            // only reads the images; behind the stash stores it cost the pace-setting layer of a pipeline ~0.3 us per step.
            __syncwarp();
            if (lane == 0) red_release_add_u32(wf.img_flags + (size_t)((w0 + wq) / kRows) * (size_t)T + t, 1u);
          }
#ifndef PAULE_EXPERIMENT_NO_STASH   // timing experiment only (results are then useless to the backward pass)
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
              if (h_out != nullptr) h_out[o] = hv[k];   // NULL: the caller only consumes the bf16 images of h
            }
          }
#endif
          if (tl0) TL(13, q, t)   // epilogue: quarter's stash stores issued
          TRACE(4)
        }
      }
    }
    if (blockIdx.x == 0 && tl0) TRACE_DUMP(8)
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kEpiW) tmem_dealloc<512>(tmem);
}

}  // namespace tc
}  // namespace paule

using namespace paule;
using namespace paule::tc;

namespace paule {
namespace tc {

int pack_v2(const float* w_ih, const float* w_hh, int64_t I, uint8_t* packed, cudaStream_t s) {
  pack_fwd2_kernel<<<kFwd2Groups, 256, 0, s>>>(w_hh, packed + kPackedFwd2Off);
  pack_bwd2_kernel<<<kBwd2Groups * 4, 256, 0, s>>>(w_hh, packed + kPackedBwd2Off);
  if (w_ih != nullptr && I >= 1 && I <= kXK) {   // fused input projection (layers fed by the cps or the mel)
    pack_x_kernel<<<kFwd2Groups, 256, 0, s>>>(w_ih, (int)I, packed + kPackedXOff);
  } else {
    PAULE_CUDA(cudaMemsetAsync(packed + kPackedXOff, 0, (size_t)kFwd2Groups * kXSliceBytes, s));
  }
  PAULE_LAUNCH_CHECK("pack v2 kernels");
  return PAULE_OK;
}

// B = words of the whole batch (row stride of the time-major tensors); the launch covers words [seg0, seg0 + seg_words)
// (seg_words = 0: all of them), in balanced passes of this layout when they exceed one launch
template <int NS, int QS, bool FUSED, int EG>
int launch_fwd2(float* gates, const void* packed, const float* bias, const void* x_img, float* h, float* c, void* xchg,
                void* h_img_seq, int64_t T, int64_t B, cudaStream_t s, WaveFlags wf, int64_t seg0 = 0, int64_t seg_words = 0) {
  static unsigned long long attr_set = 0ull;
  constexpr int NQ = NS * QS;
  const int smem_own = (int)sizeof(Fwd2Smem<NS, QS>) + 1024;
  const int smem = smem_own > kExclusiveSmemBytes ? smem_own : kExclusiveSmemBytes;   // one CTA per SM, whatever runs beside it
  if (once_per_device(attr_set)) {
    PAULE_CUDA(cudaFuncSetAttribute(tc_lstm_fwd2_kernel<NS, QS, FUSED, EG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  const int64_t seg_end = seg_words > 0 ? seg0 + seg_words : B;
  const int64_t gw = (int64_t)kWq * NQ, pw = pass_words(seg_end - seg0, kMaxQ, NQ);
  // words are independent: batches larger than one launch run as consecutive, balanced passes
  for (int64_t r0 = seg0; r0 < seg_end; r0 += pw) {
    const int Bv = (int)((seg_end - r0 < pw) ? (seg_end - r0) : pw);
    const int ng = (int)((Bv + gw - 1) / gw);
    // exchange blocks of this pass start with the phase bit set (0x4040 per value): never a valid first step.  The status
    // word in the header is NOT cleared here: it starts at zero (zero-filled scratch) and stays set once a watchdog fired.
    PAULE_CUDA(cudaMemsetAsync(reinterpret_cast<uint8_t*>(xchg) + kXchgHeader, 0x40, (size_t)ng * 2 * NQ * kLLBlockBytes, s));
    int Ti = (int)T, Bsi = (int)B, w0 = (int)r0, Bvi = Bv, Qtot = (int)((B + kWq - 1) / kWq);
    float* gp = gates + r0 * 4 * kH;
    float* hp = h ? h + r0 * kH : nullptr;
    float* cp = c + r0 * kH;
    uint8_t* xc = reinterpret_cast<uint8_t*>(xchg);
    uint8_t* is = reinterpret_cast<uint8_t*>(h_img_seq);
    const uint8_t* pk = reinterpret_cast<const uint8_t*>(packed) + kPackedFwd2Off;
    const uint8_t* pkx = reinterpret_cast<const uint8_t*>(packed) + kPackedXOff;
    const uint8_t* xi = reinterpret_cast<const uint8_t*>(x_img);
    void* args[] = {&gp, &pk, &hp, &cp, &xc, &is, &Ti, &Bvi, &Bsi, &w0, &pkx, &xi, &bias, &Qtot, &wf};
    PAULE_CUDA(cudaLaunchCooperativeKernel((void*)tc_lstm_fwd2_kernel<NS, QS, FUSED, EG>, dim3(kFwd2Groups * ng),
                                           dim3(f2_threads(EG)), args, (size_t)smem, s));
  }
  return PAULE_OK;
}

// multi-slot layouts run two epilogue groups (PAULE_RNN_EG=1 restores one group: A/B timing).  Measured per step at
// 128 / 192 / 256 / 384 / 1024 words: 3.19 / 3.51 / 4.65 / 6.13 / 18.4 us against 3.46 / 3.83 / 4.59 / 6.48 / 19.1 us with one
// group; four independent slots (4,1) instead of two lock-step pairs (2,2): 6.42 us at 384 words, one lock-step pair (1,2)
// instead of two slots (2,1): 4.40 against 3.23 us at 128 words -- no gain, not built.  Also measured and dropped: reading the
// exchange block once before probing (the loader usually returns to a slot after its data has landed): +20-30 % per step.
// Also measured and dropped (round 2, profiles/r2_fwd3_flag_tma_ab.txt): a flag + TMA exchange for these layouts -- cell warps
// store h_t into the bf16 image, one release store per CTA and quarter publishes it, one issuer warp per slot polls the 23 flag
// words and pulls the operand with twelve bulk copies.  ~60 instead of ~4 000 exchange-side instructions per slot visit, but the
// chain (release fence behind the stash stores -> acquire poll -> bulk copy) is 5 us against 2.2 us and four independent slots
// did not hide it: 9.3 against 5.7 us per step at 384 words.
// Also measured and dropped (profiles/r2_xchg_prefetch_ab.txt): reading the NEXT visit's block one visit ahead into registers
// (values validate themselves, so a speculative read is safe).  The timeline (tools/tc_timeline.py) shows a visit costing 1.7 us,
// but the data lands only ~0.8 us before the visit begins: the speculative read mostly finds the old phase, and its extra loads
// on lines that are being written slow every exchange -- 3.35 -> 4.52 us at 128 words, 5.7 -> 11.5 us at 384.
template <bool FUSED>
int dispatch_fwd2(float* gates, const void* packed, const float* bias, const void* x_img, float* h, float* c, void* xchg,
                  void* h_img_seq, int64_t T, int64_t B, cudaStream_t s, WaveFlags wf, int nq_min, int64_t seg0 = 0,
                  int64_t seg_words = 0) {
  static const bool one_group = getenv("PAULE_RNN_EG") != nullptr && atoi(getenv("PAULE_RNN_EG")) == 1;
  // PAULE_FWD_LAYOUT=<NS><QS><EG> (e.g. 222) forces a layout (A/B timing, tools/ab_groups.sh)
  static const int forced = getenv("PAULE_FWD_LAYOUT") ? atoi(getenv("PAULE_FWD_LAYOUT")) : 0;
#define PAULE_FWD_CASE(NS_, QS_, EG_) \
  return launch_fwd2<NS_, QS_, FUSED, EG_>(gates, packed, bias, x_img, h, c, xchg, h_img_seq, T, B, s, wf, seg0, seg_words)
  // more words than one launch holds: cut into passes by summed step time (tc_lstm.cuh, plan_passes); PAULE_RNN_BALANCED=1 or a
  // forced layout keeps the balanced passes of one layout
  static const bool balanced = getenv("PAULE_RNN_BALANCED") != nullptr || getenv("PAULE_RNN_NQ") != nullptr;
  if (seg_words == 0 && nq_min <= 1 && forced == 0 && !balanced && B > (int64_t)kMaxQ * kWq * 4) {
    PassPlan pp;
    if (plan_passes(B, kMaxQ, kFwdStepUs, &pp)) {
      int64_t r0 = 0;
      for (int i = 0; i < pp.n; ++i) {
        const int rc = dispatch_fwd2<FUSED>(gates, packed, bias, x_img, h, c, xchg, h_img_seq, T, B, s, wf, pp.nq[i], r0, pp.words[i]);
        if (rc != PAULE_OK) return rc;
        r0 += pp.words[i];
      }
      return PAULE_OK;
    }
  }
  switch (nq_min > 1 ? 0 : forced) {
    case 111: PAULE_FWD_CASE(1, 1, 1);
    case 211: PAULE_FWD_CASE(2, 1, 1);
    case 212: PAULE_FWD_CASE(2, 1, 2);
    case 311: PAULE_FWD_CASE(3, 1, 1);
    case 221: PAULE_FWD_CASE(2, 2, 1);
    case 222: PAULE_FWD_CASE(2, 2, 2);
    default: break;
  }
  int nq = choose_nq(seg_words > 0 ? seg_words : B, kMaxQ);
  if (seg_words > 0) nq = nq_min;    // a pass of the plan: its layout was chosen there
  if (nq < nq_min) nq = nq_min;      // a co-resident kernel (layer wavefront) leaves this one fewer SMs: more quarters per CTA
  if (nq == 1) PAULE_FWD_CASE(1, 1, 1);
  if (one_group) {
    if (nq == 2) PAULE_FWD_CASE(2, 1, 1);
    if (nq == 3) PAULE_FWD_CASE(3, 1, 1);
    PAULE_FWD_CASE(2, 2, 1);
  }
  if (nq == 2) PAULE_FWD_CASE(2, 1, 2);
  if (nq == 3) PAULE_FWD_CASE(3, 1, 1);   // 4.52 vs 4.61 us with two groups
  PAULE_FWD_CASE(2, 2, 2);
#undef PAULE_FWD_CASE
}

// quarters per CTA (1..4) such that B words fit ONE launch of at most max_ctas CTAs (0: the device); 0 = impossible
static int nq_for(int64_t B, int max_ctas) {
  const int64_t quarters = (B + kWq - 1) / kWq;
  for (int nq = 1; nq <= 4; ++nq) {
    const int64_t groups = (quarters + nq - 1) / nq;
    if (groups <= kMaxQ && (max_ctas <= 0 || groups * kFwd2Groups <= max_ctas)) return nq;
  }
  return 0;
}

int fwd2_pass_plan(int64_t B, int32_t* nq_out, int32_t* words_out, int cap_out) {
  static const bool balanced = getenv("PAULE_RNN_BALANCED") != nullptr || getenv("PAULE_RNN_NQ") != nullptr ||
                               getenv("PAULE_FWD_LAYOUT") != nullptr;
  const int64_t cap = (int64_t)kMaxQ * kWq * 4;
  PassPlan pp;
  if (!(B > cap && !balanced && plan_passes(B, kMaxQ, kFwdStepUs, &pp))) {   // one layout, balanced passes
    const int nq = choose_nq(B, kMaxQ);
    const int64_t pw = pass_words(B, kMaxQ, nq);
    pp.n = 0;
    for (int64_t r0 = 0; r0 < B && pp.n < 96; r0 += pw) { pp.nq[pp.n] = nq; pp.words[pp.n] = (int)(B - r0 < pw ? B - r0 : pw); ++pp.n; }
  }
  for (int i = 0; i < pp.n && i < cap_out; ++i) { nq_out[i] = pp.nq[i]; words_out[i] = pp.words[i]; }
  return pp.n;
}
int fwd2_passes(int64_t B) { return fwd2_pass_plan(B, nullptr, nullptr, 0); }

int fwd2_ctas(int64_t B, int max_ctas) {
  const int nq = nq_for(B, max_ctas);
  if (nq == 0) return 0;
  const int64_t quarters = (B + kWq - 1) / kWq;
  return (int)((quarters + nq - 1) / nq) * kFwd2Groups;
}

int lstm_seq_fwd2(float* gates, const void* packed, float* h, float* c, void* xchg, void* h_img_seq, int64_t T, int64_t B,
                  cudaStream_t s, WaveFlags wf, int max_ctas) {
  int nq_min = 1;
  if (max_ctas > 0) {
    nq_min = nq_for(B, max_ctas);
    if (nq_min == 0) return PAULE_ERR_ARG;
  }
  return dispatch_fwd2<false>(gates, packed, nullptr, nullptr, h, c, xchg, h_img_seq, T, B, s, wf, nq_min);
}

int lstm_seq_fwd2x(float* gates, const void* packed, const float* bias, const void* x_img, float* h, float* c, void* xchg,
                   void* h_img_seq, int64_t T, int64_t B, cudaStream_t s, WaveFlags wf, int max_ctas) {
  int nq_min = 1;
  if (max_ctas > 0) {
    nq_min = nq_for(B, max_ctas);
    if (nq_min == 0) return PAULE_ERR_ARG;
  }
  return dispatch_fwd2<true>(gates, packed, bias, x_img, h, c, xchg, h_img_seq, T, B, s, wf, nq_min);
}

int x_image(const float* x, void* img, int64_t T, int64_t B, int64_t I, cudaStream_t s) {
  const int64_t total = T * B * I;
  if (total == 0) return PAULE_OK;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  x_image_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, reinterpret_cast<uint8_t*>(img), T, B, (int)I);
  PAULE_LAUNCH_CHECK("x_image_kernel");
  return PAULE_OK;
}

}  // namespace tc
}  // namespace paule
