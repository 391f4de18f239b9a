// Shapes of the H=720 persistent-RNN kernels (forward: tc_lstm_fwd2.cu, backward: tc_lstm_bwd2.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#ifdef PAULE_TC_TRACE
#define TRACE_DECL uint64_t tr_last = globaltimer_ns(); uint64_t tr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define TRACE(i) { const uint64_t _n = globaltimer_ns(); tr_acc[i] += _n - tr_last; tr_last = _n; }
#define TRACE_DUMP(base) { uint64_t* _o = reinterpret_cast<uint64_t*>(xchg + kXchgTraceOff) + (base); for (int _i = 0; _i < 8; ++_i) _o[_i] = tr_acc[_i]; }
#else
#define TRACE_DECL
#define TRACE(i) {}
#define TRACE_DUMP(base) {}
#endif

// -DPAULE_TC_TIMELINE: two threads of CTA 0 (loader warp 3's lane 0, the first cell warp's lane 0) record (event, slot, step,
// globaltimer) tuples of steps [kTlFirst, kTlFirst + kTlSteps) into the free part of the exchange header (tools/tc_timeline.py
// prints them as a timeline).  Each thread counts its own events in a register and owns a range of the 64 slots: plain
// fire-and-forget stores -- an atomic slot counter costs the recording thread an L2 round trip (~0.45 us) per event, which
// distorted the very chain it measured.  Debug builds only.
#ifdef PAULE_TC_TIMELINE
#define TL_DECL(base_, max_) unsigned int tl_n = 0u; const unsigned int tl_base = (base_), tl_max = (max_);
#define TL(ev, slot, step)                                                                                      \
  {                                                                                                             \
    if (blockIdx.x == 0 && (step) >= kTlFirst && (step) < kTlFirst + kTlSteps && !kTlSkip(ev) && tl_n < tl_max) { \
      uint64_t* _o = reinterpret_cast<uint64_t*>(xchg + kXchgTlOff) + 2 * (size_t)(tl_base + tl_n);               \
      _o[0] = ((uint64_t)(ev) << 32) | ((uint64_t)(slot) << 16) | (uint64_t)(step);                              \
      _o[1] = globaltimer_ns();                                                                                  \
      ++tl_n;                                                                                                    \
    }                                                                                                            \
  }
#else
#define TL_DECL(base_, max_)
#define TL(ev, slot, step) {}
#endif

namespace paule {
namespace tc {
#ifndef PAULE_TL_STEPS
#define PAULE_TL_STEPS 3
#endif
constexpr int kTlFirst = 60, kTlSteps = PAULE_TL_STEPS, kXchgTlOff = 3072;   // = kXchgTraceOff (1 KB: 64 events)
constexpr unsigned int kTlLoaderEvents = 26, kTlCellEvents = 38;              // slots of the two recording threads
// -DPAULE_TL_BRIEF: drop the "starts waiting" / "stash stores issued" events (four-quarter layouts: 62 events are two steps)
#ifdef PAULE_TL_BRIEF
__host__ __device__ constexpr bool kTlSkip(int ev) { return ev == 10 || ev == 13; }
#else
__host__ __device__ constexpr bool kTlSkip(int) { return false; }
#endif

constexpr int kH = 720;                 // hidden size the tensor-core path is built for (paule/paule.py:124,167)
constexpr int kKPad = 768;              // K padded to 12 k-blocks of 64
constexpr int kNumKB = kKPad / 64;      // 12
constexpr int kRows = 64;               // words per image group (rows of a bf16 UMMA image, the batched GEMMs' half M tile)

// exchange buffer: header + exchange blocks
constexpr int kXchgHeader = 4096;   // status word, trace / timeline words
constexpr int kXchgErrOff = 2048;    // int: 0 ok, 1 exchange watchdog fired, 2 mbarrier watchdog fired (every wait loop of the
                                     // kernels bails out once this word is non-zero: results are invalid)
constexpr int kXchgClampOff = 2052;  // int, sticky and informational (must NOT stop the kernels): 1 = the BPTT kernel clamped a
                                     // recurrent gradient to the +-1.5 bound of the in-band exchange, or it was NaN / Inf
constexpr int kXchgTraceOff = 3072;  // PAULE_TC_TRACE builds only
constexpr int kXchgImageBytes = kNumKB * kRows * 128;  // 98304: one bf16 image [12 kb][64 rows][128 B]

// ---- v2 persistent kernels (tc_lstm_fwd2.cu / tc_lstm_bwd2.cu): weights are the resident A operand (M = 128 rows),
// 16 words are the N dimension, and the recurrent state travels between CTAs as 8-byte {bf16x2, step tag} elements
// that carry their own validity (no counters, no release/acquire chain, no TMA on the exchange).
constexpr int kWq = 16;                              // words per CTA ("word quarter" of a 64-word group)
constexpr int kMaxQ = 6;                             // word quarters per forward launch: 23 x 6 = 138 CTAs <= 148 SMs
constexpr int kMaxQBwd = 5;                          // backward: clusters of 4 -> at most 132 co-resident CTAs; 24 x 5 = 120
constexpr int kV2M = 128;                            // UMMA M: gate rows (forward) / hidden units (backward) per CTA
constexpr int kV2SliceBytes = kNumKB * kV2M * 128;   // 196608: resident A operand [128, 768] bf16 = 384 TMEM columns
constexpr int kV2WCols = kKPad / 2;                  // 384 32-bit TMEM columns (two bf16 per column)
constexpr int kV2AccCol = 0;                         // accumulator [128 lanes, 16 columns]
constexpr int kV2WCol = 64;                          // first column of the resident weights
constexpr int kV2BBytes = kNumKB * kWq * 128;        // 24576: B operand [16 words, 768] bf16
constexpr int kFwd2Groups = (kH + 31) / 32;          // 23 CTAs per word quarter, 32 hidden units each
constexpr int kBwd2Groups = (kH + 127) / 128;        // 6 clusters of 4 (one CTA per gate) per word quarter, 128 units each
// exchange block per (quarter, parity[, gate]): dense bf16 [12 kb][16 words][64 units], phase bit in bit 14 of every value
constexpr int kLLBlockBytes = kNumKB * kWq * 128;     // 24576
constexpr size_t kLLBytes = (size_t)kMaxQ * 4 * 2 * 4 * kLLBlockBytes;   // up to 6 groups x 4 quarters; backward: 4 gate images per parity

// offsets of the v2 weight images inside the packed buffer of one layer
constexpr size_t kPackedFwd2Off = 0;
constexpr size_t kPackedBwd2Off = kPackedFwd2Off + (size_t)kFwd2Groups * kV2SliceBytes;
// fused input projection (K = input_size <= 64 padded to 64): per forward CTA a [128 gate rows, 64] bf16 slice = 32 more
// TMEM columns behind the W_hh slice, stored in the same tcgen05.st order [4 column octets][128 rows][8 x u32]
constexpr int kXK = 64;                              // K of the fused input projection (one k-block)
constexpr int kXCols = kXK / 2;                      // 32 TMEM columns
constexpr int kXWCol = kV2WCol + kV2WCols;           // 448: first column of the input-projection weights
constexpr int kXSliceBytes = kV2M * kXK * 2;         // 16384
constexpr int kXBlockBytes = kWq * 128;              // 2048: x_t of one word quarter as a B operand [16 rows][128 B]
constexpr size_t kPackedXOff = kPackedBwd2Off + (size_t)kBwd2Groups * 4 * kV2SliceBytes;
constexpr size_t kPackedBytes = kPackedXOff + (size_t)kFwd2Groups * kXSliceBytes;


// quarters per CTA for a batch of B words when a launch holds at most max_groups groups: the smallest of 1..4 that
// fits the batch into one launch (latency first, and the most CTAs), 4 beyond that.  PAULE_RNN_NQ overrides (testing).
inline int choose_nq(int64_t B, int max_groups) {
  static const int forced = getenv("PAULE_RNN_NQ") ? atoi(getenv("PAULE_RNN_NQ")) : 0;
  if (forced >= 1 && forced <= 4) return forced;
  const int64_t cap = (int64_t)max_groups * kWq * 4;          // most words one launch can hold
  const int64_t n_pass = (B + cap - 1) / cap;
  const int64_t per = (B + n_pass - 1) / n_pass;              // words per (balanced) pass
  for (int nq = 1; nq < 4; ++nq)
    if (per <= (int64_t)max_groups * kWq * nq) return nq;
  return 4;
}
// words per launch when B words are split into the fewest passes of at most max_groups groups of 16 nq words, balanced
inline int64_t pass_words(int64_t B, int max_groups, int nq) {
  const int64_t gw = (int64_t)kWq * nq, cap = (int64_t)max_groups * gw;
  const int64_t n_pass = (B + cap - 1) / cap;
  const int64_t per = (B + n_pass - 1) / n_pass;
  return (per + gw - 1) / gw * gw;
}

// Batches beyond one launch: the words are independent, so the batch is cut into passes, each with its own quarters per CTA.
// A multi-quarter layout's step time does not depend on how many of its word groups are occupied, so
// balancing the words over the passes wastes groups: 1024 words through the BPTT kernel are 4 balanced passes of 256 words at
// 4 quarters per CTA (4 x 7.05 us per step, 4 of 5 groups used) -- or 4 full passes of 240 words at 3 quarters plus one
// latency-layout pass of 64 (4 x 5.2 + 2.85 us).  plan_passes minimises the summed step time over the passes with a small
// dynamic program over word quarters; cost[nq] = measured us per step of the layout with nq quarters per CTA.
struct PassPlan {
  int n = 0;
  int nq[96];       // quarters per CTA of pass i
  int words[96];    // words of pass i (multiples of 16 except the last)
};
inline bool plan_passes(int64_t B, int max_groups, const float* cost, PassPlan* out) {
  const int64_t Q64 = (B + kWq - 1) / kWq;
  if (Q64 > 2048) return false;
  const int Q = (int)Q64;
  static thread_local float best[2049];
  static thread_local unsigned char pick[2049];
  best[0] = 0.f;
  for (int q = 1; q <= Q; ++q) {
    best[q] = 1e30f;
    for (int nq = 1; nq <= 4; ++nq) {
      const int cap = max_groups * nq, rest = q > cap ? q - cap : 0;
      const float c = best[rest] + cost[nq] + 1e-3f * (float)nq;   // ties: the smaller layout
      if (c < best[q]) { best[q] = c; pick[q] = (unsigned char)nq; }
    }
  }
  out->n = 0;
  int64_t left = B;
  for (int q = Q; q > 0;) {
    if (out->n >= 96) return false;
    const int nq = pick[q], cap = max_groups * nq, take = q > cap ? cap : q;
    const int64_t w = (int64_t)take * kWq < left ? (int64_t)take * kWq : left;
    out->nq[out->n] = nq;
    out->words[out->n] = (int)w;
    ++out->n;
    left -= w;
    q -= take;
  }
  return true;
}
// measured us per step by quarters per CTA (tools/rnn_time.py, B200): forward (1,1) / (2,1) / (3,1) / (2,2), backward NQ = 1..4
// at the layouts' full capacity (96 / 192 / 288 / 384 and 80 / 160 / 240 / 320 words)
constexpr float kFwdStepUs[5] = {0.f, 2.23f, 3.64f, 4.71f, 5.79f};
constexpr float kBwdStepUs[5] = {0.f, 2.87f, 3.67f, 5.00f, 6.45f};

// Probe polling variants, measured at 64 words (forward step 2.42 us with the defaults): two probe loads in flight 2.75 us
// (more polling traffic slows every exchange), a 40 / 120 ns back-off between failed probes 2.39 / 2.41 us (neutral).
// Later A/B (same box, two builds each): 300 ns -> forward 2.27 -> 2.19 us at 64 words, but 1.74 -> 1.83 us at one word and
// nothing in the backward or the multi-slot layouts; 600 ns is worse everywhere.  The forward kernel therefore backs off
// 300 ns only when the launch has at least four word quarters polling (run-time argument of xchg_fetch_kblock).
// Forward launches of at most four words (one quarter, 23 CTAs, one 16-byte load per lane per k-block) poll with the block
// read itself (prober = false for every lane, 300 ns between incomplete reads): the probe round trip disappears from the
// chain -- 1.74 -> 1.63 us per step at one word.  With a full quarter (16 words) the same is 4 % slower (2.14 -> 2.22 us), at
// 64 words 2.19 -> 2.55-2.74 us; in the backward kernel it is within noise (2.03 -> 2.00 us) and not used.
// Not polling at all for the first 400 / 700 / 1000 ns after the warp's own MMA issue (nothing can arrive sooner) changed
// neither kernel (2.16-2.20 / 2.82-2.87 us): the polls that matter are the ones in flight when the data lands.
#ifndef PAULE_PROBE_PIPELINE
#define PAULE_PROBE_PIPELINE 0
#endif
#ifndef PAULE_PROBE_BACKOFF_NS
#define PAULE_PROBE_BACKOFF_NS 0
#endif

#ifdef __CUDACC__
// Weight image of one CTA as tcgen05.st wants it: [48 column octets][128 rows][8 x u32]; u32 column c of row m holds the
// bf16 pair (k = 2c, 2c+1).  Warp w (lanes 32w..32w+31) copies its rows with one 32-byte load + one x8 store per octet.
__device__ __forceinline__ void load_weights_to_tmem(const uint8_t* __restrict__ slice, uint32_t tmem, int warp, int lane,
                                                     int first_col = kV2WCol, int n_cols = kV2WCols) {
  const uint4* src = reinterpret_cast<const uint4*>(slice) + (size_t)(warp * 32 + lane) * 2;
  const uint32_t dst = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)first_col;
#pragma unroll 4
  for (int c8 = 0; c8 < n_cols / 8; ++c8) {
    const uint4 a = __ldg(src + (size_t)c8 * kV2M * 2), b = __ldg(src + (size_t)c8 * kV2M * 2 + 1);
    tmem_st_x8(dst + (uint32_t)(c8 * 8), a, b);
  }
  tmem_st_wait();
}

// Pulls k-block kb of a [16 NQ words x 768] bf16 operand ([16 NQ rows][128 B]) out of an exchange block into the swizzled
// UMMA layout at `dst`, waiting until every value carries phase bit `phase`.  Called by one full warp.  Lanes 0..15
// first poll ONE value of each of the 16 writer warps of the k-block (`probe_off`: byte offset of this lane's probe; a
// warp's values leave the SM with one store instruction, so they become visible together), then the block is read once
// with 16-byte loads (all in flight together); stragglers are re-read.  Only the first `rows` rows (valid words) are
// published and read.  Returns false when the watchdog fired.
template <int NQ>
__device__ __forceinline__ bool xchg_fetch_kblock(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int kb,
                                                  uint32_t phase, int lane, uint32_t probe_off, bool prober,
                                                  int rows, volatile int* err, uint64_t* trace = nullptr,
                                                  unsigned int backoff_ns = 0u, unsigned int reread_ns = 0u) {
  // k-block 11 holds units 704..767: only 704..735 have a writer (chunks 0..3), the rest stays zero in shared memory
  const int c = lane & 7;
  const bool active = (kb < kNumKB - 1) || (c < 4);
  uint32_t pending = 0u;
#pragma unroll
  for (int i = 0; i < 4 * NQ; ++i) pending |= (4 * i < rows) ? (1u << i) : 0u;
  {
    const uint8_t* pp = src + probe_off;
    uint64_t t0 = 0;
    bool ok = !prober;
#if PAULE_PROBE_PIPELINE
    // two probe loads in flight: the second is issued before the first is examined, which halves the sampling period
    uint32_t nxt = ok ? 0u : xchg_load(pp);
    for (unsigned int spin = 0;; ++spin) {
      const uint32_t cur = nxt;
      if (!ok) nxt = xchg_load(pp);
      if (!ok) ok = (cur & kPhaseMask) == phase;
      if (__all_sync(0xffffffffu, ok)) break;
#else
    for (unsigned int spin = 0;; ++spin) {
      if (!ok) ok = (xchg_load(pp) & kPhaseMask) == phase;
      if (__all_sync(0xffffffffu, ok)) break;
#if PAULE_PROBE_BACKOFF_NS > 0
      __nanosleep(PAULE_PROBE_BACKOFF_NS);   // fewer polls in flight -> shorter L2 queues for the loads that matter
#else
      if (backoff_ns) __nanosleep(backoff_ns);   // caller's choice (see tc_lstm_fwd2.cu)
#endif
#endif
      if ((spin & 1023u) == 1023u) {
        if (t0 == 0) t0 = globaltimer_ns();
        if (*err != 0) return false;
        if (globaltimer_ns() - t0 > kWatchdogNs) { *err = 1; return false; }
      }
    }
  }
  uint64_t t0 = 0;
  if (trace) { trace[0] = globaltimer_ns(); }
  for (unsigned int spin = 0; pending != 0u; ++spin) {
    if (trace) trace[1] += 1;
#if PAULE_XCHG_LOAD == 2
    // asynchronous 16-byte copies straight to the swizzled position, then validate + clear the phase bits in place
#pragma unroll
    for (int i = 0; i < 4 * NQ; ++i) {
      const int row = 4 * i + (lane >> 3);
      if (((pending >> i) & 1u) && active && row < rows)
        cp_async_cg16(dst + (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4), src + (size_t)(i * 32 + lane) * 16);
    }
    cp_async_wait_all();
#pragma unroll
    for (int i = 0; i < 4 * NQ; ++i) {
      if ((pending >> i) & 1u) {
        const int row = 4 * i + (lane >> 3);
        const bool mine = active && row < rows;
        uint4* sp = reinterpret_cast<uint4*>(dst + (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4));
        uint4 v = make_uint4(phase, phase, phase, phase);
        if (mine) v = *sp;
        const bool ok = ((v.x & kPhaseMask) == phase) && ((v.y & kPhaseMask) == phase) && ((v.z & kPhaseMask) == phase) &&
                        ((v.w & kPhaseMask) == phase);
        if (__all_sync(0xffffffffu, ok)) {
          pending &= ~(1u << i);
          if (mine) *sp = make_uint4(v.x & ~kPhaseMask, v.y & ~kPhaseMask, v.z & ~kPhaseMask, v.w & ~kPhaseMask);
        }
      }
    }
#else
    uint4 v[4 * NQ];
#pragma unroll
    for (int i = 0; i < 4 * NQ; ++i)
      if (((pending >> i) & 1u) && active && 4 * i + (lane >> 3) < rows) v[i] = xchg_load4(src + (size_t)(i * 32 + lane) * 16);
#pragma unroll
    for (int i = 0; i < 4 * NQ; ++i) {
      if ((pending >> i) & 1u) {
        const int row = 4 * i + (lane >> 3);
        const bool ok = !active || row >= rows || (((v[i].x & kPhaseMask) == phase) && ((v[i].y & kPhaseMask) == phase) &&
                                                   ((v[i].z & kPhaseMask) == phase) && ((v[i].w & kPhaseMask) == phase));
        if (__all_sync(0xffffffffu, ok)) {
          pending &= ~(1u << i);
          if (active && row < rows) {
            *reinterpret_cast<uint4*>(dst + (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4)) =
                make_uint4(v[i].x & ~kPhaseMask, v[i].y & ~kPhaseMask, v[i].z & ~kPhaseMask, v[i].w & ~kPhaseMask);
          }
        }
      }
    }
#endif
    if (reread_ns && pending != 0u) __nanosleep(reread_ns);   // probe-less polling (single-quarter launches): pace the re-reads
    if ((spin & 255u) == 255u) {
      if (t0 == 0) t0 = globaltimer_ns();
      if (*err != 0) return false;
      if (globaltimer_ns() - t0 > kWatchdogNs) { *err = 1; return false; }
    }
  }
  return true;
}

// words per CTA group (16 NQ) for a batch of B words: small batches keep one quarter per CTA (lowest latency), large
// ones put 2 or 4 quarters on a CTA (the grid is limited to 6 / 5 groups, so this is what scales the throughput)
#endif  // __CUDACC__

// ---- layer wavefront: per-step arrival counters next to the bf16 images.  A producing recurrent kernel
// increments `img_flags[(word group) * T + t]` (release) from every epilogue warp once its stores of step t are out; a
// consuming kernel / streaming GEMM polls them (acquire).  A consuming recurrent kernel with a fused input projection waits
// for `x_flags[(word group) * x_pairs + t / 2] >= 4` (the streaming GEMM's four epilogue warps) before it fetches x_t.
struct WaveFlags {
  unsigned int* img_flags;        // producer side, or nullptr
  const unsigned int* x_flags;    // consumer side, or nullptr
  int x_pairs;                    // pairs of steps per word group in x_flags
  unsigned int x_target;          // arrivals that complete a pair (4 epilogue warps x column tiles of the streaming GEMM)
};
constexpr WaveFlags kNoWave = {nullptr, nullptr, 0, 0u};
constexpr int kWaveArrivalsPerQuarter = kFwd2Groups * 8;   // 23 unit groups x 8 epilogue warps per (step, word quarter)
// Kernels that run side by side in the wavefront all allocate the SM's whole tensor memory: two of their CTAs on one SM would
// park the second in tcgen05.alloc until the first exits -- a deadlock when the first (transitively) waits for the second.
// Requesting more than half of the SM's shared memory makes every such CTA the only one on its SM.
constexpr int kExclusiveSmemBytes = 117 * 1024;

// host entry points of the kernels (called from the C ABI in tc_lstm_api.cu)
int pack_v2(const float* w_ih, const float* w_hh, int64_t I, uint8_t* packed, cudaStream_t s);
int x_image(const float* x, void* img, int64_t T, int64_t B, int64_t I, cudaStream_t s);
int lstm_seq_fwd2(float* gates, const void* packed, float* h, float* c, void* xchg, void* h_img_seq, int64_t T, int64_t B,
                  cudaStream_t s, WaveFlags wf = kNoWave, int max_ctas = 0);
// fused input projection: gates is output only (the activated-gate stash), pre-activations = W_hh h + W_ih x_t + bias
int lstm_seq_fwd2x(float* gates, const void* packed, const float* bias, const void* x_img, float* h, float* c, void* xchg,
                   void* h_img_seq, int64_t T, int64_t B, cudaStream_t s, WaveFlags wf = kNoWave, int max_ctas = 0);
// CTAs lstm_seq_fwd2x launches for B words when at most max_ctas (0: no limit) may be used; 0 = does not fit one launch
int fwd2_ctas(int64_t B, int max_ctas);
// launches of one recurrent layer over B words (the pass plan above; 1 up to 384 / 320 words)
int fwd2_passes(int64_t B);
int bwd2_passes(int64_t B);
// the passes themselves (at most cap entries written; returns their number)
int fwd2_pass_plan(int64_t B, int32_t* nq_out, int32_t* words_out, int cap);
int bwd2_pass_plan(int64_t B, int32_t* nq_out, int32_t* words_out, int cap);
// batched tcgen05 GEMM over image sequences (tc_gemm.cu): batch mode with a status word, streaming mode for the wavefront
int gemm_img(const void* a_img, const void* packed_b, const float* bias, float* C, int64_t steps, int64_t B, int64_t N,
             int64_t nseg, int accumulate, int* status, cudaStream_t s);
int gemm_img_kb(const void* a_img, const void* packed_b, const float* bias, float* C, int64_t steps, int64_t B, int64_t N, int KB,
                int accumulate, int* status, cudaStream_t s);
int gemm_img_stream(const void* a_img, const void* packed_b, const float* bias, float* C, int64_t steps, int64_t B, int64_t N,
                    int64_t nseg, const unsigned int* src_flags, const unsigned int* src_target, int src_per_step,
                    unsigned int* dst_flags, void* x_out, int n_par, int* status, cudaStream_t s, int reverse = 0,
                    int accumulate = 0, int n_ct = 1);
int gemm_stream_ctas(int64_t B, int64_t N, int n_par, int n_ct = 1);
unsigned int gemm_stream_arrivals(int64_t N);
int lstm_seq_bwd2(float* gates, const float* c, const void* packed, const float* dh_seq, int dh_mode, const float* dh_last,
                  void* xchg, void* da_img_seq, int64_t T, int64_t B, int keep_da, cudaStream_t s, WaveFlags wf = kNoWave,
                  int force_nq = 0);
int bwd2_ctas(int64_t B, int nq);
int bwd2_default_nq(int64_t B);
constexpr int kWaveArrivalsPerQuarterBwd = kBwd2Groups * 4 * 8;   // 24 CTAs x 8 cell warps per (step, word quarter)

}  // namespace tc
}  // namespace paule
