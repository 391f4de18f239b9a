// Batched gate GEMMs on tcgen05 (sm_100a): the "GEMMs over all timesteps" of the planning step.
//
//   C[(t, b), n] (+)= sum_k A[(t, b), k] * W[n, k] + bias[n]          M = steps x words, fp32 output
//
//   * embedder layer-1 input projection   Xp1 = h0 W_ih1^T + b      (K = 720,  N = 2880)     [K5 of SURVEY 2.2]
//   * BPTT input gradients                dX  = dA  W_ih            (K = 2880, N = 720/60/30) [K8]
//
//   A is never re-laid-out: the persistent-RNN kernels already wrote h_t (forward) / dA_t (backward) as bf16 UMMA
//   images, one per time step and 64-word group ([k-blocks][64 rows][128 B], SWIZZLE_128B).  An M = 128 tile is two
//   consecutive time steps of one group, fetched with two 8 KB TMA bulk copies per 64-wide k-block; W comes from a
//   pre-packed bf16 image.  Warp-specialised, persistent over output tiles:
//     warp 0  TMA producer  (4-stage ring of {A 16 KB, B BN x 128 B}, full/empty mbarriers)
//     warp 1  MMA issuer    (tcgen05.mma M=128, N=BN, K=16; tcgen05.commit frees the stage / publishes the accumulator)
//     warps 2-9 epilogue    (tcgen05.ld 32 columns at a time, 32 x 32 transpose through shared memory, + bias, fp32 row
//                           segments of 128 B to HBM; two warps per TMEM lane group take alternate column chunks -- with four
//                           epilogue warps the K = 720 projection was bound by the epilogue, 7.4 us per tile against 2.8 us of MMA)
//   with two TMEM accumulator buffers (2 x 256 columns) so that the epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_lstm.cuh"

namespace paule {
namespace tc {

constexpr int kGemmStages = 4;
constexpr int kGemmEpiWarps = 8;                 // two per TMEM lane group: they split the 32-column chunks of a tile
constexpr int kGemmThreads = 32 * (2 + kGemmEpiWarps);
constexpr int kSegLen = kH;          // real columns per K segment
constexpr int kSegPad = kKPad;       // padded columns per K segment (12 k-blocks)

// largest tile width <= 256 that is a multiple of 16 and divides the padded N -- except N = 720 (dX of embedder layer 1):
// three 240-wide column tiles x 50 row tiles (cfg 2) are 150 tiles on 148 CTAs, i.e. two rounds for two tiles; five 144-wide
// ones are 250 tiles = two rounds of a 40 % cheaper tile.  At large M the narrower tile re-reads A more often (+13 % on
// that one GEMM at 256 words x 400 frames, 0.2 % of the step): the small-batch case wins.  N = 2880 (input projection of
// embedder layer 1): 160-wide tiles measured fastest at both ends.
__host__ __device__ inline int pick_bn(int n_pad) {
  if (n_pad == 720) return 144;
  if (n_pad == 2880) {
#ifndef __CUDA_ARCH__
    static const int forced = getenv("PAULE_GEMM_BN") ? atoi(getenv("PAULE_GEMM_BN")) : 0;   // A/B timing
    if (forced >= 16 && forced <= 240 && forced % 16 == 0 && 2880 % forced == 0) return forced;
#endif
    return 160;
  }
  // N = 2880, single-CTA kernel, measured (tools/gemm_time.py, 64 / 256 words): 240: 56.3 / 172 us, 192: 51.9 / 161, 160: 50.3 / 159, 144: 52.1 / 172
  for (int bn = 240; bn >= 16; bn -= 16)   // 240: four stages + the eight epilogue staging tiles fit the 227 KB of an SM
    if (n_pad % bn == 0) return bn;
  return 16;
}
__host__ __device__ inline int pad_n(int n) { return (n + 15) / 16 * 16; }

// W [N, nseg*seg_len] fp32 row-major -> image [n_tile][kb (nseg * seg_pad / 64)][BN rows][128 B], zero padded
// (seg_len = 720, seg_pad = 768: the hidden-size segments; seg_len = K <= 64, seg_pad = 64: one k-block, the narrow-K GEMM)
__global__ void pack_gemm_b_kernel(const float* __restrict__ W, uint8_t* __restrict__ img, int N, int nseg, int BN, int seg_len,
                                   int seg_pad) {
  const int KB = nseg * seg_pad / 64;
  const int nt = blockIdx.x;
  const size_t tile_bytes = (size_t)KB * BN * 128;
  for (int e = threadIdx.x; e < BN * nseg * seg_pad; e += blockDim.x) {
    const int r = e / (nseg * seg_pad), kp = e % (nseg * seg_pad);
    const int seg = kp / seg_pad, ks = kp % seg_pad;
    const int n = nt * BN + r;
    const float v = (n < N && ks < seg_len) ? W[(size_t)n * (nseg * seg_len) + seg * seg_len + ks] : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(img + (size_t)nt * tile_bytes + umma_offset(BN, r, kp)) = __float2bfloat16_rn(v);
  }
}

// x [steps, B, I <= 64] fp32 -> A operand images of a narrow-K GEMM: one k-block per (64-word group, step),
// [grp][step][64 rows][128 B] bf16, SWIZZLE_128B, columns >= I zero.  A thread converts one 16-byte chunk (8 columns) of a row.
__global__ void a_image_kernel(const float* __restrict__ x, uint8_t* __restrict__ img, int64_t steps, int64_t B, int I) {
  const int64_t total = steps * B * 8;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e & 7);
    const int64_t tb = e >> 3, b = tb % B, t = tb / B;
    const float* src = x + tb * I + 8 * c;
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = 8 * c + 2 * j;
      const __nv_bfloat162 pr = __floats2bfloat162_rn(k < I ? __ldg(src + 2 * j) : 0.f, k + 1 < I ? __ldg(src + 2 * j + 1) : 0.f);
      w[j] = *reinterpret_cast<const uint32_t*>(&pr);
    }
    const int row = (int)(b % kRows);
    uint8_t* dst = img + ((size_t)(b / kRows) * (size_t)steps + (size_t)t) * (kRows * 128) + (size_t)row * 128 + ((c ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

constexpr int kStageLd = 36;   // floats per staged row: 32 + 4 (16-byte aligned rows, conflict-free float4 phases)
struct GemmBars {
  uint64_t full[kGemmStages];
  uint64_t empty[kGemmStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  // epilogue staging, one 32 x 32 tile per warp: a thread owns one accumulator ROW after tcgen05.ld, but rows of C are
  // N floats apart -- storing from registers makes every warp store touch 32 lines at 16 bytes each (1 TB/s on the K = 720
  // projection).  Through this transpose every store instruction writes four 128-byte row segments with full sectors.
  alignas(16) float stage[kGemmEpiWarps][32][kStageLd];
};

// Streaming mode (layer wavefront): the GEMM runs CONCURRENTLY with the recurrent kernel that produces its A
// operand and with the recurrent kernel that consumes its output.
//   * a CTA owns one (64-word group, column tile) and every `n_par`-th pair of output steps, in time order;
//   * before the first TMA of a tile the producer warp polls the arrival counter of the LAST source step the tile reads
//     (`src_flags[grp * src_steps + step]`, incremented with release semantics by every epilogue warp of the producing kernel
//     once its image stores of that step are out; steps complete in order) until it reaches `src_target[grp]`;
//   * the epilogue also emits the tile as bf16 operand blocks of the consumer's fused input projection
//     (`x_out`: [steps][ceil(B/16)][16 rows][128 B], SWIZZLE_128B, N <= 64 columns) and then increments `dst_flags[grp * n_pairs +
//     pair]` (release; one arrival per epilogue warp).
struct GemmStream {
  const unsigned int* src_flags;   // [n_groups][src_steps]
  const unsigned int* src_target;  // [n_groups] arrivals that complete a source step
  unsigned int* dst_flags;         // [n_groups][n_pairs], or nullptr
  uint8_t* x_out;                  // consumer operand blocks, or nullptr
  int* status;                     // sticky status word (watchdog), or nullptr
  int src_per_step;                // source steps per output step (2: pooled post_linear over frame pairs)
  int n_par;                       // CTAs that alternate over the step pairs of one (group, column tile)
  int n_ct;                        // column tiles per CTA (1: one CTA per column tile; 2: half as many CTAs, each walks two tiles per pair)
  int reverse;                     // the producer runs backwards in time (BPTT): pairs from the last to the first
};

template <bool STREAM>
__global__ void __launch_bounds__(kGemmThreads, 1)
tc_gemm_img_kernel(const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img, const float* __restrict__ bias,
                   float* __restrict__ C, int steps, int B, int N, int KB, int BN, int accumulate, GemmStream st) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t stage_bytes = 16384u + (uint32_t)BN * 128u;   // A [128 x 128 B] + B [BN x 128 B]; BN % 8 == 0 keeps 1 KB alignment
  GemmBars& bars = *reinterpret_cast<GemmBars*>(base + (size_t)kGemmStages * stage_bytes);
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  // a timed-out TMA / MMA wait is recorded in the caller's sticky status word when it gave one (planner.check() reads it)
  __shared__ int local_err;
  volatile int* err = (st.status != nullptr) ? reinterpret_cast<volatile int*>(st.status) : reinterpret_cast<volatile int*>(&local_err);
  if (threadIdx.x == 0) local_err = 0;

  const int n_groups = (B + kRows - 1) / kRows;
  const int n_pairs = (steps + 1) / 2;
  const int n_nt = pad_n(N) / BN;
  // tile walk: batch mode = all tiles strided over the grid; streaming = this CTA's (group, column tile), every n_par-th pair
  const int total = STREAM ? (n_pairs - (int)(blockIdx.x % st.n_par) + st.n_par - 1) / st.n_par * st.n_ct : n_groups * n_pairs * n_nt;
  const int tile_first = STREAM ? 0 : (int)blockIdx.x, tile_stride = STREAM ? 1 : (int)gridDim.x;
  auto coords = [&](int tile, int& nt, int& sp, int& grp) {
    if (STREAM) {
      const int cta = (int)blockIdx.x / st.n_par, n_col = n_nt / st.n_ct;   // CTAs per group that share the columns
      nt = cta % n_col + (tile % st.n_ct) * n_col; grp = cta / n_col; sp = (int)(blockIdx.x % st.n_par) + (tile / st.n_ct) * st.n_par;
      if (st.reverse) sp = n_pairs - 1 - sp;
    } else {
      nt = tile % n_nt; const int rest = tile / n_nt; sp = rest % n_pairs; grp = rest / n_pairs;
    }
  };

  if (tid == 0) {
    for (int i = 0; i < kGemmStages; ++i) { mbar_init(&bars.full[i], 1); mbar_init(&bars.empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars.tmem_full[i], 1); mbar_init(&bars.tmem_empty[i], kGemmEpiWarps); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(&bars.tmem_base);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = bars.tmem_base;

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform control flow, one elected lane issues) =====================
    int s = 0;
    uint32_t ph = 0;
    for (int tile = tile_first; tile < total; tile += tile_stride) {
      int nt, sp, grp;
      coords(tile, nt, sp, grp);
      const int t0 = 2 * sp;
      const bool two = (t0 + 1 < steps);
      const uint8_t* a0 = a_img + ((size_t)(grp * steps + t0) * KB) * (kRows * 128);
      const uint8_t* bt = b_img + (size_t)nt * KB * BN * 128;
      if (STREAM) {
        // the producing recurrent kernel has finished the last source step this tile reads (steps complete in order)
        const int src_steps = steps * st.src_per_step;
        // forward producers finish the tile's LAST source step last, reverse-time producers its FIRST one
        const int last = st.reverse ? t0 * st.src_per_step : min((t0 + 2) * st.src_per_step, src_steps) - 1;
        const unsigned int* f = st.src_flags + (size_t)grp * src_steps + last;
        const unsigned int want = st.src_target[grp];
        uint64_t w0 = 0;
        for (unsigned int spin = 0; ld_acquire_u32(f) < want; ++spin) {
          __nanosleep(200);
          if ((spin & 255u) == 255u) {
            if (w0 == 0) w0 = globaltimer_ns();
            if (*err != 0) break;
            if (globaltimer_ns() - w0 > kWatchdogNs) { *err = 1; break; }
          }
        }
        fence_proxy_async_global();   // the acquire above orders the generic proxy; the TMA reads go through the async proxy
      }
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&bars.empty[s], ph ^ 1u, err);
        if (elect_one_sync()) {
          uint8_t* sa = base + (size_t)s * stage_bytes;
          mbar_arrive_expect_tx(&bars.full[s], (two ? 16384u : 8192u) + (uint32_t)BN * 128u);
          bulk_g2s(sa, a0 + (size_t)kb * (kRows * 128), kRows * 128, &bars.full[s]);
          if (two) bulk_g2s(sa + 8192, a0 + ((size_t)KB + kb) * (kRows * 128), kRows * 128, &bars.full[s]);
          bulk_g2s(sa + 16384, bt + (size_t)kb * BN * 128, (uint32_t)BN * 128u, &bars.full[s]);
        }
        __syncwarp();
        if (++s == kGemmStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Warp-uniform control flow and elect.sync: descriptors stay in uniform registers, so the four UTCHMMA of a k-block
    // issue back to back (under `if (lane == 0)` each one sits in an ELECT / R2UR / BRA.U.ANY loop, ~90 ns apiece).
    const uint32_t idesc = make_idesc_bf16(128, BN);
    int s = 0;
    uint32_t ph = 0;
    int local = 0;
    for (int tile = tile_first; tile < total; tile += tile_stride, ++local) {
      const int ab = local & 1;
      mbar_wait(&bars.tmem_empty[ab], (uint32_t)(((local >> 1) & 1) ^ 1), err);   // epilogue drained this buffer
      tcgen05_fence_after();
      const uint32_t d = tmem + (uint32_t)(ab * 256);
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&bars.full[s], ph, err);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(base + (size_t)s * stage_bytes);
        const uint64_t da = make_smem_desc_sw128(sa);
        const uint64_t db = make_smem_desc_sw128(sa + 16384u);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
          umma_commit(&bars.empty[s]);          // stage reusable once these MMAs have read it
          if (kb == KB - 1) umma_commit(&bars.tmem_full[ab]);   // accumulator complete
        }
        __syncwarp();
        if (++s == kGemmStages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> HBM =====================
    const int lg = warp & 3;                     // TMEM lane group this warp may access
    const int ew = warp - 2;                     // epilogue warp 0..7: stage buffer; ew >> 2 = which alternate column chunks
    const int r = lg * 32 + lane;                // tile row: two time steps x 64 words
    int local = 0;
    for (int tile = tile_first; tile < total; tile += tile_stride, ++local) {
      int nt, sp, grp;
      coords(tile, nt, sp, grp);
      const int ab = local & 1;
      const int t = 2 * sp + (r >> 6), b = grp * kRows + (r & 63);
      const bool valid = (t < steps) && (b < B);
      mbar_wait(&bars.tmem_full[ab], (uint32_t)((local >> 1) & 1), err);
      tcgen05_fence_after();
      float* crow = C + ((size_t)t * B + (valid ? b : 0)) * N;
      const uint32_t taddr = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(ab * 256);
      for (int c0 = 32 * (ew >> 2); c0 < BN; c0 += 32 * (kGemmEpiWarps / 4)) {
        float v[32];
        if (BN - c0 >= 32) {
          tmem_ld_x32(taddr + (uint32_t)c0, v);
        } else {   // BN is a multiple of 16: one 16-column tail
          float v16[16];
          tmem_ld_x16(taddr + (uint32_t)c0, v16);
#pragma unroll
          for (int i = 0; i < 16; ++i) { v[i] = v16[i]; v[16 + i] = 0.f; }
        }
        const int n0 = nt * BN + c0;
        const int lim = (BN - c0 >= 32) ? 32 : 16;
        if (STREAM && st.x_out != nullptr && valid) {
          // the consumer's operand block of (step t, word quarter b / 16): row b % 16, 16-byte chunk j ^ (row % 8) holds columns
          // 8 j .. 8 j + 7 as bf16 (+ bias; columns >= N are zero: the packed weights' pad rows and no bias)
          const int row = b & (kWq - 1);
          uint8_t* blk = st.x_out + ((size_t)t * ((B + kWq - 1) / kWq) + (size_t)(b / kWq)) * kXBlockBytes + (size_t)row * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t w[4];
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
              const int i0 = 8 * j + 2 * e2, na = n0 + i0, nb = na + 1;
              const float fa = (i0 < lim && na < N) ? v[i0] + (bias ? __ldg(bias + na) : 0.f) : 0.f;
              const float fb = (i0 + 1 < lim && nb < N) ? v[i0 + 1] + (bias ? __ldg(bias + nb) : 0.f) : 0.f;
              const __nv_bfloat162 pr = __floats2bfloat162_rn(fa, fb);
              w[e2] = *reinterpret_cast<const uint32_t*>(&pr);
            }
            const int chunk = (n0 >> 3) + j;
            if (chunk < 8) *reinterpret_cast<uint4*>(blk + ((chunk ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        if ((N & 3) == 0) {
          // ---- coalesced path: own row -> staging, then 8 instructions x (4 rows x 128 B)
          float* srow = &bars.stage[ew][lane][0];
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(srow + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          __syncwarp();
          const int cc = (lane & 7) * 4, n = n0 + cc;
          const bool col_ok = cc < lim && n < N;
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias && col_ok) bv = __ldg(reinterpret_cast<const float4*>(bias + n));
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = 4 * i + (lane >> 3), r2 = lg * 32 + rr;
            const int t2 = 2 * sp + (r2 >> 6), b2 = grp * kRows + (r2 & 63);
            if (col_ok && t2 < steps && b2 < B) {
              float4 o = *reinterpret_cast<const float4*>(&bars.stage[ew][rr][cc]);
              o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
              float4* dst = reinterpret_cast<float4*>(C + ((size_t)t2 * B + b2) * N + n);
              if (accumulate) { const float4 p = *dst; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
              *dst = o;
            }
          }
          __syncwarp();   // the staging tile is rewritten by the next chunk
        } else if (valid) {
          // (static indexing: a run-time index into v[] would put the whole array into local memory for every variant)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int n = n0 + i;
            if (i < lim && n < N) {
              float o = v[i] + (bias ? __ldg(bias + n) : 0.f);
              if (accumulate) o += crow[n];
              crow[n] = o;
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars.tmem_empty[ab]);
        // streaming: this warp's rows of the pair (fp32 C and operand blocks) are out -- release them to the consumer
        if (STREAM && st.dst_flags != nullptr) red_release_add_u32(st.dst_flags + (size_t)grp * n_pairs + sp, 1u);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ---- CTA-pair variant of the batch mode (wide outputs: N = 2880 / 720) ------------------------------------------------------
// ncu on the K = 720, N = 2880 projection (profiles/r2b_ncu_gategemm.txt): tensor pipe 36 % active, L2 and crossbar at 26-41 % of
// their peaks, producer and MMA warps waiting on each other -- every SM takes in ~41 B per cycle, about what an SM takes in
// under cuBLAS at its peak, but a 128 x 160 tile only does 71 FLOP per byte received.  With tcgen05.mma.cta_group::2 the two SMs
// of a cluster run ONE M = 256 MMA: each CTA stages its own 128 rows of A and only HALF of the B tile (80 rows), i.e. 26 instead
// of 36 KB per k-block for the same arithmetic (98 FLOP per byte), and the ring holds 6 stages instead of 4.
//   rank 0 (leader): producer, MMA issuer (M = 256), epilogue of the first row tile of the pair
//   rank 1 (peer):   producer, forwarder (its "stage landed" -> the leader's peer_full barrier), epilogue of the second row tile
// tcgen05.commit.cta_group::2 with a multicast mask frees the stage / publishes the accumulator in both CTAs.
constexpr int kGemm2MaxStages = 8;
struct Gemm2Bars {
  uint64_t full[kGemm2MaxStages];     // this CTA's loads of the stage landed (tx-count)
  uint64_t peer_full[kGemm2MaxStages];   // leader only: the peer's loads of the stage landed (remote arrive by the peer's forwarder)
  uint64_t empty[kGemm2MaxStages];    // the pair's MMAs have read the stage (commit, multicast to both CTAs)
  uint64_t tmem_full[2];              // accumulator complete (commit, multicast)
  uint64_t tmem_empty[2];             // leader only: both CTAs' epilogue warps have drained the buffer
  uint32_t tmem_base;
  alignas(16) float stage[kGemmEpiWarps][32][kStageLd];
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
tc_gemm_img2_kernel(const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img, const float* __restrict__ bias,
                    float* __restrict__ C, int steps, int B, int N, int KB, int BN, int accumulate, int* status,
                    int kGemm2Stages) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t b_half = (uint32_t)BN * 64u;                  // this CTA's half of the B tile: BN / 2 rows x 128 B
  const uint32_t stage_bytes = 16384u + b_half;                // BN % 16 == 0 keeps 1 KB alignment
  Gemm2Bars& bars = *reinterpret_cast<Gemm2Bars*>(base + (size_t)kGemm2Stages * stage_bytes);
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const uint32_t rank = cluster_ctarank_u32();
  const bool leader = rank == 0;
  __shared__ int local_err;
  volatile int* err = (status != nullptr) ? reinterpret_cast<volatile int*>(status) : reinterpret_cast<volatile int*>(&local_err);
  if (threadIdx.x == 0) local_err = 0;

  const int n_groups = (B + kRows - 1) / kRows;
  const int n_pairs = (steps + 1) / 2;
  const int n_nt = pad_n(N) / BN;
  const int n_rt = n_groups * n_pairs;                         // row tiles (two time steps of one 64-word group)
  const int total = ((n_rt + 1) / 2) * n_nt;                   // pair tiles: column tile fastest, two consecutive row tiles
  const int tile_first = (int)(blockIdx.x >> 1), tile_stride = (int)(gridDim.x >> 1);
  auto coords = [&](int tile, int& nt, int& sp, int& grp, bool& valid_rt) {
    nt = tile % n_nt;
    int rt = 2 * (tile / n_nt) + (int)rank;
    valid_rt = rt < n_rt;
    if (!valid_rt) rt = n_rt - 1;                              // odd count: the last pair's second tile is computed and dropped
    sp = rt % n_pairs; grp = rt / n_pairs;
  };

  if (tid == 0) {
    for (int i = 0; i < kGemm2Stages; ++i) { mbar_init(&bars.full[i], 1); mbar_init(&bars.peer_full[i], 1); mbar_init(&bars.empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars.tmem_full[i], 1); mbar_init(&bars.tmem_empty[i], 2 * kGemmEpiWarps); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta<512>(&bars.tmem_base);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs' barriers are initialised before any remote arrive / multicast commit targets them
  tcgen05_fence_after();
  const uint32_t tmem = bars.tmem_base;

  if (warp == 0) {
    // ===================== TMA producer: own A rows, own half of the B tile =====================
    int s = 0;
    uint32_t ph = 0;
    for (int tile = tile_first; tile < total; tile += tile_stride) {
      int nt, sp, grp; bool vr;
      coords(tile, nt, sp, grp, vr);
      const int t0 = 2 * sp;
      const bool two = (t0 + 1 < steps);
      const uint8_t* a0 = a_img + ((size_t)(grp * steps + t0) * KB) * (kRows * 128);
      const uint8_t* bt = b_img + (size_t)nt * KB * BN * 128 + (size_t)rank * b_half;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&bars.empty[s], ph ^ 1u, err);
        if (elect_one_sync()) {
          uint8_t* sa = base + (size_t)s * stage_bytes;
          mbar_arrive_expect_tx(&bars.full[s], (two ? 16384u : 8192u) + b_half);
          bulk_g2s(sa, a0 + (size_t)kb * (kRows * 128), kRows * 128, &bars.full[s]);
          if (two) bulk_g2s(sa + 8192, a0 + ((size_t)KB + kb) * (kRows * 128), kRows * 128, &bars.full[s]);
          bulk_g2s(sa + 16384, bt + (size_t)kb * BN * 128, b_half, &bars.full[s]);
        }
        __syncwarp();
        if (++s == kGemm2Stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (M = 256 over the pair) =====================
    const uint32_t idesc = make_idesc_bf16(256, BN);
    int s = 0;
    uint32_t ph = 0;
    int local = 0;
    for (int tile = tile_first; tile < total; tile += tile_stride, ++local) {
      const int ab = local & 1;
      mbar_wait(&bars.tmem_empty[ab], (uint32_t)(((local >> 1) & 1) ^ 1), err);   // both epilogues drained this buffer
      tcgen05_fence_after();
      const uint32_t d = tmem + (uint32_t)(ab * 256);
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&bars.full[s], ph, err);
        mbar_wait(&bars.peer_full[s], ph, err);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(base + (size_t)s * stage_bytes);
        const uint64_t da = make_smem_desc_sw128(sa);
        const uint64_t db = make_smem_desc_sw128(sa + 16384u);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_2cta(d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
          umma_commit_2cta(&bars.empty[s], 3);
          if (kb == KB - 1) umma_commit_2cta(&bars.tmem_full[ab], 3);
        }
        __syncwarp();
        if (++s == kGemm2Stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== peer: forward "my part of the stage landed" to the leader =====================
    int s = 0;
    uint32_t ph = 0;
    for (int tile = tile_first; tile < total; tile += tile_stride) {
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&bars.full[s], ph, err);
        if (elect_one_sync()) mbar_arrive_remote(&bars.peer_full[s], 0u);
        __syncwarp();
        if (++s == kGemm2Stages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue of this CTA's 128 rows: TMEM -> staging -> HBM =====================
    const int lg = warp & 3;
    const int ew = warp - 2;
    int local = 0;
    for (int tile = tile_first; tile < total; tile += tile_stride, ++local) {
      int nt, sp, grp; bool vr;
      coords(tile, nt, sp, grp, vr);
      const int ab = local & 1;
      mbar_wait(&bars.tmem_full[ab], (uint32_t)((local >> 1) & 1), err);
      tcgen05_fence_after();
      const uint32_t taddr = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(ab * 256);
      for (int c0 = 32 * (ew >> 2); c0 < BN && vr; c0 += 32 * (kGemmEpiWarps / 4)) {
        float v[32];
        if (BN - c0 >= 32) {
          tmem_ld_x32(taddr + (uint32_t)c0, v);
        } else {
          float v16[16];
          tmem_ld_x16(taddr + (uint32_t)c0, v16);
#pragma unroll
          for (int i = 0; i < 16; ++i) { v[i] = v16[i]; v[16 + i] = 0.f; }
        }
        const int n0 = nt * BN + c0;
        const int lim = (BN - c0 >= 32) ? 32 : 16;
        float* srow = &bars.stage[ew][lane][0];
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(srow + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        __syncwarp();
        const int cc = (lane & 7) * 4, n = n0 + cc;
        const bool col_ok = cc < lim && n < N;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias && col_ok) bv = __ldg(reinterpret_cast<const float4*>(bias + n));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = 4 * i + (lane >> 3), r2 = lg * 32 + rr;
          const int t2 = 2 * sp + (r2 >> 6), b2 = grp * kRows + (r2 & 63);
          if (col_ok && t2 < steps && b2 < B) {
            float4 o = *reinterpret_cast<const float4*>(&bars.stage[ew][rr][cc]);
            o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
            float4* dst = reinterpret_cast<float4*>(C + ((size_t)t2 * B + b2) * N + n);
            if (accumulate) { const float4 p = *dst; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
            *dst = o;
          }
        }
        __syncwarp();
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&bars.tmem_empty[ab]);
        else mbar_arrive_remote(&bars.tmem_empty[ab], 0u);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();          // no CTA of the pair leaves (or frees tensor memory) while the other may still signal it
  if (warp == 1) tmem_dealloc_2cta<512>(tmem);
}

}  // namespace tc
}  // namespace paule

using namespace paule;
using namespace paule::tc;

extern "C" size_t paule_tc_gemm_packed_bytes(int64_t N, int64_t nseg) {
  if (N <= 0 || nseg <= 0) return 0;
  return (size_t)pad_n((int)N) * (size_t)nseg * kSegPad * 2;
}

extern "C" int paule_tc_gemm_pack(const float* W, void* packed, int64_t N, int64_t nseg, paule_stream_t stream) {
  PAULE_REQUIRE(W && packed && N > 0 && nseg > 0 && nseg <= 4);
  PAULE_REQUIRE(reinterpret_cast<uintptr_t>(packed) % 16 == 0);
  const int np = pad_n((int)N), bn = pick_bn(np);
  pack_gemm_b_kernel<<<np / bn, 256, 0, as_stream(stream)>>>(W, reinterpret_cast<uint8_t*>(packed), (int)N, (int)nseg, bn,
                                                               kSegLen, kSegPad);
  PAULE_LAUNCH_CHECK("pack_gemm_b_kernel");
  return PAULE_OK;
}

// narrow-K GEMM (K <= 64, one k-block): weights W [N, K] and the fp32 -> image conversion of its A operand
extern "C" size_t paule_tc_gemm_packed_bytes_k64(int64_t N) { return N <= 0 ? 0 : (size_t)pad_n((int)N) * 64 * 2; }
extern "C" int paule_tc_gemm_pack_k64(const float* W, void* packed, int64_t N, int64_t K, paule_stream_t stream) {
  PAULE_REQUIRE(W && packed && N > 0 && K > 0 && K <= 64);
  PAULE_REQUIRE(reinterpret_cast<uintptr_t>(packed) % 16 == 0);
  const int np = pad_n((int)N), bn = pick_bn(np);
  pack_gemm_b_kernel<<<np / bn, 256, 0, as_stream(stream)>>>(W, reinterpret_cast<uint8_t*>(packed), (int)N, 1, bn, (int)K, 64);
  PAULE_LAUNCH_CHECK("pack_gemm_b_kernel");
  return PAULE_OK;
}
extern "C" size_t paule_tc_a_image_bytes(int64_t steps, int64_t B) {
  if (steps <= 0 || B <= 0) return 0;
  return (size_t)((B + kRows - 1) / kRows) * (size_t)steps * (kRows * 128);
}
// x [steps, B, I <= 64] fp32 -> img (>= paule_tc_a_image_bytes, ZERO-FILLED once by the owner: rows of words >= B)
extern "C" int paule_tc_a_image(const float* x, void* img, int64_t steps, int64_t B, int64_t I, paule_stream_t stream) {
  PAULE_REQUIRE(x && img && steps >= 0 && B > 0 && I > 0 && I <= 64 && (I % 2) == 0);
  PAULE_REQUIRE(reinterpret_cast<uintptr_t>(img) % 16 == 0);
  const int64_t total = steps * B * 8;
  if (total == 0) return PAULE_OK;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  a_image_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, reinterpret_cast<uint8_t*>(img), steps, B, (int)I);
  PAULE_LAUNCH_CHECK("a_image_kernel");
  return PAULE_OK;
}
// C [steps, B, N] (+)= A W^T + bias over the images of paule_tc_a_image and the weights of paule_tc_gemm_pack_k64
extern "C" int paule_tc_gemm_img_k64(const void* a_img, const void* packed_b, const float* bias, float* C, int64_t steps,
                                     int64_t B, int64_t N, int accumulate, paule_stream_t stream) {
  return paule::tc::gemm_img_kb(a_img, packed_b, bias, C, steps, B, N, 1, accumulate, nullptr, as_stream(stream));
}

namespace paule {
namespace tc {

static int gemm_attrs() {
  static unsigned long long attr_set = 0ull;
  if (once_per_device(attr_set)) {   // the widest tile (BN = 240) bounds every launch
    const int smem_max = kGemmStages * (16384 + 240 * 128) + (int)sizeof(GemmBars) + 1024 + 16;
    PAULE_CUDA(cudaFuncSetAttribute(tc_gemm_img_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    PAULE_CUDA(cudaFuncSetAttribute(tc_gemm_img_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
  }
  return PAULE_OK;
}

int gemm_stream_ctas(int64_t B, int64_t N, int n_par, int n_ct) {
  const int np = pad_n((int)N), bn = pick_bn(np);
  if (n_ct < 1 || (np / bn) % n_ct != 0) n_ct = 1;
  return (int)((B + kRows - 1) / kRows) * (np / bn / n_ct) * n_par;
}
// arrivals that complete one pair of output steps in dst_flags: one per epilogue warp and column tile
unsigned int gemm_stream_arrivals(int64_t N) {
  const int np = pad_n((int)N), bn = pick_bn(np);
  return (unsigned int)kGemmEpiWarps * (unsigned int)(np / bn);
}

// the GEMM of paule_tc_gemm_img in streaming mode (see GemmStream); `status` may be NULL
int gemm_img_stream(const void* a_img, const void* packed_b, const float* bias, float* C, int64_t steps, int64_t B, int64_t N,
                    int64_t nseg, const unsigned int* src_flags, const unsigned int* src_target, int src_per_step,
                    unsigned int* dst_flags, void* x_out, int n_par, int* status, cudaStream_t s, int reverse, int accumulate,
                    int n_ct) {
  PAULE_REQUIRE(a_img && packed_b && C && steps > 0 && B > 0 && N > 0 && nseg > 0 && nseg <= 4 && src_flags && src_target);
  PAULE_REQUIRE(n_par >= 1 && (x_out == nullptr || N <= kXK));
  PAULE_TRY(gemm_attrs());
  const int np = pad_n((int)N), bn = pick_bn(np), KB = (int)nseg * kNumKB;
  const int smem_own = kGemmStages * (16384 + bn * 128) + (int)sizeof(GemmBars) + 1024 + 16;
  const int smem = smem_own > kExclusiveSmemBytes ? smem_own : kExclusiveSmemBytes;   // never shares an SM (tensor memory)
  if (n_ct < 1 || (np / bn) % n_ct != 0) n_ct = 1;
  GemmStream st{src_flags, src_target, dst_flags, reinterpret_cast<uint8_t*>(x_out), status, src_per_step, n_par, n_ct, reverse};
  tc_gemm_img_kernel<true><<<gemm_stream_ctas(B, N, n_par, n_ct), kGemmThreads, (size_t)smem, s>>>(
      reinterpret_cast<const uint8_t*>(a_img), reinterpret_cast<const uint8_t*>(packed_b), bias, C, (int)steps, (int)B, (int)N,
      KB, bn, accumulate, st);
  PAULE_LAUNCH_CHECK("tc_gemm_img_kernel<stream>");
  return PAULE_OK;
}

}  // namespace tc
}  // namespace paule

extern "C" int paule_tc_gemm_img(const void* a_img, const void* packed_b, const float* bias, float* C, int64_t steps,
                                 int64_t B, int64_t N, int64_t nseg, int accumulate, paule_stream_t stream) {
  return paule::tc::gemm_img(a_img, packed_b, bias, C, steps, B, N, nseg, accumulate, nullptr, as_stream(stream));
}

// paule_tc_gemm_img with the caller's sticky status word: a timed-out TMA / MMA wait is recorded there (planner.check())
int paule::tc::gemm_img(const void* a_img, const void* packed_b, const float* bias, float* C, int64_t steps, int64_t B,
                        int64_t N, int64_t nseg, int accumulate, int* status, cudaStream_t stream) {
  PAULE_REQUIRE(nseg > 0 && nseg <= 4);
  return gemm_img_kb(a_img, packed_b, bias, C, steps, B, N, (int)nseg * kNumKB, accumulate, status, stream);
}

// the batch GEMM with an explicit number of 64-wide k-blocks per row of A
int paule::tc::gemm_img_kb(const void* a_img, const void* packed_b, const float* bias, float* C, int64_t steps, int64_t B,
                           int64_t N, int KB, int accumulate, int* status, cudaStream_t stream) {
  PAULE_REQUIRE(a_img && packed_b && C && steps >= 0 && B > 0 && N > 0 && KB > 0 && KB <= 4 * kNumKB);
  PAULE_REQUIRE(reinterpret_cast<uintptr_t>(a_img) % 16 == 0 && reinterpret_cast<uintptr_t>(packed_b) % 16 == 0);
  PAULE_REQUIRE(reinterpret_cast<uintptr_t>(C) % 16 == 0);
  if (steps == 0) return PAULE_OK;
  PAULE_TRY(gemm_attrs());
  const int np = pad_n((int)N), bn = pick_bn(np);
  const int n_groups = (int)((B + kRows - 1) / kRows), n_pairs = (int)((steps + 1) / 2);
  const int64_t total = (int64_t)n_groups * n_pairs * (np / bn);
  // wide outputs: the CTA-pair kernel (PAULE_GEMM_2CTA=0: the single-CTA kernel, A/B timing)
  static const bool pair_on = getenv("PAULE_GEMM_2CTA") == nullptr || atoi(getenv("PAULE_GEMM_2CTA")) != 0;
  if (pair_on && bn >= 128 && (N % 4) == 0 && n_groups * n_pairs >= 2) {
    static unsigned long long attr2 = 0ull;
    const int smem_cap = 227 * 1024 - 256;   // opt-in maximum less the kernel's static shared memory
    if (once_per_device(attr2))
      PAULE_CUDA(cudaFuncSetAttribute(tc_gemm_img2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_cap));
    // as many stages as fit next to the barriers and the epilogue staging tiles (BN = 160: 7, BN = 144: 7, BN = 240: 6)
    static const int forced_stages = getenv("PAULE_GEMM2_STAGES") ? atoi(getenv("PAULE_GEMM2_STAGES")) : 0;
    const int fixed = (int)sizeof(Gemm2Bars) + 1024 + 16, per_stage = 16384 + bn * 64;
    int stages = (smem_cap - fixed) / per_stage;
    if (stages > kGemm2MaxStages) stages = kGemm2MaxStages;
    if (forced_stages >= 2 && forced_stages <= stages) stages = forced_stages;
    const int smem2 = stages * per_stage + fixed;
    const int64_t pairs = (int64_t)((n_groups * n_pairs + 1) / 2) * (np / bn);
    const int max_clusters = sm_count() / 2;
    const int clusters = (int)(pairs < max_clusters ? pairs : max_clusters);
    tc_gemm_img2_kernel<<<2 * clusters, kGemmThreads, (size_t)smem2, stream>>>(
        reinterpret_cast<const uint8_t*>(a_img), reinterpret_cast<const uint8_t*>(packed_b), bias, C, (int)steps, (int)B,
        (int)N, KB, bn, accumulate, status, stages);
    PAULE_LAUNCH_CHECK("tc_gemm_img2_kernel");
    return PAULE_OK;
  }
  const int smem = kGemmStages * (16384 + bn * 128) + (int)sizeof(GemmBars) + 1024 + 16;
  const int grid = (int)((total < (int64_t)sm_count()) ? total : (int64_t)sm_count());
  GemmStream st{};
  st.status = status;
  tc_gemm_img_kernel<false><<<grid, kGemmThreads, (size_t)smem, stream>>>(
      reinterpret_cast<const uint8_t*>(a_img), reinterpret_cast<const uint8_t*>(packed_b), bias, C, (int)steps, (int)B,
      (int)N, KB, bn, accumulate, st);
  PAULE_LAUNCH_CHECK("tc_gemm_img_kernel");
  return PAULE_OK;
}
