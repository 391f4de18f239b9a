// sm_100a building blocks for the tensor-core path: tcgen05 (UMMA) descriptors, TMEM allocation and loads,
// mbarriers, 1-D TMA bulk copies, proxy fences and a watchdog-guarded grid barrier.
//
// Operand convention used by every tcgen05 kernel in this directory ("UMMA image"):
//   a bf16 matrix [rows, K] (K-major) is stored as K/64 "k-blocks"; one k-block is rows x 64 bf16 = rows x 128 B,
//   row r at byte r*128, and inside each row the eight 16-byte chunks are XOR-swizzled with (r % 8)
//   (the canonical SWIZZLE_128B K-major layout, Swizzle<3,4,3>).  rows must be a multiple of 8 and a k-block must
//   start on a 1024-byte boundary in shared memory.  Producers write this image directly into HBM, so that a
//   consumer moves a whole k-block with ONE cp.async.bulk (no tensor map, no re-layout) and hands it to tcgen05.mma.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace paule {
namespace tc {

constexpr int kKB = 64;              // bf16 elements per k-block row (128 bytes)
constexpr uint32_t kRowBytes = 128;  // bytes per row of a k-block

// byte offset of element (r, j) inside a UMMA image with `rows` rows per k-block
__host__ __device__ __forceinline__ size_t umma_offset(int rows, int r, int j) {
  const int kb = j >> 6, c = (j & 63) >> 3, e = j & 7;
  return (size_t)kb * rows * kRowBytes + (size_t)r * kRowBytes + (size_t)((c ^ (r & 7)) << 4) + (size_t)e * 2;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- shared-memory matrix descriptor (K-major, SWIZZLE_128B): cute::UMMA::SmemDescriptor bit layout
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address, 16-byte units
  d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major; CuTe sets 1)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: 8 rows x 128 B between row groups
  d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                            // layout type SWIZZLE_128B
  return d;
}

// ---- instruction descriptor for kind::f16, bf16 x bf16 -> fp32, both operands K-major (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- CTA pair (cta_group::2): one tcgen05.mma spans the two SMs of a cluster of 2 -- M = 256 (128 rows in each CTA's tensor
// memory), A from each CTA's own shared memory, B split over the two CTAs (N/2 rows each).  Issued by the leader CTA only.
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this shared-memory offset in every CTA of `cta_mask` when all previously issued pair MMAs completed
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_dst) {  // one full warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t tmem_addr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "r"(kCols) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster.  Default semantics (as CUTLASS'
// ClusterBarrier::arrive(cta_id)): the explicit .release.cluster form compiles to MEMBAR.ALL.GPU + ERRBAR in front of the arrive
// (~1 us per call, measured: it serialised the CTA-pair GEMM's per-stage hand-over).  What the arrive publishes here was
// written by the async proxy (TMA, tcgen05.ld) and completed on an mbarrier / tcgen05 fence before this thread arrives.
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// wait with cluster-scope acquire: the phase may have been completed by the other CTA of the cluster
__device__ __forceinline__ bool mbar_try_wait_cl(uint64_t* bar, uint32_t parity) {
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

// A operand in tensor memory ("TS"): [M = 128 lanes, K] bf16, 32-bit column c of lane m holds k = 2c (low half) and 2c+1;
// one instruction consumes K = 16 = 8 columns.  Measured (tools/mma_bench.cu): 12 warps x 4 MMAs (M=128, N=16) into one
// accumulator complete in 653 cycles with A in TMEM against 2082 with A streamed from shared memory.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// One elected lane of a converged warp.  Inside `if (elect_one_sync())` the compiler knows the region is warp-uniform, so
// descriptors computed from uniform values stay in uniform registers and tcgen05.mma / tcgen05.commit issue directly;
// under a plain `if (lane == 0)` every UTCHMMA is wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (~90 ns each).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// warp index as a value the compiler treats as warp-uniform
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// ---- TMEM
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem_addr) {  // the same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "r"(kCols) : "memory");
}

// 32 lanes x 32 bit, 16 consecutive columns: thread i of the warp receives lane (base_lane + i)
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 32 lanes x 32 bit, 8 consecutive columns
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint4& a, const uint4& b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(a.x),
               "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ void tmem_zero_x8(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(z)
               : "memory");
}

// zero 16 columns of this warp's 32 TMEM lanes (used to re-arm accumulator tiles that several MMA issuers add into)
__device__ __forceinline__ void tmem_zero_x16(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
      "r"(z)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

constexpr uint64_t kWatchdogNs = 4000000000ull;  // 4 s: a stuck barrier sets the error flag instead of hanging the GPU

// returns false on watchdog expiry
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, volatile int* err_flag) {
  uint64_t t0 = 0;
  for (unsigned int spin = 0;; ++spin) {
    if (mbar_try_wait(bar, parity)) return true;   // try_wait suspends in hardware: this is not a busy poll
    if ((spin & 63u) == 63u) {                     // watchdog bookkeeping stays off the wake-up path
      if (t0 == 0) t0 = globaltimer_ns();
      if (*err_flag != 0) return false;
      if (globaltimer_ns() - t0 > kWatchdogNs) {
        *err_flag = 2;
        return false;
      }
    }
  }
}

__device__ __forceinline__ bool mbar_wait_cl(uint64_t* bar, uint32_t parity, volatile int* err_flag) {
  uint64_t t0 = 0;
  for (unsigned int spin = 0;; ++spin) {
    if (mbar_try_wait_cl(bar, parity)) return true;
    if ((spin & 63u) == 63u) {
      if (t0 == 0) t0 = globaltimer_ns();
      if (*err_flag != 0) return false;
      if (globaltimer_ns() - t0 > kWatchdogNs) {
        *err_flag = 2;
        return false;
      }
    }
  }
}

// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank_u32() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Orders this thread's generic-proxy view of GLOBAL memory (the acquire that observed the producers' release) before
// its subsequent async-proxy (TMA) reads of global memory.  PAULE_PROXY_FENCE: 0 none, 1 .global (default), 2 full.
#ifndef PAULE_PROXY_FENCE
#define PAULE_PROXY_FENCE 1
#endif
__device__ __forceinline__ void fence_proxy_async_global() {
#if PAULE_PROXY_FENCE == 1
  asm volatile("fence.proxy.async.global;" ::: "memory");
#elif PAULE_PROXY_FENCE == 2
  asm volatile("fence.proxy.async;" ::: "memory");
#endif
}

// ---- grid barrier: monotonic arrival counter in global memory (zeroed by the host before the launch)
__device__ __forceinline__ void grid_arrive(unsigned int* counter) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}
__device__ __forceinline__ bool grid_wait(const unsigned int* counter, unsigned int target, volatile int* err_flag) {
  unsigned int v;
  uint64_t t0 = 0;
  for (unsigned int spin = 0;; ++spin) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    if (v >= target) return true;
    if ((spin & 255u) == 255u) {   // watchdog bookkeeping stays off the polling path
      if (t0 == 0) t0 = globaltimer_ns();
      if (*err_flag != 0) return false;
      if (globaltimer_ns() - t0 > kWatchdogNs) {
        *err_flag = 1;
        return false;
      }
    }
  }
}

// ---- flag-based hand-over between concurrently running kernels (layer wavefront): release-increment / acquire-poll
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- self-validating exchange of bf16 operands between CTAs.  Every exchanged value satisfies |x| < 2 (h is a product of
// a sigmoid and a tanh; gradients are clamped), so bit 14 of its bf16 encoding -- the top exponent bit -- is always 0
// and is free to carry a phase bit.  Consecutive occupants of a ping-pong buffer are steps t-2 and t, which differ in
// (t >> 1) & 1: a reader that finds the expected phase bit in a 2-byte value holds that step's value.  No flag word, no
// fence on the writer, no acquire on the reader, any access width; buffers start as 0x40 bytes (phase bit set, the
// first occupants carry phase 0).
constexpr uint32_t kPhaseMask = 0x40004000u;
__device__ __forceinline__ uint32_t phase_bits(int step) { return ((step >> 1) & 1) ? kPhaseMask : 0u; }
__device__ __forceinline__ void xchg_store(void* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t xchg_load(const void* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// bulk read of an exchange block: 0 = ld.relaxed.gpu (default), 1 = ld.global.cg (compiles to the same LDG.STRONG.GPU),
// 2 = cp.async.cg straight into the swizzled shared-memory position + in-place validation.  Measured at B = 64 / 384 words:
// 2.43 / 6.3 us per forward step with 0, 2.34 / 6.5 with 2 -- the fetch mechanism is not what bounds the step.
#ifndef PAULE_XCHG_LOAD
#define PAULE_XCHG_LOAD 0
#endif
__device__ __forceinline__ uint4 xchg_load4(const void* p) {
  uint4 v;
#if PAULE_XCHG_LOAD != 1
  asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
#else
  // L2-coherent weak load (bypasses the non-coherent L1): every value validates itself, so no ordering is needed
  asm volatile("ld.global.cg.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
#endif
  return v;
}
// 16-byte asynchronous copy global -> shared through L2 only (LDGSTS.BYPASS): not a "strong" access, so it streams at the
// normal load rate; the copied values validate themselves afterwards (phase bit), which is all the ordering they need
__device__ __forceinline__ void cp_async_cg16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// named barrier over a subset of the CTA's warps (id 1..15; id 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) {
  // 1 - 2/(1+e^{2x}); exact limits at +-inf, abs error ~1e-7
  return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x));
}

}  // namespace tc
}  // namespace paule
