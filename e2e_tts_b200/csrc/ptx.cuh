// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (tiled + bulk), tcgen05 (alloc / mma / commit /
// ld / fences) and the UMMA shared-memory / instruction descriptors.  Nothing here is reference-derived:
// the reference (InterlinkLabs/e2e-tts) is pure eager PyTorch and ships no native code.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <type_traits>

namespace e2e {

// ---------------------------------------------------------------------------------------------------
// Debug / watchdog: every spin in the kernels is bounded.  A wait that exceeds the budget records a code
// in this global and traps, so a protocol bug turns into a CUDA error instead of a hung GPU box.
// ---------------------------------------------------------------------------------------------------
__device__ unsigned int g_watchdog_code = 0;
__device__ unsigned int* g_watchdog_host = nullptr;  // optional zero-copy host word that survives the trap

#ifndef E2E_WATCHDOG_CYCLES
#define E2E_WATCHDOG_CYCLES (4000000000ll)   // ~2 s at 1.9 GHz
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

// ---------------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)  // (an explicit suspend-time hint was measured: 2 % slower, wake-up latency)
      : "memory");
  return ok != 0;
}
// Wait.  Production builds poll mbarrier.try_wait with a short nanosleep between polls: measured on B200 over the
// full forward (ms/step, same box): clock-bounded spin 5.29, bare try_wait spin 5.26, sleep 20 ns 5.255, sleep 64 ns
// 5.22 - the waiting producer / issuer / epilogue warps otherwise take issue slots and shared-memory (mbarrier)
// bandwidth from the warps that have work.  With -DE2E_WATCHDOG (the standalone device tests) every spin is
// bounded by the clock: a wait that exceeds the budget records `code` (the call site) in g_watchdog_code and traps.
#ifndef E2E_POLL_SLEEP
#define E2E_POLL_SLEEP 64
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t code) {
#ifdef E2E_WATCHDOG
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > E2E_WATCHDOG_CYCLES) {
      atomicCAS(&g_watchdog_code, 0u, code | 0x80000000u);
      if (g_watchdog_host) {
        *reinterpret_cast<volatile unsigned int*>(g_watchdog_host) = code | 0x80000000u;
      }
      __threadfence_system();
      __trap();
    }
  }
#elif E2E_POLL_SLEEP > 0
  (void)code;
  while (!mbar_try_wait(bar, parity)) __nanosleep(E2E_POLL_SLEEP);
#else
  (void)code;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "E2E_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra E2E_DONE;\n\t"
      "bra E2E_WAIT;\n\t"
      "E2E_DONE:\n\t"
      "}"
      ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
#endif
}

// ---------------------------------------------------------------------------------------------------
// Thread-block clusters (CTA pairs for tcgen05 cta_group::2)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// All threads of every CTA of the cluster.
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster.
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
// Wait with cluster-scope acquire: the barrier is arrived on by the peer CTA as well.
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, uint32_t code) {
#ifdef E2E_WATCHDOG
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > E2E_WATCHDOG_CYCLES) {
      atomicCAS(&g_watchdog_code, 0u, code | 0x80000000u);
      if (g_watchdog_host) {
        *reinterpret_cast<volatile unsigned int*>(g_watchdog_host) = code | 0x80000000u;
      }
      __threadfence_system();
      __trap();
    }
  }
#elif E2E_POLL_SLEEP > 0
  (void)code;
  while (!mbar_try_wait_cluster(bar, parity)) __nanosleep(E2E_POLL_SLEEP);
#else
  (void)code;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "E2E_WAITC:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra E2E_DONEC;\n\t"
      "bra E2E_WAITC;\n\t"
      "E2E_DONEC:\n\t"
      "}"
      ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
#endif
}

// ---------------------------------------------------------------------------------------------------
// Register re-balancing between warpgroups (all four warps of a warpgroup execute the same instruction).  The
// kernels launch 640 threads x 96 registers; the producer / issuer warpgroup gives most of its share back and the
// four epilogue warpgroups take 104 each (128 x 56 + 512 x 104 = 60 416 <= 61 440).
// ---------------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

// ---------------------------------------------------------------------------------------------------
// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start (CTA by CTA, as SMs free up) before the previous kernel of the stream has finished.  griddep_wait() blocks
// until that previous grid has completed and its memory is visible: every role calls it before its first access to
// an activation tensor, so barrier init, TMEM allocation, descriptor prefetch and the first weight tiles overlap the
// previous layer's tail.  griddep_launch() lets the NEXT kernel start early in the same way.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------
// Proxy / tcgen05 fences
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
// Bulk prefetch of a contiguous global range into L2 (no shared-memory destination, no completion to wait for):
// `bytes` a multiple of 16, `gptr` 16-byte aligned.
__device__ __forceinline__ void bulk_prefetch_l2(const void* gptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(gptr)), "r"(bytes) : "memory");
}
// 3-D tiled load: coordinates are (c0 = innermost element index, c1, c2); out-of-bounds elements
// (including negative coordinates) are zero-filled, which is exactly Conv1d's per-layer zero padding.
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2)
      : "memory");
}
// CTA-pair forms (tcgen05 cta_group::2): executed by both CTAs of the pair, each writes its OWN shared memory,
// and the transaction bytes are counted on the mbarrier of the pair's even CTA (bit 24 of a shared::cluster
// address selects the CTA inside the pair), where the one MMA-issuing thread waits.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2,
                                                 uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0),
        "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0),
        "r"(c1)
      : "memory");
}
// 3-D tiled STORE shared -> global (bulk async-group completion).  Rows / channels outside the tensor are clipped.
// The shared-memory source must have been written (and fence.proxy.async'ed) by the threads that produced it.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest `N` committed bulk groups have finished READING their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// ... have completed entirely (global writes performed)
template <int N>
__device__ __forceinline__ void bulk_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// 1-D bulk copy global -> shared (size multiple of 16 B, both addresses 16-B aligned).
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---------------------------------------------------------------------------------------------------
// TMEM allocation (warp-collective)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// CTA-pair forms: the same-numbered warp of BOTH CTAs executes them; both get the same TMEM address.
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// UMMA descriptors
// ---------------------------------------------------------------------------------------------------
// Shared-memory operand descriptor, K-major, swizzled (SWIZZLE_128B when a row is 128 B, SWIZZLE_64B when it
// is 64 B).  Rows are packed densely, 8-row groups are `8*row_bytes` apart (SBO).  The swizzle is a function
// of the absolute shared-memory address (bits [7,10) xor-ed into bits [4,7)), so any start address that is a
// whole number of rows (and/or 32-B K-steps) into a 1024-B-aligned buffer addresses the right bytes: this is
// what lets one resident activation slab serve every tap of a dilated convolution.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t row_bytes, uint32_t base_off) {
  const uint64_t layout = (row_bytes == 128) ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
  const uint64_t sbo = (8u * row_bytes) >> 4;
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);        // start address, bits [0,14)
  d |= static_cast<uint64_t>(1) << 16;                            // LBO (ignored for swizzled K-major)
  d |= sbo << 32;                                                 // SBO, bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                            // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(base_off & 7u) << 49;                // matrix base offset
  d |= layout << 61;                                              // swizzle mode
  return d;
}

// Instruction descriptor for kind::f16: BF16 x BF16 -> FP32 (or, f16 != 0, FP16 x FP16 -> FP32: same rate, 11
// significand bits instead of 8), both operands K-major.  Bits [7,10) / [10,13) = A / B format: 0 = f16, 1 = bf16.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N, int f16 = 0) {
  return (1u << 4) | (f16 ? 0u : ((1u << 7) | (1u << 10))) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// The 16-bit operand format of the tensor-core path (activations in HBM / shared memory, packed weights): bf16 by
// default, fp16 when the generator was created with operand_dtype = fp16.  Two values <-> one 32-bit word, first value in
// the low half.  The fp16 conversion saturates to +-65504 instead of producing inf.
// The format is a run-time property of the generator but uniform over a launch: the epilogues branch ONCE per
// 16-column item into code specialised on it (a per-value `f16 ? a : b` costs a predicated-off instruction per value).
template <bool F16>
__device__ __forceinline__ uint32_t pack16t(float a, float b) {
  if (F16) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
  }
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <bool F16>
__device__ __forceinline__ void unpack16t(uint32_t w, float& lo, float& hi) {
  if (F16) {
    const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&w));
    lo = v.x;
    hi = v.y;
  } else {
    lo = __uint_as_float(w << 16);   // bf16 -> fp32 is a 16-bit shift
    hi = __uint_as_float(w & 0xffff0000u);
  }
}
__device__ __forceinline__ uint32_t pack16(float a, float b, int f16) {
  return f16 ? pack16t<true>(a, b) : pack16t<false>(a, b);
}
__device__ __forceinline__ void unpack16(uint32_t w, float& lo, float& hi, int f16) {
  if (f16) unpack16t<true>(w, lo, hi);
  else unpack16t<false>(w, lo, hi);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the accumulate predicate fixed to true (no setp in the issue loop).
__device__ __forceinline__ void umma_bf16_acc(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.eq.b32 p, 0, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc)
      : "memory");
}
// CTA-pair MMA (M = 256: rows 0-127 from the even CTA's A operand, 128-255 from the odd CTA's; each CTA holds
// N/2 rows of B at the same shared-memory offset; each CTA's TMEM receives its own 128 rows).  Issued by one
// thread of the even CTA.
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_acc_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.eq.b32 p, 0, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc)
      : "memory");
}
// One lane of a converged warp (the elected leader issues the tcgen05 / TMA instructions).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xFFFFFFFF;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
// Arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// CTA-pair commit: arrives on the mbarrier at this shared-memory offset in BOTH CTAs of the pair.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .b16 m;\n\t"
      "mov.b16 m, 3;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t"
      "}"
      ::"r"(smem_u32(bar))
      : "memory");
}

// TMEM -> registers: each thread of the warp reads 32 consecutive fp32 columns of its own lane
// (lane = 32*(warp_id%4) + laneid).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 256-bit global accesses (LDG.E.256 / STG.E.256, 32-byte aligned): a lane moves one full 32-byte sector per
// instruction, which halves the L1TEX sector work of the row-per-lane epilogue accesses.
__device__ __forceinline__ void ld_global_256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p)
               : "memory");
}
__device__ __forceinline__ void st_global_256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

// Byte offset -> swizzled byte offset inside a 1024-B-aligned buffer (Swizzle<B,4,3>): mask 7 = 128B,
// 3 = 64B.  Used by everything that writes operand bytes with ordinary stores (weight packer, fused epilogues).
__host__ __device__ __forceinline__ uint32_t swizzle_off(uint32_t off, uint32_t mask) {
  return off ^ (((off >> 7) & mask) << 4);
}

}  // namespace e2e
