// Micro-benchmark of tcgen05.mma issue/execution rate from shared-memory operands (kind::f16, M=128, cta_group::1).
// Answers: how many cycles does one 128 x N x 16 MMA take when A starts (a) on an 8-row boundary and (b) on an
// arbitrary row of a swizzled slab, and how much does the single-thread issue loop add?
//   bench_umma            -> table over row bytes {128, 64, 32} x M in {128, 64} x N in {32,64,128,256} x {aligned,
//                            row-shifted} x {1 CTA, 148 CTAs}   (round 2 added the 32-byte rows and M = 64)
#include <cstdio>
#include <cstdlib>
#include "../../e2e_tts_b200/csrc/ptx.cuh"

using namespace e2e;

template <int ROWB>
__global__ void __launch_bounds__(128, 1) umma_rate(int M, int N, int shifted, int n_taps, int mt, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  constexpr int KS = ROWB / 32;
  constexpr uint32_t ROW16 = ROWB >> 4;
  constexpr uint32_t DESC_HI = ((8u * ROWB) >> 4) | (1u << 14) | ((ROWB == 128 ? 2u : (ROWB == 64 ? 4u : 6u)) << 29);
  // zero the operand area (A slab 600 rows, B tile 256 rows)
  for (int i = threadIdx.x; i < (600 + 256) * ROWB / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(M, N);
    const uint32_t a0 = ((smem_u32(smem) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t b0 = a0 + (600 * ROWB >> 4);
    const long long t0 = clock64();
    for (int tap = 0; tap < n_taps; ++tap) {
      const uint32_t a_lo = a0 + (shifted ? (tap % 11) + 1 : 8 * (tap % 8)) * ROW16;
      for (int m = 0; m < mt; ++m) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const uint64_t da = (static_cast<uint64_t>(DESC_HI) << 32) | (a_lo + m * (128 * ROW16) + ks * 2);
          const uint64_t db = (static_cast<uint64_t>(DESC_HI) << 32) | (b0 + ks * 2);
          umma_bf16(tmem + m * N, da, db, idesc, 1u);
        }
      }
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0, 0x900);
    const long long t2 = clock64();
    if (blockIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  unsigned long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(umma_rate<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(umma_rate<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(umma_rate<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("rowb    M    N  mt shifted ctas | mmas  issue_cyc/mma  total_cyc/mma  (ideal %s)\n", "max(128,M)*N/256");
  const int Ns[4] = {32, 64, 128, 256};
  for (int rowb : {128, 64, 32})
    for (int M : {128, 64})
      for (int ni = 0; ni < 4; ++ni)
        for (int shifted = 0; shifted < 2; ++shifted)
          for (int ctas : {1, 148}) {
            const int N = Ns[ni];
            if (M == 64 && (shifted || ctas == 1)) continue;   // M = 64: the aligned full-chip rows are enough
            const int mt = 512 / N > 4 ? 4 : 512 / N;
            const int taps = 64;
            const int n_mma = taps * mt * (rowb / 32);
            for (int rep = 0; rep < 2; ++rep) {
              if (rowb == 128)
                umma_rate<128><<<ctas, 128, 150 * 1024>>>(M, N, shifted, taps, mt, d);
              else if (rowb == 64)
                umma_rate<64><<<ctas, 128, 150 * 1024>>>(M, N, shifted, taps, mt, d);
              else
                umma_rate<32><<<ctas, 128, 150 * 1024>>>(M, N, shifted, taps, mt, d);
              cudaError_t e = cudaDeviceSynchronize();
              if (e != cudaSuccess) {
                printf("CUDA error: %s\n", cudaGetErrorString(e));
                return 1;
              }
            }
            unsigned long long h[2];
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("%4d %4d %4d %3d %7d %4d | %4d  %8.1f      %8.1f       (%d)\n", rowb, M, N, mt, shifted, ctas, n_mma,
                   (double)h[0] / n_mma, (double)h[1] / n_mma, 128 * N / 256);
          }
  return 0;
}
