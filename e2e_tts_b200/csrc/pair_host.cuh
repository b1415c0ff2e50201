// Host side of the fused residual-pair kernel (pair_tc.cuh): geometry planning and launch.
#pragma once
#include "conv_host.cuh"
#include "pair_tc.cuh"

namespace e2e {

struct PairPlan {
  PairParams p{};
  CUtensorMap tm{};
  dim3 grid{};
  int smem_bytes = 0;
  int rowb = 128, mt = 1;
};

// The fused pair needs 4 TMEM accumulators of 128*MT x C fp32 (two units in flight x two convs) and two input
// slabs + two intermediate slabs in shared memory: supported for C in {32, 64, 128} with MT = 128 / C.
inline bool pair_supported(int C, int k, int d) {
  return (C == 32 || C == 64 || C == 128) && (k & 1) && k <= kMaxTaps && (k - 1) * d <= 120;
}

constexpr int kPairTailBytes = 512 + 1024;  // mbarriers + the two staged bias vectors

inline int plan_pair(PairPlan& plan, int C, int k, int d, int B, int T, int n_sms = 148) {
  if (!pair_supported(C, k, d)) return fail(-2, "fused pair: unsupported channel count / kernel size");
  PairParams& p = plan.p;
  plan.rowb = C == 32 ? 64 : 128;
  plan.mt = 128 / C;
  const int rowb = plan.rowb, mt = plan.mt;
  p.T = T;
  p.B = B;
  p.panels = C == 32 ? 1 : C / 64;
  p.nt = C;
  p.taps = k;
  p.dil = d;
  p.r_out = 128 * mt - (k - 1);
  p.m_rows = (128 * mt + (k - 1) + 15) / 16 * 16;
  const int need = 128 * mt + (k - 1) * d;
  const int tile_bytes = C * rowb;
  const int total_tiles = p.panels * k;
  const int budget = kSmemLimit - 1024 - kPairTailBytes;
  for (int box = 128; box >= 16; box >>= 1) {
    const int rows = (need + box - 1) / box * box;
    if (rows - need > 32 && box > 16) continue;
    const int slabs = 2 * p.panels * (rows + p.m_rows) * rowb;
    for (int tpc = 32768 / tile_bytes > 0 ? 32768 / tile_bytes : 1; tpc >= 1; tpc >>= 1) {
      int t = tpc > total_tiles ? total_tiles : tpc;
      const int stage_bytes = t * tile_bytes;
      int stages = (budget - slabs) / stage_bytes;
      if (stages > 4) stages = 4;
      if (stages < 2) continue;
      p.a_rows = rows;
      p.box_rows = box;
      p.tiles_per_chunk = t;
      p.n_chunks = (total_tiles + t - 1) / t;
      p.n_stages = stages;
      p.stage_bytes = stage_bytes;
      p.tiles_per_b = (T + p.r_out - 1) / p.r_out;
      p.n_units = B * p.tiles_per_b;
      plan.smem_bytes = 1024 + slabs + stages * stage_bytes + kPairTailBytes;
      plan.grid = dim3(p.n_units < n_sms ? p.n_units : n_sms, 1, 1);
      return 0;
    }
  }
  return fail(-3, "fused pair does not fit shared memory");
}

typedef void (*PairKernelFn)(const CUtensorMap, const PairParams);

inline PairKernelFn pair_kernel_for(int rowb, int mt) {
  if (rowb == 64) return pair_tc_kernel<64, 4>;
  return mt == 2 ? pair_tc_kernel<128, 2> : pair_tc_kernel<128, 1>;
}

inline int pair_kernels_init() {
  static int done_for_device = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (done_for_device == dev) return 0;
  PairKernelFn fns[3] = {pair_tc_kernel<64, 4>, pair_tc_kernel<128, 2>, pair_tc_kernel<128, 1>};
  for (PairKernelFn f : fns) {
    cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e != cudaSuccess) return fail((int)e, std::string("cudaFuncSetAttribute(pair): ") + cudaGetErrorString(e));
  }
  done_for_device = dev;
  return 0;
}

inline int launch_pair(const PairPlan& plan, cudaStream_t st) {
  int rc = pair_kernels_init();
  if (rc) return rc;
  pair_kernel_for(plan.rowb, plan.mt)<<<plan.grid, kConvThreads, plan.smem_bytes, st>>>(plan.tm, plan.p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, std::string("pair_tc launch: ") + cudaGetErrorString(e));
  return 0;
}

}  // namespace e2e
