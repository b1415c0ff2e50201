// Host side of the fused whole-ResBlock1 kernel (rb_tc.cuh): when to use it, geometry planning and launch.
#pragma once
#include <cstdlib>
#include "conv_host.cuh"
#include "rb_tc.cuh"

namespace e2e {

struct RbPlan {
  RbParams p{};
  CUtensorMap tm{};   // input activation [B][T][C]
  dim3 grid{};
  int smem_bytes = 0;
  int rowb = 128, mt = 1;
};

constexpr int kRbTailBytes = 512 + 1024;  // mbarriers (+ 1 KB spare)

// Receptive-field halo of the chain per side: sum_i (k-1)/2 * (d_i + 1).
inline int rb_halo(int k, const int* dil, int n_pairs) {
  int h = 0;
  for (int i = 0; i < n_pairs; ++i) h += (k - 1) / 2 * (dil[i] + 1);
  return h;
}

// The fused chain needs per lane TMEM acc + x of 128*MT x C fp32 each (two lanes: 4 * MT * C = 512 columns) and two
// slabs per lane: C in {32, 64, 128} with MT = 128 / C.  It recomputes the halo at both ends of every unit, so it is
// used where that is cheap: the k = 3 resblocks (H = 12 rows per side), whose pair launches are epilogue-bound.
// Same-box A/B of the whole forward in the >= 2 s power-capped regime (16 x 5 s, ms per pass;
// profiles/r02_experiments_notes.md §3): no fused chain 5.12, k = 3 at C <= 64 4.93, k = 3 at C <= 128 4.89 (the
// default), + k = 7 at C <= 64 4.99-5.00, + k = 11 at C = 32 5.09: for k >= 7 the launches are MMA-bound and every
// halo row costs.  (Isolated launches rank C = 128, k = 3 the other way round, 320 vs 302 us: the chain saves DRAM
// traffic and launches, which the power-capped, PDL-overlapped forward rewards.)
// E2E_RB_FUSION=0 disables the fused chain, =2 forces it wherever it fits; E2E_RB_KMAX / E2E_RB_CMAX move the rule.
inline bool rb_supported(int C, int k, const int* dil, int n_pairs) {
  static const char* e = std::getenv("E2E_RB_FUSION");
  if (e && e[0] == '0') return false;
  if (!(C == 32 || C == 64 || C == 128) || !(k & 1) || k > kMaxTaps || n_pairs < 1 || n_pairs > kRbMaxPairs) return false;
  int dmax = 1;
  for (int i = 0; i < n_pairs; ++i) dmax = dil[i] > dmax ? dil[i] : dmax;
  if ((k - 1) / 2 * dmax > 64) return false;
  const int mt = 128 / C;
  const int r_out = 128 * mt - 2 * rb_halo(k, dil, n_pairs);
  if (r_out < 32) return false;
  if (e && e[0] == '2') return true;
  static const char* ek = std::getenv("E2E_RB_KMAX");   // experiments: largest kernel size / channel count that is fused
  static const char* ec = std::getenv("E2E_RB_CMAX");
  const int kmax = ek ? atoi(ek) : 3, cmax = ec ? atoi(ec) : 128;
  return C <= cmax && k <= kmax && r_out * 2 >= 128 * mt;
}

inline int plan_rb(RbPlan& plan, int C, int k, const int* dil, int n_pairs, int B, int T, int n_sms = 148) {
  if (!(C == 32 || C == 64 || C == 128) || n_pairs < 1 || n_pairs > kRbMaxPairs)
    return fail(-2, "fused resblock: unsupported channel count / chain length");
  RbParams& p = plan.p;
  plan.rowb = C == 32 ? 64 : 128;
  plan.mt = 128 / C;
  const int rowb = plan.rowb, mt = plan.mt;
  p.T = T;
  p.B = B;
  p.panels = C == 32 ? 1 : C / 64;
  p.nt = C;
  p.taps = k;
  p.n_pairs = n_pairs;
  int dmax = 1;
  for (int i = 0; i < n_pairs; ++i) {
    p.dil[i] = dil[i];
    dmax = dil[i] > dmax ? dil[i] : dmax;
  }
  p.halo = rb_halo(k, dil, n_pairs);
  p.padr = ((k - 1) / 2 * dmax + 7) / 8 * 8;
  p.slab_rows = 128 * mt + 2 * p.padr;
  p.box_rows = 8;
  for (int b = 256; b >= 8; b -= 8)
    if (p.slab_rows % b == 0) {
      p.box_rows = b;
      break;
    }
  p.r_out = 128 * mt - 2 * p.halo;
  if (p.r_out < 8) return fail(-2, "fused resblock: halo larger than the unit");
  const int tile_bytes = C * rowb;
  const int total_tiles = p.panels * k;
  const int slabs = 4 * p.panels * p.slab_rows * rowb;   // P and Q, two lanes
  const int budget = kSmemLimit - 1024 - kRbTailBytes - slabs;
  for (int tpc = 32768 / tile_bytes > 0 ? 32768 / tile_bytes : 1; tpc >= 1; tpc >>= 1) {
    const int t = tpc > total_tiles ? total_tiles : tpc;
    const int stage_bytes = t * tile_bytes;
    int stages = budget / stage_bytes;
    if (stages > 4) stages = 4;
    if (stages < 2) continue;
    p.tiles_per_chunk = t;
    p.n_chunks = (total_tiles + t - 1) / t;
    p.n_stages = stages;
    p.stage_bytes = stage_bytes;
    p.tiles_per_b = (T + p.r_out - 1) / p.r_out;
    p.n_units = B * p.tiles_per_b;
    plan.smem_bytes = 1024 + slabs + stages * stage_bytes + kRbTailBytes;
    int grid = p.n_units;
    if (grid > n_sms) grid = n_sms;
    plan.grid = dim3(grid, 1, 1);
    return 0;
  }
  return fail(-3, "fused resblock does not fit shared memory");
}

typedef void (*RbKernelFn)(const CUtensorMap, const RbParams);

inline RbKernelFn rb_kernel_for(int rowb, int mt) {
  if (rowb == 64) return rb_tc_kernel<64, 4>;
  return mt == 2 ? rb_tc_kernel<128, 2> : rb_tc_kernel<128, 1>;
}

inline int rb_kernels_init() {
  static int done_for_device = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (done_for_device == dev) return 0;
  RbKernelFn fns[3] = {rb_kernel_for(64, 4), rb_kernel_for(128, 2), rb_kernel_for(128, 1)};
  for (RbKernelFn f : fns) {
    cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e != cudaSuccess) return fail((int)e, std::string("cudaFuncSetAttribute(rb): ") + cudaGetErrorString(e));
  }
  done_for_device = dev;
  return 0;
}

inline int launch_rb(const RbPlan& plan, cudaStream_t st) {
  int rc = rb_kernels_init();
  if (rc) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = plan.grid;
  cfg.blockDim = dim3(kConvThreads, 1, 1);
  cfg.dynamicSmemBytes = plan.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // see griddep_wait() in ptx.cuh
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, rb_kernel_for(plan.rowb, plan.mt), plan.tm, plan.p);
  if (e != cudaSuccess) return fail((int)e, std::string("rb_tc launch: ") + cudaGetErrorString(e));
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, std::string("rb_tc launch: ") + cudaGetErrorString(e));
  return 0;
}

}  // namespace e2e
