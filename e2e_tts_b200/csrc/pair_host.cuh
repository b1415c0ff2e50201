// Host side of the fused residual-pair kernel (pair_tc.cuh): geometry planning and launch.
#pragma once
#include <cstdlib>
#include "conv_host.cuh"
#include "pair_tc.cuh"

namespace e2e {

struct PairPlan {
  PairParams p{};
  CUtensorMap tm{};            // input activation [B][T][C]
  CUtensorMap tm_w1{}, tm_w2{};  // packed weight images as [rows][rowb] matrices (CTA-pair form only)
  CUtensorMap tm_out{}, tm_out2{};  // output activation [B][T][C]: boxes of min(r_out, 256) / r_out - 256 rows
  dim3 grid{};
  int smem_bytes = 0;
  int rowb = 128, mt = 1;
  int cg = 1;                  // CTAs per MMA (tcgen05 cta_group): 2 = CTA pairs in 2-CTA clusters
  bool staged = false;         // c2 epilogue staged in shared memory + TMA store (see pair_tc.cuh)
};

// CTA pairs pay off where the MMAs dominate and the one-CTA form is bound by shared-memory traffic (B operand
// reads + the weight stream).  Measured in the full forward on B200 (16 x 5 s, per launch, one-CTA -> pair):
//   C = 128: k = 11  307 -> 257 us, k = 7  201 -> 190 us, k = 3  122 -> 151 us (epilogue-latency bound: loses to the
//   C =  64: k = 11  182 -> 169 us, k = 7  124 -> 151 us                         extra cross-CTA hops)
//   C =  32: k = 11  147 -> 139 us, k = 7  106 -> 134 us
// E2E_PAIR_CG=1|2 overrides for experiments.
inline int pair_cta_group(int C, int k) {
  const char* e = std::getenv("E2E_PAIR_CG");
  if (e && (e[0] == '1' || e[0] == '2')) return e[0] - '0';
  if (C >= 128) return k >= 7 ? 2 : 1;
  return k >= 11 ? 2 : 1;
}

// The fused pair needs 4 TMEM accumulators of 128*MT x C fp32 (two units in flight x two convs) and two input
// slabs + two intermediate slabs in shared memory: supported for C in {32, 64, 128} with MT = 128 / C.
inline bool pair_supported(int C, int k, int d) {
  return (C == 32 || C == 64 || C == 128) && (k & 1) && k <= kMaxTaps && (k - 1) * d <= 120;
}

constexpr int kPairTailBytes = 512 + 1024;  // mbarriers (+ 1 KB spare; the biases are kernel parameters)

inline int plan_pair(PairPlan& plan, int C, int k, int d, int B, int T, int n_sms = 148) {
  if (!pair_supported(C, k, d)) return fail(-2, "fused pair: unsupported channel count / kernel size");
  PairParams& p = plan.p;
  plan.rowb = C == 32 ? 64 : 128;
  plan.mt = 128 / C;
  plan.cg = pair_cta_group(C, k);
  {
    const char* e = std::getenv("E2E_PAIR_STAGED");   // 0 | 1 overrides for experiments
    // Off by default since the biases moved to the constant bank: per launch, staged -> direct stores,
    // C = 128 k = 3 111.5 -> 105.8 us, C = 64 k = 3 95 -> 82.5 us (before that change staged won: 125 -> 111, 106 -> 100).
    plan.staged = e ? (e[0] == '1' && C >= 64) : false;
  }
  const int rowb = plan.rowb, mt = plan.mt, cg = plan.cg;
  p.T = T;
  p.B = B;
  p.panels = C == 32 ? 1 : C / 64;
  p.nt = C;
  p.taps = k;
  p.dil = d;
  p.r_out = 128 * mt - (k - 1);
  p.m_rows = (128 * mt + (k - 1) + 15) / 16 * 16;
  const int need = 128 * mt + (k - 1) * d;
  const int tile_bytes = C * rowb / cg;  // per CTA: a pair splits every weight tile
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
      int grid = (p.n_units + cg - 1) / cg * cg;
      if (grid > n_sms) grid = n_sms / cg * cg;
      plan.grid = dim3(grid, 1, 1);
      return 0;
    }
  }
  return fail(-3, "fused pair does not fit shared memory");
}

// Fills plan.tm_w1 / tm_w2 (needed by the CTA-pair form; harmless otherwise).  w1 / w2 = packed images of c1 / c2.
inline int pair_weight_maps(PairPlan& plan, const void* w1, const void* w2) {
  const PairParams& p = plan.p;
  const int rows = p.panels * p.taps * p.nt;
  int rc = make_weight_tensor_map(&plan.tm_w1, w1, rows, plan.rowb, p.nt / 2);
  if (rc) return rc;
  return make_weight_tensor_map(&plan.tm_w2, w2, rows, plan.rowb, p.nt / 2);
}

// Fills plan.tm_out / tm_out2: the result tile of a unit (r_out rows) leaves shared memory as TMA stores of one
// 64-channel (or 32-channel) panel each; a box is at most 256 rows high, so tall units (C = 32: r_out ~ 500) use a
// second map for the rows past 256 (its shared-memory source starts on an 8-row boundary, as the swizzle needs).
inline int pair_output_maps(PairPlan& plan, const void* out, int B, int T, int C) {
  const int r1 = plan.p.r_out < 256 ? plan.p.r_out : 256;
  int rc = make_act_tensor_map(&plan.tm_out, out, B, T, C, plan.rowb / 2, r1);
  if (rc) return rc;
  const int r2 = plan.p.r_out - r1;
  return make_act_tensor_map(&plan.tm_out2, out, B, T, C, plan.rowb / 2, r2 > 0 ? r2 : 8);
}

typedef void (*PairKernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                             const CUtensorMap, const PairParams);

inline PairKernelFn pair_kernel_for(int rowb, int mt, int cg, bool staged = false) {
  if (rowb == 64) return cg == 2 ? pair_tc_kernel<64, 4, 2, false> : pair_tc_kernel<64, 4, 1, false>;
  if (staged) {
    if (cg == 2) return mt == 2 ? pair_tc_kernel<128, 2, 2, true> : pair_tc_kernel<128, 1, 2, true>;
    return mt == 2 ? pair_tc_kernel<128, 2, 1, true> : pair_tc_kernel<128, 1, 1, true>;
  }
  if (cg == 2) return mt == 2 ? pair_tc_kernel<128, 2, 2, false> : pair_tc_kernel<128, 1, 2, false>;
  return mt == 2 ? pair_tc_kernel<128, 2, 1, false> : pair_tc_kernel<128, 1, 1, false>;
}

inline int pair_kernels_init() {
  static int done_for_device = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (done_for_device == dev) return 0;
  PairKernelFn fns[10] = {pair_kernel_for(64, 4, 1),         pair_kernel_for(64, 4, 2),
                          pair_kernel_for(128, 2, 1, false), pair_kernel_for(128, 1, 1, false),
                          pair_kernel_for(128, 2, 2, false), pair_kernel_for(128, 1, 2, false),
                          pair_kernel_for(128, 2, 1, true),  pair_kernel_for(128, 1, 1, true),
                          pair_kernel_for(128, 2, 2, true),  pair_kernel_for(128, 1, 2, true)};
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
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = plan.grid;
  cfg.blockDim = dim3(kConvThreads, 1, 1);
  cfg.dynamicSmemBytes = plan.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = plan.cg;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // see griddep_wait() in ptx.cuh
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, pair_kernel_for(plan.rowb, plan.mt, plan.cg, plan.staged), plan.tm, plan.tm_w1, plan.tm_w2,
                                     plan.tm_out, plan.tm_out2, plan.p);
  if (e != cudaSuccess) return fail((int)e, std::string("pair_tc launch: ") + cudaGetErrorString(e));
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, std::string("pair_tc launch: ") + cudaGetErrorString(e));
  return 0;
}

}  // namespace e2e
