// Host side of the four-time-steps-per-row fused pair kernel of the C = 32 stage (pair_tz.cuh): Toeplitz window packer,
// geometry planning and launch.
#pragma once
#include <cstdlib>
#include "conv_host.cuh"
#include "pair_tz.cuh"

namespace e2e {

struct TzPlan {
  TzParams p{};
  CUtensorMap tm{};   // input activation viewed as [B][T/4][128]
  dim3 grid{};
  int smem_bytes = 0;
};

constexpr int kTzTailBytes = 256 + 1024;  // mbarriers (+ 1 KB spare)

// Bytes of the sliding-window array of a k-tap, dilation-1 convolution: blocks y = -(h + 3) .. h + 3, h = (k - 1) / 2.
inline size_t tz_window_bytes(int k) { return (size_t)(k + 2 * (kTzG - 1)) * kTzBlock; }

// Sliding-window weight array (pair_tz.cuh): block index y + h + 3 holds W[tap offset -y] = tap m = h - y as a
// [32 co][32 ci] K-major tile of 64-byte rows in the SWIZZLE_64B pattern; the 3 blocks at either end are zero.
// wg = [co][tap][ci] fp32 (folded weights), exactly what pack_conv_weights takes.
inline void pack_tz_window(const float* wg, int k, uint8_t* out, int f16 = 0) {
  const int h = (k - 1) / 2;
  const int blocks = k + 2 * (kTzG - 1);
  memset(out, 0, (size_t)blocks * kTzBlock);
  for (int y = -h; y <= h; ++y) {
    const int m = h - y;
    uint8_t* blk = out + (size_t)(y + h + kTzG - 1) * kTzBlock;
    for (int co = 0; co < kTzC; ++co)
      for (int ci = 0; ci < kTzC; ++ci) {
        const float v = wg[((size_t)co * k + m) * kTzC + ci];
        const uint16_t hv = f16 ? f32_to_f16_rn(v) : f32_to_bf16_rn(v);
        memcpy(blk + swizzle_off((uint32_t)(co * 64 + ci * 2), 3u), &hv, 2);
      }
  }
}

// c2 loses (k - 1) / 2 time steps per side (c1 reads real context rows of the slab): whole super-rows are dropped.
inline int tz_halo(int k) { return ((k - 1) / 2 + kTzG - 1) / kTzG; }

inline int tz_w1_bytes(int k, int d) { return d == 1 ? (int)tz_window_bytes(k) : k * kTzBlock; }

// super-rows of context on either side of the 128 computed rows: the farthest tap shift, rounded up to a multiple of 4
inline int tz_padr(int k, int d) {
  const int reach = ((k - 1) / 2 * d + kTzG - 1) / kTzG;
  const int r = (reach + 3) / 4 * 4;
  return r < 4 ? 4 : r;
}

inline int tz_smem_bytes(int k, int d) {
  const int slab_rows = 128 + 2 * tz_padr(k, d);
  return 1024 + 4 * 2 * slab_rows * 128 + tz_w1_bytes(k, d) + (int)tz_window_bytes(k) + kTzTailBytes;
}

// C = 32 stages whose length is a multiple of 4 (always: the stage length is T_mel times the upsampling product).
// E2E_TZ=0 disables the kernel (A/B switch).
inline bool tz_supported(int C, int k, int d, int T) {
  static const char* e = std::getenv("E2E_TZ");
  if (e && e[0] == '0') return false;
  if (C != kTzC || !(k & 1) || k < 3 || k > kMaxTaps || d < 1 || T % kTzG) return false;
  if (128 - 2 * tz_halo(k) < 64 || tz_padr(k, d) > 32) return false;
  return tz_smem_bytes(k, d) <= kSmemLimit;
}

inline int plan_tz(TzPlan& plan, int k, int d, int B, int T, int n_sms = 148) {
  if (!(k & 1) || k < 3 || d < 1 || T % kTzG) return fail(-2, "pair_tz: unsupported kernel size / dilation / stage length");
  TzParams& p = plan.p;
  p.T = T;
  p.T4 = T / kTzG;
  p.B = B;
  p.taps = k;
  p.dil = d;
  p.halo = tz_halo(k);
  p.padr = tz_padr(k, d);
  p.slab_rows = 128 + 2 * p.padr;
  p.r_out = 128 - 2 * p.halo;
  if (p.r_out < 8 || p.slab_rows > 256) return fail(-2, "pair_tz: halo larger than the unit");
  p.w1_toep = d == 1;
  p.w1_bytes = tz_w1_bytes(k, d);
  p.w2_bytes = (int)tz_window_bytes(k);
  p.dbg = 0;
  plan.smem_bytes = tz_smem_bytes(k, d);
  if (plan.smem_bytes > kSmemLimit) return fail(-3, "pair_tz does not fit shared memory");
  p.tiles_per_b = (p.T4 + p.r_out - 1) / p.r_out;
  p.n_units = B * p.tiles_per_b;
  int grid = p.n_units;
  if (grid > n_sms) grid = n_sms;
  plan.grid = dim3(grid, 1, 1);
  return 0;
}

// Tensor map of the launch: the [B][T][32] activation as [B][T/4][128], one box = one 64-super-channel panel of a slab.
inline int tz_input_map(TzPlan& plan, const void* in) {
  return make_act_tensor_map(&plan.tm, in, plan.p.B, plan.p.T4, 128, 64, plan.p.slab_rows);
}

inline int tz_kernels_init() {
  static int done_for_device = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (done_for_device == dev) return 0;
  cudaError_t e = cudaFuncSetAttribute(pair_tz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
  if (e != cudaSuccess) return fail((int)e, std::string("cudaFuncSetAttribute(pair_tz): ") + cudaGetErrorString(e));
  done_for_device = dev;
  return 0;
}

inline int launch_tz(const TzPlan& plan, cudaStream_t st) {
  int rc = tz_kernels_init();
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, pair_tz_kernel, plan.tm, plan.p);
  if (e != cudaSuccess) return fail((int)e, std::string("pair_tz launch: ") + cudaGetErrorString(e));
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, std::string("pair_tz launch: ") + cudaGetErrorString(e));
  return 0;
}

}  // namespace e2e
