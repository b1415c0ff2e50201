// Host side of the tcgen05 convolution: TMA tensor-map creation (driver entry point fetched at run time,
// so the library does not link libcuda), launch-plan sizing and the weight packer that writes the swizzled
// shared-memory images the kernel streams with 1-D bulk copies.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "conv_tc.cuh"
#include "errors.h"

namespace e2e {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// Tensor map over a channels-last bf16 activation tensor [B][T][C]: dims (C, T, B), box (ch_box, box_rows, 1).
// ch_box = 64 -> SWIZZLE_128B, ch_box = 32 -> SWIZZLE_64B.  Out-of-bounds -> zeros.
inline int make_act_tensor_map(CUtensorMap* tm, const void* base, int B, int T, int C, int ch_box, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return fail(-10, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)T * C * 2};
  cuuint32_t box[3] = {(cuuint32_t)ch_box, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapSwizzle sw = ch_box == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) B=%d T=%d C=%d box=%dx%d", (int)r, B, T, C, ch_box,
             box_rows);
    return fail(-11, buf);
  }
  return 0;
}

// 2-D tensor map over a packed weight image: rows of `rowb` bytes, box = half a tile (nt/2 rows), no swizzle (the
// image already holds the swizzled bytes).
inline int make_weight_tensor_map(CUtensorMap* tm, const void* base, int rows, int rowb, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return fail(-10, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)(rowb / 2), (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)rowb};
  cuuint32_t box[2] = {(cuuint32_t)(rowb / 2), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(-11, "cuTensorMapEncodeTiled failed for a weight image");
  return 0;
}

// Static description of one convolution layer as the GEMM the kernel runs.
struct ConvShape {
  int cin;       // padded input channels (32, or a multiple of 64)
  int n_total;   // output columns (C_out, or u*C_out for polyphase ConvTranspose)
  int nt;        // columns per CTA
  int taps;      // taps per N tile
  std::vector<int> shifts;  // [n_tiles][taps] row shifts
};

constexpr int kSmemLimit = 232448;  // 227 KB opt-in maximum per CTA on sm_100

struct ConvPlan {
  ConvParams p{};
  CUtensorMap tm{};
  CUtensorMap tm_w{};  // packed weight image as a [rows][rowb] matrix (CTA-pair form only)
  CUtensorMap tm_out{};  // out_act as [B][T][n_total], box = 64 columns x 128 rows (staged kernels store through it)
  dim3 grid{};
  int smem_bytes = 0;
  int cg = 1;          // CTAs per MMA (tcgen05 cta_group): 2 = CTA pairs in 2-CTA clusters
  int staged = 0;      // 1: out_act leaves through shared memory and TMA stores (conv_tc.cuh, STAGED)
};

// CTA pairs for the unfused convolutions: only layers with a single N tile (both CTAs of a pair must use the same
// weight tiles), and where the weight stream + B operand reads dominate shared-memory traffic: wide N.
// E2E_CONV_CG=1|2 overrides for experiments.
inline int conv_cta_group(const ConvShape& s) {
  if (s.n_total != s.nt || s.nt < 128 || (s.nt / 2) % 8) return 1;
  const char* e = std::getenv("E2E_CONV_CG");
  if (e && (e[0] == '1' || e[0] == '2')) return e[0] - '0';
  return s.nt >= 256 && s.taps >= 7 ? 2 : 1;
}

// Fill every geometry field of plan.p from the layer shape and the problem size.  `mt_pref` = preferred
// number of 128-row tiles per unit (1, 2 or 4; reduced until TMEM and shared memory fit); `n_sms` sizes the
// persistent grid.
inline int plan_conv_impl(ConvPlan& plan, const ConvShape& s, int B, int T, int mt_pref, int n_sms, int cg,
                          int staged) {
  ConvParams& p = plan.p;
  if (cg == 0) cg = conv_cta_group(s);
  plan.cg = cg;
  if (staged && (s.nt % 64 != 0 || s.cin == 32)) staged = 0;  // 64-column groups; 128-byte-row kernels only
  plan.staged = staged;
  p.staged = staged;
  const int stage_bytes_out = staged ? kStageBufs * kStageBufBytes : 0;
  if (s.cin != 32 && s.cin % 64 != 0) return fail(-2, "cin must be 32 or a multiple of 64");
  if (s.nt % 32 != 0 || s.nt > 256 || s.n_total % s.nt != 0) return fail(-2, "bad N tiling");
  const int n_tiles = s.n_total / s.nt;
  if (n_tiles > kMaxNTiles || s.taps > kMaxTaps) return fail(-2, "too many N tiles / taps");
  p.T = T;
  p.B = B;
  p.rowb = s.cin == 32 ? 64 : 128;
  p.panels = s.cin == 32 ? 1 : s.cin / 64;
  if (p.panels > kMaxPanelSlots) return fail(-2, "too many K panels");
  p.nt = s.nt;
  p.n_total = s.n_total;
  p.n_tiles = n_tiles;
  p.taps = s.taps;
  int smin = 0, smax = 0;
  for (int i = 0; i < n_tiles; ++i)
    for (int j = 0; j < s.taps; ++j) {
      const int sh = s.shifts[i * s.taps + j];
      if (sh < -127 || sh > 127) return fail(-2, "tap shift out of range");
      p.shift[i][j] = (int8_t)sh;
      smin = sh < smin ? sh : smin;
      smax = sh > smax ? sh : smax;
    }
  p.hl = -smin;
  const int tile_bytes = s.nt * p.rowb / cg;  // per CTA: a pair splits every weight tile
  const int total_tiles = p.panels * s.taps;
  // ring stage ~ 32 KB (one tile when a tile is that large): fewer barrier round trips for the MMA issuer
  int tpc = 32768 / tile_bytes;
  if (tpc < 1) tpc = 1;
  if (tpc > total_tiles) tpc = total_tiles;
  p.tiles_per_chunk = tpc;
  p.n_chunks = (total_tiles + tpc - 1) / tpc;
  p.stage_bytes = tpc * tile_bytes;
  const int bar_bytes = 512;
  const int budget = kSmemLimit - 1024 - bar_bytes - stage_bytes_out;
  int mt0 = mt_pref >= 4 ? 4 : (mt_pref >= 2 ? 2 : 1);
  for (int mt = mt0; mt >= 1; mt >>= 1) {
    if (mt * s.nt > 512) continue;
    const int need = 128 * mt + p.hl + smax;
    for (int box = 128; box >= 16; box >>= 1) {
      const int rows = (need + box - 1) / box * box;
      if (rows - need > 32 && box > 16) continue;  // avoid large padding
      const int panel_bytes = rows * p.rowb;
      int max_slots = 2 * p.panels < kMaxPanelSlots ? 2 * p.panels : kMaxPanelSlots;
      for (int slots = max_slots; slots >= p.panels; --slots) {
        int stages = (budget - slots * panel_bytes) / p.stage_bytes;
        if (stages > 4) stages = 4;
        const int want = slots > p.panels ? 3 : 2;  // extra panel slots must not starve the weight ring
        if (stages < want) continue;
        p.mt = mt;
        p.slab_rows = rows;
        p.box_rows = box;
        p.panel_slots = slots;
        p.n_stages = stages;
        p.n_acc = 2 * mt * s.nt <= 512 ? 2 : 1;
        p.tiles_per_b = (T + 128 * mt - 1) / (128 * mt);
        p.n_units = B * p.tiles_per_b * n_tiles;
        plan.smem_bytes = 1024 + slots * panel_bytes + stages * p.stage_bytes + stage_bytes_out + bar_bytes;
        int grid = (p.n_units + cg - 1) / cg * cg;
        if (grid > n_sms) grid = n_sms / cg * cg;
        plan.grid = dim3(grid, 1, 1);
        return 0;
      }
    }
  }
  return fail(-3, "convolution does not fit shared memory / TMEM");
}

// `staged` = 1 asks for the staged TMA-store epilogue (bf16 output in the natural layout only; the caller then fills
// plan.tm_out with conv_output_map); layers whose slabs leave no room for the staging buffers fall back to direct stores.
inline int plan_conv(ConvPlan& plan, const ConvShape& s, int B, int T, int mt_pref, int n_sms = 148, int cg = 0,
                     int staged = 0) {
  if (staged && plan_conv_impl(plan, s, B, T, mt_pref, n_sms, cg, 1) == 0) return 0;
  return plan_conv_impl(plan, s, B, T, mt_pref, n_sms, cg, 0);
}

// E2E_NO_PDL=1 launches every kernel fully serialised (A/B experiments).
inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) on = std::getenv("E2E_NO_PDL") ? 0 : 1;
  return on == 1;
}

typedef void (*ConvKernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const ConvParams);

template <bool STAGED>
inline ConvKernelFn conv_kernel_for_t(int rowb, int mt, int cg) {
  if (cg == 2)  // CTA pairs: 64-channel panels only (wide layers)
    return mt == 4 ? conv_tc_kernel<128, 4, 2, STAGED>
                   : (mt == 2 ? conv_tc_kernel<128, 2, 2, STAGED> : conv_tc_kernel<128, 1, 2, STAGED>);
  return mt == 4 ? conv_tc_kernel<128, 4, 1, STAGED>
                 : (mt == 2 ? conv_tc_kernel<128, 2, 1, STAGED> : conv_tc_kernel<128, 1, 1, STAGED>);
}

inline ConvKernelFn conv_kernel_for(int rowb, int mt, int cg = 1, int staged = 0) {
  if (rowb == 128) return staged ? conv_kernel_for_t<true>(rowb, mt, cg) : conv_kernel_for_t<false>(rowb, mt, cg);
  return mt == 4 ? conv_tc_kernel<64, 4, 1, false>
                 : (mt == 2 ? conv_tc_kernel<64, 2, 1, false> : conv_tc_kernel<64, 1, 1, false>);
}

// Bias of the layer: host copy into the kernel parameters when it fits, device pointer otherwise.  `period` > 0: the
// bias repeats every `period` columns (polyphase ConvTranspose: column p * C_out + co) - only one period is passed.
inline void conv_set_bias(ConvPlan& plan, const float* d_bias, const float* h_bias, int n, int period = 0) {
  ConvParams& p = plan.p;
  p.bias = d_bias;
  p.bias_const = 0;
  p.bias_mask = 0x7fffffff;
  if (!h_bias) return;
  if (period >= 16 && period < n && (period & (period - 1)) == 0 && n % period == 0 && period <= kMaxBiasConst) {
    p.bias_mask = period - 1;
    n = period;
  }
  if (n > kMaxBiasConst) return;
  p.bias_const = 1;
  std::copy(h_bias, h_bias + n, p.cbias);
}

// Fills plan.tm_out for a staged plan (out = the bf16 [B][T][n_total] activation output).
inline int conv_output_map(ConvPlan& plan, const void* out, int B, int T) {
  if (!plan.staged) return 0;
  return make_act_tensor_map(&plan.tm_out, out, B, T, plan.p.n_total, 64, 128);
}

// Fills plan.tm_w (needed by the CTA-pair form; harmless otherwise).  w = packed image of the layer.
inline int conv_weight_map(ConvPlan& plan, const void* w) {
  const ConvParams& p = plan.p;
  const int rows = p.n_tiles * p.panels * p.taps * p.nt;
  return make_weight_tensor_map(&plan.tm_w, w, rows, p.rowb, p.nt / 2);
}

// Opt every instantiation into the 227 KB dynamic shared-memory limit (once per process and device).
inline int conv_kernels_init() {
  static int done_for_device = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (done_for_device == dev) return 0;
  const int rowbs[2] = {128, 64}, mts[3] = {1, 2, 4};
  for (int staged = 0; staged <= 1; ++staged)
    for (int cg = 1; cg <= 2; ++cg)
      for (int r : rowbs)
        for (int m : mts) {
          if ((cg == 2 || staged) && r == 64) continue;
          cudaError_t e = cudaFuncSetAttribute(conv_kernel_for(r, m, cg, staged),
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
          if (e != cudaSuccess) return fail((int)e, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
        }
  done_for_device = dev;
  return 0;
}

inline int launch_conv(const ConvPlan& plan, cudaStream_t st) {
  int rc = conv_kernels_init();
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_kernel_for(plan.p.rowb, plan.p.mt, plan.cg, plan.staged), plan.tm,
                                     plan.tm_w, plan.tm_out, plan.p);
  if (e != cudaSuccess) return fail((int)e, std::string("conv_tc launch: ") + cudaGetErrorString(e));
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, std::string("conv_tc launch: ") + cudaGetErrorString(e));
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Weight packing.  `wt(n, ci, tap_index_within_tile, n_tile)` semantics are supplied by the caller through a
// dense fp32 array  wg[n_total][taps][cin]  (already folded, already arranged per GEMM column): the packer
// only casts to bf16 and lays the bytes out as swizzled smem tiles [n_tile][panel][tap][nt][rowb].
// ---------------------------------------------------------------------------------------------------
inline uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
// fp32 -> fp16 (round to nearest even, saturating to +-65504 like the kernels' cvt.rn.satfinite; NaN kept)
inline uint16_t f32_to_f16_rn(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  x &= 0x7fffffffu;
  if (x > 0x7f800000u) return (uint16_t)(sign | 0x7e00u);      // NaN
  if (x >= 0x477ff000u) return (uint16_t)(sign | 0x7bffu);     // >= 65520 rounds past the largest finite half: saturate
  if (x < 0x38800000u) {                                       // below 2^-14: subnormal half (or zero)
    if (x < 0x33000000u) return (uint16_t)sign;                // < 2^-25 rounds to zero
    const int shift = 126 - (int)(x >> 23);                    // 14 .. 24 bits to drop
    uint32_t m = (x & 0x7fffffu) | 0x800000u;
    const uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
    m >>= shift;
    if (rem > half || (rem == half && (m & 1u))) ++m;
    return (uint16_t)(sign | m);
  }
  uint32_t h = ((x - 0x38000000u) >> 13);
  const uint32_t rem = x & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;
  return (uint16_t)(sign | h);
}
inline float bf16_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

inline size_t packed_weight_bytes(const ConvShape& s) {
  return (size_t)s.n_total * s.taps * s.cin * 2;
}

inline void pack_conv_weights(const ConvShape& s, const float* wg, uint8_t* out, int f16 = 0) {
  const int rowb = s.cin == 32 ? 64 : 128;
  const int panels = s.cin == 32 ? 1 : s.cin / 64;
  const int chp = rowb / 2;
  const uint32_t mask = rowb == 128 ? 7u : 3u;
  const int n_tiles = s.n_total / s.nt;
  const size_t tile_bytes = (size_t)s.nt * rowb;
  for (int nti = 0; nti < n_tiles; ++nti)
    for (int pn = 0; pn < panels; ++pn)
      for (int tap = 0; tap < s.taps; ++tap) {
        uint8_t* tile = out + (((size_t)nti * panels + pn) * s.taps + tap) * tile_bytes;
        for (int r = 0; r < s.nt; ++r) {
          const float* src = wg + ((size_t)(nti * s.nt + r) * s.taps + tap) * s.cin + pn * chp;
          for (int c = 0; c < chp; ++c) {
            const uint32_t off = swizzle_off((uint32_t)(r * rowb + c * 2), mask);
            const uint16_t h = f16 ? f32_to_f16_rn(src[c]) : f32_to_bf16_rn(src[c]);
            memcpy(tile + off, &h, 2);
          }
        }
      }
}

}  // namespace e2e
