// Implicit-GEMM 1-D convolution on the 5th-gen tensor cores (tcgen05 + TMEM), fed by TMA.
//
// This one kernel serves every dense contraction of the HiFi-GAN generator (reference:
// e2e_tts/models/vocoder/generator.py:37-53, layers.py:33-40):
//   * dilated Conv1d (resblock convs, conv_pre)             -> taps with row shifts (j-(k-1)/2)*d
//   * ConvTranspose1d(k=2u, stride=u, pad=u/2) (ups[i])      -> polyphase: N = u*C_out columns, every N-tile
//                                                              uses two taps with shifts {0,-1} or {0,+1}
// Data layout: activations are channels-last  [B][T][C]  bf16, so a tile of 128 consecutive time steps by 64
// channels is a K-major UMMA A operand (time on M, channels on K).  One CTA loads ONE slab of
// (128*mt + halo) rows per 64-channel panel with TMA (out-of-range rows are zero-filled by the TMA unit =
// Conv1d's per-layer zero padding) and serves every tap from that slab by offsetting the UMMA descriptor's
// start address by `shift` rows.  Weights are pre-packed on the host as ready-to-use swizzled smem images
// [n_tile][panel][tap][nt rows][row bytes] and streamed through an mbarrier ring with 1-D bulk copies.
// Accumulators (mt tiles of 128 x nt fp32) live in TMEM; four epilogue warps read them back with
// tcgen05.ld and fuse bias, residual add, the 3-way resblock sum / 3, LeakyReLU and the bf16 cast.
#pragma once
#include "ptx.cuh"

namespace e2e {

constexpr int kMaxTaps = 16;
constexpr int kMaxNTiles = 16;
constexpr int kMaxPanels = 8;
constexpr int kMaxStages = 8;
constexpr int kConvThreads = 192;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue

struct ConvParams {
  int T;                // time steps per utterance (rows); input and output have the same row count
  int B;                // utterances
  int panels;           // K panels (64 channels each, or one 32-channel panel)
  int rowb;             // bytes per panel row: 128 (64 ch, SWIZZLE_128B) or 64 (32 ch, SWIZZLE_64B)
  int nt;               // output columns per CTA (UMMA N), multiple of 16, <= 256
  int n_total;          // total output columns (row stride of the outputs)
  int mt;               // 128-row M tiles per CTA (mt*nt <= 512 TMEM columns)
  int taps;             // taps per N tile
  int hl;               // rows of left halo in the slab  (= max(0, -min shift))
  int slab_rows;        // rows per panel in shared memory (multiple of box_rows)
  int box_rows;         // TMA box height
  int tiles_per_chunk;  // weight tiles (one tap of one panel) per ring stage
  int n_chunks;         // ring transactions per CTA
  int n_stages;         // ring depth
  int stage_bytes;      // bytes per ring stage (multiple of 1024)
  int tmem_cols;        // power of two >= max(32, mt*nt)
  float divisor;        // epilogue: 0 = none, else out /= divisor (generator.py:48, xs / num_kernels)
  float slope;          // LeakyReLU slope applied to out_act
  int8_t shift[kMaxNTiles][kMaxTaps];  // row shift of each tap, per N tile
  const uint8_t* w;     // packed weights
  const float* bias;    // [n_total]
  const float* res_in;  // fp32 [B][T][n_total] residual (x of `xt + x`, layers.py:39) or nullptr
  const float* sum_in;  // fp32 running sum over resblocks (generator.py:44-47) or nullptr
  float* out_f32;       // fp32 [B][T][n_total] or nullptr
  __nv_bfloat16* out_act;  // bf16 [B][T][n_total] = leaky_relu(out, slope) or nullptr
};

#ifdef E2E_TRACE
// Debug build only: per-CTA phase timestamps (globaltimer ns) for every 8th CTA, read back by tests/cuda.
__device__ unsigned long long g_trace[512][12];
__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define E2E_TR(slot)                                               \
  do {                                                             \
    if (trace_id >= 0) g_trace[trace_id][slot] = gtime_ns();       \
  } while (0)
#else
#define E2E_TR(slot)
#endif

__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rowb = p.rowb;
  const int panel_bytes = p.slab_rows * rowb;
  const int tile_bytes = p.nt * rowb;
  const int total_tiles = p.panels * p.taps;

  uint8_t* slab = smem;
  uint8_t* ring = slab + p.panels * panel_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + p.n_stages * p.stage_bytes);
  uint64_t* panel_full = bars;                      // [kMaxPanels]
  uint64_t* ring_full = bars + kMaxPanels;          // [kMaxStages]
  uint64_t* ring_empty = ring_full + kMaxStages;    // [kMaxStages]
  uint64_t* acc_full = ring_empty + kMaxStages;     // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int t0 = blockIdx.x * (128 * p.mt);
  const int nti = blockIdx.y;
  const int b = blockIdx.z;
#ifdef E2E_TRACE
  const int lin_ = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  const int trace_id = (lin_ % 8 == 0 && lin_ / 8 < 512) ? lin_ / 8 : -1;
  if (threadIdx.x == 0) {
    E2E_TR(0);
    if (trace_id >= 0) {
      unsigned int smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      g_trace[trace_id][10] = smid;
    }
  }
#endif

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_in);
    for (int i = 0; i < p.panels; ++i) mbar_init(&panel_full[i], 1);
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(&ring_full[i], 1);
      mbar_init(&ring_empty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) E2E_TR(1);

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      const int ch_per_panel = rowb / 2;
      const int boxes = p.slab_rows / p.box_rows;
      for (int pn = 0; pn < p.panels; ++pn) {
        mbar_arrive_expect_tx(&panel_full[pn], panel_bytes);
        for (int bx = 0; bx < boxes; ++bx)
          tma_load_3d(slab + pn * panel_bytes + bx * p.box_rows * rowb, &tm_in, pn * ch_per_panel,
                      t0 - p.hl + bx * p.box_rows, b, &panel_full[pn]);
      }
      E2E_TR(2);
      const uint8_t* wsrc = p.w + static_cast<size_t>(nti) * total_tiles * tile_bytes;
      for (int c = 0; c < p.n_chunks; ++c) {
        const int stage = c % p.n_stages;
        const uint32_t par = (c / p.n_stages) & 1;
        mbar_wait(&ring_empty[stage], par ^ 1, 0x100 + stage);
        const int first = c * p.tiles_per_chunk;
        const int ntile = min(p.tiles_per_chunk, total_tiles - first);
        const uint32_t bytes = ntile * tile_bytes;
        mbar_arrive_expect_tx(&ring_full[stage], bytes);
        bulk_load_1d(ring + stage * p.stage_bytes, wsrc + static_cast<size_t>(first) * tile_bytes, bytes,
                     &ring_full[stage]);
      }
      E2E_TR(3);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      const uint32_t idesc = umma_idesc_bf16(128, p.nt);
      const uint32_t slab_addr = smem_u32(slab);
      const uint32_t ring_addr = smem_u32(ring);
      const int ksteps = rowb / 32;
      int tile = 0;
      for (int c = 0; c < p.n_chunks; ++c) {
        const int stage = c % p.n_stages;
        const uint32_t par = (c / p.n_stages) & 1;
        mbar_wait(&ring_full[stage], par, 0x200 + stage);
        tc_fence_after_sync();
        if (c == 0) E2E_TR(4);
        const int ntile = min(p.tiles_per_chunk, total_tiles - c * p.tiles_per_chunk);
        for (int i = 0; i < ntile; ++i, ++tile) {
          const int pn = tile / p.taps;
          const int tap = tile - pn * p.taps;
          if (tap == 0) {
            mbar_wait(&panel_full[pn], 0, 0x300 + pn);
            tc_fence_after_sync();
            if (pn == 0) E2E_TR(5);
          }
          const int row0 = p.hl + p.shift[nti][tap];
          const uint32_t a_base = slab_addr + pn * panel_bytes + row0 * rowb;
          const uint32_t b_base = ring_addr + stage * p.stage_bytes + i * tile_bytes;
          for (int m = 0; m < p.mt; ++m) {
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint32_t a_addr = a_base + m * 128 * rowb + ks * 32;
              const uint64_t da = umma_smem_desc(a_addr, rowb, 0);
              const uint64_t db = umma_smem_desc(b_base + ks * 32, rowb, 0);
              umma_bf16(tmem_base + m * p.nt, da, db, idesc, (tile | ks) != 0 ? 1u : 0u);
            }
          }
        }
        umma_commit(&ring_empty[stage]);  // frees the weight stage once these MMAs have read it
      }
      E2E_TR(6);
      umma_commit(acc_full);
    }
  } else {
    // ---------------- epilogue: TMEM -> registers -> global ----------------
    const int quarter = warp & 3;  // TMEM lanes 32*quarter .. +31 are the ones this warp may read
    mbar_wait(acc_full, 0, 0x400);
    tc_fence_after_sync();
    if (threadIdx.x == 64) E2E_TR(7);
    const int nchunk = p.nt / 32;
    for (int m = 0; m < p.mt; ++m) {
      const int t = t0 + m * 128 + quarter * 32 + lane;
      const bool valid = t < p.T;
      const size_t row_off = (static_cast<size_t>(b) * p.T + (valid ? t : 0)) * p.n_total + nti * p.nt;
      for (int cc = 0; cc < nchunk; ++cc) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + m * p.nt + cc * 32, v);
        tmem_ld_wait();
        if (valid) {
          const int n0 = nti * p.nt + cc * 32;
          const size_t off = row_off + cc * 32;
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(p.bias + n0 + i);
            f[i] = __uint_as_float(v[i]) + bv.x;
            f[i + 1] = __uint_as_float(v[i + 1]) + bv.y;
            f[i + 2] = __uint_as_float(v[i + 2]) + bv.z;
            f[i + 3] = __uint_as_float(v[i + 3]) + bv.w;
          }
          if (p.res_in) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 r = *reinterpret_cast<const float4*>(p.res_in + off + i);
              f[i] += r.x; f[i + 1] += r.y; f[i + 2] += r.z; f[i + 3] += r.w;
            }
          }
          if (p.sum_in) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 r = *reinterpret_cast<const float4*>(p.sum_in + off + i);
              f[i] += r.x; f[i + 1] += r.y; f[i + 2] += r.z; f[i + 3] += r.w;
            }
          }
          if (p.divisor != 0.f) {
            const float dv = p.divisor;
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = f[i] / dv;
          }
          if (p.out_f32) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<float4*>(p.out_f32 + off + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
          }
          if (p.out_act) {
            const float s = p.slope;
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              uint32_t pk[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float a = f[i + 2 * j], c = f[i + 2 * j + 1];
                a = a > 0.f ? a : a * s;
                c = c > 0.f ? c : c * s;
                __nv_bfloat162 h = __floats2bfloat162_rn(a, c);
                pk[j] = *reinterpret_cast<uint32_t*>(&h);
              }
              *reinterpret_cast<uint4*>(p.out_act + off + i) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
        }
      }
    }
  }

  if (threadIdx.x == 64) E2E_TR(8);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
  if (threadIdx.x == 0) E2E_TR(9);
}

}  // namespace e2e
