// Implicit-GEMM 1-D convolution on the 5th-gen tensor cores (tcgen05 + TMEM), fed by TMA.
//
// This one kernel serves every dense contraction of the HiFi-GAN generator (reference:
// e2e_tts/models/vocoder/generator.py:37-53, layers.py:33-40):
//   * dilated Conv1d (resblock convs, conv_pre)             -> taps with row shifts (j-(k-1)/2)*d
//   * ConvTranspose1d(k=2u, stride=u, pad=u/2) (ups[i])      -> polyphase: N = u*C_out columns, every N-tile
//                                                              uses two taps with shifts {0,-1} or {0,+1}
// Data layout: activations are channels-last  [B][T][C]  bf16, so a tile of 128 consecutive time steps by 64
// channels is a K-major UMMA A operand (time on M, channels on K).  A work unit is (utterance, tile of 128*MT
// time steps, N tile).  For every unit ONE slab of (128*MT + halo) rows per 64-channel panel is brought in by
// TMA (out-of-range rows are zero-filled by the TMA unit = Conv1d's per-layer zero padding) and serves every
// tap by offsetting the UMMA descriptor's start address by `shift` rows.  Weights are pre-packed on the host
// as ready-to-use swizzled smem images [n_tile][panel][tap][nt rows][row bytes] and streamed with 1-D bulk
// copies.
//
// The kernel is persistent (one CTA per SM, units strided over CTAs) and warp-specialised:
//   warp 0   panel producer : TMA slab loads into a ring of panel buffers (runs ahead across units)
//   warp 1   weight producer: bulk copies into the weight ring
//   warp 2   MMA issuer     : one thread issues tcgen05.mma; the (M tile, K step) loops are unrolled and the
//                             descriptors are advanced by adding to their low word only
//   warp 3   TMEM allocator; in STAGED kernels also the store warp (TMA bulk-tensor stores of staged output tiles)
//   warps 4-19 epilogue     : tcgen05.ld the fp32 accumulators (four warps per TMEM lane quarter, 16-column
//                             items), fuse bias (kernel parameters = constant bank), residual add (prefetched
//                             while the MMAs run), resblock sum, /3, LeakyReLU, casts (epilogue.cuh)
// Accumulators are double-buffered in TMEM when 2*MT*nt <= 512 columns, so the epilogue of unit i overlaps
// the MMAs of unit i+1.
//
// CG = 2 runs the pipeline on a CTA PAIR (2-CTA cluster, tcgen05 cta_group::2), for layers with one N tile: each
// CTA owns its units, its slab panels and its epilogue; one thread of the even CTA issues M = 256 MMAs whose rows
// 0-127 / 128-255 are the two CTAs' units and whose B operand (the weight tile) is split between them, nt/2 rows
// each.  Per CTA that halves the weight bytes streamed into shared memory for every unit and the B bytes every MMA
// reads back - at C = 256 the weight stream alone (1.4 MB per 128-row unit and k = 11) is two thirds of the MMA
// operand traffic, and shared-memory bandwidth is what bounds these kernels (profiles/r01_experiments_notes.md).
#pragma once
#include "epilogue.cuh"

namespace e2e {

constexpr int kMaxTaps = 16;
constexpr int kMaxNTiles = 16;
constexpr int kMaxPanelSlots = 8;
constexpr int kMaxStages = 8;
constexpr int kMaxBiasConst = 512;        // biases up to this many columns travel as kernel parameters (constant bank)
constexpr int kStageBufs = 2;             // staged-store buffers (STAGED kernels)
constexpr int kStageBufBytes = 128 * 128; // one 128-row x 64-column bf16 tile, SWIZZLE_128B

struct ConvParams {
  int T;                // time steps per utterance (rows); input and output have the same row count
  int B;                // utterances
  int panels;           // K panels (64 channels each, or one 32-channel panel)
  int rowb;             // bytes per panel row: 128 (64 ch, SWIZZLE_128B) or 64 (32 ch, SWIZZLE_64B)
  int nt;               // output columns per unit (UMMA N), multiple of 32, <= 256
  int n_total;          // total output columns (row stride of the outputs)
  int n_tiles;          // n_total / nt
  int mt;               // 128-row M tiles per unit (== template MT)
  int taps;             // taps per N tile
  int hl;               // rows of left halo in the slab  (= max(0, -min shift))
  int slab_rows;        // rows per panel in shared memory (multiple of box_rows)
  int box_rows;         // TMA box height
  int panel_slots;      // panel buffers in the ring (>= panels)
  int tiles_per_chunk;  // weight tiles (one tap of one panel) per ring stage
  int n_chunks;         // ring transactions per unit
  int n_stages;         // ring depth
  int stage_bytes;      // bytes per ring stage (multiple of 1024)
  int n_acc;            // TMEM accumulator sets (1 or 2)
  int tiles_per_b;      // time tiles per utterance
  int n_units;          // B * tiles_per_b * n_tiles
  float divisor;        // epilogue: 0 = none, else out /= divisor (generator.py:48, xs / num_kernels)
  float slope;          // LeakyReLU slope applied to out_act
  int act_tanh;         // 1: out_act = tanh(result) instead (Postnet)
  int sum_tiled;        // sum_a is in the tiled8 layout (epilogue.cuh)
  int out_tiled;        // out_act is written in the tiled8 layout
  int f16;              // 16-bit tensors and operands are fp16 instead of bf16 (ptx.cuh pack16)
  int staged;           // == template STAGED: out_act leaves through shared memory and TMA stores (see the kernel)
  int8_t shift[kMaxNTiles][kMaxTaps];  // row shift of each tap, per N tile
  const uint8_t* w;     // packed weights
  const float* bias;    // [n_total] (device memory: used when n_total > kMaxBiasConst)
  int bias_const;       // 1: cbias holds the bias - warp-uniform constant-bank loads, off the L1 data pipe the
  int bias_mask;        //    tensor core's operand reads and every other epilogue access share; column n reads
  float cbias[kMaxBiasConst];  // cbias[n & bias_mask] (the polyphase upsamplers repeat C_out biases per phase)
  const __nv_bfloat16* res_act;  // bf16 [B][T][n_total] = leaky_relu(x, 1/res_inv_slope) of the residual x of
                                 // `xt + x` (layers.py:39), or nullptr; x is recovered by the inverse LeakyReLU
  float res_inv_slope;  // 1 / slope used when res_act was written (10 for LRELU_SLOPE = 0.1)
  const __nv_bfloat16* sum_a;  // bf16 [B][T][n_total] running sum over the stage's resblocks (generator.py:44-47), or
                               // nullptr
  float* out_f32;       // fp32 [B][T][n_total] or nullptr
  __nv_bfloat16* out_act;  // bf16 [B][T][n_total] = leaky_relu(out, slope) or nullptr
};

#ifdef E2E_TRACE
// Debug build only: per-CTA phase timestamps (globaltimer ns), read back by tests/cuda.
__device__ unsigned long long g_trace[512][16];  // [12], [13]: clock64 at kernel start / end
__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define E2E_TR(slot)                                                     \
  do {                                                                   \
    if (blockIdx.x < 512) {                                              \
      g_trace[blockIdx.x][slot] = gtime_ns();                            \
      if ((slot) == 0) g_trace[blockIdx.x][12] = clock64();              \
      if ((slot) == 7) g_trace[blockIdx.x][13] = clock64();              \
    }                                                                    \
  } while (0)
#else
#define E2E_TR(slot)
#endif

#ifdef E2E_TRACE2
// Debug build only: SM-clock cycles spent in each segment of the epilogue loop by one epilogue thread (warp 4,
// lane 0) of every CTA, accumulated in registers and written once at the end (E2E_TR2_FLUSH).
__device__ unsigned int g_trace2[512][24];
#define E2E_TR2_DECL                                                      \
  const bool tr2_on = threadIdx.x == 128 && blockIdx.x < 512;             \
  unsigned int tr2_acc[16];                                               \
  _Pragma("unroll") for (int i_ = 0; i_ < 16; ++i_) tr2_acc[i_] = 0;      \
  unsigned int tr2_last = (unsigned int)clock64();
#define E2E_TR2(k)                                                        \
  do {                                                                    \
    const unsigned int c_ = (unsigned int)clock64();                      \
    tr2_acc[k] += c_ - tr2_last;                                          \
    tr2_last = c_;                                                        \
  } while (0)
#define E2E_TR2_FLUSH                                                     \
  if (tr2_on) {                                                           \
    _Pragma("unroll") for (int i_ = 0; i_ < 16; ++i_) g_trace2[blockIdx.x][i_] = tr2_acc[i_]; \
  }
#else
#define E2E_TR2_DECL
#define E2E_TR2(k)
#define E2E_TR2_FLUSH
#endif

template <int ROWB, int MT, int CG, bool STAGED>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w,
               const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  constexpr int KS = ROWB / 32;          // UMMA K steps (16 bf16 = 32 B) per panel row
  constexpr uint32_t ROW16 = ROWB >> 4;  // row pitch in descriptor address units (16 B)
  // high word of the K-major swizzled smem descriptor: SBO = 8 rows, version 1, layout 128B / 64B
  constexpr uint32_t DESC_HI = ((8u * ROWB) >> 4) | (1u << 14) | ((ROWB == 128 ? 2u : 4u) << 29);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int panel_bytes = p.slab_rows * ROWB;
  const int tile_bytes = p.nt * ROWB / CG;  // bytes of one weight tile (one tap of one panel) held by THIS CTA
  const int total_tiles = p.panels * p.taps;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool cta_leader = rank == 0;
  // units of this CTA: u_n = CG * (cluster + n * n_clusters) + rank, n = 0 .. N-1.  N is the same for both CTAs of
  // a pair; a unit index >= n_units is a dummy (utterance index >= B: TMA zero-fills, nothing is stored).
  const int cluster_id = (int)blockIdx.x / CG, n_clusters = (int)gridDim.x / CG;
  const int n_super = (p.n_units + CG - 1) / CG;
  const int N = (n_super - cluster_id + n_clusters - 1) / n_clusters;
  const int u_first = CG * cluster_id + (int)rank, u_step = CG * n_clusters;

  uint8_t* slabs = smem;
  uint8_t* ring = slabs + p.panel_slots * panel_bytes;
  uint8_t* stage_buf = ring + p.n_stages * p.stage_bytes;  // [kStageBufs][kStageBufBytes] when STAGED
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_buf + (STAGED ? kStageBufs * kStageBufBytes : 0));
  uint64_t* panel_full = bars;                          // [kMaxPanelSlots]
  uint64_t* panel_empty = panel_full + kMaxPanelSlots;  // [kMaxPanelSlots]
  uint64_t* w_full = panel_empty + kMaxPanelSlots;      // [kMaxStages]
  uint64_t* w_empty = w_full + kMaxStages;              // [kMaxStages]
  uint64_t* acc_full = w_empty + kMaxStages;            // [2]
  uint64_t* acc_empty = acc_full + 2;                   // [2]
  uint64_t* stage_full = acc_empty + 2;                 // [kStageBufs] staged tile written by all epilogue warps
  uint64_t* stage_free = stage_full + kStageBufs;       // [kStageBufs] the TMA store has read the tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stage_free + kStageBufs);

  if (threadIdx.x == 0) {
    E2E_TR(0);
#ifdef E2E_TRACE
    if (blockIdx.x < 512) g_trace[blockIdx.x][8] = g_trace[blockIdx.x][9] = 0;
#endif
    tma_prefetch_desc(&tm_in);
    for (int i = 0; i < p.panel_slots; ++i) {
      mbar_init(&panel_full[i], 1);
      mbar_init(&panel_empty[i], 1);
    }
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kEpiWarps * CG);  // (the even CTA's copy collects both CTAs' epilogue warps)
    }
    for (int i = 0; i < kStageBufs; ++i) {
      mbar_init(&stage_full[i], kEpiWarps);
      mbar_init(&stage_free[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 3) {
    if (CG == 2) {
      tmem_alloc_pair(tmem_slot, 512);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before_sync();
  if (CG == 2) cluster_sync_all();  // the peer's mbarriers are initialised before anything arrives on them
  else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) E2E_TR(1);
  griddep_launch();  // the next layer's CTAs may take this SM as soon as this CTA exits

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- panel producer (TMA) ----------------
      griddep_wait();  // the activations are the previous kernel's output
      constexpr int ch_per_panel = ROWB / 2;
      const int boxes = p.slab_rows / p.box_rows;
      uint32_t slot = 0, par = 1;  // ring position; `par` is the parity a fresh/recycled slot is waited on
      UnitIter it;
      it.init(u_first, u_step, p.n_tiles, p.tiles_per_b);
      for (int n = 0; n < N; ++n, it.next()) {
        const int b = it.b;  // >= B for a dummy unit: every row is out of bounds and arrives as zeros
        const int t0 = it.tile * (128 * MT);
        for (int pn = 0; pn < p.panels; ++pn) {
          mbar_wait(&panel_empty[slot], par, 0x100 + slot);
          if (cta_leader) mbar_arrive_expect_tx(&panel_full[slot], panel_bytes * CG);
          uint8_t* dst = slabs + slot * panel_bytes;
          for (int bx = 0; bx < boxes; ++bx) {
            if (CG == 2)
              tma_load_3d_pair(dst + bx * p.box_rows * ROWB, &tm_in, pn * ch_per_panel, t0 - p.hl + bx * p.box_rows, b,
                               &panel_full[slot]);
            else
              tma_load_3d(dst + bx * p.box_rows * ROWB, &tm_in, pn * ch_per_panel, t0 - p.hl + bx * p.box_rows, b,
                          &panel_full[slot]);
          }
          if (++slot == (uint32_t)p.panel_slots) {
            slot = 0;
            par ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- weight producer (bulk copies) ----------------
      uint32_t stage = 0, par = 1;
      UnitIter it;
      it.init(u_first, u_step, p.n_tiles, p.tiles_per_b);
      for (int n = 0; n < N; ++n, it.next()) {
        const int nti = it.nti;  // (CG == 2: one N tile per layer, both CTAs of the pair share the weight tiles)
        const uint8_t* wsrc = p.w + static_cast<size_t>(nti) * total_tiles * tile_bytes;
        int first = 0;
        for (int c = 0; c < p.n_chunks; ++c, first += p.tiles_per_chunk) {
          mbar_wait(&w_empty[stage], par, 0x200 + stage);
          const int ntile = min(p.tiles_per_chunk, total_tiles - first);
          const uint32_t bytes = ntile * tile_bytes;
          if (CG == 2) {
            // this CTA's half (nt/2 rows) of every tile of the chunk; the packed image is a [rows][ROWB] matrix
            if (cta_leader) mbar_arrive_expect_tx(&w_full[stage], bytes * 2);
            for (int i = 0; i < ntile; ++i)
              tma_load_2d_pair(ring + stage * p.stage_bytes + i * tile_bytes, &tm_w, 0,
                               (nti * total_tiles + first + i) * p.nt + (int)rank * (p.nt / 2), &w_full[stage]);
          } else {
            mbar_arrive_expect_tx(&w_full[stage], bytes);
            bulk_load_1d(ring + stage * p.stage_bytes, wsrc + static_cast<size_t>(first) * tile_bytes, bytes,
                         &w_full[stage]);
          }
          if (++stage == (uint32_t)p.n_stages) {
            stage = 0;
            par ^= 1;
          }
        }
      }
    }
  } else if (warp == 2 && cta_leader) {
    // ---------------- MMA issuer (even CTA of a pair only) ----------------
    // The whole warp runs the loop (warp-uniform control flow and address arithmetic); one elected lane issues
    // the tcgen05 instructions.  No divisions: ring positions are counters that wrap.
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_bf16(128 * CG, p.nt, p.f16);
    auto wait_full = [&](uint64_t* bar, uint32_t par, uint32_t code) {
      if (CG == 2) mbar_wait_cluster(bar, par, code);
      else mbar_wait(bar, par, code);
    };
    auto commit = [&](uint64_t* bar) {
      if (CG == 2) umma_commit_pair(bar);
      else umma_commit(bar);
    };
    const uint32_t slab_lo = ((smem_u32(slabs) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t ring_lo = ((smem_u32(ring) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t panel16 = panel_bytes >> 4, stage16 = p.stage_bytes >> 4, tile16 = tile_bytes >> 4;
    uint32_t slot = 0, ppar = 0, stage = 0, wpar = 0, acc = 0, apar = 1;
    bool first_unit = true;
    UnitIter uit;
    uit.init(u_first, u_step, p.n_tiles, p.tiles_per_b);
    for (int n = 0; n < N; ++n, uit.next()) {
      const int nti = uit.nti;
      wait_full(&acc_empty[acc], apar, 0x300 + acc);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * (MT * p.nt);
      if (first_unit && leader) E2E_TR(2);
      int tap = 0, left = total_tiles;
      uint32_t accum = 0;
      for (int c = 0; c < p.n_chunks; ++c) {
        wait_full(&w_full[stage], wpar, 0x400 + stage);
        tc_fence_after_sync();
        const int ntile = min(p.tiles_per_chunk, left);
        left -= ntile;
        uint32_t b_lo = ring_lo + stage * stage16;
        for (int i = 0; i < ntile; ++i, b_lo += tile16) {
          if (tap == 0) {
            wait_full(&panel_full[slot], ppar, 0x500 + slot);
            tc_fence_after_sync();
          }
          const uint32_t a_lo = slab_lo + slot * panel16 + (p.hl + p.shift[nti][tap]) * ROW16;
          if (leader) {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
#pragma unroll
              for (int ks = 0; ks < KS; ++ks) {
                const uint64_t da = (static_cast<uint64_t>(DESC_HI) << 32) | (a_lo + m * (128 * ROW16) + ks * 2);
                const uint64_t db = (static_cast<uint64_t>(DESC_HI) << 32) | (b_lo + ks * 2);
                if (CG == 2) {
                  if (ks == 0)
                    umma_bf16_pair(d_tmem + m * p.nt, da, db, idesc, accum);
                  else
                    umma_bf16_acc_pair(d_tmem + m * p.nt, da, db, idesc);
                } else {
                  if (ks == 0)
                    umma_bf16(d_tmem + m * p.nt, da, db, idesc, accum);
                  else
                    umma_bf16_acc(d_tmem + m * p.nt, da, db, idesc);
                }
              }
            }
          }
          accum = 1;
          if (++tap == p.taps) {
            if (leader) commit(&panel_empty[slot]);  // slab panel consumed: the producers may refill it
            tap = 0;
            if (++slot == (uint32_t)p.panel_slots) {
              slot = 0;
              ppar ^= 1;
            }
          }
        }
        if (leader) commit(&w_empty[stage]);
        if (++stage == (uint32_t)p.n_stages) {
          stage = 0;
          wpar ^= 1;
        }
      }
      if (leader) commit(&acc_full[acc]);
      if (first_unit && leader) E2E_TR(3);
      first_unit = false;
      if (++acc == (uint32_t)p.n_acc) {
        acc = 0;
        apar ^= 1;
      }
    }
    if (leader) E2E_TR(4);
  } else if (warp == 3) {
    if (STAGED && lane == 0) {
      // ---------------- store warp (STAGED): staged 128-row x 64-column tiles -> global memory by TMA ----------
      // The epilogue warps leave each group of four 16-column items (one 64-column panel of one M tile) in a
      // SWIZZLE_128B staging buffer; this thread writes it with one bulk-tensor store (rows past the utterance end
      // are clipped by the TMA unit).  A thread's own global stores would be 32 bytes per lane at a row pitch of
      // 2 * n_total bytes - ~64 L1 wavefronts per warp access on the data pipe the tensor core's operand reads need.
      griddep_wait();
      tma_prefetch_desc(&tm_out);
      const int groups = MT * (p.nt >> 6), gpm = p.nt >> 6;
      uint32_t gi = 0;
      UnitIter uit;
      uit.init(u_first, u_step, p.n_tiles, p.tiles_per_b);
      for (int n = 0; n < N; ++n, uit.next()) {
        const int t0 = uit.tile * (128 * MT);
        for (int g = 0; g < groups; ++g, ++gi) {
          const uint32_t buf = gi % kStageBufs, par = (gi / kStageBufs) & 1;
          mbar_wait(&stage_full[buf], par, 0x800 + buf);
          const int m = g / gpm, pn = g - m * gpm;
          if (uit.b < p.B && t0 + m * 128 < p.T) {
            tma_store_3d(&tm_out, stage_buf + buf * kStageBufBytes, uit.nti * p.nt + pn * 64, t0 + m * 128, uit.b);
            bulk_commit_group();
            bulk_wait_group_read<0>();
          }
          mbar_arrive(&stage_free[buf]);
        }
      }
      bulk_wait_group<0>();
    }
  } else if (warp >= 4) {
    // ---------------- epilogue: TMEM -> registers -> global ----------------
    griddep_wait();  // residual / running-sum reads and every output store follow the previous kernel
    const int e = warp - 4;
    const int quarter = e & 3;  // == warp % 4: the TMEM lanes this warp may read
    const int part = e >> 2;    // this warp takes the items with item % 4 == part
    const int nchunk = p.nt >> 4;
    const int items = MT * nchunk;  // (m tile, 16-column chunk) pairs
    // item -> (m tile, chunk) without an integer division when nchunk is a power of two (nt = 32 / 64 / 128 / 256)
    const bool nc_pow2 = (nchunk & (nchunk - 1)) == 0;
    const int nc_shift = 31 - __clz(nchunk);
    auto m_of = [&](int item) -> int { return nc_pow2 ? (item >> nc_shift) : item / nchunk; };
    const int row_in_tile = quarter * 32 + lane;
    EpiOut eo;
    eo.sum_a = p.sum_a;
    eo.out_f32 = p.out_f32;
    eo.out_act = p.out_act;
    eo.slope = p.slope;
    eo.scale = p.divisor != 0.f ? 1.0f / p.divisor : 0.f;
    eo.inv = p.res_inv_slope;
    eo.act_tanh = p.act_tanh;
    eo.f16 = p.f16;
    uint32_t it = 0, acc = 0, apar = 0;
    uint32_t gi = 0;  // STAGED: running count of 64-column groups (the staging ring position)
    UnitIter uit;
    uit.init(u_first, u_step, p.n_tiles, p.tiles_per_b);
    for (int n = 0; n < N; ++n, ++it, uit.next()) {
      const int nti = uit.nti;
      const int b = uit.b;
      const int t0 = uit.tile * (128 * MT);
      const uint32_t d_tmem = tmem_base + acc * (MT * p.nt) + (static_cast<uint32_t>(quarter * 32) << 16);

      // residual and partial sums (16 bf16 each) are fetched one item ahead, before the accumulator is waited on
      uint4 rqa[2], rqb[2], saa[2], sab[2];
      auto item_off = [&](int item, int& n0, bool& valid) -> size_t {
        const int m = m_of(item), cc = item - m * nchunk;
        const int t = t0 + m * 128 + row_in_tile;
        valid = item < items && t < p.T && b < p.B;
        n0 = nti * p.nt + cc * 16;
        return (static_cast<size_t>(b) * p.T + (valid ? t : 0)) * p.n_total + n0;
      };
      const int t8 = (p.T + 7) >> 3, c16 = p.n_total >> 4;
      auto tiled_off = [&](int item, int n0) -> size_t {
        const int m = m_of(item);
        return tiled8_off(b, t0 + m * 128 + row_in_tile, n0 >> 4, t8, c16);
      };
      auto prefetch = [&](int item, uint4 (&rq)[2], uint4 (&sa)[2]) {
        int n0;
        bool valid;
        const size_t off = item_off(item, n0, valid);
        rq[0] = rq[1] = make_uint4(0u, 0u, 0u, 0u);
        if (valid) {
          if (p.res_act) ld_global_256(p.res_act + off, rq[0], rq[1]);  // plain loads: may be updated in place
          if (p.sum_a) ld_global_256(p.sum_a + (p.sum_tiled ? tiled_off(item, n0) : off), sa[0], sa[1]);
        }
      };
      prefetch(part, rqa, saa);  // overlaps the MMAs of this unit
      prefetch(part + 4, rqb, sab);
      mbar_wait(&acc_full[acc], apar, 0x600 + acc);
      tc_fence_after_sync();
      if (it == 0 && threadIdx.x == 128) E2E_TR(5);

      auto process = [&](int item, uint4 (&rq)[2], uint4 (&sa)[2]) {
        const int m = m_of(item), cc = item - m * nchunk;
        int n0;
        bool valid;
        const size_t off = item_off(item, n0, valid);
        uint32_t v[16];
        tmem_ld_32x16(d_tmem + m * p.nt + cc * 16, v);
        float4 bv[4];  // the bias loads are in flight together with the TMEM load
        if (p.bias_const) {
          const int nb = n0 & p.bias_mask;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            bv[i] = make_float4(p.cbias[nb + 4 * i], p.cbias[nb + 4 * i + 1], p.cbias[nb + 4 * i + 2], p.cbias[nb + 4 * i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) bv[i] = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + i);
        }
        tmem_ld_wait();
        if (STAGED) {
          // item = 4 * group + part: this warp's 16 columns of the group's 64-column tile, in the tile's swizzled layout
          uint32_t pk[8];
          epi_compute16(v, bv, rq, sa, eo, pk);
          const uint32_t buf = gi % kStageBufs, spar = (gi / kStageBufs) & 1;
          ++gi;
          uint32_t so = static_cast<uint32_t>(row_in_tile) * 128 + (cc & 3) * 32;
          so ^= ((so >> 7) & 7u) << 4;
          so += smem_u32(stage_buf) + buf * kStageBufBytes;
          mbar_wait(&stage_free[buf], spar ^ 1, 0x880 + buf);  // the previous tile in this buffer has been read
          st_shared_u4(so, make_uint4(pk[0], pk[1], pk[2], pk[3]));
          st_shared_u4(so ^ 16u, make_uint4(pk[4], pk[5], pk[6], pk[7]));
          fence_proxy_async_smem();  // the TMA unit reads the tile through the async proxy
          __syncwarp();
          if (lane == 0) mbar_arrive(&stage_full[buf]);
        } else {
          epi_finish16(v, bv, rq, sa, eo, (p.out_tiled && valid) ? tiled_off(item, n0) : off, valid);
        }
        prefetch(item + 8, rq, sa);  // refill this slot for the item after next
      };
      for (int item = part; item < items; item += 8) {
        process(item, rqa, saa);
        if (item + 4 < items) process(item + 4, rqb, sab);
      }
      // all of this warp's TMEM reads of the unit are complete (tcgen05.wait::ld above): release the accumulator
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_remote(&acc_empty[acc], 0);  // the issuing thread waits in the even CTA
        else mbar_arrive(&acc_empty[acc]);
      }
      if (it == 0 && threadIdx.x == 128) E2E_TR(6);
      if (++acc == (uint32_t)p.n_acc) {
        acc = 0;
        apar ^= 1;
      }
    }
  }

  tc_fence_before_sync();
  if (CG == 2) cluster_sync_all();  // neither CTA may exit (or free TMEM) while the pair's MMAs can still touch it
  else __syncthreads();
  if (warp == 3) {
    __syncwarp();
    if (CG == 2) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
  if (threadIdx.x == 0) E2E_TR(7);
}

}  // namespace e2e
