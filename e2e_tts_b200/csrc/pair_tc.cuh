// Fused residual pair of ResBlock1 (reference e2e_tts/models/vocoder/layers.py:34-39):
//
//     xt = c1(leaky_relu(x))          dilated conv, k taps, dilation d
//     xt = c2(leaky_relu(xt))         conv, k taps, dilation 1
//     x  = xt + x
//
// in ONE persistent tcgen05 kernel.  leaky_relu(x) arrives as a bf16 channels-last slab by TMA (rows outside the
// utterance zero-filled = c1's zero padding); c1's accumulators are read from TMEM, biased, activated, rounded to
// bf16 and written into a swizzled shared-memory slab `M` that is directly c2's UMMA A operand (rows outside
// [0,T) forced to zero = c2's zero padding, SURVEY.md §8 a'1) — the intermediate never touches HBM.  c2's epilogue
// adds the residual (recovered from the stored bf16 leaky_relu(x) by the inverse LeakyReLU), the running resblock
// sum and the /3, and writes the next activation.
//
// A unit is (utterance, 128*MT - (k-1) output rows): c1 computes 128*MT rows (the extra k-1 rows are c2's halo).
// Each CTA keeps TWO units in flight on two "lanes" (own input slab, own M slab, own TMEM accumulators) and the
// single MMA-issuing warp walks the software-pipelined job order
//     c1(u0) | c1(u1) c2(u0) | c1(u2) c2(u1) | ... | c2(u_last)
// so the epilogue of every job overlaps the MMAs of the next one.  Roles: warp 0 slab producer (TMA), warp 1
// weight producer (bulk copies, both convs, job order), warp 2 MMA issuer, warp 3 TMEM allocator (and, in
// STAGED kernels, the store warp), warps 4-19 epilogue (four per TMEM lane quarter, 16-column items, epilogue.cuh).
//
// CG = 2 runs the same pipeline on a CTA PAIR (2-CTA cluster, tcgen05 cta_group::2): each CTA owns its units, its
// slabs and its epilogue, but one thread of the even CTA issues M = 256 MMAs whose rows 0-127 / 128-255 are the two
// CTAs' units and whose B operand (the weight tile) is split between them, N/2 rows each.  Per CTA that halves the
// weight bytes streamed from L2 into shared memory and the B-operand bytes every MMA reads back out of it - the
// two things that cap the N = 128 MMAs of the C = 128 stage below the tensor-pipe rate in the one-CTA form.
// All "full" barriers the issuing thread waits on live in the even CTA (TMA .cta_group::2 loads of both CTAs
// count their bytes there, the odd CTA's epilogue warps arrive there remotely); every tcgen05.commit is
// multicast to both CTAs.
#pragma once
#include "conv_tc.cuh"

namespace e2e {

struct PairParams {
  int T, B;
  int panels;           // K panels of the C channels (C/64, or 1 for C = 32)
  int nt;               // C (output columns of both convs)
  int taps;             // k
  int dil;              // dilation of c1
  int a_rows;           // input slab rows per panel (multiple of box_rows, >= 128*MT + (k-1)*dil)
  int box_rows;
  int m_rows;           // M slab rows per panel (>= 128*MT + k-1, multiple of 8)
  int r_out;            // valid output rows per unit = 128*MT - (k-1)
  int tiles_per_chunk, n_chunks, n_stages, stage_bytes;   // weight ring geometry (same for c1 and c2)
  int tiles_per_b, n_units;
  float slope_mid;      // LeakyReLU between c1 and c2 (0.1)
  float slope;          // LeakyReLU applied to out_act
  float divisor;        // 0 = none
  float res_inv_slope;  // 1 / slope of the stored input activation
  const uint8_t* w1;    // packed weights of c1 / c2: [panel][tap][nt][rowb] swizzled images
  const uint8_t* w2;
  float bias1[128];     // biases of c1 / c2 (nt <= 128), in the kernel-parameter constant bank: every epilogue warp reads
  float bias2[128];     // its 16 columns with warp-uniform constant loads, off the shared-memory / L1 data pipe
  const __nv_bfloat16* res_act;  // == the kernel's input tensor (bf16 leaky_relu(x)), read for the residual
  const __nv_bfloat16* sum_a;    // bf16 running sum over the stage's resblocks (generator.py:44-47) or nullptr
  int sum_tiled;                 // sum_a is in the tiled8 layout (epilogue.cuh)
  int no_sum_prefetch;           // experiments (E2E_NO_SUM_PREFETCH=1): no L2 prefetch of the running sum by the producer
  int out_tiled;                 // out_act is written in the tiled8 layout (direct stores; never with STAGED)
  int f16;                       // 16-bit tensors and operands are fp16 instead of bf16 (ptx.cuh pack16)
  __nv_bfloat16* out_act;        // bf16 leaky_relu(result, slope), written through tm_out / tm_out2 (TMA stores)
};

template <int ROWB, int MT, int CG, bool STAGED>
__global__ void __launch_bounds__(kConvThreads, 1)
pair_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w1,
               const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_out,
               const __grid_constant__ CUtensorMap tm_out2, const __grid_constant__ PairParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  constexpr int KS = ROWB / 32;
  constexpr uint32_t ROW16 = ROWB >> 4;
  constexpr uint32_t DESC_HI = ((8u * ROWB) >> 4) | (1u << 14) | ((ROWB == 128 ? 2u : 4u) << 29);
  constexpr uint32_t SWZ = ROWB == 128 ? 7u : 3u;
  constexpr int CH_PANEL = ROWB / 2;
  // STAGED: the c2 epilogue stages its result in shared memory and a TMA store writes it out (pair_host.cuh picks
  // it for the epilogue-bound pairs: C >= 64 and k = 3; measured per launch in the full forward, direct -> staged:
  // C = 128 k = 3 125 -> 111 us, C = 64 k = 3 106 -> 100 us, but C = 128 k = 11 244 -> 259 us, C = 64 k = 7 126 -> 131 us,
  // C = 32 k = 11 136 -> 148 us: where the MMAs already saturate shared memory the extra staging traffic loses).
  constexpr bool kStaged = STAGED;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int a_panel_bytes = p.a_rows * ROWB;
  const int m_panel_bytes = p.m_rows * ROWB;
  const int a_lane_bytes = p.panels * a_panel_bytes;
  const int m_lane_bytes = p.panels * m_panel_bytes;
  const int tile_bytes = p.nt * ROWB / CG;  // bytes of one weight tile (one tap of one panel) held by THIS CTA
  const int total_tiles = p.panels * p.taps;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool cta_leader = rank == 0;
  const int h1 = (p.taps - 1) / 2 * p.dil, h2 = (p.taps - 1) / 2;
  const int acc_cols = MT * p.nt;  // TMEM columns per accumulator; 4 accumulators: [lane][conv]

  uint8_t* a_slab = smem;                                   // [2 lanes][panels][a_rows][ROWB]
  uint8_t* m_slab = a_slab + 2 * a_lane_bytes;              // [2 lanes][panels][m_rows][ROWB]
  uint8_t* ring = m_slab + 2 * m_lane_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + p.n_stages * p.stage_bytes);
  uint64_t* a_full = bars;             // [2][4]
  uint64_t* a_empty = a_full + 8;      // [2][4]
  uint64_t* w_full = a_empty + 8;      // [kMaxStages]
  uint64_t* w_empty = w_full + kMaxStages;
  uint64_t* acc_full = w_empty + kMaxStages;  // [2][2]
  uint64_t* acc_empty = acc_full + 4;         // [2][2]
  uint64_t* m_full = acc_empty + 4;           // [2]
  uint64_t* stage_full = m_full + 2;          // [2] the c2 epilogue has staged a unit's result in the lane's M slab
  uint64_t* stage_free = stage_full + 2;      // [2] ... and the store warp's TMA store has read it out again
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stage_free + 2);

  // units of this CTA: u_n = CG * (cluster + n * n_clusters) + rank, n = 0 .. N-1.  N is the same for both CTAs of
  // a pair; a unit index >= n_units is a dummy (utterance index B: TMA zero-fills, nothing is stored).
  const int cluster_id = (int)blockIdx.x / CG, n_clusters = (int)gridDim.x / CG;
  const int n_super = (p.n_units + CG - 1) / CG;
  const int N = (n_super - cluster_id + n_clusters - 1) / n_clusters;
  const int u_first = CG * cluster_id + (int)rank, u_step = CG * n_clusters;

  if (threadIdx.x == 0) {
    E2E_TR(0);
#ifdef E2E_TRACE
    if (blockIdx.x < 512)
      for (int i = 2; i < 12; ++i)
        if (i != 4 && i != 7) g_trace[blockIdx.x][i] = 0;
#endif
    tma_prefetch_desc(&tm_in);
    for (int i = 0; i < 8; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kEpiWarps * CG);  // (the even CTA's copy collects both CTAs' epilogue warps)
    }
    mbar_init(&m_full[0], kEpiWarps * CG);
    mbar_init(&m_full[1], kEpiWarps * CG);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&stage_full[i], kEpiWarps);   // local: every CTA stores its own units
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
      // ---------------- input slab producer (TMA) ----------------
      griddep_wait();  // the activations are the previous kernel's output
      const int boxes = p.a_rows / p.box_rows;
      UnitIter uit;
      uit.init(u_first, u_step, 1, p.tiles_per_b);
      for (int n = 0; n < N; ++n, uit.next()) {
        const int b = uit.b;  // == B for a dummy unit: every row is out of bounds and arrives as zeros
        const int t0 = uit.tile * p.r_out;
        const int ln = n & 1;
        const uint32_t par = ((n >> 1) & 1) ^ 1;
        for (int pn = 0; pn < p.panels; ++pn) {
          mbar_wait(&a_empty[ln * 4 + pn], par, 0x100 + ln * 4 + pn);
          if (cta_leader) mbar_arrive_expect_tx(&a_full[ln * 4 + pn], a_panel_bytes * CG);
          uint8_t* dst = a_slab + ln * a_lane_bytes + pn * a_panel_bytes;
          for (int bx = 0; bx < boxes; ++bx) {
            if (CG == 2)
              tma_load_3d_pair(dst + bx * p.box_rows * ROWB, &tm_in, pn * CH_PANEL, t0 - h2 - h1 + bx * p.box_rows, b,
                               &a_full[ln * 4 + pn]);
            else
              tma_load_3d(dst + bx * p.box_rows * ROWB, &tm_in, pn * CH_PANEL, t0 - h2 - h1 + bx * p.box_rows, b,
                          &a_full[ln * 4 + pn]);
          }
        }
        if (p.sum_a && !p.no_sum_prefetch && b < p.B)   // see prefetch_sum_rows (epilogue.cuh)
          prefetch_sum_rows(p.sum_a, p.sum_tiled, b, t0, t0 + p.r_out, p.T, p.nt);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- weight producer: job order c1(u0) | c1(u1) c2(u0) | ... ----------------
      uint32_t stage = 0, par = 1;
      auto stream = [&](const uint8_t* wsrc, const CUtensorMap* tmw) {
        int first = 0;
        for (int c = 0; c < p.n_chunks; ++c, first += p.tiles_per_chunk) {
          mbar_wait(&w_empty[stage], par, 0x200 + stage);
          const int ntile = min(p.tiles_per_chunk, total_tiles - first);
          const uint32_t bytes = ntile * tile_bytes;
          if (CG == 2) {
            // this CTA's half (N/2 rows) of every tile of the chunk; the packed image is a [rows][ROWB] matrix
            if (cta_leader) mbar_arrive_expect_tx(&w_full[stage], bytes * 2);
            for (int i = 0; i < ntile; ++i)
              tma_load_2d_pair(ring + stage * p.stage_bytes + i * tile_bytes, tmw, 0,
                               (first + i) * p.nt + (int)rank * (p.nt / 2), &w_full[stage]);
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
      };
      for (int n = 0; n <= N; ++n) {
        if (n < N) stream(p.w1, &tm_w1);
        if (n >= 1) stream(p.w2, &tm_w2);
      }
    }
  } else if (warp == 2 && cta_leader) {
    // ---------------- MMA issuer (warp-uniform loop, one elected lane issues; even CTA only) ----------------
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
    const uint32_t a_lo0 = ((smem_u32(a_slab) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t m_lo0 = ((smem_u32(m_slab) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t ring_lo = ((smem_u32(ring) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t stage16 = p.stage_bytes >> 4, tile16 = tile_bytes >> 4;
    uint32_t stage = 0, wpar = 0;
    // one job = one convolution of one unit
    auto job = [&](int n, int ci) {
      const int ln = n & 1;
      const uint32_t par = (n >> 1) & 1;
      const uint32_t src_lo = ci == 0 ? a_lo0 + ln * (a_lane_bytes >> 4) : m_lo0 + ln * (m_lane_bytes >> 4);
      const uint32_t panel16 = (ci == 0 ? a_panel_bytes : m_panel_bytes) >> 4;
      const uint32_t tap_rows16 = (ci == 0 ? p.dil : 1) * ROW16;  // tap j reads rows shifted by j*d (c1) / j (c2)
#ifdef E2E_TRACE
      unsigned long long tq = gtime_ns();
#endif
      wait_full(&acc_empty[ln * 2 + ci], par ^ 1, 0x300 + ln * 2 + ci);
#ifdef E2E_TRACE
      if (leader && blockIdx.x < 512) { const unsigned long long t2 = gtime_ns(); g_trace[blockIdx.x][8] += t2 - tq; tq = t2; }
#endif
      if (ci == 1) wait_full(&m_full[ln], par, 0x380 + ln);  // c1's epilogue has written the whole M slab
#ifdef E2E_TRACE
      if (leader && blockIdx.x < 512) g_trace[blockIdx.x][9] += gtime_ns() - tq;
#endif
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + (ln * 2 + ci) * acc_cols;
      int tap = 0, pn = 0, left = total_tiles;
      uint32_t accum = 0;
      for (int c = 0; c < p.n_chunks; ++c) {
#ifdef E2E_TRACE
        const unsigned long long tw = gtime_ns();
#endif
        wait_full(&w_full[stage], wpar, 0x400 + stage);
        tc_fence_after_sync();
#ifdef E2E_TRACE
        if (leader && blockIdx.x < 512) g_trace[blockIdx.x][10] += gtime_ns() - tw;
#endif
        const int ntile = min(p.tiles_per_chunk, left);
        left -= ntile;
        uint32_t b_lo = ring_lo + stage * stage16;
        for (int i = 0; i < ntile; ++i, b_lo += tile16) {
          if (tap == 0 && ci == 0) {
#ifdef E2E_TRACE
            const unsigned long long ta = gtime_ns();
#endif
            wait_full(&a_full[ln * 4 + pn], par, 0x500 + ln * 4 + pn);
            tc_fence_after_sync();
#ifdef E2E_TRACE
            if (leader && blockIdx.x < 512) g_trace[blockIdx.x][11] += gtime_ns() - ta;
#endif
          }
          const uint32_t s_lo = src_lo + pn * panel16 + tap * tap_rows16;
          if (leader) {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
#pragma unroll
              for (int ks = 0; ks < KS; ++ks) {
                const uint64_t da = (static_cast<uint64_t>(DESC_HI) << 32) | (s_lo + m * (128 * ROW16) + ks * 2);
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
            if (ci == 0 && leader) commit(&a_empty[ln * 4 + pn]);  // input panel consumed -> TMA may refill
            tap = 0;
            ++pn;
          }
        }
        if (leader) commit(&w_empty[stage]);
        if (++stage == (uint32_t)p.n_stages) {
          stage = 0;
          wpar ^= 1;
        }
      }
      if (leader) commit(&acc_full[ln * 2 + ci]);
    };
    for (int n = 0; n <= N; ++n) {
      if (n < N) job(n, 0);
      if (n >= 1) job(n - 1, 1);
    }
    if (leader) E2E_TR(4);
  } else if (warp == 3) {
    if (kStaged && lane == 0) {
      // ---------------- store warp: the staged result tiles go to global memory as TMA bulk-tensor stores -------
      // The c2 epilogue leaves a unit's bf16 result in the lane's M slab (dead between c2's last MMA and the next
      // c1 epilogue of the lane), in the slab's own swizzled panel layout, which is what a SWIZZLE_128B / _64B
      // tensor map reads.  One store per 64-channel panel moves the unit's r_out rows (rows past the utterance end
      // are clipped by the TMA unit).  No LSU wavefronts, no 32-byte-per-lane global stores.
      griddep_wait();
      tma_prefetch_desc(&tm_out);
      UnitIter uit;
      uit.init(u_first, u_step, 1, p.tiles_per_b);
      const int rows1 = p.r_out < 256 ? p.r_out : 256;
      for (int n = 0; n < N; ++n, uit.next()) {
        const int ln = n & 1;
        const uint32_t par = (n >> 1) & 1;
        mbar_wait(&stage_full[ln], par, 0x800 + ln);
        if (uit.b < p.B) {
          const int t0 = uit.tile * p.r_out;
          const uint8_t* src = m_slab + ln * m_lane_bytes;
          for (int pn = 0; pn < p.panels; ++pn) {
            tma_store_3d(&tm_out, src + pn * m_panel_bytes, pn * CH_PANEL, t0, uit.b);
            if (p.r_out > 256)
              tma_store_3d(&tm_out2, src + pn * m_panel_bytes + rows1 * ROWB, pn * CH_PANEL, t0 + rows1, uit.b);
          }
          bulk_commit_group();
          bulk_wait_group_read<0>();
        }
        mbar_arrive(&stage_free[ln]);
      }
      bulk_wait_group<0>();
    }
  } else if (warp >= 4) {
    // ---------------- epilogue ----------------
    griddep_wait();  // residual / running-sum reads and every output store follow the previous kernel
    const int e = warp - 4;
    const int quarter = e & 3;
    const int part = e >> 2;
    const int nchunk = p.nt >> 4;
    // MT * nchunk == 8 items per job for every supported (C, MT): this warp owns items `part` and `part + 4`,
    // i.e. fixed (m tile, 16-column chunk) pairs for the whole kernel
    const int mA = part / nchunk, ccA = part - mA * nchunk;
    const int mB = (part + 4) / nchunk, ccB = (part + 4) - mB * nchunk;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    EpiOut eo;
    eo.sum_a = p.sum_a;
    eo.out_f32 = nullptr;  // (this kernel writes bf16 activations only)
    eo.out_act = p.out_act;
    eo.slope = p.slope;
    eo.scale = p.divisor != 0.f ? 1.0f / p.divisor : 0.f;
    eo.inv = p.res_inv_slope;
    eo.act_tanh = 0;
    eo.f16 = p.f16;
    const float smid = p.slope_mid;
    E2E_TR2_DECL

    // one 16-column item of c1: acc + bias1 -> leaky_relu -> bf16 -> M slab (zero outside the utterance)
    auto store_mid_t = [&](auto f16tag, const uint32_t (&v)[16], const float4 (&bv)[4], int m, int cc, int t0,
                           uint8_t* mdst) {
      constexpr bool F16 = decltype(f16tag)::value;
      const int r = m * 128 + row_in_tile;   // M slab row
      const int t = t0 - h2 + r;             // global time step of this row
      const bool inside = t >= 0 && t < p.T;
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float f0 = __uint_as_float(v[4 * i]) + bv[i].x, f1 = __uint_as_float(v[4 * i + 1]) + bv[i].y;
        const float f2 = __uint_as_float(v[4 * i + 2]) + bv[i].z, f3 = __uint_as_float(v[4 * i + 3]) + bv[i].w;
        const uint32_t h0 = pack16t<F16>(fmaxf(f0, f0 * smid), fmaxf(f1, f1 * smid));
        const uint32_t h1v = pack16t<F16>(fmaxf(f2, f2 * smid), fmaxf(f3, f3 * smid));
        pk[2 * i] = inside ? h0 : 0u;
        pk[2 * i + 1] = inside ? h1v : 0u;
      }
      // 16 channels = two 16-byte chunks of this row in panel (n0 / CH_PANEL)
      const int n0 = cc * 16;
      const int pn = n0 / CH_PANEL;
      const int chunk0 = (n0 % CH_PANEL) / 8;
      uint8_t* prow = mdst + pn * m_panel_bytes;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint32_t off = static_cast<uint32_t>(r) * ROWB + (chunk0 + q) * 16;
        off ^= ((off >> 7) & SWZ) << 4;
        *reinterpret_cast<uint4*>(prow + off) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      }
    };
    auto store_mid = [&](const uint32_t (&v)[16], const float4 (&bv)[4], int m, int cc, int t0, uint8_t* mdst) {
      if (p.f16) store_mid_t(std::true_type{}, v, bv, m, cc, t0, mdst);
      else store_mid_t(std::false_type{}, v, bv, m, cc, t0, mdst);
    };
    auto param_bias = [&](const float* sb, int cc, float4 (&bv)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        bv[i] = make_float4(sb[cc * 16 + 4 * i], sb[cc * 16 + 4 * i + 1], sb[cc * 16 + 4 * i + 2], sb[cc * 16 + 4 * i + 3]);
    };

    // c1 epilogue.  The TMEM load of the second item is in flight while the first item is processed; the bias
    // comes from the kernel parameters (constant bank) - no shared-memory or L1 access.
    auto epi1 = [&](int n, int t0) {
      const int ln = n & 1;
      const uint32_t par = (n >> 1) & 1;
      uint8_t* mdst = m_slab + ln * m_lane_bytes;
#ifdef E2E_TRACE
      const unsigned long long te0 = gtime_ns();
#endif
      mbar_wait(&acc_full[ln * 2 + 0], par, 0x600 + ln * 2);
      if (kStaged) mbar_wait(&stage_free[ln], par ^ 1, 0x680 + ln);   // the result staged in this slab has been stored
      tc_fence_after_sync();
      E2E_TR2(2);
#ifdef E2E_TRACE
      const unsigned long long te1 = gtime_ns();
#endif
      const uint32_t d_tmem = tmem_base + (ln * 2 + 0) * acc_cols + lane_sel;
      uint32_t vA[16], vB[16];
      float4 bv[4];
      tmem_ld_32x16(d_tmem + mA * p.nt + ccA * 16, vA);
      param_bias(p.bias1, ccA, bv);
      tmem_ld_wait();
      E2E_TR2(3);
      tmem_ld_32x16(d_tmem + mB * p.nt + ccB * 16, vB);
      store_mid(vA, bv, mA, ccA, t0, mdst);
      E2E_TR2(4);
      param_bias(p.bias1, ccB, bv);
      tmem_ld_wait();
      E2E_TR2(5);
      store_mid(vB, bv, mB, ccB, t0, mdst);
      E2E_TR2(6);
      tc_fence_before_sync();
      fence_proxy_async_smem();  // the M slab is read by the tensor core through the async proxy
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) {  // the issuing thread waits in the even CTA
          mbar_arrive_remote(&acc_empty[ln * 2 + 0], 0);
          mbar_arrive_remote(&m_full[ln], 0);
        } else {
          mbar_arrive(&acc_empty[ln * 2 + 0]);
          mbar_arrive(&m_full[ln]);
        }
      }
      E2E_TR2(7);
#ifdef E2E_TRACE
      if (threadIdx.x == 128 && blockIdx.x < 512) {
        g_trace[blockIdx.x][2] += te1 - te0;           // epi1: waiting for the accumulator
        g_trace[blockIdx.x][3] += gtime_ns() - te1;    // epi1: work
      }
#endif
    };

    // Residual and running sum of job c2(n): 16 bf16 of this thread's row per item, both items of the warp fetched
    // one whole job early (before the c1 epilogue that precedes this c2 epilogue): the L2/HBM latency is off the
    // critical path.
    uint4 rqa[2], rqb[2], sqa[2], sqb[2];
    size_t offa = 0, offb = 0;
    bool va = false, vb = false;
    auto prefetch_res = [&](int b, int t0) {
      const int oa = mA * 128 + row_in_tile, ob = mB * 128 + row_in_tile;
      va = oa < p.r_out && t0 + oa < p.T && b < p.B;
      vb = ob < p.r_out && t0 + ob < p.T && b < p.B;
      offa = (static_cast<size_t>(b) * p.T + (va ? t0 + oa : 0)) * p.nt + ccA * 16;
      offb = (static_cast<size_t>(b) * p.T + (vb ? t0 + ob : 0)) * p.nt + ccB * 16;
      rqa[0] = rqa[1] = rqb[0] = rqb[1] = make_uint4(0u, 0u, 0u, 0u);
      const int t8 = (p.T + 7) >> 3, c16 = p.nt >> 4;
      const size_t ta = (p.sum_tiled | p.out_tiled) && va ? tiled8_off(b, t0 + oa, ccA, t8, c16) : 0;
      const size_t tb = (p.sum_tiled | p.out_tiled) && vb ? tiled8_off(b, t0 + ob, ccB, t8, c16) : 0;
      if (va && p.res_act) {   // (res_act == nullptr: timing experiments of tests/cuda only - the residual is always present)
        ld_global_256(p.res_act + offa, rqa[0], rqa[1]);
        if (p.sum_a) ld_global_256(p.sum_a + (p.sum_tiled ? ta : offa), sqa[0], sqa[1]);
      }
      if (vb && p.res_act) {
        ld_global_256(p.res_act + offb, rqb[0], rqb[1]);
        if (p.sum_a) ld_global_256(p.sum_a + (p.sum_tiled ? tb : offb), sqb[0], sqb[1]);
      }
      if (p.out_tiled) {   // from here on offa / offb address the output
        offa = ta;
        offb = tb;
      }
    };

    // c2 epilogue: acc + bias2 + residual (+ running sum, * 1/divisor) -> leaky_relu -> bf16, staged in the lane's M
    // slab at this thread's (row, columns) position in the slab's swizzled layout; the store warp moves the tile to
    // global memory with a TMA store.  (A thread's own global stores would be 32 bytes per lane at a row pitch of
    // 2*C bytes: ~64 L1 wavefronts per warp access, on the data pipe the tensor core's operand reads need.)
    auto own_off = [&](int m, int cc) -> uint32_t {
      const int n0 = cc * 16;
      uint32_t off = static_cast<uint32_t>(m * 128 + row_in_tile) * ROWB + ((n0 % CH_PANEL) / 8) * 16;
      off ^= ((off >> 7) & SWZ) << 4;
      return off + (n0 / CH_PANEL) * m_panel_bytes;
    };
    const uint32_t soA = smem_u32(m_slab) + own_off(mA, ccA), soB = smem_u32(m_slab) + own_off(mB, ccB);
    auto epi2 = [&](int n) {
      const int ln = n & 1;
      const uint32_t par = (n >> 1) & 1;
      const uint32_t lane_off = ln * m_lane_bytes;
#ifdef E2E_TRACE
      const unsigned long long tf0 = gtime_ns();
#endif
      mbar_wait(&acc_full[ln * 2 + 1], par, 0x700 + ln * 2);   // c2's MMAs are done: accumulator ready, M slab dead
      tc_fence_after_sync();
      E2E_TR2(8);
#ifdef E2E_TRACE
      const unsigned long long tf1 = gtime_ns();
#endif
      const uint32_t d_tmem = tmem_base + (ln * 2 + 1) * acc_cols + lane_sel;
      uint32_t vA[16], vB[16], pk[8];
      float4 bv[4];
      tmem_ld_32x16(d_tmem + mA * p.nt + ccA * 16, vA);
      param_bias(p.bias2, ccA, bv);
      tmem_ld_wait();
      E2E_TR2(9);
      tmem_ld_32x16(d_tmem + mB * p.nt + ccB * 16, vB);
      if (kStaged) {
        epi_compute16(vA, bv, rqa, sqa, eo, pk);
        st_shared_u4(soA + lane_off, make_uint4(pk[0], pk[1], pk[2], pk[3]));
        st_shared_u4((soA + lane_off) ^ 16u, make_uint4(pk[4], pk[5], pk[6], pk[7]));
      } else {
        epi_finish16(vA, bv, rqa, sqa, eo, offa, va);
      }
      E2E_TR2(10);
      param_bias(p.bias2, ccB, bv);
      tmem_ld_wait();
      E2E_TR2(11);
      // every TMEM read of this warp has completed: release the accumulator
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_remote(&acc_empty[ln * 2 + 1], 0);
        else mbar_arrive(&acc_empty[ln * 2 + 1]);
      }
      if (kStaged) {
        epi_compute16(vB, bv, rqb, sqb, eo, pk);
        st_shared_u4(soB + lane_off, make_uint4(pk[0], pk[1], pk[2], pk[3]));
        st_shared_u4((soB + lane_off) ^ 16u, make_uint4(pk[4], pk[5], pk[6], pk[7]));
        fence_proxy_async_smem();   // the staged tile is read by the TMA unit (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&stage_full[ln]);
      } else {
        epi_finish16(vB, bv, rqb, sqb, eo, offb, vb);
      }
      E2E_TR2(12);
#ifdef E2E_TRACE
      if (threadIdx.x == 128 && blockIdx.x < 512) {
        g_trace[blockIdx.x][5] += tf1 - tf0;           // epi2: waiting for the accumulator
        g_trace[blockIdx.x][6] += gtime_ns() - tf1;    // epi2: work
      }
#endif
    };

    UnitIter uit;
    uit.init(u_first, u_step, 1, p.tiles_per_b);
    int pb = 0, pt0 = 0;  // unit n - 1
    for (int n = 0; n <= N; ++n) {
      E2E_TR2(0);
      if (n >= 1) prefetch_res(pb, pt0);
      E2E_TR2(1);
      const int t0 = uit.tile * p.r_out;
      if (n < N) epi1(n, t0);
      if (n >= 1) epi2(n - 1);
      pb = uit.b;
      pt0 = t0;
      uit.next();
    }
    E2E_TR2_FLUSH
  }

  tc_fence_before_sync();
  __syncwarp();
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
