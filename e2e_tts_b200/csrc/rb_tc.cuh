// Whole ResBlock1 in ONE persistent tcgen05 kernel (reference e2e_tts/models/vocoder/layers.py:33-40):
//
//     for (c1, c2) in zip(convs1, convs2):                 # n_pairs = 3, dilations d_i = 1, 3, 5
//         xt = c1(leaky_relu(x)); xt = c2(leaky_relu(xt)); x = xt + x
//
// The residual stream x never leaves the SM: it lives in TMEM as fp32 and IS c2's accumulator - the c2 MMAs of
// every pair accumulate straight onto it (x_{i+1} = x_i + W2 * m), so the residual add costs nothing and is exact
// fp32 (SURVEY.md §8 a'3); c2's biases are added when x is read (x_true = x_tmem + cumulative bias, a constant per
// channel and pair).  Per unit the kernel runs 2 * n_pairs convolutions back to back:
//
//   c1_i : A = slab P (bf16 leaky_relu(x_i), rows shifted by (j - (k-1)/2) * d_i), D = acc          (TMEM, overwritten)
//          epilogue: acc + b1 -> leaky_relu -> bf16 -> slab Q;   (pair 0 only: x <- inverse-lrelu(P) seeds TMEM)
//   c2_i : A = slab Q (rows shifted by j - (k-1)/2),                                  D = x += ...  (TMEM, accumulated)
//          epilogue: x + cum. bias -> leaky_relu -> bf16 -> slab P   (last pair: + running resblock sum, / n, -> HBM)
//
// A unit is 128*MT slab rows of one utterance; the receptive field of the chain is recomputed at the unit's ends
// (halo H = sum_i (k-1)/2 * (d_i + 1) rows per side), so 128*MT - 2H rows are stored.  Intermediate rows outside
// [0, T) are forced to zero at every layer (= each Conv1d's own zero padding, SURVEY.md §8 a'1); rows whose receptive
// field leaves the slab hold finite garbage that only ever feeds other such rows.  The host uses this kernel where
// the halo is cheap (k = 3 resblocks: H = 12) - there the three pair launches were epilogue-bound on their HBM
// round trips, which are gone here: one TMA slab load in, one tile out.
//
// Two units are in flight per CTA on two "lanes" (own slabs P/Q, own TMEM acc + x: 2 * 2 * MT * C = 512 columns) and the
// MMA warp alternates lanes job by job, so every epilogue overlaps the other lane's MMAs.
// Roles: warp 0 slab producer (TMA), warp 1 weight producer, warp 2 MMA issuer, warp 3 TMEM allocator,
// warps 4-19 epilogue (four per TMEM lane quarter, 16-column items).
#pragma once
#include "conv_tc.cuh"

namespace e2e {

constexpr int kRbMaxPairs = 3;

struct RbParams {
  int T, B;
  int panels;           // K panels of the C channels (C/64, or 1 for C = 32)
  int nt;               // C
  int taps;             // k
  int n_pairs;          // (c1, c2) pairs in the chain
  int dil[kRbMaxPairs]; // dilation of c1 of every pair
  int halo;             // H: rows of receptive field per side
  int padr;             // rows before / after the 128*MT computed rows that tap shifts may touch (multiple of 8)
  int slab_rows;        // 128*MT + 2*padr
  int box_rows;         // TMA box height (divides slab_rows, multiple of 8, <= 256)
  int r_out;            // rows stored per unit = 128*MT - 2*halo
  int tiles_per_chunk, n_chunks, n_stages, stage_bytes;   // weight ring geometry (same for every conv of the chain)
  int tiles_per_b, n_units;
  float slope_mid;      // LeakyReLU inside the chain (0.1)
  float slope;          // LeakyReLU applied to out_act (1 = none)
  float divisor;        // 0 = none
  float res_inv_slope;  // 1 / slope of the stored input activation
  const uint8_t* w[2 * kRbMaxPairs];  // packed weights in job order: c1_0, c2_0, c1_1, c2_1, ...
  float bias[2 * kRbMaxPairs][128];   // c1_i: its bias; c2_i: the CUMULATIVE c2 bias of pairs 0..i (constant bank)
  const __nv_bfloat16* sum_a;    // bf16 running sum over the stage's resblocks (generator.py:44-47) or nullptr
  int sum_tiled, out_tiled;      // tiled8 layouts (epilogue.cuh)
  int f16;                       // 16-bit tensors and operands are fp16 instead of bf16 (ptx.cuh pack16)
  __nv_bfloat16* out_act;        // bf16 leaky_relu(result, slope)
};

// registers -> TMEM: 16 consecutive fp32 columns of this thread's lane
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int ROWB, int MT>
__global__ void __launch_bounds__(kConvThreads, 1)
rb_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ RbParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  constexpr int KS = ROWB / 32;
  constexpr uint32_t ROW16 = ROWB >> 4;
  constexpr uint32_t DESC_HI = ((8u * ROWB) >> 4) | (1u << 14) | ((ROWB == 128 ? 2u : 4u) << 29);
  constexpr uint32_t SWZ = ROWB == 128 ? 7u : 3u;
  constexpr int CH_PANEL = ROWB / 2;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int panel_bytes = p.slab_rows * ROWB;
  const int slab_bytes = p.panels * panel_bytes;   // one slab (P or Q) of one lane
  const int tile_bytes = p.nt * ROWB;              // one weight tile (one tap of one panel)
  const int total_tiles = p.panels * p.taps;
  const int acc_cols = MT * p.nt;                  // TMEM columns of acc (and of x); per lane: [acc][x]
  const int n_jobs = 2 * p.n_pairs;
  const int hk = (p.taps - 1) / 2;

  uint8_t* slab_p = smem;                          // [2 lanes][panels][slab_rows][ROWB]
  uint8_t* slab_q = slab_p + 2 * slab_bytes;       // [2 lanes][panels][slab_rows][ROWB]
  uint8_t* ring = slab_q + 2 * slab_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + p.n_stages * p.stage_bytes);
  uint64_t* in_full = bars;                // [2][4]
  uint64_t* in_empty = in_full + 8;        // [2]
  uint64_t* w_full = in_empty + 2;         // [kMaxStages]
  uint64_t* w_empty = w_full + kMaxStages;
  uint64_t* acc_full = w_empty + kMaxStages;   // [2]  one completion per job of the lane
  uint64_t* epi_done = acc_full + 2;           // [2]  ... and one per epilogue of the lane
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_done + 2);

  // units of this CTA: u_n = blockIdx.x + n * gridDim.x; a unit index >= n_units is a dummy (utterance index B: the
  // TMA zero-fills, nothing is stored) so that both lanes always run in lock step
  const int N0 = (p.n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int N = (N0 + 1) & ~1;
  const int u_first = (int)blockIdx.x, u_step = (int)gridDim.x;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_in);
    for (int i = 0; i < 8; ++i) mbar_init(&in_full[i], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&in_empty[i], 1 + kEpiWarps);   // the last c1's MMAs have read slab P + every epilogue warp has seeded x from it
      mbar_init(&acc_full[i], 1);
      mbar_init(&epi_done[i], kEpiWarps);
    }
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 3) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // slab Q's pad rows are never written by an epilogue: clear the slabs once so that nothing the tensor core can read
  // is uninitialised (garbage rows must stay finite: they meet zero-weight-free taps of other garbage rows only, but a
  // NaN bit pattern in fresh shared memory would still be a NaN)
  for (int i = threadIdx.x * 16; i < 2 * slab_bytes; i += kConvThreads * 16)
    *reinterpret_cast<uint4*>(slab_q + i) = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch();

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- input slab producer (TMA): slab P of lane n & 1 <- rows [t0 - H - padr, ...) ----------------
      griddep_wait();
      const int boxes = p.slab_rows / p.box_rows;
      UnitIter uit;
      uit.init(u_first, u_step, 1, p.tiles_per_b);
      for (int n = 0; n < N; ++n, uit.next()) {
        const int b = uit.b < p.B ? uit.b : p.B;
        const int ts = uit.tile * p.r_out - p.halo - p.padr;
        const int ln = n & 1;
        mbar_wait(&in_empty[ln], ((n >> 1) & 1) ^ 1, 0x100 + ln);
        for (int pn = 0; pn < p.panels; ++pn) {
          mbar_arrive_expect_tx(&in_full[ln * 4 + pn], panel_bytes);
          uint8_t* dst = slab_p + ln * slab_bytes + pn * panel_bytes;
          for (int bx = 0; bx < boxes; ++bx)
            tma_load_3d(dst + bx * p.box_rows * ROWB, &tm_in, pn * CH_PANEL, ts + bx * p.box_rows, b,
                        &in_full[ln * 4 + pn]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- weight producer: job order (pairs of units; per job both lanes) ----------------
      uint32_t stage = 0, par = 1;
      for (int n0 = 0; n0 < N; n0 += 2)
        for (int j = 0; j < n_jobs; ++j)
          for (int ln = 0; ln < 2; ++ln) {
            const uint8_t* wsrc = p.w[j];
            int first = 0;
            for (int c = 0; c < p.n_chunks; ++c, first += p.tiles_per_chunk) {
              mbar_wait(&w_empty[stage], par, 0x200 + stage);
              const int ntile = min(p.tiles_per_chunk, total_tiles - first);
              const uint32_t bytes = ntile * tile_bytes;
              mbar_arrive_expect_tx(&w_full[stage], bytes);
              bulk_load_1d(ring + stage * p.stage_bytes, wsrc + static_cast<size_t>(first) * tile_bytes, bytes,
                           &w_full[stage]);
              if (++stage == (uint32_t)p.n_stages) {
                stage = 0;
                par ^= 1;
              }
            }
          }
    }
  } else if (warp == 2) {
    // ---------------- MMA issuer (warp-uniform loop, one elected lane issues) ----------------
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_bf16(128, p.nt, p.f16);
    const uint32_t p_lo0 = ((smem_u32(slab_p) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t q_lo0 = ((smem_u32(slab_q) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t ring_lo = ((smem_u32(ring) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t stage16 = p.stage_bytes >> 4, tile16 = tile_bytes >> 4;
    uint32_t stage = 0, wpar = 0;
    uint32_t jc[2] = {0u, 0u};   // jobs issued so far per lane
    for (int n0 = 0; n0 < N; n0 += 2)
      for (int j = 0; j < n_jobs; ++j)
        for (int ln = 0; ln < 2; ++ln) {
          const int n = n0 + ln;
          const bool is_c2 = j & 1;
          const int dil = is_c2 ? 1 : p.dil[j >> 1];
          const uint32_t src_lo = (is_c2 ? q_lo0 : p_lo0) + ln * (slab_bytes >> 4);
          // first tap reads rows (padr + 128 m - hk * dil)
          const uint32_t row0_16 = (uint32_t)(p.padr - hk * dil) * ROW16;
          const uint32_t tap_rows16 = (uint32_t)dil * ROW16;
          // the previous epilogue of this lane has drained acc / x and written the slab this job reads
          mbar_wait(&epi_done[ln], (jc[ln] & 1) ^ 1, 0x300 + ln);
          tc_fence_after_sync();
          const uint32_t d_tmem = tmem_base + (ln * 2 + (is_c2 ? 1 : 0)) * acc_cols;
          int tap = 0, pn = 0, left = total_tiles;
          uint32_t accum = is_c2 ? 1u : 0u;   // c2 accumulates onto the residual stream x
          for (int c = 0; c < p.n_chunks; ++c) {
            mbar_wait(&w_full[stage], wpar, 0x400 + stage);
            tc_fence_after_sync();
            const int ntile = min(p.tiles_per_chunk, left);
            left -= ntile;
            uint32_t b_lo = ring_lo + stage * stage16;
            for (int i = 0; i < ntile; ++i, b_lo += tile16) {
              if (tap == 0 && j == 0) {
                mbar_wait(&in_full[ln * 4 + pn], (n >> 1) & 1, 0x500 + ln * 4 + pn);
                tc_fence_after_sync();
              }
              const uint32_t s_lo = src_lo + pn * (panel_bytes >> 4) + row0_16 + tap * tap_rows16;
              if (leader) {
#pragma unroll
                for (int m = 0; m < MT; ++m) {
#pragma unroll
                  for (int ks = 0; ks < KS; ++ks) {
                    const uint64_t da = (static_cast<uint64_t>(DESC_HI) << 32) | (s_lo + m * (128 * ROW16) + ks * 2);
                    const uint64_t db = (static_cast<uint64_t>(DESC_HI) << 32) | (b_lo + ks * 2);
                    if (ks == 0)
                      umma_bf16(d_tmem + m * p.nt, da, db, idesc, accum);
                    else
                      umma_bf16_acc(d_tmem + m * p.nt, da, db, idesc);
                  }
                }
              }
              accum = 1;
              if (++tap == p.taps) {
                tap = 0;
                ++pn;
              }
            }
            if (leader) umma_commit(&w_empty[stage]);
            if (++stage == (uint32_t)p.n_stages) {
              stage = 0;
              wpar ^= 1;
            }
          }
          if (leader) {
            umma_commit(&acc_full[ln]);
            // slab P is last read by the last c1 of the chain: the next unit of this lane may be loaded
            if (j == n_jobs - 2) umma_commit(&in_empty[ln]);
          }
          ++jc[ln];
        }
  } else if (warp >= 4) {
    // ---------------- epilogue ----------------
    griddep_wait();
    const int e = warp - 4;
    const int quarter = e & 3;
    const int part = e >> 2;
    const int nchunk = p.nt >> 4;
    // MT * nchunk == 8 items per job: this warp owns items `part` and `part + 4` (fixed (m tile, 16-column chunk) pairs)
    const int mA = part / nchunk, ccA = part - mA * nchunk;
    const int mB = (part + 4) / nchunk, ccB = (part + 4) - mB * nchunk;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    EpiOut eo;
    eo.sum_a = p.sum_a;
    eo.out_f32 = nullptr;
    eo.out_act = p.out_act;
    eo.slope = p.slope;
    eo.scale = p.divisor != 0.f ? 1.0f / p.divisor : 0.f;
    eo.inv = p.res_inv_slope;
    eo.act_tanh = 0;
    eo.f16 = p.f16;
    const float smid = p.slope_mid;

    // swizzled offset of this thread's 16 columns of item (m, cc) inside a slab (second 16-byte chunk = offset ^ 16)
    auto own_off = [&](int m, int cc) -> uint32_t {
      const int n0 = cc * 16;
      uint32_t off = static_cast<uint32_t>(p.padr + m * 128 + row_in_tile) * ROWB + ((n0 % CH_PANEL) / 8) * 16;
      off ^= ((off >> 7) & SWZ) << 4;
      return off + (n0 / CH_PANEL) * panel_bytes;
    };
    const uint32_t offA = own_off(mA, ccA), offB = own_off(mB, ccB);
    const uint32_t p_addr = smem_u32(slab_p), q_addr = smem_u32(slab_q);

    auto param_bias = [&](const float* sb, int cc, float4 (&bv)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        bv[i] = make_float4(sb[cc * 16 + 4 * i], sb[cc * 16 + 4 * i + 1], sb[cc * 16 + 4 * i + 2], sb[cc * 16 + 4 * i + 3]);
    };
    // acc (or x) + bias -> leaky_relu -> bf16, zero outside the utterance -> 32 bytes of a slab row
    auto act_store_t = [&](auto f16tag, const uint32_t (&v)[16], const float4 (&bv)[4], bool inside, uint32_t dst) {
      constexpr bool F16 = decltype(f16tag)::value;
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float f0 = __uint_as_float(v[4 * i]) + bv[i].x, f1 = __uint_as_float(v[4 * i + 1]) + bv[i].y;
        const float f2 = __uint_as_float(v[4 * i + 2]) + bv[i].z, f3 = __uint_as_float(v[4 * i + 3]) + bv[i].w;
        const uint32_t h0 = pack16t<F16>(fmaxf(f0, f0 * smid), fmaxf(f1, f1 * smid));
        const uint32_t h1 = pack16t<F16>(fmaxf(f2, f2 * smid), fmaxf(f3, f3 * smid));
        pk[2 * i] = inside ? h0 : 0u;
        pk[2 * i + 1] = inside ? h1 : 0u;
      }
      st_shared_u4(dst, make_uint4(pk[0], pk[1], pk[2], pk[3]));
      st_shared_u4(dst ^ 16u, make_uint4(pk[4], pk[5], pk[6], pk[7]));
    };
    auto act_store = [&](const uint32_t (&v)[16], const float4 (&bv)[4], bool inside, uint32_t dst) {
      if (p.f16) act_store_t(std::true_type{}, v, bv, inside, dst);
      else act_store_t(std::false_type{}, v, bv, inside, dst);
    };
    // x <- inverse leaky_relu of the stored input activation (16 bf16 of this thread's row) : seeds the TMEM residual
    auto seed = [&](uint32_t src, uint32_t taddr) {
      const uint4 q0 = ld_shared_u4(src), q1 = ld_shared_u4(src ^ 16u);
      const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
      uint32_t xv[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float lo, hi;
        unpack16(w[j], lo, hi, p.f16);
        xv[2 * j] = __float_as_uint(fminf(lo, lo * eo.inv));
        xv[2 * j + 1] = __float_as_uint(fminf(hi, hi * eo.inv));
      }
      tmem_st_32x16(taddr, xv);
    };

    uint32_t jc[2] = {0u, 0u};
    UnitIter uit;
    uit.init(u_first, u_step, 1, p.tiles_per_b);
    int ub[2], ut0[2];
    for (int n0 = 0; n0 < N; n0 += 2) {
      for (int ln = 0; ln < 2; ++ln, uit.next()) {
        ub[ln] = uit.b;
        ut0[ln] = uit.tile * p.r_out;   // first stored row of the unit
      }
      for (int j = 0; j < n_jobs; ++j)
        for (int ln = 0; ln < 2; ++ln) {
          const bool is_c2 = j & 1;
          const bool last = j == n_jobs - 1;
          const int b = ub[ln], t0 = ut0[ln];
          const int ts = t0 - p.halo;                // time of computed row 0
          const int rA = mA * 128 + row_in_tile, rB = mB * 128 + row_in_tile;
          const bool inA = ts + rA >= 0 && ts + rA < p.T, inB = ts + rB >= 0 && ts + rB < p.T;
          const uint32_t acc_t = tmem_base + (ln * 2) * acc_cols + lane_sel;
          const uint32_t x_t = acc_t + acc_cols;
          const uint32_t lane_off = ln * slab_bytes;
          uint32_t vA[16], vB[16];
          float4 bv[4];
          // running resblock sum of the rows this thread stores (last job only), fetched before the wait
          uint4 sqa[2], sqb[2], zq[2];
          zq[0] = zq[1] = make_uint4(0u, 0u, 0u, 0u);
          size_t goffA = 0, goffB = 0;
          bool va = false, vb = false;
          if (last) {
            va = rA >= p.halo && rA < p.halo + p.r_out && ts + rA < p.T && b < p.B;
            vb = rB >= p.halo && rB < p.halo + p.r_out && ts + rB < p.T && b < p.B;
            goffA = (static_cast<size_t>(b) * p.T + (va ? ts + rA : 0)) * p.nt + ccA * 16;
            goffB = (static_cast<size_t>(b) * p.T + (vb ? ts + rB : 0)) * p.nt + ccB * 16;
            const int t8 = (p.T + 7) >> 3, c16 = p.nt >> 4;
            const size_t ta = (p.sum_tiled | p.out_tiled) && va ? tiled8_off(b, ts + rA, ccA, t8, c16) : 0;
            const size_t tb = (p.sum_tiled | p.out_tiled) && vb ? tiled8_off(b, ts + rB, ccB, t8, c16) : 0;
            if (p.sum_a) {
              if (va) ld_global_256(p.sum_a + (p.sum_tiled ? ta : goffA), sqa[0], sqa[1]);
              if (vb) ld_global_256(p.sum_a + (p.sum_tiled ? tb : goffB), sqb[0], sqb[1]);
            }
            if (p.out_tiled) {
              goffA = ta;
              goffB = tb;
            }
          }
          mbar_wait(&acc_full[ln], jc[ln] & 1, 0x600 + ln);
          if (j == 0) {   // the seed reads slab P, which the TMA unit wrote: observe its barrier (long complete)
            const uint32_t ipar = ((n0 + ln) >> 1) & 1;
            mbar_wait(&in_full[ln * 4 + (ccA * 16) / CH_PANEL], ipar, 0x680 + ln);
            mbar_wait(&in_full[ln * 4 + (ccB * 16) / CH_PANEL], ipar, 0x688 + ln);
          }
          tc_fence_after_sync();
          const uint32_t src_t = is_c2 ? x_t : acc_t;
          tmem_ld_32x16(src_t + mA * p.nt + ccA * 16, vA);
          param_bias(p.bias[j], ccA, bv);
          tmem_ld_wait();
          tmem_ld_32x16(src_t + mB * p.nt + ccB * 16, vB);
          if (!last) {
            const uint32_t dst = (is_c2 ? p_addr : q_addr) + lane_off;
            if (j == 0) seed(p_addr + lane_off + offA, x_t + mA * p.nt + ccA * 16);
            act_store(vA, bv, inA, dst + offA);
            param_bias(p.bias[j], ccB, bv);
            tmem_ld_wait();
            if (j == 0) seed(p_addr + lane_off + offB, x_t + mB * p.nt + ccB * 16);
            act_store(vB, bv, inB, dst + offB);
            if (j == 0) {
              tmem_st_wait();
              __syncwarp();
              // this warp's reads of slab P (the seed) are done: with one pair in the chain the MMAs of job 0 are the
              // last readers the producer would otherwise wait for, and the next unit's TMA load could overtake the seed
              if (lane == 0) mbar_arrive(&in_empty[ln]);
            }
            fence_proxy_async_smem();   // the slab is read by the tensor core through the async proxy
          } else {
            epi_finish16(vA, bv, zq, sqa, eo, goffA, va);
            param_bias(p.bias[j], ccB, bv);
            tmem_ld_wait();
            epi_finish16(vB, bv, zq, sqb, eo, goffB, vb);
          }
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&epi_done[ln]);
          ++jc[ln];
        }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 3) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace e2e
