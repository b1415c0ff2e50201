// Fused residual pair of ResBlock1 for the C = 32 stage with FOUR TIME STEPS PER GEMM ROW (reference
// e2e_tts/models/vocoder/layers.py:34-39; VERDICT round 1, next #5: "put 2-4 time steps on N").
//
//     xt = c1(leaky_relu(x)); xt = c2(leaky_relu(xt)); x = xt + x          (c1: k taps, dilation d; c2: k taps, dilation 1)
//
// Why.  With time on M and C = 32 output channels on N, a 128 x 32 x 16 MMA costs 54-65 cycles - the 4 KB A-operand read
// from shared memory, not the math (profiles/r02_umma_rate_microbench.log) - so the stage sat at its MMA-issue floor.  The
// channels-last tensor [B][T][32] is, byte for byte, also [B][T/4][128]: a "super-row" R holds the time steps 4R..4R+3.
// In that view a k-tap convolution is a GEMM with N = 128 output columns (sub-step i, channel co):
//
//     out[R][(i, co)] = sum over input time offsets s (relative to 4R), s = -h .. 3+h,  h = (k-1)/2 * d:
//                       A_s[R][ci]  *  B_s[(i, co)][ci],     A_s = x[4R + s],   B_s[(i, co)] = W[tap with offset s - i]
//
//   * A_s is the SAME shared-memory slab for every s: super-row shift q = floor(s / 4) moves the descriptor's start
//     address by q rows, the sub-step j = s mod 4 selects a 64-byte column range of the 256-byte super-row.
//   * B_s is a SLIDING WINDOW over one small array: with V[y] = W[tap offset -y] (zero blocks where no tap exists),
//     B_s = blocks V[-s .. -s+3], i.e. 128 consecutive rows of a (k + 6)-block array of 2 KB blocks - the block-Toeplitz
//     matrix is never materialised (round 2 had costed it at 16x the weight bytes and rejected it).  One N = 128 MMA
//     (64 cycles) then does the work of up to four N = 32 MMAs: k = 11, d = 1: 2 * (2h + 4) = 28 MMAs per 512 time steps
//     instead of 88.
//   * A dilated c1 (d > 1) gains nothing from windows (4 consecutive offsets hold at most two taps): it runs as N = 32
//     MMAs into the 32-column block of its sub-step, from the ordinary per-tap weight image - the same MMA count as
//     pair_tc.cuh, but from the same super-row slab and into the same TMEM tile.
//   * Both weight images (<= 68 KB) are loaded into shared memory ONCE per CTA and stay resident: no per-unit weight
//     stream from L2 (pair_tc.cuh re-streams 44 KB per 512-row unit at k = 11).
//
// Pipeline.  A unit is 128 super-rows of one utterance (c1 computes all of them from a slab with `padr` super-rows of real
// context on either side; c2 loses ceil((k-1)/2 / 4) super-rows per side, so 128 - 2 * halo are stored).  Two units are
// in flight on two lanes (own slabs P / Q, own TMEM acc + x: 2 x 256 columns), in the software-pipelined job order of
// pair_tc.cuh:       MMA warp   c1(u0) | c1(u1) c2(u0) | c1(u2) c2(u1) | ...
//                    epilogue   e1(u0) | e1(u1) e2(u0) | e1(u2) e2(u1) | ...
// so every epilogue overlaps the MMAs of the other lane, and no MMA job ever waits for an epilogue that stores to HBM.
// The residual stream x lives in TMEM as fp32 (rb_tc.cuh): e1 seeds it from the input slab (inverse LeakyReLU of the
// stored activation), c2 accumulates straight onto it, e2 reads x + bias2 (+ running resblock sum, / n), applies the
// next LeakyReLU and stores.  An epilogue item = 16 channels of ONE time step of the thread's super-row, so biases and
// the tiled8 running-sum layout of the [T][32] view (epilogue.cuh) stay addressable per item; intermediate rows outside
// [0, T) are forced to zero (T % 4 == 0: a super-row is inside or outside as a whole).
// Roles: warp 0 slab producer (TMA), warp 1 weight loader, warp 2 MMA issuer, warp 3 TMEM allocator, warps 4-19 epilogue.
#pragma once
#include "rb_tc.cuh"

namespace e2e {

constexpr int kTzC = 32;       // channels of the stage this kernel serves
constexpr int kTzG = 4;        // time steps per super-row (128 / C)
constexpr int kTzBlock = 2048; // bytes of one [32 co][32 ci] weight block (64-byte rows, SWIZZLE_64B)

struct TzParams {
  int T4, B;            // super-rows per utterance (T / 4), utterances
  int T;                // time steps per utterance (tiled8 addressing of the running sums)
  int taps;             // k
  int dil;              // dilation of c1
  int halo;             // super-rows per side whose receptive field leaves the unit (not stored)
  int padr;             // super-rows before / after the 128 computed rows that tap shifts may touch (multiple of 4)
  int slab_rows;        // 128 + 2 * padr
  int r_out;            // super-rows stored per unit = 128 - 2 * halo
  int tiles_per_b, n_units;
  float slope_mid, slope, divisor, res_inv_slope;
  const uint8_t* w1;    // c1: sliding-window array (dil == 1) or per-tap image (dil > 1)
  const uint8_t* w2;    // c2: sliding-window array
  int w1_bytes, w2_bytes;
  int w1_toep;          // 1: c1 runs as sliding windows (N = 128 MMAs), 0: per-tap tiles (N = 32 MMAs)
  alignas(16) float bias1[kTzC];   // (16-byte aligned: the epilogues read them as constant-bank pairs)
  alignas(16) float bias2[kTzC];   // c2's bias: x lives in TMEM without it
  const __nv_bfloat16* sum_a;    // running sum over the stage's resblocks (generator.py:44-47) or nullptr
  int sum_tiled, out_tiled;      // tiled8 layouts of the [T][32] view (epilogue.cuh)
  int no_sum_prefetch;           // experiments (E2E_NO_SUM_PREFETCH=1): no L2 prefetch of the running sum by the producer
  int f16;                       // 16-bit tensors and operands are fp16 instead of bf16
  int dbg;                       // experiments only (tests/cuda): 1 = no seed, 2 = no intermediate slab stores
  __nv_bfloat16* out_act;        // leaky_relu(result, slope)
};

// packed fp32 pairs (FADD2 / FMUL2 on sm_100) for the adds and multiplies of the two epilogues; bit-identical to the scalar
// forms.  (The same change in the shared epilogues of conv_tc / pair_tc / rb_tc was measured over the whole forward and
// is not kept: profiles/r02_experiments_notes.md §10.)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

__global__ void __launch_bounds__(kConvThreads, 1)
pair_tz_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ TzParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  constexpr int ROWB = 128;                      // bytes per panel row of a slab: 64 super-channels = 2 time steps x 32 ch
  constexpr uint32_t A_HI = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);   // SWIZZLE_128B, 8-row groups 1024 B apart
  constexpr uint32_t B_HI = ((8u * 64u) >> 4) | (1u << 14) | (4u << 29);    // SWIZZLE_64B, 8-row groups 512 B apart
  constexpr uint32_t BLOCK16 = kTzBlock >> 4;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int panel_bytes = p.slab_rows * ROWB;
  const int slab_bytes = 2 * panel_bytes;          // one slab (P or Q) of one lane: two 64-super-channel panels
  const int hk = (p.taps - 1) / 2;

  uint8_t* slab_p = smem;                          // [2 lanes][2 panels][slab_rows][128]
  uint8_t* slab_q = slab_p + 2 * slab_bytes;
  uint8_t* wbuf = slab_q + 2 * slab_bytes;         // resident weight images: c1, then c2
  uint64_t* bars = reinterpret_cast<uint64_t*>(wbuf + p.w1_bytes + p.w2_bytes);
  uint64_t* in_full = bars;                // [2] slab P loaded
  uint64_t* in_empty = in_full + 2;        // [2] c1's MMAs have read slab P and every epilogue warp has seeded x from it
  uint64_t* w_full = in_empty + 2;         // [1]
  uint64_t* acc_full = w_full + 1;         // [2] c1 complete
  uint64_t* m_full = acc_full + 2;         // [2] e1 done: slab Q written, x seeded, acc drained
  uint64_t* x_full = m_full + 2;           // [2] c2 complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_full + 2);
  uint2* tab1 = reinterpret_cast<uint2*>(bars + 16);   // MMA schedule of c1: <= 4 * kMaxTaps entries
  uint2* tab2 = tab1 + 4 * kMaxTaps;                   // ... and of c2: 2 * hk + 4 entries

  // units of this CTA: u_n = blockIdx.x + n * gridDim.x, n = 0 .. N-1; lane of unit n = n & 1
  const int N = (p.n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int u_first = (int)blockIdx.x, u_step = (int)gridDim.x;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_in);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&in_full[i], 1);
      mbar_init(&in_empty[i], 1 + kEpiWarps);
      mbar_init(&acc_full[i], 1);
      mbar_init(&m_full[i], kEpiWarps);
      mbar_init(&x_full[i], 1);
    }
    mbar_init(w_full, 1);
    fence_mbar_init();
  }
  if (warp == 3) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // slab Q's pad rows are never written by an epilogue: clear the slabs once (nothing the tensor core reads may hold a NaN)
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
      // ---------------- input slab producer (TMA): slab P of lane n & 1 <- super-rows [t0 - halo - padr, ...) ----------------
      griddep_wait();
      UnitIter uit;
      uit.init(u_first, u_step, 1, p.tiles_per_b);
      for (int n = 0; n < N; ++n, uit.next()) {
        const int ts = uit.tile * p.r_out - p.halo - p.padr;
        const int ln = n & 1;
        mbar_wait(&in_empty[ln], ((n >> 1) & 1) ^ 1, 0x100 + ln);
        mbar_arrive_expect_tx(&in_full[ln], slab_bytes);
        uint8_t* dst = slab_p + ln * slab_bytes;
        tma_load_3d(dst, &tm_in, 0, ts, uit.b, &in_full[ln]);
        tma_load_3d(dst + panel_bytes, &tm_in, 64, ts, uit.b, &in_full[ln]);
        if (p.sum_a && !p.no_sum_prefetch) {   // rows of the [T][32] view this unit stores; see prefetch_sum_rows (epilogue.cuh)
          const int t0 = kTzG * uit.tile * p.r_out;
          prefetch_sum_rows(p.sum_a, p.sum_tiled, uit.b, t0, t0 + kTzG * p.r_out, p.T, kTzC);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- weights: both images, once (static data: no dependency on the previous kernel) ----------------
      mbar_arrive_expect_tx(w_full, (uint32_t)(p.w1_bytes + p.w2_bytes));
      for (int done = 0; done < p.w1_bytes; done += 16384)
        bulk_load_1d(wbuf + done, p.w1 + done, (uint32_t)min(16384, p.w1_bytes - done), w_full);
      for (int done = 0; done < p.w2_bytes; done += 16384)
        bulk_load_1d(wbuf + p.w1_bytes + done, p.w2 + done, (uint32_t)min(16384, p.w2_bytes - done), w_full);
    }
  } else if (warp == 2) {
    // ---------------- MMA issuer (warp-uniform loop, one elected lane issues) ----------------
    const bool leader = elect_one();
    const uint32_t idesc128 = umma_idesc_bf16(128, 128, p.f16), idesc32 = umma_idesc_bf16(128, 32, p.f16);
    const uint32_t p_lo0 = ((smem_u32(slab_p) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t q_lo0 = ((smem_u32(slab_q) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t w1_lo = ((smem_u32(wbuf) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t w2_lo = w1_lo + (p.w1_bytes >> 4);
    const int panel16 = panel_bytes >> 4, slab16 = slab_bytes >> 4;
    // The MMA schedule is the same for every unit: one table entry per MMA pair (two 16-channel K steps), built once.
    //   x = A offset (16-byte units: panel, super-row shift, 64-byte column range) | TMEM column offset << 16
    //   y = B offset (16-byte units into the job's weight image)                   | "overwrites the accumulator" << 16
    // (The issuing thread would otherwise spend ~150 cycles of dependent integer work per pair: measured 76 cycles per MMA.)
    // A operand of input time offset s: super-row shift s >> 2, 64-byte column range (s & 3) of the 256-byte super-row.
    auto a_off16 = [&](int s) -> uint32_t {
      return (uint32_t)(((s & 3) >> 1) * panel16 + (p.padr + (s >> 2)) * 8 + (s & 1) * 4);
    };
    const int h1 = hk * p.dil;
    const int g1 = p.w1_toep ? 2 * h1 + kTzG : kTzG * p.taps;
    const int g2 = 2 * hk + kTzG;
    for (int g = lane; g < g1; g += 32) {
      uint2 e;
      if (p.w1_toep) {
        // sliding windows: input time offset s = g - h reads blocks [3 + h - s, ...) of the window array, N = 128
        const int sft = g - h1;
        e.x = a_off16(sft);
        e.y = (uint32_t)(kTzG - 1 + h1 - sft) * BLOCK16 | (g == 0 ? 1u << 16 : 0u);
      } else {
        // per-tap tiles: sub-step i of the output super-row reads input offset i + (m - hk) * dil through tap m, N = 32
        const int i = g / p.taps, m = g - i * p.taps;
        e.x = a_off16(i + (m - hk) * p.dil) | (uint32_t)(i * kTzC) << 16;
        e.y = (uint32_t)m * BLOCK16 | (m == 0 ? 1u << 16 : 0u);
      }
      tab1[g] = e;
    }
    for (int g = lane; g < g2; g += 32) {
      const int sft = g - hk;
      tab2[g] = make_uint2(a_off16(sft), (uint32_t)(kTzG - 1 + hk - sft) * BLOCK16);   // c2 always accumulates (onto x)
    }
    __syncwarp();
    auto run = [&](const uint2* tab, int ng, uint32_t src_lo, uint32_t wb, uint32_t d_tmem, uint32_t idesc) {
#pragma unroll 4
      for (int g = 0; g < ng; ++g) {
        const uint2 e = tab[g];
        if (leader) {
          const uint64_t da = (static_cast<uint64_t>(A_HI) << 32) | (src_lo + (e.x & 0xFFFFu));
          const uint64_t db = (static_cast<uint64_t>(B_HI) << 32) | (wb + (e.y & 0xFFFFu));
          umma_bf16(d_tmem + (e.x >> 16), da, db, idesc, (e.y >> 16) ^ 1u);
          umma_bf16_acc(d_tmem + (e.x >> 16), da + 2, db + 2, idesc);
        }
      }
    };
    mbar_wait(w_full, 0, 0x200);
    tc_fence_after_sync();
#ifdef E2E_TZTRACE
    long long tr_m[4] = {0, 0, 0, 0};   // waiting for the input slab / for e1 (slab Q, x seed); issuing c1 / c2
    const long long tr_m0 = clock64();
#endif
    for (int n = 0; n <= N; ++n) {
      if (n < N) {
        // ---- c1(u_n): A = slab P, D = acc (overwritten; e1(u_{n-2}) drained it before c2(u_{n-2}) was issued) ----
        const int ln = n & 1;
#ifdef E2E_TZTRACE
        const long long t0 = clock64();
#endif
        mbar_wait(&in_full[ln], (n >> 1) & 1, 0x500 + ln);
        tc_fence_after_sync();
#ifdef E2E_TZTRACE
        const long long t1 = clock64();
        tr_m[0] += t1 - t0;
#endif
        const uint32_t src_lo = p_lo0 + ln * slab16;
        const uint32_t d_tmem = tmem_base + ln * 256;
        run(tab1, g1, src_lo, w1_lo, d_tmem, p.w1_toep ? idesc128 : idesc32);
        if (leader) {
          umma_commit(&acc_full[ln]);
          umma_commit(&in_empty[ln]);   // slab P: read by these MMAs (and by e1's seed, which arrives there too)
        }
#ifdef E2E_TZTRACE
        tr_m[2] += clock64() - t1;
#endif
      }
      if (n >= 1) {
        // ---- c2(u_{n-1}): A = slab Q, D = x (accumulated onto the seeded residual) ----
        const int ln = (n - 1) & 1;
#ifdef E2E_TZTRACE
        const long long t0 = clock64();
#endif
        mbar_wait(&m_full[ln], ((n - 1) >> 1) & 1, 0x300 + ln);
        tc_fence_after_sync();
#ifdef E2E_TZTRACE
        const long long t1 = clock64();
        tr_m[1] += t1 - t0;
#endif
        run(tab2, g2, q_lo0 + ln * slab16, w2_lo, tmem_base + ln * 256 + 128, idesc128);
        if (leader) umma_commit(&x_full[ln]);
#ifdef E2E_TZTRACE
        tr_m[3] += clock64() - t1;
#endif
      }
    }
#ifdef E2E_TZTRACE
    if (leader && blockIdx.x == 73)
      printf("  trace cta 73 MMA warp, cycles per unit (%d units): wait slab=%lld wait e1=%lld issue c1=%lld issue c2=%lld | total=%lld\n",
             N, tr_m[0] / N, tr_m[1] / N, tr_m[2] / N, tr_m[3] / N, (clock64() - tr_m0) / N);
#endif
  } else if (warp >= 4) {
    // ---------------- epilogue ----------------
    griddep_wait();
    const int e = warp - 4;
    const int quarter = e & 3;
    const int part = e >> 2;
    // 8 items of 16 columns per job: this warp owns items `part` and `part + 4`.  Item cc = 16 channels (half cc & 1) of
    // sub-step cc >> 1 of the super-row.
    const int ccA = part, ccB = part + 4;
    const int row = quarter * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    const float smid = p.slope_mid, inv = p.res_inv_slope, slope = p.slope;
    const float scale = p.divisor != 0.f ? 1.0f / p.divisor : 1.0f;
    const bool use_scale = p.divisor != 0.f;

    // swizzled offset of this thread's 16 columns of item cc inside a slab (second 16-byte chunk = offset ^ 16)
    auto own_off = [&](int cc) -> uint32_t {
      const int n0 = cc * 16;
      uint32_t off = static_cast<uint32_t>(p.padr + row) * ROWB + ((n0 % 64) / 8) * 16;
      off ^= ((off >> 7) & 7u) << 4;
      return off + (n0 / 64) * panel_bytes;
    };
    const uint32_t offA = own_off(ccA), offB = own_off(ccB);
    const uint32_t p_addr = smem_u32(slab_p), q_addr = smem_u32(slab_q);
    // biases of this warp's two items: both items are the same channel half (part & 1) of different sub-steps
    const int c0 = (part & 1) * 16;

    // acc + bias1 -> leaky_relu -> 16-bit -> 32 bytes of a slab Q row (zeros outside the utterance)
    auto mid_store_t = [&](auto f16tag, const uint32_t (&v)[16], bool inside, uint32_t dst) {
      constexpr bool F16 = decltype(f16tag)::value;
      if (inside) {
        uint32_t pk[8];
        const f32x2 s2 = f2_pack(smid, smid);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const f32x2 f = f2_add(f2_pack(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])),
                                               f2_pack(p.bias1[c0 + 2 * i], p.bias1[c0 + 2 * i + 1]));
          const f32x2 g = f2_mul(f, s2);
          float f0, f1, g0, g1;
          f2_unpack(f, f0, f1);
          f2_unpack(g, g0, g1);
          pk[i] = pack16t<F16>(fmaxf(f0, g0), fmaxf(f1, g1));
        }
        st_shared_u4(dst, make_uint4(pk[0], pk[1], pk[2], pk[3]));
        st_shared_u4(dst ^ 16u, make_uint4(pk[4], pk[5], pk[6], pk[7]));
      } else {
        st_shared_u4(dst, make_uint4(0u, 0u, 0u, 0u));
        st_shared_u4(dst ^ 16u, make_uint4(0u, 0u, 0u, 0u));
      }
    };
    auto mid_store = [&](const uint32_t (&v)[16], bool inside, uint32_t dst) {
      if (p.f16) mid_store_t(std::true_type{}, v, inside, dst);
      else mid_store_t(std::false_type{}, v, inside, dst);
    };
    // x <- inverse leaky_relu of the stored input activation (16 values of this thread's super-row): seeds the TMEM residual
    auto seed = [&](uint32_t src, uint32_t taddr) {
      const uint4 q0 = ld_shared_u4(src), q1 = ld_shared_u4(src ^ 16u);
      const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
      uint32_t xv[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float lo, hi;
        unpack16(w[j], lo, hi, p.f16);
        xv[2 * j] = __float_as_uint(fminf(lo, lo * inv));
        xv[2 * j + 1] = __float_as_uint(fminf(hi, hi * inv));
      }
      tmem_st_32x16(taddr, xv);
    };
    // x + bias2 (+ running sum) (* 1/n) -> leaky_relu -> 16-bit -> 32 bytes of the output
    auto fin_store_t = [&](auto f16tag, const uint32_t (&v)[16], const uint4 (&sq)[2], bool has_sum, size_t off) {
      constexpr bool F16 = decltype(f16tag)::value;
      const uint32_t sw[8] = {sq[0].x, sq[0].y, sq[0].z, sq[0].w, sq[1].x, sq[1].y, sq[1].z, sq[1].w};
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        f32x2 f = f2_add(f2_pack(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])),
                                       f2_pack(p.bias2[c0 + 2 * i], p.bias2[c0 + 2 * i + 1]));
        if (has_sum) {
          float lo, hi;
          unpack16t<F16>(sw[i], lo, hi);
          f = f2_add(f, f2_pack(lo, hi));
        }
        if (use_scale) f = f2_mul(f, f2_pack(scale, scale));
        const f32x2 g = f2_mul(f, f2_pack(slope, slope));
        float f0, f1, g0, g1;
        f2_unpack(f, f0, f1);
        f2_unpack(g, g0, g1);
        pk[i] = pack16t<F16>(fmaxf(f0, g0), fmaxf(f1, g1));   // leaky_relu, 0 < slope <= 1
      }
      st_global_256(p.out_act + off, make_uint4(pk[0], pk[1], pk[2], pk[3]), make_uint4(pk[4], pk[5], pk[6], pk[7]));
    };
    auto fin_store = [&](const uint32_t (&v)[16], const uint4 (&sq)[2], bool has_sum, size_t off) {
      if (p.f16) fin_store_t(std::true_type{}, v, sq, has_sum, off);
      else fin_store_t(std::false_type{}, v, sq, has_sum, off);
    };

#ifdef E2E_TZTRACE
    long long tr_e[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // e1: wait, first ld, items, arrive; e2: prefetch, wait, lds, items
    const long long tr_e0 = clock64();
#endif
    UnitIter uit;
    uit.init(u_first, u_step, 1, p.tiles_per_b);
    int pb = 0, pts = 0;   // unit n - 1: utterance, super-row of computed row 0
    for (int n = 0; n <= N; ++n) {
      // ---- e2(u_{n-1}), part 1: where this thread's values go; the running sum is fetched before any waiting ----
      uint4 sqa[2], sqb[2];
      size_t goffA = 0, goffB = 0;
      bool valid = false;
#ifdef E2E_TZTRACE
      const long long tp0 = clock64();
#endif
      if (n >= 1) {
        valid = row >= p.halo && row < p.halo + p.r_out && pts + row < p.T4 && p.out_act != nullptr;
        if (valid) {
          const int R = pts + row;
          goffA = (static_cast<size_t>(pb) * p.T4 + R) * 128 + ccA * 16;
          goffB = goffA + 64;
          size_t sa = goffA, sb = goffB;
          if (p.sum_tiled | p.out_tiled) {
            // tiled8 layout of the [T][32] view: item cc = time step 4R + (cc >> 1), 16-channel chunk cc & 1
            const int t8 = (p.T + 7) >> 3;
            const size_t ta = tiled8_off(pb, kTzG * R + (ccA >> 1), ccA & 1, t8, kTzC / 16);
            const size_t tb = tiled8_off(pb, kTzG * R + (ccB >> 1), ccB & 1, t8, kTzC / 16);
            if (p.sum_tiled) {
              sa = ta;
              sb = tb;
            }
            if (p.out_tiled) {
              goffA = ta;
              goffB = tb;
            }
          }
          if (p.sum_a) {
            ld_global_256(p.sum_a + sa, sqa[0], sqa[1]);
            ld_global_256(p.sum_a + sb, sqb[0], sqb[1]);
          }
        }
      }
#ifdef E2E_TZTRACE
      const long long tp1 = clock64();
      tr_e[4] += tp1 - tp0;
#endif
      if (n < N) {
        // ---- e1(u_n): acc + bias1 -> leaky_relu -> slab Q; x <- inverse-lrelu(slab P) ----
        const int ln = n & 1;
        const uint32_t par = (n >> 1) & 1;
        const int ts = uit.tile * p.r_out - p.halo;   // super-row of computed row 0
        const bool inside = ts + row >= 0 && ts + row < p.T4;
        const uint32_t acc_t = tmem_base + ln * 256 + lane_sel;
        const uint32_t x_t = acc_t + 128;
        const uint32_t lane_off = ln * slab_bytes;
        uint32_t vA[16], vB[16];
        mbar_wait(&acc_full[ln], par, 0x600 + ln);
        mbar_wait(&in_full[ln], par, 0x680 + ln);   // the seed reads slab P, which the TMA unit wrote (long complete)
        tc_fence_after_sync();
#ifdef E2E_TZTRACE
        const long long t1 = clock64();
        tr_e[0] += t1 - tp1;
#endif
        tmem_ld_32x16(acc_t + ccA * 16, vA);
        tmem_ld_32x16(acc_t + ccB * 16, vB);
        if (!(p.dbg & 1)) {   // the seeds run while the accumulator loads are in flight
          seed(p_addr + lane_off + offA, x_t + ccA * 16);
          seed(p_addr + lane_off + offB, x_t + ccB * 16);
        }
        tmem_ld_wait();
#ifdef E2E_TZTRACE
        const long long t2 = clock64();
        tr_e[1] += t2 - t1;
#endif
        if (!(p.dbg & 2)) {
          mid_store(vA, inside, q_addr + lane_off + offA);
          mid_store(vB, inside, q_addr + lane_off + offB);
        }
        tmem_st_wait();
        fence_proxy_async_smem();   // slab Q is read by the tensor core through the async proxy
#ifdef E2E_TZTRACE
        const long long t3 = clock64();
        tr_e[2] += t3 - t2;
#endif
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&m_full[ln]);
          mbar_arrive(&in_empty[ln]);   // this warp's reads of slab P (the seed) are done
        }
#ifdef E2E_TZTRACE
        tr_e[3] += clock64() - t3;
#endif
        pb = uit.b;
        pts = ts;
        uit.next();
      }
      if (n >= 1) {
        // ---- e2(u_{n-1}), part 2: x + bias2 (+ sum) -> leaky_relu -> HBM.  Nothing waits for this epilogue: the next
        // reader / writer of x on this lane is this very thread (e1 two units later). ----
        const int ln = (n - 1) & 1;
#ifdef E2E_TZTRACE
        const long long t0 = clock64();
#endif
        mbar_wait(&x_full[ln], ((n - 1) >> 1) & 1, 0x700 + ln);
        tc_fence_after_sync();
#ifdef E2E_TZTRACE
        const long long t1 = clock64();
        tr_e[5] += t1 - t0;
#endif
        const uint32_t x_t = tmem_base + ln * 256 + 128 + lane_sel;
        uint32_t vA[16], vB[16];
        tmem_ld_32x16(x_t + ccA * 16, vA);
        tmem_ld_32x16(x_t + ccB * 16, vB);
        tmem_ld_wait();
#ifdef E2E_TZTRACE
        const long long t2 = clock64();
        tr_e[6] += t2 - t1;
#endif
        if (valid) {
          fin_store(vA, sqa, p.sum_a != nullptr, goffA);
          fin_store(vB, sqb, p.sum_a != nullptr, goffB);
        }
#ifdef E2E_TZTRACE
        tr_e[7] += clock64() - t2;
#endif
      }
    }
#ifdef E2E_TZTRACE
    if (threadIdx.x == 128 && blockIdx.x == 73)
      printf("  trace cta 73 epilogue warp 4, cycles per unit: e1: wait=%lld ld+seed=%lld store=%lld arrive=%lld | e2: prefetch=%lld wait=%lld ld=%lld store=%lld | total=%lld\n",
             tr_e[0] / N, tr_e[1] / N, tr_e[2] / N, tr_e[3] / N, tr_e[4] / N, tr_e[5] / N, tr_e[6] / N, tr_e[7] / N,
             (clock64() - tr_e0) / N);
#endif
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 3) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace e2e
