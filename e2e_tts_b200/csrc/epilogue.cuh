// Epilogue pieces shared by conv_tc.cuh and pair_tc.cuh.  An epilogue "item" is 16 consecutive output columns of
// one accumulator row: the thread that owns TMEM lane r reads them with one tcgen05.ld.32x32b.x16, and all global
// traffic of the item is 32-byte aligned 256-bit accesses (one full sector per lane and instruction):
//   residual  16 bf16 = 32 B   (leaky_relu(x) of the pair's input; x recovered by the inverse LeakyReLU)
//   sum_a     16 bf16 = 32 B   (running sum over the resblocks of the stage, generator.py:44-47)
//   out_f32   16 fp32 = 64 B
//   out_act   16 bf16 = 32 B   (leaky_relu(result, slope); slope 1 stores the plain result)
#pragma once
#include "ptx.cuh"

namespace e2e {

constexpr int kEpiWarps = 16;                      // four warps per TMEM lane quarter
constexpr int kConvThreads = (4 + kEpiWarps) * 32;  // + TMA producer, weight producer, MMA issuer, TMEM allocator

// TMEM -> registers: 16 consecutive fp32 columns of this thread's lane (asynchronous until tmem_ld_wait()).
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// Persistent-grid work iterator: unit u = (b * tiles_per_b + tile) * n_tiles + nti, visited with stride `step`.
// The decomposition is carried incrementally (four integer divisions per thread and kernel instead of several
// per unit and epilogue call).
struct UnitIter {
  int nti, tile, b;
  int s_nti, s_tile, s_b;
  int n_tiles, tiles_per_b;
  __device__ __forceinline__ void init(int u0, int step, int n_tiles_, int tiles_per_b_) {
    n_tiles = n_tiles_;
    tiles_per_b = tiles_per_b_;
    nti = u0 % n_tiles;
    int tb = u0 / n_tiles;
    b = tb / tiles_per_b;
    tile = tb - b * tiles_per_b;
    s_nti = step % n_tiles;
    tb = step / n_tiles;
    s_b = tb / tiles_per_b;
    s_tile = tb - s_b * tiles_per_b;
  }
  __device__ __forceinline__ void next() {
    nti += s_nti;
    int carry = 0;
    if (nti >= n_tiles) {
      nti -= n_tiles;
      carry = 1;
    }
    tile += s_tile + carry;
    b += s_b;
    if (tile >= tiles_per_b) {
      tile -= tiles_per_b;
      ++b;
    }
  }
};

// Tiled layout of the running-sum tensors S0 / S1 (touched only by epilogues, never by TMA): [B][ceil(T/8)][C/16][8][16].
// A warp's natural access - 32 consecutive rows x 16 channels, 32 bytes per lane - is then four contiguous 256-byte
// blocks (8 lines) instead of 32 bytes at a row pitch of 2*C bytes (32 lines, ~64 L1 wavefronts per warp access).
__device__ __forceinline__ size_t tiled8_off(int b, int t, int chunk16, int t8, int c16) {
  return ((static_cast<size_t>(b) * t8 + (t >> 3)) * c16 + chunk16) * 128 + (t & 7) * 16;
}

// The running sum a unit's last epilogue will add is read once, two or three launches after it was written: by then it
// has left L2, and the epilogue's loads (issued one job ahead) still see most of the DRAM latency - the third pair of
// every resblock ran 10-36 us longer than its siblings (profiles/r02_launches_g_time_dram.txt).  The slab producer
// therefore asks for the unit's rows [t0, t1) of utterance b - one contiguous range in either layout - to be brought
// into L2 when it loads the unit's input slab, two units ahead of that epilogue.  C = channels, T = rows per utterance.
__device__ __forceinline__ void prefetch_sum_rows(const __nv_bfloat16* sum, int tiled, int b, int t0, int t1, int T, int C) {
  if (t1 > T) t1 = T;
  if (t0 < 0) t0 = 0;
  if (t1 <= t0) return;
  size_t first, last;   // elements
  if (tiled) {
    const int t8 = (T + 7) >> 3;
    first = (static_cast<size_t>(b) * t8 + (t0 >> 3)) * C * 8;
    last = (static_cast<size_t>(b) * t8 + ((t1 + 7) >> 3)) * C * 8;
  } else {
    first = (static_cast<size_t>(b) * T + t0) * C;
    last = (static_cast<size_t>(b) * T + t1) * C;
  }
  const char* base = reinterpret_cast<const char*>(sum + first);
  size_t bytes = (last - first) * 2;
  while (bytes) {   // (C * 2 is a multiple of 64 bytes: every piece stays 16-byte aligned)
    const uint32_t n = bytes > 65536 ? 65536u : static_cast<uint32_t>(bytes);
    bulk_prefetch_l2(base, n);
    base += n;
    bytes -= n;
  }
}

struct EpiOut {
  const __nv_bfloat16* sum_a;  // bf16 running resblock sum to add (nullptr = absent)
  float* out_f32;
  __nv_bfloat16* out_act;
  float slope, scale, inv;  // scale: 0 = none, else result *= scale (1 / num_kernels)
  int act_tanh;             // out_act = bf16(tanh(result)) instead of leaky_relu (Postnet, N2)
  int f16;                  // 16-bit tensors (residual, running sum, out_act) are fp16 instead of bf16 (ptx.cuh pack16)
};

template <bool F16>
__device__ __forceinline__ void add_bf16x16(float (&f)[16], const uint4 (&q)[2]) {
  const uint32_t w[8] = {q[0].x, q[0].y, q[0].z, q[0].w, q[1].x, q[1].y, q[1].z, q[1].w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float lo, hi;
    unpack16t<F16>(w[j], lo, hi);
    f[2 * j] += lo;
    f[2 * j + 1] += hi;
  }
}

// 16-byte shared-memory accesses by shared-window address
__device__ __forceinline__ void st_shared_u4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// named barrier over `nthreads` threads (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// acc + bias + residual (+ partial sums) (* scale) -> leaky_relu -> 16 bf16 (8 packed words), for one 16-column item:
// the arithmetic of epi_finish16 without the stores (the fused pair kernel stages the result in shared memory and
// writes it to global memory with coalesced accesses).
template <bool F16>
__device__ __forceinline__ void epi_compute16_t(const uint32_t (&v)[16], const float4 (&bv)[4], const uint4 (&rq)[2],
                                                const uint4 (&sa)[2], const EpiOut& o, uint32_t (&pk)[8]) {
  float f[16];
  const uint32_t w[8] = {rq[0].x, rq[0].y, rq[0].z, rq[0].w, rq[1].x, rq[1].y, rq[1].z, rq[1].w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float lo, hi;
    unpack16t<F16>(w[j], lo, hi);
    f[2 * j] = fminf(lo, lo * o.inv);
    f[2 * j + 1] = fminf(hi, hi * o.inv);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[4 * i] += __uint_as_float(v[4 * i]) + bv[i].x;
    f[4 * i + 1] += __uint_as_float(v[4 * i + 1]) + bv[i].y;
    f[4 * i + 2] += __uint_as_float(v[4 * i + 2]) + bv[i].z;
    f[4 * i + 3] += __uint_as_float(v[4 * i + 3]) + bv[i].w;
  }
  if (o.sum_a) add_bf16x16<F16>(f, sa);
  if (o.scale != 0.f) {
    const float sc = o.scale;
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] *= sc;
  }
  const float s = o.slope;
  if (o.act_tanh) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      pk[i] = pack16t<F16>(tanhf(f[2 * i]), tanhf(f[2 * i + 1]));
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float a = f[2 * i], c = f[2 * i + 1];
    pk[i] = pack16t<F16>(fmaxf(a, a * s), fmaxf(c, c * s));  // leaky_relu, 0 < s <= 1
  }
}

__device__ __forceinline__ void epi_compute16(const uint32_t (&v)[16], const float4 (&bv)[4], const uint4 (&rq)[2],
                                              const uint4 (&sa)[2], const EpiOut& o, uint32_t (&pk)[8]) {
  if (o.f16) epi_compute16_t<true>(v, bv, rq, sa, o, pk);
  else epi_compute16_t<false>(v, bv, rq, sa, o, pk);
}

// acc + bias + residual (+ partial sums) (* scale) -> out_f32 and/or leaky_relu -> out_act, for one 16-column item.
//   v        raw accumulator bits
//   bv       bias of the 16 columns
//   rq       residual: 16 bf16 (zeros when there is none)
//   sa       running sum: 16 bf16 (only read when o.sum_a != nullptr)
//   off      element offset of (row, first column) in the [B][T][n_total] outputs
template <bool F16>
__device__ __forceinline__ void epi_finish16_t(const uint32_t (&v)[16], const float4 (&bv)[4], const uint4 (&rq)[2],
                                               const uint4 (&sa)[2], const EpiOut& o, size_t off, bool valid) {
  if (!valid) return;
  float f[16];
  const uint32_t w[8] = {rq[0].x, rq[0].y, rq[0].z, rq[0].w, rq[1].x, rq[1].y, rq[1].z, rq[1].w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    // bf16 -> fp32 is a 16-bit shift; x = min(a, a/slope) inverts leaky_relu for 0 < slope < 1
    float lo, hi;
    unpack16t<F16>(w[j], lo, hi);
    f[2 * j] = fminf(lo, lo * o.inv);
    f[2 * j + 1] = fminf(hi, hi * o.inv);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[4 * i] += __uint_as_float(v[4 * i]) + bv[i].x;
    f[4 * i + 1] += __uint_as_float(v[4 * i + 1]) + bv[i].y;
    f[4 * i + 2] += __uint_as_float(v[4 * i + 2]) + bv[i].z;
    f[4 * i + 3] += __uint_as_float(v[4 * i + 3]) + bv[i].w;
  }
  if (o.sum_a) add_bf16x16<F16>(f, sa);
  if (o.scale != 0.f) {
    const float sc = o.scale;
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] *= sc;
  }
  if (o.out_f32) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
      st_global_256(o.out_f32 + off + 8 * i,
                    make_uint4(__float_as_uint(f[8 * i]), __float_as_uint(f[8 * i + 1]), __float_as_uint(f[8 * i + 2]),
                               __float_as_uint(f[8 * i + 3])),
                    make_uint4(__float_as_uint(f[8 * i + 4]), __float_as_uint(f[8 * i + 5]),
                               __float_as_uint(f[8 * i + 6]), __float_as_uint(f[8 * i + 7])));
  }
  if (o.out_act) {
    const float s = o.slope;
    uint32_t pk[8];
    if (o.act_tanh) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        pk[i] = pack16t<F16>(tanhf(f[2 * i]), tanhf(f[2 * i + 1]));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float a = f[2 * i], c = f[2 * i + 1];
        pk[i] = pack16t<F16>(fmaxf(a, a * s), fmaxf(c, c * s));  // leaky_relu, 0 < s <= 1
      }
    }
    st_global_256(o.out_act + off, make_uint4(pk[0], pk[1], pk[2], pk[3]), make_uint4(pk[4], pk[5], pk[6], pk[7]));
  }
}

__device__ __forceinline__ void epi_finish16(const uint32_t (&v)[16], const float4 (&bv)[4], const uint4 (&rq)[2],
                                             const uint4 (&sa)[2], const EpiOut& o, size_t off, bool valid) {
  if (o.f16) epi_finish16_t<true>(v, bv, rq, sa, o, off, valid);
  else epi_finish16_t<false>(v, bv, rq, sa, o, off, valid);
}

}  // namespace e2e
