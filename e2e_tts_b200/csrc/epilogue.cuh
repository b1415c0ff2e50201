// Epilogue pieces shared by conv_tc.cuh and pair_tc.cuh.  An epilogue "item" is 16 consecutive output columns of
// one accumulator row: the thread that owns TMEM lane r reads them with one tcgen05.ld.32x32b.x16, and all global
// traffic of the item is 32-byte aligned 256-bit accesses (one full sector per lane and instruction):
//   residual  16 bf16 = 32 B   (leaky_relu(x) of the pair's input; x recovered by the inverse LeakyReLU)
//   sum_in    16 fp32 = 64 B   (running resblock sum, generator.py:44-47)
//   out_f32   16 fp32 = 64 B
//   out_act   16 bf16 = 32 B   (leaky_relu(result, slope))
#pragma once
#include "ptx.cuh"

namespace e2e {

constexpr int kEpiWarps = 16;                      // four warps per TMEM lane quarter
constexpr int kConvThreads = (4 + kEpiWarps) * 32;  // + TMA producer, weight producer, MMA issuer, TMEM allocator

// TMEM -> registers: 16 consecutive fp32 columns of this thread's lane.
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

struct EpiOut {
  const float* bias;    // + column offset already applied by the caller? no: base pointer, column added here
  const float* sum_in;  // base pointers (nullptr = absent)
  float* out_f32;
  __nv_bfloat16* out_act;
  float slope, divisor, inv;
};

// acc (+ bias) + residual (+ sum) (/ divisor) -> out_f32 and/or leaky_relu -> out_act, for one 16-column item.
//   v        raw accumulator bits
//   rq       residual: 16 bf16 (zeros when there is none)
//   sq       running sum: 16 fp32 (only read when o.sum_in != nullptr)
//   n0       first output column of the item (bias index)
//   off      element offset of (row, n0) in the [B][T][n_total] outputs
__device__ __forceinline__ void epi_finish16(const uint32_t (&v)[16], const uint4 (&rq)[2], const uint4 (&sq)[4],
                                             const EpiOut& o, int n0, size_t off, bool valid) {
  float f[16];
  const uint32_t w[8] = {rq[0].x, rq[0].y, rq[0].z, rq[0].w, rq[1].x, rq[1].y, rq[1].z, rq[1].w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    // bf16 -> fp32 is a 16-bit shift; x = min(a, a/slope) inverts leaky_relu for 0 < slope < 1
    float lo = __uint_as_float(w[j] << 16), hi = __uint_as_float(w[j] & 0xffff0000u);
    f[2 * j] = fminf(lo, lo * o.inv);
    f[2 * j + 1] = fminf(hi, hi * o.inv);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 bv = __ldg(reinterpret_cast<const float4*>(o.bias + n0) + i);
    f[4 * i] += __uint_as_float(v[4 * i]) + bv.x;
    f[4 * i + 1] += __uint_as_float(v[4 * i + 1]) + bv.y;
    f[4 * i + 2] += __uint_as_float(v[4 * i + 2]) + bv.z;
    f[4 * i + 3] += __uint_as_float(v[4 * i + 3]) + bv.w;
  }
  if (!valid) return;
  if (o.sum_in) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[4 * i] += __uint_as_float(sq[i].x);
      f[4 * i + 1] += __uint_as_float(sq[i].y);
      f[4 * i + 2] += __uint_as_float(sq[i].z);
      f[4 * i + 3] += __uint_as_float(sq[i].w);
    }
  }
  if (o.divisor != 0.f) {
    const float dv = o.divisor;
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = f[i] / dv;  // true division, like `xs / self.num_kernels`
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
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float a = f[2 * i], c = f[2 * i + 1];
      __nv_bfloat162 h = __floats2bfloat162_rn(fmaxf(a, a * s), fmaxf(c, c * s));  // leaky_relu, 0 < s < 1
      pk[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    st_global_256(o.out_act + off, make_uint4(pk[0], pk[1], pk[2], pk[3]), make_uint4(pk[4], pk[5], pk[6], pk[7]));
  }
}

}  // namespace e2e
