// STFT -> log-mel front-end in ONE pass over the audio (reference: e2e_tts/src/tools/stft.py:46-89,
// TorchSTFT.mel_spectrogram; :107-135 generate_melspecs; e2e_tts/src/tools/utils.py:22-28).
//
//   reflect-pad 384 | frame t = xp[256 t .. +1024) * periodic Hann | 1024-pt real FFT | sqrt(re^2+im^2+1e-9)
//   | sparse Slaney mel filterbank | log(max(., 1e-5)) | energy = sqrt(sum_k mag^2)
//
// The path is bandwidth-shaped (4 B in, 0.32 B out per sample) but a direct DFT would be ~2 MFLOP per frame, so the
// transform is FFT-structured on chip.  A CTA stages the audio of 32 consecutive frames in shared memory once (frames
// overlap 4x); each group of 64 threads runs the 512-point complex FFT of an even/odd-packed frame as three radix-8
// passes held in registers (8 complex values per thread, packed f32x2 arithmetic), exchanging through one padded,
// conflict-free shared-memory buffer after passes 1 and 2.  What bounds the kernel is instruction issue (~700 instructions
// per thread and frame in round 1: profiles/r02_experiments_notes.md §7), so this version spends its changes on instruction count:
//   * complex butterflies and twiddle multiplies in Blackwell's packed FP32 instructions (FADD2 / FMUL2 / FFMA2);
//   * the real-FFT recombination pairs bin k with 512 - k.  With k = k1 + 8 c + 64 d held by thread (k1, c) in register d,
//     512 - k lives in ONE other thread, (8 - k1, 7 - c) [(0, 8 - c) for k1 = 0], in register 7 - d: the k1 -> thread
//     map is permuted so that both are in the same warp, and eight shuffles bring the partner's values;
//   * the magnitudes of all 32 frames stay in shared memory ([frame][bin], odd pitch) and the mel filterbank runs once
//     per CTA with frames on the lanes: for filter m a warp walks the filter's bins, the weight is warp-uniform (one
//     broadcast read) and lane f reads mag[f][k] conflict-free; the log-mel leaves as a coalesced 128-byte row store.
// The frame energy sqrt(sum_k mag_k^2) over all 513 bins comes from Parseval's identity on the windowed samples:
// sum_{k<=512} |X_k|^2 = (1024 sum_n xw_n^2 + X_0^2 + X_512^2) / 2 (shuffle-reduced).  The index maps are emulated and
// checked on the CPU in tests/test_mel_fft_plan.py.
#include <cmath>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/e2e_tts_b200.h"
#include "errors.h"

using namespace e2e;

namespace {

constexpr int kNfft = 1024;
constexpr int kHop = 256;
constexpr int kPad = (kNfft - kHop) / 2;  // 384, stft.py:33
constexpr int kBins = kNfft / 2 + 1;      // 513
constexpr int kF = 32;                    // frames per CTA (= lanes of the filterbank pass)
constexpr int kGroups = 4;                // 64-thread FFT groups per CTA
constexpr int kThreads = kGroups * 64;
constexpr int kAudio = (kF - 1) * kHop + kNfft;  // samples staged per CTA
constexpr int kSx = 576;                  // padded complex exchange buffer (8 rows of 72)
constexpr int kMaxMels = 128;

struct MelParams {
  const float* wav;
  long long ldw, L;
  int B, T, n_mels;
  int nb;               // bins the filterbank reads: 1 + last non-zero column of the basis (<= 513)
  int magp;             // row pitch of the magnitude tile: 4 x odd >= nb (16-byte rows, conflict-free 128-bit lane reads)
  int n_w;              // packed filterbank weights
  float* mel;
  float* energy;
  int* range_flag;
  const float* window;  // [1024] periodic Hann
  const float2* tw;     // [1024] exp(-2 pi i j / 1024)
  const float* fb_w;    // [n_w] filter m: weights of bins lo[m] .. lo[m] + n[m] - 1 at off[m]; lo, n, off multiples of 4
  const int* fb_meta;   // [3][n_mels]: lo, n, off
};

__device__ __forceinline__ float sqrt_approx(float x) {  // one MUFU.SQRT: max relative error 2^-23 (PTX ISA); the
  float y;                                               // arguments are >= 1e-9, so flushing subnormals changes nothing
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Complex values travel as packed f32x2 register pairs (lo = re, hi = im) and the butterflies use Blackwell's packed
// FP32 instructions (FADD2 / FMUL2 / FFMA2: one instruction per complex add, two per complex multiply) - the kernel
// is bound by instruction issue, not by memory (profiles/r02_experiments_notes.md §7).
typedef unsigned long long c64;
__device__ __forceinline__ c64 pk2(float lo, float hi) {
  c64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void up2(c64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ c64 add2(c64 a, c64 b) {
  c64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ c64 sub2(c64 a, c64 b) {
  c64 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ c64 mul2(c64 a, c64 b) {
  c64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ c64 fma2(c64 a, c64 b, c64 c) {
  c64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ c64 swap2(c64 a) {   // (re, im) -> (im, re): folds into the consumer's operand swizzle
  float lo, hi;
  up2(a, lo, hi);
  return pk2(hi, lo);
}
// a * w for w = (wx, wy):  wx * a + (-wy, wy) * (a.im, a.re)
__device__ __forceinline__ c64 cmul2(c64 a, float wx, float wy) {
  return fma2(pk2(wx, wx), a, mul2(swap2(a), pk2(-wy, wy)));
}
// d + (-i) e = (d.re + e.im, d.im - e.re)   and   d - (-i) e
__device__ __forceinline__ c64 add_negi(c64 d, c64 e) { return fma2(swap2(e), pk2(1.f, -1.f), d); }
__device__ __forceinline__ c64 sub_negi(c64 d, c64 e) { return fma2(swap2(e), pk2(-1.f, 1.f), d); }

// 4-point DFT of (c0, c1, c2', c3) where c2' = -i * c2 when C2_NEGI (the W8^2 twiddle of dft8 folded into the adds)
template <bool C2_NEGI>
__device__ __forceinline__ void dft4(c64 c0, c64 c1, c64 c2, c64 c3, c64& y0, c64& y1, c64& y2, c64& y3) {
  const c64 d0 = C2_NEGI ? add_negi(c0, c2) : add2(c0, c2);
  const c64 d2 = C2_NEGI ? sub_negi(c0, c2) : sub2(c0, c2);
  const c64 d1 = add2(c1, c3), e = sub2(c1, c3);
  y0 = add2(d0, d1);
  y2 = sub2(d0, d1);
  y1 = add_negi(d2, e);
  y3 = sub_negi(d2, e);
}

// Radix-2 decimation-in-frequency 8-point DFT, natural-order output (28 packed instructions).
__device__ __forceinline__ void dft8(c64 (&a)[8]) {
  const float r = 0.70710678118654752440f;
  const c64 b0 = add2(a[0], a[4]), b1 = add2(a[1], a[5]), b2 = add2(a[2], a[6]), b3 = add2(a[3], a[7]);
  const c64 b4 = sub2(a[0], a[4]), b6 = sub2(a[2], a[6]);
  c64 b5 = sub2(a[1], a[5]), b7 = sub2(a[3], a[7]);
  b5 = cmul2(b5, r, -r);     // * W8^1
  b7 = cmul2(b7, -r, -r);    // * W8^3      (b6 * W8^2 = -i b6 is folded into dft4)
  dft4<false>(b0, b1, b2, b3, a[0], a[2], a[4], a[6]);
  dft4<true>(b4, b5, b6, b7, a[1], a[3], a[5], a[7]);
}

__device__ __forceinline__ void group_sync(int g) {
  asm volatile("bar.sync %0, 64;" ::"r"(g + 1) : "memory");
}

// k1 handled by the thread with (thread >> 3) == h: partners k1 <-> 8 - k1 share a warp (h < 4: {0, 1, 4, 7}; h >= 4:
// {2, 3, 6, 5}), and the two rows a half-warp reads in pass 2 differ by an odd number of rows (bank-conflict-free with
// the 72-element row pitch).
__device__ __forceinline__ int k1_of(int h) { return (0x56327410 >> (4 * h)) & 7; }   // {0,1,4,7,2,3,6,5}
__device__ __forceinline__ int h_of(int k1) { return (0x36725410 >> (4 * k1)) & 7; }  // inverse: k1 0..7 -> h

// DMAX = registers d of the last pass that hold bins the filterbank reads (bin k = k1 + 8 c + 64 d): 6 when the bank ends
// below bin 384 (the e2e-tts bank: fmax = 8 kHz -> bin 371), else 8.  A compile-time bound: a run-time `continue` costs
// more than it saves (it splits the recombination into branchy blocks and the shuffles no longer overlap; measured
// 1.68 -> 1.73 ms with the run-time bound, 1.68 -> 1.58 ms with this one, profiles/r02_experiments_notes.md §14).
template <int DMAX>
__global__ void __launch_bounds__(kThreads, 2) mel_kernel(const MelParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* audio = reinterpret_cast<float*>(smem);                    // [kAudio]
  float2* sx = reinterpret_cast<float2*>(audio + kAudio);           // [kGroups][kSx]
  float* mag_all = reinterpret_cast<float*>(sx + kGroups * kSx);    // [kF][magp]
  float* red_s = mag_all + kF * p.magp;                             // [kF][2] per-warp halves of a frame's energy sum
  float* fbw_s = red_s + kF * 2;                                    // [n_w]
  int* meta_s = reinterpret_cast<int*>(fbw_s + ((p.n_w + 3) & ~3)); // [3][n_mels]

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * kF;
  const int nf = min(kF, p.T - f0);

  // ---- stage audio (reflect padding resolved here, stft.py:60-64) and the sparse filterbank ----
  {
    const float* row = p.wav + (long long)b * p.ldw;
    const long long base = (long long)f0 * kHop - kPad;  // source index of audio[0]
    const int need = (nf - 1) * kHop + kNfft;
    bool bad = false;  // any sample outside [-1, 1] (NaN included), stft.py:56-57
    if (base >= 0 && base + need <= p.L && ((reinterpret_cast<uintptr_t>(row + base) & 15) == 0)) {
      // interior CTA, 16-byte aligned: all of a thread's 128-bit loads are issued before the first use
      const float4* src4 = reinterpret_cast<const float4*>(row + base);
      float4* dst4 = reinterpret_cast<float4*>(audio);
      const int n4 = need >> 2;  // need is a multiple of 256
      constexpr int kIt = (kAudio / 4 + kThreads - 1) / kThreads;
      float4 v[kIt];
#pragma unroll
      for (int it = 0; it < kIt; ++it) {
        const int i = tid + it * kThreads;
        v[it] = i < n4 ? __ldg(src4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int it = 0; it < kIt; ++it) {
        const int i = tid + it * kThreads;
        bad |= !(fabsf(v[it].x) <= 1.0f) | !(fabsf(v[it].y) <= 1.0f) | !(fabsf(v[it].z) <= 1.0f) |
               !(fabsf(v[it].w) <= 1.0f);
        if (i < n4) dst4[i] = v[it];
      }
    } else {
      for (int i = tid; i < need; i += kThreads) {
        long long s = base + i;
        if (s < 0) s = -s;
        if (s >= p.L) s = 2 * (p.L - 1) - s;
        const float v = row[s];
        bad |= !(fabsf(v) <= 1.0f);
        audio[i] = v;
      }
    }
    if (bad && p.range_flag) atomicOr(p.range_flag, 1);
    for (int i = tid; i < p.n_w; i += kThreads) fbw_s[i] = p.fb_w[i];
    for (int i = tid; i < 3 * p.n_mels; i += kThreads) meta_s[i] = p.fb_meta[i];
    if (nf < kF)   // frames past the clip end: the filterbank pass reads their (unwritten) rows with lanes >= nf
      for (int i = tid; i < (kF - nf) * p.magp; i += kThreads) mag_all[nf * p.magp + i] = 0.f;
    // columns nb .. magp-1 pad the rows to whole float4s: they meet zero weights, but must be finite
    for (int i = tid; i < kF * (p.magp - p.nb); i += kThreads)
      mag_all[(i / (p.magp - p.nb)) * p.magp + p.nb + i % (p.magp - p.nb)] = 0.f;
  }

  const int g = tid >> 6;   // FFT group
  const int t = tid & 63;   // thread within the group
  const int hi = t >> 3, lo = t & 7;
  const int k1 = k1_of(hi);   // passes 2 / 3 and the recombination: this thread's k1
  float2* S1 = sx + g * kSx;
  const int nb = p.nb;

  // per-thread constants: window taps, the twiddles of passes 1 and 2, the recombination twiddle and partner lane
  c64 wpair[8];
  float2 t1[8], t2[8];
#pragma unroll
  for (int n1 = 0; n1 < 8; ++n1) {
    wpair[n1] = pk2(p.window[128 * n1 + 2 * t], p.window[128 * n1 + 2 * t + 1]);
    t1[n1] = p.tw[2 * ((t * n1) & 511)];    // W_512^(n2 k1), n2 = t
    t2[n1] = p.tw[16 * ((lo * n1) & 63)];   // W_64^(b c),   b = lo
  }
  const bool self = k1 == 0 && lo == 0;      // holds k = 64 d: its partner bins 64 (8 - d) are its own registers
  const int k1p = (8 - k1) & 7, cp = k1 ? 7 - lo : (8 - lo) & 7;
  const int lane_p = ((h_of(k1p) & 3) << 3) | cp;   // lane (within this warp) of the thread that holds 512 - k
  // X[k] = (s + wk * dd / i) / 2 with s = Z[k] + conj Z[512-k], dd = Z[k] - conj Z[512-k], wk = W_1024^k =
  // W_1024^(k1 + 8 c) * W_16^d:  gbase = -i/2 * W_1024^(k1 + 8 c)
  const float2 wb = p.tw[k1 + 8 * lo];
  const c64 gbase = pk2(0.5f * wb.y, -0.5f * wb.x);
  c64* S1c = reinterpret_cast<c64*>(S1);
  __syncthreads();

  for (int fl = g; fl < nf; fl += kGroups) {
    c64 a[8];
    // pass 1: thread n2 = t, points z[64 n1 + n2] = (xw[128 n1 + 2 n2], xw[128 n1 + 2 n2 + 1])
    const c64* src = reinterpret_cast<const c64*>(audio + fl * kHop) + t;
    c64 e2 = pk2(0.f, 0.f);  // (sum of the squared windowed even samples, ... odd samples) of this thread
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) {
      a[n1] = mul2(src[64 * n1], wpair[n1]);
      e2 = fma2(a[n1], a[n1], e2);
    }
    dft8(a);
    S1c[t] = a[0];
#pragma unroll
    for (int q = 1; q < 8; ++q) S1c[q * 72 + t] = cmul2(a[q], t1[q].x, t1[q].y);
    group_sync(g);
    // pass 2: thread (k1, b = lo) reads Y[k1][8 a + b]
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = S1c[k1 * 72 + 8 * q + lo];
    dft8(a);
    group_sync(g);   // every thread of the group has read its pass-1 values: the buffer takes the pass-2 results
    S1c[hi * 72 + lo] = a[0];
#pragma unroll
    for (int c = 1; c < 8; ++c) S1c[hi * 72 + c * 9 + lo] = cmul2(a[c], t2[c].x, t2[c].y);
    group_sync(g);
    // pass 3: thread (k1, c = lo) reads U[k1][c][b], b = 0..7 (written by the 8 threads that share k1)
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = S1c[hi * 72 + lo * 9 + q];
    dft8(a);          // a[d] = Z[k1 + 8 c + 64 d]
    // real-FFT recombination.  The partner thread sends its register dd, which is Z[512 - k] for this thread's
    // d = 7 - dd; the thread that holds k = 64 d pairs its own registers d and 8 - d instead.
    float x0sq = 0.f, xnsq = 0.f;
    float* mrow = mag_all + fl * p.magp;
#pragma unroll
    for (int dd = 0; dd < 8; ++dd) {
      if (7 - dd >= DMAX) continue;   // compile-time: nobody reads these bins (nor needs the partner values for them)
      float ax, ay;
      up2(a[dd], ax, ay);
      float bx = __shfl_sync(0xffffffffu, ax, lane_p);
      float by = __shfl_sync(0xffffffffu, ay, lane_p);
      const int d = 7 - dd;
      if (self) up2(a[(8 - d) & 7], bx, by);
      const c64 zb = pk2(bx, by);
      const c64 sm = fma2(zb, pk2(1.f, -1.f), a[d]);    // Z[k] + conj Z[512-k]
      const c64 df = fma2(zb, pk2(-1.f, 1.f), a[d]);    // Z[k] - conj Z[512-k]
      // g = gbase * W_16^d
      constexpr float kC[8] = {1.f, 0.92387953251128675613f, 0.70710678118654752440f, 0.38268343236508977173f,
                               0.f, -0.38268343236508977173f, -0.70710678118654752440f, -0.92387953251128675613f};
      constexpr float kS[8] = {0.f, -0.38268343236508977173f, -0.70710678118654752440f, -0.92387953251128675613f,
                               -1.f, -0.92387953251128675613f, -0.70710678118654752440f, -0.38268343236508977173f};
      const c64 gd = cmul2(gbase, kC[d], kS[d]);
      float gx, gy;
      up2(gd, gx, gy);
      const c64 wz = cmul2(df, gx, gy);                  // wk * Zo
      const c64 x = fma2(pk2(0.5f, 0.5f), sm, wz);       // X[k]
      float xr, xi;
      up2(x, xr, xi);
      const int k = k1 + 8 * lo + 64 * d;
      if (k < nb) mrow[k] = sqrt_approx(fmaf(xr, xr, fmaf(xi, xi, 1e-9f)));   // stft.py:77
      if (d == 0 && self) {
        const c64 xp = fma2(pk2(0.5f, 0.5f), sm, mul2(wz, pk2(-1.f, -1.f)));   // X[512] = Ze - W^0 Zo (real)
        float pr, pi_;
        up2(xp, pr, pi_);
        x0sq = xr * xr;
        xnsq = pr * pr;
        if (512 < nb) mrow[512] = sqrt_approx(xnsq + 1e-9f);
      }
    }
    // energy^2 = sum_{k=0..512} (|X_k|^2 + 1e-9) = (1024 sum xw^2 + X_0^2 + X_512^2) / 2 + 513e-9   (stft.py:84)
    float e, eo;
    up2(e2, e, eo);
    e = fmaf(1024.f, e + eo, x0sq + xnsq);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((t & 31) == 0) red_s[fl * 2 + (t >> 5)] = e;   // the two warps' halves of the frame: combined after the loop
    group_sync(g);   // every thread has read its pass-3 values before the next frame's pass 1 overwrites the buffer
  }
  __syncthreads();

  // ---- sparse mel filterbank (stft.py:80) + log compression (stft.py:81, utils.py:28), frames on the lanes ----
  {
    const int warp = tid >> 5, lane = tid & 31;
    const float* mf = mag_all + lane * p.magp;
    const int* m_lo = meta_s;
    const int* m_n = meta_s + p.n_mels;
    const int* m_off = meta_s + 2 * p.n_mels;
    for (int m = warp; m < p.n_mels; m += kThreads / 32) {
      // the filter's span, widened to whole float4s of bins (zero weights in the padding): one broadcast 128-bit read
      // of four weights and one conflict-free 128-bit read of this lane's four magnitudes per four bins
      const float4* w4 = reinterpret_cast<const float4*>(fbw_s + m_off[m]);
      const float4* v4 = reinterpret_cast<const float4*>(mf + m_lo[m]);
      const int n4 = m_n[m] >> 2;
      float acc = 0.f;
#pragma unroll 4
      for (int i = 0; i < n4; ++i) {
        const float4 w = w4[i], v = v4[i];
        acc = fmaf(w.x, v.x, acc);
        acc = fmaf(w.y, v.y, acc);
        acc = fmaf(w.z, v.z, acc);
        acc = fmaf(w.w, v.w, acc);
      }
      // __logf = lg2.approx * ln 2: absolute error below 4e-6 over the mel range [1e-5, 1e3] (3 ulp of the result outside
      // [0.5, 2], 2^-21.4 inside), against the 1e-5 L1 bound of the parity tests; logf costs ~10x the instructions
      if (lane < nf) p.mel[((long long)b * p.n_mels + m) * p.T + f0 + lane] = __logf(fmaxf(acc, 1e-5f));
    }
    // (one thread per frame here instead of a divergent tail behind every frame's last barrier)
    if (p.energy && tid < nf)
      p.energy[(long long)b * p.T + f0 + tid] = sqrtf(0.5f * (red_s[2 * tid] + red_s[2 * tid + 1]) + 513e-9f);
  }
}

}  // namespace

struct e2e_mel {
  int n_fft, hop, win, n_mels;
  int nb = 0, magp = 0, n_w = 0;
  float* d_window = nullptr;
  float2* d_tw = nullptr;
  float* d_fbw = nullptr;
  int* d_meta = nullptr;
  int smem_bytes = 0;
};

extern "C" int e2e_mel_create(int32_t n_fft, int32_t hop_length, int32_t win_length, int32_t n_mels,
                              const float* mel_basis, e2e_mel** out) {
  if (!mel_basis || !out) return fail(-1, "null argument");
  if (n_fft != kNfft || win_length != kNfft || hop_length != kHop)
    return fail(-4, "mel front-end supports n_fft == win_length == 1024 and hop_length == 256 (the e2e-tts config)");
  if (n_mels < 1 || n_mels > kMaxMels) return fail(-4, "n_mels must be in [1, 128]");
  // sparse filterbank: per filter the span [lo, hi] of its non-zero bins (zeros inside the span are kept as weights)
  std::vector<int> meta(3 * (size_t)n_mels, 0);
  std::vector<float> wpk;
  int nb = 1;
  for (int r = 0; r < n_mels; ++r) {
    int lo = -1, hi = -1;
    for (int k = 0; k < kBins; ++k)
      if (mel_basis[(size_t)r * kBins + k] != 0.f) {
        if (lo < 0) lo = k;
        hi = k;
      }
    const int lo4 = lo < 0 ? 0 : lo & ~3;                       // span widened to whole groups of four bins
    const int n4 = lo < 0 ? 0 : (hi + 1 - lo4 + 3) / 4 * 4;
    meta[r] = lo4;
    meta[n_mels + r] = n4;
    meta[2 * n_mels + r] = (int)wpk.size();
    for (int k = lo4; k < lo4 + n4; ++k) wpk.push_back(k < kBins ? mel_basis[(size_t)r * kBins + k] : 0.f);
    if (hi + 1 > nb) nb = hi + 1;
  }
  if (wpk.empty()) wpk.push_back(0.f);
  e2e_mel* m = new e2e_mel;
  m->n_fft = n_fft;
  m->hop = hop_length;
  m->win = win_length;
  m->n_mels = n_mels;
  m->nb = nb;
  m->magp = (((nb + 3) / 4) | 1) * 4;
  m->n_w = (int)wpk.size();
  std::vector<float> window(kNfft);
  std::vector<float2> tw(kNfft);
  const double pi = 3.14159265358979323846;
  for (int n = 0; n < kNfft; ++n) {
    window[n] = (float)(0.5 - 0.5 * cos(2.0 * pi * n / kNfft));  // periodic Hann, stft.py:44
    tw[n] = make_float2((float)cos(2.0 * pi * n / kNfft), (float)(-sin(2.0 * pi * n / kNfft)));
  }
  // mirrors the carve-up at the top of mel_kernel
  m->smem_bytes = (kAudio + 2 * kGroups * kSx + kF * m->magp + kF * 2 + ((m->n_w + 3) & ~3) + 3 * n_mels) * 4;
  if (m->smem_bytes > 227 * 1024) {
    delete m;
    return fail(-4, "mel filterbank too dense for the shared-memory budget");
  }
  cudaError_t e = cudaSuccess;
  auto up = [&](void** dst, const void* src, size_t bytes) {
    if (e != cudaSuccess) return;
    e = cudaMalloc(dst, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
  };
  up((void**)&m->d_window, window.data(), window.size() * 4);
  up((void**)&m->d_tw, tw.data(), tw.size() * 8);
  up((void**)&m->d_fbw, wpk.data(), wpk.size() * 4);
  up((void**)&m->d_meta, meta.data(), meta.size() * 4);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(mel_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, m->smem_bytes);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(mel_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, m->smem_bytes);
  if (e != cudaSuccess) {
    e2e_mel_destroy(m);
    return fail((int)e, std::string("e2e_mel_create: ") + cudaGetErrorString(e));
  }
  *out = m;
  return 0;
}

extern "C" void e2e_mel_destroy(e2e_mel* m) {
  if (!m) return;
  cudaFree(m->d_window);
  cudaFree(m->d_tw);
  cudaFree(m->d_fbw);
  cudaFree(m->d_meta);
  delete m;
}

extern "C" int64_t e2e_mel_num_frames(const e2e_mel* m, int64_t L) {
  if (!m || L <= kPad) return 0;  // reflect padding needs L > pad (torch raises for shorter inputs)
  const int64_t padded = L + 2 * kPad;
  return padded < kNfft ? 0 : 1 + (padded - kNfft) / kHop;
}

extern "C" int e2e_mel_forward(e2e_mel* m, const float* wav, int32_t B, int64_t L, int64_t ldw, float* mel,
                               float* energy, int32_t* range_flag, void* stream) {
  if (!m || !wav || !mel) return fail(-1, "null argument");
  if (B < 1 || B > 65535) return fail(-1, "B must be in [1, 65535]");
  if (L <= kPad) return fail(-1, "input shorter than the reflect padding (need L > 384)");
  if (ldw < L) return fail(-1, "ldw < L");
  const int64_t T = e2e_mel_num_frames(m, L);
  if (T < 1 || T > 0x7fffffff) return fail(-1, "bad frame count");
  MelParams p;
  p.wav = wav;
  p.ldw = ldw;
  p.L = L;
  p.B = B;
  p.T = (int)T;
  p.n_mels = m->n_mels;
  p.nb = m->nb;
  p.magp = m->magp;
  p.n_w = m->n_w;
  p.mel = mel;
  p.energy = energy;
  p.range_flag = range_flag;
  p.window = m->d_window;
  p.tw = m->d_tw;
  p.fb_w = m->d_fbw;
  p.fb_meta = m->d_meta;
  dim3 grid((unsigned)((T + kF - 1) / kF), (unsigned)B);
  if (m->nb <= 6 * 64)
    mel_kernel<6><<<grid, kThreads, m->smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  else
    mel_kernel<8><<<grid, kThreads, m->smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, std::string("mel_kernel launch: ") + cudaGetErrorString(e));
  return 0;
}
