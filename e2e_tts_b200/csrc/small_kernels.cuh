// CUDA-core kernels at the two ends of the generator, where the work is a layout change or a dot product
// rather than a GEMM:
//   * mel_to_act:    [B,80,T] fp32 (any strides; the reference passes a transposed view, utils.py:144)
//                    -> channels-last bf16 [B][T][cin_pad] operand of conv_pre (padding channels = 0)
//   * post_conv_tanh: conv_post (Conv1d C->1, k=7, pad 3) + tanh  (generator.py:50-51).  Its input is the
//                    previous epilogue's leaky_relu(x, 0.01) in bf16 (generator.py:49), N = 1 so this is a
//                    224-term dot product per sample, bandwidth-shaped.  Output: fp32 waveform (the reference's
//                    return value), or - PostOut::pcm - the int16 PCM its caller makes of it (combine_audio,
//                    src/api/utils.py:108-117: trim to mel_len * hop, * max_wav_value, astype int16), which halves
//                    the device->host and gather bytes.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "ptx.cuh"

namespace e2e {

__global__ void mel_to_act_kernel(const float* __restrict__ mel, long long sB, long long sC, long long sT, int B,
                                  int T, int C, int cpad, __nv_bfloat16* __restrict__ out, int f16,
                                  float slope = 1.0f) {   // slope < 1: store leaky_relu(x, slope) (standalone resblocks)
  const int G = cpad / 8;
  const long long total = (long long)B * T * G;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int t = (int)(idx % T);
  const int g = (int)((idx / T) % G);
  const int b = (int)(idx / ((long long)T * G));
  uint32_t pk[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c0 = g * 8 + 2 * i;
    float a = c0 < C ? mel[b * sB + c0 * sC + t * sT] : 0.f;
    float c = c0 + 1 < C ? mel[b * sB + (c0 + 1) * sC + t * sT] : 0.f;
    a = fmaxf(a, a * slope);
    c = fmaxf(c, c * slope);
    pk[i] = pack16(a, c, f16);
  }
  *reinterpret_cast<uint4*>(out + ((long long)b * T + t) * cpad + g * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

// Where post_conv_tanh writes.  pcm == nullptr: wav[b][t] = tanh(.) fp32.  Otherwise pcm[b][t] =
// (int16) trunc(tanh(.) * scale) for t < lens[b] * hop (all t when lens == nullptr) and 0 beyond: numpy's
// astype("int16") truncates toward zero; exactly +1.0 * 32768 saturates to 32767 instead of wrapping.
struct PostOut {
  float* wav;
  int16_t* pcm;
  const int32_t* lens;  // device [B] mel frames per utterance, or nullptr
  int hop;
  float scale;
};

__device__ __forceinline__ int16_t to_pcm16(float y, float scale) {
  const int v = __float2int_rz(y * scale);
  return (int16_t)max(-32768, min(32767, v));
}

constexpr int kPostMaxW = 7 * 64;
constexpr int kPostThreads = 96;
constexpr int kPostOpt = 5;  // output samples per thread (measured, 16 x 5 s: 2 -> 55 us, 3 -> 54 us, 5 -> 44 us, 7 -> 51 us)
constexpr int kPostTile = kPostThreads * kPostOpt;  // output samples per block

// w: [k][C] fp32 (tap-major), one output channel.  A block stages the bf16 rows [t0 - half, t0 + tile + half) in
// shared memory with coalesced 16-byte loads (rows outside the utterance = 0: conv_post's zero padding) and every
// thread produces OPT adjacent samples, so each staged row is read and unpacked once for up to OPT outputs (OPT + K - 1
// rows per thread instead of OPT * K).  With an odd row pitch (in 16-byte units) and an odd OPT the lanes of a quarter
// warp start in eight different 16-byte bank groups: the 128-bit row reads are conflict-free (OPT = 2 was 2-way).
// For every output the taps are accumulated in increasing order with the channels innermost - the result does not
// depend on OPT.
// The K*C weights arrive as a kernel parameter (constant bank): every thread of a warp reads the same weight at the
// same time, which the constant cache broadcasts without touching the shared-memory pipe the staged rows need.
template <int N>
struct PostWeights {
  float w[N];
};

template <int C, int K, int OPT, int THREADS>
__global__ void __launch_bounds__(THREADS)
post_conv_tanh_kernel(const __nv_bfloat16* __restrict__ act, const __grid_constant__ PostWeights<K * C> pw, float bias,
                      int B, int T, const PostOut out, int f16) {
  constexpr int HALF = (K - 1) / 2;
  constexpr int TILE = OPT * THREADS;
  constexpr int ROWS = TILE + 2 * HALF;
  constexpr int ROW16 = C / 8;       // uint4 per row
  constexpr int PITCH = ROW16 | 1;   // odd row pitch
  __shared__ uint4 rows[ROWS * PITCH];
  const float* sw = pw.w;
  const int t0 = blockIdx.x * TILE;
  const int b = blockIdx.y;
  const uint4* src = reinterpret_cast<const uint4*>(act + (long long)b * T * C);
  for (int i = threadIdx.x; i < ROWS * ROW16; i += THREADS) {
    const int r = i / ROW16;
    const int t = t0 - HALF + r;
    rows[r * PITCH + (i - r * ROW16)] =
        (t >= 0 && t < T) ? __ldg(src + (long long)t * ROW16 + (i - r * ROW16)) : make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  const int o = OPT * threadIdx.x;  // first of this thread's outputs, relative to t0
  float acc[OPT];
#pragma unroll
  for (int q = 0; q < OPT; ++q) acc[q] = bias;
#pragma unroll
  for (int r = 0; r < K + OPT - 1; ++r) {  // staged row o + r feeds output o + q with tap r - q
    const uint4* row = rows + (o + r) * PITCH;
#pragma unroll
    for (int c8 = 0; c8 < ROW16; ++c8) {
      const uint4 v = row[c8];
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float lo, hi;
        unpack16(u[i], lo, hi, f16);
#pragma unroll
        for (int q = 0; q < OPT; ++q) {
          const int j = r - q;
          if (j >= 0 && j < K) {
            acc[q] = fmaf(lo, sw[j * C + c8 * 8 + 2 * i], acc[q]);
            acc[q] = fmaf(hi, sw[j * C + c8 * 8 + 2 * i + 1], acc[q]);
          }
        }
      }
    }
  }
  const int t = t0 + o;
  if (out.pcm) {
    const long long valid = out.lens ? (long long)out.lens[b] * out.hop : (long long)T;
    int16_t* dst = out.pcm + (long long)b * T + t;
#pragma unroll
    for (int q = 0; q < OPT; ++q)
      if (t + q < T) dst[q] = t + q < valid ? to_pcm16(tanhf(acc[q]), out.scale) : (int16_t)0;
    return;
  }
  float* dst = out.wav + (long long)b * T + t;
#pragma unroll
  for (int q = 0; q < OPT; ++q)
    if (t + q < T) dst[q] = tanhf(acc[q]);
}

// Generic fallback (any C multiple of 8, any odd k): one output per thread, rows read through L1.
__global__ void __launch_bounds__(256)
post_conv_tanh_generic_kernel(const __nv_bfloat16* __restrict__ act, const float* __restrict__ w, float bias, int B,
                              int T, int C, int k, const PostOut out, int f16) {
  __shared__ float sw[kPostMaxW];
  for (int i = threadIdx.x; i < k * C; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (t >= T) return;
  const int half = (k - 1) / 2;
  float acc = bias;
  for (int j = 0; j < k; ++j) {
    const int tt = t + j - half;
    if (tt < 0 || tt >= T) continue;
    const uint4* row = reinterpret_cast<const uint4*>(act + ((long long)b * T + tt) * C);
    const float* wj = sw + j * C;
    for (int c8 = 0; c8 < C / 8; ++c8) {
      const uint4 v = row[c8];
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float lo, hi;
        unpack16(u[i], lo, hi, f16);
        acc = fmaf(lo, wj[c8 * 8 + 2 * i], acc);
        acc = fmaf(hi, wj[c8 * 8 + 2 * i + 1], acc);
      }
    }
  }
  if (out.pcm) {
    const long long valid = out.lens ? (long long)out.lens[b] * out.hop : (long long)T;
    out.pcm[(long long)b * T + t] = t < valid ? to_pcm16(tanhf(acc), out.scale) : (int16_t)0;
  } else {
    out.wav[(long long)b * T + t] = tanhf(acc);
  }
}

// ---- Postnet tail (N2): y [B][T][ldy] fp32 (channels-last GEMM output, ldy >= C) -> out [B][T][C] fp32, optionally
// + x (the caller's `postnet(output) + output`, unsupervised_fastspeech2/model.py:188) ----
__global__ void __launch_bounds__(256)
postnet_out_kernel(const float* __restrict__ y, const float* __restrict__ x, long long total, int C, int ldy,
                   float* __restrict__ out) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long row = idx / C;
  const int c = (int)(idx - row * C);
  float v = y[row * ldy + c];
  if (x) v += x[idx];
  out[idx] = v;
}

// ---- standalone resblocks: y [B][T][C] fp32 (channels-last GEMM output) -> out [B][C][T] fp32 (the modules' layout),
// 32 x 32 tiles through shared memory so that both sides are coalesced ----
__global__ void __launch_bounds__(256)
cl_to_ncl_kernel(const float* __restrict__ y, float* __restrict__ out, int T, int C) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r, c = c0 + tx;
    tile[r][tx] = (t < T && c < C) ? y[((long long)b * T + t) * C + c] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, t = t0 + tx;
    if (c < C && t < T) out[((long long)b * C + c) * T + t] = tile[tx][r];
  }
}

// ---- iSTFTNet head (class iSTFT, generator.py:91-109) ----
// ReflectionPad1d((1, 0)) on channels-last rows: out[b][0] = in[b][1], out[b][p] = in[b][p-1]; 16 bytes per thread.
__global__ void __launch_bounds__(256)
reflect_pad_left_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int T, int C) {
  const int per_row = C / 8;
  const long long total = (long long)B * (T + 1) * per_row;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int q = (int)(idx % per_row);
  const long long row = idx / per_row;
  const int p = (int)(row % (T + 1));
  const int b = (int)(row / (T + 1));
  const int src = p == 0 ? (T > 1 ? 1 : 0) : p - 1;
  const uint4* s4 = reinterpret_cast<const uint4*>(in + ((long long)b * T + src) * C) + q;
  reinterpret_cast<uint4*>(out + ((long long)b * (T + 1) + p) * C)[q] = __ldg(s4);
}

// y: conv_post output [B][F][ldy] fp32 (channels-last, ldy >= 2*nb).  spec[b][c][f] = exp(y[.][c]),
// phase[b][c][f] = sin(y[.][nb + c]) for c < nb (generator.py:105-106).  A block transposes 32 frames through
// shared memory so both the reads (along channels) and the writes (along frames) are coalesced.
__global__ void __launch_bounds__(256)
spec_phase_kernel(const float* __restrict__ y, int B, int F, int ldy, int nb, float* __restrict__ spec,
                  float* __restrict__ phase) {
  __shared__ float tile[32][65];
  const int f0 = blockIdx.x * 32;
  const int b = blockIdx.y;
  const int nc = 2 * nb;  // <= 64
  for (int i = threadIdx.x; i < 32 * nc; i += blockDim.x) {
    const int f = i / nc, c = i - f * nc;
    if (f0 + f < F) tile[f][c] = y[((long long)b * F + f0 + f) * ldy + c];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * nc; i += blockDim.x) {
    const int c = i / 32, f = i - c * 32;
    if (f0 + f >= F) continue;
    const float v = tile[f][c];
    if (c < nb) spec[((long long)b * nb + c) * F + f0 + f] = expf(v);
    else phase[((long long)b * nb + (c - nb)) * F + f0 + f] = sinf(v);
  }
}

// inverse_stft (stft.py:138-148) = torch.istft(mag * exp(i phase), n_fft, hop, win = n_fft, periodic Hann,
// center=True): y[n] = sum_f w[i] x_f[i] / sum_f w[i]^2 with i = n + n_fft/2 - hop f and x_f = irfft of frame f,
//   x_f[i] = (1/N) (Re X_0 + (-1)^i Re X_{N/2} + 2 sum_{0<k<N/2} (Re X_k cos(2 pi k i / N) - Im X_k sin(2 pi k i / N))).
// A block produces 256 consecutive samples: the frames it touches are converted to (Re, Im) once in shared memory.
constexpr int kIstftMaxN = 64;
__global__ void __launch_bounds__(256)
istft_small_kernel(const float* __restrict__ mag, const float* __restrict__ phase, int B, int F, int N, int hop,
                   float* __restrict__ wav) {
  extern __shared__ float sm[];
  const int nb = N / 2 + 1;
  const int L = hop * (F - 1);
  const int n0 = blockIdx.x * 256;
  const int b = blockIdx.y;
  float* cs = sm;            // [N] cos(2 pi j / N)
  float* sn = cs + N;        // [N] sin(2 pi j / N)
  float* win = sn + N;       // [N] periodic Hann
  float* xr = win + N;       // [nfr][nb]
  const int g0 = n0 + N / 2;                       // untrimmed index of the block's first sample
  const int f_lo = max(0, (g0 - (N - 1) + hop - 1) / hop);
  const int f_hi = min(F - 1, (g0 + 255) / hop);
  const int nfr = f_hi - f_lo + 1;
  float* xi = xr + (256 / hop + N / hop + 2) * nb;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    float s, c;
    sincospif(2.0f * (float)j / (float)N, &s, &c);
    cs[j] = c;
    sn[j] = s;
    win[j] = 0.5f - 0.5f * c;
  }
  for (int i = threadIdx.x; i < nfr * nb; i += blockDim.x) {
    const int fr = i / nb, k = i - fr * nb;
    const long long src = ((long long)b * nb + k) * F + f_lo + fr;
    float s, c;
    sincosf(phase[src], &s, &c);
    const float m = mag[src];
    xr[fr * nb + k] = m * c;
    xi[fr * nb + k] = m * s;
  }
  __syncthreads();
  const int n = n0 + threadIdx.x;
  if (n >= L) return;
  const int g = n + N / 2;
  float num = 0.f, den = 0.f;
  const int fa = max(f_lo, (g - (N - 1) + hop - 1) / hop), fb = min(f_hi, g / hop);
  for (int f = fa; f <= fb; ++f) {
    const int i = g - hop * f;  // 0 <= i < N
    const float* re = xr + (f - f_lo) * nb;
    const float* im = xi + (f - f_lo) * nb;
    float acc = re[0] + ((i & 1) ? -re[nb - 1] : re[nb - 1]);
    for (int k = 1; k < nb - 1; ++k) {
      const int j = (k * i) & (N - 1);  // N is a power of two
      acc += 2.0f * (re[k] * cs[j] - im[k] * sn[j]);
    }
    const float w = win[i];
    num = fmaf(w, acc / (float)N, num);
    den = fmaf(w, w, den);
  }
  wav[(long long)b * L + n] = num / den;
}

}  // namespace e2e
