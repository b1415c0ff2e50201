// STFT -> log-mel front-end in ONE pass over the audio (reference: e2e_tts/src/tools/stft.py:46-89,
// TorchSTFT.mel_spectrogram; :107-135 generate_melspecs; e2e_tts/src/tools/utils.py:22-28).
//
//   reflect-pad 384 | frame t = xp[256 t .. +1024) * periodic Hann | 1024-pt real FFT | sqrt(re^2+im^2+1e-9)
//   | sparse Slaney mel filterbank | log(max(., 1e-5)) | energy = sqrt(sum_k mag^2)
//
// The path is bandwidth-shaped (4 B in, 0.32 B out per sample) but a direct DFT would be ~2 MFLOP per frame,
// so the transform is FFT-structured on chip: a CTA stages the audio of 32 consecutive frames in shared memory
// once (frames overlap 4x), and each group of 64 threads runs a 512-point complex FFT of the even/odd-packed
// frame as three radix-8 passes held in registers (8 complex values per thread), exchanging through padded,
// conflict-free shared-memory maps, followed by the real-FFT recombination pass.  The index maps are emulated
// and checked on the CPU in tests/test_mel_fft_plan.py.  Warp shuffles reduce the frame energy; the mel
// filterbank is applied in its sparse form (727 non-zeros instead of a dense 80x513 GEMM); outputs are staged
// so that the global stores of mel[b][m][t0..t0+32) are 128-byte coalesced.
#include <cmath>
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
constexpr int kF = 32;                    // frames per CTA
constexpr int kGroups = 4;                // 64-thread FFT groups per CTA
constexpr int kThreads = kGroups * 64;
constexpr int kAudio = (kF - 1) * kHop + kNfft;  // samples staged per CTA
constexpr int kSx = 576;                  // padded complex exchange buffer (8 rows of 72 / 64 rows of 9)
constexpr int kMagPad = 528;

struct MelParams {
  const float* wav;
  long long ldw, L;
  int B, T, n_mels, nnz;
  float* mel;
  float* energy;
  int* range_flag;
  const float* window;  // [1024] periodic Hann
  const float2* tw;     // [1024] exp(-2 pi i j / 1024)
  const float* fb_w;    // [nnz] packed filterbank weights
  const int* fb_lo;     // [n_mels] first bin of each filter
  const int* fb_cnt;    // [n_mels] bins per filter
  const int* fb_off;    // [n_mels] offset into fb_w
};

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }

__device__ __forceinline__ void dft4(float2 c0, float2 c1, float2 c2, float2 c3, float2& y0, float2& y1, float2& y2,
                                     float2& y3) {
  const float2 d0 = cadd(c0, c2), d2 = csub(c0, c2), d1 = cadd(c1, c3), d3 = mul_neg_i(csub(c1, c3));
  y0 = cadd(d0, d1);
  y1 = cadd(d2, d3);
  y2 = csub(d0, d1);
  y3 = csub(d2, d3);
}

// Radix-2 decimation-in-frequency 8-point DFT, natural-order output.
__device__ __forceinline__ void dft8(float2 (&a)[8]) {
  const float r = 0.70710678118654752440f;
  const float2 b0 = cadd(a[0], a[4]), b1 = cadd(a[1], a[5]), b2 = cadd(a[2], a[6]), b3 = cadd(a[3], a[7]);
  const float2 b4 = csub(a[0], a[4]);
  float2 b5 = csub(a[1], a[5]), b6 = csub(a[2], a[6]), b7 = csub(a[3], a[7]);
  b5 = make_float2(r * (b5.x + b5.y), r * (b5.y - b5.x));   // * W8^1
  b6 = mul_neg_i(b6);                                       // * W8^2
  b7 = make_float2(r * (b7.y - b7.x), -r * (b7.x + b7.y));  // * W8^3
  dft4(b0, b1, b2, b3, a[0], a[2], a[4], a[6]);
  dft4(b4, b5, b6, b7, a[1], a[3], a[5], a[7]);
}

__device__ __forceinline__ void group_sync(int g) {
  asm volatile("bar.sync %0, 64;" ::"r"(g + 1) : "memory");
}

__global__ void __launch_bounds__(kThreads, 2) mel_kernel(const MelParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* audio = reinterpret_cast<float*>(smem);                    // [kAudio]
  float2* tw_s = reinterpret_cast<float2*>(audio + kAudio);         // [512]
  float2* sx = tw_s + 512;                                          // [kGroups][2][kSx]
  float* mag_all = reinterpret_cast<float*>(sx + kGroups * 2 * kSx);  // [kGroups][kMagPad]
  float* out_s = mag_all + kGroups * kMagPad;                       // [n_mels][kF+1]
  float* energy_s = out_s + p.n_mels * (kF + 1);                    // [kF]
  float* red_s = energy_s + kF;                                     // [kGroups][2]
  float* fbw_s = red_s + kGroups * 2;                               // [nnz]
  int* fb_lo_s = reinterpret_cast<int*>(fbw_s + ((p.nnz + 3) & ~3));  // [n_mels]
  int* fb_cnt_s = fb_lo_s + p.n_mels;
  int* fb_off_s = fb_cnt_s + p.n_mels;

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * kF;
  const int nf = min(kF, p.T - f0);

  // ---- stage audio (reflect padding resolved here, stft.py:60-64), tables and the sparse filterbank ----
  {
    const float* row = p.wav + (long long)b * p.ldw;
    const long long base = (long long)f0 * kHop - kPad;  // source index of audio[0]
    const int need = (nf - 1) * kHop + kNfft;
    bool bad = false;
    for (int i = tid; i < need; i += kThreads) {
      long long s = base + i;
      if (s < 0) s = -s;
      if (s >= p.L) s = 2 * (p.L - 1) - s;
      const float v = row[s];
      bad |= !(v >= -1.0f && v <= 1.0f);
      audio[i] = v;
    }
    if (bad && p.range_flag) atomicOr(p.range_flag, 1);
    for (int i = tid; i < 512; i += kThreads) tw_s[i] = p.tw[i];
    for (int i = tid; i < p.nnz; i += kThreads) fbw_s[i] = p.fb_w[i];
    for (int i = tid; i < p.n_mels; i += kThreads) {
      fb_lo_s[i] = p.fb_lo[i];
      fb_cnt_s[i] = p.fb_cnt[i];
      fb_off_s[i] = p.fb_off[i];
    }
  }

  const int g = tid >> 6;   // FFT group
  const int t = tid & 63;   // thread within the group
  const int hi = t >> 3, lo = t & 7;
  float2* S1 = sx + g * 2 * kSx;
  float2* S2 = S1 + kSx;
  float* mag = mag_all + g * kMagPad;

  // per-thread constants: window taps and twiddles of passes 1 and 2
  float w0[8], w1[8];
  float2 t1[8], t2[8];
#pragma unroll
  for (int n1 = 0; n1 < 8; ++n1) {
    w0[n1] = p.window[128 * n1 + 2 * t];
    w1[n1] = p.window[128 * n1 + 2 * t + 1];
    t1[n1] = p.tw[2 * ((t * n1) & 511)];    // W_512^(n2 k1), n2 = t
    t2[n1] = p.tw[16 * ((lo * n1) & 63)];   // W_64^(b c),   b = lo
  }
  __syncthreads();

  for (int fl = g; fl < nf; fl += kGroups) {
    float2 a[8];
    // pass 1: thread n2 = t, points z[64 n1 + n2] = (xw[128 n1 + 2 n2], xw[128 n1 + 2 n2 + 1])
    const float2* src = reinterpret_cast<const float2*>(audio + fl * kHop) + t;
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) {
      const float2 v = src[64 * n1];
      a[n1] = make_float2(v.x * w0[n1], v.y * w1[n1]);
    }
    dft8(a);
    S1[t] = a[0];
#pragma unroll
    for (int k1 = 1; k1 < 8; ++k1) S1[k1 * 72 + t] = cmul(a[k1], t1[k1]);
    group_sync(g);
    // pass 2: thread (k1 = hi, b = lo) reads Y[k1][8 a + b]
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = S1[hi * 72 + 8 * q + lo];
    dft8(a);
    S2[hi * 72 + lo] = a[0];
#pragma unroll
    for (int c = 1; c < 8; ++c) S2[hi * 72 + c * 9 + lo] = cmul(a[c], t2[c]);
    group_sync(g);
    // pass 3: thread (k1 = hi, c = lo) reads U[k1][c][b]; result Z[k1 + 8 c + 64 d] -> S3 (aliases S1)
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = S2[hi * 72 + lo * 9 + q];
    dft8(a);
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      const int k = hi + 8 * lo + 64 * d;
      S1[k + (k >> 3)] = a[d];
    }
    group_sync(g);
    // real-FFT recombination: thread t owns bins k = t + 64 j
    float e = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = t + 64 * j;
      const int kk = (512 - k) & 511;
      const float2 za = S1[k + (k >> 3)];
      const float2 zb = S1[kk + (kk >> 3)];
      // Ze = (Z[k] + conj Z[N-k]) / 2 ; Zo = (Z[k] - conj Z[N-k]) / (2i)
      const float2 ze = make_float2(0.5f * (za.x + zb.x), 0.5f * (za.y - zb.y));
      const float2 zo = make_float2(0.5f * (za.y + zb.y), -0.5f * (za.x - zb.x));
      const float2 x = cadd(ze, cmul(tw_s[k], zo));
      const float s = x.x * x.x + x.y * x.y + 1e-9f;  // stft.py:77
      mag[k] = sqrtf(s);
      e += s;
      if (k == 0) {  // Nyquist bin: X[512] = Ze[0] - Zo[0] (purely real)
        const float xn = ze.x - zo.x;
        const float sn = xn * xn + 1e-9f;
        mag[512] = sqrtf(sn);
        e += sn;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((t & 31) == 0) red_s[g * 2 + (t >> 5)] = e;
    group_sync(g);
    // sparse mel filterbank + log compression (stft.py:80-81, utils.py:28)
    for (int m = t; m < p.n_mels; m += 64) {
      const int lo_bin = fb_lo_s[m], cnt = fb_cnt_s[m];
      const float* w = fbw_s + fb_off_s[m];
      float acc = 0.f;
      for (int i = 0; i < cnt; ++i) acc = fmaf(w[i], mag[lo_bin + i], acc);
      out_s[m * (kF + 1) + fl] = logf(fmaxf(acc, 1e-5f));
    }
    if (t == 0) energy_s[fl] = sqrtf(red_s[g * 2] + red_s[g * 2 + 1]);  // stft.py:84
  }
  __syncthreads();

  // ---- coalesced stores: mel[b][m][f0 .. f0+nf) ----
  for (int i = tid; i < p.n_mels * kF; i += kThreads) {
    const int m = i / kF, f = i - m * kF;
    if (f < nf) p.mel[((long long)b * p.n_mels + m) * p.T + f0 + f] = out_s[m * (kF + 1) + f];
  }
  if (p.energy && tid < nf) p.energy[(long long)b * p.T + f0 + tid] = energy_s[tid];
}

}  // namespace

struct e2e_mel {
  int n_fft, hop, win, n_mels, nnz;
  float* d_window = nullptr;
  float2* d_tw = nullptr;
  float* d_fbw = nullptr;
  int* d_lo = nullptr;
  int* d_cnt = nullptr;
  int* d_off = nullptr;
  int smem_bytes = 0;
};

extern "C" int e2e_mel_create(int32_t n_fft, int32_t hop_length, int32_t win_length, int32_t n_mels,
                              const float* mel_basis, e2e_mel** out) {
  if (!mel_basis || !out) return fail(-1, "null argument");
  if (n_fft != kNfft || win_length != kNfft || hop_length != kHop)
    return fail(-4, "mel front-end supports n_fft == win_length == 1024 and hop_length == 256 (the e2e-tts config)");
  if (n_mels < 1 || n_mels > 128) return fail(-4, "n_mels must be in [1, 128]");
  e2e_mel* m = new e2e_mel;
  m->n_fft = n_fft;
  m->hop = hop_length;
  m->win = win_length;
  m->n_mels = n_mels;
  std::vector<float> w;
  std::vector<int> lo(n_mels), cnt(n_mels), off(n_mels);
  for (int r = 0; r < n_mels; ++r) {
    const float* row = mel_basis + (size_t)r * kBins;
    int first = -1, last = -1;
    for (int k = 0; k < kBins; ++k)
      if (row[k] != 0.f) {
        if (first < 0) first = k;
        last = k;
      }
    lo[r] = first < 0 ? 0 : first;
    cnt[r] = first < 0 ? 0 : last - first + 1;
    off[r] = (int)w.size();
    for (int k = 0; k < cnt[r]; ++k) w.push_back(row[lo[r] + k]);
  }
  m->nnz = (int)w.size();
  if (w.empty()) w.push_back(0.f);
  std::vector<float> window(kNfft);
  std::vector<float2> tw(kNfft);
  const double pi = 3.14159265358979323846;
  for (int n = 0; n < kNfft; ++n) {
    window[n] = (float)(0.5 - 0.5 * cos(2.0 * pi * n / kNfft));  // periodic Hann, stft.py:44
    tw[n] = make_float2((float)cos(2.0 * pi * n / kNfft), (float)(-sin(2.0 * pi * n / kNfft)));
  }
  m->smem_bytes = (kAudio + 2 * 512 + kGroups * 2 * kSx * 2 + kGroups * kMagPad + n_mels * (kF + 1) + kF +
                   kGroups * 2 + ((m->nnz + 3) & ~3) + 3 * n_mels) * 4;
  if (m->smem_bytes > 113 * 1024) {
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
  up((void**)&m->d_fbw, w.data(), w.size() * 4);
  up((void**)&m->d_lo, lo.data(), lo.size() * 4);
  up((void**)&m->d_cnt, cnt.data(), cnt.size() * 4);
  up((void**)&m->d_off, off.data(), off.size() * 4);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, m->smem_bytes);
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
  cudaFree(m->d_lo);
  cudaFree(m->d_cnt);
  cudaFree(m->d_off);
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
  p.nnz = m->nnz;
  p.mel = mel;
  p.energy = energy;
  p.range_flag = range_flag;
  p.window = m->d_window;
  p.tw = m->d_tw;
  p.fb_w = m->d_fbw;
  p.fb_lo = m->d_lo;
  p.fb_cnt = m->d_cnt;
  p.fb_off = m->d_off;
  dim3 grid((unsigned)((T + kF - 1) / kF), (unsigned)B);
  mel_kernel<<<grid, kThreads, m->smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, std::string("mel_kernel launch: ") + cudaGetErrorString(e));
  return 0;
}
