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
// conflict-free shared-memory maps, followed by the real-FFT recombination pass, which produces the bins k and
// 512-k from one pair of loads (X[512-k] = conj(Ze - W^k Zo)) and only the bins the filterbank reads (0..371 for
// fmax = 8 kHz).  The frame energy sqrt(sum_k mag_k^2) over all 513 bins comes from Parseval's identity on the
// windowed samples: sum_{k<=512} |X_k|^2 = (1024 sum_n xw_n^2 + X_0^2 + X_512^2) / 2 (shuffle-reduced).  The mel
// filterbank is applied in its sparse, block form: thread t of a frame owns BPT consecutive bins and the (at most FPB)
// triangular filters that overlap them, as a dense FPB x BPT block of weights - branch-free, conflict-free, and
// deterministic (fixed-order partial sums per filter); outputs are staged so that the global stores of
// mel[b][m][t0..t0+32) are 128-byte coalesced.  The index maps are emulated and
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
constexpr int kF = 32;                    // frames per CTA
constexpr int kGroups = 4;                // 64-thread FFT groups per CTA
constexpr int kThreads = kGroups * 64;
constexpr int kAudio = (kF - 1) * kHop + kNfft;  // samples staged per CTA
constexpr int kSx = 576;                  // padded complex exchange buffer (8 rows of 72 / 64 rows of 9)
constexpr int kMagPad = 584;              // 64 blocks x pitch 9 (BPT = 8), multiple of 4
constexpr int kFbSplitMax = 12;           // partial-sum slots per filter (a filter spans <= 12 threads' bin blocks)

struct MelParams {
  const float* wav;
  long long ldw, L;
  int B, T, n_mels, nnz;
  int nb;               // bins the filterbank reads: 1 + last non-zero column of the basis
  int fb_split;         // partial-sum slots per filter = the most bin blocks any filter overlaps
  float* mel;
  float* energy;
  int* range_flag;
  const float* window;  // [1024] periodic Hann
  const float2* tw;     // [1024] exp(-2 pi i j / 1024)
  const float4* fb_w;   // [ceil(FPB*BPT/4)][64] weights of thread t's block: element j*BPT + i = basis[filter j][bin i]
  const int* fb_slot;   // [FPB][64] where filter j of thread t leaves its partial sum: mel * fb_split + (index of the
                        // thread among the filter's threads); unused j -> the scratch slot n_mels * fb_split
};

__device__ __forceinline__ float sqrt_approx(float x) {  // MUFU.SQRT: max relative error 2^-23 (PTX ISA)
  float y;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

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

// BPT = bins per thread, FPB = filters per bin block.  <6, 5> covers the e2e-tts basis (372 bins, 80 Slaney filters),
// <8, 8> any triangular bank on <= 512 bins the constructor accepts.  The magnitudes are stored with a row pitch of
// BPT | 1 words per block, so the 64 threads read their blocks with an odd stride (no bank conflicts).
template <int BPT, int FPB>
__global__ void __launch_bounds__(kThreads, 2) mel_kernel(const MelParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* audio = reinterpret_cast<float*>(smem);                    // [kAudio]
  float2* tw_s = reinterpret_cast<float2*>(audio + kAudio);         // [512]
  float2* sx = tw_s + 512;                                          // [kGroups][2][kSx]
  float* mag_all = reinterpret_cast<float*>(sx + kGroups * 2 * kSx);  // [kGroups][kMagPad]
  float* out_s = mag_all + kGroups * kMagPad;                       // [n_mels][kF+1]
  float* energy_s = out_s + ((p.n_mels * (kF + 1) + 3) & ~3);       // [kF]
  float* red_s = energy_s + kF;                                     // [kGroups][2]
  constexpr int PITCH = BPT | 1;
  constexpr int NW4 = (FPB * BPT + 3) / 4;
  const int part_n = (p.n_mels + 1) * p.fb_split;                   // + the scratch slot row
  float* part_all = red_s + kGroups * 2;                            // [kGroups][n_mels + 1][fb_split]
  float4* fb_w_s = reinterpret_cast<float4*>(part_all + ((kGroups * part_n + 3) & ~3));  // [NW4][64]
  int* fb_slot_s = reinterpret_cast<int*>(fb_w_s + NW4 * 64);       // [FPB][64]

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * kF;
  const int nf = min(kF, p.T - f0);

  // ---- stage audio (reflect padding resolved here, stft.py:60-64), tables and the sparse filterbank ----
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
    for (int i = tid; i < 512; i += kThreads) tw_s[i] = p.tw[i];
    for (int i = tid; i < NW4 * 64; i += kThreads) fb_w_s[i] = p.fb_w[i];
    for (int i = tid; i < FPB * 64; i += kThreads) fb_slot_s[i] = p.fb_slot[i];
    for (int i = tid; i < kGroups * part_n; i += kThreads) part_all[i] = 0.f;   // unused slots stay 0
    for (int i = tid; i < kGroups * kMagPad; i += kThreads) mag_all[i] = 0.f;   // bins past nb: weight 0 x finite
  }

  const int g = tid >> 6;   // FFT group
  const int t = tid & 63;   // thread within the group
  const int hi = t >> 3, lo = t & 7;
  float2* S1 = sx + g * 2 * kSx;
  float2* S2 = S1 + kSx;
  float* mag = mag_all + g * kMagPad;
  float* part = part_all + g * part_n;
  const int nb = p.nb;

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
    float e = 0.f;  // sum of the squared windowed samples of this thread
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) {
      const float2 v = src[64 * n1];
      a[n1] = make_float2(v.x * w0[n1], v.y * w1[n1]);
      e = fmaf(a[n1].x, a[n1].x, e);
      e = fmaf(a[n1].y, a[n1].y, e);
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
    // pass 3: thread (k1 = hi, c = lo) reads U[k1][c][b]; result Z[k] , k = k1 + 8 c + 64 d, -> S3[k ^ ((k >> 3) & 7)] (aliases S1;
    // the xor spreads both this scattered store and the recombination's paired loads over the banks)
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = S2[hi * 72 + lo * 9 + q];
    dft8(a);
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      const int k = hi + 8 * lo + 64 * d;
      S1[k ^ ((k >> 3) & 7)] = a[d];
    }
    group_sync(g);
    // real-FFT recombination: thread t owns the bin pairs (k, 512 - k), k = t + 64 j, j = 0..3; thread 0 also
    // bin 256.  Ze = (Z[k] + conj Z[512-k]) / 2, Zo = (Z[k] - conj Z[512-k]) / (2i), X[k] = Ze + W^k Zo,
    // X[512-k] = conj(Ze - W^k Zo).  k = 0 gives the purely real X[0] and X[512].
    float x0sq = 0.f, xnsq = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = t + 64 * j;
      const int kk = (512 - k) & 511;
      const float2 za = S1[k ^ ((k >> 3) & 7)];
      const float2 zb = S1[kk ^ ((kk >> 3) & 7)];
      const float2 ze = make_float2(0.5f * (za.x + zb.x), 0.5f * (za.y - zb.y));
      const float2 zo = make_float2(0.5f * (za.y + zb.y), -0.5f * (za.x - zb.x));
      const float2 wz = cmul(tw_s[k], zo);
      const float2 x = cadd(ze, wz), xp = csub(ze, wz);
      const float s = x.x * x.x + x.y * x.y + 1e-9f;  // stft.py:77
      const float sp = xp.x * xp.x + xp.y * xp.y + 1e-9f;
      if (k < nb) mag[k + (k / BPT) * (PITCH - BPT)] = sqrt_approx(s);
      if (512 - k < nb) mag[(512 - k) + ((512 - k) / BPT) * (PITCH - BPT)] = sqrt_approx(sp);
      if (k == 0) {
        x0sq = x.x * x.x;
        xnsq = xp.x * xp.x;
      }
    }
    if (t == 0 && 256 < nb) {
      const float2 z = S1[256];  // 256 ^ ((256 >> 3) & 7)
      mag[256 + (256 / BPT) * (PITCH - BPT)] = sqrt_approx(z.x * z.x + z.y * z.y + 1e-9f);
    }
    // energy^2 = sum_{k=0..512} (|X_k|^2 + 1e-9) = (1024 sum xw^2 + X_0^2 + X_512^2) / 2 + 513e-9   (stft.py:84)
    e = fmaf(1024.f, e, x0sq + xnsq);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((t & 31) == 0) red_s[g * 2 + (t >> 5)] = e;
    group_sync(g);
    // sparse mel filterbank (stft.py:80): thread t applies the FPB x BPT weight block of its bins and leaves one
    // partial sum per filter in that filter's slot of this thread
    {
      float mv[BPT];
#pragma unroll
      for (int i = 0; i < BPT; ++i) mv[i] = mag[t * PITCH + i];
      float w[NW4 * 4];
#pragma unroll
      for (int q = 0; q < NW4; ++q) {
        const float4 v = fb_w_s[q * 64 + t];
        w[4 * q] = v.x;
        w[4 * q + 1] = v.y;
        w[4 * q + 2] = v.z;
        w[4 * q + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < FPB; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < BPT; ++i) acc = fmaf(w[j * BPT + i], mv[i], acc);
        part[fb_slot_s[j * 64 + t]] = acc;
      }
    }
    if (t == 0) energy_s[fl] = sqrtf(0.5f * (red_s[g * 2] + red_s[g * 2 + 1]) + 513e-9f);
    group_sync(g);
    // fixed-order sum of the partials + log compression (stft.py:81, utils.py:28)
    for (int m = t; m < p.n_mels; m += 64) {
      float acc = 0.f;
      for (int q = 0; q < p.fb_split; ++q) acc += part[m * p.fb_split + q];
      out_s[m * (kF + 1) + fl] = logf(fmaxf(acc, 1e-5f));
    }
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
  int nb = 0, split = 1, variant = 0;  // variant 0: mel_kernel<6, 5>, 1: mel_kernel<8, 8>
  float* d_window = nullptr;
  float2* d_tw = nullptr;
  float4* d_fbw = nullptr;
  int* d_slot = nullptr;
  int smem_bytes = 0;
};

extern "C" int e2e_mel_create(int32_t n_fft, int32_t hop_length, int32_t win_length, int32_t n_mels,
                              const float* mel_basis, e2e_mel** out) {
  if (!mel_basis || !out) return fail(-1, "null argument");
  if (n_fft != kNfft || win_length != kNfft || hop_length != kHop)
    return fail(-4, "mel front-end supports n_fft == win_length == 1024 and hop_length == 256 (the e2e-tts config)");
  if (n_mels < 1 || n_mels > 128) return fail(-4, "n_mels must be in [1, 128]");
  // sparse filterbank in block form: thread t of a frame owns bins [t*BPT, (t+1)*BPT) and the filters overlapping them
  int nb = 1, nnz = 0;
  for (int r = 0; r < n_mels; ++r)
    for (int k = 0; k < kBins; ++k)
      if (mel_basis[(size_t)r * kBins + k] != 0.f) {
        ++nnz;
        nb = k + 1 > nb ? k + 1 : nb;
      }
  if (nb > 512) return fail(-4, "mel filterbank reaches the Nyquist bin: unsupported (fmax must be below sr/2)");
  int variant = -1, BPT = 0, FPB = 0, split = 1;
  std::vector<float> wblk;
  std::vector<int> slot;
  for (int v = 0; v < 2 && variant < 0; ++v) {
    BPT = v == 0 ? 6 : 8;
    FPB = v == 0 ? 5 : 8;
    if (64 * BPT < nb) continue;
    const int nw4 = (FPB * BPT + 3) / 4;
    // filters of every block, and the blocks of every filter
    std::vector<std::vector<int>> filt(64);
    std::vector<int> first(n_mels, -1), count(n_mels, 0);
    bool fits = true;
    for (int t = 0; t < 64 && fits; ++t)
      for (int r = 0; r < n_mels; ++r) {
        bool any = false;
        for (int i = 0; i < BPT; ++i) {
          const int k = t * BPT + i;
          any = any || (k < kBins && mel_basis[(size_t)r * kBins + k] != 0.f);
        }
        if (!any) continue;
        filt[t].push_back(r);
        if (first[r] < 0) first[r] = t;
        count[r] = t - first[r] + 1;  // blocks first..t (a gap inside a filter just leaves a zero partial)
        if ((int)filt[t].size() > FPB) fits = false;
      }
    if (!fits) continue;
    split = 1;
    for (int r = 0; r < n_mels; ++r) split = count[r] > split ? count[r] : split;
    if (split > kFbSplitMax) continue;
    variant = v;
    wblk.assign((size_t)nw4 * 4 * 64, 0.f);
    slot.assign((size_t)FPB * 64, n_mels * split);  // unused filters of a block -> the scratch slot row
    for (int t = 0; t < 64; ++t)
      for (size_t j = 0; j < filt[t].size(); ++j) {
        const int r = filt[t][j];
        slot[j * 64 + t] = r * split + (t - first[r]);
        for (int i = 0; i < BPT; ++i) {
          const int k = t * BPT + i;
          const int e = (int)j * BPT + i;  // element e of the thread's block lives in float4 e/4, lane e%4
          wblk[((size_t)(e / 4) * 64 + t) * 4 + (e % 4)] = k < kBins ? mel_basis[(size_t)r * kBins + k] : 0.f;
        }
      }
  }
  if (variant < 0)
    return fail(-4, "mel filterbank does not fit the kernel's block form (<= 8 filters per 8-bin block, <= 12 blocks per filter)");
  e2e_mel* m = new e2e_mel;
  m->n_fft = n_fft;
  m->hop = hop_length;
  m->win = win_length;
  m->n_mels = n_mels;
  m->nnz = nnz;
  m->nb = nb;
  m->split = split;
  m->variant = variant;
  std::vector<float> window(kNfft);
  std::vector<float2> tw(kNfft);
  const double pi = 3.14159265358979323846;
  for (int n = 0; n < kNfft; ++n) {
    window[n] = (float)(0.5 - 0.5 * cos(2.0 * pi * n / kNfft));  // periodic Hann, stft.py:44
    tw[n] = make_float2((float)cos(2.0 * pi * n / kNfft), (float)(-sin(2.0 * pi * n / kNfft)));
  }
  // mirrors the carve-up at the top of mel_kernel
  {
    const int part_n = (n_mels + 1) * split;
    const int nw4 = (FPB * BPT + 3) / 4;
    m->smem_bytes = (kAudio + 2 * 512 + kGroups * 2 * kSx * 2 + kGroups * kMagPad + ((n_mels * (kF + 1) + 3) & ~3) + kF +
                     kGroups * 2 + ((kGroups * part_n + 3) & ~3) + nw4 * 4 * 64 + FPB * 64) * 4;
  }
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
  up((void**)&m->d_fbw, wblk.data(), wblk.size() * 4);
  up((void**)&m->d_slot, slot.data(), slot.size() * 4);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(variant == 0 ? mel_kernel<6, 5> : mel_kernel<8, 8>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, m->smem_bytes);
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
  cudaFree(m->d_slot);
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
  p.nb = m->nb;
  p.fb_split = m->split;
  p.mel = mel;
  p.energy = energy;
  p.range_flag = range_flag;
  p.window = m->d_window;
  p.tw = m->d_tw;
  p.fb_w = m->d_fbw;
  p.fb_slot = m->d_slot;
  dim3 grid((unsigned)((T + kF - 1) / kF), (unsigned)B);
  if (m->variant == 0)
    mel_kernel<6, 5><<<grid, kThreads, m->smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  else
    mel_kernel<8, 8><<<grid, kThreads, m->smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, std::string("mel_kernel launch: ") + cudaGetErrorString(e));
  return 0;
}
