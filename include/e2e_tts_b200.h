/*
 * e2e_tts_b200 — C ABI of the B200-native synthesis hot path of InterlinkLabs/e2e-tts.
 *
 * The reference has no FFI: its "plugin API" for this path is the duck-typed Python contract
 *   HifiGan(config).forward(mel[B,80,T]) -> wav[B,1,256*T]      e2e_tts/models/vocoder/generator.py:13-53
 *   TorchSTFT(...).mel_spectrogram(wav[B,L]) -> mel[B,80,T]      e2e_tts/src/tools/stft.py:11-89
 *   generate_melspecs(y, ...)                                    e2e_tts/src/tools/stft.py:107-135
 * called from e2e_tts/src/api/utils.py:53-56,144-145 and e2e_tts/src/tools/tools_for_data.py:114,178.
 * The Python mirror of that contract (package e2e_tts_b200) binds exactly the entry points below with ctypes;
 * INTEGRATION.md shows the stub.  Plain pointers and sizes only, no torch types, no exceptions across the
 * boundary.  Every function returns 0 on success, a negative value for a bad argument / unsupported
 * configuration, or a positive cudaError_t; e2e_last_error_string() describes the last failure of the
 * calling thread.  All device pointers must belong to the current CUDA device; work is enqueued on `stream`
 * (a cudaStream_t passed as void*) and never synchronises the host.
 */
#ifndef E2E_TTS_B200_H
#define E2E_TTS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define E2E_MAX_UPSAMPLES 8
#define E2E_MAX_KERNELS 8
#define E2E_MAX_DILATIONS 8

/* Mirrors the `hifigan:` mapping of e2e_tts/config/model_config.yaml:75-82 as consumed by
 * HifiGan.__init__ (generator.py:14-35). */
typedef struct e2e_voc_config {
  int32_t in_channels;              /* 80, hard-coded at generator.py:18 */
  int32_t upsample_initial_channel; /* 512 */
  int32_t resblock;                 /* 1 -> ResBlock1 (layers.py:10-46), anything else -> ResBlock2 (layers.py:49-69) */
  int32_t num_upsamples;
  int32_t upsample_rates[E2E_MAX_UPSAMPLES];
  int32_t upsample_kernel_sizes[E2E_MAX_UPSAMPLES];
  int32_t num_kernels;
  int32_t resblock_kernel_sizes[E2E_MAX_KERNELS];
  int32_t num_dilations[E2E_MAX_KERNELS];
  int32_t resblock_dilation_sizes[E2E_MAX_KERNELS][E2E_MAX_DILATIONS];
  /* 0: HiFi-GAN head, conv_post C->1 + tanh (generator.py:33,49-51).  n > 0: iSTFTNet head (class iSTFT,
   * generator.py:65-109; `istft:` mapping, model_config.yaml:83-92) with gen_istft_n_fft = n: ReflectionPad1d((1,0)),
   * conv_post C -> n+2, exp on the first n/2+1 channels, sin on the rest; use e2e_voc_forward_spec. */
  int32_t istft_n_fft;
} e2e_voc_config;

typedef struct e2e_voc e2e_voc; /* opaque: packed 16-bit weights + launch plans of one generator */

/* Threading: a handle is NOT internally synchronised.  One host thread at a time may call into a given e2e_voc /
 * e2e_postnet / e2e_mel handle (the launch-plan cache, the one-shot profile events and the weight images are plain
 * members), and one handle should be driven from one stream at a time (a forward reuses the caller's workspace, and
 * e2e_voc_load_layer synchronises the device before it overwrites weights a forward in flight could still read).
 * Different handles are independent and may be used from different threads concurrently;
 * e2e_last_error_string() is per calling thread. */

/* Replaces HifiGan.__init__ (generator.py:14-35).  Allocates device memory for the packed weights. */
int e2e_voc_create(const e2e_voc_config* cfg, e2e_voc** out);
void e2e_voc_destroy(e2e_voc* v);

/* Replaces load_state_dict for one layer (e2e_tts/src/api/utils.py:54-55).  `name` is the reference's
 * state-dict prefix: "conv_pre", "ups.<i>", "resblocks.<n>.convs1.<m>", "resblocks.<n>.convs2.<m>",
 * "resblocks.<n>.convs.<m>" (ResBlock2), "conv_post".  `weight` is the FOLDED fp32 weight on the HOST in the
 * reference's own layout (Conv1d [C_out][C_in][k]; ConvTranspose1d [C_in][C_out][k]), i.e. g*v/||v|| already
 * applied (generator.py:55-62 semantics); `bias` is [C_out] fp32 on the host.  The call rearranges the weight
 * into GEMM columns, rounds to bf16, swizzles and uploads it (synchronous; not a hot-path call). */
int e2e_voc_load_layer(e2e_voc* v, const char* name, const float* weight, int64_t weight_numel, const float* bias,
                       int64_t bias_numel);

/* Number of layers that still lack weights (0 = ready for e2e_voc_forward). */
int e2e_voc_missing_layers(const e2e_voc* v);

/* Device scratch needed by e2e_voc_forward for a [B,80,T] input. */
size_t e2e_voc_workspace_bytes(const e2e_voc* v, int32_t B, int32_t T);

/* Replaces HifiGan.forward (generator.py:37-53).  mel: device fp32, element (b,c,t) at mel[b*sB + c*sC + t*sT]
 * (any strides: the reference passes a transposed view, utils.py:144).  wav: device fp32 [B][upsample*T]
 * contiguous (= [B,1,256*T]).  workspace: device, >= e2e_voc_workspace_bytes, 1024-byte aligned. */
int e2e_voc_forward(e2e_voc* v, const float* mel, int64_t sB, int64_t sC, int64_t sT, int32_t B, int32_t T,
                    float* wav, void* workspace, size_t workspace_bytes, void* stream);

/* e2e_voc_forward fused with the caller's post-processing (combine_audio, e2e_tts/src/api/utils.py:108-117:
 * `audio[: mel_len * hop] * max_wav_value` ... `.astype("int16")`): pcm: device int16 [B][upsample*T];
 * pcm[b][t] = (int16) trunc(wav[b][t] * max_wav_value) for t < mel_lengths[b] * upsample, 0 beyond.  mel_lengths:
 * device int32 [B] or NULL (= every utterance has T frames).  Truncation toward zero like numpy's astype; a sample
 * of exactly +1.0 saturates to 32767 (numpy wraps it).  Halves the device->host and multi-GPU gather bytes. */
int e2e_voc_forward_pcm16(e2e_voc* v, const float* mel, int64_t sB, int64_t sC, int64_t sT, int32_t B, int32_t T,
                          const int32_t* mel_lengths, float max_wav_value, int16_t* pcm, void* workspace,
                          size_t workspace_bytes, void* stream);

/* Replaces iSTFT.forward (generator.py:91-109) for a generator created with istft_n_fft = n > 0.  spec, phase:
 * device fp32 [B][n/2+1][upsample*T + 1] (the reference's return layout; the +1 frame is the reflection pad). */
int e2e_voc_forward_spec(e2e_voc* v, const float* mel, int64_t sB, int64_t sC, int64_t sT, int32_t B, int32_t T,
                         float* spec, float* phase, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces inverse_stft (e2e_tts/src/tools/stft.py:138-148): torch.istft(mag * exp(i*phase), n_fft, hop, win,
 * periodic Hann window, center=True) for the small transforms of the iSTFTNet head.  mag, phase: device fp32
 * [B][n_fft/2+1][frames]; wav: device fp32 [B][hop*(frames-1)].  Supported: win == n_fft <= 64, n_fft % hop == 0. */
int e2e_istft_forward(const float* mag, const float* phase, int32_t B, int32_t frames, int32_t n_fft, int32_t hop,
                      int32_t win, float* wav, void* stream);

/* Measurement hook: the NEXT e2e_voc_forward records `ev_begin` (a cudaEvent_t) on its stream just before its first
 * tensor-core convolution launch and `ev_end` right after its last one, then forgets both (one-shot).  bench.py
 * uses it to time the dominant kernel family inside the timed region.  Pass NULLs to cancel. */
int e2e_voc_set_profile_events(e2e_voc* v, void* ev_begin, void* ev_end);

/* Format of the tensor-core operands and of the 16-bit activation tensors between layers.  bf16 (default) is what
 * BASELINE names; fp16 runs at the same tensor-core rate with 11 instead of 8 significand bits (about 8x smaller
 * rounding error in the waveform), at the price of fp16's range: conversions saturate at +-65504
 * instead of overflowing.  Changing the format marks every layer as not loaded (the packed weights depend on it):
 * call it right after e2e_voc_create, before e2e_voc_load_layer.  fp32 accumulation, biases, residual adds and the
 * conv_post / tanh tail are unaffected. */
#define E2E_OPERAND_BF16 0
#define E2E_OPERAND_FP16 1
int e2e_voc_set_operand_dtype(e2e_voc* v, int32_t dtype);
int e2e_voc_operand_dtype(const e2e_voc* v);

/* Total upsampling factor (product of upsample_rates; 256 for the default config). */
int e2e_voc_hop(const e2e_voc* v);

/* Kernel launches enqueued by one e2e_voc_forward call at the current configuration. */
int e2e_voc_launches_per_forward(const e2e_voc* v);

/* A forward whose (B, T, input, output, workspace) combination has been seen before is captured into a CUDA graph
 * the second time and replayed afterwards (serving loops that reuse their buffers; E2E_NO_GRAPH=1 disables it).
 * Returns 1 if the handle's last forward was such a replay, 0 if it enqueued its kernels one by one. */
int e2e_voc_last_forward_was_graph(const e2e_voc* v);

/* ---- standalone residual blocks ----
 * Replace ResBlock1.forward / ResBlock2.forward (e2e_tts/models/vocoder/layers.py:33-40, 60-65) as modules of their
 * own (the generator runs the same convolutions through its fused launches).  kind: 1 = ResBlock1 (pairs of a dilated
 * and an undilated conv), 2 = ResBlock2 (one dilated conv per step).  Layer names: "convs1.<m>", "convs2.<m>" (kind 1)
 * or "convs.<m>" (kind 2), weights FOLDED fp32 [C][C][k] on the host as for e2e_voc_load_layer.  x: device fp32, element
 * (b, c, t) at x[b*sB + c*sC + t*sT]; out: device fp32 [B][C][T] contiguous. */
typedef struct e2e_resblock e2e_resblock;
int e2e_resblock_create(int32_t kind, int32_t channels, int32_t kernel_size, const int32_t* dilations,
                        int32_t n_dilations, e2e_resblock** out);
void e2e_resblock_destroy(e2e_resblock* rb);
int e2e_resblock_load_layer(e2e_resblock* rb, const char* name, const float* weight, int64_t weight_numel,
                            const float* bias, int64_t bias_numel);
size_t e2e_resblock_workspace_bytes(const e2e_resblock* rb, int32_t B, int32_t T);
int e2e_resblock_forward(e2e_resblock* rb, const float* x, int64_t sB, int64_t sC, int64_t sT, int32_t B, int32_t T,
                         float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- Postnet (N2, the step right before the vocoder) ----
 * Replaces Postnet.forward (e2e_tts/models/acoustic/unsupervised_fastspeech2/layers.py:507-563) in eval mode:
 * conv_layers x [Conv1d(kernel_size, same padding) -> BatchNorm1d -> tanh (all but the last)] on [B, T, n_channels].
 * e2e_postnet_load_layer takes layer `index`'s conv weight [C_out][C_in][k] / bias with the BatchNorm1d running
 * statistics already folded in (w * g/sqrt(var+eps), (b - mean) * g/sqrt(var+eps) + beta), fp32 on the HOST.
 * x, out: device fp32 [B][T][n_channels] contiguous; add_input != 0 fuses the caller's `postnet(output) + output`
 * (model.py:188). */
typedef struct e2e_postnet e2e_postnet;
int e2e_postnet_create(int32_t n_channels, int32_t embedding_dim, int32_t conv_layers, int32_t kernel_size,
                       e2e_postnet** out);
void e2e_postnet_destroy(e2e_postnet* pn);
int e2e_postnet_load_layer(e2e_postnet* pn, int32_t index, const float* weight, int64_t weight_numel,
                           const float* bias, int64_t bias_numel);
size_t e2e_postnet_workspace_bytes(const e2e_postnet* pn, int32_t B, int32_t T);
int e2e_postnet_forward(e2e_postnet* pn, const float* x, int32_t B, int32_t T, int32_t add_input, float* out,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- mel front-end ---- */
typedef struct e2e_mel e2e_mel; /* opaque: window, twiddles and the sparse mel filterbank on the device */

/* Replaces TorchSTFT.__init__ (stft.py:12-44).  `mel_basis` is the dense [n_mels][n_fft/2+1] fp32 filterbank
 * on the HOST (the reference gets it from librosa.filters.mel, stft.py:34-40).  Supported: n_fft == win_length
 * == 1024, hop_length == 256 (the reference's preprocessing_config.yaml:5-8), n_mels <= 128. */
int e2e_mel_create(int32_t n_fft, int32_t hop_length, int32_t win_length, int32_t n_mels, const float* mel_basis,
                   e2e_mel** out);
void e2e_mel_destroy(e2e_mel* m);

/* Frames produced for L samples: 1 + (L + 2*pad - n_fft) / hop with pad = (n_fft - hop)/2 (stft.py:33,60-76). */
int64_t e2e_mel_num_frames(const e2e_mel* m, int64_t L);

/* Replaces TorchSTFT.mel_spectrogram / generate_melspecs (stft.py:46-89,107-135).  wav: device fp32, row b at
 * wav + b*ldw, L valid samples (L > pad).  mel: device fp32 [B][n_mels][T].  energy: device fp32 [B][T] or
 * NULL.  range_flag: device int32 (or NULL); set to 1 if any sample is outside [-1, 1] (the reference asserts,
 * stft.py:56-57; the caller reads the flag after the stream is synchronised and raises). */
int e2e_mel_forward(e2e_mel* m, const float* wav, int32_t B, int64_t L, int64_t ldw, float* mel, float* energy,
                    int32_t* range_flag, void* stream);

const char* e2e_last_error_string(void);

/* Library / build identification, e.g. "e2e_tts_b200 0.1 sm_100a". */
const char* e2e_version_string(void);

#ifdef __cplusplus
}
#endif
#endif /* E2E_TTS_B200_H */
