// HiFi-GAN generator on B200: layer table, weight packing, launch plans and the forward pass.
// Mirrors the structure of the reference generator (e2e_tts/models/vocoder/generator.py:13-53 and
// layers.py:10-69) as a list of tcgen05 convolution launches plus two small CUDA-core kernels.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "../../include/e2e_tts_b200.h"
#include "pair_host.cuh"
#include "rb_host.cuh"
#include "pair_tz_host.cuh"
#include "small_kernels.cuh"

using namespace e2e;

namespace {

enum LayerKind { L_CONV = 0, L_CONVT = 1, L_POST = 2 };  // (the iSTFTNet conv_post is an L_CONV padded to 32 columns)

struct Layer {
  std::string name;
  LayerKind kind;
  int cin, cout, k, dil, u;  // reference shapes (u = stride of ConvTranspose1d)
  ConvShape shape;           // GEMM form (L_CONV / L_CONVT)
  // Narrower N tilings of the same GEMM (nt = 128, 64) with their own packed images: small batches leave most SMs
  // without a unit when a 256-column layer is one N tile per 128 rows, so make_conv_op may split N across CTAs
  std::vector<ConvShape> alt_shape;
  std::vector<uint8_t*> alt_w;
  uint8_t* d_w = nullptr;    // packed 16-bit weights (or fp32 [k][cin] for L_POST)
  uint8_t* d_wtz = nullptr;  // 32 -> 32 channel, dilation-1 convolutions: sliding-window array for pair_tz.cuh
  float* d_bias = nullptr;   // [n_total]
  std::vector<float> h_bias;  // host copy (the fused pair kernel takes its biases as kernel parameters)
  float post_bias = 0.f;
  bool loaded = false;
};

struct Op {
  int kind;   // 0 = mel_to_act, 1 = conv_tc, 2 = post, 3 = fused residual pair, 4 = fused whole ResBlock1,
              // 7 = fused residual pair of the C = 32 stage, four time steps per GEMM row (pair_tz.cuh)
  int layer;  // index into layers
  ConvPlan plan;
  PairPlan pair;
  RbPlan rb;
  TzPlan tz;
};

struct PlanKey {
  int B, T;
  const void* ws;
  bool operator<(const PlanKey& o) const {
    if (B != o.B) return B < o.B;
    if (T != o.T) return T < o.T;
    return ws < o.ws;
  }
};

// CUDA-graph replay of a whole forward: the launch sequence of one (plan, input, output) combination is captured the
// second time it is seen and replayed from then on (serving loops reuse their device buffers: e2e_tts_b200.HostPipeline,
// `out=`), which takes the ~50 kernel launches of a forward off the host's critical path - what bounds single-utterance
// latency when the kernels are short (e2e_tts/src/api/utils.py:131-145 synthesises one bucket at a time).
struct GraphKey {
  int B, T;
  const void *ws, *mel, *o1, *o2, *o3;
  long long sB, sC, sT;
  float scale;
  bool operator<(const GraphKey& o) const {
    return std::memcmp(this, &o, sizeof(GraphKey)) < 0;
  }
};
struct GraphEntry {
  cudaGraphExec_t exec = nullptr;
  int seen = 0;
  bool failed = false;
};

// Activations live in HBM ONCE, as bf16 leaky_relu(x): that tensor is both the next convolution's operand and
// (through the inverse LeakyReLU) the residual x of `xt + x`.  The running resblock sum `xs` (generator.py:44-47)
// ping-pongs between two bf16 buffers.
struct Buffers {
  __nv_bfloat16 *melA, *preA, *A0, *A1, *M, *Y, *S0, *S1;
  size_t total;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

struct e2e_voc {
  e2e_voc_config cfg;
  std::vector<Layer> layers;
  std::map<std::string, int> by_name;
  std::map<PlanKey, std::vector<Op>> plans;
  std::map<GraphKey, GraphEntry> graphs;
  int cin_pad = 0;
  int hop = 1;
  int n_sms = 148;
  int f16 = 0;   // operand / activation format of the tensor-core path: 0 = bf16, 1 = fp16 (e2e_voc_set_operand_dtype)
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;  // one-shot profiling events
  int last_launches = 0;
  int last_graph = 0;   // 1: the last forward was replayed from a captured CUDA graph
  cudaStream_t cap_stream = nullptr;   // capture-only stream (never executes anything)
  PostWeights<7 * 32> post_w{};                      // conv_post weights [k][C] for the 32-channel, k = 7 kernel
};

// Launch plans (and the graphs captured from them) embed biases and weight pointers: forget them when weights change.
static void drop_plans(e2e_voc* v) {
  v->plans.clear();
  for (auto& kv : v->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  v->graphs.clear();
}

static int pick_nt(int cout) { return cout >= 256 ? 256 : cout; }

static void add_conv(e2e_voc* v, const std::string& name, int cin, int cin_pad, int cout, int k, int dil,
                     int cout_pad = 0) {
  if (cout_pad < cout) cout_pad = cout;  // GEMM columns (zero weights / bias beyond cout)
  Layer L;
  L.name = name;
  L.kind = L_CONV;
  L.cin = cin;
  L.cout = cout;
  L.k = k;
  L.dil = dil;
  L.u = 1;
  L.shape.cin = cin_pad;
  L.shape.n_total = cout_pad;
  L.shape.nt = pick_nt(cout_pad);
  L.shape.taps = k;
  const int n_tiles = cout_pad / L.shape.nt;
  for (int i = 0; i < n_tiles; ++i)
    for (int j = 0; j < k; ++j) L.shape.shifts.push_back((j - (k - 1) / 2) * dil);
  if (L.shape.nt == 256)
    for (int nt : {128, 64}) {
      if (cout_pad / nt > kMaxNTiles) continue;
      ConvShape a = L.shape;
      a.nt = nt;
      a.shifts.clear();
      for (int i = 0; i < cout_pad / nt; ++i)
        for (int j = 0; j < k; ++j) a.shifts.push_back((j - (k - 1) / 2) * dil);
      L.alt_shape.push_back(a);
      L.alt_w.push_back(nullptr);
    }
  v->by_name[name] = (int)v->layers.size();
  v->layers.push_back(L);
}

// ConvTranspose1d(cin, cout, k = 2u, stride = u, padding = u/2) in polyphase form (SURVEY.md §8 a'5):
// output sample n = q*u + p, j0 = p + u/2:  y = W[:, :, j0]^T x[q] + (j0 < u ? W[:, :, j0+u]^T x[q-1]
//                                                                        : W[:, :, j0-u]^T x[q+1]).
// GEMM column n' = p*cout + co, so row q of the GEMM output IS the u output samples in channels-last order.
static int add_convt(e2e_voc* v, const std::string& name, int cin, int cout, int k, int u) {
  if (k != 2 * u || (u & 1)) return fail(-4, "ConvTranspose1d supported for kernel == 2*stride, even stride");
  Layer L;
  L.name = name;
  L.kind = L_CONVT;
  L.cin = cin;
  L.cout = cout;
  L.k = k;
  L.dil = 1;
  L.u = u;
  L.shape.cin = cin;
  L.shape.n_total = u * cout;
  int nt = cout;  // one phase per tile, widened while the tile stays inside one half of the phases
  while (nt * 2 <= 256 && ((u / 2) % (nt * 2 / cout)) == 0) nt *= 2;
  if (nt > 256) return fail(-4, "ConvTranspose1d output channels > 256 unsupported");
  L.shape.nt = nt;
  L.shape.taps = 2;
  const int n_tiles = L.shape.n_total / nt;
  for (int i = 0; i < n_tiles; ++i) {
    const int p = (i * nt) / cout;
    L.shape.shifts.push_back(0);
    L.shape.shifts.push_back(p < u / 2 ? -1 : +1);
  }
  v->by_name[name] = (int)v->layers.size();
  v->layers.push_back(L);
  return 0;
}

extern "C" int e2e_voc_create(const e2e_voc_config* cfg, e2e_voc** out) {
  if (!cfg || !out) return fail(-1, "null argument");
  if (cfg->num_upsamples < 1 || cfg->num_upsamples > E2E_MAX_UPSAMPLES || cfg->num_kernels < 1 ||
      cfg->num_kernels > E2E_MAX_KERNELS)
    return fail(-1, "bad num_upsamples / num_kernels");
  std::unique_ptr<e2e_voc> v(new e2e_voc);
  v->cfg = *cfg;
  const int C0 = cfg->upsample_initial_channel;
  if (C0 % 64 != 0 || C0 > 512) return fail(-4, "upsample_initial_channel must be a multiple of 64, <= 512");
  if (cfg->in_channels < 1 || cfg->in_channels > 512) return fail(-4, "in_channels out of range");
  v->cin_pad = (cfg->in_channels + 63) / 64 * 64;
  add_conv(v.get(), "conv_pre", cfg->in_channels, v->cin_pad, C0, 7, 1);
  int ch = C0;
  v->hop = 1;
  for (int i = 0; i < cfg->num_upsamples; ++i) {
    const int cout = ch / 2;
    if (cout != 32 && cout % 64 != 0) return fail(-4, "stage channel count must be 32 or a multiple of 64");
    int rc = add_convt(v.get(), "ups." + std::to_string(i), ch, cout, cfg->upsample_kernel_sizes[i],
                       cfg->upsample_rates[i]);
    if (rc) return rc;
    v->hop *= cfg->upsample_rates[i];
    ch = cout;
  }
  ch = C0;
  for (int i = 0; i < cfg->num_upsamples; ++i) {
    ch /= 2;
    for (int j = 0; j < cfg->num_kernels; ++j) {
      const int k = cfg->resblock_kernel_sizes[j];
      if (!(k & 1) || k > kMaxTaps) return fail(-4, "resblock kernel size must be odd and <= 15");
      const int nd = cfg->num_dilations[j];
      if (nd < 1 || nd > E2E_MAX_DILATIONS) return fail(-1, "bad num_dilations");
      const std::string base = "resblocks." + std::to_string(i * cfg->num_kernels + j);
      for (int m = 0; m < nd; ++m) {
        const int d = cfg->resblock_dilation_sizes[j][m];
        if ((k - 1) / 2 * d > 127) return fail(-4, "dilated receptive field too wide");
        if (cfg->resblock == 1) {
          add_conv(v.get(), base + ".convs1." + std::to_string(m), ch, ch, ch, k, d);
          add_conv(v.get(), base + ".convs2." + std::to_string(m), ch, ch, ch, k, 1);
        } else {
          add_conv(v.get(), base + ".convs." + std::to_string(m), ch, ch, ch, k, d);
        }
      }
    }
  }
  if (cfg->istft_n_fft > 0) {
    // iSTFTNet head: conv_post C -> n_fft + 2, k = 7 (generator.py:86) on the tensor cores, columns padded to 32
    const int n = cfg->istft_n_fft;
    if (n < 4 || n > 62 || (n & 1)) return fail(-4, "gen_istft_n_fft must be even, in [4, 62]");
    if (ch != 32 && ch % 64 != 0) return fail(-4, "conv_post input channels unsupported");
    add_conv(v.get(), "conv_post", ch, ch, n + 2, 7, 1, (n + 2 + 31) / 32 * 32);
  } else {
    Layer L;
    L.name = "conv_post";
    L.kind = L_POST;
    L.cin = ch;
    L.cout = 1;
    L.k = 7;
    L.dil = 1;
    L.u = 1;
    if (ch * 7 > kPostMaxW || ch % 8) return fail(-4, "conv_post input channels unsupported");
    v->by_name[L.name] = (int)v->layers.size();
    v->layers.push_back(L);
  }
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail((int)e, std::string("cudaGetDevice: ") + cudaGetErrorString(e));
  cudaDeviceGetAttribute(&v->n_sms, cudaDevAttrMultiProcessorCount, dev);
  if (v->n_sms < 1) v->n_sms = 148;
  int rc_init = conv_kernels_init();
  if (rc_init) return rc_init;
  rc_init = pair_kernels_init();
  if (rc_init) return rc_init;
  rc_init = rb_kernels_init();
  if (rc_init) return rc_init;
  rc_init = tz_kernels_init();
  if (rc_init) return rc_init;
  *out = v.release();
  return 0;
}

extern "C" void e2e_voc_destroy(e2e_voc* v) {
  if (!v) return;
  for (auto& L : v->layers) {
    if (L.d_w) cudaFree(L.d_w);
    if (L.d_wtz) cudaFree(L.d_wtz);
    if (L.d_bias) cudaFree(L.d_bias);
    for (uint8_t* w : L.alt_w)
      if (w) cudaFree(w);
  }
  for (auto& kv : v->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  if (v->cap_stream) cudaStreamDestroy(v->cap_stream);
  delete v;
}

extern "C" int e2e_voc_set_profile_events(e2e_voc* v, void* ev_begin, void* ev_end) {
  if (!v) return fail(-1, "null argument");
  v->ev_begin = reinterpret_cast<cudaEvent_t>(ev_begin);
  v->ev_end = reinterpret_cast<cudaEvent_t>(ev_end);
  return 0;
}

extern "C" int e2e_voc_hop(const e2e_voc* v) { return v ? v->hop : 0; }

extern "C" int e2e_voc_set_operand_dtype(e2e_voc* v, int32_t dtype) {
  if (!v) return fail(-1, "null argument");
  if (dtype != E2E_OPERAND_BF16 && dtype != E2E_OPERAND_FP16) return fail(-1, "operand dtype must be E2E_OPERAND_BF16 or E2E_OPERAND_FP16");
  if (v->f16 == dtype) return 0;
  v->f16 = dtype;
  for (auto& L : v->layers) L.loaded = false;   // the packed weight images depend on the format: reload every layer
  drop_plans(v);
  return 0;
}

extern "C" int e2e_voc_operand_dtype(const e2e_voc* v) { return v ? v->f16 : -1; }

extern "C" int e2e_voc_last_forward_was_graph(const e2e_voc* v) { return v ? v->last_graph : -1; }

extern "C" int e2e_voc_missing_layers(const e2e_voc* v) {
  if (!v) return -1;
  int n = 0;
  for (auto& L : v->layers) n += L.loaded ? 0 : 1;
  return n;
}

extern "C" int e2e_voc_load_layer(e2e_voc* v, const char* name, const float* weight, int64_t weight_numel,
                                  const float* bias, int64_t bias_numel) {
  if (!v || !name || !weight || !bias) return fail(-1, "null argument");
  auto it = v->by_name.find(name);
  if (it == v->by_name.end()) return fail(-5, std::string("unknown layer: ") + name);
  Layer& L = v->layers[it->second];
  const int64_t want_w = (int64_t)L.cin * L.cout * L.k;
  if (weight_numel != want_w || bias_numel != L.cout)
    return fail(-6, std::string("shape mismatch for layer ") + name);
  cudaError_t e;
  // A reload while a forward enqueued on a non-blocking stream is still reading the old weight image would tear it
  // (the copies below run on the legacy default stream, which does not order against such streams): drain the device.
  if (L.d_w && (e = cudaDeviceSynchronize()) != cudaSuccess) return fail((int)e, "cudaDeviceSynchronize");
  if (L.kind == L_POST) {
    // reference layout [1][cin][k] -> [k][cin]
    std::vector<float> w((size_t)L.k * L.cin);
    for (int c = 0; c < L.cin; ++c)
      for (int j = 0; j < L.k; ++j) w[(size_t)j * L.cin + c] = weight[(size_t)c * L.k + j];
    if (!L.d_w && (e = cudaMalloc(&L.d_w, w.size() * 4)) != cudaSuccess) return fail((int)e, "cudaMalloc");
    if ((e = cudaMemcpy(L.d_w, w.data(), w.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess)
      return fail((int)e, "cudaMemcpy");
    L.post_bias = bias[0];
    if (w.size() == sizeof(v->post_w.w) / sizeof(float)) memcpy(v->post_w.w, w.data(), sizeof(v->post_w.w));
    L.loaded = true;
    drop_plans(v);
    return 0;
  }
  const ConvShape& s = L.shape;
  std::vector<float> wg((size_t)s.n_total * s.taps * s.cin, 0.f);
  std::vector<float> bg(s.n_total);
  if (L.kind == L_CONV) {
    // [cout][cin][k] -> wg[co][tap][ci]
    for (int co = 0; co < L.cout; ++co) {
      for (int ci = 0; ci < L.cin; ++ci)
        for (int j = 0; j < L.k; ++j)
          wg[((size_t)co * s.taps + j) * s.cin + ci] = weight[((size_t)co * L.cin + ci) * L.k + j];
      bg[co] = bias[co];
    }
  } else {
    // [cin][cout][2u] -> column n = p*cout + co, tap 0 = W[..., j0], tap 1 = W[..., j0 +/- u]
    const int u = L.u;
    for (int p = 0; p < u; ++p) {
      const int j0 = p + u / 2;
      const int j1 = j0 < u ? j0 + u : j0 - u;
      for (int co = 0; co < L.cout; ++co) {
        const int n = p * L.cout + co;
        for (int ci = 0; ci < L.cin; ++ci) {
          const float* wsrc = weight + ((size_t)ci * L.cout + co) * L.k;
          wg[((size_t)n * 2 + 0) * s.cin + ci] = wsrc[j0];
          wg[((size_t)n * 2 + 1) * s.cin + ci] = wsrc[j1];
        }
        bg[n] = bias[co];
      }
    }
  }
  std::vector<uint8_t> packed(packed_weight_bytes(s));
  pack_conv_weights(s, wg.data(), packed.data(), v->f16);
  if (!L.d_w && (e = cudaMalloc(&L.d_w, packed.size())) != cudaSuccess) return fail((int)e, "cudaMalloc");
  if (!L.d_bias && (e = cudaMalloc(&L.d_bias, bg.size() * 4)) != cudaSuccess) return fail((int)e, "cudaMalloc");
  if ((e = cudaMemcpy(L.d_w, packed.data(), packed.size(), cudaMemcpyHostToDevice)) != cudaSuccess)
    return fail((int)e, "cudaMemcpy");
  if ((e = cudaMemcpy(L.d_bias, bg.data(), bg.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess)
    return fail((int)e, "cudaMemcpy");
  for (size_t a = 0; a < L.alt_shape.size(); ++a) {
    std::vector<uint8_t> pk(packed_weight_bytes(L.alt_shape[a]));
    pack_conv_weights(L.alt_shape[a], wg.data(), pk.data(), v->f16);
    if (!L.alt_w[a] && (e = cudaMalloc(&L.alt_w[a], pk.size())) != cudaSuccess) return fail((int)e, "cudaMalloc");
    if ((e = cudaMemcpy(L.alt_w[a], pk.data(), pk.size(), cudaMemcpyHostToDevice)) != cudaSuccess)
      return fail((int)e, "cudaMemcpy");
  }
  if (L.kind == L_CONV && L.cin == kTzC && L.cout == kTzC && s.cin == kTzC && s.n_total == kTzC && L.dil == 1 && L.k >= 3) {
    // the same weights as a sliding-window array (pair_tz.cuh): c2 of every C = 32 pair, and c1 where it is undilated
    std::vector<uint8_t> win(tz_window_bytes(L.k));
    pack_tz_window(wg.data(), L.k, win.data(), v->f16);
    if (!L.d_wtz && (e = cudaMalloc(&L.d_wtz, win.size())) != cudaSuccess) return fail((int)e, "cudaMalloc");
    if ((e = cudaMemcpy(L.d_wtz, win.data(), win.size(), cudaMemcpyHostToDevice)) != cudaSuccess)
      return fail((int)e, "cudaMemcpy");
  }
  L.h_bias = bg;
  L.loaded = true;
  drop_plans(v);
  return 0;
}

static size_t stage_elems(const e2e_voc* v, int B, int T) {
  // largest rows*channels product over the resblock stages
  size_t best = 0;
  int ch = v->cfg.upsample_initial_channel, rate = 1;
  for (int i = 0; i < v->cfg.num_upsamples; ++i) {
    ch /= 2;
    rate *= v->cfg.upsample_rates[i];
    const size_t e = (size_t)B * T * rate * ch;
    best = e > best ? e : best;
  }
  return best;
}

static void carve(const e2e_voc* v, int B, int T, void* ws, Buffers& b) {
  uint8_t* p = reinterpret_cast<uint8_t*>(ws);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* r = p ? p + off : nullptr;
    off += align_up(bytes, 1024);
    return r;
  };
  // + one row per utterance: the iSTFTNet head reflection-pads the last stage's tensor (generator.py:102)
  // and 7 rows: the running-sum tensors S0 / S1 are tiled in 8-row blocks (epilogue.cuh)
  const size_t E = stage_elems(v, B, T) + (size_t)B * 8 * v->cfg.upsample_initial_channel;
  b.melA = (__nv_bfloat16*)take((size_t)B * T * v->cin_pad * 2);
  b.preA = (__nv_bfloat16*)take((size_t)B * T * v->cfg.upsample_initial_channel * 2);
  b.A0 = (__nv_bfloat16*)take(E * 2);
  b.A1 = (__nv_bfloat16*)take(E * 2);
  b.M = (__nv_bfloat16*)take(E * 2);
  b.Y = (__nv_bfloat16*)take(E * 2);
  b.S0 = (__nv_bfloat16*)take(E * 2);
  b.S1 = (__nv_bfloat16*)take(E * 2);
  b.total = off;
}

extern "C" size_t e2e_voc_workspace_bytes(const e2e_voc* v, int32_t B, int32_t T) {
  if (!v || B < 1 || T < 1) return 0;
  Buffers b;
  carve(v, B, T, nullptr, b);
  return b.total;
}

// One conv launch: input activation `in` ([B][T][cin] bf16), outputs as requested.
enum : int { kSumTiled = 1, kOutTiled = 2 };  // which epilogue-only tensors of an op are in the tiled8 layout

static int make_conv_op(e2e_voc* v, std::vector<Op>& ops, int layer, int B, int T, const __nv_bfloat16* in,
                        const __nv_bfloat16* res_act, const __nv_bfloat16* sum_a, float* out_f32, __nv_bfloat16* out_act,
                        float slope, float divisor, int tiled = 0) {
  Layer& L = v->layers[layer];
  Op op;
  op.kind = 1;
  op.layer = layer;
  // N tiling: the layer's natural one, or - when that leaves most SMs idle - a narrower one (more units, cheaper MMAs).
  // Cost model per 128 rows: rounds of the persistent grid x cycles per MMA (N <= 64: 54, shared-memory A-operand bound;
  // N = 128: 64; N = 256: 128; profiles/r01_umma_rate_microbench.log) x N tiles a CTA walks per unit.
  int pick = -1;
  {
    static const char* esp = std::getenv("E2E_NO_NSPLIT");
    auto cost = [&](const ConvShape& cs) {
      const long long units = (long long)((T + 127) / 128) * (cs.n_total / cs.nt) * B;
      const long long rounds = (units + v->n_sms - 1) / v->n_sms;
      return (double)rounds * (cs.nt >= 256 ? 128.0 : (cs.nt >= 128 ? 64.0 : 54.0));
    };
    double best = cost(L.shape);
    for (size_t a = 0; a < L.alt_shape.size() && !esp; ++a) {
      const double c = cost(L.alt_shape[a]);
      if (L.alt_w[a] && c < best * 0.9) {
        best = c;
        pick = (int)a;
      }
    }
  }
  const ConvShape& s = pick >= 0 ? L.alt_shape[pick] : L.shape;
  uint8_t* const w_img = pick >= 0 ? L.alt_w[pick] : L.d_w;
  // Units per CTA of the persistent grid are quantised: pick the tile height (128*mt rows) with the best
  // balance, preferring taller tiles (fewer weight re-streams per row) when the balance is within 6 %.
  const int n_tiles = s.n_total / s.nt;
  int mt = 1;
  double best = 0.0;
  for (int cand = 1; cand <= 4; cand <<= 1) {
    if (2 * cand * s.nt > 512 && cand > 1) break;  // keep two TMEM accumulator sets (epilogue overlaps MMAs)
    const long long units = (long long)((T + 128 * cand - 1) / (128 * cand)) * n_tiles * B;
    const long long rounds = (units + v->n_sms - 1) / v->n_sms;
    // useful rows / rows the busiest CTA pays for
    const double eff = (double)T * B * n_tiles / ((double)rounds * v->n_sms * 128.0 * cand);
    if (eff >= best * 0.94) {
      if (eff > best) best = eff;
      mt = cand;
    }
  }
  // Staged TMA stores of the bf16 output (conv_tc.cuh, STAGED): natural-layout bf16 outputs only, and only where
  // the epilogue rather than the MMAs bounds the layer - few taps (k = 3 convolutions, the two-tap polyphase
  // upsamplers).  Measured per launch, 16 x 5 s: C = 256 k = 3 28.3 -> 25.5 us, ups.1 75.6 -> 64.7 us, k >= 7 unchanged,
  // conv_pre 16.5 -> 18.5 us.  E2E_CONV_STAGED=0|1 forces it off / on wherever it is possible.
  static const char* est = std::getenv("E2E_CONV_STAGED");
  const bool can_stage = out_act && !out_f32 && !(tiled & kOutTiled);
  const int staged = can_stage && (est ? est[0] == '1' : s.taps <= 3);
  int rc = plan_conv(op.plan, s, B, T, mt, v->n_sms, 0, staged);
  if (rc) return rc;
  ConvParams& p = op.plan.p;
  rc = conv_output_map(op.plan, out_act, B, T);
  if (rc) return rc;
  rc = make_act_tensor_map(&op.plan.tm, in, B, T, s.cin, p.rowb / 2, p.box_rows);
  if (rc) return rc;
  p.w = w_img;
  rc = conv_weight_map(op.plan, w_img);
  if (rc) return rc;
  conv_set_bias(op.plan, L.d_bias, L.h_bias.data(), (int)L.h_bias.size(), L.kind == L_CONVT ? L.cout : 0);
  p.res_act = res_act;
  p.res_inv_slope = 10.0f;  // 1 / LRELU_SLOPE: every residual tensor was written with slope 0.1
  p.sum_a = sum_a;
  p.out_f32 = out_f32;
  p.out_act = out_act;
  p.slope = slope;
  p.divisor = divisor;
  p.sum_tiled = sum_a != nullptr && (tiled & kSumTiled);
  p.out_tiled = out_act != nullptr && (tiled & kOutTiled);
  p.f16 = v->f16;
  ops.push_back(op);
  return 0;
}

// One fused launch for  x + c2(lrelu(c1(lrelu(x))))  (pair_tc.cuh).  `in` holds bf16 leaky_relu(x, 0.1).
static int make_pair_op(e2e_voc* v, std::vector<Op>& ops, int l1, int l2, int B, int T, const __nv_bfloat16* in,
                        const __nv_bfloat16* sum_a, float* out_f32, __nv_bfloat16* out_act, float slope, float divisor,
                        int tiled = 0) {
  const Layer& L1 = v->layers[l1];
  const Layer& L2 = v->layers[l2];
  Op op;
  op.kind = 3;
  op.layer = l1;
  int rc = plan_pair(op.pair, L1.cin, L1.k, L1.dil, B, T, v->n_sms);
  if (rc) return rc;
  PairParams& p = op.pair.p;
  rc = make_act_tensor_map(&op.pair.tm, in, B, T, L1.cin, op.pair.rowb / 2, p.box_rows);
  if (rc) return rc;
  p.w1 = L1.d_w;
  p.w2 = L2.d_w;
  rc = pair_weight_maps(op.pair, L1.d_w, L2.d_w);
  if (rc) return rc;
  if (L1.h_bias.size() > 128 || L2.h_bias.size() > 128) return fail(-2, "fused pair: more than 128 channels");
  std::copy(L1.h_bias.begin(), L1.h_bias.end(), p.bias1);
  std::copy(L2.h_bias.begin(), L2.h_bias.end(), p.bias2);
  p.res_act = in;
  p.res_inv_slope = 10.0f;
  p.sum_a = sum_a;
  if (out_f32 || !out_act) return fail(-2, "fused pair: bf16 activation output only");
  p.out_act = out_act;
  p.sum_tiled = sum_a != nullptr && (tiled & kSumTiled);
  p.out_tiled = (tiled & kOutTiled) != 0;
  p.no_sum_prefetch = std::getenv("E2E_NO_SUM_PREFETCH") != nullptr;
  if (p.out_tiled) op.pair.staged = false;   // the TMA store writes the natural layout
  rc = pair_output_maps(op.pair, out_act, B, T, L1.cin);
  if (rc) return rc;
  p.slope_mid = 0.1f;
  p.slope = slope;
  p.divisor = divisor;
  p.f16 = v->f16;
  ops.push_back(op);
  return 0;
}

// The same pair on the C = 32 stage with four time steps per GEMM row (pair_tz.cuh).
static bool tz_usable(const e2e_voc* v, int l1, int l2, int T) {
  const Layer& L1 = v->layers[l1];
  const Layer& L2 = v->layers[l2];
  if (L1.cin != kTzC || L1.k != L2.k || L2.dil != 1 || !L2.d_wtz || (L1.dil == 1 && !L1.d_wtz)) return false;
  return tz_supported(L1.cin, L1.k, L1.dil, T);
}

static int make_tz_op(e2e_voc* v, std::vector<Op>& ops, int l1, int l2, int B, int T, const __nv_bfloat16* in,
                      const __nv_bfloat16* sum_a, __nv_bfloat16* out_act, float slope, float divisor, int tiled) {
  const Layer& L1 = v->layers[l1];
  const Layer& L2 = v->layers[l2];
  Op op;
  op.kind = 7;
  op.layer = l1;
  int rc = plan_tz(op.tz, L1.k, L1.dil, B, T, v->n_sms);
  if (rc) return rc;
  TzParams& p = op.tz.p;
  rc = tz_input_map(op.tz, in);
  if (rc) return rc;
  p.w1 = L1.dil == 1 ? L1.d_wtz : L1.d_w;
  p.w2 = L2.d_wtz;
  if (L1.h_bias.size() != (size_t)kTzC || L2.h_bias.size() != (size_t)kTzC) return fail(-2, "pair_tz: 32 channels only");
  std::copy(L1.h_bias.begin(), L1.h_bias.end(), p.bias1);
  std::copy(L2.h_bias.begin(), L2.h_bias.end(), p.bias2);
  p.res_inv_slope = 10.0f;
  p.sum_a = sum_a;
  p.sum_tiled = sum_a != nullptr && (tiled & kSumTiled);
  p.out_tiled = (tiled & kOutTiled) != 0;
  p.no_sum_prefetch = std::getenv("E2E_NO_SUM_PREFETCH") != nullptr;
  p.out_act = out_act;
  p.slope_mid = 0.1f;
  p.slope = slope;
  p.divisor = divisor;
  p.f16 = v->f16;
  ops.push_back(op);
  return 0;
}

// One fused launch for a whole ResBlock1 (rb_tc.cuh): `in` holds bf16 leaky_relu(x, 0.1); l1[i] / l2[i] = layers of pair i.
static int make_rb_op(e2e_voc* v, std::vector<Op>& ops, const int* l1, const int* l2, int n_pairs, int B, int T,
                      const __nv_bfloat16* in, const __nv_bfloat16* sum_a, __nv_bfloat16* out_act, float slope,
                      float divisor, int tiled) {
  const Layer& L0 = v->layers[l1[0]];
  Op op;
  op.kind = 4;
  op.layer = l1[0];
  int dil[kRbMaxPairs] = {1, 1, 1};
  for (int i = 0; i < n_pairs; ++i) dil[i] = v->layers[l1[i]].dil;
  int rc = plan_rb(op.rb, L0.cin, L0.k, dil, n_pairs, B, T, v->n_sms);
  if (rc) return rc;
  RbParams& p = op.rb.p;
  rc = make_act_tensor_map(&op.rb.tm, in, B, T, L0.cin, op.rb.rowb / 2, p.box_rows);
  if (rc) return rc;
  std::vector<float> cum(L0.cin, 0.f);
  for (int i = 0; i < n_pairs; ++i) {
    const Layer& A = v->layers[l1[i]];
    const Layer& C2 = v->layers[l2[i]];
    if (A.h_bias.size() > 128 || C2.h_bias.size() > 128) return fail(-2, "fused resblock: more than 128 channels");
    p.w[2 * i] = A.d_w;
    p.w[2 * i + 1] = C2.d_w;
    std::copy(A.h_bias.begin(), A.h_bias.end(), p.bias[2 * i]);
    for (size_t n = 0; n < C2.h_bias.size(); ++n) {   // x lives in TMEM without the c2 biases: pass their running sum
      cum[n] += C2.h_bias[n];
      p.bias[2 * i + 1][n] = cum[n];
    }
  }
  p.res_inv_slope = 10.0f;
  p.sum_a = sum_a;
  p.sum_tiled = sum_a != nullptr && (tiled & kSumTiled);
  p.out_tiled = (tiled & kOutTiled) != 0;
  p.out_act = out_act;
  p.slope_mid = 0.1f;
  p.slope = slope;
  p.divisor = divisor;
  p.f16 = v->f16;
  ops.push_back(op);
  return 0;
}

static int build_plan(e2e_voc* v, int B, int T, void* ws, std::vector<Op>& ops) {
  Buffers bf;
  carve(v, B, T, ws, bf);
  // S0 / S1 (the running resblock sums) are written and read by epilogues only, so they live in the tiled8
  // layout (epilogue.cuh tiled8_off); E2E_NO_TILED_SUMS=1 keeps them in the natural one (A/B switch).
  const bool tiled_sums = std::getenv("E2E_NO_TILED_SUMS") == nullptr;
  const e2e_voc_config& c = v->cfg;
  const float kSlope = 0.1f;  // LRELU_SLOPE, generator.py:10 / layers.py:7
  {
    Op op;
    op.kind = 0;
    op.layer = 0;
    ops.push_back(op);
  }
  // conv_pre, then the first leaky_relu of the stage loop (generator.py:38-40)
  int rc = make_conv_op(v, ops, v->by_name["conv_pre"], B, T, bf.melA, nullptr, nullptr, nullptr, bf.preA, kSlope, 0.f);
  if (rc) return rc;
  const __nv_bfloat16* stage_in = bf.preA;
  int Ts = T;
  for (int i = 0; i < c.num_upsamples; ++i) {
    // x = ups[i](leaky_relu(x)) : writes A0 = bf16 leaky_relu(x, 0.1), the stage's shared input
    rc = make_conv_op(v, ops, v->by_name["ups." + std::to_string(i)], B, Ts, stage_in, nullptr, nullptr, nullptr,
                      bf.A0, kSlope, 0.f);
    if (rc) return rc;
    Ts *= c.upsample_rates[i];
    const bool last_stage = i + 1 == c.num_upsamples;
    // F.leaky_relu(x) before conv_post uses the default slope 0.01 (generator.py:49)
    const float out_slope = last_stage ? 0.01f : kSlope;
    for (int j = 0; j < c.num_kernels; ++j) {
      const std::string base = "resblocks." + std::to_string(i * c.num_kernels + j);
      const int nd = c.num_dilations[j];
      const __nv_bfloat16* ain = bf.A0;
      const int chs = v->layers[v->by_name[base + (c.resblock == 1 ? ".convs1.0" : ".convs.0")]].cin;
      const int ks = c.resblock_kernel_sizes[j];
      // ResBlock1 pairs run fused when the channel count allows (stages with C <= 128); the fused kernel reads
      // halo rows of its input from neighbouring tiles, so it ping-pongs between two output buffers (A1, M)
      // instead of updating in place.
      bool fused = c.resblock == 1 && std::getenv("E2E_NO_PAIR_FUSION") == nullptr;
      for (int m = 0; m < nd && fused; ++m) fused = pair_supported(chs, ks, c.resblock_dilation_sizes[j][m]);
      // Whole-resblock fusion (rb_tc.cuh: residual stream in TMEM, one launch per ResBlock1) where its halo is cheap
      // (E2E_TZ_K3=1: experiment - the k = 3 resblock of the C = 32 stage as three pair_tz launches instead of the chain)
      static const char* etz3 = std::getenv("E2E_TZ_K3");
      const bool tz_instead = etz3 && etz3[0] == '1' && chs == kTzC && tz_supported(chs, ks, 1, Ts);
      if (fused && !tz_instead && nd <= kRbMaxPairs && rb_supported(chs, ks, c.resblock_dilation_sizes[j], nd)) {
        int l1[kRbMaxPairs], l2[kRbMaxPairs];
        for (int m = 0; m < nd; ++m) {
          l1[m] = v->by_name[base + ".convs1." + std::to_string(m)];
          l2[m] = v->by_name[base + ".convs2." + std::to_string(m)];
        }
        // where the resblock's result goes: the running sum xs (bf16, S0 / S1 ping-pong) or, for the stage's last
        // resblock, x = xs / num_kernels as the next layer's activation (same rules as the pair path below)
        __nv_bfloat16* oact = (j & 1) ? bf.S1 : bf.S0;
        float slope = 1.0f, divisor = 0.f;
        const __nv_bfloat16* sum_in = j > 0 ? ((j & 1) ? bf.S0 : bf.S1) : nullptr;
        int tiled = 0;
        if (tiled_sums) tiled = kSumTiled | (j + 1 == c.num_kernels ? 0 : kOutTiled);
        if (j + 1 == c.num_kernels) {
          oact = bf.Y;
          divisor = (float)c.num_kernels;
          slope = out_slope;
        }
        rc = make_rb_op(v, ops, l1, l2, nd, B, Ts, bf.A0, sum_in, oact, slope, divisor, tiled);
        if (rc) return rc;
        continue;
      }
      for (int m = 0; m < nd; ++m) {
        const bool last = m + 1 == nd;
        // where does x_new = conv(...) + x go?
        float* of32 = nullptr;
        __nv_bfloat16* oact = fused ? ((m & 1) ? bf.M : bf.A1) : bf.A1;
        const __nv_bfloat16* sum_in = nullptr;
        float divisor = 0.f, slope = kSlope;
        int tiled = 0;
        if (last) {
          // xs += resblock_j(x) (generator.py:44-47): the running sum is kept in bf16 (slope 1 = no activation),
          // alternating between S0 and S1 so no launch reads and writes the same buffer
          oact = (j & 1) ? bf.S1 : bf.S0;
          slope = 1.0f;
          sum_in = j > 0 ? ((j & 1) ? bf.S0 : bf.S1) : nullptr;
          if (tiled_sums) tiled = kSumTiled | (j + 1 == c.num_kernels ? 0 : kOutTiled);
          if (j + 1 == c.num_kernels) {           // x = xs / num_kernels   (generator.py:48)
            oact = bf.Y;
            divisor = (float)c.num_kernels;
            slope = out_slope;
          }
        }
        if (fused) {
          const int l1 = v->by_name[base + ".convs1." + std::to_string(m)];
          const int l2 = v->by_name[base + ".convs2." + std::to_string(m)];
          if (!of32 && tz_usable(v, l1, l2, Ts))
            rc = make_tz_op(v, ops, l1, l2, B, Ts, ain, sum_in, oact, slope, divisor, tiled);
          else
            rc = make_pair_op(v, ops, l1, l2, B, Ts, ain, sum_in, of32, oact, slope, divisor, tiled);
          if (rc) return rc;
          ain = oact;
        } else if (c.resblock == 1) {
          rc = make_conv_op(v, ops, v->by_name[base + ".convs1." + std::to_string(m)], B, Ts, ain, nullptr, nullptr,
                            nullptr, bf.M, kSlope, 0.f);
          if (rc) return rc;
          rc = make_conv_op(v, ops, v->by_name[base + ".convs2." + std::to_string(m)], B, Ts, bf.M, ain, sum_in,
                            of32, oact, slope, divisor, tiled);
          if (rc) return rc;
          ain = bf.A1;
        } else {
          rc = make_conv_op(v, ops, v->by_name[base + ".convs." + std::to_string(m)], B, Ts, ain, ain, sum_in, of32,
                            oact, slope, divisor, tiled);
          if (rc) return rc;
          ain = bf.A1;
        }
      }
    }
    stage_in = bf.Y;
  }
  if (c.istft_n_fft > 0) {
    // iSTFTNet head (generator.py:101-106): ReflectionPad1d((1, 0)) -> conv_post on T_s + 1 rows -> exp / sin.
    // A0 holds the padded tensor, A1 (reinterpreted) conv_post's fp32 output [B][T_s + 1][32k].
    Op pad;
    pad.kind = 5;
    pad.layer = 0;
    ops.push_back(pad);
    const int lp = v->by_name["conv_post"];
    if ((size_t)v->layers[lp].shape.n_total * 4 > (size_t)v->layers[lp].cin * 2)
      return fail(-4, "iSTFTNet head: conv_post output does not fit the scratch tensor");
    rc = make_conv_op(v, ops, lp, B, Ts + 1, bf.A0, nullptr, nullptr, reinterpret_cast<float*>(bf.A1), nullptr, 1.0f,
                      0.f);
    if (rc) return rc;
    Op fin;
    fin.kind = 6;
    fin.layer = lp;
    ops.push_back(fin);
    return 0;
  }
  {
    Op op;
    op.kind = 2;
    op.layer = v->by_name["conv_post"];
    ops.push_back(op);
  }
  return 0;
}

extern "C" int e2e_voc_launches_per_forward(const e2e_voc* v) {
  if (!v) return -1;
  if (v->last_launches > 0) return v->last_launches;  // what the last forward actually enqueued
  const e2e_voc_config& c = v->cfg;
  int n = 3;  // mel_to_act, conv_pre, conv_post (unfused estimate before the first forward)
  for (int i = 0; i < c.num_upsamples; ++i) {
    n += 1;
    for (int j = 0; j < c.num_kernels; ++j) n += c.num_dilations[j] * (c.resblock == 1 ? 2 : 1);
  }
  return n;
}

struct SpecOut {
  float* spec = nullptr;
  float* phase = nullptr;
};

static int voc_forward_impl(e2e_voc* v, const float* mel, int64_t sB, int64_t sC, int64_t sT, int32_t B, int32_t T,
                            const PostOut& post, const SpecOut& so, void* workspace, size_t workspace_bytes,
                            void* stream) {
  if (!v || !mel || !workspace) return fail(-1, "null argument");
  if (v->cfg.istft_n_fft > 0) {
    if (!so.spec || !so.phase) return fail(-1, "this generator has the iSTFTNet head: call e2e_voc_forward_spec");
  } else if (!post.wav && !post.pcm) {
    return fail(-1, "this generator has the HiFi-GAN head: call e2e_voc_forward / e2e_voc_forward_pcm16");
  }
  if (B < 1 || T < 1) return fail(-1, "B and T must be positive");
  if (B > 65535) return fail(-1, "at most 65535 utterances per call (split the batch)");
  if ((long long)T * v->hop > 0x7fffffffLL) return fail(-1, "utterance too long");
  if (e2e_voc_missing_layers(v) != 0) return fail(-7, "e2e_voc_forward before all layers were loaded");
  if (reinterpret_cast<uintptr_t>(workspace) % 1024) return fail(-1, "workspace must be 1024-byte aligned");
  if (workspace_bytes < e2e_voc_workspace_bytes(v, B, T)) return fail(-1, "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  PlanKey key{B, T, workspace};
  auto it = v->plans.find(key);
  if (it == v->plans.end()) {
    std::vector<Op> ops;
    int rc = build_plan(v, B, T, workspace, ops);
    if (rc) return rc;
    if (std::getenv("E2E_DUMP_PLAN"))   // debugging aid: one line per launch of the new plan
      for (size_t i = 0; i < ops.size(); ++i) {
        const Op& o = ops[i];
        if (o.kind == 1)
          fprintf(stderr, "plan[%zu] conv_tc %-28s mt=%d cg=%d staged=%d nt=%d taps=%d grid=%u smem=%d units=%d\n", i,
                  v->layers[o.layer].name.c_str(), o.plan.p.mt, o.plan.cg, o.plan.staged, o.plan.p.nt, o.plan.p.taps,
                  o.plan.grid.x, o.plan.smem_bytes, o.plan.p.n_units);
        else
          fprintf(stderr, "plan[%zu] kind=%d %s\n", i, o.kind, v->layers[o.layer].name.c_str());
      }
    if (v->plans.size() > 64) drop_plans(v);
    it = v->plans.emplace(key, std::move(ops)).first;
  }
  Buffers bf;
  carve(v, B, T, workspace, bf);
  const std::vector<Op>& ops = it->second;
  size_t first_conv = ops.size(), last_conv = 0;
  for (size_t i = 0; i < ops.size(); ++i)
    if (ops[i].kind == 1 || ops[i].kind == 3 || ops[i].kind == 4 || ops[i].kind == 7) {
      first_conv = i < first_conv ? i : first_conv;
      last_conv = i;
    }
  // ---- CUDA-graph replay (see GraphKey): second sighting of the same buffers captures, later ones replay ----
  static const bool graphs_on = std::getenv("E2E_NO_GRAPH") == nullptr;
  GraphEntry* ge = nullptr;
  bool capturing = false;
  if (graphs_on && !v->ev_begin && !v->ev_end) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone) {
      GraphKey gk;
      std::memset(&gk, 0, sizeof gk);
      gk.B = B;
      gk.T = T;
      gk.ws = workspace;
      gk.mel = mel;
      gk.o1 = post.wav ? (const void*)post.wav : (const void*)post.pcm;
      gk.o2 = so.spec ? (const void*)so.spec : (const void*)post.lens;
      gk.o3 = so.phase;
      gk.sB = sB;
      gk.sC = sC;
      gk.sT = sT;
      gk.scale = post.scale;
      if (v->graphs.size() > 32 && v->graphs.find(gk) == v->graphs.end()) {
        for (auto& kv : v->graphs)
          if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        v->graphs.clear();
      }
      ge = &v->graphs[gk];
      ++ge->seen;
      if (ge->exec) {
        cudaError_t e = cudaGraphLaunch(ge->exec, st);
        if (e != cudaSuccess) return fail((int)e, std::string("cudaGraphLaunch: ") + cudaGetErrorString(e));
        v->last_launches = (int)ops.size();
        v->last_graph = 1;
        return 0;
      }
      if (ge->seen >= 2 && !ge->failed) {
        // captured on a stream of the handle's own (the caller's may be the legacy default stream, which cannot be
        // captured); nothing executes during capture, the instantiated graph is then launched on the caller's stream
        if (!v->cap_stream && cudaStreamCreateWithFlags(&v->cap_stream, cudaStreamNonBlocking) != cudaSuccess)
          v->cap_stream = nullptr;
        if (v->cap_stream && cudaStreamBeginCapture(v->cap_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
          capturing = true;
        } else {
          cudaGetLastError();
          ge->failed = true;
        }
      }
    }
  }
  v->last_graph = 0;
  const cudaStream_t user_st = st;
  if (capturing) st = v->cap_stream;
  for (size_t oi = 0; oi < ops.size(); ++oi) {
    const Op& op = ops[oi];
    if (oi == first_conv && v->ev_begin) cudaEventRecord(v->ev_begin, st);
    if (op.kind == 0) {
      const long long total = (long long)B * T * (v->cin_pad / 8);
      mel_to_act_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(mel, sB, sC, sT, B, T, v->cfg.in_channels,
                                                                        v->cin_pad, bf.melA, v->f16);
    } else if (op.kind == 1) {
      int rc = launch_conv(op.plan, st);
      if (rc) return rc;
    } else if (op.kind == 3) {
      int rc = launch_pair(op.pair, st);
      if (rc) return rc;
    } else if (op.kind == 4) {
      int rc = launch_rb(op.rb, st);
      if (rc) return rc;
    } else if (op.kind == 7) {
      int rc = launch_tz(op.tz, st);
      if (rc) return rc;
    } else if (op.kind == 5) {
      const int Ts = T * v->hop, C = v->layers[v->by_name["conv_post"]].cin;
      const long long total = (long long)B * (Ts + 1) * (C / 8);
      reflect_pad_left_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(bf.Y, bf.A0, B, Ts, C);
    } else if (op.kind == 6) {
      const Layer& L = v->layers[op.layer];
      const int F = T * v->hop + 1, nb = v->cfg.istft_n_fft / 2 + 1;
      dim3 grid((F + 31) / 32, B);
      spec_phase_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(bf.A1), B, F, L.shape.n_total, nb,
                                              so.spec, so.phase);
    } else {
      const Layer& L = v->layers[op.layer];
      const int Tout = T * v->hop;
      if (L.cin == 32 && L.k == 7) {
        dim3 grid((Tout + kPostTile - 1) / kPostTile, B);
        post_conv_tanh_kernel<32, 7, kPostOpt, kPostThreads>
            <<<grid, kPostThreads, 0, st>>>(bf.Y, v->post_w, L.post_bias, B, Tout, post, v->f16);
      } else {
        dim3 grid((Tout + 255) / 256, B);
        post_conv_tanh_generic_kernel<<<grid, 256, 0, st>>>(bf.Y, reinterpret_cast<const float*>(L.d_w), L.post_bias,
                                                            B, Tout, L.cin, L.k, post, v->f16);
      }
    }
    if (oi == last_conv && v->ev_end) cudaEventRecord(v->ev_end, st);
  }
  v->ev_begin = v->ev_end = nullptr;
  v->last_launches = (int)ops.size();
  if (capturing) {
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(st, &g);
    if (e == cudaSuccess && g) {
      e = cudaGraphInstantiate(&ge->exec, g, 0);
      cudaGraphDestroy(g);
    }
    if (e != cudaSuccess || !ge->exec) {   // capture is an optimisation: enqueue this call directly and do not try again
      cudaGetLastError();
      ge->exec = nullptr;
      ge->failed = true;
      int rc = voc_forward_impl(v, mel, sB, sC, sT, B, T, post, so, workspace, workspace_bytes, stream);
      return rc;
    }
    e = cudaGraphLaunch(ge->exec, user_st);
    if (e != cudaSuccess) return fail((int)e, std::string("cudaGraphLaunch: ") + cudaGetErrorString(e));
    v->last_graph = 1;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, std::string("e2e_voc_forward launch: ") + cudaGetErrorString(e));
  return 0;
}

extern "C" int e2e_voc_forward(e2e_voc* v, const float* mel, int64_t sB, int64_t sC, int64_t sT, int32_t B,
                               int32_t T, float* wav, void* workspace, size_t workspace_bytes, void* stream) {
  if (!wav) return fail(-1, "null argument");
  PostOut post{wav, nullptr, nullptr, v ? v->hop : 1, 1.0f};
  return voc_forward_impl(v, mel, sB, sC, sT, B, T, post, SpecOut{}, workspace, workspace_bytes, stream);
}

extern "C" int e2e_voc_forward_pcm16(e2e_voc* v, const float* mel, int64_t sB, int64_t sC, int64_t sT, int32_t B,
                                     int32_t T, const int32_t* mel_lengths, float max_wav_value, int16_t* pcm,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  if (!pcm) return fail(-1, "null argument");
  if (!(max_wav_value > 0.f) || max_wav_value > 32768.f) return fail(-1, "max_wav_value must be in (0, 32768]");
  PostOut post{nullptr, pcm, mel_lengths, v ? v->hop : 1, max_wav_value};
  return voc_forward_impl(v, mel, sB, sC, sT, B, T, post, SpecOut{}, workspace, workspace_bytes, stream);
}

extern "C" int e2e_voc_forward_spec(e2e_voc* v, const float* mel, int64_t sB, int64_t sC, int64_t sT, int32_t B,
                                    int32_t T, float* spec, float* phase, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  if (!spec || !phase) return fail(-1, "null argument");
  if (v && v->cfg.istft_n_fft <= 0) return fail(-1, "e2e_voc_forward_spec needs a generator created with istft_n_fft > 0");
  SpecOut so;
  so.spec = spec;
  so.phase = phase;
  PostOut post{nullptr, nullptr, nullptr, v ? v->hop : 1, 1.0f};
  return voc_forward_impl(v, mel, sB, sC, sT, B, T, post, so, workspace, workspace_bytes, stream);
}

extern "C" int e2e_istft_forward(const float* mag, const float* phase, int32_t B, int32_t frames, int32_t n_fft,
                                 int32_t hop, int32_t win, float* wav, void* stream) {
  if (!mag || !phase || !wav) return fail(-1, "null argument");
  if (B < 1 || B > 65535 || frames < 2) return fail(-1, "need B in [1, 65535] and at least two frames");
  if (n_fft < 4 || n_fft > kIstftMaxN || (n_fft & (n_fft - 1)) || win != n_fft || hop < 1 || n_fft % hop)
    return fail(-4, "inverse STFT supported for win == n_fft = 2^m <= 64 and hop dividing n_fft");
  const int nb = n_fft / 2 + 1;
  const int L = hop * (frames - 1);
  const size_t smem = (size_t)(3 * n_fft + 2 * (256 / hop + n_fft / hop + 2) * nb) * 4;
  if (smem > 48 * 1024) return fail(-4, "inverse STFT: hop too small for the shared-memory budget");
  dim3 grid((L + 255) / 256, B);
  istft_small_kernel<<<grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(mag, phase, B, frames, n_fft, hop, wav);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, std::string("istft launch: ") + cudaGetErrorString(e));
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// N2 (SURVEY.md §8 f): the acoustic model's Postnet, the step right before the vocoder
// (e2e_tts/models/acoustic/unsupervised_fastspeech2/layers.py:507-563): conv_layers x [Conv1d(k) -> BatchNorm1d (eval:
// folded into the conv by the caller) -> tanh (all but the last)], on [B, T, C] channels-last input - the layout the
// implicit-GEMM kernel wants, so the reference's two transposes disappear.  Same conv_tc kernel, tanh epilogue.
// ---------------------------------------------------------------------------------------------------
struct e2e_postnet {
  e2e_voc core;   // layer table, packed weights, plans
  int C = 0, H = 0, n_layers = 0, k = 0, c_pad = 0, out_pad = 0;
};

extern "C" int e2e_postnet_create(int32_t n_channels, int32_t embedding_dim, int32_t conv_layers, int32_t kernel_size,
                                  e2e_postnet** out) {
  if (!out) return fail(-1, "null argument");
  if (n_channels < 1 || n_channels > 256 || embedding_dim % 64 != 0 || embedding_dim < 64 || embedding_dim > 512)
    return fail(-4, "postnet: n_channels in [1, 256], embedding_dim a multiple of 64 <= 512");
  if (conv_layers < 2 || conv_layers > 16 || !(kernel_size & 1) || kernel_size > kMaxTaps)
    return fail(-4, "postnet: 2..16 layers, odd kernel size <= 15");
  std::unique_ptr<e2e_postnet> pn(new e2e_postnet);
  pn->C = n_channels;
  pn->H = embedding_dim;
  pn->n_layers = conv_layers;
  pn->k = kernel_size;
  pn->c_pad = (n_channels + 63) / 64 * 64;
  pn->out_pad = (n_channels + 31) / 32 * 32;
  e2e_voc* v = &pn->core;
  v->cfg = e2e_voc_config{};
  v->cfg.in_channels = n_channels;
  v->cfg.upsample_initial_channel = embedding_dim;
  v->cin_pad = pn->c_pad;
  for (int i = 0; i < conv_layers; ++i) {
    const int cin = i == 0 ? n_channels : embedding_dim, cin_pad = i == 0 ? pn->c_pad : embedding_dim;
    const bool last = i + 1 == conv_layers;
    add_conv(v, "convolutions." + std::to_string(i), cin, cin_pad, last ? n_channels : embedding_dim, kernel_size, 1,
             last ? pn->out_pad : 0);
  }
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail((int)e, std::string("cudaGetDevice: ") + cudaGetErrorString(e));
  cudaDeviceGetAttribute(&v->n_sms, cudaDevAttrMultiProcessorCount, dev);
  if (v->n_sms < 1) v->n_sms = 148;
  int rc = conv_kernels_init();
  if (rc) return rc;
  *out = pn.release();
  return 0;
}

extern "C" void e2e_postnet_destroy(e2e_postnet* pn) {
  if (!pn) return;
  for (auto& L : pn->core.layers) {
    if (L.d_w) cudaFree(L.d_w);
    if (L.d_bias) cudaFree(L.d_bias);
  }
  delete pn;
}

extern "C" int e2e_postnet_load_layer(e2e_postnet* pn, int32_t index, const float* weight, int64_t weight_numel,
                                      const float* bias, int64_t bias_numel) {
  if (!pn) return fail(-1, "null argument");
  return e2e_voc_load_layer(&pn->core, ("convolutions." + std::to_string(index)).c_str(), weight, weight_numel, bias,
                            bias_numel);
}

static void postnet_carve(const e2e_postnet* pn, int B, int T, void* ws, __nv_bfloat16** xin, __nv_bfloat16** h0,
                          __nv_bfloat16** h1, float** y, size_t* total) {
  uint8_t* p = reinterpret_cast<uint8_t*>(ws);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* r = p ? p + off : nullptr;
    off += align_up(bytes, 1024);
    return r;
  };
  const size_t rows = (size_t)B * T;
  *xin = (__nv_bfloat16*)take(rows * pn->c_pad * 2);
  *h0 = (__nv_bfloat16*)take(rows * pn->H * 2);
  *h1 = (__nv_bfloat16*)take(rows * pn->H * 2);
  *y = (float*)take(rows * pn->out_pad * 4);
  *total = off;
}

extern "C" size_t e2e_postnet_workspace_bytes(const e2e_postnet* pn, int32_t B, int32_t T) {
  if (!pn || B < 1 || T < 1) return 0;
  __nv_bfloat16 *a, *b, *c;
  float* y;
  size_t total;
  postnet_carve(pn, B, T, nullptr, &a, &b, &c, &y, &total);
  return total;
}

extern "C" int e2e_postnet_forward(e2e_postnet* pn, const float* x, int32_t B, int32_t T, int32_t add_input, float* out,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  if (!pn || !x || !out || !workspace) return fail(-1, "null argument");
  if (B < 1 || T < 1) return fail(-1, "B and T must be positive");
  e2e_voc* v = &pn->core;
  if (e2e_voc_missing_layers(v) != 0) return fail(-7, "e2e_postnet_forward before all layers were loaded");
  if (reinterpret_cast<uintptr_t>(workspace) % 1024) return fail(-1, "workspace must be 1024-byte aligned");
  if (workspace_bytes < e2e_postnet_workspace_bytes(pn, B, T)) return fail(-1, "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  __nv_bfloat16 *xin, *h0, *h1;
  float* y;
  size_t total;
  postnet_carve(pn, B, T, workspace, &xin, &h0, &h1, &y, &total);
  PlanKey key{B, T, workspace};
  auto it = v->plans.find(key);
  if (it == v->plans.end()) {
    std::vector<Op> ops;
    const __nv_bfloat16* in = xin;
    for (int i = 0; i < pn->n_layers; ++i) {
      const bool last = i + 1 == pn->n_layers;
      __nv_bfloat16* o = (i & 1) ? h1 : h0;
      int rc = make_conv_op(v, ops, v->by_name["convolutions." + std::to_string(i)], B, T, in, nullptr, nullptr,
                            last ? y : nullptr, last ? nullptr : o, 1.0f, 0.f);
      if (rc) return rc;
      ops.back().plan.p.act_tanh = last ? 0 : 1;   // torch.tanh(self.convolutions[i](x)), layers.py:558-559
      in = o;
    }
    if (v->plans.size() > 64) v->plans.clear();
    it = v->plans.emplace(key, std::move(ops)).first;
  }
  // [B, T, C] fp32 is element (b, c, t) at x[b*T*C + c + t*C]: the [B,80,T]-view kernel does the bf16 / padding pass
  {
    const long long tot = (long long)B * T * (pn->c_pad / 8);
    mel_to_act_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(x, (long long)T * pn->C, 1, pn->C, B, T, pn->C,
                                                                    pn->c_pad, xin, v->f16);
  }
  for (const Op& op : it->second) {
    int rc = launch_conv(op.plan, st);
    if (rc) return rc;
  }
  {
    const long long tot = (long long)B * T * pn->C;
    postnet_out_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(y, add_input ? x : nullptr, tot, pn->C, pn->out_pad,
                                                                     out);
  }
  v->last_launches = (int)it->second.size() + 2;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, std::string("e2e_postnet_forward launch: ") + cudaGetErrorString(e));
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Standalone residual blocks: ResBlock1.forward / ResBlock2.forward (e2e_tts/models/vocoder/layers.py:33-40,60-65) as
// callable modules, on the same tcgen05 convolution kernel as the generator (every conv its own conv_tc launch; the
// fused pair / whole-resblock kernels are the generator's arrangement).  x, out: fp32 [B][C][T] (out contiguous).
// ---------------------------------------------------------------------------------------------------
struct e2e_resblock {
  e2e_voc core;
  int kind = 1, C = 0, k = 3, nd = 0;
  int dil[E2E_MAX_DILATIONS] = {0};
};

extern "C" int e2e_resblock_create(int32_t kind, int32_t channels, int32_t kernel_size, const int32_t* dilations,
                                   int32_t n_dilations, e2e_resblock** out) {
  if (!out || !dilations) return fail(-1, "null argument");
  if (kind != 1 && kind != 2) return fail(-1, "kind must be 1 (ResBlock1) or 2 (ResBlock2)");
  if (channels != 32 && (channels % 64 != 0 || channels < 64 || channels > 512))
    return fail(-4, "resblock: channels must be 32 or a multiple of 64, <= 512");
  if (!(kernel_size & 1) || kernel_size < 1 || kernel_size > kMaxTaps) return fail(-4, "resblock: odd kernel size <= 15");
  if (n_dilations < 1 || n_dilations > E2E_MAX_DILATIONS) return fail(-1, "bad number of dilations");
  std::unique_ptr<e2e_resblock> rb(new e2e_resblock);
  rb->kind = kind;
  rb->C = channels;
  rb->k = kernel_size;
  rb->nd = n_dilations;
  e2e_voc* v = &rb->core;
  v->cfg = e2e_voc_config{};
  v->cfg.in_channels = channels;
  v->cfg.upsample_initial_channel = channels;
  v->cin_pad = channels;
  for (int m = 0; m < n_dilations; ++m) {
    const int d = dilations[m];
    if (d < 1 || (kernel_size - 1) / 2 * d > 127) return fail(-4, "resblock: dilated receptive field too wide");
    rb->dil[m] = d;
    if (kind == 1) {
      add_conv(v, "convs1." + std::to_string(m), channels, channels, channels, kernel_size, d);
      add_conv(v, "convs2." + std::to_string(m), channels, channels, channels, kernel_size, 1);
    } else {
      add_conv(v, "convs." + std::to_string(m), channels, channels, channels, kernel_size, d);
    }
  }
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail((int)e, std::string("cudaGetDevice: ") + cudaGetErrorString(e));
  cudaDeviceGetAttribute(&v->n_sms, cudaDevAttrMultiProcessorCount, dev);
  if (v->n_sms < 1) v->n_sms = 148;
  int rc = conv_kernels_init();
  if (rc) return rc;
  *out = rb.release();
  return 0;
}

extern "C" void e2e_resblock_destroy(e2e_resblock* rb) {
  if (!rb) return;
  for (auto& L : rb->core.layers) {
    if (L.d_w) cudaFree(L.d_w);
    if (L.d_wtz) cudaFree(L.d_wtz);
    if (L.d_bias) cudaFree(L.d_bias);
    for (uint8_t* w : L.alt_w)
      if (w) cudaFree(w);
  }
  delete rb;
}

extern "C" int e2e_resblock_load_layer(e2e_resblock* rb, const char* name, const float* weight, int64_t weight_numel,
                                       const float* bias, int64_t bias_numel) {
  if (!rb) return fail(-1, "null argument");
  return e2e_voc_load_layer(&rb->core, name, weight, weight_numel, bias, bias_numel);
}

static void resblock_carve(const e2e_resblock* rb, int B, int T, void* ws, __nv_bfloat16** xin, __nv_bfloat16** mid,
                           __nv_bfloat16** p0, __nv_bfloat16** p1, float** y, size_t* total) {
  uint8_t* p = reinterpret_cast<uint8_t*>(ws);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* r = p ? p + off : nullptr;
    off += align_up(bytes, 1024);
    return r;
  };
  const size_t n = (size_t)B * T * rb->C;
  *xin = (__nv_bfloat16*)take(n * 2);
  *mid = (__nv_bfloat16*)take(n * 2);
  *p0 = (__nv_bfloat16*)take(n * 2);
  *p1 = (__nv_bfloat16*)take(n * 2);
  *y = (float*)take(n * 4);
  *total = off;
}

extern "C" size_t e2e_resblock_workspace_bytes(const e2e_resblock* rb, int32_t B, int32_t T) {
  if (!rb || B < 1 || T < 1) return 0;
  __nv_bfloat16 *a, *b, *c, *d;
  float* y;
  size_t total;
  resblock_carve(rb, B, T, nullptr, &a, &b, &c, &d, &y, &total);
  return total;
}

extern "C" int e2e_resblock_forward(e2e_resblock* rb, const float* x, int64_t sB, int64_t sC, int64_t sT, int32_t B,
                                    int32_t T, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!rb || !x || !out || !workspace) return fail(-1, "null argument");
  if (B < 1 || T < 1 || B > 65535) return fail(-1, "B in [1, 65535] and T >= 1 required");
  e2e_voc* v = &rb->core;
  if (e2e_voc_missing_layers(v) != 0) return fail(-7, "e2e_resblock_forward before all layers were loaded");
  if (reinterpret_cast<uintptr_t>(workspace) % 1024) return fail(-1, "workspace must be 1024-byte aligned");
  if (workspace_bytes < e2e_resblock_workspace_bytes(rb, B, T)) return fail(-1, "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  __nv_bfloat16 *xin, *mid, *p0, *p1;
  float* y;
  size_t total;
  resblock_carve(rb, B, T, workspace, &xin, &mid, &p0, &p1, &y, &total);
  PlanKey key{B, T, workspace};
  auto it = v->plans.find(key);
  if (it == v->plans.end()) {
    std::vector<Op> ops;
    const __nv_bfloat16* ain = xin;   // bf16 leaky_relu(x, 0.1): conv operand and (inverted) residual
    for (int m = 0; m < rb->nd; ++m) {
      const bool last = m + 1 == rb->nd;
      __nv_bfloat16* nxt = (m & 1) ? p1 : p0;
      int rc;
      if (rb->kind == 1) {
        rc = make_conv_op(v, ops, v->by_name["convs1." + std::to_string(m)], B, T, ain, nullptr, nullptr, nullptr, mid,
                          0.1f, 0.f);
        if (rc) return rc;
        rc = make_conv_op(v, ops, v->by_name["convs2." + std::to_string(m)], B, T, mid, ain, nullptr, last ? y : nullptr,
                          last ? nullptr : nxt, 0.1f, 0.f);
      } else {
        rc = make_conv_op(v, ops, v->by_name["convs." + std::to_string(m)], B, T, ain, ain, nullptr, last ? y : nullptr,
                          last ? nullptr : nxt, 0.1f, 0.f);
      }
      if (rc) return rc;
      ain = nxt;
    }
    if (v->plans.size() > 16) drop_plans(v);
    it = v->plans.emplace(key, std::move(ops)).first;
  }
  {
    const long long tot = (long long)B * T * (rb->C / 8);
    mel_to_act_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(x, sB, sC, sT, B, T, rb->C, rb->C, xin, v->f16, 0.1f);
  }
  for (const Op& op : it->second) {
    int rc = launch_conv(op.plan, st);
    if (rc) return rc;
  }
  {
    dim3 grid((T + 31) / 32, (rb->C + 31) / 32, B);
    cl_to_ncl_kernel<<<grid, 256, 0, st>>>(y, out, T, rb->C);
  }
  v->last_launches = (int)it->second.size() + 2;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, std::string("e2e_resblock_forward launch: ") + cudaGetErrorString(e));
  return 0;
}

extern "C" const char* e2e_last_error_string(void) { return last_error().c_str(); }
extern "C" const char* e2e_version_string(void) { return "e2e_tts_b200 0.1 sm_100a"; }
