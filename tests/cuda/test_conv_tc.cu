// Standalone device test for the tcgen05 convolution kernel (no torch): each configuration is checked
// against a naive CUDA-core convolution over the same bf16-rounded operands.  Run one configuration per
// process (`test_conv_tc <id>`), because a watchdog trap poisons the CUDA context.
//   test_conv_tc list          -> number of configurations
//   test_conv_tc <id> [reps]   -> run configuration <id>; exit code 0 = pass
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../../e2e_tts_b200/csrc/conv_host.cuh"

using namespace e2e;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      report_watchdog();                                                               \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

static unsigned int* g_wd_host = nullptr;
static void report_watchdog() {
  if (g_wd_host && *g_wd_host) printf("WATCHDOG code 0x%x\n", *g_wd_host);
}

struct Cfg {
  const char* name;
  int cin, n_total, nt, taps, dil;  // dil > 0: symmetric dilated conv; dil == 0: polyphase (taps = 2)
  int u;                            // polyphase upsample factor (dil == 0)
  int B, T, mt;
  int res, sum, div3, f32out, actout;
  int reserved;
};

static const Cfg kCfgs[] = {
    {"gemm64 1tap", 64, 64, 64, 1, 1, 0, 1, 128, 1, 0, 0, 0, 1, 0, 0},
    {"c64 k3 d8 (shift mult of 8)", 64, 64, 64, 3, 8, 0, 2, 300, 1, 0, 0, 0, 1, 1, 0},
    {"c64 k3 d1 (odd row shift)", 64, 64, 64, 3, 1, 0, 2, 300, 1, 0, 0, 0, 1, 1, 0},
    {"c128 k11 d5 mt2 res", 128, 128, 128, 11, 5, 0, 3, 1000, 2, 1, 0, 0, 1, 1, 0},
    {"c128 k7 d3 mt2 res sum div3", 128, 128, 128, 7, 3, 0, 2, 777, 2, 1, 1, 1, 0, 1, 0},
    {"c256 k11 d5 mt2", 256, 256, 256, 11, 5, 0, 2, 700, 2, 1, 0, 0, 1, 1, 0},
    {"c256 k3 d1 mt1", 256, 256, 256, 3, 1, 0, 2, 431, 1, 0, 0, 0, 1, 1, 0},
    {"c64 k11 d3 mt4", 64, 64, 64, 11, 3, 0, 2, 1500, 4, 1, 0, 0, 1, 1, 0},
    {"c32 k7 d3 mt4 (SW64)", 32, 32, 32, 7, 3, 0, 2, 2000, 4, 1, 0, 0, 1, 1, 0},
    {"c32 k3 d1 mt4 (SW64)", 32, 32, 32, 3, 1, 0, 2, 2000, 4, 1, 1, 1, 0, 1, 0},
    {"pre 128->512 k7", 128, 512, 256, 7, 1, 0, 2, 431, 1, 0, 0, 0, 0, 1, 0},
    {"ups0 512->8x256", 512, 2048, 256, 2, 0, 8, 2, 431, 1, 0, 0, 0, 1, 1, 0},
    {"ups1 256->8x128", 256, 1024, 256, 2, 0, 8, 2, 600, 2, 0, 0, 0, 1, 1, 0},
    {"ups2 128->2x64", 128, 128, 64, 2, 0, 2, 2, 900, 2, 0, 0, 0, 1, 1, 0},
    {"ups3 64->2x32", 64, 64, 32, 2, 0, 2, 2, 900, 4, 0, 0, 0, 1, 1, 0},
    // performance shapes (B=16, 5 s): stage-1 / stage-0 / stage-2 / stage-3 k=11 convs
    {"perf c128 k11 d1 T27584", 128, 128, 128, 11, 1, 0, 16, 27584, 2, 1, 0, 0, 1, 1, 0},
    {"perf c256 k11 d1 T3448", 256, 256, 256, 11, 1, 0, 16, 3448, 2, 1, 0, 0, 1, 1, 0},
    {"perf c64 k11 d1 T55168", 64, 64, 64, 11, 1, 0, 16, 55168, 4, 1, 0, 0, 1, 1, 0},
    {"perf c32 k11 d1 T110336", 32, 32, 32, 11, 1, 0, 16, 110336, 4, 1, 0, 0, 1, 1, 0},
    {"perf c128 k3 d1 act-only", 128, 128, 128, 3, 1, 0, 16, 27584, 2, 0, 0, 0, 0, 1, 0},
    {"perf c128 k7 d3 act-only", 128, 128, 128, 7, 3, 0, 16, 27584, 2, 0, 0, 0, 0, 1, 0},
    {"perf c128 k7 d1 res f32+act", 128, 128, 128, 7, 1, 0, 16, 27584, 2, 1, 0, 0, 1, 1, 0},
    // stage-0 shapes as the forward plans them (mt = 1)
    {"perf c256 k3 d1 mt1 res", 256, 256, 256, 3, 1, 0, 16, 3448, 1, 1, 0, 0, 0, 1, 0},
    {"perf c256 k7 d3 mt1", 256, 256, 256, 7, 3, 0, 16, 3448, 1, 0, 0, 0, 0, 1, 0},
    {"perf c256 k11 d5 mt1 res sum", 256, 256, 256, 11, 5, 0, 16, 3448, 1, 1, 1, 0, 0, 1, 0},
    {"c256 k7 d1 mt1 odd units", 256, 256, 256, 7, 1, 0, 3, 300, 1, 1, 0, 0, 1, 1, 0},
};
static const int kNumCfgs = sizeof(kCfgs) / sizeof(kCfgs[0]);

__global__ void ref_conv(const __nv_bfloat16* x, const float* wg, const float* bias, const __nv_bfloat16* res,
                         const __nv_bfloat16* sum, float* out, int B, int T, int cin, int n_total, int nt, int taps,
                         const int* shifts, int div3) {
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * T * n_total;
  if (idx >= total) return;
  const int n = idx % n_total;
  const int t = (idx / n_total) % T;
  const int b = idx / ((size_t)n_total * T);
  const int nti = n / nt;
  float acc = 0.f;
  for (int j = 0; j < taps; ++j) {
    const int tt = t + shifts[nti * taps + j];
    if (tt < 0 || tt >= T) continue;
    const __nv_bfloat16* xr = x + ((size_t)b * T + tt) * cin;
    const float* wr = wg + ((size_t)n * taps + j) * cin;
    for (int c = 0; c < cin; ++c) acc += __bfloat162float(xr[c]) * wr[c];
  }
  float v = acc + bias[n];
  if (res) {
    const float a = __bfloat162float(res[idx]);
    v += a > 0.f ? a : a * 10.0f;
  }
  if (sum) v += __bfloat162float(sum[idx]);
  if (div3) v = v / 3.0f;
  out[idx] = v;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    printf("usage: %s list | <id> [reps]\n", argv[0]);
    return 1;
  }
  if (!strcmp(argv[1], "list")) {
    printf("%d\n", kNumCfgs);
    return 0;
  }
  const int id = atoi(argv[1]);
  const int reps = argc > 2 ? atoi(argv[2]) : 3;
  if (id < 0 || id >= kNumCfgs) return 1;
  Cfg c = kCfgs[id];
  {  // E2E_CONV_STAGED=1: exercise the staged TMA-store epilogue (bf16 output only) on every configuration that has one
    const char* est0 = getenv("E2E_CONV_STAGED");
    if (est0 && est0[0] == '1' && c.actout) c.f32out = 0;
  }
  printf("[cfg %d] %s: cin=%d N=%d nt=%d taps=%d dil=%d B=%d T=%d mt=%d\n", id, c.name, c.cin, c.n_total, c.nt,
         c.taps, c.dil, c.B, c.T, c.mt);

  CK(cudaSetDevice(0));
  CK(cudaHostAlloc(&g_wd_host, 4, cudaHostAllocMapped));
  *g_wd_host = 0;
  unsigned int* wd_dev = nullptr;
  CK(cudaHostGetDevicePointer(&wd_dev, g_wd_host, 0));
  CK(cudaMemcpyToSymbol(g_watchdog_host, &wd_dev, sizeof(wd_dev)));

  ConvShape s;
  s.cin = c.cin;
  s.n_total = c.n_total;
  s.nt = c.nt;
  s.taps = c.taps;
  const int n_tiles = c.n_total / c.nt;
  s.shifts.resize(n_tiles * c.taps);
  if (c.dil > 0) {
    for (int i = 0; i < n_tiles; ++i)
      for (int j = 0; j < c.taps; ++j) s.shifts[i * c.taps + j] = (j - (c.taps - 1) / 2) * c.dil;
  } else {
    // polyphase: columns n = p*C_out + co; phases p < u/2 use rows {q, q-1}, others {q, q+1}
    const int cout = c.n_total / c.u;
    for (int i = 0; i < n_tiles; ++i) {
      const int p = (i * c.nt) / cout;
      s.shifts[i * 2 + 0] = 0;
      s.shifts[i * 2 + 1] = (p < c.u / 2) ? -1 : +1;
    }
  }

  std::mt19937 rng(1234 + id);
  std::normal_distribution<float> nd(0.f, 1.f);
  const size_t nx = (size_t)c.B * c.T * c.cin, nw = (size_t)c.n_total * c.taps * c.cin,
               no = (size_t)c.B * c.T * c.n_total;
  std::vector<uint16_t> hx(nx);
  for (auto& v : hx) v = f32_to_bf16_rn(nd(rng));
  std::vector<float> hw(nw), hb(c.n_total);
  std::vector<uint16_t> hres, hsum;
  const float wscale = 1.0f / sqrtf((float)c.cin * c.taps);
  for (auto& v : hw) v = bf16_to_f32(f32_to_bf16_rn(nd(rng) * wscale));
  for (auto& v : hb) v = nd(rng) * 0.1f;
  if (c.res) {
    hres.resize(no);
    for (auto& v : hres) v = f32_to_bf16_rn(nd(rng));
  }
  if (c.sum) {
    hsum.resize(no);
    for (auto& v : hsum) v = f32_to_bf16_rn(nd(rng));
  }
  std::vector<uint8_t> hpack(packed_weight_bytes(s));
  pack_conv_weights(s, hw.data(), hpack.data());

  __nv_bfloat16 *dx, *dact = nullptr;
  float *dw, *db, *dout = nullptr, *dref;
  __nv_bfloat16 *dres = nullptr, *dsum = nullptr;
  uint8_t* dpack;
  int* dshift;
  CK(cudaMalloc(&dx, nx * 2));
  CK(cudaMalloc(&dw, nw * 4));
  CK(cudaMalloc(&db, c.n_total * 4));
  CK(cudaMalloc(&dpack, hpack.size()));
  CK(cudaMalloc(&dref, no * 4));
  CK(cudaMalloc(&dshift, s.shifts.size() * 4));
  CK(cudaMemcpy(dx, hx.data(), nx * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, hw.data(), nw * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), c.n_total * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dpack, hpack.data(), hpack.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dshift, s.shifts.data(), s.shifts.size() * 4, cudaMemcpyHostToDevice));
  if (c.res) {
    CK(cudaMalloc(&dres, no * 2));
    CK(cudaMemcpy(dres, hres.data(), no * 2, cudaMemcpyHostToDevice));
  }
  if (c.sum) {
    CK(cudaMalloc(&dsum, no * 2));
    CK(cudaMemcpy(dsum, hsum.data(), no * 2, cudaMemcpyHostToDevice));
  }
  if (c.f32out) {
    CK(cudaMalloc(&dout, no * 4));
    CK(cudaMemset(dout, 0xff, no * 4));
  }
  if (c.actout) {
    CK(cudaMalloc(&dact, no * 2));
    CK(cudaMemset(dact, 0xff, no * 2));
  }

  ConvPlan plan;
  const char* est = getenv("E2E_CONV_STAGED");
  const int want_staged = est && est[0] == '1' && c.actout && !c.f32out;
  int rc = plan_conv(plan, s, c.B, c.T, c.mt, 148, 0, want_staged);
  if (rc) {
    printf("plan_conv failed: %s\n", last_error().c_str());
    return 3;
  }
  ConvParams& p = plan.p;
  rc = make_act_tensor_map(&plan.tm, dx, c.B, c.T, c.cin, p.rowb / 2, p.box_rows);
  if (rc) {
    printf("tensor map failed: %s\n", last_error().c_str());
    return 3;
  }
  p.w = dpack;
  rc = conv_weight_map(plan, dpack);
  if (rc) {
    printf("weight tensor map failed: %s\n", last_error().c_str());
    return 3;
  }
  conv_set_bias(plan, db, getenv("E2E_CONV_BIAS_GLOBAL") ? nullptr : hb.data(), c.n_total);  // (random per column: no period)
  p.res_act = dres;
  p.res_inv_slope = 10.0f;
  p.sum_a = dsum;
  p.out_f32 = dout;
  p.out_act = dact;
  rc = conv_output_map(plan, dact, c.B, c.T);
  if (rc) {
    printf("output tensor map failed: %s\n", last_error().c_str());
    return 3;
  }
  p.slope = 0.1f;
  p.divisor = c.div3 ? 3.0f : 0.f;
  printf("  plan: staged=%d cg=%d grid=%d units=%d smem=%d mt=%d slab_rows=%d box=%d panel_slots=%d stages=%d stage_bytes=%d chunks=%d n_acc=%d hl=%d\n",
         plan.staged, plan.cg, plan.grid.x, p.n_units, plan.smem_bytes, p.mt, p.slab_rows, p.box_rows, p.panel_slots, p.n_stages,
         p.stage_bytes, p.n_chunks, p.n_acc, p.hl);

  rc = launch_conv(plan, 0);
  if (rc) {
    printf("launch failed: %s\n", last_error().c_str());
    return 3;
  }
  CK(cudaDeviceSynchronize());

  ref_conv<<<(unsigned)((no + 255) / 256), 256>>>(dx, dw, db, dres, dsum, dref, c.B, c.T, c.cin, c.n_total, c.nt,
                                                  c.taps, dshift, c.div3);
  CK(cudaDeviceSynchronize());

  std::vector<float> href(no), hout;
  std::vector<uint16_t> hact;
  CK(cudaMemcpy(href.data(), dref, no * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  double maxerr = 0, maxref = 0;
  if (c.f32out) {
    hout.resize(no);
    CK(cudaMemcpy(hout.data(), dout, no * 4, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < no; ++i) {
      const double e = fabs((double)hout[i] - href[i]);
      const bool isbad = !(e <= 2e-3 + 2e-3 * fabs(href[i]));
      if (e > maxerr || std::isnan(hout[i])) maxerr = std::isnan(hout[i]) ? 1e30 : e;
      if (fabs(href[i]) > maxref) maxref = fabs(href[i]);
      if (isbad && bad++ < 12) {
        const int n = i % c.n_total, t = (i / c.n_total) % c.T, b = i / ((size_t)c.n_total * c.T);
        printf("  f32 mismatch b=%d t=%d n=%d got=%g want=%g\n", b, t, n, hout[i], href[i]);
      }
    }
    printf("  f32: max_abs_err=%.3g (max |ref|=%.3g) bad=%d/%zu\n", maxerr, maxref, bad, no);
  }
  int bad2 = 0;
  if (c.actout) {
    hact.resize(no);
    CK(cudaMemcpy(hact.data(), dact, no * 2, cudaMemcpyDeviceToHost));
    double maxe2 = 0;
    for (size_t i = 0; i < no; ++i) {
      float r = href[i];
      r = r > 0 ? r : r * 0.1f;
      const float g = bf16_to_f32(hact[i]);
      const double e = fabs((double)g - r);
      const bool isbad = !(e <= 4e-3 + 1e-2 * fabs(r));
      if (e > maxe2) maxe2 = e;
      if (isbad && bad2++ < 12) {
        const int n = i % c.n_total, t = (i / c.n_total) % c.T, b = i / ((size_t)c.n_total * c.T);
        printf("  act mismatch b=%d t=%d n=%d got=%g want=%g\n", b, t, n, g, r);
      }
    }
    printf("  act: max_abs_err=%.3g bad=%d/%zu\n", maxe2, bad2, no);
  }

  // timing
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 2; ++i) launch_conv(plan, 0);
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) launch_conv(plan, 0);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= reps;
  const double flops = 2.0 * c.B * c.T * (double)c.n_total * c.taps * c.cin;
  printf("  time %.4f ms  -> %.1f TFLOP/s\n", ms, flops / ms * 1e-9);
#ifdef E2E_TRACE
  {
    static unsigned long long tr[512][16];
    launch_conv(plan, 0);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(tr, g_trace, sizeof(tr)));
    const int nblk = plan.grid.x;
    unsigned long long tmin = ~0ull;
    for (int i = 0; i < 512 && i < nblk; ++i) tmin = tr[i][0] < tmin ? tr[i][0] : tmin;
    printf("  trace (us rel. to first CTA start): cta | start setup | mma0begin mma0issued mmaAllIssued | epi0accFull epi0done | end\n");
    for (int i = 0; i < 512 && i < nblk; ++i) {
      if (!(i < 4 || i % 49 == 0)) continue;
      printf("  %5d |", i);
      for (int s = 0; s < 8; ++s) {
        printf(" %7.2f", (double)(tr[i][s] - tmin) * 1e-3);
        if (s == 1 || s == 4 || s == 6) printf(" |");
      }
      printf(" | unit0 waits: weights %.2f panels %.2f us\n", tr[i][8] * 1e-3, tr[i][9] * 1e-3);
    }
  }
#endif
  const bool ok = bad == 0 && bad2 == 0;
  printf("[cfg %d] %s\n", id, ok ? "PASS" : "FAIL");
  return ok ? 0 : 1;
}
