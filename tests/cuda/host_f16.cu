// Host-only check program for the weight packer's fp32 -> fp16 / bf16 conversions (conv_host.cuh): prints, for a sweep of
// fp32 bit patterns given on stdin (one hex word per line), the 16-bit results.  Run on the CPU by tests/test_host_pack.py.
#include <cstdio>
#include <cstring>
#include "../../e2e_tts_b200/csrc/conv_host.cuh"
int main() {
  unsigned int u;
  while (scanf("%x", &u) == 1) {
    float f;
    memcpy(&f, &u, 4);
    printf("%04x %04x\n", (unsigned)e2e::f32_to_f16_rn(f), (unsigned)e2e::f32_to_bf16_rn(f));
  }
  return 0;
}
