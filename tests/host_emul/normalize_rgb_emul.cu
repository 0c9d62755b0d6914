// CPU emulation harness for csrc/normalize_rgb.cuh (IEEE float32 arithmetic with explicit rounding: the host run is
// the device result).
#include "normalize_rgb.cuh"

extern "C" void occd_set_last_error(const char*) {}

extern "C" int normalize_rgb_emulate(const void* in, float* out, int H0, int W0, int H, int W, const float* mean,
                                     const float* stdv) {
  if (H > H0 || W > W0) return 1;
  nrm::Args a;
  a.in = (const unsigned char*)in; a.out = out; a.W0 = W0; a.H = H; a.W = W;
  for (int c = 0; c < 3; ++c) { a.mean[c] = mean[c]; a.stdv[c] = stdv[c]; }
  for (long long i = 0; i < (long long)H * W; ++i) nrm::body(a, i);
  return 0;
}
