// CPU emulation harness for csrc/se_fold_strip.cuh (same scheme as dw_tiled_emul.cu).
#include <vector>
#include "se_fold_strip.cuh"

extern "C" void occd_set_last_error(const char*) {}

extern "C" int se_fold_strip_emulate(long long* pool, const float* hidden, const float* w2t, const float* b2,
                                     const float* master, void* out, int B, int C, int R, int rows, int Kpad) {
  sef::Args a{pool, hidden, w2t, b2, master, (__nv_bfloat16*)out, C, R, rows, Kpad};
  std::vector<float> part(sef::kWarps * sef::kStrip);
  for (int img = 0; img < B; ++img)
    for (int x = 0; x < (Kpad + sef::kStrip - 1) / sef::kStrip; ++x) {
      memset(part.data(), 0xff, part.size() * 4);
      for (int t = 0; t < sef::kThreads; ++t) sef::phase_partial(a, x, img, t, part.data());
      for (int t = 0; t < sef::kThreads; ++t) sef::phase_fold(a, x, img, t, part.data());
    }
  return 0;
}
