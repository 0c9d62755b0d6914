// CPU emulation harness for csrc/upsample_rows.cuh (same scheme as dw_tiled_emul.cu).
#include "upsample_rows.cuh"

extern "C" void occd_set_last_error(const char*) {}

template <int CVB>
static void run(const upr::Args& a, int B) {
  const int pxb = upr::kThreads / CVB;
  for (int z = 0; z < B * a.OH; ++z)
    for (int y = 0; y < (a.CV + CVB - 1) / CVB; ++y)
      for (int x = 0; x < (a.OW + pxb - 1) / pxb; ++x)
        for (int t = 0; t < upr::kThreads; ++t) upr::body<CVB>(a, x, y, z, t);
}

extern "C" int upsample_rows_emulate(const void* in, void* out, int B, int h, int w, int OH, int OW, int C, int cs_in,
                                     int in_off, int cs_out, int out_off) {
  const int CV = (C + 7) / 8;
  const float sy = OH > 1 ? (float)(h - 1) / (float)(OH - 1) : 0.f;
  const float sx = OW > 1 ? (float)(w - 1) / (float)(OW - 1) : 0.f;
  upr::Args a{(const __nv_bfloat16*)in, (__nv_bfloat16*)out, h, w, OH, OW, CV, cs_in, in_off, cs_out, out_off, sy, sx};
  switch (upr::choose_cvb(CV)) {
    case 1: run<1>(a, B); break;
    case 2: run<2>(a, B); break;
    case 4: run<4>(a, B); break;
    case 8: run<8>(a, B); break;
    case 16: run<16>(a, B); break;
    default: run<32>(a, B); break;
  }
  return 0;
}
