// CPU emulation harness for csrc/dwconv_tiled.cuh: runs the kernel's three phases block by block, thread by
// thread, with plain arrays standing in for shared memory.  Built by tests/test_dw_tiled_host.py with nvcc
// (host code only is executed); it checks tile/halo/stride/ragged-edge index arithmetic without a GPU.
#include <vector>
#include "dwconv_tiled.cuh"

extern "C" void occd_set_last_error(const char*) {}

namespace {

template <typename T, int K, int S, int CVB, int TH, int PXV = 4>
void run(const dwt::Args& a, int B) {
  using C_ = dwt::Cfg<T, K, S, CVB, TH, PXV>;
  const int tiles_y = (a.OH + TH - 1) / TH;
  std::vector<T> tile(C_::TILE_ELEMS);
  std::vector<float> wsm(C_::W_ELEMS), red(C_::RED_ELEMS);
  for (int z = 0; z < B; ++z)
    for (int y = 0; y < (a.C + C_::CT - 1) / C_::CT; ++y)
      for (int x = 0; x < a.tiles_x * tiles_y; ++x) {
        const dwt::BlockIdx blk{x, y, z};
        // poison "shared memory" so that a read of a never-written slot shows up as NaN in the output
        memset(tile.data(), 0xff, tile.size() * sizeof(T));
        memset(wsm.data(), 0xff, wsm.size() * 4);
        memset(red.data(), 0xff, red.size() * 4);
        for (int t = 0; t < C_::NT; ++t) dwt::phase_load<T, K, S, CVB, TH, PXV>(a, blk, t, tile.data(), wsm.data());
        for (int t = 0; t < C_::NT; ++t)
          dwt::phase_compute<T, K, S, CVB, TH, PXV>(a, blk, t, tile.data(), wsm.data(), red.data());
        if (a.pool)
          for (int t = 0; t < C_::NT; ++t) dwt::phase_pool<T, K, S, CVB, TH, PXV>(a, blk, t, red.data());
      }
}

template <typename T, int K, int S>
void run_ks(const dwt::Args& a, int B, dwt::Choice ch) {
  if (ch.cvb == 4 && sizeof(T) == 4 && S == 1) run<T, K, S, 4, 16, 8>(a, B);   // fp32 stride 1: 8 outputs per thread, as the launcher
  else if (ch.cvb == 4) run<T, K, S, 4, 16>(a, B);
  else if (ch.th == 16) run<T, K, S, 8, 16>(a, B);
  else run<T, K, S, 8, 8>(a, B);
}

template <typename T>
int run_t(const dwt::Args& a, int B, int K, int stride, dwt::Choice ch) {
  if (K == 3 && stride == 1) run_ks<T, 3, 1>(a, B, ch);
  else if (K == 3 && stride == 2) run_ks<T, 3, 2>(a, B, ch);
  else if (K == 5 && stride == 1) run_ks<T, 5, 1>(a, B, ch);
  else if (K == 5 && stride == 2) run_ks<T, 5, 2>(a, B, ch);
  else return 1;
  return 0;
}

}  // namespace

// th: 0 = the launcher's own choice, 8 / 16 = forced (ignored for the CVB = 4 shape, which is always 16)
// elem_size: 2 = bf16 activations, 4 = fp32 (TF32-valued) activations
extern "C" int dw_tiled_emulate(const void* in, const float* w, const float* bias, void* out, long long* pool,
                                int elem_size, int B, int H, int W, int OH, int OW, int C, int cs_in, int cs_out, int K,
                                int stride, int pad_top, int pad_left, int act, int th, int* cvb_out, int* th_out) {
  dwt::Args a;
  a.in = in; a.w = w; a.bias = bias; a.out = out; a.pool = pool;
  a.H = H; a.W = W; a.OH = OH; a.OW = OW; a.C = C; a.cs_in = cs_in; a.cs_out = cs_out;
  a.pad_top = pad_top; a.pad_left = pad_left; a.act = act; a.tiles_x = (OW + dwt::kTW - 1) / dwt::kTW;
  dwt::Choice ch = dwt::choose(B, OH, OW, C, stride, 148, elem_size);
  if (th && ch.cvb == 8 && !(stride == 2 && th == 16)) ch.th = th;
  if (cvb_out) *cvb_out = ch.cvb;
  if (th_out) *th_out = ch.th;
  return elem_size == 4 ? run_t<float>(a, B, K, stride, ch) : run_t<__nv_bfloat16>(a, B, K, stride, ch);
}
