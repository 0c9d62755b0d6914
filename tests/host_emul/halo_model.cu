// CPU model of conv_halo_kernel (csrc/conv.cu) driven by the REAL host-side plan geometry (halo_geometry):
// tile decode, TMA box origin / zero fill / traversal stride, the row-linearised halo box, per-group A row
// offsets (tap_off16), M-tile / TMEM-lane <-> box-position mapping, validity mask and -- for the x-packed
// variant -- the lane-shifted epilogue sum.  tcgen05 / TMA themselves are not modelled: an MMA is
// "accumulator row i += A row (i + offset) . B^T", a TMA load is a strided gather with zero fill, which is what
// the validated per-tap kernel relies on as well.  Rows outside the loaded box are poisoned with NaN so that any
// dependence of a stored output on garbage rows shows up.
//
// The source file is included whole so that the static host functions are reachable; only host code runs here.
#include <cmath>
#include <vector>
#include "../../occdepth_b200/csrc/conv.cu"

static thread_local char g_err[512];
extern "C" void occd_set_last_error(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg ? msg : ""); }
extern "C" const char* halo_model_error() { return g_err; }

// desc: src[0] -> float [B][ID][IH][IW][src_cstride], weight -> float [n_taps][Cout_pad][Kpad], bias -> float
// out: float [B][OD][OH][OW][Cout] (pre-filled by the caller; only valid positions are written)
// info (optional, 10 ints): BD BH BW PD PH PW nM N_tile n_groups stages
extern "C" int halo_model(const occd_conv_desc* d, int xp, float* out, int* info) {
  occd_conv_plan pl;
  memset(&pl, 0, sizeof(pl));
  const int rc = halo_geometry(d, &pl, xp != 0);
  if (rc != OCCD_OK) return rc;
  const HaloParams& p = pl.halo;
  if (info) {
    const int v[10] = {p.BD, p.BH, p.BW, p.PD, p.PH, p.PW, p.nM, p.N_tile, p.n_taps, p.stages};
    for (int i = 0; i < 10; ++i) info[i] = v[i];
  }
  if (!out) return 0;  // geometry only
  const int KC = pl.kc, row_bytes = KC * 2;
  const int rows_alloc = p.a_stage_bytes / row_bytes;
  const int box_rows = p.PD * p.PH * p.PW;
  if (p.box_bytes != box_rows * row_bytes) return 100;
  if ((size_t)p.w_bytes + (size_t)p.stages * p.a_stage_bytes + 128 + 1024 != pl.smem || pl.smem > 227 * 1024) return 101;
  if (p.stages < 1 || p.tmem_cols > 512 || p.nM * p.N_tile > p.set_stride) return 102;
  const float* src = (const float*)d->src[0];
  const float* wgt = (const float*)d->weight;
  const float* bias = (const float*)d->bias;
  const int C = d->src_C[0], cs = d->src_cstride[0], coff = d->src_coff[0];
  const int dil = p.d;
  const int per_res = p.tilesD * p.tilesH * p.tilesW;
  const int num_tiles = p.epi.B * dil * dil * dil * per_res;
  std::vector<float> box((size_t)rows_alloc * KC);
  std::vector<float> acc((size_t)p.nM * 128 * p.N_tile);
  for (int tile = 0; tile < num_tiles; ++tile) {
    int t = tile;
    const int tw = t % p.tilesW; t /= p.tilesW;
    const int th = t % p.tilesH; t /= p.tilesH;
    const int td = t % p.tilesD; t /= p.tilesD;
    const int rc_ = t % dil; t /= dil;
    const int rb = t % dil; t /= dil;
    const int ra = t % dil; t /= dil;
    const int b = t;
    // ---- TMA: box origin exactly as the producer computes it; element strides = dilation; OOB -> 0 ----
    const int x0 = (tw * p.BW - p.hw) * dil + rc_, y0 = (th * p.BH - p.hh) * dil + rb,
              z0 = (td * p.BD - p.hd) * dil + ra + p.src_d0;
    for (size_t i = 0; i < box.size(); ++i) box[i] = NAN;
    for (int pd = 0; pd < p.PD; ++pd)
      for (int ph = 0; ph < p.PH; ++ph)
        for (int pw = 0; pw < p.PW; ++pw) {
          const int z = z0 + pd * dil, y = y0 + ph * dil, x = x0 + pw * dil;
          const bool in = z >= 0 && z < d->ID && y >= 0 && y < d->IH && x >= 0 && x < d->IW;
          float* row = &box[(size_t)((pd * p.PH + ph) * p.PW + pw) * KC];
          for (int k = 0; k < KC; ++k)
            row[k] = (in && k < C) ? src[((((size_t)b * d->ID + z) * d->IH + y) * d->IW + x) * cs + coff + k] : 0.f;
        }
    // ---- MMA groups: accumulator row i of M tile m += A row (m*128 + i + off_g) . W_g^T ----
    for (size_t i = 0; i < acc.size(); ++i) acc[i] = 0.f;
    for (int m = 0; m < p.nM; ++m)
      for (int g = 0; g < p.n_taps; ++g) {
        const int off_rows = p.tap_off16[g] * 16 / row_bytes;
        if (p.tap_off16[g] * 16 % row_bytes) return 103;
        for (int i = 0; i < 128; ++i) {
          const int arow = m * 128 + i + off_rows;
          if (arow < 0 || arow >= rows_alloc) return 104;  // would read outside the allocated stage
          const float* a = &box[(size_t)arow * KC];
          for (int n = 0; n < p.N_tile; ++n) {
            const float* w = wgt + ((size_t)g * p.N_tile + n) * d->Kpad;
            float s = 0.f;
            for (int k = 0; k < KC; ++k) s += a[k] * w[k];
            acc[((size_t)m * 128 + i) * p.N_tile + n] += s;
          }
        }
      }
    // ---- epilogue: TMEM lane (= accumulator row) -> box position -> output position ----
    for (int m = 0; m < p.nM; ++m)
      for (int q = 0; q < 4; ++q)
        for (int lane = 0; lane < 32; ++lane) {
          const int i = q * 32 + lane;
          const int R = p.R0 + m * 128 + i;
          const int pw = R % p.PW;
          const int phh = (R / p.PW) % p.PH;
          const int pd = R / (p.PW * p.PH);
          const int od = (td * p.BD + pd - p.hd) * dil + ra, oh = (th * p.BH + phh - p.hh) * dil + rb,
                    ow = (tw * p.BW + pw - p.hw) * dil + rc_;
          const bool valid = pd >= p.hd && pd < p.hd + p.BD && phh >= p.hh && phh < p.hh + p.BH && pw >= p.hw &&
                             pw < p.hw + p.BW && od < p.D && oh < p.H && ow < p.W;
          if (xp && pw != lane) return 105;  // the shuffle epilogue needs lane == box column
          if (!valid) continue;
          for (int co = 0; co < d->Cout; ++co) {
            float v;
            if (xp) {
              const int up = lane > 0 ? i - 1 : i, dn = lane < 31 ? i + 1 : i;  // __shfl_up / __shfl_down by 1
              v = acc[((size_t)m * 128 + i) * p.N_tile + p.CP + co] +
                  (acc[((size_t)m * 128 + up) * p.N_tile + co] + acc[((size_t)m * 128 + dn) * p.N_tile + 2 * p.CP + co]);
            } else {
              v = acc[((size_t)m * 128 + i) * p.N_tile + co];
            }
            out[((((size_t)b * p.D + od) * p.H + oh) * p.W + ow) * d->Cout + co] = v + bias[co];
          }
        }
  }
  return 0;
}
