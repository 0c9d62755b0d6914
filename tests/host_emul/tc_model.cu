// CPU model of conv_tc_kernel (csrc/conv.cu) driven by the real host-side plan geometry (tc_geometry): tile
// decode, TMA box origins / traversal strides / zero fill, (tap | tap-group, k-chunk) pipeline items, weight row
// addressing (N tiles, per-image weight sets), TMEM-lane <-> tile-position mapping, validity mask and -- for the
// x-packed variant -- the lane-shifted epilogue sum.  Same caveats as halo_model.cu: an MMA is modelled as
// "accumulator row i += A row i . B^T", a TMA load as a strided gather with zero fill.
#include <cmath>
#include <vector>
#include "../../occdepth_b200/csrc/conv.cu"

static thread_local char g_err[512];
extern "C" void occd_set_last_error(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg ? msg : ""); }
extern "C" const char* tc_model_error() { return g_err; }

// desc: src[s] -> float [B][ID][IH][IW][src_cstride], weight -> float [(B if per image)][n_taps][Cout_pad][Kpad],
// bias -> float;  out: float [B][ODf][OHf][OWf][Cout], pre-filled by the caller
// info (optional, 8 ints): TD TH TW N_tile n_items_per_tile stages group grid
extern "C" int tc_model(const occd_conv_desc* d, int mode, float* out, int* info) {
  const int xp = mode == 1 ? 1 : 0;
  const bool m2 = mode == 2;
  occd_conv_plan pl;
  memset(&pl, 0, sizeof(pl));
  const int rc = tc_geometry(d, &pl, mode);  // 0 per-tap, 1 x-packed, 2 M2
  if (rc != OCCD_OK) return rc;
  const TcParams& p = pl.tc;
  const int KC = pl.kc;
  int iters_per_tile = 0;
  for (int i = 0; i < p.n_taps; ++i) iters_per_tile += p.n_kchunks[p.tap_src[i]];
  if (info) {
    const int v[8] = {p.TD, p.TH, p.TW, p.N_tile, iters_per_tile, p.stages, p.group, (int)pl.grid.x};
    for (int i = 0; i < 8; ++i) info[i] = v[i];
  }
  if (p.TD * p.TH * p.TW != 128) return 100;
  if (pl.smem > 227 * 1024 || p.stages < 1 || p.group < 1 || p.tmem_cols > 512 || 2 * p.N_tile > p.tmem_cols) return 101;
  if (m2) {
    // kernel-side carve-up: stages * group * (2 * a_bytes + b_stride) + barriers must fit what the host asked for
    if (p.m2_sets < 1 || p.m2_sets > 2 || p.m2_sets * 2 * p.N_tile > p.tmem_cols) return 104;
    if ((size_t)p.stages * p.group * (2 * p.a_bytes + p.b_stride) + 16 * kMaxStages + 64 + 1024 > pl.smem) return 105;
    const long long pairs = ((long long)p.num_m_tiles + 1) / 2 * (p.Cout_pad / p.N_tile);
    if ((long long)pl.grid.x > pairs) return 106;
  } else if (p.m2_sets != 0) {
    return 107;
  }
  if (xp && (p.TW != 32 || p.TWv != 30 || p.N_tile != 3 * p.Cout_pad)) return 102;
  if (!out) return 0;
  const float* wgt = (const float*)d->weight;
  const float* bias = (const float*)d->bias;
  const int n_tiles_n = xp ? 1 : p.Cout_pad / p.N_tile;
  // M2: the kernel walks pairs (2*mp, 2*mp+1) of M tiles; per tile the arithmetic is the per-tap kernel's, except
  // that an odd tile count adds one phantom tile (decodes to b == B: zero-filled loads, masked stores)
  const int m_tiles_walked = m2 ? ((p.num_m_tiles + 1) / 2) * 2 : p.num_m_tiles;
  const int num_tiles = m_tiles_walked * n_tiles_n;
  std::vector<float> acc((size_t)128 * p.N_tile), A((size_t)128 * KC);
  for (int tile = 0; tile < num_tiles; ++tile) {
    const int nt = tile % n_tiles_n;
    int t = tile / n_tiles_n;
    const int tw = t % p.tiles_w; t /= p.tiles_w;
    const int th = t % p.tiles_h; t /= p.tiles_h;
    const int td = t % p.tiles_d; t /= p.tiles_d;
    const int b = t;
    const int iw0 = xp ? tw * 30 - 1 : tw * p.TW * p.stride[2], ih0 = th * p.TH * p.stride[1],
              id0 = td * p.TD * p.stride[0] + p.src_d0;
    const int n0 = nt * p.N_tile;
    for (size_t i = 0; i < acc.size(); ++i) acc[i] = 0.f;
    for (int tp = 0; tp < p.n_taps; ++tp) {
      const int src = p.tap_src[tp];
      const int cw = iw0 + p.tap_dx[tp], ch = ih0 + p.tap_dy[tp], cd = id0 + p.tap_dz[tp];
      const int wrow = (m2 ? 0 : b * p.w_batch_rows) + (xp ? tp * p.N_tile : tp * p.Cout_pad + n0);
      if (m2 && p.w_batch_rows != 0) return 108;
      const float* s = (const float*)d->src[src];
      const int C = d->src_C[src], cs = d->src_cstride[src], coff = d->src_coff[src];
      for (int kc = 0; kc < p.n_kchunks[src]; ++kc) {
        // TMA box {KC, TW*sx, TH*sy, TD*sz, 1} with traversal strides (sx, sy, sz): smem row (rd*TH + rh)*TW + rw
        for (int rd = 0; rd < p.TD; ++rd)
          for (int rh = 0; rh < p.TH; ++rh)
            for (int rw = 0; rw < p.TW; ++rw) {
              const int x = cw + rw * p.stride[2], y = ch + rh * p.stride[1], z = cd + rd * p.stride[0];
              const bool in = b < d->B && z >= 0 && z < d->ID && y >= 0 && y < d->IH && x >= 0 && x < d->IW;
              float* row = &A[(size_t)((rd * p.TH + rh) * p.TW + rw) * KC];
              for (int k = 0; k < KC; ++k) {
                const int c = kc * KC + k;
                row[k] = (in && c < C) ? s[((((size_t)b * d->ID + z) * d->IH + y) * d->IW + x) * cs + coff + c] : 0.f;
              }
            }
        for (int i = 0; i < 128; ++i)
          for (int n = 0; n < p.N_tile; ++n) {
            const float* w = wgt + (size_t)(wrow + n) * d->Kpad + kc * KC;
            float sacc = 0.f;
            for (int k = 0; k < KC; ++k) sacc += A[(size_t)i * KC + k] * w[k];
            acc[(size_t)i * p.N_tile + n] += sacc;
          }
      }
    }
    for (int q = 0; q < 4; ++q)
      for (int lane = 0; lane < 32; ++lane) {
        const int row = q * 32 + lane;
        const int rw = row % p.TW;
        const int rh = (row / p.TW) % p.TH;
        const int rd = row / (p.TW * p.TH);
        const int od = td * p.TD + rd, oh = th * p.TH + rh, ow = xp ? tw * 30 + rw - 1 : tw * p.TW + rw;
        const bool valid = b < p.epi.B && od < p.epi.OD && oh < p.epi.OH && ow < p.epi.OW &&
                           (!xp || (rw >= 1 && rw <= 30));
        if (xp && rw != lane) return 103;
        if (!valid) continue;
        const size_t pos = (((size_t)b * p.epi.ODf + ((size_t)od * p.epi.omul[0] + p.epi.oadd[0])) * p.epi.OHf +
                            ((size_t)oh * p.epi.omul[1] + p.epi.oadd[1])) * p.epi.OWf +
                           ((size_t)ow * p.epi.omul[2] + p.epi.oadd[2]);
        const int nlim = xp ? p.Cout_pad : p.N_tile;
        for (int c = 0; c < nlim; ++c) {
          const int n = (xp ? 0 : n0) + c;
          if (n >= d->Cout) break;
          float v;
          if (xp) {
            const int up = lane > 0 ? row - 1 : row, dn = lane < 31 ? row + 1 : row;
            v = acc[(size_t)row * p.N_tile + p.Cout_pad + c] +
                (acc[(size_t)up * p.N_tile + c] + acc[(size_t)dn * p.N_tile + 2 * p.Cout_pad + c]);
          } else {
            v = acc[(size_t)row * p.N_tile + c];
          }
          out[pos * d->Cout + n] = v + bias[n];
        }
      }
  }
  return 0;
}
