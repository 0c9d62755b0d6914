// Device-side parameter block + shared epilogue for the implicit-GEMM convolution kernels.
#pragma once
#include "common.cuh"
#include "../../include/occdepth_b200.h"

struct ConvEpi {
  // iteration space / output mapping
  int B, OD, OH, OW;
  int omul[3], oadd[3];
  int ODf, OHf, OWf;
  int Cout;  // real output channels (stores are masked to n < Cout_store)
  int Cout_store;  // channels physically written (Cout rounded up to 8, pad written as computed zeros)
  const float* bias;
  void* out0;   // channels-last, element type T of the plan
  int out0_cstride, out0_coff;
  int act;
  int out0_exact;
  const void* res1;
  int res1_cstride, res1_coff;
  const void* res2;
  int res2_cstride, res2_coff, res2_post;
  int out1_mode;
  void* out1;
  int out1_cstride, out1_coff, out1_C;
  int dbg;   // experiment switch (OCCD_DEBUG_EPI, tools/conv_bench.py only): 1 = skip the out0 stores, 2 = skip the
             // whole row epilogue (TMEM drain only)
};

__device__ __forceinline__ long long epi_pos(const ConvEpi& e, int b, int od, int oh, int ow) {
  return (((long long)b * e.ODf + ((long long)od * e.omul[0] + e.oadd[0])) * e.OHf +
          ((long long)oh * e.omul[1] + e.oadd[1])) * e.OWf +
         ((long long)ow * e.omul[2] + e.oadd[2]);
}

// Epilogue arithmetic for one output position `pos` and NV consecutive channels [n0, n0+NV) held in v[] (fp32
// accumulators): v = acc + bias + res1 (+ res2); the optional second output (pre-activation value: channels-last or
// fp32 planar) is written here; on return v[] holds the out0 values act(v) (+ res2 when it is added after the
// activation).  NV is 8 or 16; n0 is a multiple of 8.  T is the activation element type.
template <typename T, int NV>
__device__ __forceinline__ void conv_epilogue_compute(const ConvEpi& e, long long pos, int b, int n0, float* v) {
#pragma unroll
  for (int g = 0; g < NV; g += 8) {
    const int n = n0 + g;
    if (n >= e.Cout_store) break;
    float* vv = v + g;
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(e.bias + n));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(e.bias + n + 4));
    vv[0] += b0.x; vv[1] += b0.y; vv[2] += b0.z; vv[3] += b0.w;
    vv[4] += b1.x; vv[5] += b1.y; vv[6] += b1.z; vv[7] += b1.w;
    if (e.res1) {
      float r[8];
      Elem<T>::ld8(reinterpret_cast<const T*>(e.res1) + pos * e.res1_cstride + e.res1_coff + n, r);
#pragma unroll
      for (int i = 0; i < 8; ++i) vv[i] += r[i];
    }
    float r2[8];
    if (e.res2) {
      Elem<T>::ld8(reinterpret_cast<const T*>(e.res2) + pos * e.res2_cstride + e.res2_coff + n, r2);
      if (!e.res2_post) {
#pragma unroll
        for (int i = 0; i < 8; ++i) vv[i] += r2[i];
      }
    }
    if (e.out1_mode == OCCD_OUT1_CL) {
      Elem<T>::st8(reinterpret_cast<T*>(e.out1) + pos * e.out1_cstride + e.out1_coff + n, vv);
    } else if (e.out1_mode == OCCD_OUT1_F32_PLANAR) {
      const long long S = (long long)e.ODf * e.OHf * e.OWf;
      const long long sp = pos - (long long)b * S;
      float* o = reinterpret_cast<float*>(e.out1) + ((long long)b * e.out1_C + e.out1_coff + n) * S + sp;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (n + i < e.Cout) o[(long long)i * S] = vv[i];
    }
    if (e.out0) {
      apply_act8(vv, e.act);
      if (e.res2 && e.res2_post) {
#pragma unroll
        for (int i = 0; i < 8; ++i) vv[i] += r2[i];
      }
    }
  }
}

// compute + direct (per-thread, 8-channel vector) global stores of out0: SIMT kernel, halo kernel
template <typename T, int NV>
__device__ __forceinline__ void conv_epilogue_row(const ConvEpi& e, int b, int od, int oh, int ow, int n0,
                                                  float* v) {
  if (e.dbg == 2) return;
  const long long pos = epi_pos(e, b, od, oh, ow);
  conv_epilogue_compute<T, NV>(e, pos, b, n0, v);
  if (!e.out0) return;
  if (e.dbg == 1) {   // keep the arithmetic alive without storing
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += v[i];
    if (s == 123456.789f) reinterpret_cast<T*>(e.out0)[0] = Elem<T>::cvt(s);
    return;
  }
#pragma unroll
  for (int g = 0; g < NV; g += 8) {
    const int n = n0 + g;
    if (n >= e.Cout_store) break;
    T* o0 = reinterpret_cast<T*>(e.out0) + pos * e.out0_cstride + e.out0_coff + n;
    if (e.out0_exact) Elem<T>::st8_exact(o0, v + g);
    else Elem<T>::st8(o0, v + g);
  }
}

