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
  __nv_bfloat16* out0;
  int out0_cstride, out0_coff;
  int act;
  const __nv_bfloat16* res1;
  int res1_cstride, res1_coff;
  const __nv_bfloat16* res2;
  int res2_cstride, res2_coff, res2_post;
  int out1_mode;
  void* out1;
  int out1_cstride, out1_coff, out1_C;
  int wide;  // 1: every channels-last window is 32-byte aligned -> the WIDE kernel instance may be launched
};

__device__ __forceinline__ void ld256(const void* p, uint32_t* r) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void st256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void unpack16(const uint32_t* r, float* f) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    f[2 * i] = __uint_as_float(r[i] << 16);
    f[2 * i + 1] = __uint_as_float(r[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void pack16(const float* f, uint32_t* r) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    r[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
}

// Epilogue for one output position and NV consecutive channels [n0, n0+NV) held in v[] (fp32 accumulators).
// NV is 8 or 16; n0 is a multiple of 8; channels >= Cout_store are not written.
// WIDE is a compile-time switch (separate kernel instances) so that the default kernels' register allocation is
// untouched by the 256-bit path.
template <int NV, bool WIDE = false>
__device__ __forceinline__ void conv_epilogue_row(const ConvEpi& e, int b, int od, int oh, int ow, int n0,
                                                  float* v) {
  const long long pos = (((long long)b * e.ODf + ((long long)od * e.omul[0] + e.oadd[0])) * e.OHf +
                         ((long long)oh * e.omul[1] + e.oadd[1])) * e.OWf +
                        ((long long)ow * e.omul[2] + e.oadd[2]);
  if constexpr (NV == 16 && WIDE) {
    if (n0 + 16 <= e.Cout_store) {
      // 256-bit path (opt-in, OCCD_EPI_WIDE=1): identical arithmetic, half the load/store instructions and no
      // half-written 32-byte sectors on the way to L2
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const float4 bq = __ldg(reinterpret_cast<const float4*>(e.bias + n0 + 4 * q4));
        v[4 * q4] += bq.x; v[4 * q4 + 1] += bq.y; v[4 * q4 + 2] += bq.z; v[4 * q4 + 3] += bq.w;
      }
      uint32_t raw[8];
      if (e.res1) {
        float r[16];
        ld256(e.res1 + pos * e.res1_cstride + e.res1_coff + n0, raw);
        unpack16(raw, r);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += r[i];
      }
      float r2[16];
      if (e.res2) {
        ld256(e.res2 + pos * e.res2_cstride + e.res2_coff + n0, raw);
        unpack16(raw, r2);
        if (!e.res2_post) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += r2[i];
        }
      }
      if (e.out1_mode == OCCD_OUT1_BF16_CL) {
        pack16(v, raw);
        st256(reinterpret_cast<__nv_bfloat16*>(e.out1) + pos * e.out1_cstride + e.out1_coff + n0, raw);
      } else if (e.out1_mode == OCCD_OUT1_F32_PLANAR) {
        const long long S = (long long)e.ODf * e.OHf * e.OWf;
        const long long sp = pos - (long long)b * S;
        float* o = reinterpret_cast<float*>(e.out1) + ((long long)b * e.out1_C + e.out1_coff + n0) * S + sp;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (n0 + i < e.Cout) o[(long long)i * S] = v[i];
      }
      if (e.out0) {
        apply_act8(v, e.act);
        apply_act8(v + 8, e.act);
        if (e.res2 && e.res2_post) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += r2[i];
        }
        pack16(v, raw);
        st256(e.out0 + pos * e.out0_cstride + e.out0_coff + n0, raw);
      }
      return;
    }
  }
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
      unpack8(*reinterpret_cast<const uint4*>(e.res1 + pos * e.res1_cstride + e.res1_coff + n), r);
#pragma unroll
      for (int i = 0; i < 8; ++i) vv[i] += r[i];
    }
    float r2[8];
    if (e.res2) {
      unpack8(*reinterpret_cast<const uint4*>(e.res2 + pos * e.res2_cstride + e.res2_coff + n), r2);
      if (!e.res2_post) {
#pragma unroll
        for (int i = 0; i < 8; ++i) vv[i] += r2[i];
      }
    }
    if (e.out1_mode == OCCD_OUT1_BF16_CL) {
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.out1) + pos * e.out1_cstride + e.out1_coff + n) =
          pack8(vv);
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
      *reinterpret_cast<uint4*>(e.out0 + pos * e.out0_cstride + e.out0_coff + n) = pack8(vv);
    }
  }
}
