/*
 * occdepth_b200 -- C ABI of the B200 (sm_100a) kernels behind the OccDepth forward hot path.
 *
 * The reference (megvii-research/OccDepth) has no FFI: its "operator interface" for this path is the
 * sequence of PyTorch library calls inside occdepth/models/*.py.  Each entry point below names the
 * reference call sequence it replaces (file:line relative to the reference tree).
 *
 * Conventions
 *   - plain pointers + sizes only; all pointers are DEVICE pointers unless stated otherwise
 *   - nothing allocates or synchronises; kernels are enqueued on `stream` (a cudaStream_t cast to void*)
 *   - return 0 on success, an OCCD_ERR_* code otherwise; occd_last_error() returns a static message
 *   - activations are "channels-last": [B][D][H][W][C] with C contiguous (D == 1 for 2-D maps)
 */
#ifndef OCCDEPTH_B200_H
#define OCCDEPTH_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define OCCD_ABI_VERSION 1

#define OCCD_DTYPE_F32 0
#define OCCD_DTYPE_BF16 1

/* -------------------------------------------------------------------------------------------- */
/* library                                                                                      */
int occd_abi_version(void);
const char* occd_last_error(void);

/* -------------------------------------------------------------------------------------------- */
/* Stereo-SFA lift (the memory-bound 2D->3D gather)                                              */
/* replaces SFA.forward (occdepth/models/SFA.py:12-106) x len(project_res) as driven by          */
/* OccDepth._forward_2d_to_3d (occdepth/models/OccDepth.py:262-298) and the FlospDepth product   */
/* x3ds * x3ds_depth * 100 (OccDepth.py:339).                                                    */
#define OCCD_SFA_MAX_SCALES 4
#define OCCD_SFA_OUT_F32_PLANAR 0 /* [C][N]   == reference (C,X,Y,Z) fp32                  */
#define OCCD_SFA_OUT_BF16_CL 1    /* [N][cstride] bf16 channels-last (feeds the 3D net)    */
#define OCCD_SFA_OUT_F32_CL 2     /* [N][cstride] fp32 channels-last                       */

typedef struct {
  const void* feat[OCCD_SFA_MAX_SCALES]; /* per scale: [V][h][w][C] channels-last, feat_dtype      */
  int h[OCCD_SFA_MAX_SCALES];
  int w[OCCD_SFA_MAX_SCALES];
  int div[OCCD_SFA_MAX_SCALES]; /* projected_pix // div (floor), OccDepth.py:286-294               */
  int n_scales;                 /* 1..4                                                            */
  int n_views;                  /* V, 1..4                                                         */
  int C;                        /* channels, multiple of 4 (f32) / 8 (bf16), <= 256                */
  int feat_dtype;               /* OCCD_DTYPE_*                                                    */
  const int64_t* pix;           /* [V][N][P][2] (x,y) int64 -- batch["projected_pix_s"][i]         */
  const uint8_t* fov;           /* [V][N][P] bool          -- batch["fov_mask_s"][i]               */
  long long N;                  /* voxels                                                          */
  int P;                        /* pattern points per voxel (1 for pattern_id 0)                   */
  void* out;
  int out_mode;    /* OCCD_SFA_OUT_*                                                               */
  int out_cstride; /* channel stride of channels-last outputs (>= C)                               */
  int perm_nyu;    /* 1: voxel n=(i*S2+k)*S1+j is written at (i*S1+j)*S2+k (SFA.py:90-97)          */
  int S1, S2;
  const float* prior; /* optional [N] per-voxel multiplier (FlospDepth occupancy prior), or NULL   */
  float scale_const;  /* out = out * prior * scale_const when prior != NULL                        */
} occd_sfa_params;

int occd_sfa_lift_fwd(const occd_sfa_params* p, void* stream);

/* layout helpers at the module boundary: NCHW/NCDHW fp32 <-> channels-last (f32 or bf16)          */
/* in: [B][C][S] planar fp32, out: [B][S][cstride] channels-last (pad channels zero-filled)        */
int occd_planar_to_cl(const float* in, void* out, int out_dtype, long long B, int C, long long S,
                      int cstride, void* stream);
/* in: [B][S][cstride] channels-last, out: [B][C][S] planar fp32                                   */
int occd_cl_to_planar(const void* in, int in_dtype, float* out, long long B, int C, long long S,
                      int cstride, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OCCDEPTH_B200_H */
