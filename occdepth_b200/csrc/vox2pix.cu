// Data-pipeline kernels.
// occd_vox2pix_fwd: the data pipeline's voxel -> pixel index generation on the device.
// replaces occdepth/data/utils/helpers.py:94-169 (vox2pix) as called per view and per scale by the datasets
// (data/semantic_kitti/kitti_dataset.py, data/NYU/nyu_dataset.py).
#include "vox2pix.cuh"
#include "normalize_rgb.cuh"
#include "../../include/occdepth_b200.h"

namespace {

template <typename T>
int run(const void* cam_E, const float* cam_k, const float* vox_origin, double voxel_size, int X, int Y, int Z,
        int img_W, int img_H, const int* pattern, int P, long long* pix, unsigned char* fov, void* pix_z,
        cudaStream_t st) {
  v2p::Args<T> a;
  for (int i = 0; i < 12; ++i) a.E[i] = static_cast<const T*>(cam_E)[i];
  a.fx = cam_k[0]; a.fy = cam_k[4]; a.cx = cam_k[2]; a.cy = cam_k[5];
  for (int i = 0; i < 3; ++i) a.origin[i] = vox_origin[i];
  a.voxel_size = voxel_size;
  a.X = X; a.Y = Y; a.Z = Z; a.W = img_W; a.H = img_H; a.P = P;
  for (int p = 0; p < P; ++p) { a.pat[p][0] = pattern[2 * p]; a.pat[p][1] = pattern[2 * p + 1]; }
  a.pix = pix; a.fov = fov; a.pix_z = static_cast<T*>(pix_z);
  const long long N = (long long)X * Y * Z;
  v2p::vox2pix_kernel<T><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(a, N);
  OCCD_CHECK_LAUNCH();
  return OCCD_OK;
}

}  // namespace

extern "C" int occd_vox2pix_fwd(const void* cam_E, int pose_is_f32, const float* cam_k, const float* vox_origin,
                                double voxel_size, int X, int Y, int Z, int img_W, int img_H, const int* pattern,
                                int P, long long* pix, unsigned char* fov, void* pix_z, void* stream) {
  OCCD_CHECK_ARG(cam_E && cam_k && vox_origin && pattern && pix && fov, "occd_vox2pix_fwd: null argument");
  OCCD_CHECK_ARG(X > 0 && Y > 0 && Z > 0 && img_W > 0 && img_H > 0 && voxel_size > 0, "occd_vox2pix_fwd: dims");
  OCCD_CHECK_ARG(P >= 1 && P <= v2p::kMaxPattern, "occd_vox2pix_fwd: pattern size must be 1..25");
  OCCD_CHECK_ARG(((long long)X * Y * Z + 255) / 256 <= 2147483647LL, "occd_vox2pix_fwd: too many voxels");
  cudaStream_t st = (cudaStream_t)stream;
  return pose_is_f32 ? run<float>(cam_E, cam_k, vox_origin, voxel_size, X, Y, Z, img_W, img_H, pattern, P, pix, fov,
                                  pix_z, st)
                     : run<double>(cam_E, cam_k, vox_origin, voxel_size, X, Y, Z, img_W, img_H, pattern, P, pix, fov,
                                   pix_z, st);
}

// occd_normalize_rgb_u8: uint8 HWC image -> cropped, normalised float32 CHW (the datasets' `normalize_rgb`,
// data/semantic_kitti/kitti_dataset.py:164-171,376-402)
extern "C" int occd_normalize_rgb_u8(const void* in, float* out, int H0, int W0, int H, int W, const float* mean,
                                     const float* stdv, void* stream) {
  OCCD_CHECK_ARG(in && out && mean && stdv, "occd_normalize_rgb_u8: null argument");
  OCCD_CHECK_ARG(H > 0 && W > 0 && H <= H0 && W <= W0, "occd_normalize_rgb_u8: the crop must lie inside the image");
  nrm::Args a;
  a.in = (const unsigned char*)in; a.out = out; a.W0 = W0; a.H = H; a.W = W;
  for (int c = 0; c < 3; ++c) { a.mean[c] = mean[c]; a.stdv[c] = stdv[c]; }
  const long long total = (long long)H * W;
  nrm::normalize_rgb_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, total);
  OCCD_CHECK_LAUNCH();
  return OCCD_OK;
}
