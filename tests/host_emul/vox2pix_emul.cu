// CPU emulation harness for csrc/vox2pix.cuh: the kernel body is __host__ __device__ IEEE double arithmetic with
// explicit rounding / fma, so running it on the host reproduces the device result bit for bit.
#include "vox2pix.cuh"

extern "C" void occd_set_last_error(const char*) {}

template <typename T>
static int run(const void* cam_E, const float* cam_k, const float* vox_origin, double voxel_size, int X, int Y, int Z,
               int img_W, int img_H, const int* pattern, int P, long long* pix, unsigned char* fov, void* pix_z) {
  v2p::Args<T> a;
  for (int i = 0; i < 12; ++i) a.E[i] = static_cast<const T*>(cam_E)[i];
  a.fx = cam_k[0]; a.fy = cam_k[4]; a.cx = cam_k[2]; a.cy = cam_k[5];
  for (int i = 0; i < 3; ++i) a.origin[i] = vox_origin[i];
  a.voxel_size = voxel_size;
  a.X = X; a.Y = Y; a.Z = Z; a.W = img_W; a.H = img_H; a.P = P;
  for (int p = 0; p < P; ++p) { a.pat[p][0] = pattern[2 * p]; a.pat[p][1] = pattern[2 * p + 1]; }
  a.pix = pix; a.fov = fov; a.pix_z = static_cast<T*>(pix_z);
  const long long N = (long long)X * Y * Z;
  for (long long n = 0; n < N; ++n) v2p::body(a, n);
  return 0;
}

extern "C" int vox2pix_emulate(const void* cam_E, int pose_is_f32, const float* cam_k, const float* vox_origin,
                               double voxel_size, int X, int Y, int Z, int img_W, int img_H, const int* pattern, int P,
                               long long* pix, unsigned char* fov, void* pix_z) {
  if (P < 1 || P > v2p::kMaxPattern) return 1;
  return pose_is_f32 ? run<float>(cam_E, cam_k, vox_origin, voxel_size, X, Y, Z, img_W, img_H, pattern, P, pix, fov, pix_z)
                     : run<double>(cam_E, cam_k, vox_origin, voxel_size, X, Y, Z, img_W, img_H, pattern, P, pix, fov, pix_z);
}
