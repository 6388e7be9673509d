// N2 -- pinhole ray generation on the device (the reference builds its rays on the host with numpy:
// Carla/Waymo._generate_rays_multi, internal/obbpose_dataset.py:613-661).  One thread per pixel:
//   camera_dir = ((x - cx)/f, -(y - cy)/f, -1) with (cx, cy) = (W/2, H/2) (Carla) or the file's principal point (Waymo, :1882-1885);  direction_i = sum_j camera_dir_j * c2w[i][j]  (un-normalised);
//   origin = c2w[:, 3];  viewdir = direction / |direction|;
//   radius = |direction(y) - direction(y+1)| * 2 / sqrt(12); the last row reuses the spacing |dir(H-3) - dir(H-2)|, which is
//   what `np.concatenate([v, v[-2:-1, :]], 0)` appends (:640-646).
// HBM-bound: 52 B written per ray, nothing read; removes the 52 B/ray host->device copy of a frame.
#include "common.cuh"

namespace durf {

struct RayGenParams {
  int W, H, row0, row1;
  float focal, near, far;
  float pcx, pcy;            // principal point in pixels
  float c2w[12];
  float* origins; float* directions; float* viewdirs; float* radii; float* lossmult; float* near_out; float* far_out;
};

__device__ __forceinline__ void world_dir(const RayGenParams& p, float x, float y, float (&d)[3]) {
  const float cx = (x - p.pcx) / p.focal;
  const float cy = -(y - p.pcy) / p.focal;
  const float cz = -1.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) d[i] = (cx * p.c2w[4 * i] + cy * p.c2w[4 * i + 1]) + cz * p.c2w[4 * i + 2];
}

__global__ void __launch_bounds__(256)
raygen_kernel(const RayGenParams p) {
  const int64_t n = (int64_t)(p.row1 - p.row0) * p.W;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int yy = p.row0 + (int)(i / p.W), xx = (int)(i % p.W);
  const float x = (float)xx, y = (float)yy;
  float d[3], a[3], b[3];
  world_dir(p, x, y, d);
  // the reference appends v[-2:-1] of the (H-1)-row difference array, i.e. the LAST image row reuses |dir(H-3) - dir(H-2)|
  const float y0 = (yy >= p.H - 1) ? y - 2.f : y;
  world_dir(p, x, y0, a);
  world_dir(p, x, y0 + 1.f, b);
  const float e0 = a[0] - b[0], e1 = a[1] - b[1], e2 = a[2] - b[2];
  const float dx = sqrtf((e0 * e0 + e1 * e1) + e2 * e2);
  const float nrm = sqrtf((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    p.origins[3 * i + c] = p.c2w[4 * c + 3];
    p.directions[3 * i + c] = d[c];
    p.viewdirs[3 * i + c] = d[c] / nrm;
  }
  p.radii[i] = (dx * 2.f) / 3.46410161513775459f;
  p.lossmult[i] = 1.f;
  p.near_out[i] = p.near;
  p.far_out[i] = p.far;
}

}  // namespace durf

using namespace durf;

extern "C" int durf_generate_rays(durf_stream_t stream, const DurfCamera* cam, int32_t row0, int32_t row1, float* origins,
                                  float* directions, float* viewdirs, float* radii, float* lossmult, float* near, float* far) {
  DURF_REQUIRE(cam && cam->width >= 1 && cam->height >= 3 && cam->focal > 0.f, DURF_E_INVALID, "durf_generate_rays: bad camera");
  DURF_REQUIRE(row0 >= 0 && row1 >= row0 && row1 <= cam->height, DURF_E_INVALID, "durf_generate_rays: bad row range [%d,%d)", row0, row1);
  if (row1 == row0) return DURF_OK;
  DURF_REQUIRE(origins && directions && viewdirs && radii && lossmult && near && far, DURF_E_INVALID, "durf_generate_rays: null buffer");
  RayGenParams p;
  p.W = cam->width; p.H = cam->height; p.row0 = row0; p.row1 = row1; p.focal = cam->focal; p.near = cam->near; p.far = cam->far;
  p.pcx = cam->use_principal_point ? cam->cx : (float)cam->width * 0.5f;
  p.pcy = cam->use_principal_point ? cam->cy : (float)cam->height * 0.5f;
  for (int i = 0; i < 12; ++i) p.c2w[i] = cam->c2w[i];
  p.origins = origins; p.directions = directions; p.viewdirs = viewdirs; p.radii = radii; p.lossmult = lossmult; p.near_out = near; p.far_out = far;
  const int64_t n = (int64_t)(row1 - row0) * cam->width;
  raygen_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(p);
  DURF_CHECK_LAUNCH("durf_generate_rays");
  return DURF_OK;
}
