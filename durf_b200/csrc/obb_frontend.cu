// K0 -- OBB front-end: axis-angle -> rotation, world -> object ray transform, slab test, scene-graph merge.
// Replaces box_helpers.aa2matrix / world2object_rpy / ray_box_intersection (box_helpers.py:148-167, 286-341,
// 59-106) and the merge at obbpose_model.py:99-131.  One thread per ray, K rotations staged in shared memory.
// HBM-bound: 24 B in, 28 + 12K B out per ray.
#include "common.cuh"
#include <algorithm>

namespace durf {

constexpr int kMaxObjects = 64;

struct BoxFrame {
  float R[9];
  float t[3];   // R * (-p): origin of the world system in the object system (box_helpers.py:323)
};

// box_helpers.py:148-167.  theta = sqrt(max(|r|^2, 1e-12)) + 1e-12 (math.safe_norm, math.py:27-32).
__device__ __forceinline__ void rodrigues(const float* __restrict__ aa, float* R) {
  const float x = aa[0], y = aa[1], z = aa[2];
  float sq = x * x + y * y + z * z;
  sq = sq < 1e-12f ? 1e-12f : sq;
  const float th = sqrtf(sq) + 1e-12f;
  const float a = sinf(th) / th;
  const float b = (1.f - cosf(th)) / (th * th);
  // S = [[0,-z,y],[z,0,-x],[-y,x,0]],  S@S written out term by term
  const float s2[9] = {(-z) * z + y * (-y), y * x, (-z) * (-x),
                       (-x) * (-y), z * (-z) + (-x) * x, z * y,
                       x * z, (-y) * (-z), (-y) * y + x * (-x)};
  const float s[9] = {0.f, -z, y, z, 0.f, -x, -y, x, 0.f};
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = ((i % 4 == 0) ? 1.f : 0.f) + a * s[i] + b * s2[i];
}

__device__ __forceinline__ void make_frame(const float* __restrict__ box6, BoxFrame& f) {
  rodrigues(box6 + 3, f.R);
#pragma unroll
  for (int i = 0; i < 3; ++i)
    f.t[i] = f.R[i * 3 + 0] * (-box6[0]) + f.R[i * 3 + 1] * (-box6[1]) + f.R[i * 3 + 2] * (-box6[2]);
}

// jnp.minimum / jnp.maximum propagate NaN (fminf/fmaxf do not).
__device__ __forceinline__ float nmin(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fminf(a, b); }
__device__ __forceinline__ float nmax(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b); }

__global__ void aa2matrix_kernel(int K, const float* __restrict__ angles, float* __restrict__ R) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float r[9];
  rodrigues(angles + 3 * k, r);
#pragma unroll
  for (int i = 0; i < 9; ++i) R[9 * k + i] = r[i];
}

__global__ void __launch_bounds__(256)
obb_frontend_kernel(int B, int K, const float* __restrict__ origins, const float* __restrict__ dirs,
                    const float* __restrict__ box, const float* __restrict__ ext,
                    float* __restrict__ origins_s, float* __restrict__ dirs_s, int32_t* __restrict__ hit,
                    float* __restrict__ zi, float* __restrict__ zo, float* __restrict__ zo_ret,
                    float* __restrict__ nhit, float* __restrict__ origins_o, float* __restrict__ dirs_o) {
  __shared__ BoxFrame frames[kMaxObjects];
  __shared__ float s_ext[kMaxObjects * 3];
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    make_frame(box + 6 * k, frames[k]);
    s_ext[3 * k + 0] = ext[3 * k + 0];
    s_ext[3 * k + 1] = ext[3 * k + 1];
    s_ext[3 * k + 2] = ext[3 * k + 2];
  }
  __syncthreads();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float o[3] = {origins[3 * b], origins[3 * b + 1], origins[3 * b + 2]};
  const float d[3] = {dirs[3 * b], dirs[3 * b + 1], dirs[3 * b + 2]};
  float so[3] = {0.f, 0.f, 0.f}, sd[3] = {0.f, 0.f, 0.f};
  float zsum = 0.f;
  int hsum = 0;
  for (int k = 0; k < K; ++k) {
    const BoxFrame& f = frames[k];
    float oo[3], dd[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      oo[i] = (f.R[3 * i] * o[0] + f.R[3 * i + 1] * o[1] + f.R[3 * i + 2] * o[2]) + f.t[i];
      dd[i] = f.R[3 * i] * d[0] + f.R[3 * i + 1] * d[1] + f.R[3 * i + 2] * d[2];
    }
    const float nrm = sqrtf(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) dd[i] = dd[i] / nrm;           // box_helpers.py:340: unit direction
    // slab test against [-ext, +ext] (box_helpers.py:79-98)
    float tn = 0.f, tf = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float inv = 1.f / dd[i];
      const float e = s_ext[3 * k + i];
      const float tmin = (-e - oo[i]) * inv;
      const float tmax = (e - oo[i]) * inv;
      const float t0 = nmin(tmin, tmax), t1 = nmax(tmin, tmax);
      tn = (i == 0) ? t0 : nmax(tn, t0);
      tf = (i == 0) ? t1 : nmin(tf, t1);
    }
    int h = (tf > tn) ? 1 : 0;
    h *= ((tf * (float)h) > 0.f) ? 1 : 0;
    const float hf = (float)h;
    hit[(size_t)b * K + k] = h;
    zi[(size_t)b * K + k] = tn * hf;
    zo[(size_t)b * K + k] = tf * hf;
    zsum += hf * (tf * hf);
    hsum += h;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      so[i] += oo[i] * hf;
      sd[i] += dd[i] * hf;
    }
    if (origins_o) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        origins_o[((size_t)b * K + k) * 3 + i] = oo[i];
        dirs_o[((size_t)b * K + k) * 3 + i] = dd[i];
      }
    }
  }
  const float bk = (hsum == 0) ? 1.f : 0.f;                     // obbpose_model.py:115
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    origins_s[3 * b + i] = so[i] + bk * o[i];
    dirs_s[3 * b + i] = sd[i] + bk * d[i];
  }
  zo_ret[b] = zsum;
  nhit[b] = (float)hsum;
}

// box_helpers.world2object_rpy(pts, dirs, pose, rot) with EXPLICIT rotation matrices (box_helpers.py:286-341, dim=None,
// inverse=False): o_o = R o + R (-p), d_o = R d / |R d|.  pose / rot may be per object ([K,..], stride 0 over rays) or
// per ray and object ([B,K,..]).  One thread per (ray, object).
__global__ void __launch_bounds__(256)
world2object_kernel(int B, int K, const float* __restrict__ pts, const float* __restrict__ dirs,
                    const float* __restrict__ pose, int64_t pose_stride, const float* __restrict__ rot, int64_t rot_stride,
                    float* __restrict__ pts_o, float* __restrict__ dirs_o) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * K) return;
  const int b = (int)(i / K), k = (int)(i % K);
  const float* R = rot + (int64_t)b * rot_stride + 9 * k;
  const float* p = pose + (int64_t)b * pose_stride + 3 * k;
  const float o[3] = {pts[3 * b], pts[3 * b + 1], pts[3 * b + 2]};
  const float d[3] = {dirs[3 * b], dirs[3 * b + 1], dirs[3 * b + 2]};
  float oo[3], dd[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float t = R[3 * r] * (-p[0]) + R[3 * r + 1] * (-p[1]) + R[3 * r + 2] * (-p[2]);     // rotate_matrix(-pose_w, rot)
    oo[r] = (R[3 * r] * o[0] + R[3 * r + 1] * o[1] + R[3 * r + 2] * o[2]) + t;
    dd[r] = R[3 * r] * d[0] + R[3 * r + 1] * d[1] + R[3 * r + 2] * d[2];
  }
  const float nrm = sqrtf(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2]);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    pts_o[3 * i + r] = oo[r];
    dirs_o[3 * i + r] = dd[r] / nrm;
  }
}

// box_helpers.ray_box_intersection (box_helpers.py:59-106) on n = B*K (ray, box) pairs: slab test, NaN-propagating
// min / max like jnp, intersection = (t_far > t_near) * (t_far * intersection > 0).
__global__ void __launch_bounds__(256)
ray_box_kernel(int64_t n, const float* __restrict__ ray_o, const float* __restrict__ ray_d, const float* __restrict__ bmin,
               const float* __restrict__ bmax, float* __restrict__ zi, float* __restrict__ zo, int32_t* __restrict__ hit) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float tn = 0.f, tf = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float inv = 1.f / ray_d[3 * i + c];
    const float o = ray_o[3 * i + c];
    const float tmin = ((bmin ? bmin[3 * i + c] : -1.f) - o) * inv;
    const float tmax = ((bmax ? bmax[3 * i + c] : 1.f) - o) * inv;
    const float t0 = nmin(tmin, tmax), t1 = nmax(tmin, tmax);
    tn = (c == 0) ? t0 : nmax(tn, t0);
    tf = (c == 0) ? t1 : nmin(tf, t1);
  }
  int h = (tf > tn) ? 1 : 0;
  h *= ((tf * (float)h) > 0.f) ? 1 : 0;
  hit[i] = h;
  zi[i] = tn * (float)h;
  zo[i] = tf * (float)h;
}

// Warp-aggregated compaction of the rays with hit[:,k] != 0 (unordered: results are scattered back per ray).
// k < 0: one object per blockIdx.y, lists [K,B] and counts [K] (durf_compact_hits_all)
__global__ void compact_hits_kernel(int B, int K, int k, const int32_t* __restrict__ hit,
                                    int32_t* __restrict__ ray_index, int32_t* __restrict__ count) {
  if (k < 0) {
    k = blockIdx.y;
    ray_index += (size_t)k * B;
    count += k;
  }
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const bool h = (b < B) && hit[(size_t)b * K + k] != 0;
  const unsigned m = __ballot_sync(kFull, h);
  if (m == 0) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(count, __popc(m));
  base = __shfl_sync(kFull, base, __ffs(m) - 1);
  if (h) ray_index[base + __popc(m & ((1u << lane) - 1))] = b;
}

// ---- backward: dL/d(origins_s, dirs_s) -> dL/d box[k, 0:6] ------------------------------------------
// o_o = R (o - p),  d_o = R d / |R d|,  R = I + a S + b S^2 with a = sin(th)/th, b = (1-cos th)/th^2.
// One thread per ray; per-block shared reduction, then 6 atomics per object per block.
__global__ void __launch_bounds__(256)
obb_frontend_bwd_kernel(int B, int K, const float* __restrict__ origins, const float* __restrict__ dirs,
                        const float* __restrict__ box, const int32_t* __restrict__ hit,
                        const float* __restrict__ d_os, const float* __restrict__ d_ds,
                        int pose_grad, int rot_grad, float* __restrict__ d_box) {
  __shared__ BoxFrame frames[kMaxObjects];
  __shared__ float s_acc[kMaxObjects * 6];
  for (int k = threadIdx.x; k < K; k += blockDim.x) make_frame(box + 6 * k, frames[k]);
  for (int i = threadIdx.x; i < K * 6; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    const float o[3] = {origins[3 * b], origins[3 * b + 1], origins[3 * b + 2]};
    const float d[3] = {dirs[3 * b], dirs[3 * b + 1], dirs[3 * b + 2]};
    const float go[3] = {d_os[3 * b], d_os[3 * b + 1], d_os[3 * b + 2]};
    const float gd[3] = {d_ds[3 * b], d_ds[3 * b + 1], d_ds[3 * b + 2]};
    for (int k = 0; k < K; ++k) {
      if (hit[(size_t)b * K + k] == 0) continue;
      const BoxFrame& f = frames[k];
      const float* aa = box + 6 * k + 3;
      const float p[3] = {box[6 * k], box[6 * k + 1], box[6 * k + 2]};
      // dL/dR from both paths.  o_o = R q with q = o - p.
      float q[3] = {o[0] - p[0], o[1] - p[1], o[2] - p[2]};
      float u[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) u[i] = f.R[3 * i] * d[0] + f.R[3 * i + 1] * d[1] + f.R[3 * i + 2] * d[2];
      const float nrm = sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
      const float un[3] = {u[0] / nrm, u[1] / nrm, u[2] / nrm};
      const float dot = gd[0] * un[0] + gd[1] * un[1] + gd[2] * un[2];
      float gu[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) gu[i] = (gd[i] - dot * un[i]) / nrm;    // d(u/|u|)
      float gR[9];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) gR[3 * i + j] = go[i] * q[j] + gu[i] * d[j];
      if (pose_grad) {
        // d o_o / d p = -R  ->  g_p = -R^T go
#pragma unroll
        for (int j = 0; j < 3; ++j)
          atomicAdd(&s_acc[6 * k + j], -(f.R[j] * go[0] + f.R[3 + j] * go[1] + f.R[6 + j] * go[2]));
      }
      if (rot_grad) {
        const float x = aa[0], y = aa[1], z = aa[2];
        float sq = x * x + y * y + z * z;
        const bool clamped = sq < 1e-12f;
        sq = clamped ? 1e-12f : sq;
        const float th = sqrtf(sq) + 1e-12f;
        const float sn = sinf(th), cs = cosf(th);
        const float a = sn / th, bb = (1.f - cs) / (th * th);
        const float da = (cs * th - sn) / (th * th);                       // d a / d th
        const float db = (sn * th - 2.f * (1.f - cs)) / (th * th * th);    // d b / d th
        const float S[9] = {0.f, -z, y, z, 0.f, -x, -y, x, 0.f};
        float S2[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) S2[3 * i + j] = S[3 * i] * S[j] + S[3 * i + 1] * S[3 + j] + S[3 * i + 2] * S[6 + j];
        float g_a = 0.f, g_b = 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i) { g_a += gR[i] * S[i]; g_b += gR[i] * S2[i]; }
        // dL/dS = a gR + b (gR S^T + S^T gR)
        float gS[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            float t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int m = 0; m < 3; ++m) { t1 += gR[3 * i + m] * S[3 * j + m]; t2 += S[3 * m + i] * gR[3 * m + j]; }
            gS[3 * i + j] = a * gR[3 * i + j] + bb * (t1 + t2);
          }
        const float g_th = g_a * da + g_b * db;
        const float dth[3] = {clamped ? 0.f : x / sqrtf(sq), clamped ? 0.f : y / sqrtf(sq), clamped ? 0.f : z / sqrtf(sq)};
        const float gx = (gS[7] - gS[5]) + g_th * dth[0];
        const float gy = (gS[2] - gS[6]) + g_th * dth[1];
        const float gz = (gS[3] - gS[1]) + g_th * dth[2];
        atomicAdd(&s_acc[6 * k + 3], gx);
        atomicAdd(&s_acc[6 * k + 4], gy);
        atomicAdd(&s_acc[6 * k + 5], gz);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * 6; i += blockDim.x)
    if (s_acc[i] != 0.f) atomicAdd(&d_box[i], s_acc[i]);
}

}  // namespace durf

using namespace durf;

extern "C" int durf_aa2matrix_fwd(durf_stream_t stream, int32_t K, const float* angles, float* R) {
  DURF_REQUIRE(K >= 0 && angles && R, DURF_E_INVALID, "durf_aa2matrix_fwd: null argument");
  if (K == 0) return DURF_OK;
  aa2matrix_kernel<<<ceil_div(K, 64), 64, 0, (cudaStream_t)stream>>>(K, angles, R);
  DURF_CHECK_LAUNCH("durf_aa2matrix_fwd");
  return DURF_OK;
}

extern "C" int durf_world2object_fwd(durf_stream_t stream, int32_t B, int32_t K, const float* pts, const float* dirs,
                                     const float* pose, int32_t pose_per_ray, const float* rot, int32_t rot_per_ray,
                                     float* pts_o, float* dirs_o) {
  DURF_REQUIRE(B >= 0 && K >= 1, DURF_E_INVALID, "durf_world2object_fwd: bad shape B=%d K=%d", B, K);
  if (B == 0) return DURF_OK;
  DURF_REQUIRE(pts && dirs && pose && rot && pts_o && dirs_o, DURF_E_INVALID, "durf_world2object_fwd: null argument");
  world2object_kernel<<<ceil_div((int64_t)B * K, 256), 256, 0, (cudaStream_t)stream>>>(
      B, K, pts, dirs, pose, pose_per_ray ? (int64_t)K * 3 : 0, rot, rot_per_ray ? (int64_t)K * 9 : 0, pts_o, dirs_o);
  DURF_CHECK_LAUNCH("durf_world2object_fwd");
  return DURF_OK;
}

extern "C" int durf_ray_box_intersection_fwd(durf_stream_t stream, int64_t n, const float* ray_o, const float* ray_d,
                                             const float* aabb_min, const float* aabb_max, float* z_in, float* z_out,
                                             int32_t* intersection) {
  DURF_REQUIRE(n >= 0, DURF_E_INVALID, "durf_ray_box_intersection_fwd: negative size");
  if (n == 0) return DURF_OK;
  DURF_REQUIRE(ray_o && ray_d && z_in && z_out && intersection, DURF_E_INVALID, "durf_ray_box_intersection_fwd: null argument");
  ray_box_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(n, ray_o, ray_d, aabb_min, aabb_max, z_in, z_out, intersection);
  DURF_CHECK_LAUNCH("durf_ray_box_intersection_fwd");
  return DURF_OK;
}

extern "C" int durf_obb_frontend_fwd(durf_stream_t stream, int32_t B, int32_t K, const float* origins,
                                     const float* directions, const float* box, const float* ext,
                                     float* origins_s, float* dirs_s, int32_t* hit, float* zi, float* zo,
                                     float* zo_ret, float* nhit, float* origins_o, float* dirs_o) {
  DURF_REQUIRE(B >= 0 && K >= 1 && K <= kMaxObjects, DURF_E_INVALID,
               "durf_obb_frontend_fwd: need 1 <= K <= %d objects (got %d)", kMaxObjects, K);
  if (B == 0) return DURF_OK;
  DURF_REQUIRE(origins && directions && box && ext && origins_s && dirs_s && hit && zi && zo && zo_ret && nhit,
               DURF_E_INVALID, "durf_obb_frontend_fwd: null argument");
  DURF_REQUIRE((origins_o == nullptr) == (dirs_o == nullptr), DURF_E_INVALID,
               "durf_obb_frontend_fwd: origins_o and dirs_o must be given together");
  if (B == 0) return DURF_OK;
  obb_frontend_kernel<<<ceil_div(B, 256), 256, 0, (cudaStream_t)stream>>>(
      B, K, origins, directions, box, ext, origins_s, dirs_s, hit, zi, zo, zo_ret, nhit, origins_o, dirs_o);
  DURF_CHECK_LAUNCH("durf_obb_frontend_fwd");
  return DURF_OK;
}

extern "C" int durf_obb_frontend_bwd(durf_stream_t stream, int32_t B, int32_t K, const float* origins,
                                     const float* directions, const float* box, const int32_t* hit,
                                     const float* d_origins_s, const float* d_dirs_s, int32_t pose_grad,
                                     int32_t rot_grad, float* d_box) {
  DURF_REQUIRE(B >= 0 && K >= 1 && K <= kMaxObjects, DURF_E_INVALID, "durf_obb_frontend_bwd: bad K=%d", K);
  if (B == 0) return DURF_OK;
  DURF_REQUIRE(origins && directions && box && hit && d_origins_s && d_dirs_s && d_box, DURF_E_INVALID,
               "durf_obb_frontend_bwd: null argument");
  if (B == 0 || (!pose_grad && !rot_grad)) return DURF_OK;
  obb_frontend_bwd_kernel<<<ceil_div(B, 256), 256, 0, (cudaStream_t)stream>>>(
      B, K, origins, directions, box, hit, d_origins_s, d_dirs_s, pose_grad, rot_grad, d_box);
  DURF_CHECK_LAUNCH("durf_obb_frontend_bwd");
  return DURF_OK;
}

extern "C" int durf_compact_hits(durf_stream_t stream, int32_t B, int32_t K, int32_t k, const int32_t* hit,
                                 int32_t* ray_index, int32_t* count) {
  DURF_REQUIRE(B >= 0 && K >= 1 && k >= 0 && k < K && count && (B == 0 || (hit && ray_index)), DURF_E_INVALID,
               "durf_compact_hits: bad argument");
  cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int32_t), (cudaStream_t)stream);
  DURF_REQUIRE(e == cudaSuccess, DURF_E_LAUNCH, "durf_compact_hits: memset: %s", cudaGetErrorString(e));
  if (B == 0) return DURF_OK;
  compact_hits_kernel<<<ceil_div(B, 256), 256, 0, (cudaStream_t)stream>>>(B, K, k, hit, ray_index, count);
  DURF_CHECK_LAUNCH("durf_compact_hits");
  return DURF_OK;
}

extern "C" int durf_compact_hits_all(durf_stream_t stream, int32_t B, int32_t K, const int32_t* hit, int32_t* ray_index,
                                     int32_t* count) {
  DURF_REQUIRE(B >= 0 && K >= 1 && K <= kMaxObjects && count && (B == 0 || (hit && ray_index)), DURF_E_INVALID,
               "durf_compact_hits_all: bad argument");
  cudaError_t e = cudaMemsetAsync(count, 0, K * sizeof(int32_t), (cudaStream_t)stream);
  DURF_REQUIRE(e == cudaSuccess, DURF_E_LAUNCH, "durf_compact_hits_all: memset: %s", cudaGetErrorString(e));
  if (B == 0) return DURF_OK;
  compact_hits_kernel<<<dim3(ceil_div(B, 256), K), 256, 0, (cudaStream_t)stream>>>(B, K, -1, hit, ray_index, count);
  DURF_CHECK_LAUNCH("durf_compact_hits_all");
  return DURF_OK;
}

// raw[ray_index[m]] += src[m]: the scatter half of an object network evaluated into compact rows (DurfMlpArgs.accumulate == 2).
// One thread per sample; the rows of one ray are contiguous, so a warp moves 128 B of density and 384 B of colour per step.
__global__ void __launch_bounds__(256)
merge_raw_kernel(int M, int N, const int32_t* __restrict__ ray_index, const int32_t* __restrict__ count,
                 const float* __restrict__ src_rgb, const float* __restrict__ src_density, float* __restrict__ raw_rgb,
                 float* __restrict__ raw_density) {
  const int rows = count ? min(*count, M) : M;
  const int64_t total = (int64_t)rows * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / N);
    const int64_t o = (int64_t)ray_index[m] * N + (i - (int64_t)m * N);
    raw_density[o] += src_density[i];
    raw_rgb[3 * o + 0] += src_rgb[3 * i + 0];
    raw_rgb[3 * o + 1] += src_rgb[3 * i + 1];
    raw_rgb[3 * o + 2] += src_rgb[3 * i + 2];
  }
}

extern "C" int durf_mlp_merge_raw(durf_stream_t stream, int32_t M, int32_t N, const int32_t* ray_index, const int32_t* count,
                                  const float* src_rgb, const float* src_density, float* raw_rgb, float* raw_density) {
  DURF_REQUIRE(M >= 0 && N >= 1, DURF_E_INVALID, "durf_mlp_merge_raw: bad shape M=%d N=%d", M, N);
  if (M == 0) return DURF_OK;
  DURF_REQUIRE(ray_index && src_rgb && src_density && raw_rgb && raw_density, DURF_E_INVALID, "durf_mlp_merge_raw: null argument");
  const int blocks = (int)std::min<int64_t>(((int64_t)M * N + 255) / 256, 148 * 8);
  merge_raw_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(M, N, ray_index, count, src_rgb, src_density, raw_rgb, raw_density);
  DURF_CHECK_LAUNCH("durf_mlp_merge_raw");
  return DURF_OK;
}
