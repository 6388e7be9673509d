// KL -- the loss block of train_boxpose.py:94-220, forward value and gradient in one pass per level.
// RGB MSE, URF LIDAR depth loss, line-of-sight "near" and "empty" losses, sky loss and the distortion loss.
// One warp per ray, 4 consecutive samples per lane.  The reference's dense [B,N,N] distortion tensor
// (train_boxpose.py:146-151) becomes two prefix sums: for sorted s,
//   sum_ij w_i w_j |s_i - s_j| = 2 sum_i w_i (s_i W_<i - WS_<i),   d/dw_k = 2 (s_k W_<k - WS_<k + WS_>k - s_k W_>k).
// HBM-bound: reads weights + t_vals (1 KB/ray), writes d_weights (512 B/ray).
#include "common.cuh"

namespace durf {

constexpr int kQ = 4;

__device__ __forceinline__ void atomic_max_pos(float* addr, float v) {   // v >= 0
  atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
}

// Pass 1: depth_mask update (train_boxpose.py:98, 138-140), normalisers and the global max of the
// line-of-sight Gaussian (`distr.max()`, :164).  norms = {sum lossmult, sum depth_mask, sum sky_mask, max distr}.
__global__ void __launch_bounds__(128)
losses_prepare_kernel(const DurfLossArgs a, float* __restrict__ norms) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ray = blockIdx.x * 4 + warp;
  float s_lm = 0.f, s_dm = 0.f, s_sky = 0.f, mx = 0.f;
  const float eps = a.eps_dev ? *a.eps_dev : a.eps;
  if (ray < a.B) {
    const float z = a.depth_gt[ray];
    // lane 0 owns the read-modify-write of depth_mask[ray]; the other lanes get the value by shuffle
    float dm = 0.f;
    if (lane == 0) dm = (a.level == 0) ? ((z > 0.f) ? 1.f : 0.f) : a.depth_mask[ray];
    dm = __shfl_sync(kFull, dm, 0);
    const float box_mask = (z < a.zo[ray]) ? 1.f : 0.f;
    dm = dm + a.box_loss_mult * a.dyn_mask[ray] * box_mask;
    const float d0 = (z > 0.f) ? 1.f : 0.f;
    float sm = (a.sky[ray] > 0.f) ? 1.f : 0.f;
    sm = sm - d0 * sm;
    if (lane == 0) { a.depth_mask[ray] = dm; s_lm = a.lossmult[ray]; s_dm = dm; s_sky = sm; }
    const float sigma = (eps / 3.f) * (eps / 3.f);
    const float c = 1.f / (sigma * sqrtf(2.f * 3.14159265358979324f));
    for (int n = lane; n < a.N; n += 32) {
      const float t = a.t_vals[(size_t)ray * (a.N + 1) + n];
      const float near = ((t > z - eps) && (t < z + eps)) ? dm : 0.f;
      const float dist = near * (t - z);
      mx = fmaxf(mx, c * expf(-(dist * dist / (2.f * sigma * sigma))));
    }
  }
  mx = warp_max(mx);
  __shared__ float red[4][4];
  if (lane == 0) { red[warp][0] = s_lm; red[warp][1] = s_dm; red[warp][2] = s_sky; red[warp][3] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
    for (int i = 0; i < 4; ++i) { t0 += red[i][0]; t1 += red[i][1]; t2 += red[i][2]; t3 = fmaxf(t3, red[i][3]); }
    atomicAdd(&norms[0], t0); atomicAdd(&norms[1], t1); atomicAdd(&norms[2], t2); atomic_max_pos(&norms[3], t3);
  }
}

__global__ void __launch_bounds__(128)
losses_fwd_bwd_kernel(const DurfLossArgs a, const float* __restrict__ norms) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ray = blockIdx.x * 4 + warp;
  const int N = a.N;
  const bool fine = a.level == a.num_levels - 1;
  const bool vec_ok = (reinterpret_cast<uintptr_t>(a.d_weights) & 15u) == 0;   // float4 store needs 16-byte alignment
  float p_rgb = 0.f, p_d = 0.f, p_n = 0.f, p_e = 0.f, p_s = 0.f, p_dist = 0.f, p_obj = 0.f, p_dyn = 0.f;
  const float eps = a.eps_dev ? *a.eps_dev : a.eps;
  if (ray < a.B) {
    const float n_lm = norms[0];
    const float n_d = fmaxf(norms[1], 1.f);
    const float n_s = fmaxf(norms[2], 1.f);
    const float gmax = norms[3];
    const float z = a.depth_gt[ray];
    const float dm = a.depth_mask[ray];
    const float dep = a.depth[ray];
    // --- per-ray terms (lane 0 writes) ---
    {
      const float box_mask = (z < a.zo[ray]) ? 1.f : 0.f;
      const float dyn = a.dyn_mask[ray];
      const float wr = a.lossmult[ray] + a.box_loss_mult * dyn * box_mask;
      p_dyn = dyn;
      const float lam_rgb = fine ? 1.f : a.coarse_loss_mult;
      float g[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float diff = a.comp_rgb[3 * ray + c] - a.pixels[3 * ray + c];
        p_rgb += wr * (diff * diff);
        p_obj += dyn * (diff * diff);                       // obj_losses numerator (train_boxpose.py:192), logging only
        g[c] = lam_rgb * 2.f * wr * diff / n_lm;
      }
      const float lam_d = a.depth_loss_mult * (fine ? 1.f : 0.1f);
      const float dd = dep - z;
      p_d = dm * (dd * dd);
      float gdep = lam_d * 2.f * dm * dd / n_d;
      const float d0 = (z > 0.f) ? 1.f : 0.f;
      float sm = (a.sky[ray] > 0.f) ? 1.f : 0.f;
      sm = sm - d0 * sm;
      const float inner = sm * dep;
      const float clamped = fmaxf(inner, 1.f);
      const float sky_depth = sm * (1.f - (1.f / clamped));
      const float sd = sky_depth - a.sky[ray];
      p_s = sm * (sd * sd);
      const float lam_s = a.sky_loss_mult * (fine ? 10.f : 1.f);
      if (inner > 1.f) gdep += lam_s * 2.f * sm * sd / n_s * (sm * (1.f / (clamped * clamped)) * sm);
      if (lane == 0) {
        a.d_comp_rgb[3 * ray] = g[0]; a.d_comp_rgb[3 * ray + 1] = g[1]; a.d_comp_rgb[3 * ray + 2] = g[2];
        a.d_depth[ray] = gdep;
      } else { p_rgb = 0.f; p_d = 0.f; p_s = 0.f; p_obj = 0.f; p_dyn = 0.f; }
    }
    // --- per-sample terms ---
    const int n0 = lane * kQ;
    float w[kQ], s[kQ], dl[kQ], tv[kQ + 1];
#pragma unroll
    for (int q = 0; q <= kQ; ++q) tv[q] = (n0 + q <= N) ? a.t_vals[(size_t)ray * (N + 1) + n0 + q] : 0.f;
    float sw = 0.f, sws = 0.f;
#pragma unroll
    for (int q = 0; q < kQ; ++q) {
      const bool ok = n0 + q < N;
      w[q] = ok ? a.weights[(size_t)ray * N + n0 + q] : 0.f;
      s[q] = 0.5f * (tv[q] + tv[q + 1]);
      dl[q] = tv[q + 1] - tv[q];
      sw += w[q];
      sws += w[q] * s[q];
    }
    float Wlt = warp_scan_excl(sw, lane), WSlt = warp_scan_excl(sws, lane);
    float Wgt = warp_rscan_excl(sw, lane), WSgt = warp_rscan_excl(sws, lane);
    // suffix sums inside the lane
    float wsuf[kQ], wssuf[kQ];
    {
      float r0 = 0.f, r1 = 0.f;
#pragma unroll
      for (int q = kQ - 1; q >= 0; --q) { wsuf[q] = r0; wssuf[q] = r1; r0 += w[q]; r1 += w[q] * s[q]; }
    }
    const float sigma = (eps / 3.f) * (eps / 3.f);
    const float c = 1.f / (sigma * sqrtf(2.f * 3.14159265358979324f));
    const float lam_n = a.near_loss_mult * (fine ? 1.f : 0.1f);
    const float lam_e = a.empty_loss_mult * (fine ? 1.f : 0.1f);
    float gw[kQ];
#pragma unroll
    for (int q = 0; q < kQ; ++q) {
      const bool ok = n0 + q < N;
      const float t = tv[q];
      const float near = ((t > z - eps) && (t < z + eps)) ? dm : 0.f;
      const float empty = (t > z + eps) ? dm : 0.f;
      const float dist = near * (t - z);
      float g = c * expf(-(dist * dist / (2.f * sigma * sigma)));
      g = g / gmax;
      g = g * near;
      const float rn = near * w[q] - g;
      const float re = empty * w[q];
      float grad = lam_n * 2.f * rn * near / n_d + lam_e * 2.f * re * empty / n_d;
      // distortion
      const float before_w = Wlt, before_ws = WSlt;
      const float after_w = Wgt + wsuf[q], after_ws = WSgt + wssuf[q];
      const float inter = s[q] * before_w - before_ws;
      grad += a.distortion_mult * (2.f * (inter + after_ws - s[q] * after_w) + (2.f / 3.f) * w[q] * dl[q]);
      if (ok) {
        p_n += rn * rn;
        p_e += re * re;
        p_dist += 2.f * w[q] * inter + (1.f / 3.f) * (w[q] * w[q]) * dl[q];
      }
      Wlt += w[q]; WSlt += w[q] * s[q];
      gw[q] = grad;
    }
    if (N == 128 && vec_ok) *reinterpret_cast<float4*>(a.d_weights + (size_t)ray * N + n0) = make_float4(gw[0], gw[1], gw[2], gw[3]);
    else {
#pragma unroll
      for (int q = 0; q < kQ; ++q) if (n0 + q < N) a.d_weights[(size_t)ray * N + n0 + q] = gw[q];
    }
  }
  p_rgb = warp_sum(p_rgb); p_d = warp_sum(p_d); p_n = warp_sum(p_n); p_e = warp_sum(p_e); p_s = warp_sum(p_s); p_dist = warp_sum(p_dist);
  p_obj = warp_sum(p_obj); p_dyn = warp_sum(p_dyn);
  __shared__ float red[4][DURF_LP_STRIDE];
  if (lane == 0) {
    red[warp][0] = p_rgb; red[warp][1] = p_d; red[warp][2] = p_n; red[warp][3] = p_e; red[warp][4] = p_s; red[warp][5] = p_dist;
    red[warp][6] = p_obj; red[warp][7] = p_dyn;
  }
  __syncthreads();
  // Deterministic reduction: every block stores its 8 partial sums; the last block to finish (ticket) adds all of them
  // in a fixed order.  (Float atomics gave a run-to-run different loss VALUE; the gradients never depended on it.)
  float* bp = a.reduce_ws;
  unsigned* ticket = reinterpret_cast<unsigned*>(a.reduce_ws + (size_t)gridDim.x * DURF_LP_STRIDE);
  if (threadIdx.x < DURF_LP_STRIDE)
    bp[(size_t)blockIdx.x * DURF_LP_STRIDE + threadIdx.x] =
        red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
  __threadfence();
  __syncthreads();
  __shared__ unsigned last;
  if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (last) {
    __threadfence();
    // 128 threads: slot = tid & 7, 16 strided lanes per slot, then a fixed-order tree over the 16
    const int slot = threadIdx.x & 7, part = threadIdx.x >> 3;
    float acc = 0.f;
    for (unsigned b = part; b < gridDim.x; b += 16) acc += __ldcg(&bp[(size_t)b * DURF_LP_STRIDE + slot]);
    __shared__ float tree[16][DURF_LP_STRIDE];
    tree[part][slot] = acc;
    __syncthreads();
    if (threadIdx.x < DURF_LP_STRIDE) {
      float t = 0.f;
      for (int i = 0; i < 16; ++i) t += tree[i][threadIdx.x];
      a.partials[a.level * DURF_LP_STRIDE + threadIdx.x] = t;
      if (threadIdx.x == 0) *ticket = 0u;                    // ready for the next launch
    }
  }
}

// The scalar tail of loss_fn (train_boxpose.py:196-220): normalise the per-level sums and combine them with the level
// weights (fine x1, coarse x0.1 / coarse_loss_mult, sky fine x10), in the reference's order of additions.  One thread.
__global__ void losses_finalize_kernel(const DurfLossFinalizeArgs a) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int L = a.num_levels;
  float lo[2], dl[2], nl[2], el[2], sl[2], di[2], tv[2];
  for (int i = 0; i < L; ++i) {
    const float* p = a.partials + i * DURF_LP_STRIDE;
    const float* n = a.norms + i * 4;
    const float nd = fmaxf(n[1], 1.f);
    lo[i] = p[DURF_LP_RGB] / n[0];
    dl[i] = p[DURF_LP_DEPTH] / nd;
    nl[i] = p[DURF_LP_NEAR] / nd;
    el[i] = p[DURF_LP_EMPTY] / nd;
    sl[i] = p[DURF_LP_SKY] / fmaxf(n[2], 1.f);
    di[i] = p[DURF_LP_DISTR];
    tv[i] = a.tv ? a.tv[i] : 0.f;
    float* o = a.stats + i * DURF_LS_STRIDE;
    o[0] = lo[i]; o[1] = dl[i]; o[2] = nl[i]; o[3] = el[i]; o[4] = sl[i]; o[5] = di[i];
    o[6] = p[DURF_LP_OBJ] / p[DURF_LP_DYN];          // 0/0 = NaN when no ray hits a box, like the reference
    o[7] = tv[i];
  }
  const int f = L - 1;
  auto coarse = [&](const float* x) { float t = 0.f; for (int i = 0; i < f; ++i) t += x[i]; return t; };
  const float wl2 = a.weight_l2 ? *a.weight_l2 : 0.f;
  float loss = (a.coarse_loss_mult * coarse(lo) + lo[f]) + wl2;
  loss += a.sky_loss_mult * coarse(sl) + 10.0f * a.sky_loss_mult * sl[f];
  loss += a.depth_loss_mult * dl[f] + 0.1f * a.depth_loss_mult * coarse(dl);
  loss += a.near_loss_mult * nl[f] + 0.1f * a.near_loss_mult * coarse(nl);
  loss += a.empty_loss_mult * el[f] + 0.1f * a.empty_loss_mult * coarse(el);
  loss += a.tv_loss_mult * tv[f] + 0.1f * a.tv_loss_mult * coarse(tv);
  loss += a.distortion_mult * di[f] + a.distortion_mult * coarse(di);
  float* o = a.stats + L * DURF_LS_STRIDE;
  o[0] = loss; o[1] = wl2;
}

static int check(const DurfLossArgs* a, const char* who) {
  DURF_REQUIRE(a != nullptr, DURF_E_INVALID, "%s: null args", who);
  DURF_REQUIRE(a->B >= 0 && a->N >= 1 && a->N <= 128, DURF_E_INVALID, "%s: need 1 <= N <= 128", who);
  DURF_REQUIRE(a->level >= 0 && a->level < a->num_levels && a->num_levels <= 2, DURF_E_INVALID, "%s: bad level %d/%d", who,
               a->level, a->num_levels);
  DURF_REQUIRE(a->B == 0 || (a->t_vals && a->depth_gt && a->sky && a->lossmult && a->dyn_mask && a->zo && a->depth_mask), DURF_E_INVALID,
               "%s: null input", who);
  return DURF_OK;
}

}  // namespace durf

using namespace durf;

extern "C" int durf_losses_finalize(durf_stream_t stream, const DurfLossFinalizeArgs* a) {
  DURF_REQUIRE(a && a->num_levels >= 1 && a->num_levels <= 2 && a->partials && a->norms && a->stats, DURF_E_INVALID,
               "durf_losses_finalize: bad argument");
  losses_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*a);
  DURF_CHECK_LAUNCH("durf_losses_finalize");
  return DURF_OK;
}

extern "C" int64_t durf_losses_reduce_ws_floats(int32_t B) {
  return (int64_t)ceil_div(B > 0 ? B : 1, 4) * DURF_LP_STRIDE + 4;
}

extern "C" int durf_losses_prepare(durf_stream_t stream, const DurfLossArgs* args, float* norms) {
  int rc = check(args, "durf_losses_prepare");
  if (rc != DURF_OK) return rc;
  if (args->B == 0) return DURF_OK;
  DURF_REQUIRE(norms, DURF_E_INVALID, "durf_losses_prepare: null norms");
  losses_prepare_kernel<<<ceil_div(args->B, 4), 128, 0, (cudaStream_t)stream>>>(*args, norms);
  DURF_CHECK_LAUNCH("durf_losses_prepare");
  return DURF_OK;
}

extern "C" int durf_losses_fwd_bwd(durf_stream_t stream, const DurfLossArgs* args, const float* norms) {
  int rc = check(args, "durf_losses_fwd_bwd");
  if (rc != DURF_OK) return rc;
  if (args->B == 0) return DURF_OK;
  DURF_REQUIRE(norms && args->comp_rgb && args->depth && args->weights && args->pixels && args->partials && args->d_comp_rgb &&
                   args->d_depth && args->d_weights && args->reduce_ws, DURF_E_INVALID, "durf_losses_fwd_bwd: null buffer");
  if (args->B == 0) return DURF_OK;
  losses_fwd_bwd_kernel<<<ceil_div(args->B, 4), 128, 0, (cudaStream_t)stream>>>(*args, norms);
  DURF_CHECK_LAUNCH("durf_losses_fwd_bwd");
  return DURF_OK;
}
